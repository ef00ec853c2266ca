"""The hot path's kernel SOURCE (preprocess, depth sort, scan, emit, tile sort, tile ranges, forward blend, backward blend,
chain rule -- the same .cu files that are compiled for sm_100a) executed on the CPU through tests/host_emul and compared
with the oracle exactly like the -m gpu parity tests, at sizes the emulation finishes in seconds.  Same tolerances:
forward 1e-5 abs, gradients 1e-4 rel.  Differences to a GPU run: ex2.approx / rcp.approx / sqrt.approx are their libm
counterparts here, and float atomics arrive in another order.  Not a product path (see tests/host_emul/cuda_shim.h)."""
import math

import numpy as np
import pytest
import torch

import gs_test_util as U
from oracle import torch_oracle as TO


def _weights(cam, seed=5):
    gen = torch.Generator().manual_seed(seed)
    H, W = cam.image_height, cam.image_width
    return torch.randn(3, H, W, generator=gen).numpy(), torch.randn(1, H, W, generator=gen).numpy()


def _check(scene, cam, mode, with_depth_grad=True, grad_tol=U.GRAD_REL_TOL):
    args = U.make_args(scene, mode)
    wc, wd = _weights(cam)
    if not with_depth_grad:
        wd = None
    got = U.run_cuda(args, cam, wc, wd, device="cpu")
    ref = U.run_oracle(args, cam, wc, wd)
    assert (got["radii"] == ref["radii"]).all(), "radii differ"
    U.assert_image_close(got["color"], ref["color"], "color")
    U.assert_image_close(got["invdepth"], ref["invdepth"], "invdepth")
    U.assert_grads_close(got["grads"], ref["grads"], tol=grad_tol,
                         flips=U.count_flips(got["color"], ref["color"]) + U.count_flips(got["invdepth"], ref["invdepth"]))
    return got, ref


@pytest.mark.parametrize("deg", [0, 3])
def test_sh_degrees(on_host, deg):
    scene = TO.make_scene(250, seed=10 + deg, log_scale_mean=-2.4)
    _check(scene, TO.make_camera(48, 32, sh_degree=deg, bg=(0.1, 0.3, 0.6)), "sh")


@pytest.mark.parametrize("mode", ["sh", "precomp"])
def test_antialiasing_and_precomputed_inputs(on_host, mode):
    scene = TO.make_scene(250, seed=21, log_scale_mean=-2.6)
    _check(scene, TO.make_camera(56, 40, sh_degree=3, antialiasing=True, bg=(1.0, 1.0, 1.0)), mode)


def test_ragged_image_scale_modifier_no_depth_gradient(on_host):
    scene = TO.make_scene(200, seed=33, log_scale_mean=-2.4)
    _check(scene, TO.make_camera(70, 41, sh_degree=2, scale_modifier=0.7), "sh", with_depth_grad=False)


def test_large_and_anisotropic_gaussians(on_host):
    """Gaussians spanning many tiles: the exact tile culling, the patch masks and several staging rounds per tile."""
    scene = TO.make_scene(90, seed=35, log_scale_mean=-1.2, log_scale_std=1.2)
    # strongly anisotropic footprints: the float32 oracle's own gradients are 1-2e-4 from float64 here, and the order in which
    # the float atomics land (not reproducible) moves the kernels' result by as much -- hence 5e-4 instead of 1e-4
    _check(scene, TO.make_camera(80, 48, sh_degree=3, bg=(0.5, 0.5, 0.5)), "sh", grad_tol=5e-4)


def test_many_gaussians_per_tile(on_host):
    """More than one 128-record staging round in most tiles."""
    scene = TO.make_scene(480, seed=36, log_scale_mean=-2.0)
    scene["opacities"] = scene["opacities"] * 0.15                      # keep transmittance alive deep into the lists
    _check(scene, TO.make_camera(48, 32, sh_degree=1), "sh")


def test_empty_and_all_culled(on_host):
    dgr = on_host
    cam = TO.make_camera(48, 32, sh_degree=0, bg=(0.2, 0.4, 0.6))
    rast = dgr.GaussianRasterizer(U.settings_to(cam, "cpu"))
    z = lambda *s: torch.zeros(*s)
    color, radii, invd = rast(means3D=z(0, 3), means2D=z(0, 3), opacities=z(0, 1), shs=z(0, 1, 3), scales=z(0, 3), rotations=z(0, 4))
    assert radii.numel() == 0 and torch.allclose(color, cam.bg[:, None, None].expand(3, 32, 48))
    scene = TO.make_scene(60, seed=1, sh_coeffs=1)
    scene["means3D"][:, 2] -= 100.0
    got = U.run_cuda(U.make_args(scene, "sh"), cam, *_weights(cam), device="cpu")
    assert (got["radii"] == 0).all() and all(np.all(v == 0) for v in got["grads"].values() if v is not None)


def test_view_batch_path_matches_per_view_autograd(on_host):
    """gsb_forward_batch / gsb_backward_batch (parameters read once for all views, batched sorts and scans, gradients summed in
    place) against per-view render() + autograd, both on the host build."""
    import bench
    from gaussian_renderer import GradientBucket, render, render_views_backward
    dev = "cpu"
    scene = TO.make_scene(300, seed=51, log_scale_mean=-2.6)
    W, H = 48, 32
    cams = [bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(i, 3.0), dev) for i in range(2)]
    gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(i)) for i in range(2)]
    bg = torch.tensor([0.1, 0.2, 0.3])

    def run(fused):
        pc = bench.BenchGaussians(scene, 3, dev)
        bucket = GradientBucket(pc.parameters())
        if fused:
            losses = render_views_backward(cams, pc, bench.Pipe(), bg, lambda img, d, i: (img - gts[i]).abs().mean() + 0.1 * d.mean())["losses"]
        else:
            ls = []
            for i, cam in enumerate(cams):
                pkg = render(cam, pc, bench.Pipe(), bg)
                loss = (pkg["render"] - gts[i]).abs().mean() + 0.1 * pkg["depth"].mean()
                loss.backward()
                ls.append(loss.detach())
            losses = torch.stack(ls)
        return losses.numpy(), bucket.flat.numpy().copy()

    l0, g0 = run(False)
    l1, g1 = run(True)
    assert np.allclose(l0, l1, rtol=1e-6, atol=1e-7)
    assert np.abs(g0 - g1).max() <= 1e-4 * np.abs(g0).max()


@pytest.mark.parametrize("shape", [(3, 37, 53), (1, 16, 16)])
def test_fused_losses_source(on_host, shape, golden):
    dgr = on_host
    g = torch.Generator().manual_seed(7)
    img = (torch.rand(*shape, generator=g) * 1.4 - 0.2).requires_grad_(True)      # some values outside [0,1]: clamp mask
    gt = torch.rand(*shape, generator=g)
    ref = TO.photometric_loss(img.clamp(0, 1), gt, 0.2)
    (g_ref,) = torch.autograd.grad(ref, img)
    loss, grad, parts = dgr.photometric_loss_and_grad(img.detach(), gt, 0.2)
    assert abs(float(loss) - float(ref)) < 2e-6
    assert float((grad - g_ref).abs().max()) <= 1e-4 * float(g_ref.abs().max())
    import fused_ssim                                                              # drop-in: plain SSIM, input NOT clamped
    a = img.detach().clone().requires_grad_(True)
    v = fused_ssim.fused_ssim(a.unsqueeze(0), gt.unsqueeze(0))
    a_ref = img.detach().clone().requires_grad_(True)
    v_ref = TO.ssim(a_ref, gt)
    assert abs(float(v) - float(v_ref)) < 2e-6
    (ga,), (ga_ref,) = torch.autograd.grad(v, a), torch.autograd.grad(v_ref, a_ref)
    assert float((ga - ga_ref).abs().max()) <= 1e-4 * float(ga_ref.abs().max())
    if shape == (3, 37, 53):                                                       # the reference's own numbers
        x, y = torch.from_numpy(golden["loss_img"]), torch.from_numpy(golden["loss_gt"])
        l2, g2, p2 = dgr.photometric_loss_and_grad(x, y, 0.2)
        assert abs(float(l2) - float(golden["loss_total"])) < 2e-6 and abs(float(p2[2]) - float(golden["loss_ssim"])) < 2e-6
        assert np.abs(g2.numpy() - golden["loss_grad"]).max() <= 1e-4 * np.abs(golden["loss_grad"]).max()
    n = img.numel() // 4 * 4
    l1, g1 = dgr.l1_loss_and_grad(img.detach().reshape(-1)[:n].contiguous(), gt.reshape(-1)[:n].contiguous())
    x = img.detach().reshape(-1)[:n].clone().requires_grad_(True)
    r1 = (x.clamp(0, 1) - gt.reshape(-1)[:n]).abs().mean()
    (gr1,) = torch.autograd.grad(r1, x)
    assert abs(float(l1) - float(r1)) < 1e-6 and float((g1 - gr1).abs().max()) <= 1e-6


def _edge_cases():
    s = TO.make_scene(60, seed=1, log_scale_mean=-1.8)
    ties = {k: torch.cat([v, v]) for k, v in s.items()}                       # coincident gaussians: depth ties -> index order
    ties["shs"][60:] = torch.randn(60, 16, 3, generator=torch.Generator().manual_seed(2)) * 0.4
    yield "depth_ties", ties, TO.make_camera(48, 32, sh_degree=3)
    yield "image_smaller_than_a_tile", TO.make_scene(40, seed=2, log_scale_mean=-1.5), TO.make_camera(5, 3, sh_degree=1)
    o = TO.make_scene(80, seed=4, log_scale_mean=-2.0)
    o["opacities"][:40] = 0.0
    o["opacities"][40:] = 1.0
    yield "opacity_exactly_0_and_1", o, TO.make_camera(48, 32, sh_degree=2)
    yield "screen_filling", TO.make_scene(25, seed=5, log_scale_mean=0.5, log_scale_std=0.3), TO.make_camera(64, 48, sh_degree=1)
    yield "subpixel_antialiased", TO.make_scene(300, seed=6, log_scale_mean=-7.0), TO.make_camera(48, 32, sh_degree=0, antialiasing=True)
    yield "near_plane_straddling", TO.make_scene(300, seed=10, log_scale_mean=-2.5), TO.make_camera(48, 32, sh_degree=2, eye=(0.0, 0.0, -0.25))


@pytest.mark.parametrize("name,scene,cam", list(_edge_cases()), ids=lambda v: v if isinstance(v, str) else "")
def test_edge_cases(on_host, name, scene, cam):
    _check(scene, cam, "sh")


def test_needle_gaussians_are_as_exact_as_float32_allows(on_host):
    """60:1 anisotropy: A dx^2 + C dy^2 + 2 B dx dy cancels by four orders of magnitude, so ANY float32 evaluation is ~3e-4
    away from the float64 result.  The kernels stay an order of magnitude closer to the float32 oracle than that, and the
    exact tile culling does not change a pixel."""
    dgr = on_host
    s = TO.make_scene(80, seed=8, log_scale_mean=-3.0)
    s["scales"][:, 0] *= 60.0
    cam = TO.make_camera(64, 48, sh_degree=1)
    args = U.make_args(s, "sh")
    wc, wd = _weights(cam)
    ref = U.run_oracle(args, cam, wc, wd)
    with torch.no_grad():
        c64 = TO.rasterize(args["means3D"].double(), None, args["shs"].double(), None, args["opacities"].double(),
                           args["scales"].double(), args["rotations"].double(), None, cam)[0].numpy()
    oracle_err = np.abs(ref["color"] - c64).max()
    got = U.run_cuda(args, cam, wc, wd, device="cpu")
    dgr.set_option("cull", 0)
    try:
        unculled = U.run_cuda(args, cam, wc, wd, device="cpu")
    finally:
        dgr.set_option("cull", 1)
    assert np.array_equal(got["color"], unculled["color"]) and np.array_equal(got["radii"], ref["radii"])
    assert oracle_err > 5e-5                                              # the premise: float32 itself is the limit here
    assert np.abs(got["color"] - ref["color"]).max() <= 0.2 * oracle_err
    assert np.abs(got["color"] - c64).max() <= 1.2 * oracle_err
    U.assert_grads_close(got["grads"], ref["grads"], tol=1e-2)


def _batch_step(scene, cams, gts, bg, **kw):
    import bench
    from gaussian_renderer import GradientBucket, render_views_backward
    pc = bench.BenchGaussians(scene, 3, "cpu")
    bucket = GradientBucket(pc.parameters())
    out = render_views_backward(cams, pc, bench.Pipe(), bg, lambda img, d, i: (img - gts[i]).abs().mean() + 0.1 * d.mean(),
                                keep_images=True, **kw)
    return out["losses"].numpy(), bucket.flat.numpy().copy(), [im.numpy().copy() for im in out["images"]], bucket


def _batch_inputs(n=300, seed=51, views=2, W=48, H=32):
    import bench
    scene = TO.make_scene(n, seed=seed, log_scale_mean=-2.6)
    cams = [bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(i, 3.0), "cpu") for i in range(views)]
    gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(i)) for i in range(views)]
    return scene, cams, gts, torch.tensor([0.1, 0.2, 0.3])


def test_sync_free_chunked_view_batch_step(on_host):
    """gsb_forward_batch_async + gsb_backward_batch_chunked against the synchronous call: same images and gradients; the counts
    and their running maximum stay on the device; the chunk callback sees disjoint ranges covering [0, P) in order; option
    tile_order (default on: heavy tiles first) switched OFF leaves the images unchanged; a capacity that is too small truncates
    the lists without touching memory out of bounds and is reported by AsyncCapacity.check(), which also grows the capacity."""
    from gaussian_renderer import AsyncCapacity
    dgr = on_host
    scene, cams, gts, bg = _batch_inputs()
    l0, g0, im0, _ = _batch_step(scene, cams, gts, bg, overwrite=True)
    cap = AsyncCapacity("cpu")
    cap.learn([1000, 3000])
    assert cap.capacity == 1 << 20 and cap.observed_max() == 0
    cap.capacity = 4096                          # the emulation walks every block of the capacity-sized launches: keep it small
    seen_chunks = []
    dgr.set_option("tile_order", 0)
    try:
        l2, g2, im2, _ = _batch_step(scene, cams, gts, bg, overwrite=True, capacity=cap, grad_chunks=2,
                                     on_grad_chunk=lambda c, a, b: seen_chunks.append((c, a, b)))          # no read-back
    finally:
        dgr.set_option("tile_order", 1)
    assert seen_chunks == [(0, 0, 256), (1, 256, 300)]
    for a, b in zip(im0, im2):
        assert np.array_equal(a, b)
    assert np.array_equal(l0, l2) and np.abs(g0 - g2).max() <= 1e-5 * np.abs(g0).max()
    seen = cap.observed_max()
    assert 0 < seen <= cap.capacity and cap.check()
    counts = cap.counts.numpy()
    assert counts[:2].max() == seen and counts[2:16].sum() == 0
    # overflow: lists truncated at the capacity, reported afterwards
    small = AsyncCapacity("cpu", capacity=max(64, seen // 3))
    _batch_step(scene, cams, gts, bg, capacity=small)
    assert small.observed_max() == seen
    assert not small.check() and small.capacity >= seen and small.observed_max() == 0


@pytest.mark.parametrize("deg,tma", [(1, 1), (3, 0)])
def test_sh_row_layouts_source(on_host, deg, tma):
    """The two shared-memory layouts of the SH rows (csrc/preprocess.cu): TMA bulk copies + float4 access with a 12-float row
    (SH degree 1 tensors: three 16-byte units per row; every other test of this file runs the 48-float rows), and option
    pre_tma switched OFF (128-bit / scalar staging at an odd word stride).  Single view against the oracle."""
    dgr = on_host
    scene = TO.make_scene(260, seed=40 + deg, sh_coeffs=(deg + 1) ** 2, log_scale_mean=-2.4)
    dgr.set_option("pre_tma", tma)
    try:
        _check(scene, TO.make_camera(48, 32, sh_degree=deg, bg=(0.1, 0.3, 0.6)), "sh")
    finally:
        dgr.set_option("pre_tma", 1)


def test_state_buffers_are_freed_by_refcount_not_by_the_garbage_collector(on_host):
    """The allocator callback handed to the C ABI must not keep a call's buffers in a reference cycle: with the cyclic GC off
    (bench.py's timed region; any latency-sensitive loop) every step would otherwise pin ~1 GB of forward state until the next
    collection, and the caching allocator would have to cudaMalloc its way around it."""
    import gc
    import weakref
    dgr = on_host
    scene = TO.make_scene(120, seed=3, log_scale_mean=-2.2)
    cam = U.settings_to(TO.make_camera(48, 32, sh_degree=3), "cpu")
    a = U.make_args(scene, "sh")
    gc.collect()
    gc.disable()
    try:
        _, _, _, pack = dgr._forward_impl(a["means3D"], a["shs"], None, a["opacities"].reshape(-1), a["scales"], a["rotations"], None, cam)
        refs = [weakref.ref(pack[k]) for k in ("geom", "binning", "image")]
        assert all(r() is not None for r in refs)
        del pack
        assert all(r() is None for r in refs), "forward state survived its last reference: a cycle keeps it alive"
    finally:
        gc.enable()


def test_fused_reduce_scatter_source(on_host):
    """gsb_backward_batch_peer (csrc/preprocess.cu, peer output path): two "ranks" played one after the other in this process,
    each adding the gradients of ITS views into the buffer of the rank that owns the row (rank 0: gaussians [0, 256), rank 1:
    [256, 512) of a 300-gaussian cloud -- the padding rows receive zeros).  The owned rows, put side by side, equal the gradient
    of all views computed the ordinary way.  (On the host build the bulk reduce-adds are plain loops: layout, ownership and
    panel indexing are what is checked here; NVLink and the IPC mappings by the -m gpu two-GPU test.)"""
    import bench
    from gaussian_renderer import GradientBucket, render_views_backward
    scene, cams, gts, bg = _batch_inputs(n=300, views=2)
    P, rows, world = 300, 256, 2
    width = {"means3D": 3, "opacities": 1, "scales": 3, "rotations": 4, "shs": 48}
    offset, off = {}, 0
    for k, w in width.items():
        offset[k] = off
        off += rows * world * w
    bufs = [torch.zeros(off), torch.zeros(off)]
    loss = lambda mine: (lambda img, d, i: (img - gts[mine[i]]).abs().mean() + 0.1 * d.mean())
    for r in range(world):
        mine = [v for v in range(2) if v % world == r]
        pc = bench.BenchGaussians(scene, 3, "cpu")
        named = {"means3D": pc._xyz, "shs": pc._shs, "opacities": pc._opacity, "scales": pc._scaling, "rotations": pc._rotation}
        for k, p in named.items():
            p.grad = bufs[r][offset[k]:offset[k] + rows * world * width[k]].view(rows * world, width[k])[:P].view_as(p)
        render_views_backward([cams[v] for v in mine], pc, bench.Pipe(), bg, loss(mine),
                              peers=(world, r, rows, [b.data_ptr() for b in bufs]))
    # reference: all four views, ordinary path
    pc = bench.BenchGaussians(scene, 3, "cpu")
    bucket = GradientBucket(pc.parameters())
    render_views_backward(cams, pc, bench.Pipe(), bg, loss(list(range(2))))
    ref = {"means3D": pc._xyz.grad, "shs": pc._shs.grad, "opacities": pc._opacity.grad, "scales": pc._scaling.grad, "rotations": pc._rotation.grad}
    for k, w in width.items():
        full = [b[offset[k]:offset[k] + rows * world * w].view(rows * world, w) for b in bufs]
        got = torch.cat((full[0][:rows], full[1][rows:P]))                       # owned rows side by side
        want = ref[k].reshape(P, w)
        assert float((got - want).abs().max()) <= 1e-5 * float(want.abs().max()) + 1e-12, k
        assert float(full[0][rows:].abs().max()) == 0.0 and float(full[1][:rows].abs().max()) == 0.0      # nobody touched foreign rows
        assert float(full[1][P:].abs().max()) == 0.0                                                         # padding rows: zeros


def test_loss_gradient_written_in_place(on_host):
    """A loss_fn with a ``grad_out`` parameter receives its slot of the batch's gradient buffer: same result as the copying path."""
    import bench
    from gaussian_renderer import GradientBucket, render_views_backward
    dgr = on_host
    scene, cams, gts, bg = _batch_inputs(n=200)

    def run(in_place):
        pc = bench.BenchGaussians(scene, 3, "cpu")
        bucket = GradientBucket(pc.parameters())
        seen = []

        def with_slot(img, d, i, grad_out=None):
            seen.append(grad_out is not None)
            return dgr.l1_loss_and_grad(img, gts[i], grad_out=grad_out)
        plain = lambda img, d, i: dgr.l1_loss_and_grad(img, gts[i])
        out = render_views_backward(cams, pc, bench.Pipe(), bg, with_slot if in_place else plain, loss_returns_grad=True, overwrite=True)
        assert seen == ([True, True] if in_place else [])
        return out["losses"].numpy(), bucket.flat.numpy().copy()

    (l0, g0), (l1, g1) = run(False), run(True)
    assert np.array_equal(l0, l1) and np.abs(g0 - g1).max() <= 1e-6 * np.abs(g0).max()
