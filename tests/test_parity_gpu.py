"""GPU parity tests (-m gpu): the sm_100a CUDA path, called through the drop-in package and the C ABI,
against the CPU oracle (oracle/gs_oracle.c) on the same seeded inputs.

Tolerances (BASELINE.json north_star): forward 1e-5 abs, gradients 1e-4 rel; see tests/_util.py for how
threshold flips at alpha = 1/255 are counted rather than absorbed."""
import numpy as np
import pytest
import torch

from oracle import torch_oracle as TO
import gs_test_util as U

pytestmark = pytest.mark.gpu


def _weights(cam, seed=5):
    gen = torch.Generator().manual_seed(seed)
    H, W = cam.image_height, cam.image_width
    return torch.randn(3, H, W, generator=gen).numpy(), torch.randn(1, H, W, generator=gen).numpy()


def _flips(got, ref):
    return U.count_flips(got["color"], ref["color"]) + U.count_flips(got["invdepth"], ref["invdepth"])


def _check(scene, cam, mode, with_depth_grad=True, grad_tol=U.GRAD_REL_TOL):
    args = U.make_args(scene, mode)
    wc, wd = _weights(cam)
    if not with_depth_grad:
        wd = None
    got = U.run_cuda(args, cam, wc, wd)
    ref = U.run_oracle(args, cam, wc, wd)
    assert (got["radii"] == ref["radii"]).all(), "radii differ"
    U.assert_image_close(got["color"], ref["color"], "color")
    U.assert_image_close(got["invdepth"], ref["invdepth"], "invdepth")
    U.assert_grads_close(got["grads"], ref["grads"], tol=grad_tol, flips=_flips(got, ref))
    return got, ref


@pytest.mark.parametrize("deg", [0, 1, 2, 3])
def test_sh_degrees(deg):
    scene = TO.make_scene(2000, seed=10 + deg, log_scale_mean=-2.8)
    cam = TO.make_camera(160, 96, sh_degree=deg, bg=(0.1, 0.3, 0.6))
    _check(scene, cam, "sh")


@pytest.mark.parametrize("aa", [False, True])
@pytest.mark.parametrize("mode", ["sh", "precomp"])
def test_modes_and_antialiasing(aa, mode):
    scene = TO.make_scene(3000, seed=21, log_scale_mean=-3.0)
    cam = TO.make_camera(200, 120, sh_degree=3, antialiasing=aa, bg=(1.0, 1.0, 1.0))
    _check(scene, cam, mode)


def test_ragged_image_size_and_scale_modifier():
    scene = TO.make_scene(1500, seed=33, log_scale_mean=-2.6)
    cam = TO.make_camera(131, 77, sh_degree=2, scale_modifier=0.7, bg=(0.0, 0.0, 0.0))
    _check(scene, cam, "sh")


def test_no_depth_gradient():
    scene = TO.make_scene(1000, seed=34, log_scale_mean=-2.6)
    cam = TO.make_camera(96, 64, sh_degree=1)
    _check(scene, cam, "sh", with_depth_grad=False)


def test_large_and_anisotropic_gaussians():
    """Gaussians spanning many tiles, strongly anisotropic: exercises the exact tile culling."""
    scene = TO.make_scene(400, seed=35, log_scale_mean=-1.2, log_scale_std=1.2)
    cam = TO.make_camera(256, 160, sh_degree=3, bg=(0.5, 0.5, 0.5))
    _check(scene, cam, "sh")


def test_camera_inside_the_cloud():
    """Points behind / beside the camera: near cull, frustum clamp of the Jacobian."""
    scene = TO.make_scene(4000, seed=36, log_scale_mean=-3.0)
    cam = TO.make_camera(160, 120, sh_degree=3, eye=(0.1, 0.05, -0.4))
    _check(scene, cam, "sh", grad_tol=1.1e-3)   # near-plane points: the float32 oracle itself is 1e-3 from float64 here


def test_low_opacity_never_visible():
    scene = TO.make_scene(500, seed=37, log_scale_mean=-2.5)
    scene["opacities"] = scene["opacities"] * 0.003  # below 1/255: contributes nowhere
    cam = TO.make_camera(96, 64, sh_degree=0)
    got, ref = _check(scene, cam, "sh")
    assert (got["radii"] > 0).any()
    assert np.abs(got["color"]).max() == 0.0


def test_empty_input_and_all_culled():
    import diff_gaussian_rasterization as dgr
    cam = TO.make_camera(64, 48, sh_degree=0, bg=(0.2, 0.4, 0.6))
    rs = U.settings_to(cam, "cuda")
    rast = dgr.GaussianRasterizer(rs)
    z = lambda *s: torch.zeros(*s, device="cuda")
    color, radii, invd = rast(means3D=z(0, 3), means2D=z(0, 3), opacities=z(0, 1), shs=z(0, 1, 3), scales=z(0, 3), rotations=z(0, 4))
    assert radii.numel() == 0
    assert torch.allclose(color, cam.bg.cuda()[:, None, None].expand(3, 48, 64))
    scene = TO.make_scene(100, seed=1, sh_coeffs=1)
    scene["means3D"][:, 2] -= 100.0
    got = U.run_cuda(U.make_args(scene, "sh"), cam, *_weights(cam))
    assert (got["radii"] == 0).all()
    assert all(np.all(v == 0) for v in got["grads"].values() if v is not None)


def test_cull_on_off_identical_image():
    """Exact tile culling must not change a single pixel or gradient beyond atomics' summation order."""
    import diff_gaussian_rasterization as dgr
    scene = TO.make_scene(5000, seed=40, log_scale_mean=-2.5, log_scale_std=0.9)
    cam = TO.make_camera(320, 200, sh_degree=3, bg=(0.3, 0.2, 0.1))
    args = U.make_args(scene, "sh")
    wc, wd = _weights(cam)
    dgr.set_option("cull", 0)
    try:
        a = U.run_cuda(args, cam, wc, wd)
    finally:
        dgr.set_option("cull", 1)
    b = U.run_cuda(args, cam, wc, wd)
    assert np.array_equal(a["color"], b["color"]), "culling changed the image: max diff %g" % np.abs(a["color"] - b["color"]).max()
    assert np.array_equal(a["invdepth"], b["invdepth"])
    for k, v in a["grads"].items():
        if v is not None:
            # the two runs sum the same per-pixel terms in different groupings (different tile lists, float atomics)
            err = np.abs(v - b["grads"][k]).max() / (np.abs(v).max() + 1e-20)
            assert err <= 5e-4, (k, err)


def test_mark_visible():
    import diff_gaussian_rasterization as dgr
    scene = TO.make_scene(1000, seed=41)
    cam = TO.make_camera(64, 64, eye=(0.0, 0.0, -0.5))
    rast = dgr.GaussianRasterizer(U.settings_to(cam, "cuda"))
    vis = rast.markVisible(scene["means3D"].cuda()).cpu().numpy()
    pv = (scene["means3D"] @ cam.viewmatrix[:3, :3] + cam.viewmatrix[3, :3])[:, 2].numpy()
    assert (vis == (pv > 0.2)).all()


def test_python_side_sh_and_cov_switches_agree():
    """The reference's self-consistency switches (--convert_SHs_python / --compute_cov3D_python,
    gaussian_renderer/__init__.py:64-80): torch-side colour / covariance vs in-kernel must give the same image."""
    scene = TO.make_scene(2500, seed=42, log_scale_mean=-2.8)
    cam = TO.make_camera(160, 100, sh_degree=3)
    a = U.run_cuda(U.make_args(scene, "sh"), cam)
    d = scene["means3D"] - cam.campos[None]
    d = d / d.norm(dim=1, keepdim=True)
    cols = torch.clamp_min(TO.eval_sh(3, scene["shs"].transpose(1, 2), d) + 0.5, 0.0)
    cov = TO.build_covariance(scene["scales"], 1.0, scene["rotations"], normalize=True)
    args = dict(means3D=scene["means3D"], shs=None, colors_precomp=cols, opacities=scene["opacities"], scales=None,
                rotations=None, cov3D_precomp=cov)
    b = U.run_cuda(args, cam)
    U.assert_image_close(a["color"], b["color"], "python-vs-kernel SH/cov")


@pytest.mark.parametrize("n,bits", [(1, 32), (1000, 32), (4096, 13), (100003, 32), (1 << 20, 20)])
def test_radix_sort_pairs_is_stable_and_sorted(n, bits):
    import ctypes
    import diff_gaussian_rasterization as dgr
    g = torch.Generator().manual_seed(n)
    hi = (1 << bits) - 1
    keys = torch.randint(0, min(hi, 2**31 - 1) + 1, (n,), generator=g, dtype=torch.int64)
    if bits == 32:
        keys = keys * 2 + torch.randint(0, 2, (n,), generator=g)
    keys = keys.clamp_max(hi)
    k32 = keys.to(torch.uint32).cuda() if hasattr(torch, "uint32") else None
    kd = (keys & 0xffffffff).to(torch.int64)
    kdev = torch.empty(n, dtype=torch.int32, device="cuda")
    kdev.copy_(torch.from_numpy(kd.numpy().astype(np.uint32).view(np.int32)))
    vdev = torch.arange(n, dtype=torch.int32, device="cuda")
    arena = dgr._Arena(torch.device("cuda", torch.cuda.current_device()))
    rc = dgr._C.gsb_sort_pairs(kdev.data_ptr(), vdev.data_ptr(), n, 0, bits, arena.cb, None,
                               torch.cuda.current_stream().cuda_stream)
    dgr._check(rc, arena)
    torch.cuda.synchronize()
    out_k = kdev.cpu().numpy().view(np.uint32).astype(np.int64)
    out_v = vdev.cpu().numpy()
    order = np.argsort(kd.numpy(), kind="stable")
    assert np.array_equal(out_k, kd.numpy()[order])
    assert np.array_equal(out_v, order.astype(np.int32))


def test_full_size_properties_config2():
    """BASELINE.json configs[1]: 100k gaussians, 800x800, SH degree 0 -- checked against the oracle in full."""
    scene = TO.make_scene(100_000, seed=0, sh_coeffs=1, log_scale_mean=-4.3)
    cam = TO.make_camera(800, 800, sh_degree=0)
    args = U.make_args(scene, "sh")
    wc, wd = _weights(cam)
    got = U.run_cuda(args, cam, wc, wd)
    ref = U.run_oracle(args, cam, wc, wd)
    assert (got["radii"] == ref["radii"]).all()
    U.assert_image_close(got["color"], ref["color"], "color")
    U.assert_grads_close(got["grads"], ref["grads"], flips=_flips(got, ref))


def test_speculative_capacity_repair_path():
    """gsb_forward with a too-small capacity estimate must repair itself and give the exact result."""
    import diff_gaussian_rasterization as dgr
    scene = TO.make_scene(4000, seed=50, log_scale_mean=-2.8)
    cam = TO.make_camera(192, 128, sh_degree=2)
    args = U.make_args(scene, "sh")
    wc, wd = _weights(cam)
    dgr._capacity_hints.clear()
    a = U.run_cuda(args, cam, wc, wd)                 # exact path (no estimate yet)
    key = next(iter(dgr._capacity_hints))
    big = dgr._capacity_hints[key]
    b = U.run_cuda(args, cam, wc, wd)                 # speculative path, estimate large enough
    dgr._capacity_hints[key] = 100                    # far too small: forces the repair
    c = U.run_cuda(args, cam, wc, wd)
    assert dgr._capacity_hints[key] == big            # re-learned from the true count
    for other in (b, c):
        assert np.array_equal(a["color"], other["color"]) and np.array_equal(a["radii"], other["radii"])
        for k, v in a["grads"].items():
            if v is not None:
                assert np.abs(v - other["grads"][k]).max() <= 1e-4 * (np.abs(v).max() + 1e-20), k


def test_view_batch_path_matches_autograd_path():
    """render_views_backward (direct accumulation over views) == sum of per-view render() + autograd."""
    import math
    import bench
    from gaussian_renderer import GradientBucket, render, render_views_backward
    dev = torch.device("cuda", 0)
    scene = TO.make_scene(3000, seed=51, log_scale_mean=-3.0)
    W, H = 160, 96
    cams = [bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(i, 3.0), dev) for i in range(3)]
    gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(i)).to(dev) for i in range(3)]
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)

    def run(fused):
        pc = bench.BenchGaussians(scene, 3, dev)
        bucket = GradientBucket(pc.parameters())
        if fused:
            out = render_views_backward(cams, pc, bench.Pipe(), bg, lambda img, d, i: (img - gts[i]).abs().mean() + 0.1 * d.mean())
            losses = out["losses"]
        else:
            ls = []
            for i, cam in enumerate(cams):
                pkg = render(cam, pc, bench.Pipe(), bg)
                loss = (pkg["render"] - gts[i]).abs().mean() + 0.1 * pkg["depth"].mean()
                loss.backward()
                ls.append(loss.detach())
            losses = torch.stack(ls)
        return losses.cpu().numpy(), bucket.flat.cpu().numpy()

    l0, g0 = run(False)
    l1, g1 = run(True)
    assert np.allclose(l0, l1, rtol=1e-6, atol=1e-7)
    assert np.abs(g0 - g1).max() <= 1e-4 * np.abs(g0).max()


def test_view_batch_path_chains_through_activations():
    """Non-leaf inputs (activations of leaf parameters, as in scene.GaussianModel) get their gradient through one
    autograd.backward at the end of the batch."""
    import math
    import bench
    from gaussian_renderer import render, render_views_backward
    dev = torch.device("cuda", 0)
    scene = TO.make_scene(2000, seed=52, log_scale_mean=-3.0)
    W, H = 128, 80
    cams = [bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(i, 3.0), dev) for i in range(2)]
    gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(i)).to(dev) for i in range(2)]
    bg = torch.zeros(3, device=dev)

    class Model:
        active_sh_degree = 3
        max_sh_degree = 3

        def __init__(self):
            mk = lambda t: t.to(dev).clone().requires_grad_(True)
            self._xyz = mk(scene["means3D"]); self._f = mk(scene["shs"])
            self._s = mk(torch.log(scene["scales"])); self._r = mk(scene["rotations"] * 1.7)
            self._o = mk(torch.logit(scene["opacities"].clamp(1e-4, 1 - 1e-4)))
        get_xyz = property(lambda s: s._xyz)
        get_features = property(lambda s: s._f)
        get_opacity = property(lambda s: torch.sigmoid(s._o))
        get_scaling = property(lambda s: torch.exp(s._s))
        get_rotation = property(lambda s: torch.nn.functional.normalize(s._r))
        params = property(lambda s: [s._xyz, s._f, s._s, s._r, s._o])

    m0, m1 = Model(), Model()
    for i, cam in enumerate(cams):
        ((render(cam, m0, bench.Pipe(), bg)["render"] - gts[i]).abs().mean()).backward()
    render_views_backward(cams, m1, bench.Pipe(), bg, lambda img, d, i: (img - gts[i]).abs().mean())
    for p0, p1 in zip(m0.params, m1.params):
        assert p1.grad is not None
        assert (p0.grad - p1.grad).abs().max().item() <= 1e-4 * p0.grad.abs().max().item() + 1e-12


def test_full_size_config3_against_oracle():
    """BASELINE.json configs[2]: 1 M gaussians, 1920x1080, SH degree 3 (the bench workload), forward + backward against the
    CPU oracle in full.  ~10 s of oracle time on the GPU box's host cores."""
    import bench
    import math
    scene = TO.make_scene(1_000_000, seed=0, log_scale_mean=bench.LOG_SCALE_MEAN)
    R, T = bench.view_pose(0, 3.0)
    fovx = math.radians(60.0)
    fovy = 2.0 * math.atan(math.tan(fovx / 2) * 1080 / 1920)
    wvt, full, center = TO.camera_matrices(R, T, fovx, fovy)
    cam = TO.OracleSettings(1080, 1920, math.tan(fovx / 2), math.tan(fovy / 2), torch.zeros(3), 1.0, wvt, full, 3, center)
    args = U.make_args(scene, "sh")
    gen = torch.Generator().manual_seed(11)
    wc = torch.randn(3, 1080, 1920, generator=gen).numpy()
    got = U.run_cuda(args, cam, wc, None)
    ref = U.run_oracle(args, cam, wc, None)
    assert (got["radii"] == ref["radii"]).all()
    mx, frac = U.assert_image_close(got["color"], ref["color"], "color")
    U.assert_image_close(got["invdepth"], ref["invdepth"], "invdepth")
    # size-independent properties: visited instances can only shrink under exact culling; gradients of never-visible
    # gaussians are exactly zero
    vis = ref["radii"] > 0
    assert np.all(got["grads"]["means3D"][~vis] == 0) and np.all(got["grads"]["shs"][~vis] == 0)
    U.assert_grads_close(got["grads"], ref["grads"], flips=_flips(got, ref))


@pytest.mark.parametrize("nviews", [1, 5])
def test_batched_calls_match_single_view_calls(nviews):
    """gsb_forward_batch / gsb_backward_batch against the per-view calls: images identical, summed gradients equal up to
    float summation order, per-view means2D gradients and the densification statistics equal."""
    import math
    import bench
    from gaussian_renderer import GradientBucket, render_views_backward
    dev = torch.device("cuda", 0)
    scene = TO.make_scene(6000, seed=60, log_scale_mean=-3.0)
    W, H = 208, 120
    cams = [bench.BenchCamera(W, H, math.radians(55.0), *bench.view_pose(i, 3.0), dev) for i in range(nviews)]
    gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(i)).to(dev) for i in range(nviews)]
    bg = torch.tensor([0.3, 0.1, 0.2], device=dev)

    def run(batched):
        pc = bench.BenchGaussians(scene, 3, dev)
        bucket = GradientBucket(pc.parameters())
        stats = dict(xyz_gradient_accum=torch.zeros(6000, 1, device=dev), denom=torch.zeros(6000, 1, device=dev))
        out = render_views_backward(cams, pc, bench.Pipe(), bg, lambda img, d, i: (img - gts[i]).abs().mean() + 0.05 * d.mean(),
                                    densify_stats=stats, keep_images=True, batched=batched)
        return (out["losses"].cpu().numpy(), bucket.flat.cpu().numpy(), [im.cpu().numpy() for im in out["images"]],
                stats["xyz_gradient_accum"].cpu().numpy(), stats["denom"].cpu().numpy(), out["radii_max"].cpu().numpy())

    a, b = run(False), run(True)
    assert np.allclose(a[0], b[0], rtol=1e-6, atol=1e-7)
    for x, y in zip(a[2], b[2]):
        assert np.array_equal(x, y)
    # overwrite=True: the first chunk WRITES the gradient buffers, whatever they held
    pc = bench.BenchGaussians(scene, 3, dev)
    bucket = GradientBucket(pc.parameters())
    bucket.flat.fill_(123.0)
    render_views_backward(cams, pc, bench.Pipe(), bg, lambda img, d, i: (img - gts[i]).abs().mean() + 0.05 * d.mean(), overwrite=True)
    assert np.abs(bucket.flat.cpu().numpy() - b[1]).max() <= 1e-4 * np.abs(b[1]).max()
    assert np.abs(a[1] - b[1]).max() <= 1e-4 * np.abs(a[1]).max()
    assert np.abs(a[3] - b[3]).max() <= 1e-4 * (np.abs(a[3]).max() + 1e-20)
    assert np.array_equal(a[4], b[4]) and np.array_equal(a[5], b[5])


def test_batched_path_antialiasing_scale_modifier_background():
    """The view-batch kernels loop over cameras that differ in pose and background; antialiasing and scale_modifier are
    per-call settings carried per view."""
    import math
    import bench
    from gaussian_renderer import GradientBucket, render_views_backward
    dev = torch.device("cuda", 0)
    scene = TO.make_scene(3000, seed=61, log_scale_mean=-2.8)
    W, H = 144, 88
    cams = [bench.BenchCamera(W, H, math.radians(50.0), *bench.view_pose(i + 3, 3.0), dev) for i in range(3)]
    gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(i)).to(dev) for i in range(3)]
    bg = torch.tensor([0.9, 0.5, 0.1], device=dev)

    class AAPipe(bench.Pipe):
        antialiasing = True

    def run(batched):
        pc = bench.BenchGaussians(scene, 3, dev)
        pc.active_sh_degree = 2
        bucket = GradientBucket(pc.parameters())
        out = render_views_backward(cams, pc, AAPipe(), bg, lambda img, d, i: ((img - gts[i]) ** 2).mean(), scaling_modifier=0.8,
                                    keep_images=True, batched=batched)
        return [im.cpu().numpy() for im in out["images"]], bucket.flat.cpu().numpy()

    (ia, ga), (ib, gb) = run(False), run(True)
    for x, y in zip(ia, ib):
        assert np.array_equal(x, y)
    assert np.abs(ga - gb).max() <= 1e-4 * np.abs(ga).max()
    # and against the oracle for one of the views
    cam = cams[1]
    os_ = TO.OracleSettings(H, W, math.tan(cam.FoVx * 0.5), math.tan(cam.FoVy * 0.5), bg.cpu(), 0.8, cam.world_view_transform.cpu(),
                            cam.full_proj_transform.cpu(), 2, cam.camera_center.cpu(), False, False, True)
    ref = U.run_oracle(U.make_args(scene, "sh"), os_)
    U.assert_image_close(ib[1], ref["color"], "batched AA view vs oracle")


def test_debug_flag_float64_noncontiguous_and_side_stream():
    """debug=True synchronises after every kernel; float64 / non-contiguous inputs are converted; a non-default stream works."""
    import diff_gaussian_rasterization as dgr
    scene = TO.make_scene(1500, seed=62, log_scale_mean=-2.7)
    cam = TO.make_camera(112, 80, sh_degree=1)
    base = U.run_cuda(U.make_args(scene, "sh"), cam)
    rs = U.settings_to(cam._replace(debug=True), "cuda")
    dev = torch.device("cuda", 0)
    big = torch.zeros(1500, 6, dtype=torch.float64, device=dev)
    big[:, ::2] = scene["means3D"].double().to(dev)
    means_nc = big[:, ::2]                                   # float64, non-contiguous view
    assert not means_nc.is_contiguous()
    side = torch.cuda.Stream(device=dev)
    args = {k: v.to(dev) for k, v in U.make_args(scene, "sh").items() if v is not None}
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        rast = dgr.GaussianRasterizer(rs)
        color, radii, invd = rast(means3D=means_nc, means2D=torch.zeros(1500, 3, device=dev), shs=args["shs"], opacities=args["opacities"],
                                  scales=args["scales"], rotations=args["rotations"])
    side.synchronize()
    assert np.array_equal(color.cpu().numpy(), base["color"])
    assert np.array_equal(radii.cpu().numpy(), base["radii"])


@pytest.mark.parametrize("shape", [(3, 37, 53), (3, 128, 200), (1, 16, 16)])
def test_fused_photometric_loss_and_fused_ssim(shape, golden):
    """csrc/loss.cu against the reference-pinned restatement (oracle photometric_loss; golden vectors from the reference's own
    utils/loss_utils.py): loss value within 1e-6, gradient within 1e-4 rel, including clamp's gradient mask."""
    import diff_gaussian_rasterization as dgr
    import fused_ssim
    dev = torch.device("cuda", 0)
    if shape == (3, 37, 53):
        img, gt = torch.tensor(golden["loss_img"]), torch.tensor(golden["loss_gt"])
    else:
        g = torch.Generator().manual_seed(shape[1])
        img = torch.rand(*shape, generator=g) * 1.3 - 0.15          # some values outside [0,1]: clamp mask exercised
        gt = torch.rand(*shape, generator=g)
    ref_in = img.clone().requires_grad_(True)
    ref = TO.photometric_loss(ref_in, gt, 0.2)
    (ref_g,) = torch.autograd.grad(ref, ref_in)
    loss, grad, parts = dgr.photometric_loss_and_grad(img.to(dev), gt.to(dev), 0.2)
    assert abs(float(loss.item()) - float(ref)) < 2e-6
    err = (grad.cpu() - ref_g).abs().max().item() / (ref_g.abs().max().item() + 1e-20)
    assert err < 1e-4, err
    if shape == (3, 37, 53):
        assert abs(float(loss.item()) - float(golden["loss_total"])) < 2e-6
        assert np.abs(grad.cpu().numpy() - golden["loss_grad"]).max() < 1e-4 * np.abs(golden["loss_grad"]).max()
    # drop-in fused_ssim: mean SSIM with autograd, NO clamp (the input keeps its values outside [0,1], as in the reference's kernel)
    a = img.to(dev).requires_grad_(True)
    v = fused_ssim.fused_ssim(a.unsqueeze(0), gt.to(dev).unsqueeze(0))
    a_ref = img.clone().requires_grad_(True)
    v_ref = TO.ssim(a_ref, gt)
    assert abs(float(v.item()) - float(v_ref)) < 2e-6
    (ga,) = torch.autograd.grad(v, a)
    (ga_ref,) = torch.autograd.grad(v_ref, a_ref)
    assert (ga.cpu() - ga_ref).abs().max().item() <= 1e-4 * ga_ref.abs().max().item()


@pytest.mark.parametrize("opts", [{"sort_small": -1}, {"sort_small": 1}, {"pre_tma": 0}, {"tile_order": 0}],
                         ids=["sort_16_keys_per_thread", "sort_4_keys_per_thread", "pre_tma_off", "tile_order_off"])
def test_ab_options_keep_results(opts):
    """The tuning knobs change neither the image nor, beyond the order of float atomics, the gradients; single-view and
    view-batch path."""
    import math
    import bench
    import diff_gaussian_rasterization as dgr
    from gaussian_renderer import GradientBucket, render_views_backward
    dev = torch.device("cuda", 0)
    scene = TO.make_scene(20000, seed=71, log_scale_mean=-3.4)
    cam = TO.make_camera(400, 240, sh_degree=3, bg=(0.2, 0.4, 0.1))
    args = U.make_args(scene, "sh")
    wc, wd = _weights(cam)
    W, H = 320, 200
    cams = [bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(i, 3.0), dev) for i in range(3)]
    gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(i)).to(dev) for i in range(3)]

    def batch():
        pc = bench.BenchGaussians(scene, 3, dev)
        bucket = GradientBucket(pc.parameters())
        out = render_views_backward(cams, pc, bench.Pipe(), torch.zeros(3, device=dev), lambda img, d, i: (img - gts[i]).abs().mean() + 0.1 * d.mean())
        return out["losses"].cpu().numpy(), bucket.flat.cpu().numpy()

    base, (bl, bg_) = U.run_cuda(args, cam, wc, wd), batch()
    defaults = {"sort_small": 0, "pre_tma": 1, "tile_order": 1}
    try:
        for k, v in opts.items():
            dgr.set_option(k, v)
        got, (l, g) = U.run_cuda(args, cam, wc, wd), batch()
    finally:
        for k in opts:
            dgr.set_option(k, defaults[k])
    assert np.array_equal(base["color"], got["color"]) and np.array_equal(base["invdepth"], got["invdepth"])
    assert np.array_equal(base["radii"], got["radii"]) and np.array_equal(bl, l)
    for k, v in base["grads"].items():
        if v is not None:
            assert np.abs(v - got["grads"][k]).max() <= 1e-4 * (np.abs(v).max() + 1e-20), k
    assert np.abs(bg_ - g).max() <= 1e-4 * np.abs(bg_).max()
