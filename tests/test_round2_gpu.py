"""GPU tests of the round-2 additions (-m gpu): the sync-free view-batch step, the chunked gradient kernel with overlapped
reduction (two GPUs when the box has them), the heavy-tiles-first launch order, the oracle-INDEPENDENT finite-difference check of
the CUDA forward against the CUDA backward, and the training state at a million gaussians through two densifications."""
import math
import os
import sys

import numpy as np
import pytest
import torch

from oracle import torch_oracle as TO
import gs_test_util as U

pytestmark = pytest.mark.gpu


def _batch_inputs(n, views, W, H, dev, seed=51):
    import bench
    scene = TO.make_scene(n, seed=seed, log_scale_mean=-3.0)
    cams = [bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(i, 3.0), dev) for i in range(views)]
    gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(i)).to(dev) for i in range(views)]
    return scene, cams, gts, torch.tensor([0.1, 0.2, 0.3], device=dev)


def _batch_step(scene, cams, gts, bg, dev, **kw):
    import bench
    from gaussian_renderer import GradientBucket, render_views_backward
    pc = bench.BenchGaussians(scene, 3, dev)
    bucket = GradientBucket(pc.parameters())
    out = render_views_backward(cams, pc, bench.Pipe(), bg, lambda img, d, i: (img - gts[i]).abs().mean() + 0.1 * d.mean(),
                                keep_images=True, **kw)
    return out["losses"].cpu().numpy(), bucket.flat.cpu().numpy(), [im.cpu().numpy() for im in out["images"]]


def test_sync_free_view_batch_step():
    """gsb_forward_batch_async == gsb_forward_batch (images bit-equal, gradients up to atomics order); counts and their running
    maximum on the device; an undersized capacity is survived (truncated lists, no out-of-bounds access: compute-sanitizer
    clean in profiles/) and reported by AsyncCapacity.check()."""
    from gaussian_renderer import AsyncCapacity
    dev = torch.device("cuda", 0)
    scene, cams, gts, bg = _batch_inputs(20000, 5, 320, 200, dev)
    l0, g0, im0 = _batch_step(scene, cams, gts, bg, dev)
    cap = AsyncCapacity(dev)
    _batch_step(scene, cams, gts, bg, dev, capacity=cap)                 # synchronous once: learns the capacity
    assert cap.capacity >= 1 << 20 and cap.observed_max() == 0
    l1, g1, im1 = _batch_step(scene, cams, gts, bg, dev, capacity=cap)   # no read-back
    for a, b in zip(im0, im1):
        assert np.array_equal(a, b)
    assert np.array_equal(l0, l1) and np.abs(g0 - g1).max() <= 1e-4 * np.abs(g0).max()
    seen = cap.observed_max()
    assert 0 < seen <= cap.capacity and cap.check()
    assert int(cap.counts[:5].max()) == seen and int(cap.counts[5:16].sum()) == 0
    small = AsyncCapacity(dev, capacity=seen // 3)
    _batch_step(scene, cams, gts, bg, dev, capacity=small)
    torch.cuda.synchronize()
    assert small.observed_max() == seen and not small.check() and small.capacity >= seen


def test_chunked_gradient_kernel_and_tile_order():
    import diff_gaussian_rasterization as dgr
    dev = torch.device("cuda", 0)
    scene, cams, gts, bg = _batch_inputs(30000, 4, 320, 200, dev, seed=52)
    l0, g0, im0 = _batch_step(scene, cams, gts, bg, dev, overwrite=True)
    seen = []
    dgr.set_option("tile_order", 0)          # default is on (heavy tiles first): the A/B is "off"
    try:
        l1, g1, im1 = _batch_step(scene, cams, gts, bg, dev, overwrite=True, grad_chunks=4,
                                  on_grad_chunk=lambda c, a, b: seen.append((c, a, b)))
        cam = TO.make_camera(200, 120, sh_degree=3)
        sc1 = TO.make_scene(3000, seed=21, log_scale_mean=-3.0)
        gen = torch.Generator().manual_seed(5)
        wc = torch.randn(3, 120, 200, generator=gen).numpy()
        single = U.run_cuda(U.make_args(sc1, "sh"), cam, wc, None)
    finally:
        dgr.set_option("tile_order", 1)
    assert [c for c, _, _ in seen] == [0, 1, 2, 3] and seen[0][1] == 0 and seen[-1][2] == 30000
    assert all(seen[i][2] == seen[i + 1][1] for i in range(3))
    for a, b in zip(im0, im1):
        assert np.array_equal(a, b)
    assert np.array_equal(l0, l1) and np.abs(g0 - g1).max() <= 1e-4 * np.abs(g0).max()
    ref = U.run_oracle(U.make_args(sc1, "sh"), cam, wc, None)
    U.assert_image_close(single["color"], ref["color"], "tile_order single view")
    U.assert_grads_close(single["grads"], ref["grads"], flips=U.count_flips(single["color"], ref["color"]))


def test_finite_differences_of_cuda_forward_match_cuda_backward():
    """The one gradient check that needs NO oracle: central differences of the CUDA forward (loss accumulated in float64 from
    the float32 images) against the CUDA backward, on > 200 randomly chosen parameters of every kind, anti-aliasing on.

    The forward is piecewise smooth: a (pixel, gaussian) pair enters or leaves the sum when alpha crosses 1/255, and the
    analytic gradient (the reference's too) ignores that boundary term.  The loss weights are smooth and positive so that the
    interior term adds coherently while the boundary term stays a few percent (it is largest for the scales, whose growth
    pushes the whole 1/255 contour outwards: 5-7 % relative L2 measured for means / scales / rotations, 1 % for opacities,
    6e-6 for the SH coefficients, in which the image is linear), and the bounds are
    relative: an error in the derivation (sign, factor, missing term) shows as O(1) and breaks the correlation."""
    import diff_gaussian_rasterization as dgr
    dev = torch.device("cuda", 0)
    H, W = 96, 144
    scene = TO.make_scene(400, seed=91, log_scale_mean=-2.2)
    scene["opacities"] = scene["opacities"].clamp(0.05, 0.9)
    cam = U.settings_to(TO.make_camera(W, H, sh_degree=3, antialiasing=True, bg=(0.2, 0.3, 0.1), scale_modifier=0.9), dev)
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float64), torch.arange(W, dtype=torch.float64), indexing="ij")
    wc = torch.stack([1.0 + 0.5 * torch.sin(xx / 17 + c) * torch.cos(yy / 13) for c in range(3)]).to(dev)
    wd = (0.5 + 0.25 * torch.cos(xx / 23) * torch.sin(yy / 19))[None].to(dev)
    base = {k: scene[k].to(dev).clone() for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    rast = dgr.GaussianRasterizer(raster_settings=cam)

    def loss64(p):
        color, _, invd = rast(means3D=p["means3D"], means2D=torch.zeros_like(p["means3D"]), shs=p["shs"], colors_precomp=None,
                              opacities=p["opacities"], scales=p["scales"], rotations=p["rotations"], cov3D_precomp=None)
        return (color.double() * wc).sum() + (invd.double() * wd).sum(), color, invd

    leaves = {k: v.clone().requires_grad_(True) for k, v in base.items()}
    color, radii, invd = rast(means3D=leaves["means3D"], means2D=torch.zeros_like(leaves["means3D"], requires_grad=True),
                              shs=leaves["shs"], colors_precomp=None, opacities=leaves["opacities"], scales=leaves["scales"],
                              rotations=leaves["rotations"], cov3D_precomp=None)
    ((color * wc.float()).sum() + (invd * wd.float()).sum()).backward()
    grads = {k: v.grad.detach() for k, v in leaves.items()}
    visible = (radii > 0).nonzero().reshape(-1)
    assert visible.numel() > 200
    # absolute steps per parameter kind (scene extent 1, camera distance ~3, scales ~0.1, unit quaternions)
    steps = {"means3D": 2e-3, "scales": None, "rotations": 2e-2, "opacities": 1e-2, "shs": 2e-2}
    rng = np.random.default_rng(3)
    report = {}
    with torch.no_grad():
        for kind, h in steps.items():
            n = 48 if kind != "shs" else 64
            fd, an = [], []
            for _ in range(n):
                i = int(visible[rng.integers(visible.numel())])
                idx = (i,) + tuple(int(rng.integers(s)) for s in base[kind].shape[1:])
                hh = h if h is not None else 0.03 * float(base[kind][idx])      # scales: 3 % of the value
                plus = {k: v.clone() for k, v in base.items()}
                minus = {k: v.clone() for k, v in base.items()}
                plus[kind][idx] += hh
                minus[kind][idx] -= hh
                lp, cp, dp = loss64(plus)
                lm, cm, dm = loss64(minus)
                # difference accumulated per pixel in float64 (identical pixels cancel exactly: the forward is deterministic)
                d = ((cp.double() - cm.double()) * wc).sum() + ((dp.double() - dm.double()) * wd).sum()
                fd.append(float(d) / (2 * hh))
                an.append(float(grads[kind][idx]))
            fd, an = np.array(fd), np.array(an)
            rms = float(np.sqrt((an ** 2).mean()))
            rel_l2 = float(np.linalg.norm(fd - an) / (np.linalg.norm(an) + 1e-30))
            ok = np.abs(fd - an) <= 0.15 * np.abs(an) + 0.05 * rms
            report[kind] = (rel_l2, float(ok.mean()), float(np.corrcoef(fd, an)[0, 1]))
    print("finite differences vs backward (rel L2, fraction within bound, correlation):", report)
    for kind, (rel_l2, frac_ok, corr) in report.items():
        assert rel_l2 < (0.12 if kind in ("means3D", "scales", "rotations") else 0.03) and frac_ok >= 0.85 and corr > 0.995, (kind, report)


def _two_gpu_worker(rank, world, port, paths, out_path):
    for p in paths:
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    import bench
    from gaussian_renderer import AsyncCapacity, GradientBucket, render_views_backward
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    try:
        V, W, H, P = 3, 320, 200, 40000
        scene = TO.make_scene(P, seed=77, log_scale_mean=-3.2)
        bg = torch.tensor([0.1, 0.2, 0.3], device=dev)

        def cams_of(r):
            return [bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(r * V + i, 3.0), dev) for i in range(V)]

        def gts_of(r):
            return [torch.rand(3, H, W, generator=torch.Generator().manual_seed(100 * r + i)).to(dev) for i in range(V)]

        def local(r, **kw):
            pc = bench.BenchGaussians(scene, 3, dev)
            bucket = GradientBucket(pc.parameters())
            gts = gts_of(r)
            render_views_backward(cams_of(r), pc, bench.Pipe(), bg, lambda img, d, i: (img - gts[i]).abs().mean(), overwrite=True, **kw)
            return bucket

        cap = AsyncCapacity(dev)
        local(rank, capacity=cap)                         # learn the capacity (synchronous once)
        pending = []
        holder = {}

        def on_chunk(_c, p0, p1):
            pending.extend(holder["b"].all_reduce_rows(p0, p1))

        # the bucket must exist before the callback fires: build the step by hand
        pc = bench.BenchGaussians(scene, 3, dev)
        holder["b"] = GradientBucket(pc.parameters())
        gts = gts_of(rank)
        render_views_backward(cams_of(rank), pc, bench.Pipe(), bg, lambda img, d, i: (img - gts[i]).abs().mean(), overwrite=True,
                              capacity=cap, grad_chunks=4, on_grad_chunk=on_chunk)
        pending.extend(holder["b"].all_reduce_rest())
        GradientBucket.wait_all(pending)
        reduced = holder["b"].flat.clone()
        assert cap.check()
        # reference: every rank's views in THIS process, summed
        total = sum(local(r).flat for r in range(world))
        err = float((reduced - total).abs().max() / total.abs().max())
        one = local(rank)
        one.all_reduce()
        err_single_collective = float((one.flat - total).abs().max() / total.abs().max())
        # fused reduce-scatter: the gradient kernel adds every row into its owner's buffer over NVLink (peer-mapped memory),
        # then a barrier and an in-place all-gather; three steps so that both halves of the double buffer are used twice
        from gaussian_renderer.peer import PeerGradientBucket
        pc = bench.BenchGaussians(scene, 3, dev)
        named = {"means3D": pc._xyz, "shs": pc._shs, "opacities": pc._opacity, "scales": pc._scaling, "rotations": pc._rotation}
        pb = PeerGradientBucket(named)
        errs_peer = []
        for _step in range(4):
            pb.begin_step()
            render_views_backward(cams_of(rank), pc, bench.Pipe(), bg, lambda img, d, i: (img - gts[i]).abs().mean(), capacity=cap,
                                  peers=pb.table())
            pb.finish()
            flat = torch.cat([pc._xyz.grad.reshape(-1), pc._shs.grad.reshape(-1), pc._opacity.grad.reshape(-1),
                              pc._scaling.grad.reshape(-1), pc._rotation.grad.reshape(-1)])
            # `total` was laid out by GradientBucket: narrow parameters first, SH last -- compare tensor by tensor instead
            ref_pc = bench.BenchGaussians(scene, 3, dev)
            ref_b = GradientBucket(ref_pc.parameters())
            ref_b.flat.copy_(total)
            ref_flat = torch.cat([ref_pc._xyz.grad.reshape(-1), ref_pc._shs.grad.reshape(-1), ref_pc._opacity.grad.reshape(-1),
                                  ref_pc._scaling.grad.reshape(-1), ref_pc._rotation.grad.reshape(-1)])
            errs_peer.append(float((flat - ref_flat).abs().max() / ref_flat.abs().max()))
        torch.cuda.synchronize()
        pb.close()
        if rank == 0:
            with open(out_path, "w") as f:
                f.write(f"{err} {err_single_collective} {len(pending)} {max(errs_peer)}")
        assert err <= 1e-4 and err_single_collective <= 1e-4 and max(errs_peer) <= 1e-4, (err, err_single_collective, errs_peer)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_gpu_chunked_all_reduce_equals_single_process_sum(tmp_path):
    """On real hardware: two ranks, three views each, the sync-free step with the gradient kernel in four chunks, each chunk's
    SH rows all-reduced while the next computes, one collective for the narrow parameters at the end -- the reduced bucket equals the sum of both ranks'
    buckets computed in one process, and equals the single all-reduce of the whole bucket.  Then the FUSED reduce-scatter
    (gsb_backward_batch_peer: TMA bulk reduce-adds into the owner's peer-mapped buffer, barrier, in-place all-gather) gives the
    same sums on both halves of its double buffer."""
    import torch.multiprocessing as mp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    paths = [root, os.path.join(root, "gaussian-splatting_b200"), os.path.join(root, "tests")]
    out = str(tmp_path / "two_gpu.txt")
    mp.spawn(_two_gpu_worker, args=(2, 29500 + os.getpid() % 2000, paths, out), nprocs=2, join=True)
    err, err1, n, err_peer = open(out).read().split()
    print("two-GPU chunked reduction: rel err", err, "single collective", err1, "handles", n, "| fused reduce-scatter over peer memory:", err_peer)
    assert float(err) <= 1e-4 and int(n) == 5 and float(err_peer) <= 1e-4


def test_training_state_at_a_million_gaussians_through_two_densifications():
    """BASELINE.json configs[3] territory: the flat store at 1.2 M gaussians through optimizer steps and two densify_and_prune
    calls in train.py's order (backward, densify, skipped step) against oracle/model_oracle.py (pinned to the reference class):
    identical counts and row order, parameters to 2e-6, 32-bit offsets intact (59 floats x 2.4 M rows > 2^27)."""
    from types import SimpleNamespace
    from gaussian_store import GaussianModel
    from oracle.model_oracle import GROUPS, ModelOracle
    dev = torch.device("cuda", 0)
    P = 1_200_000
    g = torch.Generator().manual_seed(5)
    raw = {"xyz": torch.randn(P, 3, generator=g), "f_dc": torch.randn(P, 1, 3, generator=g) * 0.5,
           "f_rest": torch.randn(P, 15, 3, generator=g) * 0.1, "opacity": torch.rand(P, 1, generator=g) * 9.0 - 6.5,
           "scaling": torch.randn(P, 3, generator=g) - 3.6, "rotation": torch.randn(P, 4, generator=g)}
    opt = dict(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=30000,
               feature_lr=0.0025, opacity_lr=0.025, scaling_lr=0.005, rotation_lr=0.001, percent_dense=0.01)
    oracle = ModelOracle(raw["xyz"], raw["f_dc"], raw["f_rest"], raw["opacity"], raw["scaling"], raw["rotation"], opt, 4.0)
    m = GaussianModel(3).create_from_tensors(*(raw[k].to(dev) for k in ("xyz", "f_dc", "f_rest", "scaling", "rotation", "opacity")), 4.0)
    m.training_setup(SimpleNamespace(**opt))
    names = ("xyz", "features", "opacity", "scaling", "rotation")

    def grads_for(Pn, seed):
        gg = torch.Generator().manual_seed(seed)
        shapes = {"xyz": (3,), "features": (16, 3), "opacity": (1,), "scaling": (3,), "rotation": (4,)}
        return {n: torch.randn(Pn, *shapes[n], generator=gg) * 1e-3 for n in names}

    def backward(it, seed):
        gr = grads_for(m.P, seed)
        oracle.backward(it, gr)
        m.update_learning_rate(it)
        leaves = {"xyz": m.get_xyz, "features": m.get_features, "opacity": m.get_opacity, "scaling": m.get_scaling, "rotation": m.get_rotation}
        for n in names:
            leaves[n].grad.copy_(gr[n].to(dev).view_as(leaves[n]))
        m.gradients_ready()

    def densify(seed, max_screen):
        gg = torch.Generator().manual_seed(seed)
        denom = torch.randint(0, 3, (m.P, 1), generator=gg).float()
        accum = torch.rand(m.P, 1, generator=gg) * 0.0006 * denom
        oracle.grad_accum, oracle.denom = accum.clone(), denom.clone()
        m.xyz_gradient_accum, m.denom = accum.to(dev), denom.to(dev)
        draws = {}

        def draw(rows):
            draws["u"] = torch.randn(rows, 3, generator=torch.Generator().manual_seed(seed + 1))
            return draws["u"]
        info = m.densify_and_prune(0.0002, 0.005, 4.0, max_screen, unit_samples=draw)
        ref = oracle.densify_and_prune(0.0002, 0.005, 4.0, max_screen, unit_samples=draws["u"])
        assert info == ref, (info, ref)
        return info

    def compare(tag):
        got = {"xyz": m._xyz, "f_dc": m._features[:, :1], "f_rest": m._features[:, 1:], "opacity": m._opacity, "scaling": m._scaling,
               "rotation": m._rotation}
        for n in GROUPS:
            torch.testing.assert_close(got[n].cpu(), oracle.p[n].detach(), rtol=2e-6, atol=2e-6, msg=lambda s: f"{tag} {n}: {s}")

    backward(1, 11); m.optimizer_step(); oracle.optimizer_step()
    backward(2, 12)
    info1 = densify(21, None)
    m.optimizer_step(); oracle.optimizer_step()               # both skip: every parameter was replaced after the backward
    compare("after densify 1")
    backward(3, 13); m.optimizer_step(); oracle.optimizer_step()
    backward(4, 14)
    info2 = densify(22, 20)
    m.optimizer_step(); oracle.optimizer_step()
    backward(5, 15); m.optimizer_step(); oracle.optimizer_step()
    compare("end")
    assert info1["P"] > P and info2["P"] > info1["P"] and info2["n_pruned"] > 0
    assert {n: m.group_steps[n] for n in GROUPS} == {n: int(v) for n, v in oracle.steps().items()}
    print("1.2 M store:", info1, info2)
