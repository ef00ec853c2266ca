"""CPU tests (gloo, world_size 2) of the N>1 host logic: round-robin view sharding, one all-reduce of the flat
gradient bucket, and the densification-statistics reductions.  The rasterizer itself needs a
GPU; here each rank fills its bucket with a deterministic per-view stand-in gradient."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gaussian_renderer import GradientBucket, shard_views
        P = 257
        g = torch.Generator().manual_seed(0)
        params = [torch.randn(P, 3, generator=g, requires_grad=True), torch.randn(P, 16, 3, generator=g, requires_grad=True),
                  torch.randn(P, 1, generator=g, requires_grad=True), torch.randn(P, 3, generator=g, requires_grad=True),
                  torch.randn(P, 4, generator=g, requires_grad=True)]
        bucket = GradientBucket(params)
        assert bucket.flat.numel() == 59 * P
        views = list(range(7))                                   # ragged: 4 views on rank 0, 3 on rank 1
        mine = shard_views(views, rank, world)
        assert mine == [v for v in views if v % world == rank]
        # stand-in for per-view backward: loss_v = sum_p w_v * param  ->  grad = w_v (accumulated through autograd)
        for v in mine:
            loss = sum((p * float(v + 1)).sum() for p in params)
            loss.backward()
        for p in params:                                          # .grad is a view of the flat bucket
            assert p.grad.data_ptr() >= bucket.flat.data_ptr()
        bucket.all_reduce()
        expect = float(sum(v + 1 for v in views))
        ok = all(torch.allclose(p.grad, torch.full_like(p, expect)) for p in params)
        # densification statistics: sum / sum / max
        acc = torch.full((P, 1), float(rank + 1)); den = torch.full((P, 1), float(rank + 2)); rad = torch.full((P,), rank * 5, dtype=torch.int32)
        GradientBucket.reduce_densification_stats(acc, den, rad)
        ok = ok and torch.allclose(acc, torch.full((P, 1), 3.0)) and torch.allclose(den, torch.full((P, 1), 5.0)) and int(rad.max()) == 5
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_view_sharding_and_single_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}


def test_bucket_world1_is_noop():
    sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
    from gaussian_renderer import GradientBucket
    p = torch.zeros(5, 3, requires_grad=True)
    b = GradientBucket([p])
    (p * 2.0).sum().backward()
    assert torch.equal(b.all_reduce(), torch.full((15,), 2.0))
    b.zero_()
    assert float(p.grad.abs().sum()) == 0.0


def _store_worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from gaussian_renderer import shard_views
        from gaussian_store import GaussianModel, store_offsets
        P = 131
        m = GaussianModel(3)
        m._allocate(P, torch.device("cpu"))                       # layout only: the kernels are CUDA only
        views = list(range(5))
        for v in shard_views(views, rank, world):                 # stand-in for gsb_backward_batch accumulating into .grad
            for k, leaf in enumerate((m.get_xyz, m.get_features, m.get_opacity, m.get_scaling, m.get_rotation)):
                leaf.grad += float((v + 1) * (k + 1))
        dist.all_reduce(m.grad)                                   # the ONE collective of the step
        o, tot = store_offsets(P, 16), float(sum(v + 1 for v in views))
        ok = m.grad.numel() == 59 * P
        for k, (name, width) in enumerate((("xyz", 3), ("features", 48), ("opacity", 1), ("scaling", 3), ("rotation", 4))):
            seg = m.grad[o[name]:o[name] + width * P]
            ok = ok and bool(torch.all(seg == tot * (k + 1)))
        ok = ok and bool(torch.all(m.get_rotation.grad == tot * 5))        # the leaves see the reduced values
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_store_gradient_buffer_is_the_allreduce_bucket_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_store_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}


def _train_worker(rank, world, port, q, lib):
    """One rank of a 2-process data-parallel training step with the REAL kernels' source (host build) and gloo."""
    for p in (os.path.join(ROOT, "gaussian-splatting_b200"), os.path.join(ROOT, "tests", "host_emul"), ROOT, os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import math
    from types import SimpleNamespace
    import build as host_build
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        with host_build.python_layer_on_host(lib) as dgr:
            import bench
            from gaussian_renderer import render_views_backward, shard_views
            from gaussian_renderer.synthetic import make_scene
            from gaussian_store import GaussianModel
            sc = make_scene(220, seed=4, log_scale_mean=-2.5)
            op = sc["opacities"].clamp(1e-6, 1 - 1e-6).reshape(-1, 1)
            m = GaussianModel(3)
            m.active_sh_degree = 3
            m.create_from_tensors(sc["means3D"], sc["shs"][:, :1].contiguous(), sc["shs"][:, 1:].contiguous(), torch.log(sc["scales"]),
                                  sc["rotations"], torch.log(op / (1 - op)), 1.0)
            m.training_setup(SimpleNamespace(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01,
                                             position_lr_max_steps=30000, feature_lr=0.0025, opacity_lr=0.025, scaling_lr=0.005,
                                             rotation_lr=0.001, percent_dense=0.01))
            W, H, NV = 48, 32, 2
            cams = [bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(i, 3.0), "cpu") for i in range(NV)]
            gts = [torch.rand(3, H, W, generator=torch.Generator().manual_seed(100 + i)) for i in range(NV)]
            mine = shard_views(list(range(NV)), rank, world)
            m.update_learning_rate(1)
            # the reduction protocol of bench.py: rows of the features group as each chunk of the gradient kernel is enqueued,
            # one launch for the narrow groups behind the last chunk (world 1: both are no-ops)
            bucket, pending = m.gradient_bucket(), []
            render_views_backward([cams[i] for i in mine], m, bench.Pipe(), torch.zeros(3),
                                  lambda img, d, k: dgr.l1_loss_and_grad(img, gts[mine[k]]), loss_returns_grad=True, overwrite=True,
                                  grad_chunks=2, on_grad_chunk=lambda c, p0, p1: pending.extend(bucket.all_reduce_rows(p0, p1)))
            pending.extend(bucket.all_reduce_rest())
            bucket.wait_all(pending)
            assert (len(pending) >= 2) == (world > 1)
            grad = m.grad.clone()
            m.optimizer_step()
            q.put((rank, world, grad.numpy(), m.store.numpy().copy()))
    finally:
        if world > 1:
            dist.destroy_process_group()


def test_data_parallel_training_step_with_kernel_sources_gloo_world2(host_lib):
    """View-parallel step end to end on the CPU: each rank renders its views with the kernels' source (tests/host_emul),
    the chunk-wise gloo all-reduce of the store's gradient buffer, fused Adam on every rank.  The reduced gradient equals the
    single-process gradient over all views, and both ranks end with bit-identical parameters."""
    import numpy as np
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q, host_lib)) for r in range(2)]
    procs.append(ctx.Process(target=_train_worker, args=(0, 1, port + 1, q, host_lib)))     # single-process reference
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in range(3)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    single = next(g for g in got if g[1] == 1)
    ranks = sorted((g for g in got if g[1] == 2), key=lambda g: g[0])
    assert np.array_equal(ranks[0][2], ranks[1][2]) and np.array_equal(ranks[0][3], ranks[1][3])
    scale = np.abs(single[2]).max()
    assert scale > 0 and np.abs(ranks[0][2] - single[2]).max() <= 1e-5 * scale
    assert not np.array_equal(ranks[0][3], single[3] * 0)            # parameters were stepped
    moved = np.abs(ranks[0][3] - single[3])
    assert float((moved > 1e-3).mean()) < 1e-3                        # same step up to sign flips of ~zero gradients
