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
