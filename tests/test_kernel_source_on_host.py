"""Runs the SOURCE of the CUDA kernels on the CPU (tests/host_emul: the unmodified .cu files compiled with g++ against a shim
of the CUDA language -- blocks sequential, threads of a block = host threads, so barriers, shuffles, votes and shared
memory keep their meaning) and drives the real Python layer with it.  Index arithmetic, scans, sort orders, row orders
and the host classes' buffer surgery are checked against the reference fixtures and the oracles where no GPU exists; the
hot path itself is covered by tests/test_hot_path_source_on_host.py.  The GPU parity tests proper are the -m gpu files;
nothing here is a product path (the library raises on CPU tensors)."""
import ctypes
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from gaussian_store import GaussianModel, store_offsets
from oracle.model_oracle import GROUPS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "reference_model.npz")
RAW_KEY = {"xyz": "_xyz", "f_dc": "_features_dc", "f_rest": "_features_rest", "opacity": "_opacity", "scaling": "_scaling",
           "rotation": "_rotation"}
ACT = ("xyz", "features", "opacity", "scaling", "rotation")


@pytest.fixture(scope="module")
def emul(host_lib):
    lib = ctypes.CDLL(host_lib)
    vp, i64, i32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32
    lib.emul_exclusive_scan.argtypes = [vp, vp, ctypes.c_size_t, vp, vp]
    lib.emul_scan_partials.restype = ctypes.c_size_t
    lib.emul_scan_partials.argtypes = [ctypes.c_size_t]
    lib.emul_sort_scratch_bytes.restype = ctypes.c_size_t
    lib.emul_sort_scratch_bytes.argtypes = [i64, i32]
    lib.emul_sort_pairs.argtypes = [vp, vp, vp, vp, i64, vp, i32, i32, vp, i32, ctypes.c_size_t, i32]
    return lib


def test_product_layer_still_refuses_cpu_tensors(host_lib):
    """Outside the on_host fixture nothing runs on the CPU."""
    import diff_gaussian_rasterization as dgr
    with pytest.raises(RuntimeError, match="no CPU"):
        dgr.knn_mean_dist2(torch.zeros(4, 3))


def test_scan_source_multi_tile(emul):
    """> 256 chunks of 2048: the partials kernel loops over more than one tile of 256."""
    n = 2048 * 258 + 77
    rng = np.random.default_rng(0)
    x = rng.integers(0, 3, n, dtype=np.uint32)
    out = np.zeros(n, dtype=np.uint32)
    partials = np.zeros(emul.emul_scan_partials(n), dtype=np.uint32)
    total = np.zeros(1, dtype=np.uint32)
    assert emul.emul_exclusive_scan(x.ctypes.data, out.ctypes.data, n, partials.ctypes.data, total.ctypes.data) == 0
    ref = np.concatenate(([0], np.cumsum(x, dtype=np.uint64)[:-1])).astype(np.uint32)
    assert np.array_equal(out, ref) and int(total[0]) == int(x.sum())
    small = x[:5000].copy()                                          # in place, on a prefix (one more full pass would double the test)
    assert emul.emul_exclusive_scan(small.ctypes.data, small.ctypes.data, 5000, partials.ctypes.data, total.ctypes.data) == 0
    assert np.array_equal(small, ref[:5000])


def test_kernel_sources_replay_reference_fixture(on_host):
    import model_replay as R
    R.replay(R.StoreDriver(R.load_gold(), GaussianModel, "cpu"))


def test_kernel_source_visible_mask(on_host):
    g = torch.Generator().manual_seed(3)
    P = 70
    m = GaussianModel(1).create_from_tensors(torch.randn(P, 3, generator=g), torch.randn(P, 1, 3, generator=g),
                                             torch.randn(P, 3, 3, generator=g), torch.randn(P, 3, generator=g) - 3,
                                             torch.randn(P, 4, generator=g), torch.randn(P, 1, generator=g))
    m.training_setup(SimpleNamespace(position_lr_init=1e-3, position_lr_final=1e-5, position_lr_delay_mult=0.01, position_lr_max_steps=100,
                                     feature_lr=0.0025, opacity_lr=0.025, scaling_lr=0.005, rotation_lr=0.001, percent_dense=0.01))
    m.grad.copy_(torch.randn(m.grad.shape, generator=g) * 1e-3)
    before = (m.store.clone(), m.exp_avg.clone(), m.act.clone())
    vis = torch.rand(P, generator=g) < 0.5
    m.optimizer_step(visible=vis)
    assert torch.equal(m._xyz[~vis], before[0][:3 * P].view(P, 3)[~vis]) and not torch.equal(m._xyz[vis], before[0][:3 * P].view(P, 3)[vis])
    assert torch.equal(m._features[~vis], before[0][3 * P:3 * P + 12 * P].view(P, 4, 3)[~vis])
    assert torch.equal(m._rotation[~vis], before[0][-4 * P:].view(P, 4)[~vis]) and not torch.equal(m._rotation[vis], before[0][-4 * P:].view(P, 4)[vis])
    assert torch.equal(m.act[:P][~vis], before[2][:P][~vis])
    assert float(m.exp_avg[:3 * P].view(P, 3)[~vis].abs().max()) == 0.0


def test_create_from_pcd_on_host(on_host):
    """gaussian_model.py:150-172 on the store: SH DC from colours, log-scales from the 3-nn distances, opacity 0.1."""
    from oracle.knn_oracle import mean_dist2_bruteforce
    g = torch.Generator().manual_seed(9)
    pts, cols = torch.rand(400, 3, generator=g) * 4 - 2, torch.rand(400, 3, generator=g)
    m = GaussianModel(3).create_from_pcd(pts, cols, spatial_lr_scale=2.5)
    assert m.P == 400 and m.spatial_lr_scale == 2.5
    torch.testing.assert_close(m._features_dc[:, 0], (cols - 0.5) / 0.28209479177387814)
    assert float(m._features_rest.abs().max()) == 0.0
    ref = torch.from_numpy(np.log(np.sqrt(np.maximum(mean_dist2_bruteforce(pts.numpy()), 1e-7)))).float()
    torch.testing.assert_close(m._scaling, ref[:, None].repeat(1, 3), rtol=1e-5, atol=1e-5)
    assert torch.equal(m._rotation, torch.tensor([1.0, 0, 0, 0]).repeat(400, 1))
    torch.testing.assert_close(m.get_opacity.detach(), torch.full((400, 1), 0.1))
    torch.testing.assert_close(m.get_scaling.detach(), torch.exp(m._scaling))


@pytest.mark.parametrize("kind", ["uniform", "clustered", "planar", "duplicates", "tiny"])
def test_knn_source_matches_bruteforce(on_host, kind):
    from oracle.knn_oracle import mean_dist2_bruteforce
    from simple_knn._C import distCUDA2
    r = np.random.default_rng(11)
    if kind == "uniform":
        pts = r.uniform(-3, 5, (1500, 3))
    elif kind == "clustered":           # SfM-like: dense blobs and far outliers
        pts = np.concatenate([r.normal(c, s, (400, 3)) for c, s in ((0, 0.01), (10, 1.0), (-50, 0.2))] + [r.uniform(-500, 500, (30, 3))])
    elif kind == "planar":              # zero extent along z: the grid must not degenerate
        pts = np.concatenate([r.uniform(0, 1, (900, 2)), np.full((900, 1), 0.25)], axis=1)
    elif kind == "duplicates":
        base = r.uniform(0, 1, (300, 3))
        pts = np.concatenate([base, base[:120], base[:40], np.zeros((5, 3))])
    else:
        pts = r.uniform(0, 1, (3, 3))
    pts = np.ascontiguousarray(pts, dtype=np.float32)
    out = distCUDA2(torch.from_numpy(pts)).numpy()
    np.testing.assert_allclose(out, mean_dist2_bruteforce(pts), rtol=2e-5, atol=1e-12)
    if kind == "tiny":                  # one point: zero; no points: empty
        assert distCUDA2(torch.from_numpy(pts[:1])).tolist() == [0.0] and distCUDA2(torch.zeros(0, 3)).numel() == 0


@pytest.mark.parametrize("n,bits,V,small", [
    (1, 32, 1, 0), (1000, 32, 1, 0), (5000, 13, 1, 0),     # 1024-key blocks, multi-block look-back
    (5000, 13, 3, 0),                                      # view batch: three independent sorts, ragged counts
    (40000, 13, 1, 0),                                     # 40 blocks: the eight-deep look-back window wraps
    (40000, 13, 2, -1),                                    # 16 keys per thread (large-input instantiation)
    (50000, 8, 1, -1),
])
def test_radix_sort_source_is_stable(emul, n, bits, V, small):
    """The hot path's sort (csrc/radix_sort.cu, both block sizes) on the host: stable order on the sorted bits, values
    carried, untouched tails when the per-view count is below the launch size."""
    r = np.random.default_rng(n + bits)
    sv = n + 37                                                       # stride between the views' arrays
    keys = r.integers(0, 2 ** 32, V * sv, dtype=np.uint64).astype(np.uint32)
    if bits < 32:
        keys &= np.uint32((1 << bits) - 1) | np.uint32(0xF0000000)    # high garbage bits must be ignored
    vals = np.arange(V * sv, dtype=np.uint32)
    counts = np.array([n if v == 0 else max(1, n - 11 * v) for v in range(V)], dtype=np.uint64)
    k, v_ = keys.copy(), vals.copy()
    ka, va = np.zeros_like(k), np.zeros_like(v_)
    scratch = np.zeros(emul.emul_sort_scratch_bytes(n, V), dtype=np.uint8)
    rc = emul.emul_sort_pairs(k.ctypes.data, v_.ctypes.data, ka.ctypes.data, va.ctypes.data, n, counts.ctypes.data if V > 1 else None,
                              0, bits, scratch.ctypes.data, V, sv, small)
    assert rc == 0
    mask = np.uint32((1 << bits) - 1) if bits < 32 else np.uint32(0xFFFFFFFF)
    for view in range(V):
        cnt = int(counts[view]) if V > 1 else n
        seg = slice(view * sv, view * sv + cnt)
        order = np.argsort(keys[seg] & mask, kind="stable")
        assert np.array_equal(k[seg], keys[seg][order]), view
        assert np.array_equal(v_[seg], vals[seg][order]), view


def test_checkpoint_contract_of_the_reference(on_host):
    """capture() / restore(model_args, training_args) speak the reference's checkpoint format (gaussian_model.py:63-99): the tuple
    written by this class reloads into it bit for bit (including per-group step counts), and a tuple whose optimizer part is a REAL
    torch.optim.Adam.state_dict() -- what the reference's own capture() writes -- loads into the flat moments."""
    from oracle.model_oracle import ModelOracle
    g = torch.Generator().manual_seed(9)
    P = 70
    raw = {"xyz": torch.randn(P, 3, generator=g), "f_dc": torch.randn(P, 1, 3, generator=g), "f_rest": torch.randn(P, 15, 3, generator=g) * 0.1,
           "opacity": torch.randn(P, 1, generator=g), "scaling": torch.randn(P, 3, generator=g) - 3, "rotation": torch.randn(P, 4, generator=g)}
    opt = dict(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=30000,
               feature_lr=0.0025, opacity_lr=0.025, scaling_lr=0.005, rotation_lr=0.001, percent_dense=0.01)
    args = SimpleNamespace(**opt)
    m = GaussianModel(3).create_from_tensors(raw["xyz"], raw["f_dc"], raw["f_rest"], raw["scaling"], raw["rotation"], raw["opacity"], 2.0)
    m.training_setup(args)
    for it in (1, 2, 3):
        m.update_learning_rate(it)
        m.grad.copy_(torch.randn(m.grad.shape, generator=g) * 1e-3)
        m.gradients_ready()
        if it == 3:
            m.reset_opacity()                      # opacity misses this step: its step count lags
        m.optimizer_step()
    ck = m.capture()
    assert len(ck) == 12 and set(ck[10]) == {"state", "param_groups"} and [gr["name"] for gr in ck[10]["param_groups"]] == list(GROUPS)
    m2 = GaussianModel(3).restore(ck, args)
    assert m2.group_steps == m.group_steps == {"xyz": 3, "f_dc": 3, "f_rest": 3, "opacity": 2, "scaling": 3, "rotation": 3}
    assert torch.equal(m2.store, m.store) and torch.equal(m2.exp_avg, m.exp_avg) and torch.equal(m2.exp_avg_sq, m.exp_avg_sq)
    assert torch.equal(m2.act, m.act) and m2.active_sh_degree == m.active_sh_degree and m2.spatial_lr_scale == 2.0
    # a checkpoint as the reference writes it: parameters + a real Adam state_dict
    ref = ModelOracle(raw["xyz"], raw["f_dc"], raw["f_rest"], raw["opacity"], raw["scaling"], raw["rotation"], opt, 2.0)
    for it in (1, 2):
        ref.step(it, {"xyz": torch.randn(P, 3, generator=g) * 1e-3, "features": torch.randn(P, 16, 3, generator=g) * 1e-3,
                      "opacity": torch.randn(P, 1, generator=g) * 1e-3, "scaling": torch.randn(P, 3, generator=g) * 1e-3,
                      "rotation": torch.randn(P, 4, generator=g) * 1e-3})
    p = {n: ref.p[n].detach() for n in GROUPS}
    tup = (2, p["xyz"], p["f_dc"], p["f_rest"], p["scaling"], p["rotation"], p["opacity"], torch.zeros(P), torch.zeros(P, 1), torch.zeros(P, 1),
           ref.adam.state_dict(), 2.0)
    m3 = GaussianModel(3).restore(tup, args)
    mom = ref.moments()
    views_m, views_v = m3._group_views(m3.exp_avg), m3._group_views(m3.exp_avg_sq)
    for n in GROUPS:
        assert torch.equal(views_m[n], mom[n]["m"]) and torch.equal(views_v[n], mom[n]["v"]) and m3.group_steps[n] == 2
    assert m3.active_sh_degree == 2 and torch.equal(m3._xyz, p["xyz"])


def test_exposure_parameters_follow_the_reference_class(on_host, tmp_path):
    """Per-image exposures of gaussian_store.GaussianModel (create_exposures / get_exposure_from_name / exposure optimizer and
    its learning-rate schedule / exposure.json) against a fixture recorded from the reference's OWN scene/gaussian_model.py
    stepped as train.py:178-179 does (tests/golden/make_golden_exposure.py); the colour transform is the expression of the
    reference's gaussian_renderer/__init__.py:113-115, which gaussian_renderer.render(use_trained_exp=True) applies."""
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_exposure.npz"))
    names = [str(n) for n in gold["names"]]
    lr_init, lr_final, delay_steps, delay_mult, iterations = (float(x) for x in gold["sched"])
    args = SimpleNamespace(position_lr_init=0.00016, position_lr_final=0.0000016, position_lr_delay_mult=0.01, position_lr_max_steps=30000,
                           feature_lr=0.0025, opacity_lr=0.025, scaling_lr=0.005, rotation_lr=0.001, percent_dense=0.01,
                           exposure_lr_init=lr_init, exposure_lr_final=lr_final, exposure_lr_delay_steps=delay_steps,
                           exposure_lr_delay_mult=delay_mult, iterations=iterations)
    g = torch.Generator().manual_seed(1)
    pts, cols = torch.rand(40, 3, generator=g), torch.rand(40, 3, generator=g)
    m = GaussianModel(1).create_from_pcd(pts, cols, spatial_lr_scale=1.0, cam_infos=[SimpleNamespace(image_name=n) for n in names])
    assert m.exposure_mapping == {n: i for i, n in enumerate(names)} and m.pretrained_exposures is None
    assert torch.equal(m.get_exposure, torch.eye(3, 4)[None].repeat(len(names), 1, 1))
    m.training_setup(args)
    imgs, tgts = torch.from_numpy(gold["imgs"]), torch.from_numpy(gold["tgts"])
    for k, it in enumerate(gold["iters"].tolist()):
        m.update_learning_rate(it)
        assert abs(m.exposure_optimizer.param_groups[0]["lr"] - float(gold["lrs"][k])) <= 1e-12 * float(gold["lrs"][k]) + 1e-18
        i = int(gold["order"][k])
        exposure = m.get_exposure_from_name(names[i])
        img = torch.matmul(imgs[i].permute(1, 2, 0), exposure[:3, :3]).permute(2, 0, 1) + exposure[:3, 3, None, None]
        (img - tgts[i]).abs().mean().backward()
        m.exposure_optimizer.step()
        m.exposure_optimizer.zero_grad(set_to_none=True)
        assert torch.allclose(m.get_exposure.detach(), torch.from_numpy(gold["exposures"][k]), rtol=0, atol=1e-7), k
    # exposure.json: written as Scene.save does, read back as load_ply(use_train_test_exp=True) does (two levels above the ply)
    ply_dir = tmp_path / "point_cloud" / "iteration_7"
    ply_dir.mkdir(parents=True)
    m.save_exposures(str(tmp_path / "exposure.json"))
    m.save_ply(str(ply_dir / "point_cloud.ply"))
    m2 = GaussianModel(1).load_ply(str(ply_dir / "point_cloud.ply"), device="cpu", use_train_test_exp=True)
    assert set(m2.pretrained_exposures) == set(names)
    for n in names:
        assert torch.allclose(m2.get_exposure_from_name(n), m.get_exposure_from_name(n).detach(), rtol=0, atol=1e-7)
        assert not m2.get_exposure_from_name(n).requires_grad
    m3 = GaussianModel(1).load_ply(str(ply_dir / "point_cloud.ply"), device="cpu")          # without the flag: nothing loaded
    assert m3.pretrained_exposures is None
    with pytest.raises(RuntimeError, match="no exposures"):
        m3.get_exposure_from_name(names[0])
