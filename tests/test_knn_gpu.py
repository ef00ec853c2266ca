"""GPU parity of gsb_knn_mean_dist2 (the stand-in for simple_knn._C.distCUDA2, scene/gaussian_model.py:159) against
oracle/knn_oracle.py; float32 distances, tolerance 2e-5 relative."""
import numpy as np
import pytest
import torch

from oracle.knn_oracle import mean_dist2_bruteforce, mean_dist2_kdtree

pytestmark = pytest.mark.gpu


def _clouds():
    r = np.random.default_rng(5)
    yield "uniform_200k", r.uniform(-4, 4, (200_000, 3)), mean_dist2_kdtree
    blobs = [r.normal(c, s, (20_000, 3)) for c, s in ((0, 0.01), (10, 1.0), (-50, 0.2), (3, 5.0))]
    yield "sfm_like", np.concatenate(blobs + [r.uniform(-500, 500, (500, 3))]), mean_dist2_kdtree
    yield "planar", np.concatenate([r.uniform(0, 1, (3000, 2)), np.full((3000, 1), 0.25)], axis=1), mean_dist2_bruteforce
    base = r.uniform(0, 1, (1000, 3))
    yield "duplicates", np.concatenate([base, base[:400], base[:100], np.zeros((7, 3))]), mean_dist2_bruteforce
    yield "three", r.uniform(0, 1, (3, 3)), mean_dist2_bruteforce
    yield "one", r.uniform(0, 1, (1, 3)), mean_dist2_bruteforce


@pytest.mark.parametrize("name,pts,oracle", list(_clouds()), ids=lambda v: v if isinstance(v, str) else "")
def test_knn_matches_oracle(name, pts, oracle):
    from simple_knn._C import distCUDA2            # the import line of the reference (gaussian_model.py:21)
    p = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float32)).cuda()
    out = distCUDA2(p)
    assert out.shape == (len(pts),) and out.dtype == torch.float32
    ref = oracle(np.asarray(pts, dtype=np.float32))
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=2e-5, atol=1e-12)
    # the reference's use of it: clamp, sqrt, log -> initial log-scales (gaussian_model.py:159-160) must be finite
    scales = torch.log(torch.sqrt(torch.clamp_min(out, 0.0000001)))
    assert bool(torch.isfinite(scales).all())


def test_knn_empty():
    import diff_gaussian_rasterization as dgr
    assert dgr.knn_mean_dist2(torch.zeros(0, 3, device="cuda")).numel() == 0
