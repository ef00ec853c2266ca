"""``distCUDA2(points[P,3] float32 cuda) -> [P]``: mean squared distance to the three nearest neighbours, as used at
/root/reference/scene/gaussian_model.py:159 to initialise the scales."""
import diff_gaussian_rasterization as _dgr


def distCUDA2(points):
    return _dgr.knn_mean_dist2(points)
