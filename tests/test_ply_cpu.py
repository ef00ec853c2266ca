"""point_cloud.ply reader / writer (gaussian_store/ply.py) against the layout of the reference's save_ply / load_ply
(/root/reference/scene/gaussian_model.py:225-314): attribute order, channel-major SH flattening, round trip, ASCII input."""
import numpy as np
import pytest

from gaussian_store.ply import attribute_names, read_gaussian_ply, write_gaussian_ply


def _arrays(P, M, seed=0):
    r = np.random.default_rng(seed)
    f = lambda *s: r.standard_normal(s).astype(np.float32)
    return dict(xyz=f(P, 3), features_dc=f(P, 1, 3), features_rest=f(P, M - 1, 3), opacity=f(P, 1), scaling=f(P, 3), rotation=f(P, 4))


def test_attribute_order_matches_reference_list():
    n = attribute_names(16)
    assert n[:9] == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    assert n[9] == "f_rest_0" and n[53] == "f_rest_44" and n[54] == "opacity"
    assert n[55:] == ["scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"] and len(n) == 62
    assert len(attribute_names(1)) == 17


@pytest.mark.parametrize("deg", [0, 1, 3])
def test_round_trip_and_binary_layout(tmp_path, deg):
    M, P = (deg + 1) ** 2, 37
    a = _arrays(P, M, seed=deg)
    path = str(tmp_path / "pc" / "point_cloud.ply")
    write_gaussian_ply(path, **a)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", f"element vertex {P}"]
    assert lines[3:] == [f"property float {n}" for n in attribute_names(M)]
    table = np.frombuffer(body, dtype="<f4").reshape(P, len(attribute_names(M)))
    assert np.array_equal(table[:, :3], a["xyz"]) and not table[:, 3:6].any()
    assert np.array_equal(table[:, 6:9], a["features_dc"][:, 0, :])
    if M > 1:   # channel-major: f_rest_{c*(M-1)+k} = features_rest[:, k, c]   (transpose(1, 2).flatten(1), :244)
        assert np.array_equal(table[:, 9 + 1 * (M - 1) + 2], a["features_rest"][:, 2, 1])
    back = read_gaussian_ply(path, deg)
    for k, v in a.items():
        assert back[k].shape == v.shape and back[k].dtype == np.float32 and np.array_equal(back[k], v), k


def test_reads_ascii_and_reordered_double_properties(tmp_path):
    names = attribute_names(1)
    order = names[::-1]
    vals = {n: float(i) + 0.5 for i, n in enumerate(names)}
    path = tmp_path / "a.ply"
    path.write_text("\n".join(["ply", "format ascii 1.0", "comment made by hand", "element vertex 2"] +
                              [f"property double {n}" for n in order] + ["end_header"] +
                              [" ".join(repr(vals[n] + r) for n in order) for r in (0.0, 100.0)]) + "\n")
    g = read_gaussian_ply(str(path), 0)
    assert g["xyz"].tolist() == [[0.5, 1.5, 2.5], [100.5, 101.5, 102.5]]
    assert g["features_dc"].shape == (2, 1, 3) and g["features_dc"][0, 0].tolist() == [6.5, 7.5, 8.5]
    assert g["opacity"][1, 0] == 109.5 and g["scaling"][0].tolist() == [10.5, 11.5, 12.5] and g["rotation"][0].tolist() == [13.5, 14.5, 15.5, 16.5]
    with pytest.raises(ValueError, match="f_rest"):
        read_gaussian_ply(str(path), 3)
