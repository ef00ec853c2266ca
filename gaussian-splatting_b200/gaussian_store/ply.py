"""The gaussians' on-disk format (point_cloud.ply), read and written with numpy only.

Layout follows /root/reference/scene/gaussian_model.py: one ``vertex`` element whose float32 properties are, in order
(:225-237), ``x y z nx ny nz f_dc_0..2 f_rest_0..(3(M-1)-1) opacity scale_0..2 rot_0..3``; values are the RAW parameters
(log scale, logit opacity, unnormalised quaternion); normals are zeros; the SH arrays are stored CHANNEL-major
(``features.transpose(1, 2).flatten(1)``, :243-244), i.e. ``f_rest_{c*(M-1)+k}`` is coefficient k+1 of colour channel c.
The reference writes it through the ``plyfile`` package (absent from this image), whose default output for such an element
is ``format binary_little_endian 1.0`` with ``property float <name>`` lines [RECALL for the exact header text]; the reader
here accepts that and the ASCII flavour, any property order, and float / double properties (:263-314 looks properties up
by name).
"""
from __future__ import annotations

import os
from typing import Dict, List

import numpy as np

_PLY_TYPES = {"float": "f4", "float32": "f4", "double": "f8", "float64": "f8", "uchar": "u1", "uint8": "u1", "char": "i1", "int8": "i1",
              "short": "i2", "int16": "i2", "ushort": "u2", "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4"}


def attribute_names(sh_coeffs: int) -> List[str]:
    names = ["x", "y", "z", "nx", "ny", "nz"] + [f"f_dc_{i}" for i in range(3)]
    names += [f"f_rest_{i}" for i in range(3 * (sh_coeffs - 1))]
    return names + ["opacity"] + [f"scale_{i}" for i in range(3)] + [f"rot_{i}" for i in range(4)]


def write_gaussian_ply(path: str, xyz, features_dc, features_rest, opacity, scaling, rotation) -> None:
    """Arrays as in the reference's tensors: xyz [P,3], features_dc [P,1,3], features_rest [P,M-1,3], opacity [P,1],
    scaling [P,3], rotation [P,4]."""
    xyz = np.asarray(xyz, dtype=np.float32)
    P = xyz.shape[0]
    f_dc = np.asarray(features_dc, dtype=np.float32).reshape(P, -1, 3).transpose(0, 2, 1).reshape(P, -1)
    f_rest = np.asarray(features_rest, dtype=np.float32).reshape(P, -1, 3).transpose(0, 2, 1).reshape(P, -1)
    M = 1 + f_rest.shape[1] // 3
    cols = np.concatenate((xyz, np.zeros_like(xyz), f_dc, f_rest, np.asarray(opacity, dtype=np.float32).reshape(P, 1),
                           np.asarray(scaling, dtype=np.float32).reshape(P, 3), np.asarray(rotation, dtype=np.float32).reshape(P, 4)), axis=1)
    names = attribute_names(M)
    assert cols.shape[1] == len(names)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {P}"] + [f"property float {n}" for n in names] + ["end_header"]
    d = os.path.dirname(path)
    if d:
        os.makedirs(d, exist_ok=True)
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(np.ascontiguousarray(cols, dtype="<f4").tobytes())


def _read_vertex_table(path: str) -> Dict[str, np.ndarray]:
    with open(path, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{path}: not a PLY file")
        fmt, elements, cur = None, [], None
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{path}: header is not terminated")
            tok = line.decode("ascii").split()
            if not tok or tok[0] in ("comment", "obj_info"):
                continue
            if tok[0] == "format":
                fmt = tok[1]
            elif tok[0] == "element":
                cur = {"name": tok[1], "count": int(tok[2]), "props": []}
                elements.append(cur)
            elif tok[0] == "property":
                if tok[1] == "list":
                    raise ValueError(f"{path}: list properties are not part of the gaussian format")
                cur["props"].append((tok[2], _PLY_TYPES[tok[1]]))
            elif tok[0] == "end_header":
                break
        if not elements or elements[0]["name"] != "vertex":
            raise ValueError(f"{path}: the first element must be 'vertex'")
        el = elements[0]
        if fmt == "ascii":
            rows = np.loadtxt(f, dtype=np.float64, max_rows=el["count"], ndmin=2)
            return {n: rows[:, i] for i, (n, _) in enumerate(el["props"])}
        if fmt not in ("binary_little_endian", "binary_big_endian"):
            raise ValueError(f"{path}: unknown PLY format {fmt}")
        order = "<" if fmt == "binary_little_endian" else ">"
        dt = np.dtype([(n, order + t) for n, t in el["props"]])
        table = np.frombuffer(f.read(dt.itemsize * el["count"]), dtype=dt, count=el["count"])
        return {n: table[n] for n, _ in el["props"]}


def read_gaussian_ply(path: str, max_sh_degree: int) -> Dict[str, np.ndarray]:
    """Returns float32 arrays shaped like the reference's parameters (load_ply :278-312)."""
    t = _read_vertex_table(path)
    M = (max_sh_degree + 1) ** 2
    num = lambda prefix: sorted((n for n in t if n.startswith(prefix)), key=lambda n: int(n.split("_")[-1]))
    rest_names = num("f_rest_")
    if len(rest_names) != 3 * M - 3:
        raise ValueError(f"{path}: {len(rest_names)} f_rest properties, expected {3 * M - 3} for SH degree {max_sh_degree}")
    col = lambda names: np.stack([np.asarray(t[n], dtype=np.float32) for n in names], axis=1)
    P = len(t["x"])
    f_dc = col(["f_dc_0", "f_dc_1", "f_dc_2"]).reshape(P, 3, 1).transpose(0, 2, 1)
    f_rest = col(rest_names).reshape(P, 3, M - 1).transpose(0, 2, 1) if M > 1 else np.zeros((P, 0, 3), dtype=np.float32)
    return {"xyz": col(["x", "y", "z"]), "features_dc": np.ascontiguousarray(f_dc), "features_rest": np.ascontiguousarray(f_rest),
            "opacity": col(["opacity"]), "scaling": col(num("scale_")), "rotation": col(num("rot"))}
