"""ctypes driver for oracle/libgs_oracle.so (gs_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this.  See gs_oracle.c for the parity status ("parity unpinned" for the splatting
rules, pinned SH / covariance / camera formulae).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_float, c_int, c_int64, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgs_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "gs_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B", "libgs_oracle.so"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        fp = POINTER(c_float)
        L.gso_forward.restype = c_void_p
        L.gso_forward.argtypes = [c_int, c_int, c_int, fp, fp, fp, fp, fp, fp, fp, c_float, fp, fp, fp, fp,
                                  c_int, c_int, c_float, c_float, c_int, c_int, c_int,
                                  fp, POINTER(c_int), fp]
        L.gso_backward.restype = None
        L.gso_backward.argtypes = [c_void_p] + [fp] * 10
        L.gso_free.restype = None
        L.gso_free.argtypes = [c_void_p]
        L.gso_num_rendered.restype = c_int64
        L.gso_num_rendered.argtypes = [c_void_p]
        L.gso_final_T.restype = fp
        L.gso_final_T.argtypes = [c_void_p]
        L.gso_n_contrib.restype = POINTER(c_int)
        L.gso_n_contrib.argtypes = [c_void_p]
        L.gso_threads.restype = c_int
        L.gso_set_threads.restype = None
        L.gso_set_threads.argtypes = [c_int]
        _lib = L
    return _lib


def _f32(x):
    if x is None:
        return None
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(x, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(POINTER(c_float)) if a is not None else None


class COracle:
    """One forward (+ optional backward) of the CPU oracle.  Arguments mirror
    GaussianRasterizer.forward + GaussianRasterizationSettings (gaussian_renderer/__init__.py:36-110)."""

    def __init__(self, means3D, shs, colors_precomp, opacities, scales, rotations, cov3D_precomp, settings,
                 tile_rows=None):
        L = lib()
        self.s = settings
        self.P = int(means3D.shape[0])
        self.means = _f32(means3D)
        self.shs = _f32(shs)
        self.colors = _f32(colors_precomp)
        self.opac = _f32(opacities).reshape(-1)
        self.scales = _f32(scales)
        self.rots = _f32(rotations)
        self.cov = _f32(cov3D_precomp)
        self.M = int(self.shs.shape[1]) if self.shs is not None else 0
        H, W = int(settings.image_height), int(settings.image_width)
        self.H, self.W = H, W
        self.view = _f32(settings.viewmatrix).reshape(-1)
        self.proj = _f32(settings.projmatrix).reshape(-1)
        self.campos = _f32(settings.campos).reshape(-1)
        self.bg = _f32(settings.bg).reshape(-1)
        self.color = np.zeros((3, H, W), np.float32)
        self.radii = np.zeros(self.P, np.int32)
        self.invdepth = np.zeros((1, H, W), np.float32)
        ty0, ty1 = tile_rows if tile_rows is not None else (0, 0)
        self.h = L.gso_forward(self.P, self.M, int(settings.sh_degree), _p(self.means), _p(self.shs),
                               _p(self.colors), _p(self.opac), _p(self.scales), _p(self.rots), _p(self.cov),
                               float(settings.scale_modifier), _p(self.view), _p(self.proj), _p(self.campos),
                               _p(self.bg), W, H, float(settings.tanfovx), float(settings.tanfovy),
                               int(bool(settings.antialiasing)), int(ty0), int(ty1),
                               _p(self.color), self.radii.ctypes.data_as(POINTER(c_int)), _p(self.invdepth))
        self.num_rendered = int(L.gso_num_rendered(self.h))

    def final_T(self):
        return np.ctypeslib.as_array(lib().gso_final_T(self.h), shape=(self.H, self.W)).copy()

    def n_contrib(self):
        return np.ctypeslib.as_array(lib().gso_n_contrib(self.h), shape=(self.H, self.W)).copy()

    def backward(self, dL_dcolor, dL_dinvdepth=None):
        P, M = self.P, self.M
        dc = _f32(dL_dcolor)
        dd = _f32(dL_dinvdepth)
        g = dict(means3D=np.zeros((P, 3), np.float32), means2D=np.zeros((P, 3), np.float32),
                 opacities=np.zeros((P, 1), np.float32))
        g["shs"] = np.zeros((P, M, 3), np.float32) if self.shs is not None else None
        g["colors_precomp"] = np.zeros((P, 3), np.float32) if self.colors is not None else None
        g["scales"] = np.zeros((P, 3), np.float32) if self.scales is not None else None
        g["rotations"] = np.zeros((P, 4), np.float32) if self.rots is not None else None
        g["cov3D_precomp"] = np.zeros((P, 6), np.float32) if self.cov is not None else None
        lib().gso_backward(self.h, _p(dc), _p(dd), _p(g["means3D"]), _p(g["means2D"]), _p(g["shs"]),
                           _p(g["colors_precomp"]), _p(g["opacities"]), _p(g["scales"]), _p(g["rotations"]),
                           _p(g["cov3D_precomp"]))
        return g

    def close(self):
        if self.h:
            lib().gso_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def threads() -> int:
    return int(lib().gso_threads())


def set_threads(n: int) -> int:
    """OpenMP thread count of the oracle's loops (torchrun exports OMP_NUM_THREADS=1); returns the count in effect."""
    lib().gso_set_threads(int(n))
    return threads()
