"""CPU tests: the C-ABI library loads and exports every symbol include/*.h declares (no compute calls)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gaussian-splatting_b200", "libgs_b200.so")


def _declared_functions():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(", src))
    names.discard("gsb_alloc_fn")
    return sorted(names)


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(LIB)
    decl = _declared_functions()
    assert len(decl) >= 8
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    lib.gsb_abi_version.restype = ctypes.c_int32
    assert lib.gsb_abi_version() == 6


def test_struct_sizes_match_header():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
    import diff_gaussian_rasterization as dgr
    # LP64 layouts of include/gs_b200.h
    assert ctypes.sizeof(dgr._Settings) == 80
    assert ctypes.sizeof(dgr._Inputs) == 64
    assert ctypes.sizeof(dgr._State) == 128
    assert ctypes.sizeof(dgr._Grads) == 64
    assert ctypes.sizeof(dgr._AdamArgs) == 128
    assert ctypes.sizeof(dgr._DensifyArgs) == 80


def test_ctypes_mirrors_agree_with_the_c_compiler(tmp_path):
    """sizeof / offsetof of the argument structs as gcc sees include/gs_b200.h vs the ctypes mirrors."""
    import shutil
    import subprocess
    import sys
    import pytest
    gcc = shutil.which("gcc", path="/usr/bin") or shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
    import diff_gaussian_rasterization as dgr
    probes = {"GsbAdamArgs": (dgr._AdamArgs, ["P", "skip_groups", "params", "visible", "step_size", "bias2_sqrt", "beta1", "eps"]),
              "GsbDensifyArgs": (dgr._DensifyArgs, ["P", "n_children", "params", "denom", "grad_threshold", "world_limit", "scratch"]),
              "GsbSettings": (dgr._Settings, []), "GsbInputs": (dgr._Inputs, []), "GsbState": (dgr._State, []), "GsbGrads": (dgr._Grads, [])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gs_b200.h"', 'int main(void) {']
    for cname, (_, fields) in probes.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0; }")
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run([gcc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, (ct, fields) in probes.items():
        assert int(got[cname]) == ctypes.sizeof(ct), cname
        for f in fields:
            assert int(got[f"{cname}.{f}"]) == getattr(ct, f).offset, (cname, f)


def test_argument_errors_are_reported_not_crashed():
    """Host-side validation happens before any CUDA call, so it is testable without a GPU."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
    import torch
    import diff_gaussian_rasterization as dgr
    rs = dgr.GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4), torch.eye(4), 0,
                                           torch.zeros(3), False, False, False)
    r = dgr.GaussianRasterizer(rs)
    import pytest
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(means3D=torch.zeros(1, 3), means2D=torch.zeros(1, 3), opacities=torch.zeros(1, 1), scales=torch.ones(1, 3),
          rotations=torch.ones(1, 4))
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(means3D=torch.zeros(1, 3), means2D=torch.zeros(1, 3), opacities=torch.zeros(1, 1), shs=torch.zeros(1, 1, 3))
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(means3D=torch.zeros(1, 3), means2D=torch.zeros(1, 3), opacities=torch.zeros(1, 1), shs=torch.zeros(1, 1, 3),
          scales=torch.ones(1, 3), rotations=torch.ones(1, 4))
