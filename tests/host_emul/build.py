"""TEST INFRASTRUCTURE.  Compiles the UNMODIFIED sources of gaussian-splatting_b200/csrc for the host (g++ -DGSB_HOST_EMUL,
tests/host_emul/cuda_shim.h standing in for the CUDA language and runtime) into a throw-away shared library that exports
the same C ABI as libgs_b200.so, and points the Python layer at it for the duration of a test.

Nothing here is a product path: the shipped library is CUDA only and the Python layer refuses CPU tensors; the tests swap
the library handle and the three CUDA touch points (`_require_cuda`, `_current_stream`, `_device_ctx`) explicitly."""
import contextlib
import glob
import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "gaussian-splatting_b200", "csrc")
HERE = os.path.dirname(os.path.abspath(__file__))


def build_library(out_dir: str) -> str:
    gxx = shutil.which("g++", path="/usr/bin") or shutil.which("g++")
    if gxx is None:
        raise RuntimeError("no host C++ compiler")
    os.makedirs(out_dir, exist_ok=True)
    flags = ["-std=c++20", "-O1", "-fPIC", "-pthread", "-DGSB_HOST_EMUL", "-I", HERE, "-I", CSRC, "-w"]
    jobs = [(src, os.path.join(out_dir, os.path.basename(src) + ".o"), ["-x", "c++"]) for src in sorted(glob.glob(os.path.join(CSRC, "*.cu")))]
    for extra in ("emul_runtime.cpp", "emul_extra.cpp"):
        jobs.append((os.path.join(HERE, extra), os.path.join(out_dir, extra + ".o"), []))

    def compile_one(job):
        src, obj, lang = job
        r = subprocess.run([gxx, *flags, *lang, "-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode:
            raise RuntimeError(f"host build of {os.path.basename(src)} failed:\n{r.stderr[-4000:]}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as pool:
        objs = list(pool.map(compile_one, jobs))
    lib = os.path.join(out_dir, "libgs_b200_hostemul.so")
    subprocess.run([gxx, "-shared", "-pthread", "-o", lib, *objs], check=True, capture_output=True, text=True)
    return lib


@contextlib.contextmanager
def python_layer_on_host(lib_path: str):
    """The whole Python stack (rasterizer autograd function, view-batch step, GaussianModel, fused_ssim, simple_knn) driven by
    CPU tensors against the host build.  Restores the product bindings on exit."""
    import diff_gaussian_rasterization as dgr
    saved = (dgr._C, dgr._require_cuda, dgr._current_stream, dgr._device_ctx, dict(dgr._capacity_hints), dgr.speculative_binning)
    dgr._C = dgr._load(lib_path)
    dgr._require_cuda = lambda t: None
    dgr._current_stream = lambda device: 0
    dgr._device_ctx = lambda device: contextlib.nullcontext()
    dgr._capacity_hints.clear()
    # exact instance counts: the speculative capacity is rounded up to 2^20 instances, i.e. thousands of empty blocks per
    # launch -- free on a GPU, seconds each here (the repair / speculation logic itself is covered by the -m gpu tests)
    dgr.speculative_binning = False
    try:
        yield dgr
    finally:
        dgr._C, dgr._require_cuda, dgr._current_stream, dgr._device_ctx = saved[:4]
        dgr._capacity_hints.clear()
        dgr._capacity_hints.update(saved[4])
        dgr.speculative_binning = saved[5]
        dgr.release_scratch()
