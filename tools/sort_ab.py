"""A/B of alternative builds of the library on the tile sort (gsb_sort_pairs, 13-bit keys = two onesweep passes) and the
depth sort (32-bit keys = four passes): one subprocess per library (GSB_LIBRARY is read at import).

    python tools/sort_ab.py gaussian-splatting_b200/libgs_b200_base.so gaussian-splatting_b200/libgs_b200.so ...
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r'''
import os, sys, ctypes, torch
sys.path.insert(0, os.path.join(%(root)r, "gaussian-splatting_b200"))
import diff_gaussian_rasterization as dgr
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
res = []
for n, bits in ((4_400_000, 13), (35_000_000, 13), (1_000_000, 32)):
    keys0 = torch.randint(0, 2 ** min(bits, 31), (n,), device=dev, generator=g, dtype=torch.int64).to(torch.int32)
    vals0 = torch.arange(n, device=dev, dtype=torch.int32)
    def run():
        k, v = keys0.clone(), vals0.clone()
        arena = dgr._Arena(dev, torch.cuda.current_stream().cuda_stream)
        rc = dgr._C.gsb_sort_pairs(k.data_ptr(), v.data_ptr(), n, 0, bits, arena.cb, None, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        return k, v
    k, v = run()
    ks, order = torch.sort(keys0.to(torch.int64) & 0xffffffff, stable=True)
    ok = bool(torch.equal(k.to(torch.int64) & 0xffffffff, ks)) and bool(torch.equal(v.to(torch.int64), order))
    dgr.set_option("time_kernels", 2)
    dgr.kernel_time("", reset=True)
    for _ in range(10):
        run()
    torch.cuda.synchronize()
    ms, cnt = dgr.kernel_time("sort_scatter")
    dgr.set_option("time_kernels", 0)
    res.append(f"n={n} bits={bits}: {'OK ' if ok else 'WRONG '} {ms / cnt * 1e3:.1f} us/pass ({n * 16 / (ms / cnt) / 1e6:.0f} GB/s)")
print(" | ".join(res))
'''

for lib in sys.argv[1:]:
    env = dict(os.environ, GSB_LIBRARY=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=120)
    print(f"{os.path.basename(lib):<26} {out.stdout.strip() or out.stderr.strip()[-300:]}")
