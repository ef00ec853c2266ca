"""How much does sampling clocks through NVML disturb a launch-heavy CUDA loop?  (GPU box only)"""
import os, sys, time, threading, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
import torch, pynvml
import bench
from gaussian_renderer.synthetic import make_scene
import diff_gaussian_rasterization as dgr
from gaussian_renderer import GradientBucket, render_views_backward
dev = torch.device("cuda", 0)
P, W, H = 1_000_000, 1920, 1080
scene = make_scene(P, seed=0, log_scale_mean=bench.LOG_SCALE_MEAN)
pc = bench.BenchGaussians(scene, 3, dev); bucket = GradientBucket(pc.parameters())
bg = torch.zeros(3, device=dev)
cams = [bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(i, 3.0), dev) for i in range(8)]
gts = [torch.rand(3, H, W, device=dev) for _ in range(8)]
def step():
    bucket.zero_()
    render_views_backward(cams, pc, bench.Pipe(), bg, lambda img, d, i: dgr.l1_loss_and_grad(img, gts[i]), loss_returns_grad=True)
def run(n=10):
    torch.cuda.synchronize(); t = time.time()
    for _ in range(n): step()
    torch.cuda.synchronize(); return (time.time() - t) / n * 1e3
for _ in range(3): step()
print("no sampling: %.2f ms/step" % run())
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
def sampler(fn, period, flag, durs):
    while not flag[0]:
        t = time.time(); fn(); durs.append(time.time() - t); time.sleep(period)
tests = {
 "clockinfo": lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
 "reasons": lambda: pynvml.nvmlDeviceGetCurrentClocksEventReasons(h),
 "clock_current": lambda: pynvml.nvmlDeviceGetClock(h, pynvml.NVML_CLOCK_SM, pynvml.NVML_CLOCK_ID_CURRENT),
 "sleep_only": lambda: None,
}
for name, fn in tests.items():
    for period in (0.1,):
        flag, durs = [False], []
        th = threading.Thread(target=sampler, args=(fn, period, flag, durs), daemon=True); th.start()
        ms = run(); flag[0] = True; th.join()
        print("%-14s period %.2fs: %.2f ms/step, %d samples, query avg %.2f ms max %.2f ms" % (name, period, ms, len(durs), 1e3 * sum(durs) / max(1, len(durs)), 1e3 * max(durs or [0])))
print("no sampling again: %.2f ms/step" % run())

print("---- per-step pattern ----")
def seq(n, label):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t = time.time(); step(); torch.cuda.synchronize(); ts.append((time.time() - t) * 1e3)
    clk = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
    pw = pynvml.nvmlDeviceGetPowerUsage(h) / 1e3
    print(label, "min %.1f med %.1f max %.1f | sm %d MHz reasons 0x%x power %.0f W |" % (min(ts), sorted(ts)[len(ts)//2], max(ts), clk, r, pw), " ".join("%.0f" % x for x in ts))
seq(30, "plain      ")
dgr.set_option("time_kernels", 1)
seq(30, "time_kernels")
print("events recorded:", dgr.kernel_time("render_bwd", reset=True))
dgr.set_option("time_kernels", 0)
seq(30, "plain again ")
print("torch reserved GB %.2f allocated GB %.2f, num_alloc_retries %d, num cudaMalloc %d" % (torch.cuda.memory_reserved() / 2**30, torch.cuda.memory_allocated() / 2**30, torch.cuda.memory_stats()["num_alloc_retries"], torch.cuda.memory_stats()["segment.all.allocated"]))
