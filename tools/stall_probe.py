"""Where do the 20-150 ms host stalls come from?  (SCALE_r01.json: one straggler step per 20-step region at 8 GPUs; seen at
1 GPU too.)  Runs the sync-free bench step N times under different conditions and records, per step, the HOST time needed to
enqueue it and the DEVICE time between step marks:

    quiet        no helper thread at all
    nvml_full    the bench's clock sampler: nvmlDeviceGetClockInfo + nvmlDeviceGetCurrentClocksEventReasons every 100 ms
    nvml_clock   only nvmlDeviceGetClockInfo every 100 ms
    nvml_slow    the full sampler every 1000 ms
    sleeper      a Python thread that only sleeps / wakes every 100 ms (GIL hand-over without NVML)

    python tools/stall_probe.py [steps]          # one JSON line
"""
import gc
import json
import math
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
import torch  # noqa: E402
import bench  # noqa: E402
from gaussian_renderer.synthetic import make_scene  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402
from gaussian_renderer import AsyncCapacity, GradientBucket, render_views_backward  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
dev = torch.device("cuda", 0)
V, H, W = 8, 1080, 1920
scene = make_scene(1_000_000, seed=0, log_scale_mean=bench.LOG_SCALE_MEAN)
pc = bench.BenchGaussians(scene, 3, dev)
bucket = GradientBucket(pc.parameters())
bg = torch.zeros(3, device=dev)
gts = [torch.rand(3, H, W, device=dev) for _ in range(V)]
cams = [bench.BenchCamera(W, H, math.radians(60.0), *bench.view_pose(i, 3.0), dev) for i in range(V)]
cap = AsyncCapacity(dev)


def step():
    render_views_backward(cams, pc, bench.Pipe(), bg, lambda img, d, i: dgr.l1_loss_and_grad(img, gts[i]), loss_returns_grad=True,
                          overwrite=True, capacity=cap)


class Helper:
    def __init__(self, kind):
        self.kind, self.stop, self.calls, self.worst_ms = kind, False, 0, 0.0
        self.thread = None

    def start(self):
        if self.kind == "quiet":
            return
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        period = 1.0 if self.kind == "nvml_slow" else 0.1

        def loop():
            while not self.stop:
                t0 = time.perf_counter()
                if self.kind != "sleeper":
                    pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                    if self.kind in ("nvml_full", "nvml_slow"):
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                self.worst_ms = max(self.worst_ms, (time.perf_counter() - t0) * 1e3)
                self.calls += 1
                time.sleep(period)
        self.thread = threading.Thread(target=loop, daemon=True)
        self.thread.start()

    def end(self):
        self.stop = True
        if self.thread:
            self.thread.join()


def run(kind):
    for _ in range(5):
        step()
    torch.cuda.synchronize()
    helper = Helper(kind)
    helper.start()
    gc.collect(); gc.disable()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    host = []
    torch.cuda.synchronize()
    marks[0].record()
    for k in range(steps):
        if k % 20 == 0:
            torch.cuda.synchronize()      # same rhythm as the bench: the host starts every 20-step region level with the device
        t0 = time.perf_counter()
        step()
        marks[k + 1].record()
        host.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    gc.enable()
    helper.end()
    devms = [marks[k].elapsed_time(marks[k + 1]) for k in range(steps)]
    med_h, med_d = sorted(host)[steps // 2], sorted(devms)[steps // 2]
    return {"host_ms_median": round(med_h, 3), "host_ms_max": round(max(host), 2), "host_stalls_over_5ms": [round(x, 1) for x in host if x > med_h + 5.0],
            "device_ms_median": round(med_d, 3), "device_steps_over_1.3x": [(k, round(x, 1)) for k, x in enumerate(devms) if x > 1.3 * med_d],
            "helper_calls": helper.calls, "helper_worst_call_ms": round(helper.worst_ms, 2)}


out = {"steps_per_condition": steps, "capacity_ok": None}
for kind in ("quiet", "nvml_full", "quiet", "nvml_clock", "sleeper", "nvml_slow", "quiet"):
    key = kind if kind not in out else kind + "_again" + str(sum(1 for k in out if k.startswith(kind)))
    out[key] = run(kind)
out["capacity_ok"] = cap.check()
print(json.dumps(out))
