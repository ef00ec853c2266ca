"""Does a kernel of libgs_b200.so running on GPU r reach GPU (1 - r)'s memory through a CUDA-IPC mapping, with plain stores and
with float atomics?  (torchrun --nproc-per-node 2 tools/peer_probe.py)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussian-splatting_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import diff_gaussian_rasterization as dgr  # noqa: E402
from gaussian_renderer.peer import PeerBuffers  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = 1 << 20
pb = PeerBuffers(n + 256, dev)
ptrs = pb.pointers()
other = (rank + 1) % world
img = torch.full((n,), 0.25 + 0.5 * rank, device=dev)
gt = torch.zeros(n, device=dev)
with torch.cuda.device(dev):
    rc = dgr._C.gsb_l1_loss_grad(img.data_ptr(), gt.data_ptr(), n, 1.0, ptrs[other], ptrs[other] + 4 * n, torch.cuda.current_stream(dev).cuda_stream)
dgr._check(rc)
torch.cuda.synchronize()
dist.barrier()
src = (rank - 1) % world                      # who wrote into MY buffer
got_grad = pb.local[:n]
got_sum = float(pb.local[n])
exp_sum = n * (0.25 + 0.5 * src)
ok = bool((got_grad == 1.0).all()) and abs(got_sum - exp_sum) <= 1e-3 * exp_sum
print(f"rank {rank}: peer store {'ok' if bool((got_grad == 1.0).all()) else 'BAD'}, peer atomic sum {got_sum:.1f} (expected {exp_sum:.1f}) -> {'OK' if ok else 'FAIL'}", flush=True)
pb.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
