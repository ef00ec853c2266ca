"""Peer-mapped gradient buffers for the ranks of ONE node (plumbing for the fused reduce-scatter of csrc/preprocess.cu):
every rank owns one float32 buffer in its GPU's memory (gsb_peer_alloc: cudaMalloc + CUDA-IPC handle) and maps the buffers of
all other ranks into its own address space ON ITS OWN DEVICE (gsb_peer_open), so that a kernel running on GPU i adds its results
straight into GPU j's memory over NVLink / NVSwitch.  One process per GPU; the process group only carries the 64-byte handles.
The local buffer is exposed to torch as an ordinary tensor (``.local``) through the CUDA array interface."""
from __future__ import annotations

from typing import List

import torch

import diff_gaussian_rasterization as _dgr


class _Cai:          # the CUDA array interface torch.as_tensor understands
    def __init__(self, ptr: int, numel: int):
        self.__cuda_array_interface__ = {"shape": (numel,), "typestr": "<f4", "data": (ptr, False), "version": 2}


class PeerBuffers:
    def __init__(self, numel: int, device: torch.device, group=None):
        import torch.distributed as dist
        self.group, self.device, self.numel = group, device, int(numel)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self._ptr, handle = _dgr.peer_alloc(device, 4 * self.numel)
        self.local = torch.as_tensor(_Cai(self._ptr, self.numel), device=device)
        gathered: List[bytes] = [b""] * self.world
        dist.all_gather_object(gathered, handle, group=group)
        self._opened: List[int] = []
        self.ptrs: List[int] = []
        for r, h in enumerate(gathered):
            if r == self.rank:
                self.ptrs.append(self._ptr)
            else:
                p = _dgr.peer_open(device, h)
                self._opened.append(p)
                self.ptrs.append(p)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def pointers(self) -> List[int]:
        """Base address of every rank's buffer as seen from THIS process (index = rank)."""
        return list(self.ptrs)

    def close(self) -> None:
        import torch.distributed as dist
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)                   # nobody frees while a peer may still write
        for p in self._opened:
            _dgr.peer_close(self.device, p)
        self._opened = []
        dist.barrier(group=self.group)
        if self._ptr:
            self.local = None
            _dgr.peer_free(self.device, self._ptr)
            self._ptr = 0


class PeerGradientBucket:
    """The data-parallel gradient bucket whose reduction is FUSED into the backward pass (gsb_backward_batch_peer): gaussians are
    partitioned over the ranks; the kernel that writes the gradients adds every row into the buffer of its owner while it
    computes (TMA bulk reduce-adds over NVLink), so what remains of the collective is a barrier and an in-place all-gather of the
    owned rows -- half the bytes of an all-reduce, and none of NCCL's reduction kernels competing with the compute kernel.

    ``named`` maps the rasterizer's input names ("means3D", "shs", "opacities", "scales", "rotations") to the [P, ...] leaf
    tensors; their ``.grad`` is pointed at this bucket (``begin_step``).  Two buffers alternate between steps so that one barrier
    per step is enough (an owner zeroes buffer B before the barrier of the step that used A; nobody adds into B before it)."""

    ORDER = ("means3D", "opacities", "scales", "rotations", "shs")

    def __init__(self, named: dict, group=None, block: int = 128):
        import torch.distributed as dist
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.params = {k: named[k] for k in self.ORDER}
        any_p = self.params["means3D"]
        self.P, self.device = int(any_p.shape[0]), any_p.device
        self.rows_per_rank = ((self.P + self.world - 1) // self.world + block - 1) // block * block
        self.P_pad = self.rows_per_rank * self.world
        self.width = {k: int(v.numel() // max(1, v.shape[0])) for k, v in self.params.items()}
        self.offset, off = {}, 0
        for k in self.ORDER:
            self.offset[k] = off
            off += (self.P_pad * self.width[k] + 63) // 64 * 64          # 256-byte aligned tensors
        self.numel = off
        self.buffers = [PeerBuffers(self.numel, self.device, group), PeerBuffers(self.numel, self.device, group)]
        self._flag = torch.zeros(1, device=self.device)
        self.k = 0
        self.begin_step()

    def _view(self, buf: torch.Tensor, k: str, padded: bool = False) -> torch.Tensor:
        t = buf[self.offset[k]:self.offset[k] + self.P_pad * self.width[k]].view(self.P_pad, self.width[k])
        return t if padded else t[:self.P]

    def begin_step(self) -> None:
        """Points the parameters' ``.grad`` at the buffer of this step."""
        cur = self.buffers[self.k].local
        for k, p in self.params.items():
            p.grad = self._view(cur, k).view_as(p)

    def table(self):
        """The ``peers`` argument of render_views_backward for this step."""
        return (self.world, self.rank, self.rows_per_rank, self.buffers[self.k].pointers())

    def grads(self) -> dict:
        cur = self.buffers[self.k].local
        return {k: self._view(cur, k).view_as(p) for k, p in self.params.items()}

    def finish(self) -> None:
        """After the backward pass of every rank has been enqueued: barrier (all adds have landed), in-place all-gather of the
        owned rows (every rank ends with the full summed gradient), and the other buffer's owned rows zeroed for the next
        step.  Stream-ordered; no host synchronisation."""
        import torch.distributed as dist
        nxt = self.buffers[1 - self.k].local
        lo, hi = self.rank * self.rows_per_rank, (self.rank + 1) * self.rows_per_rank
        for k in self.ORDER:
            self._view(nxt, k, padded=True)[lo:hi].zero_()
        dist.all_reduce(self._flag, group=self.group)                  # barrier in stream order
        cur = self.buffers[self.k].local
        outs = [self._view(cur, k, padded=True) for k in self.ORDER]
        manager = getattr(dist, "_coalescing_manager", None)
        if manager is not None:
            with manager(group=self.group, async_ops=False):
                for o in outs:
                    dist.all_gather_into_tensor(o, o[lo:hi], group=self.group)
        else:
            for o in outs:
                dist.all_gather_into_tensor(o, o[lo:hi], group=self.group)
        self.k = 1 - self.k

    def close(self) -> None:
        for p in self.params.values():
            p.grad = None
        for b in self.buffers:
            b.close()
