"""Multi-box / parallel-replica dispatcher: independent boxes are sharded per GPU (one rank per GPU),
no data-path collective; only per-box scalars (T, E, virial, event flags) are gathered at output /
check intervals.

Replaces the reference's "one host thread, <= 6 GPUs, cudaSetDevice loop" (MSMLIB/sor/CommonGPU/
MSM_MultiGPU_Basic.F90:22,571-597) and its contiguous partition of the concatenated box array over
devices (MDLIB/sor/CommonGPU/MD_NeighborsList_GPU.F90:1576-1605).  The reference's natural grain is kept:
boxes are concatenated per device (MULTIBOX, MD_Globle_Variables_GPU.F90:601-606), cell ids are offset per
box and neighbour cells never cross boxes (MD_NeighborsList_GPU.F90:798-811,991-1065)."""
import numpy as np


def partition_boxes(nbox, world, rank):
    """Contiguous blocks: box b -> rank floor(b*world/nbox).  Returns (first_box, count)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    first = -(-rank * nbox // world)          # ceil(rank*nbox/world)
    nxt = -(-(rank + 1) * nbox // world)
    return first, nxt - first


class MultiBoxDispatcher:
    """Host-side plumbing over torch.distributed (backend nccl on GPUs, gloo in CPU tests)."""

    def __init__(self, nbox_total, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.nbox_total = int(nbox_total)
        self.first, self.count = partition_boxes(self.nbox_total, self.world, self.rank)
        self.counts = [partition_boxes(self.nbox_total, self.world, r)[1] for r in range(self.world)]

    def _device(self):
        """the rank's own GPU (LOCAL_RANK), not "whatever device is current": the library's calls select their context's device"""
        import os
        import torch
        if self.dist.get_backend(self.group) != "nccl":
            return "cpu"
        return torch.device("cuda", int(os.environ.get("LOCAL_RANK", torch.cuda.current_device())))

    def local_boxes(self):
        return range(self.first, self.first + self.count)

    def gather_box_scalars(self, local, device=None):
        """local: (count, k) per-box values of this rank -> (nbox_total, k) on every rank (box order)."""
        import torch
        local = np.asarray(local, dtype=np.float64).reshape(self.count, -1)
        k = local.shape[1]
        if self.world == 1:
            return local.copy()
        dev = device or self._device()
        mx = max(self.counts)
        buf = torch.zeros((mx, k), dtype=torch.float64, device=dev)
        if self.count:
            buf[: self.count] = torch.from_numpy(local).to(dev)
        out = [torch.empty_like(buf) for _ in range(self.world)]
        self.dist.all_gather(out, buf, group=self.group)
        return np.concatenate([o[:c].cpu().numpy() for o, c in zip(out, self.counts)], axis=0)

    def reduce_sum(self, x, device=None):
        import torch
        x = np.asarray(x, dtype=np.float64)
        if self.world == 1:
            return x.copy()
        dev = device or self._device()
        t = torch.from_numpy(x.copy()).to(dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy()

    def reduce_max(self, x, device=None):
        import torch
        x = np.asarray(x, dtype=np.float64)
        if self.world == 1:
            return x.copy()
        dev = device or self._device()
        t = torch.from_numpy(x.copy()).to(dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return t.cpu().numpy()
