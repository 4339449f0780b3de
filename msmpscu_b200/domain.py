"""Single huge box over several GPUs: z-slab domain decomposition with ghost-layer exchange (NCCL).

The reference's multi-GPU scheme for one box is "contiguous cell range per device, REPLICATED positions,
host-staged copies" (XP all-gather through the host after the predictor, MD_Globle_Variables_GPU.F90:2026-2040;
DEN all-gather between the passes, MD_EAM_ForceTable_GPU.F90:617-642).  Here every rank keeps full-size
arrays in the common cell-sorted order -- so neighbour slots, tiles and ids mean the same everywhere -- but
only its owned z-layers of cells plus ONE ghost layer on each side are kept current:

  per step      predictor (owned) -> send bottom/top layer records to the ranks below/above, receive the two
                ghost layers (NCCL send/recv on contiguous ranges of the packed {x,y,z,den} array)
                -> density pass (owned tiles) -> the same exchange again (now carrying DEN)
                -> force pass -> EPC + corrector (owned)
  per rebuild   every rank broadcasts its owned ranges of {pos, XP1, DIS, STATU}; all ranks run the same
                deterministic device cell sort on identical data, then build lists for their own layers.
  the distance-class shortcut of the tiled passes needs the GLOBAL max displacement: one 4-byte all-reduce.

A z-layer of cells is one contiguous atom range in cell order (cells are x-fastest, z-slowest), which is
what makes the exchange two plain range copies per neighbour.
"""
import numpy as np

from . import capi
from .constants import CP_KB


class _DevArray:
    """Zero-copy view of device memory for torch (torch.as_tensor understands __cuda_array_interface__)."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def slab_layers(ncz, world, rank):
    """z-layers [z0, z1) of cells owned by `rank` (same formula as mdb_dd_update)."""
    return (rank * ncz) // world, ((rank + 1) * ncz) // world


class SlabDomain:
    def __init__(self, ctx: capi.Context, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.ctx, self.device, self.group = ctx, torch.device("cuda", device), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        ctx.dd_set(self.rank, self.world)
        self.info = None
        # The library launches on torch's current stream from here on: NCCL operations issued through torch.distributed
        # are ordered against that stream on the device (the collective waits for the kernels before it, work.wait()
        # makes the stream wait for the collective), so a step needs no host synchronisation between its kernels and
        # its exchanges.
        # (a dedicated non-default torch stream: the library keeps its own stream when handed the NULL stream)
        self.stream = torch.cuda.Stream(device=self.device)
        ctx.set_stream(self.stream.cuda_stream)
        self.shared_stream = True
        self.phase_ms = None   # set to {} to collect per-phase device times (CUDA events on the step's stream)
        self._ev = []

    # ---- tensor views (pointers change at every rebuild: the re-sort double-buffers)
    def _views(self):
        t, n = self.torch, self.ctx.n
        mk = lambda f, shape, ts: t.as_tensor(_DevArray(self.ctx.devptr(f), shape, ts), device=self.device)
        self.pos = mk(capi.F_POS4, (n, 4), "<f8")
        self.xp1 = mk(capi.F_XP1, (3, n), "<f8")
        self.dis = mk(capi.F_DIS, (3, n), "<f8")
        self.statu = mk(capi.F_STATU, (n,), "<i4")
        self.d2max = mk(capi.F_D2MAX, (1,), "<i4")

    def _sync_stream(self):
        if not self.shared_stream:
            self.ctx.sync()  # a library-private stream would have to be drained before torch.distributed touches the arrays

    def rebuild(self):
        """All ranks obtain every rank's owned ranges, sort, and build the lists of their own layers."""
        with self.torch.cuda.stream(self.stream):
            return self._rebuild()

    def _rebuild(self):
        if self.world > 1 and self.info is not None:
            self._views()
            self._sync_stream()
            ranges = [None] * self.world
            mine = self.torch.tensor([self.info["a0"], self.info["a1"]], device=self.device, dtype=self.torch.int64)
            allr = [self.torch.empty_like(mine) for _ in range(self.world)]
            self.dist.all_gather(allr, mine, group=self.group)
            for r in range(self.world):
                a0, a1 = (int(v) for v in allr[r].tolist())
                ranges[r] = (a0, a1)
            # positions of every atom: the sort must be the same on all ranks
            for r, (a0, a1) in enumerate(ranges):
                if a1 > a0:
                    self.dist.broadcast(self.pos[a0:a1], src=r, group=self.group)
            # velocities, displacements and status only travel with atoms that change owner: between two rebuilds an atom moves
            # far less than a cell, so a new owned atom comes from the rank's own range or from a neighbour's boundary layer
            # -- the ghost ranges.  (Values of atoms deeper inside other slabs stay stale here and are never read.)
            i = self.info
            ops = []
            for arr in (self.xp1[0], self.xp1[1], self.xp1[2], self.dis[0], self.dis[1], self.dis[2], self.statu):
                ops += [self.dist.P2POp(self.dist.isend, arr[i["sb0"]:i["sb1"]], i["below"], self.group),
                        self.dist.P2POp(self.dist.isend, arr[i["st0"]:i["st1"]], i["above"], self.group),
                        self.dist.P2POp(self.dist.irecv, arr[i["ga0"]:i["ga1"]], i["above"], self.group),
                        self.dist.P2POp(self.dist.irecv, arr[i["gb0"]:i["gb1"]], i["below"], self.group)]
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
        nout = self.ctx.nlist_build()
        if self.world > 1:
            self.info = self.ctx.dd_info()
            self._views()
        return nout

    def exchange(self):
        """Boundary-layer records to the neighbours, ghost layers from them (positions, and DEN after pass 1)."""
        if self.world == 1:
            return
        with self.torch.cuda.stream(self.stream):
            self._exchange()

    def _exchange(self):
        i, dist = self.info, self.dist
        self._sync_stream()
        ops = [dist.P2POp(dist.isend, self.pos[i["sb0"]:i["sb1"]], i["below"], self.group),
               dist.P2POp(dist.isend, self.pos[i["st0"]:i["st1"]], i["above"], self.group),
               # order matters when both neighbours are the same rank (world == 2): its first send is its bottom layer
               dist.P2POp(dist.irecv, self.pos[i["ga0"]:i["ga1"]], i["above"], self.group),
               dist.P2POp(dist.irecv, self.pos[i["gb0"]:i["gb1"]], i["below"], self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()   # stream-level wait: the next kernel on this stream starts after the ghost layers have landed

    def step(self, itime, it0, nb_uptab, h):
        """One GMD step (Appshell/MD_Method_GenericMD_GPU.F90:596-627) on the decomposed box."""
        c = self.ctx
        mark = self._mark
        with self.torch.cuda.stream(self.stream):
            mark(None)
            c.predict(h)
            mark("predict")
            if (itime - it0) % nb_uptab == 0:
                self._rebuild()
                mark("rebuild")
            else:
                if self.world > 1:
                    self._sync_stream()
                    self.dist.all_reduce(self.d2max, op=self.dist.ReduceOp.MAX, group=self.group)
                    mark("allreduce_d2max")
                    self._exchange()
                mark("exchange_pos")
            c.force(capi.DEN)
            mark("pass1")
            if self.world > 1:
                self._exchange()
            mark("exchange_den")
            c.force(capi.FORCE | capi.NOPASS1)
            mark("pass2")
            c.epc_correct(h)
            mark("correct")

    def force_virial(self):
        """pCalPTensor on the decomposed box (CALPTENSOR_EAM_Force_Table2A_DEV, MD_EAM_ForceTable_GPU.F90:1366): density pass,
        DEN ghost exchange, force pass with the virial epilogue over the owned tiles, then ONE 72-byte all-reduce of the
        per-rank partial tensors (the reference sums its per-device partial tensors on the host, :1434-1466).
        Returns VTENSOR (3,3), identical on every rank."""
        c = self.ctx
        with self.torch.cuda.stream(self.stream):
            c.force(capi.DEN)
            if self.world > 1:
                self._exchange()
            vt = c.force(capi.FORCE | capi.VIRIAL | capi.NOPASS1)
            if self.world > 1:
                t = self.torch.as_tensor(np.ascontiguousarray(vt), device=self.device)
                self.dist.all_reduce(t, group=self.group)
                vt = t.cpu().numpy()
        return vt

    def global_t(self):
        """Cal_GlobalT_DEV (CommonGPU/MD_DiffScheme_GPU.F90:1042-1064) on the decomposed box: the EKIN kernel over the arrays of
        this rank, sum and count of EKIN >= 0 over its OWNED atoms (velocities of the other atoms are stale here), one
        all-reduce of the two numbers, CURT = 2*sum/count/(3*CP_KB) (:1062).  Identical on every rank."""
        c, t = self.ctx, self.torch
        with t.cuda.stream(self.stream):
            c.ekin()
            a0, a1 = self.owned()
            ek = t.as_tensor(_DevArray(c.devptr(capi.F_EKIN), (c.n,), "<f8"), device=self.device)[a0:a1]
            ok = ek >= 0.0
            acc = t.stack((t.where(ok, ek, t.zeros_like(ek)).sum(), ok.sum().to(t.float64)))
            if self.world > 1:
                self.dist.all_reduce(acc, group=self.group)
            s, n = (float(v) for v in acc.tolist())
        return 2.0 * s / n / (3.0 * CP_KB)

    def _mark(self, name):
        if self.phase_ms is None:
            return
        e = self.torch.cuda.Event(enable_timing=True)
        e.record(self.stream)
        self._ev.append((name, e))

    def phase_report(self):
        """Sum of device time per phase since the last call (needs phase_ms = {} before the steps)."""
        self.torch.cuda.synchronize(self.device)
        out, prev = {}, None
        for name, e in self._ev:
            if name is not None and prev is not None:
                out[name] = out.get(name, 0.0) + prev.elapsed_time(e)
            prev = e
        self._ev = []
        return out

    def owned(self):
        """(a0, a1) of this rank in CELL order."""
        return (self.info["a0"], self.info["a1"]) if self.world > 1 else (0, self.ctx.n)
