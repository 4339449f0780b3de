"""Single huge box over several GPUs: host-side handle of the library's z-slab decomposition (csrc/mdb_dd.cu).

The step loop, the ghost-layer exchanges (ncclSend / ncclRecv enqueued on the context's stream) and the local rebuild all
live in the C library (mdb_dd_build / mdb_dd_force / mdb_dd_run / mdb_dd_global_t).  This class only does what a host
program has to do around them: one process per GPU, hand the 128-byte NCCL id of rank 0 to every rank (here over
torch.distributed, which is plumbing: any launcher that can broadcast 128 bytes will do), and pick the stream.
Reference scheme replaced: replicated positions with host-staged all-gathers of XP and DEN every step
(MD_Globle_Variables_GPU.F90:2026-2040, MD_EAM_ForceTable_GPU.F90:617-642).
"""
from . import capi


def slab_layers(ncz, world, rank):
    """z-layers [z0, z1) of cells owned by `rank` (same formula as mdb_dd_update in the library)."""
    return (rank * ncz) // world, ((rank + 1) * ncz) // world


class SlabDomain:
    def __init__(self, ctx: capi.Context, device, group=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.ctx, self.device, self.group = ctx, torch.device("cuda", device), group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        # a dedicated non-default stream (the library keeps its own stream when handed the NULL stream)
        self.stream = torch.cuda.Stream(device=self.device)
        ctx.set_stream(self.stream.cuda_stream)
        ctx.dd_set(self.rank, self.world)
        if self.world > 1:
            box = [capi.dd_nccl_id() if self.rank == 0 else None]
            dist.broadcast_object_list(box, src=0, group=group)
            ctx.dd_nccl_init(box[0])

    def rebuild(self):
        """collective: first call = replicated initial build, later calls = local rebuild of this rank's slab"""
        return self.ctx.dd_build() if self.world > 1 else self.ctx.nlist_build()

    def force(self, flags=capi.FORCE):
        return self.ctx.dd_force(flags) if self.world > 1 else self.ctx.force(flags)

    def force_virial(self):
        """pCalPTensor on the decomposed box: VTENSOR (3,3), identical on every rank"""
        return self.force(capi.FORCE | capi.VIRIAL)

    def run(self, itime0, nsteps, it0, nb_uptab, h):
        """nsteps x For_One_Step (Appshell/MD_Method_GenericMD_GPU.F90:596-627) on the decomposed box"""
        if self.world > 1:
            return self.ctx.dd_run(itime0, nsteps, it0, nb_uptab, h)
        return self.ctx.run(itime0, nsteps, it0, nb_uptab, h)

    def run_sched(self, itime0, nsteps, it0, sched, h, time_s=0.0):
        """the GMD loop with its step-size / list-period schedules (capi.Sched) -> (out-of-box count, H, TIME)"""
        if self.world > 1:
            return self.ctx.dd_run_sched(itime0, nsteps, it0, sched, h, time_s)
        return self.ctx.run_sched(itime0, nsteps, it0, sched, h, time_s)

    def step(self, itime, it0, nb_uptab, h):
        return self.run(itime, 1, it0, nb_uptab, h)

    def global_t(self):
        return self.ctx.dd_global_t() if self.world > 1 else self.ctx.global_t()

    def owned(self):
        """(a0, a1) of this rank in CELL order."""
        if self.world > 1:
            i = self.ctx.dd_info()
            return i["a0"], i["a1"]
        return 0, self.ctx.n
