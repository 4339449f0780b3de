"""Worker of tests/test_gpu_dd.py (run under torchrun, one rank per GPU): a slab-decomposed run of one box
against the same box on a single GPU."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch
import torch.distributed as dist

import util
from msmpscu_b200 import capi
from msmpscu_b200.domain import SlabDomain


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    cascade = len(sys.argv) > 2 and sys.argv[2] == "cascade"   # PKA + electronic stopping + displacement-limited time step
    c = util.bcc_case((8, 8, 20), seed=404, temp=900.0)
    # atoms near the faces of the 8 z-layers of cells are shot across them: every rebuild migrates atoms between ranks
    t = (c.xp[:, 2] - c.boxlow[2]) / c.zl[2] * 8
    frac = t - np.floor(t)
    c.xp1 = c.xp1.copy()
    c.xp1[frac > 1.0 - 0.32 * c.rr / (c.zl[2] / 8), 2] = 0.03 * c.rr / 0.5e-15
    c.xp1[frac < 0.32 * c.rr / (c.zl[2] / 8), 2] = -0.03 * c.rr / 0.5e-15
    h, it0, nup = 0.5e-15, 1, 10
    epc = ([1], [300.0], [1.0e-12], [0.1], [100.0 * util.CP_EVERG])

    def make(dev):
        ctx = capi.Context(dev)
        ctx.box_set(c.nbox, c.napb, c.boxlow, c.zl, c.ifpd, c.mass)
        ctx.upload(capi.F_XP, c.xp); ctx.upload(capi.F_XP1, c.xp1)
        ctx.upload(capi.F_ITYP, c.ityp); ctx.upload(capi.F_STATU, c.statu)
        ctx.tables_set(util.product_tables(c), c.ru * c.ru)
        ctx.set_option(capi.OPT_FORCE_PATH, capi.FORCE_PATH_TILED)
        ctx.set_option(capi.OPT_TILED_BANKORDER, 1)     # (auto would leave it off for a box this small: the big runs have it on)
        ctx.nlist_init(c.nb_rm, c.mxkvois)
        ctx.epc_set(*epc)
        return ctx

    ne = 200
    etab = np.linspace(1.0, 2.0e4, ne) * util.CP_EVERG
    stab = (1.0e-27 * np.sqrt(etab / util.CP_EVERG)).reshape(-1, 1)
    sched = capi.Sched(-1, h, h, 0.05 * c.rr, nup, nup, 100)
    ipka = int(np.argmin(np.sum((c.xp - (c.boxlow + 0.5 * c.zl)) ** 2, axis=1))) + 1

    def arm(x):
        if cascade:
            x.pka_insert(ipka, 1000.0 * util.CP_EVERG, [1.0, 3.0, 5.0])
            x.stopping_set(etab, stab, np.array([[1]]), [1], [6.3e22])

    full = make(local)             # the whole box on this GPU: the reference for this test
    full.nlist_build(); arm(full); full.force(capi.FORCE)
    if cascade:
        _, h_full, t_full = full.run_sched(0, nsteps, it0, sched, h)
    else:
        full.run(0, nsteps, it0, nup, h)

    ctx = make(local)
    dom = SlabDomain(ctx, local)
    dom.rebuild()
    arm(ctx)
    dom.force(capi.FORCE)
    # step loop, exchanges (peer memory / NCCL) and local rebuilds inside the library
    if cascade:
        _, h_dd, t_dd = dom.run_sched(0, nsteps, it0, sched, h)
    else:
        dom.run(0, nsteps, it0, nup, h)
    a0, a1 = dom.owned()
    ok = True
    if cascade:
        ok &= bool(h_dd == h_full and t_dd == t_full and t_full < nsteps * h)
    gid_f, gid_d = full.download(capi.F_GID, capi.ORDER_CELL), ctx.download(capi.F_GID, capi.ORDER_CELL)
    ok &= bool(np.array_equal(gid_f[a0:a1], gid_d[a0:a1]))
    worst = {}
    for name, f, tol in (("xp", capi.F_XP, 1e-13), ("xp1", capi.F_XP1, 1e-10), ("fp", capi.F_FP, 1e-10), ("den", capi.F_DEN, 1e-10)):
        x, y = full.download(f, capi.ORDER_CELL)[a0:a1], ctx.download(f, capi.ORDER_CELL)[a0:a1]
        worst[name] = util.relerr(y, x)
        ok &= worst[name] < tol
    # per-atom energies of the owned atoms (the tiled energy pass works on owned tiles + current ghost positions)
    full.force(capi.EPOT); ctx.force(capi.EPOT)
    worst["epot"] = util.relerr(ctx.download(capi.F_EPOT, capi.ORDER_CELL)[a0:a1], full.download(capi.F_EPOT, capi.ORDER_CELL)[a0:a1])
    ok &= worst["epot"] < 1e-12
    # virial: per-rank partial tensors of the owned tiles, summed over the ranks, against the single-GPU tensor
    vt_full = full.force(capi.FORCE | capi.VIRIAL)
    vt_dd = dom.force_virial()
    worst["virial"] = util.relerr(vt_dd, vt_full)
    ok &= worst["virial"] < 1e-10
    # temperature: Cal_GlobalT over the owned atoms of every rank, one all-reduce
    t_full = full.global_t()
    worst["global_t"] = abs(dom.global_t() - t_full) / t_full
    ok &= worst["global_t"] < 1e-10
    kf, kd = full.download(capi.F_KVOIS, capi.ORDER_CELL), ctx.download(capi.F_KVOIS, capi.ORDER_CELL)
    ok &= bool(np.array_equal(kf[a0:a1], kd[a0:a1]))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    counts = torch.tensor([a1 - a0], device="cuda")
    dist.all_reduce(counts)
    print("rank %d owns [%d,%d) of %d  worst rel err %s  ok=%s" % (rank, a0, a1, ctx.n, worst, ok), flush=True)
    if rank == 0:
        print("DD_RESULT", "PASS" if int(flag.item()) == 1 and int(counts.item()) == ctx.n else "FAIL", flush=True)
    full.close(); ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
