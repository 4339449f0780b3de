#!/usr/bin/env python
"""bench.py -- headline benchmark of the MDPSCU tabulated EAM hot path on B200.

Metric (BASELINE.json): atom-steps/s, bcc W, Marinica EAM2 table force, 1 024 000 atoms (80^3 bcc
cells), NVT via the electron-phonon thermostat (MDLocalTempCtrl/EPC, T_e = 300 K), h = 0.5 fs,
neighbour list (1.2 x 1.9 a0, MAXNB 256) rebuilt every 10 MD steps.

One bench "step" = one neighbour-list period of the GMD loop = 10 MD steps (For_One_Step x 10:
predictor -> [rebuild on the first] -> density pass -> force pass -> EPC -> corrector), i.e. one
mdb_run(ctx, itime0, 10, ...) call; value counts MD steps: atoms x 10 x K / time.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells 80]

N > 1 (torchrun, one rank per GPU): every rank runs its own independent box (MDPSCU's natural
multi-GPU grain: independent boxes, no data-path collective) -> weak scaling; time = max over ranks.
--impl reference times the reference's CPU implementation of the same path (the C restatement in
oracle/, since the PGI CUDA-Fortran reference cannot be built here) on the host cores, rank 0 only.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries ONE JSON line: NCCL_DEBUG=VERSION makes NCCL itself print "NCCL version ..." there (INFO / WARN are left alone)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

MD_PER_STEP = 10      # NB_UPTAB: MD steps per neighbour-list period
H = 0.5e-15           # s
A0 = 3.1652           # Angstrom
RU_LU, NB_FAC, MXKVOIS, NTAB = 1.9, 1.2, 256, 10000
K_LIST = 112          # stored neighbours per atom for this lattice / cutoff
# canonical algorithmic bytes per atom (SURVEY.md 8d / BASELINE.md 4): int32 full list, fp64 SoA
BYTES_PASS1 = 4 * K_LIST + 44
BYTES_PASS2 = 4 * K_LIST + 68
BYTES_STEP = 8 * K_LIST + 356 + (36 + 4 * K_LIST) / MD_PER_STEP  # = 1300.4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=80, help="bcc cells per edge (80 -> 1 024 000 atoms)")
    ap.add_argument("--nbox", type=int, default=1, help="independent boxes per GPU, concatenated as MULTIBOX (configs[2] family)")
    ap.add_argument("--dd", action="store_true", help="configs[4] family: ONE box of --cells^3 bcc cells cut into z-slabs over the "
                    "ranks (ghost-layer exchange over NCCL), strong scaling; not the default bench line")
    ap.add_argument("--potential", default="w_marinica", choices=["w_marinica", "cu_setfl"],
                    help="cu_setfl: fcc Cu with the NIST setfl potential Cu1 imported through mdb_host_setfl_ftable (configs[2] family "
                         "with a real external EAM table; --cells = fcc cells per edge)")
    ap.add_argument("--path", default="auto", choices=["auto", "generic", "tiled"])
    ap.add_argument("--cpu-cells", type=int, default=32, help="edge of the CPU sample box (32 -> 65 536 atoms)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lanes", type=int, default=0, help="tiled path: lanes per atom (2/4/8), 0 = default")
    ap.add_argument("--classes", type=int, default=1, help="tiled path: distance-classified lists on/off")
    ap.add_argument("--threads", type=int, default=0, help="tiled path: threads of the pass CTA (512/768), 0 = default")
    ap.add_argument("--stages", type=int, default=0, help="tiled path: pipeline stages of the pass kernel (2/3), 0 = default")
    return ap.parse_args()


def make_case(cells, seed, nbox=1, potential="w_marinica"):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    if potential == "cu_setfl":
        return util.fcc_cu_case((cells, cells, cells), seed=seed, nbox=nbox)
    return util.bcc_case((cells, cells, cells), a0=A0, seed=seed, ru_lu=RU_LU, nb_fac=NB_FAC, mxkvois=MXKVOIS,
                         ntab=NTAB, temp=600.0, disp=0.02, nbox=nbox)


EPC = dict(enable=[1], te=[300.0], alpha=[1.0e-12], cut=[0.1], he=[100.0 * 1.60219e-12])


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (pynvml; nvidia-smi fallback)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.reasons, self.max_mhz = index, False, [], set(), None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
            while not self.stop_flag:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
                time.sleep(0.05)
        except Exception as e:  # pragma: no cover
            self.reasons.add("sampler_error:%s" % type(e).__name__)

    def result(self):
        self.stop_flag = True
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_arm(cells, steps, warmup, threads=None):
    """The reference's CPU implementation of the path (C restatement, oracle/) on the host cores."""
    from oracle import pyorc as O
    nthr = threads or os.cpu_count() or 1
    O.lib().orc_set_threads(nthr)
    c = make_case(cells, 4242)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    md = util.oracle_md(O, c)
    md.set_epc(EPC["enable"], EPC["te"], EPC["alpha"], EPC["cut"], EPC["he"])
    md.rebuild()
    md.force()
    n = c.xp.shape[0]
    it = 0
    for _ in range(warmup):
        for _ in range(MD_PER_STEP):
            md.step(it, 1, MD_PER_STEP, H)
            it += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        for _ in range(MD_PER_STEP):
            md.step(it, 1, MD_PER_STEP, H)
            it += 1
    dt = time.perf_counter() - t0
    return n * MD_PER_STEP * steps / dt, dt, n, nthr


def base_line(args, n_atoms):
    return {
        "metric": "atom-steps/sec (W EAM, 1M atoms)", "unit": "atom-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": ("configs[1]: bcc W %d atoms (%d^3 bcc cells), Marinica EAM2 table force, NVT via "
                                "MDLocalTempCtrl EPC (Te=300K), single box per GPU" % (n_atoms, args.cells)) if args.nbox == 1 else
                               ("configs[2] family: %d independent boxes x %d atoms (%d^3 bcc cells, W Marinica EAM2 tables), "
                                "NVT via EPC, per GPU" % (args.nbox, n_atoms // args.nbox, args.cells)),
                   "atoms_per_gpu": n_atoms, "md_steps_per_step": MD_PER_STEP, "h_fs": 0.5, "cutoff_a0": RU_LU,
                   "list_cutoff_a0": RU_LU * NB_FAC, "rebuild_every": MD_PER_STEP, "ntab": NTAB,
                   "l2_policy": "inputs larger than L2 (neighbour list + state ~0.6 GB per step, L2 126 MB)",
                   "parallelism": "independent box per GPU, no data-path collective"},
    }


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    val, dt, n, nthr = cpu_arm(args.cpu_cells, args.steps, args.warmup)
    line = base_line(args, args.cells ** 3 * 2)
    sample = "bcc W %d atoms (%d^3 cells), same potential/cutoffs/EPC, %d MD steps per step incl. 1 rebuild" % (
        n, args.cpu_cells, MD_PER_STEP)
    line.update({"impl": "reference", "value": val, "ms_per_step": dt / args.steps * 1e3,
                 "cpu_baseline": {"value": val, "unit": "atom-steps/s", "cores": nthr, "kind": "port", "sample": sample},
                 "e2e": {"value": val, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0, "n_gpus": args.gpus})
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist
    from msmpscu_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    c = make_case(args.cells, 12346 + rank, args.nbox, args.potential)
    n = c.xp.shape[0]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    ctx = capi.Context(local)
    # an explicit torch stream (a non-null handle): the library launches on it and the timing events are recorded on it.
    # (the null handle of torch's default stream would make the context fall back to its own stream, mdb_ctx_set_stream)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.box_set(c.nbox, c.napb, c.boxlow, c.zl, c.ifpd, c.mass)
    ctx.set_option(capi.OPT_FORCE_PATH, {"auto": 0, "generic": 1, "tiled": 2}[args.path])
    ctx.tables_set(util.product_tables(c), c.ru * c.ru)
    ctx.nlist_init(c.nb_rm, c.mxkvois)
    if args.lanes:
        ctx.set_option(capi.OPT_TILED_LANES, args.lanes)
    if args.threads:
        ctx.set_option(capi.OPT_TILED_THREADS, args.threads)
    if args.stages:
        ctx.set_option(capi.OPT_TILED_STAGES, args.stages)
    ctx.set_option(capi.OPT_TILED_CLASSES, args.classes)
    ctx.epc_set(EPC["enable"], EPC["te"], EPC["alpha"], EPC["cut"], EPC["he"])

    # host buffers in pinned memory, reference layout XP(N,3) column-major
    hx = torch.from_numpy(capi.colmajor(c.xp)).pin_memory()
    hv = torch.from_numpy(capi.colmajor(c.xp1)).pin_memory()
    hf = torch.empty(3 * n, dtype=torch.float64).pin_memory()
    ctx.upload_raw(capi.F_XP, hx.data_ptr())
    ctx.upload_raw(capi.F_XP1, hv.data_ptr())
    ctx.upload(capi.F_ITYP, c.ityp)
    ctx.upload(capi.F_STATU, c.statu)
    ctx.nlist_build()
    ctx.force(capi.FORCE)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(k):
            fn(i)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    itime = [0]

    def step_resident(_):
        ctx.run(itime[0], MD_PER_STEP, 1, MD_PER_STEP, H)
        itime[0] += MD_PER_STEP

    def step_e2e(_):
        # the reference-facing call with HOST buffers: CopyIn (XP, XP1) -> 10 x For_One_Step -> CopyOut (XP, XP1, FP)
        ctx.upload_raw(capi.F_XP, hx.data_ptr())
        ctx.upload_raw(capi.F_XP1, hv.data_ptr())
        ctx.run(itime[0], MD_PER_STEP, 1, MD_PER_STEP, H)
        itime[0] += MD_PER_STEP
        ctx.download_raw(capi.F_XP, hx.data_ptr())
        ctx.download_raw(capi.F_XP1, hv.data_ptr())
        ctx.download_raw(capi.F_FP, hf.data_ptr())

    for i in range(max(args.warmup, 3)):
        step_resident(i)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = ctx.launch_count()
    ms = timed(step_resident, args.steps)
    launches = ctx.launch_count() - l0
    clocks = sampler.result()
    value = world * n * MD_PER_STEP * args.steps / (ms * 1e-3)

    # end to end through host buffers
    step_e2e(0)
    ms_e2e = timed(step_e2e, max(3, args.steps // 4))
    e2e_val = world * n * MD_PER_STEP * max(3, args.steps // 4) / (ms_e2e * 1e-3)

    # per-kernel device times (CUDA events on the launching stream) for the roofline
    ctx.prof_reset()
    ctx.prof_enable(True)
    for i in range(3):
        step_resident(i)
    prof = ctx.prof_get()
    ctx.prof_enable(False)
    tot = sum(v[1] for v in prof.values())
    dom = max(("pass1", "pass2"), key=lambda k: prof[k][1])
    nl, tms = prof[dom]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    per_launch_bytes = (BYTES_PASS1 if dom == "pass1" else BYTES_PASS2) * n
    achieved = per_launch_bytes / (tms / nl * 1e-3) / 1e9 if nl else 0.0
    traffic, pipes = None, None
    try:  # DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (same workload only)
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if int(tj.get("atoms", -1)) == n and args.path != "generic":
            traffic = tj.get(dom)
            pipes = tj.get("pipes", {}).get(dom)
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": traffic, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s",
            "algorithmic_bytes_per_atom": BYTES_PASS1 if dom == "pass1" else BYTES_PASS2,
            "avg_launch_ms": tms / nl if nl else None,
            "kernel_share_of_step": tms / tot if tot else None,
            "whole_step": {"achieved": value / world * BYTES_STEP / 1e9, "frac": value / world * BYTES_STEP / 1e9 / peak,
                           "bytes_per_atom_step": BYTES_STEP},
            "per_class_ms_per_md_step": {k: v[1] / (3 * MD_PER_STEP) for k, v in prof.items() if v[0]}}
    if traffic and nl:
        # the kernel's OWN stream (16-bit class-limited slot list, positions from L2) and the pipe that actually bounds it
        roof["own_stream"] = {"dram_bytes_per_atom": traffic / n, "achieved_GBs": traffic / (tms / nl * 1e-3) / 1e9,
                              "frac_of_hbm_peak": traffic / (tms / nl * 1e-3) / 1e9 / peak}
    if pipes:
        roof["limiting_pipe"] = dict(pipes, source="ncu --set full capture committed under profiles/ (not measured in this run)")

    line = base_line(args, n)
    if args.potential == "cu_setfl":
        line["config"]["workload"] = ("configs[2] family with an imported NIST setfl table: %d independent boxes x %d atoms of fcc Cu "
                                      "(%d^3 cells, Cu1.eam.fs.setfl, force cutoff 6 A, list cutoff 7.2 A), NVT via EPC, per GPU"
                                      % (args.nbox, n // args.nbox, args.cells))
        line["config"].update({"cutoff_a0": 6.0 / 3.639087, "list_cutoff_a0": 7.2 / 3.639087})
        line["roofline_note"] = "canonical byte counts assume K_list = 112 (bcc W); this workload lists 134 neighbours"
    line.update({"value": value, "ms_per_step": ms / args.steps, "clocks": clocks, "gpu_launches": int(launches),
                 "e2e": {"value": e2e_val, "unit": "atom-steps/s", "h2d_bytes_per_step": 2 * 24 * n,
                         "d2h_bytes_per_step": 3 * 24 * n, "ms_per_step": ms_e2e / max(3, args.steps // 4)},
                 "roofline": roof, "force_path": args.path})
    line["config"]["atoms_total"] = n * world

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        val, dt, ncpu, nthr = cpu_arm(args.cpu_cells, 2, 1)
        line["cpu_baseline"] = {"value": val, "unit": "atom-steps/s", "cores": nthr, "kind": "port",
                                "sample": "bcc W %d atoms (%d^3 cells), same potential/cutoffs/EPC, 2 x %d MD steps "
                                          "incl. rebuilds, OpenMP over atoms" % (ncpu, args.cpu_cells, MD_PER_STEP)}
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_dd(args):
    """One box over all ranks (slab decomposition, msmpscu_b200/domain.py): the box is the same on every rank (same
    seed); value = atoms of the box x MD steps / max-over-ranks device time."""
    import torch
    import torch.distributed as dist
    from msmpscu_b200 import capi
    from msmpscu_b200.domain import SlabDomain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    c = make_case(args.cells, 777)
    n = c.xp.shape[0]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    ctx = capi.Context(local)
    ctx.box_set(c.nbox, c.napb, c.boxlow, c.zl, c.ifpd, c.mass)
    ctx.set_option(capi.OPT_FORCE_PATH, capi.FORCE_PATH_TILED)
    ctx.tables_set(util.product_tables(c), c.ru * c.ru)
    ctx.nlist_init(c.nb_rm, c.mxkvois)
    ctx.epc_set(EPC["enable"], EPC["te"], EPC["alpha"], EPC["cut"], EPC["he"])
    ctx.upload(capi.F_XP, c.xp); ctx.upload(capi.F_XP1, c.xp1)
    ctx.upload(capi.F_ITYP, c.ityp); ctx.upload(capi.F_STATU, c.statu)
    dom = SlabDomain(ctx, local)
    dom.rebuild()
    ctx.force(capi.DEN); dom.exchange(); ctx.force(capi.FORCE | capi.NOPASS1)
    stream = dom.stream
    itime = [0]

    def block(_):
        for _i in range(MD_PER_STEP):
            dom.step(itime[0], 1, MD_PER_STEP, H)
            itime[0] += 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(max(args.warmup, 3)):
        block(i)
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count()
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        block(i)
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = ctx.launch_count() - l0
    clocks = sampler.result()
    dom.phase_ms = {}
    block(0)
    phases = dom.phase_report()   # one extra block of 10 steps with per-phase CUDA events (not part of the timed region)
    dom.phase_ms = None
    if world > 1:
        allp = [None] * world
        dist.all_gather_object(allp, phases)
        phases = {k: [round(p.get(k, 0.0), 3) for p in allp] for k in phases}
    a0, a1 = dom.owned()
    line = base_line(args, n)
    line["scaling"] = "strong"
    line["config"].update({"workload": "configs[4] family: ONE bcc W box of %d atoms (%d^3 cells) cut into %d z-slabs, ghost-layer "
                                       "exchange of {x,y,z,den} records over NCCL twice per step, list rebuilt every %d steps "
                                       "(owned ranges broadcast, identical device sort on every rank)" % (n, args.cells, world, MD_PER_STEP),
                           "atoms_per_gpu": a1 - a0, "atoms_total": n,
                           "parallelism": "z-slab domain decomposition, 1 ghost cell layer per side"})
    line.update({"value": n * MD_PER_STEP * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps, "clocks": clocks,
                 "gpu_launches": int(launches), "force_path": "tiled", "mode": "dd",
                 "phase_ms_per_block_by_rank": phases})
    if rank == 0:
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.dd:
        run_dd(a)
    else:
        run_ours(a)
