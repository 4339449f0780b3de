#!/usr/bin/env python
"""bench.py -- headline benchmark of the MDPSCU tabulated EAM hot path on B200.

Metric (BASELINE.json): atom-steps/s, bcc W, Marinica EAM2 table force, 1 024 000 atoms (80^3 bcc cells), NVT via the
electron-phonon thermostat (MDLocalTempCtrl/EPC, T_e = 300 K), h = 0.5 fs, neighbour list (1.2 x 1.9 a0, MAXNB 256)
rebuilt every 10 MD steps.

One bench "step" = one OUTPUT INTERVAL of the GMD loop = 100 MD steps = 10 neighbour-list periods (For_One_Step x 100:
predictor -> [rebuild on the first step of a period] -> density pass -> force pass -> EPC -> corrector), i.e. one
mdb_run(ctx, itime0, 100, ...) call.  value counts MD steps: atoms x 100 x K / time.  (Round 1 used one list period per
step; 20 of those are a 0.13 s timed region, too short to be trusted, so the step now is the reference's output interval.)

  value      the box stays resident in HBM; CUDA events on the launching stream around the K steps
  e2e        the same K steps through the C ABI with HOST buffers: every step uploads XP, XP1 from page-locked memory
             (CopyIn_SimBox_DEV), runs its 100 MD steps and downloads XP, XP1, FP (CopyOut_SimBox_DEV).  Consecutive
             steps are independent jobs (the multi-box statistics pattern: box after box through one GPU) and alternate
             between two device contexts on two streams, so the copies of one job overlap the steps of the other; one
             host synchronisation per job.  "e2e_diagnostics" adds the same pipeline with a transfer every 10 MD steps
             (round 1's e2e unit) and the serial single-context form.
  roofline   dominant kernel: canonical algorithmic bytes per launch / its mean launch time (CUDA events, this run)
  parity_check  after the timed region, on the state just timed: tiled forces vs the generic path (all atoms) and vs
             the CPU oracle (10 000-atom range), per-atom relative error
  cpu_baseline  the reference's CPU routines (C restatement, oracle/) on all host cores, full 1 024 000-atom box

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--cells 80]

N > 1 (torchrun, one rank per GPU): every rank runs its own independent box (MDPSCU's natural multi-GPU grain: independent
boxes, no data-path collective) -> weak scaling; time = max over ranks.  --impl reference times the reference's CPU path
(Cal_NeighboreList2C + two-pass table force + Swope steps; the C restatement in oracle/, built -O3 -march=native, since
the PGI CUDA-Fortran reference cannot be built here) with all host threads, rank 0 only.
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
sys.path.insert(0, os.path.join(ROOT, "tests"))
# stdout carries ONE JSON line: NCCL_DEBUG=VERSION makes NCCL itself print "NCCL version ..." there (INFO / WARN are left alone)
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

MD_PER_PERIOD = 10    # NB_UPTAB: MD steps per neighbour-list period
PERIODS_PER_STEP = 10
MD_PER_STEP = MD_PER_PERIOD * PERIODS_PER_STEP   # one output interval
H = 0.5e-15           # s
CP_EVERG = 1.60219e-12  # MSMLIB/sor/Common/MSM_Const.F90:74
A0 = 3.1652           # Angstrom
RU_LU, NB_FAC, MXKVOIS, NTAB = 1.9, 1.2, 256, 10000
K_LIST = 112          # stored neighbours per atom for this lattice / cutoff
# canonical algorithmic bytes per atom (SURVEY.md 8d / BASELINE.md 4): int32 full list, fp64 SoA
BYTES_PASS1 = 4 * K_LIST + 44
BYTES_PASS2 = 4 * K_LIST + 68
BYTES_STEP = 8 * K_LIST + 356 + (36 + 4 * K_LIST) / MD_PER_PERIOD  # = 1300.4
PARITY_TOL = 1e-10    # BASELINE.json north_star: forces, energies, virial within 1e-10 relative in fp64


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity_check after the timed region")
    ap.add_argument("--no-c1", action="store_true", help="skip the configs[0] (8192 atoms, NVE, 1000 steps) sub-record")
    ap.add_argument("--lanes", type=int, default=0, help="tiled path: lanes per atom (2/4/8), 0 = default")
    ap.add_argument("--classes", type=int, default=1, help="tiled path: distance-classified lists on/off")
    ap.add_argument("--threads", type=int, default=0, help="tiled path: threads of the pass CTA (512/768), 0 = default")
    ap.add_argument("--pdl", type=int, default=-1, help="programmatic dependent launch of the predictor / passes (0/1), -1 = default")
    ap.add_argument("--bankorder", type=int, default=-1, help="tiled path: bank-aware order of the scanned list classes (0/1), -1 = default")
    ap.add_argument("--stages", type=int, default=0, help="tiled path: pipeline stages of the pass kernel (2/3), 0 = default")
    ap.add_argument("--c3-boxes", type=int, default=512, help="boxes of the configs[2] sub-record (512 x 16 000 atoms; 0 = skip)")
    ap.add_argument("--c4-replicas", type=int, default=96, help="replicas of the configs[3] sub-record (PARREP_Test asks for 100; 96 "
                    "divides over 1/2/4/8 GPUs; 0 = skip)")
    ap.add_argument("--dd-pka-kev", type=float, default=10.0, help="energy of the PKA of the cascade phase of the one-box run (0: skip)")
    ap.add_argument("--dd-cells", type=int, default=200, help="edge (bcc cells) of the single box of the dd_strong sub-record "
                    "(200 -> 16 M atoms; 0 = skip)")
    ap.add_argument("--cpu-budget-s", type=float, default=150.0, help="--impl reference: wall-clock budget of the timed CPU loop")
    return ap.parse_args()


def make_case(cells, seed, nbox=1, potential="w_marinica", temp=600.0):
    import util
    if potential == "cu_setfl":
        return util.fcc_cu_case((cells, cells, cells), seed=seed, nbox=nbox)
    return util.bcc_case((cells, cells, cells), a0=A0, seed=seed, ru_lu=RU_LU, nb_fac=NB_FAC, mxkvois=MXKVOIS,
                         ntab=NTAB, temp=temp, disp=0.02, nbox=nbox)


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


def bind_to_gpu_numa(index):
    """Run this rank (and first-touch its page-locked buffers) on the CPUs of the GPU's NUMA node: with 8 ranks copying
    ~120 MB per job each, buffers on the wrong socket were the e2e limiter at N = 8.  Best effort."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        txt = open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return txt
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU path (oracle/, fast build), loop in C
# ------------------------------------------------------------------------------------------------
def cpu_md(c, half=False, threads=None, epc=True):
    import util
    from oracle import pyorc as O
    md = util.oracle_cpu_md(O, c, half=half, fast=True)
    md.set_threads(threads or os.cpu_count() or 1)
    if epc:
        md.set_epc(EPC["enable"], EPC["te"], EPC["alpha"], EPC["cut"], EPC["he"])
    md.rebuild()
    md.force()
    return md


def cpu_time_periods(md, periods, it0=0):
    """periods x (10 MD steps incl. one rebuild); returns (seconds, next itime)"""
    t0 = time.perf_counter()
    md.run(it0, periods * MD_PER_PERIOD, 1, MD_PER_PERIOD, H)
    return time.perf_counter() - t0, it0 + periods * MD_PER_PERIOD


def cpu_variants(cells=32):
    """serial 1-core runs of the two CPU formulations on a bounded sample: the full list with every
    directed pair (CALFORCE_FS_Force_Table2) and the half list with Newton's third law (CALFORCE_FS_Force_Table)"""
    c = make_case(cells, 4242)
    n = c.xp.shape[0]
    out = {"sample": "bcc W %d atoms (%d^3 cells), 10 MD steps incl. one rebuild, 1 thread" % (n, cells)}
    for name, half in (("serial_full_list", False), ("serial_newton3_half_list", True)):
        md = cpu_md(c, half=half, threads=1)
        dt, _ = cpu_time_periods(md, 1)
        out[name] = n * MD_PER_PERIOD / dt
    return out


def c1_case():
    return make_case(16, 12345)


def cpu_c1(threads):
    """configs[0]: 8192 W atoms, NVE, 1000 steps on the CPU path; HARMIL drift"""
    c = c1_case()
    md = cpu_md(c, threads=threads, epc=False)
    h0 = md.harmil()
    dt, _ = cpu_time_periods(md, 100)
    h1 = md.harmil()
    return {"atoms": int(c.xp.shape[0]), "md_steps": 1000, "value": c.xp.shape[0] * 1000 / dt, "unit": "atom-steps/s",
            "harmil_erg": [h0, h1], "harmil_drift_rel": (h1 - h0) / abs(h0), "threads": threads}


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
                   "list_cutoff_a0": RU_LU * NB_FAC, "rebuild_every": MD_PER_PERIOD, "ntab": NTAB,
                   "l2_policy": "inputs larger than L2 (neighbour list + state ~0.6 GB per MD step, L2 126 MB)",
                   "parallelism": "independent box per GPU, no data-path collective"},
    }


def run_reference(args):
    """The reference's own CPU implementation of the path on the host cores: FULL box of the benchmark, all threads.
    One reference step = ONE list period (10 MD steps incl. its rebuild) of the 10 in a bench step -- a bounded sample in
    time, not in size -- so that K + W steps end within minutes; the metric is a rate, so the unit is the same."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthr = os.cpu_count() or 1
    c = make_case(args.cells, 12346, args.nbox, args.potential)
    n = c.xp.shape[0]
    t_setup = time.perf_counter()
    md = cpu_md(c, threads=nthr)
    t_setup = time.perf_counter() - t_setup
    it = 0
    wdt = 0.0
    for _ in range(max(args.warmup, 1)):
        d, it = cpu_time_periods(md, 1, it)
        wdt += d
    per = wdt / max(args.warmup, 1)
    steps = args.steps
    if per * steps > args.cpu_budget_s:          # never silently: the line says how many steps were really timed
        steps = max(2, int(args.cpu_budget_s / per))
    dt, it = cpu_time_periods(md, steps, it)
    val = n * MD_PER_PERIOD * steps / dt
    line = base_line(args, n)
    line["config"]["atoms_total"] = n
    line["config"]["md_steps_per_step"] = MD_PER_PERIOD
    line["config"]["parallelism"] = "host CPU, OpenMP over atoms, %d threads" % nthr
    sample = ("full box of %d atoms; one reference step = 1 list period = %d MD steps incl. one Cal_NeighboreList2C rebuild "
              "(a bench step of the GPU arm is %d such periods); %d steps timed after %d warm-up, loop inside C (orc_cpu_run), "
              "gcc -O3 -march=native" % (n, MD_PER_PERIOD, PERIODS_PER_STEP, steps, max(args.warmup, 1)))
    line.update({"impl": "reference", "value": val, "steps": steps, "ms_per_step": dt / steps * 1e3,
                 "cpu_baseline": {"value": val, "unit": "atom-steps/s", "cores": nthr, "kind": "port", "sample": sample,
                                  "host_cores_nproc": os.cpu_count(), "setup_list_and_force_s": t_setup},
                 "e2e": {"value": val, "unit": "atom-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0, "n_gpus": args.gpus})
    if steps != args.steps:
        line["steps_requested"] = args.steps
    if not os.environ.get("BENCH_CPU_QUICK"):   # (the GPU arm's cpu_baseline leg asks for the headline number only)
        try:
            line["cpu_baseline"]["variants_1core"] = cpu_variants()
            line["cpu_baseline"]["configs0_c1"] = cpu_c1(nthr)
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"]["variants_error"] = repr(e)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def new_ctx(c, local, args, stream):
    import util
    from msmpscu_b200 import capi
    ctx = capi.Context(local)
    # an explicit torch stream (a non-null handle): the library launches on it and the timing events are recorded on it.
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
    if args.bankorder >= 0:
        ctx.set_option(capi.OPT_TILED_BANKORDER, args.bankorder)
    if args.pdl >= 0:
        ctx.set_option(capi.OPT_PDL, args.pdl)
    ctx.set_option(capi.OPT_TILED_CLASSES, args.classes)
    ctx.epc_set(EPC["enable"], EPC["te"], EPC["alpha"], EPC["cut"], EPC["he"])
    return ctx


def parity_check(ctx, c, n):
    """On the state the timed region left behind: (1) forces / dF/drho of the tiled path against the generic path for
    every atom; (2) both against the CPU oracle on a 10 000-atom range (lists of the generic GPU path, themselves
    bit-checked against the oracle in tests/).  Per-atom relative errors with a floor of 1e-3 x the largest magnitude."""
    import util
    from msmpscu_b200 import capi
    from oracle import pyorc as O
    out = {"tolerance": PARITY_TOL, "metric": "max_i |a_i - b_i| / max(|b_i|, 1e-3 max|b|)", "atoms": n}
    active = ctx.get_option(capi.OPT_ACTIVE_PATH)
    ctx.force(capi.FORCE)
    f_t = ctx.download(capi.F_FP, capi.ORDER_ORIGINAL)
    d_t = ctx.download(capi.F_DEN, capi.ORDER_ORIGINAL)
    prev = ctx.get_option(capi.OPT_FORCE_PATH)
    ctx.set_option(capi.OPT_FORCE_PATH, capi.FORCE_PATH_GENERIC)
    ctx.nlist_build()
    ctx.force(capi.FORCE)
    f_g = ctx.download(capi.F_FP, capi.ORDER_ORIGINAL)
    d_g = ctx.download(capi.F_DEN, capi.ORDER_ORIGINAL)
    out["timed_path"] = "tiled" if active == capi.FORCE_PATH_TILED else "generic"
    out["tiled_vs_generic_force"] = util.atom_relerr(f_t, f_g)
    out["tiled_vs_generic_den"] = util.atom_relerr(d_t, d_g)
    # oracle on a range in the middle of the cell-sorted array (no z wrap inside the range's neighbourhood)
    gid = ctx.download(capi.F_GID, capi.ORDER_CELL) - 1
    xp = ctx.download(capi.F_XP, capi.ORDER_CELL)
    ityp = ctx.download(capi.F_ITYP, capi.ORDER_CELL)
    statu = ctx.download(capi.F_STATU, capi.ORDER_CELL)
    kv, ind = ctx.nlist_copyout(capi.ORDER_CELL)
    npart = min(10000, n)
    ia0 = max(0, n // 2 - npart // 2)
    f_o, d_o = O.force_range(xp, ityp, statu, kv, ind, c.zl, c.ifpd, util.oracle_tables(O, c), ia0, npart)
    sel = gid[ia0:ia0 + npart]
    out["oracle_sample_atoms"] = npart
    out["generic_vs_oracle_force"] = util.atom_relerr(f_g[sel], f_o)
    out["tiled_vs_oracle_force"] = util.atom_relerr(f_t[sel], f_o)
    out["tiled_vs_oracle_den"] = util.atom_relerr(d_t[sel], d_o)
    out["ok"] = bool(max(out["tiled_vs_generic_force"], out["tiled_vs_generic_den"], out["generic_vs_oracle_force"],
                         out["tiled_vs_oracle_force"], out["tiled_vs_oracle_den"]) <= PARITY_TOL)
    ctx.set_option(capi.OPT_FORCE_PATH, prev)
    ctx.nlist_build()
    return out


def gpu_c1(local, args):
    """configs[0]: 8192 W atoms (16^3 bcc cells), NVE, 1000 steps, rebuild every 10; HARMIL drift"""
    import torch
    import util
    from msmpscu_b200 import capi
    c = c1_case()
    n = c.xp.shape[0]
    stream = torch.cuda.current_stream()
    ctx = capi.Context(local)
    ctx.set_stream(stream.cuda_stream)
    ctx.box_set(c.nbox, c.napb, c.boxlow, c.zl, c.ifpd, c.mass)
    ctx.tables_set(util.product_tables(c), c.ru * c.ru)
    ctx.nlist_init(c.nb_rm, c.mxkvois)
    for f, a in ((capi.F_XP, c.xp), (capi.F_XP1, c.xp1), (capi.F_ITYP, c.ityp), (capi.F_STATU, c.statu)):
        ctx.upload(f, a)
    ctx.nlist_build()

    def harmil():
        ctx.force(capi.FORCE | capi.EPOT)
        ctx.ekin()
        ek = ctx.download(capi.F_EKIN)
        return float((ctx.download(capi.F_EPOT).sum() + ek[ek >= 0.0].sum()) / n)

    h0 = harmil()
    ctx.run(0, 100, 1, MD_PER_PERIOD, H)       # warm-up, part of the 1000 steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    ctx.run(100, 900, 1, MD_PER_PERIOD, H)
    e1.record(stream)
    torch.cuda.synchronize()
    h1 = harmil()
    ctx.close()
    return {"atoms": n, "md_steps": 1000, "value": n * 900 / (e0.elapsed_time(e1) * 1e-3), "unit": "atom-steps/s",
            "harmil_erg": [h0, h1], "harmil_drift_rel": (h1 - h0) / abs(h0)}


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
    numa = bind_to_gpu_numa(local)
    if world > 1:
        import datetime
        # (a rank that dies must not leave the others waiting ten minutes in a collective)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))

    c = make_case(args.cells, 12346 + rank, args.nbox, args.potential)
    n = c.xp.shape[0]
    # two device contexts on two streams: A carries the resident measurement, A and B alternate in the e2e pipeline
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.set_stream(streams[0])
    ctxs = [new_ctx(c, local, args, s) for s in streams]
    # host buffers in page-locked memory, reference layout XP(N,3) column-major; one set per context
    host = []
    for ctx in ctxs:
        hx = torch.from_numpy(capi.colmajor(c.xp)).pin_memory()
        hv = torch.from_numpy(capi.colmajor(c.xp1)).pin_memory()
        hf = torch.empty(3 * n, dtype=torch.float64).pin_memory()
        host.append((hx, hv, hf))
        ctx.upload_raw(capi.F_XP, hx.data_ptr())
        ctx.upload_raw(capi.F_XP1, hv.data_ptr())
        ctx.upload(capi.F_ITYP, c.ityp)
        ctx.upload(capi.F_STATU, c.statu)
        ctx.nlist_build()
        ctx.force(capi.FORCE)
    ctx = ctxs[0]
    stream = streams[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(ms):
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    def timed(fn, k, finish=None):
        """K calls of fn between two events on stream A; `finish` drains the pipeline (and joins stream B) before the end event"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for i in range(k):
            fn(i)
        if finish:
            finish()
        e1.record(stream)
        barrier()
        return reduce_max(e0.elapsed_time(e1))

    itime = [0, 0]

    def step_resident(_):
        ctx.run(itime[0], MD_PER_STEP, 1, MD_PER_PERIOD, H)
        itime[0] += MD_PER_STEP

    pending = [False, False]

    def make_job(md_steps):
        def job(i):
            # one independent job through the reference-facing calls with HOST buffers:
            # CopyIn (XP, XP1) -> md_steps x For_One_Step -> CopyOut (XP, XP1, FP); ONE host synchronisation per job
            k = i & 1
            cx, (hx, hv, hf) = ctxs[k], host[k]
            if pending[k]:
                cx.sync()                      # the previous job of this context has landed in its host buffers
            cx.upload_raw(capi.F_XP, hx.data_ptr())
            cx.upload_raw(capi.F_XP1, hv.data_ptr())
            cx.run_async(itime[k], md_steps, 1, MD_PER_PERIOD, H)
            itime[k] += md_steps
            cx.download_raw_async(capi.F_XP, hx.data_ptr())
            cx.download_raw_async(capi.F_XP1, hv.data_ptr())
            cx.download_raw_async(capi.F_FP, hf.data_ptr())
            pending[k] = True
        return job

    def drain():
        for k in (0, 1):
            if pending[k]:
                ctxs[k].sync()
                pending[k] = False
        ev = torch.cuda.Event()
        ev.record(streams[1])
        streams[0].wait_event(ev)

    def job_serial(_):
        hx, hv, hf = host[0]
        ctx.upload_raw(capi.F_XP, hx.data_ptr())
        ctx.upload_raw(capi.F_XP1, hv.data_ptr())
        ctx.run(itime[0], MD_PER_PERIOD, 1, MD_PER_PERIOD, H)
        itime[0] += MD_PER_PERIOD
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

    # end to end through host buffers (pipelined over the two contexts)
    job = make_job(MD_PER_STEP)
    job(0); job(1); drain()
    torch.cuda.synchronize()
    ms_e2e = timed(job, args.steps, drain)
    e2e_val = world * n * MD_PER_STEP * args.steps / (ms_e2e * 1e-3)
    # diagnostics: a transfer every list period (10 MD steps, round 1's e2e unit), pipelined and serial
    job10 = make_job(MD_PER_PERIOD)
    job10(0); job10(1); drain()
    k10 = max(10, args.steps)
    ms_e2e10 = timed(job10, k10, drain)
    job_serial(0)
    ms_ser10 = timed(job_serial, k10)
    e2e_diag = {"transfer_every_md_steps": MD_PER_PERIOD,
                "pipelined_two_contexts": world * n * MD_PER_PERIOD * k10 / (ms_e2e10 * 1e-3),
                "serial_one_context": world * n * MD_PER_PERIOD * k10 / (ms_ser10 * 1e-3),
                "ms_per_job_pipelined": ms_e2e10 / k10, "ms_per_job_serial": ms_ser10 / k10, "numa_cpulist": numa}

    # per-kernel device times (CUDA events on the launching stream) for the roofline
    ctx.prof_reset()
    ctx.prof_enable(True)
    step_resident(0)
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
            "traffic": traffic, "traffic_source": "profiles/traffic.json (ncu --set full capture of this workload, committed; "
                                                  "not re-measured by this run)" if traffic else None,
            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s",
            "algorithmic_bytes_per_atom": BYTES_PASS1 if dom == "pass1" else BYTES_PASS2,
            "avg_launch_ms": tms / nl if nl else None,
            "kernel_share_of_step": tms / tot if tot else None,
            "whole_step": {"achieved": value / world * BYTES_STEP / 1e9, "frac": value / world * BYTES_STEP / 1e9 / peak,
                           "bytes_per_atom_step": BYTES_STEP},
            "per_class_ms_per_md_step": {k: v[1] / MD_PER_STEP for k, v in prof.items() if v[0]}}
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
                         "d2h_bytes_per_step": 3 * 24 * n, "ms_per_step": ms_e2e / args.steps,
                         "how": "independent jobs alternate between two device contexts / streams; per job: CopyIn XP, XP1 from "
                                "page-locked host memory, %d MD steps, CopyOut XP, XP1, FP; one host synchronisation per job" % MD_PER_STEP},
                 "e2e_diagnostics": e2e_diag,
                 "roofline": roof, "force_path": args.path, "timed_region_s": ms * 1e-3})
    line["config"]["atoms_total"] = n * world

    if rank == 0 and not args.no_parity:
        try:
            line["parity_check"] = parity_check(ctx, c, n)
        except Exception as e:  # a parity failure must be visible in the line, not kill the measurement
            line["parity_check"] = {"ok": False, "error": repr(e)}
    if rank == 0 and world == 1 and not args.no_c1:
        try:
            line["configs0_c1"] = gpu_c1(local, args)
        except Exception as e:  # pragma: no cover
            line["configs0_c1"] = {"error": repr(e)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # in a fresh process (no CUDA context, no second OpenMP runtime beside it): the same code path as --impl reference
        import subprocess
        try:
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--cells", str(args.cells),
                                  "--nbox", str(args.nbox), "--potential", args.potential, "--steps", "2", "--warmup", "1"],
                                 capture_output=True, text=True, timeout=900, env=dict(os.environ, BENCH_CPU_QUICK="1"))
            ref = json.loads(out.stdout.strip().splitlines()[-1])
            line["cpu_baseline"] = ref["cpu_baseline"]
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": "atom-steps/s", "cores": os.cpu_count(), "kind": "port", "error": repr(e)}
    for cx in ctxs:
        cx.close()
    del ctxs, host
    torch.cuda.empty_cache()
    if args.c3_boxes > 0:
        # configs[2] on the same N GPUs: 512 boxes x 16 000 atoms sharded by the multi-box dispatcher (every rank takes part)
        try:
            line["multibox_c3"] = c3_measure(args, args.c3_boxes, 20, 5, 2)
        except Exception as e:  # pragma: no cover
            line["multibox_c3"] = {"error": repr(e)}
    if args.c4_replicas > 0:
        # configs[3] on the same N GPUs: PARREP replicas sharded per GPU, event detection on the device
        try:
            line["parrep_c4"] = c4_measure(args, args.c4_replicas, 2, 500)
        except Exception as e:  # pragma: no cover
            line["parrep_c4"] = {"error": repr(e)}
    if args.dd_cells > 0:
        # configs[4] family on the same N GPUs: one 16 M-atom box, strong scaling (collective: every rank takes part)
        try:
            line["dd_strong"] = dd_measure(args, args.dd_cells, 5, 2)
        except Exception as e:  # pragma: no cover
            line["dd_strong"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def dd_measure(args, cells, blocks, warm):
    """ONE box over all ranks (slab decomposition inside the library, csrc/mdb_dd.cu): the box is the same on every rank
    (same seed) for the initial build; afterwards a rank touches only its slab and its ghost layers.  Strong scaling:
    value = atoms of the box x MD steps / max-over-ranks device time.  One block = one list period (10 MD steps, one
    local rebuild).  Requires torch.cuda.set_device / the process group to be set up by the caller."""
    import torch
    import torch.distributed as dist
    import util
    from msmpscu_b200 import capi
    from msmpscu_b200.domain import SlabDomain

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    c = make_case(cells, 777)
    n = c.xp.shape[0]
    ctx = capi.Context(local)
    ctx.box_set(c.nbox, c.napb, c.boxlow, c.zl, c.ifpd, c.mass)
    ctx.set_option(capi.OPT_FORCE_PATH, capi.FORCE_PATH_TILED)
    ctx.tables_set(util.product_tables(c), c.ru * c.ru)
    ctx.nlist_init(c.nb_rm, c.mxkvois)
    ctx.epc_set(EPC["enable"], EPC["te"], EPC["alpha"], EPC["cut"], EPC["he"])
    ctx.upload(capi.F_XP, c.xp); ctx.upload(capi.F_XP1, c.xp1)
    ctx.upload(capi.F_ITYP, c.ityp); ctx.upload(capi.F_STATU, c.statu)
    del c.xp1
    dom = SlabDomain(ctx, local)
    dom.rebuild()
    dom.force(capi.FORCE)
    stream = dom.stream
    itime = [0]

    def block():
        dom.run(itime[0], MD_PER_PERIOD, 1, MD_PER_PERIOD, H)
        itime[0] += MD_PER_PERIOD

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(warm, 1)):
        block()
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = ctx.launch_count()
    barrier()
    e0.record(stream)
    for _ in range(blocks):
        block()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = ctx.launch_count() - l0
    clocks = sampler.result()
    # per-phase device times of one more block (CUDA events of the library's profiling classes, not part of the timed region)
    ctx.prof_reset(); ctx.prof_enable(True)
    block()
    prof = {k: round(v[1], 3) for k, v in ctx.prof_get().items() if v[0]}
    ctx.prof_enable(False)
    if world > 1:
        allp = [None] * world
        dist.all_gather_object(allp, prof)
        prof = {k: [p.get(k, 0.0) for p in allp] for k in prof}
    a0, a1 = dom.owned()
    cascade = dd_cascade(args, c, ctx, dom, stream, itime, blocks, barrier, world) if args.dd_pka_kev > 0 else None
    rec = {"value": n * MD_PER_PERIOD * blocks / (ms * 1e-3), "unit": "atom-steps/s", "scaling": "strong", "atoms_total": n,
           "atoms_owned_rank0": a1 - a0, "n_gpus": world, "blocks": blocks, "md_steps_per_block": MD_PER_PERIOD,
           "ms_per_block": ms / blocks, "gpu_launches": int(launches), "clocks": clocks,
           "phase_ms_per_block_by_rank": prof,
           "workload": "configs[4] family: ONE bcc W box of %d atoms (%d^3 cells) cut into %d z-slabs; per step two ghost-layer "
                       "exchanges of {x,y,z,den} records (ncclSend/ncclRecv enqueued by the library), per %d steps a local rebuild "
                       "of the rank's slab (no broadcast, no all-atom sort)" % (n, cells, world, MD_PER_PERIOD)}
    if cascade:
        rec["cascade"] = cascade
    ctx.close()
    return rec


def dd_cascade(args, c, ctx, dom, stream, itime, blocks, barrier, world):
    """configs[4] as SURVEY.md 8(d) specifies it: the same box with ONE primary knock-on atom at the centre (velocity along <135>),
    run the way the reference runs cascades (examples/Cascade_Test/CtrlFile300K_LOC_T.dat): displacement-limited time step
    (&STEPSIZE flag -1, hmx 0.5 fs, dmx 0.05 A), EPC + electronic stopping on the atoms, list rebuilt every 10 steps.  The
    stopping table is of the Lindhard-Scharff form S(E) = k sqrt(E) with k for W in W (the reference builds its LS_Z85 table
    from the same law); local-density model and per-atom energy-loss bookkeeping as in that file (&MDEN = -1, &SAVEELOSS)."""
    import torch
    import torch.distributed as dist
    from msmpscu_b200 import capi

    n = c.xp.shape[0]
    centre = c.boxlow + 0.5 * c.zl
    ipka = int(np.argmin(np.sum((c.xp - centre) ** 2, axis=1))) + 1                    # ORIGINAL id (1-based)
    ekev = float(args.dd_pka_kev)
    ctx.pka_insert(ipka, ekev * 1000.0 * CP_EVERG, [1.0, 3.0, 5.0])
    etab = np.linspace(20.0, 2.0e5, 20001) * CP_EVERG                                 # uniform grid as the reference's tables; &EMIN 20 eV
    k_ls = 3.25e-16 * CP_EVERG                                                       # erg cm^2 per sqrt(eV): the W->W column of the reference's examples/Cascade_Test/Stopping_table.stp (LS_Z85) is 1.0278e-17 keV cm^2 x sqrt(E / keV)
    stab = (k_ls * np.sqrt(etab / CP_EVERG)).reshape(-1, 1)
    mden = n / float(np.prod(c.zl))
    ctx.stopping_set(etab, stab, np.array([[1]]), [1], [mden])
    ctx.stopping_options(local_density=True, save_eloss=True)                        # &MDEN = -1, &SAVEELOSS = "YES" of the reference's cascade file
    sched = capi.Sched(-1, H, H, 0.05e-8, MD_PER_PERIOD, MD_PER_PERIOD, 100)          # dmx = 0.05 Angstrom (MD_Gvar.F90:947)
    st = {"h": H, "t": 0.0}

    def block():
        _, st["h"], st["t"] = dom.run_sched(itime[0], MD_PER_PERIOD, 1, sched, st["h"], st["t"])
        itime[0] += MD_PER_PERIOD

    block()                                                                          # warm-up: the first steps of the PKA
    h_first = st["h"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(blocks):
        block()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))), dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ctx.prof_reset(); ctx.prof_enable(True)
    block()
    prof = {k: round(v[1], 3) for k, v in ctx.prof_get().items() if v[0]}
    ctx.prof_enable(False)
    return {"value": n * MD_PER_PERIOD * blocks / (ms * 1e-3), "unit": "atom-steps/s", "ms_per_block": ms / blocks, "blocks": blocks,
            "pka_kev": ekev, "pka_original_id": ipka, "h_fs_after_first_block": h_first * 1e15, "h_fs_last": st["h"] * 1e15,
            "simulated_fs": st["t"] * 1e15, "phase_ms_per_block_rank0": prof,
            "inelastic_loss_ev_this_rank": float(ctx.stopping_eloss().sum() / CP_EVERG),
            "workload": "the same box with one %.0f keV PKA at the centre along <135>: displacement-limited step (hmx 0.5 fs, dmx 0.05 A, "
                        "checked every step: one mask kernel + 4-byte read-back, OR-ed over the ranks), EPC + electronic stopping "
                        "(Lindhard-Scharff form table, local-density model, per-atom energy loss accumulated) between friction and corrector, per-tile displacement "
                        "bounds for the distance-class shortcut" % ekev}


def c3_measure(args, nbox_total, cells, blocks, warm):
    """configs[2]: `nbox_total` independent boxes of 2*cells^3 atoms sharded over the ranks with MultiBoxDispatcher (contiguous
    blocks of boxes per GPU, concatenated as MULTIBOX inside a rank; no inter-GPU traffic per step), per-box temperatures gathered
    over the ranks at the end (the only collective).  No Fe table file ships with the reference: W Marinica tables
    and lattice are used and the record says so.  value = all boxes' atoms x MD steps / max-over-ranks device time."""
    import torch
    import torch.distributed as dist
    import util
    from msmpscu_b200 import capi
    from msmpscu_b200.multibox import MultiBoxDispatcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    disp = MultiBoxDispatcher(nbox_total)
    c = make_case(cells, 20000 + disp.first, nbox=disp.count)
    n = c.xp.shape[0]
    stream = torch.cuda.Stream()
    ctx = capi.Context(local)
    ctx.set_stream(stream.cuda_stream)
    ctx.box_set(c.nbox, c.napb, c.boxlow, c.zl, c.ifpd, c.mass)
    ctx.tables_set(util.product_tables(c), c.ru * c.ru)
    ctx.nlist_init(c.nb_rm, c.mxkvois)
    ctx.epc_set(EPC["enable"], EPC["te"], EPC["alpha"], EPC["cut"], EPC["he"])
    for f, a in ((capi.F_XP, c.xp), (capi.F_XP1, c.xp1), (capi.F_ITYP, c.ityp), (capi.F_STATU, c.statu)):
        ctx.upload(f, a)
    ctx.nlist_build()
    ctx.force(capi.FORCE)
    it = [0]

    def block():
        ctx.run(it[0], MD_PER_PERIOD, 1, MD_PER_PERIOD, H)
        it[0] += MD_PER_PERIOD

    for _ in range(max(warm, 1)):
        block()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(blocks):
        block()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    tbox = disp.gather_box_scalars(ctx.box_temperatures(c.nbox).reshape(-1, 1))   # (nbox_total, 1) on every rank
    path = ctx.get_option(capi.OPT_ACTIVE_PATH)
    ctx.close()
    ntot = nbox_total * c.napb
    return {"value": ntot * MD_PER_PERIOD * blocks / (ms * 1e-3), "unit": "atom-steps/s", "scaling": "strong", "n_gpus": world,
            "boxes_total": nbox_total, "boxes_this_rank": disp.count, "atoms_per_box": int(c.napb), "atoms_total": int(ntot),
            "blocks": blocks, "md_steps_per_block": MD_PER_PERIOD, "ms_per_block": ms / blocks,
            "force_path": "tiled" if path == capi.FORCE_PATH_TILED else "generic",
            "box_temperature_K": {"boxes_gathered": int(tbox.shape[0]), "mean": float(tbox.mean()), "min": float(tbox.min()),
                                  "max": float(tbox.max())},
            "workload": "configs[2]: %d independent boxes x %d atoms (%d^3 bcc cells) sharded over %d GPU(s) as contiguous blocks of "
                        "boxes (MULTIBOX inside a rank), NVT via EPC, no inter-GPU traffic per step; W Marinica EAM2 tables stand in "
                        "for the Fe EAM_NIST file the reference does not ship" % (nbox_total, c.napb, cells, world)}


def c4_measure(args, nrep_total, cycles, md_steps):
    """configs[3]: PARREP replicas (2000 W + 1 H, Bonny EAM1, examples/PARREP_Test control values: list cutoff 1.6 RU, MAXNB 400,
    rebuild every 10, quench ST 1000 steps, DRTOL 0.03 LU) sharded over the ranks by the multi-box dispatcher; per cycle every
    rank runs `md_steps` MD steps on its replicas, then the event detection ON THE DEVICE (save, steepest-descent quench,
    compare with the starting configuration, restore + rebuild), and the per-replica event flags are gathered over the ranks --
    the only collective.  value = replica atoms x MD steps / max-over-ranks wall time, detection included."""
    import torch
    import torch.distributed as dist
    import util
    from msmpscu_b200 import capi
    from msmpscu_b200.multibox import MultiBoxDispatcher

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    disp = MultiBoxDispatcher(nrep_total)
    c = util.parrep_case(disp.count, seed=3000 + disp.first)
    xini = util.neb_case("react").xp
    ctx = capi.Context(local)
    ctx.box_set(c.nbox, c.napb, c.boxlow, c.zl, c.ifpd, c.mass)
    for f, a in ((capi.F_XP, c.xp), (capi.F_XP1, c.xp1), (capi.F_ITYP, c.ityp), (capi.F_STATU, c.statu)):
        ctx.upload(f, a)
    ctx.tables_set(util.product_tables(c), c.ru * c.ru)
    ctx.nlist_init(c.nb_rm, c.mxkvois)
    ctx.nlist_build()
    ctx.epc_set([1, 1], [300.0, 300.0], [1.0e-12] * 2, [0.1] * 2, [100.0 * 1.60219e-12] * 2)
    ctx.force(capi.FORCE)
    ctx.thermalize(600.0, 20240101 + disp.first, 0)
    ctx.run(0, 20, 1, MD_PER_PERIOD, H)
    # warm-up of the event check: its device buffers (saved replicas, comparison work space) are allocated at first use, and a
    # cudaMalloc right after another context of this process released several GB has been seen to take 0.3 s
    ctx.state_save()
    ctx.compare(xini, 0.03 * c.rr, nbox=disp.count)
    ctx.state_restore()
    ctx.sync()
    it = 20
    events = 0
    quench = []
    t_md = t_det = 0.0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(cycles):
        a = time.perf_counter()
        ctx.run(it, md_steps, 1, MD_PER_PERIOD, H)
        it += md_steps
        b = time.perf_counter()
        ctx.state_save()
        ctx.sync()
        t1 = time.perf_counter()
        q_iflag, q_move, q_de = ctx.steepest(1000, 0.1, 0.1 * c.rr, 1.0e-5 * c.rr, 1.0e-5 * 1.60219e-12)
        t2 = time.perf_counter()
        fb, ibt, ncb = ctx.compare(xini, 0.03 * c.rr, nbox=disp.count)
        t3 = time.perf_counter()
        ctx.state_restore()
        ctx.sync()
        t4 = time.perf_counter()
        quench.append({"iflag": int(q_iflag), "save_s": round(t1 - b, 4), "quench_s": round(t2 - t1, 4), "compare_s": round(t3 - t2, 4),
                       "restore_s": round(t4 - t3, 4)})
        flags = disp.gather_box_scalars(fb.reshape(-1, 1).astype(np.float64))      # every rank sees every replica's flag
        events += int(flags.sum())
        d = time.perf_counter()
        t_md += b - a
        t_det += d - b
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt, t_md, t_det], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, t_md, t_det = (float(v) for v in t.tolist())
    path = ctx.get_option(capi.OPT_ACTIVE_PATH)
    ctx.close()
    ntot = nrep_total * c.napb
    return {"value": ntot * md_steps * cycles / dt, "unit": "atom-steps/s", "scaling": "strong", "n_gpus": world,
            "replicas_total": nrep_total, "replicas_this_rank": disp.count, "atoms_per_replica": int(c.napb), "cycles": cycles,
            "md_steps_per_cycle": md_steps, "wall_s": dt, "md_s": t_md, "event_detection_s": t_det, "events_flagged": events,
            "md_only_atom_steps_per_s": ntot * md_steps * cycles / t_md, "quench_per_cycle_rank0": quench,
            "force_path": "tiled" if path == capi.FORCE_PATH_TILED else "generic",
            "workload": "configs[3]: %d PARREP replicas x %d atoms (2000 W + 1 H, Bonny EAM1), %d cycles of %d MD steps + event detection "
                        "(device-side save / ST quench / compare / restore), replicas sharded over %d GPU(s), event flags gathered over the ranks"
                        % (nrep_total, c.napb, cycles, md_steps, world)}


def run_dd(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
    rec = dd_measure(args, args.cells, args.steps, max(args.warmup, 3))
    line = base_line(args, rec["atoms_total"])
    line["scaling"] = "strong"
    line["config"].update({"workload": rec["workload"], "atoms_per_gpu": rec["atoms_owned_rank0"], "atoms_total": rec["atoms_total"],
                           "md_steps_per_step": MD_PER_PERIOD, "parallelism": "z-slab domain decomposition, 1 ghost cell layer per side"})
    line.update({"value": rec["value"], "ms_per_step": rec["ms_per_block"], "clocks": rec["clocks"], "gpu_launches": rec["gpu_launches"],
                 "force_path": "tiled", "mode": "dd", "phase_ms_per_block_by_rank": rec["phase_ms_per_block_by_rank"]})
    if rec.get("cascade"):
        line["cascade"] = rec["cascade"]
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries exactly ONE line, the JSON record: libraries that print there (NCCL's version banner, ...) are sent to
    # stderr for the duration of the run
    _real_stdout = os.dup(1)
    sys.stdout.flush()
    os.dup2(2, 1)
    _print = print

    def print(*args, **kw):  # noqa: A001  (module-level print of the record goes to the real stdout)
        if kw.get("file") is None:
            os.write(_real_stdout, (" ".join(str(x) for x in args) + "\n").encode())
        else:
            _print(*args, **kw)

    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.dd:
        run_dd(a)
    else:
        run_ours(a)
