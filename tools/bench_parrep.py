#!/usr/bin/env python
"""BASELINE configs[3] shape on one GPU: R replicas of the 2000 W + 1 H box (Bonny EAM1, examples/PARREP_Test control values:
list cutoff 1.6 x RU, MAXNB 400, rebuild every 10 steps, event check every 500 steps, quench 1000 steps) as MULTIBOX.
Measures the three device phases of a PARREP cycle through the C ABI: thermalise, 500 MD steps, quench (steepest descent as
the shipped control file asks, and L-BFGS as BASELINE.json words it).  One JSON line.
usage: python tools/bench_parrep.py [replicas=100]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from msmpscu_b200 import capi  # noqa: E402


def main():
    nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    c = util.parrep_case(nrep)
    n = c.xp.shape[0]
    out = {"workload": "configs[3] shape: %d replicas x 2001 atoms (2000 W + 1 H, Bonny EAM1), list cutoff 1.6 RU, MAXNB 400" % nrep,
           "atoms": n}
    for path, name in ((capi.FORCE_PATH_AUTO, "auto"),):
        ctx = util.make_ctx(c, build=False, force_path=path)
        if os.environ.get("MDB_LANES"):
            ctx.set_option(capi.OPT_TILED_LANES, int(os.environ["MDB_LANES"]))
        if os.environ.get("MDB_THREADS"):
            ctx.set_option(capi.OPT_TILED_THREADS, int(os.environ["MDB_THREADS"]))
        ctx.nlist_build()
        out["active_path"] = "tiled" if ctx.get_option(capi.OPT_ACTIVE_PATH) == capi.FORCE_PATH_TILED else "generic"
        ctx.force(capi.FORCE)
        ctx.thermalize(600.0, 20240101, 0)
        ctx.run(0, 20, 1, 10, 0.5e-15)               # warm-up
        ctx.sync()
        t0 = time.perf_counter()
        ctx.thermalize(600.0, 20240101, 1)
        ctx.sync()
        out["thermalize_ms"] = (time.perf_counter() - t0) * 1e3
        ctx.prof_reset(); ctx.prof_enable(True)
        ctx.run(20, 50, 1, 10, 0.5e-15)
        out["md_device_ms_per_step_by_class"] = {k: round(v[1] / 50, 4) for k, v in ctx.prof_get().items() if v[0]}
        ctx.prof_enable(False)
        ctx.sync()
        t0 = time.perf_counter()
        ctx.run(70, 500, 1, 10, 0.5e-15)
        ctx.sync()
        dt = time.perf_counter() - t0
        out["md_500_steps_ms"] = dt * 1e3
        out["md_atom_steps_per_s"] = n * 500 / dt
        x = ctx.download(capi.F_XP)
        v = ctx.download(capi.F_XP1)
        rr = c.rr
        for meth in ("steepest", "lbfgs"):
            ctx.upload(capi.F_XP, x); ctx.upload(capi.F_XP1, v)
            ctx.nlist_build(); ctx.force(capi.FORCE)
            f0 = np.abs(ctx.download(capi.F_FP)).max()
            ctx.sync()
            ctx.prof_reset(); ctx.prof_enable(True)
            t0 = time.perf_counter()
            if meth == "steepest":   # &QUENCHSTEP 1000 "ST", &STEPBOUND 0.00001 / 0.1, &DELTAPOT 0.00001 (CtrlFile300K.dat)
                fl, mm, de = ctx.steepest(1000, 0.1, 0.1 * rr, 1.0e-5 * rr, 1.0e-5 * util.CP_EVERG)
                info = {"iflag": fl}
            else:
                fl, nfg, nit = ctx.lbfgs(1000, 7, 0.0, 1.0e-4 * f0)
                info = {"iflag": fl, "force_evaluations": nfg, "accepted_steps": nit}
            ctx.sync()
            dt = time.perf_counter() - t0
            info_prof = {k: [v[0], round(v[1], 3)] for k, v in ctx.prof_get().items() if v[0]}
            ctx.prof_enable(False)
            ctx.force(capi.FORCE | capi.EPOT)
            info.update({"ms": dt * 1e3, "device_ms_by_class": info_prof, "max_force_ratio": float(np.abs(ctx.download(capi.F_FP)).max() / f0)})
            out["quench_" + meth] = info
        ctx.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
