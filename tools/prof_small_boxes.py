"""Per-kernel-class time of the MD stretch of the two many-small-boxes workloads (configs[2] multibox, configs[3] PARREP
replicas) on one GPU: where a step goes when the boxes are small.  Diagnostic; prints one JSON line per workload.
usage: python tools/prof_small_boxes.py [nrep] [nbox]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import util  # noqa: E402
from msmpscu_b200 import capi  # noqa: E402

H = 0.5e-15


def measure(c, label, steps=200, epc_groups=1, opts=()):
    ctx = capi.Context(0)
    ctx.box_set(c.nbox, c.napb, c.boxlow, c.zl, c.ifpd, c.mass)
    for f, a in ((capi.F_XP, c.xp), (capi.F_XP1, c.xp1), (capi.F_ITYP, c.ityp), (capi.F_STATU, c.statu)):
        ctx.upload(f, a)
    ctx.tables_set(util.product_tables(c), c.ru * c.ru)
    for o, v in opts:
        ctx.set_option(o, v)
    ctx.nlist_init(c.nb_rm, c.mxkvois)
    ctx.nlist_build()
    g = epc_groups
    ctx.epc_set([1] * g, [300.0] * g, [1.0e-12] * g, [0.1] * g, [100.0 * 1.60219e-12] * g)
    ctx.force(capi.FORCE)
    ctx.thermalize(600.0, 20240101, 0)
    ctx.run(0, 40, 1, 10, H)
    ctx.sync()
    t0 = time.perf_counter()
    ctx.run(40, steps, 1, 10, H)
    ctx.sync()
    wall = time.perf_counter() - t0
    ctx.prof_enable(True)
    ctx.prof_reset()
    ctx.run(40 + steps, steps, 1, 10, H)
    ctx.sync()
    p = ctx.prof_get()
    ctx.prof_enable(False)
    nc3, nc, mx = ctx.cellinfo()
    out = {"workload": label, "atoms": int(c.nbox * c.napb), "boxes": int(c.nbox), "cells_per_box_edge": [int(x) for x in nc3],
           "opts": [list(o) for o in opts], "path": ctx.get_option(capi.OPT_ACTIVE_PATH),
           "ms_per_step_wall": 1e3 * wall / steps, "atom_steps_per_s": c.nbox * c.napb * steps / wall,
           "ms_per_step_by_class": {k: round(v[1] / steps, 5) for k, v in p.items() if v[0]},
           "launches_per_step": {k: v[0] / steps for k, v in p.items() if v[0]}}
    print(json.dumps(out), flush=True)
    ctx.close()


if __name__ == "__main__":
    nrep = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    nbox = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    c = util.parrep_case(nrep, seed=3000)
    measure(c, "parrep %d x 2001" % nrep, epc_groups=2)
    if len(sys.argv) > 3 and sys.argv[3] == "quick":
        sys.exit(0)
    if len(sys.argv) > 3 and sys.argv[3] == "lanes":
        measure(c, "parrep %d x 2001, lanes 8" % nrep, epc_groups=2, opts=((capi.OPT_TILED_LANES, 8),))
        measure(c, "parrep %d x 2001, lanes 8, 512 threads" % nrep, epc_groups=2, opts=((capi.OPT_TILED_LANES, 8), (capi.OPT_TILED_THREADS, 512)))
        measure(c, "parrep %d x 2001, 512 threads" % nrep, epc_groups=2, opts=((capi.OPT_TILED_THREADS, 512),))
        sys.exit(0)
    measure(c, "parrep %d x 2001, lanes 8" % nrep, epc_groups=2, opts=((capi.OPT_TILED_LANES, 8),))
    measure(c, "parrep %d x 2001, bank order on" % nrep, epc_groups=2, opts=((capi.OPT_TILED_BANKORDER, 1),))
    measure(c, "parrep %d x 2001, 512 threads" % nrep, epc_groups=2, opts=((capi.OPT_TILED_THREADS, 512),))
    measure(util.bcc_case((20, 20, 20), seed=20000, temp=600.0, nbox=nbox), "multibox %d x 16000" % nbox)
