"""Force + virial evaluation (pCalPTensor, MD_EAM_ForceTable_GPU.F90:1366) on configs[1] (1 024 000 W atoms): the tiled
pass with its virial epilogue against the generic CALPTENSOR kernels, CUDA events on the launching stream, inputs larger
than L2.  Prints one JSON line.  Usage: python tools/bench_virial.py [--cells 80] [--reps 20]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=80)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    import torch
    import bench
    import util
    from msmpscu_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench_virial.py: no CUDA device")
    c = bench.make_case(a.cells, 4711)
    n = c.xp.shape[0]
    out = {"workload": "configs[1] shape: bcc W %d atoms, force + virial per call" % n, "reps": a.reps}
    vts = {}
    for name, path in (("generic", 1), ("tiled", 2)):
        ctx = util.make_ctx(c, build=False, force_path=path)
        stream = torch.cuda.Stream()          # non-null handle: the library launches on it, the events are recorded on it
        torch.cuda.set_stream(stream)
        ctx.set_stream(stream.cuda_stream)
        ctx.nlist_build()
        for _ in range(3):
            vt = ctx.force(capi.FORCE | capi.VIRIAL)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(a.reps):
            vt = ctx.force(capi.FORCE | capi.VIRIAL)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(stream)
        for _ in range(a.reps):
            ctx.force(capi.FORCE)
        f1.record(stream)
        torch.cuda.synchronize()
        out[name] = {"force_virial_ms": ms, "force_only_ms": f0.elapsed_time(f1) / a.reps}
        vts[name] = np.asarray(vt, dtype=np.float64)
        ctx.close()
    out["virial_relerr_tiled_vs_generic"] = float(np.max(np.abs(vts["tiled"] - vts["generic"])) / np.max(np.abs(vts["generic"])))
    out["speedup"] = out["generic"]["force_virial_ms"] / out["tiled"]["force_virial_ms"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
