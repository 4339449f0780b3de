"""GPU tests added in round 2: parity at the size the bench reports, the asynchronous C-ABI calls, the capacity
check at every rebuild inside mdb_run (transient overflow -> generic path), the materialised reference-shape device
views, and the bank-aware list order.  Everything goes through the C ABI; the CPU oracle is the checker."""
import ctypes as C

import numpy as np
import pytest

import util
from msmpscu_b200 import capi

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _dev_to_host(ptr, count, dtype):
    """cudaMemcpy of `count` items from a raw device pointer (through torch: plumbing only)"""
    import torch

    class A:
        pass
    a = A()
    a.__cuda_array_interface__ = {"shape": (count,), "typestr": np.dtype(dtype).str, "data": (int(ptr), False), "version": 2}
    torch.cuda.synchronize()
    return torch.as_tensor(a, device="cuda").cpu().numpy()


def test_headline_size_tiled_vs_generic_and_oracle(oracle):
    """configs[1] itself: 1 024 000 W atoms.  The tiled plan (tile width, halo capacity, table window) of this size is not
    the plan of any small case, so the check runs at this size: all atoms tiled vs generic, 10 000 atoms vs the oracle,
    PER-ATOM relative error (floor 1e-3 of the largest force)."""
    import bench
    c = bench.make_case(80, 12346)
    n = c.xp.shape[0]
    assert n == 1024000
    ctx = util.make_ctx(c)
    assert ctx.get_option(capi.OPT_ACTIVE_PATH) == capi.FORCE_PATH_TILED
    ctx.epc_set(bench.EPC["enable"], bench.EPC["te"], bench.EPC["alpha"], bench.EPC["cut"], bench.EPC["he"])
    ctx.force(capi.FORCE)
    ctx.run(0, 12, 1, 10, bench.H)           # two rebuilds, atoms have left their lattice sites
    r = bench.parity_check(ctx, c, n)
    ctx.close()
    assert r["timed_path"] == "tiled"
    for k in ("tiled_vs_generic_force", "tiled_vs_generic_den", "generic_vs_oracle_force", "tiled_vs_oracle_force", "tiled_vs_oracle_den"):
        assert r[k] <= TOL, (k, r)


def test_async_run_and_download_equal_the_blocking_calls():
    import torch
    c = util.bcc_case((9, 9, 9), seed=31)
    n = c.xp.shape[0]
    h = 0.5e-15
    a = util.make_ctx(c)
    b = util.make_ctx(c)
    a.force(capi.FORCE); b.force(capi.FORCE)
    oob = a.run(0, 23, 1, 10, h)
    xa, va, fa = a.download(capi.F_XP), a.download(capi.F_XP1), a.download(capi.F_FP)
    hx, hv, hf = (torch.empty(3 * n, dtype=torch.float64).pin_memory() for _ in range(3))
    assert b.run_async(0, 23, 1, 10, h) == 0
    b.download_raw_async(capi.F_XP, hx.data_ptr())
    b.download_raw_async(capi.F_XP1, hv.data_ptr())
    b.download_raw_async(capi.F_FP, hf.data_ptr())
    assert b.sync() == oob                        # out-of-box count of the block, as mdb_run returns it
    assert b.sync() == 0                          # nothing pending any more
    for host, ref in ((hx, xa), (hv, va), (hf, fa)):
        assert np.array_equal(capi.from_colmajor(host.numpy(), n, 3), ref)
    a.close(); b.close()


def test_transient_overflow_inside_mdb_run_falls_back_to_generic(oracle):
    """A list that fits at the first build and overflows mxKVOIS at a LATER rebuild of the same mdb_run block (atoms of a
    small sphere fly towards its centre): the capacity counters are read at every rebuild, the overflowing build is redone
    on the generic path at once, no step runs on an incomplete list, and the trajectory is the oracle's (which truncates
    such lists in the reference's scan order, as the generic kernel does)."""
    c = util.bcc_case((8, 8, 8), seed=77, mxkvois=118, temp=100.0, disp=0.01)
    ctr = np.asarray(c.boxlow) + 0.5 * np.asarray(c.zl)
    d = c.xp - ctr
    r = np.linalg.norm(d, axis=1)
    inside = (r < 1.9 * c.rr) & (r > 0.0)
    v0 = 0.045 * c.rr / 0.5e-15                    # 0.045 a0 per step towards the centre
    c.xp1 = c.xp1.copy()
    c.xp1[inside] = -v0 * d[inside] / r[inside, None]
    h = 0.5e-15
    ctx = util.make_ctx(c)
    assert ctx.get_option(capi.OPT_ACTIVE_PATH) == capi.FORCE_PATH_TILED     # fits at the start
    ctx.force(capi.FORCE)
    ctx.run(0, 25, 1, 10, h)                       # rebuilds at ITIME = 1, 11, 21
    assert ctx.get_option(capi.OPT_ACTIVE_PATH) == capi.FORCE_PATH_GENERIC   # a later build overflowed -> generic from there on
    assert ctx.nlist_overflow() > 0                # the reference's truncation counter of the generic kernel
    md = util.oracle_md(oracle, c)
    md.rebuild(); md.force()
    for it in range(25):
        md.step(it, 1, 10, h)
    g = md.get()
    assert np.max(np.abs(ctx.download(capi.F_XP) - g["xp"])) < 1e-11 * np.max(np.abs(g["xp"]))
    assert util.relerr(ctx.download(capi.F_FP), g["fp"]) < 1e-8
    kv, _ = ctx.nlist_copyout(capi.ORDER_ORIGINAL)
    kvo, _ = md.nlist()
    assert np.array_equal(kv[g["gid"] - 1], kvo) and kv.max() == 118
    ctx.close()


@pytest.mark.parametrize("path", ["generic", "tiled"])
def test_devptr_reference_shape_views(oracle, path):
    """mdb_devptr(XP | DEN | INDI): the reference-shaped device arrays (dm_WorkSpace%XP(NPRT,3), DEN(NPRT), INDI(NAPDEV,mxKVOIS))
    that untouched CUDA-Fortran analysis code reads; materialised from the packed internal layout on request."""
    c = util.bcc_case((7, 8, 9), seed=5)
    n = c.xp.shape[0]
    ctx = util.make_ctx(c, force_path={"generic": capi.FORCE_PATH_GENERIC, "tiled": capi.FORCE_PATH_TILED}[path])
    ctx.force(capi.FORCE)
    ctx.lib.mdb_devptr.restype = C.c_void_p
    xp = _dev_to_host(ctx.devptr(capi.F_XP), 3 * n, np.float64)
    den = _dev_to_host(ctx.devptr(capi.F_DEN), n, np.float64)
    assert np.array_equal(capi.from_colmajor(xp, n, 3), ctx.download(capi.F_XP, capi.ORDER_CELL))
    assert np.array_equal(den, ctx.download(capi.F_DEN, capi.ORDER_CELL))
    indi = _dev_to_host(ctx.devptr(capi.F_INDI), n * c.mxkvois, np.int32).reshape(c.mxkvois, n)
    kv, ind = ctx.nlist_copyout(capi.ORDER_CELL)
    ref = oracle.nlist_build_dev(c.nbox, c.napb, c.xp, c.ityp, c.statu, c.boxlow, c.zl, c.ifpd,
                                 np.ascontiguousarray(c.nb_rm.T).ravel(), c.mxkvois)
    assert np.array_equal(kv, ref["kvois"])
    for w in range(int(kv.max())):
        assert np.array_equal(indi[w][kv > w], ref["indi"][w][kv > w])
    ctx.close()


@pytest.mark.parametrize("lanes", [2, 4, 8])
@pytest.mark.parametrize("name", ["bcc", "neb_WH"])
def test_bank_ordered_lists_give_the_same_sums(oracle, name, lanes):
    """MDB_OPT_TILED_BANKORDER: the list builder deals every scanned class out by slot residue (conflict-free LDS.128 of the
    staged records per half-warp).  Order inside a class is free, so forces / densities / energies must stay the oracle's."""
    c = util.bcc_case((7, 9, 11), seed=7) if name == "bcc" else util.neb_case("react")
    ref = oracle.nlist_build_dev(c.nbox, c.napb, c.xp, c.ityp, c.statu, c.boxlow, c.zl, c.ifpd,
                                 np.ascontiguousarray(c.nb_rm.T).ravel(), c.mxkvois)
    gid = ref["gid"] - 1
    fp, den, _, ep = oracle.force(c.xp[gid], c.ityp[gid], ref["statu"][gid], ref["kvois"], ref["indi"], c.zl, c.ifpd,
                                  util.oracle_tables(oracle, c), epot=True)
    ctx = util.make_ctx(c, build=False, force_path=capi.FORCE_PATH_TILED)
    ctx.set_option(capi.OPT_TILED_LANES, lanes)
    ctx.set_option(capi.OPT_TILED_BANKORDER, 1)
    ctx.nlist_build()
    ctx.force(capi.FORCE | capi.EPOT)
    assert util.atom_relerr(ctx.download(capi.F_FP, capi.ORDER_CELL), fp) < TOL
    assert util.atom_relerr(ctx.download(capi.F_DEN, capi.ORDER_CELL), den) < TOL
    assert util.atom_relerr(ctx.download(capi.F_EPOT, capi.ORDER_CELL), ep) < TOL
    kv, _ = ctx.nlist_copyout(capi.ORDER_CELL)
    assert np.array_equal(kv, ref["kvois"])
    ctx.close()


def test_cpp_host_replays_the_fortran_shim_call_sequence(tmp_path):
    """tests/cpp/replay_shims.cpp: a compiled C++ host that makes the calls of fortran/mdb_shims.F90 in the order the unchanged
    MDPSCU shell makes them (device init, box, force-class slots, list, kernel-by-kernel For_One_Step, copy-out, clear).  Its
    results must be those of the same sequence through Python/ctypes: the boundary is language-neutral."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "replay_shims")
    libdir = os.path.join(root, "msmpscu_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(root, "tests", "cpp", "replay_shims.cpp"),
                           "-L" + libdir, "-lmdpscu_b200", "-Wl,-rpath," + libdir])
    c = util.bcc_case((9, 9, 9), seed=2025)
    n = c.xp.shape[0]
    nsteps = 23
    c.xp1 = c.xp1.copy()
    c.xp1[n // 2] = np.sqrt(2.0 * 800.0 * util.CP_EVERG / c.mass[0]) * np.array([1.0, 3.0, 5.0]) / np.sqrt(35.0)   # a fast atom: the
    # displacement-limited step of Predictor_DEV (IHDUP = -2 in the replay) has to shorten the step
    cfg, out = str(tmp_path / "cfg.bin"), str(tmp_path / "out.bin")
    with open(cfg, "wb") as f:
        np.array([n], np.int32).tofile(f)
        np.array(list(c.boxlow) + list(c.zl) + [c.mass[0], c.ru, float(c.nb_rm[0, 0] / c.ru)]).tofile(f)
        capi.colmajor(c.xp).tofile(f)
        capi.colmajor(c.xp1).tofile(f)
    r = subprocess.run([exe, cfg, out, str(nsteps)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    raw = np.fromfile(out)
    scal, vt, rest = raw[:4], raw[4:13].reshape(3, 3).T, raw[13:]
    xp, xp1, fp = (capi.from_colmajor(rest[k * 3 * n:(k + 1) * 3 * n], n, 3) for k in range(3))
    epot = rest[9 * n:10 * n]
    # the same sequence through the Python binding
    ctx = util.make_ctx(c)
    ctx.epc_set([1], [300.0], [1.0e-12], [0.1], [100.0 * util.CP_EVERG])
    ctx.force(capi.FORCE)
    h = 0.5e-15
    hs = []
    for it in range(nsteps):
        if (it - 1 + 1) % 2 == 0:
            h = ctx.timestep_limit(0.5e-15, 0.05e-8)
        hs.append(h)
        ctx.predict(h)
        if (it - 1) % 10 == 0:
            ctx.nlist_build()
        ctx.force(capi.FORCE)
        ctx.epc_apply()
        ctx.correct(h)
    assert min(hs) < 0.5e-15
    ctx.ekin()
    t = ctx.global_t()
    vt_py = ctx.force(capi.FORCE | capi.EPOT | capi.VIRIAL)
    assert np.array_equal(xp, ctx.download(capi.F_XP)) and np.array_equal(xp1, ctx.download(capi.F_XP1))
    assert np.array_equal(fp, ctx.download(capi.F_FP)) and np.array_equal(epot, ctx.download(capi.F_EPOT))
    assert scal[0] == t
    assert util.relerr(vt, vt_py) < 1e-13      # per-warp partial tensors: the summation order follows the chunk scheduling
    kv, _ = ctx.nlist_copyout(capi.ORDER_ORIGINAL)
    assert int(scal[1]) == int(kv.sum())
    assert ctx.embed_overruns() == 0          # a thermal crystal never leaves the embedding table
    ctx.close()


def test_non_identity_boxshape_on_the_generic_path(oracle):
    """A sheared / strained BOXSHAPE (pressure-coupled boxes): the list kernel (fp32, :1117-1119), the density pass, the
    virial kernel and the energy kernel apply it, the force-only kernel does not (MD_EAM_ForceTable_GPU.F90:515-518,775,
    1171-1174) -- the tiled path declines such boxes and AUTO runs the generic kernels; forcing TILED is refused."""
    c = util.bcc_case((7, 8, 9), seed=21)
    bs = np.array([[1.0, 0.02, 0.0], [0.0, 1.0, 0.01], [0.0, 0.0, 0.99]])
    ctx = capi.Context(0)
    ctx.box_set(c.nbox, c.napb, c.boxlow, c.zl, c.ifpd, c.mass, boxshape=bs)
    for f, a in ((capi.F_XP, c.xp), (capi.F_XP1, c.xp1), (capi.F_ITYP, c.ityp), (capi.F_STATU, c.statu)):
        ctx.upload(f, a)
    ctx.tables_set(util.product_tables(c), c.ru * c.ru)
    ctx.nlist_init(c.nb_rm, c.mxkvois)
    ctx.nlist_build()
    assert ctx.get_option(capi.OPT_ACTIVE_PATH) == capi.FORCE_PATH_GENERIC
    nbr = np.ascontiguousarray(c.nb_rm.T).ravel()
    ref = oracle.nlist_build_dev(c.nbox, c.napb, c.xp, c.ityp, c.statu, c.boxlow, c.zl, c.ifpd, nbr, c.mxkvois, boxshape=bs)
    ident = oracle.nlist_build_dev(c.nbox, c.napb, c.xp, c.ityp, c.statu, c.boxlow, c.zl, c.ifpd, nbr, c.mxkvois)
    assert not np.array_equal(ref["kvois"], ident["kvois"])          # the shape does change the lists of this case
    kv, ind = ctx.nlist_copyout(capi.ORDER_CELL)
    assert np.array_equal(kv, ref["kvois"])
    for w in range(int(kv.max())):
        assert np.array_equal(ind[w][kv > w], ref["indi"][w][kv > w])
    gid = ref["gid"] - 1
    T = util.oracle_tables(oracle, c)
    args = (c.xp[gid], c.ityp[gid], ref["statu"][gid], ref["kvois"], ref["indi"], c.zl, c.ifpd, T)
    fp, den, _, ep = oracle.force(*args, epot=True, boxshape=bs)
    fpv, _, vt, _ = oracle.force(*args, virial=True, boxshape=bs)
    ctx.force(capi.FORCE | capi.EPOT)
    assert util.atom_relerr(ctx.download(capi.F_FP, capi.ORDER_CELL), fp) < TOL
    assert util.atom_relerr(ctx.download(capi.F_DEN, capi.ORDER_CELL), den) < TOL
    assert util.atom_relerr(ctx.download(capi.F_EPOT, capi.ORDER_CELL), ep) < TOL
    v = ctx.force(capi.FORCE | capi.VIRIAL)
    assert util.relerr(v, vt) < TOL
    assert util.atom_relerr(ctx.download(capi.F_FP, capi.ORDER_CELL), fpv) < TOL   # CALPTENSOR's forces (r2 through the shape)
    ctx.run(0, 12, 1, 10, 0.5e-15)                                    # steps run on such a box
    ctx.set_option(capi.OPT_FORCE_PATH, capi.FORCE_PATH_TILED)
    with pytest.raises(capi.MDBError) as e:
        ctx.nlist_build()
    assert e.value.code == capi.ERR_UNSUPPORTED
    ctx.close()
