"""The GMD time loop with its schedules (mdb_run_sched / mdb_dd_run_sched): displacement-limited time step (scheme II of
Predictor_DEV, CommonGPU/MD_DiffScheme_GPU.F90:633-655), step-size ramp (scheme I) and list-period ramp
(Appshell/MD_Method_GenericMD_GPU.F90:351-361), with a primary knock-on atom in the box -- against the same loop composed
from the C oracle's procedures, and the per-tile displacement bounds of cascade runs against the global bound."""
import numpy as np
import pytest

import util
from msmpscu_b200 import capi

pytestmark = pytest.mark.gpu
EV = util.CP_EVERG
FS = 1.0e-15


def _pka_case(ncell=(8, 8, 8), ekev=2.0, seed=909):
    """bcc W at 300 K with one fast atom near the box centre, velocity along <135>"""
    c = util.bcc_case(ncell, seed=seed, temp=600.0)
    r = np.linalg.norm(c.xp - (c.boxlow + 0.5 * c.zl), axis=1)
    i = int(np.argmin(r))
    d = np.array([1.0, 3.0, 5.0]) / np.sqrt(35.0)
    c.xp1 = c.xp1.copy()
    c.xp1[i] = np.sqrt(2.0 * ekev * 1000.0 * EV / c.mass[0]) * d
    return c, i


def _sched(c, ihdup, hmi=0.5, hmx=0.5, dmx_lu=0.05, nb=(10, 10, 100)):
    return capi.Sched(ihdup, hmi * FS, hmx * FS, dmx_lu * c.rr, nb[0], nb[1], nb[2])


def _oracle_loop(O, c, epc, s, itime0, nsteps, it0, h):
    """GenericMD's loop body around For_One_Step, from the oracle's own procedures"""
    m = util.oracle_md(O, c)
    m.set_epc(*epc)
    m.rebuild()
    m.force()
    t, hist = 0.0, []
    for k in range(nsteps):
        itime = itime0 + k
        if s.ihdup > 0:
            h = min(s.hmx, s.hmi * ((itime - it0 + 1) // s.ihdup + 1))
        nb = min(s.nb_uptabmx, s.nb_uptabmi * ((itime - it0 + 1) // s.nb_dbitab + 1))
        if s.ihdup < 0 and (itime - it0 + 1) % (-s.ihdup) == 0:
            th = s.hmx
            while m.check_timestep(th, th * th * 0.5, s.dmx * s.dmx):
                th = th * 0.5
            h = th
        m.step(itime, it0, nb, h)
        t += h
        hist.append(h)
    return m, h, t, hist


@pytest.mark.parametrize("ihdup", [-1, -3, 4])
def test_run_sched_matches_the_oracle_loop(oracle, ihdup):
    c, ipka = _pka_case()
    ng = len(c.mass)
    epc = ([1] * ng, [300.0] * ng, [1.0e-12] * ng, [0.1] * ng, [100.0 * EV] * ng)
    # scheme I ramps 0.1 fs -> 0.5 fs; both ramp the list period 2 -> 6
    s = _sched(c, ihdup, hmi=0.1 if ihdup > 0 else 0.5, nb=(2, 6, 5))
    nsteps, it0, h0 = 24, 1, 0.5 * FS
    m, h_ref, t_ref, hist = _oracle_loop(oracle, c, epc, s, 0, nsteps, it0, h0)
    ctx = util.make_ctx(c)
    ctx.epc_set(*epc)
    ctx.force(capi.FORCE)
    rc, h_gpu, t_gpu = ctx.run_sched(0, nsteps, it0, s, h0)
    assert rc == 0
    if ihdup < 0:
        assert min(hist) < 0.5 * FS, "the PKA must force a shorter step for this to test anything: %r" % hist
    assert h_gpu == h_ref and abs(t_gpu - t_ref) <= 1e-15 * t_ref, (h_gpu, h_ref, t_gpu, t_ref, hist)
    ref = m.get()
    for name, f, tol in (("xp", capi.F_XP, 1e-11), ("xp1", capi.F_XP1, 1e-9), ("fp", capi.F_FP, 1e-9)):
        err = util.relerr(ctx.download(f), ref[name])
        assert err < tol, "%s %g" % (name, err)
    ctx.close()


def test_per_tile_displacement_bounds_change_nothing_but_the_work():
    """the class shortcut is exact, so per-tile bounds (on), the global bound (off) and the full list (classes off) must give
    bit-identical trajectories; with the PKA in the box the global bound trips, the per-tile one only around the PKA"""
    c, ipka = _pka_case((48, 48, 48), ekev=3.0)      # 221 184 atoms: large enough for the pass times to show the work
    ng = len(c.mass)
    epc = ([1] * ng, [300.0] * ng, [1.0e-12] * ng, [0.1] * ng, [100.0 * EV] * ng)
    s = _sched(c, -1)
    out = []
    for guard, classes in ((1, 1), (0, 1), (-1, 1), (0, 0)):
        ctx = util.make_ctx(c, build=False, force_path=capi.FORCE_PATH_TILED)
        ctx.set_option(capi.OPT_TILE_GUARD, guard)
        ctx.set_option(capi.OPT_TILED_CLASSES, classes)
        ctx.nlist_build()
        ctx.epc_set(*epc)
        ctx.force(capi.FORCE)
        ctx.prof_enable(True)
        rc, h, t = ctx.run_sched(0, 40, 1, s, 0.5 * FS)
        prof = ctx.prof_get()
        out.append((ctx.download(capi.F_XP), ctx.download(capi.F_XP1), ctx.download(capi.F_FP), h, t, prof))
        ctx.close()
    for o in out[1:]:
        assert np.array_equal(o[0], out[0][0]) and np.array_equal(o[1], out[0][1]) and np.array_equal(o[2], out[0][2])
        assert o[3] == out[0][3] and o[4] == out[0][4]
    # the work: per-tile bounds must keep the passes clearly below the full-list cost the global bound falls back to
    ms = lambda p: p["pass1"][1] + p["pass2"][1]
    print("pass ms: per-tile %.3f  global %.3f  auto %.3f  full list %.3f" % tuple(ms(o[5]) for o in out))
    assert ms(out[0][5]) < 0.8 * ms(out[1][5])
    assert abs(ms(out[2][5]) - ms(out[0][5])) < 0.25 * ms(out[0][5])     # auto = on in a displacement-limited run


def _stop_tables(ng):
    ne = 200
    etab = np.linspace(1.0 * EV, 2.0e4 * EV, ne)
    nk = ng * ng
    stab = np.stack([(1.0 + 0.3 * k) * 1.0e-27 * np.sqrt(etab / EV) for k in range(nk)], axis=1)
    return etab, stab, np.arange(1, nk + 1).reshape(ng, ng)


def _make_tiled(c, epc, stop):
    ctx = util.make_ctx(c, build=False, force_path=capi.FORCE_PATH_TILED)
    ctx.set_option(capi.OPT_TILED_BANKORDER, 1)     # (auto would leave it off for a box this small: the big runs have it on)
    ctx.epc_set(*epc)
    if stop:
        ctx.stopping_set(*stop)
    return ctx


@pytest.mark.parametrize("world", [2, 3])
def test_decomposed_cascade_run_matches_single_context(world):
    """PKA + electronic stopping + displacement-limited step on z-slabs (in-process backend): owned slices bit-identical to
    the single-context mdb_run_sched, same step sizes"""
    c, ipka = _pka_case((8, 8, 20), ekev=1.0, seed=404)
    ng = len(c.mass)
    epc = ([1] * ng, [300.0] * ng, [1.0e-12] * ng, [0.1] * ng, [100.0 * EV] * ng)
    stop = _stop_tables(ng) + ([1] * ng, [6.3e22] * ng)
    s = _sched(c, -1)
    nsteps = 33
    full = _make_tiled(c, epc, stop)
    full.nlist_build(); full.force(capi.FORCE)
    rc, h_f, t_f = full.run_sched(0, nsteps, 1, s, 0.5 * FS)
    ctxs = [_make_tiled(c, epc, stop) for _ in range(world)]
    capi.dd_local_attach(ctxs)
    ctxs[0].dd_build()
    ctxs[0].dd_force(capi.FORCE)
    rc, h_d, t_d = ctxs[0].dd_run_sched(0, nsteps, 1, s, 0.5 * FS)
    assert h_d == h_f and t_d == t_f and t_f < nsteps * 0.5 * FS
    gid_f = full.download(capi.F_GID, capi.ORDER_CELL)
    owned = 0
    for r, ctx in enumerate(ctxs):
        info = ctx.dd_info()
        a0, a1 = info["a0"], info["a1"]
        owned += a1 - a0
        assert np.array_equal(ctx.download(capi.F_GID, capi.ORDER_CELL)[a0:a1], gid_f[a0:a1]), "rank %d order" % r
        for name, f in (("xp", capi.F_XP), ("xp1", capi.F_XP1), ("fp", capi.F_FP), ("den", capi.F_DEN), ("dis", capi.F_DIS)):
            x, y = full.download(f, capi.ORDER_CELL)[a0:a1], ctx.download(f, capi.ORDER_CELL)[a0:a1]
            assert np.array_equal(x, y), "rank %d field %s differs (max %g)" % (r, name, np.max(np.abs(x - y)))
    assert owned == c.xp.shape[0]
    for ctx in ctxs:
        ctx.close()
    full.close()
