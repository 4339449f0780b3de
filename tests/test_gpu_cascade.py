"""GPU tests of the cascade-physics procedures around the path: active region by cells, electronic
stopping (global-density model), PKA insertion -- against the NumPy restatements in oracle/cascade_np.py, and their
effect on the hot path (inactive atoms, skipped cells) against the C oracle."""
import numpy as np
import pytest

import util
from msmpscu_b200 import capi
from oracle import cascade_np as CN

pytestmark = pytest.mark.gpu
EV = util.CP_EVERG


def _hot_case():
    """bcc W with two fast atoms (500 eV and 80 eV): the seeds of the kinetic-energy criterion"""
    c = util.bcc_case((16, 15, 17), seed=61, temp=300.0)
    return c


@pytest.mark.parametrize("extend", [0, 1, 2])
def test_active_region_by_kinetic_energy_matches_restatement(oracle, extend):
    c = _hot_case()
    n = c.xp.shape[0]
    ctx = util.make_ctx(c)
    m = c.mass[0]
    ctx.pka_insert(137, 500.0 * EV, [1.0, 3.0, 5.0])            # <135> PKA
    ctx.pka_insert(n - 5, 80.0 * EV, [0.0, -1.0, 0.0])
    v = ctx.download(capi.F_XP1)
    assert abs(0.5 * m * np.sum(v[136] ** 2) - 500.0 * EV) < 1e-12 * 500.0 * EV
    assert np.allclose(v[136] / np.linalg.norm(v[136]), np.array([1.0, 3.0, 5.0]) / np.sqrt(35.0), rtol=0, atol=1e-15)
    nact = ctx.active_region(ekin_erg=50.0 * EV, extend=extend)
    st = ctx.download(capi.F_STATU, capi.ORDER_CELL)
    inc = ctx.download(capi.F_IC, capi.ORDER_CELL)
    ncell = ctx.cellinfo()[0]
    want = CN.activate_region_by_cells(np.ones(n, np.int32), ctx.download(capi.F_ITYP, capi.ORDER_CELL),
                                       ctx.download(capi.F_XP1, capi.ORDER_CELL), inc, c.mass, ncell, 1, c.ifpd,
                                       ekin_erg=50.0 * EV, extend=extend)
    assert np.array_equal(st, want)
    assert nact == int(np.count_nonzero(want & 1)) and 0 < nact < n
    # effect on the path: rebuild (NAAC: cells without active atoms are skipped), forces of inactive atoms vanish, and
    # everything equals the oracle run on the same STATU
    c2 = util.bcc_case((16, 15, 17), seed=61, temp=300.0)
    c2.xp1 = ctx.download(capi.F_XP1)
    c2.statu = ctx.download(capi.F_STATU)
    for path in (capi.FORCE_PATH_GENERIC, capi.FORCE_PATH_TILED):
        ctx.set_option(capi.OPT_FORCE_PATH, path)
        ctx.nlist_build()
        ctx.force(capi.FORCE | capi.EPOT)
        ref = oracle.nlist_build_dev(c2.nbox, c2.napb, c2.xp, c2.ityp, c2.statu, c2.boxlow, c2.zl, c2.ifpd,
                                     np.ascontiguousarray(c2.nb_rm.T).ravel(), c2.mxkvois)
        gid = ref["gid"] - 1
        fp, den, _, ep = oracle.force(c2.xp[gid], c2.ityp[gid], ref["statu"][gid], ref["kvois"], ref["indi"], c2.zl, c2.ifpd,
                                      util.oracle_tables(oracle, c2), epot=True)
        f = ctx.download(capi.F_FP, capi.ORDER_CELL)
        assert np.array_equal(ctx.download(capi.F_GID, capi.ORDER_CELL), ref["gid"])
        assert util.relerr(f, fp) < 1e-10 and util.relerr(ctx.download(capi.F_EPOT, capi.ORDER_CELL), ep) < 1e-10
        inactive = (ref["statu"][gid] & 1) == 0
        assert np.all(f[inactive] == 0.0) and np.any(f[~inactive] != 0.0)
        if path == capi.FORCE_PATH_GENERIC:
            kv, _ = ctx.nlist_copyout(capi.ORDER_CELL)
            assert np.array_equal(kv, ref["kvois"])
    # inactive atoms do not move
    x0 = ctx.download(capi.F_XP)
    ctx.run(0, 5, 1, 10, 0.5e-15)
    x1 = ctx.download(capi.F_XP)
    idle = (c2.statu & 1) == 0
    assert np.array_equal(x0[idle], x1[idle]) and not np.array_equal(x0[~idle], x1[~idle])
    ctx.active_all(True)
    assert np.all(ctx.download(capi.F_STATU) & 1)
    ctx.close()


def test_active_region_by_type_and_keep(oracle):
    """W + one H (NEB_Test box): the H atom is the centre particle; KEEP accumulates regions"""
    c = util.neb_case("react")
    n = c.xp.shape[0]
    ctx = util.make_ctx(c)
    ityp_c = ctx.download(capi.F_ITYP, capi.ORDER_CELL)
    inc = ctx.download(capi.F_IC, capi.ORDER_CELL)
    ncell = ctx.cellinfo()[0]
    nact = ctx.active_region(centpart=[0, 1], extend=1)
    st = ctx.download(capi.F_STATU, capi.ORDER_CELL)
    want = CN.activate_region_by_cells(np.ones(n, np.int32), ityp_c, ctx.download(capi.F_XP1, capi.ORDER_CELL), inc, c.mass, ncell, 1,
                                       c.ifpd, centpart=[0, 1], extend=1)
    assert np.array_equal(st, want) and nact == int(np.count_nonzero(want & 1)) and 0 < nact < n
    ctx.pka_insert(1, 200.0 * EV, [1.0, 0.0, 0.0])
    nact2 = ctx.active_region(ekin_erg=100.0 * EV, extend=0, keep=True)
    want2 = CN.activate_region_by_cells(want, ityp_c, ctx.download(capi.F_XP1, capi.ORDER_CELL), inc, c.mass, ncell, 1, c.ifpd,
                                        ekin_erg=100.0 * EV, extend=0, keep=True)
    assert np.array_equal(ctx.download(capi.F_STATU, capi.ORDER_CELL), want2) and nact2 >= nact
    ctx.close()


@pytest.mark.parametrize("path", [capi.FORCE_PATH_GENERIC, capi.FORCE_PATH_TILED])
def test_active_region_through_the_neighbour_list(path):
    """CP_BYNB_AR (ActiveByNeigbors0/1): seeds by kinetic energy and by type grown through the list, with KEEP"""
    c = util.neb_case("react")
    n = c.xp.shape[0]
    ctx = util.make_ctx(c, build=False, force_path=path)
    ctx.nlist_build()
    ctx.pka_insert(137, 500.0 * EV, [1.0, 3.0, 5.0])
    cell = lambda f: ctx.download(f, capi.ORDER_CELL)
    kv, indi = ctx.nlist_copyout(capi.ORDER_CELL)
    for extend in (0, 1, 2):
        nact = ctx.active_region(ekin_erg=50.0 * EV, extend=extend, by_neighbours=True)
        want = CN.activate_region_by_neighbours(np.ones(n, np.int32), cell(capi.F_ITYP), cell(capi.F_XP1), c.mass, kv, indi.T,
                                                ekin_erg=50.0 * EV, extend=extend)
        assert np.array_equal(cell(capi.F_STATU), want) and nact == int(np.count_nonzero(want & 1))
        assert (nact == 1) if extend == 0 else (1 < nact < n)
    # by type (the H atom), kept on top of the region of the fast atom
    st0 = cell(capi.F_STATU)
    cent = [0] * len(c.mass); cent[-1] = 1
    nact2 = ctx.active_region(centpart=cent, extend=1, keep=True, by_neighbours=True)
    want = CN.activate_region_by_neighbours(st0, cell(capi.F_ITYP), cell(capi.F_XP1), c.mass, kv, indi.T, centpart=cent, extend=1, keep=True)
    assert np.array_equal(cell(capi.F_STATU), want) and nact2 == int(np.count_nonzero(want & 1)) and nact2 > nact
    ctx.close()


def _stop_tables(ng):
    """a smooth synthetic E-S table (erg, erg cm^2): the reference reads such tables from its stopping libraries"""
    ne = 200
    etab = np.linspace(1.0 * EV, 2.0e4 * EV, ne)
    nk = ng * ng
    stab = np.stack([(1.0 + 0.3 * k) * 1.0e-27 * np.sqrt(etab / EV) for k in range(nk)], axis=1)
    kpair = np.arange(1, nk + 1).reshape(ng, ng)
    return etab, stab, kpair


def test_electronic_stopping_matches_restatement_and_rides_in_mdb_run():
    c = util.neb_case("react")
    n = c.xp.shape[0]
    ng = len(c.mass)
    etab, stab, kpair = _stop_tables(ng)
    enable, mden = [1] * ng, [6.3e22] * ng
    ctx = util.make_ctx(c)
    ctx.pka_insert(17, 5000.0 * EV, [1.0, 3.0, 5.0])
    ctx.pka_insert(n, 300.0 * EV, [0.0, 0.0, -1.0])             # the H atom
    ctx.force(capi.FORCE)
    f0, v = ctx.download(capi.F_FP), ctx.download(capi.F_XP1)
    ctx.stopping_set(etab, stab, kpair, enable, mden)
    ctx.stopping_apply()
    f1 = ctx.download(capi.F_FP)
    want = CN.stopping_force_gden(f0, v, c.ityp, c.statu, c.mass, etab, stab, kpair, enable, mden)
    assert np.any(f1 != f0) and np.max(np.abs(f1 - want)) <= 1e-14 * np.max(np.abs(want))
    # inside mdb_run the stopping acts between the EPC friction and the corrector: equal to the kernel-by-kernel sequence
    h = 0.5e-15
    epc = ([1] * ng, [300.0] * ng, [1.0e-12] * ng, [0.1] * ng, [100.0 * EV] * ng)
    a, b = util.make_ctx(c), util.make_ctx(c)
    for x in (a, b):
        x.pka_insert(17, 5000.0 * EV, [1.0, 3.0, 5.0])
        x.epc_set(*epc)
        x.stopping_set(etab, stab, kpair, enable, mden)
        x.force(capi.FORCE)
    a.run(0, 12, 1, 10, h)
    for it in range(12):
        b.predict(h)
        if (it - 1) % 10 == 0:
            b.nlist_build()
        b.force(capi.FORCE)
        b.epc_apply()
        b.stopping_apply()
        b.correct(h)
    for f in (capi.F_XP, capi.F_XP1, capi.F_FP):
        assert np.array_equal(a.download(f), b.download(f))
    # and it takes energy out: the same run without stopping keeps a faster PKA
    d = util.make_ctx(c)
    d.pka_insert(17, 5000.0 * EV, [1.0, 3.0, 5.0]); d.epc_set(*epc); d.force(capi.FORCE)
    d.run(0, 12, 1, 10, h)
    ke = lambda x: float(np.sum(x.download(capi.F_XP1)[16] ** 2))
    assert ke(a) < ke(d)
    for x in (ctx, a, b, d):
        x.close()


def test_local_density_stopping_and_energy_loss_match_restatement():
    """ST_MOD_LDEN_KERNEL + the ELOSS bookkeeping: W + H box (two types: per-type neighbour counts from INDI) and a one-type box
    (counts = KVOIS), on both force paths; the loss accumulates over the steps of mdb_run with the step's own H"""
    for case, pkas in ((util.neb_case("react"), ((17, 5000.0), (None, 300.0))), (util.bcc_case((9, 8, 8), seed=5, temp=300.0), ((40, 3000.0),))):
        c = case
        n = c.xp.shape[0]
        ng = len(c.mass)
        etab, stab, kpair = _stop_tables(ng)
        enable = [1] * ng
        for path in (capi.FORCE_PATH_GENERIC, capi.FORCE_PATH_TILED):
            ctx = util.make_ctx(c, build=False, force_path=path)
            ctx.nlist_build()
            for orig, ev in pkas:
                ctx.pka_insert(n if orig is None else orig, ev * EV, [1.0, 3.0, 5.0])
            ctx.force(capi.FORCE)
            ctx.stopping_set(etab, stab, kpair, enable, [0.0] * ng)
            ctx.stopping_options(local_density=True, save_eloss=True)
            f0 = ctx.download(capi.F_FP, capi.ORDER_CELL)
            v = ctx.download(capi.F_XP1, capi.ORDER_CELL)
            ityp_c, st_c = ctx.download(capi.F_ITYP, capi.ORDER_CELL), ctx.download(capi.F_STATU, capi.ORDER_CELL)
            kv, indi = ctx.nlist_copyout(capi.ORDER_CELL)
            h = 0.25e-15
            ctx.stopping_apply(h)
            f1 = ctx.download(capi.F_FP, capi.ORDER_CELL)
            want, loss = CN.stopping_force_lden(f0, v, ityp_c, st_c, c.mass, etab, stab, kpair, enable, kv, indi.T, c.nb_rm, dt=h)
            assert np.any(f1 != f0) and np.max(np.abs(f1 - want)) <= 1e-14 * np.max(np.abs(want))
            gid = ctx.download(capi.F_GID, capi.ORDER_CELL)
            el = ctx.stopping_eloss()
            assert np.count_nonzero(loss) == len(pkas) and np.max(np.abs(el[gid - 1] - loss)) <= 1e-14 * loss.max()
            # inside mdb_run: the fused end-of-step kernel == the kernel-by-kernel sequence, losses included
            a, b = util.make_ctx(c, build=False, force_path=path), util.make_ctx(c, build=False, force_path=path)
            epc = ([1] * ng, [300.0] * ng, [1.0e-12] * ng, [0.1] * ng, [100.0 * EV] * ng)
            for x in (a, b):
                x.nlist_build()
                for orig, ev in pkas:
                    x.pka_insert(n if orig is None else orig, ev * EV, [1.0, 3.0, 5.0])
                x.epc_set(*epc)
                x.stopping_set(etab, stab, kpair, enable, [0.0] * ng)
                x.stopping_options(local_density=True, save_eloss=True)
                x.force(capi.FORCE)
            a.run(0, 12, 1, 10, h)
            for it in range(12):
                b.predict(h)
                if (it - 1) % 10 == 0:
                    b.nlist_build()
                b.force(capi.FORCE)
                b.epc_apply()
                b.stopping_apply(h)
                b.correct(h)
            for fld in (capi.F_XP, capi.F_XP1, capi.F_FP):
                assert np.array_equal(a.download(fld), b.download(fld))
            ea, eb = a.stopping_eloss(), b.stopping_eloss(reset=True)
            assert np.array_equal(ea, eb) and ea.max() > 0.0 and np.count_nonzero(b.stopping_eloss()) == 0
            for x in (ctx, a, b):
                x.close()


def test_parrep_event_detection_on_the_device():
    """Do_ChangeDetect (Appshell/MD_Method_ParRep_GPU.F90:1094-1167) from its device pieces: save the replicas, quench, compare with
    SimBoxIni (Do_Compare :1241-1297, host mirror msmpscu_b200.mdlib.Do_Compare as the checker), restore replicas and list."""
    from types import SimpleNamespace
    from msmpscu_b200 import mdlib
    nrep = 4
    c = util.parrep_case(nrep)
    napb = c.napb
    xini = util.neb_case("react").xp                      # SimBoxIni: the configuration the replicas were started from
    ctx = util.make_ctx(c)
    ctx.force(capi.FORCE)
    ctx.thermalize(600.0, 777, 0)
    ctx.run(0, 40, 1, 10, 0.5e-15)
    before = {f: ctx.download(f) for f in (capi.F_XP, capi.F_XP1, capi.F_DIS, capi.F_STATU)}
    ctx.nlist_build()                                      # (the restore rebuilds the list from the restored positions, as :1158-1159)
    kv0 = ctx.download(capi.F_KVOIS, capi.ORDER_ORIGINAL)
    ctx.state_save()
    fl, mm, de = ctx.steepest(300, 0.1, 0.1 * c.rr, 1.0e-5 * c.rr, 1.0e-5 * EV)
    xq = ctx.download(capi.F_XP)
    drtol = 0.03 * c.rr                                    # STRCUT_DRTol default, MD_TypeDef_SimCtrlParam.F90:202
    ini = SimpleNamespace(NPRT=napb, RR=c.rr, ZL=np.asarray(c.zl), XP=xini)
    ctrl = SimpleNamespace(STRCUT_DRTol=0.03, IFPD=c.ifpd)
    boxes = [SimpleNamespace(XP=xq[b * napb:(b + 1) * napb]) for b in range(nrep)]
    want = mdlib.Do_Compare(ini, boxes, ctrl)
    fb, ibt, ncb, fa = ctx.compare(xini, drtol, per_atom=True, nbox=nrep)
    assert np.array_equal(fa, want)
    assert (ibt, ncb) == mdlib.Transition_Replicas(want, napb)
    # a forced event in ONE replica: the reference state with two atoms displaced by 0.1 a0 is what replica 3 is compared to ...
    x2 = xq.copy()
    x2[2 * napb + 7, 0] += 0.1 * c.rr
    x2[2 * napb + 100, 2] -= 0.05 * c.rr
    ctx.upload(capi.F_XP, x2)
    fb2, ibt2, ncb2, fa2 = ctx.compare(xini, drtol, per_atom=True, nbox=nrep)
    boxes = [SimpleNamespace(XP=x2[b * napb:(b + 1) * napb]) for b in range(nrep)]
    want2 = mdlib.Do_Compare(ini, boxes, ctrl)
    assert np.array_equal(fa2, want2) and fa2[2 * napb + 7] == 1 and fa2[2 * napb + 100] == 1
    assert fb2[2] == 1 and ibt2 >= 3 and ncb2 >= 1
    # ... and the mask takes them out again (m_pCfgCompMask)
    mask = np.ones(napb, np.int32)
    mask[[7, 100]] = 0
    fb3, _, _, fa3 = ctx.compare(xini, drtol, mask=mask, per_atom=True, nbox=nrep)
    assert fa3[2 * napb + 7] == 0 and fa3[2 * napb + 100] == 0
    assert np.array_equal(fa3, mdlib.Do_Compare(ini, boxes, ctrl, MASK=mask))
    # restore: replicas and list exactly as before the detection
    ctx.state_restore()
    for f, a in before.items():
        assert np.array_equal(ctx.download(f), a)
    assert np.array_equal(ctx.download(capi.F_KVOIS, capi.ORDER_ORIGINAL), kv0)
    ctx.run(40, 10, 1, 10, 0.5e-15)                       # and the run goes on
    ctx.close()
