"""GPU parity tests: every check goes through the C ABI (msmpscu_b200.capi) and compares with the
CPU oracle on the same seeded inputs.  Bars (BASELINE.json north_star): cell assignments and
neighbour sets bit-exact; forces, energies, virial within 1e-10 relative in fp64."""
import os

import numpy as np
import pytest

import util
from msmpscu_b200 import capi

pytestmark = pytest.mark.gpu

FORCE_RTOL = 1e-10
PATHS = {"generic": capi.FORCE_PATH_GENERIC, "tiled": capi.FORCE_PATH_TILED}


def _oracle_list(O, c):
    return O.nlist_build_dev(c.nbox, c.napb, c.xp, c.ityp, c.statu, c.boxlow, c.zl, c.ifpd,
                             np.ascontiguousarray(c.nb_rm.T).ravel(), c.mxkvois)


CASES = {
    "bcc8": lambda: util.bcc_case((8, 8, 8)),
    "bcc_7x9x11": lambda: util.bcc_case((7, 9, 11), seed=7),
    "multibox3": lambda: util.bcc_case((7, 7, 7), nbox=3, seed=99),
    "neb_WH": lambda: util.neb_case("react"),
    "bcc_fs_ackland": lambda: util.bcc_fs_case((7, 8, 9)),   # FS_TYPE kernels (MD_FS_ForceTable_GPU.F90), a real FS potential
    "parrep_replicas": lambda: util.parrep_case(3),          # configs[3] shape: W+H replicas, 1.6 x RU lists, MAXNB 400
    "fcc_cu_setfl": lambda: util.fcc_cu_case((6, 7, 8)),   # imported NIST setfl tables (EAM_NIST library), 134-entry lists
}


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("name", list(CASES))
def test_cells_and_list_bit_exact(oracle, name, path):
    """NeighboresListTest.F90:92-121 restated: identical KVOIS and identical ORDER of INDI, plus the
    cell assignment / sort (GID, NAC, IA1th) the list is built on."""
    c = CASES[name]()
    ref = _oracle_list(oracle, c)
    ctx = util.make_ctx(c, force_path=PATHS[path])
    ncell, nc, mxnac = ctx.cellinfo()
    assert ncell == ref["ncell"]
    assert np.array_equal(ctx.download(capi.F_GID, capi.ORDER_CELL), ref["gid"])
    assert np.array_equal(ctx.download(capi.F_NAC), ref["nac"])
    assert np.array_equal(ctx.download(capi.F_NAAC), ref["naac"])
    assert np.array_equal(ctx.download(capi.F_IA1TH), ref["ia1th"])
    assert np.array_equal(ctx.download(capi.F_IC, capi.ORDER_ORIGINAL), ref["inc"])
    assert mxnac == ref["nac"].max()
    kv, ind = ctx.nlist_copyout(capi.ORDER_CELL)
    assert np.array_equal(kv, ref["kvois"])
    for w in range(c.mxkvois):
        m = kv > w
        assert np.array_equal(ind[w][m], ref["indi"][w][m]), "neighbour order differs in row %d" % w
    assert ctx.nlist_overflow() == 0
    ctx.close()


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("name", list(CASES))
def test_force_energy_virial_parity(oracle, name, path):
    """CalForceTest.F90:113-147 restated for EAM: forces, DEN, EPOT, virial against the CPU oracle."""
    c = CASES[name]()
    ref = _oracle_list(oracle, c)
    gid = ref["gid"] - 1
    T = util.oracle_tables(oracle, c)
    fp, den, vt, ep = oracle.force(c.xp[gid], c.ityp[gid], ref["statu"][gid], ref["kvois"], ref["indi"], c.zl, c.ifpd,
                                   T, virial=True, epot=True)
    ctx = util.make_ctx(c, force_path=PATHS[path])
    vt_gpu = ctx.force(capi.FORCE | capi.VIRIAL | capi.EPOT)
    f_gpu = ctx.download(capi.F_FP, capi.ORDER_CELL)
    d_gpu = ctx.download(capi.F_DEN, capi.ORDER_CELL)
    e_gpu = ctx.download(capi.F_EPOT, capi.ORDER_CELL)
    assert util.relerr(f_gpu, fp) < FORCE_RTOL
    assert util.relerr(d_gpu, den) < FORCE_RTOL
    assert util.relerr(e_gpu, ep) < FORCE_RTOL
    assert util.relerr(vt_gpu, vt / c.nbox) < FORCE_RTOL
    # force-only entry point gives the same forces as the virial one
    ctx.force(capi.FORCE)
    f_gpu = ctx.download(capi.F_FP, capi.ORDER_CELL)
    assert util.relerr(f_gpu, fp) < FORCE_RTOL
    assert util.relerr(ctx.download(capi.F_DEN, capi.ORDER_CELL), den) < FORCE_RTOL
    # ORIGINAL-order download un-permutes through GID
    f_orig = ctx.download(capi.F_FP, capi.ORDER_ORIGINAL)
    assert np.array_equal(f_orig[gid], f_gpu)
    ctx.close()


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("tag", ["react", "product"])
def test_golden_known_answer_through_cuda(tag, path):
    """The reference GPU build's own output (examples/NEB_Test/GMD/*P0000_0001.0000): force [eV/LU]
    and POT [eV] to 9 digits.  That binary used Rmax = max(NB_RM)."""
    c = util.neb_case(tag, rmax_mode="NB_RM")
    ctx = util.make_ctx(c, force_path=PATHS[path])
    ctx.force(capi.FORCE | capi.EPOT)
    F = ctx.download(capi.F_FP) * c.rr / util.CP_EVERG
    POT = -ctx.download(capi.F_EPOT) / util.CP_EVERG
    assert np.max(np.abs(F - c.gold_force)) < 2e-9      # printed with 9 significant digits, |F| <= 0.39
    assert np.max(np.abs(POT / c.gold_pot - 1.0)) < 2e-9
    ctx.close()


def test_integrator_epc_ekin_bit_exact(oracle):
    """Predictor / EPC / corrector / EKIN are element-wise: un-fused arithmetic must match bit for bit."""
    c = util.bcc_case((8, 8, 8), seed=5)
    ctx = util.make_ctx(c)
    ctx.force(capi.FORCE)
    h = 0.5e-15
    fp0 = ctx.download(capi.F_FP)
    x0, v0 = ctx.download(capi.F_XP), ctx.download(capi.F_XP1)
    st0 = ctx.download(capi.F_STATU)
    assert np.array_equal(x0, c.xp) and np.array_equal(v0, c.xp1)
    ctx.predict(h)
    x1, v1, d1, st1 = oracle.predictor(x0, v0, fp0, np.zeros_like(x0), st0, c.ityp, c.mass, h, c.boxlow, c.zl, c.ifpd)
    assert np.array_equal(ctx.download(capi.F_XP), x1)
    assert np.array_equal(ctx.download(capi.F_XP1), v1)
    assert np.array_equal(ctx.download(capi.F_DIS), d1)
    assert np.array_equal(ctx.download(capi.F_STATU), st1)
    te, al, cut, he = [300.0], [1.0e-12], [0.1], [100.0 * util.CP_EVERG]
    ctx.epc_set([1], te, al, cut, he)
    ctx.epc_apply()
    f2 = oracle.epc(v1, fp0, st1, c.ityp, [1], c.mass, te, al, cut, he)
    assert np.array_equal(ctx.download(capi.F_FP), f2)
    assert not np.array_equal(f2, fp0)
    ctx.correct(h)
    v2 = oracle.corrector(v1, f2, st1, c.ityp, c.mass, h)
    assert np.array_equal(ctx.download(capi.F_XP1), v2)
    ctx.ekin()
    assert np.array_equal(ctx.download(capi.F_EKIN), oracle.ekin(v2, st1, c.ityp, c.mass))
    ctx.close()


@pytest.mark.parametrize("path", ["generic", "auto"])
def test_truncated_list_matches_reference_truncation(oracle, path):
    """mxKVOIS smaller than the true count: the reference silently keeps the first mxKVOIS in scan order.
    The tiled builder keeps no scan order, so it reports such a build and AUTO falls back to the generic
    (reference-ordered) kernels; forcing the tiled path is refused with a status code."""
    c = util.bcc_case((7, 7, 7), mxkvois=60)
    ref = _oracle_list(oracle, c)
    ctx = util.make_ctx(c, force_path=capi.FORCE_PATH_GENERIC if path == "generic" else capi.FORCE_PATH_AUTO)
    assert ctx.get_option(capi.OPT_ACTIVE_PATH) == capi.FORCE_PATH_GENERIC
    kv, ind = ctx.nlist_copyout(capi.ORDER_CELL)
    assert kv.max() == 60 and np.array_equal(kv, ref["kvois"])
    assert np.array_equal(ind, ref["indi"])
    assert ctx.nlist_overflow() == c.xp.shape[0]  # every bcc atom has > 60 neighbours inside 2.28 a0
    ctx.close()
    if path == "auto":
        with pytest.raises(capi.MDBError) as e:
            util.make_ctx(c, force_path=capi.FORCE_PATH_TILED)
        assert e.value.code == capi.ERR_UNSUPPORTED


@pytest.mark.parametrize("path", list(PATHS))
def test_md_trajectory_tracks_oracle(oracle, path):
    """For_One_Step sequence (predictor -> rebuild when MOD(ITIME-IT0,NB_UPTAB)==0 -> force -> corrector):
    25 steps with rebuilds at ITIME=1,11,21; positions/velocities stay within round-off growth of the
    oracle's, the lists built on them stay identical, and the NVE Hamiltonian drift matches."""
    c = util.bcc_case((8, 8, 8), seed=11, temp=600.0)
    h, it0, nup = 0.5e-15, 1, 10
    md = util.oracle_md(oracle, c)
    md.rebuild(); md.force()
    ctx = util.make_ctx(c, force_path=PATHS[path])
    ctx.force(capi.FORCE)
    for it in range(25):
        md.step(it, it0, nup, h)
    ctx.run(0, 25, it0, nup, h)
    ref = md.get()
    x, v = ctx.download(capi.F_XP), ctx.download(capi.F_XP1)
    assert util.relerr(x, ref["xp"]) < 1e-12
    assert util.relerr(v, ref["xp1"]) < 1e-9
    assert np.array_equal(ctx.download(capi.F_GID, capi.ORDER_CELL), ref["gid"])
    kv, ind = ctx.nlist_copyout(capi.ORDER_CELL)
    okv, oind = md.nlist()
    assert np.array_equal(kv, okv)
    assert all(np.array_equal(ind[w][kv > w], oind[w][kv > w]) for w in range(c.mxkvois))
    # Hamiltonian per atom (Common/MD_TypeDef_SimBox.F90:5155-5163)
    md.epot(); ctx.force(capi.EPOT); ctx.ekin()
    ref = md.get()
    ham_ref = (ref["epot"].sum() + ref["ekin"].sum()) / c.xp.shape[0]
    ham = (ctx.download(capi.F_EPOT).sum() + ctx.download(capi.F_EKIN).sum()) / c.xp.shape[0]
    assert abs(ham - ham_ref) < 1e-10 * abs(ham_ref)
    ctx.close()


@pytest.mark.parametrize("path", list(PATHS))
def test_out_of_box_atoms_are_parked_like_the_reference(oracle, path):
    """Non-periodic x: atoms pushed outside are flagged OUTOFBOX, counted, and placed at the end of the
    CELL order (smallest original id last), MD_NeighborsList_GPU.F90:1505-1526,1627-1637."""
    c = util.bcc_case((7, 7, 7), ifpd=(0, 1, 1), seed=3)
    c.xp = c.xp.copy()
    c.xp[[5, 77, 300], 0] += 2.0 * c.zl[0]
    c.xp[[10], 0] -= 2.0 * c.zl[0]
    ref = _oracle_list(oracle, c)
    ctx = util.make_ctx(c, build=False, force_path=PATHS[path])
    assert ctx.nlist_build() == 4 == ref["nout"]
    assert np.array_equal(ctx.download(capi.F_GID, capi.ORDER_CELL), ref["gid"])
    assert np.array_equal(ctx.download(capi.F_STATU, capi.ORDER_ORIGINAL), ref["statu"])
    kv, ind = ctx.nlist_copyout(capi.ORDER_CELL)
    n_in = c.xp.shape[0] - 4
    assert np.array_equal(kv[:n_in], ref["kvois"][:n_in])
    # forces with missing neighbours / parked atoms still match
    gid = ref["gid"] - 1
    fp, den, _, _ = oracle.force(c.xp[gid], c.ityp[gid], ref["statu"][gid], ref["kvois"], ref["indi"], c.zl, c.ifpd,
                                 util.oracle_tables(oracle, c))
    ctx.force(capi.FORCE)
    assert util.relerr(ctx.download(capi.F_FP, capi.ORDER_CELL), fp) < FORCE_RTOL
    assert np.all(ctx.download(capi.F_FP, capi.ORDER_CELL)[n_in:] == 0.0)
    ctx.close()


def test_tiled_matches_generic_at_scale():
    """128 000 atoms (too slow for the CPU oracle in a unit test): the tiled fast path against the
    bit-faithful generic kernels on the device, lists identical, forces to 1e-12."""
    c = util.bcc_case((40, 40, 40), seed=31)
    a = util.make_ctx(c, force_path=capi.FORCE_PATH_GENERIC)
    b = util.make_ctx(c, force_path=capi.FORCE_PATH_TILED)
    ka, ia = a.nlist_copyout(capi.ORDER_CELL)
    kb, ib = b.nlist_copyout(capi.ORDER_CELL)
    assert np.array_equal(ka, kb) and all(np.array_equal(ia[w][ka > w], ib[w][ka > w]) for w in range(c.mxkvois))
    a.force(capi.FORCE); b.force(capi.FORCE)
    assert util.relerr(b.download(capi.F_DEN, capi.ORDER_CELL), a.download(capi.F_DEN, capi.ORDER_CELL)) < 1e-12
    assert util.relerr(b.download(capi.F_FP, capi.ORDER_CELL), a.download(capi.F_FP, capi.ORDER_CELL)) < 1e-12
    a.run(0, 12, 1, 10, 0.5e-15); b.run(0, 12, 1, 10, 0.5e-15)
    assert np.array_equal(a.download(capi.F_GID, capi.ORDER_CELL), b.download(capi.F_GID, capi.ORDER_CELL))
    assert util.relerr(b.download(capi.F_XP), a.download(capi.F_XP)) < 1e-13
    a.close(); b.close()


def test_distance_classes_are_exact(oracle):
    """The tiled passes scan only the build-time distance classes they need while every atom has moved
    less than half the class margin since the rebuild.  (i) with and without the classes the results
    agree to round-off; (ii) a large move between rebuilds (here: the host overwrites positions) must
    switch the passes to the full list, and the forces must still match the oracle on that list."""
    c = util.bcc_case((8, 8, 8), seed=21)
    a = util.make_ctx(c, force_path=capi.FORCE_PATH_TILED)
    b = util.make_ctx(c, force_path=capi.FORCE_PATH_TILED)
    b.set_option(capi.OPT_TILED_CLASSES, 0)
    assert a.get_option(capi.OPT_ACTIVE_PATH) == capi.FORCE_PATH_TILED
    a.run(0, 8, 1, 10, 0.5e-15); b.run(0, 8, 1, 10, 0.5e-15)
    assert util.relerr(a.download(capi.F_FP), b.download(capi.F_FP)) < 1e-13
    assert util.relerr(a.download(capi.F_XP), b.download(capi.F_XP)) < 1e-14
    # (ii) stale list + moved atoms: displace every atom by up to 0.45 A without rebuilding
    ref = _oracle_list(oracle, c)
    gid = ref["gid"] - 1
    rng = np.random.default_rng(5)
    xnew = c.xp + rng.uniform(-0.45e-8, 0.45e-8, size=c.xp.shape)
    ctx = util.make_ctx(c, force_path=capi.FORCE_PATH_TILED)
    ctx.upload(capi.F_XP, xnew)           # no rebuild: the reference would use the stale list as is
    ctx.force(capi.FORCE)
    fp, den, _, _ = oracle.force(xnew[gid], c.ityp[gid], ref["statu"][gid], ref["kvois"], ref["indi"], c.zl, c.ifpd,
                                 util.oracle_tables(oracle, c))
    assert util.relerr(ctx.download(capi.F_FP, capi.ORDER_CELL), fp) < FORCE_RTOL
    assert util.relerr(ctx.download(capi.F_DEN, capi.ORDER_CELL), den) < FORCE_RTOL
    for x in (a, b, ctx):
        x.close()


def test_errors_are_status_codes_not_stops():
    ctx = capi.Context(0)
    with pytest.raises(capi.MDBError) as e:
        ctx.nlist_build()
    assert e.value.code == capi.ERR_STATE
    with pytest.raises(capi.MDBError):
        ctx.force(capi.FORCE)
    ctx.close()


def test_gmd_example_inputs_through_the_mdlib_interface(oracle):
    """examples/GMD_Test (BASELINE configs[0] family): box, control and configuration files as shipped
    (2000 W + 1 He, Bonny EAM1, 300 K control file), driven through the reference's own call sequence
    (Initialize_Globle_Variables_DEV -> Register_ForceClass -> Init_Forcetable_Dev ->
    Initialize_NeighboreList_DEV -> Cal_NeighBoreList_DEV -> For_One_Step ...) and compared with the oracle."""
    import os
    from msmpscu_b200 import inputs, mdlib
    g = util.GOLD
    box = inputs.read_box_file(os.path.join(g, "gmd_W_2000_He1_EAM1_box.dat"))
    ctl = inputs.read_ctrl_file(os.path.join(g, "gmd_CtrlFile300K.dat"), box)
    inputs.read_config(os.path.join(g, "gmd_W_2000_Tetra.cfg"), box)
    rng = np.random.default_rng(300)
    box.XP1 = rng.normal(0.0, 1.0, size=box.XP.shape) * np.sqrt(1.38054e-16 * 300.0 / box.CM[box.ITYP - 1])[:, None]
    dev = mdlib.DeviceState(0)
    fc = mdlib.Register_ForceClass(box.PotType)
    mdlib.Initialize_Globle_Variables_DEV(dev, box, ctl)
    ft = mdlib.Init_Forcetable_Dev(dev, box, ctl, fc)
    mdlib.Initialize_NeighboreList_DEV(dev, box, ctl)
    assert mdlib.Cal_NeighBoreList_DEV(dev, box, ctl) == 0
    mdlib.CalForce_ForceClass(dev, box, ctl, fc)
    # oracle with the same tables / list rule
    T = oracle.Tables(oracle.LIB_BONNY_EAM1, box.PTYPE, ctl.NUMFTABR, ctl.NUMFTABE, float(ctl.RU.max()))
    md = oracle.MD(1, box.NPRT, box.XP, box.XP1, box.ITYP, box.STATU, box.CM, box.BOXLOW, box.ZL, ctl.IFPD,
                   np.ascontiguousarray(ctl.NB_RM.T).ravel(), ctl.NB_MXNBS, T)
    md.rebuild(); md.force()
    for name in ("potr", "fpotb", "dfembd"):
        assert np.array_equal(getattr(ft, name), getattr(T, name))
    for it in range(12):
        mdlib.For_One_Step(dev, it, box, ctl, fc)
        md.step(it, ctl.IT0, ctl.NB_UPTAB, ctl.H)
    mdlib.CalEpot_ForceClass(dev, box, ctl, fc)
    mdlib.CalEKin_DEV(dev, box, ctl)
    mdlib.CopyOut_SimBox_DEV(dev)
    md.epot()
    ref = md.get()
    assert util.relerr(box.XP, ref["xp"]) < 1e-12
    assert util.relerr(box.FP, ref["fp"]) < FORCE_RTOL
    assert util.relerr(box.EPOT, ref["epot"]) < FORCE_RTOL
    th = mdlib.Cal_thermal_quantities(box)
    ham_ref = (ref["epot"].sum() + ref["ekin"].sum()) / box.NPRT
    assert abs(th["HARMIL"] - ham_ref) < 1e-10 * abs(ham_ref)
    vt = mdlib.CalPTensor_ForceClass(dev, box, ctl, fc)
    md.force(virial=True)
    assert util.relerr(vt, md.get()["vtensor"]) < FORCE_RTOL
    dev.ctx.close()


@pytest.mark.parametrize("path", ["generic", "tiled", "tiled_fused"])
def test_nvt_epc_steps_track_oracle(oracle, path):
    """BASELINE configs[1] dynamics in small: NVT through the electron-phonon thermostat (EPC friction applied
    to the fresh force, then the corrector -- fused into the force-pass epilogue on the tiled path)."""
    c = util.bcc_case((8, 8, 8), seed=77, temp=900.0)
    te, al, cut, he = [300.0], [1.0e-12], [0.1], [100.0 * util.CP_EVERG]
    md = util.oracle_md(oracle, c)
    md.set_epc([1], te, al, cut, he)
    md.rebuild(); md.force()
    ctx = util.make_ctx(c, force_path=PATHS[path.split("_")[0]])
    ctx.set_option(capi.OPT_FUSE_EPILOGUE, 1 if path.endswith("fused") else 0)
    ctx.epc_set([1], te, al, cut, he)
    ctx.force(capi.FORCE)
    for it in range(15):
        md.step(it, 1, 10, 0.5e-15)
    ctx.run(0, 15, 1, 10, 0.5e-15)
    ref = md.get()
    assert util.relerr(ctx.download(capi.F_XP), ref["xp"]) < 1e-12
    assert util.relerr(ctx.download(capi.F_XP1), ref["xp1"]) < 1e-9
    assert util.relerr(ctx.download(capi.F_FP), ref["fp"]) < FORCE_RTOL   # FP holds the force after the EPC friction
    ctx.close()


@pytest.mark.parametrize("path", list(PATHS))
def test_steepest_quench_tracks_oracle(oracle, path):
    """Do_Steepest0_Forsteps_DEV (CommonGPU/MD_SteepestScheme_GPU.F90:20-153) with the scalars and the
    stop flag kept on the device, against the CPU restatement: same stopping iteration, same configuration.
    (i) default criteria (stops on the energy criterion); (ii) tight criteria, out of steps (IFLAG = 0)."""
    c = util.bcc_case((8, 8, 8), seed=5, temp=0.0, disp=0.04)
    rr = c.rr
    for mx, mistep, midele in ((60, 1.0e-5, 1.0e-3), (12, 1.0e-9, 1.0e-12)):
        md = util.oracle_md(oracle, c)
        md.rebuild()
        fl_o, mm_o, de_o = md.steepest0(mx, 0.1, 0.1 * rr, mistep * rr, midele * util.CP_EVERG)
        ctx = util.make_ctx(c, force_path=PATHS[path])
        ctx.force(capi.FORCE)
        f0 = np.abs(ctx.download(capi.F_FP)).max()
        fl, mm, de = ctx.steepest(mx, 0.1, 0.1 * rr, mistep * rr, midele * util.CP_EVERG)
        assert fl == fl_o
        ref = md.get()
        assert util.relerr(ctx.download(capi.F_XP), ref["xp"]) < 1e-10
        assert abs(mm - mm_o) <= 1e-8 * abs(mm_o) and abs(de - de_o) <= 1e-6 * abs(de_o) + 1e-30
        ctx.force(capi.FORCE)
        assert np.abs(ctx.download(capi.F_FP)).max() < 0.5 * f0   # it did relax
        ctx.close()


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("scheme", ["cg0", "cg1", "sd1"])
def test_cg_and_linesearch_quench_track_oracle(oracle, path, scheme):
    """Do_CG0/CG1_Forsteps_DEV (CommonGPU/MD_CGScheme_GPU.F90:16-276) and Do_Steepest1_Forsteps_DEV
    (MD_SteepestScheme_GPU.F90:157-260) with device-resident scalars against the CPU restatements: same exit iteration
    (one run that converges, one that runs out of steps) and the same final configuration."""
    c = util.bcc_case((8, 8, 8), seed=5, temp=0.0, disp=0.04)
    rr = c.rr
    for mx, mistep, midele in ((80, 1.0e-5, 1.0e-3), (9, 1.0e-9, 1.0e-12)):
        md = util.oracle_md(oracle, c)
        md.rebuild()
        ctx = util.make_ctx(c, force_path=PATHS[path])
        ctx.force(capi.FORCE)
        f0 = np.abs(ctx.download(capi.F_FP)).max()
        if scheme == "sd1":
            fl_o, de_o = md.steepest1(mx, 0.1 * rr, mistep * rr)
            fl, _, de = ctx.steepest(mx, 0.1, 0.1 * rr, mistep * rr, midele * util.CP_EVERG, meth=capi.QUENCH_LSEARCH)
        else:
            ls = 1 if scheme == "cg1" else 0
            fl_o, de_o = md.cg(mx, ls, 0.1 * rr, mistep * rr, midele * util.CP_EVERG)
            fl, de = ctx.cg(mx, 0.1 * rr, mistep * rr, midele * util.CP_EVERG, meth=capi.QUENCH_LSEARCH if ls else 0)
        assert fl == fl_o, (scheme, mx, fl, fl_o)
        assert util.relerr(ctx.download(capi.F_XP), md.get()["xp"]) < 1e-10
        # per-atom energies are ~1e-11 erg: differences below 1e-22 erg are last-bit noise (CG0's clamped secant step can
        # return an atom exactly to where it started, see the oracle's header comment)
        assert abs(de - de_o) <= 1e-6 * abs(de_o) + 1e-22
        ctx.force(capi.FORCE)
        assert np.abs(ctx.download(capi.F_FP)).max() < 0.5 * f0   # it did relax
        ctx.close()


def test_steepest_quench_of_the_reference_example():
    """examples/NEB_Test (2000 W + 1 H, Bonny EAM1): the reference's own run quenches this configuration for 1000
    steps; its printed cohesive energy stays -8.89488 eV/atom (GMD/thermP0000_0001) while max|F| drops from 0.39 to
    0.08 eV/LU (ReactP0000_0001.0001).  The 2019 binary damped dynamically, so this pins the level the quench must
    reach, not its trajectory: energy must not rise, C.E. must print the same six digits, forces must drop at least
    as far."""
    c = util.neb_case("react")
    g = np.load(os.path.join(util.GOLD, "neb_gmd_react_quenched.npz"))
    ctx = util.make_ctx(c)
    ctx.force(capi.FORCE | capi.EPOT)
    e0 = ctx.download(capi.F_EPOT).sum()
    fl, mm, de = ctx.steepest(1000, 0.1, 0.1 * c.rr, 1.0e-5 * c.rr, 1.0e-3 * util.CP_EVERG)
    assert fl != 0
    ctx.force(capi.FORCE | capi.EPOT)
    e1 = ctx.download(capi.F_EPOT).sum()
    f1 = np.abs(ctx.download(capi.F_FP)).max() * c.rr / util.CP_EVERG       # eV/LU
    assert e1 <= e0
    assert "%.5E" % (e1 / util.CP_EVERG / 2001) == "-8.89488E+00"
    ce, ce_ref = e1 / util.CP_EVERG / 2001, -g["pot"].mean()
    assert ce <= ce_ref + 5e-7 and abs(ce - ce_ref) < 1e-5   # at least as deep as the reference's damped state
    assert f1 <= np.abs(g["force"]).max() * 1.05
    ctx.close()


def test_temperature_scaling_and_timestep_check_match_oracle(oracle):
    """Cal_GlobalT_DEV, VelScaling_DEV (per box, 3 boxes) and CheckTimestep_DEV (MD_DiffScheme_GPU.F90:1042-1446)
    against their restatements, with a fixed atom, a velocity-fixed component and an inactive atom in the mix."""
    c = util.bcc_case((6, 6, 6), seed=9, temp=500.0, nbox=3)
    c.statu = c.statu.copy()
    c.statu[3] |= 2 | 4 | 8          # FIXPOS xyz: not counted in T, velocity zeroed by the scaling
    c.statu[10] |= 32                # FIXVELY
    c.statu[500] = 0                 # inactive
    md = util.oracle_md(oracle, c)
    md.rebuild(); md.force()
    ctx = util.make_ctx(c)
    ctx.force(capi.FORCE)
    t0, t0_ref = ctx.global_t(), md.global_t()
    assert abs(t0 - t0_ref) < 1e-12 * t0_ref and 400.0 < t0 < 600.0
    h = 0.5e-15
    d2 = np.sum((h * c.xp1) ** 2, axis=1).max()
    for mxd2 in (0.25 * d2, 4.0 * d2):
        assert ctx.check_timestep(h, 0.5 * h * h, mxd2) == md.check_timestep(h, 0.5 * h * h, mxd2)
    assert ctx.check_timestep(h, 0.5 * h * h, 0.25 * d2) == 1 and ctx.check_timestep(h, 0.5 * h * h, 4.0 * d2) == 0
    ctx.vel_scaling(300.0)
    assert md.vel_scaling(300.0) == 0
    assert util.relerr(ctx.download(capi.F_XP1), md.get()["xp1"]) < 1e-13
    t1 = ctx.global_t()
    assert abs(t1 - md.global_t()) < 1e-12 * t1 and abs(t1 - 300.0) < 1.0   # (the FIXVEL component was zeroed after the factor was set)
    v = ctx.download(capi.F_XP1)
    assert np.all(v[3] == 0.0) and v[10, 1] == 0.0 and v[10, 0] != 0.0
    ctx.close()


def test_thermalizing_mc_matches_oracle_and_maxwell(oracle):
    """Thermalizing_MC_DEV (MD_DiffScheme_GPU.F90:1608-1805) with Philox4x32-10 uniforms keyed by the
    ORIGINAL atom id.  (i) velocities equal the CPU restatement's (same integers; log/cos differ by ulps),
    (ii) fixed / inactive atoms are treated as in the reference, (iii) per-box momentum is zero, (iv) the result does
    not depend on the sort order (thermalise before and after a list build), (v) Maxwell statistics at TI."""
    c = util.bcc_case((10, 10, 10), seed=3, temp=0.0, nbox=2)
    c.statu = c.statu.copy()
    c.statu[5] |= 16            # FIXVELX
    c.statu[7] |= 2 | 4 | 8     # FIXPOS xyz
    c.statu[11] = 0             # inactive
    ti, seed = 450.0, 0x1234ABCD5678
    md = util.oracle_md(oracle, c)
    md.rebuild()
    md.thermalize(ti, seed, 3)
    ref = md.get()["xp1"]
    ctx = util.make_ctx(c)                       # list built: cell order
    ctx.thermalize(ti, seed, 3)
    v = ctx.download(capi.F_XP1)
    assert util.relerr(v, ref) < 1e-12
    ctx2 = util.make_ctx(c, build=False)         # no sort yet: ORIGINAL order on the device
    ctx2.thermalize(ti, seed, 3)
    assert util.relerr(ctx2.download(capi.F_XP1), v) < 1e-15
    ctx2.thermalize(ti, seed, 4)                 # another draw: different numbers
    assert util.relerr(ctx2.download(capi.F_XP1), v) > 0.1
    ctx2.close()
    napb = c.napb
    for b in range(2):
        vb = v[b * napb:(b + 1) * napb]
        assert np.abs(vb.mean(axis=0)).max() < 1e-12 * np.abs(vb).max()      # one species: plain mean
    free = np.ones(len(v), bool); free[[5, 7, 11]] = False
    # not redrawn (inactive -> 0; fixed -> kept, 0 here): these hold exactly minus the removed box velocity
    assert np.array_equal(v[11], v[7]) and v[5, 0] == v[11, 0] and v[5, 1] != v[11, 1]
    assert 0 < np.abs(v[11]).max() < 0.1 * np.abs(v).max()
    m = c.mass[0]
    t_kin = m * (v[free] ** 2).sum() / (3.0 * free.sum() * 1.38054e-16)
    assert abs(t_kin - ti) < 0.03 * ti                                        # 4000 atoms: sigma ~ 1 %
    s = v[free] / np.sqrt(1.38054e-16 * ti / m)
    assert abs(s.std() - 1.0) < 0.03 and abs((s ** 4).mean() - 3.0) < 0.3     # Gaussian components
    ctx.close()


@pytest.mark.parametrize("name", ["bcc_7x9x11", "neb_WH", "fcc_cu_setfl"])
def test_atomic_stress_parity(oracle, name):
    """pCalAVStress (CAL_EAM_AtomicStress_KERNEL, MD_EAM_ForceTable_GPU.F90:1775-1925) against the oracle, on both list
    paths; and the consistency the two reference kernels imply: sum_i AP_i / 2 = VTENSOR * nbox (CALPTENSOR_KERNEL halves
    each directed pair, :1222, and COPYOUT_VIRIALTENSOR divides by the number of boxes, :1462-1463)."""
    c = CASES[name]()
    md = util.oracle_md(oracle, c)
    md.rebuild()
    ref = md.avstress()
    for path in PATHS.values():
        ctx = util.make_ctx(c, force_path=path)
        ap = ctx.atomic_stress(capi.ORDER_ORIGINAL)
        assert ap.shape == ref.shape
        assert util.relerr(ap, ref) < FORCE_RTOL
        vt = ctx.force(capi.FORCE | capi.VIRIAL)
        assert np.allclose(0.5 * ap.sum(axis=0).reshape(3, 3), np.asarray(vt) * c.nbox, rtol=1e-9,
                           atol=1e-12 * np.abs(ap).sum())
        ctx.close()


@pytest.mark.parametrize("path", list(PATHS))
@pytest.mark.parametrize("nearest", [14, 8, 300])
def test_reorder_nearest_bit_exact(oracle, path, nearest):
    """Reorder_NeighBoreList_Nearest_Dev (MD_NeighborsList_GPU.F90:1805-2066): the NEAREST closest neighbours in order
    of increasing distance, in place, identical to the restatement of the reference's insertion sort (same entries, same
    order, ties included); afterwards the force procedures follow the truncated list, as in the reference."""
    c = util.bcc_case((7, 8, 9), seed=21)
    md = util.oracle_md(oracle, c)
    md.rebuild()
    md.reorder_nearest(nearest)
    kv_o, ind_o = md.nlist()
    ctx = util.make_ctx(c, force_path=PATHS[path])
    ctx.nlist_reorder_nearest(nearest)
    kv, ind = ctx.nlist_copyout(capi.ORDER_CELL)
    assert np.array_equal(kv, kv_o) and kv.max() == min(nearest, 112)
    for w in range(kv.max()):
        m = kv > w
        assert np.array_equal(ind[w][m], ind_o[w][m]), "row %d" % w
    md.force()
    ctx.force(capi.FORCE)
    assert util.relerr(ctx.download(capi.F_FP, capi.ORDER_CELL), md.get_sorted_fp()) < FORCE_RTOL
    ctx.nlist_build()                                  # a rebuild restores the full list
    assert ctx.nlist_copyout(capi.ORDER_CELL)[0].max() == 112
    with pytest.raises(capi.MDBError):
        ctx.nlist_reorder_nearest(513)
    ctx.close()


@pytest.mark.parametrize("path", list(PATHS))
def test_dynamic_damping_tracks_oracle(oracle, path):
    """DAMPING_KERNEL (MD_DiffScheme_GPU.F90:125-185) is bit-exact; Do_DynDamp_Forsteps_DEV (:1809-1860) with the stop flag on
    the device exits at the same iteration as the restatement with the same configuration and velocities."""
    c = util.bcc_case((8, 8, 8), seed=15, temp=300.0, disp=0.04)
    c.statu = c.statu.copy(); c.statu[3] |= 2 | 8            # FIXPOSX, FIXPOSZ
    md = util.oracle_md(oracle, c)
    md.rebuild(); md.force()
    ctx = util.make_ctx(c, force_path=PATHS[path])
    ctx.force(capi.FORCE)
    md.damping(); ctx.damping()
    v = ctx.download(capi.F_XP1)
    assert np.array_equal(v, md.get()["xp1"]) and (v == 0.0).mean() > 0.3 and v[3, 0] == 0.0 and v[3, 2] == 0.0
    for mx, midele in ((60, 2.0e-4), (7, 1.0e-12)):
        fl_o, de_o = md.dyndamp(mx, 1.0e-15, midele * util.CP_EVERG)
        fl, de = ctx.dyndamp(mx, 1.0e-15, midele * util.CP_EVERG)
        assert fl == fl_o, (mx, fl, fl_o)
        ref = md.get()
        assert util.relerr(ctx.download(capi.F_XP), ref["xp"]) < 1e-12
        assert util.relerr(ctx.download(capi.F_XP1), ref["xp1"]) < 1e-9
        assert abs(de - de_o) <= 1e-6 * abs(de_o) + 1e-22
    ctx.close()


def test_c1_nve_energy_drift_matches_oracle(oracle):
    """BASELINE configs[0] shape: bcc W, 16^3 cells = 8192 atoms, Marinica EAM2, NVE, h = 0.5 fs, list
    rebuilt every 10 steps, HARMIL recorded every 10 steps for 1000 steps -- the product path (mdb_run, tiled kernels) against
    the CPU restatement of the reference's step loop: the Hamiltonian series agree to 1e-9 relative at every record, and so
    does the drift they show."""
    c = util.bcc_case((16, 16, 16), seed=12345, temp=600.0, disp=0.02)
    n = c.xp.shape[0]
    assert n == 8192
    h, it0, nup, nrec = 0.5e-15, 1, 10, 100
    try:
        oracle.lib().orc_set_threads(min(16, os.cpu_count() or 1))
    except Exception:
        pass
    md = util.oracle_md(oracle, c)
    md.rebuild(); md.force()
    ctx = util.make_ctx(c)
    ctx.force(capi.FORCE)

    def ham_ref():
        md.epot()
        r = md.get()
        return (r["epot"].sum() + r["ekin"].sum()) / n

    def ham_gpu():
        ctx.force(capi.EPOT); ctx.ekin()
        return (ctx.download(capi.F_EPOT).sum() + ctx.download(capi.F_EKIN).sum()) / n

    hr, hg = [ham_ref()], [ham_gpu()]
    it = 0
    for _ in range(nrec):
        for _k in range(nup):
            md.step(it + _k, it0, nup, h)
        ctx.run(it, nup, it0, nup, h)
        it += nup
        hr.append(ham_ref()); hg.append(ham_gpu())
    hr, hg = np.array(hr), np.array(hg)
    assert np.max(np.abs(hg - hr)) < 1e-9 * abs(hr[0])
    drift_r, drift_g = (hr - hr[0]) / abs(hr[0]), (hg - hg[0]) / abs(hg[0])
    assert np.max(np.abs(drift_r)) < 1e-4                      # velocity Verlet at 0.5 fs conserves H to ~1e-6
    assert np.max(np.abs(drift_g - drift_r)) < 1e-9
    print("C1 NVE: max |dH/H| oracle %.3e, product %.3e, max difference %.3e" % (
        np.abs(drift_r).max(), np.abs(drift_g).max(), np.abs(drift_g - drift_r).max()))
    ctx.close()


def _vacuum_slab_case():
    """bcc W slab with a vacuum gap: the crystal fills the lower 60 % of a box that is open (non-periodic) along z --
    empty cells, surface atoms with short lists, cells whose atoms are all inactive (NAAC = 0, skipped by the list kernel,
    MD_NeighborsList_GPU.F90:981-982)."""
    c = util.bcc_case((8, 8, 8), seed=31, ifpd=(1, 1, 0))
    zl = c.zl.copy(); zl[2] = c.zl[2] / 0.6
    c.zl, c.boxlow = zl, np.array([c.boxlow[0], c.boxlow[1], c.xp[:, 2].min() - 0.3 * c.rr])
    c.statu = c.statu.copy()
    corner = (c.xp[:, 0] < c.boxlow[0] + 2.66 * c.rr) & (c.xp[:, 1] < c.boxlow[1] + 2.66 * c.rr) & (c.xp[:, 2] < c.boxlow[2] + 2.66 * c.rr)
    c.statu[corner] = 0                                   # one whole corner cell inactive
    return c


EDGE_CASES = {
    "vacuum_slab": _vacuum_slab_case,
    "three_cells_per_edge": lambda: util.bcc_case((7, 7, 7), seed=5, ru_lu=1.9, nb_fac=1.2),   # 7 a0 / 2.28 -> 3 cells (the minimum)
    "one_atom_removed_and_one_added": lambda: _defect_case(),
}


def _defect_case():
    """a vacancy and a self-interstitial: ragged list lengths (the interstitial's neighbours exceed the crystal count)"""
    c = util.bcc_case((8, 8, 8), seed=8)
    x = c.xp.copy()
    x[100] = x[101] + np.array([0.5, 0.0, 0.0]) * c.rr * 0.9      # atom 100 leaves its site and sits next to atom 101
    c.xp = x
    return c


@pytest.mark.parametrize("name", list(EDGE_CASES))
def test_edge_configurations(oracle, name):
    """Empty cells, inactive cells, the minimum cell grid, defects: cells, lists (bit-exact, reference order) and forces on
    the generic path and on AUTO (tiled where the configuration fits its shared-memory plan, generic otherwise)."""
    c = EDGE_CASES[name]()
    ref = _oracle_list(oracle, c)
    gid = ref["gid"] - 1
    fp, den, _, ep = oracle.force(c.xp[gid], c.ityp[gid], ref["statu"][gid], ref["kvois"], ref["indi"], c.zl, c.ifpd,
                                  util.oracle_tables(oracle, c), epot=True)
    assert ref["kvois"].min() < ref["kvois"].max()
    for path in (capi.FORCE_PATH_GENERIC, capi.FORCE_PATH_AUTO):
        ctx = util.make_ctx(c, force_path=path)
        assert np.array_equal(ctx.download(capi.F_GID, capi.ORDER_CELL), ref["gid"])
        assert np.array_equal(ctx.download(capi.F_NAC), ref["nac"]) and np.array_equal(ctx.download(capi.F_NAAC), ref["naac"])
        kv, ind = ctx.nlist_copyout(capi.ORDER_CELL)
        assert np.array_equal(kv, ref["kvois"])
        for w in range(kv.max()):
            m = kv > w
            assert np.array_equal(ind[w][m], ref["indi"][w][m])
        ctx.force(capi.FORCE | capi.EPOT)
        assert util.relerr(ctx.download(capi.F_FP, capi.ORDER_CELL), fp) < FORCE_RTOL
        assert util.relerr(ctx.download(capi.F_EPOT, capi.ORDER_CELL), ep) < FORCE_RTOL
        ctx.close()


@pytest.mark.parametrize("path", list(PATHS))
def test_lbfgs_quench_tracks_oracle(oracle, path):
    """DO_LBFGSB_FORSTEPS_DEV (MD_LBFGSScheme_GPU.F90:177-388 = SETULB without bounds) with the vectors on the device and
    the direction formed in coefficient space, against the oracle's restatement (direct two-loop recursion): the same
    number of force evaluations and accepted steps and the same final configuration -- (i) run to max|F| <= 1e-6 of the
    start, (ii) cut short (IFLAG = 1), (iii) with a fixed atom, which stays where it was.  (The horizon is kept to a few
    dozen evaluations: over hundreds, round-off between the two ways of forming the direction eventually flips a
    line-search decision and the iterates part ways while reaching the same minimum.)"""
    base = util.bcc_case((8, 8, 8), seed=5, temp=300.0, disp=0.06)
    md0 = util.oracle_md(oracle, base); md0.rebuild(); md0.force()
    f0 = np.abs(md0.get()["fp"]).max()
    for mx, msave, tol, fix in ((400, 5, 1.0e-6, False), (23, 3, 1.0e-6, False), (400, 4, 1.0e-3, True)):
        c = util.bcc_case((8, 8, 8), seed=5, temp=300.0, disp=0.06)
        if fix:
            c.statu = c.statu.copy(); c.statu[7] |= 2 | 4 | 8
        md = util.oracle_md(oracle, c)
        md.rebuild()
        fl_o, nfg_o, nit_o = md.lbfgsb(mx, msave, 0.0, tol * f0)
        ctx = util.make_ctx(c, force_path=PATHS[path])
        fl, nfg, nit = ctx.lbfgs(mx, msave, 0.0, tol * f0)
        assert (fl, nfg, nit) == (fl_o, nfg_o, nit_o), (mx, (fl, nfg, nit), (fl_o, nfg_o, nit_o))
        # the quasi-Newton iteration amplifies the round-off between the two formulations: 1e-16 after 20 evaluations,
        # 1e-13 after 44, 2e-10 after 60 (measured); the bar here is 1e-8 on positions with identical evaluation counts
        assert util.relerr(ctx.download(capi.F_XP), md.get()["xp"]) < 1e-8
        assert np.all(ctx.download(capi.F_XP1) == 0.0)
        ctx.force(capi.FORCE)
        fp = ctx.download(capi.F_FP)
        if fix:
            assert np.array_equal(ctx.download(capi.F_XP)[7], c.xp[7])
            fp = np.delete(fp, 7, axis=0)
        if fl == 0:
            assert np.abs(fp).max() <= tol * f0 and nit >= 10
        else:
            assert fl == 1 and np.abs(fp).max() < 0.5 * f0
        ctx.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_disordered_configurations_tiled_vs_oracle(oracle, seed):
    """Strongly disordered boxes (random displacements up to 0.18 a0, a few vacancies, random box shape): the scan counts of
    the distance classes differ from atom to atom inside a warp, which is what the warp-uniform padding skip of the pass
    kernels and the class partition of the list builder must handle.  Forced TILED path against the oracle (forces, DEN,
    energies) and against the generic path on the device, at the build positions and after the atoms have moved enough to
    leave the class shortcut (displacement guard)."""
    rng = np.random.default_rng(1000 + seed)
    ncell = tuple(int(v) for v in rng.integers(7, 11, size=3))
    c = util.bcc_case(ncell, seed=50 + seed, disp=0.18, temp=2000.0)
    keep = np.ones(c.xp.shape[0], bool)
    keep[rng.choice(c.xp.shape[0], size=9, replace=False)] = False
    c.statu = c.statu.copy()
    c.statu[~keep] = 0                                   # inactive atoms: skipped as i, still listed as j (reference behaviour)
    ref = _oracle_list(oracle, c)
    gid = ref["gid"] - 1
    tabs = util.oracle_tables(oracle, c)
    fp, den, _, ep = oracle.force(c.xp[gid], c.ityp[gid], ref["statu"][gid], ref["kvois"], ref["indi"], c.zl, c.ifpd, tabs, epot=True)
    assert ref["kvois"].max() - ref["kvois"][ref["kvois"] > 0].min() >= 8
    ctx = util.make_ctx(c, force_path=capi.FORCE_PATH_TILED)
    assert ctx.get_option(capi.OPT_ACTIVE_PATH) == capi.FORCE_PATH_TILED and ctx.nlist_overflow() == 0
    ctx.force(capi.FORCE | capi.EPOT)
    assert util.relerr(ctx.download(capi.F_FP, capi.ORDER_CELL), fp) < FORCE_RTOL
    assert util.relerr(ctx.download(capi.F_DEN, capi.ORDER_CELL), den) < FORCE_RTOL
    assert util.relerr(ctx.download(capi.F_EPOT, capi.ORDER_CELL), ep) < FORCE_RTOL
    # 6 steps at 2000 K without a rebuild (NB_UPTAB = 100): the same list serves moved atoms; generic path is the witness
    gen = util.make_ctx(c, force_path=capi.FORCE_PATH_GENERIC)
    gen.force(capi.FORCE)
    for k in range(3):
        ctx.run(1 + 2 * k, 2, 0, 100, 1.0e-15)
        gen.run(1 + 2 * k, 2, 0, 100, 1.0e-15)
        assert util.relerr(ctx.download(capi.F_XP), gen.download(capi.F_XP)) < 1e-12
        assert util.relerr(ctx.download(capi.F_FP), gen.download(capi.F_FP)) < 1e-9
    ctx.close(); gen.close()


def test_parrep_cycle_through_the_mdlib_interface():
    """One parallel-replica cycle of examples/PARREP_Test driven by its own control file through the reference's procedure
    names (Appshell/MD_Method_ParRep_GPU.F90: thermalise -> MD block -> Do_Damp -> compare with the quenched start, Do_Compare /
    &DRTOL): two replicas of the 2000 W + 1 H box as MULTIBOX, 1.6 x RU lists, MAXNB 400, "ST" quench with the file's
    &STEPBOUND / &DELTAPOT.  40 fs at 300 K leave both replicas in the basin they started from: the quench returns every atom
    to within DRTOL (0.02 LU) of the quenched start, lowers the energy, and no event is flagged."""
    from msmpscu_b200 import inputs, mdlib
    g = util.GOLD
    box = inputs.read_box_file(os.path.join(g, "W_2000_H1_EAM1_box.dat"))
    ctl = inputs.read_ctrl_file(os.path.join(g, "parrep_CtrlFile300K.dat"), box)
    neb = np.load(os.path.join(g, "neb_gmd_react.npz"))
    boxes = []
    for r in range(2):
        b = inputs.read_box_file(os.path.join(g, "W_2000_H1_EAM1_box.dat"))
        b.ITYP = neb["ityp"].astype(np.int32)
        b.XP = neb["pos"] * b.RR
        b.allocate()
        boxes.append(b)
    dev = mdlib.DeviceState(0)
    fc = mdlib.Register_ForceClass(box.PotType)
    mdlib.Initialize_Globle_Variables_DEV(dev, boxes, ctl)
    mdlib.Init_Forcetable_Dev(dev, boxes, ctl, fc)
    mdlib.Initialize_NeighboreList_DEV(dev, boxes, ctl)
    assert mdlib.Cal_NeighBoreList_DEV(dev, boxes, ctl) == 0
    mdlib.CalForce_ForceClass(dev, boxes, ctl, fc)
    # the quenched start (SimBox0K of Do_ChangeDetect)
    assert mdlib.Do_Damp(dev, boxes, ctl, fc)[0] != 0
    x0 = dev.ctx.download(capi.F_XP)
    e0 = dev.ctx.download(capi.F_EPOT).sum()
    # thermalise both replicas (different numbers per replica: the key is the ORIGINAL atom id) and run 80 steps
    mdlib.Thermalizing_MC_DEV(dev, boxes, ctl, 600.0)
    v = dev.ctx.download(capi.F_XP1)
    assert not np.allclose(v[:2001], v[2001:])
    mdlib.For_Steps(dev, 0, 80, boxes, ctl)
    mdlib.CalEpot_ForceClass(dev, boxes, ctl, fc)
    e_hot = dev.ctx.download(capi.F_EPOT).sum()
    assert e_hot > e0
    mdlib.Cal_NeighBoreList_DEV(dev, boxes, ctl)
    fl = mdlib.Do_Damp(dev, boxes, ctl, fc)
    assert fl[0] != 0
    mdlib.ResetXP1(dev, boxes)
    x1 = dev.ctx.download(capi.F_XP)
    e1 = dev.ctx.download(capi.F_EPOT).sum()
    d = x1 - x0
    d -= np.round(d / box.ZL) * box.ZL
    dmax = np.sqrt((d ** 2).sum(axis=1)).max() / box.RR
    assert dmax < ctl.STRCUT_DRTol, dmax          # Do_Compare: no atom moved by more than DRTOL -> no event
    assert abs(e1 - e0) < 0.02 * 1.60219e-12      # back in the same minimum (the quench stops at 1e-5 eV per atom per step)
    dev.ctx.close()


def test_parrep_dephase_through_the_mdlib_interface():
    """Do_DePhase (Appshell/MD_Method_ParRep_GPU.F90:959-1090) through the reference's procedure names on the shipped PARREP_Test
    inputs: three replicas start from the quenched configuration, are thermalised IVTIME times IVPAS steps apart and run
    (IVTIME+1)*IVPAS steps, then quenched and compared with the start on the device; replicas that stayed in the basin are kept.
    Checked: the schedule (a raw C-ABI replay of the same calls gives the same bits), the replicas decorrelate, none is dropped
    at the file's &DRTOL, all are dropped when the tolerance is made smaller than the quench's own residual."""
    from msmpscu_b200 import inputs, mdlib
    g = util.GOLD
    path = os.path.join(g, "W_2000_H1_EAM1_box.dat")
    box = inputs.read_box_file(path)
    ctl = inputs.read_ctrl_file(os.path.join(g, "parrep_CtrlFile300K.dat"), box)
    assert (ctl.IVTIME, ctl.IVPAS, ctl.TI) == (10, 100, 300.0)       # &THERMALIZATION / &TEMPERATURE of the file
    ctl.IVTIME, ctl.IVPAS = 2, 40                                     # a short schedule for the test
    neb = np.load(os.path.join(g, "neb_gmd_react.npz"))
    nrep = 3

    def fresh():
        b = inputs.read_box_file(path)
        b.ITYP = neb["ityp"].astype(np.int32)
        b.XP = neb["pos"] * b.RR
        b.allocate()
        return b
    # SimBoxIni: the quenched start
    dev = mdlib.DeviceState(0)
    fc = mdlib.Register_ForceClass(box.PotType)
    ini = fresh()
    mdlib.Initialize_Globle_Variables_DEV(dev, [ini], ctl)
    mdlib.Init_Forcetable_Dev(dev, [ini], ctl, fc)
    mdlib.Initialize_NeighboreList_DEV(dev, [ini], ctl)
    assert mdlib.Cal_NeighBoreList_DEV(dev, [ini], ctl) == 0
    mdlib.CalForce_ForceClass(dev, [ini], ctl, fc)
    assert mdlib.Do_Damp(dev, [ini], ctl, fc)[0] != 0
    ini.XP = dev.ctx.download(capi.F_XP)
    dev.ctx.close()

    dev = mdlib.DeviceState(0)
    boxes = [fresh() for _ in range(nrep)]
    draw0 = mdlib._thermalize_draw[0]
    temps = []
    ncur, fb = mdlib.Do_DePhase(dev, ini, boxes, ctl, ForceClass=fc, log=lambda it, t: temps.append((it, t)))
    assert [it for it, _ in temps] == [1, 41, 81]
    assert temps[0][1] == 0.0 and 50.0 < temps[1][1] < 400.0 and 50.0 < temps[2][1] < 400.0   # TI redrawn, then equipartition
    assert ncur == nrep and not np.any(fb)
    assert not np.allclose(boxes[0].XP1, boxes[1].XP1) and np.abs(boxes[1].XP - boxes[2].XP).max() > 1.0e-3 * ini.RR
    # the same calls through the raw C ABI
    n = ini.NPRT
    xp_dev = dev.ctx.download(capi.F_XP)
    for ib in range(nrep):
        assert np.array_equal(boxes[ib].XP, xp_dev[ib * n:(ib + 1) * n])
    raw = mdlib.DeviceState(0)
    rb = [fresh() for _ in range(nrep)]
    for b in rb:
        b.XP = ini.XP.copy()
    mdlib.Initialize_Globle_Variables_DEV(raw, rb, ctl)
    mdlib.Init_Forcetable_Dev(raw, rb, ctl, fc)
    mdlib.Initialize_NeighboreList_DEV(raw, rb, ctl)
    c = raw.ctx
    assert c.nlist_build() == 0
    c.force(capi.FORCE)
    seed, k = int(ctl.SEED[0]), 0
    for it in range(1, 121):
        if (it - 1) % 40 == 0 and k < 2:
            c.thermalize(300.0, seed, draw0 + k)
            k += 1
        c.predict(ctl.H)
        if it % ctl.NB_UPTAB == 0:
            c.nlist_build()
        c.force(capi.FORCE)
        c.correct(ctl.H)
    assert np.array_equal(c.download(capi.F_XP), xp_dev)
    assert np.array_equal(c.download(capi.F_XP1), dev.ctx.download(capi.F_XP1))
    raw.ctx.close()
    # a tolerance below what the quench leaves (it stops on an energy criterion): every replica counts as having left
    ctl.STRCUT_DRTol = 1.0e-12
    dev.ctx.close()
    dev = mdlib.DeviceState(0)
    ncur2, fb2 = mdlib.Do_DePhase(dev, ini, boxes, ctl, ForceClass=fc)
    assert ncur2 == 0 and np.all(np.asarray(fb2) != 0)
    dev.ctx.close()


def test_slab_domain_single_rank_step_and_virial():
    """SlabDomain with one rank (no process group): its step is the same GMD step as mdb_run (For_One_Step,
    Appshell/MD_Method_GenericMD_GPU.F90:596-627) and force_virial returns the tensor of pCalPTensor
    (MD_EAM_ForceTable_GPU.F90:1366) -- the N = 1 end of the slab-decomposed path that tests/test_gpu_dd.py runs on 2 and 4 GPUs."""
    from msmpscu_b200.domain import SlabDomain
    c = util.bcc_case((8, 8, 12), seed=77, temp=600.0)
    h, it0, nup, nsteps = 0.5e-15, 1, 5, 12
    epc = ([1], [300.0], [1.0e-12], [0.1], [100.0 * util.CP_EVERG])
    full = util.make_ctx(c, force_path=capi.FORCE_PATH_TILED)
    full.epc_set(*epc)
    full.force(capi.FORCE)
    full.run(0, nsteps, it0, nup, h)
    ctx = util.make_ctx(c, build=False, force_path=capi.FORCE_PATH_TILED)
    ctx.epc_set(*epc)
    dom = SlabDomain(ctx, 0)
    assert dom.world == 1
    dom.rebuild()
    ctx.force(capi.FORCE)
    for it in range(nsteps):
        dom.step(it, it0, nup, h)
    assert dom.owned() == (0, ctx.n)
    for f, tol in ((capi.F_XP, 1e-13), (capi.F_XP1, 1e-10), (capi.F_FP, 1e-10)):
        assert util.relerr(ctx.download(f, capi.ORDER_CELL), full.download(f, capi.ORDER_CELL)) < tol
    vt = dom.force_virial()
    vt_full = full.force(capi.FORCE | capi.VIRIAL)
    assert np.abs(vt_full).max() > 0.0 and util.relerr(vt, vt_full) < 1e-10
    # Cal_GlobalT_DEV over the owned atoms (MD_DiffScheme_GPU.F90:1042-1064)
    t_full = full.global_t()
    assert 100.0 < t_full < 2000.0 and abs(dom.global_t() - t_full) < 1e-10 * t_full
    full.close(); ctx.close()


@pytest.mark.parametrize("name", ["bcc_7x9x11", "neb_WH", "fcc_cu_setfl"])
def test_tiled_launch_variants_parity(oracle, name):
    """Every selectable shape of the tiled kernels -- 2 / 4 / 8 lanes per atom, 512- / 768-thread pass CTAs, 2 / 3 pipeline
    stages, with and without the distance classes -- against the CPU oracle (CalForceTest.F90:113-147 restated): forces, DEN,
    EPOT and the virial of the force pass's virial epilogue; the neighbour counts stay the reference's bit for bit."""
    c = CASES[name]()
    ref = _oracle_list(oracle, c)
    gid = ref["gid"] - 1
    T = util.oracle_tables(oracle, c)
    fp, den, vt, ep = oracle.force(c.xp[gid], c.ityp[gid], ref["statu"][gid], ref["kvois"], ref["indi"], c.zl, c.ifpd,
                                   T, virial=True, epot=True)
    ran = 0
    for lanes in (2, 4, 8):
        for threads in (512, 768):
            for stages, classes in ((2, 1), (3, 1), (2, 0)):
                ctx = util.make_ctx(c, build=False, force_path=capi.FORCE_PATH_TILED)
                ctx.set_option(capi.OPT_TILED_LANES, lanes)
                ctx.set_option(capi.OPT_TILED_THREADS, threads)
                ctx.set_option(capi.OPT_TILED_STAGES, stages)
                ctx.set_option(capi.OPT_TILED_CLASSES, classes)
                ctx.nlist_build()
                tag = "lanes %d threads %d stages %d classes %d" % (lanes, threads, stages, classes)
                assert ctx.get_option(capi.OPT_ACTIVE_PATH) == capi.FORCE_PATH_TILED, tag
                vt_gpu = ctx.force(capi.FORCE | capi.VIRIAL | capi.EPOT)
                assert util.relerr(ctx.download(capi.F_FP, capi.ORDER_CELL), fp) < FORCE_RTOL, tag
                assert util.relerr(ctx.download(capi.F_DEN, capi.ORDER_CELL), den) < FORCE_RTOL, tag
                assert util.relerr(ctx.download(capi.F_EPOT, capi.ORDER_CELL), ep) < FORCE_RTOL, tag
                assert util.relerr(vt_gpu, vt / c.nbox) < FORCE_RTOL, tag
                ctx.force(capi.FORCE)          # the force-only kernels (pass 1 + pass 2 without the epilogue)
                assert util.relerr(ctx.download(capi.F_FP, capi.ORDER_CELL), fp) < FORCE_RTOL, tag
                kv, _ = ctx.nlist_copyout(capi.ORDER_CELL)
                assert np.array_equal(kv, ref["kvois"]), tag
                ctx.close()
                ran += 1
    assert ran == 18
