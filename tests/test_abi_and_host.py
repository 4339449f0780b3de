"""CPU tests of the boundary: the C-ABI library loads, exports every symbol include/mdpscu_b200.h declares,
refuses to run without a GPU (no CPU fallback), and its host-side table generator matches the oracle bit for bit."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import util
from msmpscu_b200 import capi, forcetable

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "mdpscu_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = capi.load()
    names = _declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "libmdpscu_b200.so does not export %s" % n
    # and the python binding table covers the header
    assert set(names) == set(capi.SYMBOLS)
    assert lib.mdb_version().decode().startswith("mdpscu_b200")


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "mdpscu_b200.h"\nint main(void){ mdb_ctx *c = 0; (void)c; return MDB_OK; }\n')
    import subprocess
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                           "-o", str(tmp_path / "t.o")])


@pytest.mark.skipif(capi.load().mdb_device_count() > 0, reason="a GPU is present")
def test_no_cpu_fallback_without_gpu():
    with pytest.raises(capi.MDBError) as e:
        capi.Context(0)
    assert e.value.code == capi.ERR_NOGPU


def test_product_tables_equal_oracle_tables(oracle):
    """Two independent implementations of Create_Interaction_ForceTable (C oracle, C++ product)."""
    for c in (util.bcc_case((4, 4, 4)), util.neb_case("react")):
        t = util.product_tables(c)
        o = util.oracle_tables(oracle, c)
        assert (t.nkind, t.nkind1, t.csi, t.rhod) == (o.nkind, o.nkind1, o.csi, o.rhod)
        assert np.array_equal(t.kpair, o.kpair) and np.array_equal(t.kembd, o.kembd)
        for name in ("potr", "fpotr", "potb", "fpotb", "fembd", "dfembd"):
            assert np.array_equal(getattr(t, name), getattr(o, name)), name


def test_table_grid_definition():
    """r_i = (i*sqrt(Rmax)/NTAB)^2, POTR = V/2*r, FPOTR = -V'*r, POTB = rho, FPOTB = -rho'
    (MD_TypeDef_ForceTable.F90:949-976): spot-check the grid against direct evaluation."""
    c = util.bcc_case((4, 4, 4))
    t = util.product_tables(c)
    potb = t.as_2d("potb")[0]
    csiv = 1.0 / t.csi
    i = 8000
    r = (i * csiv) ** 2
    # Marinica rho: sum b_k (rho_k - r)^3 H(rho_k - r), knots as float32 literals
    b = [-0.420429107805055e1, 0.518217702261442e0, 0.562720834534370e-1, 0.344164178842340e-1]
    rk = [float(np.float32(v)) for v in (2.5, 3.1, 3.5, 4.9)]
    ra = r * 1e8
    want = sum(bk * (k - ra) ** 3 for bk, k in zip(b, rk) if k - ra >= 0)
    assert abs(potb[i - 1] - want) <= 4e-16 * abs(want)
    assert t.rhod == pytest.approx(potb.max() * 20.0 / t.nembd, rel=1e-15)


def test_unknown_potential_id_is_an_error_code():
    with pytest.raises(capi.MDBError):
        forcetable.Create_Interaction_ForceTable(capi.LIB_MARINICA_EAM2, [[7]], 100, 100, 6e-8)


def test_philox_known_answers():
    """Philox4x32-10 (the generator behind mdb_thermalize): Random123's published known-answer vectors, through the
    product's host entry point and through the oracle."""
    import ctypes as C
    from msmpscu_b200 import capi
    from oracle import pyorc
    lib = capi.load()
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        out = (C.c_uint * 4)()
        assert lib.mdb_philox4x32_10((C.c_uint * 4)(*ctr), (C.c_uint * 2)(*key), out) == 0
        assert tuple(out) == want
        assert tuple(pyorc.philox4x32_10(ctr, key)) == want
    bits = (C.c_uint * 8)()
    assert lib.mdb_thermalize_bits(0x1234ABCD5678, 3, 42, bits) == 0
    assert list(bits[:4]) == pyorc.philox4x32_10((42, 3, 0, 0), (0xABCD5678, 0x1234))
    assert list(bits[4:]) == pyorc.philox4x32_10((42, 3, 1, 0), (0xABCD5678, 0x1234))


def test_bench_reference_arm_and_no_gpu_behaviour():
    """bench.py --impl reference prints one JSON line with the contract's keys (CPU restatement on the host cores); the
    product arm refuses to run without a CUDA device instead of falling back to anything."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cells", "8"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, BENCH_CPU_QUICK="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
                "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
    # the line says what was really run: atoms of the timed box, threads, MD steps per reference step
    assert line["config"]["atoms_total"] == 2 * 8 ** 3 and "%d atoms" % (2 * 8 ** 3) in line["config"]["workload"]
    assert line["cpu_baseline"]["cores"] == os.cpu_count() and line["config"]["md_steps_per_step"] == 10
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    import torch
    if not torch.cuda.is_available():
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                             text=True, timeout=300)
        assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)


def test_control_files_of_the_reference_examples_are_read():
    """The reference's own control files for BASELINE configs[0] (GMD_Test), the NEB_Test quench run and configs[3]
    (PARREP_Test) through msmpscu_b200.inputs: cut-offs, list control, time step and the quench / event-detection
    keywords (Common/MD_TypeDef_SimCtrlParam.F90:1694-2060) land in the SimMDCtrl fields mdlib.Do_Damp reads."""
    from msmpscu_b200 import inputs
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    box = inputs.read_box_file(os.path.join(g, "W_2000_H1_EAM1_box.dat"))
    neb = inputs.read_ctrl_file(os.path.join(g, "CtrlFile0K.dat"), box)
    assert neb.Quench_Steps == 1000 and neb.Quench_Meth == "ST" and not neb.Quench_LSearch
    assert neb.NB_MXNBS == 256 and neb.NB_UPTAB == 10 and abs(neb.NB_RM[0, 0] / neb.RU[0, 0] - 1.2) < 1e-12
    assert neb.SEED == [123456] and abs(neb.H - 0.5e-15) < 1e-30
    par = inputs.read_ctrl_file(os.path.join(g, "parrep_CtrlFile300K.dat"), box)
    assert par.NB_MXNBS == 400 and abs(par.NB_RM[0, 0] / par.RU[0, 0] - 1.6) < 1e-12
    assert par.Quench_Steps == 1000 and par.Quench_Meth == "ST"
    assert par.STEEPEST_MiStep == 1.0e-5 and par.STEEPEST_MxStep == 0.1 and par.STEEPEST_MiDelE == 1.0e-5
    assert par.STRCUT_DRTol == 0.02 and par.LBFGS_MSave == 7


def test_step_size_and_list_period_schedules_are_read(tmp_path):
    """the two lines of the reference's cascade control file (examples/Cascade_Test/CtrlFile300K_LOC_T.dat) that steer the time
    loop: &STEPSIZE flag / hmi / hmx [fs] / dmx [Angstrom, MD_Gvar.F90:947] and &UPDATEFRE min / max / doubling interval"""
    from msmpscu_b200 import inputs
    g = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    box = inputs.read_box_file(os.path.join(g, "W_2000_H1_EAM1_box.dat"))
    p = tmp_path / "ctrl.dat"
    p.write_text("&CTLF\n &POTENSUBCTL\n  &CUTOFF cutoff distances: A-A= 1.9\n &ENDSUBCTL\n &SECTSUBCTL #1\n  &TIMESUBCTL\n"
                 "   &STEPSIZE   use fixd step flag= -1, hmi= 0.25, hmx = 0.5, dmx = 0.05\n  &ENDSUBCTL\n  &NEIGHBSUBCTL\n"
                 "   &UPDATEFRE minimun frequcency= 5, max frequence= 10, frequecy of changing updating frequency 100\n"
                 "   &CUTOFF    cutoff between neighbors = 1.6\n  &ENDSUBCTL\n &ENDSUBCTL\n&ENDCTLF\n")
    c = inputs.read_ctrl_file(str(p), box)
    assert c.IHDUP == -1 and abs(c.HMI - 0.25e-15) < 1e-30 and abs(c.HMX - 0.5e-15) < 1e-30 and abs(c.DMX - 0.05e-8) < 1e-24
    assert abs(c.H - c.HMI) < 1e-30                       # "the time step starts from its minimum value"
    assert (c.NB_UPTABMI, c.NB_UPTABMX, c.NB_DBITAB, c.NB_UPTAB) == (5, 10, 100, 5)
    d = mdlib_defaults = __import__("msmpscu_b200.mdlib", fromlist=["SimMDCtrl"]).SimMDCtrl()
    assert (d.IHDUP, d.NB_UPTABMI, d.NB_UPTABMX, d.NB_DBITAB) == (0, 10, 10, 100000) and d.DMX == 0.1e-8   # :886-890, :940-942


def test_external_stopping_table_file_is_read(tmp_path):
    """the `&MDPSCU_STPTAB.stp` format of the reference's stopping tables (Load_STPTable / Import_STPTable): units keV and keV cm^2,
    column directives per pair, re-gridding to ETAB / STAB / KPAIR of mdb_stopping_set"""
    from msmpscu_b200 import inputs
    e = np.linspace(0.02, 200.0, 101)
    k = {"W->W": 1.0278e-17, "W->HE": 1.7337e-18, "HE->W": 3.2238e-18, "HE->HE": 8.599e-19}   # keV cm^2 per sqrt(keV), as the shipped table
    p = tmp_path / "tab.stp"
    with open(p, "w") as f:
        f.write("&MDPSCU_STPTAB.stp\n! energy in the unit of keV\n! stoping in the unit of keV*cm^2\n&NUMTABLE       4\n&NUMPOINT     101\n")
        f.write("&W->W       with COL#    2     74.00    183.84     74.00    183.84\n&W->He      with COL#    3     74.00    183.84      2.00      4.00\n")
        f.write("&He->W      with COL#    4      2.00      4.00     74.00    183.84\n&He->He     with COL#    5      2.00      4.00      2.00      4.00\n")
        f.write("!--- ENERGY        W->W            W->He           He->W           He->He\n")
        for x in e:
            f.write("  %.8E" % x + "".join("  %.8E" % (k[q] * np.sqrt(x)) for q in ("W->W", "W->HE", "HE->W", "HE->HE")) + "\n")
    ee, tabs, ids = inputs.read_stopping_table(str(p))
    kev = 1000.0 * 1.60219e-12
    assert set(tabs) == set(k) and ids["W->HE"] == (74.0, 183.84, 2.0, 4.0)
    assert np.allclose(ee, e * kev, rtol=1e-8) and np.allclose(tabs["HE->W"], k["HE->W"] * np.sqrt(e) * kev, rtol=1e-8)
    etab, stab, kpair = inputs.stopping_tables_for(str(p), ["W", "He"], 20.0, 200000.0, 2000)
    assert etab.shape == (2001,) and stab.shape == (2001, 4) and np.array_equal(kpair, [[1, 2], [3, 4]])
    assert abs(etab[0] - 20.0 * 1.60219e-12) < 1e-24 and abs(etab[-1] - 2.0e5 * 1.60219e-12) < 1e-18
    assert np.allclose(stab[:, 0], k["W->W"] * np.sqrt(etab / kev) * kev, rtol=2e-3)     # (linear re-gridding of a sqrt law on 101 points)
    with pytest.raises(ValueError):
        inputs.stopping_tables_for(str(p), ["W", "H"], 20.0, 1000.0, 10)


def _thermal_loop_restatement(XP1, STATU, ITYP, CM, PROP, BOXSHAPE, ZL, VTENSOR, EPOT, EKIN):
    """Cal_thermal_quantities_SimMDBox (Common/MD_TypeDef_SimBox.F90:5048-5170) as plain loops, statement by statement:
    the checker of the vectorised host mirror in msmpscu_b200/mdlib.py."""
    from msmpscu_b200.constants import CP_KB, CP_EVERG
    n, ng = len(STATU), len(CM)
    fixbits = (2, 4, 8)
    vv0, tcm, anprt = [0.0, 0.0, 0.0], 0.0, 0
    for k in range(ng):
        if any((PROP[k] & b) == b for b in fixbits):
            continue
        vs = [0.0, 0.0, 0.0]
        for j in range(n):
            if ITYP[j] == k + 1 and (STATU[j] & 1) == 1:
                for d in range(3):
                    vs[d] += XP1[j, d]
                tcm += CM[k]
                anprt += 1
        for d in range(3):
            vv0[d] += vs[d] * CM[k]
    vv0 = [v / tcm for v in vv0]
    cxp1 = np.zeros((n, 3))
    for i in range(n):
        if (STATU[i] & 1) == 1:
            for d in range(3):
                if (STATU[i] & fixbits[d]) == 0:
                    cxp1[i, d] = sum(BOXSHAPE[d, j] * (XP1[i, j] - vv0[d]) for j in range(3))
    ket = np.zeros((3, 3))
    for k in range(ng):
        if any((PROP[k] & b) == b for b in fixbits):
            continue
        for j in range(n):
            if ITYP[j] == k + 1 and (STATU[j] & 1) == 1:
                for k1 in range(3):
                    for k2 in range(3):
                        ket[k1, k2] += cxp1[j, k1] * cxp1[j, k2] * CM[k]
    volume = float(np.linalg.det(BOXSHAPE)) * ZL[0] * ZL[1] * ZL[2]
    temp = (ket[0, 0] + ket[1, 1] + ket[2, 2]) / (3.0 * anprt * CP_KB)
    sp0 = n * CP_KB * temp / volume * 1.0e-9
    sp1 = (VTENSOR[0, 0] + VTENSOR[1, 1] + VTENSOR[2, 2]) / 3.0 / volume * 1.0e-9
    act = [(STATU[i] & 1) == 1 for i in range(n)]
    avepot = sum(EPOT[i] for i in range(n) if act[i]) / CP_EVERG / sum(act)
    harmil = avepot * CP_EVERG + sum(EKIN[i] for i in range(n) if act[i] and (STATU[i] & 14) == 0) / anprt
    return dict(KTENSOR=ket, VOLUME=volume, TEMPERATURE=temp, SPRESS0=sp0, SPRESS1=sp1, SPRESS=sp0 + sp1,
                PTENSOR=(ket + VTENSOR) / volume, AVEPOT=avepot, HARMIL=harmil)


@pytest.mark.parametrize("sheared", [False, True])
def test_thermal_quantities_mirror(sheared):
    """The output-step host routine behind the path (temperature from the drift-free kinetic tensor, pressure from
    KTENSOR + VTENSOR, cohesive energy, Hamiltonian): three groups, one of them with a FIXPOS property (left out of the
    drift and of the tensor), atoms with single fixed components, inactive atoms, and a non-identity BOXSHAPE."""
    from msmpscu_b200 import mdlib
    from msmpscu_b200.constants import CP_AU2G
    rng = np.random.default_rng(11)
    n = 600
    ityp = np.repeat(np.array([1, 2, 3], dtype=np.int32), n // 3)
    statu = np.full(n, 1, dtype=np.int32)
    statu[rng.choice(n, 40, replace=False)] |= 2
    statu[rng.choice(n, 30, replace=False)] |= 8
    statu[rng.choice(n, 25, replace=False)] = 0            # inactive
    cm = np.array([183.84, 1.008, 4.0026]) * CP_AU2G
    prop = np.array([1, 1, 1 | 4], dtype=np.int32)          # the third group carries FIXPOSY
    shape = np.eye(3)
    if sheared:
        shape = np.array([[1.0, 0.08, -0.03], [0.0, 0.97, 0.05], [0.02, 0.0, 1.04]])
    box = mdlib.SimMDBox(NPRT=n, NGROUP=3, RR=3.1652e-8, ZL=np.array([6.0, 7.0, 8.0]) * 3.1652e-8, CM=cm, ITYP=ityp, STATU=statu,
                         BOXSHAPE=shape, PROP=prop)
    box.allocate()
    box.XP1 = rng.normal(size=(n, 3)) * 3.0e4 + np.array([1.0e3, -2.0e3, 5.0e2])      # cm/s, with a drift
    box.EPOT = -8.9 * 1.60219e-12 + rng.normal(size=n) * 1.0e-13
    box.EKIN = 0.5 * cm[ityp - 1] * (box.XP1 ** 2).sum(axis=1)
    box.VTENSOR = rng.normal(size=(3, 3)) * 1.0e-9
    got = mdlib.Cal_thermal_quantities(box)
    ref = _thermal_loop_restatement(box.XP1, statu, ityp, cm, prop, shape, box.ZL, box.VTENSOR, box.EPOT, box.EKIN)
    for key, val in ref.items():
        assert np.allclose(got[key], val, rtol=1e-12, atol=0.0), key
        assert np.allclose(getattr(box, key), val, rtol=1e-12, atol=0.0), key
    assert 200.0 < got["TEMPERATURE"] < 2.0e4 and got["SPRESS0"] > 0.0


def test_event_detection_compare():
    """Do_Compare / the replica bookkeeping of Do_ChangeDetect (Appshell/MD_Method_ParRep_GPU.F90:1241-1297, 1146-1156) against
    a per-atom loop: an atom that hopped by less than DRTOL is not an event, one that crossed a periodic face by a small step
    is not an event either (minimum image), a 0.5 LU hop is, a masked atom never is, and along a non-periodic axis the raw
    separation counts."""
    from msmpscu_b200 import mdlib
    rng = np.random.default_rng(3)
    n, rr = 200, 3.14e-8
    zl = np.array([10.0, 10.0, 10.0]) * rr
    ini = mdlib.SimMDBox(NPRT=n, NGROUP=1, RR=rr, ZL=zl, BOXLOW=-0.5 * zl)
    ini.XP = (rng.random((n, 3)) - 0.5) * zl
    ini.XP[0] = [4.995 * rr, 0.0, 0.0]
    ini.XP[4] = [0.0, 0.0, 4.995 * rr]
    ctl = mdlib.SimMDCtrl(IFPD=np.array([1, 1, 0], dtype=np.int32), STRCUT_DRTol=0.02)
    reps = []
    for r in range(3):
        b = mdlib.SimMDBox(NPRT=n, NGROUP=1, RR=rr, ZL=zl, BOXLOW=-0.5 * zl)
        b.XP = ini.XP + rng.normal(size=(n, 3)) * 0.004 * rr
        reps.append(b)
    reps[0].XP[0] = [-4.998 * rr, 0.0, 0.0]          # wrapped through the periodic x face: 0.007 LU away
    reps[1].XP[7] += [0.5 * rr, 0.0, 0.0]            # a hop
    reps[1].XP[9] += [0.0, 0.5 * rr, 0.0]            # a hop of a masked atom
    reps[2].XP[4] = [0.0, 0.0, -4.998 * rr]          # z is not periodic: 9.993 LU away
    mask = np.ones(n, dtype=np.int32); mask[9] = 0
    flag = mdlib.Do_Compare(ini, reps, ctl, mask)
    ref = np.zeros(3 * n, dtype=np.int32)
    for ib, b in enumerate(reps):
        for i in range(n):
            if mask[i] <= 0:
                continue
            sep = ini.XP[i] - b.XP[i]
            for d in range(3):
                if ctl.IFPD[d] > 0 and abs(sep[d]) > 0.5 * zl[d]:
                    sep[d] -= np.copysign(zl[d], sep[d])
            ref[ib * n + i] = 1 if (sep * sep).sum() > (0.02 * rr) ** 2 else 0
    assert np.array_equal(flag, ref)
    assert flag[0] == 0 and flag[n + 7] == 1 and flag[n + 9] == 0 and flag[2 * n + 4] == 1 and flag.sum() == 2
    assert mdlib.Transition_Replicas(flag, n) == (3, 2)
    assert mdlib.Transition_Replicas(mdlib.Do_Compare(ini, reps[:1], ctl), n) == (0, 0)


def test_ctrl_defaults_and_sections_follow_the_reference():
    """MD_TypeDef_SimCtrlParam.F90:202-204,996-997: STRCUT_DRTol = 0.03 LU, Quench_Steps = 1000 when the file is silent;
    a control file with several &SECTSUBCTL blocks yields one SimMDCtrl per section (sectCtrlParam list)"""
    from msmpscu_b200 import inputs, mdlib
    d = mdlib.SimMDCtrl()
    assert d.STRCUT_DRTol == 0.03 and d.Quench_Steps == 1000
    g = os.path.join(os.path.dirname(__file__), "golden")
    box = inputs.read_box_file(os.path.join(g, "gmd_W_8192_EAM_box.dat")) if os.path.exists(os.path.join(g, "gmd_W_8192_EAM_box.dat")) else None
    if box is None:
        import glob
        box = inputs.read_box_file(sorted(glob.glob(os.path.join(g, "gmd_*box*.dat")))[0])
    secs = inputs.read_ctrl_sections(os.path.join(g, "gmd_CtrlFile300K.dat"), box)
    assert len(secs) == 3
    assert secs[0].TEMP == 0.0 and secs[0].Quench_Steps == 1000 and secs[0].Quench_Meth == "ST"
    assert secs[1].TEMP == 300.0
    one = inputs.read_ctrl_file(os.path.join(g, "gmd_CtrlFile300K.dat"), box)
    assert one.TEMP == secs[0].TEMP and one.NB_MXNBS == 256 and one.NB_UPTAB == 10
    assert one.STRCUT_DRTol == 0.03          # the file has no &DRTOL
    # thermalisation schedule: defaults MD_TypeDef_SimCtrlParam.F90:918-920, &THERMALIZATION per MD_SimCtrlParam_GMD.F90:97-125
    assert (d.IVTIME0, d.IVTIME, d.IVPAS) == (1, 0, 50)
    assert [(s.IVTIME, s.IVPAS, s.TI) for s in secs] == [(0, 50, 0.0), (100, 100, 300.0), (0, 50, 0.0)]
    par = inputs.read_ctrl_file(os.path.join(g, "parrep_CtrlFile300K.dat"), box)
    assert (par.IVTIME, par.IVPAS, par.TI) == (10, 100, 300.0)


def test_fortran_binding_covers_the_whole_header():
    """fortran/mdb_c_binding.F90 is generated from include/mdpscu_b200.h (tools/gen_fortran_binding.py): every entry point of
    the header has a bind(C) interface with the same name and the same number of arguments, and the committed file is fresh.
    fortran/mdb_shims.F90 only calls bound entry points and defines the procedures INTEGRATION.md's table names."""
    import importlib.util
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_fortran_binding", os.path.join(root, "tools", "gen_fortran_binding.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    decls = gen.declarations()
    from msmpscu_b200 import capi
    assert {d[1] for d in decls} == set(capi.SYMBOLS), "header and capi.SYMBOLS disagree"
    text = open(os.path.join(root, "fortran", "mdb_c_binding.F90")).read()
    assert text == gen.generate(), "fortran/mdb_c_binding.F90 is stale: run python tools/gen_fortran_binding.py"
    flat = re.sub(r"&\s*\n\s*", " ", text)
    for ret, name, args in decls:
        m = re.search(r"(?:function|subroutine)\s+%s\(([^)]*)\)\s+bind\(C, name=\"%s\"\)" % (name, name), flat)
        assert m, name
        got = [a for a in m.group(1).split(",") if a.strip()]
        assert len(got) == len(args), (name, got, args)
    shims = open(os.path.join(root, "fortran", "mdb_shims.F90")).read()
    bound = {d[1] for d in decls}
    for called in set(re.findall(r"\b(mdb_[a-z0-9_]+)\s*\(", shims)):
        assert called in bound, "shim calls %s, which the header does not declare" % called
    for proc in ("INITIALIZE_EAM_Force_Table_DEV", "CALFORCE_EAM_Force_Table2A_DEV", "CALPTENSOR_EAM_Force_Table2A_DEV",
                 "UpdateEPOT_EAM_Force_Table2A_DEV", "CALDEN_EAM_Force_Table2A_DEV", "Cal_EAM_AtomicStressTensor_DEV",
                 "Clear_EAM_Force_Table_DEV", "INITIALIZE_FS_Force_Table_DEV", "CALFORCE_FS_Force_Table2A_DEV",
                 "CALPTENSOR_FS_Force_Table2A_DEV", "UpdateEPOT_FS_Force_Table2A_DEV", "CALDEN_FS_Force_Table2A_DEV",
                 "Cal_FS_AtomicStressTensor_DEV", "Clear_FS_Force_Table_DEV", "Initialize_NeighboreList_DEV", "Cal_NeighBoreList_DEV",
                 "Copyout_NeighboreList_A2", "GetCellInform", "Clear_NeighboreList_DEV", "Reorder_NeighBoreList_Nearest_Dev",
                 "Predictor_DEV", "Correction_DEV", "CalEKin_DEV", "Cal_GlobalT_DEV", "VelScaling_DEV", "Thermalizing_MC_DEV",
                 "CheckTimestep_DEV", "Do_EPCForce_DEV", "Do_ResetParam_DEV", "Initialize_GB_A_DEV", "Clear_Globle_Variables_DEV",
                 "CopyAllFrom_Devices_to_Host", "CopyAllFrom_Host_to_Devices", "Synchroniz_XP_on_Devices", "CopyIn_SimBoxA",
                 "CopyOut_SimBoxA", "Initialize_DEVICES", "Do_Steepest_Forsteps_DEV", "Do_CG_Forsteps_DEV", "DO_LBFGSB_FORSTEPS_DEV",
                 "Do_DynDamp_Forsteps_DEV", "Initialize_STMOD_DEV", "Reset_STMOD_DEV", "Do_STMOD_DEV", "Clear_STMOD_DEV",
                 "Initialize_ActiveRegion_DEV", "ActivateRegion_DEV", "Active_All_ActiveRegion_DEV", "DeActive_All_ActiveRegion_DEV",
                 "Do_ChangeDetect_DEV"):
        assert re.search(r"subroutine\s+%s\b" % proc, shims), "shim %s missing" % proc


def _fortran_logical_lines(path):
    """free-form source with comments removed and continuation lines joined"""
    import re
    out, buf = [], ""
    for line in open(path).read().splitlines():
        s = re.sub(r"!.*$", "", line).strip()
        if s.startswith("&"):
            s = s[1:]
        if s.endswith("&"):
            buf += s[:-1] + " "
            continue
        out.append(buf + s)
        buf = ""
    return out


def test_fortran_shims_are_balanced_and_call_the_binding_with_its_arity():
    """No Fortran compiler exists in this image (Oracle), so the shim layer is checked structurally: every
    module / subroutine / function / interface / type / do / if-then / select block of fortran/*.F90 closes in order, and
    every call of an mdb_* entry point in fortran/mdb_shims.F90 passes exactly as many arguments as its bind(C) interface in
    fortran/mdb_c_binding.F90 declares (which test_fortran_binding_covers_the_whole_header ties to the C header)."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    openers = (("module", r"^module\s+(?!procedure\b)\w+"), ("subroutine", r"^(pure\s+|elemental\s+|recursive\s+)*subroutine\s+\w+"),
               ("function", r"^((integer|real|logical|character|type|double)\S*\s+)*function\s+\w+"), ("interface", r"^(abstract\s+)?interface\b"),
               ("type", r"^type\s*(,|::|\s+[a-z])"), ("do", r"^(\w+\s*:\s*)?do\b"), ("if", r"^(\w+\s*:\s*)?if\s*\(.*\)\s*then$"),
               ("select", r"^select\s+case"))
    for name in ("mdb_shims.F90", "mdb_c_binding.F90"):
        stack = []
        for n, raw in enumerate(_fortran_logical_lines(os.path.join(root, "fortran", name))):
            line = raw.strip().lower()
            if not line:
                continue
            m = re.match(r"^end\s*(module|subroutine|function|interface|type|do|if|select)\b", line)
            if m:
                assert stack and stack[-1][0] == m.group(1), "%s: '%s' (logical line %d) closes %s" % (name, line, n, stack[-1:] or "nothing")
                stack.pop()
                continue
            for kind, pat in openers:
                if re.match(pat, line):
                    stack.append((kind, n))
                    break
        assert not stack, "%s: unclosed %s" % (name, stack)
    binding = "\n".join(_fortran_logical_lines(os.path.join(root, "fortran", "mdb_c_binding.F90")))
    arity = {}
    for m in re.finditer(r"(?:function|subroutine)\s+(mdb_\w+)\s*\(([^)]*)\)", binding, re.I):
        arity[m.group(1).lower()] = len([a for a in m.group(2).split(",") if a.strip()])
    assert len(arity) >= 85
    shims = "\n".join(_fortran_logical_lines(os.path.join(root, "fortran", "mdb_shims.F90")))
    checked = 0
    for m in re.finditer(r"\b(mdb_\w+)\s*\(", shims):
        name = m.group(1).lower()
        if name not in arity:
            continue
        i, depth, nargs, empty = m.end(), 1, 1, True
        while depth > 0 and i < len(shims):
            ch = shims[i]
            if ch == "(":
                depth += 1
            elif ch == ")":
                depth -= 1
            elif ch == "," and depth == 1:
                nargs += 1
            elif not ch.isspace():
                empty = False
            i += 1
        nargs = 0 if empty else nargs
        assert nargs == arity[name], "%s called with %d arguments, its interface has %d: %s" % (name, nargs, arity[name], shims[m.start():i][:160])
        checked += 1
    assert checked >= 60


def test_dephase_packs_the_surviving_replicas_in_order():
    """Pack_Replicas = the tail of Do_DePhase (Appshell/MD_Method_ParRep_GPU.F90:1078-1086): replicas 1 and 3 of four left the basin,
    so replicas 0 and 2 end up in boxes 0 and 1 and m_curReplicas = 2; the boxes behind keep what they held."""
    from types import SimpleNamespace
    from msmpscu_b200 import mdlib
    n, nb = 5, 4
    boxes = [SimpleNamespace(XP=np.full((n, 3), -1.0 - b), XP1=np.full((n, 3), -10.0 - b)) for b in range(nb)]
    xp = np.concatenate([np.full((n, 3), float(b)) for b in range(nb)])
    fields = {"XP": xp, "XP1": 10.0 * xp}
    assert mdlib.Pack_Replicas(boxes, fields, np.array([0, 1, 0, 1]), n) == 2
    assert np.all(boxes[0].XP == 0.0) and np.all(boxes[1].XP == 2.0) and np.all(boxes[1].XP1 == 20.0)
    assert np.all(boxes[2].XP == -3.0) and np.all(boxes[3].XP1 == -13.0)
    assert mdlib.Pack_Replicas(boxes, fields, np.ones(nb, dtype=int), n) == 0
    assert mdlib.Pack_Replicas(boxes, fields, np.zeros(nb, dtype=int), n) == nb and np.all(boxes[3].XP == 3.0)
