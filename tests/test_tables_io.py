"""External force tables: the NIST setfl importer and the .pair/.embd files, pinned on the reference's
own output for examples/NIST_Potentials/Cu_EAM (Cu1.eam.fs.setfl -> Cu1.eam.fs.setfl.pair/.embd written by the
reference's Export_ForceTable, 10 and 9 significant digits).  CPU only: host logic, no kernels."""
import os

import numpy as np
import pytest

import util
from msmpscu_b200 import capi, forcetable
from oracle import tables_np

PRINT_PAIR = 6e-10   # 1PE21.9: 10 significant digits
PRINT_EMBD = 6e-9    # 1PE16.8: 9 significant digits
# The file's rho(r) and r*V(r) hold 1e12 for r < 0.1 A and drop to O(10..100) within one grid step; spline values next to
# that step are differences of O(1e12) terms, so they carry an absolute round-off of ~1e12 * 2^-52 * (a few), in the
# reference's B-spline solve as much as in any other.  Rows there are compared with that absolute allowance.
JUMP_ABS = 2e-3


@pytest.fixture(scope="module")
def oracle():
    from oracle import pyorc
    return pyorc


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(util.GOLD, "cu1_setfl_table_rows.npz"))


@pytest.fixture(scope="module")
def setfl(tmp_path_factory):
    return util.cu_setfl_path(tmp_path_factory.mktemp("setfl"))


def _check_rows(pair, embd, r, rho, gold):
    sel = gold["index"] - 1
    assert np.allclose(r[sel], gold["r"], rtol=PRINT_PAIR, atol=0)
    assert np.allclose(rho[sel], gold["rho"], rtol=PRINT_EMBD, atol=1e-300)
    near_jump = gold["r"] < 0.12
    for c in range(4):
        g, m = gold["pair"][:, c], pair[sel, c]
        # derivative columns: the same round-off divided by the file's grid step (6e-4 A)
        tol = PRINT_PAIR * np.abs(g) + np.where(near_jump, JUMP_ABS if c in (0, 2) else JUMP_ABS / 6e-4, 1e-40)
        bad = np.abs(m - g) > tol
        assert not bad.any(), (c, gold["index"][bad][:5], g[bad][:5], m[bad][:5])
    for c in range(2):
        g, m = gold["embd"][:, c], embd[sel, c]
        assert np.all(np.abs(m - g) <= PRINT_EMBD * np.abs(g) + 1e-40), c


def test_oracle_setfl_matches_reference_tables(setfl, gold):
    """the NumPy/SciPy restatement against the reference's exported tables"""
    t = tables_np.setfl_tables(open(setfl).read(), 10000, 10000)
    pair, embd = tables_np.export_columns(t)
    r = (np.arange(1, 10001) / t["csi"]) ** 2 * 1e8
    _check_rows(pair[0], embd[0], r, np.arange(10000) * t["rhod"], gold)


def test_product_setfl_matches_reference_tables(setfl, gold):
    """mdb_host_setfl_ftable (C++) against the reference's exported tables"""
    info = forcetable.setfl_info(setfl)
    assert info["elements"] == ["Cu"] and info["z"][0] == 29 and info["nr"] == 10000
    assert abs(info["cutoff"] - 6.0e-8) < 1e-20 and abs(info["rhomx"] - 300.0) < 1e-12
    t = forcetable.NIST_Register_Interaction_Table(setfl, 10000, 10000)
    ergev = 1.0 / 1.60219e-12
    pair = np.stack([t.potr * 2 * ergev * 1e8, t.fpotr * ergev, t.potb, t.fpotb * 1e-8], axis=1)
    embd = np.stack([t.fembd * ergev, t.dfembd], axis=1)
    r = (np.arange(1, 10001) / t.csi) ** 2 * 1e8
    _check_rows(pair, embd, r, np.arange(10000) * t.rhod, gold)


def test_product_setfl_matches_oracle_full_precision(setfl):
    """both restatements agree far below print precision away from the 1e12 step (different algorithms: tridiagonal
    second-derivative solve in C++ vs SciPy's CubicSpline)"""
    o = tables_np.setfl_tables(open(setfl).read(), 4000, 3000, rmax=5.5e-8)
    t = forcetable.NIST_Register_Interaction_Table(setfl, 4000, 3000, rmax=5.5e-8)
    r = (np.arange(1, 4001) / t.csi) ** 2 * 1e8
    far = r > 0.2
    for name in ("potr", "fpotr", "potb", "fpotb"):
        a, b = getattr(t, name)[far], o[name][0][far]
        assert np.max(np.abs(a - b)) <= 1e-11 * np.max(np.abs(b)), name
    for name in ("fembd", "dfembd"):
        a, b = getattr(t, name), o[name][0]
        assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b)), name
    assert t.csi == o["csi"] and t.rhod == o["rhod"]


def test_export_then_import_round_trip(setfl, gold, tmp_path):
    """Export_ForceTable -> Import_ForceTable/Register_Imported_ForceTable on the same grid gives the tables back to
    print precision, and the exported text carries the reference's numbers."""
    t = forcetable.NIST_Register_Interaction_Table(setfl, 10000, 10000)
    base = str(tmp_path / "cu1")
    forcetable.Export_ForceTable(base, t)
    rows = np.array([[float(v) for v in line.split()] for line in open(base + ".pair") if line.split() and line.split()[0].isdigit()])
    erow = np.array([[float(v) for v in line.split()] for line in open(base + ".embd") if line.split() and line.split()[0].isdigit()])
    assert rows.shape == (10000, 6) and erow.shape == (10000, 4)
    _check_rows(rows[:, 2:], erow[:, 2:], rows[:, 1], erow[:, 1], gold)
    head = open(base + ".pair").read(2000)
    assert "&MDPSCU_POTTAB.Pair" in head and '&POTTYPE "EAM_TYPE"' in head and "&NUMPOINT   10000" in head
    back = forcetable.Register_Imported_ForceTable(base, [[1]], 10000, 10000, t.Rmax)
    assert back.PotType == "EAM_TYPE" and back.nkind == 1 and list(back.kpair) == [1] and list(back.kembd) == [1]
    r = (np.arange(1, 10001) / t.csi) ** 2 * 1e8
    far = r > 0.2
    for name in ("potr", "fpotr", "potb", "fpotb"):
        a, b = getattr(back, name)[far], getattr(t, name)[far]
        assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12 * np.max(np.abs(b)))) < 2e-9, name
    assert abs(back.rhod - t.rhod) < 1e-8 * t.rhod
    assert np.max(np.abs(back.fembd - t.fembd)) < 1e-8 * np.max(np.abs(t.fembd))
    # a coarser run grid (the case Register_Imported_ForceTable exists for): values stay on the source curve
    coarse = forcetable.Register_Imported_ForceTable(base, [[1]], 2500, 2500, t.Rmax)
    a, b = coarse.potr[200:], t.potr[3::4][200:]   # r_k(2500) = r_4k(10000)
    assert np.max(np.abs(a - b)) < 2e-9 * np.max(np.abs(b))


def test_import_errors_are_status_codes(tmp_path):
    with pytest.raises(capi.MDBError):
        forcetable.setfl_info(str(tmp_path / "missing.setfl"))
    with pytest.raises(capi.MDBError):
        forcetable.Register_Imported_ForceTable(str(tmp_path / "missing"), [[1]], 100, 100, 1e-8)


def test_fs_ackland_tables_match_reference_export(oracle):
    """FS_TYPE table generation (Create_Pairwise_ForceTable on the Ackland-Thetford W-W functions) against the reference's
    exported examples/use_ForceTableGen/EM_TB_WANGJUN_W-HE_2010.pair, table id 1, for the product and the oracle."""
    g = np.load(os.path.join(util.GOLD, "wangjun_fs_ww_pair_rows.npz"))
    rmax = 10.0e-8                                                       # the export's range: r_1 = 1e-7 A
    ergev = 1.0 / 1.60219e-12
    tp = forcetable.Create_Interaction_ForceTable(capi.LIB_ACKLAND_FS_W, [[1]], 10000, 10000, rmax, pot_type="FS_TYPE")
    to = oracle.Tables(oracle.LIB_ACKLAND_FS_W, [[1]], 10000, 10000, rmax, rmax=rmax, pot_type=oracle.POT_FS)
    sel = g["index"] - 1
    for t in (tp, to):
        r = (np.arange(1, 10001) / t.csi) ** 2 * 1e8
        assert np.allclose(r[sel], g["r"], rtol=PRINT_PAIR, atol=0)
        cols = np.stack([t.potr * 2 * ergev * 1e8, t.fpotr * ergev, t.potb * ergev * ergev, t.fpotb * ergev * ergev * 1e-8], axis=1)
        for c in range(4):
            assert np.all(np.abs(cols[sel, c] - g["pair"][:, c]) <= PRINT_PAIR * np.abs(g["pair"][:, c]) + 1e-30), c
    assert np.max(np.abs(tp.potr - to.potr)) <= 1e-13 * np.max(np.abs(to.potr))
    assert np.max(np.abs(tp.fpotb - to.fpotb)) <= 1e-13 * np.max(np.abs(to.fpotb))


@pytest.fixture(scope="module")
def lspt_dir(tmp_path_factory):
    import tarfile
    d = tmp_path_factory.mktemp("lspt")
    with tarfile.open(os.path.join(util.GOLD, "whhe_eam1_lspt_W.tar.xz"), "r:xz") as tf:
        tf.extractall(d, filter="data")
    return str(d)


def test_lspt_import_matches_reference_embd_and_oracle(lspt_dir):
    """The ".lspt" importer (Filedatas_Func_Lspt.F90) on the tungsten functions of examples/NIST_Potentials/WHHe_EAM1_LSPT:
    the embedding table equals the reference's exported WHHe_EAM1.lspt.embd (id 1, 9 digits) for the product and the
    NumPy/SciPy restatement, and all six tables of the two implementations agree far below print precision."""
    g = np.load(os.path.join(util.GOLD, "whhe_eam1_lspt_embd_rows.npz"))
    path = os.path.join(lspt_dir, "W_EAM1.lspt")
    info = forcetable.lspt_info(path)
    assert info["elements"] == ["W"] and abs(info["rhomx"] - 10.0) < 1e-12 and abs(info["cutoff"] - 5.4604375e-8) < 1e-20
    t = forcetable.NIST_Register_Interaction_Table(path, 10000, 10000)
    o = tables_np.lspt_tables_W(lspt_dir, 10000, 10000)
    ergev = 1.0 / 1.60219e-12
    sel = g["index"] - 1
    for fe, dfe, rhod in ((t.fembd, t.dfembd, t.rhod), (o["fembd"][0], o["dfembd"][0], o["rhod"])):
        assert np.allclose(np.arange(10000)[sel] * rhod, g["rho"], rtol=PRINT_EMBD, atol=1e-300)
        assert np.all(np.abs(fe[sel] * ergev - g["f"]) <= PRINT_EMBD * np.abs(g["f"]) + 1e-40)
        assert np.all(np.abs(dfe[sel] - g["df"]) <= PRINT_EMBD * np.abs(g["df"]) + 1e-40)
    for name in ("potr", "fpotr", "potb", "fpotb", "fembd", "dfembd"):
        a, b = getattr(t, name), o[name][0]
        assert np.max(np.abs(a - b)) <= 1e-11 * np.max(np.abs(b)), name
    assert t.csi == o["csi"] and t.rhod == o["rhod"]


def test_moldy_import_against_the_formulas(tmp_path):
    """".moldy" files (Filedatas_Func_Moldy.F90): no example ships with the reference, so the importer is checked against
    a direct NumPy evaluation of the documented functions (cubic-knot sums scaled by the lattice constant, F = -sqrt(rho))
    on the reference's table grid -- "parity unpinned" for this format."""
    a0 = 3.1652
    ak, rk = [0.96, -0.29, 0.41], [1.30, 1.20, 0.90]
    Ak, Rk = [0.70, 0.05], [1.35, 1.00]
    p = tmp_path / "W_test.moldy"
    p.write_text("W\n%s\n%s\n%s\n%s\n%.4f 183.84\n" % (" ".join("%.6f" % v for v in ak), " ".join("%.6f" % v for v in rk),
                                                          " ".join("%.6fD0" % v for v in Ak), " ".join("%.6f" % v for v in Rk), a0))
    ntab, nembd, rmax = 2000, 1500, 1.4 * a0 * 1e-8
    t = forcetable.Moldy_Register_Interaction_Table(str(p), ntab, nembd, rmax)
    ev, r = 1.60219e-12, (np.arange(1, ntab + 1) / t.csi) ** 2
    ra = r * 1e8
    H = lambda x: (x >= 0).astype(float)
    V = sum(a / a0 ** 3 * (k * a0 - ra) ** 3 * H(k * a0 - ra) for a, k in zip(ak, rk))
    dV = sum(a / a0 ** 3 * (k * a0 - ra) ** 2 * H(k * a0 - ra) for a, k in zip(ak, rk))
    rho = sum(a / a0 ** 6 * (k * a0 - ra) ** 3 * H(k * a0 - ra) for a, k in zip(Ak, Rk))
    drho = sum(a / a0 ** 6 * (k * a0 - ra) ** 2 * H(k * a0 - ra) for a, k in zip(Ak, Rk))
    close = lambda a, b: np.max(np.abs(a - b)) <= 1e-14 * np.max(np.abs(b))   # (the knot sums cancel near their zero crossings)
    assert close(t.potr, 0.5 * V * ev * r) and close(t.fpotr, 3.0 * dV * ev / 1e-8 * r)
    assert close(t.potb, rho) and close(t.fpotb, 3.0 * drho / 1e-8)
    assert abs(t.rhod - rho.max() * 20.0 / nembd) < 1e-15 * t.rhod
    g = np.arange(nembd) * t.rhod
    assert np.allclose(t.fembd[1:], -np.sqrt(g[1:]) * ev, rtol=1e-13) and t.fembd[0] == 0.0
    assert np.allclose(t.dfembd[1:], -0.5 / np.sqrt(g[1:]) * ev, rtol=1e-13)


REF_FS = "/root/reference/examples/use_ForceTableGen/EM_TB_WANGJUN_W-HE_2010"


@pytest.mark.skipif(not os.path.exists(REF_FS + ".pair"), reason="reads the reference tree (build container only)")
def test_import_of_a_reference_written_table_file():
    """Format compatibility of Import_ForceTable / Register_Imported_ForceTable with files the REFERENCE wrote: the 5-id
    FS_TYPE export examples/use_ForceTableGen/EM_TB_WANGJUN_W-HE_2010.pair/.embd (4 + 2 MB, read in place, not copied) is
    imported for PTYPE = 1 onto its own grid and must give back the Ackland-Thetford W tables this package generates, to the
    file's print precision; the FS embedding columns hold -sqrt(rho)."""
    ntab, rmax = 10000, 10.0e-8
    info_t = forcetable.Register_Imported_ForceTable(REF_FS, [[1]], ntab, 10000, rmax)
    assert info_t.PotType == "FS_TYPE" and info_t.nkind == 1 and list(info_t.kpair) == [1]
    gen = forcetable.Create_Interaction_ForceTable(capi.LIB_ACKLAND_FS_W, [[1]], ntab, 10000, rmax, pot_type="FS_TYPE")
    for name in ("potr", "fpotr", "potb", "fpotb"):
        a, b = getattr(info_t, name), getattr(gen, name)
        scale = np.max(np.abs(b))
        assert np.max(np.abs(a - b)) <= 2e-9 * scale, (name, np.max(np.abs(a - b)) / scale)
    # a two-group box that uses ids 1 (W-W), 4 (W<-He) , 3 (He<-W), 2 (He-He): kinds numbered in first-appearance order
    t2 = forcetable.Register_Imported_ForceTable(REF_FS, [[1, 4], [3, 2]], 2000, 2000, rmax)
    assert t2.nkind == 4 and list(t2.kpair) == [1, 3, 2, 4] and list(t2.kembd) == [1, 2]
    rho = np.arange(2000) * t2.rhod
    fe = t2.fembd.reshape(-1, t2.nkind1).T[0]
    assert np.allclose(fe[5:], -np.sqrt(rho[5:]), rtol=2e-5)      # FS export: F = -sqrt(RHO) (erg, RHO in erg^2), re-gridded


REF_CU = "/root/reference/examples/NIST_Potentials/Cu_EAM/Cu1.eam.fs.setfl"


@pytest.mark.skipif(not os.path.exists(REF_CU + ".pair"), reason="reads the reference tree (build container only)")
def test_import_of_the_reference_written_cu_tables(setfl):
    """The reference's own export of its setfl import (Cu1.eam.fs.setfl.pair/.embd, read in place) re-imported on the same
    grid equals this package's direct setfl import, to the file's print precision (away from the file's 1e12 -> O(10) step)."""
    t = forcetable.NIST_Register_Interaction_Table(setfl, 10000, 10000)
    back = forcetable.Register_Imported_ForceTable(REF_CU, [[1]], 10000, 10000, t.Rmax)
    assert back.PotType == "EAM_TYPE"
    r = (np.arange(1, 10001) / t.csi) ** 2 * 1e8
    far = r > 0.2
    for name in ("potr", "fpotr", "potb", "fpotb"):
        a, b = getattr(back, name)[far], getattr(t, name)[far]
        assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12 * np.max(np.abs(b)))) < 3e-9, name
    assert abs(back.rhod - t.rhod) < 1e-8 * t.rhod
    assert np.max(np.abs(back.fembd - t.fembd)) < 1e-8 * np.max(np.abs(t.fembd))
