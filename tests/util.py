"""Shared builders for test cases: the same seeded inputs go to the CPU oracle and to the CUDA path."""
import os

import numpy as np

from msmpscu_b200 import capi, forcetable, lattice
from msmpscu_b200.constants import CP_A2CM, CP_AU2G, CP_EVERG

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Case:
    pass


def bcc_case(ncell=(8, 8, 8), a0=3.1652, seed=12345, nbox=1, ru_lu=1.9, nb_fac=1.2, mxkvois=256, ntab=10000,
             temp=600.0, disp=0.02, ifpd=(1, 1, 1)):
    """bcc W, Marinica EAM2 -- the BASELINE config family (C1: ncell 16^3; C2: 80^3)."""
    c = Case()
    d = lattice.bcc_box(ncell, a0, seed, disp_lu=disp, temp_k=temp, nbox=nbox)
    c.__dict__.update(d)
    c.ifpd = list(ifpd)
    c.ng = 1
    c.ru = ru_lu * c.rr
    c.nb_rm = np.full((1, 1), nb_fac * c.ru)
    c.mxkvois = mxkvois
    c.ntab = c.nembd = ntab
    c.lib = capi.LIB_MARINICA_EAM2
    c.ptype = np.array([[1]])
    c.rmax = c.ru
    return c


def bcc_fs_case(ncell=(8, 8, 8), seed=77, nbox=1):
    """bcc W with the Finnis-Sinclair potential of Ackland & Thetford (FS_TYPE path, MD_FS_ForceTable_GPU.F90):
    id 1 of the reference's EM_TB_WANGJUN_W-HE_2010 library; RU = 1.4 a0 (rho reaches 4.400 A = 1.39 a0)."""
    c = bcc_case(ncell, seed=seed, nbox=nbox, ru_lu=1.4)
    c.lib = capi.LIB_ACKLAND_FS_W
    c.pot_type = "FS_TYPE"
    return c


def neb_case(tag="react", rmax_mode="RU"):
    """examples/NEB_Test: 2000 W + 1 H, Bonny EAM1 (the reference's own known-answer run)."""
    g = np.load(os.path.join(GOLD, "neb_gmd_%s.npz" % tag))
    c = Case()
    c.rr = 3.14 * CP_A2CM
    c.xp = g["pos"] * c.rr
    c.xp1 = np.zeros_like(c.xp)
    c.ityp = g["ityp"].astype(np.int32)
    c.statu = g["statu"].astype(np.int32)
    c.napb, c.nbox = 2001, 1
    c.zl = np.array([10.0, 10.0, 10.0]) * c.rr
    c.boxlow = -0.5 * c.zl
    c.ifpd = [1, 1, 1]
    c.ng = 2
    c.mass = np.array([183.84, 1.0]) * CP_AU2G
    c.ru = 1.9 * c.rr
    c.nb_rm = np.full((2, 2), 1.2 * c.ru)
    c.mxkvois = 256
    c.ntab = c.nembd = 10000
    c.lib = capi.LIB_BONNY_EAM1
    c.ptype = np.array([[1, 2], [4, 5]])
    c.rmax = c.ru if rmax_mode == "RU" else 1.2 * c.ru
    c.gold_force = g["force"]
    c.gold_pot = g["pot"]
    return c


def cu_setfl_path(tmpdir=None):
    """examples/NIST_Potentials/Cu_EAM/Cu1.eam.fs.setfl, unpacked from the committed xz fixture"""
    import lzma
    import tempfile
    d = tmpdir or tempfile.gettempdir()
    p = os.path.join(str(d), "Cu1.eam.fs.setfl")
    if not os.path.exists(p):
        with lzma.open(os.path.join(GOLD, "Cu1.eam.fs.setfl.xz"), "rb") as f, open(p + ".tmp%d" % os.getpid(), "wb") as g:
            g.write(f.read())
        os.replace(p + ".tmp%d" % os.getpid(), p)
    return p


def fcc_cu_case(ncell=(8, 8, 8), seed=4242, nbox=1, ntab=10000):
    """fcc Cu with the NIST setfl potential Cu1 (Mendelev 2008) imported like the reference's EAM_NIST library:
    table range = the file's cutoff (6 A), list cutoff 1.2 x."""
    c = Case()
    c.__dict__.update(lattice.fcc_box(ncell, 3.639087, seed, nbox=nbox))
    c.ifpd = [1, 1, 1]
    c.ng = 1
    c.ru = 6.0 * CP_A2CM
    c.nb_rm = np.full((1, 1), 1.2 * c.ru)
    c.mxkvois = 256
    c.ntab = c.nembd = ntab
    c.setfl = cu_setfl_path()
    c.ptype = np.array([[1]])
    c.rmax = c.ru
    return c


def parrep_case(nbox=3, seed=2024):
    """BASELINE configs[3] shape (examples/PARREP_Test/CtrlFile300K.dat): replicas of the 2000 W + 1 H box (Bonny EAM1) as
    MULTIBOX, list cutoff 1.6 x RU (about 258 neighbours), MAXNB 400.  Replicas differ by small seeded displacements."""
    c = neb_case("react")
    rng = np.random.default_rng(seed)
    xs = [c.xp + rng.uniform(-0.01, 0.01, size=c.xp.shape) * c.rr for _ in range(nbox)]
    c.xp = np.concatenate(xs)
    c.xp1 = np.zeros_like(c.xp)
    c.ityp = np.tile(c.ityp, nbox)
    c.statu = np.tile(c.statu, nbox)
    c.nbox = nbox
    c.nb_rm = np.full((2, 2), 1.6 * c.ru)
    c.mxkvois = 400
    return c


def product_tables(c):
    if getattr(c, "setfl", None):
        return forcetable.NIST_Register_Interaction_Table(c.setfl, c.ntab, c.nembd, c.ptype, rmax=c.rmax)
    return forcetable.Create_Interaction_ForceTable(c.lib, c.ptype, c.ntab, c.nembd, c.rmax,
                                                    pot_type=getattr(c, "pot_type", "EAM_TYPE"))


def oracle_tables(O, c):
    if getattr(c, "setfl", None):
        from oracle import tables_np
        t = tables_np.setfl_tables(open(c.setfl).read(), c.ntab, c.nembd, rmax=c.rmax)
        return O.Tables.from_arrays(c.ng, c.ptype, np.diag(c.ptype), c.ntab, c.nembd, t["csi"], t["rhod"], c.ru, t["potr"],
                                    t["fpotr"], t["potb"], t["fpotb"], t["fembd"], t["dfembd"])
    lib = {capi.LIB_MARINICA_EAM2: O.LIB_MARINICA_EAM2, capi.LIB_BONNY_EAM1: O.LIB_BONNY_EAM1,
           capi.LIB_ACKLAND_FS_W: O.LIB_ACKLAND_FS_W}[c.lib]
    return O.Tables(lib, c.ptype, c.ntab, c.nembd, c.ru, rmax=c.rmax,
                    pot_type=O.POT_FS if getattr(c, "pot_type", "EAM_TYPE") == "FS_TYPE" else O.POT_EAM)


def make_ctx(c, build=True, force_path=None):
    ctx = capi.Context(0)
    ctx.box_set(c.nbox, c.napb, c.boxlow, c.zl, c.ifpd, c.mass)
    ctx.upload(capi.F_XP, c.xp)
    ctx.upload(capi.F_XP1, c.xp1)
    ctx.upload(capi.F_ITYP, c.ityp)
    ctx.upload(capi.F_STATU, c.statu)
    ctx.tables_set(product_tables(c), c.ru * c.ru)
    if force_path is not None:
        ctx.set_option(capi.OPT_FORCE_PATH, force_path)
    ctx.nlist_init(c.nb_rm, c.mxkvois)
    if build:
        ctx.nlist_build()
    return ctx


def oracle_md(O, c):
    return O.MD(c.nbox, c.napb, c.xp, c.xp1, c.ityp, c.statu, c.mass, c.boxlow, c.zl, c.ifpd,
                np.ascontiguousarray(c.nb_rm.T).ravel(), c.mxkvois, oracle_tables(O, c))


def oracle_cpu_md(O, c, half=False, fast=False):
    """the reference's CPU path (original order, CPU list rule) on a case"""
    return O.CpuMD(c.xp, c.xp1, c.ityp, c.statu, c.mass, c.boxlow, c.zl, c.ifpd, np.ascontiguousarray(c.nb_rm.T).ravel(),
                   c.mxkvois, oracle_tables(O, c), half=half, fast=fast)


def atom_relerr(a, b, floor_frac=1e-3):
    """PER-ATOM relative error max_i |a_i - b_i| / max(|b_i|, floor), floor = floor_frac x the largest |b_i|.
    a, b: (N,) or (N,3) (vector norm per atom).  Stricter than relerr(): an atom with a small force / energy is
    measured against its own magnitude (down to the floor), not against the largest one in the box."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.ndim == 1:
        d, m = np.abs(a - b), np.abs(b)
    else:
        d, m = np.linalg.norm(a - b, axis=1), np.linalg.norm(b, axis=1)
    floor = floor_frac * max(float(np.max(m)), 1e-300)
    return float(np.max(d / np.maximum(m, floor)))


def relerr(a, b):
    """max |a-b| / max |b| : the parity metric for per-atom vectors"""
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
