"""Host mirror of type(MDForceTable) and of the potential libraries' Register_Interaction_Table
(reference: MDLIB/sor/Common/MD_TypeDef_ForceTable.F90:117-155, 1151-1230;
Potentials/EAM_WW_Marinica_JPCM25_2013/EAM_ForceTable_Marinica_JPCM25_2013.F90:17-83;
Potentials/EAM_WHeH_Bonny_JPCM26_2014/EAM_ForceTable_Bonny_JPCM26_2014.F90:17-140).

Table generation itself is native (mdb_host_ftable_create in csrc/host_potentials.cpp)."""
import ctypes as C
import os

import numpy as np

from . import capi

# &LIBNAME strings of the box file -> native library id
POTENTIAL_LIBS = {
    ("EAM_WW_MARINICA_JPCM25_2013", ""): capi.LIB_MARINICA_EAM2,
    ("EAM_WW_MARINICA_JPCM25_2013", "EAM2"): capi.LIB_MARINICA_EAM2,
    ("EAM_WHEH_BONNY_JPCM26_2014", ""): capi.LIB_BONNY_EAM1,
    ("EAM_WHEH_BONNY_JPCM26_2014", "EAM1"): capi.LIB_BONNY_EAM1,
}


class MDForceTable:
    """POTR, FPOTR, POTB, FPOTB(NKIND,NTAB); FEMBD, DFEMBD(NKIND1,NEMBD); KPAIR(NG,NG); KEMBD(NG);
    CSI = NTAB/sqrt(Rmax); RHOD.  Arrays are flat in Fortran (column-major) layout."""

    def __init__(self, pot_type="EAM_TYPE"):
        self.PotType = pot_type
        self.pot_type = capi.POT_EAM if pot_type == "EAM_TYPE" else capi.POT_FS

    def as_2d(self, name):
        nk = self.nkind1 if name in ("fembd", "dfembd") else self.nkind
        return getattr(self, name).reshape(-1, nk).T


def Create_Interaction_ForceTable(lib_id, ptype, ntab, nembd, rmax, rhoscal=20.0, pot_type="EAM_TYPE"):
    """ptype[i][j] = PTYPE(i+1,j+1): table id for 'density/pair at group i from group j'."""
    lib = capi.load()
    ptype = np.asarray(ptype, dtype=np.int32)
    ng = ptype.shape[0]
    t = MDForceTable(pot_type)
    pt = np.ascontiguousarray(ptype.T).ravel()
    nk = ng * ng
    t.potr, t.fpotr, t.potb, t.fpotb = (np.zeros(nk * ntab) for _ in range(4))
    t.fembd, t.dfembd = np.zeros(ng * nembd), np.zeros(ng * nembd)
    t.kpair = np.zeros(ng * ng, dtype=np.int32)
    t.kembd = np.zeros(ng, dtype=np.int32)
    nkind, nkind1, csi, rhod = C.c_int(), C.c_int(), C.c_double(), C.c_double()
    rc = lib.mdb_host_ftable_create(lib_id, ng, capi.ip(pt), int(ntab), int(nembd), float(rhoscal), float(rmax),
                                    C.byref(nkind), C.byref(nkind1), capi.ip(t.kpair), capi.ip(t.kembd),
                                    capi.dp(t.potr), capi.dp(t.fpotr), capi.dp(t.potb), capi.dp(t.fpotb),
                                    capi.dp(t.fembd), capi.dp(t.dfembd), C.byref(csi), C.byref(rhod))
    if rc != 0:
        raise capi.MDBError(rc, "mdb_host_ftable_create: potential id not registered in this library")
    t.ng, t.ntab, t.nembd = ng, int(ntab), int(nembd)
    t.nkind, t.nkind1, t.csi, t.rhod = nkind.value, nkind1.value, csi.value, rhod.value
    t.Rmax = float(rmax)
    for name in ("potr", "fpotr", "potb", "fpotb"):
        setattr(t, name, np.ascontiguousarray(getattr(t, name)[: t.nkind * ntab]))
    for name in ("fembd", "dfembd"):
        setattr(t, name, np.ascontiguousarray(getattr(t, name)[: t.nkind1 * nembd]))
    return t


def Register_Interaction_Table(SimBox, CtrlParam):
    """Reference entry point of every potential library: builds the MDForceTable for the box's
    &POTSUBCTL.  Table range Rmax = maxval(RU) as in the shipped source (MD_TypeDef_ForceTable.F90:591)."""
    key = (SimBox.PotLibname.upper(), SimBox.PotSubLibname.upper())
    if key not in POTENTIAL_LIBS:
        raise ValueError("potential library %r/%r is not available in this build" % key)
    return Create_Interaction_ForceTable(POTENTIAL_LIBS[key], SimBox.PTYPE, CtrlParam.NUMFTABR, CtrlParam.NUMFTABE,
                                         float(np.max(CtrlParam.RU)), CtrlParam.RHOSCAL, SimBox.PotType)


# ---- external tables (csrc/host_tables_io.cpp) ---------------------------------------------------------
def setfl_info(path):
    """Header of a NIST setfl file: elements, masses, cutoff [cm], RHOMX = Nrho*drho
    (Potentials/EAM_NIST/Filedatas_Func_Setfl.F90:153-205)."""
    lib = capi.load()
    ne, nrho, nr, cut, rhomx = C.c_int(), C.c_int(), C.c_int(), C.c_double(), C.c_double()
    names = C.create_string_buffer(16 * capi.MXGROUP)
    z = np.zeros(capi.MXGROUP, dtype=np.int32)
    mass, alat = np.zeros(capi.MXGROUP), np.zeros(capi.MXGROUP)
    rc = lib.mdb_host_setfl_info(os.fsencode(path), C.byref(ne), C.byref(nrho), C.byref(nr), C.byref(cut), C.byref(rhomx),
                                 names, 16, capi.ip(z), capi.dp(mass), capi.dp(alat))
    if rc != 0:
        raise capi.MDBError(rc, "mdb_host_setfl_info: cannot read %r" % (path,))
    n = ne.value
    el = [names.raw[16 * i:16 * i + 16].split(b"\0")[0].decode() for i in range(n)]
    return dict(elements=el, z=z[:n].copy(), mass=mass[:n].copy(), alat=alat[:n].copy(), nrho=nrho.value, nr=nr.value,
                cutoff=cut.value, rhomx=rhomx.value)


def lspt_info(path):
    lib = capi.load()
    ne, cut, rhomx = C.c_int(), C.c_double(), C.c_double()
    names = C.create_string_buffer(16 * capi.MXGROUP)
    rc = lib.mdb_host_lspt_info(os.fsencode(path), C.byref(ne), C.byref(cut), C.byref(rhomx), names, 16)
    if rc != 0:
        raise capi.MDBError(rc, "mdb_host_lspt_info: cannot read %r" % (path,))
    el = [names.raw[16 * i:16 * i + 16].split(b"\0")[0].decode() for i in range(ne.value)]
    return dict(elements=el, cutoff=cut.value, rhomx=rhomx.value)


def NIST_Register_Interaction_Table(path, ntab, nembd, ptype=None, rmax=0.0):
    """NIST_Register_Interaction_Table0 for a ".setfl" or ".lspt" library (Potentials/EAM_NIST/NIST_ForceTable.F90:100-137,
    332-398): tables for all NE*NE ids "I <- J" (id = (I-1)*NE + J, kind index = id), range and RHOMX from the
    file.  ptype defaults to the natural one, PTYPE(I,J) = (I-1)*NE + J."""
    lib = capi.load()
    is_lspt = str(path).lower().endswith(".lspt")   # Get_FileExtent dispatch, NIST_ForceTable.F90:111-125
    info = lspt_info(path) if is_lspt else setfl_info(path)
    ne = len(info["elements"])
    nk = ne * ne
    t = MDForceTable("EAM_TYPE")
    t.potr, t.fpotr, t.potb, t.fpotb = (np.zeros(nk * ntab) for _ in range(4))
    t.fembd, t.dfembd = np.zeros(nk * nembd), np.zeros(nk * nembd)
    nkind, csi, rhod, rmax_out = C.c_int(), C.c_double(), C.c_double(), C.c_double()
    fn = lib.mdb_host_lspt_ftable if is_lspt else lib.mdb_host_setfl_ftable
    rc = fn(os.fsencode(path), int(ntab), int(nembd), float(rmax), C.byref(nkind),
                                   capi.dp(t.potr), capi.dp(t.fpotr), capi.dp(t.potb), capi.dp(t.fpotb),
                                   capi.dp(t.fembd), capi.dp(t.dfembd), C.byref(csi), C.byref(rhod), C.byref(rmax_out))
    if rc != 0:
        raise capi.MDBError(rc, "mdb_host_setfl_ftable: cannot import %r" % (path,))
    if ptype is None:
        ptype = np.arange(1, nk + 1, dtype=np.int32).reshape(ne, ne)
    ptype = np.asarray(ptype, dtype=np.int32)
    t.ng = ptype.shape[0]
    t.kpair = np.ascontiguousarray(ptype.T).ravel().astype(np.int32)  # kind index = id
    t.kembd = np.array([ptype[i, i] for i in range(t.ng)], dtype=np.int32)
    t.ids = np.arange(1, nk + 1, dtype=np.int32)
    t.ids1 = t.ids.copy()
    t.ntab, t.nembd, t.nkind, t.nkind1 = int(ntab), int(nembd), nk, nk
    t.csi, t.rhod, t.Rmax, t.info = csi.value, rhod.value, rmax_out.value, info
    return t


def Moldy_Register_Interaction_Table(path, ntab, nembd, rmax, rhoscal=20.0):
    """Register_ForceTableProc_Moldy + Generate_NIST_ForceTalbe for a ".moldy" file (Filedatas_Func_Moldy.F90:21-153)."""
    lib = capi.load()
    t = MDForceTable("EAM_TYPE")
    t.potr, t.fpotr, t.potb, t.fpotb = (np.zeros(ntab) for _ in range(4))
    t.fembd, t.dfembd = np.zeros(nembd), np.zeros(nembd)
    csi, rhod = C.c_double(), C.c_double()
    rc = lib.mdb_host_moldy_ftable(os.fsencode(path), int(ntab), int(nembd), float(rhoscal), float(rmax), capi.dp(t.potr),
                                   capi.dp(t.fpotr), capi.dp(t.potb), capi.dp(t.fpotb), capi.dp(t.fembd), capi.dp(t.dfembd),
                                   C.byref(csi), C.byref(rhod))
    if rc != 0:
        raise capi.MDBError(rc, "mdb_host_moldy_ftable: cannot import %r" % (path,))
    t.ng, t.ntab, t.nembd, t.nkind, t.nkind1 = 1, int(ntab), int(nembd), 1, 1
    t.kpair, t.kembd = np.array([1], dtype=np.int32), np.array([1], dtype=np.int32)
    t.csi, t.rhod, t.Rmax = csi.value, rhod.value, float(rmax)
    return t


def Export_ForceTable(fname, t):
    """Common/MD_TypeDef_ForceTable.F90:1315-1459: writes fname.pair and fname.embd."""
    lib = capi.load()
    ids = np.asarray(getattr(t, "ids", np.arange(1, t.nkind + 1)), dtype=np.int32)
    ids1 = np.asarray(getattr(t, "ids1", np.arange(1, t.nkind1 + 1)), dtype=np.int32)
    rc = lib.mdb_host_ftable_export(os.fsencode(fname), t.pot_type, t.nkind, capi.ip(ids), t.ntab, t.csi, capi.dp(t.potr),
                                    capi.dp(t.fpotr), capi.dp(t.potb), capi.dp(t.fpotb), t.nkind1, capi.ip(ids1), t.nembd,
                                    t.rhod, capi.dp(t.fembd), capi.dp(t.dfembd))
    if rc != 0:
        raise capi.MDBError(rc, "mdb_host_ftable_export: cannot write %r" % (fname,))


def Register_Imported_ForceTable(fname, ptype, ntab, nembd, rmax):
    """Import_ForceTable + Register_Imported_ForceTable (Common/MD_TypeDef_ForceTable.F90:1461-1855): read
    fname.pair/.embd and re-grid the tables named by PTYPE onto the run's grid (rmax in cm)."""
    lib = capi.load()
    ptype = np.asarray(ptype, dtype=np.int32)
    ng = ptype.shape[0]
    pt = np.ascontiguousarray(ptype.T).ravel()
    nk = ng * ng
    potr, fpotr, potb, fpotb = (np.zeros(nk * ntab) for _ in range(4))
    fembd, dfembd = np.zeros(ng * nembd), np.zeros(ng * nembd)
    kpair, kembd = np.zeros(nk, dtype=np.int32), np.zeros(ng, dtype=np.int32)
    ptv, nkind, nkind1, csi, rhod = C.c_int(), C.c_int(), C.c_int(), C.c_double(), C.c_double()
    rc = lib.mdb_host_ftable_import(os.fsencode(fname), ng, capi.ip(pt), int(ntab), int(nembd), float(rmax), C.byref(ptv),
                                    C.byref(nkind), C.byref(nkind1), capi.ip(kpair), capi.ip(kembd), capi.dp(potr), capi.dp(fpotr),
                                    capi.dp(potb), capi.dp(fpotb), capi.dp(fembd), capi.dp(dfembd), C.byref(csi), C.byref(rhod))
    if rc != 0:
        raise capi.MDBError(rc, "mdb_host_ftable_import: cannot import %r (missing file or table id)" % (fname,))
    t = MDForceTable("FS_TYPE" if ptv.value == capi.POT_FS else "EAM_TYPE")
    t.ng, t.ntab, t.nembd, t.nkind, t.nkind1 = ng, int(ntab), int(nembd), nkind.value, nkind1.value
    t.csi, t.rhod, t.Rmax, t.kpair, t.kembd = csi.value, rhod.value, float(rmax), kpair, kembd
    t.potr, t.fpotr, t.potb, t.fpotb = (np.ascontiguousarray(a[: t.nkind * ntab]) for a in (potr, fpotr, potb, fpotb))
    t.fembd, t.dfembd = (np.ascontiguousarray(a[: t.nkind1 * nembd]) for a in (fembd, dfembd))
    return t
