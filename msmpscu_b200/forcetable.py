"""Host mirror of type(MDForceTable) and of the potential libraries' Register_Interaction_Table
(reference: MDLIB/sor/Common/MD_TypeDef_ForceTable.F90:117-155, 1151-1230;
Potentials/EAM_WW_Marinica_JPCM25_2013/EAM_ForceTable_Marinica_JPCM25_2013.F90:17-83;
Potentials/EAM_WHeH_Bonny_JPCM26_2014/EAM_ForceTable_Bonny_JPCM26_2014.F90:17-140).

Table generation itself is native (mdb_host_ftable_create in csrc/host_potentials.cpp)."""
import ctypes as C

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
