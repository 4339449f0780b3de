"""Host-side mirror of the reference's MDLIB interface for the hot path, over the C ABI.

Same names, argument meaning and error behaviour as the Fortran procedures the GMD / PARREP step loop
calls (paths relative to the reference root, MDLIB/sor/ prefix dropped):

    SimMDBox / SimMDCtrl                 Common/MD_TypeDef_SimBox.F90:43-132, MD_TypeDef_SimCtrlParam.F90:26-264
    Initialize_Globle_Variables_DEV      CommonGPU/MD_Globle_Variables_GPU.F90:585
    Register_ForceClass / MDForceClassGPU  CommonGPU/MD_ForceClass_Register_GPU.F90:248-259,363-442
    Init_Forcetable_Dev                  :483-559
    Initialize_NeighboreList_DEV, Cal_NeighBoreList_DEV, Copyout_NeighboreList_DEV   CommonGPU/MD_NeighborsList_GPU.F90
    CalForce_ForceClass, CalPTensor_ForceClass, CalEpot_ForceClass   MD_ForceClass_Register_GPU.F90:636-704
    Predictor_DEV, Correction_DEV, CalEKin_DEV    CommonGPU/MD_DiffScheme_GPU.F90:604,821,1000
    Do_ResetParam_DEV, Do_EPCForce_DEV   LocalTempCtrlMeths/MD_LocalTempMethod_GPU.F90:95-138
    CopyOut_SimBox_DEV                   CommonGPU/MD_SimBoxArray_GPU.F90:202-318
    For_One_Step                         Appshell/MD_Method_GenericMD_GPU.F90:496-659

The reference keeps one module-level device state per process; here that state is the `DeviceState`
object (one per GPU / rank).  Errors are exceptions carrying the ABI status code instead of `stop`.
"""
from dataclasses import dataclass, field

import numpy as np

from . import capi, forcetable
from .constants import (CP_A2CM, CP_AU2G, CP_CGS2KBAR, CP_ERGEV, CP_EVERG, CP_FS2S, CP_KB, CP_PS2S, CP_STATU_ACTIVE, CP_STATU_FIXPOS,
                        CP_STATU_FIXPOSX, CP_STATU_FIXPOSY, CP_STATU_FIXPOSZ)


# ----------------------------------------------------------------------------------------------
@dataclass
class SimMDBox:
    """The fields of type(SimMDBox) the hot path reads.  Arrays are (NPRT,3) / (NPRT,), CGS units."""
    NPRT: int = 0
    NGROUP: int = 1
    RR: float = 0.0                      # lattice unit in cm
    ZL: np.ndarray = None                # box size (3)
    BOXLOW: np.ndarray = None
    BOXUP: np.ndarray = None
    BOXSHAPE: np.ndarray = field(default_factory=lambda: np.eye(3))
    CM: np.ndarray = None                # mass per group (g)
    SYMB: list = field(default_factory=list)
    PTYPE: np.ndarray = None             # (NGROUP,NGROUP) interaction-table ids
    PotType: str = "EAM_TYPE"
    PotLibname: str = ""
    PotSubLibname: str = ""
    ITYP: np.ndarray = None
    XP: np.ndarray = None
    XP1: np.ndarray = None
    FP: np.ndarray = None
    DIS: np.ndarray = None
    EPOT: np.ndarray = None
    EKIN: np.ndarray = None
    STATU: np.ndarray = None
    VTENSOR: np.ndarray = field(default_factory=lambda: np.zeros((3, 3)))
    PROP: np.ndarray = None              # (NGROUP) per-group status bits (CP_STATU_FIXPOS...); None = none set

    def allocate(self):
        n = self.NPRT
        for name in ("XP", "XP1", "FP", "DIS"):
            if getattr(self, name) is None:
                setattr(self, name, np.zeros((n, 3)))
        for name in ("EPOT", "EKIN"):
            if getattr(self, name) is None:
                setattr(self, name, np.zeros(n))
        if self.STATU is None:
            self.STATU = np.full(n, CP_STATU_ACTIVE, dtype=np.int32)
        if self.ITYP is None:
            self.ITYP = np.ones(n, dtype=np.int32)


@dataclass
class TiCtrlParam:
    """Common/MD_TypeDef_EPCCtrl.F90:27-37 defaults."""
    TI: float = 0.0
    METH_EPC: int = 0
    EPC_HE: float = 100.0 * CP_EVERG
    EPC_ALPHA: float = 1.0 * CP_PS2S
    EPC_CUT: float = 0.1


@dataclass
class SimMDCtrl:
    """The fields of type(SimMDCtrl) the hot path reads (already converted to CGS)."""
    MULTIBOX: int = 1
    IFPD: np.ndarray = field(default_factory=lambda: np.ones(3, dtype=np.int32))
    RU: np.ndarray = None                # (NGROUP,NGROUP) force cut-offs, cm
    NB_RM: np.ndarray = None             # (NGROUP,NGROUP) list cut-offs, cm
    NB_MXNBS: int = 256
    NB_UPTAB: int = 10
    IT0: int = 1
    H: float = 0.5 * CP_FS2S
    # step-size and list-period schedules (Common/MD_TypeDef_SimCtrlParam.F90:41-47,120-122; defaults :886-890,940-942)
    IHDUP: int = 0                       # 0 fixed step; > 0 ramp HMI -> HMX every IHDUP steps; < 0 displacement-limited (DMX)
    HMI: float = 1.0 * CP_FS2S
    HMX: float = 1.0 * CP_FS2S
    DMX: float = 0.1e-8                  # cm (the control file gives Angstrom)
    NB_UPTABMI: int = 10
    NB_UPTABMX: int = 10
    NB_DBITAB: int = 100000
    NUMFTABR: int = 10000
    NUMFTABE: int = 10000
    RHOSCAL: float = 20.0
    LT_CTRL: list = field(default_factory=list)
    # quench / event detection (Common/MD_TypeDef_SimCtrlParam.F90:205-217; parsed at :1694-2060)
    SEED: list = field(default_factory=lambda: [43434])
    Quench_Steps: int = 1000             # MD_TypeDef_SimCtrlParam.F90:204,997
    Quench_Meth: str = "ST"              # "LBFGS" | "CG" | "ST" | "DYN"  (the low word of CtrlParam%Quench_Meth)
    Quench_LSearch: bool = False         # CP_DAMPSCHEME_LSEARCH ("CG-LS", "ST-LS")
    STEEPEST_Alpha: float = 0.1
    STEEPEST_MiStep: float = 1.0e-5      # LU
    STEEPEST_MxStep: float = 0.1         # LU
    STEEPEST_MiDelE: float = 0.001       # eV
    LBFGS_PGtol: float = 0.0
    LBFGS_Factr: float = 0.0
    LBFGS_MSave: int = 7
    STRCUT_DRTol: float = 0.03           # &DRTOL: displacement that counts as an event (LU); default :202,996
    DAMPTIME0: int = 0
    DAMPTIME1: int = 0
    # thermalisation schedule (Common/MD_TypeDef_SimCtrlParam.F90:89-91, defaults :918-920; &TEMPERATURE / &THERMALIZATION)
    TI: float = 0.0                      # K
    IVTIME0: int = 1
    IVTIME: int = 0                      # number of thermalisations in one section
    IVPAS: int = 50                      # steps between two thermalisations


class MDPSCUError(RuntimeError):
    """Raised where the reference prints 'MDPSCU Error: ...' and stops."""


# ----------------------------------------------------------------------------------------------
class MDForceClassGPU:
    """type(MDForceClassGPU): eight procedure slots filled by Register_ForceClass."""
    SLOTS = ("pIniForcetable", "pClrForcetable", "pCalForce", "pCalEpot0", "pCalEpot", "pCalEDen", "pCalPTensor",
             "pCalAVStress")

    def __init__(self):
        for s in self.SLOTS:
            setattr(self, s, None)
        self.ExtForces = []  # AddExtForce_ForceClass list (boost, springs): not on the hot path
        self.PotType = ""


class DeviceState:
    """The reference's module-level singletons (dm_WorkSpace, dm_Neighbors, dm_FTableWS) for one GPU."""

    def __init__(self, device=0):
        self.ctx = capi.Context(device)
        self.SimBox = None
        self.nbox = 1
        self.ForceClass = None
        self.FTable = None


gm_ForceClass = MDForceClassGPU()


def Initialize_Globle_Variables_DEV(dev: DeviceState, SimBox, CtrlParam: SimMDCtrl):
    """SimBox may be one box or a list of MULTIBOX identical-size boxes (concatenated on the device)."""
    boxes = SimBox if isinstance(SimBox, (list, tuple)) else [SimBox]
    b0 = boxes[0]
    for b in boxes:
        if b.NPRT != b0.NPRT:
            raise MDPSCUError("MDPSCU Error: boxes of one run must hold the same number of particles")
    dev.SimBox, dev.nbox = boxes, len(boxes)
    dev.ctx.box_set(len(boxes), b0.NPRT, b0.BOXLOW, b0.ZL, CtrlParam.IFPD, b0.CM, boxshape=b0.BOXSHAPE)
    CopyIn_SimBox_DEV(dev, boxes)


def CopyIn_SimBox_DEV(dev, boxes):
    cat = lambda name: np.concatenate([getattr(b, name) for b in boxes])
    dev.ctx.upload(capi.F_XP, cat("XP"))
    dev.ctx.upload(capi.F_XP1, cat("XP1"))
    dev.ctx.upload(capi.F_DIS, cat("DIS"))
    dev.ctx.upload(capi.F_FP, cat("FP"))
    dev.ctx.upload(capi.F_ITYP, cat("ITYP"))
    dev.ctx.upload(capi.F_STATU, cat("STATU"))


def CopyOut_SimBox_DEV(dev, boxes=None):
    """Un-permutes through GIDINV into the SimBox order (MD_SimBoxArray_GPU.F90:202-238)."""
    boxes = boxes or dev.SimBox
    n = boxes[0].NPRT
    for name, f in (("XP", capi.F_XP), ("XP1", capi.F_XP1), ("FP", capi.F_FP), ("DIS", capi.F_DIS), ("EPOT", capi.F_EPOT),
                    ("EKIN", capi.F_EKIN), ("STATU", capi.F_STATU)):
        a = dev.ctx.download(f, capi.ORDER_ORIGINAL)
        for i, b in enumerate(boxes):
            setattr(b, name, a[i * n:(i + 1) * n].copy())


def Register_ForceClass(pottype: str, ForceClass: MDForceClassGPU = gm_ForceClass):
    """Fills the slots for "EAM_TYPE" or "FS_TYPE" (MD_ForceClass_Register_GPU.F90:363-442)."""
    if pottype not in ("EAM_TYPE", "FS_TYPE"):
        raise MDPSCUError("MDPSCU Error: the potential type %s is not supported" % pottype)
    fc = ForceClass
    fc.PotType = pottype

    def ini(dev, SimBox, CtrlParam, FTable):
        dev.ctx.tables_set(FTable, float(np.max(CtrlParam.RU)) ** 2)
        dev.FTable = FTable

    fc.pIniForcetable = ini
    fc.pClrForcetable = lambda dev: dev.ctx.tables_clear()
    fc.pCalForce = lambda dev, SimBox, CtrlParam: dev.ctx.force(capi.FORCE)
    fc.pCalEpot0 = lambda dev, SimBox, CtrlParam: dev.ctx.force(capi.EPOT)

    def epot(dev, SimBox, CtrlParam):
        dev.ctx.force(capi.EPOT)
        return dev.ctx.download(capi.F_EPOT, capi.ORDER_ORIGINAL)

    fc.pCalEpot = epot
    fc.pCalEDen = lambda dev, SimBox, CtrlParam: dev.ctx.force(capi.DEN)

    def ptensor(dev, SimBox, CtrlParam):
        vt = dev.ctx.force(capi.FORCE | capi.VIRIAL)
        for b in (dev.SimBox or []):
            b.VTENSOR = vt.copy()
        return vt

    fc.pCalPTensor = ptensor
    fc.pCalAVStress = lambda dev, order=capi.ORDER_CELL: dev.ctx.atomic_stress(order)  # Cal_EAM_AtomicStressTensor_DEV
    return fc


def Init_Forcetable_Dev(dev, SimBox, CtrlParam, ForceClass=gm_ForceClass, FTable=None):
    if ForceClass.pIniForcetable is None:
        raise MDPSCUError("MDPSCU Error: force class is not registered")
    b0 = SimBox[0] if isinstance(SimBox, (list, tuple)) else SimBox
    FTable = FTable or forcetable.Register_Interaction_Table(b0, CtrlParam)
    ForceClass.pIniForcetable(dev, SimBox, CtrlParam, FTable)
    dev.ForceClass = ForceClass
    return FTable


def Initialize_NeighboreList_DEV(dev, SimBox, CtrlParam):
    dev.ctx.nlist_init(CtrlParam.NB_RM, CtrlParam.NB_MXNBS)


def Cal_NeighBoreList_DEV(dev, SimBox, CtrlParam):
    """Returns the number of atoms found out of the box (the reference warns / prompts)."""
    return dev.ctx.nlist_build()


def Copyout_NeighboreList_DEV(dev, order=capi.ORDER_ORIGINAL):
    return dev.ctx.nlist_copyout(order)


def AddExtForce_ForceClass(ForceClass, Tag, pForce, pEpot=None):
    """MD_ForceClass_Register_GPU.F90:273-277: external force procedures (boost, springs) cumulated after pCalForce.
    pForce(dev, SimBox, CtrlParam) adds to FP on the device (mdb_devptr gives it the arrays)."""
    ForceClass.ExtForces.append((Tag, pForce, pEpot))


def ClearExtForce_ForceClass(ForceClass):
    ForceClass.ExtForces.clear()


def CalForce_ForceClass(dev, SimBox, CtrlParam, ForceClass=gm_ForceClass):
    """Calforce_ForceClass0/1 (:636-662): pCalForce, then Cumulate_ExtForce over the registered list."""
    ForceClass.pCalForce(dev, SimBox, CtrlParam)
    for _tag, pforce, _pepot in ForceClass.ExtForces:
        pforce(dev, SimBox, CtrlParam)


def CalPTensor_ForceClass(dev, SimBox, CtrlParam, ForceClass=gm_ForceClass):
    return ForceClass.pCalPTensor(dev, SimBox, CtrlParam)


def CalEpot_ForceClass(dev, SimBox, CtrlParam, ForceClass=gm_ForceClass):
    return ForceClass.pCalEpot(dev, SimBox, CtrlParam)


def DO_LBFGSB_FORSTEPS_DEV(dev, SimBox, CtrlParam, ForceClass=gm_ForceClass, MXNUMSTEPS=1000):
    """CommonGPU/MD_LBFGSScheme_GPU.F90:177-388; returns IFLAG (0 finished, 1 out of steps).  Control values as the
    reference takes them: CtrlParam.LBFGS_MSave, LBFGS_Factr, LBFGS_PGtol (Common/MD_TypeDef_SimCtrlParam.F90:207-209)."""
    fl, _nfg, _nit = dev.ctx.lbfgs(MXNUMSTEPS, getattr(CtrlParam, "LBFGS_MSave", 7), getattr(CtrlParam, "LBFGS_Factr", 0.0),
                                   getattr(CtrlParam, "LBFGS_PGtol", 0.0))
    return fl


def Do_Damp(dev, SimBox, CtrlParam, ForceClass=gm_ForceClass):
    """Appshell/MD_Method_ParRep_GPU.F90:883-955 (the same dispatch sits in the ART / BST / TAD method classes): quench by
    the scheme named in the low word of CtrlParam.Quench_Meth, then the per-atom energies."""
    meth = getattr(CtrlParam, "Quench_Meth", "ST")
    steps = getattr(CtrlParam, "Quench_Steps", 1000)
    flags = capi.QUENCH_LSEARCH if getattr(CtrlParam, "Quench_LSearch", False) else 0
    if meth == "LBFGS":
        out = DO_LBFGSB_FORSTEPS_DEV(dev, SimBox, CtrlParam, ForceClass, steps)
    elif meth == "DYN":
        out = Do_DynDamp_Forsteps_DEV(dev, SimBox, CtrlParam, ForceClass, steps)
    elif meth == "CG":
        out = Do_CG_Forsteps_DEV(dev, SimBox, CtrlParam, ForceClass, steps, flags)
    else:
        out = Do_Steepest_Forsteps_DEV(dev, SimBox, CtrlParam, ForceClass, steps, flags)
    CalEpot_ForceClass(dev, SimBox, CtrlParam, ForceClass)
    return out


def Do_DynDamp_Forsteps_DEV(dev, SimBox, CtrlParam, ForceClass=gm_ForceClass, MXNUMSTEPS=1000):
    """CommonGPU/MD_DiffScheme_GPU.F90:1809-1860; returns (IFLAG, max energy change [eV])."""
    midele = getattr(CtrlParam, "STEEPEST_MiDelE", 0.001)
    fl, de = dev.ctx.dyndamp(MXNUMSTEPS, CtrlParam.H, midele * CP_EVERG)
    return fl, de / CP_EVERG


def Predictor_DEV(dev, ITIME, SimBox, CtrlParam):
    """MD_DiffScheme_GPU.F90:604-682: DAMPING while DAMPTIME0 <= ITIME <= DAMPTIME0 + DAMPTIME1 - 1 (:611-617), then the predictor."""
    d0, d1 = getattr(CtrlParam, "DAMPTIME0", 0), getattr(CtrlParam, "DAMPTIME1", 0)
    if d1 > 0 and d0 <= ITIME <= d0 + d1 - 1:
        dev.ctx.damping()
    dev.ctx.predict(CtrlParam.H)


def Correction_DEV(dev, ITIME, SimBox, CtrlParam):
    dev.ctx.correct(CtrlParam.H)


def CalEKin_DEV(dev, SimBox, CtrlParam):
    dev.ctx.ekin()
    return dev.ctx.download(capi.F_EKIN, capi.ORDER_ORIGINAL)


def Cal_GlobalT_DEV(dev, SimBox, CtrlParam):
    """MD_DiffScheme_GPU.F90:1042-1064: the average temperature of all boxes."""
    return dev.ctx.global_t()


def VelScaling_DEV(dev, SimBox, CtrlParam, DT):
    """MD_DiffScheme_GPU.F90:1390-1446: scale the velocities of every box to temperature DT (K)."""
    dev.ctx.vel_scaling(DT)


def CheckTimestep_DEV(dev, ITIME, SimBox, CtrlParam, TH, H2S2, DMX2):
    """MD_DiffScheme_GPU.F90:1214-1258: 1 if any atom would move more than sqrt(DMX2) with step TH."""
    return dev.ctx.check_timestep(TH, H2S2, DMX2)


def Do_ResetParam_DEV(dev, SimBox, CtrlParam):
    b0 = SimBox[0] if isinstance(SimBox, (list, tuple)) else SimBox
    lt = CtrlParam.LT_CTRL or [TiCtrlParam() for _ in range(b0.NGROUP)]
    dev.ctx.epc_set([t.METH_EPC for t in lt], [t.TI for t in lt], [t.EPC_ALPHA for t in lt], [t.EPC_CUT for t in lt],
                    [t.EPC_HE for t in lt])


def Do_EPCForce_DEV(dev, SimBox, CtrlParam):
    dev.ctx.epc_apply()


def Reorder_NeighBoreList_Nearest_Dev(dev, Nearest):
    """CommonGPU/MD_NeighborsList_GPU.F90:2016-2066: keep the Nearest closest neighbours, ordered by distance."""
    dev.ctx.nlist_reorder_nearest(Nearest)


def Cal_AtomicStressTensor_DEV(dev, order=capi.ORDER_ORIGINAL):
    """pCalAVStress of the force class (Cal_EAM_AtomicStressTensor_DEV, MD_EAM_ForceTable_GPU.F90:1973-1990): (N, 9)."""
    return dev.ctx.atomic_stress(order)


def For_One_Step(dev, ITIME, SimBox, CtrlParam, ForceClass=gm_ForceClass):
    """One GMD step with the reference's call sequence (each call is one C-ABI entry point)."""
    Predictor_DEV(dev, ITIME, SimBox, CtrlParam)
    if (ITIME - CtrlParam.IT0) % CtrlParam.NB_UPTAB == 0:
        Cal_NeighBoreList_DEV(dev, SimBox, CtrlParam)
    CalForce_ForceClass(dev, SimBox, CtrlParam, ForceClass)
    Do_EPCForce_DEV(dev, SimBox, CtrlParam)
    Correction_DEV(dev, ITIME, SimBox, CtrlParam)


def Do_Steepest_Forsteps_DEV(dev, SimBox, CtrlParam, ForceClass=gm_ForceClass, MXNUMSTEPS=1000, METH=0):
    """CommonGPU/MD_SteepestScheme_GPU.F90:263-290: quench on the current neighbour list.  Control values as the
    reference takes them from CtrlParam (defaults of Common/MD_TypeDef_SimCtrlParam.F90:213-217); returns
    (IFLAG, max move [LU], max energy change [eV])."""
    b0 = SimBox[0] if isinstance(SimBox, (list, tuple)) else SimBox
    alpha = getattr(CtrlParam, "STEEPEST_Alpha", 0.1)
    mxstep = getattr(CtrlParam, "STEEPEST_MxStep", 0.1)
    mistep = getattr(CtrlParam, "STEEPEST_MiStep", 1.0e-5)
    midele = getattr(CtrlParam, "STEEPEST_MiDelE", 0.001)
    fl, mm, de = dev.ctx.steepest(MXNUMSTEPS, alpha, mxstep * b0.RR, mistep * b0.RR, midele * CP_EVERG, METH)
    return fl, mm / b0.RR, de / CP_EVERG


_thermalize_draw = [0]


def Thermalizing_MC_DEV(dev, SimBox, CtrlParam, TI):
    """CommonGPU/MD_DiffScheme_GPU.F90:1746-1805.  The seed is CtrlParam.SEED[0] (the reference seeds its device
    generators from the same control value); every call uses a new draw number."""
    seed = int(getattr(CtrlParam, "SEED", [43434])[0]) & 0xFFFFFFFFFFFFFFFF
    dev.ctx.thermalize(TI, seed, _thermalize_draw[0])
    _thermalize_draw[0] += 1


def Do_CG_Forsteps_DEV(dev, SimBox, CtrlParam, ForceClass=gm_ForceClass, MXNUMSTEPS=1000, METH=0):
    """CommonGPU/MD_CGScheme_GPU.F90:280-296: conjugate-gradient quench (METH & CP_DAMPSCHEME_LSEARCH selects the
    line-search variant).  Returns (IFLAG, max energy change [eV])."""
    b0 = SimBox[0] if isinstance(SimBox, (list, tuple)) else SimBox
    mxstep = getattr(CtrlParam, "STEEPEST_MxStep", 0.1)
    mistep = getattr(CtrlParam, "STEEPEST_MiStep", 1.0e-5)
    midele = getattr(CtrlParam, "STEEPEST_MiDelE", 0.001)
    fl, de = dev.ctx.cg(MXNUMSTEPS, mxstep * b0.RR, mistep * b0.RR, midele * CP_EVERG, METH)
    return fl, de / CP_EVERG


def ResetXP1(dev, SimBox):
    """Appshell/MD_Method_GenericMD_GPU.F90 ResetXP1: velocities are zeroed after a quench."""
    b0 = SimBox[0] if isinstance(SimBox, (list, tuple)) else SimBox
    dev.ctx.upload(capi.F_XP1, np.zeros((dev.ctx.n, 3)))
    b0.XP1[:] = 0.0


def For_Steps(dev, ITIME0, nsteps, SimBox, CtrlParam):
    """The same sequence for nsteps steps inside the library: mdb_run for a fixed step and list period, else mdb_run_sched with
    the schedules of the time loop (Appshell/MD_Method_GenericMD_GPU.F90:351-361: &STEPSIZE flag / hmi / hmx / dmx,
    &UPDATEFRE min / max / doubling interval).  CtrlParam.H and CtrlParam.NB_UPTAB are left as the loop leaves them."""
    c = CtrlParam
    if c.IHDUP == 0 and c.NB_UPTABMI == c.NB_UPTABMX == c.NB_UPTAB:
        return dev.ctx.run(ITIME0, nsteps, c.IT0, c.NB_UPTAB, c.H)
    s = capi.Sched(c.IHDUP, c.HMI, c.HMX, c.DMX, c.NB_UPTABMI, c.NB_UPTABMX, c.NB_DBITAB)
    rc, c.H, _ = dev.ctx.run_sched(ITIME0, nsteps, c.IT0, s, c.H)
    last = ITIME0 + nsteps - 1
    c.NB_UPTAB = min(c.NB_UPTABMX, c.NB_UPTABMI * ((last - c.IT0 + 1) // c.NB_DBITAB + 1))
    return rc


def Cal_thermal_quantities(SimBox):
    """Cal_thermal_quantities_SimMDBox (Common/MD_TypeDef_SimBox.F90:5048-5170), the host routine the reference runs on output
    steps over the arrays copied out of the device: mass-weighted drift velocity of the active atoms of the groups whose
    PROP carries no FIXPOS bit (:5063-5090), kinetic tensor of the drift-free velocities (:5092-5123), VOLUME, PTENSOR =
    (KTENSOR + VTENSOR) / VOLUME (:5125-5131), TEMPERATURE from the trace of KTENSOR (:5133), SPRESS0/1 in kbar (:5150-5152),
    AVEPOT in eV over the active atoms (:5155-5156) and HARMIL in erg (:5159-5163).  Groups are addressed through ITYP (the
    reference's IPA ranges hold the atoms of one type).  As in the reference, row d of BOXSHAPE is applied to XP1(I,1:3) minus
    the d-th drift component (:5096-5101), and a fixed component contributes nothing (the reference leaves it unset).
    Also stores the results on SimBox, as the reference does."""
    n, ng = SimBox.NPRT, SimBox.NGROUP
    act = (SimBox.STATU & CP_STATU_ACTIVE) == CP_STATU_ACTIVE
    prop = np.asarray(getattr(SimBox, "PROP", None) if getattr(SimBox, "PROP", None) is not None else np.zeros(ng, dtype=np.int64))
    cm = np.asarray(SimBox.CM, dtype=np.float64)
    shape = np.asarray(SimBox.BOXSHAPE, dtype=np.float64)
    free_group = [(int(prop[k]) & CP_STATU_FIXPOS) == 0 for k in range(ng)]
    vv0, tcm, anprt = np.zeros(3), 0.0, 0
    for k in range(ng):
        if not free_group[k]:
            continue
        m = act & (SimBox.ITYP == k + 1)
        cnt = int(m.sum())
        vv0 += SimBox.XP1[m].sum(axis=0) * cm[k]
        tcm += cm[k] * cnt
        anprt += cnt
    vv0 = vv0 / tcm
    cxp1 = np.zeros((n, 3))
    for d, bit in enumerate((CP_STATU_FIXPOSX, CP_STATU_FIXPOSY, CP_STATU_FIXPOSZ)):
        m = act & ((SimBox.STATU & bit) == 0)
        cxp1[m, d] = ((SimBox.XP1[m] - vv0[d]) * shape[d]).sum(axis=1)
    ket = np.zeros((3, 3))
    for k in range(ng):
        if not free_group[k]:
            continue
        m = act & (SimBox.ITYP == k + 1)
        ket += cm[k] * (cxp1[m].T @ cxp1[m])
    b = shape
    vol = (b[0, 0] * (b[1, 1] * b[2, 2] - b[1, 2] * b[2, 1]) - b[1, 0] * (b[0, 1] * b[2, 2] - b[2, 1] * b[0, 2])
           + b[2, 0] * (b[0, 1] * b[1, 2] - b[1, 1] * b[0, 2]))
    volume = vol * SimBox.ZL[0] * SimBox.ZL[1] * SimBox.ZL[2]
    vt = np.asarray(SimBox.VTENSOR, dtype=np.float64)
    temp = (ket[0, 0] + ket[1, 1] + ket[2, 2]) / (3.0 * anprt * CP_KB)
    spress0 = (n * CP_KB * temp / volume) * CP_CGS2KBAR
    spress1 = ((vt[0, 0] + vt[1, 1] + vt[2, 2]) * (1.0 / 3.0) / volume) * CP_CGS2KBAR
    avepot = CP_ERGEV * SimBox.EPOT[act].sum() / int(act.sum())
    free = act & ((SimBox.STATU & CP_STATU_FIXPOS) == 0)
    harmil = avepot * CP_EVERG + SimBox.EKIN[free].sum() / anprt
    out = dict(TEMPERATURE=temp, TEMP=temp, KTENSOR=ket, VOLUME=volume, PTENSOR=(ket + vt) / volume, SPRESS0=spress0,
               SPRESS1=spress1, SPRESS=spress0 + spress1, AVEPOT=avepot, HARMIL=harmil)
    for key, val in out.items():
        if key != "TEMP":
            setattr(SimBox, key, val)
    return out


def Putout_Instance_Config_SimMDBox(Fhead, SimBox, Stamp=None):
    """Common/MD_TypeDef_SimBox.F90:2388-2630 -> inputs.write_config (the &BOXCFG18 file of one box)."""
    from . import inputs
    return inputs.write_config(Fhead, SimBox, Stamp)


def Putout_Instance_Thermal_Quantities_SimMDBox(ITIME, TIME, ISECT, FNAME, SimBox):
    """Common/MD_TypeDef_SimBox.F90:5172-5260 -> inputs.write_thermal_quantities (one line per call; ITIME = 0 starts the file)."""
    from . import inputs
    return inputs.write_thermal_quantities(FNAME, ITIME, TIME, ISECT, SimBox)


def Do_Compare(SimBoxIni, SimBox, CtrlParam, MASK=None):
    """Do_Compare (Appshell/MD_Method_ParRep_GPU.F90:1241-1297), the default structure comparison of the event detection
    (Do_ChangeDetect :1094-1167 calls it on the quenched replicas): Flag(IP) = 1 where atom I of replica IB sits further than
    &DRTOL from its place in SimBoxIni -- minimum image along the periodic axes, strict `>` on the squared distance, atoms
    with MASK <= 0 skipped.  STRCUT_DRTol is kept in LU on this side (inputs.read_ctrl_file), hence the RR.
    Returns Flag (NB*NPRT), replica-major; a replica with any flag set is a transition (NCB / IBT of Do_ChangeDetect)."""
    boxes = SimBox if isinstance(SimBox, (list, tuple)) else [SimBox]
    nprt = SimBoxIni.NPRT
    rc2 = (CtrlParam.STRCUT_DRTol * SimBoxIni.RR) ** 2
    box, hbox = np.asarray(SimBoxIni.ZL, dtype=np.float64), 0.5 * np.asarray(SimBoxIni.ZL, dtype=np.float64)
    per = np.asarray(CtrlParam.IFPD) > 0
    mask = np.ones(nprt, dtype=bool) if MASK is None else np.asarray(MASK) > 0
    flag = np.zeros(len(boxes) * nprt, dtype=np.int32)
    for ib, b in enumerate(boxes):
        sep = SimBoxIni.XP - b.XP
        wrap = per[None, :] & (np.abs(sep) > hbox[None, :])
        sep = np.where(wrap, sep - np.copysign(box[None, :], sep), sep)
        flag[ib * nprt:(ib + 1) * nprt] = ((sep * sep).sum(axis=1) > rc2) & mask
    return flag


def Do_ChangeDetect(dev, SimBoxIni, SimBox, CtrlParam, CtrlParamDamp=None, MASK=None, ForceClass=gm_ForceClass):
    """Do_ChangeDetect (Appshell/MD_Method_ParRep_GPU.F90:1094-1167) on the device: the replicas are saved in device memory
    (mdb_state_save) instead of host SwapBoxes, quenched (Do_Damp), compared with SimBoxIni by one kernel (mdb_compare =
    Do_Compare :1241-1297) and restored together with their list (mdb_state_restore = CopyIn_SimBox_DEV + Cal_NeighBoreList_DEV).
    Returns (IBT, NCB, flag per replica)."""
    nb = len(SimBox) if isinstance(SimBox, (list, tuple)) else int(getattr(dev.ctx, "nbox", 1))
    dev.ctx.state_save()
    Do_Damp(dev, SimBox, CtrlParamDamp or CtrlParam, ForceClass)
    fb, ibt, ncb = dev.ctx.compare(SimBoxIni.XP, CtrlParam.STRCUT_DRTol * SimBoxIni.RR, mask=MASK, nbox=nb)
    dev.ctx.state_restore()
    return ibt, ncb, fb


def Do_DePhase(dev, SimBoxIni, SimBox, CtrlParamDephase, CtrlParamDamp=None, MASK=None, ForceClass=gm_ForceClass, log=None):
    """Do_DePhase (Appshell/MD_Method_ParRep_GPU.F90:959-1090): every replica starts from the quenched configuration SimBoxIni
    (:996-998), the device box, tables and list are set up (:1001-1008), then (IVTIME+1)*IVPAS steps run with the velocities
    redrawn at the first IVTIME multiples of IVPAS (Thermalizing_MC_DEV at TI, :1024-1042) -- predictor, list every NB_UPTAB
    steps counted from ITIME = 1, force, corrector, no thermostat (:1045-1063).  The replicas are then quenched and compared
    with SimBoxIni on the device (mdb_state_save / Do_Damp / mdb_compare / mdb_state_restore instead of the host SwapSimBox
    copies, :1066-1076) and the ones that left the basin are dropped: the survivors are packed to the front of SimBox in
    replica order (:1078-1086).  Returns (m_curReplicas, flag per replica); SimBox[0:m_curReplicas] hold the dephased states."""
    c = CtrlParamDephase
    boxes = SimBox if isinstance(SimBox, (list, tuple)) else [SimBox]
    for b in boxes:                                     # Copy_SimMDBox(SimBoxIni, SimBox(IB))
        b.ITYP = np.array(SimBoxIni.ITYP, dtype=np.int32)
        b.XP = np.array(SimBoxIni.XP, dtype=np.float64)
        b.XP1 = b.DIS = b.FP = b.EPOT = b.EKIN = b.STATU = None
        b.allocate()
        if getattr(SimBoxIni, "STATU", None) is not None:
            b.STATU[:] = SimBoxIni.STATU
    Initialize_Globle_Variables_DEV(dev, boxes, c)
    Init_Forcetable_Dev(dev, boxes, c, ForceClass)
    Initialize_NeighboreList_DEV(dev, boxes, c)
    if Cal_NeighBoreList_DEV(dev, boxes, c) != 0:
        raise MDPSCUError("Do_DePhase: neighbour list failed")
    CalForce_ForceClass(dev, boxes, c, ForceClass)
    ivnum = 0
    for itime in range(1, (c.IVTIME + 1) * c.IVPAS + 1):
        if (itime - 1) % c.IVPAS == 0:
            if log is not None:                         # "temp. ... K at ... timsteps in DEPHAS" (:1030-1034)
                log(itime, Cal_GlobalT_DEV(dev, boxes, c))
            if ivnum < c.IVTIME:
                Thermalizing_MC_DEV(dev, boxes, c, c.TI)
                ivnum += 1
        Predictor_DEV(dev, itime, boxes, c)
        if itime % c.NB_UPTAB == 0:
            Cal_NeighBoreList_DEV(dev, boxes, c)
        CalForce_ForceClass(dev, boxes, c, ForceClass)
        Correction_DEV(dev, itime, boxes, c)
    _ibt, _ncb, fb = Do_ChangeDetect(dev, SimBoxIni, boxes, c, CtrlParamDamp or c, MASK, ForceClass)
    fb = np.asarray(fb)
    n = SimBoxIni.NPRT
    fields = {name: dev.ctx.download(f, capi.ORDER_ORIGINAL) for name, f in
              (("XP", capi.F_XP), ("XP1", capi.F_XP1), ("DIS", capi.F_DIS), ("FP", capi.F_FP))}
    return Pack_Replicas(boxes, fields, fb, n), fb


def Pack_Replicas(boxes, fields, flag_box, n):
    """The tail of Do_DePhase (Appshell/MD_Method_ParRep_GPU.F90:1078-1086): replicas whose flag is 0 are copied, in replica
    order, to the front of the box array (IBF0 runs over the survivors); returns m_curReplicas.  `fields` maps an array name
    of SimMDBox to the replica-major device download (nbox * n rows)."""
    ibf0 = 0
    for ib in range(len(boxes)):
        if flag_box[ib] == 0:
            for name, a in fields.items():
                getattr(boxes[ibf0], name)[:] = a[ib * n:(ib + 1) * n]
            ibf0 += 1
    return ibf0


def Transition_Replicas(Flag, NPRT):
    """The tail of Do_ChangeDetect (:1146-1156): NCB = number of replicas with any flagged atom, IBT = the last of them
    (1-based, 0 when none)."""
    nb = len(Flag) // NPRT
    hit = [ib + 1 for ib in range(nb) if np.any(Flag[ib * NPRT:(ib + 1) * NPRT] > 0)]
    return (hit[-1] if hit else 0), len(hit)
