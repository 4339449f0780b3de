"""Readers for the reference's keyword-driven input files, enough for the hot path:
box file (&BOXF ... &ENDBOX), control file (&CTLF ... &ENDCTLF), configurations (&BOXCFG14-18, &CFGXYZ or
bare rows).  Formats: DOC/BoxFile.readme, DOC/CtrlFile_GMD.readme; parsers being mirrored:
MDLIB/sor/Common/MD_TypeDef_SimBox.F90:377-795,1823-1979 and MD_TypeDef_SimCtrlParam.F90 / MD_SimCtrlParam_GMD.F90.
Unit conversion follows Check_Globle_Variables, MDLIB/sor/Common/MD_Gvar.F90:918-951:
  RR = a0[A]*1e-8, ZL = LATT*RR, BOXLOW = -ZL/2 unless &LOWB, CM = amu*1.66053e-24,
  H = fs*1e-15, RU = RU[LU]*RR, NB_RM = factor*RU."""
import re
from dataclasses import dataclass

import numpy as np

from .constants import CP_A2CM, CP_AU2G, CP_CM2A, CP_ERGEV, CP_EVERG, CP_FS2S, CP_S2PS, CP_STATU_ACTIVE
from .mdlib import SimMDBox, SimMDCtrl, TiCtrlParam

_NUM = re.compile(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eEdD][-+]?\d+)?")


def _lines(path):
    with open(path) as f:
        for raw in f:
            s = raw.split("!")[0].strip()
            if s:
                yield s


def _numbers(s):
    s = re.sub(r'"[^"]*"', " ", s)
    # numbers that are part of a word (e.g. "A-B=") are fine; identifiers with digits (#1) are stripped
    s = re.sub(r"#\s*\d+", " ", s)
    # words are dropped, but not the exponent letter of a number (1.0e-5, 1.D-5): a word starts after a non-digit
    s = re.sub(r"(?<![\d.])[A-Za-z_][A-Za-z_0-9]*", " ", s)
    return [float(t.replace("d", "e").replace("D", "e")) for t in _NUM.findall(s)]


def _strings(s):
    return re.findall(r'"([^"]*)"', s)


def _kw(s):
    m = re.match(r"&([A-Za-z_0-9:]+)", s)
    return m.group(1).upper() if m else ""


def read_box_file(path):
    """Returns a SimMDBox with sizes in CGS, masses in g, PTYPE filled from &TABLE rows."""
    box = SimMDBox()
    size = latt = None
    lowb = None
    groups, tables = [], []
    cur = None
    for s in _lines(path):
        k = _kw(s)
        if k == "SIZE":
            size = np.array(_numbers(s[5:])[:3])
        elif k == "LATT":
            latt = _numbers(s[5:])[0]
        elif k == "LOWB":
            lowb = np.array(_numbers(s[5:])[:3])
        elif k == "NGROUP":
            box.NGROUP = int(_numbers(s[7:])[0])
        elif k == "GROUPSUBCTL":
            cur = dict(natom=0, mass=0.0, symb="", stat=CP_STATU_ACTIVE)
            groups.append(cur)
        elif k == "NATOM":
            n = int(_numbers(s[6:])[0])
            if cur is not None and len(groups) and "done" not in cur:
                cur["natom"] = n
            else:
                box.NPRT = n
        elif k == "ATOMP" and cur is not None:
            nums = _numbers(s[6:])
            cur["symb"] = (_strings(s) or [""])[0]
            cur["mass"] = nums[-1]
        elif k == "ENDSUBCTL" and cur is not None:
            cur["done"] = True
            cur = None
        elif k == "TYPE":
            box.PotType = (_strings(s) or ["EAM_TYPE"])[0].upper()
        elif k == "LIBNAME":
            st = _strings(s)
            box.PotLibname = st[0] if st else ""
            box.PotSubLibname = st[1] if len(st) > 1 else ""
        elif k == "TABLE":
            tables.append([int(v) for v in _numbers(re.sub(r"&TABLE\s+\S+", "", s, count=1))])
    if box.NPRT == 0:
        box.NPRT = sum(g["natom"] for g in groups)
    box.NGROUP = box.NGROUP or len(groups)
    box.RR = latt * CP_A2CM
    box.ZL = size * box.RR
    box.BOXLOW = -0.5 * box.ZL if lowb is None else lowb * box.RR
    box.BOXUP = box.BOXLOW + box.ZL
    box.CM = np.array([g["mass"] for g in groups]) * CP_AU2G
    box.SYMB = [g["symb"] for g in groups]
    box.NA = [g["natom"] for g in groups]
    ng = box.NGROUP
    box.PTYPE = np.array([row[:ng] for row in tables[:ng]], dtype=np.int32) if tables else np.ones((ng, ng), np.int32)
    return box


def read_ctrl_file(path, box, section=1):
    """The common blocks plus ONE time section (&SECTSUBCTL #section, 1-based; default the first) of a &CTLF:GMD file
    -> SimMDCtrl (CGS).  The reference keeps one SimMDCtrl per section in a linked list (sectCtrlParam); later sections
    never overwrite the values of the one asked for.  read_ctrl_sections returns them all."""
    c = SimMDCtrl()
    ng = box.NGROUP
    ru_lu = np.full((ng, ng), 0.0)
    nb_fac = 1.2
    temp = 0.0
    epc = False
    sect = 0
    for s in _lines(path):
        k = _kw(s)
        if k == "SECTSUBCTL":
            sect += 1
            continue
        if sect not in (0, section):
            continue
        if k == "BOXS":
            c.MULTIBOX = int(_numbers(s[5:])[0])
        elif k == "CUTOFF" and "NEIGH" not in s.upper() and not np.any(ru_lu):
            v = _numbers(s[7:])
            ru_lu[:] = v[0]
            for i, x in enumerate(v[: ng * ng]):
                ru_lu[i // ng, i % ng] = x
        elif k == "CUTOFF":
            nb_fac = _numbers(s[7:])[0]
        elif k == "TABLESIZE":
            v = _numbers(s[10:])
            c.NUMFTABR = int(v[0])
            c.NUMFTABE = int(v[1]) if len(v) > 1 else int(v[0])
            if len(v) > 2:
                c.RHOSCAL = v[2]
        elif k == "TEMPERATURE":
            temp = _numbers(s[12:])[0]
        elif k in ("THERMALIZATION", "THERMAL"):                # MD_SimCtrlParam_GMD.F90:97-125 (PARREP reader :58-74 wants both numbers)
            v = _numbers(s[len(k) + 1:])
            if len(v) < 1:
                raise ValueError("MDPSCU Error: the number of thermalization circle should be set")
            c.IVTIME = int(v[0])
            if c.IVTIME > 0 and len(v) < 2:
                raise ValueError("MDPSCU Error: the timestep interval for thermalizing should be set")
            if len(v) >= 2:
                c.IVPAS = int(v[1])
            if c.IVTIME > 0 and c.IVPAS == 0:
                raise ValueError("MDPSCU Error: the interval for thermalizing cannot be zero")
        elif k == "E_P_COUPLE":
            epc = True
        elif k == "STEPSIZE":
            v = _numbers(s[9:])     # flag (IHDUP), hmi, hmx [fs], dmx [Angstrom]: MD_SimCtrlParam_GMD.F90:362-378, MD_Gvar.F90:944-947
            c.IHDUP = int(v[0])
            c.HMI, c.HMX = v[1] * CP_FS2S, (v[2] if len(v) > 2 else v[1]) * CP_FS2S
            c.DMX = max(v[3] if len(v) > 3 else 0.1, 1.0e-6) * 1.0e-8
            c.H = c.HMI               # "the time step starts from its minimum value"
        elif k == "PERIDIC":
            c.IFPD = np.array([int(x) for x in _numbers(s[8:])[:3]], dtype=np.int32)
        elif k == "MAXNB":
            c.NB_MXNBS = int(_numbers(s[6:])[0])
        elif k == "UPDATEFRE":
            v = [int(x) for x in _numbers(s[10:])]          # MD_TypeDef_SimCtrlParam.F90:1355-1372
            c.NB_UPTABMI = v[0]
            c.NB_UPTABMX = v[1] if len(v) > 1 else v[0]
            c.NB_DBITAB = v[2] if len(v) > 2 else 100000
            c.NB_UPTAB = c.NB_UPTABMI
        elif k == "RANDSEED":                                   # MD_TypeDef_SimCtrlParam.F90:1694
            c.SEED = [int(v) for v in _numbers(s[9:])] or c.SEED
        elif k in ("QUENCHSTEP", "QUICKDAMP", "QUICKDUMP", "QUENCH"):   # :2018-2060, MD_SimCtrlParam_GMD.F90:59
            v = _numbers(s[len(k) + 1:])
            if v:
                c.Quench_Steps = int(v[0])
            st = [t.upper() for t in _strings(s)]
            if st:
                m = st[0]
                c.Quench_LSearch = m.endswith("-LS")
                m = m[:-3] if c.Quench_LSearch else m
                c.Quench_Meth = {"LBFGS": "LBFGS", "CG": "CG", "ST": "ST", "DYN": "DYN", "DYNAMICS": "DYN"}.get(m, c.Quench_Meth)
        elif k in ("STEPBOUND", "STEPCOND"):                    # :1890-1903
            v = _numbers(s[len(k) + 1:])
            if len(v) == 1:
                c.STEEPEST_MxStep = v[0]
            elif len(v) >= 2:
                c.STEEPEST_MiStep, c.STEEPEST_MxStep = min(v[0], v[1]), max(v[0], v[1])
        elif k in ("DELTAPOT", "POTCOND"):                      # :1905-1912
            v = _numbers(s[len(k) + 1:])
            if v:
                c.STEEPEST_MiDelE = v[0]
        elif k == "ALPHA":                                      # :1876-1888
            v = _numbers(s[6:])
            if v:
                c.STEEPEST_Alpha = v[0]
        elif k == "PGTOL":                                      # :1747
            c.LBFGS_PGtol = (_numbers(s[6:]) or [0.0])[0]
        elif k == "FACTR":                                      # :1761
            c.LBFGS_Factr = (_numbers(s[6:]) or [10.0])[0]
        elif k == "MSAVE":                                      # :1775
            c.LBFGS_MSave = int((_numbers(s[6:]) or [7])[0])
        elif k == "DRTOL":                                      # :2000 (a bare &DRTOL keeps the default)
            v = _numbers(s[6:])
            if v:
                c.STRCUT_DRTol = v[0]
    c.RU = ru_lu * box.RR
    c.NB_RM = nb_fac * c.RU
    c.LT_CTRL = [TiCtrlParam(TI=temp, METH_EPC=1 if epc else 0) for _ in range(ng)]
    c.TEMP = c.TI = temp
    return c


def read_ctrl_sections(path, box):
    """one SimMDCtrl per &SECTSUBCTL of the file, in file order"""
    nsect = sum(1 for s in _lines(path) if _kw(s) == "SECTSUBCTL")
    return [read_ctrl_file(path, box, section=k) for k in range(1, max(nsect, 1) + 1)]


def read_config(path, box):
    """Rows 'type x y z [vx vy vz ...]' in lattice units (positions) -> fills box.ITYP / XP (/XP1 = 0).
    A &BOXCFG18 file (the reference's own output, Putout_Instance_Config_SimMDBox) is read through its column directives:
    velocities come back in cm/s (XP1*RR*CP_S2PS, Common/MD_SimBoxArray.F90:564-569), STATU, and -- for comparisons --
    FP (dyn), EPOT / EKIN (erg) and DIS (cm) with the inverse of the writer's conversions."""
    rows = []
    started = False
    header = False
    cols = {}
    with open(path) as f:
        for raw in f:
            s = raw.split("!")[0].strip()
            if not s:
                continue
            if s.startswith("&"):
                header = True
                key = s.split()[0].upper()
                if key == "&TYPE":
                    started = True
                elif key.endswith("COL") and not started:
                    p = s.split()
                    cols[key[1:-3]] = (int(p[1]) - 1, int(p[2]))
                continue
            if header and not started:
                continue
            p = s.split()
            try:
                rows.append([float(x.replace("D", "e").replace("d", "e")) for x in (p if cols else p[:4])])
            except ValueError:
                continue
    a = np.array(rows)
    if a.shape[0] != box.NPRT:
        raise ValueError("configuration holds %d atoms, the box file says %d" % (a.shape[0], box.NPRT))
    box.ITYP = a[:, 0].astype(np.int32)
    box.XP = a[:, 1:4] * box.RR
    box.allocate()
    get = lambda k: a[:, cols[k][0]:cols[k][0] + cols[k][1]] if k in cols and a.shape[1] >= cols[k][0] + cols[k][1] else None
    if get("VEL") is not None:
        box.XP1 = get("VEL") * box.RR * CP_S2PS
    if get("STAT") is not None:
        box.STATU = get("STAT")[:, 0].astype(np.int32)
    if get("FP") is not None:
        box.FP = get("FP") * CP_EVERG / box.RR
    if get("EPOT") is not None:
        box.EPOT = -get("EPOT")[:, 0] * CP_EVERG
    if get("EKIN") is not None:
        box.EKIN = get("EKIN")[:, 0] * CP_EVERG
    if get("DIS") is not None:
        box.DIS = get("DIS") * box.RR
    return box


@dataclass
class MDRecordStamp:
    """type(MDRecordStamp), Common/MD_TypeDef_RecordStamp.F90:12-31, with its defaults."""
    AppType: str = ""
    ITest: int = -1
    IBox: tuple = (-1, -1)
    ITime: int = -1
    ISect: int = -1
    Time: float = -1.0
    ScalTime: float = -1.0
    InstantTemp: float = -1.0
    ICfg: tuple = (-1, -1)
    IRec: tuple = (-1, -1)


def _e(x, w, d):
    """Fortran 1PEw.d"""
    return "%*.*E" % (w, d, x)


def write_config(fhead, box, stamp=None, date=None):
    """Putout_Instance_Config_SimMDBox (Common/MD_TypeDef_SimBox.F90:2388-2630) without extra data pads: the &BOXCFG18 file
    of one box -- record stamp (Putout_RecordStamp, MD_TypeDef_RecordStamp.F90:89-115), column directives, box lines and one
    row per atom `(I8,3X,3(1PE17.8,1X),3(1PE17.8,1X),I6,2X,3(1PE17.8,1X),1PE17.8,1X,1PE17.8,1X,3(1PE17.8,1X))` holding
    XP/RR, XP1/RR/CP_S2PS, STATU, FP*CP_ERGEV*RR, -EPOT*CP_ERGEV, EKIN*CP_ERGEV, DIS/RR (:2594-2600).  The file name is
    Fhead.NNNN when Stamp%ICfg(1) >= 0 (:2424-2428).  Returns the name."""
    import time as _time
    st = stamp if stamp is not None else MDRecordStamp()
    fname = "%s.%04d" % (fhead, st.ICfg[0]) if st.ICfg[0] >= 0 else fhead
    date = date if date is not None else _time.strftime("%Y-%m-%d,%Hh%Mm%Ss")
    rr = box.RR
    out = ["&BOXCFG18",
           '&APPTYPE        "%s" ' % st.AppType,
           "&DATE            " + date,
           "&TESTID          %12d" % st.ITest,
           "&BOXID           %12d, %6d" % tuple(st.IBox),
           "&CFGID           %12d, %6d" % tuple(st.ICfg),
           "&RECID           %12d, %6d" % tuple(st.IRec),
           "&TIMESTEPS       %12d" % st.ITime,
           "&TIMESECT #      %12d" % st.ISect,
           "&TIME (ps)       " + _e(st.Time, 12, 5),
           "&SCALTIME (ps)   " + _e(st.ScalTime, 12, 5),
           "&TEMPSET (K)     " + _e(st.InstantTemp, 12, 5),
           " ",
           "!--- Table colume information:"]
    ncol = 1
    for tag, width, kind in (("&TYPECOL", 1, "I"), ("&XYZCOL", 3, "D"), ("&VELCOL", 3, "D"), ("&STATCOL", 1, "I"),
                             ("&FPCOL", 3, "D"), ("&EPOTCOL", 1, "D"), ("&EKINCOL", 1, "D"), ("&DISCOL", 3, "D")):
        out.append("%-12s %4d %3d \"%s\"" % (tag, ncol, width, kind))
        ncol += width
    na = [int((box.ITYP == k + 1).sum()) for k in range(box.NGROUP)]
    e3 = lambda v: " ".join(_e(x, 12, 5) for x in v)      # a trailing 1X writes nothing at the end of a Fortran record
    out += ["",
            "&LATT     lattice length (in A):       " + e3([rr * CP_CM2A]),
            "&BOXLOW   low boundary of box (in LU): " + e3(np.asarray(box.BOXLOW) / rr),
            "&BOXSIZE  boxsize (in LU):             " + e3(np.asarray(box.ZL) / rr),
            "&NATOM    total number of atoms:       %8d" % box.NPRT,
            "&NGROUP   number of group of atoms:    %8d" % box.NGROUP,
            "    &NA   number of atoms in groups:   " + " ".join("%8d" % v for v in na),
            "&TEMPCAL instant temperature(K):      " + e3([float(getattr(box, "TEMPERATURE", 0.0))]),
            "!--- Configure:",
            ("&TYPE         " "POS(LU)  (x)             (y)              (z)         "
             "VEL(LU/ps)(vx)           (vy)             (vz)        " "STATU   "
             "FOR(ev/LU)(fx)           (fy)             (fz)        " "POT(ev).          " "K.E.(ev)          "
             "DISPLACE(dx )            (dy)               (dz)       ").rstrip()]
    e17 = lambda v: "".join(_e(x, 17, 8) + " " for x in v)
    xp, v = box.XP / rr, box.XP1 / rr / CP_S2PS
    f, dis = box.FP * CP_ERGEV * rr, box.DIS / rr
    pot, ke = -box.EPOT * CP_ERGEV, box.EKIN * CP_ERGEV
    for k in range(box.NPRT):
        out.append("%8d   " % box.ITYP[k] + e17(xp[k]) + e17(v[k]) + "%6d  " % box.STATU[k] + e17(f[k]) + e17([pot[k]])
                   + e17([ke[k]]) + e17(dis[k]))
    with open(fname, "w") as fh:
        fh.write("\n".join(out) + "\n")
    return fname


_THERMAL_TITLES = ("  TIMESECTION   ", "    TIME(PS)    ", "    TEMP.(K)    ", "  VOULUME(LU)   ", "   PRESS0(kb)   ", "   PRESS1(kb)   ",
                   "   PRESST(kb)   ", "    C.E.(ev)    ", "  HARMILT.(cgs) ")


def write_thermal_quantities(fname, itime, time_ps, isect, box):
    """Putout_Instance_Thermal_Quantities_SimMDBox (Common/MD_TypeDef_SimBox.F90:5172-5260): ITIME = 0 starts the file with
    the title line `(20x,10(A16))`, later calls append; one line `(1x,I8,4x,I4,8x,11(1pE14.5,2x))` per call with TIME,
    TEMPERATURE, VOLUME/RR**3, SPRESS0, SPRESS1, SPRESS, AVEPOT, HARMIL as Cal_thermal_quantities left them on the box."""
    vals = (time_ps, box.TEMPERATURE, box.VOLUME / box.RR ** 3, box.SPRESS0, box.SPRESS1, box.SPRESS, box.AVEPOT, box.HARMIL)
    line = " %8d    %4d        " % (itime, isect) + "  ".join(_e(float(v), 14, 5) for v in vals)
    with open(fname, "w" if itime == 0 else "a") as fh:
        if itime == 0:
            fh.write((" " * 20 + "".join(_THERMAL_TITLES)).rstrip() + "\n")
        fh.write(line + "\n")
    return line


def read_stopping_table(path):
    """An external stopping-cross-section table in the reference's `&MDPSCU_STPTAB.stp` format (written by Export_STPTable /
    Stop_Srim, read by Load_STPTable, Common/MD_TypeDef_StpRangTable.F90:662-825; e.g. examples/Cascade_Test/Stopping_table.stp):
    `&NUMTABLE nt`, `&NUMPOINT ne`, one line `&A->B with COL# k Z1 M1 Z2 M2` per table, then rows `E S_2 ... S_{nt+1}` with the
    energy in keV and the cross sections in keV cm^2.  Returns (E [erg], {"A->B": S [erg cm^2]}, {"A->B": (Z1, M1, Z2, M2)})."""
    kev = 1000.0 * CP_EVERG
    cols, ids = {}, {}
    rows = []
    with open(path) as f:
        first = True
        for raw in f:
            t = raw.strip()
            if not t or t.startswith("!"):
                continue
            if t.startswith("&"):
                key = t.split()[0].upper()
                if first:
                    if key != "&MDPSCU_STPTAB.STP":
                        raise ValueError("unknown format for external stop-table: the header keyword should be &MDPSCU_STPTAB.stp")
                    first = False
                    continue
                if key in ("&NUMTABLE", "&NUMPOINT"):
                    continue
                v = [float(x.replace("D", "E").replace("d", "e")) for x in re.findall(r"[-+]?\d+\.?\d*(?:[eEdD][-+]?\d+)?", t[len(t.split()[0]):])]
                if "->" in key and len(v) >= 5:                       # (Extract_Numb(STR, 5, ...): COL, Z1, M1, Z2, M2)
                    cols[key[1:]] = int(v[0])
                    ids[key[1:]] = tuple(v[1:5])
                continue
            first = False
            rows.append([float(x) for x in t.split()])
    a = np.asarray(rows, dtype=np.float64)
    if a.ndim != 2 or not cols:
        raise ValueError("no stopping tables in %s" % path)
    return a[:, 0] * kev, {k: a[:, c - 1] * kev for k, c in cols.items()}, ids


def stopping_tables_for(path, symbols, emin_ev, emax_ev, ntab):
    """What Import_STPTable (Common/MD_TypeDef_StpRangTable.F90:830-907) hands to Initialize_STMOD_DEV for a box whose groups carry
    the element `symbols`: ETAB(0:ntab) uniform between &EMIN and &EMAX, STAB(:, K) per pair table, KPAIR(NG, NG) 1-based -- the
    arguments of mdb_stopping_set.  (The reference re-grids with its spline library; here linearly, the tables are smooth.)"""
    e, tabs, _ = read_stopping_table(path)
    ng = len(symbols)
    emin, emax = emin_ev * CP_EVERG, emax_ev * CP_EVERG
    if emax > e.max() or emin < e.min():
        raise ValueError("&EMIN / &EMAX outside the energy range of the stopping data")
    etab = np.linspace(emin, emax, ntab + 1)
    keys, kpair = [], np.zeros((ng, ng), dtype=np.int32)
    for i in range(ng):
        for j in range(ng):
            k = ("%s->%s" % (symbols[i], symbols[j])).upper()
            if k not in tabs:
                raise ValueError("cannot find stopping cross section for " + k)
            if k not in keys:
                keys.append(k)
            kpair[i, j] = keys.index(k) + 1
    stab = np.stack([np.interp(etab, e, tabs[k]) for k in keys], axis=1)
    return etab, stab, kpair
