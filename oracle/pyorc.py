"""ctypes binding of the CPU oracle (oracle/liborc.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(msmpscu_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LIB_MARINICA_EAM2 = 1
LIB_BONNY_EAM1 = 2
POT_EAM = 0
POT_FS = 1

A2CM = 1.0e-8
AU2G = 1.66053e-24
KB = 1.38054e-16
EVERG = 1.60219e-12

STATU_ACTIVE = 1
STATU_OUTOFBOX = 65536

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_md_create.restype = C.c_void_p
        _LIB.orc_md_kvois.restype = c_ip
        _LIB.orc_md_indi.restype = c_ip
    return _LIB


def _d(a):
    return a.ctypes.data_as(c_dp)


def _i(a):
    return a.ctypes.data_as(c_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _colmajor(a):
    """(N,3) array -> flat column-major copy x[0:N], y, z (Fortran layout of XP(N,3))."""
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).T).ravel()


def _from_colmajor(flat, n, ncol=3):
    return np.ascontiguousarray(flat.reshape(ncol, n).T)


class Tables:
    """Mirror of MDForceTable (MD_TypeDef_ForceTable.F90:117-155), Fortran layout T(NKIND,NTAB)."""

    def __init__(self, lib_id, ptype, ntab, nembd, ru_max, rmax=None, rhoscal=20.0, pot_type=POT_EAM):
        ptype = np.asarray(ptype, dtype=np.int32)
        ng = ptype.shape[0]
        self.ng = ng
        self.ntab, self.nembd = int(ntab), int(nembd)
        self.pot_type = pot_type
        self.ptype_f = np.ascontiguousarray(ptype.T).ravel()  # (i,j) at i + ng*j
        rmax = ru_max if rmax is None else rmax
        self.rmax = rmax
        self.ru2max = ru_max * ru_max
        nk = ng * ng
        self.potr = np.zeros(nk * ntab)
        self.fpotr = np.zeros(nk * ntab)
        self.potb = np.zeros(nk * ntab)
        self.fpotb = np.zeros(nk * ntab)
        self.fembd = np.zeros(ng * nembd)
        self.dfembd = np.zeros(ng * nembd)
        self.kpair = np.zeros(ng * ng, dtype=np.int32)
        self.kembd = np.zeros(ng, dtype=np.int32)
        nkind, nkind1 = C.c_int(), C.c_int()
        csi, rhod = C.c_double(), C.c_double()
        rc = lib().orc_ftable_build(
            C.c_int(lib_id), C.c_int(ng), _i(self.ptype_f), C.c_int(ntab), C.c_int(nembd),
            C.c_double(rhoscal), C.c_double(rmax), C.byref(nkind), C.byref(nkind1),
            _i(self.kpair), _i(self.kembd), _d(self.potr), _d(self.fpotr), _d(self.potb), _d(self.fpotb),
            _d(self.fembd), _d(self.dfembd), C.byref(csi), C.byref(rhod))
        if rc != 0:
            raise RuntimeError("orc_ftable_build failed: %d" % rc)
        self.nkind, self.nkind1 = nkind.value, nkind1.value
        self.csi, self.rhod = csi.value, rhod.value
        for name in ("potr", "fpotr", "potb", "fpotb"):
            setattr(self, name, np.ascontiguousarray(getattr(self, name)[: self.nkind * ntab]))
        for name in ("fembd", "dfembd"):
            setattr(self, name, np.ascontiguousarray(getattr(self, name)[: self.nkind1 * nembd]))

    def table(self, name):
        """Return T as (NKIND, NTAB) numpy view (row = kind)."""
        nk = self.nkind1 if name in ("fembd", "dfembd") else self.nkind
        return getattr(self, name).reshape(-1, nk).T

    def cstruct(self):
        class T(C.Structure):
            _fields_ = [("pot_type", C.c_int), ("ng", C.c_int), ("nkind", C.c_int), ("ntab", C.c_int),
                        ("nkind1", C.c_int), ("nembd", C.c_int), ("csi", C.c_double), ("rhod", C.c_double),
                        ("ru2max", C.c_double), ("kpair", c_ip), ("kembd", c_ip), ("potr", c_dp),
                        ("fpotr", c_dp), ("potb", c_dp), ("fpotb", c_dp), ("fembd", c_dp), ("dfembd", c_dp)]

        return T(self.pot_type, self.ng, self.nkind, self.ntab, self.nkind1, self.nembd, self.csi, self.rhod,
                 self.ru2max, _i(self.kpair), _i(self.kembd), _d(self.potr), _d(self.fpotr), _d(self.potb),
                 _d(self.fpotb), _d(self.fembd), _d(self.dfembd))


def ncell(zl, nb_rm_max):
    out = (C.c_int * 3)()
    lib().orc_ncell(_d(_f64(zl)), C.c_double(nb_rm_max), out)
    return list(out)


def _vec3(a):
    return (C.c_double * 3)(*[float(x) for x in a])


def _ivec3(a):
    return (C.c_int * 3)(*[int(x) for x in a])


_IDENT = (C.c_double * 9)(1, 0, 0, 0, 1, 0, 0, 0, 1)


def nlist_build_dev(nbox, napb, xp, ityp, statu, boxlow, zl, ifpd, nb_rm, mxkvois):
    """Device rule.  xp (N,3) ORIGINAL order.  Returns dict with sorted-order list."""
    n = nbox * napb
    nb_rm = _f64(nb_rm)
    ng = int(round(np.sqrt(nb_rm.size)))
    nc3 = ncell(zl, nb_rm.max())
    nc = nc3[0] * nc3[1] * nc3[2] * nbox
    xpf = _colmajor(xp)
    ityp = _i32(ityp)
    statu = _i32(statu).copy()
    inc = np.zeros(n, np.int32)
    gid = np.zeros(n, np.int32)
    nac = np.zeros(nc, np.int32)
    naac = np.zeros(nc, np.int32)
    ia1th = np.zeros(nc, np.int32)
    kvois = np.zeros(n, np.int32)
    indi = np.zeros(n * mxkvois, np.int32)
    nnmax = C.c_int()
    ncell_out = (C.c_int * 3)()
    nout = lib().orc_nlist_build_dev(
        C.c_int(nbox), C.c_int(napb), _d(xpf), _i(ityp), _i(statu), _vec3(boxlow), _vec3(zl), _ivec3(ifpd),
        _IDENT, C.c_int(ng), _d(nb_rm), C.c_int(mxkvois), ncell_out, _i(inc), _i(gid), _i(nac), _i(naac),
        _i(ia1th), _i(kvois), _i(indi), C.byref(nnmax))
    return dict(ncell=list(ncell_out), inc=inc, gid=gid, nac=nac, naac=naac, ia1th=ia1th, kvois=kvois,
                indi=indi.reshape(mxkvois, n), nn_max=nnmax.value, nout=nout, statu=statu)


def nlist_build_cpu(xp, ityp, statu, boxlow, zl, ifpd, nb_rm, mxkvois):
    xp = np.asarray(xp)
    n = xp.shape[0]
    nb_rm = _f64(nb_rm)
    ng = int(round(np.sqrt(nb_rm.size)))
    xpf = _colmajor(xp)
    kvois = np.zeros(n, np.int32)
    indi = np.zeros(n * mxkvois, np.int32)
    rc = lib().orc_nlist_build_cpu(C.c_int(n), _d(xpf), _i(_i32(ityp)), _i(_i32(statu)), _vec3(boxlow), _vec3(zl),
                                   _ivec3(ifpd), _IDENT, C.c_int(ng), _d(nb_rm), C.c_int(mxkvois), _i(kvois),
                                   _i(indi))
    if rc != 0:
        raise RuntimeError("orc_nlist_build_cpu: %d" % rc)
    return kvois, indi.reshape(mxkvois, n)


def force(xp, ityp, statu, kvois, indi, zl, ifpd, tables, virial=False, epot=False):
    """Two-pass force on a given list.  xp (N,3); indi (K,N) 1-based (same order as xp).
    Returns (fp (N,3), den (N,), vtensor(3,3)|None, epot (N,)|None)."""
    xp = np.asarray(xp)
    n = xp.shape[0]
    xpf = _colmajor(xp)
    ityp, statu, kvois = _i32(ityp), _i32(statu), _i32(kvois)
    indi = _i32(indi)
    ts = tables.cstruct()
    den = np.zeros(n)
    fp = np.zeros(3 * n)
    L = lib()
    L.orc_force_pass1(C.c_int(n), C.c_int(0), C.c_int(n), _d(xpf), _i(ityp), _i(statu), _i(kvois), _i(indi),
                      C.c_int(n), _vec3(zl), _ivec3(ifpd), _IDENT, C.byref(ts), _d(den))
    vt = np.zeros(9) if virial else None
    L.orc_force_pass2(C.c_int(n), C.c_int(0), C.c_int(n), _d(xpf), _i(ityp), _i(statu), _i(kvois), _i(indi),
                      C.c_int(n), _vec3(zl), _ivec3(ifpd), _IDENT, C.byref(ts), _d(den), _d(fp), C.c_int(n),
                      _d(vt) if virial else None)
    ep = None
    if epot:
        ep = np.zeros(n)
        L.orc_force_epot(C.c_int(n), C.c_int(0), C.c_int(n), _d(xpf), _i(ityp), _i(statu), _i(kvois), _i(indi),
                         C.c_int(n), _vec3(zl), _ivec3(ifpd), _IDENT, C.byref(ts), _d(ep))
    return _from_colmajor(fp, n), den, (vt.reshape(3, 3).T if virial else None), ep


def predictor(xp, xp1, fp, dis, statu, ityp, cm, h, boxlow, zl, ifpd):
    n = xp.shape[0]
    x, v, f, d = _colmajor(xp), _colmajor(xp1), _colmajor(fp), _colmajor(dis)
    st = _i32(statu).copy()
    up = [boxlow[k] + zl[k] for k in range(3)]
    lib().orc_predictor(C.c_int(n), _d(x), _d(v), _d(f), _d(d), _i(st), _i(_i32(ityp)), _d(_f64(cm)),
                        C.c_double(h), _vec3(boxlow), _vec3(up), _vec3(zl), _ivec3(ifpd))
    return _from_colmajor(x, n), _from_colmajor(v, n), _from_colmajor(d, n), st


def corrector(xp1, fp, statu, ityp, cm, h):
    n = xp1.shape[0]
    v, f = _colmajor(xp1), _colmajor(fp)
    lib().orc_corrector(C.c_int(n), _d(v), _d(f), _i(_i32(statu)), _i(_i32(ityp)), _d(_f64(cm)), C.c_double(h))
    return _from_colmajor(v, n)


def ekin(xp1, statu, ityp, cm):
    n = xp1.shape[0]
    out = np.zeros(n)
    lib().orc_ekin(C.c_int(n), _d(_colmajor(xp1)), _i(_i32(statu)), _i(_i32(ityp)), _d(_f64(cm)), _d(out))
    return out


def epc(xp1, fp, statu, ityp, enable, cm, te, alpha, cut, he):
    n = xp1.shape[0]
    f = _colmajor(fp)
    ng = len(cm)
    lib().orc_epc(C.c_int(n), _d(_colmajor(xp1)), _d(f), _i(_i32(statu)), _i(_i32(ityp)), C.c_int(ng),
                  _i(_i32(enable)), _d(_f64(cm)), _d(_f64(te)), _d(_f64(alpha)), _d(_f64(cut)), _d(_f64(he)))
    return _from_colmajor(f, n)


class MD:
    """Whole-step driver (GMD For_One_Step) on the CPU oracle."""

    def __init__(self, nbox, napb, xp, xp1, ityp, statu, cm, boxlow, zl, ifpd, nb_rm, mxkvois, tables):
        self.n = nbox * napb
        self.tables = tables
        self._ts = tables.cstruct()
        nb_rm = _f64(nb_rm)
        ng = len(cm)
        self.h = lib().orc_md_create(
            C.c_int(nbox), C.c_int(napb), _d(_colmajor(xp)), _d(_colmajor(xp1)), _i(_i32(ityp)), _i(_i32(statu)),
            C.c_int(ng), _d(_f64(cm)), _vec3(boxlow), _vec3(zl), _ivec3(ifpd), _d(nb_rm), C.c_int(mxkvois),
            C.byref(self._ts))
        self.h = C.c_void_p(self.h)
        self.mxkvois = mxkvois

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_md_destroy(self.h)
            self.h = None

    def set_epc(self, enable, te, alpha, cut, he):
        lib().orc_md_set_epc(self.h, _i(_i32(enable)), _d(_f64(te)), _d(_f64(alpha)), _d(_f64(cut)), _d(_f64(he)))

    def rebuild(self):
        return lib().orc_md_rebuild(self.h)

    def force(self, virial=False):
        lib().orc_md_force(self.h, C.c_int(1 if virial else 0))

    def epot(self):
        lib().orc_md_epot(self.h)

    def step(self, itime, it0, nb_uptab, h):
        return lib().orc_md_step(self.h, C.c_int(itime), C.c_int(it0), C.c_int(nb_uptab), C.c_double(h))

    def global_t(self):
        lib().orc_md_global_t.restype = C.c_double
        return float(lib().orc_md_global_t(self.h))

    def vel_scaling(self, dt):
        return int(lib().orc_md_vel_scaling(self.h, C.c_double(dt)))

    def check_timestep(self, th, h2s2, mxd2):
        return int(lib().orc_md_check_timestep(self.h, C.c_double(th), C.c_double(h2s2), C.c_double(mxd2)))

    def steepest0(self, mxnumsteps, alpha0, maxdis, mindis, minepot):
        """Do_Steepest0_Forsteps_DEV; returns (IFLAG, MAXMOVE [cm], DELEPOT [erg])"""
        mm, de = C.c_double(0.0), C.c_double(0.0)
        fl = lib().orc_md_steepest0(self.h, C.c_int(mxnumsteps), C.c_double(alpha0), C.c_double(maxdis), C.c_double(mindis),
                                    C.c_double(minepot), C.byref(mm), C.byref(de))
        return int(fl), mm.value, de.value

    def get(self):
        n = self.n
        xp, xp1, fp, dis = (np.zeros(3 * n) for _ in range(4))
        epot, ekin = np.zeros(n), np.zeros(n)
        statu, gid = np.zeros(n, np.int32), np.zeros(n, np.int32)
        vt = np.zeros(9)
        lib().orc_md_get(self.h, _d(xp), _d(xp1), _d(fp), _d(epot), _d(ekin), _d(dis), _i(statu), _i(gid), _d(vt))
        return dict(xp=_from_colmajor(xp, n), xp1=_from_colmajor(xp1, n), fp=_from_colmajor(fp, n),
                    dis=_from_colmajor(dis, n), epot=epot, ekin=ekin, statu=statu, gid=gid,
                    vtensor=vt.reshape(3, 3).T)

    def nlist(self):
        n = self.n
        kv = np.ctypeslib.as_array(lib().orc_md_kvois(self.h), shape=(n,)).copy()
        ind = np.ctypeslib.as_array(lib().orc_md_indi(self.h), shape=(self.mxkvois, n)).copy()
        return kv, ind
