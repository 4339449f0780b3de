"""ctypes binding of libmdpscu_b200.so (the C ABI in include/mdpscu_b200.h).

The product path has NO CPU fallback: if the shared library is missing this module raises at
import of the symbol table, and creating a context without a CUDA device raises MDBError
(MDB_ERR_NOGPU).  Nothing here imports the CPU oracle.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmdpscu_b200.so")

MXGROUP = 10

OK = 0
ERR_CUDA, ERR_ARG, ERR_STATE, ERR_UNSUPPORTED, ERR_NOMEM, ERR_NOGPU = -1, -2, -3, -4, -5, -6
ORDER_ORIGINAL, ORDER_CELL = 0, 1
(F_XP, F_XP1, F_FP, F_DIS, F_EPOT, F_EKIN, F_DEN, F_ITYP, F_STATU, F_GID, F_GIDINV, F_IC, F_KVOIS, F_INDI,
 F_NAC, F_NAAC, F_IA1TH) = range(17)
POT_EAM, POT_FS = 0, 1
FORCE, VIRIAL, EPOT, DEN, NOPASS1 = 1, 2, 4, 8, 16
F_POS4, F_D2MAX = 17, 18
K_NAMES = ("cellsort", "nlist", "pass1", "pass2", "epot", "predict", "correct", "other", "exchange")
K_COUNT = 9
LIB_MARINICA_EAM2, LIB_BONNY_EAM1, LIB_ACKLAND_FS_W = 1, 2, 3

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)

# every symbol include/mdpscu_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "mdb_device_count": (C.c_int, []),
    "mdb_ctx_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "mdb_ctx_destroy": (None, [C.c_void_p]),
    "mdb_last_error": (C.c_char_p, [C.c_void_p]),
    "mdb_version": (C.c_char_p, []),
    "mdb_ctx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdb_ctx_stream": (C.c_void_p, [C.c_void_p]),
    "mdb_sync": (C.c_int, [C.c_void_p]),
    "mdb_box_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_ip, C.c_int, c_dp]),
    "mdb_natom": (C.c_int, [C.c_void_p]),
    "mdb_state_upload": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "mdb_state_download": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "mdb_devptr": (C.c_void_p, [C.c_void_p, C.c_int]),
    "mdb_tables_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, c_dp, c_dp, c_dp, c_dp, C.c_int,
                                 C.c_int, C.c_double, c_dp, c_dp, c_ip, c_ip, C.c_double]),
    "mdb_tables_clear": (C.c_int, [C.c_void_p]),
    "mdb_nlist_init": (C.c_int, [C.c_void_p, c_dp, C.c_int]),
    "mdb_nlist_build": (C.c_int, [C.c_void_p]),
    "mdb_nlist_copyout": (C.c_int, [C.c_void_p, c_ip, c_ip, C.c_int]),
    "mdb_nlist_cellinfo": (C.c_int, [C.c_void_p, c_ip, c_ip, c_ip]),
    "mdb_nlist_overflow": (C.c_int, [C.c_void_p]),
    "mdb_nlist_clear": (C.c_int, [C.c_void_p]),
    "mdb_force": (C.c_int, [C.c_void_p, C.c_uint, c_dp]),
    "mdb_predict": (C.c_int, [C.c_void_p, C.c_double]),
    "mdb_correct": (C.c_int, [C.c_void_p, C.c_double]),
    "mdb_ekin": (C.c_int, [C.c_void_p]),
    "mdb_epc_set": (C.c_int, [C.c_void_p, c_ip, c_dp, c_dp, c_dp, c_dp]),
    "mdb_epc_apply": (C.c_int, [C.c_void_p]),
    "mdb_step": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double]),
    "mdb_run": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "mdb_embed_overruns": (C.c_int, [C.c_void_p]),
    "mdb_state_save": (C.c_int, [C.c_void_p]),
    "mdb_state_restore": (C.c_int, [C.c_void_p]),
    "mdb_compare": (C.c_int, [C.c_void_p, c_dp, c_ip, C.c_double, c_ip, c_ip, c_ip, c_ip]),
    "mdb_box_temperatures": (C.c_int, [C.c_void_p, c_dp]),
    "mdb_active_region": (C.c_int, [C.c_void_p, C.c_int, c_ip, C.c_double, C.c_int]),
    "mdb_active_all": (C.c_int, [C.c_void_p, C.c_int]),
    "mdb_stopping_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, c_ip, c_ip, c_dp]),
    "mdb_stopping_apply": (C.c_int, [C.c_void_p, C.c_double]),
    "mdb_stopping_options": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mdb_stopping_eloss": (C.c_int, [C.c_void_p, c_dp, C.c_int]),
    "mdb_pka_insert": (C.c_int, [C.c_void_p, C.c_int, C.c_double, c_dp]),
    "mdb_dd_nccl_id": (C.c_int, [C.c_void_p]),
    "mdb_dd_nccl_init": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdb_dd_local_attach": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "mdb_dd_build": (C.c_int, [C.c_void_p]),
    "mdb_dd_force": (C.c_int, [C.c_void_p, C.c_uint, c_dp]),
    "mdb_dd_run": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "mdb_dd_global_t": (C.c_int, [C.c_void_p, c_dp]),
    "mdb_timestep_limit": (C.c_int, [C.c_void_p, C.c_double, C.c_double, c_dp]),
    "mdb_run_sched": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, c_dp, c_dp]),
    "mdb_dd_run_sched": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, c_dp, c_dp]),
    "mdb_run_async": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "mdb_state_download_async": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "mdb_global_t": (C.c_int, [C.c_void_p, c_dp]),
    "mdb_vel_scaling": (C.c_int, [C.c_void_p, C.c_double]),
    "mdb_check_timestep": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, c_ip]),
    "mdb_steepest": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, c_ip, c_dp, c_dp]),
    "mdb_nlist_reorder_nearest": (C.c_int, [C.c_void_p, C.c_int]),
    "mdb_epc_correct": (C.c_int, [C.c_void_p, C.c_double]),
    "mdb_atomic_stress": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mdb_atomic_stress_host": (C.c_int, [C.c_void_p, c_dp, C.c_int]),
    "mdb_thermalize": (C.c_int, [C.c_void_p, C.c_double, C.c_ulonglong, C.c_uint]),
    "mdb_thermalize_bits": (C.c_int, [C.c_ulonglong, C.c_uint, C.c_uint, C.POINTER(C.c_uint)]),
    "mdb_philox4x32_10": (C.c_int, [C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint)]),
    "mdb_lbfgs": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, c_ip, c_ip, c_ip]),
    "mdb_damping": (C.c_int, [C.c_void_p]),
    "mdb_dyndamp": (C.c_int, [C.c_void_p, C.c_int, C.c_double, C.c_double, c_ip, c_dp]),
    "mdb_cg": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, c_ip, c_dp]),
    "mdb_dd_set": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mdb_dd_info": (C.c_int, [C.c_void_p, c_ip]),
    "mdb_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mdb_get_option": (C.c_int, [C.c_void_p, C.c_int]),
    "mdb_prof_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "mdb_prof_reset": (C.c_int, [C.c_void_p]),
    "mdb_prof_get": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), c_dp]),
    "mdb_launch_count": (C.c_longlong, [C.c_void_p]),
    "mdb_host_ftable_create": (C.c_int, [C.c_int, C.c_int, c_ip, C.c_int, C.c_int, C.c_double, C.c_double, c_ip, c_ip,
                                         c_ip, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
    "mdb_host_setfl_info": (C.c_int, [C.c_char_p, c_ip, c_ip, c_ip, c_dp, c_dp, C.c_char_p, C.c_int, c_ip, c_dp, c_dp]),
    "mdb_host_setfl_ftable": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_double, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                        c_dp, c_dp, c_dp]),
    "mdb_host_lspt_info": (C.c_int, [C.c_char_p, c_ip, c_dp, c_dp, C.c_char_p, C.c_int]),
    "mdb_host_lspt_ftable": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_double, c_ip, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                       c_dp, c_dp, c_dp]),
    "mdb_host_moldy_ftable": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_double, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp,
                                        c_dp, c_dp]),
    "mdb_host_ftable_export": (C.c_int, [C.c_char_p, C.c_int, C.c_int, c_ip, C.c_int, C.c_double, c_dp, c_dp, c_dp, c_dp,
                                         C.c_int, c_ip, C.c_int, C.c_double, c_dp, c_dp]),
    "mdb_host_ftable_file_info": (C.c_int, [C.c_char_p, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_ip, c_dp, c_dp]),
    "mdb_host_ftable_import": (C.c_int, [C.c_char_p, C.c_int, c_ip, C.c_int, C.c_int, C.c_double, c_ip, c_ip, c_ip, c_ip, c_ip,
                                         c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]),
}

OPT_FORCE_PATH, OPT_TILED_LANES, OPT_TILED_CLASSES, OPT_ACTIVE_PATH, OPT_TILED_THREADS, OPT_FUSE_EPILOGUE, OPT_TILED_STAGES = 0, 1, 2, 3, 4, 5, 6
OPT_TILED_BANKORDER = 7
OPT_TILE_GUARD = 8
OPT_PDL = 9


class Sched(C.Structure):
    """mdb_sched: CtrlParam%IHDUP / HMI / HMX / DMX and NB_UPTABMI / NB_UPTABMX / NB_DBITAB (seconds, cm)"""
    _fields_ = [("ihdup", C.c_int), ("hmi", C.c_double), ("hmx", C.c_double), ("dmx", C.c_double),
                ("nb_uptabmi", C.c_int), ("nb_uptabmx", C.c_int), ("nb_dbitab", C.c_int)]
FORCE_PATH_AUTO, FORCE_PATH_GENERIC, FORCE_PATH_TILED = 0, 1, 2
QUENCH_LSEARCH = 65536  # CP_DAMPSCHEME_LSEARCH

_lib = None


class MDBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mdpscu_b200 error %d: %s" % (code, msg))
        self.code = code


def load():
    """Load the shared library and bind every symbol.  Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "msmpscu_b200: %s not found. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C msmpscu_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dp(a):
    return a.ctypes.data_as(c_dp)


def ip(a):
    return a.ctypes.data_as(c_ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def dd_nccl_id():
    """rank 0: the 128-byte NCCL unique id to hand to every rank's Context.dd_nccl_init"""
    buf = C.create_string_buffer(128)
    rc = load().mdb_dd_nccl_id(buf)
    if rc < 0:
        raise MDBError(rc, "mdb_dd_nccl_id failed (libnccl.so.2 not found?)")
    return buf.raw


def dd_local_attach(ctxs):
    """in-process backend of the decomposed run: all ranks' contexts in this process on one device (single-GPU tests)"""
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    rc = load().mdb_dd_local_attach(arr, len(ctxs))
    if rc < 0:
        raise MDBError(rc, ctxs[0].lib.mdb_last_error(ctxs[0].h).decode())


def colmajor(a):
    """(N,k) -> flat Fortran-order copy (the reference's array layout)."""
    a = np.asarray(a, dtype=np.float64)
    return np.ascontiguousarray(a.T).ravel() if a.ndim == 2 else np.ascontiguousarray(a)


def from_colmajor(flat, n, ncol):
    return np.ascontiguousarray(flat.reshape(ncol, n).T) if ncol > 1 else flat


class Context:
    """Thin RAII wrapper of mdb_ctx; methods map 1:1 to the C ABI."""

    def __init__(self, device=0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.mdb_ctx_create(int(device), C.byref(h))
        if rc != OK:
            msg = {ERR_NOGPU: "no CUDA device visible (the product path has no CPU fallback)"}.get(rc, "mdb_ctx_create failed")
            raise MDBError(rc, msg)
        self.h = h
        self.n = 0
        self.mxkvois = 0
        self.nc = 0

    def close(self):
        if getattr(self, "h", None):
            self.lib.mdb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc < 0:
            raise MDBError(rc, self.lib.mdb_last_error(self.h).decode())
        return rc

    # ---- box / state
    def box_set(self, nbox, napb, boxlow, boxsize, ifpd, mass, boxshape=None):
        lo, sz, m = f64(boxlow), f64(boxsize), f64(mass)
        pd = i32(ifpd)
        bs = f64(boxshape).T.ravel().copy() if boxshape is not None else None
        self._chk(self.lib.mdb_box_set(self.h, nbox, napb, dp(lo), dp(sz), dp(bs) if bs is not None else None, ip(pd),
                                       len(m), dp(m)))
        self.n = nbox * napb
        self.ng = len(m)

    def upload(self, field, arr, order=ORDER_ORIGINAL):
        if field in (F_ITYP, F_STATU):
            a = i32(arr)
        else:
            a = colmajor(arr)
        self._chk(self.lib.mdb_state_upload(self.h, field, a.ctypes.data_as(C.c_void_p), order))

    def upload_raw(self, field, ptr, order=ORDER_ORIGINAL):
        """host pointer (int) with the reference layout, e.g. a pinned torch tensor's data_ptr()."""
        self._chk(self.lib.mdb_state_upload(self.h, field, C.c_void_p(ptr), order))

    def download_raw(self, field, ptr, order=ORDER_ORIGINAL):
        self._chk(self.lib.mdb_state_download(self.h, field, C.c_void_p(ptr), order))

    def download_raw_async(self, field, ptr, order=ORDER_ORIGINAL):
        """enqueue the CopyOut of one field into page-locked host memory; complete after sync()"""
        self._chk(self.lib.mdb_state_download_async(self.h, field, C.c_void_p(ptr), order))

    def download(self, field, order=ORDER_ORIGINAL):
        n = self.n
        if field in (F_XP, F_XP1, F_FP, F_DIS):
            buf = np.empty(3 * n)
            self._chk(self.lib.mdb_state_download(self.h, field, buf.ctypes.data_as(C.c_void_p), order))
            return from_colmajor(buf, n, 3)
        if field in (F_EPOT, F_EKIN, F_DEN):
            buf = np.empty(n)
        elif field in (F_NAC, F_NAAC, F_IA1TH):
            buf = np.empty(self.cellinfo()[1], dtype=np.int32)
        else:
            buf = np.empty(n, dtype=np.int32)
        self._chk(self.lib.mdb_state_download(self.h, field, buf.ctypes.data_as(C.c_void_p), order))
        return buf

    def devptr(self, field):
        return self.lib.mdb_devptr(self.h, field)

    # ---- tables
    def tables_set(self, t, ru2max):
        """t: host MDForceTable-like object with Fortran-layout arrays (see forcetable.MDForceTable)."""
        self._chk(self.lib.mdb_tables_set(
            self.h, t.pot_type, t.nkind, t.ntab, t.csi, dp(t.potr), dp(t.fpotr), dp(t.potb), dp(t.fpotb), t.nkind1,
            t.nembd, t.rhod, dp(t.fembd), dp(t.dfembd), ip(t.kpair), ip(t.kembd), float(ru2max)))

    def tables_clear(self):
        self._chk(self.lib.mdb_tables_clear(self.h))

    # ---- neighbour list
    def nlist_init(self, nb_rm, mxkvois):
        a = f64(np.asarray(nb_rm, dtype=np.float64).T).ravel().copy()
        self._chk(self.lib.mdb_nlist_init(self.h, dp(a), int(mxkvois)))
        self.mxkvois = int(mxkvois)

    def nlist_build(self):
        return self._chk(self.lib.mdb_nlist_build(self.h))

    def nlist_copyout(self, order=ORDER_CELL):
        kv = np.empty(self.n, dtype=np.int32)
        ind = np.empty(self.n * self.mxkvois, dtype=np.int32)
        self._chk(self.lib.mdb_nlist_copyout(self.h, ip(kv), ip(ind), order))
        return kv, ind.reshape(self.mxkvois, self.n)

    def cellinfo(self):
        nc3 = (C.c_int * 3)()
        nc, mx = C.c_int(), C.c_int()
        self._chk(self.lib.mdb_nlist_cellinfo(self.h, nc3, C.byref(nc), C.byref(mx)))
        return list(nc3), nc.value, mx.value

    def nlist_overflow(self):
        return self._chk(self.lib.mdb_nlist_overflow(self.h))

    # ---- force / integrator
    def force(self, flags=FORCE):
        vt = np.zeros(9)
        self._chk(self.lib.mdb_force(self.h, flags, dp(vt)))
        return vt.reshape(3, 3).T if (flags & VIRIAL) else None

    def predict(self, h):
        self._chk(self.lib.mdb_predict(self.h, float(h)))

    def epc_correct(self, h):
        self._chk(self.lib.mdb_epc_correct(self.h, float(h)))

    def correct(self, h):
        self._chk(self.lib.mdb_correct(self.h, float(h)))

    def ekin(self):
        self._chk(self.lib.mdb_ekin(self.h))

    def epc_set(self, enable, te, alpha, cut, he):
        self._chk(self.lib.mdb_epc_set(self.h, ip(i32(enable)), dp(f64(te)), dp(f64(alpha)), dp(f64(cut)), dp(f64(he))))

    def epc_apply(self):
        self._chk(self.lib.mdb_epc_apply(self.h))

    def step(self, itime, it0, nb_uptab, h):
        return self._chk(self.lib.mdb_step(self.h, itime, it0, nb_uptab, float(h)))

    def run(self, itime0, nsteps, it0, nb_uptab, h):
        return self._chk(self.lib.mdb_run(self.h, itime0, nsteps, it0, nb_uptab, float(h)))

    def timestep_limit(self, hmx, dmx):
        """Predictor_DEV's halving loop (IHDUP < 0): the largest HMX/2^k that moves no active atom further than DMX"""
        hh = C.c_double(0.0)
        self._chk(self.lib.mdb_timestep_limit(self.h, float(hmx), float(dmx), C.byref(hh)))
        return hh.value

    def run_sched(self, itime0, nsteps, it0, sched, h, time_s=0.0):
        """the GMD time loop with its step-size / list-period schedules -> (out-of-box count, H after the block, TIME [s])"""
        hh, tt = C.c_double(float(h)), C.c_double(float(time_s))
        rc = self._chk(self.lib.mdb_run_sched(self.h, itime0, nsteps, it0, C.byref(sched), C.byref(hh), C.byref(tt)))
        return rc, hh.value, tt.value

    # ---- cascade physics (csrc/mdb_cascade.cu)
    def active_region(self, centpart=None, ekin_erg=None, extend=1, keep=False, by_neighbours=False):
        """ActivateRegion_DEV by cells (or through the neighbour list, CP_BYNB_AR); returns the number of active atoms"""
        method = (1 if centpart is not None else 0) | (2 if ekin_erg is not None else 0) | (4 if keep else 0) | (8 if by_neighbours else 0)
        cp = i32(centpart) if centpart is not None else None
        return self._chk(self.lib.mdb_active_region(self.h, method, ip(cp) if cp is not None else None,
                                                    float(ekin_erg or 0.0), int(extend)))

    def active_all(self, on=True):
        self._chk(self.lib.mdb_active_all(self.h, 1 if on else 0))

    def stopping_set(self, etab, stab, kpair, enable, mden):
        """etab (NE,), stab (NE, NK) [columns = tables], kpair (NG, NG) 1-based, enable (NG,), mden (NG,)"""
        e = f64(etab)
        st = np.ascontiguousarray(np.asarray(stab, dtype=np.float64).reshape(len(e), -1).T).ravel()   # column-major STAB(NE,NK)
        kp = np.ascontiguousarray(np.asarray(kpair, dtype=np.int32).T).ravel()
        self._stop_keep = (e, st, kp, i32(enable), f64(mden))
        self._chk(self.lib.mdb_stopping_set(self.h, len(e), st.size // len(e), dp(e), dp(st), ip(kp), ip(self._stop_keep[3]),
                                            dp(self._stop_keep[4])))

    def stopping_apply(self, dt=0.0):
        self._chk(self.lib.mdb_stopping_apply(self.h, float(dt)))

    def stopping_options(self, local_density=False, save_eloss=False):
        """Reset_STMOD_DEV's switches: local-density model (ST_CTRL%MDEN < 0), per-atom energy-loss accumulation"""
        self._chk(self.lib.mdb_stopping_options(self.h, 1 if local_density else 0, 1 if save_eloss else 0))

    def stopping_eloss(self, reset=False):
        """accumulated inelastic energy loss per atom [erg], ORIGINAL order"""
        e = np.zeros(self.n)
        self._chk(self.lib.mdb_stopping_eloss(self.h, dp(e), 1 if reset else 0))
        return e

    def pka_insert(self, orig_id, ekin_erg, direction):
        d = f64(direction)
        self._chk(self.lib.mdb_pka_insert(self.h, int(orig_id), float(ekin_erg), dp(d)))

    def box_temperatures(self, nbox):
        t = np.zeros(int(nbox))
        self._chk(self.lib.mdb_box_temperatures(self.h, dp(t)))
        return t

    # ---- PARREP event detection (Do_ChangeDetect pieces)
    def state_save(self):
        self._chk(self.lib.mdb_state_save(self.h))

    def state_restore(self):
        return self._chk(self.lib.mdb_state_restore(self.h))

    def compare(self, xp_ini, drtol, mask=None, per_atom=False, nbox=1):
        """Do_Compare: returns (flag_box, IBT, NCB[, Flag per atom])"""
        x = colmajor(xp_ini)
        fb = np.zeros(int(nbox), np.int32)
        fa = np.zeros(self.n, np.int32) if per_atom else None
        m = i32(mask) if mask is not None else None
        ibt, ncb = C.c_int(0), C.c_int(0)
        self._chk(self.lib.mdb_compare(self.h, dp(x), ip(m) if m is not None else None, float(drtol), ip(fb),
                                       ip(fa) if fa is not None else None, C.byref(ibt), C.byref(ncb)))
        return (fb, ibt.value, ncb.value, fa) if per_atom else (fb, ibt.value, ncb.value)

    def embed_overruns(self):
        """rho > RHOMX events of the density pass since the last call"""
        return self._chk(self.lib.mdb_embed_overruns(self.h))

    def run_async(self, itime0, nsteps, it0, nb_uptab, h):
        """enqueue nsteps steps; sync() returns the block's out-of-box count"""
        return self._chk(self.lib.mdb_run_async(self.h, itime0, nsteps, it0, nb_uptab, float(h)))

    def global_t(self):
        t = C.c_double(0.0)
        self._chk(self.lib.mdb_global_t(self.h, C.byref(t)))
        return t.value

    def vel_scaling(self, dt):
        self._chk(self.lib.mdb_vel_scaling(self.h, float(dt)))

    def check_timestep(self, th, h2s2, dmx2):
        fl = C.c_int(0)
        self._chk(self.lib.mdb_check_timestep(self.h, float(th), float(h2s2), float(dmx2), C.byref(fl)))
        return fl.value

    def steepest(self, mxnumsteps, alpha, maxdis, mindis, minepot, meth=0):
        """Do_Steepest_Forsteps_DEV on the current list; returns (IFLAG, MAXMOVE [cm], DELEPOT [erg])."""
        fl, mm, de = C.c_int(0), C.c_double(0.0), C.c_double(0.0)
        self._chk(self.lib.mdb_steepest(self.h, int(mxnumsteps), int(meth), float(alpha), float(maxdis), float(mindis),
                                        float(minepot), C.byref(fl), C.byref(mm), C.byref(de)))
        return fl.value, mm.value, de.value

    def nlist_reorder_nearest(self, nearest):
        """Reorder_NeighBoreList_Nearest_Dev: keep the `nearest` closest neighbours, by increasing distance, in place."""
        self._chk(self.lib.mdb_nlist_reorder_nearest(self.h, int(nearest)))

    def atomic_stress(self, order=ORDER_ORIGINAL):
        """pCalAVStress: per-atom virial tensor (n, 9), columns 11,12,13,21,...,33."""
        ap = np.zeros(9 * self.n)
        self._chk(self.lib.mdb_atomic_stress_host(self.h, dp(ap), int(order)))
        return np.ascontiguousarray(ap.reshape(9, self.n).T)

    def thermalize(self, ti, seed, draw=0):
        """Thermalizing_MC_DEV: Maxwell velocities at ti [K] + per-box momentum removal (Philox4x32-10 keyed by seed/draw/atom id)."""
        self._chk(self.lib.mdb_thermalize(self.h, float(ti), int(seed), int(draw)))

    def lbfgs(self, mxnumsteps, msave, factr, pgtol):
        """DO_LBFGSB_FORSTEPS_DEV; returns (IFLAG, force evaluations, accepted steps)."""
        fl, nfg, nit = C.c_int(0), C.c_int(0), C.c_int(0)
        self._chk(self.lib.mdb_lbfgs(self.h, int(mxnumsteps), int(msave), float(factr), float(pgtol), C.byref(fl), C.byref(nfg),
                                     C.byref(nit)))
        return fl.value, nfg.value, nit.value

    def damping(self):
        self._chk(self.lib.mdb_damping(self.h))

    def dyndamp(self, mxnumsteps, h, minepot):
        """Do_DynDamp_Forsteps_DEV; returns (IFLAG, DELEPOT [erg])."""
        fl, de = C.c_int(0), C.c_double(0.0)
        self._chk(self.lib.mdb_dyndamp(self.h, int(mxnumsteps), float(h), float(minepot), C.byref(fl), C.byref(de)))
        return fl.value, de.value

    def cg(self, mxnumsteps, maxdis, mindis, minepot, meth=0):
        """Do_CG_Forsteps_DEV on the current list (meth & QUENCH_LSEARCH: the line-search variant); returns (IFLAG, DELEPOT [erg])."""
        fl, de = C.c_int(0), C.c_double(0.0)
        self._chk(self.lib.mdb_cg(self.h, int(mxnumsteps), int(meth), float(maxdis), float(mindis), float(minepot),
                                  C.byref(fl), C.byref(de)))
        return fl.value, de.value

    def dd_set(self, rank, nranks):
        self._chk(self.lib.mdb_dd_set(self.h, int(rank), int(nranks)))

    def dd_info(self):
        out = (C.c_int * 16)()
        self._chk(self.lib.mdb_dd_info(self.h, out))
        keys = ("a0", "a1", "gb0", "gb1", "ga0", "ga1", "sb0", "sb1", "st0", "st1", "below", "above", "cell_lo", "cell_hi",
                "tile_lo", "tile_hi")
        return dict(zip(keys, list(out)))

    # ---- slab decomposition driven from the library (mdb_dd.cu)
    def dd_nccl_init(self, id128: bytes):
        buf = C.create_string_buffer(bytes(id128), 128)
        self._chk(self.lib.mdb_dd_nccl_init(self.h, buf))

    def dd_build(self):
        return self._chk(self.lib.mdb_dd_build(self.h))

    def dd_force(self, flags=FORCE):
        vt = np.zeros(9)
        self._chk(self.lib.mdb_dd_force(self.h, flags, dp(vt)))
        return vt.reshape(3, 3).T if (flags & VIRIAL) else None

    def dd_run(self, itime0, nsteps, it0, nb_uptab, h):
        return self._chk(self.lib.mdb_dd_run(self.h, itime0, nsteps, it0, nb_uptab, float(h)))

    def dd_run_sched(self, itime0, nsteps, it0, sched, h, time_s=0.0):
        hh, tt = C.c_double(float(h)), C.c_double(float(time_s))
        rc = self._chk(self.lib.mdb_dd_run_sched(self.h, itime0, nsteps, it0, C.byref(sched), C.byref(hh), C.byref(tt)))
        return rc, hh.value, tt.value

    def dd_global_t(self):
        t = C.c_double(0.0)
        self._chk(self.lib.mdb_dd_global_t(self.h, C.byref(t)))
        return t.value

    def sync(self):
        return self._chk(self.lib.mdb_sync(self.h))

    def set_stream(self, cuda_stream_ptr):
        self._chk(self.lib.mdb_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def set_option(self, opt, val):
        self._chk(self.lib.mdb_set_option(self.h, opt, val))

    def get_option(self, opt):
        return self.lib.mdb_get_option(self.h, opt)

    # ---- measurement
    def prof_enable(self, on=True):
        self._chk(self.lib.mdb_prof_enable(self.h, 1 if on else 0))

    def prof_reset(self):
        self._chk(self.lib.mdb_prof_reset(self.h))

    def prof_get(self):
        ln = (C.c_longlong * K_COUNT)()
        ms = (C.c_double * K_COUNT)()
        self._chk(self.lib.mdb_prof_get(self.h, ln, ms))
        return {K_NAMES[k]: (int(ln[k]), float(ms[k])) for k in range(K_COUNT)}

    def launch_count(self):
        return int(self.lib.mdb_launch_count(self.h))
