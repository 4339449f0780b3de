"""Slab decomposition driven from the library (mdb_dd_*), all ranks in ONE process on ONE GPU (in-process backend: the
exchanges are device-to-device copies between the ranks' contexts; every kernel, range and rebuild step is the one the
NCCL backend runs).  Each rank's owned slice must be bit-identical to the same slice of a single-context run."""
import numpy as np
import pytest

import util
from msmpscu_b200 import capi

pytestmark = pytest.mark.gpu
H, IT0, NUP = 0.5e-15, 1, 10
EPC = ([1], [300.0], [1.0e-12], [0.1], [100.0 * util.CP_EVERG])


def _make(c):
    ctx = capi.Context(0)
    ctx.box_set(c.nbox, c.napb, c.boxlow, c.zl, c.ifpd, c.mass)
    ctx.upload(capi.F_XP, c.xp); ctx.upload(capi.F_XP1, c.xp1)
    ctx.upload(capi.F_ITYP, c.ityp); ctx.upload(capi.F_STATU, c.statu)
    ctx.tables_set(util.product_tables(c), c.ru * c.ru)
    ctx.set_option(capi.OPT_FORCE_PATH, capi.FORCE_PATH_TILED)
    ctx.set_option(capi.OPT_TILED_BANKORDER, 1)     # (auto would leave it off for a box this small: the big runs have it on)
    ctx.nlist_init(c.nb_rm, c.mxkvois)
    ctx.epc_set(*EPC)
    return ctx


def _case():
    """bcc W, 8 x 8 x 20 cells (8 z-layers of list cells), hot; atoms close to a layer face are shot across it (0.03 a0 per
    step), so that every rebuild moves atoms from one rank's slab into its neighbour's"""
    c = util.bcc_case((8, 8, 20), seed=404, temp=1500.0)
    ncz = 8
    t = (c.xp[:, 2] - c.boxlow[2]) / c.zl[2] * ncz
    frac = t - np.floor(t)
    thick = c.zl[2] / ncz
    up = frac > 1.0 - 0.32 * c.rr / thick
    down = frac < 0.32 * c.rr / thick
    v = 0.03 * c.rr / H
    c.xp1 = c.xp1.copy()
    c.xp1[up, 2] = v
    c.xp1[down, 2] = -v
    return c


@pytest.mark.parametrize("world", [2, 3, 4])
def test_decomposed_run_matches_single_context(world):
    """world = 2: both neighbours of a rank are the same rank; 3: unequal slabs (2, 3, 3 layers); 4: two-layer slabs, every
    owned layer is somebody's ghost.  35 NVT steps = four rebuilds, i.e. three LOCAL rebuilds with atoms changing owner."""
    nsteps = 35
    c = _case()
    n = c.xp.shape[0]
    full = _make(c)
    full.nlist_build(); full.force(capi.FORCE)
    full.run(0, nsteps, IT0, NUP, H)
    ctxs = [_make(c) for _ in range(world)]
    capi.dd_local_attach(ctxs)
    ctxs[0].dd_build()
    ctxs[0].dd_force(capi.FORCE)
    ctxs[0].dd_run(0, nsteps, IT0, NUP, H)
    gid_f = full.download(capi.F_GID, capi.ORDER_CELL)
    kv_f = full.download(capi.F_KVOIS, capi.ORDER_CELL)
    full.force(capi.EPOT)
    ep_f = full.download(capi.F_EPOT, capi.ORDER_CELL)
    owned = 0
    moved = 0
    for r, ctx in enumerate(ctxs):
        info = ctx.dd_info()
        a0, a1 = info["a0"], info["a1"]
        owned += a1 - a0
        assert np.array_equal(ctx.download(capi.F_GID, capi.ORDER_CELL)[a0:a1], gid_f[a0:a1]), "rank %d order" % r
        for name, f in (("xp", capi.F_XP), ("xp1", capi.F_XP1), ("fp", capi.F_FP), ("den", capi.F_DEN), ("dis", capi.F_DIS)):
            x, y = full.download(f, capi.ORDER_CELL)[a0:a1], ctx.download(f, capi.ORDER_CELL)[a0:a1]
            assert np.array_equal(x, y), "rank %d field %s differs (max %g)" % (r, name, np.max(np.abs(x - y)))
        assert np.array_equal(ctx.download(capi.F_KVOIS, capi.ORDER_CELL)[a0:a1], kv_f[a0:a1])
        ctx.force(capi.EPOT)
        assert np.array_equal(ctx.download(capi.F_EPOT, capi.ORDER_CELL)[a0:a1], ep_f[a0:a1])
    assert owned == n
    # reductions over the ranks
    vt_f = full.force(capi.FORCE | capi.VIRIAL)
    vt_d = ctxs[0].dd_force(capi.FORCE | capi.VIRIAL)
    assert util.relerr(vt_d, vt_f) < 1e-12
    t_f = full.global_t()
    assert abs(ctxs[0].dd_global_t() - t_f) < 1e-12 * t_f
    for ctx in ctxs:
        ctx.close()
    full.close()


def test_atoms_do_change_owner_in_that_run():
    """the case above is only a test of the local rebuild if atoms cross slab faces: count them on the host"""
    c = _case()
    full = _make(c)
    full.nlist_build(); full.force(capi.FORCE)
    z0 = full.download(capi.F_XP)[:, 2]
    full.run(0, 35, IT0, NUP, H)
    z1 = full.download(capi.F_XP)[:, 2]
    ncz = 8
    lay = lambda z: np.floor((z - c.boxlow[2]) / c.zl[2] * ncz).astype(int) % ncz
    assert np.count_nonzero(lay(z0) != lay(z1)) > 10
    full.close()
