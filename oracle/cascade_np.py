"""NumPy restatements of the cascade-physics procedures around the hot path.

TEST INFRASTRUCTURE ONLY (see oracle/pyorc.py).  Each function cites the reference lines it follows.
"""
import numpy as np

STATU_ACTIVE = 1
STATU_OUTOFBOX = 65536

NIX = (0, -1, -1, -1, 0, 0, -1, 1, -1, 0, 1, -1, 0, 1, 1, 1, 1, 0, 0, -1, 1, -1, 0, 1, -1, 0, 1)
NIY = (0, 0, -1, 1, 1, 0, 0, 0, -1, -1, -1, 1, 1, 1, 0, 1, -1, -1, 0, 0, 0, -1, -1, -1, 1, 1, 1)
NIZ = (0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1)


def activate_region_by_cells(statu, ityp, xp1, inc, cm, ncell, nbox, ifpd, centpart=None, ekin_erg=None, extend=1, keep=False):
    """ActivateRegion_DEV -> ActiveByCells1 -> ActiveByCells0, CommonGPU/MD_ActiveRegion_GPU.F90:1165-1353.
    statu, ityp, inc (1-based cell id per atom, 0 = none) in ONE common atom order; xp1 (N,3); cm per type.
    Returns the new STATU."""
    st = np.array(statu, dtype=np.int64)
    n = st.size
    if not keep:
        st &= ~STATU_ACTIVE                                            # DeActive_All_Kernel :242-279 (ActiveByCells1 :1313-1319)
    if centpart is None and ekin_erg is None:                          # :1207-1209: no intrinsic seed method
        return st.astype(np.int32)
    inbox = (st & STATU_OUTOFBOX) != STATU_OUTOFBOX
    seeds = np.zeros(n, dtype=np.int64)
    if centpart is not None:                                           # CreateSeedByType_Kernel :513-560
        seeds += (inbox & (np.asarray(centpart)[np.asarray(ityp) - 1] > 0)).astype(np.int64)
    if ekin_erg is not None:                                           # CreateSeedByEkin_Kernel :646-700
        ek = 0.5 * np.asarray(cm)[np.asarray(ityp) - 1] * np.sum(np.asarray(xp1) ** 2, axis=1)
        seeds += (inbox & (ek >= ekin_erg)).astype(np.int64)
    ncx, ncy, ncz = (int(v) for v in ncell)
    nc0 = ncx * ncy * ncz
    mark = np.zeros(nc0 * nbox, dtype=np.int64)
    sel = (seeds > 0) & (np.asarray(inc) > 0)
    mark[np.asarray(inc)[sel] - 1] = 1                                 # MarkSeedCell_Kernel0 :1001-1045
    for _ in range(int(extend)):                                       # :1238-1287 (host loop of the reference)
        new = np.zeros_like(mark)
        for c in np.nonzero(mark)[0]:
            ib, r = divmod(int(c), nc0)
            iz, r2 = divmod(r, ncx * ncy)
            iy, ix = divmod(r2, ncx)
            for k in range(27):
                x, y, z = ix + NIX[k], iy + NIY[k], iz + NIZ[k]
                if x >= ncx:
                    x = 0 if ifpd[0] > 0 else x
                elif x < 0:
                    x = ncx - 1 if ifpd[0] > 0 else x
                if y >= ncy:
                    y = 0 if ifpd[1] > 0 else y
                elif y < 0:
                    y = ncy - 1 if ifpd[1] > 0 else y
                if z >= ncz:
                    z = 0 if ifpd[2] > 0 else z
                elif z < 0:
                    z = ncz - 1 if ifpd[2] > 0 else z
                if 0 <= x < ncx and 0 <= y < ncy and 0 <= z < ncz:
                    new[ib * nc0 + (z * ncy + y) * ncx + x] = 1
        mark = new
    act = (np.asarray(inc) > 0) & (mark[np.maximum(np.asarray(inc) - 1, 0)] > 0)
    st[act] |= STATU_ACTIVE                                            # Active_InCells_Kernel :1085-1127
    return st.astype(np.int32)


def stopping_force_gden(fp, xp1, ityp, statu, cm, etab, stab, kpair, enable, mden):
    """ST_MOD_GDEN_KERNEL, LocalTempCtrlMeths/Stopping/MD_ST_Coupling_GPU.F90:431-536.  fp, xp1 (N,3); stab (NE,NK);
    kpair (NG,NG) 1-based [moving type, medium type].  Returns the updated FP.  (Per-atom Python loop: small cases only.)"""
    f = np.array(fp, dtype=np.float64)
    etab, stab = np.asarray(etab, dtype=np.float64), np.asarray(stab, dtype=np.float64)
    ng = len(cm)
    de = etab[1] - etab[0]
    deinv = 1.0 / de
    emin, emax = etab[0], etab[-1]
    for i in range(f.shape[0]):
        kk = int(ityp[i]) - 1
        if (statu[i] & STATU_ACTIVE) != STATU_ACTIVE or enable[kk] <= 0:
            continue
        vx, vy, vz = xp1[i]
        vv = vx * vx + vy * vy + vz * vz
        ek = 0.5 * cm[kk] * vv                                          # CM2(KK)*VV :509
        if not (emin <= ek <= emax):
            continue
        ik = int((ek - emin) * deinv)                                   # IK - 1 :512
        ff = 0.0
        for ig in range(ng):                                            # :516-520
            kp = int(kpair[kk][ig]) - 1
            sk = mden[ig] / de                                          # SK(IG) :487
            ff = ff + sk * ((ek - etab[ik]) * stab[ik + 1, kp] + (etab[ik + 1] - ek) * stab[ik, kp])
        v = np.sqrt(vv)
        f[i, 0] -= ff * vx / v                                          # :523-525
        f[i, 1] -= ff * vy / v
        f[i, 2] -= ff * vz / v
    return f


def stopping_force_lden(fp, xp1, ityp, statu, cm, etab, stab, kpair, enable, kvois, indi, nb_rm, dt=0.0):
    """ST_MOD_LDEN_KERNEL / ST_MOD_ELOSS_LDEN_KERNEL, LocalTempCtrlMeths/Stopping/MD_ST_Coupling_GPU.F90:604-717,1015-1132: the local
    density of medium type g around atom i is the number of its LIST neighbours of that type (:690-694) over LVOL(i,g) = 4 pi/3
    NB_RM(i,g)^3 (Reset_STMOD_DEV :412).  All per-atom arrays in ONE order (the list's): fp, xp1 (N,3); kvois (N,), indi (N, mx)
    1-based; nb_rm (NG,NG) in cm.  Returns (updated FP, ELOSS = FF |V| DT per atom).  (Per-atom Python loop: small cases only.)"""
    f = np.array(fp, dtype=np.float64)
    etab, stab = np.asarray(etab, dtype=np.float64), np.asarray(stab, dtype=np.float64)
    nb_rm = np.asarray(nb_rm, dtype=np.float64)
    ng = len(cm)
    de = etab[1] - etab[0]
    deinv = 1.0 / de
    emin, emax = etab[0], etab[-1]
    lv = 4.0 * np.pi / 3.0 * nb_rm ** 3
    eloss = np.zeros(f.shape[0])
    for i in range(f.shape[0]):
        kk = int(ityp[i]) - 1
        if (statu[i] & STATU_ACTIVE) != STATU_ACTIVE or enable[kk] <= 0:
            continue
        vx, vy, vz = xp1[i]
        vv = vx * vx + vy * vy + vz * vz
        ek = 0.5 * cm[kk] * vv
        iiw = int(kvois[i])
        if not (emin <= ek <= emax) or iiw <= 0:                        # :688
            continue
        den = np.zeros(ng)
        for w in range(iiw):                                            # :690-694
            den[int(ityp[int(indi[i, w]) - 1]) - 1] += 1.0
        ik = int((ek - emin) * deinv)
        ff = 0.0
        for ig in range(ng):                                            # :701-705
            kp = int(kpair[kk][ig]) - 1
            ilv = 1.0 / lv[kk, ig] / de                                 # ILV :669
            ff = ff + den[ig] * ilv * ((ek - etab[ik]) * stab[ik + 1, kp] + (etab[ik + 1] - ek) * stab[ik, kp])
        v = np.sqrt(vv)
        f[i, 0] -= ff * vx / v
        f[i, 1] -= ff * vy / v
        f[i, 2] -= ff * vz / v
        eloss[i] = ff * v * dt                                          # :1130
    return f, eloss


def activate_region_by_neighbours(statu, ityp, xp1, cm, kvois, indi, centpart=None, ekin_erg=None, extend=1, keep=False):
    """ActivateRegion_DEV -> ActiveByNeigbors1 -> ActiveByNeigbors0 (CP_BYNB_AR), CommonGPU/MD_ActiveRegion_GPU.F90:887-997: the seeds
    mark themselves and the atoms of their neighbour lists (MarkSeedNeighbore_Kernel :783-842), `extend` times with the marks as the
    next seeds; marked atoms are activated (Active_Marked_Kernel :428-471).  All arrays in the list's atom order; indi (N, mx) 1-based."""
    st = np.array(statu, dtype=np.int64)
    n = st.size
    if not keep:
        st &= ~STATU_ACTIVE                                            # :984-990
    if centpart is None and ekin_erg is None:                          # :909-910
        return st.astype(np.int32)
    inbox = (st & STATU_OUTOFBOX) != STATU_OUTOFBOX
    seed = np.zeros(n, dtype=np.int64)
    if centpart is not None:
        seed += (inbox & (np.asarray(centpart)[np.asarray(ityp) - 1] > 0)).astype(np.int64)
    if ekin_erg is not None:
        ek = 0.5 * np.asarray(cm)[np.asarray(ityp) - 1] * np.sum(np.asarray(xp1) ** 2, axis=1)
        seed += (inbox & (ek >= ekin_erg)).astype(np.int64)
    mark = seed.copy()                                                 # DevMakeCopy :932
    for _ in range(int(extend)):                                       # :935-951
        for i in np.nonzero(seed > 0)[0]:
            mark[i] = 1
            for w in range(int(kvois[i])):
                mark[int(indi[i, w]) - 1] += 1
        seed = mark.copy()
    st[mark > 0] |= STATU_ACTIVE
    return st.astype(np.int32)
