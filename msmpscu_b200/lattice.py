"""Deterministic synthetic configurations for the BASELINE configs:
bcc lattices at (i+1/4) and (i+3/4) a0 with small uniform displacements and Maxwell velocities."""
import numpy as np

from .constants import CP_A2CM, CP_AU2G, CP_KB


def bcc_box(ncell, a0_ang, seed, disp_lu=0.02, temp_k=600.0, mass_amu=183.84, nbox=1):
    """ncell: (nx,ny,nz) bcc unit cells.  Returns dict with positions/velocities in CGS, box arrays.
    Positions: (i+1/4, j+1/4, k+1/4) a0 and (+1/2) -- keeps atoms off the cell faces -- shifted so the
    box is [-L/2, L/2).  Velocities: Maxwell at temp_k, zero centre-of-mass momentum per box."""
    nx, ny, nz = (int(v) for v in ncell)
    rr = a0_ang * CP_A2CM
    rng = np.random.default_rng(seed)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    base = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64)
    lat = np.concatenate([base + 0.25, base + 0.75], axis=0)
    napb = lat.shape[0]
    size_lu = np.array([nx, ny, nz], dtype=np.float64)
    xs, vs = [], []
    m = mass_amu * CP_AU2G
    for _ in range(nbox):
        x = lat + rng.uniform(-disp_lu, disp_lu, size=lat.shape) - 0.5 * size_lu
        v = rng.normal(0.0, np.sqrt(CP_KB * temp_k / m), size=lat.shape)
        v -= v.mean(axis=0)
        xs.append(x * rr)
        vs.append(v)
    zl = size_lu * rr
    return dict(xp=np.concatenate(xs), xp1=np.concatenate(vs), napb=napb, nbox=nbox, zl=zl, boxlow=-0.5 * zl, rr=rr,
                ityp=np.ones(napb * nbox, dtype=np.int32), statu=np.ones(napb * nbox, dtype=np.int32),
                mass=np.array([m]))


def fcc_box(ncell, a0_ang, seed, disp_lu=0.02, temp_k=600.0, mass_amu=63.546, nbox=1):
    """fcc twin of bcc_box: four atoms per cell at (1/4,1/4,1/4) + {0, (1/2,1/2,0), (1/2,0,1/2), (0,1/2,1/2)} a0."""
    nx, ny, nz = (int(v) for v in ncell)
    rr = a0_ang * CP_A2CM
    rng = np.random.default_rng(seed)
    i, j, k = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    base = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64) + 0.25
    lat = np.concatenate([base, base + [0.5, 0.5, 0.0], base + [0.5, 0.0, 0.5], base + [0.0, 0.5, 0.5]], axis=0)
    napb = lat.shape[0]
    size_lu = np.array([nx, ny, nz], dtype=np.float64)
    xs, vs = [], []
    m = mass_amu * CP_AU2G
    for _ in range(nbox):
        x = lat + rng.uniform(-disp_lu, disp_lu, size=lat.shape) - 0.5 * size_lu
        v = rng.normal(0.0, np.sqrt(CP_KB * temp_k / m), size=lat.shape)
        v -= v.mean(axis=0)
        xs.append(x * rr)
        vs.append(v)
    zl = size_lu * rr
    return dict(xp=np.concatenate(xs), xp1=np.concatenate(vs), napb=napb, nbox=nbox, zl=zl, boxlow=-0.5 * zl, rr=rr,
                ityp=np.ones(napb * nbox, dtype=np.int32), statu=np.ones(napb * nbox, dtype=np.int32),
                mass=np.array([m]))
