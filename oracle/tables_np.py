"""TEST INFRASTRUCTURE ONLY (see oracle/mdpscu_oracle.h): NumPy/SciPy restatement of the reference's external
force-table path, independent of the product's C++ (msmpscu_b200/csrc/host_tables_io.cpp).

  setfl_tables   Potentials/EAM_NIST/Filedatas_Func_Setfl.F90:153-296 (reader), :297-312 (DBINT4 with IBCL=IBCR=2,
                 FBCL=FBCR=0 = cubic spline with zero end curvature -> scipy CubicSpline(bc_type="natural")),
                 :320-462 (range rules, units), NIST_ForceTable.F90:332-398 (Generate_NIST_ForceTalbe),
                 Common/MD_TypeDef_ForceTable.F90:890-1056 (grids)
  export_columns Common/MD_TypeDef_ForceTable.F90:1388-1397, 1421-1441: the numbers Export_ForceTable prints

Pinned by tests/test_tables_io.py on the reference's own Cu1.eam.fs.setfl.pair/.embd (10 / 9 digits).
"""
import numpy as np
from scipy.interpolate import CubicSpline

EVERG = 1.60219e-12
A2CM = 1.0e-8
CM2A = 1.0e8


def read_setfl(text):
    lines = text.split("\n")
    ne = int(lines[3].split()[0])
    names = lines[3].split()[1:1 + ne]
    p = lines[4].split()
    nrho, drho, nr, dr, cutoff = int(p[0]), float(p[1]), int(p[2]), float(p[3]), float(p[4])
    toks = " ".join(lines[5:]).split()
    pos = 0

    def block(n):
        nonlocal pos
        raw = toks[pos:pos + n]
        pos += n
        if raw[0].upper() in ("INF", "NAN"):  # :213-231
            raw = [raw[1]] + raw[1:]
        return np.array([float(v) for v in raw])

    el = []
    for _ in range(ne):
        z, mass, alat, lat = int(toks[pos]), float(toks[pos + 1]), float(toks[pos + 2]), toks[pos + 3]
        pos += 4
        el.append(dict(z=z, mass=mass, alat=alat, lattice=lat, frho=block(nrho), rhor=block(nr)))
    v = {}
    for i in range(ne):
        for j in range(i + 1):
            v[(i, j)] = block(nr)
    return dict(ne=ne, names=names, nrho=nrho, drho=drho, nr=nr, dr=dr, cutoff=cutoff, el=el, v=v)


def _ranged(sp, lo, hi, t, hold_above):
    """spline inside [lo,hi]; end value with zero slope below; zero (or the end value) above"""
    tc = np.clip(t, lo, hi)
    f, df = sp(tc), sp(tc, 1)
    below, above = t < lo, t > hi
    df = np.where(below | above, 0.0, df)
    if not hold_above:
        f = np.where(above, 0.0, f)
    return f, df


def setfl_tables(text, ntab, nembd, rmax=None):
    s = read_setfl(text)
    ne = s["ne"]
    rho_x = np.arange(s["nrho"]) * s["drho"]
    r_x = np.arange(s["nr"]) * (s["cutoff"] / s["nr"])  # :184-187
    rmax = s["cutoff"] * A2CM if rmax is None else rmax
    csi = ntab / np.sqrt(rmax)
    csiv = 1.0 / csi
    rhod = (s["nrho"] * s["drho"]) / nembd
    r = (np.arange(1, ntab + 1) * csiv) ** 2
    ra = r * CM2A
    nk = ne * ne
    out = {k: np.zeros((nk, ntab)) for k in ("potr", "fpotr", "potb", "fpotb")}
    out["fembd"], out["dfembd"] = np.zeros((nk, nembd)), np.zeros((nk, nembd))
    for it in range(1, nk + 1):
        i = (it - 1) // ne + 1
        j = it - (i - 1) * ne
        iv, jv = (max(i, j), min(i, j))
        spv = CubicSpline(r_x, s["v"][(iv - 1, jv - 1)], bc_type="natural")
        f, df = _ranged(spv, r_x[0], r_x[-1], ra, False)
        pot = f / ra
        fpot = (df - pot) / ra
        out["potr"][it - 1] = 0.5 * pot * EVERG * r
        out["fpotr"][it - 1] = -1.0 * fpot * EVERG * CM2A * r
        frho = s["el"][i - 1]["frho"]
        if not (frho.max() == 0.0 and frho.min() == 0.0):
            spq = CubicSpline(r_x, s["el"][j - 1]["rhor"], bc_type="natural")
            f, df = _ranged(spq, r_x[0], r_x[-1], ra, False)
            out["potb"][it - 1] = f
            out["fpotb"][it - 1] = -1.0 * df * CM2A
        if i == j:
            spf = CubicSpline(rho_x, frho, bc_type="natural")
            f, df = _ranged(spf, rho_x[0], rho_x[-1], np.arange(nembd) * rhod, True)
            out["fembd"][it - 1] = f * EVERG
            out["dfembd"][it - 1] = df * EVERG
    out.update(csi=csi, rhod=rhod, rmax=rmax, ne=ne, nkind=nk, setfl=s)
    return out


def export_columns(t, fs=False):
    """what Export_ForceTable prints: pair (r*V [eV A], -r dV/dr [eV], RHO, -dRHO/dr [/A]) and embd (F [eV], dF/dRHO)"""
    ergev = 1.0 / EVERG
    rhounit = ergev * ergev if fs else 1.0
    pair = np.stack([t["potr"] * 2.0 * ergev * CM2A, t["fpotr"] * ergev, t["potb"] * rhounit, t["fpotb"] * rhounit * A2CM], axis=-1)
    embd = np.stack([t["fembd"] * ergev, t["dfembd"]], axis=-1)
    return pair, embd


def lspt_tables_W(dirpath, ntab, nembd):
    """The one-element lspt library of the fixture (F_W, rhoW, pWW): Filedatas_Func_Lspt.F90:79-300 (reader), :488-541
    (NN_Spline: the files hold V, not r*V), NIST_ForceTable.F90:332-398.  RHOMX = max rho of F_W, Rmax = max r of pWW."""
    import os
    ld = lambda fn: np.loadtxt(os.path.join(dirpath, fn))
    fw, rw, vw = ld("WHHe-EAM1-F_W.spt"), ld("WHHe-EAM1-rhoW.spt"), ld("WHHe-EAM1-pWW.spt")
    rmax = vw[:, 0].max() * A2CM
    rhomx = fw[:, 0].max()
    csi = ntab / np.sqrt(rmax)
    r = (np.arange(1, ntab + 1) / csi) ** 2
    ra = r * CM2A
    spv = CubicSpline(vw[:, 0], vw[:, 1], bc_type="natural")
    f, df = _ranged(spv, vw[0, 0], vw[-1, 0], ra, False)
    out = dict(potr=(0.5 * f * EVERG * r)[None], fpotr=(-1.0 * df * EVERG * CM2A * r)[None])
    spq = CubicSpline(rw[:, 0], rw[:, 1], bc_type="natural")
    f, df = _ranged(spq, rw[0, 0], rw[-1, 0], ra, False)
    out.update(potb=f[None], fpotb=(-1.0 * df * CM2A)[None])
    rhod = rhomx / nembd
    spf = CubicSpline(fw[:, 0], fw[:, 1], bc_type="natural")
    f, df = _ranged(spf, fw[0, 0], fw[-1, 0], np.arange(nembd) * rhod, True)
    out.update(fembd=(f * EVERG)[None], dfembd=(df * EVERG)[None], csi=csi, rhod=rhod, rmax=rmax)
    return out
