"""Regenerates the committed golden fixtures from the reference tree (run in the build
container only: /root/reference does not exist on the GPU box).

    python tests/golden/make_fixtures.py [/root/reference]

Outputs (small, committed):
  neb_gmd_react.npz / neb_gmd_product.npz
      the reference GPU build's own step-0 output for examples/NEB_Test (2000 W + 1 H,
      Bonny EAM1): per-atom type, position [LU], force [eV/LU], POT [eV] (= -EPOT)
      from examples/NEB_Test/GMD/{React,Product}P0000_0001.0000 (&BOXCFG18 columns
      TYPE, POS(3), VEL(3), STATU, FOR(3), POT, K.E., DISPLACE(3)).
  neb_gmd_react_quenched.npz
      the same run's output after 1000 damping steps (ReactP0000_0001.0001): position, force, POT.
  bonny_eam1_embd_rows.npz
      sampled rows of examples/use_ForceTableGen/EAM_WHeH_Bonny_JPCM26_2014.embd
      (Export_ForceTable output: RHO, F_k(RHO), dF_k/dRHO for the 9 ids).
  box/control text files used by those runs are copied verbatim (inputs, not source code).
  wangjun_fs_ww_pair_rows.npz
      sampled rows of examples/use_ForceTableGen/EM_TB_WANGJUN_W-HE_2010.pair (FS_TYPE export, Rmax = 10 A): the four
      columns of table id 1 = W-W (Ackland, Thetford): r*V, -r dV/dr, RHO [eV^2], -dRHO/dr.
  whhe_eam1_lspt_W.tar.xz, whhe_eam1_lspt_embd_rows.npz
      the tungsten functions of examples/NIST_Potentials/WHHe_EAM1_LSPT (F_W, rhoW, pWW .spt data files, verbatim, with a
      one-element index file) and sampled rows (RHO, F1, dF1/dRHO) of the reference's export WHHe_EAM1.lspt.embd.
  Cu1.eam.fs.setfl.xz
      examples/NIST_Potentials/Cu_EAM/Cu1.eam.fs.setfl (public NIST potential data file, an input), xz-compressed.
  cu1_setfl_table_rows.npz
      sampled rows of the reference's own import of that file, examples/NIST_Potentials/Cu_EAM/Cu1.eam.fs.setfl.pair
      and .embd (Export_ForceTable output, 10 / 9 significant digits): pins the setfl importer.
"""
import lzma
import os
import shutil
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def read_cfg18(path):
    rows = []
    with open(path) as f:
        started = False
        for line in f:
            if not started:
                if line.lstrip().upper().startswith("&TYPE "):
                    started = True
                continue
            p = line.split()
            if len(p) >= 16:
                rows.append([float(x) for x in p[:16]])
    a = np.array(rows)
    return dict(ityp=a[:, 0].astype(np.int32), pos=a[:, 1:4], vel=a[:, 4:7], statu=a[:, 7].astype(np.int32),
                force=a[:, 8:11], pot=a[:, 11], ekin=a[:, 12], dis=a[:, 13:16])


def main():
    neb = os.path.join(REF, "examples", "NEB_Test")
    for tag, name in (("react", "ReactP0000_0001.0000"), ("product", "ProductP0000_0001.0000")):
        d = read_cfg18(os.path.join(neb, "GMD", name))
        assert d["pos"].shape == (2001, 3)
        np.savez_compressed(os.path.join(HERE, "neb_gmd_%s.npz" % tag), ityp=d["ityp"], pos=d["pos"],
                            statu=d["statu"], force=d["force"], pot=d["pot"])
    # the same run after its QUICKDAMP section (1000 steps): a loose pin for the quench row -- the 2019 binary damped
    # dynamically (the file carries velocities), so only the energy level and the force drop are comparable
    d = read_cfg18(os.path.join(neb, "GMD", "ReactP0000_0001.0001"))
    np.savez_compressed(os.path.join(HERE, "neb_gmd_react_quenched.npz"), pos=d["pos"], force=d["force"], pot=d["pot"])
    for fn in ("W_2000_H1_EAM1_box.dat", "CtrlFile0K.dat"):
        shutil.copyfile(os.path.join(neb, fn), os.path.join(HERE, fn))
    with open(os.path.join(neb, "GMD", "thermP0000_0001")) as f:
        open(os.path.join(HERE, "neb_gmd_therm.txt"), "w").write(f.read())
    # header + first 10 rows of the reference-written &BOXCFG18 file, as text: pins the layout of inputs.write_config
    with open(os.path.join(neb, "GMD", "ReactP0000_0001.0000")) as f:
        open(os.path.join(HERE, "neb_ReactP0000_0001_head.txt"), "w").write("".join(f.readlines()[:42]))

    # the reference's own CPU-runnable example (BASELINE.json configs[0]): inputs only
    gmd = os.path.join(REF, "examples", "GMD_Test")
    for fn in ("W_2000_He1_EAM1_box.dat", "CtrlFile300K.dat", "W_2000_Tetra.cfg"):
        shutil.copyfile(os.path.join(gmd, fn), os.path.join(HERE, "gmd_" + fn))

    # the PARREP example's control file (BASELINE configs[3]): input only
    shutil.copyfile(os.path.join(REF, "examples", "PARREP_Test", "CtrlFile300K.dat"), os.path.join(HERE, "parrep_CtrlFile300K.dat"))

    # exported embedding table (10 significant digits)
    path = os.path.join(REF, "examples", "use_ForceTableGen", "EAM_WHeH_Bonny_JPCM26_2014.embd")
    rows = []
    with open(path) as f:
        for line in f:
            p = line.split()
            if len(p) == 20 and p[0].isdigit():
                rows.append([float(x) for x in p])
    a = np.array(rows)
    assert a.shape[0] == 10000, a.shape
    sel = np.unique(np.concatenate([np.arange(0, 64), np.arange(64, 10000, 97), [9998, 9999]]))
    np.savez_compressed(os.path.join(HERE, "bonny_eam1_embd_rows.npz"), index=a[sel, 0].astype(np.int32),
                        rho=a[sel, 1], f=a[sel, 2::2], df=a[sel, 3::2])
    # NIST setfl input + the reference's exported tables for it
    cu = os.path.join(REF, "examples", "NIST_Potentials", "Cu_EAM")
    with open(os.path.join(cu, "Cu1.eam.fs.setfl"), "rb") as f, lzma.open(os.path.join(HERE, "Cu1.eam.fs.setfl.xz"), "wb", preset=9) as g:
        g.write(f.read())

    def rows_of(path, ncol):
        out = []
        with open(path) as f:
            for line in f:
                p = line.split()
                if len(p) == ncol and p[0].isdigit():
                    out.append([float(x) for x in p])
        return np.array(out)

    pr = rows_of(os.path.join(cu, "Cu1.eam.fs.setfl.pair"), 6)
    em = rows_of(os.path.join(cu, "Cu1.eam.fs.setfl.embd"), 4)
    assert pr.shape[0] == 10000 and em.shape[0] == 10000
    sel = np.unique(np.concatenate([np.arange(0, 64), np.arange(64, 10000, 23), np.arange(1270, 1340), np.arange(9950, 10000)]))
    np.savez_compressed(os.path.join(HERE, "cu1_setfl_table_rows.npz"), index=pr[sel, 0].astype(np.int32), r=pr[sel, 1],
                        pair=pr[sel, 2:6], rho=em[sel, 1], embd=em[sel, 2:4])
    # exported Finnis-Sinclair tables: columns of table id 1 (W-W, Ackland & Thetford) of the 5-id library
    pr = rows_of(os.path.join(REF, "examples", "use_ForceTableGen", "EM_TB_WANGJUN_W-HE_2010.pair"), 22)
    assert pr.shape[0] == 10000
    sel = np.unique(np.concatenate([np.arange(0, 32), np.arange(32, 10000, 19), np.arange(5200, 5260), np.arange(6600, 6660)]))
    np.savez_compressed(os.path.join(HERE, "wangjun_fs_ww_pair_rows.npz"), index=pr[sel, 0].astype(np.int32), r=pr[sel, 1],
                        pair=pr[sel, 2:6])
    # lspt library: the W part of examples/NIST_Potentials/WHHe_EAM1_LSPT (index reduced to one element, the three data
    # files it names verbatim) and sampled rows of the reference's exported embedding table for the full library
    import io
    import tarfile
    ls = os.path.join(REF, "examples", "NIST_Potentials", "WHHe_EAM1_LSPT")
    index = ('!  W part of WHHe_EAM1.lspt (examples/NIST_Potentials/WHHe_EAM1_LSPT), same keywords\n'
             '&LSPT ----------------------------\n'
             '   NGROUP  the number of groups =  1\n'
             '   ATOMSYMBOL  group#1 = "W"\n'
             '   F_RHO #1 filename =     "WHHe-EAM1-F_W.spt"\n'
             '   RHO_R #1 filename =     "WHHe-EAM1-rhoW.spt"\n'
             '   VR  (1,1)  filename  =  "WHHe-EAM1-pWW.spt"\n')
    with tarfile.open(os.path.join(HERE, "whhe_eam1_lspt_W.tar.xz"), "w:xz", preset=9) as tf:
        ti = tarfile.TarInfo("W_EAM1.lspt"); data = index.encode(); ti.size = len(data)
        tf.addfile(ti, io.BytesIO(data))
        for fn in ("WHHe-EAM1-F_W.spt", "WHHe-EAM1-rhoW.spt", "WHHe-EAM1-pWW.spt"):
            tf.add(os.path.join(ls, fn), arcname=fn)
    em = rows_of(os.path.join(ls, "WHHe_EAM1.lspt.embd"), 20)
    assert em.shape[0] == 10000
    sel = np.unique(np.concatenate([np.arange(0, 64), np.arange(64, 10000, 23), np.arange(9950, 10000)]))
    np.savez_compressed(os.path.join(HERE, "whhe_eam1_lspt_embd_rows.npz"), index=em[sel, 0].astype(np.int32), rho=em[sel, 1],
                        f=em[sel, 2], df=em[sel, 3])
    print("fixtures written to", HERE)


if __name__ == "__main__":
    main()
