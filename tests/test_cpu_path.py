"""The reference's CPU path in the oracle: CPU lists (full / Newton-3 half), Newton-3 force, step loop.
CPU only.  These are the routines bench.py times as the CPU baseline; here they are pinned against each other and against
the device-rule oracle (which the NEB goldens pin, tests/test_oracle_golden.py)."""
import numpy as np
import pytest

import util
from oracle import pyorc as O


@pytest.fixture(scope="module")
def case():
    return util.bcc_case((6, 6, 6), seed=99)


def test_half_list_holds_every_pair_once(case):
    c = case
    full = util.oracle_cpu_md(O, c, half=False)
    half = util.oracle_cpu_md(O, c, half=True)
    assert full.rebuild() == 0 and half.rebuild() == 0
    kf, kh = full.kvois(), half.kvois()
    assert kf.sum() == 2 * kh.sum()          # Cal_NeighboreListC stores a pair once, ...2C twice
    assert kf.min() > 0


def test_newton3_force_matches_directed_force(case):
    """CALFORCE_FS_Force_Table (half list, reaction added to J) == CALFORCE_FS_Force_Table2 (every directed pair)"""
    c = case
    full = util.oracle_cpu_md(O, c, half=False)
    half = util.oracle_cpu_md(O, c, half=True)
    full.rebuild(); half.rebuild()
    full.force(epot=True); half.force(epot=True)
    _, _, f2, e2 = full.get(epot=True)
    _, _, f1, e1 = half.get(epot=True)
    assert util.atom_relerr(f1, f2) < 1e-11
    assert util.atom_relerr(e1, e2) < 1e-12


def test_cpu_path_matches_device_rule_oracle(case):
    """same forces as the cell-sorted device-rule driver (fp32 '<=' membership vs fp64 '<': no pair sits on the list edge,
    and every pair inside RU is listed by both)"""
    c = case
    cpu = util.oracle_cpu_md(O, c)
    cpu.rebuild(); cpu.force(epot=True)
    _, _, f, e = cpu.get(epot=True)
    md = util.oracle_md(O, c)
    md.rebuild(); md.force(); md.epot()
    g = md.get()
    assert util.atom_relerr(f, g["fp"]) < 1e-11
    assert util.atom_relerr(e, g["epot"]) < 1e-12


def test_cpu_run_loop_tracks_stepwise_driver(case):
    """orc_cpu_run (loop in C, CPU list) against orc_md_step x n (device list rule): same trajectory"""
    c = case
    h = 0.5e-15
    cpu = util.oracle_cpu_md(O, c)
    cpu.rebuild(); cpu.force()
    md = util.oracle_md(O, c)
    md.rebuild(); md.force()
    assert cpu.run(0, 12, 1, 10, h) == 2      # rebuilds at ITIME = 1 and 11
    for it in range(12):
        md.step(it, 1, 10, h)
    x, v, f, _ = cpu.get()
    g = md.get()
    assert np.max(np.abs(x - g["xp"])) < 1e-12 * np.max(np.abs(g["xp"]))
    assert util.atom_relerr(v, g["xp1"]) < 1e-9
    assert util.atom_relerr(f, g["fp"]) < 1e-8


def test_fast_build_agrees_with_parity_build(case):
    """liborc_fast.so (-O3 -march=native, FMA contraction) is the timing build of the same source"""
    c = case
    a = util.oracle_cpu_md(O, c, fast=False)
    b = util.oracle_cpu_md(O, c, fast=True)
    a.rebuild(); b.rebuild(); a.force(); b.force()
    assert np.array_equal(a.kvois(), b.kvois())
    assert util.atom_relerr(b.get()[2], a.get()[2]) < 1e-11
    h0 = a.harmil()
    assert abs(b.harmil() - h0) <= 1e-12 * abs(h0)
