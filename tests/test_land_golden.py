"""Land-surface step: the oracle restatement of Model_Data::updateforcing /
tReadForcing / ET (oracle/shud_oracle.c: shud_oracle_land_step) against sequences dumped from the reference itself
(tests/golden/<basin>.land.npz, made by tools/make_golden.py with oracle/ref_driver.cpp --land-seq).
Same compiler, same libm on both sides: the bar is bit-exact, every output, every step."""
import os

import numpy as np
import pytest

import oracle_lib
from shud_up_b200 import abi

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(basin):
    mesh = dict(np.load(os.path.join(GOLD, f"{basin}.mesh.npz")))
    land = dict(np.load(os.path.join(GOLD, f"{basin}.land.npz")))
    return mesh, land


@pytest.mark.parametrize("basin", ["ccw", "qhh"])
def test_oracle_land_sequence_bit_exact(basin):
    mesh, land = load(basin)
    Ne = int(mesh["Ne"][0])
    res = oracle_lib.oracle_land_seq(mesh, land)
    nstep = land["lseq_t"].size
    for name in abi.LAND_OUT:
        ref = land["lseq_" + name].reshape(nstep, Ne)
        assert np.array_equal(res[name], ref), name


def test_sequences_cover_the_branches():
    """the fixtures exercise what they claim: rain, snow accumulation, melt, interception, night and day, terrain
    factors above and below 1, lake cells"""
    mesh, land = load("ccw")
    Ne = int(mesh["Ne"][0])
    g = lambda n: land["lseq_" + n].reshape(-1, Ne)
    assert (g("qElePrep") > 0).any() and (g("qElePrep") == 0).any()
    snow = g("yEleSnow")
    assert (np.diff(snow, axis=0) > 0).any() and (np.diff(snow, axis=0) < 0).any()
    assert (g("yEleIS") > 0).any() and (g("qEleE_IC") > 0).any()
    f = g("rn_factor")
    assert (f == 0).any() and (f > 1).any() and ((f > 0) & (f < 1)).any()
    assert int(land["land_cs"][1]) == 1  # TERRAIN_RADIATION on
    meshq, landq = load("qhh")
    assert (meshq["ele_iLake"] > 0).sum() > 0
    assert np.unique(landq["lseq_tsr_den"]).size < landq["lseq_tsr_den"].size  # steps sharing a forcing interval
