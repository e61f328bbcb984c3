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


def test_oracle_frozen_soil_factors_bit_exact():
    """CRYOSPHERE = 1: the running means of the daily mean temperature over 7 / 28 days (AccTemperature.hpp) and
    the factors fu_Surf / fu_Sub they give, 800 hourly steps (both windows wrap), outputs every 40th step"""
    mesh = dict(np.load(os.path.join(GOLD, "ccw.mesh.npz")))
    land = dict(np.load(os.path.join(GOLD, "ccw.cryo.npz")))
    Ne = int(mesh["Ne"][0])
    assert int(land["land_cs"][2]) == 1 and land["lseq_t"].size == 800
    res = oracle_lib.oracle_land_seq(mesh, land)
    nkept = land["lseq_kept"].size
    for name in ("fu_Surf", "fu_Sub", "t_temp", "yEleSnow", "qEleNetPrep"):
        ref = land["lseq_" + name].reshape(nkept, Ne)
        assert np.array_equal(res[name], ref), name
    fs, fb = land["lseq_fu_Surf"], land["lseq_fu_Sub"]
    assert fs.min() == 0.0 and fs.max() == 1.0 and ((fs > 0) & (fs < 1)).any() and ((fb > 0) & (fb < 1)).any()


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
