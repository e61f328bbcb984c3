"""Land-surface step on the GPU (shud_b200_land_create / land_step / land_get, shud_up_b200/csrc/shud_land.cuh)
against sequences dumped from the reference itself (tests/golden/<basin>.land.npz): every output of every step.
Tolerance: the arithmetic is the reference's operation for operation (-fmad=false); the only difference is the
device exp()/log() (<= 1-2 ulp against glibc's), so |got - ref| <= 1e-12 * max(|ref|, largest |ref| of the array)."""
import os

import numpy as np
import pytest

import oracle_lib
from shud_up_b200 import abi

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-12


def load(basin):
    mesh = dict(np.load(os.path.join(GOLD, f"{basin}.mesh.npz")))
    land = dict(np.load(os.path.join(GOLD, f"{basin}.land.npz")))
    case = dict(np.load(os.path.join(GOLD, f"{basin}.ic.npz")))
    snap = dict(mesh); snap.update(case)
    return snap, land


@pytest.mark.parametrize("basin", ["ccw", "qhh"])
def test_land_sequence_matches_reference(basin):
    from shud_up_b200.api import ShudRHS
    snap, land = load(basin)
    Ne = int(snap["Ne"][0])
    rhs = ShudRHS(snap)
    L, keep = abi.make_land(land)
    rhs.land_create(L)
    rhs.land_set_state(land["land_yEleSnow0"], land["land_yEleIS0"])
    nstep = land["lseq_t"].size
    for k, S, keep2 in abi.land_steps(land):
        rhs.land_step(S)
        got = rhs.land_get()
        assert rhs.check()[0] == 0
        for name in abi.LAND_OUT:
            ref = land["lseq_" + name].reshape(nstep, Ne)[k]
            scale = np.maximum(np.abs(ref), np.abs(ref).max())
            bad = np.abs(got[name] - ref) > RTOL * scale
            assert not bad.any(), (basin, k, name, int(bad.sum()), float(np.abs(got[name] - ref).max()))


def test_frozen_soil_factors_match_reference():
    """CRYOSPHERE = 1 over 800 hourly steps (both running-mean windows wrap), outputs every 40th step"""
    from shud_up_b200.api import ShudRHS
    snap, _ = load("ccw")
    land = dict(np.load(os.path.join(GOLD, "ccw.cryo.npz")))
    Ne = int(snap["Ne"][0])
    rhs = ShudRHS(snap)
    L, keep = abi.make_land(land)
    assert L.cryosphere == 1
    rhs.land_create(L)
    rhs.land_set_state(land["land_yEleSnow0"], land["land_yEleIS0"])
    kept = {int(v): j for j, v in enumerate(land["lseq_kept"])}
    nk = len(kept)
    for k, S, keep2 in abi.land_steps(land):
        rhs.land_step(S)
        if k not in kept:
            continue
        got = rhs.land_get()
        for name in ("fu_Surf", "fu_Sub", "t_temp", "yEleSnow", "qEleNetPrep"):
            ref = land["lseq_" + name].reshape(nk, Ne)[kept[k]]
            scale = np.maximum(np.abs(ref), np.abs(ref).max())
            bad = np.abs(got[name] - ref) > 1e-11 * np.maximum(scale, 1e-30)
            assert not bad.any(), (k, name, float(np.abs(got[name] - ref).max()))
    assert rhs.check()[0] == 0


def test_land_step_feeds_the_rhs():
    """after land_step the RHS runs on the device-made forcing: same ydot as with the reference's arrays uploaded
    through set_forcing (qhh: lake means of qPotEvap / qElePrep included)"""
    import torch
    from shud_up_b200.api import ShudRHS
    snap, land = load("qhh")
    Ne = int(snap["Ne"][0])
    nstep = land["lseq_t"].size
    g = lambda n, k: land["lseq_" + n].reshape(nstep, Ne)[k]
    k = 2
    # arm A: the reference's land-step outputs uploaded
    a = ShudRHS(snap)
    f = dict(snap)
    f.update(qEleNetPrep=g("qEleNetPrep", k), qPotEvap=g("qPotEvap", k), qPotTran=g("qPotTran", k), t_lai=g("t_lai", k),
             fu_Surf=g("fu_Surf", k), fu_Sub=g("fu_Sub", k), qElePrep=g("qElePrep", k))
    a.set_forcing(f, qEleE_IC=g("qEleE_IC", k))
    a.prime(snap["y"])
    ya = torch.from_numpy(np.ascontiguousarray(snap["y"])).pin_memory(); yda = torch.empty_like(ya).pin_memory()
    a.f(0.0, ya, yda)
    # arm B: the device land step (BC arrays etc. first through set_forcing, then overwritten in place)
    b = ShudRHS(snap)
    b.set_forcing(snap, qEleE_IC=snap["qEleE_IC_in"])
    L, keep = abi.make_land(land)
    b.land_create(L)
    b.land_set_state(land["land_yEleSnow0"], land["land_yEleIS0"])
    for kk, S, keep2 in abi.land_steps(land):
        b.land_step(S)
        if kk == k:
            break
    b.prime(snap["y"])
    ydb = torch.empty_like(ya).pin_memory()
    b.f(0.0, ya, ydb)
    ra, rb = yda.numpy(), ydb.numpy()
    scale = np.maximum(np.abs(ra), 1e-9)
    assert np.all(np.abs(ra - rb) <= 1e-10 * scale), float((np.abs(ra - rb) / scale).max())
