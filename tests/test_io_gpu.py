"""Ingest / checkpoint through the device path: a context created from the binary mesh container gives the same
ydot as one created from the arrays; shud_b200_write_ic (device vector, device order, land-step buckets) writes the
same file as the host formatter fed with the reference-order state."""
import os

import numpy as np
import pytest

import oracle_lib
from shud_up_b200 import abi, api

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _ydot(rhs, snap):
    import torch
    rhs.set_forcing(snap, qEleE_IC=snap["qEleE_IC_in"])
    rhs.prime(snap["y"])
    y = torch.from_numpy(np.ascontiguousarray(snap["y"])).pin_memory()
    yd = torch.empty_like(y).pin_memory()
    rhs.f(0.0, y, yd)
    return yd.numpy().copy()


def test_context_from_container_and_checkpoint(tmp_path):
    import torch
    snap = oracle_lib.load_case("qhh", "rand4")
    a = api.ShudRHS(snap)
    ref = _ydot(a, snap)
    path = tmp_path / "qhh.shudb200"
    api.mesh_save(path, snap)
    ld = api.LoadedMesh(path)
    b = api.ShudRHS(ld)
    assert np.array_equal(_ydot(b, snap), ref)
    # checkpoint: device-order vector + land buckets -> the reference's text format
    land = dict(np.load(os.path.join(GOLD, "qhh.land.npz")))
    L, keep = abi.make_land(land)
    b.land_create(L)
    rng = np.random.default_rng(5)
    snow, ics = rng.uniform(0, 0.3, b.Ne), rng.uniform(0, 1e-3, b.Ne)
    b.land_set_state(snow, ics)
    st = b.torch_stream()
    with torch.cuda.stream(st):
        yr = torch.from_numpy(np.ascontiguousarray(snap["y"])).cuda()
        yd = torch.empty_like(yr)
        b.to_device_order(yr, yd)
    st.synchronize()
    f_dev, f_host = tmp_path / "dev.ic", tmp_path / "host.ic"
    b.write_ic(f_dev, 2880.0, yd)
    api.format_ic(f_host, 2880.0, b.Ne, b.Nr, b.Nl, snap["y"], ics, snow)
    assert open(f_dev, "rb").read() == open(f_host, "rb").read()
    b.close(); ld.close(); a.close()


def test_checkpoint_of_a_mesh_with_head_and_stage_boundary_conditions(tmp_path):
    """the reference prints the state summary() left behind (shud.cpp:137-157, MD_update.cpp:190-216): cells with
    iBC > 0 show their boundary head, reaches with BC > 0 their boundary stage - not the solver's frozen rows"""
    import torch
    snap = oracle_lib.load_case("ccw", "mut2")            # mutations ebc / rbc: head-BC cells and stage-BC reaches
    Ne, Nr = int(snap["Ne"][0]), int(snap["Nr"][0])
    hb, rb = np.asarray(snap["ele_iBC"]) > 0, np.asarray(snap["riv_BC"]) > 0
    assert hb.any() and rb.any()
    rhs = api.ShudRHS(snap)
    rhs.set_forcing(snap, qEleE_IC=snap["qEleE_IC_in"])
    st = rhs.torch_stream()
    with torch.cuda.stream(st):
        yr = torch.from_numpy(np.ascontiguousarray(snap["y"])).cuda()
        yd = torch.empty_like(yr)
        rhs.to_device_order(yr, yd)
    st.synchronize()
    want = np.array(snap["y"], copy=True)
    want[2 * Ne:3 * Ne][hb] = np.asarray(snap["ele_yBC"])[hb]
    want[3 * Ne:3 * Ne + Nr][rb] = np.asarray(snap["riv_yBC"])[rb]
    assert not np.array_equal(want, snap["y"])
    assert np.array_equal(rhs.summary(yd), want)
    f_dev, f_host = tmp_path / "dev.ic", tmp_path / "host.ic"
    rhs.write_ic(f_dev, 1440.0, yd)
    api.format_ic(f_host, 1440.0, Ne, Nr, 0, want)
    assert open(f_dev, "rb").read() == open(f_host, "rb").read()
    rhs.close()
