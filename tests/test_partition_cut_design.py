"""Design check for cut river trees (DESIGN.md section 6, next step of the multi-GPU path): with the closure of
partition.cut_river_closure - own cells, halo cells, replica cells, own reaches, halo reaches - an ordinary mesh made of
exactly those cells and reaches reproduces, with the CPU oracle, the single-domain ydot of every own cell and own reach
bit for bit, for plain Hilbert-range partitions that cut the river network anywhere.  Nothing crosses a partition
boundary but states: every flux is evaluated where it is needed (owner-computes)."""
import numpy as np
import pytest

import oracle_lib
from shud_up_b200 import partition


def _hilbert_ranges(mesh, nparts):
    key = partition._hilbert_key(mesh["ele_x"], mesh["ele_y"])
    order = np.argsort(key, kind="stable")
    part = np.empty(order.size, dtype=np.int64)
    part[order] = (np.arange(order.size) * nparts) // order.size
    return part


def _closed_mesh(mesh, cl):
    Ne, Nr = int(mesh["Ne"][0]), int(mesh["Nr"][0])
    cells = np.concatenate([cl["own"], cl["halo"], cl["replica"][~np.isin(cl["replica"], cl["halo"])]])
    rivs = np.concatenate([cl["riv_own"], cl["riv_halo"]])
    cnew = np.zeros(Ne, dtype=np.int64); cnew[cells] = np.arange(1, cells.size + 1)
    rnew = np.zeros(Nr, dtype=np.int64); rnew[rivs] = np.arange(1, rivs.size + 1)
    out = {}
    for k, v in mesh.items():
        v = np.asarray(v)
        if k in partition.EDGE_KEYS:
            out[k] = np.ascontiguousarray(v.reshape(3, Ne)[:, cells]).ravel()
        elif k in ("ele_nabr", "ele_lakenabr"):
            continue
        elif (k.startswith("ele_") or k in partition.CELL_DYN) and v.ndim == 1 and v.shape[0] == Ne:
            out[k] = v[cells]
        elif k.startswith("riv_") and v.ndim == 1 and v.shape[0] == Nr:
            out[k] = v[rivs]
        else:
            out[k] = v
    nab = np.asarray(mesh["ele_nabr"]).reshape(3, Ne)[:, cells]
    nab = np.where(nab > 0, cnew[np.maximum(nab - 1, 0)], 0)
    nab[:, cl["own"].size:] = 0                      # halo and replica cells: no edges evaluated for them
    out["ele_nabr"] = nab.astype(np.int32).ravel()
    out["ele_lakenabr"] = np.zeros(3 * cells.size, dtype=np.int32)
    down = np.asarray(mesh["riv_down"])[rivs]
    dn_new = np.where(down > 0, rnew[np.maximum(down - 1, 0)], down)
    dn_new = np.where((down > 0) & (dn_new == 0), -3, dn_new)   # a halo reach whose downstream reach is not held
    out["riv_down"] = dn_new.astype(np.int32)
    sg = cl["seg"]
    out["seg_iEle"] = cnew[np.asarray(mesh["seg_iEle"])[sg] - 1].astype(np.int32)
    out["seg_iRiv"] = rnew[np.asarray(mesh["seg_iRiv"])[sg] - 1].astype(np.int32)
    out["seg_length"] = np.asarray(mesh["seg_length"])[sg]
    out["seg_Cwr"] = np.asarray(mesh["seg_Cwr"])[sg]
    y = np.asarray(mesh["y"])
    out["y"] = np.concatenate([y[cells], y[Ne + cells], y[2 * Ne + cells], y[3 * Ne + rivs]])
    for k, n in (("Ne", cells.size), ("Nr", rivs.size), ("Ns", sg.size), ("Nl", 0), ("lakeon", 0)):
        out[k] = np.array([n], dtype=np.int32)
    out["lake_zmin"] = np.zeros(0); out["lake_NumEleLake"] = np.zeros(0, dtype=np.int32)
    out["lake_bathy_ptr"] = np.zeros(1, dtype=np.int32); out["lake_bathy_yi"] = np.zeros(0); out["lake_bathy_ai"] = np.zeros(0)
    return out, cells, rivs


@pytest.mark.parametrize("basin,case,nparts", [("ccw", "rand1", 2), ("ccw", "rand1", 4), ("heihe", "rand3", 3)])
def test_closure_reproduces_single_domain_with_cut_rivers(basin, case, nparts):
    mesh = oracle_lib.load_case(basin, case)
    assert int(mesh["Nl"][0]) == 0
    Ne, Nr = int(mesh["Ne"][0]), int(mesh["Nr"][0])
    ref = oracle_lib.oracle_rhs(mesh, want_diag=False)["ydot"]
    part = _hilbert_ranges(mesh, nparts)
    owners = np.zeros(Nr, dtype=int)
    n_cut = 0
    for p in range(nparts):
        cl = partition.cut_river_closure(mesh, part, p)
        owners[cl["riv_own"]] += 1
        n_cut += cl["replica"].size + cl["riv_halo"].size
        sub, cells, rivs = _closed_mesh(mesh, cl)
        out = oracle_lib.oracle_rhs(sub, want_diag=False)["ydot"]
        nc, nown, nro = cells.size, cl["own"].size, cl["riv_own"].size
        for b in range(3):
            assert np.array_equal(out[b * nc:b * nc + nown], ref[b * Ne + cl["own"]]), (p, b)
        assert np.array_equal(out[3 * nc:3 * nc + nro], ref[3 * Ne + cl["riv_own"]]), p
    assert np.all(owners == 1)     # every reach has exactly one owner
    assert n_cut > 0               # the partition really cuts the river network
