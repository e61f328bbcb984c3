"""GPU parity of the CUDA RHS (through the C ABI) against the CPU oracle and the golden
snapshots of the unmodified reference f().  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest

import oracle_lib
import parity
from shud_up_b200 import abi

pytestmark = pytest.mark.gpu

CASES = [("ccw", "ic"), ("ccw", "rand1"), ("ccw", "mut2"), ("heihe", "ic"), ("heihe", "rand3"),
         ("qhh", "ic"), ("qhh", "rand4"), ("qhh", "mut5"), ("qhh", "lakes6")]
# flux arrays compared at 1e-12 relative on their own (no cancellation inside them)
FLUX = ["qEleInfil", "qEleExfil", "qEleRecharge", "qEs", "qEu", "qEg", "qTu", "qTg", "qEleTrans", "qEleEvapo",
        "qEleETA", "u_effKH", "u_satn", "QeleSurf", "QeleSub", "QsegSurf", "QsegSub", "QrivDown", "y2LakeArea",
        "qLakeEvap", "qLakePrcp", "QLakeRivIn"]
# sums of the above: compared relative to the sum of |addends|
SUMS = ["QeleSurfTot", "QeleSubTot", "Qe2r_Surf", "Qe2r_Sub", "QrivSurf", "QrivSub", "QrivUp", "QLakeSurf", "QLakeSub"]


def _run_gpu(snap, diag=True):
    import torch
    from shud_up_b200.api import ShudRHS
    rhs = ShudRHS(snap)
    rhs.set_forcing(snap, qEleE_IC=snap["qEleE_IC_in"])
    rhs.set_carried(snap["ele_u_satn"])
    dev = torch.device("cuda:0")
    with torch.cuda.stream(rhs.torch_stream()):
        y_ref = torch.from_numpy(np.ascontiguousarray(snap["y"])).to(dev)
        y_dev = torch.empty_like(y_ref)
        yd_dev = torch.full_like(y_ref, float("nan"))
        yd_ref = torch.empty_like(y_ref)
        rhs.to_device_order(y_ref, y_dev)
        rhs.f_dev(0.0, y_dev, yd_dev, diag=diag)
        rhs.from_device_order(yd_dev, yd_ref)
    code, where = rhs.check()
    out = {"ydot": yd_ref.cpu().numpy(), "code": code, "where": where}
    if diag:
        out.update(rhs.get_diag())
    s, e = rhs.get_carried()
    out["u_satn_out"], out["qEleE_IC_out"] = s, e
    return rhs, out


@pytest.mark.parametrize("basin,case", CASES)
def test_ydot_and_fluxes_match_oracle(basin, case):
    snap = oracle_lib.load_case(basin, case)
    ref = oracle_lib.oracle_rhs(snap)
    assert ref["err"] == 0
    rhs, got = _run_gpu(snap)
    assert got["code"] == 0, got
    Ne = rhs.Ne
    scale = parity.ydot_scale(snap, ref)
    bad = parity.mismatches(got["ydot"], ref["ydot"], scale)
    # the golden file holds the reference's own ydot; the oracle equals it bit for bit (test_oracle_golden)
    assert np.array_equal(ref["ydot"], snap["ydot"])
    assert bad.size == 0, (f"{basin}.{case}: {bad.size} ydot components outside 1e-12*max(|ydot|,sum|terms|); "
                           f"first {bad[:5]} got {got['ydot'][bad[:5]]} ref {ref['ydot'][bad[:5]]}")
    lake = snap["ele_iLake"] > 0
    problems = []
    for name in FLUX:
        if ref[name].size == 0:
            continue
        b = parity.mismatches(got[name], ref[name])
        if b.size:
            problems.append((name, b.size, b[:3].tolist()))
    sums = {"QeleSurfTot": np.abs(ref["QeleSurf"]).reshape(3, Ne).sum(0) + np.abs(ref["Qe2r_Surf"]),
            "QeleSubTot": np.abs(ref["QeleSub"]).reshape(3, Ne).sum(0) + np.abs(ref["Qe2r_Sub"])}
    for name in SUMS:
        if ref[name].size == 0:
            continue
        sc = sums.get(name)
        if sc is None:  # sums over segments / reaches / bank edges: scale by the largest addend-sum available
            sc = np.full(ref[name].shape, 0.0)
            if name.startswith("Qe2r") or name.startswith("Qriv") or name.startswith("QLake"):
                sc = np.full(ref[name].shape, np.abs(ref[name]).max() if ref[name].size else 0.0)
        b = parity.mismatches(got[name], ref[name], sc)
        if b.size:
            problems.append((name, b.size, b[:3].tolist()))
    b = parity.mismatches(got["iBeta"][~lake], ref["iBeta"][~lake])
    if b.size:
        problems.append(("iBeta", b.size, b[:3].tolist()))
    for name in ("u_satn_out", "qEleE_IC_out"):
        b = parity.mismatches(got[name], ref[name])
        if b.size:
            problems.append((name, b.size, b[:3].tolist()))
    assert not problems, f"{basin}.{case}: {problems}"


@pytest.mark.parametrize("basin,case", [("ccw", "rand1"), ("qhh", "mut5")])
def test_host_entry_point_f(basin, case):
    """shud_b200_rhs: host vectors in reference order (the CVRhsFn shape), solver mode (no diag)."""
    from shud_up_b200.api import ShudRHS
    snap = oracle_lib.load_case(basin, case)
    ref = oracle_lib.oracle_rhs(snap)
    rhs = ShudRHS(snap)
    rhs.set_forcing(snap, qEleE_IC=snap["qEleE_IC_in"])
    rhs.set_carried(snap["ele_u_satn"])
    y = np.ascontiguousarray(snap["y"])
    ydot = np.full_like(y, np.nan)
    assert rhs.f(0.0, y, ydot) == 0
    bad = parity.mismatches(ydot, ref["ydot"], parity.ydot_scale(snap, ref))
    assert bad.size == 0
    # carried state moved on exactly as the reference's did
    s, e = rhs.get_carried()
    assert parity.mismatches(s, ref["u_satn_out"]).size == 0
    assert parity.mismatches(e, ref["qEleE_IC_out"]).size == 0
    # a second call on the same y reproduces the first bit for bit (deterministic sums, no atomics)
    ydot2 = np.empty_like(y)
    rhs.set_forcing(snap, qEleE_IC=snap["qEleE_IC_in"])
    rhs.set_carried(snap["ele_u_satn"])
    rhs.f(0.0, y, ydot2)
    assert np.array_equal(ydot, ydot2)


def test_prime_matches_oracle():
    from shud_up_b200.api import ShudRHS
    snap = oracle_lib.load_case("qhh", "rand4")
    rhs = ShudRHS(snap)
    rhs.prime(snap["y"])
    s, _ = rhs.get_carried()
    assert np.array_equal(s, oracle_lib.oracle_prime(snap, snap["y"]))


def test_device_error_word_mirrors_reference_exit_codes():
    """NaN in y -> the reference aborts with ERRNAN (10) from CheckNonNegative/CheckNANij
    (src/ModelData/MD_ET.cpp:394-401, MD_f.cpp:73-74); the replacement reports the same code."""
    from shud_up_b200.api import ShudRHS, ShudError
    snap = oracle_lib.load_case("ccw", "rand1")
    y = snap["y"].copy()
    y[2 * int(snap["Ne"][0]) + 5] = np.nan  # a NaN groundwater head -> NaN edge flux -> CheckNANij
    ref = oracle_lib.oracle_rhs(snap, y=y)
    assert ref["err"] == 10
    rhs = ShudRHS(snap)
    rhs.set_forcing(snap, qEleE_IC=snap["qEleE_IC_in"])
    rhs.set_carried(snap["ele_u_satn"])
    with pytest.raises(ShudError, match="code 10"):
        rhs.f(0.0, y, np.empty_like(y))


def test_device_side_output_accumulation():
    """Print_Ctrl::PrintData semantics (src/classes/Model_Control.cpp:930-962): buffer += value per SolverStep,
    buffer *= tau/NumUpdate at the interval end, reset - done on the device, one download per interval."""
    import torch
    snap = oracle_lib.load_case("qhh", "rand4")
    rhs, first = _run_gpu(snap)               # one diag RHS
    d1 = {k: first[k].copy() for k in abi.DIAG_ALL}
    rhs.output_accumulate()
    st = rhs.torch_stream()
    with torch.cuda.stream(st):
        y_ref = torch.from_numpy(np.ascontiguousarray(snap["y"] * 1.01)).cuda()
        y = torch.empty_like(y_ref); yd = torch.empty_like(y_ref)
        rhs.to_device_order(y_ref, y)
        rhs.f_dev(0.0, y, yd, diag=True)
    d2 = rhs.get_diag()
    rhs.output_accumulate()
    mean, n = rhs.output_flush(tau=1440.0)
    assert n == 2
    lake = snap["ele_iLake"] > 0
    for k in abi.DIAG_ALL:
        if d1[k].size == 0:
            continue
        want = (d1[k] + d2[k]) * (1440.0 / 2)
        got = mean[k]
        if k == "iBeta":
            want, got = want[~lake], got[~lake]
        assert np.array_equal(got, want), k
    # ... and against the ORACLE: the interval means Print_Ctrl would write from the reference's own arrays after the
    # same two f() calls (carried state handed from the first to the second)
    o1 = oracle_lib.oracle_rhs(snap)
    o2 = oracle_lib.oracle_rhs(snap, y=snap["y"] * 1.01, u_satn=o1["u_satn_out"], qEleE_IC=o1["qEleE_IC_out"])
    assert o1["err"] == 0 and o2["err"] == 0
    Ne = rhs.Ne
    addends = {"QeleSurfTot": lambda o: np.abs(o["QeleSurf"]).reshape(3, Ne).sum(0) + np.abs(o["Qe2r_Surf"]),
               "QeleSubTot": lambda o: np.abs(o["QeleSub"]).reshape(3, Ne).sum(0) + np.abs(o["Qe2r_Sub"])}
    problems = []
    for k in abi.DIAG_ALL:
        if o1[k].size == 0:
            continue
        want, got = (o1[k] + o2[k]) * (1440.0 / 2), mean[k]
        if k in addends:
            sc = (addends[k](o1) + addends[k](o2)) * (1440.0 / 2)
        elif k in SUMS:
            sc = np.full(want.shape, np.abs(want).max())
        else:
            sc = None
        if k == "iBeta":
            want, got = want[~lake], got[~lake]
        b = parity.mismatches(got, want, sc)
        if b.size:
            problems.append((k, b.size, b[:3].tolist()))
    assert not problems, problems
    # buffers were reset
    rhs.output_accumulate()
    again, n = rhs.output_flush(tau=2.0)
    assert n == 1 and np.array_equal(again["QrivDown"], d2["QrivDown"] * 2.0)


@pytest.mark.parametrize("nx,ny,lake", [(40, 30, 0.03), (37, 23, 0.0), (64, 64, 0.05)])
def test_synthetic_meshes_match_oracle(nx, ny, lake):
    """synthetic domains (bench generator): odd sizes / tail tiles, a lake with hundreds of bank edges and lake
    cells (tree reduction of the lake sums), dense river trees - ydot against the oracle at 1e-12"""
    from shud_up_b200 import synth
    mesh = synth.make(nx, ny, ntree=max(1, ny // 10), reaches_per_tree=min(nx, 40), lake_frac=lake)
    mesh["ele_u_satn"] = oracle_lib.oracle_prime(mesh, mesh["y"])
    ref = oracle_lib.oracle_rhs(mesh)
    assert ref["err"] == 0
    rhs, got = _run_gpu(mesh)
    assert got["code"] == 0
    bad = parity.mismatches(got["ydot"], ref["ydot"], parity.ydot_scale(mesh, ref))
    assert bad.size == 0, (bad[:5], got["ydot"][bad[:5]], ref["ydot"][bad[:5]])
    for name in ("QsegSurf", "QsegSub", "QrivDown", "qEleInfil", "qEleRecharge", "u_effKH"):
        assert parity.mismatches(got[name], ref[name]).size == 0, name
    if lake:
        assert int(mesh["Nl"][0]) == 1 and (mesh["ele_lakenabr"] > 0).sum() > 20


def _subset(snap, n):
    """the first n cells of a basin as a mesh of their own: neighbours outside become boundary edges, rivers dropped"""
    Ne = int(snap["Ne"][0])
    s = {}
    for k, v in snap.items():
        v = np.asarray(v)
        if k.startswith("riv_") or k.startswith("seg_"):
            s[k] = v[:0]
        elif v.ndim == 1 and v.size == Ne:
            s[k] = v[:n].copy()
        elif v.ndim == 1 and v.size == 3 * Ne:
            s[k] = v.reshape(3, Ne)[:, :n].copy().reshape(-1)
        else:
            s[k] = v
    nab = s["ele_nabr"].reshape(3, n)
    nab[nab > n] = 0
    s["ele_nabr"] = nab.reshape(-1)
    s["Ne"] = np.array([n], dtype=np.int32); s["Nr"] = np.array([0], dtype=np.int32); s["Ns"] = np.array([0], dtype=np.int32)
    y = np.asarray(snap["y"])
    s["y"] = np.concatenate([y[:n], y[Ne:Ne + n], y[2 * Ne:2 * Ne + n]])
    return s


@pytest.mark.parametrize("n", [1147, 100, 1])
def test_ragged_and_riverless_meshes(n):
    """edge cases of the launch geometry: no reaches and no segments at all (every river kernel launch is empty), a
    mesh smaller than one 128-cell tile, a single cell (all three edges on the boundary)"""
    snap = _subset(oracle_lib.load_case("ccw", "rand1"), n)
    ref = oracle_lib.oracle_rhs(snap)
    assert ref["err"] == 0
    rhs, got = _run_gpu(snap)
    assert got["code"] == 0, got
    assert got["ydot"].size == 3 * n
    bad = parity.mismatches(got["ydot"], ref["ydot"], parity.ydot_scale(snap, ref))
    assert bad.size == 0, (n, bad[:5], got["ydot"][bad[:5]], ref["ydot"][bad[:5]])


@pytest.mark.parametrize("basin,case", [("ccw", "rand1"), ("qhh", "lakes6"), ("heihe", "ic")])
def test_difference_quotient_evaluation_folded_into_the_prepass(basin, case):
    """shud_b200_rhs_dq_dev: ytemp = sigma (v ./ ewt) + y0 formed by the pre-pass and f(ytemp) - bit for bit the perturbed
    vector of shud_nv_dq_perturb and the ydot of shud_b200_rhs_dev on it (CVLS' difference-quotient J v inside SPGMR,
    src/Equations/cvode_config.cpp:172-179), on repeated calls (graph replay) and with another direction"""
    import torch
    from shud_up_b200.api import ShudRHS
    from shud_up_b200.nvector import NVectorOps
    snap = oracle_lib.load_case(basin, case)
    rhs = ShudRHS(snap)

    def reset():
        rhs.set_forcing(snap, qEleE_IC=snap["qEleE_IC_in"])
        rhs.set_carried(snap["ele_u_satn"])

    reset()
    st = rhs.torch_stream()
    rng = np.random.default_rng(11)
    y_ref = np.ascontiguousarray(snap["y"])
    n = y_ref.size
    sigma = float(np.sqrt(n))
    with torch.cuda.stream(st):
        y0 = torch.empty(n, dtype=torch.float64, device="cuda")
        rhs.to_device_order(torch.from_numpy(y_ref).cuda(), y0)
        ewt = 1.0 / (1e-4 * y0.abs() + 1e-4)
        vs = [torch.from_numpy(rng.normal(0, 1, n) / np.sqrt(n)).cuda() for _ in range(2)]
        ops = NVectorOps(0, rhs.stream_ptr, owner=rhs)
        for v in vs:
            yt_a, yd_a = torch.empty_like(y0), torch.empty_like(y0)
            yt_b, yd_b = torch.full_like(y0, float("nan")), torch.full_like(y0, float("nan"))
            ops.DQPerturb(sigma, v, ewt, y0, yt_a)
            reset()
            rhs.f_dev(0.0, yt_a, yd_a)
            st.synchronize()
            for rep in range(3):
                reset()  # both arms start from the same carried saturation / interception state
                rhs.f_dq_dev(0.0, sigma, v, ewt, y0, yt_b, yd_b)
                st.synchronize()
                assert torch.equal(yt_a, yt_b), (basin, case, rep)
                assert torch.equal(yd_a, yd_b), (basin, case, rep, float((yd_a - yd_b).abs().max()))
        # the unnormalised form: direction (1 / sqrt(ss)) v, normalised vector handed back - the solver's normalisation
        # pass done by the pre-pass
        raw = vs[0] * 3.7
        ss = (raw * raw).sum().reshape(1)
        vn_a = (1.0 / torch.sqrt(ss)) * raw
        yt_a, yd_a = torch.empty_like(y0), torch.empty_like(y0)
        ops.DQPerturb(sigma, vn_a, ewt, y0, yt_a)
        reset()
        rhs.f_dev(0.0, yt_a, yd_a)
        st.synchronize()
        vn_b, yt_b, yd_b = (torch.full_like(y0, float("nan")) for _ in range(3))
        for rep in range(2):
            reset()
            rhs.f_dq_dev(0.0, sigma, raw, ewt, y0, yt_b, yd_b, ss_dev=ss, vnorm_dev=vn_b)
            st.synchronize()
            assert torch.equal(vn_a, vn_b) and torch.equal(yt_a, yt_b) and torch.equal(yd_a, yd_b), (basin, case, rep)
        assert rhs.check()[0] == 0
        ops.close()
