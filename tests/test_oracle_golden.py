"""Pin the CPU oracle (oracle/shud_oracle.c) to the unmodified reference f():
bit-for-bit agreement with the snapshots tools/make_golden.py took from
oracle/_ref/shud_ref_serial (reference src/Model/f.cpp:2-32 compiled in place)."""
import numpy as np
import pytest

import oracle_lib
from shud_up_b200 import abi

CASES = [("ccw", "ic"), ("ccw", "rand1"), ("ccw", "mut2"), ("heihe", "ic"), ("heihe", "rand3"),
         ("qhh", "ic"), ("qhh", "rand4"), ("qhh", "mut5"), ("qhh", "lakes6")]

# reference-side names of the flux arrays in the snapshots
REF_NAMES = {"u_satn": "u_satn_out"}


def _bits_equal(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.array_equal(a.view(np.int64), b.view(np.int64)) or np.array_equal(a, b)


@pytest.mark.parametrize("basin,case", CASES)
def test_oracle_matches_reference_bitwise(basin, case):
    snap = oracle_lib.load_case(basin, case)
    out = oracle_lib.oracle_rhs(snap)
    assert out["err"] == 0
    lake_cell = snap["ele_iLake"] > 0
    bad = []
    for name in ["ydot", "qEleE_IC_out"] + abi.DIAG_ALL:
        ref = snap.get(REF_NAMES.get(name, name))
        if ref is None or ref.size == 0:
            continue
        got = out[name]
        if name in ("iBeta",):  # the reference leaves iBeta of lake cells untouched (stale)
            got, ref = got[~lake_cell], ref[~lake_cell]
        if name in ("qEleTrans", "qEleEvapo", "qEleETA"):
            pass
        if not _bits_equal(got, ref):
            d = np.abs(got - ref)
            bad.append((name, int((got != ref).sum()), float(np.nanmax(d))))
    assert not bad, f"{basin}.{case}: arrays differing from the reference: {bad}"


@pytest.mark.parametrize("basin,case", [("ccw", "rand1"), ("qhh", "mut5")])
def test_oracle_omp_equals_serial(basin, case):
    snap = oracle_lib.load_case(basin, case)
    a = oracle_lib.oracle_rhs(snap, nthreads=1)
    b = oracle_lib.oracle_rhs(snap, nthreads=4)
    for k in ["ydot", "u_satn_out", "qEleE_IC_out"] + abi.DIAG_ALL:
        assert _bits_equal(a[k], b[k]), k


def test_known_answer_checksums_first_call():
    """SURVEY.md section 8(c) / BASELINE.md: checksums of the reference's FIRST f() after IC
    (state fresh from updateforcing()+ET(): qEleE_IC not yet clipped by f_etFlux)."""
    kat = {"ccw": (-1.8228929483094446e-04, 4.9482557986667405e-04),
           "qhh": (-3.7481481531591236e-03, 5.889306927393561e-02),
           "heihe": (-0.6253409261945998, 1.0772633320206721)}
    for basin, (s, sa) in kat.items():
        snap = oracle_lib.load_case(basin, "ic")
        out = oracle_lib.oracle_rhs(snap, u_satn=snap["first_u_satn"], qEleE_IC=snap["first_qEleE_IC_in"])
        assert _bits_equal(out["ydot"], snap["first_ydot"]), basin
        # sequential (left-to-right) sums, as the survey's probe printed them
        assert np.cumsum(out["ydot"])[-1] == s and np.cumsum(np.abs(out["ydot"]))[-1] == sa, basin
        # u_satn left by a call = updateElement(y) (src/classes/Element.cpp:347-373): priming from y
        # reproduces the carried state the second call saw.  (Before the very FIRST call the
        # reference computes it from the still-uninitialised uY* scratch - first_u_satn is all 0.)
        land = snap["ele_iLake"] <= 0
        assert _bits_equal(oracle_lib.oracle_prime(snap, snap["y"])[land], snap["ele_u_satn"][land]), basin


@pytest.mark.parametrize("basin,case", CASES)
def test_oracle_first_call(basin, case):
    snap = oracle_lib.load_case(basin, case)
    out = oracle_lib.oracle_rhs(snap, u_satn=snap["first_u_satn"], qEleE_IC=snap["first_qEleE_IC_in"])
    assert _bits_equal(out["ydot"], snap["first_ydot"])
    assert _bits_equal(out["u_satn_out"], snap["ele_u_satn"])
    assert _bits_equal(out["qEleE_IC_out"], snap["qEleE_IC_in"])
