"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
entry point include/*.h declares; the product path never touches the oracle and fails loudly
(no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = []
    for h in ("shud_b200.h", "shud_nvector.h", "shud_sundials.h", "shud_cvode.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names += re.findall(r"\b(shud_(?:b200|nv|cv|spgmr)_\w+)\s*\(", txt)
        names += re.findall(r"\b(N_V\w+)\s*\(", txt)          # the vector constructors and the generic dispatch
    # function-pointer typedefs and struct members are not entry points
    skip = {"shud_nv_allreduce_fn", "shud_nv_allreduce_dev_fn", "shud_cv_rhs_fn", "N_Vector", "N_Vector_ID", "N_Vector_Ops", "N_Vector_S"}
    return sorted(n for n in set(names) if n not in skip)


@pytest.fixture(scope="module")
def lib():
    from shud_up_b200 import build
    return ctypes.CDLL(build.build())


def test_every_declared_entry_point_is_exported(lib):
    names = _declared()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_library_is_sm100a_only(lib):
    from shud_up_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    from shud_up_b200.api import ShudRHS, ShudError
    snap = oracle_lib.load_case("ccw", "ic")
    with pytest.raises(ShudError, match="no CUDA device"):
        ShudRHS(snap)
    from shud_up_b200.nvector import NVectorOps
    with pytest.raises(ShudError, match="no CUDA device"):
        NVectorOps(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "shud_up_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle_lib" not in txt and "liboracle" not in txt and "shud_oracle" not in txt, f


def test_abi_struct_layout_matches_header():
    """field order of the ctypes mirrors == declaration order in the header"""
    from shud_up_b200 import abi
    txt = open(os.path.join(ROOT, "include", "shud_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for cname, cls in (("shud_mesh", abi.ShudMesh), ("shud_forcing", abi.ShudForcing), ("shud_diag", abi.ShudDiag)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), txt, flags=re.S).group(1)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            decl = re.sub(r"^(const\s+)?(double|int32_t)\s*", "", decl)
            fields += [x.strip().lstrip("*").strip() for x in decl.split(",")]
        assert fields == [f[0] for f in cls._fields_], cname


def test_headers_are_plain_c_and_cxx(tmp_path):
    """the boundary headers compile on their own as C99 and as C++14 (the reference's language level)"""
    import subprocess
    inc = os.path.join(ROOT, "include")
    src_c = tmp_path / "t.c"
    src_c.write_text('#include "shud_b200.h"\n#include "shud_nvector.h"\nint main(void) { return SHUD_OK; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc, str(src_c)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    src_cc = tmp_path / "t.cpp"
    src_cc.write_text('#include "shud_nvector.h"\n#include "shud_b200.h"\nint main() { return SHUD_OK; }\n')
    r = subprocess.run(["g++", "-std=c++14", "-Wall", "-Werror", "-pedantic", "-fsyntax-only", "-I", inc, str(src_cc)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_new_entry_points_reject_bad_arguments_before_touching_a_device(lib):
    """argument checks of the entry points added for the fused Newton-Krylov step and the in-kernel allreduce: a NULL
    context / workspace / vector or an order outside 1..5 is SHUD_ERR_ARG (-2), with or without a GPU"""
    vp, d, i64 = ctypes.c_void_p, ctypes.c_double, ctypes.c_int64
    ERR_ARG = -2
    lib.shud_b200_rhs_dq_dev.argtypes = [vp, d, d, vp, vp, vp, vp, vp, vp, vp]
    assert lib.shud_b200_rhs_dq_dev(None, 0.0, 1.0, None, None, None, None, None, None, None) == ERR_ARG
    lib.shud_b200_dq_foldable.argtypes = [vp]
    assert lib.shud_b200_dq_foldable(None) == 0
    lib.shud_b200_p2p_mailboxes.argtypes = [vp, vp, vp, vp]
    assert lib.shud_b200_p2p_mailboxes(None, None, None, None) == ERR_ARG
    lib.shud_nv_ws_set_peer_allreduce.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp]
    assert lib.shud_nv_ws_set_peer_allreduce(None, 2, 0, None) == ERR_ARG
    lib.shud_nv_bdf_predict.argtypes = [vp, i64, ctypes.c_int, d, vp, vp, vp]
    assert lib.shud_nv_bdf_predict(None, 10, 0, 1.0, None, None, None) == ERR_ARG          # order 0
    assert lib.shud_nv_bdf_predict(None, 10, 6, 1.0, None, None, None) == ERR_ARG          # order 6
    lib.shud_nv_bdf_complete.argtypes = [vp, i64, ctypes.c_int, vp, vp, vp, d, d, vp, vp, i64, vp]
    assert lib.shud_nv_bdf_complete(None, 10, 2, None, None, None, 1e-4, 1e-4, None, None, 0, None) == ERR_ARG
    lib.shud_spgmr_newton_step.argtypes = [vp, d, d, d, vp, vp, vp, vp, vp, vp, d, i64, vp, vp, vp]
    assert lib.shud_spgmr_newton_step(None, 0.0, 1.0, 1.0, None, None, None, None, None, None, 1.0, 0, None, None, None) == ERR_ARG
    lib.shud_nv_ewt_wrms.argtypes = [vp, i64, d, d, vp, vp, i64, vp]
    assert lib.shud_nv_ewt_wrms(None, 10, 1e-4, 1e-4, None, None, 0, None) == ERR_ARG
