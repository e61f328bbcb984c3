"""Known-answer tests of the device N_Vector (the reference holds no tests at this
boundary - 'parity unpinned' - so the pin is the SUNDIALS 6 definition of each op evaluated by a
trivially-correct host loop: exact for the streaming ops, <= n*eps relative for the reductions
(tree order != sequential order))."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

NS = [1, 31, 1000, 3544, 1_000_003]


def _mk(n, seed, k=1):
    import torch
    rng = np.random.default_rng(seed)
    hs = [rng.standard_normal(n) * np.exp(rng.uniform(-3, 3, n)) for _ in range(k)]
    ds = [torch.from_numpy(h).cuda() for h in hs]
    return hs, ds


@pytest.fixture(scope="module")
def ops():
    import torch
    from shud_up_b200.nvector import NVectorOps
    st = torch.cuda.Stream()
    o = NVectorOps(0, st.cuda_stream)
    o._st = st
    yield o
    o.close()


@pytest.mark.parametrize("n", NS)
def test_streaming_ops_exact(ops, n):
    import torch
    (x, y), (dx, dy) = _mk(n, 1, 2)
    y = np.where(np.abs(y) < 1e-3, 1.0, y); dy = torch.from_numpy(y).cuda()
    torch.cuda.synchronize()
    dz = torch.empty_like(dx)
    sync = lambda: ops._st.synchronize()
    ops.N_VLinearSum(1.5, dx, -0.25, dy, dz); sync(); assert np.array_equal(dz.cpu().numpy(), 1.5 * x + (-0.25) * y)
    ops.N_VConst(3.25, dz); sync(); assert (dz.cpu().numpy() == 3.25).all()
    ops.N_VProd(dx, dy, dz); sync(); assert np.array_equal(dz.cpu().numpy(), x * y)
    ops.N_VDiv(dx, dy, dz); sync(); assert np.array_equal(dz.cpu().numpy(), x / y)
    ops.N_VScale(-2.5, dx, dz); sync(); assert np.array_equal(dz.cpu().numpy(), -2.5 * x)
    ops.N_VAbs(dx, dz); sync(); assert np.array_equal(dz.cpu().numpy(), np.abs(x))
    ops.N_VInv(dy, dz); sync(); assert np.array_equal(dz.cpu().numpy(), 1.0 / y)
    ops.N_VAddConst(dx, 0.125, dz); sync(); assert np.array_equal(dz.cpu().numpy(), x + 0.125)
    ops.N_VCompare(0.5, dx, dz); sync(); assert np.array_equal(dz.cpu().numpy(), (np.abs(x) >= 0.5).astype(float))
    # in-place forms CVODE uses: y += a x
    dw = dy.clone(); torch.cuda.synchronize()
    ops.N_VLinearSum(0.75, dx, 1.0, dw, dw); sync(); assert np.array_equal(dw.cpu().numpy(), 0.75 * x + y)


@pytest.mark.parametrize("n", NS)
def test_reductions(ops, n):
    import torch
    (x, y, w), (dx, dy, dw) = _mk(n, 2, 3)
    torch.cuda.synchronize()
    tol = lambda terms: 4 * n * np.finfo(float).eps * np.sum(np.abs(terms)) + 1e-300
    assert abs(ops.N_VDotProd(dx, dy) - np.dot(x, y)) <= tol(x * y)
    assert ops.N_VMaxNorm(dx) == np.abs(x).max()
    assert ops.N_VMin(dx) == x.min()
    assert abs(ops.N_VL1Norm(dx) - np.abs(x).sum()) <= tol(x)
    s = np.sum((x * w) ** 2)
    assert abs(ops.N_VWSqrSumLocal(dx, dw) - s) <= tol((x * w) ** 2)
    assert abs(ops.N_VWrmsNorm(dx, dw) - np.sqrt(s / n)) <= 1e-13 * np.sqrt(s / n)
    assert abs(ops.N_VWL2Norm(dx, dw) - np.sqrt(s)) <= 1e-13 * np.sqrt(s)
    idv = (np.arange(n) % 3 != 0).astype(float); did = torch.from_numpy(idv).cuda(); torch.cuda.synchronize()
    sm = np.sum(((x * w) ** 2)[idv > 0])
    assert abs(ops.N_VWrmsNormMask(dx, dw, did) - np.sqrt(sm / n)) <= 1e-13 * max(np.sqrt(sm / n), 1e-300)
    # determinism: same call, same bits
    assert ops.N_VDotProd(dx, dy) == ops.N_VDotProd(dx, dy)
    den = np.where(np.arange(n) % 5 == 0, 0.0, y); dden = torch.from_numpy(den).cuda(); torch.cuda.synchronize()
    ref = (x[den != 0] / den[den != 0]).min() if (den != 0).any() else np.finfo(float).max
    assert ops.N_VMinQuotient(dx, dden) == ref


def test_tests_and_masks(ops):
    import torch
    n = 10007
    (x, c), (dx, dc) = _mk(n, 3, 2)
    x[17] = 0.0; dx = torch.from_numpy(x).cuda()
    dz = torch.full_like(dx, -7.0); torch.cuda.synchronize()
    assert ops.N_VInvTest(dx, dz) is False
    z = dz.cpu().numpy(); nz = x != 0
    assert np.array_equal(z[nz], 1.0 / x[nz]) and z[17] == -7.0
    x[17] = 2.0; dx = torch.from_numpy(x).cuda(); torch.cuda.synchronize()
    assert ops.N_VInvTest(dx, dz) is True
    cc = np.random.default_rng(5).integers(-2, 3, n).astype(float); dcc = torch.from_numpy(cc).cuda()
    dm = torch.empty_like(dx); torch.cuda.synchronize()
    ok = ops.N_VConstrMask(dcc, dx, dm)
    bad = np.where(np.abs(cc) > 1.5, x * cc <= 0, np.where(cc != 0, x * cc < 0, False))
    assert ok == (not bad.any()) and np.array_equal(dm.cpu().numpy(), bad.astype(float))


@pytest.mark.parametrize("n,nv", [(3544, 2), (1_000_003, 6)])
def test_fused_ops(ops, n, nv):
    import torch
    hs, ds = _mk(n, 4, 2 * nv + 1)
    X, Y, x = hs[:nv], hs[nv:2 * nv], hs[-1]
    dX, dY, dxv = ds[:nv], ds[nv:2 * nv], ds[-1]
    dZ = [torch.empty_like(dxv) for _ in range(nv)]
    torch.cuda.synchronize()
    c = np.linspace(-1.5, 2.0, nv)
    sync = lambda: ops._st.synchronize()
    ops.N_VLinearCombination(c, dX, dZ[0]); sync()
    ref = c[0] * X[0]
    for k in range(1, nv):
        ref = ref + c[k] * X[k]          # SUNDIALS order: sequential accumulation
    assert np.array_equal(dZ[0].cpu().numpy(), ref)
    ops.N_VScaleAddMulti(c, dxv, dY, dZ); sync()
    for k in range(nv):
        assert np.array_equal(dZ[k].cpu().numpy(), c[k] * x + Y[k])
    d = ops.N_VDotProdMulti(dxv, dY)
    for k in range(nv):
        assert abs(d[k] - np.dot(x, Y[k])) <= 4 * n * np.finfo(float).eps * np.sum(np.abs(x * Y[k]))
        assert d[k] == ops.N_VDotProd(dxv, dY[k])  # the fused form reduces in the same tree
    ops.N_VLinearSumVectorArray(0.5, dX, -2.0, dY, dZ); sync()
    for k in range(nv):
        assert np.array_equal(dZ[k].cpu().numpy(), 0.5 * X[k] + (-2.0) * Y[k])
    ops.N_VScaleVectorArray(c, dX, dZ); sync()
    for k in range(nv):
        assert np.array_equal(dZ[k].cpu().numpy(), c[k] * X[k])
    ops.N_VConstVectorArray(1.25, dZ); sync()
    assert all((z.cpu().numpy() == 1.25).all() for z in dZ)
    nrm = ops.N_VWrmsNormVectorArray(dX, dY)
    for k in range(nv):
        r = np.sqrt(np.sum((X[k] * Y[k]) ** 2) / n)
        assert abs(nrm[k] - r) <= 1e-13 * r


def test_dq_fusions(ops):
    import torch
    n = 100_003
    (vs, ewt, y, fp, fy), (dvs, dewt, dy, dfp, dfy) = _mk(n, 9, 5)
    ewt = np.abs(ewt) + 0.5; dewt = torch.from_numpy(ewt).cuda()
    dout = torch.empty_like(dy); torch.cuda.synchronize()
    sig, gam = 37.5, 0.125
    ops.DQPerturb(sig, dvs, dewt, dy, dout); ops._st.synchronize()
    assert np.array_equal(dout.cpu().numpy(), sig * (vs / ewt) + y)
    ops.DQCombine(sig, gam, dvs, dewt, dfp, dfy, dout); ops._st.synchronize()
    jv = (1.0 / sig) * fp + (-1.0 / sig) * fy
    assert np.array_equal(dout.cpu().numpy(), ewt * (vs / ewt + (-gam) * jv))


def test_newton_fusions_known_answers(ops):
    """shud_nv_ewt / shud_nv_newton_resid / shud_nv_newton_update against the operation sequences they replace
    (cvEwtSetSS: Abs, Scale, AddConst, Inv; the Newton right-hand side; the correction + its WRMS norm)"""
    import torch
    for n in (1, 1000, 3544, 1_000_003):
        (y, f, psi, x, acor), (dy, df, dpsi, dx, dacor) = _mk(n, 21 + n % 7, 5)
        dewt, dr = torch.empty_like(dy), torch.empty_like(dy)
        torch.cuda.synchronize()
        rtol, atol, gamma = 1e-4, 1e-6, 0.37
        ops.EwtSet(rtol, atol, dy, dewt); ops._st.synchronize()
        ewt = 1.0 / (rtol * np.abs(y) + atol)
        assert np.array_equal(dewt.cpu().numpy(), ewt)                 # same operations in the same order: same bits
        ops.NewtonResid(gamma, df, dpsi, dy, dr); ops._st.synchronize()
        assert np.array_equal(dr.cpu().numpy(), gamma * f + psi - y)
        dele = ops.NewtonUpdate(dx, dewt, dy, dacor); ops._st.synchronize()
        assert np.array_equal(dy.cpu().numpy(), y + x) and np.array_equal(dacor.cpu().numpy(), acor + x)
        want = np.sqrt(np.sum((x * ewt) ** 2) / n)
        assert abs(dele - want) <= 1e-13 * want
