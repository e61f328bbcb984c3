"""The CVODE-shaped integrator (shud_up_b200/integrator.py) on the CPU: driven by the oracle RHS and a host
ops table over a few hours of ccw with the reference's own forcing sequence.  Checks that it integrates (error
control works: a tight-tolerance run is the yardstick) and that SPGMR solves a known linear system."""
import os

import numpy as np

import oracle_lib
from host_model import HostOps, OracleModel
from shud_up_b200 import driver, snapshot
from shud_up_b200.integrator import BDFKrylov


def test_spgmr_solves_linear_system():
    rng = np.random.default_rng(0)
    n = 40
    J = -np.diag(rng.uniform(0.5, 3.0, n)) + 0.05 * rng.standard_normal((n, n))
    integ = BDFKrylov(HostOps(), lambda: np.zeros(n), lambda t, y, yd: np.copyto(yd, J @ y), n, maxl=30)
    integ.ewt[:] = 1.0
    b, x, y, fy = rng.standard_normal(n), np.zeros(n), np.zeros(n), np.zeros(n)
    ok, it = integ._spgmr(b.copy(), x, 0.7, 0.0, y, fy, 1e-10)
    assert ok and np.allclose((np.eye(n) - 0.7 * J) @ x, b, atol=1e-8)


def test_linear_decay_accuracy():
    lam = np.array([0.01, 0.1, 1.0, 5.0])
    rhs = lambda t, y, yd: np.copyto(yd, -lam * y)
    integ = BDFKrylov(HostOps(), lambda: np.zeros(4), rhs, 4, rtol=1e-6, atol=1e-9, max_step=10.0, init_step=1e-3)
    integ.init(0.0, np.ones(4))
    y = integ.advance(20.0)
    assert np.allclose(y, np.exp(-lam * 20.0), rtol=2e-3, atol=1e-7)
    assert integ.stats["nst"] < 4000


def test_ccw_six_hours_with_the_oracle():
    mesh = oracle_lib.load_case("ccw", "ic")
    fseq = snapshot.load(os.path.join(oracle_lib.GOLDEN, "ccw.fseq.npz"))
    Ne = int(mesh["Ne"][0])
    fseq = {k: (v.reshape(-1, Ne) if v.size % Ne == 0 and v.size > Ne else v) for k, v in fseq.items()}
    a = driver.run(OracleModel(mesh, fseq), fseq, mesh["y"], n_steps=6)
    b = driver.run(OracleModel(mesh, fseq), fseq, mesh["y"], n_steps=6, rtol=1e-6, atol=1e-6)
    assert np.isfinite(a["y_end"]).all() and a["stats"]["nst"] >= 36
    # error control: the 1e-4 run stays within a modest multiple of its tolerance of the 1e-6 run
    ewt = 1e-4 * np.abs(b["y_end"]) + 1e-4
    assert np.sqrt(np.mean(((a["y_end"] - b["y_end"]) / ewt) ** 2)) < 20.0
    assert a["q_out"].shape == (6, 1) and np.all(a["q_out"] >= 0)
