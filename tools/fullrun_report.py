#!/usr/bin/env python
"""30-day full runs of ccw / heihe / qhh, GPU arm vs checker arm under the library's CVODE-shaped integrator
(shud_up_b200.driver.run_cv; see tests/test_fullrun_gpu.py): one JSON line per basin with the hydrograph scores, the
basin water budget of both arms, the integrator statistics and simulated days per wall second.
    python tools/fullrun_report.py [days] > profiles/r02_fullrun_gpu_vs_oracle.jsonl     (on the GPU box)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import host_cv  # noqa: E402
import oracle_lib  # noqa: E402
from shud_up_b200 import driver  # noqa: E402
from test_fullrun_gpu import compare  # noqa: E402

days = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
for basin in ("ccw", "heihe", "qhh"):
    mesh = oracle_lib.load_case(basin, "ic")
    run = dict(np.load(os.path.join(oracle_lib.GOLDEN, f"{basin}.run.npz")))
    n = min(int(round(days * 1440.0 / float(run["run_cfg"][3]))), int(run["run_cfg"][5]))
    every = int(round(60.0 / float(run["run_cfg"][3])))          # hourly samples of the hydrograph / budget
    ref = driver.run_cv(host_cv.OracleArm(mesh, run), run, n_steps=n, sample_every=every)
    arm = driver.GpuArm(mesh, run)
    gpu = driver.run_cv(arm, run, n_steps=n, sample_every=every)
    arm.close()
    c = compare(gpu, ref)
    ewt = 1e-4 * np.abs(ref["y_end"]) + 1e-4
    rec = dict(basin=basin, days=n * float(run["run_cfg"][3]) / 1440.0, t0_min=float(run["run_cfg"][4]),
               nse=c["nse"], vol_err=float(c["vol_err"]),
               wrms_end_state=float(np.sqrt(np.mean(((gpu["y_end"] - ref["y_end"]) / ewt) ** 2))),
               q_peak_m3_per_min=float(ref["q_out"].sum(axis=1).max()),
               budget_gpu=gpu["budget"], budget_oracle=ref["budget"], stats_gpu=gpu["stats"], stats_oracle=ref["stats"],
               sim_days_per_wall_s_gpu=gpu["sim_days_per_wall_s"], sim_days_per_wall_s_oracle=ref["sim_days_per_wall_s"])
    print(json.dumps(rec), flush=True)
