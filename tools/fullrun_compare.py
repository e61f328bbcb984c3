#!/usr/bin/env python
"""ccw short full run: CUDA RHS + device N_Vector vs CPU oracle + host ops under the same integrator."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib
from host_model import OracleModel
from shud_up_b200 import driver, snapshot
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
mesh = oracle_lib.load_case("ccw", "ic")
fseq = snapshot.load(os.path.join(oracle_lib.GOLDEN, "ccw.fseq.npz"))
Ne = int(mesh["Ne"][0])
fseq = {k: (v.reshape(-1, Ne) if v.size % Ne == 0 and v.size > Ne else v) for k, v in fseq.items()}
for tol in (1e-4, 1e-7):
    ref = driver.run(OracleModel(mesh, fseq), fseq, mesh["y"], n_steps=n, rtol=tol, atol=tol)
    gm = driver.GpuModel(mesh, fseq)
    gpu = driver.run(gm, fseq, mesh["y"], n_steps=n, rtol=tol, atol=tol)
    gm.close()
    ewt = tol * np.abs(ref["y_end"]) + tol
    print(json.dumps({"tol": tol, "gpu_stats": gpu["stats"], "ref_stats": ref["stats"],
                      "q_rel_diff_max": float(np.max(np.abs(gpu["q_out"] - ref["q_out"]) / np.abs(ref["q_out"]))),
                      "y_wrms_diff": float(np.sqrt(np.mean(((gpu["y_end"] - ref["y_end"]) / ewt) ** 2))),
                      "y_abs_diff_max": float(np.abs(gpu["y_end"] - ref["y_end"]).max()),
                      "sim_days_per_s_gpu": gpu["sim_days_per_wall_s"], "sim_days_per_s_cpu": ref["sim_days_per_wall_s"]}))
