"""The allreduce INSIDE the vector reduction kernels (shud_nv_ws_set_peer_allreduce, include/shud_nvector.h): two ranks
as two processes on cuda:0, their peer-to-peer blocks mapped through CUDA IPC (shud_b200_p2p_export / _connect), the
last block of every reduction exchanging the ranks' partial results through the mailboxes.  The global value must be the
ranks' LOCAL values combined in rank order - bit for bit, identical on both ranks - for sums, maxima, minima and the
multi-dot, over several rounds (both slots of the mailboxes).  On a multi-GPU box the same code runs one rank per GPU
over NVLink (bench.py --gpus N: newton_krylov block)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_reduction_kernels_combine_the_ranks_partials_through_the_mailboxes(tmp_path):
    world = 2
    procs = [subprocess.Popen([sys.executable, os.path.join(HERE, "allreduce_worker.py"), str(r), str(world), str(tmp_path)],
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = []
    try:
        for p in procs:
            outs.append(p.communicate(timeout=300)[0])
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    if any(p.returncode == 77 for p in procs):
        pytest.skip("this GPU admits one context at a time (exclusive-process mode): no second rank on it")
    for r, p in enumerate(procs):
        assert p.returncode == 0, outs[r][-2000:]
    res = [dict(np.load(os.path.join(tmp_path, f"result{r}.npz"))) for r in range(world)]
    n_global = sum(int(x["n"][0]) for x in res)
    for key in ("dot", "max", "min", "multi", "wrms"):
        assert np.array_equal(res[0]["glo_" + key], res[1]["glo_" + key]), key      # the same bits on every rank
        assert np.all(np.isfinite(res[0]["glo_" + key])), key
    # rank order: ((identity op local_0) op local_1)
    assert np.array_equal(res[0]["glo_dot"], (0.0 + res[0]["loc_dot"]) + res[1]["loc_dot"])
    assert np.array_equal(res[0]["glo_multi"], (0.0 + res[0]["loc_multi"]) + res[1]["loc_multi"])
    assert np.array_equal(res[0]["glo_max"], np.maximum(res[0]["loc_max"], res[1]["loc_max"]))
    assert np.array_equal(res[0]["glo_min"], np.minimum(res[0]["loc_min"], res[1]["loc_min"]))
    assert np.array_equal(res[0]["glo_wrms"], np.sqrt(((0.0 + res[0]["loc_wsq"]) + res[1]["loc_wsq"]) / n_global))
    again = (0.0 + res[0]["loc_again"][0]) + res[1]["loc_again"][0]
    assert np.array_equal(res[0]["glo_again"], [again, again]) and np.array_equal(res[1]["glo_again"], [again, again])
    assert not np.array_equal(res[0]["loc_dot"], res[1]["loc_dot"])                  # the ranks did hold different data
