"""bench.py contract pieces that can run without a GPU: the reference arm (the reference's algorithm on the host
cores, oracle port with OpenMP) prints ONE JSON line with the keys the driver reads, on the GPU arm's metric / unit /
workload."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "rhs_cell_updates_per_sec" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"].startswith("synthetic-1M per GPU: 1,000,000 cells / 50,000 reaches / 150,000 segments")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"roofline"', '"cpu_baseline"', '"e2e"', '"clocks"', '"gpu_launches"', '"vs_baseline"', '"scaling"'):
        assert key in src, key


def test_bench_parity_helper_partition_equals_whole_mesh():
    """bench.py's in-bench parity check at N>1 runs the CPU oracle on 'owned + halo' built from what the partition
    alone knows (halo statics + exchanged state): on the owned cells, reaches and lakes it must reproduce the oracle
    on the whole mesh bit for bit - fed with that, the helper reports n_bad == 0 and max_rel == 0; fed with a
    perturbed vector it reports the mismatch."""
    import numpy as np
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import bench
    import oracle_lib
    from shud_up_b200 import synth
    NX, NY, NT, RPT = 40, 40, 4, 40
    whole = synth.make(NX, NY, ntree=NT, reaches_per_tree=RPT)
    Ne_w, Nr_w = int(whole["Ne"][0]), int(whole["Nr"][0])
    whole["ele_u_satn"] = oracle_lib.oracle_prime(whole, whole["y"])
    ref = oracle_lib.oracle_rhs(whole, want_diag=False)["ydot"]
    pos = np.empty(whole["ele_gid"].max() + 1, dtype=np.int64)
    pos[whole["ele_gid"]] = np.arange(Ne_w)
    for r0 in (0, 20):
        loc = synth.make(NX, NY, ntree=NT, reaches_per_tree=RPT, rows=(r0, r0 + 20), stripe_rows=20)
        sel = pos[loc["own_gid"]]
        # reaches of the stripe: the river-tree bands are whole inside a stripe and numbered band by band
        nr = int(loc["Nr"][0])
        rsel = np.arange(nr) + (0 if r0 == 0 else Nr_w - nr)
        want = np.concatenate([ref[sel], ref[Ne_w + sel], ref[2 * Ne_w + sel], ref[3 * Ne_w + rsel]])
        p = bench.oracle_parity(loc, want, 1)
        assert p["n_bad"] == 0 and p["max_rel"] == 0.0 and p["n"] == want.size, p
        bad = want.copy(); bad[5] += 1e-6 * max(abs(bad[5]), 1e-6)
        assert bench.oracle_parity(loc, bad, 1)["n_bad"] == 1
    # whole mesh (N=1 form)
    p = bench.oracle_parity(whole, ref, 2)
    assert p["n_bad"] == 0 and p["max_rel"] == 0.0
