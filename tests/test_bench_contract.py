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
