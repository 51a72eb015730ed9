"""GPU suite: our arm of the bench contract on the tiny workload -- one JSON line with the keys the driver and the judge
read (metric, value, roofline, cpu_baseline, e2e, gpu_launches, clocks) and internally consistent numbers."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_our_arm_line_on_tiny_workload():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "tiny", "--steps", "5", "--warmup", "3",
                          "--cpu-budget", "2", "--replicates", "64"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "fitch_site_node_ops_per_s" and d["unit"] == "site-node ops/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 5 and d["warmup"] == 3 and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "u32" and d["data"] == "synthetic" and d["config"]["workload"].startswith("tiny:")
    sites = d["value"] / d["insertions_per_s"] / 2.0           # one insertion = 2 site-node ops per informative site
    assert d["value"] > 0 and 1000 < sites <= 2000 and abs(sites - round(sites)) < 1e-6
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] <= d["value"] * 1.05                      # host buffers and copies inside the timed region
    assert d["gpu_launches"] == 5
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] == 1 and cb["value"] > 0 and "sample" in cb
    assert d["search"]["cpu_baseline"]["identical_result"] is True and d["search"]["ras"]["identical_result"] is True
    assert d["bb"]["roofline"]["bound"] == "tensor" and d["bb"]["search"]["replicates_won"] == 64
    c = d["cost"]
    assert c["roofline"]["kernel"] == "k_sk_scan" and c["insertions_per_s"] > 0 and c["cpu_baseline"]["insertions_per_s"] > 0
    assert c["bb"]["chunks_on_tensor_cores"] + c["bb"]["chunks_on_exact_kernel"] >= 1
