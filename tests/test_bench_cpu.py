"""bench.py's CPU legs (the oracle port timed as cpu_baseline / --impl reference) on a tiny sample, and
the JSON contract of the reference arm -- no GPU."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_cpu_reference_sample_is_exact_and_bounded():
    import bench
    model, _ = bench.build_model()
    _, _, X = bench.synthetic_docs(96, 3)
    value, rows, t, threads = bench.cpu_reference(model, X, 0.2)
    assert value > 0 and rows >= 96 and threads >= 1 and t < 60


def test_top_k_matches_reference_semantics():
    import bench
    s = np.array([0.5, 0.9, 0.5, 0.2, 0.9])
    assert bench.top_k(s, 3, 0.5) == [(1, 0.9), (4, 0.9), (0, 0.5)]
    assert bench.top_k(s, 3, 2.0) == []


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--docs", "64", "--cpu-seconds", "0.2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "encrypted_comparisons_per_sec" and line["value"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["kind"] == "port"
    assert "workload" in line["config"]
