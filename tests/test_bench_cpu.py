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


def test_collection_rows_are_sharding_independent():
    import bench
    q = bench.collection_query(128, 5)
    whole = bench.collection_rows(0, 3 * bench.DOC_BLOCK + 17, 128, 5, q)
    assert np.allclose(np.linalg.norm(whole, axis=1), 1.0, atol=1e-5)
    for lo, hi in [(0, 1), (bench.DOC_BLOCK - 3, bench.DOC_BLOCK + 5), (100, 2 * bench.DOC_BLOCK + 1), (7, 7)]:
        assert np.array_equal(bench.collection_rows(lo, hi, 128, 5, q), whole[lo:hi])
    w = bench.weak_collection_rows(500, 2100, 1000)
    assert w.shape == (1600, 128)
    assert np.array_equal(w[:500], bench.synthetic_docs(1000, bench.DATA_SEED + 1)[2][500:])
    assert np.array_equal(w[1500:], bench.synthetic_docs(1000, bench.DATA_SEED + 3)[2][:100])


def test_both_arms_print_the_same_config_object():
    """`config` is a pure function of the command line and the circuit: the driver compares the two arms' objects."""
    import argparse
    import bench
    model, _ = bench.build_model()
    c = model.model.fhe_circuit
    for gpus, wl in [(1, "config2"), (2, "config4"), (8, "config4")]:
        a = argparse.Namespace(workload="auto", gpus=gpus, docs=1000, total_docs=bench.TOTAL_DOCS_CONFIG4)
        assert bench.resolve_workload(a) == wl
        cfg = bench.workload_config(a, c)
        assert cfg == bench.workload_config(argparse.Namespace(**vars(a)), c) and "workload" in cfg
        assert ("multi_gpu" in cfg) == (gpus > 1)
        assert cfg["bytes_per_comparison"] == (128 + 2) * (c.lwe.n + 1) * 8


def test_pbs_dispatch_labels_follow_the_dispatcher():
    """pbs_bench._mb2_kernels mirrors csrc/pbs.cu::launch_pbs_mb2: full waves of 4 x SMs ciphertexts on the throughput
    kernel, a remainder on the cluster kernel up to SMs / 2, on the one-CTA latency kernel up to 3 x SMs, else on the
    throughput kernel."""
    from fhe_icp_b200.pbs_bench import _mb2_kernels
    sm = 148
    assert _mb2_kernels(1, sm).startswith("pbs_kernel_mb2_pair x 1")
    assert _mb2_kernels(74, sm).startswith("pbs_kernel_mb2_pair x 74")
    assert _mb2_kernels(75, sm).startswith("pbs_kernel_mb2_wide x 75")
    assert _mb2_kernels(444, sm).startswith("pbs_kernel_mb2_wide x 444")
    assert _mb2_kernels(445, sm).startswith("pbs_kernel_mb2<1,4> x 445")
    assert _mb2_kernels(592, sm).startswith("pbs_kernel_mb2<1,4> x 592") and "+" not in _mb2_kernels(592, sm)
    both = _mb2_kernels(612, sm)
    assert "pbs_kernel_mb2<1,4> x 592" in both and "pbs_kernel_mb2_pair x 20" in both
    assert "pbs_kernel_mb2_wide x 148" in _mb2_kernels(740, sm)


def test_cpu_pbs_baseline_is_correct_on_a_small_sample():
    """bench.py's `pbs.cpu_baseline`: the oracle's keyswitch + multi-bit PBS at the stated parameter set on one ciphertext
    per host thread; every output must decrypt to the table entry."""
    import bench
    from fhe_icp_b200.params import PBS_PARAMS_4BIT as P
    params = {k: P[k] for k in ("n", "k", "N_poly", "l_pbs", "beta_pbs", "l_ks", "beta_ks", "log2_sigma_lwe", "log2_sigma_glwe")}
    r = bench.cpu_pbs_reference(params, per_thread=1)
    assert r["all_correct"] and r["value"] > 0 and r["pbs_per_sec"] >= r["value"] and r["kind"] == "port" and r["cores"] >= 1
