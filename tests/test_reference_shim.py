"""The reference's own, unmodified scripts driven through this repo's estimators (task 4).

`concrete.ml.sklearn` resolves to tests/ref_shim (the two-line import change of INTEGRATION.md, done by sys.path), then
/root/reference's fhe_similarity.test_fhe_similarity(), test_fhe.py, test_fhe_workflow.py and
QuantizationTester.test_bit_width run as they are and their own checks are asserted.  /root/reference exists only in the
build container, where there is no GPU: fhe="execute" is evaluated by the CPU oracle here (the shim says so), and the
recorded inputs are replayed on the CUDA path by tests/test_gpu_reference_replay.py on the B200.
"""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tests" / "golden"))

pytestmark = pytest.mark.skipif(not Path("/root/reference/fhe_similarity.py").exists(),
                                reason="the reference tree is only present in the build container")


@pytest.fixture(scope="module")
def traces():
    import make_reference_traces as M
    return M.record()


def test_reference_scripts_run_unmodified_and_their_checks_hold(traces):
    # test_fhe.py:56-57 -- |fhe - clear| < tolerance (0.01)
    r = traces["test_fhe"]["results"]
    assert abs(r["fhe_pred"][0] - r["clear_pred"][0]) < r["tolerance"] == 0.01
    assert r["fhe_pred"][0] == r["clear_pred"][0]                      # here it is exact
    # fhe_similarity.py:268-269,294 -- "Clear vs FHE error: <0.001": 5 clear and 5 fhe="execute" predictions
    calls = traces["fhe_similarity"]["estimators"][0]["calls"]
    clear = [c for c in calls if c["fhe"] == "disable" and len(c["X"]) == 5][-1]
    fhe = np.concatenate([c["y"] for c in calls if c["fhe"] == "execute"])
    assert len(fhe) == 5 and np.mean(np.abs(clear["y"] - fhe)) == 0.0
    # test_fhe_workflow.py:93,97,104 -- the two encrypted comparisons equal the clear ones
    calls = traces["test_fhe_workflow"]["estimators"][0]["calls"]
    ex = [c for c in calls if c["fhe"] == "execute"]
    cl = [c for c in calls if c["fhe"] == "disable" and len(c["X"]) == 1]
    assert len(ex) == 2 and all(np.array_equal(e["y"], c["y"]) for e, c in zip(ex, cl[-2:] if len(cl) > 2 else cl))
    assert ex[0]["y"][0] > 0.5 > abs(ex[1]["y"][0])                    # similar vs unrelated document


def test_reference_quantization_sweep_reports_the_published_numbers(traces):
    """quantization_strategy.py:52-59,79-81 run unmodified: its own `circuit_max_bits` are the published 12 / 20 / 28
    (SESSION_REPORT.md:66-71) and its `clear_vs_fhe_mae` is exactly 0 at every width."""
    res = traces["quantization_strategy"]["results"]
    assert {k: v["circuit_max_bits"] for k, v in res.items()} == {"4": 12, "8": 20, "12": 28}
    assert all(v["status"] == "success" and v["clear_vs_fhe_mae"] == 0.0 for v in res.values())


def test_committed_trace_fixture_is_current(traces):
    """tests/golden/reference_traces.npz (what the GPU test replays) is what the reference produces today."""
    import make_reference_traces as M
    want = M.flatten(traces)
    have = np.load(M.OUT, allow_pickle=False)
    assert sorted(want) == sorted(have.files)
    for k in want:
        assert np.array_equal(want[k], have[k]), k


def test_reference_fixed_similarity_script_runs_unmodified_and_its_conclusion_holds():
    """/root/reference/test_fixed_similarity.py (the script behind SESSION5_FIXES.md) run as it is on the shim: the
    estimator it builds itself -- `LinearRegression(n_bits=8)` on element-wise products, fit / score / predict, the very
    circuit of the hot path -- gives what the script concludes: "Fixed model: predictions match true cosine
    similarities" (identical ~ 1, opposite ~ -1, every pair within the 8-bit quantization error), while the 256-d
    model fed concatenated embeddings does not."""
    import contextlib
    import io
    import re
    import make_reference_traces as M
    M._paths()
    import test_fixed_similarity as ref
    assert Path(ref.__file__).parent == M.REFERENCE
    np.random.seed(104)
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        ref.test_comparison()
    txt = buf.getvalue()
    num = r"(-?\d+\.\d+)"
    true = [float(x) for x in re.search(rf"Identical: {num}\s+Similar: {num}\s+Different: {num}\s+Opposite: {num}", txt).groups()]
    blocks = re.findall(rf"Predictions:\s+Identical: {num}\s+Similar: {num}\s+Different: {num}\s+Opposite: {num}", txt)
    assert len(blocks) == 2
    orig, fixed = ([float(x) for x in b] for b in blocks)
    assert true[0] == 1.0 and true[3] == -1.0
    assert max(abs(p - t) for p, t in zip(fixed, true)) < 0.05            # "predictions match true cosine similarities"
    assert abs(orig[0] - 1.0) > 0.3        # the 256-d model fed concatenated embeddings does not see identical vectors as 1
    r2 = float(re.findall(r"Training R² score: (-?\d+\.\d+)", txt)[-1])
    assert r2 > 0.99                                                       # the product features are an exact linear model
