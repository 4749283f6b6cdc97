#!/usr/bin/env python3
"""Run the reference's OWN, UNMODIFIED code through this repo's estimators and record every estimator call.

    python tests/golden/make_reference_traces.py          # rewrites tests/golden/reference_traces.npz

What runs (all from /root/reference, read-only, nothing is copied):
  * fhe_similarity.test_fhe_similarity()                      (fhe_similarity.py:227-294)
  * test_fhe.py, executed as a script                         (test_fhe.py:10-64)
  * test_fhe_workflow.test_fhe_similarity_workflow()          (test_fhe_workflow.py:8-116)
  * quantization_strategy.QuantizationTester.test_bit_width on create_similarity_dataset() for 4 / 8 / 12 bits
                                                              (quantization_strategy.py:17-90,134-160,171-173)
with `concrete.ml.sklearn` resolved by tests/ref_shim to fhe_icp_b200's estimators (the two-line import change of
INTEGRATION.md, done by sys.path instead of by editing the reference).  np.random is seeded before each script (the
reference draws unseeded).  In this container there is no GPU, so the shim evaluates fhe="execute" on the CPU oracle;
the recorded INPUTS are what matters: tests/test_gpu_reference_replay.py feeds them to the CUDA path on the B200 and
asserts the reference's own checks there.  tests/test_reference_shim.py re-runs this recording and compares it with the
committed file, so the fixture cannot drift from the reference or from the estimators.
"""
from __future__ import annotations

import contextlib
import io
import os
import runpy
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
REFERENCE = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "reference_traces.npz"
SEEDS = {"fhe_similarity": 101, "test_fhe": 102, "test_fhe_workflow": 103}


def _paths():
    for p in (str(ROOT), str(ROOT / "tests" / "ref_shim"), str(REFERENCE)):
        if p not in sys.path:
            sys.path.insert(0, p)


def record() -> dict:
    """-> {script: {"estimators": [trace dict, ...], "results": {...}}}; needs /root/reference."""
    _paths()
    from concrete.ml import sklearn as shim           # tests/ref_shim
    assert shim.__file__.startswith(str(ROOT)), "concrete.ml.sklearn must resolve to tests/ref_shim"
    out = {}
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp, contextlib.redirect_stdout(io.StringIO()):
        os.chdir(tmp)                                  # the reference writes fhe_similarity_model.pkl into the cwd
        try:
            import fhe_similarity as ref_fs
            assert Path(ref_fs.__file__).parent == REFERENCE
            shim.TRACE.clear()
            np.random.seed(SEEDS["fhe_similarity"])
            ref_fs.test_fhe_similarity()
            out["fhe_similarity"] = {"estimators": list(shim.TRACE), "results": {}}

            shim.TRACE.clear()
            np.random.seed(SEEDS["test_fhe"])
            g = runpy.run_path(str(REFERENCE / "test_fhe.py"), run_name="__main__")
            out["test_fhe"] = {"estimators": list(shim.TRACE),
                               "results": {"clear_pred": np.asarray(g["clear_pred"]), "fhe_pred": np.asarray(g["fhe_pred"]),
                                           "tolerance": float(g["tolerance"])}}

            import test_fhe_workflow as ref_wf
            shim.TRACE.clear()
            np.random.seed(SEEDS["test_fhe_workflow"])
            ref_wf.test_fhe_similarity_workflow()
            out["test_fhe_workflow"] = {"estimators": list(shim.TRACE), "results": {}}

            import quantization_strategy as ref_qs
            shim.TRACE.clear()
            X, y = ref_qs.create_similarity_dataset(n_samples=500, dim=128)     # seeds itself with 42
            split = int(0.8 * len(X))
            tester = ref_qs.QuantizationTester()
            res = {}
            for n_bits in (4, 8, 12):
                r = tester.test_bit_width(X[:split], y[:split], X[split:], y[split:], n_bits)
                res[str(n_bits)] = {"status": r["status"], "circuit_max_bits": r["memory"].get("circuit_max_bits"),
                                    "clear_vs_fhe_mae": r["metrics"].get("clear_vs_fhe_mae"),
                                    "r2_score": r["metrics"].get("r2_score"), "error": r.get("error")}
            out["quantization_strategy"] = {"estimators": list(shim.TRACE), "results": res}
        finally:
            os.chdir(cwd)
    return out


def flatten(traces: dict) -> dict:
    """-> flat {key: array} for np.savez (object-free, so the file loads with allow_pickle=False).  Arrays are stored
    once under a content hash (the three quantization estimators share one training set; compile inputsets are
    prefixes of it), the JSON index refers to them as {"blob": hash, "rows": n}."""
    import hashlib
    import json
    flat, index = {}, {}

    def put(a):
        a = np.ascontiguousarray(a)
        for key, (h, full) in list(seen.items()):          # a row-prefix of something already stored?
            if full.dtype == a.dtype and full.shape[1:] == a.shape[1:] and len(a) <= len(full) and a.ndim >= 1 \
                    and np.array_equal(full[: len(a)], a):
                return {"blob": h, "rows": int(len(a))}
        h = hashlib.sha1(a.tobytes() + str((a.dtype, a.shape)).encode()).hexdigest()[:16]
        flat["blob/" + h] = a
        seen[h] = (h, a)
        return {"blob": h, "rows": int(len(a)) if a.ndim else 0}

    seen = {}
    for script, t in traces.items():
        ests = []
        for e in t["estimators"]:
            rec = {"cls": e["cls"], "args": e["args"], "kwargs": e["kwargs"], "max_bits": e["max_bits"],
                   "fit_X": put(e["fit"][0]), "fit_y": put(e["fit"][1]), "compile_X": put(e["compile"]), "calls": [],
                   # the fitted floats: LAPACK / SGD may round differently on another host CPU, and at the reference's
                   # float32 fits the last bit of coef_ decides the weight quantizer
                   "coef": put(e["coef"][0]), "intercept": repr(e["coef"][1])}
            for c in e["calls"]:
                rec["calls"].append({"fhe": c["fhe"], "X": put(c["X"]), "y": put(c["y"])})
            ests.append(rec)
        results = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in t["results"].items()}
        index[script] = {"estimators": ests, "results": results}
    flat["index_json"] = np.frombuffer(json.dumps(index).encode(), dtype=np.uint8)
    return flat


class Traces:
    """Reader of reference_traces.npz: ``index`` (JSON) + ``get(ref)`` resolving a {"blob", "rows"} reference."""

    def __init__(self, path=OUT):
        import json
        self.z = np.load(path, allow_pickle=False)
        self.index = json.loads(bytes(self.z["index_json"]).decode())

    def get(self, ref):
        a = self.z["blob/" + ref["blob"]]
        return a[: ref["rows"]] if a.ndim else a


if __name__ == "__main__":
    flat = flatten(record())
    np.savez_compressed(OUT, **flat)
    print(f"wrote {OUT} ({OUT.stat().st_size / 1e6:.2f} MB, {len(flat)} arrays)")
