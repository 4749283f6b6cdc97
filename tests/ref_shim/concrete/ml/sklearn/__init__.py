"""TEST INFRASTRUCTURE.  `from concrete.ml.sklearn import LinearRegression, SGDRegressor` resolves here when
tests/ref_shim is on sys.path: the reference's unmodified code then drives THIS repo's estimators
(fhe_icp_b200.linear_model) through the exact call sequence it uses on Concrete-ML
(fit / score / compile / predict(X[, fhe="execute"]) / coef_ / intercept_ / fhe_circuit.graph...).

Two additions, both for the tests only:
  * every call is recorded (arguments and results) in `TRACE`, from which tests/golden/make_reference_traces.py
    builds the committed fixture the GPU test replays (the reference sources cannot travel to the GPU box);
  * `fhe="execute"` runs on the B200 through the product's C-ABI whenever a CUDA device is present.  In THIS
    container there is none, so -- in the tests only -- the call is evaluated by the CPU ORACLE (oracle/, the checker),
    which lets the reference's scripts run to their own assertions here.  The product itself has no such path:
    fhe_icp_b200 raises without a CUDA device.
"""
from __future__ import annotations

import numpy as np

from fhe_icp_b200 import linear_model as _lm

TRACE: list = []          # one dict per estimator, in construction order


def _cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # pragma: no cover
        return False


def _oracle_execute(est, X) -> np.ndarray:
    """encrypt -> encrypted dot product -> decrypt of the compiled circuit, on the CPU oracle (checker)."""
    from oracle import oracle as O
    c = est.fhe_circuit
    spec = c.spec
    X = np.asarray(X)
    q = spec.input_q.quant(X)
    s = O.secret_key(c.key_seed, 2, c.lwe.n)
    base = c.next_ct_base(q.size)
    ct = O.lwe_encrypt(s, q, c.lwe.shift, c.lwe.sigma_abs, c.enc_seed, ct_base=base, stride=c.lwe.stride,
                       noise_seed=c.noise_seed)
    W = np.stack([spec.q_weights, np.ones_like(spec.q_weights)]) if c.two_outputs else spec.q_weights[None]
    out = O.lincomb(ct.reshape(len(X), spec.d, -1), W, c.lwe.n)
    m = O.lwe_decrypt(s, out, c.lwe.shift)
    qy = m[:, 0] - (int(spec.weight_q.zero_point) * m[:, 1] if c.two_outputs else 0) + int(spec.q_bias)
    return spec.dequantize_output(qy)


class _Recorded:
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._rec = {"cls": type(self).__name__, "args": list(args), "kwargs": dict(kwargs), "calls": []}
        TRACE.append(self._rec)

    def fit(self, X, y):
        self._rec["fit"] = (np.array(X, copy=True), np.array(y, copy=True))
        r = super().fit(X, y)
        self._rec["coef"] = (np.array(self.coef_, copy=True), float(self.intercept_))
        return r

    def compile(self, X_sample, **kw):
        self._rec["compile"] = np.array(X_sample, copy=True)
        circuit = super().compile(X_sample, **kw)
        self._rec["max_bits"] = int(circuit.graph.maximum_integer_bit_width())
        return circuit

    def predict(self, X, fhe="disable"):
        mode = getattr(fhe, "value", fhe)
        if mode == "execute" and not _cuda():
            if self.fhe_circuit is None:
                raise RuntimeError("The model is not compiled. Call compile() before fhe='execute'.")
            y = _oracle_execute(self, X)
            backend = "oracle"
        else:
            y = super().predict(X, fhe=fhe)
            backend = "b200" if mode == "execute" else "clear"
        self._rec["calls"].append({"X": np.array(X, copy=True), "fhe": mode, "y": np.array(y, copy=True),
                                   "backend": backend})
        return y


class LinearRegression(_Recorded, _lm.LinearRegression):
    pass


class SGDRegressor(_Recorded, _lm.SGDRegressor):
    pass
