"""Estimators that quack like the Concrete-ML ones the reference imports
(``from concrete.ml.sklearn import SGDRegressor, LinearRegression``,
/root/reference/fhe_similarity.py:5): ``fit / score / compile / predict(X, fhe=...) /
coef_ / intercept_ / fhe_circuit.graph.maximum_integer_bit_width()``.  Callers reach through
``FHESimilarityModel.model`` to these members (/root/reference/batch_operations.py:233,276;
fhe_similarity.py:94,98,120,129-130,151,191-192).

``predict(X)`` / ``fhe="disable"`` is the clear quantized circuit in numpy, exactly as in the
reference.  ``fhe="execute"`` runs encrypt -> encrypted dot product -> decrypt on the B200
through the C-ABI; there is no CPU substitute for it.
"""
from __future__ import annotations

import ctypes as C
import time
from typing import Optional

import numpy as np

from .params import DEFAULT_P_ERROR, LweParams, select_lwe_params
from .quantization import QuantizedLinearSpec, signed_bit_width

from .randomness import CiphertextIds, seed_or_fresh


class _Graph:
    def __init__(self, max_bits: int):
        self._max_bits = int(max_bits)

    def maximum_integer_bit_width(self) -> int:
        return self._max_bits


class FHECircuit:
    """Compiled circuit: integer bounds, crypto parameters and (lazily) the device model."""

    def __init__(self, spec: QuantizedLinearSpec, inputset_q: np.ndarray, p_error: float, bound_mode: str,
                 key_seed: Optional[int], enc_seed: Optional[int], device: Optional[int], noise_seed: Optional[int] = None,
                 ct_start: Optional[int] = None):
        self.spec = spec
        self.p_error = p_error
        # randomness.py: key_seed and noise_seed are client secrets, enc_seed is the PUBLIC mask seed that travels with
        # seeded ciphertexts; all default to the OS CSPRNG, fixed values are for reproducible tests and benchmarks
        self.key_seed = seed_or_fresh(key_seed)
        self.noise_seed = seed_or_fresh(noise_seed)
        self.enc_seed = seed_or_fresh(enc_seed)
        self.device = device
        self.ids = CiphertextIds(ct_start)     # never-reused ciphertext ids (random 62-bit origin per process)
        # "seeded": fresh ciphertexts travel as 8-byte bodies + public mask seed (the evaluator regenerates
        # the masks); "expanded": full (n+1)-word ciphertexts are materialised in HBM.  Same results.
        self.ciphertext_format = "seeded"
        self._sim = None
        self._sim_ctx = None
        zp_w = int(spec.weight_q.zero_point)
        self.two_outputs = zp_w != 0
        # --- integer bounds over the inputset (what Concrete's compiler measures) [EXT]
        dot = inputset_q @ spec.q_weights
        ssum = inputset_q.sum(axis=1)
        widths = [signed_bit_width(inputset_q.min(), inputset_q.max()),
                  signed_bit_width(spec.q_weights.min(), spec.q_weights.max()),
                  signed_bit_width(dot.min(), dot.max())]
        if self.two_outputs:
            widths.append(signed_bit_width(ssum.min(), ssum.max()))
        self.inputset_bits = max(widths)
        # --- guaranteed bounds over every representable input (the quantizer clips)
        amax = max(abs(spec.input_q.qmin), abs(spec.input_q.qmax))
        worst_dot = int(np.abs(spec.q_weights).sum()) * amax
        worst = max(worst_dot, spec.d * amax if self.two_outputs else 0, amax)
        self.guaranteed_bits = signed_bit_width(-worst, worst)
        self.bound_mode = bound_mode
        msg_bits = self.guaranteed_bits if bound_mode == "guaranteed" else self.inputset_bits
        w2 = float((spec.q_weights.astype(np.float64) ** 2).sum())
        self.lwe: LweParams = select_lwe_params(msg_bits, max(w2, float(spec.d)), p_error)
        full = spec.circuit(inputset_q)
        self.clear_postprocess_bits = signed_bit_width(min(full.min(), -abs(zp_w) * amax * spec.d),
                                                       max(full.max(), abs(zp_w) * amax * spec.d))
        # --- what graph.maximum_integer_bit_width() reports: the widest node of the FULL integer circuit
        # q_X @ q_W - zp_W * sum(q_X) + q_bias over the inputset -- inputs, constants and every intermediate,
        # as Concrete's graph does [EXT].  The bias constant is the wide node on the reference's configurations:
        # it reproduces the published 12 / 20 / 28 bits of /root/reference/SESSION_REPORT.md:66-71
        # (tests/test_host.py::test_published_circuit_bit_widths).  Crypto parameters are NOT sized from this
        # number: the client adds zp_W * (.) and q_bias after decryption, so only `inputset_bits` /
        # `guaranteed_bits` (the encrypted values) enter the parameter selection.
        zsum = zp_w * ssum
        nodes = [(inputset_q.min(), inputset_q.max()), (spec.q_weights.min(), spec.q_weights.max()),
                 (dot.min(), dot.max()), (ssum.min(), ssum.max()), (zp_w, zp_w), (zsum.min(), zsum.max()),
                 ((dot - zsum).min(), (dot - zsum).max()), (int(spec.q_bias), int(spec.q_bias)),
                 (full.min(), full.max())]
        self.circuit_bits = max(signed_bit_width(lo, hi) for lo, hi in nodes)
        self.graph = _Graph(self.circuit_bits)

    # ------------------------------------------------------------------ device model
    def native_spec(self):
        from . import _native as N
        s = self.spec
        return N.SimilaritySpec(
            d=s.d, n_bits=s.input_q.n_bits, n=self.lwe.n, stride=self.lwe.stride, shift=self.lwe.shift,
            two_outputs=1 if self.two_outputs else 0, sigma_abs=self.lwe.sigma_abs, x_scale=float(s.input_q.scale),
            x_zero_point=int(s.input_q.zero_point), x_offset=int(s.input_q.offset),
            w_zero_point=int(s.weight_q.zero_point), q_bias=int(s.q_bias), out_scale=float(s.out_scale),
            out_zero_point=int(s.out_zero_point), key_seed=self.key_seed, noise_seed=self.noise_seed)

    def keygen(self, force: bool = False):
        """Create the device-side model (secret key from ``key_seed``, weights).  Lazy, like
        Concrete's ``keygen(force=False)`` inside ``encrypt_run_decrypt``."""
        if self._sim is not None and not force:
            return self
        from . import _native as N
        self.release()
        ctx = N.context(self.device)
        h = C.c_void_p()
        spec = self.native_spec()
        qw = np.ascontiguousarray(self.spec.q_weights, dtype=np.int64)
        N.check(N.lib().fhe_b200_similarity_create(ctx.handle, C.byref(spec), qw.ctypes.data_as(C.POINTER(C.c_int64)),
                                                   C.byref(h)))
        self._sim, self._sim_ctx = h, ctx
        return self

    @property
    def handle(self):
        self.keygen()
        return self._sim

    def release(self):
        from . import _native as N
        if self._sim is not None:
            N.lib().fhe_b200_similarity_destroy(self._sim)
            self._sim = None
        if getattr(self, "_eval", None) is not None:
            N.lib().fhe_b200_similarity_destroy(self._eval)
            self._eval = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def next_ct_base(self, count: int) -> int:
        return self.ids.take(count)

    @property
    def ct_counter(self) -> int:
        return self.ids.next

    @ct_counter.setter
    def ct_counter(self, value: int):
        self.ids.next = int(value)

    def evaluator_handle(self):
        """Server side: the device model WITHOUT secret material (fhe_b200_similarity_create_evaluator).  It can run
        the encrypted dot products; encrypt / decrypt on it fail."""
        if getattr(self, "_eval", None) is None:
            from . import _native as N
            ctx = N.context(self.device)
            h = C.c_void_p()
            spec = self.native_spec()
            spec.key_seed = 0
            spec.noise_seed = 0
            qw = np.ascontiguousarray(self.spec.q_weights, dtype=np.int64)
            N.check(N.lib().fhe_b200_similarity_create_evaluator(ctx.handle, C.byref(spec),
                                                                 qw.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(h)))
            self._eval, self._eval_ctx = h, ctx
        return self._eval

    @property
    def wire32_supported(self) -> bool:
        """Whether the scores may travel in the 32-bit wire form without raising the decoding failure probability
        above p_error (the modulus switch adds (n/2 + 1) * 2^64 / 12 to the noise variance)."""
        import math
        from .params import z_score
        w2 = max(float((self.spec.q_weights.astype(np.float64) ** 2).sum()), float(self.spec.d) if self.two_outputs else 0.0)
        var = (self.lwe.sigma_abs ** 2) * w2 + (self.lwe.n / 2 + 1) * (2.0 ** 64) / 12.0
        return self.lwe.shift >= 32 and z_score(self.p_error) * math.sqrt(var) < 2.0 ** (self.lwe.shift - 1)

    def encrypt_run_decrypt(self, X: np.ndarray, return_q: bool = False):
        """The reference's ``predict(..., fhe="execute")`` with HOST buffers: float32 rows in,
        float64 scores out; quantize/encrypt/dot/decrypt/dequantize all run on the device."""
        from . import _native as N
        X = np.asarray(X)
        if X.dtype != np.float32:
            # the device quantizer reads float32; any other dtype is quantized on the host in that dtype -- the very
            # expression predict_clear evaluates -- so that fhe="execute" == fhe="disable" for float64 inputs too
            # (rint(x / scale + zp) can flip at a rounding boundary when x is first rounded to float32)
            return self._run_quantized(self.spec.input_q.quant(X.reshape(-1, self.spec.d)), return_q)
        X = np.ascontiguousarray(X, dtype=np.float32).reshape(-1, self.spec.d)
        B = X.shape[0]
        y = np.empty(B, dtype=np.float64)
        qy = np.empty(B, dtype=np.int64)
        base = self.next_ct_base(B * self.spec.d)
        fn = (N.lib().fhe_b200_similarity_predict_host_seeded if self.ciphertext_format == "seeded"
              else N.lib().fhe_b200_similarity_predict_host)
        N.check(fn(
            self.handle, X.ctypes.data_as(C.POINTER(C.c_float)), B, self.enc_seed, base,
            y.ctypes.data_as(C.POINTER(C.c_double)), qy.ctypes.data_as(C.POINTER(C.c_int64))))
        return (y, qy) if return_q else y


def _run_quantized(self, q: np.ndarray, return_q: bool = False):
    """encrypt -> dot -> decrypt of already quantized rows (int64 [B, d]), seeded ciphertexts, all on the device."""
    import torch
    from . import _native as N
    from . import engine as E
    dev = torch.device("cuda", N.context(self.device).device)
    B = q.shape[0]
    if getattr(self, "_key_t", None) is None or self._key_t.device != dev:
        self._key_t = E.secret_key(self.key_seed, 2, self.lwe.n, dev)
    base = self.next_ct_base(B * self.spec.d)
    bodies = E.lwe_encrypt_seeded(self._key_t, torch.as_tensor(np.ascontiguousarray(q, dtype=np.int64)), self.lwe.shift,
                                  self.lwe.sigma_abs, self.enc_seed, base, noise_seed=self.noise_seed)
    M = 2 if self.two_outputs else 1
    out = torch.empty((B, M, self.lwe.stride), dtype=torch.int64, device=dev)
    y = torch.empty(B, dtype=torch.float64, device=dev)
    qy = torch.empty(B, dtype=torch.int64, device=dev)
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    N.check(N.lib().fhe_b200_similarity_run_seeded(self.handle, C.c_void_p(bodies.data_ptr()), B, self.enc_seed, base,
                                                   C.c_void_p(out.data_ptr()), st))
    N.check(N.lib().fhe_b200_similarity_decrypt(self.handle, C.c_void_p(out.data_ptr()), B, C.c_void_p(y.data_ptr()),
                                                C.c_void_p(qy.data_ptr()), st))
    return (y.cpu().numpy(), qy.cpu().numpy()) if return_q else y.cpu().numpy()


FHECircuit._run_quantized = _run_quantized


class LinearRegression:
    """Quantized ordinary-least-squares regressor (Concrete-ML ``LinearRegression`` surface)."""

    def __init__(self, n_bits: int = 8, fit_intercept: bool = True):
        self.n_bits = int(n_bits)
        self.fit_intercept = fit_intercept
        self.coef_ = None
        self.intercept_ = None
        self.spec: Optional[QuantizedLinearSpec] = None
        self.fhe_circuit: Optional[FHECircuit] = None

    # ------------------------------------------------------------------ training (clear, setup)
    def _fit_float(self, X: np.ndarray, y: np.ndarray):
        # OLS on centred data, computed in the dtype of X (sklearn keeps float32 inputs in
        # float32, which is what the reference's float32 generator feeds it).
        # The solver is scipy.linalg.lstsq (LAPACK gelsd), the routine sklearn's LinearRegression
        # -- and therefore Concrete-ML's -- calls; numpy's lstsq is the stand-in without scipy.
        # (In float32 the two differ in the last bits of coef_, which decides whether the weight
        # quantizer is degenerate (q_W = 1) or asymmetric with a large zero-point)
        dt = X.dtype if X.dtype in (np.float32, np.float64) else np.float64
        X = X.astype(dt, copy=False)
        y = y.astype(dt, copy=False)
        try:
            from scipy.linalg import lstsq as _lstsq
            self.solver_ = "scipy.linalg.lstsq"
            solve = lambda A, b: _lstsq(A, b)[0]  # noqa: E731
        except ImportError:  # pragma: no cover
            self.solver_ = "numpy.linalg.lstsq"
            solve = lambda A, b: np.linalg.lstsq(A, b, rcond=None)[0]  # noqa: E731
        if self.fit_intercept:
            xm, ym = X.mean(axis=0), y.mean()
            coef = solve(X - xm, y - ym)
            intercept = ym - xm @ coef
        else:
            coef = solve(X, y)
            intercept = 0.0
        return coef, intercept

    def fit(self, X, y):
        X = np.asarray(X)
        y = np.asarray(y).reshape(-1)
        if X.ndim != 2 or X.shape[0] != y.shape[0]:
            raise ValueError("X must be [n_samples, n_features] and y [n_samples]")
        coef, intercept = self._fit_float(X, y)
        self.coef_ = np.asarray(coef)
        self.intercept_ = float(intercept)
        self.spec = QuantizedLinearSpec.from_fit(self.coef_, self.intercept_, X, self.n_bits)
        self.fhe_circuit = None
        return self

    def _check_fitted(self):
        if self.spec is None:
            raise RuntimeError("The model is not fitted. Call fit() first.")

    def score(self, X, y) -> float:
        """R^2 of the (clear, quantized) predictions."""
        y = np.asarray(y, dtype=np.float64).reshape(-1)
        pred = self.predict(X)
        ss_res = float(((y - pred) ** 2).sum())
        ss_tot = float(((y - y.mean()) ** 2).sum())
        return 1.0 - ss_res / ss_tot if ss_tot > 0 else 0.0

    # ------------------------------------------------------------------ compile
    def compile(self, X_sample, p_error: float = DEFAULT_P_ERROR, bound_mode: str = "guaranteed",
                key_seed: Optional[int] = None, enc_seed: Optional[int] = None, device: Optional[int] = None,
                noise_seed: Optional[int] = None, ct_start: Optional[int] = None):
        """Bound the integer circuit on the calibration inputset and choose crypto parameters.  Seeds default to the
        OS CSPRNG (like Concrete's key generation); fixed seeds / ``ct_start`` are for reproducible tests."""
        self._check_fitted()
        if bound_mode not in ("guaranteed", "inputset"):
            raise ValueError("bound_mode must be 'guaranteed' or 'inputset'")
        X_sample = np.asarray(X_sample)
        if X_sample.ndim != 2 or X_sample.shape[1] != self.spec.d:
            raise ValueError(f"inputset must be [rows, {self.spec.d}]")
        q = self.spec.input_q.quant(X_sample)
        self.fhe_circuit = FHECircuit(self.spec, q, p_error, bound_mode, key_seed, enc_seed, device,
                                      noise_seed=noise_seed, ct_start=ct_start)
        return self.fhe_circuit

    # ------------------------------------------------------------------ inference
    def quantize_input(self, X) -> np.ndarray:
        self._check_fitted()
        return self.spec.input_q.quant(np.asarray(X))

    def dequantize_output(self, q_y) -> np.ndarray:
        self._check_fitted()
        return self.spec.dequantize_output(q_y)

    def predict(self, X, fhe: str = "disable") -> np.ndarray:
        self._check_fitted()
        X = np.asarray(X)
        if X.ndim != 2 or X.shape[1] != self.spec.d:
            raise ValueError(f"expected input of shape [rows, {self.spec.d}], got {X.shape}")
        mode = getattr(fhe, "value", fhe)
        if mode in ("disable", "simulate"):
            # "simulate" has no separate noise model here: the parameters are chosen so that
            # the executed result equals the clear integer circuit (p_error 2^-40).
            return self.spec.predict_clear(X)
        if mode != "execute":
            raise ValueError(f"fhe must be 'disable', 'simulate' or 'execute', got {fhe!r}")
        if self.fhe_circuit is None:
            raise RuntimeError("The model is not compiled. Call compile() before fhe='execute'.")
        return self.fhe_circuit.encrypt_run_decrypt(X)


class SGDRegressor(LinearRegression):
    """``SGDRegressor(n_bits, max_iter, random_state)`` surface used by
    /root/reference/quantization_strategy.py:34-38.  Training is clear-side setup: plain
    averaged-free SGD on the squared loss with sklearn's default 'invscaling' schedule."""

    def __init__(self, n_bits: int = 8, max_iter: int = 1000, random_state: Optional[int] = None,
                 alpha: float = 1e-4, eta0: float = 0.01, power_t: float = 0.25, tol: float = 1e-3,
                 fit_intercept: bool = True):
        super().__init__(n_bits=n_bits, fit_intercept=fit_intercept)
        self.max_iter, self.random_state = int(max_iter), random_state
        self.alpha, self.eta0, self.power_t, self.tol = alpha, eta0, power_t, tol

    def _fit_float(self, X, y):
        # Concrete-ML's SGDRegressor is a wrapper around sklearn's: use that very fit when sklearn is importable
        # (same coefficients as the reference for a given random_state); the plain SGD below is the stand-in.
        try:
            from sklearn.linear_model import SGDRegressor as _SkSGD
        except ImportError:  # pragma: no cover
            _SkSGD = None
        if _SkSGD is not None:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")   # ConvergenceWarning at the reference's max_iter=20
                sk = _SkSGD(max_iter=self.max_iter, random_state=self.random_state, alpha=self.alpha, eta0=self.eta0,
                            power_t=self.power_t, tol=self.tol, fit_intercept=self.fit_intercept).fit(X, y)
            self.solver_ = "sklearn.linear_model.SGDRegressor"
            return np.asarray(sk.coef_, dtype=np.float64), float(np.ravel(sk.intercept_)[0])
        self.solver_ = "builtin-sgd"
        rng = np.random.RandomState(self.random_state)
        X = X.astype(np.float64)
        y = y.astype(np.float64)
        nrow, d = X.shape
        w, b, t = np.zeros(d), 0.0, 1.0
        best, bad = np.inf, 0
        for _ in range(self.max_iter):
            order = rng.permutation(nrow)
            loss = 0.0
            for i in order:
                eta = self.eta0 / (t ** self.power_t)
                err = X[i] @ w + b - y[i]
                loss += 0.5 * err * err
                w *= 1.0 - eta * self.alpha
                w -= eta * err * X[i]
                if self.fit_intercept:
                    b -= eta * err
                t += 1.0
            loss /= nrow
            if loss > best - self.tol:
                bad += 1
                if bad >= 5:
                    break
            else:
                bad = 0
            best = min(best, loss)
        return w, b


def timed(fn, *a, **kw):
    t0 = time.time()
    r = fn(*a, **kw)
    return r, time.time() - t0
