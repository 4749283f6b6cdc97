"""Drop-in for the reference's ``fhe_similarity.py``: same class, methods, metrics keys and
error behaviour (/root/reference/fhe_similarity.py:12-224), with the Concrete-ML CPU backend
replaced by the B200 engine.  Additions (the split client/server API the reference only
planned): ``seed=`` for reproducible data, ``keygen / encrypt / run / decrypt``.
"""
from __future__ import annotations

import os
import pickle
import time
from typing import Optional, Tuple

import numpy as np

from .linear_model import LinearRegression, SGDRegressor  # noqa: F401


class SeededCiphertexts:
    """Compressed fresh ciphertexts: ``bodies`` [B,d] u64 (int64 view) + the public (enc_seed, ct_base)
    from which the evaluator regenerates every mask word (ciphertext (b,j) has id ct_base + b*d + j)."""

    def __init__(self, bodies, enc_seed: int, ct_base: int):
        self.bodies, self.enc_seed, self.ct_base = bodies, int(enc_seed), int(ct_base)

    @property
    def shape(self):
        return tuple(self.bodies.shape)

    def nbytes(self) -> int:
        return self.bodies.numel() * 8


class FHESimilarityModel:
    """Optimized FHE model for similarity computation."""

    def __init__(self, input_dim: int = 256, n_bits: int = 8, similarity_type: str = 'cosine',
                 seed: Optional[int] = None, key_seed: Optional[int] = None, enc_seed: Optional[int] = None,
                 device: Optional[int] = None, verbose: bool = True, noise_seed: Optional[int] = None,
                 ct_start: Optional[int] = None):
        """``seed`` seeds the synthetic training data only.  ``key_seed`` / ``noise_seed`` (client secrets) and
        ``enc_seed`` (public mask seed) default to the OS CSPRNG, as Concrete's key generation does; fixed values and
        ``ct_start`` (first ciphertext id) are an opt-in for reproducible tests and benchmarks (randomness.py)."""
        self.input_dim = input_dim
        self.n_bits = n_bits
        self.similarity_type = similarity_type
        self.model = None
        self.compiled = False
        self.metrics = {}
        self.seed = seed
        self.key_seed, self.enc_seed, self.device = key_seed, enc_seed, device
        self.noise_seed, self.ct_start = noise_seed, ct_start
        self.verbose = verbose

    def _log(self, msg: str):
        if self.verbose:
            print(msg)

    # ---------------------------------------------------------------- data (fhe_similarity.py:34-70)
    def _prepare_training_data(self, n_samples: int = 1000) -> Tuple[np.ndarray, np.ndarray]:
        """Synthetic unit-norm pairs, X = emb1 * emb2, y = sum(X).  Draw order matches the
        reference, so ``seed=S`` reproduces what the reference generates after ``np.random.seed(S)``."""
        self._log(f"Generating {n_samples} training samples...")
        rng = np.random if self.seed is None else np.random.RandomState(self.seed)
        single_dim = self.input_dim
        emb1 = rng.randn(n_samples, single_dim).astype(np.float32)
        emb1 = emb1 / np.linalg.norm(emb1, axis=1, keepdims=True)
        emb2 = rng.randn(n_samples, single_dim).astype(np.float32)
        emb2 = emb2 / np.linalg.norm(emb2, axis=1, keepdims=True)
        mask = rng.rand(n_samples) > 0.5
        emb2[mask] = emb1[mask] + 0.2 * rng.randn(int(mask.sum()), single_dim)
        emb2 = emb2 / np.linalg.norm(emb2, axis=1, keepdims=True)
        X = emb1 * emb2
        if self.similarity_type == 'cosine':
            y = np.sum(emb1 * emb2, axis=1)
        elif self.similarity_type == 'dot':
            y = np.sum(emb1 * emb2, axis=1)
        elif self.similarity_type == 'manhattan':
            y = -np.sum(np.abs(emb1 - emb2), axis=1)
            y = (y - y.min()) / (y.max() - y.min())
        else:
            raise ValueError(f"Unknown similarity type: {self.similarity_type}")
        return X, y

    # ---------------------------------------------------------------- train (fhe_similarity.py:72-106)
    def train(self, X_train: Optional[np.ndarray] = None, y_train: Optional[np.ndarray] = None,
              n_samples: int = 1000):
        self._log("\nTraining FHE Similarity Model")
        self._log(f"  Input dimension: {self.input_dim}")
        self._log(f"  Quantization: {self.n_bits} bits")
        self._log(f"  Similarity type: {self.similarity_type}")
        if X_train is None or y_train is None:
            X_train, y_train = self._prepare_training_data(n_samples)
        self.model = LinearRegression(n_bits=self.n_bits)
        start = time.time()
        self.model.fit(X_train, y_train)
        train_time = time.time() - start
        train_score = self.model.score(X_train, y_train)
        self.metrics['train_time'] = train_time
        self.metrics['train_score'] = float(train_score)
        self._log(f"  Training completed in {train_time:.2f}s")
        self._log(f"  Training R² score: {train_score:.4f}")
        return X_train, y_train

    # ---------------------------------------------------------------- compile (fhe_similarity.py:108-140)
    def compile(self, X_sample: np.ndarray, **compile_kwargs):
        if self.model is None:
            raise RuntimeError("Model not trained. Call train() first.")
        start = time.time()
        start_memory = self._get_memory_usage()
        try:
            kw = dict(key_seed=self.key_seed, enc_seed=self.enc_seed, device=self.device, noise_seed=self.noise_seed,
                      ct_start=self.ct_start)
            kw.update(compile_kwargs)
            self.model.compile(X_sample, **kw)
            compile_time = time.time() - start
            end_memory = self._get_memory_usage()
            self.compiled = True
            self.metrics['compile_time'] = compile_time
            self.metrics['compile_memory_mb'] = end_memory - start_memory
            if hasattr(self.model, 'fhe_circuit'):
                max_bits = self.model.fhe_circuit.graph.maximum_integer_bit_width()
                self.metrics['circuit_max_bits'] = int(max_bits)
                self._log(f"  Circuit max bit-width: {max_bits}")
            c = self.model.fhe_circuit
            self._log(f"  LWE parameters: n={c.lwe.n}, log2(sigma)={c.lwe.log2_sigma:.2f}, "
                      f"Delta=2^{c.lwe.shift}, outputs={2 if c.two_outputs else 1}")
        except Exception as e:
            self._log(f"  Compilation failed: {str(e)}")
            raise

    # ---------------------------------------------------------------- predict (fhe_similarity.py:142-167)
    def predict_encrypted(self, X: np.ndarray) -> np.ndarray:
        """Predict using FHE execution.  The reference loops row by row; here every row of X
        goes through one batched encrypt -> dot -> decrypt pass on the GPU (same results)."""
        if not self.compiled:
            raise RuntimeError("Model not compiled. Call compile() first.")
        X = np.asarray(X)
        if len(X) == 0:
            return np.array([])
        start = time.time()
        pred = self.model.predict(X, fhe="execute")
        pred_time = time.time() - start
        # the reference records the first prediction of EVERY call (fhe_similarity.py:156-158)
        self._log(f"  First FHE prediction batch ({len(X)} rows) took {pred_time:.3f}s")
        self.metrics['fhe_prediction_time'] = pred_time
        return np.asarray(pred)

    def predict_clear(self, X: np.ndarray) -> np.ndarray:
        if self.model is None:
            raise RuntimeError("Model not trained.")
        return self.model.predict(X)

    # ---------------------------------------------------------------- split client / server API
    def keygen(self, key_seed: Optional[int] = None):
        if not self.compiled:
            raise RuntimeError("Model not compiled. Call compile() first.")
        c = self.model.fhe_circuit
        if key_seed is not None and key_seed != c.key_seed:
            c.key_seed = int(key_seed)
            c.keygen(force=True)
        else:
            c.keygen()
        return self

    def encrypt(self, X: np.ndarray, seeded: bool = False):
        """Client: float rows [B,d] -> device ciphertext tensor [B,d,stride] (int64 view of u64), or --
        ``seeded=True`` -- a :class:`SeededCiphertexts` (8-byte bodies + the public mask seed/ids)."""
        import ctypes as C
        import torch
        from . import _native as N
        c = self.keygen().model.fhe_circuit
        dev = torch.device("cuda", N.context(self.device).device)
        Xd = torch.as_tensor(np.ascontiguousarray(X, dtype=np.float32)).reshape(-1, c.spec.d).to(dev)
        B = Xd.shape[0]
        if seeded:
            bodies = torch.empty((B, c.spec.d), dtype=torch.int64, device=dev)
            base = c.next_ct_base(B * c.spec.d)
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            N.check(N.lib().fhe_b200_similarity_encrypt_seeded(c.handle, C.c_void_p(Xd.data_ptr()), B, c.enc_seed, base,
                                                               C.c_void_p(bodies.data_ptr()), st))
            return SeededCiphertexts(bodies, c.enc_seed, base)
        ct = torch.empty((B, c.spec.d, c.lwe.stride), dtype=torch.int64, device=dev)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        N.check(N.lib().fhe_b200_similarity_encrypt(c.handle, C.c_void_p(Xd.data_ptr()), B, c.enc_seed,
                                                    c.next_ct_base(B * c.spec.d), C.c_void_p(ct.data_ptr()), st))
        return ct

    def encrypt_products(self, query: np.ndarray, docs):
        """Client: seeded ciphertexts of X = query * docs -- the clear product the reference forms on the host before
        every circuit call (batch_operations.py:226,273) -- with the multiply done by the encryption kernel (IEEE
        float32, bit-identical to numpy's).  ``docs``: float32 rows [B,d] as a numpy array, a (pinned) host tensor --
        uploaded here -- or a tensor already resident on this GPU."""
        import ctypes as C
        import torch
        from . import _native as N
        c = self.keygen().model.fhe_circuit
        dev = torch.device("cuda", N.context(self.device).device)
        if not isinstance(docs, torch.Tensor):
            docs = torch.from_numpy(np.ascontiguousarray(docs, dtype=np.float32))
        if docs.dtype != torch.float32 or docs.dim() != 2 or docs.shape[1] != c.spec.d:
            raise ValueError(f"docs must be float32 [rows, {c.spec.d}]")
        Dd = docs.to(dev, non_blocking=True).contiguous()
        qd = torch.from_numpy(np.ascontiguousarray(query, dtype=np.float32).reshape(c.spec.d)).to(dev)
        B = Dd.shape[0]
        bodies = torch.empty((B, c.spec.d), dtype=torch.int64, device=dev)
        base = c.next_ct_base(B * c.spec.d)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        N.check(N.lib().fhe_b200_similarity_encrypt_seeded_products(c.handle, C.c_void_p(qd.data_ptr()), C.c_void_p(Dd.data_ptr()),
                                                                    B, c.enc_seed, base, C.c_void_p(bodies.data_ptr()), st))
        return SeededCiphertexts(bodies, c.enc_seed, base)

    def run(self, ct, out=None):
        """Server: ciphertexts [B,d,stride] -> encrypted scores [B,M,stride] (M = 1 or 2)."""
        import ctypes as C
        import torch
        from . import _native as N
        if not self.compiled:
            raise RuntimeError("Model not compiled. Call compile() first.")
        c = self.model.fhe_circuit
        handle = c.evaluator_handle()      # the evaluator holds no key material: run() is the server's whole job
        M = 2 if c.two_outputs else 1
        if isinstance(ct, SeededCiphertexts):
            B, dev = ct.bodies.shape[0], ct.bodies.device
            if out is None:
                out = torch.empty((B, M, c.lwe.stride), dtype=torch.int64, device=dev)
            st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            N.check(N.lib().fhe_b200_similarity_run_seeded(handle, C.c_void_p(ct.bodies.data_ptr()), B, ct.enc_seed,
                                                           ct.ct_base, C.c_void_p(out.data_ptr()), st))
            return out
        B = ct.shape[0]
        if out is None:
            out = torch.empty((B, M, c.lwe.stride), dtype=torch.int64, device=ct.device)
        st = C.c_void_p(torch.cuda.current_stream(ct.device).cuda_stream)
        N.check(N.lib().fhe_b200_similarity_run(handle, C.c_void_p(ct.data_ptr()), B, C.c_void_p(out.data_ptr()), st))
        return out

    def decrypt(self, out, return_q: bool = False):
        """Client: encrypted scores -> float64 similarities (and optionally the integers q_y)."""
        import ctypes as C
        import torch
        from . import _native as N
        c = self.keygen().model.fhe_circuit
        B = out.shape[0]
        y = torch.empty(B, dtype=torch.float64, device=out.device)
        qy = torch.empty(B, dtype=torch.int64, device=out.device)
        st = C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
        N.check(N.lib().fhe_b200_similarity_decrypt(c.handle, C.c_void_p(out.data_ptr()), B,
                                                    C.c_void_p(y.data_ptr()), C.c_void_p(qy.data_ptr()), st))
        y_np = y.cpu().numpy()
        return (y_np, qy.cpu().numpy()) if return_q else y_np

    def compress_scores(self, out, out32=None):
        """Server: encrypted scores [B,M,stride] u64 -> 32-bit wire form [B,M,stride] (int32 view of u32)."""
        import ctypes as C
        import torch
        from . import _native as N
        if not self.compiled:
            raise RuntimeError("Model not compiled. Call compile() first.")
        c = self.model.fhe_circuit
        if not c.wire32_supported:
            raise ValueError("the 32-bit wire form would raise the decoding failure probability above p_error at these "
                             "parameters; keep the 64-bit scores")
        if out32 is None:
            out32 = torch.empty(out.shape, dtype=torch.int32, device=out.device)
        st = C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
        N.check(N.lib().fhe_b200_lwe_modswitch32(N.context(self.device).handle, C.c_void_p(out.data_ptr()),
                                                 out.shape[0] * out.shape[1], c.lwe.stride, C.c_void_p(out32.data_ptr()), st))
        return out32

    def decrypt_compressed(self, out32, return_q: bool = False, to_host: bool = True):
        """Client: scores in the 32-bit wire form -> float64 similarities."""
        import ctypes as C
        import torch
        from . import _native as N
        c = self.keygen().model.fhe_circuit
        B = out32.shape[0]
        y = torch.empty(B, dtype=torch.float64, device=out32.device)
        qy = torch.empty(B, dtype=torch.int64, device=out32.device)
        st = C.c_void_p(torch.cuda.current_stream(out32.device).cuda_stream)
        N.check(N.lib().fhe_b200_similarity_decrypt32(c.handle, C.c_void_p(out32.data_ptr()), B,
                                                      C.c_void_p(y.data_ptr()), C.c_void_p(qy.data_ptr()), st))
        if not to_host:
            return (y, qy) if return_q else y
        return (y.cpu().numpy(), qy.cpu().numpy()) if return_q else y.cpu().numpy()

    @property
    def wire32_supported(self) -> bool:
        return bool(self.compiled and self.model.fhe_circuit.wire32_supported)

    @property
    def dev(self):
        import torch
        from . import _native as N
        return torch.device("cuda", N.context(self.device).device)

    # ---------------------------------------------------------------- misc (fhe_similarity.py:169-224)
    def _get_memory_usage(self) -> float:
        try:
            import psutil
            return psutil.Process(os.getpid()).memory_info().rss / 1024 / 1024
        except Exception:
            return 0.0

    def save(self, path: str):
        """Save hyper-parameters, fitted coefficients and the full quantizer spec."""
        data = {
            'input_dim': self.input_dim,
            'n_bits': self.n_bits,
            'similarity_type': self.similarity_type,
            'metrics': self.metrics,
            'model_params': {
                'coef_': self.model.coef_ if self.model is not None else None,
                'intercept_': self.model.intercept_ if self.model is not None else None,
            },
            'quantized_spec': self.model.spec.to_dict() if self.model is not None and self.model.spec else None,
        }
        with open(path, 'wb') as f:
            pickle.dump(data, f)
        self._log(f"Model parameters saved to {path}")

    @classmethod
    def load(cls, path: str) -> 'FHESimilarityModel':
        """Load a saved model.  Unlike the reference (which must retrain), the quantized spec
        is restored; ``compiled`` is still False after loading, call compile() again."""
        from .quantization import QuantizedLinearSpec
        with open(path, 'rb') as f:
            data = pickle.load(f)
        model = cls(input_dim=data['input_dim'], n_bits=data['n_bits'], similarity_type=data['similarity_type'])
        model.metrics = data['metrics']
        model.compiled = False
        if data.get('quantized_spec') is not None:
            est = LinearRegression(n_bits=model.n_bits)
            est.coef_ = data['model_params']['coef_']
            est.intercept_ = data['model_params']['intercept_']
            est.spec = QuantizedLinearSpec.from_dict(data['quantized_spec'])
            model.model = est
        return model
