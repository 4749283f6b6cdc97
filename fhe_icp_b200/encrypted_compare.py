"""Encrypted x encrypted comparison: BOTH embedding vectors are encrypted.

The reference multiplies the two embeddings in the clear before the FHE model sees them
(`batch_operations.py:226` `X = (emb1 * emb2).reshape(1, -1)`, `:273` in the search loop) and
concedes that "both documents must be available to compute products" (`SESSION5_FIXES.md:118-120`,
`test_fhe_workflow.py:108-112`).  Here the product itself is evaluated under encryption:

    x*y = floor((x+y)^2 / 4) - floor((x-y)^2 / 4)          (x+y and x-y have the same parity)

so one comparison of two d-dimensional vectors is 2d programmable bootstraps (table lookups of the
quarter square) on sums / differences of the two parties' ciphertexts, followed by a wrapping sum
of the 2d outputs into ONE score ciphertext per document, decrypted by the client.

Each party knows its own vector, so it can also send an encryption of its own squared norm; then

    2 * sum_j x_j*y_j = sum_j (x_j+y_j)^2 - sum_j x_j^2 - sum_j y_j^2

needs only d bootstraps (of the square table) per comparison.  That is the default protocol
(`encrypt_norms`, `scores(..., norm_q, norm_docs)`); both produce the same score encoding.

Quantization is of the FACTORS (signed `in_bits`-bit integers, symmetric scale), not of the product
as in the reference, so the scores are not bit-comparable to `FHESimilarityModel`'s; the parity
oracle for this path is its own clear integer model `sum_j xq_j * yq_j` (`compare_clear`) and the
CPU restatement `oracle.encrypted_product_scores`.  Decrypted scores equal the clear integer model
exactly; see DESIGN.md for the noise budget that makes this hold (l_pbs = 2).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _native as N
from . import engine as E
from .randomness import CiphertextIds, seed_or_fresh

# n=742, N=2048 as the stated 4-bit set, but two decomposition levels in the bootstrap: one PBS output
# carries noise std ~2^-22 of the torus (2^-15 with one level), so the sum of 2d = 256 outputs
# (std ~2^-18) stays 16 sigma inside the half-step 2^-14 of the 13-bit signed score.
COMPARE_PARAMS = dict(n=742, k=1, N_poly=2048, l_pbs=2, beta_pbs=15, l_ks=5, beta_ks=3,
                      log2_sigma_lwe=-17.1, log2_sigma_glwe=-51.6)
P_BITS = 4          # PBS message space: in_bits + 1 (sum/difference), offset-binary; plus one padding bit
IN_BITS = 3         # signed factors in [-4, 3]
IN_SHIFT = 63 - P_BITS
SCORE_BITS = 13     # signed score range [-4096, 4095] >= d * 16 for d <= 256
OUT_SHIFT = 64 - SCORE_BITS


def quarter_square_table(p_bits: int = P_BITS) -> np.ndarray:
    w = np.arange(1 << p_bits, dtype=np.int64) - (1 << (p_bits - 1))
    return (w * w) // 4


def square_table(p_bits: int = P_BITS) -> np.ndarray:
    w = np.arange(1 << p_bits, dtype=np.int64) - (1 << (p_bits - 1))
    return w * w


NORM_CT_BASE = 1 << 40   # ciphertext ids of the squared-norm encryptions (disjoint from the vectors')


class EncryptedCompare:
    """Client + server halves of the encrypted x encrypted cosine score.

    client: `keygen`, `quantize`, `encrypt`, `decrypt`, `dequantize`;  server: `scores` (needs only
    the Fourier bootstrapping key).  `multibit=True` uses the two-bits-per-step blind rotation
    (l_pbs <= 2)."""

    def __init__(self, input_dim: int = 128, params: dict | None = None, key_seed: int | None = None,
                 evk_seed: int | None = None, device=None, multibit: bool = True, chunk_pbs: int = 148 * 2 * 32,
                 noise_seed: int | None = None, enc_seed: int | None = None, ct_start: int | None = None):
        """Seeds default to the OS CSPRNG (randomness.py); pass fixed values only for reproducible tests.  ``enc_seed``
        is the public mask seed; ``ct_start`` the first ciphertext id (None: random origin)."""
        if input_dim * 16 >= (1 << (SCORE_BITS - 1)):
            raise ValueError("input_dim too large for the 13-bit score range")
        self.d = int(input_dim)
        self.pd = dict(params or COMPARE_PARAMS)
        self.p = E.make_pbs_params(**self.pd)
        self.dev = E._dev(device)
        self.key_seed, self.evk_seed = seed_or_fresh(key_seed), seed_or_fresh(evk_seed)
        self.noise_seed, self.enc_seed = seed_or_fresh(noise_seed), seed_or_fresh(enc_seed)
        self.ids = CiphertextIds(ct_start)
        self.multibit = bool(multibit)
        if self.multibit and self.p.l_pbs > 2:
            raise ValueError("multi-bit blind rotation is implemented for l_pbs <= 2")
        self.chunk_pbs = int(chunk_pbs)
        self.scale = None
        self.s = self.S = self.bskf = None
        self.lut = E.from_u64_numpy(E.make_lut_poly(quarter_square_table(), P_BITS, self.p.N, OUT_SHIFT), self.dev)
        # one-bootstrap protocol: (w-8)^2 at half the score unit, so that sum - norms = 2*sum(xy) * 2^(OUT_SHIFT-1)
        self.lut_sq = E.from_u64_numpy(E.make_lut_poly(square_table(), P_BITS, self.p.N, OUT_SHIFT - 1), self.dev)

    # ---- client
    def keygen(self) -> "EncryptedCompare":
        p = self.p
        self.s = E.secret_key(self.key_seed, 0, p.n, self.dev)
        self.S = E.secret_key(self.key_seed, 1, p.k * p.N, self.dev)
        if self.multibit:
            bsk = E.bsk2_gen(p, self.s, self.S, self.evk_seed)
            self.bskf = E.bsk2_to_fourier(p, bsk)
        else:
            bsk = E.bsk_gen(p, self.s, self.S, self.evk_seed)
            self.bskf = E.bsk_to_fourier(p, bsk)
        del bsk
        return self

    def fit_scale(self, X: np.ndarray, clip_sigmas: float = 2.5) -> float:
        """Symmetric scale for the signed IN_BITS quantizer: clip at `clip_sigmas` standard deviations."""
        sd = float(np.asarray(X, dtype=np.float64).std())
        self.scale = clip_sigmas * sd / (1 << (IN_BITS - 1)) if sd > 0 else 1.0
        return self.scale

    def quantize(self, X: np.ndarray) -> np.ndarray:
        if self.scale is None:
            raise RuntimeError("Quantizer not calibrated. Call fit_scale() first.")
        lo, hi = -(1 << (IN_BITS - 1)), (1 << (IN_BITS - 1)) - 1
        return np.clip(np.rint(np.asarray(X, dtype=np.float64) / self.scale), lo, hi).astype(np.int64)

    def dequantize(self, q_scores: np.ndarray) -> np.ndarray:
        return np.asarray(q_scores, dtype=np.float64) * (self.scale * self.scale)

    def encrypt(self, Xq: np.ndarray, enc_seed: int | None = None, ct_base: int | None = None) -> torch.Tensor:
        """Xq int [.., d] -> small-key ciphertexts [.., d, stride] (one LWE per dimension).  ``enc_seed`` (public mask
        seed) defaults to the engine's; ``ct_base=None`` takes fresh, never-reused ciphertext ids."""
        self._need_keys()
        m = torch.as_tensor(np.ascontiguousarray(Xq, dtype=np.int64))
        base = self.ids.take(m.numel()) if ct_base is None else ct_base
        return E.lwe_encrypt(self.s, m, IN_SHIFT, self.p.sigma_lwe_abs, self.enc_seed if enc_seed is None else enc_seed,
                             base, noise_seed=self.noise_seed)

    def encrypt_norms(self, Xq: np.ndarray, enc_seed: int | None = None, ct_base: int | None = None) -> torch.Tensor:
        """Xq int [.., d] -> big-key ciphertexts [.., kN+2] of sum_j xq_j^2 at 2^(OUT_SHIFT-1)."""
        self._need_keys()
        Xq = np.asarray(Xq, dtype=np.int64)
        m = torch.as_tensor(np.ascontiguousarray((Xq * Xq).sum(axis=-1)))
        base = self.ids.take(m.numel()) if ct_base is None else NORM_CT_BASE + ct_base
        return E.lwe_encrypt(self.S, m, OUT_SHIFT - 1, self.p.sigma_glwe_abs,
                             self.enc_seed if enc_seed is None else enc_seed, base, noise_seed=self.noise_seed)

    def decrypt(self, scores: torch.Tensor) -> np.ndarray:
        self._need_keys()
        v = E.lwe_decrypt(self.S, scores, OUT_SHIFT).cpu().numpy() & ((1 << SCORE_BITS) - 1)
        return np.where(v >= (1 << (SCORE_BITS - 1)), v - (1 << SCORE_BITS), v)

    # ---- server
    def scores(self, ct_query: torch.Tensor, ct_docs: torch.Tensor, norm_query: torch.Tensor | None = None,
               norm_docs: torch.Tensor | None = None) -> torch.Tensor:
        """ct_query [d, stride], ct_docs [B, d, stride] -> [B, kN+2] encrypted sum_j x_j*y_j (big key;
        ciphertext = first kN+1 words of a row).  With the two parties' encrypted squared norms
        (norm_query [kN+2], norm_docs [B, kN+2]) one bootstrap per dimension, otherwise two."""
        if self.bskf is None:
            raise RuntimeError("No evaluation key. Call keygen() (or load one) first.")
        p, d = self.p, self.d
        B = ct_docs.shape[0]
        words = p.n + 1
        out = torch.empty((B, E.even_stride(p.k * p.N)), dtype=torch.int64, device=self.dev)
        docs_per_chunk = max(1, self.chunk_pbs // (2 * d))
        fn = E.pbs_mb2 if self.multibit else E.pbs
        if (norm_query is None) != (norm_docs is None):
            raise ValueError("give both squared-norm ciphertexts or neither")
        if norm_query is not None:
            docs_per_chunk *= 2
            nq = norm_query.reshape(-1).contiguous()
        for b0 in range(0, B, docs_per_chunk):
            b1 = min(B, b0 + docs_per_chunk)
            if norm_query is None:
                pairs = E.pair_addsub(ct_query, ct_docs[b0:b1], words, 1 << 62)
                sq = fn(p, self.bskf, pairs.view(-1, words), self.lut)
                E.pair_diff_sum(sq.view(b1 - b0, d, 2, -1), out[b0:b1])
            else:
                sums = E.pair_add(ct_query, ct_docs[b0:b1], words, 1 << 62)
                sq = fn(p, self.bskf, sums.view(-1, words), self.lut_sq)
                E.square_sum(sq.view(b1 - b0, d, -1), nq, norm_docs[b0:b1], out[b0:b1])
        return out

    # ---- whole pipeline, float in / float out
    def compare_clear(self, q: np.ndarray, docs: np.ndarray) -> np.ndarray:
        """The clear integer model this path must reproduce exactly: sum_j xq_j * yq_j."""
        return self.quantize(docs) @ self.quantize(q)

    def similarity(self, q: np.ndarray, docs: np.ndarray, enc_seed: int | None = None, norms: bool = True) -> np.ndarray:
        docs = np.atleast_2d(docs)
        xq, yq = self.quantize(q), self.quantize(docs)
        ct_q = self.encrypt(xq, enc_seed)          # fresh ciphertext ids on every call
        ct_d = self.encrypt(yq, enc_seed)
        if not norms:
            return self.dequantize(self.decrypt(self.scores(ct_q, ct_d)))
        nq, nd = self.encrypt_norms(xq, enc_seed), self.encrypt_norms(yq, enc_seed)
        return self.dequantize(self.decrypt(self.scores(ct_q, ct_d, nq, nd)))

    def _need_keys(self):
        if self.s is None:
            raise RuntimeError("No secret key. Call keygen() first.")


BIT_SHIFT = 60      # threshold results: one bit at 2^60, so up to 7 of them can be added (score buckets)


class EncryptedThreshold:
    """`score >= T` evaluated under encryption, exactly.

    The reference thresholds decrypted scores in the clear (`batch_operations.py:278`
    `if similarity >= min_similarity`) and buckets them at 0.9 / 0.7 / 0.5 (`fhe_cli.py:169-176`).  Here
    the server returns only the encrypted outcome.  A single N=2048 bootstrap resolves ~5 message
    bits, the score has 13, so the sign of v = score - T is taken by LSB-first bit extraction: for
    i = 0..11 the ciphertext is scaled by 2^(12-i) (bit i moves to the top of the torus, the bits above
    wrap away, the bits below are already cleared), offset by 1/4, keyswitched to the small key and
    bootstrapped with a constant (sign) test polynomial, which returns bit i at its own weight with
    fresh noise; that is subtracted.  What remains is the sign bit; a 13th bootstrap maps it to
    (score >= T) * 2^60.  Every step has a decision margin of 1/4 of the torus against keyswitch +
    mod-switch noise of ~2^-8.6, so the result is exact for every score and threshold."""

    def __init__(self, ec: EncryptedCompare, score_bits: int = SCORE_BITS, out_shift: int = OUT_SHIFT,
                 scale: float | None = None):
        self.ec = ec
        self.scale = scale          # quantizer scale of the scores' producer (defaults to ec.scale)
        self.score_bits, self.out_shift = int(score_bits), int(out_shift)
        if self.score_bits + self.out_shift != 64:
            raise ValueError("the score must fill the torus: score_bits + out_shift == 64")
        p = ec.p
        ec._need_keys()
        # keyswitching key laid out for the tensor-core contraction (ks_mma.cu); bit-identical to KS32
        self.ksk32 = E.ksk_to_32(p, E.ksk_gen(p, ec.S, ec.s, ec.evk_seed))
        self.key_mma = E.ksk_to_mma(p, self.ksk32)
        self._ks_work = None
        mask = (1 << 64) - 1
        consts = [(-(1 << (self.out_shift - 1 + i))) & mask for i in range(self.score_bits - 1)] + [1 << (BIT_SHIFT - 1)]
        self.luts = E.from_u64_numpy(np.repeat(np.array(consts, dtype=np.uint64)[:, None], p.N, axis=1), ec.dev)

    def threshold_to_int(self, min_similarity: float) -> int:
        """Smallest integer score whose dequantized value is >= min_similarity."""
        lim = 1 << (self.score_bits - 1)
        scale = self.scale if self.scale is not None else self.ec.scale
        return int(np.clip(np.ceil(min_similarity / (scale ** 2) - 1e-9), -lim // 2, lim // 2 - 1))

    def ge(self, scores: torch.Tensor, T: int) -> torch.Tensor:
        """scores [B, stride] (big key) -> [B, kN+2] encrypting (score >= T) * 2^BIT_SHIFT."""
        ec, p = self.ec, self.ec.p
        words = p.k * p.N + 1
        fn = E.pbs_mb2 if ec.multibit else E.pbs
        acc = E.shl_add(scores.contiguous(), words, 0, -(int(T) << self.out_shift))     # v = score - T
        for i in range(self.score_bits):
            last = i == self.score_bits - 1
            tmp = E.shl_add(acc, words, self.score_bits - 1 - i, 1 << 62)                  # bit i on top, + 1/4
            pb = fn(p, ec.bskf, self._keyswitch(tmp), self.luts[i])
            if last:   # -/+ 2^59 for sign bit 0/1 -> (1 - sign) * 2^60
                return E.shl_add(pb, words, 0, 1 << (BIT_SHIFT - 1), out_stride=E.even_stride(words - 1))
            E.sub_plain(acc, pb, 1 << (self.out_shift - 1 + i))                            # clear bit i

    def _keyswitch(self, ct: torch.Tensor) -> torch.Tensor:
        need = int(N.lib().fhe_b200_keyswitch_mma_workspace_bytes(C.byref(self.ec.p), ct.shape[0]))
        if self._ks_work is None or self._ks_work.numel() < need:
            self._ks_work = torch.empty(need, dtype=torch.int8, device=ct.device)
        return E.keyswitch_mma(self.ec.p, self.key_mma, ct, work=self._ks_work)

    def buckets(self, scores: torch.Tensor, thresholds) -> torch.Tensor:
        """Number of thresholds each score reaches (e.g. 0.5 / 0.7 / 0.9 -> 0..3), one ciphertext per score."""
        out = None
        for T in thresholds:
            g = self.ge(scores, T)
            out = g if out is None else E.accumulate(out, g)
        return out

    def decrypt(self, bits: torch.Tensor) -> np.ndarray:
        return E.lwe_decrypt(self.ec.S, bits, BIT_SHIFT).cpu().numpy() & 15


# ================================================================================================ packed (leveled) variant
# GLWE side as above (k=1, N=2048) but 2 x 18-bit decomposition levels.  The dominant noise of an external
# product with a BINARY key is (N/4) * (2^-2*l*beta / 12) * (sum_j Q_j)^2: the key's non-zero mean correlates
# the rounding errors of all coefficients, so a constant-sign query is the worst case ((sum Q)^2 = d^2 * 2^8
# at 5-bit factors).  With 36 decomposed bits that worst case is std 2^-22.4 of the torus (measured, oracle and
# GPU) against a decoding margin of 2^-18: exact for EVERY input, not just zero-mean embeddings.  (At 2 x 15
# bits the worst case reaches 2^-17.4, inside the margin of a 15-bit score -- rejected.)
PACKED_PARAMS = dict(COMPARE_PARAMS, beta_pbs=18)
PACKED_IN_BITS = 5                     # signed factors in [-16, 15]
PACKED_SCORE_BITS = 17                 # |sum_j x_j*y_j| <= 128 * 256 = 2^15
PACKED_OUT_SHIFT = 64 - PACKED_SCORE_BITS


def pack_documents(Yq: np.ndarray, N_poly: int, slot: int) -> np.ndarray:
    """int [B, d] -> message polynomials int64 [ceil(B / (N/slot)), N]: document b of a group occupies
    coefficients slot*b .. slot*b + d - 1."""
    Yq = np.asarray(Yq, dtype=np.int64)
    B, d = Yq.shape
    per = N_poly // slot
    G = (B + per - 1) // per
    out = np.zeros((G * per, slot), dtype=np.int64)
    out[:B, :d] = Yq
    return out.reshape(G, N_poly)


def query_polynomial(xq: np.ndarray, N_poly: int) -> np.ndarray:
    """Q(X) = sum_j x_j X^(-j) = x_0 - sum_{j>=1} x_j X^(N-j): coefficient m of Q*D is sum_j x_j D_{m+j}."""
    xq = np.asarray(xq, dtype=np.int64)
    Q = np.zeros(N_poly, dtype=np.int64)
    Q[0] = xq[0]
    Q[N_poly - np.arange(1, xq.size)] = -xq[1:]
    return Q


class PackedEncryptedCompare:
    """Both vectors encrypted, no bootstrap: documents packed N/slot per GLWE ciphertext, the query a GGSW
    encryption of Q(X) = sum_j x_j X^(-j); ONE external product per GLWE yields N/slot inner products
    (coefficient slot*b of the product).  At d = 128: 16 comparisons per external product, 2 KB of
    ciphertext per document.  The noise of an external product grows with |Q|_2 instead of being reset,
    which 5-bit factors and two 18-bit decomposition levels absorb (worst case std 2^-22.4 against a
    decoding margin of 2^-18, see PACKED_PARAMS), so decrypted scores equal sum_j xq_j*yq_j exactly.

    client: keygen / fit_scale / quantize / encrypt_documents / encrypt_query / decrypt / dequantize;
    server: scores (needs no key material at all -- the GGSW IS the query)."""

    def __init__(self, input_dim: int = 128, params: dict | None = None, key_seed: int | None = None, device=None,
                 noise_seed: int | None = None, enc_seed: int | None = None, ct_start: int | None = None):
        """Seeds default to the OS CSPRNG (randomness.py); pass fixed values only for reproducible tests."""
        self.d = int(input_dim)
        self.pd = dict(params or PACKED_PARAMS)
        self.p = E.make_pbs_params(**self.pd)
        self.slot = 1 << max(0, (self.d - 1).bit_length())
        if self.slot > self.p.N:
            raise ValueError("input_dim exceeds the polynomial size")
        if self.d * (1 << (2 * PACKED_IN_BITS - 2)) >= (1 << (PACKED_SCORE_BITS - 1)):
            raise ValueError("input_dim too large for the 17-bit score range")
        self.per = self.p.N // self.slot
        self.dev = E._dev(device)
        self.key_seed = seed_or_fresh(key_seed)
        self.noise_seed, self.enc_seed = seed_or_fresh(noise_seed), seed_or_fresh(enc_seed)
        self.ids = CiphertextIds(ct_start)
        self.scale = None
        self.S = None

    # ---- client
    def keygen(self) -> "PackedEncryptedCompare":
        self.S = E.secret_key(self.key_seed, 1, self.p.k * self.p.N, self.dev)
        return self

    def fit_scale(self, X: np.ndarray, clip_sigmas: float = 2.5) -> float:
        sd = float(np.asarray(X, dtype=np.float64).std())
        self.scale = clip_sigmas * sd / (1 << (PACKED_IN_BITS - 1)) if sd > 0 else 1.0
        return self.scale

    def quantize(self, X: np.ndarray) -> np.ndarray:
        if self.scale is None:
            raise RuntimeError("Quantizer not calibrated. Call fit_scale() first.")
        lo, hi = -(1 << (PACKED_IN_BITS - 1)), (1 << (PACKED_IN_BITS - 1)) - 1
        return np.clip(np.rint(np.asarray(X, dtype=np.float64) / self.scale), lo, hi).astype(np.int64)

    def dequantize(self, q_scores: np.ndarray) -> np.ndarray:
        return np.asarray(q_scores, dtype=np.float64) * (self.scale * self.scale)

    def encrypt_documents(self, Yq: np.ndarray, enc_seed: int | None = None, id_base: int | None = None) -> torch.Tensor:
        """int [B, d] -> GLWE [ceil(B/per), 2, N] (documents packed `per` to a ciphertext).  ``id_base=None`` takes
        fresh, never-reused ciphertext ids; ``enc_seed`` (public mask seed) defaults to the engine's."""
        self._need_keys()
        polys = torch.as_tensor(pack_documents(Yq, self.p.N, self.slot))
        base = self.ids.take(polys.shape[0]) if id_base is None else id_base
        return E.glwe_encrypt_vectors(self.p, self.S, polys, PACKED_OUT_SHIFT, self.enc_seed if enc_seed is None else enc_seed,
                                      base, noise_seed=self.noise_seed)

    def encrypt_query(self, xq: np.ndarray, enc_seed: int | None = None, id_base: int | None = None) -> torch.Tensor:
        """int [d] -> Fourier GGSW of Q(X) (the only thing the server needs for this query)."""
        self._need_keys()
        base = self.ids.take((self.p.k + 1) * self.p.l_pbs) if id_base is None else id_base
        ggsw = E.ggsw_encrypt_poly(self.p, self.S, torch.as_tensor(query_polynomial(xq, self.p.N)),
                                   self.enc_seed if enc_seed is None else enc_seed, base, noise_seed=self.noise_seed)
        return E.ggsw_to_fourier(self.p, ggsw)

    def decrypt(self, products: torch.Tensor, n_docs: int) -> np.ndarray:
        self._need_keys()
        v = E.glwe_decrypt_coeffs(self.p, self.S, products, 0, self.slot, self.per, PACKED_OUT_SHIFT).cpu().numpy()
        v = v.reshape(-1)[:n_docs] & ((1 << PACKED_SCORE_BITS) - 1)
        return np.where(v >= (1 << (PACKED_SCORE_BITS - 1)), v - (1 << PACKED_SCORE_BITS), v)

    # ---- server
    def scores(self, query_ggsw_fourier: torch.Tensor, docs_glwe: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """-> GLWE [G, 2, N]; coefficient slot*b of ciphertext g encrypts the score of document g*per + b."""
        return E.glwe_ggsw_dot(self.p, query_ggsw_fourier, docs_glwe, out)

    def scores_as_lwe(self, products: torch.Tensor) -> torch.Tensor:
        """Sample-extract every packed score into its own big-key LWE ciphertext [G*per, kN+2]."""
        return E.glwe_sample_extract(self.p, products, 0, self.slot, self.per)

    # ---- whole pipeline
    def compare_clear(self, q: np.ndarray, docs: np.ndarray) -> np.ndarray:
        return self.quantize(docs) @ self.quantize(q)

    def similarity(self, q: np.ndarray, docs: np.ndarray, enc_seed: int | None = None) -> np.ndarray:
        docs = np.atleast_2d(docs)
        gq = self.encrypt_query(self.quantize(q), enc_seed)       # fresh ciphertext ids on every call
        gd = self.encrypt_documents(self.quantize(docs), enc_seed)
        return self.dequantize(self.decrypt(self.scores(gq, gd), docs.shape[0]))

    def _need_keys(self):
        if self.S is None:
            raise RuntimeError("No secret key. Call keygen() first.")
