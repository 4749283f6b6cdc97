"""CPU ORACLE -- test infrastructure, NOT product code.

ctypes loader for ``liboracle.so`` (oracle/fhe_oracle.c) plus a numpy restatement of
the quantized clear circuit that the reference's CLI actually executes
(/root/reference/batch_operations.py:233,276 -> Concrete-ML ``LinearRegression.predict``,
un-vendored: concrete-ml==1.9.0, /root/reference/requirements.txt:5; behaviour restated
from SURVEY.md Appendix A.1/A.2).

Parity status: "parity unpinned" at the ciphertext level (no upstream goldens, Concrete
not installable here).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

KIND_SK, KIND_MASK, KIND_NOISE = 1, 2, 3
PUR_INPUT, PUR_KSK, PUR_BSK, PUR_BSK2 = 0, 1, 2, 3


class PBSParams(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("k", C.c_int32), ("N", C.c_int32),
        ("l_pbs", C.c_int32), ("beta_pbs", C.c_int32),
        ("l_ks", C.c_int32), ("beta_ks", C.c_int32), ("_pad", C.c_int32),
        ("sigma_lwe_abs", C.c_double), ("sigma_glwe_abs", C.c_double),
    ]


def build(force: bool = False) -> Path:
    so = _HERE / "liboracle.so"
    src = _HERE / "fhe_oracle.c"
    if force or not so.exists() or (src.exists() and so.stat().st_mtime < src.stat().st_mtime):
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        so = build()
        L = C.CDLL(str(so))
        u8p, u32p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)
        u64p, i64p, f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_int64), C.POINTER(C.c_double)
        pp = C.POINTER(PBSParams)
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_philox4x32_r.argtypes = [u32p, u32p, C.c_int, u32p]
        L.orc_rng_block.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, u32p]
        L.orc_det_log.argtypes = [C.c_double]; L.orc_det_log.restype = C.c_double
        L.orc_det_cos2pi_k53.argtypes = [C.c_uint64]; L.orc_det_cos2pi_k53.restype = C.c_double
        L.orc_gaussian.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, C.c_double]
        L.orc_gaussian.restype = C.c_int64
        L.orc_secret_key.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, u8p]
        L.orc_lwe_encrypt_batch.argtypes = [u8p, C.c_int32, C.c_int64, i64p, C.c_int64, C.c_int32,
                                            C.c_double, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, u64p]
        L.orc_lwe_phase_batch.argtypes = [u8p, C.c_int32, C.c_int64, u64p, C.c_int64, u64p]
        L.orc_lwe_decrypt_batch.argtypes = [u8p, C.c_int32, C.c_int64, u64p, C.c_int64, C.c_int32, i64p]
        L.orc_lincomb_batch.argtypes = [u64p, C.c_int64, C.c_int32, C.c_int32, C.c_int64, i64p,
                                        C.c_int32, i64p, C.c_int32, u64p]
        L.orc_ksk_gen.argtypes = [pp, u8p, u8p, C.c_uint64, u64p]
        L.orc_bsk_gen.argtypes = [pp, u8p, u8p, C.c_uint64, u64p]
        L.orc_bsk_to_fourier.argtypes = [pp, u64p, f64p]
        L.orc_keyswitch_batch.argtypes = [pp, u64p, u64p, C.c_int64, u64p]
        L.orc_ksk_to_32.argtypes = [pp, u64p, u32p]
        L.orc_keyswitch32_batch.argtypes = [pp, u32p, u64p, C.c_int64, u64p]
        L.orc_modswitch_batch.argtypes = [pp, u64p, C.c_int64, i32p]
        L.orc_pbs_batch.argtypes = [pp, f64p, u64p, C.c_int64, u64p, i32p, u64p]
        L.orc_bsk2_gen.argtypes = [pp, u8p, u8p, C.c_uint64, u64p]
        L.orc_bsk2_to_fourier.argtypes = [pp, u64p, f64p]
        L.orc_pbs_mb2_batch.argtypes = [pp, f64p, u64p, C.c_int64, u64p, i32p, u64p]
        L.orc_glwe_encrypt_rows.argtypes = [pp, u8p, i64p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_uint64,
                                            C.c_uint64, C.c_uint64, u64p]
        L.orc_glwe_external_product_batch.argtypes = [pp, f64p, u64p, C.c_int64, u64p]
        L.orc_glwe_sample_extract.argtypes = [pp, u64p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64, u64p]
        L.orc_negacyclic_mul_fft.argtypes = [C.c_int32, i64p, u64p, u64p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(C.POINTER(ct))


# ----------------------------------------------------------------------------- RNG
def philox(ctr, key, rounds: int = 10) -> np.ndarray:
    """Philox4x32-R: 10 rounds for secret streams, MASK_ROUNDS for the public mask stream."""
    c = np.asarray(ctr, dtype=np.uint32); k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_r(_p(c, C.c_uint32), _p(k, C.c_uint32), int(rounds), _p(out, C.c_uint32))
    return out


def gaussian(seed: int, domain: int, obj: int, blk: int, sigma_abs: float) -> int:
    return int(lib().orc_gaussian(seed, domain, obj, blk, sigma_abs))


# ----------------------------------------------------------------------------- LWE
def secret_key(key_seed: int, key_id: int, dim: int) -> np.ndarray:
    s = np.zeros(dim, dtype=np.uint8)
    lib().orc_secret_key(key_seed, key_id, dim, _p(s, C.c_uint8))
    return s


def lwe_encrypt(s, msgs, shift: int, sigma_abs: float, enc_seed: int, ct_base: int = 0,
                purpose: int = PUR_INPUT, stride: int | None = None, noise_seed: int | None = None) -> np.ndarray:
    """``enc_seed``: public mask seed; ``noise_seed``: the client's secret seed of the error terms.  The oracle is a
    checker, so ``noise_seed=None`` simply reuses ``enc_seed`` (the convention the committed goldens were made with);
    the product never does that."""
    noise_seed = enc_seed if noise_seed is None else noise_seed
    s = np.ascontiguousarray(s, dtype=np.uint8)
    m = np.ascontiguousarray(np.asarray(msgs).reshape(-1), dtype=np.int64)
    n = s.size
    stride = stride or n + 1
    out = np.zeros((m.size, stride), dtype=np.uint64)
    lib().orc_lwe_encrypt_batch(_p(s, C.c_uint8), n, stride, _p(m, C.c_int64), m.size, shift,
                                float(sigma_abs), enc_seed, noise_seed, ct_base, purpose, _p(out, C.c_uint64))
    return out


def lwe_phase(s, ct) -> np.ndarray:
    s = np.ascontiguousarray(s, dtype=np.uint8)
    ct = np.ascontiguousarray(ct, dtype=np.uint64)
    stride = ct.shape[-1]
    cnt = ct.size // stride
    ph = np.zeros(cnt, dtype=np.uint64)
    lib().orc_lwe_phase_batch(_p(s, C.c_uint8), s.size, stride, _p(ct, C.c_uint64), cnt, _p(ph, C.c_uint64))
    return ph.reshape(ct.shape[:-1])


def lwe_decrypt(s, ct, shift: int) -> np.ndarray:
    s = np.ascontiguousarray(s, dtype=np.uint8)
    ct = np.ascontiguousarray(ct, dtype=np.uint64)
    stride = ct.shape[-1]
    cnt = ct.size // stride
    out = np.zeros(cnt, dtype=np.int64)
    lib().orc_lwe_decrypt_batch(_p(s, C.c_uint8), s.size, stride, _p(ct, C.c_uint64), cnt, shift,
                                _p(out, C.c_int64))
    return out.reshape(ct.shape[:-1])


def lincomb(ct, W, n: int, bias=None, shift: int = 0) -> np.ndarray:
    """ct [B,d,stride] u64, W [M,d] i64 -> [B,M,stride] u64."""
    ct = np.ascontiguousarray(ct, dtype=np.uint64)
    W = np.ascontiguousarray(np.atleast_2d(W), dtype=np.int64)
    B, d, stride = ct.shape
    M = W.shape[0]
    out = np.zeros((B, M, stride), dtype=np.uint64)
    bp = None
    if bias is not None:
        bias = np.ascontiguousarray(bias, dtype=np.int64)
        bp = _p(bias, C.c_int64)
    lib().orc_lincomb_batch(_p(ct, C.c_uint64), B, d, n, stride, _p(W, C.c_int64), M, bp, shift,
                            _p(out, C.c_uint64))
    return out


# ----------------------------------------------------------------------------- KS / PBS
def make_params(n=742, k=1, N=2048, l_pbs=1, beta_pbs=23, l_ks=5, beta_ks=3,
                log2_sigma_lwe=-17.1, log2_sigma_glwe=-51.6) -> PBSParams:
    return PBSParams(n, k, N, l_pbs, beta_pbs, l_ks, beta_ks, 0,
                     2.0 ** (64 + log2_sigma_lwe), 2.0 ** (64 + log2_sigma_glwe))


def ksk_gen(p: PBSParams, S_big, s_small, evk_seed: int) -> np.ndarray:
    S_big = np.ascontiguousarray(S_big, dtype=np.uint8); s_small = np.ascontiguousarray(s_small, dtype=np.uint8)
    out = np.zeros((p.k * p.N, p.l_ks, p.n + 1), dtype=np.uint64)
    lib().orc_ksk_gen(C.byref(p), _p(S_big, C.c_uint8), _p(s_small, C.c_uint8), evk_seed, _p(out, C.c_uint64))
    return out


def bsk_gen(p: PBSParams, s_small, S_big, evk_seed: int) -> np.ndarray:
    S_big = np.ascontiguousarray(S_big, dtype=np.uint8); s_small = np.ascontiguousarray(s_small, dtype=np.uint8)
    out = np.zeros((p.n, p.k + 1, p.l_pbs, p.k + 1, p.N), dtype=np.uint64)
    lib().orc_bsk_gen(C.byref(p), _p(s_small, C.c_uint8), _p(S_big, C.c_uint8), evk_seed, _p(out, C.c_uint64))
    return out


def bsk_to_fourier(p: PBSParams, bsk) -> np.ndarray:
    bsk = np.ascontiguousarray(bsk, dtype=np.uint64)
    out = np.zeros((p.n, p.k + 1, p.l_pbs, p.k + 1, p.N // 2, 2), dtype=np.float64)
    lib().orc_bsk_to_fourier(C.byref(p), _p(bsk, C.c_uint64), _p(out, C.c_double))
    return out


def keyswitch(p: PBSParams, ksk, ct) -> np.ndarray:
    ct = np.ascontiguousarray(ct, dtype=np.uint64); ksk = np.ascontiguousarray(ksk, dtype=np.uint64)
    B = ct.shape[0]
    out = np.zeros((B, p.n + 1), dtype=np.uint64)
    lib().orc_keyswitch_batch(C.byref(p), _p(ksk, C.c_uint64), _p(ct, C.c_uint64), B, _p(out, C.c_uint64))
    return out


def ksk_to_32(p: PBSParams, ksk) -> np.ndarray:
    ksk = np.ascontiguousarray(ksk, dtype=np.uint64)
    out = np.zeros(ksk.shape, dtype=np.uint32)
    lib().orc_ksk_to_32(C.byref(p), _p(ksk, C.c_uint64), _p(out, C.c_uint32))
    return out


def keyswitch32(p: PBSParams, ksk32, ct) -> np.ndarray:
    ct = np.ascontiguousarray(ct, dtype=np.uint64); ksk32 = np.ascontiguousarray(ksk32, dtype=np.uint32)
    B = ct.shape[0]
    out = np.zeros((B, p.n + 1), dtype=np.uint64)
    lib().orc_keyswitch32_batch(C.byref(p), _p(ksk32, C.c_uint32), _p(ct, C.c_uint64), B, _p(out, C.c_uint64))
    return out


def modswitch(p: PBSParams, ct) -> np.ndarray:
    ct = np.ascontiguousarray(ct, dtype=np.uint64)
    B = ct.shape[0]
    out = np.zeros((B, p.n + 1), dtype=np.int32)
    lib().orc_modswitch_batch(C.byref(p), _p(ct, C.c_uint64), B, _p(out, C.c_int32))
    return out


def pbs(p: PBSParams, bskf, ct, luts, lut_index=None) -> np.ndarray:
    ct = np.ascontiguousarray(ct, dtype=np.uint64); bskf = np.ascontiguousarray(bskf, dtype=np.float64)
    luts = np.ascontiguousarray(np.atleast_2d(luts), dtype=np.uint64)
    B = ct.shape[0]
    out = np.zeros((B, p.k * p.N + 1), dtype=np.uint64)
    li = None
    if lut_index is not None:
        lut_index = np.ascontiguousarray(lut_index, dtype=np.int32)
        li = _p(lut_index, C.c_int32)
    lib().orc_pbs_batch(C.byref(p), _p(bskf, C.c_double), _p(ct, C.c_uint64), B, _p(luts, C.c_uint64), li,
                        _p(out, C.c_uint64))
    return out


def bsk2_gen(p: PBSParams, s_small, S_big, evk_seed: int) -> np.ndarray:
    """Multi-bit (grouping factor 2) bootstrapping key: [n/2][3][k+1][l][k+1][N] u64."""
    S_big = np.ascontiguousarray(S_big, dtype=np.uint8); s_small = np.ascontiguousarray(s_small, dtype=np.uint8)
    out = np.zeros((p.n // 2, 3, p.k + 1, p.l_pbs, p.k + 1, p.N), dtype=np.uint64)
    lib().orc_bsk2_gen(C.byref(p), _p(s_small, C.c_uint8), _p(S_big, C.c_uint8), evk_seed, _p(out, C.c_uint64))
    return out


def bsk2_to_fourier(p: PBSParams, bsk2) -> np.ndarray:
    bsk2 = np.ascontiguousarray(bsk2, dtype=np.uint64)
    out = np.zeros((p.n // 2, 3, p.k + 1, p.l_pbs, p.k + 1, p.N // 2, 2), dtype=np.float64)
    lib().orc_bsk2_to_fourier(C.byref(p), _p(bsk2, C.c_uint64), _p(out, C.c_double))
    return out


def pbs_mb2(p: PBSParams, bskf2, ct, luts, lut_index=None) -> np.ndarray:
    ct = np.ascontiguousarray(ct, dtype=np.uint64); bskf2 = np.ascontiguousarray(bskf2, dtype=np.float64)
    luts = np.ascontiguousarray(np.atleast_2d(luts), dtype=np.uint64)
    B = ct.shape[0]
    out = np.zeros((B, p.k * p.N + 1), dtype=np.uint64)
    li = None
    if lut_index is not None:
        lut_index = np.ascontiguousarray(lut_index, dtype=np.int32)
        li = _p(lut_index, C.c_int32)
    lib().orc_pbs_mb2_batch(C.byref(p), _p(bskf2, C.c_double), _p(ct, C.c_uint64), B, _p(luts, C.c_uint64), li,
                            _p(out, C.c_uint64))
    return out


# ----------------------------------------------------------------------------- encrypted x encrypted comparison
# both vectors encrypted; replaces the clear product of batch_operations.py:226,273.
# x*y = floor((x+y)^2/4) - floor((x-y)^2/4), one table lookup per term.
def pair_addsub(q, y, words: int, offset: int) -> np.ndarray:
    """q [d,stride], y [B,d,stride] -> [B,d,2,words] = (q+y+offset, q-y+offset), offset on the body."""
    q = np.asarray(q, dtype=np.uint64)[None, :, :words]
    y = np.asarray(y, dtype=np.uint64)[:, :, :words]
    out = np.empty(y.shape[:2] + (2, words), dtype=np.uint64)
    out[:, :, 0] = q + y
    out[:, :, 1] = q - y
    out[:, :, :, words - 1] += np.uint64(offset & 0xFFFFFFFFFFFFFFFF)
    return out


def pair_diff_sum(sq) -> np.ndarray:
    """sq [B,d,2,words] -> [B,words] = sum_j (sq[:,j,0] - sq[:,j,1]) (wrapping u64)."""
    sq = np.asarray(sq, dtype=np.uint64)
    return (sq[:, :, 0] - sq[:, :, 1]).sum(axis=1, dtype=np.uint64)


def quarter_square_table(p_bits: int) -> np.ndarray:
    """table[w] = floor((w - 2^(p_bits-1))^2 / 4) for the offset-binary message w."""
    w = np.arange(1 << p_bits, dtype=np.int64) - (1 << (p_bits - 1))
    return (w * w) // 4


def encrypted_product_scores(p: PBSParams, bskf, ct_q, ct_docs, p_bits: int, out_shift: int,
                             multibit: bool = False) -> np.ndarray:
    """ct_q [d,stride], ct_docs [B,d,stride] under the small key -> [B, kN+1] under the big key:
    LWE encryptions of sum_j x_j*y_j at 2^out_shift."""
    words = p.n + 1
    B, d = np.asarray(ct_docs).shape[:2]
    pairs = pair_addsub(ct_q, ct_docs, words, 1 << (63 - 1))  # + 2^(p_bits-1) messages = half of the unsigned range
    lut = make_lut_poly(quarter_square_table(p_bits), p_bits, p.N, out_shift)
    fn = pbs_mb2 if multibit else pbs
    sq = fn(p, bskf, pairs.reshape(-1, words), lut)
    return pair_diff_sum(sq.reshape(B, d, 2, -1))


def square_table(p_bits: int) -> np.ndarray:
    w = np.arange(1 << p_bits, dtype=np.int64) - (1 << (p_bits - 1))
    return w * w


def encrypted_product_scores_norms(p: PBSParams, bskf, ct_q, ct_docs, norm_q, norm_docs, p_bits: int,
                                   out_shift: int, multibit: bool = False) -> np.ndarray:
    """One bootstrap per dimension: sum_j (x_j+y_j)^2 - |x|^2 - |y|^2 = 2 sum_j x_j*y_j, squares and norms at
    2^(out_shift-1) so that the result encodes sum_j x_j*y_j at 2^out_shift.  norm_q [>=kN+1], norm_docs [B, >=kN+1]."""
    words = p.n + 1
    big = p.k * p.N + 1
    ct_docs = np.asarray(ct_docs, dtype=np.uint64)
    B, d = ct_docs.shape[:2]
    sums = np.asarray(ct_q, dtype=np.uint64)[None, :, :words] + ct_docs[:, :, :words]
    sums[:, :, words - 1] += np.uint64(1 << 62)
    lut = make_lut_poly(square_table(p_bits), p_bits, p.N, out_shift - 1)
    fn = pbs_mb2 if multibit else pbs
    sq = fn(p, bskf, sums.reshape(-1, words), lut).reshape(B, d, big)
    return (sq.sum(axis=1, dtype=np.uint64) - np.asarray(norm_q, dtype=np.uint64).reshape(-1)[None, :big]
            - np.asarray(norm_docs, dtype=np.uint64)[:, :big])


def encrypted_ge(p: PBSParams, ksk32, bskf, scores, T: int, score_bits: int, out_shift: int, bit_shift: int,
                 multibit: bool = False) -> np.ndarray:
    """SURVEY.md 8f N3 (replaces the clear test of batch_operations.py:278): scores [B, >=kN+1] big-key
    LWE of a signed score_bits-bit value at 2^out_shift -> [B, kN+1] LWE of (score >= T) * 2^bit_shift,
    by LSB-first bit extraction (one 32-bit keyswitch + sign bootstrap per bit)."""
    words = p.k * p.N + 1
    M64 = 0xFFFFFFFFFFFFFFFF
    acc = np.array(np.asarray(scores, dtype=np.uint64)[:, :words])
    acc[:, -1] -= np.uint64((int(T) << out_shift) & M64)
    fn = pbs_mb2 if multibit else pbs
    for i in range(score_bits):
        last = i == score_bits - 1
        tmp = acc << np.uint64(score_bits - 1 - i)
        tmp[:, -1] += np.uint64(1 << 62)
        c = (1 << (bit_shift - 1)) if last else (-(1 << (out_shift - 1 + i))) & M64
        pb = fn(p, bskf, keyswitch32(p, ksk32, tmp), np.full(p.N, c, dtype=np.uint64))
        if last:
            pb[:, -1] += np.uint64(1 << (bit_shift - 1))
            return pb
        acc -= pb
        acc[:, -1] -= np.uint64(1 << (out_shift - 1 + i))


# ----------------------------------------------------------------------------- packed encrypted inner products
# Leveled variant of the both-encrypted comparison: P = N/slot documents per GLWE ciphertext
# (document b of a group occupies coefficients slot*b .. slot*b+d-1), the query as a GGSW of
# Q(X) = sum_j x_j X^(-j); coefficient slot*b of GGSW(Q) [.] GLWE(D) is sum_j x_j * y_{b,j}.
def pack_documents(Yq, N: int, slot: int) -> np.ndarray:
    Yq = np.asarray(Yq, dtype=np.int64)
    B, d = Yq.shape
    per = N // slot
    G = (B + per - 1) // per
    out = np.zeros((G * per, slot), dtype=np.int64)
    out[:B, :d] = Yq
    return out.reshape(G, N)


def query_polynomial(xq, N: int) -> np.ndarray:
    xq = np.asarray(xq, dtype=np.int64)
    Q = np.zeros(N, dtype=np.int64)
    Q[0] = xq[0]
    Q[N - np.arange(1, xq.size)] = -xq[1:]
    return Q


def glwe_encrypt_rows(p: PBSParams, S_big, msgs, mode: int, shift: int, seed: int, id_base: int = 0,
                      noise_seed: int | None = None) -> np.ndarray:
    """mode 0: msgs [rows][N] -> GLWE(msg << shift) [rows][k+1][N]; mode 1: msgs [N] -> GGSW rows [(k+1)*l][k+1][N]."""
    S_big = np.ascontiguousarray(S_big, dtype=np.uint8)
    msgs = np.ascontiguousarray(msgs, dtype=np.int64)
    rows = msgs.shape[0] if mode == 0 else (p.k + 1) * p.l_pbs
    out = np.zeros((rows, p.k + 1, p.N), dtype=np.uint64)
    lib().orc_glwe_encrypt_rows(C.byref(p), _p(S_big, C.c_uint8), _p(msgs, C.c_int64), rows, p.N if mode == 0 else 0,
                                mode, shift, seed, seed if noise_seed is None else noise_seed, id_base, _p(out, C.c_uint64))
    return out


def ggsw_to_fourier(p: PBSParams, ggsw) -> np.ndarray:
    p1 = PBSParams(1, p.k, p.N, p.l_pbs, p.beta_pbs, p.l_ks, p.beta_ks, 0, p.sigma_lwe_abs, p.sigma_glwe_abs)
    return bsk_to_fourier(p1, np.ascontiguousarray(ggsw, dtype=np.uint64))


def glwe_external_product(p: PBSParams, ggswf, glwe) -> np.ndarray:
    glwe = np.ascontiguousarray(glwe, dtype=np.uint64)
    ggswf = np.ascontiguousarray(ggswf, dtype=np.float64)
    out = np.zeros_like(glwe)
    lib().orc_glwe_external_product_batch(C.byref(p), _p(ggswf, C.c_double), _p(glwe, C.c_uint64), glwe.shape[0],
                                          _p(out, C.c_uint64))
    return out


def glwe_sample_extract(p: PBSParams, glwe, first: int, step: int, count: int, out_stride: int | None = None) -> np.ndarray:
    glwe = np.ascontiguousarray(glwe, dtype=np.uint64)
    out_stride = out_stride or p.k * p.N + 1
    out = np.zeros((glwe.shape[0] * count, out_stride), dtype=np.uint64)
    lib().orc_glwe_sample_extract(C.byref(p), _p(glwe, C.c_uint64), glwe.shape[0], first, step, count, out_stride,
                                  _p(out, C.c_uint64))
    return out


def negacyclic_mul_fft(a_small, b_torus) -> np.ndarray:
    a = np.ascontiguousarray(a_small, dtype=np.int64); b = np.ascontiguousarray(b_torus, dtype=np.uint64)
    out = np.zeros(a.size, dtype=np.uint64)
    lib().orc_negacyclic_mul_fft(a.size, _p(a, C.c_int64), _p(b, C.c_uint64), _p(out, C.c_uint64))
    return out


def negacyclic_mul_naive(a_small, b_torus) -> np.ndarray:
    """O(N^2) schoolbook product mod X^N+1, wrapping u64 (pure-python ints, small N only)."""
    a = [int(x) for x in np.asarray(a_small, dtype=np.int64)]
    b = [int(x) for x in np.asarray(b_torus, dtype=np.uint64)]
    N = len(a)
    out = [0] * N
    for i in range(N):
        if a[i] == 0:
            continue
        for j in range(N):
            k = i + j
            if k < N:
                out[k] += a[i] * b[j]
            else:
                out[k - N] -= a[i] * b[j]
    return np.array([x % (1 << 64) for x in out], dtype=np.uint64)


def make_lut_poly(table, p_bits: int, N: int, delta_out_log2: int) -> np.ndarray:
    """Accumulator polynomial for a p-bit message (+1 padding bit): box m holds
    table[m] << delta_out_log2, whole polynomial multiplied by X^(-box/2)."""
    box = N >> p_bits
    t = np.asarray(table, dtype=np.int64)
    assert t.size == (1 << p_bits)
    p0 = (np.repeat(t, box).astype(np.int64).astype(np.uint64)) << np.uint64(delta_out_log2)
    half = box // 2
    out = np.empty(N, dtype=np.uint64)
    out[: N - half] = p0[half:]
    out[N - half:] = (np.uint64(0) - p0[:half])
    return out


# ----------------------------------------------------------------------------- clear circuit
STABILITY_CONST = 1e-6


def uniform_quantizer_params(values: np.ndarray, n_bits: int, is_signed: bool):
    """Concrete-ML UniformQuantizer statistics."""
    v = np.asarray(values, dtype=np.float64)
    offset = 2 ** (n_bits - 1) if is_signed else 0
    rmin, rmax = float(v.min()), float(v.max())
    if abs(rmax - rmin) < STABILITY_CONST:
        if abs(rmax) < STABILITY_CONST:
            return 1.0, 0, offset
        return rmax, 0, offset
    scale = (rmax - rmin) / (2 ** n_bits - 1)
    zp = int(np.round((rmax * (-offset) - rmin * (2 ** n_bits - 1 - offset)) / (rmax - rmin)))
    return scale, zp, offset


def quantize(values, scale, zp, offset, n_bits) -> np.ndarray:
    q = np.rint(np.asarray(values, dtype=np.float64) / scale + zp)
    return np.clip(q, -offset, 2 ** n_bits - 1 - offset).astype(np.int64)


def clear_circuit(q_X: np.ndarray, q_W: np.ndarray, zp_W: int, q_bias: int) -> np.ndarray:
    """q_y = q_X @ q_W - zp_W * sum_j q_X[j] + q_bias  (int64)."""
    q_X = np.asarray(q_X, dtype=np.int64)
    return q_X @ np.asarray(q_W, dtype=np.int64) - int(zp_W) * q_X.sum(axis=1) + int(q_bias)


def num_threads() -> int:
    return int(lib().orc_num_threads())


def cpu_count() -> int:
    return os.cpu_count() or 1
