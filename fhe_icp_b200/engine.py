"""Torch-tensor front end over the C-ABI: torch owns device memory and streams, every
operation is a kernel of libfhe_b200.so (see include/fhe_b200.h).  u64 torus words are
carried in ``torch.int64`` tensors (same bits)."""
from __future__ import annotations

import ctypes as C
import secrets

import numpy as np
import torch

from . import _native as N


def _dev(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("fhe_icp_b200 requires a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(device)


def _stream(dev: torch.device):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t: torch.Tensor):
    return C.c_void_p(t.data_ptr())


def _ctx(dev: torch.device) -> N.Context:
    return N.context(dev.index if dev.index is not None else torch.cuda.current_device())


def even_stride(n: int) -> int:
    """Smallest even row length holding n mask words + the body."""
    return (n + 2) & ~1


def fresh_seed() -> int:
    """A 64-bit seed from the OS CSPRNG (secret keys, noise, evaluation keys, per-process ciphertext-id nonces)."""
    return secrets.randbits(64)


def _noise(noise_seed) -> int:
    # the error terms must come from a SECRET seed, never from the public mask seed: an explicit value is for
    # reproducible tests / benchmarks only
    return fresh_seed() if noise_seed is None else int(noise_seed)


def to_u64_numpy(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().view(np.uint64)


def from_u64_numpy(a: np.ndarray, device) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(device)


# ------------------------------------------------------------------------------- client side
def secret_key(key_seed: int, key_id: int, dim: int, device=None) -> torch.Tensor:
    dev = _dev(device)
    key = torch.empty(dim, dtype=torch.uint8, device=dev)
    N.check(N.lib().fhe_b200_secret_key(_ctx(dev).handle, key_seed, key_id, dim, _ptr(key), _stream(dev)))
    return key


def lwe_encrypt(key: torch.Tensor, msgs: torch.Tensor, shift: int, sigma_abs: float, enc_seed: int,
                ct_base: int = 0, purpose: int = N.PUR_INPUT, stride: int | None = None,
                noise_seed: int | None = None) -> torch.Tensor:
    """``enc_seed`` is the PUBLIC mask seed; ``noise_seed`` the client's SECRET seed of the error terms (default: a
    fresh one from the OS CSPRNG -- pass a value only to reproduce ciphertexts in tests).  Ciphertext ids
    ``ct_base + i`` must never repeat under the same seeds."""
    dev = key.device
    n = key.numel()
    stride = stride or even_stride(n)
    m = msgs.to(device=dev, dtype=torch.int64).contiguous()
    ct = torch.empty(tuple(m.shape) + (stride,), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_lwe_encrypt(_ctx(dev).handle, _ptr(key), n, stride, _ptr(m), m.numel(), shift,
                                         float(sigma_abs), enc_seed, _noise(noise_seed), ct_base, purpose, _ptr(ct),
                                         _stream(dev)))
    return ct


def lwe_phase(key: torch.Tensor, ct: torch.Tensor) -> torch.Tensor:
    dev = key.device
    n, stride = key.numel(), ct.shape[-1]
    ct = ct.contiguous()
    out = torch.empty(ct.shape[:-1], dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_lwe_phase(_ctx(dev).handle, _ptr(key), n, stride, _ptr(ct), out.numel(), _ptr(out),
                                       _stream(dev)))
    return out


def lwe_decrypt(key: torch.Tensor, ct: torch.Tensor, shift: int) -> torch.Tensor:
    dev = key.device
    n, stride = key.numel(), ct.shape[-1]
    ct = ct.contiguous()
    out = torch.empty(ct.shape[:-1], dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_lwe_decrypt(_ctx(dev).handle, _ptr(key), n, stride, _ptr(ct), out.numel(), shift,
                                         _ptr(out), _stream(dev)))
    return out


# ------------------------------------------------------------------------------- server side
def lincomb(ct: torch.Tensor, W: torch.Tensor, n: int, bias=None, shift: int = 0,
            out: torch.Tensor | None = None) -> torch.Tensor:
    """ct [B,d,stride], W [M,d] (M in {1,2}) -> [B,M,stride]."""
    dev = ct.device
    B, d, stride = ct.shape
    W = W.to(device=dev, dtype=torch.int64).reshape(-1, d).contiguous()
    M = W.shape[0]
    if out is None:
        out = torch.empty((B, M, stride), dtype=torch.int64, device=dev)
    hb = None
    if bias is not None:
        hb = (C.c_int64 * M)(*[int(x) for x in bias])
    N.check(N.lib().fhe_b200_lincomb(_ctx(dev).handle, _ptr(ct.contiguous()), B, d, n, stride, _ptr(W), M, hb,
                                     shift, _ptr(out), _stream(dev)))
    return out


# ------------------------------------------------------------------------------- seeded ciphertexts
def lwe_encrypt_seeded(key: torch.Tensor, msgs: torch.Tensor, shift: int, sigma_abs: float, enc_seed: int,
                       ct_base: int = 0, purpose: int = N.PUR_INPUT, noise_seed: int | None = None) -> torch.Tensor:
    """Bodies only (8 bytes per ciphertext); masks are regenerated from (enc_seed, ct_base + index)."""
    dev = key.device
    m = msgs.to(device=dev, dtype=torch.int64).contiguous()
    bodies = torch.empty(m.shape, dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_lwe_encrypt_seeded(_ctx(dev).handle, _ptr(key), key.numel(), _ptr(m), m.numel(), shift,
                                                float(sigma_abs), enc_seed, _noise(noise_seed), ct_base, purpose,
                                                _ptr(bodies), _stream(dev)))
    return bodies


def lwe_expand_seeded(bodies: torch.Tensor, n: int, enc_seed: int, ct_base: int = 0, purpose: int = N.PUR_INPUT,
                      stride: int | None = None) -> torch.Tensor:
    dev = bodies.device
    stride = stride or even_stride(n)
    b = bodies.contiguous()
    ct = torch.empty(tuple(b.shape) + (stride,), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_lwe_expand_seeded(_ctx(dev).handle, _ptr(b), b.numel(), n, stride, enc_seed, ct_base,
                                               purpose, _ptr(ct), _stream(dev)))
    return ct


def lincomb_seeded(bodies: torch.Tensor, W: torch.Tensor, n: int, enc_seed: int, ct_base: int = 0,
                   purpose: int = N.PUR_INPUT, bias=None, shift: int = 0, stride: int | None = None) -> torch.Tensor:
    """bodies [B,d], W [M,d] -> [B,M,stride]; identical to lincomb(lwe_expand_seeded(bodies), W)."""
    dev = bodies.device
    B, d = bodies.shape
    stride = stride or even_stride(n)
    W = W.to(device=dev, dtype=torch.int64).reshape(-1, d).contiguous()
    M = W.shape[0]
    out = torch.empty((B, M, stride), dtype=torch.int64, device=dev)
    hb = (C.c_int64 * M)(*[int(x) for x in bias]) if bias is not None else None
    N.check(N.lib().fhe_b200_lincomb_seeded(_ctx(dev).handle, _ptr(bodies.contiguous()), B, d, n, stride, enc_seed,
                                            ct_base, purpose, _ptr(W), M, hb, shift, _ptr(out), _stream(dev)))
    return out


def accumulate(acc: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    dev = acc.device
    assert acc.is_contiguous() and x.is_contiguous() and acc.numel() == x.numel()
    N.check(N.lib().fhe_b200_accumulate(_ctx(dev).handle, _ptr(acc), _ptr(x), acc.numel(), _stream(dev)))
    return acc


def pair_addsub(q: torch.Tensor, y: torch.Tensor, words: int, offset: int) -> torch.Tensor:
    """q [d,stride], y [B,d,stride] (ciphertext = first `words` words of a row) ->
    [B,d,2,words] = (q+y+offset, q-y+offset), offset on the body."""
    dev = y.device
    B, d, stride = y.shape
    assert q.shape == (d, stride) and q.is_contiguous() and y.is_contiguous() and words <= stride
    out = torch.empty((B, d, 2, words), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_lwe_pair_addsub(_ctx(dev).handle, _ptr(q), _ptr(y), B, d, words, stride,
                                             offset & 0xFFFFFFFFFFFFFFFF, _ptr(out), _stream(dev)))
    return out


def pair_diff_sum(sq: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """sq [B,d,2,words] -> [B,even_stride] = sum_j (sq[:,j,0] - sq[:,j,1]) (wrapping), rows zero-padded."""
    dev = sq.device
    B, d, two, words = sq.shape
    assert two == 2 and sq.is_contiguous()
    if out is None:
        out = torch.empty((B, even_stride(words - 1)), dtype=torch.int64, device=dev)
    assert out.is_contiguous() and out.shape[0] == B and out.shape[1] >= words
    N.check(N.lib().fhe_b200_lwe_pair_diff_sum(_ctx(dev).handle, _ptr(sq), B, d, words, out.shape[1], _ptr(out),
                                               _stream(dev)))
    return out


def pair_add(q: torch.Tensor, y: torch.Tensor, words: int, offset: int) -> torch.Tensor:
    """q [d,stride], y [B,d,stride] -> [B,d,words] = q + y + offset (offset on the body)."""
    dev = y.device
    B, d, stride = y.shape
    assert q.shape == (d, stride) and q.is_contiguous() and y.is_contiguous() and words <= stride
    out = torch.empty((B, d, words), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_lwe_pair_add(_ctx(dev).handle, _ptr(q), _ptr(y), B, d, words, stride,
                                          offset & 0xFFFFFFFFFFFFFFFF, _ptr(out), _stream(dev)))
    return out


def square_sum(sq: torch.Tensor, norm_q: torch.Tensor, norm_y: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """sq [B,d,words], norm_q [>=words], norm_y [B,norm_stride] -> [B,even_stride] = sum_j sq[:,j] - norm_q - norm_y."""
    dev = sq.device
    B, d, words = sq.shape
    assert sq.is_contiguous() and norm_q.is_contiguous() and norm_y.is_contiguous()
    assert norm_q.numel() >= words and norm_y.shape[0] == B and norm_y.shape[1] >= words
    if out is None:
        out = torch.empty((B, even_stride(words - 1)), dtype=torch.int64, device=dev)
    assert out.is_contiguous() and out.shape[0] == B and out.shape[1] >= words
    N.check(N.lib().fhe_b200_lwe_square_sum(_ctx(dev).handle, _ptr(sq), B, d, words, _ptr(norm_q), _ptr(norm_y),
                                            norm_y.shape[1], out.shape[1], _ptr(out), _stream(dev)))
    return out


def shl_add(ct: torch.Tensor, words: int, shift: int, offset: int, out_stride: int | None = None) -> torch.Tensor:
    """ct [count, stride] -> [count, out_stride or words] = (row << shift), offset added to the body,
    words beyond the ciphertext zeroed."""
    dev = ct.device
    count, stride = ct.shape
    out_stride = out_stride or words
    assert ct.is_contiguous() and words <= stride and out_stride >= words
    out = torch.empty((count, out_stride), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_lwe_shl_add(_ctx(dev).handle, _ptr(ct), stride, count, words, shift,
                                         offset & 0xFFFFFFFFFFFFFFFF, _ptr(out), out_stride, _stream(dev)))
    return out


def sub_plain(acc: torch.Tensor, x: torch.Tensor, plain: int) -> torch.Tensor:
    """acc [count, stride] -= x [count, words] (first `words` words of each row), body -= plain."""
    dev = acc.device
    count, words = x.shape
    assert acc.is_contiguous() and x.is_contiguous() and acc.shape[0] == count and acc.shape[1] >= words
    N.check(N.lib().fhe_b200_lwe_sub_plain(_ctx(dev).handle, _ptr(acc), acc.shape[1], _ptr(x), count, words,
                                           plain & 0xFFFFFFFFFFFFFFFF, _stream(dev)))
    return acc


# ------------------------------------------------------------------------------- KS / PBS
def make_pbs_params(n=742, k=1, N_poly=2048, l_pbs=1, beta_pbs=23, l_ks=5, beta_ks=3,
                    log2_sigma_lwe=-17.1, log2_sigma_glwe=-51.6) -> N.PBSParams:
    """Default = the 4-bit (message 2 + carry 2) KS->PBS set stated in DESIGN.md."""
    return N.PBSParams(n, k, N_poly, l_pbs, beta_pbs, l_ks, beta_ks, 0,
                       2.0 ** (64 + log2_sigma_lwe), 2.0 ** (64 + log2_sigma_glwe))


def ksk_gen(p: N.PBSParams, S_big: torch.Tensor, s_small: torch.Tensor, evk_seed: int) -> torch.Tensor:
    dev = S_big.device
    ksk = torch.empty((p.k * p.N, p.l_ks, p.n + 1), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_ksk_gen(_ctx(dev).handle, C.byref(p), _ptr(S_big), _ptr(s_small), evk_seed,
                                     _ptr(ksk), _stream(dev)))
    return ksk


def bsk_gen(p: N.PBSParams, s_small: torch.Tensor, S_big: torch.Tensor, evk_seed: int) -> torch.Tensor:
    dev = S_big.device
    bsk = torch.empty((p.n, p.k + 1, p.l_pbs, p.k + 1, p.N), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_bsk_gen(_ctx(dev).handle, C.byref(p), _ptr(s_small), _ptr(S_big), evk_seed,
                                     _ptr(bsk), _stream(dev)))
    return bsk


def bsk_to_fourier(p: N.PBSParams, bsk: torch.Tensor) -> torch.Tensor:
    dev = bsk.device
    bskf = torch.empty((p.n, p.k + 1, p.l_pbs, p.k + 1, p.N // 2, 2), dtype=torch.float64, device=dev)
    N.check(N.lib().fhe_b200_bsk_to_fourier(_ctx(dev).handle, C.byref(p), _ptr(bsk.contiguous()), _ptr(bskf),
                                            _stream(dev)))
    return bskf


def keyswitch(p: N.PBSParams, ksk: torch.Tensor, ct: torch.Tensor) -> torch.Tensor:
    dev = ct.device
    ct = ct.contiguous()
    B = ct.shape[0]
    assert ct.shape[1] == p.k * p.N + 1
    out = torch.empty((B, p.n + 1), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_keyswitch(_ctx(dev).handle, C.byref(p), _ptr(ksk), _ptr(ct), B, _ptr(out),
                                       _stream(dev)))
    return out


def ksk_to_32(p: N.PBSParams, ksk: torch.Tensor) -> torch.Tensor:
    """Round the keyswitching key to its top 32 torus bits (int32 view of u32)."""
    ksk32 = torch.empty(ksk.shape, dtype=torch.int32, device=ksk.device)
    N.check(N.lib().fhe_b200_ksk_to_32(_ctx(ksk.device).handle, C.byref(p), _ptr(ksk), _ptr(ksk32), _stream(ksk.device)))
    return ksk32


def keyswitch32(p: N.PBSParams, ksk32: torch.Tensor, ct: torch.Tensor) -> torch.Tensor:
    """32-bit keyswitch: same interface as :func:`keyswitch`, ~3x fewer integer operations."""
    dev = ct.device
    ct = ct.contiguous()
    B = ct.shape[0]
    assert ct.shape[1] == p.k * p.N + 1
    scratch = torch.empty((B, p.n + 1), dtype=torch.int32, device=dev)
    out = torch.empty((B, p.n + 1), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_keyswitch32(_ctx(dev).handle, C.byref(p), _ptr(ksk32), _ptr(ct), B, _ptr(scratch),
                                         _ptr(out), _stream(dev)))
    return out


def ksk_to_mma(p: N.PBSParams, ksk32: torch.Tensor) -> torch.Tensor:
    """Lay the 32-bit keyswitching key out as the byte blocks the tensor-core keyswitch reads (once per key)."""
    nbytes = int(N.lib().fhe_b200_ksk_mma_bytes(C.byref(p)))
    if nbytes == 0:
        raise ValueError("parameter set outside the tensor-core keyswitch")
    tiles = torch.empty(nbytes, dtype=torch.uint8, device=ksk32.device)
    N.check(N.lib().fhe_b200_ksk_to_mma(_ctx(ksk32.device).handle, C.byref(p), _ptr(ksk32), _ptr(tiles), _stream(ksk32.device)))
    return tiles


def keyswitch_mma(p: N.PBSParams, key_mma: torch.Tensor, ct: torch.Tensor, work: torch.Tensor | None = None) -> torch.Tensor:
    """32-bit keyswitch on the tensor cores (tcgen05 int8 contraction); bit-identical to :func:`keyswitch32`."""
    dev = ct.device
    ct = ct.contiguous()
    B = ct.shape[0]
    assert ct.shape[1] == p.k * p.N + 1
    out = torch.empty((B, p.n + 1), dtype=torch.int64, device=dev)
    if B == 0:
        return out
    need = int(N.lib().fhe_b200_keyswitch_mma_workspace_bytes(C.byref(p), B))
    if work is None or work.numel() < need:
        work = torch.empty(need, dtype=torch.int8, device=dev)
    N.check(N.lib().fhe_b200_keyswitch_mma(_ctx(dev).handle, C.byref(p), _ptr(key_mma), _ptr(ct), B, _ptr(work), _ptr(out),
                                           _stream(dev)))
    return out


def pbs(p: N.PBSParams, bskf: torch.Tensor, ct: torch.Tensor, luts: torch.Tensor,
        lut_index: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    dev = ct.device
    ct = ct.contiguous()
    B = ct.shape[0]
    assert ct.shape[1] == p.n + 1
    luts = luts.to(device=dev, dtype=torch.int64).reshape(-1, p.N).contiguous()
    if out is None:
        out = torch.empty((B, p.k * p.N + 1), dtype=torch.int64, device=dev)
    li = None
    if lut_index is not None:
        lut_index = lut_index.to(device=dev, dtype=torch.int32).contiguous()
        li = _ptr(lut_index)
    N.check(N.lib().fhe_b200_pbs(_ctx(dev).handle, C.byref(p), _ptr(bskf), _ptr(ct), B, _ptr(luts), li, _ptr(out),
                                 _stream(dev)))
    return out


# ---- multi-bit blind rotation (two key bits per CMux)
def bsk2_gen(p: N.PBSParams, s_small: torch.Tensor, S_big: torch.Tensor, evk_seed: int) -> torch.Tensor:
    dev = S_big.device
    bsk2 = torch.empty((p.n // 2, 3, p.k + 1, p.l_pbs, p.k + 1, p.N), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_bsk2_gen(_ctx(dev).handle, C.byref(p), _ptr(s_small), _ptr(S_big), evk_seed,
                                      _ptr(bsk2), _stream(dev)))
    return bsk2


def bsk2_to_fourier(p: N.PBSParams, bsk2: torch.Tensor) -> torch.Tensor:
    """-> [n/2][32 frequency blocks][3][k+1][l][k+1][32][2] f64 (the layout the kernel streams)."""
    dev = bsk2.device
    bskf2 = torch.empty((p.n // 2, 32, 3, p.k + 1, p.l_pbs, p.k + 1, 32, 2), dtype=torch.float64, device=dev)
    N.check(N.lib().fhe_b200_bsk2_to_fourier(_ctx(dev).handle, C.byref(p), _ptr(bsk2.contiguous()), _ptr(bskf2),
                                             _stream(dev)))
    return bskf2


def pbs_mb2(p: N.PBSParams, bskf2: torch.Tensor, ct: torch.Tensor, luts: torch.Tensor,
            lut_index: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    dev = ct.device
    ct = ct.contiguous()
    B = ct.shape[0]
    assert ct.shape[1] == p.n + 1
    luts = luts.to(device=dev, dtype=torch.int64).reshape(-1, p.N).contiguous()
    if out is None:
        out = torch.empty((B, p.k * p.N + 1), dtype=torch.int64, device=dev)
    li = None
    if lut_index is not None:
        lut_index = lut_index.to(device=dev, dtype=torch.int32).contiguous()
        li = _ptr(lut_index)
    N.check(N.lib().fhe_b200_pbs_mb2(_ctx(dev).handle, C.byref(p), _ptr(bskf2), _ptr(ct), B, _ptr(luts), li, _ptr(out),
                                     _stream(dev)))
    return out


def pbs_mb2_wide(p: N.PBSParams, bskf2: torch.Tensor, ct: torch.Tensor, luts: torch.Tensor,
                 lut_index: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """:func:`pbs_mb2` through the four-warps-per-polynomial latency kernel (one ciphertext per CTA, pbs_wide.cu)
    whatever the batch size.  Same key layout as :func:`pbs_mb2`."""
    dev = ct.device
    ct = ct.contiguous()
    B = ct.shape[0]
    assert ct.shape[1] == p.n + 1
    luts = luts.to(device=dev, dtype=torch.int64).reshape(-1, p.N).contiguous()
    if out is None:
        out = torch.empty((B, p.k * p.N + 1), dtype=torch.int64, device=dev)
    li = None
    if lut_index is not None:
        lut_index = lut_index.to(device=dev, dtype=torch.int32).contiguous()
        li = _ptr(lut_index)
    N.check(N.lib().fhe_b200_pbs_mb2_wide(_ctx(dev).handle, C.byref(p), _ptr(bskf2), _ptr(ct), B, _ptr(luts), li,
                                          _ptr(out), _stream(dev)))
    return out


def pbs_mb2_pair(p: N.PBSParams, bskf2: torch.Tensor, ct: torch.Tensor, luts: torch.Tensor,
                 lut_index: torch.Tensor | None = None, out: torch.Tensor | None = None) -> torch.Tensor:
    """:func:`pbs_mb2` with every ciphertext on a cluster of two CTAs (polynomial t on CTA t, spectra exchanged through
    distributed shared memory; pbs_wide.cu) whatever the batch size.  Same key layout as :func:`pbs_mb2`."""
    dev = ct.device
    ct = ct.contiguous()
    B = ct.shape[0]
    assert ct.shape[1] == p.n + 1
    luts = luts.to(device=dev, dtype=torch.int64).reshape(-1, p.N).contiguous()
    if out is None:
        out = torch.empty((B, p.k * p.N + 1), dtype=torch.int64, device=dev)
    li = None
    if lut_index is not None:
        lut_index = lut_index.to(device=dev, dtype=torch.int32).contiguous()
        li = _ptr(lut_index)
    N.check(N.lib().fhe_b200_pbs_mb2_pair(_ctx(dev).handle, C.byref(p), _ptr(bskf2), _ptr(ct), B, _ptr(luts), li,
                                          _ptr(out), _stream(dev)))
    return out


def make_lut_poly(table, p_bits: int, N_poly: int, delta_out_log2: int) -> np.ndarray:
    """Accumulator polynomial of a p-bit table lookup (message + 1 padding bit): box m holds
    table[m] << delta_out_log2 and the polynomial is multiplied by X^(-box/2)."""
    box = N_poly >> p_bits
    t = np.asarray(table, dtype=np.int64)
    if t.size != (1 << p_bits):
        raise ValueError("table must have 2^p_bits entries")
    p0 = np.repeat(t, box).astype(np.uint64) << np.uint64(delta_out_log2)
    half = box // 2
    out = np.empty(N_poly, dtype=np.uint64)
    out[: N_poly - half] = p0[half:]
    out[N_poly - half:] = np.uint64(0) - p0[:half]
    return out


# ------------------------------------------------------------------------------- packed inner products (GLWE x GGSW)
def glwe_encrypt_vectors(p: N.PBSParams, S_big: torch.Tensor, polys: torch.Tensor, shift: int, seed: int,
                         id_base: int = 0, noise_seed: int | None = None) -> torch.Tensor:
    """polys int64 [rows, N] (message polynomials) -> GLWE(poly << shift) [rows, k+1, N]."""
    dev = S_big.device
    polys = polys.to(device=dev, dtype=torch.int64).contiguous()
    rows = polys.shape[0]
    assert polys.shape[1] == p.N
    out = torch.empty((rows, p.k + 1, p.N), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_glwe_encrypt_rows(_ctx(dev).handle, C.byref(p), _ptr(S_big), _ptr(polys), rows, p.N, 0, shift,
                                               seed, _noise(noise_seed), id_base, _ptr(out), _stream(dev)))
    return out


def ggsw_encrypt_poly(p: N.PBSParams, S_big: torch.Tensor, poly: torch.Tensor, seed: int, id_base: int = 0,
                      noise_seed: int | None = None) -> torch.Tensor:
    """poly int64 [N] -> GGSW rows [(k+1)*l, k+1, N] (row t*l+lev carries poly << (64 - beta*(lev+1)) on component t)."""
    dev = S_big.device
    poly = poly.to(device=dev, dtype=torch.int64).contiguous()
    assert poly.numel() == p.N
    rows = (p.k + 1) * p.l_pbs
    out = torch.empty((rows, p.k + 1, p.N), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_glwe_encrypt_rows(_ctx(dev).handle, C.byref(p), _ptr(S_big), _ptr(poly), rows, 0, 1, 0, seed,
                                               _noise(noise_seed), id_base, _ptr(out), _stream(dev)))
    return out


def ggsw_to_fourier(p: N.PBSParams, ggsw: torch.Tensor) -> torch.Tensor:
    """GGSW rows -> Fourier [1][k+1][l][k+1][N/2][2] f64 (the bootstrapping-key transform with n = 1)."""
    p1 = N.PBSParams(1, p.k, p.N, p.l_pbs, p.beta_pbs, p.l_ks, p.beta_ks, 0, p.sigma_lwe_abs, p.sigma_glwe_abs)
    return bsk_to_fourier(p1, ggsw.contiguous())


def glwe_ggsw_dot(p: N.PBSParams, ggswf: torch.Tensor, glwe: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """GGSW (Fourier) external product with every GLWE of the batch: [G, k+1, N] -> [G, k+1, N]."""
    dev = glwe.device
    assert glwe.is_contiguous() and glwe.shape[1:] == (p.k + 1, p.N)
    if out is None:
        out = torch.empty_like(glwe)
    N.check(N.lib().fhe_b200_glwe_ggsw_dot(_ctx(dev).handle, C.byref(p), _ptr(ggswf), _ptr(glwe), glwe.shape[0],
                                           _ptr(out), _stream(dev)))
    return out


def glwe_decrypt_coeffs(p: N.PBSParams, S_big: torch.Tensor, glwe: torch.Tensor, first: int, step: int, count: int,
                        shift: int) -> torch.Tensor:
    dev = glwe.device
    assert glwe.is_contiguous()
    out = torch.empty((glwe.shape[0], count), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_glwe_decrypt_coeffs(_ctx(dev).handle, C.byref(p), _ptr(S_big), _ptr(glwe), glwe.shape[0],
                                                 first, step, count, shift, _ptr(out), _stream(dev)))
    return out


def glwe_sample_extract(p: N.PBSParams, glwe: torch.Tensor, first: int, step: int, count: int) -> torch.Tensor:
    """-> LWE rows [G*count, even_stride(kN)] under the big key."""
    dev = glwe.device
    assert glwe.is_contiguous()
    stride = even_stride(p.k * p.N)
    out = torch.empty((glwe.shape[0] * count, stride), dtype=torch.int64, device=dev)
    N.check(N.lib().fhe_b200_glwe_sample_extract(_ctx(dev).handle, C.byref(p), _ptr(glwe), glwe.shape[0], first, step,
                                                 count, stride, _ptr(out), _stream(dev)))
    return out
