"""Cryptographic parameter selection for the leveled (linear) circuit and the stated
parameter set for keyswitch + PBS.

The reference leaves this to Concrete's optimizer (``model.compile``,
/root/reference/fhe_similarity.py:120); a linear circuit only needs the LWE dimension n
and the fresh-noise level.  Security curve (128-bit, q = 2^64, binary keys), as used by the
Concrete/TFHE-rs parameter sets: log2(sigma) = -0.0265 * n + 2.6
(two published points: n=742 -> 2^-17.1, n=2048 -> 2^-51.6).
"""
from __future__ import annotations

import math
from dataclasses import asdict, dataclass

SECURITY_SLOPE = -0.0265
SECURITY_OFFSET = 2.6
MIN_LOG2_SIGMA = -62.0   # one torus step is 2^-64: below this the Gaussian is degenerate
DEFAULT_P_ERROR = 2.0 ** -40


def log2_sigma_for_dimension(n: int) -> float:
    return max(SECURITY_SLOPE * n + SECURITY_OFFSET, MIN_LOG2_SIGMA)


def z_score(p_error: float) -> float:
    """z such that P(|N(0,1)| > z) = p_error."""
    lo, hi = 0.0, 40.0
    for _ in range(200):
        mid = 0.5 * (lo + hi)
        if math.erfc(mid / math.sqrt(2.0)) > p_error:
            lo = mid
        else:
            hi = mid
    return hi


@dataclass
class LweParams:
    n: int                # LWE dimension
    stride: int           # u64 words per ciphertext row (even)
    shift: int            # log2(Delta)
    msg_bits: int         # signed width of the widest encrypted value
    log2_sigma: float     # fresh noise std relative to the torus
    p_error: float
    log2_out_noise: float  # predicted std of the noisiest output, relative to the torus

    @property
    def sigma_abs(self) -> float:
        return 2.0 ** (64.0 + self.log2_sigma)

    def to_dict(self) -> dict:
        return asdict(self)


def select_lwe_params(msg_bits: int, w_norm2_sq: float, p_error: float = DEFAULT_P_ERROR, align_words: int = 16,
                      padding_bits: int = 1) -> LweParams:
    """Smallest n (with n+1 a multiple of ``align_words``) such that a linear combination with
    squared weight norm ``w_norm2_sq`` of fresh ciphertexts decodes a ``msg_bits``-bit signed
    message with failure probability <= p_error:   z * sigma * ||w||_2 < Delta / 2."""
    if msg_bits + padding_bits >= 63:
        raise ValueError(f"NoParametersFound: a {msg_bits}-bit message does not fit the 64-bit torus")
    shift = 64 - msg_bits - padding_bits
    z = z_score(p_error)
    log2_need = (shift - 1 - 64) - math.log2(z) - 0.5 * math.log2(max(w_norm2_sq, 1.0))
    if log2_need < MIN_LOG2_SIGMA:
        raise ValueError("NoParametersFound: required noise is below the representable minimum "
                         f"(msg_bits={msg_bits}, ||w||^2={w_norm2_sq:g})")
    n_min = math.ceil((SECURITY_OFFSET - log2_need) / -SECURITY_SLOPE)
    n_min = max(n_min, 256)
    words = ((n_min + 1 + align_words - 1) // align_words) * align_words
    n = words - 1
    log2_sigma = log2_sigma_for_dimension(n)
    stride = (n + 2) & ~1
    return LweParams(n=n, stride=stride, shift=shift, msg_bits=msg_bits, log2_sigma=log2_sigma, p_error=p_error,
                     log2_out_noise=log2_sigma + 0.5 * math.log2(max(w_norm2_sq, 1.0)))


# The keyswitch + PBS parameter set (4-bit message space: 2 message + 2 carry bits, 1 padding
# bit; p_fail ~ 2^-40).  The reference's compiled circuit fixes none (it has no table lookup);
# this is the widely published TFHE-rs/Concrete "MESSAGE_2_CARRY_2_KS_PBS" set, restated.
PBS_PARAMS_4BIT = dict(n=742, k=1, N_poly=2048, l_pbs=1, beta_pbs=23, l_ks=5, beta_ks=3,
                       log2_sigma_lwe=-17.1, log2_sigma_glwe=-51.6)
