"""Uniform quantizers and the quantized linear circuit.

The reference gets these from Concrete-ML (un-vendored: concrete-ml==1.9.0,
/root/reference/requirements.txt:5) when it builds ``LinearRegression(n_bits=8)``
(/root/reference/fhe_similarity.py:88-94).  Behaviour follows SURVEY.md Appendix A.1/A.2;
every constant is an explicit, serialisable field so that a real Concrete-ML dump can be
diffed against it later.
"""
from __future__ import annotations

from dataclasses import asdict, dataclass

import numpy as np

STABILITY_CONST = 1e-6


@dataclass
class UniformQuantizer:
    n_bits: int
    is_signed: bool
    scale: float
    zero_point: int
    offset: int

    @classmethod
    def from_values(cls, values: np.ndarray, n_bits: int, is_signed: bool = True) -> "UniformQuantizer":
        v = np.asarray(values, dtype=np.float64)
        offset = 2 ** (n_bits - 1) if is_signed else 0
        rmin, rmax = float(v.min()), float(v.max())
        if abs(rmax - rmin) < STABILITY_CONST:
            # constant tensor: quantizes to 0 (if ~0) or to 1 with scale = the value
            if abs(rmax) < STABILITY_CONST:
                return cls(n_bits, is_signed, 1.0, 0, offset)
            return cls(n_bits, is_signed, rmax, 0, offset)
        scale = (rmax - rmin) / (2 ** n_bits - 1)
        zp = int(np.round((rmax * (-offset) - rmin * (2 ** n_bits - 1 - offset)) / (rmax - rmin)))
        return cls(n_bits, is_signed, scale, zp, offset)

    @property
    def qmin(self) -> int:
        return -self.offset

    @property
    def qmax(self) -> int:
        return 2 ** self.n_bits - 1 - self.offset

    def quant(self, values: np.ndarray) -> np.ndarray:
        q = np.rint(np.asarray(values, dtype=np.float64) / self.scale + self.zero_point)
        return np.clip(q, self.qmin, self.qmax).astype(np.int64)

    def dequant(self, q: np.ndarray) -> np.ndarray:
        return self.scale * (np.asarray(q, dtype=np.int64) - self.zero_point).astype(np.float64)

    def to_dict(self) -> dict:
        return asdict(self)


@dataclass
class QuantizedLinearSpec:
    """Everything the integer circuit needs: q_y = q_X @ q_W - zp_W * sum_j q_X[j] + q_bias,
    y = out_scale * (q_y - out_zero_point)."""
    d: int
    input_q: UniformQuantizer
    weight_q: UniformQuantizer
    q_weights: np.ndarray      # int64 [d]
    q_bias: int
    out_scale: float
    out_zero_point: int        # already doubled

    @classmethod
    def from_fit(cls, coef: np.ndarray, intercept: float, X_calib: np.ndarray, n_bits: int) -> "QuantizedLinearSpec":
        coef = np.asarray(coef).reshape(-1)
        d = coef.size
        input_q = UniformQuantizer.from_values(X_calib, n_bits, is_signed=True)
        weight_q = UniformQuantizer.from_values(coef, n_bits, is_signed=True)
        q_w = weight_q.quant(coef)
        out_scale = input_q.scale * weight_q.scale
        zp_out = int(input_q.zero_point) * (int(q_w.sum()) - d * int(weight_q.zero_point))
        q_bias = int(np.rint(float(intercept) / out_scale + zp_out))
        return cls(d, input_q, weight_q, q_w, q_bias, out_scale, 2 * zp_out)

    def circuit(self, q_X: np.ndarray) -> np.ndarray:
        q_X = np.asarray(q_X, dtype=np.int64)
        return q_X @ self.q_weights - int(self.weight_q.zero_point) * q_X.sum(axis=1) + int(self.q_bias)

    def dequantize_output(self, q_y: np.ndarray) -> np.ndarray:
        return self.out_scale * (np.asarray(q_y, dtype=np.int64) - int(self.out_zero_point)).astype(np.float64)

    def predict_clear(self, X: np.ndarray) -> np.ndarray:
        return self.dequantize_output(self.circuit(self.input_q.quant(X)))

    def to_dict(self) -> dict:
        return {
            "d": self.d, "input_q": self.input_q.to_dict(), "weight_q": self.weight_q.to_dict(),
            "q_weights": [int(x) for x in self.q_weights], "q_bias": int(self.q_bias),
            "out_scale": float(self.out_scale), "out_zero_point": int(self.out_zero_point),
        }

    @classmethod
    def from_dict(cls, dct: dict) -> "QuantizedLinearSpec":
        return cls(dct["d"], UniformQuantizer(**dct["input_q"]), UniformQuantizer(**dct["weight_q"]),
                   np.asarray(dct["q_weights"], dtype=np.int64), int(dct["q_bias"]), float(dct["out_scale"]),
                   int(dct["out_zero_point"]))


def signed_bit_width(lo: int, hi: int) -> int:
    """Bits of the smallest two's-complement (or unsigned, if lo >= 0) integer type holding [lo, hi]."""
    lo, hi = int(lo), int(hi)
    if lo >= 0:
        return max(1, hi.bit_length())
    return max((-lo - 1).bit_length(), hi.bit_length()) + 1
