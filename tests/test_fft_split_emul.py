"""CPU emulation of the two-warp form of the negacyclic FFT (fhe_icp_b200/csrc/fft_split.cuh: every 32-point
in-register DFT split by one radix-2 step between two warps, the closing butterflies deferred to the reader of the
shared-memory tile).  No kernel uses it yet -- this pins the index algebra for the blind-rotation redesign
 against numpy, against the exact negacyclic product and against the one-warp transform."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    d = tmp_path_factory.mktemp("emul")
    libs = []
    for name in ("fft_split_emul", "fft_emul"):
        so = d / f"lib{name}.so"
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", str(so), str(ROOT / "tests" / "emul" / f"{name}.cpp")],
                       check=True)
        libs.append(C.CDLL(str(so)))
    return libs


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_split_forward_matches_numpy_and_one_warp_form(emul):
    split, one = emul
    rng = np.random.RandomState(0)
    coef = rng.randint(-2 ** 22, 2 ** 22, size=2048).astype(np.float64)
    bins, bins1 = np.zeros(2048), np.zeros(2048)
    split.emul_split_forward(_dp(coef), _dp(bins))
    one.emul_forward(_dp(coef), _dp(bins1))
    j = np.arange(1024)
    z = (coef[:1024] + 1j * coef[1024:]) * np.exp(1j * np.pi * j / 2048)
    ref = np.fft.ifft(z) * 1024  # sum_j z_j exp(+2 pi i jk / 1024)
    got = bins[0::2] + 1j * bins[1::2]
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-14
    assert np.abs(bins - bins1).max() / np.abs(bins1).max() < 1e-14     # same bins, same natural order


def test_split_inverse_roundtrip_and_cross_form(emul):
    """split forward -> split inverse, split forward -> one-warp inverse and one-warp forward -> split inverse all
    return the coefficients: the two forms are interchangeable stage by stage (same Fourier-key layout)."""
    split, one = emul
    rng = np.random.RandomState(1)
    a = rng.randint(-2 ** 22, 2 ** 22, size=2048).astype(np.float64)
    fs, f1 = np.zeros(2048), np.zeros(2048)
    split.emul_split_forward(_dp(a), _dp(fs))
    one.emul_forward(_dp(a), _dp(f1))
    for fwd, inv in ((fs, split.emul_split_inverse), (fs, one.emul_inverse), (f1, split.emul_split_inverse)):
        back = np.zeros(2048)
        inv(_dp(fwd), _dp(back))
        assert np.abs(back - a).max() < 1e-6


def test_split_negacyclic_product(emul, O):
    split, _ = emul
    rng = np.random.RandomState(2)
    a = rng.randint(-2 ** 22, 2 ** 22, size=2048)
    b = rng.randint(-2 ** 27, 2 ** 27, size=2048)  # |a*b|*N < 2^61: no wrap in the exact product
    fa, fb = np.zeros(2048), np.zeros(2048)
    split.emul_split_forward(_dp(a.astype(np.float64)), _dp(fa))
    split.emul_split_forward(_dp(b.astype(np.float64)), _dp(fb))
    prod = (fa[0::2] + 1j * fa[1::2]) * (fb[0::2] + 1j * fb[1::2])
    fp = np.empty(2048)
    fp[0::2], fp[1::2] = prod.real, prod.imag
    res = np.zeros(2048)
    split.emul_split_inverse(_dp(fp), _dp(res))
    exact = O.negacyclic_mul_naive(a, b.astype(np.int64).view(np.uint64)).view(np.int64).astype(np.float64)
    assert np.abs(res - exact).max() / np.abs(exact).max() < 1e-12
