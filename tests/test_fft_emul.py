"""CPU emulation of one warp of the product's 32x32 negacyclic FFT (fhe_icp_b200/csrc/fft.cuh is
__host__ __device__): checks the index algebra, the swizzled transpose and the twiddle table
against numpy and against the oracle's independent radix-2 transform -- no GPU needed."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = tmp_path_factory.mktemp("emul") / "libfftemul.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-o", str(so),
                    str(ROOT / "tests" / "emul" / "fft_emul.cpp")], check=True)
    return C.CDLL(str(so))


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def test_forward_matches_numpy(emul):
    rng = np.random.RandomState(0)
    coef = rng.randint(-2 ** 22, 2 ** 22, size=2048).astype(np.float64)
    bins = np.zeros(2048)
    emul.emul_forward(_dp(coef), _dp(bins))
    j = np.arange(1024)
    z = (coef[:1024] + 1j * coef[1024:]) * np.exp(1j * np.pi * j / 2048)
    ref = np.fft.ifft(z) * 1024  # sum_j z_j exp(+2 pi i jk / 1024)
    got = bins[0::2] + 1j * bins[1::2]
    assert np.abs(got - ref).max() / np.abs(ref).max() < 1e-14


def test_roundtrip_and_negacyclic_product(emul, O):
    rng = np.random.RandomState(1)
    a = rng.randint(-2 ** 22, 2 ** 22, size=2048)
    b = rng.randint(-2 ** 27, 2 ** 27, size=2048)  # |a*b|*N < 2^61: no wrap in the exact product
    fa, fb = np.zeros(2048), np.zeros(2048)
    emul.emul_forward(_dp(a.astype(np.float64)), _dp(fa))
    emul.emul_forward(_dp(b.astype(np.float64)), _dp(fb))
    back = np.zeros(2048)
    emul.emul_inverse(_dp(fa), _dp(back))
    assert np.abs(back - a).max() < 1e-6
    prod = (fa[0::2] + 1j * fa[1::2]) * (fb[0::2] + 1j * fb[1::2])
    fp = np.empty(2048)
    fp[0::2], fp[1::2] = prod.real, prod.imag
    res = np.zeros(2048)
    emul.emul_inverse(_dp(fp), _dp(res))
    # exact negacyclic product (python ints), compared modulo rounding of the f64 pipeline
    exact = O.negacyclic_mul_naive(a, b.astype(np.int64).view(np.uint64)).view(np.int64).astype(np.float64)
    assert np.abs(res - exact).max() / np.abs(exact).max() < 1e-12


def test_bins_equal_oracle_fourier_layout(emul, O):
    """Natural-order bins == the oracle's orc_bsk_to_fourier layout (so the GPU key can be diffed)."""
    p = O.make_params(n=1, k=1, N=2048, l_pbs=1, beta_pbs=23)
    rng = np.random.RandomState(2)
    poly = rng.randint(-2 ** 62, 2 ** 62, size=(1, 2, 1, 2, 2048), dtype=np.int64).view(np.uint64)
    of = O.bsk_to_fourier(p, poly)[0, 0, 0, 0]
    bins = np.zeros(2048)
    emul.emul_forward(_dp(poly[0, 0, 0, 0].view(np.int64).astype(np.float64)), _dp(bins))
    assert np.abs(bins.reshape(1024, 2) - of).max() / np.abs(of).max() < 1e-14
