"""CPU emulation of the four-warps-per-polynomial blind rotation (fhe_icp_b200/csrc/pbs_wide.cuh, the arithmetic and the
shared-memory plan of pbs_kernel_mb2_wide): the 8 x 8 x 16 transform equals numpy's FFT, a complete multi-bit blind
rotation maps every message to LUT[m] with phases within the PBS noise bound of the oracle's, and the result does not
depend on the order in which the threads of a phase (or the two polynomials between the 256-thread barriers) run."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = tmp_path_factory.mktemp("emul") / "libpbswideemul.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(so),
                    str(ROOT / "tests" / "emul" / "pbs_wide_emul.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.emul_wide_fft.restype = C.c_int
    lib.emul_wide_fft.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    lib.emul_pbs_mb2_wide.restype = C.c_int
    lib.emul_pbs_mb2_wide.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    lib.emul_pbs_mb2_pair.restype = C.c_int
    lib.emul_pbs_mb2_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    return lib


def test_wide_transform_equals_numpy(emul):
    rng = np.random.default_rng(3)
    z = rng.integers(-(1 << 22), 1 << 22, size=1024).astype(np.float64) + 1j * rng.integers(-(1 << 22), 1 << 22, size=1024)
    j = np.arange(1024)
    want = np.fft.ifft(z * np.exp(2j * np.pi * j / 4096)) * 1024          # F[k] = sum_j z_j omega^j W^(jk), W = e^(+2 pi i/1024)
    for order in (0, 1, 2):
        got = np.zeros(1024, dtype=np.complex128)
        assert emul.emul_wide_fft(np.ascontiguousarray(z).ctypes.data, 1, order, got.ctypes.data) == 0
        assert np.abs(got - want).max() < 1e-9 * np.abs(want).max()
        back = np.zeros(1024, dtype=np.complex128)
        assert emul.emul_wide_fft(got.ctypes.data, -1, order, back.ctypes.data) == 0
        assert np.abs(back - z).max() < 1e-6                              # untwisted and scaled by 1/1024


def _setup(O, n, key_seed, evk_seed, msgs, enc_seed):
    p = O.make_params(n=n, k=1, N=2048, l_pbs=1, beta_pbs=23, log2_sigma_lwe=-30.0, log2_sigma_glwe=-51.6)
    s, S = O.secret_key(key_seed, 0, n), O.secret_key(key_seed, 1, 2048)
    of = O.bsk2_to_fourier(p, O.bsk2_gen(p, s, S, evk_seed))               # [i][g][t][l][c][M][2], natural bins
    blocks = np.ascontiguousarray(of.reshape(of.shape[0], 3, 2, 1, 2, 32, 32, 2).transpose(0, 5, 1, 2, 3, 4, 6, 7))
    ct = O.lwe_encrypt(s, msgs, 59, p.sigma_lwe_abs, enc_seed=enc_seed, ct_base=100, stride=n + 2)[:, : n + 1].copy()
    return p, S, of, blocks, ct


def test_wide_blind_rotation_equals_oracle_multibit_pbs(emul, O):
    n = 12
    msgs = np.arange(16)
    p, S, of, blocks, ct = _setup(O, n, 11, 22, msgs, 5)
    table = (np.arange(16) * 5 + 2) % 16
    lut = O.make_lut_poly(table, 4, 2048, 59)
    got = np.zeros((16, 2049), dtype=np.uint64)
    for b in range(16):
        row = np.ascontiguousarray(ct[b])
        assert emul.emul_pbs_mb2_wide(blocks.ctypes.data, row.ctypes.data, n, 23, lut.ctypes.data, 0, got[b].ctypes.data) == 0
    assert np.array_equal(O.lwe_decrypt(S, got, 59) & 15, table[msgs])
    ref = O.pbs_mb2(p, of, ct, lut)
    diff = (O.lwe_phase(S, got) - O.lwe_phase(S, ref)).view(np.int64).astype(np.float64)
    assert np.log2(np.abs(diff).max() + 1) - 64 < -12
    err = (O.lwe_phase(S, got) - (table[msgs].astype(np.uint64) << np.uint64(59))).view(np.int64).astype(np.float64)
    assert np.log2(err.std() + 1) - 64 < -13.5


def test_wide_memory_plan_is_order_independent(emul, O):
    """Two exchange buffers per polynomial written alternately, barriers where the kernel has them: whatever order the
    threads of a phase run in, and whichever polynomial runs ahead between the 256-thread barriers, the output is
    bit-identical (a hand-over without a barrier, or a read of a slot nobody wrote -- the buffers start as NaN --
    would change it)."""
    n = 8
    msgs = np.array([0, 5, 9, 15])
    p, S, of, blocks, ct = _setup(O, n, 3, 4, msgs, 6)
    table = (np.arange(16) * 3 + 1) % 16
    lut = O.make_lut_poly(table, 4, 2048, 59)
    for b in range(len(msgs)):
        row = np.ascontiguousarray(ct[b])
        outs = []
        for order in (0, 1, 2):
            got = np.full(2049, 0xDEAD, dtype=np.uint64)
            assert emul.emul_pbs_mb2_wide(blocks.ctypes.data, row.ctypes.data, n, 23, lut.ctypes.data, order, got.ctypes.data) == 0
            outs.append(got)
        assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2]), b
        assert (O.lwe_decrypt(S, outs[0][None, :], 59) & 15)[0] == table[msgs[b]]


def test_pair_kernel_pointwise_on_the_column_layout(emul, O):
    """pbs_kernel_mb2_pair (two CTAs per ciphertext) reads the key by output column and forms S_own, S_oth and F_t * S_own
    before the other polynomial's spectrum arrives: the same decryptions, and phases within rounding of the one-CTA form."""
    n = 8
    msgs = np.array([1, 6, 10, 15])
    p, S, of, blocks, ct = _setup(O, n, 5, 6, msgs, 9)
    # [i][k1][g][t'][l=1][c][32][2] -> [i][c][k1][g][t'][32][2]  (bsk2_column_split_kernel)
    cols = np.ascontiguousarray(blocks[:, :, :, :, 0].transpose(0, 4, 1, 2, 3, 5, 6))
    table = (np.arange(16) * 7 + 3) % 16
    lut = O.make_lut_poly(table, 4, 2048, 59)
    for b in range(len(msgs)):
        row = np.ascontiguousarray(ct[b])
        one = np.zeros(2049, dtype=np.uint64)
        two = np.zeros(2049, dtype=np.uint64)
        assert emul.emul_pbs_mb2_wide(blocks.ctypes.data, row.ctypes.data, n, 23, lut.ctypes.data, 0, one.ctypes.data) == 0
        assert emul.emul_pbs_mb2_pair(cols.ctypes.data, row.ctypes.data, n, 23, lut.ctypes.data, 1, two.ctypes.data) == 0
        assert (O.lwe_decrypt(S, two[None, :], 59) & 15)[0] == table[msgs[b]]
        diff = (O.lwe_phase(S, one[None, :]) - O.lwe_phase(S, two[None, :])).view(np.int64).astype(np.float64)
        assert np.log2(np.abs(diff).max() + 1) - 64 < -12
