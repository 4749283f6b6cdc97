"""CPU emulation of a complete multi-bit blind rotation in the two-warps-per-polynomial form
(fhe_icp_b200/csrc/pbs_split.cuh on fft_split.cuh; groundwork for the next blind-rotation kernel):
every message maps to LUT[m] and the phases agree with the oracle's multi-bit PBS within the PBS noise bound --
the same acceptance the GPU kernel has (tests/test_gpu_pbs.py::test_multibit_pbs)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = tmp_path_factory.mktemp("emul") / "libpbssplitemul.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(so),
                    str(ROOT / "tests" / "emul" / "pbs_split_emul.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.emul_pbs_mb2_split.restype = C.c_int
    lib.emul_pbs_mb2_split.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.emul_pbs_mb2_split_aliased.restype = C.c_int
    lib.emul_pbs_mb2_split_aliased.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    return lib


def test_split_blind_rotation_equals_oracle_multibit_pbs(emul, O):
    n = 12
    p = O.make_params(n=n, k=1, N=2048, l_pbs=1, beta_pbs=23, log2_sigma_lwe=-30.0, log2_sigma_glwe=-51.6)
    s, S = O.secret_key(11, 0, n), O.secret_key(11, 1, 2048)
    of = O.bsk2_to_fourier(p, O.bsk2_gen(p, s, S, 22))                     # [i][g][t][l][c][M][2], natural bins
    blocks = np.ascontiguousarray(of.reshape(of.shape[0], 3, 2, 1, 2, 32, 32, 2).transpose(0, 5, 1, 2, 3, 4, 6, 7))  # by frequency block
    msgs = np.arange(16)
    ct = O.lwe_encrypt(s, msgs, 59, p.sigma_lwe_abs, enc_seed=5, ct_base=100, stride=n + 2)[:, : n + 1].copy()
    table = (np.arange(16) * 5 + 2) % 16
    lut = O.make_lut_poly(table, 4, 2048, 59)
    got = np.zeros((16, 2049), dtype=np.uint64)
    for b in range(16):
        row = np.ascontiguousarray(ct[b])
        assert emul.emul_pbs_mb2_split(blocks.ctypes.data, row.ctypes.data, n, 23, lut.ctypes.data, got[b].ctypes.data) == 0
    assert np.array_equal(O.lwe_decrypt(S, got, 59) & 15, table[msgs])
    ref = O.pbs_mb2(p, of, ct, lut)
    diff = (O.lwe_phase(S, got) - O.lwe_phase(S, ref)).view(np.int64).astype(np.float64)
    assert np.log2(np.abs(diff).max() + 1) - 64 < -12
    err = (O.lwe_phase(S, got) - (table[msgs].astype(np.uint64) << np.uint64(59))).view(np.int64).astype(np.float64)
    assert np.log2(err.std() + 1) - 64 < -13.5


def test_kernel_memory_plan_is_order_independent(emul, O):
    """The kernel's shared-memory plan (csrc/pbs_split.cu): ONE region per polynomial reused as E|O tiles, half-spectra,
    exchanged pointwise halves and inverse tile, with phase boundaries exactly where the kernel has its named barriers.
    Whatever order the warps and lanes of a phase run in, the output is bit-identical to the un-aliased emulation --
    a hand-over without a barrier would make it depend on the order."""
    n = 8
    p = O.make_params(n=n, k=1, N=2048, l_pbs=1, beta_pbs=23, log2_sigma_lwe=-30.0, log2_sigma_glwe=-51.6)
    s, S = O.secret_key(3, 0, n), O.secret_key(3, 1, 2048)
    of = O.bsk2_to_fourier(p, O.bsk2_gen(p, s, S, 4))
    blocks = np.ascontiguousarray(of.reshape(of.shape[0], 3, 2, 1, 2, 32, 32, 2).transpose(0, 5, 1, 2, 3, 4, 6, 7))
    msgs = np.array([0, 5, 9, 15])
    ct = O.lwe_encrypt(s, msgs, 59, p.sigma_lwe_abs, enc_seed=6, ct_base=7, stride=n + 2)[:, : n + 1].copy()
    table = (np.arange(16) * 3 + 1) % 16
    lut = O.make_lut_poly(table, 4, 2048, 59)
    for b in range(len(msgs)):
        row = np.ascontiguousarray(ct[b])
        want = np.zeros(2049, dtype=np.uint64)
        assert emul.emul_pbs_mb2_split(blocks.ctypes.data, row.ctypes.data, n, 23, lut.ctypes.data, want.ctypes.data) == 0
        for order in (0, 1, 2):
            got = np.full(2049, 0xDEAD, dtype=np.uint64)
            assert emul.emul_pbs_mb2_split_aliased(blocks.ctypes.data, row.ctypes.data, n, 23, lut.ctypes.data, order, got.ctypes.data) == 0
            assert np.array_equal(got, want), (b, order)
        assert (O.lwe_decrypt(S, want[None, :], 59) & 15)[0] == table[msgs[b]]


def test_ring_order_gives_each_half_its_block(emul):
    """Slice s of a pair holds frequency blocks {s, 16 + s}: in step s half h reads block h of the slice and must find
    k1 = 16h + s there; the 32 positions are a permutation of the 32 blocks."""
    got = [emul.emul_split_ring_block(pos) for pos in range(32)]
    assert sorted(got) == list(range(32))
    for s in range(16):
        for h in range(2):
            assert got[2 * s + h] == 16 * h + s
