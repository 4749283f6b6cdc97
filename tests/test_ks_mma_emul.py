"""CPU emulation of the tensor-core keyswitch (fhe_icp_b200/csrc/ks_mma.cu): the operand blocks are produced by
the product's own __host__ __device__ builders (ks_mma_layout.cuh), read back through the shared-memory
descriptor addressing the MMA uses, contracted as s8 x u8 -> s32 and recombined like the kernel's epilogue.
The result must equal the oracle's 32-bit keyswitch word for word -- no GPU needed."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emul(tmp_path_factory):
    so = tmp_path_factory.mktemp("emul") / "libksmmaemul.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(so),
                    str(ROOT / "tests" / "emul" / "ks_mma_emul.cpp")], check=True)
    lib = C.CDLL(str(so))
    lib.emul_keyswitch_mma.restype = C.c_int
    lib.emul_keyswitch_mma.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    return lib


@pytest.mark.parametrize("n,l,beta,B", [(7, 5, 3, 3), (70, 3, 4, 2), (63, 2, 8, 1)])
def test_block_layout_contraction_equals_oracle_keyswitch32(emul, O, n, l, beta, B):
    """n = 7: one (mostly padded) column tile; n = 70: 284 byte columns = two tiles with a ragged last one;
    n = 63: exactly one full tile; beta = 8 exercises digits down to -128."""
    p = O.make_params(n=n, k=1, N=2048, l_ks=l, beta_ks=beta)
    kN = 2048
    rng = np.random.RandomState(n)
    ksk32 = rng.randint(0, 2 ** 32, size=(kN, l, n + 1), dtype=np.uint64).astype(np.uint32)
    ct = rng.randint(0, 2 ** 63, size=(B, kN + 1), dtype=np.uint64) * np.uint64(2) + rng.randint(0, 2, size=(B, kN + 1)).astype(np.uint64)
    ct[0, :8] = [0, 2 ** 64 - 1, 2 ** 63, 2 ** 63 - 1, 1 << (64 - l * beta), (1 << (63 - l * beta)) - 1, 1 << (63 - l * beta), 12345]
    out = np.full((B, n + 1), 0xDEAD, dtype=np.uint64)
    rc = emul.emul_keyswitch_mma(ksk32.ctypes.data, ct.ctypes.data, B, kN, n, l, beta, out.ctypes.data)
    assert rc == 0
    ref = O.keyswitch32(p, ksk32, ct)
    assert np.array_equal(out, ref)
