"""ctypes binding of the C-ABI in ``include/fhe_b200.h`` (libfhe_b200.so).

The library is built in-tree by :func:`build` (``nvcc -gencode arch=compute_100a,code=sm_100a``)
and has no dependency on torch; torch is only used by callers to own device memory and
streams.  There is no fallback: if the shared object is missing or no B200 is visible the
functions here raise.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
from pathlib import Path

import numpy as np

_PKG = Path(__file__).resolve().parent
_CSRC = _PKG / "csrc"
_SO = _PKG / "libfhe_b200.so"
_SOURCES = ["api.cu", "lwe.cu", "keys.cu", "keyswitch.cu", "ks_mma.cu", "pbs.cu", "pbs_wide.cu", "probe.cu"]
_HEADERS = ["common.cuh", "kernels.h", "fft.cuh", "lwe_device.cuh", "ks_mma_layout.cuh", "pbs_wide.cuh", "../../include/fhe_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]

OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_STATE = 0, 1, 2, 3, 4
KIND_SK, KIND_MASK, KIND_NOISE = 1, 2, 3
PUR_INPUT, PUR_KSK, PUR_BSK, PUR_BSK2 = 0, 1, 2, 3


class FheB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"fhe_b200 error {code}: {msg}")
        self.code = code


class PBSParams(C.Structure):
    _fields_ = [
        ("n", C.c_int32), ("k", C.c_int32), ("N", C.c_int32),
        ("l_pbs", C.c_int32), ("beta_pbs", C.c_int32),
        ("l_ks", C.c_int32), ("beta_ks", C.c_int32), ("_pad", C.c_int32),
        ("sigma_lwe_abs", C.c_double), ("sigma_glwe_abs", C.c_double),
    ]


class SimilaritySpec(C.Structure):
    _fields_ = [
        ("d", C.c_int32), ("n_bits", C.c_int32), ("n", C.c_int32), ("stride", C.c_int32),
        ("shift", C.c_int32), ("two_outputs", C.c_int32),
        ("sigma_abs", C.c_double), ("x_scale", C.c_double),
        ("x_zero_point", C.c_int64), ("x_offset", C.c_int64),
        ("w_zero_point", C.c_int64), ("q_bias", C.c_int64),
        ("out_scale", C.c_double), ("out_zero_point", C.c_int64),
        ("key_seed", C.c_uint64), ("noise_seed", C.c_uint64),
    ]


class Push(C.Structure):
    """``fhe_b200_push``: destination of one evaluation pushed to the client's score board."""
    _fields_ = [("d_board32", C.c_void_p), ("d_arrive", C.c_void_p), ("step", C.c_uint64), ("d_counter", C.c_void_p)]


IPC_HANDLE_BYTES = 64


_OBJ_DIR = _PKG / "build"


def _newer_than(target: Path, deps) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(p.exists() and p.stat().st_mtime > t for p in deps)


def _stale() -> bool:
    return _newer_than(_SO, [_CSRC / f for f in _SOURCES + _HEADERS])


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source for sm_100a into ``libfhe_b200.so`` (in-tree).  One object per source
    (compiled in parallel, rebuilt only when the source or a header is newer), then one link."""
    if not force and not _stale():
        return _SO
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        if _SO.exists():
            return _SO  # GPU box without a toolchain: use the prebuilt library
        raise RuntimeError("nvcc not found and libfhe_b200.so is not built")
    _OBJ_DIR.mkdir(exist_ok=True)
    headers = [_CSRC / h for h in _HEADERS]
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    procs = []
    for src in _SOURCES:
        obj = _OBJ_DIR / (Path(src).stem + ".o")
        if not force and not _newer_than(obj, [_CSRC / src] + headers):
            continue
        cmd = [nvcc, *compile_flags, "-c", "-o", str(obj), str(_CSRC / src)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, pr in procs:
        out, _ = pr.communicate()
        log.append(out)
        if pr.returncode != 0:
            for _, other in procs:
                if other.poll() is None:
                    other.kill()
            raise RuntimeError(f"nvcc failed on {src}:\n" + out)
    objs = [str(_OBJ_DIR / (Path(s).stem + ".o")) for s in _SOURCES]
    r = subprocess.run([nvcc, "-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
                        "-o", str(_SO)] + objs, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    if verbose:
        print("".join(log))
    return _SO


_LIB = None

_u8p, _i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
_vp = C.c_void_p

# name -> (restype, argtypes); this table is also what tests check against include/fhe_b200.h
SIGNATURES = {
    "fhe_b200_abi_version": (C.c_int, []),
    "fhe_b200_last_error": (C.c_char_p, []),
    "fhe_b200_ctx_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "fhe_b200_ctx_destroy": (C.c_int, [_vp]),
    "fhe_b200_device_info": (C.c_int, [_vp, _i32p, _i32p, _i32p, C.POINTER(C.c_uint64)]),
    "fhe_b200_launch_count": (C.c_uint64, [_vp]),
    "fhe_b200_probe_fp64": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "fhe_b200_secret_key": (C.c_int, [_vp, C.c_uint64, C.c_uint32, C.c_int64, _vp, _vp]),
    "fhe_b200_lwe_encrypt": (C.c_int, [_vp, _vp, C.c_int32, C.c_int64, _vp, C.c_int64, C.c_int32, C.c_double,
                                       C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, _vp, _vp]),
    "fhe_b200_lwe_phase": (C.c_int, [_vp, _vp, C.c_int32, C.c_int64, _vp, C.c_int64, _vp, _vp]),
    "fhe_b200_lwe_decrypt": (C.c_int, [_vp, _vp, C.c_int32, C.c_int64, _vp, C.c_int64, C.c_int32, _vp, _vp]),
    "fhe_b200_lincomb": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int64, _vp, C.c_int32,
                                   C.POINTER(C.c_int64), C.c_int32, _vp, _vp]),
    "fhe_b200_accumulate": (C.c_int, [_vp, _vp, _vp, C.c_int64, _vp]),
    "fhe_b200_glwe_encrypt_rows": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_uint64,
                                             C.c_uint64, C.c_uint64, _vp, _vp]),
    "fhe_b200_glwe_ggsw_dot": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, _vp, _vp]),
    "fhe_b200_glwe_decrypt_coeffs": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                               C.c_int32, _vp, _vp]),
    "fhe_b200_glwe_sample_extract": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                               _vp, _vp]),
    "fhe_b200_lwe_pair_add": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_uint64, _vp, _vp]),
    "fhe_b200_lwe_square_sum": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_int32, _vp, _vp, C.c_int64, C.c_int64, _vp, _vp]),
    "fhe_b200_lwe_shl_add": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_uint64, _vp, C.c_int64, _vp]),
    "fhe_b200_lwe_sub_plain": (C.c_int, [_vp, _vp, C.c_int64, _vp, C.c_int64, C.c_int32, C.c_uint64, _vp]),
    "fhe_b200_lwe_pair_addsub": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_uint64, _vp, _vp]),
    "fhe_b200_lwe_pair_diff_sum": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int64, _vp, _vp]),
    "fhe_b200_lwe_encrypt_seeded": (C.c_int, [_vp, _vp, C.c_int32, _vp, C.c_int64, C.c_int32, C.c_double, C.c_uint64,
                                              C.c_uint64, C.c_uint64, C.c_uint32, _vp, _vp]),
    "fhe_b200_lwe_expand_seeded": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_int64, C.c_uint64, C.c_uint64,
                                             C.c_uint32, _vp, _vp]),
    "fhe_b200_lincomb_seeded": (C.c_int, [_vp, _vp, C.c_int64, C.c_int32, C.c_int32, C.c_int64, C.c_uint64, C.c_uint64,
                                          C.c_uint32, _vp, C.c_int32, C.POINTER(C.c_int64), C.c_int32, _vp, _vp]),
    "fhe_b200_similarity_encrypt_seeded": (C.c_int, [_vp, _vp, C.c_int64, C.c_uint64, C.c_uint64, _vp, _vp]),
    "fhe_b200_similarity_encrypt_seeded_products": (C.c_int, [_vp, _vp, _vp, C.c_int64, C.c_uint64, C.c_uint64, _vp, _vp]),
    "fhe_b200_similarity_run_seeded": (C.c_int, [_vp, _vp, C.c_int64, C.c_uint64, C.c_uint64, _vp, _vp]),
    "fhe_b200_similarity_predict_host_seeded": (C.c_int, [_vp, C.POINTER(C.c_float), C.c_int64, C.c_uint64, C.c_uint64,
                                                          C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "fhe_b200_lwe_modswitch32": (C.c_int, [_vp, _vp, C.c_int64, C.c_int64, _vp, _vp]),
    "fhe_b200_similarity_decrypt32": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp]),
    "fhe_b200_ksk_gen": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, C.c_uint64, _vp, _vp]),
    "fhe_b200_bsk_gen": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, C.c_uint64, _vp, _vp]),
    "fhe_b200_ksk_words": (C.c_uint64, [C.POINTER(PBSParams)]),
    "fhe_b200_bsk_words": (C.c_uint64, [C.POINTER(PBSParams)]),
    "fhe_b200_bsk_to_fourier": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, _vp]),
    "fhe_b200_keyswitch": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, C.c_int64, _vp, _vp]),
    "fhe_b200_bsk2_words": (C.c_uint64, [C.POINTER(PBSParams)]),
    "fhe_b200_bsk2_gen": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, C.c_uint64, _vp, _vp]),
    "fhe_b200_bsk2_to_fourier": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, _vp]),
    "fhe_b200_pbs_mb2": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "fhe_b200_ksk_to_32": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, _vp]),
    "fhe_b200_keyswitch32": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, C.c_int64, _vp, _vp, _vp]),
    "fhe_b200_pbs_mb2_wide": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "fhe_b200_pbs_mb2_pair": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "fhe_b200_ksk_mma_bytes": (C.c_uint64, [C.POINTER(PBSParams)]),
    "fhe_b200_keyswitch_mma_workspace_bytes": (C.c_uint64, [C.POINTER(PBSParams), C.c_int64]),
    "fhe_b200_ksk_to_mma": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, _vp]),
    "fhe_b200_keyswitch_mma": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, C.c_int64, _vp, _vp, _vp]),
    "fhe_b200_pbs": (C.c_int, [_vp, C.POINTER(PBSParams), _vp, _vp, C.c_int64, _vp, _vp, _vp, _vp]),
    "fhe_b200_similarity_create": (C.c_int, [_vp, C.POINTER(SimilaritySpec), C.POINTER(C.c_int64), C.POINTER(_vp)]),
    "fhe_b200_similarity_create_evaluator": (C.c_int, [_vp, C.POINTER(SimilaritySpec), C.POINTER(C.c_int64), C.POINTER(_vp)]),
    "fhe_b200_similarity_wire32_supported": (C.c_int, [_vp]),
    "fhe_b200_similarity_destroy": (C.c_int, [_vp]),
    "fhe_b200_similarity_predict_host": (C.c_int, [_vp, C.POINTER(C.c_float), C.c_int64, C.c_uint64, C.c_uint64,
                                                   C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "fhe_b200_similarity_encrypt": (C.c_int, [_vp, _vp, C.c_int64, C.c_uint64, C.c_uint64, _vp, _vp]),
    "fhe_b200_similarity_run": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp]),
    "fhe_b200_similarity_decrypt": (C.c_int, [_vp, _vp, C.c_int64, _vp, _vp, _vp]),
    "fhe_b200_host_alloc": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp)]),
    "fhe_b200_host_free": (C.c_int, [_vp, _vp]),
    "fhe_b200_peer_alloc": (C.c_int, [_vp, C.c_uint64, C.POINTER(_vp), _u8p]),
    "fhe_b200_peer_open": (C.c_int, [_vp, _u8p, C.POINTER(_vp)]),
    "fhe_b200_peer_close": (C.c_int, [_vp, _vp]),
    "fhe_b200_peer_free": (C.c_int, [_vp, _vp]),
    "fhe_b200_similarity_run_push": (C.c_int, [_vp, _vp, C.c_int64, C.POINTER(Push), _vp]),
    "fhe_b200_similarity_run_seeded_push": (C.c_int, [_vp, _vp, C.c_int64, C.c_uint64, C.c_uint64, C.POINTER(Push), _vp]),
    "fhe_b200_peer_wait": (C.c_int, [_vp, _vp, C.c_int32, C.c_uint64, C.c_uint32, _vp, _vp]),
    "fhe_b200_peer_signal": (C.c_int, [_vp, _vp, C.c_int32, C.c_uint64, _vp]),
}


def lib() -> C.CDLL:
    """Load libfhe_b200.so (building it if the sources are newer).  Raises if unavailable."""
    global _LIB
    if _LIB is None:
        so = os.environ.get("FHE_B200_LIB") or build()   # FHE_B200_LIB: A/B-test an alternative build
        L = C.CDLL(str(so))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if L.fhe_b200_abi_version() != 2:
            raise RuntimeError("libfhe_b200.so ABI version mismatch")
        _LIB = L
    return _LIB


def check(code: int) -> None:
    if code != OK:
        raise FheB200Error(code, lib().fhe_b200_last_error().decode())


class Context:
    """One engine context per GPU / rank (``fhe_b200_ctx``)."""

    def __init__(self, device: int = 0):
        self._h = _vp()
        check(lib().fhe_b200_ctx_create(int(device), C.byref(self._h)))
        self.device = int(device)

    @property
    def handle(self):
        return self._h

    def device_info(self) -> dict:
        sm, ma, mi, mem = C.c_int32(), C.c_int32(), C.c_int32(), C.c_uint64()
        check(lib().fhe_b200_device_info(self._h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}

    def probe_fp64_tflops(self) -> float:
        v = C.c_double()
        check(lib().fhe_b200_probe_fp64(self._h, C.byref(v)))
        return v.value

    def launch_count(self) -> int:
        return int(lib().fhe_b200_launch_count(self._h))

    def close(self):
        if self._h:
            lib().fhe_b200_ctx_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_CTX: dict[int, Context] = {}


def context(device: int | None = None) -> Context:
    """Process-wide context for ``device`` (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    if device not in _CTX:
        _CTX[device] = Context(device)
    return _CTX[device]


def pinned_empty(shape, dtype=np.float32, device: int | None = None) -> np.ndarray:
    """A numpy array over page-locked host memory (``fhe_b200_host_alloc``).  Rows passed to ``predict_encrypted`` /
    ``fhe_b200_similarity_predict_host[_seeded]`` from such an array are uploaded from where they are (no staging copy).
    The memory is released when the array -- and every view of it -- is garbage collected."""
    import weakref
    ctx = context(device)
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) * dt.itemsize
    p = _vp()
    check(lib().fhe_b200_host_alloc(ctx.handle, max(n, 1), C.byref(p)))
    buf = (C.c_uint8 * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=int(np.prod(shape))).reshape(shape)
    fin = weakref.finalize(buf, lambda q=p.value: lib().fhe_b200_host_free(None, _vp(q)))
    fin.atexit = False          # at interpreter shutdown the CUDA runtime may already be gone: the OS reclaims the pages
    return arr


def pinned_copy(a: np.ndarray, device: int | None = None) -> np.ndarray:
    a = np.ascontiguousarray(a)
    out = pinned_empty(a.shape, a.dtype, device)
    out[...] = a
    return out
