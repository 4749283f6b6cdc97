"""Where the e2e call's time goes outside the kernels.  usage: [FHE_B200_E2E_MODE=m] e2e_ab.py [docs]
Prints ms per call of (a) the C entry point alone, (b) predict_encrypted, (c) predict_encrypted + host top-k, for
pageable and pinned input rows."""
import ctypes as C, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bench import build_model, synthetic_docs, top_k
from fhe_icp_b200 import _native as N

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
m, _ = build_model(0); c = m.model.fhe_circuit
_, _, X = synthetic_docs(docs, 5)
X = np.ascontiguousarray(X, dtype=np.float32)
ref = m.predict_clear(X)

def wall(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps * 1e3

y = np.empty(docs); qy = np.empty(docs, dtype=np.int64)
fn = N.lib().fhe_b200_similarity_predict_host_seeded
def ccall(A):
    base = c.next_ct_base(docs * c.spec.d)
    N.check(fn(c.handle, A.ctypes.data_as(C.POINTER(C.c_float)), docs, c.enc_seed, base,
               y.ctypes.data_as(C.POINTER(C.c_double)), qy.ctypes.data_as(C.POINTER(C.c_int64))))
c.ciphertext_format = "seeded"
print(f"mode={os.environ.get('FHE_B200_E2E_MODE', '0')} docs={docs}")
for name, A in (("pageable", X), ("pinned", N.pinned_copy(X))):
    a = wall(lambda: ccall(A))
    ok = np.array_equal(y, ref)
    b = wall(lambda: m.predict_encrypted(A))
    d = wall(lambda: top_k(m.predict_encrypted(A), 3, -np.inf))
    print(f"  {name:9s} C call {a:.4f} ms | predict_encrypted {b:.4f} ms | + top-k {d:.4f} ms -> {docs/d/1e3:.3f} M/s  exact={ok}")
