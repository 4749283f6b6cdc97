"""Stage timings of the e2e path (expanded vs seeded ciphertexts).  usage: e2e_profile.py [docs]"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bench import build_model, synthetic_docs

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
m, _ = build_model(0); c = m.model.fhe_circuit
_, _, X = synthetic_docs(docs, 5)

def ev_time(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

ct = m.encrypt(X); sc = m.encrypt(X, seeded=True); out = m.run(ct)
print(f"docs={docs}")
print(f"  encrypt expanded  {ev_time(lambda: m.encrypt(X)):.3f} ms (incl. H2D of X and allocation)")
print(f"  encrypt seeded    {ev_time(lambda: m.encrypt(X, seeded=True)):.3f} ms")
print(f"  run expanded      {ev_time(lambda: m.run(ct, out=out)):.3f} ms")
print(f"  run seeded        {ev_time(lambda: m.run(sc, out=out)):.3f} ms")
for fmt in ("expanded", "seeded"):
    c.ciphertext_format = fmt
    for _ in range(2): m.predict_encrypted(X)
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(10): y = m.predict_encrypted(X)
    dt = (time.perf_counter() - t) / 10
    print(f"  e2e {fmt:9s} {dt*1e3:.3f} ms -> {docs/dt/1e6:.3f} M comparisons/s  exact={np.array_equal(y, m.predict_clear(X))}")
