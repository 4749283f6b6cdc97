"""A few seeded predict_encrypted calls (1000 documents) for ncu: launch list / --set full of the client
encryption and seeded dot-product kernels.  usage: e2e_ncu.py [docs] [calls]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from bench import build_model, synthetic_docs

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m, _ = build_model(0)
m.model.fhe_circuit.ciphertext_format = "seeded"
_, _, X = synthetic_docs(docs, 5)
for _ in range(calls):
    y = m.predict_encrypted(X)
torch.cuda.synchronize()
print("exact:", bool(np.array_equal(y, m.predict_clear(X))))
