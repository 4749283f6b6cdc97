"""Time the encrypted x encrypted comparison (2d PBS per document) on one GPU.
usage: python tools/compare_profile.py [docs] [reps] [multibit(0/1)] [l_pbs]"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from fhe_icp_b200.encrypted_compare import COMPARE_PARAMS, EncryptedCompare  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 74
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
mb = bool(int(sys.argv[3])) if len(sys.argv) > 3 else False
params = dict(COMPARE_PARAMS)
if len(sys.argv) > 4:
    params["l_pbs"] = int(sys.argv[4])
    params["beta_pbs"] = {1: 23, 2: 15, 3: 10}[params["l_pbs"]]
d = 128
dev = torch.device("cuda:0")
ec = EncryptedCompare(input_dim=d, params=params, device=dev, multibit=mb, chunk_pbs=148 * 4 * 32).keygen()
rng = np.random.RandomState(0)
q = rng.randn(d) / np.sqrt(d)
docs = rng.randn(B, d) / np.sqrt(d)
ec.fit_scale(docs)
ct_q = ec.encrypt(ec.quantize(q), 1, 0)
ct_d = ec.encrypt(ec.quantize(docs), 1, d)
NORMS = len(sys.argv) > 5 and sys.argv[5] == "norms"
nq = nd = None
if NORMS:
    nq, nd = ec.encrypt_norms(ec.quantize(q), 1, 0), ec.encrypt_norms(ec.quantize(docs), 1, 1)
ec.scores(ct_q, ct_d, nq, nd)
torch.cuda.synchronize()
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sc = ec.scores(ct_q, ct_d, nq, nd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"docs={B} l_pbs={params['l_pbs']} multibit={mb}: {ms:.2f} ms -> {B / ms * 1e3:.1f} comparisons/s, "
          f"{(1 if NORMS else 2) * d * B / ms * 1e3:.0f} PBS/s")
got = ec.decrypt(sc)
print("exact:", bool(np.array_equal(got, ec.compare_clear(q, docs))))
