"""Time the packed both-encrypted comparison (one GLWE x GGSW external product per 16 documents).
usage: python tools/packed_profile.py [docs] [reps]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from fhe_icp_b200 import engine as E  # noqa: E402
from fhe_icp_b200.encrypted_compare import PackedEncryptedCompare  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
d = 128
dev = torch.device("cuda:0")
pe = PackedEncryptedCompare(input_dim=d, device=dev).keygen()
rng = np.random.RandomState(0)
xq = rng.randint(-16, 16, size=d)
nenc = min(B, 16 * 2048)                      # encrypt a bounded sample and tile it (timing only needs valid layout)
yq = rng.randint(-16, 16, size=(nenc, d))
gq = pe.encrypt_query(xq, 1)
gd_s = pe.encrypt_documents(yq, 1)
G = (B + pe.per - 1) // pe.per
gd = gd_s.repeat((G + gd_s.shape[0] - 1) // gd_s.shape[0], 1, 1)[:G].contiguous()
out = torch.empty_like(gd)
pe.scores(gq, gd, out)
torch.cuda.synchronize()
for _ in range(reps):
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    pe.scores(gq, gd, out)
    e1.record()
    raw = E.glwe_decrypt_coeffs(pe.p, pe.S, out, 0, pe.slot, pe.per, 47)
    e2.record()
    torch.cuda.synchronize()
    ms, ms2 = e0.elapsed_time(e1), e1.elapsed_time(e2)
    gb = 2 * gd.numel() * 8 / 1e9
    print(f"docs={B} ciphertexts={G}: external products {ms:.3f} ms -> {B / ms * 1e3 / 1e6:.1f} M comparisons/s "
          f"({G / ms * 1e3 / 1e6:.2f} M ext. products/s, {gb / ms * 1e3:.0f} GB/s of HBM); client decrypt kernel {ms2:.3f} ms")
ints = pe.decrypt(out, B)
want = np.tile(yq @ xq, (B + nenc - 1) // nenc)[:B]
print("exact:", bool(np.array_equal(ints, want)))
