"""Small pass over the integer kernels for `compute-sanitizer --tool memcheck` (no PBS: TMEM/TMA paths
are exercised by the regular tests)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from fhe_icp_b200 import engine as E

dev = torch.device("cuda", 0)
for n in (15, 16, 630):
    key = E.secret_key(1, 2, n, dev)
    msgs = torch.arange(-6, 6).reshape(3, 4)
    ct = E.lwe_encrypt(key, msgs, 40, 2.0 ** 20, 9)
    assert torch.equal(E.lwe_decrypt(key, ct, 40).cpu(), msgs)
    W = torch.tensor([[1, -2, 3, 4], [1, 1, 1, 1]])
    out = E.lincomb(ct, W, n, bias=[1, 2], shift=40)
    assert torch.equal(E.lwe_decrypt(key, out, 40).cpu(), msgs @ W.T + torch.tensor([1, 2]))
    b = E.lwe_encrypt_seeded(key, msgs, 40, 2.0 ** 20, 9)
    out2 = E.lincomb_seeded(b, W, n, 9, bias=[1, 2], shift=40)
    assert torch.equal(out2, out)
    assert torch.equal(E.lwe_expand_seeded(b, n, 9), ct)
    acc = out.clone().contiguous(); E.accumulate(acc, out.contiguous())
p = E.make_pbs_params(n=24, k=1, N_poly=2048, l_pbs=1, beta_pbs=23, l_ks=5, beta_ks=3, log2_sigma_lwe=-30.0)
s, S = E.secret_key(1, 0, p.n, dev), E.secret_key(1, 1, p.N, dev)
ksk = E.ksk_gen(p, S, s, 2)
ct = E.lwe_encrypt(S, torch.arange(16), 59, p.sigma_glwe_abs, 3, stride=p.N + 2)[:, : p.N + 1].contiguous()
a = E.keyswitch(p, ksk, ct); b32 = E.keyswitch32(p, E.ksk_to_32(p, ksk), ct)
assert torch.equal(E.lwe_decrypt(s, torch.nn.functional.pad(a, (0, 1)), 59).cpu(), torch.arange(16))
assert torch.equal(E.lwe_decrypt(s, torch.nn.functional.pad(b32, (0, 1)), 59).cpu(), torch.arange(16))
torch.cuda.synchronize()
print("sanitize_small OK")
