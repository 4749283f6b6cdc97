"""Run a few PBS launches at one batch size (for ncu / timing).  usage: pbs_profile.py [batch] [reps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from fhe_icp_b200 import engine as E
from fhe_icp_b200.params import PBS_PARAMS_4BIT

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
import os
PD = dict(PBS_PARAMS_4BIT)
if os.environ.get("PBS_L", "1") == "2":   # the two-level set of the encrypted x encrypted comparison
    PD.update(l_pbs=2, beta_pbs=15)
p = E.make_pbs_params(**PD)
s, S = E.secret_key(101, 0, p.n, dev), E.secret_key(101, 1, p.k * p.N, dev)
ksk, bsk = E.ksk_gen(p, S, s, 202), E.bsk_gen(p, s, S, 202)
bskf = E.bsk_to_fourier(p, bsk)
table = (np.arange(16) * 7 + 3) % 16
lut = E.from_u64_numpy(E.make_lut_poly(table, 4, p.N, 59), dev)
msgs = np.random.RandomState(0).randint(0, 16, size=B)
ct_big = E.lwe_encrypt(S, torch.as_tensor(msgs), 59, p.sigma_glwe_abs, enc_seed=303, stride=p.N + 2)[:, : p.N + 1].contiguous()
ct = E.keyswitch(p, ksk, ct_big)
out = torch.empty((B, p.N + 1), dtype=torch.int64, device=dev)
import os
MB2 = os.environ.get("PBS_MB2", "0") == "1"
WIDE = os.environ.get("PBS_WIDE", "0") == "1"
PAIR = os.environ.get("PBS_PAIR", "0") == "1"
if MB2 or WIDE or PAIR:
    bskf2 = E.bsk2_to_fourier(p, E.bsk2_gen(p, s, S, 202))
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if PAIR:
        E.pbs_mb2_pair(p, bskf2, ct, lut, out=out)
    elif WIDE:
        E.pbs_mb2_wide(p, bskf2, ct, lut, out=out)
    elif MB2:
        E.pbs_mb2(p, bskf2, ct, lut, out=out)
    else:
        E.pbs(p, bskf, ct, lut, out=out)
    e1.record(); torch.cuda.synchronize()
    print(f"B={B} {'pair' if PAIR else 'wide' if WIDE else 'mb2' if MB2 else 'pbs'} {e0.elapsed_time(e1):.3f} ms -> {B / e0.elapsed_time(e1) * 1e3:.0f} PBS/s")
z = torch.zeros((B, p.N + 2), dtype=torch.int64, device=dev); z[:, : p.N + 1] = out
dec = E.lwe_decrypt(S, z, 59).cpu().numpy() & 15
print("correct:", bool(np.array_equal(dec, table[msgs])))
