import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from fhe_icp_b200 import engine as E
from fhe_icp_b200.params import PBS_PARAMS_4BIT
dev = torch.device("cuda", 0)
p = E.make_pbs_params(**PBS_PARAMS_4BIT)
s, S = E.secret_key(101, 0, p.n, dev), E.secret_key(101, 1, p.N, dev)
ksk = E.ksk_gen(p, S, s, 202); k32 = E.ksk_to_32(p, ksk)
for B in (16, 1184, 4736):
    msgs = np.random.RandomState(0).randint(0, 16, size=B)
    ct = E.lwe_encrypt(S, torch.as_tensor(msgs), 59, p.sigma_glwe_abs, 303, stride=p.N + 2)[:, : p.N + 1].contiguous()
    for name, fn in (("ks64", lambda: E.keyswitch(p, ksk, ct)), ("ks32", lambda: E.keyswitch32(p, k32, ct))):
        out = fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        ok = bool(torch.equal(E.lwe_decrypt(s, torch.nn.functional.pad(out, (0, 1)), 59).cpu(), torch.as_tensor(msgs)))
        print(f"B={B} {name} {ms:.3f} ms -> {B/ms*1e3:.0f} KS/s correct={ok}")
