"""PBS latency / throughput by batch size for the three multi-bit blind-rotation kernels (one key set-up, many batches).
usage: pbs_batch_sweep.py [batches...]   prints ms per launch (best of 4) and correctness for the latency kernel (wide), the throughput kernel (mb2<1,4>)
and the dispatcher (FHE_B200_PBS_NO_WIDE=1 in the environment makes the dispatcher use mb2<1,4> only)."""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from fhe_icp_b200 import engine as E
from fhe_icp_b200.params import PBS_PARAMS_4BIT

batches = [int(a) for a in sys.argv[1:]] or [1, 16, 74, 148, 296, 592]
dev = torch.device("cuda", 0)
p = E.make_pbs_params(**dict(PBS_PARAMS_4BIT))
s, S = E.secret_key(101, 0, p.n, dev), E.secret_key(101, 1, p.k * p.N, dev)
ksk = E.ksk_gen(p, S, s, 202)
bskf2 = E.bsk2_to_fourier(p, E.bsk2_gen(p, s, S, 202))
table = (np.arange(16) * 7 + 3) % 16
lut = E.from_u64_numpy(E.make_lut_poly(table, 4, p.N, 59), dev)
def _no_wide(fn):
    os.environ["FHE_B200_PBS_NO_WIDE"] = "1"      # read by the dispatcher at every call
    try:
        return fn()
    finally:
        del os.environ["FHE_B200_PBS_NO_WIDE"]


kernels = {"pair": lambda ct, out: E.pbs_mb2_pair(p, bskf2, ct, lut, out=out),
           "wide": lambda ct, out: E.pbs_mb2_wide(p, bskf2, ct, lut, out=out),
           "mb2<1,4>": lambda ct, out: _no_wide(lambda: E.pbs_mb2(p, bskf2, ct, lut, out=out)),
           "dispatch": lambda ct, out: E.pbs_mb2(p, bskf2, ct, lut, out=out)}
if os.environ.get("SWEEP_ONLY"):
    kernels = {k: v for k, v in kernels.items() if k in os.environ["SWEEP_ONLY"].split(",")}
for B in batches:
    msgs = np.random.RandomState(B).randint(0, 16, size=B)
    ct_big = E.lwe_encrypt(S, torch.as_tensor(msgs), 59, p.sigma_glwe_abs, enc_seed=303, stride=p.N + 2)[:, : p.N + 1].contiguous()
    ct = E.keyswitch(p, ksk, ct_big)
    out = torch.empty((B, p.N + 1), dtype=torch.int64, device=dev)
    row = [f"B={B:5d}"]
    for name, fn in kernels.items():
        best = 1e9
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            out.zero_()
            e0.record(); fn(ct, out); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        z = torch.zeros((B, p.N + 2), dtype=torch.int64, device=dev); z[:, : p.N + 1] = out
        ok = bool(np.array_equal(E.lwe_decrypt(S, z, 59).cpu().numpy() & 15, table[msgs]))
        row.append(f"{name} {best:8.3f} ms {B / best:8.1f} k/s {'ok' if ok else 'WRONG'}")
    print(" | ".join(row), flush=True)
