"""Keyswitch throughput: integer-pipe KS32 vs the tensor-core contraction.  usage: ks_profile2.py [batch]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from fhe_icp_b200 import engine as E
from fhe_icp_b200.params import PBS_PARAMS_4BIT as P4

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
dev = torch.device("cuda", 0)
p = E.make_pbs_params(**P4)
s, S = E.secret_key(11, 0, p.n, dev), E.secret_key(11, 1, p.k * p.N, dev)
ksk = E.ksk_gen(p, S, s, 22)
ksk32 = E.ksk_to_32(p, ksk)
key_mma = E.ksk_to_mma(p, ksk32)
ct = torch.as_tensor(np.random.RandomState(0).randint(-2 ** 63, 2 ** 63 - 1, size=(B, p.k * p.N + 1), dtype=np.int64)).to(dev)

def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

same = torch.equal(E.keyswitch_mma(p, key_mma, ct), E.keyswitch32(p, ksk32, ct))
t32 = ev(lambda: E.keyswitch32(p, ksk32, ct))
tmm = ev(lambda: E.keyswitch_mma(p, key_mma, ct))
ops = 2.0 * B * p.k * p.N * p.l_ks * 4 * (p.n + 1)
print(f"batch {B}: KS32 {t32:.3f} ms ({B/t32/1e3:.3f} M/s)   tensor-core {tmm:.3f} ms ({B/tmm/1e3:.3f} M/s, "
      f"{ops/tmm/1e9:.1f} int8 TOP/s incl. digit kernel)   identical={same}")
