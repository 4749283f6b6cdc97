"""torchrun --nproc-per-node N tools/multi_gpu_packed_check.py [docs_per_gpu] : packed both-encrypted search,
GLWE ciphertexts sharded over the GPUs, query GGSW broadcast per query, checked against the clear integer model
on the client rank; prints whole-job comparisons/s (device-timed, max over ranks)."""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

from fhe_icp_b200.encrypted_compare import PackedEncryptedCompare
from fhe_icp_b200.sharded_search import ShardedPackedSearch

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d = 128
pe = PackedEncryptedCompare(input_dim=d, device=local)
if rank == 0:
    pe.keygen()                      # only the client holds the secret key
pe.fit_scale(np.array([-1.0, 1.0]) / np.sqrt(d))
n_docs = per_gpu * world + (7 if world > 1 else 0)
rng = np.random.RandomState(9)
q = rng.randn(d); q /= np.linalg.norm(q)
docs = None
if rank == 0:
    docs = rng.randn(n_docs, d)
    docs[::3] = 0.8 * q + 0.6 * docs[::3] / np.sqrt(d)
    docs /= np.linalg.norm(docs, axis=1, keepdims=True)
sp = ShardedPackedSearch(pe, docs)
assert rank == 0 or pe.S is None
ints = sp.search_scores(q if rank == 0 else None)      # warm-up + check
if rank == 0:
    assert np.array_equal(ints, pe.quantize(docs) @ pe.quantize(q))
best = None
for _ in range(3):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = sp.search(q if rank == 0 else None, top_k=5, min_similarity=0.5)
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    best = t.item() if best is None else min(best, t.item())
if rank == 0:
    want = sorted(np.flatnonzero(pe.dequantize(ints) >= 0.5), key=lambda i: (-ints[i], i))[:5]
    assert [r[0] for r in res] == [f"doc_{i}" for i in want], (res, want)
    print(f"packed both-encrypted sharded search over {world} GPU(s): {n_docs} docs per query in {best:.2f} ms -> "
          f"{n_docs / best * 1e3 / 1e6:.1f} M comparisons/s end to end (exact)", res[:2])
dist.barrier()
dist.destroy_process_group()
