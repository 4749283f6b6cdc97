"""torchrun --nproc-per-node N tools/multi_gpu_pair_check.py [docs_per_gpu] : both-encrypted search,
documents sharded over the GPUs, evaluation key broadcast from the client rank, checked against the
clear integer model on rank 0; prints whole-job comparisons/s (device-timed, max over ranks)."""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

from fhe_icp_b200.encrypted_compare import EncryptedCompare
from fhe_icp_b200.sharded_search import ShardedPairSearch

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
per_gpu = int(sys.argv[1]) if len(sys.argv) > 1 else 296
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d = 128
ec = EncryptedCompare(input_dim=d, device=local)
if rank == 0:
    ec.keygen()                      # only the client holds secret keys
ec.fit_scale(np.array([-1.0, 1.0]) / np.sqrt(d))
n_docs = per_gpu * world + (3 if world > 1 else 0)
rng = np.random.RandomState(9)
q = rng.randn(d); q /= np.linalg.norm(q)
docs = rng.randn(n_docs, d)
docs[::3] = 0.8 * q + 0.6 * docs[::3] / np.sqrt(d)
docs /= np.linalg.norm(docs, axis=1, keepdims=True)
sp = ShardedPairSearch(ec, docs if rank == 0 else None)
assert rank == 0 or ec.s is None
ints = sp.search_scores(q if rank == 0 else None)      # warm-up + check
if rank == 0:
    assert np.array_equal(ints, ec.quantize(docs) @ ec.quantize(q))
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
res = sp.search(q if rank == 0 else None, top_k=5, min_similarity=0.5)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1)], device=f"cuda:{local}")
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    want = sorted(np.flatnonzero(ec.dequantize(ints) >= 0.5), key=lambda i: (-ints[i], i))[:5]
    assert [r[0] for r in res] == [f"doc_{i}" for i in want], (res, want)
    print(f"both-encrypted sharded search over {world} GPU(s): {n_docs} docs in {t.item():.1f} ms -> "
          f"{n_docs / t.item() * 1e3:.1f} comparisons/s (exact)", res[:2])
dist.barrier()
dist.destroy_process_group()
