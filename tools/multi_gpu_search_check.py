"""torchrun --nproc-per-node N tools/multi_gpu_search_check.py : sharded search over NCCL, checked
against the clear quantized model on rank 0."""
import os
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import torch.distributed as dist

from fhe_icp_b200 import FHESimilarityModel
from fhe_icp_b200.batch_operations import rank_results
from fhe_icp_b200.sharded_search import ShardedSearch, broadcast_public_material

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = FHESimilarityModel(input_dim=128, n_bits=8, seed=3, verbose=False, device=local)
X, _ = m.train(n_samples=500)
m.compile(X[:10])
spec = broadcast_public_material(m.model.spec.to_dict() if rank == 0 else None)
assert spec["q_weights"] == m.model.spec.to_dict()["q_weights"]
# seeds come from the OS CSPRNG and differ per rank: only rank 0 (the client) ever uses its keys; the other ranks are
# key-less evaluators (ShardedSearch's default key_holders="client") and get no documents and no query
n_docs = 10007
rng = np.random.RandomState(9)
q = rng.randn(128).astype(np.float32); q /= np.linalg.norm(q)
docs = rng.randn(n_docs, 128).astype(np.float32)
docs[::3] = q + 0.3 * rng.randn(len(docs[::3]), 128)
docs /= np.linalg.norm(docs, axis=1, keepdims=True)
cdocs, cq = (docs, q) if rank == 0 else (None, None)
ss = ShardedSearch(m, cdocs)                      # seeded ciphertexts + 32-bit score gather (defaults)
res = ss.search(cq, top_k=5, min_similarity=0.5)
res_plain = ShardedSearch(m, cdocs, seeded=False, wire32=False).search(cq, top_k=5, min_similarity=0.5)
assert res == res_plain
# fused gather: scores pushed by the dot-product kernels into the client's score board (no collective)
for seeded in (True, False):
    sp = ShardedSearch(m, cdocs, seeded=seeded, gather="push")
    for rep in range(5):                            # > 2 steps: exercises the slot credits
        res_push = sp.search(cq, top_k=5, min_similarity=0.5)
        assert res_push == res, (rank, seeded, rep, res_push, res)
    sp.close()
if rank == 0:
    ref = rank_results(ss.doc_ids, m.predict_clear(q[None, :] * docs), 5, 0.5)
    assert res == ref, (res, ref)
    print(f"sharded search over {dist.get_world_size()} GPUs OK:", res[:3])
dist.barrier()
dist.destroy_process_group()
