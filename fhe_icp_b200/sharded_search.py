"""Document-sharded encrypted search across the GPUs of one box (one process per GPU).

Documents are independent units (the reference loops over them with no shared state,
/root/reference/batch_operations.py:268-279), so the collection is cut into contiguous ranges,
one per rank.  Per query: every rank evaluates the encrypted dot products of its shard, the
encrypted scores (M x stride u64 words per document) are all-gathered over NCCL/NVLink, and the
client rank decrypts, thresholds, sorts and takes the top-k -- top-k happens AFTER decryption, so
there is no encrypted top-k collective.  Public material (quantized model; evaluation keys when the
PBS path is used) is broadcast once from rank 0.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist

from .batch_operations import rank_results


def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous balanced ranges: the first n % world ranks hold one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def broadcast_public_material(obj, src: int = 0):
    """Broadcast a picklable public object (quantized spec, LWE parameters) from ``src``."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return obj
    box = [obj if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def broadcast_keys(tensors: Sequence[torch.Tensor], src: int = 0):
    """Broadcast evaluation keys (BSK / KSK device tensors, ~110 MB at the stated set) once."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in tensors:
            dist.broadcast(t, src=src)
    return tensors


class ShardedSearch:
    """``engine`` needs ``encrypt(X) -> ct``, ``run(ct) -> enc_scores [rows, M, stride]`` and
    ``decrypt(enc_scores) -> float64 scores``; FHESimilarityModel provides all three on the GPU."""

    def __init__(self, engine, docs: np.ndarray, doc_ids: Optional[List[str]] = None, client_rank: int = 0):
        self.engine = engine
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.client_rank = client_rank
        self.n_docs = int(docs.shape[0])
        self.doc_ids = doc_ids if doc_ids is not None else [f"doc_{i}" for i in range(self.n_docs)]
        self.lo, self.hi = shard_bounds(self.n_docs, self.world, self.rank)
        self.shard = np.ascontiguousarray(docs[self.lo:self.hi], dtype=np.float32)
        self.max_rows = max(shard_bounds(self.n_docs, self.world, r)[1] - shard_bounds(self.n_docs, self.world, r)[0]
                            for r in range(self.world))
        self._ct = None
        self._query = None

    def encrypt_shard(self, query: np.ndarray):
        """Client-side: X = query * docs (the clear product the reference feeds the circuit,
        batch_operations.py:273), encrypted row by row; stays resident on this rank."""
        self._query = np.asarray(query, dtype=np.float32)
        X = (self._query[None, :] * self.shard).astype(np.float32)
        self._ct = self.engine.encrypt(X) if len(X) else None
        return self._ct

    def evaluate_and_gather(self) -> Optional[torch.Tensor]:
        """Server-side: encrypted scores of this shard, all-gathered (padded to the largest shard)."""
        rows = self.hi - self.lo
        out = self.engine.run(self._ct) if rows else None
        if self.world == 1:
            return out
        ref = out if out is not None else self._probe_shape()
        padded = torch.zeros((self.max_rows,) + tuple(ref.shape[1:]), dtype=ref.dtype, device=ref.device)
        if rows:
            padded[:rows] = out
        gathered = torch.empty((self.world * self.max_rows,) + tuple(ref.shape[1:]), dtype=ref.dtype, device=ref.device)
        dist.all_gather_into_tensor(gathered, padded)
        parts = []
        for r in range(self.world):
            lo, hi = shard_bounds(self.n_docs, self.world, r)
            parts.append(gathered[r * self.max_rows: r * self.max_rows + (hi - lo)])
        return torch.cat(parts, dim=0)

    def _probe_shape(self):
        # a rank with an empty shard still needs the per-document shape for the collective
        return self.engine.run(self.engine.encrypt(np.zeros((1, self.shard.shape[1]), dtype=np.float32)))

    def search(self, query: np.ndarray, top_k: int = 5, min_similarity: float = 0.5):
        """Full query: returns the ranked list on the client rank, None elsewhere."""
        self.encrypt_shard(query)
        enc = self.evaluate_and_gather()
        if self.rank != self.client_rank:
            return None
        scores = self.engine.decrypt(enc) if enc is not None and enc.shape[0] else np.zeros(0)
        return rank_results(self.doc_ids, scores, top_k, min_similarity)
