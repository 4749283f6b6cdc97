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


def shard_bounds(n_items: int, world: int, rank: int, weights: Optional[Sequence[float]] = None) -> Tuple[int, int]:
    """Contiguous ranges, one per rank.  Without ``weights``: balanced by count (the first n % world ranks hold one
    extra item).  With ``weights`` (one non-negative number per rank): sizes proportional to them, edges rounded from
    the cumulative weights so that the ranges tile [0, n_items) exactly for every rank."""
    if weights is None:
        base, extra = divmod(n_items, world)
        lo = rank * base + min(rank, extra)
        return lo, lo + base + (1 if rank < extra else 0)
    if len(weights) != world or min(weights) < 0 or sum(weights) <= 0:
        raise ValueError("weights: one non-negative number per rank, not all zero")
    total = float(sum(weights))
    edge = lambda r: int(round(n_items * (float(sum(weights[:r])) / total)))  # noqa: E731
    return (0 if rank == 0 else edge(rank)), (n_items if rank == world - 1 else edge(rank + 1))


def client_cost_weights(world: int, client_rank: int, rho: float) -> Optional[List[float]]:
    """Shard weights that equalise the per-query time of the ranks when the client rank, besides evaluating its own
    shard, decrypts EVERY document's scores: with rho = (client cost per document) / (server cost per document) the
    ranks finish together when the client holds 1 - (world-1)*rho of an equal share and every other rank 1 + rho of
    one (sharded_search.ShardedSearch(shard_weights=...), bench.py).  rho <= 0 or a single rank: None (equal shards)."""
    if world <= 1 or rho <= 0:
        return None
    w = [1.0 + rho] * world
    w[client_rank] = max(0.0, 1.0 - (world - 1) * rho)
    return w


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


class _LazyDocIds:
    """doc_<i> names without a million-entry list (the default ids of a large synthetic collection)."""

    def __init__(self, n: int):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [f"doc_{j}" for j in range(*i.indices(self.n))]
        i = int(i)
        if not -self.n <= i < self.n:
            raise IndexError(i)
        return f"doc_{i % self.n}"


class ShardedSearch:
    """``engine`` needs ``encrypt(X) -> ct``, ``run(ct) -> enc_scores [rows, M, stride]`` and
    ``decrypt(enc_scores) -> float64 scores``; FHESimilarityModel provides all three on the GPU.

    Roles (``key_holders``):
      * ``"client"`` (default): only the client rank holds secret keys.  Per query it forms X = query * docs for the
        whole collection, encrypts it (seeded ciphertexts: 1 KB per document) and sends every rank its contiguous
        shard; the other ranks only ever call ``engine.run`` -- on a GPU engine that is the evaluator handle
        (fhe_b200_similarity_create_evaluator), which contains no key material -- and the client decrypts.
        ``docs`` is needed on the client rank only.
      * ``"all"``: every GPU of the box belongs to the key owner (the reference's trust model: one process, one
        machine).  Each rank encrypts its own shard, so the client-side encryption is sharded as well; all ranks
        must have been built from the SAME key set (explicit seeds or a loaded key file)."""

    def __init__(self, engine, docs: Optional[np.ndarray], doc_ids: Optional[List[str]] = None, client_rank: int = 0,
                 seeded: bool = True, wire32: bool = True, gather: str = "nccl", key_holders: str = "client",
                 shard_weights: Optional[Sequence[float]] = None, n_docs: Optional[int] = None,
                 collection: str = "host"):
        """``gather``: "nccl" -- scores written locally, then all-gathered; "push" -- the dot-product kernel
        of every rank stores its scores into the client GPU's score board over NVLink (score_board.py; the
        engine must be an FHESimilarityModel on a CUDA device, scores travel in the 32-bit wire form).
        ``shard_weights``: relative shard sizes (see :func:`client_cost_weights`; None: equal).  ``n_docs`` (only with
        key_holders='all'): size of the whole collection when ``docs`` holds just THIS rank's shard, so that no rank
        has to materialise documents it never encrypts.  ``collection`` (engines with ``encrypt_products``, ranks that
        encrypt their own shard): "host" -- the shard stays in pinned host memory and is uploaded with every query;
        "device" -- it is uploaded once and stays resident on the rank's GPU (the key owner's GPU: it is plaintext)."""
        if gather not in ("nccl", "push"):
            raise ValueError(f"Unknown gather mode: {gather}")
        if key_holders not in ("client", "all"):
            raise ValueError(f"Unknown key_holders: {key_holders}")
        if collection not in ("host", "device"):
            raise ValueError(f"Unknown collection placement: {collection}")
        self.collection = collection
        self._shard_t = None
        self.engine = engine
        self.key_holders = key_holders
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.client_rank = client_rank
        self.is_client = self.rank == client_rank
        # optional engine capabilities (FHESimilarityModel has both): seeded fresh ciphertexts (8-byte
        # bodies, masks regenerated by the evaluator) and the 32-bit wire form of the gathered scores (only where it
        # keeps the decoding failure probability: FHECircuit.wire32_supported)
        self.seeded = seeded and hasattr(engine, "compress_scores")
        self.wire32 = wire32 and hasattr(engine, "compress_scores") and bool(getattr(engine, "wire32_supported", True))
        self._is32 = False
        has_docs = docs is not None
        if key_holders == "client" and self.is_client and not has_docs:
            raise ValueError("the client rank needs the documents")
        if key_holders == "all" and not has_docs:
            raise ValueError("key_holders='all': every rank encrypts its own shard and needs the documents")
        if n_docs is not None and key_holders != "all":
            raise ValueError("n_docs (docs = this rank's shard only) needs key_holders='all'")
        meta = (int(n_docs if n_docs is not None else docs.shape[0]), int(docs.shape[1])) if (has_docs and self.is_client) else None
        self.n_docs, self.d = broadcast_public_material(meta, client_rank)
        self.shard_weights = list(shard_weights) if shard_weights is not None else None
        self.doc_ids = doc_ids if doc_ids is not None else _LazyDocIds(self.n_docs)
        self.lo, self.hi = self._bounds(self.rank)
        self.docs = np.ascontiguousarray(docs, dtype=np.float32) if (has_docs and (self.is_client or key_holders == "all")) else None
        if n_docs is not None:
            if self.docs.shape[0] != self.hi - self.lo:
                raise ValueError(f"rank {self.rank}: docs holds {self.docs.shape[0]} rows, its shard has {self.hi - self.lo}")
            self.shard = self.docs
        else:
            self.shard = self.docs[self.lo:self.hi] if self.docs is not None else None
        self.max_rows = max(self._bounds(r)[1] - self._bounds(r)[0] for r in range(self.world))
        self._ct = None
        self._query = None
        self._score_shape = None      # (M, stride) of one document's encrypted scores; the client probes it once
        self.board = None
        if gather == "push":
            if not self.wire32:
                raise ValueError("gather='push' moves scores in the 32-bit wire form, which these parameters do not support")
            from .score_board import PeerScoreBoard
            self.board = PeerScoreBoard(engine, max(self.max_rows, 1), client_rank=client_rank)

    def _bounds(self, r: int) -> Tuple[int, int]:
        return shard_bounds(self.n_docs, self.world, r, self.shard_weights)

    # ---- client side
    def _encrypt(self, X: np.ndarray):
        return self.engine.encrypt(X, seeded=True) if self.seeded else self.engine.encrypt(X)

    def _device(self) -> torch.device:
        dev = getattr(self.engine, "dev", None)
        if dev is not None:
            return torch.device(dev)
        if dist.is_initialized() and dist.get_backend() == "nccl":
            return torch.device("cuda", torch.cuda.current_device())
        return torch.device("cpu")

    def encrypt_shard(self, query: Optional[np.ndarray]):
        """Client side: X = query * docs (the clear product the reference feeds the circuit,
        batch_operations.py:273), encrypted row by row, each rank ending up with the ciphertexts of its shard.
        ``query`` is needed where the encryption happens (the client rank; every rank for key_holders='all')."""
        if self.key_holders == "all" or self.world == 1:
            self._query = np.asarray(query, dtype=np.float32)
            if not len(self.shard):
                self._ct = None
            elif self.seeded and hasattr(self.engine, "encrypt_products"):
                # the product query * docs is taken by the encryption kernel; the host only hands over its rows
                if self._shard_t is None:
                    t = torch.from_numpy(self.shard)
                    self._shard_t = t.to(self._device()) if self.collection == "device" else t.pin_memory()
                self._ct = self.engine.encrypt_products(self._query, self._shard_t)
            else:
                self._ct = self._encrypt((self._query[None, :] * self.shard).astype(np.float32))
            return self._ct
        metas = [None] * self.world
        mine, outgoing = None, []
        if self.is_client:
            self._query = np.asarray(query, dtype=np.float32)
            for r in range(self.world):
                lo, hi = self._bounds(r)
                if hi == lo:
                    continue
                ct = self._encrypt((self._query[None, :] * self.docs[lo:hi]).astype(np.float32))
                if hasattr(ct, "bodies"):
                    metas[r] = ("seeded", tuple(ct.bodies.shape), int(ct.enc_seed), int(ct.ct_base))
                    payload = ct.bodies
                else:
                    metas[r] = ("expanded", tuple(ct.shape))
                    payload = ct
                if r == self.rank:
                    mine = ct
                else:
                    outgoing.append((r, payload.contiguous()))
        dist.broadcast_object_list(metas, src=self.client_rank)      # public: shapes, mask seed, ciphertext ids
        if self.is_client:
            for r, payload in outgoing:
                dist.send(payload, dst=r)
            self._ct = mine
        else:
            meta = metas[self.rank]
            self._ct = None
            if meta is not None:
                buf = torch.empty(meta[1], dtype=torch.int64, device=self._device())
                dist.recv(buf, src=self.client_rank)
                if meta[0] == "seeded":
                    from .fhe_similarity import SeededCiphertexts
                    self._ct = SeededCiphertexts(buf, meta[2], meta[3])
                else:
                    self._ct = buf
        return self._ct

    # ---- server side
    def _row_shape(self) -> Tuple[int, ...]:
        """(M, stride) of one document's scores: probed once by the client (it alone can encrypt), then public."""
        if self._score_shape is None:
            shape = None
            if self.is_client:
                probe = self.engine.run(self._encrypt(np.zeros((1, self.d), dtype=np.float32)))
                shape = tuple(probe.shape[1:])
            self._score_shape = broadcast_public_material(shape, self.client_rank)
        return self._score_shape

    def evaluate_and_gather(self) -> Optional[torch.Tensor]:
        """Server-side: encrypted scores of this shard, all-gathered (padded to the largest shard)."""
        rows = self.hi - self.lo
        if self.board is not None:
            return self._push_and_collect(rows)
        out = self.engine.run(self._ct) if rows else None
        self._is32 = False
        if self.world == 1:
            return out
        row_shape = self._row_shape()
        if self.wire32:
            out = self.engine.compress_scores(out) if rows else None
            self._is32 = True
        dtype = torch.int32 if self._is32 else torch.int64
        dev = out.device if out is not None else self._device()
        padded = torch.zeros((self.max_rows,) + row_shape, dtype=dtype, device=dev)
        if rows:
            padded[:rows] = out
        gathered = torch.empty((self.world * self.max_rows,) + row_shape, dtype=dtype, device=dev)
        dist.all_gather_into_tensor(gathered, padded)
        parts = []
        for r in range(self.world):
            lo, hi = self._bounds(r)
            parts.append(gathered[r * self.max_rows: r * self.max_rows + (hi - lo)])
        return torch.cat(parts, dim=0)

    def _push_and_collect(self, rows: int) -> Optional[torch.Tensor]:
        """Fused evaluate + gather: no collective; only the client rank gets the scores back."""
        self.board.push(self._ct if rows else None)
        self._is32 = True
        if not self.is_client:
            return None
        slot = self.board.collect()
        parts = []
        for r in range(self.world):
            lo, hi = self._bounds(r)
            parts.append(slot[r * self.board.rows_max: r * self.board.rows_max + (hi - lo)])
        enc = torch.cat(parts, dim=0)     # copy out of the slot, then hand it back to the servers
        self.board.release()
        return enc

    def _push_and_decrypt(self, rows: int) -> Optional[np.ndarray]:
        """Push-mode query without a copy of the gathered ciphertexts: the client decrypts every rank's rows straight
        out of the score-board slot (one fused decrypt kernel per rank), returns the slot, and reads back 8 bytes per
        document."""
        self.board.push(self._ct if rows else None)
        self._is32 = True
        if not self.is_client:
            return None
        slot = self.board.collect()
        ys = []
        for r in range(self.world):
            lo, hi = self._bounds(r)
            if hi > lo:
                ys.append(self.engine.decrypt_compressed(slot[r * self.board.rows_max: r * self.board.rows_max + (hi - lo)],
                                                         to_host=False))
        self.board.release()               # stream-ordered after the decrypt kernels
        if not ys:
            return np.zeros(0)
        return (torch.cat(ys) if len(ys) > 1 else ys[0]).cpu().numpy()

    def close(self):
        """Collective: unmap / free the score board (push mode)."""
        if self.board is not None:
            self.board.check()
            self.board.close()
            self.board = None

    def search_scores(self, query: Optional[np.ndarray]) -> Optional[np.ndarray]:
        """Decrypted scores of the whole collection on the client rank (None elsewhere)."""
        self.encrypt_shard(query)
        if self.board is not None:
            return self._push_and_decrypt(self.hi - self.lo)
        enc = self.evaluate_and_gather()
        if not self.is_client:
            return None
        if enc is None or not enc.shape[0]:
            return np.zeros(0)
        return self.engine.decrypt_compressed(enc) if self._is32 else self.engine.decrypt(enc)

    def search(self, query: Optional[np.ndarray], top_k: int = 5, min_similarity: float = 0.5):
        """Full query: returns the ranked list on the client rank, None elsewhere (``query`` is read only where
        the encryption happens)."""
        scores = self.search_scores(query)
        if scores is None:
            return None
        return rank_results(self.doc_ids, scores, top_k, min_similarity)


class ShardedPairSearch:
    """Document-sharded search when BOTH the query and the documents are encrypted
    (encrypted_compare.EncryptedCompare; SURVEY.md 8e for the 8f-N1 path).

    Roles: the client rank owns the secret keys.  Once: it encrypts the collection (d small-key
    ciphertexts + one squared-norm ciphertext per document), sends every rank its contiguous shard and
    broadcasts the evaluation key (Fourier bootstrapping key, ~146 MB at the production set) -- server
    ranks never hold a secret key.  Per query: the client encrypts the query and broadcasts its
    ciphertexts (d*(n+1) + kN+2 words), every rank bootstraps its shard's d sums per document, the
    score ciphertexts (kN+2 words per document) are all-gathered and the client decrypts, thresholds,
    stable-sorts and takes the top-k.  No collective touches the data path between those two points.

    ``engine`` (client) needs quantize / encrypt / encrypt_norms / decrypt / dequantize and the key
    tensors ``bskf``; every rank needs ``scores``.  ``server_engine`` builds a key-less engine from the
    broadcast key on the other ranks."""

    def __init__(self, engine, docs: Optional[np.ndarray], doc_ids: Optional[List[str]] = None, client_rank: int = 0,
                 enc_seed: Optional[int] = None):
        """``enc_seed``: public mask seed (None: the engine's own); every encryption takes fresh ciphertext ids from
        the client engine's allocator, so no mask / error pair is ever reused across documents or queries."""
        self.engine = engine
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.client_rank = client_rank
        self.enc_seed = enc_seed
        is_client = self.rank == client_rank
        meta = broadcast_public_material((int(docs.shape[0]), int(docs.shape[1])) if is_client else None, client_rank)
        self.n_docs, self.d = meta
        self.doc_ids = doc_ids if doc_ids is not None else [f"doc_{i}" for i in range(self.n_docs)]
        self.lo, self.hi = shard_bounds(self.n_docs, self.world, self.rank)
        self.max_rows = max(shard_bounds(self.n_docs, self.world, r)[1] - shard_bounds(self.n_docs, self.world, r)[0]
                            for r in range(self.world))
        self._distribute_key()
        self._distribute_documents(docs if is_client else None)

    # ---- one-time setup
    def _distribute_key(self):
        e = self.engine
        if self.world == 1:
            return
        shape = broadcast_public_material(tuple(e.bskf.shape) if self.rank == self.client_rank else None, self.client_rank)
        if self.rank != self.client_rank:
            e.bskf = torch.empty(shape, dtype=torch.float64, device=e.dev)
        broadcast_keys([e.bskf], self.client_rank)

    def _distribute_documents(self, docs):
        e = self.engine
        shapes = None
        if self.rank == self.client_rank:
            pc, pn = self._probe(e)
            shapes = (tuple(pc.shape[1:]), tuple(pn.shape[1:]))
        self._row_shapes = broadcast_public_material(shapes, self.client_rank)
        self._ct = self._norms = None
        if self.rank == self.client_rank:
            yq = e.quantize(docs)
            for r in range(self.world):
                lo, hi = shard_bounds(self.n_docs, self.world, r)
                if hi == lo:
                    continue
                ct = e.encrypt(yq[lo:hi], self.enc_seed)
                nm = e.encrypt_norms(yq[lo:hi], self.enc_seed)
                if r == self.rank:
                    self._ct, self._norms = ct, nm
                else:
                    dist.send(ct.contiguous(), dst=r)
                    dist.send(nm.contiguous(), dst=r)
        elif self.hi > self.lo:
            rows = self.hi - self.lo
            self._ct = torch.empty((rows,) + self._row_shapes[0], dtype=torch.int64, device=e.dev)
            self._norms = torch.empty((rows,) + self._row_shapes[1], dtype=torch.int64, device=e.dev)
            dist.recv(self._ct, src=self.client_rank)
            dist.recv(self._norms, src=self.client_rank)

    def _probe(self, e):
        z = np.zeros((1, self.d), dtype=np.int64)
        return e.encrypt(z, self.enc_seed), e.encrypt_norms(z, self.enc_seed)

    # ---- per query
    def search_scores(self, query: Optional[np.ndarray]) -> Optional[np.ndarray]:
        """Integer scores of all documents on the client rank (None elsewhere)."""
        e = self.engine
        if self.rank == self.client_rank:
            xq = e.quantize(query)
            ct_q = e.encrypt(xq, self.enc_seed).contiguous()             # fresh ids per query
            n_q = e.encrypt_norms(xq, self.enc_seed).reshape(-1).contiguous()
        else:
            ct_q = torch.empty(self._row_shapes[0], dtype=torch.int64, device=e.dev)
            n_q = torch.empty(self._row_shapes[1], dtype=torch.int64, device=e.dev)
        if self.world > 1:
            dist.broadcast(ct_q, src=self.client_rank)
            dist.broadcast(n_q, src=self.client_rank)
        rows = self.hi - self.lo
        out = e.scores(ct_q, self._ct, n_q, self._norms) if rows else None
        if self.world > 1:
            width = self._row_shapes[1][0]
            padded = torch.zeros((self.max_rows, width), dtype=torch.int64, device=e.dev)
            if rows:
                padded[:rows] = out
            gathered = torch.empty((self.world * self.max_rows, width), dtype=torch.int64, device=e.dev)
            dist.all_gather_into_tensor(gathered, padded)
            parts = []
            for r in range(self.world):
                lo, hi = shard_bounds(self.n_docs, self.world, r)
                parts.append(gathered[r * self.max_rows: r * self.max_rows + (hi - lo)])
            out = torch.cat(parts, dim=0)
        if self.rank != self.client_rank:
            return None
        if out is None or not out.shape[0]:
            return np.zeros(0, dtype=np.int64)
        return e.decrypt(out)

    def search(self, query: Optional[np.ndarray], top_k: int = 5, min_similarity: float = 0.5):
        ints = self.search_scores(query)
        if ints is None:
            return None
        return rank_results(self.doc_ids, self.engine.dequantize(ints), top_k, min_similarity)


class ShardedPackedSearch:
    """Document-sharded search over the PACKED both-encrypted engine (PackedEncryptedCompare): the unit
    that shards is the GLWE ciphertext (engine.per documents each).  The client rank owns the secret key;
    once, it encrypts the collection and sends every rank its contiguous range of ciphertexts.  Per query
    it broadcasts the query's Fourier GGSW (128 KB -- the only per-query traffic to the servers, which hold
    no key material at all), every rank runs one external product per ciphertext of its shard, the
    products (32 KB per 16 documents) are all-gathered and the client decrypts the wanted coefficients,
    thresholds, stable-sorts and takes the top-k."""

    def __init__(self, engine, docs: Optional[np.ndarray], doc_ids: Optional[List[str]] = None, client_rank: int = 0,
                 enc_seed: Optional[int] = None, chunk_groups: int = 4096):
        self.engine = engine
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.client_rank = client_rank
        self.enc_seed = enc_seed
        is_client = self.rank == client_rank
        self.n_docs = broadcast_public_material(int(docs.shape[0]) if is_client else None, client_rank)
        self.doc_ids = doc_ids if doc_ids is not None else [f"doc_{i}" for i in range(self.n_docs)]
        per = engine.per
        self.n_groups = (self.n_docs + per - 1) // per
        self.glo, self.ghi = shard_bounds(self.n_groups, self.world, self.rank)
        self.max_groups = max(shard_bounds(self.n_groups, self.world, r)[1] - shard_bounds(self.n_groups, self.world, r)[0]
                              for r in range(self.world))
        e = engine
        shapes = None
        if is_client:
            probe = e.encrypt_documents(np.zeros((1, docs.shape[1]), dtype=np.int64), enc_seed)
            shapes = (tuple(probe.shape[1:]), tuple(e.encrypt_query(np.zeros(docs.shape[1], dtype=np.int64), enc_seed).shape))
        self._ct_shape, self._q_shape = broadcast_public_material(shapes, client_rank)
        self._ct = None
        if is_client:
            yq = e.quantize(docs)
            for r in range(self.world):
                glo, ghi = shard_bounds(self.n_groups, self.world, r)
                parts = []
                for g0 in range(glo, ghi, chunk_groups):       # bounded client memory per encryption call
                    g1 = min(ghi, g0 + chunk_groups)
                    ct = e.encrypt_documents(yq[g0 * per: min(self.n_docs, g1 * per)], enc_seed)
                    if r == self.rank:
                        parts.append(ct)
                    else:
                        dist.send(ct.contiguous(), dst=r)
                if r == self.rank and parts:
                    self._ct = torch.cat(parts, dim=0)
        elif self.ghi > self.glo:
            parts = []
            for g0 in range(self.glo, self.ghi, chunk_groups):
                g1 = min(self.ghi, g0 + chunk_groups)
                buf = torch.empty((g1 - g0,) + self._ct_shape, dtype=torch.int64, device=e.dev)
                dist.recv(buf, src=client_rank)
                parts.append(buf)
            self._ct = torch.cat(parts, dim=0)

    def search_scores(self, query: Optional[np.ndarray]) -> Optional[np.ndarray]:
        e = self.engine
        if self.rank == self.client_rank:
            gq = e.encrypt_query(e.quantize(query), self.enc_seed).contiguous()     # fresh ids per query
        else:
            gq = torch.empty(self._q_shape, dtype=torch.float64, device=e.dev)
        if self.world > 1:
            dist.broadcast(gq, src=self.client_rank)
        rows = self.ghi - self.glo
        out = e.scores(gq, self._ct) if rows else None
        if self.world > 1:
            padded = torch.zeros((self.max_groups,) + self._ct_shape, dtype=torch.int64, device=e.dev)
            if rows:
                padded[:rows] = out
            # only the client needs the products: gather (not all-gather), 32 KB per 16 documents
            bufs = None
            if self.rank == self.client_rank:
                bufs = [torch.empty_like(padded) for _ in range(self.world)]
            dist.gather(padded, bufs, dst=self.client_rank)
            if self.rank == self.client_rank:
                parts = []
                for r in range(self.world):
                    glo, ghi = shard_bounds(self.n_groups, self.world, r)
                    parts.append(bufs[r][: ghi - glo])
                out = torch.cat(parts, dim=0)
        if self.rank != self.client_rank:
            return None
        if out is None or not out.shape[0]:
            return np.zeros(0, dtype=np.int64)
        return e.decrypt(out.contiguous(), self.n_docs)

    def search(self, query: Optional[np.ndarray], top_k: int = 5, min_similarity: float = 0.5):
        ints = self.search_scores(query)
        if ints is None:
            return None
        return rank_results(self.doc_ids, self.engine.dequantize(ints), top_k, min_similarity)


class ShardedBootstrap:
    """Keyswitch + programmable bootstrap (the atomic pattern) over a batch of ciphertexts cut into
    contiguous ranges, one per rank (BASELINE.json: "PBS/sec at 1/2/4/8 B200", "the bootstrapping and
    keyswitch keys are broadcast once via NCCL over NVLink, each rank evaluates its shard").

    Ciphertexts are independent units, so the data path has NO collective: the client rank sends rank r its rows
    (point-to-point), every rank runs keyswitch + blind rotation on what it received, and the outputs travel back the same
    way.  Once, at construction, the client broadcasts the evaluation keys; server ranks hold nothing else.

    ``engine``: on the client ``key_tensors()`` returns the evaluation keys (device tensors); on every rank
    ``adopt_keys(list)`` installs them and ``bootstrap(ct, luts, lut_index)`` maps ``[rows, kN+1]`` big-key ciphertexts to
    ``[rows, kN+1]`` bootstrapped ones.  ``device`` is where this rank's buffers live."""

    def __init__(self, engine, client_rank: int = 0, device=None):
        self.engine = engine
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.client_rank = client_rank
        self.dev = torch.device(device) if device is not None else torch.device("cpu")
        self.key_bytes = 0
        self._distribute_keys()

    def _distribute_keys(self):
        e = self.engine
        is_client = self.rank == self.client_rank
        keys = list(e.key_tensors()) if is_client else None
        meta = broadcast_public_material([(tuple(t.shape), t.dtype) for t in keys] if is_client else None, self.client_rank)
        if not is_client:
            keys = [torch.empty(shape, dtype=dtype, device=self.dev) for shape, dtype in meta]
        self.key_bytes = int(sum(t.numel() * t.element_size() for t in keys))
        broadcast_keys(keys, self.client_rank)
        e.adopt_keys(keys)

    def bounds(self, B: int, r: int) -> Tuple[int, int]:
        return shard_bounds(B, self.world, r)

    def scatter(self, ct: Optional[torch.Tensor], lut_index: Optional[torch.Tensor] = None):
        """Client: ``ct [B, words]`` (and optionally one table index per ciphertext).  Returns this rank's rows."""
        is_client = self.rank == self.client_rank
        meta = broadcast_public_material((int(ct.shape[0]), int(ct.shape[1]), lut_index is not None) if is_client else None,
                                         self.client_rank)
        B, words, has_index = meta
        lo, hi = self.bounds(B, self.rank)
        if is_client:
            for r in range(self.world):
                a, b = self.bounds(B, r)
                if r == self.rank or a == b:
                    continue
                dist.send(ct[a:b].contiguous(), dst=r)
                if has_index:
                    dist.send(lut_index[a:b].to(torch.int32).contiguous(), dst=r)
            return B, ct[lo:hi], (lut_index[lo:hi] if has_index else None)
        mine = torch.empty((hi - lo, words), dtype=torch.int64, device=self.dev)
        idx = torch.empty(hi - lo, dtype=torch.int32, device=self.dev) if has_index else None
        if hi > lo:
            dist.recv(mine, src=self.client_rank)
            if has_index:
                dist.recv(idx, src=self.client_rank)
        return B, mine, idx

    def collect(self, B: int, out: torch.Tensor) -> Optional[torch.Tensor]:
        """Every rank's outputs back on the client, in the order of the batch (None elsewhere)."""
        if self.rank != self.client_rank:
            if out.shape[0]:
                dist.send(out.contiguous(), dst=self.client_rank)
            return None
        if self.world == 1:
            return out
        full = torch.empty((B, out.shape[1]), dtype=out.dtype, device=out.device)
        lo, hi = self.bounds(B, self.rank)
        full[lo:hi] = out
        for r in range(self.world):
            a, b = self.bounds(B, r)
            if r != self.rank and b > a:
                dist.recv(full[a:b], src=r)
        return full

    def evaluate(self, ct: Optional[torch.Tensor], luts: torch.Tensor, lut_index: Optional[torch.Tensor] = None):
        """Client: the bootstrapped batch ``[B, kN+1]``; server ranks: None.  ``luts`` (public) is passed on every rank."""
        B, mine, idx = self.scatter(ct, lut_index)
        out = self.engine.bootstrap(mine, luts, idx)
        return self.collect(B, out)


class BootstrapEngine:
    """What :class:`ShardedBootstrap` drives on a GPU: the tensor-core keyswitch followed by the multi-bit blind rotation
    (``fhe_b200_keyswitch_mma`` + ``fhe_b200_pbs_mb2``).  The client builds it from its evaluation keys; a server rank
    builds it from the public parameter set alone and receives the keys by broadcast."""

    def __init__(self, params: dict, device, key_mma: Optional[torch.Tensor] = None, bskf2: Optional[torch.Tensor] = None):
        from . import engine as E
        self.E = E
        self.p = E.make_pbs_params(**params)
        self.dev = torch.device(device)
        self.key_mma, self.bskf2 = key_mma, bskf2
        self._work = None

    def key_tensors(self):
        return [self.key_mma, self.bskf2]

    def adopt_keys(self, keys):
        self.key_mma, self.bskf2 = keys

    def bootstrap(self, ct: torch.Tensor, luts: torch.Tensor, lut_index: Optional[torch.Tensor] = None) -> torch.Tensor:
        E, p = self.E, self.p
        rows = ct.shape[0]
        if rows == 0:
            return torch.empty((0, p.k * p.N + 1), dtype=torch.int64, device=self.dev)
        import ctypes as C
        from . import _native as N
        need = int(N.lib().fhe_b200_keyswitch_mma_workspace_bytes(C.byref(p), rows))
        if self._work is None or self._work.numel() < need:
            self._work = torch.empty(need, dtype=torch.int8, device=self.dev)
        small = E.keyswitch_mma(p, self.key_mma, ct, work=self._work)
        return E.pbs_mb2(p, self.bskf2, small, luts, lut_index)
