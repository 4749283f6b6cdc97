"""world_size-2 gloo tests (CPU) of the document-sharded search plumbing: shard bounds, the padded
all-gather of encrypted scores, client-side decrypt + threshold + stable sort + top-k.  The engine
behind the plumbing is the CPU oracle here (test infrastructure); on the GPU box the same class runs
over FHESimilarityModel (tests/test_gpu_search.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


class OracleEngine:
    """encrypt / run / decrypt of the compiled circuit on CPU tensors, via the oracle."""

    def __init__(self, model, ct_base=0):
        from oracle import oracle as O
        self.O, self.m, self.c = O, model, model.model.fhe_circuit
        self.s = O.secret_key(self.c.key_seed, 2, self.c.lwe.n)
        self.ct_base = ct_base

    def encrypt(self, X):
        c, O = self.c, self.O
        q = self.m.model.quantize_input(X)
        ct = O.lwe_encrypt(self.s, q, c.lwe.shift, c.lwe.sigma_abs, c.enc_seed, ct_base=self.ct_base, stride=c.lwe.stride,
                           noise_seed=c.noise_seed)
        self.ct_base += q.size
        return torch.from_numpy(ct.reshape(len(X), c.spec.d, -1).view(np.int64))

    def run(self, ct):
        c = self.c
        W = np.stack([c.spec.q_weights, np.ones_like(c.spec.q_weights)]) if c.two_outputs else c.spec.q_weights[None]
        return torch.from_numpy(self.O.lincomb(ct.numpy().view(np.uint64), W, c.lwe.n).view(np.int64))

    def decrypt(self, enc):
        c = self.c
        m = self.O.lwe_decrypt(self.s, enc.numpy().view(np.uint64), c.lwe.shift)
        qy = m[:, 0] - (int(c.spec.weight_q.zero_point) * m[:, 1] if c.two_outputs else 0) + int(c.spec.q_bias)
        return c.spec.dequantize_output(qy)


def _make_problem(n_docs, **seeds):
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=3, verbose=False, **seeds)
    X, _ = m.train(n_samples=300)
    m.compile(X[:10], bound_mode="inputset")
    rng = np.random.RandomState(9)
    q = rng.randn(128).astype(np.float32); q /= np.linalg.norm(q)
    docs = rng.randn(n_docs, 128).astype(np.float32)
    docs[::3] = q + 0.3 * rng.randn(len(docs[::3]), 128)
    docs /= np.linalg.norm(docs, axis=1, keepdims=True)
    return m, q, docs


class KeylessEngine(OracleEngine):
    """What a server rank holds under key_holders='client': it can run, it must never encrypt or decrypt."""

    def __init__(self, model):
        super().__init__(model)
        self.s = None

    def encrypt(self, X):
        raise AssertionError("a server rank tried to encrypt")

    def decrypt(self, enc):
        raise AssertionError("a server rank tried to decrypt")


def _worker(rank, world, port, n_docs, ret, key_holders="client", weights=None, shard_only=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fhe_icp_b200.sharded_search import ShardedSearch, broadcast_public_material, shard_bounds
        if key_holders == "all":   # every rank belongs to the key owner: all built from the SAME explicit key set
            m, q, docs = _make_problem(n_docs, key_seed=11, noise_seed=12, enc_seed=13)
            engine = OracleEngine(m, ct_base=rank << 40)
        else:                      # default: seeds from the OS CSPRNG, different on every rank; only rank 0's matter
            m, q, docs = _make_problem(n_docs)
            engine = OracleEngine(m) if rank == 0 else KeylessEngine(m)
        spec = broadcast_public_material(m.model.spec.to_dict() if rank == 0 else None)
        assert spec["q_weights"] == m.model.spec.to_dict()["q_weights"]
        mine = docs if (rank == 0 or key_holders == "all") else None
        if shard_only:      # every rank is handed just the rows it will encrypt
            lo, hi = shard_bounds(n_docs, world, rank, weights)
            mine = docs[lo:hi]
        ss = ShardedSearch(engine, mine, key_holders=key_holders, shard_weights=weights, n_docs=n_docs if shard_only else None)
        assert (ss.lo, ss.hi) == shard_bounds(n_docs, world, rank, weights)
        res = ss.search(q if (rank == 0 or key_holders == "all") else None, top_k=4, min_similarity=0.2)
        if rank == 0:
            ret["res"] = res
        else:
            assert res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_docs,key_holders", [(7, "client"), (2, "client"), (1, "client"), (7, "all"), (1, "all")])
def test_sharded_search_world2_matches_single_process(n_docs, key_holders):
    """key_holders='client': rank 1 holds no key (its engine raises on encrypt / decrypt) and the ranks' key seeds
    differ (OS CSPRNG) -- the client encrypts every shard and ships the ciphertexts.  'all': same key set on both."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_docs, ret, key_holders), nprocs=world, join=True)
    from fhe_icp_b200.batch_operations import rank_results
    m, q, docs = _make_problem(n_docs)
    ref = rank_results([f"doc_{i}" for i in range(n_docs)], m.predict_clear(q[None, :] * docs), 4, 0.2)
    assert ret["res"] == ref


@pytest.mark.parametrize("weights,shard_only", [([0.25, 1.0], False), ([0.0, 1.0], True), ([1.0, 3.0], True)])
def test_sharded_search_world2_cost_weighted_shards(weights, shard_only):
    """Uneven (cost-weighted) contiguous shards -- including a client that keeps no documents -- and ranks that hold
    only their own rows give the single-process ranking."""
    world, port, n_docs = 2, _free_port(), 9
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_docs, ret, "all", weights, shard_only), nprocs=world, join=True)
    from fhe_icp_b200.batch_operations import rank_results
    m, q, docs = _make_problem(n_docs)
    assert ret["res"] == rank_results([f"doc_{i}" for i in range(n_docs)], m.predict_clear(q[None, :] * docs), 4, 0.2)


def test_weighted_shard_bounds_and_client_cost_weights():
    from fhe_icp_b200.sharded_search import client_cost_weights, shard_bounds
    assert client_cost_weights(1, 0, 0.1) is None and client_cost_weights(8, 0, 0.0) is None
    w = client_cost_weights(8, 0, 0.05)
    assert w[0] == pytest.approx(0.65) and w[1:] == [1.05] * 7 and sum(w) == pytest.approx(8.0)
    assert client_cost_weights(8, 3, 0.5)[3] == 0.0           # the client can be left without a shard, never negative
    for n in (0, 1, 7, 1000, 1_000_003):
        for weights in ([1.0, 1.0], w, [0.0, 2.0, 1.0], client_cost_weights(4, 2, 0.12)):
            world = len(weights)
            spans = [shard_bounds(n, world, r, weights) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(spans, spans[1:]))
            for (lo, hi), wt in zip(spans, weights):          # proportional up to rounding
                assert abs((hi - lo) - n * wt / sum(weights)) <= 1.0
    # time model: with rho the ranks finish together
    n, rho, world = 1_000_000, 0.02, 8
    sizes = [hi - lo for lo, hi in (shard_bounds(n, world, r, client_cost_weights(world, 0, rho)) for r in range(world))]
    assert sizes[0] + rho * n == pytest.approx(sizes[1], abs=2)
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 0, [1.0])


def test_shard_bounds_cover_everything():
    from fhe_icp_b200.sharded_search import shard_bounds
    for n in (0, 1, 5, 8, 1000, 1_000_003):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_rank_results_reference_semantics():
    """Filter >=, stable sort descending, top-k (batch_operations.py:278-284): ties keep index order."""
    from fhe_icp_b200.batch_operations import rank_results
    ids = ["a", "b", "c", "d", "e"]
    s = np.array([0.5, 0.9, 0.5, 0.49999, 0.9])
    assert rank_results(ids, s, 5, 0.5) == [("b", 0.9), ("e", 0.9), ("a", 0.5), ("c", 0.5)]
    assert rank_results(ids, s, 1, 0.5) == [("b", 0.9)]
    assert rank_results(ids, s, 3, 2.0) == []
    assert rank_results([], np.zeros(0), 3, 0.0) == []


# ------------------------------------------------------------------ both vectors encrypted (ShardedPairSearch)
class OraclePairEngine:
    """CPU stand-in for EncryptedCompare (same method names) over the oracle, toy parameters."""
    IN_SHIFT, OUT_SHIFT, P_BITS = 59, 51, 4

    def __init__(self, d, client):
        from oracle import oracle as O
        self.O, self.d, self.dev = O, d, torch.device("cpu")
        self.p = O.make_params(n=16, k=1, N=2048, l_pbs=2, beta_pbs=15, l_ks=4, beta_ks=4, log2_sigma_lwe=-30.0,
                               log2_sigma_glwe=-51.6)
        self.scale = 0.25
        self.bskf = None
        if client:   # only the client rank ever holds secret keys
            self.s, self.S = O.secret_key(5, 0, self.p.n), O.secret_key(5, 1, self.p.N)
            self.bskf = torch.from_numpy(O.bsk_to_fourier(self.p, O.bsk_gen(self.p, self.s, self.S, 6)))

    def quantize(self, X):
        return np.clip(np.rint(np.asarray(X, dtype=np.float64) / self.scale), -4, 3).astype(np.int64)

    def dequantize(self, q):
        return np.asarray(q, dtype=np.float64) * self.scale ** 2

    def _ids(self, count):   # fresh, never-reused ciphertext ids (what the GPU engines' CiphertextIds does)
        base = getattr(self, "_next", 0)
        self._next = base + int(count)
        return base

    def encrypt(self, Xq, enc_seed=None, ct_base=None):
        Xq = np.asarray(Xq)
        enc_seed = 7 if enc_seed is None else enc_seed
        ct_base = self._ids(Xq.size) if ct_base is None else ct_base
        ct = self.O.lwe_encrypt(self.s, Xq, self.IN_SHIFT, self.p.sigma_lwe_abs, enc_seed, ct_base, stride=self.p.n + 2,
                                noise_seed=1234)
        return torch.from_numpy(ct.reshape(Xq.shape + (-1,)).view(np.int64))

    def encrypt_norms(self, Xq, enc_seed=None, ct_base=None):
        Xq = np.asarray(Xq)
        m = (Xq * Xq).sum(axis=-1)
        enc_seed = 7 if enc_seed is None else enc_seed
        ct_base = self._ids(np.size(m)) if ct_base is None else (1 << 40) + ct_base
        ct = self.O.lwe_encrypt(self.S, m, self.OUT_SHIFT - 1, self.p.sigma_glwe_abs, enc_seed, ct_base,
                                stride=self.p.N + 2, noise_seed=1234)
        return torch.from_numpy(ct.reshape(np.shape(m) + (-1,)).view(np.int64))

    def scores(self, ct_q, ct_docs, n_q, n_docs):
        out = self.O.encrypted_product_scores_norms(self.p, self.bskf.numpy(), ct_q.numpy().view(np.uint64),
                                                    ct_docs.numpy().view(np.uint64), n_q.numpy().view(np.uint64),
                                                    n_docs.numpy().view(np.uint64), self.P_BITS, self.OUT_SHIFT)
        pad = np.zeros((out.shape[0], self.p.N + 2), dtype=np.uint64)
        pad[:, : out.shape[1]] = out
        return torch.from_numpy(pad.view(np.int64))

    def decrypt(self, sc):
        v = self.O.lwe_decrypt(self.S, sc.numpy().view(np.uint64), self.OUT_SHIFT) & 8191
        return np.where(v >= 4096, v - 8192, v)


def _pair_problem(n_docs, d=6):
    rng = np.random.RandomState(n_docs)
    return rng.uniform(-1, 1, size=d), rng.uniform(-1, 1, size=(n_docs, d))


def _pair_worker(rank, world, port, n_docs, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fhe_icp_b200.sharded_search import ShardedPairSearch
        q, docs = _pair_problem(n_docs)
        eng = OraclePairEngine(6, client=rank == 0)
        sp = ShardedPairSearch(eng, docs if rank == 0 else None)       # servers never see the documents in clear
        assert (rank == 0) or not hasattr(eng, "s")
        ints = sp.search_scores(q if rank == 0 else None)
        res = sp.search(q if rank == 0 else None, top_k=3, min_similarity=-10.0)
        if rank == 0:
            ret["ints"], ret["res"] = ints.tolist(), res
        else:
            assert ints is None and res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_docs", [5, 2, 1])
def test_sharded_pair_search_world2(n_docs):
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_pair_worker, args=(world, port, n_docs, ret), nprocs=world, join=True)
    from fhe_icp_b200.batch_operations import rank_results
    q, docs = _pair_problem(n_docs)
    eng = OraclePairEngine(6, client=True)
    clear = eng.quantize(docs) @ eng.quantize(q)
    assert ret["ints"] == clear.tolist()
    assert ret["res"] == rank_results([f"doc_{i}" for i in range(n_docs)], eng.dequantize(clear), 3, -10.0)


# ------------------------------------------------------------------ packed both-encrypted search (ShardedPackedSearch)
class OraclePackedEngine:
    """CPU stand-in for PackedEncryptedCompare (same method names) over the oracle."""
    OUT_SHIFT = 47

    def __init__(self, d, client):
        from oracle import oracle as O
        self.O, self.d, self.dev = O, d, torch.device("cpu")
        self.p = O.make_params(n=742, k=1, N=2048, l_pbs=2, beta_pbs=18)
        self.slot = 1 << (d - 1).bit_length()
        self.per = 2048 // self.slot
        self.scale = 1.0 / 16
        self.S = O.secret_key(5, 1, 2048) if client else None      # only the client rank holds the key

    def quantize(self, X):
        return np.clip(np.rint(np.asarray(X, dtype=np.float64) / self.scale), -16, 15).astype(np.int64)

    def dequantize(self, q):
        return np.asarray(q, dtype=np.float64) * self.scale ** 2

    def _ids(self, count):
        base = getattr(self, "_next", 0)
        self._next = base + int(count)
        return base

    def encrypt_documents(self, Yq, enc_seed=None, id_base=None):
        polys = self.O.pack_documents(Yq, 2048, self.slot)
        ct = self.O.glwe_encrypt_rows(self.p, self.S, polys, 0, self.OUT_SHIFT, 7 if enc_seed is None else enc_seed,
                                      self._ids(len(polys)) if id_base is None else id_base, noise_seed=1234)
        return torch.from_numpy(ct.view(np.int64))

    def encrypt_query(self, xq, enc_seed=None, id_base=None):
        gg = self.O.glwe_encrypt_rows(self.p, self.S, self.O.query_polynomial(xq, 2048), 1, 0,
                                      7 if enc_seed is None else enc_seed, self._ids(4) if id_base is None else id_base,
                                      noise_seed=1234)
        return torch.from_numpy(self.O.ggsw_to_fourier(self.p, gg))

    def scores(self, gq, glwe):
        return torch.from_numpy(self.O.glwe_external_product(self.p, gq.numpy(), glwe.numpy().view(np.uint64)).view(np.int64))

    def decrypt(self, prod, n_docs):
        lwe = self.O.glwe_sample_extract(self.p, prod.numpy().view(np.uint64), 0, self.slot, self.per, 2050)
        v = self.O.lwe_decrypt(self.S, lwe, self.OUT_SHIFT)[:n_docs] & 131071
        return np.where(v >= 65536, v - 131072, v)


def _packed_worker(rank, world, port, n_docs, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fhe_icp_b200.sharded_search import ShardedPackedSearch
        rng = np.random.RandomState(n_docs)
        q, docs = rng.uniform(-1, 1, size=100), rng.uniform(-1, 1, size=(n_docs, 100))
        eng = OraclePackedEngine(100, client=rank == 0)
        sp = ShardedPackedSearch(eng, docs if rank == 0 else None, chunk_groups=1)
        assert rank == 0 or eng.S is None
        ints = sp.search_scores(q if rank == 0 else None)
        res = sp.search(q if rank == 0 else None, top_k=4, min_similarity=-100.0)
        if rank == 0:
            ret["ints"], ret["res"] = ints.tolist(), res
        else:
            assert ints is None and res is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_docs", [53, 16, 3])
def test_sharded_packed_search_world2(n_docs):
    """53 documents = 4 ciphertexts (2 per rank, last one ragged); 16 = one ciphertext (rank 1 idle); 3 = partial."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_packed_worker, args=(world, port, n_docs, ret), nprocs=world, join=True)
    from fhe_icp_b200.batch_operations import rank_results
    rng = np.random.RandomState(n_docs)
    q, docs = rng.uniform(-1, 1, size=100), rng.uniform(-1, 1, size=(n_docs, 100))
    eng = OraclePackedEngine(100, client=True)
    clear = eng.quantize(docs) @ eng.quantize(q)
    assert ret["ints"] == clear.tolist()
    assert ret["res"] == rank_results([f"doc_{i}" for i in range(n_docs)], eng.dequantize(clear), 4, -100.0)


BOOT_TOY = dict(n=24, k=1, N=2048, l_pbs=1, beta_pbs=23, l_ks=5, beta_ks=3, log2_sigma_lwe=-30.0, log2_sigma_glwe=-51.6)


class OracleBootstrapEngine:
    """keyswitch + multi-bit PBS on CPU tensors via the oracle: what sharded_search.BootstrapEngine is on a GPU.
    The client is built with the secret keys; a server rank only ever sees what key_tensors() carries."""

    def __init__(self, client: bool):
        from oracle import oracle as O
        self.O, self.p = O, O.make_params(**BOOT_TOY)
        self.ksk = self.bskf2 = None
        if client:
            self.s, self.S = O.secret_key(5, 0, self.p.n), O.secret_key(5, 1, self.p.k * self.p.N)
            self.ksk = O.ksk_gen(self.p, self.S, self.s, 6)
            self.bskf2 = O.bsk2_to_fourier(self.p, O.bsk2_gen(self.p, self.s, self.S, 6))

    def key_tensors(self):
        return [torch.from_numpy(self.ksk.view(np.int64)), torch.from_numpy(self.bskf2)]

    def adopt_keys(self, keys):
        self.ksk, self.bskf2 = keys[0].numpy().view(np.uint64), keys[1].numpy()

    def bootstrap(self, ct, luts, lut_index=None):
        p, O = self.p, self.O
        if ct.shape[0] == 0:
            return torch.empty((0, p.k * p.N + 1), dtype=torch.int64)
        small = O.keyswitch(p, self.ksk, ct.numpy().view(np.uint64))
        li = None if lut_index is None else lut_index.numpy()
        return torch.from_numpy(O.pbs_mb2(p, self.bskf2, small, luts.numpy().view(np.uint64), li).view(np.int64))


def _boot_problem(B):
    from oracle import oracle as O
    eng = OracleBootstrapEngine(client=True)
    rng = np.random.RandomState(B)
    msgs, which = rng.randint(0, 16, size=B), rng.randint(0, 2, size=B).astype(np.int32)
    tables = np.stack([(np.arange(16) * 7 + 3) % 16, (np.arange(16) * 5 + 1) % 16])
    luts = np.stack([O.make_lut_poly(t, 4, eng.p.N, 59) for t in tables])
    ct = O.lwe_encrypt(eng.S, msgs, 59, eng.p.sigma_glwe_abs, 77, ct_base=0, stride=eng.p.N + 2)[:, : eng.p.N + 1]
    return eng, msgs, which, tables, torch.from_numpy(luts.view(np.int64)), torch.from_numpy(np.ascontiguousarray(ct).view(np.int64))


def _boot_worker(rank, world, port, B, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fhe_icp_b200.sharded_search import ShardedBootstrap
        if rank == 0:
            eng, msgs, which, tables, luts, ct = _boot_problem(B)
        else:
            eng, ct, which = OracleBootstrapEngine(client=False), None, None
            luts = _boot_problem(B)[4]                        # the tables are public
        sb = ShardedBootstrap(eng, client_rank=0)
        assert rank == 0 or not hasattr(eng, "s")             # a server rank holds evaluation keys only
        assert sb.key_bytes == sum(t.numel() * t.element_size() for t in eng.key_tensors())
        out = sb.evaluate(ct, luts, torch.from_numpy(which) if rank == 0 else None)
        one = sb.evaluate(ct, luts[:1])                       # one table for the whole batch, no index
        if rank == 0:
            ret["out"], ret["one"] = out.numpy().copy(), one.numpy().copy()
        else:
            assert out is None and one is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 2, 1])
def test_sharded_bootstrap_world2(B):
    """Keyswitch + PBS over a batch cut across two ranks (ragged shards, a rank without rows): the client gets back, in
    order, exactly the words a single process computes, and they decrypt to the chosen table's entry."""
    from oracle import oracle as O
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_boot_worker, args=(world, port, B, ret), nprocs=world, join=True)
    eng, msgs, which, tables, luts, ct = _boot_problem(B)
    single = eng.bootstrap(ct, luts, torch.from_numpy(which)).numpy()
    assert np.array_equal(ret["out"], single)
    assert np.array_equal(ret["one"], eng.bootstrap(ct, luts[:1]).numpy())
    pad = np.zeros((B, eng.p.N + 2), dtype=np.uint64)
    pad[:, : eng.p.N + 1] = ret["out"].view(np.uint64)
    assert np.array_equal(O.lwe_decrypt(eng.S, pad, 59) & 15, tables[which, msgs])


def test_score_board_layout_and_credits():
    """Host logic of the peer score board (no GPU): slots alternate, a slot is reused only after the step
    two back was consumed, and the (slot, rank) regions of the client allocation tile it without overlap."""
    from fhe_icp_b200 import score_board as sb
    assert [sb.slot_of(s) for s in (1, 2, 3, 4)] == [1, 0, 1, 0]
    assert [sb.credit_needed(s) for s in (1, 2, 3, 4, 9)] == [0, 0, 1, 2, 7]
    for s in range(3, 50):
        assert sb.slot_of(sb.credit_needed(s)) == sb.slot_of(s)
    world, rows, M, stride = 8, 1000, 2, 1424
    spans, flags = [], set()
    for slot in range(sb.SLOTS):
        for r in range(world):
            a, off = sb.board_offsets(world, rows, M, stride, slot, r)
            assert a % 8 == 0 and a + 8 <= sb.HEADER_BYTES and off % 16 == 0
            flags.add(a)
            spans.append((off, off + 4 * rows * M * stride))
    assert len(flags) == sb.SLOTS * world
    spans.sort()
    assert spans[0][0] == sb.HEADER_BYTES and spans[-1][1] == sb.board_bytes(world, rows, M, stride)
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_score_board_protocol_is_safe_and_live_under_any_schedule():
    """Model of the flag protocol built on score_board.slot_of / credit_needed: `world` servers push steps
    1..K (a push may start only when credit[slot] >= credit_needed(step)), the client consumes a step once every
    arrival flag of its slot carries that step and then writes the step into every rank's credit flag.  Under
    randomly interleaved schedules the run always finishes (no deadlock), a slot is never overwritten before the
    client consumed it, and the client reads exactly the step it waited for."""
    from fhe_icp_b200 import score_board as sb
    rng = np.random.RandomState(7)
    for world in (1, 2, 3, 8):
        for trial in range(40):
            K = 9
            nxt = [1] * world                                   # next step each server will push
            arrive = [[0] * world for _ in range(sb.SLOTS)]     # client memory
            content = [[0] * world for _ in range(sb.SLOTS)]    # which step's scores sit in (slot, rank)
            credit = [[0] * sb.SLOTS for _ in range(world)]     # each rank's memory
            consumed = 0
            for _ in range(20000):
                if consumed == K:
                    break
                actor = rng.randint(0, world + 1)
                if actor < world:                               # server `actor` tries to push its next step
                    s = nxt[actor]
                    if s > K:
                        continue
                    k = sb.slot_of(s)
                    if credit[actor][k] < sb.credit_needed(s):
                        continue                                # bounded wait: not yet
                    assert content[k][actor] <= consumed, "slot overwritten before the client consumed it"
                    content[k][actor] = s
                    arrive[k][actor] = s
                    nxt[actor] = s + 1
                else:                                           # client tries to consume step consumed + 1
                    s = consumed + 1
                    k = sb.slot_of(s)
                    if min(arrive[k]) < s:
                        continue
                    assert content[k] == [s] * world, "client read a slot holding another step's scores"
                    consumed = s
                    for r in range(world):
                        credit[r][k] = s
            assert consumed == K, f"deadlock: world={world}, consumed={consumed}, next={nxt}"
