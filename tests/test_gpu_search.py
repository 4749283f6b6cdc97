"""compare / search / encrypt-batch on the GPU (fhe="execute") against the clear quantized model the
reference CLI runs (batch_operations.py:233,276): identical scores, identical ranking."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def processor(cuda_dev):
    from fhe_icp_b200.batch_operations import BatchProcessor
    bp = BatchProcessor(fhe="execute", seed=0)
    topics = ["quantum", "cooking", "finance", "biology", "sailing"]
    texts = [f"{topics[i % 5]} document number {i} about things" for i in range(203)]
    bp.encrypt_documents(texts, [f"d{i}" for i in range(203)])
    return bp


def test_search_equals_clear_model_ranking(processor):
    from fhe_icp_b200.batch_operations import rank_results
    bp = processor
    for query, k, thr in [("quantum entanglement", 3, 0.5), ("cooking dinner", 5, 0.5), ("sailing", 10, -1e9),
                          ("unrelated words here", 5, 0.5)]:
        got = bp.search_similar(query, top_k=k, min_similarity=thr)
        q = bp.reducer.transform(bp.embedder.get_embedding(query).reshape(1, -1))[0]
        clear = bp.fhe_model.model.predict((q[None, :] * bp.storage.matrix()).astype(np.float32))
        assert got == rank_results([d["doc_id"] for d in bp.storage.list_documents()], clear, k, thr)
    assert len(bp.search_similar("quantum entanglement", top_k=3)) == 3


def test_compare_equals_clear_model(processor):
    bp = processor
    for a, b in [("d0", "d5"), ("d0", "d1"), ("d7", "d7")]:
        X = (bp.storage.load(a).encrypted_embedding * bp.storage.load(b).encrypted_embedding).reshape(1, -1)
        assert bp.compare_encrypted(a, b) == float(bp.fhe_model.model.predict(X)[0])
    assert bp.compare_encrypted("d0", "d5") > 0.7 > bp.compare_encrypted("d0", "d1")
    with pytest.raises(KeyError):
        bp.compare_encrypted("d0", "nope")


def test_no_model_raises_like_reference():
    from fhe_icp_b200.batch_operations import BatchProcessor
    bp = BatchProcessor(init_model=False)
    for call in (lambda: bp.compare_encrypted("a", "b"), lambda: bp.search_similar("q"), lambda: bp.encrypt_documents(["t"])):
        with pytest.raises(RuntimeError, match="No FHE model initialized"):
            call()


def test_cli_end_to_end(tmp_path, capsys, cuda_dev):
    from fhe_icp_b200.fhe_cli import main
    docs = [{"text": "quantum computing with qubits", "id": "q1", "metadata": {"tag": "physics"}},
            {"text": "quantum error correction", "id": "q2"}, {"text": "cooking pasta recipes", "id": "c1"}, "cooking garlic"]
    (tmp_path / "docs.json").write_text(json.dumps(docs))
    sd = str(tmp_path / "store")
    assert main(["--storage-dir", sd, "encrypt-batch", str(tmp_path / "docs.json"), "-o", str(tmp_path / "ids.json")]) == 0
    assert json.loads((tmp_path / "ids.json").read_text())[:3] == ["q1", "q2", "c1"]
    assert main(["--storage-dir", sd, "compare", "q1", "q2"]) == 0
    out_exec = capsys.readouterr().out
    assert "Similarity score:" in out_exec and "Interpretation: Very similar" in out_exec
    assert main(["--storage-dir", sd, "--fhe", "disable", "compare", "q1", "q2"]) == 0
    out_clear = capsys.readouterr().out
    line = [l for l in out_exec.splitlines() if "Similarity score" in l]
    assert line == [l for l in out_clear.splitlines() if "Similarity score" in l]
    assert main(["--storage-dir", sd, "search", "quantum supremacy", "--top-k", "3"]) == 0
    s = capsys.readouterr().out
    assert "Found 2 similar documents" in s and "1. q" in s and "Metadata: {'tag': 'physics'}" in s


def test_sharded_search_single_rank_on_gpu(processor):
    from fhe_icp_b200.batch_operations import rank_results
    from fhe_icp_b200.sharded_search import ShardedSearch
    bp = processor
    docs = bp.storage.matrix()
    q = bp.embedder.get_embedding("finance markets")
    ss = ShardedSearch(bp.fhe_model, docs, [d["doc_id"] for d in bp.storage.list_documents()])
    got = ss.search(q, top_k=3, min_similarity=0.5)
    clear = bp.fhe_model.model.predict((q[None, :] * docs).astype(np.float32))
    assert got == rank_results(ss.doc_ids, clear, 3, 0.5) and len(got) == 3


def test_large_batch_chunking_is_exact(cuda_dev):
    """5000 rows > one 2 GiB ciphertext chunk: predict_host walks chunks, results stay exact."""
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=21, verbose=False)
    X, _ = m.train(n_samples=500)
    m.compile(X[:10])
    rng = np.random.RandomState(1)
    Xb = X[rng.randint(0, 500, size=5000)] * rng.uniform(0.5, 1.5, size=(5000, 1)).astype(np.float32)
    assert np.array_equal(m.predict_encrypted(Xb), m.predict_clear(Xb))


def test_32bit_wire_form_decrypts_identically(cuda_dev):
    """Modulus-switched scores (2^64 -> 2^32, what crosses NVLink in the sharded search) decrypt to the
    same integers; the extra noise (std ~2^-29) is far below the 2^-23 decoding margin."""
    from fhe_icp_b200 import FHESimilarityModel
    for dt in (np.float32, np.float64):
        m = FHESimilarityModel(input_dim=128, n_bits=8, seed=5, verbose=False)
        X, y = m._prepare_training_data(800)
        m.train(X.astype(dt), y.astype(dt))
        m.compile(X[:10])
        out = m.run(m.encrypt(X))
        y64, q64 = m.decrypt(out, return_q=True)
        y32, q32 = m.decrypt_compressed(m.compress_scores(out), return_q=True)
        assert np.array_equal(q64, q32) and np.array_equal(y64, y32) and np.array_equal(y64, m.predict_clear(X))


def test_both_encrypted_mode_search_and_compare(processor):
    """fhe="both": query and documents both encrypted, the collection stored as ciphertexts only
    (N2).  Scores equal the clear integer model of that path exactly and rank the topic's documents first."""
    from fhe_icp_b200.batch_operations import BatchProcessor, rank_results
    bp = BatchProcessor(fhe="both", seed=0, init_model=False, fhe_model=processor.fhe_model)
    texts = [f"{['quantum', 'cooking', 'finance', 'biology', 'sailing'][i % 5]} document number {i} about things" for i in range(60)]
    bp.encrypt_documents(texts[:37], [f"d{i}" for i in range(37)])       # two calls: ragged groups (37 = 2 x 16 + 5)
    bp.encrypt_documents(texts[37:], [f"d{i}" for i in range(37, 60)])
    assert bp.storage.kind == "glwe" and not bp.storage._rows and bp.storage.n_groups == 3 + 2 and len(bp.storage) == 60
    eng = bp._pair_engine()
    M = bp.embedder.get_embeddings_batch(texts)                          # what the client knew before encrypting
    q = bp.embedder.get_embedding("quantum entanglement")
    ints = bp._pair_scores(eng.quantize(q))
    assert np.array_equal(ints, eng.compare_clear(q, M))                 # exact: the clear integer model
    got = bp.search_similar("quantum entanglement", top_k=12, min_similarity=-10.0)
    assert got == rank_results([f"d{i}" for i in range(60)], eng.dequantize(eng.compare_clear(q, M)), 12, -10.0)
    assert {i for i, _ in got} == {f"d{i}" for i in range(0, 60, 5)}     # the 12 "quantum" documents
    a, b = M[0], M[5]
    got = bp.compare_encrypted("d0", "d5")                               # document 1 decrypted client side -> GGSW query
    assert got == float(eng.dequantize(eng.compare_clear(a, b[None, :]))[0])
    assert abs(got - float(a @ b)) < 0.06 and got > bp.compare_encrypted("d0", "d1")
    assert np.array_equal(bp._decrypt_document(bp.storage.index["d41"]), eng.quantize(M[41]))
    # encrypted threshold: same documents as thresholding the decrypted scores of the same path
    small = BatchProcessor(fhe="both", seed=0, init_model=False, fhe_model=processor.fhe_model, keys=bp.keys)
    small.encrypt_documents([f"{t} document number {i} about things" for i, t in
                             enumerate(["quantum", "cooking", "finance", "quantum", "biology", "sailing", "quantum"])],
                            [f"s{i}" for i in range(7)])
    hits = small.filter_similar("quantum entanglement", 0.5)
    want = [i for i, _ in small.search_similar("quantum entanglement", top_k=10, min_similarity=0.5)]
    assert sorted(hits) == sorted(want) == ["s0", "s3", "s6"]


def test_encrypt_batch_then_search_in_a_new_process_over_stored_ciphertexts(tmp_path, cuda_dev):
    """VERDICT r1 task 7: `encrypt-batch` writes ciphertexts (collection.glwe) and a password-protected key file; a NEW
    process runs `search` / `compare` over what is on disk -- no plaintext embedding exists there -- and gets the ranking
    of the clear integer model."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    from fhe_icp_b200.batch_operations import DocumentStore, SyntheticEmbedder, rank_results
    from fhe_icp_b200.encrypted_compare import PackedEncryptedCompare
    from fhe_icp_b200.serialization import load_ciphertexts, load_keys
    root = Path(__file__).resolve().parent.parent
    sd = str(tmp_path / "store")
    topics = ["quantum", "cooking", "finance", "biology", "sailing"]
    docs = [{"id": f"doc{i}", "text": f"{topics[i % 5]} paper {i} on something", "metadata": {"n": i}} for i in range(45)]
    src = tmp_path / "docs.json"
    src.write_text(json.dumps(docs))
    env = dict(os.environ, FHE_MASTER_PASSWORD="correct horse", PYTHONPATH=str(root))
    run = lambda *a: subprocess.run([sys.executable, "-m", "fhe_icp_b200.fhe_cli", "--storage-dir", sd, "--fhe", "both", *a],
                                    env=env, capture_output=True, text=True, timeout=600)   # noqa: E731
    r = run("encrypt-batch", str(src))
    assert r.returncode == 0 and "Encrypted 45 documents successfully!" in r.stdout, r.stdout + r.stderr
    files = sorted(p.name for p in Path(sd).iterdir())
    assert files == ["collection.glwe", "index.json", "keys.fhe"]            # no embeddings.f32: nothing in the clear
    ct, header = load_ciphertexts(os.path.join(sd, "collection.glwe"), mmap=True)
    assert ct.shape == (3, 2, 2048) and header["meta"]["per"] == 16
    emb = SyntheticEmbedder(128)
    M = emb.get_embeddings_batch([d["text"] for d in docs])
    assert not any(np.float32(v).tobytes() in Path(sd, "collection.glwe").read_bytes() for v in M[0][:4])
    with pytest.raises(ValueError, match="Invalid master password"):
        load_keys(os.path.join(sd, "keys.fhe"), "wrong")
    # a new process: search over the stored ciphertexts
    r = run("search", "quantum entanglement", "--top-k", "4", "--min-similarity", "0.3")
    assert r.returncode == 0, r.stdout + r.stderr
    ref = PackedEncryptedCompare(input_dim=128, device=cuda_dev)             # only its quantizer is used (clear model)
    ref.fit_scale(np.array([-1.0, 1.0]) / np.sqrt(128))
    want = rank_results([d["id"] for d in docs], ref.dequantize(ref.compare_clear(emb.get_embedding("quantum entanglement"), M)), 4, 0.3)
    assert f"Found {len(want)} similar documents" in r.stdout
    for i, (doc_id, score) in enumerate(want, 1):
        assert f"{i}. {doc_id} (similarity: {score:.4f})" in r.stdout, r.stdout
    # ... and a third one: compare two stored documents
    r = run("compare", "doc0", "doc5")
    want_c = float(ref.dequantize(ref.compare_clear(M[0], M[5][None, :]))[0])
    assert r.returncode == 0 and f"Similarity score: {want_c:.4f}" in r.stdout, r.stdout + r.stderr
    # a wrong password cannot use the store
    bad = subprocess.run([sys.executable, "-m", "fhe_icp_b200.fhe_cli", "--storage-dir", sd, "--fhe", "both", "search", "x"],
                         env=dict(env, FHE_MASTER_PASSWORD="nope"), capture_output=True, text=True, timeout=600)
    assert bad.returncode == 1 and "Invalid master password" in bad.stderr
    assert DocumentStore(sd).kind == "glwe"


def _board_model(seed=11, rows=300):
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=seed, verbose=False)
    X, _ = m.train(n_samples=max(rows, 200))
    m.compile(X[:10])
    return m, X[:rows]


def test_score_board_push_is_bit_exact(cuda_dev):
    """Scores pushed by the dot-product kernel into the (here: local) score board equal the 32-bit wire form
    of the plain path word for word, for expanded and seeded ciphertexts, across more steps than slots."""
    import torch
    from fhe_icp_b200.score_board import PeerScoreBoard
    m, X = _board_model()
    B = X.shape[0]
    board = PeerScoreBoard(m, rows_max=B + 7)          # board rows need not equal the shard size
    try:
        for seeded in (False, True):
            ct = m.encrypt(X, seeded=seeded)
            want = m.compress_scores(m.run(ct))
            for _ in range(5):
                step = board.push(ct)
                slot = board.collect()
                assert slot.shape == (board.rows_max, board.M, board.stride)
                assert torch.equal(slot[:B], want), f"step {step}: pushed words differ from modswitch32(run())"
                assert np.array_equal(m.decrypt_compressed(slot[:B]), m.predict_clear(X))
                board.release()
        # an empty shard only flags its arrival
        board.push(None)
        board.collect()
        board.release()
        board.check()
    finally:
        board.close()


def test_score_board_waits_are_bounded(cuda_dev):
    """A flag that never comes (dead peer) must not hang the GPU: the in-stream wait times out and sets the
    status word; a flag that is already there lets the stream through.  On the client's own GPU the slot
    credit is an event, and overwriting an unreleased slot is a host-side error."""
    import ctypes as C
    import torch
    from fhe_icp_b200 import _native as N
    from fhe_icp_b200.score_board import PeerScoreBoard
    ctx = N.context(cuda_dev.index)
    flags = torch.tensor([7, 7, 3], dtype=torch.int64, device=cuda_dev)
    status = torch.zeros(1, dtype=torch.int32, device=cuda_dev)
    N.check(N.lib().fhe_b200_peer_wait(ctx.handle, C.c_void_p(flags.data_ptr()), 2, 7, 30, C.c_void_p(status.data_ptr()), None))
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    N.check(N.lib().fhe_b200_peer_wait(ctx.handle, C.c_void_p(flags.data_ptr()), 3, 7, 30, C.c_void_p(status.data_ptr()), None))
    torch.cuda.synchronize()
    assert int(status.item()) == 1                      # flags[2] = 3 never reaches 7: timed out after 30 ms
    ptrs = torch.tensor([flags.data_ptr() + 16], dtype=torch.int64, device=cuda_dev)
    N.check(N.lib().fhe_b200_peer_signal(ctx.handle, C.c_void_p(ptrs.data_ptr()), 1, 9, None))
    torch.cuda.synchronize()
    assert flags.tolist() == [7, 7, 9]
    m, X = _board_model(rows=64)
    board = PeerScoreBoard(m, rows_max=64, timeout_ms=30)
    try:
        ct = m.encrypt(X, seeded=True)
        board.push(ct)
        board.push(ct)
        with pytest.raises(RuntimeError, match="before release"):
            board.push(ct)          # the third step needs the slot of the first, which nobody released
        # ... and an OLDER release of the same slot does not count: steps 1..4 released, 5 not, 7 must be refused
        board2 = PeerScoreBoard(m, rows_max=64, timeout_ms=30)
        try:
            for step in range(1, 7):
                assert board2.push(ct) == step
                if step != 5:
                    board2.collect()
                    board2.release()
            with pytest.raises(RuntimeError, match="before release"):
                board2.push(ct)
        finally:
            board2.close()
        board.check()
    finally:
        board.close()


def test_score_board_rejects_bad_arguments(cuda_dev):
    import ctypes as C
    from fhe_icp_b200 import _native as N
    m, X = _board_model(rows=8)
    c = m.keygen().model.fhe_circuit
    ct = m.encrypt(X)
    lib = N.lib()
    push = N.Push(0, 0, 1, 0)
    assert lib.fhe_b200_similarity_run_push(c.handle, C.c_void_p(ct.data_ptr()), 8, C.byref(push), None) == N.ERR_INVALID
    assert b"null device pointer" in lib.fhe_b200_last_error()
    push = N.Push(ct.data_ptr(), ct.data_ptr(), 0, ct.data_ptr())
    assert lib.fhe_b200_similarity_run_push(c.handle, C.c_void_p(ct.data_ptr()), 8, C.byref(push), None) == N.ERR_INVALID
    push = N.Push(ct.data_ptr(), ct.data_ptr(), 1, ct.data_ptr())
    assert lib.fhe_b200_similarity_run_push(c.handle, C.c_void_p(ct.data_ptr()), 0, C.byref(push), None) == N.ERR_INVALID
    assert lib.fhe_b200_peer_wait(N.context().handle, None, 1, 1, 10, None, None) == N.ERR_INVALID
    with pytest.raises(ValueError):
        from fhe_icp_b200.sharded_search import ShardedSearch
        ShardedSearch(m, X, gather="carrier-pigeon")


def test_sharded_search_push_mode_equals_nccl_mode(processor):
    from fhe_icp_b200.sharded_search import ShardedSearch
    bp = processor
    docs = bp.storage.matrix()
    ids = [d["doc_id"] for d in bp.storage.list_documents()]
    q = bp.embedder.get_embedding("biology cells")
    want = ShardedSearch(bp.fhe_model, docs, ids).search(q, top_k=4, min_similarity=0.5)
    sp = ShardedSearch(bp.fhe_model, docs, ids, gather="push")
    try:
        for _ in range(4):
            assert sp.search(q, top_k=4, min_similarity=0.5) == want
    finally:
        sp.close()


def test_encrypt_products_equals_host_product_bit_for_bit(cuda_dev):
    """The fused product + quantize + encrypt kernel (fhe_b200_similarity_encrypt_seeded_products) yields the bodies of
    encrypt(query * docs): the float32 multiply on the device is numpy's, including values that sit near a quantizer
    rounding boundary; numpy rows, pinned host tensors and device-resident tensors are the same input."""
    import torch
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=4, key_seed=21, enc_seed=22, noise_seed=23, ct_start=0, verbose=False)
    X, _ = m.train(n_samples=300)
    m.compile(X[:10])
    c = m.model.fhe_circuit
    rng = np.random.RandomState(7)
    q = rng.randn(128).astype(np.float32)
    docs = (rng.randn(777, 128) * rng.choice([1e-3, 0.05, 1.0, 30.0], size=(777, 1))).astype(np.float32)
    docs[5] = 0.0
    docs[6, ::2] = np.float32(1e-30)          # products underflow to denormals / zero exactly as on the host
    c.ct_counter = 1000
    ref = m.encrypt((q[None, :] * docs).astype(np.float32), seeded=True)
    for src in (docs, torch.from_numpy(docs).pin_memory(), torch.from_numpy(docs).to(m.dev)):
        c.ct_counter = 1000
        got = m.encrypt_products(q, src)
        assert got.ct_base == ref.ct_base and got.enc_seed == ref.enc_seed
        assert torch.equal(got.bodies, ref.bodies)
    assert np.array_equal(m.decrypt(m.run(got)), m.predict_clear((q[None, :] * docs).astype(np.float32)))
    with pytest.raises(ValueError):
        m.encrypt_products(q, docs[:, :64])


def test_one_million_seeded_documents_full_size_properties(cuda_dev):
    """BASELINE.json configs[3] at its full size on one GPU: a collection of 1 M documents as seeded ciphertexts (1 GB of
    bodies), searched chunk by chunk.  Size-independent checks: (1) every one of the 1 M decrypted scores equals the clear
    quantized circuit and so does the top-10 ranking; (2) linearity on the raw score ciphertexts of a chunk -- the weighted
    row equals w_min * (plain-sum row) + the row of the shifted weights, word for word, i.e. what the kernel's split
    accumulation assumes; (3) the score ciphertexts of a chunk do not depend on how the collection is cut into launches."""
    import torch
    from bench import collection_query, collection_rows, build_model
    from fhe_icp_b200.batch_operations import top_indices
    m, _ = build_model(0)
    c = m.model.fhe_circuit
    d, total, chunk, seed = c.spec.d, 1_000_000, 62_500, 4242
    q = collection_query(d, seed)
    base = c.next_ct_base(total * d)
    scores = np.empty(total)
    clear = np.empty(total)
    first_out = None
    for lo in range(0, total, chunk):
        X = q[None, :] * collection_rows(lo, lo + chunk, d, seed, q)
        c.ct_counter = base + lo * d                     # ciphertext ids laid out as ONE collection
        sc = m.encrypt(X, seeded=True)
        out = m.run(sc)
        scores[lo:lo + chunk] = m.decrypt(out)
        clear[lo:lo + chunk] = m.predict_clear(X)
        if lo == 0:
            first_out = out[:1000].clone()
            # (3) the same 1000 documents evaluated in a launch of their own: identical ciphertext words
            c.ct_counter = base
            sc_small = m.encrypt(X[:1000], seeded=True)
            assert torch.equal(m.run(sc_small), first_out)
        del sc, out
    assert np.array_equal(scores, clear)                                           # (1) all 1 M scores
    assert np.array_equal(top_indices(scores, 10, -np.inf), top_indices(clear, 10, -np.inf))
    if c.two_outputs:                                                              # (2) linearity, on 1000 documents
        from fhe_icp_b200 import engine as E
        c.ct_counter = base
        X = q[None, :] * collection_rows(0, 1000, d, seed, q)
        ct = m.encrypt(X)                                                          # expanded form of the same ciphertexts
        qw = c.spec.q_weights.astype(np.int64)
        wmin = int(qw.min())
        shifted = torch.as_tensor(np.stack([qw - wmin, np.ones_like(qw)]), device=ct.device)
        parts = E.lincomb(ct, shifted, c.lwe.n, shift=c.lwe.shift)                          # rows: sum (w - wmin) ct, sum ct
        want = parts[:, 0] + wmin * parts[:, 1]                                    # wrapping int64 arithmetic == mod 2^64
        assert torch.equal(first_out[:, 0], want) and torch.equal(first_out[:, 1], parts[:, 1])
