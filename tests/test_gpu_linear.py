"""Parity of the CUDA linear path (through the C-ABI) against the CPU oracle: bit-exact on
key bits, ciphertext words, linear combinations and decrypted integers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _u64(t):
    return t.detach().cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("dim,key_id", [(1, 0), (127, 0), (128, 1), (1423, 2), (2048, 1), (4097, 5)])
def test_secret_key_bits_match_oracle(O, cuda_dev, dim, key_id):
    from fhe_icp_b200 import engine as E
    k = E.secret_key(1234567, key_id, dim, cuda_dev).cpu().numpy()
    assert np.array_equal(k, O.secret_key(1234567, key_id, dim))


@pytest.mark.parametrize("n", [15, 16, 630, 1023, 1423])
def test_encrypt_words_match_oracle(O, cuda_dev, n):
    import torch
    from fhe_icp_b200 import engine as E
    rng = np.random.RandomState(n)
    msgs = rng.randint(-128, 128, size=37)
    shift, sigma = 42, 2.0 ** (64 - 30.0)
    key = E.secret_key(99, 2, n, cuda_dev)
    stride = E.even_stride(n)
    ct = E.lwe_encrypt(key, torch.as_tensor(msgs), shift, sigma, enc_seed=555, ct_base=1000, stride=stride, noise_seed=9001)
    ref = O.lwe_encrypt(O.secret_key(99, 2, n), msgs, shift, sigma, 555, ct_base=1000, stride=stride, noise_seed=9001)
    assert np.array_equal(_u64(ct), ref)
    dec = E.lwe_decrypt(key, ct, shift).cpu().numpy()
    assert np.array_equal(dec, msgs)
    ph = _u64(E.lwe_phase(key, ct))
    assert np.array_equal(ph, O.lwe_phase(O.secret_key(99, 2, n), ref))


def test_gaussian_noise_stream_matches_oracle(O, cuda_dev):
    """Noise is the body minus <a,s> minus the message: compare 4096 samples bit for bit."""
    import torch
    from fhe_icp_b200 import engine as E
    n, shift, sigma = 16, 40, 2.0 ** 39.3
    key = E.secret_key(5, 2, n, cuda_dev)
    msgs = np.zeros(4096, dtype=np.int64)
    ct = E.lwe_encrypt(key, torch.as_tensor(msgs), shift, sigma, enc_seed=77, ct_base=1 << 40, noise_seed=78)
    ref = O.lwe_encrypt(O.secret_key(5, 2, n), msgs, shift, sigma, 77, ct_base=1 << 40, stride=E.even_stride(n), noise_seed=78)
    assert np.array_equal(_u64(ct), ref)
    noise = O.lwe_phase(O.secret_key(5, 2, n), ref).view(np.int64).astype(np.float64)
    assert abs(noise.std() / sigma - 1.0) < 0.05 and abs(noise.mean()) < 0.1 * sigma


@pytest.mark.parametrize("B,d,n,M", [(1, 128, 1423, 2), (3, 128, 1023, 1), (5, 7, 33, 2), (2, 256, 630, 2),
                                      (4, 1, 15, 1), (9, 130, 200, 2)])
def test_lincomb_bit_exact(O, cuda_dev, B, d, n, M):
    import torch
    from fhe_icp_b200 import engine as E
    rng = np.random.RandomState(B * 1000 + d)
    stride = E.even_stride(n)
    ct = rng.randint(0, 2 ** 63, size=(B, d, stride), dtype=np.int64).astype(np.uint64) * np.uint64(2) + \
        rng.randint(0, 2, size=(B, d, stride)).astype(np.uint64)
    ct[..., n + 1:] = 0
    W = rng.randint(-2 ** 40, 2 ** 40, size=(M, d))
    W[0, : min(d, 3)] = [-128, 127, 0][: min(d, 3)]
    bias = rng.randint(-1000, 1000, size=M)
    out = E.lincomb(E.from_u64_numpy(ct, cuda_dev), torch.as_tensor(W), n, bias=bias, shift=41)
    ref = O.lincomb(ct, W, n, bias=bias, shift=41)
    assert np.array_equal(_u64(out), ref)


def test_lincomb_empty_batch(cuda_dev):
    import torch
    from fhe_icp_b200 import engine as E
    ct = torch.empty((0, 8, 16), dtype=torch.int64, device=cuda_dev)
    out = E.lincomb(ct, torch.ones((1, 8), dtype=torch.int64), 15)
    assert out.shape == (0, 1, 16)


def test_accumulate(cuda_dev):
    import torch
    from fhe_icp_b200 import engine as E
    rng = np.random.RandomState(0)
    a = rng.randint(-2 ** 62, 2 ** 62, size=1001, dtype=np.int64)
    b = rng.randint(-2 ** 62, 2 ** 62, size=1001, dtype=np.int64)
    acc = torch.as_tensor(a).to(cuda_dev)
    E.accumulate(acc, torch.as_tensor(b).to(cuda_dev))
    assert np.array_equal(acc.cpu().numpy().view(np.uint64), a.view(np.uint64) + b.view(np.uint64))


@pytest.mark.parametrize("fit_dtype,n_bits", [(np.float32, 8), (np.float64, 8), (np.float32, 4), (np.float32, 12)])
def test_predict_encrypted_equals_clear_circuit(O, cuda_dev, fit_dtype, n_bits):
    """The reference's invariant (test_fhe.py:56-57; fhe_similarity.py:268-269): FHE result ==
    clear quantized result -- here exactly, on the integers and on the float scores."""
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=128, n_bits=n_bits, seed=7, verbose=False)
    X, y = m._prepare_training_data(600)
    m.train(X.astype(fit_dtype), y.astype(fit_dtype))
    m.compile(X[:10])
    Xt = X[:200]
    fhe = m.predict_encrypted(Xt)
    clear = m.predict_clear(Xt)
    assert np.array_equal(fhe, clear)
    # independent restatement of the quantizer + circuit (oracle/oracle.py)
    s = m.model.spec
    sc, zp, off = O.uniform_quantizer_params(X.astype(fit_dtype), n_bits, True)
    assert (sc, zp, off) == (s.input_q.scale, s.input_q.zero_point, s.input_q.offset)
    q = O.quantize(Xt, sc, zp, off, n_bits)
    qy = O.clear_circuit(q, s.q_weights, s.weight_q.zero_point, s.q_bias)
    y2, qy2 = m.model.fhe_circuit.encrypt_run_decrypt(Xt, return_q=True)
    assert np.array_equal(qy2, qy)
    assert np.array_equal(y2, clear)


def test_split_api_ciphertexts_match_oracle(O, cuda_dev):
    """keygen / encrypt / run / decrypt: every intermediate ciphertext equals the oracle's."""
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=3, verbose=False, ct_start=0)   # seeds from the OS CSPRNG
    X, _ = m.train(n_samples=300)
    m.compile(X[:10])
    c = m.model.fhe_circuit
    assert len({c.key_seed, c.noise_seed, c.enc_seed}) == 3
    Xt = X[:6]
    ct = m.encrypt(Xt)
    s = O.secret_key(c.key_seed, 2, c.lwe.n)
    q = m.model.quantize_input(Xt)
    ref_ct = O.lwe_encrypt(s, q, c.lwe.shift, c.lwe.sigma_abs, c.enc_seed, ct_base=0, stride=c.lwe.stride,
                           noise_seed=c.noise_seed)
    assert np.array_equal(_u64(ct).reshape(-1, c.lwe.stride), ref_ct)
    out = m.run(ct)
    W = np.stack([c.spec.q_weights, np.ones_like(c.spec.q_weights)]) if c.two_outputs else c.spec.q_weights[None]
    ref_out = O.lincomb(ref_ct.reshape(6, 128, -1), W, c.lwe.n)
    assert np.array_equal(_u64(out), ref_out)
    y, qy = m.decrypt(out, return_q=True)
    assert np.array_equal(qy, c.spec.circuit(q))
    assert np.array_equal(y, m.predict_clear(Xt))


def test_toy_linear_model_y_equals_2x(cuda_dev):
    """The reference's smoke test (test_fhe.py:10-57): y = 2x, x = 7 -> ~14, |fhe - clear| < 0.01."""
    from fhe_icp_b200 import LinearRegression
    X = np.array([[1], [2], [3], [4], [5], [6]], dtype=np.float32)
    y = np.array([2, 4, 6, 8, 10, 12], dtype=np.float32)
    model = LinearRegression(n_bits=8)
    model.fit(X, y)
    model.compile(X)
    t = np.array([[7]], dtype=np.float32)
    clear = model.predict(t)
    fhe = model.predict(t, fhe="execute")
    assert abs(fhe[0] - clear[0]) < 0.01 and fhe[0] == clear[0]
    assert abs(clear[0] - 12.0) < 0.2  # x=7 clips to the calibrated range [1,6] -> 12
    assert model.fhe_circuit.graph.maximum_integer_bit_width() >= 8


def test_noise_variance_within_bound(O, cuda_dev):
    """Output noise of the dot product: measured variance vs sum_j w_j^2 sigma^2 (north star:
    'noise variance within a stated bound')."""
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=11, verbose=False)
    X, _ = m.train(n_samples=400)
    m.compile(X[:10])
    c = m.model.fhe_circuit
    Xt = np.tile(X[:64], (8, 1))
    out = m.run(m.encrypt(Xt))
    s = O.secret_key(c.key_seed, 2, c.lwe.n)
    ph = O.lwe_phase(s, _u64(out)).view(np.int64)
    q = m.model.quantize_input(Xt)
    exp0 = (q @ c.spec.q_weights).astype(np.int64) << c.lwe.shift
    err = (ph[:, 0] - exp0).astype(np.float64)
    pred_std = c.lwe.sigma_abs * np.sqrt(float((c.spec.q_weights.astype(np.float64) ** 2).sum()))
    assert 0.8 < err.std() / pred_std < 1.2
    assert np.abs(err).max() < 2.0 ** (c.lwe.shift - 1)


@pytest.mark.parametrize("n,d,M", [(1423, 128, 2), (1023, 128, 1), (16, 5, 2), (15, 3, 1), (630, 130, 2)])
def test_seeded_ciphertexts_bit_identical_to_expanded(O, cuda_dev, n, d, M):
    """Seeded form (bodies + public seed): expand(seeded) == encrypt == oracle, and the on-the-fly
    dot product equals the dot product of the materialised ciphertexts, word for word."""
    import torch
    from fhe_icp_b200 import engine as E
    rng = np.random.RandomState(n + d)
    B, shift, sigma = 7, 42, 2.0 ** 28
    msgs = rng.randint(-128, 128, size=(B, d))
    key = E.secret_key(99, 2, n, cuda_dev)
    stride = E.even_stride(n)
    bodies = E.lwe_encrypt_seeded(key, torch.as_tensor(msgs), shift, sigma, enc_seed=555, ct_base=12345, noise_seed=31337)
    full = E.lwe_encrypt(key, torch.as_tensor(msgs), shift, sigma, enc_seed=555, ct_base=12345, stride=stride, noise_seed=31337)
    assert np.array_equal(_u64(bodies), _u64(full)[..., n])
    assert np.array_equal(_u64(E.lwe_expand_seeded(bodies, n, 555, 12345, stride=stride)), _u64(full))
    ref_ct = O.lwe_encrypt(O.secret_key(99, 2, n), msgs, shift, sigma, 555, ct_base=12345, stride=stride, noise_seed=31337)
    assert np.array_equal(_u64(full).reshape(-1, stride), ref_ct)
    W = rng.randint(-128, 128, size=(M, d))
    bias = rng.randint(-50, 50, size=M)
    got = E.lincomb_seeded(bodies, torch.as_tensor(W), n, 555, 12345, bias=bias, shift=shift, stride=stride)
    assert np.array_equal(_u64(got), O.lincomb(ref_ct.reshape(B, d, stride), W, n, bias=bias, shift=shift))
    assert np.array_equal(O.lwe_decrypt(O.secret_key(99, 2, n), _u64(got), shift), msgs @ W.T + bias)


def test_seeded_and_expanded_model_paths_agree(cuda_dev):
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=17, verbose=False)
    X, _ = m.train(n_samples=400)
    m.compile(X[:10])
    c = m.model.fhe_circuit
    assert c.ciphertext_format == "seeded"
    y_seeded = m.predict_encrypted(X)
    c.ciphertext_format = "expanded"
    y_expanded = m.predict_encrypted(X)
    assert np.array_equal(y_seeded, y_expanded) and np.array_equal(y_seeded, m.predict_clear(X))
    sc = m.encrypt(X[:50], seeded=True)
    assert sc.nbytes() == 50 * 128 * 8
    out_s = m.run(sc)
    c.ct_counter = sc.ct_base                      # same ciphertext ids for the expanded form
    out_e = m.run(m.encrypt(X[:50]))
    assert np.array_equal(_u64(out_s), _u64(out_e))
    assert np.array_equal(m.decrypt(out_s), m.predict_clear(X[:50]))


def test_edge_shapes_and_empty_batches(O, cuda_dev):
    """Ragged / degenerate shapes through the C-ABI: empty batches, a single feature, d below the
    unroll factor, odd word counts in the 32-bit wire form."""
    import ctypes as C
    import torch
    from fhe_icp_b200 import _native as N
    from fhe_icp_b200 import engine as E
    n = 33
    key = E.secret_key(4, 2, n, cuda_dev)
    stride = E.even_stride(n)
    empty = torch.empty((0, 3), dtype=torch.int64)
    assert E.lwe_encrypt(key, empty, 40, 1.0, 1).shape == (0, 3, stride)
    assert E.lwe_encrypt_seeded(key, empty, 40, 1.0, 1).shape == (0, 3)
    assert E.lincomb_seeded(torch.empty((0, 3), dtype=torch.int64, device=cuda_dev), torch.ones((1, 3), dtype=torch.int64),
                            n, 1).shape == (0, 1, stride)
    assert E.lwe_decrypt(key, torch.empty((0, stride), dtype=torch.int64, device=cuda_dev), 40).shape == (0,)
    for d in (1, 2, 7, 9):  # below / around the 8-row register pipeline
        msgs = np.arange(5 * d).reshape(5, d) - 7
        ct = E.lwe_encrypt(key, torch.as_tensor(msgs), 40, 2.0 ** 10, 3, ct_base=d)
        W = (np.arange(d) % 5 - 2).reshape(1, d)
        out = E.lincomb(ct, torch.as_tensor(W), n)
        ref = O.lincomb(_u64(ct), W, n)
        assert np.array_equal(_u64(out), ref)
        assert np.array_equal(E.lwe_decrypt(key, out, 40).cpu().numpy(), msgs @ W.T)
    # 32-bit wire form on an odd number of words
    x = torch.as_tensor(np.random.RandomState(0).randint(-2 ** 62, 2 ** 62, size=3 * stride, dtype=np.int64)).to(cuda_dev)
    y = torch.empty(3 * stride, dtype=torch.int32, device=cuda_dev)
    ctx = N.context(cuda_dev.index)
    N.check(N.lib().fhe_b200_lwe_modswitch32(ctx.handle, C.c_void_p(x.data_ptr()), 3, stride, C.c_void_p(y.data_ptr()), None))
    exp = ((x.cpu().numpy().view(np.uint64) + np.uint64(1 << 31)) >> np.uint64(32)).astype(np.uint32)
    assert np.array_equal(y.cpu().numpy().view(np.uint32), exp)


def test_invalid_arguments_on_device(cuda_dev):
    import ctypes as C
    import torch
    from fhe_icp_b200 import _native as N
    from fhe_icp_b200 import engine as E
    ctx = N.context(cuda_dev.index)
    key = E.secret_key(4, 2, 16, cuda_dev)
    ct = torch.zeros((1, 2, 17), dtype=torch.int64, device=cuda_dev)  # odd stride is rejected, not mis-read
    rc = N.lib().fhe_b200_lincomb(ctx.handle, C.c_void_p(ct.data_ptr()), 1, 2, 16, 17, C.c_void_p(ct.data_ptr()), 1, None, 0,
                                  C.c_void_p(ct.data_ptr()), None)
    assert rc == N.ERR_INVALID and b"stride" in N.lib().fhe_b200_last_error()
    with pytest.raises(N.FheB200Error):
        E.lincomb(torch.zeros((1, 2, 18), dtype=torch.int64, device=cuda_dev), torch.ones((3, 2), dtype=torch.int64), 16)  # M=3
    p = E.make_pbs_params(n=16, N_poly=1024)
    with pytest.raises(N.FheB200Error, match="N must be 2048"):
        E.ksk_gen(p, key, key, 1)


@pytest.mark.parametrize("n,stride,two", [(1, 2, 1), (5, 6, 1), (30, 32, 0), (33, 34, 1), (64, 66, 1), (127, 128, 0),
                                          (743, 744, 1), (742, 746, 1), (1423, 1424, 1)])
def test_fused_similarity_decrypt_all_row_shapes(cuda_dev, n, stride, two):
    """The fused client kernel (decrypt + decode + dequantize, one CTA per document) on row shapes that hit
    every vector width / tail case, in the native 64-bit and the 32-bit wire form: the integers equal
    m0 - zp_w*m1 + q_bias computed in numpy and the floats equal out_scale*(q - zp_out) exactly."""
    import ctypes as C
    import torch
    from fhe_icp_b200 import _native as N
    from fhe_icp_b200 import engine as E
    M = 2 if two else 1
    shift, key_seed = 44, 77 + n
    zp_w, q_bias, out_scale, out_zp = -31337, 12345, 3.25e-7, -99
    spec = N.SimilaritySpec(d=3, n_bits=8, n=n, stride=stride, shift=shift, two_outputs=two, sigma_abs=2.0 ** 20,
                            x_scale=1.0, x_zero_point=0, x_offset=128, w_zero_point=zp_w if two else 0, q_bias=q_bias,
                            out_scale=out_scale, out_zero_point=out_zp, key_seed=key_seed)
    ctx = N.context(cuda_dev.index)
    h = C.c_void_p()
    qw = (C.c_int64 * 3)(1, 2, 3)
    N.check(N.lib().fhe_b200_similarity_create(ctx.handle, C.byref(spec), qw, C.byref(h)))
    try:
        B = 37
        rng = np.random.RandomState(n)
        msgs = rng.randint(-2 ** 17, 2 ** 17, size=(B, M)).astype(np.int64)
        key = E.secret_key(key_seed, 2, n, cuda_dev)           # the model's client key (key id 2)
        ct = E.lwe_encrypt(key, torch.as_tensor(msgs), shift, 2.0 ** 20, 5, stride=stride)
        q_want = msgs[:, 0] - (zp_w * msgs[:, 1] if two else 0) + q_bias
        y_want = out_scale * (q_want - out_zp).astype(np.float64)
        y = torch.empty(B, dtype=torch.float64, device=cuda_dev)
        qy = torch.empty(B, dtype=torch.int64, device=cuda_dev)
        N.check(N.lib().fhe_b200_similarity_decrypt(h, C.c_void_p(ct.data_ptr()), B, C.c_void_p(y.data_ptr()),
                                                    C.c_void_p(qy.data_ptr()), None))
        assert np.array_equal(qy.cpu().numpy(), q_want) and np.array_equal(y.cpu().numpy(), y_want)
        ct32 = torch.empty(ct.shape, dtype=torch.int32, device=cuda_dev)
        N.check(N.lib().fhe_b200_lwe_modswitch32(ctx.handle, C.c_void_p(ct.data_ptr()), B * M, stride,
                                                 C.c_void_p(ct32.data_ptr()), None))
        y.zero_(); qy.zero_()
        N.check(N.lib().fhe_b200_similarity_decrypt32(h, C.c_void_p(ct32.data_ptr()), B, None, C.c_void_p(qy.data_ptr()), None))
        assert np.array_equal(qy.cpu().numpy(), q_want)
        # a view that starts 8 bytes into a 16-byte line exercises the unaligned wire-form path
        if stride % 4 == 2:
            sub = ct32[1:]
            N.check(N.lib().fhe_b200_similarity_decrypt32(h, C.c_void_p(sub.data_ptr()), B - 1, C.c_void_p(y.data_ptr()),
                                                          C.c_void_p(qy.data_ptr()), None))
            assert np.array_equal(qy.cpu().numpy()[:B - 1], q_want[1:])
            assert np.array_equal(y.cpu().numpy()[:B - 1], y_want[1:])
    finally:
        N.lib().fhe_b200_similarity_destroy(h)


@pytest.mark.parametrize("n_bits", [4, 8, 12])
def test_quantization_strategy_sweep_fhe_equals_clear(cuda_dev, n_bits):
    """The reference's quantization benchmark (quantization_strategy.py:28-90,134-160): SGDRegressor(n_bits,
    max_iter=20, random_state=42) on the seed-42 dataset of CONCATENATED 256-d pairs, compiled on the training
    set, then five rows predicted one at a time with fhe="execute" and compared with the clear prediction.  The
    reference records `clear_vs_fhe_mae`; here it is exactly zero, for every bit width it sweeps."""
    from fhe_icp_b200 import SGDRegressor
    np.random.seed(42)
    n_samples, dim = 500, 128
    emb1 = np.random.randn(n_samples, dim).astype(np.float32)
    emb1 = emb1 / np.linalg.norm(emb1, axis=1, keepdims=True)
    emb2 = np.zeros_like(emb1)
    for i in range(n_samples):
        emb2[i] = emb1[i] + 0.1 * np.random.randn(dim) if i % 2 == 0 else np.random.randn(dim)
    emb2 = emb2 / np.linalg.norm(emb2, axis=1, keepdims=True)
    X, y = np.hstack([emb1, emb2]), np.sum(emb1 * emb2, axis=1)
    X_train, X_test, y_train = X[:400], X[400:], y[:400]
    model = SGDRegressor(n_bits=n_bits, max_iter=20, random_state=42)
    model.fit(X_train, y_train)
    model.compile(X_train)
    assert model.fhe_circuit.graph.maximum_integer_bit_width() >= n_bits
    clear = model.predict(X_test[:5])
    fhe = np.array([model.predict(X_test[i:i + 1], fhe="execute")[0] for i in range(5)])
    assert np.mean(np.abs(clear - fhe)) == 0.0
    assert np.array_equal(model.predict(X_test, fhe="execute"), model.predict(X_test))


def test_public_material_does_not_reproduce_the_noise(O, cuda_dev):
    """the error terms of seeded ciphertexts must not be derivable from what the evaluator sees
    (bodies, public mask seed, ciphertext ids).  With the secret noise seed the oracle reproduces every body; with the
    public mask seed in its place -- all the evaluator could try -- not one error term matches, and default seeds are
    fresh for every compiled model."""
    from fhe_icp_b200 import FHESimilarityModel
    from fhe_icp_b200 import engine as E
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=4, verbose=False, ct_start=0)
    X, _ = m.train(n_samples=300)
    m.compile(X[:10])
    c = m.model.fhe_circuit
    m2 = FHESimilarityModel(input_dim=128, n_bits=8, seed=4, verbose=False)
    m2.train(n_samples=300)
    m2.compile(X[:10])
    c2 = m2.model.fhe_circuit
    assert {c.key_seed, c.noise_seed, c.enc_seed}.isdisjoint({c2.key_seed, c2.noise_seed, c2.enc_seed})
    assert c2.ct_counter != 0                                  # random id origin unless ct_start is given
    sc = m.encrypt(X[:4], seeded=True)
    assert not hasattr(sc, "noise_seed") and sc.enc_seed == c.enc_seed
    s = O.secret_key(c.key_seed, 2, c.lwe.n)
    q = m.model.quantize_input(X[:4])
    good = O.lwe_encrypt(s, q, c.lwe.shift, c.lwe.sigma_abs, sc.enc_seed, ct_base=sc.ct_base, stride=c.lwe.stride,
                         noise_seed=c.noise_seed)[:, c.lwe.n]
    guess = O.lwe_encrypt(s, q, c.lwe.shift, c.lwe.sigma_abs, sc.enc_seed, ct_base=sc.ct_base, stride=c.lwe.stride,
                          noise_seed=sc.enc_seed)[:, c.lwe.n]
    bodies = E.to_u64_numpy(sc.bodies).reshape(-1)
    assert np.array_equal(bodies, good)
    assert not np.any(bodies == guess)
    # two encryptions of the same rows never share ciphertext ids (mask + error reuse would leak m1 - m2)
    sc2 = m.encrypt(X[:4], seeded=True)
    assert sc2.ct_base >= sc.ct_base + 4 * 128 and not np.any(E.to_u64_numpy(sc2.bodies).reshape(-1) == bodies)


def test_evaluator_handle_runs_but_cannot_encrypt_or_decrypt(cuda_dev):
    """the server side of the linear path is a key-less object."""
    import ctypes as C
    import torch
    from fhe_icp_b200 import FHESimilarityModel
    from fhe_icp_b200 import _native as N
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=6, verbose=False)
    X, _ = m.train(n_samples=300)
    m.compile(X[:10])
    c = m.model.fhe_circuit
    ct = m.encrypt(X[:5])
    ev = c.evaluator_handle()
    out = torch.empty((5, 2 if c.two_outputs else 1, c.lwe.stride), dtype=torch.int64, device=ct.device)
    lib = N.lib()
    assert lib.fhe_b200_similarity_run(ev, C.c_void_p(ct.data_ptr()), 5, C.c_void_p(out.data_ptr()), None) == N.OK
    torch.cuda.synchronize()
    assert torch.equal(out, m.run(ct))
    assert np.array_equal(m.decrypt(out), m.predict_clear(X[:5]))
    y = torch.empty(5, dtype=torch.float64, device=ct.device)
    assert lib.fhe_b200_similarity_decrypt(ev, C.c_void_p(out.data_ptr()), 5, C.c_void_p(y.data_ptr()), None, None) == N.ERR_INVALID
    assert b"evaluator-only" in lib.fhe_b200_last_error()
    Xd = torch.as_tensor(X[:5]).to(ct.device)
    assert lib.fhe_b200_similarity_encrypt(ev, C.c_void_p(Xd.data_ptr()), 5, 1, 0, C.c_void_p(ct.data_ptr()), None) == N.ERR_INVALID
    hy = np.empty(5)
    assert lib.fhe_b200_similarity_predict_host_seeded(ev, X[:5].ctypes.data_as(C.POINTER(C.c_float)), 5, 1, 0,
                                                       hy.ctypes.data_as(C.POINTER(C.c_double)), None) == N.ERR_INVALID


def test_float64_inputs_execute_equals_clear(cuda_dev):
    """float64 rows (what the reference's reducer hands over) are quantized in float64, so
    fhe="execute" equals the clear model even where rounding the input to float32 first would flip a quantized value."""
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=8, verbose=False)
    X, _ = m.train(n_samples=400)
    m.compile(X[:10])
    q = m.model.spec.input_q
    rng = np.random.RandomState(0)
    # rows sitting a hair from the rounding boundaries of the input quantizer
    k = rng.randint(q.qmin + 1, q.qmax - 1, size=(64, 128)).astype(np.float64)
    X64 = (k + 0.5 - q.zero_point) * q.scale * (1 + rng.choice([-1, 1], size=k.shape) * 1e-9)
    assert np.any(q.quant(X64) != q.quant(X64.astype(np.float32)))          # the float32 detour would differ
    assert np.array_equal(m.predict_encrypted(X64), m.predict_clear(X64))
    assert np.array_equal(m.model.predict(X64, fhe="execute"), m.model.predict(X64))


def test_wire32_gate_counts_the_modswitch_noise(cuda_dev):
    """the 32-bit wire form is allowed only where z * sqrt(sigma_out^2 + (n/2+1) 2^64/12) < Delta/2.
    At 8 bits / d=128 it is; at a 12-bit quantization the message is too wide and compress_scores / the score board
    refuse instead of returning wrong scores, while the 64-bit path stays exact."""
    from fhe_icp_b200 import FHESimilarityModel
    from fhe_icp_b200.score_board import PeerScoreBoard
    ok = FHESimilarityModel(input_dim=128, n_bits=8, seed=5, verbose=False)
    X, _ = ok.train(n_samples=300)
    ok.compile(X[:10])
    assert ok.wire32_supported
    wide = FHESimilarityModel(input_dim=128, n_bits=12, seed=5, verbose=False)
    Xw, _ = wide.train(n_samples=300)
    wide.compile(Xw[:10])
    c = wide.model.fhe_circuit
    assert c.lwe.shift < 44 and not wide.wire32_supported
    out = wide.run(wide.encrypt(Xw[:16]))
    assert np.array_equal(wide.decrypt(out), wide.predict_clear(Xw[:16]))
    with pytest.raises(ValueError, match="32-bit wire form"):
        wide.compress_scores(out)
    with pytest.raises(ValueError, match="32-bit wire form"):
        PeerScoreBoard(wide, rows_max=16)


def test_pinned_and_pageable_rows_give_the_same_scores(cuda_dev):
    """predict_host[_seeded] uploads page-locked rows (fhe_b200_host_alloc / _native.pinned_copy) from where they are and
    stages pageable ones; both forms, and a strided (non-contiguous) view, return the clear circuit's scores -- and the
    results land in the caller's arrays although the client kernel stores them zero-copy into the library's buffer."""
    from fhe_icp_b200 import FHESimilarityModel
    from fhe_icp_b200 import _native as N
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=23, verbose=False)
    X, _ = m.train(n_samples=300)
    m.compile(X[:10])
    ref = m.predict_clear(X)
    Xp = N.pinned_copy(X)
    assert Xp.dtype == np.float32 and Xp.shape == X.shape and np.array_equal(Xp, X)
    for fmt in ("seeded", "expanded"):
        m.model.fhe_circuit.ciphertext_format = fmt
        assert np.array_equal(m.predict_encrypted(Xp), ref)
        assert np.array_equal(m.predict_encrypted(X), ref)
        assert np.array_equal(m.predict_encrypted(Xp[::2]), ref[::2])        # a view: made contiguous by the host mirror
        assert np.array_equal(m.predict_encrypted(Xp[7:8]), ref[7:8])        # one row from the middle of a pinned buffer
    del Xp                                                                    # the finalizer frees the pinned allocation
    import gc
    gc.collect()
    assert np.array_equal(m.predict_encrypted(X[:5]), ref[:5])
