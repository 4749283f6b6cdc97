"""Encrypted x encrypted comparison on the GPU against the CPU oracle and against
the clear integer model sum_j xq_j*yq_j it must reproduce exactly.  The reference has no such path
(it multiplies the embeddings in the clear, batch_operations.py:226,273); parity is anchored on the
clear product it computes there: decrypt(scores) == quantized dot product, for every document."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOY_L2 = dict(n=20, k=1, N_poly=2048, l_pbs=2, beta_pbs=15, l_ks=4, beta_ks=4, log2_sigma_lwe=-30.0, log2_sigma_glwe=-51.6)
TOY_L1 = dict(n=24, k=1, N_poly=2048, l_pbs=1, beta_pbs=23, l_ks=5, beta_ks=3, log2_sigma_lwe=-30.0, log2_sigma_glwe=-51.6)


def _u64(t):
    return t.detach().cpu().numpy().view(np.uint64)


def _oparams(O, d):
    return O.make_params(n=d["n"], k=d["k"], N=d["N_poly"], l_pbs=d["l_pbs"], beta_pbs=d["beta_pbs"], l_ks=d["l_ks"],
                         beta_ks=d["beta_ks"], log2_sigma_lwe=d["log2_sigma_lwe"], log2_sigma_glwe=d["log2_sigma_glwe"])


@pytest.mark.parametrize("B,d,words,stride", [(1, 1, 7, 8), (3, 5, 21, 22), (2, 128, 743, 744), (0, 4, 9, 10)])
def test_pair_glue_bit_exact(O, cuda_dev, B, d, words, stride):
    import torch
    from fhe_icp_b200 import engine as E
    rng = np.random.RandomState(B * 7 + d)
    q = rng.randint(0, 2**63, size=(d, stride)).astype(np.uint64) * np.uint64(2) + np.uint64(1)
    y = rng.randint(0, 2**63, size=(B, d, stride)).astype(np.uint64) * np.uint64(3)
    off = 1 << 62
    got = E.pair_addsub(E.from_u64_numpy(q, cuda_dev), E.from_u64_numpy(y, cuda_dev), words, off)
    ref = O.pair_addsub(q, y, words, off)
    assert got.shape == (B, d, 2, words)
    assert np.array_equal(_u64(got).reshape(ref.shape), ref)
    if B:
        s = E.pair_diff_sum(got)
        assert s.shape[1] % 2 == 0 and s.shape[1] >= words
        assert np.array_equal(_u64(s)[:, :words], O.pair_diff_sum(ref))
        assert not _u64(s)[:, words:].any()


def _pipeline(O, cuda_dev, params, d, B, seed, multibit, with_oracle=True, tol_log2=-16, norms=False):
    from fhe_icp_b200.encrypted_compare import EncryptedCompare, IN_SHIFT, OUT_SHIFT, P_BITS
    rng = np.random.RandomState(seed)
    ec = EncryptedCompare(input_dim=d, params=params, device=cuda_dev, multibit=multibit).keygen()
    q = rng.randn(d) / np.sqrt(d)
    docs = rng.randn(B, d) / np.sqrt(d)
    docs[0] = q  # identical vectors: the largest score
    ec.fit_scale(np.concatenate([q[None], docs]))
    xq, yq = ec.quantize(q), ec.quantize(docs)
    assert xq.min() >= -4 and xq.max() <= 3
    ct_q = ec.encrypt(xq, enc_seed=seed, ct_base=0)
    ct_d = ec.encrypt(yq, enc_seed=seed, ct_base=d)
    if norms:
        n_q, n_d = ec.encrypt_norms(xq, seed, 0), ec.encrypt_norms(yq, seed, 1)
        scores = ec.scores(ct_q, ct_d, n_q, n_d)
    else:
        scores = ec.scores(ct_q, ct_d)
    got = ec.decrypt(scores)
    assert np.array_equal(got, yq @ xq)                      # exact: the clear integer model
    assert np.array_equal(got, ec.compare_clear(q, docs))
    if with_oracle:
        op = _oparams(O, params)
        os_, oS = O.secret_key(ec.key_seed, 0, op.n), O.secret_key(ec.key_seed, 1, op.k * op.N)
        ns = ec.noise_seed       # client secret (OS CSPRNG); the oracle mirrors it to compare ciphertext words
        oq = O.lwe_encrypt(os_, xq, IN_SHIFT, op.sigma_lwe_abs, seed, 0, stride=ct_q.shape[-1], noise_seed=ns)
        od = O.lwe_encrypt(os_, yq, IN_SHIFT, op.sigma_lwe_abs, seed, d, stride=ct_q.shape[-1], noise_seed=ns).reshape(B, d, -1)
        assert np.array_equal(_u64(ct_q), oq) and np.array_equal(_u64(ct_d), od)   # inputs bit-identical
        if multibit:
            obskf = O.bsk2_to_fourier(op, O.bsk2_gen(op, os_, oS, ec.evk_seed))
        else:
            obskf = O.bsk_to_fourier(op, O.bsk_gen(op, os_, oS, ec.evk_seed))
        if norms:
            from fhe_icp_b200.encrypted_compare import NORM_CT_BASE
            onq = O.lwe_encrypt(oS, (xq * xq).sum(), OUT_SHIFT - 1, op.sigma_glwe_abs, seed, NORM_CT_BASE, stride=n_q.shape[-1],
                                noise_seed=ns)
            ond = O.lwe_encrypt(oS, (yq * yq).sum(axis=1), OUT_SHIFT - 1, op.sigma_glwe_abs, seed, NORM_CT_BASE + 1,
                                stride=n_q.shape[-1], noise_seed=ns)
            assert np.array_equal(_u64(n_q).reshape(onq.shape), onq) and np.array_equal(_u64(n_d), ond)
            ref = O.encrypted_product_scores_norms(op, obskf, oq, od, onq, ond, P_BITS, OUT_SHIFT, multibit=multibit)
        else:
            ref = O.encrypted_product_scores(op, obskf, oq, od, P_BITS, OUT_SHIFT, multibit=multibit)
        words = op.k * op.N + 1
        dec = O.lwe_decrypt(oS, ref, OUT_SHIFT) & 8191
        assert np.array_equal(np.where(dec >= 4096, dec - 8192, dec), got)
        diff = (O.lwe_phase(oS, _u64(scores)[:, :words]) - O.lwe_phase(oS, ref)).view(np.int64).astype(np.float64)
        assert np.log2(np.abs(diff).max() + 1) - 64 < tol_log2
    # residual noise of the score ciphertext against its decoding margin 2^-14
    ph = O.lwe_phase(O.secret_key(ec.key_seed, 1, ec.p.k * ec.p.N), _u64(scores)[:, : ec.p.k * ec.p.N + 1])
    err = (ph - (got.astype(np.int64).astype(np.uint64) << np.uint64(OUT_SHIFT))).view(np.int64).astype(np.float64)
    return np.log2(np.abs(err).max() + 1) - 64


@pytest.mark.parametrize("B,d", [(1, 1), (3, 5), (2, 16)])
def test_toy_two_levels(O, cuda_dev, B, d):
    assert _pipeline(O, cuda_dev, TOY_L2, d, B, seed=B + d, multibit=False) < -15


def test_toy_multibit_single_level(O, cuda_dev):
    # l_pbs = 1 (23-bit digits): one PBS output carries ~2^-17.5 at n=24, so only short sums stay inside
    # the decoding margin 2^-14 (4 outputs: std 2^-16.5) -- the reason the production set uses l_pbs = 2
    assert _pipeline(O, cuda_dev, TOY_L1, 2, 2, seed=5, multibit=True, tol_log2=-13) < -14


@pytest.mark.parametrize("B,d,multibit", [(1, 1, False), (3, 7, True), (2, 16, True)])
def test_toy_one_bootstrap_per_dimension(O, cuda_dev, B, d, multibit):
    assert _pipeline(O, cuda_dev, TOY_L2, d, B, seed=B + d, multibit=multibit, norms=True) < -15


def test_full_parameter_set_d128_norm_protocol(O, cuda_dev):
    """Default protocol at the production set: d = 128 bootstraps per document, one document via the oracle."""
    from fhe_icp_b200.encrypted_compare import COMPARE_PARAMS
    assert _pipeline(O, cuda_dev, COMPARE_PARAMS, 128, 1, seed=31, multibit=True, tol_log2=-15, norms=True) < -15.5
    assert _pipeline(O, cuda_dev, COMPARE_PARAMS, 128, 50, seed=32, multibit=True, with_oracle=False, norms=True) < -15.5


def test_toy_multibit_two_levels(O, cuda_dev):
    assert _pipeline(O, cuda_dev, TOY_L2, 6, 3, seed=8, multibit=True) < -15


def test_full_parameter_set_d128_multibit(O, cuda_dev):
    """The production configuration: two-level multi-bit blind rotation."""
    from fhe_icp_b200.encrypted_compare import COMPARE_PARAMS
    assert _pipeline(O, cuda_dev, COMPARE_PARAMS, 128, 37, seed=21, multibit=True, with_oracle=False) < -15.5


def test_full_parameter_set_d128(O, cuda_dev):
    """n=742, N=2048, l=2, d=128: 256 bootstraps per document; one document also through the oracle."""
    from fhe_icp_b200.encrypted_compare import COMPARE_PARAMS
    worst = _pipeline(O, cuda_dev, COMPARE_PARAMS, 128, 1, seed=11, multibit=False, with_oracle=True, tol_log2=-15)
    assert worst < -15.5
    worst = _pipeline(O, cuda_dev, COMPARE_PARAMS, 128, 40, seed=12, multibit=False, with_oracle=False)
    assert worst < -15.5   # 2.8x inside the decoding margin 2^-14 at the worst document


def test_float_pipeline_tracks_cosine(cuda_dev):
    from fhe_icp_b200.encrypted_compare import EncryptedCompare, COMPARE_PARAMS
    rng = np.random.RandomState(3)
    d, B = 128, 24
    q = rng.randn(d); q /= np.linalg.norm(q)
    docs = rng.randn(B, d)
    docs[: B // 2] = 0.8 * q + 0.6 * docs[: B // 2] / np.sqrt(d)   # half correlated with the query
    docs /= np.linalg.norm(docs, axis=1, keepdims=True)
    ec = EncryptedCompare(input_dim=d, params=COMPARE_PARAMS, device=cuda_dev).keygen()
    ec.fit_scale(np.concatenate([q[None], docs]))
    sim = ec.similarity(q, docs)
    cos = docs @ q
    assert np.abs(sim - cos).max() < 0.12          # 3-bit factors: coarse but unbiased enough to rank
    assert set(np.argsort(-sim)[: B // 2]) == set(range(B // 2))


# ---------------------------------------------------------------------------------------------- encrypted threshold (N3)
def _score_cts(ec, values):
    """Big-key encryptions of given integer scores at the score encoding (test fixture: stands in for
    the output of EncryptedCompare.scores, with the PBS-sum noise level 2^-18)."""
    import torch
    from fhe_icp_b200 import engine as E
    from fhe_icp_b200.encrypted_compare import OUT_SHIFT
    return E.lwe_encrypt(ec.S, torch.as_tensor(np.asarray(values, dtype=np.int64)), OUT_SHIFT, 2.0 ** (64 - 18),
                         enc_seed=77, ct_base=0)


@pytest.mark.parametrize("params_name,multibit", [("toy", False), ("toy", True), ("full", True)])
def test_encrypted_threshold_exact(O, cuda_dev, params_name, multibit):
    """(score >= T) under encryption equals the clear comparison for every score around every
    threshold, including score == T and the extremes of the 13-bit range; GPU == oracle."""
    from fhe_icp_b200 import engine as E
    from fhe_icp_b200.encrypted_compare import (BIT_SHIFT, COMPARE_PARAMS, OUT_SHIFT, SCORE_BITS, EncryptedCompare,
                                                EncryptedThreshold)
    params = TOY_L2 if params_name == "toy" else COMPARE_PARAMS
    ec = EncryptedCompare(input_dim=128, params=params, device=cuda_dev, multibit=multibit).keygen()
    th = EncryptedThreshold(ec)
    vals = np.array([-1536, -1, 0, 1, 2, 165, 166, 167, 255, 256, 257, 1023, 1024, 2047, 2048, -300, 777])
    cts = _score_cts(ec, vals)
    for T in ([166, 0] if params_name == "full" else [166, 0, -1536, 2047, 1, 256, -299]):
        got = th.decrypt(th.ge(cts, T))
        assert np.array_equal(got, (vals >= T).astype(np.int64)), (T, got)
    if params_name == "toy":
        op = _oparams(O, params)
        os_, oS = O.secret_key(ec.key_seed, 0, op.n), O.secret_key(ec.key_seed, 1, op.k * op.N)
        ksk32 = O.ksk_to_32(op, O.ksk_gen(op, oS, os_, ec.evk_seed))
        assert np.array_equal(th.ksk32.cpu().numpy().view(np.uint32).reshape(ksk32.shape), ksk32)
        bskf = (O.bsk2_to_fourier(op, O.bsk2_gen(op, os_, oS, ec.evk_seed)) if multibit
                else O.bsk_to_fourier(op, O.bsk_gen(op, os_, oS, ec.evk_seed)))
        ref = O.encrypted_ge(op, ksk32, bskf, _u64(cts), 166, SCORE_BITS, OUT_SHIFT, BIT_SHIFT, multibit=multibit)
        assert np.array_equal(O.lwe_decrypt(oS, ref, BIT_SHIFT) & 15, (vals >= 166).astype(np.int64))
    b = th.decrypt(th.buckets(cts, [166, 256, 1024]))
    assert np.array_equal(b, (vals >= 166).astype(int) + (vals >= 256) + (vals >= 1024))


def test_encrypted_threshold_on_real_scores(cuda_dev):
    """End to end: both vectors encrypted -> encrypted scores -> encrypted (score >= min_similarity)."""
    from fhe_icp_b200.encrypted_compare import EncryptedCompare, EncryptedThreshold
    rng = np.random.RandomState(4)
    d, B = 128, 12
    q = rng.randn(d); q /= np.linalg.norm(q)
    docs = rng.randn(B, d)
    docs[:5] = 0.85 * q + 0.5 * docs[:5] / np.sqrt(d)
    docs /= np.linalg.norm(docs, axis=1, keepdims=True)
    ec = EncryptedCompare(input_dim=d, device=cuda_dev).keygen()
    ec.fit_scale(np.array([-1.0, 1.0]) / np.sqrt(d))
    sc = ec.scores(ec.encrypt(ec.quantize(q), 3, 0), ec.encrypt(ec.quantize(docs), 3, d))
    ints = ec.decrypt(sc)
    th = EncryptedThreshold(ec)
    T = th.threshold_to_int(0.5)
    hit = th.decrypt(th.ge(sc, T))
    assert np.array_equal(hit, (ints >= T).astype(np.int64))
    assert np.array_equal(hit.astype(bool), ec.dequantize(ints) >= 0.5) and hit[:5].all() and not hit[5:].any()
