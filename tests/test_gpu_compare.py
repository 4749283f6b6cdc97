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


def _pipeline(O, cuda_dev, params, d, B, seed, multibit, with_oracle=True, tol_log2=-16):
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
    scores = ec.scores(ct_q, ct_d)
    got = ec.decrypt(scores)
    assert np.array_equal(got, yq @ xq)                      # exact: the clear integer model
    assert np.array_equal(got, ec.compare_clear(q, docs))
    if with_oracle:
        op = _oparams(O, params)
        os_, oS = O.secret_key(ec.key_seed, 0, op.n), O.secret_key(ec.key_seed, 1, op.k * op.N)
        oq = O.lwe_encrypt(os_, xq, IN_SHIFT, op.sigma_lwe_abs, seed, 0, stride=ct_q.shape[-1])
        od = O.lwe_encrypt(os_, yq, IN_SHIFT, op.sigma_lwe_abs, seed, d, stride=ct_q.shape[-1]).reshape(B, d, -1)
        assert np.array_equal(_u64(ct_q), oq) and np.array_equal(_u64(ct_d), od)   # inputs bit-identical
        if multibit:
            obskf = O.bsk2_to_fourier(op, O.bsk2_gen(op, os_, oS, ec.evk_seed))
        else:
            obskf = O.bsk_to_fourier(op, O.bsk_gen(op, os_, oS, ec.evk_seed))
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
    # l_pbs = 1 (23-bit digits): the per-PBS noise is ~2^-18 at n=24, small sums still decode
    assert _pipeline(O, cuda_dev, TOY_L1, 4, 2, seed=5, multibit=True, tol_log2=-13) < -15


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
