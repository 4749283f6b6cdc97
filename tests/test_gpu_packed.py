"""Packed both-encrypted comparison (GLWE x GGSW external product, no bootstrap) on the GPU against the
CPU oracle and the clear integer model sum_j xq_j*yq_j.  Fresh GLWE / GGSW ciphertexts are bit-identical
to the oracle's; products agree within the f64-FFT rounding bound; decrypted scores are exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _u64(t):
    return t.detach().cpu().numpy().view(np.uint64)


def _op(O, pd):
    return O.make_params(n=pd["n"], k=pd["k"], N=pd["N_poly"], l_pbs=pd["l_pbs"], beta_pbs=pd["beta_pbs"],
                         l_ks=pd["l_ks"], beta_ks=pd["beta_ks"], log2_sigma_lwe=pd["log2_sigma_lwe"],
                         log2_sigma_glwe=pd["log2_sigma_glwe"])


@pytest.mark.parametrize("B,d", [(1, 128), (16, 128), (37, 128), (5, 7), (3, 2048), (40, 100)])
def test_packed_scores_exact_and_match_oracle(O, cuda_dev, B, d):
    from fhe_icp_b200 import engine as E
    from fhe_icp_b200.encrypted_compare import PACKED_OUT_SHIFT, PackedEncryptedCompare
    rng = np.random.RandomState(B * 31 + d)
    if d == 2048:    # the score range limits d at full-scale factors; use small factors for the full-width case
        xq, yq = rng.randint(-2, 2, size=d), rng.randint(-2, 2, size=(B, d))
        import fhe_icp_b200.encrypted_compare as M
        pe = PackedEncryptedCompare.__new__(PackedEncryptedCompare)
        pe.d, pe.pd = d, dict(M.PACKED_PARAMS); pe.p = E.make_pbs_params(**pe.pd); pe.slot, pe.per = 2048, 1
        pe.dev, pe.key_seed, pe.scale, pe.S = cuda_dev, 5, 1.0, None
        pe.noise_seed, pe.enc_seed, pe.ids = 4242, 9, M.CiphertextIds(0)
        pe.keygen()
    else:
        pe = PackedEncryptedCompare(input_dim=d, device=cuda_dev, key_seed=5).keygen()
        xq, yq = rng.randint(-16, 16, size=d), rng.randint(-16, 16, size=(B, d))
        xq[0], yq[0, 0] = -16, -16
    gq = pe.encrypt_query(xq, enc_seed=9, id_base=0)
    gd = pe.encrypt_documents(yq, enc_seed=9, id_base=1 << 20)
    assert gd.shape == ((B + pe.per - 1) // pe.per, 2, 2048)
    prod = pe.scores(gq, gd)
    got = pe.decrypt(prod, B)
    assert np.array_equal(got, yq @ xq)
    # oracle: identical fresh ciphertexts, products within FFT rounding, identical decryptions
    op = _op(O, pe.pd)
    oS = O.secret_key(5, 1, 2048)
    odocs = O.glwe_encrypt_rows(op, oS, O.pack_documents(yq, 2048, pe.slot), 0, PACKED_OUT_SHIFT, 9, 1 << 20,
                                noise_seed=pe.noise_seed)
    assert np.array_equal(_u64(gd), odocs)
    oggsw = O.glwe_encrypt_rows(op, oS, O.query_polynomial(xq, 2048), 1, 0, 9, 0, noise_seed=pe.noise_seed)
    ogf = O.ggsw_to_fourier(op, oggsw)
    assert np.abs(gq.cpu().numpy() - ogf).max() / np.abs(ogf).max() < 1e-13
    oprod = O.glwe_external_product(op, ogf, odocs)
    diff = (_u64(prod) - oprod).view(np.int64).astype(np.float64)
    assert np.log2(np.abs(diff).max() + 1) - 64 < -25          # the same product up to f64 FFT rounding (18-bit digits)
    lwe = pe.scores_as_lwe(prod)
    olwe = O.glwe_sample_extract(op, _u64(prod), 0, pe.slot, pe.per, lwe.shape[1])
    assert np.array_equal(_u64(lwe), olwe)
    dec = O.lwe_decrypt(oS, olwe, PACKED_OUT_SHIFT)[:B] & 131071
    assert np.array_equal(np.where(dec >= 65536, dec - 131072, dec), got)
    # residual noise against the decoding margin 2^-18
    ph = O.lwe_phase(oS, olwe)[:B]
    err = (ph - (got.astype(np.int64).astype(np.uint64) << np.uint64(PACKED_OUT_SHIFT))).view(np.int64).astype(np.float64)
    assert np.log2(np.abs(err).max() + 1) - 64 < -20.0


@pytest.mark.parametrize("xval,yval", [(-16, -16), (15, -16), (-16, 15), (15, 15)])
def test_packed_worst_case_constant_sign_vectors(O, cuda_dev, xval, yval):
    """The binary key's non-zero mean makes a constant-sign query the worst case for the external-product
    noise ((sum_j Q_j)^2 = d^2 * 2^8).  Exactness and the stated noise bound must hold there too."""
    from fhe_icp_b200.encrypted_compare import PACKED_OUT_SHIFT, PackedEncryptedCompare
    d, B = 128, 48
    pe = PackedEncryptedCompare(input_dim=d, device=cuda_dev, key_seed=5).keygen()
    xq = np.full(d, xval)
    yq = np.full((B, d), yval)
    yq[1::2] = np.random.RandomState(0).randint(-16, 16, size=(B // 2, d))
    prod = pe.scores(pe.encrypt_query(xq, 3), pe.encrypt_documents(yq, 3))
    got = pe.decrypt(prod, B)
    assert np.array_equal(got, yq @ xq) and abs(int(got[0])) == abs(xval * yval) * d
    lwe = _u64(pe.scores_as_lwe(prod))[:B]
    ph = O.lwe_phase(O.secret_key(5, 1, 2048), lwe)
    err = (ph - (got.astype(np.int64).astype(np.uint64) << np.uint64(PACKED_OUT_SHIFT))).view(np.int64).astype(np.float64)
    assert np.log2(err.std()) - 64 < -21.5 and np.log2(np.abs(err).max() + 1) - 64 < -19.5   # margin: 2^-18


def test_packed_large_batch_and_float_pipeline(cuda_dev):
    from fhe_icp_b200.encrypted_compare import PackedEncryptedCompare
    rng = np.random.RandomState(2)
    d, B = 128, 16 * 148 * 3 + 5            # more ciphertexts than one persistent wave, ragged tail
    q = rng.randn(d); q /= np.linalg.norm(q)
    docs = rng.randn(B, d)
    docs[::7] = 0.8 * q + 0.6 * docs[::7] / np.sqrt(d)
    docs /= np.linalg.norm(docs, axis=1, keepdims=True)
    pe = PackedEncryptedCompare(input_dim=d, device=cuda_dev).keygen()
    pe.fit_scale(np.array([-1.0, 1.0]) / np.sqrt(d))
    sim = pe.similarity(q, docs)
    assert np.array_equal(sim, pe.dequantize(pe.compare_clear(q, docs)))
    cos = docs @ q
    assert np.abs(sim - cos).max() < 0.05          # 5-bit factors
    hits = np.flatnonzero(sim >= 0.5)
    assert set(hits) == set(range(0, B, 7))


def test_packed_scores_feed_the_encrypted_threshold(cuda_dev):
    """Packed scores -> sample extraction -> exact encrypted threshold (15-bit scores)."""
    from fhe_icp_b200.encrypted_compare import (PACKED_OUT_SHIFT, PACKED_SCORE_BITS, EncryptedCompare, EncryptedThreshold,
                                                PackedEncryptedCompare)
    rng = np.random.RandomState(6)
    d, B = 128, 21
    pe = PackedEncryptedCompare(input_dim=d, device=cuda_dev).keygen()
    xq, yq = rng.randint(-16, 16, size=d), rng.randint(-16, 16, size=(B, d))
    yq[3] = xq
    lwe = pe.scores_as_lwe(pe.scores(pe.encrypt_query(xq, 4), pe.encrypt_documents(yq, 4)))[:B]
    ints = yq @ xq
    ec = EncryptedCompare(input_dim=d, device=cuda_dev, key_seed=pe.key_seed).keygen()   # same big key S (key id 1)
    th = EncryptedThreshold(ec, score_bits=PACKED_SCORE_BITS, out_shift=PACKED_OUT_SHIFT)
    for T in (int(ints[3]), int(ints[3]) + 1, 0):
        assert np.array_equal(th.decrypt(th.ge(lwe.contiguous(), T)), (ints >= T).astype(np.int64)), T
