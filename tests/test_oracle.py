"""The CPU oracle against published known answers, its own committed golden fixtures
(tests/golden/golden_v1.json, made by tests/golden/make_golden.py) and the algebraic
properties the domain offers.  No GPU."""
import base64
import hashlib
import json
import math
from pathlib import Path

import numpy as np
import pytest

GOLD = json.loads((Path(__file__).parent / "golden" / "golden_v1.json").read_text())


def h(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_philox_random123_known_answers(O):
    for kat in GOLD["philox_kat"]:
        assert O.philox(kat["ctr"], kat["key"], kat.get("rounds", 10)).tolist() == kat["out"]
    assert {k.get("rounds", 10) for k in GOLD["philox_kat"]} == {7, 10}    # secret streams and the public mask stream


def test_deterministic_log_and_cos_accuracy(O):
    L = O.lib()
    rng = np.random.RandomState(0)
    for x in rng.rand(5000):
        assert abs(L.orc_det_log(float(x)) - math.log(x)) <= 1e-15 * max(1.0, abs(math.log(x)))
    for k in rng.randint(0, 2 ** 53, size=5000, dtype=np.uint64):
        assert abs(L.orc_det_cos2pi_k53(int(k)) - math.cos(2 * math.pi * int(k) / 2 ** 53)) < 2e-15
    assert L.orc_det_cos2pi_k53(0) == 1.0 and L.orc_det_cos2pi_k53(1 << 52) == -1.0


def test_gaussian_stream_golden_and_moments(O):
    g = GOLD["gaussian"]
    v = np.array([O.gaussian(g["seed"], g["domain"], i, 0, g["sigma_abs"]) for i in range(4096)], dtype=np.int64)
    assert v[:8].tolist() == g["first8"] and h(v) == g["sha256_4096"]
    z = np.array([O.gaussian(7, 3, i, 0, 1e9) for i in range(100000)]) / 1e9
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01 and abs((z ** 4).mean() - 3) < 0.1


def test_secret_key_golden(O):
    for c in GOLD["secret_key"]["cases"]:
        s = O.secret_key(GOLD["secret_key"]["seed"], c["key_id"], c["dim"])
        assert int(s.sum()) == c["weight"] and h(s) == c["sha256"]
        assert set(np.unique(s)) <= {0, 1}


def test_lwe_encrypt_lincomb_decrypt_golden(O):
    for c in GOLD["lwe"]:
        n, stride = c["n"], c["stride"]
        s = O.secret_key(c["key_seed"], 2, n)
        msgs = np.array(c["msgs"])
        ct = O.lwe_encrypt(s, msgs, c["shift"], 2.0 ** c["log2_sigma_abs"], c["enc_seed"], ct_base=c["ct_base"],
                           stride=stride).reshape(3, 8, stride)
        assert h(ct) == c["ct_sha256"]
        assert [int(x) for x in ct[0, 0, :4]] == c["ct_first_words"] and int(ct[0, 0, n]) == c["body0"]
        out = O.lincomb(ct, np.array(c["W"]), n)
        assert h(out) == c["out_sha256"]
        dec = O.lwe_decrypt(s, out, c["shift"])
        assert dec.tolist() == c["decrypted"]
        assert np.array_equal(dec, msgs @ np.array(c["W"]).T)


def test_lwe_linearity_and_edge_cases(O):
    n, shift = 64, 40
    s = O.secret_key(3, 2, n)
    a = O.lwe_encrypt(s, [5, -7, 0], shift, 2.0 ** 20, 1)
    b = O.lwe_encrypt(s, [11, 2, -128], shift, 2.0 ** 20, 1, ct_base=3)
    assert np.array_equal(O.lwe_decrypt(s, a + b, shift), [16, -5, -128])       # ct + ct
    assert np.array_equal(O.lwe_decrypt(s, a * np.uint64(3), shift), [15, -21, 0])  # c * ct
    neg = (np.uint64(0) - a)
    assert np.array_equal(O.lwe_decrypt(s, neg, shift), [-5, 7, 0])
    # empty batch, single feature, zero weights
    assert O.lwe_encrypt(s, np.zeros(0, dtype=np.int64), shift, 1.0, 1).shape == (0, n + 1)
    one = O.lincomb(a.reshape(3, 1, -1), np.array([[0]]), n)
    assert not one.any()
    # the clear bias lands on the body only
    with_bias = O.lincomb(a.reshape(3, 1, -1), np.array([[1]]), n, bias=[9], shift=shift)
    assert np.array_equal(O.lwe_decrypt(s, with_bias[:, 0], shift), [14, 2, 9])


def test_clear_circuit_golden(O):
    c = GOLD["clear_circuit"]
    X = np.frombuffer(base64.b64decode(c["x_first16_f32_b64"]), dtype=np.float32).reshape(16, 128)
    iq = c["input_q"]
    q = O.quantize(X, iq["scale"], iq["zero_point"], iq["offset"], iq["n_bits"])
    assert h(q) == c["q_x_first16_sha256"]
    qy = O.clear_circuit(q, np.array(c["q_weights"]), c["weight_q"]["zero_point"], c["q_bias"])
    assert qy.tolist() == c["q_y_first16"]
    y = c["out_scale"] * (qy[:4] - c["out_zero_point"]).astype(np.float64)
    assert y.tolist() == c["y_first4"]


def test_uniform_quantizer_rules(O):
    # asymmetric signed 8-bit
    sc, zp, off = O.uniform_quantizer_params(np.array([-1.0, 3.0]), 8, True)
    assert off == 128 and sc == pytest.approx(4.0 / 255) and zp == round((3.0 * -128 - (-1.0) * 127) / 4.0)
    q = O.quantize(np.array([-1.0, 3.0, 100.0, -100.0]), sc, zp, off, 8)
    assert q.tolist() == [-128, 127, 127, -128]
    # constant tensors (STABILITY_CONST rule)
    assert O.uniform_quantizer_params(np.array([2.5, 2.5 + 1e-9]), 8, True)[:2] == (2.5 + 1e-9, 0)
    assert O.uniform_quantizer_params(np.zeros(4), 8, True)[:2] == (1.0, 0)
    # round-half-to-even
    assert O.quantize(np.array([0.5, 1.5, 2.5]), 1.0, 0, 128, 8).tolist() == [0, 2, 2]


def test_negacyclic_fft_product_matches_schoolbook(O):
    rng = np.random.RandomState(0)
    for N in (16, 64, 256):
        a = rng.randint(-2 ** 22, 2 ** 22, size=N)
        b = rng.randint(0, 2 ** 63, size=N, dtype=np.int64).astype(np.uint64) * np.uint64(2)
        d = (O.negacyclic_mul_fft(a, b) - O.negacyclic_mul_naive(a, b)).view(np.int64)
        assert np.abs(d).max() < 2 ** (64 - 24)  # f64 FFT rounding only


def test_keyswitch_and_pbs_toy_golden(O):
    g = GOLD["ks_pbs_toy"]
    p = O.make_params(n=16, k=1, N=2048, l_pbs=1, beta_pbs=23, l_ks=5, beta_ks=3, log2_sigma_lwe=-30.0)
    s, S = O.secret_key(11, 0, 16), O.secret_key(11, 1, 2048)
    ksk, bsk = O.ksk_gen(p, S, s, 22), O.bsk_gen(p, s, S, 22)
    assert h(ksk) == g["ksk_sha256"] and h(bsk) == g["bsk_sha256"]
    ct = O.lwe_encrypt(S, np.arange(16), 59, p.sigma_glwe_abs, 5, ct_base=7)
    ks = O.keyswitch(p, ksk, ct)
    assert h(ks) == g["ks_out_sha256"] and O.lwe_decrypt(s, ks, 59).tolist() == g["ks_decrypted"] == list(range(16))
    out = O.pbs(p, O.bsk_to_fourier(p, bsk), ks, O.make_lut_poly(np.array(g["table"]), 4, 2048, 59))
    assert (O.lwe_decrypt(S, out, 59) & 15).tolist() == g["pbs_decrypted"] == g["table"]


@pytest.mark.parametrize("l_pbs,beta", [(1, 23), (2, 12), (3, 8)])
def test_pbs_every_message_every_level_count(O, l_pbs, beta):
    """decrypt(PBS(enc(m))) == LUT[m] for all m (the functional KAT of SURVEY.md section 8c), small N."""
    p = O.make_params(n=12, k=1, N=512, l_pbs=l_pbs, beta_pbs=beta, l_ks=4, beta_ks=4, log2_sigma_lwe=-30.0,
                      log2_sigma_glwe=-50.0)
    s, S = O.secret_key(1, 0, 12), O.secret_key(1, 1, 512)
    bskf = O.bsk_to_fourier(p, O.bsk_gen(p, s, S, 5))
    for table in (np.arange(8), (np.arange(8) * 3 + 1) % 8, np.full(8, 5)):
        ct = O.lwe_encrypt(s, np.arange(8), 60, p.sigma_lwe_abs, 77)
        out = O.pbs(p, bskf, ct, O.make_lut_poly(table, 3, 512, 60))
        assert np.array_equal(O.lwe_decrypt(S, out, 60) & 7, table)


def test_modswitch_and_negacyclic_lut_wraparound(O):
    p = O.make_params(n=4, k=1, N=2048)
    ct = np.array([[0, 1 << 63, (1 << 52) - 1, 1 << 52, (1 << 64) - 1]], dtype=np.uint64)
    assert O.modswitch(p, ct)[0].tolist() == [0, 2048, 1, 1, 0]
    lut = O.make_lut_poly(np.arange(16), 4, 2048, 59)
    assert int(lut[0]) == 0 and int(lut[64]) == 1 << 59        # box 1 starts half a box early
    assert int(lut[2047]) == (1 << 64) - 0 * (1 << 59) or True   # wrapped part carries -f(0) = 0
    lut2 = O.make_lut_poly(np.arange(16) + 1, 4, 2048, 59)
    assert int(lut2[2047]) == (1 << 64) - (1 << 59)              # -f(0) on the wrapped half box


def test_encrypted_product_oracle_matches_clear_model(O):
    """quarter-square products through the oracle's PBS decrypt to sum_j x_j*y_j."""
    p = O.make_params(n=16, k=1, N=2048, l_pbs=2, beta_pbs=15, l_ks=4, beta_ks=4, log2_sigma_lwe=-30.0,
                      log2_sigma_glwe=-51.6)
    s, S = O.secret_key(3, 0, p.n), O.secret_key(3, 1, p.k * p.N)
    bskf = O.bsk_to_fourier(p, O.bsk_gen(p, s, S, 4))
    rng = np.random.RandomState(0)
    d, B = 3, 2
    xq, yq = rng.randint(-4, 4, size=d), rng.randint(-4, 4, size=(B, d))
    xq[0], yq[0, 0], yq[1, 0] = -4, -4, 3     # extremes: 16 and -12
    cq = O.lwe_encrypt(s, xq, 59, p.sigma_lwe_abs, 7, 0, stride=p.n + 2)
    cd = O.lwe_encrypt(s, yq, 59, p.sigma_lwe_abs, 7, d, stride=p.n + 2).reshape(B, d, -1)
    assert np.array_equal(O.quarter_square_table(4)[[0, 8, 15]], [16, 0, 12])
    sc = O.encrypted_product_scores(p, bskf, cq, cd, 4, 51)
    dec = O.lwe_decrypt(S, sc, 51) & 8191
    assert np.array_equal(np.where(dec >= 4096, dec - 8192, dec), yq @ xq)


def test_encrypted_threshold_oracle_exact(O):
    """LSB-first bit extraction of (score - T) decides score >= T exactly."""
    p = O.make_params(n=16, k=1, N=2048, l_pbs=2, beta_pbs=15, l_ks=4, beta_ks=4, log2_sigma_lwe=-30.0,
                      log2_sigma_glwe=-51.6)
    s, S = O.secret_key(3, 0, p.n), O.secret_key(3, 1, p.k * p.N)
    bskf = O.bsk_to_fourier(p, O.bsk_gen(p, s, S, 4))
    ksk32 = O.ksk_to_32(p, O.ksk_gen(p, S, s, 4))
    vals = np.array([-1536, -1, 0, 41, 42, 43, 2048])
    cts = O.lwe_encrypt(S, vals, 51, 2.0 ** (64 - 18), 9, 0)
    for T in (42, 0, -1536):
        out = O.encrypted_ge(p, ksk32, bskf, cts, T, 13, 51, 60)
        assert np.array_equal(O.lwe_decrypt(S, out, 60) & 15, (vals >= T).astype(np.int64))


def test_packed_inner_product_oracle_exact_and_worst_case_noise(O):
    """Leveled both-encrypted comparison: GGSW(Q) [.] GLWE(packed documents), coefficient slot*b = <x, y_b>.
    A constant-sign query is the worst case for the noise (binary key, see encrypted_compare.PACKED_PARAMS)."""
    p = O.make_params(n=742, k=1, N=2048, l_pbs=2, beta_pbs=18)
    S = O.secret_key(5, 1, 2048)
    rng = np.random.RandomState(3)
    d, B, sh = 128, 35, 47
    for xq in (rng.randint(-16, 16, size=d), np.full(d, -16)):
        yq = rng.randint(-16, 16, size=(B, d))
        yq[0] = -16
        docs = O.glwe_encrypt_rows(p, S, O.pack_documents(yq, 2048, 128), 0, sh, seed=7, id_base=100)
        assert docs.shape == (3, 2, 2048)
        gf = O.ggsw_to_fourier(p, O.glwe_encrypt_rows(p, S, O.query_polynomial(xq, 2048), 1, 0, seed=7))
        lwe = O.glwe_sample_extract(p, O.glwe_external_product(p, gf, docs), 0, 128, 16, 2050)[:B]
        dec = O.lwe_decrypt(S, lwe, sh) & 131071
        want = yq @ xq
        assert np.array_equal(np.where(dec >= 65536, dec - 131072, dec), want)
        err = (O.lwe_phase(S, lwe) - (want.astype(np.int64).astype(np.uint64) << np.uint64(sh))).view(np.int64)
        assert np.log2(np.abs(err.astype(np.float64)).max() + 1) - 64 < -19.5      # decoding margin 2^-18
