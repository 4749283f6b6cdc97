"""Keyswitch + programmable bootstrap on the GPU against the CPU oracle.  Integer stages
(key generation, keyswitch) are bit-exact; the f64-FFT blind rotation is compared on the
decrypted table value (exact) and on the ciphertext phase (within a stated noise bound).
The reference has no test for these primitives (its circuit has no table lookup): parity is
anchored on decrypt(PBS(enc(m))) == LUT[m] for every m."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOY = dict(n=24, k=1, N_poly=2048, l_pbs=1, beta_pbs=23, l_ks=5, beta_ks=3, log2_sigma_lwe=-30.0, log2_sigma_glwe=-51.6)
TOY_L2 = dict(n=20, k=1, N_poly=2048, l_pbs=2, beta_pbs=15, l_ks=4, beta_ks=4, log2_sigma_lwe=-30.0, log2_sigma_glwe=-51.6)
P4 = dict(n=742, k=1, N_poly=2048, l_pbs=1, beta_pbs=23, l_ks=5, beta_ks=3, log2_sigma_lwe=-17.1, log2_sigma_glwe=-51.6)


def _u64(t):
    return t.detach().cpu().numpy().view(np.uint64)


def _oparams(O, d):
    return O.make_params(n=d["n"], k=d["k"], N=d["N_poly"], l_pbs=d["l_pbs"], beta_pbs=d["beta_pbs"], l_ks=d["l_ks"],
                         beta_ks=d["beta_ks"], log2_sigma_lwe=d["log2_sigma_lwe"], log2_sigma_glwe=d["log2_sigma_glwe"])


class Keys:
    def __init__(self, O, dev, d, key_seed=11, evk_seed=22):
        from fhe_icp_b200 import engine as E
        self.p = E.make_pbs_params(**d)
        self.op = _oparams(O, d)
        self.s = E.secret_key(key_seed, 0, d["n"], dev)
        self.S = E.secret_key(key_seed, 1, d["k"] * d["N_poly"], dev)
        self.os = O.secret_key(key_seed, 0, d["n"])
        self.oS = O.secret_key(key_seed, 1, d["k"] * d["N_poly"])
        self.ksk = E.ksk_gen(self.p, self.S, self.s, evk_seed)
        self.bsk = E.bsk_gen(self.p, self.s, self.S, evk_seed)
        self.bskf = E.bsk_to_fourier(self.p, self.bsk)
        self.evk_seed = evk_seed


@pytest.fixture(scope="module")
def toy(O, cuda_dev):
    return Keys(O, cuda_dev, TOY)


@pytest.fixture(scope="module")
def p4(O, cuda_dev):
    return Keys(O, cuda_dev, P4)


@pytest.mark.parametrize("which", ["toy", "p4"])
def test_evaluation_keys_bit_exact(O, request, which):
    K = request.getfixturevalue(which)
    assert np.array_equal(K.s.cpu().numpy(), K.os) and np.array_equal(K.S.cpu().numpy(), K.oS)
    assert np.array_equal(_u64(K.ksk), O.ksk_gen(K.op, K.oS, K.os, K.evk_seed))
    obsk = O.bsk_gen(K.op, K.os, K.oS, K.evk_seed)
    assert np.array_equal(_u64(K.bsk), obsk)
    of = O.bsk_to_fourier(K.op, obsk)
    gf = K.bskf.cpu().numpy()
    scale = np.abs(of).max()
    assert np.abs(gf - of).max() / scale < 1e-13


@pytest.mark.parametrize("which,B", [("toy", 1), ("toy", 9), ("p4", 1), ("p4", 33)])
def test_keyswitch_bit_exact(O, request, cuda_dev, which, B):
    import torch
    from fhe_icp_b200 import engine as E
    K = request.getfixturevalue(which)
    rng = np.random.RandomState(B)
    msgs = rng.randint(0, 16, size=B)
    ct = E.lwe_encrypt(K.S, torch.as_tensor(msgs), 59, K.op.sigma_glwe_abs, enc_seed=5, ct_base=7,
                       stride=K.p.k * K.p.N + 2)[:, : K.p.k * K.p.N + 1].contiguous()
    out = E.keyswitch(K.p, K.ksk, ct)
    ref = O.keyswitch(K.op, _u64(K.ksk), _u64(ct))
    assert np.array_equal(_u64(out), ref)
    assert np.array_equal(O.lwe_decrypt(K.os, ref, 59), msgs)


def _pbs_case(O, K, cuda_dev, B, table, seed, tol_log2):
    import torch
    from fhe_icp_b200 import engine as E
    p_bits = 4
    shift = 63 - p_bits
    rng = np.random.RandomState(seed)
    msgs = rng.randint(0, 16, size=B)
    msgs[: min(B, 16)] = np.arange(16)[: min(B, 16)]
    ct = E.lwe_encrypt(K.s, torch.as_tensor(msgs), shift, K.op.sigma_lwe_abs, enc_seed=seed, ct_base=100,
                       stride=K.p.n + 1 if (K.p.n + 1) % 2 == 0 else K.p.n + 2)[:, : K.p.n + 1].contiguous()
    lut = E.make_lut_poly(table, p_bits, K.p.N, shift)
    assert np.array_equal(lut, O.make_lut_poly(table, p_bits, K.p.N, shift))
    out = E.pbs(K.p, K.bskf, ct, E.from_u64_numpy(lut, cuda_dev))
    got = _u64(out)
    dec = O.lwe_decrypt(K.oS, got, shift)
    assert np.array_equal(dec & 15, np.asarray(table)[msgs] & 15)
    # phase agreement with the oracle's own f64-FFT evaluation of the same ciphertexts
    ref = O.pbs(K.op, O.bsk_to_fourier(K.op, _u64(K.bsk)), _u64(ct), lut)
    diff = (O.lwe_phase(K.oS, got) - O.lwe_phase(K.oS, ref)).view(np.int64).astype(np.float64)
    assert np.log2(np.abs(diff).max() + 1) - 64 < tol_log2
    return got


@pytest.mark.parametrize("B", [1, 2, 16, 37])
def test_pbs_toy_all_messages(O, toy, cuda_dev, B):
    # GPU and oracle run different f64 FFT schedules.  Their accumulators differ by FFT rounding
    # (~2^-21 of the torus per CMux at beta=23), which flips some 23-bit digit roundings in the next
    # CMux, so the two outputs are independent realisations of the gadget rounding noise
    # (std ~2^-19.8 per CMux): the phase difference is bounded by the PBS noise bound, not by
    # FFT rounding alone.  The decoding margin is 2^-6, so decrypted values are identical.
    _pbs_case(O, toy, cuda_dev, B, (np.arange(16) * 7 + 3) % 16, seed=B, tol_log2=-14)


def test_pbs_two_levels(O, cuda_dev):
    K = Keys(O, cuda_dev, TOY_L2)
    _pbs_case(O, K, cuda_dev, 16, (np.arange(16) * 5 + 1) % 16, seed=3, tol_log2=-22)


@pytest.mark.parametrize("B,table", [(16, list(range(16))), (16, [(3 * m * m + 1) % 16 for m in range(16)]),
                                     (200, [(m + 5) % 16 for m in range(16)])])
def test_pbs_stated_parameter_set(O, p4, cuda_dev, B, table):
    """n=742, N=2048, l=1, beta=23: every message maps to LUT[m]; the GPU and oracle phases
    agree to 2^-12 of the torus (each output carries the set's PBS noise, std ~2^-15.5, with
    independent rounding realisations; the decoding margin is 2^-6)."""
    got = _pbs_case(O, p4, cuda_dev, B, table, seed=B + 1, tol_log2=-12)
    # output noise against the analytic bound (n * (1 + kN/2) * 2^-2*beta / 12 dominates)
    msgs_dec = O.lwe_decrypt(p4.oS, got, 59)
    err = (O.lwe_phase(p4.oS, got) - (msgs_dec.astype(np.uint64) << np.uint64(59))).view(np.int64).astype(np.float64)
    assert np.log2(err.std()) - 64 < -14.0


def test_keyswitch_then_pbs_atomic_pattern(O, p4, cuda_dev):
    """Concrete's atomic pattern: big-key ciphertext -> keyswitch -> PBS -> big-key ciphertext."""
    import torch
    from fhe_icp_b200 import engine as E
    K = p4
    msgs = np.arange(16)
    ct = E.lwe_encrypt(K.S, torch.as_tensor(msgs), 59, K.op.sigma_glwe_abs, enc_seed=9, ct_base=0,
                       stride=K.p.N + 2)[:, : K.p.N + 1].contiguous()
    table = [(m * m) % 16 for m in range(16)]
    lut = E.from_u64_numpy(E.make_lut_poly(table, 4, K.p.N, 59), cuda_dev)
    out = E.pbs(K.p, K.bskf, E.keyswitch(K.p, K.ksk, ct), lut)
    assert np.array_equal(O.lwe_decrypt(K.oS, _u64(out), 59) & 15, np.asarray(table))
    # and again on the bootstrapped outputs (noise was reset, not accumulated)
    out2 = E.pbs(K.p, K.bskf, E.keyswitch(K.p, K.ksk, out), lut)
    assert np.array_equal(O.lwe_decrypt(K.oS, _u64(out2), 59) & 15, np.asarray(table)[np.asarray(table)])


def test_pbs_per_ciphertext_lut_index(O, toy, cuda_dev):
    import torch
    from fhe_icp_b200 import engine as E
    K = toy
    msgs = np.arange(16)
    ct = E.lwe_encrypt(K.s, torch.as_tensor(msgs), 59, K.op.sigma_lwe_abs, enc_seed=4, ct_base=0,
                       stride=K.p.n + 2 - (K.p.n % 2))[:, : K.p.n + 1].contiguous()
    t0, t1 = np.arange(16), (15 - np.arange(16))
    luts = np.stack([E.make_lut_poly(t0, 4, K.p.N, 59), E.make_lut_poly(t1, 4, K.p.N, 59)])
    idx = np.arange(16) % 2
    out = E.pbs(K.p, K.bskf, ct, E.from_u64_numpy(luts, cuda_dev), torch.as_tensor(idx.astype(np.int32)))
    exp = np.where(idx == 0, t0[msgs], t1[msgs])
    assert np.array_equal(O.lwe_decrypt(K.oS, _u64(out), 59) & 15, exp)


@pytest.mark.parametrize("which,B", [("toy", 5), ("p4", 1), ("p4", 40)])
def test_keyswitch32_bit_exact_and_correct(O, request, cuda_dev, which, B):
    """32-bit keyswitch (key rounded to 32 torus bits): bit-exact vs the oracle's KS32, decrypts to the
    message, and its output differs from the 64-bit keyswitch by ~2^-21.6 (std) of the torus -- the
    rounded key's error over kN*l*(n/2) terms -- against a small-key noise of 2^-17.1."""
    import torch
    from fhe_icp_b200 import engine as E
    K = request.getfixturevalue(which)
    msgs = np.random.RandomState(B).randint(0, 16, size=B)
    ct = E.lwe_encrypt(K.S, torch.as_tensor(msgs), 59, K.op.sigma_glwe_abs, enc_seed=5, ct_base=70,
                       stride=K.p.k * K.p.N + 2)[:, : K.p.k * K.p.N + 1].contiguous()
    ksk32 = E.ksk_to_32(K.p, K.ksk)
    oksk32 = O.ksk_to_32(K.op, _u64(K.ksk))
    assert np.array_equal(ksk32.cpu().numpy().view(np.uint32), oksk32)
    out = E.keyswitch32(K.p, ksk32, ct)
    ref = O.keyswitch32(K.op, oksk32, _u64(ct))
    assert np.array_equal(_u64(out), ref)
    assert np.array_equal(O.lwe_decrypt(K.os, ref, 59), msgs)
    d = (O.lwe_phase(K.os, ref) - O.lwe_phase(K.os, O.keyswitch(K.op, _u64(K.ksk), _u64(ct)))).view(np.int64)
    assert np.log2(np.abs(d).max() + 1) - 64 < -19


def test_keyswitch_generic_level_count(O, cuda_dev):
    """l_ks outside the compile-time specialisations (3,4,5) takes the runtime-level kernel."""
    import torch
    from fhe_icp_b200 import engine as E
    d = dict(TOY, l_ks=6, beta_ks=2)
    K = Keys(O, cuda_dev, d)
    msgs = np.arange(12) % 8
    ct = E.lwe_encrypt(K.S, torch.as_tensor(msgs), 60, K.op.sigma_glwe_abs, enc_seed=8, stride=K.p.N + 2)[:, : K.p.N + 1].contiguous()
    assert np.array_equal(_u64(E.keyswitch(K.p, K.ksk, ct)), O.keyswitch(K.op, _u64(K.ksk), _u64(ct)))
    k32 = E.ksk_to_32(K.p, K.ksk)
    assert np.array_equal(_u64(E.keyswitch32(K.p, k32, ct)), O.keyswitch32(K.op, O.ksk_to_32(K.op, _u64(K.ksk)), _u64(ct)))


def test_pbs_large_batch_wide_kernel(O, p4, cuda_dev):
    """Batch > 2 x SM count takes the 4-ciphertexts-per-CTA kernel (incl. dead slots in the last CTA)."""
    import torch
    from fhe_icp_b200 import engine as E
    K = p4
    B = 2 * 148 + 3
    msgs = np.random.RandomState(3).randint(0, 16, size=B)
    ct = E.lwe_encrypt(K.s, torch.as_tensor(msgs), 59, K.op.sigma_lwe_abs, enc_seed=12, stride=K.p.n + 2)[:, : K.p.n + 1].contiguous()
    table = (np.arange(16) * 11 + 5) % 16
    out = E.pbs(K.p, K.bskf, ct, E.from_u64_numpy(E.make_lut_poly(table, 4, K.p.N, 59), cuda_dev))
    assert np.array_equal(O.lwe_decrypt(K.oS, _u64(out), 59) & 15, table[msgs])


class KeysMB2:
    def __init__(self, O, dev, d, key_seed=11, evk_seed=22):
        from fhe_icp_b200 import engine as E
        self.p = E.make_pbs_params(**d)
        self.op = _oparams(O, d)
        self.s = E.secret_key(key_seed, 0, d["n"], dev)
        self.S = E.secret_key(key_seed, 1, d["N_poly"], dev)
        self.os, self.oS = O.secret_key(key_seed, 0, d["n"]), O.secret_key(key_seed, 1, d["N_poly"])
        self.bsk2 = E.bsk2_gen(self.p, self.s, self.S, evk_seed)
        self.bskf2 = E.bsk2_to_fourier(self.p, self.bsk2)
        self.evk_seed = evk_seed


P4_L2 = dict(P4, l_pbs=2, beta_pbs=15)


@pytest.mark.parametrize("which,B", [("toy", 16), ("toy", 3), ("p4", 16), ("p4", 2 * 148 + 5), ("p4", 4 * 148 + 21),
                                     ("p4", 7 * 148 + 1), ("toy_l2", 16), ("toy_l2", 5), ("p4_l2", 16), ("p4_l2", 2 * 148 + 3)])
def test_multibit_pbs(O, cuda_dev, which, B):
    """Multi-bit blind rotation (two key bits per CMux): key bit-exact vs the oracle, every message
    maps to LUT[m], phases agree with the oracle's multi-bit evaluation within the PBS noise bound.  The batch sizes
    cover the dispatcher (csrc/pbs.cu::launch_pbs_mb2): latency kernel only, a full wave of four-ciphertext CTAs plus a
    remainder on the latency kernel, and a remainder too large for it (everything on the throughput kernel)."""
    import torch
    from fhe_icp_b200 import engine as E
    d = {"toy": TOY, "p4": P4, "toy_l2": TOY_L2, "p4_l2": P4_L2}[which]
    two = d["l_pbs"] == 2
    K = KeysMB2(O, cuda_dev, d)
    obsk2 = O.bsk2_gen(K.op, K.os, K.oS, K.evk_seed)
    assert np.array_equal(_u64(K.bsk2), obsk2)
    of = O.bsk2_to_fourier(K.op, obsk2)                     # [i][g][t][l][c][M][2]
    gf = K.bskf2.cpu().numpy()                              # [i][k1][g][t][l][c][32][2]
    L = d["l_pbs"]
    ref_sliced = of.reshape(of.shape[0], 3, 2, L, 2, 32, 32, 2).transpose(0, 5, 1, 2, 3, 4, 6, 7)
    assert np.abs(gf - ref_sliced).max() / np.abs(of).max() < 1e-13
    rng = np.random.RandomState(B)
    msgs = rng.randint(0, 16, size=B)
    msgs[: min(B, 16)] = np.arange(16)[: min(B, 16)]
    ct = E.lwe_encrypt(K.s, torch.as_tensor(msgs), 59, K.op.sigma_lwe_abs, enc_seed=B, ct_base=100,
                       stride=K.p.n + 2 - (K.p.n % 2))[:, : K.p.n + 1].contiguous()
    table = (np.arange(16) * 5 + 2) % 16
    lut = E.make_lut_poly(table, 4, K.p.N, 59)
    out = E.pbs_mb2(K.p, K.bskf2, ct, E.from_u64_numpy(lut, cuda_dev))
    got = _u64(out)
    assert np.array_equal(O.lwe_decrypt(K.oS, got, 59) & 15, table[msgs])
    ref = O.pbs_mb2(K.op, of, _u64(ct)[: min(B, 32)], lut)
    diff = (O.lwe_phase(K.oS, got[: min(B, 32)]) - O.lwe_phase(K.oS, ref)).view(np.int64).astype(np.float64)
    assert np.log2(np.abs(diff).max() + 1) - 64 < (-19 if two else -12)
    err = (O.lwe_phase(K.oS, got) - (table[msgs].astype(np.uint64) << np.uint64(59))).view(np.int64).astype(np.float64)
    assert np.log2(err.std() + 1) - 64 < (-20.5 if two else -13.5)


@pytest.mark.parametrize("which,B", [("toy", 5), ("toy", 130), ("p4", 1), ("p4", 40), ("p4", 300)])
def test_keyswitch_tensor_core_bit_exact(O, request, cuda_dev, which, B):
    """The keyswitch as an int8 contraction on the tensor cores (tcgen05.mma.kind::i8: signed digits x unsigned
    key bytes -> s32 in tensor memory, byte planes recombined in the epilogue) equals the integer-pipe KS32
    kernel -- and therefore the oracle -- word for word, including ragged row tiles and the padded last
    column tile; random torus inputs also cover every digit value."""
    import torch
    from fhe_icp_b200 import engine as E
    K = request.getfixturevalue(which)
    kN = K.p.k * K.p.N
    msgs = np.random.RandomState(B).randint(0, 16, size=B)
    ct = E.lwe_encrypt(K.S, torch.as_tensor(msgs), 59, K.op.sigma_glwe_abs, enc_seed=6, ct_base=9,
                       stride=kN + 2)[:, : kN + 1].contiguous()
    ksk32 = E.ksk_to_32(K.p, K.ksk)
    key_mma = E.ksk_to_mma(K.p, ksk32)
    out = E.keyswitch_mma(K.p, key_mma, ct)
    want = E.keyswitch32(K.p, ksk32, ct)
    assert np.array_equal(_u64(out), _u64(want))
    if B <= 40:
        assert np.array_equal(_u64(out), O.keyswitch32(K.op, O.ksk_to_32(K.op, _u64(K.ksk)), _u64(ct)))
    assert np.array_equal(O.lwe_decrypt(K.os, _u64(out), 59), msgs)
    rnd = torch.as_tensor(np.random.RandomState(B + 1).randint(-2 ** 63, 2 ** 63 - 1, size=(B, kN + 1), dtype=np.int64)).to(cuda_dev)
    assert np.array_equal(_u64(E.keyswitch_mma(K.p, key_mma, rnd)), _u64(E.keyswitch32(K.p, ksk32, rnd)))
    assert E.keyswitch_mma(K.p, key_mma, rnd[:0]).shape == (0, K.p.n + 1)


def test_keyswitch_tensor_core_other_gadgets(O, cuda_dev):
    """Gadgets other than the stated 5 x 3 bits: 3 x 4 and 2 x 8 bits (digits down to -128 fill the whole s8 range)."""
    import torch
    from fhe_icp_b200 import engine as E
    for l_ks, beta_ks in [(3, 4), (2, 8)]:
        K = Keys(O, cuda_dev, dict(TOY, l_ks=l_ks, beta_ks=beta_ks))
        kN = K.p.k * K.p.N
        rnd = torch.as_tensor(np.random.RandomState(l_ks).randint(-2 ** 63, 2 ** 63 - 1, size=(37, kN + 1), dtype=np.int64)).to(cuda_dev)
        ksk32 = E.ksk_to_32(K.p, K.ksk)
        out = E.keyswitch_mma(K.p, E.ksk_to_mma(K.p, ksk32), rnd)
        assert np.array_equal(_u64(out), _u64(E.keyswitch32(K.p, ksk32, rnd)))
        assert np.array_equal(_u64(out), O.keyswitch32(K.op, O.ksk_to_32(K.op, _u64(K.ksk)), _u64(rnd)))


@pytest.mark.parametrize("kernel,which,B", [("wide", "toy", 16), ("wide", "toy", 3), ("wide", "p4", 16), ("wide", "p4", 1), ("wide", "p4", 148 + 9),
                                            ("pair", "toy", 16), ("pair", "toy", 3), ("pair", "p4", 16), ("pair", "p4", 1), ("pair", "p4", 74 + 9)])
def test_multibit_pbs_latency_kernels(O, cuda_dev, kernel, which, B):
    """Same acceptance as test_multibit_pbs for the two latency kernels of pbs_wide.cu -- "wide": one ciphertext per CTA,
    four warps per polynomial; "pair": one ciphertext per cluster of two CTAs, spectra exchanged through distributed
    shared memory: every message maps to LUT[m], phases agree with the oracle's multi-bit PBS within the noise bound,
    per-ciphertext LUT selection works, and more CTAs / clusters than fit at once (a second wave) changes nothing."""
    import torch
    from fhe_icp_b200 import engine as E
    d = {"toy": TOY, "p4": P4}[which]
    K = KeysMB2(O, cuda_dev, d)
    rng = np.random.RandomState(1000 + B)
    msgs = rng.randint(0, 16, size=B)
    msgs[: min(B, 16)] = np.arange(16)[: min(B, 16)]
    ct = E.lwe_encrypt(K.s, torch.as_tensor(msgs), 59, K.op.sigma_lwe_abs, enc_seed=B, ct_base=100,
                       stride=K.p.n + 2 - (K.p.n % 2))[:, : K.p.n + 1].contiguous()
    tables = np.stack([(np.arange(16) * 5 + 2) % 16, (np.arange(16) * 3 + 7) % 16])
    luts = np.stack([E.make_lut_poly(tb, 4, K.p.N, 59) for tb in tables])
    lut_d = E.from_u64_numpy(luts, cuda_dev)
    which_lut = rng.randint(0, 2, size=B).astype(np.int32)
    run = E.pbs_mb2_wide if kernel == "wide" else E.pbs_mb2_pair
    got = _u64(run(K.p, K.bskf2, ct, lut_d, lut_index=torch.as_tensor(which_lut)))
    want = tables[which_lut, msgs]
    assert np.array_equal(O.lwe_decrypt(K.oS, got, 59) & 15, want)
    of = O.bsk2_to_fourier(K.op, O.bsk2_gen(K.op, K.os, K.oS, K.evk_seed))
    m = min(B, 8)
    ref = np.concatenate([O.pbs_mb2(K.op, of, _u64(ct)[i:i + 1], luts[which_lut[i]]) for i in range(m)])
    diff = (O.lwe_phase(K.oS, got[:m]) - O.lwe_phase(K.oS, ref)).view(np.int64).astype(np.float64)
    assert np.log2(np.abs(diff).max() + 1) - 64 < -12
    err = (O.lwe_phase(K.oS, got) - (want.astype(np.uint64) << np.uint64(59))).view(np.int64).astype(np.float64)
    assert np.log2(np.abs(err).max() + 1) - 64 < -11
    assert run(K.p, K.bskf2, ct[:0], lut_d).shape == (0, K.p.N + 1)


def test_full_size_batch_keyswitch_pbs_composition(cuda_dev):
    """BASELINE.json configs[2] at its upper end: 65 536 ciphertexts through keyswitch + PBS twice (the atomic pattern,
    tensor-core keyswitch, the dispatcher's mix of kernels: 110 full waves of the throughput kernel and a remainder on
    the latency kernel).  Size-independent properties: every ciphertext decrypts to table1[m] after the first bootstrap and
    to table2[table1[m]] after the second (a table lookup composes); re-running part of the batch reproduces the same words
    where the dispatcher picks the same kernel (deterministic whatever the CTA grouping) and the same values where it
    picks another (a different FFT schedule rounds differently, inside the noise)."""
    import ctypes as C
    import torch
    from fhe_icp_b200 import _native as N_
    from fhe_icp_b200 import engine as E
    dev = cuda_dev
    p = E.make_pbs_params(**P4)
    s, S = E.secret_key(31, 0, p.n, dev), E.secret_key(31, 1, p.k * p.N, dev)
    ksk = E.ksk_gen(p, S, s, 32)
    key_mma = E.ksk_to_mma(p, E.ksk_to_32(p, ksk))
    bskf2 = E.bsk2_to_fourier(p, E.bsk2_gen(p, s, S, 32))
    B = 65536
    rng = np.random.RandomState(7)
    msgs = rng.randint(0, 16, size=B)
    t1, t2 = (np.arange(16) * 7 + 3) % 16, (np.arange(16) * 5 + 1) % 16
    lut1 = E.from_u64_numpy(E.make_lut_poly(t1, 4, p.N, 59), dev)
    lut2 = E.from_u64_numpy(E.make_lut_poly(t2, 4, p.N, 59), dev)
    ct_big = E.lwe_encrypt(S, torch.as_tensor(msgs), 59, p.sigma_glwe_abs, enc_seed=33, stride=p.N + 2)[:, : p.N + 1].contiguous()
    work = torch.empty(int(N_.lib().fhe_b200_keyswitch_mma_workspace_bytes(C.byref(p), B)), dtype=torch.int8, device=dev)

    def dec(x):
        z = torch.zeros((B, p.N + 2), dtype=torch.int64, device=dev)
        z[:, : p.N + 1] = x
        return E.lwe_decrypt(S, z, 59).cpu().numpy() & 15

    small = E.keyswitch_mma(p, key_mma, ct_big, work=work)
    out1 = E.pbs_mb2(p, bskf2, small, lut1)
    assert np.array_equal(dec(out1), t1[msgs])
    nsub = 4 * 148 * 3 + 17                                                    # a different partition of the same inputs:
    again = E.pbs_mb2(p, bskf2, small[:nsub], lut1)                            # 3 full waves + 17 on the cluster kernel
    assert torch.equal(again[: nsub - 17], out1[: nsub - 17])                  # same kernel: the same words
    z = torch.zeros((nsub, p.N + 2), dtype=torch.int64, device=dev)
    z[:, : p.N + 1] = again
    assert np.array_equal(E.lwe_decrypt(S, z, 59).cpu().numpy() & 15, t1[msgs[:nsub]])   # another kernel: the same values
    out2 = E.pbs_mb2(p, bskf2, E.keyswitch_mma(p, key_mma, out1, work=work), lut2)
    assert np.array_equal(dec(out2), t2[t1[msgs]])


def test_sharded_bootstrap_on_one_rank_equals_the_direct_calls(cuda_dev):
    """sharded_search.ShardedBootstrap / BootstrapEngine (what bench.py's pbs_sharded record drives at N GPUs; the two-rank
    plumbing is tests/test_sharded_search.py::test_sharded_bootstrap_world2): on one rank it returns exactly the words of
    fhe_b200_keyswitch_mma + fhe_b200_pbs_mb2, per-ciphertext tables included, and a key-less engine that adopts the
    evaluation keys computes the same."""
    import torch
    from fhe_icp_b200 import engine as E
    from fhe_icp_b200.sharded_search import BootstrapEngine, ShardedBootstrap
    dev = cuda_dev
    p = E.make_pbs_params(**P4)
    s, S = E.secret_key(41, 0, p.n, dev), E.secret_key(41, 1, p.k * p.N, dev)
    key_mma = E.ksk_to_mma(p, E.ksk_to_32(p, E.ksk_gen(p, S, s, 42)))
    bskf2 = E.bsk2_to_fourier(p, E.bsk2_gen(p, s, S, 42))
    B = 37
    rng = np.random.RandomState(3)
    msgs, which = rng.randint(0, 16, size=B), rng.randint(0, 2, size=B)
    tables = np.stack([(np.arange(16) * 7 + 3) % 16, (np.arange(16) * 5 + 1) % 16])
    luts = torch.stack([E.from_u64_numpy(E.make_lut_poly(t, 4, p.N, 59), dev) for t in tables])
    idx = torch.as_tensor(which, dtype=torch.int32, device=dev)
    ct = E.lwe_encrypt(S, torch.as_tensor(msgs), 59, p.sigma_glwe_abs, enc_seed=43, stride=p.N + 2)[:, : p.N + 1].contiguous()
    eng = BootstrapEngine(P4, dev, key_mma, bskf2)
    sb = ShardedBootstrap(eng, client_rank=0, device=dev)
    assert sb.key_bytes == key_mma.numel() * key_mma.element_size() + bskf2.numel() * bskf2.element_size()
    out = sb.evaluate(ct, luts, idx)
    assert torch.equal(out, E.pbs_mb2(p, bskf2, E.keyswitch_mma(p, key_mma, ct), luts, idx))
    z = torch.zeros((B, p.N + 2), dtype=torch.int64, device=dev)
    z[:, : p.N + 1] = out
    assert np.array_equal(E.lwe_decrypt(S, z, 59).cpu().numpy() & 15, tables[which, msgs])
    server = BootstrapEngine(P4, dev)
    server.adopt_keys(eng.key_tensors())
    assert torch.equal(server.bootstrap(ct, luts, idx), out)
    assert server.bootstrap(ct[:0], luts).shape == (0, p.k * p.N + 1)
