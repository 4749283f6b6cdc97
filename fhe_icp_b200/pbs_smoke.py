"""One small keyswitch + PBS on cuda:0, checked against the oracle's LUT evaluation
(called from __graft_entry__.smoke(), which is allowed to use the oracle as the checker)."""
import numpy as np


def run(O):
    import torch
    from . import engine as E
    dev = torch.device("cuda", 0)
    d = dict(n=32, k=1, N_poly=2048, l_pbs=1, beta_pbs=23, l_ks=5, beta_ks=3, log2_sigma_lwe=-30.0, log2_sigma_glwe=-51.6)
    p = E.make_pbs_params(**d)
    s, S = E.secret_key(1, 0, p.n, dev), E.secret_key(1, 1, p.k * p.N, dev)
    ksk, bsk = E.ksk_gen(p, S, s, 2), E.bsk_gen(p, s, S, 2)
    bskf = E.bsk_to_fourier(p, bsk)
    msgs = np.arange(16)
    ct = E.lwe_encrypt(S, torch.as_tensor(msgs), 59, p.sigma_glwe_abs, enc_seed=3, stride=p.N + 2)[:, : p.N + 1].contiguous()
    table = (np.arange(16) * 3 + 2) % 16
    lut = E.from_u64_numpy(E.make_lut_poly(table, 4, p.N, 59), dev)
    out = E.pbs(p, bskf, E.keyswitch(p, ksk, ct), lut)
    dec = O.lwe_decrypt(O.secret_key(1, 1, p.k * p.N), E.to_u64_numpy(out), 59) & 15
    assert np.array_equal(dec, table), "decrypt(PBS(KS(enc(m)))) != LUT[m]"
    # the same atomic pattern with the keyswitch on the tensor cores (tcgen05 int8 contraction): bit-identical
    # to the integer-pipe 32-bit keyswitch and to the oracle's, and the bootstrap behind it still evaluates the table
    ksk32 = E.ksk_to_32(p, ksk)
    ks_tc = E.keyswitch_mma(p, E.ksk_to_mma(p, ksk32), ct)
    assert torch.equal(ks_tc, E.keyswitch32(p, ksk32, ct)), "tensor-core keyswitch != integer keyswitch"
    op = O.make_params(n=d["n"], k=d["k"], N=d["N_poly"], l_pbs=d["l_pbs"], beta_pbs=d["beta_pbs"], l_ks=d["l_ks"],
                       beta_ks=d["beta_ks"], log2_sigma_lwe=d["log2_sigma_lwe"], log2_sigma_glwe=d["log2_sigma_glwe"])
    assert np.array_equal(E.to_u64_numpy(ks_tc), O.keyswitch32(op, O.ksk_to_32(op, E.to_u64_numpy(ksk)), E.to_u64_numpy(ct))), \
        "tensor-core keyswitch != oracle"
    dec = O.lwe_decrypt(O.secret_key(1, 1, p.k * p.N), E.to_u64_numpy(E.pbs(p, bskf, ks_tc, lut)), 59) & 15
    assert np.array_equal(dec, table), "decrypt(PBS(KS_mma(enc(m)))) != LUT[m]"
