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
