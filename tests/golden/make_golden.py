#!/usr/bin/env python3
"""Generates tests/golden/golden_v1.json from the CPU oracle with fixed seeds.

The reference tree holds no golden vectors and Concrete cannot be imported here (SURVEY.md
section 8c), so these fixtures pin the ORACLE (regression + cross-implementation: the CUDA path
is checked against the same numbers on the GPU box).  The Philox entries are the published
Random123 known-answer vectors, not oracle output.

    python tests/golden/make_golden.py        # rewrites golden_v1.json
"""
import base64
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402


def h(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def build() -> dict:
    g = {"version": 1}
    g["philox_kat"] = [
        {"ctr": [0, 0, 0, 0], "key": [0, 0], "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
        {"ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2, "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
        {"ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
         "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]},
        # Random123 kat_vectors, "philox4x32 7": the round count of the public mask stream (FHE_B200_MASK_ROUNDS)
        {"rounds": 7, "ctr": [0, 0, 0, 0], "key": [0, 0], "out": [0x5f6fb709, 0x0d893f64, 0x4f121f81, 0x4f730a48]},
        {"rounds": 7, "ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2,
         "out": [0x5207ddc2, 0x45165e59, 0x4d8ee751, 0x8c52f662]},
        {"rounds": 7, "ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
         "out": [0x4dfccaba, 0x190a87f0, 0xc47362ba, 0xb6b5242a]},
    ]
    g["gaussian"] = {"seed": 123, "domain": 3, "sigma_abs": 1e6,
                     "first8": [O.gaussian(123, 3, i, 0, 1e6) for i in range(8)],
                     "sha256_4096": h(np.array([O.gaussian(123, 3, i, 0, 1e6) for i in range(4096)], dtype=np.int64))}
    g["secret_key"] = {"seed": 1234567, "cases": [{"key_id": kid, "dim": dim, "weight": int(O.secret_key(1234567, kid, dim).sum()),
                                                     "sha256": h(O.secret_key(1234567, kid, dim))}
                                                    for kid, dim in [(0, 742), (1, 2048), (2, 1423)]]}
    # LWE encrypt / linear combination / decrypt, n = 15 (tiny) and n = 1423 (the compiled circuit)
    enc = []
    for n in (15, 1423):
        s = O.secret_key(99, 2, n)
        rng = np.random.RandomState(n)
        msgs = rng.randint(-128, 128, size=(3, 8))
        stride = (n + 2) & ~1
        ct = O.lwe_encrypt(s, msgs, 42, 2.0 ** 28, 555, ct_base=1000, stride=stride).reshape(3, 8, stride)
        W = np.stack([rng.randint(-128, 128, size=8), np.ones(8, dtype=np.int64)])
        out = O.lincomb(ct, W, n)
        enc.append({"n": n, "stride": stride, "msgs": msgs.tolist(), "W": W.tolist(), "shift": 42, "log2_sigma_abs": 28,
                    "enc_seed": 555, "ct_base": 1000, "key_seed": 99, "ct_sha256": h(ct), "out_sha256": h(out),
                    "ct_first_words": [int(x) for x in ct[0, 0, :4]], "body0": int(ct[0, 0, n]),
                    "decrypted": O.lwe_decrypt(s, out, 42).tolist()})
    g["lwe"] = enc
    # quantized clear circuit on the reference's generator (np.random.seed(42) order of draws)
    from fhe_icp_b200 import FHESimilarityModel
    m = FHESimilarityModel(input_dim=128, n_bits=8, seed=42, verbose=False)
    X, y = m.train()
    sp = m.model.spec
    q = sp.input_q.quant(X[:16])
    g["clear_circuit"] = {
        "seed": 42, "solver": m.model.solver_, "x_sha256": h(X), "input_q": sp.input_q.to_dict(),
        "weight_q": sp.weight_q.to_dict(), "q_weights_sha256": h(sp.q_weights), "q_bias": int(sp.q_bias),
        "out_scale": sp.out_scale, "out_zero_point": int(sp.out_zero_point),
        "q_weights": sp.q_weights.tolist(), "x_first16_f32_b64": base64.b64encode(X[:16].astype(np.float32).tobytes()).decode(),
        "q_x_first16_sha256": h(q), "q_y_first16": sp.circuit(q).tolist(), "y_first4": sp.predict_clear(X[:4]).tolist(),
        "note": "q_weights come from a float32 LAPACK least-squares fit and are NOT portable across BLAS builds; "
                "tests rebuild the circuit from the stored spec"}
    # keyswitch + PBS on a toy set (n = 16, N = 2048)
    p = O.make_params(n=16, k=1, N=2048, l_pbs=1, beta_pbs=23, l_ks=5, beta_ks=3, log2_sigma_lwe=-30.0)
    s, S = O.secret_key(11, 0, 16), O.secret_key(11, 1, 2048)
    ksk, bsk = O.ksk_gen(p, S, s, 22), O.bsk_gen(p, s, S, 22)
    msgs = np.arange(16)
    ct = O.lwe_encrypt(S, msgs, 59, p.sigma_glwe_abs, 5, ct_base=7)
    ks = O.keyswitch(p, ksk, ct)
    table = (np.arange(16) * 7 + 3) % 16
    out = O.pbs(p, O.bsk_to_fourier(p, bsk), ks, O.make_lut_poly(table, 4, 2048, 59))
    g["ks_pbs_toy"] = {"ksk_sha256": h(ksk), "bsk_sha256": h(bsk), "ks_out_sha256": h(ks),
                       "ks_decrypted": O.lwe_decrypt(s, ks, 59).tolist(), "table": table.tolist(),
                       "pbs_decrypted": (O.lwe_decrypt(S, out, 59) & 15).tolist()}
    return g


if __name__ == "__main__":
    out = Path(__file__).resolve().parent / "golden_v1.json"
    out.write_text(json.dumps(build(), indent=1))
    print("wrote", out)
