"""Ciphertext wire format and key persistence -- no GPU."""
import os
import stat

import numpy as np
import pytest


def test_ciphertext_file_roundtrip_and_mmap(tmp_path, O):
    from fhe_icp_b200.serialization import load_ciphertexts, save_ciphertexts
    n, shift = 63, 40
    s = O.secret_key(3, 2, n)
    msgs = np.arange(-6, 6).reshape(3, 4)
    ct = O.lwe_encrypt(s, msgs, shift, 2.0 ** 20, 9).reshape(3, 4, n + 1)
    p = str(tmp_path / "docs.ct")
    save_ciphertexts(p, ct, n, shift, log2_sigma=-44.0, meta={"doc_ids": ["a", "b", "c"]})
    back, h = load_ciphertexts(p)
    assert back.dtype == np.uint64 and np.array_equal(back, ct)
    assert (h["n"], h["stride"], h["shift"], h["meta"]["doc_ids"]) == (n, n + 1, shift, ["a", "b", "c"])
    mm, _ = load_ciphertexts(p, mmap=True)
    assert np.array_equal(np.asarray(mm), ct) and mm.offset % 16 == 0
    assert np.array_equal(O.lwe_decrypt(s, back, shift), msgs)
    # empty and corrupt inputs
    save_ciphertexts(p, np.zeros((0, n + 1), dtype=np.uint64), n, shift)
    assert load_ciphertexts(p)[0].shape == (0, n + 1)
    with pytest.raises(ValueError):
        save_ciphertexts(p, np.zeros((1, n), dtype=np.uint64), n, shift)
    open(p, "wb").write(b"garbage-garbage-garbage")
    with pytest.raises(ValueError):
        load_ciphertexts(p)


def test_key_file_roundtrip_password_and_permissions(tmp_path):
    from fhe_icp_b200 import FHESimilarityModel
    from fhe_icp_b200.serialization import keyset_from_model, load_keys, save_keys
    m = FHESimilarityModel(input_dim=16, n_bits=8, seed=1, key_seed=77, enc_seed=78, verbose=False)
    X, _ = m.train(n_samples=60)
    m.compile(X[:10])
    ks = keyset_from_model(m)
    p = str(tmp_path / "keys.bin")
    save_keys(p, ks, "correct horse")
    assert stat.S_IMODE(os.stat(p).st_mode) == 0o600          # test_suite.py:48-50
    back = load_keys(p, "correct horse")
    assert back == ks and back.key_seed == 77 and back.lwe["n"] == m.model.fhe_circuit.lwe.n
    with pytest.raises(ValueError, match="Invalid master password"):   # test_suite.py:315-316
        load_keys(p, "wrong")
    # the restored public spec reproduces the clear circuit
    from fhe_icp_b200.quantization import QuantizedLinearSpec
    spec = QuantizedLinearSpec.from_dict(back.quantized_spec)
    assert np.array_equal(spec.predict_clear(X[:5]), m.predict_clear(X[:5]))
