"""Wire / on-disk formats for ciphertexts and key material.

The reference stores *plaintext* float32 vectors as "encrypted" documents (pickle + gzip,
/root/reference/encrypted_storage.py:40-47) and persists only a config dict as "keys"
(/root/reference/key_management.py:148-166).  Here:

* ciphertext file: magic ``FHEB200C`` | u32 version | u32 header length | JSON header | raw little-
  endian u64 words ``[count][stride]`` (zero-copy ``np.memmap``-able);
* key file: every key of the engine is derived from 64-bit seeds by counter-based streams
  (DESIGN.md section 3), so the persistent secret is the seed set + parameters; it is wrapped with
  the reference's own scheme -- PBKDF2-HMAC-SHA256 (100 000 iterations, 16-byte salt) -> Fernet
  (key_management.py:49-58) -- and a wrong password raises ``ValueError`` like the reference.

Both are used by the product: ``BatchProcessor(fhe="both")`` keeps the collection as packed GLWE ciphertexts in
``<storage>/collection.glwe`` (written by ``encrypt-batch``, memory-mapped by ``search`` / ``compare`` -- no plaintext
embedding is ever stored), and the CLI's ``--keys`` option / ``keys generate`` command persist the key set.
"""
from __future__ import annotations

import base64
import json
import os
import struct
from dataclasses import asdict, dataclass, field
from typing import Optional

import numpy as np

CT_MAGIC = b"FHEB200C"
KEY_MAGIC = b"FHEB200K"
VERSION = 1


def save_ciphertexts(path: str, ct, n: int, shift: int, log2_sigma: Optional[float] = None, meta: Optional[dict] = None):
    """``ct``: numpy uint64/int64 array or torch tensor ``[..., stride]``."""
    if hasattr(ct, "detach"):
        ct = ct.detach().cpu().numpy()
    a = np.ascontiguousarray(ct).view(np.uint64)
    header = {"n": int(n), "stride": int(a.shape[-1]), "shape": list(a.shape), "shift": int(shift),
              "log2_sigma": log2_sigma, "word": "u64le", "meta": meta or {}}
    if header["stride"] < n + 1:
        raise ValueError("stride must be >= n + 1")
    hb = json.dumps(header).encode()
    pad = (-(len(CT_MAGIC) + 8 + len(hb))) % 16      # payload starts 16-byte aligned
    with open(path, "wb") as f:
        f.write(CT_MAGIC + struct.pack("<II", VERSION, len(hb) + pad) + hb + b" " * pad)
        f.write(a.astype("<u8", copy=False).tobytes())
    return header


def load_ciphertexts(path: str, mmap: bool = False):
    """Returns (uint64 array of the saved shape, header dict)."""
    with open(path, "rb") as f:
        head = f.read(16)
        if head[:8] != CT_MAGIC:
            raise ValueError("not a fhe_b200 ciphertext file")
        version, hlen = struct.unpack("<II", head[8:16])
        if version != VERSION:
            raise ValueError(f"unsupported ciphertext file version {version}")
        header = json.loads(f.read(hlen).decode())
        off = 16 + hlen
    shape = tuple(header["shape"])
    if mmap:
        return np.memmap(path, dtype="<u8", mode="r", offset=off, shape=shape), header
    data = np.fromfile(path, dtype="<u8", offset=off)
    if data.size != int(np.prod(shape)):
        raise ValueError("truncated ciphertext file")
    return data.reshape(shape), header


@dataclass
class KeySet:
    """Everything needed to regenerate the client's secret keys and the evaluation keys.  ``key_seed``, ``noise_seed``
    and ``evk_seed`` are secrets; ``enc_seed`` is the public mask seed (randomness.py)."""
    key_seed: int
    enc_seed: int
    evk_seed: int = 0
    noise_seed: int = 0
    lwe: dict = field(default_factory=dict)          # leveled path: n, stride, shift, log2_sigma
    pbs: dict = field(default_factory=dict)          # keyswitch / PBS parameter set, if used
    quantized_spec: Optional[dict] = None            # the compiled model's public part

    def to_json(self) -> bytes:
        return json.dumps(asdict(self)).encode()

    @classmethod
    def from_json(cls, b: bytes) -> "KeySet":
        return cls(**json.loads(b.decode()))

    @classmethod
    def generate(cls, **kw) -> "KeySet":
        """A fresh key set from the OS CSPRNG."""
        from .randomness import fresh_seed
        return cls(key_seed=fresh_seed(), enc_seed=fresh_seed(), evk_seed=fresh_seed(), noise_seed=fresh_seed(), **kw)


def _derive(password: str, salt: bytes) -> bytes:
    from cryptography.hazmat.primitives import hashes
    from cryptography.hazmat.primitives.kdf.pbkdf2 import PBKDF2HMAC
    kdf = PBKDF2HMAC(algorithm=hashes.SHA256(), length=32, salt=salt, iterations=100000)
    return base64.urlsafe_b64encode(kdf.derive(password.encode()))


def save_keys(path: str, keys: KeySet, password: str):
    from cryptography.fernet import Fernet
    salt = os.urandom(16)
    token = Fernet(_derive(password, salt)).encrypt(keys.to_json())
    with open(path, "wb") as f:
        f.write(KEY_MAGIC + struct.pack("<I", VERSION) + salt + token)
    os.chmod(path, 0o600)                            # like key_management.py:108,166


def load_keys(path: str, password: str) -> KeySet:
    from cryptography.fernet import Fernet, InvalidToken
    raw = open(path, "rb").read()
    if raw[:8] != KEY_MAGIC:
        raise ValueError("not a fhe_b200 key file")
    (version,) = struct.unpack("<I", raw[8:12])
    if version != VERSION:
        raise ValueError(f"unsupported key file version {version}")
    salt, token = raw[12:28], raw[28:]
    try:
        return KeySet.from_json(Fernet(_derive(password, salt)).decrypt(token))
    except InvalidToken:
        raise ValueError("Invalid master password")


def keyset_from_model(model) -> KeySet:
    c = model.model.fhe_circuit
    return KeySet(key_seed=c.key_seed, enc_seed=c.enc_seed, noise_seed=c.noise_seed, lwe=c.lwe.to_dict(),
                  quantized_spec=c.spec.to_dict())
