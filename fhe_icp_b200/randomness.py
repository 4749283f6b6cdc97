"""Client-side randomness policy: which seeds are secret, which are public, and how ciphertext ids are allotted.

The reference leaves all of this to Concrete, which draws keys and encryption randomness from a CSPRNG.  Here every
random quantity is a counter-based stream ``f(seed, kind, purpose, object id, block)``, so the
policy is about seeds and ids:

* ``key_seed``   -- SECRET.  Secret keys.
* ``noise_seed`` -- SECRET.  Error terms of every encryption.  Never equal to / derivable from ``enc_seed``: an evaluator
  who can regenerate the errors of seeded ciphertexts learns ``<a,s> + Delta*m`` exactly and solves for the key.
* ``evk_seed``   -- SECRET.  Masks and errors of the evaluation keys; only the expanded keys ever leave the client.
* ``enc_seed``   -- PUBLIC.  Masks of fresh ciphertexts; travels with seeded (compressed) ciphertexts.

All four default to the OS CSPRNG (``secrets``).  Fixed values are an explicit opt-in for reproducible tests and
benchmarks.  Ciphertext ids (the ``object id`` of a mask / error stream) must never repeat under one seed pair -- two
ciphertexts with equal mask and error differ exactly by ``Delta*(m1 - m2)`` -- so they come from one monotonic counter per
key set and process, started at a random 62-bit origin in the default (non-deterministic) mode: a persisted key set that
another process loads draws a fresh origin there (two processes overlap with probability ~ count / 2^62), so nothing has
to be written back to the key file after an encryption.
"""
from __future__ import annotations

import secrets
from typing import Optional


def fresh_seed() -> int:
    return secrets.randbits(64)


def seed_or_fresh(seed: Optional[int]) -> int:
    return fresh_seed() if seed is None else int(seed) & 0xFFFFFFFFFFFFFFFF


class CiphertextIds:
    """Monotonic allocator of ciphertext ids.  ``start=None``: random 62-bit origin (collision probability of two
    processes sharing a key set ~ count / 2^62); an integer origin is for deterministic tests."""

    def __init__(self, start: Optional[int] = None):
        self.next = secrets.randbits(62) if start is None else int(start)

    def take(self, count: int) -> int:
        base = self.next
        self.next = (self.next + max(int(count), 0)) & 0xFFFFFFFFFFFFFFFF
        return base
