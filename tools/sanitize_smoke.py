"""Small end-to-end pass over every kernel family, meant to run under compute-sanitizer (memcheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libzkp_b200 import engine  # noqa: E402

engine.init(0)
rng = np.random.default_rng(1)


def fr(n):
    v = rng.integers(0, 2**32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    v[:, 7] %= 0x30644E72
    return v.view(np.uint8).reshape(n, 32)


toxic = [3, 5, 7, 11, 13]
# batched path, both circuits, two chunks
pk_bytes, _ = engine.setup_builtin(engine.EQUALITY, 110, toxic)
pk = engine.ProvingKey(pk_bytes, validate=True, window_bits=8, max_chunk=24)
pk.circuit_builtin(engine.EQUALITY, 110)
a = rng.integers(0, 2**63, size=40, dtype=np.uint64)
proofs, cms, status = pk.prove_equality_batch(a, a, fr(40), fr(40))
assert not status.any() and proofs.any()
z = np.stack([engine.builtin_witness(engine.EQUALITY, 110, int(x), int(x)) for x in a[:3]])
assert np.array_equal(pk.prove_batch(z, fr(3), fr(3))[1], np.zeros(3, np.int32))
pk.close()
pk_bytes, _ = engine.setup_builtin(engine.MEMBERSHIP, 64, toxic)
pk = engine.ProvingKey(pk_bytes, window_bits=8)
pk.circuit_builtin(engine.MEMBERSHIP, 64)
sets = rng.integers(0, 2**63, size=(5, 64), dtype=np.uint64)
proofs, _, status = pk.prove_membership_batch(sets[:, 3].copy(), sets, np.full(5, 64, np.uint32), fr(5), fr(5))
assert not status.any()
pk.close()
# large-domain path (two-pass NTT, resident Pippenger MSMs), sharded partial / combine
rounds = 2730
pk_bytes, _ = engine.setup_builtin(engine.EQUALITY, rounds, toxic)
pk = engine.ProvingKey(pk_bytes)
pk.circuit_builtin(engine.EQUALITY, rounds)
z = engine.builtin_witness(engine.EQUALITY, rounds, 9, 9)[None]
proofs, status = pk.prove_batch(z, fr(1), fr(1))
assert not status.any() and proofs.any()
pk.close()
# stand-alone transforms
x = fr(1 << 13)
assert np.array_equal(engine.ntt(engine.ntt(x, coset=True), inverse=True, coset=True), x)
assert np.array_equal(engine.ntt(engine.ntt(fr(1 << 7)), inverse=True).shape, (128, 32))
bases = engine.generator_mul(1, fr(3000))
for resident in (True, False):
    B = engine.MsmBases(1, bases, resident_windows=resident, validate=True)
    out = B.msm(fr(3000))
    assert any(out)
    B.close()
b2 = engine.generator_mul(2, fr(300))
B = engine.MsmBases(2, b2)
assert any(B.msm(fr(300)))
B.close()
print("sanitize smoke ok")
