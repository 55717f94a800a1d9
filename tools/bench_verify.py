"""Batched device verification timing; KEEP_PK=1 keeps the proving key (and its streams/tables) alive meanwhile."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, '.')
from libzkp_b200 import engine, transforms
engine.init(0)
pk_bytes, vk_bytes = engine.setup_builtin(engine.EQUALITY, 110, transforms._toxic(1))
pk = engine.ProvingKey(pk_bytes, window_bits=12)
pk.circuit_builtin(engine.EQUALITY, 110)
n = 4096
rng = np.random.default_rng(3)
a = rng.integers(0, 2**63, size=n, dtype=np.uint64)
r = np.zeros((n, 32), np.uint8); r[:, 0] = 7
s = np.zeros((n, 32), np.uint8); s[:, 0] = 9
proofs, cms, st = pk.prove_equality_batch(a, a, r, s)
if not os.environ.get("KEEP_PK"):
    pk.close()
vk = engine.VerifyingKey(vk_bytes)
for nb in [int(v) for v in os.environ.get('BATCHES', '1,8,64,256,512,1024,2048,4096').split(',')]:
    ok = vk.verify_batch(proofs[:nb], cms[:nb])
    assert ok.all()
    bad = proofs[:nb].copy(); bad[nb // 2, 200] ^= 1
    assert not vk.verify_batch(bad, cms[:nb])[nb // 2]
    K = 3
    ts = []
    for _ in range(K):
        t0 = time.perf_counter()
        vk.verify_batch(proofs[:nb], cms[:nb])
        ts.append(round(1e3 * (time.perf_counter() - t0), 2))
    dt = sum(ts) / K / 1e3
    print(json.dumps({"batch": nb, "ms": round(1e3 * dt, 2), "each": ts, "verifies_per_s": round(nb / dt, 1)}))
