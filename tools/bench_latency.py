"""Latency of small batches (BASELINE.json configs[0]: a single prove_equality) through the host-buffer C ABI."""
import json, sys, time
import numpy as np
sys.path.insert(0, '.')
from libzkp_b200 import engine, transforms
engine.init(0)
pk_bytes, _ = engine.setup_builtin(engine.EQUALITY, 110, transforms._toxic(1))
pk = engine.ProvingKey(pk_bytes)
pk.circuit_builtin(engine.EQUALITY, 110)
rng = np.random.default_rng(3)
out = {}
for n in (1, 8, 64, 512):
    a = rng.integers(0, 2**63, size=n, dtype=np.uint64)
    r = np.zeros((n, 32), np.uint8); r[:, 0] = 7
    s = np.zeros((n, 32), np.uint8); s[:, 0] = 9
    for _ in range(3):
        pk.prove_equality_batch(a, a, r, s)
    t0 = time.perf_counter()
    K = 20
    for _ in range(K):
        proofs, _, st = pk.prove_equality_batch(a, a, r, s)
    out[n] = round(1e3 * (time.perf_counter() - t0) / K, 3)
print(json.dumps({"ms_per_call_by_batch": out}))
