"""Latency of small batches (BASELINE.json configs[0]: a single prove_equality) through the host-buffer C ABI, with
UNIFORM prover randomness r, s (tiny r, s would skip most additions of the assembly and flatter the number) and a
per-stage device-time breakdown.  LZKP_LATENCY_BATCH=0 gives the round-1 form (variable-base s*A + r*B1) for comparison."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, '.')
from libzkp_b200 import engine, transforms
engine.init(0)
pk_bytes, _ = engine.setup_builtin(engine.EQUALITY, 110, transforms._toxic(1))
pk = engine.ProvingKey(pk_bytes)
pk.circuit_builtin(engine.EQUALITY, 110)
rng = np.random.default_rng(3)
out = {}
for n in [int(v) for v in os.environ.get('BATCHES', '1,8,64,512').split(',')]:
    a = rng.integers(0, 2**63, size=n, dtype=np.uint64)
    r = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); r[:, 31] &= 0x1f
    s = rng.integers(0, 256, size=(n, 32), dtype=np.uint8); s[:, 31] &= 0x1f
    for _ in range(3):
        pk.prove_equality_batch(a, a, r, s)
    t0 = time.perf_counter()
    K = 20
    for _ in range(K):
        proofs, _, st = pk.prove_equality_batch(a, a, r, s)
    out[n] = round(1e3 * (time.perf_counter() - t0) / K, 3)
    if n == 1 or len(out) == 1:
        engine.profile_enable(True); pk.profile_read(reset=True)
        for _ in range(K):
            pk.prove_equality_batch(a, a, r, s)
        stages = {k: round(v[0] / max(v[1], 1), 4) for k, v in pk.profile_read(reset=True).items() if v[1]}
        engine.profile_enable(False)
import os
print(json.dumps({"latency_batch_limit": os.environ.get("LZKP_LATENCY_BATCH", "default (384)"), "ms_per_call_by_batch": out,
                  "stage_ms_single_proof": stages}))
