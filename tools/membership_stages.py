import json, sys
import numpy as np
sys.path.insert(0, '.')
import torch
from libzkp_b200 import engine, transforms
engine.init(0)
pk_bytes, _ = engine.setup_builtin(engine.MEMBERSHIP, 64, transforms._toxic(1))
pk = engine.ProvingKey(pk_bytes)
pk.circuit_builtin(engine.MEMBERSHIP, 64)
n = 4096
rng = np.random.default_rng(5)
sets = rng.integers(0, 2**63, size=(n, 64), dtype=np.uint64)
lens = np.full(n, 64, np.uint32)
vals = sets[np.arange(n), np.arange(n) % 64].copy()
dev = torch.device('cuda', 0)
r = transforms._uniform_fr(torch, dev, n, 7).cpu().numpy().view(np.uint8).reshape(n, 32)
s = transforms._uniform_fr(torch, dev, n, 8).cpu().numpy().view(np.uint8).reshape(n, 32)
for _ in range(2): pk.prove_membership_batch(vals, sets, lens, r, s)
engine.profile_enable(True); pk.profile_read(reset=True)
import time
t0 = time.perf_counter()
for _ in range(5): pk.prove_membership_batch(vals, sets, lens, r, s)
dt = (time.perf_counter() - t0) / 5
reg = pk.profile_read(reset=True)
print(json.dumps({"window_bits": pk.window_bits, "work": pk.work(), "ms_per_batch": 1e3 * dt, "proofs_per_s": n / dt, "stage_ms": {k: round(v[0] / max(v[1], 1), 3) for k, v in reg.items() if v[1]}}))
