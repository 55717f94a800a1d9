"""Per-rank device time of the sharded 2^20-constraint proof, measured on ONE GPU: every shard of an N-way split is
loaded side by side (together they are one key) and its share of the proof - z-only MSMs on the side streams, the
witness map on map ranks, the H range - is timed alone.  The slowest shard bounds the N-GPU proof; the spread shows
how well lzkp_pk_load_ex's cost model balances.  Usage: python tools/shard_balance.py [N] [rounds]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libzkp_b200 import engine, transforms  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 349524
engine.init(0)
dev = torch.device("cuda", 0)
pk_bytes, _ = engine.setup_builtin(engine.EQUALITY, rounds, transforms._toxic(1))
z = torch.from_numpy(engine.builtin_witness(engine.EQUALITY, rounds, 6, 6)).to(dev)
rs = transforms._uniform_fr(torch, dev, 2, 4)
out = []
st = torch.cuda.current_stream().cuda_stream
for i in range(N):
    pk = engine.ProvingKey(pk_bytes, shard_index=i, shard_count=N)
    first, count, k = pk.shard_info()
    if count[3]:
        pk.circuit_builtin(engine.EQUALITY, rounds)
    slot = torch.zeros(784, dtype=torch.uint8, device=dev)
    run = lambda: pk.prove_partial_device(z.data_ptr(), rs[0].data_ptr(), rs[1].data_ptr(), 0, slot.data_ptr(),
                                          slot.data_ptr() + 768, st, phase=3)
    ms = transforms._time(torch, run, 5, warmup=2)
    out.append({"rank": i, "ms": round(ms, 3), "points": dict(zip(("a", "b1", "l", "h", "b2"), count)), "map": bool(count[3])})
    pk.close()
print(json.dumps({"n": N, "map_ranks": k, "slowest_ms": max(o["ms"] for o in out), "mean_ms": round(sum(o["ms"] for o in out) / N, 3),
                  "map_cost": os.environ.get("LZKP_SHARD_MAP_COST", "0.11")}))
for o in out:
    print(json.dumps(o))
