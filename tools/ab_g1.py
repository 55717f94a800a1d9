"""G1 large-MSM timing (2^20 points, uniform and witness-like scalars) for the library LZKP_B200_LIB names."""
import json, os, sys, torch
sys.path.insert(0, '.')
from libzkp_b200 import engine, transforms
engine.init(0)
dev = torch.device('cuda', 0)
r = transforms.bench_msm(torch, dev, 17.251e12, 20, 1)
w = transforms.bench_msm(torch, dev, 17.251e12, 20, 1, witness_like=True)
print(json.dumps({"lib": os.environ.get("LZKP_B200_LIB", "default"), "g1_2^20_ms": r["ms"], "witness_like_ms": w["ms"], "result": r["result_hex"]}))
