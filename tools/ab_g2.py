"""G2 large-MSM timing (2^18 points) for the library LZKP_B200_LIB names."""
import json, os, sys, torch
sys.path.insert(0, '.')
from libzkp_b200 import engine, transforms
engine.init(0)
dev = torch.device('cuda', 0)
r = transforms.bench_msm(torch, dev, 17.251e12, 18, 2)
print(json.dumps({"lib": os.environ.get("LZKP_B200_LIB", "default"), "g2_2^18_ms": r["ms"], "result": r["result_hex"]}))
