import json, sys
sys.path.insert(0, '.')
import torch
from libzkp_b200 import engine, transforms
engine.init(0)
dev = torch.device('cuda', 0)
for ln in (17, 18, 20):
    r = transforms.bench_msm(torch, dev, 18.49e12, ln, 1)
    print(json.dumps({"log_n": ln, "ms": round(r["ms"], 4), "Mpts_s": round(r["points_per_s"] / 1e6, 1)}))
r = transforms.bench_msm(torch, dev, 18.49e12, 20, 1, resident=False)
print(json.dumps({"one_shot 2^20 ms": round(r["ms"], 4)}))
r = transforms.bench_msm(torch, dev, 18.49e12, 17, 2, iters=5)
print(json.dumps({"g2 2^17 ms": round(r["ms"], 4)}))
