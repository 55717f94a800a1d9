import json, sys, torch
sys.path.insert(0, '.')
from libzkp_b200 import engine, transforms
engine.init(0)
dev = torch.device('cuda', 0)
print(json.dumps(transforms.bench(torch, dev, 17.251e12, 6539.9), indent=1))
print(json.dumps(transforms.bench_msm(torch, dev, 17.251e12, 20, 1, resident=False)))
