import json, sys, torch
sys.path.insert(0, '.')
from libzkp_b200 import engine, transforms
engine.init(0)
dev = torch.device('cuda', 0)
print(json.dumps(transforms.bench_membership(torch, dev, iters=2, slots=1024)))
print(json.dumps(transforms.bench_membership(torch, dev, iters=5, slots=64)))
