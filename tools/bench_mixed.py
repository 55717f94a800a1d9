import json, sys, torch
sys.path.insert(0, '.')
from libzkp_b200 import engine, transforms
engine.init(0)
print(json.dumps(transforms.bench_mixed(torch, torch.device('cuda', 0))))
