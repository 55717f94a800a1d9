"""One 2^20-constraint proof on one GPU (BASELINE.json configs[3]): ms per proof and the per-stage timers."""
import json, sys
sys.path.insert(0, '.')
import torch
from libzkp_b200 import engine, transforms
engine.init(0)
print(json.dumps(transforms.bench_large_proof(torch, torch.device('cuda', 0))))
