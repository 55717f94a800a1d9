"""One random-linear-combination verification of 65 536 equality proofs (for an ncu launch list of its stages)."""
import sys
sys.path.insert(0, '.')
import torch
from libzkp_b200 import engine, transforms
engine.init(0)
print(transforms.bench_verify(torch, torch.device('cuda', 0), n=65536, iters=1))
