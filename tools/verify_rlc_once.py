"""Random-linear-combination verification of N equality proofs (N = 65536 or the sizes in SIZES=a,b,c): timing, or an
ncu launch list of its stages."""
import os, sys
sys.path.insert(0, '.')
import torch
from libzkp_b200 import engine, transforms
engine.init(0)
for n in [int(v) for v in os.environ.get("SIZES", "65536").split(",")]:
    r = transforms.bench_verify(torch, torch.device('cuda', 0), n=n, iters=2)
    print(n, {k: (round(v["ms"], 2), round(v["verifies_per_s"])) for k, v in r.items() if isinstance(v, dict)})
