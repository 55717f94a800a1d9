"""One MSM (resident and generic mode) and one NTT at the BASELINE sizes, for ncu launch lists."""
import sys
import torch
sys.path.insert(0, '.')
from libzkp_b200 import engine, transforms
engine.init(0)
dev = torch.device('cuda', 0)
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "msm"):
    print(transforms.bench_msm(torch, dev, 17.251e12, 20, 1, iters=1, resident=True))
    print(transforms.bench_msm(torch, dev, 17.251e12, 20, 1, iters=1, resident=False))
if which in ("all", "msm2"):
    print(transforms.bench_msm(torch, dev, 17.251e12, 18, 2, iters=1, resident=True))
if which in ("all", "ntt"):
    print(transforms.bench_ntt(torch, dev, 17.251e12, 6539.9, 22, iters=1))
