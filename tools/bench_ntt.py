import json, sys, torch
sys.path.insert(0, '.')
from libzkp_b200 import engine, transforms
engine.init(0)
dev = torch.device('cuda', 0)
for ln in (20, 22, 24):
    d = transforms.bench_ntt(torch, dev, 17.251e12, 6539.9, ln, iters=10)
    print(ln, round(d["ms"], 4), "ms", round(d["elements_per_s"] / 1e9, 3), "G/s", round(d["imad_frac_of_measured_peak"], 3))
