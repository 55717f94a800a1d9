"""A few resident G1 MSMs of 2^LOG_N points (for an ncu launch list of a shard-sized call)."""
import os, sys
sys.path.insert(0, '.')
import numpy as np, torch
from libzkp_b200 import engine, transforms
engine.init(0)
dev = torch.device('cuda', 0)
ln = int(os.environ.get("LOG_N", "17")); n = 1 << ln
ks = transforms._uniform_fr(torch, dev, n, 9).cpu().numpy().view(np.uint8).reshape(n, 32)
B = engine.MsmBases(1, engine.generator_mul(1, ks), window_bits=16, resident_windows=True)
sc = transforms._uniform_fr(torch, dev, n, 10)
out = torch.zeros(64, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    B.msm_device(sc.data_ptr(), n, out.data_ptr(), st)
torch.cuda.synchronize()
