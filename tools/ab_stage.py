"""Per-stage device time of the bench workload (4096 equality proofs) for the library LZKP_B200_LIB names.
Usage: LZKP_B200_LIB=path/to/variant.so python tools/ab_stage.py [batch] [window_bits]"""
import json, os, sys
import numpy as np
sys.path.insert(0, '.')
import torch
from libzkp_b200 import engine, transforms
P = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
wb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
engine.init(0)
pk_bytes, _ = engine.setup_builtin(engine.EQUALITY, 110, transforms._toxic(1))
pk = engine.ProvingKey(pk_bytes, window_bits=wb)
pk.circuit_builtin(engine.EQUALITY, 110)
dev = torch.device('cuda', 0)
rng = np.random.default_rng(3)
d_a = torch.from_numpy(rng.integers(0, 2**63, size=P, dtype=np.int64)).to(dev)
rs = rng.integers(0, 256, size=(2, P, 32), dtype=np.uint8); rs[:, :, 31] &= 0x0f
d_r, d_s = torch.from_numpy(rs[0]).to(dev), torch.from_numpy(rs[1]).to(dev)
d_proofs = torch.zeros((P, 256), dtype=torch.uint8, device=dev)
d_status = torch.zeros(P, dtype=torch.int32, device=dev)
st = torch.cuda.current_stream()
def step():
    pk.prove_equality_batch_device(P, d_a.data_ptr(), d_a.data_ptr(), d_r.data_ptr(), d_s.data_ptr(),
                                   d_proofs.data_ptr(), d_status.data_ptr(), st.cuda_stream)
for _ in range(3): step()
torch.cuda.synchronize()
assert int(d_status.abs().sum()) == 0
K = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(K): step()
e1.record(st); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
engine.profile_enable(True); pk.profile_read(reset=True)
for _ in range(5): step()
torch.cuda.synchronize()
reg = pk.profile_read(reset=True)
import hashlib
print(json.dumps({"lib": os.environ.get("LZKP_B200_LIB", "default"), "batch": P, "ms_per_step": round(ms, 3),
                  "proofs_per_s": round(P / ms * 1e3), "stage_ms": {k: round(v[0] / max(v[1], 1), 3) for k, v in reg.items() if v[1]},
                  "proofs_sha": hashlib.sha256(d_proofs.cpu().numpy().tobytes()).hexdigest()[:16]}))
