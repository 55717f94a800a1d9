"""torchrun entry: one large proof split across N GPUs (NCCL), checked on rank 0 against the single-GPU
engine path and timed.  Usage: torchrun --nproc-per-node N tools/run_sharded_proof.py [rounds]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from libzkp_b200 import engine  # noqa: E402
from libzkp_b200.multi import ShardedProver  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 349524
torch.cuda.set_device(local)
engine.init(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
toxic = [3, 5, 7, 11, 13]
pk_bytes, vk_bytes = engine.setup_builtin(engine.EQUALITY, rounds, toxic)
sp = ShardedProver(pk_bytes, engine.EQUALITY, rounds, rank, world, dev)
z = engine.builtin_witness(engine.EQUALITY, rounds, 6, 6) if rank == 0 else None
r, s = (17).to_bytes(32, "little"), (19).to_bytes(32, "little")
proof = sp.prove(z, r, s)
for _ in range(2):
    sp.prove(resident=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
K = 5
for _ in range(K):
    sp.prove(resident=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
dt = (time.perf_counter() - t0) / K
if rank == 0:
    full = engine.ProvingKey(pk_bytes)
    full.circuit_builtin(engine.EQUALITY, rounds)
    want, status = full.prove_batch(z[None], np.frombuffer(r, np.uint8)[None], np.frombuffer(s, np.uint8)[None])
    import json
    print(json.dumps({"world": world, "n": sp.n, "ms_per_proof": 1e3 * dt, "map_ranks": sp.map_ranks,
                      "matches_single_gpu": bool(not status.any() and want[0].tobytes() == proof)}))
if world > 1:
    dist.destroy_process_group()
