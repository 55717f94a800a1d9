"""One process, every visible GPU: lzkp_init(devices) replicates a proving key on each device and one host-buffer
batch call fans out over all of them (reference shape: src/advanced/batch.rs:110-140, one process maps a batch over
its workers).  Checks the bytes against the CPU oracle and prints one JSON line; run by
tests/test_gpu_round2.py::test_one_process_drives_every_gpu in a subprocess (the device list is process-wide)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from libzkp_b200 import engine  # noqa: E402
from oracle import c_oracle as co, zkp_oracle as po  # noqa: E402

G = torch.cuda.device_count()
engine.init(list(range(G)))
assert engine.device_count() == G
td = po.Trapdoor.from_seed(1)
toxic = (td.alpha, td.beta, td.gamma, td.delta, td.tau)
c = int(os.environ.get("LZKP_TEST_WINDOW_BITS", "12"))


def frs(seed, n):
    rng = po.SplitMix64(seed)
    return np.frombuffer(b"".join(rng.next_fr().to_bytes(32, "little") for _ in range(n)), np.uint8).reshape(n, 32).copy()


def busy():
    return [torch.cuda.memory_allocated(d) for d in range(G)]


rep = {"devices": G}
circ = co.Circuit("equality")
pk_bytes, _ = circ.setup(toxic)
opk = co.ProvingKey(pk_bytes)
free0 = [torch.cuda.mem_get_info(d)[0] for d in range(G)]
pk = engine.ProvingKey(pk_bytes, validate=True, window_bits=c)
pk.circuit_builtin(engine.EQUALITY, 110)
free1 = [torch.cuda.mem_get_info(d)[0] for d in range(G)]
rep["all_devices_launched"] = all(a - b > pk.table_bytes * 0.9 for a, b in zip(free0, free1))   # tables on every device

n = 256 * G * 2 + 37
rng = po.SplitMix64(3)
a = np.array([rng.next_u64() for _ in range(n)], np.uint64)
b = a.copy()
b[n - 5] ^= 1                                           # a failing operation in the last device's block
r, s = frs(4, n), frs(40, n)
t0 = time.perf_counter()
proofs, cms, status = pk.prove_equality_batch(a, b, r, s)
rep["equality_ms"] = 1e3 * (time.perf_counter() - t0)
idx = np.unique(np.r_[0:8, np.arange(0, n, 97), n - 8:n])          # samples from every device's block
want, wstat = co.prove_batch(circ, opk, a[idx], b[idx], None, None, r[idx], s[idx])
rep["equality_bit_exact"] = bool(np.array_equal(proofs[idx], want) and np.array_equal(status[idx] != 0, wstat != 0)
                                 and status[n - 5] == 1 and not proofs[n - 5].any()
                                 and all(cms[i].tobytes() == co.mimc_hash(int(a[i])) for i in idx if status[i] == 0))

# explicit assignments (lzkp_prove_batch) through the same fan-out
m = 512
z = np.stack([circ.assign(int(v), int(v)) for v in a[:m]])
p2, st2 = pk.prove_batch(z, r[:m], s[:m])
rep["explicit_z_bit_exact"] = bool(not st2.any() and np.array_equal(p2, proofs[:m]))

# a small call stays on the primary device and is still right
p3, _, st3 = pk.prove_equality_batch(a[:3], a[:3], r[:3], s[:3])
rep["small_call_on_primary_ok"] = bool(not st3.any() and np.array_equal(p3, proofs[:3]))
pk.close()

mc = co.Circuit("membership")
mpk_bytes, _ = mc.setup(toxic)
mpk = engine.ProvingKey(mpk_bytes, window_bits=max(8, c - 2))
mpk.circuit_builtin(engine.MEMBERSHIP, 64)
n = 256 * G + 11
sets = np.zeros((n, 64), np.uint64)
lens = np.zeros(n, np.uint32)
vals = np.zeros(n, np.uint64)
rng = po.SplitMix64(5)
for i in range(n):
    L = 1 + i % 64
    lens[i] = L
    sets[i, :L] = [rng.next_u64() for _ in range(L)]
    vals[i] = sets[i, i % L]
vals[n - 2] = 12345
r, s = frs(7, n), frs(8, n)
env, elen, est = mpk.prove_membership_enveloped(vals, sets, lens, r, s)
proofs, cms, status = mpk.prove_membership_batch(vals, sets, lens, r, s)
idx = np.unique(np.r_[0:4, np.arange(0, n, 61), n - 4:n])
want, wstat = co.prove_batch(mc, co.ProvingKey(mpk_bytes), vals[idx], None, sets[idx], lens[idx], r[idx], s[idx])
ok = np.array_equal(proofs[idx], want) and np.array_equal(status[idx] != 0, wstat != 0) and status[n - 2] == 2
ok = ok and np.array_equal(est, status) and elen[n - 2] == 0
for i in idx:
    if status[i] == 0:
        L = int(lens[i])
        ok = ok and env[i, 14 + 8 * L:14 + 8 * L + 256].tobytes() == proofs[i].tobytes()
rep["membership_bit_exact"] = bool(ok)
mpk.close()
print(json.dumps(rep))
