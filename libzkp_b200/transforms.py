"""Device-resident timing of the two stand-alone transforms BASELINE.json names next to proofs/s:
G1 MSM points/s at 2^20 and NTT elements/s at 2^22 (used by bench.py's `extra` object)."""
from __future__ import annotations

import numpy as np

from . import engine

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
IMAD_PER_MUL = 272


def _uniform_fr(torch, dev, n, seed):
    """n uniform canonical Fr elements on the device (rejection-free: top limb masked below r's top limb)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    x = torch.randint(0, 2**31 - 1, (n, 8), generator=g, device=dev, dtype=torch.int64)
    y = torch.randint(0, 2, (n, 8), generator=g, device=dev, dtype=torch.int64)
    v = (x * 2 + y) & 0xFFFFFFFF
    v[:, 7] = v[:, 7] % 0x30644E72            # strictly below the top limb of r  =>  value < r
    return v.to(torch.int32).contiguous()      # little-endian 32-bit limbs == 32 B canonical LE


def _time(torch, fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_msm(torch, dev, imad_peak, log_n=20, group=1, c=16, iters=10, resident=True, witness_like=False):
    n = 1 << log_n
    ks = _uniform_fr(torch, dev, n, 9).cpu().numpy().view(np.uint8).reshape(n, 32)
    bases = engine.generator_mul(group, ks)                      # n distinct points k_i * G
    B = engine.MsmBases(group, bases, window_bits=c, resident_windows=resident)
    sc = _uniform_fr(torch, dev, n, 10)
    if witness_like:                    # half of the scalars are 0 or 1 (booleans, unused wires)
        sc[0::4] = 0
        sc[1::4] = 0
        sc[1::4, 0] = 1
    out = torch.zeros(64 if group == 1 else 128, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ms = _time(torch, lambda: B.msm_device(sc.data_ptr(), n, out.data_ptr(), stream), iters)
    # independent check of the timed result: the bases are k_i * G, so the MSM must equal (sum k_i s_i mod r) * G -
    # host big-integer arithmetic and ONE fixed-base multiplication, nothing shared with the Pippenger pipeline
    sc_h = sc.cpu().numpy().view(np.uint8).reshape(n, 32)
    kw, sw = ks.view("<u8").reshape(n, 4), sc_h.view("<u8").reshape(n, 4)
    tot = 0
    for i in range(n):
        k = int(kw[i, 0]) | int(kw[i, 1]) << 64 | int(kw[i, 2]) << 128 | int(kw[i, 3]) << 192
        t = int(sw[i, 0]) | int(sw[i, 1]) << 64 | int(sw[i, 2]) << 128 | int(sw[i, 3]) << 192
        tot += k * t
    want = engine.generator_mul(group, np.frombuffer((tot % R_MOD).to_bytes(32, "little"), np.uint8).reshape(1, 32))[0]
    checked = bool(np.array_equal(want, out.cpu().numpy()))
    assert checked, "timed MSM result differs from (sum k_i s_i) * G"
    W = (255 + c - 1) // c
    cost = 10 if group == 1 else 28
    cost_add, cost_dbl = (14, 9) if group == 1 else (40, 25)
    # SURVEY.md 8d formula; with witness-like scalars only the mixed additions actually performed are counted:
    # a zero scalar has no non-zero digit, a scalar 1 has one, a uniform scalar W of them
    madds = n * W if not witness_like else (n // 2) * W + n // 4
    muls = madds * cost + 2 * (1 << (c - 1)) * W * cost_add + c * W * cost_dbl
    B.close()
    return {"n": n, "group": "G1" if group == 1 else "G2", "window_bits": c, "windows": W, "ms": ms,
            "points_per_s": n / (ms * 1e-3), "algorithmic_imad": muls * IMAD_PER_MUL,
            "imad_achieved_T": muls * IMAD_PER_MUL / (ms * 1e-3) / 1e12,
            "imad_frac_of_measured_peak": muls * IMAD_PER_MUL / (ms * 1e-3) / imad_peak,
            "mode": "bases resident in HBM with all window multiples" if resident else
            ("one bucket set per window, GLV split: 2n points x 128-bit scalars" if group == 1 else "one bucket set per window"),
            "scalars": "50% zero/one, 50% uniform" if witness_like else "uniform in [0, r)",
            "mixed_adds_counted": madds, "result_hex": bytes(out.cpu().numpy()).hex()[:32],
            "result_equals_sum_k_s_times_G": checked}


def bench_ntt(torch, dev, imad_peak, hbm_gbs, log_n=22, iters=10, inverse=False, coset=False):
    n = 1 << log_n
    a = _uniform_fr(torch, dev, n, 11)
    b = torch.empty_like(a)
    stream = torch.cuda.current_stream().cuda_stream
    bufs = [a, b]

    def step():
        engine.ntt_device(bufs[0].data_ptr(), bufs[1].data_ptr(), log_n, inverse, coset, stream)
        bufs.reverse()
    ms = _time(torch, step, iters)
    imad = (n // 2) * log_n * IMAD_PER_MUL
    return {"n": n, "inverse": inverse, "coset": coset, "ms": ms, "elements_per_s": n / (ms * 1e-3),
            "algorithmic_imad": imad, "imad_achieved_T": imad / (ms * 1e-3) / 1e12,
            "imad_frac_of_measured_peak": imad / (ms * 1e-3) / imad_peak,
            "algorithmic_bytes": 64 * n, "hbm_achieved_gbs": 64 * n / (ms * 1e-3) / 1e9,
            "hbm_frac_of_measured_peak": 64 * n / (ms * 1e-3) / 1e9 / hbm_gbs}


def _toxic(seed=1):
    # SplitMix64 -> five canonical non-zero scalars (synthetic benchmark key; never use a seeded key in production)
    st = [seed & 0xFFFFFFFFFFFFFFFF]

    def nxt():
        st[0] = (st[0] + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = st[0]
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)
    out = []
    while len(out) < 5:
        v = sum(nxt() << (64 * i) for i in range(4)) & ((1 << 254) - 1)
        if 0 < v < R_MOD:
            out.append(v)
    return out


def bench_large_proof(torch, dev, rounds=349524, iters=5):
    """BASELINE.json configs[3]: one proof of the 2^20-constraint MiMC-chain circuit, inputs resident in HBM."""
    import time
    t0 = time.perf_counter()
    pk_bytes, _ = engine.setup_builtin(engine.EQUALITY, rounds, _toxic(1))
    t_setup = time.perf_counter() - t0
    t0 = time.perf_counter()
    pk = engine.ProvingKey(pk_bytes)
    pk.circuit_builtin(engine.EQUALITY, rounds)
    torch.cuda.synchronize()
    t_load = time.perf_counter() - t0
    z = torch.from_numpy(engine.builtin_witness(engine.EQUALITY, rounds, 6, 6)).to(dev)
    rs = _uniform_fr(torch, dev, 2, 4)
    proof = torch.zeros(256, dtype=torch.uint8, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    engine.profile_enable(True)
    ms = _time(torch, lambda: pk.prove_batch_device(1, z.data_ptr(), rs[0].data_ptr(), rs[1].data_ptr(), proof.data_ptr(),
                                                    status.data_ptr(), stream), iters, warmup=2)
    reg = pk.profile_read(reset=True)
    engine.profile_enable(False)
    out = {"constraints": 3 * rounds + 2, "domain": pk.n, "ms_per_proof": ms, "proofs_per_s": 1e3 / ms,
           "stage_ms": {k: v[0] / max(v[1], 1) for k, v in reg.items() if v[1]}, "setup_s": t_setup, "pk_load_s": t_load,
           "pk_mib": len(pk_bytes) / 2**20, "resident_bases_gib": pk.table_bytes / 2**30, "status": int(status.item())}
    pk.close()
    return out


def bench_membership(torch, dev, n=1024, iters=5, slots=64):
    """BASELINE.json configs[2]: batch of 1024 membership proofs through the host-buffer C ABI, at the reference's
    MAX_SET_SIZE = 64 and at the 1024-slot size the config names (m = 5453, n = 8192; the reference itself rejects
    sets above 64, SURVEY.md headline fact 6)."""
    import time
    pk_bytes, _ = engine.setup_builtin(engine.MEMBERSHIP, slots, _toxic(1))
    pk = engine.ProvingKey(pk_bytes)
    pk.circuit_builtin(engine.MEMBERSHIP, slots)
    rng = np.random.default_rng(5)
    sets = rng.integers(0, 2**63, size=(n, slots), dtype=np.uint64)
    lens = np.full(n, slots, np.uint32)
    vals = sets[np.arange(n), np.arange(n) % slots].copy()
    r = _uniform_fr(torch, dev, n, 7).cpu().numpy().view(np.uint8).reshape(n, 32)
    s = _uniform_fr(torch, dev, n, 8).cpu().numpy().view(np.uint8).reshape(n, 32)
    for _ in range(2):
        proofs, _, status = pk.prove_membership_batch(vals, sets, lens, r, s)
    t0 = time.perf_counter()
    for _ in range(iters):
        proofs, _, status = pk.prove_membership_batch(vals, sets, lens, r, s)
    dt = (time.perf_counter() - t0) / iters
    out = {"batch": n, "set_slots": slots, "domain": pk.n, "max_chunk": pk.max_chunk, "ms_per_batch": 1e3 * dt, "proofs_per_s_e2e": n / dt, "failed": int((status != 0).sum()),
           "table_gb": pk.table_bytes / 1e9, "window_bits": pk.window_bits}
    pk.close()
    return out


def bench_mixed(torch, dev, n=8192, iters=3):
    """BASELINE.json configs[4] per-GPU share: a mixed batch (even index equality, odd index membership with 64 slots),
    both proving keys resident at once, grouped per circuit as process_batch does; host buffers, end to end."""
    import time
    pk_e = engine.ProvingKey(engine.setup_builtin(engine.EQUALITY, 110, _toxic(1))[0])
    pk_e.circuit_builtin(engine.EQUALITY, 110)
    pk_m = engine.ProvingKey(engine.setup_builtin(engine.MEMBERSHIP, 64, _toxic(1))[0])
    pk_m.circuit_builtin(engine.MEMBERSHIP, 64)
    half = n // 2
    rng = np.random.default_rng(7)
    a = rng.integers(0, 2**63, size=half, dtype=np.uint64)
    sets = rng.integers(0, 2**63, size=(half, 64), dtype=np.uint64)
    lens = np.full(half, 64, np.uint32)
    vals = sets[np.arange(half), np.arange(half) % 64].copy()
    fr = lambda seed: _uniform_fr(torch, dev, half, seed).cpu().numpy().view(np.uint8).reshape(half, 32)
    r1, s1, r2, s2 = fr(1), fr(2), fr(3), fr(4)

    def once():
        pe, _, se = pk_e.prove_equality_batch(a, a, r1, s1)
        pm, _, sm = pk_m.prove_membership_batch(vals, sets, lens, r2, s2)
        return int((se != 0).sum() + (sm != 0).sum())
    once()
    t0 = time.perf_counter()
    for _ in range(iters):
        failed = once()
    dt = (time.perf_counter() - t0) / iters
    out = {"batch": n, "ms_per_batch": 1e3 * dt, "proofs_per_s_e2e": n / dt, "failed": failed,
           "tables_gb": (pk_e.table_bytes + pk_m.table_bytes) / 1e9, "window_bits": [pk_e.window_bits, pk_m.window_bits]}
    pk_e.close()
    pk_m.close()
    return out


def bench_verify(torch, dev, n=4096, iters=3):
    """Batched device verification (lzkp_verify_batch) of n equality proofs produced by the engine, host buffers."""
    import time
    pk_bytes, vk_bytes = engine.setup_builtin(engine.EQUALITY, 110, _toxic(1))
    pk = engine.ProvingKey(pk_bytes, window_bits=12)
    pk.circuit_builtin(engine.EQUALITY, 110)
    rng = np.random.default_rng(3)
    a = rng.integers(0, 2**63, size=n, dtype=np.uint64)
    fr = lambda seed: _uniform_fr(torch, dev, n, seed).cpu().numpy().view(np.uint8).reshape(n, 32)
    proofs, cms, st = pk.prove_equality_batch(a, a, fr(1), fr(2))
    pk.close()
    vk = engine.VerifyingKey(vk_bytes)
    out = {}
    for nb in ((1, 64, 512, n) if n <= 4096 else (1, n)):
        ok = vk.verify_batch(proofs[:nb], cms[:nb])
        t0 = time.perf_counter()
        for _ in range(iters):
            ok = vk.verify_batch(proofs[:nb], cms[:nb])
        dt = (time.perf_counter() - t0) / iters
        out[f"batch_{nb}"] = {"ms": 1e3 * dt, "verifies_per_s": nb / dt, "all_accepted": bool(ok.all())}
    bad = cms.copy()
    bad[:, 0] ^= 1
    out["wrong_inputs_rejected"] = bool(not vk.verify_batch(proofs[:64], bad[:64]).any())
    vk.close()
    return out


def bench(torch, dev, imad_peak, hbm_gbs):
    jobs = {"msm_g1_2^20": lambda: bench_msm(torch, dev, imad_peak, 20, 1),
            "msm_g1_2^20_witness_like": lambda: bench_msm(torch, dev, imad_peak, 20, 1, witness_like=True),
            "msm_g1_2^20_one_shot_bases": lambda: bench_msm(torch, dev, imad_peak, 20, 1, resident=False),
            "msm_g2_2^18": lambda: bench_msm(torch, dev, imad_peak, 18, 2, iters=5),
            "ntt_2^22": lambda: bench_ntt(torch, dev, imad_peak, hbm_gbs, 22),
            "ntt_2^22_coset_inverse": lambda: bench_ntt(torch, dev, imad_peak, hbm_gbs, 22, inverse=True, coset=True),
            "ntt_2^20": lambda: bench_ntt(torch, dev, imad_peak, hbm_gbs, 20),
            "membership_batch_1024": lambda: bench_membership(torch, dev),
            "membership_1024_slots_batch_1024": lambda: bench_membership(torch, dev, iters=2, slots=1024),
            "mixed_batch_8192": lambda: bench_mixed(torch, dev),
            "verify_batch": lambda: bench_verify(torch, dev),
            "verify_batch_65536_rlc": lambda: bench_verify(torch, dev, n=65536, iters=2),
            "proof_2^20_constraints": lambda: bench_large_proof(torch, dev)}
    out = {}
    for name, fn in jobs.items():           # one failing workload is recorded under its key, the others still run
        try:
            out[name] = fn()
        except Exception as e:              # noqa: BLE001
            out[name] = {"error": repr(e)}
    return out
