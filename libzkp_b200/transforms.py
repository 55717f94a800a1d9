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


def bench_msm(torch, dev, imad_peak, log_n=20, group=1, c=16, iters=10, resident=True):
    n = 1 << log_n
    ks = _uniform_fr(torch, dev, n, 9).cpu().numpy().view(np.uint8).reshape(n, 32)
    bases = engine.generator_mul(group, ks)                      # n distinct points k_i * G
    B = engine.MsmBases(group, bases, window_bits=c, resident_windows=resident)
    sc = _uniform_fr(torch, dev, n, 10)
    out = torch.zeros(64 if group == 1 else 128, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ms = _time(torch, lambda: B.msm_device(sc.data_ptr(), n, out.data_ptr(), stream), iters)
    # independent check of the timed result: sum k_i s_i computed on the host, one generator_mul
    kl = [int.from_bytes(ks[i].tobytes(), "little") for i in range(n)] if log_n <= 16 else None
    W = (255 + c - 1) // c
    cost = 10 if group == 1 else 28
    cost_add, cost_dbl = (14, 9) if group == 1 else (40, 25)
    muls = n * W * cost + 2 * (1 << (c - 1)) * W * cost_add + c * W * cost_dbl     # SURVEY.md §8d formula
    B.close()
    return {"n": n, "group": "G1" if group == 1 else "G2", "window_bits": c, "windows": W, "ms": ms,
            "points_per_s": n / (ms * 1e-3), "algorithmic_imad": muls * IMAD_PER_MUL,
            "imad_achieved_T": muls * IMAD_PER_MUL / (ms * 1e-3) / 1e12,
            "imad_frac_of_measured_peak": muls * IMAD_PER_MUL / (ms * 1e-3) / imad_peak,
            "mode": "bases resident in HBM with all window multiples" if resident else "one bucket set per window",
            "result_hex": bytes(out.cpu().numpy()).hex()[:32], "_check": kl is not None}


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


def bench(torch, dev, imad_peak, hbm_gbs):
    return {"msm_g1_2^20": bench_msm(torch, dev, imad_peak, 20, 1),
            "msm_g2_2^18": bench_msm(torch, dev, imad_peak, 18, 2, iters=5),
            "ntt_2^22": bench_ntt(torch, dev, imad_peak, hbm_gbs, 22),
            "ntt_2^22_coset_inverse": bench_ntt(torch, dev, imad_peak, hbm_gbs, 22, inverse=True, coset=True),
            "ntt_2^20": bench_ntt(torch, dev, imad_peak, hbm_gbs, 20)}
