"""GPU parity tests of the stand-alone NTT and MSM entry points (lzkp_ntt, lzkp_msm_g1/_g2):
bit-exact against the CPU oracle, plus size-independent properties at the benchmark sizes."""
import numpy as np
import pytest

from libzkp_b200 import engine

pytestmark = pytest.mark.gpu


def test_ntt16_known_answers(golden, co):
    g = golden["ntt16"]
    arr = co.fr_array([int(x) for x in g["in"]])
    for name, (inv, coset) in {"fft": (0, 0), "ifft": (1, 0), "coset_fft": (0, 1), "coset_ifft": (1, 1)}.items():
        assert co.fr_list(engine.ntt(arr, inverse=inv, coset=coset)) == [int(x) for x in g[name]], name


@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 9, 10, 12, 13, 14, 16, 17, 20])
def test_ntt_matches_oracle(co, frs, log_n):
    a = frs(200 + log_n, 1 << log_n)
    for inv in (False, True):
        for coset in (False, True):
            assert np.array_equal(engine.ntt(a, inverse=inv, coset=coset), co.ntt(a, inverse=inv, coset=coset)), \
                (log_n, inv, coset)


def test_ntt_2_22_properties(co, frs):
    # BASELINE size: round trip, linearity, and the oracle on the same input
    n = 1 << 22
    a, b = frs(11, n), frs(12, n)
    fa = engine.ntt(a)
    assert np.array_equal(engine.ntt(fa, inverse=True), a)
    assert np.array_equal(engine.ntt(engine.ntt(a, coset=True), inverse=True, coset=True), a)
    assert np.array_equal(fa, co.ntt(a))
    # linearity on a slice of outputs: NTT(a + b) = NTT(a) + NTT(b) mod r
    R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    s = co.fr_array([(x + y) % R for x, y in zip(co.fr_list(a[:4096]), co.fr_list(b[:4096]))])
    full = np.concatenate([s, np.zeros((n - 4096, 32), np.uint8)])
    a2 = np.concatenate([a[:4096], np.zeros((n - 4096, 32), np.uint8)])
    b2 = np.concatenate([b[:4096], np.zeros((n - 4096, 32), np.uint8)])
    fs, f1, f2 = engine.ntt(full)[:64], engine.ntt(a2)[:64], engine.ntt(b2)[:64]
    assert co.fr_list(fs) == [(x + y) % R for x, y in zip(co.fr_list(f1), co.fr_list(f2))]


def test_msm8_known_answers(golden, co):
    g = golden["msm8"]
    sc = co.fr_array([int(x) for x in g["scalars"]])
    b1 = np.frombuffer(bytes.fromhex("".join(g["g1_bases"])), np.uint8)
    b2 = np.frombuffer(bytes.fromhex("".join(g["g2_bases"])), np.uint8)
    assert engine.msm_g1(b1, sc).hex() == g["g1"]
    assert engine.msm_g2(b2, sc).hex() == g["g2"]


@pytest.mark.parametrize("n", [0, 1, 2, 31, 333, 1000, 5000, 1 << 15])
def test_msm_g1_matches_oracle(co, frs, n):
    bases = co.g1_gen_mul(frs(9, max(n, 1)))[:n]
    sc = frs(10, max(n, 1))[:n]
    if n > 10:
        sc[3] = 0
        sc[4, :] = 0
        sc[4, 0] = 1
        bases[5] = 0
        bases[5, 63] = 0x40                                  # identity base
        bases[7] = bases[6]                                  # repeated point (P + P in a bucket)
        sc[7] = sc[6]
        bases[9] = bases[8]
        bases[9, 32:64] = np.frombuffer(
            ((0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
              - int.from_bytes(bases[8, 32:64].tobytes(), "little") % (1 << 254)) % (1 << 256)).to_bytes(32, "little"),
            np.uint8)                                        # P and -P with the same scalar
        bases[9, 63] &= 0x3F
        sc[9] = sc[8]
    assert engine.msm_g1(bases, sc) == co.msm_g1(bases, sc) if n else engine.msm_g1(bases, sc)[63] == 0x40


@pytest.mark.parametrize("n", [0, 1, 7, 333, 3000])
def test_msm_g2_matches_oracle(co, frs, n):
    bases = co.g2_gen_mul(frs(19, max(n, 1)))[:n]
    sc = frs(20, max(n, 1))[:n]
    if n:
        assert engine.msm_g2(bases, sc) == co.msm_g2(bases, sc)
    else:
        assert engine.msm_g2(bases, sc)[127] == 0x40


def test_msm_witness_like_scalars(co, frs):
    # membership-style scalars: mostly 0 / 1 / small u64
    n = 4096
    bases = co.g1_gen_mul(frs(9, n))
    sc = frs(10, n)
    sc[::2] = 0
    sc[1::4, 1:] = 0
    sc[1::4, 0] = 1
    sc[3::8, 8:] = 0
    assert engine.msm_g1(bases, sc) == co.msm_g1(bases, sc)


def test_generator_mul_matches_oracle(co, frs):
    ks = frs(9, 200)
    ks[3] = 0
    assert np.array_equal(engine.generator_mul(1, ks), co.g1_gen_mul(ks))
    assert np.array_equal(engine.generator_mul(2, ks[:50]), co.g2_gen_mul(ks[:50]))


@pytest.mark.parametrize("group,n,c", [(1, 3000, 16), (1, 3000, 11), (2, 700, 16), (1, 1, 16), (1, 40000, 13)])
def test_resident_bases_msm(co, frs, group, n, c):
    gen = co.g1_gen_mul if group == 1 else co.g2_gen_mul
    ora = co.msm_g1 if group == 1 else co.msm_g2
    small = frs(9, min(n, 4096))
    pts = gen(small)
    bases = np.concatenate([pts] * ((n + len(pts) - 1) // len(pts)))[:n]      # repeated points: P + P in a bucket
    sc = frs(10, n)
    B = engine.MsmBases(group, bases, window_bits=c, resident_windows=True, validate=True)
    assert B.msm(sc) == ora(bases, sc)
    sc2 = frs(11, n)
    sc2[::3] = 0
    assert B.msm(sc2) == ora(bases, sc2)                      # workspace reuse across calls
    if n > 10:
        assert B.msm(sc[: n // 2]) == ora(bases[: n // 2], sc[: n // 2])    # fewer scalars than bases (msm truncates)
    B.close()


def test_msm_skewed_scalars_long_runs(co, frs):
    # half of the scalars equal 1 and a quarter equal r - 1: giant buckets, the k_long_run path
    n = 1 << 16
    pts = co.g1_gen_mul(frs(9, 2048))
    bases = np.concatenate([pts] * (n // 2048))
    sc = frs(10, n)
    R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    sc[0::2] = np.frombuffer((1).to_bytes(32, "little"), np.uint8)
    sc[1::4] = np.frombuffer((R - 1).to_bytes(32, "little"), np.uint8)
    want = co.msm_g1(bases, sc)
    for resident in (True, False):
        B = engine.MsmBases(1, bases, resident_windows=resident)
        assert B.msm(sc) == want
        B.close()


def test_msm_rejects_bad_input(co, frs):
    bases = co.g1_gen_mul(frs(9, 16))
    bad = bases.copy()
    bad[3, 0] ^= 1                                            # off the curve
    with pytest.raises(Exception):
        engine.MsmBases(1, bad, validate=True)
    B = engine.MsmBases(1, bases)
    sc = frs(10, 16)
    sc[5] = 0xFF                                              # >= r
    with pytest.raises(Exception):
        B.msm(sc)
    B.close()


def test_msm_2_20_trapdoor(co, po, frs):
    # BASELINE size.  Bases k_i * G with known k_i (2^14 distinct points tiled 64 times) =>
    # MSM == (sum k_i s_i) * G: a check independent of any MSM code, resident and one-shot modes.
    n, m = 1 << 20, 1 << 14
    ks, sc = frs(9, m), frs(10, n)
    pts = engine.generator_mul(1, ks)
    assert np.array_equal(pts[:64], co.g1_gen_mul(ks[:64]))
    bases = np.concatenate([pts] * (n // m))
    kl, sl = co.fr_list(ks), co.fr_list(sc)
    tot = sum(kl[i % m] * s for i, s in enumerate(sl)) % po.R_MOD
    want = po.g1_to_bytes(po.G1.mul(po.G1_GEN, tot))
    B = engine.MsmBases(1, bases)
    assert B.msm(sc) == want
    B.close()
    assert engine.msm_g1(bases, sc) == want
    assert co.msm_g1(bases, sc) == want
