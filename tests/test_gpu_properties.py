"""Property-based GPU parity checks (hypothesis): ragged sizes and adversarial scalar / base patterns for the
MSM and NTT entry points against the CPU oracle, plus algebraic properties that hold at any size."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from libzkp_b200 import engine

pytestmark = pytest.mark.gpu

R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
SPECIAL = [0, 1, 2, R - 1, R - 2, (R - 1) // 2, 1 << 64, (1 << 128) - 1, 1 << 253]
scalar = st.one_of(st.sampled_from(SPECIAL), st.integers(0, R - 1), st.integers(0, 2**16))
COMMON = dict(deadline=None, max_examples=25, suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow])


@pytest.fixture(scope="module")
def points(co, frs):
    ks = frs(9, 64)
    g1, g2 = co.g1_gen_mul(ks), co.g2_gen_mul(ks)
    g1[7] = 0
    g1[7, 63] = 0x40                     # identity
    g2[7] = 0
    g2[7, 127] = 0x40
    return g1, g2


@settings(**COMMON)
@given(data=st.data())
def test_msm_g1_ragged(co, points, data):
    n = data.draw(st.integers(0, 200))
    idx = data.draw(st.lists(st.integers(0, 63), min_size=n, max_size=n))
    sc = co.fr_array(data.draw(st.lists(scalar, min_size=n, max_size=n))) if n else np.zeros((0, 32), np.uint8)
    bases = points[0][idx] if n else np.zeros((0, 64), np.uint8)
    got = engine.msm_g1(bases, sc)
    want = co.msm_g1(bases, sc) if n else bytes(63) + b"\x40"
    assert got == want
    if n:
        B = engine.MsmBases(1, bases, window_bits=data.draw(st.sampled_from([8, 12, 16])), resident_windows=True)
        assert B.msm(sc) == want
        B.close()


@settings(**{**COMMON, "max_examples": 10})
@given(data=st.data())
def test_msm_g2_ragged(co, points, data):
    n = data.draw(st.integers(1, 60))
    idx = data.draw(st.lists(st.integers(0, 63), min_size=n, max_size=n))
    sc = co.fr_array(data.draw(st.lists(scalar, min_size=n, max_size=n)))
    bases = points[1][idx]
    assert engine.msm_g2(bases, sc) == co.msm_g2(bases, sc)


@settings(**COMMON)
@given(data=st.data())
def test_ntt_variants(co, data):
    log_n = data.draw(st.integers(0, 10))
    n = 1 << log_n
    a = co.fr_array(data.draw(st.lists(scalar, min_size=n, max_size=n)))
    inv, coset = data.draw(st.booleans()), data.draw(st.booleans())
    assert np.array_equal(engine.ntt(a, inverse=inv, coset=coset), co.ntt(a, inverse=inv, coset=coset))
    assert np.array_equal(engine.ntt(engine.ntt(a, coset=coset), inverse=True, coset=coset), a)   # round trip


def test_msm_linearity_at_2_16(co, frs):
    # MSM(s + t) == MSM(s) + MSM(t): compare through a third MSM over the two results with scalars (1, 1)
    n = 1 << 16
    pts = engine.generator_mul(1, frs(9, n))
    s, t = frs(10, n), frs(11, n)
    st_sum = co.fr_array([(x + y) % R for x, y in zip(co.fr_list(s), co.fr_list(t))])
    B = engine.MsmBases(1, pts)
    a, b, c = B.msm(s), B.msm(t), B.msm(st_sum)
    B.close()
    one = co.fr_array([1, 1])
    assert engine.msm_g1(np.frombuffer(a + b, np.uint8).reshape(2, 64), one) == c
