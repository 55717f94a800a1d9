"""CPU-side checks of the product: the C-ABI library loads and exports exactly what
include/lzkp_b200.h declares, computing entry points fail loudly without a device, the native
circuit synthesis / MiMC / field arithmetic agree with the oracle, and the host mirror keeps the
reference's API behaviour (envelope, validation errors, batch registry)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT, has_gpu

import libzkp_b200 as zk
from libzkp_b200 import _ffi, batch, engine, proof, snark


def header_symbols():
    src = open(os.path.join(ROOT, "include", "lzkp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lzkp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _ffi.lib()
    names = header_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/lzkp_b200.h but not exported"
    assert sorted(_ffi.SIGNATURES) == names              # the Python binding covers the whole header
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (lzkp_[a-z0-9_]+)", out)))
    assert exported == names                             # and nothing undeclared leaks out


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback():
    with pytest.raises(zk.EngineError) as e:
        engine.init()
    assert e.value.code == _ffi.LZKP_E_NO_DEVICE
    with pytest.raises(zk.EngineError):
        engine.ntt(np.zeros((8, 32), np.uint8))
    with pytest.raises(zk.EngineError):
        engine.msm_g1(np.zeros((1, 64), np.uint8), np.zeros((1, 32), np.uint8))
    with pytest.raises(zk.EngineError):
        engine.ProvingKey(b"\x00" * 100)
    # the reference's failure convention: empty Vec -> ProofGenerationFailed -> RuntimeError
    snark.reset()
    snark.configure(generator=lambda prefix: (b"", b""))
    try:
        with pytest.raises(zk.ProofGenerationFailed):
            zk.prove_equality(5, 5)
        assert isinstance(zk.ProofGenerationFailed("x"), RuntimeError)
    finally:
        snark.reset()
        snark.configure()


def test_commit_value_snark_known_answers(golden):
    for v, row in golden["mimc"].items():
        assert zk.snark_commit_value(int(v)).hex() == row["commitment"]
    assert zk.snark_commit_value(42) != zk.snark_commit_value(43)          # snark.rs:618-623


@pytest.mark.parametrize("kind,param,name", [(0, 110, "equality"), (1, 64, "membership"), (1, 5, "membership"),
                                             (0, 7, "equality")])
def test_builtin_circuit_equals_oracle(co, kind, param, name):
    (m, n_inst, n_wit), mats = engine.builtin_circuit_csr(kind, param)
    c = co.Circuit(name, param)
    assert (m, n_inst, n_wit) == (c.m, c.n_inst, c.n_wit)
    for which in range(3):
        rowptr, col, val = c.matrix(which)
        assert np.array_equal(mats[which][0], rowptr)
        assert np.array_equal(mats[which][1], col)
        assert np.array_equal(mats[which][2], val)


@pytest.fixture(scope="module")
def field_shim(tmp_path_factory):
    out = tmp_path_factory.mktemp("shim") / "field_host.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-o", str(out),
                    os.path.join(ROOT, "tests", "host_shim", "field_host.cpp")], check=True)
    return C.CDLL(str(out))


def test_device_field_composition_on_host(field_shim, po):
    """field.cuh's Montgomery row composition (portable bodies) against big-int arithmetic."""
    R = 1 << 256
    rng = po.SplitMix64(77)
    for mod, fn in ((po.R_MOD, field_shim.fr_op), (po.Q_MOD, field_shim.fq_op)):
        def op(code, a, b=0):
            o = (C.c_uint8 * 32)()
            fn(code, (C.c_uint8 * 32)(*a.to_bytes(32, "little")), (C.c_uint8 * 32)(*b.to_bytes(32, "little")), o)
            return int.from_bytes(bytes(o), "little")
        rinv = pow(R, -1, mod)
        vals = [0, 1, mod - 1, mod - 2, R % mod] + [rng.next_fr() % mod for _ in range(200)]
        for i, a in enumerate(vals):
            b = vals[(i * 7 + 3) % len(vals)]
            assert op(0, a, b) == a * b * rinv % mod
            assert op(7, a) == a * a * rinv % mod
            assert op(1, a, b) == (a + b) % mod
            assert op(2, a, b) == (a - b) % mod
            assert op(3, a) == (-a) % mod
            assert op(4, a) == a * R % mod
            assert op(5, a) == a * rinv % mod
        for a in vals[1:20]:
            am = a * R % mod
            assert op(6, am) == pow(a, -1, mod) * R % mod
    assert field_shim.fq_gt((C.c_uint8 * 32)(*(5).to_bytes(32, "little")), (C.c_uint8 * 32)(*(4).to_bytes(32, "little"))) == 1
    assert field_shim.fq_gt((C.c_uint8 * 32)(*(4).to_bytes(32, "little")), (C.c_uint8 * 32)(*(4).to_bytes(32, "little"))) == 0


def test_lazy_fq2_product_on_host(field_shim, po):
    """field.cuh's double-width path (mul_wide, reduce_wide, fq2_mul_lazy: Karatsuba with two reductions)."""
    q, R = po.Q_MOD, 1 << 256
    rinv = pow(R, -1, q)
    rng = po.SplitMix64(91)
    b32 = lambda v: (C.c_uint8 * 32)(*v.to_bytes(32, "little"))
    b64 = lambda v: (C.c_uint8 * 64)(*v.to_bytes(64, "little"))
    edge = [0, 1, q - 1, q - 2, R % q, (1 << 254) % q, 0xffffffff, q >> 1]
    # operands that stress the Karatsuba halves (any 256-bit value is a legal input of the wide products)
    halves = [(1 << 128) - 1, 1 << 128, ((1 << 128) - 1) << 128, (1 << 128) + 1, ((1 << 127) << 128) | 1,
              (0xffffffff << 128) | 0xfffffffe, (1 << 256) - 1, ((1 << 128) - 1) << 128 | 1]
    vals = edge + [rng.next_fr() % q for _ in range(300)]
    wv = vals + halves
    for i, a in enumerate(wv):                         # 16-limb product, also of unreduced operands (sums < 2^256)
        for b in (wv[(i * 5 + 1) % len(wv)], (1 << 256) - 1, 2 * q - 2, halves[i % len(halves)]):
            o = (C.c_uint8 * 64)()
            field_shim.fq_mul_wide_host(b32(a), b32(b), o)
            assert int.from_bytes(bytes(o), "little") == a * b
            field_shim.fq_mul_wide_k_host(b32(a), b32(b), o)            # Karatsuba (three 128-bit products)
            assert int.from_bytes(bytes(o), "little") == a * b, (hex(a), hex(b))
    wide = [0, 1, q * R - 1, q * R - q, (q - 1) * (q - 1), 2 * (q - 1) * (q - 1), R - 1, R, q] + \
           [rng.next_fr() * rng.next_fr() % (q * R) for _ in range(300)]
    for t in wide:                                     # reduction of any T < q * 2^256
        o = (C.c_uint8 * 32)()
        field_shim.fq_reduce_wide_host(b64(t), o)
        assert int.from_bytes(bytes(o), "little") == t * rinv % q
    for i in range(len(vals)):                         # Fq2 product in Montgomery form: (a0 + a1 u)(b0 + b1 u), u^2 = -1
        a0, a1, b0, b1 = (vals[(i * k + k) % len(vals)] for k in (1, 3, 7, 11))
        o = (C.c_uint8 * 64)()
        field_shim.fq2_mul_lazy_host((C.c_uint8 * 64)(*(a0.to_bytes(32, "little") + a1.to_bytes(32, "little"))),
                                     (C.c_uint8 * 64)(*(b0.to_bytes(32, "little") + b1.to_bytes(32, "little"))), o)
        c0, c1 = int.from_bytes(bytes(o[:32]), "little"), int.from_bytes(bytes(o[32:]), "little")
        assert c0 == (a0 * b0 - a1 * b1) * rinv % q and c1 == (a0 * b1 + a1 * b0) * rinv % q
    # a*b - c*d with the reductions shared (the Y3 of the XYZZ formulas), extremes of both signs included
    pick = lambda i, k: vals[(i * k + k) % len(vals)]
    cases = [(q - 1, q - 1, 0, 0), (0, 0, q - 1, q - 1), (q - 1, q - 1, q - 1, q - 1), (1, 0, q - 1, q - 1)] + \
            [(pick(i, 1), pick(i, 3), pick(i, 7), pick(i, 11)) for i in range(len(vals))]
    for a, b, c, d in cases:
        o = (C.c_uint8 * 32)()
        field_shim.fq_msub_host(b32(a), b32(b), b32(c), b32(d), o)
        assert int.from_bytes(bytes(o), "little") == (a * b - c * d) * rinv % q


def test_windowed_scalar_mul_on_host(field_shim, po):
    """ec.cuh's signed 4-bit window scalar multiplication against the oracle's double-and-add."""
    G = po.G1_GEN
    rng = po.SplitMix64(321)
    b32 = lambda v: (C.c_uint8 * 32)(*v.to_bytes(32, "little"))
    def mul(pt, k, nbits):
        o = (C.c_uint8 * 64)()
        field_shim.g1_scalar_mul_host((C.c_uint8 * 64)(*(pt[0].to_bytes(32, "little") + pt[1].to_bytes(32, "little"))), b32(k), nbits, o)
        x, y = int.from_bytes(bytes(o[:32]), "little"), int.from_bytes(bytes(o[32:]), "little")
        return None if x == 0 and y == 0 else (x, y)
    P = po.G1.mul(G, 0x1234567)
    # digits of every shape: zero, carries running through all windows (0x88.., 0xff..), top-digit carry, r - 1
    full = [0, 1, 7, 8, 9, 15, 16, po.R_MOD - 1, po.R_MOD - 2, (1 << 253) + 5, int("8" * 63, 16) % po.R_MOD] + \
           [rng.next_fr() for _ in range(6)]
    for k in full:
        assert mul(P, k, 254) == po.G1.mul(P, k), hex(k)
    half = [0, 1, 8, (1 << 128) - 1, (1 << 127), int("8" * 32, 16), int("7" * 32, 16)] + [rng.next_fr() >> 127 for _ in range(6)]
    for k in half:
        assert mul(P, k, 128) == po.G1.mul(P, k), hex(k)


def test_glv_split_on_host(field_shim, po):
    """ec.cuh's GLV split: k == +-k1 +- k2 * lambda (mod r) with both halves below 2^128."""
    lam = 0x30644e72e131a029048b6e193fd84104cc37a73fec2bc5e9b8ca0b2d36636f23
    beta = 0x30644e72e131a0295e6dd9e7e0acccb0c28f069fbb966e3de4bd44e5607cfd48
    assert (lam * lam + lam + 1) % po.R_MOD == 0 and pow(beta, 3, po.Q_MOD) == 1 and beta != 1
    G = po.G1_GEN
    assert po.G1.mul(G, lam) == (beta * G[0] % po.Q_MOD, G[1])         # phi(P) = lambda * P
    rng = po.SplitMix64(123)
    ks = [0, 1, 2, po.R_MOD - 1, po.R_MOD - 2, lam, lam + 1, po.R_MOD // 2, 1 << 253] + [rng.next_fr() for _ in range(3000)]
    for k in ks:
        out = (C.c_uint8 * 32)()
        fl = field_shim.glv_split_host((C.c_uint8 * 32)(*k.to_bytes(32, "little")), out)
        assert fl & 4, "split must fit 128 bits"
        k1 = int.from_bytes(bytes(out[:16]), "little") * (-1 if fl & 1 else 1)
        k2 = int.from_bytes(bytes(out[16:]), "little") * (-1 if fl & 2 else 1)
        assert (k1 + k2 * lam - k) % po.R_MOD == 0


# ---------------------------------------------------------------- host mirror behaviour
def test_envelope_layout_and_limits():
    p = proof.Proof(2, b"\x01" * 256, b"\x02" * 32)
    b = p.to_bytes()
    assert len(b) == 298 and b[0] == 2 and b[1] == 2          # proof/mod.rs:23-36
    assert b[2:6] == (256).to_bytes(4, "little") and b[6:10] == (32).to_bytes(4, "little")
    q = proof.Proof.from_bytes(b)
    assert (q.version, q.scheme, q.proof, q.commitment) == (2, 2, p.proof, p.commitment)
    with pytest.raises(zk.InvalidProofFormat):
        proof.Proof.from_bytes(b[:9])
    with pytest.raises(zk.InvalidProofFormat):
        proof.Proof.from_bytes(b + b"\x00")
    with pytest.raises(TypeError):                             # PyO3 maps InvalidProofFormat -> TypeError
        proof.Proof.from_bytes(b"\x02\x02" + (1 << 30).to_bytes(4, "little") + bytes(4))


def test_validation_errors_match_reference():
    with pytest.raises(ValueError, match="values are not equal"):           # validation.rs:21-26
        zk.prove_equality(1, 2)
    with pytest.raises(ValueError, match="set cannot be empty"):            # validation.rs:47-50
        zk.prove_membership(1, [])
    with pytest.raises(ValueError, match="value 9 is not in the provided set"):
        zk.prove_membership(9, [1, 2, 3])
    with pytest.raises(ValueError, match="set size 65 exceeds maximum allowed size 64"):
        zk.prove_membership(1, list(range(65)))
    with pytest.raises(OverflowError):
        zk.prove_equality(2**64, 2**64)


def test_batch_registry_semantics():
    bid = zk.create_proof_batch()
    assert bid != 0
    zk.batch_add_equality_proof(bid, 7, 7)
    zk.batch_add_membership_proof(bid, 2, [1, 2, 3])
    zk.batch_add_membership_proof(bid, 1, list(range(100)))    # no size check at add time (batch.rs:92-95)
    with pytest.raises(ValueError):
        zk.batch_add_equality_proof(bid, 1, 2)
    with pytest.raises(ValueError, match="Invalid batch ID"):
        zk.batch_add_equality_proof(bid ^ 1, 1, 1)
    st = zk.get_batch_status(bid)
    assert st["total_operations"] == 3 and st["equality_proofs"] == 1 and st["membership_proofs"] == 2
    assert st["range_proofs"] == 0
    zk.clear_batch(bid)
    zk.clear_batch(bid)                                        # unknown id is not an error (batch.rs:175-183)
    with pytest.raises(ValueError, match="Invalid batch ID"):
        zk.process_batch(bid)
    assert zk.process_batch(zk.create_proof_batch()) == []     # empty batch -> empty list, id consumed


def test_key_dir_rules(tmp_path):
    snark.reset()
    try:
        with pytest.raises(TypeError):                          # ConfigError -> TypeError
            zk.set_snark_key_dir("")
        assert zk.set_snark_key_dir(str(tmp_path)) is True
        assert zk.set_snark_key_dir(str(tmp_path)) is True      # same value accepted (snark.rs:158-167)
        with pytest.raises(zk.ConfigError):
            zk.set_snark_key_dir(str(tmp_path / "other"))
        assert not zk.is_snark_setup_initialized()
    finally:
        snark.reset()


def test_zkp_backend_trait_input_length():
    assert zk.SnarkBackend.prove(b"\x00" * 47) == b""           # snark.rs:588-591
    assert zk.SnarkBackend.prove_equality_zk(1, 2, bytes(32)) == b""      # a != b (snark.rs:344)
    bad = (snark.R_MOD).to_bytes(32, "little")
    assert zk.SnarkBackend.prove_equality_zk(1, 1, bad) == b""  # non-canonical commitment (snark.rs:348-351)
    assert zk.SnarkBackend.prove_membership_zk(1, [], bytes(32)) == b""   # snark.rs:406
    assert zk.SnarkBackend.prove_membership_zk(1, list(range(65)), bytes(32)) == b""
    assert zk.SnarkBackend.prove_membership_zk(9, [1, 2], bytes(32)) == b""   # snark.rs:415-418


def test_builtin_witness_equals_oracle(co):
    c = co.Circuit("equality")
    for a in (0, 5, 2**64 - 1):
        assert np.array_equal(engine.builtin_witness(engine.EQUALITY, 110, a, a), c.assign(a, a))
    wrong = co.mimc_hash(99)
    assert np.array_equal(engine.builtin_witness(engine.EQUALITY, 110, 7, 7, commitment=wrong),
                          c.assign(7, 7, commitment=wrong))
    chain = co.Circuit("equality", 300)                       # the synthetic MiMC-chain circuit (config 4 shape)
    assert np.array_equal(engine.builtin_witness(engine.EQUALITY, 300, 12345, 12345), chain.assign(12345, 12345,
                          commitment=engine.builtin_witness(engine.EQUALITY, 300, 12345, 12345)[1].tobytes()))
    m = co.Circuit("membership")
    for v, s in ((2, [1, 2, 3]), (9, list(range(64))), (2**64 - 1, [2**64 - 1])):
        assert np.array_equal(engine.builtin_witness(engine.MEMBERSHIP, 64, v, set_=s), m.assign(v, set_=s))
    with pytest.raises(zk.EngineError):
        engine.builtin_witness(engine.MEMBERSHIP, 64, 7, set_=[1, 2, 3])


def test_window_count_covers_every_scalar():
    # engine.cu window_count(): W = ceil(255 / c) windows of the offset recoding (dev_util.cuh recode_offset) need
    # s + K < 2^(c*W) for every canonical s, with K = sum_w 2^(c*w + c-1); holds for BN254's r, and W-1 would not do.
    r = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    for c in range(8, 18):
        W = (255 + c - 1) // c
        K = sum(1 << (c * w + c - 1) for w in range(W))
        assert (r - 1) + K < 1 << (c * W), c
        digits = [(((r - 1) + K) >> (c * w)) % (1 << c) - (1 << (c - 1)) for w in range(W)]
        assert sum(d << (c * w) for w, d in enumerate(digits)) == r - 1
        assert all(-(1 << (c - 1)) <= d < (1 << (c - 1)) for d in digits)
        K1 = sum(1 << (c * w + c - 1) for w in range(W - 1))
        assert not (r - 1) + K1 < 1 << (c * (W - 1)), c
