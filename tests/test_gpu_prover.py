"""GPU parity tests of the batched Groth16 prover, all through the C ABI (include/lzkp_b200.h).
Bar: bit-exact — witness-map outputs, 256-byte proofs — against the CPU oracle on the same
seeded inputs, plus the reference's own round-trip / negative tests (snark.rs:634-641,
tests/integration.rs:19-44,86-90, examples/demo.rs:66-105) with the oracle's pairing verifier
standing in for ark-groth16's."""
import hashlib
import os

import numpy as np
import pytest

import libzkp_b200 as zk
from libzkp_b200 import engine, snark

pytestmark = pytest.mark.gpu

WINDOW_BITS = int(os.environ.get("LZKP_TEST_WINDOW_BITS", "12"))   # small tables: quick to build


@pytest.fixture(scope="module")
def eq_pk(eq_keys):
    pk = engine.ProvingKey(eq_keys.pk_bytes, validate=True, window_bits=WINDOW_BITS)
    pk.circuit_builtin(engine.EQUALITY, 110)
    yield pk
    pk.close()


@pytest.fixture(scope="module")
def mb_pk(mb_keys):
    pk = engine.ProvingKey(mb_keys.pk_bytes, validate=False, window_bits=WINDOW_BITS)
    pk.circuit_builtin(engine.MEMBERSHIP, 64)
    yield pk
    pk.close()


def test_pk_info(eq_pk, mb_pk):
    assert (eq_pk.n_vars, eq_pk.n_inst, eq_pk.n_wit, eq_pk.n) == (334, 2, 332, 512)
    assert (mb_pk.n_vars, mb_pk.n_inst, mb_pk.n_wit, mb_pk.n) == (653, 130, 523, 1024)
    assert eq_pk.window_bits == WINDOW_BITS


def test_witness_map_bit_exact(eq_pk, mb_pk, eq_keys, mb_keys, golden):
    zs = np.stack([eq_keys.circuit.assign(a, a) for a in (5, 42, 0, 2**64 - 1, 1234567)])
    h = eq_pk.witness_map(zs)
    for i in range(len(zs)):
        assert np.array_equal(h[i], eq_keys.circuit.witness_map(zs[i]))
    for i, case in enumerate(golden["equality"]["proofs"][:4]):
        assert hashlib.sha256(h[i].tobytes()).hexdigest() == case["h_sha256"]
    zm = np.stack([mb_keys.circuit.assign(v, set_=s) for v, s in ((2, [1, 2, 3]), (9, list(range(64))))])
    hm = mb_pk.witness_map(zm)
    for i in range(len(zm)):
        assert np.array_equal(hm[i], mb_keys.circuit.witness_map(zm[i]))


def test_witness_map_unsatisfied_assignment(eq_pk, eq_keys):
    # the prover does not check satisfiability (arkworks only does in debug builds): any z maps to SOME h
    z = eq_keys.circuit.assign(5, 5).copy()
    z[10, 0] ^= 1
    assert np.array_equal(eq_pk.witness_map(z[None])[0], eq_keys.circuit.witness_map(z))


def test_golden_proofs_bit_exact(eq_pk, mb_pk, eq_keys, mb_keys, co, golden):
    cases = golden["equality"]["proofs"]
    z = np.stack([eq_keys.circuit.assign(c["a"], c["a"]) for c in cases])
    r = co.fr_array([int(c["r"]) for c in cases])
    s = co.fr_array([int(c["s"]) for c in cases])
    proofs, status = eq_pk.prove_batch(z, r, s)              # includes the r = 0 / s = 0 edge cases
    assert not status.any()
    for i, c in enumerate(cases):
        assert proofs[i].tobytes().hex() == c["proof"], f"equality case {i}"
    cases = golden["membership"]["proofs"]
    z = np.stack([mb_keys.circuit.assign(c["value"], set_=c["set"]) for c in cases])
    proofs, status = mb_pk.prove_batch(z, co.fr_array([int(c["r"]) for c in cases]),
                                       co.fr_array([int(c["s"]) for c in cases]))
    assert not status.any()
    for i, c in enumerate(cases):
        assert proofs[i].tobytes().hex() == c["proof"], f"membership case {i}"


@pytest.mark.parametrize("n", [1, 3, 33, 130, 300, 512, 513])      # 512 | 513: latency shape | throughput shape
def test_equality_batch_device_witness_vs_oracle(eq_pk, eq_keys, co, po, frs, n):
    rng = po.SplitMix64(3)
    a = np.array([rng.next_u64() for _ in range(n)], np.uint64)
    b = a.copy()
    if n > 2:
        b[1] ^= 1                                           # a != b -> status 1, blank proof (snark.rs:344)
    r, s = frs(4, n), frs(40, n)
    proofs, cms, status = eq_pk.prove_equality_batch(a, b, r, s)
    want, wstat = co.prove_batch(eq_keys.circuit, eq_keys.opk, a, b, None, None, r, s)
    assert np.array_equal(status != 0, wstat != 0)
    assert np.array_equal(proofs, want)
    for i in range(min(n, 5)):
        if status[i] == 0:
            assert cms[i].tobytes() == co.mimc_hash(int(a[i]))
    assert n <= 2 or (status[1] == 1 and not proofs[1].any())


def test_equality_batch_with_supplied_commitments(eq_pk, eq_keys, co, frs):
    # a wrong commitment still yields a proof (of a false statement); bytes must match the oracle's
    a = np.array([11, 12], np.uint64)
    cm = np.stack([np.frombuffer(co.mimc_hash(11), np.uint8), np.frombuffer(co.mimc_hash(99), np.uint8)])
    r, s = frs(5, 2), frs(6, 2)
    proofs, _, status = eq_pk.prove_equality_batch(a, a, r, s, commitments=cm)
    assert not status.any()
    for i in range(2):
        z = eq_keys.circuit.assign(int(a[i]), int(a[i]), commitment=cm[i].tobytes())
        assert proofs[i].tobytes() == co.prove(eq_keys.circuit, eq_keys.opk, z, co.fr_list(r[i])[0], co.fr_list(s[i])[0])


def test_membership_batch_device_witness_vs_oracle(mb_pk, mb_keys, co, po, frs):
    rng = po.SplitMix64(5)
    n = 70
    sets = np.zeros((n, 64), np.uint64)
    lens = np.zeros(n, np.uint32)
    vals = np.zeros(n, np.uint64)
    for i in range(n):
        L = 1 + i % 64
        lens[i] = L
        sets[i, :L] = [rng.next_u64() for _ in range(L)]
        vals[i] = sets[i, i % L]
    vals[3] = 12345                                         # not in set -> status 2 (snark.rs:415-418)
    lens[4] = 0                                             # empty set (snark.rs:406)
    r, s = frs(7, n), frs(8, n)
    proofs, cms, status = mb_pk.prove_membership_batch(vals, sets, lens, r, s)
    want, wstat = co.prove_batch(mb_keys.circuit, mb_keys.opk, vals, None, sets, lens, r, s)
    assert status[3] == 2 and status[4] == 2
    assert np.array_equal(status != 0, wstat != 0)
    assert np.array_equal(proofs, want)


def test_noncanonical_scalar_is_rejected(eq_pk, eq_keys, frs):
    z = np.stack([eq_keys.circuit.assign(5, 5), eq_keys.circuit.assign(6, 6)])
    z[1, 7] = 0xFF                                          # >= r
    proofs, status = eq_pk.prove_batch(z, frs(1, 2), frs(2, 2))
    assert status[0] == 0 and status[1] != 0 and not proofs[1].any() and proofs[0].any()


# ---------------------------------------------------------------- the reference's own tests, mirrored
@pytest.fixture(scope="module")
def api(eq_keys, mb_keys):
    keys = {"equality_mimc": (eq_keys.pk_bytes, eq_keys.vk_bytes), "membership_mimc": (mb_keys.pk_bytes, mb_keys.vk_bytes)}
    snark.reset()
    snark.configure(window_bits=WINDOW_BITS, generator=lambda prefix: keys[prefix])
    yield zk
    snark.reset()
    snark.configure()


def verify_equality(po, keys, proof_bytes, val1, val2):
    """verify_equality (equality_proof.rs:50-57) with the oracle's pairing verifier."""
    if val1 != val2:
        return False
    p = zk.Proof.from_bytes(proof_bytes)
    cm = zk.snark_commit_value(val1)
    if p.scheme != 2 or p.commitment != cm:
        return False
    return po.verify(po.vk_from_bytes(keys.vk_bytes), po.equality_public_inputs(int.from_bytes(cm, "little")),
                     po.proof_from_bytes(p.proof))


def verify_membership(po, keys, proof_bytes, set_):
    """verify_membership (set_membership.rs:40-71)."""
    p = zk.Proof.from_bytes(proof_bytes)
    if p.scheme != 4:
        return False
    n = int.from_bytes(p.proof[:4], "little")
    emb = [int.from_bytes(p.proof[4 + 8 * i:12 + 8 * i], "little") for i in range(n)]
    if sorted(emb) != sorted(set_):
        return False
    return po.verify(po.vk_from_bytes(keys.vk_bytes),
                     po.membership_public_inputs(int.from_bytes(p.commitment, "little"), emb),
                     po.proof_from_bytes(p.proof[4 + 8 * n:]))


def test_integration_equality(api, po, eq_keys):
    proof = api.prove_equality(3, 3)                         # tests/integration.rs:19-23
    assert len(proof) == 298
    assert verify_equality(po, eq_keys, proof, 3, 3)
    proof = api.prove_equality(42, 42)                       # tests/integration.rs:25-32
    assert zk.Proof.from_bytes(proof).commitment == api.snark_commit_value(42)
    assert verify_equality(po, eq_keys, proof, 42, 42)
    assert not verify_equality(po, eq_keys, proof, 43, 43)   # tests/integration.rs:86-90
    assert api.prove_equality(42, 42) != proof               # fresh r, s every call (OsRng)
    assert api.is_snark_setup_initialized()


def test_snark_backend_roundtrip(api, po, eq_keys):
    # snark.rs:634-641
    cm = api.snark_commit_value(42)
    proof = api.SnarkBackend.prove_equality_zk(42, 42, cm)
    assert len(proof) == 256
    vk = po.vk_from_bytes(eq_keys.vk_bytes)
    assert po.verify(vk, po.equality_public_inputs(int.from_bytes(cm, "little")), po.proof_from_bytes(proof))
    wrong = api.snark_commit_value(99)
    assert not po.verify(vk, po.equality_public_inputs(int.from_bytes(wrong, "little")), po.proof_from_bytes(proof))
    data = (42).to_bytes(8, "little") * 2 + cm               # ZkpBackend::prove, snark.rs:587-606
    assert len(api.SnarkBackend.prove(data)) == 256


def test_integration_membership(api, po, mb_keys):
    proof = api.prove_membership(2, [1, 2, 3])               # tests/integration.rs:40-44
    assert len(proof) == 10 + 4 + 24 + 256 + 32
    assert verify_membership(po, mb_keys, proof, [1, 2, 3])
    assert verify_membership(po, mb_keys, proof, [3, 2, 1])  # sets compare as sorted multisets
    assert not verify_membership(po, mb_keys, proof, [1, 2, 4])


def test_process_batch_order_and_fail_fast(api, po, eq_keys, mb_keys):
    bid = api.create_proof_batch()                           # examples/demo.rs:66-105 (SNARK ops only)
    ops = [("e", 100, 100), ("m", 25, [10, 20, 25, 30, 40]), ("e", 7, 7), ("m", 1, [1]), ("e", 2**64 - 1, 2**64 - 1)]
    for o in ops:
        if o[0] == "e":
            api.batch_add_equality_proof(bid, o[1], o[2])
        else:
            api.batch_add_membership_proof(bid, o[1], o[2])
    out = api.process_batch(bid)
    assert len(out) == len(ops)
    for o, p in zip(ops, out):                               # insertion order kept
        if o[0] == "e":
            assert verify_equality(po, eq_keys, p, o[1], o[2])
        else:
            assert verify_membership(po, mb_keys, p, o[2])
    with pytest.raises(ValueError):                          # id consumed (batch.rs:111-118)
        api.process_batch(bid)
    bid = api.create_proof_batch()
    api.batch_add_equality_proof(bid, 1, 1)
    api.batch_add_membership_proof(bid, 0, list(range(100)))  # oversized set only fails at process time
    with pytest.raises(ValueError, match="exceeds maximum"):
        api.process_batch(bid)


def test_seeded_rng_gives_reproducible_bytes(api, po, eq_keys, co):
    rng1, rng2 = po.SplitMix64(2), po.SplitMix64(2)
    p1 = api.SnarkBackend.prove_equality_zk(5, 5, api.snark_commit_value(5), rng=rng1)
    r, s = rng2.next_fr(), rng2.next_fr()
    z = eq_keys.circuit.assign(5, 5)
    assert p1 == co.prove(eq_keys.circuit, eq_keys.opk, z, r, s)


def test_window_sizes_agree(eq_keys, co, frs, po):
    a = np.arange(1, 9, dtype=np.uint64)
    r, s = frs(31, 8), frs(32, 8)
    top = np.frombuffer((po.R_MOD - 1).to_bytes(32, "little"), np.uint8)
    r[0], s[1], r[2], s[2] = top, top, top, top      # largest canonical scalars: worst case of the offset recoding
    want, _ = co.prove_batch(eq_keys.circuit, eq_keys.opk, a, a, None, None, r, s)
    for c in (8, 11, 15, 16, 17):                    # 17: 4-byte digits, 15 windows
        pk = engine.ProvingKey(eq_keys.pk_bytes, window_bits=c)
        pk.circuit_builtin(engine.EQUALITY, 110)
        assert pk.window_bits == c
        proofs, _, status = pk.prove_equality_batch(a, a, r, s)
        pk.close()
        assert not status.any() and np.array_equal(proofs, want), f"c={c}"


def test_device_setup_bit_exact(eq_keys, mb_keys, trapdoor, golden):
    # lzkp_setup_builtin (device fixed-base multiplications) == the oracle's generate_parameters
    pk, vk = engine.setup_builtin(engine.EQUALITY, 110, trapdoor)
    assert hashlib.sha256(pk).hexdigest() == golden["equality"]["pk_sha256"]
    assert pk == eq_keys.pk_bytes and vk == eq_keys.vk_bytes
    pk, vk = engine.setup_builtin(engine.MEMBERSHIP, 64, trapdoor)
    assert pk == mb_keys.pk_bytes and vk == mb_keys.vk_bytes


def test_device_setup_generic_csr_and_default_keygen(eq_keys, trapdoor, po):
    (m, n_inst, n_wit), mats = engine.builtin_circuit_csr(engine.EQUALITY, 110)
    pk, vk = engine.setup(m, n_inst, n_wit, mats, trapdoor)
    assert pk == eq_keys.pk_bytes and vk == eq_keys.vk_bytes
    # default key generation (OsRng toxic waste) yields a key whose proofs verify
    snark.reset()
    snark.configure(window_bits=WINDOW_BITS)
    try:
        proof = zk.prove_equality(9, 9)
        setup = zk.SnarkBackend.get_universal_setup()
        p = zk.Proof.from_bytes(proof)
        assert po.verify(po.vk_from_bytes(setup.vk_bytes),
                         po.equality_public_inputs(int.from_bytes(p.commitment, "little")), po.proof_from_bytes(p.proof))
    finally:
        snark.reset()
        snark.configure()


def test_empty_and_single_batches(eq_pk, mb_pk, frs):
    e = np.zeros((0, 32), np.uint8)
    proofs, cms, status = eq_pk.prove_equality_batch(np.zeros(0, np.uint64), np.zeros(0, np.uint64), e, e)
    assert proofs.shape == (0, 256) and status.shape == (0,)
    proofs, status = eq_pk.prove_batch(np.zeros((0, eq_pk.n_vars * 32), np.uint8), e, e)
    assert proofs.shape == (0, 256)
    proofs, _, status = mb_pk.prove_membership_batch(np.zeros(0, np.uint64), np.zeros((0, 64), np.uint64),
                                                     np.zeros(0, np.uint32), e, e)
    assert proofs.shape == (0, 256)


def test_multi_chunk_batch_matches_oracle(eq_keys, co, po, frs):
    # more proofs than one device pass: chunks alternate between the two workspaces / streams; order is kept
    pk = engine.ProvingKey(eq_keys.pk_bytes, window_bits=WINDOW_BITS, max_chunk=64)
    pk.circuit_builtin(engine.EQUALITY, 110)
    n = 64 * 5 + 17
    rng = po.SplitMix64(33)
    a = np.array([rng.next_u64() for _ in range(n)], np.uint64)
    b = a.copy()
    b[100] ^= 1
    r, s = frs(34, n), frs(35, n)
    proofs, cms, status = pk.prove_equality_batch(a, b, r, s)
    want, wstat = co.prove_batch(eq_keys.circuit, eq_keys.opk, a, b, None, None, r, s)
    assert np.array_equal(status != 0, wstat != 0) and status[100] == 1
    assert np.array_equal(proofs, want)
    assert cms[n - 1].tobytes() == co.mimc_hash(int(a[n - 1]))
    pk.close()


def test_device_envelopes_equal_host_framing(eq_pk, mb_pk, frs, po):
    # Proof::to_bytes framing written on the device == framing built on the host from the bare proofs
    n = 40
    rng = po.SplitMix64(8)
    a = np.array([rng.next_u64() for _ in range(n)], np.uint64)
    b = a.copy()
    b[5] ^= 1
    r, s = frs(51, n), frs(52, n)
    env, lens, status = eq_pk.prove_equality_enveloped(a, b, r, s)
    proofs, cms, st2 = eq_pk.prove_equality_batch(a, b, r, s)
    assert np.array_equal(status, st2) and status[5] == 1 and lens[5] == 0
    for i in range(n):
        if status[i] == 0:
            assert lens[i] == 298
            assert env[i].tobytes() == zk.Proof(2, proofs[i].tobytes(), cms[i].tobytes()).to_bytes()
    m = 20
    sets = np.zeros((m, 64), np.uint64)
    lens_in = np.zeros(m, np.uint32)
    vals = np.zeros(m, np.uint64)
    for i in range(m):
        L = 1 + (i * 7) % 64
        lens_in[i] = L
        sets[i, :L] = [rng.next_u64() for _ in range(L)]
        vals[i] = sets[i, i % L]
    vals[3] = 424242                                          # not in its set
    r, s = frs(53, m), frs(54, m)
    env, lens, status = mb_pk.prove_membership_enveloped(vals, sets, lens_in, r, s)
    proofs, cms, st2 = mb_pk.prove_membership_batch(vals, sets, lens_in, r, s)
    assert np.array_equal(status, st2) and status[3] == 2 and lens[3] == 0
    import struct
    for i in range(m):
        if status[i] == 0:
            L = int(lens_in[i])
            payload = struct.pack("<I", L) + sets[i, :L].astype("<u8").tobytes() + proofs[i].tobytes()
            want = zk.Proof(4, payload, cms[i].tobytes()).to_bytes()
            assert lens[i] == len(want) and env[i, :lens[i]].tobytes() == want


def test_explicit_csr_circuit_and_error_paths(eq_keys, mb_keys, co, golden):
    # lzkp_circuit_load with caller-supplied CSR matrices (what a Rust shim would pass from cs.to_matrices())
    (m, n_inst, n_wit), mats = engine.builtin_circuit_csr(engine.EQUALITY, 110)
    pk = engine.ProvingKey(eq_keys.pk_bytes, window_bits=WINDOW_BITS)
    with pytest.raises(zk.EngineError):                       # proving before a circuit is bound
        pk.prove_batch(np.zeros((1, pk.n_vars * 32), np.uint8), np.zeros((1, 32), np.uint8), np.zeros((1, 32), np.uint8))
    with pytest.raises(zk.EngineError):                       # equality batch needs the builtin equality circuit
        pk.prove_equality_batch(np.zeros(1, np.uint64), np.zeros(1, np.uint64), np.zeros((1, 32), np.uint8), np.zeros((1, 32), np.uint8))
    with pytest.raises(zk.EngineError):                       # shape mismatch with the key
        pk.circuit_load(m, n_inst + 1, n_wit - 1, mats)
    pk.circuit_load(m, n_inst, n_wit, mats)
    case = golden["equality"]["proofs"][0]
    z = eq_keys.circuit.assign(case["a"], case["a"])[None]
    proofs, status = pk.prove_batch(z, co.fr_array([int(case["r"])]), co.fr_array([int(case["s"])]))
    assert not status.any() and proofs[0].tobytes().hex() == case["proof"]
    pk.close()
    # a membership key cannot take the equality circuit
    pkm = engine.ProvingKey(mb_keys.pk_bytes, window_bits=8)
    with pytest.raises(zk.EngineError):
        pkm.circuit_builtin(engine.EQUALITY, 110)
    pkm.close()


def test_pk_validation_rejects_corrupted_keys(eq_keys):
    good = bytearray(eq_keys.pk_bytes)
    with pytest.raises(zk.EngineError):
        engine.ProvingKey(bytes(good[:-7]), window_bits=8)                 # truncated
    bad = bytearray(good)
    off = 64 + 3 * 128 + 8 + 2 * 64 + 64 + 64 + 8 + 5 * 64                  # x of a_query[5]
    bad[off] ^= 1
    with pytest.raises(zk.EngineError, match="off-curve"):
        engine.ProvingKey(bytes(bad), validate=True, window_bits=8)
    bad = bytearray(good)
    bad[off + 31] |= 0x3F                                                    # coordinate >= q
    with pytest.raises(zk.EngineError):
        engine.ProvingKey(bytes(bad), window_bits=8)
