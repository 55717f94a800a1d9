"""GPU parity tests of the large-domain prover path (single proofs whose witness
map runs on the tiled NTT and whose five MSMs run as Pippenger over resident window-shifted bases)."""
import os

import numpy as np
import pytest

from libzkp_b200 import engine

pytestmark = pytest.mark.gpu


@pytest.fixture()
def force_large():
    os.environ["LZKP_FORCE_LARGE"] = "1"
    yield
    os.environ.pop("LZKP_FORCE_LARGE", None)


def test_large_path_on_reference_circuits_matches_golden(force_large, eq_keys, mb_keys, co, golden):
    # the reference's own two circuits pushed through the large-domain code path: same golden bytes
    pk = engine.ProvingKey(eq_keys.pk_bytes, validate=True)
    pk.circuit_builtin(engine.EQUALITY, 110)
    assert pk.max_chunk == 1
    cases = golden["equality"]["proofs"]
    z = np.stack([eq_keys.circuit.assign(c["a"], c["a"]) for c in cases])
    h = pk.witness_map(z[:2])
    assert np.array_equal(h[0], eq_keys.circuit.witness_map(z[0]))
    proofs, status = pk.prove_batch(z, co.fr_array([int(c["r"]) for c in cases]), co.fr_array([int(c["s"]) for c in cases]))
    assert not status.any()
    for i, c in enumerate(cases):
        assert proofs[i].tobytes().hex() == c["proof"], f"equality case {i}"
    pk.close()
    pk = engine.ProvingKey(mb_keys.pk_bytes)
    pk.circuit_builtin(engine.MEMBERSHIP, 64)
    cases = golden["membership"]["proofs"]
    z = np.stack([mb_keys.circuit.assign(c["value"], set_=c["set"]) for c in cases])
    proofs, status = pk.prove_batch(z, co.fr_array([int(c["r"]) for c in cases]), co.fr_array([int(c["s"]) for c in cases]))
    assert not status.any()
    for i, c in enumerate(cases):
        assert proofs[i].tobytes().hex() == c["proof"], f"membership case {i}"
    pk.close()


@pytest.mark.parametrize("rounds", [2730, 21845])
def test_mimc_chain_circuit_matches_oracle(co, po, trapdoor, frs, rounds):
    # synthetic MiMC-chain circuit (config 4's shape at 2^14 and 2^17): setup, witness map and proofs bit-exact
    circ = co.Circuit("equality", rounds)
    assert circ.m == 3 * rounds + 2
    pk_bytes, vk_bytes = engine.setup_builtin(engine.EQUALITY, rounds, trapdoor)
    opk_bytes, ovk_bytes = circ.setup(trapdoor)
    assert pk_bytes == opk_bytes and vk_bytes == ovk_bytes
    pk = engine.ProvingKey(pk_bytes)
    pk.circuit_builtin(engine.EQUALITY, rounds)
    # 2^14: window tables still fit -> batched path with the tiled witness map; 2^17: one proof per pass
    assert pk.n == circ.n and (pk.max_chunk == 1) == (circ.n > 1 << 16)
    z = np.stack([engine.builtin_witness(engine.EQUALITY, rounds, a, a) for a in (5, 2**64 - 1)])
    assert np.array_equal(z[0], circ.assign(5, 5, commitment=z[0, 1].tobytes()))
    assert np.array_equal(pk.witness_map(z)[1], circ.witness_map(z[1]))
    r, s = frs(4, 2), frs(40, 2)
    proofs, status = pk.prove_batch(z, r, s)
    assert not status.any()
    opk = co.ProvingKey(opk_bytes)
    for i in range(2):
        assert proofs[i].tobytes() == co.prove(circ, opk, z[i], co.fr_list(r[i])[0], co.fr_list(s[i])[0])
    # non-canonical scalar -> status, blank proof
    z[1, 9] = 0xFF
    proofs, status = pk.prove_batch(z, r, s)
    assert status[0] == 0 and status[1] != 0 and not proofs[1].any()
    pk.close()


def test_config4_2_20_constraints_proof_verifies(po, trapdoor, frs, co):
    # BASELINE.json configs[3]: 2^20-constraint synthetic circuit, single proof.  Full-size check through
    # a size-independent property: the pairing verifier accepts, and rejects another public input.
    rounds = 349524
    (m, n_inst, n_wit) = engine.builtin_circuit_shape(engine.EQUALITY, rounds)
    assert m == 1048574 and n_inst == 2
    pk_bytes, vk_bytes = engine.setup_builtin(engine.EQUALITY, rounds, trapdoor)
    assert len(pk_bytes) > 380 * 2**20
    pk = engine.ProvingKey(pk_bytes)
    pk.circuit_builtin(engine.EQUALITY, rounds)
    assert pk.n == 1 << 20
    z = engine.builtin_witness(engine.EQUALITY, rounds, 6, 6)[None]
    r, s = frs(4, 1), frs(40, 1)
    proofs, status = pk.prove_batch(z, r, s)
    assert not status.any()
    vk = po.vk_from_bytes(vk_bytes)
    cm = int.from_bytes(z[0, 1].tobytes(), "little")
    proof = po.proof_from_bytes(proofs[0].tobytes())
    assert po.verify(vk, [cm], proof)
    assert not po.verify(vk, [(cm + 1) % po.R_MOD], proof)
    # determinism + h has degree <= n - 2
    assert np.array_equal(pk.prove_batch(z, r, s)[0], proofs)
    h = pk.witness_map(z)
    assert not h[0, -1].any() and h[0, -2].any()
    pk.close()


@pytest.mark.parametrize("shards", [2, 3, 8])
def test_sharded_single_proof_partials_combine(force_large, co, trapdoor, frs, shards):
    # §8e single-proof split, emulated on one GPU: every "rank" is a pk holding only its point ranges;
    # partial sums of all shards are combined by shard 0.  Bit-exact against the oracle's proof.
    import torch
    rounds = 2730
    circ = co.Circuit("equality", rounds)
    pk_bytes, _ = circ.setup(trapdoor)
    z = engine.builtin_witness(engine.EQUALITY, rounds, 77, 77)
    r, s = frs(4, 1), frs(40, 1)
    want = co.prove(circ, co.ProvingKey(pk_bytes), z, co.fr_list(r)[0], co.fr_list(s)[0])
    dev = torch.device("cuda", 0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d_z, d_r, d_s = t(z), t(r[0]), t(s[0])
    pks = [engine.ProvingKey(pk_bytes, shard_index=i, shard_count=shards) for i in range(shards)]
    pks[0].circuit_builtin(engine.EQUALITY, rounds)
    d_h = torch.zeros((circ.n, 32), dtype=torch.uint8, device=dev)
    pks[0].witness_map_device(d_z.data_ptr(), d_h.data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(d_h.cpu().numpy(), circ.witness_map(z))
    partials = torch.zeros((shards, 768), dtype=torch.uint8, device=dev)
    status = torch.zeros(shards, dtype=torch.int32, device=dev)
    for i, pk in enumerate(pks):
        if i % 2:                                            # two-phase form: z-only MSMs first, H MSM once h is there
            pk.prove_partial_device(d_z.data_ptr(), d_r.data_ptr(), d_s.data_ptr(), 0, 0, status[i:].data_ptr(), phase=1)
            pk.prove_partial_device(0, d_r.data_ptr(), d_s.data_ptr(), d_h.data_ptr(), partials[i].data_ptr(),
                                    status[i:].data_ptr(), phase=2)
        else:
            pk.prove_partial_device(d_z.data_ptr(), d_r.data_ptr(), d_s.data_ptr(), d_h.data_ptr(),
                                    partials[i].data_ptr(), status[i:].data_ptr())
    proof = torch.zeros(256, dtype=torch.uint8, device=dev)
    pks[0].prove_combine_device(partials.data_ptr(), shards, d_r.data_ptr(), d_s.data_ptr(), proof.data_ptr())
    torch.cuda.synchronize()
    assert not status.any()
    assert proof.cpu().numpy().tobytes() == want
    for pk in pks:
        pk.close()


def test_sharded_prover_world_1(force_large, co, trapdoor, frs):
    import torch
    from libzkp_b200.multi import ShardedProver
    rounds = 2730
    circ = co.Circuit("equality", rounds)
    pk_bytes, _ = circ.setup(trapdoor)
    z = engine.builtin_witness(engine.EQUALITY, rounds, 5, 5)
    r, s = frs(4, 1), frs(40, 1)
    sp = ShardedProver(pk_bytes, engine.EQUALITY, rounds, 0, 1, torch.device("cuda", 0))
    got = sp.prove(z, r[0].tobytes(), s[0].tobytes())
    assert got == co.prove(circ, co.ProvingKey(pk_bytes), z, co.fr_list(r)[0], co.fr_list(s)[0])
    sp.close()


def test_mid_size_circuit_forced_large_path(force_large, co, trapdoor, frs):
    # the 2^14 chain circuit again, pushed onto the one-proof-per-pass path (two-pass NTT, Pippenger at n = 8k)
    rounds = 2730
    circ = co.Circuit("equality", rounds)
    pk_bytes, _ = circ.setup(trapdoor)
    pk = engine.ProvingKey(pk_bytes)
    pk.circuit_builtin(engine.EQUALITY, rounds)
    assert pk.max_chunk == 1
    z = engine.builtin_witness(engine.EQUALITY, rounds, 12, 12)[None]
    r, s = frs(4, 1), frs(40, 1)
    proofs, status = pk.prove_batch(z, r, s)
    assert not status.any()
    assert proofs[0].tobytes() == co.prove(circ, co.ProvingKey(pk_bytes), z[0], co.fr_list(r)[0], co.fr_list(s)[0])
    pk.close()


def test_membership_1024_slots_matches_oracle(co, trapdoor, frs, po):
    # BASELINE.json configs[2]: membership with a 1024-element public set (m = 5453, n = 8192).  The reference
    # itself caps sets at 64 (snark.rs:503); the same circuit re-parameterised runs on the large-domain path.
    S = 1024
    circ = co.Circuit("membership", S)
    assert (circ.m, circ.n_inst, circ.n_wit, circ.n) == (5453, 2050, 3403, 8192)
    pk_bytes, vk_bytes = engine.setup_builtin(engine.MEMBERSHIP, S, trapdoor)
    opk_bytes, ovk_bytes = circ.setup(trapdoor)
    assert pk_bytes == opk_bytes and vk_bytes == ovk_bytes
    pk = engine.ProvingKey(pk_bytes)
    pk.circuit_builtin(engine.MEMBERSHIP, S)
    assert pk.max_chunk > 1 and pk.window_bits >= 8           # batched path: window tables + tiled witness map
    rng = po.SplitMix64(5)
    n = 3
    sets = np.array([[rng.next_u64() for _ in range(S)] for _ in range(n)], np.uint64)
    lens = np.array([S, 700, 1], np.uint32)
    vals = np.array([sets[0, 1023], sets[1, 350], sets[2, 0]], np.uint64)
    r, s = frs(7, n), frs(8, n)
    proofs, cms, status = pk.prove_membership_batch(vals, sets, lens, r, s)
    assert not status.any()
    opk = co.ProvingKey(opk_bytes)
    for i in range(n):
        z = circ.assign(int(vals[i]), set_=[int(v) for v in sets[i, :lens[i]]])
        assert proofs[i].tobytes() == co.prove(circ, opk, z, co.fr_list(r[i])[0], co.fr_list(s[i])[0]), i
        assert cms[i].tobytes() == co.mimc_hash(int(vals[i]))
    pk.close()
