"""GPU tests of the batched verifier (lzkp_vk_load / lzkp_verify_batch): decisions equal the oracle's independent
pairing verifier on valid, forged and malformed proofs, and the reference's own verify tests (snark.rs:634-641,
tests/integration.rs:19-44,86-90, examples/demo.rs:66-105) pass through the Python mirror."""
import numpy as np
import pytest

import libzkp_b200 as zk
from libzkp_b200 import engine, snark

pytestmark = pytest.mark.gpu
WINDOW_BITS = 10


def test_verify_golden_and_forgeries(eq_keys, mb_keys, golden, po, co):
    vk = engine.VerifyingKey(eq_keys.vk_bytes)
    assert vk.n_pub == 1
    cases = [c for c in golden["equality"]["proofs"] if "z_sha256" in c]
    proofs = np.stack([np.frombuffer(bytes.fromhex(c["proof"]), np.uint8) for c in cases])
    cms = np.stack([np.frombuffer(co.mimc_hash(c["a"]), np.uint8) for c in cases])
    assert vk.verify_batch(proofs, cms).all()
    wrong = np.roll(cms, 1, axis=0)
    assert not vk.verify_batch(proofs, wrong).any()                       # wrong public input (snark.rs:639-640)
    ovk = po.vk_from_bytes(eq_keys.vk_bytes)
    forged = []
    for k in range(12):
        b = proofs[k % len(proofs)].copy()
        if k < 4:
            b[[5, 70, 200, 255][k]] ^= 1                                  # bit flips: off-curve / non-canonical / flag
        elif k < 8:                                                       # a valid but different curve point for A / C
            other = po.g1_to_bytes(po.G1.mul(po.G1_GEN, 1000 + k))
            off = 0 if k % 2 else 192
            b[off:off + 64] = np.frombuffer(other, np.uint8)
        elif k == 8:
            b[64:192] = np.frombuffer(po.g2_to_bytes(po.G2.mul(po.G2_GEN, 77)), np.uint8)
        elif k == 9:
            b[:64] = 0
            b[63] = 0x40                                                  # A = infinity
        elif k == 10:
            b[:] = 0
        else:
            b[31] |= 0x3F                                                 # x >= q
        forged.append(b)
    forged = np.stack(forged)
    fc = np.stack([cms[k % len(cms)] for k in range(12)])
    got = vk.verify_batch(forged, fc)
    for k in range(12):
        try:
            want = po.verify(ovk, po.equality_public_inputs(int.from_bytes(fc[k].tobytes(), "little")),
                             po.proof_from_bytes(forged[k].tobytes()))
        except Exception:
            want = False
        assert bool(got[k]) == bool(want) == False, k
    # B on the twist but outside the r-torsion subgroup (tools/pairing_proto.py): deserialization must reject it
    offsub = bytes.fromhex("02000000000000000000000000000000000000000000000000000000000000000100000000000000000000000000000000000000000000000000000000000000ce0141067f1657f01323d6d1264d1f38c5e0ce2da0ec9950b908934178721f10de5f8f80b5b618595db7a7f6137c7f775a006a5485ac3d962ab99b5979c176ab")
    assert po.G2.on_curve(po.g2_from_bytes(offsub, validate=False))
    b = proofs[0].copy()
    b[64:192] = np.frombuffer(offsub, np.uint8)
    assert not vk.verify_batch(b[None], cms[:1])[0]
    # non-canonical public input and wrong input count
    bad = cms.copy()
    bad[0] = 0xFF
    assert not vk.verify_batch(proofs[:1], bad[:1])[0]
    assert not vk.verify_batch(proofs[:2], np.zeros((2, 2, 32), np.uint8)).any()
    assert vk.verify_batch(np.zeros((0, 256), np.uint8), np.zeros((0, 1, 32), np.uint8)).shape == (0,)
    vk.close()
    # membership golden proofs
    vkm = engine.VerifyingKey(mb_keys.vk_bytes)
    assert vkm.n_pub == 129
    mc = golden["membership"]["proofs"]
    mp = np.stack([np.frombuffer(bytes.fromhex(c["proof"]), np.uint8) for c in mc])
    x = np.zeros((len(mc), 129, 32), np.uint8)
    for k, c in enumerate(mc):
        pub = po.membership_public_inputs(int.from_bytes(co.mimc_hash(c["value"]), "little"), c["set"])
        x[k] = co.fr_array(pub)
    assert vkm.verify_batch(mp, x).all()
    x[1, 5, 0] ^= 1
    assert list(vkm.verify_batch(mp, x)) == [True, False, True, True]
    vkm.close()


@pytest.fixture(scope="module")
def api(eq_keys, mb_keys):
    keys = {"equality_mimc": (eq_keys.pk_bytes, eq_keys.vk_bytes), "membership_mimc": (mb_keys.pk_bytes, mb_keys.vk_bytes)}
    snark.reset()
    snark.configure(window_bits=WINDOW_BITS, generator=lambda prefix: keys[prefix])
    yield zk
    snark.reset()
    snark.configure()


def test_reference_verify_tests_through_the_mirror(api, po, eq_keys):
    proof = api.prove_equality(3, 3)                                    # tests/integration.rs:19-23
    assert api.verify_equality(proof, 3, 3)
    proof = api.prove_equality(42, 42)                                  # tests/integration.rs:25-32
    assert api.verify_equality_with_commitment(proof, api.snark_commit_value(42))
    assert not api.verify_equality(proof, 43, 43)                       # tests/integration.rs:86-90
    assert not api.verify_equality(proof, 42, 43)
    assert not api.verify_equality(proof[:-1], 42, 42)
    tampered = bytearray(proof)
    tampered[100] ^= 1
    assert not api.verify_equality(bytes(tampered), 42, 42)
    cm = api.snark_commit_value(42)                                     # snark.rs:634-641
    raw = api.SnarkBackend.prove_equality_zk(42, 42, cm)
    assert api.SnarkBackend.verify_equality_zk(raw, cm)
    assert not api.SnarkBackend.verify_equality_zk(raw, api.snark_commit_value(99))
    assert api.SnarkBackend.verify(raw, cm)
    # agreement with the oracle's pairing verifier on the same bytes
    assert po.verify(po.vk_from_bytes(eq_keys.vk_bytes), [int.from_bytes(cm, "little")], po.proof_from_bytes(raw))
    mp = api.prove_membership(2, [1, 2, 3])                             # tests/integration.rs:40-44
    assert api.verify_membership(mp, [1, 2, 3]) and api.verify_membership(mp, [3, 1, 2])
    assert not api.verify_membership(mp, [1, 2, 4]) and not api.verify_membership(mp, [1, 2])


def test_batch_then_verify_proofs_parallel(api):
    bid = api.create_proof_batch()                                      # examples/demo.rs:66-105 (SNARK kinds)
    kinds = []
    for i in range(6):
        if i % 2 == 0:
            api.batch_add_equality_proof(bid, 100 + i, 100 + i)
            kinds.append("equality")
        else:
            api.batch_add_membership_proof(bid, 25, [10, 20, 25, 30, 40][: 3 + i % 3] if 25 in [10, 20, 25, 30, 40][: 3 + i % 3] else [25])
            kinds.append("membership")
    proofs = api.process_batch(bid)
    res = api.verify_proofs_parallel(list(zip(proofs, kinds)))
    assert res == [True] * 6
    broken = bytearray(proofs[2])
    broken[50] ^= 4
    res = api.verify_proofs_parallel([(bytes(broken), "equality"), (proofs[1], "membership"), (proofs[0], "membership")])
    assert res == [False, True, False]


def test_rlc_batched_verification_equals_independent_verification(eq_keys, mb_keys, co, po, frs, monkeypatch):
    # Large batches take the random-linear-combination path (one Miller loop per proof, one final exponentiation per
    # group of 64); forced here at a small size.  Decisions must be those of proof-by-proof verification: malformed
    # proofs are reported individually, a group holding a false proof falls back to k_verify4, clean groups pass.
    pk = engine.ProvingKey(eq_keys.pk_bytes, window_bits=WINDOW_BITS)
    pk.circuit_builtin(engine.EQUALITY, 110)
    n = 64 * 4 + 21                                                      # five groups, the last one partial
    rng = po.SplitMix64(77)
    a = np.array([rng.next_u64() for _ in range(n)], np.uint64)
    proofs, cms, st = pk.prove_equality_batch(a, a, frs(78, n), frs(79, n))
    pk.close()
    assert not st.any()
    vk = engine.VerifyingKey(eq_keys.vk_bytes)
    bad = proofs.copy()
    x = cms.copy()
    bad[3, 5] ^= 1                                                       # off-curve A                      (group 0)
    bad[70, 192:256] = proofs[71, 192:256]                               # a valid curve point, wrong C      (group 1)
    x[200] = cms[201]                                                    # wrong public input                (group 3)
    bad[270, 64:192] = np.frombuffer(po.g2_to_bytes(po.G2.mul(po.G2_GEN, 5)), np.uint8)   # wrong B            (group 4)
    want = np.ones(n, bool)
    want[[3, 70, 200, 270]] = False
    monkeypatch.setenv("LZKP_VERIFY_RLC_MIN", "1000000000")
    independent = vk.verify_batch(bad, x)
    assert np.array_equal(independent, want)
    monkeypatch.setenv("LZKP_VERIFY_RLC_MIN", "64")
    assert np.array_equal(vk.verify_batch(bad, x), want)                 # groups 0, 1, 3, 4 re-verified; group 2 by the combined check
    before = engine.kernel_launches()
    assert vk.verify_batch(proofs, cms).all()                            # all valid: no fallback at all -
    assert engine.kernel_launches() - before == 4                        # per-proof kernel, scalar sums, input shares, combined check
    assert not vk.verify_batch(proofs, np.roll(cms, 1, axis=0)).any()    # all false
    bad2 = proofs.copy()
    bad2[130:140, 0] ^= 1                                                # only malformed proofs in group 2: excluded, the rest of the group still passes combined
    w2 = np.ones(n, bool)
    w2[130:140] = False
    assert np.array_equal(vk.verify_batch(bad2, cms), w2)
    vk.close()
    # membership (129 public inputs: the vk_x combination is where the combined form saves most)
    pkm = engine.ProvingKey(mb_keys.pk_bytes, window_bits=8)
    pkm.circuit_builtin(engine.MEMBERSHIP, 64)
    m = 70
    sets = np.zeros((m, 64), np.uint64)
    lens = np.zeros(m, np.uint32)
    vals = np.zeros(m, np.uint64)
    for i in range(m):
        L = 1 + (i * 5) % 64
        lens[i] = L
        sets[i, :L] = [rng.next_u64() for _ in range(L)]
        vals[i] = sets[i, i % L]
    mp, mcm, mst = pkm.prove_membership_batch(vals, sets, lens, frs(80, m), frs(81, m))
    pkm.close()
    assert not mst.any()
    xs = np.zeros((m, 129, 32), np.uint8)
    for i in range(m):
        pub = po.membership_public_inputs(int.from_bytes(mcm[i].tobytes(), "little"), [int(v) for v in sets[i, :lens[i]]])
        xs[i] = co.fr_array(pub)
    vkm = engine.VerifyingKey(mb_keys.vk_bytes)
    assert vkm.verify_batch(mp, xs).all()
    xs[66, 3, 0] ^= 1                                                    # one set element changed (group 1)
    wm = np.ones(m, bool)
    wm[66] = False
    assert np.array_equal(vkm.verify_batch(mp, xs), wm)
    vkm.close()


def test_latency_form_equals_lane_per_proof_form(eq_keys, mb_keys, co, po, frs, monkeypatch):
    # Calls of up to three waves of one CTA per SM (444 proofs on a B200) give every proof a CTA whose lanes share each Fq12 product (coop.cuh: prepared lines for
    # -gamma / -delta, byte-window tables for vk_x, lane-parallel subgroup ladder); larger calls run one proof per lane
    # (k_verify4).  Both must decide what the oracle's pairing verifier decides (snark.rs:377-401), case by case.
    pk = engine.ProvingKey(eq_keys.pk_bytes, window_bits=WINDOW_BITS)
    pk.circuit_builtin(engine.EQUALITY, 110)
    n = 40
    rng = po.SplitMix64(91)
    a = np.array([rng.next_u64() for _ in range(n)], np.uint64)
    proofs, cms, st = pk.prove_equality_batch(a, a, frs(92, n), frs(93, n))
    pk.close()
    assert not st.any()
    bad, x = proofs.copy(), cms.copy()
    bad[1, 5] ^= 1                                                       # A off the curve
    bad[2, 70] ^= 1                                                      # B off the twist
    bad[3, 200] ^= 1                                                     # C off the curve
    bad[4, 255] ^= 0x20                                                  # stray flag bit
    bad[5, 0:64] = proofs[6, 0:64]                                       # valid points, wrong proof
    bad[7, 192:256] = proofs[8, 192:256]
    bad[9, 64:192] = np.frombuffer(po.g2_to_bytes(po.G2.mul(po.G2_GEN, 77)), np.uint8)
    bad[10, :64] = 0
    bad[10, 63] = 0x40                                                   # A = infinity
    bad[11, 64:192] = 0
    bad[11, 191] = 0x40                                                  # B = infinity
    bad[12, 192:256] = 0
    bad[12, 255] = 0x40                                                  # C = infinity
    bad[13, :] = 0
    bad[14, 31] |= 0x3F                                                  # x >= q
    offsub = bytes.fromhex("02000000000000000000000000000000000000000000000000000000000000000100000000000000000000000000000000000000000000000000000000000000ce0141067f1657f01323d6d1264d1f38c5e0ce2da0ec9950b908934178721f10de5f8f80b5b618595db7a7f6137c7f775a006a5485ac3d962ab99b5979c176ab")
    bad[15, 64:192] = np.frombuffer(offsub, np.uint8)                    # on the twist, outside the r-torsion
    x[16] = cms[17]                                                      # wrong public input
    x[18] = 0xFF                                                         # non-canonical public input
    x[19] = 0                                                            # zero public input
    want = np.ones(n, bool)
    want[[1, 2, 3, 4, 5, 7, 9, 10, 11, 12, 13, 14, 15, 16, 18, 19]] = False
    ovk = po.vk_from_bytes(eq_keys.vk_bytes)
    for k in (0, 5, 10, 11, 16, 19):                                     # the oracle's verdicts on a sample (it is slow)
        try:
            o = po.verify(ovk, [int.from_bytes(x[k].tobytes(), "little")], po.proof_from_bytes(bad[k].tobytes()))
        except Exception:
            o = False
        assert bool(o) == bool(want[k]), k
    vk = engine.VerifyingKey(eq_keys.vk_bytes)
    monkeypatch.setenv("LZKP_VERIFY_COOP_MAX", "512")
    coop = vk.verify_batch(bad, x)
    one_by_one = np.array([vk.verify_batch(bad[k:k + 1], x[k:k + 1])[0] for k in range(n)])
    monkeypatch.setenv("LZKP_VERIFY_COOP_MAX", "0")
    monkeypatch.setenv("LZKP_VERIFY_RLC_MIN", "1000000000")
    lanes = vk.verify_batch(bad, x)
    vk.close()
    assert np.array_equal(coop, want) and np.array_equal(lanes, want) and np.array_equal(one_by_one, want)
    # membership: 129 public inputs through the byte-window tables (64-bit set elements and a full-size commitment)
    pkm = engine.ProvingKey(mb_keys.pk_bytes, window_bits=8)
    pkm.circuit_builtin(engine.MEMBERSHIP, 64)
    m = 12
    sets = np.zeros((m, 64), np.uint64)
    lens = np.zeros(m, np.uint32)
    vals = np.zeros(m, np.uint64)
    for i in range(m):
        L = [1, 2, 63, 64][i % 4]
        lens[i] = L
        sets[i, :L] = [rng.next_u64() for _ in range(L)]
        vals[i] = sets[i, i % L]
    mp, mcm, mst = pkm.prove_membership_batch(vals, sets, lens, frs(94, m), frs(95, m))
    pkm.close()
    assert not mst.any()
    xs = np.zeros((m, 129, 32), np.uint8)
    for i in range(m):
        pub = po.membership_public_inputs(int.from_bytes(mcm[i].tobytes(), "little"), [int(v) for v in sets[i, :lens[i]]])
        xs[i] = co.fr_array(pub)
    xs[3, 1, 0] ^= 1                                                     # first set element changed
    xs[5, 128, 7] ^= 0x80                                                # last input changed
    xs[7, 0, 31] ^= 1                                                    # top byte of the commitment changed
    wm = np.ones(m, bool)
    wm[[3, 5, 7]] = False
    vkm = engine.VerifyingKey(mb_keys.vk_bytes)
    monkeypatch.setenv("LZKP_VERIFY_COOP_MAX", "512")
    coop_m = vkm.verify_batch(mp, xs)
    monkeypatch.setenv("LZKP_VERIFY_COOP_MAX", "0")
    monkeypatch.setenv("LZKP_VERIFY_RLC_MIN", "1000000000")
    lanes_m = vkm.verify_batch(mp, xs)
    vkm.close()
    assert np.array_equal(coop_m, wm) and np.array_equal(lanes_m, wm)


def test_verifier_forms_at_scale(eq_keys, po, frs):
    # Size-independent property at the sizes where the engine changes form (default thresholds): every valid proof is
    # accepted and exactly the corrupted ones are rejected - latency form (300), combined form with role warps (3000: the
    # Miller loop split over two warps, cooperative combined check, proof-by-proof fallback of failing groups) and combined
    # form with one warp per 32 proofs (10 240, beyond one resident wave).
    pk = engine.ProvingKey(eq_keys.pk_bytes, window_bits=WINDOW_BITS)
    pk.circuit_builtin(engine.EQUALITY, 110)
    n = 10240
    rng = po.SplitMix64(1234)
    a = np.array([rng.next_u64() for _ in range(n)], np.uint64)
    proofs, cms, st = pk.prove_equality_batch(a, a, frs(5, n), frs(6, n))
    pk.close()
    assert not st.any()
    vk = engine.VerifyingKey(eq_keys.vk_bytes)
    for m in (300, 3000, n):
        bad, x = proofs[:m].copy(), cms[:m].copy()
        want = np.ones(m, bool)
        spots = [0, 1, 63, 64, m // 3, m // 2, m - 65, m - 1]
        for t, i in enumerate(spots):
            want[i] = False
            if t % 4 == 0:
                bad[i, 7] ^= 1                                           # A off the curve
            elif t % 4 == 1:
                bad[i, 192:256] = proofs[(i + 5) % m, 192:256]           # valid point, wrong C
            elif t % 4 == 2:
                x[i] = cms[(i + 9) % m]                                  # wrong public input
            else:
                bad[i, 64:192] = proofs[(i + 3) % m, 64:192]             # valid G2 point, wrong B
        assert np.array_equal(vk.verify_batch(bad, x), want), m
        assert vk.verify_batch(proofs[:m], cms[:m]).all(), m
    vk.close()
