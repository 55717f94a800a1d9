"""Pins the oracle (CPU, no GPU): known answers, public constants, the C restatement against the
Python big-int restatement, and the independent checks of SURVEY.md §8c (naive evaluation,
trapdoor exponents, pairing verifier) that stand in for the golden vectors the reference lacks."""
import hashlib

import numpy as np
import pytest


def b32(v):
    return int(v).to_bytes(32, "little")


# SURVEY.md §8c table
MIMC_KAT = {
    0: ("10719233406374541724350467613780012666485496725628303365024396021904860786422",
        "f62e000d8bb81ca8728a7e614d5b7dc4a8b5348918e2de01214ed7d000dfb217"),
    3: ("1160370759792263048676088440676371550633181481740596225872732281670965422116",
        "2468ac5435e3c9b2686699ac0f3cdfa8d4c81c6f543e98416563ce971fbf9002"),
    5: ("13911153735594629281357134215282608399198733331326451785561755980384426960815",
        "afa7b2cc2d395b182b65895feba2543ecf01d995adf5bb2723af04fe196fc11e"),
    42: ("1476307098280144564201763483369928351521322173786022570940505771869559203198",
         "7e25e6a76fb54fc3889e4df3790213ee8bdd08c22839e163f7e243b1698f4303"),
    123: ("6681278592186177018903891809596725891250998437493079899386788466120164410949",
          "45867c291cc09a54b4a2f2b20a8de6e1d3ce779f57cf6dc80104667c3c78c50e"),
}


def test_mimc_known_answers(po, co, golden):
    for v, (h, cm) in MIMC_KAT.items():
        assert str(po.mimc_hash_native(v)) == h
        assert po.commit_value_snark(v).hex() == cm
        assert co.mimc_hash(v).hex() == cm
        assert golden["mimc"][str(v)]["commitment"] == cm
    v = 2**64 - 1
    assert co.mimc_hash(v).hex() == golden["mimc"][str(v)]["commitment"]


def test_mimc_constants(po, co, golden):
    cs = po.mimc_constants()
    assert str(cs[0]) == "3103100439830505286163135838837437170078884706383160420091567371750505951759"
    assert str(cs[109]) == "18697375384150344081613341308160534263741388935473189017064999337175353965223"
    assert [co.mimc_constant(i) for i in range(110)] == cs
    # derivation restated independently: SHA-256("libzkp_mimc_v1:" || u64_le(i)) as LE int mod r (snark.rs:186-198)
    for i in (0, 1, 57, 109):
        d = hashlib.sha256(b"libzkp_mimc_v1:" + i.to_bytes(8, "little")).digest()
        assert cs[i] == int.from_bytes(d, "little") % po.R_MOD
    assert golden["mimc_constants"]["sha256_all"] == hashlib.sha256(b"".join(b32(c) for c in cs)).hexdigest()


def test_mimc_reference_unit_tests(po):
    # snark.rs:618-631: deterministic, 42 != 43, Fr <-> 32-byte round trip
    assert po.mimc_hash_native(42) == po.mimc_hash_native(42)
    assert po.mimc_hash_native(42) != po.mimc_hash_native(43)
    f = po.mimc_hash_native(123)
    assert po.fr_from_bytes(po.fr_to_bytes(f)) == f
    assert po.fr_from_bytes(b32(po.R_MOD)) is None         # non-canonical rejected


def test_public_constants(po, golden):
    assert po.FR_TWO_ADIC_ROOT == 19103219067921713944291392827692070036145651957329286315305642004821462161904
    assert pow(po.FR_TWO_ADIC_ROOT, 1 << 28, po.R_MOD) == 1 and pow(po.FR_TWO_ADIC_ROOT, 1 << 27, po.R_MOD) != 1
    assert str(po.Domain(512).omega) == "6837567842312086091520287814181175430087169027974246751610506942214842701774"
    assert str(po.Domain(1024).omega) == "3161067157621608152362653341354432744960400845131437947728257924963983317266"
    assert po.G1.on_curve(po.G1_GEN) and po.G2.on_curve(po.G2_GEN)
    assert po.G1.mul(po.G1_GEN, po.R_MOD) is None and po.G2.mul(po.G2_GEN, po.R_MOD) is None
    # EIP-196 vector: 2 * (1, 2)
    x2 = 1368015179489954701390400359078579693043519447331113978918064868415326638035
    y2 = 9918110051302171585080402603319702774565515993150576347155970296011118125764
    assert po.G1.mul(po.G1_GEN, 2) == (x2, y2)
    assert golden["g1_double_gen"] == po.g1_to_bytes((x2, y2)).hex()


def test_circuit_shapes(po, co):
    # SURVEY.md §8(a-0)
    cs = po.equality_circuit(5, 5, po.mimc_hash_native(5))
    assert (cs.num_constraints(), cs.num_instance(), cs.num_witness()) == (332, 2, 332)
    assert cs.is_satisfied()
    c = co.Circuit("equality")
    assert (c.m, c.n_inst, c.n_wit, c.n) == (332, 2, 332, 512)
    m = co.Circuit("membership")
    assert (m.m, m.n_inst, m.n_wit, m.n) == (653, 130, 523, 1024)
    big = co.Circuit("membership", 1024)
    assert (big.m, big.n_inst, big.n_wit, big.n) == (5453, 2050, 3403, 8192)
    assert not po.equality_circuit(5, 5, po.mimc_hash_native(6)).is_satisfied()


def test_c_matrices_and_assignment_match_python(po, co):
    cs = po.equality_circuit(42, 42, po.mimc_hash_native(42))
    c = co.Circuit("equality")
    for which, M in enumerate(cs.matrices()):
        rowptr, col, val = c.matrix(which)
        vals = co.fr_list(val)
        for i, row in enumerate(M):
            got = sorted((int(col[t]), vals[t]) for t in range(rowptr[i], rowptr[i + 1]))
            assert got == sorted((cl, v % po.R_MOD) for v, cl in row if v % po.R_MOD)
    assert co.fr_list(c.assign(42, 42)) == cs.assignment()
    sel, sv, ir = po.membership_inputs(25, [10, 20, 25, 30, 40])
    csm = po.membership_circuit(25, sel, sv, ir, po.mimc_hash_native(25))
    assert csm.is_satisfied()
    m = co.Circuit("membership")
    assert co.fr_list(m.assign(25, set_=[10, 20, 25, 30, 40])) == csm.assignment()
    assert m.assign(7, set_=[1, 2, 3]) is None            # value not in set (snark.rs:415-418)


def test_fft_against_naive_evaluation(po, co, golden):
    g = golden["ntt16"]
    v = [int(x) for x in g["in"]]
    d = po.Domain(16)
    naive = [sum(c * pow(d.omega, i * j, po.R_MOD) for j, c in enumerate(v)) % po.R_MOD for i in range(16)]
    assert d.fft(list(v)) == naive == [int(x) for x in g["fft"]]
    cos = [sum(c * pow(5 * pow(d.omega, i, po.R_MOD), j, po.R_MOD) for j, c in enumerate(v)) % po.R_MOD for i in range(16)]
    assert d.coset_fft(list(v)) == cos == [int(x) for x in g["coset_fft"]]
    assert d.ifft(d.fft(list(v))) == v and d.coset_ifft(d.coset_fft(list(v))) == v
    arr = co.fr_array(v)
    for name, (inv, coset) in {"fft": (0, 0), "ifft": (1, 0), "coset_fft": (0, 1), "coset_ifft": (1, 1)}.items():
        assert co.fr_list(co.ntt(arr, inverse=inv, coset=coset)) == [int(x) for x in g[name]]


@pytest.mark.parametrize("log_n", [1, 5, 10, 13])
def test_c_ntt_matches_python(po, co, frs, log_n):
    a = frs(100 + log_n, 1 << log_n)
    d = po.Domain(1 << log_n)
    v = co.fr_list(a)
    assert co.fr_list(co.ntt(a)) == d.fft(list(v))
    assert co.fr_list(co.ntt(a, inverse=True, coset=True)) == d.coset_ifft(list(v))
    for t in (1, 3):
        assert np.array_equal(co.ntt(a, coset=True, threads=t), co.ntt(a, coset=True))


def test_msm_known_answers(po, co, golden):
    g = golden["msm8"]
    sc = co.fr_array([int(x) for x in g["scalars"]])
    b1 = np.frombuffer(bytes.fromhex("".join(g["g1_bases"])), np.uint8)
    b2 = np.frombuffer(bytes.fromhex("".join(g["g2_bases"])), np.uint8)
    assert co.msm_g1(b1, sc).hex() == g["g1"]
    assert co.msm_g2(b2, sc).hex() == g["g2"]


def test_c_msm_trapdoor_sum(po, co, frs):
    # bases k_i * G with known k_i  =>  MSM == (sum k_i s_i) * G  (SURVEY §8c check 4)
    n = 300
    ks, sc = frs(9, n), frs(10, n)
    sc[7] = 0
    sc[8] = np.frombuffer(b32(1), np.uint8)
    sc[9] = np.frombuffer(b32(po.R_MOD - 1), np.uint8)
    tot = sum(k * s for k, s in zip(co.fr_list(ks), co.fr_list(sc))) % po.R_MOD
    assert co.msm_g1(co.g1_gen_mul(ks), sc) == po.g1_to_bytes(po.G1.mul(po.G1_GEN, tot))
    assert co.msm_g2(co.g2_gen_mul(ks), sc) == po.g2_to_bytes(po.G2.mul(po.G2_GEN, tot))
    assert co.msm_g1(co.g1_gen_mul(ks), sc, threads=1) == co.msm_g1(co.g1_gen_mul(ks), sc, threads=4)


def test_setup_matches_golden_digest(eq_keys, mb_keys, golden):
    assert hashlib.sha256(eq_keys.pk_bytes).hexdigest() == golden["equality"]["pk_sha256"]
    assert eq_keys.vk_bytes.hex() == golden["equality"]["vk"]
    assert len(eq_keys.pk_bytes) == 140208
    assert hashlib.sha256(mb_keys.pk_bytes).hexdigest() == golden["membership"]["pk_sha256"]
    assert hashlib.sha256(mb_keys.vk_bytes).hexdigest() == golden["membership"]["vk_sha256"]


def test_c_prover_reproduces_golden_proofs(co, eq_keys, mb_keys, golden):
    for case in golden["equality"]["proofs"]:
        z = eq_keys.circuit.assign(case["a"], case["a"])
        if "z_sha256" in case:
            assert hashlib.sha256(z.tobytes()).hexdigest() == case["z_sha256"]
            h = eq_keys.circuit.witness_map(z)
            assert hashlib.sha256(h.tobytes()).hexdigest() == case["h_sha256"]
            assert np.all(h[-1] == 0)                      # deg h <= n - 2
        assert co.prove(eq_keys.circuit, eq_keys.opk, z, int(case["r"]), int(case["s"])).hex() == case["proof"]
    for case in golden["membership"]["proofs"]:
        z = mb_keys.circuit.assign(case["value"], set_=case["set"])
        assert co.prove(mb_keys.circuit, mb_keys.opk, z, int(case["r"]), int(case["s"])).hex() == case["proof"]


def test_pairing_verifier_accepts_and_rejects(po, eq_keys, golden):
    # snark.rs:634-641 and tests/integration.rs:25-32,86-90
    vk = po.vk_from_bytes(eq_keys.vk_bytes)
    case = golden["equality"]["proofs"][1]
    assert case["a"] == 42
    proof = po.proof_from_bytes(bytes.fromhex(case["proof"]))
    assert po.verify(vk, po.equality_public_inputs(po.mimc_hash_native(42)), proof)
    assert not po.verify(vk, po.equality_public_inputs(po.mimc_hash_native(99)), proof)


def test_c_batch_prover_matches_single(co, eq_keys, frs):
    a = np.array([3, 4, 5, 6], np.uint64)
    b = a.copy()
    b[2] = 77                                              # a != b -> status, empty proof (snark.rs:344)
    r, s = frs(21, 4), frs(22, 4)
    proofs, status = co.prove_batch(eq_keys.circuit, eq_keys.opk, a, b, None, None, r, s)
    assert list(status != 0) == [False, False, True, False]
    for i in (0, 1, 3):
        z = eq_keys.circuit.assign(int(a[i]), int(a[i]))
        single = co.prove(eq_keys.circuit, eq_keys.opk, z, co.fr_list(r[i])[0], co.fr_list(s[i])[0])
        assert proofs[i].tobytes() == single


def test_serialization_round_trip_and_flags(po, golden):
    p = po.G1.mul(po.G1_GEN, 12345)
    assert po.g1_from_bytes(po.g1_to_bytes(p)) == p
    assert po.g1_to_bytes(None)[-1] == 0x40 and po.g1_to_bytes(None)[:63] == bytes(63)
    q = po.G2.mul(po.G2_GEN, 999)
    assert po.g2_from_bytes(po.g2_to_bytes(q)) == q
    n = po.G1.neg(p)
    assert (po.g1_to_bytes(p)[-1] ^ po.g1_to_bytes(n)[-1]) & 0x80     # exactly one of +-y is "negative"
    env = po.envelope(2, bytes(256), bytes(32))
    assert len(env) == 298 and env[:2] == bytes([2, 2])    # proof/mod.rs:23-36
