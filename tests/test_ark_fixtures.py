"""Parity against the REAL reference stack (arkworks 0.5 + libzkp), through fixtures a maintainer generates with
integration/rust/fixtures-gen (this image has no Rust toolchain, so the files cannot be produced here).

While tests/golden/ark/ is absent every consuming test SKIPS with "PARITY UNPINNED": the repo's bit-exact claims then
rest on its own oracle only (DESIGN.md "Oracle").  One `cargo run` (see fixtures-gen/Cargo.toml) flips these tests on
with no code change:
  <name>_ark_pk.bin / _ark_vk.bin / _ark_proofs.bin   ark-groth16 setup + create_proof_with_reduction(r, s) on this
        repo's R1CS matrices and assignments (tests/golden/ark_inputs/): the oracle and the GPU must reproduce the
        proof BYTES from arkworks' own proving key                  (reference: src/backend/snark.rs:363-373)
  equality_mimc_{pk,vk}.bin, membership_mimc_{pk,vk}.bin, reference_proofs.bin   written by libzkp itself through its
        public API: this repo must load the key files, accept the reference's proofs and prove from its keys
                                                                    (reference: src/backend/snark.rs:40-139,343-452)
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import ark_fixtures as af  # noqa: E402

ARK = af.ARK_DIR
UNPINNED = ("PARITY UNPINNED: tests/golden/ark/%s is absent - generate it with integration/rust/fixtures-gen "
            "(needs cargo; see its Cargo.toml)")


def need(*names):
    for n in names:
        if not os.path.exists(os.path.join(ARK, n)):
            pytest.skip(UNPINNED % n)


def test_committed_fixture_inputs_match_the_engine(tmp_path, monkeypatch):
    # the inputs fixtures-gen consumes are exactly what the engine's circuit synthesis and witness generation produce
    # today (host code of the C ABI, no GPU needed); a change of either must regenerate them
    monkeypatch.setattr(af, "IN_DIR", str(tmp_path))
    af.write_inputs()
    for f in sorted(os.listdir(tmp_path)):
        want = open(os.path.join(ROOT, "tests", "golden", "ark_inputs", f), "rb").read()
        assert open(os.path.join(tmp_path, f), "rb").read() == want, f


def test_fixture_matrices_equal_the_oracles(co):
    # ... and the same matrices are the ones the CPU oracle derives independently from its gadget-level restatement
    import struct
    for name, circ in (("equality_mimc", co.Circuit("equality")), ("membership_mimc", co.Circuit("membership"))):
        raw = open(os.path.join(ROOT, "tests", "golden", "ark_inputs", name + ".r1cs"), "rb").read()
        m, n_inst, n_wit = struct.unpack_from("<III", raw, 4)
        assert (m, n_inst, n_wit) == (circ.m, circ.n_inst, circ.n_wit)


@pytest.mark.parametrize("name,kind", [("equality_mimc", "equality"), ("membership_mimc", "membership")])
def test_oracle_reproduces_arkworks_proof_bytes(co, po, name, kind):
    need(name + "_ark_pk.bin", name + "_ark_vk.bin", name + "_ark_proofs.bin")
    pk_bytes = open(os.path.join(ARK, name + "_ark_pk.bin"), "rb").read()
    proofs = open(os.path.join(ARK, name + "_ark_proofs.bin"), "rb").read()
    circ, opk = co.Circuit(kind), co.ProvingKey(pk_bytes)
    vk = po.vk_from_bytes(open(os.path.join(ARK, name + "_ark_vk.bin"), "rb").read())
    for i, (z, r, s) in enumerate(af.read_cases(name)):
        want = proofs[256 * i:256 * (i + 1)]
        got = co.prove(circ, opk, z, int.from_bytes(r, "little"), int.from_bytes(s, "little"))
        assert got == want, f"{name} case {i}: oracle proof differs from arkworks'"
        public = [int.from_bytes(z[j].tobytes(), "little") for j in range(1, circ.n_inst)]
        assert po.verify(vk, public, po.proof_from_bytes(want))


def test_oracle_verifier_accepts_reference_proofs(po):
    need("reference_proofs.bin", "equality_mimc_vk.bin", "membership_mimc_vk.bin")
    vks = {0: po.vk_from_bytes(open(os.path.join(ARK, "equality_mimc_vk.bin"), "rb").read()),
           1: po.vk_from_bytes(open(os.path.join(ARK, "membership_mimc_vk.bin"), "rb").read())}
    for kind, value, set_, cm, proof in af.read_proof_records(os.path.join(ARK, "reference_proofs.bin")):
        c = int.from_bytes(cm, "little")
        assert cm == po.commit_value_snark(value)                       # MiMC-5 commitment as the reference computes it
        public = po.equality_public_inputs(c) if kind == 0 else po.membership_public_inputs(c, set_)
        assert po.verify(vks[kind], public, po.proof_from_bytes(proof))
        assert not po.verify(vks[kind], [(public[0] + 1) % po.R_MOD] + public[1:], po.proof_from_bytes(proof))


@pytest.mark.gpu
@pytest.mark.parametrize("name,kind,param", [("equality_mimc", 0, 110), ("membership_mimc", 1, 64)])
def test_gpu_reproduces_arkworks_proof_bytes(name, kind, param):
    need(name + "_ark_pk.bin", name + "_ark_proofs.bin")
    from libzkp_b200 import engine
    pk = engine.ProvingKey(open(os.path.join(ARK, name + "_ark_pk.bin"), "rb").read(), validate=True, window_bits=12)
    pk.circuit_builtin(engine.EQUALITY if kind == 0 else engine.MEMBERSHIP, param)
    cases = af.read_cases(name)
    z = np.stack([c[0] for c in cases])
    r = np.stack([np.frombuffer(c[1], np.uint8) for c in cases])
    s = np.stack([np.frombuffer(c[2], np.uint8) for c in cases])
    proofs, status = pk.prove_batch(z, r, s)
    pk.close()
    assert not status.any()
    assert proofs.tobytes() == open(os.path.join(ARK, name + "_ark_proofs.bin"), "rb").read()


@pytest.mark.gpu
def test_gpu_loads_reference_keys_verifies_and_proves(po):
    need("reference_proofs.bin", "equality_mimc_pk.bin", "equality_mimc_vk.bin", "membership_mimc_pk.bin", "membership_mimc_vk.bin")
    from libzkp_b200 import engine
    recs = af.read_proof_records(os.path.join(ARK, "reference_proofs.bin"))
    for kind, prefix, param in ((0, "equality_mimc", 110), (1, "membership_mimc", 64)):
        vk_bytes = open(os.path.join(ARK, prefix + "_vk.bin"), "rb").read()
        pk_bytes = open(os.path.join(ARK, prefix + "_pk.bin"), "rb").read()
        assert pk_bytes[:len(vk_bytes)] == vk_bytes                     # ProvingKey starts with its VerifyingKey
        vk = engine.VerifyingKey(vk_bytes)
        mine = [r for r in recs if r[0] == kind]
        pub = []
        for _, value, set_, cm, _p in mine:
            c = int.from_bytes(cm, "little")
            x = po.equality_public_inputs(c) if kind == 0 else po.membership_public_inputs(c, set_)
            pub.append(b"".join(int(v).to_bytes(32, "little") for v in x))
        ok = vk.verify_batch(np.stack([np.frombuffer(r[4], np.uint8) for r in mine]),
                             np.stack([np.frombuffer(p, np.uint8) for p in pub]))
        assert ok.all(), f"{prefix}: the reference's own proofs must verify"
        vk.close()
        # prove the same statements from the reference's key file; the oracle's pairing check accepts under its vk
        pk = engine.ProvingKey(pk_bytes, validate=True, window_bits=12)
        pk.circuit_builtin(engine.EQUALITY if kind == 0 else engine.MEMBERSHIP, param)
        ovk = po.vk_from_bytes(vk_bytes)
        rs = np.frombuffer((12345).to_bytes(32, "little") + (67890).to_bytes(32, "little"), np.uint8).reshape(2, 32)
        for (_, value, set_, cm, _p), pbytes in zip(mine, pub):
            if kind == 0:
                proofs, _, st = pk.prove_equality_batch(np.array([value], np.uint64), np.array([value], np.uint64), rs[:1], rs[1:])
            else:
                sets = np.zeros((1, 64), np.uint64)
                sets[0, :len(set_)] = set_
                proofs, _, st = pk.prove_membership_batch(np.array([value], np.uint64), sets, np.array([len(set_)], np.uint32), rs[:1], rs[1:])
            assert not st.any()
            x = [int.from_bytes(pbytes[32 * i:32 * i + 32], "little") for i in range(len(pbytes) // 32)]
            assert po.verify(ovk, x, po.proof_from_bytes(proofs[0].tobytes()))
        pk.close()


def test_consuming_tests_work_on_stand_in_fixtures(tmp_path, monkeypatch, co, po, trapdoor):
    # The mechanism itself, exercised end to end with the ORACLE standing in for fixtures-gen (same file formats): the
    # consumers above run and pass on files laid out the way the Rust generator writes them.  This pins nothing - it
    # only guarantees that dropping the real files in needs no code change.
    import test_ark_fixtures as me
    monkeypatch.setattr(me, "ARK", str(tmp_path))
    recs = []
    for name, kind in (("equality_mimc", "equality"), ("membership_mimc", "membership")):
        circ = co.Circuit(kind)
        pk_bytes, vk_bytes = circ.setup(trapdoor)
        opk = co.ProvingKey(pk_bytes)
        (tmp_path / f"{name}_ark_pk.bin").write_bytes(pk_bytes)
        (tmp_path / f"{name}_ark_vk.bin").write_bytes(vk_bytes)
        (tmp_path / f"{name}_pk.bin").write_bytes(pk_bytes)
        (tmp_path / f"{name}_vk.bin").write_bytes(vk_bytes)
        out = b""
        for z, r, s in af.read_cases(name):
            out += co.prove(circ, opk, z, int.from_bytes(r, "little"), int.from_bytes(s, "little"))
        (tmp_path / f"{name}_ark_proofs.bin").write_bytes(out)
        if kind == "equality":
            for i, a in enumerate(af.EQ_CASES):
                recs.append((0, a, [], po.commit_value_snark(a), out[256 * i:256 * (i + 1)]))
        else:
            for i, (v, st) in enumerate(af.MB_CASES):
                recs.append((1, v, st, po.commit_value_snark(v), out[256 * i:256 * (i + 1)]))
    af.write_proof_records(str(tmp_path / "reference_proofs.bin"), recs)
    assert af.read_proof_records(str(tmp_path / "reference_proofs.bin")) == recs
    me.test_oracle_reproduces_arkworks_proof_bytes(co, po, "equality_mimc", "equality")
    me.test_oracle_reproduces_arkworks_proof_bytes(co, po, "membership_mimc", "membership")
    me.test_oracle_verifier_accepts_reference_proofs(po)


@pytest.mark.gpu
def test_gpu_consumers_work_on_stand_in_fixtures(tmp_path, monkeypatch, co, po, trapdoor):
    test_consuming_tests_work_on_stand_in_fixtures(tmp_path, monkeypatch, co, po, trapdoor)
    import test_ark_fixtures as me
    me.test_gpu_reproduces_arkworks_proof_bytes("equality_mimc", 0, 110)
    me.test_gpu_reproduces_arkworks_proof_bytes("membership_mimc", 1, 64)
    me.test_gpu_loads_reference_keys_verifies_and_proves(po)
