#!/usr/bin/env python3
"""Inputs for, and consumers of, integration/rust/fixtures-gen (the arkworks / libzkp fixture generator).

  python tools/ark_fixtures.py inputs   (CPU)  writes tests/golden/ark_inputs/<name>.r1cs and <name>.cases:
        this repo's R1CS matrices of libzkp's two circuits (lzkp_builtin_circuit_csr), full assignments z
        (lzkp_builtin_witness) and prover randomness r, s (SplitMix64 seed 2) - committed, deterministic.
  python tools/ark_fixtures.py ours     (GPU)  proves the statements of tests/golden/ark/reference_proofs.bin from the
        REFERENCE's own key files in that directory and writes ours_proofs.bin for `fixtures-gen verify`.

Formats (little-endian):
  <name>.r1cs   "LZR1", u32 m, u32 n_inst, u32 n_wit, then for A, B, C: u32 nnz, u32 rowptr[m + 1], u32 col[nnz],
                32 B canonical val[nnz].  Column 0 = One, 1..n_inst-1 = instance, then witness (ark-relations' order).
  <name>.cases  "LZCS", u32 count, u32 n_vars, then per case z[n_vars] x 32 B, r 32 B, s 32 B.
  *_proofs.bin  u32 count, then per record u32 kind (0 equality, 1 membership), u64 value, u32 set_len, u64 set[],
                32 B commitment, u32 proof_len, proof bytes.
"""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
IN_DIR = os.path.join(ROOT, "tests", "golden", "ark_inputs")
ARK_DIR = os.path.join(ROOT, "tests", "golden", "ark")

EQ_CASES = [5, 42, 0, 2**64 - 1]
MB_CASES = [(2, [1, 2, 3]), (25, [10, 20, 25, 30, 40]), (9, list(range(64)))]


def cases():
    """name -> (kind, param, [(value, other / set)])"""
    from libzkp_b200 import engine
    return {"equality_mimc": (engine.EQUALITY, 110, [(a, a) for a in EQ_CASES]),
            "membership_mimc": (engine.MEMBERSHIP, 64, MB_CASES)}


def scalars(n):
    from bench import SplitMix64
    rng = SplitMix64(2)
    return [(rng.next_fr(), rng.next_fr()) for _ in range(n)]


def write_inputs():
    from libzkp_b200 import engine
    os.makedirs(IN_DIR, exist_ok=True)
    for name, (kind, param, cs) in cases().items():
        (m, n_inst, n_wit), mats = engine.builtin_circuit_csr(kind, param)
        with open(os.path.join(IN_DIR, name + ".r1cs"), "wb") as f:
            f.write(b"LZR1" + struct.pack("<III", m, n_inst, n_wit))
            for rowptr, col, val in mats:
                f.write(struct.pack("<I", len(col)) + rowptr.astype("<u4").tobytes() + col.astype("<u4").tobytes() + val.tobytes())
        rs = scalars(len(cs))
        with open(os.path.join(IN_DIR, name + ".cases"), "wb") as f:
            f.write(b"LZCS" + struct.pack("<II", len(cs), n_inst + n_wit))
            for (v, o), (r, s) in zip(cs, rs):
                z = engine.builtin_witness(kind, param, v, o) if kind == engine.EQUALITY else \
                    engine.builtin_witness(kind, param, v, set_=o)
                f.write(z.tobytes() + r.to_bytes(32, "little") + s.to_bytes(32, "little"))
        print(name, "m =", m, "cases =", len(cs))


def read_cases(name):
    raw = open(os.path.join(IN_DIR, name + ".cases"), "rb").read()
    assert raw[:4] == b"LZCS"
    count, n_vars = struct.unpack_from("<II", raw, 4)
    rec = n_vars * 32 + 64
    out = []
    for i in range(count):
        b = raw[12 + i * rec:12 + (i + 1) * rec]
        out.append((np.frombuffer(b[:n_vars * 32], np.uint8).reshape(n_vars, 32), b[-64:-32], b[-32:]))
    return out


def read_proof_records(path):
    raw = open(path, "rb").read()
    (count,), p, out = struct.unpack_from("<I", raw, 0), 4, []
    for _ in range(count):
        kind, value, n = struct.unpack_from("<IQI", raw, p)
        p += 16
        set_ = list(struct.unpack_from("<%dQ" % n, raw, p))
        p += 8 * n
        cm = raw[p:p + 32]
        (plen,) = struct.unpack_from("<I", raw, p + 32)
        p += 36
        out.append((kind, value, set_, cm, raw[p:p + plen]))
        p += plen
    return out


def write_proof_records(path, recs):
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(recs)))
        for kind, value, set_, cm, proof in recs:
            f.write(struct.pack("<IQI", kind, value, len(set_)) + struct.pack("<%dQ" % len(set_), *set_) + cm +
                    struct.pack("<I", len(proof)) + proof)


def write_ours():
    """Proofs of the reference's statements, made by the GPU engine from the reference's key files."""
    import secrets
    from libzkp_b200 import engine
    R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
    recs = read_proof_records(os.path.join(ARK_DIR, "reference_proofs.bin"))
    keys = {}
    for kind, prefix, param in ((0, "equality_mimc", 110), (1, "membership_mimc", 64)):
        pk = engine.ProvingKey(open(os.path.join(ARK_DIR, prefix + "_pk.bin"), "rb").read(), validate=True, window_bits=12)
        pk.circuit_builtin(engine.EQUALITY if kind == 0 else engine.MEMBERSHIP, param)
        keys[kind] = pk
    out = []
    fr = lambda: np.frombuffer(secrets.randbelow(R).to_bytes(32, "little"), np.uint8)[None]
    for kind, value, set_, cm, _ in recs:
        if kind == 0:
            proofs, _, status = keys[0].prove_equality_batch(np.array([value], np.uint64), np.array([value], np.uint64), fr(), fr(),
                                                             commitments=np.frombuffer(cm, np.uint8)[None])
        else:
            sets = np.zeros((1, 64), np.uint64)
            sets[0, :len(set_)] = set_
            proofs, _, status = keys[1].prove_membership_batch(np.array([value], np.uint64), sets, np.array([len(set_)], np.uint32),
                                                               fr(), fr(), commitments=np.frombuffer(cm, np.uint8)[None])
        assert not status.any()
        out.append((kind, value, set_, cm, proofs[0].tobytes()))
    write_proof_records(os.path.join(ARK_DIR, "ours_proofs.bin"), out)
    print("wrote", len(out), "proofs; now run: cargo run --release -- verify", ARK_DIR)


if __name__ == "__main__":
    {"inputs": write_inputs, "ours": write_ours}[sys.argv[1] if len(sys.argv) > 1 else "inputs"]()
