"""Proof envelope and the two SNARK proof types (reference: src/proof/mod.rs,
src/proof/equality_proof.rs, src/proof/set_membership.rs) over the device-backed SnarkBackend."""
from __future__ import annotations

import struct
from typing import List, Optional, Sequence

from .errors import InvalidInput, InvalidProofFormat, ProofGenerationFailed
from .snark import MAX_SET_SIZE, SnarkBackend, mimc_commitment

PROOF_VERSION = 2                       # proof/mod.rs:3
MAX_PROOF_TOTAL_BYTES = 1 * 1024 * 1024  # utils/limits.rs:6
MAX_PROOF_PAYLOAD_BYTES = 900 * 1024     # utils/limits.rs:9
MAX_COMMITMENT_BYTES = 256               # utils/limits.rs:12
EQUALITY_SCHEME_ID = 2                  # equality_proof.rs:8
MEMBERSHIP_SCHEME_ID = 4                # set_membership.rs:10


class Proof:
    """``[version u8][scheme u8][proof_len u32 LE][comm_len u32 LE][proof][commitment]`` (proof/mod.rs:23-36)."""

    def __init__(self, scheme: int, proof: bytes, commitment: bytes, version: int = PROOF_VERSION):
        self.version, self.scheme, self.proof, self.commitment = version, scheme, bytes(proof), bytes(commitment)

    def to_bytes(self) -> bytes:
        if len(self.proof) > 0xFFFFFFFF or len(self.commitment) > 0xFFFFFFFF:
            return b""
        return (bytes([self.version, self.scheme]) + struct.pack("<II", len(self.proof), len(self.commitment))
                + self.proof + self.commitment)

    @staticmethod
    def from_bytes(data: bytes) -> "Proof":         # proof/mod.rs:38-85
        if len(data) > MAX_PROOF_TOTAL_BYTES:
            raise InvalidProofFormat(f"proof too large: max {MAX_PROOF_TOTAL_BYTES} bytes")
        if len(data) < 10:
            raise InvalidProofFormat("proof too short for header")
        proof_len, comm_len = struct.unpack_from("<II", data, 2)
        if proof_len > MAX_PROOF_PAYLOAD_BYTES or comm_len > MAX_COMMITMENT_BYTES:
            raise InvalidProofFormat("proof or commitment payload exceeds limit")
        if len(data) != 10 + proof_len + comm_len:
            raise InvalidProofFormat("proof byte length mismatch")
        return Proof(data[1], data[10:10 + proof_len], data[10 + proof_len:], version=data[0])


def commit_value_snark(value: int) -> bytes:        # utils/commitment.rs:14-16
    return mimc_commitment(value)


def validate_equality_params(val1: int, val2: int) -> None:     # utils/validation.rs:21-26
    if val1 != val2:
        raise InvalidInput("values are not equal")


def validate_membership_params(value: int, set_: Sequence[int]) -> None:   # utils/validation.rs:47-60
    if len(set_) == 0:
        raise InvalidInput("set cannot be empty")
    if value not in set_:
        raise InvalidInput(f"value {value} is not in the provided set")


def validate_set_size(set_: Sequence[int], max_size: int) -> None:          # utils/validation.rs:89-98
    if len(set_) > max_size:
        raise InvalidInput(f"set size {len(set_)} exceeds maximum allowed size {max_size}")


def _check_u64(*vals: int) -> None:
    for v in vals:
        if not (0 <= int(v) < 2**64):
            raise OverflowError("value does not fit in u64")     # what PyO3's u64 extraction raises


def _wrap_equality(snark_proof: bytes, commitment: bytes) -> bytes:
    if not snark_proof:
        raise ProofGenerationFailed("SNARK proof generation failed")
    return Proof(EQUALITY_SCHEME_ID, snark_proof, commitment).to_bytes()


def _wrap_membership(snark_proof: bytes, set_: Sequence[int], commitment: bytes) -> bytes:
    if not snark_proof:
        raise ProofGenerationFailed("SNARK membership proof generation failed")
    payload = struct.pack("<I", len(set_)) + b"".join(struct.pack("<Q", v) for v in set_) + snark_proof
    return Proof(MEMBERSHIP_SCHEME_ID, payload, commitment).to_bytes()


def prove_equality(val1: int, val2: int, rng=None) -> bytes:     # equality_proof.rs:10-32
    _check_u64(val1, val2)
    validate_equality_params(val1, val2)
    commitment = commit_value_snark(val1)
    return _wrap_equality(SnarkBackend.prove_equality_zk(val1, val2, commitment, rng), commitment)


def prove_membership(value: int, set_: Sequence[int], rng=None) -> bytes:   # set_membership.rs:12-38
    set_ = list(set_)
    _check_u64(value, *set_)
    validate_membership_params(value, set_)
    validate_set_size(set_, MAX_SET_SIZE)
    commitment = commit_value_snark(value)
    return _wrap_membership(SnarkBackend.prove_membership_zk(value, set_, commitment, rng), set_, commitment)


def prove_equality_many(pairs: Sequence[Sequence[int]], rng=None) -> List[bytes]:
    """The grouped fast path behind process_batch: same results as [prove_equality(a, b) ...].  Commitments,
    proofs and the envelope framing (Proof::to_bytes) all come back from one device call."""
    for a, b in pairs:
        _check_u64(a, b)
        validate_equality_params(a, b)
    out = SnarkBackend.prove_equality_enveloped_batch([p[0] for p in pairs], [p[1] for p in pairs], rng)
    for e in out:
        if not e:
            raise ProofGenerationFailed("SNARK proof generation failed")
    return out


def prove_membership_many(items: Sequence, rng=None) -> List[bytes]:
    for v, s in items:
        _check_u64(v, *s)
        validate_membership_params(v, s)
        validate_set_size(s, MAX_SET_SIZE)
    out = SnarkBackend.prove_membership_enveloped_batch([v for v, _ in items], [list(s) for _, s in items], rng)
    for e in out:
        if not e:
            raise ProofGenerationFailed("SNARK membership proof generation failed")
    return out


# ---------------------------------------------------------------- verification (equality_proof.rs:34-61, set_membership.rs:40-71)
def _parse(proof_bytes: bytes, scheme: int) -> Optional[Proof]:
    """parse_and_validate_proof (utils/proof_helpers.rs): envelope, version, scheme."""
    try:
        p = Proof.from_bytes(bytes(proof_bytes))
    except InvalidProofFormat:
        return None
    if p.version != PROOF_VERSION or p.scheme != scheme:
        return None
    return p


def _embedded_set(payload: bytes):
    """deserialize_embedded_set_prefix: u32 len || u64[len] || rest, len <= MAX_SET_SIZE."""
    if len(payload) < 4:
        return None
    n = struct.unpack_from("<I", payload, 0)[0]
    if n > MAX_SET_SIZE or len(payload) < 4 + 8 * n:
        return None
    return list(struct.unpack_from("<%dQ" % n, payload, 4)), payload[4 + 8 * n:]


def verify_equality_with_commitment(proof: bytes, expected_commitment: bytes) -> bool:
    return verify_equality_many([(proof, expected_commitment)])[0]


def verify_equality(proof: bytes, val1: int, val2: int) -> bool:
    if val1 != val2:
        return False
    return verify_equality_with_commitment(proof, commit_value_snark(val1))


def verify_equality_many(items: Sequence) -> List[bool]:
    """items: (proof_bytes, expected_commitment).  One device call for all of them."""
    out = [False] * len(items)
    idx, proofs, cms = [], [], []
    for i, (pb, cm) in enumerate(items):
        p = _parse(pb, EQUALITY_SCHEME_ID)
        if p is None or len(cm) != 32 or p.commitment != bytes(cm):
            continue
        idx.append(i); proofs.append(p.proof); cms.append(bytes(cm))
    for i, ok in zip(idx, SnarkBackend.verify_equality_zk_batch(proofs, cms)):
        out[i] = ok
    return out


def verify_membership(proof: bytes, set_: Sequence[int]) -> bool:
    return verify_membership_many([(proof, set_)])[0]


def verify_membership_many(items: Sequence) -> List[bool]:
    out = [False] * len(items)
    idx, proofs, sets, cms = [], [], [], []
    for i, (pb, set_) in enumerate(items):
        p = _parse(pb, MEMBERSHIP_SCHEME_ID)
        if p is None or len(p.commitment) != 32:
            continue
        emb = _embedded_set(p.proof)
        if emb is None or not emb[1] or len(set_) != len(emb[0]) or sorted(set_) != sorted(emb[0]):
            continue
        idx.append(i); proofs.append(emb[1]); sets.append(emb[0]); cms.append(p.commitment)
    for i, ok in zip(idx, SnarkBackend.verify_membership_zk_batch(proofs, sets, cms)):
        out[i] = ok
    return out
