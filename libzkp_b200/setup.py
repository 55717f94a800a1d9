"""Key generation for the two libzkp circuits when no key files exist (reference:
generate_equality_setup / generate_membership_setup, src/backend/snark.rs:309-339).  The toxic waste
comes from the OS CSPRNG, as the reference's OsRng does, and is dropped after the call."""
from __future__ import annotations

import secrets
from typing import Tuple

from . import engine

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
_CIRCUITS = {"equality_mimc": (engine.EQUALITY, 110), "membership_mimc": (engine.MEMBERSHIP, 64)}


def generate(prefix: str) -> Tuple[bytes, bytes]:
    kind, param = _CIRCUITS[prefix]
    toxic = [1 + secrets.randbelow(R_MOD - 1) for _ in range(5)]
    return engine.setup_builtin(kind, param, toxic)
