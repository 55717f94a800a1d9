"""Batch API (reference: src/advanced/batch.rs:18-140, utils/composition.rs:337-400).

``process_batch`` keeps the reference's contract — the batch id is consumed, outputs come back in
insertion order, the first failing operation fails the whole batch — but instead of a rayon
``par_iter`` over single proves (batch.rs:123-131) it groups the operations per circuit and issues
ONE device call per group (SURVEY.md §8b "Threading").  Only the two Groth16 operation kinds exist
here; the Bulletproofs/STARK kinds are outside this engine's scope.
"""
from __future__ import annotations

import secrets
import threading
from typing import Dict, List, Sequence, Tuple

from . import proof as _proof
from .errors import InvalidInput

_registry: Dict[int, List[Tuple]] = {}
_lock = threading.Lock()


def create_proof_batch() -> int:                    # batch.rs:36-49: random non-zero u64, unique in process
    with _lock:
        while True:
            bid = secrets.randbits(64)
            if bid != 0 and bid not in _registry:
                _registry[bid] = []
                return bid


def _with_batch(batch_id: int, op: Tuple) -> None:  # batch.rs:51-69
    with _lock:
        if batch_id not in _registry:
            raise InvalidInput(f"Invalid batch ID: {batch_id}")
        _registry[batch_id].append(op)


def batch_add_equality_proof(batch_id: int, val1: int, val2: int) -> None:   # batch.rs:78-81
    _proof._check_u64(val1, val2)
    _proof.validate_equality_params(val1, val2)
    _with_batch(batch_id, ("equality", int(val1), int(val2)))


def batch_add_membership_proof(batch_id: int, value: int, set_: Sequence[int]) -> None:   # batch.rs:92-95
    set_ = [int(v) for v in set_]
    _proof._check_u64(value, *set_)
    _proof.validate_membership_params(value, set_)  # note: no set-size check here, as in the reference
    _with_batch(batch_id, ("membership", int(value), set_))


def get_batch_status(batch_id: int) -> Dict[str, int]:
    with _lock:
        if batch_id not in _registry:
            raise InvalidInput(f"Invalid batch ID: {batch_id}")
        ops = _registry[batch_id]
        return {"total_operations": len(ops), "range_proofs": 0,          # batch.rs:143-172 (all six keys)
                "equality_proofs": sum(1 for o in ops if o[0] == "equality"), "threshold_proofs": 0,
                "membership_proofs": sum(1 for o in ops if o[0] == "membership"), "improvement_proofs": 0,
                "consistency_proofs": 0}


def clear_batch(batch_id: int) -> None:             # batch.rs:175-183: unknown ids are not an error
    with _lock:
        _registry.pop(batch_id, None)


def process_batch(batch_id: int, rng=None) -> List[bytes]:   # batch.rs:110-140
    with _lock:
        ops = _registry.pop(batch_id, None)
    if ops is None:
        raise InvalidInput(f"Invalid batch ID: {batch_id}")
    return prove_operations(ops, rng)


def prove_operations(ops: Sequence[Tuple], rng=None) -> List[bytes]:
    """Grouped proving of a list of operations: one device call per circuit, results in input order."""
    out: List[bytes] = [b""] * len(ops)
    eq = [i for i, o in enumerate(ops) if o[0] == "equality"]
    mb = [i for i, o in enumerate(ops) if o[0] == "membership"]
    # collect::<ZkpResult<Vec<_>>>() surfaces the error of the FIRST failing operation in order;
    # run both groups, remember each group's first failure, raise the earlier one.
    first_err = None
    for idx, fn, args in ((eq, _proof.prove_equality_many, [(ops[i][1], ops[i][2]) for i in eq]),
                          (mb, _proof.prove_membership_many, [(ops[i][1], ops[i][2]) for i in mb])):
        if not idx:
            continue
        try:
            for i, p in zip(idx, fn(args, rng)):
                out[i] = p
        except Exception as e:                      # noqa: BLE001
            pos = _first_failing(ops, idx, e)
            if first_err is None or pos < first_err[0]:
                first_err = (pos, e)
    if first_err is not None:
        raise first_err[1]
    return out


def _first_failing(ops, idx, err) -> int:
    """Index of the first operation of the group that fails on its own validation (else the group's first)."""
    for i in idx:
        o = ops[i]
        try:
            if o[0] == "equality":
                _proof.validate_equality_params(o[1], o[2])
            else:
                _proof.validate_membership_params(o[1], o[2])
                _proof.validate_set_size(o[2], _proof.MAX_SET_SIZE)
        except Exception:                           # noqa: BLE001
            return i
    return idx[0]
