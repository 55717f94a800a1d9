"""Multi-GPU batch proving: one process per GPU (torchrun), proofs sharded across ranks with NO
data-path collective (SURVEY.md §8e "Batched proving": independent units; pk replicated per GPU).

The reference's counterpart is the rayon `par_iter` over batch operations (src/advanced/batch.rs:123-131):
proof-level data parallelism with order-preserving collection and fail-fast error propagation.  Here
every rank proves a contiguous block of the operations on its own GPU; the only communication is the
final gather of the finished proof bytes (a few hundred bytes per proof) to every rank, through
torch.distributed (NCCL on GPUs, gloo in the CPU tests of the host logic).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of n items for `rank`: sizes differ by at most one, order preserved."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank / world size")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist


def process_operations_sharded(ops: Sequence[tuple], prove_block: Callable[[Sequence[tuple]], List[bytes]],
                               rank: Optional[int] = None, world: Optional[int] = None) -> List[bytes]:
    """Prove `ops` (the same list on every rank) with each rank handling its shard; returns, on every
    rank, all proofs in operation order.  `prove_block` turns a block of operations into proof bytes and
    raises on the first failing operation.  If any rank fails, every rank raises the error of the
    earliest failing block (the reference's collect::<ZkpResult<Vec<_>>>() surfaces the first Err)."""
    dist = _dist()
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(len(ops), rank, world)
    mine: List[bytes] = []
    err: Optional[BaseException] = None
    try:
        mine = list(prove_block(ops[lo:hi])) if hi > lo else []
        if len(mine) != hi - lo:
            raise RuntimeError("prove_block returned a wrong number of proofs")
    except Exception as e:                      # noqa: BLE001 - propagated to every rank below
        err = e
    if world == 1:
        if err is not None:
            raise err
        return mine
    gathered: List[object] = [None] * world
    dist.all_gather_object(gathered, (mine, None if err is None else (type(err).__name__, str(err))))
    for blk, e in gathered:                     # rank order == operation order: first failing block wins
        if e is not None:
            from . import errors
            cls = getattr(errors, e[0], None)
            if cls is not None and issubclass(cls, errors.ZkpError):
                raise cls(e[1]) if not issubclass(cls, errors._Runtime) else cls(e[1].split(": ", 1)[-1])
            raise (ValueError if e[0] in ("ValueError", "InvalidInput") else RuntimeError)(e[1])
    out: List[bytes] = []
    for blk, _ in gathered:
        out.extend(blk)
    return out


def process_batch_sharded(batch_id: int, rng=None) -> List[bytes]:
    """`process_batch` across all ranks of the current process group: every rank must have registered
    the same operations under `batch_id` (same insertion order)."""
    from . import batch as _batch

    with _batch._lock:
        ops = _batch._registry.pop(batch_id, None)
    if ops is None:
        from .errors import InvalidInput
        raise InvalidInput(f"Invalid batch ID: {batch_id}")
    return process_operations_sharded(ops, lambda blk: _batch.prove_operations(blk, rng))
