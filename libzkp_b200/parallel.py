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


def gather_rows(local, n_total: int, rank: Optional[int] = None, world: Optional[int] = None, device=None,
                dst: Optional[int] = None):
    """All ranks hold the rows [shard_range(n_total, rank, world)) of a 2-D uint8 / int32 array; returns the full array
    (n_total rows, operation order) on every rank - or, with `dst`, only on that rank (None elsewhere: the reference
    hands a batch's results to ONE caller, so the other ranks need not pay for the copy back).  One collective of
    equal-sized, zero-padded blocks (NCCL wants equal sizes; block sizes differ by at most one row) - bytes only, no
    pickling, so a 65 536-proof batch gathers in milliseconds.  `device` = the CUDA device of this rank for NCCL, None
    for gloo (CPU tensors)."""
    import numpy as np
    import torch
    dist = _dist()
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    local = np.ascontiguousarray(local)
    lo, hi = shard_range(n_total, rank, world)
    if local.shape[0] != hi - lo:
        raise ValueError("local block has the wrong number of rows")
    if world == 1:
        return local
    rows = -(-n_total // world)                                  # largest block
    if hi - lo == rows:
        t = torch.from_numpy(local)
    else:
        pad = np.zeros((rows,) + local.shape[1:], local.dtype)
        pad[:hi - lo] = local
        t = torch.from_numpy(pad)
    if device is not None:
        t = t.to(device, non_blocking=True)
    if dst is None:
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out.view(-1), t.view(-1))
    else:
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device) if rank == dst else None
        dist.gather(t, list(out.unbind(0)) if rank == dst else None, dst=dst)
        if rank != dst:
            return None
    out = out.cpu().numpy()
    if n_total == rows * world:                                  # even split: the gathered buffer IS the result
        return out.reshape((n_total,) + local.shape[1:])
    return np.concatenate([out[g, :shard_range(n_total, g, world)[1] - shard_range(n_total, g, world)[0]] for g in range(world)])


def prove_mixed_enveloped_sharded(pk_eq, pk_mb, eq_vals, mb_vals, mb_sets, mb_lens, r_eq, s_eq, r_mb, s_mb,
                                  rank: int, world: int, device=None, dst: Optional[int] = None):
    """BASELINE.json configs[4]: one mixed batch (the caller's even operations are equality proofs, the odd ones
    membership proofs, as process_batch groups them per circuit - src/advanced/batch.rs:123-131) sharded
    proof-parallel: rank g proves block g of each kind on its own GPU (both keys resident), libzkp envelopes are
    written on the device, and the finished bytes are gathered in operation order on every rank (dst=None) or on rank
    `dst` only (the other ranks return None).  Lengths and status words travel inside the same rows as the envelopes:
    two collectives per batch.  All array arguments are the FULL batch on every rank.
    Returns (eq_env, eq_len, eq_status, mb_env, mb_len, mb_status)."""
    ne, nm = len(eq_vals), len(mb_vals)
    elo, ehi = shard_range(ne, rank, world)
    mlo, mhi = shard_range(nm, rank, world)
    env_e, len_e, st_e = pk_eq.prove_equality_enveloped(eq_vals[elo:ehi], eq_vals[elo:ehi], r_eq[elo:ehi], s_eq[elo:ehi])
    env_m, len_m, st_m = pk_mb.prove_membership_enveloped(mb_vals[mlo:mhi], mb_sets[mlo:mhi], mb_lens[mlo:mhi],
                                                          r_mb[mlo:mhi], s_mb[mlo:mhi])
    import numpy as np

    def pack(env, lens, st):                       # [envelope bytes | len u32 | status i32] per row: one collective per kind
        meta = np.stack([lens.astype(np.uint32).view(np.int32), st.astype(np.int32)], 1).view(np.uint8).reshape(len(lens), 8)
        return np.concatenate([env, meta], 1)

    def unpack(rows, width):
        meta = np.ascontiguousarray(rows[:, width:]).view(np.int32).reshape(-1, 2)
        return rows[:, :width], meta[:, 0].view(np.uint32), meta[:, 1]
    ge = gather_rows(pack(env_e, len_e, st_e), ne, rank, world, device, dst)
    gm = gather_rows(pack(env_m, len_m, st_m), nm, rank, world, device, dst)
    if ge is None:
        return None
    return unpack(ge, env_e.shape[1]) + unpack(gm, env_m.shape[1])
