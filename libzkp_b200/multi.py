"""One large Groth16 proof split across the GPUs of a box (SURVEY.md §8e "Single large proof").

One process per GPU.  Every rank keeps only ITS cost-weighted point ranges of the five proving-key
queries (a, b1, l, h in G1; b2 in G2, ~2.8x per point) resident in HBM.  Per proof:

  1. rank 0 holds the assignment z and r, s; they are broadcast (NCCL over NVLink, 32 B x n_vars);
  2. every rank starts the MSM slices that only need z, while rank 0 runs the witness map (7 tiled NTTs - not
     worth distributing below ~2^24; rank 0 holds a smaller slice in exchange) and broadcasts h;
  3. every rank adds its slice of the H MSM -> 768 B of partial sums;
  4. the partial sums are gathered on rank 0, which adds them, assembles and serializes the proof.

The only data-path collectives are the two broadcasts and one 768-byte-per-rank gather.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _ffi, engine


class ShardedProver:
    def __init__(self, pk_bytes: bytes, kind: int, param: int, rank: int, world: int, device, window_bits: int = 0):
        import torch
        self.torch = torch
        self.rank, self.world, self.device = rank, world, device
        self.pk = engine.ProvingKey(pk_bytes, window_bits=window_bits, shard_index=rank, shard_count=world)
        if rank == 0:
            self.pk.circuit_builtin(kind, param)
        self.n_vars, self.n = self.pk.n_vars, self.pk.n
        u8 = dict(dtype=torch.uint8, device=device)
        self.z = torch.zeros((self.n_vars, 32), **u8)
        self.h = torch.zeros((self.n, 32), **u8)
        self.rs = torch.zeros((2, 32), **u8)
        self.partial = torch.zeros(_ffi.LZKP_PARTIAL_BYTES, **u8)
        self.partials = torch.zeros((world, _ffi.LZKP_PARTIAL_BYTES), **u8)
        self.status = torch.zeros(1, dtype=torch.int32, device=device)
        self.proof = torch.zeros(256, **u8)

    def prove(self, z: Optional["np.ndarray"] = None, r: Optional[bytes] = None, s: Optional[bytes] = None,
              resident: bool = False):
        """Rank 0 passes z (n_vars x 32 B), r, s (32 B each); other ranks pass nothing.  With
        resident=True the inputs are taken from self.z / self.rs on rank 0 (already on the device).
        Returns the 256 proof bytes on rank 0 (None elsewhere); raises if a scalar was not canonical."""
        torch = self.torch
        dist = torch.distributed
        st = torch.cuda.current_stream().cuda_stream
        if self.rank == 0 and not resident:
            self.z.copy_(torch.from_numpy(np.ascontiguousarray(z, np.uint8).reshape(self.n_vars, 32)), non_blocking=True)
            self.rs.copy_(torch.from_numpy(np.frombuffer(bytes(r) + bytes(s), np.uint8).reshape(2, 32).copy()))
        if self.world > 1:
            dist.broadcast(self.z, 0)
            dist.broadcast(self.rs, 0)
        # phase 1: the MSMs that only need z start on the engine's side streams on every rank ...
        self.pk.prove_partial_device(self.z.data_ptr(), self.rs[0].data_ptr(), self.rs[1].data_ptr(), 0, 0,
                                     self.status.data_ptr(), st, phase=1)
        # ... while rank 0 runs the witness map and h travels to the ranks that hold h_query ranges
        if self.rank == 0:
            self.pk.witness_map_device(self.z.data_ptr(), self.h.data_ptr(), st)
        if self.world > 1:
            dist.broadcast(self.h, 0)
        self.pk.prove_partial_device(0, self.rs[0].data_ptr(), self.rs[1].data_ptr(), self.h.data_ptr(),
                                     self.partial.data_ptr(), self.status.data_ptr(), st, phase=2)
        if self.world > 1:
            dist.all_gather_into_tensor(self.partials.view(-1), self.partial)
            dist.all_reduce(self.status, op=dist.ReduceOp.MAX)
        else:
            self.partials[0].copy_(self.partial)
        if self.rank != 0:
            return None
        self.pk.prove_combine_device(self.partials.data_ptr(), self.world, self.rs[0].data_ptr(), self.rs[1].data_ptr(),
                                     self.proof.data_ptr(), st)
        if resident:
            return self.proof
        if int(self.status.item()) != 0:
            raise ValueError("non-canonical scalar in z, r or s")
        return self.proof.cpu().numpy().tobytes()

    def close(self):
        self.pk.close()
