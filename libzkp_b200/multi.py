"""One large Groth16 proof split across the GPUs of a box (SURVEY.md §8e "Single large proof").

One process per GPU.  Every rank keeps only ITS cost-weighted point ranges of the five proving-key
queries (a, b1, l, h in G1; b2 in G2, ~2.8x per point) resident in HBM.  Per proof:

  1. rank 0 holds the assignment z and r, s; ONE broadcast carries z || r || s (NCCL over NVLink, 32 MiB at 2^20);
  2. every rank starts the MSM slices that only need z on the engine's side streams; the first `map_ranks` ranks
     (1 up to 4 GPUs, 3 at 8: lzkp_pk_shard_info) also run the witness map THEMSELVES (7 tiled NTTs - not worth
     distributing below ~2^24) and share the H query between them, in exchange for smaller z-slices.  h is never
     sent between GPUs (round 1 broadcast 32 MiB of it from rank 0, which serialised every H slice behind rank 0);
  3. a rank that holds points of the A or B1 query multiplies ITS partial sums by s and r as soon as they exist
     (k_scale_ab, beside its remaining MSMs): the two variable-base multiplications of the assembly (0.83 ms as a
     serial chain on rank 0 in the first version) are gone from the tail;
  4. ONE all_gather collects 784 B per rank: 768 B of XYZZ partial sums (A, s A + r B1, L, H, B2) + the status word;
  5. rank 0 adds the partial sums, converts A, B, C to affine and serializes the proof.

Two collectives per proof; everything between them is asynchronous on the caller's CUDA stream.
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _ffi, engine


class ShardedProver:
    SLOT = _ffi.LZKP_PARTIAL_BYTES + 16          # partial sums + status word (+ padding to 16 B)
    collectives_per_proof = 2

    def __init__(self, pk_bytes: bytes, kind: int, param: int, rank: int, world: int, device, window_bits: int = 0):
        import torch
        self.torch = torch
        self.rank, self.world, self.device = rank, world, device
        self.pk = engine.ProvingKey(pk_bytes, window_bits=window_bits, shard_index=rank, shard_count=world)
        self.first, self.count, self.map_ranks = self.pk.shard_info()
        self.holds_h = self.count[3] > 0
        if self.holds_h or rank == 0:
            self.pk.circuit_builtin(kind, param)
        self.n_vars, self.n = self.pk.n_vars, self.pk.n
        u8 = dict(dtype=torch.uint8, device=device)
        self.inp = torch.zeros((self.n_vars + 2, 32), **u8)          # z || r || s: one broadcast
        self.z, self.rs = self.inp[:self.n_vars], self.inp[self.n_vars:]
        self.slot = torch.zeros(self.SLOT, **u8)
        self.slots = torch.zeros((world, self.SLOT), **u8)
        self.proof = torch.zeros(256, **u8)

    def prove(self, z: Optional["np.ndarray"] = None, r: Optional[bytes] = None, s: Optional[bytes] = None,
              resident: bool = False):
        """Rank 0 passes z (n_vars x 32 B), r, s (32 B each); other ranks pass nothing.  With
        resident=True the inputs are taken from self.z / self.rs on rank 0 (already on the device).
        Returns the 256 proof bytes on rank 0 (None elsewhere); raises if a scalar was not canonical."""
        torch = self.torch
        dist = torch.distributed
        st = torch.cuda.current_stream().cuda_stream
        PB = _ffi.LZKP_PARTIAL_BYTES
        if self.rank == 0 and not resident:
            self.z.copy_(torch.from_numpy(np.ascontiguousarray(z, np.uint8).reshape(self.n_vars, 32)), non_blocking=True)
            self.rs.copy_(torch.from_numpy(np.frombuffer(bytes(r) + bytes(s), np.uint8).reshape(2, 32).copy()))
        if self.world > 1:
            dist.broadcast(self.inp, 0)
        # z-only MSMs on the side streams; map ranks run the witness map on this stream beside them, then their H slice
        self.pk.prove_partial_device(self.z.data_ptr(), self.rs[0].data_ptr(), self.rs[1].data_ptr(), 0,
                                     self.slot.data_ptr(), self.slot.data_ptr() + PB, st, phase=3)
        if self.world > 1:
            dist.all_gather_into_tensor(self.slots.view(-1), self.slot)
        else:
            self.slots[0].copy_(self.slot)
        if self.rank != 0:
            return None
        partials = self.slots[:, :PB].contiguous()
        self.pk.prove_combine_device(partials.data_ptr(), self.world, self.rs[0].data_ptr(), self.rs[1].data_ptr(),
                                     self.proof.data_ptr(), st)
        if resident:
            return self.proof
        status = self.slots[:, PB:PB + 4].contiguous().view(torch.int32)
        if int(status.abs().max().item()) != 0:
            raise ValueError("non-canonical scalar in z, r or s")
        return self.proof.cpu().numpy().tobytes()

    def close(self):
        self.pk.close()
