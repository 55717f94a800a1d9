"""Host logic of the multi-GPU batch path on CPU: world_size-2 gloo process group, sharding, ordered
gather and fail-fast propagation (the GPU prover is replaced by a stub: there is no CPU fallback)."""
import os
import socket
import sys

import pytest
import torch.multiprocessing as mp

from conftest import ROOT

from libzkp_b200.parallel import shard_range


def test_shard_range_partitions_in_order():
    for n in (0, 1, 2, 7, 4096, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from libzkp_b200 import batch, parallel
    from libzkp_b200.errors import InvalidInput, ProofGenerationFailed
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ops = [("equality", i, i) for i in range(11)]
        calls = []

        def stub(block):                           # stands in for the device call of this rank's shard
            calls.append(len(block))
            return [b"proof-%d" % o[1] for o in block]

        out = parallel.process_operations_sharded(ops, stub)
        assert out == [b"proof-%d" % i for i in range(11)], out          # operation order on every rank
        lo, hi = parallel.shard_range(11, rank, world)
        assert calls == [hi - lo]

        def failing(block):                        # the shard holding operation 8 fails
            if any(o[1] == 8 for o in block):
                raise ProofGenerationFailed("SNARK proof generation failed")
            return [b"x"] * len(block)

        try:
            parallel.process_operations_sharded(ops, failing)
            raise AssertionError("expected a failure on every rank")
        except ProofGenerationFailed as e:
            assert "SNARK proof generation failed" in str(e)

        # registry path: batch id consumed on every rank, unknown id rejected
        bid = 77 + 0
        with batch._lock:
            batch._registry[bid] = list(ops)
        import libzkp_b200.batch as b
        orig = b.prove_operations
        b.prove_operations = lambda blk, rng=None: [bytes([o[1]]) for o in blk]
        try:
            got = parallel.process_batch_sharded(bid)
        finally:
            b.prove_operations = orig
        assert got == [bytes([i]) for i in range(11)]
        try:
            parallel.process_batch_sharded(bid)
            raise AssertionError("batch id must be consumed")
        except InvalidInput:
            pass
        # byte gather of per-rank blocks (the proof bytes of a sharded batch): uneven blocks, operation order
        import numpy as np
        for n_total in (11, 4, 1, 0):
            lo, hi = parallel.shard_range(n_total, rank, world)
            full = (np.arange(n_total * 5, dtype=np.int64) % 251).astype(np.uint8).reshape(n_total, 5)
            got = parallel.gather_rows(full[lo:hi], n_total)
            assert got.shape == (n_total, 5) and np.array_equal(got, full), (n_total, got)
            one = parallel.gather_rows(full[lo:hi], n_total, dst=1)          # results on one rank only
            assert (one is None) == (rank != 1) and (one is None or np.array_equal(one, full))
        try:
            parallel.gather_rows(np.zeros((3, 5), np.uint8), 11)
            raise AssertionError("wrong block size must be rejected")
        except ValueError:
            pass
        q.put((rank, "ok"))
    except Exception as e:                          # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_sharded_batch_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res
