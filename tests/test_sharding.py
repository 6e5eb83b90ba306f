"""CPU-only, world size 2 over gloo: the multi-GPU protocol of DESIGN.md section 6.  Each rank
evaluates the feasibility of ITS node shard (here with the oracle standing in for the CUDA filter),
one all-reduce(sum) over zero-padded disjoint slices merges the bitmaps, and the merged result must
equal the unsharded one on every rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import workload
from tests import sharding_helpers as sharding


def test_shard_ranges_partition_the_cluster():
    for n in (1, 255, 256, 257, 4096, 65536, 262144, 1000):
        for ws in (1, 2, 3, 4, 8):
            spans = [sharding.shard_nodes(n, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]                      # contiguous, ordered: index = first-fit preference
            assert all(lo % 64 == 0 for lo, _ in spans)   # a 64-bit word never straddles two ranks
            assert sharding.words_per_bitmap(n) * 64 >= n


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_nodes, n_pods, out_q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oracle import binding as ob
    recs, speed, pods, now = workload.make_workload(3, n_nodes=n_nodes, n_pods=n_pods)
    W = sharding.words_per_bitmap(n_nodes)
    lo, hi = sharding.shard_nodes(n_nodes, rank, world)
    # this rank's slice of the per-pod feasibility bitmaps, zero elsewhere
    bits = np.zeros((n_pods, W * 64), dtype=np.uint8)
    for i in range(n_pods):
        bits[i, lo:hi] = ob.candidates(recs[lo:hi], speed, pods[i], now=1e9)
    words = np.packbits(bits, axis=1, bitorder='little').view('<u8').astype(np.int64)
    t = torch.from_numpy(words.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)                  # the one collective per batch
    merged = np.unpackbits(t.numpy().astype('<i8').view(np.uint8).reshape(n_pods, -1), axis=1, bitorder='little')
    full = np.stack([ob.candidates(recs, speed, pods[i], now=1e9) for i in range(n_pods)])
    ok = bool(np.array_equal(merged[:, :n_nodes], full))
    # replicated sweep: every rank schedules the whole batch on its full replica -> identical bindings
    b, _ = ob.solve(recs, speed, pods, now)
    digest = torch.tensor([int(np.frombuffer(b.tobytes(), dtype=np.uint8).astype(np.int64).sum())])
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    same = all(int(g) == int(digest) for g in gathered)
    if rank == 0:
        out_q.put((ok, same))
    dist.destroy_process_group()


def test_sharded_filter_allreduce_matches_unsharded_gloo():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1000, 12, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, same = q.get(timeout=5)
    assert ok and same
