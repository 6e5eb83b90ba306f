"""One rank of the multi-GPU parity check (launched by tests/test_gpu_multirank.py with torch.distributed.run).

Every rank holds a replica of the cluster, filters its shard of the nodes, takes part in the one exchange of the
batch and runs the replicated sweep (nhd_api.cu: world_size > 1).  Rank 0 compares every binding and every final
record with the C oracle; every rank compares a digest of its own results with rank 0's."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import workload
    from nhd_b200.solver import Solver, nccl_unique_id
    from tests import helpers

    rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    failures = []
    oracle_cache = {}
    # (config, nodes, pods, NHD_SHARD_MIN_PAIRS): the small clusters and one run of the full config 4 are forced onto the
    # node-sharded path (filter shards + all-gather); config 4 also runs at the default threshold, where every rank
    # filters the whole cluster and no collective is entered
    for cfg, n_nodes, n_pods, min_pairs in ((3, 4096, 384, '1'), (5, 8192, 512, '1'), (4, None, None, '1'), (4, None, None, None)):
        os.environ.pop('NHD_SHARD_MIN_PAIRS', None)
        if min_pairs is not None:
            os.environ['NHD_SHARD_MIN_PAIRS'] = min_pairs          # read by nhd_create; the same on every rank
        # one communicator per solver handle: a fresh NCCL unique id each time (they are single-use)
        idt = torch.zeros(128, dtype=torch.uint8, device='cuda')
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())
        recs, speed, pods, now = workload.make_workload(cfg, n_nodes, n_pods)
        s = Solver(speed, device=local, rank=rank, world_size=world, nccl_id=nccl_id)
        try:
            s.load_nodes(recs)
            b1 = s.solve_batch(pods[:len(pods) // 2], now[:len(pods) // 2])      # two batches: the second starts from
            b2 = s.solve_batch(pods[len(pods) // 2:], now[len(pods) // 2:])      # the state the first one committed
            final = s.read_nodes()
        finally:
            s.close()
        got = np.concatenate([b1, b2])
        names = [n for n in got.dtype.names if n != 'pad_']
        h = hashlib.sha256()
        for n in names:
            h.update(np.ascontiguousarray(got[n]).tobytes())
        h.update(final.tobytes())
        dig = torch.frombuffer(bytearray(h.digest()), dtype=torch.uint8).cuda()
        ref = dig.clone()
        dist.broadcast(ref, 0)
        if not bool((dig == ref).all()):
            failures.append(f'config {cfg}: rank {rank} differs from rank 0')
        if rank == 0:
            from oracle import binding
            if cfg not in oracle_cache:
                oracle_cache[cfg] = binding.solve(recs, speed, pods, now, threads=max(1, os.cpu_count() or 1))
            ob, orecs = oracle_cache[cfg]
            if not helpers.binding_bytes_equal(ob, got):
                failures.append(f'config {cfg}: bindings differ from the oracle: {helpers.first_binding_diff(ob, got)}')
            if orecs.tobytes() != final.tobytes():
                failures.append(f'config {cfg}: final records differ from the oracle')
            print(f'config {cfg} ({"node-sharded" if min_pairs else "default threshold"}): {len(recs)} nodes x {len(pods)} pods on {world} ranks: '
                  f'{"equal to the oracle" if not failures else failures}', flush=True)
    bad = torch.tensor([len(failures)], device='cuda')
    dist.all_reduce(bad)
    dist.destroy_process_group()
    if failures:
        print('\n'.join(failures), flush=True)
    sys.exit(1 if int(bad.item()) else 0)


if __name__ == '__main__':
    main()
