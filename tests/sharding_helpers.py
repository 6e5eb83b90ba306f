"""Node-range sharding of the snapshot predicate work across GPUs.

Host-side helpers shared by the C-ABI callers and the CPU-only protocol tests: which
super-tiles (256 nodes) a rank filters, and how the per-rank bitmap slices combine.  The
C library applies the same rule in `nhd_solve_staged` (nhd_api.cu)."""
SUPER_NODES = 256


def n_super_tiles(n_nodes: int) -> int:
    return max(1, (n_nodes + SUPER_NODES - 1) // SUPER_NODES)


def shard_super_tiles(n_nodes: int, rank: int, world_size: int):
    """[lo, hi) in super-tiles filtered by `rank` — contiguous, so global node index stays the
    first-fit preference order (Matcher.py:420)."""
    s = n_super_tiles(n_nodes)
    return s * rank // world_size, s * (rank + 1) // world_size


def shard_nodes(n_nodes: int, rank: int, world_size: int):
    lo, hi = shard_super_tiles(n_nodes, rank, world_size)
    return min(lo * SUPER_NODES, n_nodes), min(hi * SUPER_NODES, n_nodes)


def words_per_bitmap(n_nodes: int) -> int:
    return n_super_tiles(n_nodes) * SUPER_NODES // 64


def merge_slices(slices):
    """What the single ncclAllReduce(sum, u64) does: the slices are disjoint and zero elsewhere,
    so the sum is the bitwise OR."""
    out = slices[0].copy()
    for s in slices[1:]:
        assert not (out & s).any(), 'shard slices overlap'
        out += s
    return out
