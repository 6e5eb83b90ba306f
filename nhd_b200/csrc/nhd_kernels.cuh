/*
 * nhd_kernels.cuh — sm_100a kernels of the NHD placement solver.
 *
 *   ingest_kernel   AoS wire records -> tiled device layout (and the inverse, export_kernel)
 *   filter_kernel   snapshot predicate: F[type][node] bitmaps, NOGPU and BUSY bitmaps.
 *                   Node tiles are staged into shared memory with 1-D TMA bulk copies
 *                   (cp.async.bulk + mbarrier, double buffered); one thread per node,
 *                   warp ballots produce the bitmap words.  Replaces the per-node loops of
 *                   Matcher.FilterPodResources / FilterNumaTopology / IntersectResources
 *                   (nhd/Matcher.py:65-391) and NHDScheduler.InitialNodeFilter
 *                   (nhd/NHDScheduler.py:235-247).
 *   sweep_kernel    sequential select + assign: for every pod in order, first-fit over the
 *                   bitmaps (warp ballot scan), live re-validation on the node's current
 *                   record, NUMA mapping choice (CPython set order), physical core / GPU /
 *                   NIC assignment and state update.  Replaces Matcher.SelectNode /
 *                   GetNumaGroupIdx (nhd/Matcher.py:393-452), Node.SetPhysicalIdsFromMapping
 *                   (nhd/Node.py:663-841) and the SetBusy / ClaimPodNICResources calls of
 *                   NHDScheduler.AttemptScheduling (nhd/NHDScheduler.py:289,302-304).
 *
 * Device layout of the node array ("tiled AoSoA"): tiles of 32 nodes, 4 KB each:
 *     byte offset(node n, 16-byte chunk c) = (n / 32) * 4096 + c * 512 + (n % 32) * 16
 * so a warp reading chunk c of 32 consecutive nodes touches 512 contiguous bytes both in
 * global memory (coalesced, TMA-friendly) and in shared memory (bank-conflict free).
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "nhd_core.cuh"

namespace nhd {

constexpr int TILE_NODES = 32;
constexpr int TILE_BYTES = 4096;
constexpr int REC_CHUNKS = 8;
constexpr int SUPER_TILES = 8;                       /* tiles per CTA stage */
constexpr int SUPER_NODES = TILE_NODES * SUPER_TILES;  /* 256 */
constexpr int SUPER_BYTES = TILE_BYTES * SUPER_TILES;  /* 32 KB */
constexpr int FILTER_THREADS = SUPER_NODES;
constexpr int FILTER_STAGES = 2;
constexpr int TYPES_SMEM_MAX = 128;                  /* pod types cached in shared memory */

constexpr int MEMO_SLOTS = 1 << 16;

union RecU {
    nhd_node_rec r;
    uint4 q[REC_CHUNKS];
    __device__ RecU() {}
};

__device__ __forceinline__ size_t chunk_off(int node, int c)
{
    return (size_t)(node >> 5) * TILE_BYTES + (size_t)c * 512 + (size_t)(node & 31) * 16;
}

__device__ __forceinline__ void load_rec(const uint8_t* nodes, int node, RecU& u)
{
#pragma unroll
    for (int c = 0; c < REC_CHUNKS; c++)
        u.q[c] = *reinterpret_cast<const uint4*>(nodes + chunk_off(node, c));
}

/* only chunks 0-3 hold mutable state (used mask, gpu/nic/hugepage words, busy_time) */
__device__ __forceinline__ void store_rec_mutable(uint8_t* nodes, int node, const RecU& u)
{
#pragma unroll
    for (int c = 0; c < 4; c++)
        *reinterpret_cast<uint4*>(nodes + chunk_off(node, c)) = u.q[c];
}

/* ------------------------------------------------------------------ layout conversion */

__global__ void ingest_kernel(const uint4* __restrict__ aos, const int32_t* __restrict__ idx, int n,
                              uint8_t* __restrict__ tiled)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;          /* one thread per (record, chunk) */
    if (i >= n * REC_CHUNKS) return;
    int rec = i >> 3, c = i & 7;
    int node = idx ? idx[rec] : rec;
    *reinterpret_cast<uint4*>(tiled + chunk_off(node, c)) = aos[i];
}

__global__ void export_kernel(const uint8_t* __restrict__ tiled, int first, int n, uint4* __restrict__ aos)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * REC_CHUNKS) return;
    int rec = i >> 3, c = i & 7;
    aos[i] = *reinterpret_cast<const uint4*>(tiled + chunk_off(first + rec, c));
}

/* ------------------------------------------------------------------ TMA / mbarrier helpers */

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
/* 1-D bulk async copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP) */
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

/* ------------------------------------------------------------------ snapshot predicate kernel */

struct FilterArgs {
    const uint8_t* nodes;        /* tiled node array, padded to SUPER_NODES */
    const PodType* types;
    int n_types;
    int n_nodes;
    int super_lo, super_hi;      /* this rank's shard, in super-tiles */
    int words;                   /* u64 words per bitmap */
    uint64_t* bitmaps;           /* [n_types + 2][words]: F[0..T), NOGPU, BUSY */
    double now0;                 /* clock of the first pod, for the BUSY snapshot */
    double min_busy;
    double cap[NHD_MAX_SPEED_CLASSES];
};

__global__ void __launch_bounds__(FILTER_THREADS, 2)
filter_kernel(const FilterArgs a)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* stage_buf = smem;                                          /* FILTER_STAGES x SUPER_BYTES */
    PodType* s_types = reinterpret_cast<PodType*>(smem + FILTER_STAGES * SUPER_BYTES);
    __shared__ __align__(8) uint64_t full_bar[FILTER_STAGES];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_my = (a.super_hi - a.super_lo - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    if (n_my <= 0) return;

    if (tid == 0) {
        for (int s = 0; s < FILTER_STAGES; s++) mbar_init(&full_bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int it) {                                          /* thread 0 only */
        int st = a.super_lo + blockIdx.x + it * gridDim.x;
        int s = it % FILTER_STAGES;
        mbar_expect_tx(&full_bar[s], SUPER_BYTES);
        tma_load_1d(stage_buf + s * SUPER_BYTES, a.nodes + (size_t)st * SUPER_BYTES, SUPER_BYTES, &full_bar[s]);
    };
    if (tid == 0)
        for (int it = 0; it < FILTER_STAGES && it < n_my; it++) issue(it);

    /* pod types: small tables are copied to shared memory while the first tile is in flight */
    const bool types_in_smem = a.n_types <= TYPES_SMEM_MAX;
    if (types_in_smem) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.types);
        uint32_t* dst = reinterpret_cast<uint32_t*>(s_types);
        const int nw = a.n_types * (int)(sizeof(PodType) / 4);
        for (int i = tid; i < nw; i += FILTER_THREADS) dst[i] = src[i];
        __syncthreads();
    }
    const PodType* types = types_in_smem ? s_types : a.types;
    uint32_t* out32 = reinterpret_cast<uint32_t*>(a.bitmaps);
    const size_t words32 = (size_t)a.words * 2;

    for (int it = 0; it < n_my; it++) {
        const int s = it % FILTER_STAGES;
        const int st = a.super_lo + blockIdx.x + it * gridDim.x;
        mbar_wait(&full_bar[s], (it / FILTER_STAGES) & 1);

        RecU u;
        const uint8_t* tile = stage_buf + s * SUPER_BYTES + warp * TILE_BYTES;
#pragma unroll
        for (int c = 0; c < REC_CHUNKS; c++)
            u.q[c] = *reinterpret_cast<const uint4*>(tile + c * 512 + lane * 16);
        __syncthreads();                                                /* stage s fully consumed */
        if (tid == 0 && it + FILTER_STAGES < n_my) issue(it + FILTER_STAGES);

        const int node = st * SUPER_NODES + tid;
        const bool valid = node < a.n_nodes;
        const size_t w32 = (size_t)node >> 5;

        for (int t = 0; t < a.n_types; t++) {
            bool f = valid && node_feasible(u.r, types[t], a.cap);
            uint32_t bal = __ballot_sync(0xFFFFFFFFu, f);
            if (lane == 0) out32[(size_t)t * words32 + w32] = bal;
        }
        uint32_t nog = __ballot_sync(0xFFFFFFFFu, valid && u.r.n_gpus == 0);
        uint32_t bsy = __ballot_sync(0xFFFFFFFFu, valid && node_busy(u.r, a.now0, a.min_busy));
        if (lane == 0) {
            out32[(size_t)a.n_types * words32 + w32] = nog;
            out32[(size_t)(a.n_types + 1) * words32 + w32] = bsy;
        }
    }
}

/* ------------------------------------------------------------------ select + assign sweep */

struct SweepArgs {
    uint8_t* nodes;
    const PodType* types;
    const int32_t* pod_type;
    const double* now;
    nhd_binding* out;
    int n_pods, n_types, n_nodes, words;
    uint64_t* bitmaps;           /* [n_types + 2][words] */
    int32_t* cursors;            /* [n_types][2] first possibly non-empty word (pass 0 = GPU-less nodes) */
    int32_t* busy_list;          /* nodes whose BUSY bit is set */
    uint64_t* memo;              /* MEMO_SLOTS x 2 words */
    double min_busy;
    double cap[NHD_MAX_SPEED_CLASSES];
};

/* mapping memo: (K, G, stage masks) -> (gtuple index, misc NUMA) for K^(G+1) <= 32 */
__device__ __forceinline__ uint64_t memo_mix(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

__device__ bool choose_mapping_memo(uint64_t* memo, int K, int G, const TMask& ma, const TMask& mb, const TMask& mc,
                                    int* ps, int* ms)
{
    const int nq = ipow(K, G + 1);
    if (nq > 32) return choose_mapping(K, G, ma, mb, mc, ps, ms);
    const uint64_t key = (ma.w[0] & 0xFFFF) | ((mb.w[0] & 0xFFFFFFFFULL) << 16) | ((mc.w[0] & 0xFFFF) << 48);
    const uint64_t tag = 0x8000000000000000ULL | ((uint64_t)K << 8) | (uint64_t)G;     /* bit 63 = occupied */
    uint64_t h = memo_mix(key ^ (tag * 0x9E3779B97F4A7C15ULL));
    int free_slot = -1;
    for (int probe = 0; probe < 8; probe++) {
        const int s = (int)((h + probe) & (MEMO_SLOTS - 1));
        const uint64_t k0 = memo[2 * s], k1 = memo[2 * s + 1];
        if (!(k1 >> 63)) { free_slot = s; break; }
        if (k0 == key && (k1 & 0x800000000000FFFFULL) == tag) {
            const int v = (int)((k1 >> 16) & 0xFFFF);
            if (v == 0xFFFF) return false;
            *ps = v & 0xFF; *ms = v >> 8;
            return true;
        }
    }
    const bool ok = choose_mapping(K, G, ma, mb, mc, ps, ms);
    if (free_slot >= 0) {
        const uint64_t v = ok ? (uint64_t)((*ps & 0xFF) | (*ms << 8)) : 0xFFFFULL;
        memo[2 * free_slot] = key;
        memo[2 * free_slot + 1] = tag | (v << 16);
    }
    return ok;
}

__device__ bool evaluate_mapping_memo(uint64_t* memo, const nhd_node_rec& r, const PodType& t, const double* cap, Mapping* out)
{
    const int K = r.n_numa, G = t.G;
    const uint64_t gsw = t.pci ? free_gpus_per_switch(r) : 0;
    TMask ma, mb, mc;
    if (!stage_masks(r, t, cap, gsw, ma, mb, mc)) return false;
    int ps, ms;
    if (!choose_mapping_memo(memo, K, G, ma, mb, mc, &ps, &ms)) return false;
    tuple_digits(ps, K, G, out->gpu_numa);
    out->misc_numa = (uint8_t)ms;
    nic_first_fit(r, t, out->gpu_numa, K, cap, gsw, out->nic_idx, out->nic_li);
    return true;
}

/*
 * One warp walks the pods in order.  Bitmap scanning, busy-list maintenance and record
 * transfers use all 32 lanes; the evaluation of the (usually single) candidate and the
 * assignment run on lane 0.
 */
__global__ void __launch_bounds__(32, 1)
sweep_kernel(const SweepArgs a)
{
    const int lane = threadIdx.x;
    const int W = a.words;
    uint64_t* const NOGPU = a.bitmaps + (size_t)a.n_types * W;
    uint64_t* const BUSY = a.bitmaps + (size_t)(a.n_types + 1) * W;

    /* cursors start at word 0 */
    for (int i = lane; i < a.n_types * 2; i += 32) a.cursors[i] = 0;

    /* busy list from the BUSY snapshot (filter_kernel evaluated it for now[0]) */
    int n_busy = 0;
    for (int w0 = 0; w0 < W; w0 += 32) {
        uint64_t word = (w0 + lane < W) ? BUSY[w0 + lane] : 0;
        int cnt = popc64(word);
        int pre = cnt;                                   /* inclusive scan over lanes */
        for (int d = 1; d < 32; d <<= 1) {
            int v = __shfl_up_sync(0xFFFFFFFFu, pre, d);
            if (lane >= d) pre += v;
        }
        int pos = n_busy + pre - cnt;
        while (word) {
            a.busy_list[pos++] = (w0 + lane) * 64 + ctz64(word);
            word &= word - 1;
        }
        n_busy += __shfl_sync(0xFFFFFFFFu, pre, 31);
    }
    __syncwarp();
    double cur_now = a.n_pods > 0 ? a.now[0] : 0.0;

    for (int i = 0; i < a.n_pods; i++) {
        const int ti = a.pod_type[i];
        const PodType& t = a.types[ti];
        const double now = a.now[i];
        nhd_binding* bout = &a.out[i];

        /* ---- busy window bookkeeping when the clock moved (Node.py:847-850) ---- */
        if (now != cur_now) {
            if (now < cur_now) {
                /* clock went backwards: rebuild from the records */
                n_busy = 0;
                for (int n0 = 0; n0 < W * 64; n0 += 32) {
                    int n = n0 + lane;
                    bool b = false;
                    if (n < a.n_nodes) {
                        double bt = *reinterpret_cast<const double*>(a.nodes + chunk_off(n, 3));
                        b = (now - bt) < a.min_busy;
                    }
                    uint32_t bal = __ballot_sync(0xFFFFFFFFu, b);
                    if (lane == 0) reinterpret_cast<uint32_t*>(BUSY)[n0 >> 5] = bal;
                    if (b) a.busy_list[n_busy + popc32(bal & ((1u << lane) - 1))] = n;
                    n_busy += popc32(bal);
                }
            } else {
                int kept = 0;
                for (int e0 = 0; e0 < n_busy; e0 += 32) {
                    int e = e0 + lane;
                    int n = -1;
                    bool b = false;
                    if (e < n_busy) {
                        n = a.busy_list[e];
                        double bt = *reinterpret_cast<const double*>(a.nodes + chunk_off(n, 3));
                        b = (now - bt) < a.min_busy;
                        if (!b) atomicAnd(reinterpret_cast<unsigned long long*>(&BUSY[n >> 6]), ~(1ULL << (n & 63)));
                    }
                    uint32_t bal = __ballot_sync(0xFFFFFFFFu, b);
                    __syncwarp();
                    if (b) a.busy_list[kept + popc32(bal & ((1u << lane) - 1))] = n;   /* kept <= e0: no overlap hazard */
                    __syncwarp();
                    kept += popc32(bal);
                }
                n_busy = kept;
            }
            __syncwarp();
            cur_now = now;
        }

        /* ---- header of the binding ---- */
        if (lane == 0) {
            uint4 z = make_uint4(0, 0, 0, 0);
            for (int c = 0; c < 8; c++) reinterpret_cast<uint4*>(bout)[c] = z;
            bout->node = -1;
            bout->n_groups = t.G;
            bout->status = t.valid_map ? NHD_NO_CANDIDATE : NHD_BAD_MAP_TYPE;
        }
        if (!t.valid_map) continue;

        /* ---- first fit: pass 0 = GPU-less nodes for CPU-only pods (Matcher.py:412-416), pass 1 = any ---- */
        int chosen = -1;
        RecU u;
        Mapping m;
        for (int pass = t.needs_gpu ? 1 : 0; pass < 2 && chosen < 0; pass++) {
            uint64_t* F = a.bitmaps + (size_t)ti * W;
            int w0 = a.cursors[ti * 2 + pass];
            bool cursor_fixed = false;
            for (; w0 < W && chosen < 0; w0 += 32) {
                const int w = w0 + lane;
                uint64_t raw = 0;
                if (w < W) {
                    raw = F[w];
                    if (pass == 0) raw &= NOGPU[w];
                }
                if (!cursor_fixed) {
                    uint32_t nz = __ballot_sync(0xFFFFFFFFu, raw != 0);
                    if (nz) { cursor_fixed = true; if (lane == 0) a.cursors[ti * 2 + pass] = w0 + ctz32(nz); }
                    else if (lane == 0) a.cursors[ti * 2 + pass] = (w0 + 32 < W) ? w0 + 32 : W;
                }
                uint64_t word = raw;
                if (t.needs_gpu && w < W) word &= ~BUSY[w];              /* Matcher.py:107-111 */
                for (;;) {
                    uint32_t nz = __ballot_sync(0xFFFFFFFFu, word != 0);
                    if (!nz) break;
                    const int src = ctz32(nz);
                    const uint64_t cw = __shfl_sync(0xFFFFFFFFu, word, src);
                    const int node = (w0 + src) * 64 + ctz64(cw);
                    int ok = 0;
                    if (lane == 0) {
                        load_rec(a.nodes, node, u);
                        ok = node_gates(u.r, t) && evaluate_mapping_memo(a.memo, u.r, t, a.cap, &m);
                    }
                    ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
                    if (ok) { chosen = node; break; }
                    /* resources only shrink inside a batch: the node stays infeasible for this type */
                    if (lane == src) {
                        word &= ~(1ULL << (node & 63));
                        F[w] &= ~(1ULL << (node & 63));
                    }
                }
            }
        }
        if (chosen < 0) continue;

        /* ---- assignment + state update (lane 0) ---- */
        if (lane == 0) {
            nhd_binding b;
            {
                uint4 z = make_uint4(0, 0, 0, 0);
                for (int c = 0; c < 8; c++) reinterpret_cast<uint4*>(&b)[c] = z;
            }
            b.node = chosen;
            assign_pod(u.r, t, m, now, &b);
            store_rec_mutable(a.nodes, chosen, u);
            for (int c = 0; c < 8; c++) reinterpret_cast<uint4*>(bout)[c] = reinterpret_cast<uint4*>(&b)[c];
            if (a.min_busy > 0.0) {                       /* now - busy_time == 0 < MIN_BUSY_SECS */
                uint64_t bit = 1ULL << (chosen & 63);
                if (!(BUSY[chosen >> 6] & bit)) {
                    BUSY[chosen >> 6] |= bit;
                    a.busy_list[n_busy] = chosen;
                    n_busy++;
                }
            }
        }
        n_busy = __shfl_sync(0xFFFFFFFFu, n_busy, 0);
        __syncwarp();
    }
}

} // namespace nhd
