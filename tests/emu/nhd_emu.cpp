/*
 * nhd_emu.cpp — TEST INFRASTRUCTURE.  Compiles nhd_b200/csrc/nhd_core.cuh (the exact
 * scalar logic the CUDA kernels run) with g++ and drives it with a plain serial loop, so
 * that the device arithmetic can be diffed against the oracle on a machine without a GPU.
 * Never linked into the product library.
 */
#include <cstring>
#include <vector>
#include "../../nhd_b200/csrc/nhd_core.cuh"

using namespace nhd;

extern "C" {

int nhd_emu_solve(double bw, double min_busy, const double* speed, int n_nodes, nhd_node_rec* recs,
                  int n_pods, const nhd_pod* pods, const double* now, nhd_binding* out)
{
    double cap[NHD_MAX_SPEED_CLASSES];
    for (int i = 0; i < NHD_MAX_SPEED_CLASSES; i++) cap[i] = speed[i] * bw;
    for (int i = 0; i < n_pods; i++) {
        PodType t;
        make_pod_type(pods[i], t);
        nhd_binding* b = &out[i];
        std::memset(b, 0, sizeof(*b));
        b->node = -1;
        b->n_groups = t.G;
        if (!t.valid_map) { b->status = NHD_BAD_MAP_TYPE; continue; }
        int first = -1, first_nogpu = -1;
        for (int n = 0; n < n_nodes; n++) {
            if (!node_feasible(recs[n], t, cap)) continue;
            if (t.needs_gpu && node_busy(recs[n], now[i], min_busy)) continue;
            if (first < 0) first = n;
            if (first_nogpu < 0 && recs[n].n_gpus == 0) first_nogpu = n;
            if (first >= 0 && (t.needs_gpu || first_nogpu >= 0)) break;
        }
        if (first < 0) { b->status = NHD_NO_CANDIDATE; continue; }
        int node = (!t.needs_gpu && first_nogpu >= 0) ? first_nogpu : first;
        Mapping m;
        if (!evaluate_mapping(recs[node], t, cap, &m)) return -100 - i;   /* inconsistent with node_feasible */
        b->node = node;
        assign_pod(recs[node], t, m, now[i], b);
    }
    return 0;
}

int nhd_emu_feasible(double bw, const double* speed, int n_nodes, const nhd_node_rec* recs,
                     const nhd_pod* pod, uint8_t* out)
{
    double cap[NHD_MAX_SPEED_CLASSES];
    for (int i = 0; i < NHD_MAX_SPEED_CLASSES; i++) cap[i] = speed[i] * bw;
    PodType t;
    make_pod_type(*pod, t);
    for (int n = 0; n < n_nodes; n++) out[n] = node_feasible(recs[n], t, cap);
    return 0;
}

/* the register-only form of the predicate the filter kernel uses on 2-NUMA nodes for pods of <= 2 groups;
 * out[n] = 2 where that form does not apply */
int nhd_emu_feasible_k2(double bw, const double* speed, int n_nodes, const nhd_node_rec* recs,
                        const nhd_pod* pod, uint8_t* out)
{
    double cap[NHD_MAX_SPEED_CLASSES];
    for (int i = 0; i < NHD_MAX_SPEED_CLASSES; i++) cap[i] = speed[i] * bw;
    PodType t;
    make_pod_type(*pod, t);
    for (int n = 0; n < n_nodes; n++) {
        if (recs[n].n_numa != 2 || t.G < 1 || t.G > 2) { out[n] = 2; continue; }
        NodeDyn d;
        make_dyn(recs[n], d);
        Pre2 q;
        make_pre2(recs[n], d.fc[0], d.fc[1], q);
        out[n] = node_feasible_k2(recs[n], t, cap, q);
    }
    return 0;
}

/* stage_masks_k2 / nic_first_fit2 (the register-only forms resolve_kernel uses) against stage_masks_fc /
 * nic_first_fit on every applicable node: returns the number of differences, counts[0] = nodes compared,
 * counts[1] = NIC entries compared */
int nhd_emu_check_k2(double bw, const double* speed, int n_nodes, const nhd_node_rec* recs, const nhd_pod* pod, int* counts)
{
    double cap[NHD_MAX_SPEED_CLASSES];
    for (int i = 0; i < NHD_MAX_SPEED_CLASSES; i++) cap[i] = speed[i] * bw;
    PodType t;
    make_pod_type(*pod, t);
    int bad = 0;
    counts[0] = counts[1] = 0;
    if (t.G < 1 || t.G > 2) return 0;
    for (int n = 0; n < n_nodes; n++) {
        const nhd_node_rec& r = recs[n];
        if (r.n_numa != 2) continue;
        NodeDyn d;
        make_dyn(r, d);
        Pre2 q;
        make_pre2(r, d.fc[0], d.fc[1], q);
        const uint64_t gsw = t.pci ? free_gpus_per_switch(r) : 0;
        TMask ma, mb, mc;
        const bool ok = stage_masks_fc(r, d.fc, t, cap, gsw, ma, mb, mc);
        uint32_t a, b, c;
        const bool ok2 = stage_masks_k2(r, t, cap, q, a, b, c);
        counts[0]++;
        if (ok != ok2 || ma.w[0] != a || mb.w[0] != b || mc.w[0] != c) bad++;
        const int np = 1 << t.G;
        for (int pi = 0; pi < np; pi++) {
            uint8_t p[NHD_MAX_GROUPS], idx[NHD_MAX_GROUPS] = {0, 0, 0, 0}, li[NHD_MAX_GROUPS] = {0, 0, 0, 0};
            tuple_digits(pi, 2, t.G, p);
            const bool f = nic_first_fit(r, t, p, 2, cap, gsw, idx, li);
            uint32_t i2 = 0, l2 = 0;
            const bool f2 = nic_first_fit2(r, t, p[0], t.G > 1 ? p[1] : 0, cap, gsw, &i2, &l2);
            counts[1]++;
            if (f != f2) { bad++; continue; }
            if (!f) continue;
            for (int g = 0; g < t.G; g++)
                if (((i2 >> (8 * g)) & 0xFF) != idx[g] || ((l2 >> (8 * g)) & 0xFF) != li[g]) bad++;
        }
    }
    return bad;
}

int nhd_emu_choose(int K, int G, const uint64_t* a, const uint64_t* b, const uint64_t* c, int* p, int* m)
{
    TMask ma, mb, mc;
    for (int i = 0; i < 4; i++) { ma.w[i] = a[i]; mb.w[i] = b[i]; mc.w[i] = c[i]; }
    return choose_mapping(K, G, ma, mb, mc, p, m) ? 1 : 0;
}

uint64_t nhd_emu_tuple_hash(int idx, int K, int L) { return py_tuple_hash(idx, K, L); }

int nhd_emu_claim_order(const uint8_t* li, int n, uint8_t* out) { return claimed_nic_order(li, n, out); }

/* Two-stage flow exactly as the CUDA kernels run it: decisions on NodeDyn summaries, then the
 * core ids of every placed pod from the snapshot records and prefix offsets, then commit. */
int nhd_emu_solve2(double bw, double min_busy, const double* speed, int n_nodes, nhd_node_rec* recs,
                   int n_pods, const nhd_pod* pods, const double* now, nhd_binding* out)
{
    double cap[NHD_MAX_SPEED_CLASSES];
    for (int i = 0; i < NHD_MAX_SPEED_CLASSES; i++) cap[i] = speed[i] * bw;
    std::vector<NodeDyn> dyn(n_nodes);
    std::vector<nhd_node_rec> snap(recs, recs + n_nodes);
    for (int n = 0; n < n_nodes; n++) make_dyn(recs[n], dyn[n]);
    std::vector<PodType> types(n_pods);
    for (int i = 0; i < n_pods; i++) {
        PodType& t = types[i];
        make_pod_type(pods[i], t);
        nhd_binding* b = &out[i];
        std::memset(b, 0, sizeof(*b));
        b->node = -1;
        b->n_groups = t.G;
        if (!t.valid_map) { b->status = NHD_BAD_MAP_TYPE; continue; }
        b->status = NHD_NO_CANDIDATE;
        int first = -1, first_nogpu = -1;
        Mapping mf, mn;
        for (int n = 0; n < n_nodes; n++) {
            nhd_node_rec r = snap[n];
            apply_dyn(r, dyn[n]);
            if (!node_gates(r, t)) continue;
            if (t.needs_gpu && node_busy(r, now[i], min_busy)) continue;
            const uint64_t gsw = t.pci ? free_gpus_per_switch(r) : 0;
            TMask ma, mb, mc;
            const bool quick_no = summary_infeasible(t, dyn[n]);    /* must only ever fire on infeasible nodes */
            if (!stage_masks_fc(r, dyn[n].fc, t, cap, gsw, ma, mb, mc)) continue;
            int ps, ms;
            if (!choose_mapping(r.n_numa, t.G, ma, mb, mc, &ps, &ms)) continue;
            if (quick_no) return -7000 - i;
            Mapping m;
            tuple_digits(ps, r.n_numa, t.G, m.gpu_numa);
            m.misc_numa = (uint8_t)ms;
            nic_first_fit(r, t, m.gpu_numa, r.n_numa, cap, gsw, m.nic_idx, m.nic_li);
            if (first < 0) { first = n; mf = m; }
            if (first_nogpu < 0 && r.n_gpus == 0) { first_nogpu = n; mn = m; }
            if (t.needs_gpu || first_nogpu >= 0) break;
        }
        if (first < 0) continue;
        const bool use_nogpu = !t.needs_gpu && first_nogpu >= 0;
        const int node = use_nogpu ? first_nogpu : first;
        nhd_node_rec r = snap[node];
        apply_dyn(r, dyn[node]);
        b->node = node;
        assign_resources(r, dyn[node], t, use_nogpu ? mn : mf, now[i], b);
    }
    /* stage A: core ids, independent per pod */
    std::vector<M256> taken(n_pods);
    for (int i = 0; i < n_pods; i++)
        if (out[i].status == NHD_PLACED) taken[i] = assign_cores_from_snapshot(snap[out[i].node], types[i], &out[i]);
    /* commit */
    for (int i = 0; i < n_pods; i++) {
        if (out[i].node < 0) continue;
        nhd_node_rec& r = recs[out[i].node];
        if (out[i].status == NHD_PLACED)
            for (int w = 0; w < 4; w++) r.used[w] |= taken[i].w[w];
        apply_dyn(r, dyn[out[i].node]);
    }
    return 0;
}

}
