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

int nhd_emu_choose(int K, int G, const uint64_t* a, const uint64_t* b, const uint64_t* c, int* p, int* m)
{
    TMask ma, mb, mc;
    for (int i = 0; i < 4; i++) { ma.w[i] = a[i]; mb.w[i] = b[i]; mc.w[i] = c[i]; }
    return choose_mapping(K, G, ma, mb, mc, p, m) ? 1 : 0;
}

uint64_t nhd_emu_tuple_hash(int idx, int K, int L) { return py_tuple_hash(idx, K, L); }

int nhd_emu_claim_order(const uint8_t* li, int n, uint8_t* out) { return claimed_nic_order(li, n, out); }

}
