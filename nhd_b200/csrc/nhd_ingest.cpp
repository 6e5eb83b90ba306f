/*
 * nhd_ingest.cpp — node ingest: NFD label dictionary -> packed nhd_node_rec, host side, no CUDA.
 *
 * Restates what the reference does when a node appears (nhd/NHDScheduler.py:122-140):
 *   Node.ParseLabels   nhd/Node.py:468-487  = InitGroups :312-322, InitMaintenance :324-326 (+ GetMaintenance
 *                      :134-142), InitCores :328-376 (+ ParseRangeList :298-306), InitNics :378-426,
 *                      InitGpus :428-438, InitMisc :440-455
 *   Node.SetHugepages  nhd/Node.py:489-493
 * followed by what nhd_b200/packing.py:pack_node makes of the resulting Node object, so that a cluster can be
 * uploaded (nhd_load_nodes / nhd_update_nodes) straight from the labels the K8s API returns.
 *
 * Labels arrive as parallel key / value arrays IN DICTIONARY ORDER: the order of the NIC and GPU labels is the
 * order of Node.nics / Node.gpus, which the placement results depend on.
 *
 * Outcomes follow the reference:  NHD_OK;  NHD_ERR_LABELS where ParseLabels returns False (the scheduler
 * ignores the node);  NHD_ERR_INVALID where the reference would raise (malformed number, too few fields);
 * NHD_ERR_UNSUPPORTED for nodes that are fine for the reference but outside the packed layout's limits.
 */
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nhd_b200.h"

namespace {

constexpr const char* NFD = "feature.node.kubernetes.io/";
constexpr long SCHEDULABLE_NIC_SPEED_THRESH_MBPS = 11000;      /* nhd/Node.py:19 */

struct Err { int32_t code; };

/* Python int(s) / int(s, 16) on a str: optional surrounding whitespace, optional sign, digits with single
 * underscores between them (base 16: optional 0x prefix).  Anything else raises ValueError there. */
long py_int(const std::string& s, int base)
{
    size_t a = 0, b = s.size();
    auto is_ws = [](char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; };
    while (a < b && is_ws(s[a])) a++;
    while (b > a && is_ws(s[b - 1])) b--;
    bool neg = false;
    if (a < b && (s[a] == '+' || s[a] == '-')) { neg = s[a] == '-'; a++; }
    if (base == 16 && b - a >= 2 && s[a] == '0' && (s[a + 1] == 'x' || s[a + 1] == 'X')) {
        a += 2;
        if (a < b && s[a] == '_') a++;                          /* "0x_ff" is accepted */
    }
    if (a >= b) throw Err{NHD_ERR_INVALID};
    long v = 0;
    bool prev_us = true;                                        /* no leading underscore */
    for (size_t i = a; i < b; i++) {
        const char c = s[i];
        if (c == '_') { if (prev_us) throw Err{NHD_ERR_INVALID}; prev_us = true; continue; }
        int d;
        if (c >= '0' && c <= '9') d = c - '0';
        else if (base == 16 && c >= 'a' && c <= 'f') d = c - 'a' + 10;
        else if (base == 16 && c >= 'A' && c <= 'F') d = c - 'A' + 10;
        else throw Err{NHD_ERR_INVALID};
        if (v > (1L << 40)) throw Err{NHD_ERR_UNSUPPORTED};     /* far beyond anything a record can hold */
        v = v * base + d;
        prev_us = false;
    }
    if (prev_us) throw Err{NHD_ERR_INVALID};                    /* trailing underscore */
    return neg ? -v : v;
}

std::vector<std::string> split(const std::string& s, char sep)
{
    std::vector<std::string> out;
    size_t a = 0;
    for (;;) {
        const size_t p = s.find(sep, a);
        if (p == std::string::npos) { out.emplace_back(s.substr(a)); break; }
        out.emplace_back(s.substr(a, p - a));
        a = p + 1;
    }
    return out;
}

std::string lower(std::string s)
{
    for (auto& c : s) if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
    return s;
}

}  // namespace

struct nhd_ingest {
    std::vector<std::string> groups;      /* node-group name -> bit position, in order of first appearance */
    std::vector<double> speeds;           /* NIC speed (Gb/s) -> 4-bit class, in order of first appearance */

    int group_bit(const std::string& name, bool create)
    {
        for (size_t i = 0; i < groups.size(); i++) if (groups[i] == name) return (int)i;
        if (!create) return -1;
        if (groups.size() >= NHD_MAX_GROUP_NAMES) throw Err{NHD_ERR_UNSUPPORTED};
        groups.push_back(name);
        return (int)groups.size() - 1;
    }
    int speed_class(double gbps)
    {
        for (size_t i = 0; i < speeds.size(); i++) if (speeds[i] == gbps) return (int)i;
        if (speeds.size() >= NHD_MAX_SPEED_CLASSES) throw Err{NHD_ERR_UNSUPPORTED};
        speeds.push_back(gbps);
        return (int)speeds.size() - 1;
    }
};

extern "C" {

int32_t nhd_ingest_create(nhd_ingest** out)
{
    if (!out) return NHD_ERR_INVALID;
    *out = new (std::nothrow) nhd_ingest();
    return *out ? NHD_OK : NHD_ERR_INVALID;
}

int32_t nhd_ingest_destroy(nhd_ingest* g)
{
    delete g;
    return NHD_OK;
}

int32_t nhd_ingest_speed_table(const nhd_ingest* g, double out[NHD_MAX_SPEED_CLASSES], int32_t* n_classes)
{
    if (!g || !out) return NHD_ERR_INVALID;
    for (int i = 0; i < NHD_MAX_SPEED_CLASSES; i++) out[i] = i < (int)g->speeds.size() ? g->speeds[i] : 0.0;
    if (n_classes) *n_classes = (int32_t)g->speeds.size();
    return NHD_OK;
}

/* names: one '.'-joined list, the format of the NHD_GROUP label and of Node.SetGroups (Node.py:308-310) */
int32_t nhd_ingest_group_mask(nhd_ingest* g, const char* dotted_names, int32_t create, uint64_t* mask)
{
    if (!g || !dotted_names || !mask) return NHD_ERR_INVALID;
    try {
        uint64_t m = 0;
        for (const auto& name : split(dotted_names, '.')) {
            const int bit = g->group_bit(name, create != 0);
            if (bit >= 0) m |= 1ULL << bit;
        }
        *mask = m;
        return NHD_OK;
    } catch (const Err& e) {
        return e.code;
    }
}

int32_t nhd_ingest_node(nhd_ingest* g, int32_t n_labels, const char* const* keys, const char* const* values,
                        int32_t active, int32_t hugepages_alloc_gb, int32_t hugepages_free_gb,
                        nhd_node_rec* rec, nhd_node_aux* aux)
{
    (void)hugepages_alloc_gb;                                    /* Node.mem.ttl_hugepages_gb: never consulted by the matcher */
    if (!g || n_labels < 0 || (n_labels && (!keys || !values)) || !rec) return NHD_ERR_INVALID;
    for (int i = 0; i < n_labels; i++) if (!keys[i] || !values[i]) return NHD_ERR_INVALID;
    auto find = [&](const std::string& k) -> int {
        for (int i = 0; i < n_labels; i++) if (k == keys[i]) return i;
        return -1;
    };
    nhd_node_rec r;
    std::memset(&r, 0, sizeof r);
    nhd_node_aux ax;
    std::memset(&ax, 0, sizeof ax);
    ax.gw_label = -1;
    for (auto& v : ax.gpu_device_id) v = -1;
    for (auto& v : ax.nic_label) v = -1;
    try {
        /* ---- InitGroups (Node.py:312-322): names now, bits once the node is known to be usable (a rejected
         * node must not leave names behind in the cluster dictionary) ---- */
        std::vector<std::string> group_names;
        {
            const int i = find("NHD_GROUP");
            if (i < 0) group_names.emplace_back("default");
            else group_names = split(values[i], '.');
        }
        /* ---- InitMaintenance (Node.py:134-142, 324-326) ---- */
        bool maintenance = false;
        {
            const int i = find("sigproc.viasat.io/maintenance");
            if (i >= 0 && lower(values[i]) != "not_scheduled") maintenance = true;
        }
        /* ---- InitCores (Node.py:328-376) ---- */
        const int i_cores = find(std::string(NFD) + "nfd-extras-cpu.num_cores");
        const int i_sock = find(std::string(NFD) + "nfd-extras-cpu.numSockets");
        if (i_cores < 0 || i_sock < 0) return NHD_ERR_LABELS;
        const long sockets = py_int(values[i_sock], 10);
        const long cores = py_int(values[i_cores], 10);
        const bool smt = find(std::string(NFD) + "cpu-hardware_multithreading") >= 0;
        if (sockets == 0) return NHD_ERR_INVALID;               /* cores // sockets raises ZeroDivisionError */
        /* outside the packed limits (found by pack_node, i.e. only after ParseLabels has succeeded): remember,
         * keep parsing — a label problem further down takes precedence, as it does in the reference flow */
        const bool bad_cores = sockets < 1 || sockets > NHD_MAX_NUMA || cores <= 0 || cores % sockets != 0 ||
                               (smt ? 2 * cores : cores) > NHD_MAX_LCORES;
        const long n_logical = bad_cores ? 0 : (smt ? 2 * cores : cores);
        {
            const int i = find(std::string(NFD) + "nfd-extras-cpu.isolcpus");
            if (i >= 0) {
                std::vector<char> isol((size_t)n_logical, 0);
                for (const auto& rl : split(values[i], '_'))                /* underscores split the ranges (Node.py:356) */
                    for (const auto& part : split(rl, ',')) {               /* ParseRangeList (Node.py:298-306) */
                        const auto ends = split(part, '-');
                        const long lo = py_int(ends.front(), 10), hi = py_int(ends.back(), 10);
                        for (long c = lo < 0 ? 0 : lo; c <= hi && c < n_logical; c++) isol[(size_t)c] = 1;
                    }
                for (long c = 0; c < n_logical; c++)
                    if (!isol[(size_t)c]) { r.used[c >> 6] |= 1ULL << (c & 63); ax.n_reserved_cores++; }   /* OS cores (Node.py:368-371) */
            }
        }
        r.n_numa = (uint8_t)sockets;
        r.phys_cores = (uint16_t)cores;
        r.flags = (uint8_t)((smt ? NHD_NODE_SMT : 0) | (active ? NHD_NODE_ACTIVE : 0) | (maintenance ? NHD_NODE_MAINTENANCE : 0));

        /* ---- InitNics (Node.py:378-426): labels in dictionary order ---- */
        struct Dev { long numa, sw; double speed; int label; long device_id; };
        std::vector<Dev> nics, gpus;
        {
            std::vector<std::string> pfs;
            const std::string k_sriov = std::string(NFD) + "nfd-extras-sriov", k_nic = std::string(NFD) + "nfd-extras-nic";
            for (int i = 0; i < n_labels; i++)
                if (std::strstr(keys[i], k_sriov.c_str())) {
                    const auto p = split(keys[i], '.');
                    if (p.size() < 6) return NHD_ERR_INVALID;
                    pfs.push_back(p[5]);
                }
            for (int i = 0; i < n_labels; i++) {
                if (!std::strstr(keys[i], k_nic.c_str())) continue;
                const auto p = split(keys[i], '.');
                if (p.size() < 12) return NHD_ERR_INVALID;
                const std::string& ifname = p[4];
                const std::string& speed_s = p[7];
                const long numa = py_int(p[8], 10), sw = py_int(p[9], 16);
                (void)py_int(p[10], 16);                          /* card */
                (void)py_int(p[11], 10);                          /* port */
                bool is_pf = false;
                for (const auto& pf : pfs) if (pf == ifname) is_pf = true;
                if (is_pf) continue;
                const size_t at = speed_s.find("Mbs");
                if (at == std::string::npos) continue;            /* interface down */
                const long mbps = py_int(speed_s.substr(0, at), 10);
                if (mbps < SCHEDULABLE_NIC_SPEED_THRESH_MBPS) continue;
                nics.push_back(Dev{numa, sw, (double)mbps / 1e3, i, 0});
            }
        }
        /* ---- InitGpus (Node.py:428-438) ---- */
        {
            const std::string k_gpu = std::string(NFD) + "nfd-extras-gpu";
            for (int i = 0; i < n_labels; i++) {
                if (!std::strstr(keys[i], k_gpu.c_str())) continue;
                const auto p = split(keys[i], '.');
                if (p.size() < 8) return NHD_ERR_INVALID;
                const long dev = py_int(p[4], 10), numa = py_int(p[6], 10), sw = py_int(p[7], 16);
                gpus.push_back(Dev{numa, sw, 0.0, i, dev});
            }
        }
        /* ---- InitMisc (Node.py:440-455) ---- */
        {
            const int iv = find("DATA_PLANE_VLAN");
            if (iv < 0) return NHD_ERR_LABELS;
            ax.data_vlan = (int32_t)py_int(values[iv], 10);
            const int ig = find("DATA_DEFAULT_GW");
            if (ig < 0) return NHD_ERR_LABELS;
            ax.gw_label = ig;
            const int ir = find("RES_HUGEPAGES_GB");
            if (ir >= 0) ax.res_hugepages_gb = (int32_t)py_int(values[ir], 10);
        }
        if (bad_cores) return NHD_ERR_UNSUPPORTED;
        /* ---- SetHugepages (Node.py:489-493) ---- */
        r.free_hugepages_gb = hugepages_free_gb - ax.res_hugepages_gb;
        r.busy_time = 0.0;

        /* ---- packing (nhd_b200/packing.py:pack_node): group bits, then local switch ids in order of first use,
         * GPUs first ---- */
        for (const auto& name : group_names) r.group_mask |= 1ULL << g->group_bit(name, true);
        std::vector<long> switches;
        auto local_switch = [&](long sw) -> uint64_t {
            for (size_t i = 0; i < switches.size(); i++) if (switches[i] == sw) return i;
            if (switches.size() >= NHD_MAX_SWITCHES) throw Err{NHD_ERR_UNSUPPORTED};
            switches.push_back(sw);
            return switches.size() - 1;
        };
        if (gpus.size() > NHD_MAX_GPUS) return NHD_ERR_UNSUPPORTED;
        for (size_t i = 0; i < gpus.size(); i++) {
            if (gpus[i].numa < 0 || gpus[i].numa >= sockets) return NHD_ERR_UNSUPPORTED;
            /* the reference's failure unwind frees self.gpus[device_id] (Node.py:828): exact only while the device
             * ids are the list positions (same rule as packing.pack_node) */
            if (gpus[i].device_id != (long)i) return NHD_ERR_UNSUPPORTED;
            r.gpu_numa_mask[gpus[i].numa] |= (uint16_t)(1u << i);
            r.gpu_sw |= local_switch(gpus[i].sw) << (4 * i);
            ax.gpu_device_id[i] = (int32_t)gpus[i].device_id;
        }
        r.n_gpus = (uint8_t)gpus.size();
        if (nics.size() > NHD_MAX_NICS) return NHD_ERR_UNSUPPORTED;
        for (size_t i = 0; i < nics.size(); i++) {
            if (nics[i].numa < 0 || nics[i].numa >= sockets) return NHD_ERR_UNSUPPORTED;
            r.nic_numa_mask[nics[i].numa] |= 1u << i;
            r.nic_sw[i >> 4] |= local_switch(nics[i].sw) << (4 * (i & 15));
            r.nic_speed[i >> 4] |= (uint64_t)g->speed_class(nics[i].speed) << (4 * (i & 15));
            ax.nic_label[i] = nics[i].label;
        }
        r.n_nics = (uint8_t)nics.size();
    } catch (const Err& e) {
        return e.code;
    } catch (...) {
        return NHD_ERR_INVALID;
    }
    *rec = r;
    if (aux) *aux = ax;
    return NHD_OK;
}

/* ---- statistics off packed records (NHDScheduler.GetBasicNodeStats, nhd/NHDScheduler.py:355-378) ---- */
int32_t nhd_node_stats_from_records(int32_t n, const nhd_node_rec* recs, nhd_node_stats* out)
{
    if (n < 0 || (n && (!recs || !out))) return NHD_ERR_INVALID;
    for (int32_t i = 0; i < n; i++) {
        const nhd_node_rec& r = recs[i];
        nhd_node_stats s;
        std::memset(&s, 0, sizeof s);
        const bool smt = (r.flags & NHD_NODE_SMT) != 0;
        const int phys = r.phys_cores, K = r.n_numa;
        if (K < 1 || K > NHD_MAX_NUMA || phys <= 0 || phys % K || (smt ? 2 : 1) * phys > NHD_MAX_LCORES) return NHD_ERR_INVALID;
        const int per = phys / K;
        auto used = [&](int c) { return (r.used[c >> 6] >> (c & 63)) & 1ULL; };
        for (int c = 0; c < phys; c++) {
            /* Node.py:250-264: a physical core counts when it and (with SMT) its sibling are unused */
            const bool free_pair = !used(c) && (!smt || !used(c + phys));
            if (free_pair) s.free_cores_numa[c / per]++;
            /* Node.py:225-233: every logical core whose own and sibling's flags are clear */
            if (smt) { if (free_pair) s.freecpu += 2; }
            else if (!used(c)) s.freecpu += 1;
        }
        s.totalcpu = smt ? 2 * phys : phys;
        s.totalgpu = r.n_gpus;
        for (int g = 0; g < r.n_gpus; g++)
            if (!((r.gpu_used >> g) & 1)) {
                s.freegpu++;
                for (int k = 0; k < K; k++) if ((r.gpu_numa_mask[k] >> g) & 1) s.free_gpus_numa[k]++;
            }
        s.freehuge_gb = r.free_hugepages_gb;
        s.active = (r.flags & NHD_NODE_ACTIVE) ? 1 : 0;
        s.maintenance = (r.flags & NHD_NODE_MAINTENANCE) ? 1 : 0;
        s.nics_in_use = __builtin_popcount(r.nic_inuse & (r.n_nics >= 32 ? 0xFFFFFFFFu : ((1u << r.n_nics) - 1)));
        out[i] = s;
    }
    return NHD_OK;
}

}  // extern "C"
