/*
 * nhd_oracle.h — ORACLE entry points (test infrastructure only; see nhd_oracle.c).
 * Shares only the wire-record definitions with the product (include/nhd_b200.h).
 */
#ifndef NHD_ORACLE_H
#define NHD_ORACLE_H

#include "../include/nhd_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct nhd_oracle_params {
    double nic_bw_avail_percent;   /* nhd/Node.py:18 */
    double min_busy_secs;          /* nhd/Node.py:107 */
} nhd_oracle_params;

void nhd_oracle_default_params(nhd_oracle_params* p);

/* Schedule n_pods pods one after another (NHDScheduler.CheckPendingPods order,
 * nhd/NHDScheduler.py:425-441) on the cluster described by recs; recs is updated
 * in place; out[i] receives pod i's binding. Returns 0. */
int nhd_oracle_solve(const nhd_oracle_params* params, const double* speed_gbps,
                     int n_nodes, nhd_node_rec* recs,
                     int n_pods, const nhd_pod* pods, const double* now,
                     nhd_binding* out);

/* The same result with the walk over the nodes (which the reference repeats in full for every pod) split
 * over n_threads host threads; pods are still taken strictly one after another.  Used for the CPU
 * baseline of bench.py; checked bit for bit against nhd_oracle_solve (tests/test_oracle_mt.py). */
int nhd_oracle_solve_mt(const nhd_oracle_params* params, const double* speed_gbps,
                        int n_nodes, nhd_node_rec* recs,
                        int n_pods, const nhd_pod* pods, const double* now,
                        nhd_binding* out, int n_threads);

/* cand_out[n] = 1 iff node n is in filts[1] after IntersectResources for this pod
 * (nhd/Matcher.py:55) on the given (unmodified) state. */
int nhd_oracle_candidates(const nhd_oracle_params* params, const double* speed_gbps,
                          int n_nodes, const nhd_node_rec* recs,
                          const nhd_pod* pod, double now, uint8_t* cand_out);

#ifdef __cplusplus
}
#endif
#endif
