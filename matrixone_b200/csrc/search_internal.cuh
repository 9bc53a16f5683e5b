// search_internal.cuh -- declarations shared by search.cu (exact kernels) and tcsearch.cu (tensor-core candidate pass)
#pragma once
#include "common.cuh"
#include <vector>

namespace mob {

// IVF probe plan: which lists every query scans, inverted into per-list buckets of (query, slot) pairs
struct IvfPlan {
    int nprobe = 0;
    int64_t npairs = 0;            // nq * nprobe
    int64_t nvalid = 0;            // pairs that landed in a bucket (npairs minus missing / empty lists)
    int64_t *probes = nullptr;     // [nq][nprobe] list ids by ascending centroid distance (-1 = none)       (device)
    int32_t *bucket_q = nullptr;   // [npairs] query of every bucket entry, entries of one list are contiguous  (device)
    int32_t *bucket_l = nullptr;   // [npairs] list of every bucket entry                                        (device)
    int64_t *bucket_slot = nullptr;// [npairs] q * nprobe + rank of every bucket entry                          (device)
    int32_t *pair_pos = nullptr;   // [npairs] inverse map: position of pair (q, rank) in the buckets, -1 = none (device)
    std::vector<int> hcnt, hstart; // per list: bucket size / first position                                    (host)
};

int bruteforce_topk_device(ThreadCtx &t, const float *ddata, int64_t n, int dim, const float *dq, int64_t nq, int k, int metric,
                           int64_t key_base, int sqrt_out, int64_t *out_k, double *out_d);
int ivf_make_plan(ThreadCtx &t, const float *dcent, int64_t nlist, int dim, const float *dq, int64_t nq, int nprobe, int metric, IvfPlan &plan,
                  const int64_t *doffsets = nullptr);   // doffsets: drop the pairs of lists that are empty here (list-sharded index)
int ivf_exact_scan(ThreadCtx &t, const IvfPlan &plan, const float *ddata, int64_t n, int dim, const float *dq, int64_t nq,
                   const std::vector<int64_t> &offsets, const int64_t *drowids, int k, int metric, int sqrt_out, int64_t *ok, double *od);

// held (shared) by every search entry point for the duration of the call; see g_search_rw in tcsearch.cu
struct SearchReadGuard { SearchReadGuard(); ~SearchReadGuard(); SearchReadGuard(const SearchReadGuard &) = delete; SearchReadGuard &operator=(const SearchReadGuard &) = delete; };

bool tc_search_applicable(int64_t n, int dim, int64_t nq, int k, int metric);
bool tc_probe_applicable(int64_t nlist, int dim, int64_t nq, int nprobe, int metric);
// record_stats = false: a helper search inside another call (IVF centroid probe) leaves the fallback counter and kernel timer alone
int bruteforce_topk_tc_device(ThreadCtx &t, const float *ddata, int64_t n, int dim, const float *dq, int64_t nq, int k,
                              int64_t key_base, int sqrt_out, int64_t *out_k, double *out_d, bool record_stats = true, int metric = MO_METRIC_L2SQ);
bool tc_ivf_applicable(int64_t n, int dim, int64_t nq, int k, int nprobe, int metric, bool refine);
// returns the queries whose completeness proof failed in `redo` (their rows of ok/od are still filled with best-effort results).
// pass 0 = one-term product, 1 = three-term, 2 = three-term over list sub-ranges.  *nonfinite: Inf/NaN input, the error bound
// does not apply, redo = all.
int ivf_tc_scan(ThreadCtx &t, const IvfPlan &plan, const float *ddata, int64_t n, int dim, const float *dq, int64_t nq,
                const float *dcent, const int64_t *doffsets, const std::vector<int64_t> &offsets, const int64_t *drowids, int k, int sqrt_out,
                int pass, int64_t *ok, double *od, std::vector<int> &redo, bool *nonfinite);
bool tc_one_term_wanted(int k, bool record);   // ladder policy shared by brute force and IVF
void tc_one_term_report(int64_t nq, int64_t failed);

extern thread_local int g_last_tc_fallbacks;   // queries of the last tensor-core search that ended in the exact kernel
extern thread_local int g_last_tc_refined;     // IVF: queries of the last search that needed the sub-range refine pass

}  // namespace mob
