// plan.cu -- the generic fused scan -> filter -> project -> (hash) group -> aggregate operator: ONE pass over the columns for any
// plan of the shape the colexec pipeline runs per block (table_scan -> filter -> projection -> group, SURVEY.md section 8 rows
// a12-a19):
//
//   N conjunctive predicates   compare / BETWEEN of a column with constants          (filter.go:87-153, func_compare.go, operator_between.go)
//   an expression program      + - * / over columns and constants, SSA form          (evalExpression.go:575-640: one node per instruction)
//   group-by                   <= 8 key bytes packed like fillKeys                   (group/exec2.go:296-367, inthashmap.go:92-183)
//   K aggregates               SUM / AVG / COUNT / COUNT(*) / MIN / MAX              (aggexec/{sumavg2,count2,minmax2}.go)
//
// with a nulls bitmap on every input column.  tpch.cu holds the two hand-specialised instances of this operator (Q6, Q1); this file
// is the operator itself: a small warp-uniform interpreter.  Every thread owns one row per iteration; column values and instruction
// results live in a shared-memory register file vreg[slot][thread] (conflict-free: slot-major), so operand selection is one LDS
// instead of a chain of selects; null-ness of every slot is one bit of a per-thread mask.  Groups are found in a per-CTA shared
// memory hash table (first-come slots), whose partial states are folded into a global table when the CTA retires; rows whose CTA
// table is full go to the global table directly.
//
// Arithmetic is float64 (the reference's type for TPC-H style expressions over float columns): integer columns enter expressions
// converted to float64 (exact below 2^53); each instruction rounds once like the reference's separate nodes (no FMA contraction).
// Group sums are accumulated with atomics: the association order is not fixed (results agree with the serial loop to ~1e-13).
#include "common.cuh"
#include <cstring>

using namespace mob;

namespace {

constexpr int kThreads = 256;
constexpr int kCtaSlots = 256;          // per-CTA group table
constexpr uint64_t kEmptyKey = 0xffffffffffffffffull;

struct PlanCols { const uint8_t *data[MO_PLAN_MAX_COLS]; const uint64_t *nulls[MO_PLAN_MAX_COLS]; };

// global group table / result staging (device arena)
struct PlanGlobal {
    uint64_t *key; unsigned long long *first_row; unsigned long long *rows;   // [cap]
    double *acc; unsigned long long *cnt;                                      // [cap][naggs]
    uint64_t mask;                                                             // cap - 1
    unsigned *overflow;
};

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
__device__ __forceinline__ unsigned long long flt_key(double d) {
    if (d == 0.0) d = 0.0;
    unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | (1ull << 63));
}
__device__ __forceinline__ double flt_unkey(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & ~(1ull << 63)) : ~k;
    return __longlong_as_double((long long)b);
}
__device__ __forceinline__ double agg_identity(int kind) {
    if (kind == MO_AGG_MIN) return __longlong_as_double((long long)~0ull);   // ordered-key domain: all ones
    if (kind == MO_AGG_MAX) return __longlong_as_double(0ll);
    return 0.0;
}
// accumulate v into *a (shared or global)
__device__ __forceinline__ void agg_apply(int kind, double *a, double v) {
    if (kind == MO_AGG_MIN) { if (v == v) atomicMin(reinterpret_cast<unsigned long long *>(a), flt_key(v)); }
    else if (kind == MO_AGG_MAX) { if (v == v) atomicMax(reinterpret_cast<unsigned long long *>(a), flt_key(v)); }
    else if (kind == MO_AGG_SUM || kind == MO_AGG_AVG) atomicAdd(a, v);
}
// fold a partial state p into *a
__device__ __forceinline__ void agg_fold(int kind, double *a, double p) {
    if (kind == MO_AGG_MIN) atomicMin(reinterpret_cast<unsigned long long *>(a), (unsigned long long)__double_as_longlong(p));
    else if (kind == MO_AGG_MAX) atomicMax(reinterpret_cast<unsigned long long *>(a), (unsigned long long)__double_as_longlong(p));
    else if (kind == MO_AGG_SUM || kind == MO_AGG_AVG) atomicAdd(a, p);
}

__device__ __forceinline__ double load_as_f64(const uint8_t *p, int T, uint64_t r) {
    switch (T) {
    case MO_T_BOOL: case MO_T_UINT8: return (double)p[r];
    case MO_T_INT8: return (double)reinterpret_cast<const int8_t *>(p)[r];
    case MO_T_INT16: return (double)reinterpret_cast<const int16_t *>(p)[r];
    case MO_T_UINT16: return (double)reinterpret_cast<const uint16_t *>(p)[r];
    case MO_T_INT32: case MO_T_DATE: return (double)reinterpret_cast<const int32_t *>(p)[r];
    case MO_T_UINT32: return (double)reinterpret_cast<const uint32_t *>(p)[r];
    case MO_T_INT64: case MO_T_TIME: case MO_T_DATETIME: case MO_T_TIMESTAMP: return (double)reinterpret_cast<const int64_t *>(p)[r];
    case MO_T_UINT64: return (double)reinterpret_cast<const uint64_t *>(p)[r];
    case MO_T_FLOAT32: return (double)reinterpret_cast<const float *>(p)[r];
    default: return reinterpret_cast<const double *>(p)[r];
    }
}
__device__ __forceinline__ int type_bytes(int T) {
    switch (T) {
    case MO_T_BOOL: case MO_T_INT8: case MO_T_UINT8: return 1;
    case MO_T_INT16: case MO_T_UINT16: return 2;
    case MO_T_INT32: case MO_T_UINT32: case MO_T_FLOAT32: case MO_T_DATE: return 4;
    default: return 8;
    }
}

__device__ __forceinline__ uint64_t global_find(PlanGlobal &G, uint64_t key) {
    if (key == kEmptyKey) { G.key[G.mask + 1] = key; return G.mask + 1; }
    uint64_t s = mix64(key) & G.mask, probes = 0;
    for (;;) {
        uint64_t cur = G.key[s];
        if (cur == key) return s;
        if (cur == kEmptyKey) {
            cur = atomicCAS((unsigned long long *)&G.key[s], (unsigned long long)kEmptyKey, (unsigned long long)key);
            if (cur == kEmptyKey || cur == key) return s;
        }
        s = (s + 1) & G.mask;
        if (++probes > G.mask) { *G.overflow = 1; return ~0ull; }
    }
}

__global__ void plan_init_kernel(PlanGlobal G, int naggs, const mo_plan_t *P) {
    const uint64_t cap1 = G.mask + 2;
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s < cap1; s += (uint64_t)gridDim.x * blockDim.x) {
        G.key[s] = kEmptyKey; G.first_row[s] = ~0ull; G.rows[s] = 0;
        for (int a = 0; a < naggs; a++) { G.acc[s * naggs + a] = agg_identity(P->agg[a].kind); G.cnt[s * naggs + a] = 0; }
    }
}

__global__ void __launch_bounds__(kThreads)
plan_kernel(const mo_plan_t *__restrict__ Pg, PlanCols C, uint64_t n, PlanGlobal G) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ mo_plan_t P;
    for (int i = threadIdx.x; i < (int)(sizeof(mo_plan_t) / 4); i += kThreads) reinterpret_cast<uint32_t *>(&P)[i] = reinterpret_cast<const uint32_t *>(Pg)[i];
    __syncthreads();
    const int nslots = P.ncols + P.ninstr, naggs = P.naggs;
    double *vreg = reinterpret_cast<double *>(smem_raw);                                   // [nslots][kThreads]
    uint64_t *tkey = reinterpret_cast<uint64_t *>(vreg + (size_t)nslots * kThreads);       // [kCtaSlots]
    unsigned long long *tfirst = reinterpret_cast<unsigned long long *>(tkey + kCtaSlots);
    unsigned long long *trows = tfirst + kCtaSlots;
    double *tacc = reinterpret_cast<double *>(trows + kCtaSlots);                          // [kCtaSlots][naggs]
    unsigned long long *tcnt = reinterpret_cast<unsigned long long *>(tacc + (size_t)kCtaSlots * naggs);
    for (int s = threadIdx.x; s < kCtaSlots; s += kThreads) {
        tkey[s] = kEmptyKey; tfirst[s] = ~0ull; trows[s] = 0;
        for (int a = 0; a < naggs; a++) { tacc[s * naggs + a] = agg_identity(P.agg[a].kind); tcnt[s * naggs + a] = 0; }
    }
    __syncthreads();
    double *my = vreg + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * kThreads;
    for (uint64_t r = blockIdx.x * (uint64_t)kThreads + threadIdx.x; r < n; r += stride) {
        // ---- table scan: every referenced column value -> register-file slot c, null-ness -> bit c
        unsigned nullbits = 0;
        for (int c = 0; c < P.ncols; c++) {
            my[c * kThreads] = load_as_f64(C.data[c], P.col_type[c], r);
            if (C.nulls[c] && ((C.nulls[c][r >> 6] >> (r & 63)) & 1ull)) nullbits |= 1u << c;
        }
        // ---- filter: conjunction; a NULL operand makes the conjunct not-true (filter.go:125-141 keeps rows with !null && true)
        bool ok = true;
        for (int j = 0; j < P.npreds; j++) {
            const mo_plan_pred_t &q = P.pred[j];
            const double x = my[q.col * kThreads];
            bool t;
            switch (q.op) {
            case 0: t = x == q.lo; break; case 1: t = x != q.lo; break; case 2: t = x > q.lo; break;
            case 3: t = x >= q.lo; break; case 4: t = x < q.lo; break; case 5: t = x <= q.lo; break;
            default: t = x >= q.lo && x <= q.hi; break;   // BETWEEN, inclusive (operator_between.go:138-199)
            }
            ok = ok && t && !((nullbits >> q.col) & 1u);
        }
        if (!ok) continue;
        // ---- projection: SSA program, one rounding per node
        for (int i = 0; i < P.ninstr; i++) {
            const mo_plan_instr_t &in = P.instr[i];
            const int dst = P.ncols + i;
            double v; bool isnull = false;
            if (in.op == MO_PLAN_OP_COL) { v = my[in.a * kThreads]; isnull = (nullbits >> in.a) & 1u; }
            else if (in.op == MO_PLAN_OP_CONST) v = in.imm;
            else {
                const double a = my[in.a * kThreads], b = my[in.b * kThreads];
                isnull = ((nullbits >> in.a) | (nullbits >> in.b)) & 1u;
                switch (in.op) {
                case MO_PLAN_OP_ADD: v = __dadd_rn(a, b); break;
                case MO_PLAN_OP_SUB: v = __dsub_rn(a, b); break;
                case MO_PLAN_OP_MUL: v = __dmul_rn(a, b); break;
                default: if (b == 0.0) { isnull = true; v = 0.0; } else v = __ddiv_rn(a, b); break;   // x / 0 -> NULL (SELECT behaviour)
                }
            }
            my[dst * kThreads] = v;
            if (isnull) nullbits |= 1u << dst;
        }
        // ---- group key (fillKeys, has_null mode: marker byte per column; a NULL contributes the marker only)
        uint64_t key = 0;
        {
            int off = 0;
            for (int k = 0; k < P.nkeys; k++) {
                const int c = P.key_col[k];
                const int sz = type_bytes(P.col_type[c]);
                const bool isnull = (nullbits >> c) & 1u;
                if (P.has_null_keys) { if (isnull) { key |= 1ull << (8 * off); off += 1; continue; } off += 1; }
                uint64_t raw = 0;
                const uint8_t *p = C.data[c] + r * (uint64_t)sz;
                for (int b = 0; b < sz; b++) raw |= (uint64_t)p[b] << (8 * b);
                if (off < 8) key |= raw << (8 * off);
                off += sz;
            }
        }
        // ---- group slot: CTA table first, global table when it is full
        int slot = -1;
        if (key != kEmptyKey) {
            unsigned s = (unsigned)mix64(key) & (kCtaSlots - 1);
            for (int probes = 0; probes < kCtaSlots; probes++) {
                uint64_t cur = tkey[s];
                if (cur == key) { slot = (int)s; break; }
                if (cur == kEmptyKey) {
                    cur = atomicCAS((unsigned long long *)&tkey[s], (unsigned long long)kEmptyKey, (unsigned long long)key);
                    if (cur == kEmptyKey || cur == key) { slot = (int)s; break; }
                }
                s = (s + 1) & (kCtaSlots - 1);
            }
        }
        const uint64_t grow = (uint64_t)P.row_base + r;
        if (slot >= 0) {
            // warp pre-aggregation: the lanes of this warp that hit the same slot elect a leader, which adds the peers' values in lane
            // order and issues ONE atomic per aggregate -- with few groups (or none) the shared-memory atomics would otherwise serialise
            const unsigned peers = __match_any_sync(__activemask(), slot);
            const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
            const unsigned npeers = __popc(peers);
            if (tfirst[slot] > grow) atomicMin(&tfirst[slot], (unsigned long long)grow);
            if (lane == leader) atomicAdd(&trows[slot], (unsigned long long)npeers);
            for (int a = 0; a < naggs; a++) {
                const int vs = P.agg[a].value, kind = P.agg[a].kind;
                const bool has = vs < 0 || !((nullbits >> vs) & 1u);       // COUNT(*) counts every row; the others skip NULLs
                const double v = (vs >= 0 && has) ? my[vs * kThreads] : 0.0;
                if (kind == MO_AGG_MIN || kind == MO_AGG_MAX) {
                    if (has) { agg_apply(kind, &tacc[slot * naggs + a], v); atomicAdd(&tcnt[slot * naggs + a], 1ull); }
                    continue;
                }
                double sum = 0.0; unsigned cnt = 0;
                for (unsigned m = peers; m; m &= m - 1) {
                    const int src = __ffs(m) - 1;
                    const double pv = __shfl_sync(peers, v, src);
                    const unsigned pc = __shfl_sync(peers, has ? 1u : 0u, src);
                    if (pc) { sum = cnt ? __dadd_rn(sum, pv) : pv; cnt += 1; }
                }
                if (lane == leader && cnt) {
                    if (kind == MO_AGG_SUM || kind == MO_AGG_AVG) atomicAdd(&tacc[slot * naggs + a], sum);
                    atomicAdd(&tcnt[slot * naggs + a], (unsigned long long)cnt);
                }
            }
        } else {
            const uint64_t gs = global_find(G, key);
            if (gs == ~0ull) continue;
            atomicMin(&G.first_row[gs], (unsigned long long)grow);
            atomicAdd(&G.rows[gs], 1ull);
            for (int a = 0; a < naggs; a++) {
                const int vs = P.agg[a].value;
                if (vs < 0) { atomicAdd(&G.cnt[gs * naggs + a], 1ull); continue; }
                if ((nullbits >> vs) & 1u) continue;
                agg_apply(P.agg[a].kind, &G.acc[gs * naggs + a], my[vs * kThreads]);
                atomicAdd(&G.cnt[gs * naggs + a], 1ull);
            }
        }
    }
    // ---- retire: fold the CTA table into the global one
    __syncthreads();
    for (int s = threadIdx.x; s < kCtaSlots; s += kThreads) {
        if (tkey[s] == kEmptyKey || trows[s] == 0) continue;
        const uint64_t gs = global_find(G, tkey[s]);
        if (gs == ~0ull) continue;
        atomicMin(&G.first_row[gs], tfirst[s]);
        atomicAdd(&G.rows[gs], trows[s]);
        for (int a = 0; a < naggs; a++) {
            if (tcnt[s * naggs + a] == 0) continue;
            agg_fold(P.agg[a].kind, &G.acc[gs * naggs + a], tacc[s * naggs + a]);
            atomicAdd(&G.cnt[gs * naggs + a], tcnt[s * naggs + a]);
        }
    }
}

// collect the used slots, order them by first row (= the reference's group-id order) and write the result records
__global__ void plan_collect_kernel(PlanGlobal G, uint32_t *used, unsigned long long *nused) {
    for (uint64_t s = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; s <= G.mask + 1; s += (uint64_t)gridDim.x * blockDim.x)
        if (G.rows[s] != 0) { const unsigned long long i = atomicAdd(nused, 1ull); used[i] = (uint32_t)s; }
}
__global__ void plan_emit_kernel(PlanGlobal G, const mo_plan_t *P, const uint32_t *used, const unsigned long long *nused, uint8_t *res, uint64_t res_cap_groups,
                                 int sort_limit) {
    const unsigned long long ng = *nused;
    const int naggs = P->naggs;
    const size_t rec = sizeof(mo_plan_group_t) + sizeof(mo_plan_agg_value_t) * (size_t)naggs;
    mo_plan_result_header_t *H = reinterpret_cast<mo_plan_result_header_t *>(res);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        H->ngroups = (int64_t)ng; H->sorted = ng <= (unsigned long long)sort_limit ? 1 : 0; H->overflow = (*G.overflow || ng > res_cap_groups) ? 1 : 0; H->reserved = 0;
    }
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < ng; i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t s = used[i];
        unsigned long long pos = i;
        if (ng <= (unsigned long long)sort_limit) {   // rank by first row: the reference numbers groups in first-seen order
            pos = 0;
            const unsigned long long f = G.first_row[s];
            for (unsigned long long j = 0; j < ng; j++) { const unsigned long long fj = G.first_row[used[j]]; pos += (fj < f || (fj == f && j < i)) ? 1 : 0; }
        }
        if (pos >= res_cap_groups) continue;
        mo_plan_group_t *g = reinterpret_cast<mo_plan_group_t *>(res + sizeof(mo_plan_result_header_t) + rec * pos);
        g->key = G.key[s]; g->first_row = (int64_t)G.first_row[s]; g->rows = (int64_t)G.rows[s];
        mo_plan_agg_value_t *av = reinterpret_cast<mo_plan_agg_value_t *>(g + 1);
        for (int a = 0; a < naggs; a++) {
            const int kind = P->agg[a].kind;
            const unsigned long long c = G.cnt[(size_t)s * naggs + a];
            double v = G.acc[(size_t)s * naggs + a];
            if (kind == MO_AGG_MIN || kind == MO_AGG_MAX) v = c ? flt_unkey((unsigned long long)__double_as_longlong(v)) : 0.0;
            else if (kind == MO_AGG_AVG) v = c ? v / (double)c : 0.0;          // float64(sum) / float64(cnt), sumavg2.go:331
            else if (kind == MO_AGG_COUNT) v = (double)c;
            av[a].value = v; av[a].count = (int64_t)c;                          // count == 0: the aggregate is NULL (COUNT: 0)
        }
    }
}

}  // namespace

namespace mob {

// MO_XCALL_PLAN: args [0] result buffer ; [1] mo_plan_t (host) ; [2 .. 2 + ncols) the columns (+pnulls).  len = rows.  See include/mo_b200.h.
int xcall_plan(mo_xcall_args_t *args, uint64_t len) {
    ThreadCtx &t = tctx();
    if (!t.ready) return MO_RC_INTERNAL_ERROR;
    if (!args[1].pdata || args[1].dataSz < sizeof(mo_plan_t) || is_device_ptr(args[1].pdata)) { set_error("plan: host mo_plan_t descriptor missing"); return MO_RC_INVALID_ARGUMENT; }
    mo_plan_t P;
    memcpy(&P, args[1].pdata, sizeof P);
    if (P.ncols < 1 || P.ncols > MO_PLAN_MAX_COLS || P.npreds < 0 || P.npreds > MO_PLAN_MAX_PREDS || P.ninstr < 0 || P.ninstr > MO_PLAN_MAX_INSTR ||
        P.nkeys < 0 || P.nkeys > MO_PLAN_MAX_KEYS || P.naggs < 1 || P.naggs > MO_PLAN_MAX_AGGS) { set_error("plan: descriptor counts out of range"); return MO_RC_INVALID_ARGUMENT; }
    int keybytes = 0;
    auto tbytes = [](int T) { return (T == MO_T_BOOL || T == MO_T_INT8 || T == MO_T_UINT8) ? 1 : (T == MO_T_INT16 || T == MO_T_UINT16) ? 2 : (T == MO_T_INT32 || T == MO_T_UINT32 || T == MO_T_FLOAT32 || T == MO_T_DATE) ? 4 : 8; };
    for (int c = 0; c < P.ncols; c++) {
        const int T = P.col_type[c];
        if (!((T >= MO_T_INT8 && T <= MO_T_INT64) || (T >= MO_T_UINT8 && T <= MO_T_UINT64) || T == MO_T_BOOL || T == MO_T_FLOAT32 || T == MO_T_FLOAT64 || (T >= MO_T_DATE && T <= MO_T_TIMESTAMP))) { set_error("plan: column %d has unsupported type %d", c, T); return MO_RC_INVALID_ARGUMENT; }
        if (args[2 + c].dataSz < (uint64_t)tbytes(T) * len) { set_error("plan: column %d shorter than len", c); return MO_RC_INVALID_ARGUMENT; }
    }
    for (int j = 0; j < P.npreds; j++) if (P.pred[j].col < 0 || P.pred[j].col >= P.ncols || P.pred[j].op < 0 || P.pred[j].op > 6) { set_error("plan: predicate %d malformed", j); return MO_RC_INVALID_ARGUMENT; }
    for (int i = 0; i < P.ninstr; i++) {
        const mo_plan_instr_t &in = P.instr[i];
        const int lim = P.ncols + i;   // operands: columns or EARLIER instructions (SSA)
        const bool ok = in.op == MO_PLAN_OP_CONST || (in.op == MO_PLAN_OP_COL && in.a >= 0 && in.a < P.ncols) ||
                        (in.op >= MO_PLAN_OP_ADD && in.op <= MO_PLAN_OP_DIV && in.a >= 0 && in.a < lim && in.b >= 0 && in.b < lim);
        if (!ok) { set_error("plan: instruction %d malformed", i); return MO_RC_INVALID_ARGUMENT; }
    }
    for (int k = 0; k < P.nkeys; k++) {
        if (P.key_col[k] < 0 || P.key_col[k] >= P.ncols) { set_error("plan: key column %d out of range", k); return MO_RC_INVALID_ARGUMENT; }
        keybytes += tbytes(P.col_type[P.key_col[k]]) + (P.has_null_keys ? 1 : 0);
    }
    if (keybytes > 8) { set_error("plan: %d key bytes exceed the 8-byte IntHashMap key (group/exec2.go:73-118)", keybytes); return MO_RC_INVALID_ARGUMENT; }
    for (int a = 0; a < P.naggs; a++) {
        const mo_plan_agg_t &g = P.agg[a];
        if (g.kind < MO_AGG_SUM || g.kind > MO_AGG_AVG || g.value >= P.ncols + P.ninstr || (g.value < 0 && g.kind != MO_AGG_COUNT)) { set_error("plan: aggregate %d malformed", a); return MO_RC_INVALID_ARGUMENT; }
    }
    const size_t rec = sizeof(mo_plan_group_t) + sizeof(mo_plan_agg_value_t) * (size_t)P.naggs;
    if (!args[0].pdata || args[0].dataSz < sizeof(mo_plan_result_header_t) + rec) { set_error("plan: result buffer too small for one group"); return MO_RC_INVALID_ARGUMENT; }
    const uint64_t res_groups = (args[0].dataSz - sizeof(mo_plan_result_header_t)) / rec;
    const bool dev_res = is_device_ptr(args[0].pdata);
    bool dev_cols = true;
    for (int c = 0; c < P.ncols; c++) dev_cols = dev_cols && (len == 0 || is_device_ptr(args[2 + c].pdata)) && (!args[2 + c].pnulls || is_device_ptr(args[2 + c].pnulls));

    Stager st(t);
    PlanCols C;
    for (int c = 0; c < MO_PLAN_MAX_COLS; c++) { C.data[c] = nullptr; C.nulls[c] = nullptr; }
    for (int c = 0; c < P.ncols; c++) {
        C.data[c] = (const uint8_t *)st.in(args[2 + c].pdata, (size_t)tbytes(P.col_type[c]) * len);
        C.nulls[c] = (const uint64_t *)st.in(args[2 + c].pnulls, args[2 + c].pnulls ? ((len + 63) / 64) * 8 : 0);
    }
    uint8_t *dres = (uint8_t *)st.out(args[0].pdata, sizeof(mo_plan_result_header_t) + rec * res_groups);
    // global table: twice the groups the caller's buffer can take (min 1024 slots)
    uint64_t cap = 1024;
    while (cap < 2 * res_groups && cap < (1ull << 30)) cap <<= 1;
    PlanGlobal G;
    G.mask = cap - 1;
    G.key = (uint64_t *)st.tmp((cap + 1) * 8); G.first_row = (unsigned long long *)st.tmp((cap + 1) * 8); G.rows = (unsigned long long *)st.tmp((cap + 1) * 8);
    G.acc = (double *)st.tmp((cap + 1) * 8 * P.naggs); G.cnt = (unsigned long long *)st.tmp((cap + 1) * 8 * P.naggs);
    uint32_t *used = (uint32_t *)st.tmp((cap + 1) * 4);
    unsigned long long *nused = (unsigned long long *)st.tmp(16);
    mo_plan_t *dP = (mo_plan_t *)st.tmp(sizeof(mo_plan_t));
    if (st.failed) { st.finish(); return MO_RC_INTERNAL_ERROR; }
    G.overflow = (unsigned *)(nused + 1);
    MOB_CUDA_TRY(cudaMemsetAsync(nused, 0, 16, t.stream));
    // the descriptor travels through the thread's pinned staging buffer (the caller's copy may be pageable and short-lived)
    if (sizeof(mo_plan_t) > t.pinned_sz) { st.finish(); set_error("plan: descriptor larger than the staging buffer"); return MO_RC_INTERNAL_ERROR; }
    MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
    memcpy(t.pinned, &P, sizeof P);
    MOB_CUDA_TRY(cudaMemcpyAsync(dP, t.pinned, sizeof P, cudaMemcpyHostToDevice, t.stream));
    const unsigned igrid = (unsigned)((cap + 1 + 255) / 256 > (uint64_t)num_sms() * 8 ? (uint64_t)num_sms() * 8 : (cap + 1 + 255) / 256);
    plan_init_kernel<<<igrid, 256, 0, t.stream>>>(G, P.naggs, dP);
    MOB_LAUNCH_CHECK();
    const size_t smem = (size_t)(P.ncols + P.ninstr) * kThreads * 8 + (size_t)kCtaSlots * (24 + 16 * (size_t)P.naggs);
    static size_t attr_smem = 0;
    if (smem > attr_smem) { MOB_CUDA_TRY(cudaFuncSetAttribute(plan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr_smem = smem; }
    if (len) {
        int ctas = (int)((220 * 1024) / (smem + 2048));
        if (ctas > 6) ctas = 6; if (ctas < 1) ctas = 1;
        int grid = num_sms() * ctas;
        const uint64_t work = (len + kThreads - 1) / kThreads;
        if ((uint64_t)grid > work) grid = (int)work;
        cudaEventRecord(t.kev0, t.stream);
        plan_kernel<<<grid, kThreads, smem, t.stream>>>(dP, C, len, G);
        cudaEventRecord(t.kev1, t.stream);
        MOB_LAUNCH_CHECK();
    }
    plan_collect_kernel<<<igrid, 256, 0, t.stream>>>(G, used, nused);
    MOB_LAUNCH_CHECK();
    plan_emit_kernel<<<num_sms() * 2, 256, 0, t.stream>>>(G, dP, used, nused, dres, res_groups, 8192);
    MOB_LAUNCH_CHECK();
    if (dev_res && dev_cols) { st.release_async(); return MO_RC_SUCCESS; }   // asynchronous form: header.overflow reports a full table
    mo_plan_result_header_t H;
    int rc = read_back(t, &H, dres, sizeof H);
    int frc = st.finish();
    if (rc) return rc;
    if (H.overflow) { set_error("plan: more groups (%lld) than the result buffer holds (%llu)", (long long)H.ngroups, (unsigned long long)res_groups); return MO_RC_INVALID_ARGUMENT; }
    return frc;
}

}  // namespace mob
