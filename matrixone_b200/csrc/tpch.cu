// tpch.cu -- fused scan -> filter -> project -> aggregate kernels for the two OLAP shapes BASELINE.json names.
//
// The reference runs these as an operator chain that materialises a vector per expression node and compacts every
// column after each conjunct (table_scan -> filter -> projection -> group: pkg/sql/colexec/filter/filter.go:87-153,
// pkg/sql/colexec/group/exec2.go:296-367, aggexec/sumavg2.go:133-199).  Here each shape is ONE pass over the
// columns: every byte is read from HBM exactly once, nothing is written but a few hundred bytes of partials.
//
//   Q6  SUM(price*disc) WHERE date in [lo,hi) AND disc BETWEEN a AND b AND qty < c     28 B/row   (q6.sql:50-61)
//   Q1  filter date <= cutoff, group by (returnflag, linestatus), 8 aggregates          38 B/row packed keys,
//                                                                                        84 B/row varlena keys (q1.sql:1-21)
//
// Mapping: a "pair" is two consecutive rows.  Lane l of a warp owns pair (w*32 + l), so one warp instruction reads
// 64 consecutive rows: 256 B of the int32 column (LDG.64/lane) and 512 B of each float64 column (LDG.128/lane) --
// whole 32-byte sectors, nothing fetched twice.  kUnroll pairs are in flight per thread; grid = 148 x CTAs/SM
// persistent CTAs, grid-stride.  Products are rounded exactly as the reference's separate multiply / add nodes do
// (__dmul_rn/__dadd_rn, no FMA contraction), so only the summation ORDER differs from the serial Go loop.
// Reduction: registers -> warp shuffles -> shared memory -> per-CTA record; the last CTA (atomicInc ticket) folds
// the records in CTA-index order => bitwise run-to-run deterministic for a fixed grid.
#include "common.cuh"
#include <cstring>
#include <atomic>

using namespace mob;

namespace {

constexpr int kThreads = 256;

// =========================================================================================================
// Q6
// =========================================================================================================
struct Q6Rec { double sum; unsigned long long cnt; };
// caller-owned device result (asynchronous form): sum at res[0], qualifying rows at res[1] when res_words >= 2, nulls word bit 0 =
// "no row qualified" -- the layout MO_XCALL_Q6_MERGE consumes
struct Q6DevOut { double *res; uint64_t res_words; uint64_t *rnulls; };

template <int UNROLL, int CTAS>
__global__ void __launch_bounds__(kThreads, CTAS)
q6_kernel(const int32_t *__restrict__ sd, const double *__restrict__ disc, const double *__restrict__ qty,
          const double *__restrict__ price, uint64_t n, mo_q6_params_t P, Q6Rec *__restrict__ partials,
          Q6Rec *__restrict__ out, unsigned *ticket, Q6DevOut dres) {
    const uint64_t npairs = n >> 1;
    const uint64_t tid = blockIdx.x * (uint64_t)kThreads + threadIdx.x;
    const uint64_t nthreads = (uint64_t)gridDim.x * kThreads;
    double s0 = 0.0, s1 = 0.0;
    unsigned cnt = 0;  // a thread sees < 2^32 rows

    auto row = [&](int32_t d, double di, double q, double pr, double &s) {
        bool ok = (d >= P.date_lo) & (d < P.date_hi) & (di >= P.disc_lo) & (di <= P.disc_hi) & (q < P.qty_hi);
        double prod = __dmul_rn(pr, di);             // projection node l_extendedprice * l_discount
        s = __dadd_rn(s, ok ? prod : 0.0);           // SUM node; +0.0 leaves s unchanged (s is never -0.0)
        cnt += ok;
    };

    uint64_t p = tid;
    for (; p + (UNROLL - 1) * nthreads < npairs; p += UNROLL * nthreads) {
        int2 d[UNROLL]; int4 a[UNROLL], b[UNROLL], c[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            const uint64_t q = p + k * nthreads;
            d[k] = ld_stream8(sd + 2 * q);
            a[k] = ld_stream16(disc + 2 * q);
            b[k] = ld_stream16(qty + 2 * q);
            c[k] = ld_stream16(price + 2 * q);
        }
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            double di[2], qq[2], pr[2];
            memcpy(di, &a[k], 16); memcpy(qq, &b[k], 16); memcpy(pr, &c[k], 16);
            row(d[k].x, di[0], qq[0], pr[0], s0);
            row(d[k].y, di[1], qq[1], pr[1], s1);
        }
    }
    for (; p < npairs; p += nthreads) {
        int2 d = ld_stream8(sd + 2 * p);
        int4 a = ld_stream16(disc + 2 * p), b = ld_stream16(qty + 2 * p), c = ld_stream16(price + 2 * p);
        double di[2], qq[2], pr[2];
        memcpy(di, &a, 16); memcpy(qq, &b, 16); memcpy(pr, &c, 16);
        row(d.x, di[0], qq[0], pr[0], s0);
        row(d.y, di[1], qq[1], pr[1], s1);
    }
    if ((n & 1) && tid == 0) row(sd[n - 1], disc[n - 1], qty[n - 1], price[n - 1], s0);  // odd tail row

    // ---- deterministic reduction
    double s = __dadd_rn(s0, s1);
    unsigned long long c64 = cnt;
    s = warp_sum_f64(s);
    c64 = warp_sum(c64);
    __shared__ double ss[kThreads / 32];
    __shared__ unsigned long long sc[kThreads / 32];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { ss[warp] = s; sc[warp] = c64; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0; unsigned long long tc = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++) { t = __dadd_rn(t, ss[w]); tc += sc[w]; }
        partials[blockIdx.x].sum = t; partials[blockIdx.x].cnt = tc;
        __threadfence();
        unsigned tk = atomicInc(ticket, gridDim.x - 1);
        last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        __threadfence();
        double t = 0.0; unsigned long long tc = 0;
        for (unsigned bIdx = threadIdx.x; bIdx < gridDim.x; bIdx += kThreads) { t = __dadd_rn(t, partials[bIdx].sum); tc += partials[bIdx].cnt; }
        t = warp_sum_f64(t); tc = warp_sum(tc);
        __syncthreads();
        if (lane == 0) { ss[warp] = t; sc[warp] = tc; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double f = 0.0; unsigned long long fc = 0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; w++) { f = __dadd_rn(f, ss[w]); fc += sc[w]; }
            out->sum = f; out->cnt = fc;
            if (dres.res) {
                dres.res[0] = f;
                if (dres.res_words >= 2) reinterpret_cast<long long *>(dres.res)[1] = (long long)fc;
                if (dres.rnulls) dres.rnulls[0] = fc ? 0ull : 1ull;
            }
        }
    }
}

// =========================================================================================================
// Q1
// =========================================================================================================
constexpr int kQ1Vals = 5;  // sum_qty, sum_price, sum_disc_price, sum_charge, sum_disc
struct Q1Slot { unsigned long long first_row; unsigned long long cnt; double v[kQ1Vals]; unsigned key; unsigned used; };
struct Q1Rec { Q1Slot slot[MO_Q1_MAX_GROUPS]; unsigned overflow; unsigned pad; };
constexpr unsigned kEmptyKey = 0xffffffffu;
constexpr int kQ1MaxGrid = 1024;  // upper bound on the grid (the last-CTA fold keeps a [grid][8] byte map in shared memory)

// Shared state of one CTA: key -> slot dictionary (slots handed out first-come) + flags
struct Q1Shared {
    unsigned dict[MO_Q1_MAX_GROUPS];
    unsigned overflow;
    unsigned slow;   // debug: dictionary slow-path entries
};

// Shared-memory accumulators (staged kernels, G = 4): the per-thread state acc[slot][5], cnt[slot], first[slot] lives in shared memory,
// slot-major ([(slot * 5 + j) * threads + tid]: conflict-free whatever slot each lane uses), so a row costs 5 x (LDS, DADD, STS) on ITS
// slot instead of 20 indicator-FMAs over all four: ncu put the register version at 2.87 G warp instructions per SF100 pass with the
// issue slots 70 % busy -- instruction-bound, not HBM-bound.  Branch-free like the register version: rows that do not qualify add 0.0 to
// slot 0.  acc + v is the same single rounding as before, so results are bit-identical.
template <int kT>
struct Q1SmemAcc {
    double *acc; unsigned *cnt; unsigned *first;   // bases already offset by tid
    static constexpr size_t kBytes = (size_t)kT * 4 * (kQ1Vals * 8 + 4 + 4);
    __device__ __forceinline__ void bind(unsigned char *base, int tid) {
        acc = reinterpret_cast<double *>(base) + tid;
        cnt = reinterpret_cast<unsigned *>(base + (size_t)kT * 4 * kQ1Vals * 8) + tid;
        first = cnt + (size_t)kT * 4;
#pragma unroll
        for (int g = 0; g < 4; g++) {
#pragma unroll
            for (int j = 0; j < kQ1Vals; j++) acc[(g * kQ1Vals + j) * kT] = 0.0;
            cnt[g * kT] = 0u; first[g * kT] = 0xffffffffu;
        }
    }
    __device__ __forceinline__ void add(int slot, bool m, uint64_t r, double q, double pr, double t2, double t4, double di) {
        const int s = m ? slot : 0;
        double *a = acc + (size_t)s * kQ1Vals * kT;
        a[0] = __dadd_rn(a[0], m ? q : 0.0);
        a[kT] = __dadd_rn(a[kT], m ? pr : 0.0);
        a[2 * kT] = __dadd_rn(a[2 * kT], m ? t2 : 0.0);
        a[3 * kT] = __dadd_rn(a[3 * kT], m ? t4 : 0.0);
        a[4 * kT] = __dadd_rn(a[4 * kT], m ? di : 0.0);
        const unsigned c = cnt[s * kT];
        cnt[s * kT] = c + (m ? 1u : 0u);
        const unsigned f = first[s * kT];
        first[s * kT] = (m && c == 0u) ? (unsigned)r : f;     // rows of one launch are numbered below 2^32 (checked by the launcher)
    }
};

// Per-thread aggregation state (registers).  Control flow of every method is WARP-UNIFORM: all 32 lanes call row() together
// and the dictionary slow path is taken by the whole warp (ballot + one elected lane doing the CAS).  A first version let
// lanes diverge inside the insert loop; some warps then never reconverged and ran ~8x slower to the end of the kernel.
template <int G>
struct Q1Thread {
    double acc[G][kQ1Vals];
    unsigned cnt[G];
    unsigned long long first[G];
    unsigned dk[G];  // register copy of the dictionary
    Q1Shared *S;
    int32_t cutoff;
    int lane;
    bool dbg;

    __device__ __forceinline__ void init(Q1Shared *s, int32_t cut, bool debug) {
        S = s; cutoff = cut; lane = threadIdx.x & 31; dbg = debug;
#pragma unroll
        for (int g = 0; g < G; g++) {
            cnt[g] = 0; first[g] = ~0ull; dk[g] = kEmptyKey;
#pragma unroll
            for (int j = 0; j < kQ1Vals; j++) acc[g][j] = 0.0;
        }
    }

    __device__ __forceinline__ int find_slot(unsigned key, bool valid) {
        int slot = -1;
#pragma unroll
        for (int g = 0; g < G; g++) if (dk[g] == key) slot = g;
        bool settled = !valid || slot >= 0;   // a lane is settled once its key has been looked up in the shared dictionary
        unsigned miss = __ballot_sync(0xffffffffu, !settled);
        while (miss) {
            const int leader = __ffs(miss) - 1;
            const unsigned lkey = __shfl_sync(0xffffffffu, key, leader);
            // Claim or find lkey in the shared dictionary.  EVERY lane issues the same idempotent CAS sequence (same address,
            // same value): an elected-lane `if (lane == leader) { loop }` left the warp split in two groups for the rest of
            // the kernel (ncu: 16.0 active threads per instruction, profiles/r01_ncu_summary.md), doubling the issue cost.
            bool ok = false;
#pragma unroll
            for (int g = 0; g < G; g++) {
                const unsigned prev = ok ? lkey : atomicCAS(&S->dict[g], kEmptyKey, lkey);
                ok = ok || prev == kEmptyKey || prev == lkey;
            }
            if (!ok) S->overflow = 1;
            __syncwarp();
#pragma unroll
            for (int g = 0; g < G; g++) dk[g] = ((volatile unsigned *)S->dict)[g];
            slot = -1;
#pragma unroll
            for (int g = 0; g < G; g++) if (dk[g] == key) slot = g;
            if (key == lkey || slot >= 0) settled = true;   // lkey lanes are settled either way (slot found, or dictionary full)
            miss = __ballot_sync(0xffffffffu, !settled);
        }
        return slot;
    }

    // shared-memory accumulator form of row(): same dictionary, same products, 15 accumulate instructions instead of ~44
    template <int kT>
    __device__ __forceinline__ void row_smem(Q1SmemAcc<kT> &A, uint64_t r, bool inb, int32_t d, double q, double pr, double di, double tx, unsigned key) {
        const bool valid = inb && d <= cutoff;
        const int slot = find_slot(key, valid);
        const double t1 = __dsub_rn(1.0, di);
        const double t2 = __dmul_rn(pr, t1);
        const double t3 = __dadd_rn(1.0, tx);
        const double t4 = __dmul_rn(t2, t3);
        A.add(slot, valid && slot >= 0, r, q, pr, t2, t4, di);
    }
    template <int kT>
    __device__ __forceinline__ void load_from(const Q1SmemAcc<kT> &A) {   // registers <- shared state, for the common epilogue (G >= 4)
#pragma unroll
        for (int g = 0; g < 4; g++) {
#pragma unroll
            for (int j = 0; j < kQ1Vals; j++) acc[g][j] = A.acc[(g * kQ1Vals + j) * kT];
            cnt[g] = A.cnt[g * kT];
            const unsigned f = A.first[g * kT];
            first[g] = f == 0xffffffffu ? ~0ull : (unsigned long long)f;
        }
    }

    // `inb` = the row exists; the shipdate filter is folded into `valid`.
    // Group dispatch without branches: acc[g][j] = fma(v_j, ind_g, acc[g][j]) with ind_g = 1.0 for the row's slot and 0.0 for the
    // others.  fma(v, 1.0, a) == RN(v + a) (bit-identical to the reference's add) and fma(v, 0.0, a) == a for finite v, so all 32
    // lanes stay active.  Rows holding Inf/NaN would poison the other groups through 0 * Inf, so a warp that sees one takes the
    // exact predicated path.
    __device__ __forceinline__ void row(uint64_t r, bool inb, int32_t d, double q, double pr, double di, double tx, unsigned key) {
        const bool valid = inb && d <= cutoff;        // l_shipdate <= cutoff
        const int slot = find_slot(key, valid);
        const double t1 = __dsub_rn(1.0, di);         // 1 - l_discount
        const double t2 = __dmul_rn(pr, t1);          // l_extendedprice * (1 - l_discount)
        const double t3 = __dadd_rn(1.0, tx);         // 1 + l_tax
        const double t4 = __dmul_rn(t2, t3);          // ... * (1 + l_tax)
        const bool finite = (fabs(t4) + fabs(q)) < INFINITY;   // false for any Inf/NaN among q, pr, di, tx (and on overflow)
        if (__any_sync(0xffffffffu, valid && !finite)) {
#pragma unroll
            for (int g = 0; g < G; g++) {   // selects, not branches: lanes must not diverge here either
                const bool m = valid && slot == g;
                acc[g][0] = m ? __dadd_rn(acc[g][0], q) : acc[g][0];
                acc[g][1] = m ? __dadd_rn(acc[g][1], pr) : acc[g][1];
                acc[g][2] = m ? __dadd_rn(acc[g][2], t2) : acc[g][2];
                acc[g][3] = m ? __dadd_rn(acc[g][3], t4) : acc[g][3];
                acc[g][4] = m ? __dadd_rn(acc[g][4], di) : acc[g][4];
                if (m && cnt[g] == 0) first[g] = r;
                cnt[g] += m;
            }
            return;
        }
#pragma unroll
        for (int g = 0; g < G; g++) {
            const bool m = valid && slot == g;
            const double ind = __hiloint2double(m ? 0x3ff00000 : 0, 0);   // 1.0 or 0.0
            acc[g][0] = __fma_rn(q, ind, acc[g][0]);
            acc[g][1] = __fma_rn(pr, ind, acc[g][1]);
            acc[g][2] = __fma_rn(t2, ind, acc[g][2]);
            acc[g][3] = __fma_rn(t4, ind, acc[g][3]);
            acc[g][4] = __fma_rn(di, ind, acc[g][4]);
            if (m && cnt[g] == 0) first[g] = r;       // a thread visits its rows in increasing order
            cnt[g] += m;
        }
    }
};

// key of one row: packed uint8 columns (KEYMODE 0) or MatrixOne varlena cells (KEYMODE 1: 24 B, inline: bs[0]=len,
// bs[1..]=bytes; cgo/xcall.h:33-61); an empty string keys as 0
template <int KEYMODE>
__device__ __forceinline__ unsigned q1_key_scalar(const uint8_t *rf, const uint8_t *ls, uint64_t r) {
    if (KEYMODE == 0) return rf[r] | ((unsigned)ls[r] << 8);
    return (rf[24 * r] ? rf[24 * r + 1] : 0u) | ((ls[24 * r] ? (unsigned)ls[24 * r + 1] : 0u) << 8);
}

// CTA reduction + last-CTA fold, shared by both kernels.  kT = threads per CTA.
template <int G, int kT>
__device__ void q1_epilogue(Q1Thread<G> &T, Q1Shared &S, Q1Rec *__restrict__ partials, Q1Rec *__restrict__ out, unsigned *ticket,
                            unsigned long long *dbg) {
    auto stamp = [&](int k) { if (dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[blockIdx.x * 16 + k] = t; } };
    stamp(1);
    if (dbg && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) < 8) {
        unsigned long long tt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tt));
        unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
        dbg[blockIdx.x * 16 + 8 + (threadIdx.x >> 5)] = tt;
        if (threadIdx.x == 0) dbg[blockIdx.x * 16 + 5] = sm;
    }
    __syncthreads();  // dictionary final
    stamp(2);
    if (dbg && threadIdx.x == 0) dbg[blockIdx.x * 16 + 7] = S.slow;
    // Register slots are indexed by the CTA dictionary, identical for every thread of the CTA.
    constexpr int kW = kT / 32;
    __shared__ double sv[kW][G][kQ1Vals];
    __shared__ unsigned long long scnt[kW][G], sfirst[kW][G];
    __shared__ bool last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int g = 0; g < G; g++) {
        unsigned long long c64 = T.cnt[g], f = T.first[g];
        c64 = warp_sum(c64);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { unsigned long long of = __shfl_xor_sync(0xffffffffu, f, o); f = of < f ? of : f; }
#pragma unroll
        for (int j = 0; j < kQ1Vals; j++) {
            double v = warp_sum_f64(T.acc[g][j]);
            if (lane == 0) sv[warp][g][j] = v;
        }
        if (lane == 0) { scnt[warp][g] = c64; sfirst[warp][g] = f; }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        Q1Rec &R = partials[blockIdx.x];
        for (int g = 0; g < G; g++) {
            Q1Slot s; s.cnt = 0; s.first_row = ~0ull; s.key = S.dict[g]; s.used = S.dict[g] != kEmptyKey;
            for (int j = 0; j < kQ1Vals; j++) s.v[j] = 0.0;
            for (int w = 0; w < kW; w++) {
                s.cnt += scnt[w][g];
                if (sfirst[w][g] < s.first_row) s.first_row = sfirst[w][g];
                for (int j = 0; j < kQ1Vals; j++) s.v[j] = __dadd_rn(s.v[j], sv[w][g][j]);
            }
            R.slot[g] = s;
        }
        for (int g = G; g < MO_Q1_MAX_GROUPS; g++) { Q1Slot z; memset(&z, 0, sizeof z); z.key = kEmptyKey; R.slot[g] = z; }
        R.overflow = S.overflow; R.pad = 0;
        __threadfence();
        unsigned tk = atomicInc(ticket, gridDim.x - 1);
        last = (tk == gridDim.x - 1);
    }
    stamp(3);
    __syncthreads();
    if (last) {
        // Fold the CTA records (MergeGroup: re-hash partial keys + BatchMerge, mergeGroup.go:132-247) with the whole CTA:
        //  1. one thread per CTA record maps its local slots to final slots (shared dictionary, atomicCAS);
        //  2. one thread per (final slot, value) sums that value over the CTA records IN CTA-INDEX ORDER, so the result
        //     is bitwise deterministic although the dictionary slot order is not (the host sorts groups by first_row).
        __threadfence();
        __shared__ unsigned fdict[MO_Q1_MAX_GROUPS];
        __shared__ unsigned char inv[kQ1MaxGrid][MO_Q1_MAX_GROUPS];   // [cta][final slot] -> local slot or 0xff
        __shared__ unsigned f_overflow;
        if (threadIdx.x < MO_Q1_MAX_GROUPS) fdict[threadIdx.x] = kEmptyKey;
        if (threadIdx.x == 0) f_overflow = 0;
        for (unsigned i = threadIdx.x; i < gridDim.x * MO_Q1_MAX_GROUPS; i += kT) inv[i / MO_Q1_MAX_GROUPS][i % MO_Q1_MAX_GROUPS] = 0xff;
        __syncthreads();
        for (unsigned bIdx = threadIdx.x; bIdx < gridDim.x; bIdx += kT) {
            const Q1Rec &R = partials[bIdx];
            if (R.overflow) f_overflow = 1;
#pragma unroll
            for (int g = 0; g < G; g++) {
                if (!R.slot[g].used || R.slot[g].cnt == 0) continue;
                const unsigned key = R.slot[g].key;
                int dst = -1;
                for (int x = 0; x < MO_Q1_MAX_GROUPS; x++) {
                    const unsigned prev = atomicCAS(&fdict[x], kEmptyKey, key);
                    if (prev == kEmptyKey || prev == key) { dst = x; break; }
                }
                if (dst < 0) f_overflow = 1; else inv[bIdx][dst] = (unsigned char)g;
            }
        }
        __syncthreads();
        if (threadIdx.x < MO_Q1_MAX_GROUPS * 8) {
            const int sl = threadIdx.x >> 3, j = threadIdx.x & 7;
            Q1Slot &D = out->slot[sl];
            if (j < kQ1Vals) {
                double a = 0.0;
                for (unsigned bIdx = 0; bIdx < gridDim.x; bIdx++) { const unsigned g = inv[bIdx][sl]; if (g != 0xff) a = __dadd_rn(a, partials[bIdx].slot[g].v[j]); }
                D.v[j] = a;
            } else if (j == 5) {
                unsigned long long c = 0;
                for (unsigned bIdx = 0; bIdx < gridDim.x; bIdx++) { const unsigned g = inv[bIdx][sl]; if (g != 0xff) c += partials[bIdx].slot[g].cnt; }
                D.cnt = c;
            } else if (j == 6) {
                unsigned long long f = ~0ull;
                for (unsigned bIdx = 0; bIdx < gridDim.x; bIdx++) { const unsigned g = inv[bIdx][sl]; if (g != 0xff) { const unsigned long long r = partials[bIdx].slot[g].first_row; f = r < f ? r : f; } }
                D.first_row = f;
            } else {
                D.key = fdict[sl]; D.used = fdict[sl] != kEmptyKey;
                if (sl == 0) { out->overflow = f_overflow; out->pad = 0; }
            }
        }
        __syncthreads();
        stamp(4);
    }
}

// ---- variant A: direct register-staged loads (pairs of rows, UNROLL pairs in flight per thread) -------------------------
template <int G, int UNROLL, int CTAS, int KEYMODE>
__global__ void __launch_bounds__(kThreads, CTAS)
q1_kernel(const int32_t *__restrict__ sd, const double *__restrict__ qty, const double *__restrict__ price,
          const double *__restrict__ disc, const double *__restrict__ tax, const uint8_t *__restrict__ rf,
          const uint8_t *__restrict__ ls, uint64_t n, int32_t cutoff, Q1Rec *__restrict__ partials,
          Q1Rec *__restrict__ out, unsigned *ticket, unsigned long long *dbg, const Q1Rec *gate) {
    // asynchronous form: the 8-slot retry is enqueued behind the 4-slot pass and runs only if that pass overflowed its dictionary
    if (gate && !gate->overflow) return;
    if (dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[blockIdx.x * 16] = t; }
    __shared__ Q1Shared S;
    S.dict[threadIdx.x & (MO_Q1_MAX_GROUPS - 1)] = kEmptyKey;   // all threads write (same values): no lane-dependent branch
    S.overflow = 0; S.slow = 0;
    __syncthreads();
    Q1Thread<G> T;
    T.init(&S, cutoff, dbg != nullptr);
    const int lane = threadIdx.x & 31;

    auto load_keys = [&](uint64_t pair, unsigned &k0, unsigned &k1) {
        if (KEYMODE == 0) {
            unsigned short a = __ldg(reinterpret_cast<const unsigned short *>(rf) + pair);
            unsigned short b = __ldg(reinterpret_cast<const unsigned short *>(ls) + pair);
            k0 = (a & 0xffu) | ((b & 0xffu) << 8);
            k1 = (a >> 8) | ((b >> 8) << 8);
        } else {
            const uint64_t r0 = 2 * pair;   // head 8 bytes of each 24-byte cell hold len + first chars
            uint2 a0 = __ldg(reinterpret_cast<const uint2 *>(rf + 24 * r0)), a1 = __ldg(reinterpret_cast<const uint2 *>(rf + 24 * (r0 + 1)));
            uint2 b0 = __ldg(reinterpret_cast<const uint2 *>(ls + 24 * r0)), b1 = __ldg(reinterpret_cast<const uint2 *>(ls + 24 * (r0 + 1)));
            auto ch = [](uint2 h) -> unsigned { return (h.x & 0xffu) ? ((h.x >> 8) & 0xffu) : 0u; };
            k0 = ch(a0) | (ch(b0) << 8);
            k1 = ch(a1) | (ch(b1) << 8);
        }
    };

    const uint64_t npairs = n >> 1;
    const uint64_t tid = blockIdx.x * (uint64_t)kThreads + threadIdx.x;
    const uint64_t nthreads = (uint64_t)gridDim.x * kThreads;
    const uint64_t wbase = tid - lane;   // first pair of this warp: loop bounds depend on it only => uniform per warp
    uint64_t pw = wbase;
    for (; pw + 31 + (UNROLL - 1) * nthreads < npairs; pw += UNROLL * nthreads) {   // all 32 lanes x UNROLL pairs in range
        int2 d[UNROLL]; int4 a[UNROLL], b[UNROLL], c[UNROLL], e[UNROLL]; unsigned k0[UNROLL], k1[UNROLL];
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            const uint64_t q = pw + lane + k * nthreads;
            d[k] = ld_stream8(sd + 2 * q);
            a[k] = ld_stream16(qty + 2 * q);
            b[k] = ld_stream16(price + 2 * q);
            c[k] = ld_stream16(disc + 2 * q);
            e[k] = ld_stream16(tax + 2 * q);
            load_keys(q, k0[k], k1[k]);
        }
#pragma unroll
        for (int k = 0; k < UNROLL; k++) {
            const uint64_t q = pw + lane + k * nthreads;
            double qq[2], pr[2], di[2], tx[2];
            memcpy(qq, &a[k], 16); memcpy(pr, &b[k], 16); memcpy(di, &c[k], 16); memcpy(tx, &e[k], 16);
            T.row(2 * q, true, d[k].x, qq[0], pr[0], di[0], tx[0], k0[k]);
            T.row(2 * q + 1, true, d[k].y, qq[1], pr[1], di[1], tx[1], k1[k]);
        }
    }
    for (; pw < npairs; pw += nthreads) {   // ragged end: same code, lanes past the end carry inb = false
        const uint64_t q = pw + lane;
        const bool inb = q < npairs;
        int2 d = make_int2(0, 0); int4 a = make_int4(0, 0, 0, 0), b = a, c = a, e = a; unsigned k0 = 0, k1 = 0;
        if (inb) {
            d = ld_stream8(sd + 2 * q);
            a = ld_stream16(qty + 2 * q); b = ld_stream16(price + 2 * q); c = ld_stream16(disc + 2 * q); e = ld_stream16(tax + 2 * q);
            load_keys(q, k0, k1);
        }
        double qq[2], pr[2], di[2], tx[2];
        memcpy(qq, &a, 16); memcpy(pr, &b, 16); memcpy(di, &c, 16); memcpy(tx, &e, 16);
        T.row(2 * q, inb, d.x, qq[0], pr[0], di[0], tx[0], k0);
        T.row(2 * q + 1, inb, d.y, qq[1], pr[1], di[1], tx[1], k1);
    }
    if (tid < 32) {   // odd tail row: handled by warp 0 of CTA 0, lane 0 carries it
        const bool has = (n & 1) && lane == 0;
        const uint64_t r = n - 1;
        unsigned key = 0; int32_t dd = 0; double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
        if (has) { key = q1_key_scalar<KEYMODE>(rf, ls, r); dd = sd[r]; v0 = qty[r]; v1 = price[r]; v2 = disc[r]; v3 = tax[r]; }
        T.row(r, has, dd, v0, v1, v2, v3, key);
    }
    q1_epilogue<G, kThreads>(T, S, partials, out, ticket, dbg);
}

// ---- variant B: cp.async-staged tiles (packed keys only) ----------------------------------------------------------------
// Each warp owns a ring of kStages tiles of 64 rows in shared memory.  The 64-row slices of the seven columns are copied with
// per-lane cp.async (LDGSTS): 8 B/lane for the int32 column, 16 B/lane for each float64 column, 4 B/lane (half a warp) for each
// key column -- every copy instruction covers whole 32-byte sectors.  Bytes in flight per SM = warps x kStages x 2432 B,
// independent of the register file, so the loads of tile i+kStages-1 are in the air while tile i is reduced: the register
// version above could keep only ~78 KB per SM in flight and stalled on HBM latency (profiles/r01_ncu_summary.md).
// No CTA barriers in the steady state: a warp waits for ITS oldest copy group (cp.async.wait_group) and __syncwarp()s.
constexpr int kTileRows = 64;
constexpr int kTileBytes = kTileRows * 38;   // 256 (shipdate) + 4 x 512 + 2 x 64 = 2432
constexpr int kStagedThreads = 512;

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int G, int kStages, bool SMEMACC>
__global__ void __launch_bounds__(kStagedThreads, 1)
q1_staged_kernel(const int32_t *__restrict__ sd, const double *__restrict__ qty, const double *__restrict__ price,
                 const double *__restrict__ disc, const double *__restrict__ tax, const uint8_t *__restrict__ rf,
                 const uint8_t *__restrict__ ls, uint64_t n, int32_t cutoff, Q1Rec *__restrict__ partials,
                 Q1Rec *__restrict__ out, unsigned *ticket, unsigned long long *dbg) {
    extern __shared__ __align__(16) unsigned char ring[];   // [warp][stage][kTileBytes]
    if (dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[blockIdx.x * 16] = t; }
    __shared__ Q1Shared S;
    S.dict[threadIdx.x & (MO_Q1_MAX_GROUPS - 1)] = kEmptyKey;   // all threads write (same values): no lane-dependent branch
    S.overflow = 0; S.slow = 0;
    __syncthreads();
    Q1Thread<G> T;
    T.init(&S, cutoff, dbg != nullptr);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int kWarps = kStagedThreads / 32;
    unsigned char *wring = ring + (size_t)warp * kStages * kTileBytes;
    Q1SmemAcc<kStagedThreads> A;
    if (SMEMACC) A.bind(ring + (size_t)kWarps * kStages * kTileBytes, threadIdx.x);
    const uint64_t ntiles = n / kTileRows;
    const uint64_t gw = blockIdx.x * (uint64_t)kWarps + warp, nw = (uint64_t)gridDim.x * kWarps;

    auto issue = [&](uint64_t tile, int stage) {   // all lanes; one commit group per tile
        unsigned char *s = wring + stage * kTileBytes;
        const uint64_t r0 = tile * kTileRows;
        cp_async8(s + lane * 8, sd + r0 + lane * 2);
        cp_async16(s + 256 + lane * 16, qty + r0 + lane * 2);
        cp_async16(s + 768 + lane * 16, price + r0 + lane * 2);
        cp_async16(s + 1280 + lane * 16, disc + r0 + lane * 2);
        cp_async16(s + 1792 + lane * 16, tax + r0 + lane * 2);
        // keys: lanes 0-15 copy the returnflag slice, lanes 16-31 the linestatus slice (address select, no branch); the
        // linestatus area follows the returnflag area so one destination expression serves both
        cp_async4(s + 2304 + lane * 4, (lane < 16 ? rf + r0 + lane * 4 : ls + r0 + (lane - 16) * 4));
        cp_async_commit();
    };

    // prologue: kStages - 1 tiles in flight (empty groups keep the group arithmetic uniform when tiles run out)
    uint64_t next = gw;
#pragma unroll
    for (int s = 0; s < kStages - 1; s++) {
        if (next < ntiles) issue(next, s); else cp_async_commit();
        next += nw;
    }
    int stage = 0;
    for (uint64_t tile = gw; tile < ntiles; tile += nw) {
        // refill the slot consumed in the previous iteration, then wait for the oldest group
        const int fill = (stage + kStages - 1) % kStages;
        if (next < ntiles) issue(next, fill); else cp_async_commit();
        next += nw;
        cp_async_wait<kStages - 1>();
        __syncwarp();
        const unsigned char *s = wring + stage * kTileBytes;
        const uint64_t r0 = tile * kTileRows;
#pragma unroll
        for (int h = 0; h < 2; h++) {   // lane handles rows lane and lane + 32 of the tile: conflict-free LDS
            const int rl = lane + 32 * h;
            const int32_t d = *reinterpret_cast<const int32_t *>(s + rl * 4);
            const double q = *reinterpret_cast<const double *>(s + 256 + rl * 8);
            const double pr = *reinterpret_cast<const double *>(s + 768 + rl * 8);
            const double di = *reinterpret_cast<const double *>(s + 1280 + rl * 8);
            const double tx = *reinterpret_cast<const double *>(s + 1792 + rl * 8);
            const unsigned key = (unsigned)s[2304 + rl] | ((unsigned)s[2368 + rl] << 8);
            if (SMEMACC) T.row_smem(A, r0 + rl, true, d, q, pr, di, tx, key);
            else T.row(r0 + rl, true, d, q, pr, di, tx, key);
        }
        __syncwarp();   // every lane is done with this slot before any lane's next cp.async overwrites it
        stage = (stage + 1) % kStages;
    }
    cp_async_wait<0>();
    // rows past the last full tile (< 64): warp 0 of CTA 0, direct loads
    if (blockIdx.x == 0 && warp == 0) {
        for (uint64_t r = ntiles * kTileRows + lane; r - lane < n; r += 32) {
            const bool has = r < n;
            unsigned key = 0; int32_t dd = 0; double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
            if (has) { key = q1_key_scalar<0>(rf, ls, r); dd = sd[r]; v0 = qty[r]; v1 = price[r]; v2 = disc[r]; v3 = tax[r]; }
            if (SMEMACC) T.row_smem(A, r, has, dd, v0, v1, v2, v3, key);
            else T.row(r, has, dd, v0, v1, v2, v3, key);
        }
    }
    if (SMEMACC) T.load_from(A);
    q1_epilogue<G, kStagedThreads>(T, S, partials, out, ticket, dbg);
}

// ---- variant C: bulk-copy (TMA 1-D) staged tiles ---------------------------------------------------------------------------------
// Same per-warp ring as variant B, but every column slice of a tile is ONE cp.async.bulk (256 B shipdate, 4 x 512 B float64 columns,
// 2 x 64 B key bytes) issued by lane 0 and completed on a per-(warp, stage) mbarrier: 7 copy instructions per 64-row tile instead of
// 6 x 32 per-lane LDGSTS, nothing passes through L1, and the L2 -> SM crossbar moves every byte once (ncu on variant B:
// l1tex__m_xbar2l1tex_read_bytes = 1.84 x the DRAM bytes, sector hit rate 49 %).
__device__ __forceinline__ void q1_mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void q1_mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void q1_mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nQ1_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra Q1_DONE;\nbra Q1_WAIT;\nQ1_DONE:\n}"
        ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void q1_bulk_copy(void *smem, const void *gmem, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int G, int kStages>
__global__ void __launch_bounds__(kStagedThreads, 1)
q1_bulk_kernel(const int32_t *__restrict__ sd, const double *__restrict__ qty, const double *__restrict__ price,
               const double *__restrict__ disc, const double *__restrict__ tax, const uint8_t *__restrict__ rf,
               const uint8_t *__restrict__ ls, uint64_t n, int32_t cutoff, Q1Rec *__restrict__ partials,
               Q1Rec *__restrict__ out, unsigned *ticket, unsigned long long *dbg) {
    extern __shared__ __align__(16) unsigned char ring[];   // [warp][stage][kTileBytes], then the mbarriers
    if (dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[blockIdx.x * 16] = t; }
    __shared__ Q1Shared S;
    S.dict[threadIdx.x & (MO_Q1_MAX_GROUPS - 1)] = kEmptyKey;
    S.overflow = 0; S.slow = 0;
    constexpr int kWarps = kStagedThreads / 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *wring = ring + (size_t)warp * kStages * kTileBytes;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(ring + (size_t)kWarps * kStages * kTileBytes) + warp * kStages;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages; s++) q1_mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    Q1Thread<G> T;
    T.init(&S, cutoff, dbg != nullptr);
    const uint64_t ntiles = n / kTileRows;
    const uint64_t gw = blockIdx.x * (uint64_t)kWarps + warp, nw = (uint64_t)gridDim.x * kWarps;

    auto issue = [&](uint64_t tile, int stage) {   // lane 0 only
        unsigned char *s = wring + stage * kTileBytes;
        const uint64_t r0 = tile * kTileRows;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the slot was last READ through the generic proxy
        q1_mbar_expect_tx(&bars[stage], kTileBytes);
        q1_bulk_copy(s, sd + r0, 256, &bars[stage]);
        q1_bulk_copy(s + 256, qty + r0, 512, &bars[stage]);
        q1_bulk_copy(s + 768, price + r0, 512, &bars[stage]);
        q1_bulk_copy(s + 1280, disc + r0, 512, &bars[stage]);
        q1_bulk_copy(s + 1792, tax + r0, 512, &bars[stage]);
        q1_bulk_copy(s + 2304, rf + r0, 64, &bars[stage]);
        q1_bulk_copy(s + 2368, ls + r0, 64, &bars[stage]);
    };

    uint64_t next = gw;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < kStages - 1; s++) { if (next < ntiles) issue(next, s); next += nw; }
    } else next += (uint64_t)(kStages - 1) * nw;
    int stage = 0; unsigned phase = 0;
    for (uint64_t tile = gw; tile < ntiles; tile += nw) {
        const int fill = (stage + kStages - 1) % kStages;
        if (lane == 0 && next < ntiles) issue(next, fill);   // the slot consumed in the previous iteration (all lanes passed its __syncwarp)
        next += nw;
        q1_mbar_wait(&bars[stage], phase);
        const unsigned char *s = wring + stage * kTileBytes;
        const uint64_t r0 = tile * kTileRows;
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int rl = lane + 32 * h;
            const int32_t d = *reinterpret_cast<const int32_t *>(s + rl * 4);
            const double q = *reinterpret_cast<const double *>(s + 256 + rl * 8);
            const double pr = *reinterpret_cast<const double *>(s + 768 + rl * 8);
            const double di = *reinterpret_cast<const double *>(s + 1280 + rl * 8);
            const double tx = *reinterpret_cast<const double *>(s + 1792 + rl * 8);
            const unsigned key = (unsigned)s[2304 + rl] | ((unsigned)s[2368 + rl] << 8);
            T.row(r0 + rl, true, d, q, pr, di, tx, key);
        }
        __syncwarp();
        stage = (stage + 1) % kStages;
        if (stage == 0) phase ^= 1u;
    }
    if (blockIdx.x == 0 && warp == 0) {
        for (uint64_t r = ntiles * kTileRows + lane; r - lane < n; r += 32) {
            const bool has = r < n;
            unsigned key = 0; int32_t dd = 0; double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
            if (has) { key = q1_key_scalar<0>(rf, ls, r); dd = sd[r]; v0 = qty[r]; v1 = price[r]; v2 = disc[r]; v3 = tax[r]; }
            T.row(r, has, dd, v0, v1, v2, v3, key);
        }
    }
    q1_epilogue<G, kStagedThreads>(T, S, partials, out, ticket, dbg);
}

// ---- variant D: CTA-wide TMA pipeline ------------------------------------------------------------------------------------------------
// ncu on variant B shows every 32-byte sector requested twice from L2 (lts__t_sector_hit_rate 49 %, l1tex__m_xbar2l1tex_read_bytes = 1.84 x the
// DRAM bytes): the per-lane 16-byte LDGSTS requests of a warp are not merged into sector requests the way LDG.128 is.  Variant C (per-warp bulk
// copies of 64-512 bytes) fixed that but drowned the TMA unit in tiny copies (7 per 64 rows; 0.75 of the HBM rate).  Here ONE producer warp
// streams 1024-row tiles -- 7 bulk copies of 1-8 KB per tile, 4 tiles (156 KB) in flight per SM -- and the 16 consumer warps share each tile:
// full[stage] (transaction count) / empty[stage] (one arrival per consumer warp) mbarriers, no CTA barrier in the steady state.
constexpr int kCtaTileRows = 1024, kCtaStages = 4, kCtaConsumers = 16;
constexpr int kCtaTileBytes = kCtaTileRows * 38;   // 4096 shipdate + 4 x 8192 + 2 x 1024
constexpr int kCtaThreads = (kCtaConsumers + 1) * 32;

__device__ __forceinline__ void q1_mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

template <int G>
__global__ void __launch_bounds__(kCtaThreads, 1)
q1_tma_kernel(const int32_t *__restrict__ sd, const double *__restrict__ qty, const double *__restrict__ price,
              const double *__restrict__ disc, const double *__restrict__ tax, const uint8_t *__restrict__ rf,
              const uint8_t *__restrict__ ls, uint64_t n, int32_t cutoff, Q1Rec *__restrict__ partials,
              Q1Rec *__restrict__ out, unsigned *ticket, unsigned long long *dbg) {
    extern __shared__ __align__(16) unsigned char ring[];   // [stage][kCtaTileBytes], then full[] and empty[] mbarriers
    if (dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[blockIdx.x * 16] = t; }
    __shared__ Q1Shared S;
    S.dict[threadIdx.x & (MO_Q1_MAX_GROUPS - 1)] = kEmptyKey;
    S.overflow = 0; S.slow = 0;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(ring + (size_t)kCtaStages * kCtaTileBytes);
    unsigned long long *empty = full + kCtaStages;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kCtaStages; s++) { q1_mbar_init(&full[s], 1); q1_mbar_init(&empty[s], kCtaConsumers); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    Q1Thread<G> T;
    T.init(&S, cutoff, dbg != nullptr);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t ntiles = n / kCtaTileRows;
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint64_t mine = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == kCtaConsumers) {
        // ---------------- producer warp: lane 0 issues, the others idle (they rejoin at the epilogue barrier)
        if (lane == 0) {
            for (uint64_t k = 0; k < mine; k++) {
                const int stage = (int)(k % kCtaStages);
                if (k >= kCtaStages) q1_mbar_wait(&empty[stage], (unsigned)(((k / kCtaStages) - 1) & 1));   // every consumer warp released the slot
                unsigned char *s = ring + (size_t)stage * kCtaTileBytes;
                const uint64_t r0 = (blockIdx.x + k * gridDim.x) * (uint64_t)kCtaTileRows;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                q1_mbar_expect_tx(&full[stage], kCtaTileBytes);
                q1_bulk_copy(s, sd + r0, 4096, &full[stage]);
                q1_bulk_copy(s + 4096, qty + r0, 8192, &full[stage]);
                q1_bulk_copy(s + 12288, price + r0, 8192, &full[stage]);
                q1_bulk_copy(s + 20480, disc + r0, 8192, &full[stage]);
                q1_bulk_copy(s + 28672, tax + r0, 8192, &full[stage]);
                q1_bulk_copy(s + 36864, rf + r0, 1024, &full[stage]);
                q1_bulk_copy(s + 37888, ls + r0, 1024, &full[stage]);
            }
        }
    } else {
        // ---------------- consumer warps: thread t owns rows t and t + 512 of every tile
        for (uint64_t k = 0; k < mine; k++) {
            const int stage = (int)(k % kCtaStages);
            q1_mbar_wait(&full[stage], (unsigned)((k / kCtaStages) & 1));
            const unsigned char *s = ring + (size_t)stage * kCtaTileBytes;
            const uint64_t r0 = (blockIdx.x + k * gridDim.x) * (uint64_t)kCtaTileRows;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int rl = threadIdx.x + 512 * h;
                const int32_t d = *reinterpret_cast<const int32_t *>(s + rl * 4);
                const double q = *reinterpret_cast<const double *>(s + 4096 + rl * 8);
                const double pr = *reinterpret_cast<const double *>(s + 12288 + rl * 8);
                const double di = *reinterpret_cast<const double *>(s + 20480 + rl * 8);
                const double tx = *reinterpret_cast<const double *>(s + 28672 + rl * 8);
                const unsigned key = (unsigned)s[36864 + rl] | ((unsigned)s[37888 + rl] << 8);
                T.row(r0 + rl, true, d, q, pr, di, tx, key);
            }
            __syncwarp();
            if (lane == 0) q1_mbar_arrive(&empty[stage]);
        }
        // rows past the last full tile (< 1024): warp 0 of CTA 0, direct loads
        if (blockIdx.x == 0 && warp == 0) {
            for (uint64_t r = ntiles * kCtaTileRows + lane; r - lane < n; r += 32) {
                const bool has = r < n;
                unsigned key = 0; int32_t dd = 0; double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                if (has) { key = q1_key_scalar<0>(rf, ls, r); dd = sd[r]; v0 = qty[r]; v1 = price[r]; v2 = disc[r]; v3 = tax[r]; }
                T.row(r, has, dd, v0, v1, v2, v3, key);
            }
        }
    }
    q1_epilogue<G, kCtaThreads>(T, S, partials, out, ticket, dbg);
}

__host__ __device__ void q1_finalize(const Q1Rec &F, mo_q1_result_t *res, int64_t row_base = 0) {
    memset(res, 0, sizeof *res);
    int order[MO_Q1_MAX_GROUPS], ng = 0;
    for (int g = 0; g < MO_Q1_MAX_GROUPS; g++) if (F.slot[g].used && F.slot[g].cnt) order[ng++] = g;
    for (int i = 1; i < ng; i++)  // first-seen (row) order == the reference's group-id order
        for (int j = i; j > 0 && F.slot[order[j]].first_row < F.slot[order[j - 1]].first_row; j--) { int t = order[j]; order[j] = order[j - 1]; order[j - 1] = t; }
    res->ngroups = ng;
    for (int i = 0; i < ng; i++) {
        const Q1Slot &s = F.slot[order[i]];
        mo_q1_group_t &o = res->groups[i];
        o.returnflag = (uint8_t)(s.key & 0xff); o.linestatus = (uint8_t)((s.key >> 8) & 0xff);
        o.first_row = (int64_t)s.first_row + row_base;
        o.sum_qty = s.v[0]; o.sum_base_price = s.v[1]; o.sum_disc_price = s.v[2]; o.sum_charge = s.v[3]; o.sum_disc = s.v[4];
        o.count_order = (int64_t)s.cnt;
        const double c = (double)s.cnt;   // avg = float64(sum)/float64(cnt), sumavg2.go:331
        o.avg_qty = s.v[0] / c; o.avg_price = s.v[1] / c; o.avg_disc = s.v[4] / c;
    }
}

// asynchronous form: pick the pass that did not overflow, order the groups, write the caller's device result.  ngroups = -1 reports
// "more than MO_Q1_MAX_GROUPS distinct keys" (the synchronous form returns MO_RC_INVALID_ARGUMENT for that).
__global__ void q1_finalize_kernel(const Q1Rec *narrow, const Q1Rec *wide, mo_q1_result_t *res, int64_t row_base) {
    const Q1Rec *F = (narrow && !narrow->overflow) ? narrow : wide;
    if (!F || F->overflow) { memset(res, 0, sizeof *res); res->ngroups = -1; return; }
    q1_finalize(*F, res, row_base);
}

// MergeGroup on the device (mergeGroup.go:132-247): re-hash the partial group keys of `nparts` results and BatchMerge them IN ORDER
// (rank order => deterministic); groups come back in global first-seen row order.  One thread: <= 8 groups x nparts.
__global__ void q1_merge_kernel(const mo_q1_result_t *parts, uint64_t nparts, mo_q1_result_t *res) {
    mo_q1_result_t R; memset(&R, 0, sizeof R);
    bool bad = false;
    for (uint64_t p = 0; p < nparts && !bad; p++) {
        const mo_q1_result_t &P = parts[p];
        if (P.ngroups < 0 || P.ngroups > MO_Q1_MAX_GROUPS) { bad = true; break; }
        for (int64_t g = 0; g < P.ngroups; g++) {
            const mo_q1_group_t &s = P.groups[g];
            int dst = -1;
            for (int64_t x = 0; x < R.ngroups; x++) if (R.groups[x].returnflag == s.returnflag && R.groups[x].linestatus == s.linestatus) { dst = (int)x; break; }
            if (dst < 0) {
                if (R.ngroups == MO_Q1_MAX_GROUPS) { bad = true; break; }
                R.groups[R.ngroups++] = s;
                continue;
            }
            mo_q1_group_t &D = R.groups[dst];
            D.sum_qty = __dadd_rn(D.sum_qty, s.sum_qty); D.sum_base_price = __dadd_rn(D.sum_base_price, s.sum_base_price);
            D.sum_disc_price = __dadd_rn(D.sum_disc_price, s.sum_disc_price); D.sum_charge = __dadd_rn(D.sum_charge, s.sum_charge);
            D.sum_disc = __dadd_rn(D.sum_disc, s.sum_disc);
            D.count_order += s.count_order;
            if (s.first_row < D.first_row) D.first_row = s.first_row;
        }
    }
    if (bad) { memset(res, 0, sizeof *res); res->ngroups = -1; return; }
    for (int64_t i = 1; i < R.ngroups; i++)
        for (int64_t j = i; j > 0 && R.groups[j].first_row < R.groups[j - 1].first_row; j--) { mo_q1_group_t tmp = R.groups[j]; R.groups[j] = R.groups[j - 1]; R.groups[j - 1] = tmp; }
    for (int64_t i = 0; i < R.ngroups; i++) {
        mo_q1_group_t &o = R.groups[i];
        const double c = (double)o.count_order;
        o.avg_qty = o.sum_qty / c; o.avg_price = o.sum_base_price / c; o.avg_disc = o.sum_disc / c;
    }
    *res = R;
}

// MergeGroup for the Q6 shape: (sum, count) partials added in order; an empty partial is NULL and skipped (sumavg2.go:222-236)
__global__ void q6_merge_kernel(const double *parts, uint64_t nparts, Q6DevOut o) {
    double total = 0.0; long long cnt = 0; bool any = false;
    for (uint64_t p = 0; p < nparts; p++) {
        const double s = parts[2 * p]; const long long c = reinterpret_cast<const long long *>(parts)[2 * p + 1];
        if (c == 0) continue;
        total = any ? __dadd_rn(total, s) : s; any = true; cnt += c;
    }
    o.res[0] = total;
    if (o.res_words >= 2) reinterpret_cast<long long *>(o.res)[1] = cnt;
    if (o.rnulls) o.rnulls[0] = any ? 0ull : 1ull;
}

int g_q6_variant = 0, g_q1_variant = 0;  // tuning knobs (MoB200_SetTuning)
unsigned long long *g_q1_dbg = nullptr;   // optional per-CTA phase timestamps (MoB200_SetTuning("q1_debug", 1))

}  // namespace

namespace mob {

unsigned long long *q1_debug_buffer() { return g_q1_dbg; }
int g_plan_specialise = 1;
int xcall_plan(mo_xcall_args_t *args, uint64_t len);

extern int g_search_mode;
extern thread_local int g_last_tc_fallbacks, g_last_tc_refined, g_last_tc_kused;
extern int g_tc_pair_mode, g_tc_range_mb, g_tc_ladder_mode, g_tc_share_mode, g_tc_sched_mode;
extern std::atomic<int> g_one_term_skip;

int tuning_set(const char *name, int value) {
    if (!strcmp(name, "search_mode")) { g_search_mode = value; return 0; }
    if (!strcmp(name, "get_tc_fallbacks")) return g_last_tc_fallbacks;
    if (!strcmp(name, "get_tc_refined")) return g_last_tc_refined;
    if (!strcmp(name, "tc_pair")) { g_tc_pair_mode = (int)value; return 0; }
    if (!strcmp(name, "tc_range_mb")) { g_tc_range_mb = (int)value; return 0; }
    if (!strcmp(name, "tc_ladder")) { g_tc_ladder_mode = (int)value; g_one_term_skip.store(0); return 0; }
    if (!strcmp(name, "get_tc_kused")) return g_last_tc_kused;
    if (!strcmp(name, "tc_share")) { g_tc_share_mode = (int)value; return 0; }
    if (!strcmp(name, "tc_sched")) { g_tc_sched_mode = (int)value; return 0; }
    if (!strcmp(name, "q6_variant")) { g_q6_variant = value; return 0; }
    if (!strcmp(name, "plan_specialise")) { g_plan_specialise = value; return 0; }
    if (!strcmp(name, "q1_variant")) { g_q1_variant = value; return 0; }
    if (!strcmp(name, "q1_debug")) {
        if (value && !g_q1_dbg) { if (cudaMalloc((void **)&g_q1_dbg, 8 * 16 * 1024) != cudaSuccess) return -1; cudaMemset(g_q1_dbg, 0, 8 * 16 * 1024); }
        if (!value && g_q1_dbg) { cudaFree(g_q1_dbg); g_q1_dbg = nullptr; }
        return 0;
    }
    return -1;
}

// chunk of rows processed per launch when inputs are staged from the host (bounds device scratch)
static const uint64_t kHostChunkRows = 32ull << 20;

static bool aligned_to(const void *p, uintptr_t a) { return (((uintptr_t)p) & (a - 1)) == 0; }

static int launch_q6(ThreadCtx &t, const int32_t *sd, const double *disc, const double *qty, const double *price,
                     uint64_t n, const mo_q6_params_t &P, Q6Rec *hrec, Q6DevOut dres = Q6DevOut{nullptr, 0, nullptr}) {
    if (!aligned_to(sd, 8) || !aligned_to(disc, 16) || !aligned_to(qty, 16) || !aligned_to(price, 16)) {
        set_error("q6: columns must be 16-byte aligned (int32 column 8-byte)"); return MO_RC_INVALID_ARGUMENT;
    }
    int variant = g_q6_variant;
    int ctas = (variant == 1 || variant == 3) ? 2 : 4;
    int grid = num_sms() * ctas;
    uint64_t work = (n / 2 + kThreads - 1) / kThreads;
    if ((uint64_t)grid > work) grid = work ? (int)work : 1;
    Q6Rec *partials = (Q6Rec *)arena_alloc(t, sizeof(Q6Rec) * (size_t)(grid + 1));
    if (!partials) return MO_RC_INTERNAL_ERROR;
    Q6Rec *out = partials + grid;
    cudaEventRecord(t.kev0, t.stream);
    switch (variant) {
    case 1: q6_kernel<4, 2><<<grid, kThreads, 0, t.stream>>>(sd, disc, qty, price, n, P, partials, out, t.ctrl, dres); break;
    case 2: q6_kernel<2, 4><<<grid, kThreads, 0, t.stream>>>(sd, disc, qty, price, n, P, partials, out, t.ctrl, dres); break;
    case 3: q6_kernel<8, 2><<<grid, kThreads, 0, t.stream>>>(sd, disc, qty, price, n, P, partials, out, t.ctrl, dres); break;
    default: q6_kernel<4, 4><<<grid, kThreads, 0, t.stream>>>(sd, disc, qty, price, n, P, partials, out, t.ctrl, dres); break;
    }
    cudaEventRecord(t.kev1, t.stream);
    MOB_LAUNCH_CHECK();
    if (dres.res) return MO_RC_SUCCESS;   // asynchronous form: the last CTA wrote the caller's device result
    return read_back(t, hrec, out, sizeof(Q6Rec));
}

int xcall_q6(mo_xcall_args_t *args, uint64_t len) {
    ThreadCtx &t = tctx();
    if (!t.ready) return MO_RC_INTERNAL_ERROR;
    if (!args[5].pdata || args[5].dataSz < sizeof(mo_q6_params_t)) { set_error("q6: params missing"); return MO_RC_INVALID_ARGUMENT; }
    if (!args[0].pdata || args[0].dataSz < 8) { set_error("q6: result must hold 8 bytes"); return MO_RC_INVALID_ARGUMENT; }
    if (args[1].dataSz < 4 * len || args[2].dataSz < 8 * len || args[3].dataSz < 8 * len || args[4].dataSz < 8 * len) {
        set_error("q6: column shorter than len"); return MO_RC_INVALID_ARGUMENT;
    }
    mo_q6_params_t P;
    if (is_device_ptr(args[5].pdata)) { int rc = read_back(t, &P, args[5].pdata, sizeof P); if (rc) return rc; }
    else memcpy(&P, args[5].pdata, sizeof P);
    bool nullable = false;
    for (int i = 1; i <= 4; i++) nullable = nullable || args[i].pnulls != nullptr;
    if (nullable) {
        // nullable inputs: the same query through the generic fused operator (plan.cu), which carries a nulls bitmap on every column --
        // a NULL predicate operand rejects the row, a NULL product is skipped by SUM.  The specialised kernel above is the no-nulls fast path.
        mo_plan_t Q; memset(&Q, 0, sizeof Q);
        Q.ncols = 4; Q.col_type[0] = MO_T_DATE; Q.col_type[1] = Q.col_type[2] = Q.col_type[3] = MO_T_FLOAT64;
        Q.npreds = 4;
        Q.pred[0] = mo_plan_pred_t{0, 3, (double)P.date_lo, 0.0}; Q.pred[1] = mo_plan_pred_t{0, 4, (double)P.date_hi, 0.0};
        Q.pred[2] = mo_plan_pred_t{1, 6, P.disc_lo, P.disc_hi};    Q.pred[3] = mo_plan_pred_t{2, 4, P.qty_hi, 0.0};
        Q.ninstr = 1; Q.instr[0] = mo_plan_instr_t{MO_PLAN_OP_MUL, 3, 1, 0, 0.0};      // l_extendedprice * l_discount
        Q.naggs = 1; Q.agg[0] = mo_plan_agg_t{MO_AGG_SUM, 4};
        struct { mo_plan_result_header_t h; mo_plan_group_t g; mo_plan_agg_value_t a; } R;
        memset(&R, 0, sizeof R);
        mo_xcall_args_t pa[6]; memset(pa, 0, sizeof pa);
        pa[0].pdata = (uint8_t *)&R; pa[0].dataSz = sizeof R;
        pa[1].pdata = (uint8_t *)&Q; pa[1].dataSz = sizeof Q;
        pa[2] = args[1]; pa[3] = args[2]; pa[4] = args[3]; pa[5] = args[4];
        int rc = xcall_plan(pa, len);
        if (rc) return rc;
        const bool any_ = R.h.ngroups > 0 && R.a.count > 0;
        const double sum_ = any_ ? R.a.value : 0.0;
        const int64_t rows_ = R.h.ngroups > 0 ? R.g.rows : 0;
        const uint64_t nullword_ = any_ ? 0ull : 1ull;
        if (is_device_ptr(args[0].pdata)) {
            MOB_CUDA_TRY(cudaMemcpyAsync(args[0].pdata, &sum_, 8, cudaMemcpyHostToDevice, t.stream));
            if (args[0].dataSz >= 16) MOB_CUDA_TRY(cudaMemcpyAsync(args[0].pdata + 8, &rows_, 8, cudaMemcpyHostToDevice, t.stream));
            MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
        } else { memcpy(args[0].pdata, &sum_, 8); if (args[0].dataSz >= 16) memcpy(args[0].pdata + 8, &rows_, 8); }
        if (args[0].pnulls) {
            if (is_device_ptr(args[0].pnulls)) { MOB_CUDA_TRY(cudaMemcpyAsync(args[0].pnulls, &nullword_, 8, cudaMemcpyHostToDevice, t.stream)); MOB_CUDA_TRY(cudaStreamSynchronize(t.stream)); }
            else memcpy(args[0].pnulls, &nullword_, 8);
        }
        return MO_RC_SUCCESS;
    }

    const bool dev = is_device_ptr(args[1].pdata);
    for (int i = 2; i <= 4; i++) if (is_device_ptr(args[i].pdata) != dev) { set_error("q6: columns must all be host or all device"); return MO_RC_INVALID_ARGUMENT; }

    double sum = 0.0; unsigned long long cnt = 0; bool any = false;
    int rc = MO_RC_SUCCESS;
    if (dev && len && is_device_ptr(args[0].pdata) && (!args[0].pnulls || is_device_ptr(args[0].pnulls))) {
        // resident columns AND a device result: enqueue only (no read-back, no synchronisation); the result is ordered on the calling
        // thread's stream, ready for NCCL + MO_XCALL_Q6_MERGE or a later MoB200_Download
        rc = launch_q6(t, (const int32_t *)args[1].pdata, (const double *)args[2].pdata, (const double *)args[3].pdata,
                       (const double *)args[4].pdata, len, P, nullptr, Q6DevOut{(double *)args[0].pdata, args[0].dataSz / 8, args[0].pnulls});
        arena_reset(t);
        return rc;
    }
    if (dev || len == 0) {
        if (len) {
            Q6Rec r;
            rc = launch_q6(t, (const int32_t *)args[1].pdata, (const double *)args[2].pdata, (const double *)args[3].pdata,
                           (const double *)args[4].pdata, len, P, &r);
            sum = r.sum; cnt = r.cnt; any = r.cnt != 0;
        }
        arena_reset(t);
    } else {
        // host-resident columns: stream block ranges through the arena; partial sums are merged in chunk order
        // exactly like per-pipeline partials in MergeGroup (BatchMerge, sumavg2.go:222-236)
        for (uint64_t r0 = 0; r0 < len && rc == MO_RC_SUCCESS; r0 += kHostChunkRows) {
            uint64_t m = len - r0 < kHostChunkRows ? len - r0 : kHostChunkRows;
            Stager st(t);
            const int32_t *d_sd = (const int32_t *)st.in(args[1].pdata + 4 * r0, 4 * m);
            const double *d_di = (const double *)st.in(args[2].pdata + 8 * r0, 8 * m);
            const double *d_q = (const double *)st.in(args[3].pdata + 8 * r0, 8 * m);
            const double *d_p = (const double *)st.in(args[4].pdata + 8 * r0, 8 * m);
            if (st.failed) { st.finish(); return MO_RC_INTERNAL_ERROR; }
            Q6Rec r;
            rc = launch_q6(t, d_sd, d_di, d_q, d_p, m, P, &r);
            if (rc == MO_RC_SUCCESS && r.cnt) { sum = any ? sum + r.sum : r.sum; any = true; cnt += r.cnt; }
            int frc = st.finish();
            if (!rc) rc = frc;
        }
    }
    if (rc) return rc;
    uint64_t nullword = any ? 0ull : 1ull;
    int64_t c64 = (int64_t)cnt;
    if (is_device_ptr(args[0].pdata)) {
        MOB_CUDA_TRY(cudaMemcpyAsync(args[0].pdata, &sum, 8, cudaMemcpyHostToDevice, t.stream));
        if (args[0].dataSz >= 16) MOB_CUDA_TRY(cudaMemcpyAsync(args[0].pdata + 8, &c64, 8, cudaMemcpyHostToDevice, t.stream));
        MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
    } else {
        memcpy(args[0].pdata, &sum, 8);
        if (args[0].dataSz >= 16) memcpy(args[0].pdata + 8, &c64, 8);
    }
    if (args[0].pnulls) {
        if (is_device_ptr(args[0].pnulls)) { MOB_CUDA_TRY(cudaMemcpyAsync(args[0].pnulls, &nullword, 8, cudaMemcpyHostToDevice, t.stream)); MOB_CUDA_TRY(cudaStreamSynchronize(t.stream)); }
        else memcpy(args[0].pnulls, &nullword, 8);
    }
    return MO_RC_SUCCESS;
}

template <int KEYMODE>
static int launch_q1(ThreadCtx &t, const int32_t *sd, const double *qty, const double *price, const double *disc, const double *tax,
                     const uint8_t *rf, const uint8_t *ls, uint64_t n, int32_t cutoff, Q1Rec *hrec, mo_q1_result_t *dres = nullptr, int64_t row_base = 0) {
    if (!aligned_to(sd, 8) || !aligned_to(qty, 16) || !aligned_to(price, 16) || !aligned_to(disc, 16) || !aligned_to(tax, 16) ||
        !aligned_to(rf, KEYMODE ? 8 : 2) || !aligned_to(ls, KEYMODE ? 8 : 2)) {
        set_error("q1: columns must be 16-byte aligned (int32 column 8-byte, key columns 2/8-byte)"); return MO_RC_INVALID_ARGUMENT;
    }
    // variant: 0 = auto (cp.async-staged kernel, 3 stages, for packed keys; register kernel otherwise), 1 = register kernel,
    // 2 = register kernel with 8 group slots, 3 = staged with 4 stages, 4 = staged with 5 stages
    const bool can_stage = KEYMODE == 0 && aligned_to(sd, 16) && aligned_to(rf, 16) && aligned_to(ls, 16) && n >= kTileRows;
    Q1Rec *first_out = nullptr;   // asynchronous form: result record of the 4-slot pass
    for (int attempt = 0; attempt < 2; attempt++) {
        const bool wide = attempt == 1 || g_q1_variant == 2;
        const bool staged = can_stage && !wide && g_q1_variant != 1;
        int grid = num_sms() * (staged ? 1 : 2);
        const int threads = staged ? kStagedThreads : kThreads;
        uint64_t work = staged ? (n / kTileRows + (kStagedThreads / 32) - 1) / (kStagedThreads / 32) : (n / 2 + kThreads - 1) / kThreads;
        if ((uint64_t)grid > work) grid = work ? (int)work : 1;
        if (grid > kQ1MaxGrid) grid = kQ1MaxGrid;
        (void)threads;
        Q1Rec *partials = (Q1Rec *)arena_alloc(t, sizeof(Q1Rec) * (size_t)(grid + 1));
        if (!partials) return MO_RC_INTERNAL_ERROR;
        Q1Rec *out = partials + grid;
        if (attempt == 0 || !dres) cudaEventRecord(t.kev0, t.stream);
        if (staged && g_q1_variant >= 5 && g_q1_variant <= 7) {
            const int stages = g_q1_variant == 5 ? 3 : (g_q1_variant == 6 ? 4 : 5);
            const size_t smem = (size_t)(kStagedThreads / 32) * stages * kTileBytes + (size_t)(kStagedThreads / 32) * stages * 8;
            static bool battr[3] = {false, false, false};
            const int bi = g_q1_variant - 5;
            if (!battr[bi]) {
                cudaError_t e = stages == 3 ? cudaFuncSetAttribute(q1_bulk_kernel<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                              : stages == 4 ? cudaFuncSetAttribute(q1_bulk_kernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                            : cudaFuncSetAttribute(q1_bulk_kernel<4, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) { set_error("q1: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return MO_RC_INTERNAL_ERROR; }
                battr[bi] = true;
            }
            if (stages == 3) q1_bulk_kernel<4, 3><<<grid, kStagedThreads, smem, t.stream>>>(sd, qty, price, disc, tax, rf, ls, n, cutoff, partials, out, t.ctrl, g_q1_dbg);
            else if (stages == 4) q1_bulk_kernel<4, 4><<<grid, kStagedThreads, smem, t.stream>>>(sd, qty, price, disc, tax, rf, ls, n, cutoff, partials, out, t.ctrl, g_q1_dbg);
            else q1_bulk_kernel<4, 5><<<grid, kStagedThreads, smem, t.stream>>>(sd, qty, price, disc, tax, rf, ls, n, cutoff, partials, out, t.ctrl, g_q1_dbg);
        } else if (staged && (g_q1_variant == 10 || g_q1_variant == 0) && n >= (uint64_t)kCtaTileRows * 64) {   // default for packed keys (measured 1.07 vs 0.83 of the HBM rate for variant B)
            grid = num_sms();
            const uint64_t nt = n / kCtaTileRows;
            if ((uint64_t)grid > nt) grid = (int)nt;
            const size_t smem = (size_t)kCtaStages * kCtaTileBytes + 2 * kCtaStages * 8;
            static bool tattr = false;
            if (!tattr) {
                cudaError_t e = cudaFuncSetAttribute(q1_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) { set_error("q1: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return MO_RC_INTERNAL_ERROR; }
                tattr = true;
            }
            q1_tma_kernel<4><<<grid, kCtaThreads, smem, t.stream>>>(sd, qty, price, disc, tax, rf, ls, n, cutoff, partials, out, t.ctrl, g_q1_dbg);
        } else if (staged && (g_q1_variant == 8 || g_q1_variant == 9) && n < (1ull << 32)) {
            // shared-memory accumulators: 8 = 3 stages, 9 = 2 stages
            const int stages = g_q1_variant == 8 ? 3 : 2;
            const size_t smem = (size_t)(kStagedThreads / 32) * stages * kTileBytes + Q1SmemAcc<kStagedThreads>::kBytes;
            static bool sattr[2] = {false, false};
            if (!sattr[stages - 2]) {
                cudaError_t e = stages == 3 ? cudaFuncSetAttribute(q1_staged_kernel<4, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                            : cudaFuncSetAttribute(q1_staged_kernel<4, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) { set_error("q1: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return MO_RC_INTERNAL_ERROR; }
                sattr[stages - 2] = true;
            }
            if (stages == 3) q1_staged_kernel<4, 3, true><<<grid, kStagedThreads, smem, t.stream>>>(sd, qty, price, disc, tax, rf, ls, n, cutoff, partials, out, t.ctrl, g_q1_dbg);
            else q1_staged_kernel<4, 2, true><<<grid, kStagedThreads, smem, t.stream>>>(sd, qty, price, disc, tax, rf, ls, n, cutoff, partials, out, t.ctrl, g_q1_dbg);
        } else if (staged) {
            const int stages = g_q1_variant == 3 ? 4 : (g_q1_variant == 4 ? 5 : 3);   // 3 stages measured best (tools/tune.py q1)
            const size_t smem = (size_t)(kStagedThreads / 32) * stages * kTileBytes;
            static bool attr_done[3] = {false, false, false};
            if (!attr_done[stages - 3]) {
                cudaError_t e = stages == 3 ? cudaFuncSetAttribute(q1_staged_kernel<4, 3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                              : stages == 4 ? cudaFuncSetAttribute(q1_staged_kernel<4, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)
                                            : cudaFuncSetAttribute(q1_staged_kernel<4, 5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e != cudaSuccess) { set_error("q1: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return MO_RC_INTERNAL_ERROR; }
                attr_done[stages - 3] = true;
            }
            if (stages == 3) q1_staged_kernel<4, 3, false><<<grid, kStagedThreads, smem, t.stream>>>(sd, qty, price, disc, tax, rf, ls, n, cutoff, partials, out, t.ctrl, g_q1_dbg);
            else if (stages == 4) q1_staged_kernel<4, 4, false><<<grid, kStagedThreads, smem, t.stream>>>(sd, qty, price, disc, tax, rf, ls, n, cutoff, partials, out, t.ctrl, g_q1_dbg);
            else q1_staged_kernel<4, 5, false><<<grid, kStagedThreads, smem, t.stream>>>(sd, qty, price, disc, tax, rf, ls, n, cutoff, partials, out, t.ctrl, g_q1_dbg);
        } else if (!wide) {
            q1_kernel<4, 2, 2, KEYMODE><<<grid, kThreads, 0, t.stream>>>(sd, qty, price, disc, tax, rf, ls, n, cutoff, partials, out, t.ctrl, g_q1_dbg, nullptr);
        } else {
            q1_kernel<8, 2, 2, KEYMODE><<<grid, kThreads, 0, t.stream>>>(sd, qty, price, disc, tax, rf, ls, n, cutoff, partials, out, t.ctrl, g_q1_dbg, first_out);
        }
        if (attempt == 0 || !dres) cudaEventRecord(t.kev1, t.stream);
        MOB_LAUNCH_CHECK();
        if (dres) {
            // no read-back: enqueue the gated 8-slot retry behind a 4-slot pass, then the finalize kernel
            if (!wide) { first_out = out; continue; }
            q1_finalize_kernel<<<1, 1, 0, t.stream>>>(first_out, out, dres, row_base);
            MOB_LAUNCH_CHECK();
            return MO_RC_SUCCESS;
        }
        int rc = read_back(t, hrec, out, sizeof(Q1Rec));
        if (rc) return rc;
        if (!hrec->overflow) return MO_RC_SUCCESS;
        if (wide) break;  // more than MO_Q1_MAX_GROUPS distinct keys
    }
    set_error("q1: more than %d distinct group keys; the fused small-cardinality kernel does not apply", MO_Q1_MAX_GROUPS);
    return MO_RC_INVALID_ARGUMENT;
}

static void q1_merge_rec(Q1Rec &F, const Q1Rec &R, bool &first) {
    if (first) { F = R; first = false; return; }
    for (int g = 0; g < MO_Q1_MAX_GROUPS; g++) {
        const Q1Slot &s = R.slot[g];
        if (!s.used || s.cnt == 0) continue;
        int dst = -1;
        for (int x = 0; x < MO_Q1_MAX_GROUPS; x++) {
            if (F.slot[x].used && F.slot[x].key == s.key) { dst = x; break; }
            if (!F.slot[x].used || F.slot[x].key == kEmptyKey) { dst = x; F.slot[x].key = s.key; F.slot[x].used = 1; F.slot[x].cnt = 0; F.slot[x].first_row = ~0ull; for (int j = 0; j < kQ1Vals; j++) F.slot[x].v[j] = 0; break; }
        }
        if (dst < 0) { F.overflow = 1; continue; }
        Q1Slot &D = F.slot[dst];
        D.cnt += s.cnt;
        if (s.first_row < D.first_row) D.first_row = s.first_row;
        for (int j = 0; j < kQ1Vals; j++) D.v[j] = D.v[j] + s.v[j];
    }
}

int xcall_q1(mo_xcall_args_t *args, uint64_t len) {
    ThreadCtx &t = tctx();
    if (!t.ready) return MO_RC_INTERNAL_ERROR;
    if (!args[0].pdata || args[0].dataSz < sizeof(mo_q1_result_t)) { set_error("q1: result buffer too small"); return MO_RC_INVALID_ARGUMENT; }
    if (!args[8].pdata || args[8].dataSz < 4) { set_error("q1: cutoff param missing"); return MO_RC_INVALID_ARGUMENT; }
    if (args[1].dataSz < 4 * len) { set_error("q1: shipdate shorter than len"); return MO_RC_INVALID_ARGUMENT; }
    for (int i = 2; i <= 5; i++) if (args[i].dataSz < 8 * len) { set_error("q1: column %d shorter than len", i); return MO_RC_INVALID_ARGUMENT; }
    bool q1_nullable = false;
    for (int i = 1; i <= 7; i++) q1_nullable = q1_nullable || args[i].pnulls != nullptr;
    int keymode;
    if (args[6].dataSz == len && args[7].dataSz == len) keymode = 0;
    else if (args[6].dataSz == 24 * len && args[7].dataSz == 24 * len) keymode = 1;
    else { set_error("q1: key columns must be packed uint8 (len bytes) or varlena cells (24*len bytes)"); return MO_RC_INVALID_ARGUMENT; }
    int32_t cutoff;
    if (is_device_ptr(args[8].pdata)) { int rc = read_back(t, &cutoff, args[8].pdata, 4); if (rc) return rc; }
    else memcpy(&cutoff, args[8].pdata, 4);
    const bool dev = is_device_ptr(args[1].pdata);
    for (int i = 2; i <= 7; i++) if (is_device_ptr(args[i].pdata) != dev) { set_error("q1: columns must all be host or all device"); return MO_RC_INVALID_ARGUMENT; }
    const uint64_t ksz = keymode ? 24 : 1;
    int64_t row_base = 0;   // optional: params = {int32 cutoff; int32 pad; int64 row_base} -> first_row values are global row numbers
    if (args[8].dataSz >= 16 && !is_device_ptr(args[8].pdata)) memcpy(&row_base, args[8].pdata + 8, 8);
    if (q1_nullable) {
        // nullable inputs: the generic fused operator (plan.cu).  NULL shipdate rejects the row; NULL values are skipped by their aggregates;
        // a NULL key is its own group (has_null_keys; its key byte reads 0 here).  Packed uint8 keys only.
        if (keymode != 0) { set_error("q1: nullable columns need packed uint8 key columns"); return MO_RC_INVALID_ARGUMENT; }
        mo_plan_t Q; memset(&Q, 0, sizeof Q);
        Q.ncols = 7; Q.col_type[0] = MO_T_DATE; for (int c = 1; c <= 4; c++) Q.col_type[c] = MO_T_FLOAT64; Q.col_type[5] = Q.col_type[6] = MO_T_UINT8;
        Q.row_base = row_base; Q.has_null_keys = (args[6].pnulls || args[7].pnulls) ? 1 : 0;
        Q.npreds = 1; Q.pred[0] = mo_plan_pred_t{0, 5, (double)cutoff, 0.0};
        Q.ninstr = 5;
        Q.instr[0] = mo_plan_instr_t{MO_PLAN_OP_CONST, 0, 0, 0, 1.0};        // slot 7: 1
        Q.instr[1] = mo_plan_instr_t{MO_PLAN_OP_SUB, 7, 3, 0, 0.0};          // slot 8: 1 - l_discount
        Q.instr[2] = mo_plan_instr_t{MO_PLAN_OP_MUL, 2, 8, 0, 0.0};          // slot 9: l_extendedprice * (1 - l_discount)
        Q.instr[3] = mo_plan_instr_t{MO_PLAN_OP_ADD, 7, 4, 0, 0.0};          // slot 10: 1 + l_tax
        Q.instr[4] = mo_plan_instr_t{MO_PLAN_OP_MUL, 9, 10, 0, 0.0};         // slot 11: ... * (1 + l_tax)
        Q.nkeys = 2; Q.key_col[0] = 5; Q.key_col[1] = 6;
        Q.naggs = 6;
        Q.agg[0] = mo_plan_agg_t{MO_AGG_SUM, 1}; Q.agg[1] = mo_plan_agg_t{MO_AGG_SUM, 2}; Q.agg[2] = mo_plan_agg_t{MO_AGG_SUM, 9}; Q.agg[3] = mo_plan_agg_t{MO_AGG_SUM, 11};
        Q.agg[4] = mo_plan_agg_t{MO_AGG_SUM, 3}; Q.agg[5] = mo_plan_agg_t{MO_AGG_COUNT, -1};
        struct Rec { mo_plan_group_t g; mo_plan_agg_value_t a[6]; };
        struct { mo_plan_result_header_t h; Rec r[MO_Q1_MAX_GROUPS]; } R;
        memset(&R, 0, sizeof R);
        mo_xcall_args_t pa[9]; memset(pa, 0, sizeof pa);
        pa[0].pdata = (uint8_t *)&R; pa[0].dataSz = sizeof R;
        pa[1].pdata = (uint8_t *)&Q; pa[1].dataSz = sizeof Q;
        for (int c = 0; c < 7; c++) pa[2 + c] = args[1 + c];
        int rc = xcall_plan(pa, len);
        if (rc) { if (rc == MO_RC_INVALID_ARGUMENT) set_error("q1: more than %d distinct group keys", MO_Q1_MAX_GROUPS); return rc; }
        mo_q1_result_t res; memset(&res, 0, sizeof res);
        res.ngroups = R.h.ngroups;
        for (int64_t g = 0; g < R.h.ngroups; g++) {
            const Rec &r = R.r[g]; mo_q1_group_t &o = res.groups[g];
            const int shift = Q.has_null_keys ? 8 : 0;   // has_null mode: marker byte, then the value byte
            const bool n0 = Q.has_null_keys && (r.g.key & 0xff) != 0;
            o.returnflag = n0 ? 0 : (uint8_t)((r.g.key >> shift) & 0xff);
            const int off1 = Q.has_null_keys ? (n0 ? 1 : 2) : 1;
            const bool n1 = Q.has_null_keys && ((r.g.key >> (8 * off1)) & 0xff) != 0;
            o.linestatus = n1 ? 0 : (uint8_t)((r.g.key >> (8 * (off1 + (Q.has_null_keys ? 1 : 0)))) & 0xff);
            o.first_row = r.g.first_row;
            o.sum_qty = r.a[0].value; o.sum_base_price = r.a[1].value; o.sum_disc_price = r.a[2].value; o.sum_charge = r.a[3].value; o.sum_disc = r.a[4].value;
            o.avg_qty = r.a[0].count ? r.a[0].value / (double)r.a[0].count : 0.0;
            o.avg_price = r.a[1].count ? r.a[1].value / (double)r.a[1].count : 0.0;
            o.avg_disc = r.a[4].count ? r.a[4].value / (double)r.a[4].count : 0.0;
            o.count_order = r.a[5].count;
        }
        if (is_device_ptr(args[0].pdata)) { MOB_CUDA_TRY(cudaMemcpyAsync(args[0].pdata, &res, sizeof res, cudaMemcpyHostToDevice, t.stream)); MOB_CUDA_TRY(cudaStreamSynchronize(t.stream)); }
        else memcpy(args[0].pdata, &res, sizeof res);
        return MO_RC_SUCCESS;
    }

    if (dev && len && is_device_ptr(args[0].pdata)) {
        // resident columns AND a device result: enqueue only (see xcall_q6)
        mo_q1_result_t *dres = (mo_q1_result_t *)args[0].pdata;
        int rc = keymode ? launch_q1<1>(t, (const int32_t *)args[1].pdata, (const double *)args[2].pdata, (const double *)args[3].pdata, (const double *)args[4].pdata, (const double *)args[5].pdata, args[6].pdata, args[7].pdata, len, cutoff, nullptr, dres, row_base)
                         : launch_q1<0>(t, (const int32_t *)args[1].pdata, (const double *)args[2].pdata, (const double *)args[3].pdata, (const double *)args[4].pdata, (const double *)args[5].pdata, args[6].pdata, args[7].pdata, len, cutoff, nullptr, dres, row_base);
        arena_reset(t);
        return rc;
    }

    Q1Rec F; memset(&F, 0, sizeof F); bool first = true; int rc = MO_RC_SUCCESS;
    for (int g = 0; g < MO_Q1_MAX_GROUPS; g++) F.slot[g].key = kEmptyKey;
    if (dev || len == 0) {
        if (len) {
            Q1Rec R;
            rc = keymode ? launch_q1<1>(t, (const int32_t *)args[1].pdata, (const double *)args[2].pdata, (const double *)args[3].pdata, (const double *)args[4].pdata, (const double *)args[5].pdata, args[6].pdata, args[7].pdata, len, cutoff, &R)
                         : launch_q1<0>(t, (const int32_t *)args[1].pdata, (const double *)args[2].pdata, (const double *)args[3].pdata, (const double *)args[4].pdata, (const double *)args[5].pdata, args[6].pdata, args[7].pdata, len, cutoff, &R);
            if (!rc) q1_merge_rec(F, R, first);
        }
        arena_reset(t);
    } else {
        for (uint64_t r0 = 0; r0 < len && rc == MO_RC_SUCCESS; r0 += kHostChunkRows) {
            uint64_t m = len - r0 < kHostChunkRows ? len - r0 : kHostChunkRows;
            Stager st(t);
            const int32_t *d_sd = (const int32_t *)st.in(args[1].pdata + 4 * r0, 4 * m);
            const double *d_q = (const double *)st.in(args[2].pdata + 8 * r0, 8 * m);
            const double *d_p = (const double *)st.in(args[3].pdata + 8 * r0, 8 * m);
            const double *d_d = (const double *)st.in(args[4].pdata + 8 * r0, 8 * m);
            const double *d_t = (const double *)st.in(args[5].pdata + 8 * r0, 8 * m);
            const uint8_t *d_rf = (const uint8_t *)st.in(args[6].pdata + ksz * r0, ksz * m);
            const uint8_t *d_ls = (const uint8_t *)st.in(args[7].pdata + ksz * r0, ksz * m);
            if (st.failed) { st.finish(); return MO_RC_INTERNAL_ERROR; }
            Q1Rec R;
            rc = keymode ? launch_q1<1>(t, d_sd, d_q, d_p, d_d, d_t, d_rf, d_ls, m, cutoff, &R)
                         : launch_q1<0>(t, d_sd, d_q, d_p, d_d, d_t, d_rf, d_ls, m, cutoff, &R);
            if (!rc) {
                for (int g = 0; g < MO_Q1_MAX_GROUPS; g++) if (R.slot[g].used && R.slot[g].cnt) R.slot[g].first_row += r0;
                q1_merge_rec(F, R, first);
            }
            int frc = st.finish();
            if (!rc) rc = frc;
        }
    }
    if (rc) return rc;
    if (F.overflow) { set_error("q1: too many groups"); return MO_RC_INVALID_ARGUMENT; }
    mo_q1_result_t res;
    q1_finalize(F, &res, row_base);
    if (is_device_ptr(args[0].pdata)) {
        MOB_CUDA_TRY(cudaMemcpyAsync(args[0].pdata, &res, sizeof res, cudaMemcpyHostToDevice, t.stream));
        MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
    } else memcpy(args[0].pdata, &res, sizeof res);
    return MO_RC_SUCCESS;
}

// MO_XCALL_Q6_MERGE: args[0] = result as MO_XCALL_Q6_FILTER_SUM ; args[1] = len partial results of 16 bytes (sum f64, count i64)
int xcall_q6_merge(mo_xcall_args_t *args, uint64_t len) {
    ThreadCtx &t = tctx();
    if (!t.ready) return MO_RC_INTERNAL_ERROR;
    if (!args[0].pdata || args[0].dataSz < 8 || args[1].dataSz < 16 * len) { set_error("q6 merge: result needs 8 bytes, partials 16 bytes each"); return MO_RC_INVALID_ARGUMENT; }
    const bool async = is_device_ptr(args[0].pdata) && (len == 0 || is_device_ptr(args[1].pdata));
    Stager st(t);
    const double *dparts = (const double *)st.in(args[1].pdata, 16 * len);
    double *dres = (double *)st.out(args[0].pdata, args[0].dataSz >= 16 ? 16 : 8);
    uint64_t *dn = (uint64_t *)st.out(args[0].pnulls, args[0].pnulls ? 8 : 0);
    if (st.failed) { st.finish(); return MO_RC_INTERNAL_ERROR; }
    q6_merge_kernel<<<1, 1, 0, t.stream>>>(dparts, len, Q6DevOut{dres, args[0].dataSz / 8, dn});
    MOB_LAUNCH_CHECK();
    if (async) { st.release_async(); return MO_RC_SUCCESS; }
    return st.finish();
}

// MO_XCALL_Q1_MERGE: args[0] = mo_q1_result_t ; args[1] = len partial mo_q1_result_t (first_row already global: row_base param)
int xcall_q1_merge(mo_xcall_args_t *args, uint64_t len) {
    ThreadCtx &t = tctx();
    if (!t.ready) return MO_RC_INTERNAL_ERROR;
    if (!args[0].pdata || args[0].dataSz < sizeof(mo_q1_result_t) || args[1].dataSz < sizeof(mo_q1_result_t) * len) { set_error("q1 merge: buffers too small"); return MO_RC_INVALID_ARGUMENT; }
    const bool async = is_device_ptr(args[0].pdata) && (len == 0 || is_device_ptr(args[1].pdata));
    Stager st(t);
    const mo_q1_result_t *dparts = (const mo_q1_result_t *)st.in(args[1].pdata, sizeof(mo_q1_result_t) * len);
    mo_q1_result_t *dres = (mo_q1_result_t *)st.out(args[0].pdata, sizeof(mo_q1_result_t));
    if (st.failed) { st.finish(); return MO_RC_INTERNAL_ERROR; }
    q1_merge_kernel<<<1, 1, 0, t.stream>>>(dparts, len, dres);
    MOB_LAUNCH_CHECK();
    if (async) { st.release_async(); return MO_RC_SUCCESS; }
    int rc = st.finish();
    if (!rc && !is_device_ptr(args[0].pdata) && ((mo_q1_result_t *)args[0].pdata)->ngroups < 0) { set_error("q1 merge: more than %d distinct group keys", MO_Q1_MAX_GROUPS); return MO_RC_INVALID_ARGUMENT; }
    return rc;
}

}  // namespace mob
