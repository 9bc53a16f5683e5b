// goelem.cu -- the Go elementwise engine's semantics behind XCall.
//
// The mo.h surface (elementwise.cu) reproduces the C kernels of cgo/arith.c / compare.c / logic.c: overflow is a flag raised
// after the loop.  The LIVE engine of the reference is the Go one:
//   opBinaryFixedFixedToFixed[WithErrorCheck]   pkg/sql/plan/function/baseTemplate.go:457-578 / 580-728
//   overflow-checked + - *                       arithmetic.go:222-469, arithmetic_overflow_check.go:29-310
//   / and % with the division-by-zero mode       arithmetic.go:481-514, 704-762; baseTemplate.go:1369, 1436
//   compare                                      func_compare.go:285,677,804,931,1058,1185
//   between                                      operator_between.go:138-199
//   n-ary three-valued AND / OR                  logicalOperator.go:36-102, 104-168
// whose conventions differ: rows whose result-null bit is set are skipped (the result null bitmap arrives pre-filled with
// NOT selectList, baseTemplate.go:473-486, and leaves as rnulls | n1 | n2), the first offending row (in row order) fails the
// call with ErrOutOfRange / ErrDivByZero, a zero divisor makes the row NULL when the session's mode says so.
//
// funcIds and argument layout: include/mo_b200.h (MO_XCALL_GO_*).  One thread per row, fully coalesced; the first offender is
// an atomicMin over row numbers, so the error is the one the serial Go loop raises.  Rows before it hold their results, as the
// Go loop leaves them; rows after it are unspecified (Go leaves them unwritten; the batch is discarded on error either way).
#include "common.cuh"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <type_traits>

using namespace mob;

namespace {

constexpr int kThreads = 256;
enum { OP_ADD = 0, OP_SUB, OP_MUL, OP_DIV, OP_MOD };
enum { CMP_EQ = 0, CMP_NE, CMP_GT, CMP_GE, CMP_LT, CMP_LE };
constexpr unsigned long long kNoRow = ~0ull;

// ---- overflow-checked scalar arithmetic (arithmetic_overflow_check.go): returns true when the row must fail -----------------
template <typename T> __device__ __forceinline__ bool go_add(T a, T b, T &r) {
    if (std::is_floating_point<T>::value) { r = a + b; return isinf((double)r); }                          // :295-310
    else if (std::is_signed<T>::value) {
        using U = typename std::make_unsigned<typename std::conditional<std::is_floating_point<T>::value, int, T>::type>::type;
        const T s = (T)((U)a + (U)b); r = s;
        return (a > 0 && b > 0 && s <= 0) || (a < 0 && b < 0 && s >= 0);                                   // :32-38
    } else { const T s = (T)(a + b); r = s; return s < a || s < b; }
}
template <typename T> __device__ __forceinline__ bool go_sub(T a, T b, T &r) {
    if (std::is_floating_point<T>::value) { r = a - b; return isinf((double)r); }
    else if (std::is_signed<T>::value) {
        using U = typename std::make_unsigned<typename std::conditional<std::is_floating_point<T>::value, int, T>::type>::type;
        const T s = (T)((U)a - (U)b); r = s;
        return (a > 0 && b < 0 && s < 0) || (a < 0 && b > 0 && s > 0);                                     // :104-112
    } else { r = (T)(a - b); return a < b; }
}
template <typename T> __device__ __forceinline__ bool go_mul(T a, T b, T &r) {
    if (std::is_floating_point<T>::value) { r = a * b; return false; }                                     // unchecked, arithmetic.go:444-451
    else if (std::is_signed<T>::value) {
        using U = typename std::make_unsigned<typename std::conditional<std::is_floating_point<T>::value, int, T>::type>::type;
        if (a == 0 || b == 0) { r = 0; return false; }
        const T tmin = std::numeric_limits<T>::min();
        if ((a == tmin && b == (T)-1) || (b == tmin && a == (T)-1)) { r = 0; return true; }                // :183
        using W = typename std::conditional<(sizeof(T) < 4), unsigned, U>::type;   // no int promotion (uint16 * uint16 would overflow int)
        const T s = (T)((W)(U)a * (W)(U)b); r = s;
        return (T)(s / b) != a;                                                                            // :189
    } else {
        if (a == 0 || b == 0) { r = 0; return false; }
        using W2 = typename std::conditional<(sizeof(T) < 4), unsigned, T>::type;
        r = (T)((W2)a * (W2)b);
        return a > (T)(std::numeric_limits<T>::max() / b);                                                 // :262
    }
}
template <typename T> __device__ __forceinline__ T go_mod(T x, T y) {   // y != 0; arithmetic.go:704-762
    if (std::is_same<T, float>::value) return (T)fmod((double)x, (double)y);
    else if (std::is_same<T, double>::value) return (T)fmod((double)x, (double)y);
    else if (std::is_signed<T>::value) return y == (T)-1 ? (T)0 : (T)(x % y);
    else return (T)(x % y);
}
// integer-only helper so that `%` is never instantiated for floats
template <typename T, bool F = std::is_floating_point<T>::value> struct ModOp { static __device__ __forceinline__ T run(T x, T y) { return go_mod<T>(x, y); } };
template <typename T> struct ModOp<T, true> { static __device__ __forceinline__ T run(T x, T y) { return (T)fmod((double)x, (double)y); } };

template <typename T, int OP>
__global__ void __launch_bounds__(kThreads)
go_arith_kernel(T *__restrict__ r, const T *__restrict__ a, const T *__restrict__ b, uint64_t n, int c1, int c2,
                uint64_t *rnulls, int div0_null, unsigned long long *first_bad) {
    for (uint64_t i = blockIdx.x * (uint64_t)kThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kThreads) {
        if (bm_test(rnulls, i)) continue;                       // bit i is only ever changed by this very thread
        const T x = a[c1 ? 0 : i], y = b[c2 ? 0 : i];
        T v;
        bool bad = false;
        if (OP == OP_DIV || OP == OP_MOD) {
            if (y == (T)0) {
                if (div0_null) atomicOr(reinterpret_cast<unsigned long long *>(&rnulls[i >> 6]), 1ull << (i & 63));
                else atomicMin(first_bad, (unsigned long long)i);
                continue;
            }
            if (OP == OP_DIV) v = (T)(x / y);                   // floats only (the host rejects integer types)
            else v = ModOp<T>::run(x, y);
        } else if (OP == OP_ADD) bad = go_add<T>(x, y, v);
        else if (OP == OP_SUB) bad = go_sub<T>(x, y, v);
        else bad = go_mul<T>(x, y, v);
        if (bad) { atomicMin(first_bad, (unsigned long long)i); continue; }
        r[i] = v;
    }
}

template <typename T> __device__ __forceinline__ bool go_cmp(T x, T y, int op) {
    switch (op) {
    case CMP_EQ: return x == y; case CMP_NE: return x != y;
    case CMP_GT: return x > y;  case CMP_GE: return x >= y;
    case CMP_LT: return x < y;  default: return x <= y;
    }
}
// 8 consecutive rows of a column as 128-bit loads (the group never straddles a bitmap word: 8 | 64)
template <typename T> __device__ __forceinline__ void load8(const T *p, uint64_t i0, bool cst, T *out) {
    if (cst) {
        const T v = p[0];
#pragma unroll
        for (int j = 0; j < 8; j++) out[j] = v;
    } else if (sizeof(T) >= 2) {
        constexpr int NV = (int)(sizeof(T) * 8 / 16);
        const int4 *q = reinterpret_cast<const int4 *>(p + i0);
#pragma unroll
        for (int v = 0; v < NV; v++) { const int4 w = ld_stream16(q + v); memcpy(reinterpret_cast<char *>(out) + 16 * v, &w, 16); }
    } else {
        const uint64_t w = *reinterpret_cast<const uint64_t *>(p + i0);
        memcpy(out, &w, 8);
    }
}

// float32(math.Round(float64(a) * pow) / pow): math.Round = half away from zero = round(); IEEE double multiply / divide
template <typename T> __device__ __forceinline__ T round_scale(T v, double) { return v; }
template <> __device__ __forceinline__ float round_scale<float>(float v, double pw) {
    return pw > 0.0 ? (float)__ddiv_rn(round(__dmul_rn((double)v, pw)), pw) : v;
}

// compare: a thread owns 8 consecutive rows -> one 8-byte store of the bool results when none of them is null
template <typename T>
__global__ void __launch_bounds__(kThreads)
go_compare_kernel(uint8_t *__restrict__ r, const T *__restrict__ a, const T *__restrict__ b, uint64_t n, int c1, int c2,
                  const uint64_t *__restrict__ rnulls, int op, int vec_ok, double pw) {
    // pw > 0 (float32 columns declared with scale > 0, func_compare.go:725-734): both sides are rounded to `scale` decimals first
    const uint64_t ngroups = vec_ok ? n / 8 : 0;
    for (uint64_t g = blockIdx.x * (uint64_t)kThreads + threadIdx.x; g < ngroups; g += (uint64_t)gridDim.x * kThreads) {
        const uint64_t i0 = g * 8;
        const unsigned nb = (unsigned)((rnulls[i0 >> 6] >> (i0 & 63)) & 0xffu);
        if (nb == 0xffu) continue;
        T x[8], y[8];
        load8<T>(a, i0, c1 != 0, x); load8<T>(b, i0, c2 != 0, y);
        uint64_t packed = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) packed |= (uint64_t)(go_cmp<T>(round_scale<T>(x[j], pw), round_scale<T>(y[j], pw), op) ? 1 : 0) << (8 * j);
        if (nb == 0) *reinterpret_cast<uint64_t *>(r + i0) = packed;
        else {
#pragma unroll
            for (int j = 0; j < 8; j++) if (!((nb >> j) & 1u)) r[i0 + j] = (uint8_t)(packed >> (8 * j));
        }
    }
    // rows not covered by whole groups (all rows when the buffers are not 16-byte aligned)
    for (uint64_t i = ngroups * 8 + blockIdx.x * (uint64_t)kThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kThreads) {
        if (bm_test(rnulls, i)) continue;
        r[i] = go_cmp<T>(round_scale<T>(a[c1 ? 0 : i], pw), round_scale<T>(b[c2 ? 0 : i], pw), op) ? 1 : 0;
    }
}

// rnulls |= n1 | n2 (operands that are not const), tail bits cleared; or "all rows null" (a const NULL operand)
__global__ void go_nulls_kernel(uint64_t *rnulls, const uint64_t *n1, const uint64_t *n2, uint64_t n, int all_null) {
    const uint64_t nw = (n + 63) >> 6;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < nw; w += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t v = all_null ? ~0ull : rnulls[w];
        if (!all_null) { if (n1) v |= n1[w]; if (n2) v |= n2[w]; }
        if (w == nw - 1 && (n & 63)) { if (all_null) v = rnulls[w] | ((1ull << (n & 63)) - 1); else v &= (1ull << (n & 63)) - 1; }
        rnulls[w] = v;
    }
}

// between: null input -> res = false and the null bit set; one warp covers an aligned 32-row half word of the bitmap
template <typename T>
__global__ void __launch_bounds__(kThreads)
go_between_kernel(uint8_t *__restrict__ r, const T *__restrict__ col, T lo, T hi, uint64_t n, const uint64_t *__restrict__ nulls, uint32_t *rnulls32) {
    const uint64_t n32 = (n + 31) & ~31ull;
    for (uint64_t i = blockIdx.x * (uint64_t)kThreads + threadIdx.x; i < n32; i += (uint64_t)gridDim.x * kThreads) {
        const bool in = i < n;
        const bool isnull = in && bm_test(nulls, i);
        if (in) { const T c = col[i]; r[i] = (!isnull && c >= lo && c <= hi) ? 1 : 0; }
        const unsigned m = __ballot_sync(0xffffffffu, isnull);
        if ((threadIdx.x & 31) == 0 && m) rnulls32[i >> 5] |= m;
    }
}

// n-ary three-valued AND / OR, folded left to right exactly as opMultiAnd / opMultiOr do, row by row
struct LogicParam { const uint8_t *col; const uint64_t *nulls; int kind; };   // kind 0 = flat, 1 = const, 2 = const NULL
constexpr int kMaxLogicParams = 16;
struct LogicParams { LogicParam p[kMaxLogicParams]; int n; };
__global__ void __launch_bounds__(kThreads)
go_multi_logic_kernel(uint8_t *__restrict__ r, uint32_t *__restrict__ rnulls32, uint64_t n, LogicParams P, int is_or, int any_null_from) {
    const uint64_t n32 = (n + 31) & ~31ull;
    for (uint64_t i = blockIdx.x * (uint64_t)kThreads + threadIdx.x; i < n32; i += (uint64_t)gridDim.x * kThreads) {
        const bool in = i < n;
        bool val = false, isnull = false;
        if (in) {
            const LogicParam p0 = P.p[0];
            if (p0.kind == 2) { val = false; isnull = true; }
            else if (p0.kind == 1) val = p0.col[0] != 0;
            else { val = p0.col[i] != 0; isnull = bm_test(p0.nulls, i); }
            for (int k = 1; k < P.n; k++) {
                const LogicParam pk = P.p[k];
                if (pk.kind == 2) {
                    if (!is_or) { if (val) { val = false; isnull = true; } }
                    else { if (!val) isnull = true; }
                } else if (pk.kind == 1) {
                    const bool c = pk.col[0] != 0;
                    if (!is_or ? !c : c) { val = is_or != 0; isnull = false; }
                } else {
                    const bool a1 = pk.col[i] != 0, null2 = bm_test(pk.nulls, i);
                    // the vector-level "no nulls anywhere" fast path of the reference gives the same per-row result as this branch
                    if (isnull && !null2) {
                        if (!is_or) { if (!a1) { isnull = false; val = false; } }
                        else { if (a1) { isnull = false; val = true; } }
                    } else if (!isnull && null2) {
                        if (!is_or) { if (val) { isnull = true; val = false; } }
                        else { if (!val) isnull = true; }
                    } else if (!isnull && !null2) {
                        val = is_or ? (val || a1) : (val && a1);
                    }
                }
            }
            r[i] = val ? 1 : 0;
        }
        const unsigned m = __ballot_sync(0xffffffffu, in && isnull);
        if ((threadIdx.x & 31) == 0) rnulls32[i >> 5] = m;
    }
    (void)any_null_from;
}

int grid_rows(uint64_t n) {
    uint64_t b = (n + kThreads - 1) / kThreads;
    const uint64_t cap = (uint64_t)num_sms() * 16;
    return (int)(b > cap ? cap : (b ? b : 1));
}

const char *type_name(int T) {
    switch (T) {
    case MO_T_INT8: return "int8"; case MO_T_INT16: return "int16"; case MO_T_INT32: return "int32"; case MO_T_INT64: return "int64";
    case MO_T_UINT8: return "uint8"; case MO_T_UINT16: return "uint16"; case MO_T_UINT32: return "uint32"; case MO_T_UINT64: return "uint64";
    case MO_T_FLOAT32: return "float32"; case MO_T_FLOAT64: return "float64";
    }
    return "?";
}
template <typename T> void fmt_val(char *out, size_t cap, T v) {
    if (std::is_floating_point<T>::value) snprintf(out, cap, "%g", (double)v);
    else if (std::is_signed<T>::value) snprintf(out, cap, "%lld", (long long)v);
    else snprintf(out, cap, "%llu", (unsigned long long)v);
}

struct GoParams { int32_t div0_null; int32_t pad; int64_t err_row; };

template <typename T, int OP>
int run_go_arith(ThreadCtx &t, int Tid, mo_xcall_args_t *args, uint64_t len) {
    const uint64_t nwords = (len + 63) / 64;
    const bool c1 = args[1].dataSz == sizeof(T) && len > 1, c2 = args[2].dataSz == sizeof(T) && len > 1;
    if ((!c1 && args[1].dataSz < sizeof(T) * len) || (!c2 && args[2].dataSz < sizeof(T) * len) || args[0].dataSz < sizeof(T) * len || !args[0].pnulls) {
        set_error("go arithmetic: vectors shorter than len, or the result has no null bitmap"); return MO_RC_INVALID_ARGUMENT;
    }
    GoParams P{0, 0, -1};
    if (args[3].pdata && args[3].dataSz >= sizeof(GoParams)) {
        if (is_device_ptr(args[3].pdata)) { int rc = read_back(t, &P, args[3].pdata, sizeof P); if (rc) return rc; }
        else memcpy(&P, args[3].pdata, sizeof P);
    }
    Stager st(t);
    T *r = (T *)st.out(args[0].pdata, sizeof(T) * len, true);
    uint64_t *rn = (uint64_t *)st.out(args[0].pnulls, nwords * 8, true);
    const T *a = (const T *)st.in(args[1].pdata, c1 ? sizeof(T) : sizeof(T) * len);
    const T *b = (const T *)st.in(args[2].pdata, c2 ? sizeof(T) : sizeof(T) * len);
    const uint64_t *n1 = (const uint64_t *)st.in(args[1].pnulls, args[1].pnulls ? (c1 ? 8 : nwords * 8) : 0);
    const uint64_t *n2 = (const uint64_t *)st.in(args[2].pnulls, args[2].pnulls ? (c2 ? 8 : nwords * 8) : 0);
    unsigned long long *dbad = (unsigned long long *)st.tmp(16);
    if (st.failed) { st.finish(); return MO_RC_INTERNAL_ERROR; }
    // const NULL operand: every row is NULL (baseTemplate.go:600-700)
    uint64_t h1 = 0, h2 = 0;
    if (c1 && n1) { int rc = read_back(t, &h1, n1, 8); if (rc) { st.finish(); return rc; } }
    if (c2 && n2) { int rc = read_back(t, &h2, n2, 8); if (rc) { st.finish(); return rc; } }
    const bool all_null = (c1 && (h1 & 1)) || (c2 && (h2 & 1));
    go_nulls_kernel<<<(unsigned)((nwords + 255) / 256 > 1024 ? 1024 : (nwords + 255) / 256), 256, 0, t.stream>>>(rn, c1 ? nullptr : n1, c2 ? nullptr : n2, len, all_null ? 1 : 0);
    MOB_LAUNCH_CHECK();
    unsigned long long bad = kNoRow;
    if (!all_null) {
        MOB_CUDA_TRY(cudaMemsetAsync(dbad, 0xff, 8, t.stream));
        cudaEventRecord(t.kev0, t.stream);
        go_arith_kernel<T, OP><<<grid_rows(len), kThreads, 0, t.stream>>>(r, a, b, len, c1 ? 1 : 0, c2 ? 1 : 0, rn, P.div0_null, dbad);
        cudaEventRecord(t.kev1, t.stream);
        MOB_LAUNCH_CHECK();
        int rc = read_back(t, &bad, dbad, 8);
        if (rc) { st.finish(); return rc; }
    }
    T xa = 0, xb = 0;
    if (bad != kNoRow) {   // operands of the offending row for the message (moerr.NewOutOfRange(ctx, "int64", "(%d + %d)", v1, v2))
        read_back(t, &xa, a + (c1 ? 0 : bad), sizeof(T));
        read_back(t, &xb, b + (c2 ? 0 : bad), sizeof(T));
    }
    int frc = st.finish();
    if (frc) return frc;
    P.err_row = bad == kNoRow ? -1 : (int64_t)bad;
    if (args[3].pdata && args[3].dataSz >= sizeof(GoParams)) {
        if (is_device_ptr(args[3].pdata)) { MOB_CUDA_TRY(cudaMemcpyAsync(args[3].pdata, &P, sizeof P, cudaMemcpyHostToDevice, t.stream)); MOB_CUDA_TRY(cudaStreamSynchronize(t.stream)); }
        else memcpy(args[3].pdata, &P, sizeof P);
    }
    if (bad == kNoRow) return MO_RC_SUCCESS;
    if (OP == OP_DIV || OP == OP_MOD) { set_error("division by zero"); return MO_RC_DIVISION_BY_ZERO; }
    char sa[48], sb[48];
    fmt_val<T>(sa, sizeof sa, xa); fmt_val<T>(sb, sizeof sb, xb);
    set_error("data out of range: data type %s, value '(%s %c %s)'", type_name(Tid), sa, OP == OP_ADD ? '+' : (OP == OP_SUB ? '-' : '*'), sb);
    return MO_RC_OUT_OF_RANGE;
}

template <typename T>
int go_arith_op(ThreadCtx &t, int op, int Tid, mo_xcall_args_t *args, uint64_t len) {
    switch (op) {
    case OP_ADD: return run_go_arith<T, OP_ADD>(t, Tid, args, len);
    case OP_SUB: return run_go_arith<T, OP_SUB>(t, Tid, args, len);
    case OP_MUL: return run_go_arith<T, OP_MUL>(t, Tid, args, len);
    case OP_DIV:
        if (!std::is_floating_point<T>::value) break;   // "/" is defined on floats only (arithmetic.go:481-514)
        return run_go_arith<T, OP_DIV>(t, Tid, args, len);
    case OP_MOD: return run_go_arith<T, OP_MOD>(t, Tid, args, len);
    }
    set_error("go arithmetic: operator %d is not defined for type %d", op, Tid);
    return MO_RC_INVALID_ARGUMENT;
}

template <typename T>
int run_go_compare(ThreadCtx &t, int op, mo_xcall_args_t *args, uint64_t len, double pw = 0.0) {
    const uint64_t nwords = (len + 63) / 64;
    const bool c1 = args[1].dataSz == sizeof(T) && len > 1, c2 = args[2].dataSz == sizeof(T) && len > 1;
    if ((!c1 && args[1].dataSz < sizeof(T) * len) || (!c2 && args[2].dataSz < sizeof(T) * len) || args[0].dataSz < len || !args[0].pnulls) {
        set_error("go compare: vectors shorter than len, or the result has no null bitmap"); return MO_RC_INVALID_ARGUMENT;
    }
    Stager st(t);
    uint8_t *r = (uint8_t *)st.out(args[0].pdata, len, true);
    uint64_t *rn = (uint64_t *)st.out(args[0].pnulls, nwords * 8, true);
    const T *a = (const T *)st.in(args[1].pdata, c1 ? sizeof(T) : sizeof(T) * len);
    const T *b = (const T *)st.in(args[2].pdata, c2 ? sizeof(T) : sizeof(T) * len);
    const uint64_t *n1 = (const uint64_t *)st.in(args[1].pnulls, args[1].pnulls ? (c1 ? 8 : nwords * 8) : 0);
    const uint64_t *n2 = (const uint64_t *)st.in(args[2].pnulls, args[2].pnulls ? (c2 ? 8 : nwords * 8) : 0);
    if (st.failed) { st.finish(); return MO_RC_INTERNAL_ERROR; }
    uint64_t h1 = 0, h2 = 0;
    if (c1 && n1) { int rc = read_back(t, &h1, n1, 8); if (rc) { st.finish(); return rc; } }
    if (c2 && n2) { int rc = read_back(t, &h2, n2, 8); if (rc) { st.finish(); return rc; } }
    const bool all_null = (c1 && (h1 & 1)) || (c2 && (h2 & 1));
    go_nulls_kernel<<<(unsigned)((nwords + 255) / 256 > 1024 ? 1024 : (nwords + 255) / 256), 256, 0, t.stream>>>(rn, c1 ? nullptr : n1, c2 ? nullptr : n2, len, all_null ? 1 : 0);
    MOB_LAUNCH_CHECK();
    if (!all_null) {
        cudaEventRecord(t.kev0, t.stream);
        const int vec_ok = ((((uintptr_t)r) | ((uintptr_t)a) | ((uintptr_t)b)) & 15) == 0 ? 1 : 0;
        go_compare_kernel<T><<<grid_rows(vec_ok ? (len + 7) / 8 : len), kThreads, 0, t.stream>>>(r, a, b, len, c1 ? 1 : 0, c2 ? 1 : 0, rn, op, vec_ok, pw);
        cudaEventRecord(t.kev1, t.stream);
        MOB_LAUNCH_CHECK();
    }
    return st.finish();
}

template <typename T>
int run_go_between(ThreadCtx &t, mo_xcall_args_t *args, uint64_t len) {
    const uint64_t nwords = (len + 63) / 64;
    if (args[1].dataSz < sizeof(T) * len || args[2].dataSz < sizeof(T) || args[3].dataSz < sizeof(T) || args[0].dataSz < len || !args[0].pnulls) {
        set_error("go between: vectors shorter than len, bounds missing, or the result has no null bitmap"); return MO_RC_INVALID_ARGUMENT;
    }
    T lo, hi;
    if (is_device_ptr(args[2].pdata)) { int rc = read_back(t, &lo, args[2].pdata, sizeof(T)); if (rc) return rc; } else memcpy(&lo, args[2].pdata, sizeof(T));
    if (is_device_ptr(args[3].pdata)) { int rc = read_back(t, &hi, args[3].pdata, sizeof(T)); if (rc) return rc; } else memcpy(&hi, args[3].pdata, sizeof(T));
    Stager st(t);
    uint8_t *r = (uint8_t *)st.out(args[0].pdata, len, false);
    uint64_t *rn = (uint64_t *)st.out(args[0].pnulls, nwords * 8, true);
    const T *col = (const T *)st.in(args[1].pdata, sizeof(T) * len);
    const uint64_t *nulls = (const uint64_t *)st.in(args[1].pnulls, args[1].pnulls ? nwords * 8 : 0);
    if (st.failed) { st.finish(); return MO_RC_INTERNAL_ERROR; }
    cudaEventRecord(t.kev0, t.stream);
    go_between_kernel<T><<<grid_rows(len), kThreads, 0, t.stream>>>(r, col, lo, hi, len, nulls, reinterpret_cast<uint32_t *>(rn));
    cudaEventRecord(t.kev1, t.stream);
    MOB_LAUNCH_CHECK();
    return st.finish();
}

int run_go_multi_logic(ThreadCtx &t, int is_or, mo_xcall_args_t *args, uint64_t len) {
    const uint64_t nwords = (len + 63) / 64;
    int32_t nparams = 0;
    if (!args[1].pdata || args[1].dataSz < 4) { set_error("go logic: parameter count missing"); return MO_RC_INVALID_ARGUMENT; }
    if (is_device_ptr(args[1].pdata)) { int rc = read_back(t, &nparams, args[1].pdata, 4); if (rc) return rc; } else memcpy(&nparams, args[1].pdata, 4);
    if (nparams < 1 || nparams > kMaxLogicParams) { set_error("go logic: 1..%d operands", kMaxLogicParams); return MO_RC_INVALID_ARGUMENT; }
    if (args[0].dataSz < len || !args[0].pnulls) { set_error("go logic: result shorter than len, or without a null bitmap"); return MO_RC_INVALID_ARGUMENT; }
    Stager st(t);
    uint8_t *r = (uint8_t *)st.out(args[0].pdata, len, false);
    uint64_t *rn = (uint64_t *)st.out(args[0].pnulls, nwords * 8, false);
    LogicParams P; P.n = nparams;
    for (int k = 0; k < nparams; k++) {
        const mo_xcall_args_t &a = args[2 + k];
        const bool cst = a.dataSz == 1 && len > 1;
        if (!cst && a.dataSz < len) { st.finish(); set_error("go logic: operand %d shorter than len", k); return MO_RC_INVALID_ARGUMENT; }
        P.p[k].col = (const uint8_t *)st.in(a.pdata, cst ? 1 : len);
        P.p[k].nulls = (const uint64_t *)st.in(a.pnulls, a.pnulls ? (cst ? 8 : nwords * 8) : 0);
        P.p[k].kind = cst ? 1 : 0;
        if (cst && P.p[k].nulls) {
            uint64_t h = 0;
            int rc = read_back(t, &h, P.p[k].nulls, 8);
            if (rc) { st.finish(); return rc; }
            if (h & 1) P.p[k].kind = 2;
        }
    }
    if (st.failed) { st.finish(); return MO_RC_INTERNAL_ERROR; }
    MOB_CUDA_TRY(cudaMemsetAsync(rn, 0, nwords * 8, t.stream));
    cudaEventRecord(t.kev0, t.stream);
    go_multi_logic_kernel<<<grid_rows(len), kThreads, 0, t.stream>>>(r, reinterpret_cast<uint32_t *>(rn), len, P, is_or, 0);
    cudaEventRecord(t.kev1, t.stream);
    MOB_LAUNCH_CHECK();
    return st.finish();
}

}  // namespace

namespace mob {

int xcall_go_elementwise(int64_t funcId, mo_xcall_args_t *args, uint64_t len) {
    ThreadCtx &t = tctx();
    if (!t.ready) return MO_RC_INTERNAL_ERROR;
    if (len == 0) return MO_RC_SUCCESS;
    if (funcId >= 0x4000 && funcId < 0x4800) {
        const int op = (int)((funcId - 0x4000) >> 8), T = (int)(funcId & 0xff);
        switch (T) {
        case MO_T_INT8: return go_arith_op<int8_t>(t, op, T, args, len);
        case MO_T_INT16: return go_arith_op<int16_t>(t, op, T, args, len);
        case MO_T_INT32: return go_arith_op<int32_t>(t, op, T, args, len);
        case MO_T_INT64: return go_arith_op<int64_t>(t, op, T, args, len);
        case MO_T_UINT8: return go_arith_op<uint8_t>(t, op, T, args, len);
        case MO_T_UINT16: return go_arith_op<uint16_t>(t, op, T, args, len);
        case MO_T_UINT32: return go_arith_op<uint32_t>(t, op, T, args, len);
        case MO_T_UINT64: return go_arith_op<uint64_t>(t, op, T, args, len);
        case MO_T_FLOAT32: return go_arith_op<float>(t, op, T, args, len);
        case MO_T_FLOAT64: return go_arith_op<double>(t, op, T, args, len);
        }
        set_error("go arithmetic: unsupported type %d", T);
        return MO_RC_INVALID_ARGUMENT;
    }
    if (funcId >= 0x4800 && funcId < 0x5000) {
        const int op = (int)((funcId - 0x4800) >> 8), T = (int)(funcId & 0xff);
        if (op > CMP_LE) { set_error("go compare: unknown operator %d", op); return MO_RC_INVALID_ARGUMENT; }
        switch (T) {   // DATE compares as int32, TIME / DATETIME / TIMESTAMP as int64 (types.go:187-191); BOOL via false < true
        case MO_T_BOOL: case MO_T_UINT8: return run_go_compare<uint8_t>(t, op, args, len);
        case MO_T_INT8: return run_go_compare<int8_t>(t, op, args, len);
        case MO_T_INT16: return run_go_compare<int16_t>(t, op, args, len);
        case MO_T_INT32: case MO_T_DATE: return run_go_compare<int32_t>(t, op, args, len);
        case MO_T_INT64: case MO_T_TIME: case MO_T_DATETIME: case MO_T_TIMESTAMP: return run_go_compare<int64_t>(t, op, args, len);
        case MO_T_UINT16: return run_go_compare<uint16_t>(t, op, args, len);
        case MO_T_UINT32: return run_go_compare<uint32_t>(t, op, args, len);
        case MO_T_UINT64: return run_go_compare<uint64_t>(t, op, args, len);
        case MO_T_FLOAT32: return run_go_compare<float>(t, op, args, len);
        case MO_T_FLOAT64: return run_go_compare<double>(t, op, args, len);
        }
        set_error("go compare: unsupported type %d", T);
        return MO_RC_INVALID_ARGUMENT;
    }
    if (funcId >= 0x5200 && funcId < 0x5800) {   // MO_XCALL_GO_COMPARE_F32_SCALE(op, scale)
        const int op = (int)((funcId - 0x5200) >> 8), scale = (int)(funcId & 0xff);
        if (op > CMP_LE || scale < 1 || scale > 22) { set_error("go compare f32 scale: operator 0..5, scale 1..22"); return MO_RC_INVALID_ARGUMENT; }
        double pw = 1.0;
        for (int k = 0; k < scale; k++) pw *= 10.0;   // math.Pow10(scale), exact
        return run_go_compare<float>(t, op, args, len, pw);
    }
    if (funcId >= 0x5000 && funcId < 0x5100) {
        const int T = (int)(funcId & 0xff);
        switch (T) {
        case MO_T_INT8: return run_go_between<int8_t>(t, args, len);
        case MO_T_INT16: return run_go_between<int16_t>(t, args, len);
        case MO_T_INT32: case MO_T_DATE: return run_go_between<int32_t>(t, args, len);
        case MO_T_INT64: case MO_T_TIME: case MO_T_DATETIME: case MO_T_TIMESTAMP: return run_go_between<int64_t>(t, args, len);
        case MO_T_UINT8: return run_go_between<uint8_t>(t, args, len);
        case MO_T_UINT16: return run_go_between<uint16_t>(t, args, len);
        case MO_T_UINT32: return run_go_between<uint32_t>(t, args, len);
        case MO_T_UINT64: return run_go_between<uint64_t>(t, args, len);
        case MO_T_FLOAT32: return run_go_between<float>(t, args, len);
        case MO_T_FLOAT64: return run_go_between<double>(t, args, len);
        }
        set_error("go between: unsupported type %d", T);
        return MO_RC_INVALID_ARGUMENT;
    }
    if (funcId == 0x5100 || funcId == 0x5101) return run_go_multi_logic(t, (int)(funcId & 1), args, len);
    return -1;
}

}  // namespace mob
