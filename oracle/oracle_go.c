/*
 * oracle_go.c -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or called by the product
 * (matrixone_b200/, libmo_b200.so).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.
 *
 * A plain-C restatement of the Go batch loops MatrixOne runs per 8192-row block on the hot path
 * SURVEY.md section 8 names.  The Go toolchain is absent in the build container, so the Go side cannot be
 * run; each function below follows one reference function and cites it (paths relative to
 * /root/reference).  Build WITHOUT -ffast-math and with -ffp-contract=off: Go on amd64 evaluates
 * floating point exactly as written (no reassociation, no FMA fusion at GOAMD64=v1).
 *
 * Pinning: tests/test_oracle_golden.py checks these functions against the known-answer tables of the
 * reference's own Go tests (transcribed by tests/golden/extract_goldens.py into tests/golden/*.json).
 *
 * The C half of the reference (cgo/{mo,arith,compare,logic,xcall}.c) is NOT restated: it is compiled
 * unchanged from /root/reference into oracle/_ref/libmo_ref.so by oracle/build.py.
 */
#define _GNU_SOURCE
#include <stdint.h>
#include <stdbool.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <float.h>
#include <pthread.h>

#define OG_RC_OK 0
#define OG_RC_DIV_ZERO 20200     /* moerr ErrDivByZero, cgo/mo_impl.h:32 */
#define OG_RC_OUT_OF_RANGE 20201 /* moerr ErrOutOfRange, cgo/mo_impl.h:33 */
#define OG_RC_INVALID 20203
#define OG_RC_DIM_MISMATCH 20204

/* types.T ids, pkg/container/types/types.go:35-67 */
enum { T_bool = 10, T_int8 = 20, T_int16, T_int32, T_int64, T_uint8 = 25, T_uint16, T_uint32, T_uint64,
       T_float32 = 30, T_float64 = 31, T_date = 50, T_time, T_datetime, T_timestamp };

enum { OP_ADD = 0, OP_SUB, OP_MUL, OP_DIV, OP_MOD };
enum { CMP_EQ = 0, CMP_NE, CMP_GT, CMP_GE, CMP_LT, CMP_LE };

/* ---------------------------------------------------------------------------------------------
 * bitmap: pkg/common/bitmap/bitmap.go:196-229 (Add/Contains), LSB-first uint64 words, set = NULL
 * ------------------------------------------------------------------------------------------- */

/* ---------------------------------------------------------------------------------------------
 * Persistent worker pool for the multi-threaded pipelines below (the reference keeps one pipeline goroutine per core
 * alive for the query, pkg/sql/compile/scope.go:442-499; creating and joining 128 pthreads per call made the timed CPU
 * arm noisy).  Job j always runs on worker j % nworkers and worker w is pinned to the w-th CPU this process may use, so
 * the rows a worker generated (first touch) are the rows it later scans: NUMA-local without libnuma.
 * ------------------------------------------------------------------------------------------- */
#include <sched.h>
typedef void *(*og_fn)(void *);
static struct {
    pthread_mutex_t mu; pthread_cond_t start, done;
    pthread_t *th; int nth;
    og_fn fn; char *jobs; size_t jobsz; int njobs;
    uint64_t gen; int remaining;
    cpu_set_t allowed; int nallowed; int have_allowed;
} POOL = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER};
static pthread_mutex_t POOL_RUN = PTHREAD_MUTEX_INITIALIZER;

static void *pool_main(void *arg) {
    /* arg = (id, generation current when the worker was created): the first generation published after that is its first job */
    const int id = (int)((uint64_t *)arg)[0];
    uint64_t seen_gen = ((uint64_t *)arg)[1];
    free(arg);
    if (POOL.have_allowed && POOL.nallowed > 0) {
        int want = id % POOL.nallowed, seen = 0;
        for (int c = 0; c < CPU_SETSIZE; c++) {
            if (!CPU_ISSET(c, &POOL.allowed)) continue;
            if (seen++ == want) { cpu_set_t one; CPU_ZERO(&one); CPU_SET(c, &one); pthread_setaffinity_np(pthread_self(), sizeof one, &one); break; }
        }
    }
    for (;;) {
        pthread_mutex_lock(&POOL.mu);
        while (POOL.gen == seen_gen) pthread_cond_wait(&POOL.start, &POOL.mu);
        seen_gen = POOL.gen;
        og_fn fn = POOL.fn; char *jobs = POOL.jobs; size_t jobsz = POOL.jobsz; int njobs = POOL.njobs, nth = POOL.nth;
        pthread_mutex_unlock(&POOL.mu);
        for (int j = id; j < njobs; j += nth) fn(jobs + (size_t)j * jobsz);
        pthread_mutex_lock(&POOL.mu);
        if (--POOL.remaining == 0) pthread_cond_signal(&POOL.done);
        pthread_mutex_unlock(&POOL.mu);
    }
    return NULL;
}

/* run fn over njobs job records (jobsz bytes apart) on the pool; returns when all are done */
static void og_run(og_fn fn, void *jobs, size_t jobsz, int njobs) {
    if (njobs <= 0) return;
    if (njobs == 1) { fn(jobs); return; }
    pthread_mutex_lock(&POOL_RUN);
    pthread_mutex_lock(&POOL.mu);
    if (!POOL.have_allowed) {
        CPU_ZERO(&POOL.allowed);
        if (sched_getaffinity(0, sizeof POOL.allowed, &POOL.allowed) == 0) POOL.nallowed = CPU_COUNT(&POOL.allowed);
        POOL.have_allowed = 1;
    }
    if (POOL.nth < njobs) {   /* grow: worker ids are stable, so the job -> CPU map stays fixed */
        POOL.th = realloc(POOL.th, sizeof(pthread_t) * (size_t)njobs);
        for (int t = POOL.nth; t < njobs; t++) {
            uint64_t *a = malloc(16); a[0] = (uint64_t)t; a[1] = POOL.gen;
            pthread_create(&POOL.th[t], NULL, pool_main, a);
        }
        POOL.nth = njobs;
    }
    POOL.fn = fn; POOL.jobs = jobs; POOL.jobsz = jobsz; POOL.njobs = njobs;
    POOL.remaining = POOL.nth;
    POOL.gen++;
    pthread_cond_broadcast(&POOL.start);
    while (POOL.remaining) pthread_cond_wait(&POOL.done, &POOL.mu);
    pthread_mutex_unlock(&POOL.mu);
    pthread_mutex_unlock(&POOL_RUN);
}

static inline bool bm_has(const uint64_t *p, uint64_t i) { return p && ((p[i >> 6] >> (i & 63)) & 1); }
static inline void bm_add(uint64_t *p, uint64_t i) { p[i >> 6] |= (uint64_t)1 << (i & 63); }
static inline void bm_del(uint64_t *p, uint64_t i) { p[i >> 6] &= ~((uint64_t)1 << (i & 63)); }
static inline uint64_t bm_words(uint64_t n) { return (n + 63) >> 6; }

static bool bm_any(const uint64_t *p, uint64_t n) {
    if (!p) return false;
    for (uint64_t w = 0; w < bm_words(n); w++) {
        uint64_t v = p[w];
        if (w == bm_words(n) - 1 && (n & 63)) v &= (((uint64_t)1 << (n & 63)) - 1);
        if (v) return true;
    }
    return false;
}

/* ---------------------------------------------------------------------------------------------
 * Overflow-checked scalar arithmetic: pkg/sql/plan/function/arithmetic_overflow_check.go:29-310.
 * Go integer arithmetic wraps; the checks below are exact overflow detectors.
 * ------------------------------------------------------------------------------------------- */
#define DEF_SIGNED(NAME, T, UT, TMIN)                                                              \
    static inline bool add_##NAME(T a, T b, T *r) {                                                \
        T s = (T)((UT)a + (UT)b); *r = s;                                                          \
        return (a > 0 && b > 0 && s <= 0) || (a < 0 && b < 0 && s >= 0); /* :32-38 */              \
    }                                                                                              \
    static inline bool sub_##NAME(T a, T b, T *r) {                                                \
        T s = (T)((UT)a - (UT)b); *r = s;                                                          \
        return (a > 0 && b < 0 && s < 0) || (a < 0 && b > 0 && s > 0); /* :104-112 */              \
    }                                                                                              \
    static inline bool mul_##NAME(T a, T b, T *r) {                                                \
        if (a == 0 || b == 0) { *r = 0; return false; }                                            \
        if ((a == TMIN && b == -1) || (b == TMIN && a == -1)) { *r = 0; return true; } /* :183 */  \
        T s = (T)((UT)a * (UT)b); *r = s;                                                          \
        return (T)(s / b) != a; /* :189 */                                                         \
    }
#define DEF_UNSIGNED(NAME, T, TMAX)                                                                \
    static inline bool add_##NAME(T a, T b, T *r) { T s = (T)(a + b); *r = s; return s < a || s < b; }   \
    static inline bool sub_##NAME(T a, T b, T *r) { *r = (T)(a - b); return a < b; }               \
    static inline bool mul_##NAME(T a, T b, T *r) {                                                \
        if (a == 0 || b == 0) { *r = 0; return false; }                                            \
        *r = (T)(a * b); return a > (T)(TMAX / b); /* :262 */                                      \
    }
DEF_SIGNED(i8, int8_t, uint8_t, INT8_MIN)
DEF_SIGNED(i16, int16_t, uint16_t, INT16_MIN)
DEF_SIGNED(i32, int32_t, uint32_t, INT32_MIN)
DEF_SIGNED(i64, int64_t, uint64_t, INT64_MIN)
DEF_UNSIGNED(u8, uint8_t, UINT8_MAX)
DEF_UNSIGNED(u16, uint16_t, UINT16_MAX)
DEF_UNSIGNED(u32, uint32_t, UINT32_MAX)
DEF_UNSIGNED(u64, uint64_t, UINT64_MAX)
/* floats: add/sub error on +-Inf result (:295-310); mul unchecked (arithmetic.go:444-451) */
static inline bool add_f32(float a, float b, float *r) { *r = a + b; return isinf(*r); }
static inline bool sub_f32(float a, float b, float *r) { *r = a - b; return isinf(*r); }
static inline bool mul_f32(float a, float b, float *r) { *r = a * b; return false; }
static inline bool add_f64(double a, double b, double *r) { *r = a + b; return isinf(*r); }
static inline bool sub_f64(double a, double b, double *r) { *r = a - b; return isinf(*r); }
static inline bool mul_f64(double a, double b, double *r) { *r = a * b; return false; }

/*
 * og_arith: opBinaryFixedFixedToFixedWithErrorCheck, pkg/sql/plan/function/baseTemplate.go:580-728,
 * instantiated by plusFn/minusFn/multiFn (arithmetic.go:222-469).
 *   c1/c2    : operand is a const vector (length-1 data)
 *   n1/n2    : operand nulls bitmaps (NULL = none); a const-null operand is (cK && nK bit0 set)
 *   rnulls   : result nulls, pre-filled by the caller with NOT selectList (baseTemplate.go:473-486);
 *              on return holds rnulls | n1 | n2.  Must be non-NULL (bm_words(n) words).
 *   Rows whose result-null bit is set are skipped (r[i] untouched).
 *   Returns OG_RC_OUT_OF_RANGE at the FIRST offending non-null row (row index in *err_row); rows before it
 *   are written, rows after are not -- exactly what the Go loop leaves behind.
 *   OP_DIV/OP_MOD (floats, ints for MOD): division by zero -> row becomes NULL when div0_null, else
 *   OG_RC_DIV_ZERO (specialTemplateForDivFunction baseTemplate.go:1436, checkDivisionByZeroBehavior :1369).
 */
#define ARITH_LOOP(T, FN)                                                                          \
    do {                                                                                           \
        const T *at = (const T *)a, *bt = (const T *)b; T *rt = (T *)r;                            \
        for (uint64_t i = 0; i < n; i++) {                                                         \
            if (bm_has(rnulls, i)) continue;                                                       \
            T v;                                                                                   \
            if (FN(at[c1 ? 0 : i], bt[c2 ? 0 : i], &v)) { if (err_row) *err_row = (int64_t)i; return OG_RC_OUT_OF_RANGE; } \
            rt[i] = v;                                                                             \
        }                                                                                          \
        return OG_RC_OK;                                                                           \
    } while (0)

#define DIVMOD_LOOP(T, EXPR)                                                                       \
    do {                                                                                           \
        const T *at = (const T *)a, *bt = (const T *)b; T *rt = (T *)r;                            \
        for (uint64_t i = 0; i < n; i++) {                                                         \
            if (bm_has(rnulls, i)) continue;                                                       \
            T x = at[c1 ? 0 : i], y = bt[c2 ? 0 : i];                                              \
            if (y == 0) {                                                                          \
                if (div0_null) { bm_add(rnulls, i); continue; }                                    \
                if (err_row) *err_row = (int64_t)i;                                                \
                return OG_RC_DIV_ZERO;                                                             \
            }                                                                                      \
            rt[i] = (EXPR);                                                                        \
        }                                                                                          \
        return OG_RC_OK;                                                                           \
    } while (0)

int32_t og_arith(int32_t op, int32_t type, void *r, const void *a, const void *b, uint64_t n,
                 int32_t c1, int32_t c2, const uint64_t *n1, const uint64_t *n2, uint64_t *rnulls,
                 int32_t div0_null, int64_t *err_row) {
    /* null propagation, baseTemplate.go:600-700: const-null operand nulls every row */
    if ((c1 && bm_has(n1, 0)) || (c2 && bm_has(n2, 0))) {
        for (uint64_t i = 0; i < n; i++) bm_add(rnulls, i);
        return OG_RC_OK;
    }
    for (uint64_t w = 0; w < bm_words(n); w++) {
        if (!c1 && n1) rnulls[w] |= n1[w];
        if (!c2 && n2) rnulls[w] |= n2[w];
    }
    if (n & 63) rnulls[bm_words(n) - 1] &= (((uint64_t)1 << (n & 63)) - 1);
    switch (op) {
    case OP_ADD:
        switch (type) {
        case T_int8: ARITH_LOOP(int8_t, add_i8); case T_int16: ARITH_LOOP(int16_t, add_i16);
        case T_int32: ARITH_LOOP(int32_t, add_i32); case T_int64: ARITH_LOOP(int64_t, add_i64);
        case T_uint8: ARITH_LOOP(uint8_t, add_u8); case T_uint16: ARITH_LOOP(uint16_t, add_u16);
        case T_uint32: ARITH_LOOP(uint32_t, add_u32); case T_uint64: ARITH_LOOP(uint64_t, add_u64);
        case T_float32: ARITH_LOOP(float, add_f32); case T_float64: ARITH_LOOP(double, add_f64);
        } break;
    case OP_SUB:
        switch (type) {
        case T_int8: ARITH_LOOP(int8_t, sub_i8); case T_int16: ARITH_LOOP(int16_t, sub_i16);
        case T_int32: ARITH_LOOP(int32_t, sub_i32); case T_int64: ARITH_LOOP(int64_t, sub_i64);
        case T_uint8: ARITH_LOOP(uint8_t, sub_u8); case T_uint16: ARITH_LOOP(uint16_t, sub_u16);
        case T_uint32: ARITH_LOOP(uint32_t, sub_u32); case T_uint64: ARITH_LOOP(uint64_t, sub_u64);
        case T_float32: ARITH_LOOP(float, sub_f32); case T_float64: ARITH_LOOP(double, sub_f64);
        } break;
    case OP_MUL:
        switch (type) {
        case T_int8: ARITH_LOOP(int8_t, mul_i8); case T_int16: ARITH_LOOP(int16_t, mul_i16);
        case T_int32: ARITH_LOOP(int32_t, mul_i32); case T_int64: ARITH_LOOP(int64_t, mul_i64);
        case T_uint8: ARITH_LOOP(uint8_t, mul_u8); case T_uint16: ARITH_LOOP(uint16_t, mul_u16);
        case T_uint32: ARITH_LOOP(uint32_t, mul_u32); case T_uint64: ARITH_LOOP(uint64_t, mul_u64);
        case T_float32: ARITH_LOOP(float, mul_f32); case T_float64: ARITH_LOOP(double, mul_f64);
        } break;
    case OP_DIV: /* arithmetic.go:481-514: float "/" only */
        switch (type) {
        case T_float32: DIVMOD_LOOP(float, x / y); case T_float64: DIVMOD_LOOP(double, x / y);
        } break;
    case OP_MOD: /* arithmetic.go:704-762: ints use %, floats math.Mod */
        switch (type) {
        case T_int8: DIVMOD_LOOP(int8_t, (int8_t)(y == -1 ? 0 : x % y)); case T_int16: DIVMOD_LOOP(int16_t, (int16_t)(y == -1 ? 0 : x % y));
        case T_int32: DIVMOD_LOOP(int32_t, y == -1 ? 0 : x % y); case T_int64: DIVMOD_LOOP(int64_t, y == -1 ? 0 : x % y);
        case T_uint8: DIVMOD_LOOP(uint8_t, (uint8_t)(x % y)); case T_uint16: DIVMOD_LOOP(uint16_t, (uint16_t)(x % y));
        case T_uint32: DIVMOD_LOOP(uint32_t, x % y); case T_uint64: DIVMOD_LOOP(uint64_t, x % y);
        case T_float32: DIVMOD_LOOP(float, (float)fmod((double)x, (double)y)); case T_float64: DIVMOD_LOOP(double, fmod(x, y));
        } break;
    }
    return OG_RC_INVALID;
}

/*
 * og_compare: equalFn/greatThanFn/... pkg/sql/plan/function/func_compare.go:285,677,804,931,1058,1185
 * dispatching into opBinaryFixedFixedToFixed (baseTemplate.go:457-578).  Null rows are skipped.
 * DATE compares as int32, TIME/DATETIME/TIMESTAMP as int64 (types.go:187-191); BOOL via false<true.
 */
#define CMP_LOOP(T)                                                                                \
    do {                                                                                           \
        const T *at = (const T *)a, *bt = (const T *)b;                                            \
        for (uint64_t i = 0; i < n; i++) {                                                         \
            if (bm_has(rnulls, i)) continue;                                                       \
            T x = at[c1 ? 0 : i], y = bt[c2 ? 0 : i]; bool v;                                      \
            switch (op) {                                                                          \
            case CMP_EQ: v = x == y; break; case CMP_NE: v = x != y; break;                        \
            case CMP_GT: v = x > y; break;  case CMP_GE: v = x >= y; break;                        \
            case CMP_LT: v = x < y; break;  default: v = x <= y; break;                            \
            }                                                                                      \
            r[i] = v;                                                                              \
        }                                                                                          \
        return OG_RC_OK;                                                                           \
    } while (0)

int32_t og_compare(int32_t op, int32_t type, bool *r, const void *a, const void *b, uint64_t n,
                   int32_t c1, int32_t c2, const uint64_t *n1, const uint64_t *n2, uint64_t *rnulls) {
    if ((c1 && bm_has(n1, 0)) || (c2 && bm_has(n2, 0))) {
        for (uint64_t i = 0; i < n; i++) bm_add(rnulls, i);
        return OG_RC_OK;
    }
    for (uint64_t w = 0; w < bm_words(n); w++) {
        if (!c1 && n1) rnulls[w] |= n1[w];
        if (!c2 && n2) rnulls[w] |= n2[w];
    }
    if (n & 63) rnulls[bm_words(n) - 1] &= (((uint64_t)1 << (n & 63)) - 1);
    switch (type) {
    case T_bool: case T_uint8: CMP_LOOP(uint8_t);
    case T_int8: CMP_LOOP(int8_t); case T_int16: CMP_LOOP(int16_t);
    case T_int32: case T_date: CMP_LOOP(int32_t);
    case T_int64: case T_time: case T_datetime: case T_timestamp: CMP_LOOP(int64_t);
    case T_uint16: CMP_LOOP(uint16_t); case T_uint32: CMP_LOOP(uint32_t); case T_uint64: CMP_LOOP(uint64_t);
    case T_float32: CMP_LOOP(float); case T_float64: CMP_LOOP(double);
    }
    return OG_RC_INVALID;
}

/* og_compare_f32_scale: the float32 branch of the six compare functions when the column type carries scale > 0
 * (func_compare.go:207-216 eq, :725-734 gt, :852-861 ge, :979-988 ne, :1106-1115 lt, :1233-1242 le): both sides are rounded to
 * `scale` decimals first -- a = float32(math.Round(float64(a)*pow) / pow), pow = math.Pow10(scale); math.Round rounds half away
 * from zero = C round(). */
int32_t og_compare_f32_scale(int32_t op, int32_t scale, bool *r, const float *a, const float *b, uint64_t n,
                             int32_t c1, int32_t c2, const uint64_t *n1, const uint64_t *n2, uint64_t *rnulls) {
    if ((c1 && bm_has(n1, 0)) || (c2 && bm_has(n2, 0))) {
        for (uint64_t i = 0; i < n; i++) bm_add(rnulls, i);
        return OG_RC_OK;
    }
    for (uint64_t w = 0; w < bm_words(n); w++) {
        if (!c1 && n1) rnulls[w] |= n1[w];
        if (!c2 && n2) rnulls[w] |= n2[w];
    }
    if (n & 63) rnulls[bm_words(n) - 1] &= (((uint64_t)1 << (n & 63)) - 1);
    double pw = 1.0;
    for (int k = 0; k < scale; k++) pw *= 10.0;   /* math.Pow10: exact for scale <= 22 */
    for (uint64_t i = 0; i < n; i++) {
        if (bm_has(rnulls, i)) continue;
        float x = a[c1 ? 0 : i], y = b[c2 ? 0 : i]; bool v;
        x = (float)(round((double)x * pw) / pw);
        y = (float)(round((double)y * pw) / pw);
        switch (op) {
        case CMP_EQ: v = x == y; break; case CMP_NE: v = x != y; break;
        case CMP_GT: v = x > y; break;  case CMP_GE: v = x >= y; break;
        case CMP_LT: v = x < y; break;  default: v = x <= y; break;
        }
        r[i] = v;
    }
    return OG_RC_OK;
}

/*
 * og_between: opBetweenFixed, pkg/sql/plan/function/operator_between.go:138-199 (non-const column).
 * null input -> res=false AND null bit set; the sorted fast path gives the same values as the plain loop.
 */
#define BETWEEN_LOOP(T)                                                                            \
    do {                                                                                           \
        const T *c = (const T *)col; T l = *(const T *)lo, h = *(const T *)hi;                     \
        for (uint64_t i = 0; i < n; i++) {                                                         \
            if (bm_has(nulls, i)) { r[i] = false; bm_add(rnulls, i); }                             \
            else r[i] = c[i] >= l && c[i] <= h;                                                    \
        }                                                                                          \
        return OG_RC_OK;                                                                           \
    } while (0)

int32_t og_between(int32_t type, bool *r, const void *col, const void *lo, const void *hi, uint64_t n,
                   const uint64_t *nulls, uint64_t *rnulls) {
    switch (type) {
    case T_int8: BETWEEN_LOOP(int8_t); case T_int16: BETWEEN_LOOP(int16_t);
    case T_int32: case T_date: BETWEEN_LOOP(int32_t);
    case T_int64: case T_time: case T_datetime: case T_timestamp: BETWEEN_LOOP(int64_t);
    case T_uint8: BETWEEN_LOOP(uint8_t); case T_uint16: BETWEEN_LOOP(uint16_t);
    case T_uint32: BETWEEN_LOOP(uint32_t); case T_uint64: BETWEEN_LOOP(uint64_t);
    case T_float32: BETWEEN_LOOP(float); case T_float64: BETWEEN_LOOP(double);
    }
    return OG_RC_INVALID;
}

/*
 * og_multi_and / og_multi_or: opMultiAnd / opMultiOr, pkg/sql/plan/function/logicalOperator.go:36-102,
 * 104-168.  n-ary three-valued fold, left to right.  kind[k]: 0 = flat vector, 1 = const, 2 = const NULL.
 * r and rnulls are outputs (rnulls must hold bm_words(n) zeroed words).
 */
int32_t og_multi_logic(int32_t is_or, bool *r, uint64_t *rnulls, int32_t nparams, const bool *const *cols,
                       const uint64_t *const *nulls, const int32_t *kind, uint64_t n) {
    const bool *a0 = cols[0];
    if (kind[0] == 2) { for (uint64_t i = 0; i < n; i++) { r[i] = false; bm_add(rnulls, i); } }
    else if (kind[0] == 1) { for (uint64_t i = 0; i < n; i++) r[i] = a0[0]; }
    else {
        memcpy(r, a0, n);
        if (nulls[0]) memcpy(rnulls, nulls[0], bm_words(n) * 8);
    }
    for (int k = 1; k < nparams; k++) {
        const bool *a1 = cols[k];
        if (kind[k] == 2) {
            for (uint64_t i = 0; i < n; i++) {
                if (!is_or) { if (r[i]) { r[i] = false; bm_add(rnulls, i); } }
                else { if (!r[i]) bm_add(rnulls, i); }
            }
        } else if (kind[k] == 1) {
            if (!is_or ? !a1[0] : a1[0]) {
                for (uint64_t i = 0; i < n; i++) { r[i] = is_or ? true : false; bm_del(rnulls, i); }
            }
        } else if (bm_any(rnulls, n) || bm_any(nulls[k], n)) {
            for (uint64_t i = 0; i < n; i++) {
                bool null1 = bm_has(rnulls, i), null2 = bm_has(nulls[k], i);
                if (null1 && !null2) {
                    if (!is_or) { if (!a1[i]) { bm_del(rnulls, i); r[i] = false; } }
                    else { if (a1[i]) { bm_del(rnulls, i); r[i] = true; } }
                } else if (!null1 && null2) {
                    if (!is_or) { if (r[i]) { bm_add(rnulls, i); r[i] = false; } }
                    else { if (!r[i]) bm_add(rnulls, i); }
                } else if (!null1 && !null2) {
                    r[i] = is_or ? (r[i] || a1[i]) : (r[i] && a1[i]);
                }
            }
        } else {
            for (uint64_t i = 0; i < n; i++) r[i] = is_or ? (r[i] || a1[i]) : (r[i] && a1[i]);
        }
    }
    return OG_RC_OK;
}

/*
 * og_filter_sels: Filter.Call inner loop, pkg/sql/colexec/filter/filter.go:125-141 -- rows with
 * (!null && true) in row order.  Returns the count.
 */
int64_t og_filter_sels(const bool *v, const uint64_t *nulls, uint64_t n, int64_t *sels) {
    int64_t k = 0;
    for (uint64_t j = 0; j < n; j++)
        if (!bm_has(nulls, j) && v[j]) sels[k++] = (int64_t)j;
    return k;
}

/* og_shuffle_fixed: shuffle.FixedLengthShuffle, pkg/vectorize/shuffle/shuffle.go:21-26 (ws[i]=vs[sels[i]]) */
void og_shuffle_fixed(void *dst, const void *src, const int64_t *sels, int64_t nsel, int32_t szof) {
    for (int64_t i = 0; i < nsel; i++)
        memcpy((char *)dst + (size_t)i * szof, (const char *)src + (size_t)sels[i] * szof, szof);
}

/* og_nulls_filter: nulls.Filter (negate=false), pkg/container/nulls/nulls.go:264-280 */
void og_nulls_filter(const uint64_t *src, uint64_t src_len_bits, const int64_t *sels, int64_t nsel, uint64_t *dst) {
    memset(dst, 0, bm_words((uint64_t)nsel) * 8);
    if (!src) return;
    for (int64_t i = 0; i < nsel; i++) {
        if ((uint64_t)sels[i] >= src_len_bits) continue;
        if (bm_has(src, (uint64_t)sels[i])) bm_add(dst, (uint64_t)i);
    }
}

/*
 * og_group_ids: observable behaviour of intHashMapIterator.Insert + Int64HashMap
 * (pkg/common/hashmap/iterator.go:127-148, inthashmap.go:92-183, container/hashtable/int64_hash_map.go:
 * 113-158) as driven by Group.buildOneBatch (pkg/sql/colexec/group/exec2.go:325-362): every row gets a
 * 1-based group id, ids are handed out in FIRST-SEEN order.  The hash function itself is seeded randomly per
 * process (hashtable/hash.go:41-47) so only ids are observable; any exact map reproduces them.
 * keys[] are the packed <=8-byte group keys (inthashmap.go fillKeys).  State persists across batches via
 * (table_keys, *ngroups).  Returns the new group count.
 */
int64_t og_group_ids(const uint64_t *keys, uint64_t n, uint64_t *groups, uint64_t *table_keys, int64_t ngroups,
                     int64_t cap) {
    for (uint64_t i = 0; i < n; i++) {
        int64_t g = -1;
        for (int64_t j = 0; j < ngroups; j++) if (table_keys[j] == keys[i]) { g = j; break; }
        if (g < 0) { if (ngroups >= cap) return -1; g = ngroups; table_keys[ngroups++] = keys[i]; }
        groups[i] = (uint64_t)(g + 1);
    }
    return ngroups;
}

/* ---------------------------------------------------------------------------------------------
 * Aggregates.  groups[i] = 1-based group id, 0 = GroupNotMatched (skip).  State arrays are caller-owned
 * and indexed by group-1; *_isnull[g] = 1 means "no value yet" (SUM/MIN/MAX of an all-null group is NULL).
 * ------------------------------------------------------------------------------------------- */

/* og_sum_int64: sumAvgExec[int64,A].batchFillSum, pkg/sql/colexec/aggexec/sumavg2.go:133-166 with
 * int64OfCheck :89-94.  arg type A in {int8..int64}; strict row order; error at first overflowing row. */
#define SUM_I64_LOOP(A)                                                                            \
    for (uint64_t i = 0; i < n; i++) {                                                             \
        uint64_t grp = groups ? groups[i] : 1; if (grp == 0) continue;                             \
        if (bm_has(nulls, i + offset)) continue;                                                   \
        int64_t v1 = sums[grp - 1], v2 = (int64_t)((const A *)col)[i + offset];                    \
        int64_t s = (int64_t)((uint64_t)v1 + (uint64_t)v2);                                        \
        if ((v1 > 0 && v2 > 0 && s <= 0) || (v1 < 0 && v2 < 0 && s >= 0)) { if (err_row) *err_row = (int64_t)i; return OG_RC_OUT_OF_RANGE; } \
        if (isnull) isnull[grp - 1] = 0;                                                           \
        sums[grp - 1] = s; if (cnts) cnts[grp - 1] += 1;                                           \
    }
int32_t og_sum_int64(int32_t type, const void *col, const uint64_t *nulls, uint64_t offset, const uint64_t *groups,
                     uint64_t n, int64_t *sums, uint8_t *isnull, int64_t *cnts, int64_t *err_row) {
    switch (type) {
    case T_int8: SUM_I64_LOOP(int8_t) break; case T_int16: SUM_I64_LOOP(int16_t) break;
    case T_int32: SUM_I64_LOOP(int32_t) break; case T_int64: SUM_I64_LOOP(int64_t) break;
    default: return OG_RC_INVALID;
    }
    return OG_RC_OK;
}

/* og_sum_uint64: same with uint64OfCheck, sumavg2.go:96-101 */
#define SUM_U64_LOOP(A)                                                                            \
    for (uint64_t i = 0; i < n; i++) {                                                             \
        uint64_t grp = groups ? groups[i] : 1; if (grp == 0) continue;                             \
        if (bm_has(nulls, i + offset)) continue;                                                   \
        uint64_t v1 = sums[grp - 1], v2 = (uint64_t)((const A *)col)[i + offset], s = v1 + v2;     \
        if (s < v1 || s < v2) { if (err_row) *err_row = (int64_t)i; return OG_RC_OUT_OF_RANGE; }   \
        if (isnull) isnull[grp - 1] = 0;                                                           \
        sums[grp - 1] = s; if (cnts) cnts[grp - 1] += 1;                                           \
    }
int32_t og_sum_uint64(int32_t type, const void *col, const uint64_t *nulls, uint64_t offset, const uint64_t *groups,
                      uint64_t n, uint64_t *sums, uint8_t *isnull, int64_t *cnts, int64_t *err_row) {
    switch (type) {
    case T_uint8: SUM_U64_LOOP(uint8_t) break; case T_uint16: SUM_U64_LOOP(uint16_t) break;
    case T_uint32: SUM_U64_LOOP(uint32_t) break; case T_uint64: SUM_U64_LOOP(uint64_t) break;
    default: return OG_RC_INVALID;
    }
    return OG_RC_OK;
}

/* og_sum_float64: sumAvgExec[float64,A], A in {float32,float64}: sums[g] += float64(val), no check
 * (sumavg2.go:103-107,153-163); batchFillAvg additionally bumps cnts (:168-199). */
int32_t og_sum_float64(int32_t type, const void *col, const uint64_t *nulls, uint64_t offset, const uint64_t *groups,
                       uint64_t n, double *sums, uint8_t *isnull, int64_t *cnts) {
    for (uint64_t i = 0; i < n; i++) {
        uint64_t grp = groups ? groups[i] : 1; if (grp == 0) continue;
        if (bm_has(nulls, i + offset)) continue;
        double v = type == T_float32 ? (double)((const float *)col)[i + offset] : ((const double *)col)[i + offset];
        if (isnull) isnull[grp - 1] = 0;
        sums[grp - 1] = sums[grp - 1] + v;
        if (cnts) cnts[grp - 1] += 1;
    }
    return OG_RC_OK;
}

/* og_avg_flush: sumAvgExec.Flush, sumavg2.go:323-335: avg = float64(sum)/float64(cnt), cnt==0 -> NULL */
void og_avg_flush_f64(const double *sums, const int64_t *cnts, int64_t ngroups, double *avgs, uint8_t *isnull) {
    for (int64_t j = 0; j < ngroups; j++) {
        if (cnts[j] == 0) { isnull[j] = 1; avgs[j] = 0; } else { isnull[j] = 0; avgs[j] = sums[j] / (double)cnts[j]; }
    }
}
void og_avg_flush_i64(const int64_t *sums, const int64_t *cnts, int64_t ngroups, double *avgs, uint8_t *isnull) {
    for (int64_t j = 0; j < ngroups; j++) {
        if (cnts[j] == 0) { isnull[j] = 1; avgs[j] = 0; } else { isnull[j] = 0; avgs[j] = (double)sums[j] / (double)cnts[j]; }
    }
}

/* og_count: countStarExec.BatchFill (count2.go:45-61; nulls==NULL, star=1) and countColumnExec.BatchFill
 * (count2.go:120-146; skips null rows) */
void og_count(int32_t star, const uint64_t *nulls, uint64_t offset, const uint64_t *groups, uint64_t n, int64_t *vals) {
    for (uint64_t i = 0; i < n; i++) {
        uint64_t grp = groups ? groups[i] : 1; if (grp == 0) continue;
        if (!star && bm_has(nulls, i + offset)) continue;
        vals[grp - 1] += 1;
    }
}

/* og_minmax: minMaxExecFixed.BatchFill, pkg/sql/colexec/aggexec/minmax2.go:49-80: first non-null initialises;
 * then strict comp(value, agg) < 0 replaces (first-seen wins ties).  is_max flips the comparator. */
#define MINMAX_LOOP(T)                                                                             \
    for (uint64_t i = 0; i < n; i++) {                                                             \
        uint64_t grp = groups ? groups[i] : 1; if (grp == 0) continue;                             \
        if (bm_has(nulls, i + offset)) continue;                                                   \
        T v = ((const T *)col)[i + offset]; T *ag = (T *)aggs;                                     \
        if (isnull[grp - 1]) { isnull[grp - 1] = 0; ag[grp - 1] = v; }                             \
        else if (is_max ? (v > ag[grp - 1]) : (v < ag[grp - 1])) ag[grp - 1] = v;                  \
    }
int32_t og_minmax(int32_t is_max, int32_t type, const void *col, const uint64_t *nulls, uint64_t offset,
                  const uint64_t *groups, uint64_t n, void *aggs, uint8_t *isnull) {
    switch (type) {
    case T_bool: case T_uint8: MINMAX_LOOP(uint8_t) break;
    case T_int8: MINMAX_LOOP(int8_t) break; case T_int16: MINMAX_LOOP(int16_t) break;
    case T_int32: case T_date: MINMAX_LOOP(int32_t) break;
    case T_int64: case T_time: case T_datetime: case T_timestamp: MINMAX_LOOP(int64_t) break;
    case T_uint16: MINMAX_LOOP(uint16_t) break; case T_uint32: MINMAX_LOOP(uint32_t) break;
    case T_uint64: MINMAX_LOOP(uint64_t) break;
    case T_float32: MINMAX_LOOP(float) break; case T_float64: MINMAX_LOOP(double) break;
    default: return OG_RC_INVALID;
    }
    return OG_RC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Distances: pkg/vectorindex/metric/distance_func.go.  Accumulator type == element type T; the 8-way
 * (cosine: 4-way) unrolled association is kept exactly.
 * ------------------------------------------------------------------------------------------- */
#define DEF_DIST(SFX, T)                                                                           \
    /* L2DistanceSq, distance_func.go:59-95 */                                                     \
    T og_l2sq_##SFX(const T *p, const T *q, int64_t n) {                                           \
        T sum = 0; int64_t i = 0;                                                                  \
        for (; i <= n - 8; i += 8) {                                                               \
            T d0 = p[i] - q[i], d1 = p[i + 1] - q[i + 1], d2 = p[i + 2] - q[i + 2], d3 = p[i + 3] - q[i + 3]; \
            T d4 = p[i + 4] - q[i + 4], d5 = p[i + 5] - q[i + 5], d6 = p[i + 6] - q[i + 6], d7 = p[i + 7] - q[i + 7]; \
            T t0 = d0 * d0, t1 = d1 * d1, t2 = d2 * d2, t3 = d3 * d3, t4 = d4 * d4, t5 = d5 * d5, t6 = d6 * d6, t7 = d7 * d7; \
            T s01 = t0 + t1, s23 = t2 + t3, s45 = t4 + t5, s67 = t6 + t7;                          \
            T s = ((s01 + s23) + s45) + s67;                                                       \
            sum = sum + s;                                                                         \
        }                                                                                          \
        for (; i < n; i++) { T d = p[i] - q[i]; T t = d * d; sum = sum + t; }                      \
        return sum;                                                                                \
    }                                                                                              \
    /* L2Distance, distance_func.go:35-42: T(sqrt(float64(sumsq))) */                              \
    T og_l2_##SFX(const T *p, const T *q, int64_t n) { return (T)sqrt((double)og_l2sq_##SFX(p, q, n)); } \
    /* L1Distance, distance_func.go:112-154 */                                                     \
    T og_l1_##SFX(const T *p, const T *q, int64_t n) {                                             \
        T sum = 0;                                                                                 \
        for (int64_t i = 0; i < n; i++) { T d = p[i] - q[i]; if (d < 0) d = -d; sum = sum + d; }   \
        return sum;                                                                                \
    }                                                                                              \
    /* InnerProduct, distance_func.go:172-205: returns -sum; the 8 products are summed left to right */ \
    T og_ip_##SFX(const T *p, const T *q, int64_t n) {                                             \
        T sum = 0; int64_t i = 0;                                                                  \
        for (; i <= n - 8; i += 8) {                                                               \
            T t0 = p[i] * q[i], t1 = p[i + 1] * q[i + 1], t2 = p[i + 2] * q[i + 2], t3 = p[i + 3] * q[i + 3]; \
            T t4 = p[i + 4] * q[i + 4], t5 = p[i + 5] * q[i + 5], t6 = p[i + 6] * q[i + 6], t7 = p[i + 7] * q[i + 7]; \
            T s = t0 + t1; s = s + t2; s = s + t3; s = s + t4; s = s + t5; s = s + t6; s = s + t7; \
            sum = sum + s;                                                                         \
        }                                                                                          \
        for (; i < n; i++) { T t = p[i] * q[i]; sum = sum + t; }                                   \
        return -sum;                                                                               \
    }                                                                                              \
    /* shared accumulation of CosineDistance / CosineSimilarity, distance_func.go:216-262, 295-340 */ \
    static void cos_parts_##SFX(const T *p, const T *q, int64_t n, T *dot, T *n1, T *n2) {         \
        T dp = 0, a = 0, b = 0; int64_t i = 0;                                                     \
        for (; i <= n - 4; i += 4) {                                                               \
            T s;                                                                                   \
            s = p[i] * q[i] + p[i + 1] * q[i + 1]; s = s + p[i + 2] * q[i + 2]; s = s + p[i + 3] * q[i + 3]; dp = dp + s; \
            s = p[i] * p[i] + p[i + 1] * p[i + 1]; s = s + p[i + 2] * p[i + 2]; s = s + p[i + 3] * p[i + 3]; a = a + s;  \
            s = q[i] * q[i] + q[i + 1] * q[i + 1]; s = s + q[i + 2] * q[i + 2]; s = s + q[i + 3] * q[i + 3]; b = b + s;  \
        }                                                                                          \
        for (; i < n; i++) { dp = dp + p[i] * q[i]; a = a + p[i] * p[i]; b = b + q[i] * q[i]; }    \
        *dot = dp; *n1 = a; *n2 = b;                                                               \
    }                                                                                              \
    /* CosineDistance, distance_func.go:216-284: zero norm -> 1; clamp; T(1 - sim) */              \
    T og_cosdist_##SFX(const T *p, const T *q, int64_t n) {                                        \
        if (n == 0) return 0;                                                                      \
        T dp, a, b; cos_parts_##SFX(p, q, n, &dp, &a, &b);                                         \
        double den = sqrt((double)a) * sqrt((double)b);                                            \
        if (den == 0) return (T)1.0;                                                               \
        double sim = (double)dp / den;                                                             \
        if (sim > 1.0) sim = 1.0; else if (sim < -1.0) sim = -1.0;                                 \
        return (T)(1.0 - sim);                                                                     \
    }                                                                                              \
    /* CosineSimilarity, distance_func.go:295-356: zero norm -> error (*err = 1) */                \
    T og_cossim_##SFX(const T *p, const T *q, int64_t n, int32_t *err) {                           \
        *err = 0; if (n == 0) return 0;                                                            \
        T dp, a, b; cos_parts_##SFX(p, q, n, &dp, &a, &b);                                         \
        double den = sqrt((double)a) * sqrt((double)b);                                            \
        if (den == 0) { *err = 1; return 0; }                                                      \
        double sim = (double)dp / den;                                                             \
        if (sim > 1.0) sim = 1.0; else if (sim < -1.0) sim = -1.0;                                 \
        return (T)sim;                                                                             \
    }                                                                                              \
    /* NormalizeL2, distance_func.go:411-436 == moarray/external.go:262-285 */                     \
    int32_t og_normalize_l2_##SFX(const T *v, T *out, int64_t n) {                                 \
        if (n == 0) return 1;                                                                      \
        double ss = 0; for (int64_t i = 0; i < n; i++) ss = ss + (double)v[i] * (double)v[i];      \
        double norm = sqrt(ss);                                                                    \
        if (norm == 0) { memcpy(out, v, (size_t)n * sizeof(T)); return 0; }                        \
        for (int64_t i = 0; i < n; i++) out[i] = (T)((double)v[i] / norm);                         \
        return 0;                                                                                  \
    }
DEF_DIST(f32, float)
DEF_DIST(f64, double)

/* moarray.CosineSimilarity, pkg/vectorize/moarray/external.go:212-260: float32 snap to +-1 */
double og_moarray_cossim_f32(const float *p, const float *q, int64_t n, int32_t *err) {
    double c = (double)og_cossim_f32(p, q, n, err); float f = (float)c;
    if (f == 1.0f) c = 1; else if (f == -1.0f) c = -1; return c;
}
double og_moarray_cossim_f64(const double *p, const double *q, int64_t n, int32_t *err) {
    double c = og_cossim_f64(p, q, n, err); float f = (float)c;
    if (f == 1.0f) c = 1; else if (f == -1.0f) c = -1; return c;
}

enum { METRIC_L2 = 0, METRIC_IP = 1, METRIC_COS = 2, METRIC_L1 = 3, METRIC_L2SQ = 4 }; /* metric/types.go MetricType */

/* ResolveDistanceFn, distance_func.go:507-524: L2 and L2sq both resolve to L2DistanceSq */
static inline float distfn_f32(int metric, const float *p, const float *q, int64_t n) {
    switch (metric) {
    case METRIC_IP: return og_ip_f32(p, q, n);
    case METRIC_COS: return og_cosdist_f32(p, q, n);
    case METRIC_L1: return og_l1_f32(p, q, n);
    default: return og_l2sq_f32(p, q, n);
    }
}
static inline double distfn_f64(int metric, const double *p, const double *q, int64_t n) {
    switch (metric) {
    case METRIC_IP: return og_ip_f64(p, q, n);
    case METRIC_COS: return og_cosdist_f64(p, q, n);
    case METRIC_L1: return og_l1_f64(p, q, n);
    default: return og_l2sq_f64(p, q, n);
    }
}

/* og_distance_rows: the SQL builtins L2DistanceArray / L2DistanceSqArray / InnerProductArray /
 * CosineDistanceArray (pkg/sql/plan/function/func_binary.go:7763-7786,10540-10554 via
 * moarray/external.go:171-210): per row float64(metric(a_i, b_i)); b may be const (stride 0).
 * kind: 0 l2, 1 ip, 2 cosine distance, 3 l1, 4 l2sq (the METRIC_* ids).  Rows null in rnulls are skipped. */
int32_t og_distance_rows_f32(int32_t kind, double *r, const float *a, int64_t astride, const float *b, int64_t bstride,
                             int64_t dim, uint64_t n, const uint64_t *rnulls) {
    for (uint64_t i = 0; i < n; i++) {
        if (bm_has(rnulls, i)) continue;
        const float *p = a + (int64_t)i * astride, *q = b + (int64_t)i * bstride;
        r[i] = kind == 0 ? (double)og_l2_f32(p, q, dim) : (double)distfn_f32(kind, p, q, dim);
    }
    return OG_RC_OK;
}
int32_t og_distance_rows_f64(int32_t kind, double *r, const double *a, int64_t astride, const double *b, int64_t bstride,
                             int64_t dim, uint64_t n, const uint64_t *rnulls) {
    for (uint64_t i = 0; i < n; i++) {
        if (bm_has(rnulls, i)) continue;
        const double *p = a + (int64_t)i * astride, *q = b + (int64_t)i * bstride;
        r[i] = kind == 0 ? og_l2_f64(p, q, dim) : distfn_f64(kind, p, q, dim);
    }
    return OG_RC_OK;
}

/* ---------------------------------------------------------------------------------------------
 * FastMaxHeap: pkg/vectorindex/index.go:171-250 (bounded max-heap, SoA, strict '<' replace rule)
 * ------------------------------------------------------------------------------------------- */
#define DEF_HEAP(SFX, T)                                                                           \
    typedef struct { int64_t *keys; T *dist; int size, limit; } heap_##SFX;                        \
    static void heap_up_##SFX(heap_##SFX *h, int j) { /* siftUp :189-199 */                        \
        for (;;) {                                                                                 \
            int i = (j - 1) / 2;                                                                   \
            if (i == j || h->dist[j] <= h->dist[i]) break;                                         \
            T td = h->dist[i]; h->dist[i] = h->dist[j]; h->dist[j] = td;                           \
            int64_t tk = h->keys[i]; h->keys[i] = h->keys[j]; h->keys[j] = tk;                     \
            j = i;                                                                                 \
        }                                                                                          \
    }                                                                                              \
    static void heap_down_##SFX(heap_##SFX *h, int i0, int n) { /* siftDown :201-219 */            \
        int i = i0;                                                                                \
        for (;;) {                                                                                 \
            int j1 = 2 * i + 1;                                                                    \
            if (j1 >= n || j1 < 0) break;                                                          \
            int j = j1, j2 = j1 + 1;                                                               \
            if (j2 < n && h->dist[j2] > h->dist[j1]) j = j2;                                       \
            if (h->dist[j] <= h->dist[i]) break;                                                   \
            T td = h->dist[i]; h->dist[i] = h->dist[j]; h->dist[j] = td;                           \
            int64_t tk = h->keys[i]; h->keys[i] = h->keys[j]; h->keys[j] = tk;                     \
            i = j;                                                                                 \
        }                                                                                          \
    }                                                                                              \
    static void heap_push_##SFX(heap_##SFX *h, int64_t key, T d) { /* Push :223-234 */             \
        if (h->size < h->limit) { h->dist[h->size] = d; h->keys[h->size] = key; heap_up_##SFX(h, h->size); h->size++; } \
        else if (d < h->dist[0]) { h->dist[0] = d; h->keys[0] = key; heap_down_##SFX(h, 0, h->limit); } \
    }                                                                                              \
    static bool heap_pop_##SFX(heap_##SFX *h, int64_t *key, T *d) { /* Pop :237-250 */             \
        if (h->size == 0) { *key = -1; *d = 0; return false; }                                     \
        h->size--; *key = h->keys[0]; *d = h->dist[0];                                             \
        h->keys[0] = h->keys[h->size]; h->dist[0] = h->dist[h->size];                              \
        heap_down_##SFX(h, 0, h->size);                                                            \
        return true;                                                                               \
    }
DEF_HEAP(f32, float)
DEF_HEAP(f64, double)

/* og_heap_topk_f32: drive the heap directly (pins index_test.go:215-281): push all, pop into ascending order */
void og_heap_topk_f32(const float *d, const int64_t *keys, int64_t n, int limit, int64_t *okeys, float *odist) {
    int64_t *hk = malloc(sizeof(int64_t) * (size_t)limit); float *hd = malloc(sizeof(float) * (size_t)limit);
    heap_f32 h = {hk, hd, 0, limit};
    for (int64_t j = 0; j < n; j++) heap_push_f32(&h, keys ? keys[j] : j, d[j]);
    for (int j = limit - 1; j >= 0; j--) { int64_t k; float dd; heap_pop_f32(&h, &k, &dd); okeys[j] = k; odist[j] = dd; }
    free(hk); free(hd);
}

/*
 * og_bruteforce_search: GoBruteForceIndex.Search, pkg/vectorindex/brute_force/brute_force.go:248-341.
 * dataset row-major n x dim, queries nq x dim; keys = row ordinals; distances float64(dist).
 * limit==1: running min with strict '<' starting from MaxFloat (:287-303); else FastMaxHeap, popped into
 * ascending order, missing entries padded key=-1 dist=0 AT THE FRONT (:319-331).
 * Thread pool over queries (:271-275).
 */
typedef struct {
    int is_f64, metric, limit; const void *data, *queries; int64_t n, dim, q0, q1; int64_t *keys; double *dists;
    const int64_t *row_ids; /* optional: candidate row subset (IVF list scan), NULL = all rows */ int64_t nrow_ids;
} bf_job;

static void *bf_worker(void *arg) {
    bf_job *j = (bf_job *)arg;
    int limit = j->limit;
    int64_t *hk = malloc(sizeof(int64_t) * (size_t)(limit > 0 ? limit : 1));
    float *hd32 = malloc(sizeof(float) * (size_t)(limit > 0 ? limit : 1));
    double *hd64 = malloc(sizeof(double) * (size_t)(limit > 0 ? limit : 1));
    int64_t nrows = j->row_ids ? j->nrow_ids : j->n;
    for (int64_t k = j->q0; k < j->q1; k++) {
        if (!j->is_f64) {
            const float *q = (const float *)j->queries + k * j->dim, *D = (const float *)j->data;
            if (limit == 1) {
                float mind = FLT_MAX; int64_t mini = -1;
                for (int64_t x = 0; x < nrows; x++) {
                    int64_t r = j->row_ids ? j->row_ids[x] : x;
                    float d = distfn_f32(j->metric, q, D + r * j->dim, j->dim);
                    if (d < mind) { mind = d; mini = r; }
                }
                j->keys[k] = mini; j->dists[k] = (double)mind; continue;
            }
            heap_f32 h = {hk, hd32, 0, limit};
            for (int64_t x = 0; x < nrows; x++) {
                int64_t r = j->row_ids ? j->row_ids[x] : x;
                heap_push_f32(&h, r, distfn_f32(j->metric, q, D + r * j->dim, j->dim));
            }
            for (int t = limit - 1; t >= 0; t--) {
                int64_t key; float d;
                if (!heap_pop_f32(&h, &key, &d)) { j->keys[k * limit + t] = -1; j->dists[k * limit + t] = 0; continue; }
                j->keys[k * limit + t] = key; j->dists[k * limit + t] = (double)d;
            }
        } else {
            const double *q = (const double *)j->queries + k * j->dim, *D = (const double *)j->data;
            if (limit == 1) {
                double mind = DBL_MAX; int64_t mini = -1;
                for (int64_t x = 0; x < nrows; x++) {
                    int64_t r = j->row_ids ? j->row_ids[x] : x;
                    double d = distfn_f64(j->metric, q, D + r * j->dim, j->dim);
                    if (d < mind) { mind = d; mini = r; }
                }
                j->keys[k] = mini; j->dists[k] = mind; continue;
            }
            heap_f64 h = {hk, hd64, 0, limit};
            for (int64_t x = 0; x < nrows; x++) {
                int64_t r = j->row_ids ? j->row_ids[x] : x;
                heap_push_f64(&h, r, distfn_f64(j->metric, q, D + r * j->dim, j->dim));
            }
            for (int t = limit - 1; t >= 0; t--) {
                int64_t key; double d;
                if (!heap_pop_f64(&h, &key, &d)) { j->keys[k * limit + t] = -1; j->dists[k * limit + t] = 0; continue; }
                j->keys[k * limit + t] = key; j->dists[k * limit + t] = d;
            }
        }
    }
    free(hk); free(hd32); free(hd64);
    return NULL;
}

int32_t og_bruteforce_search(int32_t is_f64, int32_t metric, const void *dataset, int64_t n, int64_t dim,
                             const void *queries, int64_t nq, int32_t limit, int32_t nthreads,
                             int64_t *keys, double *dists) {
    if (limit == 0) return OG_RC_OK;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nq) nthreads = (int32_t)(nq > 0 ? nq : 1);
    bf_job *jobs = malloc(sizeof(bf_job) * (size_t)nthreads);
    int64_t per = (nq + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int64_t q0 = t * per, q1 = q0 + per; if (q1 > nq) q1 = nq; if (q0 > nq) q0 = nq;
        jobs[t] = (bf_job){is_f64, metric, limit, dataset, queries, n, dim, q0, q1, keys, dists, NULL, 0};
    }
    og_run(bf_worker, jobs, sizeof(bf_job), nthreads);
    free(jobs);
    return OG_RC_OK;
}

/*
 * og_ivf_search: IvfflatSearchIndex.Search, pkg/vectorindex/ivfflat/search.go:509-630.
 *   1. findCentroids (:292-311): brute-force top-nprobe over the centroid table with the resolved metric,
 *      one query at a time (NThreads=1) -> probed list ids (ascending distance).
 *   2. the SQL list scan "SELECT pk, dist(entry,q) ... WHERE id IN (probed) ORDER BY vec_dist LIMIT k"
 *      (:572-592) = distance of q to every entry whose list id is probed, then Top (colexec/top) k.
 *      We model the scan + ORDER BY/LIMIT as the same bounded-heap top-k over rows in storage order
 *      (list by list in probe-rank order is NOT guaranteed by SQL; we scan rows in ascending row id, i.e.
 *      table order, which is what a table scan yields).
 *   3. DistanceTransformIvfflat (metric/types.go:138-144): sqrt on the final k when the user metric is L2.
 * assign[r] = list id of row r.  f32 only (vecf32 is the BASELINE config).  One thread per query range.
 */
typedef struct { const float *data; const int32_t *assign; int64_t n, dim; const float *cent; int64_t nlist;
                 const float *queries; int64_t q0, q1; int nprobe, limit, metric, sqrt_out; int64_t *keys; double *dists; } ivf_job;

static void *ivf_worker(void *arg) {
    ivf_job *j = (ivf_job *)arg;
    int64_t *pk = malloc(sizeof(int64_t) * (size_t)j->nprobe); float *pd = malloc(sizeof(float) * (size_t)j->nprobe);
    int64_t *hk = malloc(sizeof(int64_t) * (size_t)j->limit); float *hd = malloc(sizeof(float) * (size_t)j->limit);
    uint8_t *probed = malloc((size_t)j->nlist);
    for (int64_t k = j->q0; k < j->q1; k++) {
        const float *q = j->queries + k * j->dim;
        heap_f32 hc = {pk, pd, 0, j->nprobe};
        for (int64_t c = 0; c < j->nlist; c++) heap_push_f32(&hc, c, distfn_f32(j->metric, q, j->cent + c * j->dim, j->dim));
        memset(probed, 0, (size_t)j->nlist);
        for (int t = 0; t < hc.size; t++) probed[pk[t]] = 1;
        heap_f32 h = {hk, hd, 0, j->limit};
        for (int64_t r = 0; r < j->n; r++) {
            if (!probed[j->assign[r]]) continue;
            heap_push_f32(&h, r, distfn_f32(j->metric, q, j->data + r * j->dim, j->dim));
        }
        for (int t = j->limit - 1; t >= 0; t--) {
            int64_t key; float d;
            if (!heap_pop_f32(&h, &key, &d)) { j->keys[k * j->limit + t] = -1; j->dists[k * j->limit + t] = 0; continue; }
            j->keys[k * j->limit + t] = key;
            j->dists[k * j->limit + t] = j->sqrt_out ? sqrt((double)d) : (double)d;
        }
    }
    free(pk); free(pd); free(hk); free(hd); free(probed);
    return NULL;
}

int32_t og_ivf_search_f32(const float *data, const int32_t *assign, int64_t n, int64_t dim, const float *centroids,
                          int64_t nlist, const float *queries, int64_t nq, int32_t nprobe, int32_t limit,
                          int32_t metric, int32_t sqrt_out, int32_t nthreads, int64_t *keys, double *dists) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nq) nthreads = (int32_t)(nq > 0 ? nq : 1);
    if (nprobe > nlist) nprobe = (int32_t)nlist;
    ivf_job *jobs = malloc(sizeof(ivf_job) * (size_t)nthreads);
    int64_t per = (nq + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int64_t q0 = t * per, q1 = q0 + per; if (q1 > nq) q1 = nq; if (q0 > nq) q0 = nq;
        jobs[t] = (ivf_job){data, assign, n, dim, centroids, nlist, queries, q0, q1, nprobe, limit, metric, sqrt_out, keys, dists};
    }
    og_run(ivf_worker, jobs, sizeof(ivf_job), nthreads);
    free(jobs);
    return OG_RC_OK;
}

/* og_assign_centroids: Productl2.probeRun, pkg/sql/colexec/productl2/product_l2.go:317-407 -- brute-force
 * Search(limit=1): argmin with strict '<' (first-seen centroid wins ties). */
void og_assign_centroids_f32(const float *data, int64_t n, int64_t dim, const float *centroids, int64_t nlist,
                             int32_t metric, int32_t *assign) {
    for (int64_t r = 0; r < n; r++) {
        float mind = FLT_MAX; int64_t mini = -1;
        for (int64_t c = 0; c < nlist; c++) {
            float d = distfn_f32(metric, data + r * dim, centroids + c * dim, dim);
            if (d < mind) { mind = d; mini = c; }
        }
        assign[r] = (int32_t)mini;
    }
}

/* ---------------------------------------------------------------------------------------------
 * Pipelines (what the bench's CPU baseline times).  Each follows the operator chain the reference runs
 * per 8192-row block: table_scan -> filter (conjunct by conjunct, Shrink after each) -> projection ->
 * group (aggexec BatchFill/BulkFill), one pipeline per worker over a contiguous block range
 * (pkg/sql/compile/scope.go:521-568), partial states merged at the end like MergeGroup
 * (pkg/sql/colexec/group/mergeGroup.go:132-247; sums merged in worker order).
 * ------------------------------------------------------------------------------------------- */
#define BLOCK_ROWS 8192 /* objectio.BlockMaxRows, pkg/objectio/const.go:26 */

typedef struct {
    const int32_t *shipdate; const double *discount, *quantity, *extprice; int64_t row0, row1;
    int32_t date_lo, date_hi; double disc_lo, disc_hi, qty_hi;
    double sum; int64_t nsel; uint8_t isnull;
} q6_job;

/* One pipeline: filter conjuncts in plan order -- shipdate >= lo ; shipdate < hi ; discount BETWEEN ; quantity <
 * (q6.sql:58-61) -- each followed by sels + Shrink of ALL FOUR columns (filter.go:116-152), then
 * projection l_extendedprice*l_discount (multiFn float64, unchecked) and SUM BulkFill (sumavg2.go:119-121). */
static void *q6_worker(void *arg) {
    q6_job *j = (q6_job *)arg;
    int32_t *sd = malloc(BLOCK_ROWS * 4), *sd2 = malloc(BLOCK_ROWS * 4);
    double *di = malloc(BLOCK_ROWS * 8), *qu = malloc(BLOCK_ROWS * 8), *ep = malloc(BLOCK_ROWS * 8);
    double *di2 = malloc(BLOCK_ROWS * 8), *qu2 = malloc(BLOCK_ROWS * 8), *ep2 = malloc(BLOCK_ROWS * 8), *proj = malloc(BLOCK_ROWS * 8);
    bool *bv = malloc(BLOCK_ROWS); int64_t *sels = malloc(BLOCK_ROWS * 8);
    double sum = 0; uint8_t isnull = 1; int64_t nsel_total = 0;
    for (int64_t b0 = j->row0; b0 < j->row1; b0 += BLOCK_ROWS) {
        int64_t n = j->row1 - b0 < BLOCK_ROWS ? j->row1 - b0 : BLOCK_ROWS;
        const int32_t *csd = j->shipdate + b0; const double *cdi = j->discount + b0, *cqu = j->quantity + b0, *cep = j->extprice + b0;
        int32_t *osd = sd, *osd_alt = sd2; double *odi = di, *oqu = qu, *oep = ep, *odi_alt = di2, *oqu_alt = qu2, *oep_alt = ep2;
        for (int conj = 0; conj < 4 && n > 0; conj++) {
            switch (conj) {
            case 0: for (int64_t i = 0; i < n; i++) bv[i] = csd[i] >= j->date_lo; break;
            case 1: for (int64_t i = 0; i < n; i++) bv[i] = csd[i] < j->date_hi; break;
            case 2: for (int64_t i = 0; i < n; i++) bv[i] = cdi[i] >= j->disc_lo && cdi[i] <= j->disc_hi; break;
            default: for (int64_t i = 0; i < n; i++) bv[i] = cqu[i] < j->qty_hi; break;
            }
            int64_t k = 0;
            for (int64_t i = 0; i < n; i++) if (bv[i]) sels[k++] = i;
            if (k != n) { /* Shrink / Union: gather every column */
                for (int64_t i = 0; i < k; i++) { osd[i] = csd[sels[i]]; odi[i] = cdi[sels[i]]; oqu[i] = cqu[sels[i]]; oep[i] = cep[sels[i]]; }
                csd = osd; cdi = odi; cqu = oqu; cep = oep;
                int32_t *ts = osd; osd = osd_alt; osd_alt = ts;
                double *t; t = odi; odi = odi_alt; odi_alt = t; t = oqu; oqu = oqu_alt; oqu_alt = t; t = oep; oep = oep_alt; oep_alt = t;
                n = k;
            }
        }
        if (n == 0) continue;
        for (int64_t i = 0; i < n; i++) proj[i] = cep[i] * cdi[i];
        for (int64_t i = 0; i < n; i++) { sum = sum + proj[i]; }
        isnull = 0; nsel_total += n;
    }
    j->sum = sum; j->isnull = isnull; j->nsel = nsel_total;
    free(sd); free(sd2); free(di); free(qu); free(ep); free(di2); free(qu2); free(ep2); free(proj); free(bv); free(sels);
    return NULL;
}

/* Returns 0; *sum = SUM (0 and *isnull=1 if no row qualifies), *nsel = qualifying rows. */
int32_t og_q6(const int32_t *shipdate, const double *discount, const double *quantity, const double *extprice,
              int64_t n, int32_t date_lo, int32_t date_hi, double disc_lo, double disc_hi, double qty_hi,
              int32_t nthreads, double *sum, uint8_t *isnull, int64_t *nsel) {
    if (nthreads < 1) nthreads = 1;
    int64_t nblocks = (n + BLOCK_ROWS - 1) / BLOCK_ROWS;
    if (nthreads > nblocks) nthreads = (int32_t)(nblocks > 0 ? nblocks : 1);
    q6_job *jobs = malloc(sizeof(q6_job) * (size_t)nthreads);
    int64_t per = (nblocks + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int64_t r0 = t * per * BLOCK_ROWS, r1 = r0 + per * BLOCK_ROWS; if (r0 > n) r0 = n; if (r1 > n) r1 = n;
        jobs[t] = (q6_job){shipdate, discount, quantity, extprice, r0, r1, date_lo, date_hi, disc_lo, disc_hi, qty_hi, 0, 0, 1};
    }
    og_run(q6_worker, jobs, sizeof(q6_job), nthreads);
    double s = 0; uint8_t nul = 1; int64_t ns = 0;
    for (int t = 0; t < nthreads; t++) {
        if (!jobs[t].isnull) { if (nul) { nul = 0; s = jobs[t].sum; } else s = s + jobs[t].sum; } /* BatchMerge sumavg2.go:222-236 */
        ns += jobs[t].nsel;
    }
    *sum = s; *isnull = nul; *nsel = ns;
    free(jobs);
    return OG_RC_OK;
}

/*
 * Q1 (q1.sql:1-21): filter l_shipdate <= cutoff; group by (l_returnflag, l_linestatus) packed into a <=8-byte
 * int key (group/exec2.go:73-118 -> IntHashMap); aggregates in select-list order:
 *   sum(qty) sum(price) sum(price*(1-disc)) sum(price*(1-disc)*(1+tax)) avg(qty) avg(price) avg(disc) count(*)
 * Projections are evaluated as separate full vectors (evaluateGroupByAndAggArgs, group/types2.go:267-299):
 *   t1 = 1 - disc ; t2 = price * t1 ; t3 = 1 + tax ; t4 = t2 * t3.
 * Output groups in first-seen order; state layout per group g (OG_Q1_MAXG groups max):
 *   out_keys[g] (rf | ls<<8), sums[g*7 + {0 qty,1 price,2 disc_price,3 charge,4 avg_qty_sum,5 avg_price_sum,6 avg_disc_sum}],
 *   cnts[g*4 + {0 avg_qty_cnt,1 avg_price_cnt,2 avg_disc_cnt,3 count_star}]
 */
#define OG_Q1_MAXG 64
typedef struct {
    const int32_t *shipdate; const double *qty, *price, *disc, *tax; const uint8_t *rf, *ls; int64_t row0, row1; int32_t cutoff;
    uint64_t keys[OG_Q1_MAXG]; int64_t first_row[OG_Q1_MAXG]; int64_t ng; double sums[OG_Q1_MAXG * 7]; int64_t cnts[OG_Q1_MAXG * 4]; int overflow;
} q1_job;

static void *q1_worker(void *arg) {
    q1_job *j = (q1_job *)arg;
    double *cq = malloc(BLOCK_ROWS * 8), *cp = malloc(BLOCK_ROWS * 8), *cd = malloc(BLOCK_ROWS * 8), *ct = malloc(BLOCK_ROWS * 8);
    double *t1 = malloc(BLOCK_ROWS * 8), *t2 = malloc(BLOCK_ROWS * 8), *t3 = malloc(BLOCK_ROWS * 8), *t4 = malloc(BLOCK_ROWS * 8);
    uint64_t *kk = malloc(BLOCK_ROWS * 8), *groups = malloc(BLOCK_ROWS * 8); int64_t *sels = malloc(BLOCK_ROWS * 8);
    j->ng = 0; memset(j->sums, 0, sizeof j->sums); memset(j->cnts, 0, sizeof j->cnts); j->overflow = 0;
    for (int64_t b0 = j->row0; b0 < j->row1; b0 += BLOCK_ROWS) {
        int64_t n0 = j->row1 - b0 < BLOCK_ROWS ? j->row1 - b0 : BLOCK_ROWS, n = 0;
        for (int64_t i = 0; i < n0; i++) if (j->shipdate[b0 + i] <= j->cutoff) sels[n++] = i; /* lessEqualFn + sels */
        if (n == 0) continue;
        for (int64_t i = 0; i < n; i++) { /* Shrink/Union of the surviving rows */
            int64_t s = b0 + sels[i];
            cq[i] = j->qty[s]; cp[i] = j->price[s]; cd[i] = j->disc[s]; ct[i] = j->tax[s];
            kk[i] = (uint64_t)j->rf[s] | ((uint64_t)j->ls[s] << 8);
        }
        for (int64_t i = 0; i < n; i++) t1[i] = 1.0 - cd[i];
        for (int64_t i = 0; i < n; i++) t2[i] = cp[i] * t1[i];
        for (int64_t i = 0; i < n; i++) t3[i] = 1.0 + ct[i];
        for (int64_t i = 0; i < n; i++) t4[i] = t2[i] * t3[i];
        for (int64_t m0 = 0; m0 < n; m0 += 256) { /* hashmap.UnitLimit mini-batches, group/exec2.go:325-362 */
            int64_t m = n - m0 < 256 ? n - m0 : 256;
            int64_t ng0 = j->ng;
            int64_t ng = og_group_ids(kk + m0, (uint64_t)m, groups, j->keys, j->ng, OG_Q1_MAXG);
            if (ng < 0) { j->overflow = 1; goto done; }
            for (int64_t g = ng0; g < ng; g++) j->first_row[g] = -1;
            j->ng = ng;
            for (int64_t i = 0; i < m; i++) {
                uint64_t g = groups[i] - 1;
                if (j->first_row[g] < 0) j->first_row[g] = b0 + sels[m0 + i];
            }
            const double *srcs[7] = {cq, cp, t2, t4, cq, cp, cd};
            for (int a = 0; a < 7; a++) /* one BatchFill pass per aggregate, strict row order */
                for (int64_t i = 0; i < m; i++) {
                    uint64_t g = groups[i] - 1;
                    j->sums[g * 7 + a] = j->sums[g * 7 + a] + srcs[a][m0 + i];
                    if (a >= 4) j->cnts[g * 4 + (a - 4)] += 1;
                }
            for (int64_t i = 0; i < m; i++) j->cnts[(groups[i] - 1) * 4 + 3] += 1;
        }
    }
done:
    free(cq); free(cp); free(cd); free(ct); free(t1); free(t2); free(t3); free(t4); free(kk); free(groups); free(sels);
    return NULL;
}

/* out arrays sized OG_Q1_MAXG; returns group count (first-seen order by global row index), or -1 on overflow. */
int64_t og_q1(const int32_t *shipdate, const double *qty, const double *price, const double *disc, const double *tax,
              const uint8_t *rf, const uint8_t *ls, int64_t n, int32_t cutoff, int32_t nthreads,
              uint64_t *out_keys, double *out_sums, int64_t *out_cnts, int64_t *out_first_row) {
    if (nthreads < 1) nthreads = 1;
    int64_t nblocks = (n + BLOCK_ROWS - 1) / BLOCK_ROWS;
    if (nthreads > nblocks) nthreads = (int32_t)(nblocks > 0 ? nblocks : 1);
    q1_job *jobs = calloc((size_t)nthreads, sizeof(q1_job));
    int64_t per = (nblocks + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int64_t r0 = t * per * BLOCK_ROWS, r1 = r0 + per * BLOCK_ROWS; if (r0 > n) r0 = n; if (r1 > n) r1 = n;
        jobs[t].shipdate = shipdate; jobs[t].qty = qty; jobs[t].price = price; jobs[t].disc = disc; jobs[t].tax = tax;
        jobs[t].rf = rf; jobs[t].ls = ls; jobs[t].row0 = r0; jobs[t].row1 = r1; jobs[t].cutoff = cutoff;
    }
    og_run(q1_worker, jobs, sizeof(q1_job), nthreads);
    int64_t ng = 0; int bad = 0;
    for (int t = 0; t < nthreads; t++) {
        if (jobs[t].overflow) bad = 1;
        for (int64_t g = 0; g < jobs[t].ng && !bad; g++) { /* MergeGroup: re-hash partial keys, BatchMerge */
            int64_t d = -1;
            for (int64_t x = 0; x < ng; x++) if (out_keys[x] == jobs[t].keys[g]) { d = x; break; }
            if (d < 0) {
                if (ng >= OG_Q1_MAXG) { bad = 1; break; }
                d = ng++; out_keys[d] = jobs[t].keys[g]; out_first_row[d] = jobs[t].first_row[g];
                for (int a = 0; a < 7; a++) out_sums[d * 7 + a] = 0;
                for (int a = 0; a < 4; a++) out_cnts[d * 4 + a] = 0;
            }
            for (int a = 0; a < 7; a++) out_sums[d * 7 + a] = out_sums[d * 7 + a] + jobs[t].sums[g * 7 + a];
            for (int a = 0; a < 4; a++) out_cnts[d * 4 + a] += jobs[t].cnts[g * 4 + a];
        }
    }
    free(jobs);
    return bad ? -1 : ng;
}

/* og_sum_int64_mt: config 1 -- SUM(int64 col) no group-by (H0: BulkFill per block), block-range workers,
 * partials merged with the same overflow check (BatchMerge sumavg2.go:229-234). */
typedef struct { const int64_t *col; const uint64_t *nulls; int64_t row0, row1; int64_t sum; uint8_t isnull; int32_t rc; } s64_job;
static void *s64_worker(void *arg) {
    s64_job *j = (s64_job *)arg; j->sum = 0; j->isnull = 1; j->rc = 0;
    for (int64_t b0 = j->row0; b0 < j->row1 && !j->rc; b0 += BLOCK_ROWS) {
        int64_t n = j->row1 - b0 < BLOCK_ROWS ? j->row1 - b0 : BLOCK_ROWS;
        j->rc = og_sum_int64(T_int64, j->col, j->nulls, (uint64_t)b0, NULL, (uint64_t)n, &j->sum, &j->isnull, NULL, NULL);
    }
    return NULL;
}
int32_t og_sum_int64_mt(const int64_t *col, const uint64_t *nulls, int64_t n, int32_t nthreads, int64_t *sum, uint8_t *isnull) {
    if (nthreads < 1) nthreads = 1;
    int64_t nblocks = (n + BLOCK_ROWS - 1) / BLOCK_ROWS;
    if (nthreads > nblocks) nthreads = (int32_t)(nblocks > 0 ? nblocks : 1);
    s64_job *jobs = calloc((size_t)nthreads, sizeof(s64_job));
    int64_t per = (nblocks + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int64_t r0 = t * per * BLOCK_ROWS, r1 = r0 + per * BLOCK_ROWS; if (r0 > n) r0 = n; if (r1 > n) r1 = n;
        jobs[t].col = col; jobs[t].nulls = nulls; jobs[t].row0 = r0; jobs[t].row1 = r1;
    }
    og_run(s64_worker, jobs, sizeof(s64_job), nthreads);
    int64_t s = 0; uint8_t nul = 1; int32_t rc = 0;
    for (int t = 0; t < nthreads; t++) {
        if (jobs[t].rc) rc = jobs[t].rc;
        if (!jobs[t].isnull && !rc) {
            if (nul) { nul = 0; s = jobs[t].sum; }
            else {
                int64_t v1 = s, v2 = jobs[t].sum, r = (int64_t)((uint64_t)v1 + (uint64_t)v2);
                if ((v1 > 0 && v2 > 0 && r <= 0) || (v1 < 0 && v2 < 0 && r >= 0)) rc = OG_RC_OUT_OF_RANGE;
                s = r;
            }
        }
    }
    *sum = s; *isnull = nul;
    free(jobs);
    return rc;
}



/* ---------------------------------------------------------------------------------------------
 * Decimal64 / Decimal128 batch arithmetic and decimal SUM (the reference's native TPC-H types are
 * DECIMAL(15,2)).  Decimal64 = int64 unscaled value, Decimal128 = two's-complement 128-bit {B0_63, B64_127}; the scale lives in the type.
 *   og_d64_addsub   d64Add / d64Sub, pkg/sql/plan/function/arith_decimal_fast.go:3618-3900: the lower-scale operand is scaled up by
 *                   10^diff (d64ScaleIntoRs / d64MulPow10 :4807-4890, error "scale overflow" when |x|*10^diff >= 2^63), then a plain int64
 *                   add/sub whose sign-rule overflow fails the call at the FIRST offending row; result scale = max(scale1, scale2)
 *   og_d64_mul      d64Mul :3901-4095: exact 128-bit product, result scale = min(s1+s2, max(12, s1, s2)); when that is smaller than s1+s2
 *                   the magnitude is divided by 10^k with round-half-up (d128DivPow10Once :4913-4923)
 *   og_d128_addsub  d128Add / d128Sub :111-430 (same rules on 128 bits)
 *   og_d128_mul     d128Mul / d128MulInline :545-733: 256-bit product of the magnitudes, same scale rule, "Decimal128 Mul overflow" when the
 *                   result does not fit 127 bits
 *   og_sum_d64 / og_sum_d128   sumDecimal64FastExec / sumDecimal128FastExec.batchFill, pkg/sql/colexec/aggexec/sum_decimal_fast.go: 128-bit
 *                   wrapping accumulation per group + a row count (count == 0 -> NULL); AVG divides at Flush
 * rc: 0, or OG_RC_INVALID_INPUT with *err_row = first offending row.
 * ------------------------------------------------------------------------------------------- */
#define OG_RC_INVALID_INPUT 20203
typedef struct { uint64_t lo, hi; } og_d128;
static const uint64_t OG_POW10[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull,
    10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull, 100000000000000ull, 1000000000000000ull, 10000000000000000ull,
    100000000000000000ull, 1000000000000000000ull, 10000000000000000000ull};
static inline __int128 d128_get(og_d128 v) { return (__int128)(((unsigned __int128)v.hi << 64) | v.lo); }
static inline og_d128 d128_put(__int128 x) { og_d128 r = {(uint64_t)(unsigned __int128)x, (uint64_t)((unsigned __int128)x >> 64)}; return r; }

static int d64_scale_up(int64_t x, int diff, int64_t *out) {   /* d64MulPow10: ok iff |x| * 10^diff < 2^63 */
    uint64_t sign = (uint64_t)x >> 63, mask = 0 - sign, ab = ((uint64_t)x ^ mask) + sign;
    unsigned __int128 p = (unsigned __int128)ab * OG_POW10[diff];
    if ((uint64_t)(p >> 64) | ((uint64_t)p >> 63)) return 0;
    *out = (int64_t)((((uint64_t)p) ^ mask) + sign);
    return 1;
}
int32_t og_d64_addsub(int32_t is_sub, int64_t *r, const int64_t *a, const int64_t *b, uint64_t n, int32_t c1, int32_t c2, int32_t scale1, int32_t scale2,
                      const uint64_t *n1, const uint64_t *n2, uint64_t *rnulls, int64_t *err_row) {
    if ((c1 && bm_has(n1, 0)) || (c2 && bm_has(n2, 0))) { for (uint64_t i = 0; i < n; i++) bm_add(rnulls, i); return OG_RC_OK; }
    for (uint64_t w = 0; w < bm_words(n); w++) { if (!c1 && n1) rnulls[w] |= n1[w]; if (!c2 && n2) rnulls[w] |= n2[w]; }
    if (n & 63) rnulls[bm_words(n) - 1] &= (((uint64_t)1 << (n & 63)) - 1);
    const int d1 = scale2 > scale1 ? scale2 - scale1 : 0, d2 = scale1 > scale2 ? scale1 - scale2 : 0;
    if (d1 > 18 || d2 > 18) return OG_RC_INVALID;
    for (uint64_t i = 0; i < n; i++) {
        if (bm_has(rnulls, i)) continue;
        int64_t x = a[c1 ? 0 : i], y = b[c2 ? 0 : i];
        if ((d1 && !d64_scale_up(x, d1, &x)) || (d2 && !d64_scale_up(y, d2, &y))) { if (err_row) *err_row = (int64_t)i; return OG_RC_INVALID_INPUT; }
        int64_t z = (int64_t)(is_sub ? (uint64_t)x - (uint64_t)y : (uint64_t)x + (uint64_t)y);
        uint64_t sx = (uint64_t)x >> 63, sy = (uint64_t)y >> 63, sz = (uint64_t)z >> 63;
        int ov = is_sub ? (sx != sy && sx != sz) : (sx == sy && sx != sz);
        r[i] = z;
        if (ov) { if (err_row) *err_row = (int64_t)i; return OG_RC_INVALID_INPUT; }
    }
    return OG_RC_OK;
}

static int mul_scale_adj(int s1, int s2) { int d = 12; if (s1 > d) d = s1; if (s2 > d) d = s2; if (s1 + s2 < d) d = s1 + s2; return d - s1 - s2; }
int32_t og_mul_result_scale(int32_t s1, int32_t s2) { return s1 + s2 + mul_scale_adj(s1, s2); }

/* magnitude / 10^k with the reference's round-half-up, one or two steps (d128DivPow10 :518-526) on a 256-bit magnitude held as 4 limbs */
static void mag_div_pow10_once(uint64_t m[4], uint64_t d) {
    unsigned __int128 rem = 0;
    for (int k = 3; k >= 0; k--) { unsigned __int128 cur = (rem << 64) | m[k]; m[k] = (uint64_t)(cur / d); rem = cur % d; }
    if ((uint64_t)rem >= (d + 1) >> 1) { for (int k = 0; k < 4; k++) { if (++m[k] != 0) break; } }
}
static void mag_div_pow10(uint64_t m[4], int k) {
    if (k <= 0) return;
    if (k <= 19) { mag_div_pow10_once(m, OG_POW10[k]); return; }
    mag_div_pow10_once(m, OG_POW10[19]); mag_div_pow10_once(m, OG_POW10[k - 19]);
}
int32_t og_d64_mul(og_d128 *r, const int64_t *a, const int64_t *b, uint64_t n, int32_t c1, int32_t c2, int32_t scale1, int32_t scale2,
                   const uint64_t *n1, const uint64_t *n2, uint64_t *rnulls) {
    if ((c1 && bm_has(n1, 0)) || (c2 && bm_has(n2, 0))) { for (uint64_t i = 0; i < n; i++) bm_add(rnulls, i); return OG_RC_OK; }
    for (uint64_t w = 0; w < bm_words(n); w++) { if (!c1 && n1) rnulls[w] |= n1[w]; if (!c2 && n2) rnulls[w] |= n2[w]; }
    if (n & 63) rnulls[bm_words(n) - 1] &= (((uint64_t)1 << (n & 63)) - 1);
    const int adj = mul_scale_adj(scale1, scale2);
    for (uint64_t i = 0; i < n; i++) {
        if (bm_has(rnulls, i)) continue;
        int64_t x = a[c1 ? 0 : i], y = b[c2 ? 0 : i];
        uint64_t ax = x < 0 ? 0 - (uint64_t)x : (uint64_t)x, ay = y < 0 ? 0 - (uint64_t)y : (uint64_t)y;
        unsigned __int128 p = (unsigned __int128)ax * ay;
        uint64_t m[4] = {(uint64_t)p, (uint64_t)(p >> 64), 0, 0};
        mag_div_pow10(m, -adj);
        __int128 v = (__int128)(((unsigned __int128)m[1] << 64) | m[0]);
        if ((x < 0) != (y < 0)) v = -v;
        r[i] = d128_put(v);
    }
    return OG_RC_OK;
}
int32_t og_d128_addsub(int32_t is_sub, og_d128 *r, const og_d128 *a, const og_d128 *b, uint64_t n, int32_t c1, int32_t c2, int32_t scale1, int32_t scale2,
                       const uint64_t *n1, const uint64_t *n2, uint64_t *rnulls, int64_t *err_row) {
    if ((c1 && bm_has(n1, 0)) || (c2 && bm_has(n2, 0))) { for (uint64_t i = 0; i < n; i++) bm_add(rnulls, i); return OG_RC_OK; }
    for (uint64_t w = 0; w < bm_words(n); w++) { if (!c1 && n1) rnulls[w] |= n1[w]; if (!c2 && n2) rnulls[w] |= n2[w]; }
    if (n & 63) rnulls[bm_words(n) - 1] &= (((uint64_t)1 << (n & 63)) - 1);
    const int d1 = scale2 > scale1 ? scale2 - scale1 : 0, d2 = scale1 > scale2 ? scale1 - scale2 : 0;
    if (d1 > 19 || d2 > 19) return OG_RC_INVALID;
    const __int128 MAXV = (__int128)(((unsigned __int128)1 << 127) - 1);
    for (uint64_t i = 0; i < n; i++) {
        if (bm_has(rnulls, i)) continue;
        __int128 x = d128_get(a[c1 ? 0 : i]), y = d128_get(b[c2 ? 0 : i]);
        for (int side = 0; side < 2; side++) {   /* d128ScaleUp: |v| * 10^d must stay below 2^127 (d128Mul1Limb :4890-4906) */
            const int d = side ? d2 : d1; __int128 *v = side ? &y : &x;
            if (!d) continue;
            unsigned __int128 ab = *v < 0 ? (unsigned __int128)0 - (unsigned __int128)*v : (unsigned __int128)*v;
            unsigned __int128 lim = (((unsigned __int128)1 << 127) - 1) / OG_POW10[d];
            if (ab > lim) { if (err_row) *err_row = (int64_t)i; return OG_RC_INVALID_INPUT; }
            ab *= OG_POW10[d];
            *v = *v < 0 ? -(__int128)ab : (__int128)ab;
        }
        unsigned __int128 uz = is_sub ? (unsigned __int128)x - (unsigned __int128)y : (unsigned __int128)x + (unsigned __int128)y;
        __int128 z = (__int128)uz;
        int sx = x < 0, sy = y < 0, sz = z < 0;
        int ov = is_sub ? (sx != sy && sx != sz) : (sx == sy && sx != sz);
        r[i] = d128_put(z);
        if (ov) { if (err_row) *err_row = (int64_t)i; return OG_RC_INVALID_INPUT; }
        (void)MAXV;
    }
    return OG_RC_OK;
}
int32_t og_d128_mul(og_d128 *r, const og_d128 *a, const og_d128 *b, uint64_t n, int32_t c1, int32_t c2, int32_t scale1, int32_t scale2,
                    const uint64_t *n1, const uint64_t *n2, uint64_t *rnulls, int64_t *err_row) {
    if ((c1 && bm_has(n1, 0)) || (c2 && bm_has(n2, 0))) { for (uint64_t i = 0; i < n; i++) bm_add(rnulls, i); return OG_RC_OK; }
    for (uint64_t w = 0; w < bm_words(n); w++) { if (!c1 && n1) rnulls[w] |= n1[w]; if (!c2 && n2) rnulls[w] |= n2[w]; }
    if (n & 63) rnulls[bm_words(n) - 1] &= (((uint64_t)1 << (n & 63)) - 1);
    const int adj = mul_scale_adj(scale1, scale2);
    for (uint64_t i = 0; i < n; i++) {
        if (bm_has(rnulls, i)) continue;
        __int128 x = d128_get(a[c1 ? 0 : i]), y = d128_get(b[c2 ? 0 : i]);
        unsigned __int128 ax = x < 0 ? (unsigned __int128)0 - (unsigned __int128)x : (unsigned __int128)x;
        unsigned __int128 ay = y < 0 ? (unsigned __int128)0 - (unsigned __int128)y : (unsigned __int128)y;
        uint64_t xl = (uint64_t)ax, xh = (uint64_t)(ax >> 64), yl = (uint64_t)ay, yh = (uint64_t)(ay >> 64);
        uint64_t m[4] = {0, 0, 0, 0};
        unsigned __int128 t = (unsigned __int128)xl * yl; m[0] = (uint64_t)t; unsigned __int128 carry = t >> 64;
        t = (unsigned __int128)xl * yh + carry; unsigned __int128 t2 = (unsigned __int128)xh * yl + (uint64_t)t; m[1] = (uint64_t)t2;
        carry = (t >> 64) + (t2 >> 64);
        t = (unsigned __int128)xh * yh + carry; m[2] = (uint64_t)t; m[3] = (uint64_t)(t >> 64);
        mag_div_pow10(m, -adj);
        if (m[2] | m[3] | (m[1] >> 63)) { if (err_row) *err_row = (int64_t)i; return OG_RC_INVALID_INPUT; }
        __int128 v = (__int128)(((unsigned __int128)m[1] << 64) | m[0]);
        if ((x < 0) != (y < 0)) v = -v;
        r[i] = d128_put(v);
    }
    return OG_RC_OK;
}
/* grouped SUM / AVG accumulation of a decimal column into 128-bit sums + counts (groups == NULL: one group) */
void og_sum_d64(const int64_t *col, const uint64_t *nulls, uint64_t offset, const uint64_t *groups, uint64_t n, og_d128 *sums, int64_t *cnts) {
    for (uint64_t i = 0; i < n; i++) {
        uint64_t grp = groups ? groups[i] : 1; if (grp == 0) continue;
        if (bm_has(nulls, i + offset)) continue;
        sums[grp - 1] = d128_put((__int128)((unsigned __int128)d128_get(sums[grp - 1]) + (unsigned __int128)(__int128)col[i + offset]));
        cnts[grp - 1] += 1;
    }
}
void og_sum_d128(const og_d128 *col, const uint64_t *nulls, uint64_t offset, const uint64_t *groups, uint64_t n, og_d128 *sums, int64_t *cnts) {
    for (uint64_t i = 0; i < n; i++) {
        uint64_t grp = groups ? groups[i] : 1; if (grp == 0) continue;
        if (bm_has(nulls, i + offset)) continue;
        sums[grp - 1] = d128_put((__int128)((unsigned __int128)d128_get(sums[grp - 1]) + (unsigned __int128)d128_get(col[i + offset])));
        cnts[grp - 1] += 1;
    }
}

/* ---------------------------------------------------------------------------------------------
 * Synthetic column generators: the C twin of matrixone_b200/csrc/datagen.cu (and datagen.py), bit for bit, run on the
 * worker pool with the SAME block-range split the pipelines above use, so every worker first-touches the rows it will scan.
 * Data synthesis only -- used by bench.py's CPU legs so that they neither depend on the GPU library nor spend minutes in numpy.
 * ------------------------------------------------------------------------------------------- */
static inline uint64_t gen_mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline uint64_t gen_hash3(uint64_t seed, uint64_t stream, uint64_t row) { return gen_mix64(gen_mix64(seed ^ (stream * 0xD6E8FEB86659FD93ull)) + row); }

typedef struct { uint64_t seed, row0; int64_t i0, i1; int32_t *sd; double *qty, *price, *disc, *tax; uint8_t *rf, *ls; } genl_job;
static void *genl_worker(void *arg) {
    genl_job *j = (genl_job *)arg;
    for (int64_t i = j->i0; i < j->i1; i++) {
        const uint64_t r = j->row0 + (uint64_t)i;
        const int32_t sd = 8036 + (int32_t)(gen_hash3(j->seed, 1, r) % 2526ull);
        const uint64_t q = 1 + gen_hash3(j->seed, 2, r) % 50ull;
        const uint64_t cents = q * (90000ull + gen_hash3(j->seed, 3, r) % 120001ull);
        if (j->sd) j->sd[i] = sd;
        if (j->qty) j->qty[i] = (double)q;
        if (j->price) j->price[i] = (double)cents / 100.0;
        if (j->disc) j->disc[i] = (double)(gen_hash3(j->seed, 4, r) % 11ull) / 100.0;
        if (j->tax) j->tax[i] = (double)(gen_hash3(j->seed, 5, r) % 9ull) / 100.0;
        const uint64_t h6 = gen_hash3(j->seed, 6, r);
        const int32_t receipt = sd + 1 + (int32_t)(h6 % 30ull);
        if (j->rf) j->rf[i] = receipt <= 9298 ? (((h6 >> 32) & 1ull) ? 'R' : 'A') : 'N';
        if (j->ls) j->ls[i] = sd <= 9298 ? 'F' : 'O';
    }
    return NULL;
}
void og_gen_lineitem(uint64_t seed, uint64_t row0, int64_t n, int32_t nthreads, int32_t *sd, double *qty, double *price, double *disc,
                     double *tax, uint8_t *rf, uint8_t *ls) {
    if (nthreads < 1) nthreads = 1;
    int64_t nblocks = (n + BLOCK_ROWS - 1) / BLOCK_ROWS;
    if (nthreads > nblocks) nthreads = (int32_t)(nblocks > 0 ? nblocks : 1);
    genl_job *jobs = malloc(sizeof(genl_job) * (size_t)nthreads);
    int64_t per = (nblocks + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int64_t r0 = t * per * BLOCK_ROWS, r1 = r0 + per * BLOCK_ROWS; if (r0 > n) r0 = n; if (r1 > n) r1 = n;
        jobs[t] = (genl_job){seed, row0, r0, r1, sd, qty, price, disc, tax, rf, ls};
    }
    og_run(genl_worker, jobs, sizeof(genl_job), nthreads);
    free(jobs);
}

typedef struct { uint64_t seed, row0; int64_t i0, i1; int64_t *out; } geni_job;
static void *geni_worker(void *arg) {
    geni_job *j = (geni_job *)arg;
    for (int64_t i = j->i0; i < j->i1; i++) j->out[i] = (int64_t)(int32_t)(uint32_t)(gen_hash3(j->seed, 1, j->row0 + (uint64_t)i) & 0xffffffffull);
    return NULL;
}
void og_gen_int64(uint64_t seed, uint64_t row0, int64_t n, int32_t nthreads, int64_t *out) {
    if (nthreads < 1) nthreads = 1;
    int64_t nblocks = (n + BLOCK_ROWS - 1) / BLOCK_ROWS;
    if (nthreads > nblocks) nthreads = (int32_t)(nblocks > 0 ? nblocks : 1);
    geni_job *jobs = malloc(sizeof(geni_job) * (size_t)nthreads);
    int64_t per = (nblocks + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int64_t r0 = t * per * BLOCK_ROWS, r1 = r0 + per * BLOCK_ROWS; if (r0 > n) r0 = n; if (r1 > n) r1 = n;
        jobs[t] = (geni_job){seed, row0, r0, r1, out};
    }
    og_run(geni_worker, jobs, sizeof(geni_job), nthreads);
    free(jobs);
}

typedef struct { uint64_t seed, row0; int64_t i0, i1, dim; float *out; const float *centers; int64_t ncenters; float sigma; int32_t *comp; } genv_job;
static void *genv_worker(void *arg) {
    genv_job *j = (genv_job *)arg;
    for (int64_t i = j->i0; i < j->i1; i++) {
        const uint64_t r = j->row0 + (uint64_t)i;
        const uint64_t c = j->centers ? gen_hash3(j->seed, 7, r) % (uint64_t)j->ncenters : 0;
        if (j->comp) j->comp[i] = (int32_t)c;
        for (int64_t d = 0; d < j->dim; d++) {
            const uint64_t h = gen_hash3(j->seed, 16 + (uint64_t)d, r);
            const int32_t s = (int32_t)(h & 0xffff) + (int32_t)((h >> 16) & 0xffff) + (int32_t)((h >> 32) & 0xffff) + (int32_t)((h >> 48) & 0xffff) - 131070;
            float z = (float)s * 2.6428965e-05f;
            if (j->centers) { float sz = j->sigma * z; z = j->centers[c * (uint64_t)j->dim + (uint64_t)d] + sz; }
            j->out[i * j->dim + d] = z;
        }
    }
    return NULL;
}
/* comp (optional): the mixture component of every row (a valid IVF list assignment when centers are the centroids) */
void og_gen_vectors_f32(uint64_t seed, uint64_t row0, int64_t n, int64_t dim, int32_t nthreads, float *out, const float *centers,
                        int64_t ncenters, float sigma, int32_t *comp) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = (int32_t)(n > 0 ? n : 1);
    genv_job *jobs = malloc(sizeof(genv_job) * (size_t)nthreads);
    int64_t per = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; t++) {
        int64_t r0 = t * per, r1 = r0 + per; if (r0 > n) r0 = n; if (r1 > n) r1 = n;
        jobs[t] = (genv_job){seed, row0, r0, r1, dim, out, centers, ncenters, sigma, comp};
    }
    og_run(genv_worker, jobs, sizeof(genv_job), nthreads);
    free(jobs);
}

/* Kahan-compensated fp64 sum, used by tests to separate ordering noise from real bugs. */
double og_kahan_sum(const double *v, int64_t n) {
    double s = 0, c = 0;
    for (int64_t i = 0; i < n; i++) { double y = v[i] - c; volatile double t = s + y; c = (t - s) - y; s = t; }
    return s;
}

/* ---------------------------------------------------------------------------------------------
 * Hash join, equality conditions only (test infrastructure like everything in this file).
 * ------------------------------------------------------------------------------------------- */

/* og_join_sels: GroupSels.Insert + Finalize (pkg/vm/message/joinMapMsg.go:72-125) fed the way HashmapBuilder.BuildHashmap feeds it
 * (pkg/sql/colexec/hashbuild/hashmap.go:395-412: rows whose key is NULL or got no group are skipped; Insert(v - 1, row)).
 * offsets has group_count + 2 entries.  Returns the number of inserted rows. */
int64_t og_join_sels(const uint64_t *groups, int64_t n, int64_t group_count, int32_t *offsets, int32_t *vals) {
    int32_t *tmp = malloc(sizeof(int32_t) * 2 * (size_t)(n > 0 ? n : 1));
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        if (groups[i] == 0) continue;
        tmp[2 * m] = (int32_t)(groups[i] - 1); tmp[2 * m + 1] = (int32_t)i; m++;
    }
    for (int64_t i = 0; i < group_count + 2; i++) offsets[i] = 0;
    for (int64_t i = 0; i < m; i++) offsets[tmp[2 * i] + 1]++;                    /* count occurrences per group */
    for (int64_t i = 1; i < group_count + 2; i++) offsets[i] += offsets[i - 1];    /* prefix sum */
    for (int64_t i = 0; i < m; i++) { int32_t k = tmp[2 * i]; vals[offsets[k]] = tmp[2 * i + 1]; offsets[k]++; }   /* scatter, offsets as cursors */
    for (int64_t i = group_count + 1; i >= 1; i--) offsets[i] = offsets[i - 1];    /* recover: shift right by one */
    offsets[0] = 0;
    free(tmp);
    return m;
}

/* og_join_find: intHashMapIterator.Find (pkg/common/hashmap/iterator.go, inthashmap.go): 1-based group id of every key, 0 = absent; a NULL key
 * (zvals == 0) never matches.  Any exact map reproduces the ids (the reference's hash is seeded randomly). */
void og_join_find(const uint64_t *table_keys, int64_t ngroups, const uint64_t *keys, const uint64_t *nulls, int64_t n, uint64_t *vals) {
    uint64_t cap = 16;
    while (cap < 2 * (uint64_t)ngroups) cap <<= 1;
    int64_t *slot = malloc(sizeof(int64_t) * cap);
    for (uint64_t s = 0; s < cap; s++) slot[s] = -1;
    for (int64_t g = 0; g < ngroups; g++) {
        uint64_t h = table_keys[g] * 0x9E3779B97F4A7C15ull; h ^= h >> 29;
        uint64_t s = h & (cap - 1);
        while (slot[s] >= 0) s = (s + 1) & (cap - 1);
        slot[s] = g;
    }
    for (int64_t i = 0; i < n; i++) {
        vals[i] = 0;
        if (bm_has(nulls, (uint64_t)i)) continue;
        uint64_t h = keys[i] * 0x9E3779B97F4A7C15ull; h ^= h >> 29;
        uint64_t s = h & (cap - 1);
        while (slot[s] >= 0) { if (table_keys[slot[s]] == keys[i]) { vals[i] = (uint64_t)slot[s] + 1; break; } s = (s + 1) & (cap - 1); }
    }
    free(slot);
}

/* og_join_probe: the emission loop of container.probe (pkg/sql/colexec/hashjoin/join.go:383-628) with NonEqCond == nil, for
 * join_type 0 inner, 1 left outer (EmitUnmatchedProbe), 2 left semi, 3 left anti.  vals = Find's result; offsets == NULL: HashOnUnique
 * (build row = v - 1, join.go:431-446), else sels = GetSels(v - 1) appended in order (psSelsForOneRow, join.go:520-560).
 * Result rows are (probe row, build row | -1); returns how many there are (only the first cap are written). */
int64_t og_join_probe(const uint64_t *vals, int64_t n, const int32_t *offsets, const int32_t *sels, int32_t join_type,
                      int64_t *out_probe, int64_t *out_build, int64_t cap) {
    int64_t r = 0;
#define EMIT(p, b) do { if (r < cap) { out_probe[r] = (p); out_build[r] = (b); } r++; } while (0)
    for (int64_t row = 0; row < n; row++) {
        uint64_t v = vals[row];
        int64_t nmatch = 0;
        if (v) nmatch = offsets ? (int64_t)(offsets[v] - offsets[v - 1]) : 1;
        if (v == 0 || nmatch == 0) {                           /* z == 0 || v == 0 */
            if (join_type == 1 || join_type == 3) EMIT(row, -1);   /* appendOneNotMatch */
            continue;
        }
        if (join_type == 2) { EMIT(row, -1); continue; }           /* left semi: the probe row once */
        if (join_type == 3) continue;                              /* left anti: matched rows vanish */
        if (!offsets) { EMIT(row, (int64_t)v - 1); continue; }
        for (int32_t j = offsets[v - 1]; j < offsets[v]; j++) EMIT(row, (int64_t)sels[j]);
    }
#undef EMIT
    return r;
}

/* ---------------------------------------------------------------------------------------------
 * Elkan k-means, dense (non-spherical) variant: pkg/vectorindex/ivfflat/kmeans/elkans/clusterer.go.  distFn = metric.L2Distance
 * for every metric (ResolveKmeansDistanceFnForDense, distance_func.go:452-476).  Every step is deterministic given the initial centroids:
 * the per-vector loops are independent, recalculateCentroids sums the members serially in row order in T.  The reference's
 * thread pools only partition independent rows, so one thread reproduces them.  Layouts: vectors [n][dim], centroids [k][dim],
 * lower [n][k], upper [n], recompute uint8 [n], assign int64 [n], half [k][k], minhalf [k].
 * ------------------------------------------------------------------------------------------- */
#define KMEANS_IMPL(T, SFX, MAXT)                                                                                         \
    /* initBounds, clusterer.go:458-512 */                                                                               \
    void og_km_init_bounds_##SFX(const T *vec, int64_t n, int64_t dim, const T *cent, int64_t k, T *lower, T *upper, int64_t *assign) { \
        for (int64_t x = 0; x < n; x++) {                                                                                \
            T minDist = MAXT; int64_t closest = 0;                                                                        \
            for (int64_t c = 0; c < k; c++) {                                                                            \
                T dist = og_l2_##SFX(vec + x * dim, cent + c * dim, dim);                                                 \
                lower[x * k + c] = dist;                                                                                  \
                if (dist < minDist) { minDist = dist; closest = c; }                                                      \
            }                                                                                                             \
            upper[x] = minDist; assign[x] = closest;                                                                      \
        }                                                                                                                 \
    }                                                                                                                     \
    /* computeCentroidDistances, clusterer.go:516-575: 0.5 d(c, c') (diagonal untouched) and s(c) = min over c' != c */    \
    void og_km_centroid_dists_##SFX(const T *cent, int64_t k, int64_t dim, T *half, T *minhalf) {                        \
        for (int64_t i = 0; i < k; i++)                                                                                   \
            for (int64_t j = i + 1; j < k; j++) {                                                                        \
                T dist = og_l2_##SFX(cent + i * dim, cent + j * dim, dim);                                                \
                dist *= (T)0.5;                                                                                           \
                half[i * k + j] = dist; half[j * k + i] = dist;                                                           \
            }                                                                                                             \
        for (int64_t i = 0; i < k; i++) {                                                                                 \
            T cur = (T)3.40282346638528859811704183484516925440e+38; /* T(math.MaxFloat32) */                             \
            for (int64_t j = 0; j < k; j++) { if (i == j) continue; cur = (T)fmin((double)cur, (double)half[i * k + j]); } \
            minhalf[i] = cur;                                                                                             \
        }                                                                                                                 \
    }                                                                                                                     \
    /* assignData, clusterer.go:579-676; returns the number of changes */                                               \
    int64_t og_km_assign_##SFX(const T *vec, int64_t n, int64_t dim, const T *cent, int64_t k, const T *half, const T *minhalf, \
                               T *lower, T *upper, uint8_t *recompute, int64_t *assign) {                                 \
        int64_t changes = 0;                                                                                              \
        for (int64_t x = 0; x < n; x++) {                                                                                \
            if (upper[x] <= minhalf[assign[x]]) continue;                                                                 \
            for (int64_t c = 0; c < k; c++) {                                                                            \
                if (c != assign[x] && upper[x] > lower[x * k + c] && upper[x] > half[assign[x] * k + c]) {                \
                    T dxcx;                                                                                               \
                    if (recompute[x]) {                                                                                   \
                        recompute[x] = 0;                                                                                 \
                        dxcx = og_l2_##SFX(vec + x * dim, cent + assign[x] * dim, dim);                                   \
                        upper[x] = dxcx; lower[x * k + assign[x]] = dxcx;                                                 \
                        if (upper[x] <= lower[x * k + c]) continue;                                                       \
                        if (upper[x] <= half[assign[x] * k + c]) continue;                                                \
                    } else dxcx = upper[x];                                                                               \
                    if (dxcx > lower[x * k + c] || dxcx > half[assign[x] * k + c]) {                                      \
                        T dxc = og_l2_##SFX(vec + x * dim, cent + c * dim, dim);                                          \
                        lower[x * k + c] = dxc;                                                                           \
                        if (dxc < dxcx) { upper[x] = dxc; assign[x] = c; changes++; }                                     \
                    }                                                                                                     \
                }                                                                                                         \
            }                                                                                                             \
        }                                                                                                                 \
        return changes;                                                                                                   \
    }                                                                                                                     \
    /* recalculateCentroids, clusterer.go:679-727 (normalize == false).  An empty cluster takes dim values of the caller's rnd.Float32()   \
       stream (rnd[*rnd_used ...]); returns -1 when that stream runs dry. */                                              \
    int32_t og_km_recalc_##SFX(const T *vec, int64_t n, int64_t dim, const int64_t *assign, int64_t k, T *newc, int64_t *members, \
                               const float *rnd, int64_t rnd_len, int64_t *rnd_used) {                                    \
        for (int64_t c = 0; c < k; c++) members[c] = 0;                                                                   \
        for (int64_t i = 0; i < k * dim; i++) newc[i] = 0;                                                                \
        for (int64_t x = 0; x < n; x++) {                                                                                \
            int64_t cx = assign[x]; members[cx]++;                                                                        \
            for (int64_t i = 0; i < dim; i++) newc[cx * dim + i] += vec[x * dim + i];                                     \
        }                                                                                                                 \
        for (int64_t c = 0; c < k; c++) {                                                                                 \
            if (members[c] == 0) {                                                                                        \
                for (int64_t l = 0; l < dim; l++) { if (*rnd_used >= rnd_len) return -1; newc[c * dim + l] = (T)rnd[(*rnd_used)++]; } \
            } else {                                                                                                      \
                T scale = (T)1.0 / (T)members[c];              /* metric.ScaleInPlace(v, 1.0/T(count)): v[i] *= scale */   \
                for (int64_t i = 0; i < dim; i++) newc[c * dim + i] *= scale;                                             \
            }                                                                                                             \
        }                                                                                                                 \
        return 0;                                                                                                         \
    }                                                                                                                     \
    /* updateBounds, clusterer.go:730-762 */                                                                             \
    void og_km_update_bounds_##SFX(const T *cent, const T *newc, int64_t k, int64_t dim, int64_t n, const int64_t *assign, \
                                   T *lower, T *upper, uint8_t *recompute, T *shift) {                                    \
        for (int64_t c = 0; c < k; c++) shift[c] = og_l2_##SFX(cent + c * dim, newc + c * dim, dim);                     \
        for (int64_t x = 0; x < n; x++) {                                                                                \
            for (int64_t c = 0; c < k; c++) { T s = lower[x * k + c] - shift[c]; lower[x * k + c] = s > 0 ? s : (s != s ? s : (T)0); /* T(math.Max(float64(s), 0)): -0 -> +0 */ } \
            upper[x] += shift[assign[x]]; recompute[x] = 1;                                                               \
        }                                                                                                                 \
    }                                                                                                                     \
    /* Cluster + elkansCluster, clusterer.go:330-392, from given initial centroids (InitCentroids is the caller's: it draws from Go's    \
       PCG).  cent [k][dim] in/out; assign out.  Returns the number of iterations run, or -1 (rnd stream dry). */          \
    int64_t og_km_cluster_##SFX(const T *vec, int64_t n, int64_t dim, T *cent, int64_t k, int64_t max_iter, const float *rnd, int64_t rnd_len, \
                                int64_t *assign) {                                                                        \
        T *lower = malloc(sizeof(T) * (size_t)(n * k)), *upper = malloc(sizeof(T) * (size_t)n), *half = calloc((size_t)(k * k), sizeof(T)); \
        T *minhalf = malloc(sizeof(T) * (size_t)k), *next = malloc(sizeof(T) * (size_t)(k * dim)), *shift = malloc(sizeof(T) * (size_t)k); \
        uint8_t *recompute = malloc((size_t)n); int64_t *members = malloc(sizeof(int64_t) * (size_t)k);                   \
        int64_t used = 0, iter = 0; int bad = 0;                                                                          \
        for (int64_t x = 0; x < n; x++) recompute[x] = 1;       /* NewKMeans: recompute = true, clusterer.go */           \
        og_km_init_bounds_##SFX(vec, n, dim, cent, k, lower, upper, assign);                                              \
        for (;; iter++) {                                                                                                 \
            og_km_centroid_dists_##SFX(cent, k, dim, half, minhalf);                                                      \
            int64_t changes = og_km_assign_##SFX(vec, n, dim, cent, k, half, minhalf, lower, upper, recompute, assign);   \
            if (og_km_recalc_##SFX(vec, n, dim, assign, k, next, members, rnd, rnd_len, &used)) { bad = 1; break; }       \
            og_km_update_bounds_##SFX(cent, next, k, dim, n, assign, lower, upper, recompute, shift);                     \
            memcpy(cent, next, sizeof(T) * (size_t)(k * dim));  /* km.centroids, km.nextCentroids = newCentroids, km.centroids */ \
            if (iter != 0 && (iter == max_iter || changes == 0)) break;                                                   \
        }                                                                                                                 \
        free(lower); free(upper); free(half); free(minhalf); free(next); free(shift); free(recompute); free(members);     \
        return bad ? -1 : iter + 1;                                                                                       \
    }
KMEANS_IMPL(float, f32, 3.40282346638528859811704183484516925440e+38f)
KMEANS_IMPL(double, f64, 1.79769313486231570814527423731704356798070e+308)

/* og_lz4_decode_block: LZ4 block format decoder (what compress.Decompress / lz4.UncompressBlock does, pkg/compress/compress.go:37-47), restated
 * from the published block format.  Returns the decoded size or -1 (malformed / does not fit).  Test infrastructure. */
int64_t og_lz4_decode_block(const uint8_t *src, int64_t sl, uint8_t *dst, int64_t dcap) {
    int64_t ip = 0, op = 0;
    while (ip < sl) {
        unsigned token = src[ip++];
        int64_t lit = token >> 4;
        if (lit == 15) { unsigned e; do { if (ip >= sl) return -1; e = src[ip++]; lit += e; } while (e == 255); }
        if (ip + lit > sl || op + lit > dcap) return -1;
        memcpy(dst + op, src + ip, (size_t)lit); ip += lit; op += lit;
        if (ip >= sl) break;
        if (ip + 2 > sl) return -1;
        int64_t offset = (int64_t)src[ip] | ((int64_t)src[ip + 1] << 8); ip += 2;
        int64_t mlen = token & 15;
        if (mlen == 15) { unsigned e; do { if (ip >= sl) return -1; e = src[ip++]; mlen += e; } while (e == 255); }
        mlen += 4;
        if (offset == 0 || offset > op || op + mlen > dcap) return -1;
        for (int64_t i = 0; i < mlen; i++) dst[op + i] = dst[op - offset + i];   /* byte-serial: overlapping matches replicate */
        op += mlen;
    }
    return op;
}
