// godist.cuh -- bit-exact device restatements of the Go metric loops (pkg/vectorindex/metric/distance_func.go), shared by the
// row-wise XCall kernels (distance.cu) and the candidate re-scoring of the tensor-core search path (tcsearch.cu).
// One warp per (p, q) pair: lanes own 8-element chunks, the serial `sum += chunk` chain is replayed in order with shuffles.
#pragma once
#include "common.cuh"
#include <cstring>

namespace mob {
namespace godist {

constexpr unsigned FULL = 0xffffffffu;

template <typename T, int N>
__device__ __forceinline__ void load_elems(const uint8_t *p, T *out, bool aligned, bool stream) {
    if (aligned) {
#pragma unroll
        for (int i = 0; i < (int)(N * sizeof(T)) / 16; i++) {
            int4 v = stream ? ld_stream16(p + 16 * i) : __ldg(reinterpret_cast<const int4 *>(p + 16 * i));
            memcpy(reinterpret_cast<char *>(out) + 16 * i, &v, 16);
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) {
            T v; uint8_t b[sizeof(T)];
#pragma unroll
            for (int k = 0; k < (int)sizeof(T); k++) b[k] = p[i * sizeof(T) + k];
            memcpy(&v, b, sizeof(T)); out[i] = v;
        }
    }
}
template <typename T> __device__ __forceinline__ T load1(const uint8_t *p) {
    T v; uint8_t b[sizeof(T)];
#pragma unroll
    for (int k = 0; k < (int)sizeof(T); k++) b[k] = p[k];
    memcpy(&v, b, sizeof(T)); return v;
}

__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }

// replay "sum = sum + c[l]" for l = 0..m-1 in lane order; all lanes end with the same sum
template <typename T>
__device__ __forceinline__ T chain(T sum, T c, int m) {
    if (m == 32) {
#pragma unroll
        for (int l = 0; l < 32; l++) sum = add_rn(sum, __shfl_sync(FULL, c, l));
    } else {
        for (int l = 0; l < m; l++) sum = add_rn(sum, __shfl_sync(FULL, c, l));
    }
    return sum;
}

// L2DistanceSq, distance_func.go:59-95
template <typename T>
__device__ T go_l2sq(const uint8_t *p, const uint8_t *q, int dim, int lane, bool al, bool sp, bool sq) {
    const int nch = dim >> 3;
    T sum = 0;
    for (int base = 0; base < nch; base += 32) {
        const int c = base + lane;
        T s = 0;
        if (c < nch) {
            T a[8], b[8], t[8];
            load_elems<T, 8>(p + (size_t)c * 8 * sizeof(T), a, al, sp);
            load_elems<T, 8>(q + (size_t)c * 8 * sizeof(T), b, al, sq);
#pragma unroll
            for (int j = 0; j < 8; j++) { T d = sub_rn(a[j], b[j]); t[j] = mul_rn(d, d); }
            s = add_rn(add_rn(add_rn(add_rn(t[0], t[1]), add_rn(t[2], t[3])), add_rn(t[4], t[5])), add_rn(t[6], t[7]));
        }
        sum = chain(sum, s, min(32, nch - base));
    }
    for (int i = nch << 3; i < dim; i++) {  // remainder loop, distance_func.go:88-92
        T d = sub_rn(load1<T>(p + (size_t)i * sizeof(T)), load1<T>(q + (size_t)i * sizeof(T)));
        sum = add_rn(sum, mul_rn(d, d));
    }
    return sum;
}

// InnerProduct, distance_func.go:172-205 (returns -sum)
template <typename T>
__device__ T go_ip(const uint8_t *p, const uint8_t *q, int dim, int lane, bool al, bool sp, bool sq) {
    const int nch = dim >> 3;
    T sum = 0;
    for (int base = 0; base < nch; base += 32) {
        const int c = base + lane;
        T s = 0;
        if (c < nch) {
            T a[8], b[8];
            load_elems<T, 8>(p + (size_t)c * 8 * sizeof(T), a, al, sp);
            load_elems<T, 8>(q + (size_t)c * 8 * sizeof(T), b, al, sq);
            s = add_rn(mul_rn(a[0], b[0]), mul_rn(a[1], b[1]));
#pragma unroll
            for (int j = 2; j < 8; j++) s = add_rn(s, mul_rn(a[j], b[j]));
        }
        sum = chain(sum, s, min(32, nch - base));
    }
    for (int i = nch << 3; i < dim; i++)
        sum = add_rn(sum, mul_rn(load1<T>(p + (size_t)i * sizeof(T)), load1<T>(q + (size_t)i * sizeof(T))));
    return -sum;
}

// shared accumulation of CosineDistance / CosineSimilarity, distance_func.go:216-262
template <typename T>
__device__ void go_cos_parts(const uint8_t *p, const uint8_t *q, int dim, int lane, bool al, bool sp, bool sq, T &dp, T &n1, T &n2) {
    const int nch = dim >> 2;
    dp = 0; n1 = 0; n2 = 0;
    for (int base = 0; base < nch; base += 32) {
        const int c = base + lane;
        T s0 = 0, s1 = 0, s2 = 0;
        if (c < nch) {
            T a[4], b[4];
            if (sizeof(T) == 4) { load_elems<T, 4>(p + (size_t)c * 4 * sizeof(T), a, al, sp); load_elems<T, 4>(q + (size_t)c * 4 * sizeof(T), b, al, sq); }
            else { load_elems<T, 4>(p + (size_t)c * 4 * sizeof(T), a, al, sp); load_elems<T, 4>(q + (size_t)c * 4 * sizeof(T), b, al, sq); }
            s0 = add_rn(add_rn(add_rn(mul_rn(a[0], b[0]), mul_rn(a[1], b[1])), mul_rn(a[2], b[2])), mul_rn(a[3], b[3]));
            s1 = add_rn(add_rn(add_rn(mul_rn(a[0], a[0]), mul_rn(a[1], a[1])), mul_rn(a[2], a[2])), mul_rn(a[3], a[3]));
            s2 = add_rn(add_rn(add_rn(mul_rn(b[0], b[0]), mul_rn(b[1], b[1])), mul_rn(b[2], b[2])), mul_rn(b[3], b[3]));
        }
        const int m = min(32, nch - base);
        dp = chain(dp, s0, m); n1 = chain(n1, s1, m); n2 = chain(n2, s2, m);
    }
    for (int i = nch << 2; i < dim; i++) {
        T x = load1<T>(p + (size_t)i * sizeof(T)), y = load1<T>(q + (size_t)i * sizeof(T));
        dp = add_rn(dp, mul_rn(x, y)); n1 = add_rn(n1, mul_rn(x, x)); n2 = add_rn(n2, mul_rn(y, y));
    }
}

enum Kind { K_XC_L2 = 0, K_XC_L2SQ, K_GO_L2, K_GO_L2SQ, K_GO_IP, K_GO_COSDIST, K_GO_COSSIM, K_GO_L1, K_GO_NORM };

// ---- lane-per-row batches over a per-warp cp.async ring (see distance.cu for the mapping) ---------------------------------------
constexpr int kTileRows = 32, kSliceBytes = 128, kPitch = 144;   // 144-byte row pitch: LDS.128 by the row owners is conflict-free
constexpr int kTileBytes = kTileRows * kPitch;
template <bool TWO> struct RingCfg {
    static constexpr int kStages = TWO ? 3 : 5;
    static constexpr int kStageBytes = TWO ? 2 * kTileBytes : kTileBytes + kPitch;   // [row tile][second row tile | const slice]
};

__device__ __forceinline__ void cp_async16(unsigned dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ const uint8_t *shfl_ptr(const uint8_t *p, int src) {
    unsigned long long v = (unsigned long long)(uintptr_t)p;
    unsigned lo = __shfl_sync(FULL, (unsigned)v, src), hi = __shfl_sync(FULL, (unsigned)(v >> 32), src);
    return (const uint8_t *)(uintptr_t)(((unsigned long long)hi << 32) | lo);
}
// 16 bytes global -> shared for a row whose base is not 16-byte aligned (byte loads: varlena offsets are arbitrary)
__device__ __forceinline__ void copy16_unaligned(unsigned char *dst, const uint8_t *src) {
#pragma unroll
    for (int k = 0; k < 16; k++) dst[k] = src[k];
}

// per-row accumulators of one kind; add_slice consumes EPS = 128 / sizeof(T) consecutive elements of both operands
template <typename T, int KIND> struct RowAcc {
    static constexpr bool kCos = KIND == K_GO_COSDIST || KIND == K_GO_COSSIM;
    static constexpr bool kXc = KIND == K_XC_L2 || KIND == K_XC_L2SQ;
    static constexpr bool kOne = KIND == K_GO_NORM;   // one-operand kind: the q side is never staged nor read
    static constexpr int CH = kCos ? 4 : 8;          // chunk of the Go loop (distance_func.go: 8-way, cosine 4-way)
    T sum = 0, n1 = 0, n2 = 0; double dsum = 0.0;
    __device__ __forceinline__ void chunk(const T *a, const T *b) {   // CH elements, exact association of the Go source
        if (kXc) {
#pragma unroll
            for (int j = 0; j < CH; j++) { T d = sub_rn(a[j], b[j]); dsum = __dadd_rn(dsum, (double)mul_rn(d, d)); }
        } else if (kOne) {   // NormalizeL2, distance_func.go:417-421: sumSquares += float64(val) * float64(val), strictly in index order
#pragma unroll
            for (int j = 0; j < CH; j++) dsum = __dadd_rn(dsum, __dmul_rn((double)a[j], (double)a[j]));
        } else if (kCos) {
            sum = add_rn(sum, add_rn(add_rn(add_rn(mul_rn(a[0], b[0]), mul_rn(a[1], b[1])), mul_rn(a[2], b[2])), mul_rn(a[3], b[3])));
            n1 = add_rn(n1, add_rn(add_rn(add_rn(mul_rn(a[0], a[0]), mul_rn(a[1], a[1])), mul_rn(a[2], a[2])), mul_rn(a[3], a[3])));
            n2 = add_rn(n2, add_rn(add_rn(add_rn(mul_rn(b[0], b[0]), mul_rn(b[1], b[1])), mul_rn(b[2], b[2])), mul_rn(b[3], b[3])));
        } else if (KIND == K_GO_L1) {   // L1Distance, distance_func.go:112-154: eight serial  sum += abs(p[j] - q[j])
#pragma unroll
            for (int j = 0; j < 8; j++) { T d = sub_rn(a[j], b[j]); sum = add_rn(sum, d < 0 ? -d : d); }
        } else if (KIND == K_GO_IP) {
            T c = add_rn(mul_rn(a[0], b[0]), mul_rn(a[1], b[1]));
#pragma unroll
            for (int j = 2; j < 8; j++) c = add_rn(c, mul_rn(a[j], b[j]));
            sum = add_rn(sum, c);
        } else {
            T t[8];
#pragma unroll
            for (int j = 0; j < 8; j++) { T d = sub_rn(a[j], b[j]); t[j] = mul_rn(d, d); }
            sum = add_rn(sum, add_rn(add_rn(add_rn(add_rn(t[0], t[1]), add_rn(t[2], t[3])), add_rn(t[4], t[5])), add_rn(t[6], t[7])));
        }
    }
    __device__ __forceinline__ void elem(T a, T b) {   // remainder loops of the Go functions
        if (kXc) { T d = sub_rn(a, b); dsum = __dadd_rn(dsum, (double)mul_rn(d, d)); }
        else if (kOne) dsum = __dadd_rn(dsum, __dmul_rn((double)a, (double)a));
        else if (kCos) { sum = add_rn(sum, mul_rn(a, b)); n1 = add_rn(n1, mul_rn(a, a)); n2 = add_rn(n2, mul_rn(b, b)); }
        else if (KIND == K_GO_IP) sum = add_rn(sum, mul_rn(a, b));
        else if (KIND == K_GO_L1) { T d = sub_rn(a, b); sum = add_rn(sum, d < 0 ? -d : d); }
        else { T d = sub_rn(a, b); sum = add_rn(sum, mul_rn(d, d)); }
    }
};

// One batch of 32 rows: lane r owns row r (px / pq = its operands, dim elements, good = the row takes part).  Returns the lane's
// accumulators after the whole row (slices through the ring, then the lane's own tail: whole chunks past the last full slice
// and the element remainder loop).  `ring` = this warp's RingCfg<TWO>::kStages * kStageBytes bytes of shared memory.
template <typename T, int KIND, bool TWO>
__device__ __forceinline__ RowAcc<T, KIND> row_batch(unsigned char *ring, int lane, const uint8_t *px, const uint8_t *pq, int dim, bool good) {
    using Cfg = RingCfg<TWO>;
    using Acc = RowAcc<T, KIND>;
    constexpr int NST = Cfg::kStages, EPS = kSliceBytes / (int)sizeof(T), CH = Acc::CH;
    const unsigned ring_u32 = (unsigned)__cvta_generic_to_shared(ring);
    const int piece = lane & 7, rgrp = lane >> 3;
    if (!good) dim = 0;
    const int nfull = (dim / CH) * CH;                                  // elements covered by whole chunks
    const int nsl = (int)(((size_t)nfull * sizeof(T)) / kSliceBytes);    // whole 128-byte slices of this row
    const bool alx = (((uintptr_t)px) & 15) == 0, alq = (((uintptr_t)pq) & 15) == 0;
    const int maxsl = __reduce_max_sync(FULL, nsl);
    // loader view: this lane copies piece `piece` of rows rgrp, rgrp + 4, ..., rgrp + 28
    const uint8_t *lx[8], *lq[8]; int lnsl[8]; unsigned lal = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int rr = rgrp + 4 * k;
        lx[k] = shfl_ptr(px, rr); lnsl[k] = __shfl_sync(FULL, nsl, rr);
        if (__shfl_sync(FULL, (int)alx, rr)) lal |= 1u << k;
        if (TWO) { lq[k] = shfl_ptr(pq, rr); if (__shfl_sync(FULL, (int)alq, rr)) lal |= 1u << (8 + k); }
    }
    // the const side: taken from the first row that has one (all rows share it)
    const unsigned goodmask = __ballot_sync(FULL, good);
    const int qsrc = goodmask ? __ffs((int)goodmask) - 1 : 0;
    const uint8_t *cq = TWO ? nullptr : shfl_ptr(pq, qsrc);
    const bool cq_al = TWO ? false : (__shfl_sync(FULL, (int)alq, qsrc) != 0);

    auto issue = [&](int s) {   // slice s of every row -> stage s % NST; always commits one group
        if (s < maxsl) {
            const unsigned sb = ring_u32 + (unsigned)((s % NST) * Cfg::kStageBytes);
            unsigned char *sp = ring + (size_t)(s % NST) * Cfg::kStageBytes;
            const size_t off = (size_t)s * kSliceBytes + (size_t)piece * 16;
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const int rr = rgrp + 4 * k;
                if (s < lnsl[k]) {
                    if (lal & (1u << k)) cp_async16(sb + rr * kPitch + piece * 16, lx[k] + off);
                    else copy16_unaligned(sp + rr * kPitch + piece * 16, lx[k] + off);
                    if (TWO) {
                        if (lal & (1u << (8 + k))) cp_async16(sb + kTileBytes + rr * kPitch + piece * 16, lq[k] + off);
                        else copy16_unaligned(sp + kTileBytes + rr * kPitch + piece * 16, lq[k] + off);
                    }
                }
            }
            if (!TWO && !Acc::kOne && lane < 8) {
                if (cq_al) cp_async16(sb + kTileBytes + lane * 16, cq + (size_t)s * kSliceBytes + (size_t)lane * 16);
                else copy16_unaligned(sp + kTileBytes + lane * 16, cq + (size_t)s * kSliceBytes + (size_t)lane * 16);
            }
        }
        cp_async_commit();
    };

    Acc acc;
#pragma unroll 1
    for (int s = 0; s < NST - 1; s++) issue(s);
#pragma unroll 1
    for (int s = 0; s < maxsl; s++) {
        issue(s + NST - 1);            // its stage was consumed in the previous iteration (ordered by the __syncwarp below)
        cp_async_wait<NST - 1>();      // all but the NST - 1 newest groups are complete => slice s has landed
        __syncwarp();
        if (s < nsl) {
            const unsigned char *sp = ring + (size_t)(s % NST) * Cfg::kStageBytes;
            const unsigned char *ra = sp + lane * kPitch;
            const unsigned char *rb = TWO ? sp + kTileBytes + lane * kPitch : sp + kTileBytes;   // const slice: broadcast reads
#pragma unroll
            for (int h = 0; h < 2; h++) {     // half a slice at a time keeps the register footprint down
                T a[EPS / 2], b[EPS / 2];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int4 va = *reinterpret_cast<const int4 *>(ra + h * 64 + j * 16);
                    const int4 vb = *reinterpret_cast<const int4 *>(rb + h * 64 + j * 16);
                    memcpy(reinterpret_cast<char *>(a) + 16 * j, &va, 16);
                    memcpy(reinterpret_cast<char *>(b) + 16 * j, &vb, 16);
                }
#pragma unroll
                for (int c = 0; c < EPS / 2; c += CH) acc.chunk(a + c, b + c);
            }
        }
        __syncwarp();
    }
    cp_async_wait<0>();
    // every lane finishes its own row: whole chunks past the last full slice, then the element remainder loop
    if (good) {
        const bool al = alx && alq;
        for (int c = nsl * EPS; c < nfull; c += CH) {
            T a[CH], b[CH];
            load_elems<T, CH>(px + (size_t)c * sizeof(T), a, al && CH * sizeof(T) >= 16, false);
            load_elems<T, CH>(pq + (size_t)c * sizeof(T), b, al && CH * sizeof(T) >= 16, false);
            acc.chunk(a, b);
        }
        for (int e = nfull; e < dim; e++) acc.elem(load1<T>(px + (size_t)e * sizeof(T)), load1<T>(pq + (size_t)e * sizeof(T)));
    }
    return acc;
}

}  // namespace godist
}  // namespace mob
