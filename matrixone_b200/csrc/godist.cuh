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

}  // namespace godist
}  // namespace mob
