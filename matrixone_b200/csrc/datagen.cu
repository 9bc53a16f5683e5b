// datagen.cu -- counter-based synthetic column generators for bench.py and the GPU tests.
// Every value is a pure function of (seed, stream, row) through splitmix64, evaluated with integer arithmetic and at
// most one IEEE division/multiplication, so matrixone_b200/datagen.py reproduces the columns bit for bit with numpy
// (the CPU oracle is fed the numpy twin; the GPU never needs a host copy of a 16 GB column set).
//
// lineitem shape follows TPC-H dbgen as SURVEY.md section 8(d) prescribes: l_shipdate uniform over 2526 days starting
// 1992-01-02 (days since 1970-01-01), l_quantity 1..50, l_discount 0.00..0.10, l_tax 0.00..0.08,
// l_extendedprice = quantity * U[900.00, 2100.00] in cents, l_returnflag/l_linestatus correlated with dates the way
// dbgen does (R/A when received by 1995-06-17 else N; F when shipped by 1995-06-17 else O).
#include "common.cuh"

using namespace mob;

namespace {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__host__ __device__ __forceinline__ uint64_t hash3(uint64_t seed, uint64_t stream, uint64_t row) {
    return mix64(mix64(seed ^ (stream * 0xD6E8FEB86659FD93ull)) + row);
}

constexpr int32_t kDate19920102 = 8036;   // days since 1970-01-01
constexpr int32_t kDate19950617 = 9298;

__global__ void gen_lineitem_kernel(uint64_t seed, uint64_t row0, uint64_t n, int32_t *shipdate, double *quantity, double *extendedprice,
                                    double *discount, double *tax, uint8_t *returnflag, uint8_t *linestatus) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = row0 + i;
        const int32_t sd = kDate19920102 + (int32_t)(hash3(seed, 1, r) % 2526ull);
        const uint64_t q = 1 + hash3(seed, 2, r) % 50ull;
        const uint64_t cents = q * (90000ull + hash3(seed, 3, r) % 120001ull);
        if (shipdate) shipdate[i] = sd;
        if (quantity) quantity[i] = (double)q;
        if (extendedprice) extendedprice[i] = (double)cents / 100.0;
        if (discount) discount[i] = (double)(hash3(seed, 4, r) % 11ull) / 100.0;
        if (tax) tax[i] = (double)(hash3(seed, 5, r) % 9ull) / 100.0;
        const uint64_t h6 = hash3(seed, 6, r);
        const int32_t receipt = sd + 1 + (int32_t)(h6 % 30ull);
        if (returnflag) returnflag[i] = receipt <= kDate19950617 ? (((h6 >> 32) & 1ull) ? 'R' : 'A') : 'N';
        if (linestatus) linestatus[i] = sd <= kDate19950617 ? 'F' : 'O';
    }
}

__global__ void gen_int64_kernel(uint64_t seed, uint64_t row0, uint64_t n, int64_t *out, uint64_t *nulls, uint32_t null_per_mille) {
    // one thread per 64 rows so the nulls word is written without atomics; row0 must be a multiple of 64
    const uint64_t nwords = (n + 63) >> 6;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < nwords; w += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t word = 0;
        for (int j = 0; j < 64; j++) {
            const uint64_t i = (w << 6) + j;
            if (i >= n) break;
            const uint64_t r = row0 + i;
            out[i] = (int64_t)(int32_t)(uint32_t)(hash3(seed, 1, r) & 0xffffffffull);   // uniform in [-2^31, 2^31)
            if (null_per_mille && hash3(seed, 2, r) % 1000ull < null_per_mille) word |= 1ull << j;
        }
        if (nulls) nulls[w] = word;
    }
}

__global__ void gen_vectors_kernel(uint64_t seed, uint64_t row0, uint64_t n, int64_t dim, float *out, const float *centers,
                                   int64_t ncenters, float sigma) {
    const uint64_t total = n * (uint64_t)dim;
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = e / (uint64_t)dim, j = e % (uint64_t)dim, r = row0 + i;
        const uint64_t h = hash3(seed, 16 + j, r);
        // Irwin-Hall(4) of 16-bit uniforms: integer sum, one float multiply -> approximately N(0,1), exactly reproducible
        const int32_t s = (int32_t)(h & 0xffff) + (int32_t)((h >> 16) & 0xffff) + (int32_t)((h >> 32) & 0xffff) + (int32_t)((h >> 48) & 0xffff) - 131070;
        float z = __fmul_rn((float)s, 2.6428965e-05f);   // 1 / sqrt(4 * (65536^2 - 1) / 12)
        if (centers) {
            const uint64_t c = hash3(seed, 7, r) % (uint64_t)ncenters;
            z = __fadd_rn(centers[c * (uint64_t)dim + j], __fmul_rn(sigma, z));
        }
        out[e] = z;
    }
}

__global__ void gather_rows_f32_kernel(const float *__restrict__ src, const int64_t *__restrict__ idx, uint64_t m, int64_t dim, float *__restrict__ dst) {
    const uint64_t total = m * (uint64_t)dim;
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x)
        dst[e] = src[(uint64_t)idx[e / (uint64_t)dim] * (uint64_t)dim + e % (uint64_t)dim];
}

}  // namespace

extern "C" {

int32_t MoB200_GenLineitem(uint64_t seed, uint64_t row0, uint64_t n, int32_t *shipdate, double *quantity, double *extendedprice,
                           double *discount, double *tax, uint8_t *returnflag, uint8_t *linestatus) {
    ThreadCtx &t = tctx();
    if (!t.ready) return MO_RC_INTERNAL_ERROR;
    if (n == 0) return MO_RC_SUCCESS;
    gen_lineitem_kernel<<<num_sms() * 8, 256, 0, t.stream>>>(seed, row0, n, shipdate, quantity, extendedprice, discount, tax, returnflag, linestatus);
    MOB_LAUNCH_CHECK();
    MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
    return MO_RC_SUCCESS;
}

int32_t MoB200_GenInt64(uint64_t seed, uint64_t row0, uint64_t n, int64_t *out, uint64_t *nulls, uint32_t null_per_mille) {
    ThreadCtx &t = tctx();
    if (!t.ready) return MO_RC_INTERNAL_ERROR;
    if (n == 0) return MO_RC_SUCCESS;
    if (row0 & 63) { set_error("GenInt64: row0 must be a multiple of 64"); return MO_RC_INVALID_ARGUMENT; }
    gen_int64_kernel<<<num_sms() * 8, 256, 0, t.stream>>>(seed, row0, n, out, nulls, null_per_mille);
    MOB_LAUNCH_CHECK();
    MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
    return MO_RC_SUCCESS;
}

int32_t MoB200_GenVectorsF32(uint64_t seed, uint64_t row0, uint64_t n, int64_t dim, float *out, const float *centers,
                             int64_t ncenters, float sigma) {
    ThreadCtx &t = tctx();
    if (!t.ready) return MO_RC_INTERNAL_ERROR;
    if (n == 0) return MO_RC_SUCCESS;
    search_invalidate(out, n * (uint64_t)dim * 4);
    gen_vectors_kernel<<<num_sms() * 16, 256, 0, t.stream>>>(seed, row0, n, dim, out, centers, ncenters, sigma);
    MOB_LAUNCH_CHECK();
    MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
    return MO_RC_SUCCESS;
}

// dst[i] = src[idx[i]] for whole rows (device pointers): lays an IVF dataset out list by list (ivfflat index build)
int32_t MoB200_GatherRowsF32(float *dst, const float *src, const int64_t *idx, uint64_t m, int64_t dim) {
    ThreadCtx &t = tctx();
    if (!t.ready) return MO_RC_INTERNAL_ERROR;
    if (m == 0) return MO_RC_SUCCESS;
    search_invalidate(dst, m * (uint64_t)dim * 4);
    gather_rows_f32_kernel<<<num_sms() * 16, 256, 0, t.stream>>>(src, idx, m, dim, dst);
    MOB_LAUNCH_CHECK();
    MOB_CUDA_TRY(cudaStreamSynchronize(t.stream));
    return MO_RC_SUCCESS;
}

}  // extern "C"
