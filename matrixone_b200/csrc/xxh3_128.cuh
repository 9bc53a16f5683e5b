// xxh3_128.cuh -- XXH3 128-bit hash with seed (xxHash 0.8.3, the version the reference pins: thirdparties/Makefile:26), restated for device code.
//
// The reference's bloom filter derives its k bit positions from the two halves of XXH3_128bits_withSeed(key, len, seed)
// (cgo/bloom.c:31-66); xxHash is a third-party dependency that is not part of /root/reference's own sources (only its tarball is), so
// this file restates the PUBLISHED algorithm (xxHash doc/xxhash_spec.md, "XXH3 algorithm overview") -- all four input-length classes --
// from the specification's formulas.  It is pinned to the real library: oracle/build.py compiles the reference's bloom.c + the tarball's
// xxhash.h into oracle/_ref/libbloom_ref.so, and tests/test_xxh3_host.py compares this header (compiled for the host) with it on every
// length 0..1100 and on the fixed-width integer path.  Host + device: every function is MOB_HD.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define MOB_HD __host__ __device__ __forceinline__
#else
#define MOB_HD inline
#endif

namespace mob {
namespace xxh3 {

struct Hash128 { uint64_t lo, hi; };

constexpr uint64_t P32_1 = 0x9E3779B1ull, P32_2 = 0x85EBCA77ull, P32_3 = 0xC2B2AE3Dull;
constexpr uint64_t P64_1 = 0x9E3779B185EBCA87ull, P64_2 = 0xC2B2AE3D27D4EB4Full, P64_3 = 0x165667B19E3779F9ull,
                   P64_4 = 0x85EBCA77C2B2AE63ull, P64_5 = 0x27D4EB2F165667C5ull;
constexpr uint64_t PMX1 = 0x165667919E3779F9ull, PMX2 = 0x9FB21C651E98DF25ull;

// the 192-byte default secret as 24 little-endian words (spec: "kSecret"); unaligned reads are assembled from two words
#define MOB_XXH3_SECRET_WORDS { \
    0xbe4ba423396cfeb8ull, 0x1cad21f72c81017cull, 0xdb979083e96dd4deull, 0x1f67b3b7a4a44072ull, 0x78e5c0cc4ee679cbull, 0x2172ffcc7dd05a82ull, \
    0x8e2443f7744608b8ull, 0x4c263a81e69035e0ull, 0xcb00c391bb52283cull, 0xa32e531b8b65d088ull, 0x4ef90da297486471ull, 0xd8acdea946ef1938ull, \
    0x3f349ce33f76faa8ull, 0x1d4f0bc7c7bbdcf9ull, 0x3159b4cd4be0518aull, 0x647378d9c97e9fc8ull, 0xc3ebd33483acc5eaull, 0xeb6313faffa081c5ull, \
    0x49daf0b751dd0d17ull, 0x9e68d429265516d3ull, 0xfca1477d58be162bull, 0xce31d07ad1b8f88full, 0x280416958f3acb45ull, 0x7e404bbbcafbd7afull }
#if defined(__CUDACC__)
static __device__ __constant__ uint64_t kSecretDev[24] = MOB_XXH3_SECRET_WORDS;
#endif
static const uint64_t kSecretHost[24] = MOB_XXH3_SECRET_WORDS;
#if defined(__CUDA_ARCH__)
#define kSecretW kSecretDev
#else
#define kSecretW kSecretHost
#endif

MOB_HD uint64_t secret64(int off) {   // little-endian 64-bit read at byte offset `off` of the default secret
    const int w = off >> 3, s = (off & 7) * 8;
    return s == 0 ? kSecretW[w] : (kSecretW[w] >> s) | (kSecretW[w + 1] << (64 - s));
}
MOB_HD uint32_t secret32(int off) { return (uint32_t)(kSecretW[off >> 3] >> ((off & 7) * 8)); }   // only used at offsets 0, 4, 8, 12

MOB_HD uint64_t rd64(const uint8_t *p) {
    uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i);
    return v;
}
MOB_HD uint32_t rd32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
MOB_HD uint32_t swap32(uint32_t x) { return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24); }
MOB_HD uint64_t swap64(uint64_t x) { return ((uint64_t)swap32((uint32_t)x) << 32) | swap32((uint32_t)(x >> 32)); }
MOB_HD uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

MOB_HD Hash128 mul128(uint64_t a, uint64_t b) {
    Hash128 r;
#if defined(__CUDA_ARCH__)
    r.lo = a * b; r.hi = __umul64hi(a, b);
#else
    const unsigned __int128 m = (unsigned __int128)a * b;
    r.lo = (uint64_t)m; r.hi = (uint64_t)(m >> 64);
#endif
    return r;
}
MOB_HD uint64_t mul128_fold64(uint64_t a, uint64_t b) { const Hash128 m = mul128(a, b); return m.lo ^ m.hi; }
MOB_HD uint64_t xorshift(uint64_t v, int s) { return v ^ (v >> s); }
MOB_HD uint64_t avalanche3(uint64_t h) { h = xorshift(h, 37); h *= PMX1; return xorshift(h, 32); }
MOB_HD uint64_t avalanche64(uint64_t h) { h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; return h ^ (h >> 32); }

// ---- 4..8 bytes given as the assembled 64-bit word (lo 32 bits = first 4 bytes, hi 32 bits = last 4 bytes) --------------------------------
MOB_HD Hash128 hash_4to8(uint64_t input_64, uint64_t len, uint64_t seed) {
    seed ^= (uint64_t)swap32((uint32_t)seed) << 32;
    const uint64_t bitflip = (secret64(16) ^ secret64(24)) + seed;
    const uint64_t keyed = input_64 ^ bitflip;
    Hash128 m = mul128(keyed, P64_1 + (len << 2));
    m.hi += m.lo << 1;
    m.lo ^= m.hi >> 3;
    m.lo = xorshift(m.lo, 35);
    m.lo *= PMX2;
    m.lo = xorshift(m.lo, 28);
    m.hi = avalanche3(m.hi);
    return m;
}

// the fixed-width integer path of the bloom filter: an 8-byte key (cgo/bloom.c:31-37, after the int8/16/32 -> int64 widening of :46-55)
MOB_HD Hash128 hash_u64(uint64_t key, uint64_t seed) { return hash_4to8(key, 8, seed); }

MOB_HD uint64_t mix16(const uint8_t *in, int soff, uint64_t seed) {
    return mul128_fold64(rd64(in) ^ (secret64(soff) + seed), rd64(in + 8) ^ (secret64(soff + 8) - seed));
}
MOB_HD Hash128 mix32(Hash128 acc, const uint8_t *in1, const uint8_t *in2, int soff, uint64_t seed) {
    acc.lo += mix16(in1, soff, seed);
    acc.lo ^= rd64(in2) + rd64(in2 + 8);
    acc.hi += mix16(in2, soff + 16, seed);
    acc.hi ^= rd64(in1) + rd64(in1 + 8);
    return acc;
}
MOB_HD Hash128 finish_mid(Hash128 acc, uint64_t len, uint64_t seed) {
    Hash128 h;
    h.lo = acc.lo + acc.hi;
    h.hi = acc.lo * P64_1 + acc.hi * P64_4 + (len - seed) * P64_2;
    h.lo = avalanche3(h.lo);
    h.hi = (uint64_t)0 - avalanche3(h.hi);
    return h;
}

// one 64-byte stripe into the 8 accumulators; `sec` = byte offset into the (seed-adjusted) secret
struct LongState { uint64_t acc[8]; uint64_t seed; };
MOB_HD uint64_t csecret64(int off, uint64_t seed) {   // custom secret: word pairs (lo + seed, hi - seed) of the default secret; unaligned reads assembled
    const int w = off >> 3, s = (off & 7) * 8;
    const uint64_t a = kSecretW[w] + ((w & 1) ? (uint64_t)0 - seed : seed);
    if (s == 0) return a;
    const uint64_t b = kSecretW[w + 1] + (((w + 1) & 1) ? (uint64_t)0 - seed : seed);
    return (a >> s) | (b << (64 - s));
}
MOB_HD void accumulate_stripe(LongState &S, const uint8_t *in, int sec) {
#pragma unroll
    for (int l = 0; l < 8; l++) {
        const uint64_t v = rd64(in + 8 * l);
        const uint64_t k = v ^ csecret64(sec + 8 * l, S.seed);
        S.acc[l ^ 1] += v;
        S.acc[l] += (k & 0xffffffffull) * (k >> 32);
    }
}
MOB_HD void scramble(LongState &S) {
#pragma unroll
    for (int l = 0; l < 8; l++) {
        uint64_t a = S.acc[l];
        a = xorshift(a, 47);
        a ^= csecret64(192 - 64 + 8 * l, S.seed);
        a *= P32_1;
        S.acc[l] = a;
    }
}
MOB_HD uint64_t merge_accs(const LongState &S, int soff, uint64_t start) {
    uint64_t r = start;
#pragma unroll
    for (int i = 0; i < 4; i++) r += mul128_fold64(S.acc[2 * i] ^ csecret64(soff + 16 * i, S.seed), S.acc[2 * i + 1] ^ csecret64(soff + 16 * i + 8, S.seed));
    return avalanche3(r);
}

// XXH3_128bits_withSeed over an arbitrary byte string
MOB_HD Hash128 hash_bytes(const uint8_t *in, size_t len, uint64_t seed) {
    if (len <= 16) {
        if (len > 8) {
            const uint64_t bitflipl = (secret64(32) ^ secret64(40)) - seed, bitfliph = (secret64(48) ^ secret64(56)) + seed;
            const uint64_t input_lo = rd64(in);
            uint64_t input_hi = rd64(in + len - 8);
            Hash128 m = mul128(input_lo ^ input_hi ^ bitflipl, P64_1);
            m.lo += (uint64_t)(len - 1) << 54;
            input_hi ^= bitfliph;
            m.hi += input_hi + (uint64_t)(uint32_t)input_hi * (P32_2 - 1);
            m.lo ^= swap64(m.hi);
            Hash128 h = mul128(m.lo, P64_2);
            h.hi += m.hi * P64_2;
            h.lo = avalanche3(h.lo);
            h.hi = avalanche3(h.hi);
            return h;
        }
        if (len >= 4) return hash_4to8((uint64_t)rd32(in) + ((uint64_t)rd32(in + len - 4) << 32), (uint64_t)len, seed);
        if (len) {
            const uint8_t c1 = in[0], c2 = in[len >> 1], c3 = in[len - 1];
            const uint32_t combinedl = ((uint32_t)c1 << 16) | ((uint32_t)c2 << 24) | (uint32_t)c3 | ((uint32_t)len << 8);
            const uint32_t combinedh = rotl32(swap32(combinedl), 13);
            const uint64_t bitflipl = (uint64_t)(secret32(0) ^ secret32(4)) + seed, bitfliph = (uint64_t)(secret32(8) ^ secret32(12)) - seed;
            Hash128 h;
            h.lo = avalanche64((uint64_t)combinedl ^ bitflipl);
            h.hi = avalanche64((uint64_t)combinedh ^ bitfliph);
            return h;
        }
        Hash128 h;
        h.lo = avalanche64(seed ^ secret64(64) ^ secret64(72));
        h.hi = avalanche64(seed ^ secret64(80) ^ secret64(88));
        return h;
    }
    if (len <= 128) {
        Hash128 acc; acc.lo = (uint64_t)len * P64_1; acc.hi = 0;
        for (int i = (int)((len - 1) / 32); i >= 0; i--) acc = mix32(acc, in + 16 * i, in + len - 16 * (i + 1), 32 * i, seed);
        return finish_mid(acc, (uint64_t)len, seed);
    }
    if (len <= 240) {
        Hash128 acc; acc.lo = (uint64_t)len * P64_1; acc.hi = 0;
        for (int i = 32; i < 160; i += 32) acc = mix32(acc, in + i - 32, in + i - 16, i - 32, seed);
        acc.lo = avalanche3(acc.lo); acc.hi = avalanche3(acc.hi);
        for (int i = 160; i <= (int)len; i += 32) acc = mix32(acc, in + i - 32, in + i - 16, 3 + i - 160, seed);
        acc = mix32(acc, in + len - 16, in + len - 32, 136 - 17 - 16, (uint64_t)0 - seed);
        return finish_mid(acc, (uint64_t)len, seed);
    }
    // long inputs: 1024-byte blocks of 16 stripes (secret consumed 8 bytes per stripe), scramble after every block, last stripe at a fixed secret offset
    LongState S;
    S.acc[0] = P32_3; S.acc[1] = P64_1; S.acc[2] = P64_2; S.acc[3] = P64_3; S.acc[4] = P64_4; S.acc[5] = P32_2; S.acc[6] = P64_5; S.acc[7] = P32_1;
    S.seed = seed;
    const size_t stripes_per_block = (192 - 64) / 8, block_len = 64 * stripes_per_block;
    const size_t nblocks = (len - 1) / block_len;
    for (size_t b = 0; b < nblocks; b++) {
        for (size_t s = 0; s < stripes_per_block; s++) accumulate_stripe(S, in + b * block_len + 64 * s, (int)(8 * s));
        scramble(S);
    }
    const size_t nstripes = ((len - 1) - block_len * nblocks) / 64;
    for (size_t s = 0; s < nstripes; s++) accumulate_stripe(S, in + nblocks * block_len + 64 * s, (int)(8 * s));
    accumulate_stripe(S, in + len - 64, 192 - 64 - 7);
    Hash128 h;
    h.lo = merge_accs(S, 11, (uint64_t)len * P64_1);
    h.hi = merge_accs(S, 192 - 64 - 11, ~((uint64_t)len * P64_2));
    return h;
}

}  // namespace xxh3
}  // namespace mob
