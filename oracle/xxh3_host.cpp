// TEST INFRASTRUCTURE: the device hash header (matrixone_b200/csrc/xxh3_128.cuh) compiled for the host, so that tests/test_xxh3_host.py can
// pin it to the real xxHash (oracle/_ref/libbloom_ref.so) without a GPU.  Not part of the product.
#include "xxh3_128.cuh"

extern "C" void mob_xxh3_128_bytes(const uint8_t *in, size_t len, uint64_t seed, uint64_t *out) {
    const mob::xxh3::Hash128 h = mob::xxh3::hash_bytes(in, len, seed);
    out[0] = h.lo; out[1] = h.hi;
}
extern "C" void mob_xxh3_128_u64(uint64_t key, uint64_t seed, uint64_t *out) {
    const mob::xxh3::Hash128 h = mob::xxh3::hash_u64(key, seed);
    out[0] = h.lo; out[1] = h.hi;
}
