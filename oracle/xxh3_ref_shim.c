/* TEST INFRASTRUCTURE: exports the real XXH3_128bits_withSeed of the xxHash tarball the reference pins (thirdparties/xxHash-0.8.3.tar.gz),
 * compiled into oracle/_ref/libbloom_ref.so next to the reference's bloom.c, so that the device restatement can be compared hash by hash. */
#define XXH_INLINE_ALL
#include "xxhash.h"
#include <stdint.h>
#include <stddef.h>

void ref_xxh3_128(const void *p, size_t n, uint64_t seed, uint64_t *out) {
    XXH128_hash_t h = XXH3_128bits_withSeed(p, n, seed);
    out[0] = h.low64; out[1] = h.high64;
}
