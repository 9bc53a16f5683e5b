/* mo_b200_bloom.h -- the reference's C bloom filter API (cgo/bloom.h:40-159), exported by libmo_b200.so with the same names, the same
 * struct and the same bits; what pkg/common/bloomfilter/cbloomfilter.go binds with cgo.  The filter's bitmap is worked on in a device
 * mirror (uploaded once, on first use); key / nullmap / result pointers follow the library's pointer rule (host or device).
 *
 * Bit positions: (h1 + i * h2) & (nbits - 1), i < k, (h1, h2) = XXH3_128bits_withSeed(key, len, seed) of xxHash 0.8.3
 * (cgo/bloom.c:31-74; 1/2/4-byte keys are sign-extended to int64 first, :46-55).  NULL rows (bit set in nullmap): test -> false,
 * add -> skipped.  test_and_add keeps the reference's row-by-row semantics (row i sees the bits of rows < i).
 *
 * These entry points have no error channel in the reference (void / bool).  If the GPU runtime is not available they print the
 * reason to stderr and abort() -- they never compute on the CPU and never return a made-up answer. */
#ifndef MO_B200_BLOOM_H
#define MO_B200_BLOOM_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLOOM_MAGIC "XXBF"
#define MAX_K_SEED 64

typedef struct {        /* cgo/bloom.h:35-41 */
    uint8_t magic[4];
    uint32_t k;
    uint64_t nbits;     /* a power of two */
    uint64_t seed;
    uint64_t bitmap[1]; /* nbits / 64 words follow */
} bloomfilter_t;

bloomfilter_t *bloomfilter_init(uint64_t nbits, uint32_t k);                                  /* bloom.h:48, bloom.c:98-116 */
bloomfilter_t *bloomfilter_init_with_seed(uint64_t nbits, uint32_t k, uint64_t seed);        /* bloom.h:57, bloom.c:118-130 */
void bloomfilter_free(bloomfilter_t *bf);                                                     /* bloom.h:62: also drops the device mirror */
void bloomfilter_add(bloomfilter_t *bf, const void *key, size_t len);                         /* bloom.h:67 */
void bloomfilter_add_fixed(bloomfilter_t *bf, const void *key, size_t len, size_t elemsz, size_t nitem,
                           const void *nullmap, size_t nullmaplen);                           /* bloom.h:78 */
bool bloomfilter_test(const bloomfilter_t *bf, const void *key, size_t len);                  /* bloom.h:84 */
void bloomfilter_test_fixed(const bloomfilter_t *bf, const void *key, size_t len, size_t elemsz, size_t nitem,
                            const void *nullmap, size_t nullmaplen, void *result);            /* bloom.h:89: result = bool[nitem] */
void bloomfilter_test_varlena_4b(const bloomfilter_t *bf, const void *key, size_t len, size_t nitem,
                                 const void *nullmap, size_t nullmaplen, void *result);       /* bloom.h:95: [u32 len][bytes]... */
void bloomfilter_add_varlena_4b(bloomfilter_t *bf, const void *key, size_t len, size_t nitem,
                                const void *nullmap, size_t nullmaplen);                      /* bloom.h:101 */
bool bloomfilter_test_and_add(bloomfilter_t *bf, const void *key, size_t len);                /* bloom.h:107 */
void bloomfilter_test_and_add_fixed(bloomfilter_t *bf, const void *key, size_t len, size_t elemsz, size_t nitem,
                                    const void *nullmap, size_t nullmaplen, void *result);    /* bloom.h:112 */
void bloomfilter_test_and_add_varlena_4b(bloomfilter_t *bf, const void *key, size_t len, size_t nitem,
                                         const void *nullmap, size_t nullmaplen, void *result); /* bloom.h:117 */
uint8_t *bloomfilter_marshal(const bloomfilter_t *bf, size_t *len);                           /* bloom.h:123: refreshes the host bytes from the mirror */
bloomfilter_t *bloomfilter_unmarshal(const uint8_t *buf, size_t len);                         /* bloom.h:129: adopts buf, no copy */
void bloomfilter_add_varlena(bloomfilter_t *bf, const void *keys, size_t len, size_t elemsz, size_t nitem,
                             const void *area, size_t area_len, const void *nullmap, size_t nullmaplen);   /* bloom.h:136: 24-byte varlena cells + area */
void bloomfilter_test_varlena(const bloomfilter_t *bf, const void *keys, size_t len, size_t elemsz, size_t nitem,
                              const void *area, size_t area_len, const void *nullmap, size_t nullmaplen, void *result);   /* bloom.h:143 */
void bloomfilter_test_and_add_varlena(bloomfilter_t *bf, const void *keys, size_t len, size_t elemsz, size_t nitem,
                                      const void *area, size_t area_len, const void *nullmap, size_t nullmaplen, void *result);   /* bloom.h:150 */
int bloomfilter_or(bloomfilter_t *dst, const bloomfilter_t *a, const bloomfilter_t *b);       /* bloom.h:155: 0, or 1 / 2 / 3 = nbits / seed / k differ */

static inline uint64_t bloomfilter_get_nbits(const bloomfilter_t *bf) { return bf->nbits; }  /* bloom.h:160-176 */
static inline uint64_t bloomfilter_get_seed(const bloomfilter_t *bf) { return bf->seed; }
static inline uint32_t bloomfilter_get_k(const bloomfilter_t *bf) { return bf->k; }

#ifdef __cplusplus
}
#endif
#endif
