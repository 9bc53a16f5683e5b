"""Host-side mirror of the random source the reference's k-means uses: math/rand/v2's PCG (PCG-DXSM, 128-bit state) as
`rand.New(rand.NewPCG(uint64(kmeans.DefaultRandSeed), 0))` (pkg/vectorindex/ivfflat/kmeans/elkans/initializer.go:46, clusterer.go:362,
kmeans/types.go:19 DefaultRandSeed = 1), with IntN (Lemire's multiply-shift with rejection) and Float32.  Go's standard library is not part of
/root/reference; this restates its published algorithm and is pinned by the reference's own expectations: the centroids
TestRandom_InitCentroids (initializer_test.go:36-59) and Test_Cluster (clusterer_test.go:441-470) expect are only reachable with the draws
IntN(12) = 7, 1 (tests/test_oracle_kmeans.py).  Pure host logic: it picks which vectors seed the centroids, nothing is computed here."""

_M64 = (1 << 64) - 1
_MUL = (2549297995355413924 << 64) | 4865540595714422341      # the 128-bit PCG multiplier
_INC = (6364136223846793005 << 64) | 1442695040888963407
_CHEAP = 0xDA942042E4DD58B5                                    # DXSM output multiplier


class PCG:
    def __init__(self, seed1=1, seed2=0):
        self.state = ((seed1 & _M64) << 64) | (seed2 & _M64)

    def uint64(self):
        self.state = (self.state * _MUL + _INC) & ((1 << 128) - 1)
        hi, lo = self.state >> 64, self.state & _M64
        hi ^= hi >> 32
        hi = (hi * _CHEAP) & _M64
        hi ^= hi >> 48
        return (hi * (lo | 1)) & _M64

    def intn(self, n):
        """rand.IntN on a 64-bit platform (uint64n)"""
        if n & (n - 1) == 0:
            return self.uint64() & (n - 1)
        x = self.uint64() * n
        hi, lo = x >> 64, x & _M64
        if lo < n:
            thresh = ((1 << 64) - n) % n
            while lo < thresh:
                x = self.uint64() * n
                hi, lo = x >> 64, x & _M64
        return hi

    def float32(self):
        """rand.Float32: float32(Uint32() << 8 >> 8) / (1 << 24), Uint32 = the top half of Uint64"""
        u32 = self.uint64() >> 32
        return float(((u32 << 8) & 0xFFFFFFFF) >> 8) / float(1 << 24)


def random_init_rows(n, k, seed=1):
    """Random.InitCentroids (initializer.go:45-71): k draws of IntN(n) -- repeats allowed"""
    r = PCG(seed, 0)
    return [r.intn(n) for _ in range(k)]


def empty_cluster_stream(count, seed=1):
    """the rnd.Float32() values elkansCluster's generator would hand to empty clusters (clusterer.go:362,700-707); a fresh generator per Cluster() call"""
    r = PCG(seed, 0)
    return [r.float32() for _ in range(count)]
