// Synthetic power-law bipartite graphs generated ON THE DEVICE, per row shard (BASELINE.json config 5: 10 M users /
// 5 M items / 1 B interactions - 40 GB as the reference's int64 COO, which no host-side builder here could hold).
//
// Edge e of the graph is a pure function of (seed, e): the user and the item are Zipf(alpha)-distributed ranks
// (inverse CDF of the continuous power law on [1, n]) scattered over the id range by an affine bijection, like
// synth.powerlaw_bipartite does on the host for the small shapes.  A rank that owns rows [row0, row0 + n_rows) of the
// (n_user + n_item)-node adjacency scans all edges and keeps the entries (row, col) whose row it owns - both directions
// of an edge - as 64-bit keys (row - row0) << 32 | col; sorting and deduplicating the keys (plgraph.py) yields the
// shard's CSR directly: no COO, no host pass.  matrix.py:41-67 semantics: an interaction is an edge whatever its
// multiplicity, degree = number of distinct neighbours.
#include "common.cuh"

namespace {

struct PlArgs {
    int64_t n_edges;
    uint32_t n_user, n_item;
    float cu, ci, inv;               // n^(1-alpha) - 1 for users / items, 1 / (1 - alpha)
    uint64_t mul_u, add_u, mul_i, add_i;   // id = (rank * mul + add) mod n, gcd(mul, n) = 1
    uint64_t seed;
    int64_t row0, n_rows;
    unsigned long long* total;       // [1] running count / cursor
    unsigned long long* keys;        // NULL: count only
    int64_t capacity;
};

__device__ __forceinline__ uint32_t pl_rank(uint32_t h, uint32_t n, float c, float inv) {
    const float u = (float)(h >> 8) * (1.0f / 16777216.0f);            // [0, 1)
    const float x = __powf(1.0f + u * c, inv);                          // in [1, n)
    const uint32_t r = (uint32_t)x;
    return min(n - 1u, r > 0u ? r - 1u : 0u);
}

__global__ void __launch_bounds__(256) plgraph_kernel(PlArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_iter = (a.n_edges + stride - 1) / stride;
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t it = 0; it < n_iter; ++it, e += stride) {
        unsigned long long k0 = 0, k1 = 0;
        int n_mine = 0;
        if (e < a.n_edges) {
            const uint2 h = ngcf_hash64(a.seed, (uint32_t)e, (uint32_t)((uint64_t)e >> 32), 0x504cu, 0x47524146u);
            const uint64_t u = ((uint64_t)pl_rank(h.x, a.n_user, a.cu, a.inv) * a.mul_u + a.add_u) % a.n_user;
            const uint64_t i = ((uint64_t)pl_rank(h.y, a.n_item, a.ci, a.inv) * a.mul_i + a.add_i) % a.n_item + a.n_user;
            const int64_t ru = (int64_t)u - a.row0, ri = (int64_t)i - a.row0;
            if (ru >= 0 && ru < a.n_rows) { k0 = ((unsigned long long)ru << 32) | i; ++n_mine; }
            if (ri >= 0 && ri < a.n_rows) { (n_mine ? k1 : k0) = ((unsigned long long)ri << 32) | u; ++n_mine; }
        }
        // one atomic per warp: exclusive scan of the lanes' counts
        int incl = n_mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += v;
        }
        const int warp_total = __shfl_sync(FULL_MASK, incl, 31);
        unsigned long long base = 0;
        if (lane == 31 && warp_total) base = atomicAdd(a.total, (unsigned long long)warp_total);
        base = __shfl_sync(FULL_MASK, base, 31);
        if (a.keys && n_mine) {
            const long long p = (long long)base + incl - n_mine;
            if (p < a.capacity) a.keys[p] = k0;
            if (n_mine == 2 && p + 1 < a.capacity) a.keys[p + 1] = k1;
        }
    }
}

uint64_t coprime_multiplier(uint64_t n, uint64_t start) {
    auto gcd = [](uint64_t x, uint64_t y) { while (y) { const uint64_t t = x % y; x = y; y = t; } return x; };
    uint64_t m = start % n;
    if (m < 2) m = 2;
    while (gcd(m, n) != 1) ++m;
    return m;
}

}  // namespace

extern "C" int ngcf_plgraph_entries(int64_t n_user, int64_t n_item, int64_t n_edges, double alpha, uint64_t seed,
                                    int64_t row0, int64_t n_rows, unsigned long long* total_dev,
                                    unsigned long long* keys_or_null, int64_t capacity, void* stream) {
    NGCF_REQUIRE(n_user > 0 && n_item > 0 && n_user < ((int64_t)1 << 31) && n_item < ((int64_t)1 << 31) && n_edges >= 0,
                 "plgraph: shape %lld x %lld, %lld edges", (long long)n_user, (long long)n_item, (long long)n_edges);
    NGCF_REQUIRE(alpha >= 0.0 && alpha < 1.0, "plgraph: alpha %f not in [0,1)", alpha);
    NGCF_REQUIRE(total_dev && row0 >= 0 && n_rows >= 0, "plgraph: bad shard");
    PlArgs a{};
    a.n_edges = n_edges; a.n_user = (uint32_t)n_user; a.n_item = (uint32_t)n_item;
    a.inv = (float)(1.0 / (1.0 - alpha));
    a.cu = (float)(pow((double)n_user, 1.0 - alpha) - 1.0);
    a.ci = (float)(pow((double)n_item, 1.0 - alpha) - 1.0);
    a.mul_u = coprime_multiplier((uint64_t)n_user, 2654435761ull + seed);
    a.mul_i = coprime_multiplier((uint64_t)n_item, 40503ull * 65537ull + seed);
    a.add_u = (seed * 0x9E3779B97F4A7C15ull >> 17) % (uint64_t)n_user;
    a.add_i = (seed * 0xC2B2AE3D27D4EB4Full >> 19) % (uint64_t)n_item;
    a.seed = seed; a.row0 = row0; a.n_rows = n_rows;
    a.total = total_dev; a.keys = keys_or_null; a.capacity = capacity;
    if (n_edges == 0) return NGCF_OK;
    const int grid = ngcf_num_sms() * 8;
    plgraph_kernel<<<grid, 256, 0, as_stream(stream)>>>(a);
    NGCF_LAUNCH_OK("plgraph_kernel");
    return NGCF_OK;
}
