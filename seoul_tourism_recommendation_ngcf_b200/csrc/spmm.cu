// CSR row-split SpMM  Y = L·X (+ addend) (+ sparse row-gradient rows).
// Replaces torch.mm(L, E) (NGCF.py:130: coalesce -> COO->CSR -> cusparseSpMM on CUDA) and the transposed
// product of its backward (MmBackward0).
//
// Mapping: one warp per row.  A warp is cut into 32/G groups of G lanes, G = next pow2 >= d/4; every lane
// of a group owns one 128-bit slice of the embedding row, so a group fetches a whole gathered row with
// one coalesced LDG.128 per lane and a warp keeps (32/G)*UNROLL gathered rows in flight.  (col,val) pairs
// are read 32 at a time, coalesced, with L1::no_allocate so the CSR stream does not evict embedding rows,
// then broadcast with shuffles.  Rows longer than the split threshold (power-law hubs) are reduced
// chunk-wise by a pre-pass (one warp per chunk) into hub_partial and only summed here, so no warp walks a
// hub row alone and the result is deterministic (no float atomics).
#include "common.cuh"

namespace {

constexpr int WARPS_PER_CTA = 8;
constexpr int UNROLL = 8;

struct SpmmArgs {
    const int32_t* rowptr;
    const int32_t* colidx;
    const float* vals;
    int64_t n_rows;
    const float* X;
    int64_t ldx;
    int d;
    const float* addend;
    int64_t ld_add;
    const int32_t* slot;
    const float* gsum;
    int64_t ld_gsum;
    const int32_t* hub_rows;
    const int32_t* hub_chunk_ptr;
    int32_t n_hub;
    const int32_t* hub_chunk_begin;
    const int32_t* hub_chunk_end;
    int32_t n_chunks;
    float* hub_partial;
    float* Y;
    int64_t ldy;
    int32_t split;     // rows with more than this many entries are hub rows (n_hub > 0 only)
    const int32_t* hub_chunk_row;   // row id of each chunk (needed by node dropout only)
    float drop_p;      // node dropout probability in device-RNG mode (0 = off)
    uint64_t seed;
    const uint64_t* seed_dev;
    int layer;
    int transposed;    // this CSR holds L^T: entry (row, col) here is entry (col, row) of L
};

// (col, val) of CSR position t of row `row`, with device-RNG node dropout folded in
__device__ __forceinline__ void load_entry(const SpmmArgs& a, int t, int row, uint64_t seed, int& c, float& v) {
    c = ld_stream_i32(a.colidx + t);
    v = ld_stream_f32(a.vals + t);
    if (a.drop_p > 0.f) {
        const uint32_t r0 = a.transposed ? (uint32_t)c : (uint32_t)row;
        const uint32_t c0 = a.transposed ? (uint32_t)row : (uint32_t)c;
        if (!node_keep(a.drop_p, seed, a.layer, r0, c0)) v = 0.f;
    }
}

// ---- vectorised path: d % 4 == 0, 16-byte aligned rows ---------------------------------------------
template <int G>
__device__ __forceinline__ float4 warp_gather_dot_vec(const SpmmArgs& a, int beg, int end, int row, int lane) {
    const uint64_t seed = a.drop_p > 0.f ? ngcf_seed(a.seed, a.seed_dev) : 0ull;
    constexpr int NG = 32 / G;
    const int g = lane / G, l = lane % G;
    const bool active = (l * 4) < a.d;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = beg; base < end; base += 32) {
        const int t = base + lane;
        int c = 0;
        float v = 0.f;
        if (t < end) load_entry(a, t, row, seed, c, v);
        const int cnt = min(32, end - base);
        for (int j0 = 0; j0 < cnt; j0 += NG * UNROLL) {
            float4 x[UNROLL];
            float w[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int jj = j0 + u * NG + g;
                const int cc = __shfl_sync(FULL_MASK, c, jj & 31);
                const float vv = __shfl_sync(FULL_MASK, v, jj & 31);
                const bool ok = (jj < cnt) && active;
                w[u] = ok ? vv : 0.f;
                x[u] = ok ? ld_f4(a.X + (int64_t)cc * a.ldx + l * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                acc.x = fmaf(w[u], x[u].x, acc.x);
                acc.y = fmaf(w[u], x[u].y, acc.y);
                acc.z = fmaf(w[u], x[u].z, acc.z);
                acc.w = fmaf(w[u], x[u].w, acc.w);
            }
        }
    }
#pragma unroll
    for (int off = G; off < 32; off <<= 1) {
        acc.x += __shfl_xor_sync(FULL_MASK, acc.x, off);
        acc.y += __shfl_xor_sync(FULL_MASK, acc.y, off);
        acc.z += __shfl_xor_sync(FULL_MASK, acc.z, off);
        acc.w += __shfl_xor_sync(FULL_MASK, acc.w, off);
    }
    return acc;
}

// sum of the hub partial rows [c0, c1) for this lane's slice
template <int G>
__device__ __forceinline__ float4 warp_sum_partials_vec(const SpmmArgs& a, int c0, int c1, int lane) {
    constexpr int NG = 32 / G;
    const int g = lane / G, l = lane % G;
    const bool active = (l * 4) < a.d;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int cbase = c0; cbase < c1; cbase += NG * 4) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int cidx = cbase + u * NG + g;
            x[u] = (cidx < c1 && active) ? ld_f4(a.hub_partial + (int64_t)cidx * a.d + l * 4)
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
    }
#pragma unroll
    for (int off = G; off < 32; off <<= 1) {
        acc.x += __shfl_xor_sync(FULL_MASK, acc.x, off);
        acc.y += __shfl_xor_sync(FULL_MASK, acc.y, off);
        acc.z += __shfl_xor_sync(FULL_MASK, acc.z, off);
        acc.w += __shfl_xor_sync(FULL_MASK, acc.w, off);
    }
    return acc;
}

__device__ __forceinline__ int find_hub(const int32_t* hub_rows, int n_hub, int row) {
    int lo = 0, hi = n_hub - 1;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (hub_rows[mid] < row) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int G>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) spmm_hub_vec_kernel(SpmmArgs a) {
    const int chunk = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (chunk >= a.n_chunks) return;
    float4 acc = warp_gather_dot_vec<G>(a, a.hub_chunk_begin[chunk], a.hub_chunk_end[chunk],
                                        a.drop_p > 0.f ? a.hub_chunk_row[chunk] : 0, lane);
    if (lane < G && lane * 4 < a.d) st_f4(a.hub_partial + (int64_t)chunk * a.d + lane * 4, acc);
}

template <int G>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32) spmm_rows_vec_kernel(SpmmArgs a) {
    const int64_t row = (int64_t)blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= a.n_rows) return;
    const int beg = a.rowptr[row], end = a.rowptr[row + 1];
    float4 acc;
    if (a.n_hub > 0 && end - beg > a.split) {
        const int h = find_hub(a.hub_rows, a.n_hub, (int)row);
        acc = warp_sum_partials_vec<G>(a, a.hub_chunk_ptr[h], a.hub_chunk_ptr[h + 1], lane);
    } else {
        acc = warp_gather_dot_vec<G>(a, beg, end, (int)row, lane);
    }
    if (lane < G && lane * 4 < a.d) {
        const int c = lane * 4;
        if (a.addend) {
            float4 ad = ld_f4(a.addend + row * a.ld_add + c);
            acc.x += ad.x; acc.y += ad.y; acc.z += ad.z; acc.w += ad.w;
        }
        if (a.slot) {
            const int s = a.slot[row];
            if (s >= 0) {
                float4 gs = ld_f4(a.gsum + (int64_t)s * a.ld_gsum + c);
                acc.x += gs.x; acc.y += gs.y; acc.z += gs.z; acc.w += gs.w;
            }
        }
        st_f4(a.Y + row * a.ldy + c, acc);
    }
}

// ---- scalar path: any d <= 128 (e.g. the reference's own width 65, 260-byte rows) ---------------------
constexpr int SC_MAXQ = NGCF_MAX_WIDTH / 32;   // columns per lane
constexpr int SC_UNROLL = 4;

__device__ __forceinline__ void warp_gather_dot_sc(const SpmmArgs& a, int beg, int end, int row, int lane,
                                                   float (&acc)[SC_MAXQ]) {
    const uint64_t seed = a.drop_p > 0.f ? ngcf_seed(a.seed, a.seed_dev) : 0ull;
#pragma unroll
    for (int q = 0; q < SC_MAXQ; ++q) acc[q] = 0.f;
    for (int base = beg; base < end; base += 32) {
        const int t = base + lane;
        int c = 0;
        float v = 0.f;
        if (t < end) load_entry(a, t, row, seed, c, v);
        const int cnt = min(32, end - base);
        for (int j0 = 0; j0 < cnt; j0 += SC_UNROLL) {
            float x[SC_UNROLL][SC_MAXQ];
            float w[SC_UNROLL];
#pragma unroll
            for (int u = 0; u < SC_UNROLL; ++u) {
                const int jj = j0 + u;
                const int cc = __shfl_sync(FULL_MASK, c, jj & 31);
                const float vv = __shfl_sync(FULL_MASK, v, jj & 31);
                const bool ok = jj < cnt;
                w[u] = ok ? vv : 0.f;
                const float* xr = a.X + (int64_t)cc * a.ldx;
#pragma unroll
                for (int q = 0; q < SC_MAXQ; ++q) {
                    const int col = lane + 32 * q;
                    x[u][q] = (ok && col < a.d) ? xr[col] : 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < SC_UNROLL; ++u)
#pragma unroll
                for (int q = 0; q < SC_MAXQ; ++q) acc[q] = fmaf(w[u], x[u][q], acc[q]);
        }
    }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) spmm_hub_sc_kernel(SpmmArgs a) {
    const int chunk = blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (chunk >= a.n_chunks) return;
    float acc[SC_MAXQ];
    warp_gather_dot_sc(a, a.hub_chunk_begin[chunk], a.hub_chunk_end[chunk],
                       a.drop_p > 0.f ? a.hub_chunk_row[chunk] : 0, lane, acc);
#pragma unroll
    for (int q = 0; q < SC_MAXQ; ++q) {
        const int col = lane + 32 * q;
        if (col < a.d) a.hub_partial[(int64_t)chunk * a.d + col] = acc[q];
    }
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32) spmm_rows_sc_kernel(SpmmArgs a) {
    const int64_t row = (int64_t)blockIdx.x * WARPS_PER_CTA + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= a.n_rows) return;
    const int beg = a.rowptr[row], end = a.rowptr[row + 1];
    float acc[SC_MAXQ];
    if (a.n_hub > 0 && end - beg > a.split) {
        const int h = find_hub(a.hub_rows, a.n_hub, (int)row);
#pragma unroll
        for (int q = 0; q < SC_MAXQ; ++q) acc[q] = 0.f;
        for (int cidx = a.hub_chunk_ptr[h]; cidx < a.hub_chunk_ptr[h + 1]; ++cidx) {
#pragma unroll
            for (int q = 0; q < SC_MAXQ; ++q) {
                const int col = lane + 32 * q;
                if (col < a.d) acc[q] += a.hub_partial[(int64_t)cidx * a.d + col];
            }
        }
    } else {
        warp_gather_dot_sc(a, beg, end, (int)row, lane, acc);
    }
    const int s = a.slot ? a.slot[row] : -1;
#pragma unroll
    for (int q = 0; q < SC_MAXQ; ++q) {
        const int col = lane + 32 * q;
        if (col < a.d) {
            float r = acc[q];
            if (a.addend) r += a.addend[row * a.ld_add + col];
            if (s >= 0) r += a.gsum[(int64_t)s * a.ld_gsum + col];
            a.Y[row * a.ldy + col] = r;
        }
    }
}

template <int G>
int launch_vec(const SpmmArgs& a, cudaStream_t st) {
    if (a.n_hub > 0 && a.n_chunks > 0) {
        spmm_hub_vec_kernel<G><<<(unsigned)ceil_div64(a.n_chunks, WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, st>>>(a);
        NGCF_LAUNCH_OK("spmm_hub_vec_kernel");
    }
    spmm_rows_vec_kernel<G><<<(unsigned)ceil_div64(a.n_rows, WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, st>>>(a);
    NGCF_LAUNCH_OK("spmm_rows_vec_kernel");
    return NGCF_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

// Split threshold shared with the host-side planner (plan.py): rows with more than NGCF_SPMM_SPLIT entries
// must appear in hub_rows when n_hub > 0.
extern "C" int ngcf_spmm_split_threshold(void) { return 256; }

extern "C" int ngcf_spmm(const int32_t* rowptr, const int32_t* colidx, const float* vals, int64_t n_rows,
                         const float* X, int64_t ldx, int d, const float* addend, int64_t ld_add,
                         const int32_t* slot, const float* gsum, int64_t ld_gsum, const int32_t* hub_rows,
                         const int32_t* hub_chunk_ptr, int32_t n_hub, const int32_t* hub_chunk_begin,
                         const int32_t* hub_chunk_end, const int32_t* hub_chunk_row, int32_t n_chunks,
                         float* hub_partial, float drop_p, uint64_t seed, const uint64_t* seed_dev, int layer,
                         int transposed, float* Y, int64_t ldy, void* stream) {
    NGCF_REQUIRE(rowptr && X && Y, "spmm: null pointer");
    NGCF_REQUIRE(n_rows >= 0 && n_rows < ((int64_t)1 << 31), "spmm: n_rows %lld", (long long)n_rows);
    NGCF_REQUIRE(d > 0 && d <= NGCF_MAX_WIDTH, "spmm: width %d not in [1,%d]", d, NGCF_MAX_WIDTH);
    NGCF_REQUIRE(ldx >= d && ldy >= d, "spmm: leading dimension smaller than width");
    NGCF_REQUIRE(n_hub >= 0 && n_chunks >= 0, "spmm: negative hub counts");
    NGCF_REQUIRE(n_hub == 0 || (hub_rows && hub_chunk_ptr && hub_chunk_begin && hub_chunk_end && hub_chunk_row &&
                                hub_partial), "spmm: hub arrays missing");
    NGCF_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "spmm: drop_p %f not in [0,1)", drop_p);
    NGCF_REQUIRE(layer >= 0 && layer < NGCF_MAX_LAYERS, "spmm: layer %d", layer);
    NGCF_REQUIRE(!slot || gsum, "spmm: slot given without gsum");
    if (n_rows == 0) return NGCF_OK;
    SpmmArgs a{rowptr, colidx, vals, n_rows, X, ldx, d, addend, ld_add, slot, gsum, ld_gsum, hub_rows, hub_chunk_ptr,
               n_hub, hub_chunk_begin, hub_chunk_end, n_chunks, hub_partial, Y, ldy, ngcf_spmm_split_threshold(),
               hub_chunk_row, drop_p, seed, seed_dev, layer, transposed};
    cudaStream_t st = as_stream(stream);
    const bool vec = (d % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(X) && aligned16(Y) &&
                     (!addend || (aligned16(addend) && ld_add % 4 == 0)) &&
                     (!slot || (aligned16(gsum) && ld_gsum % 4 == 0)) && (n_hub == 0 || aligned16(hub_partial));
    if (vec) {
        const int d4 = d / 4;
        if (d4 <= 1) return launch_vec<1>(a, st);
        if (d4 <= 2) return launch_vec<2>(a, st);
        if (d4 <= 4) return launch_vec<4>(a, st);
        if (d4 <= 8) return launch_vec<8>(a, st);
        if (d4 <= 16) return launch_vec<16>(a, st);
        return launch_vec<32>(a, st);
    }
    if (n_hub > 0 && n_chunks > 0) {
        spmm_hub_sc_kernel<<<(unsigned)ceil_div64(n_chunks, WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, st>>>(a);
        NGCF_LAUNCH_OK("spmm_hub_sc_kernel");
    }
    spmm_rows_sc_kernel<<<(unsigned)ceil_div64(n_rows, WARPS_PER_CTA), WARPS_PER_CTA * 32, 0, st>>>(a);
    NGCF_LAUNCH_OK("spmm_rows_sc_kernel");
    return NGCF_OK;
}
