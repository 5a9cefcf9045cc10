// Streaming SpMM: Y = L·X (+ addend) (+ sparse row-gradient rows), the vector path of ngcf_spmm for plain and
// per-step compacted entry lists (replaces torch.mm(L, E), NGCF.py:130, and MmBackward0).
//
// What bounded the first tiled kernel (spmm.cu) was not bandwidth but exposed latency: a warp walked its rows one
// after the other, each row = "issue <= 8 gathers per lane, wait, accumulate", so a tile of 16 short rows cost 4-6
// dependent L2/DRAM round trips per warp (tools/spmm_timeline.py: 5.5 us gather phase for ~13 entries per row), and
// the registers of a 64-register thread cap the gathers in flight.  Here:
//   * gathered embedding rows go global -> shared with cp.async (LDGSTS): no register is held while a row is in
//     flight, every lane keeps ST_SLOTS = 16 of its 16-byte pieces in flight in a private ring, and it reads back
//     exactly the bytes it asked for, so no barrier is needed between the copy and the FMAs;
//   * a group of G lanes (G = width/4) streams through a CONTIGUOUS block of tile rows as one entry stream: the
//     pipeline never drains at a row boundary (a finished row's sum is parked in a shared-memory output tile);
//   * the hub-chunk pass and the row pass are one launch (chunk tiles first: they are the longest), hub rows are
//     completed from the chunk partial sums by a small second kernel, in chunk order (deterministic, no atomics);
//   * the tile's output rows (+ addend, + row-gradient rows) are written by all threads together, coalesced.
#include "spmm_core.cuh"
#include "tc.cuh"

namespace {

using namespace ngcf;

constexpr int ST_THREADS = 128;
constexpr int ST_WARPS = ST_THREADS / 32;
constexpr int ST_BATCH = 4;                     // gathered rows per cp.async group
constexpr int ST_DEPTH = 4;                     // groups in flight per lane
constexpr int ST_SLOTS = ST_BATCH * ST_DEPTH;   // 16-byte ring slots per lane
constexpr int ST_CTAS = 5;                      // resident CTAs per SM (44 KB of shared memory each)

struct StreamSide {                 // one tile list: the ordinary rows, or the hub chunks
    const TileInfo* tiles;
    const int32_t* rowptr;          // plain: row pointers into `ent`; unused when trp is given
    const int2* ent;                // plain entries, or this layer's compacted survivors (tile t at its e0)
    const int32_t* trp;             // compacted: tile-relative row pointers [n_tiles][SP_TILE_ROWS + 1]
    float* Y;                       // output rows (chunk side: hub_partial)
    int64_t ldy;
};

struct StreamArgs {
    StreamSide rows, chunks;
    int n_chunk_tiles;              // CTAs [0, n_chunk_tiles) take chunk tiles, the rest row tiles
    const float* X;
    uint32_t ldx;
    int d;
    const float* addend;            // row side only
    int64_t ld_add;
    const int32_t* slot;
    const float* gsum;
    int64_t ld_gsum;
    unsigned long long* dbg;        // optional [n_ctas][4] phase stamps (tools/spmm_timeline.py)
};

__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void cp_async16_ca(uint32_t dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int G>
__global__ void __launch_bounds__(ST_THREADS, ST_CTAS) spmm_stream_kernel(StreamArgs a) {
    constexpr int NG = 32 / G;                                        // lane groups per warp
    constexpr int NGROUPS = ST_WARPS * NG;                            // per CTA
    constexpr int RPG = (SP_TILE_ROWS + NGROUPS - 1) / NGROUPS;       // tile rows per group (one contiguous block)
    __shared__ __align__(16) int2 ent_s[SP_TILE_ENT];
    __shared__ int rp_s[SP_TILE_ROWS + 1];
    __shared__ __align__(16) float out_s[SP_TILE_ROWS][4 * G];
    __shared__ __align__(16) float4 ring[ST_SLOTS][ST_THREADS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (a.dbg && tid == 0) a.dbg[blockIdx.x * 4 + 0] = gtime_ns();
    const bool chunk_side = (int)blockIdx.x < a.n_chunk_tiles;
    const StreamSide& sd = chunk_side ? a.chunks : a.rows;
    const int t = chunk_side ? (int)blockIdx.x : (int)blockIdx.x - a.n_chunk_tiles;
    const int4 raw = *reinterpret_cast<const int4*>(sd.tiles + t);
    const TileInfo ti{raw.x, raw.y, raw.z, raw.w};
    const int nr = ti.r1 - ti.r0;

    // ---- stage the tile: row pointers and entries (two independent coalesced reads) --------------------------------
    int cnt;
    if (sd.trp) {
        const int32_t* trp = sd.trp + (size_t)t * (SP_TILE_ROWS + 1);
        if (tid <= nr) rp_s[tid] = trp[tid];
        cnt = trp[nr];
    } else {
        if (tid <= nr) rp_s[tid] = sd.rowptr[ti.r0 + tid] - ti.e0;
        cnt = ti.e1 - ti.e0;
    }
    for (int i = tid; i < cnt; i += ST_THREADS) ent_s[i] = ld_stream_i2(sd.ent + ti.e0 + i);
    for (int i = tid; i < SP_TILE_ROWS * G; i += ST_THREADS)
        reinterpret_cast<float4*>(&out_s[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (a.dbg && tid == 0) a.dbg[blockIdx.x * 4 + 1] = gtime_ns();

    // ---- this group's entry stream: tile rows [q RPG, q RPG + RPG) = entries [s0, s1) ------------------------------
    const int q = warp * NG + lane / G, l = lane % G;
    int bnd[RPG];                                                     // end of each of the group's rows
#pragma unroll
    for (int k = 0; k < RPG; ++k) bnd[k] = rp_s[min(q * RPG + k + 1, nr)];
    const int s0 = rp_s[min(q * RPG, nr)], s1 = bnd[RPG - 1];
    const int my_steps = (s1 - s0 + ST_BATCH - 1) / ST_BATCH;
    const int steps = __reduce_max_sync(FULL_MASK, my_steps);         // the warp's groups run in lock step
    const char* xl = reinterpret_cast<const char*>(a.X + ((l * 4) < a.d ? l * 4 : 0));   // lanes past the width re-read column 0
    const uint32_t row_bytes = a.ldx * 4u;
    const uint32_t ring0 = tc::smem_u32(&ring[0][tid]);
    constexpr uint32_t SLOT_STRIDE = ST_THREADS * 16;

    auto next_boundary = [&](int p) {                                 // first row end > p
        int nb = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < RPG; ++k) nb = (bnd[k] > p && bnd[k] < nb) ? bnd[k] : nb;
        return nb;
    };
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int nb = next_boundary(s0);
    const int total = steps > 0 ? steps + ST_DEPTH - 1 : 0;
    for (int step = 0; step < total; ++step) {
        // issue: the gathered rows of entries [s0 + 4 step, +4) -> ring slots (step % DEPTH) * 4 ..
        {
            const int p0 = s0 + step * ST_BATCH;
            const uint32_t dst = ring0 + (uint32_t)((step % ST_DEPTH) * ST_BATCH) * SLOT_STRIDE;
#pragma unroll
            for (int u = 0; u < ST_BATCH; ++u) {
                if (p0 + u < s1) {
                    const int col = ent_s[p0 + u].x;
                    cp_async16_ca(dst + u * SLOT_STRIDE, xl + (uint64_t)(uint32_t)col * row_bytes);
                }
            }
            cp_async_commit();
        }
        cp_async_wait<ST_DEPTH - 1>();
        // consume: the group issued DEPTH - 1 steps ago
        const int cs = step - (ST_DEPTH - 1);
        const int p0 = s0 + cs * ST_BATCH;
        if (cs >= 0 && p0 < s1) {
            const float4* src = &ring[(cs % ST_DEPTH) * ST_BATCH][tid];
            if (nb > p0 + ST_BATCH) {                                 // no row ends inside this batch
#pragma unroll
                for (int u = 0; u < ST_BATCH; ++u) {
                    const float w = __int_as_float(ent_s[p0 + u].y);
                    const float4 x = src[u * ST_THREADS];
                    acc.x = fmaf(w, x.x, acc.x); acc.y = fmaf(w, x.y, acc.y);
                    acc.z = fmaf(w, x.z, acc.z); acc.w = fmaf(w, x.w, acc.w);
                }
            } else {
#pragma unroll
                for (int u = 0; u < ST_BATCH; ++u) {
                    const int p = p0 + u;
                    if (p < s1) {
                        const float w = __int_as_float(ent_s[p].y);
                        const float4 x = src[u * ST_THREADS];
                        acc.x = fmaf(w, x.x, acc.x); acc.y = fmaf(w, x.y, acc.y);
                        acc.z = fmaf(w, x.z, acc.z); acc.w = fmaf(w, x.w, acc.w);
                        if (p + 1 == nb) {                            // entry p closes its row: park the sum
                            int k = 0;
#pragma unroll
                            for (int j = 0; j < RPG; ++j) k += (bnd[j] <= p) ? 1 : 0;
                            *reinterpret_cast<float4*>(&out_s[q * RPG + k][4 * l]) = acc;
                            acc = make_float4(0.f, 0.f, 0.f, 0.f);
                            nb = next_boundary(p + 1);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();

    // ---- write the tile's rows: coalesced, + addend, + row-gradient rows (row side) -------------------------------
    const int d4 = a.d >> 2;
    for (int i = tid; i < nr * G; i += ST_THREADS) {
        const int r = i / G, c4 = i % G;
        if (c4 >= d4) continue;
        float4 v = *reinterpret_cast<const float4*>(&out_s[r][4 * c4]);
        const int64_t row = ti.r0 + r;
        if (!chunk_side) {
            if (a.addend) {
                const float4 ad = ld_f4(a.addend + row * a.ld_add + 4 * c4);
                v.x += ad.x; v.y += ad.y; v.z += ad.z; v.w += ad.w;
            }
            if (a.slot) {
                const int s = a.slot[row];
                if (s >= 0) {
                    const float4 gs = ld_f4(a.gsum + (int64_t)s * a.ld_gsum + 4 * c4);
                    v.x += gs.x; v.y += gs.y; v.z += gs.z; v.w += gs.w;
                }
            }
        }
        st_f4(sd.Y + row * sd.ldy + 4 * c4, v);
    }
    if (a.dbg && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        a.dbg[blockIdx.x * 4 + 2] = gtime_ns();
        a.dbg[blockIdx.x * 4 + 3] = smid;
    }
}

// hub rows: Y[row] = sum of the row's chunk partial sums, in chunk order (+ addend) (+ row-gradient row)
struct HubFinishArgs {
    const int32_t* hub_rows;
    const int32_t* hub_chunk_ptr;
    const float* partial;
    int n_hub, d;
    const float* addend;
    int64_t ld_add;
    const int32_t* slot;
    const float* gsum;
    int64_t ld_gsum;
    float* Y;
    int64_t ldy;
};

template <int G>
__global__ void __launch_bounds__(256) hub_finish_kernel(HubFinishArgs a) {
    const int h = (blockIdx.x * 256 + threadIdx.x) / G, l = threadIdx.x % G;
    if (h >= a.n_hub || l * 4 >= a.d) return;
    const int c0 = a.hub_chunk_ptr[h], c1 = a.hub_chunk_ptr[h + 1];
    const int64_t row = a.hub_rows[h];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int c = c0;
    for (; c + 4 <= c1; c += 4) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = ld_f4(a.partial + (int64_t)(c + u) * a.d + l * 4);
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
    }
    for (; c < c1; ++c) {
        const float4 x = ld_f4(a.partial + (int64_t)c * a.d + l * 4);
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    if (a.addend) {
        const float4 ad = ld_f4(a.addend + row * a.ld_add + l * 4);
        acc.x += ad.x; acc.y += ad.y; acc.z += ad.z; acc.w += ad.w;
    }
    if (a.slot) {
        const int s = a.slot[row];
        if (s >= 0) {
            const float4 gs = ld_f4(a.gsum + (int64_t)s * a.ld_gsum + l * 4);
            acc.x += gs.x; acc.y += gs.y; acc.z += gs.z; acc.w += gs.w;
        }
    }
    st_f4(a.Y + row * a.ldy + l * 4, acc);
}

unsigned long long* g_stream_dbg = nullptr;   // host copy of the debug buffer pointer

template <int G>
int launch_stream(const StreamArgs& a, int n_ctas, const HubFinishArgs& hf, cudaStream_t st) {
    spmm_stream_kernel<G><<<(unsigned)n_ctas, ST_THREADS, 0, st>>>(a);
    NGCF_LAUNCH_OK("spmm_stream_kernel");
    if (hf.n_hub > 0) {
        hub_finish_kernel<G><<<(unsigned)ceil_div64((int64_t)hf.n_hub * G, 256), 256, 0, st>>>(hf);
        NGCF_LAUNCH_OK("hub_finish_kernel");
    }
    return NGCF_OK;
}

}  // namespace

// vector path (d % 4 == 0, aligned) of ngcf_spmm for plain / compacted entry lists; see spmm.cu for the argument checks
int ngcf_spmm_stream(const ngcf_csr* g, const float* X, int64_t ldx, int d, const float* addend, int64_t ld_add,
                     const int32_t* slot, const float* gsum, int64_t ld_gsum, float* hub_partial,
                     const int32_t* c_ent, const int32_t* c_trp, float* Y, int64_t ldy, cudaStream_t st) {
    const bool hubs = g->n_hub > 0 && g->n_chunks > 0;
    StreamArgs a{};
    a.rows.tiles = reinterpret_cast<const TileInfo*>(g->tiles);
    a.rows.rowptr = g->rowptr;
    a.rows.ent = reinterpret_cast<const int2*>(c_ent ? c_ent : g->ent);
    a.rows.trp = c_trp;
    a.rows.Y = Y;
    a.rows.ldy = ldy;
    a.n_chunk_tiles = hubs ? g->n_chunk_tiles : 0;
    if (hubs) {
        a.chunks.tiles = reinterpret_cast<const TileInfo*>(g->chunk_tiles);
        a.chunks.rowptr = g->chunk_ptr;
        a.chunks.ent = c_ent ? reinterpret_cast<const int2*>(c_ent) + g->rowptr_nnz : reinterpret_cast<const int2*>(g->hub_ent);
        a.chunks.trp = c_trp ? c_trp + (size_t)g->n_tiles * (SP_TILE_ROWS + 1) : nullptr;
        a.chunks.Y = hub_partial;
        a.chunks.ldy = d;
    }
    a.X = X; a.ldx = (uint32_t)ldx; a.d = d;
    a.addend = addend; a.ld_add = ld_add; a.slot = slot; a.gsum = gsum; a.ld_gsum = ld_gsum;
    a.dbg = g_stream_dbg;
    HubFinishArgs hf{g->hub_rows, g->hub_chunk_ptr, hub_partial, hubs ? g->n_hub : 0, d, addend, ld_add, slot, gsum, ld_gsum, Y, ldy};
    const int n_ctas = a.n_chunk_tiles + g->n_tiles;
    if (n_ctas <= 0) return NGCF_OK;
    const int d4 = d / 4;
    if (d4 <= 1) return launch_stream<1>(a, n_ctas, hf, st);
    if (d4 <= 2) return launch_stream<2>(a, n_ctas, hf, st);
    if (d4 <= 4) return launch_stream<4>(a, n_ctas, hf, st);
    if (d4 <= 8) return launch_stream<8>(a, n_ctas, hf, st);
    if (d4 <= 16) return launch_stream<16>(a, n_ctas, hf, st);
    return launch_stream<32>(a, n_ctas, hf, st);
}

// debugging aid (tools/spmm_timeline.py): device buffer [n_ctas][4] of uint64 that every streaming-SpMM CTA stamps
extern "C" int ngcf_debug_spmm_timeline(unsigned long long* dev_buf_or_null) {
    g_stream_dbg = dev_buf_or_null;
    return NGCF_OK;
}
