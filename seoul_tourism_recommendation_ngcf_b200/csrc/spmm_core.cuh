// Device-side core of the tiled CSR SpMM (spmm.cu: L·E of the forward, L^T·gS of the backward).
//
// Layout (built once per Laplacian by plan.py, see ngcf_csr in ngcf_b200.h):
//   * entries are interleaved (col, value-bits) pairs, so one 8-byte shared-memory read yields both;
//   * rows longer than the split threshold ("hubs" of the power-law graph) are removed from the row CSR and
//     stored as fixed-size chunks = pseudo-rows of a second CSR whose results (hub_partial) are summed, in
//     order, by the owning row: no warp walks a hub alone and no float atomics are needed (deterministic);
//   * rows are grouped into tiles (<= TILE_ROWS rows, <= TILE_ENT entries).  A CTA stages a tile's row
//     pointers and entries in shared memory with two coalesced reads, which removes the
//     rowptr -> entries -> gather dependency chain from every row; node dropout (device RNG) is applied to the
//     staged values; then each warp gathers whole rows with 128-bit loads, UNROLL gathers in flight per lane.
#pragma once
#include "common.cuh"

namespace ngcf {

constexpr int SPLIT = 128;          // rows with more entries than this are hubs; also the hub chunk size
#ifndef NGCF_SPMM_TILE_ROWS
#define NGCF_SPMM_TILE_ROWS 16
#define NGCF_SPMM_TILE_ENT 256
#endif
constexpr int SP_TILE_ROWS = NGCF_SPMM_TILE_ROWS;    // SpMM tile: at most this many rows ...
constexpr int SP_TILE_ENT = NGCF_SPMM_TILE_ENT;   // ... and this many entries (staged in shared memory by one CTA)
static_assert(SP_TILE_ENT >= SPLIT, "a tile must hold the longest ordinary row");
#ifndef NGCF_SPMM_TAIL
#define NGCF_SPMM_TAIL 4       // predicated remainder batch per lane group (2, 4 and 8 measure the same within noise)
#endif
#ifndef NGCF_SPMM_UNROLL
#define NGCF_SPMM_UNROLL 4
#endif
constexpr int UNROLL = NGCF_SPMM_UNROLL;           // gathered rows in flight per lane group

// An entry's 32-bit column word also carries, in its top 5 bits, the index of the entry's row inside its tile (packed
// once by plan.py; the per-step compaction copies entries verbatim, so it survives node dropout): the streaming
// kernel needs no row pointers to know where a row ends.  Column ids therefore have 27 bits (N < 134 M nodes).
constexpr int LR_SHIFT = 27;
constexpr uint32_t COL_MASK = (1u << LR_SHIFT) - 1u;
static_assert(NGCF_SPMM_TILE_ROWS <= 32, "the local row index of an entry has 5 bits");
__device__ __forceinline__ uint32_t ent_col(int x) { return (uint32_t)x & COL_MASK; }
__device__ __forceinline__ int ent_lrow(int x) { return (int)((uint32_t)x >> LR_SHIFT); }

struct TileInfo {                   // int4: rows [r0, r1), entries [e0, e1) of one tile
    int r0, r1, e0, e1;
};

__device__ __forceinline__ int2 ld_stream_i2(const int2* p) {
    int2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}

struct DropArgs {
    float p;            // 0 = off
    uint64_t seed;      // already combined with the device counter
    int layer;
    int transposed;     // the CSR holds L^T: entry (row, col) here is entry (col, row) of L
    uint32_t row_off;   // global index of local row 0 (row-sharded runs key the RNG on global coordinates)
    const uint8_t* bits;    // optional precomputed decisions (ngcf_node_dropout_bits), indexed like `ent`; bit = layer
};

// Stages one tile: rp_s[0 .. nr] = rowptr[r0 .. r1] - e0 (tile-relative), ent_s[0 .. cnt) = entries with node
// dropout folded in.  row_key: optional per-row id used as the dropout key instead of the row index (hub chunks).
// All `nthreads` threads of the calling group take part; `sync()` is the group's barrier.
template <int NTHREADS, typename Sync>
__device__ __forceinline__ void stage_tile(const TileInfo ti, const int32_t* __restrict__ rowptr,
                                           const int2* __restrict__ ent, const int32_t* __restrict__ row_key,
                                           const DropArgs& dr, int* rp_s, int2* ent_s, int tid, Sync sync) {
    const int nr = ti.r1 - ti.r0, cnt = ti.e1 - ti.e0;
    for (int i = tid; i <= nr; i += NTHREADS) rp_s[i] = rowptr[ti.r0 + i] - ti.e0;
    if (dr.bits) {                                                // decisions drawn once per step: one byte per entry
        for (int i = tid; i < cnt; i += NTHREADS) {
            int2 e = ld_stream_i2(ent + ti.e0 + i);
            if (!((dr.bits[ti.e0 + i] >> dr.layer) & 1)) e.y = 0;
            ent_s[i] = e;
        }
    } else if (dr.p > 0.f) {
        sync();                                                   // rp_s visible: entries need their row
        for (int i = tid; i < cnt; i += NTHREADS) {
            int2 e = ld_stream_i2(ent + ti.e0 + i);
            int lo = 0, hi = nr - 1;                              // last row with rp_s[row] <= i
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (rp_s[mid] <= i) lo = mid; else hi = mid - 1;
            }
            const uint32_t r = (uint32_t)(row_key ? row_key[ti.r0 + lo] : ti.r0 + lo) + dr.row_off;
            const uint32_t r0 = dr.transposed ? ent_col(e.x) : r;
            const uint32_t c0 = dr.transposed ? r : ent_col(e.x);
            if (!node_keep(dr.p, dr.seed, dr.layer, r0, c0)) e.y = 0;
            ent_s[i] = e;
        }
    } else {
        for (int i = tid; i < cnt; i += NTHREADS) ent_s[i] = ld_stream_i2(ent + ti.e0 + i);
    }
    sync();
}

// ---- vector path: d % 4 == 0, 16-byte aligned rows.  A warp is cut into 32/G groups of G lanes; lane l of a
// group owns columns [4l, 4l+4) and a group fetches one gathered row per LDG.128. ------------------------------
// The issue rate, not memory, bounds this loop once the entries sit in shared memory (ncu: 67 % issue-active with a
// fully predicated body), so whole batches of NG*UNROLL entries run without any predicate: one LDS.64 with an
// immediate offset, one IMAD.WIDE, one LDG.128 and four FFMA per gathered row; only the remainder is predicated.
template <int G>
__device__ __forceinline__ float4 gather_row_vec(const int2* ent_s, int a, int b, const float* __restrict__ X,
                                                 uint32_t ldx, int d, int lane) {
    constexpr int NG = 32 / G;
    constexpr int TAIL = NGCF_SPMM_TAIL;
    const int g = lane / G, l = lane % G;
    // byte addressing: one IMAD.WIDE.U32 (col * row_bytes + base) per gathered row
    const char* xl = reinterpret_cast<const char*>(X + ((l * 4) < d ? l * 4 : 0));   // lanes past the width re-read column 0
    const uint32_t row_bytes = ldx * 4u;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int j0 = a;
    for (; j0 + NG * UNROLL <= b; j0 += NG * UNROLL) {
        const int2* p = ent_s + j0 + g;
        int2 e[UNROLL];
        float4 x[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) e[u] = p[u * NG];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            x[u] = ld_f4(reinterpret_cast<const float*>(xl + (uint64_t)ent_col(e[u].x) * row_bytes));
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const float w = __int_as_float(e[u].y);
            acc.x = fmaf(w, x[u].x, acc.x);
            acc.y = fmaf(w, x[u].y, acc.y);
            acc.z = fmaf(w, x[u].z, acc.z);
            acc.w = fmaf(w, x[u].w, acc.w);
        }
    }
    for (; j0 < b; j0 += NG * TAIL) {
        float4 x[TAIL];
        float w[TAIL];
#pragma unroll
        for (int u = 0; u < TAIL; ++u) {
            const int idx = j0 + u * NG + g;
            const bool ok = idx < b;
            int2 e = make_int2(0, 0);
            if (ok) e = ent_s[idx];
            w[u] = __int_as_float(e.y);
            x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) x[u] = ld_f4(reinterpret_cast<const float*>(xl + (uint64_t)ent_col(e.x) * row_bytes));
        }
#pragma unroll
        for (int u = 0; u < TAIL; ++u) {
            acc.x = fmaf(w[u], x[u].x, acc.x);
            acc.y = fmaf(w[u], x[u].y, acc.y);
            acc.z = fmaf(w[u], x[u].z, acc.z);
            acc.w = fmaf(w[u], x[u].w, acc.w);
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

// in-order sum of the hub partial rows [c0, c1) for this lane's slice (every group computes the same sum)
template <int G>
__device__ __forceinline__ float4 sum_partials_vec(const float* __restrict__ partial, int c0, int c1, int d,
                                                   int lane) {
    const int l = lane % G;
    const bool active = (l * 4) < d;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (!active) return acc;
    int c = c0;
    for (; c + 4 <= c1; c += 4) {
        float4 x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) x[u] = ld_f4(partial + (int64_t)(c + u) * d + l * 4);
#pragma unroll
        for (int u = 0; u < 4; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
    }
    for (; c < c1; ++c) {
        const float4 x = ld_f4(partial + (int64_t)c * d + l * 4);
        acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
    }
    return acc;
}

// the same sum with the warp's lane groups taking the partial rows round robin, UNROLL of them in flight per lane
// (the widest hub of a power-law graph has ~100 chunks), combined in group order: a fixed tree, deterministic.
// Partial rows were written by other CTAs during this launch: read them past L1 (ld.global.cg).
template <int G>
__device__ __forceinline__ float4 sum_partials_split(const float* partial, int c0, int c1, int d, int lane) {
    constexpr int NG = 32 / G;
    const int g = lane / G, l = lane % G;
    const bool active = (l * 4) < d;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (active) {
        int c = c0 + g;
        for (; c + (UNROLL - 1) * NG < c1; c += UNROLL * NG) {
            float4 x[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                x[u] = __ldcg(reinterpret_cast<const float4*>(partial + (int64_t)(c + u * NG) * d + l * 4));
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
        }
        for (; c < c1; c += NG) {
            const float4 x = __ldcg(reinterpret_cast<const float4*>(partial + (int64_t)c * d + l * 4));
            acc.x += x.x; acc.y += x.y; acc.z += x.z; acc.w += x.w;
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

// ---- scalar path: any d <= 128 (the reference's own width is 65: 260-byte rows).  Lane owns columns
// lane + 32q; the whole warp fetches one gathered row per step. -------------------------------------------------
constexpr int SC_MAXQ = NGCF_MAX_WIDTH / 32;
constexpr int SC_UNROLL = 4;

__device__ __forceinline__ void gather_row_sc(const int2* ent_s, int a, int b, const float* __restrict__ X,
                                              uint32_t ldx, int d, int lane, float (&acc)[SC_MAXQ]) {
#pragma unroll
    for (int q = 0; q < SC_MAXQ; ++q) acc[q] = 0.f;
    for (int j0 = a; j0 < b; j0 += SC_UNROLL) {
        float x[SC_UNROLL][SC_MAXQ];
        float w[SC_UNROLL];
#pragma unroll
        for (int u = 0; u < SC_UNROLL; ++u) {
            const int idx = j0 + u;
            const bool ok = idx < b;
            int2 e = make_int2(0, 0);
            if (ok) e = ent_s[idx];
            w[u] = __int_as_float(e.y);
            const float* xr = X + (uint64_t)ent_col(e.x) * ldx;
#pragma unroll
            for (int q = 0; q < SC_MAXQ; ++q) {
                const int col = lane + 32 * q;
                x[u][q] = (ok && col < d) ? xr[col] : 0.f;
            }
        }
#pragma unroll
        for (int u = 0; u < SC_UNROLL; ++u)
#pragma unroll
            for (int q = 0; q < SC_MAXQ; ++q) acc[q] = fmaf(w[u], x[u][q], acc[q]);
    }
}

__device__ __forceinline__ void sum_partials_sc(const float* __restrict__ partial, int c0, int c1, int d, int lane,
                                                float (&acc)[SC_MAXQ]) {
#pragma unroll
    for (int q = 0; q < SC_MAXQ; ++q) acc[q] = 0.f;
    for (int c = c0; c < c1; ++c) {
#pragma unroll
        for (int q = 0; q < SC_MAXQ; ++q) {
            const int col = lane + 32 * q;
            if (col < d) acc[q] += __ldcg(partial + (int64_t)c * d + col);
        }
    }
}

}  // namespace ngcf
