// Output-row gather with on-the-fly L2 normalisation (NGCF.py:144-156), the BPR loss with its row
// gradients in one kernel (bprloss.py:15-22), and the backward of the row gather as a slot map
// (IndexBackward, NGCF.py:151-155) instead of a dense N x D zero-fill + scatter.
#include "common.cuh"

namespace {

struct GatherArgs {
    const float* layer[NGCF_MAX_LAYERS + 1];
    int dim[NGCF_MAX_LAYERS + 1];
    int n;               // number of blocks (K+1)
    const int64_t* rows;
    int64_t row_offset;
    int64_t n_out;
    float* out;
    int64_t ld_out;
};

// up to four row sets (users, positives, negatives of a batch) gathered by ONE launch
struct GatherSetsArgs {
    const float* layer[NGCF_MAX_LAYERS + 1];
    int dim[NGCF_MAX_LAYERS + 1];
    int n;
    const int64_t* rows[4];
    int64_t row_offset[4];
    int64_t base[5];     // first global output index of set j; base[n_sets] = total
    float* out[4];
    int n_sets;
    int64_t ld_out;
};

template <typename A>
__device__ __forceinline__ void gather_row(const A& a, int64_t r, float* o, int lane) {
    int off = 0;
    for (int k = 0; k < a.n; ++k) {
        const int d = a.dim[k];
        const float* src = a.layer[k] + r * (int64_t)d;
        float v[NGCF_MAX_WIDTH / 32];
        float ss = 0.f;
#pragma unroll
        for (int q = 0; q < NGCF_MAX_WIDTH / 32; ++q) {
            const int c = lane + 32 * q;
            v[q] = c < d ? src[c] : 0.f;
            ss = fmaf(v[q], v[q], ss);
        }
        float scale = 1.f;
        if (k > 0) scale = 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);   // F.normalize(p=2, eps=1e-12), NGCF.py:144
#pragma unroll
        for (int q = 0; q < NGCF_MAX_WIDTH / 32; ++q) {
            const int c = lane + 32 * q;
            if (c < d) o[off + c] = k > 0 ? v[q] * scale : v[q];
        }
        off += d;
    }
}

// one warp per output row
__global__ void gather_concat_kernel(GatherArgs a) {
    const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= a.n_out) return;
    gather_row(a, (a.rows ? a.rows[b] : b) + a.row_offset, a.out + b * a.ld_out, lane);
}
__global__ void gather_concat_sets_kernel(GatherSetsArgs a) {
    const int64_t b = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= a.base[a.n_sets]) return;
    int j = 0;
    while (b >= a.base[j + 1]) ++j;
    const int64_t i = b - a.base[j];
    gather_row(a, a.rows[j][i] + a.row_offset[j], a.out[j] + i * a.ld_out, lane);
}

// ---- BPR ----------------------------------------------------------------------------------------------
constexpr int BPR_WARPS = 8;

__global__ void __launch_bounds__(BPR_WARPS * 32)
bpr_kernel(const float* __restrict__ u, const float* __restrict__ p, const float* __restrict__ n, int64_t batch, int D,
           float wd, float wu, float wp, float wn, float inv_bs, float* __restrict__ loss, float* __restrict__ gu,
           float* __restrict__ gp, float* __restrict__ gn) {
    __shared__ float part[BPR_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t b = (int64_t)blockIdx.x * BPR_WARPS + warp;
    float contrib = 0.f;
    if (b < batch) {
        const float* ur = u + b * D;
        const float* pr = p + b * D;
        const float* nr = n + b * D;
        float xp = 0.f, xn = 0.f, su = 0.f, sp = 0.f, sn = 0.f;
        for (int c = lane; c < D; c += 32) {
            const float a = ur[c], bb = pr[c], cc = nr[c];
            xp = fmaf(a, bb, xp); xn = fmaf(a, cc, xn);
            su = fmaf(a, a, su); sp = fmaf(bb, bb, sp); sn = fmaf(cc, cc, sn);
        }
        xp = warp_sum(xp); xn = warp_sum(xn); su = warp_sum(su); sp = warp_sum(sp); sn = warp_sum(sn);
        const float x = fabsf(xp) - fabsf(xn);                                 // bprloss.py:18
        const float logsig = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));          // F.logsigmoid, bprloss.py:19
        contrib = (-logsig + wd * (wu * su + wp * sp + wn * sn)) * inv_bs;     // bprloss.py:20-22
        if (gu) {
            const float c = -1.f / (1.f + expf(x));                           // d(-logsig)/dx = -sigmoid(-x)
            const float sgp = xp > 0.f ? 1.f : (xp < 0.f ? -1.f : 0.f);
            const float sgn = xn > 0.f ? 1.f : (xn < 0.f ? -1.f : 0.f);
            const float cp = c * sgp * inv_bs, cn = -c * sgn * inv_bs;
            const float ru = 2.f * wd * wu * inv_bs, rp = 2.f * wd * wp * inv_bs, rn = 2.f * wd * wn * inv_bs;
            for (int col = lane; col < D; col += 32) {
                const float a = ur[col], bb = pr[col], cc = nr[col];
                gu[b * D + col] = cp * bb + cn * cc + ru * a;
                gp[b * D + col] = cp * a + rp * bb;
                gn[b * D + col] = cn * a + rn * cc;
            }
        }
    }
    if (lane == 0) part[warp] = contrib;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < BPR_WARPS; ++w) s += part[w];
        atomicAdd(loss, s);
    }
}

// ---- slot map for the gather backward ---------------------------------------------------------------
struct RowSets {
    const int64_t* rows[4];
    int64_t offset[4];
    const float* g[4];
    int64_t batch[4];
    int64_t base[4];     // first gsum row of set j
    int n_sets;
    int64_t total;
};

__device__ __forceinline__ bool locate(const RowSets& s, int64_t i, int& j, int64_t& b) {
    if (i >= s.total) return false;
    j = 0;
    while (j + 1 < s.n_sets && i >= s.base[j + 1]) ++j;
    b = i - s.base[j];
    return true;
}

__global__ void rowgrad_claim_kernel(RowSets s, int32_t* __restrict__ slot) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int j; int64_t b;
    if (!locate(s, i, j, b)) return;
    slot[s.rows[j][b] + s.offset[j]] = (int32_t)i;       // racy by design: any one occurrence wins
}

// one warp per (set, row): add the row gradient into the winner's gsum row
__global__ void rowgrad_accum_kernel(RowSets s, int D, const int32_t* __restrict__ slot, float* __restrict__ gsum) {
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    int j; int64_t b;
    if (!locate(s, i, j, b)) return;
    const int32_t w = slot[s.rows[j][b] + s.offset[j]];
    const float* g = s.g[j] + b * D;
    float* dst = gsum + (int64_t)w * D;
    for (int c = lane; c < D; c += 32) atomicAdd(dst + c, g[c]);
}

// one warp per winning slot: blocks 1..K of its summed output-row gradient go through the backward of
// H = E'/max(||E'||, 1e-12) (F.normalize, NGCF.py:144) in place; block 0 (the raw table row, NGCF.py:121) stays
struct NormArgs {
    const float* layer[NGCF_MAX_LAYERS + 1];
    int dim[NGCF_MAX_LAYERS + 1];
    int n;
};
__global__ void rowgrad_normalize_kernel(RowSets s, NormArgs L, int D, const int32_t* __restrict__ slot,
                                         float* __restrict__ gsum) {
    const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    int j; int64_t b;
    if (!locate(s, i, j, b)) return;
    const int64_t row = s.rows[j][b] + s.offset[j];
    if (slot[row] != (int32_t)i) return;                                  // a duplicate of the row: its winner does it
    float* g = gsum + i * D;
    int col = L.dim[0];
    for (int k = 1; k < L.n; ++k) {
        const int d = L.dim[k];
        const float* e = L.layer[k] + row * d;
        float nrm2 = 0.f, dot = 0.f;
        for (int c = lane; c < d; c += 32) {
            nrm2 = fmaf(e[c], e[c], nrm2);
            dot = fmaf(e[c], g[col + c], dot);
        }
        nrm2 = warp_sum(nrm2);
        dot = warp_sum(dot);
        const float n = fmaxf(sqrtf(nrm2), 1e-12f);
        const float hd = dot / n;                                         // H . gH
        for (int c = lane; c < d; c += 32) g[col + c] = (g[col + c] - (e[c] / n) * hd) / n;
        col += d;
    }
}

__global__ void rowgrad_reset_kernel(RowSets s, int32_t* __restrict__ slot) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int j; int64_t b;
    if (!locate(s, i, j, b)) return;
    slot[s.rows[j][b] + s.offset[j]] = -1;
}

int fill_sets(RowSets& s, const int64_t* const* rows_host, const int64_t* offsets_host, const float* const* g_host,
              const int64_t* batch_host, int n_sets) {
    NGCF_REQUIRE(rows_host && offsets_host && batch_host, "rowgrad: null host array");
    NGCF_REQUIRE(n_sets >= 1 && n_sets <= 4, "rowgrad: n_sets %d not in [1,4]", n_sets);
    s.n_sets = n_sets;
    s.total = 0;
    for (int j = 0; j < 4; ++j) {
        s.rows[j] = nullptr; s.g[j] = nullptr; s.offset[j] = 0; s.batch[j] = 0; s.base[j] = 0;
    }
    for (int j = 0; j < n_sets; ++j) {
        NGCF_REQUIRE(batch_host[j] >= 0 && (batch_host[j] == 0 || rows_host[j]), "rowgrad: set %d has no rows", j);
        s.rows[j] = rows_host[j];
        s.offset[j] = offsets_host[j];
        s.g[j] = g_host ? g_host[j] : nullptr;
        s.batch[j] = batch_host[j];
        s.base[j] = s.total;
        s.total += batch_host[j];
    }
    NGCF_REQUIRE(s.total < ((int64_t)1 << 31), "rowgrad: too many rows");
    return NGCF_OK;
}

}  // namespace

extern "C" int ngcf_gather_concat(const float* const* layers_host, const int* dims_host, int n_layers_plus1,
                                  const int64_t* rows, int64_t row_offset, int64_t n_out, float* out, int64_t ld_out,
                                  void* stream) {
    NGCF_REQUIRE(layers_host && dims_host && out, "gather_concat: null pointer");
    NGCF_REQUIRE(n_layers_plus1 >= 1 && n_layers_plus1 <= NGCF_MAX_LAYERS + 1, "gather_concat: %d blocks", n_layers_plus1);
    GatherArgs a;
    int D = 0;
    for (int k = 0; k < n_layers_plus1; ++k) {
        NGCF_REQUIRE(layers_host[k] && dims_host[k] > 0 && dims_host[k] <= NGCF_MAX_WIDTH, "gather_concat: bad block %d", k);
        a.layer[k] = layers_host[k];
        a.dim[k] = dims_host[k];
        D += dims_host[k];
    }
    NGCF_REQUIRE(ld_out >= D, "gather_concat: ld_out %lld < total width %d", (long long)ld_out, D);
    a.n = n_layers_plus1; a.rows = rows; a.row_offset = row_offset; a.n_out = n_out; a.out = out; a.ld_out = ld_out;
    if (n_out <= 0) return NGCF_OK;
    gather_concat_kernel<<<(unsigned)ceil_div64(n_out * 32, 256), 256, 0, as_stream(stream)>>>(a);
    NGCF_LAUNCH_OK("gather_concat_kernel");
    return NGCF_OK;
}

extern "C" int ngcf_gather_concat_sets(const float* const* layers_host, const int* dims_host, int n_layers_plus1,
                                       const int64_t* const* rows_host, const int64_t* offsets_host,
                                       const int64_t* sizes_host, float* const* outs_host, int n_sets, int64_t ld_out,
                                       void* stream) {
    NGCF_REQUIRE(layers_host && dims_host && rows_host && offsets_host && sizes_host && outs_host, "gather_concat_sets: null pointer");
    NGCF_REQUIRE(n_layers_plus1 >= 1 && n_layers_plus1 <= NGCF_MAX_LAYERS + 1, "gather_concat_sets: %d blocks", n_layers_plus1);
    NGCF_REQUIRE(n_sets >= 1 && n_sets <= 4, "gather_concat_sets: n_sets %d not in [1,4]", n_sets);
    GatherSetsArgs a{};
    int D = 0;
    for (int k = 0; k < n_layers_plus1; ++k) {
        NGCF_REQUIRE(layers_host[k] && dims_host[k] > 0 && dims_host[k] <= NGCF_MAX_WIDTH, "gather_concat_sets: bad block %d", k);
        a.layer[k] = layers_host[k];
        a.dim[k] = dims_host[k];
        D += dims_host[k];
    }
    NGCF_REQUIRE(ld_out >= D, "gather_concat_sets: ld_out %lld < total width %d", (long long)ld_out, D);
    a.n = n_layers_plus1; a.n_sets = n_sets; a.ld_out = ld_out;
    int64_t total = 0;
    for (int j = 0; j < n_sets; ++j) {
        NGCF_REQUIRE(sizes_host[j] >= 0 && (sizes_host[j] == 0 || (rows_host[j] && outs_host[j])), "gather_concat_sets: set %d", j);
        a.rows[j] = rows_host[j]; a.row_offset[j] = offsets_host[j]; a.out[j] = outs_host[j];
        a.base[j] = total;
        total += sizes_host[j];
    }
    for (int j = n_sets; j <= 4; ++j) a.base[j] = total;
    if (total == 0) return NGCF_OK;
    gather_concat_sets_kernel<<<(unsigned)ceil_div64(total * 32, 256), 256, 0, as_stream(stream)>>>(a);
    NGCF_LAUNCH_OK("gather_concat_sets_kernel");
    return NGCF_OK;
}

extern "C" int ngcf_bpr_fwd_bwd(const float* u, const float* p, const float* n, int64_t batch, int D,
                                float weight_decay, float batch_size_ctor, float reg_w_u, float reg_w_p,
                                float reg_w_n, float* loss, float* gu, float* gp, float* gn, void* stream) {
    NGCF_REQUIRE(u && p && n && loss, "bpr: null pointer");
    NGCF_REQUIRE((gu && gp && gn) || (!gu && !gp && !gn), "bpr: give all three gradient buffers or none");
    NGCF_REQUIRE(batch >= 0 && D > 0 && batch_size_ctor != 0.f, "bpr: bad sizes");
    cudaStream_t st = as_stream(stream);
    NGCF_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    if (batch == 0) return NGCF_OK;
    bpr_kernel<<<(unsigned)ceil_div64(batch, BPR_WARPS), BPR_WARPS * 32, 0, st>>>(
        u, p, n, batch, D, weight_decay, reg_w_u, reg_w_p, reg_w_n, 1.0f / batch_size_ctor, loss, gu, gp, gn);
    NGCF_LAUNCH_OK("bpr_kernel");
    return NGCF_OK;
}

extern "C" int ngcf_rowgrad_scatter(const int64_t* const* rows_host, const int64_t* offsets_host,
                                    const float* const* g_host, const int64_t* batch_host, int n_sets, int D,
                                    int32_t* slot, float* gsum, void* stream) {
    NGCF_REQUIRE(g_host && slot && gsum && D > 0, "rowgrad_scatter: null pointer");
    RowSets s;
    int rc = fill_sets(s, rows_host, offsets_host, g_host, batch_host, n_sets);
    if (rc != NGCF_OK) return rc;
    for (int j = 0; j < n_sets; ++j) NGCF_REQUIRE(s.batch[j] == 0 || s.g[j], "rowgrad_scatter: set %d has no gradient", j);
    if (s.total == 0) return NGCF_OK;
    cudaStream_t st = as_stream(stream);
    NGCF_CUDA(cudaMemsetAsync(gsum, 0, sizeof(float) * (size_t)s.total * D, st));
    rowgrad_claim_kernel<<<(unsigned)ceil_div64(s.total, 256), 256, 0, st>>>(s, slot);
    NGCF_LAUNCH_OK("rowgrad_claim_kernel");
    rowgrad_accum_kernel<<<(unsigned)ceil_div64(s.total * 32, 256), 256, 0, st>>>(s, D, slot, gsum);
    NGCF_LAUNCH_OK("rowgrad_accum_kernel");
    return NGCF_OK;
}

extern "C" int ngcf_rowgrad_normalize(const int64_t* const* rows_host, const int64_t* offsets_host,
                                      const int64_t* batch_host, int n_sets, const float* const* layers_host,
                                      const int* dims_host, int n_layers_plus1, const int32_t* slot, float* gsum,
                                      int D, void* stream) {
    NGCF_REQUIRE(slot && gsum && layers_host && dims_host, "rowgrad_normalize: null pointer");
    NGCF_REQUIRE(n_layers_plus1 >= 1 && n_layers_plus1 <= NGCF_MAX_LAYERS + 1, "rowgrad_normalize: %d blocks", n_layers_plus1);
    RowSets s;
    int rc = fill_sets(s, rows_host, offsets_host, nullptr, batch_host, n_sets);
    if (rc != NGCF_OK) return rc;
    NormArgs L;
    int total = 0;
    for (int k = 0; k < n_layers_plus1; ++k) {
        NGCF_REQUIRE(layers_host[k] && dims_host[k] > 0 && dims_host[k] <= NGCF_MAX_WIDTH, "rowgrad_normalize: bad block %d", k);
        L.layer[k] = layers_host[k];
        L.dim[k] = dims_host[k];
        total += dims_host[k];
    }
    L.n = n_layers_plus1;
    NGCF_REQUIRE(total == D, "rowgrad_normalize: block widths sum to %d, D = %d", total, D);
    if (s.total == 0 || n_layers_plus1 == 1) return NGCF_OK;
    rowgrad_normalize_kernel<<<(unsigned)ceil_div64(s.total * 32, 256), 256, 0, as_stream(stream)>>>(s, L, D, slot, gsum);
    NGCF_LAUNCH_OK("rowgrad_normalize_kernel");
    return NGCF_OK;
}

extern "C" int ngcf_rowgrad_reset(const int64_t* const* rows_host, const int64_t* offsets_host,
                                  const int64_t* batch_host, int n_sets, int32_t* slot, void* stream) {
    NGCF_REQUIRE(slot, "rowgrad_reset: null pointer");
    RowSets s;
    int rc = fill_sets(s, rows_host, offsets_host, nullptr, batch_host, n_sets);
    if (rc != NGCF_OK) return rc;
    if (s.total == 0) return NGCF_OK;
    rowgrad_reset_kernel<<<(unsigned)ceil_div64(s.total, 256), 256, 0, as_stream(stream)>>>(s, slot);
    NGCF_LAUNCH_OK("rowgrad_reset_kernel");
    return NGCF_OK;
}
