// Tensor-core (tcgen05 / TMEM) version of the per-layer forward epilogue, NGCF.py:131-142:
//     E_out = Dropout(LeakyReLU((S+E)·W1^T + (S*E)·W2^T + 2 b1 + b2))
// as ONE GEMM per 128-row tile:  [128 x 2d] · [2d x d_out], A = [S+E | S*E] built on the fly, B = [W1 | W2]
// resident in shared memory for the whole (persistent) kernel, fp32 accumulator in TMEM, 3xTF32 (tc.cuh).
//
// Warp roles (416 threads, one CTA per SM):
//   warps 0-3   epilogue: tcgen05.ld the accumulator (TMEM lane = tile row), bias + LeakyReLU + dropout, store
//   warp  4     TMEM allocation; lane 0 issues every tcgen05.mma / tcgen05.commit
//   warps 5-12  loaders: coalesced 128-bit reads of S and E, split into TF32 hi/lo, written straight into the
//               128-byte-swizzled K-major operand layout
// A is pipelined per 32-wide K block: block kb of the next tile is refilled as soon as the MMAs that read it have
// completed (tcgen05.commit -> empty[kb]); the accumulator is double buffered in TMEM so the epilogue of tile t
// overlaps the loads and MMAs of tile t+1.
#include "tc.cuh"

namespace {

using namespace tc;

constexpr int TC_ROWS = 128;                 // UMMA M
constexpr int TC_EPI_WARPS = 4;
constexpr int TC_LOAD_WARPS = 8;
constexpr int TC_THREADS = (TC_EPI_WARPS + 1 + TC_LOAD_WARPS) * 32;
constexpr int TC_MAX_KB = 4;                 // K blocks of 32 TF32 (128 bytes): 2*d_in/32 <= 4  ->  d_in <= 64
constexpr int TC_A_BLOCK = TC_ROWS * 128;    // bytes of one K block of A (hi or lo)

struct FwdTcArgs {
    const float* S;
    const float* E;
    int64_t n_rows;
    int d_in, d_out;
    const float* wcat;        // [2*d_in, d_out]: wcat[k][o] = W1[o][k] (k < d_in), W2[o][k-d_in]
    const float* bias_eff;    // [d_out]
    float slope;
    const float* mess_mult;
    float mess_p;
    uint64_t seed;
    const uint64_t* seed_dev;
    int layer;
    float* E_out;
    int n_tiles;
};

struct Bars {
    uint64_t full[TC_MAX_KB];
    uint64_t empty[TC_MAX_KB];
    uint64_t tmem_full[2];
    uint64_t tmem_empty[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(TC_THREADS, 1) dense_fwd_tc_kernel(FwdTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;          // operand tiles need 1024-byte alignment
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const int d_in = a.d_in, d_out = a.d_out;
    const int KBH = d_in / 32, KB = 2 * KBH;                              // K blocks per half / in total
    const int b_block = d_out * 128;                                      // bytes of one K block of B (hi or lo)
    uint8_t* A_hi = smem;
    uint8_t* A_lo = A_hi + KB * TC_A_BLOCK;
    uint8_t* B_hi = A_lo + KB * TC_A_BLOCK;
    uint8_t* B_lo = B_hi + KB * b_block;
    float* bias_s = reinterpret_cast<float*>(B_lo + KB * b_block);
    Bars* bars = reinterpret_cast<Bars*>(bias_s + 64);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- one-off setup: barriers, TMEM, weights ----------------------------------------------------------------
    if (tid == 0) {
        for (int i = 0; i < TC_MAX_KB; ++i) {
            mbar_init(&bars->full[i], TC_LOAD_WARPS);
            mbar_init(&bars->empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&bars->tmem_full[i], 1);
            mbar_init(&bars->tmem_empty[i], TC_EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == TC_EPI_WARPS) tmem_alloc(&bars->tmem_base, 128);          // 2 accumulators x 64 fp32 columns
    // B[n][k] = wcat[k][n], split and swizzled (row n of K block kb: 32 values of k)
    for (int i = tid; i < d_out * KB * 8; i += TC_THREADS) {
        const int c = i & 7, n = (i >> 3) % d_out, kb = (i >> 3) / d_out;
        float4 w;
        const float* src = a.wcat + (int64_t)(kb * 32 + c * 4) * d_out + n;
        w.x = src[0]; w.y = src[d_out]; w.z = src[2 * d_out]; w.w = src[3 * d_out];
        float4 hi, lo;
        split_tf32(w, hi, lo);
        const uint32_t off = kb * b_block + sw128_offset(n, c);
        *reinterpret_cast<float4*>(B_hi + off) = hi;
        *reinterpret_cast<float4*>(B_lo + off) = lo;
    }
    for (int i = tid; i < 64; i += TC_THREADS) bias_s[i] = i < d_out ? a.bias_eff[i] : 0.f;
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = bars->tmem_base;

    const int n_my = (a.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

    if (warp < TC_EPI_WARPS) {
        // ======================= epilogue =======================================================================
        const uint64_t seed = a.mess_p > 0.f ? ngcf_seed(a.seed, a.seed_dev) : 0ull;
        for (int it = 0; it < n_my; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int buf = it & 1;
            mbar_wait(&bars->tmem_full[buf], (it >> 1) & 1);
            tc_fence_after_sync();
            float acc[64];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * 64;
            {
                float lo[32], hi[32];
                tmem_ld_32x32(taddr, lo);
                if (d_out > 32) tmem_ld_32x32(taddr + 32, hi);
#pragma unroll
                for (int j = 0; j < 32; ++j) { acc[j] = lo[j]; acc[32 + j] = d_out > 32 ? hi[j] : 0.f; }
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->tmem_empty[buf]);          // accumulator drained: next tile may reuse it
            const int64_t row = (int64_t)tile * TC_ROWS + warp * 32 + lane;
            if (row < a.n_rows) {
                float* out = a.E_out + row * d_out;
#pragma unroll
                for (int j = 0; j < 64; j += 4) {
                    if (j < d_out) {
                        float o[4];
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const float m = acc[j + c] + bias_s[j + c];
                            o[c] = m > 0.f ? m : a.slope * m;                       // LeakyReLU, NGCF.py:140
                        }
                        if (a.mess_mult) {
                            const float4 mm = ld_f4(a.mess_mult + row * d_out + j);
                            o[0] *= mm.x; o[1] *= mm.y; o[2] *= mm.z; o[3] *= mm.w;
                        } else if (a.mess_p > 0.f) {
                            const float4 mm = mess_multiplier4(a.mess_p, seed, a.layer, (uint64_t)(row * d_out + j) >> 2);
                            o[0] *= mm.x; o[1] *= mm.y; o[2] *= mm.z; o[3] *= mm.w;
                        }
                        st_f4(out + j, make_float4(o[0], o[1], o[2], o[3]));
                    }
                }
            }
        }
    } else if (warp == TC_EPI_WARPS) {
        // ======================= MMA issuer =====================================================================
        const uint32_t idesc = umma_idesc_tf32(TC_ROWS, d_out, 0, 0);
        for (int it = 0; it < n_my; ++it) {
            const int buf = it & 1;
            mbar_wait(&bars->tmem_empty[buf], ((it >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
            tc_fence_after_sync();
            const uint32_t tmem_d = tmem_base + buf * 64;
            uint32_t accumulate = 0;
            for (int hh = 0; hh < KBH; ++hh) {
                for (int part = 0; part < 2; ++part) {
                    const int kb = part * KBH + hh;
                    mbar_wait(&bars->full[kb], it & 1);
                    tc_fence_after_sync();
                    if (lane == 0) {
                        const uint32_t a_hi = smem_u32(A_hi + kb * TC_A_BLOCK), a_lo = smem_u32(A_lo + kb * TC_A_BLOCK);
                        const uint32_t b_hi = smem_u32(B_hi + kb * b_block), b_lo = smem_u32(B_lo + kb * b_block);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {                     // 4 x (K = 8 TF32 = 32 bytes) per block
                            const uint64_t dah = umma_desc_sw128(a_hi + k * 32, 16, 1024);
                            const uint64_t dal = umma_desc_sw128(a_lo + k * 32, 16, 1024);
                            const uint64_t dbh = umma_desc_sw128(b_hi + k * 32, 16, 1024);
                            const uint64_t dbl = umma_desc_sw128(b_lo + k * 32, 16, 1024);
                            umma_tf32(tmem_d, dah, dbh, idesc, accumulate);
                            umma_tf32(tmem_d, dal, dbh, idesc, 1);
                            umma_tf32(tmem_d, dah, dbl, idesc, 1);
                            accumulate = 1;
                        }
                        umma_commit(&bars->empty[kb]);                    // A block kb may be refilled
                    }
                    __syncwarp();
                }
            }
            if (lane == 0) umma_commit(&bars->tmem_full[buf]);            // accumulator complete
            __syncwarp();
        }
    } else {
        // ======================= loaders ========================================================================
        const int lt = tid - (TC_EPI_WARPS + 1) * 32;                     // 0 .. 255
        constexpr int LT = TC_LOAD_WARPS * 32;
        for (int it = 0; it < n_my; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int64_t row0 = (int64_t)tile * TC_ROWS;
            for (int hh = 0; hh < KBH; ++hh) {
                const int kb1 = hh, kb2 = KBH + hh;
                float4 s[4], e[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {                             // 128 rows x 8 chunks = 4 per thread
                    const int idx = q * LT + lt, r = idx >> 3, c = idx & 7;
                    const int64_t row = row0 + r;
                    s[q] = e[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (row < a.n_rows) {
                        s[q] = ld_f4(a.S + row * d_in + hh * 32 + c * 4);
                        e[q] = ld_f4(a.E + row * d_in + hh * 32 + c * 4);
                    }
                }
                mbar_wait(&bars->empty[kb1], (it & 1) ^ 1);
                mbar_wait(&bars->empty[kb2], (it & 1) ^ 1);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int idx = q * LT + lt, r = idx >> 3, c = idx & 7;
                    const uint32_t off = sw128_offset(r, c);
                    const float4 x1 = make_float4(s[q].x + e[q].x, s[q].y + e[q].y, s[q].z + e[q].z, s[q].w + e[q].w);
                    const float4 x2 = make_float4(s[q].x * e[q].x, s[q].y * e[q].y, s[q].z * e[q].z, s[q].w * e[q].w);
                    float4 hi, lo;
                    split_tf32(x1, hi, lo);
                    *reinterpret_cast<float4*>(A_hi + kb1 * TC_A_BLOCK + off) = hi;
                    *reinterpret_cast<float4*>(A_lo + kb1 * TC_A_BLOCK + off) = lo;
                    split_tf32(x2, hi, lo);
                    *reinterpret_cast<float4*>(A_hi + kb2 * TC_A_BLOCK + off) = hi;
                    *reinterpret_cast<float4*>(A_lo + kb2 * TC_A_BLOCK + off) = lo;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&bars->full[kb1]);
                    mbar_arrive(&bars->full[kb2]);
                }
            }
        }
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == TC_EPI_WARPS) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 128);
    }
}

}  // namespace

bool ngcf_dense_fwd_tc_eligible(int d_in, int d_out) {
    return (d_in == 32 || d_in == 64) && d_out >= 16 && d_out <= 64 && d_out % 16 == 0;
}

int ngcf_dense_fwd_tc(const float* S, const float* E, int64_t n_rows, int d_in, int d_out, const float* wcat,
                      const float* bias_eff, float slope, const float* mess_mult, float mess_p, uint64_t seed,
                      const uint64_t* seed_dev, int layer, float* E_out, cudaStream_t st) {
    FwdTcArgs a{S, E, n_rows, d_in, d_out, wcat, bias_eff, slope, mess_mult, mess_p, seed, seed_dev, layer, E_out,
                (int)ceil_div64(n_rows, TC_ROWS)};
    const int KB = 2 * d_in / 32;
    const size_t smem = 1024 + (size_t)KB * (2 * TC_A_BLOCK + 2 * d_out * 128) + 64 * sizeof(float) + sizeof(Bars);
    static bool attr_set = false;
    if (!attr_set) {
        NGCF_CUDA(cudaFuncSetAttribute(dense_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int grid = (int)min((int64_t)a.n_tiles, (int64_t)ngcf_num_sms());
    dense_fwd_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a);
    NGCF_LAUNCH_OK("dense_fwd_tc_kernel");
    return NGCF_OK;
}
