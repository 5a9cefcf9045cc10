// Micro-benchmark behind DESIGN.md section 4 ("what bounds the SpMM"): how fast can one B200 move randomly chosen
// 256-byte rows of an L2-resident table ([70839, 64] fp32 = 18 MB, the Gowalla-shaped E) into its SMs, by
//   (a) LDG.128, half a warp per row, U rows in flight per lane group (the path spmm_tile_kernel uses),
//   (b) cp.async.bulk (UBLKCP), one 256-byte bulk copy per row into a shared-memory ring, mbarrier completion,
//   (c) cp.async.bulk.tensor ... tile::gather4 (UTMALDG), four rows per instruction through a 2-D tensor map.
// The consumers only touch the data enough to keep the copies alive (LDG: a register sum; bulk/TMA: one LDS per row),
// so each figure is an upper bound for a gather-and-accumulate kernel on that path.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tools/l2_gather_bench.cu -o tools/l2_gather_bench
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__);     \
            exit(2);                                                                           \
        }                                                                                      \
    } while (0)

constexpr int D = 64;                 // floats per row (256 bytes)
__device__ int g_fail = 0;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint64_t* b, uint32_t parity) {
    const uint32_t addr = smem_u32(b);
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return true;
    }
    atomicExch(&g_fail, 1);
    return false;
}

// ---- (a) LDG.128 --------------------------------------------------------------------------------------------------
template <int U>
__global__ void __launch_bounds__(64) ldg_kernel(const float* __restrict__ E, const int* __restrict__ idx, int n_per_cta,
                                                 float* out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 4, l = lane & 15;
    const int* my = idx + (size_t)blockIdx.x * n_per_cta;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    // each warp: rows j = warp*2 + g, step 4 (two warps x two lane groups)
    for (int j0 = warp * 2 + g; j0 + 4 * (U - 1) < n_per_cta; j0 += 4 * U) {
        int c[U];
        float4 x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) c[u] = my[j0 + 4 * u];
#pragma unroll
        for (int u = 0; u < U; ++u) x[u] = *reinterpret_cast<const float4*>(E + (size_t)c[u] * D + l * 4);
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}

// ---- (d) the SpMM's CTA structure: one short-lived CTA per 512-entry tile; optionally the tile's ids are first staged in
// shared memory by a coalesced read + barrier (what spmm_tile_kernel / spmm_stream_kernel do), then gathered from there;
// PERSIST: the same tiles, but each CTA of a one-wave grid loops over tiles blockIdx.x, + gridDim.x, ...
template <int U, bool STAGE, bool PERSIST>
__global__ void __launch_bounds__(64, 20) tile_kernel(const float* __restrict__ E, const int* __restrict__ idx, int n_tiles,
                                                      float* out) {
    __shared__ int ids[512];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 4, l = lane & 15;
    const int grp = warp * 2 + g;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = blockIdx.x; t < n_tiles; t += PERSIST ? gridDim.x : n_tiles) {
        const int* my = idx + (size_t)t * 512;
        if (STAGE) {
            __syncthreads();
            for (int i = threadIdx.x; i < 512; i += 64) ids[i] = my[i];
            __syncthreads();
        }
        const int* src = STAGE ? ids : my;
        for (int j0 = grp * 128; j0 < (grp + 1) * 128; j0 += U) {      // contiguous quarter of the tile per lane group
            int c[U];
            float4 x[U];
#pragma unroll
            for (int u = 0; u < U; ++u) c[u] = src[j0 + u];
#pragma unroll
            for (int u = 0; u < U; ++u) x[u] = *reinterpret_cast<const float4*>(E + (size_t)c[u] * D + l * 4);
#pragma unroll
            for (int u = 0; u < U; ++u) { acc.x += x[u].x; acc.y += x[u].y; acc.z += x[u].z; acc.w += x[u].w; }
        }
    }
    if (acc.x + acc.y + acc.z + acc.w == 123.456f) out[0] = acc.x;
}

// ---- (b) cp.async.bulk, one 256-byte copy per row; (c) gather4 -------------------------------------------------------
// One producer warp issues the copies of a stage (ROWS rows) against the stage's mbarrier; CONS consumer warps wait, read
// one word per row and release the stage.
template <int ROWS, int STAGES, bool GATHER4>
__global__ void __launch_bounds__(32 * 5) bulk_kernel(const float* __restrict__ E, const int* __restrict__ idx,
                                                      int n_per_cta, float* out, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[STAGES], empty[STAGES];
    constexpr int CONS = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CONS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int* my = idx + (size_t)blockIdx.x * n_per_cta;
    const int n_stage = n_per_cta / ROWS;
    if (warp == CONS) {                                                   // producer
        for (int it = 0; it < n_stage; ++it) {
            const int s = it % STAGES;
            if (it >= STAGES && !mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1)) return;
            if (lane == 0) mbar_expect_tx(&full[s], ROWS * D * 4);
            __syncwarp();
            uint8_t* dst = smem + (size_t)s * ROWS * D * 4;
            if (GATHER4) {
                for (int r = lane * 4; r < ROWS; r += 128) {
                    const int4 c = *reinterpret_cast<const int4*>(my + it * ROWS + r);
                    asm volatile(
                        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
                        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                        ::"r"(smem_u32(dst + r * D * 4)), "l"(&tmap), "r"(0), "r"(c.x), "r"(c.y), "r"(c.z), "r"(c.w),
                        "r"(smem_u32(&full[s]))
                        : "memory");
                }
            } else {
                for (int r = lane; r < ROWS; r += 32) {
                    const int c = my[it * ROWS + r];
                    asm volatile(
                        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                        ::"r"(smem_u32(dst + r * D * 4)), "l"(E + (size_t)c * D), "r"(D * 4), "r"(smem_u32(&full[s]))
                        : "memory");
                }
            }
        }
    } else {                                                              // consumers
        float acc = 0.f;
        for (int it = 0; it < n_stage; ++it) {
            const int s = it % STAGES;
            if (!mbar_wait(&full[s], (it / STAGES) & 1)) return;
            const float* src = reinterpret_cast<const float*>(smem + (size_t)s * ROWS * D * 4);
            for (int r = warp * 32 + lane; r < ROWS; r += CONS * 32) acc += src[r * D + (lane & 15)];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 123.456f) out[0] = acc;
    }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename F>
float time_ms(F launch, int reps) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms / reps;
}

int main(int argc, char** argv) {
    const int N = 70839;
    const int n_sm = 148;
    std::vector<float> hE((size_t)N * D, 1.0f);
    float* E;
    CK(cudaMalloc(&E, (size_t)N * D * 4));
    CK(cudaMemcpy(E, hE.data(), (size_t)N * D * 4, cudaMemcpyHostToDevice));
    const size_t M = (size_t)1 << 22;                       // 4 Mi gathered rows = 1 GiB moved
    std::vector<int> hidx(M);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < M; ++i) {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        hidx[i] = (int)(s % N);
    }
    // optional: a file of int32 row ids (e.g. the column stream of the real graph's SpMM, execution order), repeated to M
    const char* src = "uniform random";
    if (argc > 1) {
        FILE* f = fopen(argv[1], "rb");
        if (f) {
            std::vector<int> file;
            int buf[4096];
            size_t n;
            while ((n = fread(buf, 4, 4096, f)) > 0) file.insert(file.end(), buf, buf + n);
            fclose(f);
            if (!file.empty()) {
                for (size_t i = 0; i < M; ++i) hidx[i] = file[i % file.size()] % N;
                src = argv[1];
            }
        }
    }
    printf("row ids: %s\n", src);
    int* idx;
    CK(cudaMalloc(&idx, M * 4));
    CK(cudaMemcpy(idx, hidx.data(), M * 4, cudaMemcpyHostToDevice));
    float* out;
    CK(cudaMalloc(&out, 64));
    const double bytes = (double)M * D * 4;

    // (a) LDG: grid = 148 * ctas_per_sm, 64 threads each
    for (int cps : {8, 16, 20, 32}) {
        const int grid = n_sm * cps, per = (int)(M / grid) / 64 * 64;
        const double b = (double)per * grid * D * 4;
        float ms4 = time_ms([&] { ldg_kernel<4><<<grid, 64>>>(E, idx, per, out); }, 5);
        float ms8 = time_ms([&] { ldg_kernel<8><<<grid, 64>>>(E, idx, per, out); }, 5);
        float ms16 = time_ms([&] { ldg_kernel<16><<<grid, 64>>>(E, idx, per, out); }, 5);
        printf("ldg   ctas/sm=%2d  U=4: %7.2f TB/s  U=8: %7.2f TB/s  U=16: %7.2f TB/s\n", cps, b / ms4 / 1e9, b / ms8 / 1e9,
               b / ms16 / 1e9);
    }
    CK(cudaGetLastError());
    {
        const int n_tiles = (int)(M / 512);
        const double b = (double)n_tiles * 512 * D * 4;
        float a1 = time_ms([&] { tile_kernel<4, false, false><<<n_tiles, 64>>>(E, idx, n_tiles, out); }, 5);
        float a2 = time_ms([&] { tile_kernel<4, true, false><<<n_tiles, 64>>>(E, idx, n_tiles, out); }, 5);
        float a3 = time_ms([&] { tile_kernel<4, true, true><<<n_sm * 20, 64>>>(E, idx, n_tiles, out); }, 5);
        float a4 = time_ms([&] { tile_kernel<4, false, true><<<n_sm * 20, 64>>>(E, idx, n_tiles, out); }, 5);
        printf("tiles of 512 (U=4, 64 threads, 20 CTAs/SM): one CTA per tile %7.2f TB/s | + staged ids %7.2f TB/s | "
               "persistent + staged %7.2f TB/s | persistent, direct ids %7.2f TB/s\n", b / a1 / 1e9, b / a2 / 1e9, b / a3 / 1e9,
               b / a4 / 1e9);
        // the real product has 2.05 M entries = 4013 tiles = 1.36 waves of 2960 CTAs: same kernels on that many tiles
        const int nt2 = 4013;
        const double b2 = (double)nt2 * 512 * D * 4;
        float c1 = time_ms([&] { tile_kernel<4, true, false><<<nt2, 64>>>(E, idx, nt2, out); }, 20);
        float c2 = time_ms([&] { tile_kernel<4, true, true><<<n_sm * 20, 64>>>(E, idx, nt2, out); }, 20);
        printf("4013 tiles (= one product): one CTA per tile, staged %7.2f TB/s (%.1f us) | persistent, staged %7.2f TB/s (%.1f us)\n",
               b2 / c1 / 1e9, c1 * 1e3, b2 / c2 / 1e9, c2 * 1e3);
    }
    CK(cudaGetLastError());

    // tensor map over E: [N rows, 64 cols] fp32, box = {64, 1} (gather4 fetches four such rows)
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    bool have_tmap = false;
    if (encode) {
        cuuint64_t dims[2] = {(cuuint64_t)D, (cuuint64_t)N};
        cuuint64_t strides[1] = {(cuuint64_t)D * 4};
        cuuint32_t box[2] = {(cuuint32_t)D, 1};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, E, dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        have_tmap = r == CUDA_SUCCESS;
        printf("cuTensorMapEncodeTiled -> %d\n", (int)r);
    }

    auto run_bulk = [&](auto kern, int rows, int stages, int cps, const char* name) {
        const size_t smem = (size_t)rows * stages * D * 4;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = n_sm * cps, per = (int)(M / grid) / rows * rows;
        const double b = (double)per * grid * D * 4;
        float ms = time_ms([&] { kern<<<grid, 160, smem>>>(E, idx, per, out, tmap); }, 5);
        cudaError_t e = cudaGetLastError();
        int fail = 0;
        CK(cudaMemcpyFromSymbol(&fail, g_fail, sizeof(int)));
        printf("%-7s rows/stage=%3d stages=%d ctas/sm=%d (%3zu KB in flight/SM): %7.2f TB/s%s%s\n", name, rows, stages, cps,
               smem * cps >> 10, b / ms / 1e9, fail ? "  [mbarrier TIMEOUT]" : "", e != cudaSuccess ? cudaGetErrorString(e) : "");
        fail = 0;
        CK(cudaMemcpyToSymbol(g_fail, &fail, sizeof(int)));
    };
    run_bulk(bulk_kernel<128, 2, false>, 128, 2, 1, "bulk");
    run_bulk(bulk_kernel<128, 4, false>, 128, 4, 1, "bulk");
    run_bulk(bulk_kernel<128, 6, false>, 128, 6, 1, "bulk");
    run_bulk(bulk_kernel<128, 3, false>, 128, 3, 2, "bulk");
    run_bulk(bulk_kernel<64, 4, false>, 64, 4, 3, "bulk");
    run_bulk(bulk_kernel<32, 4, false>, 32, 4, 6, "bulk");
    if (have_tmap) {
        run_bulk(bulk_kernel<128, 2, true>, 128, 2, 1, "gather4");
        run_bulk(bulk_kernel<128, 4, true>, 128, 4, 1, "gather4");
        run_bulk(bulk_kernel<128, 6, true>, 128, 6, 1, "gather4");
        run_bulk(bulk_kernel<128, 3, true>, 128, 3, 2, "gather4");
        run_bulk(bulk_kernel<64, 4, true>, 64, 4, 3, "gather4");
    }
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
