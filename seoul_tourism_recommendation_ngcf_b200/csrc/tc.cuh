// sm_100a tensor-core plumbing used by the dense parts of the NGCF layer: mbarrier, tcgen05 (UMMA) descriptors for
// K-major / MN-major TF32 operands in 128-byte-swizzled shared memory, TMEM allocation and TMEM -> register loads.
//
// Numerics: the reference computes the W1/W2 products in fp32 (cuBLAS SGEMM / MKL).  A single TF32 product has a
// 2^-11 relative rounding error per operand, which breaks the 1e-4 parity bound, so every product here is the
// error-compensated "3xTF32" split  a·b ~= a_hi·b_hi + a_lo·b_hi + a_hi·b_lo  with a_hi = a rounded to TF32 and
// a_lo = a - a_hi (exact in fp32): three tcgen05.mma per K step into the same fp32 TMEM accumulator.
#pragma once
#include "common.cuh"

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Waits until the phase with the given parity has completed (try_wait suspends the warp in hardware; an explicit
// nanosleep back-off between polls was measured and changes nothing).  A protocol bug would otherwise hang the GPU;
// after ~2^28 polls (seconds) the kernel traps instead so the launch fails with an error.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; ; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (spin > (1u << 28)) __trap();
    }
}
// transaction-count arrival: the barrier's phase completes once this arrival AND `bytes` of async-copy data have landed
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// ---- TMA (cp.async.bulk.tensor): one thread copies a whole 2-D box global -> shared memory, swizzled as the tensor map
// says, completion signalled on an mbarrier in bytes; out-of-range rows of the box are zero-filled.  The tensor map is a
// __grid_constant__ kernel parameter built on the host (tmap.h).
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void* tmap, int x, int y, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst_smem), "l"(tmap), "r"(x), "r"(y), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}

// generic-proxy shared-memory writes -> visible to the async proxy (tcgen05.mma operand reads)
#ifdef NGCF_NO_PROXY_FENCE
__device__ __forceinline__ void fence_proxy_async_smem() {}
#else
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// ---- cp.async (LDGSTS): 16 bytes global -> shared without a register stage; src_bytes < 16 zero-fills the rest ------
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- tcgen05 ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// whole-warp calls
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// Shared-memory matrix descriptor, 128-byte swizzle, tile base 1024-byte aligned (cute::UMMA::SmemDescriptor:
// start address >> 4 in bits [0,14), leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46),
// version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64)).
//   K-major operand  (rows of 128 bytes = 32 TF32 along K; 8-row groups of 1024 bytes): LBO unused (1), SBO = bytes
//   between 8-row groups.
//   MN-major operand (rows of 128 bytes = 32 TF32 along M/N; 8 K-rows per 1024-byte atom): LBO = bytes between
//   32-element M/N groups, SBO = bytes between 8-row K groups.
//   32-bit MN-major operands exist in ONE layout only, SWIZZLE_128B_BASE32B = 1 (CUTLASS sm100_common.inl: "for
//   mn-major tf32 operands, SW128_32B is the only available smem layout"): rows of 128 bytes along M/N, 4 K-rows
//   per 512-byte atom, the 32-byte chunk index XORed with the K-row index inside the atom (Swizzle<2,5,2> on byte
//   addresses); LBO = bytes between 32-element M/N groups, SBO = bytes between 4-row K groups.
constexpr uint32_t UMMA_SW128 = 2, UMMA_SW128_BASE32B = 1;
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return umma_desc(saddr, lbo_bytes, sbo_bytes, UMMA_SW128);
}

// Instruction descriptor for kind::tf32, fp32 accumulate (cute::UMMA::InstrDescriptor): c_format F32 = 1 at [4,6),
// a/b_format TF32 = 2 at [7,10)/[10,13), a/b_major at 15/16 (0 = K-major, 1 = MN-major), N >> 3 at [17,23),
// M >> 4 at [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] · B[smem]; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the mbarrier receives one arrival when every tcgen05 operation this thread issued so far has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// TMEM -> registers: lane i of the warp receives 32 consecutive fp32 columns of TMEM lane (taddr.lane + i).
// A warp may only touch the 32-lane quarter (warp_id % 4) of TMEM.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- 3xTF32 operand split and the swizzled operand layouts -----------------------------------------------------------
// hi = x rounded to nearest TF32, lo = the exactly representable remainder rounded to TF32 as well, so the tensor
// core's operand truncation changes neither.  Round-to-nearest keeps the split errors zero-mean: with a truncating
// split they all share the sign of x and add up coherently over the ~10^5-row weight-gradient sums.
// (cvt.rna.tf32.f32 is not a native sm_100a instruction: ptxas expands it to add / inf-test / select / mask.  For
// finite inputs the add-half-ulp-and-mask below is the same round-to-nearest in two integer instructions; an
// infinity stays an infinity, and the operands here are finite by construction.)
__device__ __forceinline__ float rna_tf32(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = rna_tf32(x);
    lo = rna_tf32(x - hi);
}
__device__ __forceinline__ void split_tf32(const float4& x, float4& hi, float4& lo) {
    split_tf32(x.x, hi.x, lo.x);
    split_tf32(x.y, hi.y, lo.y);
    split_tf32(x.z, hi.z, lo.z);
    split_tf32(x.w, hi.w, lo.w);
}

// Cheap split for the streamed A operands of the forward GEMM and of the backward's GEMM 1 (the loader warps are
// instruction-bound): hi = x as it is — the tensor core ignores the low 13 mantissa bits of a TF32 operand, i.e.
// truncates — and lo = x - trunc(x), exact in fp32 (its own truncation to TF32 loses < 2^-21 |x|).  Truncation errors
// share the sign of x, which is harmless over these K <= 128 contractions (measured ~1e-6 vs float64) but is why the
// weight-gradient GEMM (contraction over ~10^5 rows) keeps the round-to-nearest split above.
__device__ __forceinline__ void split_tf32_trunc(const float4& x, float4& hi, float4& lo) {
    hi = x;
    lo.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
    lo.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
    lo.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
    lo.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
}

// byte offset of the 16-byte chunk c (0..7) of row r inside a K-major (or MN-major) 128-byte-swizzled block whose
// rows are 128 bytes: 8-row atoms of 1024 bytes, chunk index XORed with the row index inside the atom
__device__ __forceinline__ uint32_t sw128_offset(uint32_t r, uint32_t c) {
    return (r >> 3) * 1024u + (r & 7u) * 128u + ((c ^ (r & 7u)) << 4);
}

// byte offset of the 16-byte chunk c (0..7) of K-row r inside an MN-major SWIZZLE_128B_BASE32B block (128-byte rows)
__device__ __forceinline__ uint32_t sw128b32_offset(uint32_t r, uint32_t c) {
    return r * 128u + ((((c >> 1) ^ (r & 3u)) << 5) | ((c & 1u) << 4));
}

}  // namespace tc
