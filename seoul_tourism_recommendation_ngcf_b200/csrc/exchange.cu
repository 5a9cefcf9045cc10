// Row-shard exchange over NVLink peer memory (multi-GPU runs).
//
// The reference has no multi-device code; the row partition prescribed for this path needs, per layer and direction,
// every rank's block of rows of one [N_pad, d] matrix on every rank (an all-gather).  Round 1 issued a bare
// ncclAllGather between the two kernels that depend on it (~50 us apiece inside the step's CUDA graph at 8 ranks).
// Here the owner PUSHES its rows: one kernel stores the block straight into every peer's copy of the matrix through
// peer-mapped pointers (symmetric memory), with two flag rounds around the stores:
//
//   enter : rank r tells every peer "I have reached exchange #e" - everything r launched before it has completed, so
//           r's copy of the matrix may be overwritten; a CTA storing to peer p first waits for p's word.
//   done  : after a CTA's stores it fences (system scope) and counts itself; the last CTA tells every peer "my rows of
//           #e have landed" and waits for the same word from all of them, so the kernel ends with the whole matrix
//           present locally and stream order does the rest.
//
// Flags are monotonically increasing exchange numbers (never reset), kept in symmetric memory next to a local counter,
// so a captured CUDA graph replays correctly.  A wait that does not complete within ~2^27 polls traps instead of hanging
// the GPU.  No kernel ever waits for another kernel of the SAME GPU.
#include "common.cuh"

namespace {

constexpr int EX_MAX_WORLD = 16;
constexpr int EX_THREADS = 256;

struct PushArgs {
    float* dst[EX_MAX_WORLD];        // base of the [N_pad, d] matrix on every rank (peer-mapped; [rank] = local)
    uint32_t* flags[EX_MAX_WORLD];   // base of every rank's flag block: uint32 [2][EX_MAX_WORLD] = {enter, done} x source rank
    uint32_t* local_state;           // [2] = {exchange number of the last completed exchange, CTA counter}
    int world, rank;
    int64_t elem0;                   // first element of this rank's block inside the matrix
    int64_t n_elem;                  // elements of the block (multiple of 4)
    int ctas_per_peer;
    float* mc;                       // multicast mapping of the matrix (one store reaches every rank's copy), or NULL
    // selected-rows variant: only the rows named by the lists (list[j] + list_off) that this rank owns travel
    const int64_t* lists[4];
    int64_t list_off[4], list_n[4];
    int n_lists;
    int64_t row0, n_rows;
    int d;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// one store that the NVSwitch replicates into every rank's copy of the matrix (NVLS multicast)
__device__ __forceinline__ void multimem_st_f4(float* p, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
// waits until *p >= want (wrap-around safe for 2^31 exchanges)
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t want) {
    for (uint32_t spin = 0; (int32_t)(ld_acquire_sys(p) - want) < 0; ++spin) {
        if (spin > (1u << 27)) __trap();
        if (spin > 64) __nanosleep(64);
    }
}

template <bool MC, bool SEL>
__global__ void __launch_bounds__(EX_THREADS) push_rows_kernel(PushArgs a) {
    __shared__ uint32_t s_epoch;
    __shared__ int s_last;
    const int tid = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();                      // the block to send is complete; every earlier reader of the local matrix is done
    if (tid == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(a.local_state) + 1u;
    __syncthreads();
    const uint32_t e = s_epoch;
    uint32_t* my_flags = a.flags[a.rank];
    // ---- enter: one CTA announces this rank to every peer ---------------------------------------------------------------
    if (blockIdx.x == 0 && tid < a.world && tid != a.rank) st_release_sys(a.flags[tid] + a.rank, e);
    // ---- stores -----------------------------------------------------------------------------------------------------------
    int peer = -1, chunk = blockIdx.x, n_chunks = gridDim.x;
    if (MC || SEL) {                                                     // this CTA's stores reach every peer: all must have entered
        if (tid < a.world && tid != a.rank) wait_flag(my_flags + tid, e);
    } else {                                                             // CTA (peer slot, chunk)
        const int slot = blockIdx.x / a.ctas_per_peer;
        chunk = blockIdx.x % a.ctas_per_peer;
        n_chunks = a.ctas_per_peer;
        peer = slot + (slot >= a.rank ? 1 : 0);                          // peers in rank order, skipping myself
        if (tid == 0) wait_flag(my_flags + peer, e);
    }
    __syncthreads();
    if (!SEL) {
        const float4* src = reinterpret_cast<const float4*>(a.dst[a.rank] + a.elem0);
        float4* dst = reinterpret_cast<float4*>((MC ? a.mc : a.dst[peer]) + a.elem0);
        const int64_t n4 = a.n_elem >> 2;
        const int64_t per = (n4 + n_chunks - 1) / n_chunks;
        const int64_t i0 = (int64_t)chunk * per, i1 = min(i0 + per, n4);
        int64_t i = i0 + tid;
        constexpr int U = 8;                                             // 16-byte loads in flight per thread
        for (; i + (U - 1) * EX_THREADS < i1; i += U * EX_THREADS) {
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = src[i + u * EX_THREADS];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (MC) multimem_st_f4(reinterpret_cast<float*>(dst + i + u * EX_THREADS), v[u]);
                else dst[i + u * EX_THREADS] = v[u];
            }
        }
        for (; i < i1; i += EX_THREADS) {
            if (MC) multimem_st_f4(reinterpret_cast<float*>(dst + i), src[i]); else dst[i] = src[i];
        }
    } else {
        // one warp per listed row: rows this rank owns go to every peer (the few thousand batch rows of the last layer)
        const int lane = tid & 31;
        const int64_t w = ((int64_t)blockIdx.x * EX_THREADS + tid) >> 5, n_w = ((int64_t)gridDim.x * EX_THREADS) >> 5;
        int64_t total = 0;
        for (int q = 0; q < a.n_lists; ++q) total += a.list_n[q];
        for (int64_t j = w; j < total; j += n_w) {
            int64_t jj = j;
            int q = 0;
            while (jj >= a.list_n[q]) { jj -= a.list_n[q]; ++q; }
            const int64_t row = a.lists[q][jj] + a.list_off[q];
            if (row < a.row0 || row >= a.row0 + a.n_rows) continue;
            const float4* src = reinterpret_cast<const float4*>(a.dst[a.rank] + row * a.d);
            for (int c = lane; c < a.d / 4; c += 32) {
                const float4 v = src[c];
                if (MC) {
                    multimem_st_f4(a.mc + row * a.d + c * 4, v);
                } else {
                    for (int p = 0; p < a.world; ++p)
                        if (p != a.rank) reinterpret_cast<float4*>(a.dst[p] + row * a.d)[c] = v;
                }
            }
        }
    }
    // ---- done: the last CTA to finish its stores tells every peer, then waits for all of them --------------------------------
    __threadfence_system();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.local_state + 1, 1u) + 1u == gridDim.x);
    __syncthreads();
    if (!s_last) return;
    __threadfence_system();
    if (tid < a.world && tid != a.rank) {
        st_release_sys(a.flags[tid] + EX_MAX_WORLD + a.rank, e);
        wait_flag(my_flags + EX_MAX_WORLD + tid, e);
    }
    __syncthreads();
    if (tid == 0) {
        a.local_state[1] = 0;
        *reinterpret_cast<volatile uint32_t*>(a.local_state) = e;
        __threadfence();
    }
}

int fill_common(PushArgs& a, float* const* matrix_on_rank_host, uint32_t* const* flags_on_rank_host, uint32_t* local_state,
                int world, int rank, float* multicast_or_null, const char* who) {
    NGCF_REQUIRE(matrix_on_rank_host && flags_on_rank_host && local_state, "%s: null pointer", who);
    NGCF_REQUIRE(world >= 2 && world <= EX_MAX_WORLD && rank >= 0 && rank < world, "%s: world %d rank %d", who, world, rank);
    for (int r = 0; r < world; ++r) {
        NGCF_REQUIRE(matrix_on_rank_host[r] && flags_on_rank_host[r], "%s: null pointer for rank %d", who, r);
        NGCF_REQUIRE((reinterpret_cast<uintptr_t>(matrix_on_rank_host[r]) & 15) == 0, "%s: matrix of rank %d not 16-byte aligned", who, r);
        a.dst[r] = matrix_on_rank_host[r];
        a.flags[r] = flags_on_rank_host[r];
    }
    a.local_state = local_state;
    a.world = world; a.rank = rank;
    a.mc = multicast_or_null;
    return NGCF_OK;
}

}  // namespace

extern "C" int ngcf_exchange_flag_words(void) { return 2 * EX_MAX_WORLD; }

extern "C" int ngcf_push_rows(float* const* matrix_on_rank_host, uint32_t* const* flags_on_rank_host, uint32_t* local_state,
                              int world, int rank, int64_t row0, int64_t n_rows, int d, float* multicast_or_null, void* stream) {
    NGCF_REQUIRE(row0 >= 0 && n_rows >= 0 && d > 0 && d % 4 == 0, "push_rows: rows [%lld, +%lld) x %d (width must be a multiple of 4)",
                 (long long)row0, (long long)n_rows, d);
    PushArgs a{};
    int rc = fill_common(a, matrix_on_rank_host, flags_on_rank_host, local_state, world, rank, multicast_or_null, "push_rows");
    if (rc != NGCF_OK) return rc;
    a.elem0 = row0 * d; a.n_elem = n_rows * d;
    const int64_t bytes = a.n_elem * 4;
    if (a.mc) {
        // one store per element reaches every peer: as many CTAs as keep the link busy, at most one per SM
        const int grid = (int)min((int64_t)2 * ngcf_num_sms(), max((int64_t)1, bytes / (64 * 1024)));
        a.ctas_per_peer = grid;
        NGCF_CUDA(ngcf_launch_pdl(push_rows_kernel<true, false>, dim3((unsigned)grid), dim3(EX_THREADS), 0, as_stream(stream), a));
    } else {
        int per_peer = (int)min((int64_t)2 * ngcf_num_sms() / (world - 1), max((int64_t)1, bytes / (64 * 1024)));
        if (per_peer < 1) per_peer = 1;
        a.ctas_per_peer = per_peer;
        NGCF_CUDA(ngcf_launch_pdl(push_rows_kernel<false, false>, dim3((unsigned)(per_peer * (world - 1))), dim3(EX_THREADS), 0,
                                  as_stream(stream), a));
    }
    NGCF_LAUNCH_OK("push_rows_kernel");
    return NGCF_OK;
}

extern "C" int ngcf_push_selected_rows(float* const* matrix_on_rank_host, uint32_t* const* flags_on_rank_host, uint32_t* local_state,
                                       int world, int rank, int64_t row0, int64_t n_rows, int d,
                                       const int64_t* const* lists_host, const int64_t* list_offsets_host,
                                       const int64_t* list_sizes_host, int n_lists, float* multicast_or_null, void* stream) {
    NGCF_REQUIRE(row0 >= 0 && n_rows >= 0 && d > 0 && d % 4 == 0, "push_selected_rows: bad block / width %d", d);
    NGCF_REQUIRE(n_lists >= 1 && n_lists <= 4 && lists_host && list_offsets_host && list_sizes_host, "push_selected_rows: 1..4 lists");
    PushArgs a{};
    int rc = fill_common(a, matrix_on_rank_host, flags_on_rank_host, local_state, world, rank, multicast_or_null, "push_selected_rows");
    if (rc != NGCF_OK) return rc;
    int64_t total = 0;
    for (int q = 0; q < n_lists; ++q) {
        NGCF_REQUIRE(lists_host[q] && list_sizes_host[q] >= 0, "push_selected_rows: list %d", q);
        a.lists[q] = lists_host[q]; a.list_off[q] = list_offsets_host[q]; a.list_n[q] = list_sizes_host[q];
        total += list_sizes_host[q];
    }
    a.n_lists = n_lists; a.row0 = row0; a.n_rows = n_rows; a.d = d;
    const int grid = (int)max((int64_t)1, min((int64_t)ngcf_num_sms(), ceil_div64(total, EX_THREADS / 32)));
    if (a.mc) NGCF_CUDA(ngcf_launch_pdl(push_rows_kernel<true, true>, dim3((unsigned)grid), dim3(EX_THREADS), 0, as_stream(stream), a));
    else NGCF_CUDA(ngcf_launch_pdl(push_rows_kernel<false, true>, dim3((unsigned)grid), dim3(EX_THREADS), 0, as_stream(stream), a));
    NGCF_LAUNCH_OK("push_rows_kernel(selected)");
    return NGCF_OK;
}
