// Fused multi-tensor Adam step: the reference trains with torch.optim.Adam over
// model.parameters() (main.py:74, experiment.py:55,58), i.e. ~15 tensors of which the two embedding tables hold
// 99.6 % of the elements.  One launch updates every parameter that has a gradient (the feature tables never do,
// NGCF.py:115) and can zero the gradients in the same pass (optimizer.zero_grad, experiment.py:55).
// HBM-bound: 16 B read + 12 B written per element (+ 4 B when the gradient is zeroed).
#include "common.cuh"

namespace {

constexpr int AD_MAX = NGCF_ADAM_MAX_TENSORS;
constexpr int AD_THREADS = 256;
constexpr int AD_CHUNK = AD_THREADS * 4 * 4;        // elements per CTA iteration: 4 float4 per thread

struct AdamArgs {
    float* p[AD_MAX];
    float* g[AD_MAX];
    float* m[AD_MAX];
    float* v[AD_MAX];
    int64_t chunk0[AD_MAX + 1];                     // first chunk of tensor i (prefix sums of ceil(size / AD_CHUNK))
    int64_t size[AD_MAX];
    int n;
    double lr, beta1d, beta2d;                      // bias corrections are formed in double, like torch's Python floats
    float beta1, beta2, omb1, omb2, eps, weight_decay;
    int64_t step;
    const int64_t* step_dev;
    int zero_grads;
};

__device__ __forceinline__ void adam1(float& p, float& g, float& m, float& v, const AdamArgs& a, float step_size,
                                      float inv_sqrt_bc2) {
    float gr = g;
    if (a.weight_decay != 0.f) gr = fmaf(a.weight_decay, p, gr);
    m = fmaf(a.beta1, m, a.omb1 * gr);                                // exp_avg.lerp_(grad, 1 - beta1)
    v = fmaf(a.beta2, v, a.omb2 * gr * gr);                           // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    const float denom = sqrtf(v) * inv_sqrt_bc2 + a.eps;
    p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(AD_THREADS) adam_kernel(AdamArgs a) {
    const int64_t t = a.step + (a.step_dev ? *a.step_dev : 0);        // 1-based step count of THIS update
    const double bc1 = 1.0 - pow(a.beta1d, (double)t), bc2 = 1.0 - pow(a.beta2d, (double)t);
    const float step_size = (float)(a.lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    const int64_t n_chunks = a.chunk0[a.n];
    for (int64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        int i = 0;
        while (i + 1 < a.n && c >= a.chunk0[i + 1]) ++i;              // <= 32 tensors: a linear search is fine
        const int64_t base = (c - a.chunk0[i]) * AD_CHUNK, n = a.size[i];
        float* P = a.p[i] + base; float* G = a.g[i] + base; float* M = a.m[i] + base; float* V = a.v[i] + base;
        const int64_t left = n - base;
        const bool vec = left >= AD_CHUNK && ((reinterpret_cast<uintptr_t>(P) | reinterpret_cast<uintptr_t>(G) |
                                               reinterpret_cast<uintptr_t>(M) | reinterpret_cast<uintptr_t>(V)) & 15) == 0;
        if (vec) {
            float4 p4[4], g4[4], m4[4], v4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int o = (u * AD_THREADS + threadIdx.x) * 4;
                p4[u] = ld_f4(P + o); g4[u] = ld_f4(G + o); m4[u] = ld_f4(M + o); v4[u] = ld_f4(V + o);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int o = (u * AD_THREADS + threadIdx.x) * 4;
                adam1(p4[u].x, g4[u].x, m4[u].x, v4[u].x, a, step_size, inv_sqrt_bc2);
                adam1(p4[u].y, g4[u].y, m4[u].y, v4[u].y, a, step_size, inv_sqrt_bc2);
                adam1(p4[u].z, g4[u].z, m4[u].z, v4[u].z, a, step_size, inv_sqrt_bc2);
                adam1(p4[u].w, g4[u].w, m4[u].w, v4[u].w, a, step_size, inv_sqrt_bc2);
                st_f4(P + o, p4[u]); st_f4(M + o, m4[u]); st_f4(V + o, v4[u]);
                if (a.zero_grads) st_f4(G + o, make_float4(0.f, 0.f, 0.f, 0.f));
            }
        } else {
            for (int64_t o = threadIdx.x; o < left && o < AD_CHUNK; o += AD_THREADS) {
                float p = P[o], g = G[o], m = M[o], v = V[o];
                adam1(p, g, m, v, a, step_size, inv_sqrt_bc2);
                P[o] = p; M[o] = m; V[o] = v;
                if (a.zero_grads) G[o] = 0.f;
            }
        }
    }
}

}  // namespace

extern "C" int ngcf_adam_step(float* const* params_host, float* const* grads_host, float* const* exp_avg_host,
                              float* const* exp_avg_sq_host, const int64_t* sizes_host, int n_tensors, double lr,
                              double beta1, double beta2, double eps, double weight_decay, int64_t step,
                              const int64_t* step_dev, int zero_grads, void* stream) {
    NGCF_REQUIRE(params_host && grads_host && exp_avg_host && exp_avg_sq_host && sizes_host, "adam_step: null host array");
    NGCF_REQUIRE(n_tensors >= 0 && n_tensors <= AD_MAX, "adam_step: %d tensors not in [0,%d]", n_tensors, AD_MAX);
    NGCF_REQUIRE(lr >= 0. && beta1 >= 0. && beta1 < 1. && beta2 >= 0. && beta2 < 1. && eps >= 0.,
                 "adam_step: bad hyper-parameters");
    NGCF_REQUIRE(step >= 1 || step_dev, "adam_step: step counts from 1");
    if (n_tensors == 0) return NGCF_OK;
    AdamArgs a{};
    a.chunk0[0] = 0;
    for (int i = 0; i < n_tensors; ++i) {
        NGCF_REQUIRE(params_host[i] && grads_host[i] && exp_avg_host[i] && exp_avg_sq_host[i] && sizes_host[i] >= 0,
                     "adam_step: tensor %d has a null pointer or a negative size", i);
        a.p[i] = params_host[i]; a.g[i] = grads_host[i]; a.m[i] = exp_avg_host[i]; a.v[i] = exp_avg_sq_host[i];
        a.size[i] = sizes_host[i];
        a.chunk0[i + 1] = a.chunk0[i] + ceil_div64(sizes_host[i], AD_CHUNK);
    }
    a.n = n_tensors;
    a.lr = lr; a.beta1d = beta1; a.beta2d = beta2;
    a.beta1 = (float)beta1; a.beta2 = (float)beta2; a.omb1 = (float)(1.0 - beta1); a.omb2 = (float)(1.0 - beta2);
    a.eps = (float)eps; a.weight_decay = (float)weight_decay;
    a.step = step; a.step_dev = step_dev; a.zero_grads = zero_grads;
    const int64_t n_chunks = a.chunk0[n_tensors];
    if (n_chunks == 0) return NGCF_OK;
    const int grid = (int)min(n_chunks, (int64_t)ngcf_num_sms() * 8);
    adam_kernel<<<grid, AD_THREADS, 0, as_stream(stream)>>>(a);
    NGCF_LAUNCH_OK("adam_kernel");
    return NGCF_OK;
}
