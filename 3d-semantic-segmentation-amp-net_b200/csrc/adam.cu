// Adam over a list of tensors in ONE launch: the two optimizers of train_pointnet-attention.py:141-142 (torch.optim.Adam,
// default betas / eps, no weight decay, no amsgrad) step ~110 small tensors, 1.2 M parameters in total. torch's fused
// multi-tensor implementation needs three launches of 20-50 CTAs (65 536-element chunks, 36 tensors per launch): 135 us per
// step on B200 for 34 MB of traffic. Here the host flattens the tensors into chunks of 2048 elements, one CTA per chunk
// (~ 700 CTAs), 16-byte accesses when the four pointers allow it.
//   m = m + (1 - b1) (g - m);  v = b2 v + (1 - b2) g g;  p -= (lr / (1 - b1^t)) m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// (the operation order of torch's FusedAdamMathFunctor, fp32). t = steps[tensor] + 1 is read from device memory (a captured
// CUDA graph replays the launch with the step count of the replay, not of the capture); the caller increments the counters
// of the tensors it stepped afterwards.
#include <math.h>

#include "amp_common.cuh"

namespace amp {
namespace {

struct AdamChunk { float* p; const float* g; float* m; float* v; int n; int tensor; };   // tensor: index into the step counters
static_assert(sizeof(AdamChunk) == 40, "the host builds this table with ctypes / struct");
constexpr int kAdamChunk = 2048, kAdamThreads = 256;

__global__ void __launch_bounds__(kAdamThreads) adam_kernel(const AdamChunk* __restrict__ table, const long long* __restrict__ steps,
                                                            float lr, float b1, float b2, float eps) {
    pdl_sync();
    const AdamChunk c = table[blockIdx.x];
    const double t = (double)(steps[c.tensor] + 1);             // torch counts steps per parameter: one without a gradient is skipped
    const float bc1 = (float)(1.0 - pow((double)b1, t)), bc2s = (float)sqrt(1.0 - pow((double)b2, t));
    const float step_size = lr / bc1;
    const bool vec = (((uintptr_t)c.p | (uintptr_t)c.g | (uintptr_t)c.m | (uintptr_t)c.v) & 15) == 0;
    auto upd = [&](float& p, float g, float& m, float& v) {
        m = m + (1.f - b1) * (g - m);
        v = b2 * v + (1.f - b2) * g * g;
        p -= step_size * m / (sqrtf(v) / bc2s + eps);
    };
    if (vec) {
        const int n4 = c.n >> 2;
        for (int i = threadIdx.x; i < n4; i += kAdamThreads) {
            float4 p = reinterpret_cast<float4*>(c.p)[i], m = reinterpret_cast<float4*>(c.m)[i], v = reinterpret_cast<float4*>(c.v)[i];
            const float4 g = reinterpret_cast<const float4*>(c.g)[i];
            upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
            reinterpret_cast<float4*>(c.p)[i] = p; reinterpret_cast<float4*>(c.m)[i] = m; reinterpret_cast<float4*>(c.v)[i] = v;
        }
        for (int i = (n4 << 2) + threadIdx.x; i < c.n; i += kAdamThreads) upd(c.p[i], c.g[i], c.m[i], c.v[i]);
    } else {
        for (int i = threadIdx.x; i < c.n; i += kAdamThreads) upd(c.p[i], c.g[i], c.m[i], c.v[i]);
    }
}

}  // namespace
}  // namespace amp

extern "C" {

int32_t amp_adam_chunk_elems(void) { return amp::kAdamChunk; }

int amp_adam_step(const void* chunk_table, int64_t n_chunks, const int64_t* steps, float lr, float beta1, float beta2, float eps,
                  void* stream) {
    using namespace amp;
    const int64_t* step = steps;
    if (!chunk_table || !step || n_chunks < 1) return fail(AMP_E_BADARG, "adam_step: null table / step or no chunks");
    if (!(lr >= 0.f) || !(beta1 >= 0.f && beta1 < 1.f) || !(beta2 >= 0.f && beta2 < 1.f) || !(eps >= 0.f))
        return fail(AMP_E_BADARG, "adam_step: bad hyper-parameters");
    launch_pdl(adam_kernel, dim3((unsigned)n_chunks), dim3(kAdamThreads), 0, (cudaStream_t)stream,
               reinterpret_cast<const AdamChunk*>(chunk_table), reinterpret_cast<const long long*>(step), lr, beta1, beta2, eps);
    count_launch();
    return check_launch("adam_kernel");
}

}  // extern "C"
