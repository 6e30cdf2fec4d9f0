// Device-side batch assembly and augmentation of the training loop.
//
// Replaces, for one collated batch, the host work of train_pointnet-attention.py:390-408: shuffle_clusters (utils/utils.py:620-632),
// the per-window numpy rotate_point_cloud_z (:582-604) + shuffle_data (:607-617) and the W separate `.to(device)` copies.
// ONE launch turns the collate layout  pc [B, N, 9, W] float32 / targets [B, N, W] int64  (already on the device: a single
// H2D copy per step) into the window-major tensors the encoder loop consumes:
//     x [W, B, N, 9] float32      x[w] = rotate_z(pc[:, pperm[w], :, cperm[w]])      (contiguous [B, N, 9] per window)
//     t [B, W * N]  int64         t[b, w * N + i] = targets[b, pperm[w][i], cperm[w]]  (train_...:421 concatenation)
// The permutations and the angle are drawn on the host with numpy in the reference's order (assembly.py), so a seeded run
// sees the same tensors. Rotation arithmetic = numpy's: float32 coordinates times the float64 matrix
// [[c, s, 0], [-s, c, 0], [0, 0, 1]], products and sum in float64 without FMA, rounded to float32 on store.
#include <stdint.h>

#include "amp_common.cuh"

namespace amp {
namespace {

__global__ void __launch_bounds__(256) assemble_windows_kernel(const float* __restrict__ pc, const long long* __restrict__ targets,
                                                               const int* __restrict__ cperm, const int* __restrict__ pperm, int B, int N, int D,
                                                               int W, int rotate, double c, double s, float* __restrict__ x,
                                                               long long* __restrict__ t) {
    const long long total = (long long)W * B * N;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(e % N);
        const int b = (int)((e / N) % B);
        const int w = (int)(e / ((long long)N * B));
        const int sw = cperm[w], sr = pperm[(long long)w * N + i];
        const float* src = pc + (((long long)b * N + sr) * D) * W + sw;
        float* dst = x + e * D;
        float v[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) v[d] = src[(long long)d * W];
        if (rotate) {
            const double px = (double)v[0], py = (double)v[1];
            dst[0] = __double2float_rn(__dadd_rn(__dmul_rn(px, c), __dmul_rn(py, -s)));
            dst[1] = __double2float_rn(__dadd_rn(__dmul_rn(px, s), __dmul_rn(py, c)));
            dst[2] = v[2];
        } else {
            dst[0] = v[0]; dst[1] = v[1]; dst[2] = v[2];
        }
        for (int d = 3; d < D; ++d) dst[d] = src[(long long)d * W];
        if (t) t[(long long)b * W * N + (long long)w * N + i] = targets[((long long)b * N + sr) * W + sw];
    }
}

}  // namespace
}  // namespace amp

extern "C" {

int amp_assemble_windows_f32(const float* pc, const int64_t* targets, const int32_t* cluster_perm, const int32_t* point_perm, int64_t B,
                             int64_t N, int32_t D, int32_t W, int32_t rotate, double cos_a, double sin_a, float* x, int64_t* targets_out,
                             void* stream) {
    using namespace amp;
    if (!pc || !cluster_perm || !point_perm || !x) return fail(AMP_E_BADARG, "assemble_windows: null pointer");
    if ((targets == nullptr) != (targets_out == nullptr)) return fail(AMP_E_BADARG, "assemble_windows: targets and targets_out come together");
    if (B < 1 || N < 1 || D < 3 || W < 1 || B * N * (int64_t)W >= (1LL << 40)) return fail(AMP_E_BADARG, "assemble_windows: bad shape");
    const long long total = (long long)W * B * N;
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    assemble_windows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pc, reinterpret_cast<const long long*>(targets), cluster_perm,
                                                                              point_perm, (int)B, (int)N, D, W, rotate, cos_a, sin_a, x,
                                                                              reinterpret_cast<long long*>(targets_out));
    count_launch();
    return check_launch("assemble_windows");
}

}  // extern "C"
