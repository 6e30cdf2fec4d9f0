// Correctness probe for the fp32-class fused chain design (tc_chain32.cu), run before the product kernel was written:
//   1. tcgen05.mma with A in TENSOR MEMORY (packed bf16 pairs, lane = row, 8 columns per 16-wide K step);
//   2. split-bf16 arithmetic with both operands as hi + lo terms: D = A_hi W_hi + A_lo W_hi + A_hi W_lo, fp32 accumulate;
//   3. the epilogue -> next layer hand-over entirely inside TMEM (tcgen05.ld -> bias/ReLU -> split -> tcgen05.st -> MMA);
//   4. channel-wise max over the 128 rows of a NON-transposed tile by redux.sync.max.s32 on float bit patterns.
// One CTA, 128 threads, one 128-row tile:  x[128, 9] -> 64 (ReLU) -> 128 (ReLU) -> 256 (ReLU, max over rows).
// Every stage is compared with a float64 host evaluation of the same fp32 inputs; expected relative error ~1e-5.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I 3d-semantic-segmentation-amp-net_b200/csrc \
//        tools/ts_chain_probe.cu -o tools/build/ts_chain_probe && tools/build/ts_chain_probe
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tc_ts.cuh"

using namespace amp::tcx;

constexpr int M = 128, K0 = 16, N1 = 64, N2 = 128, N3 = 256;
constexpr int kAcc = 0, kAhi = 256, kAlo = 320;          // TMEM columns

// weights in shared memory: hi block then lo block, each K-major no-swizzle [K/8][N][8] bf16
__device__ void stage_weights(unsigned char* dst, const float* w, int N, int K, int Kreal, int tid) {
    for (int e = tid; e < N * K; e += 128) {
        const int n = e / K, k = e - n * K;
        const float v = k < Kreal ? w[n * Kreal + k] : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        const int off = ((k >> 3) * N + n) * 16 + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(dst + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(dst + N * K * 2 + off) = l;
    }
}

// all MMAs of one layer: 3 products per 16-wide K step, A from TMEM
__device__ void issue_layer(uint32_t tm, uint32_t w_addr, int K, int N, int n0, int ncols, uint32_t acc_col, uint32_t bar) {
    const uint32_t idesc = umma_idesc(128, ncols);
    const uint32_t lbo = (uint32_t)N * 16u;
    const uint64_t whi = umma_desc(w_addr + (uint32_t)n0 * 16u, lbo, 128u);
    const uint64_t wlo = umma_desc(w_addr + (uint32_t)(N * K * 2) + (uint32_t)n0 * 16u, lbo, 128u);
    const uint64_t step = (2u * lbo) >> 4;
    for (int ks = 0; ks < K / 16; ++ks) {
        umma_bf16_ts(tm + acc_col, tm + kAhi + ks * 8, whi + ks * step, idesc, ks > 0 ? 1u : 0u);
        umma_bf16_ts(tm + acc_col, tm + kAlo + ks * 8, whi + ks * step, idesc, 1u);
        umma_bf16_ts(tm + acc_col, tm + kAhi + ks * 8, wlo + ks * step, idesc, 1u);
    }
    umma_commit(bar);
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const float* __restrict__ x, const float* __restrict__ w1, const float* __restrict__ b1,
                                                        const float* __restrict__ w2, const float* __restrict__ b2,
                                                        const float* __restrict__ w3, const float* __restrict__ b3,
                                                        float* __restrict__ y1, float* __restrict__ y2, float* __restrict__ pool,
                                                        long long* __restrict__ cycles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* s_w1 = smem;                                   // 64 x 16 x 2 x 2      =   4 KB
    unsigned char* s_w2 = s_w1 + N1 * K0 * 4;                     // 128 x 64 x 2 x 2     =  32 KB
    unsigned char* s_w3 = s_w2 + N2 * N1 * 4;                     // 256 x 128 x 2 x 2    = 128 KB
    float* s_b = reinterpret_cast<float*>(s_w3 + N3 * N2 * 4);    // 64 + 128 + 256 floats
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    __shared__ unsigned int s_pool[N3];
    const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
    stage_weights(s_w1, w1, N1, K0, 9, tid);
    stage_weights(s_w2, w2, N2, N1, N1, tid);
    stage_weights(s_w3, w3, N3, N2, N2, tid);
    for (int i = tid; i < N1; i += 128) s_b[i] = b1[i];
    for (int i = tid; i < N2; i += 128) s_b[N1 + i] = b2[i];
    for (int i = tid; i < N3; i += 128) { s_b[N1 + N2 + i] = b3[i]; s_pool[i] = 0u; }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(smem_u32(&s_tmem), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = uniform_u32(s_tmem);
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    const uint32_t barrier = smem_u32(&bar);
    uint32_t phase = 0;
    const long long t0 = clock64();
    // ---- input stage: this thread's row, 9 columns padded to K0 = 16, hi / lo pairs straight into TMEM ----
    {
        float xv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) xv[j] = j < 9 ? x[tid * 9 + j] : 0.f;
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) split_pair(xv[2 * q], xv[2 * q + 1], hi[q], lo[q]);
        tmem_st8(tm + lane_addr + kAhi, hi);
        tmem_st8(tm + lane_addr + kAlo, lo);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    // ---- layers 1 and 2: accumulator -> bias + ReLU -> split -> A operand of the next layer ----
    for (int l = 0; l < 2; ++l) {
        const int K = l == 0 ? K0 : N1, N = l == 0 ? N1 : N2;
        const float* bias = l == 0 ? s_b : s_b + N1;
        float* yout = l == 0 ? y1 : y2;
        if (warp == 0) {
            tc_fence_after();
            if (elect_one_sync()) issue_layer(tm, smem_u32(l == 0 ? s_w1 : s_w2), K, N, 0, N, kAcc, barrier);
            __syncwarp();
        }
        mbar_wait_bounded(barrier, phase);
        phase ^= 1u;
        tc_fence_after();
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tm + lane_addr + kAcc + c0, v);
            tmem_wait_ld();
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const float a = fmaxf(__uint_as_float(v[2 * q]) + bias[c0 + 2 * q], 0.f);
                const float b = fmaxf(__uint_as_float(v[2 * q + 1]) + bias[c0 + 2 * q + 1], 0.f);
                yout[tid * N + c0 + 2 * q] = a;
                yout[tid * N + c0 + 2 * q + 1] = b;
                split_pair(a, b, hi[q], lo[q]);
            }
            tmem_st16(tm + lane_addr + kAhi + c0 / 2, hi);
            tmem_st16(tm + lane_addr + kAlo + c0 / 2, lo);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncthreads();
    }
    // ---- layer 3: 256 channels as two 128-column halves, max over the 128 rows by redux ----
    for (int h = 0; h < 2; ++h) {
        if (warp == 0) {
            tc_fence_after();
            if (elect_one_sync()) issue_layer(tm, smem_u32(s_w3), N2, N3, h * 128, 128, kAcc, barrier);
            __syncwarp();
        }
        mbar_wait_bounded(barrier, phase);
        phase ^= 1u;
        tc_fence_after();
        const float* bias = s_b + N1 + N2 + h * 128;
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tm + lane_addr + kAcc + c0, v);
            tmem_wait_ld();
            float mine = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float m = warp_max_relu_safe(__uint_as_float(v[j]) + bias[c0 + j]);
                if (lane == j) mine = m;
            }
            atomicMax(&s_pool[h * 128 + c0 + lane], __float_as_uint(fmaxf(mine, 0.f)));
        }
        tc_fence_before();
        __syncthreads();
    }
    const long long t1 = clock64();
    for (int i = tid; i < N3; i += 128) pool[i] = __uint_as_float(s_pool[i]);
    if (tid == 0) cycles[0] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tm, 512);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static double relerr(const std::vector<float>& got, const std::vector<double>& ref) {
    double e = 0.0, m = 0.0;
    for (size_t i = 0; i < ref.size(); ++i) { e = fmax(e, fabs((double)got[i] - ref[i])); m = fmax(m, fabs(ref[i])); }
    return e / fmax(m, 1e-30);
}

int main() {
    srand(11);
    auto rnd = [](float s) { return ((float)rand() / RAND_MAX * 2.f - 1.f) * s; };
    std::vector<float> x(M * 9), w1(N1 * 9), b1(N1), w2(N2 * N1), b2(N2), w3(N3 * N2), b3(N3);
    for (auto& v : x) v = rnd(1.f);
    for (auto& v : w1) v = rnd(0.5f);
    for (auto& v : b1) v = rnd(0.2f);
    for (auto& v : w2) v = rnd(0.2f);
    for (auto& v : b2) v = rnd(0.2f);
    for (auto& v : w3) v = rnd(0.15f);
    for (auto& v : b3) v = rnd(0.2f);
    std::vector<double> r1(M * N1), r2(M * N2), rp(N3, 0.0);
    for (int r = 0; r < M; ++r) {
        for (int n = 0; n < N1; ++n) {
            double s = b1[n];
            for (int k = 0; k < 9; ++k) s += (double)x[r * 9 + k] * w1[n * 9 + k];
            r1[r * N1 + n] = s > 0 ? s : 0;
        }
        for (int n = 0; n < N2; ++n) {
            double s = b2[n];
            for (int k = 0; k < N1; ++k) s += r1[r * N1 + k] * w2[n * N1 + k];
            r2[r * N2 + n] = s > 0 ? s : 0;
        }
        for (int n = 0; n < N3; ++n) {
            double s = b3[n];
            for (int k = 0; k < N2; ++k) s += r2[r * N2 + k] * w3[n * N2 + k];
            if (s > rp[n]) rp[n] = s;
        }
    }
    float *dx, *dw1, *db1, *dw2, *db2, *dw3, *db3, *dy1, *dy2, *dp;
    long long* dc;
    CK(cudaMalloc(&dx, x.size() * 4)); CK(cudaMalloc(&dw1, w1.size() * 4)); CK(cudaMalloc(&db1, b1.size() * 4));
    CK(cudaMalloc(&dw2, w2.size() * 4)); CK(cudaMalloc(&db2, b2.size() * 4)); CK(cudaMalloc(&dw3, w3.size() * 4));
    CK(cudaMalloc(&db3, b3.size() * 4)); CK(cudaMalloc(&dy1, r1.size() * 4)); CK(cudaMalloc(&dy2, r2.size() * 4));
    CK(cudaMalloc(&dp, N3 * 4)); CK(cudaMalloc(&dc, 8));
    CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw1, w1.data(), w1.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db1, b1.data(), b1.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw2, w2.data(), w2.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db2, b2.data(), b2.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw3, w3.data(), w3.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db3, b3.data(), b3.size() * 4, cudaMemcpyHostToDevice));
    const int smem = N1 * K0 * 4 + N2 * N1 * 4 + N3 * N2 * 4 + (N1 + N2 + N3) * 4;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    probe_kernel<<<1, 128, smem>>>(dx, dw1, db1, dw2, db2, dw3, db3, dy1, dy2, dp, dc);
    CK(cudaDeviceSynchronize());
    std::vector<float> y1(r1.size()), y2(r2.size()), pool(N3);
    long long cyc = 0;
    CK(cudaMemcpy(y1.data(), dy1, y1.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(y2.data(), dy2, y2.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(pool.data(), dp, N3 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
    const double e1 = relerr(y1, r1), e2 = relerr(y2, r2), e3 = relerr(pool, rp);
    printf("fp32-class TS chain, one 128-row tile: 9 -> 64 -> 128 -> 256 (max over rows); %lld cycles\n", cyc);
    printf("layer 1 rel err %.3e\nlayer 2 rel err %.3e\npooled  rel err %.3e\n", e1, e2, e3);
    const bool ok = e1 < 1e-4 && e2 < 1e-4 && e3 < 1e-4;
    printf("%s\n", ok ? "OK: TS operand layout, split arithmetic, TMEM hand-over and redux pooling are right" : "MISMATCH");
    return ok ? 0 : 2;
}
