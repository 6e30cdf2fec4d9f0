// How long does the MMA batch of ONE tc_layer chunk take in isolation? 12 tcgen05.mma (4 K steps x 3 split products) with
// M = 128, N = 128, K = 16, A in tensor memory (TS) or shared memory (SS), B = DISTINCT shared-memory tiles per K step
// (hi and lo, as the layer kernel has them), one commit, one waiter. tools/umma_bench.cu reuses one B tile for all MMAs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I 3d-semantic-segmentation-amp-net_b200/csrc tools/umma_batch_probe.cu -o tools/_build/umma_batch_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <vector>
#include <cuda_bf16.h>
#include "tc_ptx.cuh"
#include "tc_ts.cuh"
using namespace amp::tcx;
constexpr int REPS = 20;

// mode bit 0: A from TMEM; bit 1: B tiles distinct per K step; bit 2: other 3 warps write shared memory while the MMAs run
template <int KSTEPS>
__global__ void __launch_bounds__(128, 1) batch_kernel(int mode, long long* out, int n_cols = 128) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* s_a = smem;                       // 2 x 4 KB (hi, lo) per K step: 8 x 8 KB = 64 KB max
    unsigned char* s_b = smem + 65536;               // same
    unsigned char* s_scratch = smem + 131072;        // 32 KB written by the other warps
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = warp_index_uniform();
    // mode bit 3: pseudo-random bf16 operands (finite, |x| < 2) instead of zeros
    for (int i = tid; i < 163840 / 4; i += 128) {
        uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<uint32_t*>(smem)[i] = (mode & 8) ? ((h & 0x807f807fu) | 0x3f003f00u) : 0u;
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(smem_u32(&s_tmem), 512);
    fence_proxy_async(); tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tm = uniform_u32(s_tmem);
    { uint32_t z[16]; for (int i = 0; i < 16; ++i) z[i] = (mode & 8) ? (0x3f803f00u + (uint32_t)(tid * 131 + i) % 127u) : 0u;
      for (int c = 0; c < 256; c += 16) tmem_st16(tm + ((uint32_t)(warp * 32) << 16) + 256u + c, z);
      tmem_wait_st(); }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t idesc = umma_idesc(128, n_cols), barrier = smem_u32(&bar);
    const uint32_t lbo = (uint32_t)n_cols * 16u;       // B tile [K / 8][N][8] bf16: K-group stride
    uint32_t phase = 0;
    for (int rep = 0; rep < REPS; ++rep) {
        long long t0 = 0; bool elected = false;
        if (warp == 0) {
            tc_fence_after();
            elected = elect_one_sync();
            if (elected) {
                t0 = clock64();
                for (int ks = 0; ks < KSTEPS; ++ks) {
                    const uint32_t off = (mode & 2) ? (uint32_t)ks * 8192u : 0u;
                    const uint64_t b_hi = umma_desc(smem_u32(s_b) + off, lbo, 128u), b_lo = umma_desc(smem_u32(s_b) + off + 2u * lbo, lbo, 128u);
                    if (mode & 1) {
                        const uint32_t a_hi = tm + 256u + ks * 8u, a_lo = tm + 384u + ks * 8u;
                        umma_bf16_ts(tm, a_lo, b_hi, idesc, ks != 0); umma_bf16_ts(tm, a_hi, b_lo, idesc, 1u); umma_bf16_ts(tm, a_hi, b_hi, idesc, 1u);
                    } else {
                        const uint64_t a_hi = umma_desc(smem_u32(s_a) + off, 2048u, 128u), a_lo = umma_desc(smem_u32(s_a) + off + 4096u, 2048u, 128u);
                        umma_bf16(tm, a_lo, b_hi, idesc, ks != 0); umma_bf16(tm, a_hi, b_lo, idesc, 1u); umma_bf16(tm, a_hi, b_hi, idesc, 1u);
                    }
                }
                umma_commit(barrier);
            }
            __syncwarp();
            if (elected) { mbar_wait(barrier, phase); if (blockIdx.x == 0) out[rep] = clock64() - t0; }
            __syncwarp();
        } else if (mode & 4) {
            for (int it = 0; it < 8; ++it)
                for (int i = tid - 32; i < 32768 / 16; i += 96) reinterpret_cast<uint4*>(s_scratch)[i] = make_uint4(it, i, 0, 0);
        }
        phase ^= 1u;
        __syncthreads();
        tc_fence_after();
    }
    tc_fence_before(); __syncthreads();
    if (warp == 1) tmem_dealloc(tm, 512);
}

static int g_grid = 1;
template <int KSTEPS>
static double run(int mode, long long* d, int n_cols = 128) {
    std::vector<long long> h(REPS);
    cudaFuncSetAttribute(batch_kernel<KSTEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 163840);
    batch_kernel<KSTEPS><<<g_grid, 128, 163840>>>(mode, d, n_cols);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    cudaMemcpy(h.data(), d, REPS * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin() + 2, h.end());
    return (double)h[2 + (REPS - 2) / 2];
}
int main(int argc, char** argv) {
    if (argc > 1) g_grid = atoi(argv[1]);          // CTAs launched (one per SM): does a busy chip change the picture?
    printf("grid = %d CTAs\n", g_grid);
    long long* d; cudaMalloc(&d, REPS * sizeof(long long));
    printf("cycles from first issue to mbarrier observed (one CTA, M = N = 128, K = 16 per MMA, 3 MMAs per K step)\n");
    printf("%-44s %10s %10s %10s\n", "mode", "1 K step", "4 K steps", "8 K steps");
    const char* names[8] = {"SS, one B tile", "TS, one B tile", "SS, distinct tiles", "TS, distinct tiles", "SS, one tile + smem writers", "TS, one tile + smem writers",
                            "SS, distinct + smem writers", "TS, distinct + smem writers"};
    for (int mode = 0; mode < 8; ++mode) printf("%-44s %10.0f %10.0f %10.0f\n", names[mode], run<1>(mode, d), run<4>(mode, d), run<8>(mode, d));
    printf("the same with pseudo-random non-zero operands\n");
    for (int mode = 8; mode < 12; ++mode) printf("%-44s %10.0f %10.0f %10.0f\n", names[mode - 8], run<1>(mode, d), run<4>(mode, d), run<8>(mode, d));
    // the width of the output block: does a narrow MMA cost less? (TS and SS, one B tile so that N = 256 fits)
    printf("cycles per batch against N (M = 128, K = 16); per-MMA cost = (8 K steps - 4 K steps) / 12\n");
    printf("%-44s %10s %10s %10s %10s\n", "N", "4 K steps", "8 K steps", "TS / MMA", "SS / MMA");
    for (int n : {16, 32, 64, 128, 256}) {
        const double t4 = run<4>(1, d, n), t8 = run<8>(1, d, n), s4 = run<4>(0, d, n), s8 = run<8>(0, d, n);
        printf("%-44d %10.0f %10.0f %10.1f %10.1f\n", n, t4, t8, (t8 - t4) / 12.0, (s8 - s4) / 12.0);
    }
    return 0;
}
