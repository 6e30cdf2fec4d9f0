// Micro-benchmark behind DESIGN.md section 7: what does ONE tcgen05.mma of the shapes the fused chains use cost, and how
// fast can an epilogue read / write TMEM?  Result on B200: profiles/r01_umma_bench.txt.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I 3d-semantic-segmentation-amp-net_b200/csrc \
//        tools/umma_bench.cu -o tools/build/umma_bench && tools/build/umma_bench
//
// Measurements (clock64 of one CTA on one SM, medians over repeats):
//   1. SS mode  (A = 128-row activation tile in shared memory, B = N x 16 weights in shared memory), M = 128,
//      N in {16, 64, 128, 256}: cycles per MMA of a batch of CHAIN MMAs + commit (first issue -> mbarrier observed).
//      Floor = max(M, 128) * N / 256 cycles.
//   2. TS mode  (A = the same tile held in TMEM as packed bf16 pairs, 8 columns per K step): same batch.
//   3. SS mode with 2 / 4 independent accumulators (round robin): does the accumulate dependency cost anything? (No.)
//   4. tcgen05.ld 32x32b.x32 / .x8 and tcgen05.st 32x32b.x32: bytes per cycle per SM with 4, 8 and 16 warps active.
// Results are printed as a table; nothing is checked numerically (operands are zeros).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include <cuda_bf16.h>

#include "tc_ptx.cuh"

using namespace amp::tcx;

constexpr int CHAIN = 32, REPS = 20;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]),
        "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
        "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// MODE 0: SS, 1: TS (A from TMEM columns [256, 264)). NACC > 1: round-robin over independent accumulators (SS only), so
// consecutive MMAs do not depend on each other. One CTA of 128 threads. The MMAs are issued the way the product kernels do
// it -- the whole first warp enters (warp-uniform branch), one elected lane issues, every operand is warp-uniform or a
// compile-time constant -- because an `if (threadIdx.x == 0)` around the loop costs an elect / R2UR broadcast loop of
// ~100-250 cycles PER MMA and would be what is measured (the first version of this file did exactly that).
template <int MODE, int N, int NACC>
__global__ void __launch_bounds__(128, 1) mma_chain_kernel(long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    // A: 128 rows x 16 k bf16, K-major no-swizzle: 2 K groups x 128 rows x 16 B = 4 KB; B: N rows x 16 k: N * 32 B
    unsigned char* s_a = smem;
    unsigned char* s_b = smem + 4096;
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = warp_index_uniform();
    for (int i = tid; i < (4096 + 256 * 32) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(smem_u32(&s_tmem), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = uniform_u32(s_tmem);
    {   // TS mode: the A operand lives in TMEM columns [256, 288): zero it (each thread = one lane)
        uint32_t z[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = 0u;
        tmem_st32(tm + ((uint32_t)(warp * 32) << 16) + 256u, z);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = umma_idesc(128, N);
    const uint64_t a_d = umma_desc(smem_u32(s_a), 2048u, 128u), b_d = umma_desc(smem_u32(s_b), (uint32_t)N * 16u, 128u);
    const uint32_t barrier = smem_u32(&bar);
    uint32_t phase = 0;
    for (int rep = 0; rep < REPS; ++rep) {
        long long t0 = 0;
        bool elected = false;
        if (warp == 0) {
            tc_fence_after();
            elected = elect_one_sync();
            if (elected) {
                t0 = clock64();
#pragma unroll
                for (int i = 0; i < CHAIN; ++i) {
                    if (MODE == 0) umma_bf16(tm + (uint32_t)((i % NACC) * N), a_d, b_d, idesc, i >= NACC ? 1u : 0u);
                    else umma_bf16_ts(tm, tm + 256u, b_d, idesc, i > 0 ? 1u : 0u);
                }
                umma_commit(barrier);
            }
            __syncwarp();
        }
        mbar_wait(barrier, phase);
        phase ^= 1u;
        tc_fence_after();
        if (elected) out[rep] = clock64() - t0;
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tm, 512);
}

template <int MODE, int N, int NACC>
static double run_chain(long long* d_out) {
    std::vector<long long> h(REPS);
    if (cudaFuncSetAttribute(mma_chain_kernel<MODE, N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384) != cudaSuccess) return -1.0;
    mma_chain_kernel<MODE, N, NACC><<<1, 128, 16384>>>(d_out);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1.0;
    cudaMemcpy(h.data(), d_out, REPS * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin() + 2, h.end());
    return (double)h[2 + (REPS - 2) / 2] / CHAIN;
}
template <int N>
static void chain_row(long long* d_out) {
    const double ss = run_chain<0, N, 1>(d_out), ts = run_chain<1, N, 1>(d_out), s2 = run_chain<0, N, 2>(d_out);
    const double s4 = N <= 128 ? run_chain<0, (N <= 128 ? N : 128), 4>(d_out) : -1.0;
    printf("%6d %10.1f %10.1f %10.1f %12.1f %12.1f\n", N, 128.0 * N / 256.0, ss, ts, s2, s4);
}

// kind 0: ld x32, 1: ld x8 (4 per 32 columns), 2: st x32. `warps` warps active (each warp touches its own 32-lane quarter:
// warp % 4), every warp moves ITER * 32 columns.
__global__ void __launch_bounds__(512, 1) tmem_rw_kernel(int kind, int warps, long long* out, uint32_t* sink) {
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 1) tmem_alloc(smem_u32(&s_tmem), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
    constexpr int ITER = 64;
    uint32_t acc = 0;
    for (int rep = 0; rep < REPS; ++rep) {
        __syncthreads();
        const long long t0 = clock64();
        if (warp < warps) {
            for (int it = 0; it < ITER; ++it) {
                const uint32_t col = (uint32_t)((it * 32 + (warp >> 2) * 128) & 511);
                if (kind == 0) {
                    uint32_t v[32];
                    tmem_ld32(tm + col, v);
                    tmem_wait_ld();
                    acc ^= v[0] ^ v[31];
                } else if (kind == 1) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint32_t v[8];
                        tmem_ld8(tm + col + q * 8, v);
                        tmem_wait_ld();
                        acc ^= v[0] ^ v[7];
                    }
                } else {
                    uint32_t v[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = acc + i;
                    tmem_st32(tm + col, v);
                    tmem_wait_st();
                }
            }
        }
        __syncthreads();
        if (tid == 0) out[rep] = clock64() - t0;
    }
    if (acc == 0x12345678u) sink[tid] = acc;
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(s_tmem, 512);
}

static long long median(std::vector<long long> v) {
    std::sort(v.begin(), v.end());
    return v[v.size() / 2];
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main() {
    long long* d_out;
    uint32_t* d_sink;
    CK(cudaMalloc(&d_out, REPS * sizeof(long long)));
    CK(cudaMalloc(&d_sink, 512 * sizeof(uint32_t)));
    std::vector<long long> h(REPS);
    printf("tcgen05.mma kind::f16, M = 128, K = 16, %d MMAs + commit: cycles per MMA (first issue -> mbarrier observed)\n", CHAIN);
    printf("%6s %10s %10s %10s %12s %12s\n", "N", "floor", "SS", "TS(A=TMEM)", "SS 2 accum", "SS 4 accum");
    chain_row<16>(d_out);
    chain_row<64>(d_out);
    chain_row<128>(d_out);
    chain_row<256>(d_out);
    printf("\nTMEM epilogue traffic, bytes per cycle per SM (each warp: 64 x 32 columns x 32 lanes x 4 B)\n");
    printf("%8s %12s %12s %12s\n", "warps", "ld.x32", "ld.x8", "st.x32");
    const int Ws[3] = {4, 8, 16};
    for (int wi = 0; wi < 3; ++wi) {
        double r[3];
        for (int kind = 0; kind < 3; ++kind) {
            tmem_rw_kernel<<<1, 512>>>(kind, Ws[wi], d_out, d_sink);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), d_out, REPS * sizeof(long long), cudaMemcpyDeviceToHost));
            const double cyc = (double)median(std::vector<long long>(h.begin() + 2, h.end()));
            r[kind] = (double)Ws[wi] * 64 * 32 * 32 * 4 / cyc;
        }
        printf("%8d %12.1f %12.1f %12.1f\n", Ws[wi], r[0], r[1], r[2]);
    }
    return 0;
}
