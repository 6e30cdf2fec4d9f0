// Correctness probe for the next step of DESIGN.md section 7: tcgen05.mma with the A operand (the activation tile) in
// TENSOR MEMORY instead of shared memory. It checks the operand layout this repo would rely on before any product kernel
// is touched:
//     A [128 rows x K] bf16 lives in TMEM as packed pairs: lane = row, 32-bit column c of the K step = (A[row][2c], A[row][2c+1])
//     (low half = even k), 8 columns per 16-wide K step, written by the epilogue with tcgen05.st.32x32b (thread = lane).
//     B [N x K] bf16 in shared memory, K-major, no swizzle (the layout tc_chain.cu already uses).
// It computes D = A * B^T (fp32) for K = 64 (4 K steps) both ways -- SS (A from shared memory, the proven path) and TS (A from
// TMEM) -- and compares each with a host reference. Not run yet on hardware (round 1 ended with the GPU budget spent):
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I 3d-semantic-segmentation-amp-net_b200/csrc \
//        tools/ts_mma_probe.cu -o tools/build/ts_mma_probe && tools/build/ts_mma_probe
//
// Expected output: "SS max |err| ~1e-6 ... TS max |err| ~1e-6"; a large TS error means the assumed TMEM layout is wrong.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "tc_ptx.cuh"

using namespace amp::tcx;

constexpr int M = 128, N = 64, K = 64;

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

// a, b: row-major bf16 bit patterns ([M][K], [N][K]); d_ss, d_ts: [M][N] fp32
__global__ void __launch_bounds__(128, 1) probe_kernel(const uint16_t* __restrict__ a, const uint16_t* __restrict__ b,
                                                        float* __restrict__ d_ss, float* __restrict__ d_ts) {
    __shared__ __align__(1024) unsigned char s_a[M * K * 2];   // K-major no-swizzle: ((k / 8) * M + r) * 16 + (k % 8) * 2
    __shared__ __align__(1024) unsigned char s_b[N * K * 2];   //                     ((k / 8) * N + n) * 16 + (k % 8) * 2
    __shared__ uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = warp_index_uniform();
    for (int e = tid; e < M * K; e += 128) {
        const int r = e / K, k = e - r * K;
        *reinterpret_cast<uint16_t*>(s_a + ((k >> 3) * M + r) * 16 + (k & 7) * 2) = a[e];
    }
    for (int e = tid; e < N * K; e += 128) {
        const int n = e / K, k = e - n * K;
        *reinterpret_cast<uint16_t*>(s_b + ((k >> 3) * N + n) * 16 + (k & 7) * 2) = b[e];
    }
    if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
    if (warp == 1) tmem_alloc(smem_u32(&s_tmem), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = uniform_u32(s_tmem);
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    // A into TMEM columns [256, 256 + K / 2): this thread's row (lane = tid), 32 packed pairs = K = 64 elements
    {
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = (uint32_t)a[tid * K + 2 * c] | ((uint32_t)a[tid * K + 2 * c + 1] << 16);
        tmem_st32(tm + lane_addr + 256u, v);
        tmem_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t idesc = umma_idesc(M, N), barrier = smem_u32(&bar);
    const uint64_t a_d = umma_desc(smem_u32(s_a), (uint32_t)M * 16u, 128u), b_d = umma_desc(smem_u32(s_b), (uint32_t)N * 16u, 128u);
    const uint64_t a_step = (2u * M * 16u) >> 4, b_step = (2u * N * 16u) >> 4;          // one 16-wide K step = 2 K groups
    uint32_t phase = 0;
    for (int mode = 0; mode < 2; ++mode) {              // 0: SS into columns [0, N), 1: TS into columns [128, 128 + N)
        if (warp == 0) {
            tc_fence_after();
            if (elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < K / 16; ++ks) {
                    if (mode == 0) umma_bf16(tm, a_d + ks * a_step, b_d + ks * b_step, idesc, ks > 0 ? 1u : 0u);
                    else umma_bf16_ts(tm + 128u, tm + 256u + (uint32_t)(ks * 8), b_d + ks * b_step, idesc, ks > 0 ? 1u : 0u);
                }
                umma_commit(barrier);
            }
            __syncwarp();
        }
        mbar_wait(barrier, phase);
        phase ^= 1u;
        tc_fence_after();
        float* out = mode == 0 ? d_ss : d_ts;
#pragma unroll
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(tm + lane_addr + (uint32_t)(mode * 128 + c0), v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) out[tid * N + c0 + j] = __uint_as_float(v[j]);
        }
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
    }
    if (warp == 1) tmem_dealloc(tm, 512);
}

static uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
    uint32_t u = (uint32_t)h << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

int main() {
    std::vector<uint16_t> ha(M * K), hb(N * K);
    srand(7);
    for (auto& v : ha) v = f2bf((float)rand() / RAND_MAX * 2.f - 1.f);
    for (auto& v : hb) v = f2bf((float)rand() / RAND_MAX * 2.f - 1.f);
    std::vector<float> ref(M * N, 0.f), hs(M * N), ht(M * N);
    for (int r = 0; r < M; ++r)
        for (int n = 0; n < N; ++n) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += (double)bf2f(ha[r * K + k]) * (double)bf2f(hb[n * K + k]);
            ref[r * N + n] = (float)s;
        }
    uint16_t *da, *db;
    float *ds, *dt;
    CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2));
    CK(cudaMalloc(&ds, ref.size() * 4)); CK(cudaMalloc(&dt, ref.size() * 4));
    CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(ds, 0, ref.size() * 4)); CK(cudaMemset(dt, 0, ref.size() * 4));
    probe_kernel<<<1, 128>>>(da, db, ds, dt);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hs.data(), ds, ref.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ht.data(), dt, ref.size() * 4, cudaMemcpyDeviceToHost));
    double es = 0.0, et = 0.0;
    for (size_t i = 0; i < ref.size(); ++i) {
        es = fmax(es, fabs((double)hs[i] - ref[i]));
        et = fmax(et, fabs((double)ht[i] - ref[i]));
    }
    printf("D = A * B^T, M = %d, N = %d, K = %d, bf16 operands, fp32 accumulate\n", M, N, K);
    printf("SS (A in shared memory) max |err| = %.3e\nTS (A in tensor memory)  max |err| = %.3e\n", es, et);
    printf("%s\n", (es < 1e-4 && et < 1e-4) ? "OK: the packed-pair TMEM layout of the A operand is right" : "MISMATCH");
    return (es < 1e-4 && et < 1e-4) ? 0 : 2;
}
