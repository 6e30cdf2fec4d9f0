// Few-row linear layers in fp32: the per-cloud T-Net FC stacks (pointNet/model/pointnetAtt.py:38-40, rows = batch size)
// and the per-token projections of the attention layer (:183-190, rows = batch x blocks), forward and backward.
// With 32 .. 300 rows these GEMMs are a few MFLOP: the tiled kernels (128-row tiles, one CTA per 128 output channels)
// leave the chip idle and serialise on K, so they are weight-read / latency bound problems and get their own kernels:
//
//   small_fwd_kernel    y = epi(pro(x) . W^T) and dx = epi(pro(dy) . W) (transposed weights): a warp streams W for its
//                                                     output column against a chunk of the input rows held in shared
//                                                     memory (lane == reduction index), butterfly transpose-reduce -> lane == row
//   small_wgrad_kernel  dW = pro(dy)^T . pro(a)      thread == input channel k, 8 output channels per CTA, rows walked in
//                                                     order (deterministic, no partial buffers)
// Prologues / epilogues are those of PwParams / WgParams (BatchNorm apply or backward, ReLU mask, batch statistics).
// Because a warp holds a whole column of <= 32 rows, the BatchNorm sums are warp shuffles in a fixed order.
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "nn_common.cuh"

namespace amp {
namespace {

constexpr int SM_ROWS = 32, SM_CPW = 1, SM_COLS = 8 * SM_CPW, SM_MAXK = 512;   // one output column per warp: more CTAs, shorter chains

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// shared epilogue for one output element held by lane == row (valid when row < M); returns the value to store
struct SmallEpi {
    const PwParams& p; int M; int r0;
    __device__ __forceinline__ void run(float v, int lane, int n) const {
        const int r = r0 + lane;
        const bool ok = r < M;
        if (p.bias) v += __ldg(p.bias + n);
        if (p.accumulate && ok) v += p.Y[(long long)r * p.ldy + n];
        if (p.part_sum && !p.mask_y) {                       // forward batch statistics (rows <= 32: one tile)
            const float s = warp_sum(ok ? v : 0.f);
            const float d = ok ? v - s / (float)M : 0.f;
            const float q = warp_sum(d * d);
            if (lane == 0) { p.part_sum[n] = s; p.part_sq[n] = q; }
        }
        if (p.out_scale) v = fmaf(v, __ldg(p.out_scale + n), __ldg(p.out_shift + n));
        if (p.out_relu) v = fmaxf(v, 0.f);
        if (p.mask_y) {
            const float y = ok ? __ldg(p.mask_y + (long long)r * p.ld_mask + n) : 0.f;
            const float mu = __ldg(p.mask_mean + n);
            const float dz = (ok && fmaf(y - mu, __ldg(p.mask_scale + n), __ldg(p.mask_shift + n)) > 0.f) ? v : 0.f;
            if (p.part_sum) {
                const float is = p.mask_invstd ? __ldg(p.mask_invstd + n) : 0.f;
                const float s = warp_sum(dz), q = warp_sum(dz * (y - mu) * is);
                if (lane == 0) { p.part_sum[n] = s; p.part_sq[n] = q; }
            }
            v = dz;
        }
        if (ok && p.Y) p.Y[(long long)r * p.ldy + n] = v;
    }
};

__global__ void __launch_bounds__(256) small_fwd_kernel(const PwParams p) {
    pdl_sync();
    extern __shared__ float xs[];                 // [SM_ROWS][kc] chunk of the (prologue-applied) input rows
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int M = p.rows_per_cloud, K = p.K, N = p.Nout;
    const int r0 = blockIdx.y * SM_ROWS;
    const int KC = K < SM_MAXK ? K : SM_MAXK;
    float acc[SM_CPW][SM_ROWS];                   // this warp's output columns: c = warp + 8 j
#pragma unroll
    for (int j = 0; j < SM_CPW; ++j)
#pragma unroll
        for (int r = 0; r < SM_ROWS; ++r) acc[j][r] = 0.f;
    for (int k0 = 0; k0 < K; k0 += KC) {
        const int kc = min(KC, K - k0);
        if (k0) __syncthreads();
        // the weight loads of this chunk (2 columns x up to 16 lane-strided elements) do not depend on the input rows: they are
        // issued first, so that their latency overlaps the staging of the rows below (these layers are latency bound)
        float wv[SM_CPW][SM_MAXK / 32];
#pragma unroll
        for (int j = 0; j < SM_CPW; ++j) {
            const int n = blockIdx.x * SM_COLS + warp + 8 * j;
            // w_kn == 0: W[n][k] (a contiguous row per output); w_kn == 1: W[k][n] (transposed weights of the backward)
            const float* __restrict__ w = p.w_kn ? p.W + (long long)k0 * p.ldw + n : p.W + (long long)n * p.ldw + k0;
            const long long ws = p.w_kn ? p.ldw : 1;
#pragma unroll
            for (int i = 0; i < SM_MAXK / 32; ++i) {
                const int kk = lane + 32 * i;
                wv[j][i] = (n < N && kk < kc) ? __ldg(w + kk * ws) : 0.f;
            }
        }
        for (int kk = tid; kk < kc; kk += 256) {  // thread == input channel: prologue constants once, 32 independent row loads
            const int k = k0 + kk;
            const float m = p.in_m ? __ldg(p.in_m + k) : 0.f, a = p.in_a ? __ldg(p.in_a + k) : 1.f, b = p.in_b ? __ldg(p.in_b + k) : 0.f;
            const float cc = p.X2 ? __ldg(p.in_c + k) : 0.f;
            float x[SM_ROWS], x2[SM_ROWS];
#pragma unroll
            for (int r = 0; r < SM_ROWS; ++r) x[r] = (r0 + r < M) ? __ldg(p.X + (long long)(r0 + r) * p.ldx + k) : 0.f;
            if (p.X2) {
#pragma unroll
                for (int r = 0; r < SM_ROWS; ++r) x2[r] = (r0 + r < M) ? __ldg(p.X2 + (long long)(r0 + r) * p.ldx + k) : 0.f;
            }
#pragma unroll
            for (int r = 0; r < SM_ROWS; ++r) {
                float v = x[r];
                if (p.X2) v = fmaf(x2[r] - m, cc, fmaf(v, a, b));
                else if (p.in_a) v = fmaf(v - m, a, b);
                if (p.in_relu) v = fmaxf(v, 0.f);
                xs[r * kc + kk] = (r0 + r < M) ? v : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < SM_MAXK / 32; ++i) {
            const int kk = lane + 32 * i;
            if (32 * i < kc) {                               // warp-uniform
                const int ks = kk < kc ? kk : 0;
#pragma unroll
                for (int r = 0; r < SM_ROWS; ++r) {
                    const float xv = xs[r * kc + ks];
#pragma unroll
                    for (int j = 0; j < SM_CPW; ++j) acc[j][r] = fmaf(xv, wv[j][i], acc[j][r]);
                }
            }
        }
    }
    const SmallEpi epi{p, M, r0};
#pragma unroll
    for (int j = 0; j < SM_CPW; ++j) {
        const int n = blockIdx.x * SM_COLS + warp + 8 * j;
        if (n >= N) break;
        // butterfly: after the five steps lane L holds the full sum of row L
#pragma unroll
        for (int o = 16, n2 = 16; o >= 1; o >>= 1, n2 >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < n2; ++i) {
                const float send = up ? acc[j][i] : acc[j][i + n2];
                const float keep = up ? acc[j][i + n2] : acc[j][i];
                acc[j][i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
        epi.run(acc[j][0], lane, n);
    }
}

// dW[n, k] = sum_r dy'[r, n] * a'[r, k], db[n] = sum_r dy'[r, n]: thread == k, 8 output channels n per CTA
constexpr int SW_N = 8;
__global__ void __launch_bounds__(256) small_wgrad_kernel(const WgParams p) {
    pdl_sync();
    __shared__ float ys[SM_ROWS][SW_N];
    const int tid = threadIdx.x;
    const int M = p.rows_per_cloud, K = p.K, N = p.Nout;
    const int n0 = blockIdx.x * SW_N;
    float acc[SW_N];
#pragma unroll
    for (int j = 0; j < SW_N; ++j) acc[j] = 0.f;
    float bsum = 0.f;
    const bool has_k = tid < K;
    const float aa = (p.a_a && has_k) ? __ldg(p.a_a + tid) : 1.f, ab = (p.a_b && has_k) ? __ldg(p.a_b + tid) : 0.f;
    const float am = (p.a_m && has_k) ? __ldg(p.a_m + tid) : 0.f;
    for (int rb = 0; rb < M; rb += SM_ROWS) {
        float a[SM_ROWS];
#pragma unroll
        for (int r = 0; r < SM_ROWS; ++r) {
            float v = 0.f;
            if (has_k && rb + r < M) {
                v = __ldg(p.A + (long long)(rb + r) * p.lda + tid);
                if (p.a_a) v = fmaf(v - am, aa, ab);
                if (p.a_relu) v = fmaxf(v, 0.f);
            }
            a[r] = v;
        }
        {
            const int r = tid / SW_N, j = tid % SW_N, n = n0 + j;      // 256 threads == 32 rows x 8 channels
            float v = 0.f;
            if (rb + r < M && n < N) {
                const long long off = (long long)(rb + r) * p.lddy + n;
                v = __ldg(p.dY + off);
                if (p.y_a) v = fmaf(v, __ldg(p.y_a + n), __ldg(p.y_b + n));
                if (p.Y2) v = fmaf(__ldg(p.Y2 + off) - (p.y_m ? __ldg(p.y_m + n) : 0.f), __ldg(p.y_c + n), v);
            }
            __syncthreads();                                           // previous tile fully consumed
            ys[r][j] = v;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SM_ROWS; ++r) {
#pragma unroll
            for (int j = 0; j < SW_N; ++j) acc[j] = fmaf(ys[r][j], a[r], acc[j]);
        }
        if (tid < SW_N) {
#pragma unroll
            for (int r = 0; r < SM_ROWS; ++r) bsum += ys[r][tid];
        }
    }
    if (has_k) {
#pragma unroll
        for (int j = 0; j < SW_N; ++j) {
            const int n = n0 + j;
            if (n < N) {
                float* dst = p.dW + (p.w_kn ? (long long)tid * p.ldw + n : (long long)n * p.ldw + tid);
                *dst = p.accumulate ? *dst + acc[j] : acc[j];
            }
        }
    }
    if (p.db && tid < SW_N && n0 + tid < N) p.db[n0 + tid] = p.accumulate ? p.db[n0 + tid] + bsum : bsum;
}

// Narrow input (K <= 16: the 3 xyz / 9 raw feature columns, per-cloud or shared weights) over many rows, exact fp32:
// memory bound on dy'. One CTA per (cloud, slab) partial of wgrad_reduce_kernel; thread == (output channel n, row phase),
// the slab's input rows sit in shared memory, the row phases are added in a fixed order.
constexpr int NW_MAXK = 16, NW_SLAB_MAX = 512, NW_T = 512;   // 16 warps: the kernel is bound by the latency of its dy' loads
__global__ void __launch_bounds__(NW_T) narrow_wgrad_kernel(const WgParams p, int slabs, int SLAB) {
    pdl_sync();
    // input rows of the slab; reused for the phase partials (NW_T x (NW_MAXK + 1) floats)
    __shared__ float as[(NW_SLAB_MAX * NW_MAXK > NW_T * (NW_MAXK + 1)) ? NW_SLAB_MAX * NW_MAXK : NW_T * (NW_MAXK + 1)];
    float* red = as;
    const int tid = threadIdx.x;
    const int K = p.K, N = p.Nout, rows = p.rows_per_cloud;
    const int unit = blockIdx.x, cloud = unit / slabs, slab = unit - cloud * slabs;
    const int r_begin = slab * SLAB, r_end = min(rows, r_begin + SLAB);
    const long long cloud_row = (long long)cloud * rows;
    for (int e = tid; e < (r_end - r_begin) * K; e += NW_T) {
        const int r = e / K, k = e - r * K;
        float v = __ldg(p.A + (cloud_row + r_begin + r) * p.lda + k);
        if (p.a_a) v = fmaf(v - (p.a_m ? __ldg(p.a_m + k) : 0.f), __ldg(p.a_a + k), __ldg(p.a_b + k));
        if (p.a_relu) v = fmaxf(v, 0.f);
        as[e] = v;
    }
    __syncthreads();
    const int phases = NW_T / N, n = tid % N, ph = tid / N;      // N in {64, 128, 256}
    float acc[NW_MAXK], bsum = 0.f;
#pragma unroll
    for (int k = 0; k < NW_MAXK; ++k) acc[k] = 0.f;
    const float ya = p.y_a ? __ldg(p.y_a + n) : 1.f, yb = p.y_b ? __ldg(p.y_b + n) : 0.f;
    const float yc = p.Y2 ? __ldg(p.y_c + n) : 0.f, ym = (p.Y2 && p.y_m) ? __ldg(p.y_m + n) : 0.f;
    constexpr int NW_INFLIGHT = 16;                                  // rows per step: their loads are issued together
    for (int rb = r_begin + ph; rb < r_end; rb += NW_INFLIGHT * phases) {
        float dv[NW_INFLIGHT], y2[NW_INFLIGHT];
#pragma unroll
        for (int u = 0; u < NW_INFLIGHT; ++u) {
            const int r = rb + u * phases;
            const long long off = (cloud_row + r) * p.lddy + n;
            dv[u] = r < r_end ? __ldg(p.dY + off) : 0.f;
            y2[u] = (p.Y2 && r < r_end) ? __ldg(p.Y2 + off) : ym;
        }
#pragma unroll
        for (int u = 0; u < NW_INFLIGHT; ++u) {
            const int r = rb + u * phases;
            if (r < r_end) {
                float v = dv[u];
                if (p.y_a) v = fmaf(v, ya, yb);
                if (p.Y2) v = fmaf(y2[u] - ym, yc, v);
                bsum += v;
                const float* a = as + (r - r_begin) * K;
#pragma unroll
                for (int k = 0; k < NW_MAXK; ++k)
                    if (k < K) acc[k] = fmaf(v, a[k], acc[k]);
            }
        }
    }
    __syncthreads();                                             // everybody is done reading the input rows
#pragma unroll
    for (int k = 0; k < NW_MAXK; ++k) red[tid * (NW_MAXK + 1) + k] = acc[k];
    red[tid * (NW_MAXK + 1) + NW_MAXK] = bsum;
    __syncthreads();
    float* part = p.partials + (long long)unit * ((long long)N * K + N);
    for (int e = tid; e < N * (K + 1); e += NW_T) {
        const int nn = e / (K + 1), k = e - nn * (K + 1);
        float s = 0.f;
        for (int q = 0; q < phases; ++q) s += red[(q * N + nn) * (NW_MAXK + 1) + (k < K ? k : NW_MAXK)];
        if (k < K) part[(long long)nn * K + k] = s;
        else part[(long long)N * K + nn] = s;
    }
}

// Narrow INPUT forward (K <= 12: the xyz / 9-feature columns into the first Conv1d of the encoder and of the two T-Nets,
// pointnetAtt.py:31, :90), 64 output channels, exact fp32: memory bound on the 256-byte output rows. One warp owns a
// 128-row tile (the unit of the BatchNorm partials): the input rows go through shared memory, lane = (channel quad, row
// parity) keeps its 4 x K weights in registers and writes one 16-byte piece per row, so a warp store covers two whole rows.
// Epilogue: raw value + per-tile BatchNorm sums (training; one pass with a shift, merged like tc_layer) or folded
// BatchNorm + ReLU (eval).
// MASK: the same shape as an INPUT-GRADIENT layer (dz = (dy W) masked by the ReLU of the saved activation, dropout of the
// forward re-applied, per-tile sums for the BatchNorm backward): the 5-class logits layer's gradient, [B, C, rows] in.
template <int KP, bool MASK>
__global__ void __launch_bounds__(128) narrow_fwd_kernel(const PwParams p) {
    pdl_sync();
    __shared__ __align__(16) float xs[4][128 * KP];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, q = lane & 15, par = lane >> 4;
    const int K = p.K, rows = p.rows_per_cloud;
    const int tpc = (rows + 127) >> 7, n_tiles = p.n_clouds * tpc;
    float* xw = xs[warp];
    const bool stats = p.part_sum != nullptr;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f), bs = sh;
    if (p.out_scale) { sc = __ldg(reinterpret_cast<const float4*>(p.out_scale) + q); sh = __ldg(reinterpret_cast<const float4*>(p.out_shift) + q); }
    if (p.bias) bs = __ldg(reinterpret_cast<const float4*>(p.bias) + q);
    float msc[4] = {0.f, 0.f, 0.f, 0.f}, msh[4] = {0.f, 0.f, 0.f, 0.f}, mmu[4] = {0.f, 0.f, 0.f, 0.f}, mis[4] = {0.f, 0.f, 0.f, 0.f};
    if (MASK) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            msc[j] = __ldg(p.mask_scale + q * 4 + j); msh[j] = __ldg(p.mask_shift + q * 4 + j);
            mmu[j] = __ldg(p.mask_mean + q * 4 + j); mis[j] = __ldg(p.mask_invstd + q * 4 + j);
        }
    }
    const unsigned long long dseed = MASK ? eff_seed(p.out_drop_seed, p.drop_off) : 0ull;
    int w_cloud = -1;
    float w[4][KP];
    for (int tile = blockIdx.x * 4 + warp; tile < n_tiles; tile += gridDim.x * 4) {
        const int cloud = tile / tpc, r0 = (tile - cloud * tpc) << 7, valid = min(128, rows - r0);
        const long long row_base = (long long)cloud * rows + r0;
        if (w_cloud != (p.w_cloud_stride ? cloud : 0)) {           // this lane's 4 x K weights (per cloud for the folded input transform)
            w_cloud = p.w_cloud_stride ? cloud : 0;
            const float* __restrict__ W = p.W + (long long)w_cloud * p.w_cloud_stride;
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int k = 0; k < KP; ++k)
                    w[j][k] = k < K ? __ldg(p.w_kn ? W + (long long)k * p.ldw + q * 4 + j : W + (long long)(q * 4 + j) * p.ldw + k) : 0.f;
        }
        __syncwarp();                                              // the previous tile's reads of xw are done
        if (MASK && p.x_transposed) {                              // [cloud][K][rows]: consecutive lanes read consecutive rows
            for (int e = lane; e < 128 * KP; e += 32) {
                const int k = e >> 7, r = e & 127;
                xw[r * KP + k] = (r < valid && k < K) ? __ldg(p.X + ((long long)cloud * K + k) * rows + r0 + r) : 0.f;
            }
        } else {
            for (int e = lane; e < 128 * KP; e += 32) {
                const int r = e / KP, k = e - r * KP;
                xw[e] = (r < valid && k < K) ? __ldg(p.X + (row_base + r) * p.ldx + k) : 0.f;
            }
        }
        __syncwarp();
        float shift[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
        float* __restrict__ yrow = p.Y + row_base * p.ldy + q * 4;
#pragma unroll 4
        for (int it = 0; it < 64; ++it) {
            const int r = 2 * it + par;
            float x[KP];
#pragma unroll
            for (int k4 = 0; k4 < KP / 4; ++k4) {
                const float4 t = *reinterpret_cast<const float4*>(xw + r * KP + k4 * 4);
                x[k4 * 4] = t.x; x[k4 * 4 + 1] = t.y; x[k4 * 4 + 2] = t.z; x[k4 * 4 + 3] = t.w;
            }
            float v[4] = {bs.x, bs.y, bs.z, bs.w};
#pragma unroll
            for (int k = 0; k < KP; ++k)
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = fmaf(x[k], w[j][k], v[j]);
            if (MASK) {
                if (r < valid) {
                    const float4 ym4 = __ldg(reinterpret_cast<const float4*>(p.mask_y + (row_base + r) * p.ld_mask + q * 4));
                    const float ym[4] = {ym4.x, ym4.y, ym4.z, ym4.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float dz = v[j];
                        if (p.out_drop_p > 0.f) dz *= dropout_keep(dseed, (unsigned long long)(row_base + r) * 64 + q * 4 + j, p.out_drop_p);
                        dz = fmaf(ym[j] - mmu[j], msc[j], msh[j]) > 0.f ? dz : 0.f;
                        s1[j] += dz;
                        s2[j] = fmaf(dz, (ym[j] - mmu[j]) * mis[j], s2[j]);
                        v[j] = dz;
                    }
                    *reinterpret_cast<float4*>(yrow + (long long)r * p.ldy) = make_float4(v[0], v[1], v[2], v[3]);
                }
            } else if (r < valid) {
                if (stats) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (it == 0) shift[j] = v[j];
                        const float d = v[j] - shift[j];
                        s1[j] += d;
                        s2[j] = fmaf(d, d, s2[j]);
                    }
                }
                if (p.out_scale) { v[0] = fmaf(v[0], sc.x, sh.x); v[1] = fmaf(v[1], sc.y, sh.y); v[2] = fmaf(v[2], sc.z, sh.z); v[3] = fmaf(v[3], sc.w, sh.w); }
                if (p.out_relu) { v[0] = fmaxf(v[0], 0.f); v[1] = fmaxf(v[1], 0.f); v[2] = fmaxf(v[2], 0.f); v[3] = fmaxf(v[3], 0.f); }
                *reinterpret_cast<float4*>(yrow + (long long)r * p.ldy) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
        if (MASK) {                                                // sum(dz) and sum(dz * xhat) of the tile: even rows + odd rows
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float o_s1 = __shfl_xor_sync(0xffffffffu, s1[j], 16), o_s2 = __shfl_xor_sync(0xffffffffu, s2[j], 16);
                if (par == 0) {
                    p.part_sum[(long long)tile * 64 + q * 4 + j] = s1[j] + o_s1;
                    p.part_sq[(long long)tile * 64 + q * 4 + j] = s2[j] + o_s2;
                }
            }
        } else if (stats) {                                        // merge the even-row and the odd-row half (lanes q and q + 16)
            const float na = (float)((valid + 1) >> 1), nb = (float)(valid >> 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float o_shift = __shfl_xor_sync(0xffffffffu, shift[j], 16), o_s1 = __shfl_xor_sync(0xffffffffu, s1[j], 16);
                const float o_s2 = __shfl_xor_sync(0xffffffffu, s2[j], 16);
                if (par == 0) {
                    const float sum_a = fmaf(shift[j], na, s1[j]), sum_b = fmaf(o_shift, nb, o_s1);
                    float m2 = s2[j] - s1[j] * s1[j] / na;
                    if (nb > 0.f) {
                        const float delta = sum_b / nb - sum_a / na;
                        m2 += o_s2 - o_s1 * o_s1 / nb + delta * delta * (na * nb / (na + nb));
                    }
                    p.part_sum[(long long)tile * 64 + q * 4 + j] = sum_a + sum_b;
                    p.part_sq[(long long)tile * 64 + q * 4 + j] = m2;
                }
            }
        }
    }
}

// Narrow OUTPUT forward (Nout <= 8, K = 64: conv_4 of the head, pointnetAtt.py:206-207, logits stored [B, C, rows]), exact
// fp32: memory bound on the 256-byte input rows. CTA = 128-row tile; 16 lanes read one row (float4 each, 8 rows in flight per
// thread), apply the prologue (BatchNorm + ReLU + Dropout of the raw input in training), form the partial dot products of
// their 4 channels with every class and reduce them over the 16 lanes; the tile's logits leave through shared memory so
// that the transposed store is contiguous along the rows.
constexpr int NOF_MAXN = 8;
__global__ void __launch_bounds__(256) narrow_out_fwd_kernel(const PwParams p) {
    pdl_sync();
    __shared__ float outs[NOF_MAXN][128];
    const int tid = threadIdx.x, q = tid & 15, rsub = tid >> 4, k = q * 4;
    const int rows = p.rows_per_cloud, N = p.Nout;
    const int tpc = (rows + 127) >> 7, tile = blockIdx.x, cloud = tile / tpc, r0 = (tile - cloud * tpc) << 7, valid = min(128, rows - r0);
    const long long row_base = (long long)cloud * rows + r0;
    float w[NOF_MAXN][4];
#pragma unroll
    for (int n = 0; n < NOF_MAXN; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) w[n][j] = n < N ? __ldg(p.W + (long long)n * p.ldw + k + j) : 0.f;
    float4 aa = make_float4(1.f, 1.f, 1.f, 1.f), ab = make_float4(0.f, 0.f, 0.f, 0.f), am = ab;
    if (p.in_a) {
        aa = __ldg(reinterpret_cast<const float4*>(p.in_a + k)); ab = __ldg(reinterpret_cast<const float4*>(p.in_b + k));
        if (p.in_m) am = __ldg(reinterpret_cast<const float4*>(p.in_m + k));
    }
    float4 xv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = rsub + 16 * i;
        xv[i] = r < valid ? __ldg(reinterpret_cast<const float4*>(p.X + (row_base + r) * p.ldx + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int r = rsub + 16 * i;
        float a[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
        if (p.in_a) {
            a[0] = fmaf(a[0] - am.x, aa.x, ab.x); a[1] = fmaf(a[1] - am.y, aa.y, ab.y);
            a[2] = fmaf(a[2] - am.z, aa.z, ab.z); a[3] = fmaf(a[3] - am.w, aa.w, ab.w);
        }
        if (p.in_relu) {
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] = fmaxf(a[j], 0.f);
        }
        if (p.in_drop_p > 0.f) {
            const unsigned long long di = (unsigned long long)(row_base + r) * 64 + k;
#pragma unroll
            for (int j = 0; j < 4; ++j) a[j] *= dropout_keep(eff_seed(p.in_drop_seed, p.drop_off), di + j, p.in_drop_p);
        }
#pragma unroll
        for (int n = 0; n < NOF_MAXN; ++n) {
            if (n < N) {
                float s = fmaf(a[3], w[n][3], fmaf(a[2], w[n][2], fmaf(a[1], w[n][1], a[0] * w[n][0])));
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (q == 0) outs[n][r] = s + (p.bias ? __ldg(p.bias + n) : 0.f);
            }
        }
    }
    __syncthreads();
    for (int e = tid; e < N * 128; e += 256) {
        const int n = e >> 7, r = e & 127;
        if (r < valid) {
            if (p.y_transposed) p.Y[((long long)cloud * N + n) * rows + r0 + r] = outs[n][r];
            else p.Y[(row_base + r) * p.ldy + n] = outs[n][r];
        }
    }
}

// Narrow OUTPUT (Nout <= 8: the 5 class logits of the head's last layer, [B, C, N]-transposed gradients) over many rows:
// memory bound on the activation. thread == (4 consecutive input channels, row phase): 16-byte activation loads, 8 in flight
// per thread (32 KB per CTA); the per-row gradients are broadcast reads of the slab's staged [row][class] table.
constexpr int NO_MAXN = 8;
__global__ void __launch_bounds__(256) narrow_out_wgrad_kernel(const WgParams p, int slabs, int SLAB) {
    pdl_sync();
    __shared__ float dys[NW_SLAB_MAX * NO_MAXN];                 // the slab's gradients [row][class], staged once (coalesced)
    __shared__ float4 red[256];
    __shared__ float bred[64 * NO_MAXN];
    const int tid = threadIdx.x;
    const int K = p.K, N = p.Nout, rows = p.rows_per_cloud;
    const int unit = blockIdx.x, cloud = unit / slabs, slab = unit - cloud * slabs;
    const int r_begin = slab * SLAB, r_end = min(rows, r_begin + SLAB), nr = r_end - r_begin;
    const long long cloud_row = (long long)cloud * rows;
    if (p.dy_transposed) {                                       // [B, C, rows]: consecutive threads = consecutive rows of one class
        for (int n = 0; n < N; ++n) {
            const float* src = p.dY + ((long long)cloud * N + n) * rows + r_begin;
            for (int r = tid; r < nr; r += 256) dys[r * NO_MAXN + n] = __ldg(src + r);
        }
    } else {
        for (int e = tid; e < nr * N; e += 256) {
            const int r = e / N, n = e - r * N;
            dys[r * NO_MAXN + n] = __ldg(p.dY + (cloud_row + r_begin + r) * p.lddy + n);
        }
    }
    __syncthreads();
    const int kq = K >> 2, phases = 256 / kq, q = tid % kq, ph = tid / kq, k = q * 4;     // K in {64, 128, 256}: 16 / 8 / 4 row phases
    float4 aa = make_float4(1.f, 1.f, 1.f, 1.f), ab = make_float4(0.f, 0.f, 0.f, 0.f), am = ab;
    if (p.a_a) {
        aa = __ldg(reinterpret_cast<const float4*>(p.a_a + k)); ab = __ldg(reinterpret_cast<const float4*>(p.a_b + k));
        am = __ldg(reinterpret_cast<const float4*>(p.a_m + k));
    }
    float acc[NO_MAXN][4], bsum[NO_MAXN];
#pragma unroll
    for (int n = 0; n < NO_MAXN; ++n) { acc[n][0] = acc[n][1] = acc[n][2] = acc[n][3] = 0.f; bsum[n] = 0.f; }
#pragma unroll 1
    for (int rb = r_begin + ph; rb < r_end; rb += 8 * phases) {
        float4 av[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = rb + u * phases;
            av[u] = r < r_end ? __ldg(reinterpret_cast<const float4*>(p.A + (cloud_row + r) * p.lda + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int r = rb + u * phases;
            if (r < r_end) {
                float a[4] = {av[u].x, av[u].y, av[u].z, av[u].w};
                if (p.a_a) {
                    a[0] = fmaf(a[0] - am.x, aa.x, ab.x); a[1] = fmaf(a[1] - am.y, aa.y, ab.y);
                    a[2] = fmaf(a[2] - am.z, aa.z, ab.z); a[3] = fmaf(a[3] - am.w, aa.w, ab.w);
                }
                if (p.a_relu) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) a[j] = fmaxf(a[j], 0.f);
                }
                if (p.a_drop_p > 0.f) {
                    const unsigned long long di = (unsigned long long)(cloud_row + r) * K + k;
#pragma unroll
                    for (int j = 0; j < 4; ++j) a[j] *= dropout_keep(eff_seed(p.a_drop_seed, p.drop_off), di + j, p.a_drop_p);
                }
                const float* dyr = dys + (r - r_begin) * NO_MAXN;
#pragma unroll
                for (int n = 0; n < NO_MAXN; ++n) {
                    if (n < N) {
                        const float dy = dyr[n];
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[n][j] = fmaf(dy, a[j], acc[n][j]);
                        bsum[n] += dy;
                    }
                }
            }
        }
    }
    // fixed-order sums over the row phases, one class at a time
    float* part = p.partials + (long long)unit * ((long long)N * K + N);
#pragma unroll
    for (int n = 0; n < NO_MAXN; ++n) {
        if (n < N) {
            __syncthreads();
            red[tid] = make_float4(acc[n][0], acc[n][1], acc[n][2], acc[n][3]);
            __syncthreads();
            if (tid < K) {
                const float* rf = reinterpret_cast<const float*>(red);
                float sum = 0.f;
                for (int qq = 0; qq < phases; ++qq) sum += rf[(qq * kq + (tid >> 2)) * 4 + (tid & 3)];
                part[(long long)n * K + tid] = sum;
            }
        }
    }
    // bias partial: the threads of one phase all saw the same rows; add the phases (channel quad 0) in order
    if (q == 0) {
#pragma unroll
        for (int n = 0; n < NO_MAXN; ++n) bred[ph * NO_MAXN + n] = bsum[n];
    }
    __syncthreads();
    if (tid < N) {
        float sum = 0.f;
        for (int qq = 0; qq < phases; ++qq) sum += bred[qq * NO_MAXN + tid];
        part[(long long)N * K + tid] = sum;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Eval-mode T-Net FC stack (pointNet/model/pointnetAtt.py:38-46) in ONE launch: relu(bn_4(fc_1)) -> relu(bn_5(fc_2)) ->
// fc_3 + bias + identity. One cluster of 8 CTAs per tile of 32 clouds; every CTA computes a slice of each layer's output
// columns for all rows of the tile (the small_fwd scheme), the intermediate [32 x 256] / [32 x 128] activations go through
// global scratch, and a cluster barrier (release / acquire) separates the layers. Replaces 4 dependent launches.
// ---------------------------------------------------------------------------------------------------------------
struct TnetFcArgs {
    const float* pooled;                                   // [B, 256]
    const float* fc1; const float* s4; const float* t4;    // [256, 256]; folded BatchNorm scale / shift [256]
    const float* fc2; const float* s5; const float* t5;    // [128, 256]; [128]
    const float* fc3w; const float* fc3b;                  // [d * d, 128]; [d * d]
    float* h1; float* h2; float* out;                      // scratch [B, 256], [B, 128]; result [B, d * d]
    int B, d, do_fc3;
};

// weights of the NC consecutive output columns of this warp (lane <-> k mod 32), issued before the tile of the layer is
// available so that their DRAM latency overlaps the previous layer / the cluster barrier
template <int NC, int KI>
__device__ __forceinline__ void tnet_fc_load_w(float (&wv)[NC][KI], const float* __restrict__ W, int n0, int n_end) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < KI; ++i) wv[c][i] = (n0 + c < n_end) ? __ldg(W + (long long)(n0 + c) * (32 * KI) + lane + 32 * i) : 0.f;
}

template <int NC, int KI>
__device__ __forceinline__ void tnet_fc_layer(float* xs, const float* __restrict__ X, int r0, int M, const float (&wv)[NC][KI], int n0,
                                              int n_end, const float* __restrict__ bias, const float* __restrict__ scale,
                                              const float* __restrict__ shift, bool relu, int eye_d, float* __restrict__ Y, int ldy,
                                              unsigned char* __restrict__ pk = nullptr, long long pk_stride = 0) {
    constexpr int K = 32 * KI;
    const int tid = threadIdx.x, lane = tid & 31;
    __syncthreads();                                       // previous layer's reads of xs are done
    for (int k = tid; k < K; k += 256) {
#pragma unroll
        for (int r = 0; r < SM_ROWS; ++r) xs[r * K + k] = (r0 + r < M) ? __ldcg(X + (long long)(r0 + r) * K + k) : 0.f;
    }
    __syncthreads();
    if (n0 >= n_end) return;
    float acc[NC][SM_ROWS];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int r = 0; r < SM_ROWS; ++r) acc[c][r] = 0.f;
#pragma unroll
    for (int i = 0; i < KI; ++i) {
#pragma unroll
        for (int r = 0; r < SM_ROWS; ++r) {
            const float x = xs[r * K + lane + 32 * i];     // one shared-memory read feeds the NC columns of the warp
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[c][r] = fmaf(x, wv[c][i], acc[c][r]);
        }
    }
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int o = 16, n2 = 16; o >= 1; o >>= 1, n2 >>= 1) {
            const bool up = (lane & o) != 0;
#pragma unroll
            for (int i = 0; i < n2; ++i) {
                const float send = up ? acc[c][i] : acc[c][i + n2];
                const float keep = up ? acc[c][i + n2] : acc[c][i];
                acc[c][i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
            }
        }
        const int n = n0 + c;
        if (n < n_end) {
            float v = acc[c][0] + (bias ? __ldg(bias + n) : 0.f);
            if (scale) v = fmaf(v, __ldg(scale + n), __ldg(shift + n));
            if (relu) v = fmaxf(v, 0.f);
            if (eye_d && n % (eye_d + 1) == 0) v += 1.f;  // + identity on the diagonal of the d x d transform
            if (r0 + lane < M) {
                Y[(long long)(r0 + lane) * ldy + n] = v;
                if (pk) {
                    // the 64 x 64 transform F = Y[row] as the B operand of local = h @ F (weight [n'][k'] = F[k'][n'], n = k' * 64 + n'),
                    // split into bf16 hi + lo blocks [K/8][64][8] (tc_chain32.cuh): fused pack, no extra launch
                    const int kq = n >> 6, nq = n & 63;
                    const __nv_bfloat16 h = __float2bfloat16_rn(v);
                    unsigned char* dst = pk + (long long)(r0 + lane) * pk_stride + ((kq >> 3) * 64 + nq) * 16 + (kq & 7) * 2;
                    *reinterpret_cast<__nv_bfloat16*>(dst) = h;
                    *reinterpret_cast<__nv_bfloat16*>(dst + 64 * 64 * 2) = __float2bfloat16_rn(v - __bfloat162float(h));
                }
            }
        }
    }
}

__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(256) tnet_fc_eval_kernel(const TnetFcArgs a) {
    pdl_sync();
    extern __shared__ float xs[];                          // [32][256]
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank(), warp = threadIdx.x >> 5;
    const int r0 = (blockIdx.x / 8) * SM_ROWS;
    float w1[4][8], w2[2][8], w3[1][4];
    tnet_fc_load_w<4, 8>(w1, a.fc1, rank * 32 + warp * 4, 256);
    tnet_fc_layer<4, 8>(xs, a.pooled, r0, a.B, w1, rank * 32 + warp * 4, 256, nullptr, a.s4, a.t4, true, 0, a.h1, 256);
    tnet_fc_load_w<2, 8>(w2, a.fc2, rank * 16 + warp * 2, 128);
    __threadfence();
    cluster.sync();
    tnet_fc_layer<2, 8>(xs, a.h1, r0, a.B, w2, rank * 16 + warp * 2, 128, nullptr, a.s5, a.t5, true, 0, a.h2, 128);
    if (!a.do_fc3) return;
    const int nout = a.d * a.d;                            // <= 64: one column per warp
    tnet_fc_load_w<1, 4>(w3, a.fc3w, rank * 8 + warp, nout);
    __threadfence();
    cluster.sync();
    tnet_fc_layer<1, 4>(xs, a.h2, r0, a.B, w3, rank * 8 + warp, nout, a.fc3b, nullptr, nullptr, false, a.d, a.out, nout);
}

// The same stack as ONE launch of 128 resident CTAs: one output column per warp for fc_1 / fc_2 (the first 32 / 16 CTAs), four
// per warp for fc_3 (4096 columns of the 64 x 64 transform = every warp; 9 columns of the 3 x 3 one = three warps), grid
// barriers between the layers, every layer's weight rows fetched before the predecessor kernel is waited for. Replaces the
// cluster launch (8 CTAs pulling all weights, ~5 us per layer) plus, for the feature T-Net, the separate fc_3 launch.
constexpr int TG_CTAS = 128;
__global__ void __launch_bounds__(256) tnet_fc_grid_kernel(const TnetFcArgs a, unsigned char* __restrict__ pk, long long pk_stride,
                                                           unsigned int* bars) {
    pdl_trigger();
    extern __shared__ float xs[];                          // [32][256]
    const int warp = threadIdx.x >> 5, gw = blockIdx.x * 8 + warp;
    const int nout = a.d * a.d;
    float w1[1][8], w2[1][8], w3[4][4];
    tnet_fc_load_w<1, 8>(w1, a.fc1, gw, 256);              // parameters: safe to read before the predecessor has finished
    tnet_fc_load_w<1, 8>(w2, a.fc2, gw, 128);
    tnet_fc_load_w<4, 4>(w3, a.fc3w, gw * 4, nout);
    pdl_wait();
    if (blockIdx.x * 8 < 256)
        for (int r0 = 0; r0 < a.B; r0 += SM_ROWS)
            tnet_fc_layer<1, 8>(xs, a.pooled, r0, a.B, w1, gw, 256, nullptr, a.s4, a.t4, true, 0, a.h1, 256);
    grid_barrier(bars + 0, gridDim.x);
    if (blockIdx.x * 8 < 128)
        for (int r0 = 0; r0 < a.B; r0 += SM_ROWS)
            tnet_fc_layer<1, 8>(xs, a.h1, r0, a.B, w2, gw, 128, nullptr, a.s5, a.t5, true, 0, a.h2, 128);
    grid_barrier(bars + 1, gridDim.x);
    if (blockIdx.x * 32 < nout)
        for (int r0 = 0; r0 < a.B; r0 += SM_ROWS)
            tnet_fc_layer<4, 4>(xs, a.h2, r0, a.B, w3, gw * 4, nout, a.fc3b, nullptr, nullptr, false, a.d, a.out, nout, pk, pk_stride);
}

// fc_3 of the 64 x 64 feature transform (4096 outputs, K = 128) + bias + identity, written as the module output [B, 64, 64]
// AND as the packed per-cloud operand of the fused chains: 128 CTAs x 32 output columns, all rows of a 32-cloud tile per pass
__global__ void __launch_bounds__(256) tnet_fc3_pack_kernel(const float* __restrict__ h2, const float* __restrict__ w, const float* __restrict__ b,
                                                            int B, float* __restrict__ out, unsigned char* __restrict__ pk, long long pk_stride) {
    pdl_sync();
    extern __shared__ float xs[];                          // [32][128]
    const int warp = threadIdx.x >> 5;
    const int n0 = blockIdx.x * 32 + warp * 4;
    float w3[4][4];
    tnet_fc_load_w<4, 4>(w3, w, n0, 4096);
    for (int r0 = 0; r0 < B; r0 += SM_ROWS)
        tnet_fc_layer<4, 4>(xs, h2, r0, B, w3, n0, 4096, b, nullptr, nullptr, false, 64, out, 4096, pk, pk_stride);
}

}  // namespace

int tnet_fc3_pack(const float* h2, const float* w, const float* b, int B, float* out, unsigned char* pk, long long pk_stride, cudaStream_t st) {
    if (!h2 || !w || !b || !out || !pk || B < 1) return fail(AMP_E_BADARG, "tnet_fc3_pack: null pointer");
    launch_pdl(tnet_fc3_pack_kernel, dim3(128), dim3(256), sizeof(float) * SM_ROWS * 128, st, h2, w, b, B, out, pk, pk_stride);
    count_launch();
    return check_launch("tnet_fc3_pack");
}

int tnet_fc_grid(const float* pooled, int B, const float* fc1, const float* s4, const float* t4, const float* fc2, const float* s5,
                 const float* t5, const float* fc3w, const float* fc3b, int d, float* h1, float* h2, float* out, unsigned char* pk,
                 long long pk_stride, unsigned int* bars, cudaStream_t st) {
    if (path_disabled("tnet_grid")) return 0;
    if (!pooled || !fc1 || !fc2 || !fc3w || !fc3b || !h1 || !h2 || !out || !bars || B < 1) return fail(AMP_E_BADARG, "tnet_fc_grid: null pointer");
    if (d * d > TG_CTAS * 32 || (pk && d != 64)) return 0;
    const size_t smem = sizeof(float) * SM_ROWS * 256;
    static int capacity = -1;                              // the grid barriers need the whole grid resident at once
    if (capacity < 0) {
        int dev = 0, sms = 0, per_sm = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tnet_fc_grid_kernel, 256, smem) != cudaSuccess)
            capacity = 0;
        else
            capacity = sms * per_sm;
    }
    if (capacity < TG_CTAS) return 0;
    TnetFcArgs a{pooled, fc1, s4, t4, fc2, s5, t5, fc3w, fc3b, h1, h2, out, B, d, 1};
    launch_pdl(tnet_fc_grid_kernel, dim3(TG_CTAS), dim3(256), smem, st, a, pk, pk_stride, bars);
    count_launch();
    count_path("tnet_grid");
    const int rc = check_launch("tnet_fc_grid");
    return rc == AMP_OK ? 1 : rc;
}

// Eval-mode T-Net FC stack in one cluster launch; with fc3_inside == 0 the caller runs fc_3 (wide) itself on h2.
int tnet_fc_eval(const float* pooled, int B, const float* fc1, const float* s4, const float* t4, const float* fc2, const float* s5,
                 const float* t5, const float* fc3w, const float* fc3b, int d, int fc3_inside, float* h1, float* h2, float* out,
                 cudaStream_t st) {
    if (!pooled || !fc1 || !fc2 || !h1 || !h2 || B < 1 || (fc3_inside && (!fc3w || !fc3b || !out)))
        return fail(AMP_E_BADARG, "tnet_fc_eval: null pointer");
    TnetFcArgs a{pooled, fc1, s4, t4, fc2, s5, t5, fc3w, fc3b, h1, h2, out, B, d, fc3_inside};
    const size_t smem = sizeof(float) * SM_ROWS * 256;
    launch_pdl(tnet_fc_eval_kernel, dim3((unsigned)(8 * ((B + SM_ROWS - 1) / SM_ROWS))), dim3(256), smem, st, a);
    count_launch();
    return check_launch("tnet_fc_eval");
}

// Forward of the narrow-output layer (class logits): 1 = launched, 0 = not eligible.
int narrow_out_fwd_try(const PwParams& p, cudaStream_t st) {
    if (path_disabled("narrow_out_fwd")) return 0;
    if (p.Nout > NOF_MAXN || p.K != 64 || !p.Y || p.x_transposed || p.X2 || p.in_c || p.out_scale || p.out_relu || p.out_drop_p != 0.f ||
        p.mask_y || p.pool_mode || p.part_sum || p.accumulate || p.group_rows || p.n_groups > 1 || p.bias_group_stride != 0 ||
        p.w_cloud_stride != 0 || p.w_kn)
        return 0;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if ((long long)p.n_clouds * p.rows_per_cloud < 2048 || p.ldx % 4 || !al16(p.X)) return 0;
    if (p.in_a && (!p.in_b || !al16(p.in_a) || !al16(p.in_b) || (p.in_m && !al16(p.in_m)))) return 0;
    const long long tiles = (long long)p.n_clouds * ((p.rows_per_cloud + 127) / 128);
    if (tiles > 0x7fffffffLL) return 0;
    launch_pdl(narrow_out_fwd_kernel, dim3((unsigned)tiles), dim3(256), 0, st, p);
    count_launch();
    const int rc = check_launch("narrow_out_fwd");
    return rc == AMP_OK ? 1 : rc;
}

// Forward of the narrow-input layers: 1 = launched, 0 = not eligible (the caller continues down the dispatch list).
int narrow_fwd_try(const PwParams& p, cudaStream_t st) {
    if (path_disabled("narrow_fwd")) return 0;
    if (p.K > 12 || p.Nout != 64 || !p.Y || p.y_transposed || p.X2 || p.in_a || p.in_relu || p.in_drop_p != 0.f ||
        p.pool_mode || p.accumulate || p.group_rows || p.n_groups > 1 || p.bias_group_stride != 0 || (p.out_relu && !p.out_scale))
        return 0;
    if (p.mask_y) {           // input-gradient form
        if (!p.part_sum || !p.mask_invstd || p.bias || p.out_scale || p.out_relu || p.w_cloud_stride || p.ld_mask % 4 ||
            (reinterpret_cast<uintptr_t>(p.mask_y) & 15))
            return 0;
    } else if (p.x_transposed || p.out_drop_p != 0.f) {
        return 0;
    }
    if ((long long)p.n_clouds * p.rows_per_cloud < 2048 || p.ldy % 4 || (reinterpret_cast<uintptr_t>(p.Y) & 15)) return 0;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if ((p.bias && !al16(p.bias)) || (p.out_scale && (!al16(p.out_scale) || !al16(p.out_shift)))) return 0;
    const long long tiles = (long long)p.n_clouds * ((p.rows_per_cloud + 127) / 128);
    long long grid = (tiles + 3) / 4;
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    if (p.mask_y) {
        if (p.K <= 8) launch_pdl(narrow_fwd_kernel<8, true>, dim3((unsigned)grid), dim3(128), 0, st, p);
        else launch_pdl(narrow_fwd_kernel<12, true>, dim3((unsigned)grid), dim3(128), 0, st, p);
        count_path("narrow_dgrad");
    } else if (p.K <= 4) launch_pdl(narrow_fwd_kernel<4, false>, dim3((unsigned)grid), dim3(128), 0, st, p);
    else launch_pdl(narrow_fwd_kernel<12, false>, dim3((unsigned)grid), dim3(128), 0, st, p);
    count_launch();
    const int rc = check_launch("narrow_fwd");
    return rc == AMP_OK ? 1 : rc;
}

int narrow_out_wgrad_try(const WgParams& p, int slabs, int SLAB, cudaStream_t st) {
    if (path_disabled("narrow_out_wgrad")) return 0;
    if (p.Nout > NO_MAXN || p.y_a || p.Y2 || p.dbg || SLAB > NW_SLAB_MAX) return 0;
    if (p.K != 64 && p.K != 128 && p.K != 256) return 0;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (p.lda % 4 || !al16(p.A) || (p.a_a && (!al16(p.a_a) || !al16(p.a_b) || !al16(p.a_m)))) return 0;   // 16-byte loads
    launch_pdl(narrow_out_wgrad_kernel, dim3((unsigned)(p.n_clouds * slabs)), dim3(256), 0, st, p, slabs, SLAB);
    count_launch();
    const int rc = check_launch("narrow_out_wgrad");
    return rc == AMP_OK ? 1 : rc;
}

// partial pass of wgrad() for narrow inputs: 1 = launched (the caller still runs wgrad_reduce), 0 = not eligible
int narrow_wgrad_try(const WgParams& p, int slabs, int SLAB, cudaStream_t st) {
    if (path_disabled("narrow_wgrad")) return 0;
    if (p.K > NW_MAXK || SLAB > NW_SLAB_MAX || p.dy_transposed || p.a_drop_p != 0.f) return 0;
    if (p.Nout != 64 && p.Nout != 128 && p.Nout != 256) return 0;
    launch_pdl(narrow_wgrad_kernel, dim3((unsigned)(p.n_clouds * slabs)), dim3(NW_T), 0, st, p, slabs, SLAB);
    count_launch();
    const int rc = check_launch("narrow_wgrad");
    return rc == AMP_OK ? 1 : rc;
}

// Split-K pair for few rows (<= 32) and a long reduction with transposed weights W[k][n] (input gradients of wide FC
// layers: 32 x 4096 -> 128 took 78 us as eight serial 512-wide chunks on 16 CTAs): one CTA per 32-wide reduction slice
// writes a partial [Nout][32 rows]; the second kernel adds the slices in a fixed order (deterministic) and runs the epilogue.
constexpr int SK_KS = 32;
__global__ void __launch_bounds__(128) small_splitk_kernel(const PwParams p, float* __restrict__ part) {
    pdl_sync();
    __shared__ __align__(16) float xs[32][SK_KS + 4];
    const int tid = threadIdx.x, M = p.rows_per_cloud, N = p.Nout;
    const int s = blockIdx.x, k0 = s * SK_KS;
    for (int e = tid; e < 32 * SK_KS; e += 128) {
        const int r = e / SK_KS, k = e - r * SK_KS;
        xs[r][k] = r < M ? __ldg(p.X + (long long)r * p.ldx + k0 + k) : 0.f;
    }
    __syncthreads();
    for (int n = tid; n < N; n += 128) {
        float w[SK_KS];
#pragma unroll
        for (int k = 0; k < SK_KS; ++k) w[k] = __ldg(p.W + (long long)(k0 + k) * p.ldw + n);
        float* dst = part + ((long long)s * N + n) * 32;
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int k4 = 0; k4 < SK_KS / 4; ++k4) {
                const float4 x = *reinterpret_cast<const float4*>(&xs[r][4 * k4]);
                acc = fmaf(x.x, w[4 * k4], acc); acc = fmaf(x.y, w[4 * k4 + 1], acc);
                acc = fmaf(x.z, w[4 * k4 + 2], acc); acc = fmaf(x.w, w[4 * k4 + 3], acc);
            }
            dst[r] = acc;
        }
    }
}

__global__ void __launch_bounds__(256) small_splitk_reduce_kernel(const PwParams p, const float* __restrict__ part, int n_slices) {
    pdl_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * 8 + warp, N = p.Nout;
    if (n >= N) return;
    float v = 0.f;
    for (int s0 = 0; s0 < n_slices; s0 += 16) {                  // 16 loads in flight, added in slice order
        float t[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) t[u] = s0 + u < n_slices ? part[((long long)(s0 + u) * N + n) * 32 + lane] : 0.f;
#pragma unroll
        for (int u = 0; u < 16; ++u) v += t[u];
    }
    const SmallEpi epi{p, p.rows_per_cloud, 0};
    epi.run(v, lane, n);
}

int small_linear_try(const PwParams& p, cudaStream_t st) {
    if (path_disabled(p.w_kn ? "small_dgrad" : "small_fwd")) return 0;
    if (p.n_clouds != 1 || p.rows_per_cloud > 1024 || p.x_transposed || p.y_transposed || p.w_cloud_stride || p.group_rows ||
        p.pool_mode || p.in_drop_p != 0.f || p.out_drop_p != 0.f || !p.Y)
        return 0;
    const bool stats = p.part_sum != nullptr;
    if ((stats || p.mask_y) && p.rows_per_cloud > SM_ROWS) return 0;      // the warp-level sums cover one 32-row tile
    const dim3 grid_rows((p.rows_per_cloud + SM_ROWS - 1) / SM_ROWS);
    if (p.w_kn && !p.bias && p.K >= 1024 && p.K % SK_KS == 0 && p.rows_per_cloud <= 32 && !p.in_a && !p.X2 && !p.in_relu && p.splitk_ws &&
        p.splitk_floats >= (size_t)(p.K / SK_KS) * p.Nout * 32 && !path_disabled("small_splitk")) {
        const int n_slices = p.K / SK_KS;
        launch_pdl(small_splitk_kernel, dim3((unsigned)n_slices), dim3(128), 0, st, p, p.splitk_ws);
        launch_pdl(small_splitk_reduce_kernel, dim3((unsigned)((p.Nout + 7) / 8)), dim3(256), 0, st, p, (const float*)p.splitk_ws, n_slices);
        count_launch(2);
        const int rc = check_launch("small_splitk");
        return rc == AMP_OK ? 1 : rc;
    }
    {
        if (p.w_kn && p.bias) return 0;
        dim3 grid((p.Nout + SM_COLS - 1) / SM_COLS, grid_rows.x);
        const size_t smem = sizeof(float) * SM_ROWS * (p.K < SM_MAXK ? p.K : SM_MAXK);
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(small_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(float) * SM_ROWS * SM_MAXK));
            attr_set = true;
        }
        launch_pdl(small_fwd_kernel, grid, dim3(256), smem, st, p);
    }
    count_launch();
    const int rc = check_launch("small_linear");
    return rc == AMP_OK ? 1 : rc;
}

int small_wgrad_try(const WgParams& p, cudaStream_t st) {
    if (path_disabled("small_wgrad")) return 0;
    if (p.n_clouds != 1 || p.rows_per_cloud > 1024 || p.dy_transposed || p.per_cloud || p.dbg || p.K > 256 || p.a_drop_p != 0.f) return 0;
    launch_pdl(small_wgrad_kernel, dim3((unsigned)((p.Nout + SW_N - 1) / SW_N)), dim3(256), 0, st, p);
    count_launch();
    const int rc = check_launch("small_wgrad");
    return rc == AMP_OK ? 1 : rc;
}

}  // namespace amp
