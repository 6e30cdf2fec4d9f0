// Point-wise linear layer on the tcgen05 tensor cores at fp32-class accuracy ("split bf16": every fp32 operand is
// the sum of two bf16 terms, hi + lo, and the product is accumulated as  lo*hi + hi*lo + hi*hi  in the fp32 TMEM
// accumulator; the dropped lo*lo term and the representation residual are ~2^-16 relative, so a layer is good to
// ~1e-5 where plain bf16 gives ~3e-3). Same contract as pw_linear_kernel (PwParams, nn_common.cuh): it is the
// large-row path of pw_linear() for the training forward, the input-gradient GEMMs of the backward and the fp32
// eval forward of pointNet/model/pointnetAtt.py's Conv1d(k=1) layers (:31-33, :90-103, :203-206).
//
// The GEMM runs TRANSPOSED:  D^T[n, r] = sum_k W[n, k] * pro(X)[r, k]   (UMMA M = 128 output channels, N = 128 rows)
// so that a TMEM lane (= epilogue thread) is an output channel and everything BatchNorm needs -- the sum over the
// rows of the tile, the squared deviations, the max / min with their row index, the ReLU-mask sums of the backward --
// is a per-thread loop over accumulator columns with per-channel constants in registers; stores and mask reads are
// coalesced across the warp (consecutive lanes = consecutive channels of one row).
//
// CTA = 256 threads = 2 independent warpgroups ("slots": own B-operand chunk buffers, 256 TMEM columns, mbarrier),
// weights (hi + lo, K-major no-swizzle core-matrix layout) resident in shared memory for all tiles of the CTA,
// K walked in chunks of 64 (stage chunk -> 3 MMAs per 16-wide K step -> commit -> wait).
#include <type_traits>

#include "nn_common.cuh"
#include "tc_ptx.cuh"

namespace amp {
namespace {
using namespace tcx;

constexpr int TL_THREADS = 256, TL_ROWS = 128, TL_KC = 64;
constexpr int TL_MAX_WELEMS = 32768;                 // Mpad * K
constexpr int TL_BHALF = TL_ROWS * TL_KC * 2;        // bytes of the hi (or lo) half of one B chunk
constexpr int TL_MAX_SMEM = 232448, TL_MIN_SMEM = 120 * 1024;

struct TlPlan { int w_lo, b0, tab, bar, total; };
__host__ __device__ inline TlPlan tl_plan(int Mpad, int K) {
    TlPlan s;
    const int wbytes = Mpad * K * 2;
    s.w_lo = wbytes;
    s.b0 = 2 * wbytes;
    s.tab = s.b0 + 4 * TL_BHALF;
    s.bar = s.tab + 16 * K;
    s.total = s.bar + 64;
    return s;
}

// v = hi + lo (+ 2^-18 residual): one packed convert per pair, the bf16 -> fp32 widening is a shift / mask
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(a, b);
    lo = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
__device__ __forceinline__ void split_store8(const float (&v)[8], uint4* hi_dst, uint4* lo_dst) {
    uint4 h, l;
    split_pair(v[0], v[1], h.x, l.x); split_pair(v[2], v[3], h.y, l.y);
    split_pair(v[4], v[5], h.z, l.z); split_pair(v[6], v[7], h.w, l.w);
    *hi_dst = h; *lo_dst = l;
}

// epilogue specialisations (bit mask): compile-time so that the per-element loop carries no dead branches
enum { TL_STATS = 1, TL_POOL2 = 2, TL_AFFINE = 4, TL_POOL1 = 8, TL_MASK = 16, TL_DROP = 32, TL_ACC = 64 };

template <int MODE>
__global__ void __launch_bounds__(TL_THREADS, 1) tc_layer_kernel(const __grid_constant__ PwParams p, const int Mpad) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wg = warp >> 2, wtid = tid & 127;
    const int K = p.K, Nout = p.Nout, rows = p.rows_per_cloud;
    const TlPlan sp = tl_plan(Mpad, K);
    __nv_bfloat16* s_whi = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* s_wlo = reinterpret_cast<__nv_bfloat16*>(smem + sp.w_lo);
    unsigned char* s_bhi = smem + sp.b0 + wg * 2 * TL_BHALF;
    unsigned char* s_blo = s_bhi + TL_BHALF;
    float* s_a = reinterpret_cast<float*>(smem + sp.tab);       // prologue constants per input channel
    float* s_b = s_a + K; float* s_c = s_b + K; float* s_m = s_c + K;
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + sp.bar);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + sp.bar + 32);
    const uint32_t mbar = smem_u32(&s_bar[wg]);

    if (tid == 0) {
        mbar_init(smem_u32(&s_bar[0]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(s_tmem), 512);
    for (int k = tid; k < K; k += TL_THREADS) {
        s_a[k] = p.in_a ? __ldg(p.in_a + k) : 1.f;
        s_b[k] = p.in_b ? __ldg(p.in_b + k) : 0.f;
        s_c[k] = p.in_c ? __ldg(p.in_c + k) : 0.f;
        s_m[k] = p.in_m ? __ldg(p.in_m + k) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;

    const int tpc = (rows + TL_ROWS - 1) / TL_ROWS;
    const int n_mt = Mpad >> 7;
    const int lrow = (warp & 3) * 32 + lane;                     // staging: tile row; epilogue: channel lane
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t slot_col = tmem_base + (uint32_t)(wg * 256);
    const uint32_t whi_addr = smem_u32(s_whi), wlo_addr = smem_u32(s_wlo), bhi_addr = smem_u32(s_bhi), blo_addr = smem_u32(s_blo);
    const uint32_t w_lbo = (uint32_t)Mpad * 16u;
    const uint32_t idesc = umma_idesc(128, 128);
    const bool per_cloud_w = p.w_cloud_stride != 0;
    const bool has_pro = p.in_a != nullptr;
    uint32_t phase = 0;

    // weights -> (hi, lo) core-matrix layout: one thread per (channel n, 8 consecutive k) = one 16-byte row of a core matrix
    auto stage_weights = [&](int cloud) {
        const float* __restrict__ W = p.W + (long long)cloud * p.w_cloud_stride;
        const int k8n = K >> 3, total = Mpad * k8n;
        uint4* whi4 = reinterpret_cast<uint4*>(s_whi);
        uint4* wlo4 = reinterpret_cast<uint4*>(s_wlo);
#pragma unroll 4
        for (int e = tid; e < total; e += TL_THREADS) {
            int n, k8;
            if (p.w_kn == 0) { n = e / k8n; k8 = e - n * k8n; } else { n = e & (Mpad - 1); k8 = e / Mpad; }
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = 0.f;
            if (n < Nout) {
                if (p.w_kn == 0) {
                    const float* src = W + (long long)n * p.ldw + k8 * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = __ldg(src + i);
                } else {
                    const float* src = W + (long long)(k8 * 8) * p.ldw + n;
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = __ldg(src + (long long)i * p.ldw);
                }
            }
            split_store8(v, whi4 + k8 * Mpad + n, wlo4 + k8 * Mpad + n);
        }
        fence_proxy_async();
    };

    // pull this thread's row of a later tile into L2 while the current tile is being worked on (no registers held)
    auto prefetch_tile = [&](int cloud, int t) {
        const long long r = (long long)cloud * rows + t * TL_ROWS + lrow;
        if (t * TL_ROWS + lrow < rows) {
            const char* x = reinterpret_cast<const char*>(p.X + r * p.ldx);
            for (int b = 0; b < K * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(x + b));
            if (p.X2) {
                const char* x2 = reinterpret_cast<const char*>(p.X2 + r * p.ldx);
                for (int b = 0; b < K * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(x2 + b));
            }
        }
    };

    auto process = [&](int cloud, int t) {
        const int row0 = t * TL_ROWS;
        const int valid = min(TL_ROWS, rows - row0);
        const long long row_base = (long long)cloud * rows + row0;        // global row of tile row 0
        const long long tile = (long long)cloud * tpc + t;
        // ---------------- K chunks: stage B operand (hi, lo), MMA ----------------
        const bool row_ok = lrow < valid;
        if (row_ok && (MODE & (TL_MASK | TL_ACC))) {          // rows the epilogue of THIS tile will read: start them towards L2 now
            if (MODE & TL_MASK) {
                const char* m = reinterpret_cast<const char*>(p.mask_y + (row_base + lrow) * p.ld_mask);
                for (int b = 0; b < Nout * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(m + b));
            }
            if (MODE & TL_ACC) {
                const char* y = reinterpret_cast<const char*>(p.Y + (row_base + lrow) * p.ldy);
                for (int b = 0; b < Nout * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(y + b));
            }
        }
        const float* __restrict__ xrow = p.X + (row_base + lrow) * p.ldx;
        const float* __restrict__ x2row = p.X2 ? p.X2 + (row_base + lrow) * p.ldx : nullptr;
        for (int kc = 0; kc * TL_KC < K; ++kc) {
            const int kcur = min(TL_KC, K - kc * TL_KC);
            // 32 input channels per batch: all global loads of the batch are issued before the first use
            for (int b0 = 0; b0 < kcur; b0 += 32) {
                const int k0 = kc * TL_KC + b0;
                const int ng = min(4, (kcur - b0) >> 3);               // 8-channel groups in this batch (2 or 4)
                float4 xa[8], ya[8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    xa[i] = (row_ok && (i >> 1) < ng) ? __ldg(reinterpret_cast<const float4*>(xrow + k0) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                if (x2row) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        ya[i] = (row_ok && (i >> 1) < ng) ? __ldg(reinterpret_cast<const float4*>(x2row + k0) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    if (c < ng) {
                        const int k = k0 + c * 8;
                        float v[8] = {xa[2 * c].x, xa[2 * c].y, xa[2 * c].z, xa[2 * c].w, xa[2 * c + 1].x, xa[2 * c + 1].y, xa[2 * c + 1].z, xa[2 * c + 1].w};
                        if (has_pro) {
                            const float4 a0 = *reinterpret_cast<const float4*>(s_a + k), a1 = *reinterpret_cast<const float4*>(s_a + k + 4);
                            const float4 q0 = *reinterpret_cast<const float4*>(s_b + k), q1 = *reinterpret_cast<const float4*>(s_b + k + 4);
                            const float4 m0 = *reinterpret_cast<const float4*>(s_m + k), m1 = *reinterpret_cast<const float4*>(s_m + k + 4);
                            const float pa[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                            const float pb[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                            const float pm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
                            if (x2row) {
                                const float4 c0 = *reinterpret_cast<const float4*>(s_c + k), c1 = *reinterpret_cast<const float4*>(s_c + k + 4);
                                const float y2[8] = {ya[2 * c].x, ya[2 * c].y, ya[2 * c].z, ya[2 * c].w, ya[2 * c + 1].x, ya[2 * c + 1].y, ya[2 * c + 1].z, ya[2 * c + 1].w};
                                const float pc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = fmaf(y2[i] - pm[i], pc[i], fmaf(v[i], pa[i], pb[i]));
                            } else {
#pragma unroll
                                for (int i = 0; i < 8; ++i) v[i] = fmaf(v[i] - pm[i], pa[i], pb[i]);
                            }
                        }
                        if (p.in_relu) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
                        }
                        if (p.in_drop_p > 0.f) {
#pragma unroll
                            for (int i = 0; i < 8; ++i)
                                v[i] *= dropout_keep(p.in_drop_seed, (unsigned long long)(row_base + lrow) * K + k + i, p.in_drop_p);
                        }
                        if (!row_ok) {
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = 0.f;
                        }
                        const int c8 = (b0 >> 3) + c;
                        split_store8(v, reinterpret_cast<uint4*>(s_bhi) + c8 * TL_ROWS + lrow, reinterpret_cast<uint4*>(s_blo) + c8 * TL_ROWS + lrow);
                    }
                }
            }
            fence_proxy_async();
            tc_fence_before();
            wg_bar_sync(wg);
            if (wtid == 0) {
                tc_fence_after();
                for (int mt = 0; mt < n_mt; ++mt) {
                    const uint32_t d = slot_col + (uint32_t)mt * 128u;
                    for (int ks = 0; ks < (kcur >> 4); ++ks) {
                        const uint32_t kg = (uint32_t)(kc * (TL_KC >> 4) + ks);
                        const uint32_t woff = (uint32_t)mt * 2048u + kg * 2u * w_lbo;
                        const uint64_t a_hi = umma_desc(whi_addr + woff, w_lbo, 128u), a_lo = umma_desc(wlo_addr + woff, w_lbo, 128u);
                        const uint64_t b_hi = umma_desc(bhi_addr + (uint32_t)ks * 4096u, 2048u, 128u);
                        const uint64_t b_lo = umma_desc(blo_addr + (uint32_t)ks * 4096u, 2048u, 128u);
                        umma_bf16(d, a_lo, b_hi, idesc, (kc | ks) != 0);
                        umma_bf16(d, a_hi, b_lo, idesc, 1u);
                        umma_bf16(d, a_hi, b_hi, idesc, 1u);
                    }
                }
                umma_commit(mbar);
            }
            __syncwarp();
            mbar_wait(mbar, phase);
            phase ^= 1u;
            tc_fence_after();
        }

        // ---------------- epilogue: lane = output channel, columns = rows of the tile ----------------
        int g_tile = 0;                                        // row group of the tile (groups are tile aligned)
        if (p.bias && p.group_rows)
            for (int q = 1; q < p.n_groups; ++q) g_tile += (row0 >= __ldg(p.group_rows + q)) ? 1 : 0;
        for (int mt = 0; mt < n_mt; ++mt) {
            const int n = mt * 128 + lrow;
            const bool n_ok = n < Nout;
            const int nn = n_ok ? n : 0;
            const uint32_t tcol = slot_col + lane_addr + (uint32_t)mt * 128u;
            const float bias_u = p.bias ? __ldg(p.bias + ((long long)cloud * p.n_groups + g_tile) * p.bias_group_stride + nn) : 0.f;
            float mean_t = 0.f;
            if (MODE & TL_STATS) {
                float s = 0.f;
                for (int c0 = 0; c0 < valid; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(tcol + (uint32_t)c0, v);
                    tmem_wait_ld();
                    const bool full = c0 + 32 <= valid;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (full || c0 + j < valid) s += __uint_as_float(v[j]) + bias_u;
                }
                mean_t = s / (float)valid;
                if (n_ok) p.part_sum[tile * Nout + n] = s;
            }
            float osc = 1.f, osh = 0.f, msc = 0.f, msh = 0.f, mmu = 0.f, mis = 0.f;
            if (MODE & TL_AFFINE) { osc = __ldg(p.out_scale + nn); osh = __ldg(p.out_shift + nn); }
            if (MODE & TL_MASK) {
                msc = __ldg(p.mask_scale + nn); msh = __ldg(p.mask_shift + nn); mmu = __ldg(p.mask_mean + nn);
                mis = p.mask_invstd ? __ldg(p.mask_invstd + nn) : 0.f;
            }
            float q = 0.f, s2 = 0.f, q2 = 0.f;
            float vmax = -INFINITY, vmin = INFINITY;
            int rmax = 0, rmin = 0;
            const bool store = p.Y != nullptr && n_ok;
            const float floor_v = ((MODE & TL_AFFINE) && p.out_relu) ? 0.f : -INFINITY;
            auto chunk = [&](int c0, auto full_tag) {
                constexpr bool full = decltype(full_tag)::value;
                uint32_t v[32];
                tmem_ld32(tcol + (uint32_t)c0, v);
                float ym[32], yo[32];
                if (MODE & TL_MASK) {
                    const float* __restrict__ My = p.mask_y + (row_base + c0) * p.ld_mask + nn;
#pragma unroll
                    for (int j = 0; j < 32; ++j) { ym[j] = (full || c0 + j < valid) ? __ldg(My) : 0.f; My += p.ld_mask; }
                }
                if (MODE & TL_ACC) {
                    const float* __restrict__ Yo = p.Y + (row_base + c0) * p.ldy + nn;
#pragma unroll
                    for (int j = 0; j < 32; ++j) { yo[j] = (full || c0 + j < valid) ? *Yo : 0.f; Yo += p.ldy; }
                }
                tmem_wait_ld();
                float* __restrict__ yp = p.Y ? p.Y + (row_base + c0) * p.ldy + nn : nullptr;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int r = c0 + j;
                    const bool ok = full || r < valid;
                    float x = __uint_as_float(v[j]) + bias_u;
                    if (MODE & TL_ACC) x += yo[j];
                    if (MODE & TL_STATS) { const float d = ok ? x - mean_t : 0.f; q = fmaf(d, d, q); }
                    if (MODE & TL_POOL2) {
                        if (ok && x > vmax) { vmax = x; rmax = r; }
                        if (ok && x < vmin) { vmin = x; rmin = r; }
                    }
                    if (MODE & TL_AFFINE) x = fmaxf(fmaf(x, osc, osh), floor_v);
                    if (MODE & TL_MASK) {
                        float dz = x;
                        if (MODE & TL_DROP)
                            dz *= dropout_keep(p.out_drop_seed, (unsigned long long)(row_base + r) * Nout + n, p.out_drop_p);
                        dz = (ok && fmaf(ym[j] - mmu, msc, msh) > 0.f) ? dz : 0.f;
                        s2 += dz;
                        q2 = fmaf(dz, (ym[j] - mmu) * mis, q2);
                        x = dz;
                    }
                    if ((MODE & TL_POOL1) && ok && x > vmax) { vmax = x; rmax = r; }
                    if (store && ok) *yp = x;
                    yp += p.ldy;
                }
            };
            for (int c0 = 0; c0 < valid; c0 += 32) {
                if (c0 + 32 <= valid) chunk(c0, std::true_type{});
                else chunk(c0, std::false_type{});
            }
            if (n_ok) {
                if (MODE & TL_STATS) p.part_sq[tile * Nout + n] = q;
                if ((MODE & TL_MASK) && p.part_sum) { p.part_sum[tile * Nout + n] = s2; p.part_sq[tile * Nout + n] = q2; }
                if (MODE & (TL_POOL1 | TL_POOL2)) {
                    const unsigned long long kmax = ((unsigned long long)ordered_bits(vmax) << 32) | (0xffffffffu - (unsigned)(row0 + rmax));
                    atomicMax(p.pool_max + (long long)cloud * Nout + n, kmax);
                    if (MODE & TL_POOL2) {
                        const unsigned long long kmin = ((unsigned long long)(~ordered_bits(vmin)) << 32) | (0xffffffffu - (unsigned)(row0 + rmin));
                        atomicMax(p.pool_min + (long long)cloud * Nout + n, kmin);
                    }
                }
            }
        }
        // accumulator reads done before this slot's next MMA overwrites the columns
        tc_fence_before();
        wg_bar_sync(wg);
    };

    if (!per_cloud_w) {
        stage_weights(0);
        __syncthreads();
        const int n_tiles = p.n_clouds * tpc;
        for (int tile = blockIdx.x * 2 + wg; tile < n_tiles; tile += gridDim.x * 2) {
            const int nxt = tile + gridDim.x * 2;
            if (nxt < n_tiles) prefetch_tile(nxt / tpc, nxt % tpc);
            process(tile / tpc, tile % tpc);
        }
    } else {
        for (int cloud = blockIdx.x; cloud < p.n_clouds; cloud += gridDim.x) {
            __syncthreads();                  // every MMA that read the previous cloud's weights has been waited for
            stage_weights(cloud);
            __syncthreads();
            for (int t = wg; t < tpc; t += 2) {
                if (t + 2 < tpc) prefetch_tile(cloud, t + 2);
                process(cloud, t);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

// Large-row path of pw_linear(): returns 1 when the launch was taken, 0 when the shape is not eligible, < 0 on error.
int tc_layer_try(const PwParams& p, cudaStream_t st) {
    if (path_disabled("tc_layer")) return 0;
    const long long total_rows = (long long)p.n_clouds * p.rows_per_cloud;
    const int Mpad = (p.Nout + 127) / 128 * 128;
    if (total_rows < 2048 || p.K % 16 || p.K < 16 || p.K > 256 || p.Nout > 256 || (long long)Mpad * p.K > TL_MAX_WELEMS) return 0;
    if (p.x_transposed || p.y_transposed || p.ldx % 4 || (reinterpret_cast<uintptr_t>(p.X) & 15)) return 0;
    if (p.X2 && (reinterpret_cast<uintptr_t>(p.X2) & 15)) return 0;
    if (p.group_rows && p.n_groups > 64) return 0;
    if (p.pool_mode && !p.pool_max) return 0;
    const TlPlan sp = tl_plan(Mpad, p.K);
    if (sp.total > TL_MAX_SMEM) return 0;
    if (p.bias && p.group_rows && !p.groups_tile_aligned) return 0;
    int mode = 0;
    if (p.part_sum && !p.mask_y) mode |= TL_STATS;
    if (p.pool_mode == 2) mode |= TL_POOL2;
    if (p.pool_mode == 1) mode |= TL_POOL1;
    if (p.out_scale) mode |= TL_AFFINE;
    if (p.mask_y) mode |= TL_MASK;
    if (p.mask_y && p.out_drop_p > 0.f) mode |= TL_DROP;
    if (p.accumulate && p.Y) mode |= TL_ACC;
    if (!p.out_scale && p.out_relu) return 0;
    if (!p.mask_y && p.out_drop_p > 0.f) return 0;
    PwParams q = p;
    if (q.n_groups < 1) q.n_groups = 1;
    const long long n_tiles = (long long)p.n_clouds * ((p.rows_per_cloud + TL_ROWS - 1) / TL_ROWS);
    long long grid = p.w_cloud_stride ? p.n_clouds : (n_tiles + 1) / 2;
    if (grid > kNumSMs) grid = kNumSMs;
    const int smem_bytes = sp.total < TL_MIN_SMEM ? TL_MIN_SMEM : sp.total;
    switch (mode) {
#define TL_CASE(M) case M: { \
        static bool attr_set = false; \
        if (!attr_set) { \
            cudaError_t e = cudaFuncSetAttribute(tc_layer_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, TL_MAX_SMEM); \
            if (e != cudaSuccess) return fail(AMP_E_CUDA, "tc_layer: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); \
            attr_set = true; \
        } \
        tc_layer_kernel<M><<<(int)grid, TL_THREADS, smem_bytes, st>>>(q, Mpad); \
        break; }
        TL_CASE(0) TL_CASE(TL_ACC) TL_CASE(TL_STATS) TL_CASE(TL_STATS | TL_POOL2) TL_CASE(TL_AFFINE) TL_CASE(TL_AFFINE | TL_POOL1)
        TL_CASE(TL_MASK) TL_CASE(TL_MASK | TL_ACC) TL_CASE(TL_MASK | TL_DROP)
#undef TL_CASE
        default: return 0;        // an epilogue combination without a specialisation: CUDA-core path
    }
    count_launch();
    const int rc = check_launch("tc_layer_kernel");
    return rc == AMP_OK ? 1 : rc;
}

}  // namespace amp
