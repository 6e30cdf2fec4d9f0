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
// CTA = 512 threads = 2 independent slots of 256 threads (own raw / B-operand chunk buffers, 256 TMEM columns, mbarrier,
// tile stream); weights (hi + lo, K-major no-swizzle core-matrix layout) resident in shared memory for all tiles of the
// CTA; K walked in chunks of 64 (32 / 16 when the buffers would not fit next to the weights):
//   cp.async raw fp32 chunk (issued one chunk ahead: in flight during the MMAs and the epilogue of the previous one)
//   -> prologue + bf16 hi / lo split into the B operand (warp = 8-channel K group, lanes = rows: conflict free)
//   -> 3 MMAs per 16-wide K step, issued by one elected lane -> commit -> mbarrier wait -> epilogue (MODE).
#include <stdlib.h>

#include <type_traits>

#include <cuda_fp16.h>

#include <algorithm>

#include "nn_common.cuh"
#include "tc_ptx.cuh"
#include "tc_ts.cuh"

namespace amp {
namespace {
using namespace tcx;

constexpr int TL_THREADS = 512, TL_SLOT = 256, TL_ROWS = 128;
constexpr int TL_MAX_WELEMS = 32768;                 // Mpad * K
constexpr int TL_MAX_SMEM = 232448, TL_MIN_SMEM = 120 * 1024;

// Shared memory: weights (hi, lo) | per slot: B operand chunk (hi, lo) | per slot: raw fp32 chunk(s) filled by cp.async |
// prologue tables | exchange | barriers. `kcw` = input channels per chunk (64, 32 or 16: the widest that fits), `x2` = the
// prologue reads a second tensor (BatchNorm backward), which gets its own raw buffer.
struct TlPlan { int w_lo, b0, bhalf, raw0, rawsz, tab, exch, bar, total; };
__host__ __device__ inline TlPlan tl_plan(int Mpad, int K, int kcw, int x2, int ts) {
    TlPlan s;
    const int wbytes = ts ? 0 : Mpad * K * 2;  // ts: the weights live in tensor memory
    s.w_lo = wbytes;
    s.b0 = 2 * wbytes;
    s.bhalf = TL_ROWS * kcw * 2;                // bytes of the hi (or lo) half of one B chunk
    s.raw0 = s.b0 + 4 * s.bhalf;
    s.rawsz = TL_ROWS * (kcw * 4 + 16);         // one raw fp32 chunk; rows padded by 16 B (conflict-free column reads)
    s.tab = s.raw0 + 2 * (1 + x2) * s.rawsz;
    s.exch = s.tab + 16 * K;                    // [slot][which][half][128] floats: cross-warpgroup sums of the epilogue
    s.bar = s.exch + 2 * 3 * 2 * 128 * 4;
    s.total = s.bar + 64;
    return s;
}

// v = hi + lo (+ residual): one packed convert per pair.
//   H16 == false: bf16 terms (8 + 8 mantissa bits, 2^-17 residual, fp32 exponent range): gradients of any magnitude
//   H16 == true:  fp16 terms (11 + 11 bits, 2^-23 residual = fp32 class; saturating converts): the TRAINING FORWARD, whose
//                 operands are BatchNorm-normalised activations and weights (|x| << 65504; a lo term that underflows to an
//                 fp16 subnormal costs <= 3e-8 absolute). The forward needs this: max-pool winners closer than the forward
//                 error flip, and one flipped winner moves a gradient by percents (tests/test_nn_backward_tc_gpu.py).
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <bool H16>
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    if (H16) {
        hi = pack_f16x2_sat(a, b);
        const __half2 h = *reinterpret_cast<const __half2*>(&hi);
        lo = pack_f16x2_sat(a - __low2float(h), b - __high2float(h));
    } else {
        hi = pack_bf16x2(a, b);
        lo = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
    }
}
template <bool H16>
__device__ __forceinline__ void split_store8(const float (&v)[8], uint4* hi_dst, uint4* lo_dst) {
    uint4 h, l;
    split_pair<H16>(v[0], v[1], h.x, l.x); split_pair<H16>(v[2], v[3], h.y, l.y);
    split_pair<H16>(v[4], v[5], h.z, l.z); split_pair<H16>(v[6], v[7], h.w, l.w);
    *hi_dst = h; *lo_dst = l;
}
// instruction descriptor, kind::f16 with fp16 operands (a_format = b_format = 0), D = fp32, both K-major
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// out of line: the hash is ~25 instructions and only the head's dropout layers use it (code size = cold-start time here)
__device__ __noinline__ float dropout_keep_ool(unsigned long long seed, unsigned long long idx, float p) { return dropout_keep(seed, idx, p); }

// epilogue specialisations (bit mask): compile-time so that the per-element loop carries no dead branches
enum { TL_STATS = 1, TL_POOL2 = 2, TL_AFFINE = 4, TL_POOL1 = 8, TL_MASK = 16, TL_DROP = 32, TL_ACC = 64, TL_FP16 = 128 };

// Debug bisection of the phase timeline (AMP_TL_DBG, only read when AMP_LAYER_PROF passes a profile buffer; results are
// garbage): 1 = no convert, 2 = no cp.async, 4 = no epilogue, 8 = slot 0 only. profiles/r02_tc_layer_phase_bisect.txt
__device__ int g_tl_dbg = 0;

template <int MODE>
__global__ void __launch_bounds__(TL_THREADS, 1) tc_layer_kernel(const __grid_constant__ PwParams p, const int Mpad, const int kcw, const int cloud_split, const int ts, long long* prof_buf) {
    pdl_trigger();
    extern __shared__ __align__(1024) unsigned char smem[];
    // two slots of 256 threads (two warpgroups each): `wg` = slot, `sub` = which warpgroup of the slot, `wtid` = thread in slot
    const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31, wg = warp >> 3, sub = (warp >> 2) & 1, wtid = tid & (TL_SLOT - 1);
    const int K = p.K, Nout = p.Nout, rows = p.rows_per_cloud;
    const bool has_x2 = p.X2 != nullptr;
    const int dbg = prof_buf ? g_tl_dbg : 0;
    const TlPlan sp = tl_plan(Mpad, K, kcw, has_x2 ? 1 : 0, ts);
    float* s_exch = reinterpret_cast<float*>(smem + sp.exch) + wg * (3 * 2 * 128);
    auto slot_sync = [&]() { asm volatile("bar.sync %0, 256;" ::"r"(wg + 1) : "memory"); };
    __nv_bfloat16* s_whi = reinterpret_cast<__nv_bfloat16*>(smem);
    __nv_bfloat16* s_wlo = reinterpret_cast<__nv_bfloat16*>(smem + sp.w_lo);
    unsigned char* s_bhi = smem + sp.b0 + wg * 2 * sp.bhalf;
    unsigned char* s_blo = s_bhi + sp.bhalf;
    unsigned char* s_raw = smem + sp.raw0 + wg * (has_x2 ? 2 : 1) * sp.rawsz;      // this slot's raw chunk (then the X2 chunk)
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
    pdl_wait();                                                 // first global-memory access below
    for (int k = tid; k < K; k += TL_THREADS) {
        s_a[k] = p.in_a ? __ldg(p.in_a + k) : 1.f;
        s_b[k] = p.in_b ? __ldg(p.in_b + k) : 0.f;
        s_c[k] = p.in_c ? __ldg(p.in_c + k) : 0.f;
        s_m[k] = p.in_m ? __ldg(p.in_m + k) : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*s_tmem);

    const int tpc = (rows + TL_ROWS - 1) / TL_ROWS;
    const int n_mt = Mpad >> 7;
    const int lrow = (warp & 3) * 32 + lane;                     // staging: tile row; epilogue: channel lane
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t slot_col = tmem_base + (uint32_t)(wg * 256);
    const uint32_t whi_addr = smem_u32(s_whi), wlo_addr = smem_u32(s_wlo), bhi_addr = smem_u32(s_bhi), blo_addr = smem_u32(s_blo);
    const uint32_t w_lbo = (uint32_t)Mpad * 16u;
    constexpr bool H16 = (MODE & TL_FP16) != 0;
    const uint32_t idesc = H16 ? umma_idesc_f16(128, 128) : umma_idesc(128, 128);
    const bool per_cloud_w = p.w_cloud_stride != 0;
    const bool has_pro = p.in_a != nullptr;
    uint32_t phase = 0;

    // weights -> (hi, lo) core-matrix layout: one thread per (channel n, 8 consecutive k) = one 16-byte row of a core matrix
    // ts (Mpad == 128, K <= 256): the weights (the A operand: M = output channels) live in TENSOR MEMORY, columns
    // [128, 128 + K / 2) = hi and [384, 384 + K / 2) = lo of the allocation (the two slots' accumulators sit at [0, 128) and
    // [256, 384)): lane = output channel, one 32-bit column = two consecutive input channels. The MMAs then read only the
    // activation operand from shared memory -- in the SS form a 128 x 128 x 16 MMA reads 8 KB of operands per 67 cycles, i.e.
    // the whole shared-memory bandwidth, shared with every copy / conversion of the other slot -- and the weights' shared
    // memory (up to 128 KB) goes to wider input chunks (64 instead of 32 channels at K = 256).
    constexpr uint32_t kWHiCol = 128u, kWLoCol = 384u;
    auto stage_weights_tmem = [&](int cloud) {
        const float* __restrict__ W = p.W + (long long)cloud * p.w_cloud_stride;
        const int n = (warp & 3) * 32 + lane;                     // this thread's TMEM lane = output channel
        const int kq = K >> 2, k0 = (warp >> 2) * kq;             // the four warps of a lane quadrant split K
        const bool n_ok = n < Nout;
        const bool wvec = p.w_kn == 0 && (p.ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
#pragma unroll 1
        for (int c0 = 0; c0 < kq; c0 += 16) {
            float v[16];
            const int k = k0 + c0;
            if (!n_ok) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 0.f;
            } else if (wvec) {
                const float4* src = reinterpret_cast<const float4*>(W + (long long)n * p.ldw + k);
#pragma unroll
                for (int i = 0; i < 4; ++i) { const float4 a = __ldg(src + i); v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w; }
            } else if (p.w_kn == 0) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __ldg(W + (long long)n * p.ldw + k + i);
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __ldg(W + (long long)(k + i) * p.ldw + n);
            }
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) split_pair<H16>(v[2 * i], v[2 * i + 1], hi[i], lo[i]);
            tmem_st8(tmem_base + lane_addr + kWHiCol + (uint32_t)(k >> 1), hi);
            tmem_st8(tmem_base + lane_addr + kWLoCol + (uint32_t)(k >> 1), lo);
        }
        tmem_wait_st();
        tc_fence_before();
    };
    auto stage_weights = [&](int cloud) {
        if (ts) { stage_weights_tmem(cloud); return; }
        const float* __restrict__ W = p.W + (long long)cloud * p.w_cloud_stride;
        const int k8n = K >> 3, total = Mpad * k8n;
        uint4* whi4 = reinterpret_cast<uint4*>(s_whi);
        uint4* wlo4 = reinterpret_cast<uint4*>(s_wlo);
        const bool wvec = p.w_kn == 0 && (p.ldw & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
        // WI items per thread and pass, their loads issued together (a 256 x 128 layer is 8 items per thread: every pass
        // is a memory round trip in the kernel's prologue)
        constexpr int WI = 4;
#pragma unroll 1
        for (int e0 = tid; e0 < total; e0 += WI * TL_THREADS) {
            float v[WI][8];
            int dst[WI];
#pragma unroll
            for (int u = 0; u < WI; ++u) {
                const int e = e0 + u * TL_THREADS;
                int n, k8;
                if (p.w_kn == 0) { n = e / k8n; k8 = e - n * k8n; } else { n = e & (Mpad - 1); k8 = e / Mpad; }
                dst[u] = e < total ? k8 * Mpad + n : -1;
#pragma unroll
                for (int i = 0; i < 8; ++i) v[u][i] = 0.f;
                if (e < total && n < Nout) {
                    if (wvec) {
                        const float4* src = reinterpret_cast<const float4*>(W + (long long)n * p.ldw + k8 * 8);
                        const float4 a = __ldg(src), c = __ldg(src + 1);
                        v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w; v[u][4] = c.x; v[u][5] = c.y; v[u][6] = c.z; v[u][7] = c.w;
                    } else if (p.w_kn == 0) {
                        const float* src = W + (long long)n * p.ldw + k8 * 8;
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[u][i] = __ldg(src + i);
                    } else {
                        const float* src = W + (long long)(k8 * 8) * p.ldw + n;
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[u][i] = __ldg(src + (long long)i * p.ldw);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < WI; ++u)
                if (dst[u] >= 0) split_store8<H16>(v[u], whi4 + dst[u], wlo4 + dst[u]);
        }
        fence_proxy_async();
    };

    // Input staging is asynchronous: the slot copies a raw fp32 chunk into shared memory with cp.async (coalesced: the lanes of
    // a half warp read the contiguous bytes of one row), converts it into the bf16 hi / lo operand, and issues the copy of the
    // NEXT chunk (same tile, or the first chunk of this slot's next tile) right away, so that it is in flight during the MMAs
    // and the epilogue of the current one. (One TMA bulk copy per row instead of cp.async measured slower: the per-lane
    // addresses serialise the issue.)
    //   copy:    thread = (channel quad q, rows rsub + rstep * i)
    //   convert: warp = one 8-channel K group (and a block of rows), lanes = consecutive rows: 16-byte operand stores of a
    //            warp are contiguous, raw reads are conflict free thanks to the 16-byte row padding
    const int qn = kcw >> 2, q = wtid & (qn - 1), rsub = wtid / qn, rstep = TL_SLOT / qn, n_rit = TL_ROWS / rstep;
    const int raw_ld = kcw * 4 + 16;                                  // bytes per raw row
    const int ng = kcw >> 3, cg = (wtid >> 5) & (ng - 1), crow0 = ((wtid >> 5) / ng) * (16 * ng) + lane, n_cit = ng >> 1;
    const uint32_t raw_addr = smem_u32(s_raw);
    auto issue_chunk = [&](int cloud, int t, int kc) {
        const int row0 = t * TL_ROWS, valid = min(TL_ROWS, rows - row0);
        const int k = kc * kcw + q * 4;
        if (k < K && !(dbg & 2)) {
            const long long row_base = (long long)cloud * rows + row0;
            const float* __restrict__ xb = p.X + row_base * p.ldx + k;
            const float* __restrict__ x2b = has_x2 ? p.X2 + row_base * p.ldx + k : nullptr;
#pragma unroll 1
            for (int i = 0; i < n_rit; ++i) {
                const int r = rsub + rstep * i;
                const bool ok = r < valid;
                const uint32_t dst = raw_addr + (uint32_t)(r * raw_ld + q * 16);
                cp_async16(dst, ok ? xb + (long long)r * p.ldx : p.X, ok ? 16u : 0u);
                if (has_x2) cp_async16(dst + (uint32_t)sp.rawsz, ok ? x2b + (long long)r * p.ldx : p.X2, ok ? 16u : 0u);
            }
        }
        cp_async_commit();
    };

    int pi = 0;
    const bool prof = prof_buf != nullptr && blockIdx.x == 0 && tid == 0;
#define TL_PROF() do { if (prof && pi < 250) prof_buf[pi++] = clock64(); } while (0)
    auto process = [&](int cloud, int t, bool have_next, int cn, int tn) {
        TL_PROF();                                             // tile start
        const int row0 = t * TL_ROWS;
        const int valid = min(TL_ROWS, rows - row0);
        const long long row_base = (long long)cloud * rows + row0;        // global row of tile row 0
        const long long tile = (long long)cloud * tpc + t;
        // ---------------- K chunks: stage B operand (hi, lo), MMA ----------------
        const bool row_ok = lrow < valid;
        if (sub == 0 && row_ok && (MODE & (TL_MASK | TL_ACC))) {          // rows the epilogue of THIS tile will read: start them towards L2 now
            if (MODE & TL_MASK) {
                const char* m = reinterpret_cast<const char*>(p.mask_y + (row_base + lrow) * p.ld_mask);
#pragma unroll 1
                for (int b = 0; b < Nout * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(m + b));
            }
            if (MODE & TL_ACC) {
                const char* y = reinterpret_cast<const char*>(p.Y + (row_base + lrow) * p.ldy);
#pragma unroll 1
                for (int b = 0; b < Nout * 4; b += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(y + b));
            }
        }
        for (int kc = 0; kc * kcw < K; ++kc) {
            const int kcur = min(kcw, K - kc * kcw);
            {
                // convert the raw chunk: prologue (BatchNorm / BatchNorm backward, ReLU, dropout), split into bf16 hi + lo;
                // a thread turns 8 consecutive channels of a row into one 16-byte K-group row of each operand half
                cp_async_wait_all();
                slot_sync();                                             // every thread's part of the raw chunk has landed
                const bool g_ok = cg * 8 < kcur && !(dbg & 1);
                const int k = kc * kcw + (g_ok ? cg * 8 : 0);
                if (g_ok) {
                    float ca[8], cb[8], cm[8], cc[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { ca[j] = s_a[k + j]; cb[j] = s_b[k + j]; cm[j] = s_m[k + j]; cc[j] = s_c[k + j]; }
                    uint4* dhi = reinterpret_cast<uint4*>(s_bhi) + cg * TL_ROWS;
                    uint4* dlo = reinterpret_cast<uint4*>(s_blo) + cg * TL_ROWS;
#pragma unroll 1
                    for (int i = 0; i < n_cit; ++i) {
                        const int r = crow0 + 32 * i;
                        const float4* rx = reinterpret_cast<const float4*>(s_raw + r * raw_ld + cg * 32);
                        const float4 x0 = rx[0], x1 = rx[1];
                        float v[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
                        if (has_pro) {
                            if (has_x2) {
                                const float4* ry = reinterpret_cast<const float4*>(s_raw + sp.rawsz + r * raw_ld + cg * 32);
                                const float4 y0 = ry[0], y1 = ry[1];
                                const float y[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = fmaf(y[j] - cm[j], cc[j], fmaf(v[j], ca[j], cb[j]));
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j] - cm[j], ca[j], cb[j]);
                            }
                        }
                        if (p.in_relu) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                        }
                        if (p.in_drop_p > 0.f) {
                            const unsigned long long di = (unsigned long long)(row_base + r) * K + k;
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] *= dropout_keep_ool(eff_seed(p.in_drop_seed, p.drop_off), di + j, p.in_drop_p);
                        }
                        if (r >= valid) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[j] = 0.f;
                        }
                        uint4 h4, l4;
                        split_pair<H16>(v[0], v[1], h4.x, l4.x); split_pair<H16>(v[2], v[3], h4.y, l4.y);
                        split_pair<H16>(v[4], v[5], h4.z, l4.z); split_pair<H16>(v[6], v[7], h4.w, l4.w);
                        dhi[r] = h4; dlo[r] = l4;
                        if (p.split_dump) {                    // [tile][hi, lo][K / 8][128 rows]: a warp writes 512 contiguous bytes
                            uint4* g = reinterpret_cast<uint4*>(p.split_dump) + ((tile * 2) * (long long)(K >> 3) + (kc * ng + cg)) * TL_ROWS + r;
                            g[0] = h4;
                            g[(long long)(K >> 3) * TL_ROWS] = l4;
                        }
                    }
                }
            }
            TL_PROF();                                         // chunk staged
            fence_proxy_async();
            tc_fence_before();
            slot_sync();
            if ((warp & 7) == 0) {                         // first warp of the slot; one elected lane issues (uniform descriptors)
                tc_fence_after();
                if (elect_one_sync()) {
                for (int mt = 0; mt < n_mt; ++mt) {
                    const uint32_t d = slot_col + (uint32_t)mt * 128u;
                    for (int ks = 0; ks < (kcur >> 4); ++ks) {
                        const uint32_t kg = (uint32_t)(kc * (kcw >> 4) + ks);
                        const uint32_t woff = (uint32_t)mt * 2048u + kg * 2u * w_lbo;
                        const uint64_t a_hi = umma_desc(whi_addr + woff, w_lbo, 128u), a_lo = umma_desc(wlo_addr + woff, w_lbo, 128u);
                        const uint64_t b_hi = umma_desc(bhi_addr + (uint32_t)ks * 4096u, 2048u, 128u);
                        const uint64_t b_lo = umma_desc(blo_addr + (uint32_t)ks * 4096u, 2048u, 128u);
                        if (ts) {
                            const uint32_t t_hi = tmem_base + kWHiCol + kg * 8u, t_lo = tmem_base + kWLoCol + kg * 8u;
                            umma_bf16_ts(d, t_lo, b_hi, idesc, (kc | ks) != 0);
                            umma_bf16_ts(d, t_hi, b_lo, idesc, 1u);
                            umma_bf16_ts(d, t_hi, b_hi, idesc, 1u);
                        } else {
                            umma_bf16(d, a_lo, b_hi, idesc, (kc | ks) != 0);
                            umma_bf16(d, a_hi, b_lo, idesc, 1u);
                            umma_bf16(d, a_hi, b_hi, idesc, 1u);
                        }
                    }
                }
                umma_commit(mbar);
                }
                __syncwarp();
            }
            // the raw chunk is consumed by the whole slot (barrier above): start the copy of the next one
            if ((kc + 1) * kcw < K) issue_chunk(cloud, t, kc + 1);
            else if (have_next) issue_chunk(cn, tn, 0);
            TL_PROF();                                         // MMAs issued
            __syncwarp();
            // ONE thread of the slot polls the mbarrier, the other 255 block in the named barrier (a hardware wait): no
            // try_wait traffic of 8 warps on the shared-memory pipe while the MMAs fetch their operand through it
            if ((warp & 7) == 0) {
                if (lane == 0) mbar_wait(mbar, phase);
                __syncwarp();
            }
            slot_sync();
            phase ^= 1u;
            tc_fence_after();
            TL_PROF();                                         // MMAs complete
        }

        // ---------------- epilogue: lane = output channel, columns = rows of the tile ----------------
        int g_tile = 0;                                        // row group of the tile (groups are tile aligned)
        if (p.bias && p.group_rows)
            for (int q = 1; q < p.n_groups; ++q) g_tile += (row0 >= __ldg(p.group_rows + q)) ? 1 : 0;
        for (int mt = 0; mt < ((dbg & 4) ? 0 : n_mt); ++mt) {
            const int n = mt * 128 + lrow;
            const bool n_ok = n < Nout;
            const int nn = n_ok ? n : 0;
            const uint32_t tcol = slot_col + lane_addr + (uint32_t)mt * 128u;
            const float bias_u = p.bias ? __ldg(p.bias + ((long long)cloud * p.n_groups + g_tile) * p.bias_group_stride + nn) : 0.f;
            // The loops below are ROLLED over 8-column pieces (tcgen05.ld x8): these kernels run two tiles per slot, so the
            // first pass through the code is an instruction-cache miss stream; 4x less code beats 4x fewer loop branches.
            // this warpgroup's half of the tile rows: accumulator columns [c_lo, c_hi); sums are combined through shared memory
            const int c_lo = sub * 64, c_hi = min(valid, c_lo + 64);
            // BatchNorm statistics in ONE pass over the accumulator: sums of (x - shift) and (x - shift)^2 with shift = the first
            // value of this thread's rows (a sample of the same distribution, so nothing cancels), turned into the tile's sum
            // and centred sum of squares when the two row halves are merged (pairwise update of Chan et al.)
            float shift = 0.f, s1 = 0.f;
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
            const float* __restrict__ My = (MODE & TL_MASK) ? p.mask_y + row_base * p.ld_mask + nn : nullptr;
            float* __restrict__ yp = p.Y ? p.Y + row_base * p.ldy + nn : nullptr;
            const int ldy32 = (int)p.ldy, ldm32 = (int)p.ld_mask;     // tile-local offsets fit 32 bits: one IMAD per element, not a 64-bit multiply
            // software pipeline: the TMEM load (and the accumulate target) of piece c0 + 8 is in flight while piece c0 is
            // processed; the saved activations that feed the ReLU mask / BatchNorm-backward sums are TWO pieces ahead (two
            // buffers, the loop body is instantiated twice so that no pending register is ever moved): their ~1.8 k cycles of
            // latency were half of the backward epilogue with a one-piece lead
            uint32_t vn[8];
            float yon[8], ymA[8], ymB[8];
            auto load_mask = [&](float (&buf)[8], int c0) {
                const float* mq = My + c0 * ldm32;                      // running pointers: one 64-bit add per element
#pragma unroll
                for (int j = 0; j < 8; ++j, mq += ldm32) buf[j] = (n_ok && c0 + j < c_hi) ? __ldg(mq) : 0.f;
            };
            auto issue = [&](int c0) {
                tmem_ld8(tcol + (uint32_t)c0, vn);
                if (MODE & TL_ACC) {
                    const float* yq = yp + c0 * ldy32;
#pragma unroll
                    for (int j = 0; j < 8; ++j, yq += ldy32) yon[j] = (n_ok && c0 + j < valid) ? *yq : 0.f;
                }
            };
            auto piece = [&](int c0, float (&ymq)[8]) {
                uint32_t v[8];
                float ym[8], yo[8];
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j) { v[j] = vn[j]; yo[j] = yon[j]; ym[j] = (MODE & TL_MASK) ? ymq[j] : 0.f; }
                if (c0 + 8 < c_hi) issue(c0 + 8);
                if ((MODE & TL_MASK) && c0 + 16 < c_hi) load_mask(ymq, c0 + 16);
                float* ys = yp + c0 * ldy32;
#pragma unroll
                for (int j = 0; j < 8; ++j, ys += ldy32) {
                    const int r = c0 + j;
                    const bool ok = r < valid;
                    float x = __uint_as_float(v[j]) + bias_u;
                    if (MODE & TL_ACC) x += yo[j];
                    if (MODE & TL_STATS) {
                        if (j == 0 && c0 == c_lo) shift = x;
                        const float d = ok ? x - shift : 0.f;
                        s1 += d;
                        q = fmaf(d, d, q);
                    }
                    if (MODE & TL_POOL2) {
                        if (ok && x > vmax) { vmax = x; rmax = r; }
                        if (ok && x < vmin) { vmin = x; rmin = r; }
                    }
                    if (MODE & TL_AFFINE) x = fmaxf(fmaf(x, osc, osh), floor_v);
                    if (MODE & TL_MASK) {
                        float dz = x;
                        if (MODE & TL_DROP)
                            dz *= dropout_keep_ool(eff_seed(p.out_drop_seed, p.drop_off), (unsigned long long)(row_base + r) * Nout + n, p.out_drop_p);
                        dz = (ok && fmaf(ym[j] - mmu, msc, msh) > 0.f) ? dz : 0.f;
                        s2 += dz;
                        q2 = fmaf(dz, (ym[j] - mmu) * mis, q2);
                        x = dz;
                    }
                    if ((MODE & TL_POOL1) && ok && x > vmax) { vmax = x; rmax = r; }
                    if (store && ok) *ys = x;
                }
            };
            if (c_lo < c_hi) {
                issue(c_lo);
                if (MODE & TL_MASK) { load_mask(ymA, c_lo); load_mask(ymB, c_lo + 8); }
            }
            if (MODE & TL_MASK) {
#pragma unroll 1
                for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
                    piece(c0, ymA);
                    if (c0 + 8 < c_hi) piece(c0 + 8, ymB);
                }
            } else {
#pragma unroll 1
                for (int c0 = c_lo; c0 < c_hi; c0 += 8) piece(c0, ymA);
            }
            if (MODE & (TL_STATS | TL_MASK)) {                           // combine the two row halves in a fixed order
                s_exch[sub * 128 + lrow] = (MODE & TL_STATS) ? shift : 0.f;
                s_exch[256 + sub * 128 + lrow] = (MODE & TL_STATS) ? s1 : s2;
                s_exch[512 + sub * 128 + lrow] = (MODE & TL_STATS) ? q : q2;
                slot_sync();
                if (n_ok && sub == 0) {
                    if (MODE & TL_STATS) {
                        const float na = (float)min(valid, 64), nb = (float)(valid - min(valid, 64));
                        const float sa = s_exch[256 + lrow], sb = s_exch[256 + 128 + lrow];
                        const float sum_a = fmaf(s_exch[lrow], na, sa), sum_b = fmaf(s_exch[128 + lrow], nb, sb);
                        float m2 = s_exch[512 + lrow] - sa * sa / na;
                        if (nb > 0.f) {
                            const float delta = sum_b / nb - sum_a / na;
                            m2 += s_exch[512 + 128 + lrow] - sb * sb / nb + delta * delta * (na * nb / (na + nb));
                        }
                        p.part_sum[tile * Nout + n] = sum_a + sum_b;
                        p.part_sq[tile * Nout + n] = m2;
                    } else if (p.part_sum) {
                        p.part_sum[tile * Nout + n] = s_exch[256 + lrow] + s_exch[256 + 128 + lrow];
                        p.part_sq[tile * Nout + n] = s_exch[512 + lrow] + s_exch[512 + 128 + lrow];
                    }
                }
                slot_sync();                                             // exchange buffer free for the next M tile
            }
            if (n_ok && c_lo < c_hi) {
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
        TL_PROF();                                             // epilogue done
        tc_fence_before();
        slot_sync();
    };

    // One loop, one call site of stage_weights / process (they are big: a second inlined copy doubles the cold-start
    // instruction fetch). Shared weights: tiles strided over (CTA, slot). Per-cloud weights: a cloud is cut into cloud_split
    // units of consecutive tile pairs, units strided over CTAs (32 clouds fill 128 SMs instead of 32); the two slots take
    // alternating tiles of the unit and both pass the CTA barriers around the weight restaging.
    const int n_tiles = p.n_clouds * tpc;
    const int tps = (tpc + 1) >> 1;
    const int tpu = (tps + cloud_split - 1) / cloud_split;     // tile pairs per unit
    const int n_units = p.n_clouds * cloud_split;
    const int n_it = per_cloud_w ? ((n_units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * tpu
                                 : (n_tiles - (int)blockIdx.x * 2 + (int)gridDim.x * 2 - 1) / ((int)gridDim.x * 2);
    auto work = [&](int it, int& cloud, int& t) -> bool {
        if (per_cloud_w) {
            const int ci = it / tpu, j = it - ci * tpu;
            const int u = blockIdx.x + ci * gridDim.x;
            cloud = u / cloud_split;
            const int jj = (u - cloud * cloud_split) * tpu + j;
            t = jj * 2 + wg;
            return u < n_units && jj < tps && t < tpc && !((dbg & 8) && wg);
        }
        const int tile = blockIdx.x * 2 + wg + it * gridDim.x * 2;
        cloud = tile / tpc;
        t = tile - cloud * tpc;
        return tile < n_tiles && !((dbg & 8) && wg);
    };
    bool pending = false;                     // the raw chunk 0 of this slot's tile of iteration `it` is already in flight
#pragma unroll 1
    for (int it = 0; it < n_it; ++it) {
        int cloud, t, cn = 0, tn = 0;
        const bool have = work(it, cloud, t);
        if (have && !pending) issue_chunk(cloud, t, 0);
        if (per_cloud_w ? (it % tpu == 0) : (it == 0)) {
            __syncthreads();                  // every MMA that read the previous weights has been waited for
            stage_weights(per_cloud_w ? cloud : 0);
            __syncthreads();
        }
        const bool have_next = it + 1 < n_it && work(it + 1, cn, tn);
        if (have) { process(cloud, t, have_next, cn, tn); pending = have_next; }
    }
    if (prof) prof_buf[255] = pi;
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace

bool& tc_layer_dumped() {
    static thread_local bool flag = false;
    return flag;
}

// Large-row path of pw_linear(): returns 1 when the launch was taken, 0 when the shape is not eligible, < 0 on error.
int tc_layer_try(const PwParams& p, cudaStream_t st) {
    if (path_disabled("tc_layer")) return 0;
    const long long total_rows = (long long)p.n_clouds * p.rows_per_cloud;
    const int Mpad = (p.Nout + 127) / 128 * 128;
    // K: whole 64-channel chunks (every wide layer of the network has K = 64, 128 or 256; other widths take the CUDA-core kernel)
    if (total_rows < 2048 || p.K % 64 || p.K > 256 || p.Nout > 256 || (long long)Mpad * p.K > TL_MAX_WELEMS) return 0;
    if (p.x_transposed || p.y_transposed || p.ldx % 4 || (reinterpret_cast<uintptr_t>(p.X) & 15)) return 0;
    if (p.X2 && (reinterpret_cast<uintptr_t>(p.X2) & 15)) return 0;
    if (p.group_rows && p.n_groups > 64) return 0;
    if (p.pool_mode && !p.pool_max) return 0;
    const int ts = (Mpad == 128 && p.K <= 256 && !path_disabled("tc_layer_ts")) ? 1 : 0;   // weights in tensor memory (see the kernel)
    int kcw = 64;                              // widest input chunk whose buffers fit next to the weights
    while (kcw > 16 && tl_plan(Mpad, p.K, kcw, p.X2 ? 1 : 0, ts).total > TL_MAX_SMEM) kcw >>= 1;
    const TlPlan sp = tl_plan(Mpad, p.K, kcw, p.X2 ? 1 : 0, ts);
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
    if (p.fp16_split && !(mode & (TL_MASK | TL_ACC | TL_AFFINE))) mode |= TL_FP16;
    if (!p.out_scale && p.out_relu) return 0;
    if (!p.mask_y && p.out_drop_p > 0.f) return 0;
    PwParams q = p;
    if (q.n_groups < 1) q.n_groups = 1;
    const long long n_tiles = (long long)p.n_clouds * ((p.rows_per_cloud + TL_ROWS - 1) / TL_ROWS);
    // per-cloud weights: cut every cloud into units of tile pairs until the units fill the machine
    const int tpc_h = (p.rows_per_cloud + TL_ROWS - 1) / TL_ROWS, tps_h = (tpc_h + 1) / 2;
    int cloud_split = 1;
    if (p.w_cloud_stride) cloud_split = (int)std::max(1LL, std::min((long long)tps_h, (long long)kNumSMs / p.n_clouds));
    long long grid = p.w_cloud_stride ? (long long)p.n_clouds * cloud_split : (n_tiles + 1) / 2;
    if (grid > kNumSMs) grid = kNumSMs;
    const int smem_bytes = sp.total < TL_MIN_SMEM ? TL_MIN_SMEM : sp.total;
    static const bool want_prof = getenv("AMP_LAYER_PROF") != nullptr;       // debugging aid: phase timeline of CTA 0 / thread 0
    static long long* dprof = nullptr;
    if (want_prof && !dprof) cudaMalloc(&dprof, 256 * sizeof(long long));
    if (want_prof) cudaMemsetAsync(dprof, 0, 256 * sizeof(long long), st);
    if (want_prof) { const int d = getenv("AMP_TL_DBG") ? atoi(getenv("AMP_TL_DBG")) : 0; cudaMemcpyToSymbolAsync(g_tl_dbg, &d, sizeof d, 0, cudaMemcpyHostToDevice, st); }
    switch (mode) {
#define TL_CASE(M) case M: { \
        static bool attr_set = false; \
        if (!attr_set) { \
            cudaError_t e = cudaFuncSetAttribute(tc_layer_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, TL_MAX_SMEM); \
            if (e != cudaSuccess) return fail(AMP_E_CUDA, "tc_layer: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); \
            attr_set = true; \
        } \
        launch_pdl(tc_layer_kernel<M>, dim3((unsigned)((int)grid)), dim3(TL_THREADS), smem_bytes, st, q, Mpad, kcw, cloud_split, ts, want_prof ? dprof : nullptr); \
        break; }
        TL_CASE(0) TL_CASE(TL_ACC) TL_CASE(TL_STATS) TL_CASE(TL_STATS | TL_POOL2) TL_CASE(TL_AFFINE) TL_CASE(TL_AFFINE | TL_POOL1)
        TL_CASE(TL_MASK) TL_CASE(TL_MASK | TL_ACC) TL_CASE(TL_MASK | TL_DROP)
        TL_CASE(TL_FP16) TL_CASE(TL_FP16 | TL_STATS) TL_CASE(TL_FP16 | TL_STATS | TL_POOL2)
#undef TL_CASE
        default: return 0;        // an epilogue combination without a specialisation: CUDA-core path
    }
    if (want_prof) {
        long long h[256];
        cudaMemcpyAsync(h, dprof, sizeof h, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        fprintf(stderr, "[tc_layer prof] mode=%d K=%d N=%d rows=%lld kcw=%d ts=%d:", mode, p.K, p.Nout, (long long)p.n_clouds * p.rows_per_cloud, kcw, ts);
        for (int i = 1; i < (int)h[255] && i < 40; ++i) fprintf(stderr, " %lld", h[i] - h[i - 1]);
        fprintf(stderr, "\n");
    }
    count_launch();
    count_path("tc_layer");
    if (ts) count_path("tc_layer_ts");
    if (p.split_dump) tc_layer_dumped() = true;
    if (mode & TL_MASK) count_path("tc_layer_dgrad");
    const int rc = check_launch("tc_layer_kernel");
    return rc == AMP_OK ? 1 : rc;
}

}  // namespace amp
