// Fused shared-MLP chain kernel on tcgen05 tensor cores (see tc_chain.cuh for the contract).
//
// Layout of one CTA (512 threads, 1 CTA per SM):
//   * the packed bf16 weights of the whole chain live in shared memory for the CTA's lifetime (TMA bulk copies
//     global -> shared, completion on an mbarrier);
//   * two independent "slots", each served by 256 threads (two warpgroups): a 32 KB activation buffer (the K-major A
//     operand of the next layer), a region for per-cloud weights, 256 TMEM columns of fp32 accumulators and one
//     mbarrier. A slot walks its own stream of 128-point tiles: stage input -> for every layer { one thread issues the
//     tcgen05.mma K-steps (+ one more against a constant block of ones, whose other operand is the layer's bias K group,
//     read with a zero leading-dimension stride: the bias add costs no ALU work) and commits to the mbarrier; the 256 threads wait, read the accumulator with tcgen05.ld
//     (thread == TMEM lane == point row; the two warpgroups take alternating 32-column chunks), pack to bf16,
//     ReLU on the packed pairs, and write the next A operand }. While one slot is in its epilogue the other slot's
//     MMAs run, so tensor pipe and CUDA cores overlap without cross-slot synchronisation.
//   * pooled layers run TRANSPOSED: D^T[channel, point] = W[channel, :] . act[point, :], so that TMEM lanes are
//     channels and the max over the points of the tile is a per-thread register reduction (no shuffles); one
//     atomicMax per (cloud, channel, half tile).
// Operand layout in shared memory (both A and B): K-major, no swizzle, 8 x 16-byte core matrices:
//   element (row r, k) at byte ((k / 8) * rows + r) * 16 + (k % 8) * 2      -> SBO = 128 B, LBO = rows * 16 B.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_chain.cuh"
#include "tc_ptx.cuh"

namespace amp {
namespace {
using namespace tcx;

constexpr int kThreads = 512, kSlotThreads = 256;
constexpr int kActBytes = kTcTileRows * 128 * 2;          // 128 rows x up to 128 channels of bf16
constexpr int kOnesBytes = kTcTileRows * 16 * 2;          // constant A / B operand of the bias MMA: [2][128][8] bf16
constexpr int kTmemCols = 512;
constexpr int kSlotCols = 256;
constexpr int kMaxSmem = 232448;                          // 227 KB per CTA on sm_100
constexpr int kMinSmem = 120 * 1024;                      // more than half an SM: one CTA (one TMEM owner) per SM

__host__ __device__ inline int align_i(int v, int a) { return (v + a - 1) / a * a; }

struct SmemPlan { int w, wc, act0, ones, bar, total; };
__host__ __device__ inline SmemPlan smem_plan(int wblob_bytes, int wcloud_bytes) {
    SmemPlan s;
    s.w = align_i(wblob_bytes, 128);
    s.wc = align_i(wcloud_bytes, 128);
    s.act0 = s.w + 2 * s.wc;
    s.ones = s.act0 + 2 * kActBytes;
    s.bar = s.ones + kOnesBytes;
    s.total = s.bar + 64;
    return s;
}

__device__ __forceinline__ void slot_bar_sync(int slot) { asm volatile("bar.sync %0, 256;" ::"r"(slot + 1) : "memory"); }
// max(x, 0) on a packed bf16 pair
__device__ __forceinline__ uint32_t relu_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    v = __hmax2(v, __float2bfloat162_rn(0.f));
    return *reinterpret_cast<uint32_t*>(&v);
}

// Rare epilogue variants of one 32-column chunk (per-block bias of the head, fp32 store of the local features, logits):
// kept out of line so that their address arithmetic is not hoisted in front of every layer's hot loop.
__device__ __noinline__ void epilogue_slow(const TcChainParams& p, const TcOp& op, const uint32_t (&v)[32], uint32_t (&u)[16], const float* gb,
                                           int c0, int nc, bool row_ok, long long grow, int cloud, int row_in_cloud) {
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (gb) {
        const float4* g4 = reinterpret_cast<const float4*>(gb + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            if (q * 4 < nc) {
                const float4 b = __ldg(g4 + q);
                f[4 * q] += b.x; f[4 * q + 1] += b.y; f[4 * q + 2] += b.z; f[4 * q + 3] += b.w;
            }
        }
    }
    if (op.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (op.store_f32 && row_ok) {
        float4* o = reinterpret_cast<float4*>(p.out_f32 + grow * p.out_ld + p.out_col0 + c0);
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q * 4 < nc) o[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
    }
    if (op.store_logits && c0 == 0 && row_ok) {
        float* lp = p.logits + (long long)cloud * p.n_classes * p.rows_per_cloud + row_in_cloud;
#pragma unroll
        for (int n = 0; n < 32; ++n) {
            if (n < p.n_classes) *lp = f[n];
            lp += p.rows_per_cloud;
        }
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) u[q] = pack_bf16x2(f[2 * q], f[2 * q + 1]);
}

__global__ void __launch_bounds__(kThreads, 1) tc_chain_kernel(const __grid_constant__ TcChainParams p) {
    pdl_trigger();
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
    const int slot = warp >> 3, sub = (warp >> 2) & 1, stid = tid & (kSlotThreads - 1);
    const SmemPlan sp = smem_plan(p.wblob_bytes, p.wcloud_bytes);
    unsigned char* s_w = smem;
    unsigned char* s_wc = smem + sp.w + slot * sp.wc;
    unsigned char* s_act = smem + sp.act0 + slot * kActBytes;
    uint4* s_ones = reinterpret_cast<uint4*>(smem + sp.ones);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + sp.bar);       // [0] weights, [1 + slot] MMA completion
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + sp.bar + 32);
    const uint32_t wbar = smem_u32(&s_bar[0]), mbar = smem_u32(&s_bar[1 + slot]);

    if (tid == 0) {
        mbar_init(wbar, 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        mbar_init(smem_u32(&s_bar[2]), 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(s_tmem), kTmemCols);
    // ones block: K group 0 = [1, 1, 0, 0, 0, 0, 0, 0] per row (multiplies the bias hi and lo terms), K group 1 = 0
    if (tid < 256) s_ones[tid] = tid < 128 ? make_uint4(0x3f803f80u, 0u, 0u, 0u) : make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*s_tmem);
    pdl_wait();                                                  // first global-memory access below (the packed weights)
    if (tid == 0 && p.wblob_bytes > 0) {
        mbar_expect_tx(wbar, (uint32_t)p.wblob_bytes);
        for (int off = 0; off < p.wblob_bytes; off += 32768) {
            const int n = min(32768, p.wblob_bytes - off);
            bulk_g2s(smem_u32(s_w + off), p.wblob + off, (uint32_t)n, wbar);
        }
    }
    bool w_ready = p.wblob_bytes == 0;

    const int rows = p.rows_per_cloud;
    const int tiles_per_cloud = (rows + kTcTileRows - 1) / kTcTileRows;
    const int n_tiles = p.n_clouds * tiles_per_cloud;
    const int row = (warp & 3) * 32 + lane;                      // this thread's TMEM lane == tile row
    const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t slot_col = tmem_base + (uint32_t)(slot * kSlotCols);
    const uint32_t act_addr = smem_u32(s_act), ones_addr = smem_u32(s_ones);
    uint32_t phase = 0;
    int cur_cloud = -1;
    int pi = 0;
    const bool prof = p.prof != nullptr && blockIdx.x == 0 && stid == 0 && slot == 0;
#define AMP_PROF() do { if (prof && pi < 250) p.prof[pi++] = clock64(); } while (0)
    AMP_PROF();

    for (int tile = blockIdx.x * 2 + slot; tile < n_tiles; tile += gridDim.x * 2) {
        const int cloud = tile / tiles_per_cloud;
        const int row0 = (tile - cloud * tiles_per_cloud) * kTcTileRows;
        const int valid = min(kTcTileRows, rows - row0);
        const bool row_ok = row < valid;
        const long long grow = (long long)cloud * rows + row0 + row;

        // ---- per-cloud weights (all MMAs of the previous tile have completed: safe to overwrite) ----
        if (p.wcloud_bytes > 0 && cloud != cur_cloud) {
            const uint4* src = reinterpret_cast<const uint4*>(p.wcloud + (long long)cloud * p.wcloud_stride);
            uint4* dst = reinterpret_cast<uint4*>(s_wc);
            for (int i = stid; i < p.wcloud_bytes / 16; i += kSlotThreads) dst[i] = __ldg(src + i);
            cur_cloud = cloud;
        }
        // ---- input stage: fp32 rows -> bf16 A operand [K/8][128][8] ----
        if (p.in_mode == 0) {
            if (sub == 0) {
                const float* src = p.in_x + grow * p.in_ld;
                float xv[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) xv[j] = (row_ok && j < p.in_k) ? __ldg(src + j) : 0.f;
                float hi[9], lo[9];
#pragma unroll
                for (int j = 0; j < 9; ++j) {
                    const __nv_bfloat16 h = __float2bfloat16_rn(xv[j]);
                    hi[j] = __bfloat162float(h);
                    lo[j] = xv[j] - hi[j];
                }
                const float one = p.in_bias ? 1.f : 0.f;        // constant columns in_k, in_k + 1 multiply the folded bias
                uint4* dst = reinterpret_cast<uint4*>(s_act);
                if (p.op[0].K == 16) {          // in_k <= 6: k 0..7 = hi (+ ones), k 8..15 = lo
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (j == p.in_k || j == p.in_k + 1) hi[j] = one;
                    dst[0 * 128 + row] = make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(hi[4], hi[5]),
                                                    pack_bf16x2(hi[6], hi[7]));
                    dst[1 * 128 + row] = make_uint4(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]),
                                                    pack_bf16x2(lo[6], lo[7]));
                } else {                        // K == 32: k 0..15 = hi (9 used), k 16..31 = lo
                    dst[0 * 128 + row] = make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(hi[4], hi[5]),
                                                    pack_bf16x2(hi[6], hi[7]));
                    dst[1 * 128 + row] = make_uint4(pack_bf16x2(hi[8], one), pack_bf16x2(one, 0.f), 0u, 0u);   // k 9, 10 = ones
                    dst[2 * 128 + row] = make_uint4(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]),
                                                    pack_bf16x2(lo[6], lo[7]));
                    dst[3 * 128 + row] = make_uint4(pack_bf16x2(lo[8], 0.f), 0u, 0u, 0u);
                }
            }
        } else {
            // coalesced: 8 consecutive lanes read 128 contiguous bytes of one row; a lane's float4 is half of a 16-byte
            // K group of the operand, written as one 8-byte store
            const int K0 = p.op[0].K, f4_per_row = K0 >> 2, total = kTcTileRows * f4_per_row;
            const float* base = p.in_x + ((long long)cloud * rows + row0) * p.in_ld;
            for (int i = stid; i < total; i += kSlotThreads) {
                const int r = i / f4_per_row, q = i - r * f4_per_row;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < valid) a = __ldg(reinterpret_cast<const float4*>(base + (long long)r * p.in_ld) + q);
                uint2* dst = reinterpret_cast<uint2*>(s_act + ((q >> 1) * 128 + r) * 16 + (q & 1) * 8);
                *dst = make_uint2(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w));
            }
        }
        // group of this row (per-block bias of the segmentation head)
        int group = 0;
        if (p.n_groups > 1) {
            const int r = row0 + (row_ok ? row : 0);
            for (int g = 1; g < p.n_groups; ++g) group += (r >= __ldg(p.group_rows + g)) ? 1 : 0;
        }
        AMP_PROF();            // staged
        fence_proxy_async();
        tc_fence_before();
        slot_bar_sync(slot);
        AMP_PROF();            // after barrier

        for (int l = 0; l < p.n_ops; ++l) {
            const TcOp& op = p.op[l];
            // ---- MMA issue: the first warp of the slot enters, one elected lane issues (uniform descriptors) ----
            if ((warp & 7) == 0) {
                if (!w_ready) { mbar_wait(wbar, 0); w_ready = true; }
                tc_fence_after();
                if (elect_one_sync()) {
                // descriptors advance linearly with the K step: build them once, then one 64-bit add per MMA
                const uint32_t wreg = op.w_cloud ? smem_u32(s_wc) : smem_u32(s_w);
                const uint32_t w_lbo = (uint32_t)op.N * 16u;
                const int ksteps = op.K >> 4;
                const uint64_t act_d = umma_desc(act_addr, 2048u, 128u), act_step = 4096u >> 4;
                const uint64_t w_d = umma_desc(wreg + (uint32_t)op.w_off, w_lbo, 128u), w_step = (2u * w_lbo) >> 4;
                if (!op.pool) {
                    const uint32_t idesc = umma_idesc(128, op.N);
                    for (int k = 0; k < ksteps; ++k) umma_bf16(slot_col, act_d + k * act_step, w_d + k * w_step, idesc, k > 0);
                    if (op.b_off >= 0)          // bias K group: both K halves alias the same 8 columns (LBO 0), ones half 1 is zero
                        umma_bf16(slot_col, umma_desc(ones_addr, 2048u, 128u), umma_desc(wreg + (uint32_t)op.b_off, 0u, 128u), idesc, 1u);
                } else {
                    const uint32_t idesc = umma_idesc(128, 128);
                    for (int mt = 0; mt < (op.N >> 7); ++mt) {
                        const uint64_t w_m = w_d + (uint64_t)(mt * (2048 >> 4));
                        for (int k = 0; k < ksteps; ++k)
                            umma_bf16(slot_col + (uint32_t)mt * 128u, w_m + k * w_step, act_d + k * act_step, idesc, k > 0);
                    }
                }
                umma_commit(mbar);
                }
                __syncwarp();
            }
            AMP_PROF();        // issued
            __syncwarp();
            mbar_spin(mbar, phase);                // every warp polls (a single polling warp + bar.sync measured 2x slower)
            phase ^= 1u;
            tc_fence_after();
            AMP_PROF();        // MMA complete

            // ---- epilogue: the two warpgroups of the slot take alternating 32-column chunks ----
            if (!op.pool) {
                const float* gb = op.bias_grouped ? p.gbias + ((long long)cloud * p.n_groups + group) * op.N : nullptr;
                const bool plain = !op.store_logits || p.n_classes <= 16;
                // fp32 copy of a layer output (the local features): a thread holds 32 consecutive channels of ONE row, so direct
                // stores touch 32 rows (lines) per instruction. When the upper half of the activation buffer is free
                // (K <= 64 and at most 64 output channels kept) the chunk is staged there (16 KB, XOR-swizzled 16-byte pieces)
                // and leaves as whole 128-byte row segments, 4 rows per warp instruction.
                const bool stage_f32 = op.store_f32 && op.K <= 64 && (!op.write_act || op.N <= 64);
                for (int it = 0; it * 64 < op.N; ++it) {
                    const int c0 = it * 64 + sub * 32;
                    uint32_t v[32];
                    if (c0 < op.N) {
                    tmem_ld32(slot_col + lane_addr + (uint32_t)c0, v);
                    tmem_wait_ld();
                    const int nc = min(32, op.N - c0);           // 16 or 32
                    uint32_t u[16];
                    if (plain) {                                  // hot path: pack, ReLU on the packed pairs
                        if (gb) {                                 // per-block bias of the head's first layer
                            const float4* g4 = reinterpret_cast<const float4*>(gb + c0);
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                if (q * 4 < nc) {
                                    const float4 b = __ldg(g4 + q);
                                    v[4 * q] = __float_as_uint(__uint_as_float(v[4 * q]) + b.x);
                                    v[4 * q + 1] = __float_as_uint(__uint_as_float(v[4 * q + 1]) + b.y);
                                    v[4 * q + 2] = __float_as_uint(__uint_as_float(v[4 * q + 2]) + b.z);
                                    v[4 * q + 3] = __float_as_uint(__uint_as_float(v[4 * q + 3]) + b.w);
                                }
                            }
                        }
                        if (op.store_logits && c0 == 0 && row_ok) {   // [B, C, rows] logits: consecutive lanes = consecutive rows
                            float* lp = p.logits + (long long)cloud * p.n_classes * rows + row0 + row;
#pragma unroll
                            for (int n = 0; n < 16; ++n)
                                if (n < p.n_classes) lp[n * rows] = __uint_as_float(v[n]);
                        }
                        if (op.store_f32 && row_ok && !stage_f32) {   // fp32 copy of the layer output, before bf16 rounding
                            const float fl = op.relu ? 0.f : -INFINITY;
                            float4* o = reinterpret_cast<float4*>(p.out_f32 + grow * p.out_ld + p.out_col0 + c0);
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                if (q * 4 < nc)
                                    o[q] = make_float4(fmaxf(__uint_as_float(v[4 * q]), fl), fmaxf(__uint_as_float(v[4 * q + 1]), fl),
                                                       fmaxf(__uint_as_float(v[4 * q + 2]), fl), fmaxf(__uint_as_float(v[4 * q + 3]), fl));
                        }
#pragma unroll
                        for (int q = 0; q < 16; ++q) u[q] = pack_bf16x2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1]));
                        if (op.relu) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) u[q] = relu_bf16x2(u[q]);
                        }
                    } else {                                      // copies: only these live on the stack, v / u stay in registers
                        uint32_t vv[32], uu[16];
#pragma unroll
                        for (int j = 0; j < 32; ++j) vv[j] = v[j];
                        epilogue_slow(p, op, vv, uu, gb, c0, nc, row_ok, grow, cloud, row0 + row);
#pragma unroll
                        for (int q = 0; q < 16; ++q) u[q] = uu[q];
                    }
                    if (op.write_act) {
                        uint4* dst = reinterpret_cast<uint4*>(s_act);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (q * 8 < nc) dst[((c0 >> 3) + q) * 128 + row] = make_uint4(u[4 * q], u[4 * q + 1], u[4 * q + 2], u[4 * q + 3]);
                    }
                    }
                    if (stage_f32) {
                        float4* stg = reinterpret_cast<float4*>(s_act + kActBytes / 2);
                        const float fl = op.relu ? 0.f : -INFINITY;
#pragma unroll 1
                        for (int round = 0; round < 2; ++round) {         // the two warpgroups' chunks take turns in the buffer
                            const int cr = it * 64 + round * 32;
                            if (sub == round && cr < op.N) {
#pragma unroll
                                for (int q = 0; q < 8; ++q)
                                    stg[row * 8 + (q ^ (row & 7))] =
                                        make_float4(fmaxf(__uint_as_float(v[4 * q]), fl), fmaxf(__uint_as_float(v[4 * q + 1]), fl),
                                                    fmaxf(__uint_as_float(v[4 * q + 2]), fl), fmaxf(__uint_as_float(v[4 * q + 3]), fl));
                            }
                            slot_bar_sync(slot);
                            if (cr < op.N) {
                                const int piece = stid & 7, ncr = min(32, op.N - cr);
                                float* ob = p.out_f32 + ((long long)cloud * rows + row0) * p.out_ld + p.out_col0 + cr + piece * 4;
#pragma unroll
                                for (int ps = 0; ps < 4; ++ps) {
                                    const int r = ps * 32 + (stid >> 3);
                                    if (r < valid && piece * 4 < ncr)
                                        *reinterpret_cast<float4*>(ob + (long long)r * p.out_ld) = stg[r * 8 + (piece ^ (r & 7))];
                                }
                            }
                            slot_bar_sync(slot);
                        }
                    }
                }
            } else {
                // lanes = channels; this warpgroup's half of the tile's points: columns [sub * 64, sub * 64 + 64)
                for (int mt = 0; mt < (op.N >> 7); ++mt) {
                    float m = -3.0e38f;
                    for (int c0 = sub * 64; c0 < sub * 64 + 64; c0 += 32) {
                        if (c0 >= valid) break;
                        uint32_t v[32];
                        tmem_ld32(slot_col + lane_addr + (uint32_t)(mt * 128 + c0), v);
                        tmem_wait_ld();
                        if (c0 + 32 <= valid) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c0 + j < valid) m = fmaxf(m, __uint_as_float(v[j]));
                        }
                    }
                    if (sub * 64 < valid) {
                        const int ch = mt * 128 + row;
                        const float r = fmaxf(m + (p.pool_bias ? __ldg(p.pool_bias + ch) : 0.f), 0.f);
                        atomicMax(p.pool + (long long)cloud * op.N + ch, __float_as_uint(r));
                    }
                }
            }
            // accumulator reads and A-operand writes of this layer are done before the next MMA is issued
            AMP_PROF();        // epilogue done
            tc_fence_before();
            fence_proxy_async();
            AMP_PROF();        // fences done
            slot_bar_sync(slot);
            AMP_PROF();        // barrier done
        }
    }
    if (prof) p.prof[255] = pi;
    if (tid == 0 && !w_ready) mbar_wait(wbar, 0);        // never leave with a bulk copy in flight
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

__global__ void tc_pack_kernel(const TcPackTable t, unsigned char* __restrict__ dst) {
    pdl_sync();
    const TcPackJob& j = t.job[blockIdx.y];
    const int cloud = blockIdx.z;
    if (cloud > 0 && j.src_cloud_stride == 0) return;
    const float* src = j.src + (long long)cloud * j.src_cloud_stride;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst + j.dst_off + (long long)cloud * j.dst_cloud_stride);
    __nv_bfloat16* bout = j.bias_dst_off >= 0 ? reinterpret_cast<__nv_bfloat16*>(dst + j.bias_dst_off + (long long)cloud * j.dst_cloud_stride) : nullptr;
    const bool has_bias = j.bias || j.bias_shift;
    const int n_w = j.Npad * j.Kpad, total = n_w + (bout ? j.Npad * 8 : 0);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int ee = e < n_w ? e : e - n_w;
        const int k8 = ee & 7, n = (ee >> 3) % j.Npad, kc = (ee >> 3) / j.Npad;
        const int k = kc * 8 + k8;
        float b = 0.f, bhi = 0.f;
        if (has_bias && n < j.N) {
            b = (j.bias_scale ? j.bias_scale[n] : 1.f) * (j.bias ? j.bias[n] : 0.f) + (j.bias_shift ? j.bias_shift[n] : 0.f);
            bhi = __bfloat162float(__float2bfloat16_rn(b));
        }
        if (e >= n_w) {                                   // separate bias K group: hi at k8 = 0, lo at k8 = 1
            bout[ee] = __float2bfloat16_rn(k8 == 0 ? bhi : (k8 == 1 ? b - bhi : 0.f));
            continue;
        }
        float v = 0.f;
        int ks = k < j.K ? k : -1;
        if (j.split_in_k) {
            const int half = j.Kpad >> 1;
            ks = k < j.split_in_k ? k : ((k >= half && k < half + j.split_in_k) ? k - half : -1);
        }
        if (n < j.N && ks >= 0) {
            v = j.transposed ? src[(long long)ks * j.ld + n] : src[(long long)n * j.ld + ks];
            if (j.scale) v *= j.scale[n];
        }
        if (j.bias_col >= 0 && has_bias) {                // bias folded into two spare weight columns (input stage ones)
            if (k == j.bias_col) v = bhi;
            else if (k == j.bias_col + 1) v = b - bhi;
        }
        out[e] = __float2bfloat16_rn(v);
    }
}

}  // namespace

int tc_chain_launch(const TcChainParams& p, cudaStream_t st) {
    if (p.n_ops < 1 || p.n_ops > kTcMaxOps) return fail(AMP_E_BADARG, "tc_chain: bad op count %d", p.n_ops);
    for (int l = 0; l < p.n_ops; ++l) {
        const TcOp& o = p.op[l];
        if (o.K < 16 || o.K > 128 || o.K % 16 || o.N < 16 || o.N > 256 || o.N % 16)
            return fail(AMP_E_BADARG, "tc_chain: op %d has unsupported shape K=%d N=%d", l, o.K, o.N);
        if (o.pool && (o.N % 128 || !o.relu || !p.pool)) return fail(AMP_E_BADARG, "tc_chain: op %d cannot pool", l);
        if (o.write_act && (o.N > 128 || l + 1 >= p.n_ops || p.op[l + 1].K != o.N))
            return fail(AMP_E_BADARG, "tc_chain: op %d does not feed op %d", l, l + 1);
        if (o.w_off % 128 || (o.b_off >= 0 && o.b_off % 128)) return fail(AMP_E_BADARG, "tc_chain: op %d weights are not 128-byte aligned", l);
        if (o.pool && o.b_off >= 0) return fail(AMP_E_BADARG, "tc_chain: op %d: pooled ops take pool_bias", l);
        if (o.store_logits && (!p.logits || p.n_classes < 1 || p.n_classes > o.N || p.n_classes > 32))
            return fail(AMP_E_BADARG, "tc_chain: op %d cannot store logits", l);
        if (o.store_f32 && (!p.out_f32 || p.out_ld % 4 || p.out_col0 % 4 || ((uintptr_t)p.out_f32 & 15)))
            return fail(AMP_E_BADARG, "tc_chain: op %d output rows are not 16-byte aligned", l);
        if (o.bias_grouped && (!p.gbias || ((uintptr_t)p.gbias & 15) || (p.n_groups > 1 && !p.group_rows)))
            return fail(AMP_E_BADARG, "tc_chain: op %d grouped bias missing", l);
    }
    if (p.in_mode == 0) {
        if (!((p.op[0].K == 16 && p.in_k >= 1 && p.in_k <= 8) || (p.op[0].K == 32 && p.in_k == 9)))
            return fail(AMP_E_BADARG, "tc_chain: split input needs K=16 (<= 8 columns) or K=32 (9 columns)");
    } else if (p.in_ld % 4 || ((uintptr_t)p.in_x & 15)) {
        return fail(AMP_E_BADARG, "tc_chain: input rows are not 16-byte aligned");
    }
    if (p.wblob_bytes % 16 || p.wcloud_bytes % 16 || p.wcloud_stride % 16 || ((uintptr_t)p.wblob & 15) || ((uintptr_t)p.wcloud & 15))
        return fail(AMP_E_BADARG, "tc_chain: packed weights are not 16-byte aligned");
    if (p.n_clouds < 1 || p.rows_per_cloud < 1) return fail(AMP_E_BADARG, "tc_chain: empty input");
    const SmemPlan sp = smem_plan(p.wblob_bytes, p.wcloud_bytes);
    if (sp.total > kMaxSmem) return fail(AMP_E_BADARG, "tc_chain: chain needs %d bytes of shared memory", sp.total);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e != cudaSuccess) return fail(AMP_E_CUDA, "tc_chain: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const long long n_tiles = (long long)p.n_clouds * ((p.rows_per_cloud + kTcTileRows - 1) / kTcTileRows);
    const int grid = (int)((n_tiles + 1) / 2 < kNumSMs ? (n_tiles + 1) / 2 : kNumSMs);
    const int smem_bytes = sp.total < kMinSmem ? kMinSmem : sp.total;
    static const bool want_prof = getenv("AMP_CHAIN_PROF") != nullptr;
    if (want_prof) {                                     // debugging aid: synchronous, prints the phase timeline of CTA 0 / slot 0
        static long long* dprof = nullptr;
        if (!dprof) cudaMalloc(&dprof, 256 * sizeof(long long));
        cudaMemsetAsync(dprof, 0, 256 * sizeof(long long), st);
        TcChainParams q = p;
        q.prof = dprof;
        launch_pdl(tc_chain_kernel, dim3((unsigned)(grid)), dim3(kThreads), smem_bytes, st, q);
        long long h[256];
        cudaMemcpyAsync(h, dprof, sizeof h, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        fprintf(stderr, "[tc_chain prof] ops=%d tiles=%lld:", p.n_ops, n_tiles);
        for (int i = 1; i < (int)h[255] && i < 250; ++i) fprintf(stderr, " %lld", h[i] - h[i - 1]);
        fprintf(stderr, "\n");
        count_launch();
        return check_launch("tc_chain_kernel");
    }
    launch_pdl(tc_chain_kernel, dim3((unsigned)(grid)), dim3(kThreads), smem_bytes, st, p);
    count_launch();
    count_path("tc_chain");
    return check_launch("tc_chain_kernel");
}

int tc_pack_weights(const TcPackTable& t, unsigned char* dst, cudaStream_t st) {
    if (t.n < 1 || t.n > TcPackTable::kMax) return fail(AMP_E_BADARG, "tc_pack_weights: bad job count");
    int mx = 0;
    for (int i = 0; i < t.n; ++i) {
        const int e = t.job[i].Npad * (t.job[i].Kpad + 8);
        if (e > mx) mx = e;
        if (t.job[i].Npad % 16 || t.job[i].Kpad % 16 || t.job[i].dst_off % 128)
            return fail(AMP_E_BADARG, "tc_pack_weights: job %d is not tile aligned", i);
    }
    int bx = (mx + 255) / 256;
    if (bx > 32) bx = 32;
    launch_pdl(tc_pack_kernel, dim3(bx, t.n, t.n_clouds < 1 ? 1 : t.n_clouds), dim3(256), 0, st, t, dst);
    count_launch();
    return check_launch("tc_pack_kernel");
}

}  // namespace amp

// ---------------------------------------------------------------------------------------------------------------
// C ABI: one point-wise linear layer on the tensor cores (unit of the chains above, also used by the tests)
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

size_t amp_tc_linear_workspace_bytes(int32_t K, int32_t N) {
    const int Kp = (K + 15) / 16 * 16, Np = (N + 15) / 16 * 16;
    return (size_t)amp::tc_packed_bytes(Np, Kp) + (size_t)amp::tc_bias_bytes(Np) + 512;
}

int amp_tc_linear_bf16(const float* x, int64_t n_clouds, int64_t rows_per_cloud, int32_t K, const float* w, const float* bias,
                       int32_t N, int32_t relu, float* y, uint32_t* pool_max, void* workspace, size_t workspace_bytes,
                       void* stream) {
    using namespace amp;
    if (!x || !w || !workspace || (!y && !pool_max)) return fail(AMP_E_BADARG, "tc_linear: null pointer");
    if (K < 16 || K > 128 || K % 16 || N < 16 || N > 256 || N % 16)
        return fail(AMP_E_BADARG, "tc_linear: K must be a multiple of 16 in [16, 128], N a multiple of 16 in [16, 256]");
    if (n_clouds < 1 || rows_per_cloud < 1 || n_clouds * rows_per_cloud > (1LL << 31) / 256)
        return fail(AMP_E_BADARG, "tc_linear: unsupported row count");
    if (workspace_bytes < amp_tc_linear_workspace_bytes(K, N)) return fail(AMP_E_WORKSPACE, "tc_linear: workspace too small");
    if (pool_max && (!relu || N % 128)) return fail(AMP_E_BADARG, "tc_linear: pooling needs relu and N % 128 == 0");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    TcPackTable pt{};
    pt.n = 1; pt.n_clouds = 1;
    const int wbytes = tc_packed_bytes(N, K);
    const bool bias_group = bias && !pool_max;            // pooled layers add their bias in the epilogue
    pt.job[0] = TcPackJob{w, K, 0, nullptr, bias, nullptr, nullptr, -1, bias_group ? wbytes : -1, N, K, N, K, 0, 0, 0, 0};
    int rc = tc_pack_weights(pt, base, st);
    if (rc != AMP_OK) return rc;
    TcChainParams p{};
    p.n_ops = 1;
    p.op[0] = TcOp{K, N, 0, 0, bias_group ? wbytes : -1, relu, 0, 0, pool_max ? 0 : 1, pool_max ? 1 : 0, 0};
    p.in_mode = 1; p.in_x = x; p.in_ld = K; p.in_k = K;
    p.wblob = base; p.wblob_bytes = wbytes + (bias_group ? tc_bias_bytes(N) : 0);
    p.pool_bias = pool_max ? bias : nullptr;
    p.n_groups = 1;
    p.out_f32 = y; p.out_ld = N; p.out_col0 = 0;
    p.pool = pool_max;
    p.n_clouds = (int)n_clouds; p.rows_per_cloud = (int)rows_per_cloud;
    return tc_chain_launch(p, st);
}

}  // extern "C"
