// Fused shared-MLP chain kernel at fp32-class accuracy (see tc_chain32.cuh for the contract).
//
// Layout of one CTA (288 threads, 1 CTA per SM):
//   * warps 0-3 / 4-7 = two "slots" of 128 threads. A slot walks its own stream of 128-point tiles; per tile and layer
//     one elected lane of the slot's first warp issues the tcgen05.mma K steps (three per step: A_hi W_hi, A_lo W_hi,
//     A_hi W_lo) and commits to the slot's mbarrier; the slot's 4 warps wait and run the epilogue: thread = TMEM lane =
//     point row, 64 accumulator columns per pass (two tcgen05.ld in flight before one wait: with 8 warps per slot and
//     16-column pieces the epilogue was bound by TMEM round-trip latency, 4.2 k cycles for a 128-channel layer), bias, ReLU,
//     split into bf16 hi + lo pairs, written back into TENSOR MEMORY as the next layer's A operand (tcgen05.st). While
//     one slot is in its epilogue the other slot's MMAs run.
//   * TMEM columns of a slot (256): [0, 128) fp32 accumulator, [128, 192) A_hi, [192, 256) A_lo (K <= 128 as packed pairs).
//   * warp 8 = producer: stages the resident weights (TMA bulk copies, one mbarrier per layer so that the first layers
//     start while the big last one is still arriving) and, for the streamed pooled layer, keeps a two-stage ring of 32 KB
//     weight chunks full. Both slots consume the same chunk sequence (one pass per pair of tiles); a stage is released
//     by one tcgen05.commit per slot (mbarrier count 2).
//   * per-cloud weights (folded conv_1, feature transform) are double buffered per slot and fetched one tile ahead by TMA
//     bulk copies; the 3 / 9 input columns of the next tile are fetched one tile ahead into registers.
// Weight layout in shared memory (B operand): K-major, no swizzle, 8 x 16-byte core matrices, hi block then lo block:
//   element (n, k) at ((k / 8) * N + n) * 16 + (k % 8) * 2     -> SBO = 128 B, LBO = N * 16 B.
#include <stdlib.h>

#include "tc_chain32.cuh"
#include "tc_ts.cuh"

namespace amp {
namespace {
using namespace tcx;

constexpr int kThreads = 288, kSlotThreads = 128, kRows = 128;
constexpr int kAcc = 0, kAhi = 128, kAlo = 192, kSlotCols = 256;
constexpr int kStageBytes = kRows * 64 * 4;               // fp32 staging of a 64-channel output tile (store_f32)
constexpr int kMaxSmem = 232448;                          // 227 KB per CTA on sm_100
constexpr int kMinSmem = 120 * 1024;                      // more than half an SM: one CTA (one TMEM owner) per SM

__host__ __device__ inline int align_i(int v, int a) { return (v + a - 1) / a * a; }

struct Plan { int w, wc, bias, stage, ring, bar, gb, total; };
__host__ __device__ inline Plan plan_of(int wblob_bytes, int wcloud_bytes, int n_bias, int stage_f32, int stream) {
    Plan s;
    s.w = 0;
    s.wc = align_i(wblob_bytes, 128);
    s.bias = s.wc + 4 * align_i(wcloud_bytes, 128);          // two slots x two buffers
    s.stage = s.bias + align_i(n_bias * 4, 128);
    s.ring = s.stage + (stage_f32 ? 2 * kStageBytes : 0);
    s.bar = s.ring + (stream ? 2 * kT32ChunkBytes : 0);
    s.gb = s.bar + 256;                                       // per slot: the tile's grouped-bias row (128 floats)
    s.total = s.gb + 2 * 512;
    return s;
}

__device__ __forceinline__ void slot_bar_sync(int slot) { asm volatile("bar.sync %0, 128;" ::"r"(slot + 1) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// the three products of every 16-wide K step of one [128 x ncols] output block; A from TMEM, W = hi / lo descriptors
__device__ __forceinline__ void issue_block(uint32_t acc, uint32_t ahi, uint32_t alo, uint64_t whi, uint64_t wlo, uint64_t step,
                                            int ksteps, uint32_t idesc) {
    for (int k = 0; k < ksteps; ++k) {
        umma_bf16_ts(acc, ahi + (uint32_t)k * 8u, whi + k * step, idesc, k > 0 ? 1u : 0u);
        umma_bf16_ts(acc, alo + (uint32_t)k * 8u, whi + k * step, idesc, 1u);
        umma_bf16_ts(acc, ahi + (uint32_t)k * 8u, wlo + k * step, idesc, 1u);
    }
}

__global__ void __launch_bounds__(kThreads, 1) tc_chain32_kernel(const __grid_constant__ T32Params p) {
    pdl_trigger();
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = warp_index_uniform(), lane = tid & 31;
    int pi = 0;
    const bool prof = p.prof != nullptr && blockIdx.x == 0 && tid == 0;
#define T32_PROF() do { if (prof && pi < 250) p.prof[pi++] = clock64(); } while (0)
    T32_PROF();                                                  // 0: kernel entry
    int stage_f32 = 0, stream = 0;
    for (int l = 0; l < p.n_ops; ++l) { stage_f32 |= p.op[l].store_f32; stream |= p.op[l].w_stream; }
    const Plan sp = plan_of(p.wblob_bytes, p.wcloud_bytes, p.n_bias, stage_f32, stream);
    float* s_bias = reinterpret_cast<float*>(smem + sp.bias);
    // [1 + slot] MMA done, [3 + s] ring full, [5 + s] ring free, [8 + op] resident weights of op (so that the first layers
    // start while the large last layer of the chain is still streaming in)
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + sp.bar);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + sp.bar + 192);
    const uint32_t wbar0 = smem_u32(&s_bar[8]);
    const uint32_t full0 = smem_u32(&s_bar[3]), free0 = smem_u32(&s_bar[5]);

    if (tid == 0) {
        for (int l = 0; l < kT32MaxOps; ++l) mbar_init(wbar0 + 8 * l, 1);
        for (int i = 16; i < 20; ++i) mbar_init(smem_u32(&s_bar[i]), 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        mbar_init(smem_u32(&s_bar[2]), 1);
        mbar_init(full0, 1); mbar_init(full0 + 8, 1);
        mbar_init(free0, 2); mbar_init(free0 + 8, 2);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(s_tmem), 512);
    pdl_wait();                                                  // first global-memory access below
    for (int i = tid; i < p.n_bias; i += kThreads) s_bias[i] = __ldg(p.bias + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = uniform_u32(*s_tmem);
    T32_PROF();                                                  // 1: prologue done (barriers, TMEM, bias table)

    const int rows = p.rows_per_cloud;
    const int tiles_per_cloud = (rows + kRows - 1) / kRows;
    const int n_tiles = p.n_clouds * tiles_per_cloud;
    const int tile_stride = (int)gridDim.x * 2;

    if (warp == 8) {
        // ---- producer: resident weights once, then the streamed chunks of every round of this CTA ----
        if (elect_one_sync()) {
            for (int l = 0; l < p.n_ops; ++l) {                  // resident weights, one barrier per op, in layer order
                const T32Op& op = p.op[l];
                if (op.w_cloud || op.w_stream) continue;
                const int bytes = op.N * op.K * 4;
                mbar_expect_tx(wbar0 + 8 * l, (uint32_t)bytes);
                for (int off = 0; off < bytes; off += 32768) {
                    const int n = min(32768, bytes - off);
                    bulk_g2s(smem_u32(smem + sp.w + op.w_off + off), p.wblob + op.w_off + off, (uint32_t)n, wbar0 + 8 * l);
                }
            }
            if (stream) {
                int rounds = 0;
                for (int t = (int)blockIdx.x * 2; t < n_tiles; t += tile_stride) ++rounds;
                for (int g = 0; g < rounds * 4; ++g) {
                    const int s = g & 1, u = g >> 1;
                    if (u > 0) mbar_wait_bounded(free0 + 8 * s, (uint32_t)((u - 1) & 1));
                    mbar_expect_tx(full0 + 8 * s, (uint32_t)kT32ChunkBytes);
                    bulk_g2s(smem_u32(smem + sp.ring + s * kT32ChunkBytes), p.wstream + (size_t)(g & 3) * kT32ChunkBytes,
                             (uint32_t)kT32ChunkBytes, full0 + 8 * s);
                }
            }
        }
        __syncwarp();
    } else {
        const int slot = warp >> 2, stid = tid & (kSlotThreads - 1);
        const int row = stid;                                        // this thread's TMEM lane == tile row (warp % 4 == lane quadrant)
        const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t tcol = tmem_base + (uint32_t)(slot * kSlotCols);
        const uint32_t mbar = smem_u32(&s_bar[1 + slot]);
        const int wc_bytes = align_i(p.wcloud_bytes, 128);
        // per-cloud weights: two buffers per slot, filled by TMA bulk copies one tile ahead; barriers [16 + 2 * slot + buffer]
        const uint32_t wc_addr0 = smem_u32(smem + sp.wc + slot * 2 * wc_bytes);
        const uint32_t cbar0 = smem_u32(&s_bar[16 + 2 * slot]);
        float4* s_stage = reinterpret_cast<float4*>(smem + sp.stage + slot * kStageBytes);
        float* s_gb = reinterpret_cast<float*>(smem + sp.gb + slot * 512);
        const uint32_t w_addr = smem_u32(smem + sp.w), ring_addr = smem_u32(smem + sp.ring);
        const bool issuer_warp = (warp & 3) == 0;
        uint32_t phase = 0;
        uint32_t w_ready = 0;                                        // bit l: resident weights of op l have landed
        for (int l = 0; l < p.n_ops; ++l)
            if (p.op[l].w_cloud || p.op[l].w_stream) w_ready |= 1u << l;
        auto tile_of = [&](int r) { return (int)blockIdx.x * 2 + slot + r * tile_stride; };
        auto request_cloud = [&](int r) {                            // one elected lane of the slot's first warp
            const int t = tile_of(r);
            if (p.wcloud_bytes > 0 && t < n_tiles) {
                const int cl = t / tiles_per_cloud, b = r & 1;
                mbar_expect_tx(cbar0 + 8 * b, (uint32_t)p.wcloud_bytes);
                bulk_g2s(wc_addr0 + (uint32_t)(b * wc_bytes), p.wcloud + (long long)cl * p.wcloud_stride, (uint32_t)p.wcloud_bytes, cbar0 + 8 * b);
            }
        };
        // narrow input rows (in_mode 0) are fetched one tile ahead into registers
        float xn[10];
        auto fetch_rows = [&](int r) {
            const int t = tile_of(r);
#pragma unroll
            for (int j = 0; j < 10; ++j) xn[j] = 0.f;
            if (p.in_mode == 0 && t < n_tiles) {
                const int cl = t / tiles_per_cloud, r0 = (t - cl * tiles_per_cloud) * kRows, vd = min(kRows, rows - r0);
                const float* src = p.in_x + ((long long)cl * rows + r0 + (row < vd ? row : vd - 1)) * p.in_ld;
#pragma unroll
                for (int j = 0; j < 10; ++j)
                    if (j < p.in_k) xn[j] = __ldg(src + j);
            }
        };
        if (issuer_warp) {
            if (elect_one_sync()) request_cloud(0);
            __syncwarp();
        }
        fetch_rows(0);

        for (int r = 0;; ++r) {
            const int tile = tile_of(r);
            if (tile >= n_tiles) {
                // the other slot still has a tile in this round: release the streamed chunks it shares with this slot
                if (stream && slot == 1 && tile - 1 < n_tiles && issuer_warp) {
                    if (elect_one_sync()) {
                        for (int c = 0; c < 4; ++c) {
                            const int g = r * 4 + c, s = g & 1;
                            mbar_wait_bounded(full0 + 8 * s, (uint32_t)((g >> 1) & 1));
                            mbar_arrive(free0 + 8 * s);
                        }
                    }
                    __syncwarp();
                }
                break;
            }
            const int cloud = tile / tiles_per_cloud;
            const int row0 = (tile - cloud * tiles_per_cloud) * kRows;
            const int valid = min(kRows, rows - row0);
            const bool row_ok = row < valid;
            // rows past the end of the cloud repeat its last row: they change no maximum and are never stored
            const uint32_t wc_addr = wc_addr0 + (uint32_t)((r & 1) * wc_bytes);

            // ---- input stage: fp32 row -> bf16 hi / lo pairs straight into tensor memory ----
            if (p.in_mode == 0) {
                uint32_t hi[8], lo[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (q < 5) split_pair(xn[2 * q], xn[2 * q + 1], hi[q], lo[q]);
                    else { hi[q] = 0u; lo[q] = 0u; }
                }
                tmem_st8(tcol + lane_addr + kAhi, hi);
                tmem_st8(tcol + lane_addr + kAlo, lo);
            } else {
                const long long srow = (long long)cloud * rows + row0 + (row_ok ? row : valid - 1);
                const float4* src = reinterpret_cast<const float4*>(p.in_x + srow * p.in_ld);
                float4 a[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) a[q] = __ldg(src + q);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        split_pair(a[8 * h + q].x, a[8 * h + q].y, hi[2 * q], lo[2 * q]);
                        split_pair(a[8 * h + q].z, a[8 * h + q].w, hi[2 * q + 1], lo[2 * q + 1]);
                    }
                    tmem_st16(tcol + lane_addr + kAhi + h * 16, hi);
                    tmem_st16(tcol + lane_addr + kAlo + h * 16, lo);
                }
            }
            // the next tile's per-cloud weights (other buffer: its readers, the MMAs of the previous tile, are done) and rows
            if (issuer_warp) {
                if (elect_one_sync()) request_cloud(r + 1);
                __syncwarp();
            }
            fetch_rows(r + 1);
            // group of this row (per-block bias of the segmentation head); when the whole tile lies in one block (the usual
            // case: blocks of 2048 points, tiles of 128) its bias row goes to shared memory once and the first layer's epilogue
            // reads it like an ordinary bias table instead of 32 global loads per thread
            int group = 0;
            bool gb_uniform = false;
            if (p.gbias) {
                int g_first = 0, g_last = 0;
                if (p.n_groups > 1) {
                    const int rr = row0 + (row_ok ? row : valid - 1);
                    for (int g = 1; g < p.n_groups; ++g) {
                        const int start = __ldg(p.group_rows + g);
                        group += (rr >= start) ? 1 : 0;
                        g_first += (row0 >= start) ? 1 : 0;
                        g_last += (row0 + valid - 1 >= start) ? 1 : 0;
                    }
                }
                gb_uniform = g_first == g_last;                          // computed alike by every thread of the slot
                if (gb_uniform && stid < p.op[0].N) s_gb[stid] = __ldg(p.gbias + ((long long)cloud * p.n_groups + g_first) * p.op[0].N + stid);
            }
            tmem_wait_st();
            tc_fence_before();
            T32_PROF();                                          // input staged
            slot_bar_sync(slot);
            T32_PROF();                                          // slot barrier

            for (int l = 0; l < p.n_ops; ++l) {
                // a register copy of the op: read through the reference, every `if (op.x)` of the unrolled epilogue is its own
                // constant-bank load -> compare -> branch chain (~70 cycles each, 20 of them per 64 columns: measured 1.4 k cycles)
                const T32Op op = p.op[l];
                const int parts = op.pool ? (op.N >> 7) : 1;
                for (int part = 0; part < parts; ++part) {
                    // ---- MMA issue: the first warp of the slot enters, one elected lane issues (uniform descriptors) ----
                    if (issuer_warp) {
                        if (!(w_ready & (1u << l))) { mbar_wait_bounded(wbar0 + 8 * l, 0); w_ready |= 1u << l; }
                        if (op.w_cloud && part == 0) mbar_wait_bounded(cbar0 + 8 * (r & 1), (uint32_t)((r >> 1) & 1));
                        tc_fence_after();
                        if (elect_one_sync()) {
                            const int ksteps = op.K >> 4;
                            const uint32_t acc = tcol + kAcc, ahi = tcol + kAhi, alo = tcol + kAlo;
                            if (!op.w_stream) {
                                const int ncols = op.pool ? 128 : op.N;
                                const uint32_t wb = (op.w_cloud ? wc_addr : w_addr) + (uint32_t)op.w_off + (uint32_t)(part * 128 * 16);
                                const uint32_t lbo = (uint32_t)op.N * 16u;
                                issue_block(acc, ahi, alo, umma_desc(wb, lbo, 128u), umma_desc(wb + (uint32_t)(op.N * op.K * 2), lbo, 128u),
                                            (uint64_t)((2u * lbo) >> 4), ksteps, umma_idesc(128, ncols));
                            } else {
                                const uint32_t lbo = (uint32_t)kT32ChunkChannels * 16u;
                                const uint32_t idesc = umma_idesc(128, kT32ChunkChannels);
                                for (int c = 0; c < 2; ++c) {
                                    const int g = r * 4 + part * 2 + c, s = g & 1;
                                    mbar_wait_bounded(full0 + 8 * s, (uint32_t)((g >> 1) & 1));
                                    const uint32_t wb = ring_addr + (uint32_t)(s * kT32ChunkBytes);
                                    issue_block(acc + (uint32_t)(c * kT32ChunkChannels), ahi, alo, umma_desc(wb, lbo, 128u),
                                                umma_desc(wb + (uint32_t)(kT32ChunkBytes / 2), lbo, 128u), (uint64_t)((2u * lbo) >> 4), ksteps, idesc);
                                    umma_commit(free0 + 8 * s);       // stage reusable once these MMAs (of both slots) have read it
                                }
                            }
                            umma_commit(mbar);
                        }
                        __syncwarp();
                    }
                    T32_PROF();                                  // issued
                    mbar_wait_bounded(mbar, phase);
                    phase ^= 1u;
                    tc_fence_after();
                    T32_PROF();                                  // MMAs complete

                    if (!op.pool) {
                        const float* gb = op.bias_grouped ? p.gbias + ((long long)cloud * p.n_groups + group) * op.N : nullptr;
                        // a thread owns a whole row: 64 accumulator columns per pass, both TMEM loads in flight before the wait.
                        // The common layer (bias + ReLU + hand-over, N a multiple of 64) gets a branch-free pass: with one
                        // epilogue warp per scheduler every uniform branch of the generic pass below (5 per 16 columns) is
                        // exposed latency (measured: 1.4 k cycles for the bias / ReLU section of 64 columns)
                        const bool gb_plain = op.bias_grouped && gb_uniform && l == 0 && op.bias_off < 0;
                        const bool plain = (gb_plain || (op.bias_off >= 0 && !op.bias_grouped)) && op.relu && op.write_act &&
                                           !op.store_logits && !op.store_f32 && (op.N & 63) == 0;
                        if (plain) {
                            for (int c0 = 0; c0 < op.N; c0 += 64) {
                                uint32_t v[64];
                                tmem_ld32(tcol + lane_addr + kAcc + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                                tmem_ld32(tcol + lane_addr + kAcc + (uint32_t)(c0 + 32), *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                                tmem_wait_ld();
                                T32_PROF();                      // (epilogue detail) accumulator columns in registers
                                const float4* b4 = reinterpret_cast<const float4*>((gb_plain ? s_gb : s_bias + op.bias_off) + c0);
#pragma unroll
                                for (int q = 0; q < 16; ++q) {
                                    const float4 bb = b4[q];
                                    v[4 * q] = __float_as_uint(fmaxf(__uint_as_float(v[4 * q]) + bb.x, 0.f));
                                    v[4 * q + 1] = __float_as_uint(fmaxf(__uint_as_float(v[4 * q + 1]) + bb.y, 0.f));
                                    v[4 * q + 2] = __float_as_uint(fmaxf(__uint_as_float(v[4 * q + 2]) + bb.z, 0.f));
                                    v[4 * q + 3] = __float_as_uint(fmaxf(__uint_as_float(v[4 * q + 3]) + bb.w, 0.f));
                                }
                                T32_PROF();                      // (epilogue detail) bias / ReLU done
#pragma unroll
                                for (int h = 0; h < 2; ++h) {
                                    uint32_t hi[16], lo[16];
#pragma unroll
                                    for (int q = 0; q < 16; ++q)
                                        split_pair(__uint_as_float(v[32 * h + 2 * q]), __uint_as_float(v[32 * h + 2 * q + 1]), hi[q], lo[q]);
                                    tmem_st16(tcol + lane_addr + kAhi + (uint32_t)((c0 >> 1) + 16 * h), hi);
                                    tmem_st16(tcol + lane_addr + kAlo + (uint32_t)((c0 >> 1) + 16 * h), lo);
                                }
                            }
                            T32_PROF();                          // (epilogue detail) split + tcgen05.st issued
                            tmem_wait_st();
                        } else {
                        for (int c0 = 0; c0 < op.N; c0 += 64) {
                            const int nc = min(64, op.N - c0);           // 16, 32 or 64
                            uint32_t v[64];
                            tmem_ld32(tcol + lane_addr + kAcc + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                            if (nc > 32) tmem_ld32(tcol + lane_addr + kAcc + (uint32_t)(c0 + 32), *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                            tmem_wait_ld();
                            T32_PROF();                          // (epilogue detail) accumulator columns in registers
#pragma unroll
                            for (int h = 0; h < 4; ++h) {                // 16 columns at a time
                                if (h * 16 < nc) {
                                    const int cc = c0 + 16 * h;
                                    uint32_t* vv = &v[16 * h];
                                    if (op.bias_off >= 0) {
                                        const float4* b4 = reinterpret_cast<const float4*>(s_bias + op.bias_off + cc);
#pragma unroll
                                        for (int q = 0; q < 4; ++q) {
                                            const float4 b = b4[q];
                                            vv[4 * q] = __float_as_uint(__uint_as_float(vv[4 * q]) + b.x);
                                            vv[4 * q + 1] = __float_as_uint(__uint_as_float(vv[4 * q + 1]) + b.y);
                                            vv[4 * q + 2] = __float_as_uint(__uint_as_float(vv[4 * q + 2]) + b.z);
                                            vv[4 * q + 3] = __float_as_uint(__uint_as_float(vv[4 * q + 3]) + b.w);
                                        }
                                    }
                                    if (gb) {
                                        const float4* g4 = reinterpret_cast<const float4*>(gb + cc);
#pragma unroll
                                        for (int q = 0; q < 4; ++q) {
                                            const float4 b = __ldg(g4 + q);
                                            vv[4 * q] = __float_as_uint(__uint_as_float(vv[4 * q]) + b.x);
                                            vv[4 * q + 1] = __float_as_uint(__uint_as_float(vv[4 * q + 1]) + b.y);
                                            vv[4 * q + 2] = __float_as_uint(__uint_as_float(vv[4 * q + 2]) + b.z);
                                            vv[4 * q + 3] = __float_as_uint(__uint_as_float(vv[4 * q + 3]) + b.w);
                                        }
                                    }
                                    if (op.relu) {
#pragma unroll
                                        for (int j = 0; j < 16; ++j) vv[j] = __float_as_uint(fmaxf(__uint_as_float(vv[j]), 0.f));
                                    }
                                    if (op.store_logits && c0 == 0 && h < 2 && row_ok) {   // [B, C, rows]: consecutive lanes = consecutive rows
                                        float* lp = p.logits + (long long)cloud * p.n_classes * rows + row0 + row;
#pragma unroll
                                        for (int n = 0; n < 16; ++n)
                                            if (16 * h + n < p.n_classes) lp[(long long)(16 * h + n) * rows] = __uint_as_float(vv[n]);
                                    }
                                    if (op.store_f32) {                  // 64 channels: 16 x 16-byte pieces per row, XOR-swizzled
#pragma unroll
                                        for (int q = 0; q < 4; ++q)
                                            s_stage[row * 16 + ((4 * h + q) ^ (row & 15))] =
                                                make_float4(__uint_as_float(vv[4 * q]), __uint_as_float(vv[4 * q + 1]), __uint_as_float(vv[4 * q + 2]),
                                                            __uint_as_float(vv[4 * q + 3]));
                                    }
                                }
                            }
                            T32_PROF();                          // (epilogue detail) bias / ReLU / stores to memory done
                            if (op.write_act) {
#pragma unroll
                                for (int h = 0; h < 2; ++h) {            // 32 columns -> 16 packed hi + 16 packed lo columns
                                    if (h * 32 < nc) {
                                        uint32_t hi[16], lo[16];
#pragma unroll
                                        for (int q = 0; q < 16; ++q)
                                            split_pair(__uint_as_float(v[32 * h + 2 * q]), __uint_as_float(v[32 * h + 2 * q + 1]), hi[q], lo[q]);
                                        if (nc - h * 32 >= 32) {
                                            tmem_st16(tcol + lane_addr + kAhi + (uint32_t)((c0 >> 1) + 16 * h), hi);
                                            tmem_st16(tcol + lane_addr + kAlo + (uint32_t)((c0 >> 1) + 16 * h), lo);
                                        } else {
                                            uint32_t h8[8], l8[8];
#pragma unroll
                                            for (int q = 0; q < 8; ++q) { h8[q] = hi[q]; l8[q] = lo[q]; }
                                            tmem_st8(tcol + lane_addr + kAhi + (uint32_t)((c0 >> 1) + 16 * h), h8);
                                            tmem_st8(tcol + lane_addr + kAlo + (uint32_t)((c0 >> 1) + 16 * h), l8);
                                        }
                                    }
                                }
                            }
                        }
                        T32_PROF();                              // (epilogue detail) split + tcgen05.st issued
                        if (op.write_act) tmem_wait_st();
                        if (op.store_f32) {                              // whole 256-byte row segments leave coalesced
                            slot_bar_sync(slot);
                            float* ob = p.out_f32 + ((long long)cloud * rows + row0) * p.out_ld + p.out_col0;
                            for (int i = stid; i < kRows * 16; i += kSlotThreads) {
                                const int rr = i >> 4, piece = i & 15;
                                if (rr < valid) *reinterpret_cast<float4*>(ob + (long long)rr * p.out_ld + piece * 4) = s_stage[rr * 16 + (piece ^ (rr & 15))];
                            }
                        }
                        }
                    } else {
                        // max over the 128 rows of the tile, 32 rows per warp by redux; lane j keeps channels j, 32 + j, 64 + j, 96 + j
                        const float* b = s_bias + op.bias_off + part * 128;
                        for (int c0 = 0; c0 < 128; c0 += 64) {
                            uint32_t v[64];
                            tmem_ld32(tcol + lane_addr + kAcc + (uint32_t)c0, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                            tmem_ld32(tcol + lane_addr + kAcc + (uint32_t)(c0 + 32), *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                            tmem_wait_ld();
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                float mine = 0.f;
#pragma unroll
                                for (int q = 0; q < 8; ++q) {
                                    const float4 bb = *reinterpret_cast<const float4*>(b + c0 + 32 * h + 4 * q);
                                    const float m0 = warp_max_relu_safe(__uint_as_float(v[32 * h + 4 * q]) + bb.x);
                                    const float m1 = warp_max_relu_safe(__uint_as_float(v[32 * h + 4 * q + 1]) + bb.y);
                                    const float m2 = warp_max_relu_safe(__uint_as_float(v[32 * h + 4 * q + 2]) + bb.z);
                                    const float m3 = warp_max_relu_safe(__uint_as_float(v[32 * h + 4 * q + 3]) + bb.w);
                                    if (lane == 4 * q) mine = m0;
                                    if (lane == 4 * q + 1) mine = m1;
                                    if (lane == 4 * q + 2) mine = m2;
                                    if (lane == 4 * q + 3) mine = m3;
                                }
                                atomicMax(p.pool + (long long)cloud * op.N + part * 128 + c0 + 32 * h + lane, __float_as_uint(fmaxf(mine, 0.f)));
                            }
                        }
                    }
                    // accumulator reads and A-operand writes of this layer are done before the next MMA is issued
                    tc_fence_before();
                    T32_PROF();                                  // epilogue done
                    slot_bar_sync(slot);
                    T32_PROF();                                  // slot barrier
                }
            }
        }
        if (issuer_warp)                                           // never leave with a bulk copy in flight
            for (int l = 0; l < p.n_ops; ++l)
                if (!(w_ready & (1u << l))) mbar_wait_bounded(wbar0 + 8 * l, 0);
    }
    T32_PROF();                                                  // all tiles done
    if (prof) p.prof[255] = pi;
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

__global__ void t32_pack_kernel(const T32PackTable t, unsigned char* __restrict__ dst) {
    pdl_sync();
    const T32PackJob& j = t.job[blockIdx.y];
    const int cloud = blockIdx.z;
    if (cloud > 0 && j.src_cloud_stride == 0) return;
    const float* src = j.src + (long long)cloud * j.src_cloud_stride;
    unsigned char* out = dst + j.dst_off + (long long)cloud * j.dst_cloud_stride;
    const int total = j.Npad * j.Kpad;
    const int cn = j.chunk_n > 0 ? j.chunk_n : j.Npad;           // rows per independent block
    const int block_bytes = cn * j.Kpad * 2;                      // one hi (or lo) block
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k = e % j.Kpad, n = e / j.Kpad;
        float v = 0.f;
        if (n < j.N && k < j.K) {
            v = j.transposed ? src[(long long)k * j.ld + n] : src[(long long)n * j.ld + k];
            if (j.scale) v *= j.scale[n];
        }
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
        const int c = n / cn, nn = n - c * cn;
        unsigned char* blk = out + (long long)c * 2 * block_bytes;
        const int off = ((k >> 3) * cn + nn) * 16 + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(blk + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(blk + block_bytes + off) = l;
    }
}

// bmm + cat + conv_1 (pointnetAtt.py:85-90) folded into per-cloud conv_1 weights, BatchNorm scale applied, split and packed:
//   W1eff[b][c][i] = W1[c][3 + i] + (i < 3 ? sum_j W1[c][j] * T[b][i][j] : 0),  i < 9 (K padded to 16)
__global__ void t32_fold_w1_kernel(const float* __restrict__ W1, const float* __restrict__ T, const float* __restrict__ scale,
                                   unsigned char* __restrict__ dst, long long dst_stride) {
    pdl_sync();
    const int b = blockIdx.x;
    const float* t = T + b * 9;
    unsigned char* out = dst + (long long)b * dst_stride;
    for (int e = threadIdx.x; e < 64 * 16; e += blockDim.x) {
        const int c = e >> 4, i = e & 15;
        float v = 0.f;
        if (i < 9) {
            v = W1[c * 12 + 3 + i];
            if (i < 3) v += W1[c * 12] * t[i * 3] + W1[c * 12 + 1] * t[i * 3 + 1] + W1[c * 12 + 2] * t[i * 3 + 2];
            v *= scale[c];
        }
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const int off = ((i >> 3) * 64 + c) * 16 + (i & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(out + off) = h;
        *reinterpret_cast<__nv_bfloat16*>(out + 64 * 16 * 2 + off) = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

__global__ void t32_bias_kernel(const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                                int n, int n_pad, float* __restrict__ dst) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_pad) dst[i] = i < n ? (scale ? scale[i] : 1.f) * (bias ? bias[i] : 0.f) + (shift ? shift[i] : 0.f) : 0.f;
}

}  // namespace

int t32_affine_bias(const float* bias, const float* scale, const float* shift, int n, int n_pad, float* dst, cudaStream_t st) {
    launch_pdl(t32_bias_kernel, dim3((unsigned)((n_pad + 127) / 128)), dim3(128), 0, st, bias, scale, shift, n, n_pad, dst);
    count_launch();
    return check_launch("t32_bias_kernel");
}

int t32_fold_w1(const float* W1, const float* T, const float* scale, int n_clouds, unsigned char* dst, long long dst_stride, cudaStream_t st) {
    launch_pdl(t32_fold_w1_kernel, dim3((unsigned)n_clouds), dim3(256), 0, st, W1, T, scale, dst, dst_stride);
    count_launch();
    return check_launch("t32_fold_w1_kernel");
}

int tc_chain32_launch(const T32Params& p, cudaStream_t st) {
    if (p.n_ops < 1 || p.n_ops > kT32MaxOps) return fail(AMP_E_BADARG, "tc_chain32: bad op count %d", p.n_ops);
    int stage_f32 = 0, stream = 0;
    for (int l = 0; l < p.n_ops; ++l) {
        const T32Op& o = p.op[l];
        if (o.K < 16 || o.K > 128 || o.K % 16 || o.N < 16 || o.N % 16 || (!o.pool && o.N > 128) || (o.pool && o.N != 128 && o.N != 256))
            return fail(AMP_E_BADARG, "tc_chain32: op %d has unsupported shape K=%d N=%d", l, o.K, o.N);
        if (o.pool && (!o.relu || !p.pool || o.bias_off < 0 || o.write_act || l + 1 != p.n_ops))
            return fail(AMP_E_BADARG, "tc_chain32: op %d cannot pool", l);
        if (o.w_stream && (!o.pool || o.N != 256 || o.K != 128 || !p.wstream || ((uintptr_t)p.wstream & 15)))
            return fail(AMP_E_BADARG, "tc_chain32: op %d cannot stream its weights", l);
        if (o.write_act && (l + 1 >= p.n_ops || p.op[l + 1].K != o.N)) return fail(AMP_E_BADARG, "tc_chain32: op %d does not feed op %d", l, l + 1);
        if (o.w_off % 128) return fail(AMP_E_BADARG, "tc_chain32: op %d weights are not 128-byte aligned", l);
        if (!o.w_stream && o.w_off + o.N * o.K * 4 > (o.w_cloud ? p.wcloud_bytes : p.wblob_bytes))
            return fail(AMP_E_BADARG, "tc_chain32: op %d weights lie outside the packed buffer", l);
        if (o.bias_off >= 0 && (o.bias_off % 4 || o.bias_off + o.N > p.n_bias || !p.bias))
            return fail(AMP_E_BADARG, "tc_chain32: op %d bias outside the table", l);
        if (o.store_logits && (!p.logits || p.n_classes < 1 || p.n_classes > o.N || p.n_classes > 32))
            return fail(AMP_E_BADARG, "tc_chain32: op %d cannot store logits", l);
        if (o.store_f32 && (o.N != 64 || !p.out_f32 || p.out_ld % 4 || p.out_col0 % 4 || ((uintptr_t)p.out_f32 & 15)))
            return fail(AMP_E_BADARG, "tc_chain32: op %d output rows are not 16-byte aligned 64-channel rows", l);
        if (o.bias_grouped && (!p.gbias || ((uintptr_t)p.gbias & 15) || (p.n_groups > 1 && !p.group_rows)))
            return fail(AMP_E_BADARG, "tc_chain32: op %d grouped bias missing", l);
        stage_f32 |= o.store_f32; stream |= o.w_stream;
    }
    if (p.in_mode == 0) {
        if (p.op[0].K != 16 || p.in_k < 1 || p.in_k > 10) return fail(AMP_E_BADARG, "tc_chain32: narrow input needs K = 16 and at most 10 columns");
    } else if (p.op[0].K != 64 || p.in_ld % 4 || ((uintptr_t)p.in_x & 15)) {
        return fail(AMP_E_BADARG, "tc_chain32: wide input needs K = 64 and 16-byte aligned rows");
    }
    if (p.wblob_bytes % 16 || p.wcloud_bytes % 16 || p.wcloud_stride % 16 || ((uintptr_t)p.wblob & 15) || ((uintptr_t)p.wcloud & 15))
        return fail(AMP_E_BADARG, "tc_chain32: packed weights are not 16-byte aligned");
    if (p.n_clouds < 1 || p.rows_per_cloud < 1) return fail(AMP_E_BADARG, "tc_chain32: empty input");
    const Plan sp = plan_of(p.wblob_bytes, p.wcloud_bytes, p.n_bias, stage_f32, stream);
    if (sp.total > kMaxSmem) return fail(AMP_E_BADARG, "tc_chain32: chain needs %d bytes of shared memory", sp.total);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc_chain32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem);
        if (e != cudaSuccess) return fail(AMP_E_CUDA, "tc_chain32: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    const long long n_tiles = (long long)p.n_clouds * ((p.rows_per_cloud + kRows - 1) / kRows);
    const int grid = (int)((n_tiles + 1) / 2 < kNumSMs ? (n_tiles + 1) / 2 : kNumSMs);
    const int smem_bytes = sp.total < kMinSmem ? kMinSmem : sp.total;
    static const bool want_prof = getenv("AMP_CHAIN32_PROF") != nullptr;
    if (want_prof) {                                     // debugging aid: synchronous, prints the phase timeline of CTA 0 / thread 0
        static long long* dprof = nullptr;
        if (!dprof) cudaMalloc(&dprof, 256 * sizeof(long long));
        cudaMemsetAsync(dprof, 0, 256 * sizeof(long long), st);
        T32Params q = p;
        q.prof = dprof;
        launch_pdl(tc_chain32_kernel, dim3((unsigned)grid), dim3(kThreads), smem_bytes, st, q);
        long long h[256];
        cudaMemcpyAsync(h, dprof, sizeof h, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        fprintf(stderr, "[tc_chain32 prof] ops=%d tiles=%lld smem=%d:", p.n_ops, n_tiles, smem_bytes);
        for (int i = 1; i < (int)h[255] && i < 250; ++i) fprintf(stderr, " %lld", h[i] - h[i - 1]);
        fprintf(stderr, "\n");
        count_launch();
        return check_launch("tc_chain32_kernel");
    }
    launch_pdl(tc_chain32_kernel, dim3((unsigned)grid), dim3(kThreads), smem_bytes, st, p);
    count_launch();
    count_path("tc_chain32");
    return check_launch("tc_chain32_kernel");
}

int t32_pack_weights(const T32PackTable& t, unsigned char* dst, cudaStream_t st) {
    if (t.n < 1 || t.n > T32PackTable::kMax) return fail(AMP_E_BADARG, "t32_pack_weights: bad job count");
    int mx = 0;
    for (int i = 0; i < t.n; ++i) {
        const T32PackJob& j = t.job[i];
        if (j.Npad % 16 || j.Kpad % 16 || j.dst_off % 128 || (j.chunk_n && j.Npad % j.chunk_n))
            return fail(AMP_E_BADARG, "t32_pack_weights: job %d is not tile aligned", i);
        if (j.Npad * j.Kpad > mx) mx = j.Npad * j.Kpad;
    }
    int bx = (mx + 255) / 256;
    if (bx > 32) bx = 32;
    launch_pdl(t32_pack_kernel, dim3(bx, t.n, t.n_clouds < 1 ? 1 : t.n_clouds), dim3(256), 0, st, t, dst);
    count_launch();
    return check_launch("t32_pack_kernel");
}

}  // namespace amp
