// Fused shared-MLP chain kernel on tcgen05 tensor cores (see tc_chain.cuh for the contract).
//
// Layout of one CTA (256 threads = 2 independent warpgroups, 1 CTA per SM):
//   * the packed bf16 weights of the whole chain live in shared memory for the CTA's lifetime (TMA bulk copies
//     global -> shared, completion on an mbarrier);
//   * each warpgroup owns a "slot": a 32 KB activation buffer (the K-major A operand of the next layer), a region for
//     per-cloud weights, 256 TMEM columns of fp32 accumulators and one mbarrier. It walks its own stream of 128-point
//     tiles: stage input -> for every layer { one thread issues the tcgen05.mma K-steps and commits to the mbarrier;
//     all 128 threads wait, read the accumulator with tcgen05.ld (thread == TMEM lane == point row), add bias, ReLU,
//     write bf16 back as the next A operand } . While one warpgroup is in its epilogue the other one's MMAs run, so
//     tensor pipe and CUDA cores overlap without any cross-warpgroup synchronisation.
//   * pooled layers run TRANSPOSED: D^T[channel, point] = W[channel, :] . act[point, :], so that TMEM lanes are
//     channels and the max over the points of the tile is a per-thread register reduction (no shuffles); one
//     atomicMax per (cloud, channel, tile).
// Operand layout in shared memory (both A and B): K-major, no swizzle, 8 x 16-byte core matrices:
//   element (row r, k) at byte ((k / 8) * rows + r) * 16 + (k % 8) * 2      -> SBO = 128 B, LBO = rows * 16 B.
#include "tc_chain.cuh"
#include "tc_ptx.cuh"

namespace amp {
namespace {
using namespace tcx;

constexpr int kThreads = 256;
constexpr int kActBytes = kTcTileRows * 128 * 2;          // 128 rows x up to 128 channels of bf16
constexpr int kTmemCols = 512;
constexpr int kSlotCols = 256;
constexpr int kMaxSmem = 232448;                          // 227 KB per CTA on sm_100
constexpr int kMinSmem = 120 * 1024;                      // more than half an SM: one CTA (one TMEM owner) per SM

__host__ __device__ inline int align_i(int v, int a) { return (v + a - 1) / a * a; }

// ---------------------------------------------------------------------------------------------------------------
// the chain kernel
// ---------------------------------------------------------------------------------------------------------------
struct SmemPlan { int w, wc, act0, tab, bar, total; };
__host__ __device__ inline SmemPlan smem_plan(int wblob_bytes, int wcloud_bytes, int n_table_floats) {
    SmemPlan s;
    s.w = align_i(wblob_bytes, 128);
    s.wc = align_i(wcloud_bytes, 128);
    s.act0 = s.w + 2 * s.wc;
    s.tab = s.act0 + 2 * kActBytes;
    s.bar = s.tab + align_i(n_table_floats * 4, 16);
    s.total = s.bar + 64;
    return s;
}

__global__ void __launch_bounds__(kThreads, 1) tc_chain_kernel(const __grid_constant__ TcChainParams p) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, wg = warp >> 2, wtid = tid & 127;
    const SmemPlan sp = smem_plan(p.wblob_bytes, p.wcloud_bytes, p.n_table_floats);
    unsigned char* s_w = smem;
    unsigned char* s_wc = smem + sp.w + wg * sp.wc;
    unsigned char* s_act = smem + sp.act0 + wg * kActBytes;
    float* s_tab = reinterpret_cast<float*>(smem + sp.tab);
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + sp.bar);       // [0] weights, [1 + wg] MMA completion
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + sp.bar + 32);
    const uint32_t wbar = smem_u32(&s_bar[0]), mbar = smem_u32(&s_bar[1 + wg]);

    if (tid == 0) {
        mbar_init(wbar, 1);
        mbar_init(smem_u32(&s_bar[1]), 1);
        mbar_init(smem_u32(&s_bar[2]), 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(s_tmem), kTmemCols);
    for (int i = tid; i < p.n_table_floats; i += kThreads) s_tab[i] = __ldg(p.tables + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *s_tmem;
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
    const uint32_t slot_col = tmem_base + (uint32_t)(wg * kSlotCols);
    const uint32_t act_addr = smem_u32(s_act);
    uint32_t phase = 0;
    int cur_cloud = -1;

    for (int tile = blockIdx.x * 2 + wg; tile < n_tiles; tile += gridDim.x * 2) {
        const int cloud = tile / tiles_per_cloud;
        const int row0 = (tile - cloud * tiles_per_cloud) * kTcTileRows;
        const int valid = min(kTcTileRows, rows - row0);
        const bool row_ok = row < valid;
        const long long grow = (long long)cloud * rows + row0 + row;

        // ---- per-cloud weights (all MMAs of the previous tile have completed: safe to overwrite) ----
        if (p.wcloud_bytes > 0 && cloud != cur_cloud) {
            const uint4* src = reinterpret_cast<const uint4*>(p.wcloud + (long long)cloud * p.wcloud_stride);
            uint4* dst = reinterpret_cast<uint4*>(s_wc);
            for (int i = wtid; i < p.wcloud_bytes / 16; i += 128) dst[i] = __ldg(src + i);
            cur_cloud = cloud;
        }
        // ---- input stage: fp32 rows -> bf16 A operand [K/8][128][8] ----
        if (p.in_mode == 0) {
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
            uint4* dst = reinterpret_cast<uint4*>(s_act);
            if (p.op[0].K == 16) {          // in_k <= 3 ... 8: k 0..7 = hi, k 8..15 = lo
                dst[0 * 128 + row] = make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(hi[4], hi[5]),
                                                pack_bf16x2(hi[6], hi[7]));
                dst[1 * 128 + row] = make_uint4(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]),
                                                pack_bf16x2(lo[6], lo[7]));
            } else {                        // K == 32: k 0..15 = hi (9 used), k 16..31 = lo
                dst[0 * 128 + row] = make_uint4(pack_bf16x2(hi[0], hi[1]), pack_bf16x2(hi[2], hi[3]), pack_bf16x2(hi[4], hi[5]),
                                                pack_bf16x2(hi[6], hi[7]));
                dst[1 * 128 + row] = make_uint4(pack_bf16x2(hi[8], 0.f), 0u, 0u, 0u);
                dst[2 * 128 + row] = make_uint4(pack_bf16x2(lo[0], lo[1]), pack_bf16x2(lo[2], lo[3]), pack_bf16x2(lo[4], lo[5]),
                                                pack_bf16x2(lo[6], lo[7]));
                dst[3 * 128 + row] = make_uint4(pack_bf16x2(lo[8], 0.f), 0u, 0u, 0u);
            }
        } else {
            const float4* src = reinterpret_cast<const float4*>(p.in_x + grow * p.in_ld);
            uint4* dst = reinterpret_cast<uint4*>(s_act);
            const int chunks = p.op[0].K >> 3;
            for (int c = 0; c < chunks; ++c) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
                if (row_ok) { a = __ldg(src + 2 * c); b = __ldg(src + 2 * c + 1); }
                dst[c * 128 + row] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
            }
        }
        // group of this row (per-block bias of the segmentation head)
        int group = 0;
        if (p.n_groups > 1) {
            const int r = row0 + (row_ok ? row : 0);
            for (int g = 1; g < p.n_groups; ++g) group += (r >= __ldg(p.group_rows + g)) ? 1 : 0;
        }
        fence_proxy_async();
        tc_fence_before();
        wg_bar_sync(wg);

        for (int l = 0; l < p.n_ops; ++l) {
            const TcOp& op = p.op[l];
            // ---- MMA issue: one thread per warpgroup ----
            if (wtid == 0) {
                if (!w_ready) { mbar_wait(wbar, 0); w_ready = true; }
                tc_fence_after();
                const uint32_t wbase = op.w_cloud ? smem_u32(s_wc) + (uint32_t)op.w_off : smem_u32(s_w) + (uint32_t)op.w_off;
                const uint32_t w_lbo = (uint32_t)op.N * 16u;
                const int ksteps = op.K >> 4;
                if (!op.pool) {
                    const uint32_t idesc = umma_idesc(128, op.N);
                    for (int k = 0; k < ksteps; ++k)
                        umma_bf16(slot_col, umma_desc(act_addr + (uint32_t)k * 4096u, 2048u, 128u),
                                  umma_desc(wbase + (uint32_t)k * 2u * w_lbo, w_lbo, 128u), idesc, k > 0);
                } else {
                    const uint32_t idesc = umma_idesc(128, 128);
                    for (int mt = 0; mt < (op.N >> 7); ++mt)
                        for (int k = 0; k < ksteps; ++k)
                            umma_bf16(slot_col + (uint32_t)mt * 128u,
                                      umma_desc(wbase + (uint32_t)mt * 2048u + (uint32_t)k * 2u * w_lbo, w_lbo, 128u),
                                      umma_desc(act_addr + (uint32_t)k * 4096u, 2048u, 128u), idesc, k > 0);
                }
                umma_commit(mbar);
            }
            __syncwarp();
            mbar_wait(mbar, phase);
            phase ^= 1u;
            tc_fence_after();

            // ---- epilogue ----
            if (!op.pool) {
                const float* gb = op.bias_grouped ? p.gbias + ((long long)cloud * p.n_groups + group) * op.N : nullptr;
                for (int c0 = 0; c0 < op.N; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(slot_col + lane_addr + (uint32_t)c0, v);
                    tmem_wait_ld();
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    const int nc = min(32, op.N - c0);           // 16 or 32
                    if (op.bias_off >= 0) {
                        const float4* b4 = reinterpret_cast<const float4*>(s_tab + op.bias_off + c0);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (q * 4 < nc) {
                                const float4 b = b4[q];
                                f[4 * q] += b.x; f[4 * q + 1] += b.y; f[4 * q + 2] += b.z; f[4 * q + 3] += b.w;
                            }
                        }
                    }
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
                    if (op.write_act) {
                        uint4* dst = reinterpret_cast<uint4*>(s_act);
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            if (q * 8 < nc)
                                dst[((c0 >> 3) + q) * 128 + row] =
                                    make_uint4(pack_bf16x2(f[8 * q], f[8 * q + 1]), pack_bf16x2(f[8 * q + 2], f[8 * q + 3]),
                                               pack_bf16x2(f[8 * q + 4], f[8 * q + 5]), pack_bf16x2(f[8 * q + 6], f[8 * q + 7]));
                    }
                    if (op.store_logits && c0 == 0 && row_ok) {
#pragma unroll
                        for (int n = 0; n < 32; ++n)
                            if (n < p.n_classes)
                                p.logits[((long long)cloud * p.n_classes + n) * rows + row0 + row] = f[n];
                    }
                }
            } else {
                for (int mt = 0; mt < (op.N >> 7); ++mt) {
                    float m = -3.0e38f;
                    for (int c0 = 0; c0 < kTcTileRows; c0 += 32) {
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
                    const int ch = mt * 128 + row;
                    float r = m + (op.bias_off >= 0 ? s_tab[op.bias_off + ch] : 0.f);
                    r = fmaxf(r, 0.f);
                    atomicMax(p.pool + (long long)cloud * op.N + ch, __float_as_uint(r));
                }
            }
            // accumulator reads and A-operand writes of this layer are done before the next MMA is issued
            tc_fence_before();
            fence_proxy_async();
            wg_bar_sync(wg);
        }
    }
    if (tid == 0 && !w_ready) mbar_wait(wbar, 0);        // never leave with a bulk copy in flight
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

__global__ void tc_pack_kernel(const TcPackTable t, unsigned char* __restrict__ dst) {
    const TcPackJob& j = t.job[blockIdx.y];
    const int cloud = blockIdx.z;
    if (cloud > 0 && j.src_cloud_stride == 0) return;
    const float* src = j.src + (long long)cloud * j.src_cloud_stride;
    __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(dst + j.dst_off + (long long)cloud * j.dst_cloud_stride);
    const int total = j.Npad * j.Kpad;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k8 = e & 7, n = (e >> 3) % j.Npad, kc = (e >> 3) / j.Npad;
        const int k = kc * 8 + k8;
        int ks = k < j.K ? k : -1;
        if (j.split_in_k) {
            const int half = j.Kpad >> 1;
            ks = k < j.split_in_k ? k : ((k >= half && k < half + j.split_in_k) ? k - half : -1);
        }
        float v = 0.f;
        if (n < j.N && ks >= 0) {
            v = j.transposed ? src[(long long)ks * j.ld + n] : src[(long long)n * j.ld + ks];
            if (j.scale) v *= j.scale[n];
        }
        out[e] = __float2bfloat16_rn(v);
    }
}

__global__ void tc_bias_kernel(const TcBiasTable t, float* __restrict__ dst) {
    const TcBiasJob& j = t.job[blockIdx.x];
    for (int i = threadIdx.x; i < j.npad; i += blockDim.x) {
        float v = 0.f;
        if (i < j.n) v = (j.scale ? j.scale[i] : 1.f) * (j.bias ? j.bias[i] : 0.f) + (j.shift ? j.shift[i] : 0.f);
        dst[j.dst_off + i] = v;
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
        if (o.bias_off >= 0 && (o.bias_off % 4 || o.bias_off + o.N > p.n_table_floats))
            return fail(AMP_E_BADARG, "tc_chain: op %d bias table out of range", l);
        if (o.w_off % 128) return fail(AMP_E_BADARG, "tc_chain: op %d weights are not 128-byte aligned", l);
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
    const SmemPlan sp = smem_plan(p.wblob_bytes, p.wcloud_bytes, p.n_table_floats);
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
    tc_chain_kernel<<<grid, kThreads, smem_bytes, st>>>(p);
    count_launch();
    return check_launch("tc_chain_kernel");
}

int tc_pack_weights(const TcPackTable& t, unsigned char* dst, cudaStream_t st) {
    if (t.n < 1 || t.n > TcPackTable::kMax) return fail(AMP_E_BADARG, "tc_pack_weights: bad job count");
    int mx = 0;
    for (int i = 0; i < t.n; ++i) {
        const int e = t.job[i].Npad * t.job[i].Kpad;
        if (e > mx) mx = e;
        if (t.job[i].Npad % 16 || t.job[i].Kpad % 16 || t.job[i].dst_off % 128)
            return fail(AMP_E_BADARG, "tc_pack_weights: job %d is not tile aligned", i);
    }
    int bx = (mx + 255) / 256;
    if (bx > 32) bx = 32;
    tc_pack_kernel<<<dim3(bx, t.n, t.n_clouds < 1 ? 1 : t.n_clouds), 256, 0, st>>>(t, dst);
    count_launch();
    return check_launch("tc_pack_kernel");
}

int tc_bias_tables(const TcBiasTable& t, float* dst, cudaStream_t st) {
    if (t.n < 1 || t.n > TcBiasTable::kMax) return fail(AMP_E_BADARG, "tc_bias_tables: bad job count");
    tc_bias_kernel<<<t.n, 128, 0, st>>>(t, dst);
    count_launch();
    return check_launch("tc_bias_kernel");
}

}  // namespace amp

// ---------------------------------------------------------------------------------------------------------------
// C ABI: one point-wise linear layer on the tensor cores (unit of the chains above, also used by the tests)
// ---------------------------------------------------------------------------------------------------------------
extern "C" {

size_t amp_tc_linear_workspace_bytes(int32_t K, int32_t N) {
    const int Kp = (K + 15) / 16 * 16, Np = (N + 15) / 16 * 16;
    return (size_t)amp::tc_packed_bytes(Np, Kp) + (size_t)Np * 4 + 512;
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
    const int wbytes = tc_packed_bytes(N, K);
    float* tab = reinterpret_cast<float*>(base + wbytes);
    TcPackTable pt{};
    pt.n = 1; pt.n_clouds = 1;
    pt.job[0] = TcPackJob{w, K, 0, nullptr, N, K, N, K, 0, 0, 0, 0};
    int rc = tc_pack_weights(pt, base, st);
    if (rc != AMP_OK) return rc;
    TcBiasTable bt{};
    bt.n = 1; bt.job[0] = TcBiasJob{bias, nullptr, nullptr, bias ? N : 0, N, 0};
    rc = tc_bias_tables(bt, tab, st);
    if (rc != AMP_OK) return rc;
    TcChainParams p{};
    p.n_ops = 1;
    p.op[0] = TcOp{K, N, 0, 0, 0, relu, 0, 0, pool_max ? 0 : 1, pool_max ? 1 : 0, 0};
    p.in_mode = 1; p.in_x = x; p.in_ld = K; p.in_k = K;
    p.tables = tab; p.n_table_floats = N;
    p.wblob = base; p.wblob_bytes = wbytes;
    p.n_groups = 1;
    p.out_f32 = y; p.out_ld = N; p.out_col0 = 0;
    p.pool = pool_max;
    p.n_clouds = (int)n_clouds; p.rows_per_cloud = (int)rows_per_cloud;
    return tc_chain_launch(p, st);
}

}  // extern "C"
