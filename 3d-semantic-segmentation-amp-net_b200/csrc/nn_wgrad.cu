// Weight / bias gradients of the point-wise linear layers in fp32 on the CUDA cores:
//   dW[n, k] = sum_r dY[r, n] * A[r, k],  db[n] = sum_r dY[r, n]
// (the autograd backward of nn.Conv1d(k=1) / nn.Linear / torch.bmm of pointNet/model/pointnetAtt.py, driven
// by loss.backward() in pointNet/self-attention/train_pointnet-attention.py:467).
// Both operands take the same prologues as pw_linear, so neither dY (BatchNorm backward applied) nor A
// (BatchNorm + ReLU + Dropout applied) is materialised.
//
// Deterministic reduction over the rows: CTA (slab of 512 rows of one cloud) x (128 x 64 tile of dW)
// writes a partial; wgrad_reduce sums the partials in a fixed order.
#include "nn_common.cuh"

namespace amp {
namespace {

constexpr int kDefaultSlab = 512, TNo = 128, TKo = 64, RB = 16, NT = 256, PAD = 4;

__global__ void __launch_bounds__(NT)
wgrad_partial_kernel(const WgParams p, int n_tiles, int k_tiles, int slabs, int SLAB) {
    pdl_sync();
    __shared__ __align__(16) float Ys[2][RB][TNo + PAD];
    __shared__ __align__(16) float As[2][RB][TKo + PAD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int slab = blockIdx.x, cloud = blockIdx.z;
    const int nt = blockIdx.y / k_tiles, kt = blockIdx.y % k_tiles;
    const int n0 = nt * TNo, k0 = kt * TKo;
    const int rows = p.rows_per_cloud;
    const int r_begin = slab * SLAB, r_end = min(rows, r_begin + SLAB);
    const long long cloud_row = (long long)cloud * rows;

    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) bsum[i] = 0.f;
    const bool do_bias = (kt == 0) && (p.db != nullptr || p.dbg != nullptr);

    float yr[8], ar[4];
    auto g_load = [&](int rb) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int e = tid + i * NT, n = e & (TNo - 1), r = rb + (e >> 7);
            float v = 0.f;
            if (r < r_end && n0 + n < p.Nout) {
                const long long off = p.dy_transposed ? ((long long)cloud * p.Nout + n0 + n) * rows + r
                                                      : (cloud_row + r) * p.lddy + n0 + n;
                v = __ldg(p.dY + off);
                if (p.y_a) v = fmaf(v, __ldg(p.y_a + n0 + n), __ldg(p.y_b + n0 + n));
                if (p.Y2) v = fmaf(__ldg(p.Y2 + off) - (p.y_m ? __ldg(p.y_m + n0 + n) : 0.f), __ldg(p.y_c + n0 + n), v);
            }
            yr[i] = v;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * NT, k = e & (TKo - 1), r = rb + (e >> 6);
            float v = 0.f;
            if (r < r_end && k0 + k < p.K) {
                v = __ldg(p.A + (cloud_row + r) * p.lda + k0 + k);
                if (p.a_a) v = fmaf(v - (p.a_m ? __ldg(p.a_m + k0 + k) : 0.f), __ldg(p.a_a + k0 + k), __ldg(p.a_b + k0 + k));
                if (p.a_relu) v = fmaxf(v, 0.f);
                if (p.a_drop_p > 0.f)
                    v *= dropout_keep(eff_seed(p.a_drop_seed, p.drop_off), (unsigned long long)(cloud_row + r) * p.K + k0 + k, p.a_drop_p);
            }
            ar[i] = v;
        }
    };
    auto s_store = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int e = tid + i * NT; Ys[buf][e >> 7][e & (TNo - 1)] = yr[i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) { const int e = tid + i * NT; As[buf][e >> 6][e & (TKo - 1)] = ar[i]; }
    };

    const int nsteps = (r_end - r_begin + RB - 1) / RB;
    if (nsteps > 0) {
        g_load(r_begin);
        s_store(0);
    }
    __syncthreads();
    for (int s = 0; s < nsteps; ++s) {
        const int buf = s & 1;
        if (s + 1 < nsteps) g_load(r_begin + (s + 1) * RB);
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const float4 y0 = *reinterpret_cast<const float4*>(&Ys[buf][r][ty * 8]);
            const float4 y1 = *reinterpret_cast<const float4*>(&Ys[buf][r][ty * 8 + 4]);
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][r][tx * 4]);
            const float y[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
            const float a[4] = {a0.x, a0.y, a0.z, a0.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(y[i], a[j], acc[i][j]);
            }
            if (do_bias && tx == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) bsum[i] += y[i];
            }
        }
        if (s + 1 < nsteps) s_store(buf ^ 1);
        __syncthreads();
    }
    const long long chunk = (long long)cloud * slabs + slab;
    float* part = p.partials + chunk * ((long long)p.Nout * p.K + p.Nout);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int n = n0 + ty * 8 + i;
        if (n >= p.Nout) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + tx * 4 + j;
            if (k < p.K) part[(long long)n * p.K + k] = acc[i][j];
        }
        if (do_bias && tx == 0) part[(long long)p.Nout * p.K + n] = bsum[i];
    }
}

// one thread per dW element (and per db / dbg element): fixed-order sum over the chunks
// fixed-order sum of `count` floats spaced `stride` apart; the loads of 8 terms are issued before they are added
__device__ __forceinline__ float chunk_sum(const float* __restrict__ src, long long stride, long long count) {
    float s = 0.f;
    long long ch = 0;
    for (; ch + 8 <= count; ch += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (ch + u) * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; ch < count; ++ch) s += __ldg(src + ch * stride);
    return s;
}

// A block reduces 32 consecutive outputs: lanes = outputs (coalesced reads of every partial), the 8 warps take 8 consecutive
// slices of the chunk list, so a thread's loads are all in flight at once and even a 64 x 64 layer fills the machine; the
// slices are then added in a fixed order (deterministic).
__device__ __forceinline__ void wgrad_reduce_body(const WgParams& p, int slabs, int SLAB, long long vblock, long long vgrid, float (*red)[32]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long per = (long long)p.Nout * p.K + p.Nout;
    const int out_clouds = p.per_cloud ? p.n_clouds : 1;
    const long long n_w = (long long)out_clouds * p.Nout * p.K;
    const long long n_b = p.db ? (long long)out_clouds * p.Nout : 0;
    const long long n_g = p.dbg ? (long long)p.n_clouds * p.n_groups * p.Nout : 0;
    const long long count = (long long)(p.per_cloud ? 1 : p.n_clouds) * slabs;       // chunks per output
    const long long cs = (count + 7) / 8, ch_lo = min(count, w * cs), ch_hi = min(count, ch_lo + cs);
    for (long long base = vblock * 32; base < n_w + n_b; base += vgrid * 32) {
        const long long i = base + lane;
        const float* src = nullptr;
        float* dst = nullptr;
        if (i < n_w) {
            const int k = (int)(i % p.K), n = (int)((i / p.K) % p.Nout), c = (int)(i / ((long long)p.K * p.Nout));
            src = p.partials + (long long)(p.per_cloud ? c : 0) * slabs * per + (long long)n * p.K + k;
            dst = p.dW + (long long)c * p.w_cloud_stride + (p.w_kn ? (long long)k * p.ldw + n : (long long)n * p.ldw + k);
        } else if (i < n_w + n_b) {
            const long long e = i - n_w;
            const int n = (int)(e % p.Nout), c = (int)(e / p.Nout);
            src = p.partials + (long long)(p.per_cloud ? c : 0) * slabs * per + (long long)p.Nout * p.K + n;
            dst = p.db + e;
        }
        red[w][lane] = src ? chunk_sum(src + ch_lo * per, per, ch_hi - ch_lo) : 0.f;
        __syncthreads();
        if (w == 0 && dst) {
            float s = red[0][lane];
#pragma unroll
            for (int q = 1; q < 8; ++q) s += red[q][lane];
            *dst = p.accumulate ? *dst + s : s;
        }
        __syncthreads();
    }
    // per-(cloud, block) bias gradients of the head: a handful of slabs each
    for (long long e = vblock * (long long)blockDim.x + threadIdx.x; e < n_g; e += vgrid * blockDim.x) {
        const int n = (int)(e % p.Nout), g = (int)((e / p.Nout) % p.n_groups), c = (int)(e / ((long long)p.Nout * p.n_groups));
        const int r_lo = p.group_rows[g], r_hi = (g + 1 < p.n_groups) ? p.group_rows[g + 1] : p.rows_per_cloud;
        float s = 0.f;
        for (int sl = r_lo / SLAB; sl < (r_hi + SLAB - 1) / SLAB; ++sl)
            s += p.partials[((long long)c * slabs + sl) * per + (long long)p.Nout * p.K + n];
        p.dbg[e] = s;
    }
}


__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const WgParams p, int slabs, int SLAB) {
    pdl_sync();
    __shared__ float red[8][32];
    wgrad_reduce_body(p, slabs, SLAB, blockIdx.x, gridDim.x, red);
}

// the queued reductions of a WgDeferScope in one launch: block -> (job, block within the job)
constexpr int kWgMaxJobs = 24;
struct WgJobs { int n; int block_end[kWgMaxJobs]; int slabs[kWgMaxJobs]; int SLAB[kWgMaxJobs]; WgParams p[kWgMaxJobs]; };
__global__ void __launch_bounds__(256) wgrad_reduce_jobs_kernel(const __grid_constant__ WgJobs jobs) {
    pdl_sync();
    __shared__ float red[8][32];
    int j = 0;
    while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.block_end[j]) ++j;       // block-uniform
    const int b0 = j ? jobs.block_end[j - 1] : 0;
    wgrad_reduce_body(jobs.p[j], jobs.slabs[j], jobs.SLAB[j], (int)blockIdx.x - b0, jobs.block_end[j] - b0, red);
}

thread_local WgDeferScope* t_defer = nullptr;
thread_local WgJobs t_jobs;

}  // namespace

size_t wgrad_workspace_floats(int n_clouds, int rows_per_cloud, int Nout, int K, int slab_rows) {
    const int SLAB = slab_rows > 0 ? slab_rows : kDefaultSlab;
    const size_t slabs = (size_t)(rows_per_cloud + SLAB - 1) / SLAB;
    return (size_t)n_clouds * slabs * ((size_t)Nout * K + Nout);
}

WgDeferScope::WgDeferScope(float* pool_, size_t pool_floats_, cudaStream_t st_) : pool(pool_), pool_floats(pool_floats_), used(0), st(st_), prev(t_defer) {
    t_defer = this;
    t_jobs.n = 0;
}
WgDeferScope::~WgDeferScope() { t_defer = prev; }
int WgDeferScope::flush() {
    WgJobs& jb = t_jobs;
    used = 0;
    if (jb.n == 0) return AMP_OK;
    const int blocks = jb.block_end[jb.n - 1];
    launch_pdl(wgrad_reduce_jobs_kernel, dim3((unsigned)blocks), dim3(256), 0, st, jb);
    jb.n = 0;
    count_launch();
    return check_launch("wgrad_reduce_jobs");
}

int wgrad(const WgParams& p_in, cudaStream_t st) {
    if (!p_in.dY || !p_in.A || !p_in.dW || !p_in.partials) return fail(AMP_E_BADARG, "wgrad: null operand");
    // a parameter gradient inside a WgDeferScope: partials into a slice of the scope's pool, reduction queued
    WgParams p = p_in;
    bool deferred = false;
    if (t_defer && t_defer->st == st && !p.per_cloud && !p.accumulate && !p.dbg && p.n_clouds >= 1 && p.rows_per_cloud >= 1 && p.Nout >= 1 && p.K >= 1) {
        const size_t need = (wgrad_workspace_floats(p.n_clouds, p.rows_per_cloud, p.Nout, p.K, p.slab_rows) + 63) & ~(size_t)63;
        if (need <= t_defer->pool_floats) {
            if (t_defer->used + need > t_defer->pool_floats || t_jobs.n == kWgMaxJobs) {
                const int rc = t_defer->flush();
                if (rc != AMP_OK) return rc;
            }
            p.partials = t_defer->pool + t_defer->used;
            p.partial_floats = need;
            deferred = true;
        }
    }
    if (p.K < 1 || p.Nout < 1 || p.n_clouds < 1 || p.rows_per_cloud < 1) return fail(AMP_E_BADARG, "wgrad: bad shape");
    if (p.n_clouds > 65535) return fail(AMP_E_BADARG, "wgrad: more than 65535 clouds in one launch");
    const int SLAB = p.slab_rows > 0 ? p.slab_rows : kDefaultSlab;
    if (p.partial_floats < wgrad_workspace_floats(p.n_clouds, p.rows_per_cloud, p.Nout, p.K, p.slab_rows))
        return fail(AMP_E_WORKSPACE, "wgrad: partial-sum workspace too small");
    if ((p.y_a == nullptr) != (p.y_b == nullptr) || (p.a_a == nullptr) != (p.a_b == nullptr))
        return fail(AMP_E_BADARG, "wgrad: scale and shift must come together");
    if (p.Y2 && (!p.y_a || !p.y_c)) return fail(AMP_E_BADARG, "wgrad: Y2 needs y_a, y_b, y_c");
    if (p.dbg && (!p.group_rows || p.n_groups < 1)) return fail(AMP_E_BADARG, "wgrad: dbg needs group_rows");
    const int slabs = (p.rows_per_cloud + SLAB - 1) / SLAB;
    const int n_tiles = (p.Nout + TNo - 1) / TNo, k_tiles = (p.K + TKo - 1) / TKo;
    int rc = small_wgrad_try(p_in, st);                  // few rows: direct deterministic kernel, no partials (nn_small.cu)
    if (rc != 0) return rc < 0 ? rc : AMP_OK;
    rc = narrow_wgrad_try(p, slabs, SLAB, st);           // K <= 16 over many rows: memory-bound exact fp32 partial pass
    if (rc == 0) rc = narrow_out_wgrad_try(p, slabs, SLAB, st);   // Nout <= 8 (class logits)
    if (rc == 0) rc = tc_wgrad_try(p, slabs, SLAB, st);  // tensor-core partial pass when the shape allows (nn_tc_wgrad.cu)
    if (rc < 0) return rc;
    if (rc == 0) {
        dim3 grid(slabs, n_tiles * k_tiles, p.n_clouds);
        launch_pdl(wgrad_partial_kernel, grid, dim3(NT), 0, st, p, n_tiles, k_tiles, slabs, SLAB);
        count_launch();
        count_path("wgrad_partial");
        rc = check_launch("wgrad_partial");
        if (rc) return rc;
    }
    const long long total = (long long)(p.per_cloud ? p.n_clouds : 1) * p.Nout * (p.K + 1) +
                            (p.dbg ? (long long)p.n_clouds * p.n_groups * p.Nout : 0);
    long long blocks = (total + 31) / 32;                  // 32 outputs per block
    if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
    if (deferred) {
        WgJobs& jb = t_jobs;
        const int j = jb.n++;
        jb.p[j] = p; jb.slabs[j] = slabs; jb.SLAB[j] = SLAB;
        jb.block_end[j] = (j ? jb.block_end[j - 1] : 0) + (int)blocks;
        t_defer->used += p.partial_floats;
        return AMP_OK;
    }
    launch_pdl(wgrad_reduce_kernel, dim3((unsigned)((unsigned)blocks)), dim3(256), 0, st, p, slabs, SLAB);
    count_launch();
    return check_launch("wgrad_reduce");
}

// slab size for row groups: the largest divisor of every group size that is <= 512 (group boundaries must fall
// on slab boundaries for dbg)
int wgrad_group_slab(const int* group_sizes, int n_groups) {
    long long g = 0;
    for (int i = 0; i < n_groups; ++i) { long long a = group_sizes[i], b = g; while (b) { long long t = a % b; a = b; b = t; } g = a; }
    int best = 1;
    for (int d = 1; d <= kDefaultSlab; ++d) if (g % d == 0) best = d;
    return best;
}

}  // namespace amp
