// Small kernels around the point-wise linear layers: BatchNorm folding / batch statistics, pooled
// value decode, T-Net identity and input-transform folding, the global-feature broadcast, the
// positional encoding and the per-(cloud, head) attention core.
// Reference: pointNet/model/pointnetAtt.py:28-47, 80-112, 176-209.
#include <algorithm>

#include "nn_common.cuh"

namespace amp {
namespace {

__global__ void bn_fold_eval_kernel(const BnTable table, float eps) {
    pdl_sync();
    const BnDesc d = table.d[blockIdx.x];
    for (int c = threadIdx.x; c < d.C; c += blockDim.x) {
        const float s = d.gamma[c] / sqrtf(d.var[c] + eps);
        d.scale[c] = s;
        d.shift[c] = d.beta[c] - d.mean[c] * s;
    }
}

// The two BatchNorm finalize kernels sit between every pair of layer kernels of a training step (~ 45 launches per
// step): they are pure latency. 128 threads per channel, two channels per CTA (32 .. 128 CTAs instead of 8 .. 32 with one
// warp per channel): a thread keeps 4 tiles in registers (one round of loads), fixed-order double sums through shuffles
// and shared memory (deterministic), no integer division per tile.
constexpr int BF_TPC = 128, BF_CH = 2, BF_REG = 4;

__device__ __forceinline__ double bf_channel_sum(double v, double* s_red, int ch, int li) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((li & 31) == 0) s_red[ch * (BF_TPC / 32) + (li >> 5)] = v;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < BF_TPC / 32; ++w) r += s_red[ch * (BF_TPC / 32) + w];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(BF_TPC * BF_CH)
bn_finalize_train_kernel(const float* __restrict__ part_sum, const float* __restrict__ part_m2, int tiles,
                         int tiles_per_cloud, int rows_per_cloud, long long count, int C,
                         const float* __restrict__ gamma, float* running_mean, float* running_var,
                         long long* nbt, float momentum, float eps, float* __restrict__ scale,
                         float* __restrict__ save_mean, float* __restrict__ save_invstd) {
    pdl_sync();
    // the per-tile (sum, M2) pairs are combined exactly (Chan et al.): M2 = sum_t [ M2_t + n_t (mean_t - mean)^2 ]
    __shared__ double s_red[BF_CH * (BF_TPC / 32)];
    const int ch = threadIdx.x / BF_TPC, li = threadIdx.x - ch * BF_TPC;
    const int c = blockIdx.x * BF_CH + ch;
    const bool ok = c < C;
    float ps[BF_REG], pq[BF_REG];
#pragma unroll
    for (int i = 0; i < BF_REG; ++i) {
        const int t = li + BF_TPC * i;
        ps[i] = (ok && t < tiles) ? part_sum[(long long)t * C + c] : 0.f;
        pq[i] = (ok && t < tiles) ? part_m2[(long long)t * C + c] : 0.f;
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < BF_REG; ++i) s += (double)ps[i];
    if (ok)
        for (int t = li + BF_TPC * BF_REG; t < tiles; t += BF_TPC) s += (double)part_sum[(long long)t * C + c];
    s = bf_channel_sum(s, s_red, ch, li);
    const double n = (double)count;
    const double mean = s / n;
    const int nt_last = rows_per_cloud - (tiles_per_cloud - 1) * 128;        // rows of the last tile of a cloud
    const double inv_full = 1.0 / 128.0, inv_last = 1.0 / (double)nt_last;     // (no fp64 divide per tile)
    // position of this thread's first tile inside its cloud, advanced by BF_TPC tiles per step (no division per tile)
    int pos = li % tiles_per_cloud;
    const int step = BF_TPC % tiles_per_cloud;
    double m2 = 0.0;
#pragma unroll
    for (int i = 0; i < BF_REG; ++i) {
        const int t = li + BF_TPC * i;
        if (ok && t < tiles) {
            const bool last = pos == tiles_per_cloud - 1;
            const double d = (double)ps[i] * (last ? inv_last : inv_full) - mean;
            m2 += (double)pq[i] + (double)(last ? nt_last : 128) * d * d;
        }
        pos += step; if (pos >= tiles_per_cloud) pos -= tiles_per_cloud;
    }
    if (ok)
        for (int t = li + BF_TPC * BF_REG; t < tiles; t += BF_TPC) {
            const bool last = pos == tiles_per_cloud - 1;
            const double d = (double)part_sum[(long long)t * C + c] * (last ? inv_last : inv_full) - mean;
            m2 += (double)part_m2[(long long)t * C + c] + (double)(last ? nt_last : 128) * d * d;
            pos += step; if (pos >= tiles_per_cloud) pos -= tiles_per_cloud;
        }
    m2 = bf_channel_sum(m2, s_red, ch, li);
    if (ok && li == 0) {
        const double var = m2 / n;
        const float invstd = (float)(1.0 / sqrt(var + (double)eps));
        scale[c] = gamma[c] * invstd;
        if (save_mean) { save_mean[c] = (float)mean; save_invstd[c] = invstd; }
        if (running_mean) {
            const double unbiased = count > 1 ? m2 / (n - 1.0) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
        }
    }
    if (nbt && blockIdx.x == 0 && threadIdx.x == 0) *nbt += 1;
}

__global__ void pool_decode_kernel(const unsigned long long* __restrict__ pmax, const unsigned long long* __restrict__ pmin,
                                   int mode, const float* __restrict__ scale, const float* __restrict__ shift,
                                   const float* __restrict__ mean, int total, int C, float* __restrict__ pooled,
                                   int* __restrict__ arg) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = i % C;
    unsigned long long key = pmax[i];
    float v;
    if (mode == 1) {
        v = unordered_bits((unsigned)(key >> 32));
    } else {
        const float s = scale[c], t = shift[c];
        if (s >= 0.f) {
            v = unordered_bits((unsigned)(key >> 32));
        } else {
            key = pmin[i];
            v = unordered_bits(~(unsigned)(key >> 32));
        }
        v = fmaxf(fmaf(v - (mean ? mean[c] : 0.f), s, t), 0.f);
    }
    pooled[i] = v;
    if (arg) arg[i] = (int)(0xffffffffu - (unsigned)(key & 0xffffffffull));
}

__global__ void add_identity_kernel(float* t, int total, int d) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int e = i % (d * d);
    if (e / d == e % d) t[i] += 1.f;
}

__global__ void fold_input_transform_kernel(const float* __restrict__ W1, const float* __restrict__ T, int n_clouds,
                                            float* __restrict__ W1eff) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;        // over clouds * 64 * 9
    if (i >= n_clouds * 64 * 9) return;
    const int b = i / (64 * 9), c = (i / 9) % 64, k = i % 9;
    float v = W1[c * 12 + 3 + k];
    if (k < 3) {
        const float* t = T + b * 9 + k * 3;                        // row k of T_b
        v += W1[c * 12 + 0] * t[0] + W1[c * 12 + 1] * t[1] + W1[c * 12 + 2] * t[2];
    }
    W1eff[i] = v;
}

__global__ void broadcast_rows_kernel(const float* __restrict__ g, int rows_per_cloud, int C4, float4* __restrict__ out,
                                      long long ldo4, long long total) {
    pdl_sync();
    // total = clouds * rows * C4 float4 elements
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4);
        const long long row = i / C4;
        const long long b = row / rows_per_cloud;
        out[row * ldo4 + c] = reinterpret_cast<const float4*>(g)[b * C4 + c];
    }
}

// 256 % C4 == 0: a thread keeps ONE column piece of its cloud in a register and only stores (a pure write stream: the generic
// kernel above spends three 64-bit divisions and a load per 16-byte store). grid = (row chunks, clouds).
__global__ void __launch_bounds__(256) broadcast_rows_cols_kernel(const float* __restrict__ g, int rows_per_cloud, int C4,
                                                                  float4* __restrict__ out, long long ldo4, int rows_per_cta) {
    pdl_sync();
    const int b = blockIdx.y, c = threadIdx.x % C4, rstep = 256 / C4;
    const float4 v = reinterpret_cast<const float4*>(g)[(long long)b * C4 + c];
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(rows_per_cloud, r0 + rows_per_cta);
    float4* o = out + ((long long)b * rows_per_cloud + r0 + threadIdx.x / C4) * ldo4 + c;
    for (int r = r0 + threadIdx.x / C4; r < r1; r += rstep, o += (long long)rstep * ldo4) __stcs(o, v);
}

__global__ void posenc_add_kernel(const float* __restrict__ gl, long long gl_ld, const float* __restrict__ cent, const float* __restrict__ w1,
                                  const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                  int n_clouds, int n_tokens, int E, float* __restrict__ tokens,
                                  float* __restrict__ h_pre) {
    pdl_sync();
    // one block per token (b, w)
    const int b = blockIdx.x / n_tokens, w = blockIdx.x % n_tokens;
    __shared__ float h[16];
    if (threadIdx.x < 16) {
        const float cx = cent[(b * n_tokens + w) * 2], cy = cent[(b * n_tokens + w) * 2 + 1];
        float v = w1[threadIdx.x * 2] * cx + w1[threadIdx.x * 2 + 1] * cy + b1[threadIdx.x];
        if (h_pre) h_pre[(b * n_tokens + w) * 16 + threadIdx.x] = v;
        h[threadIdx.x] = v > 0.f ? v : 0.01f * v;                 // F.leaky_relu_ default slope
    }
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        float v = b2[e];
#pragma unroll
        for (int k = 0; k < 16; ++k) v = fmaf(w2[e * 16 + k], h[k], v);
        tokens[((long long)b * n_tokens + w) * E + e] = gl[((long long)w * n_clouds + b) * gl_ld + e] + v;
    }
}

// one warp per (cloud, head, query token); lanes own head-dim elements (hd = 32 for E = 256, 8 heads)
__global__ void attention_core_kernel(const float* __restrict__ qkv, const unsigned char* __restrict__ key_mask,
                                      float drop_p, unsigned long long drop_seed_arg, const unsigned long long* drop_off,
                                      int n_clouds, int L, int E,
                                      int heads, float* __restrict__ out, float* __restrict__ probs) {
    pdl_sync();
    const unsigned long long drop_seed = eff_seed(drop_seed_arg, drop_off);
    const int hd = E / heads;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int total = n_clouds * heads * L;
    if (warp >= total) return;
    const int i = warp % L, h = (warp / L) % heads, b = warp / (L * heads);
    const float scale = rsqrtf((float)hd);
    const float* base = qkv + (long long)b * L * 3 * E;
    // scores: lane j handles key j, j + 32, ... (L <= 64 in practice; loop for generality)
    float m = -INFINITY;
    extern __shared__ float sm[];
    float* sc = sm + (threadIdx.x >> 5) * L;
    for (int j = lane; j < L; j += 32) {
        float s = 0.f;
        const float* q = base + (long long)i * 3 * E + h * hd;
        const float* k = base + (long long)j * 3 * E + E + h * hd;
        for (int d = 0; d < hd; ++d) s = fmaf(q[d] * scale, k[d], s);
        if (key_mask && key_mask[b * L + j]) s = -INFINITY;
        sc[j] = s;
        m = fmaxf(m, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float sum = 0.f;
    for (int j = lane; j < L; j += 32) {
        const float e = expf(sc[j] - m);
        sc[j] = e;
        sum += e;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncwarp();
    const float inv = 1.f / sum;
    const long long prow = (((long long)b * heads + h) * L + i) * L;
    for (int d = lane; d < hd; d += 32) {
        float o = 0.f;
        for (int j = 0; j < L; ++j) {
            float pj = sc[j] * inv;
            if (drop_p > 0.f) pj *= dropout_keep(drop_seed, (unsigned long long)(prow + j), drop_p);
            o = fmaf(pj, base[(long long)j * 3 * E + 2 * E + h * hd + d], o);
        }
        out[((long long)b * L + i) * E + h * hd + d] = o;
    }
    if (probs)
        for (int j = lane; j < L; j += 32) probs[prow + j] = sc[j] * inv;
}


// BatchNorm backward: per-channel sums -> dgamma, dbeta and the coefficients of dy = c1 dz + c2 y + c3
__global__ void __launch_bounds__(BF_TPC * BF_CH)
bn_backward_finalize_kernel(const float* __restrict__ part_sum, const float* __restrict__ part_sq, int tiles,
                            long long count, int C, const float* __restrict__ gamma,
                            const float* __restrict__ mean, const float* __restrict__ invstd,
                            float* dgamma, float* dbeta, int accumulate, float* __restrict__ c1,
                            float* __restrict__ c2, float* __restrict__ c3) {
    pdl_sync();
    __shared__ double s_red[BF_CH * (BF_TPC / 32)];
    const int ch = threadIdx.x / BF_TPC, li = threadIdx.x - ch * BF_TPC;
    const int c = blockIdx.x * BF_CH + ch;
    const bool ok = c < C;
    double s = 0.0, q = 0.0;
    {
        float ps[BF_REG], pq[BF_REG];             // all loads of the first 512 tiles in flight together
#pragma unroll
        for (int i = 0; i < BF_REG; ++i) {
            const int t = li + BF_TPC * i;
            ps[i] = (ok && t < tiles) ? part_sum[(long long)t * C + c] : 0.f;
            pq[i] = (ok && t < tiles) ? part_sq[(long long)t * C + c] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < BF_REG; ++i) { s += (double)ps[i]; q += (double)pq[i]; }
        if (ok)
            for (int t = li + BF_TPC * BF_REG; t < tiles; t += BF_TPC) {
                s += (double)part_sum[(long long)t * C + c];
                q += (double)part_sq[(long long)t * C + c];
            }
    }
    s = bf_channel_sum(s, s_red, ch, li);
    q = bf_channel_sum(q, s_red, ch, li);
    if (ok && li == 0) {
        const double n = (double)count;
        const double g = gamma[c], is = invstd[c];
        if (dgamma) dgamma[c] = accumulate ? dgamma[c] + (float)q : (float)q;
        if (dbeta) dbeta[c] = accumulate ? dbeta[c] + (float)s : (float)s;
        const double a = g * is;
        c1[c] = (float)a;                       // dy = c1 dz + c2 (y - mean) + c3
        c2[c] = (float)(-a * is * q / n);
        c3[c] = (float)(-a * s / n);
        (void)mean;
    }
}

// one warp per channel, lanes = clouds: scatter dpool through the saved argmax rows, apply the ReLU mask, sum for BatchNorm
// backward. The argmax -> activation -> store chain of every (cloud, channel) is independent, so the whole kernel is two
// memory round trips; the per-channel sums are reduced over the lanes in a fixed (tree) order.
__global__ void pool_scatter_bwd_kernel(const float* __restrict__ dpool, const int* __restrict__ arg, const float* __restrict__ y,
                                        const float* __restrict__ scale, const float* __restrict__ shift,
                                        const float* __restrict__ mean, const float* __restrict__ invstd, int n_clouds,
                                        int rows, int C, float* __restrict__ dz, float* __restrict__ part_sum,
                                        float* __restrict__ part_sq) {
    pdl_sync();
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= C) return;
    const float sc = scale[c], sh = shift[c], mu = mean[c], is = invstd[c];
    float s = 0.f, q = 0.f;
    for (int b = lane; b < n_clouds; b += 32) {
        const long long off = ((long long)b * rows + arg[b * C + c]) * C + c;
        const float yv = y[off];
        const float g = fmaf(yv - mu, sc, sh) > 0.f ? dpool[b * C + c] : 0.f;
        dz[off] = g;
        s += g;
        q = fmaf(g, (yv - mu) * is, q);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
        part_sum[c] = s;
        part_sq[c] = q;
    }
}

__global__ void fold_input_transform_bwd_kernel(const float* __restrict__ dW1eff, const float* __restrict__ W1,
                                                const float* __restrict__ T, int n_clouds, float* __restrict__ dW1,
                                                float* __restrict__ dT) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < 64 * 12) {
        const int c = i / 12, j = i % 12;
        float s = 0.f;
        if (j >= 3) {
            for (int b = 0; b < n_clouds; ++b) s += dW1eff[(b * 64 + c) * 9 + (j - 3)];
        } else {
            for (int b = 0; b < n_clouds; ++b)
                for (int r = 0; r < 3; ++r) s = fmaf(dW1eff[(b * 64 + c) * 9 + r], T[b * 9 + r * 3 + j], s);
        }
        dW1[i] = s;
    } else if (i < 64 * 12 + n_clouds * 9) {
        const int e = i - 64 * 12, b = e / 9, r = (e % 9) / 3, j = e % 3;
        float s = 0.f;
        for (int c = 0; c < 64; ++c) s = fmaf(dW1eff[(b * 64 + c) * 9 + r], W1[c * 12 + j], s);
        dT[e] = s;
    }
}

// stage 1: slab (128 rows) column sums; stage 2: sum of the slabs
__global__ void colsum_stage1_kernel(const float* __restrict__ dout, long long ldo, int rows, int C, int slabs,
                                     float* __restrict__ scratch) {
    pdl_sync();
    const int b = blockIdx.y, sl = blockIdx.x;
    const int r_lo = sl * 128, r_hi = min(rows, r_lo + 128);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        // 16 loads in flight, added in row order (the sum is the same as a plain loop's, bit for bit)
        float s = 0.f;
        const float* src = dout + ((long long)b * rows + r_lo) * ldo + c;
        int r = r_lo;
        for (; r + 16 <= r_hi; r += 16, src += 16 * ldo) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = __ldg(src + j * ldo);
#pragma unroll
            for (int j = 0; j < 16; ++j) s += v[j];
        }
        for (; r < r_hi; ++r, src += ldo) s += __ldg(src);
        scratch[((long long)b * slabs + sl) * C + c] = s;
    }
}
__global__ void colsum_stage2_kernel(const float* __restrict__ scratch, int slabs, int C, int total, float* __restrict__ dg) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int b = i / C, c = i % C;
    float s = 0.f;
    for (int sl = 0; sl < slabs; ++sl) s += scratch[((long long)b * slabs + sl) * C + c];
    dg[i] = s;
}

// block per token: dgl copy + dpre[t, k] = leaky'(h_pre) * sum_e dtok[t, e] w2[e, k]
__global__ void posenc_bwd_token_kernel(const float* __restrict__ dtok, const float* __restrict__ h_pre,
                                        const float* __restrict__ w2, int n_clouds, int n_tokens, int E,
                                        float* __restrict__ dgl, float* __restrict__ dpre) {
    pdl_sync();
    const int b = blockIdx.x / n_tokens, w = blockIdx.x % n_tokens;
    const long long t = (long long)b * n_tokens + w;
    for (int e = threadIdx.x; e < E; e += blockDim.x) dgl[((long long)w * n_clouds + b) * E + e] = dtok[t * E + e];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int k = warp; k < 16; k += (blockDim.x >> 5)) {
        float s = 0.f;
        for (int e = lane; e < E; e += 32) s = fmaf(dtok[t * E + e], w2[e * 16 + k], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) dpre[t * 16 + k] = h_pre[t * 16 + k] > 0.f ? s : 0.01f * s;
    }
}
// one thread per parameter-gradient element, fixed-order sum over the tokens
__global__ void posenc_bwd_param_kernel(const float* __restrict__ dtok, const float* __restrict__ cent,
                                        const float* __restrict__ h_pre, const float* __restrict__ dpre, int T, int E,
                                        float* __restrict__ dfc1_w, float* __restrict__ dfc1_b,
                                        float* __restrict__ dfc2_w, float* __restrict__ dfc2_b) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n2w = E * 16;
    if (i < n2w) {
        const int e = i / 16, k = i % 16;
        float s = 0.f;
        for (int t = 0; t < T; ++t) {
            const float hp = h_pre[t * 16 + k];
            s = fmaf(dtok[(long long)t * E + e], hp > 0.f ? hp : 0.01f * hp, s);
        }
        dfc2_w[i] = s;
    } else if (i < n2w + E) {
        const int e = i - n2w;
        float s = 0.f;
        for (int t = 0; t < T; ++t) s += dtok[(long long)t * E + e];
        dfc2_b[e] = s;
    } else if (i < n2w + E + 32) {
        const int e = i - n2w - E, k = e / 2, j = e % 2;
        float s = 0.f;
        for (int t = 0; t < T; ++t) s = fmaf(dpre[t * 16 + k], cent[t * 2 + j], s);
        dfc1_w[e] = s;
    } else if (i < n2w + E + 48) {
        const int k = i - n2w - E - 32;
        float s = 0.f;
        for (int t = 0; t < T; ++t) s += dpre[t * 16 + k];
        dfc1_b[k] = s;
    }
}

// one block per (cloud, head): dS / Pd in shared memory, then dQ, dK, dV
__global__ void attention_core_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ qkv,
                                          const float* __restrict__ probs, float drop_p, unsigned long long drop_seed_arg,
                                          const unsigned long long* drop_off, int L, int E, int heads, float* __restrict__ dqkv) {
    pdl_sync();
    const unsigned long long drop_seed = eff_seed(drop_seed_arg, drop_off);
    extern __shared__ float sm[];
    float* dS = sm;            // [L][L]
    float* Pd = sm + L * L;    // [L][L]
    const int hd = E / heads;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const float scale = rsqrtf((float)hd);
    const float* base = qkv + (long long)b * L * 3 * E;
    const float* dO = dout + (long long)b * L * E + h * hd;
    const long long pbase = ((long long)b * heads + h) * L * L;
    for (int e = threadIdx.x; e < L * L; e += blockDim.x) {
        const int i = e / L, j = e % L;
        float s = 0.f;
        for (int d = 0; d < hd; ++d) s = fmaf(dO[(long long)i * E + d], base[(long long)j * 3 * E + 2 * E + h * hd + d], s);
        const float keep = drop_p > 0.f ? dropout_keep(drop_seed, (unsigned long long)(pbase + e), drop_p) : 1.f;
        const float pr = probs[pbase + e];
        dS[e] = s * keep;      // dP
        Pd[e] = pr * keep;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        float dot = 0.f;
        for (int j = 0; j < L; ++j) dot = fmaf(dS[i * L + j], probs[pbase + i * L + j], dot);
        for (int j = 0; j < L; ++j) dS[i * L + j] = probs[pbase + i * L + j] * (dS[i * L + j] - dot);
    }
    __syncthreads();
    float* dq = dqkv + (long long)b * L * 3 * E;
    for (int e = threadIdx.x; e < L * hd; e += blockDim.x) {
        const int i = e / hd, d = e % hd;
        float sq = 0.f, sk = 0.f, sv = 0.f;
        for (int j = 0; j < L; ++j) {
            sq = fmaf(dS[i * L + j], base[(long long)j * 3 * E + E + h * hd + d], sq);          // dQ[i] += dS[i,j] K[j]
            sk = fmaf(dS[j * L + i], base[(long long)j * 3 * E + h * hd + d], sk);              // dK[i] += dS[j,i] Q[j]
            sv = fmaf(Pd[j * L + i], dO[(long long)j * E + d], sv);                             // dV[i] += Pd[j,i] dO[j]
        }
        dq[(long long)i * 3 * E + h * hd + d] = sq * scale;
        dq[(long long)i * 3 * E + E + h * hd + d] = sk * scale;
        dq[(long long)i * 3 * E + 2 * E + h * hd + d] = sv;
    }
}

}  // namespace

int bn_fold_eval(const BnDesc* table, int n_layers, float eps, cudaStream_t st) {
    if (n_layers > BnTable::kMax) return fail(AMP_E_BADARG, "bn_fold_eval: more than %d layers", BnTable::kMax);
    BnTable t;
    for (int i = 0; i < n_layers; ++i) t.d[i] = table[i];
    launch_pdl(bn_fold_eval_kernel, dim3((unsigned)(n_layers)), dim3(256), 0, st, t, eps);
    count_launch();
    return check_launch("bn_fold_eval");
}

int bn_finalize_train(const float* part_sum, const float* part_m2, int n_clouds, int rows_per_cloud, int C,
                      const float* gamma, float* running_mean, float* running_var, long long* nbt, float momentum,
                      float eps, float* scale, float* save_mean, float* save_invstd, cudaStream_t st) {
    const int tpc = (rows_per_cloud + 127) / 128;
    launch_pdl(bn_finalize_train_kernel, dim3((unsigned)((C + BF_CH - 1) / BF_CH)), dim3(BF_TPC * BF_CH), 0, st, part_sum, part_m2, n_clouds * tpc, tpc, rows_per_cloud,
                                                         (long long)n_clouds * rows_per_cloud, C, gamma, running_mean,
                                                         running_var, nbt, momentum, eps, scale, save_mean, save_invstd);
    count_launch();
    return check_launch("bn_finalize_train");
}

int bn_backward_finalize(const float* part_sum, const float* part_sq, int tiles, long long count, int C,
                         const float* gamma, const float* mean, const float* invstd, float* dgamma, float* dbeta,
                         int accumulate, float* c1, float* c2, float* c3, cudaStream_t st) {
    launch_pdl(bn_backward_finalize_kernel, dim3((unsigned)((C + BF_CH - 1) / BF_CH)), dim3(BF_TPC * BF_CH), 0, st, part_sum, part_sq, tiles, count, C, gamma, mean, invstd,
                                                            dgamma, dbeta, accumulate, c1, c2, c3);
    count_launch();
    return check_launch("bn_backward_finalize");
}

int pool_decode(const unsigned long long* pmax, const unsigned long long* pmin, int mode, const float* scale,
                const float* shift, const float* mean, int n_clouds, int C, float* pooled, int* arg, cudaStream_t st) {
    const int total = n_clouds * C;
    launch_pdl(pool_decode_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, pmax, pmin, mode, scale, shift, mean, total, C, pooled, arg);
    count_launch();
    return check_launch("pool_decode");
}

int pool_scatter_bwd(const float* dpool, const int* arg, const float* y, const float* scale, const float* shift,
                     const float* mean, const float* invstd, int n_clouds, int rows_per_cloud, int C, float* dz,
                     float* part_sum, float* part_sq, cudaStream_t st) {
    launch_pdl(pool_scatter_bwd_kernel, dim3((unsigned)((C + 7) / 8)), dim3(256), 0, st, dpool, arg, y, scale, shift, mean, invstd, n_clouds,
                                                         rows_per_cloud, C, dz, part_sum, part_sq);
    count_launch();
    return check_launch("pool_scatter_bwd");
}

int add_identity(float* t, int n_mats, int d, cudaStream_t st) {
    const int total = n_mats * d * d;
    launch_pdl(add_identity_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, t, total, d);
    count_launch();
    return check_launch("add_identity");
}

int fold_input_transform(const float* W1, const float* T, int n_clouds, float* W1eff, cudaStream_t st) {
    const int total = n_clouds * 64 * 9;
    launch_pdl(fold_input_transform_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, W1, T, n_clouds, W1eff);
    count_launch();
    return check_launch("fold_input_transform");
}

int fold_input_transform_bwd(const float* dW1eff, const float* W1, const float* T, int n_clouds, float* dW1,
                             float* dT, cudaStream_t st) {
    const int total = 64 * 12 + n_clouds * 9;
    launch_pdl(fold_input_transform_bwd_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, st, dW1eff, W1, T, n_clouds, dW1, dT);
    count_launch();
    return check_launch("fold_input_transform_bwd");
}

int broadcast_rows(const float* g, int n_clouds, int rows_per_cloud, int C, float* out, long long ldo, cudaStream_t st) {
    if ((C & 3) || (ldo & 3) || (reinterpret_cast<uintptr_t>(out) & 15) || (reinterpret_cast<uintptr_t>(g) & 15))
        return fail(AMP_E_BADARG, "broadcast_rows: needs 16-byte aligned rows");
    const long long total = (long long)n_clouds * rows_per_cloud * (C / 4);
    if (C / 4 <= 256 && 256 % (C / 4) == 0 && n_clouds <= 65535) {
        // ~8 CTAs per SM over the whole launch, at least 4 passes per thread
        const int rstep = 256 / (C / 4);
        int chunks = (int)std::min<long long>((rows_per_cloud + 4 * rstep - 1) / (4 * rstep), std::max<long long>(1, (long long)kNumSMs * 8 / n_clouds));
        const int rows_per_cta = ((rows_per_cloud + chunks - 1) / chunks + rstep - 1) / rstep * rstep;
        chunks = (rows_per_cloud + rows_per_cta - 1) / rows_per_cta;
        launch_pdl(broadcast_rows_cols_kernel, dim3((unsigned)chunks, (unsigned)n_clouds), dim3(256), 0, st, g, rows_per_cloud, C / 4,
                   reinterpret_cast<float4*>(out), ldo / 4, rows_per_cta);
        count_launch();
        return check_launch("broadcast_rows");
    }
    long long blocks = (total + 255) / 256;
    if (blocks > kNumSMs * 16) blocks = kNumSMs * 16;
    launch_pdl(broadcast_rows_kernel, dim3((unsigned)((unsigned)blocks)), dim3(256), 0, st, g, rows_per_cloud, C / 4, reinterpret_cast<float4*>(out), ldo / 4, total);
    count_launch();
    return check_launch("broadcast_rows");
}

size_t colsum_scratch_floats(int n_clouds, int rows_per_cloud, int C) {
    return (size_t)n_clouds * ((rows_per_cloud + 127) / 128) * C;
}

int colsum_rows(const float* dout, long long ldo, int n_clouds, int rows_per_cloud, int C, float* dg, float* scratch,
                cudaStream_t st) {
    const int slabs = (rows_per_cloud + 127) / 128;
    if (n_clouds > 65535) return fail(AMP_E_BADARG, "colsum_rows: more than 65535 clouds");
    launch_pdl(colsum_stage1_kernel, dim3(slabs, n_clouds), dim3(256), 0, st, dout, ldo, rows_per_cloud, C, slabs, scratch);
    count_launch();
    int rc = check_launch("colsum_stage1");
    if (rc) return rc;
    const int total = n_clouds * C;
    launch_pdl(colsum_stage2_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, scratch, slabs, C, total, dg);
    count_launch();
    return check_launch("colsum_stage2");
}

int posenc_add(const float* gl, long long gl_ld, const float* centroids, const float* fc1_w, const float* fc1_b, const float* fc2_w,
               const float* fc2_b, int n_clouds, int n_tokens, int E, float* tokens, float* h_pre, cudaStream_t st) {
    launch_pdl(posenc_add_kernel, dim3((unsigned)(n_clouds * n_tokens)), dim3(128), 0, st, gl, gl_ld, centroids, fc1_w, fc1_b, fc2_w, fc2_b, n_clouds, n_tokens, E,
                                                           tokens, h_pre);
    count_launch();
    return check_launch("posenc_add");
}

int posenc_bwd(const float* dtokens, const float* centroids, const float* h_pre, const float* fc2_w, int n_clouds,
               int n_tokens, int E, float* dgl, float* dpre_scratch, float* dfc1_w, float* dfc1_b, float* dfc2_w,
               float* dfc2_b, cudaStream_t st) {
    launch_pdl(posenc_bwd_token_kernel, dim3((unsigned)(n_clouds * n_tokens)), dim3(128), 0, st, dtokens, h_pre, fc2_w, n_clouds, n_tokens, E, dgl, dpre_scratch);
    count_launch();
    int rc = check_launch("posenc_bwd_token");
    if (rc) return rc;
    const int total = E * 16 + E + 48;
    launch_pdl(posenc_bwd_param_kernel, dim3((unsigned)((total + 127) / 128)), dim3(128), 0, st, dtokens, centroids, h_pre, dpre_scratch, n_clouds * n_tokens, E,
                                                               dfc1_w, dfc1_b, dfc2_w, dfc2_b);
    count_launch();
    return check_launch("posenc_bwd_param");
}

int attention_core(const float* qkv, const unsigned char* key_mask, float drop_p, unsigned long long drop_seed,
                   int n_clouds, int n_tokens, int E, int heads, float* out, float* probs, cudaStream_t st) {
    if (E % heads) return fail(AMP_E_BADARG, "attention_core: E %% heads != 0");
    if (n_tokens > 1024) return fail(AMP_E_BADARG, "attention_core: more than 1024 tokens per cloud");
    const int warps = n_clouds * heads * n_tokens;
    const int wpb = 4;
    launch_pdl(attention_core_kernel, dim3((unsigned)((warps + wpb - 1) / wpb)), dim3(wpb * 32), wpb * n_tokens * sizeof(float), st, 
        qkv, key_mask, drop_p, drop_seed, dropout_offset(), n_clouds, n_tokens, E, heads, out, probs);
    count_launch();
    return check_launch("attention_core");
}

int attention_core_bwd(const float* dout, const float* qkv, const float* probs, float drop_p,
                       unsigned long long drop_seed, int n_clouds, int n_tokens, int E, int heads, float* dqkv,
                       cudaStream_t st) {
    if (n_tokens > 64) return fail(AMP_E_BADARG, "attention_core_bwd: more than 64 tokens per cloud");
    launch_pdl(attention_core_bwd_kernel, dim3((unsigned)(n_clouds * heads)), dim3(128), 2 * n_tokens * n_tokens * sizeof(float), st, 
        dout, qkv, probs, drop_p, drop_seed, dropout_offset(), n_tokens, E, heads, dqkv);
    count_launch();
    return check_launch("attention_core_bwd");
}

}  // namespace amp
