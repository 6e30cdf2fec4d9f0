// Eval-mode attention tail of the segmentation head (pointNet/model/pointnetAtt.py:183-200) in ONE launch:
//   tokens = global features + positional encoding of the window centroids   (fc_1 -> leaky_relu -> fc_2, :183-185)
//   qkv = in_proj(tokens); a = per-head softmax(QK^T)V over the windows of a sample; g_w = out_proj(a)   (nn.MultiheadAttention, :187-190)
//   cb = scale_2 * (W2[:, 64:] g_w + b_2) + shift_2      (the per-block bias the fused head adds instead of the repeat / cat of :192-200)
//      = scale_2 * (Wc a + bc) + shift_2 with Wc = W2[:, 64:] W_out, bc = W2[:, 64:] b_out + b_2 folded once per parameter
//        version (seg_fold_out, kept in the pack cache): out_proj is not a phase of its own
// These are five dependent launches of a few microseconds each (B x W = 32 tokens): 33 of the forward's 232 us on B200, all
// of it launch + cold-cache latency. Here 96 CTAs (one in_proj output column per warp) stay resident across the three phases,
// separated by grid barriers (one 32-bit counter per barrier in a zeroed scratch word, release / acquire at gpu scope); the
// weight rows of ALL phases are fetched before the first phase, so after a barrier a phase only waits for the few KB of
// activations the other CTAs have just written (read with ld.global.cg: L1 is not coherent across SMs).
// Arithmetic per output follows the small-row kernels (nn_small.cu): lane <-> k mod 32, fp32 FMA chain, shuffle tree.
#include <stdint.h>

#include "nn_common.cuh"

namespace amp {
namespace {

constexpr int ST_ROWS = 32, ST_THREADS = 256, ST_E = 256, ST_KI = ST_E / 32, ST_CTAS = 96, ST_MAXW = 64;

struct SegTailArgs {
    const float* gl; long long gl_ld; const float* cent;          // [W, B, gl_ld] view of the global features; [B, W, 2]
    const float *fc1w, *fc1b, *fc2w, *fc2b;                       // positional encoding: [16, 2], [16], [E, 16], [E]
    const float *inw, *inb;                                       // [3E, E], [3E]
    const float *wc, *bc, *s2, *t2;                               // folded out_proj + conv_2 global part [hid, E], [hid]; folded bn_2
    const unsigned char* key_mask;                                // [B, W] or null
    float *qkv, *attn_o, *cb;                                     // [T, 3E], [T, E], [T, hid]
    unsigned int* bar;                                            // 2 zeroed counters
    int B, W, heads, hid;
};

__device__ __forceinline__ void load_row(float (&wv)[ST_KI], const float* __restrict__ w_row, bool ok) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < ST_KI; ++i) wv[i] = ok ? __ldg(w_row + lane + 32 * i) : 0.f;
}

// dot products of the staged tile xs[32][E] with one weight row: lane r returns row r's sum
__device__ __forceinline__ float tile_dot(const float* xs, const float (&wv)[ST_KI]) {
    const int lane = threadIdx.x & 31;
    float acc[ST_ROWS];
#pragma unroll
    for (int r = 0; r < ST_ROWS; ++r) acc[r] = 0.f;
#pragma unroll
    for (int i = 0; i < ST_KI; ++i)
#pragma unroll
        for (int r = 0; r < ST_ROWS; ++r) acc[r] = fmaf(xs[r * ST_E + lane + 32 * i], wv[i], acc[r]);
#pragma unroll
    for (int o = 16, n2 = 16; o >= 1; o >>= 1, n2 >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < n2; ++i) {
            const float send = up ? acc[i] : acc[i + n2];
            const float keep = up ? acc[i + n2] : acc[i];
            acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return acc[0];
}

__device__ __forceinline__ void stage_rows(float* xs, const float* X, int r0, int T) {
    __syncthreads();                                               // previous readers of xs are done
    for (int k = threadIdx.x; k < ST_E; k += ST_THREADS)
#pragma unroll
        for (int r = 0; r < ST_ROWS; ++r) xs[r * ST_E + k] = (r0 + r < T) ? __ldcg(X + (long long)(r0 + r) * ST_E + k) : 0.f;
    __syncthreads();
}

__global__ void __launch_bounds__(ST_THREADS) seg_tail_eval_kernel(const SegTailArgs a) {
    pdl_trigger();
    extern __shared__ float xs[];                                  // [32][E] tile, then [8 warps][ST_MAXW] attention scores
    float* sc_all = xs + ST_ROWS * ST_E;
    __shared__ float s_h[ST_ROWS][16];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gw = blockIdx.x * (ST_THREADS / 32) + warp, n_warps = gridDim.x * (ST_THREADS / 32);
    const int T = a.B * a.W, E = ST_E;
    // the weight rows of every phase (parameters: never written by a kernel of this stream) before the predecessor is waited for
    float w_in[ST_KI], w_cb[ST_KI];
    load_row(w_in, a.inw + (long long)gw * E, gw < 3 * E);
    load_row(w_cb, a.wc + (long long)gw * E, gw < a.hid);
    pdl_wait();

    // ---- phase 1: tokens (every CTA builds the tile itself: 32 x 256 x 16 FMAs) -> qkv column gw ----
    for (int r0 = 0; r0 < T; r0 += ST_ROWS) {
        __syncthreads();
        for (int i = tid; i < ST_ROWS * 16; i += ST_THREADS) {
            const int r = i >> 4, j = i & 15, t = r0 + r;
            float v = 0.f;
            if (t < T) {
                v = __ldg(a.fc1w + j * 2) * __ldg(a.cent + t * 2) + __ldg(a.fc1w + j * 2 + 1) * __ldg(a.cent + t * 2 + 1) + __ldg(a.fc1b + j);
                v = v > 0.f ? v : 0.01f * v;                       // F.leaky_relu_ default slope
            }
            s_h[r][j] = v;
        }
        __syncthreads();
        for (int e = tid; e < E; e += ST_THREADS) {
            float g[ST_ROWS];                                      // the 32 global-feature loads of this column in flight together
            {
                int b = r0 / a.W, w = r0 - b * a.W;
#pragma unroll
                for (int r = 0; r < ST_ROWS; ++r) {
                    g[r] = (r0 + r < T) ? __ldcg(a.gl + ((long long)w * a.B + b) * a.gl_ld + e) : 0.f;
                    if (++w == a.W) { w = 0; ++b; }
                }
            }
            float w2[16];
            const float4* w2v = reinterpret_cast<const float4*>(a.fc2w + e * 16);
#pragma unroll
            for (int k = 0; k < 4; ++k) { const float4 q = __ldg(w2v + k); w2[4 * k] = q.x; w2[4 * k + 1] = q.y; w2[4 * k + 2] = q.z; w2[4 * k + 3] = q.w; }
            const float b2 = __ldg(a.fc2b + e);
#pragma unroll
            for (int r = 0; r < ST_ROWS; ++r) {
                float v = b2;
#pragma unroll
                for (int k = 0; k < 16; ++k) v = fmaf(w2[k], s_h[r][k], v);
                xs[r * E + e] = (r0 + r < T) ? g[r] + v : 0.f;
            }
        }
        __syncthreads();
        if (gw < 3 * E) {
            const float v = tile_dot(xs, w_in) + __ldg(a.inb + gw);
            if (r0 + lane < T) a.qkv[(long long)(r0 + lane) * 3 * E + gw] = v;
        }
    }
    grid_barrier(a.bar + 0, gridDim.x);

    // ---- phase 2: one warp per (sample, head, query window); lanes own head-dim elements ----
    {
        const int hd = E / a.heads, L = a.W;
        const float scale = rsqrtf((float)hd);
        float* sc = sc_all + warp * ST_MAXW;
        for (int task = gw; task < a.B * a.heads * L; task += n_warps) {
            const int i = task % L, h = (task / L) % a.heads, b = task / (L * a.heads);
            const float* base = a.qkv + (long long)b * L * 3 * E;
            // scores: lanes own head-dim elements (coalesced L2 reads of q and of every key), one shuffle tree per key
            float m = -INFINITY;
            {
                const float* q = base + (long long)i * 3 * E + h * hd;
                for (int j = 0; j < L; ++j) {
                    const float* k = base + (long long)j * 3 * E + E + h * hd;
                    float s = 0.f;
                    for (int d = lane; d < hd; d += 32) s = fmaf(__ldcg(q + d) * scale, __ldcg(k + d), s);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                    if (a.key_mask && a.key_mask[b * L + j]) s = -INFINITY;
                    if (lane == 0) sc[j] = s;
                    m = fmaxf(m, s);
                }
                __syncwarp();
            }
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
            for (int d = lane; d < hd; d += 32) {
                float o = 0.f;
                for (int j = 0; j < L; ++j) o = fmaf(sc[j] * inv, __ldcg(base + (long long)j * 3 * E + 2 * E + h * hd + d), o);
                a.attn_o[((long long)b * L + i) * E + h * hd + d] = o;
            }
            __syncwarp();
        }
    }
    grid_barrier(a.bar + 1, gridDim.x);

    // ---- phase 3: per-block bias of the fused head from the attention output (out_proj folded into Wc) ----
    if (blockIdx.x * (ST_THREADS / 32) < a.hid) {
        for (int r0 = 0; r0 < T; r0 += ST_ROWS) {
            stage_rows(xs, a.attn_o, r0, T);
            if (gw < a.hid) {
                float v = tile_dot(xs, w_cb) + __ldcg(a.bc + gw);
                v = fmaf(v, __ldg(a.s2 + gw), __ldg(a.t2 + gw));
                if (r0 + lane < T) a.cb[(long long)(r0 + lane) * a.hid + gw] = v;
            }
        }
    }
}

// Wc[n][k] = sum_j W2g[n][j] W_out[j][k], bc[n] = sum_j W2g[n][j] b_out[j] + b_2[n]  (W2g = conv_2's columns of the global
// feature): one CTA per output channel n, thread = k; runs when the pack cache is (re)built
__global__ void __launch_bounds__(ST_E) seg_fold_out_kernel(const float* __restrict__ c2w, long long c2_ld, const float* __restrict__ c2b,
                                                            const float* __restrict__ outw, const float* __restrict__ outb, float* __restrict__ wc,
                                                            float* __restrict__ bc) {
    pdl_sync();
    const int n = blockIdx.x, k = threadIdx.x;
    const float* wrow = c2w + (long long)n * c2_ld;
    float acc = 0.f;
    for (int j = 0; j < ST_E; ++j) acc = fmaf(__ldg(wrow + j), __ldg(outw + (long long)j * ST_E + k), acc);
    wc[(long long)n * ST_E + k] = acc;
    if (k < 32) {
        float b = 0.f;
        for (int j = k; j < ST_E; j += 32) b = fmaf(__ldg(wrow + j), __ldg(outb + j), b);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) b += __shfl_xor_sync(0xffffffffu, b, o);
        if (k == 0) bc[n] = b + __ldg(c2b + n);
    }
}

}  // namespace

bool seg_tail_eligible(int W, int E, int heads, int hid) {
    return !path_disabled("seg_tail") && E == ST_E && W >= 1 && W <= ST_MAXW && heads >= 1 && E % heads == 0 && hid <= ST_CTAS * 8;
}

int seg_fold_out(const float* c2w, long long c2_ld, const float* c2b, const float* outw, const float* outb, int E, int hid, float* wc,
                 float* bc, cudaStream_t st) {
    if (E != ST_E) return fail(AMP_E_BADARG, "seg_fold_out: embed_dim %d", E);
    launch_pdl(seg_fold_out_kernel, dim3((unsigned)hid), dim3(ST_E), 0, st, c2w, c2_ld, c2b, outw, outb, wc, bc);
    count_launch();
    return check_launch("seg_fold_out");
}

// 1 = launched, 0 = shape not eligible (the caller runs the five separate launches)
int seg_tail_eval(const float* gl, long long gl_ld, const float* cent, const float* fc1w, const float* fc1b, const float* fc2w,
                  const float* fc2b, const float* inw, const float* inb, const float* wc, const float* bc, const float* s2,
                  const float* t2, const unsigned char* key_mask, int B, int W, int E, int heads, int hid, float* qkv, float* attn_o,
                  float* cb, unsigned int* bar, cudaStream_t st) {
    if (!seg_tail_eligible(W, E, heads, hid) || (long long)B * W > (1 << 20)) return 0;
    const size_t smem = sizeof(float) * (ST_ROWS * ST_E + (ST_THREADS / 32) * ST_MAXW);
    // the grid barriers need the whole grid resident at once: checked against the device, not assumed (a MIG slice or a smaller part
    // falls back to the separate launches)
    static int capacity = -1;
    if (capacity < 0) {
        int dev = 0, sms = 0, per_sm = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seg_tail_eval_kernel, ST_THREADS, smem) != cudaSuccess)
            capacity = 0;
        else
            capacity = sms * per_sm;
    }
    if (capacity < ST_CTAS) return 0;
    SegTailArgs a{gl, gl_ld, cent, fc1w, fc1b, fc2w, fc2b, inw, inb, wc, bc, s2, t2, key_mask, qkv, attn_o, cb, bar, B, W, heads, hid};
    cudaError_t e = cudaMemsetAsync(bar, 0, 4 * sizeof(unsigned int), st);
    if (e != cudaSuccess) return fail(AMP_E_CUDA, "seg_tail_eval: cudaMemsetAsync: %s", cudaGetErrorString(e));
    launch_pdl(seg_tail_eval_kernel, dim3(ST_CTAS), dim3(ST_THREADS), smem, st, a);
    count_launch();
    count_path("seg_tail");
    const int rc = check_launch("seg_tail_eval");
    return rc == AMP_OK ? 1 : rc;
}

}  // namespace amp
