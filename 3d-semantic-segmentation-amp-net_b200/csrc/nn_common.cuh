// Shared declarations of the PointNet-attention kernels (sm_100a).
//
// Replaces the ATen op chain of pointNet/model/pointnetAtt.py:28-47 (TransformationNet.forward),
// :80-112 (BasePointNet.forward) and :176-209 (SegmentationWithAttention.forward) of the reference,
// and the autograd backward driven by pointNet/self-attention/train_pointnet-attention.py:467.
#pragma once
#include "amp_common.cuh"

namespace amp {

// ---------------------------------------------------------------------------------------------
// Point-wise linear layer  Y[c, r, :] = epilogue( prologue(X[c, r, :K]) @ W^T + bias )
// (a Conv1d(kernel 1) / Linear of the reference over the rows of a [clouds, rows, channels] tensor;
// the same kernel computes the input gradients dX = dY @ W of the backward pass)
// ---------------------------------------------------------------------------------------------
struct PwParams {
    // input rows: X[(cloud * rows_per_cloud + r) * ldx + k], k < K
    // (x_transposed: X[(cloud * K + k) * rows_per_cloud + r], the [B, C, N] layout of the logit gradients)
    const float* X; long long ldx; int K; int x_transposed;
    // prologue per input channel (in_m optional, 0 when null):
    //   without X2:  v = (X - in_m[k]) * in_a[k] + in_b[k]                      forward, training mode: BatchNorm
    //                (+ ReLU if in_relu, * dropout keep-scale if in_drop_p > 0)  of the previous layer on its raw output
    //   with X2:     v = X * in_a[k] + in_b[k] + in_c[k] * (X2 - in_m[k])       backward: BatchNorm backward
    //                dy = c1 * dz + c3 + c2 * (y - mean)  with X = dz, X2 = y
    const float* in_a; const float* in_b; const float* in_c; const float* in_m; const float* X2; int in_relu;
    float in_drop_p; unsigned long long in_drop_seed;
    // optional device word added to every dropout seed of the call when non-null (amp_set_dropout_offset: a training step
    // captured in a CUDA graph draws a new mask per replay by bumping that word on the device)
    const unsigned long long* drop_off;
    // weights: w_kn == 0: W[n * ldw + k] (PyTorch [out, in]);  w_kn == 1: W[k * ldw + n] ([in, out]);
    // per-cloud weights when w_cloud_stride != 0
    const float* W; long long ldw; long long w_cloud_stride; int w_kn;
    // optional bias[(cloud * n_groups + g) * bias_group_stride + n]; g = group of row r by group_rows
    // (group_rows[g] = first row of group g; the per-cluster bias of the segmentation head)
    const float* bias; long long bias_group_stride; const int* group_rows; int n_groups;
    int groups_tile_aligned;            // host hint: every group starts on a multiple of 128 rows (tensor-core path)
    // tensor-core path: split the operands into fp16 (not bf16) hi + lo terms: fp32-class products (2^-23) for operands of
    // moderate magnitude -- set by the training FORWARD (normalised activations, weights), never for gradients
    int fp16_split;
    // accumulate: acc += Y (previous content) before the rest of the epilogue (gradient fan-in)
    int accumulate;
    // forward epilogue per output channel: y = y * out_scale[n] + out_shift[n]; then ReLU if out_relu
    const float* out_scale; const float* out_shift; int out_relu;
    // backward epilogue: the value is the gradient w.r.t. dropout(relu(bn(y))) of THIS tensor position:
    //   * dropout keep-scale (out_drop_p > 0), zeroed where (mask_y - mask_mean) * mask_scale + mask_shift <= 0;
    //   part_sum += dz, part_sq += dz * (mask_y - mask_mean) * mask_invstd   (BatchNorm backward sums)
    const float* mask_y; long long ld_mask; const float* mask_scale; const float* mask_shift;
    const float* mask_mean; const float* mask_invstd;
    float out_drop_p; unsigned long long out_drop_seed;
    float* Y; long long ldy;            // optional store
    int y_transposed;                   // store Y[(cloud * Nout + n) * rows_per_cloud + r]  ([B, C, N] logits)
    int n_clouds; int rows_per_cloud; int Nout;
    // pooling over the rows of a cloud: 0 none; 1 max of the final value; 2 max and min of the raw value
    int pool_mode; unsigned long long* pool_max; unsigned long long* pool_min;   // packed [clouds, Nout]
    // optional per-tile partials [tiles, Nout]: forward = sum of the raw value and its sum of squared deviations
    // from the TILE mean (combined exactly by bn_finalize_train: no E[x^2] - mean^2 cancellation);
    // backward (mask_y given) = sum of dz and of dz * xhat
    float* part_sum; float* part_sq;
    // optional: the tensor-core layer kernel also writes the split 16-bit operand it builds from X (after the prologue) to
    // global memory, per 128-row tile: [tile][hi, lo][K / 8][128 rows] 16-byte pieces of 8 consecutive channels. The weight
    // gradient of the same layer reads it back instead of redoing the prologue and the split (WgParams::dy_split).
    void* split_dump;
    // optional scratch for the split-K path of the few-row kernels (long reductions, e.g. the 4096-wide gradient of the
    // 64 x 64 transform's fc_3): (K / 32) * Nout * 32 floats
    float* splitk_ws; size_t splitk_floats;
};

int pw_linear(const PwParams& p, cudaStream_t st);
// eval-mode T-Net FC stack (fc_1 + bn_4 + relu, fc_2 + bn_5 + relu [, fc_3 + bias + identity]) in one cluster launch (nn_small.cu)
int tnet_fc_eval(const float* pooled, int B, const float* fc1, const float* s4, const float* t4, const float* fc2, const float* s5,
                 const float* t5, const float* fc3w, const float* fc3b, int d, int fc3_inside, float* h1, float* h2, float* out,
                 cudaStream_t st);
// fc_3 of the feature transform (+ bias + identity) -> out [B, 64, 64] and the packed split-bf16 per-cloud operand (nn_small.cu)
int tnet_fc3_pack(const float* h2, const float* w, const float* b, int B, float* out, unsigned char* pk, long long pk_stride, cudaStream_t st);
// the whole eval-mode FC stack of a T-Net (fc_1, fc_2, fc_3 + identity [+ the packed operand]) in ONE launch of 128 resident CTAs
// with grid barriers between the layers (nn_small.cu): 1 = launched, 0 = not eligible; `bars` = 2 zeroed words
int tnet_fc_grid(const float* pooled, int B, const float* fc1, const float* s4, const float* t4, const float* fc2, const float* s5,
                 const float* t5, const float* fc3w, const float* fc3b, int d, float* h1, float* h2, float* out, unsigned char* pk,
                 long long pk_stride, unsigned int* bars, cudaStream_t st);

#ifdef __CUDACC__
// Grid-wide barrier for kernels whose whole grid is resident (the launcher checks occupancy x SM count): every CTA has passed
// this point and its global writes are visible (read them with ld.global.cg). One zeroed 32-bit counter per barrier. CTAs that
// arrive first spin; a wait beyond ~10 s traps instead of hanging the device.
__device__ __forceinline__ void grid_barrier(unsigned int* ctr, unsigned int n_ctas) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned int v;
        for (int spin = 0;; ++spin) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= n_ctas) break;
            if (spin > (1 << 24)) __trap();
        }
        __threadfence();
    }
    __syncthreads();
}
#endif
// forward of the narrow-input (K <= 12) 64-channel layers over many rows, exact fp32 (nn_small.cu)
int narrow_fwd_try(const PwParams& p, cudaStream_t st);
// forward of the narrow-output (class logits, K = 64) layer over many rows, exact fp32 (nn_small.cu)
int narrow_out_fwd_try(const PwParams& p, cudaStream_t st);
int small_linear_try(const PwParams& p, cudaStream_t st);   // nn_small.cu: few-row layers; 1 launched, 0 not eligible, < 0 error
int tc_layer_try(const PwParams& p, cudaStream_t st);   // nn_tc_layer.cu: 1 = launched, 0 = not eligible, < 0 = error
int pw_tiles(int n_clouds, int rows_per_cloud);      // number of row tiles (= rows of part_sum)

// ---------------------------------------------------------------------------------------------
// Weight gradient  dW[c?][n, k] = sum_r proY(dY)[c, r, n] * proA(A)[c, r, k]   (+ db[n] = sum_r proY(dY)[c, r, n])
// Deterministic: every (cloud, row slab) writes a partial, wgrad reduces the partials in fixed order.
// ---------------------------------------------------------------------------------------------
struct WgParams {
    // dY operand with the BatchNorm-backward prologue (see PwParams): v = dY * y_a[n] + y_b[n] + y_c[n] * (Y2 - y_m[n])
    const float* dY; long long lddy; int Nout; const float* y_a; const float* y_b; const float* y_c; const float* y_m; const float* Y2;
    int dy_transposed;                  // dY[(cloud * Nout + n) * rows_per_cloud + r]   ([B, C, N] logit gradients)
    // A operand with the forward prologue: a = (A - a_m[k]) * a_a[k] + a_b[k]; ReLU if a_relu; dropout if a_drop_p > 0
    const float* A; long long lda; int K; const float* a_a; const float* a_b; const float* a_m; int a_relu;
    float a_drop_p; unsigned long long a_drop_seed;
    const unsigned long long* drop_off;          // see PwParams
    int n_clouds; int rows_per_cloud;
    int per_cloud;                      // 1: one dW per cloud (bmm weights), dW[cloud][n * ldw + k]
    int w_kn;                           // 0: dW[n * ldw + k]; 1: dW[k * ldw + n]
    float* dW; long long ldw; long long w_cloud_stride;
    float* db;                          // optional [Nout] (summed over clouds; with per_cloud: [clouds, Nout])
    int accumulate;                     // dW += / db += instead of =
    // optional row groups (segmentation head): dbg[(cloud * n_groups + g) * Nout + n] = sum over the rows of group g
    const int* group_rows; int n_groups; float* dbg;
    int slab_rows;                      // rows per partial (0 = 512); must divide every group size when dbg is used
    float* partials; size_t partial_floats;   // workspace from wgrad_workspace_floats()
    // optional: dy' already split into bf16 hi + lo by the input-gradient kernel of the same layer (PwParams::split_dump
    // layout, tiles of 128 rows per cloud); dY / Y2 / y_* are then not read by the tensor-core kernel
    const void* dy_split;
};
// did the last tensor-core layer launch of this thread write its PwParams::split_dump? (reset by the caller)
bool& tc_layer_dumped();
size_t wgrad_workspace_floats(int n_clouds, int rows_per_cloud, int Nout, int K, int slab_rows = 0);
// Deferred reduction of weight-gradient partials. While a scope is alive on the calling thread, wgrad() runs only the partial
// pass of a PARAMETER gradient (per_cloud == 0, accumulate == 0, no dbg: nothing else in the backward reads the result) into its own
// slice of `pool` and queues the fixed-order reduction; flush() reduces all queued jobs in ONE launch (a backward pass has
// ~13 of them, each a few microseconds of pure launch latency). Jobs that do not fit are reduced at once as before.
struct WgDeferScope {
    WgDeferScope(float* pool, size_t pool_floats, cudaStream_t st);
    ~WgDeferScope();                    // drops the scope; queued jobs must have been flushed (flush() returns the error code)
    int flush();
    WgDeferScope(const WgDeferScope&) = delete;
    WgDeferScope& operator=(const WgDeferScope&) = delete;
    float* pool; size_t pool_floats, used; cudaStream_t st; WgDeferScope* prev;
};
int wgrad_group_slab(const int* group_sizes, int n_groups);
int wgrad(const WgParams& p, cudaStream_t st);
int small_wgrad_try(const WgParams& p, cudaStream_t st);   // nn_small.cu
int narrow_wgrad_try(const WgParams& p, int slabs, int SLAB, cudaStream_t st);   // nn_small.cu
int narrow_out_wgrad_try(const WgParams& p, int slabs, int SLAB, cudaStream_t st);   // nn_small.cu
int tc_wgrad_try(const WgParams& p, int slabs, int SLAB, cudaStream_t st);   // nn_tc_wgrad.cu: 1 launched, 0 not eligible, < 0 error

// packed (value, row) keys for the pooling atomics: larger key = larger value, then lower row
__host__ __device__ inline unsigned int ordered_bits(float v) {
#ifdef __CUDA_ARCH__
    unsigned int u = __float_as_uint(v);
#else
    union { float f; unsigned int u; } c; c.f = v; unsigned int u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ inline float unordered_bits(unsigned int u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; unsigned int u; } c; c.u = u; return c.f;
#endif
}

// seed of a dropout site: the by-value seed of the call plus the optional device-side offset (see PwParams::drop_off)
__device__ __forceinline__ unsigned long long eff_seed(unsigned long long seed, const unsigned long long* off) {
    return off ? seed + __ldg(off) : seed;
}
// Counter-based dropout: keep-scale (0 or 1/(1-p)) of element `idx` of a tensor under `seed`.
// Forward and backward call it with the same (seed, idx), so no mask is stored.
__device__ __forceinline__ float dropout_keep(unsigned long long seed, unsigned long long idx, float p) {
    unsigned int h = (unsigned int)idx ^ (unsigned int)seed;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    h += (unsigned int)(idx >> 32) * 0x9e3779b1u + (unsigned int)(seed >> 32);
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    const float u = (float)(h >> 8) * (1.0f / 16777216.0f);
    return u >= p ? 1.0f / (1.0f - p) : 0.0f;
}

// ---------------------------------------------------------------------------------------------
// small kernels (nn_glue.cu)
// ---------------------------------------------------------------------------------------------
// eval BatchNorm fold: scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale
struct BnDesc { const float* gamma; const float* beta; const float* mean; const float* var; float* scale; float* shift; int C; };
struct BnTable { static constexpr int kMax = 16; BnDesc d[kMax]; };     // passed to the kernel by value (no H2D copy)
int bn_fold_eval(const BnDesc* host_table, int n_layers, float eps, cudaStream_t st);
// training BatchNorm: batch statistics from the per-tile partials of pw_linear (tile t of a cloud holds
// min(128, rows_per_cloud - 128 t) rows); writes scale = gamma * invstd for the consumer's prologue
// (applied as (y - mean) * scale + beta), saves mean / inv-std, updates the running statistics in place
int bn_finalize_train(const float* part_sum, const float* part_m2, int n_clouds, int rows_per_cloud, int C,
                      const float* gamma, float* running_mean, float* running_var,
                      long long* num_batches_tracked, float momentum, float eps, float* scale,
                      float* save_mean, float* save_invstd, cudaStream_t st);
// BatchNorm backward coefficients from the partial sums of dz and dz * xhat:
//   dbeta = sum dz, dgamma = sum dz * xhat, and  dy = c1 * dz + c2 * (y - mean) + c3  per channel
int bn_backward_finalize(const float* part_sum, const float* part_sq, int tiles, long long count, int C,
                         const float* gamma, const float* mean, const float* invstd, float* dgamma, float* dbeta,
                         int accumulate, float* c1, float* c2, float* c3, cudaStream_t st);
// pooled[b, c] and arg[b, c] from the packed keys; mode 1: key holds the final value; mode 2: raw max/min,
// pooled = relu(scale * ((scale >= 0 ? max : min) - mean) + shift)   (mean may be null)
int pool_decode(const unsigned long long* pmax, const unsigned long long* pmin, int mode, const float* scale,
                const float* shift, const float* mean, int n_clouds, int C, float* pooled, int* arg, cudaStream_t st);
// backward of max-pool + ReLU + BatchNorm statistics for the pooled layer: dz is zero except at the arg rows;
// writes dz dense [clouds * rows, C] (zero-filled by the caller) and the partial sums as ONE tile row
int pool_scatter_bwd(const float* dpool, const int* arg, const float* y, const float* scale, const float* shift,
                     const float* mean, const float* invstd, int n_clouds, int rows_per_cloud, int C, float* dz,
                     float* part_sum, float* part_sq, cudaStream_t st);
// t[i] += 1 on the diagonal of each d x d block   (fc_3 output + identity, pointnetAtt.py:42-46)
int add_identity(float* t, int n_mats, int d, cudaStream_t st);
// W1eff[b][c][i] = W1[c][3 + i] + (i < 3 ? sum_j W1[c][j] * T[b][i][j] : 0)   (bmm + cat + conv_1 of :85-90 folded)
int fold_input_transform(const float* W1, const float* T, int n_clouds, float* W1eff, cudaStream_t st);
// backward of the fold: dW1[c][3+i] (+)= sum_b dW1eff[b][c][i]; dW1[c][j] (+)= sum_b sum_i dW1eff[b][c][i] T[b][i][j];
// dT[b][i][j] = sum_c dW1eff[b][c][i] W1[c][j]
int fold_input_transform_bwd(const float* dW1eff, const float* W1, const float* T, int n_clouds, float* dW1,
                             float* dT, cudaStream_t st);
// out[b, r, 0:C] = g[b, 0:C] for every row r (the repeat + cat of :109-110)
int broadcast_rows(const float* g, int n_clouds, int rows_per_cloud, int C, float* out, long long ldo, cudaStream_t st);
// eval-mode attention tail of the segmentation head in one launch (nn_seg_tail.cu): 1 = launched, 0 = not eligible, < 0 = error;
// wc / bc = out_proj folded into conv_2's global-feature columns (seg_fold_out, once per parameter version)
bool seg_tail_eligible(int W, int E, int heads, int hid);
int seg_fold_out(const float* c2w, long long c2_ld, const float* c2b, const float* outw, const float* outb, int E, int hid, float* wc,
                 float* bc, cudaStream_t st);
int seg_tail_eval(const float* gl, long long gl_ld, const float* cent, const float* fc1w, const float* fc1b, const float* fc2w,
                  const float* fc2b, const float* inw, const float* inb, const float* wc, const float* bc, const float* s2,
                  const float* t2, const unsigned char* key_mask, int B, int W, int E, int heads, int hid, float* qkv, float* attn_o,
                  float* cb, unsigned int* bar, cudaStream_t st);
// dg[b, c] = sum_r dout[b, r, c]  (backward of the broadcast), deterministic two-stage column sum
int colsum_rows(const float* dout, long long ldo, int n_clouds, int rows_per_cloud, int C, float* dg, float* scratch,
                cudaStream_t st);
size_t colsum_scratch_floats(int n_clouds, int rows_per_cloud, int C);
// tokens[b, w, :] = gl[w, b, :] + fc2(leaky_relu(fc1(centroids[b, w])))      (:183-185); h_pre [B*W, 16] optional
int posenc_add(const float* gl, long long gl_ld, const float* centroids, const float* fc1_w, const float* fc1_b, const float* fc2_w,
               const float* fc2_b, int n_clouds, int n_tokens, int E, float* tokens, float* h_pre, cudaStream_t st);
// backward of posenc_add: dgl[w, b, :] = dtokens[b, w, :]; gradients of fc1 / fc2
int posenc_bwd(const float* dtokens, const float* centroids, const float* h_pre, const float* fc2_w, int n_clouds,
               int n_tokens, int E, float* dgl, float* dpre_scratch /*[B*W,16]*/, float* dfc1_w, float* dfc1_b, float* dfc2_w,
               float* dfc2_b, cudaStream_t st);
// softmax(q k^T / sqrt(hd)) v per (cloud, head); qkv [clouds * tokens, 3E]; key_mask [clouds, tokens] (uint8, 1 = ignore)
// probs [clouds, heads, L, L] (softmax output before dropout) optional; dropout on the probabilities when drop_p > 0
int attention_core(const float* qkv, const unsigned char* key_mask, float drop_p, unsigned long long drop_seed,
                   int n_clouds, int n_tokens, int E, int heads, float* out, float* probs, cudaStream_t st);
int attention_core_bwd(const float* dout, const float* qkv, const float* probs, float drop_p,
                       unsigned long long drop_seed, int n_clouds, int n_tokens, int E, int heads, float* dqkv,
                       cudaStream_t st);

}  // namespace amp
