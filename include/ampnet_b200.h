/* ampnet_b200 -- C ABI of the B200-native (sm_100a) AMP-Net hot path.
 *
 * This header is the drop-in boundary: plain pointers and sizes, no torch types, no C++
 * exceptions. Every entry point replaces one reference interface (cited per function, paths
 * relative to the reference repo marionacaros/3D-semantic-segmentation-AMP-Net).
 *
 * Conventions
 *   - all data pointers are DEVICE pointers owned by the caller (allocated e.g. by torch.empty);
 *     the library never allocates, frees or retains them;
 *   - `stream` is a cudaStream_t passed as void*; every call only ENQUEUES work on it and
 *     never synchronises;
 *   - return value 0 = success, negative = AMP_E_* below; amp_last_error() returns a
 *     thread-local human-readable message for the last failure on the calling thread;
 *   - workspace sizes are obtained from the matching *_workspace_bytes() call.
 */
#ifndef AMPNET_B200_H_
#define AMPNET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMP_OK            0
#define AMP_E_BADARG     -1   /* shape / size / alignment rejected */
#define AMP_E_WORKSPACE  -2   /* workspace missing or too small   */
#define AMP_E_CUDA       -3   /* launch failed (cudaGetLastError) */
#define AMP_E_ARCH       -4   /* device is not sm_100             */

/* arithmetic of the PointNet-attention forward (amp_encoder_fwd / amp_seg_fwd `precision`):
 *   AMP_PREC_FP32  the parity path (logits within 1e-3 relative of the fp32 reference; measured 6e-6), train + eval.
 *                  The shared-MLP layers run on the tcgen05 tensor cores at fp32-class accuracy: every fp32 operand is split
 *                  into two 16-bit terms and lo*hi + hi*lo + hi*hi is accumulated in fp32. Eval: bf16 terms (~2^-17 per
 *                  product), fused chains with the activations resident in tensor memory (tc_chain32). Training forward: fp16
 *                  terms (2^-23: max-pool ties must not flip), one kernel per layer with batch statistics in the epilogue;
 *                  backward: bf16 terms (gradient magnitudes need the exponent range). The few-row FC / token layers and all
 *                  BatchNorm statistics are plain fp32 on the CUDA cores.
 *   AMP_PREC_FP32_STRICT  as AMP_PREC_FP32 but no split-bf16 arithmetic: every layer of the call is plain fp32 FMA on the
 *                  CUDA cores (3-4x slower): the arithmetic of the reference itself, for gradient comparisons at the 1e-4 level.
 *   AMP_PREC_BF16  eval only: fused tcgen05 chains, bf16 operands (BatchNorm folded into the weights), fp32 accumulate;
 *                  logits within ~3e-3 .. 7e-3 of the reference. */
#define AMP_PREC_FP32     0
#define AMP_PREC_BF16     1
#define AMP_PREC_FP32_STRICT 2

const char* amp_last_error(void);
/* ABI version of this header: major*1000 + minor. */
int amp_abi_version(void);
/* Number of kernel launches the library has enqueued since process start (all threads);
 * bench.py uses the difference over its timed region for `gpu_launches`. */
int64_t amp_launch_count(void);
/* Debugging / test hooks (no reference counterpart).
 *   amp_path_count(name)        launches served so far by the kernel family `name`: "tc_layer" (split-bf16 tcgen05 layer),
 *                               "tc_layer_dgrad" (its input-gradient modes), "tc_wgrad", "tc_chain" (bf16 fused chains),
 *                               "tc_chain32" (fp32-class fused chains), "pw_linear" / "wgrad_partial" (CUDA-core tiles)
 *   amp_debug_set_disabled(csv) switches optional fast paths off by name ("tc_layer,tc_wgrad,..."; same names as the
 *                               AMP_DISABLE environment variable, which it overrides); NULL restores the environment's list.
 *                               The generic kernels then serve the call: tests compare both on the same inputs. */
int64_t amp_path_count(const char* name);
int amp_debug_set_disabled(const char* csv);
/* Dropout seed offset on the device (no reference counterpart; the reference draws its masks from torch's generator on
 * every call, pointNet/model/pointnetAtt.py:167,188). amp_seg_fwd / amp_seg_bwd take their dropout seed BY VALUE, which a
 * captured CUDA graph would replay unchanged. When `device_u64` is non-null, every dropout site of the following calls (of
 * any thread: autograd runs the backward on a worker thread) uses seed + *device_u64, read on the device at execution time: a graph-captured training step bumps that
 * word inside the graph and draws a fresh mask per replay. The word must stay valid while such work is enqueued or
 * captured; forward and backward of one step must see the same value. NULL switches the offset off. */
int amp_set_dropout_offset(const void* device_u64);

/* ------------------------------------------------------------------------------------------
 * Farthest-point sampling.  Replaces utils/utils.py:889-933 `fps(pc, n_samples)` and its
 * driver data_proc/sample_fps.py:12-34 (float32 input) for a batch of B equally sized clouds.
 *   pc          [B, P, row_stride] AoS rows; only columns 0:3 are read (utils.py:894)
 *   out_idx     [B, S] int64: indices in pick order; out_idx[b,0] == start_idx (utils.py:907)
 *   status      [B] int32 (may be NULL): 0 ok, 1 = cloud holds a non-finite coordinate
 * Semantics pinned by oracle/fps_oracle.py: picked points leave the candidate set, ties go to
 * the lowest index, distance = (dx*dx + dy*dy) + dz*dz without FMA in the input precision.
 * Requires 1 <= S <= P, 0 <= start_idx < P, P < 2^31.
 * ------------------------------------------------------------------------------------------ */
size_t amp_fps_workspace_bytes(int64_t B, int64_t P, int32_t elem_bytes);
int amp_fps_f32(const float* pc, int64_t B, int64_t P, int64_t row_stride, int32_t S,
                int32_t start_idx, int64_t* out_idx, int32_t* status,
                void* workspace, size_t workspace_bytes, void* stream);
int amp_fps_f64(const double* pc, int64_t B, int64_t P, int64_t row_stride, int32_t S,
                int32_t start_idx, int64_t* out_idx, int32_t* status,
                void* workspace, size_t workspace_bytes, void* stream);
/* Row gather pc[idx] (utils.py:933 `return pc[sample_inds]`): out [B, S, row_elems]. */
int amp_gather_rows(const void* pc, int64_t B, int64_t P, int64_t row_elems, int32_t elem_bytes,
                    const int64_t* idx, int64_t S, void* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * K-means block split.  Replaces the `KMeansConstrained(...).fit_predict(in_pc[:, i_f])` call
 * of data_proc/3_kmeans.py:78-82 and utils/utils.py:500-505 (third-party solver; restated by
 * oracle/kmeans_oracle.py, which defines the deterministic rules used here).
 * ------------------------------------------------------------------------------------------ */
/* Assignment step: labels[i] = argmin_j ((x_i - c_j)^2).sum(-1), first minimum wins.
 *   feats [n,3] f32, centroids [k,3] f32 (1 <= k <= 64), labels [n] int32, min_d2 [n] f32|NULL */
int amp_kmeans_assign_f32(const float* feats, const float* centroids, int64_t n, int32_t k,
                          int32_t* labels, float* min_d2, void* stream);
/* Column gather feats[i,:] = pc[i, cols[0..2]] (`in_pc[:, i_f]`, 3_kmeans.py:81-82). */
int amp_kmeans_gather_feats_f32(const float* pc, int64_t n, int64_t row_stride,
                                int32_t c0, int32_t c1, int32_t c2, float* feats, void* stream);
/* Whole constrained k-means for W independent windows (one CTA per window, working state in shared memory):
 *   feats      [sum n_w, 3] f32, window w owns rows [offsets[w], offsets[w+1])
 *   offsets    [W+1] int64 (device)
 *   ks         [W] int32 (device), 1 <= k_w <= kmax <= 32
 *   size_min / size_max  0 = unconstrained on that side (3_kmeans.py:78: both 2048; utils.py:500: min only)
 *   n_init     restarts (3_kmeans.py:78-80 passes n_init=5): restart r seeds the farthest-point initialisation at row
 *              (r * n_w) / n_init, the run with the smallest inertia (fixed-point sum of squared distances) wins, the
 *              earliest on a tie; n_init > 1 needs windows that fit the on-chip kernel (<= ~28 000 points)
 *   labels     [sum n_w] int32 out; centroids [W, kmax, 3] f32 out; n_iter [W] int32 out
 * A window whose constraints cannot be met (k_w > n_w, size_max * k_w < n_w, size_min * k_w > n_w, k_w outside [1, kmax]) gets
 * labels -1, zero centroids and n_iter -1; the other windows of the call are unaffected.
 * Sums are fixed point (rint(x * 2^32) in int64): coordinates must satisfy sum |x| < 2^31 per window (the reference's
 * inputs are normalised to [-1, 1]; clustering.py checks the range).
 */
size_t amp_kmeans_workspace_bytes(int64_t total_points, int64_t W, int32_t kmax);
int amp_kmeans_constrained_f32(const float* feats, const int64_t* offsets, const int32_t* ks,
                               int64_t W, int64_t total_points, int64_t max_window_points, int32_t kmax,
                               int32_t size_min, int32_t size_max, int32_t max_iter, double tol, int32_t n_init,
                               int32_t* labels, float* centroids, int32_t* n_iter,
                               void* workspace, size_t workspace_bytes, void* stream);
/* Stable regroup by label (3_kmeans.py:88-101, utils.py:508-517): order[] lists the rows of each
 * window sorted by (label, original index); counts [W, kmax] int32; xy_mean [W, kmax, 2] f32 is
 * utils.py:538-543 get_cluster_centroid of each group (mean of columns 0 and 1 of pc). */
int amp_kmeans_regroup(const int32_t* labels, const int64_t* offsets, const int32_t* ks, int64_t W,
                       int32_t kmax, const float* pc, int64_t row_stride,
                       int64_t* order, int32_t* counts, float* xy_mean, void* stream);

/* ------------------------------------------------------------------------------------------
 * Data preparation in front of the block split (float64, bit-exact to the reference's numpy arithmetic).
 * Replaces the numerical core of data_proc/1_get_windows_split.py:53-80 (LAS tile -> W x W metre windows) and
 * data_proc/2_preprocessing_filter_norm.py:40-104 (class / outlier filter, 13-column row, normalisation); LAS I/O and
 * the md5-keyed NIR join stay on the host (the NIR value is an input column).
 *   amp_minmax_f64            out4 = (min x, max x, min y, max y); x / y element i at x[i * stride]; workspace 32 bytes
 *   amp_window_ids_f64        ids[i] = iy * nx + ix of the window (x0 + ix*wx, x0 + (ix+1)*wx) x (y0 + iy*wy, ...) that holds
 *                             point i, bounds strict on both sides (:58-62), -1 when no window takes it; the host derives
 *                             x0 = round(min x), nx = len(range(x0, round(max x), wx)) (Python semantics) from amp_minmax_f64
 *   amp_window_partition      stable counting sort by window id: order[] = point indices, window w = order[offsets[w] ..
 *                             offsets[w + 1]) in original order, dropped points behind offsets[n_bins]
 *   amp_filter_normalize_f64  cols [P, 10] = (x, y, z, HeightAboveGround, class, intensity, red, green, blue, nir) float64;
 *                             per window: drop classes 2, 7, 8, 13, 24, 30 and HAG outside [0, max_z], write rows
 *                             (x', y', hag / max_z, class, clip(intensity / max_intensity), r, g, b / 65536, clip(nir / 65535),
 *                             clip((ndvi + 1) / 2), x, y, z) with x', y' = 2 (v - min) / (max - min) - 1 over the kept rows;
 *                             a window with no kept row or a zero x / y extent writes nothing (:56, :92).
 *                             out [sum kept, 13] f64 (caller sizes it for P rows), out_offsets [n_windows + 1] int64 (device).
 * ------------------------------------------------------------------------------------------ */
int amp_minmax_f64(const double* x, const double* y, int64_t n, int64_t stride, double* out4, void* workspace,
                   size_t workspace_bytes, void* stream);
int amp_window_ids_f64(const double* x, const double* y, int64_t n, int64_t stride, double x0, double y0, int32_t wx, int32_t wy,
                       int32_t nx, int32_t ny, int32_t* ids, void* stream);
size_t amp_window_partition_workspace_bytes(int64_t n);
int amp_window_partition(const int32_t* ids, int64_t n, int32_t n_bins, int64_t* order, int64_t* offsets, void* workspace,
                         size_t workspace_bytes, void* stream);
size_t amp_filter_normalize_workspace_bytes(int64_t n_windows);
int amp_filter_normalize_f64(const double* cols, const int64_t* order, const int64_t* offsets, int64_t n_windows, double max_z,
                             double max_intensity, double* out, int64_t* out_offsets, void* workspace, size_t workspace_bytes,
                             void* stream);

/* ------------------------------------------------------------------------------------------
 * Training-batch assembly and augmentation on the device.  Replaces the host work of
 * pointNet/self-attention/train_pointnet-attention.py:390-408 for one collated batch: shuffle_clusters (utils/utils.py:620-632),
 * per-window rotate_point_cloud_z (:582-604) + shuffle_data (:607-617) and the W per-window host-to-device copies.
 *   pc            [B, N, D, W] f32, the collate_seq_padd layout (collate_fns.py:4-55), D >= 3     targets [B, N, W] int64 | NULL
 *   cluster_perm  [W] int32: output window w reads input window cluster_perm[w]
 *   point_perm    [W, N] int32: output row i of window w reads input row point_perm[w][i]
 *   rotate        != 0: columns 0:3 times [[c, s, 0], [-s, c, 0], [0, 0, 1]] in float64 (numpy's arithmetic), rounded to f32
 *   x             [W, B, N, D] f32 out (one contiguous [B, N, D] encoder input per window)
 *   targets_out   [B, W * N] int64 out (window-major concatenation of train_pointnet-attention.py:421) | NULL
 * ------------------------------------------------------------------------------------------ */
int amp_assemble_windows_f32(const float* pc, const int64_t* targets, const int32_t* cluster_perm, const int32_t* point_perm, int64_t B,
                             int64_t N, int32_t D, int32_t W, int32_t rotate, double cos_a, double sin_a, float* x, int64_t* targets_out,
                             void* stream);

/* ------------------------------------------------------------------------------------------
 * PointNet encoder.  Replaces BasePointNet.forward (pointNet/model/pointnetAtt.py:80-112, with its two
 * TransformationNets :28-47) as called by train_pointnet-attention.py:410 / test_pointnet_att_segmen.py:164,
 * and the autograd backward loss.backward() runs through it (train_pointnet-attention.py:467).
 *   params      host array of amp_encoder_param_count() DEVICE pointers, one per state_dict entry of
 *               BasePointNet(point_dimension=3, global_feat_dim=256) in state_dict order
 *               (amp_encoder_param_name(i) names entry i); float32, num_batches_tracked int64
 *   x           [B, N, 9] f32        out [B, N, 320] f32 = [global 256 (repeated) | local 64] (:109-110)
 *   feat_t      [B, 64, 64] f32 feature transform (second module output, :94)
 *   training    0: eval (BatchNorm running statistics);  1: train (batch statistics; running_mean /
 *               running_var / num_batches_tracked updated in place; `saved` filled for backward)
 *   precision   AMP_PREC_FP32 | AMP_PREC_BF16 (eval only) | AMP_PREC_FP32_STRICT
 *   pack_cache  amp_encoder_pack_bytes() bytes owned by the caller, or NULL: the BatchNorm-folded, hi / lo split weights of the
 *               fused eval chains. pack_valid = 0: (re)build it in this call (first call, or parameters changed since);
 *               pack_valid = 1: reuse. Only the eval AMP_PREC_FP32 path reads it. NULL: rebuilt every call in the workspace.
 *   saved       amp_encoder_saved_bytes() bytes, kept by the caller between fwd and bwd (training only)
 *   workspace   amp_encoder_workspace_bytes() bytes of scratch (may be reused after the call's work completes)
 * amp_encoder_bwd: grads = host array of DEVICE pointers like params (NULL for the BatchNorm buffers), every
 * gradient is overwritten (not accumulated); d_out [B, N, 320]; d_feat_t [B, 64, 64] or NULL.
 * ------------------------------------------------------------------------------------------ */
int amp_encoder_param_count(void);
const char* amp_encoder_param_name(int i);
size_t amp_encoder_saved_bytes(int64_t B, int64_t N, int32_t training);
size_t amp_encoder_workspace_bytes(int64_t B, int64_t N, int32_t training);
size_t amp_encoder_pack_bytes(void);
int amp_encoder_fwd(const void* const* params, const float* x, int64_t B, int64_t N, int32_t training, int32_t precision,
                    float* out, float* feat_t, void* saved, size_t saved_bytes, void* workspace,
                    size_t workspace_bytes, void* pack_cache, size_t pack_bytes, int32_t pack_valid, void* stream);
int amp_encoder_bwd(const void* const* params, void* const* grads, const float* x, const float* out,
                    const float* feat_t, const float* d_out, const float* d_feat_t, int64_t B, int64_t N,
                    void* saved, size_t saved_bytes, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Attention + segmentation head.  Replaces SegmentationWithAttention.forward
 * (pointNet/model/pointnetAtt.py:176-209) as called by train_pointnet-attention.py:435 /
 * test_pointnet_att_segmen.py:176-177, and its autograd backward.
 *   params      host array of amp_seg_param_count() DEVICE pointers in state_dict order of
 *               SegmentationWithAttention(256, heads, num_classes, local_dim=64)
 *   gl_feats    [W, B, E] f32 (sequence first, as the script passes it)      lo_feats [B, rows, 64] f32
 *   gl_ld, lo_ld  floats between consecutive rows of gl_feats (row w * B + b, >= E) and lo_feats (row b * rows + r, >= 64,
 *               a multiple of 4, 16-byte aligned rows): the scripts pass slices of the encoder output
 *               (train_pointnet-attention.py:427-433 `out[:, :, -64:]`, `out[:, 0, :-64]`), which are read in place
 *   centroids   [B, W, 2] f32      np_cluster HOST int32 [W] points per block, sum = rows
 *   group_rows  DEVICE int32 [W]: first row of each block (exclusive prefix sum of np_cluster)
 *   key_padding_mask  DEVICE uint8 [B, W] (1 = ignore key) or NULL
 *   logits      [B, num_classes, rows] f32
 *   training    1: batch-statistics BatchNorm + dropout(dropout_p) driven by the counter-based generator
 *               seeded with `seed` (pass the same seed to amp_seg_bwd)
 *   precision   AMP_PREC_FP32 | AMP_PREC_BF16 (eval only, num_classes <= 32) | AMP_PREC_FP32_STRICT
 *   pack_cache  amp_seg_pack_bytes(num_classes) bytes or NULL, pack_valid: as for amp_encoder_fwd
 *   saved       amp_seg_saved_bytes() bytes (needed in both modes; kept for backward in training)
 * amp_seg_bwd: d_logits [B, num_classes, rows]; writes d_gl_feats [W, B, E], d_lo_feats [B, rows, 64] and every
 * parameter gradient (overwritten). Needs block sizes sharing a factor >= 64 points and W <= 64.
 * ------------------------------------------------------------------------------------------ */
size_t amp_seg_pack_bytes(int32_t num_classes);
int amp_seg_param_count(void);
const char* amp_seg_param_name(int i);
size_t amp_seg_saved_bytes(int64_t B, int64_t W, int64_t rows, int32_t embed_dim, int32_t heads);
size_t amp_seg_workspace_bytes(int64_t B, int64_t W, int64_t rows, int32_t embed_dim, int32_t training);
int amp_seg_fwd(const void* const* params, const float* gl_feats, int64_t gl_ld, const float* lo_feats, int64_t lo_ld,
                const float* centroids,
                const int32_t* np_cluster, const int32_t* group_rows, const uint8_t* key_padding_mask, int64_t B,
                int64_t W, int64_t rows, int32_t embed_dim, int32_t heads, int32_t num_classes, int32_t training,
                int32_t precision, float dropout_p, uint64_t seed, float* logits, void* saved, size_t saved_bytes,
                void* workspace, size_t workspace_bytes, void* pack_cache, size_t pack_bytes, int32_t pack_valid, void* stream);
int amp_seg_bwd(const void* const* params, void* const* grads, const float* lo_feats, int64_t lo_ld, const float* centroids,
                const int32_t* np_cluster, const int32_t* group_rows, const float* d_logits, int64_t B, int64_t W,
                int64_t rows, int32_t embed_dim, int32_t heads, int32_t num_classes, float dropout_p, uint64_t seed,
                float* d_gl_feats, float* d_lo_feats, void* saved, size_t saved_bytes, void* workspace,
                size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * One point-wise linear layer (Conv1d kernel 1 / Linear of pointnetAtt.py) on the tcgen05 tensor cores, the
 * building block of the fused bf16 chains used by amp_encoder_fwd / amp_seg_fwd with precision AMP_PREC_BF16:
 *   y[c, r, :] = act(x[c, r, :K] @ w[N, K]^T + bias)      operands rounded to bf16, fp32 accumulate
 *   x [n_clouds, rows_per_cloud, K] f32, w [N, K] f32, bias [N] f32 | NULL, y [n_clouds, rows_per_cloud, N] f32 | NULL
 *   pool_max   [n_clouds, N] uint32 | NULL: bit patterns of max over the rows of each cloud (needs relu, N % 128 == 0;
 *              zero-initialised by the caller; the layer then runs channels x points and y is not written)
 * K multiple of 16 in [16, 128], N multiple of 16 in [16, 256].
 * ------------------------------------------------------------------------------------------ */
size_t amp_tc_linear_workspace_bytes(int32_t K, int32_t N);
int amp_tc_linear_bf16(const float* x, int64_t n_clouds, int64_t rows_per_cloud, int32_t K, const float* w,
                       const float* bias, int32_t N, int32_t relu, float* y, uint32_t* pool_max, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Weight / bias gradient of one point-wise linear layer, the unit amp_encoder_bwd / amp_seg_bwd are built from
 * (autograd backward of nn.Conv1d(k=1), pointnetAtt.py; driven by train_pointnet-attention.py:467):
 *   dw[n, k] = sum over all rows of dy[r, n] * a[r, k]        db[n] = sum dy[r, n]   (db may be NULL)
 *   dy [n_clouds, rows_per_cloud, N] f32, a [n_clouds, rows_per_cloud, K] f32, dw [N, K], deterministic summation.
 * Runs on the tcgen05 tensor cores (split bf16, fp32-class accuracy) when K % 16 == 0, N % 8 == 0, K, N <= 256 and there
 * are at least 2048 rows; on the CUDA cores otherwise. */
size_t amp_wgrad_workspace_bytes(int64_t n_clouds, int64_t rows_per_cloud, int32_t N, int32_t K);
int amp_wgrad_f32(const float* dy, const float* a, int64_t n_clouds, int64_t rows_per_cloud, int32_t N, int32_t K, float* dw,
                  float* db, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Adam step over a list of tensors in one launch.  Replaces `optimizer_pointnet.step()` / `optimizer_att.step()`
 * (train_pointnet-attention.py:468-469; torch.optim.Adam built at :141-142 with lr only: betas (0.9, 0.999), eps 1e-8, no
 * weight decay, no amsgrad) when the loop uses ampnet_b200.FusedAdam.
 *   chunk_table  device array of n_chunks records {float* p; const float* g; float* m; float* v; int32 n; int32 tensor} (40 bytes):
 *                consecutive pieces of at most amp_adam_chunk_elems() elements of tensor number `tensor` each
 *   steps        device int64 per tensor: updates done so far; this one uses t = steps[tensor] + 1 (bias corrections
 *                1 - beta^t, counted per parameter like torch: a tensor without a gradient is not stepped). The caller
 *                increments the counters of the tensors it passed afterwards.
 * ------------------------------------------------------------------------------------------ */
int32_t amp_adam_chunk_elems(void);
int amp_adam_step(const void* chunk_table, int64_t n_chunks, const int64_t* steps, float lr, float beta1, float beta2, float eps,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AMPNET_B200_H_ */
