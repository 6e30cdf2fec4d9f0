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

const char* amp_last_error(void);
/* ABI version of this header: major*1000 + minor. */
int amp_abi_version(void);
/* Number of kernel launches the library has enqueued since process start (all threads);
 * bench.py uses the difference over its timed region for `gpu_launches`. */
int64_t amp_launch_count(void);

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
/* Whole constrained k-means for W independent windows (one CTA cluster per window):
 *   feats      [sum n_w, 3] f32, window w owns rows [offsets[w], offsets[w+1])
 *   offsets    [W+1] int64 (device)
 *   ks         [W] int32 (device), 1 <= k_w <= kmax <= 32
 *   size_min / size_max  0 = unconstrained on that side (3_kmeans.py:78: both 2048; utils.py:500: min only)
 *   labels     [sum n_w] int32 out; centroids [W, kmax, 3] f32 out; n_iter [W] int32 out
 */
size_t amp_kmeans_workspace_bytes(int64_t total_points, int64_t W, int32_t kmax);
int amp_kmeans_constrained_f32(const float* feats, const int64_t* offsets, const int32_t* ks,
                               int64_t W, int64_t total_points, int64_t max_window_points, int32_t kmax,
                               int32_t size_min, int32_t size_max, int32_t max_iter, double tol,
                               int32_t* labels, float* centroids, int32_t* n_iter,
                               void* workspace, size_t workspace_bytes, void* stream);
/* Stable regroup by label (3_kmeans.py:88-101, utils.py:508-517): order[] lists the rows of each
 * window sorted by (label, original index); counts [W, kmax] int32; xy_mean [W, kmax, 2] f32 is
 * utils.py:538-543 get_cluster_centroid of each group (mean of columns 0 and 1 of pc). */
int amp_kmeans_regroup(const int32_t* labels, const int64_t* offsets, const int32_t* ks, int64_t W,
                       int32_t kmax, const float* pc, int64_t row_stride,
                       int64_t* order, int32_t* counts, float* xy_mean, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AMPNET_B200_H_ */
