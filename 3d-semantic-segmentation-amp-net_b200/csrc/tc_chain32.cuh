// Fused shared-MLP chains at fp32-class accuracy on the 5th-generation tensor cores: the AMP_PREC_FP32 eval forward.
//
// Same role as tc_chain.cuh (one kernel runs a whole chain of Conv1d(k=1) layers of pointNet/model/pointnetAtt.py over
// 128-point tiles: :31-35 T-Net convs + max-pool, :90-104 encoder convs + max-pool, :203-207 segmentation head), but
//   * every operand is a pair of bf16 terms, x = hi + lo, and each 16-wide K step issues three tcgen05.mma
//     (A_hi W_hi + A_lo W_hi + A_hi W_lo, fp32 accumulate): ~2^-17 relative per product, logits within ~1e-5 of the
//     fp32 reference instead of the ~5e-3 of plain bf16;
//   * the activations of a tile never touch shared memory: the epilogue writes the next layer's A operand straight into
//     TENSOR MEMORY (tcgen05.st, packed bf16 pairs) and the next tcgen05.mma reads A from there ("TS" form), so the only
//     shared-memory traffic of a layer is its weights;
//   * max-pooled layers are NOT transposed: the max over the 128 rows of a tile is a warp-wide redux on the float bit
//     patterns of the accumulator columns;
//   * weights that do not fit next to the others (the 256 x 128 pooled layer of the last encoder chain, 128 KB as hi + lo)
//     are streamed from L2 through a two-stage ring by a producer warp, one pass per PAIR of tiles.
// Eval mode only: BatchNorm folded into the packed weights (scale) and a per-channel bias (shift).
#pragma once
#include "nn_common.cuh"

namespace amp {

constexpr int kT32MaxOps = 6;
constexpr int kT32ChunkChannels = 64;                    // streamed weights: channels per ring stage
constexpr int kT32ChunkBytes = kT32ChunkChannels * 128 * 2 * 2;   // K = 128, hi + lo = 32 KB

struct T32Op {
    int K, N;           // per tile: D[128, N] = A[128, K] W[N, K]^T;  K % 16 == 0 in [16, 128]; N % 16 == 0, <= 128 (pooled: 128 or 256)
    int w_off;          // byte offset of the packed weights (hi block [K/8][N][8] bf16, then the lo block) in the resident blob,
    int w_cloud;        // or in the per-cloud block when w_cloud != 0
    int w_stream;       // pooled op with N = 256, K = 128 whose weights come through the ring (T32Params.wstream)
    int bias_off;       // >= 0: float offset of the op's bias in T32Params.bias (added in the epilogue); < 0: none
    int relu;
    int bias_grouped;   // + gbias[(cloud * n_groups + group(row)) * N + n] (per-block bias of the segmentation head)
    int write_act;      // result (split into hi + lo) becomes the A operand of the next op
    int store_f32;      // result rows -> out_f32[row * out_ld + out_col0 + n]  (N == 64)
    int pool;           // atomicMax(pool[cloud * N + n], max over the rows of the tile of relu(. + bias))
    int store_logits;   // first n_classes columns -> logits[(cloud * n_classes + n) * rows_per_cloud + r]
};

struct T32Params {
    int n_ops;
    T32Op op[kT32MaxOps];
    // input stage: 0 = the first in_k (<= 10) columns of fp32 rows x[row * in_ld + k], zero-padded to K = 16;
    //              1 = 64 fp32 columns per row (in_ld % 4 == 0, 16-byte aligned rows)
    int in_mode;
    const float* in_x; long long in_ld; int in_k;
    const unsigned char* wblob; int wblob_bytes;        // resident packed weights, staged once per CTA by TMA bulk copies
    const unsigned char* wstream;                       // 4 chunks of kT32ChunkBytes (64 channels each: hi then lo)
    const unsigned char* wcloud; long long wcloud_stride; int wcloud_bytes;   // per-cloud packed weights
    const float* bias; int n_bias;                      // bias table of the chain (copied to shared memory)
    const float* gbias; const int* group_rows; int n_groups;
    float* out_f32; long long out_ld; int out_col0;
    unsigned int* pool;                                 // [clouds, N] bit patterns of non-negative floats, zero-initialised
    float* logits; int n_classes;
    int n_clouds, rows_per_cloud;
    long long* prof;                                    // debugging aid (AMP_CHAIN32_PROF=1): phase timestamps of CTA 0 / slot 0
};

int tc_chain32_launch(const T32Params& p, cudaStream_t st);

// Weight packing for the chains above: fp32 W[n, k] (src[n * ld + k], or src[k * ld + n] when transposed), optionally scaled
// per output row (BatchNorm fold), split into bf16 hi + lo blocks [Kpad/8][Npad][8] at dst + dst_off (hi) and
// dst + dst_off + Npad * Kpad * 2 (lo); chunk_n != 0: independent blocks of chunk_n output rows (the streamed layout:
// chunk c at dst_off + c * chunk_n * Kpad * 4, hi then lo). Rows n >= N and columns k >= K are zero. Per cloud when
// src_cloud_stride != 0 (grid z = cloud).
struct T32PackJob {
    const float* src; long long ld; long long src_cloud_stride; const float* scale;
    int N, K, Npad, Kpad, transposed, chunk_n; long long dst_off; long long dst_cloud_stride;
};
struct T32PackTable { static constexpr int kMax = 16; int n; int n_clouds; T32PackJob job[kMax]; };
int t32_pack_weights(const T32PackTable& t, unsigned char* dst, cudaStream_t st);
inline int t32_packed_bytes(int Npad, int Kpad) { return Npad * Kpad * 4; }
// per-cloud conv_1 weights with the 3 x 3 input transform folded in (fold_input_transform of nn_common.cuh), scaled by the
// BatchNorm scale and packed: 64 x 16 hi + lo block at dst + cloud * dst_stride
int t32_fold_w1(const float* W1, const float* T, const float* scale, int n_clouds, unsigned char* dst, long long dst_stride, cudaStream_t st);
// dst[i] = scale[i] * bias[i] + shift[i] for i < n (null pointers: 1, 0, 0), zero for n <= i < n_pad: a bias-table entry
int t32_affine_bias(const float* bias, const float* scale, const float* shift, int n, int n_pad, float* dst, cudaStream_t st);

}  // namespace amp
