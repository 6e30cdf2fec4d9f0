// Fused shared-MLP chains on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands, fp32 accumulate.
//
// One kernel runs a whole chain of Conv1d(k=1) layers of pointNet/model/pointnetAtt.py (:31-35 T-Net convs + max-pool,
// :90-104 encoder convs + max-pool, :203-207 segmentation head) over 128-point tiles: the activations of a tile
// never leave the SM between layers (TMEM accumulator -> registers -> bias/ReLU -> bf16 -> shared-memory A
// operand of the next tcgen05.mma), the weights of the chain are staged once per CTA by TMA bulk copies, and the
// channel-wise max over the points of a cloud is taken in the epilogue of a transposed (channels x points) MMA.
// Eval mode only: BatchNorm is folded into the packed weights (scale) and a per-channel bias (shift).
#pragma once
#include <cuda_bf16.h>

#include "nn_common.cuh"

namespace amp {

constexpr int kTcMaxOps = 8;
constexpr int kTcTileRows = 128;

struct TcOp {
    int K, N;           // per tile: D[128, N] = A[128, K] * W[N, K]^T (+ bias);  K % 16 == 0, N % 16 == 0, 16 <= N <= 256
    int w_off;          // byte offset of the packed bf16 weights ([K/8][N][8]) in the shared blob, or in the per-cloud
    int w_cloud;        // block when w_cloud != 0
    int b_off;          // >= 0: byte offset (same blob as the weights) of the bias K group [N][8] (bf16 hi at k8 = 0, lo at
                        // k8 = 1): one more tcgen05.mma accumulates it against the constant ones block, so the bias add
                        // costs no ALU work. Pooled ops take their bias from pool_bias in the epilogue instead.
    int relu;
    int bias_grouped;   // + gbias[(cloud * n_groups + group(row)) * N + n] in the epilogue (per-block bias of the head)
    int write_act;      // bf16 result becomes the A operand of the next op
    int store_f32;      // result rows -> out_f32[row * out_ld + out_col0 + n]
    int pool;           // transposed MMA (channels x points); atomicMax(pool[cloud * N + n], max over the valid rows);
                        // needs relu (values >= 0 order like their bit patterns) and N % 128 == 0
    int store_logits;   // first n_classes columns -> logits[(cloud * n_classes + n) * rows_per_cloud + r]
};

struct TcChainParams {
    int n_ops;
    TcOp op[kTcMaxOps];
    // input stage
    //   0: the first in_k (3 or 9) columns of fp32 rows x[row * in_ld + k], split into bf16 hi + lo terms:
    //      A[:, k] = hi_k, A[:, Kpad / 2 + k] = lo_k, Kpad = op[0].K (16 for in_k = 3, 32 for in_k = 9); the packed
    //      weights of op[0] repeat W[:, k] at both positions, so the layer sees the input at ~16 bits of mantissa
    //   1: fp32 rows x[row * in_ld + k], k < op[0].K, rounded to bf16 (in_ld % 4 == 0, 16-byte aligned rows)
    //      in_bias != 0: columns in_k and in_k + 1 of the hi half are set to 1.0 and the packed weights carry the bias of
    //      op[0] there (TcPackJob.bias_col), so op[0] needs no bias K group
    int in_mode;
    const float* in_x; long long in_ld; int in_k; int in_bias;
    const unsigned char* wblob; int wblob_bytes;        // packed weights of the chain, staged by TMA bulk copies
    const unsigned char* wcloud; long long wcloud_stride; int wcloud_bytes;   // per-cloud packed weights
    const float* gbias; const int* group_rows; int n_groups;
    float* out_f32; long long out_ld; int out_col0;
    unsigned int* pool; const float* pool_bias;         // pool_bias[n]: added to the pooled op's maximum before the ReLU
    float* logits; int n_classes;
    int n_clouds, rows_per_cloud;
    long long* prof;                                    // debugging aid (AMP_CHAIN_PROF=1): phase timestamps of CTA 0 / slot 0
};

int tc_chain_launch(const TcChainParams& p, cudaStream_t st);

// weight packing: fp32 W[n, k] (element at src[n * ld + k], or src[k * ld + n] when transposed), optionally scaled per
// output row (BatchNorm fold) -> bf16 [Kpad/8][Npad][8] at dst + dst_off (rows n >= N and columns k >= K zero-filled);
// split_in_k != 0: column k and column Kpad/2 + k both hold W[:, k] for k < split_in_k (input stage 0).
// Bias b[n] = bias_scale[n] * bias[n] + bias_shift[n] (null pointers: 1, 0, 0) as bf16 hi + lo terms, either into the
// weight columns bias_col, bias_col + 1 (input stage 0 with in_bias) or as a separate K group [Npad][8] at
// dst + bias_dst_off (bias_dst_off >= 0). Per cloud when src_cloud_stride != 0.
struct TcPackJob {
    const float* src; long long ld; long long src_cloud_stride; const float* scale;
    const float* bias; const float* bias_scale; const float* bias_shift; int bias_col; long long bias_dst_off;
    int N, K, Npad, Kpad, transposed, split_in_k; long long dst_off; long long dst_cloud_stride;
};
struct TcPackTable { static constexpr int kMax = 12; int n; int n_clouds; TcPackJob job[kMax]; };
int tc_pack_weights(const TcPackTable& t, unsigned char* dst, cudaStream_t st);

inline int tc_packed_bytes(int Npad, int Kpad) { return Npad * Kpad * 2; }
inline int tc_bias_bytes(int Npad) { return Npad * 16; }

}  // namespace amp
