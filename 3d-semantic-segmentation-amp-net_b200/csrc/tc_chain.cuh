// Fused shared-MLP chains on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 operands, fp32 accumulate.
//
// One kernel runs a whole chain of Conv1d(k=1) layers of pointNet/model/pointnetAtt.py (:31-35 T-Net convs + max-pool,
// :90-104 encoder convs + max-pool, :203-207 segmentation head) over 128-point tiles: the activations of a tile
// never leave the SM between layers (TMEM accumulator -> registers -> BatchNorm/ReLU -> bf16 -> shared-memory A
// operand of the next tcgen05.mma), the weights of the chain are staged once per CTA by TMA bulk copies, and the
// channel-wise max over the points of a cloud is taken in the epilogue.
#pragma once
#include <cuda_bf16.h>

#include "nn_common.cuh"

namespace amp {

constexpr int kTcMaxOps = 6;
constexpr int kTcTileRows = 128;

struct TcOp {
    int K, N;           // per tile: D[128, N] = A[128, K] * W[N, K]^T;  K % 16 == 0, N % 16 == 0, 16 <= N <= 256
    int w_off;          // byte offset of the packed bf16 weights ([K/8][N][8]) in the shared blob; -1: per-cloud weights
    int affine_off;     // float offset of scale[N] then shift[N] in the tables; -1: none
    int relu;
    int bias_grouped;   // + gbias[(cloud * n_groups + group(row)) * N + n]   (per-block bias of the segmentation head)
    int write_act;      // bf16 result becomes the A operand of the next op
    int store_f32;      // result rows -> out_f32[row * out_ld + out_col0 + n]
    int store_bf16;     // result rows -> out_bf16[row * N + n]
    int pool;           // atomicMax(pool[cloud * N + n], max over the valid rows)   (values are >= 0 after ReLU)
    int store_logits;   // first n_classes columns -> logits[(cloud * n_classes + n) * rows_per_cloud + r]
};

struct TcChainParams {
    int n_ops;
    TcOp op[kTcMaxOps];
    // input stage: 0 = 64-channel layer on the CUDA cores from fp32 x (x[row * in_ld + k], k < in_k; weights in_w[n * in_k + k],
    //              per cloud when in_w_cloud_stride != 0), BatchNorm scale/shift at in_affine_off, ReLU
    //              1 = bf16 rows in_bf16[row * 64 + k];   2 = fp32 rows in_x[row * in_ld + k], k < 64
    int in_mode;
    const float* in_x; long long in_ld; int in_k;
    const float* in_w; long long in_w_cloud_stride; int in_affine_off;
    const __nv_bfloat16* in_bf16;
    const float* tables; int n_table_floats;            // BatchNorm scale / shift (and plain biases), staged in shared memory
    const unsigned char* wblob; int wblob_bytes;        // packed weights of the chain, staged by TMA bulk copy
    const unsigned char* wcloud; int wcloud_bytes;      // per-cloud packed weights (feature transform), wcloud + cloud * wcloud_bytes
    const float* gbias; const int* group_rows; int n_groups;
    float* out_f32; long long out_ld; int out_col0;
    __nv_bfloat16* out_bf16;
    unsigned int* pool;
    float* logits; int n_classes;
    int n_clouds, rows_per_cloud;
};

int tc_chain_launch(const TcChainParams& p, cudaStream_t st);

// weight packing: fp32 [N, K] (row stride ld; element (n, k) at src[n * ld + k], or src[k * ld + n] when transposed)
// -> bf16 [K/8][Npad][8] at dst (rows n >= N and columns k >= K zero-filled); per-cloud when src_cloud_stride != 0
struct TcPackJob { const float* src; long long ld; long long src_cloud_stride; int N, K, Npad, Kpad, transposed; long long dst_off; long long dst_cloud_stride; };
struct TcPackTable { static constexpr int kMax = 8; int n; int n_clouds; TcPackJob job[kMax]; };
int tc_pack_weights(const TcPackTable& t, unsigned char* dst, cudaStream_t st);

// T-Net FC stack in eval mode (pointnetAtt.py:38-46): pooled [B, 256] -> relu(bn4(fc1)) -> relu(bn5(fc2)) -> fc3 + bias + I
int tnet_fc_eval(const float* pooled, const float* fc1, const float* s4, const float* t4, const float* fc2, const float* s5,
                 const float* t5, const float* fc3_w, const float* fc3_b, int n_clouds, int d, float* h2_scratch, float* out,
                 cudaStream_t st);

}  // namespace amp
