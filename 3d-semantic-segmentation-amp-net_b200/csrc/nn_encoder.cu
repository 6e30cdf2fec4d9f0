// PointNet encoder: forward and backward host orchestration behind amp_encoder_fwd / amp_encoder_bwd.
//
// Replaces BasePointNet.forward (pointNet/model/pointnetAtt.py:80-112) with its two TransformationNets
// (:28-47) and, for training, the autograd backward that loss.backward() runs through it
// (pointNet/self-attention/train_pointnet-attention.py:467).
//
// Forward = 21 fused point-wise layers (nn_linear.cu). Eval: BatchNorm folded to scale/shift in the GEMM
// epilogue, activations stored once as final values, max-pools taken in the epilogue (pooled layers are
// never stored). Training: raw layer outputs are kept for backward, the batch statistics come out of the
// producing GEMM's epilogue and BatchNorm + ReLU are applied in the consumer's prologue.
// The 3x3 input transform is folded into per-cloud conv_1 weights (bmm + cat of :85-86 never materialise).
#include <limits>

#include "nn_layout.cuh"
#include "tc_chain.cuh"
#include "tc_chain32.cuh"

namespace amp {
namespace {

constexpr float kBnEps = 1e-5f, kBnMomentum = 0.1f;

struct EncCtx {
    const void* const* P;       // parameter pointers, state_dict order
    void* const* Gd;            // gradient pointers, same order (backward only)
    cudaStream_t st;
    int B, N;
    bool train;
    EncSaved S;
    float *part_sum, *part_sq;
    unsigned long long *pmax, *pmin;
    // backward-only
    float *k1, *k2, *k3;        // BatchNorm-backward coefficient tables [kEncBnTotal]
    float* wg; size_t wg_floats;
    void* split_dump;           // [tiles][hi, lo][<= 32 groups][128 rows] x 16 B: dy' of the layer being differentiated, split bf16
    const float* pf(int i) const { return reinterpret_cast<const float*>(P[i]); }
    float* gf(int i) const { return reinterpret_cast<float*>(Gd[i]); }
};

#define AMP_TRY(expr) do { int rc_ = (expr); if (rc_ != AMP_OK) return rc_; } while (0)
#define AMP_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(AMP_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); } while (0)

// one Conv1d / Linear (+ BatchNorm + ReLU) (+ max-pool over the rows of each cloud)
int fwd_layer(EncCtx& c, const float* X, long long ldx, int K, int in_bn, const float* W, long long ldw,
              long long wstride, int w_kn, const float* bias, float* Y, long long ldy, int Nout, int out_bn,
              float* pooled, int* arg, int clouds, int rows) {
    PwParams p{};
    p.X = X; p.ldx = ldx; p.K = K;
    if (c.train && in_bn >= 0) {
        const int o = enc_bn_offset(in_bn);
        p.in_a = c.S.scale + o; p.in_b = c.pf(kEncBnParam[in_bn] + BN_B); p.in_m = c.S.mean + o; p.in_relu = 1;
    }
    p.W = W; p.ldw = ldw; p.w_cloud_stride = wstride; p.w_kn = w_kn;
    p.bias = bias; p.n_groups = 1;
    p.fp16_split = c.train ? 1 : 0;
    p.Y = Y; p.ldy = ldy; p.n_clouds = clouds; p.rows_per_cloud = rows; p.Nout = Nout;
    const int o = out_bn >= 0 ? enc_bn_offset(out_bn) : 0;
    if (out_bn >= 0) {
        if (c.train) {
            p.part_sum = c.part_sum; p.part_sq = c.part_sq;
            if (pooled) {
                p.pool_mode = 2; p.pool_max = c.pmax; p.pool_min = c.pmin;
                AMP_CUDA(cudaMemsetAsync(c.pmax, 0, sizeof(unsigned long long) * clouds * Nout, c.st));
                AMP_CUDA(cudaMemsetAsync(c.pmin, 0, sizeof(unsigned long long) * clouds * Nout, c.st));
            }
        } else {
            p.out_scale = c.S.scale + o; p.out_shift = c.S.shift + o; p.out_relu = 1;
            if (pooled) {
                p.pool_mode = 1; p.pool_max = c.pmax; p.Y = nullptr;
                AMP_CUDA(cudaMemsetAsync(c.pmax, 0, sizeof(unsigned long long) * clouds * Nout, c.st));
            }
        }
    }
    AMP_TRY(pw_linear(p, c.st));
    if (out_bn >= 0 && c.train) {
        const int pb = kEncBnParam[out_bn];
        AMP_TRY(bn_finalize_train(c.part_sum, c.part_sq, clouds, rows, Nout, c.pf(pb + BN_W),
                                  const_cast<float*>(c.pf(pb + BN_RM)), const_cast<float*>(c.pf(pb + BN_RV)),
                                  reinterpret_cast<long long*>(const_cast<void*>(c.P[pb + BN_NBT])), kBnMomentum, kBnEps,
                                  c.S.scale + o, c.S.mean + o, c.S.invstd + o, c.st));
    }
    if (pooled) {
        if (c.train)
            AMP_TRY(pool_decode(c.pmax, c.pmin, 2, c.S.scale + o, c.pf(kEncBnParam[out_bn] + BN_B), c.S.mean + o, clouds, Nout,
                                pooled, arg, c.st));
        else
            AMP_TRY(pool_decode(c.pmax, c.pmin, 1, nullptr, nullptr, nullptr, clouds, Nout, pooled, arg, c.st));
    }
    return AMP_OK;
}

// eval-mode FC stack of a T-Net (:38-46) in one cluster launch; the wide fc_3 of the 64 x 64 transform stays a separate launch
int tnet_fc_stack_eval(EncCtx& c, int pbase, int L1, int d, const float* pool, float* f1, float* f2, float* out) {
    const int o4 = enc_bn_offset(L1 + 3), o5 = enc_bn_offset(L1 + 4);
    const int inside = d * d <= 64 ? 1 : 0;
    AMP_TRY(tnet_fc_eval(pool, c.B, c.pf(pbase + T_FC1), c.S.scale + o4, c.S.shift + o4, c.pf(pbase + T_FC2), c.S.scale + o5, c.S.shift + o5,
                         c.pf(pbase + T_FC3W), c.pf(pbase + T_FC3B), d, inside, f1, f2, out, c.st));
    if (inside) return AMP_OK;
    AMP_TRY(fwd_layer(c, f2, 128, 128, -1, c.pf(pbase + T_FC3W), 128, 0, 0, c.pf(pbase + T_FC3B), out, d * d, d * d, -1, nullptr, nullptr, 1, c.B));
    return add_identity(out, c.B, d, c.st);
}

// TransformationNet.forward (pointnetAtt.py:28-47): A = input rows (with the BatchNorm+ReLU of layer a_bn pending
// in training mode), out = [B, d*d] transform (+ identity)
int tnet_fwd(EncCtx& c, int pbase, int L1, int d, const float* A, long long lda, int K, int a_bn, float* y1, float* y2,
             float* y3, float* pool, int* arg, float* f1, float* f2, float* out) {
    const int B = c.B, N = c.N;
    AMP_TRY(fwd_layer(c, A, lda, K, a_bn, c.pf(pbase + T_CONV1), K, 0, 0, nullptr, y1, 64, 64, L1, nullptr, nullptr, B, N));
    AMP_TRY(fwd_layer(c, y1, 64, 64, L1, c.pf(pbase + T_CONV2), 64, 0, 0, nullptr, y2, 128, 128, L1 + 1, nullptr, nullptr, B, N));
    AMP_TRY(fwd_layer(c, y2, 128, 128, L1 + 1, c.pf(pbase + T_CONV3), 128, 0, 0, nullptr, y3, 256, 256, L1 + 2, pool, arg, B, N));
    if (!c.train) return tnet_fc_stack_eval(c, pbase, L1, d, pool, f1, f2, out);
    AMP_TRY(fwd_layer(c, pool, 256, 256, -1, c.pf(pbase + T_FC1), 256, 0, 0, nullptr, f1, 256, 256, L1 + 3, nullptr, nullptr, 1, B));
    AMP_TRY(fwd_layer(c, f1, 256, 256, L1 + 3, c.pf(pbase + T_FC2), 256, 0, 0, nullptr, f2, 128, 128, L1 + 4, nullptr, nullptr, 1, B));
    AMP_TRY(fwd_layer(c, f2, 128, 128, L1 + 4, c.pf(pbase + T_FC3W), 128, 0, 0, c.pf(pbase + T_FC3B), out, d * d, d * d, -1,
                      nullptr, nullptr, 1, B));
    return add_identity(out, B, d, c.st);
}

// packed chain blobs: weights [K/8][N][8] bf16, each followed (128-byte aligned) by its bias K group [N][8] where the
// layer's bias goes through the tensor pipe (tc_chain.cuh)
constexpr int kBlob1 = 2048 + (16384 + 2048) + 65536;                       // it.conv_1 (64 x 16 split, bias folded), it.conv_2 + bias, it.conv_3
constexpr int kBlob2 = (8192 + 1024) * 2 + (16384 + 2048) + 65536;          // conv_2, ft.conv_1, ft.conv_2 (+ biases), ft.conv_3
constexpr int kBlob3 = (8192 + 1024) * 2 + (16384 + 2048) + (32768 + 2048) + 65536;   // conv_2 .. conv_5 (+ biases), conv_6
constexpr int kWcW1 = 4096, kWcF = 8192, kWcStride = kWcW1 + kWcF;   // per cloud: W1eff (64 x 32 split, bias folded), F^T (64 x 64)

struct EncTc {
    float *scale, *shift;
    unsigned char *blobs, *wcloud;
    float *pools;                       // it_pool | ft_pool | G, [B, 256] each
    float *f1, *f2, *T, *W1eff;
};

EncTc enc_tc_carve(Arena& a, long long B) {
    EncTc t{};
    t.scale = a.take<float>(kEncBnTotal); t.shift = a.take<float>(kEncBnTotal);
    t.blobs = a.take<unsigned char>(kBlob1 + kBlob2 + kBlob3);
    t.wcloud = a.take<unsigned char>((size_t)B * kWcStride);
    t.pools = a.take<float>((size_t)B * 256 * 3);
    t.f1 = a.take<float>(B * 256); t.f2 = a.take<float>(B * 128); t.T = a.take<float>(B * 9 + 7); t.W1eff = a.take<float>(B * 576);
    return t;
}

size_t enc_fused32_bytes(long long B);       // per-call buffers + in-workspace pack of the fp32-class fused eval path (below)

size_t fwd_ws_bytes(long long B, long long N, bool training) {
    Arena a(nullptr, std::numeric_limits<size_t>::max());
    const size_t tiles = (size_t)pw_tiles((int)B, (int)N);
    a.take<float>(tiles * 256); a.take<float>(tiles * 256);
    a.take<unsigned long long>(B * 256); a.take<unsigned long long>(B * 256);
    if (!training) { enc_carve(a, B, N, false); enc_tc_carve(a, B); }
    return a.off + 256 + (training ? 0 : enc_fused32_bytes(B));
}

struct BwdWs {
    float *gA, *gB, *g64, *dG, *dF, *dT, *dW1eff, *d_f2, *d_f1, *dpool, *colsum;
    float* wg_pool; size_t wg_pool_floats;       // slices for the deferred weight-gradient reductions (WgDeferScope)
};

size_t wg_floats_needed(long long B, long long N) {
    size_t m = wgrad_workspace_floats((int)B, (int)N, 256, 128);
    const size_t fc3 = wgrad_workspace_floats(1, (int)B, 4096, 128);
    const size_t w1 = wgrad_workspace_floats((int)B, (int)N, 64, 64);
    if (fc3 > m) m = fc3;
    if (w1 > m) m = w1;
    return m;
}

BwdWs bwd_carve(Arena& a, EncCtx* c, long long B, long long N) {
    BwdWs w{};
    const size_t M = (size_t)B * N;
    const size_t tiles = (size_t)pw_tiles((int)B, (int)N);
    float* ps = a.take<float>(tiles * 256); float* pq = a.take<float>(tiles * 256);
    float* k1 = a.take<float>(kEncBnTotal); float* k2 = a.take<float>(kEncBnTotal); float* k3 = a.take<float>(kEncBnTotal);
    const size_t wgf = wg_floats_needed(B, N);
    float* wg = a.take<float>(wgf);
    if (c) { c->part_sum = ps; c->part_sq = pq; c->k1 = k1; c->k2 = k2; c->k3 = k3; c->wg = wg; c->wg_floats = wgf; }
    w.gA = a.take<float>(M * 256); w.gB = a.take<float>(M * 256); w.g64 = a.take<float>(M * 64);
    w.dG = a.take<float>(B * 256); w.dF = a.take<float>(B * 4096); w.dT = a.take<float>(B * 9 + 7);
    w.dW1eff = a.take<float>(B * 576); w.d_f2 = a.take<float>(B * 128); w.d_f1 = a.take<float>(B * 256);
    w.dpool = a.take<float>(B * 256);
    w.colsum = a.take<float>(colsum_scratch_floats((int)B, (int)N, 256));
    // every parameter gradient of one backward pass keeps its partials until the common reduction: sum over the layers
    // of clouds x slabs x (Nout x K + Nout) = 4.7 x the largest layer (256 x 128)
    w.wg_pool_floats = 5 * wgrad_workspace_floats((int)B, (int)N, 256, 128) + 64 * 32;
    w.wg_pool = a.take<float>(w.wg_pool_floats);
    // the input-gradient kernel leaves dy' (split bf16, up to 256 channels) here for the weight gradient of the same layer
    void* dump = a.take<unsigned char>(tiles * (size_t)(2 * 32 * 128 * 16));
    if (c) c->split_dump = dump;
    return w;
}

// BatchNorm-backward sums -> dgamma / dbeta of layer L and the coefficients its consumers apply
int bwd_finalize(EncCtx& c, int L, int tiles, long long count) {
    const int o = enc_bn_offset(L), pb = kEncBnParam[L];
    return bn_backward_finalize(c.part_sum, c.part_sq, tiles, count, kEncBnChannels[L], c.pf(pb + BN_W), c.S.mean + o,
                                c.S.invstd + o, c.gf(pb + BN_W), c.gf(pb + BN_B), 0, c.k1 + o, c.k2 + o, c.k3 + o, c.st);
}

// Backward through  y_out = act_in(A) @ W^T (+ b)  given dz (gradient w.r.t. the BatchNorm OUTPUT of layer L_out
// after its ReLU mask; L_out < 0: dz is the plain gradient of y_out):
//   dW (+ db) = dy^T act_in(A),   dA = mask_in(dy @ W)   with dy = bn_backward(dz) applied on the fly.
// act_in = ReLU(BatchNorm_{L_in}(.)) when L_in >= 0. dA gets the ReLU mask and the BatchNorm-backward sums of L_in
// unless defer (the tensor has another gradient contribution still to come).
int bwd_step(EncCtx& c, const float* dz, int Nout, int L_out, const float* y_out, const float* A, long long lda, int K,
             int L_in, const float* W, float* dW, float* db, float* dA, long long ldda, int accumulate, bool defer,
             int clouds, int rows) {
    const int oo = L_out >= 0 ? enc_bn_offset(L_out) : 0, oi = L_in >= 0 ? enc_bn_offset(L_in) : 0;
    WgParams g{};
    g.dY = dz; g.lddy = Nout; g.Nout = Nout;
    if (L_out >= 0) { g.y_a = c.k1 + oo; g.y_b = c.k3 + oo; g.y_c = c.k2 + oo; g.y_m = c.S.mean + oo; g.Y2 = y_out; }
    g.A = A; g.lda = lda; g.K = K;
    if (L_in >= 0) { g.a_a = c.S.scale + oi; g.a_b = c.pf(kEncBnParam[L_in] + BN_B); g.a_m = c.S.mean + oi; g.a_relu = 1; }
    g.n_clouds = clouds; g.rows_per_cloud = rows;
    g.dW = dW; g.ldw = K; g.db = db;
    g.partials = c.wg; g.partial_floats = c.wg_floats;
    if (!dA) return wgrad(g, c.st);
    PwParams p{};
    p.X = dz; p.ldx = Nout; p.K = Nout;
    if (L_out >= 0) { p.in_a = c.k1 + oo; p.in_b = c.k3 + oo; p.in_c = c.k2 + oo; p.in_m = c.S.mean + oo; p.X2 = y_out; }
    p.W = W; p.ldw = K; p.w_kn = 1;
    p.n_groups = 1;
    p.accumulate = accumulate;
    p.Y = dA; p.ldy = ldda; p.n_clouds = clouds; p.rows_per_cloud = rows; p.Nout = K;
    p.splitk_ws = c.wg; p.splitk_floats = c.wg_floats;         // (the weight-gradient partials above have been reduced: stream order)
    const bool mask = L_in >= 0 && !defer;
    if (mask) {
        p.mask_y = A; p.ld_mask = lda; p.mask_scale = c.S.scale + oi; p.mask_shift = c.pf(kEncBnParam[L_in] + BN_B);
        p.mask_mean = c.S.mean + oi; p.mask_invstd = c.S.invstd + oi;
        p.part_sum = c.part_sum; p.part_sq = c.part_sq;
    }
    // The input gradient runs first: its tensor-core kernel builds dy' = bn_backward(dz) split into bf16 hi + lo as its own
    // operand and leaves a copy in c.split_dump, which the weight-gradient kernel of this layer then copies instead of
    // reading dz and y_out again and redoing the prologue and the split (two thirds of its instructions for a 256-wide dy').
    const bool want_dump = c.split_dump && Nout % 8 == 0 && Nout <= 256 && !path_disabled("wgrad_presplit");
    if (want_dump) p.split_dump = c.split_dump;
    tc_layer_dumped() = false;
    AMP_TRY(pw_linear(p, c.st));
    if (want_dump && tc_layer_dumped()) g.dy_split = c.split_dump;
    if (mask) AMP_TRY(bwd_finalize(c, L_in, pw_tiles(clouds, rows), (long long)clouds * rows));
    return wgrad(g, c.st);
}

// backward of max-pool over the rows + ReLU + BatchNorm of the pooled layer L (raw output y [M, 256])
int bwd_pool(EncCtx& c, const float* dpool, const int* arg, const float* y, int L, float* dz) {
    const int o = enc_bn_offset(L);
    const size_t M = (size_t)c.B * c.N;
    AMP_CUDA(cudaMemsetAsync(dz, 0, M * 256 * sizeof(float), c.st));
    AMP_TRY(pool_scatter_bwd(dpool, arg, y, c.S.scale + o, c.pf(kEncBnParam[L] + BN_B), c.S.mean + o, c.S.invstd + o, c.B, c.N, 256, dz,
                             c.part_sum, c.part_sq, c.st));
    return bwd_finalize(c, L, 1, (long long)M);
}

// backward of TransformationNet given dM [B, d*d]; gS [M,256] and gT [M,128] are scratch
int tnet_bwd(EncCtx& c, BwdWs& w, int pbase, int L1, int d, const float* dM, const float* y1, const float* y2,
             const float* y3, const float* pool, const int* arg, const float* f1, const float* f2, const float* A,
             long long lda, int K, int a_bn, float* dX, long long lddx, float* gS, float* gT) {
    const int B = c.B, N = c.N;
    AMP_TRY(bwd_step(c, dM, d * d, -1, nullptr, f2, 128, 128, L1 + 4, c.pf(pbase + T_FC3W), c.gf(pbase + T_FC3W),
                     c.gf(pbase + T_FC3B), w.d_f2, 128, 0, false, 1, B));
    AMP_TRY(bwd_step(c, w.d_f2, 128, L1 + 4, f2, f1, 256, 256, L1 + 3, c.pf(pbase + T_FC2), c.gf(pbase + T_FC2), nullptr,
                     w.d_f1, 256, 0, false, 1, B));
    AMP_TRY(bwd_step(c, w.d_f1, 256, L1 + 3, f1, pool, 256, 256, -1, c.pf(pbase + T_FC1), c.gf(pbase + T_FC1), nullptr,
                     w.dpool, 256, 0, false, 1, B));
    AMP_TRY(bwd_pool(c, w.dpool, arg, y3, L1 + 2, gS));
    AMP_TRY(bwd_step(c, gS, 256, L1 + 2, y3, y2, 128, 128, L1 + 1, c.pf(pbase + T_CONV3), c.gf(pbase + T_CONV3), nullptr,
                     gT, 128, 0, false, B, N));
    AMP_TRY(bwd_step(c, gT, 128, L1 + 1, y2, y1, 64, 64, L1, c.pf(pbase + T_CONV2), c.gf(pbase + T_CONV2), nullptr,
                     gS, 64, 0, false, B, N));
    // conv_1: the input gradient (if any) accumulates onto dX and closes the fan-in of its BatchNorm layer
    AMP_TRY(bwd_step(c, gS, 64, L1, y1, A, lda, K, a_bn, c.pf(pbase + T_CONV1), c.gf(pbase + T_CONV1), nullptr,
                     dX, lddx, 1, false, B, N));
    return AMP_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// bf16 tensor-core eval path (AMP_PREC_BF16): three fused tcgen05 chains split at the cloud-wide max-pools
//   chain 1: xyz -> input T-Net convs -> max          chain 2: x -> conv_1, conv_2 -> feature T-Net convs -> max
//   chain 3: x -> conv_1, conv_2 -> bmm(F) = local features (stored) -> conv_3 .. conv_6 -> max
// conv_1 / conv_2 are recomputed in chain 3 instead of storing their activations (4.7 kMAC per point against a
// 128 B/point round trip); the T-Net FC stacks run on the CUDA cores in fp32 (B rows only).
// ---------------------------------------------------------------------------------------------------------------
inline TcOp tc_op(int K, int N, int w_off, int w_cloud, int b_off, int relu) {
    TcOp o{}; o.K = K; o.N = N; o.w_off = w_off; o.w_cloud = w_cloud; o.b_off = b_off; o.relu = relu; o.write_act = 1;
    return o;
}
// pack job of a BatchNorm-folded layer: weights scaled by the BN scale, bias = BN shift
inline TcPackJob tc_job(const float* w, int ld, const float* scale, const float* shift, int N, int K, int Kpad, int split_in_k,
                        int bias_col, long long bias_dst_off, long long dst_off) {
    return TcPackJob{w, ld, 0, scale, nullptr, nullptr, shift, bias_col, bias_dst_off, N, K, N, Kpad, 0, split_in_k, dst_off, 0};
}

int tnet_fc_fwd(EncCtx& c, int pbase, int L1, int d, const float* pool, float* f1, float* f2, float* out) {
    return tnet_fc_stack_eval(c, pbase, L1, d, pool, f1, f2, out);
}

int encoder_fwd_bf16(EncCtx& c, const float* x, float* out, float* feat_t, Arena& wa) {
    const int B = c.B, N = c.N;
    EncTc t = enc_tc_carve(wa, B);
    if (!wa.ok()) return fail(AMP_E_WORKSPACE, "encoder_fwd: workspace too small");
    c.S.scale = t.scale; c.S.shift = t.shift;
    {
        BnDesc d[L_ENC_BN];
        for (int L = 0; L < L_ENC_BN; ++L) {
            const int pb = kEncBnParam[L], o = enc_bn_offset(L);
            d[L] = BnDesc{c.pf(pb + BN_W), c.pf(pb + BN_B), c.pf(pb + BN_RM), c.pf(pb + BN_RV), t.scale + o, t.shift + o, kEncBnChannels[L]};
        }
        AMP_TRY(bn_fold_eval(d, L_ENC_BN, kBnEps, c.st));
    }
    auto sc = [&](int L) { return t.scale + enc_bn_offset(L); };
    auto sh = [&](int L) { return t.shift + enc_bn_offset(L); };
    // shared weights of the three chains, BatchNorm scale folded into the rows, BatchNorm shift as the bias K group
    const int b1 = 0, b2 = kBlob1, b3 = kBlob1 + kBlob2;
    // chain 1 layout
    const int c1_w1 = 0, c1_w2 = 2048, c1_b2 = c1_w2 + 16384, c1_w3 = c1_b2 + 2048;
    // chain 2 layout
    const int c2_w2 = 0, c2_b2 = 8192, c2_f1 = c2_b2 + 1024, c2_fb1 = c2_f1 + 8192, c2_f2 = c2_fb1 + 1024, c2_fb2 = c2_f2 + 16384,
              c2_f3 = c2_fb2 + 2048;
    // chain 3 layout
    const int c3_w2 = 0, c3_b2 = 8192, c3_w3 = c3_b2 + 1024, c3_b3 = c3_w3 + 8192, c3_w4 = c3_b3 + 1024, c3_b4 = c3_w4 + 16384,
              c3_w5 = c3_b4 + 2048, c3_b5 = c3_w5 + 32768, c3_w6 = c3_b5 + 2048;
    {
        TcPackTable pt{};
        pt.n = 12; pt.n_clouds = 1;
        pt.job[0] = tc_job(c.pf(E_IT + T_CONV1), 3, sc(L_IT1), sh(L_IT1), 64, 3, 16, 3, 3, -1, b1 + c1_w1);
        pt.job[1] = tc_job(c.pf(E_IT + T_CONV2), 64, sc(L_IT2), sh(L_IT2), 128, 64, 64, 0, -1, b1 + c1_b2, b1 + c1_w2);
        pt.job[2] = tc_job(c.pf(E_IT + T_CONV3), 128, sc(L_IT3), nullptr, 256, 128, 128, 0, -1, -1, b1 + c1_w3);
        pt.job[3] = tc_job(c.pf(E_CONV2), 64, sc(L_C2), sh(L_C2), 64, 64, 64, 0, -1, b2 + c2_b2, b2 + c2_w2);
        pt.job[4] = tc_job(c.pf(E_FT + T_CONV1), 64, sc(L_FT1), sh(L_FT1), 64, 64, 64, 0, -1, b2 + c2_fb1, b2 + c2_f1);
        pt.job[5] = tc_job(c.pf(E_FT + T_CONV2), 64, sc(L_FT2), sh(L_FT2), 128, 64, 64, 0, -1, b2 + c2_fb2, b2 + c2_f2);
        pt.job[6] = tc_job(c.pf(E_FT + T_CONV3), 128, sc(L_FT3), nullptr, 256, 128, 128, 0, -1, -1, b2 + c2_f3);
        pt.job[7] = tc_job(c.pf(E_CONV2), 64, sc(L_C2), sh(L_C2), 64, 64, 64, 0, -1, b3 + c3_b2, b3 + c3_w2);
        pt.job[8] = tc_job(c.pf(E_CONV3), 64, sc(L_C3), sh(L_C3), 64, 64, 64, 0, -1, b3 + c3_b3, b3 + c3_w3);
        pt.job[9] = tc_job(c.pf(E_CONV4), 64, sc(L_C4), sh(L_C4), 128, 64, 64, 0, -1, b3 + c3_b4, b3 + c3_w4);
        pt.job[10] = tc_job(c.pf(E_CONV5), 128, sc(L_C5), sh(L_C5), 128, 128, 128, 0, -1, b3 + c3_b5, b3 + c3_w5);
        pt.job[11] = tc_job(c.pf(E_CONV6), 128, sc(L_C6), nullptr, 256, 128, 128, 0, -1, -1, b3 + c3_w6);
        AMP_TRY(tc_pack_weights(pt, t.blobs, c.st));
    }
    AMP_CUDA(cudaMemsetAsync(t.pools, 0, sizeof(float) * B * 256 * 3, c.st));
    float* it_pool = t.pools; float* ft_pool = t.pools + (size_t)B * 256; float* G = t.pools + (size_t)B * 512;
    TcChainParams base{};
    base.in_mode = 0; base.in_x = x; base.in_ld = 9; base.in_bias = 1; base.n_groups = 1; base.n_clouds = B; base.rows_per_cloud = N;
    // chain 1: input T-Net convs on xyz + max-pool (:31-35)
    {
        TcChainParams p = base;
        p.in_k = 3; p.n_ops = 3;
        p.wblob = t.blobs + b1; p.wblob_bytes = kBlob1;
        p.op[0] = tc_op(16, 64, c1_w1, 0, -1, 1);
        p.op[1] = tc_op(64, 128, c1_w2, 0, c1_b2, 1);
        p.op[2] = tc_op(128, 256, c1_w3, 0, -1, 1); p.op[2].write_act = 0; p.op[2].pool = 1;
        p.pool = reinterpret_cast<unsigned int*>(it_pool); p.pool_bias = sh(L_IT3);
        AMP_TRY(tc_chain_launch(p, c.st));
    }
    AMP_TRY(tnet_fc_fwd(c, E_IT, L_IT1, 3, it_pool, t.f1, t.f2, t.T));
    // bmm + cat + conv_1 (:85-90) as per-cloud conv_1 weights, packed with the bn_1 scale; bn_1 shift rides in two spare input columns
    AMP_TRY(fold_input_transform(c.pf(E_CONV1), t.T, B, t.W1eff, c.st));
    {
        TcPackTable pt{};
        pt.n = 1; pt.n_clouds = B;
        pt.job[0] = TcPackJob{t.W1eff, 9, 576, sc(L_C1), nullptr, nullptr, sh(L_C1), 9, -1, 64, 9, 64, 32, 0, 9, 0, kWcStride};
        AMP_TRY(tc_pack_weights(pt, t.wcloud, c.st));
    }
    // chain 2: conv_1, conv_2, feature T-Net convs + max-pool (:90-94)
    {
        TcChainParams p = base;
        p.in_k = 9; p.n_ops = 5;
        p.wblob = t.blobs + b2; p.wblob_bytes = kBlob2;
        p.wcloud = t.wcloud; p.wcloud_stride = kWcStride; p.wcloud_bytes = kWcW1;
        p.op[0] = tc_op(32, 64, 0, 1, -1, 1);
        p.op[1] = tc_op(64, 64, c2_w2, 0, c2_b2, 1);
        p.op[2] = tc_op(64, 64, c2_f1, 0, c2_fb1, 1);
        p.op[3] = tc_op(64, 128, c2_f2, 0, c2_fb2, 1);
        p.op[4] = tc_op(128, 256, c2_f3, 0, -1, 1); p.op[4].write_act = 0; p.op[4].pool = 1;
        p.pool = reinterpret_cast<unsigned int*>(ft_pool); p.pool_bias = sh(L_FT3);
        AMP_TRY(tc_chain_launch(p, c.st));
    }
    AMP_TRY(tnet_fc_fwd(c, E_FT, L_FT1, 64, ft_pool, t.f1, t.f2, feat_t));
    {
        TcPackTable pt{};
        pt.n = 1; pt.n_clouds = B;
        pt.job[0] = TcPackJob{feat_t, 64, 4096, nullptr, nullptr, nullptr, nullptr, -1, -1, 64, 64, 64, 64, 1, 0, kWcW1, kWcStride};
        AMP_TRY(tc_pack_weights(pt, t.wcloud, c.st));
    }
    // chain 3: conv_1, conv_2, bmm with the feature transform (= local features, :96-97), conv_3 .. conv_6 + max-pool
    {
        TcChainParams p = base;
        p.in_k = 9; p.n_ops = 7;
        p.wblob = t.blobs + b3; p.wblob_bytes = kBlob3;
        p.wcloud = t.wcloud; p.wcloud_stride = kWcStride; p.wcloud_bytes = kWcStride;
        p.op[0] = tc_op(32, 64, 0, 1, -1, 1);
        p.op[1] = tc_op(64, 64, c3_w2, 0, c3_b2, 1);
        p.op[2] = tc_op(64, 64, kWcW1, 1, -1, 0); p.op[2].store_f32 = 1;
        p.op[3] = tc_op(64, 64, c3_w3, 0, c3_b3, 1);
        p.op[4] = tc_op(64, 128, c3_w4, 0, c3_b4, 1);
        p.op[5] = tc_op(128, 128, c3_w5, 0, c3_b5, 1);
        p.op[6] = tc_op(128, 256, c3_w6, 0, -1, 1); p.op[6].write_act = 0; p.op[6].pool = 1;
        p.out_f32 = out; p.out_ld = 320; p.out_col0 = 256;
        p.pool = reinterpret_cast<unsigned int*>(G); p.pool_bias = sh(L_C6);
        AMP_TRY(tc_chain_launch(p, c.st));
    }
    // repeat + cat (:109-110)
    return broadcast_rows(G, B, N, 256, out, 320, c.st);
}

// ---------------------------------------------------------------------------------------------------------------
// fp32-class fused eval path (AMP_PREC_FP32, eval): the same chains as above on tc_chain32_kernel (split-bf16 operands,
// activations resident in tensor memory), chain 3 cut at the local features (a module output that is stored anyway):
//   chain 1: xyz -> input T-Net convs -> max            chain 2: x -> conv_1, conv_2 -> feature T-Net convs -> max
//   chain 3a: x -> conv_1, conv_2 -> bmm(F) = local features (stored)
//   chain 3b: local -> conv_3 .. conv_5 -> conv_6 (weights streamed) -> max
// The BatchNorm-folded, hi / lo split weights live in a caller-owned pack cache that is rebuilt only when the caller says
// the parameters changed (pack_valid == 0).
// ---------------------------------------------------------------------------------------------------------------
constexpr int kP32W_it1 = 0, kP32W_it2 = kP32W_it1 + 64 * 16 * 4, kP32W_it3 = kP32W_it2 + 128 * 64 * 4, kP32Blob1 = kP32W_it3 + 256 * 128 * 4;
constexpr int kP32W_c2 = 0, kP32W_ft1 = kP32W_c2 + 64 * 64 * 4, kP32W_ft2 = kP32W_ft1 + 64 * 64 * 4, kP32W_ft3 = kP32W_ft2 + 128 * 64 * 4,
              kP32Blob2 = kP32W_ft3 + 256 * 128 * 4;
constexpr int kP32Blob3a = 64 * 64 * 4;                                                   // conv_2
constexpr int kP32W_c3 = 0, kP32W_c4 = kP32W_c3 + 64 * 64 * 4, kP32W_c5 = kP32W_c4 + 128 * 64 * 4, kP32Blob3b = kP32W_c5 + 128 * 128 * 4;
constexpr int kP32Stream = 256 * 128 * 4;                                                 // conv_6, 4 chunks of 64 channels
constexpr int kP32CloudW1 = 64 * 16 * 4, kP32CloudF = 64 * 64 * 4, kP32CloudStride = kP32CloudW1 + kP32CloudF;

struct EncPack { float *scale, *shift; unsigned char *blob1, *blob2, *blob3a, *blob3b, *stream; };
struct EncT32 { unsigned char* wcloud; float *pools, *f1, *f2, *T, *W1eff; unsigned int* bars; };      // per-call buffers
EncT32 enc_t32_carve(Arena& a, long long B) {
    EncT32 t{};
    t.wcloud = a.take<unsigned char>((size_t)B * kP32CloudStride);
    t.pools = a.take<float>((size_t)B * 256 * 3 + 16);    // + the grid-barrier words of the two T-Net FC launches (zeroed with the pools)
    t.bars = reinterpret_cast<unsigned int*>(t.pools + (size_t)B * 256 * 3);
    t.f1 = a.take<float>(B * 256); t.f2 = a.take<float>(B * 128); t.T = a.take<float>(B * 9 + 7); t.W1eff = a.take<float>(B * 576);
    return t;
}
EncPack enc_pack_carve(Arena& a) {
    EncPack k{};
    k.scale = a.take<float>(kEncBnTotal); k.shift = a.take<float>(kEncBnTotal);
    k.blob1 = a.take<unsigned char>(kP32Blob1); k.blob2 = a.take<unsigned char>(kP32Blob2); k.blob3a = a.take<unsigned char>(kP32Blob3a);
    k.blob3b = a.take<unsigned char>(kP32Blob3b); k.stream = a.take<unsigned char>(kP32Stream);
    return k;
}
size_t enc_pack_bytes() {
    Arena a(nullptr, std::numeric_limits<size_t>::max());
    enc_pack_carve(a);
    return a.off + 256;
}

size_t enc_fused32_bytes(long long B) {
    Arena a(nullptr, std::numeric_limits<size_t>::max());
    enc_t32_carve(a, B); enc_pack_carve(a);
    return a.off + 512;
}

inline T32Op t32_op(int K, int N, int w_off, int w_cloud, int bias_off, int relu) {
    T32Op o{}; o.K = K; o.N = N; o.w_off = w_off; o.w_cloud = w_cloud; o.bias_off = bias_off; o.relu = relu; o.write_act = 1;
    return o;
}
inline T32PackJob t32_job(const float* w, int ld, const float* scale, int N, int K, int Kpad, long long dst_off, int chunk_n = 0) {
    return T32PackJob{w, ld, 0, scale, N, K, N, Kpad, 0, chunk_n, dst_off, 0};
}

int encoder_pack_fp32(EncCtx& c, const EncPack& k) {
    BnDesc d[L_ENC_BN];
    for (int L = 0; L < L_ENC_BN; ++L) {
        const int pb = kEncBnParam[L], o = enc_bn_offset(L);
        d[L] = BnDesc{c.pf(pb + BN_W), c.pf(pb + BN_B), c.pf(pb + BN_RM), c.pf(pb + BN_RV), k.scale + o, k.shift + o, kEncBnChannels[L]};
    }
    AMP_TRY(bn_fold_eval(d, L_ENC_BN, kBnEps, c.st));
    auto sc = [&](int L) { return k.scale + enc_bn_offset(L); };
    // one pack launch per destination buffer (dst offsets are relative to it)
    {
        T32PackTable pt{}; pt.n = 3; pt.n_clouds = 1;
        pt.job[0] = t32_job(c.pf(E_IT + T_CONV1), 3, sc(L_IT1), 64, 3, 16, kP32W_it1);
        pt.job[1] = t32_job(c.pf(E_IT + T_CONV2), 64, sc(L_IT2), 128, 64, 64, kP32W_it2);
        pt.job[2] = t32_job(c.pf(E_IT + T_CONV3), 128, sc(L_IT3), 256, 128, 128, kP32W_it3);
        AMP_TRY(t32_pack_weights(pt, k.blob1, c.st));
    }
    {
        T32PackTable pt{}; pt.n = 4; pt.n_clouds = 1;
        pt.job[0] = t32_job(c.pf(E_CONV2), 64, sc(L_C2), 64, 64, 64, kP32W_c2);
        pt.job[1] = t32_job(c.pf(E_FT + T_CONV1), 64, sc(L_FT1), 64, 64, 64, kP32W_ft1);
        pt.job[2] = t32_job(c.pf(E_FT + T_CONV2), 64, sc(L_FT2), 128, 64, 64, kP32W_ft2);
        pt.job[3] = t32_job(c.pf(E_FT + T_CONV3), 128, sc(L_FT3), 256, 128, 128, kP32W_ft3);
        AMP_TRY(t32_pack_weights(pt, k.blob2, c.st));
    }
    {
        T32PackTable pt{}; pt.n = 1; pt.n_clouds = 1;
        pt.job[0] = t32_job(c.pf(E_CONV2), 64, sc(L_C2), 64, 64, 64, 0);
        AMP_TRY(t32_pack_weights(pt, k.blob3a, c.st));
    }
    {
        T32PackTable pt{}; pt.n = 3; pt.n_clouds = 1;
        pt.job[0] = t32_job(c.pf(E_CONV3), 64, sc(L_C3), 64, 64, 64, kP32W_c3);
        pt.job[1] = t32_job(c.pf(E_CONV4), 64, sc(L_C4), 128, 64, 64, kP32W_c4);
        pt.job[2] = t32_job(c.pf(E_CONV5), 128, sc(L_C5), 128, 128, 128, kP32W_c5);
        AMP_TRY(t32_pack_weights(pt, k.blob3b, c.st));
    }
    {
        T32PackTable pt{}; pt.n = 1; pt.n_clouds = 1;
        pt.job[0] = t32_job(c.pf(E_CONV6), 128, sc(L_C6), 256, 128, 128, 0, kT32ChunkChannels);
        AMP_TRY(t32_pack_weights(pt, k.stream, c.st));
    }
    return AMP_OK;
}

int encoder_fwd_fp32_fused(EncCtx& c, const float* x, float* out, float* feat_t, Arena& wa, void* pack, size_t pack_bytes, int pack_valid) {
    const int B = c.B, N = c.N;
    EncT32 t = enc_t32_carve(wa, B);                     // per-call buffers (pools, FC activations, per-cloud packed weights)
    EncPack k;
    if (pack) {
        if (pack_bytes < enc_pack_bytes()) return fail(AMP_E_WORKSPACE, "encoder_fwd: pack cache too small");
        Arena pa(pack, pack_bytes);
        k = enc_pack_carve(pa);
    } else {
        k = enc_pack_carve(wa);
        pack_valid = 0;
    }
    if (!wa.ok()) return fail(AMP_E_WORKSPACE, "encoder_fwd: workspace too small");
    if (!pack_valid) AMP_TRY(encoder_pack_fp32(c, k));
    c.S.scale = k.scale; c.S.shift = k.shift;            // tnet_fc_stack_eval reads the folded bn_4 / bn_5 from here
    auto sc = [&](int L) { return k.scale + enc_bn_offset(L); };
    AMP_CUDA(cudaMemsetAsync(t.pools, 0, sizeof(float) * ((size_t)B * 256 * 3 + 16), c.st));
    float* it_pool = t.pools; float* ft_pool = t.pools + (size_t)B * 256; float* G = t.pools + (size_t)B * 512;
    T32Params base{};
    base.in_mode = 0; base.in_x = x; base.in_ld = 9; base.n_groups = 1; base.n_clouds = B; base.rows_per_cloud = N;
    base.bias = k.shift; base.n_bias = kEncBnTotal;      // every op's bias is a slice of the folded-BatchNorm shift table
    // chain 1: input T-Net convs on xyz + max-pool (:31-35)
    {
        T32Params p = base;
        p.in_k = 3; p.n_ops = 3;
        p.wblob = k.blob1; p.wblob_bytes = kP32Blob1;
        p.op[0] = t32_op(16, 64, kP32W_it1, 0, enc_bn_offset(L_IT1), 1);
        p.op[1] = t32_op(64, 128, kP32W_it2, 0, enc_bn_offset(L_IT2), 1);
        p.op[2] = t32_op(128, 256, kP32W_it3, 0, enc_bn_offset(L_IT3), 1); p.op[2].write_act = 0; p.op[2].pool = 1;
        p.pool = reinterpret_cast<unsigned int*>(it_pool);
        AMP_TRY(tc_chain32_launch(p, c.st));
    }
    {
        const int o4 = enc_bn_offset(L_IT1 + 3), o5 = enc_bn_offset(L_IT1 + 4);
        const int rc = tnet_fc_grid(it_pool, B, c.pf(E_IT + T_FC1), k.scale + o4, k.shift + o4, c.pf(E_IT + T_FC2), k.scale + o5, k.shift + o5,
                                    c.pf(E_IT + T_FC3W), c.pf(E_IT + T_FC3B), 3, t.f1, t.f2, t.T, nullptr, 0, t.bars, c.st);
        if (rc < 0) return rc;
        if (rc == 0) AMP_TRY(tnet_fc_stack_eval(c, E_IT, L_IT1, 3, it_pool, t.f1, t.f2, t.T));
    }
    // bmm + cat + conv_1 (:85-90) as per-cloud conv_1 weights (bn_1 scale folded in), packed per cloud
    AMP_TRY(t32_fold_w1(c.pf(E_CONV1), t.T, sc(L_C1), B, t.wcloud, kP32CloudStride, c.st));
    // chain 2: conv_1, conv_2, feature T-Net convs + max-pool (:90-94)
    {
        T32Params p = base;
        p.in_k = 9; p.n_ops = 5;
        p.wblob = k.blob2; p.wblob_bytes = kP32Blob2;
        p.wcloud = t.wcloud; p.wcloud_stride = kP32CloudStride; p.wcloud_bytes = kP32CloudW1;
        p.op[0] = t32_op(16, 64, 0, 1, enc_bn_offset(L_C1), 1);
        p.op[1] = t32_op(64, 64, kP32W_c2, 0, enc_bn_offset(L_C2), 1);
        p.op[2] = t32_op(64, 64, kP32W_ft1, 0, enc_bn_offset(L_FT1), 1);
        p.op[3] = t32_op(64, 128, kP32W_ft2, 0, enc_bn_offset(L_FT2), 1);
        p.op[4] = t32_op(128, 256, kP32W_ft3, 0, enc_bn_offset(L_FT3), 1); p.op[4].write_act = 0; p.op[4].pool = 1;
        p.pool = reinterpret_cast<unsigned int*>(ft_pool);
        AMP_TRY(tc_chain32_launch(p, c.st));
    }
    {   // fc_1, fc_2 in one cluster launch; fc_3 + identity + the packed operand of local = h @ F in a second one
        const int o4 = enc_bn_offset(L_FT1 + 3), o5 = enc_bn_offset(L_FT1 + 4);
        const int rc = tnet_fc_grid(ft_pool, B, c.pf(E_FT + T_FC1), k.scale + o4, k.shift + o4, c.pf(E_FT + T_FC2), k.scale + o5, k.shift + o5,
                                    c.pf(E_FT + T_FC3W), c.pf(E_FT + T_FC3B), 64, t.f1, t.f2, feat_t, t.wcloud + kP32CloudW1, kP32CloudStride,
                                    t.bars + 4, c.st);
        if (rc < 0) return rc;
        if (rc == 0) {
            AMP_TRY(tnet_fc_eval(ft_pool, B, c.pf(E_FT + T_FC1), k.scale + o4, k.shift + o4, c.pf(E_FT + T_FC2), k.scale + o5, k.shift + o5,
                                 c.pf(E_FT + T_FC3W), c.pf(E_FT + T_FC3B), 64, 0, t.f1, t.f2, feat_t, c.st));
            AMP_TRY(tnet_fc3_pack(t.f2, c.pf(E_FT + T_FC3W), c.pf(E_FT + T_FC3B), B, feat_t, t.wcloud + kP32CloudW1, kP32CloudStride, c.st));
        }
    }
    // chain 3a: conv_1, conv_2, bmm with the feature transform = local features (:96-97), stored into out[:, :, 256:320]
    {
        T32Params p = base;
        p.in_k = 9; p.n_ops = 3;
        p.wblob = k.blob3a; p.wblob_bytes = kP32Blob3a;
        p.wcloud = t.wcloud; p.wcloud_stride = kP32CloudStride; p.wcloud_bytes = kP32CloudStride;
        p.op[0] = t32_op(16, 64, 0, 1, enc_bn_offset(L_C1), 1);
        p.op[1] = t32_op(64, 64, 0, 0, enc_bn_offset(L_C2), 1);
        p.op[2] = t32_op(64, 64, kP32CloudW1, 1, -1, 0); p.op[2].write_act = 0; p.op[2].store_f32 = 1;
        p.out_f32 = out; p.out_ld = 320; p.out_col0 = 256;
        AMP_TRY(tc_chain32_launch(p, c.st));
    }
    // chain 3b: conv_3 .. conv_6 on the local features + global max-pool (:100-106)
    {
        T32Params p = base;
        p.in_mode = 1; p.in_x = out + 256; p.in_ld = 320; p.in_k = 64; p.n_ops = 4;
        p.wblob = k.blob3b; p.wblob_bytes = kP32Blob3b; p.wstream = k.stream;
        p.op[0] = t32_op(64, 64, kP32W_c3, 0, enc_bn_offset(L_C3), 1);
        p.op[1] = t32_op(64, 128, kP32W_c4, 0, enc_bn_offset(L_C4), 1);
        p.op[2] = t32_op(128, 128, kP32W_c5, 0, enc_bn_offset(L_C5), 1);
        p.op[3] = t32_op(128, 256, 0, 0, enc_bn_offset(L_C6), 1); p.op[3].write_act = 0; p.op[3].pool = 1; p.op[3].w_stream = 1;
        p.pool = reinterpret_cast<unsigned int*>(G);
        AMP_TRY(tc_chain32_launch(p, c.st));
    }
    // repeat + cat (:109-110)
    return broadcast_rows(G, B, N, 256, out, 320, c.st);
}

int check_sizes(int64_t B, int64_t N, const char* who) {
    if (B < 1 || N < 1 || B > 65535 || B * N > (1LL << 31) / 320)
        return fail(AMP_E_BADARG, "%s: unsupported shape B=%lld N=%lld", who, (long long)B, (long long)N);
    return AMP_OK;
}

const char* kTnetNames[T_COUNT] = {
    "conv_1.weight", "conv_2.weight", "conv_3.weight",
    "bn_1.weight", "bn_1.bias", "bn_1.running_mean", "bn_1.running_var", "bn_1.num_batches_tracked",
    "bn_2.weight", "bn_2.bias", "bn_2.running_mean", "bn_2.running_var", "bn_2.num_batches_tracked",
    "bn_3.weight", "bn_3.bias", "bn_3.running_mean", "bn_3.running_var", "bn_3.num_batches_tracked",
    "bn_4.weight", "bn_4.bias", "bn_4.running_mean", "bn_4.running_var", "bn_4.num_batches_tracked",
    "bn_5.weight", "bn_5.bias", "bn_5.running_mean", "bn_5.running_var", "bn_5.num_batches_tracked",
    "fc_1.weight", "fc_2.weight", "fc_3.weight", "fc_3.bias"};
const char* kBnSuffix[BN_STRIDE] = {"weight", "bias", "running_mean", "running_var", "num_batches_tracked"};

}  // namespace
}  // namespace amp

extern "C" {

int amp_encoder_param_count(void) { return amp::E_COUNT; }

const char* amp_encoder_param_name(int i) {
    using namespace amp;
    static thread_local char buf[96];
    if (i < 0 || i >= E_COUNT) return nullptr;
    if (i < E_FT) { snprintf(buf, sizeof buf, "input_transform.%s", kTnetNames[i]); return buf; }
    if (i < E_CONV1) { snprintf(buf, sizeof buf, "feature_transform.%s", kTnetNames[i - E_FT]); return buf; }
    if (i < E_BN1) { snprintf(buf, sizeof buf, "conv_%d.weight", i - E_CONV1 + 1); return buf; }
    snprintf(buf, sizeof buf, "bn_%d.%s", (i - E_BN1) / BN_STRIDE + 1, kBnSuffix[(i - E_BN1) % BN_STRIDE]);
    return buf;
}

size_t amp_encoder_saved_bytes(int64_t B, int64_t N, int32_t training) {
    if (!training) return 0;
    amp::Arena a(nullptr, std::numeric_limits<size_t>::max());
    amp::enc_carve(a, B, N, true);
    return a.off + 256;
}

size_t amp_encoder_workspace_bytes(int64_t B, int64_t N, int32_t training) {
    size_t f = amp::fwd_ws_bytes(B, N, training != 0);
    if (training) {
        amp::Arena a(nullptr, std::numeric_limits<size_t>::max());
        amp::bwd_carve(a, nullptr, B, N);
        if (a.off + 256 > f) f = a.off + 256;
    }
    return f;
}

size_t amp_encoder_pack_bytes(void) { return amp::enc_pack_bytes(); }

int amp_encoder_fwd(const void* const* params, const float* x, int64_t B, int64_t N, int32_t training, int32_t precision,
                    float* out, float* feat_t, void* saved, size_t saved_bytes, void* workspace,
                    size_t workspace_bytes, void* pack_cache, size_t pack_bytes, int32_t pack_valid, void* stream) {
    using namespace amp;
    if (!params || !x || !out || !feat_t || !workspace) return fail(AMP_E_BADARG, "encoder_fwd: null pointer");
    AMP_TRY(check_sizes(B, N, "encoder_fwd"));
    if (precision != AMP_PREC_FP32 && precision != AMP_PREC_BF16 && precision != AMP_PREC_FP32_STRICT)
        return fail(AMP_E_BADARG, "encoder_fwd: unknown precision %d", precision);
    StrictScope strict(precision == AMP_PREC_FP32_STRICT);
    if (precision == AMP_PREC_FP32_STRICT) precision = AMP_PREC_FP32;
    if (training && precision != AMP_PREC_FP32)
        return fail(AMP_E_BADARG, "encoder_fwd: the bf16 tensor-core path is eval-only; training runs in AMP_PREC_FP32");
    if (training && B * N < 2) return fail(AMP_E_BADARG, "encoder_fwd: BatchNorm in training mode needs more than 1 row");
    if (training && B < 2) return fail(AMP_E_BADARG, "encoder_fwd: training mode needs B >= 2 (BatchNorm over the T-Net FC rows)");
    for (int i = 0; i < E_COUNT; ++i)
        if (!params[i]) return fail(AMP_E_BADARG, "encoder_fwd: parameter %d (%s) is null", i, amp_encoder_param_name(i));
    if (workspace_bytes < fwd_ws_bytes(B, N, training != 0)) return fail(AMP_E_WORKSPACE, "encoder_fwd: workspace too small");
    if (training && (!saved || saved_bytes < amp_encoder_saved_bytes(B, N, 1)))
        return fail(AMP_E_WORKSPACE, "encoder_fwd: saved-for-backward buffer too small");
    PdlScope pdl(training == 0);
    EncCtx c{};
    c.P = params; c.st = (cudaStream_t)stream; c.B = (int)B; c.N = (int)N; c.train = training != 0;
    Arena wa(workspace, workspace_bytes);
    if (precision == AMP_PREC_BF16) return encoder_fwd_bf16(c, x, out, feat_t, wa);
    if (!c.train && precision == AMP_PREC_FP32 && !path_disabled("tc_chain32"))
        return encoder_fwd_fp32_fused(c, x, out, feat_t, wa, pack_cache, pack_bytes, pack_valid);
    const size_t tiles = (size_t)pw_tiles(c.B, c.N);
    c.part_sum = wa.take<float>(tiles * 256); c.part_sq = wa.take<float>(tiles * 256);
    c.pmax = wa.take<unsigned long long>(B * 256); c.pmin = wa.take<unsigned long long>(B * 256);
    if (c.train) { Arena sa(saved, saved_bytes); c.S = enc_carve(sa, B, N, true); }
    else c.S = enc_carve(wa, B, N, false);
    EncSaved& S = c.S;
    if (!c.train) {
        BnDesc t[L_ENC_BN];
        for (int L = 0; L < L_ENC_BN; ++L) {
            const int pb = kEncBnParam[L], o = enc_bn_offset(L);
            t[L] = BnDesc{c.pf(pb + BN_W), c.pf(pb + BN_B), c.pf(pb + BN_RM), c.pf(pb + BN_RV), S.scale + o, S.shift + o,
                          kEncBnChannels[L]};
        }
        AMP_TRY(bn_fold_eval(t, L_ENC_BN, kBnEps, c.st));
    }
    const int Bi = c.B, Ni = c.N;
    // input T-Net on xyz (:83-84)
    AMP_TRY(tnet_fwd(c, E_IT, L_IT1, 3, x, 9, 3, -1, S.it_y1, S.it_y2, S.it_y3, S.it_pool, S.it_arg, S.it_f1, S.it_f2, S.T));
    // bmm + cat + conv_1 (:85-90) as per-cloud conv_1 weights
    AMP_TRY(fold_input_transform(c.pf(E_CONV1), S.T, Bi, S.W1eff, c.st));
    AMP_TRY(fwd_layer(c, x, 9, 9, -1, S.W1eff, 9, 576, 0, nullptr, S.c1, 64, 64, L_C1, nullptr, nullptr, Bi, Ni));
    AMP_TRY(fwd_layer(c, S.c1, 64, 64, L_C1, c.pf(E_CONV2), 64, 0, 0, nullptr, S.c2, 64, 64, L_C2, nullptr, nullptr, Bi, Ni));
    // feature T-Net (:94); its output is the second module output
    AMP_TRY(tnet_fwd(c, E_FT, L_FT1, 64, S.c2, 64, 64, L_C2, S.ft_y1, S.ft_y2, S.ft_y3, S.ft_pool, S.ft_arg, S.ft_f1, S.ft_f2, feat_t));
    // local features = x @ F (:96-97), written straight into out[:, :, 256:320]
    float* local = out + 256;
    AMP_TRY(fwd_layer(c, S.c2, 64, 64, L_C2, feat_t, 64, 4096, 1, nullptr, local, 320, 64, -1, nullptr, nullptr, Bi, Ni));
    // conv_3 .. conv_6 + global max-pool (:100-106)
    AMP_TRY(fwd_layer(c, local, 320, 64, -1, c.pf(E_CONV3), 64, 0, 0, nullptr, S.c3, 64, 64, L_C3, nullptr, nullptr, Bi, Ni));
    AMP_TRY(fwd_layer(c, S.c3, 64, 64, L_C3, c.pf(E_CONV4), 64, 0, 0, nullptr, S.c4, 128, 128, L_C4, nullptr, nullptr, Bi, Ni));
    AMP_TRY(fwd_layer(c, S.c4, 128, 128, L_C4, c.pf(E_CONV5), 128, 0, 0, nullptr, S.c5, 128, 128, L_C5, nullptr, nullptr, Bi, Ni));
    AMP_TRY(fwd_layer(c, S.c5, 128, 128, L_C5, c.pf(E_CONV6), 128, 0, 0, nullptr, S.c6, 256, 256, L_C6, S.G, S.g_arg, Bi, Ni));
    // repeat + cat (:109-110)
    return broadcast_rows(S.G, Bi, Ni, 256, out, 320, c.st);
}

int amp_encoder_bwd(const void* const* params, void* const* grads, const float* x, const float* out,
                    const float* feat_t, const float* d_out, const float* d_feat_t, int64_t B, int64_t N,
                    void* saved, size_t saved_bytes, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace amp;
    if (!params || !grads || !x || !out || !feat_t || !d_out || !saved || !workspace)
        return fail(AMP_E_BADARG, "encoder_bwd: null pointer");
    AMP_TRY(check_sizes(B, N, "encoder_bwd"));
    if (saved_bytes < amp_encoder_saved_bytes(B, N, 1)) return fail(AMP_E_WORKSPACE, "encoder_bwd: saved buffer too small");
    if (workspace_bytes < amp_encoder_workspace_bytes(B, N, 1)) return fail(AMP_E_WORKSPACE, "encoder_bwd: workspace too small");
    for (int i = 0; i < E_COUNT; ++i) {
        const bool is_buf = i >= E_BN1 ? (i - E_BN1) % BN_STRIDE >= BN_RM
                          : i < E_CONV1 ? ((i % T_COUNT) >= T_BN1 && (i % T_COUNT) < T_FC1 && ((i % T_COUNT) - T_BN1) % BN_STRIDE >= BN_RM)
                                        : false;
        if (!params[i] || (!is_buf && !grads[i]))
            return fail(AMP_E_BADARG, "encoder_bwd: parameter / gradient %d (%s) is null", i, amp_encoder_param_name(i));
    }
    EncCtx c{};
    c.P = params; c.Gd = grads; c.st = (cudaStream_t)stream; c.B = (int)B; c.N = (int)N; c.train = true;
    Arena sa(saved, saved_bytes);
    c.S = enc_carve(sa, B, N, true);
    Arena wa(workspace, workspace_bytes);
    BwdWs w = bwd_carve(wa, &c, B, N);
    EncSaved& S = c.S;
    const int Bi = c.B, Ni = c.N;
    const size_t M = (size_t)B * N;
    const float* local = out + 256;
    WgDeferScope defer(w.wg_pool, w.wg_pool_floats, c.st);     // parameter-gradient reductions: one launch at the end

    // out = [G broadcast | local]: dG = column sums of d_out[:, :, :256]
    AMP_TRY(colsum_rows(d_out, 320, Bi, Ni, 256, w.dG, w.colsum, c.st));
    AMP_TRY(bwd_pool(c, w.dG, S.g_arg, S.c6, L_C6, w.gA));
    AMP_TRY(bwd_step(c, w.gA, 256, L_C6, S.c6, S.c5, 128, 128, L_C5, c.pf(E_CONV6), c.gf(E_CONV6), nullptr, w.gB, 128, 0, false, Bi, Ni));
    AMP_TRY(bwd_step(c, w.gB, 128, L_C5, S.c5, S.c4, 128, 128, L_C4, c.pf(E_CONV5), c.gf(E_CONV5), nullptr, w.gA, 128, 0, false, Bi, Ni));
    AMP_TRY(bwd_step(c, w.gA, 128, L_C4, S.c4, S.c3, 64, 64, L_C3, c.pf(E_CONV4), c.gf(E_CONV4), nullptr, w.gB, 64, 0, false, Bi, Ni));
    // d local = d_out[:, :, 256:] + conv_3 input gradient
    AMP_CUDA(cudaMemcpy2DAsync(w.gA, 64 * sizeof(float), d_out + 256, 320 * sizeof(float), 64 * sizeof(float), M,
                               cudaMemcpyDeviceToDevice, c.st));
    AMP_TRY(bwd_step(c, w.gB, 64, L_C3, S.c3, local, 320, 64, -1, c.pf(E_CONV3), c.gf(E_CONV3), nullptr, w.gA, 64, 1, false, Bi, Ni));
    // local = relu(bn_2(c2)) @ F[b]:  dF[b] = a2^T dlocal (+ d_feat_t),  d a2 = dlocal @ F[b]^T
    if (d_feat_t) AMP_CUDA(cudaMemcpyAsync(w.dF, d_feat_t, sizeof(float) * B * 4096, cudaMemcpyDeviceToDevice, c.st));
    else AMP_CUDA(cudaMemsetAsync(w.dF, 0, sizeof(float) * B * 4096, c.st));
    {
        const int o2 = enc_bn_offset(L_C2);
        WgParams g{};
        g.dY = w.gA; g.lddy = 64; g.Nout = 64;
        g.A = S.c2; g.lda = 64; g.K = 64; g.a_a = S.scale + o2; g.a_b = c.pf(E_BN2 + BN_B); g.a_m = S.mean + o2; g.a_relu = 1;
        g.n_clouds = Bi; g.rows_per_cloud = Ni; g.per_cloud = 1; g.w_kn = 1;
        g.dW = w.dF; g.ldw = 64; g.w_cloud_stride = 4096; g.accumulate = 1;
        g.partials = c.wg; g.partial_floats = c.wg_floats;
        AMP_TRY(wgrad(g, c.st));
        PwParams p{};
        p.X = w.gA; p.ldx = 64; p.K = 64;
        p.W = feat_t; p.ldw = 64; p.w_cloud_stride = 4096; p.w_kn = 0; p.n_groups = 1;
        p.Y = w.g64; p.ldy = 64; p.n_clouds = Bi; p.rows_per_cloud = Ni; p.Nout = 64;
        AMP_TRY(pw_linear(p, c.st));
    }
    // feature T-Net; closes the fan-in of bn_2 (its conv_1 input gradient accumulates onto g64)
    AMP_TRY(tnet_bwd(c, w, E_FT, L_FT1, 64, w.dF, S.ft_y1, S.ft_y2, S.ft_y3, S.ft_pool, S.ft_arg, S.ft_f1, S.ft_f2, S.c2, 64, 64,
                     L_C2, w.g64, 64, w.gA, w.gB));
    AMP_TRY(bwd_step(c, w.g64, 64, L_C2, S.c2, S.c1, 64, 64, L_C1, c.pf(E_CONV2), c.gf(E_CONV2), nullptr, w.gA, 64, 0, false, Bi, Ni));
    // conv_1 with per-cloud folded weights: dW1eff[b] = dy1[b]^T x[b], then unfold into conv_1.weight and dT
    {
        const int o1 = enc_bn_offset(L_C1);
        WgParams g{};
        g.dY = w.gA; g.lddy = 64; g.Nout = 64; g.y_a = c.k1 + o1; g.y_b = c.k3 + o1; g.y_c = c.k2 + o1; g.y_m = S.mean + o1; g.Y2 = S.c1;
        g.A = x; g.lda = 9; g.K = 9;
        g.n_clouds = Bi; g.rows_per_cloud = Ni; g.per_cloud = 1;
        g.dW = w.dW1eff; g.ldw = 9; g.w_cloud_stride = 576;
        g.partials = c.wg; g.partial_floats = c.wg_floats;
        AMP_TRY(wgrad(g, c.st));
        AMP_TRY(fold_input_transform_bwd(w.dW1eff, c.pf(E_CONV1), S.T, Bi, c.gf(E_CONV1), w.dT, c.st));
    }
    // input T-Net (no input gradient: x is data)
    AMP_TRY(tnet_bwd(c, w, E_IT, L_IT1, 3, w.dT, S.it_y1, S.it_y2, S.it_y3, S.it_pool, S.it_arg, S.it_f1, S.it_f2, x, 9, 3, -1,
                     nullptr, 0, w.gB, w.gA));
    return defer.flush();
}

}  // extern "C"
