// Attention + segmentation head: forward and backward host orchestration behind amp_seg_fwd / amp_seg_bwd.
//
// Replaces SegmentationWithAttention.forward (pointNet/model/pointnetAtt.py:176-209): positional encoding
// (:183-185), nn.MultiheadAttention over the W block tokens of each window (:187-190), the repeat/cat loop
// (:192-200) and conv_2/bn_2/relu/dropout, conv_3/bn_3/relu/dropout, conv_4 (:203-207); and the autograd
// backward through them (pointNet/self-attention/train_pointnet-attention.py:467).
//
// The [B, sumN, 320] concatenation is never built: conv_2 is split as
//   conv_2(cat(local, g_w)) = W2[:, :64] local + (W2[:, 64:] g_w + b2)
// and the second term is a per-(window, block) bias [B, W, 128] computed once per token.
#include <limits>

#include "nn_layout.cuh"
#include "tc_chain.cuh"
#include "tc_chain32.cuh"

namespace amp {
namespace {

constexpr float kBnEps = 1e-5f, kBnMomentum = 0.1f;
constexpr int kSegBn2 = 0, kSegBn3 = 128, kSegBnTotal = 192;   // offsets into the scale/shift tables
// bf16 tensor-core head: conv_2 local half (128 x 64), conv_3 (64 x 128), conv_4 (<= 32 x 64) packed; biases of conv_3 / conv_4
constexpr int kSegTcBlob = 16384 + (16384 + 1024) + (4096 + 512);

// fp32-class fused head (tc_chain32): conv_2 local half (128 x 64), conv_3 (64 x 128), conv_4 (<= 32 x 64), hi + lo split;
// bias table = [scale3 * b3 + shift3 (64) | b4 (Cp)]; kept in a caller-owned pack cache between calls
constexpr int kSeg32W2 = 0, kSeg32W3 = 128 * 64 * 4, kSeg32W4 = kSeg32W3 + 64 * 128 * 4;
struct SegPack { float *scale, *shift, *bias, *wc, *bc; unsigned char* blob; };   // wc / bc: seg_fold_out (E = 256, hid = 128)
SegPack seg_pack_carve(Arena& a, int Cp) {
    SegPack k{};
    k.scale = a.take<float>(kSegBnTotal); k.shift = a.take<float>(kSegBnTotal); k.bias = a.take<float>(64 + Cp);
    k.wc = a.take<float>(128 * 256); k.bc = a.take<float>(128);
    k.blob = a.take<unsigned char>(kSeg32W4 + Cp * 64 * 4);
    return k;
}
size_t seg_pack_bytes(int num_classes) {
    Arena a(nullptr, std::numeric_limits<size_t>::max());
    seg_pack_carve(a, (num_classes + 15) / 16 * 16);
    return a.off + 256;
}

#define AMP_TRY(expr) do { int rc_ = (expr); if (rc_ != AMP_OK) return rc_; } while (0)
#define AMP_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return fail(AMP_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); } while (0)

struct SegSaved {
    float *tokens, *h_pre, *qkv, *probs, *attn_o, *g_w, *cb;      // per token
    float *y2, *y3;                                               // per point (raw in training, final in eval)
    float *scale, *shift, *mean, *invstd;                         // [kSegBnTotal]
};

SegSaved seg_carve(Arena& a, long long B, long long W, long long R, int E, int heads, int hid) {
    SegSaved s{};
    const size_t T = (size_t)B * W, M = (size_t)B * R;
    s.tokens = a.take<float>(T * E); s.h_pre = a.take<float>(T * 16); s.qkv = a.take<float>(T * 3 * E);
    s.probs = a.take<float>((size_t)B * heads * W * W); s.attn_o = a.take<float>(T * E); s.g_w = a.take<float>(T * E);
    s.cb = a.take<float>(T * hid);
    s.y2 = a.take<float>(M * hid); s.y3 = a.take<float>(M * 64);
    s.scale = a.take<float>(kSegBnTotal); s.shift = a.take<float>(kSegBnTotal);
    s.mean = a.take<float>(kSegBnTotal); s.invstd = a.take<float>(kSegBnTotal);
    return s;
}

struct SegWs {
    unsigned char* tc_blob;                                     // bf16 tensor-core path: packed head weights + bias K groups
    unsigned char* pack32;                                      // fp32-class fused path without a caller-owned pack cache
    float *part_sum, *part_sq, *k1, *k2, *k3, *wg; size_t wg_floats;
    float* wg_pool; size_t wg_pool_floats;
    void* split_dump;                                           // dy' of the layer being differentiated, split bf16 (PwParams::split_dump)
    float *dz3, *dz2, *dcb, *dg_w, *dattn_o, *dqkv, *dtokens, *dpre;
};

constexpr int kMinGroupSlab = 64;
size_t seg_wg_floats(long long B, long long W, long long R, int E, int hid) {
    size_t m = wgrad_workspace_floats((int)B, (int)R, hid, 64, R < kMinGroupSlab ? (int)R : kMinGroupSlab);
    size_t o = wgrad_workspace_floats((int)B, (int)R, 64, hid);
    if (o > m) m = o;
    o = wgrad_workspace_floats(1, (int)(B * W), 3 * E, E);
    if (o > m) m = o;
    return m;
}

SegWs seg_ws_carve(Arena& a, long long B, long long W, long long R, int E, int hid, bool backward) {
    SegWs w{};
    const size_t T = (size_t)B * W, M = (size_t)B * R;
    const size_t tiles = (size_t)pw_tiles((int)B, (int)R);
    w.part_sum = a.take<float>(tiles * 128); w.part_sq = a.take<float>(tiles * 128);
    if (!backward) {
        w.tc_blob = a.take<unsigned char>(kSegTcBlob);
        w.pack32 = a.take<unsigned char>(seg_pack_bytes(32));          // used when the caller passes no pack cache
        return w;
    }
    w.k1 = a.take<float>(kSegBnTotal); w.k2 = a.take<float>(kSegBnTotal); w.k3 = a.take<float>(kSegBnTotal);
    w.wg_floats = seg_wg_floats(B, W, R, E, hid);
    w.wg = a.take<float>(w.wg_floats);
    w.wg_pool_floats = 2 * w.wg_floats + 64 * 16;              // deferred parameter-gradient reductions (WgDeferScope)
    w.wg_pool = a.take<float>(w.wg_pool_floats);
    w.split_dump = a.take<unsigned char>(tiles * (size_t)(2 * 16 * 128 * 16));          // up to 128 channels
    w.dz3 = a.take<float>(M * 64); w.dz2 = a.take<float>(M * hid);
    w.dcb = a.take<float>(T * hid); w.dg_w = a.take<float>(T * E); w.dattn_o = a.take<float>(T * E);
    w.dqkv = a.take<float>(T * 3 * E); w.dtokens = a.take<float>(T * E); w.dpre = a.take<float>(T * 16);
    return w;
}

struct SegShape { int64_t B, W, R; int E, heads, C, hid; };

int seg_check(const SegShape& s, const int32_t* np_cluster, const char* who) {
    if (s.B < 1 || s.W < 1 || s.W > 1024 || s.R < 1 || s.B > 65535 || s.B * s.R > (1LL << 31) / 320)
        return fail(AMP_E_BADARG, "%s: unsupported shape B=%lld W=%lld rows=%lld", who, (long long)s.B, (long long)s.W, (long long)s.R);
    if (s.E != 256 || s.heads < 1 || s.E % s.heads || s.hid != s.E / 2 || s.C < 1 || s.C > 64)
        return fail(AMP_E_BADARG, "%s: unsupported head E=%d heads=%d classes=%d", who, s.E, s.heads, s.C);
    long long sum = 0;
    for (int i = 0; i < s.W; ++i) {
        if (np_cluster[i] < 1) return fail(AMP_E_BADARG, "%s: empty block %d", who, i);
        sum += np_cluster[i];
    }
    if (sum != s.R) return fail(AMP_E_BADARG, "%s: sum(np_cluster)=%lld != rows=%lld", who, sum, (long long)s.R);
    return AMP_OK;
}

const char* kSegNames[S_COUNT] = {
    "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "attention.in_proj_weight", "attention.in_proj_bias",
    "attention.out_proj.weight", "attention.out_proj.bias", "conv_2.weight", "conv_2.bias", "conv_3.weight", "conv_3.bias",
    "conv_4.weight", "conv_4.bias",
    "bn_2.weight", "bn_2.bias", "bn_2.running_mean", "bn_2.running_var", "bn_2.num_batches_tracked",
    "bn_3.weight", "bn_3.bias", "bn_3.running_mean", "bn_3.running_var", "bn_3.num_batches_tracked"};

inline const float* pf(const void* const* P, int i) { return reinterpret_cast<const float*>(P[i]); }
inline float* gf(void* const* G, int i) { return reinterpret_cast<float*>(G[i]); }

}  // namespace
}  // namespace amp

extern "C" {

size_t amp_seg_pack_bytes(int32_t num_classes) { return amp::seg_pack_bytes(num_classes); }
int amp_seg_param_count(void) { return amp::S_COUNT; }
const char* amp_seg_param_name(int i) { return (i < 0 || i >= amp::S_COUNT) ? nullptr : amp::kSegNames[i]; }

size_t amp_seg_saved_bytes(int64_t B, int64_t W, int64_t rows, int32_t embed_dim, int32_t heads) {
    amp::Arena a(nullptr, std::numeric_limits<size_t>::max());
    amp::seg_carve(a, B, W, rows, embed_dim, heads, embed_dim / 2);
    return a.off + 256;
}

size_t amp_seg_workspace_bytes(int64_t B, int64_t W, int64_t rows, int32_t embed_dim, int32_t training) {
    amp::Arena a(nullptr, std::numeric_limits<size_t>::max());
    amp::seg_ws_carve(a, B, W, rows, embed_dim, embed_dim / 2, training != 0);
    return a.off + 256;
}

int amp_seg_fwd(const void* const* params, const float* gl_feats, int64_t gl_ld, const float* lo_feats, int64_t lo_ld,
                const float* centroids,
                const int32_t* np_cluster, const int32_t* group_rows, const uint8_t* key_padding_mask, int64_t B,
                int64_t W, int64_t rows, int32_t embed_dim, int32_t heads, int32_t num_classes, int32_t training,
                int32_t precision, float dropout_p, uint64_t seed, float* logits, void* saved, size_t saved_bytes,
                void* workspace, size_t workspace_bytes, void* pack_cache, size_t pack_bytes, int32_t pack_valid, void* stream) {
    using namespace amp;
    if (precision != AMP_PREC_FP32 && precision != AMP_PREC_BF16 && precision != AMP_PREC_FP32_STRICT)
        return fail(AMP_E_BADARG, "seg_fwd: unknown precision %d", precision);
    StrictScope strict(precision == AMP_PREC_FP32_STRICT);
    if (precision == AMP_PREC_FP32_STRICT) precision = AMP_PREC_FP32;
    if (precision == AMP_PREC_BF16 && training)
        return fail(AMP_E_BADARG, "seg_fwd: the bf16 tensor-core path is eval-only; training runs in AMP_PREC_FP32");
    if (precision == AMP_PREC_BF16 && num_classes > 32) return fail(AMP_E_BADARG, "seg_fwd: the bf16 head supports up to 32 classes");
    if (!params || !gl_feats || !lo_feats || !centroids || !np_cluster || !group_rows || !logits || !saved || !workspace)
        return fail(AMP_E_BADARG, "seg_fwd: null pointer");
    if (gl_ld < embed_dim || lo_ld < 64 || (lo_ld & 3) || (reinterpret_cast<uintptr_t>(lo_feats) & 15))
        return fail(AMP_E_BADARG, "seg_fwd: gl_ld >= embed_dim, lo_ld >= 64 and a multiple of 4, lo_feats 16-byte aligned");
    const SegShape sh{B, W, rows, embed_dim, heads, num_classes, embed_dim / 2};
    AMP_TRY(seg_check(sh, np_cluster, "seg_fwd"));
    if (dropout_p < 0.f || dropout_p >= 1.f) return fail(AMP_E_BADARG, "seg_fwd: dropout p must be in [0, 1)");
    for (int i = 0; i < S_COUNT; ++i)
        if (!params[i]) return fail(AMP_E_BADARG, "seg_fwd: parameter %d (%s) is null", i, kSegNames[i]);
    if (saved_bytes < amp_seg_saved_bytes(B, W, rows, embed_dim, heads)) return fail(AMP_E_WORKSPACE, "seg_fwd: saved buffer too small");
    if (workspace_bytes < amp_seg_workspace_bytes(B, W, rows, embed_dim, 0)) return fail(AMP_E_WORKSPACE, "seg_fwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const bool train = training != 0;
    PdlScope pdl(!train);
    const float dp = train ? dropout_p : 0.f;
    const int E = embed_dim, hid = sh.hid, Bi = (int)B, Wi = (int)W, Ri = (int)rows, T = Bi * Wi;
    Arena sa(saved, saved_bytes);
    SegSaved S = seg_carve(sa, B, W, rows, E, heads, hid);
    Arena wa(workspace, workspace_bytes);
    SegWs ws = seg_ws_carve(wa, B, W, rows, E, hid, false);

    const bool fused32 = !train && precision == AMP_PREC_FP32 && num_classes <= 32 && hid == 128 && !path_disabled("tc_chain32");
    const int Cp = (num_classes + 15) / 16 * 16;
    SegPack k{};
    bool tail_done = false;
    if (fused32) {
        // fused head on the tensor cores at fp32-class accuracy: bn_2 / bn_3 folded into the packed weights and the biases;
        // the per-block bias becomes cb' = scale2 * (W2[:, 64:] g_w + b2) + shift2
        if (pack_cache) {
            if (pack_bytes < seg_pack_bytes(num_classes)) return fail(AMP_E_WORKSPACE, "seg_fwd: pack cache too small");
            Arena pa(pack_cache, pack_bytes);
            k = seg_pack_carve(pa, Cp);
        } else {
            Arena pa(ws.pack32, seg_pack_bytes(32));
            k = seg_pack_carve(pa, Cp);
            pack_valid = 0;
        }
        if (!pack_valid) {
            BnDesc t[2] = {
                {pf(params, S_BN2 + BN_W), pf(params, S_BN2 + BN_B), pf(params, S_BN2 + BN_RM), pf(params, S_BN2 + BN_RV), k.scale + kSegBn2, k.shift + kSegBn2, hid},
                {pf(params, S_BN3 + BN_W), pf(params, S_BN3 + BN_B), pf(params, S_BN3 + BN_RM), pf(params, S_BN3 + BN_RV), k.scale + kSegBn3, k.shift + kSegBn3, 64}};
            AMP_TRY(bn_fold_eval(t, 2, kBnEps, st));
            T32PackTable pt{}; pt.n = 3; pt.n_clouds = 1;
            pt.job[0] = T32PackJob{pf(params, S_C2W), 64 + E, 0, k.scale + kSegBn2, hid, 64, hid, 64, 0, 0, kSeg32W2, 0};
            pt.job[1] = T32PackJob{pf(params, S_C3W), hid, 0, k.scale + kSegBn3, 64, hid, 64, hid, 0, 0, kSeg32W3, 0};
            pt.job[2] = T32PackJob{pf(params, S_C4W), 64, 0, nullptr, num_classes, 64, Cp, 64, 0, 0, kSeg32W4, 0};
            AMP_TRY(t32_pack_weights(pt, k.blob, st));
            AMP_TRY(t32_affine_bias(pf(params, S_C3B), k.scale + kSegBn3, k.shift + kSegBn3, 64, 64, k.bias, st));
            AMP_TRY(t32_affine_bias(pf(params, S_C4B), nullptr, nullptr, num_classes, Cp, k.bias + 64, st));
            if (E == 256)      // whatever W and the debug switches of THIS call: the cache outlives it
                AMP_TRY(seg_fold_out(pf(params, S_C2W) + 64, 64 + E, pf(params, S_C2B), pf(params, S_OUTW), pf(params, S_OUTB), E, hid, k.wc, k.bc, st));
        }
        // positional encoding, attention and the per-block bias in one launch (nn_seg_tail.cu); the barrier words live in the
        // h_pre slot of the saved buffer, which only the training backward reads
        const int rc = seg_tail_eval(gl_feats, gl_ld, centroids, pf(params, S_FC1W), pf(params, S_FC1B), pf(params, S_FC2W), pf(params, S_FC2B),
                                     pf(params, S_INW), pf(params, S_INB), k.wc, k.bc, k.scale + kSegBn2, k.shift + kSegBn2, key_padding_mask,
                                     Bi, Wi, E, heads, hid, S.qkv, S.attn_o, S.cb, reinterpret_cast<unsigned int*>(S.h_pre), st);
        if (rc < 0) return rc;
        tail_done = rc == 1;
    }
    if (!tail_done) {
    // positional encoding + token-major layout (:183-185)
    AMP_TRY(posenc_add(gl_feats, gl_ld, centroids, pf(params, S_FC1W), pf(params, S_FC1B), pf(params, S_FC2W), pf(params, S_FC2B), Bi, Wi, E,
                       S.tokens, S.h_pre, st));
    // nn.MultiheadAttention (:187-190): in_proj, per-head softmax(QK^T)V, out_proj
    {
        PwParams p{};
        p.X = S.tokens; p.ldx = E; p.K = E; p.W = pf(params, S_INW); p.ldw = E; p.bias = pf(params, S_INB); p.n_groups = 1;
        p.Y = S.qkv; p.ldy = 3 * E; p.n_clouds = 1; p.rows_per_cloud = T; p.Nout = 3 * E;
        AMP_TRY(pw_linear(p, st));
    }
    AMP_TRY(attention_core(S.qkv, key_padding_mask, dp, seed + 1, Bi, Wi, E, heads, S.attn_o, S.probs, st));
    {
        PwParams p{};
        p.X = S.attn_o; p.ldx = E; p.K = E; p.W = pf(params, S_OUTW); p.ldw = E; p.bias = pf(params, S_OUTB); p.n_groups = 1;
        p.Y = S.g_w; p.ldy = E; p.n_clouds = 1; p.rows_per_cloud = T; p.Nout = E;
        AMP_TRY(pw_linear(p, st));
    }
    }
    // per-block bias of conv_2: cb[b, w, :] = W2[:, 64:] g_w[b, w] + b2   (the repeat/cat of :192-200 folded away)
    {
        PwParams p{};
        p.X = S.g_w; p.ldx = E; p.K = E; p.W = pf(params, S_C2W) + 64; p.ldw = 64 + E; p.bias = pf(params, S_C2B); p.n_groups = 1;
        p.Y = S.cb; p.ldy = hid; p.n_clouds = 1; p.rows_per_cloud = T; p.Nout = hid;
        if (precision != AMP_PREC_BF16 && !fused32) AMP_TRY(pw_linear(p, st));
    }
    if (fused32) {
        {
            PwParams p{};
            p.X = S.g_w; p.ldx = E; p.K = E; p.W = pf(params, S_C2W) + 64; p.ldw = 64 + E; p.bias = pf(params, S_C2B); p.n_groups = 1;
            p.out_scale = k.scale + kSegBn2; p.out_shift = k.shift + kSegBn2;
            p.Y = S.cb; p.ldy = hid; p.n_clouds = 1; p.rows_per_cloud = T; p.Nout = hid;
            if (!tail_done) AMP_TRY(pw_linear(p, st));
        }
        T32Params p{};
        p.n_ops = 3;
        p.op[0] = T32Op{64, hid, kSeg32W2, 0, 0, -1, 1, 1, 1, 0, 0, 0};
        p.op[1] = T32Op{hid, 64, kSeg32W3, 0, 0, 0, 1, 0, 1, 0, 0, 0};
        p.op[2] = T32Op{64, Cp, kSeg32W4, 0, 0, 64, 0, 0, 0, 0, 0, 1};
        p.in_mode = 1; p.in_x = lo_feats; p.in_ld = lo_ld; p.in_k = 64;
        p.wblob = k.blob; p.wblob_bytes = kSeg32W4 + Cp * 64 * 4;
        p.bias = k.bias; p.n_bias = 64 + Cp;
        p.gbias = S.cb; p.group_rows = group_rows; p.n_groups = Wi;
        p.logits = logits; p.n_classes = num_classes;
        p.n_clouds = Bi; p.rows_per_cloud = Ri;
        return tc_chain32_launch(p, st);
    }
    if (precision == AMP_PREC_BF16) {
        // fused head on the tensor cores: bn_2 / bn_3 folded into the packed weights and the biases; the per-block bias
        // becomes cb' = scale2 * (W2[:, 64:] g_w + b2) + shift2
        BnDesc t[2] = {
            {pf(params, S_BN2 + BN_W), pf(params, S_BN2 + BN_B), pf(params, S_BN2 + BN_RM), pf(params, S_BN2 + BN_RV), S.scale + kSegBn2, S.shift + kSegBn2, hid},
            {pf(params, S_BN3 + BN_W), pf(params, S_BN3 + BN_B), pf(params, S_BN3 + BN_RM), pf(params, S_BN3 + BN_RV), S.scale + kSegBn3, S.shift + kSegBn3, 64}};
        AMP_TRY(bn_fold_eval(t, 2, kBnEps, st));
        {
            PwParams p{};
            p.X = S.g_w; p.ldx = E; p.K = E; p.W = pf(params, S_C2W) + 64; p.ldw = 64 + E; p.bias = pf(params, S_C2B); p.n_groups = 1;
            p.out_scale = S.scale + kSegBn2; p.out_shift = S.shift + kSegBn2;
            p.Y = S.cb; p.ldy = hid; p.n_clouds = 1; p.rows_per_cloud = T; p.Nout = hid;
            AMP_TRY(pw_linear(p, st));
        }
        const int Cp = (num_classes + 15) / 16 * 16;
        const int o_w2 = 0, o_w3 = 16384, o_b3 = o_w3 + 16384, o_w4 = o_b3 + 1024, o_b4 = o_w4 + 4096;
        TcPackTable pt{};
        pt.n = 3; pt.n_clouds = 1;
        pt.job[0] = TcPackJob{pf(params, S_C2W), 64 + E, 0, S.scale + kSegBn2, nullptr, nullptr, nullptr, -1, -1, hid, 64, hid, 64, 0, 0, o_w2, 0};
        pt.job[1] = TcPackJob{pf(params, S_C3W), hid, 0, S.scale + kSegBn3, pf(params, S_C3B), S.scale + kSegBn3, S.shift + kSegBn3, -1, o_b3,
                              64, hid, 64, hid, 0, 0, o_w3, 0};
        pt.job[2] = TcPackJob{pf(params, S_C4W), 64, 0, nullptr, pf(params, S_C4B), nullptr, nullptr, -1, o_b4, num_classes, 64, Cp, 64, 0, 0, o_w4, 0};
        AMP_TRY(tc_pack_weights(pt, ws.tc_blob, st));
        TcChainParams p{};
        p.n_ops = 3;
        p.op[0] = TcOp{64, hid, o_w2, 0, -1, 1, 1, 1, 0, 0, 0};
        p.op[1] = TcOp{hid, 64, o_w3, 0, o_b3, 1, 0, 1, 0, 0, 0};
        p.op[2] = TcOp{64, Cp, o_w4, 0, o_b4, 0, 0, 0, 0, 0, 1};
        p.in_mode = 1; p.in_x = lo_feats; p.in_ld = lo_ld; p.in_k = 64;
        p.wblob = ws.tc_blob; p.wblob_bytes = o_b4 + Cp * 16;
        p.gbias = S.cb; p.group_rows = group_rows; p.n_groups = Wi;
        p.logits = logits; p.n_classes = num_classes;
        p.n_clouds = Bi; p.rows_per_cloud = Ri;
        return tc_chain_launch(p, st);
    }
    if (!train) {
        BnDesc t[2] = {
            {pf(params, S_BN2 + BN_W), pf(params, S_BN2 + BN_B), pf(params, S_BN2 + BN_RM), pf(params, S_BN2 + BN_RV), S.scale + kSegBn2, S.shift + kSegBn2, hid},
            {pf(params, S_BN3 + BN_W), pf(params, S_BN3 + BN_B), pf(params, S_BN3 + BN_RM), pf(params, S_BN3 + BN_RV), S.scale + kSegBn3, S.shift + kSegBn3, 64}};
        AMP_TRY(bn_fold_eval(t, 2, kBnEps, st));
    }
    auto finalize = [&](int bn, int off, int C) {
        return bn_finalize_train(ws.part_sum, ws.part_sq, Bi, Ri, C, pf(params, bn + BN_W),
                                 const_cast<float*>(pf(params, bn + BN_RM)), const_cast<float*>(pf(params, bn + BN_RV)),
                                 reinterpret_cast<long long*>(const_cast<void*>(params[bn + BN_NBT])), kBnMomentum, kBnEps,
                                 S.scale + off, S.mean + off, S.invstd + off, st);
    };
    // conv_2 on the local half + per-block bias, bn_2, relu (:203)
    {
        PwParams p{};
        p.X = lo_feats; p.ldx = lo_ld; p.K = 64; p.W = pf(params, S_C2W); p.ldw = 64 + E;
        p.bias = S.cb; p.bias_group_stride = hid; p.group_rows = group_rows; p.n_groups = Wi;
        p.groups_tile_aligned = 1;
        for (int i = 0; i < Wi; ++i) if (np_cluster[i] % 128) p.groups_tile_aligned = 0;
        p.fp16_split = train ? 1 : 0;
        p.Y = S.y2; p.ldy = hid; p.n_clouds = Bi; p.rows_per_cloud = Ri; p.Nout = hid;
        if (train) { p.part_sum = ws.part_sum; p.part_sq = ws.part_sq; }
        else { p.out_scale = S.scale + kSegBn2; p.out_shift = S.shift + kSegBn2; p.out_relu = 1; }
        AMP_TRY(pw_linear(p, st));
        if (train) AMP_TRY(finalize(S_BN2, kSegBn2, hid));
    }
    // dropout, conv_3, bn_3, relu (:204-205)
    {
        PwParams p{};
        p.X = S.y2; p.ldx = hid; p.K = hid;
        if (train) { p.in_a = S.scale + kSegBn2; p.in_b = pf(params, S_BN2 + BN_B); p.in_m = S.mean + kSegBn2; p.in_relu = 1; p.in_drop_p = dp; p.drop_off = dropout_offset(); p.in_drop_seed = seed + 2; }
        p.W = pf(params, S_C3W); p.ldw = hid; p.bias = pf(params, S_C3B); p.n_groups = 1;
        p.fp16_split = train ? 1 : 0;
        p.Y = S.y3; p.ldy = 64; p.n_clouds = Bi; p.rows_per_cloud = Ri; p.Nout = 64;
        if (train) { p.part_sum = ws.part_sum; p.part_sq = ws.part_sq; }
        else { p.out_scale = S.scale + kSegBn3; p.out_shift = S.shift + kSegBn3; p.out_relu = 1; }
        AMP_TRY(pw_linear(p, st));
        if (train) AMP_TRY(finalize(S_BN3, kSegBn3, 64));
    }
    // dropout, conv_4 -> logits [B, C, rows] (:206-207)
    {
        PwParams p{};
        p.X = S.y3; p.ldx = 64; p.K = 64;
        if (train) { p.in_a = S.scale + kSegBn3; p.in_b = pf(params, S_BN3 + BN_B); p.in_m = S.mean + kSegBn3; p.in_relu = 1; p.in_drop_p = dp; p.drop_off = dropout_offset(); p.in_drop_seed = seed + 3; }
        p.W = pf(params, S_C4W); p.ldw = 64; p.bias = pf(params, S_C4B); p.n_groups = 1;
        p.Y = logits; p.y_transposed = 1; p.n_clouds = Bi; p.rows_per_cloud = Ri; p.Nout = num_classes;
        AMP_TRY(pw_linear(p, st));
    }
    return AMP_OK;
}

int amp_seg_bwd(const void* const* params, void* const* grads, const float* lo_feats, int64_t lo_ld, const float* centroids,
                const int32_t* np_cluster, const int32_t* group_rows, const float* d_logits, int64_t B, int64_t W,
                int64_t rows, int32_t embed_dim, int32_t heads, int32_t num_classes, float dropout_p, uint64_t seed,
                float* d_gl_feats, float* d_lo_feats, void* saved, size_t saved_bytes, void* workspace,
                size_t workspace_bytes, void* stream) {
    using namespace amp;
    if (!params || !grads || !lo_feats || !centroids || !np_cluster || !group_rows || !d_logits || !d_gl_feats || !d_lo_feats ||
        !saved || !workspace)
        return fail(AMP_E_BADARG, "seg_bwd: null pointer");
    if (lo_ld < 64 || (lo_ld & 3) || (reinterpret_cast<uintptr_t>(lo_feats) & 15))
        return fail(AMP_E_BADARG, "seg_bwd: lo_ld >= 64 and a multiple of 4, lo_feats 16-byte aligned");
    const SegShape sh{B, W, rows, embed_dim, heads, num_classes, embed_dim / 2};
    AMP_TRY(seg_check(sh, np_cluster, "seg_bwd"));
    if (W > 64) return fail(AMP_E_BADARG, "seg_bwd: more than 64 blocks per window");
    const int gslab = wgrad_group_slab(np_cluster, (int)W);
    if (gslab < kMinGroupSlab && gslab != rows)
        return fail(AMP_E_BADARG, "seg_bwd: block sizes must share a factor >= %d points (training blocks are equal-sized)", kMinGroupSlab);
    for (int i = 0; i < S_COUNT; ++i) {
        const bool is_buf = i >= S_BN2 && (i - S_BN2) % BN_STRIDE >= BN_RM;
        if (!params[i] || (!is_buf && !grads[i])) return fail(AMP_E_BADARG, "seg_bwd: parameter / gradient %d (%s) is null", i, kSegNames[i]);
    }
    if (saved_bytes < amp_seg_saved_bytes(B, W, rows, embed_dim, heads)) return fail(AMP_E_WORKSPACE, "seg_bwd: saved buffer too small");
    if (workspace_bytes < amp_seg_workspace_bytes(B, W, rows, embed_dim, 1)) return fail(AMP_E_WORKSPACE, "seg_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const float dp = dropout_p;
    const int E = embed_dim, hid = sh.hid, Bi = (int)B, Wi = (int)W, Ri = (int)rows, T = Bi * Wi, C = num_classes;
    Arena sa(saved, saved_bytes);
    SegSaved S = seg_carve(sa, B, W, rows, E, heads, hid);
    Arena wa(workspace, workspace_bytes);
    SegWs ws = seg_ws_carve(wa, B, W, rows, E, hid, true);
    const int tiles = pw_tiles(Bi, Ri);
    const long long count = (long long)Bi * Ri;
    WgDeferScope defer(ws.wg_pool, ws.wg_pool_floats, st);       // parameter-gradient reductions: one launch at the end

    // conv_4: dW4 / db4 from the [B, C, rows] logit gradients; d a3 -> dropout, relu mask, bn_3 sums
    {
        WgParams g{};
        g.dY = d_logits; g.Nout = C; g.dy_transposed = 1;
        g.A = S.y3; g.lda = 64; g.K = 64; g.a_a = S.scale + kSegBn3; g.a_b = pf(params, S_BN3 + BN_B); g.a_m = S.mean + kSegBn3; g.a_relu = 1;
        g.a_drop_p = dp; g.drop_off = dropout_offset(); g.a_drop_seed = seed + 3;
        g.n_clouds = Bi; g.rows_per_cloud = Ri; g.dW = gf(grads, S_C4W); g.ldw = 64; g.db = gf(grads, S_C4B);
        g.partials = ws.wg; g.partial_floats = ws.wg_floats;
        AMP_TRY(wgrad(g, st));
        PwParams p{};
        p.X = d_logits; p.K = C; p.x_transposed = 1; p.W = pf(params, S_C4W); p.ldw = 64; p.w_kn = 1; p.n_groups = 1;
        p.mask_y = S.y3; p.ld_mask = 64; p.mask_scale = S.scale + kSegBn3; p.mask_shift = pf(params, S_BN3 + BN_B);
        p.mask_mean = S.mean + kSegBn3; p.mask_invstd = S.invstd + kSegBn3; p.out_drop_p = dp; p.drop_off = dropout_offset(); p.out_drop_seed = seed + 3;
        p.part_sum = ws.part_sum; p.part_sq = ws.part_sq;
        p.Y = ws.dz3; p.ldy = 64; p.n_clouds = Bi; p.rows_per_cloud = Ri; p.Nout = 64;
        AMP_TRY(pw_linear(p, st));
        AMP_TRY(bn_backward_finalize(ws.part_sum, ws.part_sq, tiles, count, 64, pf(params, S_BN3 + BN_W), S.mean + kSegBn3,
                                     S.invstd + kSegBn3, gf(grads, S_BN3 + BN_W), gf(grads, S_BN3 + BN_B), 0, ws.k1 + kSegBn3,
                                     ws.k2 + kSegBn3, ws.k3 + kSegBn3, st));
    }
    // conv_3
    {
        WgParams g{};
        g.dY = ws.dz3; g.lddy = 64; g.Nout = 64; g.y_a = ws.k1 + kSegBn3; g.y_b = ws.k3 + kSegBn3; g.y_c = ws.k2 + kSegBn3; g.y_m = S.mean + kSegBn3; g.Y2 = S.y3;
        g.A = S.y2; g.lda = hid; g.K = hid; g.a_a = S.scale + kSegBn2; g.a_b = pf(params, S_BN2 + BN_B); g.a_m = S.mean + kSegBn2; g.a_relu = 1;
        g.a_drop_p = dp; g.drop_off = dropout_offset(); g.a_drop_seed = seed + 2;
        g.n_clouds = Bi; g.rows_per_cloud = Ri; g.dW = gf(grads, S_C3W); g.ldw = hid; g.db = gf(grads, S_C3B);
        g.partials = ws.wg; g.partial_floats = ws.wg_floats;
        // the input gradient first: its kernel leaves dy' split for the weight gradient (see bwd_step of nn_encoder.cu)
        const bool want_dump = !path_disabled("wgrad_presplit");
        PwParams p{};
        if (want_dump) p.split_dump = ws.split_dump;
        p.X = ws.dz3; p.ldx = 64; p.K = 64; p.in_a = ws.k1 + kSegBn3; p.in_b = ws.k3 + kSegBn3; p.in_c = ws.k2 + kSegBn3; p.in_m = S.mean + kSegBn3; p.X2 = S.y3;
        p.W = pf(params, S_C3W); p.ldw = hid; p.w_kn = 1; p.n_groups = 1;
        p.mask_y = S.y2; p.ld_mask = hid; p.mask_scale = S.scale + kSegBn2; p.mask_shift = pf(params, S_BN2 + BN_B);
        p.mask_mean = S.mean + kSegBn2; p.mask_invstd = S.invstd + kSegBn2; p.out_drop_p = dp; p.drop_off = dropout_offset(); p.out_drop_seed = seed + 2;
        p.part_sum = ws.part_sum; p.part_sq = ws.part_sq;
        p.Y = ws.dz2; p.ldy = hid; p.n_clouds = Bi; p.rows_per_cloud = Ri; p.Nout = hid;
        tc_layer_dumped() = false;
        AMP_TRY(pw_linear(p, st));
        if (want_dump && tc_layer_dumped()) g.dy_split = ws.split_dump;
        AMP_TRY(wgrad(g, st));
        AMP_TRY(bn_backward_finalize(ws.part_sum, ws.part_sq, tiles, count, hid, pf(params, S_BN2 + BN_W), S.mean + kSegBn2,
                                     S.invstd + kSegBn2, gf(grads, S_BN2 + BN_W), gf(grads, S_BN2 + BN_B), 0, ws.k1 + kSegBn2,
                                     ws.k2 + kSegBn2, ws.k3 + kSegBn2, st));
    }
    // conv_2, local half: dW2[:, :64], per-block bias gradient dcb, d lo_feats
    {
        WgParams g{};
        g.dY = ws.dz2; g.lddy = hid; g.Nout = hid; g.y_a = ws.k1 + kSegBn2; g.y_b = ws.k3 + kSegBn2; g.y_c = ws.k2 + kSegBn2; g.y_m = S.mean + kSegBn2; g.Y2 = S.y2;
        g.A = lo_feats; g.lda = lo_ld; g.K = 64;
        g.n_clouds = Bi; g.rows_per_cloud = Ri; g.dW = gf(grads, S_C2W); g.ldw = 64 + E;
        g.group_rows = group_rows; g.n_groups = Wi; g.dbg = ws.dcb; g.slab_rows = gslab;
        g.partials = ws.wg; g.partial_floats = ws.wg_floats;
        const bool want_dump = gslab % 128 == 0 && !path_disabled("wgrad_presplit");
        PwParams p{};
        if (want_dump) p.split_dump = ws.split_dump;
        p.X = ws.dz2; p.ldx = hid; p.K = hid; p.in_a = ws.k1 + kSegBn2; p.in_b = ws.k3 + kSegBn2; p.in_c = ws.k2 + kSegBn2; p.in_m = S.mean + kSegBn2; p.X2 = S.y2;
        p.W = pf(params, S_C2W); p.ldw = 64 + E; p.w_kn = 1; p.n_groups = 1;
        p.Y = d_lo_feats; p.ldy = 64; p.n_clouds = Bi; p.rows_per_cloud = Ri; p.Nout = 64;
        tc_layer_dumped() = false;
        AMP_TRY(pw_linear(p, st));
        if (want_dump && tc_layer_dumped()) g.dy_split = ws.split_dump;
        AMP_TRY(wgrad(g, st));
    }
    // conv_2, global half through the per-block bias: dW2[:, 64:], db2, d g_w
    {
        WgParams g{};
        g.dY = ws.dcb; g.lddy = hid; g.Nout = hid; g.A = S.g_w; g.lda = E; g.K = E;
        g.n_clouds = 1; g.rows_per_cloud = T; g.dW = gf(grads, S_C2W) + 64; g.ldw = 64 + E; g.db = gf(grads, S_C2B);
        g.partials = ws.wg; g.partial_floats = ws.wg_floats;
        AMP_TRY(wgrad(g, st));
        PwParams p{};
        p.X = ws.dcb; p.ldx = hid; p.K = hid; p.W = pf(params, S_C2W) + 64; p.ldw = 64 + E; p.w_kn = 1; p.n_groups = 1;
        p.Y = ws.dg_w; p.ldy = E; p.n_clouds = 1; p.rows_per_cloud = T; p.Nout = E;
        AMP_TRY(pw_linear(p, st));
    }
    // out_proj
    {
        WgParams g{};
        g.dY = ws.dg_w; g.lddy = E; g.Nout = E; g.A = S.attn_o; g.lda = E; g.K = E;
        g.n_clouds = 1; g.rows_per_cloud = T; g.dW = gf(grads, S_OUTW); g.ldw = E; g.db = gf(grads, S_OUTB);
        g.partials = ws.wg; g.partial_floats = ws.wg_floats;
        AMP_TRY(wgrad(g, st));
        PwParams p{};
        p.X = ws.dg_w; p.ldx = E; p.K = E; p.W = pf(params, S_OUTW); p.ldw = E; p.w_kn = 1; p.n_groups = 1;
        p.Y = ws.dattn_o; p.ldy = E; p.n_clouds = 1; p.rows_per_cloud = T; p.Nout = E;
        AMP_TRY(pw_linear(p, st));
    }
    AMP_TRY(attention_core_bwd(ws.dattn_o, S.qkv, S.probs, dp, seed + 1, Bi, Wi, E, heads, ws.dqkv, st));
    // in_proj
    {
        WgParams g{};
        g.dY = ws.dqkv; g.lddy = 3 * E; g.Nout = 3 * E; g.A = S.tokens; g.lda = E; g.K = E;
        g.n_clouds = 1; g.rows_per_cloud = T; g.dW = gf(grads, S_INW); g.ldw = E; g.db = gf(grads, S_INB);
        g.partials = ws.wg; g.partial_floats = ws.wg_floats;
        AMP_TRY(wgrad(g, st));
        PwParams p{};
        p.X = ws.dqkv; p.ldx = 3 * E; p.K = 3 * E; p.W = pf(params, S_INW); p.ldw = E; p.w_kn = 1; p.n_groups = 1;
        p.Y = ws.dtokens; p.ldy = E; p.n_clouds = 1; p.rows_per_cloud = T; p.Nout = E;
        AMP_TRY(pw_linear(p, st));
    }
    // positional encoding
    AMP_TRY(posenc_bwd(ws.dtokens, centroids, S.h_pre, pf(params, S_FC2W), Bi, Wi, E, d_gl_feats, ws.dpre, gf(grads, S_FC1W),
                      gf(grads, S_FC1B), gf(grads, S_FC2W), gf(grads, S_FC2B), st));
    return defer.flush();
}

}  // extern "C"
