// Parameter order and buffer layouts shared by the encoder / segmentation-head host code.
//
// Parameter tables follow the state_dict order of the reference modules (pointNet/model/pointnetAtt.py:
// TransformationNet :10-26, BasePointNet :59-78, SegmentationWithAttention :160-174); the Python side
// passes one device pointer per state_dict entry in that order (amp_encoder_param_name(i) /
// amp_seg_param_name(i) let it check the order by name).
#pragma once
#include "nn_common.cuh"

namespace amp {

// BatchNorm block inside a table: weight, bias, running_mean, running_var, num_batches_tracked
enum { BN_W = 0, BN_B = 1, BN_RM = 2, BN_RV = 3, BN_NBT = 4, BN_STRIDE = 5 };
// TransformationNet block
enum { T_CONV1 = 0, T_CONV2 = 1, T_CONV3 = 2, T_BN1 = 3, T_BN2 = 8, T_BN3 = 13, T_BN4 = 18, T_BN5 = 23,
       T_FC1 = 28, T_FC2 = 29, T_FC3W = 30, T_FC3B = 31, T_COUNT = 32 };
// BasePointNet
enum { E_IT = 0, E_FT = 32, E_CONV1 = 64, E_CONV2 = 65, E_CONV3 = 66, E_CONV4 = 67, E_CONV5 = 68, E_CONV6 = 69,
       E_BN1 = 70, E_BN2 = 75, E_BN3 = 80, E_BN4 = 85, E_BN5 = 90, E_BN6 = 95, E_COUNT = 100 };
// SegmentationWithAttention
enum { S_FC1W = 0, S_FC1B, S_FC2W, S_FC2B, S_INW, S_INB, S_OUTW, S_OUTB, S_C2W, S_C2B, S_C3W, S_C3B, S_C4W, S_C4B,
       S_BN2 = 14, S_BN3 = 19, S_COUNT = 24 };

// BatchNorm layers of the encoder in a fixed order (index into the scale/shift / saved-stat tables)
enum { L_IT1 = 0, L_IT2, L_IT3, L_IT4, L_IT5, L_FT1, L_FT2, L_FT3, L_FT4, L_FT5, L_C1, L_C2, L_C3, L_C4, L_C5, L_C6, L_ENC_BN = 16 };
static const int kEncBnChannels[L_ENC_BN] = {64, 128, 256, 256, 128, 64, 128, 256, 256, 128, 64, 64, 64, 128, 128, 256};
static const int kEncBnParam[L_ENC_BN] = {E_IT + T_BN1, E_IT + T_BN2, E_IT + T_BN3, E_IT + T_BN4, E_IT + T_BN5,
                                          E_FT + T_BN1, E_FT + T_BN2, E_FT + T_BN3, E_FT + T_BN4, E_FT + T_BN5,
                                          E_BN1, E_BN2, E_BN3, E_BN4, E_BN5, E_BN6};
constexpr int kEncBnTotal = 2368;      // sum of kEncBnChannels
inline int enc_bn_offset(int layer) { int o = 0; for (int i = 0; i < layer; ++i) o += kEncBnChannels[i]; return o; }

// bump allocator over a caller-provided buffer (256-byte aligned pieces); with base == nullptr it only measures
struct Arena {
    char* base; size_t size; size_t off;
    Arena(void* b, size_t s) : base(reinterpret_cast<char*>(b)), size(s), off(0) {}
    template <typename T> T* take(size_t n) {
        off = align_up(off, 256);
        T* p = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return p;
    }
    bool ok() const { return off <= size; }
};

// Everything the encoder forward produces besides its outputs. Training: carved from the `saved` buffer the
// caller keeps for backward (raw pre-BatchNorm layer outputs, pooled argmax, batch statistics). Eval: carved
// from the workspace, per-point buffers aliased (activations are final values and die after their consumer).
struct EncSaved {
    float *it_y1, *it_y2, *it_y3, *c1, *c2, *ft_y1, *ft_y2, *ft_y3, *c3, *c4, *c5, *c6;   // per point
    float *it_pool, *it_f1, *it_f2, *T, *W1eff, *ft_pool, *ft_f1, *ft_f2, *G;             // per cloud
    int *it_arg, *ft_arg, *g_arg;
    float *scale, *shift, *mean, *invstd;                                                  // [kEncBnTotal]
};

inline EncSaved enc_carve(Arena& a, long long B, long long N, bool training) {
    EncSaved s{};
    const size_t M = (size_t)B * N;
    if (training) {
        s.it_y1 = a.take<float>(M * 64); s.it_y2 = a.take<float>(M * 128); s.it_y3 = a.take<float>(M * 256);
        s.c1 = a.take<float>(M * 64); s.c2 = a.take<float>(M * 64);
        s.ft_y1 = a.take<float>(M * 64); s.ft_y2 = a.take<float>(M * 128); s.ft_y3 = a.take<float>(M * 256);
        s.c3 = a.take<float>(M * 64); s.c4 = a.take<float>(M * 128); s.c5 = a.take<float>(M * 128);
        s.c6 = a.take<float>(M * 256);
    } else {
        float* b64a = a.take<float>(M * 64); float* b64b = a.take<float>(M * 64);
        float* b128a = a.take<float>(M * 128); float* b128b = a.take<float>(M * 128);
        s.it_y1 = s.c1 = s.ft_y1 = s.c3 = b64a;
        s.c2 = b64b;
        s.it_y2 = s.ft_y2 = s.c4 = b128a;
        s.c5 = b128b;
        s.it_y3 = s.ft_y3 = s.c6 = nullptr;      // pooled in the epilogue, never stored
    }
    s.it_pool = a.take<float>(B * 256); s.it_f1 = a.take<float>(B * 256); s.it_f2 = a.take<float>(B * 128);
    s.T = a.take<float>(B * 9 + 7); s.W1eff = a.take<float>(B * 576);
    s.ft_pool = a.take<float>(B * 256); s.ft_f1 = a.take<float>(B * 256); s.ft_f2 = a.take<float>(B * 128);
    s.G = a.take<float>(B * 256);
    s.it_arg = a.take<int>(B * 256); s.ft_arg = a.take<int>(B * 256); s.g_arg = a.take<int>(B * 256);
    s.scale = a.take<float>(kEncBnTotal); s.shift = a.take<float>(kEncBnTotal);
    s.mean = a.take<float>(kEncBnTotal); s.invstd = a.take<float>(kEncBnTotal);
    return s;
}

}  // namespace amp
