// Point-wise linear layer in fp32 on the CUDA cores: the parity-mode path of every Conv1d(k=1) /
// Linear of the reference (pointNet/model/pointnetAtt.py:31-40, 90-103, 203-207) with BatchNorm,
// ReLU, Dropout, bias, the max-pool over the points of a cloud (MaxPool1d, :35, :104) and the BatchNorm
// batch statistics fused into the GEMM prologue / epilogue, so an activation is read once and
// written once. The same kernel computes the input gradients dX = dY @ W of the backward pass with
// the BatchNorm-backward apply in its prologue and the ReLU mask + BatchNorm-backward sums in its
// epilogue.
//
// Tiling: CTA = 128 rows x BN (64|128) output channels, K in chunks of 16, 256 threads, each thread
// an 8 x (BN/16) register tile; double-buffered shared memory with register prefetch. The tile never
// straddles two clouds (grid.z = cloud), so pooling reduces inside the tile and finishes with one
// 64-bit atomicMax per (cloud, channel) on a packed (value, row) key: first maximum wins, like
// MaxPool1d's argmax.
#include "nn_common.cuh"

namespace amp {
namespace {

constexpr int BM = 128, BK = 16, NT = 256, PAD = 4;

template <int BN>
__global__ void __launch_bounds__(NT, 2)
pw_linear_kernel(const PwParams p) {
    pdl_sync();
    constexpr int TN = BN / 16;                 // output channels per thread: 8 or 4
    constexpr int WREG = BN * BK / NT;          // W elements staged per thread per chunk
    constexpr int XREG = BM * BK / NT;          // = 8
    __shared__ __align__(16) float Xs[2][BK][BM + PAD];
    __shared__ __align__(16) float Ws[2][BK][BN + PAD];

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int cloud = blockIdx.z;
    const int r0 = blockIdx.x * BM;             // first row of the tile inside the cloud
    const int n0 = blockIdx.y * BN;
    const int rows = p.rows_per_cloud;
    const long long row_base = (long long)cloud * rows + r0;
    const float* __restrict__ X = p.x_transposed ? p.X + (long long)cloud * p.K * rows + r0 : p.X + row_base * p.ldx;
    const float* __restrict__ X2 = p.X2 ? p.X2 + row_base * p.ldx : nullptr;
    const float* __restrict__ W = p.W + (long long)cloud * p.w_cloud_stride;
    const int K = p.K;
    const int nk = (K + BK - 1) / BK;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float xr[XREG], wr[WREG];
    auto g_load = [&](int kt) {
        const int k0 = kt * BK;
#pragma unroll
        for (int i = 0; i < XREG; ++i) {
            const int e = tid + i * NT, k = e & (BK - 1), r = e >> 4;
            float v = 0.f;
            if (k0 + k < K && r0 + r < rows) {
                const long long off = p.x_transposed ? (long long)(k0 + k) * rows + r : (long long)r * p.ldx + k0 + k;
                v = __ldg(X + off);
                const float m = p.in_m ? __ldg(p.in_m + k0 + k) : 0.f;
                if (X2) v = fmaf(__ldg(X2 + off) - m, __ldg(p.in_c + k0 + k), fmaf(v, __ldg(p.in_a + k0 + k), __ldg(p.in_b + k0 + k)));
                else if (p.in_a) v = fmaf(v - m, __ldg(p.in_a + k0 + k), __ldg(p.in_b + k0 + k));
                if (p.in_relu) v = fmaxf(v, 0.f);
                if (p.in_drop_p > 0.f)
                    v *= dropout_keep(eff_seed(p.in_drop_seed, p.drop_off), (unsigned long long)(row_base + r) * K + k0 + k, p.in_drop_p);
            }
            xr[i] = v;
        }
#pragma unroll
        for (int i = 0; i < WREG; ++i) {
            const int e = tid + i * NT;
            float v = 0.f;
            if (p.w_kn == 0) {
                const int k = e & (BK - 1), n = e >> 4;
                if (k0 + k < K && n0 + n < p.Nout) v = __ldg(W + (long long)(n0 + n) * p.ldw + k0 + k);
            } else {
                const int n = e % BN, k = e / BN;
                if (k0 + k < K && n0 + n < p.Nout) v = __ldg(W + (long long)(k0 + k) * p.ldw + n0 + n);
            }
            wr[i] = v;
        }
    };
    auto s_store = [&](int buf) {
#pragma unroll
        for (int i = 0; i < XREG; ++i) {
            const int e = tid + i * NT;
            Xs[buf][e & (BK - 1)][e >> 4] = xr[i];
        }
#pragma unroll
        for (int i = 0; i < WREG; ++i) {
            const int e = tid + i * NT;
            if (p.w_kn == 0) Ws[buf][e & (BK - 1)][e >> 4] = wr[i];
            else Ws[buf][e / BN][e % BN] = wr[i];
        }
    };

    g_load(0);
    s_store(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) g_load(kt + 1);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&Xs[buf][k][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&Xs[buf][k][ty * 8 + 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float b[TN];
            const float4 b0 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
            if (TN == 8) {
                const float4 b1 = *reinterpret_cast<const float4*>(&Ws[buf][k][64 + tx * 4]);
                b[TN - 4] = b1.x; b[TN - 3] = b1.y; b[TN - 2] = b1.z; b[TN - 1] = b1.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) s_store(buf ^ 1);
        __syncthreads();
    }

    // ---------------- epilogue ----------------
    // column of acc[.][j]: n0 + cl(j), cl(j) = (j < 4 ? tx*4 + j : 64 + tx*4 + (j-4));  row of acc[i][.]: r0 + ty*8 + i
    int cl[TN], col[TN];
#pragma unroll
    for (int j = 0; j < TN; ++j) { cl[j] = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)); col[j] = n0 + cl[j]; }
    const int rl0 = ty * 8;                                   // first local row of this thread
    // bias (per group of rows when group_rows is given)
    if (p.bias) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = r0 + rl0 + i;
            int g = 0;
            if (p.group_rows) {
                for (int q = 1; q < p.n_groups; ++q) g += (r >= __ldg(p.group_rows + q)) ? 1 : 0;
            }
            const float* bp = p.bias + ((long long)cloud * p.n_groups + g) * p.bias_group_stride;
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (col[j] < p.Nout) acc[i][j] += __ldg(bp + col[j]);
        }
    }
    if (p.accumulate && p.Y) {
        const float* __restrict__ Yo = p.Y + row_base * p.ldy;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (r0 + rl0 + i >= rows) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j)
                if (col[j] < p.Nout) acc[i][j] += Yo[(long long)(rl0 + i) * p.ldy + col[j]];
        }
    }
    float* red = &Xs[0][0][0];                              // 2*16*132 floats >= 2*16*128
    const long long tile = (long long)cloud * gridDim.x + blockIdx.x;
    // per-tile sum and sum of squared deviations from the tile mean of the raw value (BatchNorm batch
    // statistics), fixed summation order
    if (p.part_sum && !p.mask_y) {
        __syncthreads();
        const int n_t = min(BM, rows - r0);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += (r0 + rl0 + i < rows) ? acc[i][j] : 0.f;
            red[ty * BN + cl[j]] = s;
        }
        __syncthreads();
        if (tid < BN) {
            float s = 0.f;
#pragma unroll
            for (int t = 0; t < 16; ++t) s += red[t * BN + tid];
            red[16 * BN + tid] = s / (float)n_t;
            if (n0 + tid < p.Nout) p.part_sum[tile * p.Nout + n0 + tid] = s;
        }
        __syncthreads();
        float mt[TN];
#pragma unroll
        for (int j = 0; j < TN; ++j) mt[j] = red[16 * BN + cl[j]];
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float q = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float d = (r0 + rl0 + i < rows) ? acc[i][j] - mt[j] : 0.f;
                q = fmaf(d, d, q);
            }
            red[ty * BN + cl[j]] = q;
        }
        __syncthreads();
        if (tid < BN && n0 + tid < p.Nout) {
            float q = 0.f;
#pragma unroll
            for (int t = 0; t < 16; ++t) q += red[t * BN + tid];
            p.part_sq[tile * p.Nout + n0 + tid] = q;
        }
    }
    if (p.pool_mode == 2) {                                  // max / min of the raw value
        __syncthreads();
        unsigned long long* r64 = reinterpret_cast<unsigned long long*>(red);   // [16][BN] keys (16 KB)
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                unsigned long long best = 0ull;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = r0 + rl0 + i;
                    if (r < rows) {
                        unsigned int ob = ordered_bits(acc[i][j]);
                        if (pass) ob = ~ob;
                        const unsigned long long key = ((unsigned long long)ob << 32) | (0xffffffffu - (unsigned)r);
                        best = key > best ? key : best;
                    }
                }
                r64[ty * BN + cl[j]] = best;
            }
            __syncthreads();
            if (tid < BN && n0 + tid < p.Nout) {
                unsigned long long best = 0ull;
#pragma unroll
                for (int t = 0; t < 16; ++t) { const unsigned long long k = r64[t * BN + tid]; best = k > best ? k : best; }
                atomicMax((pass ? p.pool_min : p.pool_max) + (long long)cloud * p.Nout + n0 + tid, best);
            }
            __syncthreads();
        }
    }
    // output affine + ReLU (eval BatchNorm folded to scale/shift)
    if (p.out_scale) {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            if (col[j] < p.Nout) {
                const float s = __ldg(p.out_scale + col[j]), t = __ldg(p.out_shift + col[j]);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i][j] = fmaf(acc[i][j], s, t);
            }
        }
    }
    if (p.out_relu) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaxf(acc[i][j], 0.f);
    }
    // backward: dropout keep-scale, ReLU mask, BatchNorm-backward sums
    if (p.mask_y) {
        const float* __restrict__ My = p.mask_y + row_base * p.ld_mask;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            float s = 0.f, q = 0.f;
            if (col[j] < p.Nout) {
                const float ms = __ldg(p.mask_scale + col[j]), mt = __ldg(p.mask_shift + col[j]);
                const float mu = __ldg(p.mask_mean + col[j]);
                const float is = p.mask_invstd ? __ldg(p.mask_invstd + col[j]) : 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rl0 + i;
                    float dz = 0.f;
                    if (r0 + r < rows) {
                        const float y = My[(long long)r * p.ld_mask + col[j]];
                        dz = acc[i][j];
                        if (p.out_drop_p > 0.f)
                            dz *= dropout_keep(eff_seed(p.out_drop_seed, p.drop_off), (unsigned long long)(row_base + r) * p.Nout + col[j], p.out_drop_p);
                        dz = (fmaf(y - mu, ms, mt) > 0.f) ? dz : 0.f;
                        s += dz;
                        q = fmaf(dz, (y - mu) * is, q);
                    }
                    acc[i][j] = dz;
                }
            }
            if (p.part_sum) { red[ty * BN + cl[j]] = s; red[16 * BN + ty * BN + cl[j]] = q; }
        }
        if (p.part_sum) {
            __syncthreads();
            if (tid < BN && n0 + tid < p.Nout) {
                float s = 0.f, q = 0.f;
#pragma unroll
                for (int t = 0; t < 16; ++t) { s += red[t * BN + tid]; q += red[16 * BN + t * BN + tid]; }
                p.part_sum[tile * p.Nout + n0 + tid] = s;
                p.part_sq[tile * p.Nout + n0 + tid] = q;
            }
        }
    }
    if (p.pool_mode == 1) {                                  // max of the final value
        __syncthreads();
        unsigned long long* r64 = reinterpret_cast<unsigned long long*>(red);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            unsigned long long best = 0ull;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = r0 + rl0 + i;
                if (r < rows) {
                    const unsigned long long key = ((unsigned long long)ordered_bits(acc[i][j]) << 32) | (0xffffffffu - (unsigned)r);
                    best = key > best ? key : best;
                }
            }
            r64[ty * BN + cl[j]] = best;
        }
        __syncthreads();
        if (tid < BN && n0 + tid < p.Nout) {
            unsigned long long best = 0ull;
#pragma unroll
            for (int t = 0; t < 16; ++t) { const unsigned long long k = r64[t * BN + tid]; best = k > best ? k : best; }
            atomicMax(p.pool_max + (long long)cloud * p.Nout + n0 + tid, best);
        }
    }
    if (p.Y) {
        if (p.y_transposed) {
            float* __restrict__ Y = p.Y + (long long)cloud * p.Nout * rows;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                if (col[j] >= p.Nout) continue;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (r0 + rl0 + i < rows) Y[(long long)col[j] * rows + r0 + rl0 + i] = acc[i][j];
            }
        } else {
            float* __restrict__ Y = p.Y + row_base * p.ldy;
            const bool vec = ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0) && (n0 + BN <= p.Nout);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rl0 + i;
                if (r0 + r >= rows) continue;
                if (vec) {
                    *reinterpret_cast<float4*>(Y + (long long)r * p.ldy + col[0]) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                    if (TN == 8)
                        *reinterpret_cast<float4*>(Y + (long long)r * p.ldy + col[TN - 4]) =
                            make_float4(acc[i][TN - 4], acc[i][TN - 3], acc[i][TN - 2], acc[i][TN - 1]);
                } else {
#pragma unroll
                    for (int j = 0; j < TN; ++j)
                        if (col[j] < p.Nout) Y[(long long)r * p.ldy + col[j]] = acc[i][j];
                }
            }
        }
    }
}


}  // namespace

int pw_tiles(int n_clouds, int rows_per_cloud) { return n_clouds * ((rows_per_cloud + BM - 1) / BM); }

int pw_linear(const PwParams& p, cudaStream_t st) {
    if (!p.X || !p.W) return fail(AMP_E_BADARG, "pw_linear: null operand");
    if (p.K < 1 || p.Nout < 1 || p.n_clouds < 1 || p.rows_per_cloud < 1)
        return fail(AMP_E_BADARG, "pw_linear: bad shape K=%d Nout=%d clouds=%d rows=%d", p.K, p.Nout, p.n_clouds, p.rows_per_cloud);
    if (p.n_clouds > 65535) return fail(AMP_E_BADARG, "pw_linear: more than 65535 clouds in one launch");
    if ((p.in_a == nullptr) != (p.in_b == nullptr) || (p.out_scale == nullptr) != (p.out_shift == nullptr))
        return fail(AMP_E_BADARG, "pw_linear: scale and shift must come together");
    if (p.X2 && (!p.in_a || !p.in_c || p.x_transposed)) return fail(AMP_E_BADARG, "pw_linear: X2 needs in_a, in_b, in_c");
    if ((p.pool_mode && !p.pool_max) || (p.pool_mode == 2 && !p.pool_min) || ((p.part_sum == nullptr) != (p.part_sq == nullptr)))
        return fail(AMP_E_BADARG, "pw_linear: pooling / statistics buffers missing");
    if (p.mask_y && (!p.mask_scale || !p.mask_shift || !p.mask_mean)) return fail(AMP_E_BADARG, "pw_linear: mask needs scale, shift and mean");
    if (p.mask_y && p.part_sum && !p.mask_invstd) return fail(AMP_E_BADARG, "pw_linear: backward sums need invstd");
    if (p.y_transposed && p.accumulate) return fail(AMP_E_BADARG, "pw_linear: accumulate into a transposed output");
    if (p.group_rows && p.n_groups > 1024) return fail(AMP_E_BADARG, "pw_linear: more than 1024 row groups");
    // few rows (per-cloud FC stacks, per-token projections): weight-streaming kernels (nn_small.cu)
    {
        const int rc = small_linear_try(p, st);
        if (rc != 0) return rc < 0 ? rc : AMP_OK;
    }
    // many rows, 3 / 9 input columns, 64 channels: exact fp32, output-bandwidth bound (nn_small.cu)
    {
        const int rc = narrow_fwd_try(p, st);
        if (rc != 0) return rc < 0 ? rc : AMP_OK;
    }
    // many rows, a handful of output channels (class logits): exact fp32, input-bandwidth bound (nn_small.cu)
    {
        const int rc = narrow_out_fwd_try(p, st);
        if (rc != 0) return rc < 0 ? rc : AMP_OK;
    }
    // many rows, tensor-core friendly K: split-bf16 tcgen05 path (nn_tc_layer.cu)
    {
        const int rc = tc_layer_try(p, st);
        if (rc != 0) return rc < 0 ? rc : AMP_OK;
    }
    PwParams q = p;
    if (q.n_groups < 1) q.n_groups = 1;
    const bool wide = p.Nout > 64;
    const int bn = wide ? 128 : 64;
    dim3 grid((p.rows_per_cloud + BM - 1) / BM, (p.Nout + bn - 1) / bn, p.n_clouds);
    count_path("pw_linear");
    if (wide) launch_pdl(pw_linear_kernel<128>, grid, dim3(NT), 0, st, q);
    else launch_pdl(pw_linear_kernel<64>, grid, dim3(NT), 0, st, q);
    count_launch();
    return check_launch("pw_linear");
}

}  // namespace amp
