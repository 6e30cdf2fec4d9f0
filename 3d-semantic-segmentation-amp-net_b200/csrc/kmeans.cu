// K-means block split for sm_100a.
//
// Replaces the `KMeansConstrained(...).fit_predict(in_pc[:, i_f])` call of the reference
// (data_proc/3_kmeans.py:78-82, utils/utils.py:500-505) and the regroup that follows it
// (3_kmeans.py:88-101, utils.py:507-517). The solver behind that call is third-party and random;
// the deterministic rules implemented here are the ones DEFINED in oracle/kmeans_oracle.py and
// are reproduced bit-for-bit:
//   assign   argmin_j (d0*d0 + d1*d1) + d2*d2, float32, no FMA, first minimum
//   init     farthest-point sampling of k rows on the 3 features, start row 0
//   update   order-independent fixed-point sums (rint(x * 2^32) as int64), mean in float64
//   stop     centre shift (float64, fixed order) <= tol * mean variance
//   balance  capacity rounds with an exact (d^2, index) radix select per over-subscribed cluster
#include "amp_common.cuh"

namespace amp {
namespace {

constexpr int kKMax = 32;
constexpr double kFix = 4294967296.0;  // 2^32

__device__ __forceinline__ float sqd3(float x0, float x1, float x2, float c0, float c1, float c2) {
    float d0 = __fsub_rn(x0, c0), d1 = __fsub_rn(x1, c1), d2 = __fsub_rn(x2, c2);
    return __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
}

// ---------------------------------------------------------------------------------------------
// Stand-alone assignment step: HBM-bound stream, 12 B read + 4 B (+4 B) written per point.
// Each thread owns 4 consecutive points = three 16-byte loads, one 16-byte label store.
// ---------------------------------------------------------------------------------------------
constexpr int kAssignThreads = 256;

__global__ void __launch_bounds__(kAssignThreads)
kmeans_assign_kernel(const float* __restrict__ feats, const float* __restrict__ cent, long long n, int k,
                     int* __restrict__ labels, float* __restrict__ min_d2) {
    __shared__ float sc[64 * 3];
    for (int i = threadIdx.x; i < k * 3; i += blockDim.x) sc[i] = cent[i];
    __syncthreads();
    const long long nquad = n >> 2;
    const float4* f4 = reinterpret_cast<const float4*>(feats);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nquad;
         q += (long long)gridDim.x * blockDim.x) {
        float4 a = __ldg(f4 + 3 * q), b = __ldg(f4 + 3 * q + 1), c = __ldg(f4 + 3 * q + 2);
        float px[4] = {a.x, a.w, b.z, c.y};
        float py[4] = {a.y, b.x, b.w, c.z};
        float pz[4] = {a.z, b.y, c.x, c.w};
        float best[4];
        int bj[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) { best[p] = INFINITY; bj[p] = 0; }
#pragma unroll 3
        for (int j = 0; j < k; ++j) {
            float c0 = sc[3 * j], c1 = sc[3 * j + 1], c2 = sc[3 * j + 2];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                float d = sqd3(px[p], py[p], pz[p], c0, c1, c2);
                if (d < best[p]) { best[p] = d; bj[p] = j; }
            }
        }
        reinterpret_cast<int4*>(labels)[q] = make_int4(bj[0], bj[1], bj[2], bj[3]);
        if (min_d2) reinterpret_cast<float4*>(min_d2)[q] = make_float4(best[0], best[1], best[2], best[3]);
    }
    // tail (n % 4 points), first block only
    if (blockIdx.x == 0) {
        for (long long i = (nquad << 2) + threadIdx.x; i < n; i += blockDim.x) {
            float x0 = feats[3 * i], x1 = feats[3 * i + 1], x2 = feats[3 * i + 2];
            float best = INFINITY;
            int bj = 0;
            for (int j = 0; j < k; ++j) {
                float d = sqd3(x0, x1, x2, sc[3 * j], sc[3 * j + 1], sc[3 * j + 2]);
                if (d < best) { best = d; bj = j; }
            }
            labels[i] = bj;
            if (min_d2) min_d2[i] = best;
        }
    }
}

// Small k (the block split uses k = ceil(points / 2048), 9 on configs[3]): centroids live in registers, the loop over the
// centroids is fully unrolled, and half of the arithmetic goes through the packed fp32 pair instructions of sm_100
// (FFMA2: two points per instruction, each half rounded exactly like the scalar operation). Packed instructions only
// issue to one of the two fp32 pipes, scalar ones to either, so the work is split: the subtractions and the third
// product stay scalar, the first two products and the two additions are packed. ptxas contracts mul.f32x2 + add.f32x2
// into one FFMA2 even with .rn, which would change the result, so products and sums are written as fused multiply-adds
// that restate the unfused operations exactly, rn(a * b) = fma(a, b, -0) and rn(a + b) = fma(a, 1, b), with the
// constants -0 and 1 arriving as kernel arguments (opaque to the compiler).
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float lo, float hi) {
    f32x2_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) {
    f32x2_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int K>
__global__ void __launch_bounds__(kAssignThreads)
kmeans_assign_small_k_kernel(const float* __restrict__ feats, const float* __restrict__ cent, long long n, int* __restrict__ labels,
                             float* __restrict__ min_d2, float neg_zero, float plus_one) {
    float cx[K], cy[K], cz[K];
#pragma unroll
    for (int j = 0; j < K; ++j) { cx[j] = __ldg(cent + 3 * j); cy[j] = __ldg(cent + 3 * j + 1); cz[j] = __ldg(cent + 3 * j + 2); }
    const f32x2_t neg0 = pack2(neg_zero, neg_zero), one = pack2(plus_one, plus_one);
    const long long nquad = n >> 2;
    const float4* f4 = reinterpret_cast<const float4*>(feats);
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < nquad; q += (long long)gridDim.x * blockDim.x) {
        const float4 a = __ldg(f4 + 3 * q), b = __ldg(f4 + 3 * q + 1), c = __ldg(f4 + 3 * q + 2);
        const float px[4] = {a.x, a.w, b.z, c.y};
        const float py[4] = {a.y, b.x, b.w, c.z};
        const float pz[4] = {a.z, b.y, c.x, c.w};
        float best[4];
        int bj[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) { best[p] = INFINITY; bj[p] = 0; }
#pragma unroll
        for (int j = 0; j < K; ++j) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {             // points (0, 1) and (2, 3)
                const int p0 = 2 * h, p1 = 2 * h + 1;
                const float d0a = __fsub_rn(px[p0], cx[j]), d0b = __fsub_rn(px[p1], cx[j]);
                const float d1a = __fsub_rn(py[p0], cy[j]), d1b = __fsub_rn(py[p1], cy[j]);
                const float d2a = __fsub_rn(pz[p0], cz[j]), d2b = __fsub_rn(pz[p1], cz[j]);
                const f32x2_t d0 = pack2(d0a, d0b), d1 = pack2(d1a, d1b);
                const f32x2_t m2 = pack2(__fmul_rn(d2a, d2a), __fmul_rn(d2b, d2b));
                const f32x2_t s = fma2(fma2(fma2(d0, d0, neg0), one, fma2(d1, d1, neg0)), one, m2);
                float da, db;
                unpack2(s, da, db);
                if (da < best[p0]) { best[p0] = da; bj[p0] = j; }
                if (db < best[p1]) { best[p1] = db; bj[p1] = j; }
            }
        }
        reinterpret_cast<int4*>(labels)[q] = make_int4(bj[0], bj[1], bj[2], bj[3]);
        if (min_d2) reinterpret_cast<float4*>(min_d2)[q] = make_float4(best[0], best[1], best[2], best[3]);
    }
    if (blockIdx.x == 0) {                            // tail (n % 4 points)
        for (long long i = (nquad << 2) + threadIdx.x; i < n; i += blockDim.x) {
            const float x0 = feats[3 * i], x1 = feats[3 * i + 1], x2 = feats[3 * i + 2];
            float best = INFINITY;
            int bj = 0;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const float d = sqd3(x0, x1, x2, cx[j], cy[j], cz[j]);
                if (d < best) { best = d; bj = j; }
            }
            labels[i] = bj;
            if (min_d2) min_d2[i] = best;
        }
    }
}

template <int K>
void launch_assign_small_k(const float* feats, const float* cent, long long n, int* labels, float* min_d2, unsigned blocks,
                           cudaStream_t st) {
    kmeans_assign_small_k_kernel<K><<<blocks, kAssignThreads, 0, st>>>(feats, cent, n, labels, min_d2, -0.0f, 1.0f);
}

__global__ void gather_feats_kernel(const float* __restrict__ pc, long long n, long long row_stride,
                                    int c0, int c1, int c2, float* __restrict__ feats) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        const float* p = pc + i * row_stride;
        feats[3 * i] = p[c0];
        feats[3 * i + 1] = p[c1];
        feats[3 * i + 2] = p[c2];
    }
}

// ---------------------------------------------------------------------------------------------
// Whole constrained k-means of one window per CTA.
// ---------------------------------------------------------------------------------------------
constexpr int kWinThreads = 1024;

struct WinShared {
    float cent[kKMax * 3];
    float cent_new[kKMax * 3];
    unsigned long long sums[kKMax * 3];
    unsigned long long mom[6];
    int counts[kKMax];
    int room[kKMax];
    int nprop[kKMax];
    int over[kKMax];
    unsigned long long prefix[kKMax];
    int rank[kKMax];
    unsigned hist[kKMax * 256];
    int red_bits[32];
    unsigned red_idx[32];
    int n_todo;
    int n_open;
    int flag;
    int pick;
    double tol_abs;
};

__device__ __forceinline__ void block_argmax(WinShared& s, int kb, unsigned idx, int& out_idx) {
    // (kb desc, idx asc) over the block; result broadcast through s.pick
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int m = __reduce_max_sync(0xffffffffu, kb);
    unsigned c = (kb == m) ? idx : 0xffffffffu;
    unsigned mi = __reduce_min_sync(0xffffffffu, c);
    if (lane == 0) { s.red_bits[warp] = m; s.red_idx[warp] = mi; }
    __syncthreads();
    if (warp == 0) {
        int kb2 = s.red_bits[lane];
        unsigned i2 = s.red_idx[lane];
        int m2 = __reduce_max_sync(0xffffffffu, kb2);
        unsigned c2 = (kb2 == m2) ? i2 : 0xffffffffu;
        unsigned mi2 = __reduce_min_sync(0xffffffffu, c2);
        if (lane == 0) s.pick = (int)mi2;
    }
    __syncthreads();
    out_idx = s.pick;
}

// Capacity rounds (oracle/kmeans_oracle.py::_capacity_rounds). labels < 0 = unassigned.
__device__ void capacity_rounds(WinShared& s, const float* __restrict__ x, int n, int k,
                                int* __restrict__ labels, int* __restrict__ prop, float* __restrict__ pd) {
    const int tid = threadIdx.x;
    while (true) {
        if (tid == 0) { s.n_todo = 0; s.n_open = 0; }
        if (tid < k) s.nprop[tid] = 0;
        __syncthreads();
        if (tid < k && s.room[tid] > 0) atomicAdd(&s.n_open, 1);
        // proposals
        int local_todo = 0;
        for (int i = tid; i < n; i += kWinThreads) {
            if (labels[i] >= 0) continue;
            ++local_todo;
            float x0 = x[3 * i], x1 = x[3 * i + 1], x2 = x[3 * i + 2];
            float best = INFINITY;
            int bj = -1;
            for (int j = 0; j < k; ++j) {
                if (s.room[j] <= 0) continue;
                float d = sqd3(x0, x1, x2, s.cent[3 * j], s.cent[3 * j + 1], s.cent[3 * j + 2]);
                if (d < best || bj < 0) { best = d; bj = j; }
            }
            prop[i] = bj;
            pd[i] = best;
            if (bj >= 0) atomicAdd(&s.nprop[bj], 1);
        }
        if (local_todo) atomicAdd(&s.n_todo, local_todo);
        __syncthreads();
        if (s.n_todo == 0 || s.n_open == 0) break;   // block-uniform
        // which clusters are over-subscribed; set up the radix select
        if (tid < k) {
            s.over[tid] = (s.room[tid] > 0 && s.nprop[tid] > s.room[tid]) ? 1 : 0;
            s.prefix[tid] = 0ull;
            s.rank[tid] = s.room[tid];
        }
        if (tid == 0) s.flag = 0;
        __syncthreads();
        if (tid < k && s.over[tid]) s.flag = 1;
        __syncthreads();
        if (s.flag) {
            // exact room[j]-th smallest 64-bit key (d^2 bits << 32 | index), 8 bits per pass
            for (int pass = 0; pass < 8; ++pass) {
                const int shift = 56 - 8 * pass;
                for (int i = tid; i < k * 256; i += kWinThreads) s.hist[i] = 0u;
                __syncthreads();
                for (int i = tid; i < n; i += kWinThreads) {
                    if (labels[i] >= 0) continue;
                    int j = prop[i];
                    if (j < 0 || !s.over[j]) continue;
                    unsigned long long key = ((unsigned long long)__float_as_uint(pd[i]) << 32) | (unsigned)i;
                    bool match = (pass == 0) || ((key >> (shift + 8)) == s.prefix[j]);
                    if (match) atomicAdd(&s.hist[j * 256 + (int)((key >> shift) & 255ull)], 1u);
                }
                __syncthreads();
                if (tid < k && s.over[tid]) {
                    int rk = s.rank[tid];
                    int cum = 0, dgt = 0;
                    for (; dgt < 256; ++dgt) {
                        int h = (int)s.hist[tid * 256 + dgt];
                        if (cum + h >= rk) break;
                        cum += h;
                    }
                    s.prefix[tid] = (s.prefix[tid] << 8) | (unsigned long long)dgt;
                    s.rank[tid] = rk - cum;
                }
                __syncthreads();
            }
        }
        // accept
        for (int i = tid; i < n; i += kWinThreads) {
            if (labels[i] >= 0) continue;
            int j = prop[i];
            if (j < 0) continue;
            bool ok = true;
            if (s.over[j]) {
                unsigned long long key = ((unsigned long long)__float_as_uint(pd[i]) << 32) | (unsigned)i;
                ok = key <= s.prefix[j];
            }
            if (ok) labels[i] = j;
        }
        __syncthreads();
        if (tid < k) s.room[tid] -= (s.nprop[tid] < s.room[tid]) ? s.nprop[tid] : s.room[tid];
        __syncthreads();
    }
    __syncthreads();
}

__device__ void count_labels(WinShared& s, int n, int k, const int* __restrict__ labels) {
    const int tid = threadIdx.x;
    if (tid < k) s.counts[tid] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kWinThreads) {
        int l = labels[i];
        if (l >= 0) atomicAdd(&s.counts[l], 1);
    }
    __syncthreads();
}

__device__ void plain_assign(WinShared& s, const float* __restrict__ x, int n, int k,
                             int* __restrict__ labels, bool only_unassigned) {
    for (int i = threadIdx.x; i < n; i += kWinThreads) {
        if (only_unassigned && labels[i] >= 0) continue;
        float x0 = x[3 * i], x1 = x[3 * i + 1], x2 = x[3 * i + 2];
        float best = INFINITY;
        int bj = 0;
        for (int j = 0; j < k; ++j) {
            float d = sqd3(x0, x1, x2, s.cent[3 * j], s.cent[3 * j + 1], s.cent[3 * j + 2]);
            if (d < best) { best = d; bj = j; }
        }
        labels[i] = bj;
    }
    __syncthreads();
}

__device__ void constrained_assign(WinShared& s, const float* __restrict__ x, int n, int k,
                                   int size_min, int size_max, int* __restrict__ labels,
                                   int* __restrict__ prop, float* __restrict__ pd) {
    const int tid = threadIdx.x;
    if (size_max > 0) {
        for (int i = tid; i < n; i += kWinThreads) labels[i] = -1;
        if (size_min > 0 && size_min < size_max) {
            if (tid < k) s.room[tid] = size_min;
            __syncthreads();
            capacity_rounds(s, x, n, k, labels, prop, pd);
            count_labels(s, n, k, labels);
            if (tid < k) s.room[tid] = size_max - s.counts[tid];
            __syncthreads();
            capacity_rounds(s, x, n, k, labels, prop, pd);
        } else {
            if (tid < k) s.room[tid] = size_max;
            __syncthreads();
            capacity_rounds(s, x, n, k, labels, prop, pd);
        }
        return;
    }
    plain_assign(s, x, n, k, labels, false);
    if (size_min <= 0) return;
    count_labels(s, n, k, labels);
    if (tid == 0) s.flag = 0;
    __syncthreads();
    if (tid < k && s.counts[tid] < size_min) s.flag = 1;
    __syncthreads();
    if (!s.flag) return;
    for (int i = tid; i < n; i += kWinThreads) labels[i] = -1;
    if (tid < k) s.room[tid] = size_min;
    __syncthreads();
    capacity_rounds(s, x, n, k, labels, prop, pd);
    plain_assign(s, x, n, k, labels, true);
}

__global__ void __launch_bounds__(kWinThreads, 1)
kmeans_window_kernel(const float* __restrict__ feats, const long long* __restrict__ offsets,
                     const int* __restrict__ ks, int kmax, int size_min, int size_max, int max_iter,
                     double tol, int* __restrict__ labels_all, float* __restrict__ centroids,
                     int* __restrict__ n_iter, int* __restrict__ prop_all, float* __restrict__ pd_all) {
    __shared__ WinShared s;
    const int w = blockIdx.x, tid = threadIdx.x;
    const long long off = offsets[w];
    const int n = (int)(offsets[w + 1] - off);
    const int k = ks[w];
    const float* x = feats + 3 * off;
    int* labels = labels_all + off;
    int* prop = prop_all + off;
    float* pd = pd_all + off;
    // a window the constraints cannot be met on (or with more clusters than points): labels -1, n_iter -1, no out-of-range write
    if (k < 1 || k > kmax || k > n || (size_max > 0 && (long long)size_max * k < n) || (size_min > 0 && (long long)size_min * k > n)) {
        for (int i = tid; i < n; i += kWinThreads) labels[i] = -1;
        for (int i = tid; i < kmax * 3; i += kWinThreads) centroids[(long long)w * kmax * 3 + i] = 0.0f;
        if (tid == 0) n_iter[w] = -1;
        return;
    }

    // ---- A. fixed-point moments -> tol_abs ----
    if (tid < 6) s.mom[tid] = 0ull;
    __syncthreads();
    {
        long long a[6] = {0, 0, 0, 0, 0, 0};
        for (int i = tid; i < n; i += kWinThreads) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                double v = (double)x[3 * i + d];
                a[d] += __double2ll_rn(__dmul_rn(v, kFix));
                a[3 + d] += __double2ll_rn(__dmul_rn(__dmul_rn(v, v), kFix));
            }
        }
#pragma unroll
        for (int d = 0; d < 6; ++d) atomicAdd(&s.mom[d], (unsigned long long)a[d]);
    }
    __syncthreads();
    if (tid == 0) {
        double acc = 0.0;
        const double dn = (double)n;
        for (int d = 0; d < 3; ++d) {
            double m1 = __ddiv_rn(__ddiv_rn(__ll2double_rn((long long)s.mom[d]), kFix), dn);
            double m2 = __ddiv_rn(__ddiv_rn(__ll2double_rn((long long)s.mom[3 + d]), kFix), dn);
            acc = __dadd_rn(acc, __dsub_rn(m2, __dmul_rn(m1, m1)));
        }
        s.tol_abs = __dmul_rn(__ddiv_rn(acc, 3.0), tol);
    }

    // ---- B. init: farthest-point sampling of k rows, start row 0 (pd[] = running min distance) ----
    for (int i = tid; i < n; i += kWinThreads) pd[i] = (i == 0) ? -1.0f : INFINITY;
    int last = 0;
    if (tid < 3) s.cent[tid] = x[tid];
    __syncthreads();
    for (int c = 1; c < k; ++c) {
        const float lx = x[3 * last], ly = x[3 * last + 1], lz = x[3 * last + 2];
        float bd = -1.0f;
        unsigned bi = 0xffffffffu;
        for (int i = tid; i < n; i += kWinThreads) {
            float d = pd[i];
            float nd = sqd3(lx, ly, lz, x[3 * i], x[3 * i + 1], x[3 * i + 2]);
            d = (nd < d) ? nd : d;
            pd[i] = d;
            if (d > bd) { bd = d; bi = (unsigned)i; }
        }
        block_argmax(s, __float_as_int(bd), bi, last);
        if (tid == 0) pd[last] = -1.0f;
        if (tid < 3) s.cent[3 * c + tid] = x[3 * last + tid];
        __syncthreads();
    }

    // ---- C. Lloyd iterations with the size constraint ----
    int it = 0;
    for (it = 1; it <= max_iter; ++it) {
        constrained_assign(s, x, n, k, size_min, size_max, labels, prop, pd);
        for (int i = tid; i < k * 3; i += kWinThreads) s.sums[i] = 0ull;
        if (tid < k) s.counts[tid] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += kWinThreads) {
            int l = labels[i];
            if (l < 0) continue;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                long long q = __double2ll_rn(__dmul_rn((double)x[3 * i + d], kFix));
                atomicAdd(&s.sums[3 * l + d], (unsigned long long)q);
            }
            atomicAdd(&s.counts[l], 1);
        }
        __syncthreads();
        if (tid == 0) {
            double shift = 0.0;
            for (int j = 0; j < k; ++j) {
                for (int d = 0; d < 3; ++d) {
                    float co = s.cent[3 * j + d];
                    float cn = co;
                    if (s.counts[j] > 0)
                        cn = __double2float_rn(__ddiv_rn(__ddiv_rn(__ll2double_rn((long long)s.sums[3 * j + d]), kFix),
                                                         (double)s.counts[j]));
                    double t = __dsub_rn((double)cn, (double)co);
                    shift = __dadd_rn(shift, __dmul_rn(t, t));
                    s.cent[3 * j + d] = cn;
                }
            }
            s.flag = (shift <= s.tol_abs) ? 1 : 0;
        }
        __syncthreads();
        if (s.flag) break;
    }
    if (it > max_iter) it = max_iter;
    __syncthreads();
    // ---- D. final labels with the final centroids ----
    constrained_assign(s, x, n, k, size_min, size_max, labels, prop, pd);
    for (int i = tid; i < kmax * 3; i += kWinThreads)
        centroids[(long long)w * kmax * 3 + i] = (i < k * 3) ? s.cent[i] : 0.0f;
    if (tid == 0) n_iter[w] = it;
}

// ---------------------------------------------------------------------------------------------
// Stable regroup of one window per CTA: order[] = rows sorted by (label, original index),
// counts per label and the (mean x, mean y) of every group (fixed-point sums).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kWinThreads, 1)
kmeans_regroup_kernel(const int* __restrict__ labels_all, const long long* __restrict__ offsets,
                      const int* __restrict__ ks, int kmax, const float* __restrict__ pc,
                      long long row_stride, long long* __restrict__ order, int* __restrict__ counts,
                      float* __restrict__ xy_mean) {
    __shared__ int s_cnt[kKMax];
    __shared__ int s_base[kKMax];
    __shared__ int s_run[kKMax];
    __shared__ int s_wcnt[32 * kKMax];
    __shared__ unsigned long long s_sum[kKMax * 2];
    const int w = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long off = offsets[w];
    const int n = (int)(offsets[w + 1] - off);
    const int k = ks[w];
    const int* labels = labels_all + off;
    if (tid < kKMax) { s_cnt[tid] = 0; s_run[tid] = 0; }
    if (tid < 2 * kKMax) s_sum[tid] = 0ull;
    __syncthreads();
    for (int i = tid; i < n; i += kWinThreads) {
        int l = labels[i];
        if (l < 0 || l >= kKMax) continue;                 // unassigned (infeasible window): belongs to no group
        atomicAdd(&s_cnt[l], 1);
        if (pc) {
            const float* p = pc + (off + i) * row_stride;
            atomicAdd(&s_sum[2 * l], (unsigned long long)__double2ll_rn(__dmul_rn((double)p[0], kFix)));
            atomicAdd(&s_sum[2 * l + 1], (unsigned long long)__double2ll_rn(__dmul_rn((double)p[1], kFix)));
        }
    }
    __syncthreads();
    if (tid == 0) {
        int acc = 0;
        for (int j = 0; j < kKMax; ++j) { s_base[j] = acc; acc += s_cnt[j]; }
    }
    if (tid < kmax) {
        counts[(long long)w * kmax + tid] = (tid < k) ? s_cnt[tid] : 0;
        if (xy_mean) {
            float mx = 0.f, my = 0.f;
            if (tid < k && s_cnt[tid] > 0) {
                mx = __double2float_rn(__ddiv_rn(__ddiv_rn(__ll2double_rn((long long)s_sum[2 * tid]), kFix), (double)s_cnt[tid]));
                my = __double2float_rn(__ddiv_rn(__ddiv_rn(__ll2double_rn((long long)s_sum[2 * tid + 1]), kFix), (double)s_cnt[tid]));
            }
            xy_mean[((long long)w * kmax + tid) * 2] = mx;
            xy_mean[((long long)w * kmax + tid) * 2 + 1] = my;
        }
    }
    __syncthreads();
    for (int base = 0; base < n; base += kWinThreads) {
        const int i = base + tid;
        int l = (i < n) ? labels[i] : -1;
        if (l >= kKMax) l = -1;
        for (int j = tid; j < 32 * kKMax; j += kWinThreads) s_wcnt[j] = 0;
        __syncthreads();
        // rank among equal labels inside the warp, in lane (= index) order
        unsigned peers = __match_any_sync(0xffffffffu, l);
        int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
        if (l >= 0 && rank_in_warp == 0) s_wcnt[warp * kKMax + l] = __popc(peers);
        __syncthreads();
        if (tid < kKMax) {   // exclusive prefix over warps for label `tid`
            int acc = s_run[tid];
            for (int ww = 0; ww < 32; ++ww) {
                int c = s_wcnt[ww * kKMax + tid];
                s_wcnt[ww * kKMax + tid] = acc;
                acc += c;
            }
            s_run[tid] = acc;
        }
        __syncthreads();
        if (l >= 0) order[off + s_base[l] + s_wcnt[warp * kKMax + l] + rank_in_warp] = off + i;
        __syncthreads();
    }
}

}  // namespace
// on-chip variant (kmeans_window.cu): 1 = launched, 0 = window too large, < 0 = error
int kmeans_window_fast_cap(bool with_x, int kmax);
int kmeans_window_fast_try(const float* feats, const long long* offsets, const int* ks, long long W, long long max_window_points, int kmax,
                           int size_min, int size_max, int max_iter, double tol, int n_init, int* labels, float* centroids, int* n_iter,
                           cudaStream_t st);
}  // namespace amp

extern "C" {

int amp_kmeans_assign_f32(const float* feats, const float* centroids, int64_t n, int32_t k,
                          int32_t* labels, float* min_d2, void* stream) {
    if (!feats || !centroids || !labels) return amp::fail(AMP_E_BADARG, "kmeans_assign: null pointer");
    if (k < 1 || k > 64) return amp::fail(AMP_E_BADARG, "kmeans_assign: k=%d not in [1,64]", k);
    if (n < 0) return amp::fail(AMP_E_BADARG, "kmeans_assign: n < 0");
    if (n == 0) return AMP_OK;
    if (((uintptr_t)feats & 15) || ((uintptr_t)labels & 15) || (min_d2 && ((uintptr_t)min_d2 & 15)))
        return amp::fail(AMP_E_BADARG, "kmeans_assign: feats/labels/min_d2 must be 16-byte aligned");
    long long quads = (n + 3) / 4;
    long long blocks = (quads + amp::kAssignThreads - 1) / amp::kAssignThreads;
    long long cap = (long long)amp::kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    cudaStream_t st = (cudaStream_t)stream;
    switch (k) {
#define AMP_ASSIGN_K(K) case K: amp::launch_assign_small_k<K>(feats, centroids, n, labels, min_d2, (unsigned)blocks, st); break;
        AMP_ASSIGN_K(1) AMP_ASSIGN_K(2) AMP_ASSIGN_K(3) AMP_ASSIGN_K(4) AMP_ASSIGN_K(5) AMP_ASSIGN_K(6) AMP_ASSIGN_K(7) AMP_ASSIGN_K(8)
        AMP_ASSIGN_K(9) AMP_ASSIGN_K(10) AMP_ASSIGN_K(11) AMP_ASSIGN_K(12) AMP_ASSIGN_K(13) AMP_ASSIGN_K(14) AMP_ASSIGN_K(15) AMP_ASSIGN_K(16)
#undef AMP_ASSIGN_K
        default:
            amp::kmeans_assign_kernel<<<(unsigned)blocks, amp::kAssignThreads, 0, st>>>(feats, centroids, n, k, labels, min_d2);
    }
    amp::count_launch();
    return amp::check_launch("kmeans_assign");
}

int amp_kmeans_gather_feats_f32(const float* pc, int64_t n, int64_t row_stride, int32_t c0, int32_t c1,
                                int32_t c2, float* feats, void* stream) {
    if (!pc || !feats) return amp::fail(AMP_E_BADARG, "gather_feats: null pointer");
    if (c0 < 0 || c1 < 0 || c2 < 0 || c0 >= row_stride || c1 >= row_stride || c2 >= row_stride)
        return amp::fail(AMP_E_BADARG, "gather_feats: column out of range");
    if (n <= 0) return AMP_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > amp::kNumSMs * 8) blocks = amp::kNumSMs * 8;
    amp::gather_feats_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pc, n, row_stride, c0, c1, c2, feats);
    amp::count_launch();
    return amp::check_launch("gather_feats");
}

size_t amp_kmeans_workspace_bytes(int64_t total_points, int64_t W, int32_t kmax) {
    (void)W; (void)kmax;
    return (size_t)total_points * 8 + 256;   // prop int32 + pd float32 per point
}

int amp_kmeans_constrained_f32(const float* feats, const int64_t* offsets, const int32_t* ks, int64_t W,
                               int64_t total_points, int64_t max_window_points, int32_t kmax,
                               int32_t size_min, int32_t size_max, int32_t max_iter, double tol, int32_t n_init,
                               int32_t* labels, float* centroids, int32_t* n_iter, void* workspace,
                               size_t workspace_bytes, void* stream) {
    if (!feats || !offsets || !ks || !labels || !centroids || !n_iter)
        return amp::fail(AMP_E_BADARG, "kmeans_constrained: null pointer");
    if (kmax < 1 || kmax > amp::kKMax) return amp::fail(AMP_E_BADARG, "kmeans_constrained: kmax=%d not in [1,%d]", kmax, amp::kKMax);
    if (W < 1 || total_points < 1) return amp::fail(AMP_E_BADARG, "kmeans_constrained: empty input");
    if (max_window_points >= (1LL << 31)) return amp::fail(AMP_E_BADARG, "kmeans_constrained: window too large");
    if (max_iter < 1 || n_init < 1 || n_init > 64) return amp::fail(AMP_E_BADARG, "kmeans_constrained: max_iter >= 1 and 1 <= n_init <= 64");
    if (size_min < 0 || size_max < 0 || (size_max > 0 && size_min > size_max))
        return amp::fail(AMP_E_BADARG, "kmeans_constrained: size_min / size_max must be >= 0 and size_min <= size_max");
    size_t need = amp_kmeans_workspace_bytes(total_points, W, kmax);
    if (!workspace || workspace_bytes < need)
        return amp::fail(AMP_E_WORKSPACE, "kmeans_constrained: workspace %zu < %zu", workspace_bytes, need);
    {
        const int rc = amp::kmeans_window_fast_try(feats, reinterpret_cast<const long long*>(offsets), ks, W, max_window_points, kmax, size_min,
                                                   size_max, max_iter, tol, n_init, labels, centroids, n_iter, (cudaStream_t)stream);
        if (rc != 0) return rc < 0 ? rc : AMP_OK;
    }
    if (n_init > 1)
        return amp::fail(AMP_E_BADARG, "kmeans_constrained: n_init > 1 needs windows of at most %d points (the on-chip kernel)",
                         amp::kmeans_window_fast_cap(false, kmax));
    int* prop = reinterpret_cast<int*>(workspace);
    float* pd = reinterpret_cast<float*>(prop + total_points);
    amp::kmeans_window_kernel<<<(unsigned)W, amp::kWinThreads, 0, (cudaStream_t)stream>>>(
        feats, reinterpret_cast<const long long*>(offsets), ks, kmax, size_min, size_max, max_iter, tol,
        labels, centroids, n_iter, prop, pd);
    amp::count_launch();
    return amp::check_launch("kmeans_constrained");
}

int amp_kmeans_regroup(const int32_t* labels, const int64_t* offsets, const int32_t* ks, int64_t W,
                       int32_t kmax, const float* pc, int64_t row_stride, int64_t* order, int32_t* counts,
                       float* xy_mean, void* stream) {
    if (!labels || !offsets || !ks || !order || !counts)
        return amp::fail(AMP_E_BADARG, "kmeans_regroup: null pointer");
    if (kmax < 1 || kmax > amp::kKMax) return amp::fail(AMP_E_BADARG, "kmeans_regroup: kmax");
    if (W < 1) return amp::fail(AMP_E_BADARG, "kmeans_regroup: W < 1");
    if (xy_mean && !pc) return amp::fail(AMP_E_BADARG, "kmeans_regroup: xy_mean needs pc");
    amp::kmeans_regroup_kernel<<<(unsigned)W, amp::kWinThreads, 0, (cudaStream_t)stream>>>(
        labels, reinterpret_cast<const long long*>(offsets), ks, kmax, pc, row_stride,
        reinterpret_cast<long long*>(order), counts, xy_mean);
    amp::count_launch();
    return amp::check_launch("kmeans_regroup");
}

}  // extern "C"
