// Constrained k-means of one window per CTA with the whole working state on chip (the latency-bound stage of configs[3]).
//
// Same algorithm, same arithmetic and therefore the same labels / centroids as kmeans_window_kernel (kmeans.cu) and
// oracle/kmeans_oracle.py -- FPS init, capacity rounds with an exact (d^2, index) radix select, order-independent
// fixed-point centroid sums -- restructured for latency: a window runs ~10 Lloyd iterations x ~k capacity rounds x
// ~6 block-wide passes one after the other, so what matters is the cost of ONE pass:
//   * running min-distance, proposal and label of every point live in shared memory (6 B / point), and the coordinates too
//     when the window is small enough (18 B / point); the previous kernel kept all of it in global memory;
//   * the digit search of a radix pass is a warp-parallel scan of the 256 bins (it was a serial loop of one thread per
//     cluster: ~6 k cycles per pass), the histogram is warp-aggregated (most proposals of a round share their top bytes),
//     and the passes over the index half of the key run only when two candidates tie at the cut (4 passes instead of 8);
//   * centroid sums: every thread sums its own points per cluster in registers, one warp reduction and one shared-memory
//     atomic per (warp, cluster, dimension) instead of one 64-bit shared-memory atomic (a CAS loop) per point and dimension.
// Replaces the solver call of data_proc/3_kmeans.py:78-82 / utils/utils.py:500-505 (see kmeans.cu).
#include <stdint.h>
#include <stdlib.h>

#include "amp_common.cuh"

namespace amp {
namespace {

constexpr int kKMax = 32;
constexpr double kFix = 4294967296.0;  // 2^32
constexpr int kT = 1024;

__device__ __forceinline__ float sqd3(float x0, float x1, float x2, float c0, float c1, float c2) {
    float d0 = __fsub_rn(x0, c0), d1 = __fsub_rn(x1, c1), d2 = __fsub_rn(x2, c2);
    return __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
}

// (The bins of a histogram row are read with scalar loads: a 16-byte vector load of `hist`, which the compiler emits when it
// may assume the dynamic shared-memory window is 16-byte aligned, faulted with "misaligned address" on B200.)
struct WS {
    float cent[kKMax * 3];
    unsigned long long sums[kKMax * 3];
    unsigned long long mom[6];
    int counts[kKMax];
    int room[kKMax];
    int nprop[kKMax];
    int over[kKMax];
    int done[kKMax];
    unsigned long long prefix[kKMax];
    int rank[kKMax];
    int red_bits[32];
    unsigned red_idx[32];
    int n_todo, n_open, flag, flag2, pick, best_it;
    float best_cent[kKMax * 3];
    unsigned long long inertia;
    long long best_inertia;
    double tol_abs, shift;
};

struct State {                      // per-point state in shared memory
    float* pd; signed char* lab; signed char* prop; const float* x;   // x: shared copy (XS) or the global rows
    unsigned* hist;                 // [kmax][256] radix histogram rows (sized by the launch's largest k, not by kKMax: 27 KB
                                    // more for the points, which keeps 10 k-point windows' coordinates on chip)
};

__device__ __forceinline__ long long warp_sum_ll(long long v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void block_argmax(WS& s, int kb, unsigned idx, int& out_idx) {
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

// warp-aggregated shared-memory counter increment: one atomic per distinct address in the warp
__device__ __forceinline__ void agg_inc(unsigned* base, int slot, bool active) {
    const unsigned peers = __match_any_sync(0xffffffffu, active ? slot : -1);
    if (active && (__ffs(peers) - 1) == (int)(threadIdx.x & 31)) atomicAdd(base + slot, (unsigned)__popc(peers));
}

// Capacity rounds (oracle/kmeans_oracle.py::_capacity_rounds). lab < 0 = unassigned.
__device__ void capacity_rounds(WS& s, const State& st, int n, int k) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    while (true) {
        if (tid == 0) { s.n_todo = 0; s.n_open = 0; }
        if (tid < k) s.nprop[tid] = 0;
        __syncthreads();
        if (tid < k && s.room[tid] > 0) atomicAdd(&s.n_open, 1);
        // proposals: nearest cluster that still has room
        int local_todo = 0;
        for (int i0 = 0; i0 < n; i0 += kT) {
            const int i = i0 + tid;
            const bool todo = i < n && st.lab[i] < 0;
            int bj = -1;
            if (todo) {
                ++local_todo;
                const float x0 = st.x[3 * i], x1 = st.x[3 * i + 1], x2 = st.x[3 * i + 2];
                float best = INFINITY;
                for (int j = 0; j < k; ++j) {
                    if (s.room[j] <= 0) continue;
                    const float d = sqd3(x0, x1, x2, s.cent[3 * j], s.cent[3 * j + 1], s.cent[3 * j + 2]);
                    if (d < best || bj < 0) { best = d; bj = j; }
                }
                st.prop[i] = (signed char)bj;
                st.pd[i] = best;
            }
            agg_inc(reinterpret_cast<unsigned*>(s.nprop), bj, todo && bj >= 0);
        }
        local_todo = __reduce_add_sync(0xffffffffu, local_todo);
        if (lane == 0 && local_todo) atomicAdd(&s.n_todo, local_todo);
        __syncthreads();
        if (s.n_todo == 0 || s.n_open == 0) break;   // block-uniform
        if (tid < k) {
            s.over[tid] = (s.room[tid] > 0 && s.nprop[tid] > s.room[tid]) ? 1 : 0;
            s.prefix[tid] = 0ull;
            s.rank[tid] = s.room[tid];
            s.done[tid] = 0;
        }
        if (tid == 0) s.flag = 0;
        __syncthreads();
        if (tid < k && s.over[tid]) s.flag = 1;
        __syncthreads();
        if (s.flag) {
            // exact room[j]-th smallest 64-bit key (d^2 bits << 32 | index), 8 bits per pass; the four passes over the index
            // half run only if some cluster's cut falls inside a group of equal distances
            for (int pass = 0; pass < 8; ++pass) {
                const int shift = 56 - 8 * pass;
                for (int i = tid; i < k * 256; i += kT) st.hist[i] = 0u;
                if (tid == 0) s.flag2 = 0;
                __syncthreads();
                for (int i0 = 0; i0 < n; i0 += kT) {
                    const int i = i0 + tid;
                    bool act = false;
                    int slot = 0;
                    if (i < n && st.lab[i] < 0) {
                        const int j = st.prop[i];
                        if (j >= 0 && s.over[j] && !s.done[j]) {
                            const unsigned long long key = ((unsigned long long)__float_as_uint(st.pd[i]) << 32) | (unsigned)i;
                            act = (pass == 0) || ((key >> (shift + 8)) == s.prefix[j]);
                            slot = j * 256 + (int)((key >> shift) & 255ull);
                        }
                    }
                    // top byte (sign + exponent): a handful of distinct digits per warp -> one aggregated atomic per digit;
                    // lower bytes are spread over the bins: plain atomics (match_any loops once per distinct value)
                    if (pass == 0) agg_inc(st.hist, slot, act);
                    else if (act) atomicAdd(st.hist + slot, 1u);
                }
                __syncthreads();
                if (warp < k && s.over[warp] && !s.done[warp]) {        // one warp per cluster: lane owns 8 consecutive bins
                    const volatile unsigned* h = st.hist + warp * 256 + lane * 8;     // (scalar loads: see the note at WS)
                    int c[8], tot = 0;
#pragma unroll
                    for (int b = 0; b < 8; ++b) { c[b] = (int)h[b]; tot += c[b]; }
                    int incl = tot;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
                    const int before = incl - tot, rk = s.rank[warp];
                    // the lane whose bins hold the rk-th element: before < rk <= before + tot (first such lane)
                    const bool mine = before < rk && rk <= incl;
                    if (mine) {
                        // first bin b with before + c[0..b] >= rk; selects only (no dynamically indexed local array)
                        int cum = before, dgt = 7, cdg = c[7], run = before;
                        bool found = false;
#pragma unroll
                        for (int b = 0; b < 8; ++b) {
                            const bool hit = !found && run + c[b] >= rk;
                            if (hit) { dgt = b; cum = run; cdg = c[b]; }
                            found = found || hit;
                            run += c[b];
                        }
                        s.prefix[warp] = (s.prefix[warp] << 8) | (unsigned long long)(lane * 8 + dgt);
                        s.rank[warp] = rk - cum;
                        if (pass == 3) {                                 // distance fully determined
                            if (rk - cum == cdg) {                       // the whole group of equal distances fits: no tie at the cut
                                s.prefix[warp] = (s.prefix[warp] << 32) | 0xffffffffull;
                                s.done[warp] = 1;
                            } else {
                                s.flag2 = 1;
                            }
                        }
                    }
                }
                __syncthreads();
                if (pass == 3 && !s.flag2) break;                        // block-uniform
            }
        }
        // accept
        for (int i = tid; i < n; i += kT) {
            if (st.lab[i] >= 0) continue;
            const int j = st.prop[i];
            if (j < 0) continue;
            bool ok = true;
            if (s.over[j]) {
                const unsigned long long key = ((unsigned long long)__float_as_uint(st.pd[i]) << 32) | (unsigned)i;
                ok = key <= s.prefix[j];
            }
            if (ok) st.lab[i] = (signed char)j;
        }
        __syncthreads();
        if (tid < k) s.room[tid] -= (s.nprop[tid] < s.room[tid]) ? s.nprop[tid] : s.room[tid];
        __syncthreads();
    }
    __syncthreads();
}

__device__ void count_labels(WS& s, const State& st, int n, int k) {
    const int tid = threadIdx.x;
    if (tid < k) s.counts[tid] = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += kT) {
        const int i = i0 + tid;
        const int l = i < n ? st.lab[i] : -1;
        agg_inc(reinterpret_cast<unsigned*>(s.counts), l, l >= 0);
    }
    __syncthreads();
}

__device__ void plain_assign(WS& s, const State& st, int n, int k, bool only_unassigned) {
    for (int i = threadIdx.x; i < n; i += kT) {
        if (only_unassigned && st.lab[i] >= 0) continue;
        const float x0 = st.x[3 * i], x1 = st.x[3 * i + 1], x2 = st.x[3 * i + 2];
        float best = INFINITY;
        int bj = 0;
        for (int j = 0; j < k; ++j) {
            const float d = sqd3(x0, x1, x2, s.cent[3 * j], s.cent[3 * j + 1], s.cent[3 * j + 2]);
            if (d < best) { best = d; bj = j; }
        }
        st.lab[i] = (signed char)bj;
    }
    __syncthreads();
}

__device__ void constrained_assign(WS& s, const State& st, int n, int k, int size_min, int size_max) {
    const int tid = threadIdx.x;
    if (size_max > 0) {
        for (int i = tid; i < n; i += kT) st.lab[i] = -1;
        if (size_min > 0 && size_min < size_max) {
            if (tid < k) s.room[tid] = size_min;
            __syncthreads();
            capacity_rounds(s, st, n, k);
            count_labels(s, st, n, k);
            if (tid < k) s.room[tid] = size_max - s.counts[tid];
            __syncthreads();
            capacity_rounds(s, st, n, k);
        } else {
            if (tid < k) s.room[tid] = size_max;
            __syncthreads();
            capacity_rounds(s, st, n, k);
        }
        return;
    }
    plain_assign(s, st, n, k, false);
    if (size_min <= 0) return;
    count_labels(s, st, n, k);
    if (tid == 0) s.flag = 0;
    __syncthreads();
    if (tid < k && s.counts[tid] < size_min) s.flag = 1;
    __syncthreads();
    if (!s.flag) return;
    for (int i = tid; i < n; i += kT) st.lab[i] = -1;
    if (tid < k) s.room[tid] = size_min;
    __syncthreads();
    capacity_rounds(s, st, n, k);
    plain_assign(s, st, n, k, true);
}

template <bool XS>
__global__ void __launch_bounds__(kT, 1)
kmeans_window_fast_kernel(const float* __restrict__ feats, const long long* __restrict__ offsets, const int* __restrict__ ks, int kmax,
                          int size_min, int size_max, int max_iter, double tol, int cap, int* __restrict__ labels_all,
                          float* __restrict__ centroids, int* __restrict__ n_iter, int n_init) {
    extern __shared__ __align__(8) unsigned char smem[];
    WS& s = *reinterpret_cast<WS*>(smem);
    unsigned* s_hist = reinterpret_cast<unsigned*>(smem + ((sizeof(WS) + 15) & ~(size_t)15));
    float* s_pd = reinterpret_cast<float*>(s_hist + (size_t)kmax * 256);
    float* s_x = s_pd + cap;                                   // [3 * cap] when XS
    signed char* s_lab = reinterpret_cast<signed char*>(XS ? s_x + 3 * (size_t)cap : s_x);
    signed char* s_prop = s_lab + cap;
    signed char* s_best = s_prop + cap;                        // labels of the best restart so far (n_init > 1)
    const int w = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const long long off = offsets[w];
    const int n = (int)(offsets[w + 1] - off);
    const int k = ks[w];
    const float* gx = feats + 3 * off;
    int* labels = labels_all + off;
    // a window the constraints cannot be met on (or with more clusters than points): labels -1, n_iter -1, no out-of-range write
    if (k < 1 || k > kmax || k > n || (size_max > 0 && (long long)size_max * k < n) || (size_min > 0 && (long long)size_min * k > n)) {
        for (int i = tid; i < n; i += kT) labels[i] = -1;
        for (int i = tid; i < kmax * 3; i += kT) centroids[(long long)w * kmax * 3 + i] = 0.0f;
        if (tid == 0) n_iter[w] = -1;
        return;
    }
    if (XS) for (int i = tid; i < 3 * n; i += kT) s_x[i] = gx[i];
    State st{s_pd, s_lab, s_prop, XS ? s_x : gx, s_hist};

    // ---- A. fixed-point moments -> tol_abs ----
    if (tid < 6) s.mom[tid] = 0ull;
    __syncthreads();
    {
        long long a[6] = {0, 0, 0, 0, 0, 0};
        for (int i = tid; i < n; i += kT) {
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const double v = (double)st.x[3 * i + d];
                a[d] += __double2ll_rn(__dmul_rn(v, kFix));
                a[3 + d] += __double2ll_rn(__dmul_rn(__dmul_rn(v, v), kFix));
            }
        }
#pragma unroll
        for (int d = 0; d < 6; ++d) {
            const long long t = warp_sum_ll(a[d]);
            if (lane == 0) atomicAdd(&s.mom[d], (unsigned long long)t);
        }
    }
    __syncthreads();
    if (tid == 0) {
        double acc = 0.0;
        const double dn = (double)n;
        for (int d = 0; d < 3; ++d) {
            const double m1 = __ddiv_rn(__ddiv_rn(__ll2double_rn((long long)s.mom[d]), kFix), dn);
            const double m2 = __ddiv_rn(__ddiv_rn(__ll2double_rn((long long)s.mom[3 + d]), kFix), dn);
            acc = __dadd_rn(acc, __dsub_rn(m2, __dmul_rn(m1, m1)));
        }
        s.tol_abs = __dmul_rn(__ddiv_rn(acc, 3.0), tol);
    }

    if (n_init < 1) n_init = 1;
    for (int restart = 0; restart < n_init; ++restart) {
    // ---- B. init: farthest-point sampling of k rows from the restart's start row (pd[] = running min distance) ----
    const int start = (int)(((long long)restart * n) / n_init);
    for (int i = tid; i < n; i += kT) s_pd[i] = (i == start) ? -1.0f : INFINITY;
    int last = start;
    if (tid < 3) s.cent[tid] = st.x[3 * start + tid];
    __syncthreads();
    for (int c = 1; c < k; ++c) {
        const float lx = st.x[3 * last], ly = st.x[3 * last + 1], lz = st.x[3 * last + 2];
        float bd = -1.0f;
        unsigned bi = 0xffffffffu;
        for (int i = tid; i < n; i += kT) {
            float d = s_pd[i];
            const float nd = sqd3(lx, ly, lz, st.x[3 * i], st.x[3 * i + 1], st.x[3 * i + 2]);
            d = (nd < d) ? nd : d;
            s_pd[i] = d;
            if (d > bd) { bd = d; bi = (unsigned)i; }
        }
        block_argmax(s, __float_as_int(bd), bi, last);
        if (tid == 0) s_pd[last] = -1.0f;
        if (tid < 3) s.cent[3 * c + tid] = st.x[3 * last + tid];
        __syncthreads();
    }

    // ---- C. Lloyd iterations with the size constraint ----
    int it = 0;
    for (it = 1; it <= max_iter; ++it) {
        constrained_assign(s, st, n, k, size_min, size_max);
        for (int i = tid; i < k * 3; i += kT) s.sums[i] = 0ull;
        if (tid < k) s.counts[tid] = 0;
        __syncthreads();
        // per-thread sums of its own points, cluster by cluster; one warp reduction + one atomic per (warp, cluster, dim)
        for (int j = 0; j < k; ++j) {
            long long a0 = 0, a1 = 0, a2 = 0;
            int cn = 0;
            for (int i = tid; i < n; i += kT) {
                if (s_lab[i] == j) {
                    a0 += __double2ll_rn(__dmul_rn((double)st.x[3 * i], kFix));
                    a1 += __double2ll_rn(__dmul_rn((double)st.x[3 * i + 1], kFix));
                    a2 += __double2ll_rn(__dmul_rn((double)st.x[3 * i + 2], kFix));
                    ++cn;
                }
            }
            a0 = warp_sum_ll(a0); a1 = warp_sum_ll(a1); a2 = warp_sum_ll(a2);
            cn = __reduce_add_sync(0xffffffffu, cn);
            if (lane == 0 && cn) {
                atomicAdd(&s.sums[3 * j], (unsigned long long)a0);
                atomicAdd(&s.sums[3 * j + 1], (unsigned long long)a1);
                atomicAdd(&s.sums[3 * j + 2], (unsigned long long)a2);
                atomicAdd(&s.counts[j], cn);
            }
        }
        __syncthreads();
        if (tid == 0) {
            double shift = 0.0;
            for (int j = 0; j < k; ++j) {
                for (int d = 0; d < 3; ++d) {
                    const float co = s.cent[3 * j + d];
                    float cn = co;
                    if (s.counts[j] > 0)
                        cn = __double2float_rn(__ddiv_rn(__ddiv_rn(__ll2double_rn((long long)s.sums[3 * j + d]), kFix), (double)s.counts[j]));
                    const double t = __dsub_rn((double)cn, (double)co);
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
    constrained_assign(s, st, n, k, size_min, size_max);
    if (n_init == 1) {
        for (int i = tid; i < n; i += kT) labels[i] = (int)s_lab[i];
        for (int i = tid; i < kmax * 3; i += kT) centroids[(long long)w * kmax * 3 + i] = (i < k * 3) ? s.cent[i] : 0.0f;
        if (tid == 0) n_iter[w] = it;
        return;
    }
    // ---- E. restarts: keep the run with the smallest fixed-point inertia (earliest on a tie) ----
    if (tid == 0) s.inertia = 0ull;
    __syncthreads();
    {
        long long a = 0;
        for (int i = tid; i < n; i += kT) {
            const int l = s_lab[i];
            const float d = sqd3(st.x[3 * i], st.x[3 * i + 1], st.x[3 * i + 2], s.cent[3 * l], s.cent[3 * l + 1], s.cent[3 * l + 2]);
            a += __double2ll_rn(__dmul_rn((double)d, kFix));
        }
        a = warp_sum_ll(a);
        if (lane == 0) atomicAdd(&s.inertia, (unsigned long long)a);
    }
    __syncthreads();
    const bool better = restart == 0 || (long long)s.inertia < s.best_inertia;     // block-uniform
    __syncthreads();
    if (better) {
        for (int i = tid; i < n; i += kT) s_best[i] = s_lab[i];
        for (int i = tid; i < k * 3; i += kT) s.best_cent[i] = s.cent[i];
        if (tid == 0) { s.best_inertia = (long long)s.inertia; s.best_it = it; }
    }
    __syncthreads();
    }   // restarts
    for (int i = tid; i < n; i += kT) labels[i] = (int)s_best[i];
    for (int i = tid; i < kmax * 3; i += kT) centroids[(long long)w * kmax * 3 + i] = (i < k * 3) ? s.best_cent[i] : 0.0f;
    if (tid == 0) n_iter[w] = s.best_it;
}

constexpr size_t kMaxSmem = 232448;
inline size_t ws_bytes(int kmax) { return ((sizeof(WS) + 15) & ~(size_t)15) + (size_t)kmax * 256 * sizeof(unsigned); }

}  // namespace

// Largest window (points) the on-chip kernel takes: with the coordinates in shared memory, and without.
int kmeans_window_fast_cap(bool with_x, int kmax) {
    const size_t per = with_x ? 19 : 7;
    return (int)(((kMaxSmem - ws_bytes(kmax) - 64) / per) & ~(size_t)15);
}

// 1 = launched, 0 = window too large for the on-chip kernel (the caller runs kmeans_window_kernel), < 0 = error
int kmeans_window_fast_try(const float* feats, const long long* offsets, const int* ks, long long W, long long max_window_points, int kmax,
                           int size_min, int size_max, int max_iter, double tol, int n_init, int* labels, float* centroids, int* n_iter,
                           cudaStream_t st) {
    if (path_disabled("kmeans_fast")) return 0;
    if (kmax < 1 || kmax > kKMax) return 0;
    const bool xs = max_window_points <= kmeans_window_fast_cap(true, kmax);
    if (!xs && max_window_points > kmeans_window_fast_cap(false, kmax)) return 0;
    const int cap = (int)((max_window_points + 15) & ~15LL);
    const size_t smem = ws_bytes(kmax) + (size_t)cap * (xs ? 19 : 7) + 64;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(kmeans_window_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(kmeans_window_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem);
        if (e != cudaSuccess) return fail(AMP_E_CUDA, "kmeans_window_fast: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        attr_set = true;
    }
    if (xs)
        kmeans_window_fast_kernel<true><<<(unsigned)W, kT, smem, st>>>(feats, offsets, ks, kmax, size_min, size_max, max_iter, tol, cap, labels,
                                                                        centroids, n_iter, n_init);
    else
        kmeans_window_fast_kernel<false><<<(unsigned)W, kT, smem, st>>>(feats, offsets, ks, kmax, size_min, size_max, max_iter, tol, cap, labels,
                                                                         centroids, n_iter, n_init);
    count_launch();
    count_path("kmeans_fast");
    const int rc = check_launch("kmeans_window_fast");
    return rc == AMP_OK ? 1 : rc;
}

}  // namespace amp
