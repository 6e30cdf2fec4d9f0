// Data preparation in front of the block split, float64 like the reference's numpy code, bit-exact:
//
//   amp_minmax_f64            min / max of the tile's x and y                      (1_get_windows_split.py:53-54)
//   amp_window_ids_f64        window id of every point, strict bounds on both sides (1_get_windows_split.py:57-62)
//   amp_window_partition      stable counting sort of the point indices by window id (the boolean-mask gathers of :62-77
//                             keep the original order inside a window)
//   amp_filter_normalize_f64  per window: drop ground / noise classes and HAG outliers, build the 13-column row,
//                             normalise x / y to [-1, 1], HAG / max_z, clip intensity / NIR / NDVI
//                             (2_preprocessing_filter_norm.py:40-104)
//
// All of it is streaming byte / float64 work: one pass over the points per kernel, coalesced, no tensor cores.
// The oracle (oracle/dataprep_oracle.py) is pinned bit-for-bit to the unmodified reference functions.
#include <stdint.h>

#include "amp_common.cuh"

namespace amp {
namespace {

constexpr int kThreads = 256;
constexpr int kSortThreads = 1024, kSortWarps = kSortThreads / 32;

// order-preserving map double -> uint64 (for atomicMin / atomicMax)
__device__ __forceinline__ unsigned long long ord_bits(double v) {
    const unsigned long long u = (unsigned long long)__double_as_longlong(v);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline double ord_value(unsigned long long k) {
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    union { unsigned long long u; double d; } c; c.u = u; return c.d;
#endif
}

// keys[0..3] = ordered bits of min x, max x, min y, max y (initialised by the host wrapper: ~0, 0, ~0, 0)
__global__ void minmax_kernel(const double* __restrict__ x, const double* __restrict__ y, long long n, long long stride,
                              unsigned long long* __restrict__ keys) {
    unsigned long long mnx = ~0ull, mxx = 0ull, mny = ~0ull, mxy = 0ull;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const unsigned long long kx = ord_bits(x[i * stride]), ky = ord_bits(y[i * stride]);
        mnx = min(mnx, kx); mxx = max(mxx, kx); mny = min(mny, ky); mxy = max(mxy, ky);
    }
    for (int o = 16; o >= 1; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(keys + 0, mnx); atomicMax(keys + 1, mxx); atomicMin(keys + 2, mny); atomicMax(keys + 3, mxy);
    }
}
__global__ void minmax_init_kernel(unsigned long long* __restrict__ keys) {
    if (threadIdx.x < 4) keys[threadIdx.x] = (threadIdx.x & 1) ? 0ull : ~0ull;
}
__global__ void minmax_decode_kernel(const unsigned long long* __restrict__ keys, double* __restrict__ out) {
    if (threadIdx.x < 4) out[threadIdx.x] = ord_value(keys[threadIdx.x]);
}

// index of the open interval (lo0 + j * w, lo0 + (j + 1) * w) that holds v, -1 when v lies on a bound or outside [0, n)
__device__ __forceinline__ int open_cell(double v, double lo0, double w, int n) {
    int j = (int)floor((v - lo0) / w);
    if (j < -1 || j > n) return -1;
#pragma unroll
    for (int t = -1; t <= 1; ++t) {                       // the quotient may be one off after rounding: test the neighbours exactly
        const int jj = j + t;
        const double lo = lo0 + (double)jj * w;           // integers: exact
        if (jj >= 0 && jj < n && v > lo && v < lo + w) return jj;
    }
    return -1;
}

__global__ void window_ids_kernel(const double* __restrict__ x, const double* __restrict__ y, long long n, long long stride, double x0, double y0,
                                  double wx, double wy, int nx, int ny, int* __restrict__ ids) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int ix = open_cell(x[i * stride], x0, wx, nx), iy = open_cell(y[i * stride], y0, wy, ny);
        ids[i] = (ix < 0 || iy < 0) ? -1 : iy * nx + ix;
    }
}

// ---- stable LSD radix sort of (key, index) by 8-bit digits: histogram per chunk, scan, ordered scatter ----
__device__ __forceinline__ unsigned int digit_of(int key, int n_bins, int shift) {
    const unsigned int k = key < 0 ? (unsigned int)n_bins : (unsigned int)key;     // dropped points sort behind every window
    return (k >> shift) & 255u;
}

__global__ void radix_hist_kernel(const int* __restrict__ keys, long long n, long long chunk, int n_bins, int shift, int* __restrict__ hist) {
    __shared__ int s_h[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_h[i] = 0;
    __syncthreads();
    const long long b = blockIdx.x * chunk, e = min(n, b + chunk);
    for (long long i = b + threadIdx.x; i < e; i += blockDim.x) atomicAdd(&s_h[digit_of(keys[i], n_bins, shift)], 1);
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[(long long)blockIdx.x * 256 + i] = s_h[i];
}

// base[c][d] = (points with a smaller digit) + (points with digit d in earlier chunks); one CTA of 256 threads
__global__ void radix_scan_kernel(int* __restrict__ hist, int n_chunks) {
    __shared__ long long s_tot[256];
    const int d = threadIdx.x;
    long long t = 0;
    for (int c = 0; c < n_chunks; ++c) t += hist[(long long)c * 256 + d];
    s_tot[d] = t;
    __syncthreads();
    long long before = 0;
    for (int j = 0; j < d; ++j) before += s_tot[j];
    long long run = before;
    for (int c = 0; c < n_chunks; ++c) {
        const int h = hist[(long long)c * 256 + d];
        hist[(long long)c * 256 + d] = (int)run;
        run += h;
    }
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const int* __restrict__ keys_in, const int* __restrict__ vals_in,
                                                                     long long n, long long chunk, int n_bins, int shift,
                                                                     const int* __restrict__ base, int* __restrict__ keys_out,
                                                                     int* __restrict__ vals_out) {
    __shared__ int s_run[256];
    __shared__ int s_w[kSortWarps * 256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 256; i += kSortThreads) s_run[i] = base[(long long)blockIdx.x * 256 + i];
    const long long b = blockIdx.x * chunk, e = min(n, b + chunk);
    for (long long t0 = b; t0 < e; t0 += kSortThreads) {
        for (int i = tid; i < kSortWarps * 256; i += kSortThreads) s_w[i] = 0;
        __syncthreads();
        const long long i = t0 + tid;
        const bool ok = i < e;
        const int key = ok ? keys_in[i] : 0;
        const int val = ok ? (vals_in ? vals_in[i] : (int)i) : 0;
        const unsigned int d = ok ? digit_of(key, n_bins, shift) : 256u;
        // rank among equal digits inside the warp, in lane (= index) order
        const unsigned int peers = __match_any_sync(0xffffffffu, d);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (ok && rank == 0) s_w[warp * 256 + d] = __popc(peers);
        __syncthreads();
        if (tid < 256) {                                   // exclusive prefix over the warps for digit `tid`
            int acc = s_run[tid];
            for (int w = 0; w < kSortWarps; ++w) {
                const int c = s_w[w * 256 + tid];
                s_w[w * 256 + tid] = acc;
                acc += c;
            }
            s_run[tid] = acc;
        }
        __syncthreads();
        if (ok) {
            const int dst = s_w[warp * 256 + d] + rank;
            keys_out[dst] = key;
            vals_out[dst] = val;
        }
        __syncthreads();
    }
}

// offsets[w] = first position of window w in the sorted order (w = 0 .. n_bins; offsets[n_bins] = number of kept points)
__global__ void window_offsets_kernel(const int* __restrict__ keys_sorted, long long n, int n_bins, long long* __restrict__ offsets,
                                      const int* __restrict__ vals_sorted, long long* __restrict__ order) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i <= n; i += (long long)gridDim.x * blockDim.x) {
        const int cur = i < n ? (keys_sorted[i] < 0 ? n_bins : keys_sorted[i]) : n_bins + 1;
        const int prev = i > 0 ? (keys_sorted[i - 1] < 0 ? n_bins : keys_sorted[i - 1]) : -1;
        for (int w = prev + 1; w <= cur && w <= n_bins; ++w) offsets[w] = i;
        if (i < n) order[i] = vals_sorted[i];
    }
}

// ---- filter + normalise, one CTA per window -------------------------------------------------------------------
// cols [P, 10] float64 = (x, y, z, hag, class, intensity, red, green, blue, nir); rows of window w = order[offsets[w] .. offsets[w + 1])
__device__ __forceinline__ bool keep_row(const double* __restrict__ r, double max_z) {
    const int c = (int)r[4];
    const bool cls_ok = c != 2 && c != 7 && c != 8 && c != 13 && c != 24 && c != 30;
    return cls_ok && r[3] <= max_z && r[3] >= 0.0;
}

// pass A: kept count and min / max of x, y over the kept rows of every window
__global__ void __launch_bounds__(kThreads) filter_stats_kernel(const double* __restrict__ cols, const long long* __restrict__ order,
                                                                const long long* __restrict__ offsets, double max_z,
                                                                long long* __restrict__ kept, double* __restrict__ stats) {
    __shared__ unsigned long long s_k[4];
    __shared__ int s_cnt;
    const int w = blockIdx.x;
    if (threadIdx.x == 0) { s_k[0] = ~0ull; s_k[1] = 0ull; s_k[2] = ~0ull; s_k[3] = 0ull; s_cnt = 0; }
    __syncthreads();
    unsigned long long mnx = ~0ull, mxx = 0ull, mny = ~0ull, mxy = 0ull;
    int cnt = 0;
    for (long long i = offsets[w] + threadIdx.x; i < offsets[w + 1]; i += kThreads) {
        const double* r = cols + order[i] * 10;
        if (keep_row(r, max_z)) {
            const unsigned long long kx = ord_bits(r[0]), ky = ord_bits(r[1]);
            mnx = min(mnx, kx); mxx = max(mxx, kx); mny = min(mny, ky); mxy = max(mxy, ky);
            ++cnt;
        }
    }
    for (int o = 16; o >= 1; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&s_k[0], mnx); atomicMax(&s_k[1], mxx); atomicMin(&s_k[2], mny); atomicMax(&s_k[3], mxy);
        atomicAdd(&s_cnt, cnt);
    }
    __syncthreads();
    if (threadIdx.x == 0) kept[w] = s_cnt;
    if (threadIdx.x < 4) stats[w * 4 + threadIdx.x] = s_cnt > 0 ? ord_value(s_k[threadIdx.x]) : 0.0;
}

// exclusive scan of the kept counts (one CTA; a window whose x or y extent is zero yields no rows: 2_preprocessing...:92)
__global__ void filter_scan_kernel(const long long* __restrict__ kept, const double* __restrict__ stats, int n_windows,
                                   long long* __restrict__ out_offsets) {
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int w = 0; w < n_windows; ++w) {
            out_offsets[w] = run;
            const bool flat = stats[w * 4 + 1] - stats[w * 4] == 0.0 || stats[w * 4 + 3] - stats[w * 4 + 2] == 0.0;
            run += (kept[w] > 0 && !flat) ? kept[w] : 0;
        }
        out_offsets[n_windows] = run;
    }
}

// pass B: stable compaction of the kept rows + the 13-column normalised row
__global__ void __launch_bounds__(kThreads) filter_write_kernel(const double* __restrict__ cols, const long long* __restrict__ order,
                                                                const long long* __restrict__ offsets, double max_z, double max_intensity,
                                                                const double* __restrict__ stats, const long long* __restrict__ out_offsets,
                                                                double* __restrict__ out) {
    __shared__ int s_warp[kThreads / 32];
    __shared__ long long s_base;
    const int w = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (out_offsets[w + 1] == out_offsets[w]) return;                  // nothing kept, or a flat window
    const double xmin = stats[w * 4], xmax = stats[w * 4 + 1], ymin = stats[w * 4 + 2], ymax = stats[w * 4 + 3];
    const double xr = __dsub_rn(xmax, xmin), yr = __dsub_rn(ymax, ymin);
    if (threadIdx.x == 0) s_base = out_offsets[w];
    __syncthreads();
    for (long long t0 = offsets[w]; t0 < offsets[w + 1]; t0 += kThreads) {
        const long long i = t0 + threadIdx.x;
        const double* r = i < offsets[w + 1] ? cols + order[i] * 10 : nullptr;
        const bool keep = r != nullptr && keep_row(r, max_z);
        const unsigned int m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int j = 0; j < kThreads / 32; ++j) { const int c = s_warp[j]; if (j < warp) before += c; total += c; }
        if (keep) {
            double* o = out + (s_base + before + __popc(m & ((1u << lane) - 1u))) * 13;
            const double x = r[0], y = r[1], z = r[2], hag = r[3], red = r[6], nir = r[9];
            // numpy: 2 * ((x - min) / (max - min)) - 1   (2 * t is exact, one rounding in the subtraction)
            o[0] = __dsub_rn(__dmul_rn(2.0, __ddiv_rn(__dsub_rn(x, xmin), xr)), 1.0);
            o[1] = __dsub_rn(__dmul_rn(2.0, __ddiv_rn(__dsub_rn(y, ymin), yr)), 1.0);
            o[2] = __ddiv_rn(hag, max_z);
            o[3] = r[4];
            o[4] = fmin(fmax(__ddiv_rn(r[5], max_intensity), 0.0), 1.0);
            o[5] = __ddiv_rn(red, 65536.0);
            o[6] = __ddiv_rn(r[7], 65536.0);
            o[7] = __ddiv_rn(r[8], 65536.0);
            o[8] = fmin(fmax(__ddiv_rn(nir, 65535.0), 0.0), 1.0);
            const double ndvi = __ddiv_rn(__dadd_rn(__ddiv_rn(__dsub_rn(nir, red), __dadd_rn(nir, red)), 1.0), 2.0);
            o[9] = ndvi != ndvi ? ndvi : fmin(fmax(ndvi, 0.0), 1.0);   // np.clip keeps NaN (0 / 0 when nir == red == 0)
            o[10] = x; o[11] = y; o[12] = z;
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
}

inline unsigned int grid_for(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    if (b > kNumSMs * 8) b = kNumSMs * 8;
    return (unsigned int)(b < 1 ? 1 : b);
}

}  // namespace
}  // namespace amp

extern "C" {

int amp_minmax_f64(const double* x, const double* y, int64_t n, int64_t stride, double* out4, void* workspace, size_t workspace_bytes,
                   void* stream) {
    using namespace amp;
    if (!x || !y || !out4 || !workspace) return fail(AMP_E_BADARG, "minmax: null pointer");
    if (n < 1 || stride < 1) return fail(AMP_E_BADARG, "minmax: empty input");
    if (workspace_bytes < 32) return fail(AMP_E_WORKSPACE, "minmax: workspace needs 32 bytes");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(workspace);
    minmax_init_kernel<<<1, 32, 0, st>>>(keys);
    minmax_kernel<<<grid_for(n, kThreads), kThreads, 0, st>>>(x, y, n, stride, keys);
    minmax_decode_kernel<<<1, 32, 0, st>>>(keys, out4);
    count_launch(3);
    return check_launch("minmax");
}

int amp_window_ids_f64(const double* x, const double* y, int64_t n, int64_t stride, double x0, double y0, int32_t wx, int32_t wy,
                       int32_t nx, int32_t ny, int32_t* ids, void* stream) {
    using namespace amp;
    if (!x || !y || !ids) return fail(AMP_E_BADARG, "window_ids: null pointer");
    if (n < 1 || stride < 1 || wx < 1 || wy < 1 || nx < 0 || ny < 0 || (long long)nx * ny > (1 << 24))
        return fail(AMP_E_BADARG, "window_ids: bad grid %d x %d", nx, ny);
    window_ids_kernel<<<grid_for(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(x, y, n, stride, x0, y0, (double)wx, (double)wy, nx, ny, ids);
    count_launch();
    return check_launch("window_ids");
}

size_t amp_window_partition_workspace_bytes(int64_t n) {
    const long long chunks = (n + 8191) / 8192 < 1 ? 1 : (n + 8191) / 8192;
    return (size_t)n * 16 + (size_t)chunks * 256 * 4 + 1024;
}

int amp_window_partition(const int32_t* ids, int64_t n, int32_t n_bins, int64_t* order, int64_t* offsets, void* workspace,
                         size_t workspace_bytes, void* stream) {
    using namespace amp;
    if (!ids || !order || !offsets || !workspace) return fail(AMP_E_BADARG, "window_partition: null pointer");
    if (n < 1 || n >= (1LL << 31) || n_bins < 1 || n_bins > (1 << 24)) return fail(AMP_E_BADARG, "window_partition: bad size");
    if (workspace_bytes < amp_window_partition_workspace_bytes(n)) return fail(AMP_E_WORKSPACE, "window_partition: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const long long chunk = 8192;
    const int n_chunks = (int)((n + chunk - 1) / chunk);
    int* ka = reinterpret_cast<int*>(workspace);
    int* va = ka + n; int* kb = va + n; int* vb = kb + n;
    int* hist = reinterpret_cast<int*>((reinterpret_cast<uintptr_t>(vb + n) + 255) & ~(uintptr_t)255);
    const int* kin = ids; const int* vin = nullptr;
    int* kout = ka; int* vout = va;
    int passes = 1;
    while ((1LL << (8 * passes)) <= n_bins) ++passes;                  // digits needed for keys 0 .. n_bins (n_bins = dropped)
    for (int p = 0; p < passes; ++p) {
        radix_hist_kernel<<<n_chunks, 256, 0, st>>>(kin, n, chunk, n_bins, 8 * p, hist);
        radix_scan_kernel<<<1, 256, 0, st>>>(hist, n_chunks);
        radix_scatter_kernel<<<n_chunks, kSortThreads, 0, st>>>(kin, vin, n, chunk, n_bins, 8 * p, hist, kout, vout);
        kin = kout; vin = vout;
        if (kout == ka) { kout = kb; vout = vb; } else { kout = ka; vout = va; }
        count_launch(3);
    }
    window_offsets_kernel<<<grid_for(n + 1, kThreads), kThreads, 0, st>>>(kin, n, n_bins, reinterpret_cast<long long*>(offsets), vin,
                                                                         reinterpret_cast<long long*>(order));
    count_launch();
    return check_launch("window_partition");
}

size_t amp_filter_normalize_workspace_bytes(int64_t n_windows) { return (size_t)n_windows * (8 + 32) + 256; }

int amp_filter_normalize_f64(const double* cols, const int64_t* order, const int64_t* offsets, int64_t n_windows, double max_z,
                             double max_intensity, double* out, int64_t* out_offsets, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace amp;
    if (!cols || !order || !offsets || !out || !out_offsets || !workspace) return fail(AMP_E_BADARG, "filter_normalize: null pointer");
    if (n_windows < 1 || n_windows > 65535 * 16) return fail(AMP_E_BADARG, "filter_normalize: bad window count");
    if (workspace_bytes < amp_filter_normalize_workspace_bytes(n_windows)) return fail(AMP_E_WORKSPACE, "filter_normalize: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    long long* kept = reinterpret_cast<long long*>(workspace);
    double* stats = reinterpret_cast<double*>(kept + n_windows);
    filter_stats_kernel<<<(unsigned)n_windows, kThreads, 0, st>>>(cols, reinterpret_cast<const long long*>(order),
                                                                  reinterpret_cast<const long long*>(offsets), max_z, kept, stats);
    filter_scan_kernel<<<1, 32, 0, st>>>(kept, stats, (int)n_windows, reinterpret_cast<long long*>(out_offsets));
    filter_write_kernel<<<(unsigned)n_windows, kThreads, 0, st>>>(cols, reinterpret_cast<const long long*>(order),
                                                                  reinterpret_cast<const long long*>(offsets), max_z, max_intensity, stats,
                                                                  reinterpret_cast<const long long*>(out_offsets), out);
    count_launch(3);
    return check_launch("filter_normalize");
}

}  // extern "C"
