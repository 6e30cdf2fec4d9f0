// Farthest-point sampling for sm_100a: one thread-block CLUSTER per cloud, the cloud resident
// on chip for the whole run (coordinates in registers + shared memory, running min-distance in
// registers), one cluster barrier per pick.
//
// Replaces utils/utils.py:889-933 `fps` (reference). Semantics are those pinned by
// oracle/fps_oracle.py: start index given, picked points leave the candidate set, lowest index
// wins ties, d = (dx*dx + dy*dy) + dz*dz with separately rounded operations (no FMA).
//
// Data placement for a cloud of P points on a cluster of C CTAs x THREADS threads
// (TT = C*THREADS): point i belongs to cluster-thread g = i % TT, slot j = i / TT.
//   slots [0, RS)         x,y,z and min-dist in registers
//   slots [RS, RS+DS)     x,y,z in shared memory (SoA, conflict-free), min-dist in registers
//   slots >= RS+DS        x,y,z,min-dist in a global workspace (only for P > 8 * capacity)
// Each pick: every thread updates its slots against the last pick and keeps its best; warp
// argmax by redux.sync; per-CTA argmax through shared memory; every CTA posts its candidate
// (with coordinates) into every peer's shared memory (DSMEM); one cluster barrier; everybody
// reduces the C candidates locally. Candidate slots are double-buffered so one barrier per pick
// is enough.
#include <cooperative_groups.h>

#include "amp_common.cuh"

namespace cg = cooperative_groups;

namespace amp {
namespace {

constexpr unsigned kNoIdx = 0xffffffffu;

template <typename T>
struct Cand {
    T d;
    unsigned idx;
    T x, y, z;
};

__device__ __forceinline__ float sqd(float lx, float ly, float lz, float x, float y, float z) {
    float dx = __fsub_rn(lx, x), dy = __fsub_rn(ly, y), dz = __fsub_rn(lz, z);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
__device__ __forceinline__ double sqd(double lx, double ly, double lz, double x, double y, double z) {
    double dx = __dsub_rn(lx, x), dy = __dsub_rn(ly, y), dz = __dsub_rn(lz, z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// Warp-wide argmax of (d desc, idx asc). On return every lane holds the winner.
// All real distances are >= +0, picked / padded slots carry -1, so for float the IEEE bit
// pattern ordered as a signed int is the distance order.
__device__ __forceinline__ void warp_argmax(float& d, unsigned& idx) {
    int kb = __float_as_int(d);
    int m = __reduce_max_sync(0xffffffffu, kb);
    unsigned c = (kb == m) ? idx : kNoIdx;
    idx = __reduce_min_sync(0xffffffffu, c);
    d = __int_as_float(m);
}
__device__ __forceinline__ void warp_argmax(double& d, unsigned& idx) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double od = __shfl_xor_sync(0xffffffffu, d, o);
        unsigned oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (od > d || (od == d && oi < idx)) { d = od; idx = oi; }
    }
}

template <typename T>
__device__ __forceinline__ bool is_finite3(T x, T y, T z) {
    return isfinite(x) && isfinite(y) && isfinite(z);
}

template <typename T, int THREADS, int RS, int DS>
__global__ void __launch_bounds__(THREADS, 1)
fps_cluster_kernel(const T* __restrict__ pc, int P, long long row_stride, int S, int start_idx,
                   long long* __restrict__ out_idx, int* __restrict__ status,
                   T* __restrict__ ovf, int ovf_slots, int log2C) {
    constexpr int NW = THREADS / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sx = reinterpret_cast<T*>(smem_raw);
    T* sy = sx + DS * THREADS;
    T* sz = sy + DS * THREADS;
    Cand<T>* s_warp = reinterpret_cast<Cand<T>*>(sz + DS * THREADS);   // [NW]
    Cand<T>* s_clu = s_warp + NW;                                       // [2][8]

    cg::cluster_group cluster = cg::this_cluster();
    const int C = 1 << log2C;
    const int r = (int)cluster.block_rank();
    const int b = blockIdx.x >> log2C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int TT = THREADS << log2C;
    const int log2TT = log2C + 31 - __clz(THREADS);
    const int g = r * THREADS + tid;

    const T* cloud = pc + (long long)b * P * row_stride;
    T* my_ovf = ovf ? ovf + ((long long)(b * C + r) * ovf_slots) * (4 * THREADS) : nullptr;

    // ---- load the cloud into registers / shared memory / overflow; absent slots get d = -1 ----
    T rx[RS], ry[RS], rz[RS], rd[RS];
    T dd[DS > 0 ? DS : 1];
    bool bad = false;
    const T kInf = (T)INFINITY;
#pragma unroll
    for (int j = 0; j < RS; ++j) {
        int i = j * TT + g;
        rx[j] = ry[j] = rz[j] = (T)0; rd[j] = (T)-1;
        if (i < P) {
            const T* p = cloud + (long long)i * row_stride;
            rx[j] = p[0]; ry[j] = p[1]; rz[j] = p[2]; rd[j] = kInf;
            bad |= !is_finite3(rx[j], ry[j], rz[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < DS; ++j) {
        int i = (RS + j) * TT + g;
        T x = 0, y = 0, z = 0; dd[j] = (T)-1;
        if (i < P) {
            const T* p = cloud + (long long)i * row_stride;
            x = p[0]; y = p[1]; z = p[2]; dd[j] = kInf;
            bad |= !is_finite3(x, y, z);
        }
        sx[j * THREADS + tid] = x; sy[j * THREADS + tid] = y; sz[j * THREADS + tid] = z;
    }
    for (int j = 0; j < ovf_slots; ++j) {
        int i = (RS + DS + j) * TT + g;
        T x = 0, y = 0, z = 0, d = (T)-1;
        if (i < P) {
            const T* p = cloud + (long long)i * row_stride;
            x = p[0]; y = p[1]; z = p[2]; d = kInf;
            bad |= !is_finite3(x, y, z);
        }
        T* o = my_ovf + (long long)j * (4 * THREADS);
        o[tid] = x; o[THREADS + tid] = y; o[2 * THREADS + tid] = z; o[3 * THREADS + tid] = d;
    }
    if (bad && status) atomicExch(&status[b], 1);

    int last = start_idx;
    T lx, ly, lz;
    {
        const T* p = cloud + (long long)last * row_stride;
        lx = p[0]; ly = p[1]; lz = p[2];
    }
    if (g == 0) out_idx[(long long)b * S] = last;

    for (int s = 1; s < S; ++s) {
        const int par = s & 1;
        const bool own_last = (last & (TT - 1)) == g;
        const int lslot = last >> log2TT;
        T bd = (T)-1;
        int bslot = 0;
#pragma unroll
        for (int j = 0; j < RS; ++j) {
            T d = rd[j];
            T nd = sqd(lx, ly, lz, rx[j], ry[j], rz[j]);
            d = (nd < d) ? nd : d;                // picked / absent slots hold -1 and stay -1
            if (own_last && lslot == j) d = (T)-1;
            rd[j] = d;
            if (d > bd) { bd = d; bslot = j; }    // strict: lowest slot (= lowest index) wins
        }
#pragma unroll
        for (int j = 0; j < DS; ++j) {
            T d = dd[j];
            T nd = sqd(lx, ly, lz, sx[j * THREADS + tid], sy[j * THREADS + tid], sz[j * THREADS + tid]);
            d = (nd < d) ? nd : d;
            if (own_last && lslot == RS + j) d = (T)-1;
            dd[j] = d;
            if (d > bd) { bd = d; bslot = RS + j; }
        }
        for (int j = 0; j < ovf_slots; ++j) {
            T* o = my_ovf + (long long)j * (4 * THREADS);
            T d = o[3 * THREADS + tid];
            T nd = sqd(lx, ly, lz, o[tid], o[THREADS + tid], o[2 * THREADS + tid]);
            d = (nd < d) ? nd : d;
            if (own_last && lslot == RS + DS + j) d = (T)-1;
            o[3 * THREADS + tid] = d;
            if (d > bd) { bd = d; bslot = RS + DS + j; }
        }

        // ---- warp argmax; the winning lane publishes (d, idx, coords) ----
        const unsigned my_idx = (bd >= (T)0) ? (unsigned)(bslot * TT + g) : kNoIdx;
        T wd = bd;
        unsigned wi = my_idx;
        warp_argmax(wd, wi);
        const bool lane_wins = (wi == kNoIdx) ? (lane == 0) : (my_idx == wi);
        if (lane_wins) {
            T x = 0, y = 0, z = 0;
            if (wi != kNoIdx) {
                if (bslot < RS) {
#pragma unroll
                    for (int j = 0; j < RS; ++j)
                        if (j == bslot) { x = rx[j]; y = ry[j]; z = rz[j]; }
                } else if (bslot < RS + DS) {
                    int o = (bslot - RS) * THREADS + tid;
                    x = sx[o]; y = sy[o]; z = sz[o];
                } else {
                    T* o = my_ovf + (long long)(bslot - RS - DS) * (4 * THREADS);
                    x = o[tid]; y = o[THREADS + tid]; z = o[2 * THREADS + tid];
                }
            }
            Cand<T> c; c.d = wd; c.idx = wi; c.x = x; c.y = y; c.z = z;
            s_warp[warp] = c;
        }
        __syncthreads();
        // ---- CTA argmax by warp 0; post the CTA candidate into every peer's slot ----
        if (warp == 0) {
            Cand<T> c;
            c.d = (T)-1; c.idx = kNoIdx; c.x = c.y = c.z = (T)0;
            if (lane < NW) c = s_warp[lane];
            T d2 = c.d;
            unsigned i2 = c.idx;
            warp_argmax(d2, i2);
            const bool wins = (i2 == kNoIdx) ? (lane == 0) : (c.idx == i2);
            if (wins) {
                for (int p = 0; p < C; ++p) {
                    Cand<T>* dst = cluster.map_shared_rank(&s_clu[par * 8 + r], p);
                    *dst = c;
                }
            }
        }
        if (C > 1) cluster.sync(); else __syncthreads();
        // ---- everybody reduces the C CTA candidates ----
        Cand<T> best = s_clu[par * 8];
        for (int p = 1; p < C; ++p) {
            Cand<T> c = s_clu[par * 8 + p];
            if (c.d > best.d || (c.d == best.d && c.idx < best.idx)) best = c;
        }
        last = (int)best.idx;
        lx = best.x; ly = best.y; lz = best.z;
        if (g == 0) out_idx[(long long)b * S + s] = last;
    }
    if (C > 1) cluster.sync();   // no CTA may exit while a peer could still write into it
}

template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ pc, long long P, long long row_elems,
                                   const long long* __restrict__ idx, long long S, T* __restrict__ out,
                                   long long total) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        long long col = e % row_elems, row = e / row_elems;
        long long b = row / S;
        out[e] = pc[(b * P + idx[row]) * row_elems + col];
    }
}

constexpr int kThreads = 1024;
constexpr int kRS = 2;
template <typename T> struct MaxDS { static constexpr int v = 18; };
template <> struct MaxDS<double> { static constexpr int v = 8; };

template <typename T, int DS>
int launch_variant(const T* pc, int64_t B, int P, int64_t row_stride, int S, int start_idx,
                   int64_t* out_idx, int32_t* status, T* ovf, int ovf_slots, int log2C,
                   cudaStream_t st) {
    auto kern = fps_cluster_kernel<T, kThreads, kRS, DS>;
    size_t smem = (size_t)3 * DS * kThreads * sizeof(T) + sizeof(Cand<T>) * (kThreads / 32 + 16);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(AMP_E_CUDA, "fps: smem attribute: %s", cudaGetErrorString(e));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(B << log2C));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1u << log2C;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    long long* oi = reinterpret_cast<long long*>(out_idx);
    e = cudaLaunchKernelEx(&cfg, kern, pc, P, (long long)row_stride, S, start_idx, oi, status, ovf,
                           ovf_slots, log2C);
    if (e != cudaSuccess) return fail(AMP_E_CUDA, "fps launch: %s", cudaGetErrorString(e));
    count_launch();
    return AMP_OK;
}

template <typename T>
int64_t cap_per_cta(int ds) { return (int64_t)(kRS + ds) * kThreads; }

// cluster size: fill the 148 SMs, then grow until the cloud fits on chip (max 8 CTAs)
template <typename T>
int choose_log2C(int64_t B, int64_t P) {
    int lc = 0;
    while (lc < 3 && B * (2LL << lc) <= kNumSMs) ++lc;
    while (lc > 0 && P < ((int64_t)kThreads << lc)) --lc;
    while (lc < 3 && (cap_per_cta<T>(MaxDS<T>::v) << lc) < P) ++lc;
    return lc;
}

template <typename T>
int fps_impl(const T* pc, int64_t B, int64_t P, int64_t row_stride, int32_t S, int32_t start_idx,
             int64_t* out_idx, int32_t* status, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!pc || !out_idx) return fail(AMP_E_BADARG, "fps: null pointer");
    if (B < 1 || P < 1 || P >= (1LL << 31) - 1 || row_stride < 3)
        return fail(AMP_E_BADARG, "fps: bad shape B=%lld P=%lld row_stride=%lld", (long long)B,
                    (long long)P, (long long)row_stride);
    if (S < 1 || S > P) return fail(AMP_E_BADARG, "fps: n_samples=%d must be in [1, P=%lld]", S, (long long)P);
    if (start_idx < 0 || start_idx >= P) return fail(AMP_E_BADARG, "fps: start_idx out of range");
    if (status) {
        cudaError_t e = cudaMemsetAsync(status, 0, sizeof(int32_t) * B, st);
        if (e != cudaSuccess) return fail(AMP_E_CUDA, "fps: memset: %s", cudaGetErrorString(e));
    }
    const int lc = choose_log2C<T>(B, P);
    const int64_t per_cta = (P + (1LL << lc) - 1) >> lc;
    int ovf_slots = 0;
    T* ovf = nullptr;
    const int64_t cap = cap_per_cta<T>(MaxDS<T>::v);
    if (per_cta > cap) {
        ovf_slots = (int)((per_cta - cap + kThreads - 1) / kThreads);
        size_t need = (size_t)B * (1u << lc) * ovf_slots * 4 * kThreads * sizeof(T);
        if (!ws || ws_bytes < need)
            return fail(AMP_E_WORKSPACE, "fps: workspace %zu bytes < %zu needed", ws_bytes, need);
        ovf = reinterpret_cast<T*>(ws);
    }
    const int slots = (int)((per_cta + kThreads - 1) / kThreads);   // slots actually needed per thread
#define AMP_FPS_VARIANT(DS_)                                                                     \
    if (slots <= kRS + DS_ && DS_ <= MaxDS<T>::v)                                                \
        return launch_variant<T, (DS_ <= MaxDS<T>::v ? DS_ : 0)>(pc, B, (int)P, row_stride, S,   \
                                                                 start_idx, out_idx, status, ovf, \
                                                                 ovf_slots, lc, st);
    AMP_FPS_VARIANT(0)
    AMP_FPS_VARIANT(2)
    AMP_FPS_VARIANT(6)
    AMP_FPS_VARIANT(8)
    AMP_FPS_VARIANT(10)
    AMP_FPS_VARIANT(18)
#undef AMP_FPS_VARIANT
    return launch_variant<T, MaxDS<T>::v>(pc, B, (int)P, row_stride, S, start_idx, out_idx, status,
                                          ovf, ovf_slots, lc, st);
}

}  // namespace
}  // namespace amp

extern "C" {

size_t amp_fps_workspace_bytes(int64_t B, int64_t P, int32_t elem_bytes) {
    const int64_t cap = (elem_bytes == 8 ? amp::cap_per_cta<double>(amp::MaxDS<double>::v)
                                         : amp::cap_per_cta<float>(amp::MaxDS<float>::v));
    if (P <= 8 * cap) return 0;
    const int64_t per_cta = (P + 7) / 8;
    const int64_t ovf_slots = (per_cta - cap + amp::kThreads - 1) / amp::kThreads;
    return (size_t)B * 8 * ovf_slots * 4 * amp::kThreads * (size_t)elem_bytes;
}

int amp_fps_f32(const float* pc, int64_t B, int64_t P, int64_t row_stride, int32_t S,
                int32_t start_idx, int64_t* out_idx, int32_t* status, void* workspace,
                size_t workspace_bytes, void* stream) {
    return amp::fps_impl<float>(pc, B, P, row_stride, S, start_idx, out_idx, status, workspace,
                                workspace_bytes, (cudaStream_t)stream);
}

int amp_fps_f64(const double* pc, int64_t B, int64_t P, int64_t row_stride, int32_t S,
                int32_t start_idx, int64_t* out_idx, int32_t* status, void* workspace,
                size_t workspace_bytes, void* stream) {
    return amp::fps_impl<double>(pc, B, P, row_stride, S, start_idx, out_idx, status, workspace,
                                 workspace_bytes, (cudaStream_t)stream);
}

int amp_gather_rows(const void* pc, int64_t B, int64_t P, int64_t row_elems, int32_t elem_bytes,
                    const int64_t* idx, int64_t S, void* out, void* stream) {
    if (!pc || !idx || !out) return amp::fail(AMP_E_BADARG, "gather_rows: null pointer");
    if (elem_bytes != 4 && elem_bytes != 8) return amp::fail(AMP_E_BADARG, "gather_rows: elem_bytes");
    long long total = (long long)B * S * row_elems;
    if (total <= 0) return AMP_OK;
    int threads = 256;
    long long blocks = (total + threads - 1) / threads;
    if (blocks > amp::kNumSMs * 16) blocks = amp::kNumSMs * 16;
    const long long* li = reinterpret_cast<const long long*>(idx);
    if (elem_bytes == 4)
        amp::gather_rows_kernel<float><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
            (const float*)pc, P, row_elems, li, S, (float*)out, total);
    else
        amp::gather_rows_kernel<double><<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
            (const double*)pc, P, row_elems, li, S, (double*)out, total);
    amp::count_launch();
    return amp::check_launch("gather_rows");
}

}  // extern "C"
