// Farthest-point sampling for sm_100a: one thread-block CLUSTER per cloud, the cloud resident
// on chip for the whole run (coordinates in registers + shared memory, running min-distance in
// registers), no cluster-wide barrier inside a pick.
//
// Replaces utils/utils.py:889-933 `fps` (reference). Semantics are those pinned by
// oracle/fps_oracle.py: start index given, picked points leave the candidate set, lowest index
// wins ties, d = (dx*dx + dy*dy) + dz*dz with separately rounded operations (no FMA).
//
// Data placement for a cloud of P points on a cluster of C CTAs x THREADS threads
// (TT = C*THREADS): point i belongs to cluster-thread g = i % TT, slot j = i / TT.
//   slots [0, RS)         x,y,z and min-dist in registers
//   slots [RS, RS+DS)     x,y,z in shared memory (pairs of slots per 8-byte entry, conflict-free), min-dist in registers
//   slots >= RS+DS        x,y,z,min-dist in a global workspace (only for P > 8 * capacity)
// A pick:
//   * per thread: the subtract-square-add chain of TWO slots at a time on packed fp32 pairs (FFMA2, bit-identical to the
//     scalar operations), one min and one max per slot (FMNMX: the slot of the maximum is NOT tracked in the loop);
//   * per CTA: warp maxima by redux -> shared memory -> ONE block barrier -> every warp reduces the NW values. Only the
//     warp that carries the CTA maximum rescans its slots (lowest slot per lane, lowest index over the lanes by a second
//     redux) and its winning lane posts ONE candidate (d, index, x, y, z) into the table of EVERY CTA of the cluster, its own
//     included, with st.async -- a remote shared-memory store that completes transaction bytes on the receiver's
//     mbarrier. Equal maxima in several warps (rare): a shared-memory atomicMin on the index and a second block barrier;
//   * every thread waits on its CTA's mbarrier and picks the best of the C candidates itself.
// There is no barrier.cluster in the loop (round 1 had one per pick, with its MEMBAR.ALL.GPU). The candidate tables and
// mbarriers are double buffered by pick parity: a CTA can only post pick s + 2 after it has seen all posts of pick s + 1,
// which every peer sends after all its threads have passed the block barrier of pick s + 1, i.e. after they have read the
// candidates of pick s. The thread that owns the picked point retires it (d = -1) before the next pick, outside the loop.
#include <cooperative_groups.h>

#include "amp_common.cuh"

namespace cg = cooperative_groups;

namespace amp {
namespace {

constexpr unsigned kNoIdx = 0xffffffffu;

__device__ __forceinline__ float sqd(float lx, float ly, float lz, float x, float y, float z) {
    float dx = __fsub_rn(lx, x), dy = __fsub_rn(ly, y), dz = __fsub_rn(lz, z);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}
__device__ __forceinline__ double sqd(double lx, double ly, double lz, double x, double y, double z) {
    double dx = __dsub_rn(lx, x), dy = __dsub_rn(ly, y), dz = __dsub_rn(lz, z);
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
// min that keeps d when nd is NaN (non-finite input rows are reported through `status`), max likewise
__device__ __forceinline__ float keep_min(float d, float nd) { return fminf(d, nd); }
__device__ __forceinline__ double keep_min(double d, double nd) { return (nd < d) ? nd : d; }
__device__ __forceinline__ float keep_max(float b, float d) { return fmaxf(b, d); }
__device__ __forceinline__ double keep_max(double b, double d) { return (d > b) ? d : b; }

// All real distances are >= +0, picked / padded slots carry -1, so for float the IEEE bit pattern ordered as a signed int
// is the distance order.
// warp-wide maximum of the value alone
__device__ __forceinline__ float warp_max(float d) {
    return __int_as_float(__reduce_max_sync(0xffffffffu, __float_as_int(d)));
}
__device__ __forceinline__ double warp_max(double d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double od = __shfl_xor_sync(0xffffffffu, d, o);
        d = (od > d) ? od : d;
    }
    return d;
}

// Packed fp32 pairs (sm_100 FFMA2): two slots per instruction, each half rounded exactly like the scalar operation.
// Products and sums are written as fused multiply-adds that restate the unfused operations exactly,
//   rn(a - b) = fma(b, -1, a),  rn(a * a) = fma(a, a, -0),  rn(a + b) = fma(a, 1, b),
// with the constants arriving as kernel arguments (opaque to ptxas, which would otherwise contract a packed multiply
// and a packed add into one fused operation and change the rounding).
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
struct FpsConsts { float one, neg_one, neg_zero; };
// squared distances of two slots to the last pick, each half bit-identical to sqd(float...)
__device__ __forceinline__ f32x2_t sqd2(f32x2_t lx, f32x2_t ly, f32x2_t lz, f32x2_t x, f32x2_t y, f32x2_t z, f32x2_t one, f32x2_t neg_one,
                                        f32x2_t neg_zero) {
    const f32x2_t dx = fma2(x, neg_one, lx), dy = fma2(y, neg_one, ly), dz = fma2(z, neg_one, lz);
    const f32x2_t xx = fma2(dx, dx, neg_zero), yy = fma2(dy, dy, neg_zero), zz = fma2(dz, dz, neg_zero);
    return fma2(fma2(xx, one, yy), one, zz);
}

template <typename T>
__device__ __forceinline__ bool is_finite3(T x, T y, T z) {
    return isfinite(x) && isfinite(y) && isfinite(z);
}

// ---- st.async: remote shared-memory store that completes bytes on the receiver's mbarrier ----
__device__ __forceinline__ uint32_t fps_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, int rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_async(uint32_t addr, unsigned v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(addr), "r"(v), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_async(uint32_t addr, float v, uint32_t bar) { st_async(addr, __float_as_uint(v), bar); }
__device__ __forceinline__ void st_async(uint32_t addr, double v, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(addr), "l"(__double_as_longlong(v)), "r"(bar) : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t addr, unsigned a, unsigned b, unsigned c, unsigned d, uint32_t bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr), "r"(a),
                 "r"(b), "r"(c), "r"(d), "r"(bar) : "memory");
}
__device__ __forceinline__ void fps_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fps_mbar_expect(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fps_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "FPS_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra FPS_DONE_%=;\n\t"
        "bra FPS_WAIT_%=;\n\t"
        "FPS_DONE_%=:\n\t}" ::"r"(bar), "r"(parity), "r"(4000u) : "memory");   // suspend-time hint (ns): wake on completion, not by polling
}

// bytes one candidate adds to a receiver's transaction count: d, idx, x, y, z
template <typename T> struct CandBytes { static constexpr unsigned v = 4 * sizeof(T) + 4; };

#ifdef AMP_FPS_PROF
__device__ long long g_fps_prof[3 * 8 * 6];
#define FPS_PROF(k) do { if (blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 15 || warp == 31) && s >= 100 && s < 108) \
    g_fps_prof[((warp == 0 ? 0 : warp == 15 ? 1 : 2) * 8 + (s - 100)) * 6 + (k)] = clock64(); } while (0)
#else
#define FPS_PROF(k) do { } while (0)
#endif

// LC >= 0: cluster size 2^LC known at compile time (the float variants: every table offset of the exchange folds into
// an immediate); LC < 0: taken from the argument.
template <typename T, int THREADS, int RS, int DS, int LC>
__global__ void __launch_bounds__(THREADS, 1)
fps_cluster_kernel(const T* __restrict__ pc, int P, long long row_stride, int S, int start_idx,
                   long long* __restrict__ out_idx, int* __restrict__ status,
                   T* __restrict__ ovf, int ovf_slots, int log2C_arg, const FpsConsts kc) {
    const int log2C = LC >= 0 ? LC : log2C_arg;
    constexpr int NW = THREADS / 32;
    static_assert(RS % 2 == 0 && DS % 2 == 0, "slots are processed in pairs");
    // shared-memory slots as PAIRS: entry [j / 2][tid] holds slots j and j + 1 of this thread (8-byte reads feed the packed
    // arithmetic directly; conflict-free)
    struct Pair { T a, b; };
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Pair* sx = reinterpret_cast<Pair*>(smem_raw);
    Pair* sy = sx + (DS / 2) * THREADS;
    Pair* sz = sy + (DS / 2) * THREADS;
    const int C = 1 << log2C;
    // candidate tables [2][C] (one entry per CTA of the cluster and pick parity), then two mbarriers, the warp maxima
    // [2][NW] and the tie cells [2]. float: {d, idx, x, y} as one 16-byte entry (one st.async.v4 per receiver) + z;
    // double: one array per field
    constexpr bool kF32 = sizeof(T) == 4;
    unsigned char* c_base = reinterpret_cast<unsigned char*>(sz + (DS / 2) * THREADS);
    uint4* c_a = reinterpret_cast<uint4*>(c_base);                       // float only
    float* c_zf = reinterpret_cast<float*>(c_a + 2 * C);                  // float only
    T* c_d = reinterpret_cast<T*>(c_base);                                // double only from here
    T* c_x = c_d + 2 * C;
    T* c_y = c_x + 2 * C;
    T* c_z = c_y + 2 * C;
    unsigned* c_i = reinterpret_cast<unsigned*>(c_z + 2 * C);
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(c_base + ((size_t)2 * C * CandBytes<T>::v + 15) / 16 * 16);
    T* s_wmax = reinterpret_cast<T*>(s_bar + 2);
    unsigned* s_tie = reinterpret_cast<unsigned*>(s_wmax + 2 * NW);

    cg::cluster_group cluster = cg::this_cluster();
    const int r = (int)cluster.block_rank();
    const int b = blockIdx.x >> log2C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int TT = THREADS << log2C;
    const int log2TT = log2C + 31 - __clz(THREADS);
    const int g = r * THREADS + tid;
    const uint32_t bar0 = fps_smem_u32(&s_bar[0]);
    const uint32_t pick_bytes = (uint32_t)C * CandBytes<T>::v;  // one candidate per CTA and pick

    if (tid == 0) {
        fps_mbar_init(bar0, 1);
        fps_mbar_init(bar0 + 8, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fps_mbar_expect(bar0 + 8, pick_bytes);                   // pick 1 (parity 1), then pick 2 (parity 0)
        fps_mbar_expect(bar0, pick_bytes);
        s_tie[0] = kNoIdx; s_tie[1] = kNoIdx;
    }

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
        const int o = (j >> 1) * THREADS + tid;
        (&sx[o].a)[j & 1] = x; (&sy[o].a)[j & 1] = y; (&sz[o].a)[j & 1] = z;
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
    // every mbarrier of the cluster is initialised and armed before anybody posts into it
    if (C > 1) cluster.sync(); else __syncthreads();

    for (int s = 1; s < S; ++s) {
        const int par = s & 1;
        FPS_PROF(0);
        // ---- the owner of the last pick retires it: d = -1 stays -1 under min() and never wins ----
        if ((last & (TT - 1)) == g) {
            const int lslot = last >> log2TT;
#pragma unroll
            for (int j = 0; j < RS; ++j)
                if (lslot == j) rd[j] = (T)-1;
#pragma unroll
            for (int j = 0; j < DS; ++j)
                if (lslot == RS + j) dd[j] = (T)-1;
            if (lslot >= RS + DS) my_ovf[(long long)(lslot - RS - DS) * (4 * THREADS) + 3 * THREADS + tid] = (T)-1;
        }
        T bd = (T)-1;
        if constexpr (kF32) {
            const f32x2_t one = pack2(kc.one, kc.one), neg_one = pack2(kc.neg_one, kc.neg_one), neg_zero = pack2(kc.neg_zero, kc.neg_zero);
            const f32x2_t lx2 = pack2(lx, lx), ly2 = pack2(ly, ly), lz2 = pack2(lz, lz);
#pragma unroll
            for (int j = 0; j < RS; j += 2) {
                float na, nb;
                unpack2(sqd2(lx2, ly2, lz2, pack2(rx[j], rx[j + 1]), pack2(ry[j], ry[j + 1]), pack2(rz[j], rz[j + 1]), one, neg_one, neg_zero), na, nb);
                rd[j] = keep_min(rd[j], na); rd[j + 1] = keep_min(rd[j + 1], nb);
                bd = keep_max(bd, rd[j]); bd = keep_max(bd, rd[j + 1]);
            }
#pragma unroll
            for (int j = 0; j < DS; j += 2) {
                const int o = (j >> 1) * THREADS + tid;
                float na, nb;
                unpack2(sqd2(lx2, ly2, lz2, *reinterpret_cast<const f32x2_t*>(&sx[o]), *reinterpret_cast<const f32x2_t*>(&sy[o]),
                             *reinterpret_cast<const f32x2_t*>(&sz[o]), one, neg_one, neg_zero), na, nb);
                dd[j] = keep_min(dd[j], na); dd[j + 1] = keep_min(dd[j + 1], nb);
                bd = keep_max(bd, dd[j]); bd = keep_max(bd, dd[j + 1]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < RS; ++j) {
                rd[j] = keep_min(rd[j], sqd(lx, ly, lz, rx[j], ry[j], rz[j]));
                bd = keep_max(bd, rd[j]);
            }
#pragma unroll
            for (int j = 0; j < DS; ++j) {
                const int o = (j >> 1) * THREADS + tid;
                dd[j] = keep_min(dd[j], sqd(lx, ly, lz, (&sx[o].a)[j & 1], (&sy[o].a)[j & 1], (&sz[o].a)[j & 1]));
                bd = keep_max(bd, dd[j]);
            }
        }
        for (int j = 0; j < ovf_slots; ++j) {
            T* o = my_ovf + (long long)j * (4 * THREADS);
            const T d = keep_min(o[3 * THREADS + tid], sqd(lx, ly, lz, o[tid], o[THREADS + tid], o[2 * THREADS + tid]));
            o[3 * THREADS + tid] = d;
            bd = keep_max(bd, d);
        }

        // ---- CTA maximum: warp maxima by redux -> shared memory -> one block barrier -> every warp reduces the NW values ----
        FPS_PROF(1);
        const T wd = warp_max(bd);
        if (lane == 0) s_wmax[par * NW + warp] = wd;
        __syncthreads();
        const T wm = (NW == 32 || lane < NW) ? s_wmax[par * NW + (lane & (NW - 1))] : (T)-1;
        const T cm = warp_max(wm);
        const int n_tied = __popc(__ballot_sync(0xffffffffu, wm == cm));     // warps that carry the CTA maximum (uniform over the CTA)
        if (tid == 0) s_tie[par ^ 1] = kNoIdx;                               // the other parity's cell: nobody reads or writes it now
        // Only a warp that carries the CTA maximum looks for the slot (the rescan is 2 ALU instructions per slot, and the ALU
        // pipe is the busiest one of this kernel): lowest slot per lane, lowest index over the lanes by a second redux.
        unsigned my_idx = kNoIdx, wi = kNoIdx;
        int bslot = 0;
        bool poster = false;
        if (cm < (T)0) {
            poster = warp == 0 && lane == 0;                                 // nothing left in this CTA: an empty candidate
        } else if (wd == cm) {
            if (bd == wd) {
                for (int j = ovf_slots - 1; j >= 0; --j)
                    if (my_ovf[(long long)j * (4 * THREADS) + 3 * THREADS + tid] == wd) bslot = RS + DS + j;
#pragma unroll
                for (int j = DS - 1; j >= 0; --j)
                    if (dd[j] == wd) bslot = RS + j;
#pragma unroll
                for (int j = RS - 1; j >= 0; --j)
                    if (rd[j] == wd) bslot = j;
                my_idx = (unsigned)(bslot * TT + g);
            }
            wi = __reduce_min_sync(0xffffffffu, my_idx);
            if (n_tied == 1) poster = my_idx == wi;
            else if (lane == 0) atomicMin(&s_tie[par], wi);
        }
        if (cm >= (T)0 && n_tied > 1) {                                      // rare: equal distances in different warps
            __syncthreads();
            poster = wd == cm && my_idx == wi && s_tie[par] == wi;
        }
        FPS_PROF(2);
        if (poster) {
            T x = 0, y = 0, z = 0;
            if (wi != kNoIdx) {
                if (bslot < RS) {
#pragma unroll
                    for (int j = 0; j < RS; ++j)
                        if (j == bslot) { x = rx[j]; y = ry[j]; z = rz[j]; }
                } else if (bslot < RS + DS) {
                    const int q = bslot - RS, o = (q >> 1) * THREADS + tid;
                    x = (&sx[o].a)[q & 1]; y = (&sy[o].a)[q & 1]; z = (&sz[o].a)[q & 1];
                } else {
                    T* o = my_ovf + (long long)(bslot - RS - DS) * (4 * THREADS);
                    x = o[tid]; y = o[THREADS + tid]; z = o[2 * THREADS + tid];
                }
            }
            const int slot = par * C + r;
            const uint32_t a_b = bar0 + 8u * (uint32_t)par;
            if constexpr (kF32) {
                const uint32_t a_a = fps_smem_u32(&c_a[slot]), a_z = fps_smem_u32(&c_zf[slot]);
#pragma unroll
                for (int p = 0; p < C; ++p) {
                    // a CTA's shared window is contiguous in the cluster address space: one mapa, then plain offsets
                    const uint32_t off = map_to_cta(bar0, p) - bar0;
                    st_async_v4(off + a_a, __float_as_uint(cm), wi, __float_as_uint(x), __float_as_uint(y), off + a_b);
                    st_async(off + a_z, z, off + a_b);
                }
            } else {
                const uint32_t a_d = fps_smem_u32(&c_d[slot]), a_x = fps_smem_u32(&c_x[slot]), a_y = fps_smem_u32(&c_y[slot]),
                               a_z = fps_smem_u32(&c_z[slot]), a_i = fps_smem_u32(&c_i[slot]);
#pragma unroll
                for (int p = 0; p < C; ++p) {
                    const uint32_t off = map_to_cta(bar0, p) - bar0;
                    const uint32_t pb = off + a_b;
                    st_async(off + a_d, cm, pb);
                    st_async(off + a_i, wi, pb);
                    st_async(off + a_x, x, pb);
                    st_async(off + a_y, y, pb);
                    st_async(off + a_z, z, pb);
                }
            }
        }
        FPS_PROF(3);
        // ---- wait for the C candidates of this pick (one per CTA); every thread reduces them itself ----
        fps_mbar_wait(bar0 + 8u * (uint32_t)par, (uint32_t)(((s - 1) >> 1) & 1));   // k-th use of this barrier: picks 2k+1 / 2k+2
        FPS_PROF(4);
        if (tid == 0 && s + 2 < S) fps_mbar_expect(bar0 + 8u * (uint32_t)par, pick_bytes);   // re-arm for pick s + 2
        {
            T gd = (T)-2;
            unsigned gi = kNoIdx;
#pragma unroll
            for (int p = 0; p < C; ++p) {
                T d, x, y, z;
                unsigned i2;
                if constexpr (kF32) {
                    const uint4 v = c_a[par * C + p];
                    d = __uint_as_float(v.x); i2 = v.y; x = __uint_as_float(v.z); y = __uint_as_float(v.w); z = c_zf[par * C + p];
                } else {
                    d = c_d[par * C + p]; i2 = c_i[par * C + p]; x = c_x[par * C + p]; y = c_y[par * C + p]; z = c_z[par * C + p];
                }
                if (d > gd || (d == gd && i2 < gi)) { gd = d; gi = i2; lx = x; ly = y; lz = z; }
            }
            last = (int)gi;
        }
        if (g == 0) out_idx[(long long)b * S + s] = last;
        FPS_PROF(5);
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

// 512 threads per CTA (16 warps, 128 registers per thread): with 40 slots per thread instead of 20, twelve of them (not
// two) keep their coordinates in registers, so a pick reads 28 x 512 x 12 B = 172 KB of shared memory per SM instead of
// 221 KB (the shared-memory port is what bounds the slot loop), and the block barrier, the redux steps and the rescan of
// a pick run in half as many warps. (RS, DS) variants keep small clouds from looping over absent slots.
constexpr int kThreads = 512;
template <typename T> struct MaxSlots { static constexpr int rs = 12, ds = 28; };     // 40 slots x 512 threads = 20 480 points per CTA
template <> struct MaxSlots<double> { static constexpr int rs = 4, ds = 16; };         // 20 slots x 512 threads = 10 240 points per CTA

template <typename T, int RS, int DS, int LC>
int launch_kernel(const T* pc, int64_t B, int P, int64_t row_stride, int S, int start_idx,
                  int64_t* out_idx, int32_t* status, T* ovf, int ovf_slots, int log2C,
                  cudaStream_t st) {
    auto kern = fps_cluster_kernel<T, kThreads, RS, DS, LC>;
    const size_t n_cand = (size_t)(kThreads / 32) << log2C;
    size_t smem = (size_t)3 * DS * kThreads * sizeof(T) + 2 * n_cand * (4 * sizeof(T) + 4) + 8 + 16;
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
    FpsConsts kc;
    kc.one = 1.f; kc.neg_one = -1.f; kc.neg_zero = -0.f;
    e = cudaLaunchKernelEx(&cfg, kern, pc, P, (long long)row_stride, S, start_idx, oi, status, ovf,
                           ovf_slots, log2C, kc);
    if (e != cudaSuccess) return fail(AMP_E_CUDA, "fps launch: %s", cudaGetErrorString(e));
    count_launch();
    return AMP_OK;
}

template <typename T, int RS, int DS>
int launch_variant(const T* pc, int64_t B, int P, int64_t row_stride, int S, int start_idx,
                   int64_t* out_idx, int32_t* status, T* ovf, int ovf_slots, int log2C,
                   cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        switch (log2C) {
            case 0: return launch_kernel<T, RS, DS, 0>(pc, B, P, row_stride, S, start_idx, out_idx, status, ovf, ovf_slots, log2C, st);
            case 1: return launch_kernel<T, RS, DS, 1>(pc, B, P, row_stride, S, start_idx, out_idx, status, ovf, ovf_slots, log2C, st);
            case 2: return launch_kernel<T, RS, DS, 2>(pc, B, P, row_stride, S, start_idx, out_idx, status, ovf, ovf_slots, log2C, st);
            default: return launch_kernel<T, RS, DS, 3>(pc, B, P, row_stride, S, start_idx, out_idx, status, ovf, ovf_slots, log2C, st);
        }
    } else {
        return launch_kernel<T, RS, DS, -1>(pc, B, P, row_stride, S, start_idx, out_idx, status, ovf, ovf_slots, log2C, st);
    }
}

template <typename T>
int64_t cap_per_cta() { return (int64_t)(MaxSlots<T>::rs + MaxSlots<T>::ds) * kThreads; }

// cluster size: fill the 148 SMs, then grow until the cloud fits on chip (max 8 CTAs)
template <typename T>
int choose_log2C(int64_t B, int64_t P) {
    int lc = 0;
    while (lc < 3 && B * (2LL << lc) <= kNumSMs) ++lc;
    while (lc > 0 && P < ((int64_t)kThreads << lc)) --lc;
    while (lc < 3 && (cap_per_cta<T>() << lc) < P) ++lc;
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
    const int64_t cap = cap_per_cta<T>();
    if (per_cta > cap) {
        ovf_slots = (int)((per_cta - cap + kThreads - 1) / kThreads);
        size_t need = (size_t)B * (1u << lc) * ovf_slots * 4 * kThreads * sizeof(T);
        if (!ws || ws_bytes < need)
            return fail(AMP_E_WORKSPACE, "fps: workspace %zu bytes < %zu needed", ws_bytes, need);
        ovf = reinterpret_cast<T*>(ws);
    }
    const int slots = (int)((per_cta + kThreads - 1) / kThreads);   // slots actually needed per thread
#define AMP_FPS_VARIANT(RS_, DS_)                                                                \
    if (slots <= RS_ + DS_)                                                                      \
        return launch_variant<T, RS_, DS_>(pc, B, (int)P, row_stride, S, start_idx, out_idx, status, ovf, ovf_slots, lc, st);
    if constexpr (sizeof(T) == 4) {
        AMP_FPS_VARIANT(2, 0)
        AMP_FPS_VARIANT(4, 0)
        AMP_FPS_VARIANT(8, 0)
        AMP_FPS_VARIANT(12, 0)
        AMP_FPS_VARIANT(12, 4)
        AMP_FPS_VARIANT(12, 8)
        AMP_FPS_VARIANT(12, 16)
    } else {
        AMP_FPS_VARIANT(2, 0)
        AMP_FPS_VARIANT(4, 0)
        AMP_FPS_VARIANT(4, 4)
        AMP_FPS_VARIANT(4, 8)
    }
#undef AMP_FPS_VARIANT
    return launch_variant<T, MaxSlots<T>::rs, MaxSlots<T>::ds>(pc, B, (int)P, row_stride, S, start_idx, out_idx, status,
                                                               ovf, ovf_slots, lc, st);
}

}  // namespace
}  // namespace amp

#ifdef AMP_FPS_PROF
extern "C" int amp_fps_prof_dump(long long* out) {
    return (int)cudaMemcpyFromSymbol(out, amp::g_fps_prof, sizeof(long long) * 3 * 8 * 6);
}
#endif

extern "C" {

size_t amp_fps_workspace_bytes(int64_t B, int64_t P, int32_t elem_bytes) {
    const int64_t cap = (elem_bytes == 8 ? amp::cap_per_cta<double>() : amp::cap_per_cta<float>());
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
