// Inline-PTX helpers for the fp32-class fused chains (tc_chain32.cu): tcgen05.mma with the A operand in TENSOR MEMORY
// ("TS" form), tcgen05.st, split-bf16 packing, warp-wide max of float bit patterns.
//
// Operand layout of A in TMEM for kind::f16 (checked against a host reference by tools/ts_chain_probe.cu): lane = row of the
// 128-row tile, one 32-bit column holds two consecutive K elements (low half = even k), 8 columns per 16-wide K step.
#pragma once
#include "tc_ptx.cuh"

namespace amp {
namespace tcx {

// D[tmem] (+)= A[tmem] * B[smem descriptor]^T, kind::f16 (bf16 operands, fp32 accumulate), one CTA
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// x = hi + lo with hi = bf16(x), lo = bf16(x - hi): the two bf16 terms of the split-precision operands, for a pair of
// consecutive K elements (a = even k -> low half)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(a, b);
    const float ah = __uint_as_float(hi << 16), bh = __uint_as_float(hi & 0xffff0000u);
    lo = pack_bf16x2(a - ah, b - bh);
}

// max over the 32 lanes of the float whose bits are v, valid when the true maximum is >= 0 (float bit patterns order like
// signed integers for non-negative values, and every negative value compares below every non-negative one); when all 32 values
// are negative the result is some negative value, which the ReLU that follows maps to 0 either way.
__device__ __forceinline__ float warp_max_relu_safe(float v) {
    int r;
    asm volatile("redux.sync.max.s32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(__float_as_int(v)));
    return __int_as_float(r);
}

// bounded mbarrier wait: a protocol error becomes a trap (launch failure reported by the next CUDA call) instead of a hung GPU.
// try_wait suspends the thread in hardware for a while between polls: a test_wait spin by every warp of a slot floods the
// MIO pipe with SYNCS instructions and slows the OTHER slot's epilogue (its shared-memory bias loads took ~90 cycles each:
// 1.4 k cycles for 64 columns of bias + ReLU, measured with the clock64 phase profile).
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    uint32_t done;
    for (int spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > (1 << 22)) __trap();          // far beyond any legitimate wait of these kernels
    }
}

}  // namespace tcx
}  // namespace amp
