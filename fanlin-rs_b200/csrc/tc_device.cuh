// Device helpers of the tensor-core kernels (tcgen05 / TMEM / mbarrier), shared by kernels_fused_tc.cu and
// kernels_blur_tc.cu.  Included inside an anonymous namespace user: everything here is static inline device code.
#pragma once
#include <cuda.h>

#include "fused_device.cuh"
#include "fused_tc.h"

namespace fanlin {
namespace {

// Shared-memory matrix descriptor, no swizzle.  Measured on B200
// (profiles/microbench/umma_i8.cu): LBO = byte stride between core matrices along K,
// SBO = along M/N, for both the MN-major A tile and the K-major B tile.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout = 0) {
    uint64_t d = uint64_t((saddr & 0x3FFFFu) >> 4) | (uint64_t(layout) << 61);  // layout 0 = no swizzle, 2 = 128-byte swizzle
    d |= uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= uint64_t(1) << 46;  // descriptor version of sm_100
    return d;
}

// Instruction descriptor: D = S32 (2 @ bit 4), A = U8 (0 @ bit 7), B = S8 (1 @ bit 10),
// A is MN-major (bit 15), B is K-major, N >> 3 @ bit 17, M >> 4 @ bit 24.
constexpr uint32_t UMMA_IDESC = (2u << 4) | (0u << 7) | (1u << 10) | (1u << 15) | ((TC_N >> 3) << 17) | ((TC_M >> 4) << 24);

// One lane of a converged warp.  Unlike `lane == 0`, elect.sync tells ptxas that a single thread runs
// the branch, so tcgen05.mma / TMA operands stay in uniform registers: with `lane == 0` every
// UTCIMMA sat in an ELECT / BRA.U.ANY loop and issued every ~110 clk whatever its shape
// (profiles/microbench/umma_rate_elect.cu: 71 clk for the same loop, N / 2 clk from N = 192 up).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// MBAR_HINT: suspend-time hint (ns) of try_wait -- the thread may sleep in the barrier unit that long before the
// instruction returns false, instead of coming back to the issue slots every few hundred cycles (35 % of the
// warp instructions of the round-1 kernel were this loop).  MBAR_SLEEP: __nanosleep between failed tries.
#ifndef MBAR_HINT
#define MBAR_HINT 0
#endif
#ifndef MBAR_SLEEP
#define MBAR_SLEEP 0
#endif
// WD + -DMBAR_WATCHDOG (debug builds, `make EXTRA=-DMBAR_WATCHDOG`): after ~2^22 failed tries (seconds; a healthy wait is
// microseconds) the CTA traps, so that a protocol error in a kernel under development ends in a launch failure and not in
// a hung GPU.  Off in the shipped build: the counter and the trap cost 4-7 % of the C2 kernel even when only the
// single-thread role warps carry them (A/B on one box, 20 steps: 6.42-6.51 ms without, 6.67-6.96 ms with) -- the role
// threads' wait loops sit between the tensor core's completion and the next MMA / copy they issue.
template <bool WD>
__device__ __forceinline__ void mbar_wait_impl(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    (void)spins;
    while (!done) {
#if MBAR_HINT
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity), "r"(uint32_t(MBAR_HINT)) : "memory");
#else
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
#endif
#if MBAR_SLEEP
        if (!done) __nanosleep(MBAR_SLEEP);
#endif
#ifdef MBAR_WATCHDOG
        if constexpr (WD) {
            if (!done && ++spins > (1u << 22)) __trap();
        }
#endif
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { mbar_wait_impl<false>(bar, parity); }
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity) { mbar_wait_impl<true>(bar, parity); }

#ifdef TC2_PROF
#define PW(acc, ...) do { const long long t0_ = clock64(); mbar_wait(__VA_ARGS__); (acc) += clock64() - t0_; } while (0)
#define PWR(acc, ...) do { const long long t0_ = clock64(); mbar_wait_wd(__VA_ARGS__); (acc) += clock64() - t0_; } while (0)
#else
#define PW(acc, ...) mbar_wait(__VA_ARGS__)
#define PWR(acc, ...) mbar_wait_wd(__VA_ARGS__)  // role threads: with the watchdog in debug builds
#endif

__device__ __forceinline__ void ffma2(float2 &acc, float2 a, float w) {
    const float2 b = make_float2(w, w);
    asm("fma.rn.f32x2 %0, %1, %2, %0;"
        : "+l"(reinterpret_cast<unsigned long long &>(acc))
        : "l"(reinterpret_cast<const unsigned long long &>(a)), "l"(reinterpret_cast<const unsigned long long &>(b)));
}

__device__ __forceinline__ void sts8(uint32_t saddr, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" ::"r"(saddr), "r"(v) : "memory"); }

__device__ __forceinline__ float lds_f32(uint32_t saddr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ float4 lds_f32x4(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t r[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {  // {lo, hi} -> f16x2, round to nearest even
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
    float2 r;
    asm("{\n\t.reg .b16 a, b;\n\tmov.b32 {a, b}, %2;\n\tcvt.f32.f16 %0, a;\n\tcvt.f32.f16 %1, b;\n\t}\n" : "=f"(r.x), "=f"(r.y) : "r"(v));
    return r;
}
// zeroes 16 TMEM columns of the warp's 32 lanes
__device__ __forceinline__ void tmem_st16_zero(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr), "r"(z) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}


}  // namespace
}  // namespace fanlin
