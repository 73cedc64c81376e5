// Philox4x32-10 counter-based RNG and the native-mode stream layout (DESIGN.md "RNG streams").
//
// The reference draws from torch's / numpy's process-global generators (GlobalMCMC.py:39,47;
// GLMCMC.py:17); a kernel over 65,536+ independent chains needs a stateless generator keyed by the
// GLOBAL chain id so a chain's trace does not depend on how chains are sharded over GPUs.
//   counter = (chain_lo, chain_hi, block, slot), key = (seed_lo, seed_hi)
//   slot 1      : the "step block" of step `block`: ONE Philox call carries everything a
//                 d<=2 GlobalMCMC / local step draws (32x32->64 multiplies are quarter-rate on
//                 sm_100, so the bit budget is spent carefully):
//                   w.x[31:8]  24 bits  Box-Muller radius uniform, pair a     w.x[7:0]  \  16-bit branch
//                   w.z[31:8]  24 bits  Box-Muller radius uniform, pair b     w.z[7:0]  /  uniform U_b
//                   w.y[31:12] 20 bits  Box-Muller angle, pair a              w.y[11:0] \  24-bit accept
//                   w.w[31:12] 20 bits  Box-Muller angle, pair b              w.w[11:0] /  uniform U_a
//                 (disjoint bit fields of one block are independent uniform bits)
//   slot 2 + g  : extra normal blocks (4 normals each) for d > 2 and for iSIR candidates
//   slot 0x80000000 : float64 resampling uniform of step `block` (iSIR)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace glabc {

constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

constexpr uint32_t kSlotStep = 1u;
constexpr uint32_t kSlotNormal = 2u;
constexpr uint32_t kSlotU64 = 0x80000000u;

// The ten round keys are launch-uniform: the host expands them once (RoundKeys) and the kernel
// reads them from its constant bank as direct LOP3 operands — 2 IMAD.WIDE + 2 LOP3 per round.
struct RoundKeys {
    uint32_t k[10][2];
};

__host__ __device__ inline RoundKeys expand_key(uint2 key)
{
    RoundKeys rk;
    for (int r = 0; r < 10; ++r) {
        rk.k[r][0] = key.x + static_cast<uint32_t>(r) * kPhiloxW0;
        rk.k[r][1] = key.y + static_cast<uint32_t>(r) * kPhiloxW1;
    }
    return rk;
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, const RoundKeys& rk)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = static_cast<uint64_t>(kPhiloxM0) * c.x;
        const uint64_t p1 = static_cast<uint64_t>(kPhiloxM1) * c.z;
        c = make_uint4(static_cast<uint32_t>(p1 >> 32) ^ c.y ^ rk.k[r][0], static_cast<uint32_t>(p1),
                       static_cast<uint32_t>(p0 >> 32) ^ c.w ^ rk.k[r][1], static_cast<uint32_t>(p0));
    }
    return c;
}

struct Stream {
    uint32_t chain_lo, chain_hi;
    __device__ __forceinline__ uint4 block(const RoundKeys& rk, uint32_t blk, uint32_t slot) const
    {
        return philox4x32_10(make_uint4(chain_lo, chain_hi, blk, slot), rk);
    }
};

// 24-bit uniform on torch.rand's float32 grid {k * 2^-24}
__device__ __forceinline__ float u24(uint32_t w) { return __uint2float_rn(w >> 8) * 0x1p-24f; }

// MUFU.LG2 without the denormal-input fix-up __log2f carries (inputs here are never denormal)
__device__ __forceinline__ float lg2_approx(float x)
{
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// natural log of a 24-bit uniform k*2^-24 (k = 0 gives -inf, as torch.log(0) does, B-16)
__device__ __forceinline__ float log_approx(float x) { return 0.69314718055994531f * lg2_approx(x); }

// Box-Muller pair from two words, MUFU path (lg2, sqrt, sin, cos): 4 MUFU + ~8 FP per pair.
// radius uniform u1 = (k + 0.5) * 2^-24 with k = w0[31:8]; angle = 2*pi * j * 2^-20 with j = w1[31:12].
// The masked words have <= 24 significant bits, so the int->float conversions are exact and the
// low bits (used for U_b / U_a) cannot leak into the normals.
__device__ __forceinline__ void box_muller(uint32_t w0, uint32_t w1, float& n0, float& n1)
{
    const float u1 = fmaf(__uint2float_rn(w0 & 0xFFFFFF00u), 0x1p-32f, 0x1p-25f);
    const float r2 = -1.3862943611198906f * lg2_approx(u1);  // -2 ln2 * lg2(u1) = -2 ln(u1)
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(r2));
    const float a = __uint2float_rn(w1 & 0xFFFFF000u) * (6.28318530717958647692f * 0x1p-32f);
    n0 = r * __cosf(a);
    n1 = r * __sinf(a);
}

// U_b of a step block in the top 16 bits (low 16 bits are don't-care): global iff ub_hi < thr << 16
__device__ __forceinline__ uint32_t step_block_ub(const uint4& w) { return __byte_perm(w.x, w.z, 0x0400); }
// U_a of a step block: 24 bits
__device__ __forceinline__ uint32_t step_block_ua(const uint4& w) { return (w.y & 0xFFFu) | ((w.w << 12) & 0xFFF000u); }

}  // namespace glabc
