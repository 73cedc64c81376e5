// Device-side constants and log-density / simulator math of the fused model family.
//
// Two arithmetic modes (glabc_arith_mode):
//   STRICT  the reference's float32 operation order with IEEE add/mul/div/sqrt and no FMA
//           contraction (__f*_rn intrinsics are never contracted by nvcc) — replay parity mode;
//   FAST    same formulas with FMA, reciprocal multiplies and constants folded on the host.
// Reference formulas: distribution.py:166-181 (DiagGaussian), examples/Mixture.py:13-53.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/glabc.h"

namespace glabc {

// Everything a sampler kernel needs about the model and up to two DiagGaussian proposals,
// passed by value as a __grid_constant__ kernel parameter (lives in the constant bank, so the
// values are usable directly as FFMA operands without occupying registers).
struct GaussConsts {
    float loc[GLABC_MAX_DIM];
    float log_scale[GLABC_MAX_DIM];
    float scale[GLABC_MAX_DIM];
    float inv_scale[GLABC_MAX_DIM];  // FAST: 1/scale
    float nloc_inv[GLABC_MAX_DIM];   // FAST: -loc/scale, so r = fma(z, inv_scale, nloc_inv)
    float c;                         // float32(-0.5*d*log(2*pi))
    float c_fast;                    // c - sum(log_scale)
};

struct ModelConsts {
    int32_t family;
    float y_obs[GLABC_MAX_DIM];
    float noise_loc[GLABC_MAX_DIM];
    float noise_scale[GLABC_MAX_DIM];
    float eps_log_scale, eps_scale;
    float c_kern;        // float32(-0.5*log(2*pi))
    float kern_fast_c;   // c_kern - eps_log_scale
    float kern_fast_m;   // -0.5 / eps_scale^2
    GaussConsts prior;
};

// torch.sum over n < 16 contiguous float32 (ATen SumKernel.cpp row_sum, ilp_factor 4): four
// interleaved partials, tail into partial 0, then partials 1..3 folded in.  Left-to-right for n<=4.
template <int N>
__device__ __forceinline__ float torch_sum_strict(const float (&v)[N])
{
    static_assert(N >= 1 && N < 16, "vectorised path not needed for n < 16");
    float p[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    constexpr int rows = N / 4;
#pragma unroll
    for (int r = 0; r < rows; ++r)
#pragma unroll
        for (int k = 0; k < 4; ++k) p[k] = __fadd_rn(p[k], v[r * 4 + k]);
#pragma unroll
    for (int i = rows * 4; i < N; ++i) p[0] = __fadd_rn(p[0], v[i]);
#pragma unroll
    for (int k = 1; k < 4; ++k) p[0] = __fadd_rn(p[0], p[k]);
    return p[0];
}

// DiagGaussian.log_prob(z), distribution.py:176-181
template <int D, bool STRICT>
__device__ __forceinline__ float gauss_log_prob(const GaussConsts& g, const float (&z)[D])
{
    if constexpr (STRICT) {
        float t[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const float r = __fdiv_rn(__fsub_rn(z[i], g.loc[i]), g.scale[i]);
            t[i] = __fadd_rn(g.log_scale[i], __fmul_rn(0.5f, __fmul_rn(r, r)));
        }
        return __fsub_rn(g.c, torch_sum_strict<D>(t));
    } else {
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const float r = fmaf(z[i], g.inv_scale[i], g.nloc_inv[i]);
            q = fmaf(r, r, q);
        }
        return fmaf(-0.5f, q, g.c_fast);
    }
}

// DiagGaussian.forward(): z = loc + exp(log_scale)*eps; log_p from eps, distribution.py:166-174
template <int D, bool STRICT>
__device__ __forceinline__ float gauss_forward(const GaussConsts& g, const float (&eps)[D], float (&z)[D])
{
    if constexpr (STRICT) {
        float t[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            z[i] = __fadd_rn(g.loc[i], __fmul_rn(g.scale[i], eps[i]));
            t[i] = __fadd_rn(g.log_scale[i], __fmul_rn(0.5f, __fmul_rn(eps[i], eps[i])));
        }
        return __fsub_rn(g.c, torch_sum_strict<D>(t));
    } else {
        float q = 0.0f;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            z[i] = fmaf(g.scale[i], eps[i], g.loc[i]);
            q = fmaf(eps[i], eps[i], q);
        }
        return fmaf(-0.5f, q, g.c_fast);
    }
}

// generate_samples, Mixture.py:13-26: mean(theta) + (noise_loc + noise_scale*eps)
template <int D, bool STRICT>
__device__ __forceinline__ void model_simulate(const ModelConsts& m, const float (&theta)[D],
                                               const float (&eps)[D], float (&y)[D])
{
    const bool use_abs = m.family == GLABC_MODEL_ABS_NORMAL;  // launch-uniform
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const float mean = use_abs ? fabsf(theta[i]) : theta[i];
        if constexpr (STRICT) {
            y[i] = __fadd_rn(mean, __fadd_rn(m.noise_loc[i], __fmul_rn(m.noise_scale[i], eps[i])));
        } else {
            y[i] = fmaf(m.noise_scale[i], eps[i], mean + m.noise_loc[i]);
        }
    }
}

// calculate_log_kernel(y), Mixture.py:33-45: log N(||y - y_obs||_2 ; 0, eps)
template <int D, bool STRICT>
__device__ __forceinline__ float model_log_kernel(const ModelConsts& m, const float (&y)[D])
{
    if constexpr (STRICT) {
        float t[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const float dy = __fsub_rn(y[i], m.y_obs[i]);
            t[i] = __fmul_rn(dy, dy);
        }
        const float dis = __fsqrt_rn(torch_sum_strict<D>(t));
        const float r = __fdiv_rn(__fsub_rn(dis, 0.0f), m.eps_scale);
        return __fsub_rn(m.c_kern, __fadd_rn(m.eps_log_scale, __fmul_rn(0.5f, __fmul_rn(r, r))));
    } else {
        // (sqrt(s)/eps)^2 == s/eps^2 up to rounding: the sqrt is skipped
        float s = 0.0f;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const float dy = y[i] - m.y_obs[i];
            s = fmaf(dy, dy, s);
        }
        return fmaf(m.kern_fast_m, s, m.kern_fast_c);
    }
}

template <int D, bool STRICT>
__device__ __forceinline__ float model_prior(const ModelConsts& m, const float (&theta)[D])
{
    return gauss_log_prob<D, STRICT>(m.prior, theta);
}

}  // namespace glabc
