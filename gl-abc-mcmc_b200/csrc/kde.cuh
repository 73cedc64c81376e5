// K5 — weighted Gaussian KernelDensity (reference kernel_density.py:4-177), batched over independent point
// sets: one big set for the stand-alone estimator (1e5 x 1e5 pairs) or one small set per chain for AGLMCMC.
//
//   k_kde_fit       one CTA per set: normalised weights, weighted unbiased std, Silverman / Scott bandwidth
//   k_kde_logprob   CTA = (set, tile of queries); training points staged through shared memory in tiles and
//                   broadcast to all threads; each thread owns Q queries.  The pair loop is MUFU(ex2)/issue
//                   bound: d subs + d FMAs + 1 FMA + 1 ex2 + 1 add per pair, no HBM traffic to speak of.
//                   d = 2 makes the "pairwise-distance GEMM" a K=2 contraction — not a tensor-core shape
//                   (SURVEY.md 8(d)); CUDA cores + MUFU are the right pipes.
//   k_kde_cdf / k_kde_sample   inverse-CDF categorical draw + Gaussian jitter (or the reference's own
//                   torch.multinomial indices / normals in replay mode)
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/glabc.h"
#include "f32x2.cuh"
#include "philox.cuh"

namespace glabc {

constexpr uint32_t kSlotAdSample = 0x40000000u;  // + 2q: categorical uniform, + 2q + 1: jitter normals
constexpr uint32_t kSlotAdSim = 0x48000000u;     // + b: simulator normals of block candidate b
constexpr uint32_t kSlotInit = 0x30000u;         // + b*G + g: normals of the initial block

struct KdeSets {
    const float* X;         // [sets][cap][D]
    const float* weights;   // [sets][cap] normalised
    const float* bw;        // [sets][D]
    const int32_t* n;       // [sets] or nullptr (= cap)
    const int32_t* active;  // [sets] or nullptr: sets with active[s] == 0 are skipped
    int64_t sets, cap;
};

__device__ __forceinline__ double block_sum_f64(double v, double* red)
{
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        int lo = __double2loint(v), hi = __double2hiint(v);
        lo = __shfl_xor_sync(0xffffffffu, lo, off);
        hi = __shfl_xor_sync(0xffffffffu, hi, off);
        v += __hiloint2double(hi, lo);
    }
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
    for (int i = 0; i < nw; ++i) t += red[i];
    return t;
}

// fit(), kernel_density.py:70-94 with _compute_bandwidth :22-37 and weighted_std :39-68
template <int D>
__global__ void __launch_bounds__(256) k_kde_fit(const float* __restrict__ X, const float* __restrict__ w,
                                                 const int32_t* __restrict__ n_dev, const int32_t* __restrict__ active,
                                                 int64_t cap, int rule, float* __restrict__ weights_out,
                                                 float* __restrict__ lw_out, float* __restrict__ bw_out)
{
    __shared__ double red[8];
    const int64_t s = blockIdx.x;
    if (active != nullptr && active[s] == 0) return;
    const int n = n_dev != nullptr ? n_dev[s] : static_cast<int>(cap);
    if (n < 1) return;
    const float* Xs = X + s * cap * D;
    const float* ws = w != nullptr ? w + s * cap : nullptr;
    float* wo = weights_out + s * cap;

    double acc = 0.0;
    if (ws != nullptr)
        for (int j = threadIdx.x; j < n; j += blockDim.x) acc += static_cast<double>(ws[j]);
    const float swf = static_cast<float>(block_sum_f64(acc, red));
    const float uni = __fdiv_rn(1.0f, static_cast<float>(n));
    acc = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const float v = ws != nullptr ? __fdiv_rn(ws[j], swf) : uni;  // :80,83
        wo[j] = v;
        if (lw_out != nullptr) lw_out[s * cap + j] = logf(__fadd_rn(v, 1e-10f));  // :124
        acc += static_cast<double>(v);
    }
    const float s2f = static_cast<float>(block_sum_f64(acc, red));  // weighted_std normalises once more, :53
    __syncthreads();  // wo[] written above is re-read below by other threads? no: same thread, same j
    double sw2 = 0.0, mean[D];
#pragma unroll
    for (int i = 0; i < D; ++i) mean[i] = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const float wj = __fdiv_rn(wo[j], s2f);
        sw2 += static_cast<double>(__fmul_rn(wj, wj));
#pragma unroll
        for (int i = 0; i < D; ++i) mean[i] += static_cast<double>(__fmul_rn(wj, Xs[static_cast<int64_t>(j) * D + i]));
    }
    sw2 = block_sum_f64(sw2, red);
    float mf[D];
#pragma unroll
    for (int i = 0; i < D; ++i) mf[i] = static_cast<float>(block_sum_f64(mean[i], red));
    double var[D];
#pragma unroll
    for (int i = 0; i < D; ++i) var[i] = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const float wj = __fdiv_rn(wo[j], s2f);
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const float df = __fsub_rn(Xs[static_cast<int64_t>(j) * D + i], mf[i]);
            var[i] += static_cast<double>(__fmul_rn(wj, __fmul_rn(df, df)));
        }
    }
#pragma unroll
    for (int i = 0; i < D; ++i) var[i] = block_sum_f64(var[i], red);
    if (threadIdx.x == 0) {
        float corr = __fsub_rn(1.0f, static_cast<float>(sw2));  // :64
        if (corr < 1e-10f) corr = 1e-10f;
        const double h = rule == GLABC_BW_SILVERMAN ? pow(static_cast<double>(n) * (D + 2) / 4., -1. / (D + 4))
                                                    : pow(static_cast<double>(n), -1. / (D + 4));
#pragma unroll
        for (int i = 0; i < D; ++i)
            bw_out[s * D + i] = __fmul_rn(static_cast<float>(h), __fsqrt_rn(__fdiv_rn(static_cast<float>(var[i]), corr)));
    }
}

// per-set constants of log_prob
template <int D>
struct KdeConst {
    float bw[D], inv_bw[D];
    float c;       // 0.5 * d * log(2 pi) in float32 (kernel_density.py:120)
    float slog;    // torch.log(bandwidth).sum()  (:121)
};

template <int D>
__device__ __forceinline__ KdeConst<D> kde_const(const float* bw)
{
    KdeConst<D> k;
    float t[D];
#pragma unroll
    for (int i = 0; i < D; ++i) {
        k.bw[i] = bw[i];
        k.inv_bw[i] = __fdiv_rn(1.0f, bw[i]);
        t[i] = logf(bw[i]);
    }
    float s = t[0];
#pragma unroll
    for (int i = 1; i < D; ++i) s = __fadd_rn(s, t[i]);
    k.slog = s;
    k.c = __fmul_rn(__fmul_rn(0.5f, static_cast<float>(D)), logf(6.283185307179586f));
    return k;
}

// one (query, point) term in the reference's operation order, :116-124
template <int D>
__device__ __forceinline__ float kde_term_strict(const KdeConst<D>& k, const float (&x)[D], const float* Xj, float lw)
{
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < D; ++i) {
        const float df = __fdiv_rn(__fsub_rn(x[i], Xj[i]), k.bw[i]);
        s = i == 0 ? __fmul_rn(df, df) : __fadd_rn(s, __fmul_rn(df, df));
    }
    float lk = __fmul_rn(-0.5f, s);
    lk = __fsub_rn(lk, k.c);
    lk = __fsub_rn(lk, k.slog);
    return __fadd_rn(lk, lw);
}

// max-shifted logsumexp over the points of one set for one query, straight from global memory
// (slow path of the tiled kernel for far-away queries; also the single-point evaluation AGLMCMC needs)
template <int D>
__device__ __forceinline__ float kde_log_prob_scan(const KdeConst<D>& k, const float (&x)[D], const float* Xs,
                                                   const float* lw, int n, int first, int stride)
{
    float mx = -INFINITY, sum = 0.0f;  // online logsumexp
    for (int j = first; j < n; j += stride) {
        const float v = kde_term_strict<D>(k, x, Xs + static_cast<int64_t>(j) * D, lw[j]);
        if (v > mx) {
            sum = sum * expf(mx - v) + 1.0f;
            mx = v;
        } else {
            sum += expf(v - mx);
        }
    }
    return (stride == 1) ? mx + logf(sum) : (sum > 0.0f ? mx + logf(sum) : -INFINITY);
}

// combine per-lane (log-partial) values of a strided scan across the warp: logsumexp of 32 logs
__device__ __forceinline__ float warp_logsumexp(float v)
{
    float m = v;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
    if (m == -INFINITY) m = 0.0f;
    float e = expf(v - m);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) e += __shfl_xor_sync(0xffffffffu, e, off);
    return m + logf(e);
}

constexpr int kKdeTile = 1024;     // training points per shared-memory tile
constexpr int kKdeThreads = 128;

// log_prob(), kernel_density.py:96-128.  grid.x = sets * query_tiles, Q queries per thread.
template <int D, int Q, bool STRICT>
__global__ void __launch_bounds__(kKdeThreads) k_kde_logprob(KdeSets S, const float* __restrict__ lw_in,
                                                             const float* __restrict__ x, int64_t m, int64_t qtiles,
                                                             float* __restrict__ out, int ksplit = 1,
                                                             float* __restrict__ partial = nullptr)
{
    // ksplit > 1 (FAST only): grid.x = sets * qtiles * ksplit, CTA (.., ks) sums the point tiles ks, ks + ksplit, ... and
    // stores its partial sum to partial[ks][set][q]; k_kde_logprob_finish adds them in a fixed order (a small query
    // batch then still fills every SM)
    // tile: FAST [kKdeTile][D + 1] = scaled coordinates + log2-weight; STRICT [kKdeTile][D + 1] = raw + log-weight
    //       FAST rows hold a PAIR of points, component-interleaved: (x0_a, x0_b, x1_a, x1_b, ..., lw_a, lw_b), pitch PP
    constexpr int P = D + 1 + ((D + 1) & 1);  // STRICT row: 4 floats for d = 2,3; 2 for d = 1; 6 for d = 4 (8-byte aligned)
    constexpr int PP = (2 * (D + 1) + 3) & ~3;  // FAST pair row, 16-byte aligned: 4 / 8 / 8 / 12 floats for d = 1..4
    constexpr int kTileFloats = (kKdeTile * P > (kKdeTile / 2) * PP) ? kKdeTile * P : (kKdeTile / 2) * PP;
    __shared__ __align__(16) float tile[kTileFloats];
    const int ks = STRICT ? 0 : static_cast<int>(blockIdx.x % ksplit);
    const int64_t bq = STRICT ? blockIdx.x : blockIdx.x / ksplit;
    const int64_t set = bq / qtiles;
    const int64_t qt = bq - set * qtiles;
    if (S.active != nullptr && S.active[set] == 0) return;
    const int n = S.n != nullptr ? S.n[set] : static_cast<int>(S.cap);
    const float* Xs = S.X + set * S.cap * D;
    const float* ws = S.weights + set * S.cap;
    const float* lws = lw_in != nullptr ? lw_in + set * S.cap : nullptr;
    const KdeConst<D> kc = kde_const<D>(S.bw + set * D);
    constexpr float kLog2e = 1.4426950408889634f;

    float xq[Q][D];
    bool valid[Q];
    int64_t qi[Q];
#pragma unroll
    for (int u = 0; u < Q; ++u) {
        qi[u] = (qt * Q + u) * kKdeThreads + threadIdx.x;
        valid[u] = qi[u] < m;
        const int64_t src = valid[u] ? qi[u] : 0;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            const float v = x[(set * m + src) * D + i];
            xq[u][i] = STRICT ? v : v * kc.inv_bw[i];
        }
    }

    float acc[Q], mx[Q];
    f32x2 acc2[Q];   // FAST: partial sums over even / odd points
#pragma unroll
    for (int u = 0; u < Q; ++u) {
        acc[u] = 0.0f;
        mx[u] = -INFINITY;
        acc2[u] = pack2(0.0f, 0.0f);
    }
    const int passes = STRICT ? 2 : 1;
    const int tstep = STRICT ? kKdeTile : kKdeTile * ksplit;
    for (int pass = 0; pass < passes; ++pass) {
        for (int t0 = ks * kKdeTile; t0 < n; t0 += tstep) {
            const int tn = min(kKdeTile, n - t0);
            __syncthreads();
            for (int j = threadIdx.x; j < ((tn + 1) & ~1); j += kKdeThreads) {
                const bool real = j < tn;   // FAST pads an odd tile with a point of weight exp2(-inf) = 0
                const float lwj = !real ? -INFINITY : lws != nullptr ? lws[t0 + j] : logf(__fadd_rn(ws[t0 + j], 1e-10f));  // :124
                if (STRICT && !real) continue;
#pragma unroll
                for (int i = 0; i < D; ++i) {
                    const float v = real ? Xs[static_cast<int64_t>(t0 + j) * D + i] : 0.0f;
                    if constexpr (STRICT) tile[j * P + i] = v;
                    else tile[(j >> 1) * PP + 2 * i + (j & 1)] = v * kc.inv_bw[i];
                }
                if constexpr (STRICT) tile[j * P + D] = lwj;
                else tile[(j >> 1) * PP + 2 * D + (j & 1)] = lwj * kLog2e;
            }
            __syncthreads();
            if constexpr (STRICT) {
                for (int j = 0; j < tn; ++j) {
#pragma unroll
                    for (int u = 0; u < Q; ++u) {
                        const float v = kde_term_strict<D>(kc, xq[u], &tile[j * P], tile[j * P + D]);
                        if (pass == 0) mx[u] = fmaxf(mx[u], v);
                        else acc[u] += expf(v - mx[u]);
                    }
                }
            } else {
                // exp2(lw2_j - 0.5*log2e*|xs - Xs_j|^2): every term <= 1 because the weights are normalised,
                // so the fixed shift M0 = -c - slog needs no running maximum
                // two training points per 64-bit register pair: FADD2 / FMUL2 / FFMA2 (sm_100 packed FP32) halve the issue
                // slots of the distance + exponent arithmetic, leaving MUFU.EX2 (16 / clk / SM) as the bound
                const f32x2 kk = pack2(-0.5f * kLog2e, -0.5f * kLog2e);
                const int np = (tn + 1) >> 1;
#pragma unroll 4
                for (int j = 0; j < np; ++j) {
                    f32x2 p[D + 1];
                    const float* row = &tile[j * PP];
                    if constexpr (D == 1) {
                        const float4 v = *reinterpret_cast<const float4*>(row);
                        p[0] = pack2(v.x, v.y); p[1] = pack2(v.z, v.w);
                    } else {
#pragma unroll
                        for (int i = 0; i + 1 < D + 1; i += 2) {
                            const float4 v = *reinterpret_cast<const float4*>(row + 2 * i);
                            p[i] = pack2(v.x, v.y); p[i + 1] = pack2(v.z, v.w);
                        }
                        if constexpr ((D + 1) & 1) {
                            const float2 v = *reinterpret_cast<const float2*>(row + 2 * D);
                            p[D] = pack2(v.x, v.y);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < Q; ++u) {
                        f32x2 df = sub2(pack2(xq[u][0], xq[u][0]), p[0]);
                        f32x2 s2 = mul2(df, df);
#pragma unroll
                        for (int i = 1; i < D; ++i) {
                            df = sub2(pack2(xq[u][i], xq[u][i]), p[i]);
                            s2 = fma2(df, df, s2);
                        }
                        float t0, t1, e0, e1;
                        unpack2(fma2(s2, kk, p[D]), t0, t1);
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
                        acc2[u] = add2(acc2[u], pack2(e0, e1));
                    }
                }
            }
        }
        if (STRICT && pass == 0) {
#pragma unroll
            for (int u = 0; u < Q; ++u)
                if (mx[u] == -INFINITY || mx[u] == INFINITY) mx[u] = 0.0f;  // torch.logsumexp's handling of an infinite max
        }
    }
    if constexpr (!STRICT) {
#pragma unroll
        for (int u = 0; u < Q; ++u) {
            float a, b;
            unpack2(acc2[u], a, b);
            acc[u] = a + b;
        }
        if (ksplit > 1) {
#pragma unroll
            for (int u = 0; u < Q; ++u)
                if (valid[u]) partial[(static_cast<int64_t>(ks) * S.sets + set) * m + qi[u]] = acc[u];
            return;
        }
    }
#pragma unroll
    for (int u = 0; u < Q; ++u) {
        if (!valid[u]) continue;
        float r;
        if constexpr (STRICT) {
            r = __fadd_rn(mx[u], logf(acc[u]));
        } else {
            if (acc[u] > 1e-30f) {
                r = fmaf(0.6931471805599453f, lg2_approx(acc[u]), -(kc.c + kc.slog));
            } else {
                // far-away query: every term underflowed against the fixed shift — redo it max-shifted (rare)
                float xr[D];
#pragma unroll
                for (int i = 0; i < D; ++i) xr[i] = x[(set * m + qi[u]) * D + i];
                float mxs = -INFINITY, sum = 0.0f;
                for (int j = 0; j < n; ++j) {
                    const float lwj = lws != nullptr ? lws[j] : logf(__fadd_rn(ws[j], 1e-10f));
                    const float v = kde_term_strict<D>(kc, xr, Xs + static_cast<int64_t>(j) * D, lwj);
                    if (v > mxs) {
                        sum = sum * expf(mxs - v) + 1.0f;
                        mxs = v;
                    } else {
                        sum += expf(v - mxs);
                    }
                }
                r = mxs + logf(sum);
            }
        }
        out[set * m + qi[u]] = r;
    }
}

// second half of the point-split log_prob: out[set][q] = log(sum_ks partial[ks][set][q]) - c - slog, with the same
// max-shifted redo for a query whose terms all underflowed
template <int D>
__global__ void __launch_bounds__(256) k_kde_logprob_finish(KdeSets S, const float* __restrict__ lw_in, const float* __restrict__ x,
                                                            int64_t m, int ksplit, const float* __restrict__ partial,
                                                            float* __restrict__ out)
{
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= S.sets * m) return;
    const int64_t set = g / m;
    if (S.active != nullptr && S.active[set] == 0) return;
    float acc = 0.0f;
    for (int k = 0; k < ksplit; ++k) acc += partial[static_cast<int64_t>(k) * S.sets * m + g];
    const KdeConst<D> kc = kde_const<D>(S.bw + set * D);
    float r;
    if (acc > 1e-30f) {
        r = fmaf(0.6931471805599453f, lg2_approx(acc), -(kc.c + kc.slog));
    } else {
        const int n = S.n != nullptr ? S.n[set] : static_cast<int>(S.cap);
        const float* Xs = S.X + set * S.cap * D;
        const float* ws = S.weights + set * S.cap;
        const float* lws = lw_in != nullptr ? lw_in + set * S.cap : nullptr;
        float xr[D];
#pragma unroll
        for (int i = 0; i < D; ++i) xr[i] = x[g * D + i];
        float mxs = -INFINITY, sum = 0.0f;
        for (int j = 0; j < n; ++j) {
            const float lwj = lws != nullptr ? lws[j] : logf(__fadd_rn(ws[j], 1e-10f));
            const float v = kde_term_strict<D>(kc, xr, Xs + static_cast<int64_t>(j) * D, lwj);
            if (v > mxs) {
                sum = sum * expf(mxs - v) + 1.0f;
                mxs = v;
            } else {
                sum += expf(v - mxs);
            }
        }
        r = mxs + logf(sum);
    }
    out[g] = r;
}

// inclusive float64 prefix sums of the normalised weights of every set (native-mode sampling)
static __global__ void __launch_bounds__(256) k_kde_cdf(const float* __restrict__ weights, const int32_t* __restrict__ n_dev,
                                                 const int32_t* __restrict__ active, int64_t cap, double* __restrict__ cdf)
{
    __shared__ double part[256];
    const int64_t s = blockIdx.x;
    if (active != nullptr && active[s] == 0) return;
    const int n = n_dev != nullptr ? n_dev[s] : static_cast<int>(cap);
    const int per = (n + 255) / 256;
    const int lo = threadIdx.x * per, hi = min(n, lo + per);
    double t = 0.0;
    for (int j = lo; j < hi; ++j) t += static_cast<double>(weights[s * cap + j]);
    part[threadIdx.x] = t;
    __syncthreads();
    double base = 0.0;
    for (int i = 0; i < threadIdx.x; ++i) base += part[i];  // 256 sequential adds per thread: exact left-to-right order
    for (int j = lo; j < hi; ++j) {
        base += static_cast<double>(weights[s * cap + j]);
        cdf[s * cap + j] = base;
    }
}

// one native-mode draw q of set `set` (kernel_density.py:142-149): categorical index by inverse CDF on the float64 prefix
// sums, Gaussian jitter; a pure function of (chain id, round, q) through Philox, so any subset of the q's can be generated
template <int D>
__device__ __forceinline__ void kde_draw_native(const KdeSets& S, const double* __restrict__ cdf, int64_t set, int n, int round,
                                                int64_t q, const RoundKeys& rk, uint64_t chain_id_base, float (&out)[D])
{
    const uint64_t gid = chain_id_base + static_cast<uint64_t>(set);
    const Stream st{static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32)};
    const uint4 wu = st.block(rk, static_cast<uint32_t>(round), kSlotAdSample + 2u * static_cast<uint32_t>(q));
    const double u = static_cast<double>(wu.x) * 0x1p-32;
    const double* cs = cdf + set * S.cap;
    int lo = 0, hi = n - 1;  // first j with u < cdf[j]; n - 1 if rounding leaves none
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (u < cs[mid]) hi = mid; else lo = mid + 1;
    }
    const int idx = min(max(lo, 0), n - 1);
    const uint4 wn = st.block(rk, static_cast<uint32_t>(round), kSlotAdSample + 2u * static_cast<uint32_t>(q) + 1u);
    float z[4];
    box_muller(wn.x, wn.y, z[0], z[1]);
    box_muller(wn.z, wn.w, z[2], z[3]);
#pragma unroll
    for (int i = 0; i < D; ++i)
        out[i] = __fadd_rn(S.X[(set * S.cap + idx) * D + i], __fmul_rn(z[i], S.bw[set * D + i]));
}

// sample(), kernel_density.py:130-152
template <int D>
__global__ void __launch_bounds__(256) k_kde_sample(KdeSets S, const double* __restrict__ cdf, int64_t m, RoundKeys rk,
                                                    uint64_t chain_id_base, const int32_t* __restrict__ round_dev,
                                                    const int32_t* __restrict__ idx_tape, const float* __restrict__ noise_tape,
                                                    int64_t tape_set_stride, int64_t tape_set_stride_noise, int64_t tape_q_stride,
                                                    int64_t tape_round_stride_idx, int64_t tape_round_stride_noise,
                                                    int max_round, float* __restrict__ out)
{
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= S.sets * m) return;
    const int64_t set = g / m, q = g - set * m;
    if (S.active != nullptr && S.active[set] == 0) return;
    const int n = S.n != nullptr ? S.n[set] : static_cast<int>(S.cap);
    const int round = round_dev != nullptr ? min(round_dev[set], max_round) : 0;
    int idx;
    float nz[D];
    if (idx_tape != nullptr) {
        idx = idx_tape[round * tape_round_stride_idx + q * tape_q_stride + set * tape_set_stride];
#pragma unroll
        for (int i = 0; i < D; ++i)
            nz[i] = noise_tape[round * tape_round_stride_noise + (q * D + i) * tape_q_stride + set * tape_set_stride_noise];
    } else {
        const uint64_t gid = chain_id_base + static_cast<uint64_t>(set);
        const Stream st{static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32)};
        const uint4 wu = st.block(rk, static_cast<uint32_t>(round), kSlotAdSample + 2u * static_cast<uint32_t>(q));
        const double u = static_cast<double>(wu.x) * 0x1p-32;
        const double* cs = cdf + set * S.cap;
        int lo = 0, hi = n - 1;  // first j with u < cdf[j]; n - 1 if rounding leaves none
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (u < cs[mid]) hi = mid; else lo = mid + 1;
        }
        idx = lo;
        const uint4 wn = st.block(rk, static_cast<uint32_t>(round), kSlotAdSample + 2u * static_cast<uint32_t>(q) + 1u);
        float z[4];
        box_muller(wn.x, wn.y, z[0], z[1]);
        box_muller(wn.z, wn.w, z[2], z[3]);
#pragma unroll
        for (int i = 0; i < D; ++i) nz[i] = z[i];
    }
    idx = min(max(idx, 0), n - 1);
#pragma unroll
    for (int i = 0; i < D; ++i)  // samples = X[indices] + randn * bandwidth, :148-149
        out[(set * m + q) * D + i] = __fadd_rn(S.X[(set * S.cap + idx) * D + i], __fmul_rn(nz[i], S.bw[set * D + i]));
}

// host launchers (kde.cu)
cudaError_t launch_kde_fit(const float* X, const float* w, const int32_t* n, const int32_t* active, int64_t sets, int64_t cap,
                           int dim, int rule, float* weights_out, float* lw_out, float* bw_out, cudaStream_t st);
// point-split factor the public entry should use for (sets, m) queries against <= cap points, and the scratch floats it needs
int kde_logprob_split(int64_t sets, int64_t m, int64_t cap);
cudaError_t launch_kde_logprob(const KdeSets& S, const float* lw, int dim, const float* x, int64_t m, float* out, bool strict,
                               cudaStream_t st, int ksplit = 1, float* partial = nullptr);
cudaError_t launch_kde_cdf(const float* weights, const int32_t* n, const int32_t* active, int64_t sets, int64_t cap, double* cdf,
                           cudaStream_t st);
cudaError_t launch_kde_sample(const KdeSets& S, int dim, const double* cdf, int64_t m, const RoundKeys& rk, uint64_t chain_id_base,
                              const int32_t* round_dev, const int32_t* idx_tape, const float* noise_tape, int64_t tape_set_stride,
                              int64_t tape_set_stride_noise, int64_t tape_q_stride, int64_t tape_round_stride_idx,
                              int64_t tape_round_stride_noise, int max_round, float* out, cudaStream_t st);

}  // namespace glabc
