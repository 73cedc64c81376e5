// K1 — the GlobalMCMC chain step, fused: branch draw, local random-walk / global independence
// proposal, simulator draw, discrepancy, Gaussian ABC log-kernel, prior, Metropolis–Hastings
// accept/reject, trace row and statistics.  Reference: GlobalMCMC.py:37-68.
//
// One thread = one chain; the chain state (theta, y, cached log-target) stays in registers for the
// whole launch.  Both branches differ only in how theta' is formed and in the proposal-density
// correction, so the step is branch-free (selects), which matters because the branch is a
// per-thread coin flip.
#include "launch.cuh"
#include "sampler_common.cuh"

namespace glabc {

struct GlobalConsts {
    ModelConsts model;
    GaussConsts lp;  // Local_Proposal
    GaussConsts gp;  // Global_Proposal
};

template <int D>
struct Draws {
    bool is_global;
    float u_b;  // branch uniform (replay: from the tape; native: only materialised for tape_dump)
    float u_a;  // uniform of the accept test (log taken by the consumer)
    float eps_p[D], eps_s[D];
};

// replay: the tape holds the reference's own draws, [step][slot][chain]
template <int D>
__device__ __forceinline__ Draws<D> read_tape(const RunParams& r, uint32_t step_in_launch, int32_t chain)
{
    constexpr int kSlots = GLABC_TAPE_GLOBAL_SLOTS(D, D);
    const float* t = r.tape32 + (static_cast<int64_t>(step_in_launch) * kSlots) * r.n_chains + chain;
    Draws<D> d;
    d.u_b = __ldg(t);
    d.is_global = d.u_b < r.gf;  // GlobalMCMC.py:39 — float32 compare, strict
#pragma unroll
    for (int k = 0; k < D; ++k) d.eps_p[k] = __ldg(t + static_cast<int64_t>(1 + k) * r.n_chains);
#pragma unroll
    for (int k = 0; k < D; ++k) d.eps_s[k] = __ldg(t + static_cast<int64_t>(1 + D + k) * r.n_chains);
    d.u_a = __ldg(t + static_cast<int64_t>(1 + 2 * D) * r.n_chains);
    return d;
}

template <int D>
__device__ __forceinline__ void normals_for_step(const Stream& s, const RoundKeys& rk, uint32_t step, float (&eps_p)[D], float (&eps_s)[D])
{
    constexpr int kGroups = (2 * D + 3) / 4;
    float z[kGroups * 4];
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
        const uint4 w = s.block(rk, step, kSlotNormal + g);
        box_muller(w.x, w.y, z[4 * g + 0], z[4 * g + 1]);
        box_muller(w.z, w.w, z[4 * g + 2], z[4 * g + 3]);
    }
#pragma unroll
    for (int k = 0; k < D; ++k) {
        eps_p[k] = z[k];
        eps_s[k] = z[D + k];
    }
}

template <int D, bool STRICT, bool REPLAY, int LAYOUT>
__global__ void __launch_bounds__(256) k_global_mcmc(const __grid_constant__ GlobalConsts K,
                                                     const __grid_constant__ RunParams R)
{
    using Writer = typename WriterFor<D, LAYOUT>::type;
    extern __shared__ float smem[];
    const int32_t chain = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = chain < R.n_chains;
    const int32_t cidx = active ? chain : R.n_chains - 1;  // tail lanes shadow the last chain, never store

    float theta[D], y[D];
#pragma unroll
    for (int k = 0; k < D; ++k) {
        theta[k] = R.theta[static_cast<int64_t>(cidx) * D + k];
        y[k] = R.y[static_cast<int64_t>(cidx) * D + k];
    }
    // The reference recomputes prior(theta_old), kernel(y_old) and q_g(theta_old) every step
    // (GlobalMCMC.py:45-46,61); same inputs give the same float32 values, so they are cached.
    float prior_old = model_prior<D, STRICT>(K.model, theta);
    float kern_old = model_log_kernel<D, STRICT>(K.model, y);

    Writer writer(R, cidx, active, smem + (threadIdx.x >> 5) * Writer::smem_floats_per_warp);
    if (R.write_row0) writer.put(R, R.first_step - 1u, theta);

    ChainStats<D> stats;
    const Stream stream = chain_stream(R, cidx);

    auto step = [&](uint32_t i, const Draws<D>& dr) {
        // proposal: global = Global_Proposal.forward(1) (:40); local = Local_Proposal.sample(1) + theta (:56)
        float th_g[D], th_l[D], theta_p[D], y_p[D];
        const float lq_p = gauss_forward<D, STRICT>(K.gp, dr.eps_p, th_g);
        (void)gauss_forward<D, STRICT>(K.lp, dr.eps_p, th_l);
#pragma unroll
        for (int k = 0; k < D; ++k) {
            const float loc = STRICT ? __fadd_rn(th_l[k], theta[k]) : th_l[k] + theta[k];
            theta_p[k] = dr.is_global ? th_g[k] : loc;
        }
        model_simulate<D, STRICT>(K.model, theta_p, dr.eps_s, y_p);           // :41 / :57
        const float prior_p = model_prior<D, STRICT>(K.model, theta_p);        // :44 / :60
        const float kern_p = model_log_kernel<D, STRICT>(K.model, y_p);
        const float lq_old = gauss_log_prob<D, STRICT>(K.gp, theta);           // :45
        float log_acc;
        if constexpr (STRICT) {
            // left-to-right, exactly as written at GlobalMCMC.py:44-46 and :60-61
            const float base = __fadd_rn(prior_p, kern_p);
            const float g = __fsub_rn(__fsub_rn(__fsub_rn(__fadd_rn(base, lq_old), lq_p), prior_old), kern_old);
            const float l = __fsub_rn(__fsub_rn(base, prior_old), kern_old);
            log_acc = dr.is_global ? g : l;
        } else {
            const float corr = dr.is_global ? lq_old - lq_p : 0.0f;
            log_acc = (prior_p + kern_p) + corr - (prior_old + kern_old);
        }
        const float log_w = STRICT ? logf(dr.u_a) : log_approx(dr.u_a);            // :47 / :62
        const bool accept = log_w < log_acc;                                   // :49 strict <, NaN rejects

        float prev[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
            prev[k] = theta[k];
            theta[k] = accept ? theta_p[k] : theta[k];
            y[k] = accept ? y_p[k] : y[k];
        }
        prior_old = accept ? prior_p : prior_old;
        kern_old = accept ? kern_p : kern_old;
        stats.update(dr.is_global, accept, theta, prev);
        writer.put(R, i, theta);                                               // :53 / :68

        if constexpr (REPLAY) {
            if (R.debug != nullptr && active) {
                float* g = R.debug + static_cast<int64_t>(i - R.first_step) * GLABC_DEBUG_SLOTS * R.n_chains + chain;
                g[0] = static_cast<float>(static_cast<int>(dr.is_global) | (static_cast<int>(accept) << 1));
                g[static_cast<int64_t>(1) * R.n_chains] = prior_p;
                g[static_cast<int64_t>(2) * R.n_chains] = kern_p;
                g[static_cast<int64_t>(3) * R.n_chains] = log_acc;
            }
        } else {
            if (R.tape_dump != nullptr && active) {  // native draws, in tape layout, so a replay can re-run this chain
                constexpr int kSlots = GLABC_TAPE_GLOBAL_SLOTS(D, D);
                float* t = R.tape_dump + static_cast<int64_t>(i - R.first_step) * kSlots * R.n_chains + chain;
                t[0] = dr.u_b;
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    t[static_cast<int64_t>(1 + k) * R.n_chains] = dr.eps_p[k];
                    t[static_cast<int64_t>(1 + D + k) * R.n_chains] = dr.eps_s[k];
                }
                t[static_cast<int64_t>(1 + 2 * D) * R.n_chains] = dr.u_a;
            }
        }
    };

    if (R.last_step >= R.first_step) {
        if constexpr (REPLAY) {
            for (uint32_t i = R.first_step; i <= R.last_step; ++i) step(i, read_tape<D>(R, i - R.first_step, cidx));
        } else {
            // one uniform block serves two steps: walk aligned pairs (2j, 2j+1)
            for (uint32_t j = R.first_step >> 1; j <= (R.last_step >> 1); ++j) {
                const uint4 u = stream.block(R.rk, j, kSlotUniform);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint32_t i = 2u * j + h;
                    if (i < R.first_step || i > R.last_step) continue;  // launch-uniform
                    Draws<D> dr;
                    dr.is_global = ((h ? u.z : u.x) >> 8) < R.gf_threshold;
                    dr.u_b = u24(h ? u.z : u.x);
                    dr.u_a = u24(h ? u.w : u.y);
                    normals_for_step<D>(stream, R.rk, i, dr.eps_p, dr.eps_s);
                    step(i, dr);
                }
            }
        }
    }
    writer.finish(R);

    if (active) {
#pragma unroll
        for (int k = 0; k < D; ++k) {
            R.theta[static_cast<int64_t>(chain) * D + k] = theta[k];
            R.y[static_cast<int64_t>(chain) * D + k] = y[k];
        }
        if (R.stats != nullptr)
            stats.store(R.stats + static_cast<int64_t>(chain) * GLABC_NSTATS(D), R.last_step + 1u - R.first_step);
    }
}

template <int D, bool STRICT, bool REPLAY>
static cudaError_t launch_layout(const GlobalConsts& K, const RunParams& R, int layout, int block, cudaStream_t st)
{
    const int grid = (R.n_chains + block - 1) / block;
    const int warps = block / 32;
    switch (layout) {
    case GLABC_TRACE_NONE:
        k_global_mcmc<D, STRICT, REPLAY, GLABC_TRACE_NONE><<<grid, block, 0, st>>>(K, R);
        break;
    case GLABC_TRACE_TIME_MAJOR:
        k_global_mcmc<D, STRICT, REPLAY, GLABC_TRACE_TIME_MAJOR><<<grid, block, 0, st>>>(K, R);
        break;
    case GLABC_TRACE_CHAIN_MAJOR: {
        const size_t smem = sizeof(float) * ChainMajorWriter<D>::smem_floats_per_warp * warps;
        auto kern = k_global_mcmc<D, STRICT, REPLAY, GLABC_TRACE_CHAIN_MAJOR>;
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, block, smem, st>>>(K, R);
        break;
    }
    default:
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int D>
static cudaError_t launch_dim(const GlobalConsts& K, const RunParams& R, bool strict, bool replay, int layout,
                              int block, cudaStream_t st)
{
    if (replay) {
        return strict ? launch_layout<D, true, true>(K, R, layout, block, st)
                      : launch_layout<D, false, true>(K, R, layout, block, st);
    }
    return strict ? launch_layout<D, true, false>(K, R, layout, block, st)
                  : launch_layout<D, false, false>(K, R, layout, block, st);
}

cudaError_t launch_global_mcmc(const ModelConsts& model, const GaussConsts& lp, const GaussConsts& gp, int dim,
                               const RunParams& R, bool strict, bool replay, int layout, int block, cudaStream_t st)
{
    GlobalConsts K{model, lp, gp};
    switch (dim) {
    case 1: return launch_dim<1>(K, R, strict, replay, layout, block, st);
    case 2: return launch_dim<2>(K, R, strict, replay, layout, block, st);
    case 3: return launch_dim<3>(K, R, strict, replay, layout, block, st);
    case 4: return launch_dim<4>(K, R, strict, replay, layout, block, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace glabc
