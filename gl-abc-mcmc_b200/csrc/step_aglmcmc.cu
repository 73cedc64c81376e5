// Orchestration of K6 (AGLMCMC): step / adapt rounds on one stream, dispatched over theta_dim and model family.
#include "step_aglmcmc.cuh"

namespace glabc {

template <int D, int FAMILY>
static cudaError_t run_family(const AgConsts& K, const AgWorkspace& W, const AgTapes& T, const RunParams& R, int init,
                              int kde_rule, bool strict, bool replay, int layout, int block, cudaStream_t st)
{
    const int64_t C = W.C, B = W.B;
    const unsigned g_chain = static_cast<unsigned>((C + 255) / 256);
    const int64_t cb = (C * B + 255) / 256;
    if (cb > 0x7fffffffll) return cudaErrorInvalidValue;
    const unsigned g_cb = static_cast<unsigned>(cb);
    const unsigned g_step = static_cast<unsigned>((C + block - 1) / block);
    cudaError_t e;
    if (init) {
        k_ag_reset<<<g_chain, 256, 0, st>>>(W, R.first_step);
        if (replay) k_ag_block<D, FAMILY, true, true><<<g_cb, 256, 0, st>>>(K, R, W, T);
        else k_ag_block<D, FAMILY, true, false><<<g_cb, 256, 0, st>>>(K, R, W, T);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    if (R.write_row0 && layout != GLABC_TRACE_NONE) k_ag_row0<D><<<g_chain, 256, 0, st>>>(R, C, layout);
    const int64_t n_steps = static_cast<int64_t>(R.last_step) - R.first_step + 1;
    const int64_t rounds = (n_steps > 0 ? n_steps : 0) / K.S + 2;
    KdeSets S{W.kde_X, W.kde_wn, W.kde_bw, W.kde_n, W.pending, C, B};
    for (int64_t r = 0; r < rounds; ++r) {
        if (strict) {
            if (replay) k_ag_step<D, FAMILY, true, true><<<g_step, block, 0, st>>>(K, R, W, layout);
            else k_ag_step<D, FAMILY, true, false><<<g_step, block, 0, st>>>(K, R, W, layout);
        } else {
            if (replay) k_ag_step<D, FAMILY, false, true><<<g_step, block, 0, st>>>(K, R, W, layout);
            else k_ag_step<D, FAMILY, false, false><<<g_step, block, 0, st>>>(K, R, W, layout);
        }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (r == rounds - 1) break;
        // ---- adaptation of the chains that paused (AGLMCMC.py:170-249); every kernel skips the others ----
        k_ag_adapt<D><<<static_cast<unsigned>(C), 256, 0, st>>>(K, W);
        if ((e = launch_kde_fit(W.kde_X, W.kde_w, W.kde_n, W.pending, C, B, D, kde_rule, W.kde_wn, W.kde_lw, W.kde_bw, st)) != cudaSuccess)
            return e;
        if (!replay && (e = launch_kde_cdf(W.kde_wn, W.kde_n, W.pending, C, B, W.cdf, st)) != cudaSuccess) return e;
        const uint64_t base = (static_cast<uint64_t>(R.chain_hi0) << 32) | R.chain_lo0;
        if (replay) {
            e = launch_kde_sample(S, D, W.cdf, 4 * B, R.rk, base, W.n_adapt, T.ad_idx, T.ad_noise, 1, 1, C, 4 * B * C, 4 * B * D * C,
                                  T.tape_rounds - 1, W.smp, st);
            if (e != cudaSuccess) return e;
            k_ag_filter<D><<<static_cast<unsigned>(C), 256, 0, st>>>(K, W);
        } else {
            k_ag_sample_filter<D><<<static_cast<unsigned>(C), 256, 0, st>>>(K, W, S, R.rk, base);
        }
        if ((e = launch_kde_logprob(S, W.kde_lw, D, W.blk_theta, B, W.blk_lq, strict, st)) != cudaSuccess) return e;
        if (replay) k_ag_block<D, FAMILY, false, true><<<g_cb, 256, 0, st>>>(K, R, W, T);
        else k_ag_block<D, FAMILY, false, false><<<g_cb, 256, 0, st>>>(K, R, W, T);
        k_ag_commit<D><<<g_chain, 256, 0, st>>>(W, T);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int D>
static cudaError_t run_dim(const AgConsts& K, const AgWorkspace& W, const AgTapes& T, const RunParams& R, int init, int kde_rule,
                           bool strict, bool replay, int layout, int block, cudaStream_t st)
{
    if (K.model.family == GLABC_MODEL_ABS_NORMAL)
        return run_family<D, GLABC_MODEL_ABS_NORMAL>(K, W, T, R, init, kde_rule, strict, replay, layout, block, st);
    return run_family<D, GLABC_MODEL_ID_NORMAL>(K, W, T, R, init, kde_rule, strict, replay, layout, block, st);
}

template <int D, int FAMILY>
static cudaError_t block_isir_family(const AgConsts& K, const AgWorkspace& W, const RunParams& R, bool strict, int layout, int block,
                                     cudaStream_t st)
{
    const unsigned g_step = static_cast<unsigned>((W.C + block - 1) / block);
    if (R.write_row0 && layout != GLABC_TRACE_NONE) k_ag_row0<D><<<static_cast<unsigned>((W.C + 255) / 256), 256, 0, st>>>(R, W.C, layout);
    if (strict) k_ag_step<D, FAMILY, true, false, true><<<g_step, block, 0, st>>>(K, R, W, layout);
    else k_ag_step<D, FAMILY, false, false, true><<<g_step, block, 0, st>>>(K, R, W, layout);
    return cudaGetLastError();
}

cudaError_t launch_block_isir(const AgConsts& K, const AgWorkspace& W, const RunParams& R, int dim, bool strict, int layout,
                              int block, cudaStream_t st)
{
    const bool ab = K.model.family == GLABC_MODEL_ABS_NORMAL;
    switch (dim) {
    case 1: return ab ? block_isir_family<1, GLABC_MODEL_ABS_NORMAL>(K, W, R, strict, layout, block, st)
                      : block_isir_family<1, GLABC_MODEL_ID_NORMAL>(K, W, R, strict, layout, block, st);
    case 2: return ab ? block_isir_family<2, GLABC_MODEL_ABS_NORMAL>(K, W, R, strict, layout, block, st)
                      : block_isir_family<2, GLABC_MODEL_ID_NORMAL>(K, W, R, strict, layout, block, st);
    case 3: return ab ? block_isir_family<3, GLABC_MODEL_ABS_NORMAL>(K, W, R, strict, layout, block, st)
                      : block_isir_family<3, GLABC_MODEL_ID_NORMAL>(K, W, R, strict, layout, block, st);
    case 4: return ab ? block_isir_family<4, GLABC_MODEL_ABS_NORMAL>(K, W, R, strict, layout, block, st)
                      : block_isir_family<4, GLABC_MODEL_ID_NORMAL>(K, W, R, strict, layout, block, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_block_weights(const AgConsts& K, const AgWorkspace& W, const RunParams& R, int dim, uint32_t round, cudaStream_t st)
{
    const int64_t cb = (W.C * W.B + 255) / 256;
    if (cb > 0x7fffffffll) return cudaErrorInvalidValue;
    const unsigned g = static_cast<unsigned>(cb);
    const bool ab = K.model.family == GLABC_MODEL_ABS_NORMAL;
#define GLABC_BW(DD)                                                                                          \
    if (ab) k_blk_weights<DD, GLABC_MODEL_ABS_NORMAL><<<g, 256, 0, st>>>(K, R, W, round);                     \
    else k_blk_weights<DD, GLABC_MODEL_ID_NORMAL><<<g, 256, 0, st>>>(K, R, W, round)
    switch (dim) {
    case 1: GLABC_BW(1); break;
    case 2: GLABC_BW(2); break;
    case 3: GLABC_BW(3); break;
    case 4: GLABC_BW(4); break;
    default: return cudaErrorInvalidValue;
    }
#undef GLABC_BW
    return cudaGetLastError();
}

cudaError_t launch_aglmcmc(const AgConsts& K, const AgWorkspace& W, const AgTapes& T, const RunParams& R, int dim, int init,
                           int kde_rule, bool strict, bool replay, int layout, int block, cudaStream_t st)
{
    switch (dim) {
    case 1: return run_dim<1>(K, W, T, R, init, kde_rule, strict, replay, layout, block, st);
    case 2: return run_dim<2>(K, W, T, R, init, kde_rule, strict, replay, layout, block, st);
    case 3: return run_dim<3>(K, W, T, R, init, kde_rule, strict, replay, layout, block, st);
    case 4: return run_dim<4>(K, W, T, R, init, kde_rule, strict, replay, layout, block, st);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace glabc
