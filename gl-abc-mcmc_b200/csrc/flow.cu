// Host launchers of K4 (RealNVP flow on tcgen05 tensor cores, flow.cuh).
#include "flow.cuh"

namespace glabc {

cudaError_t launch_flow_pack(const float* w2, float* w2p, float* w2p_lo, int n_blocks, cudaStream_t st)
{
    const int64_t total = static_cast<int64_t>(n_blocks) * kFlowHidden * kFlowHidden;
    if (total <= 0) return cudaSuccess;
    k_flow_pack<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(w2, w2p, w2p_lo, total);
    return cudaGetLastError();
}

cudaError_t launch_flow(const FlowDev& W, bool sample, bool precise, const float* in, int64_t n, float* out_theta, float* out_lq,
                        int sm_count, cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    if (precise && (!kFlowF16 || W.w2p_lo == nullptr)) return cudaErrorInvalidValue;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_flow<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlowSmemBytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_flow<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlowSmemBytes);
        if (e != cudaSuccess) return e;
        if constexpr (kFlowF16) {
            e = cudaFuncSetAttribute(k_flow<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlowSmemBytesPrecise);
            if (e != cudaSuccess) return e;
            e = cudaFuncSetAttribute(k_flow<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlowSmemBytesPrecise);
            if (e != cudaSuccess) return e;
        }
        configured = true;
    }
    // tiles per chunk: at most what the shared-memory state holds, sized so that the chunks divide evenly over the SMs (an
    // even number: one tile per group at least)
    const int64_t tiles = (n + kFlowTile - 1) / kFlowTile;
    const int64_t waves = (tiles + int64_t(sm_count) * kFlowTilesPerCta - 1) / (int64_t(sm_count) * kFlowTilesPerCta);
    int64_t tpc = (tiles + sm_count * waves - 1) / (sm_count * waves);   // every SM gets `waves` chunks of (almost) equal size
    tpc = (tpc + 1) / 2 * 2;
    if (tpc < 2) tpc = 2;
    if (tpc > kFlowTilesPerCta) tpc = kFlowTilesPerCta;
    const int64_t chunks = (tiles + tpc - 1) / tpc;
    const unsigned grid = static_cast<unsigned>(chunks < sm_count ? chunks : sm_count);  // persistent: one CTA per SM
    if constexpr (kFlowF16) {
        if (precise) {
            if (sample) k_flow<true, true><<<grid, kFlowThreads, kFlowSmemBytesPrecise, st>>>(W, in, n, out_theta, out_lq, static_cast<int>(tpc));
            else k_flow<false, true><<<grid, kFlowThreads, kFlowSmemBytesPrecise, st>>>(W, in, n, out_theta, out_lq, static_cast<int>(tpc));
            return cudaGetLastError();
        }
    }
    if (sample) k_flow<true, false><<<grid, kFlowThreads, kFlowSmemBytes, st>>>(W, in, n, out_theta, out_lq, static_cast<int>(tpc));
    else k_flow<false, false><<<grid, kFlowThreads, kFlowSmemBytes, st>>>(W, in, n, out_theta, out_lq, static_cast<int>(tpc));
    return cudaGetLastError();
}

}  // namespace glabc
