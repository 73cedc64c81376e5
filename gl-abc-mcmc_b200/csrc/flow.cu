// Host launchers of K4 (RealNVP flow on tcgen05 tensor cores, flow.cuh).
#include "flow.cuh"

namespace glabc {

cudaError_t launch_flow_pack(const float* w2, float* w2p, int n_blocks, cudaStream_t st)
{
    const int64_t total = static_cast<int64_t>(n_blocks) * kFlowHidden * kFlowHidden;
    if (total <= 0) return cudaSuccess;
    k_flow_pack<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(w2, w2p, total);
    return cudaGetLastError();
}

cudaError_t launch_flow(const FlowDev& W, bool sample, const float* in, int64_t n, float* out_theta, float* out_lq, int sm_count,
                        cudaStream_t st)
{
    if (n <= 0) return cudaSuccess;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_flow<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlowSmemBytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_flow<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kFlowSmemBytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int64_t chunks = (n + kFlowTilesPerCta * kFlowTile - 1) / (kFlowTilesPerCta * kFlowTile);
    const unsigned grid = static_cast<unsigned>(chunks < sm_count ? chunks : sm_count);  // persistent: one CTA per SM
    if (sample) k_flow<true><<<grid, kFlowThreads, kFlowSmemBytes, st>>>(W, in, n, out_theta, out_lq);
    else k_flow<false><<<grid, kFlowThreads, kFlowSmemBytes, st>>>(W, in, n, out_theta, out_lq);
    return cudaGetLastError();
}

}  // namespace glabc
