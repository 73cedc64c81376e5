// K2 instantiations for theta_dim = 2
#include "step_isir.cuh"

namespace glabc {
template cudaError_t launch_isir_dim<2>(const ModelConsts&, const GaussConsts&, const GaussConsts&, const RunParams&, bool, bool, int, int, cudaStream_t);
}  // namespace glabc
