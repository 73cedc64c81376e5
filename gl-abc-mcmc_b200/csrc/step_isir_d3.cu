// K2 instantiations for theta_dim = 3
#include "step_isir.cuh"

namespace glabc {
template cudaError_t launch_isir_dim<3>(const ModelConsts&, const GaussConsts&, const GaussConsts&, const RunParams&, bool, bool, int, int, cudaStream_t);
}  // namespace glabc
