// Host-callable launchers of the sampler kernels (one translation unit per sampler).
#pragma once
#include <cuda_runtime.h>

#include "sampler_common.cuh"

namespace glabc {

cudaError_t launch_global_mcmc(const ModelConsts& model, const GaussConsts& lp, const GaussConsts& gp, int dim,
                               const RunParams& R, bool strict, bool replay, int layout, int block, cudaStream_t st);

cudaError_t launch_isir(const ModelConsts& model, const GaussConsts& lp, const GaussConsts& ip, int dim, const RunParams& R,
                        bool strict, bool replay, int layout, int block, cudaStream_t st);

struct MalaConsts;
cudaError_t launch_mala(const MalaConsts& K, int dim, const RunParams& R, bool strict, bool replay, int block, cudaStream_t st);

cudaError_t launch_esjd(const float* trace, int layout, int64_t rows, int64_t chains, int dim, float* out,
                        cudaStream_t st);

cudaError_t launch_summarize(const float* stats, int64_t chains, int dim, double* out, cudaStream_t st);

cudaError_t launch_resample(const float* w, int64_t n, int64_t N, float u0, int64_t* idx, unsigned long long* count, double* scratch,
                            cudaStream_t st);

cudaError_t launch_philox_kat(const uint32_t* ctr, const uint32_t* key, int64_t n, uint32_t* out, cudaStream_t st);

}  // namespace glabc
