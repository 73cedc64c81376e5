// Flat layout of the flow's FP32 parameters (= of its gradient and of each Adam moment), shared by abi.cu and flow_train.cuh.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace glabc {

struct FlowParamLayout {
    int64_t w1, b1, w2, b2, w3, b3, loc, log_scale, total;
    __host__ __device__ explicit FlowParamLayout(int L)
    {
        const int64_t H = 128;
        w1 = 0; b1 = w1 + L * H; w2 = b1 + L * H; b2 = w2 + L * H * H; w3 = b2 + L * H; b3 = w3 + L * 2 * H;
        loc = b3 + L * 2; log_scale = loc + 2; total = log_scale + 2;
    }
};

struct FlowDev;
cudaError_t launch_flow_bwd(const FlowDev& W, const float* z_final, int64_t n, float* partial, int n_slices, cudaStream_t st);
cudaError_t launch_flow_grad_reduce(const float* partial, int n_slices, int64_t total, int64_t n, float* grad, cudaStream_t st);
cudaError_t launch_flow_loss(const float* lq, int64_t n, float* loss, cudaStream_t st);
cudaError_t launch_flow_adam(float* p, float* m, float* v, const float* g, int64_t total, const float* loss, float lr, float beta1, float beta2,
                             float eps, float wd, int64_t t, cudaStream_t st);
int flow_train_chunk_samples();

}  // namespace glabc
