// TEST INFRASTRUCTURE: an independent Philox4x32-10 (NVIDIA cuRAND's device implementation) to pin
// the product's generator against.  curand_init(seed, subsequence, offset) sets key = seed and
// counter = (offset/4 lo, offset/4 hi, subsequence lo, subsequence hi); curand4() returns the block.
#include <cstdint>
#include <cuda_runtime.h>
#include <curand_kernel.h>

__global__ void k_curand_blocks(unsigned long long seed, const unsigned long long* subseq,
                                const unsigned long long* block, int n, uint32_t* out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    curandStatePhilox4_32_10_t st;
    curand_init(seed, subseq[i], 4ull * block[i], &st);
    const uint4 r = curand4(&st);
    out[4 * i] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

extern "C" __attribute__((visibility("default")))
int curand_philox_blocks(unsigned long long seed, const unsigned long long* subseq_dev,
                         const unsigned long long* block_dev, int n, uint32_t* out_dev)
{
    k_curand_blocks<<<(n + 127) / 128, 128>>>(seed, subseq_dev, block_dev, n, out_dev);
    return (int)cudaDeviceSynchronize();
}
