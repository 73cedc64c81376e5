// Systematic resampling — `resample(W, N)` of GLMCMC_NFs.py:29-40 (same code at AGLMCMC.py:30-41), the step that picks the
// training set of the flow (GLMCMC_NFs.py:114-116).  Reference semantics, pinned by tests/golden/resample.npz:
//   u_i  = (U + i) / N  in float32 (a Python float plus an int64 arange gives a float32 tensor), i = 0..N-1
//   Psum = torch.cumsum(W)  — on the CPU build a float64 running sum rounded to float32 per element (acc_type<float> = double)
//   index j is emitted once for every u_i with Psum[j-1] <= u_i < Psum[j]; u_i at or beyond the last cumulative weight are
//   dropped, so fewer than N indices come back when the weights sum to less than one.
// Device form for W of any length (the pooled block holds up to 1e9 weights): float64 block sums, a float64 scan of the block
// sums, and one thread per u_i — binary search over the block prefixes, then the float64 running sum inside the block, each
// prefix rounded to float32 before the compare as the reference's is.  A parallel float64 sum differs from the sequential
// one by ~1e-16 relative, invisible after the rounding to float32.
#include <cstdint>
#include <cuda_runtime.h>

namespace glabc {

constexpr int kRsBlock = 4096;   // weights per block sum

__global__ void __launch_bounds__(256) k_rs_block_sums(const float* __restrict__ w, int64_t n, double* __restrict__ bsum)
{
    const int64_t base = static_cast<int64_t>(blockIdx.x) * kRsBlock;
    double s = 0.0;
    for (int e = threadIdx.x; e < kRsBlock; e += 256) {
        const int64_t j = base + e;
        if (j < n) s += static_cast<double>(w[j]);
    }
    __shared__ double part[8];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += part[i];
        bsum[blockIdx.x] = t;
    }
}

// in place: bsum[b] -> sum of the blocks before b (exclusive); one thread — nb <= 2.5e5 float64 adds even for 1e9 weights
__global__ void k_rs_scan(double* __restrict__ bsum, int64_t nb, double* __restrict__ total)
{
    double run = 0.0;
    for (int64_t b = 0; b < nb; ++b) {
        const double v = bsum[b];
        bsum[b] = run;
        run += v;
    }
    *total = run;
}

__global__ void __launch_bounds__(256) k_rs_search(const float* __restrict__ w, int64_t n, const double* __restrict__ excl, int64_t nb,
                                                  const double* __restrict__ total, int64_t N, float u0, int64_t* __restrict__ idx,
                                                  unsigned long long* __restrict__ count)
{
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float u = __fdiv_rn(__fadd_rn(u0, static_cast<float>(i)), static_cast<float>(N));   // GLMCMC_NFs.py:31
    int64_t found = n;
    if (n > 0) {
        // first block whose END prefix (rounded to float32) exceeds u: prefixes are non-decreasing
        int64_t lo = 0, hi = nb;   // answer in [lo, hi]
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            const double end = mid + 1 < nb ? excl[mid + 1] : *total;
            if (static_cast<float>(end) > u) hi = mid; else lo = mid + 1;
        }
        for (int64_t b = lo; b < nb && found == n; ++b) {   // (a second block only if the two summation orders disagree on the last ulp)
            double run = excl[b];
            const int64_t j1 = min(n, (b + 1) * kRsBlock);
            for (int64_t j = b * kRsBlock; j < j1; ++j) {
                run += static_cast<double>(w[j]);
                if (static_cast<float>(run) > u) { found = j; break; }   // :34 Psum[j] > u[i]
            }
        }
    }
    idx[i] = found;
    if (found < n) atomicAdd(count, 1ull);
}

cudaError_t launch_resample(const float* w, int64_t n, int64_t N, float u0, int64_t* idx, unsigned long long* count, double* scratch,
                            cudaStream_t st)
{
    // scratch: [nb] block sums / exclusive prefixes, then the total
    const int64_t nb = (n + kRsBlock - 1) / kRsBlock;
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    if (nb > 0) k_rs_block_sums<<<static_cast<unsigned>(nb), 256, 0, st>>>(w, n, scratch);
    k_rs_scan<<<1, 1, 0, st>>>(scratch, nb, scratch + nb);
    if (N > 0) k_rs_search<<<static_cast<unsigned>((N + 255) / 256), 256, 0, st>>>(w, n, scratch, nb, scratch + nb, N, u0, idx, count);
    return cudaGetLastError();
}

}  // namespace glabc
