// Cost of tcgen05.mma variants on one SM (cycles per instruction, one issuing thread, REP sets back to back):
//   nvcc -gencode arch=compute_100a,code=sm_100a -I gl-abc-mcmc_b200/csrc -I include -o gpurun_out/mma_cost profiles/micro/mma_cost.cu
#include <cstdio>
#include <cstdlib>
#include "flow.cuh"
using namespace glabc;

constexpr int REP = 32;
// mode 0: TS N=128 chained (8 per set); 1: TS N=16 chained; 2: TS N=16, 4 independent accumulators; 3: SS N=16 (A smem), 4 accumulators;
// 4: SS N=128 chained; 5: TS N=64 chained; 6: SS N=16 chained; 7: TS N=32, 4 accumulators
template <int mode, bool ELECT>
__global__ void __launch_bounds__(128, 1) k(long long* out)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[1];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    {   // A operand in TMEM: some finite values
        uint32_t hv[32];
        for (int j = 0; j < 32; ++j) hv[j] = 0x3C003C00u;
        tmem_st32(tmem + 256u + (static_cast<uint32_t>(warp * 32) << 16), hv);
        tmem_st32(tmem + 288u + (static_cast<uint32_t>(warp * 32) << 16), hv);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (ELECT ? warp == 0 : tid == 0) {
        const uint32_t sa = smem_u32(smem), sb = sa + 32768;
        auto idesc = [](uint32_t n) { return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24); };
        bool lead = true;
        if (ELECT) {
            uint32_t p;
            asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(p));
            lead = p != 0;
        }
        const long long t0 = clock64();
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint64_t bd = umma_desc(sb + kk * 256, 128, 2048), ad = umma_desc(sa + kk * 256, 128, 2048);
                if (lead) {
                    if (mode == 0) umma_f16_ts(tmem, tmem + 256u + kk * 8, bd, idesc(128), kk > 0);
                    if (mode == 1) umma_f16_ts(tmem, tmem + 256u + kk * 8, bd, idesc(16), kk > 0);
                    if (mode == 2) umma_f16_ts(tmem + (kk & 3) * 16, tmem + 256u + kk * 8, bd, idesc(16), kk >= 4);
                    if (mode == 3) umma_f16_ss(tmem + (kk & 3) * 16, ad, bd, idesc(16), kk >= 4);
                    if (mode == 4) umma_f16_ss(tmem, ad, bd, idesc(128), kk > 0);
                    if (mode == 5) umma_f16_ts(tmem, tmem + 256u + kk * 8, bd, idesc(64), kk > 0);
                    if (mode == 6) umma_f16_ss(tmem, ad, bd, idesc(16), kk > 0);
                    if (mode == 7) umma_f16_ts(tmem + (kk & 3) * 32, tmem + 256u + kk * 8, bd, idesc(32), kk >= 4);
                }
            }
        }
        const long long t1 = clock64();
        if (lead) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bars[0])) : "memory");
        mbar_wait(smem_u32(&bars[0]), 0);
        const long long t2 = clock64();
        if (lead) {
            out[0] = t1 - t0;
            out[1] = t2 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main()
{
    long long* d;
    cudaMalloc(&d, 16 * sizeof(long long));
    const char* names[] = {"TS N=128 chained", "TS N=16 chained", "TS N=16 x4 acc", "SS N=16 x4 acc", "SS N=128 chained", "TS N=64 chained", "SS N=16 chained", "TS N=32 x4 acc"};
    auto run = [&](auto kern, const char* name, const char* how) {
        for (int pass = 0; pass < 2; ++pass) {
            kern<<<1, 128, 65536>>>(d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
            long long h[2];
            cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
            if (pass) printf("%-18s %-8s issue %.1f cyc/instr, complete %.1f cyc/instr\n", name, how, double(h[0]) / (8 * REP), double(h[1]) / (8 * REP));
        }
    };
#define RUN(m) \
    cudaFuncSetAttribute(k<m, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); \
    cudaFuncSetAttribute(k<m, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);  \
    run(k<m, false>, names[m], "thread0"); run(k<m, true>, names[m], "elect");
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
    return 0;
}
