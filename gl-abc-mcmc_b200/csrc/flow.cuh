// K4 — the RealNVP importance proposal of GLMCMC-NFs (reference GLMCMC_NFs.py:51-61,72,98,127; normflows
// pieces restated in SURVEY.md Appendix C): 32 x [AffineCouplingBlock(MLP([1,128,128,2])), Permute(swap)] over a
// trainable DiagGaussian(2) base.  sample(n) and log_prob(x) for large batches.
//
// The dense piece — the 128x128 hidden layer of every coupling MLP — runs on the 5th-generation tensor cores:
//   * a CTA holds two groups of 128 threads; a group owns tiles of 128 samples = the 128 TMEM lanes of an M=128
//     accumulator (each group has its own A buffer, accumulator columns and mbarrier, so one group's MMA overlaps the
//     other's CUDA-core phases);
//   * per coupling block: thread r computes row r of A = relu(w1 * z1[r] + b1) (the 1->128 layer is a K=1 outer
//     product: CUDA cores) straight into shared memory in the UMMA K-major core-matrix layout, one elected thread
//     issues 16 x tcgen05.mma.kind::tf32 (M128 N128 K8, FP32 accumulate in TMEM) against W2 (pre-packed in the
//     same layout, fetched per block with one cp.async.bulk into shared memory), tcgen05.commit -> mbarrier;
//   * epilogue: each warp pulls its 32 lanes x 128 columns back with tcgen05.ld, adds b2, ReLU, and contracts
//     with the two rows of W3 (N=2: CUDA cores) -> (shift, log-scale) -> affine update of z2, log-det, swap.
//   * W2 of a block is reused for kFlowTilesPerCta tiles (1,024 samples) before the next block's weights are
//     fetched, so weight traffic is 2 KB per sample from L2, nothing from HBM (the whole flow is 2.1 MB).
// TF32 operands: sample() and log_prob() evaluate the SAME deterministic network (same kernel, same rounding), so
// the log-density returned for a sample is the exact density of the map that produced it — importance weights stay
// exact whatever the operand precision; only the agreement with an fp32 evaluation of the weights is ~1e-3.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace glabc {

constexpr int kFlowHidden = 128;
constexpr int kFlowTile = 128;
constexpr int kFlowTilesPerCta = 8;
constexpr int kFlowW2Bytes = kFlowHidden * kFlowHidden * 4;           // 64 KB per block
constexpr int kFlowVecFloats = 768;                                  // w1, b1, b2 [128], w3 [2][128], b3 [2] (+pad)
constexpr int kFlowGroups = 2;                                        // two 128-thread groups, one tile in flight each
constexpr int kFlowThreads = kFlowGroups * kFlowTile;
constexpr int kFlowSmemBytes = (1 + kFlowGroups) * kFlowW2Bytes + kFlowVecFloats * 4 + 3 * kFlowTilesPerCta * kFlowTile * 4 + 64;

struct FlowDev {
    const float* w1;   // [L][128]
    const float* b1;   // [L][128]
    const float* w2p;  // [L][128*128] tf32 bits, UMMA K-major / no-swizzle core-matrix layout (flow_pack_offset)
    const float* b2;   // [L][128]
    const float* w3;   // [L][2][128]
    const float* b3;   // [L][2]
    float base_loc[2], base_log_scale[2];
    int32_t n_blocks;
};

// byte offset of element (row r, k) of a [128][128] fp32 operand in the K-major no-swizzle canonical layout:
// core matrix = 8 rows x 16 bytes, core matrices contiguous along K (LBO = 128 B), 8-row groups 4 KB apart (SBO)
__host__ __device__ constexpr uint32_t flow_pack_offset(uint32_t r, uint32_t k)
{
    return (r >> 3) * 4096u + (k >> 2) * 128u + (r & 7u) * 16u + (k & 3u) * 4u;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo_bytes >> 4) << 16) |
           (static_cast<uint64_t>(sbo_bytes >> 4) << 32) | (1ull << 46);
}

// D[tmem] (+)= A[smem] * B[smem]^T, TF32 operands, FP32 accumulate, M = 128, N = 128, K = 8
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float to_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// round-to-nearest TF32 of a non-negative float in ONE integer add: the tensor core ignores the low 13 mantissa bits
// (truncation), so adding half a TF32 ulp first rounds to nearest (cvt.rna.tf32 is a multi-instruction sequence)
__device__ __forceinline__ float round_tf32_nonneg(float x) { return __uint_as_float(__float_as_uint(x) + 0x1000u); }

// pack W2 [L][out=128][in=128] (torch Linear.weight) into the UMMA layout, rounded to TF32
static __global__ void __launch_bounds__(256) k_flow_pack(const float* __restrict__ w2, float* __restrict__ w2p, int64_t total)
{
    const int64_t g = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (g >= total) return;
    const int64_t l = g / (kFlowHidden * kFlowHidden);
    const uint32_t e = static_cast<uint32_t>(g - l * kFlowHidden * kFlowHidden);
    const uint32_t r = e / kFlowHidden, k = e % kFlowHidden;
    w2p[l * kFlowHidden * kFlowHidden + flow_pack_offset(r, k) / 4] = to_tf32(w2[g]);
}

__device__ __forceinline__ void group_sync(int group)
{
    asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(kFlowTile) : "memory");
}

// SAMPLE: in = eps [n][2] standard normals -> theta [n][2], log q [n]       (NormalizingFlow.sample)
// !SAMPLE: in = theta [n][2] -> log q [n]                                     (NormalizingFlow.log_prob)
// 256 threads = two groups of 128; each group owns its own A buffer, TMEM accumulator (128 columns) and mbarrier and
// walks its tiles independently, so the MMA of one group's tile overlaps the CUDA-core phases (layer 1, epilogue) of the
// other's, and every scheduler holds two warps.
template <bool SAMPLE>
__global__ void __launch_bounds__(kFlowThreads, 1) k_flow(const __grid_constant__ FlowDev W, const float* __restrict__ in,
                                                          int64_t n, float* __restrict__ out_theta, float* __restrict__ out_lq)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    float* sB = reinterpret_cast<float*>(smem);
    float* sVec = reinterpret_cast<float*>(smem + (1 + kFlowGroups) * kFlowW2Bytes);
    float* sState = sVec + kFlowVecFloats;  // [3][tiles][128]: z1, z2, log q
    uint64_t* bars = reinterpret_cast<uint64_t*>(sState + 3 * kFlowTilesPerCta * kFlowTile);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 1 + kFlowGroups);
    const int tid = threadIdx.x, warp = tid >> 5;
    const int group = tid >> 7, gtid = tid & (kFlowTile - 1), gwarp = gtid >> 5;
    float* sA = reinterpret_cast<float*>(smem + (1 + group) * kFlowW2Bytes);
    const uint32_t bar_w = smem_u32(&bars[0]), bar_m = smem_u32(&bars[1 + group]);
    constexpr int TS = kFlowTilesPerCta * kFlowTile;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(static_cast<uint32_t>(kFlowGroups * 128))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 1 + kFlowGroups; ++i) mbar_init(smem_u32(&bars[i]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot + static_cast<uint32_t>(group * 128);  // this group's accumulator columns
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 128, M = 128 (cute::UMMA::InstrDescriptor bit layout)
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t sA_addr = smem_u32(sA), sB_addr = smem_u32(sB);
    uint32_t ph_w = 0, ph_m = 0;
    const float c2 = -1.8378770664093453f;  // -0.5 * 2 * log(2 pi)
    const int L = W.n_blocks;
    const int64_t n_chunks = (n + TS - 1) / TS;

    for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
        int tiles = 0;
        for (int t = 0; t < kFlowTilesPerCta; ++t) {
            if ((chunk * kFlowTilesPerCta + t) * kFlowTile < n) tiles = t + 1;
            if ((t & 1) != group) continue;  // a tile's state is only ever touched by its own group
            const int64_t idx = (chunk * kFlowTilesPerCta + t) * kFlowTile + gtid;
            float a = 0.0f, b = 0.0f, lq = 0.0f;
            if (idx < n) {
                a = in[idx * 2];
                b = in[idx * 2 + 1];
                if (SAMPLE) {  // base DiagGaussian.forward: z = loc + exp(log_scale) * eps, log p from eps
                    lq = c2 - ((W.base_log_scale[0] + 0.5f * (a * a)) + (W.base_log_scale[1] + 0.5f * (b * b)));
                    a = W.base_loc[0] + expf(W.base_log_scale[0]) * a;
                    b = W.base_loc[1] + expf(W.base_log_scale[1]) * b;
                }
            }
            sState[0 * TS + t * kFlowTile + gtid] = a;
            sState[1 * TS + t * kFlowTile + gtid] = b;
            sState[2 * TS + t * kFlowTile + gtid] = lq;
        }
        for (int li = 0; li < L; ++li) {
            const int l = SAMPLE ? li : L - 1 - li;
            __syncthreads();  // both groups are done with the previous block's W2 / vectors (their MMAs were waited for)
            if (tid == 0) {
                mbar_expect_tx(bar_w, kFlowW2Bytes);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    bulk_g2s(sB_addr + q * (kFlowW2Bytes / 4), reinterpret_cast<const uint8_t*>(W.w2p) +
                             static_cast<int64_t>(l) * kFlowW2Bytes + q * (kFlowW2Bytes / 4), kFlowW2Bytes / 4, bar_w);
            }
            if (tid < kFlowHidden) {
                sVec[tid] = W.w1[l * kFlowHidden + tid];
                sVec[128 + tid] = W.b1[l * kFlowHidden + tid];
                sVec[256 + tid] = W.b2[l * kFlowHidden + tid];
            } else {
                const int u = tid - kFlowHidden;
                sVec[384 + u] = W.w3[(l * 2 + 0) * kFlowHidden + u];
                sVec[512 + u] = W.w3[(l * 2 + 1) * kFlowHidden + u];
                if (u < 2) sVec[640 + u] = W.b3[l * 2 + u];
            }
            __syncthreads();
            mbar_wait(bar_w, ph_w);
            ph_w ^= 1u;

            for (int t = group; t < tiles; t += kFlowGroups) {
                float z1 = sState[0 * TS + t * kFlowTile + gtid], z2 = sState[1 * TS + t * kFlowTile + gtid];
                if (!SAMPLE) {  // Permute(swap)^-1 precedes the coupling's inverse
                    const float tmp = z1;
                    z1 = z2;
                    z2 = tmp;
                }
                // ---- layer 1 (K = 1) on CUDA cores, written as the A operand ----
                uint8_t* rowp = reinterpret_cast<uint8_t*>(sA) + (gtid >> 3) * 4096 + (gtid & 7) * 16;
#pragma unroll 8
                for (int kc = 0; kc < kFlowHidden / 4; ++kc) {
                    const float4 w = *reinterpret_cast<const float4*>(&sVec[kc * 4]);
                    const float4 bb = *reinterpret_cast<const float4*>(&sVec[128 + kc * 4]);
                    float4 h;
                    h.x = round_tf32_nonneg(fmaxf(fmaf(w.x, z1, bb.x), 0.0f));
                    h.y = round_tf32_nonneg(fmaxf(fmaf(w.y, z1, bb.y), 0.0f));
                    h.z = round_tf32_nonneg(fmaxf(fmaf(w.z, z1, bb.z), 0.0f));
                    h.w = round_tf32_nonneg(fmaxf(fmaf(w.w, z1, bb.w), 0.0f));
                    *reinterpret_cast<float4*>(rowp + kc * 128) = h;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
                tc_fence_before();
                group_sync(group);
                // ---- layer 2 (128 x 128 x 128) on the tensor cores ----
                if (gtid == 0) {
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < kFlowHidden / 8; ++k) {
                        const uint64_t ad = umma_desc(sA_addr + k * 256, 128, 4096);
                        const uint64_t bd = umma_desc(sB_addr + k * 256, 128, 4096);
                        umma_tf32(tmem, ad, bd, idesc, k > 0 ? 1u : 0u);
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_m) : "memory");
                }
                mbar_wait(bar_m, ph_m);
                ph_m ^= 1u;
                tc_fence_after();
                // ---- bias + ReLU + layer 3 (N = 2) from TMEM; four independent partial sums per output ----
                float p0[4] = {0.0f, 0.0f, 0.0f, 0.0f}, p1[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    uint32_t v[32];
                    tmem_ld32(tmem + (static_cast<uint32_t>(gwarp * 32) << 16) + cb * 32, v);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {  // broadcast LDS.128 of b2 / W3 rows: 3 loads per 4 columns
                        const float4 b2v = *reinterpret_cast<const float4*>(&sVec[256 + cb * 32 + j]);
                        const float4 wa = *reinterpret_cast<const float4*>(&sVec[384 + cb * 32 + j]);
                        const float4 wb = *reinterpret_cast<const float4*>(&sVec[512 + cb * 32 + j]);
                        const float h0 = fmaxf(__uint_as_float(v[j]) + b2v.x, 0.0f), h1 = fmaxf(__uint_as_float(v[j + 1]) + b2v.y, 0.0f);
                        const float h2 = fmaxf(__uint_as_float(v[j + 2]) + b2v.z, 0.0f), h3 = fmaxf(__uint_as_float(v[j + 3]) + b2v.w, 0.0f);
                        p0[0] = fmaf(wa.x, h0, p0[0]); p1[0] = fmaf(wb.x, h0, p1[0]);
                        p0[1] = fmaf(wa.y, h1, p0[1]); p1[1] = fmaf(wb.y, h1, p1[1]);
                        p0[2] = fmaf(wa.z, h2, p0[2]); p1[2] = fmaf(wb.z, h2, p1[2]);
                        p0[3] = fmaf(wa.w, h3, p0[3]); p1[3] = fmaf(wb.w, h3, p1[3]);
                    }
                }
                const float sh = ((p0[0] + p0[1]) + (p0[2] + p0[3])) + sVec[640];  // shift     = param[:, 0::2]
                const float sc = ((p1[0] + p1[1]) + (p1[2] + p1[3])) + sVec[641];  // log-scale = param[:, 1::2]
                float lq = sState[2 * TS + t * kFlowTile + gtid];
                if (SAMPLE) {
                    const float z2n = fmaf(z2, expf(sc), sh);  // z2 * exp(s) + shift; log q -= log det
                    lq -= sc;
                    sState[0 * TS + t * kFlowTile + gtid] = z2n;  // Permute(swap)
                    sState[1 * TS + t * kFlowTile + gtid] = z1;
                } else {
                    const float z2n = (z2 - sh) * expf(-sc);     // inverse; log det = -s
                    lq -= sc;
                    sState[0 * TS + t * kFlowTile + gtid] = z1;
                    sState[1 * TS + t * kFlowTile + gtid] = z2n;
                }
                sState[2 * TS + t * kFlowTile + gtid] = lq;
                tc_fence_before();
                group_sync(group);  // this group's TMEM columns and A buffer are free for its next tile
            }
        }
        for (int t = group; t < tiles; t += kFlowGroups) {
            const int64_t idx = (chunk * kFlowTilesPerCta + t) * kFlowTile + gtid;
            if (idx >= n) continue;
            const float a = sState[0 * TS + t * kFlowTile + gtid], b = sState[1 * TS + t * kFlowTile + gtid];
            float lq = sState[2 * TS + t * kFlowTile + gtid];
            if (SAMPLE) {
                out_theta[idx * 2] = a;
                out_theta[idx * 2 + 1] = b;
            } else {  // + base.log_prob(z)
                const float r0 = (a - W.base_loc[0]) / expf(W.base_log_scale[0]);
                const float r1 = (b - W.base_loc[1]) / expf(W.base_log_scale[1]);
                lq += c2 - ((W.base_log_scale[0] + 0.5f * (r0 * r0)) + (W.base_log_scale[1] + 0.5f * (r1 * r1)));
            }
            out_lq[idx] = lq;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(static_cast<uint32_t>(kFlowGroups * 128))
                     : "memory");
}

cudaError_t launch_flow_pack(const float* w2, float* w2p, int n_blocks, cudaStream_t st);
cudaError_t launch_flow(const FlowDev& W, bool sample, const float* in, int64_t n, float* out_theta, float* out_lq, int sm_count,
                        cudaStream_t st);

}  // namespace glabc
