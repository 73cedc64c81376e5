// Pieces shared by the sampler kernels: the launch description, the trace writers and the
// per-chain statistics accumulators.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/glabc.h"
#include "model.cuh"
#include "philox.cuh"

namespace glabc {

// device view of glabc_run_t (validated and narrowed by the host side)
struct RunParams {
    int32_t n_chains;
    uint32_t first_step;   // first loop index i to perform (= step_base + 1)
    uint32_t last_step;    // last loop index (inclusive); last < first means no transition
    uint32_t chain_lo0, chain_hi0;  // global id of chain 0 (low/high words)
    RoundKeys rk;          // Philox round keys of the run's seed
    float gf;
    uint32_t gf_thr_hi;     // native mode: global iff (U_b16 << 16 | junk) < gf_thr_hi, U_b16 the 16-bit
                            // branch uniform and gf_thr_hi = ceil(gf * 2^16) << 16   (cf. B-15/B-16)
    int32_t gf_all_global;  // gf >= 1: every step is global (the threshold would not fit 32 bits)
    int32_t write_row0;
    int64_t trace_rows, trace_chains, trace_chain_off, trace_row_base;
    float* theta;
    float* y;
    float* aux;
    float* trace;
    float* stats;
    const float* tape32;
    const double* tape64;
    float* debug;
    float* tape_dump;
    double* tape64_dump;
    int32_t n_candidates;
    // GLMALA (K3)
    int32_t trace_layout;
    double* state64;
    const float* tape_grad0;
    float* tape_grad0_dump;
    double* debug64;
};

__device__ __forceinline__ Stream chain_stream(const RunParams& r, int32_t chain)
{
    // 64-bit add of the chain index to the global base id
    const uint64_t gid = (static_cast<uint64_t>(r.chain_hi0) << 32 | r.chain_lo0) + static_cast<uint64_t>(chain);
    return Stream{static_cast<uint32_t>(gid), static_cast<uint32_t>(gid >> 32)};
}

// ---------------------------------------------------------------------------------------------
// Trace writers.  Row index = the reference's loop index i (row 0 = initial theta,
// GlobalMCMC.py:34-35); rows are relative to trace_row_base so a run can be chunked.
// ---------------------------------------------------------------------------------------------
template <int D>
struct VecOf { using type = float; };
template <> struct VecOf<2> { using type = float2; };
template <> struct VecOf<4> { using type = float4; };

template <int D>
__device__ __forceinline__ void store_row(float* dst, const float (&v)[D])
{
    if constexpr (D == 2) {
        *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
    } else if constexpr (D == 4) {
        *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
#pragma unroll
        for (int k = 0; k < D; ++k) dst[k] = v[k];
    }
}

// TIME_MAJOR: trace[row][chain][d] — a warp's 32 chains are contiguous, one vector store per row.
template <int D>
struct TimeMajorWriter {
    float* next;  // &trace[next row][chain][0]; rows are written consecutively
    bool active;
    __device__ __forceinline__ TimeMajorWriter(const RunParams& r, int32_t chain, bool act, float*)
        : active(act)
    {
        const int64_t first_row = static_cast<int64_t>(r.first_step) - (r.write_row0 ? 1 : 0) - r.trace_row_base;
        next = r.trace + (first_row * r.trace_chains + r.trace_chain_off + chain) * D;
    }
    __device__ __forceinline__ void put(const RunParams& r, uint32_t, const float (&v)[D])
    {
        if (active) store_row<D>(next, v);
        next += r.trace_chains * D;
    }
    __device__ __forceinline__ void maybe_flush(const RunParams&, uint32_t) {}
    __device__ __forceinline__ void finish(const RunParams&) {}
    static constexpr int smem_floats_per_warp = 0;
};

// CHAIN_MAJOR: trace[chain][row][d] — out[c] is a reference-shaped [num_ite, d] chain.  A thread
// per chain would store with a stride of a whole chain, so each warp stages up to 32 rows of its
// 32 chains in shared memory (component-planar, row pitch 33 words: conflict-free both ways) and
// flushes them as 32 runs of 32*D contiguous floats, 128 B per store instruction.
template <int D>
struct ChainMajorWriter {
    static constexpr int kPitch = 33;
    static constexpr int kPlane = 32 * kPitch + (D > 1 ? 32 / D : 0);  // plane skew spreads banks on read
    static constexpr int smem_floats_per_warp = D * kPlane;
    float* tile;       // this warp's tile
    float* out;        // &trace[chain0_of_warp][0][0]
    int64_t chain_stride;
    uint32_t row0;     // absolute row of tile slot 0
    int32_t count;     // buffered rows
    int32_t lane;
    int32_t chains_in_warp;  // valid chains of this warp (tail warp may have < 32)

    __device__ __forceinline__ ChainMajorWriter(const RunParams& r, int32_t, bool, float* smem_warp)
        : tile(smem_warp), chain_stride(r.trace_rows * D), row0(0), count(0), lane(threadIdx.x & 31)
    {
        // warp-uniform geometry (tail lanes shadow another chain's state but must agree on the tile)
        const int32_t chain0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
        out = r.trace + (r.trace_chain_off + chain0) * chain_stride;
        chains_in_warp = min(32, r.n_chains - chain0);
    }

    __device__ __forceinline__ void put(const RunParams& r, uint32_t row, const float (&v)[D])
    {
        if (count == 0) row0 = row;
#pragma unroll
        for (int k = 0; k < D; ++k) tile[k * kPlane + count * kPitch + lane] = v[k];
        ++count;
    }

    // call after put(row): flushes on a 32-row boundary of the absolute row index, so every run a
    // chain receives is 32 rows long and 128B-aligned (the tile never holds more than 32 rows)
    __device__ __forceinline__ void maybe_flush(const RunParams& r, uint32_t row)
    {
        if (((row + 1u) & 31u) == 0u) flush(r);
    }

    __device__ __forceinline__ void flush(const RunParams& r)
    {
        __syncwarp();
        const int32_t n = count * D;  // floats per chain in this tile
        const int64_t off = (static_cast<int64_t>(row0) - r.trace_row_base) * D;
        for (int32_t c = 0; c < chains_in_warp; ++c) {
            float* dst = out + c * chain_stride + off;
            for (int32_t e = lane; e < n; e += 32) {
                const int32_t s = e / D, k = e - s * D;
                dst[e] = tile[k * kPlane + s * kPitch + c];
            }
        }
        __syncwarp();
        count = 0;
    }

    __device__ __forceinline__ void finish(const RunParams& r)
    {
        if (count > 0) flush(r);
    }
};

// D in {1, 2, 4}: a row is one 4/8/16-byte vector, so the tile is [32 rows][33] vectors — a lane
// stores its chain's row with one STS (consecutive lanes, conflict-free), and a flush reads column c
// with one LDS per lane (lane = row; pitch 33 keeps the 64/128-bit phases conflict-free) and writes
// chain c's 32 rows as ONE coalesced store instruction of 32 vectors (128/256/512 contiguous bytes).
template <int D>
struct ChainMajorVecWriter {
    using V = typename VecOf<D>::type;
    static constexpr int kPitch = 33;
    static constexpr int smem_floats_per_warp = 32 * kPitch * D;
    V* tile;
    V* out;            // &trace[chain0_of_warp][0] in vector units
    int64_t chain_stride;  // vectors per chain
    uint32_t row0;
    int32_t count, lane, chains_in_warp;

    __device__ __forceinline__ ChainMajorVecWriter(const RunParams& r, int32_t, bool, float* smem_warp)
        : tile(reinterpret_cast<V*>(smem_warp)), chain_stride(r.trace_rows), row0(0), count(0), lane(threadIdx.x & 31)
    {
        const int32_t chain0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
        out = reinterpret_cast<V*>(r.trace) + (r.trace_chain_off + chain0) * chain_stride;
        chains_in_warp = min(32, r.n_chains - chain0);
    }

    __device__ __forceinline__ void put(const RunParams&, uint32_t row, const float (&v)[D])
    {
        if (count == 0) row0 = row;
        V x;
        if constexpr (D == 1) x = v[0];
        else if constexpr (D == 2) x = make_float2(v[0], v[1]);
        else x = make_float4(v[0], v[1], v[2], v[3]);
        tile[count * kPitch + lane] = x;
        ++count;
    }

    __device__ __forceinline__ void maybe_flush(const RunParams& r, uint32_t row)
    {
        if (((row + 1u) & 31u) == 0u) flush(r);
    }

    __device__ __forceinline__ void flush(const RunParams& r)
    {
        __syncwarp();
        V* dst = out + (static_cast<int64_t>(row0) - r.trace_row_base) + lane;
        const V* src = tile + lane * kPitch;
        if (count == 32 && chains_in_warp == 32) {
#pragma unroll 8
            for (int32_t c = 0; c < 32; ++c) dst[c * chain_stride] = src[c];
        } else {
            for (int32_t c = 0; c < chains_in_warp; ++c)
                if (lane < count) dst[c * chain_stride] = src[c];
        }
        __syncwarp();
        count = 0;
    }

    __device__ __forceinline__ void finish(const RunParams& r)
    {
        if (count > 0) flush(r);
    }
};

// EVENTS: the chain as its moves (glabc.h GLABC_TRACE_EVENTS).  One scattered store per move instead of one row per step.
template <int D>
struct EventWriter {
    float* ev;        // &events[chain][0][0], entries of 1 + D floats
    uint32_t n;       // entries used so far + 1 (entry 0 is the header)
    uint32_t cap;     // trace_rows
    float last[D];
    bool active;
    static constexpr int smem_floats_per_warp = 0;
    __device__ __forceinline__ EventWriter(const RunParams& r, int32_t chain, bool act, float*)
        : ev(r.trace + (r.trace_chain_off + chain) * r.trace_rows * (1 + D)), n(1u), cap(static_cast<uint32_t>(r.trace_rows)), active(act)
    {
#pragma unroll
        for (int k = 0; k < D; ++k) last[k] = __int_as_float(0x7fc00000);   // NaN: the first row always differs
    }
    __device__ __forceinline__ void put(const RunParams&, uint32_t row, const float (&v)[D])
    {
        bool moved = false;
#pragma unroll
        for (int k = 0; k < D; ++k) moved |= !(v[k] == last[k]);
        if (moved) {
            if (active && n < cap) {
                float* e = ev + static_cast<int64_t>(n) * (1 + D);
                e[0] = __uint_as_float(row);
#pragma unroll
                for (int k = 0; k < D; ++k) e[1 + k] = v[k];
            }
            ++n;
#pragma unroll
            for (int k = 0; k < D; ++k) last[k] = v[k];
        }
    }
    __device__ __forceinline__ void maybe_flush(const RunParams&, uint32_t) {}
    __device__ __forceinline__ void finish(const RunParams&)
    {
        if (active) ev[0] = __uint_as_float(n - 1u);
    }
};

struct NoTraceWriter {
    __device__ __forceinline__ NoTraceWriter(const RunParams&, int32_t, bool, float*) {}
    template <int D>
    __device__ __forceinline__ void put(const RunParams&, uint32_t, const float (&)[D]) {}
    __device__ __forceinline__ void maybe_flush(const RunParams&, uint32_t) {}
    __device__ __forceinline__ void finish(const RunParams&) {}
    static constexpr int smem_floats_per_warp = 0;
};

template <int D, int LAYOUT> struct WriterFor;
template <int D> struct WriterFor<D, GLABC_TRACE_NONE> { using type = NoTraceWriter; };
template <int D> struct WriterFor<D, GLABC_TRACE_TIME_MAJOR> { using type = TimeMajorWriter<D>; };
template <int D> struct WriterFor<D, GLABC_TRACE_CHAIN_MAJOR> { using type = ChainMajorVecWriter<D>; };
template <> struct WriterFor<3, GLABC_TRACE_CHAIN_MAJOR> { using type = ChainMajorWriter<3>; };
template <int D> struct WriterFor<D, GLABC_TRACE_EVENTS> { using type = EventWriter<D>; };

// ---------------------------------------------------------------------------------------------
// Per-chain statistics (layout: GLABC_STAT_* in glabc.h)
// ---------------------------------------------------------------------------------------------
template <int D>
struct ChainStats {
    static constexpr int kTri = D * (D + 1) / 2;
    uint32_t n_global, acc_local, acc_global;
    float f_global, f_acc, f_acc_global;  // FAST path: the same counts kept as floats on the FMA pipe
                                          // (exact below 2^24 transitions per launch — host-enforced)
    float sum[D], sumsq[D], gram[kTri];

    __device__ __forceinline__ ChainStats()
        : n_global(0), acc_local(0), acc_global(0), f_global(0.0f), f_acc(0.0f), f_acc_global(0.0f)
    {
#pragma unroll
        for (int i = 0; i < D; ++i) sum[i] = sumsq[i] = 0.0f;
#pragma unroll
        for (int i = 0; i < kTri; ++i) gram[i] = 0.0f;
    }

    // theta_new is the post-decision state; delta = theta_new - theta_prev (zero unless moved)
    __device__ __forceinline__ void update(bool is_global, bool moved, const float (&theta_new)[D],
                                           const float (&theta_prev)[D])
    {
        n_global += is_global;
        acc_global += (moved && is_global);
        acc_local += (moved && !is_global);
        float dl[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            sum[i] += theta_new[i];
            sumsq[i] = fmaf(theta_new[i], theta_new[i], sumsq[i]);
            dl[i] = theta_new[i] - theta_prev[i];
        }
        int t = 0;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = i; j < D; ++j, ++t) gram[t] = fmaf(dl[i], dl[j], gram[t]);
    }

    // glob, moved: 0/1 masks
    __device__ __forceinline__ void update_masked(float glob, float moved, const float (&theta_new)[D],
                                                  const float (&theta_prev)[D])
    {
        f_global += glob;
        f_acc += moved;
        f_acc_global = fmaf(moved, glob, f_acc_global);
        float dl[D];
#pragma unroll
        for (int i = 0; i < D; ++i) {
            sum[i] += theta_new[i];
            sumsq[i] = fmaf(theta_new[i], theta_new[i], sumsq[i]);
            dl[i] = theta_new[i] - theta_prev[i];
        }
        int t = 0;
#pragma unroll
        for (int i = 0; i < D; ++i)
#pragma unroll
            for (int j = i; j < D; ++j, ++t) gram[t] = fmaf(dl[i], dl[j], gram[t]);
    }

    __device__ __forceinline__ void store(float* st, uint32_t n_steps) const
    {
        st[GLABC_STAT_STEPS] += static_cast<float>(n_steps);
        st[GLABC_STAT_GLOBAL_STEPS] += static_cast<float>(n_global) + f_global;
        st[GLABC_STAT_ACC_LOCAL] += static_cast<float>(acc_local) + (f_acc - f_acc_global);
        st[GLABC_STAT_ACC_GLOBAL] += static_cast<float>(acc_global) + f_acc_global;
#pragma unroll
        for (int i = 0; i < D; ++i) {
            st[GLABC_STAT_SUM + i] += sum[i];
            st[GLABC_STAT_SUM + D + i] += sumsq[i];
        }
#pragma unroll
        for (int i = 0; i < kTri; ++i) st[GLABC_STAT_SUM + 2 * D + i] += gram[i];
    }
};

}  // namespace glabc
