// How fast can the host cores write a dense float32 trace (run-length expansion of a piecewise-constant chain)?
// nvcc -O3 -o host_fill host_fill.cu -lpthread ; ./host_fill [threads]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include <immintrin.h>
// fill n float2 rows at p with v using 16-byte non-temporal stores (no read-for-ownership)
__attribute__((target("avx512f"))) static inline void fill_nt512(float2* p, size_t n, float2 v)
{
    size_t k = 0;
    while (k < n && (reinterpret_cast<uintptr_t>(p + k) & 63u)) p[k++] = v;       // head up to a cache line
    const __m512 vv = _mm512_castpd_ps(_mm512_set1_pd(*reinterpret_cast<const double*>(&v)));
    for (; k + 8 <= n; k += 8) _mm512_stream_ps(reinterpret_cast<float*>(p + k), vv);
    for (; k < n; ++k) p[k] = v;
}
__attribute__((target("avx"))) static inline void fill_nt256(float2* p, size_t n, float2 v)
{
    size_t k = 0;
    while (k < n && (reinterpret_cast<uintptr_t>(p + k) & 31u)) p[k++] = v;
    const __m256 vv = _mm256_castpd_ps(_mm256_set1_pd(*reinterpret_cast<const double*>(&v)));
    for (; k + 4 <= n; k += 4) _mm256_stream_ps(reinterpret_cast<float*>(p + k), vv);
    for (; k < n; ++k) p[k] = v;
}
static inline void fill_nt(float2* p, size_t n, float2 v)
{
    size_t k = 0;
    if ((reinterpret_cast<uintptr_t>(p) & 15u) && n) { p[0] = v; k = 1; }
    const __m128 vv = _mm_set_ps(v.y, v.x, v.y, v.x);
    for (; k + 2 <= n; k += 2) _mm_stream_ps(reinterpret_cast<float*>(p + k), vv);
    if (k < n) p[k] = v;
}
int main(int argc, char** argv)
{
    const size_t C = 65536, T = 10000;
    const size_t bytes = C * T * 8;
    float2* buf = nullptr;
    if (cudaHostAlloc(&buf, bytes, cudaHostAllocDefault) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    const int hw = (int)std::thread::hardware_concurrency();
    for (int nt_mode = 0; nt_mode < (__builtin_cpu_supports("avx512f") ? 4 : 3); ++nt_mode)
    for (int nt : {hw}) {
        if (nt < 1) continue;
        for (int rep = 0; rep < 2; ++rep) {
            auto t0 = std::chrono::steady_clock::now();
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t)
                th.emplace_back([=]() {
                    _mm_sfence();
                    for (size_t c = t; c < C; c += nt) {          // chain c: ~120 runs of ~83 rows
                        float2* row = buf + c * T;
                        size_t i = 0;
                        uint32_t s = (uint32_t)c * 2654435761u + 12345u;
                        while (i < T) {
                            s = s * 1664525u + 1013904223u;
                            size_t len = 20 + (s >> 25);          // 20..147
                            if (i + len > T) len = T - i;
                            const float2 v = make_float2((float)i, (float)c);
                            if (nt_mode == 3) fill_nt512(row + i, len, v); else if (nt_mode == 2) fill_nt256(row + i, len, v); else if (nt_mode == 1) fill_nt(row + i, len, v); else for (size_t k = 0; k < len; ++k) row[i + k] = v;
                            i += len;
                        }
                    }
                });
            for (auto& x : th) x.join();
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("nt_stores %d threads %3d rep %d: %.1f ms  %.1f GB/s\n", nt_mode, nt, rep, dt * 1e3, bytes / dt / 1e9);
        }
    }
    printf("hardware_concurrency %d\n", hw);
    return 0;
}
