// FP32 pipe probe.  SURVEY.md 8(d): the render path is bound by the FP32 FMA pipe, not by HBM, so
// bench.py states the achieved FMA rate next to a MEASURED peak from this kernel (same clocks, same
// box) instead of a nominal 148 SM x 128 lanes x clock figure.
#include "bas_internal.cuh"

namespace {

template <bool PACKED>
__global__ void __launch_bounds__(256)
bas_probe_fma_kernel(int iters, float* __restrict__ sink) {
    const float seed = (float)(threadIdx.x & 7) * 1e-3f;
    if (PACKED) {
        unsigned long long acc[16], a, b;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(0.999f + seed), "f"(0.998f));
        asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(1e-3f), "f"(2e-3f));
#pragma unroll
        for (int i = 0; i < 16; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(seed + i), "f"(seed - i));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(a), "l"(b));
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i]));
            s += lo + hi;
        }
        sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        float acc[32];
        const float a = 0.999f + seed, b = 1e-3f;
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = seed + i;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 32; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(acc[i]) : "f"(a), "f"(b));
        }
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) s += acc[i];
        sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

}  // namespace

extern "C" int bas_probe_fma(int packed, int blocks, int threads, int iters, float* sink_dev, void* stream) {
    BAS_CHECK_ARG(sink_dev, "null pointer");
    BAS_CHECK_ARG(blocks >= 1 && threads >= 32 && threads <= 256 && threads % 32 == 0 && iters >= 1, "launch shape");
    if (packed) bas_probe_fma_kernel<true><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev);
    else bas_probe_fma_kernel<false><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev);
    BAS_LAUNCH_CHECK();
    return 0;
}
