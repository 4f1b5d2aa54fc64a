// FP32 pipe probe.  SURVEY.md 8(d): the render path is bound by the FP32 FMA pipe, not by HBM, so
// bench.py states the achieved FMA rate next to a MEASURED peak from this kernel (same clocks, same
// box) instead of a nominal 148 SM x 128 lanes x clock figure.
#include "bas_internal.cuh"
#include "../../include/bas_probe.h"

namespace {

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// mhz (optional): SM clock the kernel actually ran at, from clock64 against the nanosecond timer
template <bool PACKED>
__global__ void __launch_bounds__(256)
bas_probe_fma_kernel(int iters, float* __restrict__ sink, float* __restrict__ mhz = nullptr) {
    const float seed = (float)(threadIdx.x & 7) * 1e-3f;
    long long c0 = 0; unsigned long long t0 = 0;
    if (mhz && blockIdx.x == 0 && threadIdx.x == 0) { t0 = global_ns(); c0 = clock64(); }
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
        if (mhz && blockIdx.x == 0 && threadIdx.x == 0) { const long long c1 = clock64(); const unsigned long long t1 = global_ns(); *mhz = 1e3f * (float)(c1 - c0) / (float)(t1 - t0); }
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
        if (mhz && blockIdx.x == 0 && threadIdx.x == 0) { const long long c1 = clock64(); const unsigned long long t1 = global_ns(); *mhz = 1e3f * (float)(c1 - c0) / (float)(t1 - t0); }
    }
}

// FIR-shaped register streams: 32 (packed) or 64 (scalar) accumulators, a 32-entry ring of taps and a
// fresh x scalar every 32 FMAs - the operand pattern of the render kernel without its loads.
// mode 2: fma.rn.f32x2 acc[r] += {x,x} * w[(r-m)&31];   mode 3: the same arithmetic as scalar fma.rn.f32.
template <int MODE>
__global__ void __launch_bounds__(128)
bas_probe_fir_kernel(int iters, float* __restrict__ sink) {
    const float seed = (float)(threadIdx.x & 7) * 1e-3f;
    float s = 0.f;
    if (MODE == 2) {
        unsigned long long acc[32], w[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(seed + i), "f"(seed - i));
            asm("mov.b64 %0, {%1, %2};" : "=l"(w[i]) : "f"(1e-3f * i + seed), "f"(2e-3f * i));
        }
        float x = 0.5f + seed;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                unsigned long long xx;
                asm volatile("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x));
#pragma unroll
                for (int r = 0; r < 32; ++r)
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[r]) : "l"(xx), "l"(w[(r - m) & 31]));
                x = x * 0.999f + 1e-4f;
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
    } else {
        float accl[32], accr[32], wl[32], wr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { accl[i] = seed + i; accr[i] = seed - i; wl[i] = 1e-3f * i + seed; wr[i] = 2e-3f * i; }
        float x = 0.5f + seed;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int m = 0; m < 32; ++m) {
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(accl[r]) : "f"(x), "f"(wl[(r - m) & 31]));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(accr[r]) : "f"(x), "f"(wr[(r - m) & 31]));
                }
                x = x * 0.999f + 1e-4f;
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) s += accl[i] + accr[i];
    }
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 4: the same 32x32 block of packed FMAs visited diagonal by diagonal: one tap pair w is reused by
// 32 consecutive FFMA2 while the x scalar and the accumulator change (x[32] lives in registers).
// mode 5: as mode 4 with scalar fma.rn.f32.
template <int MODE>
__global__ void __launch_bounds__(128)
bas_probe_diag_kernel(int iters, float* __restrict__ sink) {
    const float seed = (float)(threadIdx.x & 7) * 1e-3f;
    float s = 0.f;
    float x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) x[i] = 0.5f + seed + 1e-3f * i;
    if (MODE == 4) {
        unsigned long long acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(seed + i), "f"(seed - i));
        float wl = 1e-3f + seed, wr = 2e-3f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                unsigned long long w;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(w) : "f"(wl), "f"(wr));
#pragma unroll
                for (int m = 0; m < 32; ++m) {
                    unsigned long long xx;
                    asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(x[m]));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc[(m + j) & 31]) : "l"(xx), "l"(w));
                }
                wl = wl * 0.999f + 1e-4f; wr = wr * 0.998f + 1e-4f;
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
    } else {
        float accl[32], accr[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) { accl[i] = seed + i; accr[i] = seed - i; }
        float wl = 1e-3f + seed, wr = 2e-3f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
#pragma unroll
                for (int m = 0; m < 32; ++m) {
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(accl[(m + j) & 31]) : "f"(x[m]), "f"(wl));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(accr[(m + j) & 31]) : "f"(x[m]), "f"(wr));
                }
                wl = wl * 0.999f + 1e-4f; wr = wr * 0.998f + 1e-4f;
            }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) s += accl[i] + accr[i];
    }
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace

extern "C" int bas_probe_fma(int packed, int blocks, int threads, int iters, float* sink_dev, void* stream) {
    BAS_CHECK_ARG(sink_dev, "null pointer");
    BAS_CHECK_ARG(blocks >= 1 && threads >= 32 && threads <= 256 && threads % 32 == 0 && iters >= 1, "launch shape");
    if (packed == 4 || packed == 5) {
        BAS_CHECK_ARG(threads <= 128, "FIR probes use at most 128 threads per block");
        if (packed == 4) bas_probe_diag_kernel<4><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev);
        else bas_probe_diag_kernel<5><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev);
        BAS_LAUNCH_CHECK();
        return 0;
    }
    if (packed == 2 || packed == 3) {       // FIR-shaped streams: FMA count = blocks * threads * iters * 2048
        BAS_CHECK_ARG(threads <= 128, "FIR probes use at most 128 threads per block");
        if (packed == 2) bas_probe_fir_kernel<2><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev);
        else bas_probe_fir_kernel<3><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev);
        BAS_LAUNCH_CHECK();
        return 0;
    }
    if (packed) bas_probe_fma_kernel<true><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev);
    else bas_probe_fma_kernel<false><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev);
    BAS_LAUNCH_CHECK();
    return 0;
}

// The same FMA streams as bas_probe_fma (packed = 0 / 1), also reporting the SM clock they ran at.
extern "C" int bas_probe_clock(int packed, int blocks, int threads, int iters, float* sink_dev, float* mhz_dev, void* stream) {
    BAS_CHECK_ARG(sink_dev && mhz_dev, "null pointer");
    BAS_CHECK_ARG((packed == 0 || packed == 1) && blocks >= 1 && threads >= 32 && threads <= 256 && threads % 32 == 0 && iters >= 1, "launch shape");
    if (packed) bas_probe_fma_kernel<true><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev, mhz_dev);
    else bas_probe_fma_kernel<false><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, sink_dev, mhz_dev);
    BAS_LAUNCH_CHECK();
    return 0;
}

// ---- the render kernel's own 32x32 block (render_tiled.cuh: block_diag) on synthetic shared-memory data,
//      without the tile machinery around it: how fast the block itself can go --------------------------
#include "render_tiled.cuh"

namespace {
using namespace bas_render_detail;

// per warp: 34 input rows on the 144-byte pitch, and two filter rows of 2 x 258 taps
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
bas_probe_block_kernel(int iters, int blocks_per_iter, float* __restrict__ sink) {
    extern __shared__ __align__(16) unsigned char psm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pitch = 258;
    float2* rows = reinterpret_cast<float2*>(psm);                       // [4 rows][pitch], shared by the CTA
    float* xw = reinterpret_cast<float*>(psm + 4 * pitch * 8) + warp * 72 * kXPitch;   // 72 rows per warp
    for (int i = threadIdx.x; i < 4 * pitch; i += blockDim.x) rows[i] = make_float2(1e-3f * (i % 97), -1e-3f * (i % 89));
    for (int i = lane; i < 72 * kXPitch; i += 32) xw[i] = 1e-2f * (i % 31);
    __syncthreads();
    u64 acc[kBlk];
#pragma unroll
    for (int r = 0; r < kBlk; ++r) acc[r] = 0ull;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int d = 0; d < blocks_per_iter; ++d) {
            const int chunk = (lane + d) & 1;                             // lanes read two different filter rows
            const float2* ra = rows + chunk * pitch + 32 * d + 32;        // taps base - 31 .. base + 31 inside the row
            const float alpha = 0.0625f * ((lane + it) & 15);
            const u64 aa[1] = {pack2(alpha, alpha)};
            const float* xa = xw + (lane + 8 - d) * kXPitch;
            block_diag<1>(acc, ra, aa, ra, aa, pitch, xa, 0, xa, 0);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < kBlk; ++r) { float l, rr; unpack2(acc[r], l, rr); s += l + rr; }
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

// blocks CTAs of 128 threads, each warp running iters x 6 blocks (6 x 1024 useful packed FMAs per lane per iteration).
extern "C" int bas_probe_block(int ctas_per_sm, int blocks, int iters, float* sink_dev, void* stream) {
    BAS_CHECK_ARG(sink_dev && blocks >= 1 && iters >= 1 && ctas_per_sm >= 1 && ctas_per_sm <= 3, "launch shape");
    const size_t smem = 4 * 258 * 8 + 4 * 72 * kXPitch * 4;
    if (ctas_per_sm == 3) {
        BAS_CUDA(cudaFuncSetAttribute(bas_probe_block_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bas_probe_block_kernel<3><<<blocks, 128, smem, (cudaStream_t)stream>>>(iters, 6, sink_dev);
    } else {
        BAS_CUDA(cudaFuncSetAttribute(bas_probe_block_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        bas_probe_block_kernel<2><<<blocks, 128, smem, (cudaStream_t)stream>>>(iters, 6, sink_dev);
    }
    BAS_LAUNCH_CHECK();
    return 0;
}

// ---- time stamps in stream order (tools/peer_cost_probe.py: where do the kernels of a step run?) ----
namespace {
__global__ void bas_probe_stamp_kernel(unsigned long long* slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    *slot = t;
}
}  // namespace
extern "C" int bas_probe_stamp(unsigned long long* slot_dev, void* stream) {
    bas_probe_stamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(slot_dev);
    return (int)cudaGetLastError();
}
