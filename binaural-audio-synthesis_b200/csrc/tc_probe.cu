// FEASIBILITY PROBE, not the product path: the chunk / subchunk FIR of make_signal_move_2d
// (apply_hrtf.py:431-453) as a tensor-core contraction on tcgen05 with a 3 x TF32 split.
//
// For the boundary filter H_i (both ears) the renderer needs   sum_q  s_q * (x_q (*) H_i)   over the 32
// subchunks q of the two chunks next to boundary i, with s_q = alpha_q (chunk i-1) or 1 - alpha_q (chunk i)
// (apply_hrtf.py:442-443 distributes over the convolution).  Written as a matrix product per boundary:
//
//     D[r][n] = sum_m  T[r][m] * X[m][n],     T[r][m] = H_i[r - m]  (Toeplitz, r < K + 31, m < 32),
//                                              X[m][n] = s_n * x[32 (16 (i-1) + n) + m]        n < 32
//
// and output sample 512 (i-1) + 32 n + r receives D[r][n] (overlap-add).  M = 2 ears x (K + 31) rows,
// N = 32 columns, K_mma = 32: 60 tcgen05.mma (M128 N32 K8, kind::tf32) per boundary with the 3 x TF32 split
// (hi*hi + hi*lo + lo*hi; tf32 operands are fp32 words whose low 13 mantissa bits the tensor core ignores).
//
// The Toeplitz operand is never materialised.  With the columns reversed (m' = 31 - m) the element
// (r, m') is g[r + m'], g = the zero-padded taps: it depends on r + m' only.  In the K-major no-swizzle
// canonical layout a core matrix is 8 rows x 16 bytes, rows 16 bytes apart; an array W of 16-byte slots
// W[j] = {g[j], g[j+1], g[j+2], g[j+3]} makes slot (r + 4 kc) the k-chunk kc of row r, so the shared-memory
// descriptor  {start = &W[128 t + 8 s], stride-byte-offset = 128 (8 rows), leading-byte-offset = 64 (4 slots)}
// addresses the whole 128 x 8 operand of M-tile t, k-step s with OVERLAPPING core matrices: 10.7 KB of
// shared memory per operand instead of a 576 x 32 x 4 B = 74 KB Toeplitz block.
//
// mode 0: correct end to end (outputs accumulated with global atomics) - for the accuracy figure;
// mode 1: the same MMAs and TMEM loads, overlap-add in registers, no global accumulation - rate proxy;
// mode 2: MMAs only.
#include "bas_internal.cuh"
#include "../../include/bas_probe.h"

namespace {

typedef unsigned long long u64;

__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ u64 smem_desc(unsigned addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    // start address, leading / stride byte offsets in 16-byte units; descriptor version 1 (Blackwell); no swizzle
    return (u64)((addr >> 4) & 0x3FFF) | ((u64)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((u64)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

__device__ __forceinline__ void mma_tf32(unsigned d_tmem, u64 a_desc, u64 b_desc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
}

__device__ __forceinline__ void tmem_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

constexpr int kRowsPerEar = 320;              // ear R's rows start here: ear L's last k-chunk reads slots below 320
constexpr int kMTiles = 5;                    // 640 rows
constexpr int kGLen = 704;                    // g[u], u < 640 + 32 + padding
constexpr int kWSlots = 672;
constexpr int kTmemCols = 256;                // 5 tiles x 32 columns -> next power of two

struct TcSmem {
    float4 w[2][kWSlots];                     // hi / lo, 16-byte slots W[j] = g[j .. j + 3]
    float4 b[2][256];                         // hi / lo, unit (kchunk, n) at kchunk * 32 + n: 4 reversed, scaled samples of column n
    float g[2][kGLen];
    unsigned long long bar;
    unsigned tmem_base;
};

__global__ void __launch_bounds__(128, 1)
bas_probe_tc_kernel(const float* __restrict__ x, long long n_in, const float2* __restrict__ filt, int pitch, int K, int C,
                    float* __restrict__ out, long long out_stride, long long n_out, int mode, float* __restrict__ sink) {
    extern __shared__ __align__(128) unsigned char raw[];
    TcSmem& sm = *reinterpret_cast<TcSmem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_chunks = (int)(n_in / C);
    const int spc = C / 32;                                        // subchunks per chunk (16)

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&sm.bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&sm.tmem_base)), "n"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sm.tmem_base;
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 32, M = 128
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
    unsigned phase = 0;
    float keep = 0.f;

    for (int i = blockIdx.x; i <= n_chunks; i += gridDim.x) {
        // mode 4: operands staged once, then only the 60 MMAs per boundary - the tensor-core rate of this shape on its own
        const bool stage = mode != 4 || i == (int)blockIdx.x;
        // ---- g: the zero-padded taps of boundary filter i, ear L at u = 31 .. 31 + K - 1, ear R 320 rows later
        const float2* row = filt + (long long)i * pitch;
        if (stage) {
        for (int u = tid; u < kGLen; u += 128) {
            float v = 0.f;
            const int tl = u - 31, tr = u - 31 - kRowsPerEar;
            if (tl >= 0 && tl < K) v = row[tl].x;
            else if (tr >= 0 && tr < K) v = row[tr].y;
            const float hi = tf32_hi(v);
            sm.g[0][u] = hi;
            sm.g[1][u] = v - hi;
        }
        // ---- X: column n = subchunk n of the two chunks around the boundary, scaled by its blend weight
        for (int u = tid; u < 256; u += 128) {
            const int kchunk = u >> 5, n = u & 31;
            const int chunk = i - 1 + (n >= spc ? 1 : 0);
            const int q = n >= spc ? n - spc : n;
            const float alpha = (float)(q * 32) / (float)C;                    // apply_hrtf.py:442
            const float s = n >= spc ? 1.f - alpha : alpha;                    // weight of H_i in the blend of :443
            float hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const long long idx = (long long)chunk * C + 32 * q + 31 - (4 * kchunk + e);
                const float v = (chunk >= 0 && chunk < n_chunks) ? s * x[idx] : 0.f;
                hi[e] = tf32_hi(v);
                lo[e] = v - hi[e];
            }
            sm.b[0][u] = make_float4(hi[0], hi[1], hi[2], hi[3]);
            sm.b[1][u] = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        __syncthreads();
        for (int j = tid; j < kWSlots; j += 128) {
            sm.w[0][j] = make_float4(sm.g[0][j], sm.g[0][j + 1], sm.g[0][j + 2], sm.g[0][j + 3]);
            sm.w[1][j] = make_float4(sm.g[1][j], sm.g[1][j + 1], sm.g[1][j + 2], sm.g[1][j + 3]);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // generic-proxy writes -> visible to the tensor core
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        if (tid == 0) {
            const unsigned wa[2] = {s32(&sm.w[0][0]), s32(&sm.w[1][0])};
            const unsigned ba[2] = {s32(&sm.b[0][0]), s32(&sm.b[1][0])};
#pragma unroll 1
            for (int t = 0; t < kMTiles; ++t) {
                unsigned acc = 0;
#pragma unroll
                for (int split = 0; split < 3; ++split) {                      // hi*hi, hi*lo, lo*hi
                    const int sa = split == 2 ? 1 : 0, sb = split == 1 ? 1 : 0;
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const u64 da = smem_desc(wa[sa] + 16u * (128u * t + 8u * s), 64u, 128u);
                        const u64 db = smem_desc(ba[sb] + 16u * (64u * s), 512u, 128u);
                        mma_tf32(tmem + 32u * t, da, db, idesc, acc);
                        acc = 1;
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&sm.bar)) : "memory");
        }
        // ---- wait for the MMAs of this boundary
        {
            unsigned ok = 0;
            unsigned long long t0 = 0;
            for (unsigned spins = 0; !ok; ++spins) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                             : "=r"(ok) : "r"(s32(&sm.bar)), "r"(phase) : "memory");
                if (!ok && (spins & 1023u) == 1023u) {
                    unsigned long long now;
                    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
                    if (!t0) t0 = now;
                    else if (now - t0 > 5000000000ull) __trap();              // 5 s: fail loudly instead of hanging
                }
            }
            phase ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue: row r = 128 t + 32 warp + lane of every M-tile, 32 columns each
        if (mode != 2 && mode != 4) {
            float ola[16 + 32];                                              // mode 1: overlap-add in registers, block index n + 4 t (+ warp)
#pragma unroll
            for (int k = 0; k < 48; ++k) ola[k] = 0.f;
#pragma unroll
            for (int t = 0; t < kMTiles; ++t) {
                float v[32];
                tmem_ld32(tmem + ((unsigned)(warp * 32) << 16) + 32u * t, v);
                const int r = 128 * t + 32 * warp + lane;
                const int ear = r >= kRowsPerEar ? 1 : 0;
                const int tt = r - ear * kRowsPerEar;                        // output offset inside the subchunk's convolution
                if (mode == 0) {
                    if (tt < K + 31) {
#pragma unroll
                        for (int n = 0; n < 32; ++n) {
                            const long long p = (long long)(i - 1) * C + 32 * n + tt;
                            if (p >= 0 && p < n_out) atomicAdd(out + (long long)ear * out_stride + p, v[n]);
                        }
                    }
                } else {
#pragma unroll
                    for (int n = 0; n < 32; ++n) ola[n + 4 * (t % 3)] += v[n];   // same add count as a real overlap-add
                }
            }
            if (mode == 1) {
#pragma unroll
                for (int k = 0; k < 48; ++k) keep += ola[k];
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();                                                     // shared memory and TMEM are free for the next boundary
    }
    if (mode != 0 && sink) sink[(size_t)blockIdx.x * blockDim.x + tid] = keep;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(kTmemCols) : "memory");
}


// ------------------------------------------------------------------------------------------------------------
// mode 3: the same contraction PIPELINED and complete (no atomics): one CTA per SM owns a contiguous range of 512-sample
// output chunks and walks over the boundaries that feed it.  Nine warps, three roles, double-buffered operands and
// accumulators:   warps 5-8 stage the operands of boundary k+1 | warp 4 issues the 60 MMAs of boundary k |
//                 warps 0-3 load the accumulators of boundary k-1 from TMEM, overlap-add them in registers, add them to a
//                 2048-sample output ring in shared memory and flush the chunk that is now complete to global memory.
// ------------------------------------------------------------------------------------------------------------
struct PipeSmem {
    float4 w[2][2][kWSlots];                  // [buffer][hi / lo]
    float4 b[2][2][256];
    float g[2][kGLen];                        // stagers' scratch: padded taps, hi / lo
    float ring[2][2048];                      // output ring per ear: samples [512 (i-1), 512 (i-1) + 2048) of the boundary in flight
    unsigned long long op_full[2], op_empty[2], tm_full[2], tm_empty[2];
    unsigned tmem_base;
};

__device__ __forceinline__ void bar_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok = 0;
    unsigned long long t0 = 0;
    for (unsigned spins = 0; !ok; ++spins) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(s32(bar)), "r"(parity) : "memory");
        if (!ok && (spins & 1023u) == 1023u) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(now));
            if (!t0) t0 = now;
            else if (now - t0 > 5000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void bar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}

__global__ void __launch_bounds__(288, 1)
bas_probe_tc_pipe_kernel(const float* __restrict__ x, long long n_in, const float2* __restrict__ filt, int pitch, int K, int C,
                         float* __restrict__ out, long long out_stride, long long n_out) {
    extern __shared__ __align__(128) unsigned char raw[];
    PipeSmem& sm = *reinterpret_cast<PipeSmem*>(raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_chunks = (int)(n_in / C);
    const int spc = C / 32;
    const int out_chunks = (int)((n_out + C - 1) / C);
    // this CTA's output chunks [c0, c1) and the boundaries that feed them
    const int c0 = (int)((long long)blockIdx.x * out_chunks / gridDim.x), c1 = (int)((long long)(blockIdx.x + 1) * out_chunks / gridDim.x);
    const int i_first = c0 - 1 < 0 ? 0 : c0 - 1, i_last = c1 > n_chunks ? n_chunks : c1;
    const int nb = c1 > c0 ? i_last - i_first + 1 : 0;

    if (tid == 0) {
        for (int b = 0; b < 2; ++b) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(s32(&sm.op_full[b])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&sm.op_empty[b])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&sm.tm_full[b])) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(s32(&sm.tm_empty[b])) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int k = tid; k < 2 * 2048; k += 288) (&sm.ring[0][0])[k] = 0.f;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&sm.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = sm.tmem_base;
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);

    if (warp >= 5) {
        // ================================ stagers ================================
        const int st = tid - 160;
        for (int k = 0; k < nb; ++k) {
            const int i = i_first + k, b = k & 1, u = k >> 1;
            const float2* row = filt + (long long)i * pitch;
            asm volatile("bar.sync 2, 128;" ::: "memory");                 // the previous boundary's expansion has read g
            for (int e = st; e < kGLen; e += 128) {
                float v = 0.f;
                const int tl = e - 31, tr = e - 31 - kRowsPerEar;
                if (tl >= 0 && tl < K) v = row[tl].x;
                else if (tr >= 0 && tr < K) v = row[tr].y;
                const float hi = tf32_hi(v);
                sm.g[0][e] = hi;
                sm.g[1][e] = v - hi;
            }
            if (u >= 1) bar_wait(&sm.op_empty[b], (unsigned)((u - 1) & 1));   // the MMAs that read this buffer have completed
            for (int e = st; e < 256; e += 128) {
                const int kchunk = e >> 5, n = e & 31;
                const int chunk = i - 1 + (n >= spc ? 1 : 0);
                const int q = n >= spc ? n - spc : n;
                const float alpha = (float)(q * 32) / (float)C;
                const float s = n >= spc ? 1.f - alpha : alpha;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (chunk >= 0 && chunk < n_chunks) v = *reinterpret_cast<const float4*>(x + (long long)chunk * C + 32 * q + 28 - 4 * kchunk);
                // reversed within the row: element e' of the unit is sample 31 - (4 kchunk + e')
                const float r0 = s * v.w, r1 = s * v.z, r2 = s * v.y, r3 = s * v.x;
                const float h0 = tf32_hi(r0), h1 = tf32_hi(r1), h2 = tf32_hi(r2), h3 = tf32_hi(r3);
                sm.b[b][0][e] = make_float4(h0, h1, h2, h3);
                sm.b[b][1][e] = make_float4(r0 - h0, r1 - h1, r2 - h2, r3 - h3);
            }
            asm volatile("bar.sync 2, 128;" ::: "memory");                 // g complete
            for (int j = st; j < kWSlots; j += 128) {
                sm.w[b][0][j] = make_float4(sm.g[0][j], sm.g[0][j + 1], sm.g[0][j + 2], sm.g[0][j + 3]);
                sm.w[b][1][j] = make_float4(sm.g[1][j], sm.g[1][j + 1], sm.g[1][j + 2], sm.g[1][j + 3]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bar_arrive(&sm.op_full[b]);
        }
    } else if (warp == 4) {
        // ================================ MMA issuer ================================
        for (int k = 0; k < nb; ++k) {
            const int b = k & 1, u = k >> 1;
            bar_wait(&sm.op_full[b], (unsigned)(u & 1));
            if (u >= 1) bar_wait(&sm.tm_empty[b], (unsigned)((u - 1) & 1));   // the epilogue has drained this accumulator buffer
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const unsigned wa[2] = {s32(&sm.w[b][0][0]), s32(&sm.w[b][1][0])};
                const unsigned ba[2] = {s32(&sm.b[b][0][0]), s32(&sm.b[b][1][0])};
#pragma unroll 1
                for (int t = 0; t < kMTiles; ++t) {
                    unsigned acc = 0;
#pragma unroll
                    for (int split = 0; split < 3; ++split) {
                        const int sa = split == 2 ? 1 : 0, sb = split == 1 ? 1 : 0;
#pragma unroll
                        for (int s = 0; s < 4; ++s) {
                            mma_tf32(tmem + 256u * b + 32u * t, smem_desc(wa[sa] + 16u * (128u * t + 8u * s), 64u, 128u),
                                     smem_desc(ba[sb] + 16u * (64u * s), 512u, 128u), idesc, acc);
                            acc = 1;
                        }
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&sm.op_empty[b])) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&sm.tm_full[b])) : "memory");
            }
            __syncwarp();
        }
    } else {
        // ================================ epilogue ================================
        for (int k = 0; k < nb; ++k) {
            const int i = i_first + k, b = k & 1, u = k >> 1;
            bar_wait(&sm.tm_full[b], (unsigned)(u & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float olaL[40], olaR[40];
#pragma unroll
            for (int e = 0; e < 40; ++e) { olaL[e] = 0.f; olaR[e] = 0.f; }
#pragma unroll
            for (int t = 0; t < kMTiles; ++t) {
                float v[32];
                tmem_ld32(tmem + ((unsigned)(warp * 32) << 16) + 256u * b + 32u * t, v);
                // rows 128 t + 32 warp + lane: ear L below row 320 (block n + 4 t + warp), ear R from row 320 (block n + 4 t - 10 + warp)
                if (t < 2 || (t == 2 && warp < 2)) {
#pragma unroll
                    for (int n = 0; n < 32; ++n) olaL[n + 4 * (t < 2 ? t : 2)] += v[n];
                } else {
#pragma unroll
                    for (int n = 0; n < 32; ++n) olaR[n + 4 * (t - 2)] += v[n];
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(&sm.tm_empty[b]);                                   // the accumulators are in registers: the buffer may be overwritten
            const long long base = (long long)(i - 1) * C;                 // first output sample of this boundary (may be -512)
            const unsigned rb = (unsigned)(((i - 1) & 3) * C);             // its place in the 2048-sample ring
#pragma unroll
            for (int e = 0; e < 40; ++e) {
                atomicAdd(&sm.ring[0][(rb + 32u * (unsigned)(e + warp) + lane) & 2047u], olaL[e]);
                atomicAdd(&sm.ring[1][(rb + 32u * (unsigned)(e + warp - 2 + 2) + lane - 64u) & 2047u], olaR[e]);
            }
            asm volatile("bar.sync 3, 128;" ::: "memory");
            // chunk i - 1 is complete: no later boundary adds to it
            const int chunk = i - 1;
            {
                const int ear = tid >> 6, o = (tid & 63) * 8;               // 128 threads x 8 samples = 2 ears x 512
                float* r = &sm.ring[ear][(rb + o) & 2047u];
                const float4 a = *reinterpret_cast<float4*>(r), c = *reinterpret_cast<float4*>(r + 4);
                *reinterpret_cast<float4*>(r) = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(r + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
                if (chunk >= c0 && chunk < c1) {
                    const long long p = base + o;
                    float* dst = out + (long long)ear * out_stride + p;
                    if (p + 8 <= n_out) {
                        *reinterpret_cast<float4*>(dst) = a;
                        *reinterpret_cast<float4*>(dst + 4) = c;
                    } else {
                        const float vals[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
                        for (int e = 0; e < 8; ++e) if (p + e < n_out) dst[e] = vals[e];
                    }
                }
            }
            asm volatile("bar.sync 3, 128;" ::: "memory");
        }
        // the last CTA also owns the K - 1 sample tail after the last boundary: chunk i_last is complete, nothing follows
        for (int chunk = i_last; chunk < c1 && nb > 0; ++chunk) {
            const int ear = tid >> 6, o = (tid & 63) * 8;
            const float* r = &sm.ring[ear][(unsigned)((chunk & 3) * C + o) & 2047u];
            const long long p = (long long)chunk * C + o;
            float* dst = out + (long long)ear * out_stride + p;
            for (int e = 0; e < 8; ++e) if (p + e < n_out) dst[e] = r[e];
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

}  // namespace

extern "C" int bas_probe_tc_render(const float* x_dev, long long n_in, const float* filt_dev, int K, int C, float* out_dev,
                                   long long out_stride, long long n_out, int mode, int blocks, float* sink_dev, void* stream) {
    BAS_CHECK_ARG(x_dev && filt_dev && out_dev, "null pointer");
    BAS_CHECK_ARG(K >= 1 && K + 31 <= kRowsPerEar - 31 && C == 512 && n_in % C == 0, "probe geometry: K <= 258, chunksize 512");
    BAS_CHECK_ARG(mode >= 0 && mode <= 4 && blocks >= 1, "mode / blocks");
    if (mode == 3) {
        BAS_CHECK_ARG(out_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(out_dev) & 15) == 0 && (reinterpret_cast<uintptr_t>(x_dev) & 15) == 0, "alignment");
        const size_t smem3 = sizeof(PipeSmem) > 120 * 1024 ? sizeof(PipeSmem) : 120 * 1024;   // one CTA per SM (all 512 TMEM columns)
        BAS_CUDA(cudaFuncSetAttribute(bas_probe_tc_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        bas_probe_tc_pipe_kernel<<<blocks, 288, smem3, (cudaStream_t)stream>>>(x_dev, n_in, reinterpret_cast<const float2*>(filt_dev),
                                                                               (K + 31) / 32 * 32 + 2, K, C, out_dev, out_stride, n_out);
        BAS_LAUNCH_CHECK();
        return 0;
    }
    const int pitch = (K + 31) / 32 * 32 + 2;
    const size_t smem = sizeof(TcSmem) > 120 * 1024 ? sizeof(TcSmem) : 120 * 1024;      // one CTA per SM: TMEM is allocated per CTA
    BAS_CUDA(cudaFuncSetAttribute(bas_probe_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bas_probe_tc_kernel<<<blocks, 128, smem, (cudaStream_t)stream>>>(x_dev, n_in, reinterpret_cast<const float2*>(filt_dev), pitch, K, C,
                                                                      out_dev, out_stride, n_out, mode, sink_dev);
    BAS_LAUNCH_CHECK();
    return 0;
}
