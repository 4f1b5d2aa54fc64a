// K3/K4  render: the chunk / subchunk loops of make_signal_move_2d (apply_hrtf.py:431-453), the
// float32 cast (:459) and the peak that the normalisation needs (:462-464).
//
// The reference blends one filter per subchunk (h_q = (1-alpha_q) H_i + alpha_q H_{i+1}, :442-443),
// convolves the S input samples of that subchunk with it (:445-446) and overlap-adds (:450-453).
// Written per output sample that is
//
//     out_e[p] = sum_{k=0}^{K-1} x[p-k] * h_{q(p-k),e}[k],     q(n) = floor(n / S),
//
// i.e. the filter belongs to the INPUT sample.  Both kernels here are output-stationary: no atomics,
// no overlap-add traffic, bit-reproducible.
//
// bas_render_generic_kernel   any (C, S, K): one thread per output sample.  Reference-shaped, slow.
// bas_render_tiled_kernel     S == 32: register-tiled FP32 kernel on the packed-FMA pipe.
//
// Tiled kernel, per warp: lane l owns the 32 consecutive outputs of block b = b0 + l (both ears: 32
// fma.rn.f32x2 accumulators {L,R}).  Output block b draws on input subchunks q = b - d, d = 0..D:
//     out[32b + r] += x[32q + m] * h_q[32d + r - m]        r, m = 0..31
// All lanes of the warp sit at the same (d, m) so subchunk boundaries are warp-uniform; each lane
// keeps a 32-entry ring of taps h_q[32d + r - m] in registers and slides it by one tap per m.  Taps
// are blended on the fly from shared memory rows that hold {H_i, H_{i+1} - H_i} for both ears as one
// 16-byte entry per tap (one LDS.128 + one packed FMA per new tap); lanes in the same chunk read the
// same entry (broadcast), lanes in different chunks hit different banks (odd row stride).  The input
// tile sits in shared memory with a 128-byte XOR swizzle so the stride-32 lane pattern is conflict
// free.  FMA : LDS.128 ratio is 1087 : 71 per 32x32 block, so the kernel is bound by the FP32 pipe,
// as SURVEY.md 8(d) predicts (arithmetic intensity ~95-190 flop per HBM byte).
#include "bas_internal.cuh"

namespace {

struct RenderParams {
    const float* x; long long x_stride; long long n_valid;
    int n_src; long long n_in;
    int C, S, K;
    const float* filt; long long filt_stride; long long filt_src_stride;   // floats between sources
    const float* gains;
    long long p_begin, p_end;      // rendered output range [p_begin, p_end)
    float* out; long long out_stride;
    int mix;
    float* peaks;
};

__device__ __forceinline__ void atomic_max_nonneg(float* addr, float v) {
    // v >= 0: IEEE order equals integer order
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------
// generic kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bas_render_generic_kernel(RenderParams prm) {
    const long long p = prm.p_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = p < prm.p_end;
    const int s_first = prm.mix ? 0 : blockIdx.y;
    const int s_last = prm.mix ? prm.n_src : blockIdx.y + 1;
    float mix_l = 0.f, mix_r = 0.f;
    for (int s = s_first; s < s_last; ++s) {
        float acc_l = 0.f, acc_r = 0.f;
        if (live) {
            const float* x = prm.x + (long long)s * prm.x_stride;
            const float* f = prm.filt + (long long)s * prm.filt_src_stride;
            for (int k = 0; k < prm.K; ++k) {
                const long long n = p - k;
                if (n < 0) break;
                if (n >= prm.n_valid) continue;
                const long long chunk = n / prm.C;
                const int j = (int)(n - chunk * prm.C) / prm.S * prm.S;          // apply_hrtf.py:438
                const float alpha = (float)j / (float)prm.C;                     // :442
                const float* h0 = f + chunk * 2 * prm.filt_stride;
                const float* h1 = h0 + 2 * prm.filt_stride;
                const float xv = x[n];
                const float hl = (1.f - alpha) * h0[k] + alpha * h1[k];          // :443
                const float hr = (1.f - alpha) * h0[prm.filt_stride + k] + alpha * h1[prm.filt_stride + k];
                acc_l = fmaf(xv, hl, acc_l);
                acc_r = fmaf(xv, hr, acc_r);
            }
        }
        if (prm.peaks) {
            const float m = warp_max(fmaxf(fabsf(acc_l), fabsf(acc_r)));
            if ((threadIdx.x & 31) == 0 && m > 0.f) atomic_max_nonneg(prm.peaks + s, m);
        }
        const float g = prm.gains ? prm.gains[s] : 1.f;
        if (prm.mix) {
            mix_l = fmaf(g, acc_l, mix_l);
            mix_r = fmaf(g, acc_r, mix_r);
        } else if (live) {
            float* o = prm.out + (long long)s * 2 * prm.out_stride;
            o[p - prm.p_begin] = g * acc_l;
            o[prm.out_stride + p - prm.p_begin] = g * acc_r;
        }
    }
    if (prm.mix && live) {
        prm.out[p - prm.p_begin] = mix_l;
        prm.out[prm.out_stride + p - prm.p_begin] = mix_r;
    }
}

// ------------------------------------------------------------------------------------------------
// tiled kernel
// ------------------------------------------------------------------------------------------------
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// d = a * b + d on both halves (FFMA2 on sm_100)
__device__ __forceinline__ void fma2_acc(u64& d, u64 a, u64 b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

constexpr int kBlk = 32;                 // outputs per lane = subchunk size of the tiled kernel
constexpr int kWarpTile = 32 * kBlk;     // 1024 outputs per tile-warp
constexpr int kStagePitch = kBlk + 1;    // staging row pitch (floats): conflict-free for lane-private rows

// geometry shared by host and device
struct TileGeom {
    int D;              // last tap block: d = 0..D
    int row_len;        // 16-byte entries per filter row (odd)
    int x_rows;         // 32-sample rows of the input tile
    int max_rows;       // filter rows (chunks) a CTA can need
};

__host__ __device__ inline TileGeom tile_geom(int K, int C, int TW) {
    TileGeom g;
    g.D = (K + kBlk - 2) / kBlk;
    g.row_len = kBlk * (g.D + 2) + 1;
    g.x_rows = TW * 32 + g.D;
    g.max_rows = (g.x_rows * kBlk + C - 1) / C + 1;
    return g;
}

// Shared memory: [ area A: filter rows | input tile ]  [ mix tile (only when mixing) ].
// With one source per CTA the output staging tile aliases area A; when mixing, the per-source
// scratch of the tap splits aliases area A and the running mix has its own tile behind it.
__host__ __device__ inline size_t tile_stage_bytes(int TW) { return (size_t)2 * TW * 32 * kStagePitch * 4; }
__host__ __device__ inline size_t tile_area_a_bytes(const TileGeom& g, int TW) {
    const size_t rows_xs = (size_t)g.max_rows * g.row_len * 16 + (size_t)g.x_rows * kBlk * 4;
    const size_t stage = tile_stage_bytes(TW);
    return rows_xs > stage ? rows_xs : stage;
}
inline size_t tile_smem_bytes(const TileGeom& g, int TW, bool mix) {
    return tile_area_a_bytes(g, TW) + (mix ? tile_stage_bytes(TW) : 0);
}

// One 32x32 block of (input sample m) x (output r) products for both ears.
//   acc[r] += {x[m], x[m]} * w[(r - m) & 31],   w slot (j & 31) holds tap 32d + j for j = r - m
// rowp points at tap 32d of the lane's filter row (16-byte entries {H_L, H_R, D_L, D_R});
// xrow points at the lane's swizzled 128-byte input row; swz = row & 7.
__device__ __forceinline__ void block_32x32(u64 (&acc)[kBlk], const ulonglong2* __restrict__ rowp, u64 alpha2,
                                            const float4* __restrict__ xrow, int swz) {
    u64 w[kBlk];
#pragma unroll
    for (int j = 0; j < kBlk; ++j) {
        const ulonglong2 hd = rowp[j];
        w[j] = fma2(alpha2, hd.y, hd.x);           // H + alpha * (H_next - H)   (apply_hrtf.py:443)
    }
#pragma unroll
    for (int m4 = 0; m4 < kBlk / 4; ++m4) {
        const float4 xv = xrow[m4 ^ swz];
        const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) {
            const int m = m4 * 4 + mm;
            if (m > 0) {
                const ulonglong2 hd = rowp[-m];
                w[(kBlk - m) & 31] = fma2(alpha2, hd.y, hd.x);
            }
            const u64 xx = pack2(xs[mm], xs[mm]);
#pragma unroll
            for (int r = 0; r < kBlk; ++r) fma2_acc(acc[r], xx, w[(r - m) & 31]);
        }
    }
}

template <int TW, int TS>
__global__ void __launch_bounds__(TW * TS * 32)
bas_render_tiled_kernel(RenderParams prm) {
    extern __shared__ __align__(16) unsigned char smem[];
    const TileGeom g = tile_geom(prm.K, prm.C, TW);
    float4* rows = reinterpret_cast<float4*>(smem);
    float4* xs = rows + (size_t)g.max_rows * g.row_len;
    float* stage = reinterpret_cast<float*>(prm.mix ? smem + tile_area_a_bytes(g, TW) : smem);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tw = warp % TW, ts = warp / TW;
    constexpr int kThreads = TW * TS * 32;

    const long long p_base = prm.p_begin / kBlk * kBlk;
    const long long P0 = p_base + (long long)blockIdx.x * (TW * kWarpTile);     // first output of the CTA
    const long long n_lo = P0 - (long long)kBlk * g.D;                          // first input of the tile
    const long long q0 = n_lo / kBlk;                                           // exact: n_lo % 32 == 0 (may be < 0)
    const int spc = prm.C / kBlk;                                               // subchunks per chunk
    const long long n_chunks = prm.n_in / prm.C;
    // chunks touched by the tile's inputs, clamped to the signal
    long long c_first = n_lo < 0 ? 0 : n_lo / prm.C;
    long long c_last = (P0 + (long long)TW * kWarpTile - 1) / prm.C;
    if (c_last > n_chunks - 1) c_last = n_chunks - 1;
    if (c_first > c_last) c_first = c_last;
    const int n_rows = (int)(c_last - c_first + 1);

    const int blk = tw * 32 + lane;                 // output block of this lane inside the CTA tile
    // tap blocks of this warp: d in [d_first, d_last)
    const int d_per = (g.D + 1 + TS - 1) / TS;
    const int d_first = ts * d_per;
    const int d_last = min(g.D + 1, d_first + d_per);
    const bool warp_live = P0 + (long long)tw * kWarpTile < prm.p_end;

    const int s_first = prm.mix ? 0 : blockIdx.y;
    const int s_last = prm.mix ? prm.n_src : blockIdx.y + 1;

    if (prm.mix) {
        for (int i = tid; i < 2 * TW * 32 * kStagePitch; i += kThreads) stage[i] = 0.f;
    }

    for (int s = s_first; s < s_last; ++s) {
        // ---- stage the input tile (zero outside [0, n_valid)) ---------------------------------
        const float* x = prm.x + (long long)s * prm.x_stride;
        for (int c = tid; c < g.x_rows * 8; c += kThreads) {
            const int row = c >> 3, ch = c & 7;
            const long long n = n_lo + (long long)row * kBlk + ch * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n >= 0 && n + 3 < prm.n_valid) {
                v = *reinterpret_cast<const float4*>(x + n);
            } else if (n + 3 >= 0 && n < prm.n_valid) {
                if (n >= 0 && n < prm.n_valid) v.x = x[n];
                if (n + 1 >= 0 && n + 1 < prm.n_valid) v.y = x[n + 1];
                if (n + 2 >= 0 && n + 2 < prm.n_valid) v.z = x[n + 2];
                if (n + 3 >= 0 && n + 3 < prm.n_valid) v.w = x[n + 3];
            }
            xs[row * 8 + (ch ^ (row & 7))] = v;
        }
        // ---- stage the filter rows: {H_i, H_{i+1} - H_i} for both ears, zero padded --------------
        const float* f = prm.filt + (long long)s * prm.filt_src_stride;
        for (int e = tid; e < n_rows * g.row_len; e += kThreads) {
            const int ri = e / g.row_len, idx = e - ri * g.row_len;
            const int k = idx - kBlk;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k >= 0 && k < prm.K) {
                const float* h0 = f + (c_first + ri) * 2 * prm.filt_stride;
                const float* h1 = h0 + 2 * prm.filt_stride;
                const float l0 = h0[k], r0 = h0[prm.filt_stride + k];
                v = make_float4(l0, r0, h1[k] - l0, h1[prm.filt_stride + k] - r0);
            }
            rows[e] = v;
        }
        __syncthreads();

        // ---- the FIR ----------------------------------------------------------------------------
        u64 acc[kBlk];
#pragma unroll
        for (int r = 0; r < kBlk; ++r) acc[r] = 0ull;
        if (warp_live) {
            for (int d = d_first; d < d_last; ++d) {
                const int xrow = blk + g.D - d;                     // input row (subchunk) inside the tile
                const long long q = q0 + xrow;                      // absolute subchunk
                long long chunk = q < 0 ? 0 : q / spc;
                const int sub = q < 0 ? 0 : (int)(q - chunk * spc);
                if (chunk < c_first) chunk = c_first;               // only for rows whose samples are all zero
                if (chunk > c_last) chunk = c_last;
                const float alpha = (float)(sub * kBlk) / (float)prm.C;         // apply_hrtf.py:442
                const ulonglong2* rowp = reinterpret_cast<const ulonglong2*>(rows) +
                                         (size_t)(chunk - c_first) * g.row_len + kBlk + d * kBlk;
                block_32x32(acc, rowp, pack2(alpha, alpha), xs + xrow * 8, xrow & 7);
            }
        }

        // ---- combine tap splits, peak, mix -------------------------------------------------------
        if (!prm.mix) __syncthreads();          // staging aliases rows/xs: everyone must be done reading
        const float gain = prm.gains ? prm.gains[s] : 1.f;
        float* my_l = stage + blk * kStagePitch;
        float* my_r = stage + (TW * 32 + blk) * kStagePitch;
        if (TS == 1 && !prm.mix) {
            float pk = 0.f;
            const long long pb = P0 + (long long)blk * kBlk;
#pragma unroll
            for (int r = 0; r < kBlk; ++r) {
                float l, rr; unpack2(acc[r], l, rr);
                const bool in = pb + r >= prm.p_begin && pb + r < prm.p_end;
                if (in) pk = fmaxf(pk, fmaxf(fabsf(l), fabsf(rr)));
                my_l[r] = gain * l; my_r[r] = gain * rr;
            }
            if (prm.peaks) { pk = warp_max(pk); if (lane == 0 && pk > 0.f) atomic_max_nonneg(prm.peaks + s, pk); }
        } else {
            // per-source tile in registers -> the tap splits take turns adding into a scratch copy
            // (fixed order: deterministic), the last one takes the peak and folds into the mix.
            // Scratch = lane-private slots, so only block-level ordering between turns is needed.
            float* scr_l = my_l; float* scr_r = my_r;
            if (prm.mix) {   // scratch for the per-source sum lives in the (now idle) input tile area
                float* scr = reinterpret_cast<float*>(rows);
                __syncthreads();
                scr_l = scr + blk * kStagePitch; scr_r = scr + (TW * 32 + blk) * kStagePitch;
            }
            for (int turn = 0; turn < TS; ++turn) {
                if (ts == turn) {
                    float pk = 0.f;
                    const long long pb = P0 + (long long)blk * kBlk;
#pragma unroll
                    for (int r = 0; r < kBlk; ++r) {
                        float l, rr; unpack2(acc[r], l, rr);
                        if (turn > 0) { l += scr_l[r]; rr += scr_r[r]; }
                        if (turn == TS - 1) {
                            const bool in = pb + r >= prm.p_begin && pb + r < prm.p_end;
                            if (in) pk = fmaxf(pk, fmaxf(fabsf(l), fabsf(rr)));
                            if (prm.mix) { my_l[r] = fmaf(gain, l, my_l[r]); my_r[r] = fmaf(gain, rr, my_r[r]); }
                            else { my_l[r] = gain * l; my_r[r] = gain * rr; }
                        } else { scr_l[r] = l; scr_r[r] = rr; }
                    }
                    if (turn == TS - 1 && prm.peaks) {
                        pk = warp_max(pk);
                        if (lane == 0 && pk > 0.f) atomic_max_nonneg(prm.peaks + s, pk);
                    }
                }
                if (turn + 1 < TS) __syncthreads();
            }
        }
        __syncthreads();

        // ---- store (one source per CTA) ------------------------------------------------------------
        if (!prm.mix) {
            float* o = prm.out + (long long)s * 2 * prm.out_stride;
            for (int i = tid; i < 2 * TW * kWarpTile; i += kThreads) {
                const int e = i / (TW * kWarpTile), idx = i - e * (TW * kWarpTile);
                const long long p = P0 + idx;
                if (p >= prm.p_begin && p < prm.p_end)
                    o[e * prm.out_stride + (p - prm.p_begin)] = stage[(e * TW * 32 + (idx >> 5)) * kStagePitch + (idx & 31)];
            }
        }
    }
    if (prm.mix) {
        for (int i = tid; i < 2 * TW * kWarpTile; i += kThreads) {
            const int e = i / (TW * kWarpTile), idx = i - e * (TW * kWarpTile);
            const long long p = P0 + idx;
            if (p >= prm.p_begin && p < prm.p_end)
                prm.out[e * prm.out_stride + (p - prm.p_begin)] = stage[(e * TW * 32 + (idx >> 5)) * kStagePitch + (idx & 31)];
        }
    }
}

template <int TW, int TS>
int launch_tiled(const RenderParams& prm, cudaStream_t st) {
    const TileGeom g = tile_geom(prm.K, prm.C, TW);
    const size_t smem = tile_smem_bytes(g, TW, prm.mix != 0);
    if (smem > 227 * 1024) return BAS_E_UNSUPPORTED;
    auto kern = bas_render_tiled_kernel<TW, TS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { bas_set_error("bas_render: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    const long long p_base = prm.p_begin / kBlk * kBlk;
    const long long tiles = bas_ceil_div(prm.p_end - p_base, (long long)TW * kWarpTile);
    if (tiles > 0x7fffffffLL || prm.n_src > 65535) return BAS_E_UNSUPPORTED;
    dim3 grid((unsigned)tiles, prm.mix ? 1u : (unsigned)prm.n_src);
    kern<<<grid, TW * TS * 32, smem, st>>>(prm);
    e = cudaGetLastError();
    if (e != cudaSuccess) { bas_set_error("bas_render: tiled launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// peak / normalise  (apply_hrtf.py:462-464)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bas_peak_kernel(const float* __restrict__ v, long long n, float* __restrict__ peak) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(v[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomic_max_nonneg(peak, m);
}

__global__ void __launch_bounds__(256)
bas_normalise_kernel(float* __restrict__ v, long long n, const float* __restrict__ peak) {
    const float m = *peak;
    if (!(m > 1.f)) return;                      // apply_hrtf.py:463: only when the peak exceeds 1
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        v[i] = v[i] / m;                         // :464 true division, not a reciprocal multiply
}

}  // namespace

// variant encoding beyond the public three: BAS_RENDER_TILED | (TW << 8) | (TS << 16) picks a tile shape
// (used by the tuning sweep in bench.py; unknown shapes return BAS_E_UNSUPPORTED).
extern "C" int bas_render(const float* x_dev, long long x_stride, long long n_valid, int n_src, long long n_in,
                          int C, int S, int K, const float* filt_dev, long long filt_stride, const float* gains_dev,
                          long long p_begin, long long p_count, float* out_dev, long long out_stride, int mix,
                          float* peaks_dev, int variant, void* stream) {
    BAS_CHECK_ARG(x_dev && filt_dev && out_dev, "null pointer");
    BAS_CHECK_ARG(n_src >= 1, "n_src");
    BAS_CHECK_ARG(C >= 1 && S >= 1 && C % S == 0, "subchunksize must divide chunksize");     // apply_hrtf.py:401-402
    BAS_CHECK_ARG(K >= 1 && filt_stride >= K, "need 1 <= K <= filt_stride");
    BAS_CHECK_ARG(n_in >= C && n_in % C == 0, "n_in must be a positive multiple of C");      // apply_hrtf.py:405
    BAS_CHECK_ARG(n_valid >= 0 && n_valid <= n_in && (n_src == 1 || x_stride >= n_valid), "n_valid / x_stride");
    BAS_CHECK_ARG(p_begin >= 0 && p_count >= 0 && p_begin + p_count <= n_in + K - 1, "output range");
    BAS_CHECK_ARG(out_stride >= p_count, "out_stride < p_count");
    if (p_count == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    RenderParams prm;
    prm.x = x_dev; prm.x_stride = x_stride; prm.n_valid = n_valid; prm.n_src = n_src; prm.n_in = n_in;
    prm.C = C; prm.S = S; prm.K = K; prm.filt = filt_dev; prm.filt_stride = filt_stride;
    prm.filt_src_stride = (n_in / C + 1) * 2 * filt_stride;
    prm.gains = gains_dev; prm.p_begin = p_begin; prm.p_end = p_begin + p_count;
    prm.out = out_dev; prm.out_stride = out_stride; prm.mix = mix ? 1 : 0; prm.peaks = peaks_dev;

    const int base = variant & 0xff;
    BAS_CHECK_ARG(base == BAS_RENDER_AUTO || base == BAS_RENDER_GENERIC || base == BAS_RENDER_TILED, "variant");
    const bool tiled_ok = S == kBlk && C % kBlk == 0 && K <= 1024 &&
                          (reinterpret_cast<uintptr_t>(x_dev) & 15) == 0 && (n_src == 1 || x_stride % 4 == 0);
    if (base == BAS_RENDER_TILED && !tiled_ok) {
        bas_set_error("bas_render: tiled kernel needs S == 32, 32 | C, K <= 1024, 16-byte aligned signals");
        return BAS_E_UNSUPPORTED;
    }
    if (base != BAS_RENDER_GENERIC && tiled_ok) {
        int tw = (variant >> 8) & 0xff, tsp = (variant >> 16) & 0xff;
        if (tw == 0 && tsp == 0) {
            // default shapes: many small CTAs when one source must fill the GPU, larger tiles for a mix
            if (prm.mix) { tw = 2; tsp = 3; } else { tw = 1; tsp = 3; }
        }
        int rc = BAS_E_UNSUPPORTED;
#define BAS_TILE_CASE(TW_, TS_) if (tw == TW_ && tsp == TS_) rc = launch_tiled<TW_, TS_>(prm, st)
        BAS_TILE_CASE(1, 1); BAS_TILE_CASE(1, 3); BAS_TILE_CASE(2, 1); BAS_TILE_CASE(2, 3); BAS_TILE_CASE(4, 1);
        BAS_TILE_CASE(1, 2); BAS_TILE_CASE(2, 2); BAS_TILE_CASE(4, 2); BAS_TILE_CASE(1, 4);
#undef BAS_TILE_CASE
        if (rc != BAS_E_UNSUPPORTED || base == BAS_RENDER_TILED) {
            if (rc == BAS_E_UNSUPPORTED) bas_set_error("bas_render: tile shape %dx%d unavailable for K=%d C=%d", tw, tsp, K, C);
            return rc;
        }
    }
    const int threads = 256;
    const long long blocks = bas_ceil_div(p_count, threads);
    BAS_CHECK_ARG(blocks < 0x7fffffffLL && n_src <= 65535, "launch too large");
    dim3 grid((unsigned)blocks, prm.mix ? 1u : (unsigned)n_src);
    bas_render_generic_kernel<<<grid, threads, 0, st>>>(prm);
    BAS_LAUNCH_CHECK();
    return 0;
}

extern "C" int bas_peak(const float* v_dev, long long n, float* peak_dev, void* stream) {
    BAS_CHECK_ARG(v_dev && peak_dev && n >= 0, "bad pointer or size");
    if (n == 0) return 0;
    const long long blocks = bas_ceil_div(n, 256 * 8);
    bas_peak_kernel<<<(unsigned)(blocks > 148 * 8 ? 148 * 8 : blocks), 256, 0, (cudaStream_t)stream>>>(v_dev, n, peak_dev);
    BAS_LAUNCH_CHECK();
    return 0;
}

extern "C" int bas_normalise(float* out_dev, long long n, const float* peak_dev, void* stream) {
    BAS_CHECK_ARG(out_dev && peak_dev && n >= 0, "bad pointer or size");
    if (n == 0) return 0;
    const long long blocks = bas_ceil_div(n, 256 * 8);
    bas_normalise_kernel<<<(unsigned)(blocks > 148 * 8 ? 148 * 8 : blocks), 256, 0, (cudaStream_t)stream>>>(out_dev, n, peak_dev);
    BAS_LAUNCH_CHECK();
    return 0;
}
