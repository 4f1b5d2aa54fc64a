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
// bas_render_tiled_kernel     S == 32: persistent, TMA-fed, register-tiled kernel on the packed
//                             FP32 FMA pipe.
//
// Tiled kernel
//   work     A CTA of TW warps owns tiles of TW*1024 consecutive outputs (one 1024-output stripe
//            per warp) and walks over its tiles persistently (grid = resident CTAs).
//   staging  For every (tile, source) item one warp issues cp.async.bulk (TMA, 1-D) copies into a
//            two-stage shared-memory ring guarded by full/empty mbarriers: the input samples as
//            128-byte rows on a 144-byte pitch (conflict-free for the lane-per-row reads below) and
//            the boundary-filter rows of the chunks the tile touches as ONE contiguous copy (rows
//            come from ir_synth already interleaved {L,R} and zero padded).  The copy of item j+1
//            overlaps the arithmetic of item j; there is no __syncthreads in the kernel.
//   math     Lane l of a warp owns the 32 consecutive outputs of block b = b0 + l for both ears: 32
//            fma.rn.f32x2 accumulators {L,R}.  Output block b draws on input subchunks q = b - d:
//                out[32b + r] += x[32q + m] * h_q[32d + r - m]          r, m = 0..31
//            All lanes sit at the same (d, m), so subchunk boundaries are warp-uniform; each lane
//            keeps a 32-entry ring of blended taps in registers and slides it one tap per m.  The
//            half-empty first (d = 0, r >= m) and last (d = D, r < m) blocks are folded into ONE
//            32x32 block whose ring is initialised from the d = 0 filter and refilled from the
//            d = D filter, so a tile costs exactly ceil(K/32) blocks of 1024 packed FMAs plus 126
//            blend operations each (89 % useful at K = 256, 94 % at K = 512).
//   taps     Blended on the fly: w = H_i + alpha (H_{i+1} - H_i) from two LDS.128 (two taps each);
//            lanes in the same chunk read the same address (broadcast), other chunks other banks.
//   epilogue Direct 16-byte global stores from registers, peak by warp reduction + one atomicMax.
#include "bas_internal.cuh"

namespace {

struct RenderParams {
    const float* x; long long x_stride; long long n_valid;
    int n_src; long long n_in;
    int C, S, K;
    const float2* filt;            // [n_src][n_in/C + 1][pitch] {L, R}
    int pitch;                     // taps per filter row
    long long filt_src_stride;     // float2 entries between sources
    const float* gains;
    long long p_begin, p_end;      // rendered output range [p_begin, p_end)
    float* out; long long out_stride;
    int mix;
    float* peaks;
    long long tiles;               // tiled kernel: tiles per source
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
            const float2* f = prm.filt + (long long)s * prm.filt_src_stride;
            for (int k = 0; k < prm.K; ++k) {
                const long long n = p - k;
                if (n < 0) break;
                if (n >= prm.n_valid) continue;
                const long long chunk = n / prm.C;
                const int j = (int)(n - chunk * prm.C) / prm.S * prm.S;          // apply_hrtf.py:438
                const float alpha = (float)j / (float)prm.C;                     // :442
                const float2 h0 = f[chunk * prm.pitch + k];
                const float2 h1 = f[(chunk + 1) * prm.pitch + k];
                const float xv = x[n];
                acc_l = fmaf(xv, (1.f - alpha) * h0.x + alpha * h1.x, acc_l);    // :443, :445
                acc_r = fmaf(xv, (1.f - alpha) * h0.y + alpha * h1.y, acc_r);    // :443, :446
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
// tiled kernel: PTX helpers
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
__device__ __forceinline__ void fma2_acc(u64& d, u64 a, u64 b) {          // d += a * b  (FFMA2)
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (UBLKCP in SASS)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

constexpr int kBlk = 32;                 // outputs per lane = subchunk size of the tiled kernel
constexpr int kWarpTile = 32 * kBlk;     // 1024 outputs per warp
constexpr int kXPitch = 36;              // floats per staged input row (32 samples + 16 bytes)
constexpr int kStages = 2;

// geometry shared by host and device
struct TileGeom {
    int D;              // tap blocks per tile: ceil(K / 32)
    int x_rows;         // 32-sample input rows of a tile
    int f_rows;         // filter rows staged per tile (chunks touched + 1)
    int w_rows;         // input rows one warp reads (its 32 blocks + D)
    unsigned x_bytes, f_bytes, stage_bytes, warp_x_bytes;
};

__host__ __device__ inline TileGeom tile_geom(int K, int C, int pitch, int TW) {
    TileGeom g;
    g.D = (K + kBlk - 1) / kBlk;
    g.x_rows = TW * 32 + g.D;
    g.f_rows = (g.x_rows * kBlk + C - 1) / C + 2;
    g.w_rows = 32 + g.D;
    g.x_bytes = (unsigned)g.x_rows * kBlk * 4;            // staged linearly (one bulk copy)
    g.f_bytes = (unsigned)g.f_rows * pitch * 8;
    g.stage_bytes = g.x_bytes + g.f_bytes;
    g.warp_x_bytes = (unsigned)g.w_rows * kXPitch * 4;    // per-warp copy on the conflict-free pitch
    return g;
}
__host__ __device__ inline size_t tile_smem_bytes(const TileGeom& g, int TW) {
    return 64 + (size_t)kStages * g.stage_bytes + (size_t)TW * g.warp_x_bytes;
}
constexpr size_t kBarBytes = 64;         // 2 x full + 2 x empty mbarriers, padded

// One 32x32 block: acc[r] += x_sel[m] * w[(r - m) & 31] for both ears, where ring slot j holds
//   tap (base_a + j)        of the blend of rows (ra, ra + pitch)   for j = r - m >= 0   (x from xa)
//   tap (base_b + j - 32)   of the blend of rows (rb, rb + pitch)   for j - 32 = r - m < 0 (x from xb)
// ra/rb point at tap base_a / base_b of the lane's filter row (float2 {L,R} entries, 16-byte aligned).
// For a full block (a, b) are the same filter and base_b = base_a; for the folded first/last block
// a is the d = 0 filter (base 0) and b the d = D filter (base 32 D).
__device__ __forceinline__ void block_32x32(u64 (&acc)[kBlk], const float2* __restrict__ ra, u64 alpha_a,
                                            const float2* __restrict__ rb, u64 alpha_b, int pitch,
                                            const float* __restrict__ xa, const float* __restrict__ xb) {
    u64 w[kBlk];
#pragma unroll
    for (int j = 0; j < kBlk; j += 2) {
        const ulonglong2 h0 = *reinterpret_cast<const ulonglong2*>(ra + j);
        const ulonglong2 h1 = *reinterpret_cast<const ulonglong2*>(ra + pitch + j);
        w[j] = fma2(alpha_a, sub2(h1.x, h0.x), h0.x);          // H_i + alpha (H_{i+1} - H_i)   apply_hrtf.py:443
        w[j + 1] = fma2(alpha_a, sub2(h1.y, h0.y), h0.y);
    }
    u64 pending = 0ull;
#pragma unroll
    for (int m4 = 0; m4 < kBlk / 4; ++m4) {
        const float4 xav = *reinterpret_cast<const float4*>(xa + 4 * m4);
        const float4 xbv = *reinterpret_cast<const float4*>(xb + 4 * m4);
        const float xas[4] = {xav.x, xav.y, xav.z, xav.w};
        const float xbs[4] = {xbv.x, xbv.y, xbv.z, xbv.w};
#pragma unroll
        for (int mm = 0; mm < 4; ++mm) {
            const int m = m4 * 4 + mm;
            if (m > 0) {
                if (m & 1) {          // taps (base_b - m - 1, base_b - m) in one 16-byte load
                    const ulonglong2 h0 = *reinterpret_cast<const ulonglong2*>(rb - m - 1);
                    const ulonglong2 h1 = *reinterpret_cast<const ulonglong2*>(rb + pitch - m - 1);
                    w[(kBlk - m) & 31] = fma2(alpha_b, sub2(h1.y, h0.y), h0.y);
                    pending = fma2(alpha_b, sub2(h1.x, h0.x), h0.x);
                } else {
                    w[(kBlk - m) & 31] = pending;
                }
            }
            const u64 xxa = pack2(xas[mm], xas[mm]);
            const u64 xxb = pack2(xbs[mm], xbs[mm]);
#pragma unroll
            for (int r = 0; r < kBlk; ++r) fma2_acc(acc[r], r >= m ? xxa : xxb, w[(r - m) & 31]);
        }
    }
}

// ---- work decomposition (stream-K) -----------------------------------------------------------------
// The work of a launch is a line of SLICES grouped into GROUPS of gs slices that share one output
// tile:   one source per tile (MIX = false): group = (source, tile), slice = tap block d, gs = D
//         mixing                (MIX = true): group = tile,           slice = source,      gs = n_src
// CTA c owns the contiguous slice span [c*total/G, (c+1)*total/G).  A span boundary that falls
// inside a group splits that group between exactly two CTAs (spans are longer than a group); both
// write their partial tile to the workspace and bas_render_fixup_kernel adds the two in a fixed order,
// so results stay deterministic while every scheduler gets the same number of 32x32 blocks.  Without
// a workspace spans are rounded to group boundaries.
struct SpanInfo { long long total, n_groups; int gs; int split; };

__host__ __device__ inline long long span_begin(const SpanInfo& sp, long long c, long long G) {
    if (sp.split) return (long long)(((unsigned long long)c * (unsigned long long)sp.total) / (unsigned long long)G);
    return (long long)(((unsigned long long)c * (unsigned long long)sp.n_groups) / (unsigned long long)G) * sp.gs;
}

struct Item {
    long long tile; int src;
    int d0, d1;             // tap blocks to run
    bool group_end;         // outputs of the group are complete (for this CTA) after this item
    bool partial;           // this CTA holds only part of the group -> workspace
    int slot;               // workspace slot: 0 = group began in the previous CTA, 1 = continues in the next
};

template <int TW, bool MIX>
__global__ void __launch_bounds__(TW * 32, 1)
bas_render_tiled_kernel(RenderParams prm, SpanInfo sp, float* __restrict__ workspace) {
    extern __shared__ __align__(128) unsigned char smem[];
    const TileGeom g = tile_geom(prm.K, prm.C, prm.pitch, TW);
    u64* full_bar = reinterpret_cast<u64*>(smem);              // [kStages]
    u64* empty_bar = full_bar + kStages;                       // [kStages]
    unsigned char* stage_base = smem + kBarBytes;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* xw = reinterpret_cast<float*>(stage_base + (size_t)kStages * g.stage_bytes + (size_t)warp * g.warp_x_bytes);
    const long long p_base = prm.p_begin / kBlk * kBlk;
    const int spc = prm.C / kBlk;                              // subchunks per chunk
    const long long n_chunks = prm.n_in / prm.C;
    // split: contiguous slice span [i0, i1);  otherwise whole groups, dealt round-robin (group = c + k G)
    const long long i0 = sp.split ? span_begin(sp, blockIdx.x, gridDim.x) : 0;
    const long long i1 = sp.split ? span_begin(sp, blockIdx.x + 1, gridDim.x) : 0;
    const long long g_first = i0 / sp.gs;
    const long long my_groups = blockIdx.x < sp.n_groups ? (sp.n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long n_items = sp.split ? (i1 <= i0 ? 0 : (MIX ? i1 - i0 : (i1 - 1) / sp.gs - g_first + 1))
                                       : my_groups * (MIX ? sp.gs : 1);

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, TW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // item j of this CTA
    auto item_info = [&](long long j) {
        Item it;
        long long grp; int a, b;
        if (sp.split) {
            grp = MIX ? (i0 + j) / sp.gs : g_first + j;
            const long long lo = grp * sp.gs;
            a = (int)((i0 > lo ? i0 : lo) - lo); b = (int)((i1 < lo + sp.gs ? i1 : lo + sp.gs) - lo);
            it.src = MIX ? (int)(i0 + j - lo) : 0;
        } else {
            const long long k = MIX ? j / sp.gs : j;
            grp = blockIdx.x + k * gridDim.x;
            a = 0; b = sp.gs;
            it.src = MIX ? (int)(j - k * sp.gs) : 0;
        }
        it.partial = a > 0 || b < sp.gs;
        it.slot = a > 0 ? 0 : 1;
        if (MIX) {
            it.tile = grp;
            it.d0 = 0; it.d1 = g.D;
            it.group_end = it.src == b - 1;
        } else {
            it.src = (int)(grp / prm.tiles); it.tile = grp - (long long)it.src * prm.tiles;
            it.d0 = a; it.d1 = b;
            it.group_end = true;
        }
        return it;
    };
    // chunk range whose filter rows a tile needs (clamped to the signal)
    auto tile_chunks = [&](long long tile, long long& n_lo, long long& c_first, int& n_rows) {
        const long long P0 = p_base + tile * (TW * kWarpTile);
        n_lo = P0 - (long long)kBlk * g.D;
        c_first = n_lo < 0 ? 0 : n_lo / prm.C;
        long long c_last = (P0 + (long long)TW * kWarpTile - 1) / prm.C;
        if (c_last > n_chunks - 1) c_last = n_chunks - 1;
        if (c_first > c_last) c_first = c_last;
        n_rows = (int)(c_last - c_first + 2);                   // + the boundary after the last chunk
    };

    // ---- producer: warp 0 stages item j into ring slot j % kStages --------------------------------
    auto produce = [&](long long j) {
        const int st = (int)(j % kStages);
        const long long use = j / kStages;
        if (use > 0) mbar_wait(empty_bar + st, (unsigned)((use - 1) & 1));      // consumers left the slot
        const Item it = item_info(j);
        long long n_lo, c_first; int n_rows;
        tile_chunks(it.tile, n_lo, c_first, n_rows);
        float* xs = reinterpret_cast<float*>(stage_base + (size_t)st * g.stage_bytes);
        float2* fs = reinterpret_cast<float2*>(stage_base + (size_t)st * g.stage_bytes + g.x_bytes);
        const float* x = prm.x + (long long)it.src * prm.x_stride;
        if (lane == 0) {
            // two bulk copies per item: the in-range part of the input span, and the filter rows.
            // Samples outside [0, n_valid) are never copied; consumers zero them while re-laying out.
            const long long na = n_lo < 0 ? 0 : n_lo;
            long long nb = n_lo + (long long)g.x_rows * kBlk;
            if (nb > prm.n_valid) nb = prm.n_valid;
            const unsigned x_bytes = nb > na ? (unsigned)(nb - na) * 4 : 0;
            const unsigned f_bytes = (unsigned)n_rows * prm.pitch * 8;
            mbar_arrive_expect_tx(full_bar + st, x_bytes + f_bytes);
            if (x_bytes) bulk_g2s(xs + (na - n_lo), x + na, x_bytes, full_bar + st);
            bulk_g2s(fs, prm.filt + (long long)it.src * prm.filt_src_stride + c_first * prm.pitch, f_bytes, full_bar + st);
        }
        __syncwarp();
    };

    if (warp == 0 && n_items > 0) produce(0);

    u64 acc[kBlk];
    u64 mixacc[MIX ? kBlk : 1];
#pragma unroll
    for (int r = 0; r < kBlk; ++r) acc[r] = 0ull;
    if (MIX) {
#pragma unroll
        for (int r = 0; r < (MIX ? kBlk : 1); ++r) mixacc[r] = 0ull;
    }
    const int blk = warp * 32 + lane;                           // output block of this lane inside the tile
    const bool vec_ok = (prm.p_begin & 3) == 0 && (prm.out_stride & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(prm.out) & 15) == 0;

    for (long long j = 0; j < n_items; ++j) {
        if (warp == 0 && j + 1 < n_items) produce(j + 1);
        const int st = (int)(j % kStages);
        const Item it = item_info(j);
        long long n_lo, c_first; int n_rows;
        tile_chunks(it.tile, n_lo, c_first, n_rows);
        const long long P0 = p_base + it.tile * (TW * kWarpTile);
        const bool warp_live = P0 + (long long)warp * kWarpTile < prm.p_end;
        const float* xs = reinterpret_cast<const float*>(stage_base + (size_t)st * g.stage_bytes);
        const float2* fs = reinterpret_cast<const float2*>(stage_base + (size_t)st * g.stage_bytes + g.x_bytes);
        const long long q0 = n_lo / kBlk;                       // exact (n_lo % 32 == 0), may be negative

        mbar_wait(full_bar + st, (unsigned)((j / kStages) & 1));

        if (warp_live) {
            // re-lay this warp's input rows from the linear staging buffer onto the 144-byte pitch
            // (lane-per-row reads below are then conflict free) and zero what lies outside the signal
            const long long n_w = n_lo + (long long)warp * kWarpTile;
            const float4* lin = reinterpret_cast<const float4*>(xs) + warp * (kWarpTile / 4);
            for (int idx = lane; idx < g.w_rows * 8; idx += 32) {
                const int row = idx >> 3, ch = idx & 7;
                const long long n = n_w + (long long)row * kBlk + ch * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (n >= 0 && n + 4 <= prm.n_valid) v = lin[idx];
                *reinterpret_cast<float4*>(xw + row * kXPitch + ch * 4) = v;
            }
            __syncwarp();
        }
        if (warp_live) {
            // Filter row (chunk) and blend weight of the lane's input row.  Input rows are visited in
            // descending order (xrow = blk + D - d), so (chunk, sub) is divided once per item and then
            // stepped; rows outside the staged range belong to input rows that are all zero (clamped).
            const int q_top = (int)q0 + blk + g.D;               // absolute subchunk of the d = 0 row
            const int cf = (int)c_first;
            int chunk_a = q_top < 0 ? 0 : q_top / spc;
            int sub_a = q_top < 0 ? 0 : q_top - chunk_a * spc;
            auto row_of = [&](int chunk, const float2*& rowp) {
                int ri = chunk - cf;
                ri = ri < 0 ? 0 : (ri > n_rows - 2 ? n_rows - 2 : ri);
                rowp = fs + ri * prm.pitch;
            };
            // d = 0 is the folded block: ring initialised from the d = 0 filter (r >= m), refilled from
            // the d = D filter (r < m).  One call site keeps the unrolled body in the instruction cache.
#pragma unroll 1
            for (int d = it.d0; d < it.d1; ++d) {
                // (chunk, sub) of row q_top - d
                int chunk = chunk_a, sub = sub_a - d;
                if (q_top - d < 0) { chunk = 0; sub = 0; }
                else { while (sub < 0) { sub += spc; --chunk; } }
                const float2 *ra, *rb;
                row_of(chunk, ra);
                ra += kBlk * d;
                const float alpha = (float)(sub * kBlk) / (float)prm.C;          // apply_hrtf.py:442
                u64 aa = pack2(alpha, alpha), ab = aa;
                rb = ra;
                const int xrow_a = blk + g.D - d;
                int xrow_b = xrow_a;
                if (d == 0) {
                    xrow_b = blk;
                    const int qb = q_top - g.D;
                    const int chunk_b = qb < 0 ? 0 : qb / spc;
                    const int sub_b = qb < 0 ? 0 : qb - chunk_b * spc;
                    row_of(chunk_b, rb);
                    rb += kBlk * g.D;
                    const float alpha_b = (float)(sub_b * kBlk) / (float)prm.C;
                    ab = pack2(alpha_b, alpha_b);
                }
                block_32x32(acc, ra, aa, rb, ab, prm.pitch, xw + (xrow_a - warp * 32) * kXPitch, xw + (xrow_b - warp * 32) * kXPitch);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar + st);             // this warp is done with the slot

        // ---- peak, gain, mix, store ----------------------------------------------------------------
        const long long pb = P0 + (long long)blk * kBlk;        // first output of this lane
        const float gain = prm.gains ? prm.gains[it.src] : 1.f;
        if (prm.peaks && (MIX || !it.partial)) {                // split tiles get their peak in the fix-up
            float pk = 0.f;
#pragma unroll
            for (int r = 0; r < kBlk; ++r) {
                float l, rr; unpack2(acc[r], l, rr);
                if (pb + r >= prm.p_begin && pb + r < prm.p_end) pk = fmaxf(pk, fmaxf(fabsf(l), fabsf(rr)));
            }
            pk = warp_max(pk);
            if (lane == 0 && pk > 0.f) atomic_max_nonneg(prm.peaks + it.src, pk);
        }
        if (MIX) {
            const u64 g2 = pack2(gain, gain);
#pragma unroll
            for (int r = 0; r < kBlk; ++r) { mixacc[r & (MIX ? 31 : 0)] = fma2(g2, acc[r], mixacc[r & (MIX ? 31 : 0)]); acc[r] = 0ull; }
        }
        if (it.group_end) {
            if (it.partial) {
                // partial tile -> workspace[cta][slot][ear][TW*1024], no gain (one source per tile) / mixed
                float* w = workspace + ((long long)blockIdx.x * 2 + it.slot) * (2 * TW * kWarpTile) + blk * kBlk;
#pragma unroll
                for (int r4 = 0; r4 < kBlk; r4 += 4) {
                    float l[4], rr[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) unpack2(MIX ? mixacc[(r4 + i) & (MIX ? 31 : 0)] : acc[r4 + i], l[i], rr[i]);
                    *reinterpret_cast<float4*>(w + r4) = make_float4(l[0], l[1], l[2], l[3]);
                    *reinterpret_cast<float4*>(w + TW * kWarpTile + r4) = make_float4(rr[0], rr[1], rr[2], rr[3]);
                }
            } else {
                float* o = prm.out + (MIX ? 0 : (long long)it.src * 2 * prm.out_stride);
                const long long off = pb - prm.p_begin;
                if (vec_ok && pb >= prm.p_begin && pb + kBlk <= prm.p_end) {
#pragma unroll
                    for (int r4 = 0; r4 < kBlk; r4 += 4) {
                        float l[4], rr[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            unpack2(MIX ? mixacc[(r4 + i) & (MIX ? 31 : 0)] : acc[r4 + i], l[i], rr[i]);
                            if (!MIX) { l[i] *= gain; rr[i] *= gain; }
                        }
                        *reinterpret_cast<float4*>(o + off + r4) = make_float4(l[0], l[1], l[2], l[3]);
                        *reinterpret_cast<float4*>(o + prm.out_stride + off + r4) = make_float4(rr[0], rr[1], rr[2], rr[3]);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < kBlk; ++r) {
                        float l, rr; unpack2(MIX ? mixacc[r & (MIX ? 31 : 0)] : acc[r], l, rr);
                        if (!MIX) { l *= gain; rr *= gain; }
                        if (pb + r >= prm.p_begin && pb + r < prm.p_end) { o[off + r] = l; o[prm.out_stride + off + r] = rr; }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < kBlk; ++r) acc[r] = 0ull;
            if (MIX) {
#pragma unroll
                for (int r = 0; r < (MIX ? kBlk : 1); ++r) mixacc[r] = 0ull;
            }
        }
    }
}

// Adds the two partial tiles of every group a span boundary split (fixed order: the earlier CTA's
// part first), applies the gain, takes the peak and stores.  grid = (boundaries, 2 ears, T / 1024).
__global__ void __launch_bounds__(256)
bas_render_fixup_kernel(RenderParams prm, SpanInfo sp, const float* __restrict__ workspace, long long G, int TW) {
    const long long c = blockIdx.x + 1;                           // boundary between CTA c-1 and CTA c
    const int ear = blockIdx.y;
    const long long ic = span_begin(sp, c, G);
    if (ic % sp.gs == 0 || ic >= sp.total) return;                // boundary on a group edge: nothing was split
    const long long grp = ic / sp.gs;
    const int T = TW * kWarpTile;
    const int src = prm.mix ? 0 : (int)(grp / prm.tiles);
    const long long tile = prm.mix ? grp : grp - (long long)src * prm.tiles;
    const long long p_base = prm.p_begin / kBlk * kBlk;
    const int i = blockIdx.z * kWarpTile + threadIdx.x * 4;       // 256 threads x 4 outputs
    const long long p = p_base + tile * T + i;
    const float4 a = *reinterpret_cast<const float4*>(workspace + ((c - 1) * 2 + 1) * (2LL * T) + (long long)ear * T + i);
    const float4 b = *reinterpret_cast<const float4*>(workspace + (c * 2 + 0) * (2LL * T) + (long long)ear * T + i);
    const float gain = (!prm.mix && prm.gains) ? prm.gains[src] : 1.f;
    float* o = prm.out + (prm.mix ? 0 : (long long)src * 2 * prm.out_stride) + (long long)ear * prm.out_stride;
    const float v[4] = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
    float pk = 0.f;
    const bool vec_ok = (prm.p_begin & 3) == 0 && (prm.out_stride & 3) == 0 && (reinterpret_cast<uintptr_t>(prm.out) & 15) == 0;
    if (vec_ok && p >= prm.p_begin && p + 4 <= prm.p_end) {
        pk = fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3])));
        *reinterpret_cast<float4*>(o + (p - prm.p_begin)) = make_float4(gain * v[0], gain * v[1], gain * v[2], gain * v[3]);
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (p + e >= prm.p_begin && p + e < prm.p_end) { pk = fmaxf(pk, fabsf(v[e])); o[p + e - prm.p_begin] = gain * v[e]; }
    }
    if (prm.peaks && !prm.mix) {
        pk = warp_max(pk);
        if ((threadIdx.x & 31) == 0 && pk > 0.f) atomic_max_nonneg(prm.peaks + src, pk);
    }
}

int device_sm_count() {
    static thread_local int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    return sm_count;
}

template <int TW, bool MIX>
int launch_tiled(RenderParams prm, bool want_split, float* workspace, long long workspace_bytes, cudaStream_t st) {
    const TileGeom g = tile_geom(prm.K, prm.C, prm.pitch, TW);
    const size_t smem = tile_smem_bytes(g, TW);
    if (smem > 227 * 1024) return BAS_E_UNSUPPORTED;
    auto kern = bas_render_tiled_kernel<TW, MIX>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { bas_set_error("bas_render: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    const long long p_base = prm.p_begin / kBlk * kBlk;
    prm.tiles = bas_ceil_div(prm.p_end - p_base, (long long)TW * kWarpTile);
    SpanInfo sp;
    sp.gs = MIX ? prm.n_src : g.D;
    sp.n_groups = MIX ? prm.tiles : prm.tiles * prm.n_src;
    sp.total = sp.n_groups * sp.gs;
    // persistent grid: as many CTAs as the device keeps resident
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TW * 32, smem);
    if (e != cudaSuccess || per_sm < 1) { bas_set_error("bas_render: tile shape does not fit an SM"); cudaGetLastError(); return BAS_E_UNSUPPORTED; }
    long long grid = (long long)device_sm_count() * per_sm;
    if (grid > sp.n_groups) grid = sp.n_groups;
    // split groups between CTAs only when every span is longer than a group (then a group has at
    // most two contributors) and the caller gave a workspace
    const long long need = grid * 2 * (2LL * TW * kWarpTile) * 4;
    sp.split = (want_split && workspace && workspace_bytes >= need && grid > 1 && sp.total / grid >= sp.gs + 1) ? 1 : 0;
    kern<<<(unsigned)grid, TW * 32, smem, st>>>(prm, sp, workspace);
    e = cudaGetLastError();
    if (e != cudaSuccess) { bas_set_error("bas_render: tiled launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    if (sp.split) {
        dim3 fgrid((unsigned)(grid - 1), 2, (unsigned)TW);
        bas_render_fixup_kernel<<<fgrid, 256, 0, st>>>(prm, sp, workspace, grid, TW);
        e = cudaGetLastError();
        if (e != cudaSuccess) { bas_set_error("bas_render: fix-up launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    }
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

// variant encoding beyond the public three: BAS_RENDER_TILED | (TW << 8) picks the warps per CTA
// (tuning sweeps; unknown shapes return BAS_E_UNSUPPORTED).
extern "C" int bas_render(const float* x_dev, long long x_stride, long long n_valid, int n_src, long long n_in,
                          int C, int S, int K, const float* filt_dev, const float* gains_dev,
                          long long p_begin, long long p_count, float* out_dev, long long out_stride, int mix,
                          float* peaks_dev, int variant, void* workspace_dev, long long workspace_bytes, void* stream) {
    BAS_CHECK_ARG(x_dev && filt_dev && out_dev, "null pointer");
    BAS_CHECK_ARG(n_src >= 1, "n_src");
    BAS_CHECK_ARG(C >= 1 && S >= 1 && C % S == 0, "subchunksize must divide chunksize");     // apply_hrtf.py:401-402
    BAS_CHECK_ARG(K >= 1 && K < (1 << 20), "K");
    BAS_CHECK_ARG(n_in >= C && n_in % C == 0, "n_in must be a positive multiple of C");      // apply_hrtf.py:405
    BAS_CHECK_ARG(n_valid >= 0 && n_valid <= n_in && (n_src == 1 || x_stride >= n_valid), "n_valid / x_stride");
    BAS_CHECK_ARG(p_begin >= 0 && p_count >= 0 && p_begin + p_count <= n_in + K - 1, "output range");
    BAS_CHECK_ARG(out_stride >= p_count, "out_stride < p_count");
    BAS_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace_dev) & 15) == 0 && workspace_bytes >= 0, "workspace");
    if (p_count == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    RenderParams prm;
    prm.x = x_dev; prm.x_stride = x_stride; prm.n_valid = n_valid; prm.n_src = n_src; prm.n_in = n_in;
    prm.C = C; prm.S = S; prm.K = K; prm.filt = reinterpret_cast<const float2*>(filt_dev);
    prm.pitch = bas_filter_row_pitch(K);
    prm.filt_src_stride = (n_in / C + 1) * (long long)prm.pitch;
    prm.gains = gains_dev; prm.p_begin = p_begin; prm.p_end = p_begin + p_count;
    prm.out = out_dev; prm.out_stride = out_stride; prm.mix = mix ? 1 : 0; prm.peaks = peaks_dev; prm.tiles = 0;

    const int base = variant & 0x3f;
    BAS_CHECK_ARG(base == BAS_RENDER_AUTO || base == BAS_RENDER_GENERIC || base == BAS_RENDER_TILED, "variant");
    const bool tiled_ok = S == kBlk && C % kBlk == 0 && n_valid % 4 == 0 && (reinterpret_cast<uintptr_t>(x_dev) & 15) == 0 &&
                          (reinterpret_cast<uintptr_t>(filt_dev) & 15) == 0 && (n_src == 1 || x_stride % 4 == 0);
    if (base == BAS_RENDER_TILED && !tiled_ok) {
        bas_set_error("bas_render: tiled kernel needs S == 32, 32 | C, 4 | n_valid, 16-byte aligned signals and filter rows");
        return BAS_E_UNSUPPORTED;
    }
    if (base != BAS_RENDER_GENERIC && tiled_ok) {
        const int tw_req = (variant >> 8) & 0xff;
        int rc = BAS_E_UNSUPPORTED;
        // default: 4 warps per CTA, falling back to smaller tiles when shared memory runs out
        const int order[3] = {tw_req ? tw_req : (prm.mix ? 8 : 4), tw_req ? 0 : 2, tw_req ? 0 : 1};
        for (int i = 0; i < 3 && rc == BAS_E_UNSUPPORTED; ++i) {
            const int tw = order[i];
            // Splitting tiles between CTAs pays when a tile carries many sources (mixing); with one
            // source per tile whole tiles dealt round-robin measured faster (fewer, longer items).
            float* ws = reinterpret_cast<float*>(workspace_dev);
            const bool split = (variant & BAS_RENDER_SPLIT) || (prm.mix && !(variant & BAS_RENDER_NO_SPLIT));
            if (tw == 8) rc = prm.mix ? launch_tiled<8, true>(prm, split, ws, workspace_bytes, st) : launch_tiled<8, false>(prm, split, ws, workspace_bytes, st);
            else if (tw == 4) rc = prm.mix ? launch_tiled<4, true>(prm, split, ws, workspace_bytes, st) : launch_tiled<4, false>(prm, split, ws, workspace_bytes, st);
            else if (tw == 2) rc = prm.mix ? launch_tiled<2, true>(prm, split, ws, workspace_bytes, st) : launch_tiled<2, false>(prm, split, ws, workspace_bytes, st);
            else if (tw == 1) rc = prm.mix ? launch_tiled<1, true>(prm, split, ws, workspace_bytes, st) : launch_tiled<1, false>(prm, split, ws, workspace_bytes, st);
        }
        if (rc != BAS_E_UNSUPPORTED || base == BAS_RENDER_TILED) {
            if (rc == BAS_E_UNSUPPORTED) bas_set_error("bas_render: no tile shape fits K=%d C=%d (requested TW=%d)", K, C, tw_req);
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

extern "C" long long bas_render_workspace_bytes(void) {
    // resident warps x 2 slots x (1024 outputs x 2 ears) floats, whatever the tile shape
    return (long long)device_sm_count() * 8 * 2 * (2LL * kWarpTile) * 4;
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
