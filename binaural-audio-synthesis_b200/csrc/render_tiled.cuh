// Shared by render.cu (API, generic kernel) and render_tw*.cu (instantiations of the tiled kernel,
// one translation unit per tile width so they compile in parallel).
#pragma once
#include "bas_internal.cuh"

namespace bas_render_detail {

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
__host__ __device__ inline size_t tile_smem_bytes(const TileGeom& g, int TW, int NS) {
    return 64 + (size_t)NS * g.stage_bytes + (size_t)TW * g.warp_x_bytes;
}
constexpr size_t kBarBytes = 64;         // full + empty mbarriers of up to 2 stages, padded

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

template <int TW, bool MIX, int NS>
__global__ void __launch_bounds__(TW * 32, 1)
bas_render_tiled_kernel(RenderParams prm, SpanInfo sp, float* __restrict__ workspace) {
    extern __shared__ __align__(128) unsigned char smem[];
    const TileGeom g = tile_geom(prm.K, prm.C, prm.pitch, TW);
    u64* full_bar = reinterpret_cast<u64*>(smem);              // [NS]
    u64* empty_bar = full_bar + NS;                       // [NS]
    unsigned char* stage_base = smem + kBarBytes;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* xw = reinterpret_cast<float*>(stage_base + (size_t)NS * g.stage_bytes + (size_t)warp * g.warp_x_bytes);
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
        for (int s = 0; s < NS; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, TW); }
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

    // ---- producer: warp 0 stages item j into ring slot j % NS --------------------------------
    auto produce = [&](long long j) {
        const int st = (int)(j % NS);
        const long long use = j / NS;
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
        // two stages: the copy of item j+1 overlaps the arithmetic of item j.  One stage: the slot can
        // only be refilled after every warp (this one included) has left it - see below.
        if (NS > 1 && warp == 0 && j + 1 < n_items) produce(j + 1);
        const int st = (int)(j % NS);
        const Item it = item_info(j);
        long long n_lo, c_first; int n_rows;
        tile_chunks(it.tile, n_lo, c_first, n_rows);
        const long long P0 = p_base + it.tile * (TW * kWarpTile);
        const bool warp_live = P0 + (long long)warp * kWarpTile < prm.p_end;
        const float* xs = reinterpret_cast<const float*>(stage_base + (size_t)st * g.stage_bytes);
        const float2* fs = reinterpret_cast<const float2*>(stage_base + (size_t)st * g.stage_bytes + g.x_bytes);
        const long long q0 = n_lo / kBlk;                       // exact (n_lo % 32 == 0), may be negative

        mbar_wait(full_bar + st, (unsigned)((j / NS) & 1));

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
        if (NS == 1 && warp == 0 && j + 1 < n_items) produce(j + 1);

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
static __global__ void __launch_bounds__(256)
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

inline int device_sm_count() {
    static thread_local int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    return sm_count;
}

// Resident warps per SM this shape reaches (0: does not fit).
template <int TW, bool MIX, int NS>
int tiled_warps_per_sm(int K, int C, int pitch) {
    const TileGeom g = tile_geom(K, C, pitch, TW);
    const size_t smem = tile_smem_bytes(g, TW, NS);
    if (smem > 227 * 1024) return 0;
    auto kern = bas_render_tiled_kernel<TW, MIX, NS>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TW * 32, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    return per_sm * TW;
}

template <int TW, bool MIX, int NS>
int launch_tiled(RenderParams prm, bool want_split, float* workspace, long long workspace_bytes, cudaStream_t st) {
    const TileGeom g = tile_geom(prm.K, prm.C, prm.pitch, TW);
    const size_t smem = tile_smem_bytes(g, TW, NS);
    if (smem > 227 * 1024) return BAS_E_UNSUPPORTED;
    auto kern = bas_render_tiled_kernel<TW, MIX, NS>;
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


// explicit instantiations live in render_tw2.cu / render_tw4.cu / render_tw8.cu
#define BAS_DECLARE_TILED(TW_)                                                                              \
    extern template int launch_tiled<TW_, false, 1>(RenderParams, bool, float*, long long, cudaStream_t);    \
    extern template int launch_tiled<TW_, false, 2>(RenderParams, bool, float*, long long, cudaStream_t);    \
    extern template int launch_tiled<TW_, true, 1>(RenderParams, bool, float*, long long, cudaStream_t);     \
    extern template int launch_tiled<TW_, true, 2>(RenderParams, bool, float*, long long, cudaStream_t);     \
    extern template int tiled_warps_per_sm<TW_, false, 1>(int, int, int);                                    \
    extern template int tiled_warps_per_sm<TW_, false, 2>(int, int, int);                                    \
    extern template int tiled_warps_per_sm<TW_, true, 1>(int, int, int);                                     \
    extern template int tiled_warps_per_sm<TW_, true, 2>(int, int, int);
#define BAS_INSTANTIATE_TILED(TW_)                                                                    \
    template int launch_tiled<TW_, false, 1>(RenderParams, bool, float*, long long, cudaStream_t);    \
    template int launch_tiled<TW_, false, 2>(RenderParams, bool, float*, long long, cudaStream_t);    \
    template int launch_tiled<TW_, true, 1>(RenderParams, bool, float*, long long, cudaStream_t);     \
    template int launch_tiled<TW_, true, 2>(RenderParams, bool, float*, long long, cudaStream_t);     \
    template int tiled_warps_per_sm<TW_, false, 1>(int, int, int);                                    \
    template int tiled_warps_per_sm<TW_, false, 2>(int, int, int);                                    \
    template int tiled_warps_per_sm<TW_, true, 1>(int, int, int);                                     \
    template int tiled_warps_per_sm<TW_, true, 2>(int, int, int);

}  // namespace bas_render_detail
