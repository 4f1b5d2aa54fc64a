// Shared by render.cu (API, generic kernel) and render_tw*.cu (instantiations of the tiled kernel,
// one translation unit per tile width so they compile in parallel).
#pragma once
#include "bas_internal.cuh"

#include <cuda.h>

namespace bas_render_detail {

struct TermDev { int32_t row_shift; float weight; };      // bas_term

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
    int accumulate;                // mix only: add to what `out` already holds (source groups rendered in turn)
    float* peaks;
    long long tiles;               // tiled kernel: tiles per source
    int parts;                     // tiled kernel: warps that share one 1024-output stripe (split along the taps)
    int tmap;                      // tiled kernel: input rows arrive by tensor-map TMA (128-byte swizzle), else bulk copy + re-layout
    int box_rows, n_box;           // tensor-map path: rows per copy, copies per item
    // fused filter synthesis (FUSED kernels): the filter rows of an item are not copied from `filt` but
    // synthesised by the CTA itself from the plan terms and the L2-resident polyphase bank
    const TermDev* terms;          // [n_src][n_in/C + 1][2 ears][16] (bas_plan_build)
    const float* bank2;            // polyphase bank with every phase row stored twice: [ear][row][U][2K] (+ padding)
    int U;
    int nf;                        // FUSED: filter-row buffers in shared memory (2: producers run one item ahead)
    // routed mix (MIX kernels, several GPUs, peer.cu): a finished tile of the local mix is not stored to `out` but
    // into the receive buffer of the rank that owns its stretch of the output, over NVLink:
    //   route_table[owner] + (route_rank * 2 + ear) * route_stride + (p - owner * route_len)
    float* const* route_table;     // device array of route_n peer-mapped receive buffers, or NULL
    int route_n, route_rank;
    long long route_len, route_stride;
    // ... and, when arrive_ptrs is set, the last CTA of the launch to finish tells every rank "the tiles of rank
    // route_rank have landed" (flag[route_rank] = arrive_epoch on every peer; bas_peer_signal folded into the render)
    unsigned* const* arrive_ptrs;  // device array of route_n peer-mapped flag arrays, or NULL
    unsigned* arrive_counter;      // CTAs of this launch that have finished (zero between launches)
    unsigned arrive_epoch;
    // diagnostics (bas_render_set_trace; NULL in normal use): per CTA {start ns, end ns, SM id, items}
    unsigned long long* trace;
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
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
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
__device__ __forceinline__ bool mbar_try_wait(void* bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Waits are bounded by ELAPSED TIME, not by a spin count: a copy that never lands (a wrong byte count, a
// bad pointer) or a neighbour that never publishes must fail loudly instead of hanging the GPU, but a
// healthy launch that is merely time-sliced (MPS, ncu replay, compute-sanitizer) or whose neighbour CTA
// is not resident yet (other streams hold the SMs) must not be killed.
constexpr unsigned long long kWaitLimitNs = 30ull * 1000 * 1000 * 1000;
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    if (mbar_try_wait(bar, parity)) return;
    const unsigned long long t0 = global_ns();
    for (unsigned spins = 1; !mbar_try_wait(bar, parity); ++spins)
        if ((spins & 1023u) == 0 && global_ns() - t0 > kWaitLimitNs) __trap();
}
// spin (lane 0 of a warp) until the neighbour CTA has published its partial stripe for this launch
__device__ __forceinline__ void flag_wait(const unsigned long long* flag, unsigned long long epoch);
// TMA 1-D bulk copy global -> shared, completion counted in bytes on an mbarrier (UBLKCP in SASS)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// TMA tiled copy of a (32 floats x box_rows x 1) box of the input tensor [source][row][32]; rows outside
// the signal (negative or past the end) arrive as zeros; 128-byte swizzle (UTMALDG in SASS)
__device__ __forceinline__ void tensor_g2s(void* dst_smem, const CUtensorMap* map, int row, int src, void* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst_smem)), "l"(map), "r"(0), "r"(row), "r"(src), "r"(smem_u32(bar)) : "memory");
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

// TS = stripes (1024 outputs each) per tile = warps per CTA / warps per stripe
__host__ __device__ inline TileGeom tile_geom(int K, int C, int pitch, int TS) {
    TileGeom g;
    g.D = (K + kBlk - 1) / kBlk;
    g.x_rows = TS * 32 + g.D;
    g.f_rows = (g.x_rows * kBlk + C - 1) / C + 2;
    g.w_rows = 32 + g.D;
    g.x_bytes = (unsigned)g.x_rows * kBlk * 4;            // staged linearly (one bulk copy)
    g.f_bytes = (unsigned)g.f_rows * pitch * 8;
    g.stage_bytes = g.x_bytes + g.f_bytes;
    g.warp_x_bytes = (unsigned)g.w_rows * kXPitch * 4;    // per-warp copy on the conflict-free pitch
    return g;
}
constexpr size_t kBarBytes = 64;         // full + empty mbarriers of up to 2 stages, padded
constexpr size_t kStripeBytes = (size_t)kWarpTile * 8;      // one stripe of {L,R} partial sums
// barriers | NS stages | per-warp input rows | blend weights per subchunk | partial sums of the warps
// that share a stripe (parts > 1) | mix accumulators (MIX, one stripe per part-0 warp)
// tensor-map path: the rows of an item land densely (128-byte pitch, swizzled) in n_box copies of box_rows rows
__host__ __device__ inline int tmap_n_box(int x_rows) { return (x_rows + 255) / 256; }
__host__ __device__ inline int tmap_box_rows(int x_rows) { const int n = tmap_n_box(x_rows); return (x_rows + n - 1) / n; }
__host__ __device__ inline size_t tmap_stage_bytes(const TileGeom& g) {
    const size_t x = (size_t)tmap_n_box(g.x_rows) * tmap_box_rows(g.x_rows) * 128;
    return (x + g.f_bytes + 1023) / 1024 * 1024;                  // every stage starts on a swizzle atom
}
constexpr int kTermsPerRow = 2 * BAS_MAX_TERMS;             // both ears
// FUSED: a plan term as the producer warps use it - where its K floats start in the doubled bank (tap 0), and its weight
struct __align__(16) TermPtr { const float* p; float w; int pad; };
// FUSED kernels keep the input rows (staged by TMA) and the filter rows (synthesised by the producer warps)
// apart: NS stages of input rows, then nf buffers of filter rows and one term table
__host__ __device__ inline size_t fused_x_stage_bytes(const TileGeom& g, bool tmap) {
    if (!tmap) return (size_t)g.x_bytes;
    return ((size_t)tmap_n_box(g.x_rows) * tmap_box_rows(g.x_rows) * 128 + 1023) / 1024 * 1024;
}
__host__ __device__ inline size_t tile_smem_bytes(const TileGeom& g, int TW, int NS, int parts, int C, bool mix, bool tmap, bool fused = false, int nf = 0) {
    const int TS = TW / parts;
    size_t staging;
    if (fused) staging = (tmap ? 1024 : (size_t)TW * g.warp_x_bytes) + (size_t)NS * fused_x_stage_bytes(g, tmap);
    else staging = tmap ? 1024 + (size_t)NS * tmap_stage_bytes(g) : (size_t)NS * g.stage_bytes + (size_t)TW * g.warp_x_bytes;
    return kBarBytes + staging + (size_t)((C / kBlk * 8 + 15) / 16 * 16) +
           (fused ? (size_t)g.f_rows * kTermsPerRow * sizeof(TermPtr) + (size_t)nf * g.f_bytes : 0) +
           (parts > 1 ? (size_t)(TW - TS) * kStripeBytes : 0) + (mix ? (size_t)TS * kStripeBytes : 0);
}
// FUSED shapes: TW consumer warps + TW producer warps (whole warpgroups of four, as setmaxnreg wants), two input
// stages, 16 warps per SM launched at 128 registers and re-balanced to 168 (consumers) / 88 (producers):
// (4, 2, 2) and (8, 2, 1)
// subchunksize 16 (SUBS = 2) is compiled for two shapes only (they are fused shapes as well)
__host__ __device__ constexpr bool subs_shape_ok(int TW, int NS, int MINB) {
    return NS == 2 && ((TW == 4 && MINB == 2) || (TW == 8 && MINB == 1));
}
__host__ __device__ constexpr bool fused_shape_ok(int TW, int NS, int MINB) {
    return NS == 2 && ((TW == 4 && MINB == 2) || (TW == 8 && MINB == 1));
}

// One 32x32 block, visited diagonal by diagonal:  acc[r] += x_sel[m] * tap(r - m)  for r, m = 0..31 with
//   tap(j) = tap (base_a + j) of the blend of rows (ra, ra + pitch), x from xa,   for j = r - m >= 0
//   tap(j) = tap (base_b + j) of the blend of rows (rb, rb + pitch), x from xb,   for j = r - m <  0
// ra/rb point at tap base_a / base_b of the lane's filter row (float2 {L,R} entries, 16-byte aligned).
// For a full block (a, b) are the same filter and base_b = base_a; for the folded first/last block
// a is the d = 0 filter (base 0) and b the d = D filter (base 32 D).
// One blended tap pair {L,R} is live at a time and feeds the 32 - |j| packed FMAs of its diagonal
// (register-reuse operand); the lane's 32 input samples sit in registers as the scalar-broadcast
// operand.  No tap ring: 64 accumulator + 32 sample registers, and the loads and blends of the next
// diagonal overlap the FMAs of the current one, so a block has no serial prologue.
// SUBS subchunks per 32-sample input row (subchunksize 32 / SUBS): the blend weight belongs to the INPUT
// sample's subchunk (apply_hrtf.py:438-443), so a diagonal's tap is blended once per subchunk of the row and
// input sample m multiplies the blend of subchunk m / (32 / SUBS).  SUBS = 1: subchunksize 32 or a multiple.
template <int SUBS>
__device__ __forceinline__ void block_diag(u64 (&acc)[kBlk], const float2* __restrict__ ra, const u64 (&alpha_a)[SUBS],
                                           const float2* __restrict__ rb, const u64 (&alpha_b)[SUBS], int pitch,
                                           const float* __restrict__ xa, int ka, const float* __restrict__ xb, int kb) {
    constexpr int SL = kBlk / SUBS;              // input samples per subchunk of a row
    // xa / xb: the lane's input row; 16-byte chunk m4 of a row sits at chunk (m4 ^ key): key = row & 7 in the
    // swizzled tensor-map layout, 0 on the padded pitch
    float x[kBlk];
#pragma unroll
    for (int m4 = 0; m4 < kBlk / 4; ++m4) {
        const float4 v = *reinterpret_cast<const float4*>(xa + 4 * (m4 ^ ka));
        x[4 * m4] = v.x; x[4 * m4 + 1] = v.y; x[4 * m4 + 2] = v.z; x[4 * m4 + 3] = v.w;
    }
#pragma unroll
    for (int j = 0; j < kBlk; j += 2) {          // diagonals j, j + 1 >= 0: taps (base_a + j, base_a + j + 1) in one 16-byte load
        const ulonglong2 h0 = *reinterpret_cast<const ulonglong2*>(ra + j);
        const ulonglong2 h1 = *reinterpret_cast<const ulonglong2*>(ra + pitch + j);
        const u64 d0 = sub2(h1.x, h0.x), d1 = sub2(h1.y, h0.y);
        u64 w0[SUBS], w1[SUBS];
#pragma unroll
        for (int s = 0; s < SUBS; ++s) {
            w0[s] = fma2(alpha_a[s], d0, h0.x);                        // H_i + alpha (H_{i+1} - H_i)   apply_hrtf.py:443
            w1[s] = fma2(alpha_a[s], d1, h0.y);
        }
#pragma unroll
        for (int m = 0; m + j < kBlk; ++m) fma2_acc(acc[m + j], pack2(x[m], x[m]), w0[m / SL]);
#pragma unroll
        for (int m = 0; m + j + 1 < kBlk; ++m) fma2_acc(acc[m + j + 1], pack2(x[m], x[m]), w1[m / SL]);
    }
#pragma unroll
    for (int m4 = 0; m4 < kBlk / 4; ++m4) {      // folded block: the d = D part reads another input row
        const float4 v = *reinterpret_cast<const float4*>(xb + 4 * (m4 ^ kb));
        x[4 * m4] = v.x; x[4 * m4 + 1] = v.y; x[4 * m4 + 2] = v.z; x[4 * m4 + 3] = v.w;
    }
#pragma unroll
    for (int i = 1; i < kBlk; i += 2) {          // diagonals -i, -(i + 1): taps (base_b - i - 1, base_b - i) in one 16-byte load
        const ulonglong2 h0 = *reinterpret_cast<const ulonglong2*>(rb - i - 1);
        const ulonglong2 h1 = *reinterpret_cast<const ulonglong2*>(rb + pitch - i - 1);
        const u64 d0 = sub2(h1.x, h0.x), d1 = sub2(h1.y, h0.y);
        u64 w1[SUBS];
#pragma unroll
        for (int s = 0; s < SUBS; ++s) w1[s] = fma2(alpha_b[s], d1, h0.y);           // tap base_b - i
#pragma unroll
        for (int m = i; m < kBlk; ++m) fma2_acc(acc[m - i], pack2(x[m], x[m]), w1[m / SL]);
        if (i + 1 < kBlk) {
            u64 w0[SUBS];
#pragma unroll
            for (int s = 0; s < SUBS; ++s) w0[s] = fma2(alpha_b[s], d0, h0.x);       // tap base_b - i - 1
#pragma unroll
            for (int m = i + 1; m < kBlk; ++m) fma2_acc(acc[m - i - 1], pack2(x[m], x[m]), w0[m / SL]);
        }
    }
}

// ---- work decomposition (stream-K) -----------------------------------------------------------------
// The work of a launch is a line of SLICES grouped into GROUPS of gs slices that share one output
// tile:   one source per tile (MIX = false): group = (source, tile), slice = tap block d, gs = D
//         mixing                (MIX = true): group = tile,           slice = source,      gs = n_src
// CTA c owns the contiguous slice span [c*total/G, (c+1)*total/G).  A span boundary that falls
// inside a group splits that group between exactly two CTAs (spans are longer than a group).  The
// later CTA runs its share FIRST in its span, writes the partial stripe to the workspace and raises a
// flag (st.release.gpu); the earlier CTA reaches the group LAST in its span, waits for the flag
// (ld.acquire.gpu), adds the partial to its own sums - earlier part + later part, a fixed order - and
// finishes the tile.  CTA c only ever waits for CTA c + 1, and for work c + 1 does before anything
// else, so the wait cannot deadlock even when not all CTAs are resident at once.  Every scheduler
// gets the same number of 32x32 blocks and results stay deterministic.  Without a workspace spans
// are rounded to group boundaries.
// CHAIN kernels (mixing only): a launch with fewer tiles than CTAs (a time range of a long mix: 39 tiles x 64 sources)
// still fills the device - spans may be SHORTER than a group, down to 1 / kMaxChain of it.  A CTA in the middle of a
// tile runs its sources, waits for its successor's partial sums, adds them to its own and publishes the total to its
// predecessor: the sums of a tile travel down a chain of at most kMaxChain + 1 CTAs (all resident: the grid is never
// larger than what the device holds), later sources first, in a fixed order.  A kernel instance of its own, so that
// the code of the common case stays what it was.
constexpr int kMaxChain = 8;
struct SpanInfo { long long total, n_groups; int gs; int split; unsigned long long epoch; };

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
    bool goes_on;           // CHAIN: the group continues in the next CTA (also when it began in an earlier one)
};

__device__ __forceinline__ void flag_release(unsigned long long* flag, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(flag), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long flag_acquire(const unsigned long long* flag) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    return v;
}
__device__ __forceinline__ void flag_wait(const unsigned long long* flag, unsigned long long epoch) {
    if (flag_acquire(flag) == epoch) return;
    const unsigned long long t0 = global_ns();
    for (unsigned spins = 1; flag_acquire(flag) != epoch; ++spins) {
        __nanosleep(64);
        if ((spins & 1023u) == 0 && global_ns() - t0 > kWaitLimitNs) __trap();     // the neighbour never published
    }
}
// workspace: one flag and one stripe of {L,R} partial sums per (CTA, stripe)
__host__ __device__ inline size_t ws_flag_bytes(long long grid, int TS) { return ((size_t)grid * TS * 8 + 255) / 256 * 256; }
__host__ __device__ inline size_t ws_bytes(long long grid, int TS) { return ws_flag_bytes(grid, TS) + (size_t)grid * TS * kWarpTile * 8; }

__device__ __forceinline__ void cta_barrier(int threads) {
    asm volatile("bar.sync 1, %0;" ::"r"(threads) : "memory");
}

// TW warps per CTA, NS pipeline stages, MINB resident CTAs per SM the register budget is cut for.
// prm.parts (a divisor of TW) warps share each 1024-output stripe of a tile, each running its share of
// the item's tap blocks; their partial sums meet in shared memory in a fixed order.  A tile is then
// TW / parts stripes: smaller tiles and more of them, which is what fills the last wave of a launch
// whose tile count is a small multiple of the resident warps.
// FUSED: the boundary-filter rows an item needs are not copied from a filter-row array that a separate
// bas_ir_synth launch wrote; TW / 2 PRODUCER warps of the CTA synthesise them straight into shared memory
// (interpolate_2d's array part, apply_hrtf.py:219-281, = ir_synth.cu's weighted gather from the L2-resident
// polyphase bank) one item ahead of the TW CONSUMER warps, which run the FIR blocks.  The gathers lean on L2
// while the FMA pipe idles, the FIR blocks are the exact complement: the two share every SM instead of taking
// turns on the device.  Hand-over by mbarriers (full / empty per filter buffer and per input stage); the
// producers also issue the TMA copies of the input rows.  Same terms, same order of summation: the rows
// are bit-identical to bas_ir_synth's.
// SUBS: subchunks per 32-sample row (see block_diag): 1 for subchunksize 32, 64, 96, ...; 2 for subchunksize 16.
template <int TW, bool MIX, int NS, int MINB, bool FUSED = false, int SUBS = 1, bool CHAIN = false>
__global__ void __launch_bounds__((FUSED ? 2 * TW : TW) * 32, MINB)
bas_render_tiled_kernel(RenderParams prm, SpanInfo sp, float* __restrict__ workspace, const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int P = prm.parts, TS = TW / P;
    const TileGeom g = tile_geom(prm.K, prm.C, prm.pitch, TS);
    constexpr int PW = FUSED ? TW : 0;                    // producer warps
    u64* full_bar = reinterpret_cast<u64*>(smem);              // [NS]
    u64* empty_bar = full_bar + NS;                       // [NS]
    u64* full_f = empty_bar + NS;                         // [2]  FUSED: filter buffer written / ...
    u64* empty_f = full_f + 2;                            // [2]  ... read by every consumer warp
    // tensor-map path: stages start on a 1024-byte boundary (the 128-byte swizzle works on address bits)
    const bool tmap = prm.tmap != 0;
    unsigned char* stage_base = smem + kBarBytes;
    if (tmap) stage_base += (1024u - (smem_u32(stage_base) & 1023u)) & 1023u;
    const size_t stage_stride = FUSED ? fused_x_stage_bytes(g, tmap) : tmap ? tmap_stage_bytes(g) : (size_t)g.stage_bytes;
    const size_t x_stage_bytes = tmap ? (size_t)prm.n_box * prm.box_rows * 128 : (size_t)g.x_bytes;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int stripe = warp / P, part = warp - stripe * P;     // warps of a stripe are neighbours
    const int spc = prm.C / kBlk;                              // subchunks per chunk
    unsigned char* after_stages = stage_base + (size_t)NS * stage_stride;
    float* xw = reinterpret_cast<float*>(after_stages + (size_t)warp * g.warp_x_bytes);          // unused on the tensor-map path
    float* alpha_tab = reinterpret_cast<float*>(after_stages + (tmap ? 0 : (size_t)TW * g.warp_x_bytes));
    unsigned char* after_tab = reinterpret_cast<unsigned char*>(alpha_tab) + (spc * 8 + 15) / 16 * 16;     // room for SUBS = 2
    TermPtr* term_tab = reinterpret_cast<TermPtr*>(after_tab);     // FUSED: {start of the term's taps in bank2, weight} per (row, ear, slot)
    unsigned char* fbuf_base = after_tab + (FUSED ? (size_t)g.f_rows * kTermsPerRow * sizeof(TermPtr) : 0);       // FUSED: nf filter-row buffers
    unsigned char* after_alpha = fbuf_base + (FUSED ? (size_t)prm.nf * g.f_bytes : 0);
    // partial sums of parts 1..P-1 of every stripe: [stripe][part - 1][r][lane] {L,R}
    u64* red = reinterpret_cast<u64*>(after_alpha);
    u64* mixbuf = reinterpret_cast<u64*>(after_alpha + (P > 1 ? (size_t)(TW - TS) * kStripeBytes : 0)) + (size_t)stripe * kWarpTile;
    const long long p_base = prm.p_begin / kBlk * kBlk;
    const int T = TS * kWarpTile;                              // outputs per tile
    const long long n_chunks = prm.n_in / prm.C;
    // split: contiguous slice span [i0, i1);  otherwise whole groups, dealt round-robin (group = c + k G)
    const long long i0 = sp.split ? span_begin(sp, blockIdx.x, gridDim.x) : 0;
    const long long i1 = sp.split ? span_begin(sp, blockIdx.x + 1, gridDim.x) : 0;
    const long long g_first = i0 / sp.gs;
    const long long my_groups = blockIdx.x < sp.n_groups ? (sp.n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int n_items = (int)(sp.split ? (i1 <= i0 ? 0 : (MIX ? i1 - i0 : (i1 - 1) / sp.gs - g_first + 1))
                                       : my_groups * (MIX ? sp.gs : 1));
    const unsigned first_slice = (unsigned)(i0 - g_first * sp.gs);   // split: offset of the span inside its first group

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, TW); }
        if (FUSED) for (int s = 0; s < 2; ++s) { mbar_init(full_f + s, PW * 32); mbar_init(empty_f + s, TW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // blend weight of the subchunk that starts j samples into a chunk: alpha = j / chunksize with j a multiple of
    // the subchunksize (apply_hrtf.py:438, :442); entry i stands for samples [i, i + 1) * 32 / SUBS of the chunk
    for (int s = tid; s < spc * SUBS; s += (TW + PW) * 32)
        alpha_tab[s] = (float)((s * (kBlk / SUBS)) / prm.S * prm.S) / (float)prm.C;
    __syncthreads();
    // programmatic dependent launch: everything above overlapped the tail of the previous kernel in the
    // stream; the filter rows / plan terms it wrote are read only from here on
    bas_grid_launch_dependents();
    bas_grid_dependency_wait();
    if (prm.trace != nullptr && tid == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        prm.trace[4 * blockIdx.x] = global_ns(); prm.trace[4 * blockIdx.x + 2] = smid;
    }

    // item j of this CTA (32-bit arithmetic: the host keeps slice counts below 2^31)
    auto item_info = [&](int j) {
        Item it;
        long long grp; int a, b;
        if (sp.split) {
            if (MIX) {
                const unsigned s_abs = first_slice + (unsigned)j;          // slices since the start of group g_first
                const unsigned gi = s_abs / (unsigned)sp.gs;
                grp = g_first + gi;
                it.src = (int)(s_abs - gi * (unsigned)sp.gs);
            } else {
                grp = g_first + j;
                it.src = 0;
            }
            const long long lo = grp * sp.gs;
            a = (int)((i0 > lo ? i0 : lo) - lo); b = (int)((i1 < lo + sp.gs ? i1 : lo + sp.gs) - lo);
        } else {
            const unsigned k = MIX ? (unsigned)j / (unsigned)sp.gs : (unsigned)j;
            grp = blockIdx.x + (long long)k * gridDim.x;
            a = 0; b = sp.gs;
            it.src = MIX ? (int)((unsigned)j - k * (unsigned)sp.gs) : 0;
        }
        it.partial = a > 0 || b < sp.gs;
        it.slot = a > 0 ? 0 : 1;
        it.goes_on = b < sp.gs;
        if (MIX) {
            it.tile = grp;
            it.d0 = 0; it.d1 = g.D;
            it.group_end = it.src == b - 1;
        } else {
            const unsigned ug = (unsigned)grp, ut = (unsigned)prm.tiles;
            it.src = (int)(ug / ut); it.tile = (long long)(ug - (unsigned)it.src * ut);
            it.d0 = a; it.d1 = b;
            it.group_end = true;
        }
        return it;
    };
    // chunk range whose filter rows a tile needs (clamped to the signal)
    auto tile_chunks = [&](long long tile, long long& n_lo, long long& c_first, int& n_rows) {
        const long long P0 = p_base + tile * T;
        n_lo = P0 - (long long)kBlk * g.D;
        if (n_lo < 0) c_first = 0;
        else if (n_lo < 0x7fffffffLL) c_first = (long long)((unsigned)n_lo / (unsigned)prm.C);
        else c_first = n_lo / prm.C;
        const long long p_last = P0 + T - 1;
        long long c_last = p_last < 0x7fffffffLL ? (long long)((unsigned)p_last / (unsigned)prm.C) : p_last / prm.C;
        if (c_last > n_chunks - 1) c_last = n_chunks - 1;
        if (c_first > c_last) c_first = c_last;
        n_rows = (int)(c_last - c_first + 2);                   // + the boundary after the last chunk
    };

    // ---- producer: warp 0 stages item j into ring slot j % NS --------------------------------
    auto produce = [&](int j) {
        const int st = j % NS;
        const int use = j / NS;
        if (use > 0) mbar_wait(empty_bar + st, (unsigned)((use - 1) & 1));      // consumers left the slot
        if (lane == 0) {
            const Item it = item_info(j);
            long long n_lo, c_first; int n_rows;
            tile_chunks(it.tile, n_lo, c_first, n_rows);
            float* xs = reinterpret_cast<float*>(stage_base + (size_t)st * stage_stride);
            float2* fs = reinterpret_cast<float2*>(stage_base + (size_t)st * stage_stride + x_stage_bytes);
            const float* x = prm.x + (long long)it.src * prm.x_stride;
            const unsigned f_bytes = FUSED ? 0u : (unsigned)n_rows * prm.pitch * 8;
            if (tmap) {
                // input rows: n_box tiled copies of box_rows rows each (rows outside the signal are zero-filled by
                // the TMA unit, and count towards the transaction bytes); filter rows: one bulk copy
                mbar_arrive_expect_tx(full_bar + st, (unsigned)x_stage_bytes + f_bytes);
                const int row0 = (int)(n_lo / kBlk);                           // exact, may be negative
                for (int b = 0; b < prm.n_box; ++b)
                    tensor_g2s(reinterpret_cast<unsigned char*>(xs) + (size_t)b * prm.box_rows * 128, &xmap, row0 + b * prm.box_rows, it.src, full_bar + st);
            } else {
                // two bulk copies per item: the in-range part of the input span, and the filter rows.
                // Samples outside [0, n_valid) are never copied; consumers zero them while re-laying out.
                const long long na = n_lo < 0 ? 0 : n_lo;
                long long nb = n_lo + (long long)g.x_rows * kBlk;
                if (nb > prm.n_valid) nb = prm.n_valid;
                const unsigned x_bytes = nb > na ? (unsigned)(nb - na) * 4 : 0;
                mbar_arrive_expect_tx(full_bar + st, x_bytes + f_bytes);
                if (x_bytes) bulk_g2s(xs + (na - n_lo), x + na, x_bytes, full_bar + st);
            }
            if (!FUSED) bulk_g2s(fs, prm.filt + (long long)it.src * prm.filt_src_stride + c_first * prm.pitch, f_bytes, full_bar + st);
        }
        __syncwarp();
    };

    if (FUSED) {
        // 16 warps per SM were launched with 128 registers each; the FIR blocks need 168, the gathers get by with 88
        if (warp >= TW) asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 168;");
    }
    if (FUSED && warp >= TW) {
        // =========================== producer warps (FUSED) ===========================================
        // For every item, in order: the TMA copies of its input rows into stage j % NS (first producer warp), then its
        // filter rows into buffer j % nf:   row[m] = sum_t w_t * bank[row_t][phase_t][(m - adv_t) mod K]   per ear,
        // the terms in their fixed plan slots (a slot with weight zero reads the bank's first phase row with weight
        // zero: adds exactly nothing), summed in slot order like ir_synth.cu.  Every phase row is stored twice in
        // a row, so (m - adv) mod K is the plain index m + K - adv.
        constexpr int PT = PW * 32;
        const int ptid = tid - TW * 32;
        const int K = prm.K, K2 = 2 * prm.K;
        for (int j = 0; j < n_items; ++j) {
            if (warp == TW) produce(j);
            const Item it = item_info(j);
            long long n_lo, c_first; int n_rows;
            tile_chunks(it.tile, n_lo, c_first, n_rows);
            const int fb = j % prm.nf;
            if (j >= prm.nf) mbar_wait(empty_f + fb, (unsigned)((j / prm.nf - 1) & 1));      // consumers left the buffer
            asm volatile("bar.sync 2, %0;" ::"r"(PT) : "memory");                            // ... and every producer the table
            const TermDev* tsrc = prm.terms + ((long long)it.src * (n_chunks + 1) + c_first) * kTermsPerRow;
            for (int e = ptid; e < n_rows * kTermsPerRow; e += PT) {
                const TermDev t = tsrc[e];
                const int ear = (e & (kTermsPerRow - 1)) / BAS_MAX_TERMS;
                const int row = t.row_shift >> 20, shift = t.row_shift & 0xFFFFF;
                const int ph = (prm.U - shift % prm.U) % prm.U;
                const int adv = (shift + ph) / prm.U;
                const int off = t.weight != 0.f ? ((ear * BAS_N_DIRECTIONS + row) * prm.U + ph) * K2 + K - adv : 0;
                TermPtr tp;
                tp.p = prm.bank2 + off; tp.w = t.weight; tp.pad = 0;
                term_tab[e] = tp;
            }
            asm volatile("bar.sync 2, %0;" ::"r"(PT) : "memory");
            float2* fsw = reinterpret_cast<float2*>(fbuf_base + (size_t)fb * g.f_bytes);
            // One ROW per producer warp at a time (rows dealt round-robin); a lane owns taps lane, lane + 32, ...
            // and takes them two at a time: the table entry of a term is read once for both, the second load is the
            // first address plus 128 bytes.  The terms are summed in slot order per ear, like ir_synth.cu.
            const int pw = warp - TW;
            for (int row = pw; row < n_rows; row += PW) {
                const TermPtr* tab = term_tab + row * kTermsPerRow;
                float2* dst = fsw + row * prm.pitch;
                // 64 loads in flight per lane; the second producer warp of the scheduler is in its FMA phase meanwhile.  Few
                // instructions per gathered word matter here: the producers share the schedulers and the FMA pipe with the FIR
                // blocks.  A lane takes FOUR taps of one ear at a time (tap0, +32, +64, +96): one table read gives a term's
                // ready pointer, the four loads are that address plus immediates, and two packed FMAs add the term to the
                // four taps (same fp32 rounding per lane as scalar FMAs, terms in slot order like ir_synth.cu).
                for (int tap0 = lane; tap0 < K; tap0 += 128) {
                    u64 acc[2][2];
#pragma unroll
                    for (int ear = 0; ear < 2; ++ear) {
                        const TermPtr* te = tab + ear * BAS_MAX_TERMS;
                        u64 v[BAS_MAX_TERMS][2];
#pragma unroll
                        for (int t = 0; t < BAS_MAX_TERMS; ++t) {
                            // past the end of the row the loads stay inside the bank's padding
                            const float* q = te[t].p + (size_t)tap0;
                            v[t][0] = pack2(__ldg(q), __ldg(q + 32));
                            v[t][1] = pack2(__ldg(q + 64), __ldg(q + 96));
                        }
                        u64 a0 = 0ull, a1 = 0ull;
#pragma unroll
                        for (int t = 0; t < BAS_MAX_TERMS; ++t) {
                            const float w = te[t].w;
                            const u64 ww = pack2(w, w);
                            fma2_acc(a0, ww, v[t][0]);
                            fma2_acc(a1, ww, v[t][1]);
                        }
                        acc[ear][0] = a0; acc[ear][1] = a1;
                    }
                    float l[4], r[4];
                    unpack2(acc[0][0], l[0], l[1]); unpack2(acc[0][1], l[2], l[3]);
                    unpack2(acc[1][0], r[0], r[1]); unpack2(acc[1][1], r[2], r[3]);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (tap0 + 32 * i < K) dst[tap0 + 32 * i] = make_float2(l[i], r[i]);
                }
            }
            const int pad = prm.pitch - K;                           // zero padding taps K .. pitch - 1 of every row
            for (int e = ptid; e < n_rows * pad; e += PT) fsw[(e / pad) * prm.pitch + K + e % pad] = make_float2(0.f, 0.f);
            mbar_arrive(full_f + fb);                                // every producer thread: its own stores are released
        }
        return;
    }
    if (!FUSED && warp == 0 && n_items > 0) produce(0);

    if (MIX && part == 0) {
#pragma unroll
        for (int r = 0; r < kBlk; ++r) mixbuf[r * 32 + lane] = 0ull;
    }
    const int blk = stripe * 32 + lane;                         // output block of this lane inside the tile
    const bool vec_ok = (prm.p_begin & 3) == 0 && (prm.out_stride & 3) == 0 &&
                        (reinterpret_cast<uintptr_t>(prm.out) & 15) == 0;

    for (int j = 0; j < n_items; ++j) {
        // two stages: the copy of item j+1 overlaps the arithmetic of item j.  One stage: the slot can
        // only be refilled after every warp (this one included) has left it - see below.
        if (!FUSED && NS > 1 && warp == 0 && j + 1 < n_items) produce(j + 1);
        const int st = j % NS;
        const Item it = item_info(j);
        long long n_lo, c_first; int n_rows;
        tile_chunks(it.tile, n_lo, c_first, n_rows);
        const long long P0 = p_base + it.tile * T;
        const bool warp_live = P0 + (long long)stripe * kWarpTile < prm.p_end;
        const float* xs = reinterpret_cast<const float*>(stage_base + (size_t)st * stage_stride);
        const int fb = FUSED ? j % prm.nf : 0;
        const float2* fs = FUSED ? reinterpret_cast<const float2*>(fbuf_base + (size_t)fb * g.f_bytes)
                                 : reinterpret_cast<const float2*>(stage_base + (size_t)st * stage_stride + x_stage_bytes);
        const long long q0 = n_lo / kBlk;                       // exact (n_lo % 32 == 0), may be negative
        // this warp's share of the item's tap blocks
        const int len = it.d1 - it.d0;
        const int d_first = it.d0 + (len * part) / P, d_last = it.d0 + (len * (part + 1)) / P;

        u64 acc[kBlk];
#pragma unroll
        for (int r = 0; r < kBlk; ++r) acc[r] = 0ull;

        mbar_wait(full_bar + st, (unsigned)((j / NS) & 1));
        if (FUSED) mbar_wait(full_f + fb, (unsigned)((j / prm.nf) & 1));

        if (warp_live && d_last > d_first) {
            if (!tmap) {
                // re-lay this warp's input rows from the linear staging buffer onto the 144-byte pitch
                // (lane-per-row reads below are then conflict free) and zero what lies outside the signal
                const long long n_w = n_lo + (long long)stripe * kWarpTile;
                const float4* lin = reinterpret_cast<const float4*>(xs) + stripe * (kWarpTile / 4);
                for (int idx = lane; idx < g.w_rows * 8; idx += 32) {
                    const int row = idx >> 3, ch = idx & 7;
                    const long long n = n_w + (long long)row * kBlk + ch * 4;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (n >= 0 && n + 4 <= prm.n_valid) v = lin[idx];
                    *reinterpret_cast<float4*>(xw + row * kXPitch + ch * 4) = v;
                }
                __syncwarp();
            }
            // Filter row (chunk) and blend weight of the lane's input row.  Input rows are visited in
            // descending order (xrow = blk + D - d), so (chunk, sub) is divided once per item and then
            // stepped; rows outside the staged range belong to input rows that are all zero (clamped).
            const int q_top = (int)q0 + blk + g.D;               // absolute subchunk of the d = 0 row
            const int cf = (int)c_first;
            auto split_q = [&](int q, int& chunk, int& sub) {
                if (q < 0) { chunk = 0; sub = 0; }
                else { chunk = (int)((unsigned)q / (unsigned)spc); sub = q - chunk * spc; }
            };
            auto row_of = [&](int chunk) {
                int ri = chunk - cf;
                ri = ri < 0 ? 0 : (ri > n_rows - 2 ? n_rows - 2 : ri);
                return fs + ri * prm.pitch;
            };
            int chunk, sub;
            split_q(q_top - d_first, chunk, sub);
            // d = 0 is the folded block: diagonals >= 0 from the d = 0 filter, diagonals < 0 from the
            // d = D filter.  One call site keeps the unrolled body in the instruction cache.
#pragma unroll 1
            for (int d = d_first; d < d_last; ++d) {
                const float2* ra = row_of(chunk) + kBlk * d;
                u64 aa[SUBS], ab[SUBS];
#pragma unroll
                for (int s = 0; s < SUBS; ++s) {
                    const float alpha = alpha_tab[sub * SUBS + s];
                    aa[s] = ab[s] = pack2(alpha, alpha);
                }
                const float2* rb = ra;
                const int xrow_a = blk + g.D - d;
                int xrow_b = xrow_a;
                if (d == 0) {
                    xrow_b = blk;
                    int chunk_b, sub_b;
                    split_q(q_top - g.D, chunk_b, sub_b);
                    rb = row_of(chunk_b) + kBlk * g.D;
#pragma unroll
                    for (int s = 0; s < SUBS; ++s) {
                        const float alpha_b = alpha_tab[sub_b * SUBS + s];
                        ab[s] = pack2(alpha_b, alpha_b);
                    }
                }
                // the lane's input rows: on the padded per-warp copy, or in place in the swizzled stage
                const float* pa = tmap ? xs + xrow_a * kBlk : xw + (xrow_a - stripe * 32) * kXPitch;
                const float* pb_ = tmap ? xs + xrow_b * kBlk : xw + (xrow_b - stripe * 32) * kXPitch;
                block_diag<SUBS>(acc, ra, aa, rb, ab, prm.pitch, pa, tmap ? (xrow_a & 7) : 0, pb_, tmap ? (xrow_b & 7) : 0);
                // next row down: q - 1
                if (q_top - d - 1 < 0) { chunk = 0; sub = 0; }
                else if (--sub < 0) { sub = spc - 1; --chunk; }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty_bar + st);             // this warp is done with the slot
        if (FUSED && lane == 0) mbar_arrive(empty_f + fb);      // ... and with the filter rows
        if (!FUSED && NS == 1 && warp == 0 && j + 1 < n_items) produce(j + 1);

        // ---- the warps of a stripe add up: parts 1.. hand their sums to part 0, fixed order ---------
        if (P > 1) {
            if (part > 0) {
                u64* dst = red + ((size_t)(stripe * (P - 1) + part - 1) * kBlk) * 32 + lane;
#pragma unroll
                for (int r = 0; r < kBlk; ++r) { dst[r * 32] = acc[r]; acc[r] = 0ull; }
            }
            cta_barrier(TW * 32);
            if (part == 0) {
                for (int pp = 0; pp < P - 1; ++pp) {
                    const u64* srcp = red + ((size_t)(stripe * (P - 1) + pp) * kBlk) * 32 + lane;
#pragma unroll
                    for (int r = 0; r < kBlk; ++r) acc[r] = add2(acc[r], srcp[r * 32]);
                }
            }
            cta_barrier(TW * 32);                               // red may be overwritten by the next item
            if (part > 0) continue;
        }

        // ---- peak, gain, mix, store ----------------------------------------------------------------
        const long long pb = P0 + (long long)blk * kBlk;        // first output of this lane
        const float gain = prm.gains ? prm.gains[it.src] : 1.f;
        // a tile split between two CTAs: the later CTA (slot 0) publishes its partial sums, the earlier
        // one (slot 1) adds them to its own and finishes the tile
        const bool publish = it.partial && it.slot == 0, adopt = it.partial && it.slot == 1;
        unsigned long long* ws_flags = reinterpret_cast<unsigned long long*>(workspace);
        u64* ws_sums = reinterpret_cast<u64*>(reinterpret_cast<unsigned char*>(workspace) + ws_flag_bytes(gridDim.x, TS));
        if (!MIX && adopt && it.group_end && warp_live) {
            const long long slot = (long long)(blockIdx.x + 1) * TS + stripe;
            if (lane == 0) flag_wait(ws_flags + slot, sp.epoch);
            __syncwarp();
            const u64* srcp = ws_sums + slot * kWarpTile + lane;
#pragma unroll
            for (int r = 0; r < kBlk; ++r) acc[r] = add2(acc[r], __ldcg(srcp + r * 32));
        }
        if (prm.peaks && (MIX || !publish)) {
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
            // gain-weighted running sum over the sources of the tile, kept in shared memory; the last
            // source of the group leaves the total in acc
            const u64 g2 = pack2(gain, gain);
            const bool first_src = it.src == 0 || j == 0;
#pragma unroll
            for (int r = 0; r < kBlk; ++r) {
                const u64 prev = first_src ? 0ull : mixbuf[r * 32 + lane];
                acc[r] = fma2(g2, acc[r], prev);
                if (!it.group_end) { mixbuf[r * 32 + lane] = acc[r]; acc[r] = 0ull; }
            }
            if (adopt && it.group_end && warp_live) {           // earlier sources (here) + later sources (next CTA)
                const long long slot = (long long)(blockIdx.x + 1) * TS + stripe;
                if (lane == 0) flag_wait(ws_flags + slot, sp.epoch);
                __syncwarp();
                const u64* srcp = ws_sums + slot * kWarpTile + lane;
#pragma unroll
                for (int r = 0; r < kBlk; ++r) acc[r] = add2(acc[r], __ldcg(srcp + r * 32));
            }
        }
        if (it.group_end) {
            if (publish) {
                // partial stripe -> workspace[cta][stripe][r][lane] {L,R}, no gain (one source per tile) / mixed
                if (warp_live) {
                    const long long slot = (long long)blockIdx.x * TS + stripe;
                    u64* dstp = ws_sums + slot * kWarpTile + lane;
                    if (CHAIN && it.goes_on) {                  // middle of a chain: own sources + everything after them
                        const long long next = (long long)(blockIdx.x + 1) * TS + stripe;
                        if (lane == 0) flag_wait(ws_flags + next, sp.epoch);
                        __syncwarp();
                        const u64* srcp = ws_sums + next * kWarpTile + lane;
#pragma unroll
                        for (int r = 0; r < kBlk; ++r) acc[r] = add2(acc[r], __ldcg(srcp + r * 32));
                    }
#pragma unroll
                    for (int r = 0; r < kBlk; ++r) __stcg(dstp + r * 32, acc[r]);
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) flag_release(ws_flags + slot, sp.epoch);
                }
            } else {
                float* o = prm.out + (MIX ? 0 : (long long)it.src * 2 * prm.out_stride);
                long long off = pb - prm.p_begin, ostride = prm.out_stride;
                bool vok = vec_ok;
                if (MIX && prm.route_table != nullptr && pb >= 0) {
                    const long long owner = pb / prm.route_len;             // 32-output blocks never straddle owners
                    if (owner < prm.route_n) {
                        o = prm.route_table[owner] + (long long)prm.route_rank * 2 * prm.route_stride;
                        off = pb - owner * prm.route_len;
                        ostride = prm.route_stride;
                        vok = true;                                         // receive buffers are 16-byte aligned, strides multiples of 4
                    }
                }
                if (vok && pb >= prm.p_begin && pb + kBlk <= prm.p_end) {
#pragma unroll
                    for (int r4 = 0; r4 < kBlk; r4 += 4) {
                        float l[4], rr[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            unpack2(acc[r4 + i], l[i], rr[i]);
                            if (!MIX) { l[i] *= gain; rr[i] *= gain; }
                        }
                        if (MIX && prm.accumulate) {            // earlier source groups first, then this one
                            const float4 ol = *reinterpret_cast<const float4*>(o + off + r4);
                            const float4 orr = *reinterpret_cast<const float4*>(o + ostride + off + r4);
                            l[0] = ol.x + l[0]; l[1] = ol.y + l[1]; l[2] = ol.z + l[2]; l[3] = ol.w + l[3];
                            rr[0] = orr.x + rr[0]; rr[1] = orr.y + rr[1]; rr[2] = orr.z + rr[2]; rr[3] = orr.w + rr[3];
                        }
                        *reinterpret_cast<float4*>(o + off + r4) = make_float4(l[0], l[1], l[2], l[3]);
                        *reinterpret_cast<float4*>(o + ostride + off + r4) = make_float4(rr[0], rr[1], rr[2], rr[3]);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < kBlk; ++r) {
                        float l, rr; unpack2(acc[r], l, rr);
                        if (!MIX) { l *= gain; rr *= gain; }
                        if (pb + r >= prm.p_begin && pb + r < prm.p_end) {
                            if (MIX && prm.accumulate) { l = o[off + r] + l; rr = o[ostride + off + r] + rr; }
                            o[off + r] = l; o[ostride + off + r] = rr;
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < kBlk; ++r) acc[r] = 0ull;
        }
    }
    if (prm.trace != nullptr && tid == 0) { prm.trace[4 * blockIdx.x + 1] = global_ns(); prm.trace[4 * blockIdx.x + 3] = (unsigned long long)n_items; }
    if (MIX && prm.arrive_ptrs != nullptr) {
        // routed mix: every routed store of this CTA is ordered before its count, every count before the last CTA's
        // signal - a peer that reads the flag (ld.acquire.sys, peer.cu) then finds all tiles of this launch in place
        __threadfence_system();
        cta_barrier(TW * 32);
        if (warp == 0) {
            unsigned last = 0;
            if (lane == 0) last = atomicAdd(prm.arrive_counter, 1u) == gridDim.x - 1 ? 1u : 0u;
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                if (lane == 0) *prm.arrive_counter = 0;              // ready for the next launch
                __threadfence_system();
                if (lane < prm.route_n)
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(prm.arrive_ptrs[lane] + prm.route_rank), "r"(prm.arrive_epoch) : "memory");
            }
        }
    }
}

// Tensor map of the input signals as [source][row][32 floats] with the 128-byte swizzle; false when the
// signal does not qualify (or the driver entry point is missing): the kernel then stages with bulk copies.
bool input_tensor_map_possible(const RenderParams& prm);
bool make_input_tensor_map(CUtensorMap* map, const RenderParams& prm, int box_rows);

// Launch stamp the hand-off flags of a split launch are compared with: unique per launch of this
// process (random salt in the high bits), so stale flags in the caller's workspace never match.
unsigned long long next_epoch();

inline int device_sm_count() {
    static thread_local int sm_count = 0;
    if (!sm_count) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    }
    return sm_count;
}

// Resident CTAs per SM this shape reaches with `parts` warps per stripe (0: does not fit).
template <int TW, bool MIX, int NS, int MINB, bool FUSED, int SUBS>
int tiled_ctas_per_sm(int K, int C, int pitch, int parts, bool tmap) {
    if (parts < 1 || TW % parts) return 0;
    if constexpr ((FUSED && !fused_shape_ok(TW, NS, MINB)) || (SUBS > 1 && !subs_shape_ok(TW, NS, MINB))) {
        return 0;
    } else {
        const TileGeom g = tile_geom(K, C, pitch, TW / parts);
        size_t smem = tile_smem_bytes(g, TW, NS, parts, C, MIX, tmap, FUSED, 2);
        if (FUSED && smem > 227 * 1024) smem = tile_smem_bytes(g, TW, NS, parts, C, MIX, tmap, FUSED, 1);
        if (smem > 227 * 1024) return 0;
        auto kern = bas_render_tiled_kernel<TW, MIX, NS, MINB, FUSED, SUBS>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return 0; }
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, (FUSED ? 2 * TW : TW) * 32, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
        return per_sm;
    }
}

template <int TW, bool MIX, int NS, int MINB, bool FUSED, int SUBS>
int launch_tiled(RenderParams prm, int parts, bool want_split, float* workspace, long long workspace_bytes, cudaStream_t st) {
    if (parts < 1 || TW % parts) return BAS_E_UNSUPPORTED;
    if constexpr ((FUSED && !fused_shape_ok(TW, NS, MINB)) || (SUBS > 1 && !subs_shape_ok(TW, NS, MINB))) {
        return BAS_E_UNSUPPORTED;
    } else {
    constexpr int kThreads = (FUSED ? 2 * TW : TW) * 32;
    const int TS = TW / parts;
    const TileGeom g = tile_geom(prm.K, prm.C, prm.pitch, TS);
    // input rows by tensor-map TMA when the signal allows it (whole 32-sample rows, 16-byte aligned)
    CUtensorMap xmap;
    memset(&xmap, 0, sizeof(xmap));
    prm.n_box = tmap_n_box(g.x_rows);
    prm.box_rows = tmap_box_rows(g.x_rows);
    prm.tmap = make_input_tensor_map(&xmap, prm, prm.box_rows) ? 1 : 0;
    prm.nf = FUSED ? 2 : 0;
    size_t smem = tile_smem_bytes(g, TW, NS, parts, prm.C, MIX, prm.tmap != 0, FUSED, prm.nf);
    if (FUSED && smem > 227 * 1024) { prm.nf = 1; smem = tile_smem_bytes(g, TW, NS, parts, prm.C, MIX, prm.tmap != 0, FUSED, 1); }
    if (smem > 227 * 1024) return BAS_E_UNSUPPORTED;
    using Kernel = void (*)(RenderParams, SpanInfo, float*, const CUtensorMap);
    Kernel kern = bas_render_tiled_kernel<TW, MIX, NS, MINB, FUSED, SUBS, false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { bas_set_error("bas_render: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    const long long p_base = prm.p_begin / kBlk * kBlk;
    prm.parts = parts;
    prm.tiles = bas_ceil_div(prm.p_end - p_base, (long long)TS * kWarpTile);
    SpanInfo sp;
    sp.gs = MIX ? prm.n_src : g.D;
    sp.n_groups = MIX ? prm.tiles : prm.tiles * prm.n_src;
    sp.total = sp.n_groups * sp.gs;
    if (sp.total >= 0x7fffffffLL || prm.tiles >= 0x7fffffffLL) {
        bas_set_error("bas_render: more than 2^31 work slices in one launch; render the signal in time ranges");
        return BAS_E_UNSUPPORTED;
    }
    // persistent grid: as many CTAs as the device keeps resident
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem);
    if (e != cudaSuccess || per_sm < 1) { bas_set_error("bas_render: tile shape does not fit an SM"); cudaGetLastError(); return BAS_E_UNSUPPORTED; }
    const long long resident = (long long)device_sm_count() * per_sm;
    long long grid = resident;
    if (grid > sp.n_groups) grid = sp.n_groups;
    // split groups between CTAs only when every span is longer than a group (then a group has at
    // most two contributors) and the caller gave a workspace
    const long long need = (long long)ws_bytes(grid, TS);
    sp.split = (want_split && workspace && workspace_bytes >= need && grid > 1 && sp.total / grid >= sp.gs + 1) ? 1 : 0;
    if constexpr (MIX) {
        // fewer tiles than the device holds CTAs: the CHAIN kernel, spans of at least 1 / kMaxChain of a tile's sources
        const long long min_span = (sp.gs + kMaxChain - 1) / kMaxChain + 1;
        long long grid_chain = resident < sp.total / min_span ? resident : sp.total / min_span;
        if (!sp.split && want_split && workspace && grid_chain > grid && workspace_bytes >= (long long)ws_bytes(grid_chain, TS)) {
            kern = bas_render_tiled_kernel<TW, MIX, NS, MINB, FUSED, SUBS, true>;
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) { bas_set_error("bas_render: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
            grid = grid_chain;
            sp.split = 1;
        }
    }
    sp.epoch = sp.split ? next_epoch() : 0ull;
    e = bas_launch(kern, dim3((unsigned)grid), dim3(kThreads), smem, st, prm, sp, workspace, xmap);
    if (e != cudaSuccess) { bas_set_error("bas_render: tiled launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
    }
}

// One compiled tile shape: warps per CTA x pipeline stages x CTAs per SM the registers allow.
// index = mix + 2 * fused + 4 * (subchunksize 16): [0] one source per tile, [1] mixing, [2] / [3] the same with
// fused filter synthesis, [4..7] the four again for subchunksize 16.
struct TiledShape {
    int tw, ns, minb;
    int (*ctas_per_sm[8])(int K, int C, int pitch, int parts, bool tmap);
    int (*launch[8])(RenderParams prm, int parts, bool want_split, float* workspace, long long workspace_bytes, cudaStream_t st);
};
#define BAS_TILED_FNS(FN, TW_, NS_, MINB_)                                                               \
    { FN<TW_, false, NS_, MINB_, false, 1>, FN<TW_, true, NS_, MINB_, false, 1>,                          \
      FN<TW_, false, NS_, MINB_, true, 1>, FN<TW_, true, NS_, MINB_, true, 1>,                            \
      FN<TW_, false, NS_, MINB_, false, 2>, FN<TW_, true, NS_, MINB_, false, 2>,                          \
      FN<TW_, false, NS_, MINB_, true, 2>, FN<TW_, true, NS_, MINB_, true, 2> }
#define BAS_TILED_SHAPE(TW_, NS_, MINB_)                                                                 \
    { TW_, NS_, MINB_, BAS_TILED_FNS(tiled_ctas_per_sm, TW_, NS_, MINB_), BAS_TILED_FNS(launch_tiled, TW_, NS_, MINB_) }

// defined in render_tw4.cu / render_tw6.cu / render_tw8.cu (one translation unit per tile width, so
// they compile in parallel)
const TiledShape* tiled_shapes_tw4(int* count);
const TiledShape* tiled_shapes_tw6(int* count);
const TiledShape* tiled_shapes_tw8(int* count);

}  // namespace bas_render_detail
