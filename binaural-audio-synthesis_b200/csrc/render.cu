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
//   staging  For every (tile, source) item one thread issues TMA copies into a shared-memory ring guarded
//            by full/empty mbarriers: the input samples as a tiled tensor-map copy with the 128-byte
//            swizzle (rows outside the signal arrive as zeros; lane-per-row reads are conflict free
//            through the XOR), and the boundary-filter rows of the chunks the tile touches as ONE
//            contiguous bulk copy (rows come from ir_synth already interleaved {L,R} and zero padded).
//            There is no __syncthreads in the kernel.
//   math     Lane l of a warp owns the 32 consecutive outputs of block b = b0 + l for both ears: 32
//            fma.rn.f32x2 accumulators {L,R}.  Output block b draws on input subchunks q = b - d:
//                out[32b + r] += x[32q + m] * h_q[32d + r - m]          r, m = 0..31
//            All lanes sit at the same (d, m), so subchunk boundaries are warp-uniform.  The 32x32
//            products of a block are visited diagonal by diagonal (j = r - m fixed): ONE blended tap pair
//            {L,R} is live at a time and feeds the 32 - |j| packed FMAs of its diagonal as the
//            register-reuse operand; the lane's 32 input samples sit in registers (block_diag in
//            render_tiled.cuh).  The half-empty first (d = 0, r >= m) and last (d = D, r < m) blocks are
//            folded into ONE 32x32 block (diagonals >= 0 from the d = 0 filter, < 0 from the d = D
//            filter), so a tile costs exactly ceil(K/32) blocks of 1024 packed FMAs plus 126 blend
//            operations each (89 % useful at K = 256, 94 % at K = 512).
//   taps     Blended on the fly: w = H_i + alpha (H_{i+1} - H_i) from two LDS.128 (two taps each);
//            lanes in the same chunk read the same address (broadcast), other chunks other banks.
//   epilogue Direct 16-byte global stores from registers, peak by warp reduction + one atomicMax.
#include "render_tiled.cuh"

#include <atomic>
#include <chrono>
#include <stdlib.h>

using namespace bas_render_detail;

namespace {

// ------------------------------------------------------------------------------------------------
// generic kernel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bas_render_generic_kernel(RenderParams prm) {
    bas_grid_launch_dependents();
    bas_grid_dependency_wait();
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
        if (prm.accumulate) { mix_l = prm.out[p - prm.p_begin] + mix_l; mix_r = prm.out[prm.out_stride + p - prm.p_begin] + mix_r; }
        prm.out[p - prm.p_begin] = mix_l;
        prm.out[prm.out_stride + p - prm.p_begin] = mix_r;
    }
}

// ------------------------------------------------------------------------------------------------
// peak / normalise  (apply_hrtf.py:462-464)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bas_peak_kernel(const float* __restrict__ v, long long n, float* __restrict__ peak) {
    bas_grid_launch_dependents();
    bas_grid_dependency_wait();
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(v[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomic_max_nonneg(peak, m);
}

__global__ void __launch_bounds__(256)
bas_normalise_kernel(float* __restrict__ v, long long n, const float* __restrict__ peak) {
    bas_grid_launch_dependents();
    bas_grid_dependency_wait();
    const float m = *peak;
    if (!(m > 1.f)) return;                      // apply_hrtf.py:463: only when the peak exceeds 1
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        v[i] = v[i] / m;                         // :464 true division, not a reciprocal multiply
}

}  // namespace

// variant encoding beyond the public ones: BAS_RENDER_TILED | (TW << 8) | (NS << 16) requests a tile
// shape (tuning sweeps; unknown shapes return BAS_E_UNSUPPORTED).
// filt_dev: filter rows written by bas_ir_synth(BAS_IR_ROWS); or, fused, terms_dev + bank_pp2_dev + U: the
// tiled kernel synthesises the rows itself.
// diagnostics: per-CTA time stamps of the tiled kernel (tools/cta_trace.py)
static unsigned long long* g_render_trace = nullptr;
extern "C" int bas_render_set_trace(unsigned long long* trace_dev) { g_render_trace = trace_dev; return 0; }

static int render_common(const float* x_dev, long long x_stride, long long n_valid, int n_src, long long n_in,
                         int C, int S, int K, const float* filt_dev, const bas_term* terms_dev, const float* bank_pp2_dev, int U,
                         const float* gains_dev, long long p_begin, long long p_count, float* out_dev, long long out_stride, int mix,
                         float* peaks_dev, int variant, void* workspace_dev, long long workspace_bytes, void* stream,
                         const bas_route* route = nullptr) {
    const bool fused = filt_dev == nullptr;
    if (route && route->n > 1) {
        BAS_CHECK_ARG(mix == 1 && route->table_dev && route->rank >= 0 && route->rank < route->n && route->n <= 32, "route: mixing renders only, rank < n <= 32");
        BAS_CHECK_ARG(route->len > 0 && route->len % 32 == 0 && route->stride >= route->len && route->stride % 4 == 0, "route: slice length / stride");
        BAS_CHECK_ARG(route->len * route->n >= p_begin + p_count, "route: the slices do not cover the output");
        BAS_CHECK_ARG(!route->arrive_ptrs_dev || route->arrive_counter_dev, "route: arrival signal without a counter");
    }
    BAS_CHECK_ARG(x_dev && out_dev && (filt_dev || (terms_dev && bank_pp2_dev && U >= 1)), "null pointer");
    BAS_CHECK_ARG(n_src >= 1, "n_src");
    BAS_CHECK_ARG(mix == 0 || mix == 1 || mix == BAS_MIX_ACCUMULATE, "mix must be 0, 1 or BAS_MIX_ACCUMULATE");
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
    prm.out = out_dev; prm.out_stride = out_stride; prm.mix = mix ? 1 : 0; prm.accumulate = mix == BAS_MIX_ACCUMULATE ? 1 : 0; prm.peaks = peaks_dev; prm.tiles = 0; prm.parts = 1; prm.tmap = 0; prm.box_rows = 0; prm.n_box = 0;
    prm.terms = reinterpret_cast<const TermDev*>(terms_dev); prm.bank2 = bank_pp2_dev; prm.U = U; prm.nf = 0;
    const bool routed = route && route->n > 1;
    prm.route_table = routed ? route->table_dev : nullptr; prm.route_n = routed ? route->n : 0; prm.route_rank = routed ? route->rank : 0;
    prm.route_len = routed ? route->len : 1; prm.route_stride = routed ? route->stride : 0;
    prm.arrive_ptrs = routed ? route->arrive_ptrs_dev : nullptr; prm.arrive_counter = routed ? route->arrive_counter_dev : nullptr;
    prm.arrive_epoch = routed ? route->arrive_epoch : 0u;
    prm.trace = g_render_trace;

    const int base = variant & 0x3f;
    BAS_CHECK_ARG(base == BAS_RENDER_AUTO || base == BAS_RENDER_GENERIC || base == BAS_RENDER_TILED, "variant");
    // the tiled kernel works on 32-sample input rows: a subchunk is a whole number of rows (32, 64, ...) or half a row (16)
    const int subs = S == kBlk / 2 ? 2 : 1;
    const bool tiled_ok = (S % kBlk == 0 || S == kBlk / 2) && C % kBlk == 0 && n_valid % 4 == 0 && (reinterpret_cast<uintptr_t>(x_dev) & 15) == 0 &&
                          (fused || (reinterpret_cast<uintptr_t>(filt_dev) & 15) == 0) && (n_src == 1 || x_stride % 4 == 0);
    if ((base == BAS_RENDER_TILED || fused) && !tiled_ok) {
        bas_set_error("bas_render: tiled kernel needs subchunksize 16 or a multiple of 32, 32 | C, 4 | n_valid, 16-byte aligned signals and filter rows");
        return BAS_E_UNSUPPORTED;
    }
    if (base != BAS_RENDER_GENERIC && tiled_ok) {
        // Tile shape: TW warps per CTA x NS pipeline stages x CTAs per SM, and `parts` warps per
        // 1024-output stripe.  Variant bits 8..15 request a TW, 16..23 an NS, 24..27 a CTA count per SM,
        // 28..30 parts (code n: 2^(n-1)); otherwise the shape with the lowest estimated time is taken.
        const int tw_req = (variant >> 8) & 0xff, ns_req = (variant >> 16) & 0xff, cta_req = (variant >> 24) & 0xf;
        const int parts_code = (variant >> 28) & 0x7, parts_req = parts_code ? 1 << (parts_code - 1) : 0;
        const int mixi = (prm.mix ? 1 : 0) + (fused ? 2 : 0) + (subs == 2 ? 4 : 0);
        const int D = (K + kBlk - 1) / kBlk;
        const long long p_base = p_begin / kBlk * kBlk;
        const bool split = (variant & BAS_RENDER_SPLIT) || (prm.mix && !(variant & BAS_RENDER_NO_SPLIT));
        const bool can_split = split && workspace_dev != nullptr;
        const bool use_tmap = input_tensor_map_possible(prm);
        const TiledShape* best = nullptr;
        int best_parts = 0;
        double best_cost = 0.0;
        for (int tu = 0; tu < 3; ++tu) {
            int n_shapes = 0;
            const TiledShape* shapes = tu == 0 ? tiled_shapes_tw4(&n_shapes) : tu == 1 ? tiled_shapes_tw6(&n_shapes) : tiled_shapes_tw8(&n_shapes);
            for (int i = 0; i < n_shapes; ++i) {
                const TiledShape& sh = shapes[i];
                if ((tw_req && sh.tw != tw_req) || (ns_req && sh.ns != ns_req) || (cta_req && sh.minb != cta_req)) continue;
                for (int parts = 1; parts <= sh.tw && parts <= 8; ++parts) {
                    if (sh.tw % parts || (parts_req && parts != parts_req)) continue;
                    if (!parts_req && parts > D) continue;                  // nothing left to share
                    const int ctas = sh.ctas_per_sm[mixi](K, C, prm.pitch, parts, use_tmap);
                    if (ctas < 1) continue;
                    // Measured on B200 (tools/tune_render.py): every shape that keeps 8 or 12 warps per SM
                    // busy lands within a few per cent; what separates them is how evenly the tiles of
                    // this launch fill the device.  cost = waves of tiles x blocks a warp runs per tile x
                    // warps that share a scheduler's FMA pipe, with small measured preferences on top.
                    const int ts = sh.tw / parts;
                    const double tiles = (double)bas_ceil_div(prm.p_end - p_base, (long long)ts * 1024) * (prm.mix ? 1 : n_src);
                    const double slots = (double)device_sm_count() * ctas;
                    const double blocks_per_tile = (double)((D + parts - 1) / parts) * (prm.mix ? n_src : 1);
                    double waves = tiles / slots;
                    // mixing with a workspace: a tile's sources may be dealt to a chain of CTAs (render_tiled.cuh), so
                    // fewer tiles than CTAs still fill the device
                    const bool chain = prm.mix && can_split && tiles * n_src / slots >= (double)((n_src + 7) / 8 + 1);
                    if (!(can_split && (waves > 1.0 || chain))) {
                        // whole tiles, dealt round-robin; the last wave only occupies part of every SM and
                        // its CTAs then run with fewer neighbours on the FMA pipe
                        const double full = (double)(long long)waves, rest = waves - full;
                        waves = full + (rest > 0.0 ? (rest * ctas <= 1.0 ? 1.0 / ctas + 0.25 : rest * ctas <= 2.0 ? 2.0 / ctas + 0.15 : 1.0) : 0.0);
                    }
                    const int wps = (ctas * sh.tw + 3) / 4;
                    double cost = waves * blocks_per_tile * wps;
                    // fused: with room for ONE filter-row buffer only (long filters: K = 512 rows of an 8-stripe tile) the
                    // producers cannot run an item ahead - measured 141 us per source against 131 us for the same CTA with
                    // two warps per stripe (half the tile, two buffers; tools/mix_bench.py 32 512 16)
                    const bool one_buffer = fused && tile_smem_bytes(tile_geom(K, C, prm.pitch, ts), sh.tw, sh.ns, parts, C, prm.mix != 0, use_tmap, true, 2) > 227 * 1024;
                    cost *= 1.0 + 0.25 * (parts - 1) + (sh.ns == 1 && ctas == 1 ? 0.02 : 0.0) + (wps < 2 ? 0.3 : 0.0) + (one_buffer ? 0.30 : 0.0) +
                            (prm.mix ? (sh.tw == 8 ? -0.05 : sh.minb == 3 ? 0.03 : 0.0) : (sh.tw == 4 ? 0.0 : 0.05));
                    if (!best || cost < best_cost) { best = &sh; best_parts = parts; best_cost = cost; }
                }
            }
        }
        int rc = BAS_E_UNSUPPORTED;
        if (best) rc = best->launch[mixi](prm, best_parts, split, reinterpret_cast<float*>(workspace_dev), workspace_bytes, st);
        if (rc != BAS_E_UNSUPPORTED || base == BAS_RENDER_TILED) {
            if (rc == BAS_E_UNSUPPORTED && !best)
                bas_set_error("bas_render: no tile shape fits K=%d C=%d (requested TW=%d NS=%d CTAs/SM=%d parts=%d)", K, C, tw_req, ns_req, cta_req, parts_req);
            return rc;
        }
    }
    if (fused) { bas_set_error("bas_render_fused: no tile shape fits K=%d C=%d", K, C); return BAS_E_UNSUPPORTED; }
    if (routed) { bas_set_error("bas_render: routed mixes need the tiled kernel"); return BAS_E_UNSUPPORTED; }
    const int threads = 256;
    const long long blocks = bas_ceil_div(p_count, threads);
    BAS_CHECK_ARG(blocks < 0x7fffffffLL && n_src <= 65535, "launch too large");
    dim3 grid((unsigned)blocks, prm.mix ? 1u : (unsigned)n_src);
    cudaError_t e = bas_launch(bas_render_generic_kernel, grid, dim3(threads), 0, st, prm);
    if (e != cudaSuccess) { bas_set_error("bas_render: generic launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

extern "C" int bas_render(const float* x_dev, long long x_stride, long long n_valid, int n_src, long long n_in,
                          int C, int S, int K, const float* filt_dev, const float* gains_dev,
                          long long p_begin, long long p_count, float* out_dev, long long out_stride, int mix,
                          float* peaks_dev, int variant, void* workspace_dev, long long workspace_bytes, void* stream) {
    BAS_CHECK_ARG(filt_dev, "null pointer");
    return render_common(x_dev, x_stride, n_valid, n_src, n_in, C, S, K, filt_dev, nullptr, nullptr, 0, gains_dev, p_begin, p_count,
                         out_dev, out_stride, mix, peaks_dev, variant, workspace_dev, workspace_bytes, stream);
}

extern "C" int bas_render_fused(const float* x_dev, long long x_stride, long long n_valid, int n_src, long long n_in,
                                int C, int S, int K, const bas_term* terms_dev, const float* bank_pp2_dev, int U,
                                const float* gains_dev, long long p_begin, long long p_count, float* out_dev, long long out_stride,
                                int mix, float* peaks_dev, int variant, void* workspace_dev, long long workspace_bytes, void* stream) {
    BAS_CHECK_ARG(terms_dev && bank_pp2_dev, "null pointer");
    BAS_CHECK_ARG(U >= 1 && (long long)U * K < (1 << 20) && (2LL * BAS_N_DIRECTIONS * U * 2 * K) < 0x7fffffffLL, "U");
    return render_common(x_dev, x_stride, n_valid, n_src, n_in, C, S, K, nullptr, terms_dev, bank_pp2_dev, U, gains_dev, p_begin, p_count,
                         out_dev, out_stride, mix, peaks_dev, variant, workspace_dev, workspace_bytes, stream);
}

extern "C" int bas_render_routed(const float* x_dev, long long x_stride, long long n_valid, int n_src, long long n_in,
                                 int C, int S, int K, const float* filt_dev, const bas_term* terms_dev, const float* bank_pp2_dev, int U,
                                 const float* gains_dev, long long p_begin, long long p_count, float* out_dev, long long out_stride,
                                 float* peaks_dev, int variant, void* workspace_dev, long long workspace_bytes, const bas_route* route,
                                 void* stream) {
    BAS_CHECK_ARG(route, "null route");
    BAS_CHECK_ARG(filt_dev || (terms_dev && bank_pp2_dev), "need filter rows, or plan terms and the doubled bank");
    return render_common(x_dev, x_stride, n_valid, n_src, n_in, C, S, K, filt_dev, filt_dev ? nullptr : terms_dev, bank_pp2_dev, U, gains_dev,
                         p_begin, p_count, out_dev, out_stride, 1, peaks_dev, variant, workspace_dev, workspace_bytes, stream, route);
}

extern "C" int bas_render_fused_supported(int C, int S) { return (S % kBlk == 0 || S == kBlk / 2) && S >= 1 && C % kBlk == 0 && C % S == 0 ? 1 : 0; }

extern "C" int bas_render_fused_shape(int variant) {
    // 1 when the (TW, NS, CTAs per SM) a variant word requests is one the fused kernel is compiled for (or none is requested)
    const int tw = (variant >> 8) & 0xff, ns = (variant >> 16) & 0xff, ctas = (variant >> 24) & 0xf;
    if (!tw && !ns && !ctas) return 1;
    return fused_shape_ok(tw, ns ? ns : 2, ctas ? ctas : (tw == 4 ? 2 : 1)) ? 1 : 0;
}

extern "C" int bas_render_fused_fits(int K, int C, int S, int mix, int variant) {
    // 1 when bas_render_fused has a tile shape for this geometry (shared memory for the input stages, at least one
    // filter-row buffer and, when mixing, the running sums) - callers decide up front instead of falling back mid-job
    if (K < 1 || !bas_render_fused_supported(C, S) || (variant & 0x3f) == BAS_RENDER_GENERIC || !bas_render_fused_shape(variant)) return 0;
    // the answer only depends on the arguments (and the device): remember the last one per host thread - the host
    // pipeline asks once per phase of every make_signal_move_2d call
    static thread_local int last_key[5] = {0, 0, 0, 0, 0}, last_dev = -1, last_answer = 0;
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev == last_dev && last_key[0] == K && last_key[1] == C && last_key[2] == S && last_key[3] == (mix ? 1 : 0) && last_key[4] == variant) return last_answer;
    struct Remember {
        int *key, *dev_slot, *answer_slot, k0, k1, k2, k3, k4, dev, answer = 0;
        ~Remember() { key[0] = k0; key[1] = k1; key[2] = k2; key[3] = k3; key[4] = k4; *dev_slot = dev; *answer_slot = answer; }
    } remember{last_key, &last_dev, &last_answer, K, C, S, mix ? 1 : 0, variant, dev};
    const int tw_req = (variant >> 8) & 0xff;
    const int parts_code = (variant >> 28) & 0x7, parts = parts_code ? 1 << (parts_code - 1) : 1;
    const int idx = (mix ? 1 : 0) + 2 + (S == kBlk / 2 ? 4 : 0);
    const int pitch = bas_filter_row_pitch(K);
    for (int tu = 0; tu < 3; ++tu) {
        int n_shapes = 0;
        const TiledShape* shapes = tu == 0 ? tiled_shapes_tw4(&n_shapes) : tu == 1 ? tiled_shapes_tw6(&n_shapes) : tiled_shapes_tw8(&n_shapes);
        for (int i = 0; i < n_shapes; ++i) {
            if (tw_req && shapes[i].tw != tw_req) continue;
            if (shapes[i].tw % parts) continue;
            if (shapes[i].ctas_per_sm[idx](K, C, pitch, parts, true) > 0 || shapes[i].ctas_per_sm[idx](K, C, pitch, parts, false) > 0) { remember.answer = 1; return 1; }
        }
    }
    return 0;
}

extern "C" long long bas_bank2_floats(int U, int K) {
    // every phase row twice + padding for the second tap of a gather pass (render_tiled.cuh)
    return U < 1 || K < 1 ? BAS_E_ARG : 2LL * BAS_N_DIRECTIONS * U * 2 * K + 1024;
}

namespace bas_render_detail {
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (getenv("BAS_NO_TMAP")) return (EncodeTiledFn) nullptr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return (EncodeTiledFn) nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

bool input_tensor_map_possible(const RenderParams& prm) {
    return encode_tiled_fn() != nullptr && prm.n_valid >= kBlk && prm.n_valid % kBlk == 0 &&
           (reinterpret_cast<uintptr_t>(prm.x) & 15) == 0 && (prm.n_src == 1 || prm.x_stride % 4 == 0) &&
           prm.n_valid / kBlk < (1LL << 31) && prm.n_src < (1 << 30);
}

bool make_input_tensor_map(CUtensorMap* map, const RenderParams& prm, int box_rows) {
    if (!input_tensor_map_possible(prm) || box_rows < 1 || box_rows > 256) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)kBlk, (cuuint64_t)(prm.n_valid / kBlk), (cuuint64_t)prm.n_src};
    const cuuint64_t source_stride = (cuuint64_t)(prm.n_src == 1 ? prm.n_valid : prm.x_stride) * 4;
    const cuuint64_t strides[2] = {(cuuint64_t)kBlk * 4, source_stride};
    const cuuint32_t box[3] = {(cuuint32_t)kBlk, (cuuint32_t)box_rows, 1u};
    const cuuint32_t elem[3] = {1u, 1u, 1u};
    const CUresult rc = encode_tiled_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(prm.x), dims, strides, box, elem,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return rc == CUDA_SUCCESS;
}

unsigned long long next_epoch() {
    static std::atomic<unsigned long long> counter{0};
    static const unsigned long long salt = [] {
        const auto t = std::chrono::high_resolution_clock::now().time_since_epoch().count();
        unsigned long long z = (unsigned long long)t + 0x9E3779B97F4A7C15ull * (unsigned long long)(uintptr_t)&counter;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return (z ^ (z >> 31)) | 1ull;
    }();
    return (salt << 24) ^ (++counter);
}
}  // namespace bas_render_detail

extern "C" long long bas_render_workspace_bytes(void) {
    // resident stripes (at most 16 warps per SM): a flag and 1024 {L,R} partial sums each, whatever the shape
    return (long long)ws_bytes((long long)device_sm_count() * 16, 1);
}

extern "C" int bas_peak(const float* v_dev, long long n, float* peak_dev, void* stream) {
    BAS_CHECK_ARG(v_dev && peak_dev && n >= 0, "bad pointer or size");
    if (n == 0) return 0;
    const long long blocks = bas_ceil_div(n, 256 * 8);
    BAS_CUDA(bas_launch(bas_peak_kernel, dim3((unsigned)(blocks > 148 * 8 ? 148 * 8 : blocks)), dim3(256), 0, (cudaStream_t)stream,
                        v_dev, n, peak_dev));
    return 0;
}

extern "C" int bas_normalise(float* out_dev, long long n, const float* peak_dev, void* stream) {
    BAS_CHECK_ARG(out_dev && peak_dev && n >= 0, "bad pointer or size");
    if (n == 0) return 0;
    const long long blocks = bas_ceil_div(n, 256 * 8);
    BAS_CUDA(bas_launch(bas_normalise_kernel, dim3((unsigned)(blocks > 148 * 8 ? 148 * 8 : blocks)), dim3(256), 0, (cudaStream_t)stream,
                        out_dev, n, peak_dev));
    return 0;
}
