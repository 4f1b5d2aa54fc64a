// K1  plan_build: one fp64 thread per trajectory point.
//
// Replaces the scalar arithmetic of interpolate_2d (apply_hrtf.py:199-215, :244-252, :261-266,
// :272-273), of delay_signal_float (apply_hrtf.py:149-151) and of
// sphere.azim_to_interpolation_params (sphere.py:78-121).  The arithmetic itself lives in
// plan_math.h and is also compiled for the host (bas_plan_build_host) from the same source.
//
// COMPILE THIS FILE WITH -fmad=false: a fused multiply-add would change floor()/ceil() inputs.
#include <limits.h>

#include "bas_internal.cuh"
#include "plan_math.h"

static_assert(sizeof(BasTerm) == sizeof(bas_term), "term layout");
static_assert(sizeof(BasTrace) == sizeof(bas_trace), "trace layout");
static_assert(BAS_MAX_TERMS == 16 && BAS_N_DIR == BAS_N_DIRECTIONS, "constants");

// ---- device kernel ---------------------------------------------------------------------------
// One thread per (trajectory point, ear): the ear-independent ring lookups are recomputed by both
// threads of a point (cheap) so that all 2*n_points threads run independently; 64-thread CTAs spread
// a 5169-point trajectory over every SM.
// status[0]: OR of the error bits of every point; status[1]: bitwise complement of the minimum over failing
// points of (index << 3 | error bits of that point, both ears) - the reference raises at the FIRST bad
// trajectory point (apply_hrtf.py:429/:435 call interpolate_2d in time order), so the host needs the
// bits of the earliest one.  Stored complemented (atomicMax) so that ALL-ZERO words mean "no error" and
// one memset clears the status together with the peaks that follow it.  The two ear threads of a point
// are neighbouring lanes.
__device__ __forceinline__ void report_error(int* __restrict__ status, int err, long long idx) {
    const int both = err | __shfl_xor_sync(__activemask(), err, 1);
    if (both && status) {
        atomicOr(status, both);
        const long long capped = idx > BAS_STATUS_MAX_INDEX ? BAS_STATUS_MAX_INDEX : idx;
        atomicMax(reinterpret_cast<unsigned*>(status) + 1, ~(unsigned)(capped << 3 | (both & 7)));
    }
}

__device__ __forceinline__ void
plan_thread(const double* __restrict__ diffs_l, const double* __restrict__ diffs_r, int U, long long L,
            double elev, double azim, int kind, long long p, int ear,
            BasTerm* __restrict__ terms, BasTrace* __restrict__ trace, int* __restrict__ status, long long point_offset) {
    const BasPointGeom g = bas_point_geom(elev, azim, kind);
    BasTerm local[BAS_MAX_TERMS];
    long long lo[6], hi[6];
    const int err = g.err | bas_plan_point_ear(ear ? diffs_r : diffs_l, U, L, g, local, lo, hi);
    // 16 x 8 B = 128 B per (point, ear), written as 8 x 16 B
    int4* dst = reinterpret_cast<int4*>(terms + (p * 2 + ear) * BAS_MAX_TERMS);
#pragma unroll
    for (int i = 0; i < BAS_MAX_TERMS / 2; ++i)
        dst[i] = make_int4(local[2 * i].row_shift, __float_as_int(local[2 * i].weight),
                           local[2 * i + 1].row_shift, __float_as_int(local[2 * i + 1].weight));
    if (trace) {
        BasTrace* tr = trace + p;
        if (ear == 0) {
            tr->rows[0] = g.top.before; tr->rows[1] = g.top.after; tr->rows[2] = g.bot.before; tr->rows[3] = g.bot.after;
            tr->alpha_top = g.top.alpha; tr->alpha_bot = g.bot.alpha; tr->a = g.a; tr->pad = 0;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) { tr->lo[ear][i] = lo[i]; tr->hi[ear][i] = hi[i]; }
        if (ear == 0) tr->err = 0;
        __syncwarp();
        if (err) atomicOr(&tr->err, err);
    }
    report_error(status, err, p + point_offset);
}

__global__ void __launch_bounds__(64)
bas_plan_kernel(const double* __restrict__ diffs_l, const double* __restrict__ diffs_r, int U, long long L,
                const double* __restrict__ elev, const double* __restrict__ azim,
                const uint8_t* __restrict__ az_kind, int az_kind_all, long long n_points,
                BasTerm* __restrict__ terms, BasTrace* __restrict__ trace, int* __restrict__ status,
                long long point_offset, long long run, long long run_stride) {
    bas_grid_launch_dependents();
    bas_grid_dependency_wait();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long q = t >> 1;
    if (q >= n_points) return;
    // the n_points points are `run` consecutive entries out of every `run_stride` (the same stretch of every
    // source's trajectory); run == run_stride: one contiguous list
    const long long p = run == run_stride ? q : (q / run) * run_stride + q % run;
    plan_thread(diffs_l, diffs_r, U, L, elev[p], azim[p], az_kind ? (int)az_kind[p] : az_kind_all, p, (int)(t & 1),
                terms, trace, status, point_offset);
}

// Directions carried in the kernel's own parameters (up to 32,764 bytes since CUDA 12.1): they reach
// the device with the launch, through the command stream, instead of through a host -> device copy.
// The host pipeline uses this while the copy engine is busy uploading the signal - a copy of the
// directions would queue behind megabytes of samples (pipeline.cu).
constexpr int kInlinePoints = 1984;                            // 2 x 1984 x 8 B = 31,744 B
struct InlineDirs { double elev[kInlinePoints]; double azim[kInlinePoints]; };

__global__ void __launch_bounds__(64)
bas_plan_inline_kernel(const __grid_constant__ InlineDirs dirs, const double* __restrict__ diffs_l,
                       const double* __restrict__ diffs_r, int U, long long L, int az_kind_all, int n_points,
                       BasTerm* __restrict__ terms, int* __restrict__ status, long long point_offset) {
    bas_grid_launch_dependents();
    bas_grid_dependency_wait();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = t >> 1;
    if (p >= n_points) return;
    plan_thread(diffs_l, diffs_r, U, L, dirs.elev[p], dirs.azim[p], az_kind_all, p, t & 1, terms, nullptr, status, point_offset);
}

// A range of a longer direction list: failing indices are reported as point_offset + i and the status
// words are only reset on request (the host pipeline plans a trajectory in several phases).
int bas_plan_build_range(const double* diffs_left_dev, const double* diffs_right_dev, int U, int L,
                         const double* elev_dev, const double* azim_dev, const uint8_t* az_kind_dev,
                         int az_kind_all, long long n_points, bas_term* terms_dev, bas_trace* trace_dev,
                         int* status_dev, long long point_offset, int reset_status, void* stream) {
    return bas_plan_build_runs(diffs_left_dev, diffs_right_dev, U, L, elev_dev, azim_dev, az_kind_dev, az_kind_all, 1, n_points, n_points,
                               terms_dev, trace_dev, status_dev, point_offset, reset_status, stream);
}

// `rows` runs of `run` consecutive points, run_stride entries apart in every array (the same stretch of the
// trajectory of every source), in ONE launch.
int bas_plan_build_runs(const double* diffs_left_dev, const double* diffs_right_dev, int U, int L,
                        const double* elev_dev, const double* azim_dev, const uint8_t* az_kind_dev,
                        int az_kind_all, long long rows, long long run, long long run_stride, bas_term* terms_dev, bas_trace* trace_dev,
                        int* status_dev, long long point_offset, int reset_status, void* stream) {
    const long long n_points = rows * run;
    BAS_CHECK_ARG(rows >= 0 && run >= 0 && (rows <= 1 || run_stride >= run), "runs");
    if (rows <= 1) run_stride = run;
    BAS_CHECK_ARG(diffs_left_dev && diffs_right_dev && elev_dev && azim_dev && terms_dev, "null pointer");
    BAS_CHECK_ARG(U >= 1 && L >= U && L % U == 0 && L < (1 << 20), "need 1 <= U, U | L, L < 2^20");
    BAS_CHECK_ARG(az_kind_all >= 0 && az_kind_all <= 2, "az_kind_all");
    BAS_CHECK_ARG(n_points >= 0, "n_points");
    if (n_points == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (status_dev && reset_status) {
        BAS_CUDA(cudaMemsetAsync(status_dev, 0, 2 * sizeof(int), st));     // a memset keeps the call capturable in a CUDA graph
    }
    const int threads = 64;
    const long long blocks = bas_ceil_div(2 * n_points, threads);
    BAS_CHECK_ARG(blocks < 0x7fffffffLL, "too many points for one launch");
    BAS_CUDA(bas_launch(bas_plan_kernel, dim3((unsigned)blocks), dim3(threads), 0, st,
                        diffs_left_dev, diffs_right_dev, U, (long long)L, elev_dev, azim_dev, az_kind_dev, az_kind_all, n_points,
                        reinterpret_cast<BasTerm*>(terms_dev), reinterpret_cast<BasTrace*>(trace_dev), status_dev, point_offset,
                        run, run_stride));
    return 0;
}

// HOST direction arrays, one az kind for all: planned in launches of kInlinePoints directions each.
int bas_plan_build_inline(const double* diffs_left_dev, const double* diffs_right_dev, int U, int L,
                          const double* elev_host, const double* azim_host, int az_kind_all, long long n_points,
                          bas_term* terms_dev, int* status_dev, long long point_offset, void* stream) {
    BAS_CHECK_ARG(diffs_left_dev && diffs_right_dev && elev_host && azim_host && terms_dev, "null pointer");
    BAS_CHECK_ARG(U >= 1 && L >= U && L % U == 0 && L < (1 << 20), "need 1 <= U, U | L, L < 2^20");
    BAS_CHECK_ARG(az_kind_all >= 0 && az_kind_all <= 2, "az_kind_all");
    BAS_CHECK_ARG(n_points >= 0, "n_points");
    static thread_local InlineDirs dirs;                        // copied into the launch by cudaLaunchKernel
    for (long long first = 0; first < n_points; first += kInlinePoints) {
        const int m = (int)(n_points - first < kInlinePoints ? n_points - first : kInlinePoints);
        memcpy(dirs.elev, elev_host + first, (size_t)m * 8);
        memcpy(dirs.azim, azim_host + first, (size_t)m * 8);
        bas_plan_inline_kernel<<<(unsigned)bas_ceil_div(2 * m, 64), 64, 0, (cudaStream_t)stream>>>(
            dirs, diffs_left_dev, diffs_right_dev, U, (long long)L, az_kind_all, m,
            reinterpret_cast<BasTerm*>(terms_dev) + first * 2 * BAS_MAX_TERMS, status_dev, point_offset + first);
        BAS_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int bas_plan_build(const double* diffs_left_dev, const double* diffs_right_dev, int U, int L,
                              const double* elev_dev, const double* azim_dev, const uint8_t* az_kind_dev,
                              int az_kind_all, long long n_points, bas_term* terms_dev, bas_trace* trace_dev,
                              int* status_dev, void* stream) {
    return bas_plan_build_range(diffs_left_dev, diffs_right_dev, U, L, elev_dev, azim_dev, az_kind_dev, az_kind_all,
                                n_points, terms_dev, trace_dev, status_dev, 0, 1, stream);
}

// ---- host twins (same plan_math.h) -----------------------------------------------------------
extern "C" int bas_plan_build_host(const double* diffs_left, const double* diffs_right, int U, int L,
                                   const double* elev, const double* azim, const uint8_t* az_kind,
                                   int az_kind_all, long long n_points, bas_term* terms, bas_trace* trace) {
    BAS_CHECK_ARG(diffs_left && diffs_right && elev && azim && terms, "null pointer");
    BAS_CHECK_ARG(U >= 1 && L >= U && L % U == 0 && L < (1 << 20), "need 1 <= U, U | L, L < 2^20");
    int err = 0;
    for (long long p = 0; p < n_points; ++p) {
        const int kind = az_kind ? (int)az_kind[p] : az_kind_all;
        err |= bas_plan_point(diffs_left, diffs_right, U, L, elev[p], azim[p], kind,
                              reinterpret_cast<BasTerm*>(terms) + p * 2 * BAS_MAX_TERMS,
                              trace ? reinterpret_cast<BasTrace*>(trace) + p : nullptr);
    }
    return err;
}

extern "C" int bas_ring_lookup_host(double ring_elev, double azim, int az_kind, int* before, double* alpha, int* after) {
    BAS_CHECK_ARG(before && alpha && after, "null pointer");
    BAS_CHECK_ARG(az_kind >= 0 && az_kind <= 2, "az_kind");
    // sphere.py:88 clips the elevation, :90-98 match it against the table with 1e-5 tolerance
    double e = ring_elev;
    if (e < bas_ring_elev(0)) e = bas_ring_elev(0);
    if (e > bas_ring_elev(BAS_N_RING - 1)) e = bas_ring_elev(BAS_N_RING - 1);
    int r = -1;
    for (int k = 0; k < BAS_N_RING; ++k)
        if (fabs(bas_ring_elev(k) - e) < 0.00001) r = k;
    // the azimuth assertion (sphere.py:87) precedes the ring check (:100-101)
    BasRing rg;
    const int err = bas_ring_lookup(r < 0 ? 0 : r, azim, az_kind, &rg);
    if (err) return err;
    if (r < 0) {
        bas_set_error("bas_ring_lookup_host: elevation %.9g is not a grid ring", ring_elev);
        return BAS_E_ARG;
    }
    *before = rg.before; *alpha = rg.alpha; *after = rg.after;
    return 0;
}

// Ring plans on the device: one thread per (item, ear), same bas_plan_ring_ear as the host twin below.
// rows: {before, after} per item; weights: {alpha, 1 - alpha} per item (the caller evaluates 1 - alpha in
// alpha's own precision, apply_hrtf.py:90); delays: the two values the reference returns (:106).
__global__ void __launch_bounds__(64)
bas_plan_ring_kernel(const double* __restrict__ diffs_l, const double* __restrict__ diffs_r, int U, long long L,
                     const int* __restrict__ rows, const double* __restrict__ weights, long long n,
                     BasTerm* __restrict__ terms, double* __restrict__ delays, int* __restrict__ status) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = t >> 1;
    const int ear = (int)(t & 1);
    if (i >= n) return;
    BasTerm local[BAS_MAX_TERMS];
    long long lo[2], hi[2];
    double d = 0.0;
    const int err = bas_plan_ring_ear(ear ? diffs_r : diffs_l, U, L, rows[2 * i], rows[2 * i + 1], weights[2 * i], weights[2 * i + 1],
                                      local, &d, lo, hi);
    int4* dst = reinterpret_cast<int4*>(terms + (i * 2 + ear) * BAS_MAX_TERMS);
#pragma unroll
    for (int k = 0; k < BAS_MAX_TERMS / 2; ++k)
        dst[k] = make_int4(local[2 * k].row_shift, __float_as_int(local[2 * k].weight),
                           local[2 * k + 1].row_shift, __float_as_int(local[2 * k + 1].weight));
    delays[2 * i + ear] = d;
    report_error(status, err, i);
}

extern "C" int bas_plan_ring(const double* diffs_left_dev, const double* diffs_right_dev, int U, int L, const int* rows_dev,
                             const double* weights_dev, long long n, bas_term* terms_dev, double* delays_dev, int* status_dev,
                             void* stream) {
    BAS_CHECK_ARG(diffs_left_dev && diffs_right_dev && rows_dev && weights_dev && terms_dev && delays_dev, "null pointer");
    BAS_CHECK_ARG(U >= 1 && L >= U && L % U == 0 && L < (1 << 20), "need 1 <= U, U | L, L < 2^20");
    BAS_CHECK_ARG(n >= 0 && n < 0x3fffffffLL, "n");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (status_dev) BAS_CUDA(cudaMemsetAsync(status_dev, 0, 2 * sizeof(int), st));
    bas_plan_ring_kernel<<<(unsigned)bas_ceil_div(2 * n, 64), 64, 0, st>>>(diffs_left_dev, diffs_right_dev, U, (long long)L, rows_dev, weights_dev,
                                                                         n, reinterpret_cast<BasTerm*>(terms_dev), delays_dev, status_dev);
    BAS_LAUNCH_CHECK();
    return 0;
}

extern "C" int bas_plan_ring_host(const double* diffs_left, const double* diffs_right, int U, int L, int before,
                                  int after, double alpha, double one_minus_alpha, bas_term* terms,
                                  double* delays, int64_t* lo, int64_t* hi) {
    BAS_CHECK_ARG(diffs_left && diffs_right && terms && delays, "null pointer");
    BAS_CHECK_ARG(before >= 0 && before < BAS_N_DIR && after >= 0 && after < BAS_N_DIR, "row index");
    BAS_CHECK_ARG(U >= 1 && L >= U && L % U == 0 && L < (1 << 20), "need 1 <= U, U | L, L < 2^20");
    int err = 0;
    for (int e = 0; e < 2; ++e) {
        long long l2[2], h2[2];
        err |= bas_plan_ring_ear(e ? diffs_right : diffs_left, U, L, before, after, alpha, one_minus_alpha,
                                 reinterpret_cast<BasTerm*>(terms) + e * BAS_MAX_TERMS, delays + e, l2, h2);
        if (lo && hi) { lo[2 * e] = l2[0]; lo[2 * e + 1] = l2[1]; hi[2 * e] = h2[0]; hi[2 * e + 1] = h2[1]; }
    }
    return err;
}

// ---- delay_signal_float, apply_hrtf.py:127-165, as a callable of its own ----------------------
// out[i] = (1 - a) * x[(i*D - before) mod n] + a * x[(i*D - after) mod n]: the two np.roll's (:156-157),
// the decimation (:160-163) and the blend (:165) in float64, products rounded separately like numpy
// (this file is compiled with -fmad=false), so the result is bit-identical to the reference's.
__global__ void __launch_bounds__(256)
bas_delay_signal_kernel(const double* __restrict__ x, long long n, long long before, long long after, double a,
                        long long D, long long n_out, double* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    long long ib = (i * D - before) % n, ia = (i * D - after) % n;
    ib += ib < 0 ? n : 0;
    ia += ia < 0 ? n : 0;
    out[i] = (1 - a) * x[ib] + a * x[ia];
}

extern "C" int bas_delay_signal_float(const double* in_dev, long long n, long long before, long long after, double a,
                                      int downsample, double* out_dev, void* stream) {
    BAS_CHECK_ARG(in_dev && out_dev, "null pointer");
    BAS_CHECK_ARG(n >= 0 && n < (1LL << 40), "n");
    if (n == 0) return 0;
    const long long D = downsample > 1 ? downsample : 1;            // :160 only decimates when downsample > 1
    const long long n_out = (n + D - 1) / D;                        // len(np.arange(0, n, D))
    const long long blocks = bas_ceil_div(n_out, 256);
    BAS_CHECK_ARG(blocks < 0x7fffffffLL, "signal too long for one launch");
    // reduce the shifts once on the host so the kernel's products cannot overflow
    bas_delay_signal_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in_dev, n, before % n, after % n, a, D, n_out, out_dev);
    BAS_LAUNCH_CHECK();
    return 0;
}
