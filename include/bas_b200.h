/* bas_b200.h - C ABI of the B200-native moving-source binaural renderer.
 *
 * The reference (mbjd/binaural-audio-synthesis) has no FFI of its own; its boundary is the
 * Python function surface of apply_hrtf.py.  Each entry point below names the reference
 * function (file:line in /root/reference) whose work it performs.  The Python package
 * `binaural-audio-synthesis_b200` binds these with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative BAS_E_* code on a bad argument, or a
 *     positive cudaError_t; bas_last_error() returns the message of the last failure on the
 *     calling thread;
 *   - "dev" pointers are device (HBM) pointers, "host" pointers are host pointers; the library
 *     never allocates, frees or retains caller memory;
 *   - `stream` is a cudaStream_t passed as void*; all device work is asynchronous on it;
 *   - functions are thread-safe for distinct streams.
 */
#ifndef BAS_B200_H
#define BAS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BAS_ABI_VERSION 1

#define BAS_N_DIRECTIONS 187      /* rows of the measurement grid, sphere.py:127-315 */
#define BAS_MAX_TERMS 16          /* merged gather terms per ear per trajectory point */

/* argument errors */
#define BAS_E_ARG (-1)
#define BAS_E_UNSUPPORTED (-2)
#define BAS_E_NO_DEVICE (-3)

/* scalar type of the azimuth a trajectory returned (NumPy-2 promotion) */
#define BAS_AZ_PYFLOAT 0          /* Python float / int  : sphere.py:103-119 evaluate in float32 */
#define BAS_AZ_F64 1              /* np.float64          : evaluate in float64 */
#define BAS_AZ_F32 2              /* np.float32          : float32 including the modulo at :86 */

/* error bits reported per plan (the reference's own failure modes) */
#define BAS_ERR_AZIM_ASSERT 1     /* sphere.py:87       assert azim >= 0 */
#define BAS_ERR_VERT_ASSERT 2     /* apply_hrtf.py:266  assert 0 <= a <= 1 */
#define BAS_ERR_NONFINITE 4       /* apply_hrtf.py:149  int(floor(nan)) */

/* One gather term: out[m] += weight * bank[row][(m*U - shift) mod L] */
typedef struct bas_term {
    int32_t row_shift;            /* (row << 20) | shift, 0 <= shift < L */
    float weight;
} bas_term;

/* Every integer of the parity contract for one trajectory point. */
typedef struct bas_trace {
    int32_t rows[4];              /* top_before, top_after, bot_before, bot_after (apply_hrtf.py:214-215) */
    int32_t err;
    int32_t pad;
    double alpha_top, alpha_bot, a;
    int64_t lo[2][6];             /* per ear: floor of {top -d, top alpha*d, bot -d, bot alpha*d, -dv, (1-a)dv} */
    int64_t hi[2][6];             /* per ear: ceil  of the same */
} bas_trace;

int bas_abi_version(void);
int bas_last_error(char* buf, size_t len);
int bas_device_count(void);

/* ---- bank upload: load_irs_and_delaydiffs, apply_hrtf.py:23-46 --------------------------
 * irs_dev:   n_rows x L doubles, row-major (one ear), already truncated to L = K*U samples.
 * out_dev:   n_rows x U x K floats, polyphase order: out[row][n % U][n / U] = irs[row][n].  */
int bas_bank_to_polyphase(const double* irs_dev, int n_rows, int L, int U, float* out_dev, void* stream);

/* ---- plan: the scalar part of interpolate_2d, apply_hrtf.py:199-215, :244-252, :261-266,
 *      :272-273, :149-151 and sphere.azim_to_interpolation_params, sphere.py:78-121 -----------
 * diffs_*_dev: 187 x 187 doubles row-major.  elev/azim: n_points doubles (radians).
 * az_kind_dev: n_points bytes (BAS_AZ_*), or NULL to use az_kind_all for every point.
 * terms_dev:   n_points x 2 x BAS_MAX_TERMS.   trace_dev: n_points records or NULL.
 * status_dev:  2 ints: OR of all error bits, and the lowest failing point index (or INT_MAX). */
int bas_plan_build(const double* diffs_left_dev, const double* diffs_right_dev, int U, int L,
                   const double* elev_dev, const double* azim_dev, const uint8_t* az_kind_dev,
                   int az_kind_all, long long n_points, bas_term* terms_dev, bas_trace* trace_dev,
                   int* status_dev, void* stream);

/* Same arithmetic on the host (identical source, plan_math.h) - host pointers throughout.
 * Used for the scalar entry points and by CPU-side index-parity tests.  Returns OR of error bits. */
int bas_plan_build_host(const double* diffs_left, const double* diffs_right, int U, int L,
                        const double* elev, const double* azim, const uint8_t* az_kind,
                        int az_kind_all, long long n_points, bas_term* terms, bas_trace* trace);

/* sphere.azim_to_interpolation_params, sphere.py:78-121 (host scalar).  ring_elev must be one of
 * the ten grid elevations (radians, 1e-5 tolerance); returns BAS_E_ARG otherwise (the reference
 * raises ValueError, sphere.py:100-101), BAS_ERR_AZIM_ASSERT if the azimuth assertion fails. */
int bas_ring_lookup_host(double ring_elev, double azim, int az_kind, int* before, double* alpha, int* after);

/* delay_compensated_interpolation_with_delaydiff, apply_hrtf.py:53-106, scalar part for both ears.
 * one_minus_alpha is passed separately because the reference evaluates it in the dtype of alpha.
 * terms: 2 x BAS_MAX_TERMS (host).  delays: the two returned delays (apply_hrtf.py:106).
 * lo/hi: per ear floor/ceil of {-d, alpha*d} (may be NULL). */
int bas_plan_ring_host(const double* diffs_left, const double* diffs_right, int U, int L, int before,
                       int after, double alpha, double one_minus_alpha, bas_term* terms,
                       double* delays, int64_t* lo, int64_t* hi);

/* ---- IR synthesis: the array part of interpolate_2d (apply_hrtf.py:219-281) and of
 *      delay_compensated_interpolation_with_delaydiff (apply_hrtf.py:86-102) ------------------
 * bank_pp_dev: 2 x 187 x U x K floats (left ear block, then right), polyphase.
 * mode BAS_IR_UPSAMPLED: out[point][ear][n], n < K*U  (return_upsampled=True), row stride out_stride;
 * mode BAS_IR_PLANAR:    out[point][ear][m], m < K, row stride out_stride >= K floats (tail zeroed);
 * mode BAS_IR_ROWS:      out[point][m][ear], m < bas_filter_row_pitch(K) (taps >= K zeroed): the
 *                        filter-row layout bas_render consumes; out_stride is ignored. */
#define BAS_IR_UPSAMPLED 0
#define BAS_IR_PLANAR 1
#define BAS_IR_ROWS 2
int bas_ir_synth(const float* bank_pp_dev, int U, int K, const bas_term* terms_dev, long long n_points,
                 int mode, float* out_dev, long long out_stride, void* stream);

/* Taps per filter row in the BAS_IR_ROWS layout: K rounded up to a multiple of 32, plus 2 (the
 * row pitch in bytes, 8 * pitch, is a multiple of 16 but not of 128). */
int bas_filter_row_pitch(int K);

/* ---- renderer: the chunk/subchunk loops of make_signal_move_2d, apply_hrtf.py:431-453 --------
 * Output-stationary form of the overlap-add:
 *     out_e[p] = sum_k x[p-k] * h_{q(p-k),e}[k],   h_q = (1-alpha_q) H_i + alpha_q H_{i+1}
 * x_dev:     n_src signals, x_stride floats apart; samples >= n_valid are treated as zero (the
 *            zero padding to n_in of apply_hrtf.py:405-406 is implicit).
 * filt_dev:  n_src x (n_in/C + 1) filter rows in the BAS_IR_ROWS layout (16-byte aligned).
 * gains_dev: n_src floats multiplying each source before mixing, or NULL for 1.
 * Output samples p_begin <= p < p_begin + p_count (0 <= p, p_begin + p_count <= n_in + K - 1):
 * out_dev:   mix=0: n_src x 2 x out_stride (planar L then R), out[s][e][p - p_begin];
 *            mix=1: 2 x out_stride holding the gain-weighted sum over sources (deterministic order).
 * peaks_dev: n_src floats, max |out| of each source over the rendered range BEFORE gain and mixing
 *            (for apply_hrtf.py:462), or NULL.  Must be zeroed by the caller (atomic max).
 * variant:   BAS_RENDER_AUTO / _GENERIC (any C, S, K) / _TILED (needs S == 32, C % 32 == 0,
 *            16-byte aligned signals; BAS_E_UNSUPPORTED otherwise).
 * workspace_dev / workspace_bytes: see bas_render_workspace_bytes (16-byte aligned, or NULL). */
#define BAS_RENDER_AUTO 0
#define BAS_RENDER_GENERIC 1
#define BAS_RENDER_TILED 2
#define BAS_RENDER_SPLIT 0x40       /* OR-ed into variant: always balance by splitting tiles between CTAs */
#define BAS_RENDER_NO_SPLIT 0x80    /* OR-ed into variant: never split a tile between CTAs */
int bas_render(const float* x_dev, long long x_stride, long long n_valid, int n_src, long long n_in,
               int C, int S, int K, const float* filt_dev, const float* gains_dev,
               long long p_begin, long long p_count, float* out_dev, long long out_stride, int mix,
               float* peaks_dev, int variant, void* workspace_dev, long long workspace_bytes, void* stream);

/* Scratch the tiled renderer may use to balance work across SMs (a tile split between two CTAs is
 * summed there in a fixed order, so results stay deterministic).  workspace_dev may be NULL: tiles
 * are then never split.  Contents need no initialisation and carry nothing between calls. */
long long bas_render_workspace_bytes(void);

/* apply_hrtf.py:462-464: divide n floats by *peak_dev when it exceeds 1 (no-op otherwise). */
int bas_normalise(float* out_dev, long long n, const float* peak_dev, void* stream);

/* max over n floats of |v| into *peak_dev (atomic max; zero it first). */
int bas_peak(const float* v_dev, long long n, float* peak_dev, void* stream);

/* FP32 pipe probe used by bench.py to state the measured FMA peak beside the HBM roofline:
 * every thread runs `iters` rounds of 16 independent dependent-chain FMAs.
 * packed=0: fma.rn.f32 (32 FMA per thread per round) ; packed=1: fma.rn.f32x2 (same FMA count).
 * sink_dev receives one float per thread so the work cannot be elided.  FMA count of the launch =
 * blocks * threads * iters * 32. */
int bas_probe_fma(int packed, int blocks, int threads, int iters, float* sink_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BAS_B200_H */
