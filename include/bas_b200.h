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

#define BAS_ABI_VERSION 5

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

/* ---- bank builder: the offline preprocessing of upsample_irs.m (the input side of the path) ----
 * hrir_dev: n_rows x n doubles, one ear's measured HRIRs (content_m, upsample_irs.m:25-26).
 * h_dev:    the resampling FIR, n_taps = 2 Lh + 1 doubles (zero phase; designed on the host, see
 *           bank_builder.py), applied as y[m] = sum_k h[Lh + m - k U] x[k].
 * bas_bank_upsample    -> out_dev: n_rows x (n U) doubles = resample(row, U, 1)      (:42-43)
 * bas_bank_delay_diffs -> diffs_dev: n_rows x n_rows doubles, delaydifference of every pair i < j
 *                         (:59-77, parabolic peak :88-101), antisymmetrised (:31-32), zero diagonal. */
int bas_bank_upsample(const double* hrir_dev, int n_rows, int n, int U, const double* h_dev, int n_taps,
                      double* out_dev, void* stream);
int bas_bank_delay_diffs(const double* hrir_dev, int n_rows, int n, int U, const double* h_dev, int n_taps,
                         double* diffs_dev, void* stream);

/* ---- plan: the scalar part of interpolate_2d, apply_hrtf.py:199-215, :244-252, :261-266,
 *      :272-273, :149-151 and sphere.azim_to_interpolation_params, sphere.py:78-121 -----------
 * diffs_*_dev: 187 x 187 doubles row-major.  elev/azim: n_points doubles (radians).
 * az_kind_dev: n_points bytes (BAS_AZ_*), or NULL to use az_kind_all for every point.
 * terms_dev:   n_points x 2 x BAS_MAX_TERMS.   trace_dev: n_points records or NULL.
 * status_dev:  2 ints, zeroed by the call: [0] OR of the error bits of all points; [1] the bitwise
 *              complement of min over failing points of (point index << 3 | error bits of that point) -
 *              index and bits of the EARLIEST failing point, which is where the reference would have
 *              raised.  All-zero words = no error (indices are capped at BAS_STATUS_MAX_INDEX). */
#define BAS_STATUS_MAX_INDEX 0x0fffffff
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

/* delay_compensated_interpolation_with_delaydiff, apply_hrtf.py:53-106, scalar part on the device for n
 * (before, after, alpha) triples: rows_dev {before, after} x n (grid rows, 0..186), weights_dev
 * {alpha, 1 - alpha} x n (1 - alpha evaluated by the caller in alpha's own precision, :90),
 * terms_dev n x 2 x BAS_MAX_TERMS, delays_dev n x 2 (the returned delays, :106), status_dev as in
 * bas_plan_build.  Used by the ring entry points and by the legacy renderer make_signal_move (:334). */
int bas_plan_ring(const double* diffs_left_dev, const double* diffs_right_dev, int U, int L, const int* rows_dev,
                  const double* weights_dev, long long n, bas_term* terms_dev, double* delays_dev, int* status_dev,
                  void* stream);

/* delay_compensated_interpolation_with_delaydiff, apply_hrtf.py:53-106, scalar part for both ears (host
 * twin of bas_plan_ring, same source; CPU-side index-parity tests).
 * one_minus_alpha is passed separately because the reference evaluates it in the dtype of alpha.
 * terms: 2 x BAS_MAX_TERMS (host).  delays: the two returned delays (apply_hrtf.py:106).
 * lo/hi: per ear floor/ceil of {-d, alpha*d} (may be NULL). */
int bas_plan_ring_host(const double* diffs_left, const double* diffs_right, int U, int L, int before,
                       int after, double alpha, double one_minus_alpha, bas_term* terms,
                       double* delays, int64_t* lo, int64_t* hi);

/* ---- delay_signal_float, apply_hrtf.py:127-165, as a callable of its own (inside interpolate_2d its
 *      gathers are fused into bas_ir_synth) ----------------------------------------------------
 * in_dev: n doubles.  before = int(floor(samples)), after = int(ceil(samples)), a = samples - before are
 * evaluated by the caller exactly as :149-151 do.  out_dev: ceil(n / D) doubles, D = max(downsample, 1):
 *     out[i] = (1 - a) * in[(i*D - before) mod n] + a * in[(i*D - after) mod n]
 * (circular like np.roll, :156-157; float64, products rounded separately: bit-identical to numpy). */
int bas_delay_signal_float(const double* in_dev, long long n, long long before, long long after, double a,
                           int downsample, double* out_dev, void* stream);

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
 *            mix=1: 2 x out_stride holding the gain-weighted sum over sources (deterministic order);
 *            mix=BAS_MIX_ACCUMULATE: the same sum ADDED to what out_dev already holds, so that groups
 *            of sources can be rendered in turn (earlier groups first) into one mix.
 * peaks_dev: n_src floats, max |out| of each source over the rendered range BEFORE gain and mixing
 *            (for apply_hrtf.py:462), or NULL.  Must be zeroed by the caller (atomic max).
 * variant:   BAS_RENDER_AUTO / _GENERIC (any C, S, K) / _TILED (needs S == 32, C % 32 == 0,
 *            16-byte aligned signals; BAS_E_UNSUPPORTED otherwise).
 * workspace_dev / workspace_bytes: see bas_render_workspace_bytes (16-byte aligned, or NULL). */
#define BAS_RENDER_AUTO 0
#define BAS_RENDER_GENERIC 1
#define BAS_RENDER_TILED 2
#define BAS_MIX_ACCUMULATE 2
#define BAS_RENDER_SPLIT 0x40       /* OR-ed into variant: always balance by splitting tiles between CTAs */
#define BAS_RENDER_NO_SPLIT 0x80    /* OR-ed into variant: never split a tile between CTAs */
int bas_render(const float* x_dev, long long x_stride, long long n_valid, int n_src, long long n_in,
               int C, int S, int K, const float* filt_dev, const float* gains_dev,
               long long p_begin, long long p_count, float* out_dev, long long out_stride, int mix,
               float* peaks_dev, int variant, void* workspace_dev, long long workspace_bytes, void* stream);

/* The same renderer with the filter rows synthesised inside the tiled kernel (interpolate_2d's array part,
 * apply_hrtf.py:219-281, fused into the chunk loop of :431-453): instead of filt_dev it takes the plan terms
 * of bas_plan_build for n_src x (n_in/C + 1) chunk boundaries and the bank in the layout of
 * bas_bank2_floats(): the polyphase bank of bas_bank_to_polyphase with every phase row stored twice in a
 * row, [ear][row][U][2K] floats followed by padding (total bas_bank2_floats(U, K) floats), so that the
 * circular index (m - adv) mod K is the plain index m + K - adv.  Rows are bit-identical to bas_ir_synth's.
 * Needs the tiled kernel (S == 32, 32 | C, see bas_render_fused_supported); BAS_E_UNSUPPORTED otherwise. */
int bas_render_fused(const float* x_dev, long long x_stride, long long n_valid, int n_src, long long n_in,
                     int C, int S, int K, const bas_term* terms_dev, const float* bank_pp2_dev, int U,
                     const float* gains_dev, long long p_begin, long long p_count, float* out_dev, long long out_stride,
                     int mix, float* peaks_dev, int variant, void* workspace_dev, long long workspace_bytes, void* stream);
int bas_render_fused_supported(int C, int S);
int bas_render_fused_shape(int variant);      /* 1: the tile shape `variant` requests is compiled for the fused kernel */
int bas_render_fused_fits(int K, int C, int S, int mix, int variant);   /* 1: some fused tile shape fits this geometry (needs a device) */
long long bas_bank2_floats(int U, int K);

/* ---- by-source sharding over several GPUs: the per-rank mixes summed over peer memory ----
 * The reference is single-process; a mix of many sources is the sum of independent make_signal_move_2d calls
 * (apply_hrtf.py:356-466), so sources shard over GPUs and the (2, N_out) per-rank mixes must be added.  Here
 * the addition is fused with the render: the output is cut into n slices of `len` samples, slice o owned by rank
 * o; bas_render_routed is bas_render(mix = 1) / bas_render_fused whose epilogue stores every finished tile into the
 * OWNER's receive buffer over NVLink (table_dev[o], peer-mapped, laid out [writer rank][ear][stride]) while the
 * other tiles are still being computed.  Then, on every rank, in stream order:
 *   bas_peer_signal(arrived_ptrs, n, rank, epoch)   "my tiles have landed": one release-store per peer
 *   bas_peer_reduce(...)                            waits for all n writers, sums their slices in rank order
 *                                                   (deterministic), stores the sum into the result buffers of
 *                                                   result_ptrs_dev[0 .. n_results) - every rank (replicated mix) or
 *                                                   this rank only (mix left sharded by time) - and its last CTA
 *                                                   signals "slice done" to every peer.  arrive_ptrs_dev != NULL
 *                                                   folds bas_peer_signal into this launch
 *   bas_peer_wait(done_flags, n, epoch)             replicated: all n slices are in this rank's result buffer;
 *                                                   sharded: called before the NEXT routed render, it keeps a fast
 *                                                   rank from overwriting receive buffers an owner is still summing
 * epoch counts steps (any value that increases by one per step).  Flags, receive and result buffers are the
 * caller's (symmetric allocations; flags and counter zeroed once); every wait is bounded by elapsed time. */
typedef struct bas_route {
    float* const* table_dev;      /* device array: receive buffer of every rank (peer-mapped pointers) */
    int n, rank;                  /* ranks; this rank */
    long long len;                /* output samples per slice (multiple of 32); n * len >= N_out */
    long long stride;             /* floats between the ear rows of a receive block (>= len, multiple of 4) */
    /* optional: fold bas_peer_signal into the render - the last CTA of the launch to finish stores arrive_epoch
     * into flag [rank] of every rank's arrived-array.  Set it on the LAST routed render of a step only. */
    unsigned* const* arrive_ptrs_dev;   /* device array: arrived-flags of every rank (peer-mapped), or NULL */
    unsigned* arrive_counter_dev;       /* one word of this rank, zeroed once; the launches of a stream share it */
    unsigned arrive_epoch;
    unsigned reserved;
} bas_route;
int bas_render_routed(const float* x_dev, long long x_stride, long long n_valid, int n_src, long long n_in,
                      int C, int S, int K, const float* filt_dev, const bas_term* terms_dev, const float* bank_pp2_dev, int U,
                      const float* gains_dev, long long p_begin, long long p_count, float* out_dev, long long out_stride,
                      float* peaks_dev, int variant, void* workspace_dev, long long workspace_bytes, const bas_route* route,
                      void* stream);
int bas_peer_signal(unsigned* const* flag_ptrs_dev, int n, int slot, unsigned epoch, void* stream);
int bas_peer_reduce(const float* recv_dev, int n, long long stride, long long valid, float* const* result_ptrs_dev, int n_results,
                    long long result_stride, long long slice_begin, const unsigned* arrived_dev, unsigned epoch,
                    unsigned* const* done_ptrs_dev, int rank, unsigned* counter_dev, unsigned* const* arrive_ptrs_dev,
                    void* stream);
int bas_peer_wait(const unsigned* flags_dev, int n, unsigned epoch, void* stream);
/* The same wait as n stream memory operations (cuStreamWaitValue32, >=): nothing is resident on the SMs while a peer
 * is late.  BAS_E_UNSUPPORTED when the driver does not offer it: use bas_peer_wait. */
int bas_peer_stream_wait(const unsigned* flags_dev, int n, unsigned epoch, void* stream);

/* ---- one render step as one call: apply_hrtf.py:429-435 feeding :438-453 and :459-464 ----------
 * bas_render_step enqueues, on `stream`: [BAS_STEP_PLAN] one memset of small_dev, bas_plan_build for the
 * n_src x (n_in/C + 1) directions, and (unless BAS_STEP_FUSED) bas_ir_synth into filt_dev;
 * [BAS_STEP_RENDER] bas_render / bas_render_fused of output samples [p_begin, p_begin + p_count);
 * [BAS_STEP_NORMALISE] (mix == 0) bas_normalise of every source by its own peak.  The kernels of one call
 * are chained with programmatic dependent launch.  A job whose output is rendered in several time segments
 * calls once with BAS_STEP_PLAN and then once per segment with BAS_STEP_RENDER. */
#define BAS_STEP_PLAN 1
#define BAS_STEP_RENDER 2
#define BAS_STEP_NORMALISE 4
#define BAS_STEP_FUSED 8
typedef struct bas_step_job {
    int n_src;                    /* mono sources */
    int C, S, K, U;               /* chunksize, subchunksize, taps, upsampling of the bank */
    int mix;                      /* 0, 1 or BAS_MIX_ACCUMULATE, as in bas_render */
    int variant;                  /* bas_render variant */
    int az_kind_all;              /* BAS_AZ_* of every direction when az_kind_dev is NULL */
    int flags;                    /* BAS_STEP_* */
    int reserved;
    long long n_valid, n_in;      /* samples per source / rounded up to a multiple of C */
    long long x_stride;           /* floats between sources in x_dev */
    long long p_begin, p_count;   /* rendered output range */
    long long out_stride;         /* floats between the ear rows of out_dev */
    const float* x_dev;
    const double* elev_dev;       /* n_src x (n_in/C + 1) directions (radians) */
    const double* azim_dev;
    const uint8_t* az_kind_dev;   /* or NULL */
    const double* diffs_left_dev; /* bank: delay tables, polyphase HRIRs, and the doubled layout for FUSED */
    const double* diffs_right_dev;
    const float* bank_pp_dev;
    const float* bank_pp2_dev;
    bas_term* terms_dev;          /* scratch: n_src x (n_in/C + 1) x 2 x BAS_MAX_TERMS */
    float* filt_dev;              /* scratch: n_src x (n_in/C + 1) filter rows (unused with BAS_STEP_FUSED) */
    const float* gains_dev;       /* or NULL */
    float* out_dev;
    int32_t* small_dev;           /* 2 + n_src words: status pair (bas_plan_build), then per-source peaks */
    void* workspace_dev;          /* bas_render workspace (may be NULL) */
    long long workspace_bytes;
    const bas_route* route;       /* mixing over several GPUs: route finished tiles to their owners (or NULL) */
} bas_step_job;
int bas_render_step(const bas_step_job* job, void* stream);

/* Scratch the tiled renderer may use to balance work across SMs: a tile split between two CTAs is
 * handed from one to the other through it (partial sums + a release/acquire flag per stripe, added
 * in a fixed order, so results stay deterministic).  workspace_dev may be NULL: tiles are then never
 * split.  Contents need no initialisation and carry nothing between calls; concurrent bas_render
 * calls (different streams) need a workspace each. */
long long bas_render_workspace_bytes(void);

/* Diagnostics, off by default: while trace_dev is non-NULL every CTA of the tiled render kernels launched by this
 * process stores {start ns, end ns, SM id, work items} (4 x u64 per CTA, room for 4 x 148 x 4 CTAs) - where does the
 * tail of a launch go (tools/cta_trace.py).  Pass NULL to switch it off again. */
int bas_render_set_trace(unsigned long long* trace_dev);

/* apply_hrtf.py:462-464: divide n floats by *peak_dev when it exceeds 1 (no-op otherwise). */
int bas_normalise(float* out_dev, long long n, const float* peak_dev, void* stream);

/* max over n floats of |v| into *peak_dev (atomic max; zero it first). */
int bas_peak(const float* v_dev, long long n, float* peak_dev, void* stream);

/* ---- host <-> HBM copies of the segment pipeline (make_signal_move_2d takes and returns host
 *      arrays, apply_hrtf.py:356, :459-466) ---------------------------------------------------
 * Asynchronous copy of `rows` rows of width_bytes each on `stream`; to_device != 0: host -> device,
 * else device -> host.  Pinned host memory is copied by DMA without staging. */
int bas_copy_2d(void* dst, long long dst_pitch_bytes, const void* src, long long src_pitch_bytes,
                long long width_bytes, long long rows, int to_device, void* stream);

/* Page-lock / release a caller-owned host array in place (cudaHostRegister): the input signal of
 * make_signal_move_2d (apply_hrtf.py:356) is an ordinary pageable ndarray; registered once, its
 * upload is a direct DMA.  Registering an already registered range, or releasing an unregistered
 * one, succeeds. */
int bas_host_register(void* host, long long bytes);
int bas_host_unregister(void* host);

/* ---- make_signal_move_2d with host buffers (apply_hrtf.py:356-466): upload, plan, ir_synth,
 *      segmented render and download as one pipeline over three streams ------------------------
 * The output range is cut into 1..8 PHASES along time (p_cuts[0..n_phases]).
 * bas_pipeline_upload enqueues the signal upload phase by phase (and the zero padding of
 * apply_hrtf.py:405-406) on stream_up and returns at once.  For every phase, in order, the caller
 * evaluates the trajectory at the phase's chunk boundaries into dirs_host (the reference's
 * elev_azim_function is host code, apply_hrtf.py:429/:435) and calls bas_pipeline_phase: it enqueues
 * the copy of directions [pt_begin, pt_end) of every source, their plan and filter rows, the render
 * of output samples [p_from, p_to) in time segments on stream_main, and the copy of each finished
 * segment to out_host on stream_down.  Only the last phase blocks: it returns when out_host and
 * small_host are complete.  All calls of a job must come from one host thread.  The peak division of
 * apply_hrtf.py:462-464 is NOT applied: the caller reads the peaks from small_host and, in the
 * rare case that one exceeds 1, runs bas_normalise on the arena's output block and copies again. */
typedef struct bas_pipeline_job {
    int n_src;                    /* mono sources of equal length */
    int C, S, K, U;               /* chunksize, subchunksize, taps, upsampling of the bank */
    int mix;                      /* 0: n_src x 2 x p_count result, 1: one 2 x p_count mix */
    int variant;                  /* bas_render variant */
    int az_kind_all;              /* BAS_AZ_* of every direction when az_kind_host is NULL */
    long long n;                  /* samples per source */
    long long n_in;               /* n rounded up to a multiple of C */
    long long p_begin, p_count;   /* rendered output range */
    long long x_host_stride;      /* floats between sources in x_host */
    long long segment_bytes;      /* target size of a downloaded segment (0: a single segment) */
    const float* x_host;          /* n_src x n (pinned for DMA without staging), or NULL with x_dev */
    const float* x_dev;           /* optional HBM-resident n_src x n_in signals (zero padded) */
    const double* dirs_host;      /* elev[n_src * n_pts] then azim[n_src * n_pts], n_pts = n_in / C + 1 */
    const uint8_t* az_kind_host;  /* n_src * n_pts bytes, or NULL: az_kind_all for every direction
                                     (then the directions travel inside the plan launches) */
    const double* diffs_left_dev; /* bank: delay tables and polyphase HRIRs (bas_bank_to_polyphase) */
    const double* diffs_right_dev;
    const float* bank_pp_dev;
    float* out_host;              /* [n_rows][2][p_count], n_rows = mix ? 1 : n_src (pinned) */
    int32_t* small_host;          /* 2 + n_src words (pinned): error bits, first failing direction,
                                     then the per-source peaks as float bits */
    void* arena_dev;              /* bas_pipeline_arena_bytes() bytes, 256-byte aligned */
    long long arena_bytes;
    void* workspace_dev;          /* bas_render workspace (may be NULL) */
    long long workspace_bytes;
    void* stream_main;            /* three distinct streams */
    void* stream_up;
    void* stream_down;
    const float* bank_pp2_dev;    /* bank in the bas_bank2_floats layout, or NULL.  When given and
                                     bas_render_fused_supported(C, S), no filter rows are written: the render
                                     kernel synthesises them (bas_render_fused) */
} bas_pipeline_job;

/* Bytes of device scratch a job needs.  offsets (may be NULL) receives the byte offsets of
 * {directions, az kinds, plan terms, filter rows, status+peaks, signals, output} and, in
 * offsets[7], the output row stride in floats (out[row][ear][stride]). */
long long bas_pipeline_arena_bytes(int n_src, long long n_in, int C, int K, int mix, long long p_count,
                                   int resident_x, long long* offsets);
int bas_pipeline_upload(const bas_pipeline_job* job, int n_phases, const long long* p_cuts);
int bas_pipeline_phase(const bas_pipeline_job* job, int phase, int n_phases, long long pt_begin, long long pt_end,
                       long long p_from, long long p_to);

/* Debug aid: returns (in buf) the timeline of the pipeline calls made on this thread since tracing was
 * enabled - host time of every enqueue, device time of its completion - then clears it and switches
 * tracing on or off. */
int bas_pipeline_trace(int enable, char* buf, size_t len);

/* Asynchronous byte fill of device memory on `stream` (zeroing peaks before bas_render). */
int bas_memset(void* dev, int value, long long bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BAS_B200_H */
