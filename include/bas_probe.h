/* bas_probe.h - measurement kernels (libbas_probe.so).  NOT part of the product library: nothing in
 * the package loads it; bench.py and tools/ use it to state measured pipe ceilings beside the
 * render kernel's achieved rate. */
#ifndef BAS_PROBE_H
#define BAS_PROBE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int bas_last_error(char* buf, size_t len);

/* FP32 pipe probe used by bench.py to state the measured FMA peak beside the HBM roofline:
 * every thread runs `iters` rounds of 16 independent dependent-chain FMAs.
 * packed=0: fma.rn.f32 (32 FMA per thread per round) ; packed=1: fma.rn.f32x2 (same FMA count).
 * sink_dev receives one float per thread so the work cannot be elided.  FMA count of the launch =
 * blocks * threads * iters * 32. */
int bas_probe_fma(int packed, int blocks, int threads, int iters, float* sink_dev, void* stream);

/* bas_probe_fma (packed = 0 / 1) that also reports the SM clock (MHz) the stream ran at, measured on the
 * device (clock64 against the nanosecond global timer): what the FMA peak has to be read against. */
int bas_probe_clock(int packed, int blocks, int threads, int iters, float* sink_dev, float* mhz_dev, void* stream);

/* The render kernel's own 32x32 block on synthetic shared-memory data, without the tile machinery:
 * `blocks` CTAs of 4 warps, every warp runs iters x 6 blocks of 1024 useful packed FMAs per lane.
 * ctas_per_sm (1..3) selects the register budget the block is compiled for. */
int bas_probe_block(int ctas_per_sm, int blocks, int iters, float* sink_dev, void* stream);

/* FEASIBILITY PROBE (tc_probe.cu): the chunk / subchunk FIR of make_signal_move_2d (apply_hrtf.py:431-453) for ONE source as a
 * Toeplitz contraction on tcgen05 (kind::tf32, 3 x TF32 split, accumulators in TMEM), the Toeplitz operand addressed through
 * overlapping shared-memory descriptors instead of being materialised.  x_dev: n_in samples (multiple of C = 512); filt_dev: the
 * (n_in / C + 1) filter rows of bas_ir_synth(BAS_IR_ROWS); out_dev: 2 x out_stride floats, ZEROED by the caller.
 * mode 0: correct end to end (global atomics) - accuracy; mode 1: MMAs + TMEM loads + register overlap-add, nothing stored -
 * rate proxy; mode 2: MMAs only.  blocks: CTAs (one per SM).  K <= 258. */
int bas_probe_tc_render(const float* x_dev, long long n_in, const float* filt_dev, int K, int C, float* out_dev,
                        long long out_stride, long long n_out, int mode, int blocks, float* sink_dev, void* stream);

/* One-thread kernel that stores %globaltimer (ns) into *slot_dev: a time stamp in stream order. */
int bas_probe_stamp(unsigned long long* slot_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BAS_PROBE_H */
