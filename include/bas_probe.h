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

#ifdef __cplusplus
}
#endif
#endif /* BAS_PROBE_H */
