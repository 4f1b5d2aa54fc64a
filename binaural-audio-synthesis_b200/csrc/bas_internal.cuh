// Internal helpers shared by the translation units of libbas_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/bas_b200.h"

void bas_set_error(const char* fmt, ...);

#define BAS_CHECK_ARG(cond, msg)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            bas_set_error("%s: bad argument: %s", __func__, msg);  \
            return BAS_E_ARG;                                      \
        }                                                          \
    } while (0)

#define BAS_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            bas_set_error("%s: %s failed: %s", __func__, #call, cudaGetErrorString(e_)); \
            return (int)e_;                                                              \
        }                                                                                \
    } while (0)

#define BAS_LAUNCH_CHECK()                                                              \
    do {                                                                                \
        cudaError_t e_ = cudaGetLastError();                                            \
        if (e_ != cudaSuccess) {                                                        \
            bas_set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(e_)); \
            return (int)e_;                                                             \
        }                                                                               \
    } while (0)

int bas_plan_build_range(const double* diffs_left_dev, const double* diffs_right_dev, int U, int L,
                         const double* elev_dev, const double* azim_dev, const uint8_t* az_kind_dev,
                         int az_kind_all, long long n_points, bas_term* terms_dev, bas_trace* trace_dev,
                         int* status_dev, long long point_offset, int reset_status, void* stream);

int bas_plan_build_runs(const double* diffs_left_dev, const double* diffs_right_dev, int U, int L,
                        const double* elev_dev, const double* azim_dev, const uint8_t* az_kind_dev,
                        int az_kind_all, long long rows, long long run, long long run_stride, bas_term* terms_dev, bas_trace* trace_dev,
                        int* status_dev, long long point_offset, int reset_status, void* stream);

int bas_plan_build_inline(const double* diffs_left_dev, const double* diffs_right_dev, int U, int L,
                          const double* elev_host, const double* azim_host, int az_kind_all, long long n_points,
                          bas_term* terms_dev, int* status_dev, long long point_offset, void* stream);

static inline long long bas_ceil_div(long long a, long long b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch ---------------------------------------------------------------
// bas_render_step chains plan -> ir_synth -> render -> normalise on one stream.  Inside such a chain a
// kernel is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may become resident
// while the previous kernel drains (launch latency, barrier / table set-up overlap its tail), and
// bas_grid_dependency_wait() - griddepcontrol.wait, the first thing each kernel does before it touches
// anything an earlier kernel wrote or still reads - holds them until that kernel has completed and
// flushed.  Outside a chain (the flag is off) kernels launch with full stream serialization and the
// wait returns at once.  The flag is per host thread; bas_render_step sets it around its launches.
bool& bas_pdl_flag();
struct BasPdlScope {
    bool saved;
    explicit BasPdlScope(bool on) : saved(bas_pdl_flag()) { bas_pdl_flag() = on; }
    ~BasPdlScope() { bas_pdl_flag() = saved; }
};

#ifdef __CUDACC__
__device__ __forceinline__ void bas_grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void bas_grid_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t bas_launch(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = bas_pdl_flag() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#endif
