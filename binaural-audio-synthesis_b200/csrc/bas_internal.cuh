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

int bas_plan_build_inline(const double* diffs_left_dev, const double* diffs_right_dev, int U, int L,
                          const double* elev_host, const double* azim_host, int az_kind_all, long long n_points,
                          bas_term* terms_dev, int* status_dev, long long point_offset, void* stream);

static inline long long bas_ceil_div(long long a, long long b) { return (a + b - 1) / b; }
