// K2  ir_synth: the array part of interpolate_2d (apply_hrtf.py:219-281) and of
// delay_compensated_interpolation_with_delaydiff (apply_hrtf.py:86-102).
//
// The reference builds one interpolated HRIR by chaining twelve delay_signal_float calls
// (apply_hrtf.py:127-165: two np.roll's over a length-L row each) and three linear blends, and only
// then throws away U-1 of every U samples (apply_hrtf.py:160-163).  All of those steps are linear
// and circular on length-L rows, so the decimated result is a weighted gather
//
//     out_e[m] = sum_t w_t * bank_e[row_t][(m*U - s_t) mod L],     m = 0..K-1
//
// with at most 16 distinct (row, shift) pairs per ear after merging equal ones (plan_math.h).  On
// the polyphase bank (bank_kernel.cu) term t reads phase row (-s_t mod U) starting at column
// (m - ceil(s_t / U)) mod K: K contiguous floats, one coalesced 128-byte line per warp load.  Only
// the K surviving samples are ever formed: work and traffic scale with K, not with L = K*U.
//
// Bound: L2 bandwidth (the bank, <= 12.25 MB, is L2 resident): 2*16*K*4 bytes read per point.
#include "bas_internal.cuh"

struct BasTermDev { int32_t row_shift; float weight; };

namespace {

constexpr int kThreads = 128;
constexpr int kMaxTerms = BAS_MAX_TERMS;

// decimated output: one CTA per (point, ear)
__global__ void __launch_bounds__(kThreads)
bas_ir_synth_kernel(const float* __restrict__ bank_pp, int U, int K, const BasTermDev* __restrict__ terms,
                    float* __restrict__ out, long long out_stride) {
    bas_grid_launch_dependents();
    bas_grid_dependency_wait();
    const long long point = blockIdx.x;
    const int ear = blockIdx.y;
    const int L = U * K;
    __shared__ int s_base[kMaxTerms];     // float offset of the phase row inside the ear's bank
    __shared__ int s_adv[kMaxTerms];      // column advance a = ceil(shift / U)
    __shared__ float s_w[kMaxTerms];
    __shared__ int s_n;
    if (threadIdx.x < 32) {
        // compact the non-zero terms (warp 0), so the inner loop is as short as the plan allows
        BasTermDev t; t.row_shift = 0; t.weight = 0.f;
        if (threadIdx.x < kMaxTerms) t = terms[(point * 2 + ear) * kMaxTerms + threadIdx.x];
        const bool live = t.weight != 0.f;
        const unsigned mask = __ballot_sync(0xffffffffu, live);
        if (live) {
            const int slot = __popc(mask & ((1u << threadIdx.x) - 1u));
            const int row = t.row_shift >> 20, shift = t.row_shift & 0xFFFFF;
            const int ph = (U - shift % U) % U;
            s_base[slot] = row * L + ph * K;
            s_adv[slot] = (shift + ph) / U;
            s_w[slot] = t.weight;
        }
        if (threadIdx.x == 0) s_n = __popc(mask);
    }
    __syncthreads();
    const int n_terms = s_n;
    const float* bank = bank_pp + (size_t)ear * BAS_N_DIRECTIONS * L;
    float* dst = out + (point * 2 + ear) * out_stride;
    for (int m = threadIdx.x; m < out_stride; m += kThreads) {
        float acc = 0.f;
        if (m < K) {
            for (int t = 0; t < n_terms; ++t) {
                int j = m - s_adv[t];
                j += (j < 0) ? K : 0;
                acc = fmaf(s_w[t], __ldg(bank + s_base[t] + j), acc);
            }
        }
        dst[m] = acc;                      // columns K..out_stride-1 are zero padding
    }
}

// filter-row layout for the renderer: both ears per thread, out[point][m][ear] with taps K..pitch-1
// zeroed (bas_filter_row_pitch).  One CTA per trajectory point.  All 2 x 16 gathers of a tap are
// issued as independent loads before the first is consumed (measured: the kernel then moves its 32 KB
// per point at the L2's ~6.3 TB/s whatever the cache policy of the loads): terms stay in their fixed plan slots, a slot with weight zero reads column 0 of
// the bank with weight zero (adds exactly nothing), and the sum runs over the slots in order.
__global__ void __launch_bounds__(kThreads)
bas_ir_synth_rows_kernel(const float* __restrict__ bank_pp, int U, int K, int pitch, long long n_points,
                         const BasTermDev* __restrict__ terms, float2* __restrict__ out) {
    bas_grid_launch_dependents();
    bas_grid_dependency_wait();
    const int L = U * K;
    __shared__ int s_base[2 * kMaxTerms];        // float offset of the phase row inside the bank (both ears)
    __shared__ int s_adv[2 * kMaxTerms];         // column advance a = ceil(shift / U)
    __shared__ float s_w[2 * kMaxTerms];
    const long long point = blockIdx.x;
    if (threadIdx.x < 2 * kMaxTerms) {
        const int ear = threadIdx.x / kMaxTerms;
        const BasTermDev t = terms[point * 2 * kMaxTerms + threadIdx.x];
        const bool live = t.weight != 0.f;
        const int row = t.row_shift >> 20, shift = t.row_shift & 0xFFFFF;
        const int ph = (U - shift % U) % U;
        s_base[threadIdx.x] = live ? (ear * BAS_N_DIRECTIONS + row) * L + ph * K : 0;
        s_adv[threadIdx.x] = live ? (shift + ph) / U : 0;
        s_w[threadIdx.x] = t.weight;
    }
    __syncthreads();
    float2* dst = out + point * pitch;
    for (int m = threadIdx.x; m < pitch; m += kThreads) {
        float acc_l = 0.f, acc_r = 0.f;
        if (m < K) {
            float v[2 * kMaxTerms];
#pragma unroll
            for (int t = 0; t < 2 * kMaxTerms; ++t) {
                int j = m - s_adv[t];
                j += (j < 0) ? K : 0;
                v[t] = __ldg(bank_pp + s_base[t] + j);
            }
#pragma unroll
            for (int t = 0; t < kMaxTerms; ++t) {
                acc_l = fmaf(s_w[t], v[t], acc_l);
                acc_r = fmaf(s_w[kMaxTerms + t], v[kMaxTerms + t], acc_r);
            }
        }
        dst[m] = make_float2(acc_l, acc_r);
    }
}

// return_upsampled=True: all L samples (apply_hrtf.py:97-99), used only by the ring entry point
__global__ void __launch_bounds__(kThreads)
bas_ir_synth_full_kernel(const float* __restrict__ bank_pp, int U, int K, const BasTermDev* __restrict__ terms,
                         float* __restrict__ out, long long out_stride) {
    bas_grid_launch_dependents();
    bas_grid_dependency_wait();
    const long long point = blockIdx.y;
    const int ear = blockIdx.z;
    const int L = U * K;
    const int n = blockIdx.x * kThreads + threadIdx.x;
    if (n >= out_stride) return;
    const BasTermDev* tp = terms + (point * 2 + ear) * kMaxTerms;
    const float* bank = bank_pp + (size_t)ear * BAS_N_DIRECTIONS * L;
    float acc = 0.f;
    if (n < L) {
        for (int t = 0; t < kMaxTerms; ++t) {
            const BasTermDev term = tp[t];
            if (term.weight == 0.f) continue;
            const int row = term.row_shift >> 20, shift = term.row_shift & 0xFFFFF;
            int idx = n - shift;
            idx += (idx < 0) ? L : 0;
            acc = fmaf(term.weight, __ldg(bank + (size_t)row * L + (idx % U) * K + idx / U), acc);
        }
    }
    out[(point * 2 + ear) * out_stride + n] = acc;
}

}  // namespace

extern "C" int bas_filter_row_pitch(int K) { return K < 1 ? BAS_E_ARG : (K + 31) / 32 * 32 + 2; }

extern "C" int bas_ir_synth(const float* bank_pp_dev, int U, int K, const bas_term* terms_dev, long long n_points,
                            int mode, float* out_dev, long long out_stride, void* stream) {
    BAS_CHECK_ARG(bank_pp_dev && terms_dev && out_dev, "null pointer");
    BAS_CHECK_ARG(U >= 1 && K >= 1 && (long long)U * K < (1 << 20), "need U, K >= 1 and U*K < 2^20");
    BAS_CHECK_ARG(n_points >= 0 && n_points < 0x7fffffffLL, "n_points");
    BAS_CHECK_ARG(mode == BAS_IR_UPSAMPLED || mode == BAS_IR_PLANAR || mode == BAS_IR_ROWS, "mode");
    if (mode != BAS_IR_ROWS)
        BAS_CHECK_ARG(out_stride >= (mode == BAS_IR_PLANAR ? K : U * K) && out_stride < 0x7fffffffLL, "out_stride too small");
    if (n_points == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const BasTermDev* terms = reinterpret_cast<const BasTermDev*>(terms_dev);
    if (mode == BAS_IR_ROWS) {
        BAS_CHECK_ARG((reinterpret_cast<uintptr_t>(out_dev) & 15) == 0, "filter rows must be 16-byte aligned");
        BAS_CUDA(bas_launch(bas_ir_synth_rows_kernel, dim3((unsigned)n_points), dim3(kThreads), 0, st,
                            bank_pp_dev, U, K, bas_filter_row_pitch(K), n_points, terms, reinterpret_cast<float2*>(out_dev)));
    } else if (mode == BAS_IR_PLANAR) {
        dim3 grid((unsigned)n_points, 2);
        BAS_CUDA(bas_launch(bas_ir_synth_kernel, grid, dim3(kThreads), 0, st, bank_pp_dev, U, K, terms, out_dev, out_stride));
    } else {
        BAS_CHECK_ARG(n_points <= 65535, "n_points <= 65535 when mode is BAS_IR_UPSAMPLED");
        dim3 grid((unsigned)bas_ceil_div(out_stride, kThreads), (unsigned)n_points, 2);
        BAS_CUDA(bas_launch(bas_ir_synth_full_kernel, grid, dim3(kThreads), 0, st, bank_pp_dev, U, K, terms, out_dev, out_stride));
    }
    BAS_LAUNCH_CHECK();
    return 0;
}
