// Bank builder: the offline preprocessing of upsample_irs.m on the device.
//
//   irs_*(i, :)      = resample(hrir(i, :), U, 1)                                  upsample_irs.m:42-43
//   diffs_*(i, j)    = delaydifference(hrir(i, :), hrir(j, :), U), i < j            :22-28, :59-77
//                      then diffs - diffs'                                          :31-32
//
// resample(x, U, 1) is a zero-phase polyphase interpolation with an odd-length FIR h (2 Lh + 1 taps,
// designed on the host: bank_builder.py restates the filter design):
//     y[n] = sum_k h[Lh + n - k U] x[k],      n = 0 .. len(x) U - 1.
// delaydifference cross-correlates two HRIRs (fftconv(fliplr(a), b), :68), resamples the correlation
// the same way, takes the first maximum, refines it with a parabola through its neighbours (:88-101)
// and re-centres on zero lag (:70-76).  Everything is fp64: the delay tables feed floor()/ceil().
#include "bas_internal.cuh"

namespace {

// y[row][n] for one ear: one thread per output sample
__global__ void __launch_bounds__(256)
bas_upsample_kernel(const double* __restrict__ x, int n_rows, int n, int U, const double* __restrict__ h, int Lh,
                    double* __restrict__ y) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_row = (long long)n * U;
    if (idx >= per_row * n_rows) return;
    const int row = (int)(idx / per_row);
    const int m = (int)(idx - (long long)row * per_row);
    // taps t = Lh + m - k U in [0, 2 Lh]  <=>  k in [ceil((m - Lh) / U), floor((m + Lh) / U)]
    int k_lo = m - Lh; k_lo = k_lo <= 0 ? 0 : (k_lo + U - 1) / U;
    int k_hi = (m + Lh) / U; if (k_hi > n - 1) k_hi = n - 1;
    const double* xr = x + (long long)row * n;
    double acc = 0.0;
    for (int k = k_lo; k <= k_hi; ++k) acc += h[Lh + m - k * U] * xr[k];
    y[idx] = acc;
}

// one CTA per pair (i < j): correlation in shared memory, upsampled on the fly, first maximum, parabola
constexpr int kPairThreads = 256;

__global__ void __launch_bounds__(kPairThreads)
bas_delay_diff_kernel(const double* __restrict__ x, int n_rows, int n, int U, const double* __restrict__ h, int Lh,
                      double* __restrict__ diffs) {
    extern __shared__ double sm[];
    double* a = sm;                    // n
    double* b = a + n;                 // n
    double* c = b + n;                 // 2n - 1 lags
    __shared__ double s_val[kPairThreads];
    __shared__ int s_idx[kPairThreads];
    // pair index -> (i, j), i < j, row-major over the upper triangle
    long long p = blockIdx.x;
    int i = 0;
    while (p >= n_rows - 1 - i) { p -= n_rows - 1 - i; ++i; }
    const int j = i + 1 + (int)p;
    for (int t = threadIdx.x; t < n; t += kPairThreads) { a[t] = x[(long long)i * n + t]; b[t] = x[(long long)j * n + t]; }
    __syncthreads();
    // c[l] = sum_t a[t] b[t + l - (n - 1)]: fftconv(fliplr(a), b), upsample_irs.m:68
    const int n_lags = 2 * n - 1;
    for (int l = threadIdx.x; l < n_lags; l += kPairThreads) {
        const int s = l - (n - 1);
        const int t0 = s < 0 ? -s : 0, t1 = s > 0 ? n - s : n;
        double acc = 0.0;
        for (int t = t0; t < t1; ++t) acc += a[t] * b[t + s];
        c[l] = acc;
    }
    __syncthreads();
    // upsampled correlation cu[m] = sum_k h[Lh + m - k U] c[k]; first maximum (Octave's max, :71)
    auto cu = [&](int m) {
        int k_lo = m - Lh; k_lo = k_lo <= 0 ? 0 : (k_lo + U - 1) / U;
        int k_hi = (m + Lh) / U; if (k_hi > n_lags - 1) k_hi = n_lags - 1;
        double acc = 0.0;
        for (int k = k_lo; k <= k_hi; ++k) acc += h[Lh + m - k * U] * c[k];
        return acc;
    };
    const int n_up = n_lags * U;
    double best = -1.0e308; int best_m = 0;
    for (int m = threadIdx.x; m < n_up; m += kPairThreads) {
        const double v = cu(m);
        if (v > best) { best = v; best_m = m; }              // ascending m per thread: keeps the first maximum
    }
    s_val[threadIdx.x] = best; s_idx[threadIdx.x] = best_m;
    __syncthreads();
    for (int off = kPairThreads / 2; off > 0; off >>= 1) {
        if (threadIdx.x < off) {
            const double v = s_val[threadIdx.x + off]; const int m = s_idx[threadIdx.x + off];
            if (v > s_val[threadIdx.x] || (v == s_val[threadIdx.x] && m < s_idx[threadIdx.x])) { s_val[threadIdx.x] = v; s_idx[threadIdx.x] = m; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int pk = s_idx[0];                               // 0-based index of the maximum
        double d = nan("");
        if (pk > 0 && pk < n_up - 1) {                         // parabolic_interpolation, :88-101
            const double v0 = cu(pk - 1), v1 = s_val[0], v2 = cu(pk + 1);
            const double pa = 0.5 * (v0 + v2 - 2.0 * v1), pb = 0.5 * (v2 - v0);
            const double frac = -pb / (2.0 * pa);
            // :72-76 with Octave's 1-based peak_index: (pk + 1 + frac - 1) / U - (n - 1)
            d = ((double)pk + 1.0 + frac - 1.0) / (double)U - (double)(n - 1);
        }
        diffs[(long long)i * n_rows + j] = d;                  // upper triangle (:22-28)
    }
}

// diffs = upper - upper'  (upsample_irs.m:31-32), in place
__global__ void bas_antisym_kernel(double* __restrict__ d, int n_rows) {
    const int i = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_rows || j <= i) return;
    const double v = d[(long long)i * n_rows + j];
    d[(long long)j * n_rows + i] = -v;          // the diagonal stays zero: the matrix is zero-filled first
}

}  // namespace

extern "C" int bas_bank_upsample(const double* hrir_dev, int n_rows, int n, int U, const double* h_dev, int n_taps,
                                 double* out_dev, void* stream) {
    BAS_CHECK_ARG(hrir_dev && h_dev && out_dev, "null pointer");
    BAS_CHECK_ARG(n_rows >= 1 && n >= 1 && U >= 1 && n_taps >= 1 && (n_taps & 1), "need rows, samples, U >= 1 and an odd filter length");
    const long long total = (long long)n_rows * n * U;
    BAS_CHECK_ARG(total < (1LL << 40), "too large");
    bas_upsample_kernel<<<(unsigned)bas_ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(hrir_dev, n_rows, n, U, h_dev, n_taps / 2, out_dev);
    BAS_LAUNCH_CHECK();
    return 0;
}

extern "C" int bas_bank_delay_diffs(const double* hrir_dev, int n_rows, int n, int U, const double* h_dev, int n_taps,
                                    double* diffs_dev, void* stream) {
    BAS_CHECK_ARG(hrir_dev && h_dev && diffs_dev, "null pointer");
    BAS_CHECK_ARG(n_rows >= 2 && n_rows <= 4096 && n >= 2 && n <= 2048 && U >= 1 && n_taps >= 1 && (n_taps & 1), "geometry");
    cudaStream_t st = (cudaStream_t)stream;
    BAS_CUDA(cudaMemsetAsync(diffs_dev, 0, (size_t)n_rows * n_rows * sizeof(double), st));
    const size_t smem = (size_t)(4 * n - 1) * sizeof(double);
    BAS_CUDA(cudaFuncSetAttribute(bas_delay_diff_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long pairs = (long long)n_rows * (n_rows - 1) / 2;
    bas_delay_diff_kernel<<<(unsigned)pairs, kPairThreads, smem, st>>>(hrir_dev, n_rows, n, U, h_dev, n_taps / 2, diffs_dev);
    BAS_LAUNCH_CHECK();
    dim3 grid((unsigned)bas_ceil_div(n_rows, 128), (unsigned)n_rows);
    bas_antisym_kernel<<<grid, 128, 0, st>>>(diffs_dev, n_rows);
    BAS_LAUNCH_CHECK();
    return 0;
}
