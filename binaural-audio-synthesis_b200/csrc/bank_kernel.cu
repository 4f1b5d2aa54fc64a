// K0  bank upload: the device half of load_irs_and_delaydiffs (apply_hrtf.py:23-46).
//
// The reference keeps each ear's upsampled HRIRs as a 187 x L float64 matrix (L = samples_to_keep * U,
// apply_hrtf.py:43-44).  Every consumer on the hot path reads a row at stride U after a shift
// (delay_signal_float with downsample=U, apply_hrtf.py:160-163), so the bank is stored in HBM in
// POLYPHASE order as fp32:
//
//     pp[row][n % U][n / U] = (float) irs[row][n]
//
// A gather term "bank[row][(m*U - s) mod L], m = 0..K-1" then becomes a contiguous (circularly
// rotated) run of K floats of phase row (-s mod U): fully coalesced.
#include "bas_internal.cuh"

__global__ void __launch_bounds__(256)
bas_polyphase_kernel(const double* __restrict__ irs, int L, int U, int K, float* __restrict__ pp) {
    const int row = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;       // coalesced read of the source row
    if (n >= L) return;
    const double v = irs[(size_t)row * L + n];
    pp[(size_t)row * L + (size_t)(n % U) * K + n / U] = (float)v;
}

extern "C" int bas_bank_to_polyphase(const double* irs_dev, int n_rows, int L, int U, float* out_dev, void* stream) {
    BAS_CHECK_ARG(irs_dev && out_dev, "null pointer");
    BAS_CHECK_ARG(n_rows >= 1 && n_rows <= 65535, "n_rows");
    BAS_CHECK_ARG(U >= 1 && L >= U && L % U == 0 && L < (1 << 20), "need 1 <= U, U | L, L < 2^20");
    const int threads = 256;
    dim3 grid((unsigned)bas_ceil_div(L, threads), (unsigned)n_rows);
    bas_polyphase_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(irs_dev, L, U, L / U, out_dev);
    BAS_LAUNCH_CHECK();
    return 0;
}
