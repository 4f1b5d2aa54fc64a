// Host <-> HBM transfers of the render pipeline.  make_signal_move_2d takes a host signal and returns
// a host array (apply_hrtf.py:356, :459-466); the host side cuts the signal into time segments and
// overlaps upload, render and download of consecutive segments on three streams.  These are the
// copies it issues: plain asynchronous DMA, one call each (no batched-copy API).
#include "bas_internal.cuh"

extern "C" int bas_copy_2d(void* dst, long long dst_pitch_bytes, const void* src, long long src_pitch_bytes,
                           long long width_bytes, long long rows, int to_device, void* stream) {
    BAS_CHECK_ARG(dst && src, "null pointer");
    BAS_CHECK_ARG(width_bytes >= 0 && rows >= 0, "negative extent");
    if (width_bytes == 0 || rows == 0) return 0;
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    if (rows == 1 || (dst_pitch_bytes == width_bytes && src_pitch_bytes == width_bytes)) {
        BAS_CUDA(cudaMemcpyAsync(dst, src, (size_t)(width_bytes * rows), kind, (cudaStream_t)stream));
        return 0;
    }
    BAS_CHECK_ARG(dst_pitch_bytes >= width_bytes && src_pitch_bytes >= width_bytes, "pitch smaller than a row");
    BAS_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch_bytes, src, (size_t)src_pitch_bytes, (size_t)width_bytes,
                               (size_t)rows, kind, (cudaStream_t)stream));
    return 0;
}

extern "C" int bas_memset(void* dev, int value, long long bytes, void* stream) {
    BAS_CHECK_ARG(dev && bytes >= 0, "bad pointer or size");
    if (bytes) BAS_CUDA(cudaMemsetAsync(dev, value, (size_t)bytes, (cudaStream_t)stream));
    return 0;
}

// Page-lock a caller's host array in place so that the upload of make_signal_move_2d's input
// (apply_hrtf.py:356: in_signal is an ordinary, pageable ndarray) is a direct DMA instead of a
// staged copy.  The Python host keeps one registration per live array (dropped when the array dies).
extern "C" int bas_host_register(void* host, long long bytes) {
    BAS_CHECK_ARG(host && bytes > 0, "bad pointer or size");
    cudaError_t e = cudaHostRegister(host, (size_t)bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) {
        cudaGetLastError();
        bas_set_error("bas_host_register: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

extern "C" int bas_host_unregister(void* host) {
    BAS_CHECK_ARG(host, "null pointer");
    cudaError_t e = cudaHostUnregister(host);
    if (e != cudaSuccess && e != cudaErrorHostMemoryNotRegistered) {
        cudaGetLastError();
        bas_set_error("bas_host_unregister: %s", cudaGetErrorString(e));
        return (int)e;
    }
    cudaGetLastError();
    return 0;
}
