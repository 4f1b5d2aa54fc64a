// Mix reduction over peer memory (NVLink / NVSwitch), fused with the render kernel.
//
// By-source sharding: every rank mixes its own sources; the per-rank mixes (2, N_out) must be
// summed.  Instead of a collective after the render, the render kernel's epilogue ROUTES every finished tile of
// the local mix straight into the receive buffer of the rank that owns that stretch of the output (plain
// 16-byte stores to peer-mapped memory: the transfer overlaps the FIR math tile by tile, render_tiled.cuh).
// What is left for the end of the step are three small kernels on each rank:
//
//   bas_peer_signal   "my partial tiles have landed everywhere"   one release-store per peer (or folded into
//                     bas_peer_reduce: arrive_ptrs_dev)
//   bas_peer_reduce   wait for every writer's signal; sum the N partial slices in RANK ORDER (deterministic,
//                     unlike a ring or tree); store the reduced slice into the result buffer of n_results ranks
//                     (every rank: the mix replicated, an all-reduce; only this rank: the mix left sharded by time,
//                     a reduce-scatter); the last CTA signals "slice of owner o done" to every peer
//   bas_peer_wait     wait for every owner's signal.  Replicated: at the end of the step - the full mix is now in this
//                     rank's result buffer.  Sharded: at the START of the next step, before the routed render writes
//                     into the owners' receive buffers again (flow control, off the critical path)
//
// Memory: caller-owned symmetric buffers (torch.distributed._symmetric_memory supplies the peer mappings;
// this library never allocates).  All waits are bounded by elapsed time (trap instead of hanging the GPU).
#include "bas_internal.cuh"
#include <cuda.h>

namespace {

__device__ __forceinline__ unsigned long long peer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
constexpr unsigned long long kPeerWaitNs = 20ull * 1000 * 1000 * 1000;

// flags count steps: a flag "has reached" epoch e when (int)(flag - e) >= 0 (wrap-safe)
__device__ __forceinline__ void wait_reached(const unsigned* flag, unsigned epoch) {
    if ((int)(ld_acquire_sys(flag) - epoch) >= 0) return;
    const unsigned long long t0 = peer_ns();
    for (unsigned spins = 1; (int)(ld_acquire_sys(flag) - epoch) < 0; ++spins) {
        __nanosleep(100);
        if ((spins & 255u) == 0 && peer_ns() - t0 > kPeerWaitNs) __trap();      // a peer never arrived: fail loudly
    }
}

__global__ void bas_peer_signal_kernel(unsigned* const* __restrict__ flag_ptrs, int n, int slot, unsigned epoch) {
    bas_grid_launch_dependents();
    bas_grid_dependency_wait();
    __threadfence_system();
    if ((int)threadIdx.x < n) st_release_sys(flag_ptrs[threadIdx.x] + slot, epoch);
}

__global__ void bas_peer_wait_kernel(const unsigned* __restrict__ flags, int n, unsigned epoch) {
    bas_grid_launch_dependents();
    bas_grid_dependency_wait();
    if ((int)threadIdx.x < n) wait_reached(flags + threadIdx.x, epoch);
}

// recv: this rank's receive buffer [writer][ear][stride]; sums over writers in rank order the `valid` outputs of this
// rank's slice and stores them at result[ear][slice_begin + p] of every rank
__global__ void __launch_bounds__(256)
bas_peer_reduce_kernel(const float* __restrict__ recv, int n, long long stride, long long valid, float* const* __restrict__ result_ptrs,
                       int n_results, long long result_stride, long long slice_begin, const unsigned* __restrict__ arrived, unsigned epoch,
                       unsigned* const* __restrict__ done_ptrs, int rank, unsigned* __restrict__ counter,
                       unsigned* const* __restrict__ arrive_ptrs) {
    bas_grid_launch_dependents();
    bas_grid_dependency_wait();
    // folded bas_peer_signal: this rank's render (the previous kernel in the stream) has completed
    if (arrive_ptrs && blockIdx.x == 0) {
        __threadfence_system();
        if ((int)threadIdx.x < n) st_release_sys(arrive_ptrs[threadIdx.x] + rank, epoch);
    }
    if ((int)threadIdx.x < n) wait_reached(arrived + threadIdx.x, epoch);
    __syncthreads();
    const long long quads = (valid + 3) / 4;                           // stride and slice_begin are multiples of 4
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < 2 * quads; q += (long long)gridDim.x * blockDim.x) {
        const int ear = q >= quads;
        const long long p = (q - ear * quads) * 4;
        float4 acc = *reinterpret_cast<const float4*>(recv + (long long)ear * stride + p);
        for (int w = 1; w < n; ++w) {
            const float4 v = *reinterpret_cast<const float4*>(recv + ((long long)w * 2 + ear) * stride + p);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        for (int r = 0; r < n_results; ++r)
            *reinterpret_cast<float4*>(result_ptrs[r] + (long long)ear * result_stride + slice_begin + p) = acc;
    }
    // the last CTA to finish tells every peer that this owner's slice is complete
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *counter = 0;                            // ready for the next step
        __threadfence_system();
        if ((int)threadIdx.x < n) st_release_sys(done_ptrs[threadIdx.x] + rank, epoch);
    }
}

}  // namespace

extern "C" int bas_peer_signal(unsigned* const* flag_ptrs_dev, int n, int slot, unsigned epoch, void* stream) {
    BAS_CHECK_ARG(flag_ptrs_dev && n >= 1 && n <= 32 && slot >= 0, "bad arguments");
    BAS_CUDA(bas_launch(bas_peer_signal_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, flag_ptrs_dev, n, slot, epoch));
    return 0;
}

extern "C" int bas_peer_wait(const unsigned* flags_dev, int n, unsigned epoch, void* stream) {
    BAS_CHECK_ARG(flags_dev && n >= 1 && n <= 32, "bad arguments");
    BAS_CUDA(bas_launch(bas_peer_wait_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, flags_dev, n, epoch));
    return 0;
}

// Stream-ordered wait without a resident kernel: n cuStreamWaitValue32(GEQ) operations - the stream's front end
// polls the flags, no SM is held while a peer is late (a spinning bas_peer_wait / bas_peer_reduce CTA would keep a
// persistent render CTA of the next step off its SM).  GEQ is the same wrap-safe comparison as wait_reached.
typedef CUresult (*StreamWaitValue32Fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamWaitValue32Fn stream_wait_fn() {
    static StreamWaitValue32Fn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (getenv("BAS_NO_STREAM_WAIT")) return (StreamWaitValue32Fn) nullptr;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return (StreamWaitValue32Fn) nullptr;
        }
        return (StreamWaitValue32Fn)p;
    }();
    return fn;
}

extern "C" int bas_peer_stream_wait(const unsigned* flags_dev, int n, unsigned epoch, void* stream) {
    BAS_CHECK_ARG(flags_dev && n >= 1 && n <= 32, "bad arguments");
    StreamWaitValue32Fn fn = stream_wait_fn();
    if (!fn) { bas_set_error("bas_peer_stream_wait: cuStreamWaitValue32 is not available"); return BAS_E_UNSUPPORTED; }
    for (int w = 0; w < n; ++w) {
        const CUresult r = fn((CUstream)stream, (CUdeviceptr)(uintptr_t)(flags_dev + w), epoch, CU_STREAM_WAIT_VALUE_GEQ);
        if (r != CUDA_SUCCESS) { bas_set_error("bas_peer_stream_wait: cuStreamWaitValue32 failed (%d)", (int)r); return BAS_E_UNSUPPORTED; }
    }
    return 0;
}

extern "C" int bas_peer_reduce(const float* recv_dev, int n, long long stride, long long valid, float* const* result_ptrs_dev, int n_results,
                               long long result_stride, long long slice_begin, const unsigned* arrived_dev, unsigned epoch,
                               unsigned* const* done_ptrs_dev, int rank, unsigned* counter_dev, unsigned* const* arrive_ptrs_dev,
                               void* stream) {
    BAS_CHECK_ARG(recv_dev && result_ptrs_dev && arrived_dev && done_ptrs_dev && counter_dev, "null pointer");
    BAS_CHECK_ARG(n >= 1 && n <= 32 && rank >= 0 && rank < n && n_results >= 1 && n_results <= n, "ranks");
    BAS_CHECK_ARG(valid >= 0 && stride >= valid && stride % 4 == 0 && slice_begin % 4 == 0 && result_stride % 4 == 0, "geometry (multiples of 4 floats)");
    const long long quads = 2 * ((valid + 3) / 4);
    long long blocks = bas_ceil_div(quads > 0 ? quads : 1, 256 * 4);
    if (blocks > 148 * 4) blocks = 148 * 4;
    BAS_CUDA(bas_launch(bas_peer_reduce_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, recv_dev, n, stride, valid,
                        result_ptrs_dev, n_results, result_stride, slice_begin, arrived_dev, epoch, done_ptrs_dev, rank, counter_dev, arrive_ptrs_dev));
    return 0;
}
