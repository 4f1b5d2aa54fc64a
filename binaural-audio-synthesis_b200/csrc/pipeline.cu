// Host-buffer form of make_signal_move_2d (apply_hrtf.py:356-466): mono host signals in, a host
// float32 result out, everything in between on the device.
//
//   upload   (stream_up)    x_host -> HBM phase by phase, zero padding to n_in (apply_hrtf.py:405-406)
//   phase i  (stream_main)  directions of the phase -> plan_build -> ir_synth -> [wait upload i] ->
//                           render, cut into time segments
//            (stream_down)  each finished segment -> out_host while the next one renders
//
// The job is cut into a few PHASES along time.  The caller evaluates the trajectory of phase i on the
// host (the reference's elev_azim_function is host code, apply_hrtf.py:429/:435), calls
// bas_pipeline_phase(i) - which only enqueues work - and goes on to evaluate phase i+1 while phase i
// renders and travels back.  PCIe is the long pole of the call (8 B per output pair going back); the
// pipeline gets the device -> host direction busy as early as possible and keeps it busy.  No memory is allocated here: the caller passes a device arena of
// bas_pipeline_arena_bytes() bytes (layout below) and pinned host buffers.
#include "bas_internal.cuh"

#include <chrono>
#include <string>
#include <vector>

namespace {

// ---- optional timeline (tools/e2e_timeline.py): host time of every enqueue and device time of its
//      completion.  Off unless bas_pipeline_trace(1) was called on this thread. ---------------------
struct Mark { std::string label; double host_us; cudaEvent_t ev; };
thread_local bool g_trace = false;
thread_local std::vector<Mark> g_marks;

void mark(const char* what, long long a, long long b, cudaStream_t st, bool host_only = false) {
    if (!g_trace) return;
    Mark m;
    char buf[64];
    snprintf(buf, sizeof(buf), what, a, b);
    m.label = buf;
    m.host_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count();
    m.ev = nullptr;
    if (!host_only && cudaEventCreate(&m.ev) == cudaSuccess) cudaEventRecord(m.ev, st);
    g_marks.push_back(m);
}

constexpr long long kAlign = 256;
constexpr long long kSegAlign = 8192;          // output samples: whole render tiles for every tile width

inline long long up_to(long long v, long long a) { return (v + a - 1) / a * a; }

struct Layout {
    long long dirs, kinds, terms, filt, small, x, out, total;
    long long n_pts, n_dirs, stride;
    int pitch, n_rows;
};

Layout layout_of(int n_src, long long n_in, int C, int K, int mix, long long p_count, int resident_x) {
    Layout l;
    l.n_pts = n_in / C + 1;
    l.n_dirs = l.n_pts * n_src;
    l.pitch = bas_filter_row_pitch(K);
    l.n_rows = mix ? 1 : n_src;
    l.stride = up_to(p_count > 0 ? p_count : 1, 4);
    long long o = 0;
    l.dirs = o;  o += up_to(l.n_dirs * 16, kAlign);
    l.kinds = o; o += up_to(l.n_dirs, kAlign);
    l.terms = o; o += up_to(l.n_dirs * 2 * BAS_MAX_TERMS * (long long)sizeof(bas_term), kAlign);
    l.filt = o;  o += up_to(l.n_dirs * l.pitch * 8, kAlign);
    l.small = o; o += up_to((2 + n_src) * 4LL, kAlign);
    l.x = o;     o += resident_x ? 0 : up_to((long long)n_src * n_in * 4, kAlign);
    l.out = o;   o += up_to(l.n_rows * 2 * l.stride * 4, kAlign);
    l.total = o;
    return l;
}

// cross-stream edges of the pipeline: one event set per (thread, device)
constexpr int kMaxPhases = 8;
struct Events { cudaEvent_t upload[kMaxPhases] = {}, segment = nullptr; };
Events* events_for_current_device() {
    static thread_local Events ev[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    Events& e = ev[dev];
    if (!e.segment) {
        for (int i = 0; i < kMaxPhases; ++i)
            if (cudaEventCreateWithFlags(&e.upload[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&e.segment, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    }
    return &e;
}

int check_job(const bas_pipeline_job* j) {
    BAS_CHECK_ARG(j, "null job");
    BAS_CHECK_ARG(j->n_src >= 1 && j->n >= 1 && j->C >= 1 && j->S >= 1 && j->C % j->S == 0, "geometry");
    BAS_CHECK_ARG(j->n_in >= j->n && j->n_in % j->C == 0 && j->n_in - j->n < j->C, "n_in must be n rounded up to a multiple of C");
    BAS_CHECK_ARG(j->K >= 1 && j->U >= 1, "K, U");
    BAS_CHECK_ARG(j->p_begin >= 0 && j->p_count >= 0 && j->p_begin + j->p_count <= j->n_in + j->K - 1, "output range");
    BAS_CHECK_ARG(j->arena_dev && (reinterpret_cast<uintptr_t>(j->arena_dev) & 255) == 0, "arena must be 256-byte aligned");
    const Layout l = layout_of(j->n_src, j->n_in, j->C, j->K, j->mix, j->p_count, j->x_dev != nullptr);
    BAS_CHECK_ARG(j->arena_bytes >= l.total, "arena too small (bas_pipeline_arena_bytes)");
    BAS_CHECK_ARG(j->x_dev || (j->x_host && j->x_host_stride >= j->n), "signals");
    return 0;
}

}  // namespace

extern "C" long long bas_pipeline_arena_bytes(int n_src, long long n_in, int C, int K, int mix, long long p_count,
                                              int resident_x, long long* offsets) {
    if (n_src < 1 || n_in < 1 || C < 1 || K < 1 || p_count < 0 || n_in % C) return BAS_E_ARG;
    const Layout l = layout_of(n_src, n_in, C, K, mix, p_count, resident_x);
    if (offsets) {
        offsets[0] = l.dirs; offsets[1] = l.kinds; offsets[2] = l.terms; offsets[3] = l.filt;
        offsets[4] = l.small; offsets[5] = l.x; offsets[6] = l.out; offsets[7] = l.stride;
    }
    return l.total;
}

extern "C" int bas_pipeline_upload(const bas_pipeline_job* j, int n_phases, const long long* p_cuts) {
    if (int rc = check_job(j)) return rc;
    BAS_CHECK_ARG(n_phases >= 1 && n_phases <= kMaxPhases && p_cuts, "1..8 phases");
    BAS_CHECK_ARG(p_cuts[0] == j->p_begin && p_cuts[n_phases] == j->p_begin + j->p_count, "phase cuts must span the output range");
    for (int i = 0; i < n_phases; ++i) BAS_CHECK_ARG(p_cuts[i] <= p_cuts[i + 1], "phase cuts must ascend");
    if (j->x_dev || j->p_count == 0) return 0;
    const Layout l = layout_of(j->n_src, j->n_in, j->C, j->K, j->mix, j->p_count, 0);
    Events* ev = events_for_current_device();
    BAS_CHECK_ARG(ev, "cannot create events on the current device");
    cudaStream_t up = (cudaStream_t)j->stream_up;
    float* x = reinterpret_cast<float*>(static_cast<char*>(j->arena_dev) + l.x);
    // the caller's earlier work on the main stream (previous users of the arena) comes first
    BAS_CUDA(cudaEventRecord(ev->upload[0], (cudaStream_t)j->stream_main));
    BAS_CUDA(cudaStreamWaitEvent(up, ev->upload[0], 0));
    mark("upload begins", 0, 0, up);
    if (j->n_in > j->n)
        BAS_CUDA(cudaMemset2DAsync(x + j->n, (size_t)j->n_in * 4, 0, (size_t)(j->n_in - j->n) * 4, (size_t)j->n_src, up));
    // First input sample the render kernels READ: the tiled kernel works on whole 32-sample rows and on
    // ceil(K/32) tap blocks, so it multiplies samples back to p_begin/32*32 - 32*ceil(K/32) by taps that are
    // zero padding - those samples must be initialised (0 * NaN = NaN), hence they are uploaded too.
    long long lo = j->p_begin / 32 * 32 - 32LL * ((j->K + 31) / 32);
    lo = lo < 0 ? 0 : lo;
    for (int i = 0; i < n_phases; ++i) {
        long long hi = p_cuts[i + 1] < j->n ? p_cuts[i + 1] : j->n;      // outputs < cut need inputs < cut
        if (hi > lo) {
            if (int rc = bas_copy_2d(x + lo, j->n_in * 4, j->x_host + lo, j->x_host_stride * 4, (hi - lo) * 4, j->n_src, 1, up)) return rc;
            lo = hi;
        }
        BAS_CUDA(cudaEventRecord(ev->upload[i], up));
        mark("upload of phase %lld", i, 0, up);
    }
    return 0;
}

extern "C" int bas_pipeline_phase(const bas_pipeline_job* j, int phase, int n_phases, long long pt_begin, long long pt_end,
                                  long long p_from, long long p_to) {
    if (int rc = check_job(j)) return rc;
    BAS_CHECK_ARG(j->dirs_host && j->out_host && j->small_host, "null host buffer");
    BAS_CHECK_ARG(j->diffs_left_dev && j->diffs_right_dev && j->bank_pp_dev, "null bank pointer");
    const Layout l = layout_of(j->n_src, j->n_in, j->C, j->K, j->mix, j->p_count, j->x_dev != nullptr);
    BAS_CHECK_ARG(n_phases >= 1 && n_phases <= kMaxPhases && phase >= 0 && phase < n_phases, "phase");
    BAS_CHECK_ARG(0 <= pt_begin && pt_begin <= pt_end && pt_end <= l.n_pts, "direction range");
    BAS_CHECK_ARG(j->p_begin <= p_from && p_from <= p_to && p_to <= j->p_begin + j->p_count, "output range of the phase");
    // every input sample below p_to must have its chunk's two boundary filters planned
    const long long last_in = (p_to < j->n_in ? p_to : j->n_in) - 1;
    BAS_CHECK_ARG(p_to == p_from || last_in < 0 || last_in / j->C + 2 <= pt_end, "phase renders past its planned directions");
    Events* ev = events_for_current_device();
    BAS_CHECK_ARG(ev, "cannot create events on the current device");
    cudaStream_t mainst = (cudaStream_t)j->stream_main, down = (cudaStream_t)j->stream_down;
    char* arena = static_cast<char*>(j->arena_dev);
    bas_term* terms = reinterpret_cast<bas_term*>(arena + l.terms);
    float* filt = reinterpret_cast<float*>(arena + l.filt);
    int* small = reinterpret_cast<int*>(arena + l.small);
    const float* x = j->x_dev ? j->x_dev : reinterpret_cast<const float*>(arena + l.x);
    float* out = reinterpret_cast<float*>(arena + l.out);
    float* peaks = reinterpret_cast<float*>(small + 2);
    const long long m = pt_end - pt_begin;
    const bool fused = j->bank_pp2_dev != nullptr && bas_render_fused_fits(j->K, j->C, j->S, j->mix, j->variant);
    mark("phase %lld called", phase, 0, nullptr, true);

    if (phase == 0) {                   // status words and peaks: all zero
        BAS_CUDA(cudaMemsetAsync(small, 0, (size_t)(2 + j->n_src) * 4, mainst));
        mark("phase %lld memset", phase, 0, mainst);
    }
    // directions of this phase -> plan -> filter rows.  With one az kind for all directions (the usual
    // case) they travel inside the plan launches themselves (bas_plan_build_inline); a host -> device
    // copy would queue behind the signal upload on the same copy engine, and in-place reads of mapped
    // host memory wait behind the download traffic.
    if (m > 0) {
        double* dirs = reinterpret_cast<double*>(arena + l.dirs);
        uint8_t* kinds = reinterpret_cast<uint8_t*>(arena + l.kinds);
        // One source, one az kind (make_signal_move_2d): the directions ride in the plan launches themselves.
        // Several sources: ONE plan launch per phase that reads the directions in place from the caller's pinned
        // buffer (mapped host memory - a few bytes per point over PCIe, next to megabytes of signal upload);
        // if the buffer is not device-accessible they are copied first.
        const bool inline_dirs = j->n_src == 1 && !j->az_kind_host;
        const double* elev_src = nullptr;
        const uint8_t* kinds_src = nullptr;
        if (!inline_dirs) {
            void* mapped = nullptr;
            void* mapped_kinds = nullptr;
            const bool direct = cudaHostGetDevicePointer(&mapped, const_cast<double*>(j->dirs_host), 0) == cudaSuccess &&
                                (!j->az_kind_host || cudaHostGetDevicePointer(&mapped_kinds, const_cast<uint8_t*>(j->az_kind_host), 0) == cudaSuccess);
            if (direct) {
                elev_src = static_cast<const double*>(mapped);
                kinds_src = static_cast<const uint8_t*>(mapped_kinds);
            } else {
                cudaGetLastError();
                if (int rc = bas_copy_2d(dirs + pt_begin, l.n_pts * 8, j->dirs_host + pt_begin, l.n_pts * 8, m * 8, 2LL * j->n_src, 1, mainst)) return rc;
                if (j->az_kind_host)
                    if (int rc = bas_copy_2d(kinds + pt_begin, l.n_pts, j->az_kind_host + pt_begin, l.n_pts, m, j->n_src, 1, mainst)) return rc;
                elev_src = dirs;
                kinds_src = j->az_kind_host ? kinds : nullptr;
            }
            if (int rc = bas_plan_build_runs(j->diffs_left_dev, j->diffs_right_dev, j->U, j->K * j->U, elev_src + pt_begin, elev_src + l.n_dirs + pt_begin,
                                             kinds_src ? kinds_src + pt_begin : nullptr, j->az_kind_all, j->n_src, m, l.n_pts,
                                             terms + pt_begin * 2 * BAS_MAX_TERMS, nullptr, small, pt_begin, 0, mainst)) return rc;
            mark("phase %lld plan (one launch)", phase, 0, mainst);
        }
        for (int s = 0; s < j->n_src; ++s) {
            const long long first = (long long)s * l.n_pts + pt_begin;
            bas_term* t = terms + first * 2 * BAS_MAX_TERMS;
            if (inline_dirs) {
                if (int rc = bas_plan_build_inline(j->diffs_left_dev, j->diffs_right_dev, j->U, j->K * j->U, j->dirs_host + first,
                                                   j->dirs_host + l.n_dirs + first, j->az_kind_all, m, t, small, first, mainst)) return rc;
                mark("phase %lld plan of source 0", phase, 0, mainst);
            }
            if (!fused)
                if (int rc = bas_ir_synth(j->bank_pp_dev, j->U, j->K, t, m, BAS_IR_ROWS, filt + first * l.pitch * 2, j->K, mainst)) return rc;
        }
    }
    mark("phase %lld filter rows", phase, 0, mainst);
    if (!j->x_dev && j->p_count > 0) BAS_CUDA(cudaStreamWaitEvent(mainst, ev->upload[phase], 0));

    // segments: render on main, download on down
    if (p_to > p_from) {
        const long long count = p_to - p_from;
        const long long bytes_per_sample = 8LL * l.n_rows;
        long long n_seg = j->segment_bytes > 0 ? (count * bytes_per_sample + j->segment_bytes / 2) / j->segment_bytes : 1;
        if (n_seg < 1) n_seg = 1;
        if (n_seg > 64) n_seg = 64;
        const long long step = up_to((count + n_seg - 1) / n_seg, kSegAlign);
        for (long long pa = p_from; pa < p_to;) {
            long long pb = pa / kSegAlign * kSegAlign + step;     // cuts on the tile grid
            if (pb > p_to) pb = p_to;
            if (fused) {
                if (int rc = bas_render_fused(x, j->n_in, j->n_in, j->n_src, j->n_in, j->C, j->S, j->K, terms, j->bank_pp2_dev, j->U, nullptr,
                                              pa, pb - pa, out + (pa - j->p_begin), l.stride, j->mix, peaks, j->variant, j->workspace_dev,
                                              j->workspace_bytes, mainst)) return rc;
            } else if (int rc = bas_render(x, j->n_in, j->n_in, j->n_src, j->n_in, j->C, j->S, j->K, filt, nullptr, pa, pb - pa,
                                           out + (pa - j->p_begin), l.stride, j->mix, peaks, j->variant, j->workspace_dev,
                                           j->workspace_bytes, mainst)) return rc;
            BAS_CUDA(cudaEventRecord(ev->segment, mainst));
            BAS_CUDA(cudaStreamWaitEvent(down, ev->segment, 0));
            mark("render [%lld, %lld)", pa, pb, mainst);
            if (int rc = bas_copy_2d(j->out_host + (pa - j->p_begin), j->p_count * 4, out + (pa - j->p_begin), l.stride * 4,
                                     (pb - pa) * 4, 2LL * l.n_rows, 0, down)) return rc;
            mark("download [%lld, %lld)", pa, pb, down);
            pa = pb;
        }
    }
    if (phase == n_phases - 1) {
        BAS_CUDA(cudaEventRecord(ev->segment, mainst));
        BAS_CUDA(cudaStreamWaitEvent(down, ev->segment, 0));
        BAS_CUDA(cudaMemcpyAsync(j->small_host, small, (size_t)(2 + j->n_src) * 4, cudaMemcpyDeviceToHost, down));
        BAS_CUDA(cudaStreamSynchronize(down));
        mark("synchronised", 0, 0, nullptr, true);
    }
    return 0;
}

// Timeline of the pipeline calls made on this thread since tracing was switched on (debug aid).
extern "C" int bas_pipeline_trace(int enable, char* buf, size_t len) {
    if (buf && len) {
        std::string out;
        cudaEvent_t first = nullptr;
        for (const Mark& m : g_marks) if (m.ev) { first = m.ev; break; }
        const double h0 = g_marks.empty() ? 0.0 : g_marks.front().host_us;
        for (const Mark& m : g_marks) {
            char line[160];
            float ms = 0.f;
            if (m.ev && first) { cudaEventSynchronize(m.ev); cudaEventElapsedTime(&ms, first, m.ev); }
            if (m.ev) snprintf(line, sizeof(line), "%-34s enqueued %8.1f us   done on device %8.1f us\n", m.label.c_str(), m.host_us - h0, ms * 1e3);
            else snprintf(line, sizeof(line), "%-34s host     %8.1f us\n", m.label.c_str(), m.host_us - h0);
            out += line;
        }
        strncpy(buf, out.c_str(), len - 1);
        buf[len - 1] = 0;
    }
    for (Mark& m : g_marks) if (m.ev) cudaEventDestroy(m.ev);
    g_marks.clear();
    g_trace = enable != 0;
    return 0;
}
