// One render step as ONE call: the scalar part of interpolate_2d for every chunk boundary
// (bas_plan_build), the filter rows (bas_ir_synth - or none, when the render kernel synthesises them
// itself), the chunk / subchunk FIR of make_signal_move_2d (bas_render / bas_render_fused) and the peak
// division (bas_normalise), enqueued back to back on one stream by one host call.
//
// apply_hrtf.py:429-435 (one interpolate_2d per chunk boundary) feeding :438-453 (the subchunk loop) and
// :459-464 (cast, peak, divide).
//
// Compared with issuing the same launches one C-ABI call at a time from the host language:
//   - one host call and one memset (status words and peaks are laid out together and all-zero means
//     "no error, no peak") instead of seven stream operations;
//   - the kernels after the first are launched with programmatic stream serialization: their CTAs
//     become resident while the previous kernel drains and wait (griddepcontrol.wait) until it has
//     completed, so launch latency and per-CTA set-up overlap the tail of the kernel before.
#include "bas_internal.cuh"

extern "C" int bas_render_step(const bas_step_job* j, void* stream) {
    BAS_CHECK_ARG(j, "null job");
    BAS_CHECK_ARG(j->n_src >= 1 && j->C >= 1 && j->S >= 1 && j->C % j->S == 0 && j->K >= 1 && j->U >= 1, "geometry");
    BAS_CHECK_ARG(j->n_in >= j->C && j->n_in % j->C == 0, "n_in must be a positive multiple of C");
    BAS_CHECK_ARG(j->small_dev && j->terms_dev, "null scratch pointer");
    const long long n_pts = j->n_in / j->C + 1;
    const long long n_dirs = n_pts * j->n_src;
    const bool fused = (j->flags & BAS_STEP_FUSED) != 0;
    int* status = j->small_dev;
    float* peaks = reinterpret_cast<float*>(j->small_dev + 2);
    cudaStream_t st = (cudaStream_t)stream;

    if (j->flags & BAS_STEP_PLAN) {
        BAS_CHECK_ARG(j->elev_dev && j->azim_dev && j->diffs_left_dev && j->diffs_right_dev, "null plan input");
        BAS_CUDA(cudaMemsetAsync(j->small_dev, 0, (size_t)(2 + j->n_src) * 4, st));
        {
            BasPdlScope first(false);           // the first kernel of the step follows a memset: full serialization
            if (int rc = bas_plan_build_range(j->diffs_left_dev, j->diffs_right_dev, j->U, j->K * j->U, j->elev_dev, j->azim_dev,
                                              j->az_kind_dev, j->az_kind_all, n_dirs, j->terms_dev, nullptr, status, 0, 0, st)) return rc;
        }
        if (!fused) {
            BAS_CHECK_ARG(j->filt_dev && j->bank_pp_dev, "null filter-row scratch or bank");
            BasPdlScope chained(true);
            if (int rc = bas_ir_synth(j->bank_pp_dev, j->U, j->K, j->terms_dev, n_dirs, BAS_IR_ROWS, j->filt_dev, j->K, st)) return rc;
        }
    }
    if ((j->flags & BAS_STEP_RENDER) && j->p_count > 0) {
        BasPdlScope chained((j->flags & BAS_STEP_PLAN) != 0);      // a render-only call follows whatever the caller enqueued
        int rc;
        if (j->route && j->route->n > 1)
            rc = bas_render_routed(j->x_dev, j->x_stride, j->n_valid, j->n_src, j->n_in, j->C, j->S, j->K, fused ? nullptr : j->filt_dev,
                                   j->terms_dev, j->bank_pp2_dev, j->U, j->gains_dev, j->p_begin, j->p_count, j->out_dev, j->out_stride,
                                   peaks, j->variant, j->workspace_dev, j->workspace_bytes, j->route, st);
        else if (fused)
            rc = bas_render_fused(j->x_dev, j->x_stride, j->n_valid, j->n_src, j->n_in, j->C, j->S, j->K, j->terms_dev, j->bank_pp2_dev,
                                  j->U, j->gains_dev, j->p_begin, j->p_count, j->out_dev, j->out_stride, j->mix, peaks, j->variant,
                                  j->workspace_dev, j->workspace_bytes, st);
        else
            rc = bas_render(j->x_dev, j->x_stride, j->n_valid, j->n_src, j->n_in, j->C, j->S, j->K, j->filt_dev, j->gains_dev,
                            j->p_begin, j->p_count, j->out_dev, j->out_stride, j->mix, peaks, j->variant, j->workspace_dev,
                            j->workspace_bytes, st);
        if (rc) return rc;
        if ((j->flags & BAS_STEP_NORMALISE) && !j->mix) {
            // apply_hrtf.py:462-464 per source; the kernel returns at once unless the source's peak exceeds 1
            BasPdlScope after_render(true);
            for (int s = 0; s < j->n_src; ++s)
                if (int rc2 = bas_normalise(j->out_dev + (long long)s * 2 * j->out_stride, 2 * j->out_stride, peaks + s, st)) return rc2;
        }
    }
    return 0;
}
