// Per-trajectory-point plan arithmetic, shared by the fp64 device kernel (plan_kernel.cu) and
// the host scalar helpers of the C-ABI (cabi_host.cpp).
//
// For one direction (elev, azim) this computes everything that is *scalar* in the
// reference's interpolate_2d: the four grid rows and ring weights (sphere.py:78-121), the
// vertical weight (apply_hrtf.py:199-211, :261-266), every fractional delay and its
// floor/ceil split (apply_hrtf.py:82-83, :94-95, :149-151, :246-252, :272-273), and flattens
// the nested 2-tap delays into at most 16 distinct (row, shift, weight) gather terms per ear:
//
//     out_e[m] = sum_t w_t * bank_e[row_t][(m*U - shift_t) mod L]          
//
// All integers are produced by the reference's own operation order in IEEE double (or single
// where NumPy-2 weak-scalar promotion makes the reference compute in float32), so they are
// bit-identical.  This translation unit must be compiled WITHOUT floating-point contraction
// (nvcc -fmad=false, gcc -ffp-contract=off).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define BAS_HD __host__ __device__ __forceinline__
#else
#define BAS_HD inline
#endif

// Scalar type of the azimuth the trajectory function returned (SURVEY.md section 5 "dtype
// hazard").  A Python float/int is a weak scalar: sphere.py:103-105,119 then compare and
// divide in float32.  An np.float64 keeps those in float64.  An np.float32 additionally
// makes the modulo of sphere.py:86 a float32 operation.
// The BAS_AZ_* kinds and BAS_ERR_* bits (sphere.py:87 azimuth assert, apply_hrtf.py:266 vertical
// assert, apply_hrtf.py:149-150 int(floor(nan))) are the public constants of include/bas_b200.h.
#include "../../include/bas_b200.h"

#define BAS_N_DIR 187
#define BAS_N_RING 10

struct BasTerm {
    int32_t row_shift;   // (row << 20) | shift, shift already reduced to [0, L)
    float weight;
};

// Everything the parity contract calls an index, for tests (one record per point).
struct BasTrace {
    int32_t rows[4];          // top_before, top_after, bot_before, bot_after  (apply_hrtf.py:214-215)
    int32_t err;
    int32_t pad;
    double alpha_top, alpha_bot, a;      // ring weights (sphere.py:119) and vertical weight (:262)
    // per ear: floor/ceil of the six delays in the order
    //   top remove (-d), top restore (alpha d), bot remove, bot restore   (apply_hrtf.py:86-87, :98-99)
    //   vertical remove (-dv), vertical restore ((1-a) dv)                (apply_hrtf.py:254-255, :276-277)
    int64_t lo[2][6];
    int64_t hi[2][6];
};

struct BasRing { int before, after; double alpha, one_minus_alpha; };

BAS_HD double bas_ring_elev(int r) {
    // np.deg2rad([-45,-30,-15,0,15,30,45,60,75,90])  (apply_hrtf.py:199); the two literals at
    // :204 and :209 are equal to the first and last entry.
    switch (r) {
        case 0: return -0x1.921fb54442d18p-1;
        case 1: return -0x1.0c152382d7365p-1;
        case 2: return -0x1.0c152382d7365p-2;
        case 3: return 0.0;
        case 4: return 0x1.0c152382d7365p-2;
        case 5: return 0x1.0c152382d7365p-1;
        case 6: return 0x1.921fb54442d18p-1;
        case 7: return 0x1.0c152382d7365p+0;
        case 8: return 0x1.4f1a6c638d03fp+0;
        default: return 0x1.921fb54442d18p+0;
    }
}
BAS_HD int bas_ring_start(int r) { return r < 7 ? 24 * r : (r == 7 ? 168 : (r == 8 ? 180 : 186)); }
BAS_HD int bas_ring_count(int r) { return r < 7 ? 24 : (r == 7 ? 12 : (r == 8 ? 6 : 1)); }

// float32 azimuth of point k on ring r: float32(deg) * float32(2*pi/360)  (sphere.py:315-318)
BAS_HD float bas_ring_azim(int r, int k) {
    const float deg = (float)(k * (360 / bas_ring_count(r)));
    return deg * 0x1.1df46ap-6f;
}

// Python / numpy float modulo for a positive divisor: fmod, then shift a negative remainder
// (CPython float_rem, numpy npy_divmod).
BAS_HD double bas_pymod(double x, double y) {
    double m = fmod(x, y);
    if (m != 0.0) { if (m < 0.0) m += y; } else { m = 0.0; }
    return m;
}
BAS_HD float bas_pymodf(float x, float y) {
    float m = fmodf(x, y);
    if (m != 0.0f) { if (m < 0.0f) m += y; } else { m = 0.0f; }
    return m;
}

// sphere.py:78-121 on ring r.  Returns nonzero on the reference's assertion failure.
BAS_HD int bas_ring_lookup(int r, double azim, int az_kind, BasRing* out) {
    const double two_pi = 0x1.921fb54442d18p+2;
    double az64 = 0.0;
    float az32 = 0.0f;
    if (az_kind == BAS_AZ_F32) {
        az32 = bas_pymodf((float)azim, (float)two_pi);              // float32 % weak python float
        if (!(az32 >= 0.0f)) return BAS_ERR_AZIM_ASSERT;
    } else {
        az64 = bas_pymod(azim, two_pi);                            // sphere.py:86
        if (!(az64 >= 0.0)) return BAS_ERR_AZIM_ASSERT;            // sphere.py:87
        az32 = (float)az64;                                        // weak scalar -> table dtype
    }
    if (r == BAS_N_RING - 1) {                                     // sphere.py:92-93
        out->before = 186; out->after = 186; out->alpha = 0.0; out->one_minus_alpha = 1.0;
        return 0;
    }
    const int start = bas_ring_start(r), count = bas_ring_count(r);
    int kb = 0, ka = -1;
    for (int k = 0; k < count; ++k) {
        const float t = bas_ring_azim(r, k);
        const bool le = (az_kind == BAS_AZ_F64) ? ((double)t <= az64) : (t <= az32);
        if (le) kb = k;                                            // max index with azim_row <= azim  (:103)
        else if (ka < 0) ka = k;                                   // min index with azim_row >  azim  (:104)
    }
    if (ka < 0) ka = 0;                                            // wrap to the ring's first row (:105-109)
    const float az_before = bas_ring_azim(r, kb);
    const float az_after = bas_ring_azim(r, ka);
    // denominator is float32 - float32, or weak 2*pi - float32: float32 either way (:115-119)
    const float den = (az_after < az_before ? (float)two_pi : az_after) - az_before;
    out->before = start + kb;
    out->after = start + ka;
    if (az_kind == BAS_AZ_F64) {
        out->alpha = (az64 - (double)az_before) / (double)den;
        out->one_minus_alpha = 1.0 - out->alpha;                   // apply_hrtf.py:90 in float64
    } else {
        const float a32 = (az32 - az_before) / den;
        out->alpha = (double)a32;
        out->one_minus_alpha = (double)(1.0f - a32);               // apply_hrtf.py:90: int - np.float32 stays float32
    }
    return 0;
}

struct BasDelay { long long lo, hi; double frac; };

// apply_hrtf.py:149-151
BAS_HD BasDelay bas_split_delay(double d, int* err) {
    BasDelay s;
    if (!(fabs(d) < 4.0e18)) { *err |= BAS_ERR_NONFINITE; d = 0.0; }
    const double fl = floor(d);
    s.lo = (long long)fl;
    s.hi = (long long)ceil(d);
    s.frac = d - fl;
    return s;
}

struct BasRawTerm { int row; long long shift; double w; };

// The six gather terms of one ring interpolation (apply_hrtf.py:82-102), and the delay it
// reports back in base-rate samples (:106).
BAS_HD double bas_ring_terms(const double* diffs, int upsampling, const BasRing& rg, BasRawTerm* t,
                             long long* lo, long long* hi, int* err) {
    const double d = (double)upsampling * diffs[rg.before * BAS_N_DIR + rg.after];   // :82-83
    const BasDelay rem = bas_split_delay(-d, err);                                   // :86-87
    const double d_back = rg.alpha * d;                                              // :94-95
    const BasDelay res = bas_split_delay(d_back, err);                               // :98-102
    lo[0] = rem.lo; hi[0] = rem.hi; lo[1] = res.lo; hi[1] = res.hi;
    const double w2[2] = {1.0 - res.frac, res.frac};
    const long long s2[2] = {res.lo, res.hi};
    const double w1[2] = {1.0 - rem.frac, rem.frac};
    const long long s1[2] = {rem.lo, rem.hi};
    int n = 0;
    for (int i = 0; i < 2; ++i) {
        t[n].row = rg.before; t[n].shift = s2[i]; t[n].w = w2[i] * rg.one_minus_alpha; ++n;
        for (int j = 0; j < 2; ++j) {
            t[n].row = rg.after; t[n].shift = s2[i] + s1[j]; t[n].w = w2[i] * (rg.alpha * w1[j]); ++n;
        }
    }
    return d_back / (double)upsampling;                                              // :106
}

BAS_HD void bas_merge_term(BasTerm* out, int* n_out, long long* keys, double* wsum, int row, long long shift,
                           double w, long long L) {
    long long s = shift % L;
    if (s < 0) s += L;
    const long long key = ((long long)row << 32) | s;
    for (int i = 0; i < *n_out; ++i)
        if (keys[i] == key) { wsum[i] += w; return; }
    if (*n_out < BAS_MAX_TERMS) {
        keys[*n_out] = key; wsum[*n_out] = w; ++*n_out;
    }
    (void)out;
}

// Ring mode: delay_compensated_interpolation_with_delaydiff (apply_hrtf.py:53-106) for one ear.
// Writes up to BAS_MAX_TERMS merged terms and the reported delay.
BAS_HD int bas_plan_ring_ear(const double* diffs, int upsampling, long long L, int before, int after,
                             double alpha, double one_minus_alpha, BasTerm* terms, double* delay_out,
                             long long* lo, long long* hi) {
    int err = 0;
    BasRing rg; rg.before = before; rg.after = after; rg.alpha = alpha; rg.one_minus_alpha = one_minus_alpha;
    BasRawTerm raw[6];
    *delay_out = bas_ring_terms(diffs, upsampling, rg, raw, lo, hi, &err);
    long long keys[BAS_MAX_TERMS]; double wsum[BAS_MAX_TERMS]; int n = 0;
    for (int i = 0; i < 6; ++i) bas_merge_term(terms, &n, keys, wsum, raw[i].row, raw[i].shift, raw[i].w, L);
    for (int i = 0; i < BAS_MAX_TERMS; ++i) {
        if (i < n) { terms[i].row_shift = (int32_t)(((keys[i] >> 32) << 20) | (keys[i] & 0xFFFFF)); terms[i].weight = (float)wsum[i]; }
        else { terms[i].row_shift = 0; terms[i].weight = 0.0f; }
    }
    return err;
}

// Full 2-D plan for one point: interpolate_2d (apply_hrtf.py:171-281).
// diffs_l / diffs_r: 187x187 row-major doubles.  terms: [2][BAS_MAX_TERMS].  trace may be null.
BAS_HD int bas_plan_point(const double* diffs_l, const double* diffs_r, int upsampling, long long L,
                          double elev, double azim, int az_kind, BasTerm* terms, BasTrace* trace) {
    int err = 0;
    // apply_hrtf.py:201-211: bracketing rings (comparisons in float64; NaN selects -45 / +90)
    int r_lo = -1, r_hi = -1;
    for (int r = 0; r < BAS_N_RING; ++r) {
        if (bas_ring_elev(r) <= elev) r_lo = r;
        if (r_hi < 0 && bas_ring_elev(r) >= elev) r_hi = r;
    }
    if (r_lo < 0) r_lo = 0;
    if (r_hi < 0) r_hi = BAS_N_RING - 1;
    BasRing top, bot;
    top.before = top.after = bot.before = bot.after = 0; top.alpha = bot.alpha = 0.0;
    top.one_minus_alpha = bot.one_minus_alpha = 1.0;
    err |= bas_ring_lookup(r_hi, azim, az_kind, &top);             // :214
    err |= bas_ring_lookup(r_lo, azim, az_kind, &bot);             // :215
    double a = 0.0;                                                // :261-266
    if (bas_ring_elev(r_hi) > bas_ring_elev(r_lo)) {
        a = (elev - bas_ring_elev(r_lo)) / (bas_ring_elev(r_hi) - bas_ring_elev(r_lo));
        if (!(0.0 <= a && a <= 1.0)) { err |= BAS_ERR_VERT_ASSERT; a = 0.0; }
    }
    if (trace) {
        trace->rows[0] = top.before; trace->rows[1] = top.after;
        trace->rows[2] = bot.before; trace->rows[3] = bot.after;
        trace->alpha_top = top.alpha; trace->alpha_bot = bot.alpha; trace->a = a; trace->pad = 0;
    }
    for (int e = 0; e < 2; ++e) {
        const double* diffs = e ? diffs_r : diffs_l;
        BasRawTerm tt[6], bt[6];
        long long lo[6], hi[6];
        const double d_top = bas_ring_terms(diffs, upsampling, top, tt, lo + 0, hi + 0, &err);   // :219
        const double d_bot = bas_ring_terms(diffs, upsampling, bot, bt, lo + 2, hi + 2, &err);   // :220
        // :246-252   U * (-delay_top + diffs[top_before, bot_before] + delay_bot)
        const double dv = (double)upsampling * ((-d_top + diffs[top.before * BAS_N_DIR + bot.before]) + d_bot);
        const BasDelay vr = bas_split_delay(-dv, &err);                         // :254-255
        const double one_minus_a = 1.0 - a;
        const BasDelay vs = bas_split_delay(one_minus_a * dv, &err);            // :272-277
        lo[4] = vr.lo; hi[4] = vr.hi; lo[5] = vs.lo; hi[5] = vs.hi;
        if (trace) for (int i = 0; i < 6; ++i) { trace->lo[e][i] = lo[i]; trace->hi[e][i] = hi[i]; }
        long long keys[BAS_MAX_TERMS]; double wsum[BAS_MAX_TERMS]; int n = 0;
        const double w4[2] = {1.0 - vs.frac, vs.frac};
        const long long s4[2] = {vs.lo, vs.hi};
        const double w3[2] = {1.0 - vr.frac, vr.frac};
        const long long s3[2] = {vr.lo, vr.hi};
        for (int i = 0; i < 2; ++i) {
            for (int k = 0; k < 6; ++k)                                          // a * hrtf_top   (:268-269)
                bas_merge_term(0, &n, keys, wsum, tt[k].row, s4[i] + tt[k].shift, w4[i] * (a * tt[k].w), L);
            for (int j = 0; j < 2; ++j)
                for (int k = 0; k < 6; ++k)                                      // (1-a) * bottom_nodelay
                    bas_merge_term(0, &n, keys, wsum, bt[k].row, s4[i] + s3[j] + bt[k].shift,
                                   w4[i] * (one_minus_a * (w3[j] * bt[k].w)), L);
        }
        BasTerm* out = terms + e * BAS_MAX_TERMS;
        for (int i = 0; i < BAS_MAX_TERMS; ++i) {
            if (i < n) { out[i].row_shift = (int32_t)(((keys[i] >> 32) << 20) | (keys[i] & 0xFFFFF)); out[i].weight = (float)wsum[i]; }
            else { out[i].row_shift = 0; out[i].weight = 0.0f; }
        }
    }
    if (trace) trace->err = err;
    return err;
}
