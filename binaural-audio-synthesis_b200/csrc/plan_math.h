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

// ---- gather terms ---------------------------------------------------------------------------
// A two-tap fractional delay by d is the polynomial P(z) = (1-f) + f z in the shift operator z,
// anchored at shift floor(d) (apply_hrtf.py:149-165; when d is an integer f = 0 and the second tap,
// which the reference places at ceil(d) = floor(d), carries weight zero).  Chained delays multiply
// their polynomials and add their anchors, so each bank row contributes a short run of CONSECUTIVE
// shifts whose weights are the coefficients of a product of 2..4 such polynomials:
//     ring:  before row   (1-alpha)    P_restore                 2 shifts   (apply_hrtf.py:90-102)
//            after row     alpha       P_restore P_remove        3 shifts   (apply_hrtf.py:86-102)
//     2-D :  top rows      a           P_vrestore x ring         3 + 4      (apply_hrtf.py:268-277)
//            bottom rows  (1-a)        P_vrestore P_vremove x ring   4 + 5  (apply_hrtf.py:254-277)
// 16 terms per ear, in fixed slots, with no searching or merging.
struct BasPoly2 { double c0, c1; };                          // (1-f) + f z
BAS_HD BasPoly2 bas_poly(const BasDelay& d) { BasPoly2 p; p.c0 = 1.0 - d.frac; p.c1 = d.frac; return p; }

BAS_HD int32_t bas_pack_term(int row, long long shift, long long L) {
    long long s = shift % L;
    if (s < 0) s += L;
    return (int32_t)(((long long)row << 20) | s);
}
// A run of n consecutive shifts base, base + 1, ...: one modulo (64-bit division is a subroutine on the
// device), then increments that wrap at L.
#define BAS_PACK_RUN(out, first, n, row, base, L)                                        \
    do {                                                                                 \
        long long s_ = (base) % (L);                                                     \
        if (s_ < 0) s_ += (L);                                                           \
        _Pragma("unroll") for (int i_ = 0; i_ < (n); ++i_) {                             \
            (out)[(first) + i_].row_shift = (int32_t)(((long long)(row) << 20) | s_);    \
            if (++s_ == (L)) s_ = 0;                                                     \
        }                                                                                \
    } while (0)

// out[0..n] = scale * in[0..n-1] * (p.c0 + p.c1 z)
#define BAS_POLY_MUL(out, in, n, p)                                  \
    do {                                                             \
        (out)[0] = (in)[0] * (p).c0;                                 \
        _Pragma("unroll") for (int i_ = 1; i_ < (n); ++i_)           \
            (out)[i_] = (in)[i_] * (p).c0 + (in)[i_ - 1] * (p).c1;   \
        (out)[n] = (in)[(n) - 1] * (p).c1;                           \
    } while (0)

struct BasRingEar {          // scalar results of one ring interpolation for one ear
    BasDelay rem, res;       // remove (-d) and restore (alpha d) delays
    double delay_out;        // alpha d / U, what the reference returns (apply_hrtf.py:106)
};

BAS_HD BasRingEar bas_ring_ear(const double* diffs, int upsampling, const BasRing& rg, int* err) {
    BasRingEar r;
    const double d = (double)upsampling * diffs[rg.before * BAS_N_DIR + rg.after];   // :82-83
    r.rem = bas_split_delay(-d, err);                                                // :86-87
    const double d_back = rg.alpha * d;                                              // :94-95
    r.res = bas_split_delay(d_back, err);                                            // :98-102
    r.delay_out = d_back / (double)upsampling;                                       // :106
    return r;
}

// Ring mode: delay_compensated_interpolation_with_delaydiff (apply_hrtf.py:53-106) for one ear.
BAS_HD int bas_plan_ring_ear(const double* diffs, int upsampling, long long L, int before, int after,
                             double alpha, double one_minus_alpha, BasTerm* terms, double* delay_out,
                             long long* lo, long long* hi) {
    int err = 0;
    BasRing rg; rg.before = before; rg.after = after; rg.alpha = alpha; rg.one_minus_alpha = one_minus_alpha;
    const BasRingEar re = bas_ring_ear(diffs, upsampling, rg, &err);
    *delay_out = re.delay_out;
    lo[0] = re.rem.lo; hi[0] = re.rem.hi; lo[1] = re.res.lo; hi[1] = re.res.hi;
    const BasPoly2 p_res = bas_poly(re.res), p_rem = bas_poly(re.rem);
    double wb[2] = {one_minus_alpha * p_res.c0, one_minus_alpha * p_res.c1};
    double t1[2] = {alpha * p_res.c0, alpha * p_res.c1}, wa[3];
    BAS_POLY_MUL(wa, t1, 2, p_rem);
    for (int i = 0; i < BAS_MAX_TERMS; ++i) { terms[i].row_shift = 0; terms[i].weight = 0.0f; }
    for (int i = 0; i < 2; ++i) { terms[i].row_shift = bas_pack_term(before, re.res.lo + i, L); terms[i].weight = (float)wb[i]; }
    for (int i = 0; i < 3; ++i) { terms[2 + i].row_shift = bas_pack_term(after, re.res.lo + re.rem.lo + i, L); terms[2 + i].weight = (float)wa[i]; }
    return err;
}

struct BasPointGeom {        // the ear-independent part of a 2-D plan
    BasRing top, bot;
    double a;
    int err;
};

// apply_hrtf.py:201-215, :261-266 and sphere.py:78-121
BAS_HD BasPointGeom bas_point_geom(double elev, double azim, int az_kind) {
    BasPointGeom g;
    g.err = 0;
    // apply_hrtf.py:201-211: bracketing rings (comparisons in float64; NaN selects -45 / +90)
    int r_lo = -1, r_hi = -1;
    for (int r = 0; r < BAS_N_RING; ++r) {
        if (bas_ring_elev(r) <= elev) r_lo = r;
        if (r_hi < 0 && bas_ring_elev(r) >= elev) r_hi = r;
    }
    if (r_lo < 0) r_lo = 0;
    if (r_hi < 0) r_hi = BAS_N_RING - 1;
    g.top.before = g.top.after = g.bot.before = g.bot.after = 0; g.top.alpha = g.bot.alpha = 0.0;
    g.top.one_minus_alpha = g.bot.one_minus_alpha = 1.0;
    g.err |= bas_ring_lookup(r_hi, azim, az_kind, &g.top);             // :214
    g.err |= bas_ring_lookup(r_lo, azim, az_kind, &g.bot);             // :215
    g.a = 0.0;                                                         // :261-266
    if (bas_ring_elev(r_hi) > bas_ring_elev(r_lo)) {
        g.a = (elev - bas_ring_elev(r_lo)) / (bas_ring_elev(r_hi) - bas_ring_elev(r_lo));
        if (!(0.0 <= g.a && g.a <= 1.0)) { g.err |= BAS_ERR_VERT_ASSERT; g.a = 0.0; }
    }
    return g;
}

// One ear of interpolate_2d (apply_hrtf.py:219-277): 16 terms, and the six floor/ceil pairs.
BAS_HD int bas_plan_point_ear(const double* diffs, int upsampling, long long L, const BasPointGeom& g,
                              BasTerm* out, long long* lo, long long* hi) {
    int err = 0;
    const BasRingEar rt = bas_ring_ear(diffs, upsampling, g.top, &err);                      // :219
    const BasRingEar rb = bas_ring_ear(diffs, upsampling, g.bot, &err);                      // :220
    // :246-252   U * (-delay_top + diffs[top_before, bot_before] + delay_bot)
    const double dv = (double)upsampling * ((-rt.delay_out + diffs[g.top.before * BAS_N_DIR + g.bot.before]) + rb.delay_out);
    const BasDelay vr = bas_split_delay(-dv, &err);                                         // :254-255
    const double one_minus_a = 1.0 - g.a;
    const BasDelay vs = bas_split_delay(one_minus_a * dv, &err);                            // :272-277
    lo[0] = rt.rem.lo; hi[0] = rt.rem.hi; lo[1] = rt.res.lo; hi[1] = rt.res.hi;
    lo[2] = rb.rem.lo; hi[2] = rb.rem.hi; lo[3] = rb.res.lo; hi[3] = rb.res.hi;
    lo[4] = vr.lo; hi[4] = vr.hi; lo[5] = vs.lo; hi[5] = vs.hi;
    const BasPoly2 p4 = bas_poly(vs), p3 = bas_poly(vr);
    const BasPoly2 p2t = bas_poly(rt.res), p1t = bas_poly(rt.rem), p2b = bas_poly(rb.res), p1b = bas_poly(rb.rem);
    // top: a * P4 * P2t * {(1-alpha_t), alpha_t * P1t}
    double s1[1], s2[2], tb3[3], ta3[3], ta4[4];
    s1[0] = g.a;
    BAS_POLY_MUL(s2, s1, 1, p4);
    double sb[2] = {s2[0] * g.top.one_minus_alpha, s2[1] * g.top.one_minus_alpha};
    double sa[2] = {s2[0] * g.top.alpha, s2[1] * g.top.alpha};
    BAS_POLY_MUL(tb3, sb, 2, p2t);
    BAS_POLY_MUL(ta3, sa, 2, p2t);
    BAS_POLY_MUL(ta4, ta3, 3, p1t);
    // bottom: (1-a) * P4 * P3 * P2b * {(1-alpha_b), alpha_b * P1b}
    double u1[1], u2[2], u3[3], bb4[4], ba4[4], ba5[5];
    u1[0] = one_minus_a;
    BAS_POLY_MUL(u2, u1, 1, p4);
    BAS_POLY_MUL(u3, u2, 2, p3);
    double ub[3] = {u3[0] * g.bot.one_minus_alpha, u3[1] * g.bot.one_minus_alpha, u3[2] * g.bot.one_minus_alpha};
    double ua[3] = {u3[0] * g.bot.alpha, u3[1] * g.bot.alpha, u3[2] * g.bot.alpha};
    BAS_POLY_MUL(bb4, ub, 3, p2b);
    BAS_POLY_MUL(ba4, ua, 3, p2b);
    BAS_POLY_MUL(ba5, ba4, 4, p1b);
    const long long base_tb = vs.lo + rt.res.lo, base_ta = base_tb + rt.rem.lo;
    const long long base_bb = vs.lo + vr.lo + rb.res.lo, base_ba = base_bb + rb.rem.lo;
    BAS_PACK_RUN(out, 0, 3, g.top.before, base_tb, L);
    BAS_PACK_RUN(out, 3, 4, g.top.after, base_ta, L);
    BAS_PACK_RUN(out, 7, 4, g.bot.before, base_bb, L);
    BAS_PACK_RUN(out, 11, 5, g.bot.after, base_ba, L);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 3; ++i) out[i].weight = (float)tb3[i];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 4; ++i) { out[3 + i].weight = (float)ta4[i]; out[7 + i].weight = (float)bb4[i]; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 5; ++i) out[11 + i].weight = (float)ba5[i];
    return err;
}

// Full 2-D plan for one point: interpolate_2d (apply_hrtf.py:171-281), both ears (host path).
// diffs_l / diffs_r: 187x187 row-major doubles.  terms: [2][BAS_MAX_TERMS].  trace may be null.
BAS_HD int bas_plan_point(const double* diffs_l, const double* diffs_r, int upsampling, long long L,
                          double elev, double azim, int az_kind, BasTerm* terms, BasTrace* trace) {
    const BasPointGeom g = bas_point_geom(elev, azim, az_kind);
    int err = g.err;
    if (trace) {
        trace->rows[0] = g.top.before; trace->rows[1] = g.top.after;
        trace->rows[2] = g.bot.before; trace->rows[3] = g.bot.after;
        trace->alpha_top = g.top.alpha; trace->alpha_bot = g.bot.alpha; trace->a = g.a; trace->pad = 0;
    }
    for (int e = 0; e < 2; ++e) {
        long long lo[6], hi[6];
        err |= bas_plan_point_ear(e ? diffs_r : diffs_l, upsampling, L, g, terms + e * BAS_MAX_TERMS, lo, hi);
        if (trace) for (int i = 0; i < 6; ++i) { trace->lo[e][i] = lo[i]; trace->hi[e][i] = hi[i]; }
    }
    if (trace) trace->err = err;
    return err;
}
