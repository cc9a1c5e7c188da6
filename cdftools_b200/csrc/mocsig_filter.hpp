// mocsig_filter.hpp -- constants of tiers 1 and 2 of K2's bin function (mocsig_kernel.cuh) and the host code that derives
// them, including the RIGOROUS error bound that makes the fp32 tier safe.  Plain C++ (no CUDA types beyond float2): the
// same header serves the kernel (the struct travels in the kernel parameters) and the host.
//
// The reference evaluates (src/eos.f90:848-882)
//     sigma = ((dlr3*h + dlr2)*h + dlr1)*h + dlr0 + dlref - 1000 ,  dlr_k = sum_ij EOSijk s^i t^j ,
//     t = T/40, s = sqrt(|S + deltaS| * r1_S0), h = pref * 1e-4,
// in fp64 and bins it with an fp32 formula (src/cdfmocsig.f90:399-403).  For a fixed h this is ONE bivariate polynomial
//     P(s,t) = sum_ij c_ij s^i t^j ,  c_ij = sum_k EOSijk h^k          (28 terms, total degree 6)
// whatever the reference depth.  Tier 2 evaluates P in fp64 with FMAs.  Tier 1 evaluates it in fp32 after a change of
// variables to the domain box: u = (s - s0)/hs, v = (t - t0)/ht in [-1,1], P(s,t) = c'_00 + sum_(a,b)!=(0,0) c'_ab u^a v^b.
// The c'_ab are O(10) where the c_ij are O(10^3) with cancellation, which is what makes fp32 good to ~2e-5 kg/m3.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef __CUDACC__
struct float2 { float x, y; };
#endif

namespace cdfgpu {

constexpr int kSigPolyTerms = 28;

struct SigFilter {
    // ---- tier 1 (fp32, packed: every constant duplicated so that one 64-bit constant load feeds an FFMA2) ----
    float2 c32[kSigPolyTerms];   // c'_ab in Horner order (see sig_horner_order), c'_00 replaced by 0
    float2 ta, tb;               // v = fma(T, ta, tb)
    float2 sr, sdr, nsr, nsdr;   // x = fma(S, sr, sdr) = (S + deltaS) * r1_S0;  -x = fma(S, nsr, nsdr)
    float2 ns0, ihs;             // u = (sqrt(x) + ns0) * ihs,  ns0 = -s0
    float2 qscale, qoff;         // q = fma(P', qscale, qoff) = (sigma - sigmin) / sigstp
    float margin32;              // accept only |q - rint(q)| > margin32
    float qc, qh;                // ... and |q - qc| < qh  (1 < q < nbins)
    int tier1;                   // 0: tier 1 off
    // ---- tier 2 (fp64) ----
    double c64[kSigPolyTerms];   // c_ij in Horner order
    double rdeltaS, r1_S0;
    double inv_sigstp, qoffset;  // q = fma(P, inv_sigstp, qoffset), qoffset = (dlref - 1000 - sigmin) / sigstp
    double half_m_margin;        // accept only |frac(q) - 0.5| < 0.5 - margin64
    int tier2;                   // 0: tier 2 off
    // ---- diagnostics (tests) ----
    float Tmin, Tmax, Smin, Smax;   // domain box of tier 1
    double err32;                   // bound on |sigma_tier1 - sigma_reference|, kg/m3
    double margin64;
};

// Horner order of the 28 terms: for j = 6..0 (power of t / v), i = 6-j..0 (power of s / u):
//   acc = c[0]; for j = 5..0 { q = c[next]; for i = 5-j..0: q = q*s + c[next]; acc = acc*t + q }
inline void sig_horner_order(int (&pi)[kSigPolyTerms], int (&pj)[kSigPolyTerms])
{
    int n = 0;
    for (int j = 6; j >= 0; --j)
        for (int i = 6 - j; i >= 0; --i) { pi[n] = i; pj[n] = j; ++n; }
}

// index of EOSijk in the 52-entry tables of include/cdf_eos_coeffs.h (k-major, then j, then i); the h^k block holds the
// terms with i + j <= sig_eos_degree(k) = 6, 4, 2, 1 (src/eos.f90:853-877: dlr0 .. dlr3)
inline int sig_eos_degree(int k) { return k == 0 ? 6 : k == 1 ? 4 : k == 2 ? 2 : 1; }
inline int sig_eos_index(int i, int j, int k)
{
    int n = 0;
    for (int kk = 0; kk < k; ++kk) { const int d = sig_eos_degree(kk); n += (d + 1) * (d + 2) / 2; }
    const int d = sig_eos_degree(k);
    for (int jj = 0; jj < j; ++jj) n += d - jj + 1;
    return n + i;
}

// Reference profile at pref (src/eos.f90:843-845): dlh = pref * r1_Z0 and dlref, plain fp64 operations in the reference's
// order (volatile: no contraction, no reassociation on the host).
inline void sig_reference_profile(const double *R0, float pref, double *dlh, double *dlref)
{
    volatile double h = (double)pref * 1.e-4;
    volatile double a = R0[5] * h;
    a = a + R0[4]; a = a * h; a = a + R0[3]; a = a * h; a = a + R0[2]; a = a * h; a = a + R0[1]; a = a * h;
    a = a + R0[0]; a = a * h;
    *dlh = h;
    *dlref = a;
}

// Margin (bin units) that covers the fp32 roundings of the reference's own bin formula (src/cdfmocsig.f90:399-403):
//   z = fl32(sigma), a = fl32(z - sigmin), q = fl32(a / sigstp), bin = INT(q).  With S = largest |sigma| whose bin is not
//   decided by the clamp, |q - (sigma - sigmin)/sigstp| <= (2^-24 S + 2^-24 (S+|sigmin|) + dsig)/sigstp + 2^-24 (nbins+2),
//   where dsig = 1e-9 kg/m3 bounds |sigma_fp64_fma - sigma_reference| (both are roundings of the same polynomial; > 10^3
//   times the observed difference).  The margin is that bound times 2, plus the fp64 slop of forming q.  A non-positive or
//   non-finite step, or a margin above 1/8, gives -1 (no filter).
inline double sig_reference_margin(float sigmin, float sigstp, int nbins)
{
    const double stp = (double)sigstp, smin = (double)sigmin;
    if (!(stp > 0.0 && stp < 1e30 && fabs(smin) < 1e30)) return -1.0;
    const double S = fabs(smin) + (nbins + 2.0) * stp;
    const double u = 5.9604644775390625e-08;  // 2^-24
    const double bound = (u * S + u * (S + fabs(smin)) + 1e-9) / stp + u * (nbins + 2.0);
    const double m = 2.0 * bound + 1e-9;
    return m < 0.125 ? m : -1.0;
}

// qmargin_ref: the margin (q units) that covers the fp32 roundings of the reference's own bin formula plus 1e-9 kg/m3 of
// fp64 evaluation slop (api_mocsig.inc); < 0 switches both tiers off.
inline void sig_filter_build(SigFilter &f, const double *coef52, double rdeltaS, double r1_S0, float pref, double dlref, float sigmin,
                             float sigstp, int nbins, double qmargin_ref)
{
    memset(&f, 0, sizeof(f));
    f.rdeltaS = rdeltaS; f.r1_S0 = r1_S0;
    f.margin64 = qmargin_ref;
    if (!(qmargin_ref >= 0.0)) return;
    typedef long double ld;
    const ld h = (ld)((double)pref * 1.e-4);
    // ---- c_ij = sum_k EOSijk h^k ----
    ld c[7][7];
    for (int i = 0; i < 7; ++i)
        for (int j = 0; j < 7; ++j) {
            c[i][j] = 0.0L;
            ld hk = 1.0L;
            for (int k = 0; k <= 3 && i + j <= sig_eos_degree(k); ++k) { c[i][j] += (ld)coef52[sig_eos_index(i, j, k)] * hk; hk *= h; }
        }
    int pi[kSigPolyTerms], pj[kSigPolyTerms];
    sig_horner_order(pi, pj);
    for (int n = 0; n < kSigPolyTerms; ++n) f.c64[n] = (double)c[pi[n]][pj[n]];
    const double stp = (double)sigstp;
    f.inv_sigstp = 1.0 / stp;
    f.qoffset = (dlref - 1000.0 - (double)sigmin) * f.inv_sigstp;
    f.half_m_margin = 0.5 - qmargin_ref;
    f.tier2 = 1;

    // ---- tier 1: domain box, change of variables through the fp32 constants the kernel will really use ----
    f.Tmin = -4.0f; f.Tmax = 42.0f; f.Smin = 1.0f; f.Smax = 45.0f;
    const ld smin = sqrtl(((ld)f.Smin + rdeltaS) * r1_S0), smax = sqrtl(((ld)f.Smax + rdeltaS) * r1_S0);
    const ld tmin = (ld)f.Tmin / 40.0L, tmax = (ld)f.Tmax / 40.0L;
    auto dup = [](float v) { float2 r; r.x = v; r.y = v; return r; };
    f.ns0 = dup(-(float)((smin + smax) / 2.0L));
    f.ihs = dup((float)(2.0L / (smax - smin)));
    f.ta = dup((float)(1.0L / (40.0L * (tmax - tmin) / 2.0L)));
    f.tb = dup((float)(-((tmin + tmax) / 2.0L) / ((tmax - tmin) / 2.0L)));
    f.sr = dup((float)r1_S0);
    f.sdr = dup((float)(rdeltaS * r1_S0));
    f.nsr = dup(-f.sr.x);
    f.nsdr = dup(-f.sdr.x);
    // the affine maps the fp32 constants define EXACTLY: s = s0 + hs*u, t = t0 + ht*v
    const ld s0 = -(ld)f.ns0.x, hs = 1.0L / (ld)f.ihs.x;
    const ld ht = 1.0L / (40.0L * (ld)f.ta.x), t0 = -(ld)f.tb.x * ht;
    auto binom = [](int n, int k) { ld r = 1.0L; for (int q = 1; q <= k; ++q) r = r * (ld)(n - k + q) / (ld)q; return r; };
    ld cs[7][7];
    for (int a = 0; a < 7; ++a)
        for (int b = 0; b < 7; ++b) {
            cs[a][b] = 0.0L;
            if (a + b > 6) continue;
            for (int i = a; i < 7; ++i)
                for (int j = b; i + j < 7; ++j)
                    cs[a][b] += c[i][j] * binom(i, a) * powl(s0, i - a) * powl(hs, a) * binom(j, b) * powl(t0, j - b) * powl(ht, b);
        }
    ld sumc = 0.0L, Du = 0.0L, Dv = 0.0L;
    for (int n = 0; n < kSigPolyTerms; ++n) {
        const int a = pi[n], b = pj[n];
        const float cf = (a == 0 && b == 0) ? 0.0f : (float)cs[a][b];
        f.c32[n].x = cf; f.c32[n].y = cf;
        sumc += fabsl((ld)cf);
        Du += a * fabsl((ld)cf);
        Dv += b * fabsl((ld)cf);
    }
    // ---- error bound of the kernel's operation sequence (sigma_bins_f32x8), u = 2^-24 ----
    //  v^ = fma(T,ta,tb): ONE rounding (ta, tb are exact in the map above)                          |dv| <= u
    //  x^ = fma(S,sr,sdr): sr, sdr rounded (rel. u each) + one rounding                             rel. 2u (+u^2)
    //  sqrt: rsqrt.approx (rel. 2^-22) -> x*y (u) -> one Newton step, residual 2*(1.5*2^-22)^2 < 2^-40, + one rounding
    //        => s^ = s (1 + e), |e| <= u (from x) + u + 2^-40
    //  u^ = (s^ - s0) * ihs: one rounding each (s0, ihs exact in the map)                           |du| <= |e| smax / hs + 2u
    //  Horner with FMAs: every coefficient passes through at most 8 roundings; with the rounding of the coefficient itself
    //        |P^ - P(u^,v^)| <= 12 u sum|c'|  (gamma_9 + 1, rounded up), |u^|,|v^| <= 1 is CHECKED on the computed values
    //  P(u^,v^) - P(u,v) <= Du |du| + Dv |dv|   (mean value theorem on the box; second order terms in the 1e-4 slack)
    const ld u = ldexpl(1.0L, -24);
    const ld es = 2.0L * u + ldexpl(1.0L, -40);
    const ld du = es * smax / hs + 2.0L * u, dv = u;
    const ld EP = (12.0L * u * sumc + Du * du + Dv * dv) * (1.0L + 1e-4L);
    f.err32 = (double)EP;
    //  q^ = fma(P^, qscale, qoff): qscale, qoff rounded (rel. u each) + one rounding, |q^| < nbins + 1 when accepted
    const ld qoffl = (cs[0][0] + (ld)dlref - 1000.0L - (ld)sigmin) / (ld)stp;
    f.qscale = dup((float)(1.0L / (ld)stp));
    f.qoff = dup((float)qoffl);
    const ld Eq = EP / (ld)stp + u * ((ld)nbins + 2.0L) + u * sumc / (ld)stp + u * fabsl(qoffl);
    const ld m32 = ((ld)qmargin_ref + Eq) * 1.05L;
    f.margin32 = (float)m32;
    if ((ld)f.margin32 < m32) f.margin32 = nextafterf(f.margin32, 1.0f);
    f.qc = 0.5f * ((float)nbins + 1.0f);
    f.qh = 0.5f * ((float)nbins - 1.0f);
    // usable only while the magic-number rounding is valid (q < 2^22), the margin leaves room and the constants are finite
    f.tier1 = (m32 < 0.25L && nbins >= 2 && nbins < (1 << 21) && fabsl(qoffl) < 1e6L && isfinite(f.qscale.x) && isfinite(f.qoff.x)) ? 1 : 0;
}

}  // namespace cdfgpu
