// filter_emul.cpp -- host emulation of tiers 1 and 2 of K2's bin function (cdftools_b200/csrc/mocsig_kernel.cuh:
// sigma_bins_f32x8, sigma_bin_f64), operation for operation, on the constants the library itself derives
// (cdftools_b200/csrc/mocsig_filter.hpp).  Test infrastructure: lets the CPU suite check the fp32 tier's error bound and
// its bin decisions against the oracle without a GPU.  The only operation that cannot be reproduced bit for bit is
// rsqrt.approx (MUFU.RSQ); it is emulated by a correctly rounded 1/sqrt perturbed by +-2^-22 (the PTX bound) on request.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../cdftools_b200/csrc/mocsig_filter.hpp"
#include "../../include/cdf_eos_coeffs.h"

using namespace cdfgpu;
static SigFilter F;
static int NBINS;
static float SPS;

extern "C" {

// info8 as cdfmocsig_gpu_filter_info (first five entries) + dlref, qoff, sum|c32|
int filter_build(int teos10, float pref, float sigmin, float sigstp, int nbins, double *info8)
{
    double dlh, dlref;
    sig_reference_profile(CDF_EOS_R0, pref, &dlh, &dlref);
    const double m = sig_reference_margin(sigmin, sigstp, nbins);
    sig_filter_build(F, teos10 ? CDF_TEOS10_COEF : CDF_EOS80_COEF, teos10 ? CDF_TEOS10_RDELTAS : CDF_EOS80_RDELTAS,
                     teos10 ? CDF_TEOS10_R1_S0 : CDF_EOS80_R1_S0, pref, dlref, sigmin, sigstp, nbins, m);
    NBINS = nbins;
    SPS = 0.0f;
    info8[0] = F.tier1; info8[1] = F.tier2; info8[2] = F.err32; info8[3] = F.margin32; info8[4] = F.margin64;
    info8[5] = dlref; info8[6] = F.qoff.x;
    double sc = 0;
    for (int n = 0; n < kSigPolyTerms; ++n) sc += fabs((double)F.c32[n].x);
    info8[7] = sc;
    return 0;
}

// tier 1 on n cells: ib = trunc(q), ok = accepted, sig32 = the fp32 tier's sigma - sigmin in bin units (q).
// rsq_pert: relative perturbation applied to the reciprocal square root (0, +2^-22, -2^-22 ...).
void filter_tier1(const float *T, const float *S, long n, double rsq_pert, int *ib, unsigned char *ok, float *q32)
{
    for (long e = 0; e < n; ++e) {
        const float v = fmaf(T[e], F.ta.x, F.tb.x);
        const float x = fmaf(S[e], F.sr.x, F.sdr.x);
        const float nx = fmaf(S[e], F.nsr.x, F.nsdr.x);
        const float y = (float)((1.0 / sqrt((double)x)) * (1.0 + rsq_pert));
        float sx = x * y;
        const float er = fmaf(sx, sx, nx);
        sx = fmaf(er, y * -0.5f, sx);
        const float u = (sx + F.ns0.x) * F.ihs.x;
        float acc = F.c32[0].x;
        int idx = 1;
        for (int j = 5; j >= 0; --j) {
            float q = F.c32[idx++].x;
            for (int i = 5 - j; i >= 0; --i) q = fmaf(q, u, F.c32[idx++].x);
            acc = fmaf(acc, v, q);
        }
        const float qq = fmaf(acc, F.qscale.x, F.qoff.x);
        const float r = qq + 12582912.0f;
        const float d = qq - (r + -12582912.0f);
        int32_t rb, db;
        memcpy(&rb, &r, 4);
        memcpy(&db, &d, 4);
        ib[e] = (rb - 0x4B400000) + (db >> 31);
        ok[e] = (fabsf(d) > F.margin32) && (fabsf(qq - F.qc) < F.qh) && (fmaxf(fabsf(u), fabsf(v)) <= 1.0f);
        q32[e] = qq;
    }
}

void filter_tier2(const float *T, const float *S, long n, int *ib, unsigned char *ok, double *q64)
{
    for (long e = 0; e < n; ++e) {
        const double t = (double)T[e] * (1.0 / 40.0);
        const double s = sqrt(fabs((double)S[e] + F.rdeltaS) * F.r1_S0);
        double acc = F.c64[0];
        int idx = 1;
        for (int j = 5; j >= 0; --j) {
            double q = F.c64[idx++];
            for (int i = 5 - j; i >= 0; --i) q = fma(q, s, F.c64[idx++]);
            acc = fma(acc, t, q);
        }
        const double qa = fma(acc, F.inv_sigstp, F.qoffset);
        const int i = (int)qa;
        const double fr = qa - (double)i;
        ib[e] = i;
        ok[e] = (unsigned)(i - 1) < (unsigned)(NBINS - 1) && fabs(fr - 0.5) < F.half_m_margin && S[e] != 0.0f && S[e] != SPS;
        q64[e] = qa;
    }
}
}
