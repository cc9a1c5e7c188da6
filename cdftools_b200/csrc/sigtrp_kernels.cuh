// sigtrp_kernels.cuh -- K7: cdfsigtrp, transport across a section in density classes.
// Replaces the compute part of one section, src/cdfsigtrp.f90:559-627.  Section arrays are (npts, npk) in the reference
// = [npk][npts] here (the along-section index is the fast one, so consecutive threads read consecutive columns).
//
//   dsig(ji,jk)  = sigmai(zt,zs,refdep)*zmask | sigmantr*zmask | -zt*zmask            (:560-567)  kernel (a), one thread per cell
//   dsig(ji,0)   = dsig(ji,1) - 1.e-4 ; land cells: dsig(jk) = dsig(jk-1) + 1.e-5     (:569-578)  kernel (b), one thread per column
//   dhiso(ji,l)  = depth at which the column first reaches dsigma_lev(l)              (:581-601)  kernel (c), one thread per
//   dwtrp(ji,l)  = transport from the surface down to dhiso(ji,l)                     (:604-618)              (column, level l)
//   dwtrpbin     = dwtrp(l+1) - dwtrp(l) ; dtrpbin(l) = SUM_ji dwtrpbin               (:621-627)  kernel (d), one thread per bin,
//                                                                                                 ji ascending as SUM does
// All fp64 arithmetic is spelled with __dmul_rn / __dadd_rn / __ddiv_rn: no contraction, the reference's association.
// The work is O(npts * nbins * nk) per section (1e6 .. 1e8 operations): latency-bound, not HBM-bound -- the kernels exist
// so that the tool is a drop-in on the same library, they are not a roofline item.
#pragma once
#include "common.cuh"
#include "eos_device.cuh"

namespace cdfgpu {

struct SigtrpParams {
    const float *__restrict__ eu;      // (npts)
    const float *__restrict__ de3;     // (npk, npts)  REAL(8) in the reference, holding REAL(4) file values
    const double *__restrict__ ddepu;  // (npk+1, npts), row 0 = the reference's ddepu(:,0) = 0
    const float *__restrict__ gdepw;   // (npk)
    const float *__restrict__ ddepw;   // (npk, npts) or NULL (-brk: the column's own w depths)
    const float *__restrict__ zu, *__restrict__ zt, *__restrict__ zs, *__restrict__ zmask;   // (npk, npts)
    const double *__restrict__ lev;    // (nbins+1) dsigma_lev
    double *dsig;                      // (nk+1, npts)
    double *dhiso, *dwtrp;             // (nbins+1, npts)
    double *dwtrpbin;                  // (nbins, npts)
    double *dtrpbin;                   // (nbins)
    double dlh, dlref;
    int npts, npk, nk, nbins;
};

// MODE 0: sigma0 (refdep == 0), 1: sigmai(refdep), 2: neutral density, 3: -temp
template <int MODE>
__global__ void sigtrp_density_kernel(const SigtrpParams p)
{
    const size_t n = (size_t)p.nk * p.npts;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x) {
        double d;
        if (MODE == 3) {
            d = (double)__fmul_rn(-p.zt[c], p.zmask[c]);
        } else {
            // sigma0: dlh = 0 drops the pressure terms exactly -- for finite T, S.  The reference still multiplies them by 0,
            // so an Inf in T or S gives NaN there (Inf * 0), not +-Inf: those cells take the full expression.
            const float t = p.zt[c], sal = p.zs[c];
            const bool fin = (__float_as_uint(t) & 0x7f800000u) != 0x7f800000u && (__float_as_uint(sal) & 0x7f800000u) != 0x7f800000u;
            const double s = MODE == 2             ? eos_sigma_neutral(t, sal)
                             : (MODE == 0 && fin) ? eos_sigma_exact<true>(t, sal, p.dlh, p.dlref)
                                                  : eos_sigma_exact<false>(t, sal, p.dlh, p.dlref);
            d = __dmul_rn(s, (double)p.zmask[c]);
        }
        p.dsig[(size_t)p.npts + c] = d;
    }
}

__global__ void sigtrp_fill_kernel(const SigtrpParams p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.npts) return;
    double prev = __dadd_rn(p.dsig[(size_t)p.npts + i], -(double)1.e-4f);
    p.dsig[i] = prev;
    for (int k = 1; k <= p.nk; ++k) {
        const size_t c = (size_t)k * p.npts + i;
        double d = p.dsig[c];
        if (p.zmask[c - p.npts] == 0.0f) {
            d = __dadd_rn(prev, (double)1.e-5f);
            p.dsig[c] = d;
        }
        prev = d;
    }
}

__global__ void sigtrp_iso_kernel(const SigtrpParams p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x, iso = blockIdx.y;
    if (i >= p.npts) return;
    const double dsigma = p.lev[iso];
    double h = p.ddepu[(size_t)p.npk * p.npts + i];
    double d0 = p.dsig[i];
    for (int k = 1; k <= p.nk; ++k) {
        const double d1 = p.dsig[(size_t)k * p.npts + i];
        if (!(d1 < dsigma)) {
            const double dalfa = __ddiv_rn(__dadd_rn(dsigma, -d0), __dadd_rn(d1, -d0));
            if (fabs(dalfa) > 1.1 || dalfa < 0.0) {
                h = 0.0;
            } else {
                const double a = __dmul_rn(p.ddepu[(size_t)k * p.npts + i], dalfa);
                const double b = __dmul_rn(__dadd_rn(1.0, -dalfa), p.ddepu[(size_t)(k - 1) * p.npts + i]);
                h = __dadd_rn(a, b);
            }
            break;
        }
        d0 = d1;
    }
    p.dhiso[(size_t)iso * p.npts + i] = h;
    const double e = (double)p.eu[i];
    double w = 0.0;
    for (int k = 1; k <= p.nk - 1; ++k) {
        const float gw1 = p.ddepw ? p.ddepw[(size_t)k * p.npts + i] : p.gdepw[k];
        const double u = (double)p.zu[(size_t)(k - 1) * p.npts + i];
        if ((double)gw1 < h) {
            w = __dadd_rn(w, __dmul_rn(__dmul_rn(e, (double)p.de3[(size_t)(k - 1) * p.npts + i]), u));
        } else {
            const float gw0 = p.ddepw ? p.ddepw[(size_t)(k - 1) * p.npts + i] : p.gdepw[k - 1];
            w = __dadd_rn(w, __dmul_rn(__dmul_rn(e, __dadd_rn(h, -(double)gw0)), u));
            break;
        }
    }
    p.dwtrp[(size_t)iso * p.npts + i] = w;
}

// one warp per bin: coalesced differences, then lane 0 adds them in ji order (what SUM does) through shuffles
__global__ void sigtrp_bins_kernel(const SigtrpParams p)
{
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (b >= p.nbins) return;   // warp-uniform
    double s = 0.0;
    for (int i0 = 0; i0 < p.npts; i0 += 32) {
        const int i = i0 + lane;
        double d = 0.0;
        if (i < p.npts) {
            d = __dadd_rn(p.dwtrp[(size_t)(b + 1) * p.npts + i], -p.dwtrp[(size_t)b * p.npts + i]);
            p.dwtrpbin[(size_t)b * p.npts + i] = d;
        }
        const int n = min(32, p.npts - i0);
        for (int l = 0; l < n; ++l) s = __dadd_rn(s, __shfl_sync(kFull, d, l));
    }
    if (lane == 0) p.dtrpbin[b] = s;
}

}  // namespace cdfgpu
