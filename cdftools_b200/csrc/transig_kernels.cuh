// transig_kernels.cuh -- K6: cdftransig_xy3d, volume transport of every grid cell binned in density space.
// Replaces the frame body of src/cdftransig_xy3d.f90:396-461 (one time frame, all levels).
//
//   dens   = sigmai(zt, zs, pref)                          REAL(8), the exact chain of eos.f90:842-882 (no FMA)
//   zdensu = fl32(0.5*(dens(i) + dens(i+1))) * zmasku      (last column: zdensu(2) when E-W periodic, else 0)
//   zdensv = fl32(0.5*(dens(j) + dens(j+1))) * zmaskv      (last row: 0 -- never assigned in the reference)
//   ijb    = clamp(INT((zdens - ds1min)/ds1scalmin) + 1, 1, nsigmax) ; ibin = itab(ijb)
//   dusigsig(i,j,ibinu) += dble(fl32(e2u * fl32(zu*e3u)))  ; dvsigsig(i,j,ibinv) += dble(fl32(e1v * fl32(zv*e3v)))
//
// Two passes per frame: (a) density of every cell into an fp64 scratch record (each density is needed by three cells;
// at ~120 fp64 operations it is cheaper to store 8 bytes than to evaluate it three times), (b) one thread per cell reads
// its three densities (neighbours come from L1 / L2), bins, and adds into the (nbins, ny, nx) accumulators, which stay
// resident across the frames of a run.  Every (i,j) column belongs to one thread (levels in sequence): plain
// read-modify-write, no atomics, the reference's order of additions within a frame.
// HBM-bound: about 80 bytes per level-cell (16 dens write + read, 16 velocities + metrics, 2 masks, 2 scattered RMWs).
#pragma once
#include "common.cuh"
#include "eos_device.cuh"   // exact EOS chain (eos_sigma_exact)

namespace cdfgpu {

template <bool SIGMA0>
__global__ void transig_dens_kernel(const float *__restrict__ zt, const float *__restrict__ zs, double dlh, double dlref, size_t n,
                                    double *__restrict__ dens)
{
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x)
        dens[c] = eos_sigma_exact<SIGMA0>(zt[c], zs[c], dlh, dlref);
}

struct TransigParams {
    const double *__restrict__ dens;                      // (nz-1, ny, nx)
    const float *__restrict__ zu, *__restrict__ zv;       // (nz-1, ny, nx)
    const float *__restrict__ e3u, *__restrict__ e3v;     // (nz-1, ny, nx)
    const float *__restrict__ e2u, *__restrict__ e1v;     // (ny, nx)
    uint8_t *masku, *maskv;                               // (nz-1, ny, nx) 0 / 1
    const int *__restrict__ itab;                         // (nsigmax) 1-based bins, 0 = none
    double *dusig, *dvsig;                                // (nbins, ny, nx)
    double ds1min, ds1scalmin;
    int nx, ny, nzm1, nsigmax, lperio, set_masks;
};

__device__ __forceinline__ int transig_step(float zdens, double ds1min, double ds1scalmin, int nsigmax)
{
    const double q = __ddiv_rn(__dadd_rn((double)zdens, -ds1min), ds1scalmin);
    // INT(): truncation toward zero; out of range and NaN give INT_MIN (x86 CVTTSD2SI)
    int ijb = (q > -2147483649.0 && q < 2147483648.0) ? __double2int_rz(q) : (int)0x80000000;
    ijb = (ijb == (int)0x80000000) ? ijb : ijb + 1;      // INT_MIN + 1 stays far below 1 either way; avoid the wrap
    return min(max(ijb, 1), nsigmax);
}

__global__ void transig_scatter_kernel(const TransigParams p)
{
    const size_t nxy = (size_t)p.nx * p.ny;
    // one thread per (i,j) column, levels in sequence: the column's accumulators are touched by this thread only
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nxy; c += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(c / p.nx), i = (int)(c - (size_t)j * p.nx);
        const float e2 = p.e2u[c], e1 = p.e1v[c];
        for (int k = 0; k < p.nzm1; ++k) {
            const size_t c3 = (size_t)k * nxy + c;
            const float u = p.zu[c3], v = p.zv[c3];
            float mu, mv;
            if (p.set_masks) {
                mu = (u == 0.0f) ? 0.0f : 1.0f;
                mv = (v == 0.0f) ? 0.0f : 1.0f;
                p.masku[c3] = (uint8_t)mu;
                p.maskv[c3] = (uint8_t)mv;
            } else {
                mu = (float)p.masku[c3];
                mv = (float)p.maskv[c3];
            }
            const double d0 = p.dens[c3];
            float zdu = 0.0f, zdv = 0.0f;
            if (i < p.nx - 1) zdu = __double2float_rn(__dmul_rn(0.5, __dadd_rn(d0, p.dens[c3 + 1])));
            else if (p.lperio && p.nx > 2) {
                const size_t r = c3 - (size_t)i;   // first cell of the row
                zdu = __double2float_rn(__dmul_rn(0.5, __dadd_rn(p.dens[r + 1], p.dens[r + 2])));
            }
            if (j < p.ny - 1) zdv = __double2float_rn(__dmul_rn(0.5, __dadd_rn(d0, p.dens[c3 + p.nx])));
            zdu = __fmul_rn(zdu, mu);
            zdv = __fmul_rn(zdv, mv);
            const int bu = p.itab[transig_step(zdu, p.ds1min, p.ds1scalmin, p.nsigmax) - 1];
            const int bv = p.itab[transig_step(zdv, p.ds1min, p.ds1scalmin, p.nsigmax) - 1];
            const float pu = __fmul_rn(e2, __fmul_rn(u, p.e3u[c3]));
            const float pv = __fmul_rn(e1, __fmul_rn(v, p.e3v[c3]));
            if (bu >= 1) {
                double *a = p.dusig + (size_t)(bu - 1) * nxy + c;
                *a = __dadd_rn(*a, (double)pu);
            }
            if (bv >= 1) {
                double *a = p.dvsig + (size_t)(bv - 1) * nxy + c;
                *a = __dadd_rn(*a, (double)pv);
            }
        }
    }
}

}  // namespace cdfgpu
