// moc_decomp.cuh -- cdfmoc -decomp: barotropic / geostrophic-shear / ageostrophic split of the MOC.
// Replaces src/cdfmoc.f90:353,360-365,390-517.  Not a bandwidth showcase: six small kernels that
// reproduce the reference's REAL(4)/REAL(8) chains operation by operation; the zonal sums reuse K1 (noscan).
//   dvbt(i,j)   = sum_k dble(fl32(e3m*zv)) / hdep            hdep = fl32 running sum of e3m          (:363-364,:397-401)
//   dmoc_bt     = -sum_i dble(fl32(fl32(e1v*e3m)*mask)) * dvbt, scanned /1.d6                        (:403-419)
//   zsig0       = fl32(sigmai_dep(T,S,gdept(k)) * tmask)                                             (:443)
//   dgeo, dvgeo, zv_geo : 4-point stencil, vertical recurrence from the bottom                       (:447-461)
//   dmoc_sh     = K1 on zv_geo  -  (weighted sum with the vertical mean of zv_geo), scanned /1.e6    (:464-510)
//   dmoc_ag     = dmoc - dmoc_sh - dmoc_bt                                                           (:516)
#pragma once
#include "common.cuh"
#include "eos_device.cuh"   // exact EOS chain (eos_sigma_exact)

namespace cdfgpu {

// vertical mean of fl32(e3m*zv): one thread per column; `descending` = the level order the reference sums in
__global__ void decomp_vbar_kernel(const float *__restrict__ e3m, const float *__restrict__ zv, size_t nxy, int nzm1,
                                   int descending, float *__restrict__ hdep, int write_hdep, double *__restrict__ dvbt)
{
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nxy; c += (size_t)gridDim.x * blockDim.x) {
        double dv = 0.0;
        float hd = 0.0f;
        for (int n = 0; n < nzm1; ++n) {
            const int k = descending ? nzm1 - 1 - n : n;
            const float e = e3m[(size_t)k * nxy + c];
            dv = __dadd_rn(dv, (double)__fmul_rn(e, zv[(size_t)k * nxy + c]));
            hd = __fadd_rn(hd, e);
        }
        if (write_hdep) hdep[c] = hd;
        const float h = write_hdep ? hd : hdep[c];
        dvbt[c] = (h != 0.0f) ? __ddiv_rn(dv, (double)h) : 0.0;
    }
}

// raw(b,j,k) = -sum_i dble(fl32(fl32(e1v*e3m)*real(mask_b))) * w(i,j): one warp per (j,k) row
template <int NB>
__global__ void decomp_weighted_rows_kernel(const float *__restrict__ area, const int16_t *__restrict__ ibmask,
                                            const double *__restrict__ w, int nx, int ny, int nzm1,
                                            double *__restrict__ raw)
{
    const int lane = threadIdx.x & 31;
    const int nrows = ny * nzm1;
    for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += gridDim.x * (blockDim.x >> 5)) {
        const int k = r / ny, j = r - k * ny;
        double acc[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[b] = 0.0;
        for (int i = lane; i < nx; i += 32) {
            const size_t c = (size_t)j * nx + i;
            const float a = area[(size_t)k * ny * nx + c];
            const double wv = w[c];
#pragma unroll
            for (int b = 0; b < NB; ++b)
                acc[b] = __dadd_rn(acc[b], -__dmul_rn((double)__fmul_rn(a, (float)ibmask[c * NB + b]), wv));
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const double t = warp_sum(acc[b]);
            if (lane == b) raw[((size_t)k * ny + j) * NB + b] = t;
        }
    }
}

// psi(k) = psi(k+1) + (raw(k) - sub(k))/1e6 from the bottom, one thread per (j,b); level nz-1 = 0
__global__ void decomp_scan_kernel(double *__restrict__ d, const double *__restrict__ sub, int nyb, int nz)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nyb) return;
    double psi = 0.0;
    d[(size_t)(nz - 1) * nyb + t] = 0.0;
    for (int k = nz - 2; k >= 0; --k) {
        double v = d[(size_t)k * nyb + t];
        if (sub) v = __dadd_rn(v, -sub[(size_t)k * nyb + t]);
        psi = __dadd_rn(psi, __ddiv_rn(v, 1.0e6));
        d[(size_t)k * nyb + t] = psi;
    }
}

__global__ void decomp_ageo_kernel(const double *__restrict__ tot, const double *__restrict__ sh,
                                   const double *__restrict__ bt, double *__restrict__ ag, size_t n)
{
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x)
        ag[c] = __dadd_rn(__dadd_rn(tot[c], -sh[c]), -bt[c]);
}

// zsig0 = fl32(sigmai_dep(T,S,gdept(k)) * tmask) for one level
__global__ void decomp_sigma_kernel(const float *__restrict__ zt, const float *__restrict__ zs,
                                    const int16_t *__restrict__ tmask, double dlh, double dlref, size_t nxy,
                                    float *__restrict__ sig)
{
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nxy; c += (size_t)gridDim.x * blockDim.x)
        sig[c] = __double2float_rn(__dmul_rn(eos_sigma_exact<false>(zt[c], zs[c], dlh, dlref), (double)tmask[c]));
}

// one level of the geostrophic recurrence (:447-461); border cells of zv keep the last V level read (:357)
__global__ void decomp_geo_level_kernel(const float *__restrict__ sig, const int16_t *__restrict__ um,
                                        const float *__restrict__ e1u, const float *__restrict__ zcoef,
                                        const int16_t *__restrict__ ibmask, int nb, const float *__restrict__ e3m_k,
                                        const float *__restrict__ zv_last, const double *__restrict__ dv_do,
                                        double *__restrict__ dv_up, float *__restrict__ zvgeo_k, int nx, int ny)
{
    const size_t nxy = (size_t)nx * ny;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nxy; c += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(c / nx), i = (int)(c - (size_t)j * nx);
        if (i == 0 || i == nx - 1 || j == 0 || j == ny - 1) { zvgeo_k[c] = zv_last[c]; continue; }
        const size_t cn = c + nx;
        const int su = um[cn - 1] + um[cn] + um[c - 1] + um[c];
        const float zmsv = __fdiv_rn(1.0f, (float)(su > 1 ? su : 1));
        const float t1 = __fdiv_rn(__fmul_rn(__fsub_rn(sig[cn], sig[cn - 1]), (float)um[cn - 1]), e1u[cn - 1]);
        const float t2 = __fdiv_rn(__fmul_rn(__fsub_rn(sig[cn + 1], sig[cn]), (float)um[cn]), e1u[cn]);
        const float t3 = __fdiv_rn(__fmul_rn(__fsub_rn(sig[c], sig[c - 1]), (float)um[c - 1]), e1u[c - 1]);
        const float t4 = __fdiv_rn(__fmul_rn(__fsub_rn(sig[c + 1], sig[c]), (float)um[c]), e1u[c]);
        const double dgeo = (double)__fadd_rn(__fadd_rn(__fadd_rn(t1, t2), t3), t4);
        double d = __dmul_rn((double)zcoef[c], dgeo);
        d = __dmul_rn(d, (double)zmsv);
        d = __dmul_rn(d, (double)ibmask[c * nb]);
        d = __dmul_rn(d, (double)e3m_k[c]);
        const double up = __dadd_rn(dv_do[c], d);
        dv_up[c] = up;
        zvgeo_k[c] = __double2float_rn(__dmul_rn(0.5, __dadd_rn(up, dv_do[c])));
    }
}

}  // namespace cdfgpu
