// mocsig_kernel.cuh -- K2: density-space MOC.  Replaces the loop nest src/cdfmocsig.f90:366-475 and the
// sigmai_dep / sigmantr calls it makes (src/eos.f90:842-882, :661-684).
//
// Per level-cell (i in 2..nx-1, j in 2..ny-1 of the reference's 1-based indices):
//   zv,zt,zs : == missing value -> 0 ; zv += zveiv (-eiv)                              (cdfmocsig.f90:375-384)
//   p     = fl32(zv * fl32(e1v*e3v))                                                   (:390, :420)
//   dens  = sigmai_dep(zt,zs,pref)  fp64, no FMA contraction                           (eos.f90:842-882)
//   z     = fl32(dens * itmask) ; ibin = clamp(INT(fl32(fl32(z-sigmin)/sigstp)),1,nbins) (:399-403)
//   H(b,ibin,j) += (0.d0 - dble(p)) * ibmask(b,i,j)                                    (:420, :439)
// then  psi(b,nbins,j)=H/1e6 ; psi(b,bin,j) = psi(b,bin+1,j) + H(b,bin,j)/1e6           (:471-475)
//
// Design: one CTA per latitude row j (all levels), persistent grid of one CTA per SM fed by a ticket.  The
// (level, 16-byte vector) pairs of the row are dealt round-robin to the CTA's threads, so the work is balanced to
// one vector.  The basin masks are folded at setup into ONE byte per (j,i): the index of the distinct mask tuple
// ("pattern") at that cell, 0 = no basin, 255 = cell excluded (i=1, i=nx, outside the row).  A cell therefore
// costs ONE fp64 add into a shared-memory histogram hist[bin][pattern]; runs of equal (bin,pattern) along i are
// first merged with a segmented warp scan so that smooth fields issue one shared atomic per run.  Cells with
// pattern 0 or p == 0 contribute exactly nothing in the reference and skip the EOS altogether.
// Epilogue per row: patterns -> basins, /1e6, cumulative sum from the densest bin, coalesced store.
//
// Bit-exactness of the bins: eos_sigma_exact() evaluates the reference expression with __dmul_rn/__dadd_rn only
// (never contracted), IEEE sqrt, IEEE fp32 subtract/divide, truncation with x86 CVTTSS2SI semantics.
#pragma once
#include "common.cuh"
#include "../../include/cdf_eos_coeffs.h"

namespace cdfgpu {

constexpr int kSigThreads = 1024;
constexpr int kSigMaxPat = 32;

struct EosConst {
    double c[CDF_EOS_NCOEF];
    double r0[6];
    double rdeltaS, r1_S0;
};
__constant__ EosConst c_eos;
__constant__ double c_patw[kSigMaxPat][CDFGPU_MAX_BASINS];  // weight of pattern p for basin b (= mask value)

struct SigParams {
    const float *__restrict__ zv, *__restrict__ zt, *__restrict__ zs;  // (nz-1, ny, nx) raw file values
    const float *__restrict__ zveiv;                                   // optional, may be null
    const float *__restrict__ area;                                    // (nz-1, ny, nx) fl32(e1v*e3v)
    const uint32_t *__restrict__ patw;                                 // [4][ny][pitchw] pattern bytes, shifted copies
    double *__restrict__ out;                                          // (ny, nbins, nb)
    int *tickets;                                                      // [2]
    int nx, ny, nz, nb, nbins, npat, pitchw;
    int parity;
    int j_first_global, ny_global;
    int eos;  // CDFGPU_EOS_*
    float sigmin, sigstp, pref;
    float spv, spt, sps;
    double dlh, dlref;
};

// ---- equation of state ----------------------------------------------------------------------------------------
#define CE(i, j, k) c_eos.c[I_EOS##i##j##k]
#define DM(a, b) __dmul_rn((a), (b))
#define DA(a, b) __dadd_rn((a), (b))

// dlr0 of eos.f90:871-877 -- the only polynomial needed when pref == 0.
__device__ __forceinline__ double eos_dlr0(double t, double s)
{
    double a = DA(DA(DM(CE(0, 6, 0), t), DM(CE(1, 5, 0), s)), CE(0, 5, 0));
    a = DA(DA(DM(a, t), DM(DA(DM(CE(2, 4, 0), s), CE(1, 4, 0)), s)), CE(0, 4, 0));
    a = DA(DA(DM(a, t), DM(DA(DM(DA(DM(CE(3, 3, 0), s), CE(2, 3, 0)), s), CE(1, 3, 0)), s)), CE(0, 3, 0));
    a = DA(DA(DM(a, t), DM(DA(DM(DA(DM(DA(DM(CE(4, 2, 0), s), CE(3, 2, 0)), s), CE(2, 2, 0)), s), CE(1, 2, 0)), s)),
           CE(0, 2, 0));
    a = DA(DA(DM(a, t),
              DM(DA(DM(DA(DM(DA(DM(DA(DM(CE(5, 1, 0), s), CE(4, 1, 0)), s), CE(3, 1, 0)), s), CE(2, 1, 0)), s), CE(1, 1, 0)), s)),
           CE(0, 1, 0));
    a = DA(DA(DM(a, t),
              DM(DA(DM(DA(DM(DA(DM(DA(DM(DA(DM(CE(6, 0, 0), s), CE(5, 0, 0)), s), CE(4, 0, 0)), s), CE(3, 0, 0)), s), CE(2, 0, 0)), s),
                    CE(1, 0, 0)),
                 s)),
           CE(0, 0, 0));
    return a;
}
__device__ __forceinline__ double eos_dlr1(double t, double s)
{
    double a = DA(DA(DM(CE(0, 4, 1), t), DM(CE(1, 3, 1), s)), CE(0, 3, 1));
    a = DA(DA(DM(a, t), DM(DA(DM(CE(2, 2, 1), s), CE(1, 2, 1)), s)), CE(0, 2, 1));
    a = DA(DA(DM(a, t), DM(DA(DM(DA(DM(CE(3, 1, 1), s), CE(2, 1, 1)), s), CE(1, 1, 1)), s)), CE(0, 1, 1));
    a = DA(DA(DM(a, t), DM(DA(DM(DA(DM(DA(DM(CE(4, 0, 1), s), CE(3, 0, 1)), s), CE(2, 0, 1)), s), CE(1, 0, 1)), s)),
           CE(0, 0, 1));
    return a;
}
__device__ __forceinline__ double eos_dlr2(double t, double s)
{
    double a = DA(DA(DM(CE(0, 2, 2), t), DM(CE(1, 1, 2), s)), CE(0, 1, 2));
    a = DA(DA(DM(a, t), DM(DA(DM(CE(2, 0, 2), s), CE(1, 0, 2)), s)), CE(0, 0, 2));
    return a;
}
__device__ __forceinline__ double eos_dlr3(double t, double s)
{
    return DA(DA(DM(CE(0, 1, 3), t), DM(CE(1, 0, 3), s)), CE(0, 0, 3));
}

// sigmai_dep for one cell (eos.f90:848-882).  SIGMA0: pref == 0, then dlh == 0 and dlr == dlr0 exactly.
template <bool SIGMA0>
__device__ __forceinline__ double eos_sigma_exact(float tem, float sal, double dlh, double dlref)
{
    const double t = DM((double)tem, 1.0 / 40.0);
    const double s = __dsqrt_rn(DM(fabs(DA((double)sal, c_eos.rdeltaS)), c_eos.r1_S0));
    double dlr = eos_dlr0(t, s);
    if (!SIGMA0) {
        const double r1 = eos_dlr1(t, s), r2 = eos_dlr2(t, s), r3 = eos_dlr3(t, s);
        dlr = DA(DM(DA(DM(DA(DM(r3, dlh), r2), dlh), r1), dlh), dlr);
    }
    const double dltm = (sal == 0.0f) ? 0.0 : 1.0;
    return DM(DA(DA(dlr, dlref), -1000.0), dltm);
}

// sigmantr for one cell (eos.f90:663-682).
__device__ __forceinline__ double eos_sigma_neutral(float tem, float sal)
{
    const double t = (double)tem, s = (double)sal;
    const double sr = __dsqrt_rn(fabs(s));
    const double r1 = DA(DM(DA(DM(DA(DM(-4.3159255086706703e-4, t), 8.1157118782170051e-2), t), 2.2280832068441331e-1), t),
                         1002.3063688892480e0);
    // -a*s - b*t - c  ==  ((-a*s) - (b*t)) - c
    const double r2 = DM(DA(DA(DM(-1.7052298331414675e-7, s), -DM(3.1710675488863952e-3, t)), -1.0304537539692924e-4), s);
    const double r3 = DA(DM(DA(DM(DA(DM(DA(DM(-2.3850178558212048e-9, t), -1.6212552470310961e-7), t), 7.8717799560577725e-5), t),
                               4.3907692647825900e-5), t), 1.0);
    const double r4 = DM(DA(DM(DA(DM(DM(-2.2744455733317707e-9, t), t), 6.0399864718597388e-6), t), -5.1268124398160734e-4), s);
    const double r5 = DM(DM(DA(DM(DM(-1.3409379420216683e-9, t), t), -3.6138532339703262e-5), s), sr);
    return DA(__ddiv_rn(DA(r1, r2), DA(DA(r3, r4), r5)), -1000.0);
}

// bin index 1..nbins of one cell from scrubbed T,S (cdfmocsig.f90:393-403).
template <int EOS, bool SIGMA0>
__device__ __forceinline__ int sigma_bin(float tem, float sal, const SigParams &p)
{
    const double dens = (EOS == CDFGPU_EOS_NEUTRAL) ? eos_sigma_neutral(tem, sal)
                                                    : eos_sigma_exact<SIGMA0>(tem, sal, p.dlh, p.dlref);
    const double itm = (sal == p.sps) ? 0.0 : 1.0;  // itmask, evaluated on the scrubbed salinity (:393-394)
    const float z = __double2float_rn(DM(dens, itm));
    const float q = __fdiv_rn(__fsub_rn(z, p.sigmin), p.sigstp);
    // INT(): truncation toward zero; out-of-range and NaN give INT_MIN like x86 CVTTSS2SI (what gfortran emits)
    int ib = (q >= 2147483648.0f || q < -2147483648.0f || q != q) ? (int)0x80000000 : __float2int_rz(q);
    ib = max(ib, 1);
    ib = min(ib, p.nbins);
    return ib;
}

__device__ __forceinline__ float scrub(float x, float spval) { return (x == spval) ? 0.0f : x; }

// ---- histogram accumulation ---------------------------------------------------------------------------------
// One item per lane: key (>=0: bin*kSigMaxPat+pattern, <0: nothing), val.  Runs of equal keys over adjacent lanes
// are summed with a segmented scan; the last lane of each run issues one shared-memory atomic.
__device__ __forceinline__ void hist_add_runs(double *hist, int key, double val, int lane)
{
    const int prev = __shfl_up_sync(kFull, key, 1);
    const unsigned heads = __ballot_sync(kFull, lane == 0 || key != prev || key < 0);
    const int start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const double o = __shfl_up_sync(kFull, val, d);
        if (lane - d >= start) val += o;
    }
    const bool tail = (lane == 31) || ((heads >> (lane + 1)) & 1u);
    if (tail && key >= 0) atomicAdd(hist + key, val);
}

template <int EOS, bool SIGMA0>
__device__ __forceinline__ void sig_cell(const SigParams &p, float v, float ve, float t, float s, float a, uint32_t pat,
                                         int &key, double &val, unsigned *s_poison)
{
    key = -1;
    val = 0.0;
    v = scrub(v, p.spv);
    if (p.zveiv) v = __fadd_rn(v, ve);
    const float pr = __fmul_rn(v, a);
    if (pat == 255u || pr == 0.0f) return;         // excluded cell, or an exact zero contribution
    const bool finite = (__float_as_uint(pr) & 0x7f800000u) != 0x7f800000u;
    if (pat == 0u && finite) return;               // no basin covers the cell: (0-p)*0 == 0
    t = scrub(t, p.spt);
    s = scrub(s, p.sps);
    const int ib = sigma_bin<EOS, SIGMA0>(t, s, p);
    if (!finite) {  // NaN/Inf transport: basin b is poisoned iff (0-p)*mask_b is NaN (NaN*x, or Inf*0)
        unsigned bits = 0u;
        for (int b = 0; b < p.nb; ++b) {
            const double c = __dmul_rn(0.0 - (double)pr, c_patw[pat][b]);
            if (c != c) bits |= 1u << b;
        }
        if (bits) atomicOr(s_poison + (ib - 1), bits);
        if (pr != pr || pat == 0u) return;  // an Inf with a covering basin still accumulates below
    }
    key = (ib - 1) * kSigMaxPat + (int)pat;
    val = 0.0 - (double)pr;
}

template <int EOS, bool SIGMA0>
__global__ void __launch_bounds__(kSigThreads, 1) mocsig_eos_hist_scan_kernel(const SigParams p)
{
    extern __shared__ double s_mem[];
    double *hist = s_mem;                                           // [nbins][kSigMaxPat]
    double *comb = hist + (size_t)p.nbins * kSigMaxPat;             // [nbins][nb]
    unsigned *s_poison = reinterpret_cast<unsigned *>(comb + (size_t)p.nbins * p.nb);  // [nbins]
    __shared__ int s_ticket[2];

    const int tid = threadIdx.x, lane = tid & 31;
    const int nzm1 = p.nz - 1;
    const uint64_t pol = make_evict_first_policy();
    const int NV = (p.nx + 6) >> 2;  // vectors per row, upper bound over the 4 alignments
    int *ticket = p.tickets + p.parity;
    if (blockIdx.x == 0 && tid == 0) p.tickets[p.parity ^ 1] = 0;

    if (tid == 0) s_ticket[0] = atomicAdd(ticket, 1);
    int tsel = 0;
    const int total = nzm1 * NV;  // (level, vector) pairs of one latitude row
    for (;;) {
        for (int t = tid; t < p.nbins * kSigMaxPat; t += kSigThreads) hist[t] = 0.0;
        for (int t = tid; t < p.nbins; t += kSigThreads) s_poison[t] = 0u;
        __syncthreads();
        const int j = s_ticket[tsel];
        if (j >= p.ny) break;
        if (tid == 0) s_ticket[tsel ^ 1] = atomicAdd(ticket, 1);  // next row's ticket, latency hidden by this row
        tsel ^= 1;
        const int jg = j + p.j_first_global;
        const bool skip_row = (p.ny_global > 1) && (jg == 0 || jg == p.ny_global - 1);  // jj = 2..npjglo-1 only

        if (!skip_row) {
            int k = 0, vi = tid;
            while (vi >= NV) { vi -= NV; ++k; }
            // warp-uniform trip count: lanes past the end idle through the collectives
            for (int t = tid; t - lane < total; t += kSigThreads) {
                int key[4] = {-1, -1, -1, -1};
                double val[4] = {0.0, 0.0, 0.0, 0.0};
                if (t < total) {
                    const size_t e0 = ((size_t)k * p.ny + j) * (size_t)p.nx;
                    const int s = (int)(e0 & 3);
                    const int nvec = (s + p.nx + 3) >> 2;
                    if (vi < nvec) {
                        const size_t off = (e0 - s) + 4 * (size_t)vi;
                        const uint32_t pw = __ldg(p.patw + ((size_t)s * p.ny + j) * p.pitchw + vi);
                        if (pw != 0xffffffffu) {
                            const float4 v4 = ld_stream_f4(reinterpret_cast<const float4 *>(p.zv + off), pol);
                            const float4 a4 = ld_stream_f4(reinterpret_cast<const float4 *>(p.area + off), pol);
                            const float4 t4 = ld_stream_f4(reinterpret_cast<const float4 *>(p.zt + off), pol);
                            const float4 s4 = ld_stream_f4(reinterpret_cast<const float4 *>(p.zs + off), pol);
                            float4 e4 = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (p.zveiv) e4 = ld_stream_f4(reinterpret_cast<const float4 *>(p.zveiv + off), pol);
                            sig_cell<EOS, SIGMA0>(p, v4.x, e4.x, t4.x, s4.x, a4.x, pw & 255u, key[0], val[0], s_poison);
                            sig_cell<EOS, SIGMA0>(p, v4.y, e4.y, t4.y, s4.y, a4.y, (pw >> 8) & 255u, key[1], val[1], s_poison);
                            sig_cell<EOS, SIGMA0>(p, v4.z, e4.z, t4.z, s4.z, a4.z, (pw >> 16) & 255u, key[2], val[2], s_poison);
                            sig_cell<EOS, SIGMA0>(p, v4.w, e4.w, t4.w, s4.w, a4.w, pw >> 24, key[3], val[3], s_poison);
                        }
                    }
                }
                if (__any_sync(kFull, (key[0] & key[1] & key[2] & key[3]) >= 0)) {
                    // merge the lane's own 4 cells when they share a key (the common case in smooth fields)
                    const bool same = key[0] == key[1] && key[1] == key[2] && key[2] == key[3];
                    if (__all_sync(kFull, same)) {
                        hist_add_runs(hist, key[0], (val[0] + val[1]) + (val[2] + val[3]), lane);
                    } else {
#pragma unroll
                        for (int c = 0; c < 4; ++c) hist_add_runs(hist, key[c], val[c], lane);
                    }
                }
                vi += kSigThreads;
                while (vi >= NV) { vi -= NV; ++k; }
            }
        }
        __syncthreads();
        // patterns -> basins, /1e6 (dmoc/1.e6), poison handling
        for (int t = tid; t < p.nbins * p.nb; t += kSigThreads) {
            const int bin = t / p.nb, b = t - bin * p.nb;
            double h = 0.0;
            for (int q = 1; q < p.npat; ++q) {
                const double w = c_patw[q][b];
                if (w != 0.0) h += hist[bin * kSigMaxPat + q] * w;
            }
            if ((s_poison[bin] >> b) & 1u) h = __longlong_as_double(0x7ff8000000000000LL);  // NaN transport
            comb[t] = h / 1.0e6;
        }
        __syncthreads();
        if (tid < p.nb) {  // integrate from the densest bin, sequentially as the reference does (:472-475)
            double psi = comb[(p.nbins - 1) * p.nb + tid];
            for (int bin = p.nbins - 2; bin >= 0; --bin) {
                psi = psi + comb[bin * p.nb + tid];
                comb[bin * p.nb + tid] = psi;
            }
        }
        __syncthreads();
        double *o = p.out + (size_t)j * p.nbins * p.nb;
        for (int t = tid; t < p.nbins * p.nb; t += kSigThreads) o[t] = comb[t];
        __syncthreads();
    }
}

// Diagnostic kernel for the parity tests: ibin of every cell of a record, same device function as the fused kernel.
template <int EOS, bool SIGMA0>
__global__ void mocsig_bins_kernel(const SigParams p, int32_t *__restrict__ ibin, size_t n)
{
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x)
        ibin[c] = sigma_bin<EOS, SIGMA0>(scrub(p.zt[c], p.spt), scrub(p.zs[c], p.sps), p);
}

// setup kernel: area = fl32(e1v * e3v) (cdfmocsig.f90:390), e3v NOT masked.
__global__ void sig_prep_area_kernel(const float *__restrict__ e1v, const float *__restrict__ e3v,
                                     float *__restrict__ area, size_t nxy)
{
    const size_t k = blockIdx.y;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nxy; c += (size_t)gridDim.x * blockDim.x)
        area[k * nxy + c] = __fmul_rn(e1v[c], e3v[k * nxy + c]);
}

#undef CE
#undef DM
#undef DA

}  // namespace cdfgpu
