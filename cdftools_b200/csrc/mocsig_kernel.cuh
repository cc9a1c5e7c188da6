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
// Design: one CTA (16 warps) per latitude row j (all levels), persistent grid of one CTA per SM fed by a ticket.
// The row's (level, 256-cell window) pairs are dealt round-robin to the CTA's warps.  The basin masks are folded at setup into ONE byte per (j,i): the index of the distinct mask tuple
// ("pattern") at that cell, 0 = no basin, 255 = cell excluded (i=1, i=nx, outside the row).  A cell therefore
// costs at most ONE fp64 add into a warp-private shared-memory histogram hist[bin][pattern] (no atomics, bitwise
// reproducible); a lane owns 8 consecutive cells and merges equal (bin,pattern) neighbours in registers first.
// Cells with pattern 0 or p == 0 contribute exactly nothing in the reference and skip the EOS altogether.
// Epilogue per row: patterns -> basins, /1e6, cumulative sum from the densest bin, coalesced store.
//
// Bit-exactness of the bins: eos_sigma_exact() evaluates the reference expression with __dmul_rn/__dadd_rn only
// (never contracted), IEEE sqrt, IEEE fp32 subtract/divide, truncation with x86 CVTTSS2SI semantics.
#pragma once
#include "common.cuh"
#include "../../include/cdf_eos_coeffs.h"

namespace cdfgpu {

constexpr int kSigWarps = 24;
constexpr int kSigThreads = kSigWarps * 32;
constexpr int kSigMaxPat = 32;
constexpr int kSigWinVec = 64;  // one warp step = 64 float4 vectors = 256 cells, 8 consecutive cells per lane

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
    double *__restrict__ out_iso;                                      // -isodep: (ny, nbins, nb) mean isopycnal depth
    const float *__restrict__ gdep;                                    // -isodep: -gdept(k), (nz)
    int *tickets;                                                      // [2]
    int nx, ny, nz, nb, nbins, npat, npat1, pitchw;  // npat incl. the all-zero pattern; npat1 = max(npat-1,1)
    int parity;
    int j_first_global, ny_global;
    int eos;  // CDFGPU_EOS_*
    float sigmin, sigstp, pref;
    float spv, spt, sps;
    double dlh, dlref;
    double inv_sigstp, qmargin;  // fast bin filter: 1/sigstp and the safety distance to a bin edge (q units); <0: off
    double dlref_m1000, nbins_d, half_m_margin;   // dlref - 1000, (double)nbins, 0.5 - qmargin
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

// ---- fast bin: FMA-evaluated polynomial + distance-to-bin-edge guard ------------------------------------------
// The bin index is a monotone step function of sigma.  sigma is evaluated here with fused multiply-adds (half the
// fp64 instructions of the reference's mul/add chain, more ILP); the result differs from the reference value by
// rounding only.  q = (sigma - sigmin)/sigstp is formed in fp64 and compared with the nearest integer: when it is
// farther than `qmargin` (set up on the host from the fp32 roundings of the reference's own bin formula plus a
// generous bound on the evaluation difference, see api_mocsig.inc) the truncation is provably the reference's bin;
// otherwise (about 1e-4 of the cells) the cell is re-evaluated with the exact chain.  Bit-exactness is kept.
#define FM(a, b, c) fma((a), (b), (c))
__device__ __forceinline__ double eos_dlr0_fma(double t, double s)
{
    const double q6 = CE(0, 6, 0);
    const double q5 = FM(CE(1, 5, 0), s, CE(0, 5, 0));
    const double q4 = FM(FM(CE(2, 4, 0), s, CE(1, 4, 0)), s, CE(0, 4, 0));
    const double q3 = FM(FM(FM(CE(3, 3, 0), s, CE(2, 3, 0)), s, CE(1, 3, 0)), s, CE(0, 3, 0));
    const double q2 = FM(FM(FM(FM(CE(4, 2, 0), s, CE(3, 2, 0)), s, CE(2, 2, 0)), s, CE(1, 2, 0)), s, CE(0, 2, 0));
    const double q1 = FM(FM(FM(FM(FM(CE(5, 1, 0), s, CE(4, 1, 0)), s, CE(3, 1, 0)), s, CE(2, 1, 0)), s, CE(1, 1, 0)), s, CE(0, 1, 0));
    const double q0 = FM(FM(FM(FM(FM(FM(CE(6, 0, 0), s, CE(5, 0, 0)), s, CE(4, 0, 0)), s, CE(3, 0, 0)), s, CE(2, 0, 0)), s, CE(1, 0, 0)), s, CE(0, 0, 0));
    return FM(FM(FM(FM(FM(FM(q6, t, q5), t, q4), t, q3), t, q2), t, q1), t, q0);
}
__device__ __forceinline__ double eos_dlr123_fma(double t, double s, double h)
{
    const double a4 = CE(0, 4, 1);
    const double a3 = FM(CE(1, 3, 1), s, CE(0, 3, 1));
    const double a2 = FM(FM(CE(2, 2, 1), s, CE(1, 2, 1)), s, CE(0, 2, 1));
    const double a1 = FM(FM(FM(CE(3, 1, 1), s, CE(2, 1, 1)), s, CE(1, 1, 1)), s, CE(0, 1, 1));
    const double a0 = FM(FM(FM(FM(CE(4, 0, 1), s, CE(3, 0, 1)), s, CE(2, 0, 1)), s, CE(1, 0, 1)), s, CE(0, 0, 1));
    const double r1 = FM(FM(FM(FM(a4, t, a3), t, a2), t, a1), t, a0);
    const double b1 = FM(CE(1, 1, 2), s, CE(0, 1, 2));
    const double b0 = FM(FM(CE(2, 0, 2), s, CE(1, 0, 2)), s, CE(0, 0, 2));
    const double r2 = FM(FM(CE(0, 2, 2), t, b1), t, b0);
    const double r3 = FM(CE(0, 1, 3), t, FM(CE(1, 0, 3), s, CE(0, 0, 3)));
    return FM(FM(r3, h, r2), h, r1) * h;   // (dlr3*h + dlr2)*h + dlr1)*h, to be added to dlr0
}

// sqrt for the fast path: fp32 reciprocal-sqrt seed + two Newton steps in fp64 (branch-free, ~1 ulp).  The exact
// path keeps the correctly rounded __dsqrt_rn; the difference is covered by the guard (|dsigma/ds| < 4e3).
__device__ __forceinline__ double fast_sqrt_pos(double x)
{
    double y = (double)rsqrtf((float)x);
    y = y * FM(-0.5 * x, y * y, 1.5);
    y = y * FM(-0.5 * x, y * y, 1.5);
    const double sx = x * y;
    return FM(FM(-sx, sx, x), 0.5 * y, sx);   // one correction step on sqrt itself
}

template <int EOS, bool SIGMA0>
__device__ __forceinline__ int sigma_bin_fast(float tem, float sal, const SigParams &p)
{
    if (EOS == CDFGPU_EOS_NEUTRAL || p.qmargin < 0.0) return sigma_bin<EOS, SIGMA0>(tem, sal, p);
    const double t = (double)tem * (1.0 / 40.0);
    const double s = fast_sqrt_pos(fabs((double)sal + c_eos.rdeltaS) * c_eos.r1_S0);
    double dlr = eos_dlr0_fma(t, s);
    if (!SIGMA0) dlr += eos_dlr123_fma(t, s, p.dlh);
    const double qa = ((dlr + p.dlref_m1000) - (double)p.sigmin) * p.inv_sigstp;
    // common case: strictly inside the bin range, farther than the margin from a bin edge, salinity not a mask value
    const int ib = __double2int_rz(qa);
    const double fr = qa - (double)ib;
    const bool ok = (qa > 1.0) && (qa < p.nbins_d) && (fabs(fr - 0.5) < p.half_m_margin) && (sal != 0.0f) && (sal != p.sps);
    if (ok) return ib;
    return sigma_bin<EOS, SIGMA0>(tem, sal, p);   // clamps, land, NaN/Inf, near-edge cells: the reference chain
}
#undef FM

__device__ __forceinline__ float scrub(float x, float spval) { return (x == spval) ? 0.0f : x; }

// ---- histogram accumulation ---------------------------------------------------------------------------------
// Every warp owns a private histogram hist[(bin-1)*P + (pattern-1)] in shared memory, so no atomics are needed and
// the result is bitwise reproducible.  A lane owns 8 CONSECUTIVE cells per step; equal (bin,pattern) neighbours are
// merged in registers first (run-length), then the surviving entries of the 32 lanes are combined:
//   * all live lanes share one key (smooth fields): one shuffle-tree sum, one read-modify-write;
//   * otherwise __match_any groups the lanes by key: groups of >= 4 lanes are summed with a shuffle tree each,
//     the remaining small groups write in <= 3 conflict-free rounds ordered by rank.
__device__ __noinline__ void hist_flush(double *hist, int key, double val, int lane)
{
    const bool live = key >= 0;
    const unsigned act = __ballot_sync(kFull, live);
    if (act == 0u) return;  // warp-uniform
    const int first = __ffs(act) - 1;
    const int k0 = __shfl_sync(kFull, key, first);
    if (__ballot_sync(kFull, live && key != k0) == 0u) {
        const double t = warp_sum(live ? val : 0.0);
        if (lane == first) hist[k0] += t;
        __syncwarp();
        return;
    }
    const unsigned peers = __match_any_sync(kFull, key);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    const bool big = live && __popc(peers) >= 4;
    unsigned leaders = __ballot_sync(kFull, big && rank == 0);
    while (leaders) {
        const int L = __ffs(leaders) - 1;
        leaders &= leaders - 1u;
        const int kk = __shfl_sync(kFull, key, L);
        const double t = warp_sum((live && key == kk) ? val : 0.0);
        if (lane == L) hist[kk] += t;
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const bool mine = live && !big && rank == r;
        if (__any_sync(kFull, mine)) {
            if (mine) hist[key] += val;
            __syncwarp();
        }
    }
}

// transport of one cell in fp32, exactly as the reference forms it: scrub, optional bolus velocity, zv*zarea
__device__ __forceinline__ float sig_transport(const SigParams &p, float v, float ve, float a)
{
    v = scrub(v, p.spv);
    if (p.zveiv) v = __fadd_rn(v, ve);
    return __fmul_rn(v, a);
}

// per-warp staging area in shared memory (one 256-cell window)
struct SigStage {
    float ct[8 * 32];       // compacted scrubbed temperature of the cells that need a bin
    float cs[8 * 32];       // compacted scrubbed salinity
    uint16_t bin[8 * 32];   // bin of cell (c, lane), written by the dense EOS pass
    uint8_t slot[8 * 32];   // compacted entry -> c*32+lane
};

template <int EOS, bool SIGMA0, bool ISO>
__global__ void __launch_bounds__(kSigThreads, 1) mocsig_eos_hist_scan_kernel(const SigParams p)
{
    extern __shared__ double s_mem[];
    const int nwarps = blockDim.x >> 5, nthreads = blockDim.x;     // chosen by the host so that shared memory fits
    const int hsize = p.nbins * p.npat1;                            // one private histogram
    constexpr int NH = ISO ? 3 : 1;                                  // histograms per warp: transport [, depth*area, area]
    double *hist_all = s_mem;                                       // [nwarps][NH][nbins][npat1]
    double *comb = hist_all + (size_t)nwarps * NH * hsize;        // [nbins][nb]
    SigStage *stage_all = reinterpret_cast<SigStage *>(comb + (((size_t)p.nbins * p.nb + 1) & ~(size_t)1));  // 16-B aligned
    unsigned *s_poison = reinterpret_cast<unsigned *>(stage_all + nwarps);  // [nbins]
    __shared__ int s_ticket[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nzm1 = p.nz - 1;
    const uint64_t pol = make_evict_first_policy();
    double *hist = hist_all + (size_t)warp * NH * hsize;
    SigStage &st = stage_all[warp];
    const int NV = (p.nx + 6) >> 2;                 // vectors per row, upper bound over the 4 alignments
    const int wpr = (NV + kSigWinVec - 1) / kSigWinVec;  // windows per (level) row
    const int total = nzm1 * wpr;                   // windows of one latitude row j
    int *ticket = p.tickets + p.parity;
    if (blockIdx.x == 0 && tid == 0) p.tickets[p.parity ^ 1] = 0;

    if (tid == 0) s_ticket[0] = atomicAdd(ticket, 1);
    int tsel = 0;
    for (;;) {
        for (int t = tid; t < nwarps * NH * hsize; t += nthreads) hist_all[t] = 0.0;
        for (int t = tid; t < p.nbins; t += nthreads) s_poison[t] = 0u;
        __syncthreads();
        const int j = s_ticket[tsel];
        if (j >= p.ny) break;
        if (tid == 0) s_ticket[tsel ^ 1] = atomicAdd(ticket, 1);  // next row's ticket, latency hidden by this row
        tsel ^= 1;
        const int jg = j + p.j_first_global;
        const bool skip_row = (p.ny_global > 1) && (jg == 0 || jg == p.ny_global - 1);  // jj = 2..npjglo-1 only

        if (!skip_row) {
            // window id = win*(nz-1) + k (level fastest): a warp's static share w, w+nwarps, ... then sweeps all
            // longitude sectors and depths, so land/ocean contrasts do not load some warps of the CTA more than others
            // (with the sector-fastest order and nwarps a multiple of the sectors per row, a warp kept ONE sector).
            int win = 0, k = warp;
            while (k >= nzm1) { k -= nzm1; ++win; }
            for (int w = warp; w < total; w += nwarps) {   // warp-uniform loop: one window of 64 vectors per trip
                const size_t e0 = ((size_t)k * p.ny + j) * (size_t)p.nx;
                const int s = (int)(e0 & 3);
                const int nvec = (s + p.nx + 3) >> 2;
                const int v0 = win * kSigWinVec + 2 * lane;   // this lane's two consecutive vectors = 8 cells
                uint32_t pw0 = 0xffffffffu, pw1 = 0xffffffffu;
                unsigned need = 0u;                            // bit c: cell c needs a bin (contributes)
                float tt[8], ss[8], pr[8], ar[ISO ? 8 : 1];
#pragma unroll
                for (int c = 0; c < 8; ++c) pr[c] = 0.0f;
                // ---- phase A: load, transport, which cells contribute -----------------------------------------
                if (v0 < nvec) {
                    const size_t off = (e0 - s) + 4 * (size_t)v0;
                    const uint32_t *pwp = p.patw + ((size_t)s * p.ny + j) * p.pitchw + v0;
                    const bool two = v0 + 1 < nvec;
                    pw0 = __ldg(pwp);
                    if (two) pw1 = __ldg(pwp + 1);
                    if ((pw0 & pw1) != 0xffffffffu) {
                        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        const float4 *pv = reinterpret_cast<const float4 *>(p.zv + off);
                        const float4 *pa = reinterpret_cast<const float4 *>(p.area + off);
                        const float4 *pt = reinterpret_cast<const float4 *>(p.zt + off);
                        const float4 *ps = reinterpret_cast<const float4 *>(p.zs + off);
                        const float4 va = ld_stream_f4(pv, pol), vb = two ? ld_stream_f4(pv + 1, pol) : z4;
                        const float4 aa = ld_stream_f4(pa, pol), ab = two ? ld_stream_f4(pa + 1, pol) : z4;
                        const float4 ta = ld_stream_f4(pt, pol), tb = two ? ld_stream_f4(pt + 1, pol) : z4;
                        const float4 sa = ld_stream_f4(ps, pol), sb = two ? ld_stream_f4(ps + 1, pol) : z4;
                        float4 ea = z4, eb = z4;
                        if (p.zveiv) {
                            const float4 *pe = reinterpret_cast<const float4 *>(p.zveiv + off);
                            ea = ld_stream_f4(pe, pol);
                            if (two) eb = ld_stream_f4(pe + 1, pol);
                        }
                        pr[0] = sig_transport(p, va.x, ea.x, aa.x); pr[1] = sig_transport(p, va.y, ea.y, aa.y);
                        pr[2] = sig_transport(p, va.z, ea.z, aa.z); pr[3] = sig_transport(p, va.w, ea.w, aa.w);
                        pr[4] = sig_transport(p, vb.x, eb.x, ab.x); pr[5] = sig_transport(p, vb.y, eb.y, ab.y);
                        pr[6] = sig_transport(p, vb.z, eb.z, ab.z); pr[7] = sig_transport(p, vb.w, eb.w, ab.w);
                        tt[0] = ta.x; tt[1] = ta.y; tt[2] = ta.z; tt[3] = ta.w; tt[4] = tb.x; tt[5] = tb.y; tt[6] = tb.z; tt[7] = tb.w;
                        ss[0] = sa.x; ss[1] = sa.y; ss[2] = sa.z; ss[3] = sa.w; ss[4] = sb.x; ss[5] = sb.y; ss[6] = sb.z; ss[7] = sb.w;
                        if (ISO) { ar[0] = aa.x; ar[1] = aa.y; ar[2] = aa.z; ar[3] = aa.w; ar[4] = ab.x; ar[5] = ab.y; ar[6] = ab.z; ar[7] = ab.w; }
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const uint32_t pat = ((c < 4 ? pw0 : pw1) >> (8 * (c & 3))) & 255u;
                            const bool finite = (__float_as_uint(pr[c]) & 0x7f800000u) != 0x7f800000u;
                            // excluded cell, exact zero, or finite transport not covered by any basin: contributes nothing.
                            // With -isodep every covered cell carries area weight into its bin, whatever its transport.
                            if (ISO ? (pat != 255u && (pat != 0u || (!finite && pr[c] != 0.0f)))
                                    : (pat != 255u && pr[c] != 0.0f && (pat != 0u || !finite))) need |= 1u << c;
                        }
                    }
                }
                // ---- compaction: the cells that need the EOS, in (lane, c) order, over the whole warp ------------
                const int mine = __popc(need);
                int base = mine;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int o = __shfl_up_sync(kFull, base, d);
                    if (lane >= d) base += o;
                }
                const int nneed = __shfl_sync(kFull, base, 31);
                base -= mine;
                if (nneed > 0) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (need & (1u << c)) {
                            const int e = base + __popc(need & ((1u << c) - 1u));
                            st.ct[e] = tt[c];
                            st.cs[e] = ss[c];
                            st.slot[e] = (uint8_t)(c * 32 + lane);
                        }
                    __syncwarp();
                    // ---- phase B: dense EOS + bin over the compacted list ---------------------------------------
                    for (int e = lane; e < nneed; e += 32)
                        st.bin[st.slot[e]] =
                            (uint16_t)sigma_bin_fast<EOS, SIGMA0>(scrub(st.ct[e], p.spt), scrub(st.cs[e], p.sps), p);
                    __syncwarp();
                    // ---- phase C: the lane's 8 consecutive cells, all in registers: key/value per cell, equal
                    // (bin,pattern) neighbours merged (a run's sum travels to its last cell), then the surviving entries
                    // are flushed to the warp's private histogram (hist_flush is a warp collective) ----------------
                    int key[8];
                    bool mocadd[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        key[c] = -1;
                        mocadd[c] = false;
                        if (need & (1u << c)) {
                            const uint32_t pat = ((c < 4 ? pw0 : pw1) >> (8 * (c & 3))) & 255u;
                            const int ib = st.bin[c * 32 + lane];
                            bool add = true;
                            if ((__float_as_uint(pr[c]) & 0x7f800000u) == 0x7f800000u) {
                                // NaN/Inf transport: basin b is poisoned iff (0-p)*mask_b is NaN (NaN*x, or Inf*0)
                                unsigned bits = 0u;
                                for (int b = 0; b < p.nb; ++b) {
                                    const double cc = __dmul_rn(0.0 - (double)pr[c], c_patw[pat][b]);
                                    if (cc != cc) bits |= 1u << b;
                                }
                                if (bits) atomicOr(s_poison + (ib - 1), bits);
                                add = !(pr[c] != pr[c] || pat == 0u);  // an Inf with a covering basin still accumulates
                            }
                            mocadd[c] = add;
                            if (add || (ISO && pat != 0u)) key[c] = (ib - 1) * p.npat1 + (int)pat - 1;
                        }
                    }
                    // one pass per histogram: values of the 8 cells, run merge (a run's sum travels to its last cell),
                    // flush of the surviving entries (hist_flush is a warp collective)
                    auto pass = [&](double *h, auto value_of) {
                        int kk[8];
                        double val[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) { kk[c] = key[c]; val[c] = (key[c] >= 0) ? value_of(c) : 0.0; }
#pragma unroll
                        for (int c = 1; c < 8; ++c)
                            if (kk[c] == kk[c - 1] && kk[c] >= 0) {
                                val[c] += val[c - 1];
                                kk[c - 1] = -1;
                            }
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (__any_sync(kFull, kk[c] >= 0)) hist_flush(h, kk[c], val[c], lane);
                    };
                    pass(hist, [&](int c) { return mocadd[c] ? 0.0 - (double)pr[c] : 0.0; });
                    if (ISO) {   // cdfmocsig.f90:427-428: gdep(jk)*itmask*zarea and itmask*zarea, REAL(4) chains
                        const float gk = p.gdep[k];
                        pass(hist + hsize, [&](int c) {
                            const float itm = (scrub(ss[c], p.sps) == p.sps) ? 0.0f : 1.0f;
                            return (double)__fmul_rn(__fmul_rn(gk, itm), ar[c]);
                        });
                        pass(hist + 2 * hsize, [&](int c) {
                            const float itm = (scrub(ss[c], p.sps) == p.sps) ? 0.0f : 1.0f;
                            return (double)__fmul_rn(itm, ar[c]);
                        });
                    }
                }
                k += nwarps;
                while (k >= nzm1) { k -= nzm1; ++win; }
            }
        }
        __syncthreads();
        // private histograms -> one (fixed order: deterministic), patterns -> basins, /1e6, poison handling
        auto combine = [&](int which, int bin, int b) {
            double h = 0.0;
            for (int q = 1; q < p.npat; ++q) {
                const double wgt = c_patw[q][b];
                if (wgt != 0.0) {
                    double hq = 0.0;
                    for (int w = 0; w < nwarps; ++w) hq += hist_all[((size_t)w * NH + which) * hsize + bin * p.npat1 + q - 1];
                    h += hq * wgt;
                }
            }
            return h;
        };
        for (int t = tid; t < p.nbins * p.nb; t += nthreads) {
            const int bin = t / p.nb, b = t - bin * p.nb;
            double h = combine(0, bin, b);
            if ((s_poison[bin] >> b) & 1u) h = __longlong_as_double(0x7ff8000000000000LL);  // NaN transport
            comb[t] = h / 1.0e6;
            if (ISO) {   // depi = depi / wdep where wdep /= 0, else rp_spval (cdfmocsig.f90:463-469); no cumsum
                const double d = combine(1, bin, b), w = combine(2, bin, b);
                p.out_iso[(size_t)j * p.nbins * p.nb + t] = (w != 0.0) ? d / w : 99999.0;
            }
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
        for (int t = tid; t < p.nbins * p.nb; t += nthreads) o[t] = comb[t];
        __syncthreads();
    }
}

// Diagnostic kernel for the parity tests: ibin of every cell of a record, same device function as the fused kernel.
template <int EOS, bool SIGMA0>
__global__ void mocsig_bins_kernel(const SigParams p, int32_t *__restrict__ ibin, size_t n)
{
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (size_t)gridDim.x * blockDim.x)
        ibin[c] = sigma_bin_fast<EOS, SIGMA0>(scrub(p.zt[c], p.spt), scrub(p.zs[c], p.sps), p);
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
