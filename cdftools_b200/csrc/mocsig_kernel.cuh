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

#ifndef CDF_SIG_WARPS
#define CDF_SIG_WARPS 16
#endif
constexpr int kSigWarps = CDF_SIG_WARPS;
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
    const uint8_t *__restrict__ cov8;                                  // [4][ny][pitchw/2] coverage bits of each 8-cell group
    double *__restrict__ out;                                          // (ny, nbins, nb)
    double *__restrict__ out_iso;                                      // -isodep: (ny, nbins, nb) mean isopycnal depth
    const float *__restrict__ gdep;                                    // -isodep: -gdept(k), (nz)
    int *tickets;                                                      // [3] row tickets, one per launch generation
    int pdl;                                                           // launched with programmatic stream serialization
    int nx, ny, nz, nb, nbins, npat, npat1, pitchw;  // npat incl. the all-zero pattern; npat1 = max(npat-1,1)
    unsigned npat1_magic;                            // ceil(2^32 / npat1): k / npat1 == umulhi(k, magic) for k < 2^32 / npat1
    int parity;
    int j_first_global, ny_global;
    int eos;  // CDFGPU_EOS_*
    float sigmin, sigstp, pref;
    float spv, spt, sps;
    double dlh, dlref;
    double inv_sigstp, qmargin;  // fast bin filter: 1/sigstp and the safety distance to a bin edge (q units); <0: off
    double dlref_m1000, nbins_d, half_m_margin;   // dlref - 1000, (double)nbins, 0.5 - qmargin
    double qoffset;                               // (dlref - 1000 - sigmin) / sigstp
    size_t patplane;                              // words per pre-shifted pattern plane = ny * pitchw
    int scrub_ts;                                 // T or S missing value is not zero
    int scrub_v;                                  // V missing value is not zero
    int use_table;                                // per-warp flush table allocated (hist_flush_table)
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

// One 256-cell window (level k, window `win` of latitude row j), fully general: any mix of covered / uncovered cells,
// NaN / Inf transports (poison), -isodep.  Cell-granular compaction through the warp's staging area.
template <int EOS, bool SIGMA0, bool ISO>
__device__ __forceinline__ void sig_window_general(const SigParams &p, SigStage &st, double *hist, unsigned *s_poison,
                                                   int hsize, int j, int k, int win, int lane, uint64_t pol)
{
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
}


// ---- group-granular compaction (the production path) ---------------------------------------------------------
// A lane's 8 consecutive cells form a GROUP.  The sweep over a window reads only the pattern bytes, V and the area,
// forms the fp32 transports and queues the groups that hold at least one contributing cell in a warp-private ring
// (descriptor + the 8 transports).  As soon as 32 groups are queued the warp runs one DENSE iteration: every lane
// pops one group, loads its T and S, evaluates the 8 bins in registers, merges equal (bin,pattern) neighbours and
// flushes to the warp's private histogram.  T/S of windows without contributing cells are never read, the EOS runs
// on full warps whatever the land/ocean geometry, and nothing per cell goes through shared memory.
// Windows holding a NaN/Inf transport (poison semantics) take sig_window_general() instead.
#ifndef CDF_SIG_EOS_BATCH
#define CDF_SIG_EOS_BATCH 4
#endif
constexpr int kSigEosBatch = CDF_SIG_EOS_BATCH;   // cells evaluated together (coefficient-major)
constexpr int kSigRing = 64;                      // >= 31 left over + 32 pushed
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
struct SigQueue {
    uint4 desc[kSigRing];   // x: float4-vector index of the group's first vector, y: need | two<<8 | k<<16, z/w: pattern words
    float4 pa[kSigRing];    // transports of cells 0..3
    float4 pb[kSigRing];    // transports of cells 4..7
};
union SigScratch {
    SigStage st;
    SigQueue q;
};
#ifndef CDF_SIG_TABLE_MIN_LANES
#define CDF_SIG_TABLE_MIN_LANES 4
#endif
constexpr int kSigTableMinLanes = CDF_SIG_TABLE_MIN_LANES;
#ifndef CDF_SIG_TAB_WIDE
#define CDF_SIG_TAB_WIDE 1
#endif
// hist_flush_table(): 2 KB per warp, allocated when SigParams::use_table.  8 bins x 32 columns: every lane has its own
// column, one round per band (default; sigma0 ORCA025 white noise 0.64 ms), or with CDF_SIG_TAB_WIDE=0 16 bins x 16 columns,
// the half-warps taking turns (0.68 ms)
constexpr int kSigTabRows = CDF_SIG_TAB_WIDE ? 8 : 16, kSigTabCols = CDF_SIG_TAB_WIDE ? 32 : 16;

// This warp's flush table: recomputed from the kernel's shared-memory layout where it is needed (rare path) instead of
// being carried in a register through the row loop (the kernel sits at the 128-register cap).
template <bool ISO>
__device__ __forceinline__ double *sig_tab_ptr(const SigParams &p)
{
    extern __shared__ double s_mem[];
    const size_t nwarps = blockDim.x >> 5;
    double *comb = s_mem + nwarps * (ISO ? 3 : 1) * ((size_t)p.nbins * p.npat1);
    SigScratch *stage_all = reinterpret_cast<SigScratch *>(comb + (((size_t)p.nbins * p.nb + 1) & ~(size_t)1));
    return reinterpret_cast<double *>(stage_all + nwarps) + (size_t)(threadIdx.x >> 5) * (kSigTabRows * kSigTabCols);
}

// Flush of a dense step in which MANY lanes hold several runs each (fields that change density class from cell to cell).
// hist_flush() is a chain of dependent warp collectives per round of entries (ballot, match, one shuffle tree per key
// group): with 8 rounds and only 4 warps per scheduler to hide them, those latencies -- not the EOS -- are what such a
// step costs.  Here every lane adds its (up to 8) entries of one mask pattern into its own column of a small per-warp
// table tab[bin - blo][lane]: no two lanes ever touch the same
// word, so there is nothing to resolve; then lane r sums row r and adds it to the warp's private
// histogram.  No atomics, fixed order of additions: bitwise reproducible.  The entries are flushed band by band (a band =
// kSigTabRows consecutive bins from the lowest bin still pending: a step that holds groups of two windows, i.e. of two
// levels, has two clusters of bins), one pass per mask pattern present in the band.  The table is left zeroed.
// Shared memory is not free here: the carve-out steps (... 132, 164, 196, 228 KB) take the space from L1, and L1 bounds
// the bytes the sweep keeps in flight (sigma0 / 104 bins, smooth fields: 0.515 ms with 113 KB, 0.520 with 145 KB, 0.565
// with 177 KB), so the host enables the table only while the kernel stays within the 164 KB step (api_mocsig.inc).
__device__ __forceinline__ bool hist_flush_table(double *hist, double *tab, const int (&kk)[8], const double (&val)[8], int lane,
                                                 int npat1, unsigned npat1_magic)
{
    int bn[8], pt[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int k = max(kk[c], 0);
        bn[c] = (npat1 == 1) ? k : (int)__umulhi((unsigned)k, npat1_magic);   // k / npat1 (magic = ceil(2^32 / npat1))
        pt[c] = (kk[c] >= 0) ? k - bn[c] * npat1 : -1;                        // -1: no entry (or already flushed)
    }
    double *const mycol = tab + (lane & (kSigTabCols - 1));
    // A dense step may hold groups of two or more windows, i.e. of different LEVELS (a warp's consecutive windows are
    // nwarps levels apart): their bins form separate clusters.  Each trip of this loop flushes the band of
    // kSigTabRows bins that starts at the lowest bin still pending, one pass per mask pattern present in the band.
    for (;;) {
        int lo = 0x7fffffff;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (pt[c] >= 0) lo = min(lo, bn[c]);
        lo = __reduce_min_sync(kFull, lo);
        if (lo == 0x7fffffff) break;               // warp-uniform: everything is flushed
        unsigned pm = 0u;
        int top = lo;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (pt[c] >= 0 && bn[c] - lo < kSigTabRows) { pm |= 1u << pt[c]; top = max(top, bn[c]); }
        pm = __reduce_or_sync(kFull, pm);
        const int span = __reduce_max_sync(kFull, top) - lo;
        double *col = mycol - (size_t)lo * kSigTabCols;
        while (pm) {                               // warp-uniform
            const int q = __ffs(pm) - 1;
            pm &= pm - 1u;
#if CDF_SIG_TAB_WIDE
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (pt[c] == q && bn[c] - lo < kSigTabRows) col[bn[c] * kSigTabCols] += val[c];
            __syncwarp();
            {   // four lanes per row, 8 columns each
                const int r = lane >> 2, qtr = lane & 3;
                double s = 0.0;
                if (r <= span) {
                    double *row = tab + r * kSigTabCols + qtr * 8;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int e = (i + r) & 7;
                        s += row[e];
                        row[e] = 0.0;
                    }
                }
                s += __shfl_xor_sync(kFull, s, 1);
                s += __shfl_xor_sync(kFull, s, 2);
                if (qtr == 0 && r <= span) hist[(lo + r) * npat1 + q] += s;
            }
            __syncwarp();
#else
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                if ((lane >> 4) == half) {
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        if (pt[c] == q && bn[c] - lo < kSigTabRows) col[bn[c] * kSigTabCols] += val[c];
                }
                __syncwarp();
            }
            if (lane <= span) {
                double *row = tab + lane * kSigTabCols;
                double s = 0.0;
#pragma unroll
                for (int i = 0; i < kSigTabCols; ++i) {
                    const int e = (i + lane) & (kSigTabCols - 1);   // skewed: conflict-free
                    s += row[e];
                    row[e] = 0.0;
                }
                hist[(lo + lane) * npat1 + q] += s;
            }
            __syncwarp();
#endif
        }
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (bn[c] - lo < kSigTabRows) pt[c] = -1;   // this band is done
    }
    return true;
}

#define FM(a, b, c) fma((a), (b), (c))
// Fast bins of four cells at once, coefficient-major: sm_100 has no constant-bank operands on DFMA, every coefficient
// costs a load instruction, so each one is fetched once and applied to the four cells.  Per cell the operations and
// their order are those of eos_dlr0_fma / eos_dlr123_fma.  Returns the mask of cells whose bin is provably the
// reference's (see sigma_bin_fast).
template <bool SIGMA0, int NB>
__device__ __forceinline__ unsigned sigma_bins_try(const float *tem, const float *sal, const SigParams &p, int *ib)
{
    double t[NB], s[NB], acc[NB], q[NB];
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        t[c] = (double)tem[c] * (1.0 / 40.0);
        const double x = fabs((double)sal[c] + c_eos.rdeltaS) * c_eos.r1_S0;
        float y0;
        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"((float)x));
        const double y = (double)y0;                // 1/sqrt(x) to 2^-22
        const double sx = x * y;                    // sqrt(x) to 2^-22
        s[c] = FM(FM(-sx, sx, x), 0.5 * y, sx);     // one Newton step on the root: ~2^-43 relative (|dsigma/ds| < 4e3: far
                                                    // below the 1e-9 kg/m3 evaluation slop the margin allows for)
    }
#define ALL4(expr) _Pragma("unroll") for (int c = 0; c < NB; ++c) { expr; }
    ALL4(q[c] = FM(CE(1, 5, 0), s[c], CE(0, 5, 0)))
    ALL4(acc[c] = FM(CE(0, 6, 0), t[c], q[c]))
    ALL4(q[c] = FM(CE(2, 4, 0), s[c], CE(1, 4, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(0, 4, 0)))
    ALL4(acc[c] = FM(acc[c], t[c], q[c]))
    ALL4(q[c] = FM(CE(3, 3, 0), s[c], CE(2, 3, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(1, 3, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(0, 3, 0)))
    ALL4(acc[c] = FM(acc[c], t[c], q[c]))
    ALL4(q[c] = FM(CE(4, 2, 0), s[c], CE(3, 2, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(2, 2, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(1, 2, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(0, 2, 0)))
    ALL4(acc[c] = FM(acc[c], t[c], q[c]))
    ALL4(q[c] = FM(CE(5, 1, 0), s[c], CE(4, 1, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(3, 1, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(2, 1, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(1, 1, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(0, 1, 0)))
    ALL4(acc[c] = FM(acc[c], t[c], q[c]))
    ALL4(q[c] = FM(CE(6, 0, 0), s[c], CE(5, 0, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(4, 0, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(3, 0, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(2, 0, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(1, 0, 0)))
    ALL4(q[c] = FM(q[c], s[c], CE(0, 0, 0)))
    ALL4(acc[c] = FM(acc[c], t[c], q[c]))
    if (!SIGMA0) {   // + ((dlr3*h + dlr2)*h + dlr1)*h
        double r1[NB];
        const double h = p.dlh;
        ALL4(q[c] = FM(CE(1, 3, 1), s[c], CE(0, 3, 1)))
        ALL4(r1[c] = FM(CE(0, 4, 1), t[c], q[c]))
        ALL4(q[c] = FM(CE(2, 2, 1), s[c], CE(1, 2, 1)))
        ALL4(q[c] = FM(q[c], s[c], CE(0, 2, 1)))
        ALL4(r1[c] = FM(r1[c], t[c], q[c]))
        ALL4(q[c] = FM(CE(3, 1, 1), s[c], CE(2, 1, 1)))
        ALL4(q[c] = FM(q[c], s[c], CE(1, 1, 1)))
        ALL4(q[c] = FM(q[c], s[c], CE(0, 1, 1)))
        ALL4(r1[c] = FM(r1[c], t[c], q[c]))
        ALL4(q[c] = FM(CE(4, 0, 1), s[c], CE(3, 0, 1)))
        ALL4(q[c] = FM(q[c], s[c], CE(2, 0, 1)))
        ALL4(q[c] = FM(q[c], s[c], CE(1, 0, 1)))
        ALL4(q[c] = FM(q[c], s[c], CE(0, 0, 1)))
        ALL4(r1[c] = FM(r1[c], t[c], q[c]))
        double r2[NB];
        ALL4(q[c] = FM(CE(1, 1, 2), s[c], CE(0, 1, 2)))
        ALL4(r2[c] = FM(CE(0, 2, 2), t[c], q[c]))
        ALL4(q[c] = FM(CE(2, 0, 2), s[c], CE(1, 0, 2)))
        ALL4(q[c] = FM(q[c], s[c], CE(0, 0, 2)))
        ALL4(r2[c] = FM(r2[c], t[c], q[c]))
        ALL4(q[c] = FM(CE(1, 0, 3), s[c], CE(0, 0, 3)))
        ALL4(q[c] = FM(CE(0, 1, 3), t[c], q[c]))
        ALL4(acc[c] += FM(FM(q[c], h, r2[c]), h, r1[c]) * h)
    }
#undef ALL4
    unsigned ok = 0u;
#pragma unroll
    for (int c = 0; c < NB; ++c) {
        const double qa = FM(acc[c], p.inv_sigstp, p.qoffset);   // (sigma - sigmin) / sigstp
        const int i = __double2int_rz(qa);
        const double fr = qa - (double)i;
        ib[c] = i;
        // strictly inside the bin range, farther than the margin from a bin edge, salinity not a mask value
        if ((unsigned)(i - 1) < (unsigned)(p.nbins - 1) && fabs(fr - 0.5) < p.half_m_margin && sal[c] != 0.0f && sal[c] != p.sps)
            ok |= 1u << c;
    }
    return ok;
}
#undef FM

// The reference chain, out of line (rare: cells near a bin edge, clamped cells, land values).
template <int EOS, bool SIGMA0>
__device__ __noinline__ int sigma_bin_slow(float tem, float sal, float sigmin, float sigstp, float sps, int nbins, double dlh,
                                           double dlref)
{
    const double dens = (EOS == CDFGPU_EOS_NEUTRAL) ? eos_sigma_neutral(tem, sal) : eos_sigma_exact<SIGMA0>(tem, sal, dlh, dlref);
    const double itm = (sal == sps) ? 0.0 : 1.0;
    const float z = __double2float_rn(__dmul_rn(dens, itm));
    const float q = __fdiv_rn(__fsub_rn(z, sigmin), sigstp);
    int ib = (q >= 2147483648.0f || q < -2147483648.0f || q != q) ? (int)0x80000000 : __float2int_rz(q);
    return min(max(ib, 1), nbins);
}

// Bins of the 8 cells of a group (scrubbed T, S; `need` = the cells whose bin is wanted): the production bin function.
// Fast path in batches of kSigEosBatch cells, the reference chain for the cells it does not vouch for.  The fused kernel
// and the diagnostic kernel behind cdfmocsig_gpu_bins_device (bit-exact sweep in the parity tests) both call this.
template <int EOS, bool SIGMA0>
__device__ __forceinline__ void sig_group_bins(const float *tt, const float *sv, unsigned need, const SigParams &p, int *ib)
{
    unsigned bad = 0xffu;
    if (EOS != CDFGPU_EOS_NEUTRAL && p.qmargin >= 0.0) {
        unsigned ok = 0u;
#pragma unroll
        for (int c0 = 0; c0 < 8; c0 += kSigEosBatch)
            ok |= sigma_bins_try<SIGMA0, kSigEosBatch>(tt + c0, sv + c0, p, ib + c0) << c0;
        bad = ~ok & 0xffu;
    }
    bad &= need;
    if (bad) {
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (bad & (1u << c))
                ib[c] = sigma_bin_slow<EOS, SIGMA0>(tt[c], sv[c], p.sigmin, p.sigstp, p.sps, p.nbins, p.dlh, p.dlref);
    }
}

// hist_flush with a shortcut for a single live lane (a bin boundary inside one lane's group)
__device__ __forceinline__ void hist_flush_few(double *hist, int key, double val, int lane)
{
    const bool live = key >= 0;
    const unsigned act = __ballot_sync(kFull, live);
    if ((act & (act - 1u)) == 0u) {   // zero or one live lane
        if (live) hist[key] += val;
        __syncwarp();
        return;
    }
    hist_flush(hist, key, val, lane);
}

// One sweep step: transports and contribution mask of this lane's group in window (k, win) of row j; queues the group
// when it contributes.  Reads V, the area and ONE byte of the pre-computed coverage plane (bit c: cell c of the group
// lies in some basin and is not an excluded column).  Returns false when the window must take the general path (a
// non-finite transport somewhere in the warp); nothing is queued then.
template <bool ISO>
__device__ __forceinline__ bool sig_sweep_window(const SigParams &p, SigQueue &q, int j, int k, int win, int lane, uint64_t pol,
                                                 int &qtail)
{
    const uint32_t r = (uint32_t)k * (uint32_t)p.ny + (uint32_t)j;
    const uint64_t e0 = (uint64_t)r * (uint32_t)p.nx;
    const uint32_t s = (uint32_t)e0 & 3u;
    const int nvec = (int)((s + (uint32_t)p.nx + 3u) >> 2);
    const int v0 = win * kSigWinVec + 2 * lane;
    const uint32_t voff = (uint32_t)(e0 >> 2) + (uint32_t)v0;
    const uint32_t pidx = (s * (uint32_t)p.ny + (uint32_t)j) * (uint32_t)p.pitchw + (uint32_t)v0;   // pattern word of the group
    const bool two = v0 + 1 < nvec;
    unsigned need = 0u;
    float zacc = 0.0f;
    float pr[8];
    if (v0 < nvec) {
        const unsigned cov = p.cov8[pidx >> 1];
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 *pv = reinterpret_cast<const float4 *>(p.zv) + voff;
        const float4 *pa = reinterpret_cast<const float4 *>(p.area) + voff;
        float4 va = ld_stream_f4(pv, pol), aa = ld_stream_f4(pa, pol);
        float4 vb = z4, ab = z4;
        if (two) { vb = ld_stream_f4(pv + 1, pol); ab = ld_stream_f4(pa + 1, pol); }
        float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w};
        const float a[8] = {aa.x, aa.y, aa.z, aa.w, ab.x, ab.y, ab.z, ab.w};
        if (p.scrub_v) {   // a zero missing value needs no scrub
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = scrub(v[c], p.spv);
        }
        if (p.zveiv) {     // -eiv: the bolus velocity is read here (rare option; its latency is not hidden)
            const float4 *pe = reinterpret_cast<const float4 *>(p.zveiv) + voff;
            const float4 ea = ld_stream_f4(pe, pol), eb = two ? ld_stream_f4(pe + 1, pol) : z4;
            const float e[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = __fadd_rn(v[c], e[c]);
        }
        unsigned nz = 0u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            pr[c] = __fmul_rn(v[c], a[c]);
            zacc = __fmaf_rn(pr[c], 0.0f, zacc);   // NaN iff some transport of the group is NaN / Inf
            if (pr[c] != 0.0f) nz |= 1u << c;
        }
        need = ISO ? cov : (cov & nz);
    }
    if (__any_sync(kFull, zacc != zacc)) return false;
    const bool has = need != 0u;
    const unsigned b = __ballot_sync(kFull, has);
    if (has) {
        const int pos = (qtail + __popc(b & ((1u << lane) - 1u))) & (kSigRing - 1);
        q.desc[pos] = make_uint4(voff, need | (two ? 0x100u : 0u) | ((uint32_t)k << 16), pidx, 0u);
        q.pa[pos] = make_float4(pr[0], pr[1], pr[2], pr[3]);
        q.pb[pos] = make_float4(pr[4], pr[5], pr[6], pr[7]);
        prefetch_l2(reinterpret_cast<const float4 *>(p.zt) + voff + 1);   // the dense step finds T / S in L2
        prefetch_l2(reinterpret_cast<const float4 *>(p.zs) + voff + 1);
    }
    qtail += __popc(b);
    return true;
}

// A popped group: descriptor, transports, and its T / S (/ area) loads in flight.
template <bool ISO>
struct SigDense {
    uint4 d;
    float4 ta, tb, sa, sb, aa, ab;
    uint32_t pw0, pw1;   // pattern bytes of the 8 cells
    int slot;            // ring slot (the transports are read from it after the EOS)
    bool active;
};

template <bool ISO>
__device__ __forceinline__ void sig_dense_pop(const SigParams &p, const SigQueue &q, int qhead, int n, int lane, uint64_t pol,
                                              SigDense<ISO> &D)
{
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    D.active = lane < n;
    D.d = make_uint4(0u, 0u, 0u, 0u);
    D.slot = 0;
    D.pw0 = 0u; D.pw1 = 0u;
    D.ta = z4; D.tb = z4; D.sa = z4; D.sb = z4; D.aa = z4; D.ab = z4;
    if (D.active) {
        const int e = (qhead + lane) & (kSigRing - 1);
        D.slot = e;
        D.d = q.desc[e];
        const bool two = (D.d.y & 0x100u) != 0u;
        D.pw0 = __ldg(p.patw + D.d.z);
        if (two) D.pw1 = __ldg(p.patw + D.d.z + 1);
        const float4 *pt = reinterpret_cast<const float4 *>(p.zt) + D.d.x;
        const float4 *ps = reinterpret_cast<const float4 *>(p.zs) + D.d.x;
        D.ta = ld_stream_f4(pt, pol);
        D.sa = ld_stream_f4(ps, pol);
        if (two) { D.tb = ld_stream_f4(pt + 1, pol); D.sb = ld_stream_f4(ps + 1, pol); }
        if (ISO) {
            const float4 *pa = reinterpret_cast<const float4 *>(p.area) + D.d.x;
            D.aa = ld_stream_f4(pa, pol);
            if (two) D.ab = ld_stream_f4(pa + 1, pol);
        }
    }
}

// EOS, bins, keys, in-lane run merge and flush of a popped group (warp collective).
template <int EOS, bool SIGMA0, bool ISO>
__device__ __forceinline__ void sig_dense_compute(const SigParams &p, const SigQueue &q, const SigDense<ISO> &D, double *hist,
                                                  int hsize, int lane)
{
    int key[8];
    float pr[8], ss[ISO ? 8 : 1], ar[ISO ? 8 : 1];
#pragma unroll
    for (int c = 0; c < 8; ++c) { key[c] = -1; pr[c] = 0.0f; }
    if (D.active) {
        unsigned need = D.d.y & 255u;
        float tt[8], sv[8];
        tt[0] = D.ta.x; tt[1] = D.ta.y; tt[2] = D.ta.z; tt[3] = D.ta.w; tt[4] = D.tb.x; tt[5] = D.tb.y; tt[6] = D.tb.z; tt[7] = D.tb.w;
        sv[0] = D.sa.x; sv[1] = D.sa.y; sv[2] = D.sa.z; sv[3] = D.sa.w; sv[4] = D.sb.x; sv[5] = D.sb.y; sv[6] = D.sb.z; sv[7] = D.sb.w;
        if (ISO) {
            ar[0] = D.aa.x; ar[1] = D.aa.y; ar[2] = D.aa.z; ar[3] = D.aa.w; ar[4] = D.ab.x; ar[5] = D.ab.y; ar[6] = D.ab.z; ar[7] = D.ab.w;
        }
        if (p.scrub_ts) {   // missing values other than zero (a zero missing value needs no scrub: x == 0 -> 0)
#pragma unroll
            for (int c = 0; c < 8; ++c) { tt[c] = scrub(tt[c], p.spt); sv[c] = scrub(sv[c], p.sps); }
        }
        if (ISO) {
            // -isodep queues every covered cell; those that are masked (itmask = 0), carry no transport and have a finite
            // area add exact zeros to all three histograms: drop them before the EOS (the dry cells below the sea floor)
            const float4 qa = q.pa[D.slot], qb = q.pb[D.slot];
            pr[0] = qa.x; pr[1] = qa.y; pr[2] = qa.z; pr[3] = qa.w; pr[4] = qb.x; pr[5] = qb.y; pr[6] = qb.z; pr[7] = qb.w;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                ss[c] = sv[c];
                const bool finite_area = (__float_as_uint(ar[c]) & 0x7f800000u) != 0x7f800000u;
                if (sv[c] == p.sps && pr[c] == 0.0f && finite_area) need &= ~(1u << c);
            }
        }
        int ib[8];
        sig_group_bins<EOS, SIGMA0>(tt, sv, need, p, ib);
        if (!ISO) {
            const float4 qa = q.pa[D.slot], qb = q.pb[D.slot];
            pr[0] = qa.x; pr[1] = qa.y; pr[2] = qa.z; pr[3] = qa.w; pr[4] = qb.x; pr[5] = qb.y; pr[6] = qb.z; pr[7] = qb.w;
        }
        // keys; a cell that does not contribute is transparent (inherits its left neighbour's key, adds zero), so that a
        // lane's last run always ends in cell 7
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int pat = (int)__byte_perm(c < 4 ? D.pw0 : D.pw1, 0u, 0x4440u + (c & 3));
            if (need & (1u << c)) key[c] = ib[c] * p.npat1 + (pat - 1 - p.npat1);
            else {
                key[c] = (c > 0) ? key[c > 0 ? c - 1 : 0] : -1;
                pr[c] = 0.0f;
                if (ISO) ar[c] = 0.0f;
            }
        }
    }
    auto pass = [&](double *h, auto value_of) {
        int kk[8];
        double val[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { kk[c] = key[c]; val[c] = value_of(c); }
#pragma unroll
        for (int c = 1; c < 8; ++c)
            if (kk[c] == kk[c - 1]) {
                val[c] += val[c - 1];
                kk[c - 1] = -1;
            }
        bool extra = false;
#pragma unroll
        for (int c = 0; c < 7; ++c) extra |= kk[c] >= 0;
        const unsigned xl = __ballot_sync(kFull, extra);
        if (xl == 0u) {   // one run per lane (smooth fields): a single warp-wide flush
            hist_flush(h, kk[7], val[7], lane);
            return;
        }
        // several lanes with several runs each: the per-lane-column table; a few bin boundaries inside single lanes stay
        // with the cheap single-lane flushes below
        if (p.use_table && __popc(xl) >= kSigTableMinLanes &&
            hist_flush_table(h, sig_tab_ptr<ISO>(p), kk, val, lane, p.npat1, p.npat1_magic))
            return;
        hist_flush(h, kk[7], val[7], lane);
#pragma unroll
        for (int c = 0; c < 7; ++c)
            if (__any_sync(kFull, kk[c] >= 0)) hist_flush_few(h, kk[c], val[c], lane);
    };
    pass(hist, [&](int c) { return 0.0 - (double)pr[c]; });
    if (ISO) {   // cdfmocsig.f90:427-428: gdep(jk)*itmask*zarea and itmask*zarea, REAL(4) chains
        const float gk = p.gdep[D.d.y >> 16];
        pass(hist + hsize, [&](int c) {
            const float itm = (ss[c] == p.sps) ? 0.0f : 1.0f;
            return (double)__fmul_rn(__fmul_rn(gk, itm), ar[c]);
        });
        pass(hist + 2 * hsize, [&](int c) {
            const float itm = (ss[c] == p.sps) ? 0.0f : 1.0f;
            return (double)__fmul_rn(itm, ar[c]);
        });
    }
}

// sig_window_general out of line, for the rare windows that hold a NaN / Inf transport
template <int EOS, bool SIGMA0, bool ISO>
__device__ __noinline__ void sig_window_general_call(const SigParams &p, SigStage &st, double *hist, unsigned *s_poison, int hsize,
                                                     int j, int k, int win, int lane)
{
    sig_window_general<EOS, SIGMA0, ISO>(p, st, hist, s_poison, hsize, j, k, win, lane, make_evict_first_policy());
}

// All windows of latitude row j that belong to this warp (w = warp, warp + nwarps, ...; window id = win*(nz-1) + k).
// One loop with a single pop / compute site: while windows remain, a dense iteration runs whenever 32 groups are
// queued; afterwards the same loop drains the ring.  Windows with a non-finite transport are only flagged here and
// handled by a second pass over the warp's windows (the order of the additions into the private histogram does not
// matter for determinism: it is fixed by the data).
template <int EOS, bool SIGMA0, bool ISO>
__device__ __forceinline__ void sig_row_queue(const SigParams &p, SigScratch &scr, double *hist, unsigned *s_poison, int hsize,
                                              int j, int warp, int nwarps, int total, int lane, uint64_t pol)
{
    const int nzm1 = p.nz - 1;
    int qhead = 0, qtail = 0;   // warp-uniform ring positions
    int win = 0, k = warp;
    bool any_weird = false;
    while (k >= nzm1) { k -= nzm1; ++win; }
    // One trip: sweep one window and queue its groups, then run a dense iteration if 32 groups are queued; after the
    // last window the same loop drains the ring (single pop / compute site).
    for (int w = warp; w < total || qtail != qhead;) {
        if (w < total) {
            if (!sig_sweep_window<ISO>(p, scr.q, j, k, win, lane, pol, qtail)) any_weird = true;
            w += nwarps;
            k += nwarps;
            while (k >= nzm1) { k -= nzm1; ++win; }
            __syncwarp();
        }
        const int cnt = qtail - qhead;
        const int n = (w < total) ? (cnt >= 32 ? 32 : 0) : min(32, cnt);
        if (n) {
            SigDense<ISO> D;
            sig_dense_pop<ISO>(p, scr.q, qhead, n, lane, pol, D);
            qhead += n;
            sig_dense_compute<EOS, SIGMA0, ISO>(p, scr.q, D, hist, hsize, lane);
        }
    }
    if (any_weird) {   // rare: NaN / Inf transports (poison semantics) -- redo exactly those windows with the general code
        __syncwarp();
        win = 0; k = warp;
        while (k >= nzm1) { k -= nzm1; ++win; }
        for (int w = warp; w < total; w += nwarps) {
            int dummy = 0;
            if (!sig_sweep_window<ISO>(p, scr.q, j, k, win, lane, pol, dummy)) {
                __syncwarp();
                sig_window_general_call<EOS, SIGMA0, ISO>(p, scr.st, hist, s_poison, hsize, j, k, win, lane);
            }
            __syncwarp();
            k += nwarps;
            while (k >= nzm1) { k -= nzm1; ++win; }
        }
    }
}

template <int EOS, bool SIGMA0, bool ISO, bool QUEUE>
__global__ void __launch_bounds__(kSigThreads, 1) mocsig_eos_hist_scan_kernel(const __grid_constant__ SigParams p)
{
    extern __shared__ double s_mem[];
    const int nwarps = blockDim.x >> 5, nthreads = blockDim.x;     // chosen by the host so that shared memory fits
    const int hsize = p.nbins * p.npat1;                            // one private histogram
    constexpr int NH = ISO ? 3 : 1;                                  // histograms per warp: transport [, depth*area, area]
    double *hist_all = s_mem;                                       // [nwarps][NH][nbins][npat1]
    double *comb = hist_all + (size_t)nwarps * NH * hsize;        // [nbins][nb]
    SigScratch *stage_all = reinterpret_cast<SigScratch *>(comb + (((size_t)p.nbins * p.nb + 1) & ~(size_t)1));  // 16-B aligned
    double *tab_all = reinterpret_cast<double *>(stage_all + nwarps);       // [nwarps][kSigTabRows*kSigTabCols] when use_table
    unsigned *s_poison = reinterpret_cast<unsigned *>(tab_all + (p.use_table ? (size_t)nwarps * kSigTabRows * kSigTabCols : 0));  // [nbins]
    __shared__ int s_ticket[2];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nzm1 = p.nz - 1;
    const uint64_t pol = make_evict_first_policy();
    double *hist = hist_all + (size_t)warp * NH * hsize;
    SigStage &st = stage_all[warp].st;
    if (p.use_table) {
        double *tab = sig_tab_ptr<ISO>(p);
        for (int t = lane; t < kSigTabRows * kSigTabCols; t += 32) tab[t] = 0.0;
        __syncwarp();
    }
    const int NV = (p.nx + 6) >> 2;                 // vectors per row, upper bound over the 4 alignments
    const int wpr = (NV + kSigWinVec - 1) / kSigWinVec;  // windows per (level) row
    const int total = nzm1 * wpr;                   // windows of one latitude row j
    // Programmatic dependent launch, as in K1 (moc_kernel.cuh): a row takes a CTA ~70 us on ORCA025, so the CTAs of a launch
    // retire over a window that long while the SMs idle; with PDL the CTAs of the next launch of the stream move in as
    // they leave.  A dependent CTA works through its first row entirely in shared memory and executes
    // griddepcontrol.wait before the row's epilogue stores; three ticket generations rotate, block 0 re-arms the next
    // one (and only then releases its own dependents) once the predecessor is complete, and no CTA retires earlier.
    int *ticket = p.tickets + p.parity;
    if (blockIdx.x == 0) {
        if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
        if (tid == 0) {
            p.tickets[(p.parity + 1) % 3] = 0;
            __threadfence();
        }
    }
    if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    bool first_row = p.pdl != 0 && blockIdx.x != 0;

    if (tid == 0) s_ticket[0] = atomicAdd(ticket, 1);
    int tsel = 0;
    for (;;) {
        for (int t = tid; t < nwarps * NH * hsize; t += nthreads) hist_all[t] = 0.0;
        for (int t = tid; t < p.nbins; t += nthreads) s_poison[t] = 0u;
        __syncthreads();
        const int j = s_ticket[tsel];
        if (j >= p.ny) {
            if (p.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
            break;
        }
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
            if (QUEUE) {
                sig_row_queue<EOS, SIGMA0, ISO>(p, stage_all[warp], hist, s_poison, hsize, j, warp, nwarps, total, lane, pol);
            } else {
                for (int w = warp; w < total; w += nwarps) {   // warp-uniform loop: one window of 64 vectors per trip
                    sig_window_general<EOS, SIGMA0, ISO>(p, st, hist, s_poison, hsize, j, k, win, lane, pol);
                    k += nwarps;
                    while (k >= nzm1) { k -= nzm1; ++win; }
                }
            }
        }
        __syncthreads();
        if (first_row) {   // PDL: the predecessor must be complete before this launch stores anything
            asm volatile("griddepcontrol.wait;" ::: "memory");
            first_row = false;
        }
        // private histograms -> one (fixed order w = 0..nwarps-1: deterministic), in place in warp 0's copy; then
        // patterns -> basins, /1e6, poison handling
        for (int t = tid; t < NH * hsize; t += nthreads) {
            double hq = 0.0;
            for (int w = 0; w < nwarps; ++w) hq += hist_all[(size_t)w * NH * hsize + t];
            hist_all[t] = hq;
        }
        __syncthreads();
        auto combine = [&](int which, int bin, int b) {
            double h = 0.0;
            for (int q = 1; q < p.npat; ++q) {
                const double wgt = c_patw[q][b];
                if (wgt != 0.0) h += hist_all[which * hsize + bin * p.npat1 + q - 1] * wgt;
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
        if (tid < p.nb) {  // integrate from the densest bin, sequentially as the reference does (:472-475); the loads
                           // of 8 bins are issued ahead of the dependent chain of additions
            double psi = comb[(p.nbins - 1) * p.nb + tid];
            int bin = p.nbins - 2;
            for (; bin >= 7; bin -= 8) {
                double h[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) h[u] = comb[(bin - u) * p.nb + tid];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    psi = psi + h[u];
                    comb[(bin - u) * p.nb + tid] = psi;
                }
            }
            for (; bin >= 0; --bin) {
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

// Diagnostic kernel for the parity tests: ibin of every cell of a record, through the same group function as the fused
// kernel (8 consecutive cells per thread).
template <int EOS, bool SIGMA0>
__global__ void mocsig_bins_kernel(const __grid_constant__ SigParams p, int32_t *__restrict__ ibin, size_t n)
{
    const size_t ngroups = (n + 7) / 8;
    for (size_t g8 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g8 < ngroups; g8 += (size_t)gridDim.x * blockDim.x) {
        float tt[8], sv[8];
        int ib[8];
        unsigned need = 0u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const size_t e = g8 * 8 + c;
            const bool in = e < n;
            tt[c] = scrub(in ? p.zt[e] : 0.0f, p.spt);
            sv[c] = scrub(in ? p.zs[e] : 0.0f, p.sps);
            if (in) need |= 1u << c;
        }
        sig_group_bins<EOS, SIGMA0>(tt, sv, need, p, ib);
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (need & (1u << c)) ibin[g8 * 8 + c] = ib[c];
    }
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
