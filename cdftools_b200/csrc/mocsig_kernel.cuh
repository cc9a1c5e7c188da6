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
// Second generation (round 2).  The first generation (git history: mocsig_kernel.cuh of round 1) spent ~117 instructions
// per cell, ~33 of them on the fp64 EOS and the rest on a sweep / queue / pop machinery that compacted wet groups for the
// fp64 pipe.  This one removes the reason for the machinery instead:
//
//   * THREE-TIER BIN FUNCTION.  Tier 1 evaluates the density polynomial in fp32 with packed FFMA2 (two cells per
//     instruction): for a fixed reference depth the 52-term (s,t,h) polynomial collapses to 28 (s,t) terms, which the host
//     re-centres on the (T,S) domain box so that every term is O(10) instead of O(10^3) (mocsig_filter.hpp) -- the fp32
//     result is then within ~2e-5 kg/m3 of the reference's fp64 value, with a rigorous bound derived on the host.  A cell's
//     bin is accepted when q = (sigma - sigmin)/sigstp is farther from an integer than that bound; the ~0.2 % that are not
//     go to tier 2 (the same 28-term polynomial in fp64 FMA, margin from the fp32 roundings of the reference's own bin
//     formula), and what tier 2 cannot vouch for (~1e-4 of those, plus land / clamped / out-of-domain / NaN cells) takes
//     tier 3, the reference chain operation for operation.  Bins stay bit-exact; sigma2 costs the same as sigma0.
//   * NO QUEUE.  A warp owns a 256-cell window (8 consecutive cells per lane = one 32-byte sector per field and lane) and
//     runs a three-stage register pipeline over its windows: V / area loads of window n+2 and T / S loads of window n+1
//     are in flight while window n is evaluated (T / S sectors of lanes without a contributing cell are never requested).
//   * LAZY LANE-COLUMN TABLE.  A lane merges equal (bin, pattern) neighbours in registers and adds what is left into ITS
//     OWN column of a small per-warp table tab[bin - lo][lane] (no two lanes share a word: no atomics, no collectives, fixed
//     order of additions = bitwise reproducible).  The table is folded into the warp's private histogram only when a cell
//     falls outside its band of bins / its mask pattern, or at the end of the work unit.
//   * WORK UNIT = (latitude row j, chunk of levels), handed out by ticket; the host picks the chunking so that a launch
//     has >= ~8 units per CTA (a 383-row ORCA12 band no longer quantises to 2.6 waves).  With more than one chunk per row
//     the units write partial histograms and the CTA that completes a row folds them in chunk order (deterministic).
#pragma once
#include "common.cuh"
#include "eos_device.cuh"
#include "mocsig_filter.hpp"

namespace cdfgpu {

#ifndef CDF_SIG_WARPS
#define CDF_SIG_WARPS 16   // most warps of a CTA (the host launches 16 x 1 CTA or 8 x 2 CTAs per SM: 128 registers either way)
#endif
#ifndef CDF_SIG_MINCTAS
#define CDF_SIG_MINCTAS 1
#endif
#ifndef CDF_SIG_TAB_ROWS
#define CDF_SIG_TAB_ROWS 8
#endif
#ifndef CDF_SIG_TAB_SLACK
#define CDF_SIG_TAB_SLACK 2
#endif
constexpr int kSigWarps = CDF_SIG_WARPS;
constexpr int kSigThreads = kSigWarps * 32;
constexpr int kSigMinCtas = CDF_SIG_MINCTAS;
constexpr int kSigMaxPat = 32;
constexpr int kSigWinVec = 64;                  // one warp step = 64 float4 vectors = 256 cells, 8 consecutive cells per lane
constexpr int kSigTabRows = CDF_SIG_TAB_ROWS;   // bins covered by the lane-column table
constexpr int kSigTabSlack = CDF_SIG_TAB_SLACK; // a new band starts this many bins below the lowest pending bin
#ifndef CDF_SIG_PREFETCH_VA
#define CDF_SIG_PREFETCH_VA 1   // V / area of window n+2 in flight during window n (0: only T / S of window n+1)
#endif
#ifndef CDF_SIG_SEG_WIN
#define CDF_SIG_SEG_WIN 16
#endif
constexpr int kSigSegWin = CDF_SIG_SEG_WIN;     // windows per run grabbed by a warp, at most
static_assert(kSigTabRows == 8 || kSigTabRows == 16, "table reduce is written for 8 or 16 rows");

__constant__ double c_patw[kSigMaxPat][CDFGPU_MAX_BASINS];  // weight of pattern p for basin b (= mask value)

struct SigParams {
    const float *__restrict__ zv, *__restrict__ zt, *__restrict__ zs;  // (nz-1, ny, nx) raw file values
    const float *__restrict__ zveiv;                                   // optional, may be null
    const float *__restrict__ area;                                    // (nz-1, ny, nx) fl32(e1v*e3v), zero outside the basins
    const uint32_t *__restrict__ patw;                                 // [4][ny][pitchw] pattern bytes, shifted copies
    double *__restrict__ out;                                          // (ny, nbins, nb)
    double *__restrict__ out_iso;                                      // -isodep: (ny, nbins, nb) mean isopycnal depth
    const float *__restrict__ gdep;                                    // -isodep: -gdept(k), (nz)
    int *tickets;                                                      // [3] unit tickets, one per launch generation
    int *col;                                                          // [ny] chunks of a row that have been folded (nchunk > 1)
    double *part;                                                      // [ny*nchunk][NH*hsize] partial histograms (nchunk > 1)
    unsigned *partp;                                                   // [ny*nchunk][nbins]    partial poison bits  (nchunk > 1)
    int pdl;          // launched with programmatic stream serialization
    int pdl_early;    // ... and the inputs were complete before the predecessor was launched: loads may precede the wait
    int nx, ny, nz, nb, nbins, npat, npat1, pitchw;  // npat incl. the all-zero pattern; npat1 = max(npat-1,1)
    int nchunk, nkc, nunits;                         // chunks per row, levels per chunk, ny * nchunk
    int parity;
    int j_first_global, ny_global;
    int sigma0;       // pref == 0: dlr == dlr0 in the reference chain
    float sigmin, sigstp;
    float spv, spt, sps;
    double dlh, dlref;
    int scrub_ts;     // T or S missing value is not zero
    int scrub_v;      // V missing value is not zero
    SigFilter f;      // tiers 1 and 2 of the bin function (mocsig_filter.hpp)
};

// ---- packed fp32 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2: two cells per instruction) ----------------------------
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 dup2(float x) { return pk2(x, x); }
__device__ __forceinline__ u64 ld2(const float2 &v) { return *reinterpret_cast<const u64 *>(&v); }   // a duplicated constant
__device__ __forceinline__ float rsqrt_approx(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// ---- tier 1: fp32 bins of the 8 cells of a group ------------------------------------------------------------------
// Operation for operation the sequence whose error mocsig_filter.hpp bounds (sig_filter_build): any change here must be
// mirrored there.  ib[] holds trunc(q); returns the mask of the cells that are wanted (wanted(c)) and whose bin is NOT
// provably the reference's.
template <class Wanted>
__device__ __forceinline__ unsigned sigma_bins_f32x8(const float (&tt)[8], const float (&sv)[8], const SigParams &p, int (&ib)[8],
                                                     Wanted wanted)
{
    const SigFilter &f = p.f;
    u64 u[4], v[4], acc[4], q[4];
    {
        const u64 ta = ld2(f.ta), tb = ld2(f.tb), sr = ld2(f.sr), sdr = ld2(f.sdr), nsr = ld2(f.nsr), nsdr = ld2(f.nsdr);
        const u64 ns0 = ld2(f.ns0), ihs = ld2(f.ihs), mhalf = dup2(-0.5f);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const u64 T2 = pk2(tt[2 * g], tt[2 * g + 1]), S2 = pk2(sv[2 * g], sv[2 * g + 1]);
            v[g] = fma2(T2, ta, tb);                 // v = (T/40 - t0) / ht
            const u64 x = fma2(S2, sr, sdr);         // x = (S + deltaS) * r1_S0
            const u64 nx = fma2(S2, nsr, nsdr);      // -x, exactly
            float x0, x1;
            upk2(x, x0, x1);
            const u64 y = pk2(rsqrt_approx(x0), rsqrt_approx(x1));
            u64 sx = mul2(x, y);                     // sqrt(x) to 2^-22
            const u64 e = fma2(sx, sx, nx);          // sx^2 - x
            sx = fma2(e, mul2(y, mhalf), sx);        // one Newton step: sqrt(x) to ~1 ulp
            u[g] = mul2(add2(sx, ns0), ihs);         // u = (sqrt(x) - s0) / hs
        }
    }
    {   // P(u,v) - c00 = sum_j v^j Q_j(u), Horner in u inside Horner in v, coefficient-major over the four cell pairs
        const float2 *c = f.c32;
#define ALL4(expr) _Pragma("unroll") for (int g = 0; g < 4; ++g) { expr; }
        {
            const u64 c0 = ld2(c[0]);
            ALL4(acc[g] = c0)
        }
        int idx = 1;
#pragma unroll
        for (int j = 5; j >= 0; --j) {
            {
                const u64 ck = ld2(c[idx]);
                ALL4(q[g] = ck)
            }
            ++idx;
#pragma unroll
            for (int i = 5 - j; i >= 0; --i) {
                const u64 ck = ld2(c[idx]);
                ALL4(q[g] = fma2(q[g], u[g], ck))
                ++idx;
            }
            ALL4(acc[g] = fma2(acc[g], v[g], q[g]))
        }
#undef ALL4
    }
    unsigned bad = 0u;
    const u64 qs = ld2(f.qscale), qo = ld2(f.qoff), magic = dup2(12582912.0f), nmagic = dup2(-12582912.0f);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        const u64 qq = fma2(acc[g], qs, qo);        // q = (sigma - sigmin) / sigstp
        const u64 r = add2(qq, magic);              // low mantissa bits = q rounded to the nearest integer (0 <= q < 2^22)
        const u64 d = sub2(qq, add2(r, nmagic));    // q - rint(q), exact
        float qe[2], re[2], de[2], ue[2], ve[2];
        upk2(qq, qe[0], qe[1]); upk2(r, re[0], re[1]); upk2(d, de[0], de[1]); upk2(u[g], ue[0], ue[1]); upk2(v[g], ve[0], ve[1]);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            // trunc(q) = rint(q) - (q < rint(q))
            ib[2 * g + e] = (__float_as_int(re[e]) - 0x4B400000) + (__float_as_int(de[e]) >> 31);
            // farther than the error bound from a bin edge, strictly inside the bin range, (T,S) inside the domain box the
            // bound was derived on (NaN fails every comparison)
            const bool good = (fabsf(de[e]) > f.margin32) && (fabsf(__fsub_rn(qe[e], f.qc)) < f.qh) &&
                              (fmaxf(fabsf(ue[e]), fabsf(ve[e])) <= 1.0f);
            bad |= good ? 0u : 1u << (2 * g + e);
        }
    }
    unsigned wm = 0u;   // (mask arithmetic instead of a branch per cell)
#pragma unroll
    for (int c = 0; c < 8; ++c) wm |= wanted(c) ? 1u << c : 0u;
    return bad & wm;
}

// ---- tier 2: the same 28-term polynomial in fp64 FMA, one cell ------------------------------------------------------
// Differs from the reference value by rounding only (~1e-12 kg/m3; the margin allows 1e-9).  True: *ibout is the bin.
__device__ __forceinline__ bool sigma_bin_f64(float tem, float sal, const SigParams &p, int *ibout)
{
    const SigFilter &f = p.f;
    const double t = (double)tem * (1.0 / 40.0);
    const double s = sqrt(fabs((double)sal + f.rdeltaS) * f.r1_S0);
    double acc = f.c64[0];
    int idx = 1;
#pragma unroll
    for (int j = 5; j >= 0; --j) {
        double q = f.c64[idx++];
#pragma unroll
        for (int i = 5 - j; i >= 0; --i) q = fma(q, s, f.c64[idx++]);
        acc = fma(acc, t, q);
    }
    const double qa = fma(acc, f.inv_sigstp, f.qoffset);   // (sigma - sigmin) / sigstp
    const int i = __double2int_rz(qa);
    const double fr = qa - (double)i;
    *ibout = i;
    // strictly inside the bin range, farther than the margin from a bin edge, salinity not a mask value
    return (unsigned)(i - 1) < (unsigned)(p.nbins - 1) && fabs(fr - 0.5) < f.half_m_margin && sal != 0.0f && sal != p.sps;
}

// ---- tier 3: the reference chain, out of line (cells near a bin edge, clamped cells, land values, NaN) --------------
template <bool NEUTRAL>
__device__ __noinline__ int sigma_bin_slow(float tem, float sal, float sigmin, float sigstp, float sps, int nbins, double dlh,
                                           double dlref, int sigma0)
{
    const double dens = NEUTRAL  ? eos_sigma_neutral(tem, sal)
                        : sigma0 ? eos_sigma_exact<true>(tem, sal, dlh, dlref)
                                 : eos_sigma_exact<false>(tem, sal, dlh, dlref);
    const double itm = (sal == sps) ? 0.0 : 1.0;  // itmask, evaluated on the scrubbed salinity (cdfmocsig.f90:393-394)
    const float z = __double2float_rn(__dmul_rn(dens, itm));
    const float q = __fdiv_rn(__fsub_rn(z, sigmin), sigstp);
    // INT(): truncation toward zero; out-of-range and NaN give INT_MIN like x86 CVTTSS2SI (what gfortran emits)
    int ib = (q >= 2147483648.0f || q < -2147483648.0f || q != q) ? (int)0x80000000 : __float2int_rz(q);
    return min(max(ib, 1), nbins);
}

// Bins (1..nbins) of the wanted cells of a group from scrubbed T, S: the production bin function.  The fused kernel and
// the diagnostic kernel behind cdfmocsig_gpu_bins_device (bit-exact sweeps in the parity tests) both call this.
// wanted(c): does cell c need a bin.  stats (diagnostic kernel only): [0] cells asked for, [1] cells tier 1 did not accept,
// [2] cells that took tier 3.
template <bool NEUTRAL, class Wanted>
__device__ __forceinline__ void sig_group_bins(const float (&tt)[8], const float (&sv)[8], const SigParams &p, int (&ib)[8], Wanted wanted,
                                               unsigned long long *stats = nullptr)
{
    unsigned bad = 0u;
    if (!NEUTRAL && p.f.tier1) {
        bad = sigma_bins_f32x8(tt, sv, p, ib, wanted);
    } else {
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (wanted(c)) bad |= 1u << c;
    }
    if (stats) {
        unsigned n = 0u;
#pragma unroll
        for (int c = 0; c < 8; ++c) n += wanted(c) ? 1u : 0u;
        atomicAdd(stats + 0, (unsigned long long)n);
        atomicAdd(stats + 1, (unsigned long long)__popc(bad));
    }
    while (bad) {   // per lane: one rejected cell per trip
        const int c = __ffs(bad) - 1;
        bad &= bad - 1u;
        float t = tt[0], s = sv[0];
#pragma unroll
        for (int e = 1; e < 8; ++e)
            if (c == e) { t = tt[e]; s = sv[e]; }
        int b;
        if (NEUTRAL || !p.f.tier2 || !sigma_bin_f64(t, s, p, &b)) {
            b = sigma_bin_slow<NEUTRAL>(t, s, p.sigmin, p.sigstp, p.sps, p.nbins, p.dlh, p.dlref, p.sigma0);
            if (stats) atomicAdd(stats + 2, 1ull);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (c == e) ib[e] = b;
    }
}

__device__ __forceinline__ float scrub(float x, float spval) { return (x == spval) ? 0.0f : x; }

// ---- lane-column table ------------------------------------------------------------------------------------------------
// A histogram KEY is (pattern << 16) + bin (1-based mask pattern and bin; bins < 65536), -1 = none.  The table tab[r][lane] covers the
// kSigTabRows keys base .. base + kSigTabRows - 1 (one mask pattern, a band of bins); it is all zero whenever !dirty.
struct SigTabState {   // warp-uniform
    int base;          // first key of the band; -1: none yet
    bool dirty;
};

__device__ __forceinline__ void sig_tab_flush(double *hist, double *tab, SigTabState &st, int npat1, int nbins, int lane)
{
    if (!st.dirty || st.base < 0) { st.dirty = false; return; }   // warp-uniform
    __syncwarp();
    constexpr int LPR = 32 / kSigTabRows;   // lanes per row:     8 rows -> 4, 16 rows -> 2
    constexpr int CPL = 32 / LPR;           // columns per lane:  8 rows -> 8, 16 rows -> 16
    const int r = lane / LPR, part = lane % LPR;
    double *row = tab + r * 32 + part * CPL;
    // skewed start so that the 16 lanes of a half-warp (64-bit accesses go half-warp by half-warp) hit 16 distinct bank pairs
    const int skew = (kSigTabRows == 8) ? (r + 4 * (part >> 1)) : (r + 8 * part);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
        const int e = (i + skew) & (CPL - 1);
        s += row[e];
        row[e] = 0.0;
    }
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) s += __shfl_xor_sync(kFull, s, o);
    const int bin = (st.base & 0xffff) - 1 + r;   // keys hold 1-based bins and patterns
    if (part == 0 && bin < nbins) hist[bin * npat1 + (st.base >> 16) - 1] += s;
    __syncwarp();
    st.dirty = false;
}

// Entries of one lane: kk[c] = key or -1, val[c] the fp64 value; `extra` (warp-uniform): some lane holds an entry in c < 7.
// Adds them to the warp's table, re-basing the band (fold into `hist`, new base) as often as needed.  Warp collective.
__device__ __forceinline__ void sig_accumulate(double *hist, double *tab, SigTabState &st, int (&kk)[8], const double (&val)[8],
                                               bool extra, int npat1, int nbins, int lane)
{
    double *col = tab + lane;
    bool miss = false;   // an entry that does not fit the band
    if (extra) {
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const unsigned r = (unsigned)(kk[c] - st.base);
            if (r < (unsigned)kSigTabRows) {
                col[r * 32] += val[c];
                kk[c] = -1;
            }
            miss = miss || kk[c] >= 0;
        }
    }
    {
        const unsigned r = (unsigned)(kk[7] - st.base);
        if (r < (unsigned)kSigTabRows) {
            col[r * 32] += val[7];
            kk[7] = -1;
        }
        miss = miss || kk[7] >= 0;
    }
    st.dirty = true;   // (a flush of an untouched table adds zeros)
    if (!__any_sync(kFull, miss)) return;
#pragma unroll 1
    for (;;) {
        // the band moves to the lowest key still pending; entries that fit go in
        int m = 0x7fffffff;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (kk[c] >= 0) m = min(m, kk[c]);
        m = __reduce_min_sync(kFull, m);
        if (m == 0x7fffffff) break;   // warp-uniform
        sig_tab_flush(hist, tab, st, npat1, nbins, lane);
        st.base = max(m - kSigTabSlack, (m & ~0xffff) + 1);   // a few bins of slack below, never below bin 1 of the pattern
        st.dirty = true;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const unsigned r = (unsigned)(kk[c] - st.base);
            if (r < (unsigned)kSigTabRows) {
                col[r * 32] += val[c];
                kk[c] = -1;
            }
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

// ---- the three pipeline stages of a window ------------------------------------------------------------------------------
// Register budget: 16 warps / SM leave 128 registers per thread, and the pipeline keeps V / area of window n+2 and T / S of
// windows n+1 and n in registers (48); everything else a later stage needs is packed (flags) or re-derived: the fp32
// transports wait in shared memory (32 B per thread and pipeline slot), the pattern words are re-read (L1 hits).
constexpr unsigned kSigTwo = 0x100u, kSigValid = 0x200u, kSigPoison = 0x400u, kSigWanted = 0x800u, kSigEdge = 0x1000u;   // flag bits
// bit c (0..3) of the result: byte c of a pattern word is a real pattern (neither 0 = no basin nor 255 = excluded column)
__device__ __forceinline__ unsigned sig_covered(uint32_t pw)
{
    unsigned m = 0u;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const unsigned b = (pw >> (8 * c)) & 255u;
        if (b != 0u && b != 255u) m |= 1u << c;
    }
    return m;
}
struct SigS1 {          // V / area loads in flight
    float4 va, vb, aa, ab;
    uint32_t voff;      // float4 index of the lane's first vector in the flat (nz-1, ny, nx) arrays
    uint32_t pidx;      // word index of the lane's first pattern word
    unsigned flags;     // kSigValid / kSigTwo: the lane's first / second vector overlaps the row; level << 16
    bool live;          // warp-uniform: the slot holds a window
};
struct SigS2 {          // transports known (shared memory), T / S loads in flight
    float4 ta, tb, sa, sb;
    uint32_t voff, pidx;
    unsigned need;      // kSigTwo, kSigValid; kSigWanted: some cell of the lane contributes (T / S requested); kSigPoison
                        // (warp-uniform): some transport of the window is NaN / Inf -> general path; bits 16..31: level
    bool live;
};

// The stream of windows of one warp inside a work unit.  The unit's windows are numbered level by level (w = kk * wpr +
// win); a warp grabs RUNS of consecutive windows from a shared-memory ticket, so its consecutive windows are neighbours at
// one depth -- their cells fall into the same few density classes and the table's band holds -- and the runs shrink
// towards the end of the unit (guided self-scheduling: remaining / (2 * warps), between 1 and kSigSegWin windows), so the
// warps of a CTA reach the unit's barrier together whatever the land / ocean geometry.
struct SigStream {
    uint32_t voff0, pidx0;   // of window 0 of the current level row, lane 0
    int nvec;                // vectors of the level row
    int w, w_end;            // next window, end of the run (unit-wide numbering)
    int win;                 // next window within its level row
    int k;
};

__device__ __forceinline__ void sig_row_geometry(const SigParams &p, SigStream &s, int j)
{
    const uint32_t r = (uint32_t)s.k * (uint32_t)p.ny + (uint32_t)j;
    const uint64_t e0 = (uint64_t)r * (uint32_t)p.nx;
    const uint32_t sh = (uint32_t)e0 & 3u;
    s.nvec = (int)((sh + (uint32_t)p.nx + 3u) >> 2);
    s.voff0 = (uint32_t)(e0 >> 2);
    s.pidx0 = (sh * (uint32_t)p.ny + (uint32_t)j) * (uint32_t)p.pitchw;
}

__device__ __forceinline__ void sig_next_window(const SigParams &p, SigStream &s, SigS1 &a, int *s_seg, int j, int k0, int total,
                                                int nwarps, int wpr, int lane, uint64_t pol)
{
    a.live = true;
    if (s.w >= s.w_end) {   // warp-uniform: grab the next run of windows
        int start = 0, n = 0;
        if (lane == 0) {
            const int rem = total - *reinterpret_cast<volatile int *>(s_seg);
            n = max(1, min(kSigSegWin, rem / (2 * nwarps)));
            start = atomicAdd(s_seg, n);
        }
        start = __shfl_sync(kFull, start, 0);
        n = __shfl_sync(kFull, n, 0);
        if (start >= total) {
            a.live = false;
            s.w = 0; s.w_end = 0;   // stays exhausted
        } else {
            const int kk = start / wpr;
            s.w = start;
            s.w_end = min(total, start + n);
            s.win = start - kk * wpr;
            s.k = k0 + kk;
            sig_row_geometry(p, s, j);
        }
    } else if (s.win == wpr) {   // the run continues on the next level row
        s.win = 0;
        ++s.k;
        sig_row_geometry(p, s, j);
    }
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    a.va = z4; a.vb = z4; a.aa = z4; a.ab = z4;
    a.flags = 0u; a.voff = 0u; a.pidx = 0u;
    if (a.live) {
        const int v0 = s.win * kSigWinVec + 2 * lane;
        a.voff = s.voff0 + (uint32_t)v0;
        a.pidx = s.pidx0 + (uint32_t)v0;
        a.flags = (unsigned)s.k << 16;
        if (v0 < s.nvec) a.flags |= kSigValid;
        if (v0 + 1 < s.nvec) a.flags |= kSigTwo;
        if (v0 == 0 || v0 + 2 >= s.nvec) a.flags |= kSigEdge;   // the row's first / last vector: may hold cells of the neighbour rows
        ++s.win;
        ++s.w;
        if (a.flags & kSigValid) {
            const float4 *pv = reinterpret_cast<const float4 *>(p.zv) + a.voff;
            const float4 *pa = reinterpret_cast<const float4 *>(p.area) + a.voff;
            a.va = ld_stream_f4(pv, pol);
            a.aa = ld_stream_f4(pa, pol);
            if (a.flags & kSigTwo) { a.vb = ld_stream_f4(pv + 1, pol); a.ab = ld_stream_f4(pa + 1, pol); }
        }
    }
}

// transports + contribution mask from V / area (waits for them), T / S loads of the lanes that need them
template <bool ISO>
__device__ __forceinline__ void sig_stage2(const SigParams &p, const SigS1 &a, SigS2 &d, float4 *prs, int lane, uint64_t pol)
{
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    d.live = a.live;
    d.voff = a.voff; d.pidx = a.pidx;
    d.need = a.flags;
    d.ta = z4; d.tb = z4; d.sa = z4; d.sb = z4;
    float zacc = 0.0f;
    if (a.flags & kSigValid) {
        // The resident area field is ZERO outside every basin and in the excluded columns (sig_prep_area_kernel): those
        // cells, like exact zeros of V, contribute nothing in the reference, and their transports come out as +-0 here.
        float pr[8];
        if (p.scrub_v || p.zveiv || (a.flags & kSigEdge)) {   // (uniform or rare)
            float v[8] = {a.va.x, a.va.y, a.va.z, a.va.w, a.vb.x, a.vb.y, a.vb.z, a.vb.w};
            const float ar[8] = {a.aa.x, a.aa.y, a.aa.z, a.aa.w, a.ab.x, a.ab.y, a.ab.z, a.ab.w};
            if (p.scrub_v) {   // a zero missing value needs no scrub
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = scrub(v[c], p.spv);
            }
            if (p.zveiv) {     // -eiv: the bolus velocity is read here (rare option; its latency is not hidden)
                const float4 *pe = reinterpret_cast<const float4 *>(p.zveiv) + a.voff;
                const float4 ea = ld_stream_f4(pe, pol), eb = (a.flags & kSigTwo) ? ld_stream_f4(pe + 1, pol) : z4;
                const float e[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
#pragma unroll
                for (int c = 0; c < 8; ++c) v[c] = __fadd_rn(v[c], e[c]);
            }
            if (a.flags & kSigEdge) {   // 16-byte vectors on the flat array straddle the row ends: cells of the neighbour rows out
                const uint32_t pw0 = __ldg(p.patw + a.pidx), pw1 = (a.flags & kSigTwo) ? __ldg(p.patw + a.pidx + 1) : 0xffffffffu;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if ((((c < 4 ? pw0 : pw1) >> (8 * (c & 3))) & 255u) == 255u) v[c] = 0.0f;
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) pr[c] = __fmul_rn(v[c], ar[c]);
        } else {
            pr[0] = __fmul_rn(a.va.x, a.aa.x); pr[1] = __fmul_rn(a.va.y, a.aa.y); pr[2] = __fmul_rn(a.va.z, a.aa.z); pr[3] = __fmul_rn(a.va.w, a.aa.w);
            pr[4] = __fmul_rn(a.vb.x, a.ab.x); pr[5] = __fmul_rn(a.vb.y, a.ab.y); pr[6] = __fmul_rn(a.vb.z, a.ab.z); pr[7] = __fmul_rn(a.vb.w, a.ab.w);
        }
        unsigned any = 0u;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            zacc = __fmaf_rn(pr[c], 0.0f, zacc);   // NaN iff some transport of the group is NaN / Inf
            any |= __float_as_uint(pr[c]);
        }
        bool wanted = (any << 1) != 0u;            // some transport is not +-0
        if (ISO) {   // -isodep: every cell inside a basin carries area weight into its class, whatever its transport
            const uint32_t pw0 = __ldg(p.patw + a.pidx), pw1 = (a.flags & kSigTwo) ? __ldg(p.patw + a.pidx + 1) : 0xffffffffu;
            wanted = wanted || sig_covered(pw0) != 0u || sig_covered(pw1) != 0u;
        }
        if (wanted) {
            d.need |= kSigWanted;
            prs[0] = make_float4(pr[0], pr[1], pr[2], pr[3]);
            prs[blockDim.x] = make_float4(pr[4], pr[5], pr[6], pr[7]);
        }
    }
    if (__any_sync(kFull, zacc != zacc)) d.need |= kSigPoison;
    if ((d.need & kSigWanted) != 0u && !(d.need & kSigPoison)) {
        const float4 *pt = reinterpret_cast<const float4 *>(p.zt) + d.voff;
        const float4 *ps = reinterpret_cast<const float4 *>(p.zs) + d.voff;
        d.ta = ld_stream_f4(pt, pol);
        d.sa = ld_stream_f4(ps, pol);
        if (d.need & kSigTwo) { d.tb = ld_stream_f4(pt + 1, pol); d.sb = ld_stream_f4(ps + 1, pol); }
    }
}

// One window, fully general and slow: any mix of covered / uncovered cells, NaN / Inf transports (the reference's poison
// semantics: Inf*0 and NaN*x turn a basin's class sum into NaN), -isodep.  Every bin through the reference chain; the lanes
// add their cells to the warp's private histogram one lane at a time (fixed order, no conflicts).
template <bool NEUTRAL, bool ISO>
__device__ __noinline__ void sig_window_general(const SigParams &p, double *hist, unsigned *s_poison, int hsize, uint32_t voff,
                                                uint32_t pidx, bool valid, bool two, int k, int lane)
{
    const uint64_t pol = make_evict_first_policy();
    int key[8];
    double val[8], v1[ISO ? 8 : 1], v2[ISO ? 8 : 1];
#pragma unroll
    for (int c = 0; c < 8; ++c) { key[c] = -1; val[c] = 0.0; }
    if (valid) {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t pw0 = __ldg(p.patw + pidx), pw1 = two ? __ldg(p.patw + pidx + 1) : 0xffffffffu;
        const float4 *pv = reinterpret_cast<const float4 *>(p.zv) + voff;
        const float4 *pa = reinterpret_cast<const float4 *>(p.area) + voff;
        const float4 *pt = reinterpret_cast<const float4 *>(p.zt) + voff;
        const float4 *ps = reinterpret_cast<const float4 *>(p.zs) + voff;
        const float4 va = ld_stream_f4(pv, pol), vb = two ? ld_stream_f4(pv + 1, pol) : z4;
        const float4 aa = ld_stream_f4(pa, pol), ab = two ? ld_stream_f4(pa + 1, pol) : z4;
        const float4 ta = ld_stream_f4(pt, pol), tb = two ? ld_stream_f4(pt + 1, pol) : z4;
        const float4 sa = ld_stream_f4(ps, pol), sb = two ? ld_stream_f4(ps + 1, pol) : z4;
        float4 ea = z4, eb = z4;
        if (p.zveiv) {
            const float4 *pe = reinterpret_cast<const float4 *>(p.zveiv) + voff;
            ea = ld_stream_f4(pe, pol);
            if (two) eb = ld_stream_f4(pe + 1, pol);
        }
        const float v[8] = {va.x, va.y, va.z, va.w, vb.x, vb.y, vb.z, vb.w}, e[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
        const float ar[8] = {aa.x, aa.y, aa.z, aa.w, ab.x, ab.y, ab.z, ab.w};
        const float tt[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w}, ss[8] = {sa.x, sa.y, sa.z, sa.w, sb.x, sb.y, sb.z, sb.w};
        const float gk = ISO ? p.gdep[k] : 0.0f;
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
            const uint32_t pat = ((c < 4 ? pw0 : pw1) >> (8 * (c & 3))) & 255u;
            const float pr = sig_transport(p, v[c], e[c], ar[c]);
            const bool finite = (__float_as_uint(pr) & 0x7f800000u) != 0x7f800000u;
            // excluded cell, exact zero, or finite transport not covered by any basin: contributes nothing.
            const bool need = ISO ? (pat != 255u && (pat != 0u || (!finite && pr != 0.0f)))
                                  : (pat != 255u && pr != 0.0f && (pat != 0u || !finite));
            if (!need) continue;
            const float sc = scrub(ss[c], p.sps);
            const int ib = sigma_bin_slow<NEUTRAL>(scrub(tt[c], p.spt), sc, p.sigmin, p.sigstp, p.sps, p.nbins, p.dlh, p.dlref, p.sigma0);
            bool add = true;
            if (!finite) {
                // NaN/Inf transport: basin b is poisoned iff (0-p)*mask_b is NaN (NaN*x, or Inf*0)
                unsigned bits = 0u;
                for (int b = 0; b < p.nb; ++b) {
                    const double cc = __dmul_rn(0.0 - (double)pr, c_patw[pat][b]);
                    if (cc != cc) bits |= 1u << b;
                }
                if (bits) atomicOr(s_poison + (ib - 1), bits);
                add = !(pr != pr || pat == 0u);   // an Inf with a covering basin still accumulates
            }
            if (add || (ISO && pat != 0u)) key[c] = (ib - 1) * p.npat1 + (int)pat - 1;
            val[c] = add ? 0.0 - (double)pr : 0.0;
            if (ISO) {   // cdfmocsig.f90:427-428: gdep(jk)*itmask*zarea and itmask*zarea, REAL(4) chains
                const float itm = (sc == p.sps) ? 0.0f : 1.0f;
                v1[c] = (double)__fmul_rn(__fmul_rn(gk, itm), ar[c]);
                v2[c] = (double)__fmul_rn(itm, ar[c]);
            }
        }
    }
    __syncwarp();
    for (int src = 0; src < 32; ++src) {
        if (lane == src) {
#pragma unroll
            for (int c = 0; c < 8; ++c)
                if (key[c] >= 0) {
                    hist[key[c]] += val[c];
                    if (ISO) { hist[hsize + key[c]] += v1[c]; hist[2 * hsize + key[c]] += v2[c]; }
                }
        }
        __syncwarp();
    }
}

// EOS, bins, keys, run merge and table accumulation of a window whose T / S have arrived (warp collective).
template <bool NEUTRAL, bool ISO>
__device__ __forceinline__ void sig_stage3(const SigParams &p, const SigS2 &d, const float4 *prs, double *hist, double *tab,
                                           SigTabState &st, unsigned *s_poison, int hsize, int lane, uint64_t pol)
{
    const int k = (int)(d.need >> 16);
    if (d.need & kSigPoison) {   // warp-uniform, rare: the general code adds straight into the private histogram
        sig_tab_flush(hist, tab, st, p.npat1, p.nbins, lane);
        sig_window_general<NEUTRAL, ISO>(p, hist, s_poison, hsize, d.voff, d.pidx, (d.need & kSigValid) != 0u, (d.need & kSigTwo) != 0u,
                                         k, lane);
        return;
    }
    const bool wanted = (d.need & kSigWanted) != 0u;
    if (!__any_sync(kFull, wanted)) return;   // nothing of this window contributes
    float tt[8] = {d.ta.x, d.ta.y, d.ta.z, d.ta.w, d.tb.x, d.tb.y, d.tb.z, d.tb.w};
    float sv[8] = {d.sa.x, d.sa.y, d.sa.z, d.sa.w, d.sb.x, d.sb.y, d.sb.z, d.sb.w};
    float ar[ISO ? 8 : 1];
    if (p.scrub_ts) {   // missing values other than zero (a zero missing value needs no scrub: x == 0 -> 0)
#pragma unroll
        for (int c = 0; c < 8; ++c) { tt[c] = scrub(tt[c], p.spt); sv[c] = scrub(sv[c], p.sps); }
    }
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 qa = z4, qb = z4;
    uint32_t pw0 = 0u, pw1 = 0u;
    if (wanted) {
        qa = prs[0]; qb = prs[blockDim.x];
        pw0 = __ldg(p.patw + d.pidx);
        if (d.need & kSigTwo) pw1 = __ldg(p.patw + d.pidx + 1);
    }
    const float pr[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
    unsigned need = 0u;   // -isodep only: bit c = cell c contributes
    if (ISO && wanted) {
        // -isodep takes every covered cell; those that are masked (itmask = 0), carry no transport and have a finite area
        // add exact zeros to all three histograms: drop them before the EOS (the dry cells below the sea floor)
        need = sig_covered(pw0) | (((d.need & kSigTwo) ? sig_covered(pw1) : 0u) << 4);
        const float4 *pa = reinterpret_cast<const float4 *>(p.area) + d.voff;
        const float4 aa = ld_stream_f4(pa, pol), ab = (d.need & kSigTwo) ? ld_stream_f4(pa + 1, pol) : z4;
        ar[0] = aa.x; ar[1] = aa.y; ar[2] = aa.z; ar[3] = aa.w; ar[4] = ab.x; ar[5] = ab.y; ar[6] = ab.z; ar[7] = ab.w;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const bool finite_area = (__float_as_uint(ar[c]) & 0x7f800000u) != 0x7f800000u;
            if (sv[c] == p.sps && pr[c] == 0.0f && finite_area) need &= ~(1u << c);
        }
    }
    // does cell c contribute?  The area field is zero outside the basins, so without -isodep: its transport is not +-0.
    auto contributes = [&](int c) { return ISO ? ((need >> c) & 1u) != 0u : pr[c] != 0.0f; };
    int ib[8];
    sig_group_bins<NEUTRAL>(tt, sv, p, ib, contributes);
    __syncwarp();
    // Keys (pattern << 16) + bin (both 1-based); a cell that does not contribute is transparent (it takes the key before it
    // and adds +-0), so equal keys around it still merge.
    int key[8];
    {
        int prev = -1;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int kc = (int)__byte_perm(c < 4 ? pw0 : pw1, 0u, 0x4044u + ((c & 3) << 8)) + ib[c];
            key[c] = contributes(c) ? kc : prev;
            prev = key[c];
        }
    }
    // Run merge in registers (a run's sum travels to its last cell), then the surviving entries into the lane's column of
    // the table.  On smooth fields a lane ends up with one entry, in cell 7.
    auto pass = [&](double *h, auto value_of) {
        int kk[8];
        double val[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { kk[c] = key[c]; val[c] = value_of(c); }
        bool extra = false;
#pragma unroll
        for (int c = 1; c < 8; ++c) {
            if (kk[c] == kk[c - 1]) {
                val[c] += val[c - 1];
                kk[c - 1] = -1;
            }
            extra = extra || kk[c - 1] >= 0;
        }
        sig_accumulate(h, tab, st, kk, val, __any_sync(kFull, extra), p.npat1, p.nbins, lane);
    };
    if (!ISO) {
        pass(hist, [&](int c) { return 0.0 - (double)pr[c]; });   // (+-0 where the cell does not contribute)
    } else {   // cdfmocsig.f90:427-428: gdep(jk)*itmask*zarea and itmask*zarea, REAL(4) chains; one table, three histograms
        const float gk = p.gdep[k];
        pass(hist, [&](int c) { return contributes(c) ? 0.0 - (double)pr[c] : 0.0; });
        sig_tab_flush(hist, tab, st, p.npat1, p.nbins, lane);
        pass(hist + hsize, [&](int c) {
            const float itm = (sv[c] == p.sps) ? 0.0f : 1.0f;
            return contributes(c) ? (double)__fmul_rn(__fmul_rn(gk, itm), ar[c]) : 0.0;
        });
        sig_tab_flush(hist + hsize, tab, st, p.npat1, p.nbins, lane);
        pass(hist + 2 * hsize, [&](int c) {
            const float itm = (sv[c] == p.sps) ? 0.0f : 1.0f;
            return contributes(c) ? (double)__fmul_rn(itm, ar[c]) : 0.0;
        });
        sig_tab_flush(hist + 2 * hsize, tab, st, p.npat1, p.nbins, lane);
    }
}

template <bool NEUTRAL, bool ISO>
__global__ void __launch_bounds__(kSigThreads, kSigMinCtas) mocsig_eos_hist_scan_kernel(const __grid_constant__ SigParams p)
{
    extern __shared__ double s_mem[];
    const int nwarps = blockDim.x >> 5, nthreads = blockDim.x;     // <= kSigWarps: chosen by the host
    const int hsize = p.nbins * p.npat1;                            // one private histogram
    constexpr int NH = ISO ? 3 : 1;                                 // histograms per warp: transport [, depth*area, area]
    double *hist_all = s_mem;                                       // [nwarps][NH][nbins][npat1]
    double *tab_all = hist_all + (size_t)nwarps * NH * hsize;       // [nwarps][kSigTabRows][32]
    double *comb = tab_all + (size_t)nwarps * kSigTabRows * 32;     // [nbins][nb]
    unsigned *s_poison = reinterpret_cast<unsigned *>(comb + (size_t)p.nbins * p.nb);   // [nbins]
    // fp32 transports of the two windows past stage 2, per thread: [slot][half][thread], 16-byte aligned
    float4 *s_prs = reinterpret_cast<float4 *>(s_mem + (((size_t)nwarps * (NH * hsize + kSigTabRows * 32) + (size_t)p.nbins * p.nb +
                                                          (p.nbins + 1) / 2 + 1) & ~(size_t)1));
    __shared__ int s_ticket[2];
    __shared__ int s_last, s_seg;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nzm1 = p.nz - 1;
    const uint64_t pol = make_evict_first_policy();
    double *hist = hist_all + (size_t)warp * NH * hsize;
    double *tab = tab_all + (size_t)warp * kSigTabRows * 32;
    for (int t = lane; t < kSigTabRows * 32; t += 32) tab[t] = 0.0;
    const int NV = (p.nx + 6) >> 2;                       // vectors per row, upper bound over the 4 alignments
    const int wpr = (NV + kSigWinVec - 1) / kSigWinVec;   // windows per (level) row
    // Programmatic dependent launch, as in K1 (moc_kernel.cuh): the CTAs of a launch retire over the time of a work unit
    // while the SMs idle; with PDL the CTAs of the next launch of the stream move in as they leave.  A dependent CTA works
    // through its first unit entirely in shared memory and executes griddepcontrol.wait before it stores anything; three
    // ticket generations rotate, block 0 re-arms the next one (and only then releases its own dependents) once the
    // predecessor is complete, and no CTA retires earlier.  When the inputs may have been produced by the predecessor
    // (byte swap, -vvl area build, a caller's kernel) the host clears pdl_early and every CTA waits before its first load.
    int *ticket = p.tickets + p.parity;
    bool must_wait = p.pdl != 0;
    if (p.pdl && (blockIdx.x == 0 || !p.pdl_early)) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        must_wait = false;
    }
    if (blockIdx.x == 0) {
        if (tid == 0) {
            p.tickets[(p.parity + 1) % 3] = 0;
            __threadfence();
        }
        __syncthreads();   // the re-armed ticket is visible before any warp of this CTA releases the dependents
    }
    if (p.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (tid == 0) s_ticket[0] = atomicAdd(ticket, 1);
    int tsel = 0;
    for (;;) {
        for (int t = tid; t < nwarps * NH * hsize; t += nthreads) hist_all[t] = 0.0;
        for (int t = tid; t < p.nbins; t += nthreads) s_poison[t] = 0u;
        if (tid == 0) s_seg = 0;
        __syncthreads();
        const int unit = s_ticket[tsel];
        if (unit >= p.nunits) {
            if (must_wait) asm volatile("griddepcontrol.wait;" ::: "memory");
            break;
        }
        if (tid == 0) s_ticket[tsel ^ 1] = atomicAdd(ticket, 1);  // next unit's ticket, latency hidden by this unit
        tsel ^= 1;
        const int j = unit / p.nchunk, ch = unit - j * p.nchunk;
        const int k0 = ch * p.nkc, nk = min(p.nkc, nzm1 - k0);
        const int jg = j + p.j_first_global;
        const bool skip_row = (p.ny_global > 1) && (jg == 0 || jg == p.ny_global - 1);  // jj = 2..npjglo-1 only (:405-407)

        if (!skip_row && nk > 0) {
            const int total = nk * wpr;   // windows of the unit, handed to the warps in runs by a shared-memory ticket
            SigTabState st;
            st.base = -0x40000000; st.dirty = false;
            SigStream ws;
            ws.w = 0; ws.w_end = 0; ws.win = 0; ws.k = 0; ws.nvec = 0; ws.voff0 = 0u; ws.pidx0 = 0u;
            SigS1 a;
            SigS2 x, y;
            float4 *const prx = s_prs + tid, *const pry = s_prs + 2 * nthreads + tid;
            sig_next_window(p, ws, a, &s_seg, j, k0, total, nwarps, wpr, lane, pol);
            sig_stage2<ISO>(p, a, x, prx, lane, pol);
#if CDF_SIG_PREFETCH_VA
            sig_next_window(p, ws, a, &s_seg, j, k0, total, nwarps, wpr, lane, pol);
#endif
            // x: window n (T / S in flight, transports in slot `slot`), a: window n+1 (V / area in flight)
            int slot = 0;
#pragma unroll 1
            while (x.live) {
#if CDF_SIG_PREFETCH_VA
                sig_stage2<ISO>(p, a, y, slot ? prx : pry, lane, pol);
                sig_next_window(p, ws, a, &s_seg, j, k0, total, nwarps, wpr, lane, pol);
#else
                sig_next_window(p, ws, a, &s_seg, j, k0, total, nwarps, wpr, lane, pol);
                sig_stage2<ISO>(p, a, y, slot ? prx : pry, lane, pol);
#endif
                sig_stage3<NEUTRAL, ISO>(p, x, slot ? pry : prx, hist, tab, st, s_poison, hsize, lane, pol);
                x = y;
                slot ^= 1;
            }
            sig_tab_flush(hist, tab, st, p.npat1, p.nbins, lane);
        }
        __syncthreads();
        if (must_wait) {   // PDL: the predecessor must be complete before this launch stores anything
            asm volatile("griddepcontrol.wait;" ::: "memory");
            must_wait = false;
        }
        // private histograms -> one (fixed order w = 0..nwarps-1: deterministic), in place in warp 0's copy
        for (int t = tid; t < NH * hsize; t += nthreads) {
            double hq = 0.0;
            for (int w = 0; w < nwarps; ++w) hq += hist_all[(size_t)w * NH * hsize + t];
            hist_all[t] = hq;
        }
        __syncthreads();
        if (p.nchunk > 1) {
            // this unit's partial histogram to global memory; the CTA that delivers the row's last chunk folds them in
            // chunk order (the order of arrival does not matter: bitwise reproducible)
            double *mine = p.part + (size_t)unit * NH * hsize;
            for (int t = tid; t < NH * hsize; t += nthreads) mine[t] = hist_all[t];
            unsigned *minep = p.partp + (size_t)unit * p.nbins;
            for (int t = tid; t < p.nbins; t += nthreads) minep[t] = s_poison[t];
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const int old = atomicAdd(p.col + j, 1);
                s_last = (old == p.nchunk - 1);
                if (s_last) p.col[j] = 0;   // re-armed for the next launch (stream order: nobody else touches it now)
            }
            __syncthreads();
            if (!s_last) continue;
            __threadfence();
            const double *rowp = p.part + (size_t)j * p.nchunk * NH * hsize;
            for (int t = tid; t < NH * hsize; t += nthreads) {
                double hq = 0.0;
                for (int c = 0; c < p.nchunk; ++c) hq += __ldcg(rowp + (size_t)c * NH * hsize + t);
                hist_all[t] = hq;
            }
            const unsigned *rowpp = p.partp + (size_t)j * p.nchunk * p.nbins;
            for (int t = tid; t < p.nbins; t += nthreads) {
                unsigned bits = 0u;
                for (int c = 0; c < p.nchunk; ++c) bits |= __ldcg(rowpp + (size_t)c * p.nbins + t);
                s_poison[t] = bits;
            }
            __syncthreads();
        }
        // patterns -> basins, /1e6, poison handling
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
                const double dd = combine(1, bin, b), w = combine(2, bin, b);
                p.out_iso[(size_t)j * p.nbins * p.nb + t] = (w != 0.0) ? dd / w : 99999.0;
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
template <bool NEUTRAL>
__global__ void mocsig_bins_kernel(const __grid_constant__ SigParams p, int32_t *__restrict__ ibin, size_t n, unsigned long long *stats)
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
        sig_group_bins<NEUTRAL>(tt, sv, p, ib, [&](int c) { return (need & (1u << c)) != 0u; }, stats);
#pragma unroll
        for (int c = 0; c < 8; ++c)
            if (need & (1u << c)) ibin[g8 * 8 + c] = ib[c];
    }
}

// setup kernel: area = fl32(e1v * e3v) (cdfmocsig.f90:390), e3v NOT masked -- and ZERO where the cell lies in no basin or in
// an excluded column (pattern byte 0 / 255; patw plane 0 is the unshifted one): such a cell contributes (0 - p) * 0 to every
// basin, i.e. nothing, so the fused kernel can take "transport is not +-0" as "the cell contributes" without reading the
// coverage.  A non-finite area is kept: its transport poisons the basins in the reference and takes the general path here.
__global__ void sig_prep_area_kernel(const float *__restrict__ e1v, const float *__restrict__ e3v, const uint32_t *__restrict__ patw,
                                     int nx, int pitchw, float *__restrict__ area, size_t nxy)
{
    const size_t k = blockIdx.y;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < nxy; c += (size_t)gridDim.x * blockDim.x) {
        const size_t j = c / (size_t)nx;
        const int i = (int)(c - j * (size_t)nx);
        const unsigned pat = (patw[j * (size_t)pitchw + (i >> 2)] >> (8 * (i & 3))) & 255u;
        const float a = __fmul_rn(e1v[c], e3v[k * nxy + c]);
        const bool finite = (__float_as_uint(a) & 0x7f800000u) != 0x7f800000u;
        area[k * nxy + c] = (pat != 0u && pat != 255u) || !finite ? a : 0.0f;
    }
}

}  // namespace cdfgpu
