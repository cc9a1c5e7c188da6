// moc_kernel_class.cuh -- K1, class formulation (EXPERIMENT, $CDFGPU_K1=class; not the default): the same zonal integral
// and vertical scan as moc_zonal_scan_kernel (moc_kernel.cuh; src/cdfmoc.f90:352-388) with fewer fp64 operations per cell.
//
// Idea: K1 is co-bound by instruction issue: per cell and per basin one SEL + one DFMA, and a DFMA
// occupies the issue port for two cycles -- 15 of the ~31 issue slots of a cell for 5 basins.  But the basin masks are
// not independent: every cell belongs to exactly one CLASS (distinct tuple of mask bits), classes are constant over long
// runs of i, and cells whose area is zero at every level (land at the surface) are TRANSPARENT: they add exactly +-0 to
// whatever sum they join.  The host therefore classifies every 128-cell segment of a row (the 32 float4 vectors one warp
// step covers, for each of the 4 row alignments) as
//     1..8  all non-transparent cells of the segment are in this class
//     0     they are in no basin                 (nothing to add)
//     0xFE  the segment is transparent           (nothing to add)
//     0xFF  mixed                                (a class boundary between wet cells: sector cuts, the i=1 / i=nx columns)
// and the warp, which gets the segment bytes of a trip with one uniform load (fetched one trip ahead), adds the four
// products of a lane's vector straight to the accumulator of that class behind a warp-uniform branch (4 F2F + 4 DADD per
// vector instead of 4 x 5 x (SEL + DFMA)).  Mixed segments take the per-cell path with one-hot class bytes.  At the end of
// a row the class sums are reduced across the warp and recombined into basin sums (basin b = sum of the classes that
// carry bit b).  Only the ORDER of the fp64 additions differs from the basin kernel; a NaN / Inf anywhere in the row still
// sends it to the literal chain.
//
// Measured (ORCA025, synthetic masks: 7 classes because the sub-basin masks live on T points and the global one on V
// points; 29 % of the segments mixed): parity-green (tests/test_gpu_cdfmoc.py, test_golden.py with $CDFGPU_K1=class),
// fp64 pipe 24 % -> 8 %, but 0.216 ms per record against 0.160 ms for the basin kernel (0.313 ms before the segment bytes
// were prefetched and the running sum replaced by direct class accumulation).  ncu: 97 M warp-instructions instead of
// 86 M -- the uniform branches per 128-cell segment cost register moves at every merge (IMAD.MOV 4.6 per cell), the
// branchy code keeps the four vector steps of a trip from overlapping (one dependent DADD chain per segment), and
// long-scoreboard stalls per issue double (10.5 vs 5.4).  The basin kernel's straight-line 5-chain code hides latency
// better than this formulation saves work.  Kept for the record.
#pragma once
#include "moc_kernel.cuh"

namespace cdfgpu {

constexpr int kMocMaxClasses = 8;                       // non-zero classes (one-hot byte)
__constant__ uint32_t c_k1_class_bits[kMocMaxClasses];  // class q+1 -> bit b set iff basin b contains the class

struct MocClassParams {
    MocParams m;                          // maskw is unused here
    const uint32_t *__restrict__ classw;  // [4][ny][pitchw] one-hot class bytes (0 = no basin / outside the row)
    const uint8_t *__restrict__ segtab;   // [4][ny][segpitch] segment classification (see above)
    int segpitch;                         // multiple of 4, >= trips * 4
    int nb;                               // basins
};

// literal chain for a row with runtime nb (rare: NaN / Inf in the row); stores the raw basin sums
__device__ __noinline__ void row_general_store_rt(const MocParams &p, int nb, int j, int k, int lane)
{
    const size_t e0 = ((size_t)k * p.ny + j) * (size_t)p.nx;
    double acc[CDFGPU_MAX_BASINS];
#pragma unroll
    for (int b = 0; b < CDFGPU_MAX_BASINS; ++b) acc[b] = 0.0;
    for (int i = lane; i < p.nx; i += kWarp) {
        const float a = p.area[e0 + i];
        const float v = p.zv[e0 + i];
        const int16_t *m = p.ibmask + ((size_t)j * p.nx + i) * nb;
#pragma unroll
        for (int b = 0; b < CDFGPU_MAX_BASINS; ++b)
            if (b < nb) acc[b] -= (double)__fmul_rn(__fmul_rn(a, (float)m[b]), v);
    }
#pragma unroll
    for (int b = 0; b < CDFGPU_MAX_BASINS; ++b) {
        if (b < nb) {
            const double t = warp_sum(acc[b]);
            if (lane == b) p.out[((size_t)k * p.ny + j) * nb + b] = t;
        }
    }
}

template <int NC, int MINB>
__global__ void __launch_bounds__(kMocThreads, MINB) moc_zonal_class_kernel(const MocClassParams q)
{
    constexpr int U = 4;                // vectors per lane and trip: one 32-bit word of segment bytes per trip
    const MocParams &p = q.m;
    const int nb = q.nb;
    const int lane = lane_id();
    const int nzm1 = p.nz - 1;
    const uint64_t pol = make_evict_first_policy();
    // ---- work units: sharded tickets with stealing, exactly as in moc_zonal_scan_kernel --------------------------
    const int chunks_per_col = (nzm1 + p.chunk - 1) / p.chunk;
    const int nunits = p.ny * chunks_per_col;
    int *tickets = p.tickets + p.parity * (kTicketShards * kTicketStride);
    {
        int *other = p.tickets + ((p.parity + 1) % 3) * (kTicketShards * kTicketStride);
        if (blockIdx.x == 0 && threadIdx.x < kTicketShards) other[threadIdx.x * kTicketStride] = 0;
    }
    int shard = (blockIdx.x * (kMocThreads / 32) + (threadIdx.x >> 5)) % kTicketShards;
    auto take = [&](int sh) { return atomicAdd(tickets + sh * kTicketStride, 1) * kTicketShards + sh; };
    int u = 0;
    if (lane == 0) u = take(shard);
    u = __shfl_sync(kFull, u, 0);
    for (;;) {
        if (u >= nunits) {
            const int seen = __ldcg(tickets + lane * kTicketStride) * kTicketShards + lane;
            unsigned live = __ballot_sync(kFull, seen < nunits);
            if (live == 0u) break;
            live = (live >> shard) | (shard ? (live << (32 - shard)) : 0u);
            shard = (shard + __ffs(live) - 1) % kTicketShards;
            if (lane == 0) u = take(shard);
            u = __shfl_sync(kFull, u, 0);
            continue;
        }
        int unext = 0;
        if (lane == 0) unext = take(shard);
        const int j = u / chunks_per_col;
        const int k0 = (u - j * chunks_per_col) * p.chunk;
        const int k1 = min(k0 + p.chunk, nzm1);
        for (int k = k0; k < k1; ++k) {
            // ---- one (j,k) row ----------------------------------------------------------------------------------
            const size_t e0 = ((size_t)k * p.ny + j) * (size_t)p.nx;
            const int s = (int)(e0 & 3);
            const size_t a0 = e0 - s;
            const int nvec = (s + p.nx + 3) >> 2;
            const float4 *__restrict__ v4 = reinterpret_cast<const float4 *>(p.zv + a0);
            const float4 *__restrict__ a4 = reinterpret_cast<const float4 *>(p.area + a0);
            const size_t rowsel = (size_t)s * p.ny + j;
            const uint32_t *__restrict__ cw = q.classw + rowsel * p.pitchw;
            const uint32_t *__restrict__ sg = reinterpret_cast<const uint32_t *>(q.segtab + rowsel * q.segpitch);
            double acc[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) acc[c] = 0.0;
            float badf = 0.0f;
            int trip = 0;
            uint32_t segw_next = __ldg(sg);   // segment bytes are fetched one trip ahead: their latency never gates the
                                              // trip's streaming loads
            for (int v0 = lane; v0 < nvec; v0 += kWarp * U, ++trip) {
                float4 vv[U], aa[U];
                uint32_t mm[U];
#pragma unroll
                for (int w = 0; w < U; ++w) {
                    const int vi = v0 + w * kWarp;
                    if (vi < nvec) {
                        vv[w] = ld_stream_f4(v4 + vi, pol);
                        aa[w] = ld_stream_f4(a4 + vi, pol);
                    } else {
                        vv[w] = make_float4(0.f, 0.f, 0.f, 0.f);
                        aa[w] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                const uint32_t segw = segw_next;          // the trip's four segment bytes (warp-uniform)
                segw_next = __ldg(sg + trip + 1);         // (the table row is padded by one word)
#pragma unroll
                for (int w = 0; w < U; ++w) {
                    const int vi = v0 + w * kWarp;
                    const bool mixed = ((segw >> (8 * w)) & 0xffu) == 0xffu;
                    mm[w] = (mixed && vi < nvec) ? __ldg(cw + vi) : 0u;
                }
#pragma unroll
                for (int w = 0; w < U; ++w) {
                    const int sc = (int)((segw >> (8 * w)) & 0xffu);   // warp-uniform
                    if (sc == 0xff) {            // mixed segment: per cell, one-hot class bytes
                        moc_cell<NC, 0>(aa[w].x, vv[w].x, mm[w], acc, badf);
                        moc_cell<NC, 8>(aa[w].y, vv[w].y, mm[w], acc, badf);
                        moc_cell<NC, 16>(aa[w].z, vv[w].z, mm[w], acc, badf);
                        moc_cell<NC, 24>(aa[w].w, vv[w].w, mm[w], acc, badf);
                    } else {
                        const float p0 = __fmul_rn(aa[w].x, vv[w].x), p1 = __fmul_rn(aa[w].y, vv[w].y);
                        const float p2 = __fmul_rn(aa[w].z, vv[w].z), p3 = __fmul_rn(aa[w].w, vv[w].w);
                        if (sc >= 1 && sc <= NC) {
                            // one class for the whole segment: a NaN / Inf product makes the class sum non-finite, which
                            // the end of the row checks -- no per-cell flag needed here
#pragma unroll
                            for (int c = 0; c < NC; ++c)
                                if (sc == c + 1) {
                                    acc[c] -= (double)p0;
                                    acc[c] -= (double)p1;
                                    acc[c] -= (double)p2;
                                    acc[c] -= (double)p3;
                                }
                        } else {                 // nothing to add; the reference still propagates NaN / Inf from here
                            badf = __fmaf_rn(p0, 0.0f, badf);
                            badf = __fmaf_rn(p1, 0.0f, badf);
                            badf = __fmaf_rn(p2, 0.0f, badf);
                            badf = __fmaf_rn(p3, 0.0f, badf);
                        }
                    }
                }
            }
            bool bad = badf != badf;
#pragma unroll
            for (int c = 0; c < NC; ++c) bad |= (__double2hiint(acc[c]) & 0x7ff00000) == 0x7ff00000;   // NaN or Inf
            if (p.general || __any_sync(kFull, bad)) {
                row_general_store_rt(p, nb, j, k, lane);
            } else {
                double tb = 0.0;   // lane b: raw sum of basin b = sum of the classes that carry bit b
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    const double t = warp_sum(acc[c]);
                    if ((c_k1_class_bits[c] >> lane) & 1u) tb += t;
                }
                if (lane < nb) p.out[((size_t)k * p.ny + j) * nb + lane] = tb;
            }
        }
        // ---- end of unit: publish the rows, count them; the warp that completes column j integrates it vertically
        __syncwarp();
        int done = 0;
        if (lane == 0) {
            int old;
            asm volatile("atom.add.release.gpu.global.s32 %0, [%1], %2;" : "=r"(old) : "l"(p.col_done + j), "r"(k1 - k0) : "memory");
            done = old + (k1 - k0);
        }
        done = __shfl_sync(kFull, done, 0);
        if (done == nzm1) {
            if (!p.noscan) {
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                if (lane < nb) {   // psi(k) = psi(k+1) + T(k)/1.d6 in the reference's order (cdfmoc.f90:385); loads 8 ahead
                    double psi = 0.0;
                    p.out[((size_t)nzm1 * p.ny + j) * nb + lane] = 0.0;  // dmoc(:,:,npk) stays 0
                    int kk = nzm1 - 1;
                    for (; kk >= 7; kk -= 8) {
                        double h[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) h[r] = __ldcg(p.out + ((size_t)(kk - r) * p.ny + j) * nb + lane) / 1.0e6;
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            psi = psi + h[r];
                            p.out[((size_t)(kk - r) * p.ny + j) * nb + lane] = psi;
                        }
                    }
                    for (; kk >= 0; --kk) {
                        psi = psi + __ldcg(p.out + ((size_t)kk * p.ny + j) * nb + lane) / 1.0e6;
                        p.out[((size_t)kk * p.ny + j) * nb + lane] = psi;
                    }
                }
            }
            if (lane == 0) p.col_done[j] = 0;  // self-reset for the next launch
            __syncwarp();
        }
        u = __shfl_sync(kFull, unext, 0);
    }
}

}  // namespace cdfgpu
