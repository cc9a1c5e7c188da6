// moc_kernel_pipe.cuh -- K1, asynchronously staged (EXPERIMENT, not the default): same arithmetic, work distribution and
// epilogue as moc_zonal_scan_kernel (moc_kernel.cuh; src/cdfmoc.f90:352-388), but a warp's loads never drain.
//
// Idea: in moc_zonal_scan_kernel a warp alternates between a load phase (one batch of 16-byte loads into registers, then
// ~1 us of HBM latency) and a compute phase; while it computes it has nothing in flight, and at the end of every (j,k)
// row its pipeline drains (warp reduction, store, next row's address set-up).  Here the (unit, level, trip) iteration
// space of a warp is FLAT -- across the rows of a work unit and across work units -- and every warp owns a ring of S
// stages in shared memory, filled either with cp.async (LDGSTS; BULK = false) or with TMA bulk copies that complete on
// the stage's mbarrier (cp.async.bulk, one elected lane issues three copies per batch -- V, area, mask words; BULK =
// true): S-1 batches are always outstanding while one is consumed.  A lane reads back exactly the bytes it (or the
// TMA on its behalf) wrote, so the only synchronisation is cp.async.wait_group / the mbarrier phase.
//
// Measured (ORCA025, B200; tools/k1_variants.py, gpurun_out/k1_variants_[ab].log): bit-identical results, but
// 0.25-0.36 ms per record for every depth S = 2..6, batch size and occupancy (16-32 warps / SM), with either feed,
// against 0.160 ms for the register-staged kernel.  ncu (`--set full`, variants 10 and 23): 117-130 M warp-instructions
// per record instead of 86 M (LDS, stage addressing, per-batch cursor logic), issue slots 58-63 % busy with
// `not_selected` the top stall, DRAM 40 %.  So K1 is co-bound by instruction issue, not by bytes in flight: a DFMA
// occupies the issue port for two cycles on sm_100 (fp64 at half rate), which makes the per-basin SEL + DFMA pairs 15 of
// the ~31 issue slots a cell costs, and 24 resident warps of the register kernel already cover the HBM latency.  Any
// staging scheme that adds instructions per cell loses.  Kept selectable ($CDFGPU_K1_VARIANT=10 / 23) for the record.
#pragma once
#include "moc_kernel.cuh"
#include "moc_kernel_tma.cuh"   // mbar_init / mbar_expect_tx / mbar_wait / tma_load_1d

namespace cdfgpu {

// 16-byte asynchronous copy, L2 evict-first, zero-fill when !in (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async16_zfill(void *dst, const void *src, bool in, uint64_t pol)
{
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"((unsigned)__cvta_generic_to_shared(dst)),
                 "l"(src), "r"(in ? 16 : 0), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(void *dst, const void *src, bool in)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src),
                 "r"(in ? 4 : 0)
                 : "memory");
}

// warp-uniform position in the flat (unit, level, trip) stream
struct MocCursor {
    int j, k, k0, k1, trip;
};

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <int U>
struct alignas(128) MocStage {   // one batch of one warp: lane l owns element [u][l] of each array
    float4 v[U][kWarp];
    float4 a[U][kWarp];
    uint32_t m[U][kWarp];
};

template <int U>
__device__ __forceinline__ void moc_issue_batch(const MocParams &p, const MocCursor &c, int lane, uint64_t pol, MocStage<U> &st)
{
    const size_t e0 = ((size_t)c.k * p.ny + c.j) * (size_t)p.nx;
    const int s = (int)(e0 & 3);
    const size_t a0 = e0 - s;
    const int nvec = (s + p.nx + 3) >> 2;
    const float4 *__restrict__ v4 = reinterpret_cast<const float4 *>(p.zv + a0);
    const float4 *__restrict__ a4 = reinterpret_cast<const float4 *>(p.area + a0);
    const uint32_t *__restrict__ mw = p.maskw + ((size_t)s * p.ny + c.j) * p.pitchw;
    const int v0 = c.trip * (kWarp * U) + lane;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int vi = v0 + u * kWarp;
        const bool in = vi < nvec;
        const int vs = in ? vi : 0;   // keep the (unread) source address inside the row
        cp_async16_zfill(&st.v[u][lane], v4 + vs, in, pol);
        cp_async16_zfill(&st.a[u][lane], a4 + vs, in, pol);
        cp_async4_zfill(&st.m[u][lane], mw + vs, in);
    }
}

// The same batch by TMA: vectors [trip*32U, trip*32U + nvalid) of the row, contiguous in global and in the stage.
// Called by ONE lane.  Returns nvalid (0 when this alignment of the row has no vector left in its last trip).
template <int U>
__device__ __forceinline__ int moc_issue_batch_bulk(const MocParams &p, const MocCursor &c, uint64_t pol_stream, uint64_t pol_keep,
                                                    MocStage<U> &st, uint64_t *bar)
{
    const size_t e0 = ((size_t)c.k * p.ny + c.j) * (size_t)p.nx;
    const int s = (int)(e0 & 3);
    const size_t a0 = e0 - s;
    const int nvec = (s + p.nx + 3) >> 2;
    const int first = c.trip * (kWarp * U);
    const int nvalid = max(0, min(nvec - first, kWarp * U));
    if (nvalid == 0) {
        mbar_arrive(bar);
        return 0;
    }
    const uint32_t bytes_v = (uint32_t)nvalid * 16u;
    const uint32_t bytes_m = (uint32_t)((nvalid + 3) & ~3) * 4u;   // the mask planes are padded to a multiple of 4 words
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage's previous readers (generic proxy) are done
    mbar_expect_tx(bar, 2u * bytes_v + bytes_m);
    tma_load_1d(&st.v[0][0], p.zv + a0 + 4 * (size_t)first, bytes_v, bar, pol_stream);
    tma_load_1d(&st.a[0][0], p.area + a0 + 4 * (size_t)first, bytes_v, bar, pol_stream);
    tma_load_1d(&st.m[0][0], p.maskw + ((size_t)s * p.ny + c.j) * p.pitchw + first, bytes_m, bar, pol_keep);
    return nvalid;
}

template <int NB, int U, int S, int MINB, bool BULK>
__global__ void __launch_bounds__(kMocThreads, MINB) moc_zonal_scan_pipe_kernel(const MocParams p)
{
    extern __shared__ __align__(128) unsigned char s_raw[];
    constexpr int kWarps = kMocThreads / 32;
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    MocStage<U> *ring = reinterpret_cast<MocStage<U> *>(s_raw) + (size_t)warp * S;
    int4 *meta = reinterpret_cast<int4 *>(s_raw + sizeof(MocStage<U>) * (size_t)S * kWarps) + warp * S;   // per stage: j, k | flags, nvalid
    uint64_t *bars = reinterpret_cast<uint64_t *>(s_raw + (sizeof(MocStage<U>) + sizeof(int4)) * (size_t)S * kWarps) + warp * S;
    uint64_t pol_keep = 0;
    if (BULK) {
        pol_keep = make_evict_last_policy();
        if (lane == 0)
            for (int i = 0; i < S; ++i) mbar_init(bars + i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
    }
    const int nzm1 = p.nz - 1;
    const uint64_t pol = make_evict_first_policy();
    const int tpr = (((p.nx + 6) >> 2) + kWarp * U - 1) / (kWarp * U);   // trips per row (upper bound over the alignments)
    const int chunks_per_col = (nzm1 + p.chunk - 1) / p.chunk;
    const int nunits = p.ny * chunks_per_col;
    int *tickets = p.tickets + p.parity * (kTicketShards * kTicketStride);
    {   // re-arm the next launch's counters
        int *other = p.tickets + ((p.parity + 1) % 3) * (kTicketShards * kTicketStride);
        if (blockIdx.x == 0 && threadIdx.x < kTicketShards) other[threadIdx.x * kTicketStride] = 0;
    }
    // ---- work units: sharded tickets with stealing, exactly as in moc_zonal_scan_kernel --------------------------
    int shard = (blockIdx.x * kWarps + warp) % kTicketShards;
    auto take = [&](int sh) { return atomicAdd(tickets + sh * kTicketStride, 1) * kTicketShards + sh; };
    int unext = 0;   // lane 0: the ticket requested ahead
    if (lane == 0) unext = take(shard);
    auto next_unit = [&](MocCursor &c) -> bool {   // next work unit into the cursor; false when every shard is dry
        int u = __shfl_sync(kFull, unext, 0);
        while (u >= nunits) {
            const int seen = __ldcg(tickets + lane * kTicketStride) * kTicketShards + lane;
            unsigned live = __ballot_sync(kFull, seen < nunits);
            if (live == 0u) return false;
            live = (live >> shard) | (shard ? (live << (32 - shard)) : 0u);
            shard = (shard + __ffs(live) - 1) % kTicketShards;
            if (lane == 0) u = take(shard);
            u = __shfl_sync(kFull, u, 0);
        }
        if (lane == 0) unext = take(shard);   // the unit after this one: its latency hides behind this unit's rows
        c.j = u / chunks_per_col;
        c.k0 = (u - c.j * chunks_per_col) * p.chunk;
        c.k1 = min(c.k0 + p.chunk, nzm1);
        c.k = c.k0;
        c.trip = 0;
        return true;
    };
    auto advance = [&](MocCursor &c) -> bool {
        if (++c.trip < tpr) return true;
        c.trip = 0;
        if (++c.k < c.k1) return true;
        return next_unit(c);
    };

    MocCursor P;
    bool more = next_unit(P);
    int sp = 0, sc = 0;          // ring positions of the producer / consumer
    int outstanding = 0;         // batches issued and not yet consumed
    auto produce = [&]() {
        if (more) {
            if (!BULK) moc_issue_batch<U>(p, P, lane, pol, ring[sp]);
            if (lane == 0) {
                int nvalid = kWarp * U;
                if (BULK) nvalid = moc_issue_batch_bulk<U>(p, P, pol, pol_keep, ring[sp], bars + sp);
                const int rowend = (P.trip == tpr - 1) ? 1 : 0;
                const int unitend = (rowend && P.k == P.k1 - 1) ? 1 : 0;
                meta[sp] = make_int4(P.j, P.k | (rowend << 16) | (unitend << 17) | ((P.k1 - P.k0) << 18), nvalid, 0);
            }
            sp = (sp + 1 == S) ? 0 : sp + 1;
            ++outstanding;
            more = advance(P);
        }
        if (!BULK) cp_async_commit();   // one group per call, empty or not: the group arithmetic below stays uniform
    };
#pragma unroll 1
    for (int i = 0; i < S - 1; ++i) produce();

    double acc[NB];
    float badf = 0.0f;
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.0;

#pragma unroll 1
    uint32_t cphase = 0;         // BULK: parity of the consumer's current pass over the ring
    while (outstanding > 0) {
        produce();
        if (BULK) mbar_wait(bars + sc, cphase);
        else cp_async_wait<S - 1>();   // everything but the newest S-1 groups has landed: the batch at `sc` is complete
        __syncwarp();             // lane 0's meta word
        const MocStage<U> &st = ring[sc];
        const int4 mt = meta[sc];
        if (sc + 1 == S) { sc = 0; cphase ^= 1u; } else ++sc;
        --outstanding;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!BULK || u * kWarp + lane < mt.z) {
                const float4 v = st.v[u][lane], a = st.a[u][lane];
                const uint32_t m = st.m[u][lane];
                moc_cell<NB, 0>(a.x, v.x, m, acc, badf);
                moc_cell<NB, 8>(a.y, v.y, m, acc, badf);
                moc_cell<NB, 16>(a.z, v.z, m, acc, badf);
                moc_cell<NB, 24>(a.w, v.w, m, acc, badf);
            }
        }
        if (BULK) __syncwarp();   // every lane has read its part of the stage before lane 0 refills it
        if (!(mt.y & 0x10000)) continue;
        // ---- end of row (j,k) -----------------------------------------------------------------------------------
        const int j = mt.x, k = mt.y & 0xffff;
        if (__any_sync(kFull, badf != badf)) {
            row_general_store<NB>(p, j, k, lane);
        } else {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
                const double t = warp_sum(acc[b]);
                if (lane == b) p.out[((size_t)k * p.ny + j) * NB + b] = t;
            }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[b] = 0.0;
        badf = 0.0f;
        if (!(mt.y & 0x20000)) continue;
        // ---- end of unit: publish the rows, count them; the warp that completes column j integrates it vertically
        const int cnt = mt.y >> 18;
        __syncwarp();
        int done = 0;
        if (lane == 0) {
            int old;
            asm volatile("atom.add.release.gpu.global.s32 %0, [%1], %2;" : "=r"(old) : "l"(p.col_done + j), "r"(cnt) : "memory");
            done = old + cnt;
        }
        done = __shfl_sync(kFull, done, 0);
        if (done != nzm1) continue;
        if (!p.noscan) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            if (lane < NB) {   // psi(k) = psi(k+1) + T(k)/1.d6 in the reference's order (cdfmoc.f90:385); loads 8 ahead
                double psi = 0.0;
                p.out[((size_t)nzm1 * p.ny + j) * NB + lane] = 0.0;  // dmoc(:,:,npk) stays 0
                int kk = nzm1 - 1;
                for (; kk >= 7; kk -= 8) {
                    double h[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) h[q] = __ldcg(p.out + ((size_t)(kk - q) * p.ny + j) * NB + lane) / 1.0e6;
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        psi = psi + h[q];
                        p.out[((size_t)(kk - q) * p.ny + j) * NB + lane] = psi;
                    }
                }
                for (; kk >= 0; --kk) {
                    psi = psi + __ldcg(p.out + ((size_t)kk * p.ny + j) * NB + lane) / 1.0e6;
                    p.out[((size_t)kk * p.ny + j) * NB + lane] = psi;
                }
            }
        }
        if (lane == 0) p.col_done[j] = 0;  // self-reset for the next launch
        __syncwarp();
    }
}

}  // namespace cdfgpu
