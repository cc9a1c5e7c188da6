// moc_kernel_tma.cuh -- K1 (fast path): depth-space MOC as a TMA-fed streaming reduction.
// Replaces the loop nest src/cdfmoc.f90:352-388 for 0/1 basin masks (the general INTEGER(2) case and rows that
// contain NaN/Inf fall back to the literal chain in moc_kernel.cuh).
//
//   T(b,j,k) = sum_i -dble( fl32( fl32( fl32(e1v*e3m) * real(ibmask(b)) ) * zv ) )       (cdfmoc.f90:373-374)
//   psi(b,j,nz) = 0 ; psi(b,j,k) = psi(b,j,k+1) + T(b,j,k)/1.d6 , k = nz-1..1            (cdfmoc.f90:385)
//
// Data movement: every warp runs its own D-stage ring in shared memory.  One elected lane issues 1-D bulk async
// copies (cp.async.bulk, the TMA unit; SASS UBLKCP) of a row tile of V, of the resident area = fl32(e1v*e3m) and of
// the row's mask-class bytes, completing on an mbarrier; the bytes in flight (D-1 stages x warps x 13 KB per SM)
// are decoupled from the register file, so HBM latency is covered without occupancy.  L2 policy: V and area are
// evict-first (read once), the mask-class plane is evict-last (re-read by every level).
//
// Arithmetic: a lane owns L CONSECUTIVE cells of the tile (L = 4 mod 8 words makes the 16-byte shared loads
// conflict-free).  The five basin masks are folded at setup into one byte per (j,i): the index of the distinct
// 0/1 mask tuple at that cell ("class", 0 = no basin, 255 = outside the row).  Along a lane's cells the class
// changes only at coasts and basin limits, so the lane keeps ONE running fp64 sum (4 FMUL + 4 F2F + 4 DADD per
// 16-byte vector) and flushes it into a per-lane per-class slot in shared memory when the class changes.  At the end
// of the row the class sums are reduced across lanes (shuffle tree) and combined into basins.  The fp32 product is
// the reference's own chain for mask = 1 (fl32(area*1) == area); mask = 0 cells add -(+-0) in the reference, i.e.
// nothing.  Non-finite class sums (a NaN/Inf somewhere in the row) send the row to the literal chain.
//
// Work distribution: persistent grid, units of `chunk` consecutive levels of one latitude row j handed out by an
// atomic ticket (fetched one unit ahead, so the ring never drains at unit boundaries); the warp that completes the
// last level of a column integrates it bottom-up in the reference's sequential order.
#pragma once
#include "common.cuh"
#include "moc_kernel.cuh"

namespace cdfgpu {

constexpr int kTmaStages = 3;
constexpr int kTmaMaxClasses = 8;   // distinct mask tuples incl. the all-zero one; more -> general kernel
constexpr int kTmaMaxLaneCells = 60;

__constant__ uint32_t c_class_bits[kTmaMaxClasses];  // class -> bit b set iff basin b covers the class

struct MocTmaParams {
    const float *__restrict__ zv;        // (nz-1, ny, nx)
    const float *__restrict__ area;      // (nz-1, ny, nx) fl32(e1v*e3m)
    const uint8_t *__restrict__ classes; // [4][ny][cpitch] class bytes, copy s shifted by s cells, 255 outside the row
    const int16_t *__restrict__ ibmask;  // (ny, nx, nb) for the per-row fallback
    double *__restrict__ out;            // (nz, ny, nb)
    int *tickets;                        // [2]
    int *col_done;                       // [ny]
    int nx, ny, nz, nclass;
    int lane_cells;                      // L: cells per lane per tile, L % 8 == 4
    int ntile;                           // tiles per row
    int cpitch;                          // bytes per class row = ntile * 32 * L
    int parity, chunk;
    int warps;                           // warps per CTA (shared-memory carve-up)
};

// ---- PTX wrappers: mbarrier + bulk async copy (TMA) --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t pol)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
        : "memory");
}
__device__ __forceinline__ uint64_t make_evict_last_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}

// The deterministic sequence of tiles a warp walks: units come from the ticket counter, a unit is `chunk` levels of
// one latitude row, a level-row is `ntile` tiles.  Producer and consumer each hold one cursor over the same sequence.
struct TileCursor {
    int j, k, k1, tile;  // current tile; k1 = end of the unit's level range
    bool valid;
};

template <int NB>
__global__ void __launch_bounds__(256, 1) moc_zonal_scan_tma_kernel(const MocTmaParams p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const int nzm1 = p.nz - 1;
    const int L = p.lane_cells;
    const int TC = 32 * L;                         // cells per tile
    // per-warp carve-up: [stages][V TC f32 | area TC f32 | class TC u8] | class sums [nclass][32] f64 | scan | barriers
    const size_t stage_bytes = (size_t)TC * 9;
    const size_t warp_bytes = kTmaStages * stage_bytes + (size_t)kTmaMaxClasses * 32 * 8 + (size_t)nzm1 * NB * 8 + 64;
    unsigned char *wbase = smem_raw + (size_t)warp * ((warp_bytes + 127) & ~(size_t)127);
    double *csum = reinterpret_cast<double *>(wbase + kTmaStages * stage_bytes);
    double *sc = csum + kTmaMaxClasses * 32;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sc + (size_t)nzm1 * NB);

    const uint64_t pol_stream = make_evict_first_policy();
    const uint64_t pol_keep = make_evict_last_policy();
    const int chunks_per_col = (nzm1 + p.chunk - 1) / p.chunk;
    const int nunits = p.ny * chunks_per_col;
    int *ticket = p.tickets + p.parity;
    if (blockIdx.x == 0 && threadIdx.x == 0) p.tickets[p.parity ^ 1] = 0;  // re-arm the next launch's counter

    if (lane == 0)
        for (int s = 0; s < kTmaStages; ++s) mbar_init(bars + s, 1);
    for (int t = lane; t < kTmaMaxClasses * 32; t += 32) csum[t] = 0.0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();

    // ---- unit look-ahead: the ticket after the producer's current unit is requested early and only read (shuffled
    // out of lane 0) when the producer crosses the unit boundary, so the atomic's latency is never exposed ----------
    auto unit_to_cursor = [&](int u, TileCursor &c) {
        c.valid = u < nunits;
        if (!c.valid) return;
        c.j = u / chunks_per_col;
        c.k = (u - c.j * chunks_per_col) * p.chunk;
        c.k1 = min(c.k + p.chunk, nzm1);
        c.tile = 0;
    };
    int ahead_l0 = 0;          // meaningful in lane 0 only
    if (lane == 0) ahead_l0 = atomicAdd(ticket, 1);
    TileCursor prod;
    auto next_unit = [&]() {
        const int u = __shfl_sync(kFull, ahead_l0, 0);
        unit_to_cursor(u, prod);
        if (prod.valid && lane == 0) ahead_l0 = atomicAdd(ticket, 1);
        if (!prod.valid) ahead_l0 = nunits;
    };
    next_unit();

    // metadata of the tile sitting in each stage (registers, indexed statically by the unrolled stage loop)
    int m_j[kTmaStages], m_k[kTmaStages], m_tile[kTmaStages], m_k1[kTmaStages];
    bool m_valid[kTmaStages];

    auto produce = [&](int s, uint32_t) {
        m_valid[s] = prod.valid;
        if (!prod.valid) return;
        m_j[s] = prod.j; m_k[s] = prod.k; m_tile[s] = prod.tile; m_k1[s] = prod.k1;
        const size_t e0 = ((size_t)prod.k * p.ny + prod.j) * (size_t)p.nx;
        const int sh = (int)(e0 & 3);
        const int first = prod.tile * TC;                       // first shifted cell index of the tile
        const int ncell = min(TC, ((sh + p.nx + 3) & ~3) - first);   // valid cells, multiple of 4
        if (lane == 0) {
            unsigned char *st = wbase + (size_t)s * stage_bytes;
            const uint32_t bytes_f = (uint32_t)ncell * 4u;
            const uint32_t bytes_c = (uint32_t)((ncell + 15) & ~15);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(bars + s, 2u * bytes_f + bytes_c);
            tma_load_1d(st, p.zv + (e0 - sh) + first, bytes_f, bars + s, pol_stream);
            tma_load_1d(st + (size_t)TC * 4, p.area + (e0 - sh) + first, bytes_f, bars + s, pol_stream);
            tma_load_1d(st + (size_t)TC * 8, p.classes + ((size_t)sh * p.ny + prod.j) * p.cpitch + first, bytes_c,
                        bars + s, pol_keep);
        }
        // advance the producer cursor
        if (++prod.tile == p.ntile || prod.tile * TC >= sh + p.nx) {
            prod.tile = 0;
            if (++prod.k == prod.k1) next_unit();
        }
    };

#pragma unroll
    for (int s = 0; s < kTmaStages; ++s) produce(s, 0);

    double cur = 0.0;        // running sum of the lane's current class
    int curc = 255;          // current class (255 = outside the row: dropped)
    uint32_t phase = 0;
    bool running = true;
    while (running) {
#pragma unroll
        for (int s = 0; s < kTmaStages; ++s) {
            if (!m_valid[s]) { running = false; break; }
            mbar_wait(bars + s, phase);
            const int j = m_j[s], k = m_k[s], tile = m_tile[s], unit_k1 = m_k1[s];
            const size_t e0 = ((size_t)k * p.ny + j) * (size_t)p.nx;
            const int sh = (int)(e0 & 3);
            const int first = tile * TC;
            const int ncell = min(TC, ((sh + p.nx + 3) & ~3) - first);
            const unsigned char *st = wbase + (size_t)s * stage_bytes;
            const float4 *sv = reinterpret_cast<const float4 *>(st) + lane * (L / 4);
            const float4 *sa = reinterpret_cast<const float4 *>(st + (size_t)TC * 4) + lane * (L / 4);
            const uint32_t *sm = reinterpret_cast<const uint32_t *>(st + (size_t)TC * 8) + lane * (L / 4);
            const int nv = max(0, min(L, ncell - lane * L)) >> 2;   // this lane's vectors in the tile
            uint32_t curw = (uint32_t)curc * 0x01010101u;
            for (int v = 0; v < nv; ++v) {
                const float4 vv = sv[v], aa = sa[v];
                const uint32_t mw = sm[v];
                const float p0 = __fmul_rn(aa.x, vv.x), p1 = __fmul_rn(aa.y, vv.y);
                const float p2 = __fmul_rn(aa.z, vv.z), p3 = __fmul_rn(aa.w, vv.w);
                if (mw == curw) {
                    cur += ((double)p0 + (double)p1) + ((double)p2 + (double)p3);
                } else {
                    const float pp[4] = {p0, p1, p2, p3};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int cls = (int)((mw >> (8 * c)) & 255u);
                        if (cls != curc) {
                            if (curc != 255) csum[curc * 32 + lane] += cur;
                            cur = 0.0;
                            curc = cls;
                        }
                        cur += (double)pp[c];
                    }
                    curw = (uint32_t)curc * 0x01010101u;
                }
            }
            const bool row_done = (tile + 1 == p.ntile) || ((tile + 1) * TC >= sh + p.nx);
            __syncwarp();
            produce(s, phase);   // refill this stage with the tile kTmaStages ahead (all lanes finished reading it)

            if (row_done) {
                if (curc != 255) csum[curc * 32 + lane] += cur;
                cur = 0.0;
                curc = 255;
                __syncwarp();
                double tb = 0.0;
                bool bad = false;
                for (int q = 0; q < p.nclass; ++q) {
                    double x = csum[q * 32 + lane];
                    csum[q * 32 + lane] = 0.0;
                    x = warp_sum(x);
                    bad = bad || !(fabs(x) <= 1.7976931348623157e308);   // NaN or Inf
                    if ((c_class_bits[q] >> lane) & 1u) tb += x;
                }
                if (bad) {
                    MocParams g;   // literal chain for this row (NaN/Inf semantics of the reference)
                    g.zv = p.zv; g.area = p.area; g.maskw = nullptr; g.ibmask = p.ibmask; g.out = p.out;
                    g.tickets = nullptr; g.col_done = nullptr;
                    g.nx = p.nx; g.ny = p.ny; g.nz = p.nz; g.pitchw = 0; g.parity = 0; g.chunk = 1; g.general = 1; g.noscan = 0;
                    row_general_store<NB>(g, j, k, lane);
                } else if (lane < NB) {
                    p.out[((size_t)k * p.ny + j) * NB + lane] = 0.0 - tb;
                }
                // the unit's last level: publish the rows, count them; the warp that completes column j integrates it
                if (k + 1 == unit_k1) {
                    __syncwarp();
                    int done = 0;
                    const int k0 = (k / p.chunk) * p.chunk;
                    if (lane == 0) {
                        __threadfence();
                        done = atomicAdd(p.col_done + j, k + 1 - k0) + (k + 1 - k0);
                    }
                    done = __shfl_sync(kFull, done, 0);
                    if (done == nzm1) {
                        __threadfence();
                        for (int t = lane; t < nzm1 * NB; t += kWarp) {
                            const int kk = t / NB, b = t - kk * NB;
                            sc[t] = __ldcg(p.out + ((size_t)kk * p.ny + j) * NB + b) / 1.0e6;  // dmoc(:,jj,jk)/1.d6
                        }
                        __syncwarp();
                        if (lane < NB) {
                            double psi = 0.0;
                            p.out[((size_t)nzm1 * p.ny + j) * NB + lane] = 0.0;  // dmoc(:,:,npk) stays 0
                            for (int kk = nzm1 - 1; kk >= 0; --kk) {
                                psi = psi + sc[kk * NB + lane];  // dmoc(:,jj,jk+1) + dmoc(:,jj,jk)/1.d6
                                p.out[((size_t)kk * p.ny + j) * NB + lane] = psi;
                            }
                        }
                        if (lane == 0) p.col_done[j] = 0;  // self-reset for the next launch
                        __syncwarp();
                    }
                }
            }
        }
        phase ^= 1u;
    }
}

}  // namespace cdfgpu
