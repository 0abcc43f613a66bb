// scan_tma.cu - first-dimension scan (multiplyQueryByDatabase, src/spiral.cpp:628-999) with the database streamed through
// shared memory by the TMA engine (cp.async.bulk + mbarrier ring) instead of through registers.
//
// Why: k_scan_spiral keeps 24 accumulator registers per thread, so ptxas holds only ~5 of its 8 unrolled 16-byte loads in flight
// and an SM has ~80 KiB outstanding - 6.0-6.5 TB/s, while a read-only stream with 130 KiB outstanding per SM reaches 7.1-7.3 TB/s
// on the same GPU (k_scan_pack_wide, scripts/micro/pipes.cu).  Here the bytes in flight are a ring of NS tiles of 16 KiB per CTA,
// filled by one producer lane; the 256 consumer threads read the tiles with conflict-free 128-bit shared loads and run the same
// 12 MACs per 16 bytes.  Same arithmetic, same results, same output layout as k_scan_spiral.
//
//   CTA = one z-slice x 256 database columns (or the whole shard when narrower): 8 consumer warps + 1 producer warp
//   smem = NS x 16 KiB tiles | the z-slice of the query (dim0 x 64 B, one bulk copy) | 2 NS + 1 mbarriers
// Measured (cfg1 2 GiB / cfg5 8 GiB, scan stage between CUDA events): k_scan_spiral 0.356 / 1.346 ms; this kernel 0.336 / 1.269 ms
// = 6.39 / 6.77 TB/s.  Tiles of 2 / 4 / 8 rows, rings of 2..6 slots and 4 or 8 consumer warps are within 3 % of each other once an
// SM holds ~100 KiB of tiles; requesting the first tiles before griddepcontrol.wait (the database is constant) bought nothing.
// compute-sanitizer: memcheck clean; racecheck flags every bulk-copy write / shared read pair - it does not model completion by
// transaction count (mbarrier complete_tx), which is what orders them.
#include "kernels.cuh"

namespace sb200 {

// ICB = database columns per CTA (256; or the whole shard when it is narrower: 128 / 64 / 32 columns - a database sharded over
// several GPUs, a small second dimension).  The 256 consumer threads are 256 / ICB row groups x ICB columns: a 16 KiB tile is
// R = 1024 / ICB rows, every thread takes 4 of them, and the row groups' partial sums meet in shared memory after the last tile.
template <int ICB, int NS>
__global__ void __launch_bounds__(288) k_scan_spiral_tma(uint32_t *__restrict__ out, const uint64_t *__restrict__ query,
                                                        const uint64_t *__restrict__ db, int dim0, int IC, int zmask) {
    using namespace tc;                                 // mbarrier / bulk-copy wrappers (tc_scan.cu)
    pdl_begin();
    constexpr int R = 1024 / ICB, CG = 256 / ICB, CW = 8;               // rows per tile, row groups, consumer warps (producer: warp 8)
    extern __shared__ __align__(128) uint8_t sm_raw[];
    uint4 *tiles = reinterpret_cast<uint4 *>(sm_raw);                   // [NS][R][ICB] = NS x 16 KiB
    uint4 *qs = tiles + NS * 1024;                                      // [dim0][4]: 64 bytes per (z, j)
    const uint32_t full0 = smem_u32(qs + (size_t)dim0 * 4), empty0 = full0 + 8 * NS, qbar = empty0 + 8 * NS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int z = blockIdx.x, cb = blockIdx.y;
    if (tid == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, CW); }
        mbar_init(qbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int T = dim0 / R;                                             // tiles of this z-slice
    if (warp == CW) {
        if (lane == 0) {
            // zmask = 2047 for an explicit database; an implicit one holds zmask + 1 slices (reference src/spiral.cpp:647)
            const uint4 *dbz = reinterpret_cast<const uint4 *>(db) + ((size_t)(z & zmask) * dim0) * IC + cb * ICB;
            const uint64_t stream = policy_evict_first(), keep = policy_evict_last();     // evict_normal / evict_unchanged: 2 % slower
            const bool contiguous = IC == ICB;                          // the CTA owns whole rows: a tile is one 16 KiB run
            auto issue = [&](int t) {
                const int s = t % NS;
                mbar_expect_tx(full0 + 8 * s, 16384);
                if (contiguous) bulk_g2s(smem_u32(tiles + s * 1024), dbz + (size_t)t * 1024, 16384, full0 + 8 * s, stream);
                else {
#pragma unroll
                    for (int r = 0; r < R; r++)
                        bulk_g2s(smem_u32(tiles + s * 1024 + r * ICB), dbz + (size_t)(t * R + r) * IC, ICB * 16, full0 + 8 * s, stream);
                }
            };
            pdl_wait();
            int t = 0;
            for (; t < NS && t < T; t++) issue(t);
            mbar_expect_tx(qbar, (uint32_t)dim0 * 64);
            bulk_g2s(smem_u32(qs), reinterpret_cast<const uint4 *>(query) + (size_t)z * dim0 * 4, (uint32_t)dim0 * 64, qbar, keep);
            for (; t < T; t++) {
                mbar_wait(empty0 + 8 * (t % NS), ((t / NS) - 1) & 1);   // the eight consumer warps are done with this slot
                issue(t);
            }
        }
        return;
    }
    pdl_wait();
    const int col = tid % ICB, rg = tid / ICB;
    const uint32_t c32p = (uint32_t)((1ull << 32) % kP), c32b = (uint32_t)((1ull << 32) % kB);
    uint64_t acc[3][2];
#pragma unroll
    for (int r = 0; r < 3; r++) acc[r][0] = acc[r][1] = 0;
    mbar_wait(qbar, 0);
    for (int t = 0; t < T; t++) {
        const int s = t % NS;
        mbar_wait(full0 + 8 * s, (t / NS) & 1);
        const uint4 *tile = tiles + s * 1024 + rg * ICB + col;
        const uint4 *qz = qs + ((size_t)t * R + rg) * 4;
#pragma unroll
        for (int k = 0; k < 4; k++) {                                   // this thread's rows of the tile: rg, rg + CG, ...
            const uint4 d = tile[k * CG * ICB];
            const uint4 q0 = qz[k * CG * 4 + 0], q1 = qz[k * CG * 4 + 1], q2 = qz[k * CG * 4 + 2], q3 = qz[k * CG * 4 + 3];
            // d = (m0.p, m0.b, m1.p, m1.b); q0 = m0:(r0.p r0.b r1.p r1.b) q1 = m0:(r2.p r2.b - -) q2,q3 = m1
            acc[0][0] += (uint64_t)q0.x * d.x;  acc[0][1] += (uint64_t)q0.y * d.y;
            acc[1][0] += (uint64_t)q0.z * d.x;  acc[1][1] += (uint64_t)q0.w * d.y;
            acc[2][0] += (uint64_t)q1.x * d.x;  acc[2][1] += (uint64_t)q1.y * d.y;
            acc[0][0] += (uint64_t)q2.x * d.z;  acc[0][1] += (uint64_t)q2.y * d.w;
            acc[1][0] += (uint64_t)q2.z * d.z;  acc[1][1] += (uint64_t)q2.w * d.w;
            acc[2][0] += (uint64_t)q3.x * d.z;  acc[2][1] += (uint64_t)q3.y * d.w;
        }
        // The slot may be refilled the moment the last warp has arrived, by the ASYNC proxy.  ptxas schedules the arrive right
        // behind the last shared load's ISSUE (the MACs that consume it come later), so without this fence the engine overwrote
        // tiles whose loads were still in flight - whole z-slices wrong whenever the producer was waiting for a free slot.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
        if (((t + 1) & 15) == 0) {                                      // every 16 tiles = 64 rows of this thread = 128 products of < 2^56 on top of < 2^61
#pragma unroll
            for (int r = 0; r < 3; r++) {
                acc[r][0] = (acc[r][0] & 0xffffffffull) + (uint64_t)(uint32_t)(acc[r][0] >> 32) * c32p;
                acc[r][1] = (acc[r][1] & 0xffffffffull) + (uint64_t)(uint32_t)(acc[r][1] >> 32) * c32b;
            }
        }
    }
    if (CG > 1) {
        // partial sums of the row groups: folded below 2^61 each (CG <= 8), summed by group 0 through the (now idle) tile ring
#pragma unroll
        for (int r = 0; r < 3; r++) {
            acc[r][0] = (acc[r][0] & 0xffffffffull) + (uint64_t)(uint32_t)(acc[r][0] >> 32) * c32p;
            acc[r][1] = (acc[r][1] & 0xffffffffull) + (uint64_t)(uint32_t)(acc[r][1] >> 32) * c32b;
        }
        uint64_t *ps = reinterpret_cast<uint64_t *>(tiles);            // [CG - 1][6][ICB]
        asm volatile("bar.sync 1, 256;" ::: "memory");                 // every consumer is past its last tile read
        if (rg > 0) {
#pragma unroll
            for (int r = 0; r < 3; r++) { ps[((rg - 1) * 6 + 2 * r) * ICB + col] = acc[r][0]; ps[((rg - 1) * 6 + 2 * r + 1) * ICB + col] = acc[r][1]; }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (rg > 0) return;
        for (int g = 0; g < CG - 1; g++) {
#pragma unroll
            for (int r = 0; r < 3; r++) { acc[r][0] += ps[(g * 6 + 2 * r) * ICB + col]; acc[r][1] += ps[(g * 6 + 2 * r + 1) * ICB + col]; }
        }
    }
    const int ic = cb * ICB + col;
    const int i = ic >> 1, c = ic & 1;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        uint32_t *o = out + ((((size_t)i * kN1 + r) * kN2 + c) * 2) * kN + z;
        o[0] = reduce_u64(acc[r][0], 0);
        o[kN] = reduce_u64(acc[r][1], 1);
    }
}

// dynamic shared memory above 48 KiB is opt-in per kernel; also called when peers are connected, so that no attribute call or lazy
// module load happens while another shard's kernel is spinning on a flag (see preload_exchange_kernels, api.cu)
void scan_tma_prepare() {
    static bool done = false;
    if (done) return;
#define SB200_TMA_ATTR(ICBv) \
    cudaFuncSetAttribute(k_scan_spiral_tma<ICBv, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); \
    cudaFuncSetAttribute(k_scan_spiral_tma<ICBv, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); \
    cudaFuncSetAttribute(k_scan_spiral_tma<ICBv, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    SB200_TMA_ATTR(256) SB200_TMA_ATTR(128) SB200_TMA_ATTR(64) SB200_TMA_ATTR(32)
#undef SB200_TMA_ATTR
    done = true;
}
// returns false when the shape is outside this kernel's domain (the caller then uses k_scan_spiral / k_scan_spiral_jsplit)
bool launch_scan_spiral_tma(uint32_t *out, const uint64_t *query, const uint64_t *db, size_t dim0, size_t num_per, cudaStream_t s, size_t z_slices) {
    static const bool on = [] { const char *e = getenv("SB200_SCAN_TMA"); return !(e && *e == '0'); }();
    const int IC = (int)num_per * 2;
    const int ICB = IC % 256 == 0 ? 256 : IC;           // narrower shards: one CTA owns whole rows
    if (!on || (ICB != 256 && ICB != 128 && ICB != 64 && ICB != 32) || dim0 < 64 || dim0 > 1024 || (dim0 & (dim0 - 1))) return false;
    const int zmask = (int)(z_slices ? z_slices : (size_t)kN) - 1;
    const dim3 grid(kN, IC / ICB);
    // ring depth: as many 16 KiB tiles as fit next to the query slice with 3 (slice <= 16 KiB) or 2 CTAs on an SM
    const size_t slice = dim0 * 64, per_cta = (size_t)227 * 1024 / (slice <= 16384 ? 3 : 2);
    int NS = (int)((per_cta - slice - 1024 - 128) / 16384);
    NS = NS < 3 ? 3 : NS > 5 ? 5 : NS;
    scan_tma_prepare();
    auto go = [&](auto kernel) {
        const size_t smem = (size_t)NS * 16384 + slice + (2 * NS + 1) * 8;
        count_launch();
        launch_pdl_impl(kernel, grid, dim3(288), smem, s, out, query, db, (int)dim0, IC, zmask);
    };
    note_kernel("k_scan_spiral_tma");
#define SB200_TMA_GO(ICBv) { if (NS == 3) go(k_scan_spiral_tma<ICBv, 3>); else if (NS == 4) go(k_scan_spiral_tma<ICBv, 4>); else go(k_scan_spiral_tma<ICBv, 5>); }
    if (ICB == 256) SB200_TMA_GO(256) else if (ICB == 128) SB200_TMA_GO(128) else if (ICB == 64) SB200_TMA_GO(64) else SB200_TMA_GO(32)
#undef SB200_TMA_GO
    return true;
}

}  // namespace sb200
