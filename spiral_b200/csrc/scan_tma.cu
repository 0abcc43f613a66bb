// scan_tma.cu - first-dimension scan (multiplyQueryByDatabase, src/spiral.cpp:628-999) with the database streamed through
// shared memory by the TMA engine (cp.async.bulk + mbarrier ring) instead of through registers.
//
// Why: k_scan_spiral keeps 24 accumulator registers per thread, so ptxas holds only ~5 of its 8 unrolled 16-byte loads in flight
// and an SM has ~80 KiB outstanding - 6.0-6.5 TB/s, while a read-only stream with 130 KiB outstanding per SM reaches 7.1-7.3 TB/s
// on the same GPU (k_scan_pack_wide, scripts/micro/pipes.cu).  Here the bytes in flight are a ring of NS tiles of R database rows
// (R x 4 KiB) per CTA, filled by one producer lane; the 128 consumer threads read the tiles with conflict-free 128-bit shared
// loads and run the same 12 MACs per 16 bytes.  Same arithmetic, same results, same output layout as k_scan_spiral.
//
//   CTA = one z-slice x 256 database columns: 256 / U consumer threads (thread t owns columns t + u * 256 / U) + 1 producer warp
//   smem = NS x R x 4 KiB tiles | the z-slice of the query (dim0 x 64 B, one bulk copy) | 2 NS + 1 mbarriers
// Measured (cfg1 2 GiB / cfg5 8 GiB, scan stage between CUDA events): k_scan_spiral 0.356 / 1.346 ms; this kernel 0.336 / 1.269 ms
// = 6.39 / 6.77 TB/s.  Tile and ring shapes R x NS in {2,4,8} x {2..6} and U in {1,2} are within 3 % of each other once an SM
// holds ~100 KiB of tiles; requesting the first tiles before griddepcontrol.wait (the database is constant) bought nothing.
#include "kernels.cuh"

namespace sb200 {

template <int R, int NS, int U>                        // U columns per consumer thread: 2 -> 4 consumer warps, 1 -> 8
__global__ void __launch_bounds__(256 / U + 32) k_scan_spiral_tma(uint32_t *__restrict__ out, const uint64_t *__restrict__ query,
                                                        const uint64_t *__restrict__ db, int dim0, int IC, int zmask) {
    using namespace tc;                                 // mbarrier / bulk-copy wrappers (tc_scan.cu)
    pdl_begin();
    extern __shared__ __align__(128) uint8_t sm_raw[];
    uint4 *tiles = reinterpret_cast<uint4 *>(sm_raw);                   // [NS][R][256]
    uint4 *qs = tiles + NS * R * 256;                                   // [dim0][4]: 64 bytes per (z, j)
    const uint32_t full0 = smem_u32(qs + (size_t)dim0 * 4), empty0 = full0 + 8 * NS, qbar = empty0 + 8 * NS;
    constexpr int CT = 256 / U, CW = CT / 32;           // consumer threads / warps; the producer is warp CW
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
            const uint4 *dbz = reinterpret_cast<const uint4 *>(db) + ((size_t)(z & zmask) * dim0) * IC + cb * 256;
            const uint64_t stream = policy_evict_first(), keep = policy_evict_last();     // evict_normal / evict_unchanged: 2 % slower
            auto issue = [&](int t) {
                const int s = t % NS;
                mbar_expect_tx(full0 + 8 * s, R * 4096);
#pragma unroll
                for (int r = 0; r < R; r++)
                    bulk_g2s(smem_u32(tiles + (s * R + r) * 256), dbz + (size_t)(t * R + r) * IC, 4096, full0 + 8 * s, stream);
            };
            pdl_wait();
            int t = 0;
            for (; t < NS && t < T; t++) issue(t);
            mbar_expect_tx(qbar, (uint32_t)dim0 * 64);
            bulk_g2s(smem_u32(qs), reinterpret_cast<const uint4 *>(query) + (size_t)z * dim0 * 4, (uint32_t)dim0 * 64, qbar, keep);
            for (; t < T; t++) {
                mbar_wait(empty0 + 8 * (t % NS), ((t / NS) - 1) & 1);   // the four consumer warps are done with this slot
                issue(t);
            }
        }
        return;
    }
    pdl_wait();
    const uint32_t c32p = (uint32_t)((1ull << 32) % kP), c32b = (uint32_t)((1ull << 32) % kB);
    uint64_t acc[U][3][2];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int r = 0; r < 3; r++) acc[u][r][0] = acc[u][r][1] = 0;
    mbar_wait(qbar, 0);
    for (int t = 0; t < T; t++) {
        const int s = t % NS;
        mbar_wait(full0 + 8 * s, (t / NS) & 1);
        const uint4 *tile = tiles + s * R * 256 + tid;
        const uint4 *qz = qs + (size_t)t * R * 4;
#pragma unroll
        for (int r = 0; r < R; r++) {
            uint4 d[U];
#pragma unroll
            for (int u = 0; u < U; u++) d[u] = tile[r * 256 + u * CT];
            const uint4 q0 = qz[r * 4 + 0], q1 = qz[r * 4 + 1], q2 = qz[r * 4 + 2], q3 = qz[r * 4 + 3];
#pragma unroll
            for (int u = 0; u < U; u++) {
                // d = (m0.p, m0.b, m1.p, m1.b); q0 = m0:(r0.p r0.b r1.p r1.b) q1 = m0:(r2.p r2.b - -) q2,q3 = m1
                acc[u][0][0] += (uint64_t)q0.x * d[u].x;  acc[u][0][1] += (uint64_t)q0.y * d[u].y;
                acc[u][1][0] += (uint64_t)q0.z * d[u].x;  acc[u][1][1] += (uint64_t)q0.w * d[u].y;
                acc[u][2][0] += (uint64_t)q1.x * d[u].x;  acc[u][2][1] += (uint64_t)q1.y * d[u].y;
                acc[u][0][0] += (uint64_t)q2.x * d[u].z;  acc[u][0][1] += (uint64_t)q2.y * d[u].w;
                acc[u][1][0] += (uint64_t)q2.z * d[u].z;  acc[u][1][1] += (uint64_t)q2.w * d[u].w;
                acc[u][2][0] += (uint64_t)q3.x * d[u].z;  acc[u][2][1] += (uint64_t)q3.y * d[u].w;
            }
        }
        // The slot may be refilled the moment the fourth warp has arrived, by the ASYNC proxy.  ptxas schedules the arrive right
        // behind the last shared load's ISSUE (the MACs that consume it come later), so without this fence the engine overwrote
        // tiles whose loads were still in flight - whole z-slices wrong whenever the producer was waiting for a free slot.
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
        if ((((t + 1) * R) & 63) == 0) {                                // every 64 rows = 128 products of < 2^56 on top of < 2^61
#pragma unroll
            for (int u = 0; u < U; u++)
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    acc[u][r][0] = (acc[u][r][0] & 0xffffffffull) + (uint64_t)(uint32_t)(acc[u][r][0] >> 32) * c32p;
                    acc[u][r][1] = (acc[u][r][1] & 0xffffffffull) + (uint64_t)(uint32_t)(acc[u][r][1] >> 32) * c32b;
                }
        }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int ic = cb * 256 + tid + u * CT;
        const int i = ic >> 1, c = ic & 1;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            uint32_t *o = out + ((((size_t)i * kN1 + r) * kN2 + c) * 2) * kN + z;
            o[0] = reduce_u64(acc[u][r][0], 0);
            o[kN] = reduce_u64(acc[u][r][1], 1);
        }
    }
}

// dynamic shared memory above 48 KiB is opt-in per kernel; also called when peers are connected, so that no attribute call or lazy
// module load happens while another shard's kernel is spinning on a flag (see preload_exchange_kernels, api.cu)
void scan_tma_prepare() {
    static bool done = false;
    if (done) return;
    cudaFuncSetAttribute(k_scan_spiral_tma<4, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_scan_spiral_tma<4, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(k_scan_spiral_tma<4, 5, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    done = true;
}
// returns false when the shape is outside this kernel's domain (the caller then uses k_scan_spiral)
bool launch_scan_spiral_tma(uint32_t *out, const uint64_t *query, const uint64_t *db, size_t dim0, size_t num_per, cudaStream_t s, size_t z_slices) {
    static const bool on = [] { const char *e = getenv("SB200_SCAN_TMA"); return !(e && *e == '0'); }();
    const int IC = (int)num_per * 2;
    if (!on || IC % 256 || dim0 < 64 || dim0 > 512 || (dim0 & (dim0 - 1))) return false;
    const int zmask = (int)(z_slices ? z_slices : (size_t)kN) - 1;
    const dim3 grid(kN, IC / 256);
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
    if (NS == 3)      go(k_scan_spiral_tma<4, 3, 1>);
    else if (NS == 4) go(k_scan_spiral_tma<4, 4, 1>);
    else              go(k_scan_spiral_tma<4, 5, 1>);
    return true;
}

}  // namespace sb200
