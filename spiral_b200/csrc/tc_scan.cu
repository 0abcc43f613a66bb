// tc_scan.cu - batched first dimension on the 5th-generation tensor cores (north_star "batched-query
// variant ... dense contraction on int8-limb tcgen05 tiles").
//
// Same function as k_scan_spiral (multiplyQueryByDatabase, reference src/spiral.cpp:628-999), for up to 16 queries
// that share ONE pass over the database:
//     out_q[i][r][c][n][z] = sum_{k=(j,m)} query_q[z][k][r]_n * DB[z][k][ic]_n   mod prime_n        (ic = 2i + c)
// For a fixed (z, n) this is a (IC x K) . (K x 3*BQ) integer matrix product with 28-bit operands.  Every residue is
// split into four unsigned byte limbs (the bytes of the little-endian u32 the database already stores), the 16 limb
// products run as exact u8 x u8 -> s32 tcgen05.mma (kind::i8) with accumulators in tensor memory, and the epilogue
// recombines  sum_w 2^(8w) T_w  modulo the prime (T_w = sum of the limb products of weight w = a + b; < 2^28, exact).
//
// Layouts (chosen so that the copy engine, not threads, moves every byte):
//   database  "DB_tc"  : for item it = (z, n, mt):  [kc][a][128 rows (ic)][128 B of k]  - 16 KiB tiles stored in HBM
//                        EXACTLY as the shared-memory image the MMA wants (K-major, SWIZZLE_128B), in consumption
//                        order, so one item is one contiguous 32*KC KiB run fetched by cp.async.bulk (UBLKCP).
//                        Same 8 bytes per NTT coefficient as the scan layout: the four limbs ARE the four bytes.
//   queries   "Q_tc"   : for (z, n): [kc][4*NB rows][128 B of k], row = b*NB + 3*q + r (b = query limb), same swizzle.
//   accumulators       : TMEM columns [w*NB, (w+1)*NB), w = 0..6.  The MMA for database limb a writes the window
//                        [a*NB, a*NB + 4*NB): query limb b lands on weight a + b, so equal weights add up in place.
// CTA = 2 + NB/4 warps, persistent, one per SM: warp 0 producer (bulk copies into an 8 x 16 KiB A ring and a 3-deep B ring),
// warp 1 MMA issuer (one thread) + TMEM allocation, then four epilogue warps per 16 accumulator columns
// (TMEM -> registers, accumulators released at once, -> Barrett -> global).
#include "kernels.cuh"
#include "common.cuh"

namespace sb200 {
namespace tc {

constexpr int kM = 128;                    // database columns per tile (UMMA M)
constexpr int kKB = 128;                   // bytes of k per tile row = one SWIZZLE_128B atom = 4 MMA k-steps
constexpr int kATile = kM * kKB;           // 16 KiB
constexpr int kAStages = 8;
constexpr int kBStages = 3;
constexpr int kMaxBatch = 16;

struct OutPtrs { uint32_t *out[kMaxBatch]; int count; };

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// (no labels inside the asm: a loop unrolled by the compiler holds several copies of this block in one function, and branch
// targets named alike - even in separate { } scopes - mis-synchronised k_scan_spiral_tma whenever its ring depth was a power of two)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
// global -> shared bulk copy (the TMA engine, no tensor map: source tiles are already the shared-memory image)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {     // arrives on `bar` when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
    return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
           ((uint64_t)2 << 61);
}
// instruction descriptor: u8 x u8 -> s32, both operands K-major, M = 128 (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t idesc_i8(int n) { return (2u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kM >> 4) << 24); }
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int NB> struct Shape {
    static constexpr int kBRows = 4 * NB;
    static constexpr int kBTile = kBRows * kKB;
    static constexpr int kCols = 7 * NB;
    static constexpr int kAlloc = kCols <= 128 ? 128 : kCols <= 256 ? 256 : 512;
    static constexpr size_t kSmem = 1024 + (size_t)kAStages * kATile + (size_t)kBStages * kBTile + 256;
};

// byte offset of (row, 16-byte chunk) inside a SWIZZLE_128B tile whose base is 1024-byte aligned
__host__ __device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t chunk) { return row * 128u + ((chunk ^ (row & 7u)) << 4); }

// sum_w 2^(8w) T_w  mod q from the seven weight accumulators of one output (T_w < 2^28): the powers 2^32, 2^40, 2^48 are
// replaced by their residues so the whole sum stays below 2^59, then one Barrett step with floor(2^59/q) on the top 32 bits.
// Results go to the tile-order buffer T1[q][g][r][ic], g = (plane*2048 + z)*2 + n: the 32 lanes of a warp own 32 consecutive ic, so every store
// instruction writes one full 128-byte line (the reference order [i][r][c][n][z] would be 32 separate 4-byte sectors).
template <int CB, int RQ>
__device__ __forceinline__ void epilogue_store(uint32_t *__restrict__ t1, int count, uint32_t IC, uint32_t sq, const uint32_t (&v)[7][16], uint32_t off, int n) {
    constexpr uint32_t c4p = (uint32_t)((1ull << 32) % kP), c5p = (uint32_t)((1ull << 40) % kP), c6p = (uint32_t)((1ull << 48) % kP);
    constexpr uint32_t c4b = (uint32_t)((1ull << 32) % kB), c5b = (uint32_t)((1ull << 40) % kB), c6b = (uint32_t)((1ull << 48) % kB);
    constexpr uint32_t mup = (uint32_t)((1ull << 59) / kP), mub = (uint32_t)((1ull << 59) / kB);
    const uint32_t c4 = n ? c4b : c4p, c5 = n ? c5b : c5p, c6 = n ? c6b : c6p, mu = n ? mub : mup, q = n ? kB : kP;
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const int col = CB * 16 + e, qi = col / RQ, r = col - RQ * qi;
        if (qi < count) {
            uint64_t x = (uint64_t)v[0][e] + ((uint64_t)v[1][e] << 8) + ((uint64_t)v[2][e] << 16) + ((uint64_t)v[3][e] << 24);
            x += (uint64_t)v[4][e] * c4 + (uint64_t)v[5][e] * c5 + (uint64_t)v[6][e] * c6;
            const uint32_t qhat = __umulhi((uint32_t)(x >> 27), mu);
            uint32_t res = (uint32_t)x - qhat * q;                  // true remainder + at most 3q  (< 2^30)
            res = min(res, res - 2 * q);
            res = min(res, res - q);
            t1[(size_t)(off + (uint32_t)qi * sq + (uint32_t)r * IC)] = res;     // host checks that T1 has < 2^32 words
        }
    }
}

// T1[q][g][r][ic] -> dev-NTT outputs: 32 x 32 shared-memory transpose, 128-byte lines both ways.
//   Spiral (RQ = 3): out_q[i][r][c][n][z] with ic = 2i + c           (multiplyQueryByDatabase's output order)
//   Pack   (RQ = 2): out_q[plane][i = ic][r][n][z]                   (fastMultiplyQueryByDatabaseDim1's 2x1 MatPolys)
template <int RQ>
__global__ void __launch_bounds__(256) k_tc_untile(const __grid_constant__ OutPtrs outs, const uint32_t *__restrict__ t1, int IC, int planes,
                                                   size_t out_plane_polys) {
    pdl_prologue();
    __shared__ uint32_t tile[32][33];
    const int z0 = blockIdx.x * 32, ic0 = blockIdx.y * 32;
    int bz = blockIdx.z;
    const int r = bz % RQ; bz /= RQ;
    const int n = bz & 1; bz >>= 1;
    const int plane = bz % planes, q = bz / planes;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const size_t sq = (size_t)planes * 2 * kN * RQ * IC;
    const uint32_t *src = t1 + (size_t)q * sq + ((((size_t)plane * kN + z0) * 2 + n) * RQ + r) * IC + ic0;
#pragma unroll
    for (int k = 0; k < 4; k++) tile[ty + 8 * k][tx] = __ldg(src + (size_t)(ty + 8 * k) * 2 * RQ * IC + tx);
    __syncthreads();
    uint32_t *dst = outs.out[q];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int ic = ic0 + ty + 8 * k;
        size_t poly;
        if (RQ == 3) { const int i = ic >> 1, c = ic & 1; poly = ((size_t)i * kN1 + r) * kN2 + c; }
        else         poly = (size_t)plane * out_plane_polys + (size_t)ic * 2 + r;
        dst[(poly * 2 + n) * kN + z0 + tx] = tile[tx][ty + 8 * k];
    }
}

template <int NB, int RQ>
__global__ void __launch_bounds__(64 + 8 * NB, 1) k_scan_tc(uint32_t *__restrict__ t1, int count, const uint8_t *__restrict__ q_tc,
                                                         const uint8_t *__restrict__ db_tc, int KC, int MT, int n_items) {
    using S = Shape<NB>;
    pdl_prologue();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base, b_smem = base + kAStages * kATile;
    const uint32_t bars = b_smem + kBStages * S::kBTile;
    // barrier map (8 bytes each): a_full[8] a_empty[8] b_full[3] b_empty[3] tmem_full tmem_empty ; then the TMEM base address
    const uint32_t a_full = bars, a_empty = bars + 8 * kAStages, b_full = bars + 16 * kAStages, b_empty = b_full + 8 * kBStages;
    const uint32_t t_full = b_empty + 8 * kBStages, t_empty = t_full + 8, tmem_slot = t_empty + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kAStages; i++) { mbar_init(a_full + 8 * i, 1); mbar_init(a_empty + 8 * i, 1); }
        for (int i = 0; i < kBStages; i++) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
        mbar_init(t_full, 1); mbar_init(t_empty, 8 * NB);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(S::kAlloc) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot));

    if (warp == 0) {                                                // ===== producer (lane 0 issues, the warp stays together) =====
        const uint64_t pol_db = policy_evict_first(), pol_q = policy_evict_last();
        uint32_t as = 0, aph = 0, bs = 0, bph = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            if (lane == 0) {
                const uint8_t *asrc = db_tc + (size_t)it * KC * 4 * kATile;
                const uint8_t *bsrc = q_tc + (size_t)((it / MT) % (2 * kN)) * KC * S::kBTile;   // every plane scans the same query tiles
                for (int kc = 0; kc < KC; kc++) {
                    mbar_wait(b_empty + 8 * bs, bph ^ 1);
                    mbar_expect_tx(b_full + 8 * bs, S::kBTile);
                    bulk_g2s(b_smem + bs * S::kBTile, bsrc + (size_t)kc * S::kBTile, S::kBTile, b_full + 8 * bs, pol_q);
                    if (++bs == kBStages) { bs = 0; bph ^= 1; }
                    for (int a = 0; a < 4; a++) {
                        mbar_wait(a_empty + 8 * as, aph ^ 1);
                        mbar_expect_tx(a_full + 8 * as, kATile);
                        bulk_g2s(a_smem + as * kATile, asrc + (size_t)(kc * 4 + a) * kATile, kATile, a_full + 8 * as, pol_db);
                        if (++as == kAStages) { as = 0; aph ^= 1; }
                    }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {                                         // ===== MMA issuer (one thread) =====
        uint32_t as = 0, aph = 0, bs = 0, bph = 0, tph = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            if (lane == 0) {
                mbar_wait(t_empty, tph ^ 1);                        // epilogue has drained the previous item's accumulators
                tc_fence_after();
                for (int kc = 0; kc < KC; kc++) {
                    mbar_wait(b_full + 8 * bs, bph);
                    const uint32_t b_addr = b_smem + bs * S::kBTile;
                    for (int a = 0; a < 4; a++) {
                        mbar_wait(a_full + 8 * as, aph);
                        tc_fence_after();
                        const uint32_t a_addr = a_smem + as * kATile;
#pragma unroll
                        for (int ks = 0; ks < 4; ks++) {
                            const uint64_t ad = smem_desc_sw128(a_addr + ks * 32), bd = smem_desc_sw128(b_addr + ks * 32);
                            const bool first = (kc == 0 && ks == 0);
                            if (a == 0) {
                                mma_i8(tmem, ad, bd, idesc_i8(4 * NB), first ? 0u : 1u);
                            } else if (!first) {
                                mma_i8(tmem + a * NB, ad, bd, idesc_i8(4 * NB), 1u);
                            } else {                                // weight block a+3 sees its first product here: overwrite it
                                mma_i8(tmem + a * NB, ad, bd, idesc_i8(3 * NB), 1u);
                                mma_i8(tmem + a * NB + 3 * NB, ad, smem_desc_sw128(b_addr + 3 * NB * kKB + ks * 32), idesc_i8(NB), 0u);
                            }
                        }
                        tc_commit(a_empty + 8 * as);
                        if (++as == kAStages) { as = 0; aph ^= 1; }
                    }
                    tc_commit(b_empty + 8 * bs);
                    if (++bs == kBStages) { bs = 0; bph ^= 1; }
                }
                tc_commit(t_full);
                tph ^= 1;
            }
            __syncwarp();
        }
    } else {                                                        // ===== epilogue: 4 warps per 16-column chunk =====
        const int e = warp - 2, quarter = warp & 3, cb = e >> 2;   // TMEM lanes are tied to warp % 4; chunk cb of NB/16
        const int row = quarter * 32 + lane;
        uint32_t tph = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const int mt = it % MT, g = it / MT, n = g & 1;          // g = (plane*2048 + z)*2 + n
            const uint32_t IC = (uint32_t)MT * kM;
            const uint32_t sq = (uint32_t)(n_items / MT) * RQ * IC;  // T1 words per query
            const uint32_t off = ((uint32_t)g * RQ) * IC + (uint32_t)(mt * kM + row);   // T1 word index of (q = 0, r = 0)
            mbar_wait(t_full, tph);
            tc_fence_after();
            const uint32_t trow = tmem + ((uint32_t)(quarter * 32) << 16);
            uint32_t v[7][16];
#pragma unroll
            for (int w = 0; w < 7; w++) tmem_ld16(trow + w * NB + cb * 16, v[w]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(t_empty);                                   // accumulators are in registers: the next item may start
            if (cb == 0)      epilogue_store<0, RQ>(t1, count, IC, sq, v, off, n);
            else if (cb == 1) epilogue_store<1, RQ>(t1, count, IC, sq, v, off, n);
            else              epilogue_store<2, RQ>(t1, count, IC, sq, v, off, n);
            tph ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(S::kAlloc) : "memory");
    }
}

// ---- database: scan layout DB'[z][j][ic][m] PB64 -> DB_tc.  One thread per (z, kc, ic, 16-byte chunk): 8 j x 2 m.
__global__ void __launch_bounds__(256) k_db_to_tc(uint8_t *__restrict__ db_tc, const uint4 *__restrict__ db, int dim0, int IC) {
    const int KC = dim0 * 2 / kKB, MT = IC / kM;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;           // ((z*KC + kc)*8 + chunk)*IC + ic
    if (idx >= (size_t)kN * KC * 8 * IC) return;
    const int ic = (int)(idx % IC);
    const int chunk = (int)((idx / IC) & 7);
    const int kc = (int)((idx / IC / 8) % KC);
    const int z = (int)(idx / IC / 8 / KC);
    uint4 d[8];
#pragma unroll
    for (int jj = 0; jj < 8; jj++) d[jj] = db[((size_t)z * dim0 + kc * 64 + chunk * 8 + jj) * IC + ic];
    const int mt = ic / kM, row = ic % kM;
#pragma unroll
    for (int n = 0; n < 2; n++)
#pragma unroll
        for (int a = 0; a < 4; a++) {
            uint32_t w[4];
#pragma unroll
            for (int g = 0; g < 4; g++) {                                        // bytes 4g .. 4g+3 = (j = 2g, m = 0,1), (j = 2g+1, m = 0,1)
                const uint32_t e0 = n ? d[2 * g].y : d[2 * g].x, e1 = n ? d[2 * g].w : d[2 * g].z;
                const uint32_t e2 = n ? d[2 * g + 1].y : d[2 * g + 1].x, e3 = n ? d[2 * g + 1].w : d[2 * g + 1].z;
                w[g] = ((e0 >> (8 * a)) & 0xff) | (((e1 >> (8 * a)) & 0xff) << 8) | (((e2 >> (8 * a)) & 0xff) << 16) | (((e3 >> (8 * a)) & 0xff) << 24);
            }
            const size_t item = ((size_t)z * 2 + n) * MT + mt;
            uint8_t *tile = db_tc + ((item * KC + kc) * 4 + a) * (size_t)kATile;
            *reinterpret_cast<uint4 *>(tile + sw128(row, chunk)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
}

// ---- queries -> their RQ x 4 rows of every Q_tc tile.  Source: PB64 words query[z][k][kstride] with the RQ ciphertext rows
// first (Spiral: reorientCiphertexts layout [z][j][m][4], k = 2j + m, kstride 4;  Pack: reorientCiphertextsDim1 layout
// [z][j][2], k = j, kstride 2).  One thread per (z, kc, 16-byte chunk, r): 16 k's under both primes -> 8 chunks (n, b).
struct QueryPtrs { const uint64_t *query[kMaxBatch]; int first_slot; };
template <int NB>
__global__ void __launch_bounds__(256) k_query_to_tc(uint8_t *__restrict__ q_tc, const __grid_constant__ QueryPtrs qp, int K, int kstride, int RQ) {
    using S = Shape<NB>;
    pdl_prologue();
    const int KC = K / kKB;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;           // ((z*KC + kc)*8 + chunk)*RQ + r
    if (idx >= (size_t)kN * KC * 8 * RQ) return;
    const int q = qp.first_slot + blockIdx.y;
    const uint64_t *__restrict__ query = qp.query[blockIdx.y];
    const int r = (int)(idx % RQ), chunk = (int)((idx / RQ) & 7), kc = (int)((idx / RQ / 8) % KC), z = (int)(idx / RQ / 8 / KC);
    uint64_t e[16];
    const uint64_t *src = query + ((size_t)z * K + kc * kKB + chunk * 16) * kstride + r;
#pragma unroll
    for (int t = 0; t < 16; t++) e[t] = __ldg(src + (size_t)t * kstride);
#pragma unroll
    for (int n = 0; n < 2; n++)
#pragma unroll
        for (int b = 0; b < 4; b++) {
            uint32_t w[4];
#pragma unroll
            for (int g = 0; g < 4; g++) {
                uint32_t acc = 0;
#pragma unroll
                for (int t = 0; t < 4; t++) acc |= (uint32_t)((e[4 * g + t] >> (32 * n + 8 * b)) & 0xff) << (8 * t);
                w[g] = acc;
            }
            uint8_t *tile = q_tc + (((size_t)z * 2 + n) * KC + kc) * (size_t)S::kBTile;
            *reinterpret_cast<uint4 *>(tile + sw128(b * NB + RQ * q + r, chunk)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
}

inline int nb_for(int rows) { return rows <= 16 ? 16 : rows <= 32 ? 32 : 48; }      // accumulator block width for `rows` = RQ * capacity

}  // namespace tc

// ---- geometry of one tensor-core scan: K bytes of k per row, IC database columns, RQ ciphertext rows per query
TcGeom tc_geom_spiral(size_t dim0, size_t num_per) { return TcGeom{dim0 * 2, num_per * 2, 3, 1, 0, 4}; }
TcGeom tc_geom_pack(size_t dim0, size_t num_per, size_t planes, size_t out_plane_polys) { return TcGeom{dim0, num_per, 2, planes, out_plane_polys, 2}; }
int tc_geom_ok(const TcGeom &g) { return g.K >= (size_t)tc::kKB && g.K % tc::kKB == 0 && g.IC >= (size_t)tc::kM && g.IC % tc::kM == 0 && g.planes >= 1; }
size_t tc_db_bytes_g(const TcGeom &g) { return g.planes * (size_t)kN * 2 * g.IC * g.K * 4; }
size_t tc_query_bytes_g(const TcGeom &g, int capacity) { return (size_t)kN * 2 * (g.K / tc::kKB) * 4 * tc::nb_for(g.RQ * capacity) * tc::kKB; }
size_t tc_scratch_bytes_g(const TcGeom &g, int count) { return (size_t)count * g.planes * 2 * kN * g.RQ * g.IC * sizeof(uint32_t); }

// src_plane_words: u64 words between consecutive planes of the scan-layout source (16-byte pairs of consecutive k per column)
void launch_db_to_tc_g(uint8_t *db_tc, const uint64_t *db, const TcGeom &g, size_t src_plane_words, cudaStream_t s) {
    const size_t n = (size_t)kN * (g.K / tc::kKB) * 8 * g.IC, plane_bytes = (size_t)kN * 2 * g.IC * g.K * 4;
    for (size_t p = 0; p < g.planes; p++) {
        count_launch();
        note_kernel("k_db_to_tc"); tc::k_db_to_tc<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(db_tc + p * plane_bytes, reinterpret_cast<const uint4 *>(db + p * src_plane_words),
                                                                   (int)(g.K / 2), (int)g.IC);
    }
}

// capacity fixes the tile shape (NB); queries[b] goes to batch slot first_slot + b (< capacity); one launch for all of them
void launch_queries_to_tc_g(uint8_t *q_tc, const uint64_t *const *queries, int count, int first_slot, int capacity, const TcGeom &g, cudaStream_t s) {
    const size_t n = (size_t)kN * (g.K / tc::kKB) * 8 * g.RQ;
    const dim3 grid((unsigned)((n + 255) / 256), (unsigned)count);
    tc::QueryPtrs qp;
    for (int b = 0; b < tc::kMaxBatch; b++) qp.query[b] = b < count ? queries[b] : nullptr;
    qp.first_slot = first_slot;
    count_launch();
    switch (tc::nb_for(g.RQ * capacity)) {
        case 16: launch_pdl(tc::k_query_to_tc<16>, grid, dim3(256), 0, s, q_tc, qp, (int)g.K, (int)g.kstride, g.RQ); break;
        case 32: launch_pdl(tc::k_query_to_tc<32>, grid, dim3(256), 0, s, q_tc, qp, (int)g.K, (int)g.kstride, g.RQ); break;
        default: launch_pdl(tc::k_query_to_tc<48>, grid, dim3(256), 0, s, q_tc, qp, (int)g.K, (int)g.kstride, g.RQ); break;
    }
}

template <int NB, int RQ>
static int launch_scan_tc_nb(uint32_t *t1, int count, const uint8_t *q_tc, const uint8_t *db_tc, int KC, int MT, size_t groups, cudaStream_t s) {
    static int sms = 0;
    static bool attr = false;
    if (!attr) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (cudaFuncSetAttribute(tc::k_scan_tc<NB, RQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::Shape<NB>::kSmem) != cudaSuccess) return -2;
        attr = true;
    }
    const size_t items = groups * MT;
    if (items > 0x7fffffffull) return -1;
    const int n_items = (int)items;
    const int grid = n_items < sms ? n_items : sms;
    count_launch();
    launch_pdl((tc::k_scan_tc<NB, RQ>), dim3(grid), dim3(64 + 8 * NB), tc::Shape<NB>::kSmem, s, t1, count, q_tc, db_tc, KC, MT, n_items);
    return 0;
}

// out[q]: device outputs in the order of the single-query scans (launch_scan_spiral / launch_scan_pack); count <= capacity <= 16;
// q_tc built with the same capacity; scratch: tc_scratch_bytes_g(g, count) bytes for the tile-order results before the transpose
int launch_scan_tc_g(uint32_t *const *out, int count, int capacity, const uint8_t *q_tc, const uint8_t *db_tc, const TcGeom &g,
                     uint32_t *scratch, cudaStream_t s) {
    if (count < 1 || count > capacity || capacity > tc::kMaxBatch || !tc_geom_ok(g) || !scratch) return -1;
    if (tc_scratch_bytes_g(g, count) / 4 > 0xffffffffull) return -1;          // the kernels index T1 with 32-bit words
    tc::OutPtrs o;
    for (int b = 0; b < tc::kMaxBatch; b++) o.out[b] = b < count ? out[b] : nullptr;
    o.count = count;
    const int KC = (int)(g.K / tc::kKB), MT = (int)(g.IC / tc::kM);
    const size_t groups = g.planes * 2 * kN;
    int rc;
    if (g.RQ == 3) switch (tc::nb_for(3 * capacity)) {
        case 16: rc = launch_scan_tc_nb<16, 3>(scratch, count, q_tc, db_tc, KC, MT, groups, s); break;
        case 32: rc = launch_scan_tc_nb<32, 3>(scratch, count, q_tc, db_tc, KC, MT, groups, s); break;
        default: rc = launch_scan_tc_nb<48, 3>(scratch, count, q_tc, db_tc, KC, MT, groups, s); break;
    } else switch (tc::nb_for(2 * capacity)) {
        case 16: rc = launch_scan_tc_nb<16, 2>(scratch, count, q_tc, db_tc, KC, MT, groups, s); break;
        default: rc = launch_scan_tc_nb<32, 2>(scratch, count, q_tc, db_tc, KC, MT, groups, s); break;
    }
    if (rc) return rc;
    count_launch();
    const dim3 grid(kN / 32, (unsigned)(g.IC / 32), (unsigned)(count * g.planes * 2 * g.RQ));
    if (g.RQ == 3) launch_pdl(tc::k_tc_untile<3>, grid, dim3(256), 0, s, o, (const uint32_t *)scratch, (int)g.IC, (int)g.planes, g.out_plane_polys);
    else           launch_pdl(tc::k_tc_untile<2>, grid, dim3(256), 0, s, o, (const uint32_t *)scratch, (int)g.IC, (int)g.planes, g.out_plane_polys);
    return 0;
}

// ---- Spiral-shaped wrappers (dim0, num_per)
int tc_shape_ok(size_t dim0, size_t num_per) { return tc_geom_ok(tc_geom_spiral(dim0, num_per)); }
size_t tc_query_bytes(size_t dim0, int capacity) { return tc_query_bytes_g(tc_geom_spiral(dim0, 64), capacity); }
size_t tc_scratch_bytes(size_t num_per, int count) { return tc_scratch_bytes_g(tc_geom_spiral(64, num_per), count); }
void launch_db_to_tc(uint8_t *db_tc, const uint64_t *db, size_t dim0, size_t num_per, cudaStream_t s) {
    launch_db_to_tc_g(db_tc, db, tc_geom_spiral(dim0, num_per), 0, s);
}
void launch_queries_to_tc(uint8_t *q_tc, const uint64_t *const *queries, int count, int first_slot, int capacity, size_t dim0, cudaStream_t s) {
    launch_queries_to_tc_g(q_tc, queries, count, first_slot, capacity, tc_geom_spiral(dim0, 64), s);
}
void launch_query_to_tc(uint8_t *q_tc, const uint64_t *query, int q, int capacity, size_t dim0, cudaStream_t s) {
    launch_queries_to_tc(q_tc, &query, 1, q, capacity, dim0, s);
}
int launch_scan_tc(uint32_t *const *out, int count, int capacity, const uint8_t *q_tc, const uint8_t *db_tc, size_t dim0, size_t num_per,
                   uint32_t *scratch, cudaStream_t s) {
    return launch_scan_tc_g(out, count, capacity, q_tc, db_tc, tc_geom_spiral(dim0, num_per), scratch, s);
}

}  // namespace sb200
