// pack_kernels.cu - SpiralPack / SpiralStreamPack server path (reference src/testing.cpp): 2x1 Regev
// ciphertexts, 1x1 plaintexts, out_n^2 database planes, packing into one (out_n+1) x out_n response.
//
// Reference functions replaced:
//   convertDb                         src/testing.cpp:316-340  -> k_db_build_pack / transpose from the reference layout
//   reorientCiphertextsDim1           src/testing.cpp:342-362  -> k_reorient_dim1
//   fastMultiplyQueryByDatabaseDim1   src/testing.cpp:364-593  -> k_scan_pack
//   foldCiphertextsDim1               src/testing.cpp:596-624  -> generic fold kernels (spiral_kernels.cu), unsigned digits
//   regevToSimpleGsw + negation       src/testing.cpp:108-140, 1027-1032 -> k_simple_gsw_accum + k_gsw_negate
//   pack                              src/testing.cpp:198-241  -> k_pack_accum
#include "kernels.cuh"
#include "ntt.cuh"

namespace sb200 {

__device__ __forceinline__ uint4 ld_stream_u4p(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint64_t pack_pb3(uint32_t p, uint32_t b) { return (uint64_t)p | ((uint64_t)b << 32); }

// ============================================================================================
// Database.  Scan layout per plane: DBP[z][jp][i][s] PB64, jp = j/2, s = j&1, item = j*num_per + i.
// Consecutive threads (i) read consecutive 16-byte (j even, j odd) pairs.
// ============================================================================================
__global__ void __launch_bounds__(kNttThreads) k_db_build_pack(uint64_t *__restrict__ db, const uint16_t *__restrict__ pts,
                                                               int dim0, int num_per, uint32_t p_db) {
    pdl_prologue();
    extern __shared__ __align__(16) uint32_t dyn[];
    uint32_t(*sm)[kPlaneWords] = reinterpret_cast<uint32_t(*)[kPlaneWords]>(dyn);
    uint32_t *stash = dyn + 2 * kPlaneWords;                       // [s*2 + ipar][n][z]
    const int n = plane_of_thread(), lt = lane_in_plane();
    const uint32_t q = modulus(n);
    const int ih = blockIdx.x, jp = blockIdx.y;                    // i = 2*ih + ipar, j = 2*jp + s
    for (int poly = 0; poly < 4; poly++) {
        const int s = poly >> 1, ipar = poly & 1;
        if (2 * ih + ipar >= num_per) continue;                    // num_per == 1 (a one-column shard): no odd column
        const size_t item = (size_t)(2 * jp + s) * num_per + (2 * ih + ipar);
        const uint16_t *src = pts + item * kN;
        uint32_t v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            uint32_t c = src[nat_pos(lt, k)];                      // centre-lift as src/testing.cpp:857-868
            v[k] = (c >= p_db / 2) ? (q - ((p_db - c) % q)) % q : c % q;
        }
        ntt_forward_plane(v, sm[n], lt, n);
        store_ntt_regs(v, stash + (poly * 2 + n) * kN, lt);
    }
    __syncthreads();
    const size_t JP = dim0 / 2;
    for (int z = threadIdx.x; z < kN; z += kNttThreads) {
        ulonglong2 a, b;        // i even: (s=0, s=1) ; i odd: (s=0, s=1)
        a.x = pack_pb3(stash[(0 * 2 + 0) * kN + z], stash[(0 * 2 + 1) * kN + z]);
        a.y = pack_pb3(stash[(2 * 2 + 0) * kN + z], stash[(2 * 2 + 1) * kN + z]);
        b.x = pack_pb3(stash[(1 * 2 + 0) * kN + z], stash[(1 * 2 + 1) * kN + z]);
        b.y = pack_pb3(stash[(3 * 2 + 0) * kN + z], stash[(3 * 2 + 1) * kN + z]);
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(db) + ((size_t)z * JP + jp) * num_per + 2 * ih;
        dst[0] = a;
        if (num_per > 1) dst[1] = b;
    }
}
void launch_db_build_pack(uint64_t *db_plane, const uint16_t *pts_plane, size_t dim0, size_t num_per, uint32_t p_db, cudaStream_t s) {
    const size_t smem = (2 * kPlaneWords + 4 * 2 * kN) * sizeof(uint32_t);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_db_build_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
    count_launch(); launch_pdl(k_db_build_pack, dim3(dim3((unsigned)((num_per + 1) / 2), (unsigned)(dim0 / 2))), dim3(kNttThreads), smem, s, db_plane, pts_plane, (int)dim0, (int)num_per, p_db);
}

// one item of a plane replaced in place (planting known records in a synthetic database): centre-lift + NTT of its polynomial,
// then its 2048 words go to DBP[z][j/2][i][j&1]
__global__ void __launch_bounds__(kNttThreads) k_db_set_item_pack(uint64_t *__restrict__ db, const uint16_t *__restrict__ poly,
                                                                  int dim0, int num_per, uint32_t p_db, int i, int j) {
    pdl_prologue();
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    __shared__ __align__(16) uint32_t stash[2][kN];
    const int n = plane_of_thread(), lt = lane_in_plane();
    const uint32_t q = modulus(n);
    uint32_t v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t c = poly[nat_pos(lt, k)];
        v[k] = (c >= p_db / 2) ? (q - ((p_db - c) % q)) % q : c % q;
    }
    ntt_forward_plane(v, sm[n], lt, n);
    store_ntt_regs(v, stash[n], lt);
    __syncthreads();
    const size_t JP = dim0 / 2;
    for (int z = threadIdx.x; z < kN; z += kNttThreads)
        db[(((size_t)z * JP + j / 2) * num_per + i) * 2 + (j & 1)] = pack_pb3(stash[0][z], stash[1][z]);
}
void launch_db_set_item_pack(uint64_t *db_plane, const uint16_t *poly, size_t dim0, size_t num_per, uint32_t p_db, size_t i, size_t j, cudaStream_t s) {
    count_launch(); launch_pdl(k_db_set_item_pack, dim3(1), dim3(kNttThreads), 0, s, db_plane, poly, (int)dim0, (int)num_per, p_db, (int)i, (int)j);
}

// ============================================================================================
// reorientCiphertextsDim1: selected 2x1 dev-NTT cts -> query[z][j][r] PB64
// ============================================================================================
__global__ void k_reorient_dim1(uint64_t *__restrict__ out, const uint32_t *__restrict__ cv, const int *__restrict__ ct_idx, int dim0) {
    pdl_prologue();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // (j, z), z fastest
    if (idx >= (size_t)dim0 * kN) return;
    const int z = (int)(idx % kN), j = (int)(idx / kN);
    const uint32_t *ct = cv + (size_t)ct_idx[j] * 2 * 2 * kN;
    ulonglong2 w;
    w.x = pack_pb3(ct[z], ct[kN + z]);
    w.y = pack_pb3(ct[2 * kN + z], ct[3 * kN + z]);
    reinterpret_cast<ulonglong2 *>(out)[(size_t)z * dim0 + j] = w;
}
void launch_reorient_dim1(uint64_t *out, const uint32_t *cv, const int *ct_idx, size_t dim0, cudaStream_t s) {
    const size_t n = dim0 * kN;
    count_launch(); launch_pdl(k_reorient_dim1, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, out, cv, ct_idx, (int)dim0);
}

// ============================================================================================
// First-dimension scan, Pack variant (fastMultiplyQueryByDatabaseDim1):
//   out[plane][i][r][n][z] = sum_j query[z][j][r]_n * DBP[plane][z][j][i]_n      (4 MACs per 8 bytes)
// CTA = 256 threads on one (plane, z): thread = (js, i) with i the fast index; with few i-columns
// (SpiralStreamPack: num_per = 8) the j axis is split over JS = 256/ICT thread groups and the partial
// sums are combined through shared memory, so every lane still streams 16-byte pairs.
// ============================================================================================
constexpr int kPackScanThreads = 256;
// PG planes per CTA: the staged query slice of one z serves PG planes (SpiralStreamPack: the 32 KiB slice is a quarter of the
// 128 KiB of database one (plane, z) streams - re-staging it per plane cost 15 % of the scan)
__global__ void __launch_bounds__(kPackScanThreads) k_scan_pack(uint32_t *__restrict__ out, const uint64_t *__restrict__ query,
                                                                const uint64_t *__restrict__ db, int dim0, int IC, int ICT, int JC,
                                                                size_t plane_words, size_t out_plane_polys, int planes, int PG) {
    pdl_prologue();
    extern __shared__ __align__(16) uint4 qs[];        // [JC pairs][2] uint4 (+ the reduction area behind it when the j axis is split)
    const int tid = threadIdx.x, JS = kPackScanThreads / ICT;
    const int i = tid % ICT, js = tid / ICT;
    const int z = blockIdx.x, i0 = blockIdx.y * ICT, plane0 = blockIdx.z * PG;
    const int JP = dim0 / 2;
    const uint32_t c32p = (uint32_t)((1ull << 32) % kP), c32b = (uint32_t)((1ull << 32) % kB);
    const uint4 *qg = reinterpret_cast<const uint4 *>(query) + (size_t)z * JP * 2;
    const bool whole = JC == JP;                       // the slice fits: staged once for all PG planes
    uint4 *rs = qs + (size_t)JC * 2;                   // [js][i]
    for (int pg = 0; pg < PG && plane0 + pg < planes; pg++) {
        const int plane = plane0 + pg;
        uint64_t acc[2][2] = {{0, 0}, {0, 0}};
        const uint4 *dbz = reinterpret_cast<const uint4 *>(db + plane * plane_words) + ((size_t)z * JP) * IC + i0 + i;
        int since_fold = 0;
        for (int jc0 = 0; jc0 < JP; jc0 += JC) {
            if (!whole || pg == 0) {
                __syncthreads();
                for (int e = tid; e < JC * 2; e += kPackScanThreads) qs[e] = __ldg(qg + (size_t)jc0 * 2 + e);
                __syncthreads();
            }
#pragma unroll 8
            for (int jj = js; jj < JC; jj += JS) {
                const uint4 d = ld_stream_u4p(dbz + (size_t)(jc0 + jj) * IC);      // (j0.p, j0.b, j1.p, j1.b)
                const uint4 q0 = qs[jj * 2], q1 = qs[jj * 2 + 1];                  // j0:(r0.p r0.b r1.p r1.b), j1
                acc[0][0] += (uint64_t)q0.x * d.x;  acc[0][1] += (uint64_t)q0.y * d.y;
                acc[1][0] += (uint64_t)q0.z * d.x;  acc[1][1] += (uint64_t)q0.w * d.y;
                acc[0][0] += (uint64_t)q1.x * d.z;  acc[0][1] += (uint64_t)q1.y * d.w;
                acc[1][0] += (uint64_t)q1.z * d.z;  acc[1][1] += (uint64_t)q1.w * d.w;
                if (++since_fold == 60) {
                    since_fold = 0;
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        acc[r][0] = (acc[r][0] & 0xffffffffull) + (uint64_t)(uint32_t)(acc[r][0] >> 32) * c32p;
                        acc[r][1] = (acc[r][1] & 0xffffffffull) + (uint64_t)(uint32_t)(acc[r][1] >> 32) * c32b;
                    }
                }
            }
        }
        uint32_t red[4] = {reduce_u64(acc[0][0], 0), reduce_u64(acc[0][1], 1), reduce_u64(acc[1][0], 0), reduce_u64(acc[1][1], 1)};
        if (JS > 1) {
            rs[js * ICT + i] = make_uint4(red[0], red[1], red[2], red[3]);
            __syncthreads();
            if (js == 0) {
                uint64_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
                for (int k = 0; k < JS; k++) { const uint4 v = rs[k * ICT + i]; s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w; }
                red[0] = reduce_u64(s0, 0); red[1] = reduce_u64(s1, 1); red[2] = reduce_u64(s2, 0); red[3] = reduce_u64(s3, 1);
            }
            __syncthreads();                               // rs is rewritten by the next plane
        }
        if (js == 0) {
            uint32_t *o = out + ((size_t)plane * out_plane_polys + (size_t)(i0 + i) * 2) * 2 * kN + z;
            o[0] = red[0]; o[kN] = red[1];                // row 0: planes p, b
            o[2 * kN] = red[2]; o[3 * kN] = red[3];       // row 1
        }
    }
}
// Wide planes (SpiralPack: 256 or more columns per z-slice, the whole query slice of a z in <= 32 KiB): the shape is known at
// compile time, so the loop carries no stride multiplies, no fold counter and no j-split bookkeeping - one 128-bit database
// load per column, two shared-memory query reads per pair of columns, 8 MACs per load.  CTA = T threads x U columns on one
// (plane, z), 8 CTAs per SM.  The generic kernel issued 56 % of the scheduler's slots for 45 % FMA-pipe work and ran into the
// power cap on 10 ms scans (profiles/r02_scan_ncu.md); accumulators fold after every 32 pairs (64 products < 2^62 on top of < 2^60).
template <int U, int T, int UNR>
__global__ void __launch_bounds__(T, 1024 / T) k_scan_pack_wide(uint32_t *__restrict__ out, const uint64_t *__restrict__ query,
                                                              const uint64_t *__restrict__ db, int JP, int IC, size_t plane_words, size_t out_plane_polys) {
    pdl_prologue();
    extern __shared__ __align__(16) uint4 qs[];        // [JP][2]
    const int tid = threadIdx.x, z = blockIdx.x, i0 = blockIdx.y * (T * U) + tid, plane = blockIdx.z;
    const uint32_t c32p = (uint32_t)((1ull << 32) % kP), c32b = (uint32_t)((1ull << 32) % kB);
    const uint4 *qg = reinterpret_cast<const uint4 *>(query) + (size_t)z * JP * 2;
    for (int e = tid; e < JP * 2; e += T) qs[e] = __ldg(qg + e);
    __syncthreads();
    const uint4 *dbz = reinterpret_cast<const uint4 *>(db + plane * plane_words) + ((size_t)z * JP) * IC + i0;
    uint64_t acc[U][2][2];
#pragma unroll
    for (int u = 0; u < U; u++) acc[u][0][0] = acc[u][0][1] = acc[u][1][0] = acc[u][1][1] = 0;
    for (int jb = 0; jb < JP; jb += 32) {
        const uint4 *row = dbz + (size_t)jb * IC;
        const uint4 *qb = qs + jb * 2;
#pragma unroll UNR
        for (int jj = 0; jj < 32; jj++) {
            uint4 d[U];
#pragma unroll
            for (int u = 0; u < U; u++) d[u] = ld_stream_u4p(row + (size_t)jj * IC + u * T);      // (j0.p, j0.b, j1.p, j1.b)
            const uint4 q0 = qb[jj * 2], q1 = qb[jj * 2 + 1];                                     // j0:(r0.p r0.b r1.p r1.b), j1
#pragma unroll
            for (int u = 0; u < U; u++) {
                acc[u][0][0] += (uint64_t)q0.x * d[u].x;  acc[u][0][1] += (uint64_t)q0.y * d[u].y;
                acc[u][1][0] += (uint64_t)q0.z * d[u].x;  acc[u][1][1] += (uint64_t)q0.w * d[u].y;
                acc[u][0][0] += (uint64_t)q1.x * d[u].z;  acc[u][0][1] += (uint64_t)q1.y * d[u].w;
                acc[u][1][0] += (uint64_t)q1.z * d[u].z;  acc[u][1][1] += (uint64_t)q1.w * d[u].w;
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++)
#pragma unroll
            for (int r = 0; r < 2; r++) {
                acc[u][r][0] = (acc[u][r][0] & 0xffffffffull) + (uint64_t)(uint32_t)(acc[u][r][0] >> 32) * c32p;
                acc[u][r][1] = (acc[u][r][1] & 0xffffffffull) + (uint64_t)(uint32_t)(acc[u][r][1] >> 32) * c32b;
            }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
        uint32_t *o = out + ((size_t)plane * out_plane_polys + (size_t)(i0 + u * T) * 2) * 2 * kN + z;
        o[0] = reduce_u64(acc[u][0][0], 0); o[kN] = reduce_u64(acc[u][0][1], 1);                  // row 0: planes p, b
        o[2 * kN] = reduce_u64(acc[u][1][0], 0); o[3 * kN] = reduce_u64(acc[u][1][1], 1);         // row 1
    }
}
// Narrow planes (SpiralStreamPack: 8 columns per z-slice, 25 planes): one WARP owns a whole (plane, z) - its 32 lanes cover
// 32 / IC consecutive 16-byte rows of IC columns, i.e. 512 contiguous bytes per load step - and the CTA's warps take different
// planes of the same z, sharing the staged query slice.  No j-split across warps: the partial sums of the 32 / IC row groups
// meet in two shuffles at the very end, so nothing synchronises after the staging and every warp streams its 128 KiB without
// a bubble (the generic kernel reduces 32 partial sums through shared memory and two barriers per plane).
template <int IC, int UNR, int S>                      // S warps share one (plane, z): each takes a contiguous 1/S of its rows
__global__ void __launch_bounds__(512) k_scan_pack_narrow(uint32_t *__restrict__ out, const uint64_t *__restrict__ query, const uint64_t *__restrict__ db,
                                                          int JP, size_t plane_words, size_t out_plane_polys, int planes, int row_stride, int nslab) {
    pdl_prologue();
    extern __shared__ __align__(16) uint4 qs[];        // [JP][2], then (S > 1) the partial sums [warp][IC]
    constexpr int G = 32 / IC;                         // row groups per warp
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, z = blockIdx.x;
    const int W = (int)(blockDim.x >> 5) / S;          // (plane, slab) pairs per CTA
    // planes wider than 32 columns (a SpiralPack plane shared out over 2 or 4 GPUs: 128 / 64 columns) are cut into nslab slabs of
    // IC = 32 columns; a warp then owns one (plane, slab, z) and its rows are row_stride apart instead of back to back
    const int vp = blockIdx.y * W + warp / S, split = warp % S;
    const int plane = vp / nslab, slab = vp % nslab;
    const uint4 *qg = reinterpret_cast<const uint4 *>(query) + (size_t)z * JP * 2;
    for (int e = tid; e < JP * 2; e += blockDim.x) qs[e] = __ldg(qg + e);
    __syncthreads();
    const bool live = vp < planes * nslab;
    const int g = lane / IC, i = lane % IC;
    const uint32_t c32p = (uint32_t)((1ull << 32) % kP), c32b = (uint32_t)((1ull << 32) % kB);
    uint64_t a00 = 0, a01 = 0, a10 = 0, a11 = 0;
    if (live) {
        const uint4 *dbz = reinterpret_cast<const uint4 *>(db + plane * plane_words) + ((size_t)z * JP) * row_stride + slab * IC + g * row_stride + i;   // row jj0 + g, column i
        const int j_begin = split * (JP / S), j_end = j_begin + JP / S;
        for (int jb = j_begin; jb < j_end; jb += G * 32) {
            const uint4 *row = dbz + (size_t)jb * row_stride;
            const uint4 *qb = qs + (jb + g) * 2;
#pragma unroll 1
            for (int k0 = 0; k0 < 32; k0 += UNR) {
                uint4 d[UNR];
#pragma unroll
                for (int u = 0; u < UNR; u++) d[u] = ld_stream_u4p(row + (size_t)(k0 + u) * G * row_stride);     // (j0.p, j0.b, j1.p, j1.b) of row jb + k*G + g
#pragma unroll
                for (int u = 0; u < UNR; u++) {
                    const uint4 q0 = qb[(k0 + u) * G * 2], q1 = qb[(k0 + u) * G * 2 + 1];
                    a00 += (uint64_t)q0.x * d[u].x;  a01 += (uint64_t)q0.y * d[u].y;
                    a10 += (uint64_t)q0.z * d[u].x;  a11 += (uint64_t)q0.w * d[u].y;
                    a00 += (uint64_t)q1.x * d[u].z;  a01 += (uint64_t)q1.y * d[u].w;
                    a10 += (uint64_t)q1.z * d[u].z;  a11 += (uint64_t)q1.w * d[u].w;
                }
            }
            a00 = (a00 & 0xffffffffull) + (uint64_t)(uint32_t)(a00 >> 32) * c32p;  a01 = (a01 & 0xffffffffull) + (uint64_t)(uint32_t)(a01 >> 32) * c32b;
            a10 = (a10 & 0xffffffffull) + (uint64_t)(uint32_t)(a10 >> 32) * c32p;  a11 = (a11 & 0xffffffffull) + (uint64_t)(uint32_t)(a11 >> 32) * c32b;
        }
    }
    uint32_t r0 = reduce_u64(a00, 0), r1 = reduce_u64(a01, 1), r2 = reduce_u64(a10, 0), r3 = reduce_u64(a11, 1);
#pragma unroll
    for (int off = IC; off < 32; off <<= 1) {          // sums of at most 4 residues < 2^30
        r0 += __shfl_xor_sync(0xffffffffu, r0, off); r1 += __shfl_xor_sync(0xffffffffu, r1, off);
        r2 += __shfl_xor_sync(0xffffffffu, r2, off); r3 += __shfl_xor_sync(0xffffffffu, r3, off);
    }
    if (S > 1) {                                       // the other warps of this plane hand their sums over (< 2^30 each, S <= 4)
        uint4 *ps = qs + (size_t)JP * 2;
        if (split != 0 && g == 0) ps[warp * IC + i] = make_uint4(r0, r1, r2, r3);
        __syncthreads();
        if (split == 0 && g == 0) {
#pragma unroll
            for (int k = 1; k < S; k++) { const uint4 v = ps[(warp + k) * IC + i]; r0 += v.x; r1 += v.y; r2 += v.z; r3 += v.w; }
        }
    }
    if (live && split == 0 && g == 0) {
        uint32_t *o = out + ((size_t)plane * out_plane_polys + (size_t)(slab * IC + i) * 2) * 2 * kN + z;
        o[0] = reduce_u64(r0, 0); o[kN] = reduce_u64(r1, 1);                      // row 0: planes p, b
        o[2 * kN] = reduce_u64(r2, 0); o[3 * kN] = reduce_u64(r3, 1);             // row 1
    }
}
void launch_scan_pack(uint32_t *out, const uint64_t *query, const uint64_t *db, size_t dim0, size_t num_per, size_t planes,
                      size_t db_plane_words, size_t out_plane_polys, cudaStream_t s) {
    const int IC = (int)num_per;
    // SB200_PACK_SCAN_SHAPED=0: the generic kernel everywhere (experiments)
    static const bool shaped = [] { const char *e = getenv("SB200_PACK_SCAN_SHAPED"); return !(e && *e == '0'); }();
    const int JPs = (int)dim0 / 2;
    if (shaped && IC % 256 == 0 && JPs % 32 == 0 && (size_t)JPs * 32 <= 32768) {
        // measured at SpiralPack cfg3 (64 GiB): 9.63 ms with 256 threads x 1 column, 9.71 ms with 128 x 2, 11.1 ms generic
        count_launch();
        launch_pdl((k_scan_pack_wide<1, 256, 8>), dim3(kN, IC / 256, (unsigned)planes), dim3(256), (size_t)JPs * 32, s, out, query, db, JPs, IC, db_plane_words, out_plane_polys);
        return;
    }
    const int ICs = IC > 32 ? 32 : IC, nslab = IC / ICs;               // slab width / slabs per plane (narrow kernel)
    if (shaped && (IC == 8 || IC == 16 || IC == 32 || IC == 64 || IC == 128) && JPs % ((32 / ICs) * 32) == 0 && (size_t)JPs * 32 <= 32768) {
        const size_t vplanes = planes * nslab;
        int W = vplanes <= 8 ? (int)vplanes : 8;
        if (vplanes > 8) { int best = 1 << 30; for (int w = 8; w >= 4; w--) { const int waste = (int)(((vplanes + w - 1) / w) * w - vplanes); if (waste < best) { best = waste; W = w; } } }
        // S warps per (plane, slab, z): twice the loads in flight where the rows allow it (the 32 KiB query slice caps an SM at 6 CTAs)
        static const int s_env = [] { const char *e = getenv("SB200_PACK_SCAN_SPLIT"); return e ? atoi(e) : 2; }();
        const int S = (s_env == 2 && JPs % (2 * (32 / ICs) * 32) == 0) ? 2 : 1;
        const dim3 grid(kN, (unsigned)((vplanes + W - 1) / W));
        const size_t smem_n = (size_t)JPs * 32 + (S > 1 ? (size_t)W * S * ICs * 16 : 0);
        static bool attr_n = false;
        if (!attr_n) {
            cudaFuncSetAttribute(k_scan_pack_narrow<8, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            cudaFuncSetAttribute(k_scan_pack_narrow<16, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            cudaFuncSetAttribute(k_scan_pack_narrow<32, 8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            attr_n = true;
        }
        count_launch();
#define SB200_NARROW(ICv, Sv) launch_pdl((k_scan_pack_narrow<ICv, 8, Sv>), grid, dim3(32 * W * Sv), smem_n, s, out, query, db, JPs, db_plane_words, out_plane_polys, (int)planes, IC, nslab)
        if (ICs == 8)       { if (S == 2) SB200_NARROW(8, 2); else SB200_NARROW(8, 1); }
        else if (ICs == 16) { if (S == 2) SB200_NARROW(16, 2); else SB200_NARROW(16, 1); }
        else                { if (S == 2) SB200_NARROW(32, 2); else SB200_NARROW(32, 1); }
#undef SB200_NARROW
        return;
    }
    const int ICT = IC < kPackScanThreads ? IC : kPackScanThreads;
    const int JP = (int)dim0 / 2, JS = kPackScanThreads / ICT;
    int JC = JP;
    while ((size_t)JC * 32 > 32768) JC >>= 1;
    const size_t smem = (size_t)JC * 32 + (JS > 1 ? (size_t)JS * ICT * 16 : 0);
    // planes per CTA: amortise the query staging where it is a visible share of the CTA's traffic (narrow planes), but keep
    // at least ~4 waves of CTAs on the 148 SMs
    int PG = 1;
    if (JC == JP && IC <= 64) { PG = 5; while (PG > 1 && (size_t)kN * (IC / ICT) * ((planes + PG - 1) / PG) < 148 * 7 * 2) PG--; }
    static const int pg_env = [] { const char *e = getenv("SB200_PACK_SCAN_PG"); return e ? atoi(e) : 0; }();
    if (pg_env > 0 && JC == JP) PG = pg_env;
    dim3 grid(kN, IC / ICT, (unsigned)((planes + PG - 1) / PG));
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_scan_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); attr_set = true; }
    count_launch(); launch_pdl(k_scan_pack, dim3(grid), dim3(kPackScanThreads), smem, s, out, query, db, (int)dim0, IC, ICT, JC, db_plane_words, out_plane_polys, (int)planes, PG);
}

// ============================================================================================
// regevToSimpleGsw accumulate:  gsw[d] (2 x 2*ell):  col 2j+1 = c_inp ; col 2j = V (2 x 2*t_conv) * ginv
// ginv: [2*t_conv][nbits] polys (row jj + 2k), bit index b = d*ell + j
// ============================================================================================
__global__ void k_simple_gsw_accum(uint32_t *__restrict__ gsw, const uint32_t *__restrict__ cv, const int *__restrict__ ct_idx,
                                   const uint32_t *__restrict__ ginv, const uint32_t *__restrict__ V, int t_conv, int ell, int nu2) {
    pdl_prologue();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // (b, n, z)
    const int nbits = ell * nu2;
    if (idx >= (size_t)nbits * 2 * kN) return;
    const int nz = (int)(idx % (2 * kN)), b = (int)(idx / (2 * kN)), n = nz >= kN;
    const int d = b / ell, j = b % ell, mc2 = 2 * t_conv, cols = 2 * ell;
    uint64_t acc[2] = {0, 0};
    for (int m = 0; m < mc2; m++) {
        const uint32_t g = ginv[((size_t)m * nbits + b) * 2 * kN + nz];
        acc[0] += (uint64_t)V[((size_t)0 * mc2 + m) * 2 * kN + nz] * g;
        acc[1] += (uint64_t)V[((size_t)1 * mc2 + m) * 2 * kN + nz] * g;
        if ((m & 127) == 127) { acc[0] = reduce_u64(acc[0], n); acc[1] = reduce_u64(acc[1], n); }
    }
    const uint32_t *cin = cv + (size_t)ct_idx[b] * 2 * 2 * kN;
    uint32_t *o = gsw + (size_t)d * 2 * cols * 2 * kN;
#pragma unroll
    for (int r = 0; r < 2; r++) {
        o[((size_t)r * cols + 2 * j) * 2 * kN + nz] = reduce_u64(acc[r], n);
        o[((size_t)r * cols + 2 * j + 1) * 2 * kN + nz] = cin[(size_t)r * 2 * kN + nz];
    }
}
// poly_idx: 2*nbits entries (row-0 polys of the bit ciphertexts, then their row-1 polys)
void launch_regev_to_simple_gsw(uint32_t *gsw_out, uint32_t *gsw_neg_out, const uint32_t *cv, const int *ct_idx, const int *poly_idx,
                                int nu2, int ell, const uint32_t *V, int t_conv, uint64_t *scratch_raw, uint32_t *scratch_ntt, cudaStream_t s) {
    const int nbits = nu2 * ell;
    if (!nbits) return;
    launch_from_ntt_indexed(scratch_raw, cv, poly_idx, 2 * (size_t)nbits, s);             // raw as (rdim = 2) x nbits
    launch_gadget_ntt(scratch_ntt, scratch_raw, 2 * t_conv, 2, nbits, s);
    const size_t n = (size_t)nbits * 2 * kN;
    count_launch(); launch_pdl(k_simple_gsw_accum, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, gsw_out, cv, ct_idx, scratch_ntt, V, t_conv, ell, nu2);
    if (gsw_neg_out) launch_gsw_negate(gsw_neg_out, gsw_out, nu2, ell, 2, s);
}

// ============================================================================================
// pack:  result[row][c] = sum_r ( W_r[row][:] * NTT(G^-1(ct_{r,c} row 0)) ) + [row >= 1] NTT(ct_{row-1,c} row 1)
// ginv: [t_conv][n*n] polys ; ct2: [n*n] polys (NTT of the second rows) ; v_W: [n][(n+1) x t_conv]
// ============================================================================================
__global__ void k_pack_accum(uint32_t *__restrict__ result, const uint32_t *__restrict__ ginv, const uint32_t *__restrict__ ct2,
                             const uint32_t *__restrict__ vW, int out_n, int t_conv) {
    pdl_prologue();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // (row, c, n, z)
    const int rows = out_n + 1, nn = out_n * out_n;
    if (idx >= (size_t)rows * out_n * 2 * kN) return;
    const int nz = (int)(idx % (2 * kN)), rc = (int)(idx / (2 * kN)), n = nz >= kN;
    const int row = rc / out_n, c = rc % out_n;
    uint64_t acc = 0;
    int cnt = 0;
    for (int r = 0; r < out_n; r++) {
        const uint32_t *W = vW + ((size_t)r * rows + row) * t_conv * 2 * kN;
        for (int k = 0; k < t_conv; k++) {
            acc += (uint64_t)W[(size_t)k * 2 * kN + nz] * ginv[((size_t)k * nn + r * out_n + c) * 2 * kN + nz];
            if (++cnt == 128) { cnt = 0; acc = reduce_u64(acc, n); }
        }
    }
    if (row >= 1) acc += ct2[((size_t)(row - 1) * out_n + c) * 2 * kN + nz];
    result[idx] = reduce_u64(acc, n);
}
// v_ct_raw: n*n cts (2 polys each: row 0, row 1) raw ; scratch_raw: 2*n*n polys ; scratch_ntt: (t_conv + 1) * n*n polys
__global__ void k_split_rows(uint64_t *__restrict__ rows01, const uint64_t *__restrict__ cts, int count) {
    pdl_prologue();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // (ct, row, z)
    if (idx >= (size_t)count * 2 * kN) return;
    const int z = (int)(idx % kN), row = (int)((idx / kN) & 1), ct = (int)(idx / (2 * kN));
    rows01[((size_t)row * count + ct) * kN + z] = cts[idx];
}
void launch_pack(uint32_t *result, const uint64_t *v_ct_raw, const uint32_t *vW, int out_n, int t_conv,
                 uint64_t *scratch_raw, uint32_t *scratch_ntt, cudaStream_t s) {
    const int nn = out_n * out_n;
    const size_t n1 = (size_t)nn * 2 * kN;
    count_launch(); launch_pdl(k_split_rows, dim3((unsigned)((n1 + 255) / 256)), dim3(256), 0, s, scratch_raw, v_ct_raw, nn);
    launch_gadget_ntt(scratch_ntt, scratch_raw, t_conv, 1, nn, s);                                   // digits of the first rows
    launch_to_ntt(scratch_ntt + (size_t)t_conv * nn * 2 * kN, scratch_raw + (size_t)nn * kN, nn, s);  // second rows
    const size_t n2 = (size_t)(out_n + 1) * out_n * 2 * kN;
    count_launch(); launch_pdl(k_pack_accum, dim3((unsigned)((n2 + 255) / 256)), dim3(256), 0, s, result, scratch_ntt, scratch_ntt + (size_t)t_conv * nn * 2 * kN, vW, out_n, t_conv);
}

}  // namespace sb200
