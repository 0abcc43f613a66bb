// spiral_kernels.cu - database preprocessing, first-dimension scan, folding, query expansion and
// Regev->GSW conversion kernels of the Spiral (matrix-Regev, n = 2) server path.
//
// Reference functions replaced (paths relative to the reference tree):
//   load_db                       src/spiral.cpp:1028-1172   -> k_db_build_spiral / k_db_from_reference
//   reorientCiphertexts           src/spiral.cpp:410-433     -> k_reorient_query (and fused in k_scal_to_mat_accum)
//   multiplyQueryByDatabase       src/spiral.cpp:628-999     -> k_scan_spiral
//   nttInvAndCrtLiftCiphertexts   src/spiral.cpp:437-453     -> k_from_ntt (ntt_kernels.cu)
//   foldOneFurtherDimension       src/spiral.cpp:1349-1410   -> k_fold_decomp_ntt + k_fold_mac_intt
//   expandImproved                src/spiral.cpp:1664-1743   -> k_expand_prep + k_expand_digits + k_expand_accum
//   scalToMat / regevToGSW        src/spiral.cpp:1850-2025   -> k_from_ntt_indexed + k_gadget_ntt + *_accum
//   GSW negation                  src/spiral.cpp:2361-2378   -> k_gsw_negate
#include "kernels.cuh"
#include "ntt.cuh"

namespace sb200 {

__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint64_t pack_pb(uint32_t p, uint32_t b) { return (uint64_t)p | ((uint64_t)b << 32); }

// ============================================================================================
// Database preprocessing.
// Scan layout: DB'[z][j][ic][m] PB64, ic = ii*n2 + c, database item i = j*num_per + ii, plaintext
// entry (m, c).  One (z, j) row is IC*16 bytes; consecutive threads of the scan read consecutive
// 16-byte (m=0, m=1) pairs -> fully coalesced 128-bit loads.
// ============================================================================================
constexpr int kDbStashWords = 4 * 2 * kN;   // 4 polys x 2 planes

__global__ void __launch_bounds__(kNttThreads) k_db_build_spiral(uint64_t *__restrict__ db, const uint16_t *__restrict__ pts,
                                                                 int nu1, int nu2, uint32_t p_db, size_t item_begin) {
    pdl_prologue();
    extern __shared__ __align__(16) uint32_t dyn[];
    uint32_t(*sm)[kPlaneWords] = reinterpret_cast<uint32_t(*)[kPlaneWords]>(dyn);
    uint32_t *stash = dyn + 2 * kPlaneWords;                       // [poly][n][z]
    const int n = plane_of_thread(), lt = lane_in_plane();
    const uint32_t q = modulus(n);
    const size_t item = item_begin + blockIdx.x;
    const size_t num_per = (size_t)1 << nu2, dim0 = (size_t)1 << nu1;
    const uint16_t *src = pts + (size_t)blockIdx.x * 4 * kN;
    for (int poly = 0; poly < 4; poly++) {
        uint32_t v[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            // centre-lift: v >= p_db/2 -> v - p_db (+Q)   (reference src/spiral.cpp:1116-1127); Q = 0 mod p, b
            uint32_t c = src[poly * kN + nat_pos(lt, k)];
            v[k] = (c >= p_db / 2) ? (q - ((p_db - c) % q)) % q : c % q;
        }
        ntt_forward_plane(v, sm[n], lt, n);
        store_ntt_regs(v, stash + (poly * 2 + n) * kN, lt);
    }
    __syncthreads();
    const size_t ii = item % num_per, j = item / num_per, IC = num_per * 2;
    for (int z = threadIdx.x; z < kN; z += kNttThreads) {
        // 32 contiguous bytes: (c=0: m=0, m=1), (c=1: m=0, m=1); poly index = m*n2 + c
        ulonglong2 c0, c1;
        c0.x = pack_pb(stash[(0 * 2 + 0) * kN + z], stash[(0 * 2 + 1) * kN + z]);   // m=0,c=0
        c0.y = pack_pb(stash[(2 * 2 + 0) * kN + z], stash[(2 * 2 + 1) * kN + z]);   // m=1,c=0
        c1.x = pack_pb(stash[(1 * 2 + 0) * kN + z], stash[(1 * 2 + 1) * kN + z]);   // m=0,c=1
        c1.y = pack_pb(stash[(3 * 2 + 0) * kN + z], stash[(3 * 2 + 1) * kN + z]);   // m=1,c=1
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(db + (((size_t)z * dim0 + j) * IC + ii * 2) * 2);
        dst[0] = c0;
        dst[1] = c1;
    }
}
void launch_db_build_spiral(uint64_t *db, const uint16_t *pts, int nu1, int nu2, uint32_t p_db,
                            size_t item_begin, size_t item_count, cudaStream_t s) {
    const size_t smem = (2 * kPlaneWords + kDbStashWords) * sizeof(uint32_t);
    static bool attr_set = false;
    if (!attr_set) { cudaFuncSetAttribute(k_db_build_spiral, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr_set = true; }
    if (item_count) { count_launch(); launch_pdl(k_db_build_spiral, dim3((unsigned)item_count), dim3(kNttThreads), smem, s, db, pts, nu1, nu2, p_db, item_begin); }
}

// reference layout B[z][ii][c][j][m] -> DB'[z][j][ic][m]: per z a (IC x dim0) -> (dim0 x IC) transpose of 16-byte elements
__global__ void k_db_from_reference(uint64_t *__restrict__ db, const uint64_t *__restrict__ Bref, size_t dim0, size_t IC, size_t z_begin) {
    pdl_prologue();
    // per z: src is (IC rows x dim0 cols) of 16-byte elements, dst is (dim0 x IC)
    __shared__ ulonglong2 tile[32][33];
    const size_t z = blockIdx.z;      // z relative to the chunk for Bref, absolute z_begin + z for db
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(Bref) + z * IC * dim0;
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(db) + (z_begin + z) * IC * dim0;
    const size_t j0 = (size_t)blockIdx.x * 32, ic0 = (size_t)blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        size_t ic = ic0 + r, j = j0 + threadIdx.x;
        if (ic < IC && j < dim0) tile[r][threadIdx.x] = src[ic * dim0 + j];
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        size_t j = j0 + r, ic = ic0 + threadIdx.x;
        if (ic < IC && j < dim0) dst[j * IC + ic] = tile[threadIdx.x][r];
    }
}
void launch_db_from_reference(uint64_t *db, const uint64_t *B_ref, size_t dim0, size_t ic, size_t z_begin,
                              size_t z_count, cudaStream_t s) {
    if (!z_count) return;
    dim3 grid((unsigned)((dim0 + 31) / 32), (unsigned)((ic + 31) / 32), (unsigned)z_count);
    count_launch(); launch_pdl(k_db_from_reference, dim3(grid), dim3(dim3(32, 8)), 0, s, db, B_ref, dim0, ic, z_begin);
}

void launch_db_to_reference(uint64_t *B_ref, const uint64_t *db, size_t dim0, size_t ic, cudaStream_t s) {
    // the same 16-byte transpose with the roles of (ic, j) swapped
    dim3 grid((unsigned)((ic + 31) / 32), (unsigned)((dim0 + 31) / 32), (unsigned)kN);
    count_launch(); launch_pdl(k_db_from_reference, dim3(grid), dim3(dim3(32, 8)), 0, s, B_ref, db, ic, dim0, 0);
}

// ============================================================================================
// reorientCiphertexts: dev-NTT cts [j][r][m][n][z] -> query[z][j][m][4] PB64 (r = 3 lane zero)
// ============================================================================================
__global__ void k_reorient_query(uint64_t *__restrict__ out, const uint32_t *__restrict__ cts, size_t dim0) {
    pdl_prologue();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // (jm, z), z fastest
    if (idx >= dim0 * 2 * kN) return;
    const size_t z = idx % kN, jm = idx / kN, j = jm >> 1, m = jm & 1;
    uint64_t w[4];
#pragma unroll
    for (int r = 0; r < 3; r++) {
        const uint32_t *p = cts + (((j * kN1 + r) * 2 + m) * 2) * (size_t)kN;
        w[r] = pack_pb(p[z], p[kN + z]);
    }
    w[3] = 0;
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out + ((z * dim0 + j) * 2 + m) * 4);
    dst[0] = make_ulonglong2(w[0], w[1]);
    dst[1] = make_ulonglong2(w[2], w[3]);
}
void launch_reorient_query(uint64_t *out, const uint32_t *cts, size_t dim0, cudaStream_t s) {
    size_t n = dim0 * 2 * kN;
    count_launch(); launch_pdl(k_reorient_query, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, out, cts, dim0);
}

// ============================================================================================
// First-dimension scan  (multiplyQueryByDatabase, src/spiral.cpp:628-999)
//   out[i][r][c][n][z] = sum_{j,m} query[z][j][m][r]_n * DB[z][i][c][j][m]_n  mod prime_n
// HBM-bound: every database byte is read exactly once per query; 6 32x32->64 MACs per 8 bytes.
// CTA = 128 threads covering ZT z-slices x ICT ic-columns; each thread owns U ic-columns of one z
// and streams its 16-byte (m=0,m=1) pairs down the j axis with 128-bit no-allocate loads, while
// the query slice for those z sits in shared memory and is broadcast to the warp.
// Accumulators are unreduced u64 (products < 2^56); they are folded to < 2^61 every 64 j
// (128 products) with one 32x32 multiply, and Barrett-reduced once at the end.
// ============================================================================================
constexpr int kScanFoldEvery = 64;

__device__ __forceinline__ uint64_t fold_acc(uint64_t a, uint32_t c32) {      // a mod q preserved, result < 2^61
    return (a & 0xffffffffull) + (uint64_t)(uint32_t)(a >> 32) * c32;
}

template <int U, int kScanThreads, int UNR, bool kFullTile>      // kFullTile: ICT == U * kScanThreads (one z-slice per CTA), strides fold into immediates
__global__ void __launch_bounds__(kScanThreads) k_scan_spiral(uint32_t *__restrict__ out, const uint64_t *__restrict__ query,
                                                             const uint64_t *__restrict__ db, int dim0, int IC, int ICT,
                                                             int ZT, int JC, int zmask) {
    pdl_prologue();
    extern __shared__ __align__(16) uint4 qs[];        // [ZT][JC][4] uint4 = 64 bytes per (z, j)
    const int tid = threadIdx.x;
    const int TZ = kFullTile ? kScanThreads : ICT / U; // threads per z-slice; thread owns columns icl + u*TZ
    const int zl = tid / TZ, icl = tid % TZ;
    const int z0 = blockIdx.x * ZT, z = z0 + zl;
    const int ic0 = blockIdx.y * ICT + icl;            // first owned column
    const bool active = tid < ZT * TZ;                 // tiny IC: surplus threads only help staging
    const uint32_t c32p = (uint32_t)((1ull << 32) % kP), c32b = (uint32_t)((1ull << 32) % kB);

    uint64_t acc[U][3][2];
#pragma unroll
    for (int u = 0; u < U; u++)
#pragma unroll
        for (int r = 0; r < 3; r++) acc[u][r][0] = acc[u][r][1] = 0;

    // zmask = 2047 for an explicit database; an implicit (--random-data) one holds only zmask + 1 slices and NTT coefficient z
    // reads slice z mod (zmask + 1)  (reference src/spiral.cpp:647)
    const uint4 *dbz = reinterpret_cast<const uint4 *>(db) + ((size_t)(z & zmask) * dim0) * IC + ic0;
    const uint4 *qg = reinterpret_cast<const uint4 *>(query);

    for (int jc0 = 0; jc0 < dim0; jc0 += JC) {
        __syncthreads();
        // stage the query rows [z0 .. z0+ZT) x [jc0 .. jc0+JC) : 4 uint4 per (z, j)
        for (int e = tid; e < ZT * JC * 4; e += kScanThreads) {
            const int zz = e / (JC * 4), rem = e % (JC * 4);
            qs[e] = __ldg(qg + ((size_t)(z0 + zz) * dim0 + jc0) * 4 + rem);
        }
        __syncthreads();
        const uint4 *qz = qs + (size_t)zl * JC * 4;
        if (active)
#pragma unroll UNR
        for (int jj = 0; jj < JC; jj++) {
            const int j = jc0 + jj;
            uint4 d[U];
#pragma unroll
            for (int u = 0; u < U; u++) d[u] = ld_stream_u4(dbz + (size_t)j * IC + u * TZ);
            const uint4 q0 = qz[jj * 4 + 0], q1 = qz[jj * 4 + 1], q2 = qz[jj * 4 + 2], q3 = qz[jj * 4 + 3];
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
            if ((jj & (kScanFoldEvery - 1)) == kScanFoldEvery - 1) {
#pragma unroll
                for (int u = 0; u < U; u++)
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        acc[u][r][0] = fold_acc(acc[u][r][0], c32p);
                        acc[u][r][1] = fold_acc(acc[u][r][1], c32b);
                    }
            }
        }
    }
    if (active)
#pragma unroll
    for (int u = 0; u < U; u++) {
        const int ic = ic0 + u * TZ;
        const int i = ic >> 1, c = ic & 1;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            uint32_t *o = out + ((((size_t)i * kN1 + r) * kN2 + c) * 2) * kN + z;
            o[0] = reduce_u64(acc[u][r][0], 0);
            o[kN] = reduce_u64(acc[u][r][1], 1);
        }
    }
}


// Narrow shards (16..64 database columns per z-slice: a database sharded over many GPUs, or a small second dimension).  With two
// columns per thread a z-slice occupies only IC/2 threads, the whole grid a quarter (or less) of the full-tile scan's threads, and
// the loads in flight no longer cover the HBM latency (4.4-4.6 TB/s).  Here the j axis of a z-slice is split over JS thread groups
// of the same CTA (interleaved rows, so the groups together still stream consecutive 1 KiB rows), which restores 128 threads per
// z-slice; the JS partial sums (folded below 2^61 each) meet in shared memory and group 0 reduces and stores.
//   thread = js * (ZT*TZ) + zl * TZ + icl,  TZ = IC/2 threads per z-slice,  JS * ZT * TZ = 128
__global__ void __launch_bounds__(128) k_scan_spiral_jsplit(uint32_t *__restrict__ out, const uint64_t *__restrict__ query,
                                                            const uint64_t *__restrict__ db, int dim0, int IC, int ZT, int JS, int JC, int zmask) {
    pdl_prologue();
    extern __shared__ __align__(16) uint4 qs[];        // [ZT][JC][4] uint4, later reused for the reduction
    const int tid = threadIdx.x;
    const int TZ = IC >> 1, per = ZT * TZ;
    const int js = tid / per, loc = tid % per, zl = loc / TZ, icl = loc % TZ;
    const int z0 = blockIdx.x * ZT, z = z0 + zl;
    const uint32_t c32p = (uint32_t)((1ull << 32) % kP), c32b = (uint32_t)((1ull << 32) % kB);
    uint64_t acc[2][3][2];
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
        for (int r = 0; r < 3; r++) acc[u][r][0] = acc[u][r][1] = 0;
    const uint4 *dbz = reinterpret_cast<const uint4 *>(db) + ((size_t)(z & zmask) * dim0) * IC + icl;
    const uint4 *qg = reinterpret_cast<const uint4 *>(query);
    int it = 0;
    for (int jc0 = 0; jc0 < dim0; jc0 += JC) {
        __syncthreads();
        for (int e = tid; e < ZT * JC * 4; e += 128) {
            const int zz = e / (JC * 4), rem = e % (JC * 4);
            qs[e] = __ldg(qg + ((size_t)(z0 + zz) * dim0 + jc0) * 4 + rem);
        }
        __syncthreads();
        const uint4 *qz = qs + (size_t)zl * JC * 4;
#pragma unroll 4
        for (int jj = js; jj < JC; jj += JS, it++) {
            const size_t row = (size_t)(jc0 + jj) * IC;
            const uint4 d0 = ld_stream_u4(dbz + row), d1 = ld_stream_u4(dbz + row + TZ);
            const uint4 q0 = qz[jj * 4 + 0], q1 = qz[jj * 4 + 1], q2 = qz[jj * 4 + 2], q3 = qz[jj * 4 + 3];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const uint4 d = u == 0 ? d0 : d1;
                acc[u][0][0] += (uint64_t)q0.x * d.x;  acc[u][0][1] += (uint64_t)q0.y * d.y;
                acc[u][1][0] += (uint64_t)q0.z * d.x;  acc[u][1][1] += (uint64_t)q0.w * d.y;
                acc[u][2][0] += (uint64_t)q1.x * d.x;  acc[u][2][1] += (uint64_t)q1.y * d.y;
                acc[u][0][0] += (uint64_t)q2.x * d.z;  acc[u][0][1] += (uint64_t)q2.y * d.w;
                acc[u][1][0] += (uint64_t)q2.z * d.z;  acc[u][1][1] += (uint64_t)q2.w * d.w;
                acc[u][2][0] += (uint64_t)q3.x * d.z;  acc[u][2][1] += (uint64_t)q3.y * d.w;
            }
            if ((it & (kScanFoldEvery - 1)) == kScanFoldEvery - 1) {
#pragma unroll
                for (int u = 0; u < 2; u++)
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        acc[u][r][0] = fold_acc(acc[u][r][0], c32p);
                        acc[u][r][1] = fold_acc(acc[u][r][1], c32b);
                    }
            }
        }
    }
    // partial sums of the JS groups -> shared memory (each folded below 2^61, so up to 4 of them add without overflow)
    __syncthreads();
    uint64_t *red = reinterpret_cast<uint64_t *>(qs);   // [js][12][per]
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
        for (int r = 0; r < 3; r++) {
            red[((size_t)js * 12 + (u * 3 + r) * 2 + 0) * per + loc] = fold_acc(acc[u][r][0], c32p);
            red[((size_t)js * 12 + (u * 3 + r) * 2 + 1) * per + loc] = fold_acc(acc[u][r][1], c32b);
        }
    __syncthreads();
    if (js == 0) {
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int ic = icl + u * TZ;
            const int i = ic >> 1, c = ic & 1;
#pragma unroll
            for (int r = 0; r < 3; r++) {
                uint64_t a0 = 0, a1 = 0;
                for (int g = 0; g < JS; g++) {
                    a0 += red[((size_t)g * 12 + (u * 3 + r) * 2 + 0) * per + loc];
                    a1 += red[((size_t)g * 12 + (u * 3 + r) * 2 + 1) * per + loc];
                }
                uint32_t *o = out + ((((size_t)i * kN1 + r) * kN2 + c) * 2) * kN + z;
                o[0] = reduce_u64(a0, 0);
                o[kN] = reduce_u64(a1, 1);
            }
        }
    }
}
// The same with the shard's shape as compile-time constants (IC = 16 / 32 / 64 columns, JS j-groups, ZT z-slices per CTA) and eight
// resident CTAs per SM.  ncu of the run-time-shaped kernel at 64 columns (profiles/r02_scan_ncu.md): 72 registers -> 7 CTAs per SM,
// 103 warp instructions per 32 bytes of database against 49 in the full-tile scan (64-bit row-address multiplies, run-time strides),
// issue slots 44 % busy at 4.9 TB/s - an integer-overhead problem on top of the HBM stream, not a memory one.
template <int IC, int JS, int ZT>
__global__ void __launch_bounds__(128, 8) k_scan_spiral_jsplit_t(uint32_t *__restrict__ out, const uint64_t *__restrict__ query,
                                                                 const uint64_t *__restrict__ db, int dim0, int JC, int zmask) {
    pdl_prologue();
    extern __shared__ __align__(16) uint4 qs[];        // [ZT][JC][4] uint4, later reused for the reduction
    constexpr int TZ = IC / 2, per = ZT * TZ;
    static_assert(JS * per == 128, "one CTA = 128 threads");
    const int tid = threadIdx.x;
    const int js = tid / per, loc = tid % per, zl = loc / TZ, icl = loc % TZ;
    const int z0 = blockIdx.x * ZT, z = z0 + zl;
    const uint32_t c32p = (uint32_t)((1ull << 32) % kP), c32b = (uint32_t)((1ull << 32) % kB);
    uint64_t acc[2][3][2];
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
        for (int r = 0; r < 3; r++) acc[u][r][0] = acc[u][r][1] = 0;
    const uint4 *dbp = reinterpret_cast<const uint4 *>(db) + ((size_t)(z & zmask) * dim0 + js) * IC + icl;   // row js of this z-slice
    const uint4 *qg = reinterpret_cast<const uint4 *>(query);
    int it = 0;
    for (int jc0 = 0; jc0 < dim0; jc0 += JC) {
        __syncthreads();
        for (int e = tid; e < ZT * JC * 4; e += 128) {
            const int zz = e / (JC * 4), rem = e % (JC * 4);
            qs[e] = __ldg(qg + ((size_t)(z0 + zz) * dim0 + jc0) * 4 + rem);
        }
        __syncthreads();
        const uint4 *qz = qs + (size_t)zl * JC * 4 + js * 4;
#pragma unroll 4
        for (int jj = js; jj < JC; jj += JS, it++, dbp += JS * IC, qz += JS * 4) {
            const uint4 d0 = ld_stream_u4(dbp), d1 = ld_stream_u4(dbp + TZ);
            const uint4 q0 = qz[0], q1 = qz[1], q2 = qz[2], q3 = qz[3];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const uint4 d = u == 0 ? d0 : d1;
                acc[u][0][0] += (uint64_t)q0.x * d.x;  acc[u][0][1] += (uint64_t)q0.y * d.y;
                acc[u][1][0] += (uint64_t)q0.z * d.x;  acc[u][1][1] += (uint64_t)q0.w * d.y;
                acc[u][2][0] += (uint64_t)q1.x * d.x;  acc[u][2][1] += (uint64_t)q1.y * d.y;
                acc[u][0][0] += (uint64_t)q2.x * d.z;  acc[u][0][1] += (uint64_t)q2.y * d.w;
                acc[u][1][0] += (uint64_t)q2.z * d.z;  acc[u][1][1] += (uint64_t)q2.w * d.w;
                acc[u][2][0] += (uint64_t)q3.x * d.z;  acc[u][2][1] += (uint64_t)q3.y * d.w;
            }
            if ((it & (kScanFoldEvery - 1)) == kScanFoldEvery - 1) {
#pragma unroll
                for (int u = 0; u < 2; u++)
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        acc[u][r][0] = fold_acc(acc[u][r][0], c32p);
                        acc[u][r][1] = fold_acc(acc[u][r][1], c32b);
                    }
            }
        }
    }
    __syncthreads();
    uint64_t *red = reinterpret_cast<uint64_t *>(qs);   // [js][12][per]
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
        for (int r = 0; r < 3; r++) {
            red[((size_t)js * 12 + (u * 3 + r) * 2 + 0) * per + loc] = fold_acc(acc[u][r][0], c32p);
            red[((size_t)js * 12 + (u * 3 + r) * 2 + 1) * per + loc] = fold_acc(acc[u][r][1], c32b);
        }
    __syncthreads();
    // every thread finishes 12 / JS of its column pair's 12 sums (the run-time kernel left the whole epilogue to group 0)
    for (int o = js; o < 12; o += JS) {
        uint64_t a = 0;
#pragma unroll
        for (int g = 0; g < JS; g++) a += red[((size_t)g * 12 + o) * per + loc];
        const int u = o / 6, r = (o % 6) / 2, n = o & 1;
        const int ic = icl + u * TZ, i = ic >> 1, c = ic & 1;
        out[((((size_t)i * kN1 + r) * kN2 + c) * 2 + n) * kN + z] = reduce_u64(a, n);
    }
}
void launch_scan_spiral(uint32_t *out, const uint64_t *query, const uint64_t *db, size_t dim0, size_t num_per, cudaStream_t s, size_t z_slices) {
    if (launch_scan_spiral_tma(out, query, db, dim0, num_per, s, z_slices)) return;     // wide shards: database tiles through the TMA ring
    const int zmask = (int)(z_slices ? z_slices : (size_t)kN) - 1;
    // Tiling measured on B200 at cfg1 (profiles/r01_kernel_times_warm.md): 128 threads x 2 columns, unroll 4 is the
    // best of {64x4, 128x2, 256x1} x {unroll 2, 4, 8}; capping residency to make the 2048 CTAs an exact two waves is slower
    // (8 CTAs per SM: 0.360 ms, 7: 0.390, 6: 0.414 - round 2, profiles/r02_expansion_chains.md).  One column per thread at 44
    // registers (11 CTAs per SM, more loads in flight) is within 1 % of this shape at 2 GiB and 8 GiB: at 6.5 TB/s the scan's
    // 0.75 IMAD.WIDE per byte keep the multiplier pipe 56 % busy (30 IMAD.WIDE per SM per clock measured, scripts/micro/pipes.cu).
    // The query slice is staged in chunks of at most 16 KiB per CTA: with 32 KiB (first dimensions of 512 and 1024) only 6 CTAs
    // fit an SM and the scan drops from 6.45 to 5.6-6.0 TB/s (profiles/r01_scan_shapes.md).
    // Narrow shards (IC = 64 or 128 columns: a small second dimension, or a database sharded over many GPUs) keep two columns per
    // thread - one broadcast read of the query slice per two database loads - by giving each z-slice IC/2 threads (a whole number
    // of warps) and putting 2 or 4 z-slices into the CTA; below 64 columns one column per thread remains.
    const int IC = (int)num_per * 2;
    static const bool jsplit = [] { const char *e = getenv("SB200_SCAN_JSPLIT"); return !(e && *e == '0'); }();
    if (jsplit && IC >= 16 && IC <= 64 && dim0 >= 4) {
        const int TZ = IC / 2;
        int JS = 128 / TZ, ZT = 1;
        if (JS > 4) { ZT = JS / 4; JS = 4; }              // at most 4 partial sums per output; the rest of the CTA takes more z-slices
        while (JS > (int)dim0) { JS >>= 1; ZT <<= 1; }
        int JC = (int)dim0;
        while ((size_t)ZT * JC * 64 > 16384 && JC > kScanFoldEvery) JC >>= 1;
        const size_t smem = std::max((size_t)ZT * JC * 64, (size_t)128 * 12 * 8);
        count_launch();
        static const bool fixed = [] { const char *e = getenv("SB200_SCAN_JSPLIT_FIXED"); return !(e && *e == '0'); }();
        if (fixed && JS == 4 && JC % 4 == 0) {             // the common shapes (first dimension >= 4) with compile-time strides
            if (IC == 64 && ZT == 1) { launch_pdl((k_scan_spiral_jsplit_t<64, 4, 1>), dim3(kN), dim3(128), smem, s, out, query, db, (int)dim0, JC, zmask); return; }
            if (IC == 32 && ZT == 2) { launch_pdl((k_scan_spiral_jsplit_t<32, 4, 2>), dim3(kN / 2), dim3(128), smem, s, out, query, db, (int)dim0, JC, zmask); return; }
            if (IC == 16 && ZT == 4) { launch_pdl((k_scan_spiral_jsplit_t<16, 4, 4>), dim3(kN / 4), dim3(128), smem, s, out, query, db, (int)dim0, JC, zmask); return; }
        }
        launch_pdl(k_scan_spiral_jsplit, dim3(kN / ZT), dim3(128), smem, s, out, query, db, (int)dim0, IC, ZT, JS, JC, zmask);
        return;
    }
    static const bool t64 = [] { const char *e = getenv("SB200_SCAN_T64"); return !(e && *e == '0'); }();
    const int T = (t64 && IC == 64) ? 64 : 128;        // 64-column shards: smaller CTAs spread more evenly over the 148 SMs
    const int U = IC >= 64 ? 2 : 1;
    const int ICT = IC < T * U ? IC : T * U;
    int ZT = (T * U) / ICT;
    if (ZT > 8) ZT = 8;
    static const size_t smem_cap = [] { const char *e = getenv("SB200_SCAN_SMEM"); size_t v = e ? (size_t)atol(e) : 0; return v >= 4096 ? v : (size_t)16384; }();
    int JC = (int)dim0;
    while ((size_t)ZT * JC * 64 > smem_cap && JC > kScanFoldEvery) JC >>= 1;
    const size_t smem = (size_t)ZT * JC * 64;
    dim3 grid(kN / ZT, IC / ICT);
    if (JC < (int)dim0) note_kernel("k_scan_spiral[query slice staged in chunks]");     // visible in sb200_kernel_log
    count_launch();
    if (U == 2 && ICT == 2 * T) launch_pdl((k_scan_spiral<2, 128, 4, true>), grid, dim3(128), smem, s, out, query, db, (int)dim0, IC, ICT, ZT, JC, zmask);
    else if (U == 2 && T == 64) launch_pdl((k_scan_spiral<2, 64, 4, false>), grid, dim3(64), smem, s, out, query, db, (int)dim0, IC, ICT, ZT, JC, zmask);
    else if (U == 2)            launch_pdl((k_scan_spiral<2, 128, 4, false>), grid, dim3(128), smem, s, out, query, db, (int)dim0, IC, ICT, ZT, JC, zmask);
    else                        launch_pdl((k_scan_spiral<1, 128, 4, false>), grid, dim3(128), smem, s, out, query, db, (int)dim0, IC, ICT, ZT, JC, zmask);
}

// ---- batched first dimension: BQ queries answered in ONE pass over the database ------------------------------
// (SURVEY 8f, rank 1).  The single-query scan uses ~30 % of the FMA pipe at full HBM bandwidth, so several queries can
// share every 16-byte database load: per load 12*BQ MACs against BQ query slices in shared memory.  CTA = 128 threads,
// one column each, on one (z, 128-column tile); accumulators BQ x 6 u64 per thread.
struct ScanBatchArgs {
    const uint64_t *query[8];
    uint32_t *out[8];
};
template <int BQ>
__global__ void __launch_bounds__(128) k_scan_spiral_batched(const __grid_constant__ ScanBatchArgs a, const uint64_t *__restrict__ db,
                                                             int dim0, int IC, int JC) {
    pdl_prologue();
    extern __shared__ __align__(16) uint4 qs[];        // [BQ][JC][4]
    const int tid = threadIdx.x, z = blockIdx.x, ic = blockIdx.y * 128 + tid;
    const uint32_t c32p = (uint32_t)((1ull << 32) % kP), c32b = (uint32_t)((1ull << 32) % kB);
    uint64_t acc[BQ][3][2];
#pragma unroll
    for (int b = 0; b < BQ; b++)
#pragma unroll
        for (int r = 0; r < 3; r++) acc[b][r][0] = acc[b][r][1] = 0;
    const uint4 *dbz = reinterpret_cast<const uint4 *>(db) + ((size_t)z * dim0) * IC + ic;
    for (int jc0 = 0; jc0 < dim0; jc0 += JC) {
        __syncthreads();
#pragma unroll
        for (int b = 0; b < BQ; b++) {
            const uint4 *qg = reinterpret_cast<const uint4 *>(a.query[b]) + ((size_t)z * dim0 + jc0) * 4;
            for (int e = tid; e < JC * 4; e += 128) qs[(size_t)b * JC * 4 + e] = __ldg(qg + e);
        }
        __syncthreads();
#pragma unroll 2
        for (int jj = 0; jj < JC; jj++) {
            const uint4 d = ld_stream_u4(dbz + (size_t)(jc0 + jj) * IC);
#pragma unroll
            for (int b = 0; b < BQ; b++) {
                const uint4 *q = qs + ((size_t)b * JC + jj) * 4;
                const uint4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
                acc[b][0][0] += (uint64_t)q0.x * d.x;  acc[b][0][1] += (uint64_t)q0.y * d.y;
                acc[b][1][0] += (uint64_t)q0.z * d.x;  acc[b][1][1] += (uint64_t)q0.w * d.y;
                acc[b][2][0] += (uint64_t)q1.x * d.x;  acc[b][2][1] += (uint64_t)q1.y * d.y;
                acc[b][0][0] += (uint64_t)q2.x * d.z;  acc[b][0][1] += (uint64_t)q2.y * d.w;
                acc[b][1][0] += (uint64_t)q2.z * d.z;  acc[b][1][1] += (uint64_t)q2.w * d.w;
                acc[b][2][0] += (uint64_t)q3.x * d.z;  acc[b][2][1] += (uint64_t)q3.y * d.w;
            }
            if ((jj & (kScanFoldEvery - 1)) == kScanFoldEvery - 1) {
#pragma unroll
                for (int b = 0; b < BQ; b++)
#pragma unroll
                    for (int r = 0; r < 3; r++) {
                        acc[b][r][0] = fold_acc(acc[b][r][0], c32p);
                        acc[b][r][1] = fold_acc(acc[b][r][1], c32b);
                    }
            }
        }
    }
    const int i = ic >> 1, c = ic & 1;
#pragma unroll
    for (int b = 0; b < BQ; b++)
#pragma unroll
        for (int r = 0; r < 3; r++) {
            uint32_t *o = a.out[b] + ((((size_t)i * kN1 + r) * kN2 + c) * 2) * kN + z;
            o[0] = reduce_u64(acc[b][r][0], 0);
            o[kN] = reduce_u64(acc[b][r][1], 1);
        }
}
// count in {2, 4}; requires IC = 2 * num_per to be a multiple of 128
int launch_scan_spiral_batched(uint32_t *const *out, const uint64_t *const *query, int count, const uint64_t *db, size_t dim0,
                               size_t num_per, cudaStream_t s) {
    const int IC = (int)num_per * 2;
    if ((count != 2 && count != 4) || IC % 128) return -1;
    ScanBatchArgs a;
    for (int b = 0; b < count; b++) { a.query[b] = query[b]; a.out[b] = out[b]; }
    int JC = (int)dim0;
    while ((size_t)count * JC * 64 > 65536 && JC > kScanFoldEvery) JC >>= 1;
    const size_t smem = (size_t)count * JC * 64;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_scan_spiral_batched<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        cudaFuncSetAttribute(k_scan_spiral_batched<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        attr_set = true;
    }
    dim3 grid(kN, IC / 128);
    count_launch();
    if (count == 2) launch_pdl((k_scan_spiral_batched<2>), grid, dim3(128), smem, s, a, db, (int)dim0, IC, JC);
    else            launch_pdl((k_scan_spiral_batched<4>), grid, dim3(128), smem, s, a, db, (int)dim0, IC, JC);
    return 0;
}

// ============================================================================================
// Folding  (foldOneFurtherDimension, src/spiral.cpp:1349-1410)
//   C'[i] = Qneg (x) G^-1(C[i]) + Q (x) G^-1(C[num_per + i])
// k_fold_decomp_ntt : signed base-2^bits_per digits with the reference's per-half carry chain
//                     (split_and_crt, :270-341) -> forward NTT -> scratch[ct][m][c]
// k_fold_mac_intt   : 2*m2-term pointwise dot product (cpu_mul_query_by_ct, :464-582), add,
//                     inverse NTT and CRT lift (:1374-1407) fused; writes the folded ct in place.
// ============================================================================================
// Signed digit k of `val` as a residue modulo q (lazy, < 4q), following split_and_crt's carry chain
// (reference src/spiral.cpp:282-329).  The digit is either a small value d <= 2^bp or d + Q - 2^bp;
// Q = 0 (mod p, b), so for bp <= 27 the residue is d or d + q - 2^bp with no 64-bit reduction.
struct SignedDigitPlan {          // everything that depends only on (k, t): built once per CTA
    uint32_t off0, offk, lowbits, bits_per;
    uint64_t K, lowmask, mask, half;
    bool guard;
};
__device__ __forceinline__ SignedDigitPlan make_signed_digit_plan(int k, int t, uint32_t bits_per) {
    // The carry chain "piece > 2^(bp-1) -> carry" is ordinary carry propagation of adding the constant
    // K = sum_j (2^(bp-1) - 1) * 2^(j*bp) to the lower digits, so the carry into digit k is one bit of (L + K):
    // no loop per coefficient.  The first half restarts at digit 0, the second at digit t/2 (carry reset, :282,312).
    SignedDigitPlan p;
    const int half_elems = t / 2;
    const int k0 = k < half_elems ? 0 : half_elems;
    p.bits_per = bits_per;
    p.lowbits = (uint32_t)(k - k0) * bits_per;
    p.mask = (1ull << bits_per) - 1;
    p.off0 = min((uint32_t)k0 * bits_per, 64u) & 63;
    p.offk = min((uint32_t)k * bits_per, 64u) & 63;
    p.lowmask = p.lowbits ? ((1ull << p.lowbits) - 1) : 0;
    p.half = 1ull << (bits_per - 1);
    p.K = 0;
    for (uint32_t j = 0; j < p.lowbits; j += bits_per) p.K |= (p.half - 1) << j;
    p.guard = (k < half_elems) ? (k + 1 < half_elems) : true;                 // first half: k < num_elems/2 - 1
    return p;
}
__device__ __forceinline__ uint32_t signed_digit_res(uint64_t val, const SignedDigitPlan &p, uint32_t q, int n) {
    const uint64_t carry = (((val >> p.off0) & p.lowmask) + p.K) >> p.lowbits;   // 0 or 1 (K = 0, lowmask = 0 when lowbits = 0)
    const uint64_t piece = ((val >> p.offk) & p.mask) + (p.lowbits ? carry : 0);
    const bool wrapped = piece > p.half && p.guard;
    if (p.bits_per <= 27) return wrapped ? (uint32_t)piece + q - (1u << p.bits_per) : (uint32_t)piece;
    return raw_to_res(wrapped ? piece + kQ - (1ull << p.bits_per) : piece, n);
}
// the same digit as a small signed integer (bits_per <= 27): piece, or piece - 2^bp where the reference adds Q - 2^bp
__device__ __forceinline__ int32_t signed_digit_small(uint64_t val, const SignedDigitPlan &p) {
    const uint64_t carry = (((val >> p.off0) & p.lowmask) + p.K) >> p.lowbits;
    const uint64_t piece = ((val >> p.offk) & p.mask) + (p.lowbits ? carry : 0);
    const bool wrapped = piece > p.half && p.guard;
    return wrapped ? (int32_t)piece - (int32_t)(1u << p.bits_per) : (int32_t)piece;
}
// Generic over the ciphertext shape so the Pack variant (foldCiphertextsDim1, src/testing.cpp:596-624:
// 2x1 ciphertexts, UNSIGNED gadget_invert digits, out_n^2 planes batched) shares the kernels:
//   R x Cc ciphertext, GSW is R x (R*t), digit k of input row r lands in row r + k*R.
struct FoldShape {
    int R, Cc, t, is_signed;
    int np;              // ciphertexts per plane AFTER this round
    int planes;          // independent planes folded by the same GSW ciphertext
    int plane_stride;    // ciphertexts between consecutive planes in `cts`
    int cmux;            // 1: resident path, C_lo + Q (x) (G^-1(C_hi) - G^-1(C_lo)); 0: the reference's two-product form
};
__global__ void __launch_bounds__(kNttThreads, 5) k_fold_decomp_ntt(uint32_t *__restrict__ scratch, const uint64_t *__restrict__ cts, FoldShape fs, int cts_per_plane) {
    pdl_begin();
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    const int n = plane_of_thread(), lt = lane_in_plane();
    prefetch_twiddles(c_ntt.fwd[n], lt);                   // constant: fetched while the previous kernel is still running
    pdl_wait();
    const int RC = fs.R * fs.Cc;
    // 1-D grid with the digit index fastest: the t CTAs that decompose the same polynomial run back to back, so all but the
    // first find it in L2 (with k slowest the re-reads were DRAM traffic: 8x the ciphertext bytes at SpiralPack sizes)
    const int ctpoly = (int)(blockIdx.x / (unsigned)fs.t), k = (int)(blockIdx.x % (unsigned)fs.t);   // ctpoly = (plane*cts_per_plane + ct)*RC + r*Cc + c
    const int ctd = ctpoly / RC, rc = ctpoly % RC, r = rc / fs.Cc, c = rc % fs.Cc;
    const int plane = ctd / cts_per_plane, ctl = ctd % cts_per_plane;
    const uint32_t bits_per = get_bits_per(fs.t);
    const uint64_t mask = (1ull << bits_per) - 1;
    const uint64_t *src = cts + ((size_t)(plane * fs.plane_stride + ctl) * RC + rc) * kN;
    const SignedDigitPlan plan = make_signed_digit_plan(k, fs.t, bits_per);
    const uint32_t q = modulus(n);
    uint32_t v[16];
    if (!fs.cmux) {
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const uint64_t val = __ldg(src + nat_pos(lt, e));
            v[e] = fs.is_signed ? signed_digit_res(val, plan, q, n)
                                : (bits_per <= 29 ? (uint32_t)gadget_digit(val, k, bits_per, mask) : raw_to_res(gadget_digit(val, k, bits_per, mask), n));
        }
    } else if (bits_per <= 27) {
        // CMux form: ctl < np is the LOW ciphertext i, its partner is np + i; one NTT of the digit difference.
        // A digit is a small integer s (|s| <= 2^bp; the reference's "s + Q" for negative digits is s modulo both primes),
        // so the difference s_hi - s_lo is the same small integer under both primes: each prime plane extracts HALF of the
        // coefficients and the halves meet in shared memory.  When a digit and the carry chain below it live in one 32-bit
        // word of the coefficient (t * bp = 64 with word-aligned halves, e.g. t_GSW = 8) the extraction is 32-bit arithmetic.
        __shared__ int32_t dsm[kN];
        const uint64_t *hi = src + (size_t)fs.np * RC * kN;
        const int half_elems = fs.t / 2;
        const bool word32 = fs.is_signed ? (bits_per * (uint32_t)half_elems == 32u)
                                         : ((uint32_t)k * bits_per / 32u == ((uint32_t)(k + 1) * bits_per - 1u) / 32u && (uint32_t)(k + 1) * bits_per <= 64u);
        if (word32) {
            const int wsel = fs.is_signed ? (k >= half_elems) : (int)((uint32_t)k * bits_per / 32u);
            const uint32_t sh = fs.is_signed ? plan.lowbits : ((uint32_t)k * bits_per) & 31u;
            const uint32_t m32 = (uint32_t)mask, K32 = (uint32_t)plan.K, lm32 = (uint32_t)plan.lowmask, half32 = (uint32_t)plan.half;
            const uint32_t *lo32 = reinterpret_cast<const uint32_t *>(src) + wsel, *hi32 = reinterpret_cast<const uint32_t *>(hi) + wsel;
            uint32_t av[8], bv[8];                       // all 16 loads in flight before the first use
#pragma unroll
            for (int e8 = 0; e8 < 8; e8++) { const int idx = nat_pos(lt, 8 * n + e8); av[e8] = __ldg(lo32 + 2 * idx); bv[e8] = __ldg(hi32 + 2 * idx); }
#pragma unroll
            for (int e8 = 0; e8 < 8; e8++) {
                const int idx = nat_pos(lt, 8 * n + e8);
                const uint32_t a = av[e8], b = bv[e8];
                int32_t sa, sb;
                if (fs.is_signed) {
                    const uint32_t pa = ((a >> sh) & m32) + (((a & lm32) + K32) >> sh), pb = ((b >> sh) & m32) + (((b & lm32) + K32) >> sh);
                    sa = (pa > half32 && plan.guard) ? (int32_t)pa - (int32_t)(1u << bits_per) : (int32_t)pa;
                    sb = (pb > half32 && plan.guard) ? (int32_t)pb - (int32_t)(1u << bits_per) : (int32_t)pb;
                } else { sa = (int32_t)((a >> sh) & m32); sb = (int32_t)((b >> sh) & m32); }
                dsm[idx] = sb - sa;
            }
        } else {
            uint64_t av[8], bv[8];
#pragma unroll
            for (int e8 = 0; e8 < 8; e8++) { const int idx = nat_pos(lt, 8 * n + e8); av[e8] = __ldg(src + idx); bv[e8] = __ldg(hi + idx); }
#pragma unroll
            for (int e8 = 0; e8 < 8; e8++) {
                const int idx = nat_pos(lt, 8 * n + e8);
                const uint64_t a = av[e8], b = bv[e8];
                int32_t sa, sb;
                if (fs.is_signed) { sa = signed_digit_small(a, plan); sb = signed_digit_small(b, plan); }
                else { sa = (int32_t)gadget_digit(a, k, bits_per, mask); sb = (int32_t)gadget_digit(b, k, bits_per, mask); }
                dsm[idx] = sb - sa;
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 16; e++) v[e] = (uint32_t)(dsm[nat_pos(lt, e)] + (int32_t)(2 * q));   // in (0, 4q): a valid lazy NTT input
    } else {
        const uint64_t *hi = src + (size_t)fs.np * RC * kN;
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const uint64_t a = __ldg(src + nat_pos(lt, e)), b = __ldg(hi + nat_pos(lt, e));
            uint32_t da, db;
            if (fs.is_signed) { da = signed_digit_res(a, plan, q, n); db = signed_digit_res(b, plan, q, n); }
            else { da = raw_to_res(gadget_digit(a, k, bits_per, mask), n); db = raw_to_res(gadget_digit(b, k, bits_per, mask), n); }
            v[e] = db + 2 * q - da;          // in (0, 4q): a valid lazy NTT input
        }
    }
    ntt_forward_plane(v, sm[n], lt, n);
    const int m2 = fs.R * fs.t, row = r + k * fs.R;
    store_ntt_regs(v, scratch + ((((size_t)ctd * m2 + row) * fs.Cc + c) * 2 + n) * kN, lt);
}
// Pointwise 2*m2-term dot product for every output polynomial.  CTA = 64 uint4 columns (256 coefficients) x 4 term
// groups: each thread owns every 4th term, so its <= 12 (t_GSW = 8) 128-bit load pairs are all independent and in
// flight together; the four partial sums meet in shared memory.  16 CTAs per output polynomial keep late rounds
// (a handful of ciphertexts) spread over the chip instead of serialising a 48-step load->MAC chain.
constexpr int kMacCols = 64, kMacGroups = 4;
__global__ void __launch_bounds__(256) k_fold_mac(uint32_t *__restrict__ out, const uint32_t *__restrict__ scratch,
                                                  const uint32_t *__restrict__ q_dev, const uint32_t *__restrict__ qneg_dev, FoldShape fs) {
    pdl_begin();
    __shared__ ulonglong2 part[kMacGroups - 1][kMacCols][2];
    const int RC = fs.R * fs.Cc;
    const int op = blockIdx.x >> 4, seg = blockIdx.x & 15;
    const int id = op / RC, rc = op % RC, r = rc / fs.Cc, c = rc % fs.Cc;
    const int plane = id / fs.np, i = id % fs.np;
    const int m2 = fs.R * fs.t;
    const int col = threadIdx.x & (kMacCols - 1), grp = threadIdx.x >> 6;
    const int w4 = seg * kMacCols + col, n = w4 >= 512;
    const size_t qs = 2 * kN / 4, cs = (size_t)fs.Cc * 2 * kN / 4;
    const uint4 *Qn = reinterpret_cast<const uint4 *>(qneg_dev + (size_t)r * m2 * 2 * kN) + w4;
    const uint4 *Qp = reinterpret_cast<const uint4 *>(q_dev + (size_t)r * m2 * 2 * kN) + w4;
    const uint4 *C0 = reinterpret_cast<const uint4 *>(scratch + (((size_t)(plane * 2 * fs.np + i) * m2) * fs.Cc + c) * 2 * kN) + w4;
    const uint4 *C1 = reinterpret_cast<const uint4 *>(scratch + (((size_t)(plane * 2 * fs.np + fs.np + i) * m2) * fs.Cc + c) * 2 * kN) + w4;
    uint64_t acc[4] = {0, 0, 0, 0};
    int cnt = 0;
    if (fs.cmux && m2 <= 6 * kMacGroups) {
        // short chains (t_GSW <= 8: at most six terms per thread): the GSW words - complete before this graph starts - are in
        // registers before the digits of the previous kernel exist
        const uint4 *Cd = reinterpret_cast<const uint4 *>(scratch + (((size_t)(plane * fs.np + i) * m2) * fs.Cc + c) * 2 * kN) + w4;
        uint4 x[6];
#pragma unroll
        for (int t = 0; t < 6; t++) { const int m = grp + t * kMacGroups; x[t] = m < m2 ? __ldg(Qp + m * qs) : make_uint4(0, 0, 0, 0); }
        pdl_wait();
#pragma unroll
        for (int t = 0; t < 6; t++) {
            const int m = grp + t * kMacGroups;
            if (m < m2) {
                const uint4 y = __ldg(Cd + m * cs);
                acc[0] += (uint64_t)x[t].x * y.x; acc[1] += (uint64_t)x[t].y * y.y;
                acc[2] += (uint64_t)x[t].z * y.z; acc[3] += (uint64_t)x[t].w * y.w;
            }
        }
    } else if (fs.cmux) {        // scratch holds one difference-digit set per OUTPUT ciphertext (dense index plane*np + i)
        pdl_wait();
        const uint4 *Cd = reinterpret_cast<const uint4 *>(scratch + (((size_t)(plane * fs.np + i) * m2) * fs.Cc + c) * 2 * kN) + w4;
#pragma unroll 4
        for (int m = grp; m < m2; m += kMacGroups) {
            const uint4 x = __ldg(Qp + m * qs), y = __ldg(Cd + m * cs);
            acc[0] += (uint64_t)x.x * y.x; acc[1] += (uint64_t)x.y * y.y;
            acc[2] += (uint64_t)x.z * y.z; acc[3] += (uint64_t)x.w * y.w;
            if (++cnt == 120) {
                cnt = 0;
#pragma unroll
                for (int e = 0; e < 4; e++) acc[e] = reduce_u64(acc[e], n);
            }
        }
    } else {
    pdl_wait();
#pragma unroll 4
    for (int tm = grp; tm < 2 * m2; tm += kMacGroups) {
        const int h = tm >= m2, m = h ? tm - m2 : tm;
        const uint4 x = __ldg((h ? Qp : Qn) + m * qs), y = __ldg((h ? C1 : C0) + m * cs);
        acc[0] += (uint64_t)x.x * y.x; acc[1] += (uint64_t)x.y * y.y;
        acc[2] += (uint64_t)x.z * y.z; acc[3] += (uint64_t)x.w * y.w;
        if (++cnt == 120) {
            cnt = 0;
#pragma unroll
            for (int e = 0; e < 4; e++) acc[e] = reduce_u64(acc[e], n);
        }
    }
    }
#pragma unroll
    for (int e = 0; e < 4; e++) acc[e] = reduce_u64(acc[e], n);
    if (grp > 0) {
        part[grp - 1][col][0] = make_ulonglong2(acc[0], acc[1]);
        part[grp - 1][col][1] = make_ulonglong2(acc[2], acc[3]);
    }
    __syncthreads();
    if (grp == 0) {
#pragma unroll
        for (int g = 0; g < kMacGroups - 1; g++) {
            const ulonglong2 a = part[g][col][0], b = part[g][col][1];
            acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
        }
        reinterpret_cast<uint4 *>(out + (size_t)op * 2 * kN)[w4] =
            make_uint4(reduce_u64(acc[0], n), reduce_u64(acc[1], n), reduce_u64(acc[2], n), reduce_u64(acc[3], n));
    }
}
// Throughput variant for rounds with many outputs (CMux form only): CTA = 256 uint4 columns, every thread walks all m2
// terms of its column, so there is no partial-sum exchange and one Barrett reduction per output word instead of eight.
__global__ void __launch_bounds__(256) k_fold_mac_wide(uint32_t *__restrict__ out, const uint32_t *__restrict__ scratch,
                                                       const uint32_t *__restrict__ q_dev, FoldShape fs) {
    pdl_prologue();
    const int RC = fs.R * fs.Cc;
    const int op = blockIdx.x >> 2, seg = blockIdx.x & 3;
    const int id = op / RC, rc = op % RC, r = rc / fs.Cc, c = rc % fs.Cc;
    const int m2 = fs.R * fs.t;
    const int w4 = seg * 256 + threadIdx.x, n = w4 >= 512;
    const size_t qs = 2 * kN / 4, cs = (size_t)fs.Cc * 2 * kN / 4;
    const uint4 *Qp = reinterpret_cast<const uint4 *>(q_dev + (size_t)r * m2 * 2 * kN) + w4;
    const uint4 *Cd = reinterpret_cast<const uint4 *>(scratch + (((size_t)id * m2) * fs.Cc + c) * 2 * kN) + w4;   // id = plane*np + i
    uint64_t acc[4] = {0, 0, 0, 0};
    for (int m0 = 0; m0 < m2; m0 += 96) {                    // 96 products of < 2^56 stay below 2^63
        const int m1 = m0 + 96 < m2 ? m0 + 96 : m2;
#pragma unroll 8
        for (int m = m0; m < m1; m++) {
            const uint4 x = __ldg(Qp + m * qs), y = __ldg(Cd + m * cs);
            acc[0] += (uint64_t)x.x * y.x; acc[1] += (uint64_t)x.y * y.y;
            acc[2] += (uint64_t)x.z * y.z; acc[3] += (uint64_t)x.w * y.w;
        }
        if (m1 < m2) {
#pragma unroll
            for (int e = 0; e < 4; e++) acc[e] = reduce_u64(acc[e], n);
        }
    }
    reinterpret_cast<uint4 *>(out + (size_t)op * 2 * kN)[w4] =
        make_uint4(reduce_u64(acc[0], n), reduce_u64(acc[1], n), reduce_u64(acc[2], n), reduce_u64(acc[3], n));
}
// inverse NTT + CRT lift of the dense MAC outputs back into the (strided) ciphertext array
__global__ void __launch_bounds__(kNttThreads) k_fold_lift(uint64_t *__restrict__ cts, const uint32_t *__restrict__ macout, FoldShape fs) {
    pdl_begin();
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    const int n = plane_of_thread(), lt = lane_in_plane();
    prefetch_twiddles(c_ntt.inv[n], lt);
    const int RC = fs.R * fs.Cc;
    const int id = blockIdx.x / RC, rc = blockIdx.x % RC;
    const int plane = id / fs.np, i = id % fs.np;
    uint64_t *dst = cts + ((size_t)(plane * fs.plane_stride + i) * RC + rc) * kN;
    // NOTE: only data that is constant for the whole query may be read before pdl_wait(): early launches cascade (this kernel can be
    // resident while the lift TWO rounds back is still writing), so C_lo - written "two kernels ago" - is read after the wait
    pdl_wait();
    uint64_t clo[8];
    if (fs.cmux) {
#pragma unroll
        for (int k = 0; k < 8; k++) clo[k] = dst[threadIdx.x + 256 * k];      // in flight during the inverse transform
    }
    uint32_t v[16];
    load_ntt_regs(v, macout + ((size_t)blockIdx.x * 2 + n) * kN, lt);
    ntt_inverse_plane(v, sm[n], lt, n);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) sm[n][nat_pos(lt, k)] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int z = threadIdx.x + 256 * k;
        uint64_t val = crt_compose(sm[0][z], sm[1][z]);
        if (fs.cmux) {                      // + C_lo (G * G^-1(C_lo) = C_lo mod Q), canonical result
            uint64_t lo = clo[k];
            lo = lo >= kQ ? lo - kQ : lo;
            val += lo;
            val = val >= kQ ? val - kQ : val;
        }
        dst[z] = val;
    }
}
// digits of every input ciphertext + the dense MAC outputs (one per output polynomial)
size_t fold_scratch_words_generic(size_t cts_in, int R, int Cc, int t) { return cts_in * (size_t)R * t * Cc * 2 * kN + (cts_in / 2 + 1) * (size_t)R * Cc * 2 * kN; }
void launch_fold_round_generic(uint64_t *cts, int R, int Cc, int t, int is_signed, size_t np_after, size_t planes, size_t plane_stride,
                               const uint32_t *q_dev, const uint32_t *qneg_dev, uint32_t *scratch, cudaStream_t s) {
    // qneg_dev == nullptr selects the CMux form (resident servers): with Qneg = G - Q (mod Q) and an exact gadget
    // decomposition,  Qneg (x) G^-1(C_lo) + Q (x) G^-1(C_hi) = C_lo + Q (x) (G^-1(C_hi) - G^-1(C_lo))  (mod Q),
    // and the output is the canonical representative either way - bit-identical, half the NTTs and MAC traffic.
    const int cmux = qneg_dev == nullptr;
    FoldShape fs{R, Cc, t, is_signed, (int)np_after, (int)planes, (int)plane_stride, cmux};
    const int RC = R * Cc, cpp = (int)((cmux ? 1 : 2) * np_after);
    count_launch(); launch_pdl(k_fold_decomp_ntt, dim3((unsigned)(planes * cpp * RC * t)), dim3(kNttThreads), 0, s, scratch, cts, fs, cpp);
    uint32_t *macout = scratch + (size_t)planes * cpp * R * t * Cc * 2 * kN;
    const size_t outputs = planes * np_after * RC;
    count_launch();
    if (cmux && outputs >= 384) launch_pdl(k_fold_mac_wide, dim3((unsigned)(outputs * 4)), dim3(256), 0, s, macout, (const uint32_t *)scratch, q_dev, fs);
    else launch_pdl(k_fold_mac, dim3((unsigned)(outputs * 16)), dim3(256), 0, s, macout, scratch, q_dev, qneg_dev, fs);
    count_launch(); launch_pdl(k_fold_lift, dim3((unsigned)(planes * np_after * RC)), dim3(kNttThreads), 0, s, cts, macout, fs);
}
// reference reorient_Q layout (packed [z][r*m2 + m], src/spiral.cpp:388-400) -> dev-NTT [r*m2 + m][n][z]
__global__ void k_unreorient_q(uint32_t *__restrict__ out, const uint64_t *__restrict__ q_reor, int rm_count) {
    pdl_prologue();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // (rm, z), z fastest
    if (idx >= (size_t)rm_count * kN) return;
    const int z = (int)(idx % kN), rm = (int)(idx / kN);
    const uint64_t w = q_reor[(size_t)z * rm_count + rm];
    out[((size_t)rm * 2) * kN + z] = (uint32_t)w % kP;
    out[((size_t)rm * 2 + 1) * kN + z] = (uint32_t)(w >> 32) % kB;
}
void launch_unreorient_q(uint32_t *out, const uint64_t *q_reor, int rm_count, cudaStream_t s) {
    const size_t n = (size_t)rm_count * kN;
    count_launch(); launch_pdl(k_unreorient_q, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, out, q_reor, rm_count);
}
// scan layout -> reference layout (inverse of launch_db_from_reference), whole database
void launch_db_to_reference(uint64_t *B_ref, const uint64_t *db, size_t dim0, size_t ic, cudaStream_t s);

size_t fold_scratch_words(size_t num_per_half, int t_gsw) { return fold_scratch_words_generic(2 * num_per_half, kN1, kN2, t_gsw); }
// split_and_crt alone (reference src/spiral.cpp:270-341) on `count` ciphertexts: scratch[ct][m][c] dev-NTT
void launch_fold_decomp_only(uint32_t *scratch, const uint64_t *cts, size_t count, int t_gsw, cudaStream_t s) {
    FoldShape fs{kN1, kN2, t_gsw, 1, (int)count, 1, (int)count, 0};
    if (count) { count_launch(); launch_pdl(k_fold_decomp_ntt, dim3((unsigned)(count * 6 * t_gsw)), dim3(kNttThreads), 0, s, scratch, cts, fs, (int)count); }
}
void launch_fold_round(uint64_t *cts, size_t num_per, const uint32_t *q_dev, const uint32_t *qneg_dev,
                       int t_gsw, uint32_t *scratch, cudaStream_t s) {
    launch_fold_round_generic(cts, kN1, kN2, t_gsw, 1, num_per, 1, 2 * num_per, q_dev, qneg_dev, scratch, s);
}

}  // namespace sb200
