// ntt.cuh - CTA-level negacyclic 2048-point NTT / inverse NTT over the two CRT primes.
//
// Same transform as the reference's ntt_forward/ntt_inverse (src/core.cpp:254-514): same
// primitive roots (psi_p = 66687, psi_b = 158221), natural-order input, bit-reversed output,
// so NTT-domain data can cross the boundary unchanged.  The schedule is B200-native:
//   * one CTA transforms one polynomial under BOTH primes: 256 threads = 2 planes x 128 threads,
//     16 coefficients per thread held in registers;
//   * the 11 butterfly stages run as three register passes (4 + 4 + 3 stages) with two exchanges
//     through a padded, bank-conflict-free shared-memory plane (8 KiB + padding per prime);
//   * Harvey lazy butterflies with Shoup twiddles (w, floor(w*2^32/q)) fetched as one 8-byte load;
//   * each plane synchronises on its own named barrier, so the two primes never wait on each other.
// Outputs are always canonical ([0,q)); the reference's AVX2 path may leave q for 0, which is
// equal modulo q and indistinguishable to every consumer.
#pragma once
#include "common.cuh"

namespace sb200 {

constexpr int kNttThreads = 256;          // threads per CTA for every NTT-based kernel
constexpr int kPlaneThreads = 128;        // threads per prime plane
constexpr int kPlaneWords = kN + (kN >> 7) * 8;   // 2048 + 16*8 padding words

struct NttTables {
    const uint2 *fwd[2];     // fwd[n][m + i] = (w, w') with w = psi^bitrev(m+i)           [2048 entries]
    const uint2 *inv[2];     // inv[n][h + i] = (w, w') with w = psi^-bitrev(h+i) (no 1/2)  [2048 entries]
    uint2 ninv[2];           // N^-1 mod q (Shoup pair)
    uint2 inv_last[2];       // inv[n][1] * N^-1 (Shoup pair) - last inverse stage with the scaling folded in
};
// The library is built as ONE translation unit (spiral_b200.cu), so these are the only copies.
__constant__ NttTables c_ntt;
__constant__ uint2 c_fwd_head[2][16];   // fwd[n][0..15]: uniform twiddles of the first forward pass
__constant__ uint2 c_inv_head[2][16];   // inv[n][0..15]: uniform twiddles of the last inverse pass

__device__ __forceinline__ int phys(int i) { return i + ((i >> 7) << 3); }

__device__ __forceinline__ void plane_sync(int plane) {     // named barriers 1 and 2 (0 is __syncthreads)
    if (plane == 0) asm volatile("bar.sync 1, %0;" ::"n"(kPlaneThreads) : "memory");
    else            asm volatile("bar.sync 2, %0;" ::"n"(kPlaneThreads) : "memory");
}

// Cooley-Tukey (forward) butterfly: x,y in [0,4q) -> [0,4q)
__device__ __forceinline__ void bf_fwd(uint32_t &x, uint32_t &y, uint2 w, uint32_t q, uint32_t q2) {
    uint32_t cx = min(x, x - q2);
    uint32_t t = mul_shoup_lazy(y, w.x, w.y, q);
    x = cx + t;
    y = cx + q2 - t;
}
// Gentleman-Sande (inverse) butterfly: u,v in [0,2q) -> [0,2q)
__device__ __forceinline__ void bf_inv(uint32_t &u, uint32_t &v, uint2 w, uint32_t q, uint32_t q2) {
    uint32_t s = u + v;
    uint32_t t = u - v + q2;
    u = min(s, s - q2);
    v = mul_shoup_lazy(t, w.x, w.y, q);
}

// Warm the twiddles of the later passes while the first pass computes (each is an L2 round trip otherwise).
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_twiddles(const uint2 *tw, int lt) {
    const int blk = lt >> 3;
    prefetch_l1(tw + 16 + blk); prefetch_l1(tw + 32 + 2 * blk); prefetch_l1(tw + 64 + 4 * blk); prefetch_l1(tw + 128 + 8 * blk);
    prefetch_l1(tw + 256 + lt); prefetch_l1(tw + 256 + 128 + lt);
    prefetch_l1(tw + 512 + 2 * lt); prefetch_l1(tw + 512 + 256 + 2 * lt);
    prefetch_l1(tw + 1024 + 4 * lt); prefetch_l1(tw + 1024 + 512 + 4 * lt);
}

// Forward NTT of one prime plane.
//   in : v[k] = a[lt + 128*k]  (natural coefficient order), any value < 4q
//   out: v[k] = A[8*lt + k] for k < 8, A[1024 + 8*lt + (k-8)] for k >= 8  (reference NTT order), canonical
// `pl` is this plane's shared buffer (kPlaneWords words).  All 128 threads of the plane must call.
__device__ __forceinline__ void ntt_forward_plane(uint32_t (&v)[16], uint32_t *pl, int lt, int n) {
    const uint32_t q = modulus(n), q2 = 2 * q;
    const uint2 *__restrict__ tw = c_ntt.fwd[n];
    prefetch_twiddles(tw, lt);
    // pass A: stages 0..3, distances 1024,512,256,128 = 8,4,2,1 register steps
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int half = 8 >> s;
#pragma unroll
        for (int k = 0; k < 16; k++)
            if ((k & half) == 0) bf_fwd(v[k], v[k + half], c_fwd_head[n][(1 << s) + (k >> (4 - s))], q, q2);
    }
    plane_sync(n);     // previous users of `pl` are done
#pragma unroll
    for (int k = 0; k < 16; k++) pl[phys(lt + 128 * k)] = v[k];
    plane_sync(n);
    // pass B: stages 4..7, distances 64,32,16,8
    const int blk = lt >> 3, jp = lt & 7;
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = pl[phys(blk * 128 + jp + 8 * k)];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const int half = 8 >> s;
        const uint2 *base = tw + (16 << s) + (blk << s);
#pragma unroll
        for (int k = 0; k < 16; k++)
            if ((k & half) == 0) bf_fwd(v[k], v[k + half], __ldg(base + (k >> (4 - s))), q, q2);
    }
#pragma unroll
    for (int k = 0; k < 16; k++) pl[phys(blk * 128 + jp + 8 * k)] = v[k];
    plane_sync(n);
    // pass C: stages 8..10, distances 4,2,1 inside two blocks of 8 contiguous coefficients
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int base = h * 1024 + 8 * lt;
        const uint4 a = *reinterpret_cast<const uint4 *>(pl + phys(base));
        const uint4 b = *reinterpret_cast<const uint4 *>(pl + phys(base) + 4);
        uint32_t *u = &v[8 * h];
        u[0] = a.x; u[1] = a.y; u[2] = a.z; u[3] = a.w; u[4] = b.x; u[5] = b.y; u[6] = b.z; u[7] = b.w;
        const uint2 w8 = __ldg(tw + 256 + (base >> 3));
#pragma unroll
        for (int k = 0; k < 4; k++) bf_fwd(u[k], u[k + 4], w8, q, q2);
        const uint2 w9a = __ldg(tw + 512 + (base >> 2)), w9b = __ldg(tw + 512 + (base >> 2) + 1);
        bf_fwd(u[0], u[2], w9a, q, q2); bf_fwd(u[1], u[3], w9a, q, q2);
        bf_fwd(u[4], u[6], w9b, q, q2); bf_fwd(u[5], u[7], w9b, q, q2);
#pragma unroll
        for (int k = 0; k < 4; k++) bf_fwd(u[2 * k], u[2 * k + 1], __ldg(tw + 1024 + (base >> 1) + k), q, q2);
    }
#pragma unroll
    for (int k = 0; k < 16; k++) { uint32_t x = min(v[k], v[k] - q2); v[k] = min(x, x - q); }
}

// Inverse NTT of one prime plane (scaling by N^-1 included).
//   in : v[k] = A[8*lt + k] (k<8), A[1024 + 8*lt + k-8] (k>=8), any value < 2q
//   out: v[k] = a[lt + 128*k], canonical
__device__ __forceinline__ void ntt_inverse_plane(uint32_t (&v)[16], uint32_t *pl, int lt, int n) {
    const uint32_t q = modulus(n), q2 = 2 * q;
    const uint2 *__restrict__ tw = c_ntt.inv[n];
    prefetch_twiddles(tw, lt);
    // pass C': distances 1,2,4
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int base = h * 1024 + 8 * lt;
        uint32_t *u = &v[8 * h];
#pragma unroll
        for (int k = 0; k < 4; k++) bf_inv(u[2 * k], u[2 * k + 1], __ldg(tw + 1024 + (base >> 1) + k), q, q2);
        const uint2 w9a = __ldg(tw + 512 + (base >> 2)), w9b = __ldg(tw + 512 + (base >> 2) + 1);
        bf_inv(u[0], u[2], w9a, q, q2); bf_inv(u[1], u[3], w9a, q, q2);
        bf_inv(u[4], u[6], w9b, q, q2); bf_inv(u[5], u[7], w9b, q, q2);
        const uint2 w8 = __ldg(tw + 256 + (base >> 3));
#pragma unroll
        for (int k = 0; k < 4; k++) bf_inv(u[k], u[k + 4], w8, q, q2);
    }
    plane_sync(n);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int base = h * 1024 + 8 * lt;
        const uint32_t *u = &v[8 * h];
        *reinterpret_cast<uint4 *>(pl + phys(base)) = make_uint4(u[0], u[1], u[2], u[3]);
        *reinterpret_cast<uint4 *>(pl + phys(base) + 4) = make_uint4(u[4], u[5], u[6], u[7]);
    }
    plane_sync(n);
    // pass B': distances 8,16,32,64
    const int blk = lt >> 3, jp = lt & 7;
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = pl[phys(blk * 128 + jp + 8 * k)];
#pragma unroll
    for (int s = 3; s >= 0; s--) {
        const int half = 8 >> s;
        const uint2 *base = tw + (16 << s) + (blk << s);
#pragma unroll
        for (int k = 0; k < 16; k++)
            if ((k & half) == 0) bf_inv(v[k], v[k + half], __ldg(base + (k >> (4 - s))), q, q2);
    }
#pragma unroll
    for (int k = 0; k < 16; k++) pl[phys(blk * 128 + jp + 8 * k)] = v[k];
    plane_sync(n);
    // pass A': distances 128,256,512,1024; the last stage carries N^-1
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = pl[phys(lt + 128 * k)];
#pragma unroll
    for (int s = 3; s >= 1; s--) {
        const int half = 8 >> s;
#pragma unroll
        for (int k = 0; k < 16; k++)
            if ((k & half) == 0) bf_inv(v[k], v[k + half], c_inv_head[n][(1 << s) + (k >> (4 - s))], q, q2);
    }
    const uint2 ni = c_ntt.ninv[n], wl = c_ntt.inv_last[n];
#pragma unroll
    for (int k = 0; k < 8; k++) {
        uint32_t s = v[k] + v[k + 8];                   // < 4q
        uint32_t t = v[k] - v[k + 8] + q2;              // < 4q
        uint32_t a = mul_shoup_lazy(s, ni.x, ni.y, q);
        uint32_t b = mul_shoup_lazy(t, wl.x, wl.y, q);
        v[k] = min(a, a - q);
        v[k + 8] = min(b, b - q);
    }
}

// Convenience: which plane / local thread am I (blockDim.x == kNttThreads).
__device__ __forceinline__ int plane_of_thread() { return threadIdx.x >> 7; }
__device__ __forceinline__ int lane_in_plane() { return threadIdx.x & 127; }

// NTT-order position held in register k by plane-thread lt (output of forward / input of inverse)
__device__ __forceinline__ int ntt_pos(int lt, int k) { return (k < 8) ? (8 * lt + k) : (1024 + 8 * lt + (k - 8)); }
// natural-order coefficient index held in register k (input of forward / output of inverse)
__device__ __forceinline__ int nat_pos(int lt, int k) { return lt + 128 * k; }

// Load / store a plane in device NTT format (u32 [2][2048], plane n at +n*2048) using 128-bit accesses.
__device__ __forceinline__ void load_ntt_regs(uint32_t (&v)[16], const uint32_t *__restrict__ plane_ptr, int lt) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint4 *src = reinterpret_cast<const uint4 *>(plane_ptr + h * 1024 + 8 * lt);
        uint4 a = __ldg(src), b = __ldg(src + 1);
        uint32_t *u = &v[8 * h];
        u[0] = a.x; u[1] = a.y; u[2] = a.z; u[3] = a.w; u[4] = b.x; u[5] = b.y; u[6] = b.z; u[7] = b.w;
    }
}
__device__ __forceinline__ void store_ntt_regs(const uint32_t (&v)[16], uint32_t *__restrict__ plane_ptr, int lt) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
        uint4 *dst = reinterpret_cast<uint4 *>(plane_ptr + h * 1024 + 8 * lt);
        const uint32_t *u = &v[8 * h];
        dst[0] = make_uint4(u[0], u[1], u[2], u[3]);
        dst[1] = make_uint4(u[4], u[5], u[6], u[7]);
    }
}

}  // namespace sb200
