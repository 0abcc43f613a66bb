// common.cuh - ring constants and modular arithmetic for the Spiral server path on sm_100a.
//
// Ring: Z_Q[x]/(x^2048+1), Q = p*b with the reference's two 28-bit NTT primes
// (reference include/values.h:7-41).  All arithmetic is exact unsigned-integer modular
// arithmetic; every routine here returns canonical representatives unless it says "lazy".
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace sb200 {

constexpr int      kLogN = 11;
constexpr int      kN    = 1 << kLogN;                 // poly_len            values.h:11
constexpr uint32_t kP    = 268369921u;                 // p_i = 2^28-2^16+1   values.h:13
constexpr uint32_t kB    = 249561089u;                 // b_i                 values.h:21
constexpr uint64_t kQ    = 66974689739603969ull;       // Q_i = p*b           values.h:41
constexpr uint64_t kPsiP = 66687ull;                   // primitive 4096-th roots of unity that
constexpr uint64_t kPsiB = 158221ull;                  // regenerate src/constants.cpp:16
constexpr uint64_t kCr1P = 68736257792ull;             // floor(2^64/p)       values.h:59
constexpr uint64_t kCr1B = 73916747789ull;             // floor(2^64/b)       values.h:61
constexpr uint32_t kBInvModP = 163640210u;             // b^-1 mod p          values.h:25
constexpr uint32_t kBInvModPShoup = 2618882725u;       // floor(b^-1 * 2^32 / p)
constexpr uint32_t kPInvModB = 97389680u;              // p^-1 mod b          values.h:24
constexpr int kN0 = 2, kN1 = 3, kN2 = 2;               // values.h:67-69

__host__ __device__ __forceinline__ constexpr uint32_t modulus(int n) { return n == 0 ? kP : kB; }

// ---- u64 -> [0,q): Barrett with floor(2^64/q) (one correction), as reference include/poly.h:137-146
__device__ __forceinline__ uint32_t reduce_u64(uint64_t x, int n) {
    const uint64_t cr = n == 0 ? kCr1P : kCr1B;
    const uint32_t q = modulus(n);
    uint64_t hi = __umul64hi(x, cr);
    uint32_t t = (uint32_t)x - (uint32_t)hi * q;        // true remainder estimate is < 2q < 2^32
    return t >= q ? t - q : t;
}

// ---- a*b mod q for a,b < 2^32 with a*b < 2^60 (any a,b < 2^30): 32-bit Barrett
// mu = floor(2^57 / q) < 2^30; x>>27 < 2^33 needs care, so inputs are kept < 2^29 (a*b < 2^58).
__device__ __forceinline__ uint32_t mulmod(uint32_t a, uint32_t b, int n) {
    return reduce_u64((uint64_t)a * b, n);
}

// ---- Shoup multiplication: w < q fixed with wp = floor(w*2^32/q); y any u32.  Lazy result in [0,2q).
__device__ __forceinline__ uint32_t mul_shoup_lazy(uint32_t y, uint32_t w, uint32_t wp, uint32_t q) {
    uint32_t h = __umulhi(y, wp);
    return y * w - h * q;
}
__device__ __forceinline__ uint32_t csub(uint32_t x, uint32_t q) {   // x in [0,2q) -> [0,q)
    return min(x, x - q);                                           // unsigned wrap trick
}

// ---- CRT lift (x mod p, y mod b) -> canonical value in [0,Q).  Garner form of the reference's
// crt_compose (src/poly.cpp:344-353): same canonical result, no 128-bit Barrett.
//   v = y + b * ((x - y) * b^-1 mod p)
__device__ __forceinline__ uint64_t crt_compose(uint32_t x, uint32_t y) {
    uint32_t ymp = y >= kP ? y - kP : y;                 // y < b < p, kept for safety with y == b
    uint32_t d = x >= ymp ? x - ymp : x + kP - ymp;      // (x - y) mod p
    uint32_t t = csub(mul_shoup_lazy(d, kBInvModP, kBInvModPShoup, kP), kP);   // d * b^-1 mod p with the constant's Shoup companion
    return (uint64_t)y + (uint64_t)kB * t;
}

// ---- raw coefficient in [0, 2^64) -> residues
__device__ __forceinline__ uint32_t raw_to_res(uint64_t v, int n) { return reduce_u64(v, n); }

__host__ __device__ __forceinline__ uint32_t get_bits_per(uint32_t dim) {   // reference include/util.h:34-38
    if (dim == 56) return 1;
    return 56 / dim + 1;
}

// unsigned gadget digit k of val (reference src/util.cpp:135-141).  off == 64 would behave as the
// reference's x86 shift (by 0); unreachable for every surveyed parameter set.
__device__ __forceinline__ uint64_t gadget_digit(uint64_t val, int k, uint32_t bits_per, uint64_t mask) {
    uint32_t off = min((uint32_t)k * bits_per, 64u);
    return (val >> (off & 63)) & mask;
}

// ---- programmatic dependent launch (sm_90+): every kernel of the query path starts with pdl_prologue().
// launch_dependents lets the NEXT kernel of the stream / graph be scheduled while this one is still running;
// wait blocks until the PREVIOUS kernel has completed and its writes are visible, so nothing is consumed early.
// The dozens of short dependent kernels of one query thereby overlap their launch + ramp-up latencies.
struct TraceCtl { unsigned long long *buf; unsigned int *counter; unsigned int cap; };   // timeline trace (ntt_kernels.cu); buf == nullptr: off
__constant__ TraceCtl c_trace;
__device__ __forceinline__ unsigned long long global_timer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// every kernel starts with pdl_prologue(); the macro passes the call site's line, which names the kernel in the trace
#define pdl_prologue() pdl_prologue_at(__LINE__, true)
// Split form for the kernels of the latency chains: pdl_begin() lets the dependents launch, then the kernel issues loads of data
// that is CONSTANT FOR THE WHOLE QUERY (index lists, keys, twiddles, permutations) and only then pdl_wait()s; that part runs while
// the predecessor is still computing.  Nothing written during the query may be touched before the wait, however many kernels ago:
// early launches cascade, so a kernel can be resident while a kernel several links back in the chain is still running.
#define pdl_begin() unsigned long long pdl_t0_ = pdl_begin_at()
#define pdl_wait() pdl_wait_at(__LINE__, pdl_t0_)
// for kernels that may SPIN on a flag written by another GPU (or, in single-device tests, by another stream): their dependents must
// not be launched early - a pre-launched grid would sit in griddepcontrol.wait holding SM slots the flag's producer may need
#define pdl_prologue_no_early_dependents() pdl_prologue_at(__LINE__, false)
__device__ __forceinline__ unsigned long long pdl_begin_at() {
    const bool tr = c_trace.buf != nullptr && threadIdx.x == 0 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0;
    unsigned long long t0 = 0;
    if (tr) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0)); }
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    return t0;
}
__device__ __forceinline__ void pdl_wait_at(int line, unsigned long long t0);
__device__ __forceinline__ void pdl_prologue_at(int line, bool early_dependents) {
    const bool tr = c_trace.buf != nullptr && threadIdx.x == 0 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0;
    unsigned long long t0 = 0;
    if (tr) t0 = global_timer_ns();
    if (early_dependents) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tr) {
        const unsigned int i = atomicAdd(c_trace.counter, 1u);
        if (i < c_trace.cap) {
            c_trace.buf[3 * i] = t0; c_trace.buf[3 * i + 1] = global_timer_ns();
            c_trace.buf[3 * i + 2] = (unsigned long long)(gridDim.x & 0xFFFFFu) | ((unsigned long long)(gridDim.y & 0xFFFFu) << 20) |
                                     ((unsigned long long)(blockDim.x & 0xFFFu) << 36) | ((unsigned long long)(line & 0xFFFF) << 48);
        }
    }
}

// phase marks inside a kernel (timeline trace only): CTA (0,0,0), thread 0 appends {0, now, 0xF000 | id in the line field}
__device__ __forceinline__ void trace_mark(int id) {
    if (c_trace.buf != nullptr && threadIdx.x == 0 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) {
        const unsigned int i = atomicAdd(c_trace.counter, 1u);
        if (i < c_trace.cap) { c_trace.buf[3 * i] = 0; c_trace.buf[3 * i + 1] = global_timer_ns(); c_trace.buf[3 * i + 2] = (unsigned long long)(0xF000 | (id & 0xFFF)) << 48; }
    }
}
__device__ __forceinline__ void pdl_wait_at(int line, unsigned long long t0) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (c_trace.buf != nullptr && threadIdx.x == 0 && (blockIdx.x | blockIdx.y | blockIdx.z) == 0) {
        const unsigned int i = atomicAdd(c_trace.counter, 1u);
        if (i < c_trace.cap) {
            c_trace.buf[3 * i] = t0; c_trace.buf[3 * i + 1] = global_timer_ns();
            c_trace.buf[3 * i + 2] = (unsigned long long)(gridDim.x & 0xFFFFFu) | ((unsigned long long)(gridDim.y & 0xFFFFu) << 20) |
                                     ((unsigned long long)(blockDim.x & 0xFFFu) << 36) | ((unsigned long long)(line & 0xFFFF) << 48);
        }
    }
}

}  // namespace sb200
