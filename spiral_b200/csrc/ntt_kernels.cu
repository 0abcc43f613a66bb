// ntt_kernels.cu - twiddle tables + the batched MatPoly primitives built on the CTA-level NTT.
#include "kernels.cuh"
#include "ntt.cuh"
#include <algorithm>
#include <atomic>
#include <cstring>
#include <string>
#include <cstdlib>
#include <mutex>
#include <vector>

namespace sb200 {

// --------------------------------------------------------------------------------------------
// host-side table generation (regenerates reference src/constants.cpp:16 from psi; see SURVEY
// appendix A: fwd[bitrev11(i)] = psi^i, inverse rows carry psi^-i (the reference folds 1/2 into
// each inverse twiddle, we fold N^-1 into the last stage instead - identical modulo q).
// --------------------------------------------------------------------------------------------
static uint64_t h_powmod(uint64_t a, uint64_t e, uint64_t q) {
    unsigned __int128 r = 1, x = a % q;
    while (e) { if (e & 1) r = r * x % q; x = x * x % q; e >>= 1; }
    return (uint64_t)r;
}
static uint32_t h_bitrev11(uint32_t x) {
    uint32_t r = 0;
    for (int i = 0; i < kLogN; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
static uint2 h_shoup(uint64_t w, uint64_t q) {
    return make_uint2((uint32_t)w, (uint32_t)((((unsigned __int128)w) << 32) / q));
}

static std::atomic<uint64_t> g_launch_count{0};
void count_launch(int n) { g_launch_count.fetch_add((uint64_t)n, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launch_count.load(); }

// distinct kernel names launched since the last reset: call sites pass string literals, so a pointer table is enough
static std::mutex g_names_mutex;
static std::vector<const char *> g_names;
void note_kernel(const char *name) {
    static thread_local const char *last = nullptr;
    if (name == last) return;
    last = name;
    std::lock_guard<std::mutex> lock(g_names_mutex);
    for (const char *n : g_names) if (n == name) return;
    g_names.push_back(name);
}
size_t kernel_log(char *buf, size_t cap) {       // ';'-separated (template instances contain commas), sorted; returns the length needed
    std::vector<std::string> v;
    { std::lock_guard<std::mutex> lock(g_names_mutex); for (const char *n : g_names) v.emplace_back(n); }
    for (auto &n : v) { while (!n.empty() && (n.front() == '(' || n.front() == ' ')) n.erase(n.begin()); while (!n.empty() && (n.back() == ')' || n.back() == ' ')) n.pop_back(); }
    std::sort(v.begin(), v.end());
    v.erase(std::unique(v.begin(), v.end()), v.end());
    std::string out;
    for (size_t i = 0; i < v.size(); i++) { if (i) out += ";"; out += v[i]; }
    if (buf && cap) { size_t n = std::min(cap - 1, out.size()); memcpy(buf, out.data(), n); buf[n] = 0; }
    return out.size() + 1;
}
void kernel_log_reset() { std::lock_guard<std::mutex> lock(g_names_mutex); g_names.clear(); }

// timeline trace (profiling): when enabled, CTA 0 of every kernel appends {time it was scheduled, time its dependencies were
// resolved (griddepcontrol.wait returned), grid/block shape} - the dependency-resolved times of consecutive kernels give
// the cost of each link of a launch chain INSIDE a replayed graph, which ncu (serialised, cold) cannot show
static unsigned long long *g_trace_dev = nullptr;
static unsigned int g_trace_cap = 0;
int trace_enable(unsigned int capacity) {
    TraceCtl ctl{};
    if (g_trace_dev) { cudaFree(g_trace_dev); g_trace_dev = nullptr; g_trace_cap = 0; }
    if (capacity) {
        if (cudaMalloc(&g_trace_dev, (size_t)capacity * 3 * 8 + 8) != cudaSuccess) return -1;
        cudaMemset(g_trace_dev, 0, (size_t)capacity * 3 * 8 + 8);
        g_trace_cap = capacity;
        ctl.buf = g_trace_dev + 1; ctl.counter = reinterpret_cast<unsigned int *>(g_trace_dev); ctl.cap = capacity;
    }
    cudaDeviceSynchronize();
    return cudaMemcpyToSymbol(c_trace, &ctl, sizeof ctl) == cudaSuccess ? 0 : -1;
}
size_t trace_read(unsigned long long *out, size_t max_records, int reset) {
    if (!g_trace_dev) return 0;
    cudaDeviceSynchronize();
    unsigned int n = 0;
    cudaMemcpy(&n, g_trace_dev, 4, cudaMemcpyDeviceToHost);
    size_t m = std::min<size_t>(std::min<size_t>(n, g_trace_cap), max_records);
    if (out && m) cudaMemcpy(out, g_trace_dev + 1, m * 3 * 8, cudaMemcpyDeviceToHost);
    if (reset) cudaMemset(g_trace_dev, 0, 8);
    return m;
}

int &launch_priority() { static thread_local int p = kNoPriority; return p; }
LaunchPriority::LaunchPriority(bool high, int level) : saved(launch_priority()) {
    // measured (profiles/r02_expansion_chains.md): strict priorities starve the low-priority chain whenever the other graph has
    // nodes pending, even dependency-blocked ones - the odd chain then spills into the scan.  Off unless SB200_PRIO=1.
    static const int mode = [] { const char *e = getenv("SB200_PRIO"); return e ? atoi(e) : 0; }();
    if (mode != level && !(level == 3 && mode == 2)) return;
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) { cudaGetLastError(); return; }
    launch_priority() = high ? greatest : least;
}
bool pdl_enabled() { static int v = -1; if (v < 0) { const char *e = getenv("SB200_NO_PDL"); v = (e && *e == '1') ? 0 : 1; } return v == 1; }

static std::mutex g_tab_mutex;
static bool g_tab_ready[64] = {false};

int init_tables() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    std::lock_guard<std::mutex> lock(g_tab_mutex);
    if (dev < 64 && g_tab_ready[dev]) return 0;
    NttTables h;
    uint2 fwd_head[2][16], inv_head[2][16];
    for (int n = 0; n < 2; n++) {
        const uint64_t q = n == 0 ? kP : kB, psi = n == 0 ? kPsiP : kPsiB;
        const uint64_t psi_inv = h_powmod(psi, q - 2, q);
        std::vector<uint2> fwd(kN), inv(kN);
        unsigned __int128 a = 1, b = 1;
        for (uint32_t i = 0; i < (uint32_t)kN; i++) {
            uint32_t br = h_bitrev11(i);
            fwd[br] = h_shoup((uint64_t)a, q);
            inv[br] = h_shoup((uint64_t)b, q);
            a = a * psi % q;
            b = b * psi_inv % q;
        }
        const uint64_t ninv = h_powmod(kN, q - 2, q);
        h.ninv[n] = h_shoup(ninv, q);
        h.inv_last[n] = h_shoup((uint64_t)((unsigned __int128)inv[1].x * ninv % q), q);
        for (int i = 0; i < 16; i++) { fwd_head[n][i] = fwd[i]; inv_head[n][i] = inv[i]; }
        uint2 *d_fwd = nullptr, *d_inv = nullptr;
        if (cudaMalloc(&d_fwd, kN * sizeof(uint2)) != cudaSuccess) return -2;
        if (cudaMalloc(&d_inv, kN * sizeof(uint2)) != cudaSuccess) return -2;
        cudaMemcpy(d_fwd, fwd.data(), kN * sizeof(uint2), cudaMemcpyHostToDevice);
        cudaMemcpy(d_inv, inv.data(), kN * sizeof(uint2), cudaMemcpyHostToDevice);
        h.fwd[n] = d_fwd;
        h.inv[n] = d_inv;
    }
    if (cudaMemcpyToSymbol(c_ntt, &h, sizeof(h)) != cudaSuccess) return -3;
    if (cudaMemcpyToSymbol(c_fwd_head, fwd_head, sizeof(fwd_head)) != cudaSuccess) return -3;
    if (cudaMemcpyToSymbol(c_inv_head, inv_head, sizeof(inv_head)) != cudaSuccess) return -3;
    if (cudaDeviceSynchronize() != cudaSuccess) return -4;
    if (dev < 64) g_tab_ready[dev] = true;
    return 0;
}

// --------------------------------------------------------------------------------------------
// boundary format conversion
// --------------------------------------------------------------------------------------------
__global__ void k_ntt_u64_to_dev(uint32_t *__restrict__ out, const uint64_t *__restrict__ in, size_t nwords) {
    pdl_prologue();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    int n = (int)((i / kN) & 1);
    out[i] = reduce_u64(in[i], n);       // the reference may hand over q for 0 (and lazy digits) - canonicalise
}
__global__ void k_ntt_dev_to_u64(uint64_t *__restrict__ out, const uint32_t *__restrict__ in, size_t nwords) {
    pdl_prologue();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nwords) out[i] = in[i];
}
void launch_ntt_u64_to_dev(uint32_t *out, const uint64_t *in, size_t npolys, cudaStream_t s) {
    size_t n = npolys * 2 * kN;
    if (n) { count_launch(); launch_pdl(k_ntt_u64_to_dev, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, out, in, n); }
}
void launch_ntt_dev_to_u64(uint64_t *out, const uint32_t *in, size_t npolys, cudaStream_t s) {
    size_t n = npolys * 2 * kN;
    if (n) { count_launch(); launch_pdl(k_ntt_dev_to_u64, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, out, in, n); }
}

// --------------------------------------------------------------------------------------------
// to_ntt: raw u64 -> residues -> forward NTT (reference src/poly.cpp:291-329; the "no_reduce"
// variant is the same map modulo q).  One CTA per polynomial.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNttThreads) k_to_ntt(uint32_t *__restrict__ out, const uint64_t *__restrict__ raw) {
    pdl_prologue();
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    const int n = plane_of_thread(), lt = lane_in_plane();
    const uint64_t *src = raw + (size_t)blockIdx.x * kN;
    uint32_t v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = raw_to_res(__ldg(src + nat_pos(lt, k)), n);
    ntt_forward_plane(v, sm[n], lt, n);
    store_ntt_regs(v, out + ((size_t)blockIdx.x * 2 + n) * kN, lt);
}
void launch_to_ntt(uint32_t *out, const uint64_t *raw, size_t npolys, cudaStream_t s) {
    if (npolys) { count_launch(); launch_pdl(k_to_ntt, dim3((unsigned)npolys), dim3(kNttThreads), 0, s, out, raw); }
}

// --------------------------------------------------------------------------------------------
// from_ntt: inverse NTT under both primes + CRT lift to [0,Q) (reference src/poly.cpp:357-377,
// nttInvAndCrtLiftCiphertexts src/spiral.cpp:437-453).  One CTA per polynomial.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void intt_crt_store(uint32_t (&v)[16], uint32_t (*sm)[kPlaneWords], uint64_t *__restrict__ dst) {
    const int n = plane_of_thread(), lt = lane_in_plane();
    ntt_inverse_plane(v, sm[n], lt, n);
    __syncthreads();                                  // both planes done reading sm
#pragma unroll
    for (int k = 0; k < 16; k++) sm[n][nat_pos(lt, k)] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
        int z = threadIdx.x + 256 * k;
        dst[z] = crt_compose(sm[0][z], sm[1][z]);
    }
}
__global__ void __launch_bounds__(kNttThreads) k_from_ntt(uint64_t *__restrict__ raw, const uint32_t *__restrict__ in) {
    pdl_begin();
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    const int n = plane_of_thread(), lt = lane_in_plane();
    prefetch_twiddles(c_ntt.inv[n], lt);
    pdl_wait();
    uint32_t v[16];
    load_ntt_regs(v, in + ((size_t)blockIdx.x * 2 + n) * kN, lt);
    intt_crt_store(v, sm, raw + (size_t)blockIdx.x * kN);
}
void launch_from_ntt(uint64_t *raw, const uint32_t *in, size_t npolys, cudaStream_t s) {
    if (npolys) { count_launch(); launch_pdl(k_from_ntt, dim3((unsigned)npolys), dim3(kNttThreads), 0, s, raw, in); }
}

// --------------------------------------------------------------------------------------------
// multiply / add (reference src/poly.cpp:34-78, :138-155): pointwise over z, 4 coefficients / thread
// --------------------------------------------------------------------------------------------
__global__ void k_matmul(uint32_t *__restrict__ out, const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                         int rs, int ms, int cs) {
    pdl_prologue();
    // blockIdx.x = (r*cs + c)*4 + quarter ; 256 threads x uint4 = 1024 words = quarter of a dev-NTT poly
    const int rc = blockIdx.x >> 2, quarter = blockIdx.x & 3;
    const int r = rc / cs, c = rc % cs;
    const int w4 = quarter * 256 + threadIdx.x;                 // uint4 index inside the poly (0..1023)
    const int n = w4 >= 512;
    uint64_t acc[4] = {0, 0, 0, 0};
    for (int m = 0; m < ms; m++) {
        const uint4 x = __ldg(reinterpret_cast<const uint4 *>(a + (size_t)(r * ms + m) * 2 * kN) + w4);
        const uint4 y = __ldg(reinterpret_cast<const uint4 *>(b + (size_t)(m * cs + c) * 2 * kN) + w4);
        acc[0] += (uint64_t)x.x * y.x; acc[1] += (uint64_t)x.y * y.y;
        acc[2] += (uint64_t)x.z * y.z; acc[3] += (uint64_t)x.w * y.w;
        if ((m & 127) == 127) {
#pragma unroll
            for (int e = 0; e < 4; e++) acc[e] = reduce_u64(acc[e], n);
        }
    }
    uint4 o = make_uint4(reduce_u64(acc[0], n), reduce_u64(acc[1], n), reduce_u64(acc[2], n), reduce_u64(acc[3], n));
    reinterpret_cast<uint4 *>(out + (size_t)rc * 2 * kN)[w4] = o;
}
void launch_matmul(uint32_t *out, const uint32_t *a, const uint32_t *b, int rs, int ms, int cs, cudaStream_t s) {
    if (rs * cs) { count_launch(); launch_pdl(k_matmul, dim3(rs * cs * 4), dim3(256), 0, s, out, a, b, rs, ms, cs); }
}

__global__ void k_add(uint32_t *__restrict__ out, const uint32_t *__restrict__ a, const uint32_t *__restrict__ b, size_t nwords) {
    pdl_prologue();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nwords) return;
    const uint32_t q = modulus((int)((i / kN) & 1));
    out[i] = csub(a[i] + b[i], q);
}
void launch_add(uint32_t *out, const uint32_t *a, const uint32_t *b, size_t npolys, cudaStream_t s) {
    size_t n = npolys * 2 * kN;
    if (n) { count_launch(); launch_pdl(k_add, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, out, a, b, n); }
}

// --------------------------------------------------------------------------------------------
// automorph (reference src/poly.cpp:240-261): x -> x^t on raw coefficients; negated entries are
// Q - a, so a == 0 becomes Q (kept on purpose: the value is later gadget-decomposed as Q).
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ void automorph_target(int i, uint32_t t, int &rem, bool &neg) {
    uint32_t it = (uint32_t)i * t;
    rem = (int)(it & (kN - 1));
    neg = (it >> kLogN) & 1;
}
__global__ void k_automorph(uint64_t *__restrict__ out, const uint64_t *__restrict__ in, uint32_t t) {
    pdl_prologue();
    const uint64_t *src = in + (size_t)blockIdx.x * kN;
    uint64_t *dst = out + (size_t)blockIdx.x * kN;
    for (int i = threadIdx.x; i < kN; i += blockDim.x) {
        int rem; bool neg;
        automorph_target(i, t, rem, neg);
        uint64_t a = src[i];
        dst[rem] = neg ? kQ - a : a;
    }
}
void launch_automorph(uint64_t *out, const uint64_t *in, size_t npolys, uint32_t t, cudaStream_t s) {
    if (npolys) { count_launch(); launch_pdl(k_automorph, dim3((unsigned)npolys), dim3(256), 0, s, out, in, t); }
}

// --------------------------------------------------------------------------------------------
// gadget_invert (+ to_ntt): unsigned base-2^bits_per digits (reference src/util.cpp:114-150).
// in : raw (rdim x cols); out: (mx x cols), digit k of input row j lands in row j + k*rdim.
// grid = (cols, mx): blockIdx.y = output row.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNttThreads) k_gadget_ntt(uint32_t *__restrict__ out, const uint64_t *__restrict__ raw,
                                                            int mx, int rdim, int cols) {
    pdl_begin();
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    const int n = plane_of_thread(), lt = lane_in_plane();
    prefetch_twiddles(c_ntt.fwd[n], lt);
    pdl_wait();
    const int col = blockIdx.x, row = blockIdx.y;
    const int j = row % rdim, k = row / rdim;
    const uint32_t bits_per = get_bits_per(mx / rdim);
    const uint64_t mask = (1ull << bits_per) - 1;
    const uint64_t *src = raw + ((size_t)j * cols + col) * kN;
    uint32_t v[16];
#pragma unroll
    for (int e = 0; e < 16; e++) {
        uint64_t d = gadget_digit(__ldg(src + nat_pos(lt, e)), k, bits_per, mask);
        v[e] = bits_per >= 28 ? raw_to_res(d, n) : (uint32_t)d;     // digits < 2^28 are already < 4q
    }
    ntt_forward_plane(v, sm[n], lt, n);
    store_ntt_regs(v, out + (((size_t)row * cols + col) * 2 + n) * kN, lt);
}
void launch_gadget_ntt(uint32_t *out, const uint64_t *raw, int mx, int rdim, int cols, cudaStream_t s) {
    if (mx * cols) { count_launch(); launch_pdl(k_gadget_ntt, dim3(dim3(cols, mx)), dim3(kNttThreads), 0, s, out, raw, mx, rdim, cols); }
}
__global__ void k_gadget_raw(uint64_t *__restrict__ out, const uint64_t *__restrict__ raw, int mx, int rdim, int cols) {
    pdl_prologue();
    const int col = blockIdx.x, row = blockIdx.y;
    const int j = row % rdim, k = row / rdim;
    const uint32_t bits_per = get_bits_per(mx / rdim);
    const uint64_t mask = (1ull << bits_per) - 1;
    const uint64_t *src = raw + ((size_t)j * cols + col) * kN;
    uint64_t *dst = out + ((size_t)row * cols + col) * kN;
    for (int z = threadIdx.x; z < kN; z += blockDim.x) dst[z] = gadget_digit(src[z], k, bits_per, mask);
}
void launch_gadget_raw(uint64_t *out, const uint64_t *raw, int mx, int rdim, int cols, cudaStream_t s) {
    if (mx * cols) { count_launch(); launch_pdl(k_gadget_raw, dim3(dim3(cols, mx)), dim3(256), 0, s, out, raw, mx, rdim, cols); }
}

// --------------------------------------------------------------------------------------------
// rescale / getRescaled (reference src/poly.cpp:578-601): modulus switch with C truncation and
// sign-dependent rounding.  inp_mod <= 2^57, out_mod < 2^36: the product needs 128-bit signed
// arithmetic, done here on the magnitude with an exact restoring division.
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t udiv128_64(uint64_t hi, uint64_t lo, uint64_t d) {   // (hi:lo) / d, hi < d
    uint64_t qt = 0, r = hi;
#pragma unroll 1
    for (int i = 63; i >= 0; i--) {
        const uint64_t top = r >> 63;
        r = (r << 1) | ((lo >> i) & 1);
        if (top || r >= d) { r -= d; qt |= (1ull << i); }
    }
    return qt;
}
__device__ __forceinline__ uint64_t rescale_one(uint64_t a, uint64_t inp_mod, uint64_t out_mod) {
    int64_t inp_val = (int64_t)(a % inp_mod);
    if (inp_val >= (int64_t)(inp_mod / 2)) inp_val -= (int64_t)inp_mod;
    const bool negv = inp_val < 0;
    const uint64_t mag = negv ? (uint64_t)(-inp_val) : (uint64_t)inp_val;
    // |val + sign*(inp_mod/2)| = mag*out_mod + inp_mod/2 ; C division truncates toward zero, so the
    // quotient's magnitude is floor(that / inp_mod) and its sign is the sign of inp_val
    uint64_t lo = mag * out_mod, hi = __umul64hi(mag, out_mod);
    const uint64_t half = inp_mod / 2;
    lo += half; hi += (lo < half);
    const uint64_t qmag = udiv128_64(hi, lo, inp_mod);          // < out_mod + 1
    // result = (+-qmag + (inp_mod/out_mod)*out_mod + 2*out_mod) % out_mod, then (+out_mod) % out_mod
    const uint64_t r = qmag % out_mod;
    return (negv && r != 0) ? out_mod - r : r;
}
__global__ void k_rescale(uint64_t *__restrict__ out, const uint64_t *__restrict__ in, size_t n, uint64_t inp_mod, uint64_t out_mod) {
    pdl_prologue();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = rescale_one(in[i] % kQ, inp_mod, out_mod);
}
// the response's two modulus switches in one launch: the first n0 coefficients to mod0 (row 0 -> q'), the next n1 to mod1 (-> 4p)
__global__ void k_rescale2(uint64_t *__restrict__ out, const uint64_t *__restrict__ in, size_t n0, size_t n1, uint64_t inp_mod, uint64_t mod0, uint64_t mod1) {
    pdl_prologue();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n0 + n1) out[i] = rescale_one(in[i] % kQ, inp_mod, i < n0 ? mod0 : mod1);
}
void launch_rescale2(uint64_t *out, const uint64_t *in, size_t n0, size_t n1, uint64_t inp_mod, uint64_t mod0, uint64_t mod1, cudaStream_t s) {
    if (n0 + n1) { count_launch(); launch_pdl(k_rescale2, dim3((unsigned)((n0 + n1 + 255) / 256)), dim3(256), 0, s, out, in, n0, n1, inp_mod, mod0, mod1); }
}
void launch_rescale(uint64_t *out, const uint64_t *in, size_t ncoeffs, uint64_t inp_mod, uint64_t out_mod, cudaStream_t s) {
    if (ncoeffs) { count_launch(); launch_pdl(k_rescale, dim3((unsigned)((ncoeffs + 255) / 256)), dim3(256), 0, s, out, in, ncoeffs, inp_mod, out_mod); }
}

// --------------------------------------------------------------------------------------------
// modswitch (reference src/spiral.cpp:40-78) and bit packing (write_arbitrary_bits, src/core.cpp:32-52).
// The reference evaluates round((long double)v * q' / Q) in x87 extended precision: the product (up to 93 bits) is
// rounded to a 64-bit significand, the quotient again, then roundl.  Restated in integers so the GPU agrees bit for
// bit:  P' = RNE64(v * q'),  I = floor(P'/Q),  R = P' mod Q,  f = 64 - bitlen(I);
//       result = I + 1  iff  2R >= Q  or  (I > 0 and Q - 2R <= floor(Q / 2^f))      (derivation: oracle/spiral_oracle.c)
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t modswitch_one(uint64_t val, uint64_t qp) {
    uint64_t lo = val * qp, hi = __umul64hi(val, qp);
    if (hi) {                                               // first rounding: 64 significant bits, ties to even
        const int s = 64 - __clzll((long long)hi);          // 1 <= s <= 37
        uint64_t m = (hi << (64 - s)) | (lo >> s);
        const uint64_t rem = lo & ((1ull << s) - 1), half = 1ull << (s - 1);
        if (rem > half || (rem == half && (m & 1))) m++;
        if (m == 0) { hi = 1ull << s; lo = 0; }             // significand overflowed to 2^64
        else { hi = m >> (64 - s); lo = m << s; }
    }
    const uint64_t I = udiv128_64(hi, lo, kQ);              // hi < 2^38 < Q
    const uint64_t R = lo - I * kQ;
    bool up;
    if (2 * R >= kQ) up = true;
    else if (I == 0) up = false;
    else up = (kQ - 2 * R) <= (kQ >> __clzll((long long)I));
    return I + (up ? 1 : 0);
}
// One thread per OUTPUT word: gathers the (at most 64/bits + 2) values overlapping it.  MODE 0: plain values,
// MODE 1: modswitch_one(in[i], qp) computed on the fly.
template <int MODE>
__global__ void k_bitpack(uint64_t *__restrict__ out, const uint64_t *__restrict__ in, size_t n, uint32_t bits, uint64_t qp) {
    pdl_prologue();
    const size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t nwords = (n * bits + 63) / 64;
    if (w >= nwords) return;
    const uint64_t mask = bits >= 64 ? ~0ull : (1ull << bits) - 1;
    const size_t lo_bit = w * 64;
    size_t i = lo_bit / bits;
    uint64_t word = 0;
    for (; i < n && i * bits < lo_bit + 64; i++) {
        uint64_t v = in[i];
        if (MODE == 1) v = modswitch_one(v, qp);
        v &= mask;
        const size_t b = i * bits;
        word |= b >= lo_bit ? v << (b - lo_bit) : v >> (lo_bit - b);
    }
    out[w] = word;
}
void launch_bitpack(uint64_t *out, const uint64_t *in, size_t n, uint32_t bits, cudaStream_t s) {
    const size_t nwords = (n * bits + 63) / 64;
    if (nwords) { count_launch(); launch_pdl((k_bitpack<0>), dim3((unsigned)((nwords + 127) / 128)), dim3(128), 0, s, out, in, n, bits, (uint64_t)0); }
}
void launch_modswitch(uint64_t *out, const uint64_t *in_raw, size_t n, uint32_t bits, uint64_t qprime, cudaStream_t s) {
    const size_t nwords = (n * bits + 63) / 64;
    if (nwords) { count_launch(); launch_pdl((k_bitpack<1>), dim3((unsigned)((nwords + 127) / 128)), dim3(128), 0, s, out, in_raw, n, bits, qprime); }
}

}  // namespace sb200
