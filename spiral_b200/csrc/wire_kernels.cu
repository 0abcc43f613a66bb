// wire_kernels.cu - wire and on-disk formats of the query path.
//
// The reference keeps everything in process: it only ACCOUNTS for the query / response sizes
// (print_summary, src/spiral.cpp:219-234, b_per_elem = poly_len * logQ / 8) and leaves the database file
// I/O as `// TODO` (load_db, src/spiral.cpp:1095-1162).  The formats below are what a client/server split
// needs; they are stated independently in oracle/wire_format.c and built on the reference's bit order
// (read/write_arbitrary_bits, src/core.cpp:20-52: value i occupies bits [i*bits, (i+1)*bits) of a
// little-endian word stream).
//
//   query, SEEDED : "SB2Q" | kind=1 | 32-byte seed | row 1 at 56 bits per coefficient          (14 376 B)
//   query, FULL   : "SB2Q" | kind=2 | row 0 | row 1, both at 56 bits per coefficient            (28 680 B)
//   records       : flat bit stream, log2(p_db) bits per plaintext coefficient, item-major, polynomial-major
//                   inside an item (Spiral item = 2x2 polynomials, Pack item = out_n^2 polynomials)
//
// A query is the 2x1 Regev ciphertext of getRegevSample (src/client.cpp:141-157): row 0 = -a is uniformly
// random, so a client may derive it from a seed - directly in NTT form, one ChaCha20 block (RFC 8439) per
// NTT slot - and send only row 1.  The server regenerates row 0 into the same dev-NTT buffer the expansion
// reads; row 1 is unpacked and transformed by the same CTA-level NTT as everything else.
#include "kernels.cuh"
#include "ntt.cuh"

namespace sb200 {

constexpr uint32_t kWireQueryMagic = 0x51324253u;      // "SB2Q"
constexpr uint32_t kWireSeeded = 1, kWireFull = 2;
constexpr size_t kWireHeaderBytes = 8, kWireSeedBytes = 32, kWireRowBytes = (size_t)kN * 56 / 8;

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }   // one SHF on the device
__host__ __device__ __forceinline__ void chacha_qr(uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
    a += b; d ^= a; d = rotl32(d, 16);
    c += d; b ^= c; b = rotl32(b, 12);
    a += b; d ^= a; d = rotl32(d, 8);
    c += d; b ^= c; b = rotl32(b, 7);
}
// RFC 8439 section 2.3 block function (also used on the host: the client derives per-query wire seeds with it)
__host__ __device__ __forceinline__ void chacha20_block(uint32_t (&x)[16], const uint32_t (&key)[8], uint32_t counter, uint32_t n0, uint32_t n1, uint32_t n2) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key[0], key[1], key[2], key[3],
                      key[4], key[5], key[6], key[7], counter, n0, n1, n2};
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = s[i];
#pragma unroll
    for (int r = 0; r < 10; r++) {
        chacha_qr(x[0], x[4], x[8], x[12]); chacha_qr(x[1], x[5], x[9], x[13]); chacha_qr(x[2], x[6], x[10], x[14]); chacha_qr(x[3], x[7], x[11], x[15]);
        chacha_qr(x[0], x[5], x[10], x[15]); chacha_qr(x[1], x[6], x[11], x[12]); chacha_qr(x[2], x[7], x[8], x[13]); chacha_qr(x[3], x[4], x[9], x[14]);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] += s[i];
}
// uniform residue below q from one block: the FIRST of the 16 words whose low 28 bits are < q (all rejected, probability < 2^-60:
// the last word reduced)
__device__ __forceinline__ uint32_t uniform_from_block(const uint32_t (&x)[16], uint32_t q) {
    uint32_t v = (x[15] & 0x0FFFFFFFu) % q;
#pragma unroll
    for (int k = 15; k >= 0; k--) { const uint32_t c = x[k] & 0x0FFFFFFFu; if (c < q) v = c; }
    return v;
}

// One launch turns a wire query into cv[0] (dev-NTT, [row][prime][2048]).
//   CTA 0                 : row 1 - 14 336 packed bytes -> shared memory -> 56-bit coefficients -> forward NTT under both primes
//   CTA 1 (FULL)          : row 0 the same way
//   CTAs 1..16 (SEEDED)   : row 0 - one ChaCha20 block per (prime, slot); first of the 16 words whose low 28 bits are < q
__global__ void __launch_bounds__(kNttThreads) k_query_from_wire(uint32_t *__restrict__ cv, const uint8_t *__restrict__ wire, uint32_t kind) {
    pdl_prologue();
    __shared__ __align__(16) uint64_t packed[kWireRowBytes / 8];
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    if (kind == kWireSeeded && blockIdx.x > 0) {
        const int slot = (blockIdx.x - 1) * kNttThreads + threadIdx.x;       // n * 2048 + z
        const int n = slot >> kLogN;
        const uint32_t q = modulus(n);
        const uint32_t *seed = reinterpret_cast<const uint32_t *>(wire + kWireHeaderBytes);
        uint32_t key[8], x[16];
#pragma unroll
        for (int i = 0; i < 8; i++) key[i] = __ldg(seed + i);
        chacha20_block(x, key, (uint32_t)slot, kWireQueryMagic, 0u, 0u);
        cv[slot] = uniform_from_block(x, q);                                   // row 0 = polynomial 0: [n][z]
        return;
    }
    const int row = blockIdx.x == 0 ? 1 : 0;
    const size_t off = kind == kWireSeeded ? kWireHeaderBytes + kWireSeedBytes : kWireHeaderBytes + (size_t)row * kWireRowBytes;
    const uint64_t *src = reinterpret_cast<const uint64_t *>(wire + off);
    for (int i = threadIdx.x; i < (int)(kWireRowBytes / 8); i += kNttThreads) packed[i] = __ldg(src + i);
    __syncthreads();
    const int n = plane_of_thread(), lt = lane_in_plane();
    uint32_t v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const uint32_t bit = (uint32_t)nat_pos(lt, k) * 56u, w = bit >> 6, o = bit & 63u;
        uint64_t val = packed[w] >> o;
        if (o > 8) val |= packed[w + 1] << (64 - o);
        v[k] = raw_to_res(val & ((1ull << 56) - 1), n);
    }
    ntt_forward_plane(v, sm[n], lt, n);
    store_ntt_regs(v, cv + ((size_t)row * 2 + n) * kN, lt);
}
void launch_query_from_wire(uint32_t *cv, const uint8_t *wire, uint32_t kind, cudaStream_t s) {
    const unsigned grid = kind == kWireSeeded ? 1 + 2 * kN / kNttThreads : 2;
    count_launch(); launch_pdl(k_query_from_wire, dim3(grid), dim3(kNttThreads), 0, s, cv, wire, kind);
}

// records -> plaintext coefficients.  Staged records hold n_items whole items of `polys` polynomials; coefficient
// (item, poly, z) goes to out[item * out_item_stride + poly * out_poly_stride + z] (Spiral: item-major chunks for
// k_db_build_spiral; Pack: one plane-major array for k_db_build_pack).  `rec` is padded by 4 readable bytes.
__global__ void k_records_to_pts(uint16_t *__restrict__ out, const uint8_t *__restrict__ rec, size_t n_items, int polys, uint32_t bits,
                                 size_t out_item_stride, size_t out_poly_stride) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_items * polys * kN) return;
    const size_t z = idx & (kN - 1), ip = idx >> kLogN, poly = ip % polys, item = ip / polys;
    const size_t bit = idx * bits, byte = bit >> 3;
    const uint32_t w = (uint32_t)rec[byte] | ((uint32_t)rec[byte + 1] << 8) | ((uint32_t)rec[byte + 2] << 16);
    out[item * out_item_stride + poly * out_poly_stride + z] = (uint16_t)((w >> (bit & 7)) & ((1u << bits) - 1));
}
void launch_records_to_pts(uint16_t *out, const uint8_t *rec, size_t n_items, int polys, uint32_t bits, size_t out_item_stride,
                           size_t out_poly_stride, cudaStream_t s) {
    const size_t n = n_items * polys * kN;
    if (!n) return;
    count_launch();
    note_kernel("k_records_to_pts"); k_records_to_pts<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(out, rec, n_items, polys, bits, out_item_stride, out_poly_stride);
}

// integrity word of a database snapshot: sum of all 64-bit words modulo 2^64
__global__ void k_sum64(unsigned long long *__restrict__ acc, const uint64_t *__restrict__ words, size_t n) {
    uint64_t s = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) s += words[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(acc, (unsigned long long)s);
}
void launch_sum64(unsigned long long *acc, const uint64_t *words, size_t n, cudaStream_t s) {
    count_launch();
    note_kernel("k_sum64"); k_sum64<<<148 * 8, 256, 0, s>>>(acc, words, n);
}

}  // namespace sb200
