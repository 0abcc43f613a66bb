// client_kernels.cu - the Spiral CLIENT on the GPU: key generation, public parameters, query
// generation and response decoding - the 0.3-1.4 s client-side costs of the reference's summaries, and its only use of
// Intel HEXL (decoding, src/util.cpp:231-241).
//
// Reference statements replaced:
//   keygen                         src/client.cpp:23-47         -> k_client_gauss_raw
//   getRegevSample / encryptSimpleRegev(Matrix)   :141-227      -> k_client_regev_cols (R = 1)
//   get_fresh_public_key_raw       src/client.cpp:49-68         -> k_client_regev_cols (R = 2)
//   getPublicEncryptions           src/client.cpp:271-290       -> host loop over rounds (api_client.cu)
//   W / V generation               src/spiral.cpp:2207-2290     -> the same kernel with gadget-scaled messages
//   query encoding                 src/spiral.cpp:2098-2157     -> sparse sigma on the host + the same kernel
//   check_final decode             src/spiral.cpp:1428-1476 (to_ntt_qprime / mul_over_qprime / from_ntt_qprime) -> k_client_decode
//
// The reference draws from an unseeded std::random_device, so no two runs agree; here every random polynomial is a
// ChaCha20 stream named by an object id (nonce = {"SB2C", object, stream}), which lets 256-thread CTAs sample
// independently and lets oracle/client_sim.c (so_client_new_chacha) restate the client bit for bit.  Uniform
// polynomials are drawn directly in NTT form (the NTT + CRT map is a bijection on uniform values), so a column
// costs ONE forward NTT (of its noise) instead of three.
#include "kernels.cuh"
#include "ntt.cuh"

namespace sb200 {

constexpr uint32_t kClientMagic = 0x43324253u;        // "SB2C"
__constant__ uint64_t c_gauss_thr[128];               // floor(cdf[k] * 2^53) + 1 (discrete Gaussian, width 6.4, src/core.cpp:182-207)

struct ClientKey { uint32_t w[8]; };

// Gaussian coefficient z of (object, stream): U = top 53 bits of block words (1,0); value = -64 + #{k : U >= thr[k]}
__device__ __forceinline__ int gauss_coeff(const ClientKey &key, uint32_t obj, uint32_t sub, uint32_t z) {
    uint32_t x[16];
    chacha20_block(x, key.w, z, kClientMagic, obj, sub);
    const uint64_t U = (((uint64_t)x[1] << 32) | x[0]) >> 11;
    int v = -64;
#pragma unroll 8
    for (int k = 0; k < 128; k++) v += (U >= c_gauss_thr[k]) ? 1 : 0;
    return v;
}

// raw Gaussian polynomials (keygen): out[poly][z] in [0,Q), object = obj_base + poly
__global__ void k_client_gauss_raw(uint64_t *__restrict__ out, ClientKey key, uint32_t obj_base, uint32_t sub) {
    const uint32_t z = blockIdx.x * blockDim.x + threadIdx.x, poly = blockIdx.y;
    const int v = gauss_coeff(key, obj_base + poly, sub, z);
    out[(size_t)poly * kN + z] = v < 0 ? kQ - (uint64_t)(-v) : (uint64_t)v;
}

// One CTA = one column of an (1 + R) x cols matrix of Regev encryptions under the R secret rows S:
//   row 0     = -a, uniform in NTT form: slot (n, z) from block(counter = n*2048 + z, nonce {"SB2C", obj, 0}) - or, for a wire
//               query, from block(key = wire seed, nonce {"SB2Q", 0, 0}) as k_query_from_wire regenerates it
//   row 1 + r = a * S_r + NTT(e_r) + scal[r][col] * msg_r,   e_r = Gaussian (obj, sub_e + r)
// out: dev-NTT, polynomial (row, col) at (row * out_cols + col).
struct RegevArgs {
    ClientKey key, ukey;            // noise / uniform keys (ukey used when wire != 0)
    uint32_t obj_base;              // object of column c = obj_base + c
    uint32_t sub_e, wire;
    int R, out_cols, col_begin;
    const uint32_t *S;              // [R] dev-NTT secret rows
    const uint32_t *msg;            // [R] dev-NTT message bases (nullptr: no message)
    const uint32_t *scal;           // [R][ncols][2] residues of the per-column scalar (nullptr: 1)
    int ncols;
    int col_stride;                 // 0 / 1: consecutive columns; s: launch column c is matrix column col_begin + c*s, object obj_base + c*s
    int ct_major;                   // 0: polynomial (row, col) at row*out_cols + col (a matrix); 1: at col*(1+R) + row (a vector of ciphertexts)
};
__global__ void __launch_bounds__(kNttThreads) k_client_regev_cols(uint32_t *__restrict__ out, RegevArgs a) {
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    const int n = plane_of_thread(), lt = lane_in_plane();
    const uint32_t q = modulus(n);
    const int c = blockIdx.x, cs = a.col_stride > 1 ? a.col_stride : 1, col = a.col_begin + c * cs;
    const uint32_t obj = a.obj_base + (uint32_t)(c * cs);
    const size_t rows = 1 + (size_t)a.R;
    uint32_t row0[16];
#pragma unroll
    for (int k = 0; k < 16; k++) {
        uint32_t x[16];
        const uint32_t slot = (uint32_t)n * kN + (uint32_t)ntt_pos(lt, k);
        if (a.wire) chacha20_block(x, a.ukey.w, slot, 0x51324253u, 0u, 0u);
        else chacha20_block(x, a.key.w, slot, kClientMagic, obj, 0u);
        row0[k] = uniform_from_block(x, q);
    }
    store_ntt_regs(row0, out + ((a.ct_major ? (size_t)col * rows : (size_t)col) * 2 + n) * kN, lt);
    for (int r = 0; r < a.R; r++) {
        uint32_t e[16];
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const int v = gauss_coeff(a.key, obj, a.sub_e + (uint32_t)r, (uint32_t)nat_pos(lt, k));
            e[k] = v < 0 ? q - (uint32_t)(-v) : (uint32_t)v;
        }
        __syncthreads();                                   // the previous row's NTT is done with sm
        ntt_forward_plane(e, sm[n], lt, n);
        uint32_t s[16], m[16];
        load_ntt_regs(s, a.S + ((size_t)r * 2 + n) * kN, lt);
        uint32_t sc = 1;
        if (a.msg) {
            load_ntt_regs(m, a.msg + ((size_t)r * 2 + n) * kN, lt);
            if (a.scal) sc = a.scal[((size_t)r * a.ncols + c) * 2 + n];
        }
#pragma unroll
        for (int k = 0; k < 16; k++) {
            uint64_t acc = (uint64_t)(row0[k] ? q - row0[k] : 0u) * s[k] + e[k];
            if (a.msg) acc += (uint64_t)m[k] * sc;
            e[k] = reduce_u64(acc, n);
        }
        store_ntt_regs(e, out + ((a.ct_major ? (size_t)col * rows + 1 + r : (size_t)(1 + r) * a.out_cols + col) * 2 + n) * kN, lt);
    }
}

// Decoding (check_final): pt[r][c] = round-and-reduce of  Sp_r * resp_row0[c] (negacyclic, mod q')  combined with rows 1..2.
// grid (8, 4): blockIdx.y = r*2 + c, 256 output coefficients per CTA; the two 2048-term operands sit in shared memory.
__global__ void __launch_bounds__(256) k_client_decode(uint64_t *__restrict__ pt, const uint64_t *__restrict__ resp, const uint64_t *__restrict__ Sp_raw,
                                                       uint64_t qp, uint64_t p_db, int nn) {      // nn = n2 (Spiral: 2) or out_n (Pack variants)
    // operands stay 64-bit and the 2048 products (each < qp^2 <= 2^74 for the 37-bit moduli of qprime_mods) are summed in
    // 128 bits: exact for every q' the reference's table holds (oracle/client_sim.c does the same with __int128)
    __shared__ uint64_t sa[kN], sb[kN];
    const int r = blockIdx.y / nn, c = blockIdx.y % nn;
    for (int i = threadIdx.x; i < kN; i += 256) {
        const uint64_t raw = Sp_raw[(size_t)r * kN + i];                   // recentring of to_ntt_qprime (src/util.cpp:224-229)
        const int64_t a = raw >= kQ / 2 ? (int64_t)raw - (int64_t)kQ : (int64_t)raw;
        int64_t m = a % (int64_t)qp;
        sa[i] = (uint64_t)(m < 0 ? m + (int64_t)qp : m);
        sb[i] = resp[(size_t)c * kN + i] % qp;
    }
    __syncthreads();
    const int k = blockIdx.x * 256 + threadIdx.x;
    unsigned __int128 pos128 = 0, neg128 = 0;
#pragma unroll 8
    for (int i = 0; i < kN; i++) {
        const int j = k - i;
        const unsigned __int128 prod = (unsigned __int128)sa[i] * sb[j & (kN - 1)];
        if (j >= 0) pos128 += prod; else neg128 += prod;                    // x^N = -1
    }
    const uint64_t pos = (uint64_t)(pos128 % qp), neg = (uint64_t)(neg128 % qp);
    const uint64_t sp = (pos + qp - neg) % qp;
    const uint64_t q1 = 4 * p_db, denom = qp * (q1 / p_db);
    const uint64_t rest = resp[(size_t)(nn + r * nn + c) * kN + k];
    const int64_t vf = sp >= qp / 2 ? (int64_t)sp - (int64_t)qp : (int64_t)sp;
    const int64_t vr = rest >= q1 / 2 ? (int64_t)rest - (int64_t)q1 : (int64_t)rest;
    const int64_t rr = vf * (int64_t)q1 + vr * (int64_t)qp;
    const int64_t half = (int64_t)(denom / 2);
    int64_t res = (rr + (rr >= 0 ? half : -half)) / (int64_t)denom;       // C truncating division, as the reference
    res = (res + (int64_t)((denom / p_db) * p_db) + 2 * (int64_t)p_db) % (int64_t)p_db;
    pt[(size_t)(r * nn + c) * kN + k] = (uint64_t)res;
}

void launch_client_gauss_raw(uint64_t *out, const ClientKey &key, uint32_t obj_base, uint32_t sub, int npolys, cudaStream_t s) {
    count_launch();
    note_kernel("k_client_gauss_raw"); k_client_gauss_raw<<<dim3(kN / 256, npolys), 256, 0, s>>>(out, key, obj_base, sub);
}
void launch_client_regev_cols(uint32_t *out, const RegevArgs &a, cudaStream_t s) {
    if (a.ncols <= 0) return;
    count_launch();
    note_kernel("k_client_regev_cols"); k_client_regev_cols<<<a.ncols, kNttThreads, 0, s>>>(out, a);
}
void launch_client_decode(uint64_t *pt, const uint64_t *resp, const uint64_t *Sp_raw, uint64_t qp, uint64_t p_db, int nn, cudaStream_t s) {
    count_launch();
    note_kernel("k_client_decode"); k_client_decode<<<dim3(kN / 256, nn * nn), 256, 0, s>>>(pt, resp, Sp_raw, qp, p_db, nn);
}

}  // namespace sb200
