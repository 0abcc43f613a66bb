// api_client.cu - C-ABI of the GPU client (kernels in client_kernels.cu).  Part of the single
// translation unit spiral_b200.cu.  The client holds its secret key in HBM; public parameters, queries and decoded
// records cross the boundary as host buffers in the formats the server entries take (ref-NTT matrices, wire query).
#include <cmath>

struct sb200_client {
    sb200_params prm;
    int device = 0;
    size_t g = 0, stopround = 0, n_right = 0;
    size_t sp_rows = kN0;                       // rows of S': n0 = 2 (Spiral), out_n (Pack variants)
    bool pack = false;
    ClientKey key{};
    DBuf<uint64_t> sr_raw, Sp_raw;              // 1 and 2 polynomials, raw (small signed values mod Q)
    DBuf<uint32_t> sr_ntt, Sp_ntt;              // the same in dev-NTT form
    cudaStream_t st = nullptr;
    ~sb200_client() { if (st) cudaStreamDestroy(st); }
};

namespace {
enum { CC_KEYS = 1, CC_W_RIGHT = 2, CC_W_LEFT = 3, CC_W_CONV = 4, CC_V_CONV = 5, CC_QUERY = 6, CC_PACK_W = 7, CC_QUERY_DIRECT = 8 };
inline uint32_t cc_obj(uint32_t cls, size_t idx) { return (cls << 24) | (uint32_t)idx; }

// discrete Gaussian of width 6.4 on [-64, 64] (src/core.cpp:182-207) as integer thresholds on a 53-bit uniform
void gaussian_thresholds(uint64_t *thr) {
    double total = 0, acc = 0;
    for (int i = -64; i <= 64; i++) total += exp(-M_PI * (double)i * i / (6.4 * 6.4));
    for (int i = -64; i < 64; i++) {
        acc += exp(-M_PI * (double)i * i / (6.4 * 6.4)) / total;
        thr[i + 64] = (uint64_t)floor(acc * 9007199254740992.0) + 1;
    }
}
inline uint64_t mulmod_u64(uint64_t a, uint64_t b, uint64_t m) { return (uint64_t)((unsigned __int128)a * b % m); }
inline uint64_t inv_pow2_mod_Q(size_t e) {            // (2^e)^-1 mod Q, Q odd
    uint64_t half = (kQ + 1) / 2, r = 1;
    for (size_t i = 0; i < e; i++) r = mulmod_u64(r, half, kQ);
    return r;
}
// residues of the gadget constant 2^(bits_per * j) (0 when the shift leaves 64 bits, src/util.cpp:101)
inline void gadget_scalar(uint32_t bits_per, size_t j, uint32_t out[2]) {
    const uint64_t sh = (uint64_t)bits_per * j;
    const uint64_t v = sh >= 64 ? 0 : 1ull << sh;
    out[0] = (uint32_t)(v % kP); out[1] = (uint32_t)(v % kB);
}
int client_matrix_down(sb200_client *c, uint64_t *host, const uint32_t *dev, size_t npolys) {
    DBuf<uint64_t> tmp;
    const size_t chunk = 2048;
    CU(tmp.alloc(std::min(chunk, npolys) * PLW));
    for (size_t o = 0; o < npolys; o += chunk) {
        const size_t n = std::min(chunk, npolys - o);
        launch_ntt_dev_to_u64(tmp.p, dev + o * PLW, n, c->st); CHECK_LAUNCH();
        CU(cudaMemcpyAsync(host + o * PLW, tmp.p, n * PLW * 8, cudaMemcpyDeviceToHost, c->st));
        CU(cudaStreamSynchronize(c->st));
    }
    return SB200_OK;
}
// getPublicEncryptions (src/client.cpp:271-290): `count` rounds of t Regev encryptions of tau_i(s) * 2^(bits_per k)
int client_expansion_keys(sb200_client *c, uint64_t *W_host, size_t count, uint32_t t, uint32_t cls) {
    DBuf<uint32_t> W, tau_ntt, scal;
    DBuf<uint64_t> tau_raw;
    CU(W.alloc(count * 2 * t * PLW)); CU(tau_ntt.alloc(PLW)); CU(tau_raw.alloc(kN)); CU(scal.alloc((size_t)t * 2));
    std::vector<uint32_t> hs((size_t)t * 2);
    for (uint32_t k = 0; k < t; k++) gadget_scalar(get_bits_per(t), k, &hs[2 * k]);
    CU(cudaMemcpyAsync(scal.p, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice, c->st));
    for (size_t i = 0; i < count; i++) {
        launch_automorph(tau_raw.p, c->sr_raw.p, 1, (uint32_t)((kN >> i) + 1), c->st);
        launch_to_ntt(tau_ntt.p, tau_raw.p, 1, c->st);
        RegevArgs a{};
        a.key = c->key; a.obj_base = cc_obj(cls, i * t); a.sub_e = 1; a.wire = 0; a.R = 1; a.out_cols = (int)t; a.col_begin = 0;
        a.S = c->sr_ntt.p; a.msg = tau_ntt.p; a.scal = scal.p; a.ncols = (int)t;
        launch_client_regev_cols(W.p + i * 2 * t * PLW, a, c->st); CHECK_LAUNCH();
    }
    return client_matrix_down(c, W_host, W.p, count * 2 * t);
}
}  // namespace

extern "C" int sb200_client_gaussian_thresholds(uint64_t *out128) {
    if (!out128) return fail(SB200_ERR_ARG, "null argument");
    gaussian_thresholds(out128);
    return SB200_OK;
}
static int client_create_impl(sb200_client **out, const sb200_params *prm, int device, const uint8_t *seed32, bool pack) {
    if (!out || !prm || !seed32) return fail(SB200_ERR_ARG, "client_create: null argument");
    if (prm->t_gsw == 0 || prm->t_conv == 0 || prm->t_exp == 0 || prm->t_exp_right == 0) return fail(SB200_ERR_ARG, "client_create: zero gadget length");
    if (pack && (prm->out_n == 0 || prm->out_n > 16 || prm->nu1 < 1 || prm->nu1 > 11)) return fail(SB200_ERR_ARG, "pack_client_create: 1 <= out_n <= 16 and 1 <= nu1 <= 11 required");
    if (!pack && ((size_t)1 << prm->nu1) + (size_t)prm->t_gsw * prm->nu2 > (size_t)kN) return fail(SB200_ERR_ARG, "client_create: 2^nu1 + t_GSW*nu2 exceeds the 2048 query slots");
    if (sb200_arb_qprime(prm->qp_bits) == 0) return fail(SB200_ERR_ARG, "client_create: no response modulus for qp_bits = %u (14..36)", prm->qp_bits);
    if (prm->p_db == 0 || prm->p_db > 65536) return fail(SB200_ERR_ARG, "client_create: p_db must be in [1, 65536]");
    int rc = sb200_init(device);
    if (rc) return rc;
    sb200_client *c = new sb200_client();
    c->prm = *prm; c->device = device; c->pack = pack;
    const size_t nbits = (size_t)prm->t_gsw * prm->nu2, dim0 = (size_t)1 << prm->nu1;       // src/spiral.cpp:2076-2085
    c->g = ceil_log2(nbits + dim0);
    if (pack) {                                 // testHighRate, src/testing.cpp:795-798: always the stopround form
        c->sp_rows = prm->out_n;
        c->stopround = ceil_log2(nbits ? nbits : 1);
        c->n_right = c->stopround + 1;
    } else {
        c->stopround = nbits > dim0 ? 0 : ceil_log2(nbits);
        c->n_right = c->stopround > 0 ? c->stopround + 1 : c->g;
    }
    for (int i = 0; i < 8; i++) memcpy(&c->key.w[i], seed32 + 4 * i, 4);
    uint64_t thr[128];
    gaussian_thresholds(thr);
    cudaError_t e = cudaMemcpyToSymbol(c_gauss_thr, thr, sizeof(thr));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = c->sr_raw.alloc(kN);
    if (e == cudaSuccess) e = c->Sp_raw.alloc(c->sp_rows * (size_t)kN);
    if (e == cudaSuccess) e = c->sr_ntt.alloc(PLW);
    if (e == cudaSuccess) e = c->Sp_ntt.alloc(c->sp_rows * PLW);
    if (e != cudaSuccess) { delete c; return fail(SB200_ERR_CUDA, "client_create: %s", cudaGetErrorString(e)); }
    // keygen (src/client.cpp:23-47): s and S' from the error distribution
    launch_client_gauss_raw(c->sr_raw.p, c->key, cc_obj(CC_KEYS, 0), 0, 1, c->st);
    launch_client_gauss_raw(c->Sp_raw.p, c->key, cc_obj(CC_KEYS, 1), 0, (int)c->sp_rows, c->st);
    launch_to_ntt(c->sr_ntt.p, c->sr_raw.p, 1, c->st);
    launch_to_ntt(c->Sp_ntt.p, c->Sp_raw.p, c->sp_rows, c->st);
    e = cudaStreamSynchronize(c->st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { delete c; return fail(SB200_ERR_CUDA, "client_create: %s", cudaGetErrorString(e)); }
    *out = c;
    return SB200_OK;
}
extern "C" int sb200_client_create(sb200_client **out, const sb200_params *prm, int device, const uint8_t *seed32) {
    return client_create_impl(out, prm, device, seed32, false);
}
extern "C" int sb200_pack_client_create(sb200_client **out, const sb200_params *prm, int device, const uint8_t *seed32) {
    return client_create_impl(out, prm, device, seed32, true);
}
extern "C" void sb200_client_destroy(sb200_client *c) { delete c; }
extern "C" int sb200_client_secret(sb200_client *c, uint64_t *sr_raw_host, uint64_t *Sp_raw_host) {
    if (!c || !sr_raw_host || !Sp_raw_host) return fail(SB200_ERR_ARG, "client_secret: null argument");
    CU(c->sr_raw.down(sr_raw_host, kN));
    CU(c->Sp_raw.down(Sp_raw_host, c->sp_rows * (size_t)kN));
    return SB200_OK;
}
extern "C" int sb200_client_public_param_polys(const sb200_client *c, size_t *out4) {
    if (!c || !out4) return fail(SB200_ERR_ARG, "null argument");
    out4[0] = c->g * 2 * c->prm.t_exp; out4[1] = c->n_right * 2 * c->prm.t_exp_right;
    out4[2] = out4[3] = 3 * 2 * (size_t)c->prm.t_conv;
    return SB200_OK;
}
extern "C" int sb200_client_public_params(sb200_client *c, uint64_t *W_exp_left, uint64_t *W_exp_right, uint64_t *W_conv, uint64_t *V_conv) {
    if (!c || !W_exp_left || !W_exp_right || !W_conv || !V_conv) return fail(SB200_ERR_ARG, "client_public_params: null argument");
    if (c->pack) return fail(SB200_ERR_STATE, "client_public_params: this is a Pack client (sb200_pack_client_public_params)");
    CU(cudaSetDevice(c->device));
    TRY(client_expansion_keys(c, W_exp_right, c->n_right, c->prm.t_exp_right, CC_W_RIGHT));
    TRY(client_expansion_keys(c, W_exp_left, c->g, c->prm.t_exp, CC_W_LEFT));
    const size_t mc = c->prm.t_conv, m = 2 * mc;
    DBuf<uint32_t> M, msg, scal;
    CU(M.alloc(3 * m * PLW)); CU(msg.alloc(2 * PLW)); CU(scal.alloc(2 * m * 2));
    std::vector<uint32_t> hs(2 * m * 2, 0);
    {   // W = P + [0 ; s0 * G]  (src/spiral.cpp:2207-2217), G = gadget(n0, m): G[r][r + 2j] = 2^(bits_per j)
        const uint32_t bp = get_bits_per((uint32_t)(m / kN0));
        for (size_t r = 0; r < (size_t)kN0; r++)
            for (size_t k = 0; k < m; k++)
                if (k % kN0 == r) gadget_scalar(bp, k / kN0, &hs[(r * m + k) * 2]);
        CU(cudaMemcpyAsync(scal.p, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice, c->st));
        CU(cudaMemcpyAsync(msg.p, c->sr_ntt.p, PLW * 4, cudaMemcpyDeviceToDevice, c->st));
        CU(cudaMemcpyAsync(msg.p + PLW, c->sr_ntt.p, PLW * 4, cudaMemcpyDeviceToDevice, c->st));
        RegevArgs a{};
        a.key = c->key; a.obj_base = cc_obj(CC_W_CONV, 0); a.sub_e = 1; a.R = kN0; a.out_cols = (int)m; a.col_begin = 0;
        a.S = c->Sp_ntt.p; a.msg = msg.p; a.scal = scal.p; a.ncols = (int)m;
        launch_client_regev_cols(M.p, a, c->st); CHECK_LAUNCH();
        TRY(client_matrix_down(c, W_conv, M.p, 3 * m));
    }
    {   // V = P + [0 ; S' * [s0 * gv | gv]]  (src/spiral.cpp:2274-2290), gv = gadget(1, t_conv)
        const uint32_t bp = get_bits_per((uint32_t)mc);
        std::fill(hs.begin(), hs.end(), 0u);
        for (size_t r = 0; r < (size_t)kN0; r++)
            for (size_t k = 0; k < mc; k++) gadget_scalar(bp, k, &hs[(r * mc + k) * 2]);
        CU(cudaMemcpyAsync(scal.p, hs.data(), 2 * mc * 2 * 4, cudaMemcpyHostToDevice, c->st));
        launch_matmul(msg.p, c->Sp_ntt.p, c->sr_ntt.p, kN0, 1, 1, c->st);                      // S' * s0
        RegevArgs a{};
        a.key = c->key; a.sub_e = 1; a.R = kN0; a.out_cols = (int)m; a.S = c->Sp_ntt.p; a.scal = scal.p; a.ncols = (int)mc;
        a.obj_base = cc_obj(CC_V_CONV, 0); a.col_begin = 0; a.msg = msg.p;
        launch_client_regev_cols(M.p, a, c->st);                                               // columns [0, mc): S' s0 gv_k
        a.obj_base = cc_obj(CC_V_CONV, mc); a.col_begin = (int)mc; a.msg = c->Sp_ntt.p;
        launch_client_regev_cols(M.p, a, c->st); CHECK_LAUNCH();                               // columns [mc, 2mc): S' gv_k
        TRY(client_matrix_down(c, V_conv, M.p, 3 * m));
    }
    return SB200_OK;
}
// The wire seed fixes row 0 = -a of the query ciphertext.  Two queries of one client with the same `a` leak the difference of
// their messages (row1 - row1' = e - e' + sigma - sigma'), so the seed must be fresh per query: it is derived from the client key
// and the query counter - one ChaCha20 block, nonce {"SB2C", CC_QUERY | query_id, "wsee"} - and a monotonic query_id is the only
// thing a caller has to keep.
static const uint32_t kWireSeedStream = 0x65657377u;
extern "C" int sb200_client_wire_seed(const sb200_client *c, uint32_t query_id, uint8_t *seed32_out) {
    if (!c || !seed32_out) return fail(SB200_ERR_ARG, "client_wire_seed: null argument");
    if (query_id >= (1u << 24)) return fail(SB200_ERR_ARG, "client_wire_seed: query_id must be below 2^24");
    uint32_t x[16];
    chacha20_block(x, c->key.w, 0u, kClientMagic, cc_obj(CC_QUERY, query_id), kWireSeedStream);
    memcpy(seed32_out, x, 32);
    return SB200_OK;
}
static int client_wire_from_sigma(sb200_client *c, const std::vector<uint64_t> &sigma, uint32_t query_id, const uint8_t *wire_seed32, uint8_t *wire_out);
// query encoding (src/spiral.cpp:2098-2157) + encryptSimpleRegev, straight into the SEEDED wire form
extern "C" int sb200_client_query_wire(sb200_client *c, size_t idx_target, uint32_t query_id, const uint8_t *wire_seed32, uint8_t *wire_out) {
    if (!c || !wire_out) return fail(SB200_ERR_ARG, "client_query_wire: null argument");
    uint8_t derived[32];
    if (!wire_seed32) {                      // the safe default: row-0 seed from (client key, query_id)
        TRY(sb200_client_wire_seed(c, query_id, derived));
        wire_seed32 = derived;
    }
    const sb200_params &p = c->prm;
    const size_t fd = p.nu2, ell = p.t_gsw, dim0 = (size_t)1 << p.nu1;
    if (c->pack && dim0 + ell * fd > (size_t)kN) return fail(SB200_ERR_ARG, "pack_client_query_wire: 2^nu1 + t_GSW*nu2 exceeds the 2048 query slots (direct upload only)");
    if (idx_target >= (dim0 << fd)) return fail(SB200_ERR_ARG, "client_query_wire: index %zu outside the 2^%zu records", idx_target, (size_t)(p.nu1 + fd));
    if (query_id >= (1u << 24)) return fail(SB200_ERR_ARG, "client_query_wire: query_id must be below 2^24");
    CU(cudaSetDevice(c->device));
    const size_t idx_dim0 = idx_target >> fd, idx_further = idx_target & (((size_t)1 << fd) - 1);
    const uint32_t bits_per = get_bits_per((uint32_t)ell);
    std::vector<uint64_t> sigma(kN, 0);
    const uint64_t scale_k = kQ / p.p_db;
    if (c->pack || c->stopround != 0) {       // Pack variants always use the two-scale encoding (src/testing.cpp:987-1004)
        const uint64_t inv_first = inv_pow2_mod_Q(c->g), inv_rest = inv_pow2_mod_Q(c->stopround + 1);
        sigma[2 * idx_dim0] = mulmod_u64(scale_k % kQ, inv_first, kQ);
        for (size_t i = 0; i < fd; i++)
            for (size_t j = 0; j < ell; j++)
                if ((idx_further >> i) & 1) sigma[2 * (i * ell + j) + 1] = mulmod_u64((1ull << (bits_per * j)) % kQ, inv_rest, kQ);
    } else {
        const uint64_t inv = inv_pow2_mod_Q(c->g);
        sigma[idx_dim0] = mulmod_u64(scale_k % kQ, inv, kQ);
        size_t ctr = 0;
        for (size_t i = 0; i < fd; i++)
            for (size_t j = 0; j < ell; j++, ctr++)
                if ((idx_further >> i) & 1) sigma[dim0 + ctr] = mulmod_u64((1ull << (bits_per * j)) % kQ, inv, kQ);
    }
    return client_wire_from_sigma(c, sigma, query_id, wire_seed32, wire_out);
}
static int client_wire_from_sigma(sb200_client *c, const std::vector<uint64_t> &sigma, uint32_t query_id, const uint8_t *wire_seed32, uint8_t *wire_out) {
    DBuf<uint64_t> sig_raw, row1_raw, packed;
    DBuf<uint32_t> sig_ntt, ct;
    CU(sig_raw.alloc(kN)); CU(row1_raw.alloc(kN)); CU(packed.alloc(kWireRowBytes / 8)); CU(sig_ntt.alloc(PLW)); CU(ct.alloc(2 * PLW));
    CU(cudaMemcpyAsync(sig_raw.p, sigma.data(), kN * 8, cudaMemcpyHostToDevice, c->st));
    launch_to_ntt(sig_ntt.p, sig_raw.p, 1, c->st);
    RegevArgs a{};
    a.key = c->key; a.wire = 1; a.obj_base = cc_obj(CC_QUERY, query_id); a.sub_e = 1; a.R = 1; a.out_cols = 1; a.col_begin = 0;
    for (int i = 0; i < 8; i++) memcpy(&a.ukey.w[i], wire_seed32 + 4 * i, 4);
    a.S = c->sr_ntt.p; a.msg = sig_ntt.p; a.scal = nullptr; a.ncols = 1;
    launch_client_regev_cols(ct.p, a, c->st);
    launch_from_ntt(row1_raw.p, ct.p + PLW, 1, c->st);
    launch_bitpack(packed.p, row1_raw.p, kN, 56, c->st); CHECK_LAUNCH();
    const uint32_t magic = kWireQueryMagic;
    memcpy(wire_out, &magic, 4);
    wire_out[4] = (uint8_t)kWireSeeded; wire_out[5] = wire_out[6] = wire_out[7] = 0;
    memcpy(wire_out + kWireHeaderBytes, wire_seed32, kWireSeedBytes);
    CU(cudaMemcpyAsync(wire_out + kWireHeaderBytes + kWireSeedBytes, packed.p, kWireRowBytes, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return SB200_OK;
}
// decoding (check_final, src/spiral.cpp:1428-1476): total_resp 3x2 raw (row 0 mod q', rows 1-2 mod 4p) -> 2x2 plaintext polynomials
extern "C" int sb200_client_decode(sb200_client *c, const uint64_t *total_resp_host, uint64_t *pt_out_host) {
    if (!c || !total_resp_host || !pt_out_host) return fail(SB200_ERR_ARG, "client_decode: null argument");
    CU(cudaSetDevice(c->device));
    // Spiral: 3x2 response -> 2x2 plaintext polynomials; Pack variants: (out_n+1) x out_n -> out_n x out_n (item of plane i*out_n + j)
    const size_t nn = c->pack ? c->prm.out_n : (size_t)kN2, rw = (nn + 1) * nn * (size_t)kN, pw = nn * nn * (size_t)kN;
    DBuf<uint64_t> resp, pt;
    CU(resp.alloc(rw)); CU(pt.alloc(pw));
    CU(cudaMemcpyAsync(resp.p, total_resp_host, rw * 8, cudaMemcpyHostToDevice, c->st));
    launch_client_decode(pt.p, resp.p, c->Sp_raw.p, sb200_arb_qprime(c->prm.qp_bits), c->prm.p_db, (int)nn, c->st); CHECK_LAUNCH();
    CU(cudaMemcpyAsync(pt_out_host, pt.p, pw * 8, cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    return SB200_OK;
}


// ---- SpiralPack / SpiralStreamPack client (testHighRate's client statements, src/testing.cpp:904-1005, 1086-1122) ---------------
// Same counter-based randomness as the Spiral client; oracle/client_sim.c (so_pack_client_new_chacha) states it in plain C.
//   keys        S' has out_n rows (keygen(S, Sp, sr, out_n), src/client.cpp:23-47)
//   v_W[i]      (out_n+1) x t_conv: fresh public key under S' + [0 ; s0 * g_vec in row 1+i]   (:905-912), objects (7, i*t_conv + k)
//   V           2 x 2*t_conv: column i encrypts s0^2 * G[0][i] (i even) or s0 * G[1][i] (i odd) under s0   (:918-931), objects (5, i)
//   queries     the packed single ciphertext (:987-1005, wire form as the Spiral client's) or the direct upload (:962-985):
//               2^nu1 Regev ciphertexts + nu2 GSW ciphertexts (2 x 2*ell), objects (8, ciphertext number)
//   decode      out_n x out_n products S'_r * resp_row0[c] over q' and the reference's rounding (:1086-1118)
extern "C" int sb200_pack_client_public_param_polys(const sb200_client *c, size_t *out4) {
    if (!c || !out4 || !c->pack) return fail(SB200_ERR_ARG, "pack_client_public_param_polys: needs a Pack client");
    const size_t n = c->prm.out_n;
    out4[0] = c->g * 2 * c->prm.t_exp; out4[1] = c->n_right * 2 * c->prm.t_exp_right;
    out4[2] = 2 * 2 * (size_t)c->prm.t_conv; out4[3] = n * (n + 1) * c->prm.t_conv;
    return SB200_OK;
}
extern "C" int sb200_pack_client_public_params(sb200_client *c, uint64_t *W_exp_left, uint64_t *W_exp_right, uint64_t *V, uint64_t *v_W) {
    if (!c || !v_W || !c->pack) return fail(SB200_ERR_ARG, "pack_client_public_params: needs a Pack client and v_W");
    CU(cudaSetDevice(c->device));
    const size_t n = c->prm.out_n, mc = c->prm.t_conv;
    const uint32_t bp = get_bits_per((uint32_t)mc);
    {   // packing keys
        DBuf<uint32_t> M, msg, scal;
        CU(M.alloc((n + 1) * mc * PLW)); CU(msg.alloc(n * PLW)); CU(scal.alloc(n * mc * 2));
        for (size_t r = 0; r < n; r++) CU(cudaMemcpyAsync(msg.p + r * PLW, c->sr_ntt.p, PLW * 4, cudaMemcpyDeviceToDevice, c->st));
        for (size_t i = 0; i < n; i++) {
            std::vector<uint32_t> hs(n * mc * 2, 0);
            for (size_t k = 0; k < mc; k++) gadget_scalar(bp, k, &hs[(i * mc + k) * 2]);             // the message sits in row 1 + i only
            CU(cudaMemcpyAsync(scal.p, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice, c->st));
            RegevArgs a{};
            a.key = c->key; a.obj_base = cc_obj(CC_PACK_W, i * mc); a.sub_e = 1; a.R = (int)n; a.out_cols = (int)mc; a.col_begin = 0;
            a.S = c->Sp_ntt.p; a.msg = msg.p; a.scal = scal.p; a.ncols = (int)mc;
            launch_client_regev_cols(M.p, a, c->st); CHECK_LAUNCH();
            TRY(client_matrix_down(c, v_W + i * (n + 1) * mc * PLW, M.p, (n + 1) * mc));           // synchronises: hs may go out of scope
        }
    }
    if (!W_exp_left && !W_exp_right && !V) return SB200_OK;                                        // direct-upload client
    if (!W_exp_left || !W_exp_right || !V) return fail(SB200_ERR_ARG, "pack_client_public_params: expansion keys and V come together");
    TRY(client_expansion_keys(c, W_exp_left, c->g, c->prm.t_exp, CC_W_LEFT));
    TRY(client_expansion_keys(c, W_exp_right, c->n_right, c->prm.t_exp_right, CC_W_RIGHT));
    {   // V: even columns carry s0^2 * 2^(bp * i/2), odd columns s0 * 2^(bp * i/2)  (G = gadget(2, 2*t_conv))
        const size_t m = 2 * mc;
        DBuf<uint32_t> M, s0sq, scal;
        CU(M.alloc(2 * m * PLW)); CU(s0sq.alloc(PLW)); CU(scal.alloc(mc * 2));
        std::vector<uint32_t> hs(mc * 2);
        for (size_t k = 0; k < mc; k++) gadget_scalar(bp, k, &hs[k * 2]);
        CU(cudaMemcpyAsync(scal.p, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice, c->st));
        launch_matmul(s0sq.p, c->sr_ntt.p, c->sr_ntt.p, 1, 1, 1, c->st);
        RegevArgs a{};
        a.key = c->key; a.sub_e = 1; a.R = 1; a.out_cols = (int)m; a.S = c->sr_ntt.p; a.scal = scal.p; a.ncols = (int)mc; a.col_stride = 2;
        a.obj_base = cc_obj(CC_V_CONV, 0); a.col_begin = 0; a.msg = s0sq.p;
        launch_client_regev_cols(M.p, a, c->st);
        a.obj_base = cc_obj(CC_V_CONV, 1); a.col_begin = 1; a.msg = c->sr_ntt.p;
        launch_client_regev_cols(M.p, a, c->st); CHECK_LAUNCH();
        TRY(client_matrix_down(c, V, M.p, 2 * m));
    }
    return SB200_OK;
}
extern "C" int sb200_pack_client_query_wire(sb200_client *c, size_t idx_target, uint32_t query_id, const uint8_t *wire_seed32, uint8_t *wire_out) {
    if (!c || !c->pack) return fail(SB200_ERR_ARG, "pack_client_query_wire: needs a Pack client");
    return sb200_client_query_wire(c, idx_target, query_id, wire_seed32, wire_out);
}
// direct upload: v_firstdim = 2^nu1 ciphertexts (2x1 ref-NTT each), v_folding = nu2 x (2 x 2*ell) ref-NTT - the arguments of
// sb200_pack_server_answer_direct / _upload_direct.  query_id (< 2^8) selects a fresh block of 2^16 objects per query.
extern "C" int sb200_pack_client_query_direct(sb200_client *c, size_t idx_target, uint32_t query_id, uint64_t *v_firstdim, uint64_t *v_folding) {
    if (!c || !c->pack || !v_firstdim) return fail(SB200_ERR_ARG, "pack_client_query_direct: needs a Pack client and output buffers");
    const sb200_params &p = c->prm;
    const size_t fd = p.nu2, ell = p.t_gsw, dim0 = (size_t)1 << p.nu1;
    if (fd && !v_folding) return fail(SB200_ERR_ARG, "pack_client_query_direct: v_folding missing");
    if (idx_target >= (dim0 << fd)) return fail(SB200_ERR_ARG, "pack_client_query_direct: index %zu outside the 2^%zu records", idx_target, (size_t)(p.nu1 + fd));
    if (query_id >= 256 || dim0 + 2 * ell * fd > 65536) return fail(SB200_ERR_ARG, "pack_client_query_direct: query_id must be below 2^8");
    CU(cudaSetDevice(c->device));
    const size_t idx_dim0 = idx_target >> fd, idx_further = idx_target & (((size_t)1 << fd) - 1);
    const uint32_t bits_per = get_bits_per((uint32_t)ell), qbase = query_id << 16;
    const uint64_t scale_k = kQ / p.p_db;
    DBuf<uint32_t> ones, cts, scal;
    DBuf<uint64_t> one_raw;
    CU(ones.alloc(PLW)); CU(one_raw.alloc(kN)); CU(cts.alloc(std::max(dim0 * 2, 2 * 2 * ell) * PLW)); CU(scal.alloc(std::max(dim0, ell) * 2));
    CU(cudaMemsetAsync(one_raw.p, 0, kN * 8, c->st));
    const uint64_t one = 1;
    CU(cudaMemcpyAsync(one_raw.p, &one, 8, cudaMemcpyHostToDevice, c->st));
    launch_to_ntt(ones.p, one_raw.p, 1, c->st);                                  // NTT of the constant 1
    {   // first dimension: Regev encryptions of scale_k * [i == idx_dim0]
        std::vector<uint32_t> hs(dim0 * 2, 0);
        hs[idx_dim0 * 2] = (uint32_t)(scale_k % kP); hs[idx_dim0 * 2 + 1] = (uint32_t)(scale_k % kB);
        CU(cudaMemcpyAsync(scal.p, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice, c->st));
        RegevArgs a{};
        a.key = c->key; a.obj_base = cc_obj(CC_QUERY_DIRECT, qbase); a.sub_e = 1; a.R = 1; a.out_cols = (int)dim0; a.col_begin = 0; a.ct_major = 1;
        a.S = c->sr_ntt.p; a.msg = ones.p; a.scal = scal.p; a.ncols = (int)dim0;
        launch_client_regev_cols(cts.p, a, c->st); CHECK_LAUNCH();
        TRY(client_matrix_down(c, v_firstdim, cts.p, dim0 * 2));
    }
    for (size_t i = 0; i < fd; i++) {   // GSW ciphertext of bit i: column 2j+1 encrypts val_j = bit * 2^(bits_per j), column 2j encrypts s0 * val_j
        const uint64_t bit = (idx_further >> i) & 1;
        std::vector<uint32_t> hs(ell * 2, 0);
        for (size_t j = 0; j < ell; j++) { const uint64_t val = ((uint64_t)1 << (bits_per * j)) * bit; hs[2 * j] = (uint32_t)(val % kP); hs[2 * j + 1] = (uint32_t)(val % kB); }
        CU(cudaMemcpyAsync(scal.p, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice, c->st));
        RegevArgs a{};
        a.key = c->key; a.sub_e = 1; a.R = 1; a.out_cols = (int)(2 * ell); a.S = c->sr_ntt.p; a.scal = scal.p; a.ncols = (int)ell; a.col_stride = 2;
        a.obj_base = cc_obj(CC_QUERY_DIRECT, qbase + dim0 + i * ell * 2); a.col_begin = 0; a.msg = c->sr_ntt.p;
        launch_client_regev_cols(cts.p, a, c->st);
        a.obj_base = cc_obj(CC_QUERY_DIRECT, qbase + dim0 + i * ell * 2 + 1); a.col_begin = 1; a.msg = ones.p;
        launch_client_regev_cols(cts.p, a, c->st); CHECK_LAUNCH();
        TRY(client_matrix_down(c, v_folding + i * 2 * 2 * ell * PLW, cts.p, 2 * 2 * ell));
    }
    return SB200_OK;
}
extern "C" int sb200_pack_client_decode(sb200_client *c, const uint64_t *total_resp_host, uint64_t *pt_out_host) {
    if (!c || !c->pack) return fail(SB200_ERR_ARG, "pack_client_decode: needs a Pack client");
    return sb200_client_decode(c, total_resp_host, pt_out_host);
}
