/*
 * client_sim.c - a minimal Spiral CLIENT (key generation, public parameters, query, decoding).
 *
 * TEST INFRASTRUCTURE ONLY (part of liboracle.so).  The client is out of scope for the product
 * (SURVEY section 2); this restatement exists so the tests can drive the CUDA server with REAL
 * encryptions and check the one end-to-end property the reference itself checks: the decoded
 * record equals the planted record ("Is correct?: 1", src/spiral.cpp:1494).  Randomness comes
 * from a seeded splitmix64 instead of the reference's unseeded std::random_device.
 *
 * Restates: keygen src/client.cpp:23-47, getRegevSample :141-157, encryptSimpleRegev :170-186,
 * encryptSimpleRegevMatrix :209-227, get_fresh_public_key_raw :49-68, getPublicEncryptions
 * :271-290, W / V generation src/spiral.cpp:2207-2290, query encoding :2098-2157, decoding
 * (check_final) :1428-1476, discrete Gaussian src/core.cpp:182-207.
 */
#include "client_sim.h"
#include "wire_format.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define N SO_N
#define PL (2 * (size_t)SO_N)
typedef unsigned __int128 u128;

struct so_client {
    so_params prm;
    so_rng rng;
    int nonoise;
    uint64_t *Sp;       /* n0 x 1 raw */
    uint64_t *sr;       /* 1 x 1 raw  */
    double cdf[2 * 64 + 1];
    /* counter-based randomness (so_client_new_chacha): every random polynomial is a ChaCha20 stream named by an object id,
     * so a parallel implementation (the CUDA client, spiral_b200/csrc/client_kernels.cu) can be compared bit for bit */
    int chacha;
    uint32_t key[8];
    uint64_t thr[128];  /* Gaussian thresholds: floor(cdf[k] * 2^53) + 1 */
};
#define CC_MAGIC 0x43324253u            /* "SB2C" */
#define CC_OBJ(cls, idx) (((uint32_t)(cls) << 24) | (uint32_t)(idx))
enum { CC_KEYS = 1, CC_W_RIGHT = 2, CC_W_LEFT = 3, CC_W_CONV = 4, CC_V_CONV = 5, CC_QUERY = 6, CC_PACK_W = 7, CC_QUERY_DIRECT = 8 };

static uint64_t *xalloc(size_t words) {
    uint64_t *p = (uint64_t *)calloc(words ? words : 1, sizeof(uint64_t));
    if (!p) abort();
    return p;
}

/* discrete Gaussian, width 6.4, support [-64, 64]  (src/core.cpp:182-207) */
static void build_cdf(so_client *c) {
    double total = 0, acc = 0;
    for (int i = -64; i <= 64; i++) total += exp(-M_PI * (double)i * i / (6.4 * 6.4));
    for (int i = -64; i <= 64; i++) { acc += exp(-M_PI * (double)i * i / (6.4 * 6.4)) / total; c->cdf[i + 64] = acc; }
}
static uint64_t sample_u64(so_client *c) {                 /* src/client.cpp:7-10 */
    if (c->nonoise) return 0;
    double u = (double)(so_rng_next(&c->rng) >> 11) / 9007199254740992.0;
    int k = 0;
    while (k < 128 && c->cdf[k] < u) k++;
    int64_t v = k - 64;
    return (uint64_t)((v + (int64_t)SO_Q) % (int64_t)SO_Q);
}
/* uniform polynomial in NTT form ([2][N]): slot (n, z) from block(counter = n*N + z, nonce = {"SB2C", obj, sub}), the first of
 * its 16 words whose low 28 bits are below the prime (as so_wire_seeded_row0) */
static void cc_uniform_ntt(const so_client *c, uint32_t obj, uint32_t sub, uint64_t *out) {
    uint32_t nonce[3] = {CC_MAGIC, obj, sub}, blk[16];
    for (uint32_t n = 0; n < 2; n++) {
        const uint32_t q = (uint32_t)(n == 0 ? SO_P : SO_B);
        for (uint32_t z = 0; z < N; z++) {
            so_chacha20_block(c->key, n * N + z, nonce, blk);
            uint32_t v = (blk[15] & 0x0FFFFFFFu) % q;
            for (int k = 0; k < 16; k++) { uint32_t w = blk[k] & 0x0FFFFFFFu; if (w < q) { v = w; break; } }
            out[n * N + z] = v;
        }
    }
}
/* discrete Gaussian polynomial in raw form: coefficient z from block(counter = z, nonce = {"SB2C", obj, sub}): U = the top 53 bits
 * of words (1,0); value = -64 + #{k < 128 : U >= thr[k]} - the inverse-CDF walk of sample_u64 on u = U / 2^53, in integers */
static void cc_gauss_raw(const so_client *c, uint32_t obj, uint32_t sub, uint64_t *out) {
    uint32_t nonce[3] = {CC_MAGIC, obj, sub}, blk[16];
    for (uint32_t z = 0; z < N; z++) {
        so_chacha20_block(c->key, z, nonce, blk);
        uint64_t U = (((uint64_t)blk[1] << 32) | blk[0]) >> 11;
        int64_t v = -64;
        for (int k = 0; k < 128; k++) v += (U >= c->thr[k]);
        out[z] = v < 0 ? SO_Q - (uint64_t)(-v) : (uint64_t)v;
    }
}
static void noise(so_client *c, uint64_t *E, size_t npolys) { for (size_t i = 0; i < npolys * N; i++) E[i] = sample_u64(c); }

so_client *so_client_new(const so_params *prm, uint64_t seed, int nonoise) {
    so_client *c = (so_client *)calloc(1, sizeof(so_client));
    c->prm = *prm; c->rng.s = seed; c->nonoise = nonoise;
    build_cdf(c);
    c->Sp = xalloc(SO_N0 * N); c->sr = xalloc(N);
    /* keygen draws the key from the error distribution even under --nonoise; keep it non-trivial */
    int saved = c->nonoise; c->nonoise = 0;
    for (size_t m = 0; m < N; m++) c->sr[m] = sample_u64(c) % SO_Q;
    for (size_t r = 0; r < SO_N0; r++) for (size_t m = 0; m < N; m++) c->Sp[r * N + m] = sample_u64(c) % SO_Q;
    c->nonoise = saved;
    return c;
}
so_client *so_client_new_chacha(const so_params *prm, const uint8_t seed[32]) {
    so_client *c = (so_client *)calloc(1, sizeof(so_client));
    c->prm = *prm; c->chacha = 1;
    for (int i = 0; i < 8; i++) c->key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
    build_cdf(c);
    for (int k = 0; k < 128; k++) c->thr[k] = (uint64_t)floor(c->cdf[k] * 9007199254740992.0) + 1;
    c->Sp = xalloc(SO_N0 * N); c->sr = xalloc(N);
    cc_gauss_raw(c, CC_OBJ(CC_KEYS, 0), 0, c->sr);
    for (size_t r = 0; r < SO_N0; r++) cc_gauss_raw(c, CC_OBJ(CC_KEYS, 1 + r), 0, &c->Sp[r * N]);
    return c;
}
void so_client_gaussian_thresholds(const so_client *c, uint64_t *out128) { memcpy(out128, c->thr, sizeof(c->thr)); }
void so_client_secret(const so_client *c, uint64_t *sr, uint64_t *Sp) {
    memcpy(sr, c->sr, N * sizeof(uint64_t)); memcpy(Sp, c->Sp, SO_N0 * N * sizeof(uint64_t));
}
/* the same for a Pack client: S' has sp_rows = out_n rows */
void so_client_secret_n(const so_client *c, uint64_t *sr, uint64_t *Sp, size_t sp_rows) {
    memcpy(sr, c->sr, N * sizeof(uint64_t)); memcpy(Sp, c->Sp, sp_rows * N * sizeof(uint64_t));
}
void so_client_free(so_client *c) { if (c) { free(c->Sp); free(c->sr); free(c); } }

/* P (2x1 NTT) = [to_ntt(Q - a) ; a*s + e]   (getRegevSample) */
/* (-row 0) * s + e, row 0 given in NTT form, e in raw form */
static void regev_row1_from_row0(so_client *c, uint64_t *P_ntt, const uint64_t *e_raw) {
    uint64_t *a_ntt = xalloc(PL), *s_ntt = xalloc(PL), *e_ntt = xalloc(PL);
    for (size_t z = 0; z < N; z++) {
        a_ntt[z] = (SO_P - P_ntt[z]) % SO_P;
        a_ntt[N + z] = (SO_B - P_ntt[N + z]) % SO_B;
    }
    so_to_ntt(s_ntt, c->sr, 1); so_to_ntt(e_ntt, e_raw, 1);
    so_multiply(P_ntt + PL, a_ntt, s_ntt, 1, 1, 1);
    so_add(P_ntt + PL, P_ntt + PL, e_ntt, 1);
    free(a_ntt); free(s_ntt); free(e_ntt);
}
static void regev_sample(so_client *c, uint64_t *P_ntt, uint32_t obj) {
    if (c->chacha) {                                   /* row 0 = -a drawn directly in NTT form */
        uint64_t *e = xalloc(N);
        cc_uniform_ntt(c, obj, 0, P_ntt);
        cc_gauss_raw(c, obj, 1, e);
        regev_row1_from_row0(c, P_ntt, e);
        free(e);
        return;
    }
    uint64_t *a = xalloc(N), *e = xalloc(N), *a_inv = xalloc(N);
    uint64_t *a_ntt = xalloc(PL), *s_ntt = xalloc(PL), *e_ntt = xalloc(PL), *b = xalloc(PL);
    so_fill_uniform_raw(a, N, &c->rng);
    noise(c, e, 1);
    so_invert(a_inv, a, 1);
    so_to_ntt(a_ntt, a, 1); so_to_ntt(s_ntt, c->sr, 1); so_to_ntt(e_ntt, e, 1);
    so_multiply(b, a_ntt, s_ntt, 1, 1, 1);
    so_add(b, b, e_ntt, 1);
    so_to_ntt(P_ntt, a_inv, 1);
    memcpy(P_ntt + PL, b, PL * sizeof(uint64_t));
    free(a); free(e); free(a_inv); free(a_ntt); free(s_ntt); free(e_ntt); free(b);
}
/* encryptSimpleRegev(sigma): 2x1 NTT */
static void encrypt_simple_regev(so_client *c, uint64_t *out_ntt, const uint64_t *sigma_raw, uint32_t obj) {
    uint64_t *sig_ntt = xalloc(PL);
    regev_sample(c, out_ntt, obj);
    so_to_ntt(sig_ntt, sigma_raw, 1);
    so_add(out_ntt + PL, out_ntt + PL, sig_ntt, 1);
    free(sig_ntt);
}
/* encryptSimpleRegevMatrix(s, mat_to_enc (1 x m NTT)) -> 2 x m NTT */
static void encrypt_simple_regev_matrix(so_client *c, uint64_t *out, const uint64_t *mat_ntt, size_t m, uint32_t obj_base) {
    uint64_t P[2 * 2 * SO_N];
    for (size_t i = 0; i < m; i++) {
        regev_sample(c, P, obj_base + (uint32_t)i);
        memcpy(&out[(0 * m + i) * PL], P, PL * sizeof(uint64_t));
        so_add(&out[(1 * m + i) * PL], P + PL, &mat_ntt[i * PL], 1);
    }
}
/* get_fresh_public_key_raw(Sp, m) -> to_ntt(P), P = [Q - A ; Sp*A + E]  (n1 x m) */
static void fresh_public_key_ntt(so_client *c, uint64_t *P_ntt, size_t m, uint32_t cls) {
    if (c->chacha) {       /* column k: row 0 = -A_k uniform in NTT form (object (cls, k), stream 0), E[r][k] Gaussian (stream 1 + r) */
        uint64_t *a_ntt = xalloc(PL), *e = xalloc(N), *e_ntt = xalloc(PL), *Sp_ntt = xalloc(SO_N0 * PL), *prod = xalloc(PL);
        so_to_ntt(Sp_ntt, c->Sp, SO_N0);
        for (size_t k = 0; k < m; k++) {
            uint64_t *row0 = &P_ntt[k * PL];
            cc_uniform_ntt(c, CC_OBJ(cls, k), 0, row0);
            for (size_t z = 0; z < N; z++) { a_ntt[z] = (SO_P - row0[z]) % SO_P; a_ntt[N + z] = (SO_B - row0[N + z]) % SO_B; }
            for (size_t r = 0; r < SO_N0; r++) {
                cc_gauss_raw(c, CC_OBJ(cls, k), 1 + (uint32_t)r, e);
                so_to_ntt(e_ntt, e, 1);
                so_multiply(prod, &Sp_ntt[r * PL], a_ntt, 1, 1, 1);
                so_add(&P_ntt[((1 + r) * m + k) * PL], prod, e_ntt, 1);
            }
        }
        free(a_ntt); free(e); free(e_ntt); free(Sp_ntt); free(prod);
        return;
    }
    uint64_t *A = xalloc(m * N), *E = xalloc(SO_N0 * m * N), *A_ntt = xalloc(m * PL), *E_ntt = xalloc(SO_N0 * m * PL);
    uint64_t *Sp_ntt = xalloc(SO_N0 * PL), *Bp = xalloc(SO_N0 * m * PL), *P_raw = xalloc(SO_N1 * m * N);
    so_fill_uniform_raw(A, m * N, &c->rng);
    noise(c, E, SO_N0 * m);
    so_to_ntt(A_ntt, A, m); so_to_ntt(E_ntt, E, SO_N0 * m); so_to_ntt(Sp_ntt, c->Sp, SO_N0);
    so_multiply(Bp, Sp_ntt, A_ntt, SO_N0, 1, m);
    so_add(Bp, E_ntt, Bp, SO_N0 * m);
    uint64_t *A_back = xalloc(m * N);
    so_from_ntt(A_back, A_ntt, m);
    so_invert(P_raw, A_back, m);
    so_from_ntt(P_raw + m * N, Bp, SO_N0 * m);
    so_to_ntt(P_ntt, P_raw, SO_N1 * m);
    free(A); free(E); free(A_ntt); free(E_ntt); free(Sp_ntt); free(Bp); free(P_raw); free(A_back);
}

size_t so_client_w_exp_right_count(const so_params *prm) {
    size_t g, stop; so_spiral_expansion_shape(prm, &g, &stop);
    return stop > 0 ? stop + 1 : g;
}

static void expansion_keys(so_client *c, uint64_t *W, size_t count, uint32_t t, uint32_t cls) {   /* getPublicEncryptions */
    uint64_t *G = xalloc((size_t)t * N), *G_ntt = xalloc((size_t)t * PL), *tau = xalloc(N), *tau_ntt = xalloc(PL), *prod = xalloc((size_t)t * PL);
    so_build_gadget(G, 1, t);
    so_to_ntt(G_ntt, G, t);
    for (size_t i = 0; i < count; i++) {
        so_automorph(tau, c->sr, 1, (N >> i) + 1);
        so_to_ntt(tau_ntt, tau, 1);
        so_multiply(prod, tau_ntt, G_ntt, 1, 1, t);
        encrypt_simple_regev_matrix(c, &W[i * 2 * t * PL], prod, t, CC_OBJ(cls, i * t));
    }
    free(G); free(G_ntt); free(tau); free(tau_ntt); free(prod);
}

void so_client_spiral_pub_params(so_client *c, uint64_t *W_exp_left, uint64_t *W_exp_right, uint64_t *W_conv, uint64_t *V_conv) {
    const so_params *p = &c->prm;
    size_t g, stop; so_spiral_expansion_shape(p, &g, &stop);
    expansion_keys(c, W_exp_right, so_client_w_exp_right_count(p), p->t_exp_right, CC_W_RIGHT);   /* order as :2093-2094 */
    expansion_keys(c, W_exp_left, g, p->t_exp, CC_W_LEFT);

    size_t mc = p->t_conv, m = 2 * mc;
    uint64_t *s0_ntt = xalloc(PL), *Sp_ntt = xalloc(SO_N0 * PL);
    so_to_ntt(s0_ntt, c->sr, 1); so_to_ntt(Sp_ntt, c->Sp, SO_N0);
    {   /* W = P + [0 ; s0 * G_scale]   (:2207-2217) */
        uint64_t *G = xalloc(SO_N0 * m * N), *G_ntt = xalloc(SO_N0 * m * PL), *s0G = xalloc(SO_N0 * m * PL), *P = xalloc(SO_N1 * m * PL);
        so_build_gadget(G, SO_N0, m);
        so_to_ntt(G_ntt, G, SO_N0 * m);
        so_mul_by_const(s0G, s0_ntt, G_ntt, SO_N0 * m);
        fresh_public_key_ntt(c, P, m, CC_W_CONV);
        memcpy(W_conv, P, m * PL * sizeof(uint64_t));
        so_add(W_conv + m * PL, P + m * PL, s0G, SO_N0 * m);
        free(G); free(G_ntt); free(s0G); free(P);
    }
    {   /* V = P + [0 ; Sp * [s0*gv | gv]]   (:2274-2290) */
        uint64_t *gv = xalloc(mc * N), *gv_ntt = xalloc(mc * PL), *tog = xalloc(m * PL), *res = xalloc(SO_N0 * m * PL), *P = xalloc(SO_N1 * m * PL);
        so_build_gadget(gv, 1, mc);
        so_to_ntt(gv_ntt, gv, mc);
        so_mul_by_const(tog, s0_ntt, gv_ntt, mc);
        memcpy(tog + mc * PL, gv_ntt, mc * PL * sizeof(uint64_t));
        so_multiply(res, Sp_ntt, tog, SO_N0, 1, m);
        fresh_public_key_ntt(c, P, m, CC_V_CONV);
        memcpy(V_conv, P, m * PL * sizeof(uint64_t));
        so_add(V_conv + m * PL, P + m * PL, res, SO_N0 * m);
        free(gv); free(gv_ntt); free(tog); free(res); free(P);
    }
    free(s0_ntt); free(Sp_ntt);
}

static int64_t inv_mod(int64_t a, int64_t b) {             /* src/util.cpp:272-286 */
    int64_t b0 = b, t, q, x0 = 0, x1 = 1;
    if (b == 1) return 1;
    while (a > 1) { q = a / b; t = b; b = a % b; a = t; t = x0; x0 = x1 - q * x0; x1 = t; }
    if (x1 < 0) x1 += b0;
    return x1;
}

static uint64_t *query_sigma(so_client *c, size_t idx_target) {   /* :2098-2157 */
    const so_params *p = &c->prm;
    size_t g, stop; so_spiral_expansion_shape(p, &g, &stop);
    size_t fd = p->nu2, ell = p->t_gsw, dim0 = (size_t)1 << p->nu1;
    size_t idx_dim0 = idx_target >> fd, idx_further = idx_target & (((size_t)1 << fd) - 1);
    uint32_t bits_per = so_get_bits_per((uint32_t)ell);
    uint64_t scale_k = SO_Q / p->p_db;
    uint64_t *sigma = xalloc(N);
    if (stop != 0) {
        sigma[2 * idx_dim0] = scale_k % SO_Q;
        for (size_t i = 0; i < fd; i++) {
            uint64_t bit = (idx_further >> i) & 1;
            for (size_t j = 0; j < ell; j++) sigma[2 * (i * ell + j) + 1] = ((uint64_t)1 << (bits_per * j)) * bit % SO_Q;
        }
        uint64_t inv_first = (uint64_t)inv_mod((int64_t)1 << g, (int64_t)SO_Q), inv_rest = (uint64_t)inv_mod((int64_t)1 << (stop + 1), (int64_t)SO_Q);
        for (size_t i = 0; i < N / 2; i++) {
            sigma[2 * i] = (uint64_t)((u128)sigma[2 * i] * inv_first % SO_Q);
            sigma[2 * i + 1] = (uint64_t)((u128)sigma[2 * i + 1] * inv_rest % SO_Q);
        }
    } else {
        sigma[idx_dim0] = scale_k % SO_Q;
        size_t ctr = 0;
        for (size_t i = 0; i < fd; i++) {
            uint64_t bit = (idx_further >> i) & 1;
            for (size_t j = 0; j < ell; j++) sigma[dim0 + ctr++] = ((uint64_t)1 << (bits_per * j)) * bit % SO_Q;
        }
        uint64_t inv = (uint64_t)inv_mod((int64_t)1 << g, (int64_t)SO_Q);
        for (size_t i = 0; i < N; i++) sigma[i] = (uint64_t)((u128)sigma[i] * inv % SO_Q);
    }
    return sigma;
}
void so_client_spiral_query(so_client *c, size_t idx_target, uint64_t *query_cv) {
    uint64_t *sigma = query_sigma(c, idx_target);
    encrypt_simple_regev(c, query_cv, sigma, CC_OBJ(CC_QUERY, 0));
    free(sigma);
}
/* counter-based client: seeded wire query with the wire seed and the query number chosen by the caller
 * (row 0 from the wire seed as so_wire_seeded_row0, noise = Gaussian object (CC_QUERY, query_id), stream 1) */
static void chacha_wire_from_sigma(so_client *c, uint64_t *sigma, uint32_t query_id, const uint8_t wire_seed[32], uint8_t *wire);
void so_client_chacha_query_wire(so_client *c, size_t idx_target, uint32_t query_id, const uint8_t wire_seed[32], uint8_t *wire) {
    chacha_wire_from_sigma(c, query_sigma(c, idx_target), query_id, wire_seed, wire);
}
static void chacha_wire_from_sigma(so_client *c, uint64_t *sigma, uint32_t query_id, const uint8_t wire_seed[32], uint8_t *wire) {
    uint64_t *P = xalloc(2 * PL), *e = xalloc(N), *sig_ntt = xalloc(PL);
    so_wire_seeded_row0(wire_seed, P);
    cc_gauss_raw(c, CC_OBJ(CC_QUERY, query_id), 1, e);
    regev_row1_from_row0(c, P, e);
    so_to_ntt(sig_ntt, sigma, 1);
    so_add(P + PL, P + PL, sig_ntt, 1);
    so_wire_query_pack_seeded(wire_seed, P + PL, wire);
    free(sigma); free(P); free(e); free(sig_ntt);
}
/* The same query in its wire form (wire_format.h).  SEEDED: row 0 of the Regev sample (the uniformly random -a,
 * getRegevSample src/client.cpp:141-157) is drawn from a 32-byte seed directly in NTT form, so only the seed and
 * row 1 = a*s + e + sigma travel; FULL: an ordinary query with both rows packed at 56 bits per coefficient. */
void so_client_spiral_query_wire(so_client *c, size_t idx_target, uint32_t kind, uint8_t *wire) {
    uint64_t *sigma = query_sigma(c, idx_target);
    if (kind == SO_WIRE_QUERY_FULL) {
        uint64_t *cv = xalloc(2 * PL);
        encrypt_simple_regev(c, cv, sigma, CC_OBJ(CC_QUERY, 0));
        so_wire_query_pack_full(cv, wire);
        free(cv);
    } else {
        uint8_t seed[SO_WIRE_SEED_BYTES];
        for (int i = 0; i < 4; i++) { uint64_t w = so_rng_next(&c->rng); memcpy(seed + 8 * i, &w, 8); }
        uint64_t *row0 = xalloc(PL), *a_ntt = xalloc(PL), *s_ntt = xalloc(PL), *e = xalloc(N), *e_ntt = xalloc(PL);
        uint64_t *b = xalloc(PL), *sig_ntt = xalloc(PL);
        so_wire_seeded_row0(seed, row0);
        for (size_t z = 0; z < N; z++) {                       /* a = -(row 0) */
            a_ntt[z] = (SO_P - row0[z]) % SO_P;
            a_ntt[N + z] = (SO_B - row0[N + z]) % SO_B;
        }
        noise(c, e, 1);
        so_to_ntt(s_ntt, c->sr, 1); so_to_ntt(e_ntt, e, 1); so_to_ntt(sig_ntt, sigma, 1);
        so_multiply(b, a_ntt, s_ntt, 1, 1, 1);
        so_add(b, b, e_ntt, 1);
        so_add(b, b, sig_ntt, 1);
        so_wire_query_pack_seeded(seed, b, wire);
        free(row0); free(a_ntt); free(s_ntt); free(e); free(e_ntt); free(b); free(sig_ntt);
    }
    free(sigma);
}

/* negacyclic product over Z_q' (what to_ntt_qprime / mul_over_qprime / from_ntt_qprime compute) */
static void negacyclic_mul_acc(uint64_t *res, const uint64_t *a, const uint64_t *b, uint64_t q) {
    for (size_t i = 0; i < N; i++) {
        if (!a[i]) continue;
        for (size_t j = 0; j < N; j++) {
            uint64_t prod = (uint64_t)((u128)a[i] * b[j] % q);
            size_t k = i + j;
            if (k < N) res[k] = (res[k] + prod) % q; else res[k - N] = (res[k - N] + q - prod) % q;
        }
    }
}

void so_client_spiral_decode(so_client *c, const uint64_t *total_resp, uint64_t *out_pt) {   /* :1428-1476 */
    const so_params *p = &c->prm;
    uint64_t qp = so_arb_qprime(p->qp_bits), q_1 = 4 * p->p_db;
    uint64_t *Sp_q = xalloc(SO_N0 * N), *s_prod = xalloc(SO_N0 * SO_N2 * N);
    for (size_t i = 0; i < SO_N0 * N; i++) {               /* to_ntt_qprime's recentering, src/util.cpp:224-229 */
        int64_t a = (int64_t)c->Sp[i];
        if (a >= (int64_t)(SO_Q / 2)) a -= (int64_t)SO_Q;
        Sp_q[i] = (uint64_t)((a + (int64_t)((SO_Q / qp) * qp) + 2 * (int64_t)qp) % (int64_t)qp);
    }
    for (size_t r = 0; r < SO_N0; r++)
        for (size_t cc = 0; cc < SO_N2; cc++)
            negacyclic_mul_acc(&s_prod[(r * SO_N2 + cc) * N], &Sp_q[r * N], &total_resp[cc * N], qp);
    const uint64_t *rest = total_resp + SO_N2 * N;
    for (size_t i = 0; i < SO_N0 * SO_N2 * N; i++) {
        int64_t vf = (int64_t)s_prod[i]; if (vf >= (int64_t)(qp / 2)) vf -= (int64_t)qp;
        int64_t vr = (int64_t)rest[i];   if (vr >= (int64_t)(q_1 / 2)) vr -= (int64_t)q_1;
        uint64_t denom = qp * (q_1 / p->p_db);
        int64_t r = vf * (int64_t)q_1 + vr * (int64_t)qp;
        int64_t sign = r >= 0 ? 1 : -1;
        __int128 res = ((__int128)r + sign * ((int64_t)denom / 2)) / (__int128)denom;
        res = (res + (denom / p->p_db) * p->p_db + 2 * p->p_db) % p->p_db;
        out_pt[i] = (uint64_t)res;
    }
    free(Sp_q); free(s_prod);
}

/* ===========================================================================================
 * SpiralPack / SpiralStreamPack client (testHighRate's client statements, src/testing.cpp:777-1155): keys for an
 * out_n x out_n packed response, packing keys v_W, expansion keys + V, packed or direct-upload query, decoding.
 * Restates: keygen(S, Sp, sr, out_n) src/client.cpp:23-47; get_fresh_public_key_raw_arb / encryptMatrixArbitrary
 * src/testing.cpp:141-197; v_W, V and the query :905-1005; decoding :1086-1118.
 * =========================================================================================== */
so_client *so_pack_client_new(const so_params *prm, uint64_t seed) {
    so_client *c = (so_client *)calloc(1, sizeof(so_client));
    c->prm = *prm; c->rng.s = seed;
    build_cdf(c);
    size_t n = prm->out_n;
    c->Sp = xalloc(n * N); c->sr = xalloc(N);
    for (size_t m = 0; m < N; m++) c->sr[m] = sample_u64(c) % SO_Q;
    for (size_t r = 0; r < n; r++) for (size_t m = 0; m < N; m++) c->Sp[r * N + m] = sample_u64(c) % SO_Q;
    return c;
}
/* the Pack client with counter-based randomness (the statement a CUDA Pack client is to be compared with): keys as
 * so_client_new_chacha (S' has out_n rows), packing keys from objects (CC_PACK_W, i*t_conv + k), direct-upload ciphertexts from
 * (CC_QUERY_DIRECT, ciphertext number) */
so_client *so_pack_client_new_chacha(const so_params *prm, const uint8_t seed[32]) {
    so_client *c = (so_client *)calloc(1, sizeof(so_client));
    c->prm = *prm; c->chacha = 1;
    for (int i = 0; i < 8; i++) c->key[i] = (uint32_t)seed[4 * i] | ((uint32_t)seed[4 * i + 1] << 8) | ((uint32_t)seed[4 * i + 2] << 16) | ((uint32_t)seed[4 * i + 3] << 24);
    build_cdf(c);
    for (int k = 0; k < 128; k++) c->thr[k] = (uint64_t)floor(c->cdf[k] * 9007199254740992.0) + 1;
    size_t n = prm->out_n;
    c->Sp = xalloc(n * N); c->sr = xalloc(N);
    cc_gauss_raw(c, CC_OBJ(CC_KEYS, 0), 0, c->sr);
    for (size_t r = 0; r < n; r++) cc_gauss_raw(c, CC_OBJ(CC_KEYS, 1 + r), 0, &c->Sp[r * N]);
    return c;
}
/* P = [-A ; Sp*A + E], (out_n+1) x m, NTT form   (get_fresh_public_key_raw_arb + to_ntt); obj_base: object of column 0 */
static void fresh_public_key_arb_ntt(so_client *c, uint64_t *P_ntt, size_t m, uint32_t obj_base) {
    size_t n = c->prm.out_n;
    if (c->chacha) {
        uint64_t *a_ntt = xalloc(PL), *e = xalloc(N), *e_ntt = xalloc(PL), *Sp_ntt = xalloc(n * PL), *prod = xalloc(PL);
        so_to_ntt(Sp_ntt, c->Sp, n);
        for (size_t k = 0; k < m; k++) {
            uint64_t *row0 = &P_ntt[k * PL];
            cc_uniform_ntt(c, obj_base + (uint32_t)k, 0, row0);
            for (size_t z = 0; z < N; z++) { a_ntt[z] = (SO_P - row0[z]) % SO_P; a_ntt[N + z] = (SO_B - row0[N + z]) % SO_B; }
            for (size_t r = 0; r < n; r++) {
                cc_gauss_raw(c, obj_base + (uint32_t)k, 1 + (uint32_t)r, e);
                so_to_ntt(e_ntt, e, 1);
                so_multiply(prod, &Sp_ntt[r * PL], a_ntt, 1, 1, 1);
                so_add(&P_ntt[((1 + r) * m + k) * PL], prod, e_ntt, 1);
            }
        }
        free(a_ntt); free(e); free(e_ntt); free(Sp_ntt); free(prod);
        return;
    }
    uint64_t *A = xalloc(m * N), *A_inv = xalloc(m * N), *E = xalloc(n * m * N), *A_ntt = xalloc(m * PL), *E_ntt = xalloc(n * m * PL);
    uint64_t *Sp_ntt = xalloc(n * PL), *Bp = xalloc(n * m * PL);
    so_fill_uniform_raw(A, m * N, &c->rng);
    noise(c, E, n * m);
    so_to_ntt(A_ntt, A, m); so_to_ntt(E_ntt, E, n * m); so_to_ntt(Sp_ntt, c->Sp, n);
    so_multiply(Bp, Sp_ntt, A_ntt, n, 1, m);
    so_add(Bp, E_ntt, Bp, n * m);
    so_invert(A_inv, A, m);
    so_to_ntt(P_ntt, A_inv, m);
    memcpy(P_ntt + m * PL, Bp, n * m * PL * sizeof(uint64_t));
    free(A); free(A_inv); free(E); free(A_ntt); free(E_ntt); free(Sp_ntt); free(Bp);
}
/* W_exp_left: g x (2 x t_exp); W_exp_right: (stopround+1) x (2 x t_exp_right); V: 2 x 2*t_conv; v_W: out_n x ((out_n+1) x t_conv).
 * The first three may be NULL for a direct-upload client. */
void so_pack_client_pub_params(so_client *c, uint64_t *W_exp_left, uint64_t *W_exp_right, uint64_t *V, uint64_t *v_W) {
    const so_params *p = &c->prm;
    size_t n = p->out_n, mc = p->t_conv, g, stop;
    so_pack_expansion_shape(p, &g, &stop);
    uint64_t *s0_ntt = xalloc(PL), *gv = xalloc(mc * N), *gv_ntt = xalloc(mc * PL), *prod = xalloc(mc * PL);
    so_to_ntt(s0_ntt, c->sr, 1);
    so_build_gadget(gv, 1, mc);
    so_to_ntt(gv_ntt, gv, mc);
    so_mul_by_const(prod, s0_ntt, gv_ntt, mc);                          /* s0 * g_vec */
    for (size_t i = 0; i < n; i++) {                                       /* v_W[i] = P + [0 ; AG], AG row i = s0 * g_vec  (:905-912) */
        uint64_t *W = &v_W[i * (n + 1) * mc * PL];
        fresh_public_key_arb_ntt(c, W, mc, CC_OBJ(CC_PACK_W, i * mc));
        so_add(&W[(1 + i) * mc * PL], &W[(1 + i) * mc * PL], prod, mc);
    }
    if (W_exp_left && W_exp_right && V) {
        expansion_keys(c, W_exp_left, g, p->t_exp, CC_W_LEFT);
        expansion_keys(c, W_exp_right, stop + 1, p->t_exp_right, CC_W_RIGHT);
        /* V (:918-931): column i encrypts s0^2 * G[0][i] (i even) or s0 * G[1][i] (i odd), G = gadget(2, 2*t_conv) */
        size_t m = 2 * mc;
        uint64_t *G = xalloc(2 * m * N), *s0sq = xalloc(PL), *cst = xalloc(N), *cst_ntt = xalloc(PL), *sig_ntt = xalloc(PL), *sigma = xalloc(N);
        uint64_t ct[2 * 2 * SO_N];
        so_build_gadget(G, 2, m);
        so_multiply(s0sq, s0_ntt, s0_ntt, 1, 1, 1);
        for (size_t i = 0; i < m; i++) {
            memset(cst, 0, N * sizeof(uint64_t));
            cst[0] = (i % 2 == 0) ? G[i * N] : G[(m + i) * N];
            so_to_ntt(cst_ntt, cst, 1);
            so_multiply(sig_ntt, (i % 2 == 0) ? s0sq : s0_ntt, cst_ntt, 1, 1, 1);
            so_from_ntt(sigma, sig_ntt, 1);
            encrypt_simple_regev(c, ct, sigma, CC_OBJ(CC_V_CONV, i));
            memcpy(&V[(0 * m + i) * PL], ct, PL * sizeof(uint64_t));
            memcpy(&V[(1 * m + i) * PL], ct + PL, PL * sizeof(uint64_t));
        }
        free(G); free(s0sq); free(cst); free(cst_ntt); free(sig_ntt); free(sigma);
    }
    free(s0_ntt); free(gv); free(gv_ntt); free(prod);
}
/* packed single-ciphertext query (:987-1005) */
static uint64_t *pack_query_sigma(so_client *c, size_t idx_target) {
    const so_params *p = &c->prm;
    size_t g, stop; so_pack_expansion_shape(p, &g, &stop);
    size_t fd = p->nu2, ell = p->t_gsw;
    size_t idx_dim0 = idx_target >> fd, idx_further = idx_target & (((size_t)1 << fd) - 1);
    uint32_t bits_per = so_get_bits_per((uint32_t)ell);
    uint64_t *sigma = xalloc(N);
    sigma[2 * idx_dim0] = (SO_Q / p->p_db) % SO_Q;
    for (size_t i = 0; i < fd; i++) {
        uint64_t bit = (idx_further >> i) & 1;
        for (size_t j = 0; j < ell; j++) sigma[2 * (i * ell + j) + 1] = ((uint64_t)1 << (bits_per * j)) * bit;
    }
    uint64_t inv_first = (uint64_t)inv_mod((int64_t)1 << g, (int64_t)SO_Q), inv_rest = (uint64_t)inv_mod((int64_t)1 << (stop + 1), (int64_t)SO_Q);
    for (size_t i = 0; i < N / 2; i++) {
        sigma[2 * i] = (uint64_t)((u128)sigma[2 * i] * inv_first % SO_Q);
        sigma[2 * i + 1] = (uint64_t)((u128)sigma[2 * i + 1] * inv_rest % SO_Q);
    }
    return sigma;
}
void so_pack_client_query(so_client *c, size_t idx_target, uint64_t *query_cv) {
    uint64_t *sigma = pack_query_sigma(c, idx_target);
    encrypt_simple_regev(c, query_cv, sigma, CC_OBJ(CC_QUERY, 0));
    free(sigma);
}
/* counter-based client: the packed query in SEEDED wire form (as so_client_chacha_query_wire) */
void so_pack_client_chacha_query_wire(so_client *c, size_t idx_target, uint32_t query_id, const uint8_t wire_seed[32], uint8_t *wire) {
    chacha_wire_from_sigma(c, pack_query_sigma(c, idx_target), query_id, wire_seed, wire);
}
/* direct upload (:962-985): v_firstdim = 2^nu1 cts (2x1 NTT); v_folding = nu2 x (2 x 2*ell) NTT */
void so_pack_client_query_direct(so_client *c, size_t idx_target, uint64_t *v_firstdim, uint64_t *v_folding) {
    const so_params *p = &c->prm;
    size_t fd = p->nu2, ell = p->t_gsw, dim0 = (size_t)1 << p->nu1;
    size_t idx_dim0 = idx_target >> fd, idx_further = idx_target & (((size_t)1 << fd) - 1);
    uint32_t bits_per = so_get_bits_per((uint32_t)ell);
    uint64_t *sigma = xalloc(N), *cst_ntt = xalloc(PL), *s0_ntt = xalloc(PL), *prod = xalloc(PL);
    uint64_t ct[2 * 2 * SO_N];
    so_to_ntt(s0_ntt, c->sr, 1);
    for (size_t i = 0; i < dim0; i++) {
        memset(sigma, 0, N * sizeof(uint64_t));
        sigma[0] = i == idx_dim0 ? SO_Q / p->p_db : 0;
        encrypt_simple_regev(c, &v_firstdim[i * 2 * PL], sigma, CC_OBJ(CC_QUERY_DIRECT, i));
    }
    for (size_t i = 0; i < fd; i++) {
        uint64_t bit = (idx_further >> i) & 1;
        uint64_t *gsw = &v_folding[i * 2 * 2 * ell * PL];
        for (size_t j = 0; j < ell; j++) {
            uint64_t val = ((uint64_t)1 << (bits_per * j)) * bit;
            memset(sigma, 0, N * sizeof(uint64_t)); sigma[0] = val;
            encrypt_simple_regev(c, ct, sigma, CC_OBJ(CC_QUERY_DIRECT, dim0 + (i * ell + j) * 2 + 1));   /* column 2j+1: val */
            memcpy(&gsw[(0 * 2 * ell + 2 * j + 1) * PL], ct, PL * sizeof(uint64_t));
            memcpy(&gsw[(1 * 2 * ell + 2 * j + 1) * PL], ct + PL, PL * sizeof(uint64_t));
            so_to_ntt(cst_ntt, sigma, 1);
            so_multiply(prod, s0_ntt, cst_ntt, 1, 1, 1);
            so_from_ntt(sigma, prod, 1);
            encrypt_simple_regev(c, ct, sigma, CC_OBJ(CC_QUERY_DIRECT, dim0 + (i * ell + j) * 2));       /* column 2j: s0 * val */
            memcpy(&gsw[(0 * 2 * ell + 2 * j) * PL], ct, PL * sizeof(uint64_t));
            memcpy(&gsw[(1 * 2 * ell + 2 * j) * PL], ct + PL, PL * sizeof(uint64_t));
        }
    }
    free(sigma); free(cst_ntt); free(s0_ntt); free(prod);
}
/* total_resp: (out_n+1) x out_n raw -> out_pt: out_n x out_n polynomials (item (i,j) = plane i*out_n + j)  (:1086-1118) */
void so_pack_client_decode(so_client *c, const uint64_t *total_resp, uint64_t *out_pt) {
    const so_params *p = &c->prm;
    size_t n = p->out_n;
    uint64_t qp = so_arb_qprime(p->qp_bits), q_1 = 4 * p->p_db;
    uint64_t *Sp_q = xalloc(n * N), *s_prod = xalloc(n * n * N);
    for (size_t i = 0; i < n * N; i++) {
        int64_t a = (int64_t)c->Sp[i];
        if (a >= (int64_t)(SO_Q / 2)) a -= (int64_t)SO_Q;
        Sp_q[i] = (uint64_t)((a + (int64_t)((SO_Q / qp) * qp) + 2 * (int64_t)qp) % (int64_t)qp);
    }
    for (size_t r = 0; r < n; r++)
        for (size_t cc = 0; cc < n; cc++)
            negacyclic_mul_acc(&s_prod[(r * n + cc) * N], &Sp_q[r * N], &total_resp[cc * N], qp);
    const uint64_t *rest = total_resp + n * N;
    for (size_t i = 0; i < n * n * N; i++) {
        int64_t vf = (int64_t)s_prod[i]; if (vf >= (int64_t)(qp / 2)) vf -= (int64_t)qp;
        int64_t vr = (int64_t)rest[i];   if (vr >= (int64_t)(q_1 / 2)) vr -= (int64_t)q_1;
        uint64_t denom = qp * (q_1 / p->p_db);
        int64_t r = vf * (int64_t)q_1 + vr * (int64_t)qp;
        int64_t sign = r >= 0 ? 1 : -1;
        __int128 res = ((__int128)r + sign * ((int64_t)denom / 2)) / (__int128)denom;
        res = (res + (denom / p->p_db) * p->p_db + 2 * p->p_db) % p->p_db;
        out_pt[i] = (uint64_t)res;
    }
    free(Sp_q); free(s_prod);
}
