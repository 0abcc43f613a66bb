/*
 * golden_cases.c - seeded input generators + oracle evaluation for the parity cases.
 *
 * TEST INFRASTRUCTURE ONLY (part of liboracle.so).  One table of cases is shared by
 *   - oracle/ref_golden.cpp : feeds the SAME inputs to the unmodified reference (oracle/_ref),
 *     checks reference == oracle in-process and writes tests/golden/ref_digests_<cfg>.json;
 *   - tests/test_oracle_golden.py : recomputes the oracle digests on CPU and compares them with
 *     the committed reference digests (pins the oracle without /root/reference);
 *   - tests/test_gpu_parity.py : feeds the same inputs to the CUDA path through the C-ABI.
 * Inputs are uniform random ring elements (not valid encryptions): the arithmetic being pinned
 * is exact modular arithmetic, for which uniform inputs are the hardest case.
 */
#include "golden_cases.h"
#include <stdlib.h>
#include <string.h>

#define N SO_N
#define PL (2 * (size_t)SO_N)

static uint64_t *xalloc(size_t words) {
    /* 64-byte aligned: the reference's AVX paths use aligned loads on query/DB buffers */
    size_t bytes = ((words ? words : 1) * sizeof(uint64_t) + 63) & ~(size_t)63;
    uint64_t *p = (uint64_t *)aligned_alloc(64, bytes);
    if (!p) abort();
    memset(p, 0, bytes);
    return p;
}
static void set_in(so_case_io *io, int k, size_t words) { io->in[k] = xalloc(words); io->in_words[k] = words; }
static void set_out(so_case_io *io, size_t words, int kind) { io->out = xalloc(words); io->out_words = words; io->out_kind = kind; }

static void fill_packed(uint64_t *out, size_t n, so_rng *r) {
    for (size_t i = 0; i < n; i++) out[i] = (so_rng_next(r) % SO_P) | ((so_rng_next(r) % SO_B) << 32);
}
/* raw coefficients with a sprinkling of zeros and Q-1 so the Q - 0 = Q quirk and the sign
 * boundaries of the signed decomposition are exercised */
static void fill_raw_edgy(uint64_t *out, size_t n, so_rng *r) {
    so_fill_uniform_raw(out, n, r);
    for (size_t i = 0; i < n; i += 37) out[i] = 0;
    for (size_t i = 5; i < n; i += 101) out[i] = SO_Q - 1;
    for (size_t i = 11; i < n; i += 211) out[i] = (so_rng_next(r) & 0xffff);
}

/* modswitch (src/spiral.cpp:40-78) computes round((long double)v * q' / Q) in x87 arithmetic: the product is rounded to
 * a 64-bit significand, then the quotient, then roundl.  Uniform inputs never get close enough to k + 1/2 for the
 * double rounding to show (probability ~2^-45 per coefficient), so the case plants coefficients built to sit just below
 * a half: P' = M * 2^s (a representable product) with P' mod Q = (Q-1)/2 - t for small t, and val = round(P' / q')
 * whenever val * q' really rounds to P'.  About 40 % of them round differently from exact integer rounding. */
typedef unsigned __int128 gc_u128;
static uint64_t gc_invmod(uint64_t a, uint64_t m) {
    __int128 t = 0, nt = 1, r = m, nr = a % m;
    while (nr) { __int128 q = r / nr, tmp = t - q * nt; t = nt; nt = tmp; tmp = r - q * nr; r = nr; nr = tmp; }
    if (t < 0) t += m;
    return (uint64_t)t;
}
static size_t fill_modswitch_adversarial(uint64_t *out, size_t want, uint64_t qp, so_rng *r) {
    size_t n = 0;
    if (qp == 0 || qp % SO_P == 0 || qp % SO_B == 0) return 0;
    for (unsigned pass = 0; pass < 4096 && n < want; pass++)
        for (unsigned s = 1; s <= 37 && n < want; s++) {
            uint64_t inv2s = gc_invmod((uint64_t)(((gc_u128)1 << s) % SO_Q), SO_Q);
            uint64_t t = so_rng_next(r) % 3000;
            uint64_t R = (SO_Q - 1) / 2 - t, M0 = (uint64_t)((gc_u128)R * inv2s % SO_Q);
            for (uint64_t k = 0; k < 300 && n < want; k++) {
                gc_u128 M = (gc_u128)M0 + (gc_u128)k * SO_Q;
                if (M < ((gc_u128)1 << 63) || M >= ((gc_u128)1 << 64)) continue;
                gc_u128 Pp = M << s, val = (Pp + qp / 2) / qp, P = val * qp;
                if (val > SO_Q) continue;
                unsigned L = 0; for (gc_u128 v = P; v; v >>= 1) L++;
                if (L != 64 + s) continue;
                gc_u128 m = P >> s, rem = P & (((gc_u128)1 << s) - 1), half = (gc_u128)1 << (s - 1);
                if (rem > half || (rem == half && (m & 1))) m++;
                if ((m << s) == Pp) out[n++] = (uint64_t)val;
            }
        }
    return n;
}

static const char *NAMES[SO_CASE_COUNT] = {
    "ntt_forward", "ntt_inverse", "to_ntt", "from_ntt", "multiply", "automorph", "gadget_invert",
    "rescale", "reorient_ciphertexts", "first_dim", "ntt_inv_crt_lift", "split_and_crt", "fold_one",
    "expand_full", "expand_stopround", "scal_to_mat", "regev_to_gsw", "load_db", "convert_db",
    "reorient_dim1", "first_dim_pack", "fold_dim1", "regev_to_simple_gsw", "pack", "to_ntt_no_reduce", "modswitch",
};
int so_case_count(void) { return SO_CASE_COUNT; }
const char *so_case_name(int id) { return (id >= 0 && id < SO_CASE_COUNT) ? NAMES[id] : "?"; }

/* fixed small shapes (kept in one place so harness, CPU tests and GPU tests agree) */
void so_case_shape(int id, const so_params *p, so_case_shape_t *s) {
    memset(s, 0, sizeof(*s));
    s->npolys = 4; s->dim0 = 32; s->num_per = 4; s->g = 3; s->stopround = 0; s->max_bits_right = 0;
    switch (id) {
    case SO_CASE_FOLD_ONE: s->num_per = 2; s->cur_dim = 1; break;
    case SO_CASE_SPLIT_AND_CRT: case SO_CASE_NTT_INV_CRT: s->num_per = 2; break;
    case SO_CASE_EXPAND_STOP: s->g = 4; s->stopround = 2; s->max_bits_right = 2; break;
    case SO_CASE_LOAD_DB: s->dim0 = 4; s->num_per = 2; break;
    case SO_CASE_REORIENT: s->dim0 = 4; break;
    case SO_CASE_FOLD_DIM1: s->num_per = 4; break;
    default: break;
    }
    (void)p;
}

void so_case_make_inputs(int id, const so_params *p, uint64_t seed, so_case_io *io) {
    memset(io, 0, sizeof(*io));
    so_rng r = {seed * 0x100 + (uint64_t)id};
    so_case_shape_t s; so_case_shape(id, p, &s);
    size_t m2 = (size_t)SO_N1 * p->t_gsw;
    switch (id) {
    case SO_CASE_NTT_FWD: case SO_CASE_NTT_INV: case SO_CASE_FROM_NTT:
        set_in(io, 0, s.npolys * PL); so_fill_uniform_ntt(io->in[0], s.npolys, &r); break;
    case SO_CASE_TO_NTT: case SO_CASE_AUTOMORPH: case SO_CASE_RESCALE:
        set_in(io, 0, s.npolys * N); fill_raw_edgy(io->in[0], s.npolys * N, &r); break;
    case SO_CASE_TO_NTT_NR:
        set_in(io, 0, s.npolys * N); so_fill_uniform_mod(io->in[0], s.npolys * N, 1ull << 29, &r); break;
    case SO_CASE_MODSWITCH: {                                   /* furtherDimsLocals.cts: n1 x n2 raw, values in [0, Q] */
        size_t n = SO_N1 * SO_N2 * N;
        set_in(io, 0, n); fill_raw_edgy(io->in[0], n, &r);
        io->in[0][1] = SO_Q; io->in[0][2] = SO_Q / 2; io->in[0][3] = SO_Q / 2 + 1; io->in[0][4] = 1;
        uint64_t *adv = xalloc(1024);
        size_t cnt = fill_modswitch_adversarial(adv, 1024, so_arb_qprime(p->qp_bits), &r);
        for (size_t i = 0; i < cnt; i++) io->in[0][8 + 5 * i] = adv[i];      /* 5 and qp_bits coprime to 64: all word phases */
        free(adv);
        break; }
    case SO_CASE_MULTIPLY:
        set_in(io, 0, 2 * 3 * PL); so_fill_uniform_ntt(io->in[0], 6, &r);
        set_in(io, 1, 3 * 2 * PL); so_fill_uniform_ntt(io->in[1], 6, &r); break;
    case SO_CASE_GADGET_INVERT:
        set_in(io, 0, 2 * N); fill_raw_edgy(io->in[0], 2 * N, &r); break;
    case SO_CASE_REORIENT:
        set_in(io, 0, s.dim0 * SO_N1 * 2 * PL); so_fill_uniform_ntt(io->in[0], s.dim0 * SO_N1 * 2, &r); break;
    case SO_CASE_FIRST_DIM:
        set_in(io, 0, s.dim0 * 2 * 4 * N); fill_packed(io->in[0], s.dim0 * 2 * 4 * N, &r);
        for (size_t i = 3; i < s.dim0 * 2 * 4 * N; i += 4) io->in[0][i] = 0;      /* r = 3 padding lane */
        set_in(io, 1, s.dim0 * s.num_per * 4 * N); fill_packed(io->in[1], s.dim0 * s.num_per * 4 * N, &r); break;
    case SO_CASE_NTT_INV_CRT:
        set_in(io, 0, s.num_per * 6 * PL); so_fill_uniform_ntt(io->in[0], s.num_per * 6, &r); break;
    case SO_CASE_SPLIT_AND_CRT:
        set_in(io, 0, s.num_per * 6 * N); fill_raw_edgy(io->in[0], s.num_per * 6 * N, &r); break;
    case SO_CASE_FOLD_ONE:
        set_in(io, 0, 2 * s.num_per * 6 * N); fill_raw_edgy(io->in[0], 2 * s.num_per * 6 * N, &r);
        set_in(io, 1, (s.cur_dim + 1) * SO_N1 * m2 * PL); fill_packed(io->in[1], io->in_words[1], &r);
        set_in(io, 2, (s.cur_dim + 1) * SO_N1 * m2 * PL); fill_packed(io->in[2], io->in_words[2], &r); break;
    case SO_CASE_EXPAND_FULL: case SO_CASE_EXPAND_STOP: {
        size_t cnt = (size_t)1 << s.g;
        set_in(io, 0, cnt * 2 * PL); so_fill_uniform_ntt(io->in[0], 2, &r);
        set_in(io, 1, s.g * 2 * p->t_exp * PL); so_fill_uniform_ntt(io->in[1], s.g * 2 * p->t_exp, &r);
        set_in(io, 2, s.g * 2 * p->t_exp_right * PL); so_fill_uniform_ntt(io->in[2], s.g * 2 * p->t_exp_right, &r);
        break; }
    case SO_CASE_SCAL_TO_MAT:
        set_in(io, 0, 2 * PL); so_fill_uniform_ntt(io->in[0], 2, &r);
        set_in(io, 1, SO_N1 * 2 * p->t_conv * PL); so_fill_uniform_ntt(io->in[1], SO_N1 * 2 * p->t_conv, &r); break;
    case SO_CASE_REGEV_TO_GSW:
        set_in(io, 0, p->t_gsw * 2 * PL); so_fill_uniform_ntt(io->in[0], p->t_gsw * 2, &r);
        set_in(io, 1, SO_N1 * 2 * p->t_conv * PL); so_fill_uniform_ntt(io->in[1], SO_N1 * 2 * p->t_conv, &r);
        set_in(io, 2, SO_N1 * 2 * p->t_conv * PL); so_fill_uniform_ntt(io->in[2], SO_N1 * 2 * p->t_conv, &r); break;
    case SO_CASE_LOAD_DB:
        set_in(io, 0, s.dim0 * s.num_per * 4 * N); so_fill_uniform_mod(io->in[0], io->in_words[0], p->p_db, &r); break;
    case SO_CASE_CONVERT_DB:
        set_in(io, 0, s.dim0 * s.num_per * PL); so_fill_uniform_ntt(io->in[0], s.dim0 * s.num_per, &r); break;
    case SO_CASE_REORIENT_DIM1:
        set_in(io, 0, 2 * s.dim0 * 2 * PL); so_fill_uniform_ntt(io->in[0], 2 * s.dim0 * 2, &r); break;
    case SO_CASE_FIRST_DIM_PACK:
        set_in(io, 0, s.dim0 * 2 * N); fill_packed(io->in[0], io->in_words[0], &r);
        set_in(io, 1, s.dim0 * s.num_per * N); fill_packed(io->in[1], io->in_words[1], &r); break;
    case SO_CASE_FOLD_DIM1:
        set_in(io, 0, s.num_per * 2 * N); fill_raw_edgy(io->in[0], io->in_words[0], &r);
        set_in(io, 1, 2 * 2 * 2 * p->t_gsw * PL); so_fill_uniform_ntt(io->in[1], 2 * 2 * 2 * p->t_gsw, &r);
        set_in(io, 2, 2 * 2 * 2 * p->t_gsw * PL); so_fill_uniform_ntt(io->in[2], 2 * 2 * 2 * p->t_gsw, &r); break;
    case SO_CASE_REGEV_TO_SGSW:
        set_in(io, 0, (2 * (2 * p->t_gsw) + 2) * 2 * PL); so_fill_uniform_ntt(io->in[0], (2 * (2 * p->t_gsw) + 2) * 2, &r);
        set_in(io, 1, 2 * 2 * p->t_conv * PL); so_fill_uniform_ntt(io->in[1], 2 * 2 * p->t_conv, &r); break;
    case SO_CASE_PACK: {
        size_t n = p->out_n;
        set_in(io, 0, n * n * 2 * N); fill_raw_edgy(io->in[0], io->in_words[0], &r);
        set_in(io, 1, n * (n + 1) * p->t_conv * PL); so_fill_uniform_ntt(io->in[1], n * (n + 1) * p->t_conv, &r);
        break; }
    default: break;
    }
}

void so_case_run_oracle(int id, const so_params *p, so_case_io *io) {
    so_case_shape_t s; so_case_shape(id, p, &s);
    size_t m2 = (size_t)SO_N1 * p->t_gsw;
    switch (id) {
    case SO_CASE_NTT_FWD:
        set_out(io, io->in_words[0], SO_KIND_NTT); memcpy(io->out, io->in[0], io->in_words[0] * 8);
        for (size_t i = 0; i < s.npolys; i++) so_ntt_forward(&io->out[i * PL]);
        break;
    case SO_CASE_NTT_INV:
        set_out(io, io->in_words[0], SO_KIND_NTT); memcpy(io->out, io->in[0], io->in_words[0] * 8);
        for (size_t i = 0; i < s.npolys; i++) so_ntt_inverse(&io->out[i * PL]);
        break;
    case SO_CASE_TO_NTT:
        set_out(io, s.npolys * PL, SO_KIND_NTT); so_to_ntt(io->out, io->in[0], s.npolys); break;
    case SO_CASE_TO_NTT_NR:
        set_out(io, s.npolys * PL, SO_KIND_NTT); so_to_ntt_no_reduce(io->out, io->in[0], s.npolys); break;
    case SO_CASE_MODSWITCH:
        set_out(io, so_packed_words(io->in_words[0], p->qp_bits), SO_KIND_RAW);
        so_modswitch(io->out, io->in[0], io->in_words[0], p->qp_bits); break;
    case SO_CASE_FROM_NTT:
        set_out(io, s.npolys * N, SO_KIND_RAW); so_from_ntt(io->out, io->in[0], s.npolys); break;
    case SO_CASE_MULTIPLY:
        set_out(io, 2 * 2 * PL, SO_KIND_NTT); so_multiply(io->out, io->in[0], io->in[1], 2, 3, 2); break;
    case SO_CASE_AUTOMORPH:
        set_out(io, s.npolys * N, SO_KIND_RAW); so_automorph(io->out, io->in[0], s.npolys, N / 4 + 1); break;
    case SO_CASE_GADGET_INVERT:
        set_out(io, 2 * p->t_conv * N, SO_KIND_RAW); so_gadget_invert(io->out, io->in[0], 2 * p->t_conv, 2, 1); break;
    case SO_CASE_RESCALE:
        set_out(io, s.npolys * N, SO_KIND_RAW);
        so_get_rescaled(io->out, io->in[0], 2 * N, SO_Q, so_arb_qprime(p->qp_bits));
        so_get_rescaled(&io->out[2 * N], &io->in[0][2 * N], 2 * N, SO_Q, 4 * p->p_db);
        break;
    case SO_CASE_REORIENT:
        set_out(io, s.dim0 * 2 * 4 * N, SO_KIND_PACKED); so_reorient_ciphertexts(io->out, io->in[0], s.dim0, 4); break;
    case SO_CASE_FIRST_DIM:
        set_out(io, s.num_per * 6 * PL, SO_KIND_NTT);
        so_multiply_query_by_database(io->out, io->in[0], io->in[1], s.dim0, s.num_per); break;
    case SO_CASE_NTT_INV_CRT: {
        uint64_t *tmp = xalloc(io->in_words[0]); memcpy(tmp, io->in[0], io->in_words[0] * 8);
        set_out(io, s.num_per * 6 * N, SO_KIND_RAW); so_ntt_inv_and_crt_lift(io->out, tmp, s.num_per); free(tmp);
        break; }
    case SO_CASE_SPLIT_AND_CRT:
        set_out(io, s.num_per * m2 * 2 * PL, SO_KIND_NTT); so_split_and_crt(io->out, io->in[0], s.num_per, p->t_gsw); break;
    case SO_CASE_FOLD_ONE: {
        uint64_t *tmp = xalloc(io->in_words[0]); memcpy(tmp, io->in[0], io->in_words[0] * 8);
        so_fold_one_further_dimension(s.cur_dim, s.num_per, io->in[1], io->in[2], tmp, p->t_gsw);
        set_out(io, s.num_per * 6 * N, SO_KIND_RAW); memcpy(io->out, tmp, s.num_per * 6 * N * 8); free(tmp);
        break; }
    case SO_CASE_EXPAND_FULL: case SO_CASE_EXPAND_STOP:
        set_out(io, io->in_words[0], SO_KIND_NTT); memcpy(io->out, io->in[0], io->in_words[0] * 8);
        so_expand_improved(io->out, s.g, p->t_exp, io->in[1], io->in[2], p->t_exp_right, s.max_bits_right, s.stopround);
        break;
    case SO_CASE_SCAL_TO_MAT:
        set_out(io, SO_N1 * 2 * PL, SO_KIND_NTT); so_scal_to_mat(io->out, io->in[0], io->in[1], p->t_conv); break;
    case SO_CASE_REGEV_TO_GSW:
        set_out(io, SO_N1 * m2 * PL, SO_KIND_NTT);
        so_regev_to_gsw(io->out, io->in[0], p->t_conv, p->t_gsw, io->in[1], io->in[2]); break;
    case SO_CASE_LOAD_DB: {
        uint32_t nu1 = 0, nu2 = 0; while (((size_t)1 << nu1) < s.dim0) nu1++; while (((size_t)1 << nu2) < s.num_per) nu2++;
        set_out(io, s.dim0 * s.num_per * 4 * N, SO_KIND_PACKED); so_load_db(io->out, io->in[0], nu1, nu2, p->p_db);
        break; }
    case SO_CASE_CONVERT_DB:
        set_out(io, s.dim0 * s.num_per * N, SO_KIND_PACKED);
        so_convert_db(io->out, io->in[0], s.dim0 * s.num_per, s.dim0, s.num_per); break;
    case SO_CASE_REORIENT_DIM1:
        set_out(io, s.dim0 * 2 * N, SO_KIND_PACKED); so_reorient_ciphertexts_dim1(io->out, io->in[0], s.dim0, 2); break;
    case SO_CASE_FIRST_DIM_PACK:
        set_out(io, s.num_per * 2 * PL, SO_KIND_NTT); so_fast_multiply_dim1(io->out, io->in[1], io->in[0], s.dim0, s.num_per); break;
    case SO_CASE_FOLD_DIM1: {
        uint64_t *tmp = xalloc(io->in_words[0]); memcpy(tmp, io->in[0], io->in_words[0] * 8);
        so_fold_ciphertexts_dim1(tmp, s.num_per, io->in[1], io->in[2], p->t_gsw);
        set_out(io, 2 * N, SO_KIND_RAW); memcpy(io->out, tmp, 2 * N * 8); free(tmp);
        break; }
    case SO_CASE_REGEV_TO_SGSW:
        set_out(io, 2 * 2 * 2 * p->t_gsw * PL, SO_KIND_NTT);
        so_regev_to_simple_gsw(io->out, io->in[0], io->in[1], p->t_conv, p->t_gsw, 2, 2, 1); break;
    case SO_CASE_PACK:
        set_out(io, (p->out_n + 1) * p->out_n * PL, SO_KIND_NTT);
        so_pack(io->out, p->out_n, p->t_conv, io->in[0], io->in[1]); break;
    default: break;
    }
}

/* canonical digest: raw exact; NTT / packed reduced modulo the respective prime first */
uint64_t so_digest_kind(const uint64_t *w, size_t words, int kind) {
    if (kind == SO_KIND_RAW) return so_fnv1a64(w, words);
    if (kind == SO_KIND_NTT) return so_fnv1a64_ntt(w, words / PL);
    uint64_t *tmp = xalloc(words);
    for (size_t i = 0; i < words; i++) tmp[i] = ((w[i] & 0xffffffffull) % SO_P) | (((w[i] >> 32) % SO_B) << 32);
    uint64_t h = so_fnv1a64(tmp, words);
    free(tmp);
    return h;
}
uint64_t so_case_digest(const so_case_io *io) { return so_digest_kind(io->out, io->out_words, io->out_kind); }

void so_case_free(so_case_io *io) {
    for (int k = 0; k < SO_CASE_MAX_IN; k++) free(io->in[k]);
    free(io->out);
    memset(io, 0, sizeof(*io));
}
