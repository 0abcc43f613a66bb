/*
 * spiral_oracle.h - CPU restatement of the Spiral server-side query-answering path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is linked into, imported by or executed
 * from the product (spiral_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load liboracle.so, and only as the checker.
 *
 * Every function restates one function of the reference (menonsamir/spiral, C++/AVX) in plain
 * C with RUNTIME scheme parameters (the reference fixes them with -D macros,
 * include/values.h:78-93).  The file:line each one follows is given at its definition in
 * spiral_oracle.c.  Parity is PINNED: tests/test_oracle_vs_ref.py checks these functions
 * against golden digests produced by the unmodified reference itself (oracle/_ref, built by
 * oracle/Makefile from /root/reference; generator = oracle/ref_golden.cpp).
 *
 * Data layouts are the reference's own (include/poly.h:24-64):
 *   NTT form : data[(r*cols + c)*2*N + n*N + z]   n = 0 -> mod p, n = 1 -> mod b, one u64 each
 *   raw form : data[(r*cols + c)*N + z]           u64 in [0, Q]
 */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_N 2048u                         /* include/values.h:11  poly_len            */
#define SO_LOGN 11u
#define SO_P 268369921ull                  /* include/values.h:13  p_i                 */
#define SO_B 249561089ull                  /* include/values.h:21  b_i                 */
#define SO_Q 66974689739603969ull          /* include/values.h:41  Q_i = p*b (56 bit)  */
#define SO_LOGQ 56u
#define SO_N0 2u                           /* include/values.h:67-69                   */
#define SO_N1 3u
#define SO_N2 2u

typedef struct so_params {
    uint32_t nu1;          /* num_expansions: first dimension is 2^nu1            */
    uint32_t nu2;          /* further_dims                                         */
    uint32_t t_gsw;        /* TGSW   (m2 = n1 * t_gsw)                             */
    uint32_t t_conv;       /* TCONV  (m_conv)                                      */
    uint32_t t_exp;        /* TEXP   (m_exp, left key-switch gadget length)        */
    uint32_t t_exp_right;  /* TEXPRIGHT                                            */
    uint32_t qp_bits;      /* QPBITS -> arb_qprime = qprime_mods[qp_bits]          */
    uint32_t out_n;        /* OUTN (Pack variants)                                 */
    uint64_t p_db;         /* PVALUE                                               */
} so_params;

/* ---- constants / scalar arithmetic ------------------------------------------------------ */
const uint64_t *so_tables(void);                       /* 8 x 2048, same layout as src/constants.cpp:16 */
uint64_t so_arb_qprime(uint32_t qp_bits);              /* include/values.h:74-76 */
uint32_t so_get_bits_per(uint32_t dim);                /* include/util.h:34-38   */
uint64_t so_barrett_coeff(uint64_t val, int n);        /* include/poly.h:137-153 */
uint64_t so_crt_compose(uint64_t x, uint64_t y);       /* src/poly.cpp:344-353   */
uint64_t so_rescale(uint64_t a, uint64_t inp_mod, uint64_t out_mod);   /* src/poly.cpp:578-591 */

/* ---- NTT (src/core.cpp:254-514) ---------------------------------------------------------- */
void so_ntt_forward(uint64_t *op);                     /* in place on [2][N]; canonical output */
void so_ntt_inverse(uint64_t *op);

/* ---- MatPoly algebra on flat buffers (src/poly.cpp) ------------------------------------- */
void so_to_ntt(uint64_t *out_ntt, const uint64_t *in_raw, size_t npolys);
void so_to_ntt_no_reduce(uint64_t *out_ntt, const uint64_t *in_raw, size_t npolys);
void so_from_ntt(uint64_t *out_raw, const uint64_t *in_ntt, size_t npolys);
void so_multiply(uint64_t *out, const uint64_t *a, const uint64_t *b, size_t rs, size_t ms, size_t cs);
void so_add(uint64_t *out, const uint64_t *a, const uint64_t *b, size_t npolys);
void so_mul_by_const(uint64_t *out, const uint64_t *single, const uint64_t *a, size_t npolys);
void so_automorph(uint64_t *out_raw, const uint64_t *in_raw, size_t npolys, uint64_t t);
void so_invert(uint64_t *out_raw, const uint64_t *in_raw, size_t npolys);
void so_build_gadget(uint64_t *G_raw, size_t rows, size_t cols);                 /* src/util.cpp:89-112  */
void so_gadget_invert(uint64_t *out_raw, const uint64_t *in_raw, size_t mx, size_t rdim, size_t cols); /* src/util.cpp:114-150 */
void so_get_rescaled(uint64_t *out, const uint64_t *in, size_t ncoeffs, uint64_t inp_mod, uint64_t out_mod); /* src/poly.cpp:593-601 */
/* bit I/O (src/core.cpp:20-52) and the response bit-packer modswitch (src/spiral.cpp:40-78):
 * round((long double)v * arb_qprime / Q) packed at qp_bits bits per coefficient, n1*n2*N coefficients. */
uint64_t so_read_arbitrary_bits(const uint64_t *p, size_t bit_offs, size_t num_bits);
void so_write_arbitrary_bits(uint64_t *p, uint64_t val, size_t bit_offs, size_t num_bits);
uint64_t so_modswitch_coeff(uint64_t val, uint64_t qprime);          /* x87 arithmetic restated in integers */
uint64_t so_modswitch_coeff_x87(uint64_t val, uint64_t qprime);      /* the literal long double statements (x86 only; self-check) */
void so_modswitch(uint64_t *out_words, const uint64_t *in_raw, size_t ncoeffs, uint32_t qp_bits);
size_t so_packed_words(size_t ncoeffs, uint32_t bits);               /* u64 words holding ncoeffs values of `bits` bits */

/* ---- Spiral / SpiralStream server path (src/spiral.cpp) --------------------------------- */
void so_encode_plaintext(uint64_t *out_raw, const uint64_t *pt, size_t ncoeffs, uint64_t p_db);  /* :1116-1127 */
void so_load_db(uint64_t *B, const uint64_t *pts, uint32_t nu1, uint32_t nu2, uint64_t p_db);    /* :1028-1172 */
void so_reorient_ciphertexts(uint64_t *out, const uint64_t *inp, size_t dim0, size_t n1_padded);/* :410-433  */
void so_multiply_query_by_database(uint64_t *out, const uint64_t *reoriented, const uint64_t *db,
                                   size_t dim0, size_t num_per);                                  /* :628-999  */
void so_ntt_inv_and_crt_lift(uint64_t *cts_raw, uint64_t *scratch_ntt, size_t num_per);          /* :437-453  */
void so_split_and_crt(uint64_t *out, const uint64_t *in, size_t num_per, uint32_t t_gsw);        /* :270-341  */
void so_reorient_C(uint64_t *out, const uint64_t *inp, size_t num_per, uint32_t t_gsw);          /* :345-384  */
void so_reorient_Q(uint64_t *out, const uint64_t *inp, uint32_t t_gsw);                           /* :388-400  */
void so_cpu_mul_query_by_ct(uint64_t *C_next, const uint64_t *Q, const uint64_t *C, size_t num_per, uint32_t t_gsw); /* :464-582 */
void so_cpu_crt(uint64_t *out, const uint64_t *inp, size_t num_polys);                            /* :586-593  */
void so_cpu_crt_to_ucompressed_and_ntt(uint64_t *out, const uint64_t *inp, size_t num_polys);    /* :597-609  */
/* cts: raw, 2*num_per cts of n1 x n2 on entry (num_per = count AFTER halving), folded in place */
void so_fold_one_further_dimension(size_t cur_dim, size_t num_per, const uint64_t *q_reor,
                                   const uint64_t *q_neg_reor, uint64_t *cts_raw, uint32_t t_gsw); /* :1349-1410 */
/* cv: 2^g cts of 2x1 NTT (first filled, rest zero); W_left[r]: 2 x t_exp, W_right[r]: 2 x t_exp_right (NTT) */
void so_expand_improved(uint64_t *cv, size_t g, uint32_t t_exp, const uint64_t *W_left,
                        const uint64_t *W_right, uint32_t t_exp_right, size_t max_bits_right,
                        size_t stopround);                                                        /* :1664-1743 */
void so_scal_to_mat(uint64_t *out_reg, const uint64_t *cv, const uint64_t *W, uint32_t t_conv);  /* :1850-1885 */
void so_regev_to_gsw(uint64_t *out, const uint64_t *cv_v, uint32_t t_conv, uint32_t t,
                     const uint64_t *W, const uint64_t *V);                                       /* :1985-2025 */
/* G2 - Q on raw coefficients then NTT; q_crtd: nu2 x (n1 x m2) raw. (:2361-2378) */
void so_gsw_negate(uint64_t *q_neg_ntt, const uint64_t *q_crtd, uint32_t nu2, uint32_t t_gsw);

/* Whole server pipeline for one query (server-side statements of runConversionImproved :2040-2335,
 * process_crtd_query :2337-2406, process_query_fast :1584-1629 and check_final's modulus switch
 * :1441-1447).  query_cv: 2x1 NTT.  pub params NTT.  Returns 0 on success.
 *   final_ct_raw : n1 x n2 raw (furtherDimsLocals.cts after folding)
 *   total_resp   : n1 x n2 raw, row 0 mod arb_qprime, rows 1.. mod 4*p_db
 * dbg_first_dim_raw (optional, may be NULL): num_per cts raw after nttInvAndCrtLift. */
int so_spiral_answer(const so_params *prm, const uint64_t *query_cv, const uint64_t *W_exp_left,
                     const uint64_t *W_exp_right, const uint64_t *W_conv, const uint64_t *V_conv,
                     const uint64_t *B, uint64_t *final_ct_raw, uint64_t *total_resp,
                     uint64_t *dbg_first_dim_raw);
/* g / stopround exactly as runConversionImproved derives them (:2079-2085) */
void so_spiral_expansion_shape(const so_params *prm, size_t *g, size_t *stopround);

/* ---- SpiralPack / SpiralStreamPack server path (src/testing.cpp) ------------------------ */
void so_convert_db(uint64_t *db_buf, const uint64_t *db_ntt, size_t count, size_t dim0, size_t num_per); /* :316-340 */
void so_reorient_ciphertexts_dim1(uint64_t *out, const uint64_t *v_firstdim, size_t dim0, size_t idx_factor); /* :342-362 */
void so_fast_multiply_dim1(uint64_t *out, const uint64_t *db, const uint64_t *v_firstdim, size_t dim0, size_t num_per); /* :364-593 */
/* v_cts: 2^nu2 cts (2x1 raw); v_folding[d], v_folding_neg[d]: 2 x 2*ell NTT */
void so_fold_ciphertexts_dim1(uint64_t *v_cts, size_t count, const uint64_t *v_folding,
                              const uint64_t *v_folding_neg, uint32_t ell);                        /* :596-624 */
void so_regev_to_simple_gsw(uint64_t *v_gsw, const uint64_t *v_inp, const uint64_t *V, uint32_t t_conv,
                            uint32_t ell, uint32_t further_dims, size_t idx_factor, size_t idx_offset); /* :108-140 */
void so_simple_gsw_negate(uint64_t *neg, const uint64_t *v_folding, uint32_t further_dims, uint32_t ell); /* :1027-1032 */
void so_pack(uint64_t *result, uint32_t out_n, uint32_t t_conv, const uint64_t *v_ct, const uint64_t *v_W); /* :198-241 */
void so_pack_expansion_shape(const so_params *prm, size_t *g, size_t *stopround);                 /* :795-798 */
/* Whole Pack-variant server pipeline (testHighRate :1007-1081, server statements only).
 * do_expansion: query is one packed ct (query_cv 2x1 NTT) else direct upload:
 *   v_firstdim (2^nu1 cts 2x1 NTT) and v_folding_direct (nu2 x [2 x 2*ell] NTT).
 * db_planes: out_n^2 buffers in convertDb layout, concatenated.
 * total_resp: (out_n+1) x out_n raw. result_cts (optional): out_n^2 cts 2x1 raw. */
int so_pack_answer(const so_params *prm, int do_expansion, const uint64_t *query_cv,
                   const uint64_t *W_exp_left, const uint64_t *W_exp_right, const uint64_t *V,
                   const uint64_t *v_firstdim, const uint64_t *v_folding_direct, const uint64_t *v_W,
                   const uint64_t *db_planes, uint64_t *total_resp, uint64_t *result_cts);

/* ---- digest used by the golden fixtures -------------------------------------------------- */
uint64_t so_fnv1a64(const uint64_t *words, size_t nwords);
/* same, but each NTT-form word is first reduced mod its prime (layout [poly][2][N]);
 * the reference's AVX2 NTT may leave q instead of 0 (src/core.cpp:342-349). */
uint64_t so_fnv1a64_ntt(const uint64_t *ntt_words, size_t npolys);

/* ---- deterministic input generator shared by tests and oracle/ref_golden.cpp ------------ */
typedef struct so_rng { uint64_t s; } so_rng;
uint64_t so_rng_next(so_rng *r);                                  /* splitmix64 */
void so_fill_uniform_raw(uint64_t *out, size_t ncoeffs, so_rng *r);     /* uniform [0,Q) */
void so_fill_uniform_ntt(uint64_t *out, size_t npolys, so_rng *r);      /* [poly][2][N], uniform [0,p) x [0,b) */
void so_fill_uniform_mod(uint64_t *out, size_t n, uint64_t mod, so_rng *r);

#ifdef __cplusplus
}
#endif
