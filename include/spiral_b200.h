/*
 * spiral_b200.h - C-ABI of libspiral_b200.so: the B200 (sm_100a) implementation of Spiral's
 * server-side query answering.  Plain pointers and sizes only; no torch / C++ types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the reference
 * tree menonsamir/spiral).  The reference has no FFI: its path sits behind C++ free functions
 * with external linkage, so a maintainer binds these symbols from thin definitions of those
 * functions (INTEGRATION.md shows the stubs; spiral_b200/csrc/host_mirror.cpp is that file).
 *
 * Three tiers:
 *   sb200_dev_*     device pointers + stream: the kernels, for callers that keep data in HBM
 *   sb200_<refname> host pointers in the REFERENCE's layouts: one call = H2D, kernels, D2H
 *   sb200_server_*  a resident server: database + public parameters live in HBM, one query in,
 *                   one response out
 *
 * Layouts.  "ref-NTT": uint64_t data[(r*cols+c)*2*2048 + n*2048 + z], n = 0 mod p / 1 mod b
 * (reference include/poly.h:24-64).  "raw": uint64_t data[(r*cols+c)*2048 + z] in [0,Q].
 * "dev-NTT": uint32_t with the same index order (residues are 28-bit).  "PB64": one uint64_t =
 * residue mod p | residue mod b << 32 (reference src/spiral.cpp:429).
 *
 * All functions return 0 on success and a negative code on failure; sb200_last_error() gives the
 * text.  There is NO CPU fallback: without a CUDA device every compute entry fails with
 * SB200_ERR_NO_DEVICE.  `stream` is a cudaStream_t passed as void* (NULL = default stream).
 */
#ifndef SPIRAL_B200_H
#define SPIRAL_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_OK 0
#define SB200_ERR_NO_DEVICE (-1)
#define SB200_ERR_CUDA (-2)
#define SB200_ERR_ARG (-3)
#define SB200_ERR_STATE (-4)

#define SB200_POLY_LEN 2048

/* Scheme parameters: the reference's compile-time -D values (include/values.h:78-93) at run time. */
typedef struct sb200_params {
    uint32_t nu1;          /* num_expansions (argv[1])   */
    uint32_t nu2;          /* further_dims   (argv[2])   */
    uint32_t t_gsw;        /* TGSW      */
    uint32_t t_conv;       /* TCONV     */
    uint32_t t_exp;        /* TEXP      */
    uint32_t t_exp_right;  /* TEXPRIGHT */
    uint32_t qp_bits;      /* QPBITS    */
    uint32_t out_n;        /* OUTN (Pack variants) */
    uint64_t p_db;         /* PVALUE    */
} sb200_params;

/* ---- library -------------------------------------------------------------------------- */
int sb200_init(int device);                 /* cudaSetDevice + twiddle tables for that device */
const char *sb200_last_error(void);
int sb200_abi_version(void);
uint64_t sb200_arb_qprime(uint32_t qp_bits);           /* include/values.h:74-76 */
/* kernels launched by this process since load (our own launches only) */
uint64_t sb200_launch_count(void);
/* names of the distinct kernels launched since the last reset, ';'-separated and sorted; returns the bytes needed
   (tests print the kernel set a shape dispatched to) */
size_t sb200_kernel_log(char *buf, size_t cap);
void sb200_kernel_log_reset(void);
/* timeline trace (profiling): with capacity > 0 the first CTA of every kernel records {ns it was scheduled, ns its
   dependencies were resolved, gridDim.x | gridDim.y << 20 | blockDim.x << 36 | source line of the kernel's prologue << 48}; read
   copies up to max_records x 3 words.
   Works inside replayed CUDA graphs, where ncu's serialised timings say nothing about a dependent chain.  0 = off. */
int sb200_trace_enable(uint32_t capacity);
size_t sb200_trace_read(uint64_t *out, size_t max_records, int reset);

/* ---- tier 1: device pointers ---------------------------------------------------------- */
int sb200_dev_ntt_from_ref(uint32_t *out, const uint64_t *in_ref_ntt, size_t npolys, void *stream);
int sb200_dev_ntt_to_ref(uint64_t *out_ref_ntt, const uint32_t *in, size_t npolys, void *stream);
int sb200_dev_to_ntt(uint32_t *out, const uint64_t *raw, size_t npolys, void *stream);          /* src/poly.cpp:291-329 */
int sb200_dev_from_ntt(uint64_t *raw, const uint32_t *in, size_t npolys, void *stream);        /* src/poly.cpp:357-377 */
int sb200_dev_multiply(uint32_t *out, const uint32_t *a, const uint32_t *b, int rs, int ms, int cs, void *stream); /* src/poly.cpp:34-78 */
int sb200_dev_automorph(uint64_t *out, const uint64_t *in, size_t npolys, uint32_t t, void *stream);   /* src/poly.cpp:240-261 */
int sb200_dev_gadget_ntt(uint32_t *out, const uint64_t *raw, int mx, int rdim, int cols, void *stream); /* src/util.cpp:114-150 + to_ntt */
int sb200_dev_rescale(uint64_t *out, const uint64_t *in, size_t ncoeffs, uint64_t inp_mod, uint64_t out_mod, void *stream); /* src/poly.cpp:578-601 */
/* modswitch: round((long double)v * arb_qprime / Q) in the reference's x87 arithmetic, bit-packed at qp_bits (src/spiral.cpp:40-78) */
int sb200_dev_modswitch(uint64_t *out_words, const uint64_t *cts_raw, size_t ncoeffs, uint32_t qp_bits, void *stream);
/* write_arbitrary_bits over a buffer: n values of `bits` bits -> sb200_packed_words(n, bits) words (src/core.cpp:32-52) */
int sb200_dev_bitpack(uint64_t *out_words, const uint64_t *values, size_t n, uint32_t bits, void *stream);
size_t sb200_packed_words(size_t ncoeffs, uint32_t bits);
/* response wire format: modulus-switched response, row 0 at qp_bits bits, other rows at log2(4*p_db) bits (sizes: src/spiral.cpp:229-232) */
size_t sb200_packed_response_words(size_t row0_coeffs, size_t rest_coeffs, uint32_t qp_bits, uint64_t p_db);
int sb200_dev_pack_response(uint64_t *packed, const uint64_t *total_resp, size_t row0_coeffs, size_t rest_coeffs,
                            uint32_t qp_bits, uint64_t p_db, void *stream);
/* client-side inverse (plain host code; read_arbitrary_bits, src/core.cpp:20-30) */
int sb200_unpack_response(uint64_t *total_resp_host, const uint64_t *packed_host, size_t row0_coeffs, size_t rest_coeffs,
                          uint32_t qp_bits, uint64_t p_db);
/* database: plaintext items (u16 coefficients < p_db, [item][m*2+c][2048]) -> scan layout (load_db, src/spiral.cpp:1028-1172) */
int sb200_dev_db_build(uint64_t *db, const uint16_t *pts, uint32_t nu1, uint32_t nu2, uint32_t p_db,
                       size_t item_begin, size_t item_count, void *stream);
/* database already in the reference layout B[z][ii][c][j][m] (src/spiral.cpp:1139-1153), z_count slices at B_chunk */
int sb200_dev_db_from_reference(uint64_t *db, const uint64_t *B_chunk, size_t dim0, size_t num_per,
                                size_t z_begin, size_t z_count, void *stream);
size_t sb200_db_words(uint32_t nu1, uint32_t nu2);       /* uint64 words of the scan-layout database */
int sb200_dev_reorient_query(uint64_t *out, const uint32_t *cts, size_t dim0, void *stream);   /* src/spiral.cpp:410-433 */
int sb200_dev_first_dim(uint32_t *out, const uint64_t *query, const uint64_t *db, size_t dim0, size_t num_per, void *stream); /* src/spiral.cpp:628-999 */
/* batched first dimension on the tcgen05 tensor cores: multiplyQueryByDatabase (src/spiral.cpp:628-999) for up
   to 16 queries in ONE database pass.  db_tc: limb-tile copy of the scan-layout database (same byte count); q_tc: the batch's
   query tiles (sb200_tc_query_bytes, zero-initialised by the caller); out[q]: as sb200_dev_first_dim. */
int sb200_tc_supported(size_t dim0, size_t num_per);      /* 2*dim0 and 2*num_per must be multiples of 128 */
size_t sb200_tc_query_bytes(size_t dim0, int capacity);
int sb200_dev_db_to_tc(uint8_t *db_tc, const uint64_t *db, size_t dim0, size_t num_per, void *stream);
int sb200_dev_query_to_tc(uint8_t *q_tc, const uint64_t *query, int q, int capacity, size_t dim0, void *stream);
size_t sb200_tc_scratch_bytes(size_t num_per, int count);   /* tile-order results of one pass (transposed into out[q] by a second kernel) */
int sb200_dev_first_dim_tc(uint32_t *const *out, int count, int capacity, const uint8_t *q_tc, const uint8_t *db_tc,
                           size_t dim0, size_t num_per, uint32_t *scratch, void *stream);
size_t sb200_fold_scratch_words(size_t num_per_after, uint32_t t_gsw);   /* uint32 words */
int sb200_dev_fold_round(uint64_t *cts, size_t num_per_after, const uint32_t *q, const uint32_t *q_neg,
                         uint32_t t_gsw, uint32_t *scratch, void *stream);                      /* src/spiral.cpp:1349-1410 */

/* ---- tier 2: host pointers, reference layouts (what the interposed reference functions call) */
int sb200_to_ntt(uint64_t *out_ref_ntt, const uint64_t *raw, size_t npolys);
int sb200_from_ntt(uint64_t *raw, const uint64_t *in_ref_ntt, size_t npolys);
int sb200_ntt_forward(uint64_t *io_ref_ntt, size_t npolys);     /* ntt_forward, src/core.cpp:247 (values mod q) */
int sb200_ntt_inverse(uint64_t *io_ref_ntt, size_t npolys);     /* ntt_inverse, src/core.cpp:419 */
int sb200_multiply(uint64_t *out, const uint64_t *a, const uint64_t *b, int rs, int ms, int cs);
int sb200_automorph(uint64_t *out_raw, const uint64_t *in_raw, size_t npolys, uint32_t t);
int sb200_gadget_invert(uint64_t *out_raw, const uint64_t *in_raw, int mx, int rdim, int cols);   /* src/util.cpp:114-150 */
int sb200_getRescaled(uint64_t *out, const uint64_t *in, size_t ncoeffs, uint64_t inp_mod, uint64_t out_mod);
/* modswitch(furtherDimsLocals.result, furtherDimsLocals.cts): n1 x n2 x 2048 raw coefficients in, 192 * qp_bits words out (src/spiral.cpp:40) */
int sb200_modswitch(uint64_t *out_words, const uint64_t *cts_raw, uint32_t qp_bits);
/* load_db: pts = total_n items of n0*n2 polys, u64 coefficients < p_db (reference generate_random_pt); B in reference layout */
int sb200_load_db(uint64_t *B_ref_layout, const uint64_t *pts, uint32_t nu1, uint32_t nu2, uint64_t p_db);
int sb200_reorientCiphertexts(uint64_t *out, const uint64_t *inp_ref_ntt, size_t dim0, size_t n1_padded);
int sb200_multiplyQueryByDatabase(uint64_t *out_ref_ntt, const uint64_t *reoriented, const uint64_t *database,
                                  size_t dim0, size_t num_per);
/* the same for `count` (<= 16) reoriented queries against one database in ONE tensor-core pass (sb200_dev_first_dim_tc) */
int sb200_multiplyQueryByDatabase_batched(uint64_t *const *out_ref_ntt, const uint64_t *const *reoriented, int count,
                                          const uint64_t *database, size_t dim0, size_t num_per);
int sb200_nttInvAndCrtLiftCiphertexts(uint64_t *cts_raw, const uint64_t *scratch_ref_ntt, size_t num_per);
int sb200_split_and_crt(uint64_t *out_ref_ntt, const uint64_t *in_raw, size_t num_per, uint32_t t_gsw);   /* src/spiral.cpp:270-341 */
/* q / q_neg: the reference's reoriented GSW buffers (reorient_Q layout, stride n1*m2*2*2048 words per dimension) */
int sb200_foldOneFurtherDimension(size_t cur_dim, size_t num_per, const uint64_t *q, const uint64_t *q_neg,
                                  uint64_t *cts_raw, uint32_t t_gsw);
/* cv: 2^g cts (2x1 ref-NTT); W_left: g x (2 x t_exp); W_right: g x (2 x t_exp_right) */
int sb200_expandImproved(uint64_t *cv, size_t g, uint32_t t_exp, const uint64_t *W_left, const uint64_t *W_right,
                         uint32_t t_exp_right, size_t max_bits_right, size_t stopround);            /* src/spiral.cpp:1664-1743 */
int sb200_scalToMat(uint64_t *out_reg, const uint64_t *cv, const uint64_t *W, uint32_t t_conv);     /* src/spiral.cpp:1850-1885 */
int sb200_regevToGSW(uint64_t *out, const uint64_t *cv_v, uint32_t t_conv, uint32_t t, const uint64_t *W,
                     const uint64_t *V);                                                            /* src/spiral.cpp:1985-2025 */

typedef struct sb200_server sb200_server;      /* the resident server, tier 3 below */
/* resident forms of the leaves of process_query_fast (src/spiral.cpp:1584-1629) for a drop-in host that must keep the reference's
 * call sequence: the intermediates stay in HBM between the leaves and the harness's own host pointers are only
 * KEYS naming what is resident.  A leaf whose input key does not match returns SB200_ERR_STATE - call the stateless form then.
 *   reorientCiphertexts      uploads the 2^nu1 converted ciphertexts, leaves the reoriented query resident under out_key
 *   multiplyQueryByDatabase  scans the resident database with the resident query; result resident under out_key
 *   nttInvAndCrtLift...      lifts the resident scan output; ciphertexts resident under cts_key (= furtherDimsLocals.cts)
 *   reorient_Q               one GSW ciphertext (n1 x m2 ref-NTT, the INPUT of the reference's reorient_Q) resident under out_key
 *   foldOneFurtherDimension  one fold round (the reference's two-product form, arbitrary q_neg); the last round (num_per == 1)
 *                            downloads the surviving ciphertext into cts_host, which check_final / modswitch read */
int sb200_resident_reorientCiphertexts(sb200_server *srv, const void *out_key, const uint64_t *inp_ref_ntt_host, size_t dim0);
int sb200_resident_multiplyQueryByDatabase(sb200_server *srv, const void *out_key, const void *reoriented_key);
int sb200_resident_nttInvAndCrtLiftCiphertexts(sb200_server *srv, const void *cts_key, const void *scratch_key);
int sb200_resident_reorient_Q(sb200_server *srv, const void *out_key, const uint64_t *inp_ref_ntt_host);
int sb200_resident_foldOneFurtherDimension(sb200_server *srv, size_t num_per, const void *q_key, const void *q_neg_key, uint64_t *cts_host);

/* SpiralPack / SpiralStreamPack leaves (src/testing.cpp) */
int sb200_convertDb(uint64_t *db_buf, const uint64_t *db_ref_ntt, size_t count, size_t dim0, size_t num_per);          /* :316-340 */
int sb200_reorientCiphertextsDim1(uint64_t *out, const uint64_t *v_firstdim, size_t count, size_t dim0, size_t idx_factor); /* :342-362 */
int sb200_fastMultiplyQueryByDatabaseDim1(uint64_t *out_ref_ntt, const uint64_t *db, const uint64_t *v_firstdim,
                                          size_t dim0, size_t num_per);                                                   /* :364-593 */
/* `count` (<= 16) reoriented queries against one plane in ONE tensor-core pass; outputs as fastMultiplyQueryByDatabaseDim1's */
int sb200_fastMultiplyQueryByDatabaseDim1_batched(uint64_t *const *out_ref_ntt, const uint64_t *db_buf, const uint64_t *const *v_firstdim_reoriented,
                                                  int count, size_t dim0, size_t num_per);
/* v_cts: count cts (2x1 raw), result left in v_cts[0]; v_folding / v_folding_neg: log2(count) x (2 x 2*ell) ref-NTT */
int sb200_foldCiphertextsDim1(uint64_t *v_cts, size_t count, const uint64_t *v_folding, const uint64_t *v_folding_neg, uint32_t ell); /* :596-624 */
int sb200_regevToSimpleGsw(uint64_t *v_gsw, const uint64_t *v_inp, size_t count_inp, const uint64_t *V, uint32_t t_conv,
                           uint32_t ell, uint32_t further_dims, size_t idx_factor, size_t idx_offset);                  /* :108-140 */
int sb200_pack(uint64_t *result_ref_ntt, uint32_t out_n, uint32_t t_conv, const uint64_t *v_ct_raw, const uint64_t *v_W); /* :198-241 */

/* ---- tier 3: resident server ----------------------------------------------------------- */
typedef struct sb200_server sb200_server;
/* shard `rank` of `world` owns second-dimension indices ii = rank (mod world); world = 1 -> whole database */
int sb200_server_create(sb200_server **out, const sb200_params *prm, int device, int rank, int world);
void sb200_server_destroy(sb200_server *srv);
/* another server over the SAME resident database (own workspaces, keys, graphs): one per concurrent client */
int sb200_server_create_view(sb200_server **out, sb200_server *parent);
/* database from plaintext items (u16 coefficients, this shard's items only, order j-major: item = j*local_num_per + ii_local) */
int sb200_server_load_db_items(sb200_server *srv, const uint16_t *pts_host, size_t item_begin, size_t item_count);
/* database handed over in the reference's layout (the WHOLE B of load_db); the shard's rows are extracted */
int sb200_server_load_db_reference(sb200_server *srv, const uint64_t *B_host);
/* the reference's --random-data ("implicit database") mode, src/spiral.cpp:1032-1081, 1274-1282: B_slices holds `working_set`
 * z-slices in load_db's layout (dummyWorkingSet = min(2^25 / total_n, 2048), a power of two) and the scan reads slice
 * z mod working_set (:647, the AVX-512 statement) - half the memory, the same 8 B x 2048 x 4 x 2^(nu1+nu2) algorithmic bytes */
int sb200_server_load_db_implicit(sb200_server *srv, const uint64_t *B_slices_host, size_t working_set);
/* the same database generated on the device: every record the constant `value` (< p_db / 2) in coefficient 0 of its four polynomials */
int sb200_server_load_db_implicit_constant(sb200_server *srv, uint64_t value, size_t working_set);
size_t sb200_server_db_slices(const sb200_server *srv);     /* 2048 for an explicit database */
uint64_t *sb200_server_db_ptr(sb200_server *srv);          /* device pointer of the scan-layout shard */
/* public parameters, ref-NTT host buffers: W_exp_left g x (2 x t_exp), W_exp_right (stopround+1 or g) x (2 x t_exp_right),
 * W_conv 3 x 2*t_conv, V_conv 3 x 2*t_conv  (runConversionImproved, src/spiral.cpp:2093-2300) */
int sb200_server_set_public_params(sb200_server *srv, const uint64_t *W_exp_left, const uint64_t *W_exp_right,
                                   const uint64_t *W_conv, const uint64_t *V_conv);
/* polynomial counts (2*2048 words each) of those four matrices, in that order: what the server will READ.  Public parameters
 * arrive from an untrusted client - compare the sizes received with these before calling sb200_server_set_public_params. */
int sb200_server_public_param_polys(const sb200_server *srv, size_t *out4);
/* one query end to end: H2D of the packed query ciphertext (2x1 ref-NTT, 64 KiB), all server stages, D2H of the
 * modulus-switched response (3x2 raw, 96 KiB).  world == 1 only. */
int sb200_server_answer(sb200_server *srv, const uint64_t *query_cv_host, uint64_t *total_resp_host, void *stream);
/* all server stages of the query last uploaded (sb200_server_upload_query[_wire]) in one call, response left in total_resp_dev
 * (NULL: the server's own buffer); marks: NULL or four cudaEvent_t recorded before the expansion, before and after the
 * first-dimension scan and at the end.  Sharded servers (world > 1, peers connected with sb200_server_xchg_connect): every rank
 * calls it; the exchange over NVLink peer memory and rank 0's tail folds are part of the call, the response lands on rank 0. */
int sb200_server_process(sb200_server *srv, uint64_t *total_resp_dev, void *stream, void *const *marks);
/* captures and instantiates every CUDA graph sb200_server_process will replay, without running anything (the first query then
 * costs what the others do; needs the public parameters and, on a sharded server, connected peers) */
int sb200_server_prepare(sb200_server *srv, uint64_t *total_resp_dev, void *stream);
/* same, response in the wire format (sb200_dev_pack_response): 20 KiB instead of 96 KiB at cfg1 */
int sb200_server_answer_packed(sb200_server *srv, const uint64_t *query_cv_host, uint64_t *packed_resp_host, void *stream);
size_t sb200_server_packed_response_bytes(const sb200_server *srv);
/* staged variants (device-resident between stages; used by bench.py and the multi-GPU path) */
int sb200_server_upload_query(sb200_server *srv, const uint64_t *query_cv_host, void *stream);
int sb200_server_expand_and_convert(sb200_server *srv, void *stream);      /* expansion + ScalToMat + RegevToGSW (+negation) */
int sb200_server_first_dim(sb200_server *srv, void *stream);               /* scan + INTT + CRT lift */
int sb200_server_scan(sb200_server *srv, void *stream);                    /* multiplyQueryByDatabase only (src/spiral.cpp:628) */
/* batched first dimension: `count` (2 or 4) servers sharing one database answered in ONE database pass */
int sb200_server_scan_batched(sb200_server *const *servers, int count, void *stream);
/* tensor-core variant: build the limb-tile database copy once (database owner; capacity <= 16 queries per pass) ... */
int sb200_server_enable_tc(sb200_server *srv, int capacity);
/* ... optionally keep ONLY that copy: the scan-layout copy is freed (one database's worth of HBM instead of two) and a single query's
   first dimension - sb200_server_scan / _process / _answer* - becomes a one-query pass of the tensor-core kernel.  Servers sharing
   the database must then scan on one stream; loading the database again drops the tensor-core state; snapshots need the scan layout */
int sb200_server_tc_only(sb200_server *srv);
/* ... then answer the first dimension of up to `capacity` servers sharing that database with one tcgen05 pass */
int sb200_server_scan_batched_tc(sb200_server *const *servers, int count, void *stream);
int sb200_server_lift(sb200_server *srv, void *stream);                    /* nttInvAndCrtLiftCiphertexts only (src/spiral.cpp:437) */
/* interposed multiplyQueryByDatabase: host reoriented query in, ref-NTT host ciphertexts out, database stays resident */
int sb200_server_scan_host(sb200_server *srv, const uint64_t *reoriented_host, uint64_t *out_ref_ntt_host);
int sb200_server_copy_partial(sb200_server *srv, uint64_t *dst_dev, void *stream);   /* D2D copy of the shard's surviving ct */
int sb200_server_load_db_random(sb200_server *srv, uint64_t seed);         /* synthetic uniform database (benchmarks) */
int sb200_server_fold_local(sb200_server *srv, void *stream);              /* local fold rounds; leaves 1 ct per shard */
uint64_t *sb200_server_partial_ct(sb200_server *srv);                      /* device ptr: this shard's surviving ct (3x2 raw) */
/* rank 0: `gathered` = world cts (device, order = rank); runs the last log2(world) folds + modulus switch */
int sb200_server_fold_tail(sb200_server *srv, uint64_t *gathered_dev, uint64_t *total_resp_dev, void *stream);
/* the exchange step over NVLink peer memory (no NCCL, no host round trip): every rank stores its surviving ciphertext
 * straight into rank 0's HBM and publishes a flag; rank 0 waits, runs the tail folds + modulus switch.
 * Setup once: each rank exports a handle, all ranks connect with the world handles in rank order. */
size_t sb200_server_xchg_handle_bytes(void);
int sb200_server_xchg_export(sb200_server *srv, void *handle_out);
int sb200_server_xchg_connect(sb200_server *srv, const void *all_handles);                 /* one process per GPU (cudaIpc) */
int sb200_server_xchg_connect_local(sb200_server *srv, sb200_server *const *all_servers);  /* shards inside one process */
int sb200_server_exchange_and_tail(sb200_server *srv, uint64_t *total_resp_dev, void *stream);
int sb200_server_xchg_error(sb200_server *srv, void *stream);      /* 0 ok; 1/2/3 = a bounded spin timed out (4 s) */
/* 1 when the last expansion was sharded: connected peers and 2^nu1 / world a multiple of 8 - each rank then expands and converts
 * only the first-dimension ciphertexts j = rank (mod world) and stores them into every rank's query buffer over peer memory */
int sb200_server_expansion_sharded(const sb200_server *srv);
/* after a time-out the shards are out of step and every later exchange would time out too: with no query in flight EVERY rank
 * calls this (epochs, flags, acks, error word start over).  sb200_server_download on a sharded server returns SB200_ERR_STATE
 * instead of a garbage response while the error word is set. */
int sb200_server_xchg_reset(sb200_server *srv);
int sb200_server_download(sb200_server *srv, uint64_t *dst_host, const uint64_t *src_dev, size_t words, void *stream);
/* debug taps (device pointers): raw cts after the first dimension; final ct before modulus switch */
uint64_t *sb200_server_first_dim_cts(sb200_server *srv);
size_t sb200_server_query_bytes(const sb200_server *srv);
size_t sb200_server_response_bytes(const sb200_server *srv);


/* ---- tier 3, Pack variants (testHighRate, src/testing.cpp:777-1155; server statements :1007-1081) ---- */
typedef struct sb200_pack_server sb200_pack_server;
int sb200_pack_server_create(sb200_pack_server **out, const sb200_params *prm, int device);
/* shard `rank` of `world`: second-dimension indices ii = rank (mod world) of every plane */
int sb200_pack_server_create_sharded(sb200_pack_server **out, const sb200_params *prm, int device, int rank, int world);
/* plane sharding ("Pack alternative"; src/testing.cpp:1045-1061 runs the planes as independent trials): rank r holds
 * the WHOLE planes p = r (mod world), scans and folds them alone; the one exchange is a folded 32 KiB ciphertext per plane to
 * rank 0, which packs.  Plane arguments of the load entries stay GLOBAL indices; a plane of another rank is SB200_ERR_ARG. */
int sb200_pack_server_create_plane_sharded(sb200_pack_server **out, const sb200_params *prm, int device, int rank, int world);
int sb200_pack_server_owns_plane(const sb200_pack_server *srv, size_t plane);
size_t sb200_pack_server_local_planes(const sb200_pack_server *srv);
void sb200_pack_server_destroy(sb200_pack_server *srv);
/* one of the out_n^2 database planes: this shard's 2^nu1 * (2^nu2 / world) items of one polynomial each (u16 coefficients
 * < p_db), j-major: item = j * local_num_per + ii_local */
int sb200_pack_server_load_plane_items(sb200_pack_server *srv, size_t plane, const uint16_t *pts_host);
/* one item of a loaded plane replaced in place: first-dimension index j, second-dimension index ii_local inside this shard */
int sb200_pack_server_set_plane_item(sb200_pack_server *srv, size_t plane, size_t j, size_t ii_local, const uint16_t *poly_host);
/* the WHOLE plane in the reference's convertDb layout (src/testing.cpp:316-340); the shard's rows are extracted */
int sb200_pack_server_load_plane_reference(sb200_pack_server *srv, size_t plane, const uint64_t *db_buf_host);
int sb200_pack_server_load_random(sb200_pack_server *srv, uint64_t seed);     /* synthetic plaintexts generated on the device */
/* W_exp_left g x (2 x t_exp), W_exp_right (stopround+1) x (2 x t_exp_right), V 2 x 2*t_conv (all three may be NULL for
 * direct-upload clients), v_W out_n x ((out_n+1) x t_conv); ref-NTT host buffers */
int sb200_pack_server_set_public_params(sb200_pack_server *srv, const uint64_t *W_exp_left, const uint64_t *W_exp_right,
                                        const uint64_t *V, const uint64_t *v_W);
/* total_resp: (out_n+1) x out_n raw; result_cts (optional): out_n^2 folded cts (2x1 raw) before packing.  world == 1 only. */
int sb200_pack_server_answer(sb200_pack_server *srv, const uint64_t *query_cv_host, uint64_t *total_resp_host,
                             uint64_t *result_cts_host, void *stream);
int sb200_pack_server_answer_direct(sb200_pack_server *srv, const uint64_t *v_firstdim_host, const uint64_t *v_folding_host,
                                    uint64_t *total_resp_host, uint64_t *result_cts_host, void *stream);
/* staged variants (device-resident between stages; bench.py and the multi-GPU path) */
int sb200_pack_server_upload_query(sb200_pack_server *srv, const uint64_t *query_cv_host, void *stream);
int sb200_pack_server_expand_and_convert(sb200_pack_server *srv, void *stream);   /* coefficientExpansion + reorientCiphertextsDim1 + regevToSimpleGsw */
int sb200_pack_server_upload_direct(sb200_pack_server *srv, const uint64_t *v_firstdim_host, const uint64_t *v_folding_host, void *stream);
int sb200_pack_server_scan(sb200_pack_server *srv, void *stream);                 /* fastMultiplyQueryByDatabaseDim1, all planes (src/testing.cpp:364) */
/* several clients over one resident set of planes, and their first dimensions in ONE tensor-core pass (as sb200_server_*_tc) */
int sb200_pack_server_create_view(sb200_pack_server **out, sb200_pack_server *parent);
int sb200_pack_server_enable_tc(sb200_pack_server *srv, int capacity);     /* needs dim0 and num_per (per shard) multiples of 128 */
int sb200_pack_server_tc_only(sb200_pack_server *srv);                       /* as sb200_server_tc_only: SpiralPack cfg3 = 64 GiB resident, not 128 */
int sb200_pack_server_scan_batched_tc(sb200_pack_server *const *servers, int count, void *stream);
/* interposed fastMultiplyQueryByDatabaseDim1 against ONE resident plane: host reoriented query in, ref-NTT host cts out */
int sb200_pack_server_scan_plane_host(sb200_pack_server *srv, size_t plane, const uint64_t *v_firstdim_host, uint64_t *out_ref_ntt_host);
int sb200_pack_server_fold_local(sb200_pack_server *srv, void *stream);           /* from_ntt + local fold rounds: out_n^2 surviving cts */
uint64_t *sb200_pack_server_partial_cts(sb200_pack_server *srv);                  /* device ptr: [plane] 2x1 raw cts of this shard */
size_t sb200_pack_server_partial_words(const sb200_pack_server *srv);
int sb200_pack_server_copy_partial(sb200_pack_server *srv, uint64_t *dst_dev, void *stream);
/* rank 0: gathered = [world][plane] cts (device, rank order); last log2(world) folds + pack + modulus switch */
int sb200_pack_server_fold_tail(sb200_pack_server *srv, const uint64_t *gathered_dev, uint64_t *total_resp_dev, void *stream);
uint64_t *sb200_pack_server_result_cts(sb200_pack_server *srv);                   /* device ptr: folded per-plane cts after fold_tail */
uint64_t *sb200_pack_server_response_ptr(sb200_pack_server *srv);                 /* device ptr: the server's own response buffer */
int sb200_pack_server_download(sb200_pack_server *srv, uint64_t *dst_host, const uint64_t *src_dev, size_t words, void *stream);
/* the exchange step over NVLink peer memory, as sb200_server_xchg_*: every rank stores its surviving ciphertexts into rank 0's
 * HBM and raises a flag; rank 0 waits, runs the tail (last log2(world) folds under second-dimension sharding; nothing under
 * plane sharding), packs and switches the modulus.  Handles carry the exchange buffer and the query buffer of a rank. */
size_t sb200_pack_server_xchg_handle_bytes(void);
int sb200_pack_server_xchg_export(sb200_pack_server *srv, void *handle_out);
int sb200_pack_server_xchg_connect(sb200_pack_server *srv, const void *all_handles);                     /* one process per GPU (cudaIpc) */
int sb200_pack_server_xchg_connect_local(sb200_pack_server *srv, sb200_pack_server *const *all_servers);  /* shards inside one process */
int sb200_pack_server_exchange_and_tail(sb200_pack_server *srv, uint64_t *total_resp_dev, void *stream);
int sb200_pack_server_xchg_error(sb200_pack_server *srv, void *stream);    /* 0 ok; 1/2/3 = a bounded spin timed out (4 s) */
/* sharded direct upload: this rank's 1/world of the first-dimension ciphertexts (j in [rank, rank + 1) * 2^nu1 / world) + all GSW
 * ciphertexts; the reorientation kernel stores the slice into EVERY rank's query buffer over peer memory (fused all-gather) */
int sb200_pack_server_upload_direct_split(sb200_pack_server *srv, const uint64_t *v_firstdim_slice_host, const uint64_t *v_folding_host, void *stream);
/* all server stages of the query last uploaded in ONE call (every rank of a sharded server calls it; response on rank 0, in
 * total_resp_dev or the server's own buffer when NULL); marks as sb200_server_process */
int sb200_pack_server_process(sb200_pack_server *srv, uint64_t *total_resp_dev, void *stream, void *const *marks);
/* as sb200_server_prepare: build the graphs _process replays without running anything */
int sb200_pack_server_prepare(sb200_pack_server *srv, uint64_t *total_resp_dev, void *stream);
/* 1 when the packed-query expansion (coefficientExpansion, src/testing.cpp:40-105) is shared out: a column-sharded server with
 * connected peers expands only the first-dimension ciphertexts j = rank (mod world) and stores them, reoriented, into every rank's
 * query buffer over NVLink (the GSW-bit chain stays whole on every rank) */
int sb200_pack_server_expansion_sharded(const sb200_pack_server *srv);
size_t sb200_pack_server_db_bytes(const sb200_pack_server *srv);                  /* this shard */
size_t sb200_pack_server_response_words(const sb200_pack_server *srv);


/* ---- wire and on-disk formats ------------------------------------------------------------------
 * The reference keeps client and server in one process: it only ACCOUNTS for the sizes of what would travel
 * (print_summary, src/spiral.cpp:219-234) and leaves the database file I/O of load_db as `// TODO`
 * (src/spiral.cpp:1095-1162).  These entries are what a client/server split binds instead; all multi-byte fields are
 * little endian and bit-packed values follow write_arbitrary_bits (src/core.cpp:32-52): value i occupies bits
 * [i*bits, (i+1)*bits) of the stream.
 *
 * Query (the 2x1 Regev ciphertext of getRegevSample, src/client.cpp:141-157; row 0 = -a is uniformly random):
 *   bytes 0-3 "SB2Q", bytes 4-5 kind, bytes 6-7 zero, then
 *   kind 1 SEEDED : 32-byte seed | row 1 raw at 56 bits per coefficient                      (8 + 32 + 14336 bytes)
 *                   row 0 in NTT form, slot (prime n, z): ChaCha20 block (RFC 8439; key = seed, counter = n*2048 + z,
 *                   nonce = "SB2Q",0,0), the first of its 16 words whose low 28 bits are below the prime
 *   kind 2 FULL   : row 0 | row 1, both raw at 56 bits per coefficient                        (8 + 2*14336 bytes)
 * Response: sb200_dev_pack_response above.
 * Records: the database as a flat bit stream, log2(p_db) bits per plaintext coefficient, item-major, polynomial-major
 *   inside an item (Spiral: the 2x2 polynomials (m*2+c) of item j*2^nu2 + ii; Pack: the out_n^2 plane polynomials).
 * Snapshot: 64-byte header (magic "SB2D", version, kind, nu1, nu2, rank, world, out_n, p_db, words, sum of words) +
 *   the preprocessed scan-layout shard exactly as it sits in HBM; loading verifies the parameters and the sum. */
#define SB200_WIRE_QUERY_SEEDED 1u
#define SB200_WIRE_QUERY_FULL 2u
size_t sb200_wire_query_bytes(uint32_t kind);                       /* 0 for an unknown kind */
/* device copy of a wire query (header included) -> cv[0], the 2x1 dev-NTT ciphertext the expansion reads */
int sb200_dev_query_from_wire(uint32_t *cv_dev, const uint8_t *wire_dev, uint32_t kind, void *stream);
/* as sb200_server_upload_query, the query in wire form; a malformed buffer is SB200_ERR_ARG and nothing is uploaded */
int sb200_server_upload_query_wire(sb200_server *srv, const uint8_t *wire_host, size_t bytes, void *stream);
/* one query entirely in wire form: wire query in, sb200_server_packed_response_bytes() of packed response out */
int sb200_server_answer_wire(sb200_server *srv, const uint8_t *wire_host, size_t bytes, uint64_t *packed_resp_host, void *stream);
int sb200_pack_server_upload_query_wire(sb200_pack_server *srv, const uint8_t *wire_host, size_t bytes, void *stream);
/* as sb200_pack_server_answer, the query in wire form */
int sb200_pack_server_answer_wire(sb200_pack_server *srv, const uint8_t *wire_host, size_t bytes, uint64_t *total_resp_host,
                                  uint64_t *result_cts_host, void *stream);
uint64_t *sb200_pack_server_db_ptr(sb200_pack_server *srv);        /* device pointer of the scan-layout planes of this shard */
/* load_db's `has_data` branch (src/spiral.cpp:1108-1111): the WHOLE database as records, in memory or in a file; a sharded
 * server reads only its own items (ii = rank mod world) */
size_t sb200_server_record_stream_bytes(const sb200_server *srv);
int sb200_server_load_db_records(sb200_server *srv, const uint8_t *records_host, size_t bytes);
int sb200_server_load_db_records_file(sb200_server *srv, const char *path);
/* load_db's `has_file && !load` / `has_file && load` branches (src/spiral.cpp:1095-1097, 1159-1161): the preprocessed shard */
int sb200_server_save_db(sb200_server *srv, const char *path);
int sb200_server_load_db_snapshot(sb200_server *srv, const char *path);
size_t sb200_pack_server_record_stream_bytes(const sb200_pack_server *srv);
int sb200_pack_server_load_db_records(sb200_pack_server *srv, const uint8_t *records_host, size_t bytes);
int sb200_pack_server_load_db_records_file(sb200_pack_server *srv, const char *path);
int sb200_pack_server_save_db(sb200_pack_server *srv, const char *path);
int sb200_pack_server_load_db_snapshot(sb200_pack_server *srv, const char *path);


/* ---- the client on the GPU ------------------------------------------------------------------------
 * Key generation (src/client.cpp:23-47), public parameters (getPublicEncryptions src/client.cpp:271-290; W and V,
 * src/spiral.cpp:2207-2290), query encoding + encryption (src/spiral.cpp:2098-2157, encryptSimpleRegev src/client.cpp:170-186)
 * and decoding (check_final, src/spiral.cpp:1428-1476 - the reference's only use of Intel HEXL).  Spiral / SpiralStream
 * parameter sets (matrix-Regev responses).  The reference draws from an unseeded std::random_device; here all randomness
 * derives from the 32-byte seed given at creation: every random polynomial is a ChaCha20 (RFC 8439) stream with nonce
 * {"SB2C", object id, stream}, uniform polynomials drawn directly in NTT form, Gaussian ones (width 6.4, support [-64, 64],
 * src/core.cpp:182-207) by inverse CDF on 53-bit uniforms.  oracle/client_sim.c (so_client_new_chacha) states the same client
 * in plain C; the two agree bit for bit. */
typedef struct sb200_client sb200_client;
int sb200_client_create(sb200_client **out, const sb200_params *prm, int device, const uint8_t *seed32);
void sb200_client_destroy(sb200_client *c);
/* polynomial counts of the four public-parameter matrices (W_exp_left, W_exp_right, W_conv, V_conv), 2*2048 words each */
int sb200_client_public_param_polys(const sb200_client *c, size_t *out4);
/* ref-NTT host buffers, exactly what sb200_server_set_public_params takes */
int sb200_client_public_params(sb200_client *c, uint64_t *W_exp_left, uint64_t *W_exp_right, uint64_t *W_conv, uint64_t *V_conv);
/* SEEDED wire query for record idx_target; wire_out: sb200_wire_query_bytes(1) bytes.
 * FRESHNESS (privacy-critical): query_id (< 2^24) names the noise stream AND, with wire_seed32 == NULL (the recommended call),
 * the seed of row 0 = -a, which is then derived from the client key and query_id (sb200_client_wire_seed) - so ONE monotonic
 * counter per client key is all a caller keeps, and it must never repeat.  An explicit wire_seed32 is for tests / callers with
 * their own randomness; it too must never repeat under one client key: two queries sharing `a` reveal sigma - sigma' (the
 * difference of the queried indices) to the server. */
int sb200_client_query_wire(sb200_client *c, size_t idx_target, uint32_t query_id, const uint8_t *wire_seed32, uint8_t *wire_out);
/* the row-0 seed sb200_client_query_wire derives for query_id: ChaCha20 block (key = client seed, counter 0, nonce {"SB2C",
 * 6 << 24 | query_id, "wsee"}), first 32 bytes */
int sb200_client_wire_seed(const sb200_client *c, uint32_t query_id, uint8_t *seed32_out);
/* total_resp: 3x2 raw response (sb200_server_answer, or sb200_unpack_response of the packed one) -> 2x2 plaintext polynomials */
int sb200_client_decode(sb200_client *c, const uint64_t *total_resp_host, uint64_t *pt_out_host);
/* ---- the SpiralPack / SpiralStreamPack client (testHighRate's client statements, src/testing.cpp:904-1005, 1086-1122): keys with
 * out_n rows of S', packing keys v_W (:905-912), expansion keys + V (:913-931), the packed query (:987-1005) or the direct upload
 * (:962-985), out_n x out_n decoding (:1086-1118).  Same counter-based randomness; oracle/client_sim.c (so_pack_client_new_chacha)
 * states it in plain C.  Destroy / wire seed / secret tap: the sb200_client_* entries. */
int sb200_pack_client_create(sb200_client **out, const sb200_params *prm, int device, const uint8_t *seed32);
/* polynomial counts of W_exp_left, W_exp_right, V, v_W - the arguments of sb200_pack_server_set_public_params */
int sb200_pack_client_public_param_polys(const sb200_client *c, size_t *out4);
/* W_exp_left, W_exp_right and V may all be NULL (direct-upload client) */
int sb200_pack_client_public_params(sb200_client *c, uint64_t *W_exp_left, uint64_t *W_exp_right, uint64_t *V, uint64_t *v_W);
/* SEEDED wire query (as sb200_client_query_wire; needs 2^nu1 + t_GSW*nu2 <= 2048) */
int sb200_pack_client_query_wire(sb200_client *c, size_t idx_target, uint32_t query_id, const uint8_t *wire_seed32, uint8_t *wire_out);
/* direct upload: v_firstdim = 2^nu1 ciphertexts (2x1 ref-NTT), v_folding = nu2 x (2 x 2*t_GSW) ref-NTT; query_id < 2^8 selects a
 * fresh block of 2^16 randomness objects (never reuse one under the same client seed) */
int sb200_pack_client_query_direct(sb200_client *c, size_t idx_target, uint32_t query_id, uint64_t *v_firstdim, uint64_t *v_folding);
/* total_resp (out_n+1) x out_n raw -> out_n x out_n plaintext polynomials (entry i*out_n + j = the record's polynomial of that plane) */
int sb200_pack_client_decode(sb200_client *c, const uint64_t *total_resp_host, uint64_t *pt_out_host);
/* test taps: the secret key (raw, 1 + 2 polynomials; 1 + out_n for a Pack client) and the sampler's 128 integer thresholds */
int sb200_client_secret(sb200_client *c, uint64_t *sr_raw_host, uint64_t *Sp_raw_host);
int sb200_client_gaussian_thresholds(uint64_t *out128);

#ifdef __cplusplus
}
#endif
#endif /* SPIRAL_B200_H */
