/* client_sim.h - minimal Spiral client for end-to-end tests.  TEST INFRASTRUCTURE ONLY (see client_sim.c). */
#pragma once
#include "spiral_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct so_client so_client;
so_client *so_client_new(const so_params *prm, uint64_t seed, int nonoise);
/* the same client with counter-based randomness (ChaCha20 streams named by object ids, see client_sim.c): the statement the
 * CUDA client (include/spiral_b200.h, sb200_client_*) is compared with bit for bit */
so_client *so_client_new_chacha(const so_params *prm, const uint8_t seed[32]);
void so_client_gaussian_thresholds(const so_client *c, uint64_t *out128);
void so_client_secret(const so_client *c, uint64_t *sr_raw, uint64_t *Sp_raw);
void so_client_secret_n(const so_client *c, uint64_t *sr_raw, uint64_t *Sp_raw, size_t sp_rows);
void so_client_chacha_query_wire(so_client *c, size_t idx_target, uint32_t query_id, const uint8_t wire_seed[32], uint8_t *wire);
void so_client_free(so_client *c);
size_t so_client_w_exp_right_count(const so_params *prm);
/* W_exp_left: g x (2 x t_exp); W_exp_right: count x (2 x t_exp_right); W_conv, V_conv: 3 x 2*t_conv (all ref-NTT) */
void so_client_spiral_pub_params(so_client *c, uint64_t *W_exp_left, uint64_t *W_exp_right, uint64_t *W_conv, uint64_t *V_conv);
void so_client_spiral_query(so_client *c, size_t idx_target, uint64_t *query_cv);
/* kind: SO_WIRE_QUERY_SEEDED / SO_WIRE_QUERY_FULL (wire_format.h); wire: so_wire_query_bytes(kind) bytes */
void so_client_spiral_query_wire(so_client *c, size_t idx_target, uint32_t kind, uint8_t *wire);
void so_client_spiral_decode(so_client *c, const uint64_t *total_resp, uint64_t *out_pt);
/* ---- SpiralPack / SpiralStreamPack client (testHighRate's client statements, src/testing.cpp:777-1155) ---- */
so_client *so_pack_client_new(const so_params *prm, uint64_t seed);
so_client *so_pack_client_new_chacha(const so_params *prm, const uint8_t seed[32]);     /* counter-based randomness (see so_client_new_chacha) */
void so_pack_client_chacha_query_wire(so_client *c, size_t idx_target, uint32_t query_id, const uint8_t wire_seed[32], uint8_t *wire);
/* W_exp_left: g x (2 x t_exp), W_exp_right: (stopround+1) x (2 x t_exp_right), V: 2 x 2*t_conv (all three NULL for a direct-upload
 * client); v_W: out_n x ((out_n+1) x t_conv); all ref-NTT */
void so_pack_client_pub_params(so_client *c, uint64_t *W_exp_left, uint64_t *W_exp_right, uint64_t *V, uint64_t *v_W);
void so_pack_client_query(so_client *c, size_t idx_target, uint64_t *query_cv);
void so_pack_client_query_direct(so_client *c, size_t idx_target, uint64_t *v_firstdim, uint64_t *v_folding);
void so_pack_client_decode(so_client *c, const uint64_t *total_resp, uint64_t *out_pt);
#ifdef __cplusplus
}
#endif
