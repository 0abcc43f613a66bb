/* golden_cases.h - shared parity-case table (TEST INFRASTRUCTURE ONLY, see golden_cases.c). */
#pragma once
#include "spiral_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif

enum {
    SO_CASE_NTT_FWD = 0, SO_CASE_NTT_INV, SO_CASE_TO_NTT, SO_CASE_FROM_NTT, SO_CASE_MULTIPLY,
    SO_CASE_AUTOMORPH, SO_CASE_GADGET_INVERT, SO_CASE_RESCALE, SO_CASE_REORIENT, SO_CASE_FIRST_DIM,
    SO_CASE_NTT_INV_CRT, SO_CASE_SPLIT_AND_CRT, SO_CASE_FOLD_ONE, SO_CASE_EXPAND_FULL,
    SO_CASE_EXPAND_STOP, SO_CASE_SCAL_TO_MAT, SO_CASE_REGEV_TO_GSW, SO_CASE_LOAD_DB,
    SO_CASE_CONVERT_DB, SO_CASE_REORIENT_DIM1, SO_CASE_FIRST_DIM_PACK, SO_CASE_FOLD_DIM1,
    SO_CASE_REGEV_TO_SGSW, SO_CASE_PACK, SO_CASE_TO_NTT_NR, SO_CASE_MODSWITCH, SO_CASE_COUNT
};
enum { SO_KIND_RAW = 0, SO_KIND_NTT = 1, SO_KIND_PACKED = 2 };
#define SO_CASE_MAX_IN 4

typedef struct so_case_io {
    uint64_t *in[SO_CASE_MAX_IN];
    size_t in_words[SO_CASE_MAX_IN];
    uint64_t *out;
    size_t out_words;
    int out_kind;
} so_case_io;

typedef struct so_case_shape_t {
    size_t npolys, dim0, num_per, g, stopround, max_bits_right, cur_dim;
} so_case_shape_t;

int so_case_count(void);
const char *so_case_name(int id);
void so_case_shape(int id, const so_params *p, so_case_shape_t *s);
void so_case_make_inputs(int id, const so_params *p, uint64_t seed, so_case_io *io);
void so_case_run_oracle(int id, const so_params *p, so_case_io *io);
uint64_t so_case_digest(const so_case_io *io);
uint64_t so_digest_kind(const uint64_t *w, size_t words, int kind);
void so_case_free(so_case_io *io);

#ifdef __cplusplus
}
#endif
