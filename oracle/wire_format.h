/*
 * wire_format.h - CPU statement of the wire / on-disk formats of SURVEY section 8f #2.
 *
 * TEST INFRASTRUCTURE ONLY (part of liboracle.so, see spiral_oracle.h).
 *
 * The reference only ACCOUNTS for these formats (query / response sizes printed by print_summary,
 * src/spiral.cpp:219-234) and leaves the database file I/O as `// TODO` (src/spiral.cpp:1095-1162);
 * it never serialises anything.  The formats are therefore defined by this project
 * (include/spiral_b200.h, "wire formats"); this file is their independent plain-C statement, built on
 * the reference's own bit I/O (read/write_arbitrary_bits, src/core.cpp:20-52) and pinned by the
 * RFC 8439 ChaCha20 block-function test vector.  What IS pinned to the reference: the expanded
 * query is an ordinary 2x1 Regev ciphertext whose answer must decode to the planted record, and a
 * database loaded from records must equal load_db's (src/spiral.cpp:1028-1172) on the same plaintexts.
 */
#pragma once
#include "spiral_oracle.h"
#ifdef __cplusplus
extern "C" {
#endif

#define SO_WIRE_QUERY_MAGIC 0x51324253u      /* "SB2Q" little endian */
#define SO_WIRE_QUERY_SEEDED 1u              /* header | 32-byte seed | row 1 at 56 bits per coefficient  */
#define SO_WIRE_QUERY_FULL 2u                /* header | row 0 | row 1, both at 56 bits per coefficient   */
#define SO_WIRE_HEADER_BYTES 8u
#define SO_WIRE_SEED_BYTES 32u
#define SO_WIRE_ROW_BYTES (SO_N * SO_LOGQ / 8u)        /* 14336 = b_per_elem of src/spiral.cpp:219 */

/* RFC 8439 section 2.3 block function: key 8 words, counter, nonce 3 words -> 16 words */
void so_chacha20_block(const uint32_t key[8], uint32_t counter, const uint32_t nonce[3], uint32_t out[16]);
/* row 0 of a seeded query in NTT form ([2][N] u64): slot (n, z) = first of the 16 words of ChaCha20 block
 * (key = seed, counter = n*N + z, nonce = {"SB2Q", 0, 0}) whose low 28 bits are below the prime; if all 16
 * are rejected (probability < 2^-60) the last word's low 28 bits reduced modulo the prime. */
void so_wire_seeded_row0(const uint8_t seed[32], uint64_t *row0_ntt);
size_t so_wire_query_bytes(uint32_t kind);
/* wire -> 2x1 ref-NTT query ciphertext; returns 0, or -1 on a malformed buffer */
int so_wire_query_expand(const uint8_t *wire, size_t bytes, uint64_t *query_cv);
/* 2x1 ref-NTT query ciphertext -> FULL wire form (both rows from_ntt'd and packed) */
void so_wire_query_pack_full(const uint64_t *query_cv, uint8_t *wire);
/* header + seed + row 1 (NTT form in, from_ntt'd and packed at 56 bits) -> SEEDED wire form */
void so_wire_query_pack_seeded(const uint8_t seed[32], const uint64_t *row1_ntt, uint8_t *wire);

/* records (flat little-endian bit stream, log2(p_db) bits per coefficient, item-major, polynomial-major inside an
 * item) -> plaintext coefficients u64 (what generate_random_pt would have produced, src/util.cpp:77-86) */
void so_records_to_plaintexts(uint64_t *pts, const uint8_t *records, size_t ncoeffs, uint64_t p_db);

#ifdef __cplusplus
}
#endif
