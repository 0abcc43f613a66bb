/* wire_format.c - see wire_format.h.  TEST INFRASTRUCTURE ONLY. */
#include "wire_format.h"
#include <stdlib.h>
#include <string.h>

#define N SO_N
#define PL (2 * (size_t)SO_N)

static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define QR(a, b, c, d) do { a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
                            a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7); } while (0)

void so_chacha20_block(const uint32_t key[8], uint32_t counter, const uint32_t nonce[3], uint32_t out[16]) {   /* RFC 8439 2.3 */
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) s[4 + i] = key[i];
    s[12] = counter; s[13] = nonce[0]; s[14] = nonce[1]; s[15] = nonce[2];
    uint32_t x[16];
    memcpy(x, s, sizeof(x));
    for (int r = 0; r < 10; r++) {
        QR(x[0], x[4], x[8], x[12]); QR(x[1], x[5], x[9], x[13]); QR(x[2], x[6], x[10], x[14]); QR(x[3], x[7], x[11], x[15]);
        QR(x[0], x[5], x[10], x[15]); QR(x[1], x[6], x[11], x[12]); QR(x[2], x[7], x[8], x[13]); QR(x[3], x[4], x[9], x[14]);
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

static uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

void so_wire_seeded_row0(const uint8_t seed[32], uint64_t *row0) {
    uint32_t key[8], nonce[3] = {SO_WIRE_QUERY_MAGIC, 0, 0}, blk[16];
    for (int i = 0; i < 8; i++) key[i] = le32(seed + 4 * i);
    for (uint32_t n = 0; n < 2; n++) {
        const uint32_t q = (uint32_t)(n == 0 ? SO_P : SO_B);
        for (uint32_t z = 0; z < N; z++) {
            so_chacha20_block(key, n * N + z, nonce, blk);
            uint32_t v = (blk[15] & 0x0FFFFFFFu) % q;
            for (int k = 0; k < 16; k++) {
                uint32_t c = blk[k] & 0x0FFFFFFFu;
                if (c < q) { v = c; break; }
            }
            row0[n * N + z] = v;
        }
    }
}

size_t so_wire_query_bytes(uint32_t kind) {
    if (kind == SO_WIRE_QUERY_SEEDED) return SO_WIRE_HEADER_BYTES + SO_WIRE_SEED_BYTES + SO_WIRE_ROW_BYTES;
    if (kind == SO_WIRE_QUERY_FULL) return SO_WIRE_HEADER_BYTES + 2 * SO_WIRE_ROW_BYTES;
    return 0;
}

static void unpack_row(uint64_t *row_ntt, const uint8_t *packed) {
    uint64_t *words = (uint64_t *)malloc(SO_WIRE_ROW_BYTES + 16), *raw = (uint64_t *)malloc(N * sizeof(uint64_t));
    memset(words, 0, SO_WIRE_ROW_BYTES + 16);
    memcpy(words, packed, SO_WIRE_ROW_BYTES);
    for (size_t i = 0; i < N; i++) raw[i] = so_read_arbitrary_bits(words, i * SO_LOGQ, SO_LOGQ);
    so_to_ntt(row_ntt, raw, 1);                       /* reduces any 56-bit value modulo both primes */
    free(words); free(raw);
}
static void pack_row(uint8_t *packed, const uint64_t *row_ntt) {
    uint64_t *words = (uint64_t *)calloc(SO_WIRE_ROW_BYTES / 8 + 2, 8), *raw = (uint64_t *)malloc(N * sizeof(uint64_t));
    so_from_ntt(raw, row_ntt, 1);
    for (size_t i = 0; i < N; i++) so_write_arbitrary_bits(words, raw[i], i * SO_LOGQ, SO_LOGQ);
    memcpy(packed, words, SO_WIRE_ROW_BYTES);
    free(words); free(raw);
}

int so_wire_query_expand(const uint8_t *wire, size_t bytes, uint64_t *query_cv) {
    if (!wire || bytes < SO_WIRE_HEADER_BYTES || le32(wire) != SO_WIRE_QUERY_MAGIC) return -1;
    const uint32_t kind = (uint32_t)wire[4] | ((uint32_t)wire[5] << 8);
    if (wire[6] || wire[7] || so_wire_query_bytes(kind) == 0 || bytes != so_wire_query_bytes(kind)) return -1;
    const uint8_t *p = wire + SO_WIRE_HEADER_BYTES;
    if (kind == SO_WIRE_QUERY_SEEDED) {
        so_wire_seeded_row0(p, query_cv);
        unpack_row(query_cv + PL, p + SO_WIRE_SEED_BYTES);
    } else {
        unpack_row(query_cv, p);
        unpack_row(query_cv + PL, p + SO_WIRE_ROW_BYTES);
    }
    return 0;
}

static void put_header(uint8_t *wire, uint32_t kind) {
    const uint32_t m = SO_WIRE_QUERY_MAGIC;
    wire[0] = (uint8_t)m; wire[1] = (uint8_t)(m >> 8); wire[2] = (uint8_t)(m >> 16); wire[3] = (uint8_t)(m >> 24);
    wire[4] = (uint8_t)kind; wire[5] = (uint8_t)(kind >> 8); wire[6] = 0; wire[7] = 0;
}
void so_wire_query_pack_full(const uint64_t *query_cv, uint8_t *wire) {
    put_header(wire, SO_WIRE_QUERY_FULL);
    pack_row(wire + SO_WIRE_HEADER_BYTES, query_cv);
    pack_row(wire + SO_WIRE_HEADER_BYTES + SO_WIRE_ROW_BYTES, query_cv + PL);
}
/* used by client_sim.c: header + seed + packed row 1 */
void so_wire_query_pack_seeded(const uint8_t seed[32], const uint64_t *row1_ntt, uint8_t *wire) {
    put_header(wire, SO_WIRE_QUERY_SEEDED);
    memcpy(wire + SO_WIRE_HEADER_BYTES, seed, SO_WIRE_SEED_BYTES);
    pack_row(wire + SO_WIRE_HEADER_BYTES + SO_WIRE_SEED_BYTES, row1_ntt);
}

void so_records_to_plaintexts(uint64_t *pts, const uint8_t *records, size_t ncoeffs, uint64_t p_db) {
    uint32_t bits = 0;
    while (((uint64_t)1 << bits) < p_db) bits++;
    for (size_t k = 0; k < ncoeffs; k++) {
        uint64_t v = 0;
        for (uint32_t b = 0; b < bits; b++) {
            size_t bit = k * bits + b;
            v |= (uint64_t)((records[bit >> 3] >> (bit & 7)) & 1) << b;
        }
        pts[k] = v;
    }
}
