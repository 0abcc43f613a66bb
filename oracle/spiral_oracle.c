/*
 * spiral_oracle.c - CPU restatement (plain C) of the Spiral server-side query-answering path.
 *
 * TEST INFRASTRUCTURE ONLY - see spiral_oracle.h.  Parity is pinned against the unmodified
 * reference (oracle/_ref) by tests/golden/ref_digests.json (generator: oracle/ref_golden.cpp).
 *
 * Each function names the reference lines it restates (paths relative to /root/reference).
 * The restatement keeps the reference's observable arithmetic, including its quirks:
 *   - automorph / invert produce Q - a, hence Q (not 0) for a == 0      src/poly.cpp:256,279
 *   - split_and_crt's signed digits with a per-half carry reset          src/spiral.cpp:282,312
 *   - shift clamp min(k*bits_per, 64)                                    src/util.cpp:139
 *   - rescale's truncating division with sign-dependent rounding         src/poly.cpp:578-591
 * NTT-domain outputs here are always canonical ([0,q)); the reference's AVX2 NTT may leave the
 * value q in place of 0 (strict '>' at src/core.cpp:342-349), so NTT-domain buffers are
 * compared modulo the prime and raw-domain buffers exactly.
 */
#include "spiral_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define N SO_N
#define P SO_P
#define B_ SO_B
#define Q SO_Q
typedef unsigned __int128 u128;

static const uint64_t PSI_P = 66687ull, PSI_B = 158221ull;   /* primitive 4096-th roots */
static const uint64_t CR1_P = 68736257792ull;                /* include/values.h:59 */
static const uint64_t CR1_B = 73916747789ull;                /* include/values.h:61 */
static const uint64_t CR0_Q = 7906011006380390721ull;        /* include/values.h:26 */
static const uint64_t CR1_Q = 275ull;                        /* include/values.h:27 */
static const u128 PA_INV_B = (u128)97389680ull * SO_P;       /* include/values.h:24 */
static const u128 B_INV_PA = (u128)163640210ull * SO_B;      /* include/values.h:25 */

static uint64_t *xalloc(size_t words) {
    uint64_t *p = (uint64_t *)calloc(words ? words : 1, sizeof(uint64_t));
    if (!p) abort();
    return p;
}

/* ------------------------------------------------------------------------------------------
 * tables (src/constants.cpp:16; layout comment src/core.cpp:6-17), regenerated from psi.
 * ---------------------------------------------------------------------------------------- */
static uint64_t g_tables[8 * N];
static int g_tables_ready = 0;

static uint64_t powmod(uint64_t a, uint64_t e, uint64_t q) {
    uint64_t r = 1; a %= q;
    while (e) { if (e & 1) r = (uint64_t)((u128)r * a % q); a = (uint64_t)((u128)a * a % q); e >>= 1; }
    return r;
}
static uint32_t bitrev11(uint32_t x) {
    uint32_t r = 0;
    for (unsigned i = 0; i < SO_LOGN; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}
const uint64_t *so_tables(void) {
    if (g_tables_ready) return g_tables;
    const uint64_t mods[2] = {P, B_}, psis[2] = {PSI_P, PSI_B};
    for (int c = 0; c < 2; c++) {
        uint64_t q = mods[c], psi = psis[c];
        uint64_t psi_inv = powmod(psi, q - 2, q), half = (q + 1) / 2;
        uint64_t *inv = &g_tables[(size_t)c * 2 * N], *inv_s = inv + N;
        uint64_t *fwd = &g_tables[(size_t)(4 + c * 2) * N], *fwd_s = fwd + N;
        uint64_t a = 1, b = 1;
        for (uint32_t i = 0; i < N; i++) {
            uint32_t br = bitrev11(i);
            fwd[br] = a;                                    /* psi^i at bit-reversed slot       */
            inv[br] = (uint64_t)((u128)b * half % q);       /* psi^-i / 2 at bit-reversed slot  */
            a = (uint64_t)((u128)a * psi % q);
            b = (uint64_t)((u128)b * psi_inv % q);
        }
        for (uint32_t i = 0; i < N; i++) {
            fwd_s[i] = (uint64_t)(((u128)fwd[i] << 32) / q);
            inv_s[i] = (uint64_t)(((u128)inv[i] << 32) / q);
        }
    }
    g_tables_ready = 1;
    return g_tables;
}

uint64_t so_arb_qprime(uint32_t qp_bits) {                  /* include/values.h:74-76 */
    static const uint64_t qprime_mods[37] = {
        0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 12289, 12289, 61441, 65537, 65537, 520193, 786433,
        786433, 3604481, 7340033, 16515073, 33292289, 67043329, 132120577, 268369921, 469762049,
        1073479681, 2013265921, 4293918721ull, 8588886017ull, 17175674881ull, 34359214081ull,
        68718428161ull};
    return qp_bits < 37 ? qprime_mods[qp_bits] : 0;
}

uint32_t so_get_bits_per(uint32_t dim) {                    /* include/util.h:34-38 */
    if (dim == SO_LOGQ) return 1;
    return (uint32_t)floor(SO_LOGQ / (double)dim) + 1;
}

static inline uint64_t barrett_raw_u64(uint64_t input, uint64_t cr1, uint64_t modulus) {   /* include/poly.h:137-146 */
    uint64_t hi = (uint64_t)(((u128)input * cr1) >> 64);
    uint64_t t = input - hi * modulus;
    return t >= modulus ? t - modulus : t;
}
uint64_t so_barrett_coeff(uint64_t val, int n) {            /* include/poly.h:148-153 */
    return n == 0 ? barrett_raw_u64(val, CR1_P, P) : barrett_raw_u64(val, CR1_B, B_);
}

static inline uint64_t barrett_reduction_u128(u128 val) {   /* src/poly.cpp:11-32 */
    uint64_t zx = (uint64_t)val, zy = (uint64_t)(val >> 64);
    uint64_t carry = (uint64_t)(((u128)zx * CR0_Q) >> 64);
    u128 t2 = (u128)zx * CR1_Q;
    uint64_t tmp1 = (uint64_t)t2 + carry;
    uint64_t tmp3 = (uint64_t)(t2 >> 64) + (tmp1 < (uint64_t)t2 ? 1 : 0);
    t2 = (u128)zy * CR0_Q;
    uint64_t s = tmp1 + (uint64_t)t2;
    carry = (uint64_t)(t2 >> 64) + (s < tmp1 ? 1 : 0);
    tmp1 = zy * CR1_Q + tmp3 + carry;
    uint64_t r = zx - tmp1 * Q;
    r -= Q * (uint64_t)(r >= Q);
    return r;
}
uint64_t so_crt_compose(uint64_t x, uint64_t y) {           /* src/poly.cpp:344-353 */
    u128 val = (u128)x * B_INV_PA;
    val += (u128)y * PA_INV_B;
    return barrett_reduction_u128(val);
}

uint64_t so_rescale(uint64_t a, uint64_t inp_mod, uint64_t out_mod) {   /* src/poly.cpp:578-591 */
    int64_t inp_val = (int64_t)(a % inp_mod);
    if (inp_val >= (int64_t)(inp_mod / 2)) inp_val -= (int64_t)inp_mod;
    int64_t sign = inp_val >= 0 ? 1 : -1;
    __int128 val = inp_val * (__int128)out_mod;
    __int128 result = (val + sign * ((int64_t)inp_mod / 2)) / (__int128)inp_mod;
    result = (result + (inp_mod / out_mod) * out_mod + 2 * out_mod) % out_mod;
    return (uint64_t)((result + out_mod) % out_mod);
}
void so_get_rescaled(uint64_t *out, const uint64_t *in, size_t ncoeffs, uint64_t inp_mod, uint64_t out_mod) {
    for (size_t i = 0; i < ncoeffs; i++) out[i] = so_rescale(in[i] % Q, inp_mod, out_mod);   /* src/poly.cpp:593-601 */
}

/* ------------------------------------------------------------------------------------------
 * bit I/O (src/core.cpp:20-52) and modswitch (src/spiral.cpp:40-78)
 * ---------------------------------------------------------------------------------------- */
uint64_t so_read_arbitrary_bits(const uint64_t *p, size_t bit_offs, size_t num_bits) {   /* src/core.cpp:20-30 */
    size_t word_off = bit_offs / 64, in_word = bit_offs % 64;
    uint64_t mask = (1ull << num_bits) - 1;
    if (in_word + num_bits <= 64) return (p[word_off] >> in_word) & mask;
    u128 v = (u128)p[word_off] | ((u128)p[word_off + 1] << 64);
    return (uint64_t)(v >> in_word) & mask;
}
void so_write_arbitrary_bits(uint64_t *p, uint64_t val, size_t bit_offs, size_t num_bits) {   /* src/core.cpp:32-52 */
    size_t word_off = bit_offs / 64, in_word = bit_offs % 64;
    uint64_t mask = (1ull << num_bits) - 1;
    val &= mask;
    p[word_off] = (p[word_off] & ~(mask << in_word)) | (val << in_word);
    if (in_word + num_bits > 64) {
        size_t spill = in_word + num_bits - 64;                     /* high bits land in the next word */
        p[word_off + 1] = (p[word_off + 1] & ~((1ull << spill) - 1)) | (val >> (num_bits - spill));
    }
}
size_t so_packed_words(size_t ncoeffs, uint32_t bits) { return (ncoeffs * bits + 63) / 64; }

/* modswitch's arithmetic (src/spiral.cpp:58-64) is x87 extended precision:
 *     long double d = (long double)val * (long double)arb_qprime / (long double)Q_i;  result = round(d)   [roundl]
 * i.e. P' = RNE64(val * q')  (the exact product can need 93 bits; round-to-nearest-even to a 64-bit significand),
 *      D  = RNE64(P' / Q), result = floor(D + 1/2) (half away from zero, D >= 0).
 * Restated exactly in integers: with I = floor(P'/Q), R = P' mod Q and f = 64 - bitlen(I) fraction bits in D,
 * D >= I + 1/2  <=>  R/Q >= 1/2 - 2^-(f+1)  (I + 1/2 is representable and its significand is even, so the tie of the
 * second rounding goes up)  <=>  2R >= Q  or  Q - 2R <= floor(Q / 2^f).  For I = 0 the bound is 2R >= Q. */
static unsigned bitlen_u128(u128 v) { unsigned n = 0; while (v) { n++; v >>= 1; } return n; }
uint64_t so_modswitch_coeff(uint64_t val, uint64_t qprime) {
    u128 P128 = (u128)val * qprime;
    unsigned L = bitlen_u128(P128);
    if (L > 64) {                                                   /* first rounding: product to 64 significant bits */
        unsigned s = L - 64;
        u128 m = P128 >> s, rem = P128 & (((u128)1 << s) - 1), half = (u128)1 << (s - 1);
        if (rem > half || (rem == half && (m & 1))) m++;
        P128 = m << s;
    }
    uint64_t I = (uint64_t)(P128 / Q), R = (uint64_t)(P128 % Q);
    int up;
    if (2 * R >= Q) up = 1;
    else if (I == 0) up = 0;
    else {
        unsigned f = 64 - bitlen_u128(I);
        up = (Q - 2 * R) <= (f >= 64 ? 0 : (Q >> f));
    }
    return I + (uint64_t)up;
}
uint64_t so_modswitch_coeff_x87(uint64_t val, uint64_t qprime) {
#if defined(__x86_64__) || defined(__i386__)
    long double d = ((long double)(int64_t)val) * ((long double)qprime) / ((long double)Q);
    return (uint64_t)(int64_t)roundl(d);
#else
    return so_modswitch_coeff(val, qprime);
#endif
}
void so_modswitch(uint64_t *out_words, const uint64_t *in_raw, size_t ncoeffs, uint32_t qp_bits) {   /* src/spiral.cpp:40-78 */
    uint64_t qprime = so_arb_qprime(qp_bits);
    size_t bit_offs = 0;
    for (size_t i = 0; i < ncoeffs; i++) {          /* r, c, m loops of the reference = one linear pass */
        so_write_arbitrary_bits(out_words, so_modswitch_coeff(in_raw[i], qprime), bit_offs, qp_bits);
        bit_offs += qp_bits;
    }
}

/* ------------------------------------------------------------------------------------------
 * NTT  (src/core.cpp:254-416 forward, :426-514 inverse; scalar statements)
 * ---------------------------------------------------------------------------------------- */
void so_ntt_forward(uint64_t *op_all) {
    const uint64_t *tb = so_tables();
    for (int cm = 0; cm < 2; cm++) {
        const uint64_t *fw = &tb[(size_t)N * 4 + (size_t)cm * N * 2];
        uint64_t *op = &op_all[(size_t)cm * N];
        uint32_t q = cm == 0 ? (uint32_t)P : (uint32_t)B_, q2 = 2 * q;
        for (unsigned mm = 0; mm < SO_LOGN; mm++) {
            size_t m = (size_t)1 << mm, t = N >> (mm + 1);
            for (size_t i = 0; i < m; i++) {
                uint64_t W = fw[m + i], Wp = fw[N + m + i];
                for (size_t j = 0; j < t; j++) {
                    uint64_t *px = &op[2 * i * t + j], *py = &op[2 * i * t + t + j];
                    uint32_t x = (uint32_t)*px, y = (uint32_t)*py;
                    uint32_t cx = x - (q2 * (uint32_t)(x >= q2));
                    uint64_t Qv = ((uint64_t)y * Wp) >> 32;
                    Qv = W * y - Qv * q;
                    *px = cx + Qv;
                    *py = cx + (q2 - Qv);
                }
            }
        }
        for (size_t i = 0; i < N; i++) {
            op[i] -= (uint64_t)(op[i] >= q2) * q2;
            op[i] -= (uint64_t)(op[i] >= q) * q;
        }
    }
}

void so_ntt_inverse(uint64_t *op_all) {
    const uint64_t *tb = so_tables();
    for (int cm = 0; cm < 2; cm++) {
        const uint64_t *iv = &tb[(size_t)cm * N * 2];
        uint64_t *op = &op_all[(size_t)cm * N];
        uint64_t q = cm == 0 ? P : B_, q2 = 2 * q;
        size_t t = 1;
        for (size_t m = N; m > 1; m >>= 1) {
            size_t j1 = 0, h = m >> 1;
            for (size_t i = 0; i < h; i++) {
                uint64_t W = iv[h + i], Wp = iv[N + h + i];
                uint64_t *U = op + j1, *V = U + t;
                for (size_t j = 0; j < t; j++) {
                    uint64_t T = q2 - *V + *U;
                    uint64_t cu = *U + *V - (q2 * (uint64_t)((*U << 1) >= T));
                    *U++ = (cu + (q * (T & 1))) >> 1;
                    uint64_t H = (T * Wp) >> 32;
                    *V++ = W * T - H * q;
                }
                j1 += (t << 1);
            }
            t <<= 1;
        }
        for (size_t i = 0; i < N; i++) {
            op[i] -= (uint64_t)(op[i] >= q2) * q2;
            op[i] -= (uint64_t)(op[i] >= q) * q;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * MatPoly algebra  (src/poly.cpp)
 * ---------------------------------------------------------------------------------------- */
void so_to_ntt(uint64_t *out, const uint64_t *in, size_t npolys) {             /* :311-329 */
    for (size_t p = 0; p < npolys; p++) {
        for (int n = 0; n < 2; n++)
            for (size_t z = 0; z < N; z++) out[p * 2 * N + n * N + z] = so_barrett_coeff(in[p * N + z], n);
        so_ntt_forward(&out[p * 2 * N]);
    }
}
void so_to_ntt_no_reduce(uint64_t *out, const uint64_t *in, size_t npolys) {   /* :291-309 */
    for (size_t p = 0; p < npolys; p++) {
        for (int n = 0; n < 2; n++)
            for (size_t z = 0; z < N; z++) out[p * 2 * N + n * N + z] = in[p * N + z];
        so_ntt_forward(&out[p * 2 * N]);
    }
}
void so_from_ntt(uint64_t *out, const uint64_t *in, size_t npolys) {           /* :357-377 */
    uint64_t scratch[2 * N];
    for (size_t p = 0; p < npolys; p++) {
        memcpy(scratch, &in[p * 2 * N], sizeof(scratch));
        so_ntt_inverse(scratch);
        for (size_t z = 0; z < N; z++) out[p * N + z] = so_crt_compose(scratch[z], scratch[N + z]);
    }
}
void so_multiply(uint64_t *out, const uint64_t *a, const uint64_t *b, size_t rs, size_t ms, size_t cs) {  /* :34-78 */
    for (size_t r = 0; r < rs; r++)
        for (size_t c = 0; c < cs; c++) {
            uint64_t *acc = &out[(r * cs + c) * 2 * N];
            memset(acc, 0, 2 * N * sizeof(uint64_t));
            for (size_t m = 0; m < ms; m++) {
                const uint64_t *x = &a[(r * ms + m) * 2 * N], *y = &b[(m * cs + c) * 2 * N];
                for (size_t z = 0; z < 2 * N; z++) acc[z] += x[z] * y[z];   /* wrapping u64, as the reference */
            }
            for (size_t z = 0; z < N; z++) { acc[z] %= P; acc[N + z] %= B_; }
        }
}
void so_add(uint64_t *out, const uint64_t *a, const uint64_t *b, size_t npolys) {   /* :138-155 */
    for (size_t p = 0; p < npolys; p++)
        for (int n = 0; n < 2; n++)
            for (size_t z = 0; z < N; z++) {
                size_t i = p * 2 * N + n * N + z;
                out[i] = so_barrett_coeff(a[i] + b[i], n);
            }
}
void so_mul_by_const(uint64_t *out, const uint64_t *single, const uint64_t *a, size_t npolys) {   /* :190-214 */
    for (size_t p = 0; p < npolys; p++)
        for (int n = 0; n < 2; n++)
            for (size_t z = 0; z < N; z++) {
                size_t i = p * 2 * N + n * N + z;
                out[i] = so_barrett_coeff(a[i] * single[n * N + z], n);
            }
}
void so_automorph(uint64_t *out, const uint64_t *in, size_t npolys, uint64_t t) {   /* :240-261 */
    for (size_t p = 0; p < npolys; p++)
        for (size_t i = 0; i < N; i++) {
            uint64_t num = (i * t) / N, rem = (i * t) % N;
            out[p * N + rem] = (num % 2 == 0) ? in[p * N + i] : Q - in[p * N + i];
        }
}
void so_invert(uint64_t *out, const uint64_t *in, size_t npolys) {                 /* :269-283 */
    for (size_t i = 0; i < npolys * N; i++) out[i] = Q - in[i];
}

void so_build_gadget(uint64_t *G, size_t nx, size_t m) {                           /* src/util.cpp:89-106 */
    memset(G, 0, nx * m * N * sizeof(uint64_t));
    size_t num_elems = m / nx;
    uint32_t bits_per = so_get_bits_per((uint32_t)num_elems);
    for (size_t i = 0; i < nx; i++)
        for (size_t j = 0; j < num_elems; j++) {
            if ((uint64_t)bits_per * j >= 64) continue;
            G[(i * m + (i + j * nx)) * N] = 1ull << (bits_per * j);
        }
}

static inline uint64_t shr_clamped(uint64_t val, size_t bit_offs) {
    /* min(k*bits_per, 64) (src/util.cpp:139, src/spiral.cpp:285); a 64-bit shift by 64 is what
     * x86 SHR executes as a shift by 0.  Never reached for the surveyed parameter sets. */
    return val >> (bit_offs & 63);
}
void so_gadget_invert(uint64_t *out, const uint64_t *in, size_t mx, size_t rdim, size_t m) {   /* src/util.cpp:114-150 */
    size_t num_elems = mx / rdim;
    uint32_t bits_per = so_get_bits_per((uint32_t)num_elems);
    uint64_t mask = (1ull << bits_per) - 1;
    for (size_t i = 0; i < m; i++)
        for (size_t j = 0; j < rdim; j++)
            for (size_t z = 0; z < N; z++) {
                uint64_t val = in[(j * m + i) * N + z];
                for (size_t k = 0; k < num_elems; k++) {
                    size_t row = j + k * rdim;
                    size_t bit_offs = k * bits_per < 64 ? k * bits_per : 64;
                    out[(row * m + i) * N + z] = shr_clamped(val, bit_offs) & mask;
                }
            }
}

/* ------------------------------------------------------------------------------------------
 * Spiral / SpiralStream server path  (src/spiral.cpp)
 * ---------------------------------------------------------------------------------------- */
void so_encode_plaintext(uint64_t *out, const uint64_t *pt, size_t ncoeffs, uint64_t p_db) {   /* :1116-1127 */
    for (size_t i = 0; i < ncoeffs; i++) {
        int64_t val = (int64_t)pt[i];
        if (val >= (int64_t)(p_db / 2)) val -= (int64_t)p_db;
        if (val < 0) val += (int64_t)Q;
        out[i] = (uint64_t)val;
    }
}

void so_load_db(uint64_t *Bbuf, const uint64_t *pts, uint32_t nu1, uint32_t nu2, uint64_t p_db) {   /* :1083-1156 */
    size_t dim0 = (size_t)1 << nu1, num_per = (size_t)1 << nu2, total_n = dim0 * num_per;
    uint64_t *raw = xalloc(SO_N0 * SO_N2 * N), *enc = xalloc(SO_N0 * SO_N2 * 2 * N);
    for (size_t i = 0; i < total_n; i++) {
        so_encode_plaintext(raw, &pts[i * SO_N0 * SO_N2 * N], SO_N0 * SO_N2 * N, p_db);
        so_to_ntt(enc, raw, SO_N0 * SO_N2);
        size_t ii = i % num_per, j = i / num_per;
        for (size_t m = 0; m < SO_N0; m++)
            for (size_t c = 0; c < SO_N2; c++) {
                const uint64_t *BB = &enc[(m * SO_N2 + c) * 2 * N];
                for (size_t z = 0; z < N; z++) {
                    size_t idx = z * (num_per * SO_N2 * dim0 * SO_N0) + ii * (SO_N2 * dim0 * SO_N0) +
                                 c * (dim0 * SO_N0) + j * SO_N0 + m;
                    Bbuf[idx] = BB[z] | (BB[N + z] << 32);
                }
            }
    }
    free(raw); free(enc);
}

void so_reorient_ciphertexts(uint64_t *out, const uint64_t *inp, size_t dim0, size_t n1_padded) {   /* :410-433 */
    for (size_t j = 0; j < dim0; j++)
        for (size_t r = 0; r < SO_N1; r++)
            for (size_t m = 0; m < 2; m++)
                for (size_t z = 0; z < N; z++) {
                    size_t in_i = j * (SO_N1 * 2 * 2 * N) + r * (2 * 2 * N) + m * (2 * N);
                    size_t out_i = z * (dim0 * 2 * n1_padded) + j * (2 * n1_padded) + m * n1_padded + r;
                    out[out_i] = inp[in_i + z] | (inp[in_i + N + z] << 32);
                }
}

void so_multiply_query_by_database(uint64_t *out, const uint64_t *reor, const uint64_t *db,
                                   size_t dim0, size_t num_per) {   /* :932-997 (scalar statement) */
    const size_t n1p = 4;
    for (size_t z = 0; z < N; z++) {
        size_t a_base = z * (2 * dim0 * n1p);
        size_t b_idx = z * (num_per * SO_N2 * dim0 * SO_N0);
        for (size_t i = 0; i < num_per; i++)
            for (size_t c = 0; c < SO_N2; c++) {
                u128 s0[3] = {0, 0, 0}, s1[3] = {0, 0, 0};
                for (size_t jm = 0; jm < dim0 * 2; jm++) {
                    uint64_t b = db[b_idx++];
                    const uint64_t *va = &reor[a_base + jm * n1p];
                    uint32_t b_lo = (uint32_t)b, b_hi = (uint32_t)(b >> 32);
                    for (int r = 0; r < 3; r++) {
                        s0[r] += (uint64_t)(uint32_t)va[r] * b_lo;
                        s1[r] += (uint64_t)(uint32_t)(va[r] >> 32) * b_hi;
                    }
                }
                for (int r = 0; r < 3; r++) {
                    size_t idx = i * (SO_N1 * SO_N2 * 2 * N) + (size_t)r * (SO_N2 * 2 * N) + c * (2 * N) + z;
                    out[idx] = (uint64_t)(s0[r] % P);
                    out[idx + N] = (uint64_t)(s1[r] % B_);
                }
            }
    }
}

void so_cpu_crt(uint64_t *out, const uint64_t *inp, size_t num_polys) {   /* :586-593 */
    for (size_t i = 0; i < num_polys; i++)
        for (size_t j = 0; j < N; j++) out[i * N + j] = so_crt_compose(inp[i * 2 * N + j], inp[i * 2 * N + N + j]);
}
void so_cpu_crt_to_ucompressed_and_ntt(uint64_t *out, const uint64_t *inp, size_t num_polys) {   /* :597-609 */
    for (size_t i = 0; i < num_polys; i++) {
        for (size_t j = 0; j < N; j++) {
            out[i * 2 * N + j] = inp[i * N + j] % P;
            out[i * 2 * N + N + j] = inp[i * N + j] % B_;
        }
        so_ntt_forward(&out[i * 2 * N]);
    }
}
void so_ntt_inv_and_crt_lift(uint64_t *cts, uint64_t *scratch, size_t num_per) {   /* :437-453 */
    for (size_t i = 0; i < num_per * SO_N1 * SO_N2; i++) so_ntt_inverse(&scratch[2 * N * i]);
    so_cpu_crt(cts, scratch, num_per * SO_N1 * SO_N2);
}

void so_split_and_crt(uint64_t *out, const uint64_t *in, size_t num_per, uint32_t t_gsw) {   /* :270-341 */
    size_t m2 = (size_t)SO_N1 * t_gsw, num_elems = t_gsw;
    uint32_t bits_per = so_get_bits_per(t_gsw);
    uint64_t mask = (1ull << bits_per) - 1;
    for (size_t i = 0; i < num_per; i++)
        for (size_t r = 0; r < SO_N1; r++)
            for (size_t c = 0; c < SO_N2; c++) {
                const uint64_t *src = &in[i * (SO_N1 * SO_N2 * N) + r * (SO_N2 * N) + c * N];
                for (int half = 0; half < 2; half++) {
                    size_t k0 = half == 0 ? 0 : num_elems / 2, k1 = half == 0 ? num_elems / 2 : num_elems;
                    for (size_t z = 0; z < N; z++) {
                        uint64_t val = src[z], carry = 0;
                        for (size_t k = k0; k < k1; k++) {
                            size_t row = r + k * SO_N1;
                            size_t bit_offs = k * bits_per < 64 ? k * bits_per : 64;
                            uint64_t piece = shr_clamped(val, bit_offs) & mask;
                            piece += carry;
                            carry = 0;
                            int last_guard = half == 0 ? (k < (num_elems / 2 - 1)) : 1;
                            if (piece > (uint64_t)((1 << bits_per) / 2) && last_guard) {
                                piece += Q - ((uint64_t)1 << bits_per);
                                carry = 1;
                            }
                            size_t oi = i * (m2 * SO_N2 * 2 * N) + row * (SO_N2 * 2 * N) + c * (2 * N);
                            out[oi + z] = so_barrett_coeff(piece, 0);
                            out[oi + N + z] = so_barrett_coeff(piece, 1);
                        }
                    }
                    for (size_t k = k0; k < k1; k++) {
                        size_t row = r + k * SO_N1;
                        so_ntt_forward(&out[i * (m2 * SO_N2 * 2 * N) + row * (SO_N2 * 2 * N) + c * (2 * N)]);
                    }
                }
            }
}

void so_reorient_C(uint64_t *out, const uint64_t *inp, size_t num_per, uint32_t t_gsw) {   /* :345-384 */
    size_t m2 = (size_t)SO_N1 * t_gsw;
    for (size_t i = 0; i < num_per; i++)
        for (size_t m = 0; m < m2; m++)
            for (size_t c = 0; c < SO_N2; c++)
                for (size_t z = 0; z < N; z++) {
                    size_t in_i = i * (m2 * SO_N2 * 2 * N) + m * (SO_N2 * 2 * N) + c * (2 * N);
                    size_t out_i = z * (num_per * SO_N2 * m2) + i * (SO_N2 * m2) + c * m2 + m;
                    out[out_i] = inp[in_i + z] | (inp[in_i + N + z] << 32);
                }
}
void so_reorient_Q(uint64_t *out, const uint64_t *inp, uint32_t t_gsw) {   /* :388-400 */
    size_t m2 = (size_t)SO_N1 * t_gsw;
    for (size_t r = 0; r < SO_N1; r++)
        for (size_t m = 0; m < m2; m++)
            for (size_t z = 0; z < N; z++) {
                size_t in_i = r * (m2 * 2 * N) + m * (2 * N);
                out[z * (SO_N1 * m2) + r * m2 + m] = inp[in_i + z] | (inp[in_i + N + z] << 32);
            }
}
void so_cpu_mul_query_by_ct(uint64_t *C_next, const uint64_t *Qr, const uint64_t *C, size_t num_per, uint32_t t_gsw) {   /* :464-582 */
    size_t m2 = (size_t)SO_N1 * t_gsw;
    for (size_t z = 0; z < N; z++)
        for (size_t i = 0; i < num_per; i++)
            for (size_t r = 0; r < SO_N1; r++)
                for (size_t c = 0; c < SO_N2; c++) {
                    const uint64_t *Cp = &C[z * (num_per * SO_N2 * m2) + i * (SO_N2 * m2) + c * m2];
                    const uint64_t *Qp = &Qr[z * (SO_N1 * m2) + r * m2];
                    uint64_t s0 = 0, s1 = 0;
                    for (size_t m = 0; m < m2; m++) {
                        s0 += (uint64_t)(uint32_t)Qp[m] * (uint32_t)Cp[m];
                        s1 += (uint64_t)(uint32_t)(Qp[m] >> 32) * (uint32_t)(Cp[m] >> 32);
                    }
                    uint64_t *o = &C_next[(i * (SO_N1 * SO_N2) + r * SO_N2 + c) * 2 * N];
                    o[z] = so_barrett_coeff(s0, 0);
                    o[N + z] = so_barrett_coeff(s1, 1);
                }
}

void so_fold_one_further_dimension(size_t cur_dim, size_t num_per, const uint64_t *q_reor,
                                   const uint64_t *q_neg_reor, uint64_t *g_C, uint32_t t_gsw) {   /* :1349-1410 */
    size_t m2 = (size_t)SO_N1 * t_gsw, ct = SO_N1 * SO_N2;
    uint64_t *big1 = xalloc(num_per * m2 * SO_N2 * 2 * N), *big2 = xalloc(num_per * m2 * SO_N2 * N);
    uint64_t *int1 = xalloc(num_per * ct * 2 * N), *int2 = xalloc(num_per * ct * 2 * N);
    so_split_and_crt(big1, &g_C[num_per * ct * N], num_per, t_gsw);
    so_reorient_C(big2, big1, num_per, t_gsw);
    /* per-dimension stride is the reference's (n1*m2*crt_count*poly_len words, :1365), twice the packed size */
    so_cpu_mul_query_by_ct(int2, &q_reor[cur_dim * (SO_N1 * m2 * 2 * N)], big2, num_per, t_gsw);
    so_split_and_crt(big1, g_C, num_per, t_gsw);
    so_reorient_C(big2, big1, num_per, t_gsw);
    so_cpu_mul_query_by_ct(int1, &q_neg_reor[cur_dim * (SO_N1 * m2 * 2 * N)], big2, num_per, t_gsw);
    for (size_t i = 0; i < num_per * ct; i++)
        for (size_t z = 0; z < N; z++)
            for (int n = 0; n < 2; n++) {
                size_t idx = i * 2 * N + n * N + z;
                int1[idx] = so_barrett_coeff(int1[idx] + int2[idx], n);
            }
    for (size_t i = 0; i < num_per * ct; i++) so_ntt_inverse(&int1[2 * N * i]);
    so_cpu_crt(g_C, int1, num_per * ct);
    free(big1); free(big2); free(int1); free(int2);
}

/* neg1s_mp[r] = to_ntt(invert(x^(N - 2^r)))  (src/spiral.cpp:184-192, src/testing.cpp:755-765) */
static void make_neg1(uint64_t *out_ntt, size_t r) {
    uint64_t *raw = xalloc(N), *inv = xalloc(N);
    raw[N - ((size_t)1 << r)] = 1;
    so_invert(inv, raw, 1);
    so_to_ntt(out_ntt, inv, 1);
    free(raw); free(inv);
}

void so_expand_improved(uint64_t *cv, size_t g, uint32_t t_exp, const uint64_t *W_left,
                        const uint64_t *W_right, uint32_t t_exp_right, size_t max_bits_right,
                        size_t stopround) {   /* :1664-1743 == src/testing.cpp:40-105 */
    const size_t CT = 2 * 2 * N;   /* one 2x1 NTT ciphertext */
    uint64_t *c = xalloc(2 * N), *c_auto = xalloc(2 * N), *c1_ntt = xalloc(2 * N), *neg1 = xalloc(2 * N);
    size_t tmax = t_exp > t_exp_right ? t_exp : t_exp_right;
    uint64_t *ginv = xalloc(tmax * N), *ginv_ntt = xalloc(tmax * 2 * N), *Wg = xalloc(2 * 2 * N);
    for (size_t r = 0; r < g; r++) {
        size_t num_in = (size_t)1 << r, num_out = 2 * num_in;
        uint64_t t = (N / ((uint64_t)1 << r)) + 1;
        const uint64_t *Wl = &W_left[r * (2 * (size_t)t_exp * 2 * N)];
        const uint64_t *Wr = &W_right[r * (2 * (size_t)t_exp_right * 2 * N)];
        make_neg1(neg1, r);
        for (size_t i = 0; i < num_out; i++) {
            if (stopround > 0 && r > stopround && (i % 2) == 1) continue;
            if (stopround > 0 && r == stopround && (i % 2) == 1 && i / 2 > max_bits_right) continue;
            const uint64_t *W = (i % 2) == 0 ? Wl : Wr;
            size_t gd = (i % 2) == 0 ? t_exp : t_exp_right;
            if (i < num_in) so_mul_by_const(&cv[(num_in + i) * CT], neg1, &cv[i * CT], 2);
            so_from_ntt(c, &cv[i * CT], 2);
            so_automorph(c_auto, c, 2, t);
            so_to_ntt(c1_ntt, &c_auto[N], 1);
            so_gadget_invert(ginv, c_auto, gd, 1, 1);
            so_to_ntt_no_reduce(ginv_ntt, ginv, gd);
            so_multiply(Wg, W, ginv_ntt, 2, gd, 1);
            size_t idx = 0;
            for (size_t j = 0; j < 2; j++)
                for (int n = 0; n < 2; n++)
                    for (size_t z = 0; z < N; z++) {
                        cv[i * CT + idx] = so_barrett_coeff(cv[i * CT + idx] + Wg[idx] + j * c1_ntt[n * N + z], n);
                        idx++;
                    }
        }
    }
    free(c); free(c_auto); free(c1_ntt); free(neg1); free(ginv); free(ginv_ntt); free(Wg);
}

/* prod[r][c] = sum_k W[r][2k+c] * a_k, the effect of special_distribute (:1834-1848) + multiply */
static void scal_to_mat_fast(uint64_t *out_reg, const uint64_t *cv, const uint64_t *ginv_raw_ntt,
                             const uint64_t *W, uint32_t t_conv) {   /* :1887-1917 */
    size_t mc2 = 2 * (size_t)t_conv;
    uint64_t *dist = xalloc(mc2 * 2 * 2 * N), *prod = xalloc(SO_N1 * 2 * 2 * N), *pad = xalloc(SO_N1 * 2 * 2 * N);
    for (size_t i = 0; i < t_conv; i++) {
        memcpy(&dist[((2 * i) * 2 + 0) * 2 * N], &ginv_raw_ntt[i * 2 * N], 2 * N * sizeof(uint64_t));
        memcpy(&dist[((2 * i + 1) * 2 + 1) * 2 * N], &ginv_raw_ntt[i * 2 * N], 2 * N * sizeof(uint64_t));
    }
    so_multiply(prod, W, dist, SO_N1, mc2, 2);
    memcpy(&pad[(1 * 2 + 0) * 2 * N], &cv[2 * N], 2 * N * sizeof(uint64_t));   /* place(cv_1, 1, 0) */
    memcpy(&pad[(2 * 2 + 1) * 2 * N], &cv[2 * N], 2 * N * sizeof(uint64_t));   /* place(cv_1, 2, 1) */
    so_add(out_reg, prod, pad, SO_N1 * 2);
    free(dist); free(prod); free(pad);
}

void so_scal_to_mat(uint64_t *out_reg, const uint64_t *cv, const uint64_t *W, uint32_t t_conv) {   /* :1850-1885 */
    uint64_t *raw = xalloc(N), *ginv = xalloc((size_t)t_conv * N), *ginv_ntt = xalloc((size_t)t_conv * 2 * N);
    so_from_ntt(raw, cv, 1);
    so_gadget_invert(ginv, raw, t_conv, 1, 1);
    so_to_ntt_no_reduce(ginv_ntt, ginv, t_conv);
    scal_to_mat_fast(out_reg, cv, ginv_ntt, W, t_conv);
    free(raw); free(ginv); free(ginv_ntt);
}

void so_regev_to_gsw(uint64_t *out, const uint64_t *cv_v, uint32_t t_conv, uint32_t t,
                     const uint64_t *W, const uint64_t *V) {   /* :1985-2025 */
    const size_t CT = 2 * 2 * N, PL = 2 * N;
    size_t cols = (size_t)SO_N1 * t;
    uint64_t *cv_raw = xalloc(2 * N), *ginv = xalloc((size_t)t_conv * N), *g_ntt = xalloc((size_t)t_conv * PL);
    uint64_t *chat = xalloc(2 * (size_t)t_conv * t * PL), *s2m = xalloc(SO_N1 * 2 * PL);
    uint64_t *prod = xalloc(SO_N1 * (size_t)t * PL), *result = xalloc(SO_N1 * cols * PL);
    for (size_t i = 0; i < t; i++) {
        const uint64_t *cvi = &cv_v[i * CT];
        so_from_ntt(cv_raw, cvi, 2);
        so_gadget_invert(ginv, cv_raw, t_conv, 1, 1);
        so_to_ntt_no_reduce(g_ntt, ginv, t_conv);
        for (size_t k = 0; k < t_conv; k++) memcpy(&chat[(k * t + i) * PL], &g_ntt[k * PL], PL * sizeof(uint64_t));
        scal_to_mat_fast(s2m, cvi, g_ntt, W, t_conv);
        for (size_t r = 0; r < SO_N1; r++)
            for (size_t c = 0; c < 2; c++)
                memcpy(&result[(r * cols + t + 2 * i + c) * PL], &s2m[(r * 2 + c) * PL], PL * sizeof(uint64_t));
        so_gadget_invert(ginv, &cv_raw[N], t_conv, 1, 1);
        so_to_ntt_no_reduce(g_ntt, ginv, t_conv);
        for (size_t k = 0; k < t_conv; k++) memcpy(&chat[((t_conv + k) * t + i) * PL], &g_ntt[k * PL], PL * sizeof(uint64_t));
    }
    so_multiply(prod, V, chat, SO_N1, 2 * (size_t)t_conv, t);
    for (size_t r = 0; r < SO_N1; r++)
        for (size_t c = 0; c < t; c++) memcpy(&result[(r * cols + c) * PL], &prod[(r * t + c) * PL], PL * sizeof(uint64_t));
    for (size_t i = 0; i < t; i++)        /* permute :2019-2022 */
        for (size_t r = 0; r < SO_N1; r++) {
            memcpy(&out[(r * cols + 3 * i) * PL], &result[(r * cols + i) * PL], PL * sizeof(uint64_t));
            memcpy(&out[(r * cols + 3 * i + 1) * PL], &result[(r * cols + t + 2 * i) * PL], 2 * PL * sizeof(uint64_t));
        }
    free(cv_raw); free(ginv); free(g_ntt); free(chat); free(s2m); free(prod); free(result);
}

void so_gsw_negate(uint64_t *q_neg_ntt, const uint64_t *q_crtd, uint32_t nu2, uint32_t t_gsw) {   /* :2361-2378 */
    size_t m2 = (size_t)SO_N1 * t_gsw, per = SO_N1 * m2 * N;
    uint64_t *G2 = xalloc(per), *neg = xalloc((size_t)nu2 * per);
    so_build_gadget(G2, SO_N1, m2);
    for (size_t j = 0; j < nu2; j++)
        for (size_t i = 0; i < per; i++) {
            long val = (long)G2[i] - (long)q_crtd[j * per + i];
            if (val < 0) val += (long)Q;
            neg[j * per + i] = (uint64_t)val;
        }
    so_cpu_crt_to_ucompressed_and_ntt(q_neg_ntt, neg, (size_t)nu2 * SO_N1 * m2);
    free(G2); free(neg);
}

static size_t ceil_log2(size_t x) { size_t g = 0; while (((size_t)1 << g) < x) g++; return g; }

void so_spiral_expansion_shape(const so_params *prm, size_t *g, size_t *stopround) {   /* :2076-2085 */
    size_t num_expanded = (size_t)1 << prm->nu1, ell = prm->t_gsw;
    size_t num_bits_to_gen = ell * prm->nu2 + num_expanded;
    *g = ceil_log2(num_bits_to_gen);
    *stopround = ceil_log2(ell * prm->nu2);
    if (ell * prm->nu2 > num_expanded) *stopround = 0;
}

int so_spiral_answer(const so_params *prm, const uint64_t *query_cv, const uint64_t *W_exp_left,
                     const uint64_t *W_exp_right, const uint64_t *W_conv, const uint64_t *V_conv,
                     const uint64_t *Bdb, uint64_t *final_ct_raw, uint64_t *total_resp,
                     uint64_t *dbg_first_dim_raw) {
    const size_t CT = 2 * 2 * N, PL = 2 * N;
    size_t dim0 = (size_t)1 << prm->nu1, num_per = (size_t)1 << prm->nu2, fd = prm->nu2;
    size_t ell = prm->t_gsw, m2 = SO_N1 * ell;
    size_t g, stopround;
    so_spiral_expansion_shape(prm, &g, &stopround);
    size_t num_bits_to_gen = ell * fd + dim0;

    /* --- expansion (runConversionImproved :2159-2177) --- */
    uint64_t *round_cv = xalloc(((size_t)1 << g) * CT);
    memcpy(round_cv, query_cv, CT * sizeof(uint64_t));
    so_expand_improved(round_cv, g, prm->t_exp, W_exp_left, W_exp_right, prm->t_exp_right, ell * fd, stopround);
    uint64_t *cv_v = xalloc(num_bits_to_gen * CT);
    if (stopround != 0) {                                      /* reorderFromStopround :2027-2036 */
        for (size_t i = 0; i < dim0; i++) memcpy(&cv_v[i * CT], &round_cv[(2 * i) * CT], CT * sizeof(uint64_t));
        for (size_t i = 0; i < ell * fd; i++) memcpy(&cv_v[(dim0 + i) * CT], &round_cv[(2 * i + 1) * CT], CT * sizeof(uint64_t));
    } else {
        memcpy(cv_v, round_cv, num_bits_to_gen * CT * sizeof(uint64_t));
    }
    free(round_cv);

    /* --- ScalToMat x dim0 (:2232-2253) --- */
    size_t creg = SO_N1 * SO_N0 * PL;
    uint64_t *exp_cts = xalloc(dim0 * creg);
    for (size_t i = 0; i < dim0; i++) so_scal_to_mat(&exp_cts[i * creg], &cv_v[i * CT], W_conv, prm->t_conv);

    /* --- RegevToGSW x nu2 (:2311-2331) --- */
    size_t per_ntt = SO_N1 * m2 * PL, per_raw = SO_N1 * m2 * N;
    uint64_t *gQ_ntt = xalloc(fd * per_ntt), *gQ_crtd = xalloc(fd * per_raw);
    for (size_t i = 0; i < fd; i++) {
        size_t slot = fd - 1 - i;
        so_regev_to_gsw(&gQ_ntt[slot * per_ntt], &cv_v[(dim0 + i * ell) * CT], prm->t_conv, (uint32_t)ell, W_conv, V_conv);
        so_from_ntt(&gQ_crtd[slot * per_raw], &gQ_ntt[slot * per_ntt], SO_N1 * m2);
    }
    free(cv_v);

    /* --- negation + reorient_Q (process_crtd_query :2361-2386) --- */
    uint64_t *gQ_neg_ntt = xalloc(fd * per_ntt);
    so_gsw_negate(gQ_neg_ntt, gQ_crtd, (uint32_t)fd, (uint32_t)ell);
    uint64_t *gQ = xalloc(fd * per_ntt), *gQ_neg = xalloc(fd * per_ntt);
    for (size_t j = 0; j < fd; j++) {
        so_reorient_Q(&gQ[j * per_ntt], &gQ_ntt[j * per_ntt], (uint32_t)ell);
        so_reorient_Q(&gQ_neg[j * per_ntt], &gQ_neg_ntt[j * per_ntt], (uint32_t)ell);
    }
    free(gQ_ntt); free(gQ_crtd); free(gQ_neg_ntt);

    /* --- process_query_fast (:1584-1629) --- */
    uint64_t *reor = xalloc(dim0 * 2 * 4 * N);                 /* n1_padded = 4, r = 3 lane stays zero */
    so_reorient_ciphertexts(reor, exp_cts, dim0, 4);
    free(exp_cts);
    size_t ctw = SO_N1 * SO_N2;
    uint64_t *scratch = xalloc(num_per * ctw * PL), *cts = xalloc(num_per * ctw * N);
    so_multiply_query_by_database(scratch, reor, Bdb, dim0, num_per);
    so_ntt_inv_and_crt_lift(cts, scratch, num_per);
    free(reor); free(scratch);
    if (dbg_first_dim_raw) memcpy(dbg_first_dim_raw, cts, num_per * ctw * N * sizeof(uint64_t));
    size_t cur_dim = 0, np = num_per;
    while (np >= 2) {
        np /= 2;
        so_fold_one_further_dimension(cur_dim, np, gQ, gQ_neg, cts, (uint32_t)ell);
        cur_dim++;
    }
    free(gQ); free(gQ_neg);
    memcpy(final_ct_raw, cts, ctw * N * sizeof(uint64_t));
    free(cts);

    /* --- modulus switch (check_final :1441-1447) --- */
    uint64_t q_1 = 4 * prm->p_db, qp = so_arb_qprime(prm->qp_bits);
    so_get_rescaled(total_resp, final_ct_raw, SO_N2 * N, Q, qp);
    so_get_rescaled(&total_resp[SO_N2 * N], &final_ct_raw[SO_N2 * N], (SO_N1 - 1) * SO_N2 * N, Q, q_1);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * SpiralPack / SpiralStreamPack server path  (src/testing.cpp)
 * ---------------------------------------------------------------------------------------- */
void so_convert_db(uint64_t *db_buf, const uint64_t *db_ntt, size_t count, size_t dim0, size_t num_per) {   /* :316-340, pt 1x1 */
    for (size_t i = 0; i < count; i++) {
        size_t ii = i % num_per, j = i / num_per;
        const uint64_t *BB = &db_ntt[i * 2 * N];
        for (size_t z = 0; z < N; z++) db_buf[z * (num_per * dim0) + ii * dim0 + j] = BB[z] | (BB[N + z] << 32);
    }
}
void so_reorient_ciphertexts_dim1(uint64_t *out, const uint64_t *v, size_t dim0, size_t idx_factor) {   /* :342-362, ct 2x1 */
    for (size_t j = 0; j < dim0; j++) {
        const uint64_t *ct = &v[(j * idx_factor) * 2 * 2 * N];
        for (size_t r = 0; r < 2; r++)
            for (size_t z = 0; z < N; z++)
                out[z * (dim0 * 2) + j * 2 + r] = (ct[r * 2 * N + z] % P) | ((ct[r * 2 * N + N + z] % B_) << 32);
    }
}
void so_fast_multiply_dim1(uint64_t *out, const uint64_t *db, const uint64_t *vf, size_t dim0, size_t num_per) {   /* :537-592 */
    for (size_t z = 0; z < N; z++) {
        size_t a_base = z * dim0 * 2, b_idx = z * num_per * dim0;
        for (size_t i = 0; i < num_per; i++) {
            u128 s00 = 0, s01 = 0, s10 = 0, s11 = 0;
            for (size_t j = 0; j < dim0; j++) {
                uint64_t b = db[b_idx++], a0 = vf[a_base + j * 2], a1 = vf[a_base + j * 2 + 1];
                uint32_t b_lo = (uint32_t)b, b_hi = (uint32_t)(b >> 32);
                s00 += (uint64_t)(uint32_t)a0 * b_lo;  s01 += (uint64_t)(uint32_t)a1 * b_lo;
                s10 += (uint64_t)(uint32_t)(a0 >> 32) * b_hi;  s11 += (uint64_t)(uint32_t)(a1 >> 32) * b_hi;
            }
            uint64_t *o = &out[i * 2 * 2 * N];
            o[z] = (uint64_t)(s00 % P);          o[2 * N + z] = (uint64_t)(s01 % P);
            o[N + z] = (uint64_t)(s10 % B_);     o[2 * N + N + z] = (uint64_t)(s11 % B_);
        }
    }
}

void so_fold_ciphertexts_dim1(uint64_t *v_cts, size_t count, const uint64_t *v_folding,
                              const uint64_t *v_folding_neg, uint32_t ell) {   /* :596-624 */
    const size_t PL = 2 * N;
    size_t further_dims = ceil_log2(count), mx = 2 * (size_t)ell, gsw = 2 * mx * PL;
    uint64_t *ginv = xalloc(mx * N), *ginv_ntt = xalloc(mx * PL), *prod = xalloc(2 * PL), *sum = xalloc(2 * PL);
    size_t num_per = count;
    for (size_t cur = 0; cur < further_dims; cur++) {
        num_per /= 2;
        for (size_t i = 0; i < num_per; i++) {
            so_gadget_invert(ginv, &v_cts[i * 2 * N], mx, 2, 1);
            so_to_ntt(ginv_ntt, ginv, mx);
            so_multiply(prod, &v_folding_neg[(further_dims - 1 - cur) * gsw], ginv_ntt, 2, mx, 1);
            so_gadget_invert(ginv, &v_cts[(num_per + i) * 2 * N], mx, 2, 1);
            so_to_ntt(ginv_ntt, ginv, mx);
            so_multiply(sum, &v_folding[(further_dims - 1 - cur) * gsw], ginv_ntt, 2, mx, 1);
            so_add(sum, sum, prod, 2);
            so_from_ntt(&v_cts[i * 2 * N], sum, 2);
        }
    }
    free(ginv); free(ginv_ntt); free(prod); free(sum);
}

void so_regev_to_simple_gsw(uint64_t *v_gsw, const uint64_t *v_inp, const uint64_t *V, uint32_t t_conv,
                            uint32_t ell, uint32_t further_dims, size_t idx_factor, size_t idx_offset) {   /* :108-140 */
    const size_t PL = 2 * N, CT = 2 * PL;
    size_t mc2 = 2 * (size_t)t_conv, cols = 2 * (size_t)ell;
    uint64_t *raw = xalloc(2 * N), *ginv = xalloc(mc2 * N), *ginv_ntt = xalloc(mc2 * PL), *tmp = xalloc(2 * PL);
    for (size_t i = 0; i < further_dims; i++) {
        uint64_t *ct = &v_gsw[i * 2 * cols * PL];
        for (size_t j = 0; j < ell; j++) {
            const uint64_t *c_inp = &v_inp[(idx_factor * (i * ell + j) + idx_offset) * CT];
            for (size_t r = 0; r < 2; r++) memcpy(&ct[(r * cols + 2 * j + 1) * PL], &c_inp[r * PL], PL * sizeof(uint64_t));
            so_from_ntt(raw, c_inp, 2);
            so_gadget_invert(ginv, raw, mc2, 2, 1);
            so_to_ntt(ginv_ntt, ginv, mc2);
            so_multiply(tmp, V, ginv_ntt, 2, mc2, 1);
            for (size_t r = 0; r < 2; r++) memcpy(&ct[(r * cols + 2 * j) * PL], &tmp[r * PL], PL * sizeof(uint64_t));
        }
    }
    free(raw); free(ginv); free(ginv_ntt); free(tmp);
}

void so_simple_gsw_negate(uint64_t *neg, const uint64_t *v_folding, uint32_t further_dims, uint32_t ell) {   /* :1027-1032 */
    size_t cols = 2 * (size_t)ell, np = 2 * cols;
    uint64_t *gadget = xalloc(np * N), *gadget_ntt = xalloc(np * 2 * N);
    uint64_t *raw = xalloc(np * N), *inv = xalloc(np * N), *inv_ntt = xalloc(np * 2 * N);
    so_build_gadget(gadget, 2, cols);
    so_to_ntt(gadget_ntt, gadget, np);
    for (size_t i = 0; i < further_dims; i++) {
        so_from_ntt(raw, &v_folding[i * np * 2 * N], np);
        so_invert(inv, raw, np);
        so_to_ntt(inv_ntt, inv, np);
        so_add(&neg[i * np * 2 * N], gadget_ntt, inv_ntt, np);
    }
    free(gadget); free(gadget_ntt); free(raw); free(inv); free(inv_ntt);
}

void so_pack(uint64_t *result, uint32_t out_n, uint32_t t_conv, const uint64_t *v_ct, const uint64_t *v_W) {   /* :198-241 */
    const size_t PL = 2 * N;
    size_t rows = out_n + 1;
    uint64_t *v_int = xalloc(rows * PL), *ginv = xalloc((size_t)t_conv * N), *ginv_ntt = xalloc((size_t)t_conv * PL);
    uint64_t *prod = xalloc(rows * PL), *ct2_ntt = xalloc(PL);
    for (size_t c = 0; c < out_n; c++) {
        memset(v_int, 0, rows * PL * sizeof(uint64_t));
        for (size_t r = 0; r < out_n; r++) {
            const uint64_t *W = &v_W[r * rows * t_conv * PL];
            const uint64_t *ct = &v_ct[(r * out_n + c) * 2 * N];
            so_to_ntt(ct2_ntt, &ct[N], 1);
            so_gadget_invert(ginv, ct, t_conv, 1, 1);
            so_to_ntt(ginv_ntt, ginv, t_conv);
            so_multiply(prod, W, ginv_ntt, rows, t_conv, 1);
            so_add(&v_int[(1 + r) * PL], &v_int[(1 + r) * PL], ct2_ntt, 1);   /* add_into(v_int, v_int, ct_2_ntt, 1 + r, 0) */
            so_add(v_int, v_int, prod, rows);
        }
        for (size_t r = 0; r < rows; r++) memcpy(&result[(r * out_n + c) * PL], &v_int[r * PL], PL * sizeof(uint64_t));
    }
    free(v_int); free(ginv); free(ginv_ntt); free(prod); free(ct2_ntt);
}

void so_pack_expansion_shape(const so_params *prm, size_t *g, size_t *stopround) {   /* :795-798 */
    size_t ell = prm->t_gsw;
    *g = ceil_log2(ell * prm->nu2 + ((size_t)1 << prm->nu1));
    *stopround = ceil_log2(ell * prm->nu2);
}

int so_pack_answer(const so_params *prm, int do_expansion, const uint64_t *query_cv,
                   const uint64_t *W_exp_left, const uint64_t *W_exp_right, const uint64_t *V,
                   const uint64_t *v_firstdim, const uint64_t *v_folding_direct, const uint64_t *v_W,
                   const uint64_t *db_planes, uint64_t *total_resp, uint64_t *result_cts) {
    const size_t PL = 2 * N, CT = 2 * PL;
    size_t dim0 = (size_t)1 << prm->nu1, num_per = (size_t)1 << prm->nu2, fd = prm->nu2;
    size_t ell = prm->t_gsw, out_n = prm->out_n, trials = out_n * out_n;
    size_t gsw = 2 * 2 * ell * PL;
    uint64_t *reor = xalloc(dim0 * 2 * N), *v_folding = xalloc(fd * gsw), *v_folding_neg = xalloc(fd * gsw);
    if (do_expansion) {                                        /* :1007-1025 */
        size_t g, stopround;
        so_pack_expansion_shape(prm, &g, &stopround);
        uint64_t *v = xalloc(((size_t)1 << g) * CT);
        memcpy(v, query_cv, CT * sizeof(uint64_t));
        so_expand_improved(v, g, prm->t_exp, W_exp_left, W_exp_right, prm->t_exp_right, ell * fd, stopround);
        so_reorient_ciphertexts_dim1(reor, v, dim0, 2);
        so_regev_to_simple_gsw(v_folding, v, V, prm->t_conv, (uint32_t)ell, (uint32_t)fd, 2, 1);
        free(v);
    } else {                                                   /* :962-985 */
        so_reorient_ciphertexts_dim1(reor, v_firstdim, dim0, 1);
        memcpy(v_folding, v_folding_direct, fd * gsw * sizeof(uint64_t));
    }
    so_simple_gsw_negate(v_folding_neg, v_folding, (uint32_t)fd, (uint32_t)ell);

    uint64_t *v_out = xalloc(num_per * CT), *v_out_raw = xalloc(num_per * 2 * N), *v_res = xalloc(trials * 2 * N);
    for (size_t trial = 0; trial < trials; trial++) {          /* :1045-1062 */
        so_fast_multiply_dim1(v_out, &db_planes[trial * N * dim0 * num_per], reor, dim0, num_per);
        so_from_ntt(v_out_raw, v_out, num_per * 2);
        so_fold_ciphertexts_dim1(v_out_raw, num_per, v_folding, v_folding_neg, (uint32_t)ell);
        memcpy(&v_res[trial * 2 * N], v_out_raw, 2 * N * sizeof(uint64_t));
    }
    if (result_cts) memcpy(result_cts, v_res, trials * 2 * N * sizeof(uint64_t));

    size_t rows = out_n + 1;                                    /* :1066-1081 */
    uint64_t *packed = xalloc(rows * out_n * PL), *ct_inp = xalloc(rows * out_n * N);
    so_pack(packed, (uint32_t)out_n, prm->t_conv, v_res, v_W);
    so_from_ntt(ct_inp, packed, rows * out_n);
    so_get_rescaled(total_resp, ct_inp, out_n * N, Q, so_arb_qprime(prm->qp_bits));
    so_get_rescaled(&total_resp[out_n * N], &ct_inp[out_n * N], (rows - 1) * out_n * N, Q, 4 * prm->p_db);
    free(reor); free(v_folding); free(v_folding_neg); free(v_out); free(v_out_raw); free(v_res);
    free(packed); free(ct_inp);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * digests + deterministic inputs
 * ---------------------------------------------------------------------------------------- */
uint64_t so_fnv1a64(const uint64_t *w, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) {
        uint64_t v = w[i];
        for (int b = 0; b < 8; b++) { h ^= (v & 0xff); h *= 1099511628211ull; v >>= 8; }
    }
    return h;
}
uint64_t so_fnv1a64_ntt(const uint64_t *w, size_t npolys) {
    uint64_t h = 1469598103934665603ull;
    for (size_t p = 0; p < npolys; p++)
        for (int n = 0; n < 2; n++)
            for (size_t z = 0; z < N; z++) {
                uint64_t v = w[p * 2 * N + n * N + z] % (n == 0 ? P : B_);
                for (int b = 0; b < 8; b++) { h ^= (v & 0xff); h *= 1099511628211ull; v >>= 8; }
            }
    return h;
}
uint64_t so_rng_next(so_rng *r) {
    uint64_t z = (r->s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
void so_fill_uniform_mod(uint64_t *out, size_t n, uint64_t mod, so_rng *r) {
    for (size_t i = 0; i < n; i++) out[i] = so_rng_next(r) % mod;
}
void so_fill_uniform_raw(uint64_t *out, size_t n, so_rng *r) { so_fill_uniform_mod(out, n, Q, r); }
void so_fill_uniform_ntt(uint64_t *out, size_t npolys, so_rng *r) {
    for (size_t p = 0; p < npolys; p++) {
        so_fill_uniform_mod(&out[p * 2 * N], N, P, r);
        so_fill_uniform_mod(&out[p * 2 * N + N], N, B_, r);
    }
}
