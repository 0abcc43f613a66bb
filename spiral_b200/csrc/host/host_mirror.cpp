// host_mirror.cpp - C++ host side above the C-ABI: definitions of the reference's own server
// functions (same names, same signatures, hence the same mangled symbols) that forward to
// libspiral_b200.so.  Linked into an executable AHEAD of the reference library, the dynamic linker
// routes the reference harness's calls (process_query_fast, runConversionImproved, testHighRate ...)
// to these definitions: the reference's client, parameters and "Is correct?" check drive the CUDA
// path unchanged.  No arithmetic happens here.
//
// Build (oracle/Makefile, target `driver`): one executable per parameter set because the reference
// fixes TGSW etc. at compile time; this file receives the same -D values.
//
// SB200_PARITY=1 in the environment: every mirrored call ALSO runs the reference's own function
// (dlsym(RTLD_NEXT, ...)) on a copy of the inputs and aborts on the first differing byte
// (NTT-domain buffers are compared modulo the prime).
#include <dlfcn.h>
#include <chrono>
#include <cstdio>
#include "../../../include/spiral_b200.h"
#include "ref_abi.h"

#if !defined(TGSW) || !defined(TCONV) || !defined(TEXP) || !defined(TEXPRIGHT) || !defined(QPBITS) || !defined(PVALUE)
#error "compile with the reference's own -DTGSW= -DTCONV= -DTEXP= -DTEXPRIGHT= -DQPBITS= -DPVALUE= values"
#endif

static const size_t N = 2048, PL = 4096;
static const uint64_t P = 268369921ull, Bq = 249561089ull;

// resident mode (default when not checking parity): the leaves of process_query_fast keep their intermediates in HBM between
// calls (sb200_resident_*); SB200_RESIDENT=0 restores one H2D + D2H per leaf
static bool resident_mode() { static int v = -1; if (v < 0) { const char *e = getenv("SB200_RESIDENT"); v = (e && *e == '0') ? 0 : 1; } return v == 1; }
static bool parity_mode() { static int v = -1; if (v < 0) { const char *e = getenv("SB200_PARITY"); v = (e && *e == '1') ? 1 : 0; } return v == 1; }
static void die(const char *what) { fprintf(stderr, "[spiral_b200] %s: %s\n", what, sb200_last_error()); abort(); }
#define OKAY(call) do { if ((call) != 0) die(#call); } while (0)
template <typename F> static F next_sym(const char *mangled) {
    void *p = dlsym(RTLD_NEXT, mangled);
    if (!p) { fprintf(stderr, "[spiral_b200] reference symbol %s not found\n", mangled); abort(); }
    return (F)p;
}
static void cmp_ntt(const char *what, const uint64_t *a, const uint64_t *b, size_t npolys) {
    for (size_t i = 0; i < npolys * PL; i++) {
        uint64_t q = ((i / N) & 1) ? Bq : P;
        if (a[i] % q != b[i] % q) { fprintf(stderr, "[spiral_b200] PARITY FAIL %s word %zu: %llu vs %llu\n", what, i, (unsigned long long)a[i], (unsigned long long)b[i]); abort(); }
    }
    fprintf(stderr, "[spiral_b200] parity ok: %s (%zu polys, mod q)\n", what, npolys);
}
static void cmp_packed(const char *what, const uint64_t *a, const uint64_t *b, size_t words) {
    for (size_t i = 0; i < words; i++)
        if ((a[i] & 0xffffffffull) % P != (b[i] & 0xffffffffull) % P || (a[i] >> 32) % Bq != (b[i] >> 32) % Bq) {
            fprintf(stderr, "[spiral_b200] PARITY FAIL %s word %zu\n", what, i); abort(); }
    fprintf(stderr, "[spiral_b200] parity ok: %s (%zu packed words)\n", what, words);
}
static void cmp_raw(const char *what, const uint64_t *a, const uint64_t *b, size_t words) {
    if (memcmp(a, b, words * 8) != 0) { fprintf(stderr, "[spiral_b200] PARITY FAIL %s (raw, %zu words)\n", what, words); abort(); }
    fprintf(stderr, "[spiral_b200] parity ok: %s (%zu raw words, exact)\n", what, words);
}

// ---- reference globals the mirror reads (defined by the reference library) ---------------------
extern uint64_t *B;                       // src/spiral.cpp:1017
extern size_t num_expansions, further_dims;
extern bool random_data;                  // src/spiral.cpp:18   (--random-data: the implicit database)
extern size_t dummyWorkingSet;            // src/util.cpp:3      (z-slices B holds in that mode)

static sb200_server *g_srv = nullptr;     // resident database shard (whole database, world = 1)

// ---- parity mode, tier 3: the RESIDENT server answers the harness's own query --------------------------------------
// The leaves below see everything a real server would receive: the packed query ciphertext (first argument of the
// expansion), the expansion keys, the conversion keys W and V.  In parity mode they are captured, the resident server
// (sb200_server_answer: expansion, conversion, scan, CMux folds, modulus switch - the path bench.py times) answers the
// same query against the same database, and its response is memcmp'd with the harness's own total_resp at the two
// getRescaled calls of check_final (src/spiral.cpp:1441-1447) / testHighRate (src/testing.cpp:1074-1081).
struct Tier3Capture {
    std::vector<uint64_t> query, W_left, W_right, W_conv, V, vW, v_firstdim, v_folding;
    int expansions = 0;
    bool answered = false, failed = false;
    std::vector<uint64_t> resp, result_cts;
};
static Tier3Capture g_t3;
static void tier3_note(const char *what) { fprintf(stderr, "[spiral_b200] tier-3: %s\n", what); }
static void tier3_spiral_answer() {
    if (g_t3.answered || g_t3.failed) return;
    if (!g_srv || g_t3.expansions != 1 || g_t3.query.empty() || g_t3.W_conv.empty() || g_t3.V.empty()) {
        g_t3.failed = true; tier3_note("resident cross-check skipped (query not a single packed ciphertext)"); return;
    }
    OKAY(sb200_server_set_public_params(g_srv, g_t3.W_left.data(), g_t3.W_right.data(), g_t3.W_conv.data(), g_t3.V.data()));
    g_t3.resp.assign(6 * N, 0);
    OKAY(sb200_server_answer(g_srv, g_t3.query.data(), g_t3.resp.data(), nullptr));
    g_t3.answered = true;
}
// `got` = the harness's modulus-switched rows; the resident response holds row 0 in words [0, 2N), rows 1.. after it
static void tier3_compare(const char *what, const uint64_t *got, size_t offset_words, size_t words) {
    if (!g_t3.answered) return;
    if (memcmp(g_t3.resp.data() + offset_words, got, words * 8) != 0) {
        fprintf(stderr, "[spiral_b200] PARITY FAIL tier-3 resident server, %s (%zu words)\n", what, words); abort();
    }
    fprintf(stderr, "[spiral_b200] parity ok: tier-3 resident server %s (%zu raw words, exact)\n", what, words);
}

// ---- load_db (src/spiral.cpp:1028): the reference generates B + its bookkeeping globals, then the
// database is made resident in HBM once.
void load_db() {
    next_sym<void (*)()>("_Z7load_dbv")();
    sb200_params prm = {};
    prm.nu1 = (uint32_t)num_expansions; prm.nu2 = (uint32_t)further_dims;
    prm.t_gsw = TGSW; prm.t_conv = TCONV; prm.t_exp = TEXP; prm.t_exp_right = TEXPRIGHT;
    prm.qp_bits = QPBITS; prm.out_n = 2; prm.p_db = PVALUE;
    OKAY(sb200_server_create(&g_srv, &prm, 0, 0, 1));
    if (random_data) {        // implicit database: B holds dummyWorkingSet slices, the scan reads slice z mod dummyWorkingSet (:647)
        OKAY(sb200_server_load_db_implicit(g_srv, B, dummyWorkingSet));
        fprintf(stderr, "[spiral_b200] implicit database resident on the GPU (%zu of 2048 slices, %zu MiB)\n", dummyWorkingSet,
                (sb200_db_words(prm.nu1, prm.nu2) / 2048 * dummyWorkingSet * 8) >> 20);
        return;
    }
    OKAY(sb200_server_load_db_reference(g_srv, B));
    fprintf(stderr, "[spiral_b200] database resident on the GPU (%zu MiB)\n", (sb200_db_words(prm.nu1, prm.nu2) * 8) >> 20);
}

// ---- reorientCiphertexts (src/spiral.cpp:410)
void reorientCiphertexts(uint64_t *out, const uint64_t *inp, size_t dim0, size_t n1_padded) {
    if (g_srv && resident_mode() && !parity_mode() && n1_padded == 4 && sb200_resident_reorientCiphertexts(g_srv, out, inp, dim0) == 0) return;
    OKAY(sb200_reorientCiphertexts(out, inp, dim0, n1_padded));
    if (parity_mode()) {
        std::vector<uint64_t> ref(dim0 * 2 * n1_padded * N, 0);
        next_sym<void (*)(uint64_t *, const uint64_t *, size_t, size_t)>("_Z19reorientCiphertextsPmPKmmm")(ref.data(), inp, dim0, n1_padded);
        cmp_packed("reorientCiphertexts", out, ref.data(), ref.size());
    }
}

// ---- multiplyQueryByDatabase (src/spiral.cpp:628): scan of the RESIDENT database when `database`
// is the buffer registered by load_db, otherwise the stateless call
void multiplyQueryByDatabase(uint64_t *output, const uint64_t *reorientedCiphertexts, const uint64_t *database, size_t dim0, size_t num_per) {
    // resident: the reoriented query is already in HBM (left there by reorientCiphertexts), the result stays there for the lift
    if (g_srv && database == B && resident_mode() && !parity_mode() && sb200_resident_multiplyQueryByDatabase(g_srv, output, reorientedCiphertexts) == 0) return;
    if (g_srv && database == B) OKAY(sb200_server_scan_host(g_srv, reorientedCiphertexts, output));
    else OKAY(sb200_multiplyQueryByDatabase(output, reorientedCiphertexts, database, dim0, num_per));
    if (parity_mode()) {
        std::vector<uint64_t> ref(num_per * 6 * PL, 0);
        next_sym<void (*)(uint64_t *, const uint64_t *, const uint64_t *, size_t, size_t)>("_Z23multiplyQueryByDatabasePmPKmS1_mm")(ref.data(), reorientedCiphertexts, database, dim0, num_per);
        cmp_ntt("multiplyQueryByDatabase", output, ref.data(), num_per * 6);
    }
}

// ---- nttInvAndCrtLiftCiphertexts (src/spiral.cpp:437): locals passed by value
void nttInvAndCrtLiftCiphertexts(size_t num_per, FurtherDimsLocals locals) {
    if (g_srv && resident_mode() && !parity_mode() && sb200_resident_nttInvAndCrtLiftCiphertexts(g_srv, locals.cts, locals.scratch_cts1) == 0) return;
    std::vector<uint64_t> ref_in;
    if (parity_mode()) ref_in.assign(locals.scratch_cts1, locals.scratch_cts1 + num_per * 6 * PL);
    OKAY(sb200_nttInvAndCrtLiftCiphertexts(locals.cts, locals.scratch_cts1, num_per));
    if (parity_mode()) {
        std::vector<uint64_t> got(locals.cts, locals.cts + num_per * 6 * N);
        memcpy(locals.scratch_cts1, ref_in.data(), ref_in.size() * 8);
        next_sym<void (*)(size_t, FurtherDimsLocals)>("_Z27nttInvAndCrtLiftCiphertextsm17FurtherDimsLocals")(num_per, locals);
        cmp_raw("nttInvAndCrtLiftCiphertexts", got.data(), locals.cts, got.size());
    }
}

// ---- foldOneFurtherDimension (src/spiral.cpp:1349)
void foldOneFurtherDimension(size_t cur_dim, size_t num_per, const uint64_t *query_ct, const uint64_t *query_ct_neg, FurtherDimsLocals locals) {
    const size_t q_stride = (size_t)3 * 3 * TGSW * 2 * N;       // words between the dimensions' reoriented GSW ciphertexts (:2385-2386)
    if (g_srv && resident_mode() && !parity_mode() &&
        sb200_resident_foldOneFurtherDimension(g_srv, num_per, query_ct + cur_dim * q_stride, query_ct_neg + cur_dim * q_stride, locals.cts) == 0) return;
    std::vector<uint64_t> before;
    if (parity_mode()) before.assign(locals.cts, locals.cts + 2 * num_per * 6 * N);
    OKAY(sb200_foldOneFurtherDimension(cur_dim, num_per, query_ct, query_ct_neg, locals.cts, TGSW));
    if (parity_mode()) {
        std::vector<uint64_t> got(locals.cts, locals.cts + num_per * 6 * N);
        memcpy(locals.cts, before.data(), before.size() * 8);
        next_sym<void (*)(size_t, size_t, const uint64_t *, const uint64_t *, FurtherDimsLocals)>("_Z23foldOneFurtherDimensionmmPKmS0_17FurtherDimsLocals")(cur_dim, num_per, query_ct, query_ct_neg, locals);
        cmp_raw("foldOneFurtherDimension", got.data(), locals.cts, got.size());
    }
}

// ---- reorient_Q (src/spiral.cpp:388): the host relayout stays the reference's; in resident mode the GSW ciphertext it was handed
// (n1 x m2, NTT form) also goes to HBM, registered under the output pointer foldOneFurtherDimension will pass back
void reorient_Q(uint64_t *out, const uint64_t *inp) {
    next_sym<void (*)(uint64_t *, const uint64_t *)>("_Z10reorient_QPmPKm")(out, inp);
    if (g_srv && resident_mode() && !parity_mode()) OKAY(sb200_resident_reorient_Q(g_srv, out, inp));
}

// ---- expandImproved (src/spiral.cpp:1664): returns its own elapsed microseconds (appendix C.1)
static std::vector<uint64_t> flatten(const std::vector<MatPoly> &v, size_t count) {
    std::vector<uint64_t> out;
    for (size_t i = 0; i < count; i++) out.insert(out.end(), v[i].data, v[i].data + v[i].words());
    return out;
}
double expandImproved(std::vector<MatPoly> &cv_v, size_t g, size_t m_exp, const std::vector<MatPoly> &W_left_v,
                      const std::vector<MatPoly> &W_right_v, size_t max_bits_to_gen_right, size_t stopround) {
    auto t0 = std::chrono::high_resolution_clock::now();
    const size_t ncts = (size_t)1 << g, n_right = stopround > 0 ? stopround + 1 : g;
    std::vector<uint64_t> cv = flatten(cv_v, ncts), Wl = flatten(W_left_v, g), Wr = flatten(W_right_v, n_right);
    std::vector<MatPoly> ref_cv;
    if (parity_mode()) for (size_t i = 0; i < ncts; i++) { MatPoly c(2, 1); memcpy(c.data, cv_v[i].data, 2 * PL * 8); ref_cv.push_back(c); }
    if (parity_mode() && g_t3.expansions++ == 0) { g_t3.query.assign(cv.begin(), cv.begin() + 2 * PL); g_t3.W_left = Wl; g_t3.W_right = Wr; }
    const uint32_t t_right = (uint32_t)W_right_v[0].cols;
    OKAY(sb200_expandImproved(cv.data(), g, (uint32_t)m_exp, Wl.data(), Wr.data(), t_right, max_bits_to_gen_right, stopround));
    for (size_t i = 0; i < ncts; i++) memcpy(cv_v[i].data, &cv[i * 2 * PL], 2 * PL * 8);
    double us = (double)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::high_resolution_clock::now() - t0).count();
    if (parity_mode()) {
        next_sym<double (*)(std::vector<MatPoly> &, size_t, size_t, const std::vector<MatPoly> &, const std::vector<MatPoly> &, size_t, size_t)>(
            "_Z14expandImprovedRSt6vectorI7MatPolySaIS0_EEmmRKS2_S5_mm")(ref_cv, g, m_exp, W_left_v, W_right_v, max_bits_to_gen_right, stopround);
        std::vector<uint64_t> ref = flatten(ref_cv, ncts);
        cmp_ntt("expandImproved", cv.data(), ref.data(), ncts * 2);
    }
    return us;
}

// ---- regevToGSW (src/spiral.cpp:1985)
void regevToGSW(size_t m_conv, size_t t, MatPoly &out, const std::vector<MatPoly> &cv_v, size_t cv_v_offset, const MatPoly &W, const MatPoly &V) {
    std::vector<uint64_t> cv;
    for (size_t i = 0; i < t; i++) cv.insert(cv.end(), cv_v[cv_v_offset + i].data, cv_v[cv_v_offset + i].data + 2 * PL);
    std::vector<uint64_t> res(3 * 3 * t * PL);
    if (parity_mode() && g_t3.V.empty()) { g_t3.W_conv.assign(W.data, W.data + W.words()); g_t3.V.assign(V.data, V.data + V.words()); }
    OKAY(sb200_regevToGSW(res.data(), cv.data(), (uint32_t)m_conv, (uint32_t)t, W.data, V.data));
    if (parity_mode()) {
        MatPoly ref(3, 3 * t);
        next_sym<void (*)(size_t, size_t, MatPoly &, const std::vector<MatPoly> &, size_t, const MatPoly &, const MatPoly &)>(
            "_Z10regevToGSWmmR7MatPolyRKSt6vectorIS_SaIS_EEmRKS_S7_")(m_conv, t, ref, cv_v, cv_v_offset, W, V);
        cmp_ntt("regevToGSW", res.data(), ref.data, 3 * 3 * t);
    }
    // the reference assigns `out = result_permuted` (a fresh calloc, include/poly.h:48-58)
    out.rows = 3; out.cols = 3 * t; out.isNTT = true;
    out.data = (uint64_t *)calloc(res.size(), sizeof(uint64_t));
    memcpy(out.data, res.data(), res.size() * 8);
}

// ---- scalToMat (src/spiral.cpp:1850, the 14-argument form runConversionImproved calls at :2231; the scratch
// matrices are the reference's own working storage and are left untouched)
void scalToMat(size_t m_conv, MatPoly &out_reg, const MatPoly &cv, const MatPoly &W, MatPoly &cv_0, MatPoly &cv_1, MatPoly &cv_ntti,
               MatPoly &square_cv, MatPoly &ginv_c, MatPoly &ginv_c_nttd, MatPoly &prod_W_ginv, MatPoly &padded_cv_1,
               MatPoly &ginv_c_raw, MatPoly &ginv_c_raw_nttd) {
    OKAY(sb200_scalToMat(out_reg.data, cv.data, W.data, (uint32_t)m_conv));
    if (parity_mode()) {
        static int shown = 0;
        MatPoly ref(3, 2);
        next_sym<void (*)(size_t, MatPoly &, const MatPoly &, const MatPoly &, MatPoly &, MatPoly &, MatPoly &, MatPoly &, MatPoly &, MatPoly &,
                          MatPoly &, MatPoly &, MatPoly &, MatPoly &)>("_Z9scalToMatmR7MatPolyRKS_S2_S0_S0_S0_S0_S0_S0_S0_S0_S0_S0_")(
            m_conv, ref, cv, W, cv_0, cv_1, cv_ntti, square_cv, ginv_c, ginv_c_nttd, prod_W_ginv, padded_cv_1, ginv_c_raw, ginv_c_raw_nttd);
        for (size_t i = 0; i < 6 * PL; i++) {
            const uint64_t q = ((i / N) & 1) ? Bq : P;
            if (out_reg.data[i] % q != ref.data[i] % q) { fprintf(stderr, "[spiral_b200] PARITY FAIL scalToMat word %zu\n", i); abort(); }
        }
        if (!shown++) fprintf(stderr, "[spiral_b200] parity ok: scalToMat (6 polys, mod q; every call is checked)\n");
        free(ref.data);
    }
}

// ---- modswitch (src/spiral.cpp:40): the QPBITS-bit response packer
void modswitch(uint64_t *out, const uint64_t *inp) {
    OKAY(sb200_modswitch(out, inp, QPBITS));
    if (parity_mode()) {
        const size_t words = sb200_packed_words(3 * 2 * N, QPBITS);
        std::vector<uint64_t> ref(words + 1, 0);
        next_sym<void (*)(uint64_t *, const uint64_t *)>("_Z9modswitchPmPKm")(ref.data(), inp);
        cmp_raw("modswitch", out, ref.data(), words);
    }
}

static bool g_pack_tier3();              // Pack variants: the resident pack server has answered (below)
static bool g_pack_mode();               // Pack variants: a resident pack server exists
// ---- getRescaled (src/poly.cpp:593): modulus switch of the response
MatPoly getRescaled(const MatPoly &a, uint64_t inp_mod, uint64_t out_mod) {
    // the reference's `MatPoly b = a;` is a SHALLOW copy (implicit copy constructor): it rescales a's storage in place and
    // returns a matrix sharing it - reproduced, since callers may rely on either name
    const size_t n = a.rows * a.cols * N;
    std::vector<uint64_t> in(a.data, a.data + n);
    MatPoly b = a;
    OKAY(sb200_getRescaled(b.data, in.data(), n, inp_mod, out_mod));
    if (parity_mode()) {
        MatPoly a2(a.rows, a.cols, false);
        memcpy(a2.data, in.data(), n * 8);
        MatPoly ref = next_sym<MatPoly (*)(const MatPoly &, uint64_t, uint64_t)>("_Z11getRescaledRK7MatPolymm")(a2, inp_mod, out_mod);
        cmp_raw("getRescaled", b.data, ref.data, n);
        free(a2.data);
        // the two calls of check_final / testHighRate: first row (1 x cols) -> arb_qprime, rest rows -> 4 * p_db
        if (g_pack_tier3()) {
            const size_t cols = a.cols;
            if (a.rows == 1) tier3_compare("response row 0", b.data, 0, n);
            else tier3_compare("response rows 1..", b.data, cols * N, n);
        } else if (!g_pack_mode() && a.cols == 2 && (a.rows == 1 || a.rows == 2)) {
            if (a.rows == 1) { tier3_spiral_answer(); tier3_compare("response row 0", b.data, 0, n); }
            else tier3_compare("response rows 1-2", b.data, 2 * N, n);
        }
    }
    return b;
}

// =================================================================================================
// SpiralPack / SpiralStreamPack leaves (src/testing.cpp), driven by testHighRate (`--high-rate`)
// =================================================================================================
#ifndef OUTN
#define OUTN 4
#endif
static sb200_pack_server *g_pack = nullptr;                 // resident planes, in convertDb call order
static std::vector<const uint64_t *> g_plane_bufs;          // the db_buf pointers convertDb returned

// ---- convertDb (src/testing.cpp:316): relayout on the GPU; the plane also becomes resident in HBM
uint64_t *convertDb(const std::vector<MatPoly> &db, size_t dim0, size_t num_per) {
    const size_t count = db.size();
    std::vector<uint64_t> flat = flatten(db, count);
    uint64_t *buf = (uint64_t *)malloc(count * N * sizeof(uint64_t));
    OKAY(sb200_convertDb(buf, flat.data(), count, dim0, num_per));
    if (parity_mode()) {
        uint64_t *ref = next_sym<uint64_t *(*)(const std::vector<MatPoly> &, size_t, size_t)>("_Z9convertDbRKSt6vectorI7MatPolySaIS0_EEmm")(db, dim0, num_per);
        cmp_packed("convertDb", buf, ref, count * N);
        free(ref);
    }
    if (!g_pack) {
        sb200_params prm = {};
        prm.nu1 = (uint32_t)num_expansions; prm.nu2 = (uint32_t)further_dims;
        prm.t_gsw = TGSW; prm.t_conv = TCONV; prm.t_exp = TEXP; prm.t_exp_right = TEXPRIGHT;
        prm.qp_bits = QPBITS; prm.out_n = OUTN; prm.p_db = PVALUE;
        if (((size_t)1 << prm.nu1) == dim0 && ((size_t)1 << prm.nu2) == num_per) OKAY(sb200_pack_server_create(&g_pack, &prm, 0));
    }
    if (g_pack && g_plane_bufs.size() < (size_t)OUTN * OUTN) {
        OKAY(sb200_pack_server_load_plane_reference(g_pack, g_plane_bufs.size(), buf));
        g_plane_bufs.push_back(buf);
        fprintf(stderr, "[spiral_b200] database plane %zu resident on the GPU (%zu MiB)\n", g_plane_bufs.size() - 1, (count * N * 8) >> 20);
    }
    return buf;
}

// ---- coefficientExpansion (src/testing.cpp:40): same body as expandImproved, no return value
void coefficientExpansion(std::vector<MatPoly> &cv_v, size_t g, size_t m_exp, const std::vector<MatPoly> &W_left_v,
                          const std::vector<MatPoly> &W_right_v, size_t max_bits_to_gen_right, size_t stopround) {
    const size_t ncts = (size_t)1 << g;
    size_t n_right = stopround > 0 ? stopround + 1 : g;
    if (n_right > W_right_v.size()) n_right = W_right_v.size();
    std::vector<uint64_t> cv = flatten(cv_v, ncts), Wl = flatten(W_left_v, g), Wr = flatten(W_right_v, n_right);
    std::vector<MatPoly> ref_cv;
    if (parity_mode()) for (size_t i = 0; i < ncts; i++) { MatPoly c(2, 1); memcpy(c.data, cv_v[i].data, 2 * PL * 8); ref_cv.push_back(c); }
    if (parity_mode() && g_t3.expansions++ == 0) { g_t3.query.assign(cv.begin(), cv.begin() + 2 * PL); g_t3.W_left = Wl; g_t3.W_right = Wr; }
    OKAY(sb200_expandImproved(cv.data(), g, (uint32_t)m_exp, Wl.data(), Wr.data(), (uint32_t)W_right_v[0].cols, max_bits_to_gen_right, stopround));
    for (size_t i = 0; i < ncts; i++) memcpy(cv_v[i].data, &cv[i * 2 * PL], 2 * PL * 8);
    if (parity_mode()) {
        next_sym<void (*)(std::vector<MatPoly> &, size_t, size_t, const std::vector<MatPoly> &, const std::vector<MatPoly> &, size_t, size_t)>(
            "_Z20coefficientExpansionRSt6vectorI7MatPolySaIS0_EEmmRKS2_S5_mm")(ref_cv, g, m_exp, W_left_v, W_right_v, max_bits_to_gen_right, stopround);
        std::vector<uint64_t> ref = flatten(ref_cv, ncts);
        cmp_ntt("coefficientExpansion", cv.data(), ref.data(), ncts * 2);
        for (auto &m : ref_cv) free(m.data);
    }
}

// ---- reorientCiphertextsDim1 (src/testing.cpp:342)
void reorientCiphertextsDim1(uint64_t *out, const std::vector<MatPoly> &v_firstdim, size_t dim0, size_t idx_factor) {
    const size_t count = v_firstdim.size();
    std::vector<uint64_t> flat = flatten(v_firstdim, count);
    if (parity_mode() && idx_factor == 1 && count == dim0) g_t3.v_firstdim = flat;       // direct upload: the 2^nu1 first-dimension ciphertexts
    OKAY(sb200_reorientCiphertextsDim1(out, flat.data(), count, dim0, idx_factor));
    if (parity_mode()) {
        std::vector<uint64_t> ref(dim0 * 2 * N, 0);
        next_sym<void (*)(uint64_t *, const std::vector<MatPoly> &, size_t, size_t)>("_Z23reorientCiphertextsDim1PmRKSt6vectorI7MatPolySaIS1_EEmm")(ref.data(), v_firstdim, dim0, idx_factor);
        cmp_packed("reorientCiphertextsDim1", out, ref.data(), ref.size());
    }
}

// ---- regevToSimpleGsw (src/testing.cpp:108): appends further_dims GSW ciphertexts (2 x 2*ell) to v_gsw
void regevToSimpleGsw(std::vector<MatPoly> &v_gsw, const std::vector<MatPoly> &v_inp, const MatPoly &V, size_t m_conv, size_t ell,
                      size_t fdims, size_t idx_factor, size_t idx_offset) {
    const size_t count = v_inp.size(), per = 2 * 2 * ell;
    std::vector<uint64_t> flat = flatten(v_inp, count), res(fdims * per * PL);
    if (parity_mode() && g_t3.V.empty()) g_t3.V.assign(V.data, V.data + V.words());
    OKAY(sb200_regevToSimpleGsw(res.data(), flat.data(), count, V.data, (uint32_t)m_conv, (uint32_t)ell, (uint32_t)fdims, idx_factor, idx_offset));
    if (parity_mode()) {
        std::vector<MatPoly> ref;
        next_sym<void (*)(std::vector<MatPoly> &, const std::vector<MatPoly> &, const MatPoly &, size_t, size_t, size_t, size_t, size_t)>(
            "_Z16regevToSimpleGswRSt6vectorI7MatPolySaIS0_EERKS2_RKS0_mmmmm")(ref, v_inp, V, m_conv, ell, fdims, idx_factor, idx_offset);
        std::vector<uint64_t> rf = flatten(ref, fdims);
        cmp_ntt("regevToSimpleGsw", res.data(), rf.data(), fdims * per);
        for (auto &m : ref) free(m.data);
    }
    for (size_t d = 0; d < fdims; d++) {
        MatPoly ct(2, 2 * ell);
        memcpy(ct.data, &res[d * per * PL], per * PL * 8);
        v_gsw.push_back(ct);
    }
}

// ---- fastMultiplyQueryByDatabaseDim1 (src/testing.cpp:364): scan of the RESIDENT plane when `db` is a buffer convertDb
// returned, otherwise the stateless call
void fastMultiplyQueryByDatabaseDim1(std::vector<MatPoly> &out, const uint64_t *db, const uint64_t *v_firstdim, size_t dim0, size_t num_per) {
    std::vector<uint64_t> res(num_per * 2 * PL);
    size_t plane = g_plane_bufs.size();
    for (size_t k = 0; k < g_plane_bufs.size(); k++) if (g_plane_bufs[k] == db) plane = k;
    if (g_pack && plane < g_plane_bufs.size()) OKAY(sb200_pack_server_scan_plane_host(g_pack, plane, v_firstdim, res.data()));
    else OKAY(sb200_fastMultiplyQueryByDatabaseDim1(res.data(), db, v_firstdim, dim0, num_per));
    if (parity_mode()) {
        std::vector<MatPoly> ref;
        for (size_t i = 0; i < num_per; i++) ref.emplace_back(2, 1);
        next_sym<void (*)(std::vector<MatPoly> &, const uint64_t *, const uint64_t *, size_t, size_t)>(
            "_Z31fastMultiplyQueryByDatabaseDim1RSt6vectorI7MatPolySaIS0_EEPKmS5_mm")(ref, db, v_firstdim, dim0, num_per);
        std::vector<uint64_t> rf = flatten(ref, num_per);
        cmp_ntt("fastMultiplyQueryByDatabaseDim1", res.data(), rf.data(), num_per * 2);
        for (auto &m : ref) free(m.data);
    }
    for (size_t i = 0; i < num_per; i++) memcpy(out[i].data, &res[i * 2 * PL], 2 * PL * 8);
}

// ---- foldCiphertextsDim1 (src/testing.cpp:596): result left in v_cts[0]
void foldCiphertextsDim1(std::vector<MatPoly> &v_cts, const std::vector<MatPoly> &v_folding, const std::vector<MatPoly> &v_folding_neg) {
    const size_t count = v_cts.size(), fd = v_folding.size();
    if (fd == 0) return;
    const uint32_t ell = (uint32_t)(v_folding[0].cols / 2);
    std::vector<uint64_t> cts = flatten(v_cts, count), f = flatten(v_folding, fd), fn = flatten(v_folding_neg, fd);
    if (parity_mode() && g_t3.v_folding.empty()) g_t3.v_folding = f;
    std::vector<MatPoly> ref_cts;
    if (parity_mode()) for (size_t i = 0; i < count; i++) { MatPoly c(2, 1, false); memcpy(c.data, v_cts[i].data, 2 * N * 8); ref_cts.push_back(c); }
    OKAY(sb200_foldCiphertextsDim1(cts.data(), count, f.data(), fn.data(), ell));
    memcpy(v_cts[0].data, cts.data(), 2 * N * 8);
    if (parity_mode()) {
        next_sym<void (*)(std::vector<MatPoly> &, const std::vector<MatPoly> &, const std::vector<MatPoly> &)>(
            "_Z19foldCiphertextsDim1RSt6vectorI7MatPolySaIS0_EERKS2_S5_")(ref_cts, v_folding, v_folding_neg);
        cmp_raw("foldCiphertextsDim1", v_cts[0].data, ref_cts[0].data, 2 * N);
        for (auto &m : ref_cts) free(m.data);
    }
}

// ---- pack (src/testing.cpp:198)
void pack(MatPoly &result, size_t out_n, size_t m_conv, const std::vector<MatPoly> &v_ct, const std::vector<MatPoly> &v_W) {
    std::vector<uint64_t> cts = flatten(v_ct, out_n * out_n), W = flatten(v_W, out_n);
    OKAY(sb200_pack(result.data, (uint32_t)out_n, (uint32_t)m_conv, cts.data(), W.data()));
    if (parity_mode()) {
        MatPoly ref(out_n + 1, out_n);
        next_sym<void (*)(MatPoly &, size_t, size_t, const std::vector<MatPoly> &, const std::vector<MatPoly> &)>(
            "_Z4packR7MatPolymmRKSt6vectorIS_SaIS_EES5_")(ref, out_n, m_conv, v_ct, v_W);
        cmp_ntt("pack", result.data, ref.data, (out_n + 1) * out_n);
        free(ref.data);
        // tier 3: everything the resident pack server needs has passed through the leaves by now
        if (g_pack && g_plane_bufs.size() == out_n * out_n && !g_t3.answered && !g_t3.failed) {
            const bool direct = g_t3.expansions == 0 && !g_t3.v_firstdim.empty();
            const bool packed = g_t3.expansions == 1 && !g_t3.query.empty() && !g_t3.V.empty();
            if (!direct && !packed) { g_t3.failed = true; tier3_note("resident cross-check skipped (no complete query captured)"); return; }
            g_t3.resp.assign((out_n + 1) * out_n * N, 0); g_t3.result_cts.assign(out_n * out_n * 2 * N, 0);
            if (direct) {
                OKAY(sb200_pack_server_set_public_params(g_pack, nullptr, nullptr, nullptr, W.data()));
                OKAY(sb200_pack_server_answer_direct(g_pack, g_t3.v_firstdim.data(), g_t3.v_folding.empty() ? nullptr : g_t3.v_folding.data(),
                                                     g_t3.resp.data(), g_t3.result_cts.data(), nullptr));
            } else {
                OKAY(sb200_pack_server_set_public_params(g_pack, g_t3.W_left.data(), g_t3.W_right.data(), g_t3.V.data(), W.data()));
                OKAY(sb200_pack_server_answer(g_pack, g_t3.query.data(), g_t3.resp.data(), g_t3.result_cts.data(), nullptr));
            }
            g_t3.answered = true;
            cmp_raw("tier-3 resident pack server, folded per-plane ciphertexts", g_t3.result_cts.data(), cts.data(), cts.size());
        }
    }
}
static bool g_pack_tier3() { return g_pack != nullptr && g_t3.answered; }
static bool g_pack_mode() { return g_pack != nullptr; }

static void report_at_exit() {          // testHighRate ends in exit(0) (src/spiral.cpp:1337-1340): report from an atexit handler
    fflush(stdout);
    fprintf(stderr, "[spiral_b200] %llu CUDA kernel launches\n", (unsigned long long)sb200_launch_count());
    std::vector<char> names(sb200_kernel_log(nullptr, 0) + 1);
    sb200_kernel_log(names.data(), names.size());
    fprintf(stderr, "[spiral_b200] kernels: %s\n", names.data());
}

// ---- driver entry: run the reference's own main (client + harness) with the definitions above bound
int main(int argc, char **argv) {
    typedef int (*main_fn)(int, char **);
    main_fn ref_main = (main_fn)dlsym(RTLD_NEXT, "main");
    if (!ref_main) { fprintf(stderr, "reference main not found: %s\n", dlerror()); return 3; }
    if (sb200_init(0) != 0) { fprintf(stderr, "[spiral_b200] %s\n", sb200_last_error()); return 4; }
    fprintf(stderr, "[spiral_b200] driving the reference harness with the CUDA server path%s\n", parity_mode() ? " (parity mode)" : "");
    atexit(report_at_exit);
    ref_main(argc, argv);
    return 0;
}
