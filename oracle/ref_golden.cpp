// ref_golden.cpp - golden-digest generator.  TEST INFRASTRUCTURE ONLY.
//
// Links the UNMODIFIED reference (oracle/_ref/libspiral_ref_<cfg>_<isa>.so) and liboracle.so,
// feeds every parity case of oracle/golden_cases.c to the reference's own functions, checks
// reference == oracle in-process (raw buffers exactly, NTT-domain buffers modulo the prime) and
// prints one JSON document with the reference digests.  Run by oracle/make_golden.sh in the
// build container (where /root/reference exists); the output is committed as
// tests/golden/ref_digests_<cfg>.json so the CPU test-suite can pin the oracle without the
// reference sources.  Only the reference's HEADERS are included (for MatPoly and prototypes).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
#include <unistd.h>
#include <fcntl.h>

#include "poly.h"      // reference: MatPoly, multiply, to_ntt, from_ntt, automorph, getRescaled ...
#include "util.h"      // reference: gadget_invert, buildGadget
#include "client.h"
#include "testing.h"   // reference: pack
#include "golden_cases.h"

// ---- reference symbols that have no header declaration (src/spiral.cpp, src/testing.cpp) ----
extern size_t num_expansions, further_dims, total_n, IDX_TARGET;
extern uint64_t *B;
class FurtherDimsLocals {   // layout-compatible with include/spiral.h:86-127 (passed by value)
public:
    uint64_t *result, *cts, *scratch_cts1, *scratch_cts2, *scratch_cts_double1, *scratch_cts_double2;
    size_t num_per, num_bytes_C;
    FurtherDimsLocals(size_t np) : num_per(np) {
        num_bytes_C = sizeof(uint64_t) * np * n1 * n2 * 2 * poly_len;
        result = (uint64_t *)calloc(2 * num_bytes_C, 1);
        cts = (uint64_t *)calloc(2 * num_bytes_C, 1);
        scratch_cts1 = (uint64_t *)calloc(num_bytes_C, 1);
        scratch_cts2 = (uint64_t *)calloc(num_bytes_C, 1);
        scratch_cts_double1 = (uint64_t *)calloc(m2 / n1 * num_bytes_C, 1);
        scratch_cts_double2 = (uint64_t *)calloc(m2 / n1 * num_bytes_C, 1);
    }
};
void modswitch(uint64_t *out, const uint64_t *inp);      // src/spiral.cpp:40
void setup_constants();
void set_neg1s();
void load_db();
void reorientCiphertexts(uint64_t *out, const uint64_t *inp, size_t dim0, size_t n1_padded);
void multiplyQueryByDatabase(uint64_t *output, const uint64_t *reorientedCiphertexts, const uint64_t *database, size_t dim0, size_t num_per);
void nttInvAndCrtLiftCiphertexts(size_t num_per, FurtherDimsLocals furtherDimsLocals);
void split_and_crt(uint64_t *out, const uint64_t *in, size_t num_per);
void foldOneFurtherDimension(size_t cur_dim, size_t num_per, const uint64_t *query_ct, const uint64_t *query_ct_neg, FurtherDimsLocals locals);
double expandImproved(std::vector<MatPoly> &cv_v, size_t g, size_t m_exp, const std::vector<MatPoly> &W_left_v, const std::vector<MatPoly> &W_right_v, size_t max_bits_to_gen_right, size_t stopround);
void coefficientExpansion(std::vector<MatPoly> &cv_v, size_t g, size_t m_exp, const std::vector<MatPoly> &W_left_v, const std::vector<MatPoly> &W_right_v, size_t max_bits_to_gen_right, size_t stopround);
void scalToMat(size_t m_conv, MatPoly &out_reg, const MatPoly &cv, const MatPoly &W);
void regevToGSW(size_t m_conv, size_t t, MatPoly &out, const std::vector<MatPoly> &cv_v, size_t cv_v_offset, const MatPoly &W, const MatPoly &V);
uint64_t *convertDb(const std::vector<MatPoly> &db, size_t dim0, size_t num_per);
void reorientCiphertextsDim1(uint64_t *out, const std::vector<MatPoly> &v_firstdim, size_t dim0, size_t idx_factor);
void fastMultiplyQueryByDatabaseDim1(std::vector<MatPoly> &out, const uint64_t *db, const uint64_t *v_firstdim, size_t dim0, size_t num_per);
void foldCiphertextsDim1(std::vector<MatPoly> &v_cts, const std::vector<MatPoly> &v_folding, const std::vector<MatPoly> &v_folding_neg);
void regevToSimpleGsw(std::vector<MatPoly> &v_gsw, const std::vector<MatPoly> &v_inp, const MatPoly &V, size_t m_conv, size_t ell, size_t further_dims, size_t idx_factor, size_t idx_offset);

// ---- rand() interposer: load_db draws plaintext coefficients from libc rand() ----------------
static const uint64_t *g_rand_feed = nullptr;
static size_t g_rand_pos = 0, g_rand_len = 0;
extern "C" int rand(void) {
    if (g_rand_feed && g_rand_pos < g_rand_len) return (int)g_rand_feed[g_rand_pos++];
    return 4;
}

// the reference chats on stdout; mute it around reference calls
static int g_saved_stdout = -1;
static void mute() { fflush(stdout); g_saved_stdout = dup(1); int nul = open("/dev/null", O_WRONLY); dup2(nul, 1); close(nul); }
static void unmute() { fflush(stdout); dup2(g_saved_stdout, 1); close(g_saved_stdout); }

static const size_t PLW = 2 * poly_len;
static MatPoly mk(size_t r, size_t c, bool ntt, const uint64_t *src) {
    MatPoly m(r, c, ntt);
    memcpy(m.data, src, r * c * (ntt ? 2 : 1) * poly_len * sizeof(uint64_t));
    return m;
}
static std::vector<uint64_t> flat(const std::vector<MatPoly> &v) {
    std::vector<uint64_t> out;
    for (auto &m : v) { size_t w = m.rows * m.cols * (m.isNTT ? 2 : 1) * poly_len; out.insert(out.end(), m.data, m.data + w); }
    return out;
}

static std::vector<uint64_t> run_reference(int id, const so_params *p, so_case_io *io) {
    so_case_shape_t s; so_case_shape(id, p, &s);
    std::vector<uint64_t> out;
    switch (id) {
    case SO_CASE_NTT_FWD: case SO_CASE_NTT_INV:
        out.assign(io->in[0], io->in[0] + io->in_words[0]);
        for (size_t i = 0; i < s.npolys; i++) (id == SO_CASE_NTT_FWD ? ntt_forward : ntt_inverse)(&out[i * PLW]);
        break;
    case SO_CASE_TO_NTT: { MatPoly a = mk(1, s.npolys, false, io->in[0]); MatPoly o(1, s.npolys); to_ntt(o, a); out = flat({o}); break; }
    case SO_CASE_MODSWITCH:
        out.assign((n1 * n2 * poly_len * bits_to_hold_arb_qprime + 63) / 64 + 1, 0);     // +1: the 128-bit store of write_arbitrary_bits
        modswitch(out.data(), io->in[0]);
        out.pop_back(); break;
    case SO_CASE_TO_NTT_NR: { MatPoly a = mk(1, s.npolys, false, io->in[0]); MatPoly o(1, s.npolys); to_ntt_no_reduce(o, a); out = flat({o}); break; }
    case SO_CASE_FROM_NTT: { MatPoly a = mk(1, s.npolys, true, io->in[0]); MatPoly o(1, s.npolys, false); from_ntt(o, a); out = flat({o}); break; }
    case SO_CASE_MULTIPLY: { MatPoly a = mk(2, 3, true, io->in[0]), b = mk(3, 2, true, io->in[1]); MatPoly o(2, 2); multiply(o, a, b); out = flat({o}); break; }
    case SO_CASE_AUTOMORPH: { MatPoly a = mk(1, s.npolys, false, io->in[0]); MatPoly o(1, s.npolys, false); automorph(o, a, poly_len / 4 + 1); out = flat({o}); break; }
    case SO_CASE_GADGET_INVERT: { MatPoly a = mk(2, 1, false, io->in[0]); MatPoly o(2 * p->t_conv, 1, false); gadget_invert(2 * p->t_conv, o, a, 2); out = flat({o}); break; }
    case SO_CASE_RESCALE: {
        MatPoly a = mk(1, 2, false, io->in[0]), b = mk(1, 2, false, io->in[0] + 2 * poly_len);
        MatPoly ra = getRescaled(a, Q_i, arb_qprime), rb = getRescaled(b, Q_i, 4 * p_db);
        out = flat({ra, rb}); break; }
    case SO_CASE_REORIENT:
        out.assign(s.dim0 * 2 * 4 * poly_len, 0);
        reorientCiphertexts(out.data(), io->in[0], s.dim0, 4); break;
    case SO_CASE_FIRST_DIM:
        out.assign(s.num_per * 6 * PLW, 0);
        multiplyQueryByDatabase(out.data(), io->in[0], io->in[1], s.dim0, s.num_per); break;
    case SO_CASE_NTT_INV_CRT: {
        FurtherDimsLocals L(s.num_per);
        memcpy(L.scratch_cts1, io->in[0], io->in_words[0] * 8);
        nttInvAndCrtLiftCiphertexts(s.num_per, L);
        out.assign(L.cts, L.cts + s.num_per * 6 * poly_len); break; }
    case SO_CASE_SPLIT_AND_CRT:
        out.assign(s.num_per * m2 * 2 * PLW, 0);
        split_and_crt(out.data(), io->in[0], s.num_per); break;
    case SO_CASE_FOLD_ONE: {
        FurtherDimsLocals L(s.num_per);
        memcpy(L.cts, io->in[0], io->in_words[0] * 8);
        foldOneFurtherDimension(s.cur_dim, s.num_per, io->in[1], io->in[2], L);
        out.assign(L.cts, L.cts + s.num_per * 6 * poly_len); break; }
    case SO_CASE_EXPAND_FULL: case SO_CASE_EXPAND_STOP: {
        std::vector<MatPoly> cv, Wl, Wr;
        for (size_t i = 0; i < ((size_t)1 << s.g); i++) cv.push_back(mk(2, 1, true, io->in[0] + i * 2 * PLW));
        for (size_t r = 0; r < s.g; r++) {
            Wl.push_back(mk(2, m_exp, true, io->in[1] + r * 2 * m_exp * PLW));
            Wr.push_back(mk(2, m_exp_right, true, io->in[2] + r * 2 * m_exp_right * PLW));
        }
        // Spiral and Pack variants share the body; exercise both entry points
        std::vector<MatPoly> cv2; for (auto &m : cv) { MatPoly c; c = m; cv2.push_back(c); }
        expandImproved(cv, s.g, m_exp, Wl, Wr, s.max_bits_right, s.stopround);
        coefficientExpansion(cv2, s.g, m_exp, Wl, Wr, s.max_bits_right, s.stopround);
        out = flat(cv);
        std::vector<uint64_t> out2 = flat(cv2);
        if (so_digest_kind(out.data(), out.size(), SO_KIND_NTT) != so_digest_kind(out2.data(), out2.size(), SO_KIND_NTT)) {
            fprintf(stderr, "expandImproved != coefficientExpansion\n"); exit(1);
        }
        break; }
    case SO_CASE_SCAL_TO_MAT: {
        MatPoly cv = mk(2, 1, true, io->in[0]), W = mk(n1, 2 * m_conv, true, io->in[1]); MatPoly o(n1, n0);
        scalToMat(m_conv, o, cv, W); out = flat({o}); break; }
    case SO_CASE_REGEV_TO_GSW: {
        std::vector<MatPoly> cv; for (size_t i = 0; i < t_GSW; i++) cv.push_back(mk(2, 1, true, io->in[0] + i * 2 * PLW));
        MatPoly W = mk(n1, 2 * m_conv, true, io->in[1]), V = mk(n1, 2 * m_conv, true, io->in[2]); MatPoly o(n1, m2);
        regevToGSW(m_conv, t_GSW, o, cv, 0, W, V); out = flat({o}); break; }
    case SO_CASE_LOAD_DB: {
        num_expansions = 0; while (((size_t)1 << num_expansions) < s.dim0) num_expansions++;
        further_dims = 0; while (((size_t)1 << further_dims) < s.num_per) further_dims++;
        total_n = s.dim0 * s.num_per; IDX_TARGET = 1; random_data = false;
        g_rand_feed = io->in[0]; g_rand_pos = 0; g_rand_len = io->in_words[0];
        load_db();
        g_rand_feed = nullptr;
        out.assign(B, B + total_n * 4 * poly_len); break; }
    case SO_CASE_CONVERT_DB: {
        std::vector<MatPoly> db; for (size_t i = 0; i < s.dim0 * s.num_per; i++) db.push_back(mk(1, 1, true, io->in[0] + i * PLW));
        uint64_t *buf = convertDb(db, s.dim0, s.num_per); out.assign(buf, buf + s.dim0 * s.num_per * poly_len); break; }
    case SO_CASE_REORIENT_DIM1: {
        std::vector<MatPoly> v; for (size_t i = 0; i < 2 * s.dim0; i++) v.push_back(mk(2, 1, true, io->in[0] + i * 2 * PLW));
        out.assign(s.dim0 * 2 * poly_len, 0); reorientCiphertextsDim1(out.data(), v, s.dim0, 2); break; }
    case SO_CASE_FIRST_DIM_PACK: {
        std::vector<MatPoly> o; for (size_t i = 0; i < s.num_per; i++) o.emplace_back(2, 1);
        fastMultiplyQueryByDatabaseDim1(o, io->in[1], io->in[0], s.dim0, s.num_per); out = flat(o); break; }
    case SO_CASE_FOLD_DIM1: {
        std::vector<MatPoly> cts, f, fn;
        for (size_t i = 0; i < s.num_per; i++) cts.push_back(mk(2, 1, false, io->in[0] + i * 2 * poly_len));
        for (size_t d = 0; d < 2; d++) {
            f.push_back(mk(2, 2 * t_GSW, true, io->in[1] + d * 2 * 2 * t_GSW * PLW));
            fn.push_back(mk(2, 2 * t_GSW, true, io->in[2] + d * 2 * 2 * t_GSW * PLW));
        }
        foldCiphertextsDim1(cts, f, fn); out = flat({cts[0]}); break; }
    case SO_CASE_REGEV_TO_SGSW: {
        std::vector<MatPoly> v, gsw; for (size_t i = 0; i < 4 * t_GSW + 2; i++) v.push_back(mk(2, 1, true, io->in[0] + i * 2 * PLW));
        MatPoly V = mk(2, 2 * m_conv, true, io->in[1]);
        regevToSimpleGsw(gsw, v, V, m_conv, t_GSW, 2, 2, 1); out = flat(gsw); break; }
    case SO_CASE_PACK: {
        std::vector<MatPoly> cts, Ws;
        for (size_t i = 0; i < out_n * out_n; i++) cts.push_back(mk(2, 1, false, io->in[0] + i * 2 * poly_len));
        for (size_t i = 0; i < out_n; i++) Ws.push_back(mk(out_n + 1, m_conv, true, io->in[1] + i * (out_n + 1) * m_conv * PLW));
        MatPoly res(out_n + 1, out_n); pack(res, out_n, m_conv, cts, Ws); out = flat({res}); break; }
    }
    return out;
}

int main(int argc, char **argv) {
    const char *cfg = argc > 1 ? argv[1] : "cfg?";
    uint64_t seed = argc > 2 ? strtoull(argv[2], nullptr, 10) : 20220368ull;
    so_params p;
    p.nu1 = 0; p.nu2 = 0; p.t_gsw = t_GSW; p.t_conv = m_conv; p.t_exp = m_exp; p.t_exp_right = m_exp_right;
    p.qp_bits = bits_to_hold_arb_qprime; p.out_n = out_n; p.p_db = p_db;

    scratch = (uint64_t *)malloc(crt_count * poly_len * sizeof(uint64_t));
    fprintf(stderr, "[ref_golden %s] setup\n", cfg);
    mute();
    setup_constants();
    set_neg1s();
    unmute();

    // tables: regenerated-from-psi vs the reference's constants (src/constants.cpp:16)
    const uint64_t *mine = so_tables();
    int tables_ok = memcmp(mine, tables, 8 * poly_len * sizeof(uint64_t)) == 0;

    std::string json = "{\n  \"config\": \"" + std::string(cfg) + "\",\n";
    char buf[512];
    snprintf(buf, sizeof buf, "  \"params\": {\"t_gsw\": %u, \"t_conv\": %u, \"t_exp\": %u, \"t_exp_right\": %u, \"qp_bits\": %u, \"out_n\": %u, \"p_db\": %llu},\n  \"seed\": %llu,\n  \"tables_match_reference\": %s,\n  \"tables_digest\": \"%016llx\",\n  \"cases\": {\n",
             p.t_gsw, p.t_conv, p.t_exp, p.t_exp_right, p.qp_bits, p.out_n, (unsigned long long)p.p_db,
             (unsigned long long)seed, tables_ok ? "true" : "false",
             (unsigned long long)so_fnv1a64(tables, 8 * poly_len));
    json += buf;

    int failures = tables_ok ? 0 : 1;
    for (int id = 0; id < so_case_count(); id++) {
        so_case_io io;
        so_case_make_inputs(id, &p, seed, &io);
        so_case_run_oracle(id, &p, &io);
        mute();
        std::vector<uint64_t> ref = run_reference(id, &p, &io);
        unmute();
        uint64_t d_ref = so_digest_kind(ref.data(), ref.size(), io.out_kind);
        uint64_t d_orc = so_case_digest(&io);
        bool same = ref.size() == io.out_words && d_ref == d_orc;
        bool exact = same && memcmp(ref.data(), io.out, io.out_words * 8) == 0;
        if (!same) failures++;
        fprintf(stderr, "[ref_golden %s] %-22s words=%zu ref=%016llx oracle=%016llx %s%s\n", cfg, so_case_name(id),
                ref.size(), (unsigned long long)d_ref, (unsigned long long)d_orc, same ? "OK" : "MISMATCH",
                (same && !exact) ? " (equal mod q only)" : "");
        snprintf(buf, sizeof buf, "    \"%s\": {\"id\": %d, \"words\": %zu, \"kind\": %d, \"digest\": \"%016llx\"}%s\n",
                 so_case_name(id), id, ref.size(), io.out_kind, (unsigned long long)d_ref, id + 1 < so_case_count() ? "," : "");
        json += buf;
        so_case_free(&io);
    }
    json += "  }\n}\n";
    if (argc > 3) { FILE *f = fopen(argv[3], "w"); fputs(json.c_str(), f); fclose(f); }
    else fputs(json.c_str(), stderr);
    fprintf(stderr, "[ref_golden %s] %s (%d failures)\n", cfg, failures ? "FAILED" : "all cases match", failures);
    return failures ? 1 : 0;
}
