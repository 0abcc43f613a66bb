// api.cu - the C-ABI of libspiral_b200.so (include/spiral_b200.h).
#include "../../include/spiral_b200.h"
#include "kernels.cuh"
#include "common.cuh"
#include "ntt.cuh"
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace sb200;

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver ? SB200_ERR_NO_DEVICE : SB200_ERR_CUDA, \
                                           "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CHECK_LAUNCH() CU(cudaGetLastError())

static int ensure_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(SB200_ERR_NO_DEVICE, "no CUDA device: libspiral_b200 has no CPU fallback (%s)", e == cudaSuccess ? "0 devices" : cudaGetErrorString(e));
    }
    int rc = init_tables();
    if (rc != 0) return fail(SB200_ERR_CUDA, "twiddle-table initialisation failed (%d)", rc);
    return SB200_OK;
}
#define NEED_DEVICE() do { int rc_ = ensure_device(); if (rc_) return rc_; } while (0)

extern "C" const char *sb200_last_error(void) { return g_err.c_str(); }
extern "C" int sb200_abi_version(void) { return 1; }
extern "C" uint64_t sb200_launch_count(void) { return launch_count(); }
extern "C" size_t sb200_kernel_log(char *buf, size_t cap) { return kernel_log(buf, cap); }
extern "C" void sb200_kernel_log_reset(void) { kernel_log_reset(); }
extern "C" int sb200_trace_enable(uint32_t capacity) { int rc_ = ensure_device(); if (rc_) return rc_; return trace_enable(capacity) ? fail(SB200_ERR_CUDA, "trace_enable failed") : SB200_OK; }
extern "C" size_t sb200_trace_read(uint64_t *out, size_t max_records, int reset) { return trace_read(reinterpret_cast<unsigned long long *>(out), max_records, reset); }
extern "C" uint64_t sb200_arb_qprime(uint32_t qp_bits) {     // reference include/values.h:74-76
    static const uint64_t qprime_mods[37] = {
        0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 12289, 12289, 61441, 65537, 65537, 520193, 786433,
        786433, 3604481, 7340033, 16515073, 33292289, 67043329, 132120577, 268369921, 469762049,
        1073479681, 2013265921, 4293918721ull, 8588886017ull, 17175674881ull, 34359214081ull, 68718428161ull};
    return qp_bits < 37 ? qprime_mods[qp_bits] : 0;
}
extern "C" int sb200_init(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) { cudaGetLastError(); return fail(SB200_ERR_NO_DEVICE, "no CUDA device: libspiral_b200 has no CPU fallback"); }
    if (device < 0 || device >= n) return fail(SB200_ERR_ARG, "device %d out of range (%d devices)", device, n);
    CU(cudaSetDevice(device));
    return ensure_device();
}

// ---------------------------------------------------------------------------------------------
// tier 1: device pointers
// ---------------------------------------------------------------------------------------------
static inline cudaStream_t S(void *s) { return (cudaStream_t)s; }

extern "C" int sb200_dev_ntt_from_ref(uint32_t *out, const uint64_t *in, size_t npolys, void *stream) {
    NEED_DEVICE(); launch_ntt_u64_to_dev(out, in, npolys, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" int sb200_dev_ntt_to_ref(uint64_t *out, const uint32_t *in, size_t npolys, void *stream) {
    NEED_DEVICE(); launch_ntt_dev_to_u64(out, in, npolys, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" int sb200_dev_to_ntt(uint32_t *out, const uint64_t *raw, size_t npolys, void *stream) {
    NEED_DEVICE(); launch_to_ntt(out, raw, npolys, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" int sb200_dev_from_ntt(uint64_t *raw, const uint32_t *in, size_t npolys, void *stream) {
    NEED_DEVICE(); launch_from_ntt(raw, in, npolys, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" int sb200_dev_multiply(uint32_t *out, const uint32_t *a, const uint32_t *b, int rs, int ms, int cs, void *stream) {
    NEED_DEVICE(); launch_matmul(out, a, b, rs, ms, cs, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" int sb200_dev_automorph(uint64_t *out, const uint64_t *in, size_t npolys, uint32_t t, void *stream) {
    NEED_DEVICE(); launch_automorph(out, in, npolys, t, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" int sb200_dev_gadget_ntt(uint32_t *out, const uint64_t *raw, int mx, int rdim, int cols, void *stream) {
    NEED_DEVICE();
    if (rdim <= 0 || mx % rdim) return fail(SB200_ERR_ARG, "gadget_ntt: mx %% rdim != 0");
    launch_gadget_ntt(out, raw, mx, rdim, cols, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" int sb200_dev_rescale(uint64_t *out, const uint64_t *in, size_t n, uint64_t inp_mod, uint64_t out_mod, void *stream) {
    NEED_DEVICE(); launch_rescale(out, in, n, inp_mod, out_mod, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" size_t sb200_packed_words(size_t ncoeffs, uint32_t bits) { return (ncoeffs * bits + 63) / 64; }
extern "C" int sb200_dev_modswitch(uint64_t *out_words, const uint64_t *cts_raw, size_t ncoeffs, uint32_t qp_bits, void *stream) {
    NEED_DEVICE();
    if (qp_bits == 0 || qp_bits > 36 || !sb200_arb_qprime(qp_bits)) return fail(SB200_ERR_ARG, "modswitch: no arb_qprime for %u bits", qp_bits);
    launch_modswitch(out_words, cts_raw, ncoeffs, qp_bits, sb200_arb_qprime(qp_bits), S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" int sb200_dev_bitpack(uint64_t *out_words, const uint64_t *values, size_t n, uint32_t bits, void *stream) {
    NEED_DEVICE();
    if (bits == 0 || bits > 63) return fail(SB200_ERR_ARG, "bitpack: bits must be in [1, 63]");
    launch_bitpack(out_words, values, n, bits, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" size_t sb200_db_words(uint32_t nu1, uint32_t nu2) { return ((size_t)kN << (nu1 + nu2)) * kN0 * kN2; }
extern "C" int sb200_dev_db_build(uint64_t *db, const uint16_t *pts, uint32_t nu1, uint32_t nu2, uint32_t p_db,
                                  size_t item_begin, size_t item_count, void *stream) {
    NEED_DEVICE();
    if (p_db == 0 || p_db > 65536) return fail(SB200_ERR_ARG, "db_build: p_db must be in [1, 65536]");
    launch_db_build_spiral(db, pts, (int)nu1, (int)nu2, p_db, item_begin, item_count, S(stream));
    CHECK_LAUNCH(); return SB200_OK;
}
extern "C" int sb200_dev_db_from_reference(uint64_t *db, const uint64_t *B_chunk, size_t dim0, size_t num_per,
                                           size_t z_begin, size_t z_count, void *stream) {
    NEED_DEVICE(); launch_db_from_reference(db, B_chunk, dim0, num_per * kN2, z_begin, z_count, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" int sb200_dev_reorient_query(uint64_t *out, const uint32_t *cts, size_t dim0, void *stream) {
    NEED_DEVICE(); launch_reorient_query(out, cts, dim0, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" int sb200_dev_first_dim(uint32_t *out, const uint64_t *query, const uint64_t *db, size_t dim0, size_t num_per, void *stream) {
    NEED_DEVICE();
    if (!dim0 || !num_per || (dim0 & (dim0 - 1)) || (num_per & (num_per - 1))) return fail(SB200_ERR_ARG, "first_dim: dim0 and num_per must be powers of two");
    launch_scan_spiral(out, query, db, dim0, num_per, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
// ---- batched first dimension on tensor cores (tc_scan.cu)
extern "C" int sb200_tc_supported(size_t dim0, size_t num_per) { return tc_shape_ok(dim0, num_per); }
extern "C" size_t sb200_tc_query_bytes(size_t dim0, int capacity) { return tc_query_bytes(dim0, capacity); }
extern "C" int sb200_dev_db_to_tc(uint8_t *db_tc, const uint64_t *db, size_t dim0, size_t num_per, void *stream) {
    NEED_DEVICE();
    if (!db_tc || !db || !tc_shape_ok(dim0, num_per)) return fail(SB200_ERR_ARG, "db_to_tc: needs 2*dim0 and 2*num_per to be multiples of 128");
    launch_db_to_tc(db_tc, db, dim0, num_per, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" int sb200_dev_query_to_tc(uint8_t *q_tc, const uint64_t *query, int q, int capacity, size_t dim0, void *stream) {
    NEED_DEVICE();
    if (!q_tc || !query || q < 0 || q >= capacity || capacity > 16 || (dim0 * 2) % 128) return fail(SB200_ERR_ARG, "query_to_tc: bad slot, capacity or dim0");
    launch_query_to_tc(q_tc, query, q, capacity, dim0, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}
extern "C" size_t sb200_tc_scratch_bytes(size_t num_per, int count) { return tc_scratch_bytes(num_per, count); }
extern "C" int sb200_dev_first_dim_tc(uint32_t *const *out, int count, int capacity, const uint8_t *q_tc, const uint8_t *db_tc,
                                      size_t dim0, size_t num_per, uint32_t *scratch, void *stream) {
    NEED_DEVICE();
    if (!out || !q_tc || !db_tc || !scratch) return fail(SB200_ERR_ARG, "first_dim_tc: null argument");
    const int rc = launch_scan_tc(out, count, capacity, q_tc, db_tc, dim0, num_per, scratch, S(stream));
    if (rc == -1) return fail(SB200_ERR_ARG, "first_dim_tc: needs 1 <= count <= capacity <= 16, 2*dim0 and 2*num_per multiples of 128");
    if (rc) return fail(SB200_ERR_CUDA, "first_dim_tc: cannot reserve %zu bytes of shared memory", (size_t)0);
    CHECK_LAUNCH(); return SB200_OK;
}
extern "C" size_t sb200_fold_scratch_words(size_t num_per_after, uint32_t t_gsw) { return fold_scratch_words(num_per_after, (int)t_gsw); }
extern "C" int sb200_dev_fold_round(uint64_t *cts, size_t num_per_after, const uint32_t *q, const uint32_t *q_neg,
                                    uint32_t t_gsw, uint32_t *scratch, void *stream) {
    NEED_DEVICE(); launch_fold_round(cts, num_per_after, q, q_neg, (int)t_gsw, scratch, S(stream)); CHECK_LAUNCH(); return SB200_OK;
}

// ---------------------------------------------------------------------------------------------
// tier 2: host pointers in the reference's layouts.  Small RAII device buffers + sync copies.
// ---------------------------------------------------------------------------------------------
namespace {
template <typename T>
struct DBuf {
    T *p = nullptr; size_t n = 0;
    DBuf() {}
    explicit DBuf(size_t count) { alloc(count); }
    ~DBuf() { if (p) cudaFree(p); }
    DBuf(const DBuf &) = delete; DBuf &operator=(const DBuf &) = delete;
    cudaError_t alloc(size_t count) { n = count; return cudaMalloc(&p, (count ? count : 1) * sizeof(T)); }
    cudaError_t up(const T *h, size_t count) { return cudaMemcpy(p, h, count * sizeof(T), cudaMemcpyHostToDevice); }
    cudaError_t down(T *h, size_t count) const { return cudaMemcpy(h, p, count * sizeof(T), cudaMemcpyDeviceToHost); }
};
constexpr size_t PLW = 2 * (size_t)kN;   // words of one NTT-form polynomial

// ref-NTT host buffer -> dev-NTT device buffer
int up_ntt(DBuf<uint32_t> &dst, const uint64_t *host, size_t npolys) {
    DBuf<uint64_t> tmp;
    CU(tmp.alloc(npolys * PLW)); CU(tmp.up(host, npolys * PLW));
    CU(dst.alloc(npolys * PLW));
    launch_ntt_u64_to_dev(dst.p, tmp.p, npolys, 0); CHECK_LAUNCH();
    CU(cudaDeviceSynchronize());
    return SB200_OK;
}
int down_ntt(uint64_t *host, const uint32_t *dev, size_t npolys) {
    DBuf<uint64_t> tmp;
    CU(tmp.alloc(npolys * PLW));
    launch_ntt_dev_to_u64(tmp.p, dev, npolys, 0); CHECK_LAUNCH();
    CU(tmp.down(host, npolys * PLW));
    return SB200_OK;
}
#define TRY(x) do { int rc__ = (x); if (rc__) return rc__; } while (0)
}  // namespace

extern "C" int sb200_to_ntt(uint64_t *out, const uint64_t *raw, size_t npolys) {
    NEED_DEVICE();
    DBuf<uint64_t> d_raw; DBuf<uint32_t> d_out;
    CU(d_raw.alloc(npolys * kN)); CU(d_raw.up(raw, npolys * kN)); CU(d_out.alloc(npolys * PLW));
    launch_to_ntt(d_out.p, d_raw.p, npolys, 0); CHECK_LAUNCH();
    return down_ntt(out, d_out.p, npolys);
}
extern "C" int sb200_from_ntt(uint64_t *raw, const uint64_t *in, size_t npolys) {
    NEED_DEVICE();
    DBuf<uint32_t> d_in; DBuf<uint64_t> d_raw;
    TRY(up_ntt(d_in, in, npolys)); CU(d_raw.alloc(npolys * kN));
    launch_from_ntt(d_raw.p, d_in.p, npolys, 0); CHECK_LAUNCH();
    CU(d_raw.down(raw, npolys * kN));
    return SB200_OK;
}
// ntt_forward / ntt_inverse on ref-NTT buffers: forward = to_ntt of the plane values; inverse without CRT.
// Implemented through the raw path: forward treats each plane as a raw polynomial reduced mod its own prime.
__global__ void __launch_bounds__(sb200::kNttThreads) k_ntt_only(uint32_t *io, int inverse) {
    pdl_prologue();
    __shared__ __align__(16) uint32_t sm[2][sb200::kPlaneWords];
    const int n = sb200::plane_of_thread(), lt = sb200::lane_in_plane();
    uint32_t *pl = io + ((size_t)blockIdx.x * 2 + n) * sb200::kN;
    uint32_t v[16];
    if (!inverse) {
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] = pl[sb200::nat_pos(lt, k)];
        sb200::ntt_forward_plane(v, sm[n], lt, n);
        sb200::plane_sync(n);
        sb200::store_ntt_regs(v, pl, lt);
    } else {
        sb200::load_ntt_regs(v, pl, lt);
        sb200::ntt_inverse_plane(v, sm[n], lt, n);
        sb200::plane_sync(n);
#pragma unroll
        for (int k = 0; k < 16; k++) pl[sb200::nat_pos(lt, k)] = v[k];
    }
}
static int ntt_only(uint64_t *io, size_t npolys, int inverse) {
    NEED_DEVICE();
    DBuf<uint32_t> d;
    TRY(up_ntt(d, io, npolys));
    if (npolys) { count_launch(); launch_pdl(k_ntt_only, dim3((unsigned)npolys), dim3(kNttThreads), 0, 0, d.p, inverse); }
    CHECK_LAUNCH();
    return down_ntt(io, d.p, npolys);
}
extern "C" int sb200_ntt_forward(uint64_t *io, size_t npolys) { return ntt_only(io, npolys, 0); }
extern "C" int sb200_ntt_inverse(uint64_t *io, size_t npolys) { return ntt_only(io, npolys, 1); }

extern "C" int sb200_multiply(uint64_t *out, const uint64_t *a, const uint64_t *b, int rs, int ms, int cs) {
    NEED_DEVICE();
    DBuf<uint32_t> da, db, dout;
    TRY(up_ntt(da, a, (size_t)rs * ms)); TRY(up_ntt(db, b, (size_t)ms * cs)); CU(dout.alloc((size_t)rs * cs * PLW));
    launch_matmul(dout.p, da.p, db.p, rs, ms, cs, 0); CHECK_LAUNCH();
    return down_ntt(out, dout.p, (size_t)rs * cs);
}
extern "C" int sb200_automorph(uint64_t *out, const uint64_t *in, size_t npolys, uint32_t t) {
    NEED_DEVICE();
    DBuf<uint64_t> di(npolys * kN), dout(npolys * kN);
    CU(di.up(in, npolys * kN));
    launch_automorph(dout.p, di.p, npolys, t, 0); CHECK_LAUNCH();
    CU(dout.down(out, npolys * kN));
    return SB200_OK;
}
extern "C" int sb200_gadget_invert(uint64_t *out, const uint64_t *in, int mx, int rdim, int cols) {
    NEED_DEVICE();
    if (rdim <= 0 || mx % rdim) return fail(SB200_ERR_ARG, "gadget_invert: mx %% rdim != 0");
    DBuf<uint64_t> di((size_t)rdim * cols * kN), dout((size_t)mx * cols * kN);
    CU(di.up(in, (size_t)rdim * cols * kN));
    launch_gadget_raw(dout.p, di.p, mx, rdim, cols, 0); CHECK_LAUNCH();
    CU(dout.down(out, (size_t)mx * cols * kN));
    return SB200_OK;
}
extern "C" int sb200_getRescaled(uint64_t *out, const uint64_t *in, size_t n, uint64_t inp_mod, uint64_t out_mod) {
    NEED_DEVICE();
    DBuf<uint64_t> di(n), dout(n);
    CU(di.up(in, n));
    launch_rescale(dout.p, di.p, n, inp_mod, out_mod, 0); CHECK_LAUNCH();
    CU(dout.down(out, n));
    return SB200_OK;
}
// modswitch(furtherDimsLocals.result, furtherDimsLocals.cts) (src/spiral.cpp:40, call :2460): n1 x n2 x 2048 coefficients
extern "C" int sb200_modswitch(uint64_t *out_words, const uint64_t *cts_raw, uint32_t qp_bits) {
    NEED_DEVICE();
    const size_t n = (size_t)kN1 * kN2 * kN, words = sb200_packed_words(n, qp_bits);
    DBuf<uint64_t> di(n), dout(words);
    CU(di.up(cts_raw, n));
    TRY(sb200_dev_modswitch(dout.p, di.p, n, qp_bits, nullptr));
    CU(dout.down(out_words, words));
    return SB200_OK;
}
// ---- response wire format (sizes as print_summary counts them, src/spiral.cpp:229-232): the modulus-switched
// response bit-packed, row 0 at qp_bits bits per coefficient, the remaining rows at log2(4 * p_db) bits
static uint32_t rest_bits_of(uint64_t p_db) { uint32_t b = 0; while ((1ull << b) < 4 * p_db) b++; return b; }
extern "C" size_t sb200_packed_response_words(size_t row0_coeffs, size_t rest_coeffs, uint32_t qp_bits, uint64_t p_db) {
    return sb200_packed_words(row0_coeffs, qp_bits) + sb200_packed_words(rest_coeffs, rest_bits_of(p_db));
}
extern "C" int sb200_dev_pack_response(uint64_t *packed, const uint64_t *total_resp, size_t row0_coeffs, size_t rest_coeffs,
                                       uint32_t qp_bits, uint64_t p_db, void *stream) {
    NEED_DEVICE();
    if (qp_bits == 0 || qp_bits > 63 || p_db == 0) return fail(SB200_ERR_ARG, "pack_response: bad parameters");
    launch_bitpack(packed, total_resp, row0_coeffs, qp_bits, S(stream));
    launch_bitpack(packed + sb200_packed_words(row0_coeffs, qp_bits), total_resp + row0_coeffs, rest_coeffs, rest_bits_of(p_db), S(stream));
    CHECK_LAUNCH(); return SB200_OK;
}
// client-side helper (plain host code, no device): the inverse of sb200_dev_pack_response = read_arbitrary_bits (src/core.cpp:20-30)
extern "C" int sb200_unpack_response(uint64_t *total_resp, const uint64_t *packed, size_t row0_coeffs, size_t rest_coeffs,
                                     uint32_t qp_bits, uint64_t p_db) {
    if (!total_resp || !packed || qp_bits == 0 || qp_bits > 63 || p_db == 0) return fail(SB200_ERR_ARG, "unpack_response: bad argument");
    auto rd = [](const uint64_t *p, size_t off, uint32_t b) {
        const size_t w = off / 64, in = off % 64;
        uint64_t v = p[w] >> in;
        if (in + b > 64) v |= p[w + 1] << (64 - in);
        return v & ((1ull << b) - 1);
    };
    for (size_t i = 0; i < row0_coeffs; i++) total_resp[i] = rd(packed, i * qp_bits, qp_bits);
    const uint64_t *seg = packed + sb200_packed_words(row0_coeffs, qp_bits);
    const uint32_t rb = rest_bits_of(p_db);
    for (size_t i = 0; i < rest_coeffs; i++) total_resp[row0_coeffs + i] = rd(seg, i * rb, rb);
    return SB200_OK;
}
extern "C" int sb200_load_db(uint64_t *B, const uint64_t *pts, uint32_t nu1, uint32_t nu2, uint64_t p_db) {
    NEED_DEVICE();
    if (p_db == 0 || p_db > 65536) return fail(SB200_ERR_ARG, "load_db: p_db must be in [1, 65536]");
    const size_t total_n = (size_t)1 << (nu1 + nu2), words = sb200_db_words(nu1, nu2);
    std::vector<uint16_t> h16(total_n * 4 * kN);
    for (size_t i = 0; i < h16.size(); i++) h16[i] = (uint16_t)pts[i];
    DBuf<uint16_t> dp(h16.size()); DBuf<uint64_t> ddb(words), dref(words);
    CU(dp.up(h16.data(), h16.size()));
    launch_db_build_spiral(ddb.p, dp.p, (int)nu1, (int)nu2, (uint32_t)p_db, 0, total_n, 0); CHECK_LAUNCH();
    launch_db_to_reference(dref.p, ddb.p, (size_t)1 << nu1, ((size_t)1 << nu2) * kN2, 0); CHECK_LAUNCH();
    CU(dref.down(B, words));
    return SB200_OK;
}
extern "C" int sb200_reorientCiphertexts(uint64_t *out, const uint64_t *inp, size_t dim0, size_t n1_padded) {
    NEED_DEVICE();
    if (n1_padded != 4) return fail(SB200_ERR_ARG, "reorientCiphertexts: n1_padded must be 4");
    DBuf<uint32_t> din; DBuf<uint64_t> dout(dim0 * 2 * 4 * kN);
    TRY(up_ntt(din, inp, dim0 * kN1 * 2));
    launch_reorient_query(dout.p, din.p, dim0, 0); CHECK_LAUNCH();
    CU(dout.down(out, dim0 * 2 * 4 * kN));
    return SB200_OK;
}
extern "C" int sb200_multiplyQueryByDatabase(uint64_t *out, const uint64_t *reoriented, const uint64_t *database,
                                             size_t dim0, size_t num_per) {
    NEED_DEVICE();
    const size_t qwords = dim0 * 2 * 4 * kN, dbwords = dim0 * num_per * 4 * kN, opolys = num_per * 6;
    DBuf<uint64_t> dq(qwords), dref(dbwords), ddb(dbwords); DBuf<uint32_t> dout(opolys * PLW);
    CU(dq.up(reoriented, qwords)); CU(dref.up(database, dbwords));
    launch_db_from_reference(ddb.p, dref.p, dim0, num_per * kN2, 0, kN, 0); CHECK_LAUNCH();
    launch_scan_spiral(dout.p, dq.p, ddb.p, dim0, num_per, 0); CHECK_LAUNCH();
    return down_ntt(out, dout.p, opolys);
}
// `count` reoriented queries against one database in ONE pass (tensor-core path); outputs as multiplyQueryByDatabase's
extern "C" int sb200_multiplyQueryByDatabase_batched(uint64_t *const *out, const uint64_t *const *reoriented, int count,
                                                     const uint64_t *database, size_t dim0, size_t num_per) {
    NEED_DEVICE();
    if (!out || !reoriented || count < 1 || count > 16) return fail(SB200_ERR_ARG, "multiplyQueryByDatabase_batched: 1 <= count <= 16");
    if (!tc_shape_ok(dim0, num_per)) return fail(SB200_ERR_ARG, "multiplyQueryByDatabase_batched: needs 2*dim0 and 2*num_per to be multiples of 128");
    const size_t qwords = dim0 * 2 * 4 * kN, dbwords = dim0 * num_per * 4 * kN, opolys = num_per * 6;
    DBuf<uint64_t> dq(qwords), dref(dbwords), ddb(dbwords); DBuf<uint8_t> dtc(dbwords * 8), qtc(tc_query_bytes(dim0, count));
    DBuf<uint32_t> dout((size_t)count * opolys * PLW), dt1(tc_scratch_bytes(num_per, count) / 4);
    CU(dref.up(database, dbwords));
    launch_db_from_reference(ddb.p, dref.p, dim0, num_per * kN2, 0, kN, 0); CHECK_LAUNCH();
    launch_db_to_tc(dtc.p, ddb.p, dim0, num_per, 0); CHECK_LAUNCH();
    CU(cudaMemset(qtc.p, 0, qtc.n));
    uint32_t *o[16];
    for (int b = 0; b < count; b++) {
        CU(dq.up(reoriented[b], qwords));
        launch_query_to_tc(qtc.p, dq.p, b, count, dim0, 0); CHECK_LAUNCH();
        CU(cudaDeviceSynchronize());
        o[b] = dout.p + (size_t)b * opolys * PLW;
    }
    if (launch_scan_tc(o, count, count, qtc.p, dtc.p, dim0, num_per, dt1.p, 0)) return fail(SB200_ERR_CUDA, "multiplyQueryByDatabase_batched: launch failed");
    CHECK_LAUNCH();
    for (int b = 0; b < count; b++) TRY(down_ntt(out[b], o[b], opolys));
    return SB200_OK;
}
extern "C" int sb200_nttInvAndCrtLiftCiphertexts(uint64_t *cts_raw, const uint64_t *scratch, size_t num_per) {
    return sb200_from_ntt(cts_raw, scratch, num_per * 6);
}
extern "C" int sb200_foldOneFurtherDimension(size_t cur_dim, size_t num_per, const uint64_t *q, const uint64_t *q_neg,
                                             uint64_t *cts_raw, uint32_t t_gsw) {
    NEED_DEVICE();
    const int rm = 3 * 3 * (int)t_gsw;
    const size_t stride = (size_t)rm * 2 * kN;              // reference stride between dimensions (n1*m2*crt_count*poly_len)
    DBuf<uint64_t> dq((size_t)rm * kN), dqn((size_t)rm * kN), dcts(2 * num_per * 6 * kN);
    DBuf<uint32_t> q_dev((size_t)rm * PLW), qn_dev((size_t)rm * PLW), scratch(fold_scratch_words(num_per, (int)t_gsw));
    CU(dq.up(q + cur_dim * stride, (size_t)rm * kN)); CU(dqn.up(q_neg + cur_dim * stride, (size_t)rm * kN));
    CU(dcts.up(cts_raw, 2 * num_per * 6 * kN));
    launch_unreorient_q(q_dev.p, dq.p, rm, 0); launch_unreorient_q(qn_dev.p, dqn.p, rm, 0); CHECK_LAUNCH();
    launch_fold_round(dcts.p, num_per, q_dev.p, qn_dev.p, (int)t_gsw, scratch.p, 0); CHECK_LAUNCH();
    CU(dcts.down(cts_raw, num_per * 6 * kN));
    return SB200_OK;
}
extern "C" int sb200_split_and_crt(uint64_t *out, const uint64_t *in_raw, size_t num_per, uint32_t t_gsw) {
    NEED_DEVICE();
    // the decomposition kernel of the fold, run on num_per ciphertexts; scratch layout == reference's (i, m, c, n, z)
    DBuf<uint64_t> dcts(num_per * 6 * kN); DBuf<uint32_t> scratch(num_per * 3 * t_gsw * 2 * PLW);
    CU(dcts.up(in_raw, num_per * 6 * kN));
    launch_fold_decomp_only(scratch.p, dcts.p, num_per, (int)t_gsw, 0); CHECK_LAUNCH();
    return down_ntt(out, scratch.p, num_per * 3 * t_gsw * 2);
}

extern "C" int sb200_expandImproved(uint64_t *cv, size_t g, uint32_t t_exp, const uint64_t *W_left, const uint64_t *W_right,
                                    uint32_t t_exp_right, size_t max_bits_right, size_t stopround) {
    NEED_DEVICE();
    ExpandPlan plan{(int)g, (int)t_exp, (int)t_exp_right, (int)stopround, (int)max_bits_right};
    const size_t ncts = (size_t)1 << g;
    const size_t n_right = stopround > 0 ? stopround + 1 : g;
    std::vector<int> list(expand_active_total(plan)), offs(g), cnt(g);
    const int maxcnt = expand_build_lists(plan, list.data(), offs.data(), cnt.data());
    const int tmax = plan.t_left > plan.t_right ? plan.t_left : plan.t_right;
    DBuf<uint32_t> dcv, dWl, dWr, neg1(2 * g * PLW), c1((size_t)maxcnt * PLW), ginv((size_t)maxcnt * tmax * PLW);
    DBuf<uint64_t> c0((size_t)maxcnt * kN); DBuf<int> dlist(list.size());
    TRY(up_ntt(dcv, cv, ncts * 2)); TRY(up_ntt(dWl, W_left, g * 2 * t_exp)); TRY(up_ntt(dWr, W_right, n_right * 2 * t_exp_right));
    CU(dlist.up(list.data(), list.size()));
    build_neg1(neg1.p, (int)g, 0);
    std::vector<uint16_t> hperm(g * kN); build_automorph_perms(hperm.data(), (int)g);
    DBuf<uint16_t> dperm(hperm.size()); CU(dperm.up(hperm.data(), hperm.size()));
    launch_expand(dcv.p, plan, dWl.p, dWr.p, neg1.p, dperm.p, c0.p, c1.p, ginv.p, dlist.p, offs.data(), cnt.data(), 0); CHECK_LAUNCH();
    return down_ntt(cv, dcv.p, ncts * 2);
}
extern "C" int sb200_scalToMat(uint64_t *out_reg, const uint64_t *cv, const uint64_t *W, uint32_t t_conv) {
    NEED_DEVICE();
    DBuf<uint32_t> dcv, dW, dout(6 * PLW), sntt((size_t)t_conv * PLW); DBuf<uint64_t> sraw(kN); DBuf<int> idx(2);
    TRY(up_ntt(dcv, cv, 2)); TRY(up_ntt(dW, W, 3 * 2 * (size_t)t_conv));
    const int h[2] = {0, 0};
    CU(idx.up(h, 2));
    launch_scal_to_mat_ntt(dout.p, dcv.p, idx.p, idx.p + 1, 1, dW.p, (int)t_conv, sraw.p, sntt.p, 0); CHECK_LAUNCH();
    return down_ntt(out_reg, dout.p, 6);
}
extern "C" int sb200_regevToGSW(uint64_t *out, const uint64_t *cv_v, uint32_t t_conv, uint32_t t, const uint64_t *W, const uint64_t *V) {
    NEED_DEVICE();
    const int nbits = (int)t;
    DBuf<uint32_t> dcv, dW, dV, dout((size_t)3 * 3 * t * PLW), sntt((size_t)2 * t_conv * nbits * PLW);
    DBuf<uint64_t> sraw((size_t)2 * nbits * kN); DBuf<int> ct_idx(nbits), poly_idx(2 * nbits);
    TRY(up_ntt(dcv, cv_v, (size_t)t * 2)); TRY(up_ntt(dW, W, 3 * 2 * (size_t)t_conv)); TRY(up_ntt(dV, V, 3 * 2 * (size_t)t_conv));
    std::vector<int> hc(nbits), hp(2 * nbits);
    for (int b = 0; b < nbits; b++) { hc[b] = b; hp[b] = 2 * b; hp[nbits + b] = 2 * b + 1; }
    CU(ct_idx.up(hc.data(), nbits)); CU(poly_idx.up(hp.data(), 2 * nbits));
    launch_regev_to_gsw(dout.p, nullptr, dcv.p, ct_idx.p, poly_idx.p, 1, (int)t, dW.p, dV.p, (int)t_conv, sraw.p, sntt.p, 0); CHECK_LAUNCH();
    return down_ntt(out, dout.p, (size_t)3 * 3 * t);
}

// ---------------------------------------------------------------------------------------------
// tier 3: resident server
// ---------------------------------------------------------------------------------------------
// A query is ~60 dependent kernel launches on fixed buffers: each stage is captured once into a CUDA
// graph and replayed, which removes the per-launch gaps of the latency-bound expansion / fold chains.
// SB200_NO_GRAPH=1 falls back to plain launches (same kernels, same results).
namespace {
struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    int launches = 0;
    const void *key0 = nullptr, *key1 = nullptr;
    ~GraphSlot() { if (exec) cudaGraphExecDestroy(exec); }
};
bool graphs_enabled() {
    static int v = -1;
    if (v < 0) { const char *e = getenv("SB200_NO_GRAPH"); v = (e && *e == '1') ? 0 : 1; }
    return v == 1;
}
// sb200_server_prepare: build (capture + instantiate) the graphs of a query without running anything
static thread_local bool tl_prepare_only = false;
template <typename F>
int run_stage(GraphSlot &slot, cudaStream_t st, const void *k0, const void *k1, F &&body) {
    if (!graphs_enabled()) { if (!tl_prepare_only) { body(st); CHECK_LAUNCH(); } return SB200_OK; }
    {   // the caller is capturing (a whole-query graph around this stage, or a user's own capture): enqueue the stage's kernels
        // into that capture - never this stage's own executable graph, which would become a child-graph node
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(st, &cs);
        if (cs != cudaStreamCaptureStatusNone) { body(st); CHECK_LAUNCH(); return SB200_OK; }
    }
    if (slot.exec && (slot.key0 != k0 || slot.key1 != k1)) { cudaGraphExecDestroy(slot.exec); slot.exec = nullptr; }
    if (!slot.exec) {
        const uint64_t before = launch_count();
        cudaGraph_t graph = nullptr;
        CU(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        body(st);
        cudaError_t e = cudaStreamEndCapture(st, &graph);
        if (e != cudaSuccess || !graph) { cudaGetLastError(); return fail(SB200_ERR_CUDA, "stream capture failed: %s", cudaGetErrorString(e)); }
        slot.launches = (int)(launch_count() - before);
        count_launch(-slot.launches);                       // capture enqueues nothing; replays are counted below
        // node priorities (LaunchPriority scopes during capture) only count when the graph is instantiated with this flag; without
        // it every node runs at the priority of the stream the graph is launched into
        e = cudaGraphInstantiateWithFlags(&slot.exec, graph, cudaGraphInstantiateFlagUseNodePriority);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { slot.exec = nullptr; return fail(SB200_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e)); }
        slot.key0 = k0; slot.key1 = k1;
    }
    if (tl_prepare_only) { CU(cudaGraphUpload(slot.exec, st)); return SB200_OK; }       // the first launch then has nothing left to set up
    CU(cudaGraphLaunch(slot.exec, st));
    count_launch(slot.launches);
    return SB200_OK;
}
}  // namespace

struct sb200_server {
    sb200_params prm;
    int device = 0, rank = 0, world = 1, log_world = 0;
    size_t dim0 = 0, num_per = 0, local_num_per = 0;
    size_t g = 0, stopround = 0;
    ExpandPlan plan{};
    std::vector<int> offs, cnt;
    int maxcnt = 0, tmax = 0;
    bool have_db = false, have_params = false;
    size_t z_slices = 0;                                // 0 / 2048: explicit database; an implicit (--random-data) one holds fewer slices
    const sb200_server *db_owner = nullptr;           // views (sb200_server_create_view) scan another server's resident database
    // device memory
    DBuf<uint64_t> db;                                  // scan layout shard
    DBuf<uint8_t> db_tc, q_tc;                          // tensor-core path (sb200_server_enable_tc): limb-tile database, batched query tiles
    DBuf<uint32_t> tc_t1;                               // ... and the tile-order scan results of one pass
    int tc_capacity = 0;
    bool tc_only = false;                               // sb200_server_tc_only: the limb-tile copy is the only one, every scan runs on k_scan_tc
    DBuf<uint32_t> W_left, W_right, W_conv, V_conv, neg1;
    DBuf<uint64_t> q_stage;                             // uploaded query (ref-NTT)
    DBuf<uint8_t> q_wire;                               // uploaded query in its wire form (wire_kernels.cu), header included
    uint32_t wire_kind = 0;                             // 0: q_stage holds the query; 1 / 2: q_wire does (seeded / full)
    DBuf<uint32_t> cv, c1, ginv, conv_ntt, gsw, scan_out, fold_scratch;
    DBuf<uint64_t> c0, conv_raw, query, cts, resp, final_ct;
    DBuf<int> lists, ct_idx_first, poly_idx_first, ct_idx_bits, poly_idx_bits;
    // even / odd chains of the expansion tree (stopround > 0): own lists, the odd chain its own scratch (it runs on aux_stream)
    bool split_chains = false;
    DBuf<int> lists_e, lists_o;
    std::vector<int> offs_e, cnt_e, offs_o, cnt_o;
    DBuf<uint32_t> c1_o, ginv_o;
    DBuf<uint64_t> c0_o;
    DBuf<uint16_t> perms;
    GraphSlot g_convert, g_convert_wire[2], g_lift_fold, g_tail;
    // split chains as two independent graphs (even chain on the caller's stream, odd chain + RegevToGSW on aux_stream, joined
    // only before the folds); [0] in-memory query, [1] / [2] wire kinds
    GraphSlot g_even[3], g_odd[3];
    DBuf<uint32_t> cv_o;                                // the odd chain's own ciphertext array
    bool join_pending = false;                          // ev_join not yet waited for on the main stream
    // sharded expansion (world > 1, peers connected): this rank expands and converts only the first-dimension ciphertexts
    // j = rank (mod world) and stores them into every rank's query buffer (fused all-gather)
    DBuf<int> lists_es, ct_idx_first_s, poly_idx_first_s;
    std::vector<int> offs_es, cnt_es;
    // ... and of the odd chain only the ancestors of the GSW bits b = rank (mod world); its RegevToGSW columns go to every rank
    DBuf<int> lists_os, ct_idx_bits_s, poly_idx_bits_s, bit_ids_s;
    std::vector<int> offs_os, cnt_os;
    int nbits_local = 0;
    std::vector<void *> peer_query, peer_xb, peer_gsw;  // every rank's query buffer / exchange header / GSW buffer as seen from here
    bool shard_eligible = false, query_sharded = false, gsw_wait_pending = false;
    // resident drop-in (sb200_resident_*): which HOST buffers of the reference harness currently live in this server's device
    // buffers instead (keys are the harness's own pointers; the host copies are stale until a leaf downloads them)
    const void *res_query_key = nullptr, *res_scan_key = nullptr, *res_cts_key = nullptr;
    std::vector<std::pair<const void *, DBuf<uint32_t> *>> res_gsw;     // reorient_Q outputs -> dev-NTT GSW ciphertexts
    DBuf<uint32_t> res_cts_in;                                          // staging of the 2^nu1 ScalToMat outputs (dev-NTT)
    // stream = NULL means "the legacy default stream", which cannot be captured: such calls run on this
    // BLOCKING stream instead (implicitly ordered with legacy-default-stream work, e.g. torch's default stream)
    cudaStream_t own_stream = nullptr, aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    DBuf<uint64_t> conv_raw2;
    DBuf<uint32_t> conv_ntt2;
    // peer-memory exchange (xchg_kernels.cu)
    DBuf<uint8_t> xchg;                       // this rank's XchgBuf (+ slots)
    DBuf<unsigned int> xchg_state;            // [0] epoch, [1] error
    DBuf<unsigned int *> xchg_acks;           // rank 0: device array of pointers to every rank's ack word
    DBuf<uint64_t> gathered;                  // rank 0: private copy of the world surviving ciphertexts
    void *xchg_target = nullptr;              // rank 0's XchgBuf as seen from this rank
    std::vector<void *> ipc_opened;
    bool xchg_connected = false;
    GraphSlot g_xchg;
    ~sb200_server() {
        for (auto &e : res_gsw) delete e.second;
        if (own_stream) cudaStreamDestroy(own_stream);
        if (aux_stream) cudaStreamDestroy(aux_stream);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        for (void *p : ipc_opened) cudaIpcCloseMemHandle(p);
    }
};
static inline cudaStream_t ES(sb200_server *s, void *stream) { return stream ? (cudaStream_t)stream : s->own_stream; }

static size_t ceil_log2(size_t x) { size_t g = 0; while (((size_t)1 << g) < x) g++; return g; }

extern "C" int sb200_server_create(sb200_server **out, const sb200_params *prm, int device, int rank, int world) {
    if (!out || !prm) return fail(SB200_ERR_ARG, "server_create: null argument");
    if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world) return fail(SB200_ERR_ARG, "server_create: world must be a power of two and 0 <= rank < world");
    if (((size_t)1 << prm->nu2) < (size_t)world) return fail(SB200_ERR_ARG, "server_create: 2^nu2 < world");
    if (prm->t_gsw == 0 || prm->t_conv == 0 || prm->t_exp == 0 || prm->t_exp_right == 0) return fail(SB200_ERR_ARG, "server_create: zero gadget length");
    if (sb200_arb_qprime(prm->qp_bits) == 0) return fail(SB200_ERR_ARG, "server_create: no response modulus for qp_bits = %u (14..36)", prm->qp_bits);
    if (prm->p_db == 0 || prm->p_db > 65536) return fail(SB200_ERR_ARG, "server_create: p_db must be in [1, 65536]");
    if (prm->nu1 > 11 || ((size_t)1 << prm->nu1) + (size_t)prm->t_gsw * prm->nu2 > (size_t)kN)
        return fail(SB200_ERR_ARG, "server_create: 2^nu1 + t_GSW*nu2 exceeds the 2048 slots of a packed query");
    int rc = sb200_init(device);
    if (rc) return rc;
    sb200_server *s = new sb200_server();
    s->prm = *prm; s->device = device; s->rank = rank; s->world = world; s->log_world = (int)ceil_log2((size_t)world);
    s->dim0 = (size_t)1 << prm->nu1; s->num_per = (size_t)1 << prm->nu2; s->local_num_per = s->num_per / world;
    // expansion shape exactly as runConversionImproved derives it (reference src/spiral.cpp:2076-2085, qe_rest == 0)
    const size_t ell = prm->t_gsw, nbits = ell * prm->nu2;
    s->g = ceil_log2(nbits + s->dim0);
    s->stopround = ceil_log2(nbits);
    if (nbits > s->dim0) s->stopround = 0;
    s->plan = ExpandPlan{(int)s->g, (int)prm->t_exp, (int)prm->t_exp_right, (int)s->stopround, (int)nbits};
    std::vector<int> list(expand_active_total(s->plan));
    s->offs.resize(s->g); s->cnt.resize(s->g);
    s->maxcnt = expand_build_lists(s->plan, list.data(), s->offs.data(), s->cnt.data());
    s->tmax = s->plan.t_left > s->plan.t_right ? s->plan.t_left : s->plan.t_right;

    const size_t ncts = (size_t)1 << s->g, m2 = 3 * ell, n_right = s->stopround > 0 ? s->stopround + 1 : s->g;
    const size_t conv_cols = std::max(s->dim0, 2 * nbits);
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    A(s->W_left.alloc(s->g * 2 * prm->t_exp * PLW)); A(s->W_right.alloc(n_right * 2 * prm->t_exp_right * PLW));
    A(s->W_conv.alloc(3 * 2 * prm->t_conv * PLW)); A(s->V_conv.alloc(3 * 2 * prm->t_conv * PLW)); A(s->neg1.alloc(2 * s->g * PLW));
    A(s->q_stage.alloc(2 * PLW)); A(s->q_wire.alloc(kWireHeaderBytes + 2 * kWireRowBytes + 8)); A(s->cv.alloc(ncts * 2 * PLW)); A(s->c1.alloc((size_t)s->maxcnt * PLW));
    A(s->ginv.alloc(expand_ginv_polys(s->plan, s->cnt.data()) * PLW)); A(s->c0.alloc((size_t)s->maxcnt * kN));
    A(s->conv_raw.alloc(conv_cols * kN)); A(s->conv_ntt.alloc(conv_cols * prm->t_conv * PLW));
    A(s->gsw.alloc(prm->nu2 * 3 * m2 * PLW));
    A(s->query.alloc(s->dim0 * 2 * 4 * kN)); A(s->scan_out.alloc(s->local_num_per * 6 * PLW));
    const size_t cts_n = std::max(s->local_num_per, (size_t)world);
    A(s->cts.alloc(cts_n * 6 * kN)); A(s->final_ct.alloc(6 * kN)); A(s->resp.alloc(6 * kN));
    const size_t fold_np = std::max(s->local_num_per / 2, (size_t)world / 2);
    A(s->fold_scratch.alloc(fold_scratch_words(fold_np ? fold_np : 1, (int)ell)));
    A(s->lists.alloc(list.size())); A(s->ct_idx_first.alloc(s->dim0)); A(s->poly_idx_first.alloc(s->dim0));
    A(s->ct_idx_bits.alloc(nbits ? nbits : 1)); A(s->poly_idx_bits.alloc(nbits ? 2 * nbits : 1));
    if (e != cudaSuccess) { delete s; return fail(SB200_ERR_CUDA, "server_create: device allocation failed: %s", cudaGetErrorString(e)); }
    // index lists (reorderFromStopround, reference src/spiral.cpp:2027-2036)
    std::vector<int> cf(s->dim0), pf(s->dim0), cb(nbits), pb(2 * nbits);
    for (size_t j = 0; j < s->dim0; j++) { cf[j] = (int)(s->stopround ? 2 * j : j); pf[j] = 2 * cf[j]; }
    for (size_t b = 0; b < nbits; b++) { cb[b] = (int)(s->stopround ? 2 * b + 1 : s->dim0 + b); pb[b] = 2 * cb[b]; pb[nbits + b] = 2 * cb[b] + 1; }
    A(s->lists.up(list.data(), list.size())); A(s->ct_idx_first.up(cf.data(), cf.size())); A(s->poly_idx_first.up(pf.data(), pf.size()));
    if (nbits) { A(s->ct_idx_bits.up(cb.data(), cb.size())); A(s->poly_idx_bits.up(pb.data(), pb.size())); }
    if (s->stopround > 0 && s->g > 1) {
        std::vector<int> le(list.size()), lo(list.size());
        s->offs_e.resize(s->g); s->cnt_e.resize(s->g); s->offs_o.resize(s->g); s->cnt_o.resize(s->g);
        const int max_odd = expand_split_lists(s->plan, list.data(), s->offs.data(), s->cnt.data(), le.data(), s->offs_e.data(), s->cnt_e.data(),
                                               lo.data(), s->offs_o.data(), s->cnt_o.data());
        A(s->lists_e.alloc(le.size())); A(s->lists_o.alloc(lo.size()));
        A(s->lists_e.up(le.data(), le.size())); A(s->lists_o.up(lo.data(), lo.size()));
        A(s->c0_o.alloc((size_t)std::max(max_odd, 1) * kN)); A(s->c1_o.alloc((size_t)std::max(max_odd, 1) * PLW));
        A(s->ginv_o.alloc((size_t)std::max(max_odd, 1) * prm->t_exp_right * PLW));
        A(s->cv_o.alloc(((size_t)2 << s->stopround) * 2 * PLW));
        s->split_chains = true;
        // this rank's share of the even chain: output i = 2 j'' of round r is kept when j'' = rank modulo min(world, 2^r) - the
        // ancestors of the first-dimension ciphertexts j = rank (mod world) and nothing else
        if (world > 1 && (s->dim0 / world) % 8 == 0 && s->dim0 / world >= 8 && prm->t_conv == 4) {
            std::vector<int> ls;
            s->offs_es.resize(s->g); s->cnt_es.resize(s->g);
            for (size_t r = 0; r < s->g; r++) {
                s->offs_es[r] = (int)ls.size();
                const size_t m = std::min((size_t)world, (size_t)1 << r);
                for (int k = 0; k < s->cnt_e[r]; k++) {
                    const int i = le[s->offs_e[r] + k];
                    if (((size_t)(i / 2)) % m == (size_t)rank % m) ls.push_back(i);
                }
                s->cnt_es[r] = (int)ls.size() - s->offs_es[r];
            }
            const size_t cl = s->dim0 / world;
            std::vector<int> cfs(cl), pfs(cl);
            for (size_t jl = 0; jl < cl; jl++) { cfs[jl] = (int)(2 * ((size_t)rank + (size_t)world * jl)); pfs[jl] = 2 * cfs[jl]; }
            A(s->lists_es.alloc(ls.size())); A(s->lists_es.up(ls.data(), ls.size()));
            A(s->ct_idx_first_s.alloc(cl)); A(s->ct_idx_first_s.up(cfs.data(), cl));
            A(s->poly_idx_first_s.alloc(cl)); A(s->poly_idx_first_s.up(pfs.data(), cl));
            std::vector<int> lo_s;
            s->offs_os.resize(s->g); s->cnt_os.resize(s->g);
            for (size_t r = 0; r < s->g; r++) {
                s->offs_os[r] = (int)lo_s.size();
                const size_t m = std::min((size_t)world, (size_t)1 << r);
                for (int k = 0; k < s->cnt_o[r]; k++) {
                    const int i = lo[s->offs_o[r] + k];
                    if (((size_t)(i / 2)) % m == (size_t)rank % m) lo_s.push_back(i);
                }
                s->cnt_os[r] = (int)lo_s.size() - s->offs_os[r];
            }
            std::vector<int> cbs, bids;
            for (size_t b = (size_t)rank; b < nbits; b += (size_t)world) { bids.push_back((int)b); cbs.push_back((int)(2 * b + 1)); }
            s->nbits_local = (int)bids.size();
            std::vector<int> pbs(2 * bids.size());
            for (size_t l = 0; l < bids.size(); l++) { pbs[l] = 2 * cbs[l]; pbs[bids.size() + l] = 2 * cbs[l] + 1; }
            A(s->lists_os.alloc(std::max(lo_s.size(), (size_t)1))); A(s->lists_os.up(lo_s.data(), lo_s.size()));
            A(s->ct_idx_bits_s.alloc(std::max(cbs.size(), (size_t)1))); A(s->ct_idx_bits_s.up(cbs.data(), cbs.size()));
            A(s->poly_idx_bits_s.alloc(std::max(pbs.size(), (size_t)1))); A(s->poly_idx_bits_s.up(pbs.data(), pbs.size()));
            A(s->bit_ids_s.alloc(std::max(bids.size(), (size_t)1))); A(s->bit_ids_s.up(bids.data(), bids.size()));
            s->shard_eligible = nbits >= (size_t)world;
        }
    }
    { std::vector<uint16_t> hperm(s->g * kN); build_automorph_perms(hperm.data(), (int)s->g);
      A(s->perms.alloc(hperm.size())); A(s->perms.up(hperm.data(), hperm.size())); }
    build_neg1(s->neg1.p, (int)s->g, 0);
    A(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamDefault));
    A(cudaStreamCreateWithFlags(&s->aux_stream, cudaStreamNonBlocking));
    A(cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming)); A(cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming));
    A(s->conv_raw2.alloc(std::max((size_t)1, 2 * nbits) * kN)); A(s->conv_ntt2.alloc(std::max((size_t)1, 2 * nbits) * prm->t_conv * PLW));
    if (world > 1) {
        if (world > 16) { delete s; return fail(SB200_ERR_ARG, "server_create: world > 16 not supported by the peer exchange"); }
        A(s->xchg.alloc(xchg_buffer_bytes(world))); A(s->xchg_state.alloc(2)); A(s->xchg_acks.alloc(world)); A(s->gathered.alloc((size_t)world * 6 * kN));
        if (e == cudaSuccess) { A(cudaMemset(s->xchg.p, 0, xchg_buffer_bytes(world))); A(cudaMemset(s->xchg_state.p, 0, 2 * sizeof(unsigned int))); }
    }
    A(cudaDeviceSynchronize());
    if (e != cudaSuccess) { delete s; return fail(SB200_ERR_CUDA, "server_create: setup failed: %s", cudaGetErrorString(e)); }
    *out = s;
    return SB200_OK;
}
extern "C" void sb200_server_destroy(sb200_server *s) { delete s; }
// A second server over the SAME resident database (own workspaces, own public parameters, own streams / graphs):
// one per concurrent client, so several queries can be in flight on one GPU.  The parent must outlive its views.
extern "C" int sb200_server_create_view(sb200_server **out, sb200_server *parent) {
    if (!out || !parent) return fail(SB200_ERR_ARG, "create_view: null argument");
    const sb200_server *owner = parent->db_owner ? parent->db_owner : parent;
    int rc = sb200_server_create(out, &parent->prm, parent->device, parent->rank, parent->world);
    if (rc) return rc;
    (*out)->db_owner = owner;
    return SB200_OK;
}

static inline const uint64_t *server_db(const sb200_server *s) { return s->db_owner ? s->db_owner->db.p : s->db.p; }
static inline bool server_has_db(const sb200_server *s) { return s->db_owner ? s->db_owner->have_db : s->have_db; }
static inline size_t server_z_slices(const sb200_server *s) { const sb200_server *o = s->db_owner ? s->db_owner : s; return o->z_slices ? o->z_slices : (size_t)kN; }
static int server_alloc_db(sb200_server *s, size_t z_slices = 0) {
    if (s->db_owner) return fail(SB200_ERR_STATE, "this server is a view: load the database through its parent");
    const size_t want = z_slices ? z_slices : (size_t)kN;
    if (s->db_tc.p) {                                   // a (re)load makes the limb-tile copy stale: drop it (enable again afterwards)
        cudaFree(s->db_tc.p); s->db_tc.p = nullptr;
        s->tc_capacity = 0;
        if (s->tc_only) { s->tc_only = false; s->have_db = false; }     // ... and it was the only copy
    }
    if (s->db.p && server_z_slices(s) != want) { cudaFree(s->db.p); s->db.p = nullptr; s->have_db = false; }
    s->z_slices = z_slices;
    if (s->db.p) return SB200_OK;
    CU(s->db.alloc(s->dim0 * s->local_num_per * 4 * want));
    return SB200_OK;
}
extern "C" int sb200_server_load_db_items(sb200_server *s, const uint16_t *pts, size_t item_begin, size_t item_count) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    CU(cudaSetDevice(s->device));
    TRY(server_alloc_db(s));
    const size_t total_local = s->dim0 * s->local_num_per;
    if (item_begin + item_count > total_local) return fail(SB200_ERR_ARG, "load_db_items: range exceeds the shard (%zu items)", total_local);
    const size_t chunk = 4096;                                           // items per staging buffer (64 MiB)
    DBuf<uint16_t> stage(std::min(chunk, item_count ? item_count : 1) * 4 * kN);
    uint32_t nu2_local = (uint32_t)ceil_log2(s->local_num_per);
    for (size_t o = 0; o < item_count; o += chunk) {
        const size_t n = std::min(chunk, item_count - o);
        CU(stage.up(pts + o * 4 * kN, n * 4 * kN));
        launch_db_build_spiral(s->db.p, stage.p, (int)s->prm.nu1, (int)nu2_local, (uint32_t)s->prm.p_db, item_begin + o, n, 0);
        CHECK_LAUNCH();
        CU(cudaDeviceSynchronize());
    }
    s->have_db = true;
    return SB200_OK;
}
static int server_load_db_reference_slices(sb200_server *s, const uint64_t *B, size_t slices) {
    if (!s || !B) return fail(SB200_ERR_ARG, "null argument");
    CU(cudaSetDevice(s->device));
    TRY(server_alloc_db(s, slices == (size_t)kN ? 0 : slices));
    // per z-slice the reference holds num_per rows (ii) of n2*dim0*n0 words; the shard takes rows ii = rank (mod world)
    const size_t row_words = kN2 * s->dim0 * kN0, zc = std::min((size_t)16, slices);
    DBuf<uint64_t> stage(zc * s->local_num_per * row_words);
    for (size_t z0 = 0; z0 < slices; z0 += zc) {
        if (s->world == 1) {
            CU(stage.up(B + z0 * s->num_per * row_words, zc * s->num_per * row_words));
        } else {
            for (size_t z = 0; z < zc; z++)
                CU(cudaMemcpy2D(stage.p + z * s->local_num_per * row_words, row_words * 8,
                                B + ((z0 + z) * s->num_per + s->rank) * row_words, (size_t)s->world * row_words * 8,
                                row_words * 8, s->local_num_per, cudaMemcpyHostToDevice));
        }
        launch_db_from_reference(s->db.p, stage.p, s->dim0, s->local_num_per * kN2, z0, zc, 0);
        CHECK_LAUNCH();
        CU(cudaDeviceSynchronize());
    }
    s->have_db = true;
    return SB200_OK;
}
extern "C" int sb200_server_load_db_reference(sb200_server *s, const uint64_t *B) { return server_load_db_reference_slices(s, B, kN); }
// The implicit database of the reference's --random-data mode (src/spiral.cpp:1032-1081, 1274-1282): B holds only `working_set`
// z-slices (a power of two <= 2048; dummyWorkingSet = min(2^25 / total_n, 2048)) and the scan reads slice z mod working_set
// (:647, the AVX-512 statement).  The algorithmic size of a scan stays the full 2^(nu1+nu2) x 4 x 2048 words.
extern "C" int sb200_server_load_db_implicit(sb200_server *s, const uint64_t *B_slices_host, size_t working_set) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (!working_set || working_set > (size_t)kN || (working_set & (working_set - 1))) return fail(SB200_ERR_ARG, "load_db_implicit: working_set must be a power of two <= 2048");
    if (s->db_tc.p) return fail(SB200_ERR_STATE, "load_db_implicit: the tensor-core copy was built for an explicit database");
    return server_load_db_reference_slices(s, B_slices_host, working_set);
}
// every record the same constant polynomial `value` (< p_db / 4) in coefficient 0 of its four polynomials, as load_db generates
// it (:1034-1072): every word of every slice is the same residue pair, written on the device
extern "C" int sb200_server_load_db_implicit_constant(sb200_server *s, uint64_t value, size_t working_set) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (!working_set || working_set > (size_t)kN || (working_set & (working_set - 1))) return fail(SB200_ERR_ARG, "load_db_implicit: working_set must be a power of two <= 2048");
    if (value >= s->prm.p_db / 2) return fail(SB200_ERR_ARG, "load_db_implicit_constant: value must be below p_db / 2");
    if (s->db_tc.p) return fail(SB200_ERR_STATE, "load_db_implicit: the tensor-core copy was built for an explicit database");
    CU(cudaSetDevice(s->device));
    TRY(server_alloc_db(s, working_set == (size_t)kN ? 0 : working_set));
    const uint64_t word = (value % kP) | ((value % kB) << 32);
    const size_t words = s->dim0 * s->local_num_per * 4 * working_set;
    std::vector<uint64_t> h(1 << 16, word);
    for (size_t o = 0; o < words; o += h.size()) CU(cudaMemcpy(s->db.p + o, h.data(), std::min(h.size(), words - o) * 8, cudaMemcpyHostToDevice));
    s->have_db = true;
    return SB200_OK;
}
extern "C" size_t sb200_server_db_slices(const sb200_server *s) { return s ? server_z_slices(s) : 0; }
extern "C" uint64_t *sb200_server_db_ptr(sb200_server *s) { return s ? const_cast<uint64_t *>(server_db(s)) : nullptr; }

static int server_up_ntt(sb200_server *s, DBuf<uint32_t> &dst, const uint64_t *host, size_t npolys) {
    (void)s;
    const size_t chunk = 1024;
    DBuf<uint64_t> tmp(std::min(chunk, npolys ? npolys : 1) * PLW);
    for (size_t o = 0; o < npolys; o += chunk) {
        const size_t n = std::min(chunk, npolys - o);
        CU(tmp.up(host + o * PLW, n * PLW));
        launch_ntt_u64_to_dev(dst.p + o * PLW, tmp.p, n, 0); CHECK_LAUNCH();
        CU(cudaDeviceSynchronize());
    }
    return SB200_OK;
}
// polynomial counts (2*2048 words each) of the four matrices sb200_server_set_public_params reads: a host that receives them
// from an untrusted client checks the sizes against these before handing the pointers over
extern "C" int sb200_server_public_param_polys(const sb200_server *s, size_t *out4) {
    if (!s || !out4) return fail(SB200_ERR_ARG, "server_public_param_polys: null argument");
    const size_t n_right = s->stopround > 0 ? s->stopround + 1 : s->g;
    out4[0] = s->g * 2 * s->prm.t_exp; out4[1] = n_right * 2 * s->prm.t_exp_right;
    out4[2] = out4[3] = 3 * 2 * (size_t)s->prm.t_conv;
    return SB200_OK;
}
extern "C" int sb200_server_set_public_params(sb200_server *s, const uint64_t *W_exp_left, const uint64_t *W_exp_right,
                                              const uint64_t *W_conv, const uint64_t *V_conv) {
    if (!s || !W_exp_left || !W_exp_right || !W_conv || !V_conv) return fail(SB200_ERR_ARG, "set_public_params: null argument");
    CU(cudaSetDevice(s->device));
    const size_t n_right = s->stopround > 0 ? s->stopround + 1 : s->g;
    TRY(server_up_ntt(s, s->W_left, W_exp_left, s->g * 2 * s->prm.t_exp));
    TRY(server_up_ntt(s, s->W_right, W_exp_right, n_right * 2 * s->prm.t_exp_right));
    TRY(server_up_ntt(s, s->W_conv, W_conv, 3 * 2 * (size_t)s->prm.t_conv));
    TRY(server_up_ntt(s, s->V_conv, V_conv, 3 * 2 * (size_t)s->prm.t_conv));
    s->have_params = true;
    return SB200_OK;
}

extern "C" int sb200_server_upload_query(sb200_server *s, const uint64_t *query_cv_host, void *stream) {
    if (!s || !query_cv_host) return fail(SB200_ERR_ARG, "upload_query: null argument");
    CU(cudaMemcpyAsync(s->q_stage.p, query_cv_host, 2 * PLW * sizeof(uint64_t), cudaMemcpyHostToDevice, ES(s, stream)));
    s->wire_kind = 0;
    return SB200_OK;      // the narrowing into cv[0] is the first node of the expand_and_convert stage
}
static int join_odd_chain(sb200_server *s, cudaStream_t st);
// expansion + conversion.  defer_join: the odd chain (GSW bits -> RegevToGSW) is only needed by the folds, so the one-call paths
// leave it running on the side stream across the scan and join in sb200_server_fold_local.
static int expand_and_convert_impl(sb200_server *s, void *stream, bool defer_join) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (!s->have_params) return fail(SB200_ERR_STATE, "expand_and_convert: public parameters not set");
    cudaStream_t main_st = ES(s, stream);
    if (s->split_chains) {
        // The expansion tree splits at its root into two independent chains: the even ciphertexts (first dimension: t_left digits,
        // all g rounds, then ScalToMat) and the odd ones (GSW bits: t_right = 56 digits per key switch, rounds 0..stopround, then
        // RegevToGSW).  Each is its own graph on its own stream, each starts from the uploaded query (the odd chain works on a
        // private ciphertext array); the scan waits for the even chain only, the folds join the odd chain.
        const int wk = (int)s->wire_kind;
        const bool shard = s->world > 1 && s->xchg_connected && s->shard_eligible;
        static const bool skip_odd = [] { const char *e = getenv("SB200_PROFILE_SKIP_ODD_CHAIN"); return e && *e == '1'; }();   // timing experiments only: answers are wrong
        CU(cudaEventRecord(s->ev_fork, main_st));
        CU(cudaStreamWaitEvent(s->aux_stream, s->ev_fork, 0));
        const int odd_end = (int)s->stopround + 1;
        // The odd chain's 56-digit rounds are thousands of NTT CTAs that sit ahead of the even chain's next kernels in the block
        // scheduler's queue and cost the even chain ~90 us.  Measured alternatives (profiles/r02_expansion_chains.md): strict launch
        // priorities starve the odd chain, which then spills into the scan; digit kernels confined to SB200_ODD_SLOTS CTA slots
        // per SM make the contention last longer (1.01 ms/query at 2 slots against 0.94 unlimited).  Default: unlimited.
        static const int odd_slots = [] { const char *e = getenv("SB200_ODD_SLOTS"); return e ? atoi(e) : 0; }();
        static const int odd_chunk = [] { const char *e = getenv("SB200_ODD_CHUNK"); return e ? atoi(e) : 740; }();   // ~one wave of NTT CTAs
        auto odd_chain = [&](cudaStream_t st) {
            LaunchPriority low(false);                  // no-op unless SB200_PRIO=1 (experiments)
            if (wk) launch_query_from_wire(s->cv_o.p, s->q_wire.p, s->wire_kind, st);
            else launch_ntt_u64_to_dev(s->cv_o.p, s->q_stage.p, 2, st);
            if (shard) {
                launch_expand(s->cv_o.p, s->plan, s->W_left.p, s->W_right.p, s->neg1.p, s->perms.p, s->c0_o.p, s->c1_o.p, s->ginv_o.p, s->lists_os.p,
                              s->offs_os.data(), s->cnt_os.data(), st, 0, odd_end, 1, 1, odd_slots * 148, odd_chunk);
                GswTargets tg{};
                tg.ntargets = s->world;
                for (int t = 0; t < s->world; t++) {
                    const int r = (s->rank + t) % s->world;
                    tg.gsw[t] = (uint32_t *)s->peer_gsw[r];
                    tg.flag[t] = xchg_gflag_ptr(s->peer_xb[r], s->rank);
                }
                tg.arrive = xchg_arrive_aux_ptr(s->xchg.p); tg.ack = xchg_ack_ptr(s->xchg.p);
                tg.epoch = s->xchg_state.p; tg.error = s->xchg_state.p + 1;
                launch_regev_to_gsw_sharded(tg, s->cv_o.p, s->ct_idx_bits_s.p, s->poly_idx_bits_s.p, s->bit_ids_s.p, s->nbits_local, (int)s->prm.nu2,
                                            (int)s->prm.t_gsw, s->W_conv.p, s->V_conv.p, (int)s->prm.t_conv, s->conv_raw2.p, s->conv_ntt2.p, st);
                return;
            }
            launch_expand(s->cv_o.p, s->plan, s->W_left.p, s->W_right.p, s->neg1.p, s->perms.p, s->c0_o.p, s->c1_o.p, s->ginv_o.p, s->lists_o.p,
                          s->offs_o.data(), s->cnt_o.data(), st, 0, odd_end, 1, 1, odd_slots * 148, odd_chunk);
            // no GSW negation on the resident path: the fold uses the CMux form (launch_fold_round_generic)
            launch_regev_to_gsw(s->gsw.p, nullptr, s->cv_o.p, s->ct_idx_bits.p, s->poly_idx_bits.p, (int)s->prm.nu2, (int)s->prm.t_gsw,
                                s->W_conv.p, s->V_conv.p, (int)s->prm.t_conv, s->conv_raw2.p, s->conv_ntt2.p, st);
        };
        // Also measured and not adopted: the chain's last rounds (or RegevToGSW alone) as a second graph that starts with the scan.
        // The scan keeps every register file full (8 CTAs x 8 K registers per SM; at 7 / 6 CTAs it takes 0.390 / 0.414 ms instead
        // of 0.360) and the lift's CTAs, launched early, take the slots that free up - the deferred work runs after the scan.
        if (!skip_odd) TRY(run_stage(s->g_odd[wk], s->aux_stream, shard ? (const void *)s->xchg.p : nullptr, nullptr, odd_chain));
        TRY(run_stage(s->g_even[wk], main_st, shard ? (const void *)s->xchg.p : nullptr, nullptr, [&](cudaStream_t st) {
            LaunchPriority high(true);                  // no-op unless SB200_PRIO=1 (experiments)
            if (wk) launch_query_from_wire(s->cv.p, s->q_wire.p, s->wire_kind, st);     // wire query -> cv[0]
            else launch_ntt_u64_to_dev(s->cv.p, s->q_stage.p, 2, st);                   // uploaded query (ref-NTT) -> cv[0]
            if (!shard) {
                launch_expand(s->cv.p, s->plan, s->W_left.p, s->W_right.p, s->neg1.p, s->perms.p, s->c0.p, s->c1.p, s->ginv.p, s->lists_e.p,
                              s->offs_e.data(), s->cnt_e.data(), st, 0, (int)s->g, 0, 1);
                LaunchPriority tail(true, 2);
                launch_scal_to_mat_reoriented(s->query.p, s->cv.p, s->ct_idx_first.p, s->poly_idx_first.p, s->dim0, s->W_conv.p,
                                              (int)s->prm.t_conv, s->conv_raw.p, s->conv_ntt.p, st);
                return;
            }
            launch_expand(s->cv.p, s->plan, s->W_left.p, s->W_right.p, s->neg1.p, s->perms.p, s->c0.p, s->c1.p, s->ginv.p, s->lists_es.p,
                          s->offs_es.data(), s->cnt_es.data(), st, 0, (int)s->g, 0, 1);
            ScalTargets tg{};
            tg.ntargets = s->world;
            for (int t = 0; t < s->world; t++) {                       // own buffer first, then the peers in ring order
                const int r = (s->rank + t) % s->world;
                tg.query[t] = (uint64_t *)s->peer_query[r];
                tg.flag[t] = xchg_qflag_ptr(s->peer_xb[r], s->rank);
            }
            tg.arrive = xchg_arrive_ptr(s->xchg.p, 0); tg.ack = xchg_ack_ptr(s->xchg.p);
            tg.epoch = s->xchg_state.p; tg.error = s->xchg_state.p + 1;
            launch_scal_to_mat_sharded(tg, s->cv.p, s->ct_idx_first_s.p, s->poly_idx_first_s.p, s->dim0, s->dim0 / s->world, s->rank, s->world,
                                       s->W_conv.p, s->conv_raw.p, s->conv_ntt.p, st);
        }));
        CU(cudaEventRecord(s->ev_join, s->aux_stream));
        s->query_sharded = shard;
        s->gsw_wait_pending = shard;
        if (defer_join) s->join_pending = true;
        else { s->join_pending = true; TRY(join_odd_chain(s, main_st)); }
        return SB200_OK;
    }
    s->query_sharded = false;
    GraphSlot &slot = s->wire_kind ? s->g_convert_wire[s->wire_kind - 1] : s->g_convert;
    return run_stage(slot, main_st, nullptr, nullptr, [&](cudaStream_t st) {
        if (s->wire_kind) launch_query_from_wire(s->cv.p, s->q_wire.p, s->wire_kind, st);     // wire query -> cv[0]
        else launch_ntt_u64_to_dev(s->cv.p, s->q_stage.p, 2, st);                             // uploaded query (ref-NTT) -> cv[0]
        // stopround == 0 (more GSW bits than first-dimension slots): one tree; RegevToGSW forks to the side stream after the last
        // round that touches its ciphertexts (a fork/join pair inside the captured graph)
        const int fork_round = s->stopround > 0 ? (int)s->stopround + 1 : (int)s->g;
        launch_expand(s->cv.p, s->plan, s->W_left.p, s->W_right.p, s->neg1.p, s->perms.p, s->c0.p, s->c1.p, s->ginv.p, s->lists.p,
                      s->offs.data(), s->cnt.data(), st, 0, fork_round);
        cudaEventRecord(s->ev_fork, st);
        cudaStreamWaitEvent(s->aux_stream, s->ev_fork, 0);
        // no GSW negation on the resident path: the fold uses the CMux form (launch_fold_round_generic)
        launch_regev_to_gsw(s->gsw.p, nullptr, s->cv.p, s->ct_idx_bits.p, s->poly_idx_bits.p, (int)s->prm.nu2, (int)s->prm.t_gsw,
                            s->W_conv.p, s->V_conv.p, (int)s->prm.t_conv, s->conv_raw2.p, s->conv_ntt2.p, s->aux_stream);
        cudaEventRecord(s->ev_join, s->aux_stream);
        launch_expand(s->cv.p, s->plan, s->W_left.p, s->W_right.p, s->neg1.p, s->perms.p, s->c0.p, s->c1.p, s->ginv.p, s->lists.p,
                      s->offs.data(), s->cnt.data(), st, fork_round, (int)s->g);
        launch_scal_to_mat_reoriented(s->query.p, s->cv.p, s->ct_idx_first.p, s->poly_idx_first.p, s->dim0, s->W_conv.p,
                                      (int)s->prm.t_conv, s->conv_raw.p, s->conv_ntt.p, st);
        cudaStreamWaitEvent(st, s->ev_join, 0);
    });
}
extern "C" int sb200_server_expand_and_convert(sb200_server *s, void *stream) { return expand_and_convert_impl(s, stream, false); }
extern "C" int sb200_server_expansion_sharded(const sb200_server *s) { return s && s->query_sharded; }
static int join_odd_chain(sb200_server *s, cudaStream_t st) {
    if (s->join_pending) { CU(cudaStreamWaitEvent(st, s->ev_join, 0)); s->join_pending = false; }
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(st, &cs);
    if (s->gsw_wait_pending && (!tl_prepare_only || cs != cudaStreamCaptureStatusNone)) {   // sharded conversion: every rank's GSW columns must have landed here
        launch_flag_wait(s->xchg.p, s->world, s->xchg_state.p, s->xchg_state.p + 1, 1, 0, st);
    }
    s->gsw_wait_pending = false;
    return SB200_OK;
}
static inline sb200_server *server_owner(sb200_server *s) { return s->db_owner ? const_cast<sb200_server *>(s->db_owner) : s; }
// first dimension of ONE server's converted query: the streaming kernel on the scan-layout database, or - when only the limb-tile
// copy is resident (sb200_server_tc_only) - a one-query pass of the tensor-core kernel
static int scan_one(sb200_server *s, cudaStream_t st) {
    sb200_server *owner = server_owner(s);
    if (owner->tc_only) {
        uint32_t *o[1] = {s->scan_out.p}; const uint64_t *q[1] = {s->query.p};
        launch_queries_to_tc(owner->q_tc.p, q, 1, 0, owner->tc_capacity, owner->dim0, st);
        if (launch_scan_tc(o, 1, owner->tc_capacity, owner->q_tc.p, owner->db_tc.p, owner->dim0, owner->local_num_per, owner->tc_t1.p, st))
            return fail(SB200_ERR_CUDA, "scan (tensor-core copy only): launch failed");
    } else {
        LaunchPriority first(true, 3);
        launch_scan_spiral(s->scan_out.p, s->query.p, server_db(s), s->dim0, s->local_num_per, st, server_z_slices(s));
    }
    CHECK_LAUNCH();
    return SB200_OK;
}
extern "C" int sb200_server_scan(sb200_server *s, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (!server_has_db(s)) return fail(SB200_ERR_STATE, "scan: database not loaded");
    // sharded expansion: every rank's slice of the query must have landed in this rank's buffer
    if (s->query_sharded) launch_flag_wait(s->xchg.p, s->world, s->xchg_state.p, s->xchg_state.p + 1, 0, 0, ES(s, stream));
    return scan_one(s, ES(s, stream));
}
// One pass over the database for `count` (2 or 4) servers that share it (a parent and its views): every server's own
// converted query is scanned, every server's own ciphertext buffer is filled.  Launched on `stream`.
extern "C" int sb200_server_scan_batched(sb200_server *const *servers, int count, void *stream) {
    if (!servers || (count != 2 && count != 4)) return fail(SB200_ERR_ARG, "scan_batched: count must be 2 or 4");
    sb200_server *s0 = servers[0];
    if (!s0 || !server_has_db(s0)) return fail(SB200_ERR_STATE, "scan_batched: database not loaded");
    if (server_z_slices(s0) != (size_t)kN) return fail(SB200_ERR_STATE, "scan_batched: needs an explicit database");
    if (server_owner(s0)->tc_only) return fail(SB200_ERR_STATE, "scan_batched: the scan-layout copy was released (sb200_server_tc_only): use sb200_server_scan_batched_tc");
    const uint64_t *q[4]; uint32_t *o[4];
    for (int b = 0; b < count; b++) {
        if (!servers[b] || server_owner(servers[b]) != server_owner(s0)) return fail(SB200_ERR_ARG, "scan_batched: servers must share one database");
        q[b] = servers[b]->query.p; o[b] = servers[b]->scan_out.p;
    }
    if (launch_scan_spiral_batched(o, q, count, server_db(s0), s0->dim0, s0->local_num_per, ES(s0, stream)) != 0)
        return fail(SB200_ERR_ARG, "scan_batched: needs 2 * num_per to be a multiple of 128");
    CHECK_LAUNCH();
    return SB200_OK;
}
// Tensor-core batched first dimension: build the limb-tile copy of the resident database (owner only) for batches of up
// to `capacity` (<= 16) queries.  Costs a second database-sized allocation.
extern "C" int sb200_server_enable_tc(sb200_server *s, int capacity) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->db_owner) return fail(SB200_ERR_STATE, "enable_tc: call it on the server that owns the database");
    if (!s->have_db) return fail(SB200_ERR_STATE, "enable_tc: database not loaded");
    if (server_z_slices(s) != (size_t)kN) return fail(SB200_ERR_STATE, "enable_tc: needs an explicit database");
    if (capacity < 1 || capacity > 16) return fail(SB200_ERR_ARG, "enable_tc: capacity must be in [1, 16]");
    if (!tc_shape_ok(s->dim0, s->local_num_per)) return fail(SB200_ERR_ARG, "enable_tc: needs 2*dim0 and 2*num_per (per shard) to be multiples of 128");
    CU(cudaSetDevice(s->device));
    if (!s->db_tc.p) {                                  // (after sb200_server_tc_only the copy exists and only the staging is resized)
        CU(s->db_tc.alloc(s->db.n * sizeof(uint64_t)));
        launch_db_to_tc(s->db_tc.p, s->db.p, s->dim0, s->local_num_per, s->own_stream); CHECK_LAUNCH();
    }
    if (s->q_tc.p) { cudaFree(s->q_tc.p); s->q_tc.p = nullptr; }
    if (s->tc_t1.p) { cudaFree(s->tc_t1.p); s->tc_t1.p = nullptr; }
    CU(s->tc_t1.alloc(tc_scratch_bytes(s->local_num_per, capacity) / 4));
    CU(s->q_tc.alloc(tc_query_bytes(s->dim0, capacity)));
    CU(cudaMemsetAsync(s->q_tc.p, 0, s->q_tc.n, s->own_stream));
    CU(cudaStreamSynchronize(s->own_stream));
    s->tc_capacity = capacity;
    return SB200_OK;
}
// One tensor-core pass over the database for `count` (<= capacity) servers that share it: converts every server's
// reoriented query into its rows of the batched tiles, then one k_scan_tc launch fills every server's scan output.
extern "C" int sb200_server_scan_batched_tc(sb200_server *const *servers, int count, void *stream) {
    if (!servers || count < 1) return fail(SB200_ERR_ARG, "scan_batched_tc: no servers");
    sb200_server *s0 = servers[0];
    if (!s0) return fail(SB200_ERR_ARG, "scan_batched_tc: null server");
    sb200_server *owner = const_cast<sb200_server *>(s0->db_owner ? s0->db_owner : s0);
    if (!owner->tc_capacity) return fail(SB200_ERR_STATE, "scan_batched_tc: call sb200_server_enable_tc on the database owner first");
    if (count > owner->tc_capacity) return fail(SB200_ERR_ARG, "scan_batched_tc: %d queries exceed the enabled capacity %d", count, owner->tc_capacity);
    uint32_t *o[16]; const uint64_t *qs[16];
    cudaStream_t st = ES(s0, stream);
    for (int b = 0; b < count; b++) {
        if (!servers[b] || server_owner(servers[b]) != owner) return fail(SB200_ERR_ARG, "scan_batched_tc: servers must share one database");
        qs[b] = servers[b]->query.p; o[b] = servers[b]->scan_out.p;
    }
    launch_queries_to_tc(owner->q_tc.p, qs, count, 0, owner->tc_capacity, owner->dim0, st);
    if (launch_scan_tc(o, count, owner->tc_capacity, owner->q_tc.p, owner->db_tc.p, owner->dim0, owner->local_num_per, owner->tc_t1.p, st))
        return fail(SB200_ERR_CUDA, "scan_batched_tc: launch failed");
    CHECK_LAUNCH();
    return SB200_OK;
}
// Keep ONLY the limb-tile copy of the database (it holds the same residues, as u8 limbs in MMA tile order): the scan-layout
// copy is freed, so a server that batches on the tensor cores occupies one database's worth of HBM, not two, and a single query's
// first dimension (sb200_server_scan / _process / _answer*) becomes a one-query pass of k_scan_tc (0.41 ms instead of 0.36 ms
// at 2 GiB).  Servers sharing this database must issue their scans on one stream: they share the query-tile and result staging.
// Peak footprint is still two copies while sb200_server_enable_tc converts; loading the database again drops the tensor-core state.
extern "C" int sb200_server_tc_only(sb200_server *s) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->db_owner) return fail(SB200_ERR_STATE, "tc_only: call it on the server that owns the database");
    if (!s->tc_capacity || !s->db_tc.p) return fail(SB200_ERR_STATE, "tc_only: call sb200_server_enable_tc first");
    if (s->tc_only) return SB200_OK;
    CU(cudaSetDevice(s->device));
    CU(cudaDeviceSynchronize());                        // no scan of the copy being freed may be in flight
    cudaFree(s->db.p); s->db.p = nullptr;
    s->tc_only = true;
    return SB200_OK;
}
extern "C" int sb200_server_lift(sb200_server *s, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    launch_from_ntt(s->cts.p, s->scan_out.p, s->local_num_per * 6, ES(s, stream));
    CHECK_LAUNCH();
    return SB200_OK;
}
extern "C" int sb200_server_first_dim(sb200_server *s, void *stream) {
    TRY(sb200_server_scan(s, stream));
    return sb200_server_lift(s, stream);
}
extern "C" int sb200_server_scan_host(sb200_server *s, const uint64_t *reoriented_host, uint64_t *out_ref_ntt_host) {
    // multiplyQueryByDatabase against the resident database with a host-side reoriented query (interposed reference call)
    if (!s || !reoriented_host || !out_ref_ntt_host) return fail(SB200_ERR_ARG, "scan_host: null argument");
    if (!server_has_db(s)) return fail(SB200_ERR_STATE, "scan_host: database not loaded");
    CU(cudaMemcpy(s->query.p, reoriented_host, s->dim0 * 2 * 4 * kN * sizeof(uint64_t), cudaMemcpyHostToDevice));
    TRY(scan_one(s, 0));
    return down_ntt(out_ref_ntt_host, s->scan_out.p, s->local_num_per * 6);
}
extern "C" int sb200_server_copy_partial(sb200_server *s, uint64_t *dst_dev, void *stream) {
    if (!s || !dst_dev) return fail(SB200_ERR_ARG, "copy_partial: null argument");
    CU(cudaMemcpyAsync(dst_dev, s->cts.p, 6 * (size_t)kN * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ES(s, stream)));
    return SB200_OK;
}
// synthetic plaintext coefficients, uniform in [0, p_db): counter-based (splitmix64), 4 values per thread
__global__ void k_fill_random_u16(uint16_t *__restrict__ dst, size_t n4, uint32_t p_db, uint64_t seed) {
    pdl_prologue();
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    uint64_t x = seed + (i + 1) * 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull; x = (x ^ (x >> 27)) * 0x94d049bb133111ebull; x ^= x >> 31;
    ushort4 o;
    o.x = (uint16_t)((x & 0xffff) % p_db); o.y = (uint16_t)(((x >> 16) & 0xffff) % p_db);
    o.z = (uint16_t)(((x >> 32) & 0xffff) % p_db); o.w = (uint16_t)((x >> 48) % p_db);
    reinterpret_cast<ushort4 *>(dst)[i] = o;
}
extern "C" int sb200_server_load_db_random(sb200_server *s, uint64_t seed) {
    // synthetic database for benchmarks: uniform plaintext coefficients generated ON THE DEVICE chunk by chunk, then the normal
    // preprocessing (centre-lift, CRT, NTT, scan layout)
    if (!s) return fail(SB200_ERR_ARG, "null server");
    CU(cudaSetDevice(s->device));
    TRY(server_alloc_db(s));
    const size_t total_local = s->dim0 * s->local_num_per, chunk = 8192;
    DBuf<uint16_t> d;
    CU(d.alloc(std::min(chunk, total_local) * 4 * kN));
    const uint32_t nu2_local = (uint32_t)ceil_log2(s->local_num_per);
    for (size_t o = 0; o < total_local; o += chunk) {
        const size_t n = std::min(chunk, total_local - o), n4 = n * 4 * kN / 4;
        count_launch(); launch_pdl(k_fill_random_u16, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, 0, d.p, n4, (uint32_t)s->prm.p_db,
                                   seed * 0x100000001b3ull + o * 0x632be59bd9b4e019ull + 0x1234567ull);
        launch_db_build_spiral(s->db.p, d.p, (int)s->prm.nu1, (int)nu2_local, (uint32_t)s->prm.p_db, o, n, 0);
        CHECK_LAUNCH();
        CU(cudaDeviceSynchronize());
    }
    s->have_db = true;
    return SB200_OK;
}
static void fold_rounds(sb200_server *s, uint64_t *cts, size_t count, size_t first_dim, cudaStream_t st) {
    const size_t per = 3 * 3 * (size_t)s->prm.t_gsw * PLW;
    size_t np = count, d = first_dim;
    while (np >= 2) {
        np /= 2;
        launch_fold_round(cts, np, s->gsw.p + d * per, nullptr, (int)s->prm.t_gsw, s->fold_scratch.p, st);   // CMux form
        d++;
    }
}
extern "C" int sb200_server_fold_local(sb200_server *s, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    TRY(join_odd_chain(s, ES(s, stream)));                      // the GSW ciphertexts come from the side stream
    return run_stage(s->g_lift_fold, ES(s, stream), nullptr, nullptr, [&](cudaStream_t st) { fold_rounds(s, s->cts.p, s->local_num_per, 0, st); });
}
extern "C" uint64_t *sb200_server_partial_ct(sb200_server *s) { return s ? s->cts.p : nullptr; }
extern "C" uint64_t *sb200_server_first_dim_cts(sb200_server *s) { return s ? s->cts.p : nullptr; }
extern "C" int sb200_server_fold_tail(sb200_server *s, uint64_t *gathered, uint64_t *resp_dev, void *stream) {
    if (!s || !gathered || !resp_dev) return fail(SB200_ERR_ARG, "fold_tail: null argument");
    TRY(join_odd_chain(s, ES(s, stream)));
    return run_stage(s->g_tail, ES(s, stream), gathered, resp_dev, [&](cudaStream_t st) {
        fold_rounds(s, gathered, (size_t)s->world, s->prm.nu2 - s->log_world, st);
        // modulus switch (check_final, reference src/spiral.cpp:1441-1447): row 0 -> arb_qprime, rows 1.. -> 4*p_db
        launch_rescale2(resp_dev, gathered, 2 * (size_t)kN, 4 * (size_t)kN, kQ, sb200_arb_qprime(s->prm.qp_bits), 4 * s->prm.p_db, st);
    });
}
// ---- peer-memory exchange: setup -----------------------------------------------------------------
extern "C" size_t sb200_server_xchg_handle_bytes(void) { return 3 * sizeof(cudaIpcMemHandle_t); }
// a rank's handle = its exchange buffer + its query buffer + its GSW buffer (the sharded expansion stores converted query slices
// and GSW columns into the peers')
extern "C" int sb200_server_xchg_export(sb200_server *s, void *handle_out) {
    if (!s || !handle_out || s->world < 2) return fail(SB200_ERR_ARG, "xchg_export: needs a sharded server");
    cudaIpcMemHandle_t h[3];
    CU(cudaIpcGetMemHandle(&h[0], s->xchg.p));
    CU(cudaIpcGetMemHandle(&h[1], s->query.p));
    CU(cudaIpcGetMemHandle(&h[2], s->gsw.p));
    memcpy(handle_out, h, sizeof h);
    return SB200_OK;
}
// CUDA loads kernels lazily, at their first launch, and loading may need a context-wide synchronisation - which never comes while
// a kernel of this context is spinning on a flag that the not-yet-loaded kernel (or anything queued behind the load) would set.
// Everything a sharded query launches is therefore loaded when the peers are connected, before any flag is waited for.
template <typename K> static void preload_kernel(K *k) { cudaFuncAttributes a; if (cudaFuncGetAttributes(&a, k) != cudaSuccess) cudaGetLastError(); }
static void preload_exchange_kernels() {
    preload_kernel(k_xchg_push); preload_kernel(k_xchg_wait); preload_kernel(k_xchg_ack); preload_kernel(k_flag_wait); preload_kernel(k_query_wait);
    preload_kernel(k_reorient_dim1_allgather); preload_kernel(k_scal_to_mat_accum_tiled); preload_kernel(k_regev_to_gsw_accum_sharded);
    preload_kernel(k_rescale2); preload_kernel(k_fold_decomp_ntt); preload_kernel(k_fold_mac); preload_kernel(k_fold_mac_wide); preload_kernel(k_fold_lift);
    preload_kernel(k_from_ntt); preload_kernel(k_from_ntt_indexed); preload_kernel(k_gadget_ntt); preload_kernel(k_expand_prep); preload_kernel(k_expand_digits);
    preload_kernel(k_expand_accum); preload_kernel(k_expand_accum_wide); preload_kernel(k_ntt_u64_to_dev); preload_kernel(k_query_from_wire);
    preload_kernel(k_scan_spiral_jsplit); preload_kernel(k_scan_spiral<2, 128, 4, true>); preload_kernel(k_scan_spiral<2, 128, 4, false>);
    preload_kernel(k_scan_spiral<2, 64, 4, false>); preload_kernel(k_scan_spiral<1, 128, 4, false>);
    scan_tma_prepare(); preload_kernel(k_scan_pack_wide<1, 256, 8>); preload_kernel(k_scan_pack_narrow<8, 8, 2>); preload_kernel(k_scan_pack_narrow<8, 8, 1>);
    preload_kernel(k_scan_pack); preload_kernel(k_pack_accum); preload_kernel(k_split_rows); preload_kernel(k_simple_gsw_accum); preload_kernel(k_reorient_dim1);
    cudaGetLastError();
}
static int xchg_finish_connect(sb200_server *s, const std::vector<void *> &bufs, const std::vector<void *> &queries, const std::vector<void *> &gsws) {
    preload_exchange_kernels();
    s->peer_xb = bufs; s->peer_query = queries; s->peer_gsw = gsws;
    s->xchg_target = bufs[0];
    if (s->rank == 0) {
        std::vector<unsigned int *> acks(s->world);
        for (int r = 0; r < s->world; r++) acks[r] = reinterpret_cast<unsigned int *>(reinterpret_cast<uint8_t *>(bufs[r]) + xchg_ack_offset());
        CU(s->xchg_acks.up(acks.data(), s->world));
    }
    s->xchg_connected = true;
    return SB200_OK;
}
// all_handles: world handles in rank order (each rank's sb200_server_xchg_export), one process per GPU
extern "C" int sb200_server_xchg_connect(sb200_server *s, const void *all_handles) {
    if (!s || !all_handles || s->world < 2) return fail(SB200_ERR_ARG, "xchg_connect: needs a sharded server");
    std::vector<void *> bufs(s->world, nullptr), queries(s->world, nullptr), gsws(s->world, nullptr);
    for (int r = 0; r < s->world; r++) {
        if (r == s->rank) { bufs[r] = s->xchg.p; queries[r] = s->query.p; gsws[r] = s->gsw.p; continue; }
        cudaIpcMemHandle_t h[3];
        memcpy(h, reinterpret_cast<const uint8_t *>(all_handles) + (size_t)r * sizeof h, sizeof h);
        void *pb = nullptr, *pq = nullptr, *pg = nullptr;
        CU(cudaIpcOpenMemHandle(&pb, h[0], cudaIpcMemLazyEnablePeerAccess)); s->ipc_opened.push_back(pb);
        CU(cudaIpcOpenMemHandle(&pq, h[1], cudaIpcMemLazyEnablePeerAccess)); s->ipc_opened.push_back(pq);
        CU(cudaIpcOpenMemHandle(&pg, h[2], cudaIpcMemLazyEnablePeerAccess)); s->ipc_opened.push_back(pg);
        bufs[r] = pb; queries[r] = pq; gsws[r] = pg;
    }
    return xchg_finish_connect(s, bufs, queries, gsws);
}
// same-process variant (several shards driven from one process, e.g. tests on one device): direct pointers
extern "C" int sb200_server_xchg_connect_local(sb200_server *s, sb200_server *const *all_servers) {
    if (!s || !all_servers || s->world < 2) return fail(SB200_ERR_ARG, "xchg_connect_local: needs a sharded server");
    std::vector<void *> bufs(s->world, nullptr), queries(s->world, nullptr), gsws(s->world, nullptr);
    for (int r = 0; r < s->world; r++) {
        if (!all_servers[r] || all_servers[r]->world != s->world || all_servers[r]->rank != r) return fail(SB200_ERR_ARG, "xchg_connect_local: server %d mismatched", r);
        bufs[r] = all_servers[r]->xchg.p; queries[r] = all_servers[r]->query.p; gsws[r] = all_servers[r]->gsw.p;
        if (all_servers[r]->device != s->device) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, s->device, all_servers[r]->device));
            if (!can) return fail(SB200_ERR_CUDA, "xchg_connect_local: no peer access %d -> %d", s->device, all_servers[r]->device);
            cudaError_t pe = cudaDeviceEnablePeerAccess(all_servers[r]->device, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CU(pe);
            cudaGetLastError();
        }
    }
    return xchg_finish_connect(s, bufs, queries, gsws);
}
// every rank: push the surviving ciphertext into rank 0's HBM; rank 0 additionally waits for all shards, runs the
// tail folds and the modulus switch into resp_dev (ignored on other ranks)
extern "C" int sb200_server_exchange_and_tail(sb200_server *s, uint64_t *resp_dev, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->world < 2) return sb200_server_fold_tail(s, s->cts.p, resp_dev, stream);
    if (!s->xchg_connected) return fail(SB200_ERR_STATE, "exchange_and_tail: peers not connected (sb200_server_xchg_connect)");
    if (s->rank == 0 && !resp_dev) return fail(SB200_ERR_ARG, "exchange_and_tail: rank 0 needs a response buffer");
    TRY(join_odd_chain(s, ES(s, stream)));
    const bool late_ack = s->query_sharded;       // sharded expansion: peers write into this rank's query / GSW buffers, so rank 0
                                                  // acknowledges only once its tail folds no longer read them
    return run_stage(s->g_xchg, ES(s, stream), resp_dev, late_ack ? (const void *)s->xchg.p : nullptr, [&](cudaStream_t st) {
        launch_xchg_push(s->xchg_target, s->xchg.p, s->cts.p, s->xchg_state.p, s->rank, s->world, s->xchg_state.p + 1, st);
        if (s->rank == 0) {
            launch_xchg_wait(s->xchg.p, s->xchg_acks.p, s->gathered.p, s->xchg_state.p, s->world, s->xchg_state.p + 1, st, late_ack ? 0 : 1);
            fold_rounds(s, s->gathered.p, (size_t)s->world, s->prm.nu2 - s->log_world, st);
            launch_rescale2(resp_dev, s->gathered.p, 2 * (size_t)kN, 4 * (size_t)kN, kQ, sb200_arb_qprime(s->prm.qp_bits), 4 * s->prm.p_db, st);
            if (late_ack) launch_xchg_ack(s->xchg_acks.p, s->xchg_state.p, s->world, s->xchg_state.p + 1, st);
        }
    });
}
// 0 = ok, 1 = push timed out waiting for a free slot, 2 = rank 0 timed out waiting for a shard (synchronises the stream)
extern "C" int sb200_server_xchg_error(sb200_server *s, void *stream) {
    if (!s || s->world < 2) return 0;
    unsigned int st[2] = {0, 0};
    if (cudaStreamSynchronize(ES(s, stream)) != cudaSuccess) return -1;
    if (cudaMemcpy(st, s->xchg_state.p, sizeof st, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int)st[1];
}

extern "C" int sb200_server_download(sb200_server *s, uint64_t *dst, const uint64_t *src, size_t words, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    CU(cudaMemcpyAsync(dst, src, words * sizeof(uint64_t), cudaMemcpyDeviceToHost, ES(s, stream)));
    CU(cudaStreamSynchronize(ES(s, stream)));
    if (s->world > 1 && s->xchg_connected) {       // a response produced after a timed-out exchange is garbage: never hand it out as SB200_OK
        unsigned int st[2] = {0, 0};
        CU(cudaMemcpy(st, s->xchg_state.p, sizeof st, cudaMemcpyDeviceToHost));
        if (st[1]) return fail(SB200_ERR_STATE, "peer exchange timed out (error %u: %s); the shards are out of step - call sb200_server_xchg_reset on every rank",
                               st[1], st[1] == 1 ? "no free slot on rank 0" : "a shard's ciphertext never arrived");
    }
    return SB200_OK;
}
// after a timed-out exchange: EVERY rank calls this with no query in flight; epochs, flags, acks and the error word start over
extern "C" int sb200_server_xchg_reset(sb200_server *s) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->world < 2) return SB200_OK;
    CU(cudaSetDevice(s->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemset(s->xchg.p, 0, xchg_header_bytes()));
    CU(cudaMemset(s->xchg_state.p, 0, 2 * sizeof(unsigned int)));
    CU(cudaDeviceSynchronize());
    return SB200_OK;
}
// ---------------------------------------------------------------------------------------------
// resident drop-in: the leaf functions of process_query_fast (src/spiral.cpp:1584-1629) driven by the
// UNMODIFIED harness, with every intermediate staying in HBM between the leaves.  The harness's host buffers are only KEYS here:
// a leaf whose input key matches what the previous leaf left on the device skips the upload, and nothing is downloaded until
// un-interposed host code needs it (the final ciphertext, 96 KiB).  A mismatch returns SB200_ERR_STATE and the caller falls back
// to the stateless sb200_<refname> call, so correctness never depends on the harness's call order.
// ---------------------------------------------------------------------------------------------
extern "C" int sb200_resident_reorientCiphertexts(sb200_server *s, const void *out_key, const uint64_t *inp_ref_ntt_host, size_t dim0) {
    if (!s || !out_key || !inp_ref_ntt_host) return fail(SB200_ERR_ARG, "resident reorientCiphertexts: null argument");
    if (dim0 != s->dim0 || s->world != 1) return fail(SB200_ERR_STATE, "resident reorientCiphertexts: shape differs from the resident server");
    CU(cudaSetDevice(s->device));
    if (!s->res_cts_in.p) CU(s->res_cts_in.alloc(s->dim0 * kN1 * 2 * PLW));
    TRY(server_up_ntt(s, s->res_cts_in, inp_ref_ntt_host, s->dim0 * kN1 * 2));
    launch_reorient_query(s->query.p, s->res_cts_in.p, s->dim0, 0); CHECK_LAUNCH();
    CU(cudaDeviceSynchronize());
    s->res_query_key = out_key; s->res_scan_key = s->res_cts_key = nullptr;
    return SB200_OK;
}
extern "C" int sb200_resident_multiplyQueryByDatabase(sb200_server *s, const void *out_key, const void *reoriented_key) {
    if (!s || !out_key) return fail(SB200_ERR_ARG, "resident multiplyQueryByDatabase: null argument");
    if (!s->res_query_key || reoriented_key != s->res_query_key) return fail(SB200_ERR_STATE, "resident multiplyQueryByDatabase: the query is not resident");
    if (!server_has_db(s)) return fail(SB200_ERR_STATE, "resident multiplyQueryByDatabase: database not loaded");
    TRY(scan_one(s, 0));
    CU(cudaDeviceSynchronize());                       // the harness's timer brackets the call
    s->res_scan_key = out_key; s->res_cts_key = nullptr;
    return SB200_OK;
}
extern "C" int sb200_resident_nttInvAndCrtLiftCiphertexts(sb200_server *s, const void *cts_key, const void *scratch_key) {
    if (!s || !cts_key) return fail(SB200_ERR_ARG, "resident nttInvAndCrtLiftCiphertexts: null argument");
    if (!s->res_scan_key || scratch_key != s->res_scan_key) return fail(SB200_ERR_STATE, "resident nttInvAndCrtLiftCiphertexts: the scan output is not resident");
    launch_from_ntt(s->cts.p, s->scan_out.p, s->local_num_per * 6, 0); CHECK_LAUNCH();
    CU(cudaDeviceSynchronize());
    s->res_cts_key = cts_key;
    return SB200_OK;
}
// reorient_Q(out, inp) (src/spiral.cpp:388-400): inp = one GSW ciphertext (n1 x m2, ref-NTT); its dev-NTT copy is registered under `out`
extern "C" int sb200_resident_reorient_Q(sb200_server *s, const void *out_key, const uint64_t *inp_ref_ntt_host) {
    if (!s || !out_key || !inp_ref_ntt_host) return fail(SB200_ERR_ARG, "resident reorient_Q: null argument");
    CU(cudaSetDevice(s->device));
    const size_t polys = (size_t)kN1 * kN1 * s->prm.t_gsw;
    DBuf<uint32_t> *buf = nullptr;
    for (auto &e : s->res_gsw) if (e.first == out_key) buf = e.second;
    if (!buf) {
        if (s->res_gsw.size() >= 4 * (size_t)s->prm.nu2 + 8) {        // keys of earlier queries (the harness mallocs fresh buffers per query)
            for (auto &e : s->res_gsw) delete e.second;
            s->res_gsw.clear();
        }
        buf = new DBuf<uint32_t>();
        cudaError_t e = buf->alloc(polys * PLW);
        if (e != cudaSuccess) { delete buf; return fail(SB200_ERR_CUDA, "resident reorient_Q: %s", cudaGetErrorString(e)); }
        s->res_gsw.emplace_back(out_key, buf);
    }
    return server_up_ntt(s, *buf, inp_ref_ntt_host, polys);
}
// foldOneFurtherDimension on the resident ciphertexts; q_key / q_neg_key: the reorient_Q outputs of THIS dimension.  When one
// ciphertext is left (num_per == 1) it is downloaded into cts_host (the harness's check_final / modswitch read it there).
extern "C" int sb200_resident_foldOneFurtherDimension(sb200_server *s, size_t num_per, const void *q_key, const void *q_neg_key, uint64_t *cts_host) {
    if (!s || !cts_host) return fail(SB200_ERR_ARG, "resident foldOneFurtherDimension: null argument");
    if (!s->res_cts_key || (const void *)cts_host != s->res_cts_key) return fail(SB200_ERR_STATE, "resident foldOneFurtherDimension: the ciphertexts are not resident");
    const uint32_t *q = nullptr, *qn = nullptr;
    for (auto &e : s->res_gsw) { if (e.first == q_key) q = e.second->p; if (e.first == q_neg_key) qn = e.second->p; }
    if (!q || !qn) return fail(SB200_ERR_STATE, "resident foldOneFurtherDimension: the GSW ciphertexts are not resident");
    if (2 * num_per > s->local_num_per) return fail(SB200_ERR_ARG, "resident foldOneFurtherDimension: num_per too large");
    launch_fold_round(s->cts.p, num_per, q, qn, (int)s->prm.t_gsw, s->fold_scratch.p, 0); CHECK_LAUNCH();   // the reference's two-product form
    if (num_per == 1) CU(cudaMemcpy(cts_host, s->cts.p, 6 * (size_t)kN * 8, cudaMemcpyDeviceToHost));
    else CU(cudaDeviceSynchronize());
    return SB200_OK;
}

extern "C" int sb200_server_answer(sb200_server *s, const uint64_t *query_cv_host, uint64_t *total_resp_host, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->world != 1) return fail(SB200_ERR_STATE, "server_answer: single-shard call on a sharded server (use the staged API)");
    TRY(sb200_server_upload_query(s, query_cv_host, stream));
    TRY(sb200_server_process(s, s->resp.p, stream, nullptr));
    return sb200_server_download(s, total_resp_host, s->resp.p, 6 * kN, stream);
}
// One resident query in ONE call (the query is already in q_stage / q_wire): all server stages into total_resp_dev.  A serving loop
// written in a scripting language issues one foreign call per query instead of six, so the host stays ahead of the ~1 ms of GPU work.
// marks (optional): four cudaEvent_t recorded before the expansion, before the scan, after the scan and at the end.
// Capture and instantiate every CUDA graph sb200_server_process will replay (expansion chains, folds, exchange + tail) without
// running anything: the first query then costs what the others do.  Needs the public parameters (and connected peers on a sharded
// server, since the sharded expansion is a different graph); total_resp_dev as it will be passed to sb200_server_process.
extern "C" int sb200_server_prepare(sb200_server *s, uint64_t *total_resp_dev, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (!s->have_params) return fail(SB200_ERR_STATE, "server_prepare: public parameters not set");
    if (s->world > 1 && !s->xchg_connected) return fail(SB200_ERR_STATE, "server_prepare: sharded server without connected peers");
    CU(cudaSetDevice(s->device));
    uint64_t *resp = total_resp_dev ? total_resp_dev : s->resp.p;
    tl_prepare_only = true;
    int rc = expand_and_convert_impl(s, stream, true);
    if (!rc) rc = sb200_server_fold_local(s, stream);
    if (!rc) rc = sb200_server_exchange_and_tail(s, resp, stream);
    tl_prepare_only = false;
    s->join_pending = s->gsw_wait_pending = s->query_sharded = false;
    CU(cudaStreamSynchronize(s->aux_stream));
    CU(cudaStreamSynchronize(ES(s, stream)));
    return rc;
}
static int process_stages(sb200_server *s, uint64_t *resp, void *stream, void *const *marks) {
    cudaStream_t st = ES(s, stream);
    if (marks) CU(cudaEventRecord((cudaEvent_t)marks[0], st));
    TRY(expand_and_convert_impl(s, stream, true));
    if (marks) CU(cudaEventRecord((cudaEvent_t)marks[1], st));
    TRY(sb200_server_scan(s, stream));
    if (marks) CU(cudaEventRecord((cudaEvent_t)marks[2], st));
    TRY(sb200_server_lift(s, stream));
    TRY(sb200_server_fold_local(s, stream));
    // one shard: tail = modulus switch; several: push this shard's ciphertext to rank 0 over NVLink, rank 0 folds the rest
    TRY(sb200_server_exchange_and_tail(s, resp, stream));
    if (marks) CU(cudaEventRecord((cudaEvent_t)marks[3], st));
    return SB200_OK;
}
// Measured and not adopted (profiles/r02_expansion_chains.md): the whole query captured as ONE graph (both chains, scan, folds).
// Device-resident it runs as fast as the per-stage graphs (0.933 vs 0.935 ms - a graph-to-graph boundary costs nothing that the
// programmatic launches do not already hide), and with host buffers it is slower (1.00 vs 0.96 ms): the first kernel cannot start
// before all ~200 nodes are submitted, while the first small stage graph starts at once and the rest is submitted under it.
extern "C" int sb200_server_process(sb200_server *s, uint64_t *total_resp_dev, void *stream, void *const *marks) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->world > 1 && !s->xchg_connected) return fail(SB200_ERR_STATE, "server_process: sharded server without connected peers (sb200_server_xchg_connect)");
    return process_stages(s, total_resp_dev ? total_resp_dev : s->resp.p, stream, marks);
}
// the same query with the response in its wire format: 20 KiB instead of 96 KiB cross PCIe at cfg1
extern "C" int sb200_server_answer_packed(sb200_server *s, const uint64_t *query_cv_host, uint64_t *packed_resp_host, void *stream) {
    if (!s || !packed_resp_host) return fail(SB200_ERR_ARG, "server_answer_packed: null argument");
    if (s->world != 1) return fail(SB200_ERR_STATE, "server_answer_packed: single-shard call on a sharded server (use the staged API)");
    TRY(sb200_server_upload_query(s, query_cv_host, stream));
    TRY(sb200_server_process(s, s->resp.p, stream, nullptr));
    TRY(sb200_dev_pack_response(s->final_ct.p, s->resp.p, 2 * (size_t)kN, 4 * (size_t)kN, s->prm.qp_bits, s->prm.p_db, ES(s, stream)));
    return sb200_server_download(s, packed_resp_host, s->final_ct.p, sb200_packed_response_words(2 * (size_t)kN, 4 * (size_t)kN, s->prm.qp_bits, s->prm.p_db), stream);
}
extern "C" size_t sb200_server_packed_response_bytes(const sb200_server *s) {
    return s ? 8 * sb200_packed_response_words(2 * (size_t)kN, 4 * (size_t)kN, s->prm.qp_bits, s->prm.p_db) : 0;
}
extern "C" size_t sb200_server_query_bytes(const sb200_server *) { return 2 * PLW * sizeof(uint64_t); }
extern "C" size_t sb200_server_response_bytes(const sb200_server *) { return 6 * (size_t)kN * sizeof(uint64_t); }
