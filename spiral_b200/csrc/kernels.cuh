// kernels.cuh - launch wrappers for every CUDA kernel of the Spiral server path.
// Device data formats:
//   NTT form ("dev-NTT") : uint32_t [poly][2][2048]  plane 0 = residues mod p, plane 1 = mod b
//                          (the reference's [n][z] layout narrowed to the 28-bit residues)
//   raw form             : uint64_t [poly][2048] in [0, Q]
//   packed ("PB64")      : uint64_t, low 32 bits = residue mod p, high 32 bits = residue mod b
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace sb200 {

void count_launch(int n = 1);        // our own kernel launches since load (sb200_launch_count)
uint64_t launch_count();
bool pdl_enabled();                    // SB200_NO_PDL=1 disables programmatic dependent launch
// Launch priority of the calling thread's next launches (captured into graph kernel nodes as well): the expansion runs two
// independent chains on two streams - the one the scan waits for is launched at the device's greatest priority, so that when
// both have CTAs pending the critical chain's are dispatched first.
constexpr int kNoPriority = 1 << 30;
int &launch_priority();
struct LaunchPriority {
    int saved;
    explicit LaunchPriority(bool high, int level = 1);   // active when SB200_PRIO == level
    ~LaunchPriority() { launch_priority() = saved; }
};
void note_kernel(const char *name);    // distinct kernel names launched since the last reset (sb200_kernel_log)
template <typename... KArgs, typename... Args>
inline void launch_pdl_impl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (launch_priority() != kNoPriority) {            // LaunchPriority scope: kernels of a latency-critical chain go first
        attr[1].id = cudaLaunchAttributePriority;
        attr[1].val.priority = launch_priority();
        cfg.numAttrs = 2;
    }
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
// every launch of the library goes through this macro: it records the kernel's name (a template instance is written in
// parentheses at the call site so its commas survive the preprocessor)
#define launch_pdl(kernel, ...) (::sb200::note_kernel(#kernel), ::sb200::launch_pdl_impl(kernel, __VA_ARGS__))
size_t kernel_log(char *buf, size_t cap);
void kernel_log_reset();
int trace_enable(unsigned int capacity);                 // timeline trace of dependency-resolved times (profiling; 0 = off)
size_t trace_read(unsigned long long *out, size_t max_records, int reset);
int init_tables();   // builds twiddle tables on the current device (idempotent, per device)

// ---- format conversion at the boundary (reference u64-per-residue NTT layout <-> dev-NTT)
void launch_ntt_u64_to_dev(uint32_t *out, const uint64_t *in, size_t npolys, cudaStream_t s);
void launch_ntt_dev_to_u64(uint64_t *out, const uint32_t *in, size_t npolys, cudaStream_t s);

// ---- MatPoly primitives (reference src/poly.cpp) on batches of polynomials
void launch_to_ntt(uint32_t *out, const uint64_t *raw, size_t npolys, cudaStream_t s);          // to_ntt / to_ntt_no_reduce
void launch_from_ntt(uint64_t *raw, const uint32_t *in, size_t npolys, cudaStream_t s);        // from_ntt (INTT + CRT lift)
void launch_matmul(uint32_t *out, const uint32_t *a, const uint32_t *b, int rs, int ms, int cs, cudaStream_t s);
void launch_add(uint32_t *out, const uint32_t *a, const uint32_t *b, size_t npolys, cudaStream_t s);
void launch_automorph(uint64_t *out, const uint64_t *in, size_t npolys, uint32_t t, cudaStream_t s);
// raw (rdim x cols) -> NTT'd digits (mx x cols), row j + k*rdim  (gadget_invert + to_ntt)
void launch_gadget_ntt(uint32_t *out, const uint64_t *raw, int mx, int rdim, int cols, cudaStream_t s);
void launch_gadget_raw(uint64_t *out, const uint64_t *raw, int mx, int rdim, int cols, cudaStream_t s);
void launch_rescale(uint64_t *out, const uint64_t *in, size_t ncoeffs, uint64_t inp_mod, uint64_t out_mod, cudaStream_t s);
void launch_rescale2(uint64_t *out, const uint64_t *in, size_t n0, size_t n1, uint64_t inp_mod, uint64_t mod0, uint64_t mod1, cudaStream_t s);
// write_arbitrary_bits over a whole buffer: n values of `bits` bits -> ceil(n*bits/64) words (reference src/core.cpp:32-52)
void launch_bitpack(uint64_t *out, const uint64_t *in, size_t n, uint32_t bits, cudaStream_t s);
// modswitch (reference src/spiral.cpp:40-78): round(v * qprime / Q) in the reference's x87 arithmetic, bit-packed
void launch_modswitch(uint64_t *out, const uint64_t *in_raw, size_t n, uint32_t bits, uint64_t qprime, cudaStream_t s);

// ---- wire / on-disk formats (wire_kernels.cu)
// wire query (device copy, header included; kind 1 = seeded, 2 = full) -> cv[0] in dev-NTT form
void launch_query_from_wire(uint32_t *cv, const uint8_t *wire, uint32_t kind, cudaStream_t s);
// staged records (n_items whole items of `polys` polynomials, `bits` per coefficient, 4 readable padding bytes) -> u16 coefficients
void launch_records_to_pts(uint16_t *out, const uint8_t *rec, size_t n_items, int polys, uint32_t bits, size_t out_item_stride,
                           size_t out_poly_stride, cudaStream_t s);
void launch_sum64(unsigned long long *acc, const uint64_t *words, size_t n, cudaStream_t s);   // *acc += sum of words (mod 2^64)

// ---- database preprocessing
// plaintext items (values < p_db, one u16 per coefficient, item-major [item][n0*n2][2048]) ->
// scan layout DB'[z][j][ic][m] PB64 with ic = ii*n2 + c, item = j*num_per + ii.
void launch_db_build_spiral(uint64_t *db, const uint16_t *pts, int nu1, int nu2, uint32_t p_db,
                            size_t item_begin, size_t item_count, cudaStream_t s);
// reference layout B[z][ii][c][j][m] (src/spiral.cpp:1139-1153) -> scan layout, rows_z z-slices at a time
void launch_db_from_reference(uint64_t *db, const uint64_t *B_ref, size_t dim0, size_t ic, size_t z_begin,
                              size_t z_count, cudaStream_t s);

void launch_db_to_reference(uint64_t *B_ref, const uint64_t *db, size_t dim0, size_t ic, cudaStream_t s);
void launch_unreorient_q(uint32_t *out, const uint64_t *q_reor, int rm_count, cudaStream_t s);

// ---- first dimension
// query: [z][j][m][4] PB64 (reference reorientCiphertexts layout);  db: scan layout;
// out: dev-NTT [i][r][c] (num_per x 3 x 2 polys)
void launch_reorient_query(uint64_t *out, const uint32_t *cts, size_t dim0, cudaStream_t s);
// z_slices: slices the database buffer holds (2048; an implicit database holds a power of two fewer and slice z mod z_slices is read)
void launch_scan_spiral(uint32_t *out, const uint64_t *query, const uint64_t *db, size_t dim0, size_t num_per, cudaStream_t s, size_t z_slices = 0);
// the same scan with the database streamed through shared memory by the TMA engine (scan_tma.cu); false: shape outside its domain
bool launch_scan_spiral_tma(uint32_t *out, const uint64_t *query, const uint64_t *db, size_t dim0, size_t num_per, cudaStream_t s, size_t z_slices);
void scan_tma_prepare();

int launch_scan_spiral_batched(uint32_t *const *out, const uint64_t *const *query, int count, const uint64_t *db, size_t dim0,
                               size_t num_per, cudaStream_t s);   // count in {2,4}: queries sharing one database pass

// ---- batched first dimension on tcgen05 tensor cores (tc_scan.cu): up to 16 queries per database pass
// generic geometry: K = bytes of k per database row (Spiral 2*dim0, Pack dim0), IC = database columns (Spiral 2*num_per, Pack num_per),
// RQ = ciphertext rows per query (3 / 2), planes (Pack: out_n^2), kstride = PB64 words between consecutive k in the reoriented query
struct TcGeom { size_t K, IC; int RQ; size_t planes, out_plane_polys, kstride; };
TcGeom tc_geom_spiral(size_t dim0, size_t num_per);
TcGeom tc_geom_pack(size_t dim0, size_t num_per, size_t planes, size_t out_plane_polys);
int tc_geom_ok(const TcGeom &g);                                      // K and IC multiples of 128
size_t tc_db_bytes_g(const TcGeom &g);
size_t tc_query_bytes_g(const TcGeom &g, int capacity);
size_t tc_scratch_bytes_g(const TcGeom &g, int count);
void launch_db_to_tc_g(uint8_t *db_tc, const uint64_t *db, const TcGeom &g, size_t src_plane_words, cudaStream_t s);
void launch_queries_to_tc_g(uint8_t *q_tc, const uint64_t *const *queries, int count, int first_slot, int capacity, const TcGeom &g, cudaStream_t s);
int launch_scan_tc_g(uint32_t *const *out, int count, int capacity, const uint8_t *q_tc, const uint8_t *db_tc, const TcGeom &g,
                     uint32_t *scratch, cudaStream_t s);
// Spiral-shaped wrappers
int tc_shape_ok(size_t dim0, size_t num_per);                       // needs 2*dim0 % 128 == 0 and 2*num_per % 128 == 0
size_t tc_query_bytes(size_t dim0, int capacity);                   // Q_tc bytes for a batch of up to `capacity` queries
void launch_db_to_tc(uint8_t *db_tc, const uint64_t *db, size_t dim0, size_t num_per, cudaStream_t s);   // scan layout -> DB_tc (same size)
void launch_query_to_tc(uint8_t *q_tc, const uint64_t *query, int q, int capacity, size_t dim0, cudaStream_t s);
void launch_queries_to_tc(uint8_t *q_tc, const uint64_t *const *queries, int count, int first_slot, int capacity, size_t dim0, cudaStream_t s);
size_t tc_scratch_bytes(size_t num_per, int count);                 // tile-order results of one pass, before the transpose
int launch_scan_tc(uint32_t *const *out, int count, int capacity, const uint8_t *q_tc, const uint8_t *db_tc, size_t dim0, size_t num_per,
                   uint32_t *scratch, cudaStream_t s);

// ---- folding (Spiral): cts raw [2*num_per][3][2][2048] -> first num_per folded in place.
// q_dev / qneg_dev: dev-NTT (3 x 3*t_gsw) GSW ciphertext of THIS round.
void launch_fold_round(uint64_t *cts, size_t num_per, const uint32_t *q_dev, const uint32_t *qneg_dev,
                       int t_gsw, uint32_t *scratch, cudaStream_t s);
size_t fold_scratch_words(size_t num_per_half, int t_gsw);
size_t fold_scratch_words_generic(size_t cts_in, int R, int Cc, int t);
// R x Cc ciphertexts, `planes` independent planes (plane p's ciphertexts start at p*plane_stride), signed (Spiral) or unsigned (Pack) digits
void launch_fold_round_generic(uint64_t *cts, int R, int Cc, int t, int is_signed, size_t np_after, size_t planes, size_t plane_stride,
                               const uint32_t *q_dev, const uint32_t *qneg_dev, uint32_t *scratch, cudaStream_t s);
void launch_fold_decomp_only(uint32_t *scratch, const uint64_t *cts, size_t count, int t_gsw, cudaStream_t s);

// ---- expansion (reference expandImproved / coefficientExpansion)
struct ExpandPlan { int g, t_left, t_right, stopround, max_bits_right; };
size_t expand_active_total(const ExpandPlan &p);
int expand_build_lists(const ExpandPlan &p, int *list, int *offs, int *cnt);   // host arrays; returns max count
size_t expand_ginv_polys(const ExpandPlan &p, const int *cnt);                 // polynomials of digit scratch (ginv) launch_expand needs
// cv: dev-NTT [2^g][2]; W_left: [g][2][t_left]; W_right: [g or stopround+1][2][t_right]; neg1: [g] polys
// c0_raw: maxcnt*2048 u64; c1_ntt: maxcnt polys; ginv: maxcnt*max(t) polys; list_dev: device copy of list
void build_automorph_perms(uint16_t *perm_host, int g);      // host: g x 2048 slot permutations (one per round)
void launch_expand(uint32_t *cv, const ExpandPlan &p, const uint32_t *W_left, const uint32_t *W_right,
                   const uint32_t *neg1, const uint16_t *perms, uint64_t *c0_raw, uint32_t *c1_ntt, uint32_t *ginv,
                   const int *list_dev, const int *offs, const int *cnt, cudaStream_t s, int r_begin = 0, int r_end = -1, int parity = -1,
                   int store_self = 0, int slot_limit = 0, int digit_chunk = 0);   // slot_limit > 0: the digit kernels never hold more than that
                                                                                   // many CTAs; digit_chunk > 0: big digit rounds in launches of ~that many
int expand_split_lists(const ExpandPlan &p, const int *list, const int *offs, const int *cnt, int *list_e, int *offs_e, int *cnt_e,
                       int *list_o, int *offs_o, int *cnt_o);                    // even / odd chains of the expansion tree
void build_neg1(uint32_t *neg1_dev, int count, cudaStream_t s);   // neg1[r] = NTT(-x^(N-2^r)), r < count, then their Shoup companions (2 * count polys)

// ---- conversion
void launch_from_ntt_indexed(uint64_t *raw, const uint32_t *in, const int *poly_idx, size_t count, cudaStream_t s);
// ct_idx[j] = ciphertext index inside cv, poly_idx[j] = 2*ct_idx[j] (row-0 polynomial)
void launch_scal_to_mat_reoriented(uint64_t *query_out, const uint32_t *cv, const int *ct_idx, const int *poly_idx, size_t dim0,
                                   const uint32_t *W, int t_conv, uint64_t *scratch_raw, uint32_t *scratch_ntt, cudaStream_t s);
void launch_scal_to_mat_ntt(uint32_t *out, const uint32_t *cv, const int *ct_idx, const int *poly_idx, size_t count,
                            const uint32_t *W, int t_conv, uint64_t *scratch_raw, uint32_t *scratch_ntt, cudaStream_t s);
// poly_idx: 2*nbits entries (row-0 polys of the nu2*t_gsw bit ciphertexts, then their row-1 polys)
void launch_regev_to_gsw(uint32_t *gsw_out, uint32_t *gsw_neg_out, const uint32_t *cv, const int *ct_idx, const int *poly_idx,
                         int nu2, int t_gsw, const uint32_t *W, const uint32_t *V, int t_conv,
                         uint64_t *scratch_raw, uint32_t *scratch_ntt, cudaStream_t s);
void launch_gsw_negate(uint32_t *neg, const uint32_t *gsw, int count, int ell, int rows, cudaStream_t s);

}  // namespace sb200
