// query_kernels.cu - database-independent part of query answering: coefficient expansion of the
// packed query, ScalToMat / RegevToGSW conversion and GSW negation, all on the device so that
// only the single query ciphertext crosses PCIe per query.
//
// Reference functions replaced:
//   expandImproved / coefficientExpansion   src/spiral.cpp:1664-1743, src/testing.cpp:40-105
//   scalToMat (+special_distribute)         src/spiral.cpp:1834-1885
//   regevToGSW                              src/spiral.cpp:1985-2025
//   GSW negation, cpu_crt_to_ucompressed_and_ntt   src/spiral.cpp:2361-2378, :597-609
//   regevToSimpleGsw, GSW negation (Pack)   src/testing.cpp:108-140, :1027-1032
#include "kernels.cuh"
#include "ntt.cuh"
#include <vector>

namespace sb200 {

__device__ __forceinline__ uint64_t pack_pb2(uint32_t p, uint32_t b) { return (uint64_t)p | ((uint64_t)b << 32); }

// ============================================================================================
// Expansion, one round = three kernels over the round's ACTIVE ciphertexts (host-built list):
//   k_expand_prep   : (slot,row) cv[i] (or x^(-2^r) * cv[i - 2^r], stored) -> INTT -> CRT ->
//                     automorph (negation as Q - a, 0 -> Q kept) -> row 0: raw to c0_raw[slot]
//                                                                  row 1: NTT to c1_ntt[slot]
//   k_expand_digits : (slot,k)   digit k of c0_raw[slot] -> NTT -> ginv[slot][k]
//   k_expand_accum  : cv[i][row] += sum_k W[row][k] * ginv[slot][k] + row * c1_ntt[slot]
// ============================================================================================
__global__ void __launch_bounds__(kNttThreads, 4) k_expand_prep(uint32_t *__restrict__ cv, const int *__restrict__ active, int num_in,
                                                             const uint32_t *__restrict__ neg1, const uint32_t *__restrict__ neg1_shoup, uint32_t tpow,
                                                             const uint16_t *__restrict__ perm, uint64_t *__restrict__ c0_raw, uint32_t *__restrict__ c1_ntt,
                                                             int store_self) {
    pdl_begin();
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    const int n = plane_of_thread(), lt = lane_in_plane();
    // constant during a query, loaded while the previous kernel is still running: the active list, the round's neg1 (+ Shoup
    // companion), the inverse twiddles
    const int slot = blockIdx.x, row = blockIdx.y, i = active[slot];
    const uint32_t q = modulus(n);
    uint32_t v[16], w[16], ws[16];
    load_ntt_regs(w, neg1 + n * kN, lt);
    load_ntt_regs(ws, neg1_shoup + n * kN, lt);
    // row 1: the slot permutation of this thread's 16 output positions (the loop below used to fetch them one by one, a chain of 16
    // dependent L2 round trips - 5 us, the critical path of the whole kernel)
    uint16_t pm[16];
    if (row == 0) prefetch_twiddles(c_ntt.inv[n], lt);
    else {
#pragma unroll
        for (int k = 0; k < 16; k++) pm[k] = __ldg(perm + lt + kPlaneThreads * k);
    }
    pdl_wait();
    if (i < num_in) {
        // the reference writes cv[2^r + i] = x^(-2^r) * cv[i] whenever i is processed, even if 2^r + i
        // itself is skipped later in the round (src/spiral.cpp:1709) - so the producer stores it
        // neg1 is a constant of the round: the product is a Shoup multiplication (3 multiplies) instead of a 64-bit Barrett
        uint32_t nb[16];
        load_ntt_regs(v, cv + (((size_t)i * 2 + row) * 2 + n) * kN, lt);
#pragma unroll
        for (int e = 0; e < 16; e++) nb[e] = csub(mul_shoup_lazy(v[e], w[e], ws[e], q), q);
        store_ntt_regs(nb, cv + (((size_t)(i + num_in) * 2 + row) * 2 + n) * kN, lt);
    } else {
        // same product recomputed locally: no dependence on the producer CTA of this launch
        uint32_t a[16];
        load_ntt_regs(a, cv + (((size_t)(i - num_in) * 2 + row) * 2 + n) * kN, lt);
#pragma unroll
        for (int e = 0; e < 16; e++) v[e] = csub(mul_shoup_lazy(a[e], w[e], ws[e], q), q);
        trace_mark(1);
        // a chain that follows only its own subtree (odd / even graphs, a rank's share of a sharded expansion) may process
        // output i without its sibling i - num_in: then nobody else stores the base value the accumulation adds to
        if (store_self) store_ntt_regs(v, cv + (((size_t)i * 2 + row) * 2 + n) * kN, lt);
    }
    if (row == 1) {
        // NTT(automorph(c_1)) is a permutation of the NTT slots of c_1 (x -> x^t maps evaluation point
        // psi^e to psi^(e*t)): no inverse / forward transform needed, identical values modulo q.
        plane_sync(n);
#pragma unroll
        for (int k = 0; k < 16; k++) sm[n][ntt_pos(lt, k)] = v[k];
        plane_sync(n);
        uint32_t *dst = c1_ntt + ((size_t)slot * 2 + n) * kN;
#pragma unroll
        for (int k = 0; k < 16; k++) dst[lt + kPlaneThreads * k] = sm[n][pm[k]];
        return;
    }
    trace_mark(2);
    ntt_inverse_plane(v, sm[n], lt, n);
    trace_mark(3);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) sm[n][nat_pos(lt, k)] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int z = threadIdx.x + 256 * k;
        uint64_t val = crt_compose(sm[0][z], sm[1][z]);
        const uint32_t it = (uint32_t)z * tpow;
        const int rem = (int)(it & (kN - 1));
        if ((it >> kLogN) & 1) val = kQ - val;                 // 0 -> Q on purpose (reference src/poly.cpp:256)
        c0_raw[(size_t)slot * kN + rem] = val;
    }
    trace_mark(4);
}

__global__ void __launch_bounds__(kNttThreads) k_expand_digits(uint32_t *__restrict__ ginv, const uint64_t *__restrict__ c0_raw,
                                                               const int *__restrict__ active, int t_left, int t_right, int tmax, int cnt, int k_begin) {
    pdl_begin();
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    const int n = plane_of_thread(), lt = lane_in_plane();
    prefetch_twiddles(c_ntt.fwd[n], lt);                   // constant: fetched while the previous kernel is still running
    pdl_wait();
    // grid (cnt, tmax), or - for a chain that must not crowd out a concurrent, more urgent one - a smaller 1-D grid whose CTAs
    // walk the cnt x tmax items: the kernel then holds a fixed number of CTA slots instead of queueing thousands of CTAs ahead
    // of the other chain's
    // k_begin >= 0: 2-D grid (cnt, digits of this launch), digit k = k_begin + blockIdx.y
    const int total = cnt * tmax;
    for (int item = k_begin >= 0 ? (int)(blockIdx.x * tmax + k_begin + blockIdx.y) : (int)blockIdx.x; item < total; item += k_begin >= 0 ? total : (int)gridDim.x) {
        const int slot = item / tmax, k = item % tmax, i = active[slot];
        const int gd = (i & 1) ? t_right : t_left;
        if (k >= gd) continue;
        const uint32_t bits_per = get_bits_per(gd);
        const uint64_t mask = (1ull << bits_per) - 1;
        const uint64_t *src = c0_raw + (size_t)slot * kN;
        uint32_t v[16];
#pragma unroll
        for (int e = 0; e < 16; e++) {
            const uint64_t d = gadget_digit(__ldg(src + nat_pos(lt, e)), k, bits_per, mask);
            v[e] = bits_per >= 28 ? raw_to_res(d, n) : (uint32_t)d;
        }
        trace_mark(5);
        ntt_forward_plane(v, sm[n], lt, n);
        trace_mark(6);
        store_ntt_regs(v, ginv + (((size_t)slot * tmax + k) * 2 + n) * kN, lt);
        trace_mark(7);
    }
}

__global__ void __launch_bounds__(256) k_expand_accum(uint32_t *__restrict__ cv, const int *__restrict__ active, const uint32_t *__restrict__ ginv,
                               const uint32_t *__restrict__ c1_ntt, const uint32_t *__restrict__ W_left,
                               const uint32_t *__restrict__ W_right, int t_left, int t_right, int tmax) {
    pdl_begin();
    // grid (slot, row*16 + segment); CTA = 64 uint4 columns x 4 digit groups (same split as k_fold_mac: every
    // thread's loads are independent, partial sums meet in shared memory)
    __shared__ ulonglong2 part[3][64][2];
    const int slot = blockIdx.x, row = blockIdx.y >> 4, seg = blockIdx.y & 15, i = active[slot];
    const int col = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const int w4 = seg * 64 + col, n = w4 >= 512;
    const int gd = (i & 1) ? t_right : t_left;
    const uint4 *W = reinterpret_cast<const uint4 *>(((i & 1) ? W_right : W_left) + (size_t)row * gd * 2 * kN) + w4;
    const uint4 *G = reinterpret_cast<const uint4 *>(ginv + (size_t)slot * tmax * 2 * kN) + w4;
    uint64_t acc[4] = {0, 0, 0, 0};
    if (gd <= 8) {
        // short chains (t_left = 8): this thread's key words (constant during a query) are in registers before the digits exist
        uint4 x0 = make_uint4(0, 0, 0, 0), x1 = x0;
        if (grp < gd) x0 = __ldg(W + (size_t)grp * (2 * kN / 4));
        if (grp + 4 < gd) x1 = __ldg(W + (size_t)(grp + 4) * (2 * kN / 4));
        pdl_wait();
        if (grp < gd) {
            const uint4 y = __ldg(G + (size_t)grp * (2 * kN / 4));
            acc[0] += (uint64_t)x0.x * y.x; acc[1] += (uint64_t)x0.y * y.y; acc[2] += (uint64_t)x0.z * y.z; acc[3] += (uint64_t)x0.w * y.w;
        }
        if (grp + 4 < gd) {
            const uint4 y = __ldg(G + (size_t)(grp + 4) * (2 * kN / 4));
            acc[0] += (uint64_t)x1.x * y.x; acc[1] += (uint64_t)x1.y * y.y; acc[2] += (uint64_t)x1.z * y.z; acc[3] += (uint64_t)x1.w * y.w;
        }
    } else {
        pdl_wait();
#pragma unroll 4
        for (int k = grp; k < gd; k += 4) {
            const uint4 x = __ldg(W + (size_t)k * (2 * kN / 4)), y = __ldg(G + (size_t)k * (2 * kN / 4));
            acc[0] += (uint64_t)x.x * y.x; acc[1] += (uint64_t)x.y * y.y;
            acc[2] += (uint64_t)x.z * y.z; acc[3] += (uint64_t)x.w * y.w;
        }
    }
    if (grp > 0) {
        part[grp - 1][col][0] = make_ulonglong2(acc[0], acc[1]);
        part[grp - 1][col][1] = make_ulonglong2(acc[2], acc[3]);
    }
    __syncthreads();
    if (grp != 0) return;
#pragma unroll
    for (int g = 0; g < 3; g++) {
        const ulonglong2 a = part[g][col][0], b = part[g][col][1];
        acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
    }
    uint4 *dst = reinterpret_cast<uint4 *>(cv + ((size_t)i * 2 + row) * 2 * kN) + w4;
    const uint4 cur = *dst;
    uint4 add = make_uint4(0, 0, 0, 0);
    if (row == 1) add = __ldg(reinterpret_cast<const uint4 *>(c1_ntt + (size_t)slot * 2 * kN) + w4);
    uint4 o;
    o.x = reduce_u64(acc[0] + cur.x + add.x, n); o.y = reduce_u64(acc[1] + cur.y + add.y, n);
    o.z = reduce_u64(acc[2] + cur.z + add.z, n); o.w = reduce_u64(acc[3] + cur.w + add.w, n);
    *dst = o;
}

// Throughput variant for rounds with many active ciphertexts: CTA = 256 uint4 columns of one slot, every thread walks all
// digits of its column for BOTH rows, so each digit word is loaded once (not once per row) and there is no partial-sum exchange.
__global__ void __launch_bounds__(256) k_expand_accum_wide(uint32_t *__restrict__ cv, const int *__restrict__ active, const uint32_t *__restrict__ ginv,
                                                           const uint32_t *__restrict__ c1_ntt, const uint32_t *__restrict__ W_left,
                                                           const uint32_t *__restrict__ W_right, int t_left, int t_right, int tmax) {
    pdl_prologue();
    const int slot = blockIdx.x, seg = blockIdx.y, i = active[slot];
    const int w4 = seg * 256 + threadIdx.x, n = w4 >= 512;
    const int gd = (i & 1) ? t_right : t_left;
    const size_t ps = 2 * kN / 4;                                     // uint4 per polynomial
    const uint4 *W0 = reinterpret_cast<const uint4 *>((i & 1) ? W_right : W_left) + w4;
    const uint4 *W1 = W0 + (size_t)gd * ps;
    const uint4 *G = reinterpret_cast<const uint4 *>(ginv + (size_t)slot * tmax * 2 * kN) + w4;
    uint64_t a0[4] = {0, 0, 0, 0}, a1[4] = {0, 0, 0, 0};
#pragma unroll 4
    for (int k = 0; k < gd; k++) {                                    // gd <= 56 products of < 2^56: no intermediate reduction
        const uint4 y = __ldg(G + (size_t)k * ps), x0 = __ldg(W0 + (size_t)k * ps), x1 = __ldg(W1 + (size_t)k * ps);
        a0[0] += (uint64_t)x0.x * y.x; a0[1] += (uint64_t)x0.y * y.y; a0[2] += (uint64_t)x0.z * y.z; a0[3] += (uint64_t)x0.w * y.w;
        a1[0] += (uint64_t)x1.x * y.x; a1[1] += (uint64_t)x1.y * y.y; a1[2] += (uint64_t)x1.z * y.z; a1[3] += (uint64_t)x1.w * y.w;
    }
    uint4 *d0 = reinterpret_cast<uint4 *>(cv + ((size_t)i * 2 + 0) * 2 * kN) + w4, *d1 = d0 + ps;
    const uint4 c0 = *d0, c1 = *d1, add = __ldg(reinterpret_cast<const uint4 *>(c1_ntt + (size_t)slot * 2 * kN) + w4);
    *d0 = make_uint4(reduce_u64(a0[0] + c0.x, n), reduce_u64(a0[1] + c0.y, n), reduce_u64(a0[2] + c0.z, n), reduce_u64(a0[3] + c0.w, n));
    *d1 = make_uint4(reduce_u64(a1[0] + c1.x + add.x, n), reduce_u64(a1[1] + c1.y + add.y, n), reduce_u64(a1[2] + c1.z + add.z, n),
                     reduce_u64(a1[3] + c1.w + add.w, n));
}

// neg1[r] = NTT(invert(x^(N - 2^r))) = NTT(-x^(N-2^r))   (reference src/spiral.cpp:184-192), followed at + count polynomials by
// the Shoup companions floor(neg1 * 2^32 / q)
__global__ void __launch_bounds__(kNttThreads) k_build_neg1(uint32_t *__restrict__ neg1) {
    pdl_prologue();
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    const int n = plane_of_thread(), lt = lane_in_plane(), r = blockIdx.x;
    const uint32_t q = modulus(n);
    const int idx = kN - (1 << r);
    uint32_t v[16];
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = (nat_pos(lt, k) == idx) ? q - 1 : 0;
    ntt_forward_plane(v, sm[n], lt, n);
    store_ntt_regs(v, neg1 + ((size_t)r * 2 + n) * kN, lt);
    uint32_t ws[16];
#pragma unroll
    for (int k = 0; k < 16; k++) ws[k] = (uint32_t)(((uint64_t)v[k] << 32) / q);
    store_ntt_regs(ws, neg1 + ((size_t)(gridDim.x + r) * 2 + n) * kN, lt);
}
void build_neg1(uint32_t *neg1_dev, int count, cudaStream_t s) {   // neg1_dev: 2 * count polynomials
    if (count) { count_launch(); launch_pdl(k_build_neg1, dim3(count), dim3(kNttThreads), 0, s, neg1_dev); }
}

// host side: per-round active lists, exactly the reference's skip rules (src/spiral.cpp:1701-1702)
static int build_active(int *dst, int r, int stopround, int max_bits_right) {
    const int num_out = 2 << r;
    int cnt = 0;
    for (int i = 0; i < num_out; i++) {
        if (stopround > 0 && r > stopround && (i % 2) == 1) continue;
        if (stopround > 0 && r == stopround && (i % 2) == 1 && i / 2 > max_bits_right) continue;
        dst[cnt++] = i;
    }
    return cnt;
}
size_t expand_active_total(const ExpandPlan &p) {
    size_t total = 0;
    for (int r = 0; r < p.g; r++) total += (size_t)2 << r;
    return total;
}
// fills host arrays: list (concatenated), offs[r], cnt[r]; returns the max count
int expand_build_lists(const ExpandPlan &p, int *list, int *offs, int *cnt) {
    int o = 0, mx = 0;
    for (int r = 0; r < p.g; r++) {
        offs[r] = o;
        cnt[r] = build_active(list + o, r, p.stopround, p.max_bits_right);
        if (cnt[r] > mx) mx = cnt[r];
        o += cnt[r];
    }
    return mx;
}

// perm[r][pos] = NTT slot whose value lands in slot `pos` after the round-r automorphism x -> x^(N/2^r + 1).
// Host-side: slot `pos` of the reference NTT order holds the evaluation at psi^e(pos); e() is read off a
// transform of the polynomial x computed with the same butterfly network as ntt.cuh.
void build_automorph_perms(uint16_t *perm_host, int g) {
    const uint64_t q = kP, psi = kPsiP;
    auto mulm = [&](uint64_t a, uint64_t b) { return (uint64_t)((unsigned __int128)a * b % q); };
    std::vector<uint64_t> pw(2 * kN);
    pw[0] = 1;
    for (int i = 1; i < 2 * kN; i++) pw[i] = mulm(pw[i - 1], psi);
    auto brev = [](uint32_t x) { uint32_t r = 0; for (int i = 0; i < kLogN; i++) { r = (r << 1) | (x & 1); x >>= 1; } return r; };
    // forward transform of a(x) = x, Cooley-Tukey with w[m + i] = psi^bitrev11(m + i) (reference src/core.cpp:254-270)
    std::vector<uint64_t> a(kN, 0);
    a[1] = 1;
    for (int mm = 0; mm < kLogN; mm++) {
        const int m = 1 << mm, t = kN >> (mm + 1);
        for (int i = 0; i < m; i++) {
            const uint64_t w = pw[brev((uint32_t)(m + i))];
            for (int j = 0; j < t; j++) {
                uint64_t &x = a[2 * i * t + j], &y = a[2 * i * t + t + j];
                const uint64_t wy = mulm(w, y), nx = (x + wy) % q, ny = (x + q - wy) % q;
                x = nx; y = ny;
            }
        }
    }
    std::vector<int> expo_of_slot(kN), slot_of_expo(2 * kN, -1);
    for (int e = 1; e < 2 * kN; e += 2) {
        for (int pos = 0; pos < kN; pos++) if (a[pos] == pw[e]) { expo_of_slot[pos] = e; slot_of_expo[e] = pos; break; }
    }
    for (int r = 0; r < g; r++) {
        const uint32_t t = (uint32_t)(kN >> r) + 1;
        for (int pos = 0; pos < kN; pos++) perm_host[(size_t)r * kN + pos] = (uint16_t)slot_of_expo[(expo_of_slot[pos] * (uint64_t)t) % (2 * kN)];
    }
}

// polynomials the digit scratch must hold: max over rounds of (active ciphertexts x digits per slot of that round)
// even / odd halves of the per-round lists (rounds >= 1; round 0 stays whole).  Returns the largest odd count.
int expand_split_lists(const ExpandPlan &p, const int *list, const int *offs, const int *cnt, int *list_e, int *offs_e, int *cnt_e,
                       int *list_o, int *offs_o, int *cnt_o) {
    int oe = 0, oo = 0, mx = 0;
    for (int r = 0; r < p.g; r++) {
        offs_e[r] = oe; offs_o[r] = oo; cnt_e[r] = cnt_o[r] = 0;
        for (int k = 0; k < cnt[r]; k++) {
            const int i = list[offs[r] + k];
            if (i & 1) { list_o[oo++] = i; cnt_o[r]++; } else { list_e[oe++] = i; cnt_e[r]++; }
        }
        if (cnt_o[r] > mx) mx = cnt_o[r];
    }
    return mx;
}
size_t expand_ginv_polys(const ExpandPlan &p, const int *cnt) {
    const int tmax = p.t_left > p.t_right ? p.t_left : p.t_right;
    size_t need = 1;
    for (int r = 0; r < p.g; r++) {
        const bool any_odd = !(p.stopround > 0 && r > p.stopround);
        const size_t polys = (size_t)cnt[r] * (any_odd ? tmax : p.t_left);
        if (polys > need) need = polys;
    }
    return need;
}
void launch_expand(uint32_t *cv, const ExpandPlan &p, const uint32_t *W_left, const uint32_t *W_right,
                   const uint32_t *neg1, const uint16_t *perms, uint64_t *c0_raw, uint32_t *c1_ntt, uint32_t *ginv,
                   const int *list_dev, const int *offs, const int *cnt, cudaStream_t s, int r_begin, int r_end, int parity, int store_self, int slot_limit, int digit_chunk) {
    // parity: -1 = the lists hold every active ciphertext; 0 / 1 = they hold only the even / odd ones (expand_split_lists):
    // after round 0 the even chain (first-dimension ciphertexts, t_left digits) and the odd chain (GSW bits, t_right digits)
    // never touch each other's ciphertexts, so the two can run on different streams with their own scratch.
    const int tmax = p.t_left > p.t_right ? p.t_left : p.t_right;
    if (r_end < 0 || r_end > p.g) r_end = p.g;
    for (int r = r_begin; r < r_end; r++) {
        if (!cnt[r]) continue;
        const int *act = list_dev + offs[r];
        const uint32_t tpow = (uint32_t)(kN >> r) + 1;
        const uint32_t *Wl = W_left + (size_t)r * 2 * p.t_left * 2 * kN;
        const uint32_t *Wr = W_right + (size_t)r * 2 * p.t_right * 2 * kN;
        // digits needed this round: t_right only if some odd ciphertext is active
        const bool any_odd = parity != 0 && !(p.stopround > 0 && r > p.stopround);
        const int ty = parity == 1 ? p.t_right : any_odd ? tmax : p.t_left;
        count_launch(); launch_pdl(k_expand_prep, dim3(dim3(cnt[r], 2)), dim3(kNttThreads), 0, s, cv, act, 1 << r, neg1 + (size_t)r * 2 * kN, neg1 + (size_t)(p.g + r) * 2 * kN, tpow, perms + (size_t)r * kN, c0_raw, c1_ntt, store_self);
        // ginv is indexed [slot][ty]: rounds past stopround only hold t_left digits per slot (see expand_ginv_polys)
        if (slot_limit > 0 && cnt[r] * ty > slot_limit) {
            count_launch(); launch_pdl(k_expand_digits, dim3(slot_limit), dim3(kNttThreads), 0, s, ginv, c0_raw, act, p.t_left, p.t_right, ty, cnt[r], -1);
        } else {
            // digit_chunk > 0: a round with thousands of digit NTTs goes out as several launches of about digit_chunk CTAs, so that
            // a concurrent chain's kernels wait behind one wave of this chain's CTAs, not behind all of them
            int per = ty;
            if (digit_chunk > 0 && cnt[r] * ty > 2 * digit_chunk) { per = digit_chunk / cnt[r]; if (per < 1) per = 1; }
            for (int k0 = 0; k0 < ty; k0 += per) {
                const int kn = k0 + per < ty ? per : ty - k0;
                count_launch(); launch_pdl(k_expand_digits, dim3(dim3(cnt[r], kn)), dim3(kNttThreads), 0, s, ginv, c0_raw, act, p.t_left, p.t_right, ty, cnt[r], k0);
            }
        }
        count_launch();                                   // rounds with right slots (56-term chains) stay on the split kernel
        if (((cnt[r] >= 64 && !any_odd) || cnt[r] >= 512) && tmax <= 128) launch_pdl(k_expand_accum_wide, dim3(dim3(cnt[r], 4)), dim3(256), 0, s, cv, act, ginv, c1_ntt, Wl, Wr, p.t_left, p.t_right, ty);
        else launch_pdl(k_expand_accum, dim3(dim3(cnt[r], 32)), dim3(256), 0, s, cv, act, ginv, c1_ntt, Wl, Wr, p.t_left, p.t_right, ty);
    }
}

// ============================================================================================
// Conversion.  Shared front end: INTT+CRT of selected polynomials, then NTT'd gadget digits.
// ============================================================================================
// raw[slot] = from_ntt(poly src[poly_idx[slot]])
__global__ void __launch_bounds__(kNttThreads) k_from_ntt_indexed(uint64_t *__restrict__ raw, const uint32_t *__restrict__ in,
                                                                  const int *__restrict__ poly_idx) {
    pdl_begin();
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    const int n = plane_of_thread(), lt = lane_in_plane();
    const int poly = poly_idx[blockIdx.x];                 // constant index list: read before the previous kernel is done
    prefetch_twiddles(c_ntt.inv[n], lt);
    pdl_wait();
    uint32_t v[16];
    load_ntt_regs(v, in + ((size_t)poly * 2 + n) * kN, lt);
    ntt_inverse_plane(v, sm[n], lt, n);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) sm[n][nat_pos(lt, k)] = v[k];
    __syncthreads();
    uint64_t *dst = raw + (size_t)blockIdx.x * kN;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int z = threadIdx.x + 256 * k;
        dst[z] = crt_compose(sm[0][z], sm[1][z]);
    }
}
void launch_from_ntt_indexed(uint64_t *raw, const uint32_t *in, const int *poly_idx, size_t count, cudaStream_t s) {
    if (count) { count_launch(); launch_pdl(k_from_ntt_indexed, dim3((unsigned)count), dim3(kNttThreads), 0, s, raw, in, poly_idx); }
}

// scalToMat for dim0 ciphertexts, written straight into the scan's query layout
// query[z][j][m=c][r] (reorientCiphertexts fused):   out[r][c] = sum_k W[r][2k+c]*ginv[k][j] + [r == c+1] cv1
// ginv: [t_conv][dim0] polys ; cv row 1 taken from cv[ct_idx[j]]
__global__ void k_scal_to_mat_accum(uint64_t *__restrict__ query, const uint32_t *__restrict__ cv, const int *__restrict__ ct_idx,
                                    const uint32_t *__restrict__ ginv, const uint32_t *__restrict__ W, int t_conv, int dim0) {
    pdl_prologue();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // (j, z), z fastest
    if (idx >= (size_t)dim0 * kN) return;
    const int z = (int)(idx % kN), j = (int)(idx / kN);
    uint64_t acc[3][2][2];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 2; c++) acc[r][c][0] = acc[r][c][1] = 0;
    const int wc = 2 * t_conv;
    for (int k = 0; k < t_conv; k++) {
        const uint32_t *g = ginv + ((size_t)k * dim0 + j) * 2 * kN;
        const uint32_t gp = g[z], gb = g[kN + z];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 2; c++) {
                const uint32_t *w = W + ((size_t)r * wc + 2 * k + c) * 2 * kN;
                acc[r][c][0] += (uint64_t)w[z] * gp;
                acc[r][c][1] += (uint64_t)w[kN + z] * gb;
            }
        if ((k & 127) == 127) {
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 2; c++) { acc[r][c][0] = reduce_u64(acc[r][c][0], 0); acc[r][c][1] = reduce_u64(acc[r][c][1], 1); }
        }
    }
    const uint32_t *cv1 = cv + ((size_t)ct_idx[j] * 2 + 1) * 2 * kN;
    const uint32_t c1p = cv1[z], c1b = cv1[kN + z];
    acc[1][0][0] += c1p; acc[1][0][1] += c1b;     // place(cv_1, 1, 0)
    acc[2][1][0] += c1p; acc[2][1][1] += c1b;     // place(cv_1, 2, 1)
#pragma unroll
    for (int c = 0; c < 2; c++) {
        uint64_t w[4];
#pragma unroll
        for (int r = 0; r < 3; r++) w[r] = pack_pb2(reduce_u64(acc[r][c][0], 0), reduce_u64(acc[r][c][1], 1));
        w[3] = 0;
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(query + (((size_t)z * dim0 + j) * 2 + c) * 4);
        dst[0] = make_ulonglong2(w[0], w[1]);
        dst[1] = make_ulonglong2(w[2], w[3]);
    }
}
// The same, tiled: CTA = 32 z x 8 j.  Operands are read with z across the lanes (coalesced), results go through shared memory
// and leave with j across the lanes: 512 contiguous bytes per z instead of 64-byte pieces 16 KiB apart (the scattered form spent
// its time in the store path: 2 M sixteen-byte stores to 2 M different sectors).
// Sharded expansion: this rank converts only the `count` ciphertexts jl whose first-dimension index is j = j_off + j_stride * jl and
// stores them into the query buffer of EVERY rank (peer pointers: the all-gather is fused into the kernel that produces the
// data); the last CTA to finish raises this rank's flag on every peer (k_query_wait is the consumer side).
struct ScalTargets {
    uint64_t *query[16];                // [0] = this rank's own buffer; ntargets == 1: unsharded
    unsigned int *flag[16];             // flag[t]: this rank's arrival flag inside target t's exchange header
    int ntargets;
    unsigned int *arrive;               // this rank's CTA arrival counter
    const unsigned int *ack;            // rank 0's acknowledgement of the previous query, as seen by this rank
    const unsigned int *epoch;          // exchange epoch of the last completed query
    unsigned int *error;
};
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__global__ void __launch_bounds__(256) k_scal_to_mat_accum_tiled(const __grid_constant__ ScalTargets tg, const uint32_t *__restrict__ cv, const int *__restrict__ ct_idx,
                                                                 const uint32_t *__restrict__ ginv, const uint32_t *__restrict__ W, int dim0, int count,
                                                                 int j_off, int j_stride, int jtiles) {
    pdl_prologue_no_early_dependents();
    constexpr int TC = 4;                                             // t_conv of every Spiral parameter set that reaches this kernel
    constexpr int kRowWords = 8 * 8 + 1;                              // 8 j x 8 words, padded: lanes (z) land on different banks
    __shared__ __align__(16) uint64_t tile[32 * kRowWords];
    __shared__ int ok;
    const int zl = threadIdx.x & 31, jl = threadIdx.x >> 5, lane = zl, wid = jl;
    const int z = blockIdx.x * 32 + zl;
    unsigned int e = 0;
    if (tg.ntargets > 1) {
        // the peers' query buffers may be overwritten only after every rank has finished scanning the previous query: rank 0's
        // acknowledgement of that query's exchange (bounded spin, as in xchg_kernels.cu)
        if (threadIdx.x == 0) {
            e = *tg.epoch + 1;
            ok = 1;
            if (e > 1) {
                unsigned long long t0 = global_timer_ns();
                while ((int)(ld_acquire_sys_u32(tg.ack) - (e - 1)) < 0) {
                    __nanosleep(200);
                    if (global_timer_ns() - t0 > 4000000000ull) { ok = 0; break; }
                }
            }
        }
        __syncthreads();
        if (!ok) { if (threadIdx.x == 0) *tg.error = 3; return; }
    }
    // the conversion key W (3 x 2*t_conv) at this z under both primes stays in registers for all the j of this CTA
    uint32_t wp[3][2][TC], wb[3][2][TC];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
            for (int k = 0; k < TC; k++) {
                const uint32_t *w = W + ((size_t)r * (2 * TC) + 2 * k + c) * 2 * kN;
                wp[r][c][k] = __ldg(w + z); wb[r][c][k] = __ldg(w + kN + z);
            }
    for (int jt = 0; jt < jtiles; jt++) {
        const int j0 = (blockIdx.y * jtiles + jt) * 8, j = j0 + jl;           // local ciphertext numbers
        uint32_t gp[TC], gb[TC];
#pragma unroll
        for (int k = 0; k < TC; k++) { const uint32_t *g = ginv + ((size_t)k * count + j) * 2 * kN; gp[k] = __ldg(g + z); gb[k] = __ldg(g + kN + z); }
        const uint32_t *cv1 = cv + ((size_t)ct_idx[j] * 2 + 1) * 2 * kN;
        const uint32_t c1p = __ldg(cv1 + z), c1b = __ldg(cv1 + kN + z);
        uint64_t *row = tile + zl * kRowWords + jl * 8;
#pragma unroll
        for (int c = 0; c < 2; c++) {
#pragma unroll
            for (int r = 0; r < 3; r++) {
                uint64_t ap = (r == c + 1) ? c1p : 0, ab = (r == c + 1) ? c1b : 0;      // place(cv_1, 1, 0), place(cv_1, 2, 1)
#pragma unroll
                for (int k = 0; k < TC; k++) { ap += (uint64_t)wp[r][c][k] * gp[k]; ab += (uint64_t)wb[r][c][k] * gb[k]; }
                row[c * 4 + r] = pack_pb2(reduce_u64(ap, 0), reduce_u64(ab, 1));
            }
            row[c * 4 + 3] = 0;
        }
        __syncthreads();
        // write-out: warp w takes rows z = w, w + 8, ...; lane l the 16-byte chunk l of the 8 x 64 bytes [z][j(j0 .. j0+8)][m][4]
        // (one 512-byte run when j_stride == 1)
        const size_t jg = (size_t)j_off + (size_t)j_stride * (j0 + (lane >> 2));
#pragma unroll
        for (int zz = wid; zz < 32; zz += 8) {
            const uint64_t *src = tile + zz * kRowWords + lane * 2;
            const ulonglong2 val = make_ulonglong2(src[0], src[1]);
            const size_t o = (((size_t)(blockIdx.x * 32 + zz) * dim0 + jg) * 2) * 4 / 2 + (lane & 3);     // in 16-byte units
            for (int t = 0; t < tg.ntargets; t++) reinterpret_cast<ulonglong2 *>(tg.query[t])[o] = val;
        }
        __syncthreads();
    }
    if (tg.ntargets > 1) {
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(tg.arrive, 1u) == gridDim.x * gridDim.y - 1) {
                *tg.arrive = 0;
                __threadfence_system();
                for (int t = 0; t < tg.ntargets; t++) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(tg.flag[t]), "r"(e) : "memory");
            }
        }
    }
}
// same product but emitted as dev-NTT MatPoly (3 x 2) per ciphertext - the reference's scalToMat output
__global__ void k_scal_to_mat_ntt(uint32_t *__restrict__ out, const uint32_t *__restrict__ cv, const int *__restrict__ ct_idx,
                                  const uint32_t *__restrict__ ginv, const uint32_t *__restrict__ W, int t_conv, int count) {
    pdl_prologue();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // (j, n, z)
    if (idx >= (size_t)count * 2 * kN) return;
    const int nz = (int)(idx % (2 * kN)), j = (int)(idx / (2 * kN)), n = nz >= kN;
    const int wc = 2 * t_conv;
    uint64_t acc[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    for (int k = 0; k < t_conv; k++) {
        const uint32_t g = ginv[((size_t)k * count + j) * 2 * kN + nz];
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
            for (int c = 0; c < 2; c++) acc[r][c] += (uint64_t)W[((size_t)r * wc + 2 * k + c) * 2 * kN + nz] * g;
        if ((k & 127) == 127) {
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 2; c++) acc[r][c] = reduce_u64(acc[r][c], n);
        }
    }
    const uint32_t c1 = cv[((size_t)ct_idx[j] * 2 + 1) * 2 * kN + nz];
    acc[1][0] += c1; acc[2][1] += c1;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 2; c++) out[(((size_t)j * 3 + r) * 2 + c) * 2 * kN + nz] = reduce_u64(acc[r][c], n);
}

// regevToGSW accumulate: for GSW ciphertext d (stored at slot dslot = nu2-1-d) and digit index jj < ell
//   col 3jj     = sum_k V[r][k]*g0[k] + V[r][t_conv+k]*g1[k]
//   col 3jj+1+c = sum_k W[r][2k+c]*g0[k] + [r == c+1] cv1
// g0/g1: NTT'd digits of rows 0/1 of cv (bit index b = d*ell + jj): ginv[row][k][b]
__global__ void k_regev_to_gsw_accum(uint32_t *__restrict__ gsw, const uint32_t *__restrict__ cv, const int *__restrict__ ct_idx,
                                     const uint32_t *__restrict__ ginv, const uint32_t *__restrict__ W, const uint32_t *__restrict__ V,
                                     int t_conv, int ell, int nu2) {
    pdl_prologue();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // (b, n, z)
    const int nbits = ell * nu2;
    if (idx >= (size_t)nbits * 2 * kN) return;
    const int nz = (int)(idx % (2 * kN)), b = (int)(idx / (2 * kN)), n = nz >= kN;
    const int d = b / ell, jj = b % ell, wc = 2 * t_conv, m2 = 3 * ell;
    uint64_t s2m[3][2] = {{0, 0}, {0, 0}, {0, 0}}, pr[3] = {0, 0, 0};
    for (int k = 0; k < t_conv; k++) {
        const uint32_t g0 = ginv[(((size_t)0 * t_conv + k) * nbits + b) * 2 * kN + nz];
        const uint32_t g1 = ginv[(((size_t)1 * t_conv + k) * nbits + b) * 2 * kN + nz];
#pragma unroll
        for (int r = 0; r < 3; r++) {
            pr[r] += (uint64_t)V[((size_t)r * wc + k) * 2 * kN + nz] * g0 + (uint64_t)V[((size_t)r * wc + t_conv + k) * 2 * kN + nz] * g1;
#pragma unroll
            for (int c = 0; c < 2; c++) s2m[r][c] += (uint64_t)W[((size_t)r * wc + 2 * k + c) * 2 * kN + nz] * g0;
        }
        if ((k & 63) == 63) {
#pragma unroll
            for (int r = 0; r < 3; r++) {
                pr[r] = reduce_u64(pr[r], n);
#pragma unroll
                for (int c = 0; c < 2; c++) s2m[r][c] = reduce_u64(s2m[r][c], n);
            }
        }
    }
    const uint32_t c1 = cv[((size_t)ct_idx[b] * 2 + 1) * 2 * kN + nz];
    s2m[1][0] += c1; s2m[2][1] += c1;
    uint32_t *o = gsw + (size_t)(nu2 - 1 - d) * 3 * m2 * 2 * kN;
#pragma unroll
    for (int r = 0; r < 3; r++) {
        o[((size_t)r * m2 + 3 * jj) * 2 * kN + nz] = reduce_u64(pr[r], n);
        o[((size_t)r * m2 + 3 * jj + 1) * 2 * kN + nz] = reduce_u64(s2m[r][0], n);
        o[((size_t)r * m2 + 3 * jj + 2) * 2 * kN + nz] = reduce_u64(s2m[r][1], n);
    }
}

// The same for a rank's share of a sharded conversion: `count` bits, local bit l is the global bit bit_ids[l] (the ciphertexts of
// the odd chain this rank expanded); the 9 words of every (bit, slot) go into the GSW buffer of EVERY rank (peer pointers), and
// the last CTA raises this rank's flag on each (gflags; k_flag_wait before the folds is the consumer side).
struct GswTargets {
    uint32_t *gsw[16];
    unsigned int *flag[16];
    int ntargets;
    unsigned int *arrive;
    const unsigned int *ack, *epoch;
    unsigned int *error;
};
__global__ void __launch_bounds__(256) k_regev_to_gsw_accum_sharded(const __grid_constant__ GswTargets tg, const uint32_t *__restrict__ cv, const int *__restrict__ ct_idx,
                                                                    const int *__restrict__ bit_ids, const uint32_t *__restrict__ ginv, const uint32_t *__restrict__ W,
                                                                    const uint32_t *__restrict__ V, int t_conv, int ell, int nu2, int count) {
    pdl_prologue_no_early_dependents();
    __shared__ int ok;
    unsigned int e = 0;
    if (threadIdx.x == 0) {        // the peers' GSW buffers are free once rank 0 has finished the previous query (its acknowledgement)
        e = *tg.epoch + 1;
        ok = 1;
        if (e > 1) {
            unsigned long long t0 = global_timer_ns();
            while ((int)(ld_acquire_sys_u32(tg.ack) - (e - 1)) < 0) {
                __nanosleep(200);
                if (global_timer_ns() - t0 > 4000000000ull) { ok = 0; break; }
            }
        }
    }
    __syncthreads();
    if (!ok) { if (threadIdx.x == 0) *tg.error = 4; return; }
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // (l, n, z)
    if (idx < (size_t)count * 2 * kN) {
        const int nz = (int)(idx % (2 * kN)), l = (int)(idx / (2 * kN)), n = nz >= kN;
        const int b = bit_ids[l], d = b / ell, jj = b % ell, wc = 2 * t_conv, m2 = 3 * ell;
        uint64_t s2m[3][2] = {{0, 0}, {0, 0}, {0, 0}}, pr[3] = {0, 0, 0};
        for (int k = 0; k < t_conv; k++) {
            const uint32_t g0 = ginv[(((size_t)0 * t_conv + k) * count + l) * 2 * kN + nz];
            const uint32_t g1 = ginv[(((size_t)1 * t_conv + k) * count + l) * 2 * kN + nz];
#pragma unroll
            for (int r = 0; r < 3; r++) {
                pr[r] += (uint64_t)V[((size_t)r * wc + k) * 2 * kN + nz] * g0 + (uint64_t)V[((size_t)r * wc + t_conv + k) * 2 * kN + nz] * g1;
#pragma unroll
                for (int c = 0; c < 2; c++) s2m[r][c] += (uint64_t)W[((size_t)r * wc + 2 * k + c) * 2 * kN + nz] * g0;
            }
            if ((k & 63) == 63) {
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    pr[r] = reduce_u64(pr[r], n);
#pragma unroll
                    for (int c = 0; c < 2; c++) s2m[r][c] = reduce_u64(s2m[r][c], n);
                }
            }
        }
        const uint32_t c1 = cv[((size_t)ct_idx[l] * 2 + 1) * 2 * kN + nz];
        s2m[1][0] += c1; s2m[2][1] += c1;
        const size_t base = (size_t)(nu2 - 1 - d) * 3 * m2 * 2 * kN;
#pragma unroll
        for (int r = 0; r < 3; r++) {
            const uint32_t v0 = reduce_u64(pr[r], n), v1 = reduce_u64(s2m[r][0], n), v2 = reduce_u64(s2m[r][1], n);
            const size_t o = base + ((size_t)r * m2 + 3 * jj) * 2 * kN + nz;
            for (int t = 0; t < tg.ntargets; t++) {
                uint32_t *g = tg.gsw[t];
                g[o] = v0; g[o + 2 * kN] = v1; g[o + 4 * kN] = v2;
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(tg.arrive, 1u) == gridDim.x - 1) {
            *tg.arrive = 0;
            __threadfence_system();
            for (int t = 0; t < tg.ntargets; t++) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(tg.flag[t]), "r"(e) : "memory");
        }
    }
}

// GSW negation: Qneg = NTT(G2 - from_ntt(Q))  per polynomial (d, r, m); G2 = buildGadget(n1, m2)
// (reference src/spiral.cpp:2361-2378; `long` subtraction, +Q when negative)
__global__ void __launch_bounds__(kNttThreads) k_gsw_negate(uint32_t *__restrict__ neg, const uint32_t *__restrict__ gsw, int ell, int rows) {
    pdl_prologue();
    __shared__ __align__(16) uint32_t sm[2][kPlaneWords];
    __shared__ __align__(16) uint32_t au[2][kN];
    const int n = plane_of_thread(), lt = lane_in_plane();
    const int m2 = rows * ell;
    const int poly = blockIdx.x, rm = poly % (rows * m2), r = rm / m2, m = rm % m2;
    uint32_t v[16];
    load_ntt_regs(v, gsw + ((size_t)poly * 2 + n) * kN, lt);
    ntt_inverse_plane(v, sm[n], lt, n);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) sm[n][nat_pos(lt, k)] = v[k];
    __syncthreads();
    const uint32_t bits_per = get_bits_per(ell);
    const int kk = m / rows;
    uint64_t gad = 0;                                              // buildGadget: G[r][r + kk*rows] = 2^(bits_per*kk)
    if ((m % rows) == r && (uint64_t)bits_per * kk < 64) gad = 1ull << (bits_per * kk);
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int z = threadIdx.x + 256 * k;
        const uint64_t val = crt_compose(sm[0][z], sm[1][z]);
        long long d = (long long)(z == 0 ? gad : 0) - (long long)val;
        if (d < 0) d += (long long)kQ;
        au[0][z] = raw_to_res((uint64_t)d, 0);
        au[1][z] = raw_to_res((uint64_t)d, 1);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) v[k] = au[n][nat_pos(lt, k)];
    ntt_forward_plane(v, sm[n], lt, n);
    store_ntt_regs(v, neg + ((size_t)poly * 2 + n) * kN, lt);
}
// poly_idx: 2*count entries (row-0 polynomials of this rank's bit ciphertexts, then their row-1 polynomials)
void launch_regev_to_gsw_sharded(const GswTargets &tg, const uint32_t *cv, const int *ct_idx, const int *poly_idx, const int *bit_ids, int count,
                                 int nu2, int t_gsw, const uint32_t *W, const uint32_t *V, int t_conv, uint64_t *scratch_raw, uint32_t *scratch_ntt, cudaStream_t s) {
    if (!count) return;
    launch_from_ntt_indexed(scratch_raw, cv, poly_idx, 2 * (size_t)count, s);
    launch_gadget_ntt(scratch_ntt, scratch_raw, t_conv, 1, count, s);
    launch_gadget_ntt(scratch_ntt + (size_t)t_conv * count * 2 * kN, scratch_raw + (size_t)count * kN, t_conv, 1, count, s);
    const size_t n = (size_t)count * 2 * kN;
    count_launch(); launch_pdl(k_regev_to_gsw_accum_sharded, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, tg, cv, ct_idx, bit_ids, scratch_ntt, W, V, t_conv, t_gsw, nu2, count);
}
void launch_gsw_negate(uint32_t *neg, const uint32_t *gsw, int count, int ell, int rows, cudaStream_t s) {
    const int polys = count * rows * rows * ell;
    if (polys) { count_launch(); launch_pdl(k_gsw_negate, dim3(polys), dim3(kNttThreads), 0, s, neg, gsw, ell, rows); }
}

void launch_scal_to_mat_reoriented(uint64_t *query_out, const uint32_t *cv, const int *ct_idx, const int *poly_idx, size_t dim0,
                                   const uint32_t *W, int t_conv, uint64_t *scratch_raw, uint32_t *scratch_ntt, cudaStream_t s) {
    launch_from_ntt_indexed(scratch_raw, cv, poly_idx, dim0, s);
    launch_gadget_ntt(scratch_ntt, scratch_raw, t_conv, 1, (int)dim0, s);
    const size_t n = dim0 * kN;
    count_launch();
    if (dim0 % 8 == 0 && t_conv == 4) {
        const int jtiles = dim0 % 32 == 0 ? 4 : 1;                    // j-tiles of 8 per CTA: the key registers are amortised over 32 j
        ScalTargets tg{};
        tg.query[0] = query_out; tg.ntargets = 1;
        launch_pdl(k_scal_to_mat_accum_tiled, dim3(kN / 32, (unsigned)(dim0 / 8 / jtiles)), dim3(256), 0, s, tg, cv, ct_idx, scratch_ntt, W, (int)dim0, (int)dim0, 0, 1, jtiles);
    }
    else launch_pdl(k_scal_to_mat_accum, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, query_out, cv, ct_idx, scratch_ntt, W, t_conv, (int)dim0);
}
// a rank's share of a sharded conversion: `count` (multiple of 8) ciphertexts, first-dimension index j = j_off + j_stride * jl,
// t_conv == 4; results are stored into every target of `tg` and this rank's flag is raised on each (see the kernel)
void launch_scal_to_mat_sharded(const ScalTargets &tg, const uint32_t *cv, const int *ct_idx, const int *poly_idx, size_t dim0, size_t count,
                                int j_off, int j_stride, const uint32_t *W, uint64_t *scratch_raw, uint32_t *scratch_ntt, cudaStream_t s) {
    launch_from_ntt_indexed(scratch_raw, cv, poly_idx, count, s);
    launch_gadget_ntt(scratch_ntt, scratch_raw, 4, 1, (int)count, s);
    const int jtiles = count % 32 == 0 ? 4 : 1;
    // (measured, timing only: writing each rank's rows as ONE contiguous block of the peers' buffers instead of every world-th 64-byte
    // row would save 23 us of cfg5's 100 us / 3 us of cfg1's 49 us at 2 GPUs - the all-gather is bound by the 16-32 MiB it moves)
    count_launch();
    launch_pdl(k_scal_to_mat_accum_tiled, dim3(kN / 32, (unsigned)(count / 8 / jtiles)), dim3(256), 0, s, tg, cv, ct_idx, scratch_ntt, W, (int)dim0, (int)count, j_off, j_stride, jtiles);
}
void launch_scal_to_mat_ntt(uint32_t *out, const uint32_t *cv, const int *ct_idx, const int *poly_idx, size_t count,
                            const uint32_t *W, int t_conv, uint64_t *scratch_raw, uint32_t *scratch_ntt, cudaStream_t s) {
    launch_from_ntt_indexed(scratch_raw, cv, poly_idx, count, s);
    launch_gadget_ntt(scratch_ntt, scratch_raw, t_conv, 1, (int)count, s);
    const size_t n = count * 2 * kN;
    count_launch(); launch_pdl(k_scal_to_mat_ntt, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, out, cv, ct_idx, scratch_ntt, W, t_conv, (int)count);
}
// poly_idx: 2*nbits entries - first nbits = row-0 polys of the bit ciphertexts, next nbits = row-1 polys
void launch_regev_to_gsw(uint32_t *gsw_out, uint32_t *gsw_neg_out, const uint32_t *cv, const int *ct_idx, const int *poly_idx,
                         int nu2, int t_gsw, const uint32_t *W, const uint32_t *V, int t_conv,
                         uint64_t *scratch_raw, uint32_t *scratch_ntt, cudaStream_t s) {
    const int nbits = nu2 * t_gsw;
    if (!nbits) return;
    launch_from_ntt_indexed(scratch_raw, cv, poly_idx, 2 * (size_t)nbits, s);
    // raw viewed as (rdim = 1) x (2*nbits) -> digits [k][2*nbits]; reinterpret as [row][k][nbits] needs row-major
    // split, so run the two rows separately: ginv[row][k][b]
    launch_gadget_ntt(scratch_ntt, scratch_raw, t_conv, 1, nbits, s);
    launch_gadget_ntt(scratch_ntt + (size_t)t_conv * nbits * 2 * kN, scratch_raw + (size_t)nbits * kN, t_conv, 1, nbits, s);
    const size_t n = (size_t)nbits * 2 * kN;
    count_launch(); launch_pdl(k_regev_to_gsw_accum, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, gsw_out, cv, ct_idx, scratch_ntt, W, V, t_conv, t_gsw, nu2);
    if (gsw_neg_out) launch_gsw_negate(gsw_neg_out, gsw_out, nu2, t_gsw, 3, s);
}

}  // namespace sb200
