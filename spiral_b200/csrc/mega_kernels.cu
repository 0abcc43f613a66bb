// mega_kernels.cu - the whole coefficient expansion (all rounds) as ONE persistent dataflow kernel.
//
// expandImproved (reference src/spiral.cpp:1664-1743) is g dependent rounds of three dependent phases; as
// separate launches that is 3g short kernels whose launch + ramp-up + drain latencies dominate the early
// rounds (2..32 ciphertexts).  Here every phase instance is a WORK ITEM in a host-built table
//     PREP(round, slot, row)  ->  DIGIT(round, slot, k)  ->  ACCUM(round, slot, row, quarter)
// ordered round-major.  Persistent CTAs pull items with one atomicAdd (so an item's producers are always
// already resident: no deadlock), spin on per-(round, slot) completion counters with acquire loads, run the
// same device code as the per-phase kernels, and publish with fence + atomicAdd.  Independent slots and the
// tail of one round / head of the next overlap freely; the critical path is the chain of CTA latencies.
// Every round has its own scratch region (c0, c1, digits), so there are no write-after-read hazards.
// Data produced inside the kernel is read with ld.global.cg (L2), never through the non-coherent path.
#include "kernels.cuh"
#include "ntt.cuh"
#include <vector>

namespace sb200 {

enum { MEGA_PREP = 0, MEGA_DIGIT = 1, MEGA_ACCUM = 2 };
struct MegaItem { int type, round, slot, part; };     // part: row (PREP), k (DIGIT), row*4 + quarter (ACCUM)
struct MegaRound {
    int num_in, cnt, tstride, list_off;                // tstride: digit polys per slot in this round's region
    uint32_t tpow;
    size_t c_off, g_off;                                // slot base of this round in c0/c1 (slots) and ginv (polys)
};
constexpr int kMegaMaxRounds = 16;
struct MegaArgs {
    uint32_t *cv;
    const uint32_t *W_left, *W_right, *neg1;
    const uint16_t *perms;
    uint64_t *c0;
    uint32_t *c1, *ginv;
    const int *active;              // concatenated active lists
    const uint8_t *store_partner;   // per (round, slot): 1 if PREP of i < num_in must store cv[i + num_in]
    const int *partner_slot;        // per (round, slot): slot of the ACTIVE partner i + num_in (reads cv[i] in its PREP), or -1
    const MegaItem *items;
    int n_items, t_left, t_right, n_cts;
    int *next;                      // work counter
    int *prep_done, *digit_done;    // [round][slot]  (slot index inside the round, stride n_cts)
    int *accum_done;                // [round][ct]
    MegaRound rounds[kMegaMaxRounds];
};

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void wait_counter(const int *p, int target) {     // called by thread 0 only
    while (ld_acquire(p) < target) __nanosleep(64);
}
__device__ __forceinline__ void load_ntt_regs_cg(uint32_t (&v)[16], const uint32_t *plane_ptr, int lt) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const uint4 *src = reinterpret_cast<const uint4 *>(plane_ptr + h * 1024 + 8 * lt);
        uint4 a = __ldcg(src), b = __ldcg(src + 1);
        uint32_t *u = &v[8 * h];
        u[0] = a.x; u[1] = a.y; u[2] = a.z; u[3] = a.w; u[4] = b.x; u[5] = b.y; u[6] = b.z; u[7] = b.w;
    }
}

struct MegaSmem {
    union {
        uint32_t planes[2][kPlaneWords];
        ulonglong2 part[3][64][2];
    };
    int item;
};

__device__ __forceinline__ void mega_prep(const MegaArgs &a, const MegaRound &R, int round, int slot, int row, MegaSmem &sm) {
    const int n = plane_of_thread(), lt = lane_in_plane();
    const int i = a.active[R.list_off + slot];
    const int src = i < R.num_in ? i : i - R.num_in;
    if (round > 0 && threadIdx.x == 0) wait_counter(a.accum_done + (size_t)(round - 1) * a.n_cts + src, 8);   // 2 rows x 4 quarters
    __syncthreads();
    uint32_t v[16], w[16];
    load_ntt_regs_cg(v, a.cv + (((size_t)src * 2 + row) * 2 + n) * kN, lt);
    load_ntt_regs(w, a.neg1 + ((size_t)round * 2 + n) * kN, lt);
    if (i < R.num_in) {
        if (a.store_partner[(size_t)round * a.n_cts + slot]) {      // partner 2^r + i is skipped this round: store it here (src/spiral.cpp:1709)
            uint32_t nb[16];
#pragma unroll
            for (int e = 0; e < 16; e++) nb[e] = mulmod(v[e], w[e], n);
            store_ntt_regs(nb, a.cv + (((size_t)(i + R.num_in) * 2 + row) * 2 + n) * kN, lt);
        }
    } else {
#pragma unroll
        for (int e = 0; e < 16; e++) v[e] = mulmod(v[e], w[e], n);
        store_ntt_regs(v, a.cv + (((size_t)i * 2 + row) * 2 + n) * kN, lt);   // the active partner stores its own copy (before its ACCUM)
    }
    const size_t cslot = R.c_off + slot;
    if (row == 1) {
        const uint16_t *perm = a.perms + (size_t)round * kN;
        plane_sync(n);
#pragma unroll
        for (int k = 0; k < 16; k++) sm.planes[n][ntt_pos(lt, k)] = v[k];
        plane_sync(n);
        uint32_t *dst = a.c1 + (cslot * 2 + n) * kN;
        for (int pos = lt; pos < kN; pos += kPlaneThreads) dst[pos] = sm.planes[n][perm[pos]];
        return;
    }
    ntt_inverse_plane(v, sm.planes[n], lt, n);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; k++) sm.planes[n][nat_pos(lt, k)] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const int z = threadIdx.x + 256 * k;
        uint64_t val = crt_compose(sm.planes[0][z], sm.planes[1][z]);
        const uint32_t it = (uint32_t)z * R.tpow;
        if ((it >> kLogN) & 1) val = kQ - val;                 // 0 -> Q on purpose (reference src/poly.cpp:256)
        a.c0[cslot * kN + (it & (kN - 1))] = val;
    }
}

__device__ __forceinline__ void mega_digit(const MegaArgs &a, const MegaRound &R, int round, int slot, int k, MegaSmem &sm) {
    const int n = plane_of_thread(), lt = lane_in_plane();
    const int i = a.active[R.list_off + slot];
    const int gd = (i & 1) ? a.t_right : a.t_left;
    if (threadIdx.x == 0) wait_counter(a.prep_done + (size_t)round * a.n_cts + slot, 2);
    __syncthreads();
    const uint32_t bits_per = get_bits_per(gd);
    const uint64_t mask = (1ull << bits_per) - 1;
    const uint64_t *src = a.c0 + (R.c_off + slot) * kN;
    uint32_t v[16];
#pragma unroll
    for (int e = 0; e < 16; e++) {
        const uint64_t d = gadget_digit(__ldcg(src + nat_pos(lt, e)), k, bits_per, mask);
        v[e] = bits_per >= 28 ? raw_to_res(d, n) : (uint32_t)d;
    }
    ntt_forward_plane(v, sm.planes[n], lt, n);
    store_ntt_regs(v, a.ginv + ((R.g_off + (size_t)slot * R.tstride + k) * 2 + n) * kN, lt);
}

__device__ __forceinline__ void mega_accum(const MegaArgs &a, const MegaRound &R, int round, int slot, int part, MegaSmem &sm) {
    const int row = part >> 2, quarter = part & 3;
    const int i = a.active[R.list_off + slot];
    const int gd = (i & 1) ? a.t_right : a.t_left;
    if (threadIdx.x == 0) {
        wait_counter(a.digit_done + (size_t)round * a.n_cts + slot, gd);
        wait_counter(a.prep_done + (size_t)round * a.n_cts + slot, 2);
        // this item overwrites cv[i]; the partner's PREP (i + 2^r) still reads the OLD cv[i]: wait for it
        const int ps = a.partner_slot[(size_t)round * a.n_cts + slot];
        if (ps >= 0) wait_counter(a.prep_done + (size_t)round * a.n_cts + ps, 2);
    }
    __syncthreads();
    const int col = threadIdx.x & 63, grp = threadIdx.x >> 6;
    const uint32_t *Wbase = ((i & 1) ? a.W_right + (size_t)round * 2 * a.t_right * 2 * kN : a.W_left + (size_t)round * 2 * a.t_left * 2 * kN) + (size_t)row * gd * 2 * kN;
    const uint32_t *Gbase = a.ginv + (R.g_off + (size_t)slot * R.tstride) * 2 * kN;
    for (int s4 = 0; s4 < 4; s4++) {
        const int w4 = (quarter * 4 + s4) * 64 + col, n = w4 >= 512;
        const uint4 *W = reinterpret_cast<const uint4 *>(Wbase) + w4;
        const uint4 *G = reinterpret_cast<const uint4 *>(Gbase) + w4;
        uint64_t acc[4] = {0, 0, 0, 0};
#pragma unroll 4
        for (int k = grp; k < gd; k += 4) {
            const uint4 x = __ldg(W + (size_t)k * (2 * kN / 4)), y = __ldcg(G + (size_t)k * (2 * kN / 4));
            acc[0] += (uint64_t)x.x * y.x; acc[1] += (uint64_t)x.y * y.y;
            acc[2] += (uint64_t)x.z * y.z; acc[3] += (uint64_t)x.w * y.w;
        }
        if (grp > 0) {
            sm.part[grp - 1][col][0] = make_ulonglong2(acc[0], acc[1]);
            sm.part[grp - 1][col][1] = make_ulonglong2(acc[2], acc[3]);
        }
        __syncthreads();
        if (grp == 0) {
#pragma unroll
            for (int g = 0; g < 3; g++) {
                const ulonglong2 p0 = sm.part[g][col][0], p1 = sm.part[g][col][1];
                acc[0] += p0.x; acc[1] += p0.y; acc[2] += p1.x; acc[3] += p1.y;
            }
            uint4 *dst = reinterpret_cast<uint4 *>(a.cv + ((size_t)i * 2 + row) * 2 * kN) + w4;
            const uint4 cur = __ldcg(dst);
            uint4 add = make_uint4(0, 0, 0, 0);
            if (row == 1) add = __ldcg(reinterpret_cast<const uint4 *>(a.c1 + (R.c_off + slot) * 2 * kN) + w4);
            uint4 o;
            o.x = reduce_u64(acc[0] + cur.x + add.x, n); o.y = reduce_u64(acc[1] + cur.y + add.y, n);
            o.z = reduce_u64(acc[2] + cur.z + add.z, n); o.w = reduce_u64(acc[3] + cur.w + add.w, n);
            *dst = o;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kNttThreads) k_expand_mega(const __grid_constant__ MegaArgs a) {
    pdl_prologue();
    __shared__ __align__(16) MegaSmem sm;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sm.item = atomicAdd(a.next, 1);
        __syncthreads();
        const int id = sm.item;
        if (id >= a.n_items) return;
        const MegaItem it = a.items[id];
        const MegaRound &R = a.rounds[it.round];
        int *done;
        if (it.type == MEGA_PREP) { mega_prep(a, R, it.round, it.slot, it.part, sm); done = a.prep_done + (size_t)it.round * a.n_cts + it.slot; }
        else if (it.type == MEGA_DIGIT) { mega_digit(a, R, it.round, it.slot, it.part, sm); done = a.digit_done + (size_t)it.round * a.n_cts + it.slot; }
        else { mega_accum(a, R, it.round, it.slot, it.part, sm); done = a.accum_done + (size_t)it.round * a.n_cts + a.active[R.list_off + it.slot]; }
        __syncthreads();                 // every thread's global writes happen-before thread 0's fence
        if (threadIdx.x == 0) { __threadfence(); atomicAdd(done, 1); }
    }
}

// ---- host side -------------------------------------------------------------------------------------
struct MegaPlan {
    std::vector<MegaItem> items;
    std::vector<uint8_t> store_partner;      // [round][slot] with stride n_cts
    std::vector<int> partner_slot;
    MegaRound rounds[kMegaMaxRounds];
    size_t c_slots = 0, g_polys = 0;         // totals for the per-round scratch regions
    int n_cts = 0, g = 0;
};
void build_mega_plan(MegaPlan &mp, const ExpandPlan &p, const int *list, const int *offs, const int *cnt, int r_begin, int r_end) {
    mp.g = p.g; mp.n_cts = 1 << p.g;
    mp.items.clear();
    mp.store_partner.assign((size_t)p.g * mp.n_cts, 0);
    mp.partner_slot.assign((size_t)p.g * mp.n_cts, -1);
    mp.c_slots = 0; mp.g_polys = 0;
    const int tmax = p.t_left > p.t_right ? p.t_left : p.t_right;
    for (int r = 0; r < p.g; r++) {
        MegaRound &R = mp.rounds[r];
        const bool any_odd = !(p.stopround > 0 && r > p.stopround);
        R.num_in = 1 << r; R.cnt = cnt[r]; R.tstride = any_odd ? tmax : p.t_left; R.list_off = offs[r];
        R.tpow = (uint32_t)(kN >> r) + 1;
        R.c_off = mp.c_slots; R.g_off = mp.g_polys;
        mp.c_slots += cnt[r]; mp.g_polys += (size_t)cnt[r] * R.tstride;
        if (r < r_begin || r >= r_end) continue;
        std::vector<int> slot_of(2 << r, -1);
        for (int s = 0; s < cnt[r]; s++) slot_of[list[offs[r] + s]] = s;
        for (int s = 0; s < cnt[r]; s++) {
            const int i = list[offs[r] + s];
            if (i >= R.num_in) continue;
            if (slot_of[i + R.num_in] < 0) mp.store_partner[(size_t)r * mp.n_cts + s] = 1;
            else mp.partner_slot[(size_t)r * mp.n_cts + s] = slot_of[i + R.num_in];
        }
        for (int s = 0; s < cnt[r]; s++) for (int row = 0; row < 2; row++) mp.items.push_back({MEGA_PREP, r, s, row});
        for (int s = 0; s < cnt[r]; s++) {
            const int gd = (list[offs[r] + s] & 1) ? p.t_right : p.t_left;
            for (int k = 0; k < gd; k++) mp.items.push_back({MEGA_DIGIT, r, s, k});
        }
        for (int s = 0; s < cnt[r]; s++) for (int part = 0; part < 8; part++) mp.items.push_back({MEGA_ACCUM, r, s, part});
    }
}

void launch_expand_mega(const MegaArgs &args, int n_sm, cudaStream_t s) {
    // counters (work index + completion counts) are zeroed by the caller on the same stream
    count_launch();
    launch_pdl(k_expand_mega, dim3((unsigned)(n_sm * 4)), dim3(kNttThreads), 0, s, args);
}

}  // namespace sb200
