// xchg_kernels.cu - the one exchange step of the sharded path, over NVLink / NVSwitch peer memory.
//
// After the local fold rounds every GPU holds ONE surviving ciphertext (3x2 raw, 96 KiB).  Instead of a host-driven
// NCCL gather, each rank's k_xchg_push stores its ciphertext straight into rank 0's HBM through a peer mapping
// (cudaIpc), then publishes a system-scope release flag; rank 0's k_xchg_wait acquires the flags, pulls the
// ciphertexts into its private fold buffer and acknowledges, and the tail folds follow in the same stream / graph.
// No host round trip, no collective launch: the exchange costs one small kernel on each side.
//
// Protocol (all counters are per-query epochs, every rank processes the same query sequence):
//   slot = epoch % kXchgSlots ;  push waits until ack >= epoch - kXchgSlots (slot free), writes, sets flag[slot][rank] = epoch
//   wait  spins until flag[slot][r] == epoch for all r, copies, then stores ack = epoch on every rank (peer writes)
// Every spin is bounded (kXchgTimeoutNs): on timeout the kernel raises an error word instead of hanging the GPU.
#include "kernels.cuh"
#include "common.cuh"

namespace sb200 {

constexpr int kXchgSlots = 2;
constexpr int kXchgMaxWorld = 16;
constexpr unsigned long long kXchgTimeoutNs = 4000000000ull;     // 4 s
constexpr size_t kCtWords = 6 * (size_t)kN;

struct XchgBuf {                 // lives in ONE cudaMalloc allocation per rank (exported through cudaIpc)
    unsigned int flags[kXchgSlots][kXchgMaxWorld];
    unsigned int ack;            // last epoch rank 0 has consumed (written remotely by rank 0)
    unsigned int arrive[2];      // CTA arrival counters of this rank's multi-CTA push / wait kernels
    unsigned int qflags[kXchgMaxWorld];   // query all-gather (sharded expansion / Pack direct upload): epoch of the slice rank r has stored here
    unsigned int gflags[kXchgMaxWorld];   // GSW all-gather (sharded RegevToGSW): epoch of the columns rank r has stored here
    unsigned int arrive_aux;              // CTA arrival counter of kernels on the side stream
    unsigned int pad[12];
    // followed by slots[kXchgSlots][world][slot_words] u64 (only rank 0's copy is used as the target)
};
static_assert(sizeof(XchgBuf) % 16 == 0, "slots must stay 16-byte aligned");
__host__ __device__ inline size_t xchg_bytes_w(int world, size_t slot_words) { return sizeof(XchgBuf) + (size_t)kXchgSlots * world * slot_words * 8; }
__host__ __device__ inline size_t xchg_bytes(int world) { return xchg_bytes_w(world, kCtWords); }
__device__ __forceinline__ uint64_t *xchg_slot(XchgBuf *b, int slot, int world, int r, size_t slot_words) {
    return reinterpret_cast<uint64_t *>(b + 1) + ((size_t)slot * world + r) * slot_words;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int *p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// returns false on timeout
__device__ __forceinline__ bool spin_until_ge(const unsigned int *p, unsigned int target) {
    const unsigned long long t0 = now_ns();
    while ((int)(ld_acquire_sys(p) - target) < 0) {
        __nanosleep(200);
        if (now_ns() - t0 > kXchgTimeoutNs) return false;
    }
    return true;
}

// gridDim.x CTAs of 1024 threads; `target` = rank 0's XchgBuf (peer mapping, or local on rank 0), `mine` = this rank's own XchgBuf.
// Every CTA stores its share of the `words`-word payload into slot [epoch % 2][rank]; the last CTA to finish publishes the flag.
__global__ void __launch_bounds__(1024) k_xchg_push(XchgBuf *target, XchgBuf *mine, const uint64_t *ct, unsigned int *epoch,
                                                    int rank, int world, unsigned int *error, size_t words, size_t slot_words) {
    pdl_prologue_no_early_dependents();
    __shared__ int ok;
    const unsigned int e = *epoch + 1;
    const int slot = e % kXchgSlots;
    if (threadIdx.x == 0) ok = (e <= (unsigned)kXchgSlots) ? 1 : spin_until_ge(&mine->ack, e - kXchgSlots);
    __syncthreads();
    if (ok) {
        const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(ct);
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(xchg_slot(target, slot, world, rank, slot_words));
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words / 2; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!ok) *error = 1;
        __threadfence();
        if (atomicAdd(&mine->arrive[0], 1u) == gridDim.x - 1) {          // every CTA's stores are fenced before its arrival
            mine->arrive[0] = 0;
            __threadfence_system();
            if (ld_acquire_sys(error) == 0) st_release_sys(&target->flags[slot][rank], e);
            *epoch = e;
        }
    }
}

// rank 0 only: `mine` = rank 0's XchgBuf, `acks[r]` = pointer to rank r's XchgBuf::ack (peer mappings), out = world x slot_words.
// gridDim.x CTAs: each waits for all flags, copies its share of the slots into the private buffer; the last one acknowledges.
__global__ void __launch_bounds__(1024) k_xchg_wait(XchgBuf *mine, unsigned int *const *acks, uint64_t *out, const unsigned int *epoch,
                                                    int world, unsigned int *error, size_t slot_words, int ack_now) {
    pdl_prologue_no_early_dependents();
    __shared__ int ok;
    const unsigned int e = *epoch;                 // already advanced by this rank's own push (same stream)
    const int slot = e % kXchgSlots;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < world) { if (!spin_until_ge(&mine->flags[slot][threadIdx.x], e)) ok = 0; }
    __syncthreads();
    if (ok) {
        const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(xchg_slot(mine, slot, world, 0, slot_words));
        ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out);
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)world * slot_words / 2; i += (size_t)gridDim.x * blockDim.x) dst[i] = __ldcg(src + i);
        __threadfence_system();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (!ok) *error = 2;
        __threadfence();
        if (atomicAdd(&mine->arrive[1], 1u) == gridDim.x - 1) {
            mine->arrive[1] = 0;
            __threadfence_system();
            if (ack_now && ld_acquire_sys(error) == 0) for (int r = 0; r < world; r++) st_release_sys(acks[r], e);
        }
    }
}
// rank 0, at the very END of a query whose expansion was sharded: the acknowledgement that lets the other ranks overwrite this
// rank's query and GSW buffers (with it sent right after the gather, rank 0's tail folds could still be reading the GSW ciphertexts)
__global__ void k_xchg_ack(unsigned int *const *acks, const unsigned int *epoch, int world, const unsigned int *error) {
    pdl_prologue_no_early_dependents();
    if ((int)threadIdx.x < world && ld_acquire_sys(error) == 0) { __threadfence_system(); st_release_sys(acks[threadIdx.x], *epoch); }
}
// waits until all `world` flags of this rank's header (qflags: which = 0, gflags: which = 1) have reached the current query
__global__ void k_flag_wait(XchgBuf *mine, int world, const unsigned int *epoch, unsigned int *error, int which, int advanced) {
    pdl_prologue_no_early_dependents();
    const unsigned int e = *epoch + (advanced ? 0u : 1u);          // advanced: this rank's push of the query has already run
    const unsigned int *f = which ? mine->gflags : mine->qflags;
    if ((int)threadIdx.x < world && !spin_until_ge(&f[threadIdx.x], e)) *error = 3 + which;
}

static inline unsigned xchg_grid(size_t words) { size_t g = (words * 8 + 131071) / 131072; return (unsigned)(g < 1 ? 1 : g > 32 ? 32 : g); }   // ~128 KiB per CTA
void launch_xchg_push_w(void *target, void *mine, const uint64_t *ct, unsigned int *epoch, int rank, int world, unsigned int *error,
                        size_t words, size_t slot_words, cudaStream_t s) {
    count_launch();
    launch_pdl(k_xchg_push, dim3(xchg_grid(words)), dim3(1024), 0, s, (XchgBuf *)target, (XchgBuf *)mine, ct, epoch, rank, world, error, words, slot_words);
}
void launch_xchg_wait_w(void *mine, unsigned int *const *acks, uint64_t *out, const unsigned int *epoch, int world, unsigned int *error,
                        size_t slot_words, cudaStream_t s, int ack_now = 1) {
    count_launch();
    launch_pdl(k_xchg_wait, dim3(xchg_grid((size_t)world * slot_words)), dim3(1024), 0, s, (XchgBuf *)mine, acks, out, epoch, world, error, slot_words, ack_now);
}
void launch_xchg_ack(unsigned int *const *acks, const unsigned int *epoch, int world, const unsigned int *error, cudaStream_t s) {
    count_launch(); launch_pdl(k_xchg_ack, dim3(1), dim3(32), 0, s, acks, epoch, world, error);
}
void launch_flag_wait(void *mine, int world, const unsigned int *epoch, unsigned int *error, int which, int advanced, cudaStream_t s) {
    count_launch(); launch_pdl(k_flag_wait, dim3(1), dim3(32), 0, s, (XchgBuf *)mine, world, epoch, error, which, advanced);
}
void launch_xchg_push(void *target, void *mine, const uint64_t *ct, unsigned int *epoch, int rank, int world, unsigned int *error, cudaStream_t s) {
    launch_xchg_push_w(target, mine, ct, epoch, rank, world, error, kCtWords, kCtWords, s);
}
void launch_xchg_wait(void *mine, unsigned int *const *acks, uint64_t *out, const unsigned int *epoch, int world, unsigned int *error, cudaStream_t s, int ack_now = 1) {
    launch_xchg_wait_w(mine, acks, out, epoch, world, error, kCtWords, s, ack_now);
}
size_t xchg_buffer_bytes(int world) { return xchg_bytes(world); }
size_t xchg_buffer_bytes_w(int world, size_t slot_words) { return xchg_bytes_w(world, slot_words); }
size_t xchg_ack_offset() { return offsetof(XchgBuf, ack); }
size_t xchg_header_bytes() { return sizeof(XchgBuf); }
unsigned int *xchg_qflag_ptr(void *buf, int rank) { return &reinterpret_cast<XchgBuf *>(buf)->qflags[rank]; }   // address arithmetic only
unsigned int *xchg_arrive_ptr(void *buf, int k) { return &reinterpret_cast<XchgBuf *>(buf)->arrive[k]; }
unsigned int *xchg_ack_ptr(void *buf) { return &reinterpret_cast<XchgBuf *>(buf)->ack; }
unsigned int *xchg_gflag_ptr(void *buf, int rank) { return &reinterpret_cast<XchgBuf *>(buf)->gflags[rank]; }
unsigned int *xchg_arrive_aux_ptr(void *buf) { return &reinterpret_cast<XchgBuf *>(buf)->arrive_aux; }

// ---- query all-gather over peer memory (Pack direct upload, sharded): every rank uploads and reorients only its 1/world
// slice of the first-dimension ciphertexts and stores it into EVERY rank's query buffer, then raises qflags[rank] there;
// k_query_wait (first kernel of the scan's stream order) waits until all world slices of this epoch have landed.
struct QueryPeers { uint64_t *query[kXchgMaxWorld]; XchgBuf *xb[kXchgMaxWorld]; };
// ct_idx (nullable): position in cv of this rank's k-th ciphertext (a sharded EXPANSION keeps the leaves j = rank + world * k of the
// tree at cv[2 j]; an uploaded slice is simply cv[k]); its row in the query is j_begin + k * j_stride
__global__ void __launch_bounds__(256) k_reorient_dim1_allgather(const __grid_constant__ QueryPeers peers, const uint32_t *__restrict__ cv,
                                                                 const int *__restrict__ ct_idx, int dim0, int j_begin, int j_stride, int j_count,
                                                                 int rank, int world, const unsigned int *epoch, XchgBuf *mine) {
    pdl_prologue_no_early_dependents();
    // this query is number e of the exchange sequence; its slices may only overwrite the peers' query buffers once rank 0 has
    // gathered query e - 1 from EVERY rank (ack >= e - 1: all scans of the previous query are over)
    __shared__ int ok;
    const unsigned int qepoch = *epoch + 1;
    if (threadIdx.x == 0) ok = qepoch <= 1 ? 1 : spin_until_ge(&mine->ack, qepoch - 1);
    __syncthreads();
    if (!ok) return;                                   // the missing flag makes every rank's k_query_wait report the time-out
    // thread = (z, j-pair): 32 bytes of two consecutive j per store, to every peer
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // (z, jl/2), j fastest: contiguous runs per z
    const int half = j_count / 2;
    if (idx < (size_t)kN * half) {
        const int z = (int)(idx / half), jl = (int)(idx % half) * 2;
        ulonglong2 w[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const uint32_t *ct = cv + (size_t)(ct_idx ? ct_idx[jl + t] : jl + t) * 2 * 2 * kN;
            w[t].x = (uint64_t)ct[z] | ((uint64_t)ct[kN + z] << 32);
            w[t].y = (uint64_t)ct[2 * kN + z] | ((uint64_t)ct[3 * kN + z] << 32);
        }
        const size_t o = (size_t)z * dim0 + j_begin + (size_t)jl * j_stride;
        for (int r = 0; r < world; r++) {
            ulonglong2 *q = reinterpret_cast<ulonglong2 *>(peers.query[(rank + r) % world]);      // start with the own copy, spread the peers
            q[o] = w[0]; q[o + j_stride] = w[1];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&mine->arrive[0], 1u) == gridDim.x - 1) {
            mine->arrive[0] = 0;
            __threadfence_system();
            for (int r = 0; r < world; r++) st_release_sys(&peers.xb[r]->qflags[rank], qepoch);
        }
    }
}
__global__ void k_query_wait(XchgBuf *mine, int world, const unsigned int *epoch, unsigned int *error) {
    pdl_prologue_no_early_dependents();
    if ((int)threadIdx.x < world && !spin_until_ge(&mine->qflags[threadIdx.x], *epoch + 1)) *error = 3;
}
void launch_reorient_dim1_allgather(const QueryPeers &peers, const uint32_t *cv, const int *ct_idx, size_t dim0, size_t j_begin, size_t j_stride, size_t j_count,
                                    int rank, int world, const unsigned int *qepoch, void *mine, cudaStream_t s) {
    const size_t n = (size_t)kN * (j_count / 2);
    count_launch();
    launch_pdl(k_reorient_dim1_allgather, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, s, peers, cv, ct_idx, (int)dim0, (int)j_begin, (int)j_stride, (int)j_count,
               rank, world, qepoch, (XchgBuf *)mine);
}
void launch_query_wait(void *mine, int world, const unsigned int *qepoch, unsigned int *error, cudaStream_t s) {
    count_launch();
    launch_pdl(k_query_wait, dim3(1), dim3(32), 0, s, (XchgBuf *)mine, world, qepoch, error);
}

}  // namespace sb200
