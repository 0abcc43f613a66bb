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
    unsigned int pad[31];
    // followed by slots[kXchgSlots][world][kCtWords] u64 (only rank 0's copy is used as the target)
};
__host__ __device__ inline size_t xchg_bytes(int world) { return sizeof(XchgBuf) + (size_t)kXchgSlots * world * kCtWords * 8; }
__device__ __forceinline__ uint64_t *xchg_slot(XchgBuf *b, int slot, int world, int r) {
    return reinterpret_cast<uint64_t *>(b + 1) + ((size_t)slot * world + r) * kCtWords;
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

// one CTA of 1024 threads; `target` = rank 0's XchgBuf (peer mapping, or local on rank 0), `mine` = this rank's own XchgBuf
__global__ void __launch_bounds__(1024) k_xchg_push(XchgBuf *target, XchgBuf *mine, const uint64_t *ct, unsigned int *epoch,
                                                    int rank, int world, unsigned int *error) {
    pdl_prologue();
    __shared__ int ok;
    const unsigned int e = *epoch + 1;
    const int slot = e % kXchgSlots;
    if (threadIdx.x == 0) ok = (e <= (unsigned)kXchgSlots) ? 1 : spin_until_ge(&mine->ack, e - kXchgSlots);
    __syncthreads();
    if (!ok) { if (threadIdx.x == 0) { *error = 1; *epoch = e; } return; }
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(ct);
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(xchg_slot(target, slot, world, rank));
    for (int i = threadIdx.x; i < (int)(kCtWords / 2); i += blockDim.x) dst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) { st_release_sys(&target->flags[slot][rank], e); *epoch = e; }
}

// rank 0 only: `mine` = rank 0's XchgBuf, `acks[r]` = pointer to rank r's XchgBuf::ack (peer mappings), out = world x ct
__global__ void __launch_bounds__(1024) k_xchg_wait(XchgBuf *mine, unsigned int *const *acks, uint64_t *out, const unsigned int *epoch,
                                                    int world, unsigned int *error) {
    pdl_prologue();
    __shared__ int ok;
    const unsigned int e = *epoch;                 // already advanced by this rank's own push (same stream)
    const int slot = e % kXchgSlots;
    if (threadIdx.x == 0) ok = 1;
    __syncthreads();
    if (threadIdx.x < world) { if (!spin_until_ge(&mine->flags[slot][threadIdx.x], e)) ok = 0; }
    __syncthreads();
    if (!ok) { if (threadIdx.x == 0) *error = 2; return; }
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(xchg_slot(mine, slot, world, 0));
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out);
    for (size_t i = threadIdx.x; i < (size_t)world * kCtWords / 2; i += blockDim.x) dst[i] = __ldcg(src + i);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) st_release_sys(acks[threadIdx.x], e);
}

void launch_xchg_push(void *target, void *mine, const uint64_t *ct, unsigned int *epoch, int rank, int world, unsigned int *error, cudaStream_t s) {
    count_launch();
    launch_pdl(k_xchg_push, dim3(1), dim3(1024), 0, s, (XchgBuf *)target, (XchgBuf *)mine, ct, epoch, rank, world, error);
}
void launch_xchg_wait(void *mine, unsigned int *const *acks, uint64_t *out, const unsigned int *epoch, int world, unsigned int *error, cudaStream_t s) {
    count_launch();
    launch_pdl(k_xchg_wait, dim3(1), dim3(1024), 0, s, (XchgBuf *)mine, acks, out, epoch, world, error);
}
size_t xchg_buffer_bytes(int world) { return xchg_bytes(world); }
size_t xchg_ack_offset() { return offsetof(XchgBuf, ack); }

}  // namespace sb200
