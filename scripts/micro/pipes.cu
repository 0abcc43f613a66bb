// micro-benchmarks behind the scan's roofline discussion: IMAD.WIDE.U32 issue rate and the read-only
// HBM bandwidth of a B200.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu ; ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int CH> __global__ void __launch_bounds__(128) k_wide(uint64_t *out, uint32_t a0, uint32_t b0, int iters) {
    uint64_t acc[CH]; uint32_t a[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { acc[i] = i; a[i] = a0 + i * 7 + threadIdx.x; }
    uint32_t b = b0 + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) acc[i] += (uint64_t)a[i] * b;
        b += 3;
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= acc[i];
    if (s == 0x1234567) out[0] = s;
}
template <int CH> __global__ void __launch_bounds__(128) k_lo(uint32_t *out, uint32_t a0, uint32_t b0, int iters) {
    uint32_t acc[CH], a[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { acc[i] = i; a[i] = a0 + i * 7 + threadIdx.x; }
    uint32_t b = b0 + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) acc[i] += a[i] * b;
        b += 3;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s ^= acc[i];
    if (s == 0x1234567) out[0] = s;
}
template <int CH> __global__ void __launch_bounds__(128) k_dfma(double *out, double a0, double b0, int iters) {
    double acc[CH], a[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { acc[i] = i; a[i] = a0 + i * 7 + threadIdx.x; }
    double b = b0 + blockIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) acc[i] = fma(a[i], b, acc[i]);
        b += 3.0;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += acc[i];
    if (s == 0.1234567) out[0] = s;
}
__device__ __forceinline__ uint4 ld_na(const uint4 *p) {
    uint4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r;
}
template <int UNR> __global__ void __launch_bounds__(128) k_read(uint32_t *out, const uint4 *in, size_t per_cta) {   // per_cta uint4, contiguous per CTA
    const uint4 *p = in + (size_t)blockIdx.x * per_cta + threadIdx.x;
    uint32_t s = 0;
    for (size_t o = 0; o < per_cta; o += 128 * UNR) {
        uint4 v[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) v[u] = ld_na(p + o + u * 128);
#pragma unroll
        for (int u = 0; u < UNR; u++) s ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (s == 0x1234567) out[0] = s;
}
template <typename F> float timed(F f, int reps) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int r = 0; r < reps; r++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / reps;
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int sms = pr.multiProcessorCount; const double ghz = clk_khz / 1e6;
    uint64_t *o64; uint32_t *o32; cudaMalloc(&o64, 64); cudaMalloc(&o32, 64);
    const int iters = 8192;
    for (int ctas = 1; ctas <= 8; ctas *= 2) {
        float ms = timed([&] { k_wide<16><<<sms * ctas, 128>>>(o64, 3, 5, iters); }, 5);
        double ops = (double)sms * ctas * 128 * 16.0 * iters;
        printf("IMAD.WIDE.U32 (+64-bit add), 16 chains, %d CTAs x 128 thr / SM: %.1f per SM per clk (at %.3f GHz nominal)\n", ctas, ops / (ms * 1e-3) / sms / (ghz * 1e9), ghz);
        ms = timed([&] { k_dfma<16><<<sms * ctas, 128>>>((double *)o64, 3.0, 5.0, iters); }, 5);
        printf("DFMA (fp64), 16 chains, %d CTAs x 128 thr / SM: %.1f per SM per clk\n", ctas, ops / (ms * 1e-3) / sms / (ghz * 1e9));
        ms = timed([&] { k_lo<16><<<sms * ctas, 128>>>(o32, 3, 5, iters); }, 5);
        printf("IMAD (32-bit), 16 chains, %d CTAs x 128 thr / SM: %.1f per SM per clk\n", ctas, ops / (ms * 1e-3) / sms / (ghz * 1e9));
    }
    const size_t bytes = (size_t)16 << 30;
    uint4 *buf; if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 1, bytes);
    for (int per_sm = 4; per_sm <= 16; per_sm *= 2) {
        const int grid = sms * per_sm * 4;                      // several waves
        const size_t per_cta = (bytes / 16 / grid) / (128 * 8) * (128 * 8);
        float ms4 = timed([&] { k_read<4><<<grid, 128>>>(o32, buf, per_cta); }, 3);
        float ms8 = timed([&] { k_read<8><<<grid, 128>>>(o32, buf, per_cta); }, 3);
        const double b = (double)per_cta * 16 * grid;
        printf("read-only stream, %d CTAs (contiguous %.1f MiB each): unroll 4 %.0f GB/s, unroll 8 %.0f GB/s\n", grid, per_cta * 16 / 1048576.0, b / (ms4 * 1e-3) / 1e9, b / (ms8 * 1e-3) / 1e9);
    }
    uint4 *dst; if (cudaMalloc(&dst, bytes / 2) == cudaSuccess) {
        float ms = timed([&] { cudaMemcpyAsync(dst, buf, bytes / 2, cudaMemcpyDeviceToDevice); }, 3);
        printf("cudaMemcpy D2D %.0f GB/s (read + write bytes)\n", 2.0 * (bytes / 2) / (ms * 1e-3) / 1e9);
    }
    return 0;
}
