// spiral_b200.cu - single translation unit of libspiral_b200.so (kernels share __constant__ tables).
#include "ntt_kernels.cu"
#include "spiral_kernels.cu"
#include "tc_scan.cu"
#include "query_kernels.cu"
#include "pack_kernels.cu"
#include "xchg_kernels.cu"
#include "api.cu"
#include "api_pack.cu"
