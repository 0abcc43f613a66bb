// ref_abi.h - ABI-compatible declarations of the reference's argument types, written for the host
// mirror (host_mirror.cpp).  Only the memory layout matters: the reference passes these by reference
// or by value across its own translation units, and the dynamic linker binds calls by mangled name.
//   MatPoly            include/poly.h:24-64        {size_t rows, cols; uint64_t *data; bool isNTT}
//   FurtherDimsLocals  include/spiral.h:86-127     six buffers + two sizes, passed BY VALUE
//   ExpansionLocals    include/spiral.h:129-198    five buffers + two sizes, passed BY VALUE
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

struct MatPoly {
    size_t rows;
    size_t cols;
    uint64_t *data;
    bool isNTT;
    MatPoly() : rows(0), cols(0), data(nullptr), isNTT(true) {}
    MatPoly(size_t r, size_t c, bool ntt = true) : rows(r), cols(c), isNTT(ntt) {
        data = (uint64_t *)calloc(r * c * (ntt ? 2 : 1) * 2048, sizeof(uint64_t));
    }
    size_t words() const { return rows * cols * (isNTT ? 2 : 1) * 2048; }
};
static_assert(sizeof(MatPoly) == 32, "MatPoly layout must match the reference");

class FurtherDimsLocals {
public:
    uint64_t *result, *cts, *scratch_cts1, *scratch_cts2, *scratch_cts_double1, *scratch_cts_double2;
    size_t num_per, num_bytes_C;
};
class ExpansionLocals {
public:
    uint64_t *cts, *scratch_cts1, *scratch_cts2, *small_coeff_polys, *reoriented_ciphertexts;
    size_t n1_padded, split;
};
