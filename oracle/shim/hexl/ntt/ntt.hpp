// Decode-only stand-in for Intel HEXL's intel::hexl::NTT (HEXL 1.2.1 is the
// reference's vcpkg dependency; it is not installed in this image and there is
// no network).  TEST INFRASTRUCTURE: used only to compile the unmodified
// reference into oracle/_ref/.  The reference touches HEXL at exactly three
// call sites, all on the CLIENT decode path (src/spiral.cpp:1237,
// src/util.cpp:231,241): forward NTT -> pointwise product -> inverse NTT over
// arb_qprime.  That composition does not depend on the NTT's output ordering or
// root choice, so any correct negacyclic transform reproduces HEXL's results
// exactly.  This one is a plain O(n log n) Cooley-Tukey/Gentleman-Sande pair.
#pragma once
#include <cstdint>
#include <cstddef>
#include <vector>

namespace intel { namespace hexl {

class NTT {
    uint64_t n_, q_;
    std::vector<uint64_t> psi_pow_, psi_inv_pow_;   // bit-reversed powers of psi / psi^-1
    uint64_t n_inv_;

    static uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)((__uint128_t)a * b % q); }
    static uint64_t powmod(uint64_t a, uint64_t e, uint64_t q) {
        uint64_t r = 1; a %= q;
        while (e) { if (e & 1) r = mulmod(r, a, q); a = mulmod(a, a, q); e >>= 1; }
        return r;
    }
    static size_t bitrev(size_t x, unsigned bits) {
        size_t r = 0;
        for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
        return r;
    }
public:
    NTT(uint64_t n, uint64_t q) : n_(n), q_(q), psi_pow_(n), psi_inv_pow_(n) {
        // find a primitive 2n-th root of unity: g^((q-1)/2n) with g^((q-1)/2) == -1 ... any psi with psi^n == -1
        uint64_t psi = 0;
        for (uint64_t g = 2; g < q; g++) {
            uint64_t c = powmod(g, (q - 1) / (2 * n), q);
            if (powmod(c, n, q) == q - 1) { psi = c; break; }
        }
        uint64_t psi_inv = powmod(psi, q - 2, q);
        unsigned bits = 0; while ((1ull << bits) < n) bits++;
        uint64_t a = 1, b = 1;
        for (size_t i = 0; i < n; i++) {
            psi_pow_[bitrev(i, bits)] = a;
            psi_inv_pow_[bitrev(i, bits)] = b;
            a = mulmod(a, psi, q);
            b = mulmod(b, psi_inv, q);
        }
        n_inv_ = powmod(n, q - 2, q);
    }

    void ComputeForward(uint64_t *out, const uint64_t *in, uint64_t, uint64_t) {
        if (out != in) for (size_t i = 0; i < n_; i++) out[i] = in[i];
        size_t t = n_;
        for (size_t m = 1; m < n_; m <<= 1) {
            t >>= 1;
            for (size_t i = 0; i < m; i++) {
                uint64_t w = psi_pow_[m + i];
                for (size_t j = 2 * i * t; j < 2 * i * t + t; j++) {
                    uint64_t u = out[j] % q_, v = mulmod(out[j + t] % q_, w, q_);
                    out[j] = (u + v) % q_;
                    out[j + t] = (u + q_ - v) % q_;
                }
            }
        }
    }

    void ComputeInverse(uint64_t *out, const uint64_t *in, uint64_t, uint64_t) {
        if (out != in) for (size_t i = 0; i < n_; i++) out[i] = in[i];
        size_t t = 1;
        for (size_t m = n_; m > 1; m >>= 1) {
            size_t h = m >> 1, j1 = 0;
            for (size_t i = 0; i < h; i++) {
                uint64_t w = psi_inv_pow_[h + i];
                for (size_t j = j1; j < j1 + t; j++) {
                    uint64_t u = out[j] % q_, v = out[j + t] % q_;
                    out[j] = (u + v) % q_;
                    out[j + t] = mulmod((u + q_ - v) % q_, w, q_);
                }
                j1 += 2 * t;
            }
            t <<= 1;
        }
        for (size_t i = 0; i < n_; i++) out[i] = mulmod(out[i], n_inv_, q_);
    }
};

}}  // namespace intel::hexl
