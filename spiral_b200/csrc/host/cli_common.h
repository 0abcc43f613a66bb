// cli_common.h - shared by pir_server.cpp / pir_client.cpp: the two halves of a real client/server split, written against the
// C-ABI only (include/spiral_b200.h).  The reference runs both halves in one process (do_test, src/spiral.cpp:2408); its CLI
// reserves `--server` / `--client` / `--loaddata` / `--input` flags (src/spiral.cpp:1252-1300) without implementing them.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/spiral_b200.h"

[[noreturn]] inline void die(const std::string &what) {
    fprintf(stderr, "error: %s%s%s\n", what.c_str(), *sb200_last_error() ? ": " : "", sb200_last_error());
    exit(1);
}
#define OK(call) do { if ((call) != 0) die(#call); } while (0)

// "nu1,nu2,t_gsw,t_conv,t_exp,t_exp_right,qp_bits,p_db"  (the reference's argv[1..2] and -D macros, include/values.h:78-93)
inline sb200_params parse_params(const char *s) {
    unsigned long long v[8];
    if (sscanf(s, "%llu,%llu,%llu,%llu,%llu,%llu,%llu,%llu", &v[0], &v[1], &v[2], &v[3], &v[4], &v[5], &v[6], &v[7]) != 8)
        die("--params wants nu1,nu2,t_gsw,t_conv,t_exp,t_exp_right,qp_bits,p_db");
    sb200_params p{};
    p.nu1 = (uint32_t)v[0]; p.nu2 = (uint32_t)v[1]; p.t_gsw = (uint32_t)v[2]; p.t_conv = (uint32_t)v[3];
    p.t_exp = (uint32_t)v[4]; p.t_exp_right = (uint32_t)v[5]; p.qp_bits = (uint32_t)v[6]; p.out_n = 2; p.p_db = v[7];
    return p;
}
inline std::vector<uint8_t> read_file(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) die(std::string("cannot open ") + path);
    fseek(f, 0, SEEK_END); long n = ftell(f); fseek(f, 0, SEEK_SET);
    std::vector<uint8_t> b((size_t)n);
    if (n && fread(b.data(), 1, (size_t)n, f) != (size_t)n) die(std::string("short read on ") + path);
    fclose(f);
    return b;
}
inline void write_file(const char *path, const void *p, size_t n) {
    FILE *f = fopen(path, "wb");
    if (!f || fwrite(p, 1, n, f) != n) die(std::string("cannot write ") + path);
    fclose(f);
}
// public-parameter file: "SB2P", four polynomial counts, then the four ref-NTT matrices (W_exp_left, W_exp_right, W_conv, V_conv)
struct PubHeader { uint32_t magic; uint32_t reserved; uint64_t polys[4]; };
constexpr uint32_t kPubMagic = 0x50324253u;
inline const char *arg(int argc, char **argv, const char *name, const char *dflt = nullptr) {
    for (int i = 2; i + 1 < argc; i++) if (!strcmp(argv[i], name)) return argv[i + 1];
    if (!dflt) die(std::string("missing ") + name);
    return dflt;
}
inline std::vector<const char *> args(int argc, char **argv, const char *name) {
    std::vector<const char *> out;
    for (int i = 2; i + 1 < argc; i++) if (!strcmp(argv[i], name)) out.push_back(argv[i + 1]);
    return out;
}
