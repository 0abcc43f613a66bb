// pir_server.cpp - the server half of a client/server split (C-ABI only): database from a record file or a snapshot, public
// parameters from the client's file, wire queries in, packed responses out.
//   pir_server --params P (--db records.bin | --snapshot db.sb2d) [--save-snapshot db.sb2d] --pp pp.bin
//              --query q0.bin [--query q1.bin ...] --out-prefix resp      -> resp.0, resp.1, ...
// Replaces, for a deployment, what do_test (src/spiral.cpp:2408) does in one process: load_db, the server statements of
// runConversionImproved / process_query_fast, modswitch.
#include <chrono>

#include "cli_common.h"

int main(int argc, char **argv) {
    if (argc < 2) die("usage: pir_server --params ... (see the file header)");
    // arg() skips a sub-command word (pir_client has one): give it one here too
    std::vector<char *> av{argv[0], (char *)"serve"};
    for (int i = 1; i < argc; i++) av.push_back(argv[i]);
    const int ac = (int)av.size();
    const sb200_params prm = parse_params(arg(ac, av.data(), "--params"));
    const int device = atoi(arg(ac, av.data(), "--device", "0"));
    OK(sb200_init(device));
    sb200_server *srv = nullptr;
    OK(sb200_server_create(&srv, &prm, device, 0, 1));
    const char *db = arg(ac, av.data(), "--db", ""), *snap = arg(ac, av.data(), "--snapshot", "");
    auto t0 = std::chrono::steady_clock::now();
    if (*snap) OK(sb200_server_load_db_snapshot(srv, snap));
    else if (*db) OK(sb200_server_load_db_records_file(srv, db));
    else die("one of --db / --snapshot is required");
    fprintf(stderr, "database resident in HBM after %.1f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    const char *save = arg(ac, av.data(), "--save-snapshot", "");
    if (*save) OK(sb200_server_save_db(srv, save));

    const std::vector<uint8_t> pp = read_file(arg(ac, av.data(), "--pp"));
    PubHeader h;
    if (pp.size() < sizeof(h)) die("--pp: not a public-parameter file");
    memcpy(&h, pp.data(), sizeof(h));
    // the file comes from the client: its four counts must be exactly what THIS server (shapes derived from --params) will read
    size_t expect[4], total = 0;
    OK(sb200_server_public_param_polys(srv, expect));
    if (h.magic != kPubMagic) die("--pp: not a public-parameter file");
    for (int i = 0; i < 4; i++) {
        if (h.polys[i] != expect[i]) die("--pp: public parameters were generated for other scheme parameters (matrix " + std::to_string(i) + ")");
        total += expect[i];
    }
    if (pp.size() != sizeof(h) + total * 2 * SB200_POLY_LEN * 8) die("--pp: truncated or padded public-parameter file");
    const uint64_t *m[4], *p = reinterpret_cast<const uint64_t *>(pp.data() + sizeof(h));
    for (int i = 0; i < 4; i++) { m[i] = p; p += h.polys[i] * 2 * SB200_POLY_LEN; }
    OK(sb200_server_set_public_params(srv, m[0], m[1], m[2], m[3]));

    const std::string prefix = arg(ac, av.data(), "--out-prefix");
    std::vector<uint64_t> packed(sb200_server_packed_response_bytes(srv) / 8);
    int n = 0;
    for (const char *q : args(ac, av.data(), "--query")) {
        const std::vector<uint8_t> wire = read_file(q);
        t0 = std::chrono::steady_clock::now();
        OK(sb200_server_answer_wire(srv, wire.data(), wire.size(), packed.data(), nullptr));
        fprintf(stderr, "query %d answered in %.3f ms (host clock, %zu bytes in, %zu bytes out)\n", n,
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(), wire.size(), packed.size() * 8);
        write_file((prefix + "." + std::to_string(n++)).c_str(), packed.data(), packed.size() * 8);
    }
    sb200_server_destroy(srv);
    return 0;
}
