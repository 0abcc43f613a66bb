// pir_client.cpp - the client half of a client/server split, on the GPU through sb200_client_* (C-ABI only).
//   pir_client keygen --params P --seed seed.bin --out pp.bin
//   pir_client query  --params P --seed seed.bin --idx N --query-id K [--wire-seed ws.bin] --out query.bin   (K: never reuse)
//   pir_client decode --params P --seed seed.bin --in resp.bin --out item.bin
// The 32-byte seed IS the client's state: keys are re-derived from it (the reference's keygen draws from an unseeded
// std::random_device and keeps the key in globals, src/client.cpp:3-5,23-47).  item.bin holds the decoded plaintext matrix in the
// record-stream format (log2(p_db) bits per coefficient).
#include "cli_common.h"

int main(int argc, char **argv) {
    if (argc < 2) die("usage: pir_client keygen|query|decode --params ... (see the file header)");
    const std::string cmd = argv[1];
    const sb200_params prm = parse_params(arg(argc, argv, "--params"));
    const std::vector<uint8_t> seed = read_file(arg(argc, argv, "--seed"));
    if (seed.size() != 32) die("--seed wants a 32-byte file");
    OK(sb200_init(atoi(arg(argc, argv, "--device", "0"))));
    sb200_client *c = nullptr;
    OK(sb200_client_create(&c, &prm, atoi(arg(argc, argv, "--device", "0")), seed.data()));
    if (cmd == "keygen") {
        PubHeader h{kPubMagic, 0, {0, 0, 0, 0}};
        size_t polys[4];
        OK(sb200_client_public_param_polys(c, polys));
        size_t total = 0;
        for (int i = 0; i < 4; i++) { h.polys[i] = polys[i]; total += polys[i]; }
        std::vector<uint64_t> buf(sizeof(h) / 8 + total * 2 * SB200_POLY_LEN);
        memcpy(buf.data(), &h, sizeof(h));
        uint64_t *m[4], *p = buf.data() + sizeof(h) / 8;
        for (int i = 0; i < 4; i++) { m[i] = p; p += polys[i] * 2 * SB200_POLY_LEN; }
        OK(sb200_client_public_params(c, m[0], m[1], m[2], m[3]));
        write_file(arg(argc, argv, "--out"), buf.data(), buf.size() * 8);
    } else if (cmd == "query") {
        // --query-id is mandatory: it names the noise stream and (without --wire-seed) the row-0 seed; reusing one under the
        // same client seed breaks the privacy of both queries.  --wire-seed (a 32-byte file) overrides the derived row-0 seed.
        const uint32_t qid = (uint32_t)strtoul(arg(argc, argv, "--query-id"), nullptr, 10);
        std::vector<uint8_t> ws;
        if (*arg(argc, argv, "--wire-seed", "")) { ws = read_file(arg(argc, argv, "--wire-seed")); if (ws.size() != 32) die("--wire-seed wants a 32-byte file"); }
        std::vector<uint8_t> wire(sb200_wire_query_bytes(SB200_WIRE_QUERY_SEEDED));
        OK(sb200_client_query_wire(c, strtoull(arg(argc, argv, "--idx"), nullptr, 10), qid, ws.empty() ? nullptr : ws.data(), wire.data()));
        write_file(arg(argc, argv, "--out"), wire.data(), wire.size());
    } else if (cmd == "decode") {
        const std::vector<uint8_t> packed = read_file(arg(argc, argv, "--in"));
        const size_t n0 = 2 * SB200_POLY_LEN, n1 = 4 * SB200_POLY_LEN;
        if (packed.size() != 8 * sb200_packed_response_words(n0, n1, prm.qp_bits, prm.p_db)) die("--in is not a packed response for these parameters");
        std::vector<uint64_t> words(packed.size() / 8), resp(6 * SB200_POLY_LEN), pt(4 * SB200_POLY_LEN);
        memcpy(words.data(), packed.data(), packed.size());
        OK(sb200_unpack_response(resp.data(), words.data(), n0, n1, prm.qp_bits, prm.p_db));
        OK(sb200_client_decode(c, resp.data(), pt.data()));
        uint32_t bits = 0;
        while ((1ull << bits) < prm.p_db) bits++;
        std::vector<uint8_t> item((pt.size() * bits + 7) / 8, 0);
        for (size_t k = 0; k < pt.size(); k++)
            for (uint32_t b = 0; b < bits; b++)
                if ((pt[k] >> b) & 1) item[(k * bits + b) >> 3] |= (uint8_t)(1u << ((k * bits + b) & 7));
        write_file(arg(argc, argv, "--out"), item.data(), item.size());
    } else {
        die("unknown command " + cmd);
    }
    sb200_client_destroy(c);
    return 0;
}
