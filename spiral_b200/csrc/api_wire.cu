// api_wire.cu - C-ABI of the wire / on-disk formats (kernels in wire_kernels.cu).
// Part of the single translation unit spiral_b200.cu (uses the server structs of api.cu / api_pack.cu).
//
//   * queries in wire form (seed-compressed or fully packed) for both resident servers;
//   * databases from a flat stream of records, in memory or in a file: the reference's `has_data` / `has_file`
//     branches of load_db, left as `// TODO` (src/spiral.cpp:1095-1162);
//   * snapshots of the PREPROCESSED database (`has_file && load` / `has_file && !load` of the same function):
//     header + the scan-layout shard as it sits in HBM + a device-computed integrity word.
// Host side of the I/O: pread into two pinned buffers, so reading chunk k+1 overlaps the H2D copy and the
// preprocessing kernels of chunk k.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

extern "C" size_t sb200_wire_query_bytes(uint32_t kind) {
    if (kind == kWireSeeded) return kWireHeaderBytes + kWireSeedBytes + kWireRowBytes;
    if (kind == kWireFull) return kWireHeaderBytes + 2 * kWireRowBytes;
    return 0;
}
extern "C" int sb200_dev_query_from_wire(uint32_t *cv_dev, const uint8_t *wire_dev, uint32_t kind, void *stream) {
    NEED_DEVICE();
    if (!cv_dev || !wire_dev || sb200_wire_query_bytes(kind) == 0) return fail(SB200_ERR_ARG, "query_from_wire: bad argument");
    launch_query_from_wire(cv_dev, wire_dev, kind, S(stream)); CHECK_LAUNCH();
    return SB200_OK;
}
namespace {
int wire_query_kind(const uint8_t *wire, size_t bytes, uint32_t *kind) {
    if (!wire || bytes < kWireHeaderBytes) return fail(SB200_ERR_ARG, "wire query: buffer shorter than its header");
    uint32_t magic; memcpy(&magic, wire, 4);
    if (magic != kWireQueryMagic) return fail(SB200_ERR_ARG, "wire query: bad magic 0x%08x", magic);
    const uint32_t k = (uint32_t)wire[4] | ((uint32_t)wire[5] << 8);
    if (wire[6] || wire[7] || sb200_wire_query_bytes(k) == 0) return fail(SB200_ERR_ARG, "wire query: unknown kind %u", k);
    if (bytes != sb200_wire_query_bytes(k)) return fail(SB200_ERR_ARG, "wire query: %zu bytes, kind %u needs %zu", bytes, k, sb200_wire_query_bytes(k));
    *kind = k;
    return SB200_OK;
}
}  // namespace
extern "C" int sb200_server_upload_query_wire(sb200_server *s, const uint8_t *wire_host, size_t bytes, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    uint32_t kind = 0;
    TRY(wire_query_kind(wire_host, bytes, &kind));
    CU(cudaMemcpyAsync(s->q_wire.p, wire_host, bytes, cudaMemcpyHostToDevice, ES(s, stream)));
    s->wire_kind = kind;
    return SB200_OK;      // unpacking / seed expansion is the first node of the expand_and_convert stage
}
// the whole exchange in wire form: 14 KiB query in, QPBITS-packed response out
extern "C" int sb200_server_answer_wire(sb200_server *s, const uint8_t *wire_host, size_t bytes, uint64_t *packed_resp_host, void *stream) {
    if (!s || !packed_resp_host) return fail(SB200_ERR_ARG, "server_answer_wire: null argument");
    if (s->world != 1) return fail(SB200_ERR_STATE, "server_answer_wire: single-shard call on a sharded server (use the staged API)");
    TRY(sb200_server_upload_query_wire(s, wire_host, bytes, stream));
    TRY(sb200_server_process(s, s->resp.p, stream, nullptr));
    TRY(sb200_dev_pack_response(s->final_ct.p, s->resp.p, 2 * (size_t)kN, 4 * (size_t)kN, s->prm.qp_bits, s->prm.p_db, ES(s, stream)));
    return sb200_server_download(s, packed_resp_host, s->final_ct.p, sb200_packed_response_words(2 * (size_t)kN, 4 * (size_t)kN, s->prm.qp_bits, s->prm.p_db), stream);
}
extern "C" int sb200_pack_server_upload_query_wire(sb200_pack_server *s, const uint8_t *wire_host, size_t bytes, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    uint32_t kind = 0;
    TRY(wire_query_kind(wire_host, bytes, &kind));
    CU(cudaMemcpyAsync(s->q_wire.p, wire_host, bytes, cudaMemcpyHostToDevice, PS(s, stream)));
    s->wire_kind = kind;
    return SB200_OK;
}
extern "C" int sb200_pack_server_answer_wire(sb200_pack_server *s, const uint8_t *wire_host, size_t bytes, uint64_t *total_resp_host,
                                             uint64_t *result_cts_host, void *stream) {
    if (!s || !total_resp_host) return fail(SB200_ERR_ARG, "pack answer_wire: null argument");
    TRY(sb200_pack_server_upload_query_wire(s, wire_host, bytes, stream));
    TRY(sb200_pack_server_expand_and_convert(s, stream));
    return pack_process(s, total_resp_host, result_cts_host, stream);
}
extern "C" uint64_t *sb200_pack_server_db_ptr(sb200_pack_server *s) { return s ? const_cast<uint64_t *>(pack_db(s)) : nullptr; }

// ---------------------------------------------------------------------------------------------
// records and snapshots
// ---------------------------------------------------------------------------------------------
namespace {
struct ByteSource {            // a record stream in memory or in a file
    const uint8_t *mem = nullptr; int fd = -1; size_t size = 0;
    int read(void *dst, size_t off, size_t n) const {
        if (off + n > size) return fail(SB200_ERR_ARG, "record stream: read of %zu bytes at %zu beyond its %zu bytes", n, off, size);
        if (mem) { memcpy(dst, mem + off, n); return SB200_OK; }
        for (size_t done = 0; done < n;) {
            ssize_t r = pread(fd, (char *)dst + done, n - done, (off_t)(off + done));
            if (r <= 0) return fail(SB200_ERR_ARG, "record stream: pread failed at offset %zu", off + done);
            done += (size_t)r;
        }
        return SB200_OK;
    }
};
struct FileFd {
    int fd = -1;
    ~FileFd() { if (fd >= 0) close(fd); }
};
struct PinnedPair {            // two pinned staging buffers + the events that guard their reuse
    void *p[2] = {nullptr, nullptr}; cudaEvent_t ev[2] = {nullptr, nullptr}; cudaStream_t st = nullptr;
    ~PinnedPair() {
        for (int i = 0; i < 2; i++) { if (p[i]) cudaFreeHost(p[i]); if (ev[i]) cudaEventDestroy(ev[i]); }
        if (st) cudaStreamDestroy(st);
    }
    cudaError_t init(size_t bytes) {
        for (int i = 0; i < 2; i++) {
            cudaError_t e = cudaHostAlloc(&p[i], bytes, cudaHostAllocDefault); if (e != cudaSuccess) return e;
            e = cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming); if (e != cudaSuccess) return e;
        }
        return cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    }
};
uint32_t coeff_bits(uint64_t p_db) {               // log2(p_db) for a power of two <= 2^16, else 0
    uint32_t b = 0;
    while (b <= 16 && (1ull << b) != p_db) b++;
    return b >= 1 && b <= 16 ? b : 0;
}
// Walk this shard's items (local order L = j*local_num_per + ii_local <-> global item j*num_per + rank + ii_local*world) in chunks:
// fill a pinned buffer from the source, copy it to `dev`, hand the chunk to `consume(dev, L0, n, stream)`.
template <typename F>
int stream_records(const ByteSource &src, size_t item_bytes, size_t dim0, size_t num_per, int rank, int world, size_t chunk_items, F &&consume) {
    const size_t lnp = num_per / (size_t)world, total = dim0 * lnp;
    if (src.size != dim0 * num_per * item_bytes)
        return fail(SB200_ERR_ARG, "record stream: %zu bytes, the database holds %zu items of %zu bytes", src.size, dim0 * num_per, item_bytes);
    chunk_items = std::min(chunk_items, total);
    PinnedPair pin;
    CU(pin.init(chunk_items * item_bytes));
    DBuf<uint8_t> dev;
    CU(dev.alloc(chunk_items * item_bytes + 4));
    for (size_t L0 = 0, k = 0; L0 < total; L0 += chunk_items, k++) {
        const int b = (int)(k & 1);
        const size_t n = std::min(chunk_items, total - L0);
        if (k >= 2) CU(cudaEventSynchronize(pin.ev[b]));
        if (world == 1) {
            TRY(src.read(pin.p[b], L0 * item_bytes, n * item_bytes));
        } else {
            for (size_t i = 0; i < n; i++) {
                const size_t L = L0 + i, j = L / lnp, iil = L % lnp;
                TRY(src.read((uint8_t *)pin.p[b] + i * item_bytes, (j * num_per + (size_t)rank + iil * (size_t)world) * item_bytes, item_bytes));
            }
        }
        CU(cudaMemcpyAsync(dev.p, pin.p[b], n * item_bytes, cudaMemcpyHostToDevice, pin.st));
        CU(cudaEventRecord(pin.ev[b], pin.st));
        TRY(consume(dev.p, L0, n, pin.st));
    }
    CU(cudaStreamSynchronize(pin.st));
    CHECK_LAUNCH();
    return SB200_OK;
}
int open_source(const char *path, FileFd &f, ByteSource &src) {
    if (!path) return fail(SB200_ERR_ARG, "null path");
    f.fd = open(path, O_RDONLY);
    if (f.fd < 0) return fail(SB200_ERR_ARG, "cannot open %s", path);
    struct stat st;
    if (fstat(f.fd, &st) != 0) return fail(SB200_ERR_ARG, "cannot stat %s", path);
    src.fd = f.fd; src.size = (size_t)st.st_size;
    return SB200_OK;
}

int server_load_records(sb200_server *s, const ByteSource &src) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    const uint32_t bits = coeff_bits(s->prm.p_db);
    if (!bits) return fail(SB200_ERR_ARG, "load_db_records: p_db must be a power of two <= 65536");
    CU(cudaSetDevice(s->device));
    TRY(server_alloc_db(s));
    const size_t item_bytes = 4 * (size_t)kN * bits / 8, chunk = 4096;
    DBuf<uint16_t> pts;
    CU(pts.alloc(std::min(chunk, s->dim0 * s->local_num_per) * 4 * kN));
    const uint32_t nu2_local = (uint32_t)ceil_log2(s->local_num_per);
    TRY(stream_records(src, item_bytes, s->dim0, s->num_per, s->rank, s->world, chunk, [&](const uint8_t *dev, size_t L0, size_t n, cudaStream_t st) {
        launch_records_to_pts(pts.p, dev, n, 4, bits, 4 * (size_t)kN, (size_t)kN, st);
        launch_db_build_spiral(s->db.p, pts.p, (int)s->prm.nu1, (int)nu2_local, (uint32_t)s->prm.p_db, L0, n, st);
        return SB200_OK;
    }));
    s->have_db = true;
    return SB200_OK;
}
int pack_server_load_records(sb200_pack_server *s, const ByteSource &src) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->db_owner) return fail(SB200_ERR_STATE, "this pack server is a view: load the database through its parent");
    if (s->shard_planes) return fail(SB200_ERR_STATE, "plane-sharded pack servers load their planes with sb200_pack_server_load_plane_*");
    const uint32_t bits = coeff_bits(s->prm.p_db);
    if (!bits) return fail(SB200_ERR_ARG, "load_db_records: p_db must be a power of two <= 65536");
    NEED_SCAN_PLANES(s, "pack load_db_records");
    CU(cudaSetDevice(s->device));
    const size_t items = s->dim0 * s->local_num_per, item_bytes = s->planes * (size_t)kN * bits / 8;
    DBuf<uint16_t> pts;                                        // [plane][local item][2048]
    CU(pts.alloc(s->planes * items * kN));
    const size_t chunk = std::max<size_t>(1, ((size_t)32 << 20) / item_bytes);
    TRY(stream_records(src, item_bytes, s->dim0, s->num_per, s->rank, s->world, chunk, [&](const uint8_t *dev, size_t L0, size_t n, cudaStream_t st) {
        launch_records_to_pts(pts.p + L0 * kN, dev, n, (int)s->planes, bits, (size_t)kN, items * kN, st);
        return SB200_OK;
    }));
    for (size_t p = 0; p < s->planes; p++) {
        launch_db_build_pack(s->db.p + p * s->plane_words, pts.p + p * items * kN, s->dim0, s->local_num_per, (uint32_t)s->prm.p_db, 0); CHECK_LAUNCH();
        s->plane_loaded[p] = true;
    }
    CU(cudaDeviceSynchronize());
    return SB200_OK;
}

// ---- snapshots of the preprocessed database --------------------------------------------------
struct SnapHeader {
    uint32_t magic, version, kind, nu1, nu2, rank, world, out_n;
    uint64_t p_db, words, sum, reserved;
};
static_assert(sizeof(SnapHeader) == 64, "snapshot header is 64 bytes");
constexpr uint32_t kSnapMagic = 0x44324253u;      // "SB2D"
constexpr size_t kSnapChunk = (size_t)64 << 20;

int device_sum64(const uint64_t *dev, size_t words, uint64_t *out) {
    DBuf<unsigned long long> acc;
    CU(acc.alloc(1));
    CU(cudaMemset(acc.p, 0, 8));
    launch_sum64(acc.p, dev, words, 0); CHECK_LAUNCH();
    unsigned long long h = 0;
    CU(acc.down(&h, 1));
    *out = (uint64_t)h;
    return SB200_OK;
}
int write_all(int fd, const void *p, size_t n) {
    for (size_t done = 0; done < n;) {
        ssize_t r = write(fd, (const char *)p + done, n - done);
        if (r <= 0) return fail(SB200_ERR_ARG, "snapshot: write failed");
        done += (size_t)r;
    }
    return SB200_OK;
}
int snapshot_save(const char *path, SnapHeader h, const uint64_t *dev, size_t words) {
    if (!path) return fail(SB200_ERR_ARG, "null path");
    h.magic = kSnapMagic; h.version = 1; h.words = words; h.reserved = 0;
    TRY(device_sum64(dev, words, &h.sum));
    FileFd f;
    f.fd = open(path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (f.fd < 0) return fail(SB200_ERR_ARG, "cannot create %s", path);
    TRY(write_all(f.fd, &h, sizeof(h)));
    PinnedPair pin;
    CU(pin.init(kSnapChunk));
    const size_t bytes = words * 8, nchunks = (bytes + kSnapChunk - 1) / kSnapChunk;
    auto issue = [&](size_t k) {
        const size_t o = k * kSnapChunk, n = std::min(kSnapChunk, bytes - o);
        cudaMemcpyAsync(pin.p[k & 1], (const uint8_t *)dev + o, n, cudaMemcpyDeviceToHost, pin.st);
        cudaEventRecord(pin.ev[k & 1], pin.st);
    };
    if (nchunks) issue(0);
    for (size_t k = 0; k < nchunks; k++) {
        if (k + 1 < nchunks) issue(k + 1);                     // the next D2H copy runs while this chunk is written
        CU(cudaEventSynchronize(pin.ev[k & 1]));
        TRY(write_all(f.fd, pin.p[k & 1], std::min(kSnapChunk, bytes - k * kSnapChunk)));
    }
    CHECK_LAUNCH();
    return SB200_OK;
}
int snapshot_load(const char *path, const SnapHeader &want, uint64_t *dev, size_t words) {
    FileFd f; ByteSource src;
    TRY(open_source(path, f, src));
    SnapHeader h;
    if (src.size < sizeof(h)) return fail(SB200_ERR_ARG, "snapshot %s: shorter than its header", path);
    TRY(src.read(&h, 0, sizeof(h)));
    if (h.magic != kSnapMagic || h.version != 1) return fail(SB200_ERR_ARG, "snapshot %s: bad magic / version", path);
    if (h.kind != want.kind || h.nu1 != want.nu1 || h.nu2 != want.nu2 || h.rank != want.rank || h.world != want.world ||
        h.out_n != want.out_n || h.p_db != want.p_db || h.words != words)
        return fail(SB200_ERR_ARG, "snapshot %s: written for another server (kind %u nu1 %u nu2 %u shard %u/%u out_n %u p_db %llu, %llu words)", path,
                    h.kind, h.nu1, h.nu2, h.rank, h.world, h.out_n, (unsigned long long)h.p_db, (unsigned long long)h.words);
    if (src.size != sizeof(h) + words * 8) return fail(SB200_ERR_ARG, "snapshot %s: truncated", path);
    PinnedPair pin;
    CU(pin.init(kSnapChunk));
    const size_t bytes = words * 8;
    for (size_t o = 0, k = 0; o < bytes; o += kSnapChunk, k++) {
        const size_t n = std::min(kSnapChunk, bytes - o);
        if (k >= 2) CU(cudaEventSynchronize(pin.ev[k & 1]));
        TRY(src.read(pin.p[k & 1], sizeof(h) + o, n));
        CU(cudaMemcpyAsync((uint8_t *)dev + o, pin.p[k & 1], n, cudaMemcpyHostToDevice, pin.st));
        CU(cudaEventRecord(pin.ev[k & 1], pin.st));
    }
    CU(cudaStreamSynchronize(pin.st));
    uint64_t sum = 0;
    TRY(device_sum64(dev, words, &sum));
    if (sum != h.sum) return fail(SB200_ERR_ARG, "snapshot %s: integrity word mismatch (file says %016llx, data sums to %016llx)", path,
                                  (unsigned long long)h.sum, (unsigned long long)sum);
    return SB200_OK;
}
SnapHeader server_snap_header(const sb200_server *s) {
    SnapHeader h{};
    h.kind = 1; h.nu1 = s->prm.nu1; h.nu2 = s->prm.nu2; h.rank = (uint32_t)s->rank; h.world = (uint32_t)s->world; h.out_n = 0; h.p_db = s->prm.p_db;
    return h;
}
SnapHeader pack_snap_header(const sb200_pack_server *s) {
    SnapHeader h{};
    h.kind = 2; h.nu1 = s->prm.nu1; h.nu2 = s->prm.nu2; h.rank = (uint32_t)s->rank; h.world = (uint32_t)s->world; h.out_n = s->prm.out_n; h.p_db = s->prm.p_db;
    return h;
}
}  // namespace

extern "C" size_t sb200_server_record_stream_bytes(const sb200_server *s) {
    return s && coeff_bits(s->prm.p_db) ? s->dim0 * s->num_per * 4 * (size_t)kN * coeff_bits(s->prm.p_db) / 8 : 0;
}
extern "C" int sb200_server_load_db_records(sb200_server *s, const uint8_t *records_host, size_t bytes) {
    if (!records_host) return fail(SB200_ERR_ARG, "load_db_records: null argument");
    ByteSource src; src.mem = records_host; src.size = bytes;
    return server_load_records(s, src);
}
extern "C" int sb200_server_load_db_records_file(sb200_server *s, const char *path) {
    FileFd f; ByteSource src;
    TRY(open_source(path, f, src));
    return server_load_records(s, src);
}
extern "C" int sb200_server_save_db(sb200_server *s, const char *path) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->db_owner || !s->have_db) return fail(SB200_ERR_STATE, "save_db: this server owns no loaded database");
    if (server_z_slices(s) != (size_t)kN) return fail(SB200_ERR_STATE, "save_db: an implicit database has no snapshot form");
    if (!s->db.p) return fail(SB200_ERR_STATE, "save_db: the scan-layout copy was released (sb200_server_tc_only); snapshots hold that layout");
    CU(cudaSetDevice(s->device));
    return snapshot_save(path, server_snap_header(s), s->db.p, s->dim0 * s->local_num_per * 4 * kN);
}
extern "C" int sb200_server_load_db_snapshot(sb200_server *s, const char *path) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    CU(cudaSetDevice(s->device));
    TRY(server_alloc_db(s));
    s->have_db = false;
    TRY(snapshot_load(path, server_snap_header(s), s->db.p, s->dim0 * s->local_num_per * 4 * kN));
    s->have_db = true;
    return SB200_OK;
}
extern "C" size_t sb200_pack_server_record_stream_bytes(const sb200_pack_server *s) {
    return s && coeff_bits(s->prm.p_db) ? s->dim0 * s->num_per * s->planes * (size_t)kN * coeff_bits(s->prm.p_db) / 8 : 0;
}
extern "C" int sb200_pack_server_load_db_records(sb200_pack_server *s, const uint8_t *records_host, size_t bytes) {
    if (s && s->shard_planes) return fail(SB200_ERR_STATE, "plane-sharded pack servers load their planes with sb200_pack_server_load_plane_*");
    if (!records_host) return fail(SB200_ERR_ARG, "load_db_records: null argument");
    ByteSource src; src.mem = records_host; src.size = bytes;
    return pack_server_load_records(s, src);
}
extern "C" int sb200_pack_server_load_db_records_file(sb200_pack_server *s, const char *path) {
    FileFd f; ByteSource src;
    TRY(open_source(path, f, src));
    return pack_server_load_records(s, src);
}
extern "C" int sb200_pack_server_save_db(sb200_pack_server *s, const char *path) {
    if (s && s->shard_planes) return fail(SB200_ERR_STATE, "plane-sharded pack servers load their planes with sb200_pack_server_load_plane_*");
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->db_owner) return fail(SB200_ERR_STATE, "save_db: this pack server is a view");
    for (size_t p = 0; p < s->planes; p++) if (!s->plane_loaded[p]) return fail(SB200_ERR_STATE, "save_db: plane %zu not loaded", p);
    NEED_SCAN_PLANES(s, "pack save_db");
    CU(cudaSetDevice(s->device));
    return snapshot_save(path, pack_snap_header(s), s->db.p, s->planes * s->plane_words);
}
extern "C" int sb200_pack_server_load_db_snapshot(sb200_pack_server *s, const char *path) {
    if (s && s->shard_planes) return fail(SB200_ERR_STATE, "plane-sharded pack servers load their planes with sb200_pack_server_load_plane_*");
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->db_owner) return fail(SB200_ERR_STATE, "this pack server is a view: load the database through its parent");
    NEED_SCAN_PLANES(s, "pack load_db_snapshot");
    CU(cudaSetDevice(s->device));
    for (size_t p = 0; p < s->planes; p++) s->plane_loaded[p] = false;
    TRY(snapshot_load(path, pack_snap_header(s), s->db.p, s->planes * s->plane_words));
    for (size_t p = 0; p < s->planes; p++) s->plane_loaded[p] = true;
    return SB200_OK;
}
