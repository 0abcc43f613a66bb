// api_pack.cu - C-ABI entry points of the SpiralPack / SpiralStreamPack path (include/spiral_b200.h).
// Included by spiral_b200.cu after api.cu (shares its helpers: DBuf, up_ntt, down_ntt, CU, TRY, fail).

namespace sb200 {
void launch_db_build_pack(uint64_t *db_plane, const uint16_t *pts_plane, size_t dim0, size_t num_per, uint32_t p_db, cudaStream_t s);
void launch_reorient_dim1(uint64_t *out, const uint32_t *cv, const int *ct_idx, size_t dim0, cudaStream_t s);
void launch_db_set_item_pack(uint64_t *db_plane, const uint16_t *poly, size_t dim0, size_t num_per, uint32_t p_db, size_t i, size_t j, cudaStream_t s);
void launch_scan_pack(uint32_t *out, const uint64_t *query, const uint64_t *db, size_t dim0, size_t num_per, size_t planes,
                      size_t db_plane_words, size_t out_plane_polys, cudaStream_t s);
void launch_regev_to_simple_gsw(uint32_t *gsw_out, uint32_t *gsw_neg_out, const uint32_t *cv, const int *ct_idx, const int *poly_idx,
                                int nu2, int ell, const uint32_t *V, int t_conv, uint64_t *scratch_raw, uint32_t *scratch_ntt, cudaStream_t s);
void launch_pack(uint32_t *result, const uint64_t *v_ct_raw, const uint32_t *vW, int out_n, int t_conv,
                 uint64_t *scratch_raw, uint32_t *scratch_ntt, cudaStream_t s);
}

// dev-NTT 1x1 plaintexts -> the reference's convertDb layout db_buf[z][ii][j] (pure relayout)
__global__ void k_convert_db_ref(uint64_t *__restrict__ out, const uint32_t *__restrict__ in, size_t count, size_t dim0, size_t num_per) {
    pdl_prologue();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // (item, z)
    if (idx >= count * sb200::kN) return;
    const size_t z = idx % sb200::kN, i = idx / sb200::kN, ii = i % num_per, j = i / num_per;
    const uint32_t *p = in + i * 2 * sb200::kN;
    out[z * (num_per * dim0) + ii * dim0 + j] = (uint64_t)p[z] | ((uint64_t)p[sb200::kN + z] << 32);
}

extern "C" int sb200_convertDb(uint64_t *db_buf, const uint64_t *db_ntt, size_t count, size_t dim0, size_t num_per) {
    NEED_DEVICE();
    DBuf<uint32_t> din; DBuf<uint64_t> dout(count * kN);
    TRY(up_ntt(din, db_ntt, count));
    const size_t n = count * kN;
    count_launch(); launch_pdl(k_convert_db_ref, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, 0, dout.p, din.p, count, dim0, num_per); CHECK_LAUNCH();
    CU(dout.down(db_buf, count * kN));
    return SB200_OK;
}
extern "C" int sb200_reorientCiphertextsDim1(uint64_t *out, const uint64_t *v_firstdim, size_t count, size_t dim0, size_t idx_factor) {
    NEED_DEVICE();
    if (dim0 * idx_factor > count + idx_factor - 1 && (dim0 - 1) * idx_factor >= count) return fail(SB200_ERR_ARG, "reorientCiphertextsDim1: not enough ciphertexts");
    DBuf<uint32_t> dcv; DBuf<uint64_t> dout(dim0 * 2 * kN); DBuf<int> idx(dim0);
    TRY(up_ntt(dcv, v_firstdim, count * 2));
    std::vector<int> h(dim0);
    for (size_t j = 0; j < dim0; j++) h[j] = (int)(j * idx_factor);
    CU(idx.up(h.data(), dim0));
    launch_reorient_dim1(dout.p, dcv.p, idx.p, dim0, 0); CHECK_LAUNCH();
    CU(dout.down(out, dim0 * 2 * kN));
    return SB200_OK;
}
extern "C" int sb200_fastMultiplyQueryByDatabaseDim1(uint64_t *out, const uint64_t *db, const uint64_t *v_firstdim, size_t dim0, size_t num_per) {
    NEED_DEVICE();
    if (dim0 < 2 || num_per < 1 || (dim0 & (dim0 - 1)) || (num_per & (num_per - 1))) return fail(SB200_ERR_ARG, "fastMultiplyQueryByDatabaseDim1: dim0 >= 2 and powers of two required");
    const size_t words = dim0 * num_per * kN;
    DBuf<uint64_t> dq(dim0 * 2 * kN), dref(words), ddb(words); DBuf<uint32_t> dout(num_per * 2 * PLW);
    CU(dq.up(v_firstdim, dim0 * 2 * kN)); CU(dref.up(db, words));
    // reference db_buf[z][ii][j] -> DBP[z][jp][i][s]: per z a (num_per x dim0/2) -> (dim0/2 x num_per) transpose of 16-byte pairs
    launch_db_from_reference(ddb.p, dref.p, dim0 / 2, num_per, 0, kN, 0); CHECK_LAUNCH();
    launch_scan_pack(dout.p, dq.p, ddb.p, dim0, num_per, 1, words, num_per * 2, 0); CHECK_LAUNCH();
    return down_ntt(out, dout.p, num_per * 2);
}
extern "C" int sb200_foldCiphertextsDim1(uint64_t *v_cts, size_t count, const uint64_t *v_folding, const uint64_t *v_folding_neg, uint32_t ell) {
    NEED_DEVICE();
    size_t fd = 0; while (((size_t)1 << fd) < count) fd++;
    const size_t gsw_polys = 2 * 2 * (size_t)ell;
    DBuf<uint64_t> dcts(count * 2 * kN); DBuf<uint32_t> df, dfn, scratch(fold_scratch_words_generic(count, 2, 1, (int)ell));
    CU(dcts.up(v_cts, count * 2 * kN));
    TRY(up_ntt(df, v_folding, fd * gsw_polys)); TRY(up_ntt(dfn, v_folding_neg, fd * gsw_polys));
    size_t np = count;
    for (size_t cur = 0; cur < fd; cur++) {
        np /= 2;
        const size_t d = fd - 1 - cur;
        launch_fold_round_generic(dcts.p, 2, 1, (int)ell, 0, np, 1, count, df.p + d * gsw_polys * PLW, dfn.p + d * gsw_polys * PLW, scratch.p, 0);
    }
    CHECK_LAUNCH();
    CU(dcts.down(v_cts, 2 * kN));            // the reference keeps the result in v_cts[0]
    return SB200_OK;
}
extern "C" int sb200_regevToSimpleGsw(uint64_t *v_gsw, const uint64_t *v_inp, size_t count_inp, const uint64_t *V, uint32_t t_conv,
                                      uint32_t ell, uint32_t further_dims, size_t idx_factor, size_t idx_offset) {
    NEED_DEVICE();
    const int nbits = (int)(ell * further_dims);
    if ((size_t)(nbits - 1) * idx_factor + idx_offset >= count_inp) return fail(SB200_ERR_ARG, "regevToSimpleGsw: not enough input ciphertexts");
    DBuf<uint32_t> dcv, dV, dout((size_t)further_dims * 2 * 2 * ell * PLW), sntt((size_t)2 * t_conv * nbits * PLW);
    DBuf<uint64_t> sraw((size_t)2 * nbits * kN); DBuf<int> ct_idx(nbits), poly_idx(2 * nbits);
    TRY(up_ntt(dcv, v_inp, count_inp * 2)); TRY(up_ntt(dV, V, 2 * 2 * (size_t)t_conv));
    std::vector<int> hc(nbits), hp(2 * nbits);
    for (int b = 0; b < nbits; b++) { hc[b] = (int)(idx_factor * b + idx_offset); hp[b] = 2 * hc[b]; hp[nbits + b] = 2 * hc[b] + 1; }
    CU(ct_idx.up(hc.data(), nbits)); CU(poly_idx.up(hp.data(), 2 * nbits));
    launch_regev_to_simple_gsw(dout.p, nullptr, dcv.p, ct_idx.p, poly_idx.p, (int)further_dims, (int)ell, dV.p, (int)t_conv, sraw.p, sntt.p, 0); CHECK_LAUNCH();
    return down_ntt(v_gsw, dout.p, (size_t)further_dims * 2 * 2 * ell);
}
extern "C" int sb200_pack(uint64_t *result, uint32_t out_n, uint32_t t_conv, const uint64_t *v_ct, const uint64_t *v_W) {
    NEED_DEVICE();
    const size_t nn = (size_t)out_n * out_n, rows = out_n + 1;
    DBuf<uint64_t> dct(nn * 2 * kN), sraw(nn * 2 * kN); DBuf<uint32_t> dW, dres(rows * out_n * PLW), sntt((t_conv + 1) * nn * PLW);
    CU(dct.up(v_ct, nn * 2 * kN)); TRY(up_ntt(dW, v_W, out_n * rows * t_conv));
    launch_pack(dres.p, dct.p, dW.p, (int)out_n, (int)t_conv, sraw.p, sntt.p, 0); CHECK_LAUNCH();
    return down_ntt(result, dres.p, rows * out_n);
}

// ---------------------------------------------------------------------------------------------
// resident Pack server (testHighRate's server statements, src/testing.cpp:1007-1081)
//
// Sharding: rank g of `world` owns the second-dimension indices ii = g (mod world) of EVERY plane, so all
// fold rounds for bits >= log2(world) pair ciphertexts on the same GPU; one exchange of `planes` surviving 2x1
// ciphertexts per GPU (32 KiB each) precedes the last log2(world) rounds, the packing and the modulus switch on rank 0.
// ---------------------------------------------------------------------------------------------
// gathered [world][planes] ciphertexts (2 x 2048 u64 each) -> [planes][world]: the fold kernels want a plane's ciphertexts contiguous
__global__ void k_transpose_cts(uint64_t *__restrict__ out, const uint64_t *__restrict__ in, int world, int planes) {
    pdl_prologue();
    const int ct = blockIdx.x, r = ct / planes, p = ct % planes;          // input ciphertext (r, p)
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(in) + (size_t)ct * sb200::kN;
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out) + ((size_t)p * world + r) * sb200::kN;
    for (int i = threadIdx.x; i < sb200::kN; i += blockDim.x) dst[i] = src[i];
}

// plane sharding: gathered [world][slot_words] holds rank r's planes (p = r, r + world, ...) in local order -> plane order
__global__ void k_gather_planes(uint64_t *__restrict__ out, const uint64_t *__restrict__ in, int world, size_t slot_words) {
    pdl_prologue();
    const int p = blockIdx.x, r = p % world, l = p / world;
    const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(in + (size_t)r * slot_words) + (size_t)l * sb200::kN;
    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(out) + (size_t)p * sb200::kN;
    for (int i = threadIdx.x; i < sb200::kN; i += blockDim.x) dst[i] = src[i];
}

struct sb200_pack_server {
    sb200_params prm;
    int device = 0, rank = 0, world = 1, log_world = 0;
    size_t dim0 = 0, num_per = 0, local_num_per = 0, planes = 0, plane_words = 0;
    size_t g = 0, stopround = 0;
    ExpandPlan plan{};
    std::vector<int> offs, cnt;
    int maxcnt = 0, tmax = 0;
    bool have_params = false;
    std::vector<bool> plane_loaded;
    const sb200_pack_server *db_owner = nullptr;        // views (sb200_pack_server_create_view) scan another server's resident planes
    DBuf<uint64_t> db;
    DBuf<uint8_t> db_tc, q_tc;                          // tensor-core path (sb200_pack_server_enable_tc)
    DBuf<uint32_t> tc_t1;
    int tc_capacity = 0;
    bool tc_only = false;                               // sb200_pack_server_tc_only: only the limb-tile planes are resident
    DBuf<uint32_t> W_left, W_right, V, vW, neg1;
    DBuf<uint64_t> stage;
    DBuf<uint8_t> q_wire;                               // query in its wire form (wire_kernels.cu)
    uint32_t wire_kind = 0;
    DBuf<uint32_t> cv, c1, ginv, conv_ntt, gsw, scan_out, fold_scratch, packed;
    DBuf<uint64_t> c0, conv_raw, query, cts, result_cts, tail_cts, packed_raw, resp;
    DBuf<int> lists, ct_idx_first, ct_idx_direct, ct_idx_bits, poly_idx_bits;
    DBuf<uint16_t> perms;
    GraphSlot g_convert, g_convert_wire[2], g_fold, g_tail, g_xchg;
    // sharded expansion (column-sharded servers with connected peers): this rank expands only the ancestors of the first-dimension
    // ciphertexts j = rank (mod world) - plus the whole odd chain, the GSW bits are few - and k_reorient_dim1_allgather stores its
    // 2^nu1 / world reoriented ciphertexts into every rank's query buffer (as the Spiral server's ScalToMat does)
    DBuf<int> lists_sh, ct_idx_first_s;
    std::vector<int> offs_sh, cnt_sh;
    bool expand_shard_eligible = false;
    GraphSlot g_convert_sh[3];
    cudaStream_t own_stream = nullptr;
    // sharding: 0 = second dimension strided (ii = rank mod world of EVERY plane; one exchange, then log2(world) tail folds);
    // 1 = whole planes (plane p lives on rank p mod world: no exchange before the packing; src/testing.cpp:1045-1061 runs
    // the out_n^2 planes as independent trials).  `planes` counts the LOCAL planes, `planes_total` = out_n^2.
    int shard_planes = 0;
    size_t planes_total = 0;
    std::vector<int> plane_ids;                          // global plane index of each local plane
    int query_mode = 0;                                  // 1 packed query, 2 direct upload, 3 direct upload split over the ranks
    bool query_wait_pending = false;
    // peer-memory exchange (xchg_kernels.cu), as sb200_server's
    DBuf<uint8_t> xchg;
    DBuf<unsigned int> xchg_state;                       // [0] epoch, [1] error
    DBuf<unsigned int *> xchg_acks;
    DBuf<uint64_t> gathered;                             // rank 0: [world][slot_words]
    size_t slot_words = 0;
    void *xchg_target = nullptr;
    std::vector<void *> ipc_opened;
    bool xchg_connected = false;
    QueryPeers qpeers{};                                 // every rank's query buffer + exchange header (split direct upload)
    ~sb200_pack_server() {
        if (own_stream) cudaStreamDestroy(own_stream);
        for (void *p : ipc_opened) cudaIpcCloseMemHandle(p);
    }
    int local_plane(size_t global_plane) const {
        for (size_t k = 0; k < plane_ids.size(); k++) if ((size_t)plane_ids[k] == global_plane) return (int)k;
        return -1;
    }
};
static inline cudaStream_t PS(sb200_pack_server *s, void *stream) { return stream ? (cudaStream_t)stream : s->own_stream; }
static inline const sb200_pack_server *pack_owner(const sb200_pack_server *s) { return s->db_owner ? s->db_owner : s; }
static inline const uint64_t *pack_db(const sb200_pack_server *s) { return pack_owner(s)->db.p; }
// loaders and the plane-at-a-time host path work on the scan-layout planes
#define NEED_SCAN_PLANES(s, what) do { if (!(s)->db.p) return fail(SB200_ERR_STATE, what ": the scan-layout planes were released (sb200_pack_server_tc_only); create a new server to load another database"); } while (0)
static inline bool pack_plane_loaded(const sb200_pack_server *s, size_t p) { return pack_owner(s)->plane_loaded[p]; }
static inline TcGeom pack_geom(const sb200_pack_server *s) { return tc_geom_pack(s->dim0, s->local_num_per, s->planes, s->local_num_per * 2); }

static int pack_server_create_impl(sb200_pack_server **out, const sb200_params *prm, int device, int rank, int world, const sb200_pack_server *parent,
                                   int shard_planes = 0) {
    if (!out || !prm) return fail(SB200_ERR_ARG, "pack_server_create: null argument");
    if (prm->out_n == 0 || prm->nu1 < 1) return fail(SB200_ERR_ARG, "pack_server_create: out_n >= 1 and nu1 >= 1 required");
    if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world) return fail(SB200_ERR_ARG, "pack_server_create: world must be a power of two and 0 <= rank < world");
    if (!shard_planes && ((size_t)1 << prm->nu2) < (size_t)world) return fail(SB200_ERR_ARG, "pack_server_create: 2^nu2 < world");
    if (shard_planes && (size_t)prm->out_n * prm->out_n < (size_t)world) return fail(SB200_ERR_ARG, "pack_server_create: fewer planes than ranks");
    if (world > 16) return fail(SB200_ERR_ARG, "pack_server_create: world > 16 not supported by the peer exchange");
    if (prm->t_gsw == 0 || prm->t_conv == 0 || prm->t_exp == 0 || prm->t_exp_right == 0) return fail(SB200_ERR_ARG, "pack_server_create: zero gadget length");
    if (sb200_arb_qprime(prm->qp_bits) == 0) return fail(SB200_ERR_ARG, "pack_server_create: no response modulus for qp_bits = %u (14..36)", prm->qp_bits);
    if (prm->p_db == 0 || prm->p_db > 65536) return fail(SB200_ERR_ARG, "pack_server_create: p_db must be in [1, 65536]");
    if (prm->nu1 > 11) return fail(SB200_ERR_ARG, "pack_server_create: nu1 > 11");
    int rc = sb200_init(device);
    if (rc) return rc;
    sb200_pack_server *s = new sb200_pack_server();
    s->prm = *prm; s->device = device; s->rank = rank; s->world = world; s->log_world = (int)ceil_log2((size_t)world);
    s->dim0 = (size_t)1 << prm->nu1; s->num_per = (size_t)1 << prm->nu2;
    s->shard_planes = shard_planes && world > 1;
    s->planes_total = (size_t)prm->out_n * prm->out_n;
    if (s->shard_planes) {
        s->local_num_per = s->num_per; s->log_world = 0;
        for (size_t pl = (size_t)rank; pl < s->planes_total; pl += (size_t)world) s->plane_ids.push_back((int)pl);
    } else {
        s->local_num_per = s->num_per / world;
        for (size_t pl = 0; pl < s->planes_total; pl++) s->plane_ids.push_back((int)pl);
    }
    s->planes = s->plane_ids.size();
    s->slot_words = (s->shard_planes ? (s->planes_total + world - 1) / world : s->planes_total) * 2 * (size_t)kN;
    s->plane_words = s->dim0 * s->local_num_per * kN;
    s->plane_loaded.assign(s->planes, false);
    const size_t ell = prm->t_gsw, nbits = ell * prm->nu2;
    // expansion shape (testHighRate :795-798)
    s->g = ceil_log2(nbits + s->dim0);
    s->stopround = ceil_log2(nbits ? nbits : 1);
    s->plan = ExpandPlan{(int)s->g, (int)prm->t_exp, (int)prm->t_exp_right, (int)s->stopround, (int)nbits};
    std::vector<int> list(expand_active_total(s->plan));
    s->offs.resize(s->g); s->cnt.resize(s->g);
    s->maxcnt = expand_build_lists(s->plan, list.data(), s->offs.data(), s->cnt.data());
    s->tmax = s->plan.t_left > s->plan.t_right ? s->plan.t_left : s->plan.t_right;
    const size_t ncts = std::max((size_t)1 << s->g, s->dim0 + nbits), rows = prm->out_n + 1;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    if (parent) s->db_owner = parent; else A(s->db.alloc(s->planes * s->plane_words));
    A(s->W_left.alloc(s->g * 2 * prm->t_exp * PLW)); A(s->W_right.alloc((s->stopround + 1) * 2 * prm->t_exp_right * PLW));
    A(s->V.alloc(2 * 2 * prm->t_conv * PLW)); A(s->vW.alloc(prm->out_n * rows * prm->t_conv * PLW)); A(s->neg1.alloc(2 * s->g * PLW));
    A(s->stage.alloc((size_t)1024 * PLW)); A(s->q_wire.alloc(kWireHeaderBytes + 2 * kWireRowBytes + 8));
    A(s->cv.alloc(ncts * 2 * PLW)); A(s->c1.alloc((size_t)s->maxcnt * PLW)); A(s->ginv.alloc(expand_ginv_polys(s->plan, s->cnt.data()) * PLW));
    A(s->c0.alloc((size_t)s->maxcnt * kN));
    const size_t conv_polys = std::max(2 * nbits, s->planes_total * 2);
    A(s->conv_raw.alloc(std::max(conv_polys, (size_t)1) * kN));
    A(s->conv_ntt.alloc(std::max((size_t)2 * prm->t_conv * nbits, (prm->t_conv + 1) * s->planes_total) * PLW));
    A(s->gsw.alloc(std::max(prm->nu2, 1u) * 2 * 2 * ell * PLW));
    A(s->query.alloc(s->dim0 * 2 * kN)); A(s->scan_out.alloc(s->planes * s->local_num_per * 2 * PLW));
    A(s->cts.alloc(s->planes * s->local_num_per * 2 * kN)); A(s->result_cts.alloc(s->planes_total * 2 * kN));
    A(s->tail_cts.alloc(std::max(s->planes * (size_t)world, s->planes_total) * 2 * kN));
    if (world > 1) {
        A(s->xchg.alloc(xchg_buffer_bytes_w(world, s->slot_words))); A(s->xchg_state.alloc(2)); A(s->xchg_acks.alloc(world));
        A(s->gathered.alloc((size_t)world * s->slot_words));
        if (e == cudaSuccess) { A(cudaMemset(s->xchg.p, 0, xchg_buffer_bytes_w(world, s->slot_words))); A(cudaMemset(s->xchg_state.p, 0, 2 * sizeof(unsigned int))); }
    }
    A(s->fold_scratch.alloc(fold_scratch_words_generic(std::max(s->planes * std::max(s->local_num_per, (size_t)world), (size_t)2), 2, 1, (int)ell)));
    A(s->packed.alloc(rows * prm->out_n * PLW)); A(s->packed_raw.alloc(rows * prm->out_n * kN)); A(s->resp.alloc(rows * prm->out_n * kN));
    A(s->lists.alloc(list.size())); A(s->ct_idx_first.alloc(s->dim0)); A(s->ct_idx_direct.alloc(s->dim0));
    A(s->ct_idx_bits.alloc(nbits ? nbits : 1)); A(s->poly_idx_bits.alloc(nbits ? 2 * nbits : 1));
    if (e != cudaSuccess) { delete s; return fail(SB200_ERR_CUDA, "pack_server_create: device allocation failed: %s", cudaGetErrorString(e)); }
    A(s->lists.up(list.data(), list.size()));
    if (world > 1 && !s->shard_planes && s->stopround > 0 && s->dim0 % (2 * (size_t)world) == 0 && s->dim0 / world >= 2) {
        // even output i = 2 j'' of round r is kept when j'' = rank modulo min(world, 2^r) (the ancestors of the leaves j = rank mod
        // world); odd outputs (the GSW-bit chain) are all kept
        std::vector<int> ls;
        s->offs_sh.resize(s->g); s->cnt_sh.resize(s->g);
        for (size_t r = 0; r < s->g; r++) {
            s->offs_sh[r] = (int)ls.size();
            const size_t m = std::min((size_t)world, (size_t)1 << r);
            for (int k = 0; k < s->cnt[r]; k++) {
                const int i = list[s->offs[r] + k];
                if ((i & 1) || ((size_t)(i / 2)) % m == (size_t)rank % m) ls.push_back(i);
            }
            s->cnt_sh[r] = (int)ls.size() - s->offs_sh[r];
        }
        const size_t cl = s->dim0 / world;
        std::vector<int> cfs(cl);
        for (size_t jl = 0; jl < cl; jl++) cfs[jl] = (int)(2 * ((size_t)rank + (size_t)world * jl));
        A(s->lists_sh.alloc(ls.size())); A(s->lists_sh.up(ls.data(), ls.size()));
        A(s->ct_idx_first_s.alloc(cl)); A(s->ct_idx_first_s.up(cfs.data(), cl));
        s->expand_shard_eligible = true;
    }
    // ciphertext selections: packed query = reorientCiphertextsDim1(..., 2) and regevToSimpleGsw(..., 2, 1) (:1018, :1024);
    // direct upload = the 2^nu1 first-dimension ciphertexts as they arrive
    std::vector<int> cf(s->dim0), cd(s->dim0), cb(nbits), pb(2 * nbits);
    for (size_t j = 0; j < s->dim0; j++) { cf[j] = (int)(2 * j); cd[j] = (int)j; }
    for (size_t b = 0; b < nbits; b++) { cb[b] = (int)(2 * b + 1); pb[b] = 2 * cb[b]; pb[nbits + b] = 2 * cb[b] + 1; }
    A(s->ct_idx_first.up(cf.data(), cf.size())); A(s->ct_idx_direct.up(cd.data(), cd.size()));
    if (nbits) { A(s->ct_idx_bits.up(cb.data(), cb.size())); A(s->poly_idx_bits.up(pb.data(), pb.size())); }
    { std::vector<uint16_t> hperm(s->g * kN); build_automorph_perms(hperm.data(), (int)s->g);
      A(s->perms.alloc(hperm.size())); A(s->perms.up(hperm.data(), hperm.size())); }
    build_neg1(s->neg1.p, (int)s->g, 0);
    A(cudaStreamCreateWithFlags(&s->own_stream, cudaStreamDefault));
    A(cudaDeviceSynchronize());
    if (e != cudaSuccess) { delete s; return fail(SB200_ERR_CUDA, "pack_server_create: setup failed: %s", cudaGetErrorString(e)); }
    *out = s;
    return SB200_OK;
}
extern "C" int sb200_pack_server_create_sharded(sb200_pack_server **out, const sb200_params *prm, int device, int rank, int world) {
    return pack_server_create_impl(out, prm, device, rank, world, nullptr);
}
// plane sharding ("Pack alternative"): rank r holds the WHOLE planes p = r (mod world) and runs their scans and all
// nu2 fold rounds alone; the only exchange is one folded 32 KiB ciphertext per plane to rank 0, which packs.  The shape for
// SpiralStreamPack (cfg4: 25 planes of 8 columns - splitting 8 columns over 8 GPUs starves the scan).
extern "C" int sb200_pack_server_create_plane_sharded(sb200_pack_server **out, const sb200_params *prm, int device, int rank, int world) {
    return pack_server_create_impl(out, prm, device, rank, world, nullptr, 1);
}
extern "C" int sb200_pack_server_owns_plane(const sb200_pack_server *s, size_t plane) { return s && s->local_plane(plane) >= 0; }
extern "C" size_t sb200_pack_server_local_planes(const sb200_pack_server *s) { return s ? s->planes : 0; }
// A second query context (own keys, scratch, stream) over the parent's resident planes: one per concurrent client.
extern "C" int sb200_pack_server_create_view(sb200_pack_server **out, sb200_pack_server *parent) {
    if (!out || !parent) return fail(SB200_ERR_ARG, "pack create_view: null argument");
    const sb200_pack_server *owner = pack_owner(parent);
    return pack_server_create_impl(out, &owner->prm, owner->device, owner->rank, owner->world, owner, owner->shard_planes);
}
extern "C" int sb200_pack_server_create(sb200_pack_server **out, const sb200_params *prm, int device) {
    return sb200_pack_server_create_sharded(out, prm, device, 0, 1);
}
extern "C" void sb200_pack_server_destroy(sb200_pack_server *s) { delete s; }

// pts: this shard's items of the plane, j-major: item = j * local_num_per + ii_local (ii = rank + world * ii_local)
extern "C" int sb200_pack_server_load_plane_items(sb200_pack_server *s, size_t plane, const uint16_t *pts) {
    if (!s || !pts || plane >= s->planes_total) return fail(SB200_ERR_ARG, "load_plane_items: bad argument");
    if (s->db_owner) return fail(SB200_ERR_STATE, "this pack server is a view: load the database through its parent");
    if (s->local_plane(plane) < 0) return fail(SB200_ERR_ARG, "load_plane_items: plane %zu lives on rank %zu (plane sharding)", plane, plane % s->world);
    plane = (size_t)s->local_plane(plane);
    NEED_SCAN_PLANES(s, "pack load_plane_items");
    CU(cudaSetDevice(s->device));
    const size_t items = s->dim0 * s->local_num_per;
    DBuf<uint16_t> d(items * kN);
    CU(d.up(pts, items * kN));
    launch_db_build_pack(s->db.p + plane * s->plane_words, d.p, s->dim0, s->local_num_per, (uint32_t)s->prm.p_db, 0); CHECK_LAUNCH();
    CU(cudaDeviceSynchronize());
    s->plane_loaded[plane] = true;
    return SB200_OK;
}
// one item of a loaded plane replaced: first-dimension index j, second-dimension index ii_local inside this shard (global
// ii = rank + world * ii_local under second-dimension sharding), its polynomial as 2048 u16 coefficients < p_db
extern "C" int sb200_pack_server_set_plane_item(sb200_pack_server *s, size_t plane, size_t j, size_t ii_local, const uint16_t *poly_host) {
    if (!s || !poly_host || plane >= s->planes_total || j >= s->dim0 || ii_local >= s->local_num_per) return fail(SB200_ERR_ARG, "set_plane_item: bad argument");
    if (s->db_owner) return fail(SB200_ERR_STATE, "this pack server is a view: load the database through its parent");
    if (s->local_plane(plane) < 0) return fail(SB200_ERR_ARG, "set_plane_item: plane %zu lives on rank %zu (plane sharding)", plane, plane % s->world);
    const size_t lp = (size_t)s->local_plane(plane);
    if (!s->plane_loaded[lp]) return fail(SB200_ERR_STATE, "set_plane_item: plane %zu not loaded", plane);
    NEED_SCAN_PLANES(s, "pack set_plane_item");
    CU(cudaSetDevice(s->device));
    DBuf<uint16_t> d(kN);
    CU(d.up(poly_host, kN));
    launch_db_set_item_pack(s->db.p + lp * s->plane_words, d.p, s->dim0, s->local_num_per, (uint32_t)s->prm.p_db, ii_local, j, 0); CHECK_LAUNCH();
    CU(cudaDeviceSynchronize());
    return SB200_OK;
}
// db_buf: the WHOLE plane in the reference's convertDb layout db_buf[z][ii][j]; the shard's rows ii = rank (mod world) are taken
extern "C" int sb200_pack_server_load_plane_reference(sb200_pack_server *s, size_t plane, const uint64_t *db_buf) {
    if (!s || !db_buf || plane >= s->planes_total) return fail(SB200_ERR_ARG, "load_plane_reference: bad argument");
    if (s->db_owner) return fail(SB200_ERR_STATE, "this pack server is a view: load the database through its parent");
    if (s->local_plane(plane) < 0) return fail(SB200_ERR_ARG, "load_plane_reference: plane %zu lives on rank %zu (plane sharding)", plane, plane % s->world);
    plane = (size_t)s->local_plane(plane);
    NEED_SCAN_PLANES(s, "pack load_plane_reference");
    CU(cudaSetDevice(s->device));
    const size_t zc = 64, row = s->local_num_per * s->dim0;
    DBuf<uint64_t> stage(zc * row);
    for (size_t z0 = 0; z0 < (size_t)kN; z0 += zc) {
        if (s->world == 1 || s->shard_planes) {
            CU(stage.up(db_buf + z0 * row, zc * row));
        } else {
            for (size_t z = 0; z < zc; z++)
                CU(cudaMemcpy2D(stage.p + z * row, s->dim0 * 8, db_buf + ((z0 + z) * s->num_per + s->rank) * s->dim0,
                                (size_t)s->world * s->dim0 * 8, s->dim0 * 8, s->local_num_per, cudaMemcpyHostToDevice));
        }
        launch_db_from_reference(s->db.p + plane * s->plane_words, stage.p, s->dim0 / 2, s->local_num_per, z0, zc, 0); CHECK_LAUNCH();
        CU(cudaDeviceSynchronize());
    }
    s->plane_loaded[plane] = true;
    return SB200_OK;
}
// synthetic database (benchmarks): plaintext coefficients generated ON THE DEVICE, then the normal preprocessing
// (centre-lift, CRT, NTT, scan layout) - a 64 GiB cfg3 database is built in well under a second of GPU time
extern "C" int sb200_pack_server_load_random(sb200_pack_server *s, uint64_t seed) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->db_owner) return fail(SB200_ERR_STATE, "this pack server is a view: load the database through its parent");
    NEED_SCAN_PLANES(s, "pack load_random");
    CU(cudaSetDevice(s->device));
    const size_t items = s->dim0 * s->local_num_per, n4 = items * kN / 4;
    DBuf<uint16_t> d;
    CU(d.alloc(items * kN));
    for (size_t p = 0; p < s->planes; p++) {
        count_launch(); launch_pdl(k_fill_random_u16, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, 0, d.p, n4, (uint32_t)s->prm.p_db,
                                   seed * 0x100000001b3ull + (uint64_t)s->plane_ids[p] * 0x632be59bd9b4e019ull);
        launch_db_build_pack(s->db.p + p * s->plane_words, d.p, s->dim0, s->local_num_per, (uint32_t)s->prm.p_db, 0); CHECK_LAUNCH();
        CU(cudaDeviceSynchronize());
        s->plane_loaded[p] = true;
    }
    return SB200_OK;
}
static int pack_up(sb200_pack_server *s, DBuf<uint32_t> &dst, size_t dst_off_polys, const uint64_t *host, size_t npolys, cudaStream_t st) {
    const size_t chunk = 1024;
    for (size_t o = 0; o < npolys; o += chunk) {
        const size_t n = std::min(chunk, npolys - o);
        CU(cudaMemcpyAsync(s->stage.p, host + o * PLW, n * PLW * 8, cudaMemcpyHostToDevice, st));
        launch_ntt_u64_to_dev(dst.p + (dst_off_polys + o) * PLW, s->stage.p, n, st); CHECK_LAUNCH();
        if (o + chunk < npolys) CU(cudaStreamSynchronize(st));        // the staging buffer is reused by the next chunk
    }
    return SB200_OK;
}
// W_exp_left: g x (2 x t_exp); W_exp_right: (stopround+1) x (2 x t_exp_right); V: 2 x 2*t_conv; v_W: out_n x ((out_n+1) x t_conv).
// Expansion keys and V may be NULL for direct-upload (SpiralStreamPack) clients.
extern "C" int sb200_pack_server_set_public_params(sb200_pack_server *s, const uint64_t *W_exp_left, const uint64_t *W_exp_right,
                                                   const uint64_t *V, const uint64_t *v_W) {
    if (!s || !v_W) return fail(SB200_ERR_ARG, "pack set_public_params: null argument");
    if ((W_exp_left || W_exp_right || V) && s->dim0 + (size_t)s->prm.t_gsw * s->prm.nu2 > (size_t)kN)
        return fail(SB200_ERR_ARG, "pack set_public_params: 2^nu1 + t_GSW*nu2 exceeds the 2048 slots of a packed query (direct upload only)");
    CU(cudaSetDevice(s->device));
    cudaStream_t st = s->own_stream;
    if (W_exp_left) { TRY(pack_up(s, s->W_left, 0, W_exp_left, s->g * 2 * s->prm.t_exp, st)); CU(cudaStreamSynchronize(st)); }
    if (W_exp_right) { TRY(pack_up(s, s->W_right, 0, W_exp_right, (s->stopround + 1) * 2 * s->prm.t_exp_right, st)); CU(cudaStreamSynchronize(st)); }
    if (V) { TRY(pack_up(s, s->V, 0, V, 2 * 2 * (size_t)s->prm.t_conv, st)); CU(cudaStreamSynchronize(st)); }
    TRY(pack_up(s, s->vW, 0, v_W, (size_t)s->prm.out_n * (s->prm.out_n + 1) * s->prm.t_conv, st)); CU(cudaStreamSynchronize(st));
    s->have_params = (W_exp_left && W_exp_right && V);
    return SB200_OK;
}

// ---- staged query answering (device-resident between stages; bench.py and the multi-GPU path) -------------------
// packed single-ciphertext query (SpiralPack): H2D of the 2x1 ref-NTT ciphertext
extern "C" int sb200_pack_server_upload_query(sb200_pack_server *s, const uint64_t *query_cv_host, void *stream) {
    if (!s || !query_cv_host) return fail(SB200_ERR_ARG, "pack upload_query: null argument");
    CU(cudaMemcpyAsync(s->stage.p, query_cv_host, 2 * PLW * 8, cudaMemcpyHostToDevice, PS(s, stream)));
    s->wire_kind = 0; s->query_mode = 1;
    return SB200_OK;
}
// coefficientExpansion + reorientCiphertextsDim1 + regevToSimpleGsw (src/testing.cpp:1015-1024)
extern "C" int sb200_pack_server_expand_and_convert(sb200_pack_server *s, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (!s->have_params) return fail(SB200_ERR_STATE, "pack expand_and_convert: expansion keys / V not set");
    if (s->world > 1 && s->xchg_connected && s->expand_shard_eligible) {
        TRY(run_stage(s->g_convert_sh[s->wire_kind], PS(s, stream), s->xchg.p, nullptr, [&](cudaStream_t st) {
            if (s->wire_kind) launch_query_from_wire(s->cv.p, s->q_wire.p, s->wire_kind, st);
            else launch_ntt_u64_to_dev(s->cv.p, s->stage.p, 2, st);
            // store_self: an even output may be processed without its sibling, which would otherwise have stored the base value
            launch_expand(s->cv.p, s->plan, s->W_left.p, s->W_right.p, s->neg1.p, s->perms.p, s->c0.p, s->c1.p, s->ginv.p, s->lists_sh.p,
                          s->offs_sh.data(), s->cnt_sh.data(), st, 0, -1, -1, 1);
            // the GSW conversion first: the all-gather may wait for rank 0's acknowledgement of the previous query
            launch_regev_to_simple_gsw(s->gsw.p, nullptr, s->cv.p, s->ct_idx_bits.p, s->poly_idx_bits.p, (int)s->prm.nu2, (int)s->prm.t_gsw, s->V.p,
                                       (int)s->prm.t_conv, s->conv_raw.p, s->conv_ntt.p, st);
            launch_reorient_dim1_allgather(s->qpeers, s->cv.p, s->ct_idx_first_s.p, s->dim0, (size_t)s->rank, (size_t)s->world, s->dim0 / s->world,
                                           s->rank, s->world, s->xchg_state.p, s->xchg.p, st);
        }));
        s->query_mode = 1;
        s->query_wait_pending = true;        // the scan waits for every rank's rows of THIS query
        return SB200_OK;
    }
    GraphSlot &slot = s->wire_kind ? s->g_convert_wire[s->wire_kind - 1] : s->g_convert;
    s->query_wait_pending = false;
    return run_stage(slot, PS(s, stream), nullptr, nullptr, [&](cudaStream_t st) {
        if (s->wire_kind) launch_query_from_wire(s->cv.p, s->q_wire.p, s->wire_kind, st);
        else launch_ntt_u64_to_dev(s->cv.p, s->stage.p, 2, st);
        launch_expand(s->cv.p, s->plan, s->W_left.p, s->W_right.p, s->neg1.p, s->perms.p, s->c0.p, s->c1.p, s->ginv.p, s->lists.p,
                      s->offs.data(), s->cnt.data(), st);
        launch_reorient_dim1(s->query.p, s->cv.p, s->ct_idx_first.p, s->dim0, st);
        launch_regev_to_simple_gsw(s->gsw.p, nullptr, s->cv.p, s->ct_idx_bits.p, s->poly_idx_bits.p, (int)s->prm.nu2, (int)s->prm.t_gsw, s->V.p,
                                   (int)s->prm.t_conv, s->conv_raw.p, s->conv_ntt.p, st);
    });
}
// direct upload (SpiralStreamPack): 2^nu1 first-dimension cts + nu2 GSW cts arrive already expanded (src/testing.cpp:960-1005)
extern "C" int sb200_pack_server_upload_direct(sb200_pack_server *s, const uint64_t *v_firstdim_host, const uint64_t *v_folding_host, void *stream) {
    if (!s || !v_firstdim_host) return fail(SB200_ERR_ARG, "pack upload_direct: null argument");
    cudaStream_t st = PS(s, stream);
    const size_t ell = s->prm.t_gsw, fd = s->prm.nu2;
    if (fd && !v_folding_host) return fail(SB200_ERR_ARG, "pack upload_direct: GSW ciphertexts missing");
    TRY(pack_up(s, s->cv, 0, v_firstdim_host, s->dim0 * 2, st));
    launch_reorient_dim1(s->query.p, s->cv.p, s->ct_idx_direct.p, s->dim0, st);
    if (fd) { CU(cudaStreamSynchronize(st)); TRY(pack_up(s, s->gsw, 0, v_folding_host, fd * 2 * 2 * ell, st)); }
    CHECK_LAUNCH();
    s->query_mode = 2;
    return SB200_OK;
}
// Sharded direct upload: every rank scans with the WHOLE first-dimension query, but uploads only ITS 1/world of it - the
// ciphertexts j in [rank * 2^nu1 / world, (rank + 1) * 2^nu1 / world) - and the reorientation kernel stores that slice straight
// into every rank's query buffer over NVLink peer memory (an all-gather fused into the kernel that produces the data; each
// PCIe link carries 1/world of the 2^nu1 x 64 KiB).  v_firstdim_slice_host: this rank's ciphertexts only; v_folding_host: all
// nu2 GSW ciphertexts (small).  Needs connected peers; the scan waits for all slices (k_query_wait).
extern "C" int sb200_pack_server_upload_direct_split(sb200_pack_server *s, const uint64_t *v_firstdim_slice_host, const uint64_t *v_folding_host, void *stream) {
    if (!s || !v_firstdim_slice_host) return fail(SB200_ERR_ARG, "pack upload_direct_split: null argument");
    if (s->world < 2) return sb200_pack_server_upload_direct(s, v_firstdim_slice_host, v_folding_host, stream);
    if (!s->xchg_connected) return fail(SB200_ERR_STATE, "pack upload_direct_split: peers not connected (sb200_pack_server_xchg_connect)");
    if (s->dim0 % (2 * (size_t)s->world)) return fail(SB200_ERR_ARG, "pack upload_direct_split: 2^nu1 must be a multiple of 2 * world");
    cudaStream_t st = PS(s, stream);
    const size_t ell = s->prm.t_gsw, fd = s->prm.nu2, jc = s->dim0 / s->world;
    if (fd && !v_folding_host) return fail(SB200_ERR_ARG, "pack upload_direct_split: GSW ciphertexts missing");
    // GSW ciphertexts first: the all-gather kernel may wait for rank 0's acknowledgement of the previous query, nothing should queue behind it
    if (fd) { TRY(pack_up(s, s->gsw, 0, v_folding_host, fd * 2 * 2 * ell, st)); CU(cudaStreamSynchronize(st)); }
    TRY(pack_up(s, s->cv, 0, v_firstdim_slice_host, jc * 2, st));
    launch_reorient_dim1_allgather(s->qpeers, s->cv.p, nullptr, s->dim0, (size_t)s->rank * jc, 1, jc, s->rank, s->world, s->xchg_state.p, s->xchg.p, st);
    CHECK_LAUNCH();
    s->query_mode = 3;
    s->query_wait_pending = true;        // the next scan waits for every rank's slice of THIS upload (later scans reuse the query)
    return SB200_OK;
}
// fastMultiplyQueryByDatabaseDim1 for all out_n^2 planes of the shard in one launch (src/testing.cpp:1045-1052)
extern "C" int sb200_pack_server_scan(sb200_pack_server *s, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    for (size_t p = 0; p < s->planes; p++) if (!pack_plane_loaded(s, p)) return fail(SB200_ERR_STATE, "pack scan: database plane %zu not loaded", p);
    if (s->query_wait_pending && !tl_prepare_only) {
        launch_query_wait(s->xchg.p, s->world, s->xchg_state.p, s->xchg_state.p + 1, PS(s, stream));
        s->query_wait_pending = false;
    }
    sb200_pack_server *owner = const_cast<sb200_pack_server *>(pack_owner(s));
    if (owner->tc_only) {                               // only the limb-tile planes are resident: a one-query pass of k_scan_tc
        uint32_t *o[1] = {s->scan_out.p}; const uint64_t *q[1] = {s->query.p};
        const TcGeom g = pack_geom(owner);
        launch_queries_to_tc_g(owner->q_tc.p, q, 1, 0, owner->tc_capacity, g, PS(s, stream));
        if (launch_scan_tc_g(o, 1, owner->tc_capacity, owner->q_tc.p, owner->db_tc.p, g, owner->tc_t1.p, PS(s, stream)))
            return fail(SB200_ERR_CUDA, "pack scan (tensor-core copy only): launch failed");
    } else {
        launch_scan_pack(s->scan_out.p, s->query.p, pack_db(s), s->dim0, s->local_num_per, s->planes, s->plane_words, s->local_num_per * 2, PS(s, stream));
    }
    CHECK_LAUNCH();
    return SB200_OK;
}
// interposed fastMultiplyQueryByDatabaseDim1 on ONE resident plane: host reoriented query in, ref-NTT host ciphertexts out
extern "C" int sb200_pack_server_scan_plane_host(sb200_pack_server *s, size_t plane, const uint64_t *v_firstdim_host, uint64_t *out_ref_ntt_host) {
    if (!s || !v_firstdim_host || !out_ref_ntt_host || s->local_plane(plane) < 0) return fail(SB200_ERR_ARG, "pack scan_plane_host: bad argument");
    plane = (size_t)s->local_plane(plane);
    if (!pack_plane_loaded(s, plane)) return fail(SB200_ERR_STATE, "pack scan_plane_host: database plane %zu not loaded", plane);
    if (!pack_db(s)) return fail(SB200_ERR_STATE, "pack scan_plane_host: the scan-layout planes were released (sb200_pack_server_tc_only)");
    CU(cudaMemcpy(s->query.p, v_firstdim_host, s->dim0 * 2 * kN * sizeof(uint64_t), cudaMemcpyHostToDevice));
    launch_scan_pack(s->scan_out.p, s->query.p, pack_db(s) + plane * s->plane_words, s->dim0, s->local_num_per, 1, s->plane_words, s->local_num_per * 2, 0);
    CHECK_LAUNCH();
    return down_ntt(out_ref_ntt_host, s->scan_out.p, s->local_num_per * 2);
}
static void pack_fold_rounds(sb200_pack_server *s, uint64_t *cts, size_t count, size_t plane_stride, size_t first_round, cudaStream_t st) {
    const size_t ell = s->prm.t_gsw, fd = s->prm.nu2, gsw_polys = 2 * 2 * ell;
    size_t np = count, cur = first_round;
    while (np >= 2) {
        np /= 2;
        const size_t d = fd - 1 - cur;                        // v_folding[nu2 - 1 - cur] (foldCiphertextsDim1, :607-610)
        launch_fold_round_generic(cts, 2, 1, (int)ell, 0, np, s->planes, plane_stride, s->gsw.p + d * gsw_polys * PLW,
                                  nullptr, s->fold_scratch.p, st);            // CMux form, no negated GSW needed
        cur++;
    }
}
// ---- tensor-core batched first dimension (tc_scan.cu): fastMultiplyQueryByDatabaseDim1 of every plane for up to 16 clients
// in ONE pass over the planes.  Needs dim0 and num_per (per shard) to be multiples of 128 (SpiralPack shapes; SpiralStreamPack's
// 8-column planes cannot fill a 128-row MMA tile and keep the streaming kernel).
extern "C" int sb200_pack_server_enable_tc(sb200_pack_server *s, int capacity) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->db_owner) return fail(SB200_ERR_STATE, "pack enable_tc: call it on the server that owns the database");
    for (size_t p = 0; p < s->planes; p++) if (!s->plane_loaded[p]) return fail(SB200_ERR_STATE, "pack enable_tc: database plane %zu not loaded", p);
    if (capacity < 1 || capacity > 16) return fail(SB200_ERR_ARG, "pack enable_tc: capacity must be in [1, 16]");
    const TcGeom g = pack_geom(s);
    if (!tc_geom_ok(g)) return fail(SB200_ERR_ARG, "pack enable_tc: needs dim0 and num_per (per shard) to be multiples of 128");
    if (tc_scratch_bytes_g(g, capacity) / 4 > 0xffffffffull) return fail(SB200_ERR_ARG, "pack enable_tc: capacity too large for this shape");
    CU(cudaSetDevice(s->device));
    if (!s->db_tc.p) {
        CU(s->db_tc.alloc(tc_db_bytes_g(g)));
        launch_db_to_tc_g(s->db_tc.p, s->db.p, g, s->plane_words, s->own_stream); CHECK_LAUNCH();
    }
    if (s->q_tc.p) { cudaFree(s->q_tc.p); s->q_tc.p = nullptr; }
    if (s->tc_t1.p) { cudaFree(s->tc_t1.p); s->tc_t1.p = nullptr; }
    CU(s->tc_t1.alloc(tc_scratch_bytes_g(g, capacity) / 4));
    CU(s->q_tc.alloc(tc_query_bytes_g(g, capacity)));
    CU(cudaMemsetAsync(s->q_tc.p, 0, s->q_tc.n, s->own_stream));
    CU(cudaStreamSynchronize(s->own_stream));
    s->tc_capacity = capacity;
    return SB200_OK;
}
// Keep ONLY the limb-tile planes (as sb200_server_tc_only): SpiralPack's 64 GiB of planes then occupy 64 GiB, not 128, of the 180.
extern "C" int sb200_pack_server_tc_only(sb200_pack_server *s) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->db_owner) return fail(SB200_ERR_STATE, "pack tc_only: call it on the server that owns the database");
    if (!s->tc_capacity || !s->db_tc.p) return fail(SB200_ERR_STATE, "pack tc_only: call sb200_pack_server_enable_tc first");
    if (s->tc_only) return SB200_OK;
    CU(cudaSetDevice(s->device));
    CU(cudaDeviceSynchronize());
    cudaFree(s->db.p); s->db.p = nullptr;
    s->tc_only = true;
    return SB200_OK;
}
extern "C" int sb200_pack_server_scan_batched_tc(sb200_pack_server *const *servers, int count, void *stream) {
    if (!servers || count < 1 || !servers[0]) return fail(SB200_ERR_ARG, "pack scan_batched_tc: no servers");
    sb200_pack_server *s0 = servers[0];
    sb200_pack_server *owner = const_cast<sb200_pack_server *>(pack_owner(s0));
    if (!owner->tc_capacity) return fail(SB200_ERR_STATE, "pack scan_batched_tc: call sb200_pack_server_enable_tc on the database owner first");
    if (count > owner->tc_capacity) return fail(SB200_ERR_ARG, "pack scan_batched_tc: %d queries exceed the enabled capacity %d", count, owner->tc_capacity);
    uint32_t *o[16]; const uint64_t *qs[16];
    for (int b = 0; b < count; b++) {
        if (!servers[b] || pack_owner(servers[b]) != owner) return fail(SB200_ERR_ARG, "pack scan_batched_tc: servers must share one database");
        qs[b] = servers[b]->query.p; o[b] = servers[b]->scan_out.p;
    }
    cudaStream_t st = PS(s0, stream);
    const TcGeom g = pack_geom(owner);
    launch_queries_to_tc_g(owner->q_tc.p, qs, count, 0, owner->tc_capacity, g, st);
    if (launch_scan_tc_g(o, count, owner->tc_capacity, owner->q_tc.p, owner->db_tc.p, g, owner->tc_t1.p, st))
        return fail(SB200_ERR_CUDA, "pack scan_batched_tc: launch failed");
    CHECK_LAUNCH();
    return SB200_OK;
}
// `count` reoriented queries (reorientCiphertextsDim1 layout) against ONE plane in the reference's convertDb layout, one pass
extern "C" int sb200_fastMultiplyQueryByDatabaseDim1_batched(uint64_t *const *out, const uint64_t *db, const uint64_t *const *v_firstdim,
                                                             int count, size_t dim0, size_t num_per) {
    NEED_DEVICE();
    if (!out || !db || !v_firstdim || count < 1 || count > 16) return fail(SB200_ERR_ARG, "fastMultiplyQueryByDatabaseDim1_batched: 1 <= count <= 16");
    const TcGeom g = tc_geom_pack(dim0, num_per, 1, num_per * 2);
    if (!tc_geom_ok(g)) return fail(SB200_ERR_ARG, "fastMultiplyQueryByDatabaseDim1_batched: needs dim0 and num_per to be multiples of 128");
    const size_t words = dim0 * num_per * kN, qwords = dim0 * 2 * kN;
    DBuf<uint64_t> dq((size_t)count * qwords), dref(words), ddb(words); DBuf<uint8_t> dtc(tc_db_bytes_g(g)), qtc(tc_query_bytes_g(g, count));
    DBuf<uint32_t> dout((size_t)count * num_per * 2 * PLW), dt1(tc_scratch_bytes_g(g, count) / 4);
    CU(dref.up(db, words));
    launch_db_from_reference(ddb.p, dref.p, dim0 / 2, num_per, 0, kN, 0); CHECK_LAUNCH();
    launch_db_to_tc_g(dtc.p, ddb.p, g, words, 0); CHECK_LAUNCH();
    CU(cudaMemset(qtc.p, 0, qtc.n));
    uint32_t *o[16]; const uint64_t *qs[16];
    for (int b = 0; b < count; b++) {
        CU(cudaMemcpy(dq.p + (size_t)b * qwords, v_firstdim[b], qwords * 8, cudaMemcpyHostToDevice));
        qs[b] = dq.p + (size_t)b * qwords; o[b] = dout.p + (size_t)b * num_per * 2 * PLW;
    }
    launch_queries_to_tc_g(qtc.p, qs, count, 0, count, g, 0); CHECK_LAUNCH();
    if (launch_scan_tc_g(o, count, count, qtc.p, dtc.p, g, dt1.p, 0)) return fail(SB200_ERR_CUDA, "fastMultiplyQueryByDatabaseDim1_batched: launch failed");
    CHECK_LAUNCH();
    for (int b = 0; b < count; b++) TRY(down_ntt(out[b], o[b], num_per * 2));
    return SB200_OK;
}
// from_ntt of every scan output + the local fold rounds (src/testing.cpp:1055-1058); leaves `planes` surviving cts
extern "C" int sb200_pack_server_fold_local(sb200_pack_server *s, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    return run_stage(s->g_fold, PS(s, stream), nullptr, nullptr, [&](cudaStream_t st) {
        launch_from_ntt(s->cts.p, s->scan_out.p, s->planes * s->local_num_per * 2, st);
        pack_fold_rounds(s, s->cts.p, s->local_num_per, s->local_num_per, 0, st);
        cudaMemcpy2DAsync(s->result_cts.p, 2 * kN * 8, s->cts.p, s->local_num_per * 2 * kN * 8, 2 * kN * 8, s->planes, cudaMemcpyDeviceToDevice, st);
    });
}
extern "C" uint64_t *sb200_pack_server_partial_cts(sb200_pack_server *s) { return s ? s->result_cts.p : nullptr; }
extern "C" size_t sb200_pack_server_partial_words(const sb200_pack_server *s) { return s ? s->planes * 2 * kN : 0; }
extern "C" int sb200_pack_server_copy_partial(sb200_pack_server *s, uint64_t *dst_dev, void *stream) {
    if (!s || !dst_dev) return fail(SB200_ERR_ARG, "pack copy_partial: null argument");
    CU(cudaMemcpyAsync(dst_dev, s->result_cts.p, s->planes * 2 * kN * 8, cudaMemcpyDeviceToDevice, PS(s, stream)));
    return SB200_OK;
}
// rank 0: gathered = [world][planes] surviving cts (device, rank order; world == 1: the server's own partial buffer).
// Last log2(world) fold rounds, pack (:1066-1072), modulus switch (:1074-1081) -> total_resp_dev ((out_n+1) x out_n raw).
// gathered: [world][slot_words] (rank order).  Second-dimension sharding: slot = planes_total ciphertexts, the last log2(world)
// fold rounds run here; plane sharding: slot r = rank r's planes in local order, they are only put back into plane order.
static void pack_tail_body(sb200_pack_server *s, const uint64_t *gathered_dev, uint64_t *total_resp_dev, cudaStream_t st) {
    const size_t out_n = s->prm.out_n, rows = out_n + 1;
    const uint64_t *final_cts = gathered_dev;
    if (s->world > 1 && s->shard_planes) {
        count_launch(); launch_pdl(k_gather_planes, dim3((unsigned)s->planes_total), dim3(256), 0, st, s->tail_cts.p, gathered_dev, s->world, s->slot_words);
        cudaMemcpyAsync(s->result_cts.p, s->tail_cts.p, s->planes_total * 2 * kN * 8, cudaMemcpyDeviceToDevice, st);
        final_cts = s->result_cts.p;
    } else if (s->world > 1) {
        count_launch(); launch_pdl(k_transpose_cts, dim3((unsigned)(s->world * s->planes)), dim3(256), 0, st, s->tail_cts.p, gathered_dev, s->world, (int)s->planes);
        pack_fold_rounds(s, s->tail_cts.p, (size_t)s->world, (size_t)s->world, s->prm.nu2 - s->log_world, st);
        cudaMemcpy2DAsync(s->result_cts.p, 2 * kN * 8, s->tail_cts.p, (size_t)s->world * 2 * kN * 8, 2 * kN * 8, s->planes, cudaMemcpyDeviceToDevice, st);
        final_cts = s->result_cts.p;
    }
    launch_pack(s->packed.p, final_cts, s->vW.p, (int)out_n, (int)s->prm.t_conv, s->conv_raw.p, s->conv_ntt.p, st);
    launch_from_ntt(s->packed_raw.p, s->packed.p, rows * out_n, st);
    launch_rescale2(total_resp_dev, s->packed_raw.p, out_n * (size_t)kN, (rows - 1) * out_n * (size_t)kN, kQ, sb200_arb_qprime(s->prm.qp_bits), 4 * s->prm.p_db, st);
}
extern "C" int sb200_pack_server_fold_tail(sb200_pack_server *s, const uint64_t *gathered_dev, uint64_t *total_resp_dev, void *stream) {
    if (!s || !gathered_dev || !total_resp_dev) return fail(SB200_ERR_ARG, "pack fold_tail: null argument");
    return run_stage(s->g_tail, PS(s, stream), gathered_dev, total_resp_dev, [&](cudaStream_t st) { pack_tail_body(s, gathered_dev, total_resp_dev, st); });
}
// ---- peer-memory exchange for the Pack servers (same protocol and kernels as sb200_server_xchg_*, payload = this rank's surviving
// ciphertexts: planes_total x 32 KiB under second-dimension sharding, its own planes under plane sharding)
extern "C" size_t sb200_pack_server_xchg_handle_bytes(void) { return 2 * sizeof(cudaIpcMemHandle_t); }
extern "C" int sb200_pack_server_xchg_export(sb200_pack_server *s, void *handle_out) {
    if (!s || !handle_out || s->world < 2) return fail(SB200_ERR_ARG, "pack xchg_export: needs a sharded server");
    cudaIpcMemHandle_t h[2];
    CU(cudaIpcGetMemHandle(&h[0], s->xchg.p));
    CU(cudaIpcGetMemHandle(&h[1], s->query.p));
    memcpy(handle_out, h, sizeof h);
    return SB200_OK;
}
static int pack_xchg_finish(sb200_pack_server *s, const std::vector<void *> &bufs, const std::vector<void *> &queries) {
    preload_exchange_kernels(); preload_kernel(k_gather_planes); preload_kernel(k_transpose_cts);
    s->xchg_target = bufs[0];
    for (int r = 0; r < s->world; r++) { s->qpeers.query[r] = (uint64_t *)queries[r]; s->qpeers.xb[r] = (XchgBuf *)bufs[r]; }
    if (s->rank == 0) {
        std::vector<unsigned int *> acks(s->world);
        for (int r = 0; r < s->world; r++) acks[r] = reinterpret_cast<unsigned int *>(reinterpret_cast<uint8_t *>(bufs[r]) + xchg_ack_offset());
        CU(s->xchg_acks.up(acks.data(), s->world));
    }
    s->xchg_connected = true;
    return SB200_OK;
}
extern "C" int sb200_pack_server_xchg_connect(sb200_pack_server *s, const void *all_handles) {
    if (!s || !all_handles || s->world < 2) return fail(SB200_ERR_ARG, "pack xchg_connect: needs a sharded server");
    std::vector<void *> bufs(s->world, nullptr), queries(s->world, nullptr);
    for (int r = 0; r < s->world; r++) {
        if (r == s->rank) { bufs[r] = s->xchg.p; queries[r] = s->query.p; continue; }
        cudaIpcMemHandle_t h[2];
        memcpy(h, reinterpret_cast<const uint8_t *>(all_handles) + (size_t)r * sizeof h, sizeof h);
        void *pb = nullptr, *pq = nullptr;
        CU(cudaIpcOpenMemHandle(&pb, h[0], cudaIpcMemLazyEnablePeerAccess)); s->ipc_opened.push_back(pb);
        CU(cudaIpcOpenMemHandle(&pq, h[1], cudaIpcMemLazyEnablePeerAccess)); s->ipc_opened.push_back(pq);
        bufs[r] = pb; queries[r] = pq;
    }
    return pack_xchg_finish(s, bufs, queries);
}
extern "C" int sb200_pack_server_xchg_connect_local(sb200_pack_server *s, sb200_pack_server *const *all) {
    if (!s || !all || s->world < 2) return fail(SB200_ERR_ARG, "pack xchg_connect_local: needs a sharded server");
    std::vector<void *> bufs(s->world, nullptr), queries(s->world, nullptr);
    for (int r = 0; r < s->world; r++) {
        if (!all[r] || all[r]->world != s->world || all[r]->rank != r || all[r]->shard_planes != s->shard_planes) return fail(SB200_ERR_ARG, "pack xchg_connect_local: server %d mismatched", r);
        bufs[r] = all[r]->xchg.p; queries[r] = all[r]->query.p;
        if (all[r]->device != s->device) {
            int can = 0;
            CU(cudaDeviceCanAccessPeer(&can, s->device, all[r]->device));
            if (!can) return fail(SB200_ERR_CUDA, "pack xchg_connect_local: no peer access %d -> %d", s->device, all[r]->device);
            cudaError_t pe = cudaDeviceEnablePeerAccess(all[r]->device, 0);
            if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) CU(pe);
            cudaGetLastError();
        }
    }
    return pack_xchg_finish(s, bufs, queries);
}
// every rank: push the surviving ciphertexts into rank 0's HBM; rank 0 also waits for all ranks and runs the tail (the last fold
// rounds under second-dimension sharding, then pack + modulus switch) into resp_dev (ignored on the other ranks)
extern "C" int sb200_pack_server_exchange_and_tail(sb200_pack_server *s, uint64_t *resp_dev, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->world < 2) return sb200_pack_server_fold_tail(s, s->result_cts.p, resp_dev ? resp_dev : s->resp.p, stream);
    if (!s->xchg_connected) return fail(SB200_ERR_STATE, "pack exchange_and_tail: peers not connected (sb200_pack_server_xchg_connect)");
    if (s->rank == 0 && !resp_dev) resp_dev = s->resp.p;
    return run_stage(s->g_xchg, PS(s, stream), resp_dev, nullptr, [&](cudaStream_t st) {
        launch_xchg_push_w(s->xchg_target, s->xchg.p, s->result_cts.p, s->xchg_state.p, s->rank, s->world, s->xchg_state.p + 1,
                           s->planes * 2 * (size_t)kN, s->slot_words, st);
        if (s->rank == 0) {
            launch_xchg_wait_w(s->xchg.p, s->xchg_acks.p, s->gathered.p, s->xchg_state.p, s->world, s->xchg_state.p + 1, s->slot_words, st);
            pack_tail_body(s, s->gathered.p, resp_dev, st);
        }
    });
}
extern "C" int sb200_pack_server_xchg_error(sb200_pack_server *s, void *stream) {
    if (!s || s->world < 2) return 0;
    unsigned int st[2] = {0, 0};
    if (cudaStreamSynchronize(PS(s, stream)) != cudaSuccess) return -1;
    if (cudaMemcpy(st, s->xchg_state.p, sizeof st, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int)st[1];
}
// all server stages of the query last uploaded (sb200_pack_server_upload_query / _upload_direct / _upload_direct_split) in ONE
// call, sharded servers included (every rank calls it; the response lands on rank 0).  marks: NULL or four cudaEvent_t recorded
// before the expansion, before and after the first-dimension scan and at the end.
// Capture and instantiate the graphs sb200_pack_server_process will replay, without running anything (as sb200_server_prepare):
// several shards driven from one process on one device must not build graphs while another shard's kernel spins on a flag.
extern "C" int sb200_pack_server_prepare(sb200_pack_server *s, uint64_t *total_resp_dev, void *stream) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (s->world > 1 && !s->xchg_connected) return fail(SB200_ERR_STATE, "pack prepare: sharded server without connected peers");
    CU(cudaSetDevice(s->device));
    const int mode = s->query_mode; const bool pending = s->query_wait_pending;
    tl_prepare_only = true;
    int rc = s->have_params ? sb200_pack_server_expand_and_convert(s, stream) : SB200_OK;
    if (!rc) rc = sb200_pack_server_fold_local(s, stream);
    if (!rc) rc = sb200_pack_server_exchange_and_tail(s, total_resp_dev, stream);
    tl_prepare_only = false;
    s->query_mode = mode; s->query_wait_pending = pending;
    CU(cudaStreamSynchronize(PS(s, stream)));
    return rc;
}
extern "C" int sb200_pack_server_expansion_sharded(const sb200_pack_server *s) { return s && s->world > 1 && s->xchg_connected && s->expand_shard_eligible && s->have_params; }
extern "C" int sb200_pack_server_process(sb200_pack_server *s, uint64_t *total_resp_dev, void *stream, void *const *marks) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    if (!s->query_mode) return fail(SB200_ERR_STATE, "pack process: no query uploaded");
    if (s->world > 1 && !s->xchg_connected) return fail(SB200_ERR_STATE, "pack process: sharded server without connected peers (sb200_pack_server_xchg_connect)");
    cudaStream_t st = PS(s, stream);
    if (marks) CU(cudaEventRecord((cudaEvent_t)marks[0], st));
    if (s->query_mode == 1) TRY(sb200_pack_server_expand_and_convert(s, stream));
    if (marks) CU(cudaEventRecord((cudaEvent_t)marks[1], st));
    TRY(sb200_pack_server_scan(s, stream));
    if (marks) CU(cudaEventRecord((cudaEvent_t)marks[2], st));
    TRY(sb200_pack_server_fold_local(s, stream));
    TRY(sb200_pack_server_exchange_and_tail(s, total_resp_dev, stream));
    if (marks) CU(cudaEventRecord((cudaEvent_t)marks[3], st));
    return SB200_OK;
}
extern "C" uint64_t *sb200_pack_server_result_cts(sb200_pack_server *s) { return s ? s->result_cts.p : nullptr; }
extern "C" uint64_t *sb200_pack_server_response_ptr(sb200_pack_server *s) { return s ? s->resp.p : nullptr; }
extern "C" int sb200_pack_server_download(sb200_pack_server *s, uint64_t *dst_host, const uint64_t *src_dev, size_t words, void *stream) {
    if (!s || !dst_host || !src_dev) return fail(SB200_ERR_ARG, "pack download: null argument");
    CU(cudaMemcpyAsync(dst_host, src_dev, words * 8, cudaMemcpyDeviceToHost, PS(s, stream)));
    CU(cudaStreamSynchronize(PS(s, stream)));
    return SB200_OK;
}

static int pack_process(sb200_pack_server *s, uint64_t *resp_host, uint64_t *result_cts_host, void *stream) {
    if (s->world != 1) return fail(SB200_ERR_STATE, "pack answer: single-shard call on a sharded server (use the staged API)");
    const size_t out_n = s->prm.out_n, rows = out_n + 1;
    TRY(sb200_pack_server_scan(s, stream));
    TRY(sb200_pack_server_fold_local(s, stream));
    TRY(sb200_pack_server_fold_tail(s, s->result_cts.p, s->resp.p, stream));
    if (result_cts_host) CU(cudaMemcpyAsync(result_cts_host, s->result_cts.p, s->planes_total * 2 * kN * 8, cudaMemcpyDeviceToHost, PS(s, stream)));
    return sb200_pack_server_download(s, resp_host, s->resp.p, rows * out_n * kN, stream);
}
// packed single-ciphertext query (SpiralPack): expansion + conversion + processing; 64 KiB in, (out_n+1) x out_n x 16 KiB out
extern "C" int sb200_pack_server_answer(sb200_pack_server *s, const uint64_t *query_cv_host, uint64_t *total_resp_host,
                                        uint64_t *result_cts_host, void *stream) {
    if (!s || !query_cv_host || !total_resp_host) return fail(SB200_ERR_ARG, "pack answer: null argument");
    TRY(sb200_pack_server_upload_query(s, query_cv_host, stream));
    TRY(sb200_pack_server_expand_and_convert(s, stream));
    return pack_process(s, total_resp_host, result_cts_host, stream);
}
extern "C" int sb200_pack_server_answer_direct(sb200_pack_server *s, const uint64_t *v_firstdim_host, const uint64_t *v_folding_host,
                                               uint64_t *total_resp_host, uint64_t *result_cts_host, void *stream) {
    if (!s || !v_firstdim_host || !total_resp_host) return fail(SB200_ERR_ARG, "pack answer_direct: null argument");
    TRY(sb200_pack_server_upload_direct(s, v_firstdim_host, v_folding_host, stream));
    return pack_process(s, total_resp_host, result_cts_host, stream);
}
extern "C" size_t sb200_pack_server_db_bytes(const sb200_pack_server *s) { return s ? s->planes * s->plane_words * 8 : 0; }
extern "C" size_t sb200_pack_server_response_words(const sb200_pack_server *s) { return s ? (size_t)(s->prm.out_n + 1) * s->prm.out_n * kN : 0; }
