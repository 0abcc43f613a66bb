// api_pack.cu - C-ABI entry points of the SpiralPack / SpiralStreamPack path (include/spiral_b200.h).
// Included by spiral_b200.cu after api.cu (shares its helpers: DBuf, up_ntt, down_ntt, CU, TRY, fail).

namespace sb200 {
void launch_db_build_pack(uint64_t *db_plane, const uint16_t *pts_plane, size_t dim0, size_t num_per, uint32_t p_db, cudaStream_t s);
void launch_reorient_dim1(uint64_t *out, const uint32_t *cv, const int *ct_idx, size_t dim0, cudaStream_t s);
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
// ---------------------------------------------------------------------------------------------
struct sb200_pack_server {
    sb200_params prm;
    int device = 0;
    size_t dim0 = 0, num_per = 0, planes = 0, plane_words = 0;
    size_t g = 0, stopround = 0;
    ExpandPlan plan{};
    std::vector<int> offs, cnt;
    int maxcnt = 0, tmax = 0;
    bool have_params = false;
    std::vector<bool> plane_loaded;
    DBuf<uint64_t> db;
    DBuf<uint32_t> W_left, W_right, V, vW, neg1;
    DBuf<uint64_t> stage;
    DBuf<uint32_t> cv, c1, ginv, conv_ntt, gsw, scan_out, fold_scratch, packed;
    DBuf<uint64_t> c0, conv_raw, query, cts, result_cts, packed_raw, resp;
    DBuf<int> lists, ct_idx_first, ct_idx_bits, poly_idx_bits;
    DBuf<uint16_t> perms;
};

extern "C" int sb200_pack_server_create(sb200_pack_server **out, const sb200_params *prm, int device) {
    if (!out || !prm) return fail(SB200_ERR_ARG, "pack_server_create: null argument");
    if (prm->out_n == 0 || prm->nu1 < 1) return fail(SB200_ERR_ARG, "pack_server_create: out_n >= 1 and nu1 >= 1 required");
    int rc = sb200_init(device);
    if (rc) return rc;
    sb200_pack_server *s = new sb200_pack_server();
    s->prm = *prm; s->device = device;
    s->dim0 = (size_t)1 << prm->nu1; s->num_per = (size_t)1 << prm->nu2; s->planes = (size_t)prm->out_n * prm->out_n;
    s->plane_words = s->dim0 * s->num_per * kN;
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
    A(s->db.alloc(s->planes * s->plane_words));
    A(s->W_left.alloc(s->g * 2 * prm->t_exp * PLW)); A(s->W_right.alloc((s->stopround + 1) * 2 * prm->t_exp_right * PLW));
    A(s->V.alloc(2 * 2 * prm->t_conv * PLW)); A(s->vW.alloc(prm->out_n * rows * prm->t_conv * PLW)); A(s->neg1.alloc(s->g * PLW));
    A(s->stage.alloc(std::max((size_t)2, (size_t)1024) * PLW));
    A(s->cv.alloc(ncts * 2 * PLW)); A(s->c1.alloc((size_t)s->maxcnt * PLW)); A(s->ginv.alloc((size_t)s->maxcnt * s->tmax * PLW));
    A(s->c0.alloc((size_t)s->maxcnt * kN));
    const size_t conv_polys = std::max(2 * nbits, s->planes * 2);
    A(s->conv_raw.alloc(std::max(conv_polys, (size_t)1) * kN));
    A(s->conv_ntt.alloc(std::max((size_t)2 * prm->t_conv * nbits, (prm->t_conv + 1) * s->planes) * PLW));
    A(s->gsw.alloc(std::max(prm->nu2, 1u) * 2 * 2 * ell * PLW));
    A(s->query.alloc(s->dim0 * 2 * kN)); A(s->scan_out.alloc(s->planes * s->num_per * 2 * PLW));
    A(s->cts.alloc(s->planes * s->num_per * 2 * kN)); A(s->result_cts.alloc(s->planes * 2 * kN));
    A(s->fold_scratch.alloc(fold_scratch_words_generic(std::max(s->planes * s->num_per, (size_t)2), 2, 1, (int)ell)));
    A(s->packed.alloc(rows * prm->out_n * PLW)); A(s->packed_raw.alloc(rows * prm->out_n * kN)); A(s->resp.alloc(rows * prm->out_n * kN));
    A(s->lists.alloc(list.size())); A(s->ct_idx_first.alloc(s->dim0)); A(s->ct_idx_bits.alloc(nbits ? nbits : 1)); A(s->poly_idx_bits.alloc(nbits ? 2 * nbits : 1));
    if (e != cudaSuccess) { delete s; return fail(SB200_ERR_CUDA, "pack_server_create: device allocation failed: %s", cudaGetErrorString(e)); }
    A(s->lists.up(list.data(), list.size()));
    { std::vector<uint16_t> hperm(s->g * kN); build_automorph_perms(hperm.data(), (int)s->g);
      A(s->perms.alloc(hperm.size())); A(s->perms.up(hperm.data(), hperm.size())); }
    build_neg1(s->neg1.p, (int)s->g, 0);
    A(cudaDeviceSynchronize());
    if (e != cudaSuccess) { delete s; return fail(SB200_ERR_CUDA, "pack_server_create: setup failed: %s", cudaGetErrorString(e)); }
    *out = s;
    return SB200_OK;
}
extern "C" void sb200_pack_server_destroy(sb200_pack_server *s) { delete s; }

extern "C" int sb200_pack_server_load_plane_items(sb200_pack_server *s, size_t plane, const uint16_t *pts) {
    if (!s || !pts || plane >= s->planes) return fail(SB200_ERR_ARG, "load_plane_items: bad argument");
    if (s->num_per < 2) return fail(SB200_ERR_ARG, "load_plane_items: num_per >= 2 required (use load_plane_reference)");
    const size_t items = s->dim0 * s->num_per;
    DBuf<uint16_t> d(items * kN);
    CU(d.up(pts, items * kN));
    launch_db_build_pack(s->db.p + plane * s->plane_words, d.p, s->dim0, s->num_per, (uint32_t)s->prm.p_db, 0); CHECK_LAUNCH();
    CU(cudaDeviceSynchronize());
    s->plane_loaded[plane] = true;
    return SB200_OK;
}
extern "C" int sb200_pack_server_load_plane_reference(sb200_pack_server *s, size_t plane, const uint64_t *db_buf) {
    if (!s || !db_buf || plane >= s->planes) return fail(SB200_ERR_ARG, "load_plane_reference: bad argument");
    const size_t zc = 64, row = s->num_per * s->dim0;
    DBuf<uint64_t> stage(zc * row);
    for (size_t z0 = 0; z0 < (size_t)kN; z0 += zc) {
        CU(stage.up(db_buf + z0 * row, zc * row));
        launch_db_from_reference(s->db.p + plane * s->plane_words, stage.p, s->dim0 / 2, s->num_per, z0, zc, 0); CHECK_LAUNCH();
        CU(cudaDeviceSynchronize());
    }
    s->plane_loaded[plane] = true;
    return SB200_OK;
}
extern "C" int sb200_pack_server_load_random(sb200_pack_server *s, uint64_t seed) {
    if (!s) return fail(SB200_ERR_ARG, "null server");
    const size_t items = s->dim0 * s->num_per;
    std::vector<uint16_t> h(items * kN);
    uint64_t x = seed * 0x9e3779b97f4a7c15ull + 99;
    for (size_t p = 0; p < s->planes; p++) {
        for (size_t i = 0; i < h.size(); i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; h[i] = (uint16_t)((x >> 20) % s->prm.p_db); }
        TRY(sb200_pack_server_load_plane_items(s, p, h.data()));
    }
    return SB200_OK;
}
static int pack_up(sb200_pack_server *s, DBuf<uint32_t> &dst, size_t dst_off_polys, const uint64_t *host, size_t npolys, cudaStream_t st) {
    const size_t chunk = 1024;
    for (size_t o = 0; o < npolys; o += chunk) {
        const size_t n = std::min(chunk, npolys - o);
        CU(cudaMemcpyAsync(s->stage.p, host + o * PLW, n * PLW * 8, cudaMemcpyHostToDevice, st));
        launch_ntt_u64_to_dev(dst.p + (dst_off_polys + o) * PLW, s->stage.p, n, st); CHECK_LAUNCH();
        CU(cudaStreamSynchronize(st));
    }
    return SB200_OK;
}
// W_exp_left: g x (2 x t_exp); W_exp_right: (stopround+1) x (2 x t_exp_right); V: 2 x 2*t_conv; v_W: out_n x ((out_n+1) x t_conv).
// Expansion keys and V may be NULL for direct-upload (SpiralStreamPack) clients.
extern "C" int sb200_pack_server_set_public_params(sb200_pack_server *s, const uint64_t *W_exp_left, const uint64_t *W_exp_right,
                                                   const uint64_t *V, const uint64_t *v_W) {
    if (!s || !v_W) return fail(SB200_ERR_ARG, "pack set_public_params: null argument");
    if (W_exp_left) TRY(pack_up(s, s->W_left, 0, W_exp_left, s->g * 2 * s->prm.t_exp, 0));
    if (W_exp_right) TRY(pack_up(s, s->W_right, 0, W_exp_right, (s->stopround + 1) * 2 * s->prm.t_exp_right, 0));
    if (V) TRY(pack_up(s, s->V, 0, V, 2 * 2 * (size_t)s->prm.t_conv, 0));
    TRY(pack_up(s, s->vW, 0, v_W, (size_t)s->prm.out_n * (s->prm.out_n + 1) * s->prm.t_conv, 0));
    s->have_params = (W_exp_left && W_exp_right && V);
    return SB200_OK;
}

static int pack_process(sb200_pack_server *s, uint64_t *resp_host, uint64_t *result_cts_host, cudaStream_t st) {
    for (size_t p = 0; p < s->planes; p++) if (!s->plane_loaded[p]) return fail(SB200_ERR_STATE, "pack answer: database plane %zu not loaded", p);
    const size_t ell = s->prm.t_gsw, fd = s->prm.nu2, gsw_polys = 2 * 2 * ell, out_n = s->prm.out_n, rows = out_n + 1;
    // first dimension for all planes at once, lift, fold (planes batched), keep ct 0 of every plane
    launch_scan_pack(s->scan_out.p, s->query.p, s->db.p, s->dim0, s->num_per, s->planes, s->plane_words, s->num_per * 2, st);
    launch_from_ntt(s->cts.p, s->scan_out.p, s->planes * s->num_per * 2, st);
    size_t np = s->num_per;
    for (size_t cur = 0; cur < fd; cur++) {
        np /= 2;
        const size_t d = fd - 1 - cur;
        launch_fold_round_generic(s->cts.p, 2, 1, (int)ell, 0, np, s->planes, s->num_per, s->gsw.p + d * gsw_polys * PLW,
                                  nullptr, s->fold_scratch.p, st);            // CMux form, no negated GSW needed
    }
    CU(cudaMemcpy2DAsync(s->result_cts.p, 2 * kN * 8, s->cts.p, s->num_per * 2 * kN * 8, 2 * kN * 8, s->planes, cudaMemcpyDeviceToDevice, st));
    launch_pack(s->packed.p, s->result_cts.p, s->vW.p, (int)out_n, (int)s->prm.t_conv, s->conv_raw.p, s->conv_ntt.p, st);
    launch_from_ntt(s->packed_raw.p, s->packed.p, rows * out_n, st);
    launch_rescale(s->resp.p, s->packed_raw.p, out_n * (size_t)kN, kQ, sb200_arb_qprime(s->prm.qp_bits), st);
    launch_rescale(s->resp.p + out_n * kN, s->packed_raw.p + out_n * kN, (rows - 1) * out_n * (size_t)kN, kQ, 4 * s->prm.p_db, st);
    CHECK_LAUNCH();
    if (result_cts_host) CU(cudaMemcpyAsync(result_cts_host, s->result_cts.p, s->planes * 2 * kN * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(resp_host, s->resp.p, rows * out_n * kN * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SB200_OK;
}
// packed single-ciphertext query (SpiralPack): expansion + conversion + processing
extern "C" int sb200_pack_server_answer(sb200_pack_server *s, const uint64_t *query_cv_host, uint64_t *total_resp_host,
                                        uint64_t *result_cts_host, void *stream) {
    if (!s || !query_cv_host || !total_resp_host) return fail(SB200_ERR_ARG, "pack answer: null argument");
    if (!s->have_params) return fail(SB200_ERR_STATE, "pack answer: expansion keys / V not set");
    cudaStream_t st = S(stream);
    const size_t ell = s->prm.t_gsw, nbits = ell * s->prm.nu2;
    std::vector<int> cf(s->dim0), cb(nbits), pb(2 * nbits);
    for (size_t j = 0; j < s->dim0; j++) cf[j] = (int)(2 * j);                              // reorientCiphertextsDim1(..., 2)
    for (size_t b = 0; b < nbits; b++) { cb[b] = (int)(2 * b + 1); pb[b] = 2 * cb[b]; pb[nbits + b] = 2 * cb[b] + 1; }   // regevToSimpleGsw(..., 2, 1)
    CU(cudaMemcpyAsync(s->ct_idx_first.p, cf.data(), cf.size() * 4, cudaMemcpyHostToDevice, st));
    if (nbits) { CU(cudaMemcpyAsync(s->ct_idx_bits.p, cb.data(), cb.size() * 4, cudaMemcpyHostToDevice, st));
                 CU(cudaMemcpyAsync(s->poly_idx_bits.p, pb.data(), pb.size() * 4, cudaMemcpyHostToDevice, st)); }
    CU(cudaMemcpyAsync(s->stage.p, query_cv_host, 2 * PLW * 8, cudaMemcpyHostToDevice, st));
    launch_ntt_u64_to_dev(s->cv.p, s->stage.p, 2, st);
    launch_expand(s->cv.p, s->plan, s->W_left.p, s->W_right.p, s->neg1.p, s->perms.p, s->c0.p, s->c1.p, s->ginv.p, s->lists.p, s->offs.data(), s->cnt.data(), st);
    launch_reorient_dim1(s->query.p, s->cv.p, s->ct_idx_first.p, s->dim0, st);
    launch_regev_to_simple_gsw(s->gsw.p, nullptr, s->cv.p, s->ct_idx_bits.p, s->poly_idx_bits.p, (int)s->prm.nu2, (int)ell, s->V.p,
                               (int)s->prm.t_conv, s->conv_raw.p, s->conv_ntt.p, st);
    CHECK_LAUNCH();
    CU(cudaStreamSynchronize(st));          // host index vectors go out of scope
    return pack_process(s, total_resp_host, result_cts_host, st);
}
// direct upload (SpiralStreamPack): 2^nu1 first-dimension cts + nu2 GSW cts arrive already expanded
extern "C" int sb200_pack_server_answer_direct(sb200_pack_server *s, const uint64_t *v_firstdim_host, const uint64_t *v_folding_host,
                                               uint64_t *total_resp_host, uint64_t *result_cts_host, void *stream) {
    if (!s || !v_firstdim_host || !total_resp_host) return fail(SB200_ERR_ARG, "pack answer_direct: null argument");
    cudaStream_t st = S(stream);
    const size_t ell = s->prm.t_gsw, fd = s->prm.nu2;
    TRY(pack_up(s, s->cv, 0, v_firstdim_host, s->dim0 * 2, st));
    std::vector<int> cf(s->dim0);
    for (size_t j = 0; j < s->dim0; j++) cf[j] = (int)j;
    CU(cudaMemcpyAsync(s->ct_idx_first.p, cf.data(), cf.size() * 4, cudaMemcpyHostToDevice, st));
    launch_reorient_dim1(s->query.p, s->cv.p, s->ct_idx_first.p, s->dim0, st);
    if (fd) {
        if (!v_folding_host) return fail(SB200_ERR_ARG, "pack answer_direct: GSW ciphertexts missing");
        TRY(pack_up(s, s->gsw, 0, v_folding_host, fd * 2 * 2 * ell, st));
    }
    CHECK_LAUNCH();
    CU(cudaStreamSynchronize(st));
    return pack_process(s, total_resp_host, result_cts_host, st);
}
extern "C" size_t sb200_pack_server_db_bytes(const sb200_pack_server *s) { return s ? s->planes * s->plane_words * 8 : 0; }
extern "C" size_t sb200_pack_server_response_words(const sb200_pack_server *s) { return s ? (size_t)(s->prm.out_n + 1) * s->prm.out_n * kN : 0; }
