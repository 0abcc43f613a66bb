"""GPU suite: SpiralPack / SpiralStreamPack resident server against the oracle's whole Pack pipeline
(so_pack_answer = testHighRate's server statements).  Inputs are uniform ring elements of the right
shapes: the arithmetic is exact, so equality on arbitrary inputs is the strongest parity statement;
both database ingest paths and both query modes (packed + expansion, direct upload) are covered."""
import ctypes as C

import numpy as np
import pytest

from spiral_b200 import SpiralParams
from spiral_b200.lib import check
from tests import oracle_lib as ol

pytestmark = pytest.mark.gpu
N, PL = ol.N, 2 * ol.N
P64 = C.POINTER(C.c_uint64)
P16 = C.POINTER(C.c_uint16)


def p(a):
    return a.ctypes.data_as(P64)


def rnd_ntt(rng, npolys):
    a = np.empty((npolys, 2, N), dtype=np.uint64)
    a[:, 0, :] = rng.integers(0, ol.P, size=(npolys, N), dtype=np.uint64)
    a[:, 1, :] = rng.integers(0, ol.B, size=(npolys, N), dtype=np.uint64)
    return np.ascontiguousarray(a.reshape(-1))


def build_planes(oracle, prm, rng, dim0, num_per, planes):
    """Plaintext planes (u16) + the reference-layout db_buf the oracle consumes."""
    items = dim0 * num_per
    pts = rng.integers(0, prm.p_db, size=(planes, items, N), dtype=np.uint64)
    db = np.zeros(planes * items * N, dtype=np.uint64)
    for pl in range(planes):
        enc = np.zeros(items * N, dtype=np.uint64)
        oracle.so_encode_plaintext(p(enc), p(np.ascontiguousarray(pts[pl].reshape(-1))), items * N, prm.p_db)
        ntt = np.zeros(items * PL, dtype=np.uint64)
        oracle.so_to_ntt(p(ntt), p(enc), items)
        view = db[pl * items * N:(pl + 1) * items * N]
        oracle.so_convert_db(p(view), p(ntt), items, dim0, num_per)
    return pts, db


@pytest.mark.parametrize("dim0,num_per", [(64, 256), (128, 512), (512, 256), (32, 256), (256, 8), (2048, 8), (512, 16), (64, 32), (128, 8),
                                          (64, 64), (256, 128), (128, 64), (64, 4)])
def test_shaped_plane_scans_match_oracle(sb, oracle, dim0, num_per):
    """fastMultiplyQueryByDatabaseDim1 through the compile-time-shaped kernels - k_scan_pack_wide at SpiralPack widths (>= 256
    columns per z-slice), k_scan_pack_narrow (one warp per plane and z) at SpiralStreamPack widths (8 / 16 / 32 columns, and 64 / 128 as slabs of 32) - and,
    for the shapes outside their domains, the generic one, against the oracle; all-maximal residues in two columns and one query
    row, so the accumulators see their largest sums."""
    rng = np.random.default_rng(dim0 * 7 + num_per)
    words = dim0 * num_per * N

    def rnd_pb(shape):
        return rng.integers(0, ol.P, size=shape, dtype=np.uint64) | (rng.integers(0, ol.B, size=shape, dtype=np.uint64) << np.uint64(32))
    big = np.uint64((ol.P - 1) | ((ol.B - 1) << 32))
    db = rnd_pb((N, num_per, dim0))                      # convertDb layout db_buf[z][ii][j]  (src/testing.cpp:316-340)
    db[:, 3 % num_per, :] = big
    db[:, num_per - 1, :] = big
    q = rnd_pb((N, dim0, 2))                             # reorientCiphertextsDim1 layout [z][j][r]
    q[:, :, 1] = big
    db = np.ascontiguousarray(db.reshape(-1)); q = np.ascontiguousarray(q.reshape(-1))
    assert db.size == words
    got = np.zeros(num_per * 2 * 2 * N, dtype=np.uint64)
    want = np.zeros_like(got)
    sb.sb200_kernel_log_reset()
    check(sb.sb200_fastMultiplyQueryByDatabaseDim1(p(got), p(db), p(q), dim0, num_per), sb)
    from spiral_b200.lib import kernel_log
    kernels = kernel_log(sb)
    jp = dim0 // 2
    if num_per % 256 == 0:
        shaped = "k_scan_pack_wide" if jp % 32 == 0 and jp * 32 <= 32768 else None
    else:
        slab = min(num_per, 32)                          # planes of 64 / 128 columns go as slabs of 32
        shaped = "k_scan_pack_narrow" if num_per >= 8 and jp % ((32 // slab) * 32) == 0 and jp * 32 <= 32768 else None
    if shaped:
        assert any(shaped in k for k in kernels) and "k_scan_pack" not in kernels, (shaped, kernels)
    else:
        assert "k_scan_pack" in kernels, kernels
    oracle.so_fast_multiply_dim1(p(want), p(db), p(q), dim0, num_per)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, f"{bad.size} of {want.size} words differ, first at {bad[:5]}"


@pytest.mark.parametrize("cfg,nu1,nu2,mode", [("cfg1", 5, 2, "expand"), ("cfg3", 4, 1, "expand"), ("cfg4", 5, 2, "direct"),
                                              ("cfg4", 6, 3, "direct"), ("cfg1", 5, 3, "direct"), ("cfg5", 5, 1, "expand"),
                                              ("cfg4", 8, 3, "direct")])     # 25 planes x 8 columns, 128 pairs: k_scan_pack_narrow, 5 warps per CTA
def test_pack_server_matches_oracle(sb, oracle, cfg, nu1, nu2, mode):
    prm = ol.make_params(cfg, nu1, nu2)
    rng = np.random.default_rng(nu1 * 100 + nu2)
    dim0, num_per, n = 1 << nu1, 1 << nu2, prm.out_n
    planes, ell, fd = n * n, prm.t_gsw, nu2
    pts, db = build_planes(oracle, prm, rng, dim0, num_per, planes)
    g, stop = C.c_size_t(), C.c_size_t()
    oracle.so_pack_expansion_shape(C.byref(prm), C.byref(g), C.byref(stop))
    g, stop = g.value, stop.value
    vW = rnd_ntt(rng, n * (n + 1) * prm.t_conv)
    W_left = rnd_ntt(rng, g * 2 * prm.t_exp)
    W_right = rnd_ntt(rng, (stop + 1) * 2 * prm.t_exp_right)
    V = rnd_ntt(rng, 2 * 2 * prm.t_conv)
    query = rnd_ntt(rng, 2)
    v_first = rnd_ntt(rng, dim0 * 2)
    v_fold = rnd_ntt(rng, max(fd, 1) * 2 * 2 * ell)
    want = np.zeros((n + 1) * n * N, dtype=np.uint64)
    want_cts = np.zeros(planes * 2 * N, dtype=np.uint64)
    rc = oracle.so_pack_answer(C.byref(prm), int(mode == "expand"), p(query), p(W_left), p(W_right), p(V), p(v_first), p(v_fold),
                               p(vW), p(db), p(want), p(want_cts))
    assert rc == 0

    h = C.c_void_p()
    sp = SpiralParams(nu1, nu2, prm.t_gsw, prm.t_conv, prm.t_exp, prm.t_exp_right, prm.qp_bits, prm.out_n, prm.p_db)
    check(sb.sb200_pack_server_create(C.byref(h), C.byref(sp), 0), sb)
    items = dim0 * num_per
    for pl in range(planes):
        if pl % 2 == 0:
            a = np.ascontiguousarray(pts[pl].astype(np.uint16))
            check(sb.sb200_pack_server_load_plane_items(h, pl, a.ctypes.data_as(P16)), sb)
        else:
            view = np.ascontiguousarray(db[pl * items * N:(pl + 1) * items * N])
            check(sb.sb200_pack_server_load_plane_reference(h, pl, p(view)), sb)
    got = np.zeros_like(want)
    got_cts = np.zeros_like(want_cts)
    if mode == "expand":
        check(sb.sb200_pack_server_set_public_params(h, p(W_left), p(W_right), p(V), p(vW)), sb)
        check(sb.sb200_pack_server_answer(h, query.ctypes.data, got.ctypes.data, got_cts.ctypes.data, None), sb)
    else:
        check(sb.sb200_pack_server_set_public_params(h, None, None, None, p(vW)), sb)
        check(sb.sb200_pack_server_answer_direct(h, v_first.ctypes.data, v_fold.ctypes.data, got.ctypes.data, got_cts.ctypes.data, None), sb)
    assert sb.sb200_pack_server_response_words(h) == want.size
    sb.sb200_pack_server_destroy(h)
    assert np.array_equal(got_cts, want_cts), "folded per-plane ciphertexts differ"
    assert np.array_equal(got, want), "packed + modulus-switched response differs"


@pytest.mark.parametrize("cfg,nu1,nu2,mode,world", [("cfg3", 4, 2, "expand", 2), ("cfg3", 3, 2, "expand", 4), ("cfg4", 5, 3, "direct", 2),
                                                    ("cfg4", 4, 3, "direct", 8), ("cfg1", 4, 1, "expand", 2)])
def test_sharded_pack_servers_reproduce_oracle(sb, oracle, cfg, nu1, nu2, mode, world):
    """world shards on ONE device (strided over the second dimension of every plane): staged API, device-side gather of
    the per-plane surviving ciphertexts, tail folds + pack + modulus switch on shard 0 == the oracle's whole pipeline.
    Covers one-column shards (num_per / world == 1) and both database ingest paths."""
    from spiral_b200.server import PackServer
    prm = ol.make_params(cfg, nu1, nu2)
    rng = np.random.default_rng(nu1 * 1000 + nu2 * 10 + world)
    dim0, num_per, n = 1 << nu1, 1 << nu2, prm.out_n
    planes, ell, fd = n * n, prm.t_gsw, nu2
    pts, db = build_planes(oracle, prm, rng, dim0, num_per, planes)
    g, stop = C.c_size_t(), C.c_size_t()
    oracle.so_pack_expansion_shape(C.byref(prm), C.byref(g), C.byref(stop))
    g, stop = g.value, stop.value
    vW = rnd_ntt(rng, n * (n + 1) * prm.t_conv)
    W_left, W_right, V = rnd_ntt(rng, g * 2 * prm.t_exp), rnd_ntt(rng, (stop + 1) * 2 * prm.t_exp_right), rnd_ntt(rng, 2 * 2 * prm.t_conv)
    query, v_first, v_fold = rnd_ntt(rng, 2), rnd_ntt(rng, dim0 * 2), rnd_ntt(rng, max(fd, 1) * 2 * 2 * ell)
    want = np.zeros((n + 1) * n * N, dtype=np.uint64)
    want_cts = np.zeros(planes * 2 * N, dtype=np.uint64)
    assert oracle.so_pack_answer(C.byref(prm), int(mode == "expand"), p(query), p(W_left), p(W_right), p(V), p(v_first), p(v_fold),
                                 p(vW), p(db), p(want), p(want_cts)) == 0

    sp = SpiralParams(nu1, nu2, prm.t_gsw, prm.t_conv, prm.t_exp, prm.t_exp_right, prm.qp_bits, prm.out_n, prm.p_db)
    shards = [PackServer(sp, rank=r, world=world) for r in range(world)]
    items = dim0 * num_per
    for srv in shards:
        for pl in range(planes):
            if (pl + srv.rank) % 2 == 0:
                srv.load_plane_items(pl, srv.shard_items(pts[pl]).astype(np.uint16))
            else:
                srv.load_plane_reference(pl, np.ascontiguousarray(db[pl * items * N:(pl + 1) * items * N]))
        if mode == "expand":
            srv.set_public_params(W_left, W_right, V, vW)
        else:
            srv.set_public_params(None, None, None, vW)
    import torch
    words = shards[0].partial_words
    gathered = torch.zeros(world * words, dtype=torch.int64, device="cuda")
    resp = torch.zeros(shards[0].response_words, dtype=torch.int64, device="cuda")
    for rep in range(2):                                        # second pass replays the captured graphs
        for srv in shards:
            if mode == "expand":
                srv.upload_query_ptr(query.ctypes.data)
                srv.expand_and_convert()
            else:
                srv.upload_direct_ptr(v_first.ctypes.data, v_fold.ctypes.data)
            srv.scan()
            srv.fold_local()
            srv.copy_partial(gathered.data_ptr() + srv.rank * words * 8)
        torch.cuda.synchronize()
        shards[0].fold_tail(gathered.data_ptr(), resp.data_ptr())
        torch.cuda.synchronize()
        got = resp.cpu().numpy().view(np.uint64)
        got_cts = shards[0].download(shards[0].result_cts_ptr(), planes * 2 * N)
        assert np.array_equal(got_cts, want_cts), f"folded per-plane ciphertexts differ (pass {rep})"
        assert np.array_equal(got, want), f"packed + modulus-switched response differs (pass {rep})"
    with pytest.raises(Exception):
        shards[0].answer(query)                                 # single-shard call on a sharded server must fail loudly
    for srv in shards:
        srv.close()


def test_pack_device_random_database_is_deterministic_and_in_range(sb):
    """sb200_pack_server_load_random (device-side generator): same seed -> same scan output, different seed -> different."""
    from spiral_b200.server import PackServer
    sp = SpiralParams(4, 2, 3, 56, 56, 56, 27, 2, 65536)
    rng = np.random.default_rng(5)
    vW = rnd_ntt(rng, 2 * 3 * 56)
    v_first, v_fold = rnd_ntt(rng, 16 * 2), rnd_ntt(rng, 2 * 2 * 2 * 3)
    outs = []
    for seed in (7, 7, 8):
        srv = PackServer(sp)
        srv.load_random(seed)
        srv.set_public_params(None, None, None, vW)
        outs.append(srv.answer_direct(v_first, v_fold))
        srv.close()
    assert np.array_equal(outs[0], outs[1])
    assert not np.array_equal(outs[0], outs[2])


@pytest.mark.parametrize("cfg,nu1,nu2,mode,world,shard", [
    ("cfg3", 4, 2, "expand", 2, "nu2"), ("cfg3", 3, 2, "expand", 4, "planes"), ("cfg4", 5, 3, "direct", 8, "planes"),
    ("cfg4", 5, 3, "split", 4, "planes"), ("cfg4", 4, 3, "split", 2, "nu2"), ("cfg1", 4, 1, "expand", 2, "planes"), ("cfg4", 5, 2, "direct", 4, "nu2"),
    ("cfg3", 5, 3, "expand", 4, "nu2"), ("cfg1", 5, 3, "expand", 8, "nu2"), ("cfg5", 4, 1, "expand", 2, "nu2")])
def test_pack_peer_exchange_and_one_call_process(sb, oracle, cfg, nu1, nu2, mode, world, shard):
    """The exchange step inside the product (sb200_pack_server_exchange_and_tail, same peer-memory protocol as the Spiral server),
    both shardings - second dimension (tail folds on rank 0) and whole planes (no fold after the exchange) -, the split direct
    upload (each shard uploads 1/world of the first-dimension ciphertexts, the reorientation kernel all-gathers them through
    peer stores) and the one-call form sb200_pack_server_process: world shards on ONE device == the oracle's whole pipeline."""
    from spiral_b200.server import PackServer
    prm = ol.make_params(cfg, nu1, nu2)
    rng = np.random.default_rng(nu1 * 1000 + nu2 * 10 + world)
    dim0, num_per, n = 1 << nu1, 1 << nu2, prm.out_n
    planes, ell, fd = n * n, prm.t_gsw, nu2
    pts, db = build_planes(oracle, prm, rng, dim0, num_per, planes)
    g, stop = C.c_size_t(), C.c_size_t()
    oracle.so_pack_expansion_shape(C.byref(prm), C.byref(g), C.byref(stop))
    g, stop = g.value, stop.value
    vW = rnd_ntt(rng, n * (n + 1) * prm.t_conv)
    W_left, W_right, V = rnd_ntt(rng, g * 2 * prm.t_exp), rnd_ntt(rng, (stop + 1) * 2 * prm.t_exp_right), rnd_ntt(rng, 2 * 2 * prm.t_conv)
    sp = SpiralParams(nu1, nu2, prm.t_gsw, prm.t_conv, prm.t_exp, prm.t_exp_right, prm.qp_bits, prm.out_n, prm.p_db)
    shards = [PackServer(sp, rank=r, world=world, shard=shard) for r in range(world)]
    items = dim0 * num_per
    for srv in shards:
        for pl in range(planes):
            if not srv.owns_plane(pl):
                with pytest.raises(Exception):
                    srv.load_plane_reference(pl, np.ascontiguousarray(db[pl * items * N:(pl + 1) * items * N]))
                continue
            if (pl + srv.rank) % 2 == 0:
                srv.load_plane_items(pl, (pts[pl] if shard == "planes" else srv.shard_items(pts[pl])).astype(np.uint16))
            else:
                srv.load_plane_reference(pl, np.ascontiguousarray(db[pl * items * N:(pl + 1) * items * N]))
        srv.set_public_params(*((W_left, W_right, V, vW) if mode == "expand" else (None, None, None, vW)))
    assert sum(int(sb.sb200_pack_server_local_planes(s.h)) for s in shards) == (planes if shard == "planes" else planes * world)
    for srv in shards:
        srv.xchg_connect_local(shards)
    import torch
    streams = [torch.cuda.Stream() for _ in shards]
    # column-sharded servers share out the EXPANSION of a packed query too: each expands the first-dimension ciphertexts
    # j = rank (mod world) and all-gathers them, reoriented, through peer stores
    for srv in shards:
        assert srv.expansion_sharded() == (mode == "expand" and shard == "nu2" and dim0 % (2 * world) == 0)
    for srv, st in zip(shards, streams):                        # every graph exists before any shard's kernel waits for another's
        srv.prepare(None, st.cuda_stream)
    jc = dim0 // world
    for rep in range(3):                                        # later passes replay the captured graphs and reuse the exchange slots
        query, v_first, v_fold = rnd_ntt(rng, 2), rnd_ntt(rng, dim0 * 2), rnd_ntt(rng, max(fd, 1) * 2 * 2 * ell)
        want = np.zeros((n + 1) * n * N, dtype=np.uint64)
        want_cts = np.zeros(planes * 2 * N, dtype=np.uint64)
        assert oracle.so_pack_answer(C.byref(prm), int(mode == "expand"), p(query), p(W_left), p(W_right), p(V), p(v_first), p(v_fold),
                                     p(vW), p(db), p(want), p(want_cts)) == 0
        # all uploads first (the split upload stores into EVERY shard's query buffer), then the one-call form, rank 0 last:
        # its wait kernel needs the other shards' pushes, and everything shares one device here
        for srv, st in zip(shards, streams):
            if mode == "expand":
                srv.upload_query_ptr(query.ctypes.data, st.cuda_stream)
            elif mode == "direct":
                srv.upload_direct_ptr(v_first.ctypes.data, v_fold.ctypes.data, st.cuda_stream)
            else:
                sl = np.ascontiguousarray(v_first.reshape(dim0, -1)[srv.rank * jc:(srv.rank + 1) * jc].reshape(-1))
                srv.upload_direct_split_ptr(sl.ctypes.data, v_fold.ctypes.data, st.cuda_stream)
                st.synchronize()
        for srv, st in list(zip(shards, streams))[::-1]:
            srv.process(None, st.cuda_stream)
        torch.cuda.synchronize()
        assert all(s.xchg_error() == 0 for s in shards)
        got = shards[0].download(sb.sb200_pack_server_response_ptr(shards[0].h), shards[0].response_words)
        got_cts = shards[0].download(shards[0].result_cts_ptr(), planes * 2 * N)
        assert np.array_equal(got_cts, want_cts), f"folded per-plane ciphertexts differ (pass {rep})"
        assert np.array_equal(got, want), f"packed + modulus-switched response differs (pass {rep})"
    for srv in shards:
        srv.close()
