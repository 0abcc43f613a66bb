"""GPU suite: the resident server (tier-3 C-ABI) on real encryptions.
  * response bit-exact with the oracle's whole pipeline (and the intermediate first-dimension cts)
  * decoded record == planted record (the reference's "Is correct?: 1" gate)
  * both database ingest paths (plaintext items, reference-layout buffer) give the same answers
  * edge indices (first / last record), stopround == 0 and != 0 shapes, odd t_GSW."""
import ctypes as C

import numpy as np
import pytest

from spiral_b200 import SpiralParams
from spiral_b200.lib import check
from spiral_b200.server import SpiralServer
from tests import oracle_lib as ol

pytestmark = pytest.mark.gpu


def sb_params(so):
    return SpiralParams(so.nu1, so.nu2, so.t_gsw, so.t_conv, so.t_exp, so.t_exp_right, so.qp_bits, so.out_n, so.p_db)


# the last two shapes have 64 / 128 database columns per z: the two-columns-per-thread scan with several z-slices per CTA
@pytest.mark.parametrize("cfg,nu1,nu2", [("cfg1", 2, 2), ("cfg1", 4, 1), ("cfg5", 3, 2), ("cfg1", 5, 3), ("cfg1", 6, 2), ("cfg1", 1, 1), ("cfg4", 2, 2), ("cfg3", 4, 2), ("cfg1", 7, 4),
                                         ("cfg1", 3, 5), ("cfg1", 2, 6)])
def test_server_matches_oracle_and_decodes(sb, oracle, cfg, nu1, nu2):
    s = ol.SpiralSession(oracle, cfg, nu1, nu2, seed=11)
    Bbuf = s.reference_db()
    srv = SpiralServer(sb_params(s.prm))
    srv.load_db_items(s.pts.astype(np.uint16))
    srv.set_public_params(s.W_left, s.W_right, s.W_conv, s.V_conv)
    for idx in (0, s.total_n - 1, s.total_n // 3):
        q = s.query(idx)
        want_resp, _, want_first = s.oracle_answer(q, Bbuf)
        got = srv.answer(q)
        assert np.array_equal(got, want_resp), f"response differs at idx {idx}"
        if cfg != "cfg4":      # cfg4's parameters are a Pack-variant set: as a Spiral query they exceed the noise budget
            assert np.array_equal(s.decode(got), s.pts[idx]), f"decode failed at idx {idx}"
    srv.close()
    s.close()


def test_first_dim_tap_and_reference_layout_ingest(sb, oracle):
    s = ol.SpiralSession(oracle, "cfg1", 4, 2, seed=5)
    Bbuf = s.reference_db()
    q = s.query(9)
    want_resp, _, want_first = s.oracle_answer(q, Bbuf)
    a = SpiralServer(sb_params(s.prm))
    a.load_db_reference(Bbuf)                      # the reference's own B buffer
    a.set_public_params(s.W_left, s.W_right, s.W_conv, s.V_conv)
    a.upload_query(q)
    a.expand_and_convert()
    a.first_dim()
    assert np.array_equal(a.first_dim_cts(), want_first), "raw cts after the first dimension differ"
    assert np.array_equal(a.answer(q), want_resp)
    a.close()
    s.close()


def test_sharded_servers_reproduce_single_gpu_answer(sb, oracle):
    """world = 2 and 4 on ONE device: strided shards + gather + tail folds == unsharded response."""
    import torch
    s = ol.SpiralSession(oracle, "cfg1", 3, 3, seed=9)
    Bbuf = s.reference_db()
    q = s.query(41)
    want_resp, _, _ = s.oracle_answer(q, Bbuf)
    for world in (2, 4):
        servers = [SpiralServer(sb_params(s.prm), rank=r, world=world) for r in range(world)]
        gathered = torch.empty(world * 6 * ol.N, dtype=torch.int64, device="cuda")
        resp = torch.empty(6 * ol.N, dtype=torch.int64, device="cuda")
        for r, srv in enumerate(servers):
            if r % 2 == 0:
                srv.load_db_items(srv.shard_items(s.pts).astype(np.uint16))
            else:
                srv.load_db_reference(Bbuf)
            srv.set_public_params(s.W_left, s.W_right, s.W_conv, s.V_conv)
            srv.upload_query(q); srv.expand_and_convert(); srv.first_dim(); srv.fold_local()
            torch.cuda.synchronize()
            part = torch.from_numpy(srv.download(srv.partial_ct_ptr(), 6 * ol.N).view(np.int64)).cuda()
            gathered[r * 6 * ol.N:(r + 1) * 6 * ol.N] = part
        servers[0].fold_tail(gathered.data_ptr(), resp.data_ptr())
        torch.cuda.synchronize()
        got = resp.cpu().numpy().view(np.uint64)
        assert np.array_equal(got, want_resp), f"world={world} (host-side gather)"
        # the same exchange through the peer-memory kernels (flags + stores into rank 0's buffer), three queries in a
        # row so the slot / epoch / ack protocol wraps around; non-root shards push first, rank 0 last
        for srv in servers:
            srv.xchg_connect_local(servers)
        for idx in (41, 7, 60, 41):
            q2 = s.query(idx)
            want2, _, _ = s.oracle_answer(q2, Bbuf)
            for srv in servers:
                srv.upload_query(q2); srv.expand_and_convert(); srv.first_dim(); srv.fold_local()
            resp.zero_()
            torch.cuda.synchronize()
            for srv in reversed(servers):
                srv.exchange_and_tail(resp.data_ptr() if srv.rank == 0 else None)
            assert all(srv.xchg_error() == 0 for srv in servers)
            torch.cuda.synchronize()
            got2 = resp.cpu().numpy().view(np.uint64)
            assert np.array_equal(got2, want2), f"world={world} idx={idx} (peer-memory exchange)"
        # and as ONE call per shard (sb200_server_process: all stages + the exchange), the form bench.py times
        q3 = s.query(13)
        want3, _, _ = s.oracle_answer(q3, Bbuf)
        resp.zero_()
        torch.cuda.synchronize()
        for srv in reversed(servers):
            srv.upload_query(q3)
            srv.process(resp.data_ptr() if srv.rank == 0 else None)
        assert all(srv.xchg_error() == 0 for srv in servers)
        torch.cuda.synchronize()
        assert np.array_equal(resp.cpu().numpy().view(np.uint64), want3), f"world={world} (one call per shard)"
        for srv in servers:
            srv.close()
    s.close()


def test_process_is_the_staged_calls_in_one(sb, oracle):
    """sb200_server_process (with and without stage events) leaves the response sb200_server_answer returns."""
    import ctypes
    import torch
    s = ol.SpiralSession(oracle, "cfg1", 4, 2, seed=31)
    srv = SpiralServer(sb_params(s.prm))
    srv.load_db_items(s.pts.astype(np.uint16))
    srv.set_public_params(s.W_left, s.W_right, s.W_conv, s.V_conv)
    resp = torch.zeros(6 * ol.N, dtype=torch.int64, device="cuda")
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for e in evs:
            e.record()
        marks = (ctypes.c_void_p * 4)(*[e.cuda_event for e in evs])
        # without marks the whole query is ONE graph (captured at the first call, replayed afterwards); with marks the stages are
        # separate graphs; the staged calls are the third form - all three against the oracle, on different queries
        for k, use_marks in enumerate((False, True, False, True, False)):
            q = s.query(11 + 3 * k)
            want, _, _ = s.oracle_answer(q, s.reference_db())
            assert np.array_equal(srv.answer(q), want)
            resp.zero_()
            srv.upload_query(q, stream.cuda_stream)
            srv.process(resp.data_ptr(), stream.cuda_stream, marks if use_marks else None)
            stream.synchronize()
            assert np.array_equal(resp.cpu().numpy().view(np.uint64), want), f"process, marks={use_marks}"
            resp.zero_()
            srv.upload_query(q, stream.cuda_stream)
            srv.expand_and_convert(stream.cuda_stream); srv.first_dim(stream.cuda_stream); srv.fold_local(stream.cuda_stream)
            srv.fold_tail(srv.partial_ct_ptr(), resp.data_ptr(), stream.cuda_stream)
            stream.synchronize()
            assert np.array_equal(resp.cpu().numpy().view(np.uint64), want), "staged calls"
        assert 0 < evs[1].elapsed_time(evs[2]) < evs[0].elapsed_time(evs[3])          # scan inside the whole query
    with pytest.raises(Exception, match="connected"):
        shard = SpiralServer(sb_params(s.prm), rank=0, world=2)
        shard.process(resp.data_ptr())
    srv.close()
    s.close()


def test_views_share_the_database_and_run_concurrently(sb, oracle):
    """Two clients (different keys) in flight on one GPU: a view scans its parent's resident database."""
    import torch
    a = ol.SpiralSession(oracle, "cfg1", 4, 3, seed=21)
    b = ol.SpiralSession(oracle, "cfg1", 4, 3, seed=22)
    b.pts = a.pts                                   # same database, different client keys
    Bbuf = a.reference_db()
    srv = SpiralServer(sb_params(a.prm))
    srv.load_db_items(a.pts.astype(np.uint16))
    view = srv.view()
    srv.set_public_params(a.W_left, a.W_right, a.W_conv, a.V_conv)
    view.set_public_params(b.W_left, b.W_right, b.W_conv, b.V_conv)
    with pytest.raises(Exception):
        view.load_db_items(a.pts.astype(np.uint16))  # a view never owns a database
    qa, qb = a.query(100), b.query(23)
    want_a, _, _ = a.oracle_answer(qa, Bbuf)
    want_b, _, _ = b.oracle_answer(qb, Bbuf)
    sa, sb_ = torch.cuda.Stream(), torch.cuda.Stream()
    ra = torch.empty(6 * ol.N, dtype=torch.int64, device="cuda")
    rb = torch.empty(6 * ol.N, dtype=torch.int64, device="cuda")
    for _ in range(3):                               # interleave the stages of the two clients on two streams
        srv.upload_query(qa, sa.cuda_stream); view.upload_query(qb, sb_.cuda_stream)
        srv.expand_and_convert(sa.cuda_stream); view.expand_and_convert(sb_.cuda_stream)
        view.first_dim(sb_.cuda_stream); srv.first_dim(sa.cuda_stream)
        srv.fold_local(sa.cuda_stream); view.fold_local(sb_.cuda_stream)
        view.fold_tail(view.partial_ct_ptr(), rb.data_ptr(), sb_.cuda_stream)
        srv.fold_tail(srv.partial_ct_ptr(), ra.data_ptr(), sa.cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(ra.cpu().numpy().view(np.uint64), want_a)
        assert np.array_equal(rb.cpu().numpy().view(np.uint64), want_b)
    assert np.array_equal(a.decode(ra.cpu().numpy().view(np.uint64)), a.pts[100])
    assert np.array_equal(b.decode(rb.cpu().numpy().view(np.uint64)), a.pts[23])
    view.close(); srv.close(); a.close(); b.close()


@pytest.mark.parametrize("count", [2, 4])
def test_batched_scan_equals_individual_scans(sb, count):
    """several queries answered in one database pass give exactly the single-query ciphertexts."""
    import torch
    from spiral_b200 import SpiralParams
    nu1, nu2 = 4, 6                                   # 2 * num_per = 128 columns
    prm = SpiralParams(nu1, nu2, 8, 4, 8, 56, 20, 2, 256)
    rng = np.random.default_rng(77)
    N = ol.N

    def rnd(npolys):
        a = np.empty((npolys, 2, N), dtype=np.uint64)
        a[:, 0, :] = rng.integers(0, ol.P, size=(npolys, N), dtype=np.uint64)
        a[:, 1, :] = rng.integers(0, ol.B, size=(npolys, N), dtype=np.uint64)
        return np.ascontiguousarray(a.reshape(-1))
    srv = SpiralServer(prm)
    srv.load_db_items(rng.integers(0, 256, size=(1 << (nu1 + nu2), 4, N), dtype=np.uint16))
    servers = [srv] + [srv.view() for _ in range(count - 1)]
    nbits = 8 * nu2
    g = int(np.ceil(np.log2(nbits + (1 << nu1))))
    n_right = g                                       # nbits > dim0 -> stopround = 0
    single = []
    for s_ in servers:
        s_.set_public_params(rnd(g * 2 * 8), rnd(n_right * 2 * 56), rnd(3 * 2 * 4), rnd(3 * 2 * 4))
        s_.upload_query(rnd(2)); s_.expand_and_convert(); s_.first_dim()
        single.append(s_.first_dim_cts())
    torch.cuda.synchronize()
    SpiralServer.scan_batched(servers)
    for s_ in servers:
        s_.lift()
    for k, s_ in enumerate(servers):
        assert np.array_equal(s_.first_dim_cts(), single[k]), f"query {k} of {count}"
    assert len({c.tobytes() for c in single}) == count, "the queries must differ for the test to mean anything"
    for s_ in reversed(servers):
        s_.close()


def test_packed_response_wire_format(sb, oracle):
    """sb200_server_answer_packed: the modulus-switched response bit-packed on the device (row 0 at QPBITS bits, rows 1-2
    at log2(4p) bits - the size print_summary reports, src/spiral.cpp:229-232); unpacking it gives the raw response."""
    s = ol.SpiralSession(oracle, "cfg1", 4, 2, seed=21)
    srv = SpiralServer(sb_params(s.prm))
    srv.load_db_items(s.pts.astype(np.uint16))
    srv.set_public_params(s.W_left, s.W_right, s.W_conv, s.V_conv)
    q = s.query(17)
    raw = srv.answer(q)
    nbytes = sb.sb200_server_packed_response_bytes(srv.h)
    N = ol.N
    assert nbytes == (2 * N * s.prm.qp_bits + 4 * N * 10) // 8 == 20480
    packed = np.zeros(nbytes // 8, dtype=np.uint64)
    assert sb.sb200_server_answer_packed(srv.h, q.ctypes.data, packed.ctypes.data, None) == 0, sb.sb200_last_error()
    back = np.zeros(6 * N, dtype=np.uint64)
    P64 = C.POINTER(C.c_uint64)
    assert sb.sb200_unpack_response(back.ctypes.data_as(P64), packed.ctypes.data_as(P64), 2 * N, 4 * N, s.prm.qp_bits, s.prm.p_db) == 0
    assert np.array_equal(back, raw)
    assert np.array_equal(s.decode(back), s.pts[17])
    # the packing agrees with the reference's own bit writer (oracle restatement of write_arbitrary_bits)
    for k in (0, 1, 63, 64, 2 * N - 1):
        assert oracle.so_read_arbitrary_bits(ol.ptr(packed), k * s.prm.qp_bits, s.prm.qp_bits) == raw[k]
    srv.close()
    s.close()


@pytest.mark.parametrize("cfg,nu1,nu2,world", [("cfg1", 6, 3, 2), ("cfg1", 5, 3, 4), ("cfg5", 6, 2, 4), ("cfg1", 7, 4, 8), ("cfg1", 7, 3, 8)])
def test_sharded_expansion_with_fused_all_gather(sb, oracle, cfg, nu1, nu2, world):
    """Connected shards with 2^nu1 / world a multiple of 8 no longer replicate the query-side work: each rank expands only the
    ancestors of the first-dimension ciphertexts j = rank (mod world), converts those, and its ScalToMat kernel stores them into
    EVERY rank's query buffer (peer stores + flags); the scan waits for all slices.  world shards on one device, several queries
    in a row (the slices of query k+1 may only land after every rank has scanned query k), == the oracle's unsharded answer."""
    import torch
    s = ol.SpiralSession(oracle, cfg, nu1, nu2, seed=21)
    Bbuf = s.reference_db()
    servers = [SpiralServer(sb_params(s.prm), rank=r, world=world) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in servers]
    for srv in servers:
        srv.load_db_items(srv.shard_items(s.pts).astype(np.uint16))
        srv.set_public_params(s.W_left, s.W_right, s.W_conv, s.V_conv)
    for srv in servers:
        srv.xchg_connect_local(servers)
    resp = torch.zeros(6 * ol.N, dtype=torch.int64, device="cuda")
    # graphs are built up front: capturing / instantiating one shard's graphs while ANOTHER shard's kernel spins on a flag (all on
    # this one device) can block the host until the spin times out; one process per GPU never meets this
    for srv, st in zip(servers, streams):
        srv.prepare(resp.data_ptr() if srv.rank == 0 else None, st.cuda_stream)
    for k, idx in enumerate((5, s.total_n - 1, 77 % s.total_n, 5)):
        q = s.query(idx)
        want, _, _ = s.oracle_answer(q, Bbuf)
        resp.zero_()
        torch.cuda.synchronize()
        # all uploads first: a host-to-device copy from pageable memory can block the host behind ANOTHER shard's kernel that is
        # spinning on a flag on this same device (one process per GPU with pinned buffers - the deployed form - never meets this)
        for srv, st in zip(servers, streams):
            srv.upload_query(q, st.cuda_stream)
        torch.cuda.synchronize()
        for srv, st in list(zip(servers, streams))[::-1]:        # rank 0 last: its wait kernel needs the others' pushes
            srv.process(resp.data_ptr() if srv.rank == 0 else None, st.cuda_stream)
        torch.cuda.synchronize()
        errs = [srv.xchg_error() for srv in servers]
        assert not any(errs), f"query {k}: exchange time-outs per rank {errs} (1 push, 2 gather, 3 query slices, 4 GSW columns)"
        assert all(sb.sb200_server_expansion_sharded(srv.h) == 1 for srv in servers), "the expansion was replicated, not sharded"
        got = resp.cpu().numpy().view(np.uint64)
        assert np.array_equal(got, want), f"query {k} (idx {idx}): sharded expansion changed the response"
        assert np.array_equal(s.decode(got), s.pts[idx])
    # the staged calls take the same path (join before returning from expand_and_convert)
    q = s.query(9)
    want, _, _ = s.oracle_answer(q, Bbuf)
    for srv, st in zip(servers, streams):
        srv.upload_query(q, st.cuda_stream)
    torch.cuda.synchronize()
    for srv, st in list(zip(servers, streams))[::-1]:            # first call of the staged form: its graphs exist already (same slots)
        srv.expand_and_convert(st.cuda_stream); srv.first_dim(st.cuda_stream); srv.fold_local(st.cuda_stream)
        srv.exchange_and_tail(resp.data_ptr() if srv.rank == 0 else None, st.cuda_stream)
    torch.cuda.synchronize()
    assert all(srv.xchg_error() == 0 for srv in servers)
    assert np.array_equal(resp.cpu().numpy().view(np.uint64), want)
    for srv in servers:
        srv.close()
    s.close()


@pytest.mark.parametrize("nu1,nu2,dws", [(4, 2, 256), (6, 2, 64), (3, 3, 2048), (5, 1, 1)])
def test_implicit_database_scan_reads_slice_z_mod_working_set(sb, oracle, nu1, nu2, dws):
    """cfg2, the reference's --random-data mode (src/spiral.cpp:647,1032-1081): the resident buffer holds `dws` z-slices and NTT
    coefficient z is multiplied with slice z mod dws.  Checked against the oracle's scan of the database written out in full."""
    rng = np.random.default_rng(nu1 * 10 + nu2)
    dim0, num_per, N = 1 << nu1, 1 << nu2, ol.N
    row = num_per * 2 * dim0 * 2                                   # words of one z-slice in load_db's layout

    def rnd_pb(n):
        return rng.integers(0, ol.P, size=n, dtype=np.uint64) | (rng.integers(0, ol.B, size=n, dtype=np.uint64) << np.uint64(32))
    slices = np.ascontiguousarray(rnd_pb(dws * row))
    full = np.ascontiguousarray(np.tile(slices.reshape(dws, row), (N // dws, 1)).reshape(-1))
    q = rnd_pb(N * dim0 * 2 * 4).reshape(N, dim0, 2, 4)
    q[..., 3] = 0
    q = np.ascontiguousarray(q.reshape(-1))
    want = np.zeros(num_per * 6 * 2 * N, dtype=np.uint64)
    oracle.so_multiply_query_by_database(ol.ptr(want), ol.ptr(q), ol.ptr(full), dim0, num_per)
    prm = ol.make_params("cfg1", nu1, nu2)
    srv = SpiralServer(sb_params(prm))
    srv.load_db_implicit(slices, dws)
    assert sb.sb200_server_db_slices(srv.h) == dws
    got = np.zeros_like(want)
    check(sb.sb200_server_scan_host(srv.h, ol.ptr(q), ol.ptr(got)), sb)
    assert np.array_equal(ol.canon(got, ol.KIND_NTT), ol.canon(want, ol.KIND_NTT))
    with pytest.raises(Exception):
        srv.enable_tc(4)                                           # the tensor-core copy needs an explicit database
    srv.load_db_reference(full)                                    # back to an explicit database on the same server
    assert sb.sb200_server_db_slices(srv.h) == N
    check(sb.sb200_server_scan_host(srv.h, ol.ptr(q), ol.ptr(got)), sb)
    assert np.array_equal(ol.canon(got, ol.KIND_NTT), ol.canon(want, ol.KIND_NTT))
    srv.close()


def test_implicit_constant_database_answers_every_index_with_the_constant(sb, oracle):
    """The whole cfg2 pipeline at a small shape: device-generated constant database, real query, decoded record == the constant."""
    s = ol.SpiralSession(oracle, "cfg1", 5, 3, seed=3)
    srv = SpiralServer(sb_params(s.prm))
    srv.load_db_implicit_constant(29, 128)
    srv.set_public_params(s.W_left, s.W_right, s.W_conv, s.V_conv)
    want = np.zeros((4, ol.N), dtype=np.uint64)
    want[:, 0] = 29
    for idx in (0, 100, s.total_n - 1):
        assert np.array_equal(s.decode(srv.answer(s.query(idx))), want)
    srv.close()
    s.close()


@pytest.mark.parametrize("dim0,num_per", [(64, 128), (128, 256), (512, 128), (32, 128), (64, 64), (512, 64), (1024, 32), (256, 16), (2048, 16)])
def test_wide_shard_scan_matches_oracle(sb, oracle, dim0, num_per):
    """multiplyQueryByDatabase (src/spiral.cpp:628-999) at shard widths of 256 database columns and more - cfg1 / cfg5's shapes
    scaled down in the first dimension - and at the narrow shards of a database spread over several GPUs (32 / 64 / 128 columns):
    the TMA-ring kernel (k_scan_spiral_tma) where the shape is in its domain (64..1024 first-dimension ciphertexts), the register-
    streaming kernels otherwise, all against the oracle on random residues with all-maximal columns and an
    all-maximal query row (largest accumulator sums, every fold step)."""
    import ctypes as C
    from spiral_b200.lib import check, kernel_log
    P64 = C.POINTER(C.c_uint64)
    rng = np.random.default_rng(dim0 * 31 + num_per)
    N = ol.N

    def rnd_pb(shape):
        return rng.integers(0, ol.P, size=shape, dtype=np.uint64) | (rng.integers(0, ol.B, size=shape, dtype=np.uint64) << np.uint64(32))
    big = np.uint64((ol.P - 1) | ((ol.B - 1) << 32))
    db = rnd_pb((N, num_per, 2, dim0, 2))               # reference layout B[z][ii][c][j][m]  (src/spiral.cpp:1139-1153)
    db[:, 0, 0] = big
    db[:, num_per - 1, 1] = big
    q = rnd_pb((N, dim0, 2, 4))                         # reorientCiphertexts layout [z][j][m][4], lane r = 3 is zero
    q[:, :, :, 1] = big
    q[..., 3] = 0
    db = np.ascontiguousarray(db.reshape(-1)); q = np.ascontiguousarray(q.reshape(-1))
    words = num_per * 6 * 2 * N
    got, want = np.zeros(words, dtype=np.uint64), np.zeros(words, dtype=np.uint64)
    sb.sb200_kernel_log_reset()
    check(sb.sb200_multiplyQueryByDatabase(got.ctypes.data_as(P64), q.ctypes.data_as(P64), db.ctypes.data_as(P64), dim0, num_per), sb)
    kernels = kernel_log(sb)
    ic = 2 * num_per
    assert ("k_scan_spiral_tma" in kernels) == (64 <= dim0 <= 1024 and (ic % 256 == 0 or ic in (32, 64, 128))), kernels
    if (dim0, num_per) == (2048, 16):                   # outside the ring kernel's domain: narrow-shard kernel, query slice in chunks
        assert any("k_scan_spiral_jsplit" in k for k in kernels), kernels
    oracle.so_multiply_query_by_database(ol.ptr(want), ol.ptr(q), ol.ptr(db), dim0, num_per)
    bad = np.nonzero(got != want)[0]
    assert bad.size == 0, f"{bad.size} of {want.size} words differ, first at {bad[:5]}"
