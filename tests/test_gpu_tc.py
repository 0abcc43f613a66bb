"""GPU suite: the batched first dimension on tcgen05 tensor cores (spiral_b200/csrc/tc_scan.cu).
multiplyQueryByDatabase for up to 16 queries in ONE database pass, computed as exact u8-limb integer matrix products.
Bar: bit-exact against the oracle (src/spiral.cpp:628-999 restated) and against the single-query scan kernel."""
import ctypes as C

import numpy as np
import pytest

from spiral_b200 import SB200Error, SpiralParams
from spiral_b200.server import SpiralServer
from tests import oracle_lib as ol

pytestmark = pytest.mark.gpu
P64 = C.POINTER(C.c_uint64)
N = ol.N


def pack(lo, hi):
    return lo.astype(np.uint64) | (hi.astype(np.uint64) << np.uint64(32))


def rnd_pb(rng, shape):
    return pack(rng.integers(0, ol.P, size=shape, dtype=np.uint64), rng.integers(0, ol.B, size=shape, dtype=np.uint64))


@pytest.mark.parametrize("count", [1, 3, 7, 16])       # tile shapes NB = 16, 16, 32, 48
def test_tc_batched_first_dim_matches_oracle(sb, oracle, count):
    dim0, num_per = 64, 64                              # the smallest shape the 128 x 128-byte tiles accept
    rng = np.random.default_rng(1000 + count)
    db = rnd_pb(rng, (N, num_per, 2, dim0, 2))          # reference layout B[z][ii][c][j][m]  (src/spiral.cpp:1139-1153)
    db[:, 0, 0] = np.uint64((ol.P - 1) | ((ol.B - 1) << 32))     # an all-maximal column: the largest limb sums possible
    db[:, 1, 1] = 0
    queries = []
    for b in range(count):
        q = rnd_pb(rng, (N, dim0, 2, 4))                # reorientCiphertexts layout [z][j][m][4], lane r = 3 is zero
        q[..., 3] = 0
        if b == 0:
            q[:, :, :, 0] = np.uint64((ol.P - 1) | ((ol.B - 1) << 32))
        queries.append(np.ascontiguousarray(q.reshape(-1)))
    db = np.ascontiguousarray(db.reshape(-1))
    words = num_per * 6 * 2 * N
    outs = [np.zeros(words, dtype=np.uint64) for _ in range(count)]
    rc = sb.sb200_multiplyQueryByDatabase_batched((P64 * count)(*[o.ctypes.data_as(P64) for o in outs]),
                                                  (P64 * count)(*[q.ctypes.data_as(P64) for q in queries]), count,
                                                  db.ctypes.data_as(P64), dim0, num_per)
    assert rc == 0, sb.sb200_last_error().decode()
    for b in range(count):
        want = np.zeros(words, dtype=np.uint64)
        oracle.so_multiply_query_by_database(ol.ptr(want), ol.ptr(queries[b]), ol.ptr(db), dim0, num_per)
        got = outs[b]
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, f"query {b} of {count}: {bad.size} of {want.size} words differ, first at {bad[:5]}: got {got[bad[:5]]} want {want[bad[:5]]}"


@pytest.mark.parametrize("count", [5, 16])
def test_tc_scan_equals_single_query_scan_kernel(sb, count):
    """Device-pointer tier: k_scan_tc against k_scan_spiral on the same resident database, two k-chunks, two column tiles
    (count 5: 16-column accumulator blocks, count 16: 48-column blocks)."""
    import torch
    dim0, num_per = 128, 128
    rng = np.random.default_rng(31)
    nu1, nu2 = 7, 7
    db_words = sb.sb200_db_words(nu1, nu2)
    db = torch.from_numpy(rnd_pb(rng, (db_words,)).view(np.int64)).cuda()        # scan layout: any residues are a valid database
    db_tc = torch.empty(db_words * 8, dtype=torch.uint8, device="cuda")
    assert sb.sb200_dev_db_to_tc(db_tc.data_ptr(), db.data_ptr(), dim0, num_per, None) == 0, sb.sb200_last_error()
    q_tc = torch.zeros(sb.sb200_tc_query_bytes(dim0, count), dtype=torch.uint8, device="cuda")
    out_words = num_per * 6 * 2 * N
    t1 = torch.empty(sb.sb200_tc_scratch_bytes(num_per, count), dtype=torch.uint8, device="cuda")
    singles, batched, keep = [], [], []
    for b in range(count):
        q = rnd_pb(rng, (N, dim0, 2, 4)); q[..., 3] = 0
        qd = torch.from_numpy(np.ascontiguousarray(q.reshape(-1)).view(np.int64)).cuda()
        keep.append(qd)
        o = torch.empty(out_words, dtype=torch.int32, device="cuda")
        assert sb.sb200_dev_first_dim(o.data_ptr(), qd.data_ptr(), db.data_ptr(), dim0, num_per, None) == 0, sb.sb200_last_error()
        singles.append(o)
        assert sb.sb200_dev_query_to_tc(q_tc.data_ptr(), qd.data_ptr(), b, count, dim0, None) == 0, sb.sb200_last_error()
        batched.append(torch.full((out_words,), -1, dtype=torch.int32, device="cuda"))
    arr = (C.c_void_p * count)(*[t.data_ptr() for t in batched])
    assert sb.sb200_dev_first_dim_tc(arr, count, count, q_tc.data_ptr(), db_tc.data_ptr(), dim0, num_per, t1.data_ptr(), None) == 0, sb.sb200_last_error()
    torch.cuda.synchronize()
    for b in range(count):
        assert torch.equal(batched[b], singles[b]), f"query {b}"
    assert sb.sb200_dev_first_dim_tc(arr, count, count, q_tc.data_ptr(), db_tc.data_ptr(), 32, num_per, t1.data_ptr(), None) == -3   # 2*dim0 < 128


def test_tc_server_batch_answers_every_client(sb, oracle):
    """Resident server: 6 clients (own keys, own queries) share one database; one tensor-core pass replaces their six
    scans and every response is bit-identical to the single-query path and decodes to the planted record."""
    import torch
    nu1, nu2, count = 6, 6, 6
    sessions = [ol.SpiralSession(oracle, "cfg1", nu1, nu2, seed=40 + k) for k in range(count)]
    a = sessions[0]
    p = a.prm
    prm = SpiralParams(p.nu1, p.nu2, p.t_gsw, p.t_conv, p.t_exp, p.t_exp_right, p.qp_bits, p.out_n, p.p_db)
    srv = SpiralServer(prm)
    srv.load_db_items(a.pts.astype(np.uint16))
    with pytest.raises(SB200Error, match="enable_tc"):
        SpiralServer.scan_batched_tc([srv])
    srv.enable_tc(8)
    servers = [srv] + [srv.view() for _ in range(count - 1)]
    idxs = [0, 1, 77, 1000, 2047, (1 << (nu1 + nu2)) - 1]
    queries = [ses.query(idx) for ses, idx in zip(sessions, idxs)]   # fresh encryption randomness per call: generate once
    want = []
    for s_, ses, q in zip(servers, sessions, queries):
        s_.set_public_params(ses.W_left, ses.W_right, ses.W_conv, ses.V_conv)
        want.append(s_.answer(q))                                    # single-query path (k_scan_spiral)
    resp = [torch.empty(6 * N, dtype=torch.int64, device="cuda") for _ in range(count)]
    for s_, q in zip(servers, queries):
        s_.upload_query(q); s_.expand_and_convert()
    torch.cuda.synchronize()
    SpiralServer.scan_batched_tc(servers)
    torch.cuda.synchronize()
    for k, s_ in enumerate(servers):
        s_.lift(); s_.fold_local(); s_.fold_tail(s_.partial_ct_ptr(), resp[k].data_ptr())
    torch.cuda.synchronize()
    for k, (ses, idx) in enumerate(zip(sessions, idxs)):
        got = resp[k].cpu().numpy().view(np.uint64)
        assert np.array_equal(got, want[k]), f"client {k}"
        assert np.array_equal(ses.decode(got), a.pts[idx]), f"client {k} decodes record {idx}"
    with pytest.raises(SB200Error, match="capacity"):
        SpiralServer.scan_batched_tc(servers + [servers[0]] * 3)
    # keep ONLY the limb-tile copy: the scan-layout database is freed (one copy in HBM, not two) and the single-query calls of
    # the owner and of its views go through a one-query pass of the tensor-core kernel - same responses, bit for bit
    free0 = torch.cuda.mem_get_info()[0]
    srv.tc_only()
    assert torch.cuda.mem_get_info()[0] - free0 >= srv.db_bytes * 9 // 10, "the scan-layout copy was not released"
    assert not sb.sb200_server_db_ptr(srv.h)
    for k in (0, 3, 5):
        assert np.array_equal(servers[k].answer(queries[k]), want[k]), f"tc_only, client {k}"
    SpiralServer.scan_batched_tc(servers)                             # the batched pass is unchanged
    torch.cuda.synchronize()
    servers[1].lift(); servers[1].fold_local(); servers[1].fold_tail(servers[1].partial_ct_ptr(), resp[1].data_ptr())
    torch.cuda.synchronize()
    assert np.array_equal(resp[1].cpu().numpy().view(np.uint64), want[1])
    with pytest.raises(SB200Error, match="released"):
        SpiralServer.scan_batched(servers[:2])
    with pytest.raises(SB200Error, match="released"):
        srv.save_db("/tmp/sb200_tc_only_should_not_exist.snap")
    srv.load_db_items(a.pts.astype(np.uint16))                        # a reload starts over: scan layout back, tensor-core state dropped
    assert np.array_equal(srv.answer(queries[0]), want[0])
    with pytest.raises(SB200Error, match="enable_tc"):
        SpiralServer.scan_batched_tc([srv])
    for s_ in reversed(servers):
        s_.close()
    for ses in sessions:
        ses.close()


# ---------------------------------------------------------------------------------------------------------------
# SpiralPack shape: 2x1 ciphertexts, 1x1 plaintexts, planes (fastMultiplyQueryByDatabaseDim1, src/testing.cpp:364-593)
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("count", [1, 8, 16])           # tile shapes NB = 16, 16, 32 (two rows per query)
def test_tc_batched_pack_first_dim_matches_oracle(sb, oracle, count):
    dim0, num_per = 128, 128
    rng = np.random.default_rng(2000 + count)
    db = rnd_pb(rng, (N, num_per, dim0))                 # convertDb layout db_buf[z][ii][j]  (src/testing.cpp:316-340)
    db[:, 0, :] = np.uint64((ol.P - 1) | ((ol.B - 1) << 32))
    db = np.ascontiguousarray(db.reshape(-1))
    queries = []
    for b in range(count):
        q = rnd_pb(rng, (N, dim0, 2))                    # reorientCiphertextsDim1 layout [z][j][r]
        if b == count - 1:
            q[:, :, 1] = np.uint64((ol.P - 1) | ((ol.B - 1) << 32))
        queries.append(np.ascontiguousarray(q.reshape(-1)))
    words = num_per * 2 * 2 * N
    outs = [np.zeros(words, dtype=np.uint64) for _ in range(count)]
    rc = sb.sb200_fastMultiplyQueryByDatabaseDim1_batched((P64 * count)(*[o.ctypes.data_as(P64) for o in outs]), db.ctypes.data_as(P64),
                                                          (P64 * count)(*[q.ctypes.data_as(P64) for q in queries]), count, dim0, num_per)
    assert rc == 0, sb.sb200_last_error().decode()
    for b in range(count):
        want = np.zeros(words, dtype=np.uint64)
        oracle.so_fast_multiply_dim1(ol.ptr(want), ol.ptr(db), ol.ptr(queries[b]), dim0, num_per)
        bad = np.nonzero(outs[b] != want)[0]
        assert bad.size == 0, f"query {b} of {count}: {bad.size} of {want.size} words differ, first at {bad[:5]}"


def test_tc_pack_server_batch_equals_single_query_path(sb):
    """Resident SpiralPack server (out_n = 2: four planes): 5 clients with their own keys share the planes; one tensor-core
    pass over all planes replaces their five scans; every packed response equals the single-query path's bit for bit."""
    import torch
    from spiral_b200.server import PackServer
    nu1, nu2, count = 7, 7, 5
    prm = SpiralParams(nu1, nu2, 8, 4, 8, 56, 20, 2, 256)
    rng = np.random.default_rng(99)

    def rnd_ntt(npolys):
        a = np.empty((npolys, 2, N), dtype=np.uint64)
        a[:, 0, :] = rng.integers(0, ol.P, size=(npolys, N), dtype=np.uint64)
        a[:, 1, :] = rng.integers(0, ol.B, size=(npolys, N), dtype=np.uint64)
        return np.ascontiguousarray(a.reshape(-1))
    srv = PackServer(prm)
    srv.load_random(7)
    with pytest.raises(SB200Error, match="enable_tc"):
        PackServer.scan_batched_tc([srv])
    srv.enable_tc(count)
    servers = [srv] + [srv.view() for _ in range(count - 1)]
    nbits = 8 * nu2
    g = int(np.ceil(np.log2(nbits + (1 << nu1))))
    stop = int(np.ceil(np.log2(nbits)))
    queries, want = [], []
    for s_ in servers:
        s_.set_public_params(rnd_ntt(g * 2 * 8), rnd_ntt((stop + 1) * 2 * 56), rnd_ntt(2 * 2 * 4), rnd_ntt(2 * 3 * 4))
        queries.append(rnd_ntt(2))
        want.append(s_.answer(queries[-1]))                        # single-query path (k_scan_pack)
    assert len({w.tobytes() for w in want}) == count
    with pytest.raises(SB200Error, match="view"):
        servers[1].load_random(3)
    for s_, q in zip(servers, queries):
        s_.upload_query_ptr(q.ctypes.data); s_.expand_and_convert()
    torch.cuda.synchronize()
    PackServer.scan_batched_tc(servers)
    torch.cuda.synchronize()
    resp = [torch.empty(srv.response_words, dtype=torch.int64, device="cuda") for _ in range(count)]
    for k, s_ in enumerate(servers):
        s_.fold_local(); s_.fold_tail(s_.partial_cts_ptr(), resp[k].data_ptr())
    torch.cuda.synchronize()
    for k in range(count):
        assert np.array_equal(resp[k].cpu().numpy().view(np.uint64), want[k]), f"client {k}"
    free0 = torch.cuda.mem_get_info()[0]
    srv.tc_only()                                                     # the limb-tile planes are the only copy from here on
    assert torch.cuda.mem_get_info()[0] - free0 >= srv.db_bytes * 9 // 10
    for k in (0, count - 1):
        assert np.array_equal(servers[k].answer(queries[k]), want[k]), f"tc_only, client {k}"
    with pytest.raises(SB200Error, match="released"):
        srv.load_random(8)
    for s_ in reversed(servers):
        s_.close()
