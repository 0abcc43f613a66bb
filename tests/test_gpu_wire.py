"""GPU suite: wire and on-disk formats through the C-ABI, against oracle/wire_format.c.
  * wire query (seed-compressed and fully packed) -> the same dev-NTT ciphertext the oracle expands, slot for slot
  * a whole exchange in wire form: wire query in, QPBITS-packed response out == the oracle's response, and the record decodes
  * database from a record stream (memory and file, sharded and not) == the database from plaintext items
  * snapshot of the preprocessed database: save -> load round trip, parameter / integrity checks fail loudly
  * Pack server: wire query + record stream against the oracle's whole Pack pipeline."""
import ctypes as C
import os

import numpy as np
import pytest

from spiral_b200 import SpiralParams
from spiral_b200.lib import SB200Error, check
from spiral_b200.server import PackServer, SpiralServer
from tests import oracle_lib as ol
from tests.test_gpu_pack import build_planes, p, rnd_ntt

pytestmark = pytest.mark.gpu
N, PL = ol.N, 2 * ol.N


def sb_params(so):
    return SpiralParams(so.nu1, so.nu2, so.t_gsw, so.t_conv, so.t_exp, so.t_exp_right, so.qp_bits, so.out_n, so.p_db)


@pytest.mark.parametrize("kind", [ol.WIRE_SEEDED, ol.WIRE_FULL])
def test_query_from_wire_matches_oracle_expansion(sb, oracle, kind):
    """tier 1: k_query_from_wire against so_wire_query_expand on client-made and on adversarial (all-ones) payloads."""
    import torch
    s = ol.SpiralSession(oracle, "cfg1", 2, 2, seed=3)
    wires = [s.query_wire(i, kind) for i in (0, 7)]
    worst = wires[0].copy(); worst[8:] = 0xFF                 # every 56-bit value = 2^56 - 1 > Q, every seed byte 0xFF
    wires.append(worst)
    for wire in wires:
        want = ol.wire_expand(oracle, wire)
        assert want is not None and sb.sb200_wire_query_bytes(kind) == wire.size
        d_wire = torch.from_numpy(np.concatenate([wire, np.zeros(8, dtype=np.uint8)])).cuda()
        d_cv = torch.zeros(2 * PL, dtype=torch.int32, device="cuda")
        check(sb.sb200_dev_query_from_wire(d_cv.data_ptr(), d_wire.data_ptr(), kind, None), sb)
        torch.cuda.synchronize()
        got = d_cv.cpu().numpy().view(np.uint32).astype(np.uint64)
        assert np.array_equal(got, ol.canon(want, ol.KIND_NTT))
    s.close()


@pytest.mark.parametrize("cfg,nu1,nu2", [("cfg1", 4, 2), ("cfg5", 3, 3), ("cfg1", 6, 2)])
def test_whole_exchange_in_wire_form(sb, oracle, cfg, nu1, nu2):
    s = ol.SpiralSession(oracle, cfg, nu1, nu2, seed=21)
    Bbuf = s.reference_db()
    srv = SpiralServer(sb_params(s.prm))
    srv.load_db_records(s.records())                           # record stream instead of plaintext items
    srv.set_public_params(s.W_left, s.W_right, s.W_conv, s.V_conv)
    for rep, (idx, kind) in enumerate([(0, ol.WIRE_SEEDED), (s.total_n - 1, ol.WIRE_FULL), (s.total_n // 2, ol.WIRE_SEEDED),
                                       (5, ol.WIRE_SEEDED), (6, ol.WIRE_FULL)]):   # kinds alternate: each has its own captured graph
        wire = s.query_wire(idx, kind)
        want, _, _ = s.oracle_answer(ol.wire_expand(oracle, wire), Bbuf)
        packed = srv.answer_wire(wire)
        assert packed.nbytes == srv.lib.sb200_server_packed_response_bytes(srv.h)
        got = srv.unpack_response(packed)
        assert np.array_equal(got, want), f"response differs (query {rep})"
        assert np.array_equal(s.decode(got), s.pts[idx]), f"decode failed (query {rep})"
    # the plain and the wire entry interleave on one server
    q = s.query(3)
    want, _, _ = s.oracle_answer(q, Bbuf)
    assert np.array_equal(srv.answer(q), want)
    wire = s.query_wire(3)
    assert np.array_equal(srv.unpack_response(srv.answer_wire(wire)), s.oracle_answer(ol.wire_expand(oracle, wire), Bbuf)[0])
    srv.close()
    s.close()


def test_malformed_wire_query_fails_loudly(sb, oracle):
    s = ol.SpiralSession(oracle, "cfg1", 2, 2, seed=2)
    srv = SpiralServer(sb_params(s.prm))
    wire = s.query_wire(1)
    for bad in (wire[:-1].copy(), np.concatenate([wire, np.zeros(1, np.uint8)])):
        with pytest.raises(SB200Error, match="bytes"):
            srv.upload_query_wire(bad)
    bad = wire.copy(); bad[1] ^= 0x40
    with pytest.raises(SB200Error, match="magic"):
        srv.upload_query_wire(bad)
    bad = wire.copy(); bad[4] = 3
    with pytest.raises(SB200Error, match="kind"):
        srv.upload_query_wire(bad)
    assert sb.sb200_wire_query_bytes(0) == 0 and sb.sb200_wire_query_bytes(3) == 0
    srv.close()
    s.close()


def _db_words(srv):
    return srv.download(srv.lib.sb200_server_db_ptr(srv.h), srv.db_bytes // 8)


@pytest.mark.parametrize("cfg,nu1,nu2,world", [("cfg1", 3, 2, 1), ("cfg1", 4, 3, 2), ("cfg4", 2, 2, 1), ("cfg1", 2, 3, 4)])
def test_record_stream_builds_the_same_database(sb, oracle, cfg, nu1, nu2, world, tmp_path):
    """memory and file record streams == plaintext-item ingest, word for word, for whole and sharded servers
    (cfg4: 16 bits per coefficient)."""
    s = ol.SpiralSession(oracle, cfg, nu1, nu2, seed=4)
    rec = s.records()
    path = str(tmp_path / "records.bin")
    rec.tofile(path)
    for rank in range(world):
        a = SpiralServer(sb_params(s.prm), rank=rank, world=world)
        a.load_db_items(a.shard_items(s.pts).astype(np.uint16))
        want = _db_words(a)
        a.close()
        for source in (rec, path):
            b = SpiralServer(sb_params(s.prm), rank=rank, world=world)
            assert b.record_stream_bytes == rec.size
            b.load_db_records(source)
            assert np.array_equal(_db_words(b), want), f"rank {rank}, source {type(source).__name__}"
            b.close()
    srv = SpiralServer(sb_params(s.prm))
    with pytest.raises(SB200Error, match="record stream"):
        srv.load_db_records(rec[:-8].copy())
    with pytest.raises(SB200Error, match="cannot open"):
        srv.load_db_records(str(tmp_path / "missing.bin"))
    srv.close()
    s.close()


def test_snapshot_round_trip_and_checks(sb, oracle, tmp_path):
    s = ol.SpiralSession(oracle, "cfg1", 4, 2, seed=8)
    Bbuf = s.reference_db()
    path = str(tmp_path / "db.sb2d")
    a = SpiralServer(sb_params(s.prm))
    with pytest.raises(SB200Error, match="no loaded database"):
        a.save_db(path)
    a.load_db_items(s.pts.astype(np.uint16))
    a.save_db(path)
    want_db = _db_words(a)
    a.close()
    assert os.path.getsize(path) == 64 + want_db.nbytes
    b = SpiralServer(sb_params(s.prm))
    b.load_db_snapshot(path)
    assert np.array_equal(_db_words(b), want_db)
    b.set_public_params(s.W_left, s.W_right, s.W_conv, s.V_conv)
    q = s.query(17)
    got = b.answer(q)
    assert np.array_equal(got, s.oracle_answer(q, Bbuf)[0]) and np.array_equal(s.decode(got), s.pts[17])
    b.close()
    # another shape / shard must refuse the file
    other = ol.make_params("cfg1", 3, 3)
    c = SpiralServer(sb_params(other))
    with pytest.raises(SB200Error, match="another server"):
        c.load_db_snapshot(path)
    c.close()
    d = SpiralServer(sb_params(s.prm), rank=1, world=2)
    with pytest.raises(SB200Error, match="another server"):
        d.load_db_snapshot(path)
    d.close()
    # a flipped payload bit is caught by the integrity word, a truncated file by its length
    raw = np.fromfile(path, dtype=np.uint8)
    raw[64 + 12345] ^= 0x10
    raw.tofile(path)
    e = SpiralServer(sb_params(s.prm))
    with pytest.raises(SB200Error, match="integrity"):
        e.load_db_snapshot(path)
    with pytest.raises(SB200Error, match="not loaded"):
        e.scan()                                               # a rejected snapshot leaves the server without a database
    raw[:-8].tofile(path)
    with pytest.raises(SB200Error, match="truncated"):
        e.load_db_snapshot(path)
    e.close()
    s.close()


@pytest.mark.parametrize("cfg,nu1,nu2,world", [("cfg3", 4, 2, 1), ("cfg4", 3, 2, 2)])
def test_pack_server_records_snapshot_and_wire_query(sb, oracle, cfg, nu1, nu2, world, tmp_path):
    """Pack server: the record stream (out_n^2 polynomials per item) gives the planes of load_plane_items; for the whole
    server a FULL wire query through the expansion path equals the oracle's Pack pipeline; snapshot round trip."""
    prm = ol.make_params(cfg, nu1, nu2)
    rng = np.random.default_rng(77)
    dim0, num_per, n = 1 << nu1, 1 << nu2, prm.out_n
    planes = n * n
    pts, db = build_planes(oracle, prm, rng, dim0, num_per, planes)            # pts: [plane][item][N]
    dt = np.uint8 if prm.p_db == 256 else np.uint16
    rec = np.ascontiguousarray(pts.transpose(1, 0, 2).astype(dt).reshape(-1)).view(np.uint8)   # item-major, planes inside
    sp = SpiralParams(nu1, nu2, prm.t_gsw, prm.t_conv, prm.t_exp, prm.t_exp_right, prm.qp_bits, prm.out_n, prm.p_db)
    path = str(tmp_path / "pack.sb2d")
    for rank in range(world):
        a = PackServer(sp, rank=rank, world=world)
        for pl in range(planes):
            a.load_plane_items(pl, a.shard_items(pts[pl]).astype(np.uint16))
        want = a.db_words()
        a.save_db(path)
        a.close()
        b = PackServer(sp, rank=rank, world=world)
        assert b.lib.sb200_pack_server_record_stream_bytes(b.h) == rec.size
        b.load_db_records(rec)
        assert np.array_equal(b.db_words(), want), f"rank {rank}: records"
        b.close()
        c = PackServer(sp, rank=rank, world=world)
        c.load_db_snapshot(path)
        assert np.array_equal(c.db_words(), want), f"rank {rank}: snapshot"
        c.close()
    if world != 1:
        return
    g, stop = C.c_size_t(), C.c_size_t()
    oracle.so_pack_expansion_shape(C.byref(prm), C.byref(g), C.byref(stop))
    g, stop = g.value, stop.value
    vW = rnd_ntt(rng, n * (n + 1) * prm.t_conv)
    W_left, W_right, V = rnd_ntt(rng, g * 2 * prm.t_exp), rnd_ntt(rng, (stop + 1) * 2 * prm.t_exp_right), rnd_ntt(rng, 2 * 2 * prm.t_conv)
    query = rnd_ntt(rng, 2)
    wire = np.zeros(oracle.so_wire_query_bytes(ol.WIRE_FULL), dtype=np.uint8)
    oracle.so_wire_query_pack_full(p(query), ol.ptr8(wire))
    dummy = rnd_ntt(rng, 2)
    want = np.zeros((n + 1) * n * N, dtype=np.uint64)
    want_cts = np.zeros(planes * 2 * N, dtype=np.uint64)
    assert oracle.so_pack_answer(C.byref(prm), 1, p(query), p(W_left), p(W_right), p(V), p(dummy), p(dummy), p(vW), p(db), p(want), p(want_cts)) == 0
    srv = PackServer(sp)
    srv.load_db_records(rec)
    srv.set_public_params(W_left, W_right, V, vW)
    for _ in range(2):                                          # the second call replays the captured graph
        assert np.array_equal(srv.answer_wire(wire), want)
    assert np.array_equal(srv.answer(query), want)              # and the plain entry still works on the same server
    srv.close()
