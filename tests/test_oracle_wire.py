"""CPU tests of the oracle's statement of the wire / on-disk formats (oracle/wire_format.c): pinned by the RFC 8439
ChaCha20 block vector, by the reference's own bit I/O and by the end-to-end property the reference checks
(decoded record == planted record, src/spiral.cpp:1494) for queries that went through the wire."""
import ctypes as C

import numpy as np
import pytest

from tests import oracle_lib as ol


@pytest.fixture(scope="module")
def lib():
    return ol.load()


def test_chacha20_block_rfc8439(lib):
    """RFC 8439 section 2.3.2: key 00..1f, nonce 00:00:00:09 00:00:00:4a 00:00:00:00, block counter 1."""
    key = (C.c_uint32 * 8)(*[int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)])
    nonce = (C.c_uint32 * 3)(0x09000000, 0x4A000000, 0)
    out = (C.c_uint32 * 16)()
    lib.so_chacha20_block(key, 1, nonce, out)
    want = [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
            0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]
    assert list(out) == want


def test_chacha20_block_against_cryptography(lib):
    """Independent implementation (OpenSSL through `cryptography`), random key / counter / nonce."""
    algorithms = pytest.importorskip("cryptography.hazmat.primitives.ciphers.algorithms")
    from cryptography.hazmat.primitives.ciphers import Cipher
    rng = np.random.default_rng(3)
    for _ in range(8):
        key = rng.integers(0, 256, 32, dtype=np.uint8).tobytes()
        nonce = rng.integers(0, 256, 12, dtype=np.uint8).tobytes()
        ctr = int(rng.integers(0, 1 << 32))
        ks = Cipher(algorithms.ChaCha20(key, ctr.to_bytes(4, "little") + nonce), mode=None).encryptor().update(bytes(64))
        out = (C.c_uint32 * 16)()
        lib.so_chacha20_block((C.c_uint32 * 8)(*np.frombuffer(key, dtype="<u4")), ctr, (C.c_uint32 * 3)(*np.frombuffer(nonce, dtype="<u4")), out)
        assert list(out) == list(np.frombuffer(ks, dtype="<u4"))


def test_seeded_row0_is_canonical_and_seed_dependent(lib):
    rows = []
    for s in (1, 2):
        seed = np.full(32, s, dtype=np.uint8)
        row = np.zeros(2 * ol.N, dtype=np.uint64)
        lib.so_wire_seeded_row0(ol.ptr8(seed), ol.ptr(row))
        assert row[:ol.N].max() < ol.P and row[ol.N:].max() < ol.B
        # uniform residues: the mean of 2048 draws sits within 5 sigma of q/2
        for part, q in ((row[:ol.N], ol.P), (row[ol.N:], ol.B)):
            assert abs(part.astype(np.float64).mean() - q / 2) < 5 * q / np.sqrt(12 * ol.N)
        rows.append(row)
    assert not np.array_equal(rows[0], rows[1])


def test_full_wire_round_trip(lib):
    s = ol.SpiralSession(lib, "cfg1", 2, 2, seed=5)
    q = s.query(3)
    wire = np.zeros(lib.so_wire_query_bytes(ol.WIRE_FULL), dtype=np.uint8)
    lib.so_wire_query_pack_full(ol.ptr(q), ol.ptr8(wire))
    assert wire.size == 8 + 2 * 14336                   # 2 * b_per_elem (src/spiral.cpp:219-225) + header
    back = ol.wire_expand(lib, wire)
    assert np.array_equal(ol.canon(back, ol.KIND_NTT), ol.canon(q, ol.KIND_NTT))
    s.close()


@pytest.mark.parametrize("kind,size", [(ol.WIRE_SEEDED, 8 + 32 + 14336), (ol.WIRE_FULL, 8 + 2 * 14336)])
def test_wire_query_decodes_to_planted_record(lib, kind, size):
    s = ol.SpiralSession(lib, "cfg1", 2, 2, seed=7)
    Bbuf = s.reference_db()
    idx = 9
    wire = s.query_wire(idx, kind)
    assert wire.size == size
    q = ol.wire_expand(lib, wire)
    assert q is not None
    resp, _, _ = s.oracle_answer(q, Bbuf)
    assert np.array_equal(s.decode(resp), s.pts[idx])
    s.close()


def test_malformed_wire_is_rejected(lib):
    s = ol.SpiralSession(lib, "cfg1", 2, 2, seed=7)
    wire = s.query_wire(1, ol.WIRE_SEEDED)
    bad = wire.copy(); bad[0] ^= 1
    assert ol.wire_expand(lib, bad) is None               # magic
    bad = wire.copy(); bad[4] = 9
    assert ol.wire_expand(lib, bad) is None               # kind
    assert ol.wire_expand(lib, wire[:-1].copy()) is None  # length
    s.close()


def test_records_to_plaintexts(lib):
    rng = np.random.default_rng(11)
    for p_db, dt in ((256, np.uint8), (65536, np.uint16), (16, None)):
        n = 4096
        vals = rng.integers(0, p_db, n, dtype=np.uint64)
        if dt is not None:
            rec = vals.astype(dt).view(np.uint8)
        else:                                              # 4 bits per coefficient: two per byte, low nibble first
            rec = (vals[0::2] | (vals[1::2] << np.uint64(4))).astype(np.uint8)
        out = np.zeros(n, dtype=np.uint64)
        lib.so_records_to_plaintexts(ol.ptr(out), ol.ptr8(np.ascontiguousarray(rec)), n, p_db)
        assert np.array_equal(out, vals)


def test_counter_based_client_end_to_end(lib):
    """The ChaCha-stream client (the statement the CUDA client is compared with): keys, public parameters and a seeded wire
    query drive the oracle's server to a response that decodes to the planted record; secrets are small and seed-dependent."""
    seed = bytes(range(32))
    s = ol.SpiralSession(lib, "cfg1", 3, 2, seed=5, chacha_seed=seed)
    sr, Sp = s.secret()
    centred = np.where(sr > ol.Q // 2, sr.astype(np.int64) - ol.Q, sr.astype(np.int64))
    assert np.abs(centred).max() <= 64 and 1.5 < centred.std() < 3.5          # width 6.4 <-> sigma = 6.4 / sqrt(2 pi) = 2.55
    Bbuf = s.reference_db()
    for qid, idx in enumerate((0, 13, s.total_n - 1)):
        wire = s.chacha_query_wire(idx, qid, bytes([qid + 1] * 32))
        resp, _, _ = s.oracle_answer(ol.wire_expand(lib, wire), Bbuf)
        assert np.array_equal(s.decode(resp), s.pts[idx])
    # the same (wire seed, query id) gives the same bytes; another query id changes only the noise, i.e. row 1
    a = s.chacha_query_wire(5, 7, bytes([9] * 32)); b = s.chacha_query_wire(5, 7, bytes([9] * 32)); c = s.chacha_query_wire(5, 8, bytes([9] * 32))
    assert np.array_equal(a, b) and np.array_equal(a[:40], c[:40]) and not np.array_equal(a, c)
    t = ol.SpiralSession(lib, "cfg1", 3, 2, seed=5, chacha_seed=bytes(range(1, 33)))
    assert not np.array_equal(t.secret()[0], sr)
    s.close(); t.close()


@pytest.mark.parametrize("cfg,nu1,nu2,direct", [("cfg3", 4, 2, False), ("cfg4", 3, 2, True)])
def test_counter_based_pack_client_end_to_end(lib, cfg, nu1, nu2, direct):
    """The ChaCha-stream SpiralPack / SpiralStreamPack client: packing keys, expansion keys + V, a seeded wire query (or the
    direct-upload ciphertexts) drive the oracle's Pack server to a response that decodes to the planted item of every plane."""
    s = ol.PackSession(lib, cfg, nu1, nu2, direct, seed=8, chacha_seed=bytes(range(32)))
    db = s.reference_planes()
    for qid, idx in enumerate((0, s.total_n - 1, 9)):
        if direct:
            query = s.query(idx)
        else:
            wire = s.chacha_query_wire(idx, qid, bytes([qid + 3] * 32))
            assert wire.size == 8 + 32 + 14336
            query = (ol.wire_expand(lib, wire), None, None)
        resp, _ = s.oracle_answer(query, db)
        assert np.array_equal(s.decode(resp), s.planted(idx)), f"idx {idx}"
    s.close()


def test_known_answers_of_the_formats(lib):
    """tests/golden/wire_kat.json (scripts/make_wire_golden.py): the byte-level behaviour of the wire query, the seed expansion,
    the counter-based client and the record unpacking is frozen - a format is a contract between machines."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.join(ol.ROOT, "scripts"))
    import make_wire_golden
    with open(os.path.join(ol.GOLDEN_DIR, "wire_kat.json")) as f:
        want = json.load(f)
    got = json.loads(json.dumps(make_wire_golden.vectors(lib)))
    assert got == want
