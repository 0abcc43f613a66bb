"""GPU suite: the client on the GPU against its plain-C statement (oracle/client_sim.c, so_client_new_chacha):
secret keys, public parameters, wire queries and decoded records agree bit for bit; and the all-GPU loop - GPU client ->
wire query -> GPU server -> packed response -> GPU decode - returns the planted record (the reference's "Is correct?: 1")."""
import numpy as np
import pytest

from spiral_b200 import SpiralParams
from spiral_b200.client import SpiralClient
from spiral_b200.lib import SB200Error, check
from spiral_b200.server import SpiralServer
from tests import oracle_lib as ol

pytestmark = pytest.mark.gpu
SEED = bytes(range(32))


def sb_params(so):
    return SpiralParams(so.nu1, so.nu2, so.t_gsw, so.t_conv, so.t_exp, so.t_exp_right, so.qp_bits, so.out_n, so.p_db)


def test_gaussian_thresholds_match_the_oracle(sb, oracle):
    s = ol.SpiralSession(oracle, "cfg1", 2, 2, chacha_seed=SEED)
    want = np.zeros(128, dtype=np.uint64)
    oracle.so_client_gaussian_thresholds(s.client, ol.ptr(want))
    got = np.zeros(128, dtype=np.uint64)
    check(sb.sb200_client_gaussian_thresholds(ol.ptr(got)), sb)
    assert np.array_equal(got, want)
    assert np.all(np.diff(got.astype(np.int64)) >= 0) and got[0] == 1 and abs(int(got[-1]) - (1 << 53)) < 64     # cdf[127] = 1 up to rounding
    s.close()


# stopround == 0 (cfg1 3,2: 16 GSW bits > 8 first-dimension slots), stopround != 0, odd t_GSW, a larger first dimension
@pytest.mark.parametrize("cfg,nu1,nu2", [("cfg1", 3, 2), ("cfg1", 6, 2), ("cfg5", 4, 3), ("cfg1", 7, 3)])
def test_client_matches_oracle_statement(sb, oracle, cfg, nu1, nu2):
    s = ol.SpiralSession(oracle, cfg, nu1, nu2, seed=3, chacha_seed=SEED)
    c = SpiralClient(sb_params(s.prm), SEED)
    sr, Sp = c.secret()
    want_sr, want_Sp = s.secret()
    assert np.array_equal(sr, want_sr) and np.array_equal(Sp, want_Sp), "secret keys differ"
    for name, got, want in zip(("W_exp_left", "W_exp_right", "W_conv", "V_conv"), c.public_params(), (s.W_left, s.W_right, s.W_conv, s.V_conv)):
        assert got.size == want.size, name
        assert np.array_equal(got, ol.canon(want, ol.KIND_NTT)), f"{name} differs"
    for qid, idx in enumerate((0, s.total_n - 1, s.total_n // 2 + 1)):
        wseed = bytes([17 * qid + 1] * 32)
        assert np.array_equal(c.query_wire(idx, qid, wseed), s.chacha_query_wire(idx, qid, wseed)), f"wire query {qid} differs"
    # decoding: a real response (oracle server) and an arbitrary one (uniform rows in range) decode identically
    Bbuf = s.reference_db()
    resp, _, _ = s.oracle_answer(ol.wire_expand(oracle, s.chacha_query_wire(5, 9, bytes([5] * 32))), Bbuf)
    assert np.array_equal(c.decode(resp), s.decode(resp)) and np.array_equal(c.decode(resp), s.pts[5])
    rng = np.random.default_rng(1)
    qp = oracle.so_arb_qprime(s.prm.qp_bits)
    rnd = np.concatenate([rng.integers(0, qp, 2 * ol.N, dtype=np.uint64), rng.integers(0, 4 * s.prm.p_db, 4 * ol.N, dtype=np.uint64)])
    assert np.array_equal(c.decode(rnd), s.decode(rnd))
    c.close()
    s.close()


@pytest.mark.parametrize("cfg,nu1,nu2", [("cfg1", 5, 3), ("cfg5", 6, 2)])
def test_all_gpu_round_trip(sb, oracle, cfg, nu1, nu2):
    """No oracle on the data path: GPU client keys + wire query, GPU server on a record stream, GPU decode."""
    prm = ol.make_params(cfg, nu1, nu2)
    sp = sb_params(prm)
    rng = np.random.default_rng(42)
    total = 1 << (nu1 + nu2)
    pts = rng.integers(0, prm.p_db, size=(total, 4, ol.N), dtype=np.uint64)
    client = SpiralClient(sp, bytes([7] * 32))
    srv = SpiralServer(sp)
    srv.load_db_records(np.ascontiguousarray(pts.astype(np.uint8).reshape(-1)))
    srv.set_public_params(*client.public_params())
    for qid, idx in enumerate((0, total - 1, 77 % total, total // 2)):
        wire = client.query_wire(idx, qid, bytes([qid + 100] * 32))
        got = client.decode(srv.unpack_response(srv.answer_wire(wire)))
        assert np.array_equal(got, pts[idx]), f"record {idx} not recovered"
    with pytest.raises(SB200Error, match="outside"):
        client.query_wire(total, 0, bytes(32))
    with pytest.raises(SB200Error, match="query_id"):
        client.query_wire(0, 1 << 24, bytes(32))
    srv.close()
    client.close()


def test_client_rejects_impossible_shapes(sb):
    from spiral_b200.lib import load_library
    with pytest.raises(SB200Error, match="query slots"):
        SpiralClient(SpiralParams(11, 4, 8, 4, 8, 56, 20, 2, 256), SEED)      # 2048 + 32 slots do not fit one polynomial
    assert load_library() is sb


def test_wire_seed_is_derived_per_query(sb, oracle):
    """row 0 = -a must be fresh per query.  With no explicit seed the client derives it from (client key, query_id):
    one ChaCha20 block, nonce {"SB2C", CC_QUERY | query_id, "wsee"} - checked against the oracle's RFC 8439 block function."""
    import ctypes as C
    prm = ol.make_params("cfg1", 4, 2)
    c = SpiralClient(sb_params(prm), SEED)
    seeds = [c.wire_seed(q) for q in (0, 1, 2, (1 << 24) - 1)]
    assert len(set(seeds)) == 4 and c.wire_seed(1) == seeds[1]
    for q, got in zip((0, 1, 2, (1 << 24) - 1), seeds):
        out = (C.c_uint32 * 16)()
        oracle.so_chacha20_block((C.c_uint32 * 8)(*np.frombuffer(SEED, dtype="<u4")), 0,
                                 (C.c_uint32 * 3)(0x43324253, (6 << 24) | q, 0x65657377), out)
        assert bytes(np.array(out[:8], dtype="<u4").tobytes()) == got
    a, b = c.query_wire(5, 1), c.query_wire(5, 2)
    assert np.array_equal(a, c.query_wire(5, 1, seeds[1])) and np.array_equal(a, c.query_wire(5, 1))
    assert not np.array_equal(a[8:40], b[8:40]), "two queries share row 0"
    other = SpiralClient(sb_params(prm), bytes([9] * 32))
    assert other.wire_seed(1) != seeds[1]
    with pytest.raises(SB200Error, match="query_id"):
        c.wire_seed(1 << 24)
    other.close()
    c.close()


# response moduli of 30+ bits (q_prime_bits 31 / 32 occur in the reference's parameter files; the table goes to 36):
# 64 products of ~qp^2 no longer fit 64 bits, the decode sums in 128 bits like the oracle
@pytest.mark.parametrize("qp_bits", [29, 31, 32, 36])
def test_decode_with_wide_response_moduli(sb, oracle, qp_bits):
    cfg = dict(ol.CONFIGS["cfg1"], qp_bits=qp_bits)
    s = ol.SpiralSession(oracle, cfg, 3, 2, seed=4, chacha_seed=SEED)
    c = SpiralClient(sb_params(s.prm), SEED)
    Bbuf = s.reference_db()
    for idx in (0, 9, s.total_n - 1):
        resp, _, _ = s.oracle_answer(ol.wire_expand(oracle, s.chacha_query_wire(idx, idx, bytes([idx + 1] * 32))), Bbuf)
        got = c.decode(resp)
        assert np.array_equal(got, s.decode(resp)), f"qp_bits {qp_bits}: GPU decode differs from the oracle"
        assert np.array_equal(got, s.pts[idx]), f"qp_bits {qp_bits}: record {idx} not recovered"
    rng = np.random.default_rng(qp_bits)
    qp = oracle.so_arb_qprime(qp_bits)
    rnd = np.concatenate([rng.integers(0, qp, 2 * ol.N, dtype=np.uint64), rng.integers(0, 4 * s.prm.p_db, 4 * ol.N, dtype=np.uint64)])
    assert np.array_equal(c.decode(rnd), s.decode(rnd))
    c.close()
    s.close()


def test_servers_reject_parameters_they_cannot_serve(sb):
    import ctypes as C
    for bad in (SpiralParams(4, 2, 8, 4, 8, 56, 13, 2, 256), SpiralParams(4, 2, 8, 4, 8, 56, 40, 2, 256),
                SpiralParams(4, 2, 8, 4, 8, 56, 20, 2, 0), SpiralParams(4, 2, 8, 4, 8, 56, 20, 2, 1 << 17), SpiralParams(11, 4, 8, 4, 8, 56, 20, 2, 256)):
        h = C.c_void_p()
        assert sb.sb200_server_create(C.byref(h), C.byref(bad), 0, 0, 1) == -3, "bad parameters accepted"
    for bad in (SpiralParams(4, 2, 8, 4, 8, 56, 13, 2, 256), SpiralParams(4, 2, 8, 4, 8, 56, 20, 2, 0)):
        h = C.c_void_p()
        assert sb.sb200_pack_server_create(C.byref(h), C.byref(bad), 0) == -3
    with pytest.raises(SB200Error, match="response modulus"):
        SpiralClient(SpiralParams(4, 2, 8, 4, 8, 56, 13, 2, 256), SEED)
