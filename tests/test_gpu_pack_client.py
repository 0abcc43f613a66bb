"""GPU suite: the Pack servers on REAL encryptions - keys, packing keys and queries from the oracle's SpiralPack / SpiralStreamPack
client (oracle/client_sim.c, testHighRate's client statements).  The response must equal the oracle's whole Pack pipeline bit for
bit and decode to the planted item of every plane (testHighRate's "Is correct? : 1", src/testing.cpp:1131); the other Pack
tests use uniform ring elements, which pin the arithmetic but not the end-to-end property."""
import numpy as np
import pytest

from spiral_b200 import SpiralParams
from spiral_b200.server import PackServer
from tests import oracle_lib as ol

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg,nu1,nu2,direct", [("cfg3", 4, 2, False), ("cfg1", 5, 1, False), ("cfg4", 3, 2, True), ("cfg4", 5, 3, True)])
def test_pack_server_on_real_encryptions(sb, oracle, cfg, nu1, nu2, direct):
    s = ol.PackSession(oracle, cfg, nu1, nu2, direct, seed=6)
    p = s.prm
    db = s.reference_planes()
    srv = PackServer(SpiralParams(nu1, nu2, p.t_gsw, p.t_conv, p.t_exp, p.t_exp_right, p.qp_bits, p.out_n, p.p_db))
    for pl in range(s.planes):
        srv.load_plane_items(pl, np.ascontiguousarray(s.pts[pl].astype(np.uint16)))
    srv.set_public_params(s.W_left, s.W_right, s.V, s.v_W)
    for idx in (0, s.total_n - 1, s.total_n // 2 + 1):
        q, v_first, v_fold = s.query(idx)
        want, want_cts = s.oracle_answer((q, v_first, v_fold), db)
        got, got_cts = srv.answer_direct(v_first, v_fold, want_cts=True) if direct else srv.answer(q, want_cts=True)
        assert np.array_equal(got_cts, want_cts), f"folded per-plane ciphertexts differ at idx {idx}"
        assert np.array_equal(got, want), f"response differs at idx {idx}"
        assert np.array_equal(s.decode(got), s.planted(idx)), f"decode failed at idx {idx}"
    srv.close()
    s.close()


SEED = bytes(range(32))


@pytest.mark.parametrize("cfg,nu1,nu2,direct", [("cfg3", 4, 2, False), ("cfg1", 5, 1, False), ("cfg4", 3, 2, True), ("cfg4", 5, 3, True)])
def test_gpu_pack_client_matches_oracle_statement(sb, oracle, cfg, nu1, nu2, direct):
    """The Pack-variant client on the GPU (sb200_pack_client_*) against its plain-C statement (so_pack_client_new_chacha): secret
    keys, packing keys, expansion keys, V, the packed wire query or the direct-upload ciphertexts, and decoding agree bit for bit."""
    from spiral_b200.client import PackClient
    s = ol.PackSession(oracle, cfg, nu1, nu2, direct, seed=6, chacha_seed=SEED)
    p = s.prm
    c = PackClient(SpiralParams(nu1, nu2, p.t_gsw, p.t_conv, p.t_exp, p.t_exp_right, p.qp_bits, p.out_n, p.p_db), SEED)
    sr, Sp = c.secret()
    want_sr, want_Sp = np.zeros(ol.N, dtype=np.uint64), np.zeros(p.out_n * ol.N, dtype=np.uint64)
    oracle.so_client_secret_n(s.client, ol.ptr(want_sr), ol.ptr(want_Sp), p.out_n)
    assert np.array_equal(sr, want_sr) and np.array_equal(Sp, want_Sp), "secret keys differ"
    got = c.public_params(direct=direct)
    for name, g_, w_ in zip(("W_exp_left", "W_exp_right", "V", "v_W"), got, (s.W_left, s.W_right, s.V, s.v_W)):
        if w_ is None:
            assert g_ is None
            continue
        assert g_.size == w_.size, name
        assert np.array_equal(g_, ol.canon(w_, ol.KIND_NTT)), f"{name} differs"
    db = s.reference_planes()
    for qid, idx in enumerate((0, s.total_n - 1, s.total_n // 2 + 1)):
        if direct:
            if qid == 0:                                  # the oracle's direct query has no query id: block 0
                v_first, v_fold = c.query_direct(idx, 0)
                _, w_first, w_fold = s.query(idx)
                assert np.array_equal(v_first, ol.canon(w_first, ol.KIND_NTT)), "direct first-dimension ciphertexts differ"
                assert np.array_equal(v_fold[:w_fold.size], ol.canon(w_fold, ol.KIND_NTT)), "direct GSW ciphertexts differ"
            else:
                v_first, v_fold = c.query_direct(idx, qid)
            resp, _ = s.oracle_answer((None, v_first, v_fold), db)
        else:
            wseed = bytes([23 * qid + 5] * 32)
            wire = c.query_wire(idx, qid, wseed)
            assert np.array_equal(wire, s.chacha_query_wire(idx, qid, wseed)), f"wire query {qid} differs"
            resp, _ = s.oracle_answer((ol.wire_expand(oracle, wire), None, None), db)
        dec = c.decode(resp)
        assert np.array_equal(dec, s.decode(resp)), "GPU decode differs from the oracle"
        assert np.array_equal(dec, s.planted(idx)), f"record {idx} not recovered"
    c.close()
    s.close()


@pytest.mark.parametrize("cfg,nu1,nu2,direct", [("cfg3", 5, 2, False), ("cfg4", 5, 3, True)])
def test_all_gpu_pack_round_trip(sb, cfg, nu1, nu2, direct):
    """No oracle on the data path: GPU Pack client keys + query, GPU Pack server, GPU decode == the planted records."""
    from spiral_b200.client import PackClient
    d = ol.CONFIGS[cfg]
    sp = SpiralParams(nu1, nu2, d["t_gsw"], d["t_conv"], d["t_exp"], d["t_exp_right"], d["qp_bits"], d["out_n"], d["p_db"])
    rng = np.random.default_rng(77)
    total, planes = 1 << (nu1 + nu2), d["out_n"] ** 2
    pts = rng.integers(0, d["p_db"], size=(planes, total, ol.N), dtype=np.uint64)
    client = PackClient(sp, bytes([3] * 32))
    srv = PackServer(sp)
    for pl in range(planes):
        srv.load_plane_items(pl, np.ascontiguousarray(pts[pl].astype(np.uint16)))
    srv.set_public_params(*client.public_params(direct=direct))
    for qid, idx in enumerate((0, total - 1, 77 % total)):
        if direct:
            resp = srv.answer_direct(*client.query_direct(idx, qid))
        else:
            resp = srv.answer_wire(client.query_wire(idx, qid))
        assert np.array_equal(client.decode(resp), pts[:, idx, :]), f"record {idx} not recovered"
    srv.close()
    client.close()
