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
