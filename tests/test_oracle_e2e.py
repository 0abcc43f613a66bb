"""CPU suite, part 3: the oracle's whole server pipeline (so_spiral_answer) composed with the test
client decodes to the planted record - the reference's own end-to-end gate ("Is correct?: 1")."""
import numpy as np
import pytest

from tests import oracle_lib as ol


@pytest.mark.parametrize("cfg,nu1,nu2,idx", [("cfg1", 2, 2, 5), ("cfg1", 4, 1, 17), ("cfg5", 3, 2, 30)])
def test_oracle_pipeline_decodes_planted_record(oracle, cfg, nu1, nu2, idx):
    s = ol.SpiralSession(oracle, cfg, nu1, nu2, seed=3)
    Bbuf = s.reference_db()
    resp, _, _ = s.oracle_answer(s.query(idx), Bbuf)
    got = s.decode(resp)
    assert np.array_equal(got, s.pts[idx]), "decoded record differs from the planted one"
    s.close()


@pytest.mark.parametrize("cfg,nu1,nu2,direct", [("cfg3", 4, 2, False), ("cfg1", 5, 1, False), ("cfg4", 3, 2, True)])
def test_pack_client_and_server_statements_decode_the_planted_items(oracle, cfg, nu1, nu2, direct):
    """testHighRate's "Is correct? : 1" (src/testing.cpp:1131) on the oracle's Pack client + Pack server: the out_n^2 decoded
    polynomials are the planted plaintexts of every plane at the queried index (packed query + expansion, and direct upload)."""
    s = ol.PackSession(oracle, cfg, nu1, nu2, direct, seed=4)
    db = s.reference_planes()
    for idx in (0, s.total_n - 1, s.total_n // 2 + 1):
        resp, _ = s.oracle_answer(s.query(idx), db)
        assert np.array_equal(s.decode(resp), s.planted(idx)), f"idx {idx}"
    s.close()
