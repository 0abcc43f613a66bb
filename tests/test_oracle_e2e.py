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
