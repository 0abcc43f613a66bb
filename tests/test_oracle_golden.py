"""CPU suite, part 1: the oracle (oracle/spiral_oracle.c) is pinned to the UNMODIFIED reference.

tests/golden/ref_digests_<cfg>.json holds FNV-1a digests of the reference's own outputs
(oracle/ref_golden.cpp linked against oracle/_ref/libspiral_ref_<cfg>_<isa>.so, both ISA builds)
for 25 seeded cases per parameter set; here the oracle recomputes every case and must match."""
import pytest

from tests import oracle_lib as ol


@pytest.mark.parametrize("cfg", sorted(ol.CONFIGS))
def test_tables_match_reference(oracle, cfg):
    g = ol.golden(cfg)
    assert g["tables_match_reference"] is True
    assert f"{oracle.so_fnv1a64(oracle.so_tables(), 8 * ol.N):016x}" == g["tables_digest"]


@pytest.mark.parametrize("cfg", sorted(ol.CONFIGS))
def test_oracle_digests_match_reference(oracle, cfg):
    g = ol.golden(cfg)
    prm = ol.make_params(cfg)
    for k in ("t_gsw", "t_conv", "t_exp", "t_exp_right", "qp_bits", "out_n", "p_db"):
        assert g["params"][k] == ol.CONFIGS[cfg][k]
    assert oracle.so_case_count() == len(g["cases"])
    bad = []
    for name, meta in g["cases"].items():
        case = ol.Case(oracle, meta["id"], prm, g["seed"])
        assert case.name == name
        if case.out.size != meta["words"] or f"{case.digest:016x}" != meta["digest"]:
            bad.append(name)
    assert not bad, f"oracle disagrees with the reference on {bad}"


def test_case_inputs_are_seed_dependent(oracle):
    prm = ol.make_params("cfg1")
    a = ol.Case(oracle, 3, prm, 1)
    b = ol.Case(oracle, 3, prm, 2)
    assert a.digest != b.digest
