"""CPU suite, part 1: the oracle (oracle/spiral_oracle.c) is pinned to the UNMODIFIED reference.

tests/golden/ref_digests_<cfg>.json holds FNV-1a digests of the reference's own outputs
(oracle/ref_golden.cpp linked against oracle/_ref/libspiral_ref_<cfg>_<isa>.so, both ISA builds)
for 26 seeded cases per parameter set; here the oracle recomputes every case and must match."""
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


@pytest.mark.parametrize("cfg", sorted(ol.CONFIGS))
def test_modswitch_integer_restatement_equals_x87(oracle, cfg):
    """modswitch (src/spiral.cpp:40-78) rounds twice in x87 extended precision before roundl.  The oracle restates it
    in integers (so a GPU can agree bit for bit); on an x86 host the literal long double statements are available too
    and must agree on the golden case's inputs, which plant coefficients that sit within 2^-45 of k + 1/2.  A large
    share of those round differently from exact integer rounding - otherwise the case would pin nothing."""
    import platform
    if platform.machine() not in ("x86_64", "AMD64", "i686"):
        pytest.skip("long double is not x87 extended precision on this host")
    g = ol.golden(cfg)
    prm = ol.make_params(cfg)
    case = ol.Case(oracle, g["cases"]["modswitch"]["id"], prm, g["seed"])
    qp = oracle.so_arb_qprime(prm.qp_bits)
    vals = [int(v) for v in case.inputs[0]]
    differs_from_exact = 0
    for k, v in enumerate(vals):
        a, b = oracle.so_modswitch_coeff(v, qp), oracle.so_modswitch_coeff_x87(v, qp)
        assert a == b, f"coefficient {v}: integer restatement {a} != x87 {b}"
        assert oracle.so_read_arbitrary_bits(ol.ptr(case.out), k * prm.qp_bits, prm.qp_bits) == a & ((1 << prm.qp_bits) - 1)
        differs_from_exact += a != (2 * v * qp + ol.Q) // (2 * ol.Q)
    assert differs_from_exact > 100
