"""CPU suite: the B200 shape model emits its choices in the form the reference's tooling reads - the sectioned
"/* name */ + JSON" text of all_parameter_choices.txt with "(log2 records, record bytes)" -> scheme -> {"params": {...}} entries
(all_parameter_choices.txt:1-80; select_params.py / run_all.py look entries up by that key and read nu_1, nu_2, p, q_prime_bits,
t_GSW, t_conv, t_exp, t_exp_right from "params")."""
import importlib.util
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("cost_model_b200", os.path.join(ROOT, "scripts", "cost_model_b200.py"))
cm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(cm)

PARAM_KEYS = {"nu_1", "nu_2", "p", "q_prime_bits", "query_size", "s_e", "t_GSW", "t_conv", "t_exp", "t_exp_right"}


def _entry(nu1, nu2, t_gsw, qp):
    return {"spiral": {"params": {"nu_1": nu1, "nu_2": nu2, "p": 256, "q_prime_bits": qp, "query_size": 14336.0, "s_e": 87.7,
                                  "t_GSW": t_gsw, "t_conv": 4, "t_exp": 8, "t_exp_right": 56}},
            "spiralstream": {"params": {"nu_1": 9, "nu_2": 6, "p": 256, "q_prime_bits": 20, "t_GSW": 4, "t_conv": 32, "t_exp": 2, "t_exp_right": 56}}}


def test_emitted_choices_have_the_reference_format_and_keep_the_database():
    model = json.load(open(os.path.join(ROOT, "profiles", "r01_cost_model_b200.json")))
    table = {"(20, 256)": _entry(8, 7, 8, 20), "(18, 30000)": _entry(9, 9, 9, 21), "(16, 256)": _entry(5, 6, 8, 20),
             "(12, 1)": {"spiral": {"params": dict(_entry(8, 7, 8, 20)["spiral"]["params"], t_GSW=5)}}}        # unknown gadget set: skipped
    out = cm.emit_choices(model, table)
    assert set(out) == {"(20, 256)", "(18, 30000)", "(16, 256)"}
    for key, ent in out.items():
        assert re.fullmatch(r"\(\d+, \d+\)", key)
        p, ref = ent["spiral"]["params"], table[key]["spiral"]["params"]
        assert set(p) == PARAM_KEYS
        assert p["nu_1"] + p["nu_2"] == ref["nu_1"] + ref["nu_2"], "the record count must not change"
        assert all(p[k] == ref[k] for k in PARAM_KEYS - {"nu_1", "nu_2"}), "gadget lengths and moduli are copied, only the split moves"
        assert ent["spiral"]["b200_model_us"] <= ent["spiral"]["reference_choice"]["b200_model_us"] + 1e-9
    # (16, 256) = 2^11 matrices with cfg1's gadgets: the reference's CPU-shaped split (5, 6) is not the B200 optimum
    assert (out["(16, 256)"]["spiral"]["params"]["nu_1"], out["(16, 256)"]["spiral"]["params"]["nu_2"]) != (5, 6) or \
        out["(16, 256)"]["spiral"]["b200_model_us"] == out["(16, 256)"]["spiral"]["reference_choice"]["b200_model_us"]
    # certified bounds are respected
    assert out["(20, 256)"]["spiral"]["params"]["nu_1"] <= 8 and out["(20, 256)"]["spiral"]["params"]["nu_2"] <= 7


def test_sectioned_text_round_trips():
    text = "/* Table */\n" + json.dumps({"(20, 256)": _entry(8, 7, 8, 20)}, indent=4) + "\n/* Asymp comp */\n{}\n" + json.dumps({"wiki": _entry(9, 11, 10, 22)}) + "\n"
    secs = cm.parse_choices(text)
    assert [n for n, _ in secs] == ["Table", "Asymp comp", "Asymp comp"]
    assert secs[0][1]["(20, 256)"]["spiral"]["params"]["t_exp_right"] == 56 and secs[2][1]["wiki"]["spiral"]["params"]["t_GSW"] == 10


def test_committed_b200_choices_parse_and_match_the_model():
    path = os.path.join(ROOT, "profiles", "b200_parameter_choices.txt")
    model = json.load(open(os.path.join(ROOT, "profiles", "r01_cost_model_b200.json")))
    secs = cm.parse_choices(open(path).read())
    assert secs and secs[0][0] == "Table"
    for _, table in secs:
        for key, ent in table.items():
            p = ent["spiral"]["params"]
            assert set(p) == PARAM_KEYS
            g = [g for g in cm.GADGETS.values() if g["t_gsw"] == p["t_GSW"] and g["qp_bits"] == p["q_prime_bits"]][0]
            best = cm.best_split(model, p["nu_1"] + p["nu_2"], g)
            assert (best[1], best[2]) == (p["nu_1"], p["nu_2"]), key
