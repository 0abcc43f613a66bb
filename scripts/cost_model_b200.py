"""Parameter selection re-fitted for B200 timings.

The reference picks (nu_1, nu_2) and the gadget lengths by minimising a cost model whose coefficients were
regressed on ITS CPU timings (select_params.py:179-198: folding = 1000*(33 + 29.6*t_GSW)*2^nu_2/2^6 us,
first dimension = 619*2^nu_2 + 9.26*2^(nu_1+nu_2) us, ...).  On a B200 the first dimension is ~1000x faster
while the latency-bound expansion / fold chains shrink far less, so the optimum moves.  This script

  measure : times the three server stages on the GPU for a grid of database shapes     (GPU box, writes JSON)
  fit     : regresses the same functional form on those timings                         (anywhere)
  select  : ranks the shapes that hold a given database by the fitted model             (anywhere)

Feasibility (the noise budget) is NOT re-derived: a shape is admitted only if it does not exceed, in either
dimension, a (nu_1, nu_2) the reference itself certified for the same gadget parameters
(all_parameter_choices.txt: noise grows with both dimensions, so smaller shapes with the same gadgets are safe).

    python scripts/cost_model_b200.py measure --out gpurun_out/shape_sweep.json
    python scripts/cost_model_b200.py fit profiles/r01_shape_sweep.json --out profiles/r01_cost_model_b200.json
    python scripts/cost_model_b200.py select profiles/r01_cost_model_b200.json --log-items 15
    python scripts/cost_model_b200.py emit profiles/r01_cost_model_b200.json /path/to/all_parameter_choices.txt --out profiles/b200_parameter_choices.txt
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# gadget sets the reference certified for p = 256, each with the largest shape it used them at (all_parameter_choices.txt,
# "spiral" entries): (20, 256) -> nu (8,7), t_GSW 8, q' 2^20 [cfg1]; (18, 30000) -> nu (9,9), t_GSW 9, q' 2^21 [cfg5];
# "wiki" -> nu (9,11), t_GSW 10, q' 2^22
GADGETS = {
    "cfg1": dict(t_gsw=8, t_conv=4, t_exp=8, t_exp_right=56, qp_bits=20, out_n=2, p_db=256, max_nu1=8, max_nu2=7),
    "cfg5": dict(t_gsw=9, t_conv=4, t_exp=8, t_exp_right=56, qp_bits=21, out_n=2, p_db=256, max_nu1=9, max_nu2=9),
    "wiki": dict(t_gsw=10, t_conv=4, t_exp=8, t_exp_right=56, qp_bits=22, out_n=2, p_db=256, max_nu1=9, max_nu2=11),
}
N = 2048


def ceil_log2(x):
    return max(0, (x - 1).bit_length())


def shape_features(nu1, nu2, g):
    """Regressors of one shape: expansion rounds and size, scan bytes, fold rounds and size."""
    ell_bits = g["t_gsw"] * nu2
    rounds = ceil_log2(ell_bits + (1 << nu1))
    nosplit = float(ell_bits > (1 << nu1))            # stopround == 0: every round carries the 56-digit right-hand key switches
    return {
        # launch chain, rounds, first-dimension ciphertexts, GSW bits, and the unsplit tree's 2^rounds wide key switches
        "exp": [1.0, float(rounds), float(1 << nu1), float(ell_bits), nosplit * float(1 << rounds)],
        # fixed + bytes / bandwidth, one bandwidth per scan tiling (launch_scan_spiral): >= 128, 64, < 64 database columns per z
        "scan": [1.0] + [float(8 * N * 4 * (1 << (nu1 + nu2))) if cls else 0.0 for cls in (nu2 >= 6, nu2 == 5, nu2 < 5)],
        "fold": [1.0, float(nu2), float(g["t_gsw"] * (1 << nu2))],                 # fixed, rounds, digit NTTs
    }


def feasible(nu1, nu2, g):
    return (1 << nu1) + g["t_gsw"] * nu2 <= N and nu1 >= 1 and nu2 >= 1


def cmd_measure(args):
    import types

    import numpy as np
    import torch

    import bench
    out = []
    # the timing grid may leave the certified region (timings do not depend on the noise budget); `select` does not
    grid = [("cfg1", a, b) for (a, b) in [(6, 7), (7, 6), (8, 5), (6, 9), (7, 8), (8, 7), (9, 6), (10, 5), (8, 9), (9, 8), (10, 7)]]
    grid += [("cfg5", a, b) for (a, b) in [(6, 9), (7, 8), (8, 7), (9, 6), (9, 8)]]
    grid += [("wiki", a, b) for (a, b) in [(5, 10), (6, 9), (7, 8), (8, 7), (9, 6)]]
    if args.shapes:
        grid = [(t.split(":")[0], int(t.split(":")[1].split(",")[0]), int(t.split(":")[1].split(",")[1])) for t in args.shapes.split(";")]
    for name, g in GADGETS.items():
        bench.WORKLOADS.setdefault(name, dict(kind="spiral", prm={k: v for k, v in g.items() if not k.startswith("max_")}))
    torch.cuda.set_device(0)
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    fake = types.SimpleNamespace(exchange="p2p")
    for cfg, nu1, nu2 in grid:
        try:
            drv = bench.SpiralDriver(fake, cfg, nu1, nu2, 0, 1, 0, torch, None, np)
        except Exception as e:  # noqa: BLE001 - a shape the server refuses is reported, the sweep goes on
            print(json.dumps({"cfg": cfg, "nu1": nu1, "nu2": nu2, "skipped": str(e)}), file=sys.stderr, flush=True)
            continue
        drv.upload(stream)

        def step(marks=None):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            evs[0].record(); drv.stage_convert(stream)
            evs[1].record(); drv.stage_scan(stream)
            evs[2].record(); drv.stage_rest(stream)
            evs[3].record()
            if marks is not None:
                marks.append(evs)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        marks = []
        for _ in range(args.steps):
            step(marks)
        torch.cuda.synchronize()
        med = lambda v: sorted(v)[len(v) // 2]  # noqa: E731
        row = {"cfg": cfg, "nu1": nu1, "nu2": nu2, "db_bytes": drv.db_bytes,
               "exp_us": 1e3 * med([m[0].elapsed_time(m[1]) for m in marks]),
               "scan_us": 1e3 * med([m[1].elapsed_time(m[2]) for m in marks]),
               "fold_us": 1e3 * med([m[2].elapsed_time(m[3]) for m in marks]),
               "total_us": 1e3 * med([m[0].elapsed_time(m[3]) for m in marks])}
        print(json.dumps(row), file=sys.stderr, flush=True)
        out.append(row)
        drv.close()
        del drv
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump({"steps": args.steps, "gpu": torch.cuda.get_device_name(0), "rows": out}, f, indent=1)


def cmd_fit(args):
    import numpy as np
    rows = json.load(open(args.sweep))["rows"]
    model = {"source": os.path.basename(args.sweep), "form": {
        "exp_us": "c0 + c1*rounds + c2*2^nu1 + c3*t_GSW*nu2 + c4*[t_GSW*nu2 > 2^nu1]*2^rounds   (rounds = ceil(log2(2^nu1 + t_GSW*nu2)))",
        "scan_us": "c0 + db_bytes * (c1 if nu2 >= 6 else c2 if nu2 == 5 else c3)   (one bandwidth per scan tiling)",
        "fold_us": "c0 + c1*nu2 + c2*t_GSW*2^nu2"}, "coef": {}, "residual_pct": {}}
    for stage in ("exp", "scan", "fold"):
        X = np.array([shape_features(r["nu1"], r["nu2"], GADGETS[r["cfg"]])[stage] for r in rows])
        y = np.array([r[stage + "_us"] for r in rows])
        coef, *_ = np.linalg.lstsq(X, y, rcond=None)
        pred = X @ coef
        model["coef"][stage] = [float(c) for c in coef]
        model["residual_pct"][stage] = float(np.max(np.abs(pred - y) / y) * 100)
    model["scan_gbs_asymptotic"] = [1e-3 / c if c > 0 else None for c in model["coef"]["scan"][1:]]
    with open(args.out, "w") as f:
        json.dump(model, f, indent=1)
    print(json.dumps(model, indent=1))


def predict(model, nu1, nu2, g):
    f = shape_features(nu1, nu2, g)
    return {k: sum(c * x for c, x in zip(model["coef"][k], f[k])) for k in ("exp", "scan", "fold")}


def reference_cost_us(nu1, nu2, g):
    """select_params.py:179-198 (CPU-fitted), expansion LUT term omitted (it does not depend on nu_2)."""
    fold = 1000 * (33 + 29.6 * g["t_gsw"]) * (2 ** nu2 / 2 ** 6)
    first = 619.13591337 * 2 ** nu2 + 9.25842148 * 2 ** (nu1 + nu2)
    comp = 185451 * (2 ** nu1 / 2 ** 9) * (g["t_conv"] / 4)
    conv = 93709 * (nu2 * g["t_gsw"] / 40) * (g["t_conv"] / 4)
    return first + fold + comp + conv


def cmd_select(args):
    model = json.load(open(args.model))
    rows, skipped = [], []
    for name, g in GADGETS.items():
        for nu1 in range(1, 12):
            nu2 = args.log_items - nu1
            if nu2 < 1 or not feasible(nu1, nu2, g) or nu1 > g["max_nu1"] or nu2 > g["max_nu2"]:
                continue
            if nu2 < 5 and model["coef"]["scan"][3] == 0.0:
                skipped.append((name, nu1, nu2))            # no timing of this scan tiling in the sweep: not ranked
                continue
            p = predict(model, nu1, nu2, g)
            rows.append((sum(p.values()), name, nu1, nu2, p, reference_cost_us(nu1, nu2, g)))
    rows.sort()
    print(f"database of 2^{args.log_items} plaintext matrices ({8192 << args.log_items} bytes of records)")
    print("gadgets  nu1 nu2   B200 model us (exp / scan / fold)      reference CPU model us")
    for tot, name, nu1, nu2, p, ref in rows:
        print(f"{name:8s} {nu1:3d} {nu2:3d}   {tot:8.1f} ({p['exp']:6.1f} / {p['scan']:6.1f} / {p['fold']:6.1f})      {ref:12.0f}")
    if skipped:
        print("not ranked (fewer than 64 columns per z-slice, no timing in the sweep): " + ", ".join(f"{n} ({a},{b})" for n, a, b in skipped))
    if rows:
        best_ref = min(rows, key=lambda r: r[5])
        print(f"B200 optimum: {rows[0][1]} nu=({rows[0][2]},{rows[0][3]});  CPU-model optimum: {best_ref[1]} nu=({best_ref[2]},{best_ref[3]})")


# ---- emit: the B200 choice in the form select_params.py's consumers read ------------------------------------------------------
# all_parameter_choices.txt is a "/* Table */" line followed by a JSON object: "(log2 records, record bytes)" -> scheme -> {"params":
# {nu_1, nu_2, p, q_prime_bits, query_size, s_e, t_GSW, t_conv, t_exp, t_exp_right}} (all_parameter_choices.txt:1-80; run_all.py
# and select_params.py:300-340 look entries up by that key).  For every "spiral" entry whose gadget lengths are a set this model
# was fitted on, the record count is kept (nu_1 + nu_2 unchanged) and the split is re-chosen by the B200 model among the shapes the
# reference certified for those gadgets; everything else in the entry is copied.
def best_split(model, total_nu, g):
    cands = []
    for nu1 in range(1, 12):
        nu2 = total_nu - nu1
        if nu2 < 1 or not feasible(nu1, nu2, g) or nu1 > g["max_nu1"] or nu2 > g["max_nu2"]:
            continue
        if nu2 < 5 and model["coef"]["scan"][3] == 0.0:
            continue                                         # scan tiling not covered by the sweep
        cands.append((sum(predict(model, nu1, nu2, g).values()), nu1, nu2))
    return min(cands) if cands else None


def emit_choices(model, table):
    """table: the reference's dict; returns a dict of the same shape with the entries this model can re-choose."""
    out = {}
    for key, schemes in table.items():
        ent = schemes.get("spiral", {}).get("params")
        if not ent:
            continue
        match = [g for g in GADGETS.values() if (g["t_gsw"], g["t_conv"], g["t_exp"], g["t_exp_right"], g["qp_bits"], g["p_db"]) ==
                 (ent["t_GSW"], ent["t_conv"], ent["t_exp"], ent["t_exp_right"], ent["q_prime_bits"], ent["p"])]
        if not match:
            continue
        best = best_split(model, ent["nu_1"] + ent["nu_2"], match[0])
        if best is None:
            continue
        new = dict(ent, nu_1=best[1], nu_2=best[2])
        ref_us = sum(predict(model, ent["nu_1"], ent["nu_2"], match[0]).values())
        out[key] = {"spiral": {"params": new, "b200_model_us": round(best[0], 1), "reference_choice": {"nu_1": ent["nu_1"], "nu_2": ent["nu_2"],
                                                                                                      "b200_model_us": round(ref_us, 1)}}}
    return out


def parse_choices(text):
    """all_parameter_choices.txt -> [(section name, dict), ...]: sections are a C comment line followed by one or more JSON objects."""
    out, pos, name, dec = [], 0, "", json.JSONDecoder()
    while True:
        while pos < len(text) and text[pos] in " \t\r\n":
            pos += 1
        if pos >= len(text):
            return out
        if text.startswith("/*", pos):
            end = text.index("*/", pos)
            name, pos = text[pos + 2:end].strip(), end + 2
            continue
        obj, pos = dec.raw_decode(text, pos)
        out.append((name, obj))


def cmd_emit(args):
    model = json.load(open(args.model))
    with open(args.out, "w") as f:
        for name, table in parse_choices(open(args.table).read()):
            out = emit_choices(model, table)
            if not out:
                continue
            f.write(f"/* {name} */\n" + json.dumps(out, indent=4, sort_keys=True) + "\n")
            for k, v in sorted(out.items()):
                p, r = v["spiral"]["params"], v["spiral"]["reference_choice"]
                print(f"[{name}] {k}: B200 ({p['nu_1']},{p['nu_2']}) {v['spiral']['b200_model_us']} us;  reference ({r['nu_1']},{r['nu_2']}) "
                      f"{r['b200_model_us']} us on the B200 model")


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    m = sub.add_parser("measure"); m.add_argument("--out", default="gpurun_out/shape_sweep.json"); m.add_argument("--steps", type=int, default=20)
    m.add_argument("--shapes", default="", help='subset, e.g. "cfg1:9,6;cfg1:10,5"')
    f = sub.add_parser("fit"); f.add_argument("sweep"); f.add_argument("--out", default="profiles/cost_model_b200.json")
    s = sub.add_parser("select"); s.add_argument("model"); s.add_argument("--log-items", type=int, default=15)
    e = sub.add_parser("emit"); e.add_argument("model"); e.add_argument("table", help="the reference's all_parameter_choices.txt")
    e.add_argument("--out", default="profiles/b200_parameter_choices.txt")
    args = ap.parse_args()
    {"measure": cmd_measure, "fit": cmd_fit, "select": cmd_select, "emit": cmd_emit}[args.cmd](args)


if __name__ == "__main__":
    main()
