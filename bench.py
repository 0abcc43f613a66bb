#!/usr/bin/env python
"""bench.py - Spiral server-side query answering on B200: ms/query and database GB/s scanned.

One "step" = one query answered against the resident database (expansion + conversion +
first-dimension scan + folding + packing / modulus switch), BASELINE.json's metric.

  python bench.py --gpus 1 --steps K --warmup W            our CUDA path   (config.workload = cfg1, the headline)
  python bench.py --impl reference ...                      the UNMODIFIED reference (oracle/_ref) on host cores
  torchrun --nproc-per-node N bench.py --gpus N ...         second dimension sharded over N GPUs
  python bench.py --workload cfg3|cfg4|cfg5 [--scaling weak|strong]   the other BASELINE.json configurations
        cfg1  Spiral 2^20 x 256 B (./spiral 8 7)                    weak scaling by default: every GPU keeps a 2 GiB shard
        cfg5  Spiral 2^22 x 256 B (./spiral 9 8, 8 GiB)             strong scaling: the database is split over the GPUs
        cfg3  SpiralPack 2^18 x 30 KB (./spiral 10 8 --high-rate, 64 GiB)            strong
        cfg4  SpiralStreamPack 2^14 x 100 KB (./spiral 11 3 --high-rate --direct-upload, 6.25 GiB)   strong
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CFG1 = dict(t_gsw=8, t_conv=4, t_exp=8, t_exp_right=56, qp_bits=20, out_n=2, p_db=256)   # SURVEY 8d, `./spiral 8 7`
N_POLY = 2048
# BASELINE.json configs.  cfg1/cfg2 share one shape (the explicit database is the harder, honest one);
# cfg1 is the default and the headline; the others are selected with --workload.
WORKLOADS = {
    "cfg1": dict(kind="spiral", nu1=8, nu2=7, scaling="weak", prm=CFG1, flags=[],
                 macros="TEXP=8 TEXPRIGHT=56 TCONV=4 TGSW=8 QPBITS=20 PVALUE=256"),
    # cfg2 = cfg1's shape with the reference's --random-data ("implicit") database: one constant record, dummyWorkingSet =
    # min(2^25 / total_n, 2048) z-slices held in memory, slice z mod dummyWorkingSet scanned (src/spiral.cpp:647,1032-1081,1274-1282)
    "cfg2": dict(kind="spiral", nu1=8, nu2=7, scaling="weak", prm=CFG1, flags=["--random-data"], implicit=True,
                 macros="TEXP=8 TEXPRIGHT=56 TCONV=4 TGSW=8 QPBITS=20 PVALUE=256"),
    "cfg5": dict(kind="spiral", nu1=9, nu2=8, scaling="strong", flags=[],
                 prm=dict(t_gsw=9, t_conv=4, t_exp=8, t_exp_right=56, qp_bits=21, out_n=2, p_db=256),
                 macros="TEXP=8 TEXPRIGHT=56 TCONV=4 TGSW=9 QPBITS=21 PVALUE=256", sample=(9, 5)),
    "cfg3": dict(kind="pack", nu1=10, nu2=8, scaling="strong", direct=False, flags=["--high-rate"],
                 prm=dict(t_gsw=8, t_conv=4, t_exp=16, t_exp_right=56, qp_bits=20, out_n=4, p_db=256),
                 macros="TEXP=16 TEXPRIGHT=56 TCONV=4 TGSW=8 QPBITS=20 PVALUE=256 OUTN=4", sample=(8, 5)),
    "cfg4": dict(kind="pack", nu1=11, nu2=3, scaling="strong", direct=True, flags=["--high-rate", "--direct-upload"],
                 prm=dict(t_gsw=3, t_conv=56, t_exp=56, t_exp_right=56, qp_bits=27, out_n=5, p_db=65536),
                 macros="TEXP=56 TEXPRIGHT=56 TCONV=56 TGSW=3 QPBITS=27 PVALUE=65536 OUTN=5", sample=(9, 3)),
}


def host_isa():
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return "avx2"
    return "avx512" if " avx512f" in flags and " avx512dq" in flags and " avx512bw" in flags and " avx512vl" in flags else "avx2"


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def ceil_log2(x):
    g = 0
    while (1 << g) < x:
        g += 1
    return g


def record_bytes(wl):
    """Plaintext bytes of one database record (PIR item) of the workload."""
    p = wl["prm"]
    bits = (p["p_db"] - 1).bit_length()
    polys = 4 if wl["kind"] == "spiral" else p["out_n"] ** 2
    return polys * N_POLY * bits // 8


def workload_name(cfg, nu1, nu2):
    wl = WORKLOADS[cfg]
    rb = record_bytes(wl)
    if wl["kind"] == "spiral":
        # the reference's own accounting (print_summary, src/spiral.cpp:209-216): n = 2^(nu1+nu2) matrix plaintexts of
        # n0*n2*2048*log2(p) bits; BASELINE.json quotes the same database as 2^(nu1+nu2+5) records x 256 B
        recs, size = f"2^{nu1 + nu2 + 5}", "256 B"
        assert rb == 32 * 256
    else:
        recs, size = f"2^{nu1 + nu2}", {"cfg3": "30 KB (32 KiB)", "cfg4": "100 KB (100 KiB)"}.get(cfg, f"{rb // 1024} KiB")
    variant = {"cfg1": "Spiral", "cfg2": "Spiral", "cfg5": "Spiral", "cfg3": "SpiralPack", "cfg4": "SpiralStreamPack"}[cfg]
    flags = (" a " + " ".join(wl["flags"])) if wl["flags"] else ""
    kind = "implicit DB (one constant record, %d of 2048 slices resident)" % implicit_slices(nu1, nu2) if wl.get("implicit") else "explicit DB"
    return f"{variant} {recs} records x {size} (./spiral {nu1} {nu2}{flags}; {wl['macros']}), {kind}"


def implicit_slices(nu1, nu2):
    """dummyWorkingSet of the reference's --random-data mode (src/spiral.cpp:1277)."""
    return min((1 << 25) >> (nu1 + nu2), N_POLY) or 1


def db_bytes_total(wl, nu1, nu2):
    """Algorithmic database bytes: 8 B per NTT coefficient = the reference's own B / db_buf footprint."""
    polys = 4 if wl["kind"] == "spiral" else wl["prm"]["out_n"] ** 2
    return 8 * N_POLY * polys << (nu1 + nu2)


def workload_shape(args, world):
    """(nu1, nu2, scaling) of the whole job at `world` GPUs."""
    wl = WORKLOADS[args.workload]
    nu1 = args.nu1 if args.nu1 is not None else wl["nu1"]
    nu2 = args.nu2 if args.nu2 is not None else wl["nu2"]
    scaling = args.scaling or wl["scaling"]
    if scaling == "weak":
        nu2 += world.bit_length() - 1                        # fixed shard per GPU: the second dimension grows
    return nu1, nu2, scaling


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference through its own harness (oracle/ref_bench.cpp)
# ------------------------------------------------------------------------------------------------
STAGE_RE = {
    "expansion": r"Main expansion\s+\(CPU.us\):\s+(\d+)",
    "conversion": r"Conversion \(CPU.us\):\s+(\d+)",
    "first_dim": r"First dimension multiply \(CPU.us\):\s+(\d+)",
    "folding": r"Folding \(CPU.us\):\s+(\d+)",
    "packing": r"Packing \(CPU.us\):\s+(\d+)",
}
STAGES = ("expansion", "conversion", "first_dim", "folding", "packing")


def run_reference(cfg, nu1, nu2, queries, timeout=3000):
    """Per-query stage times (ms) parsed from the reference's own summary (src/spiral.cpp:239-263,
    src/testing.cpp:626-683).  Spiral: `Main expansion` and `Conversion` accumulate across queries
    (+=, src/spiral.cpp:2180,2256), so they are differenced; the Pack path exits after one query
    (src/spiral.cpp:1337-1340), so it is one process per query."""
    wl = WORKLOADS[cfg]
    isa = host_isa()
    exe = os.path.join(ROOT, "oracle", "_ref", f"ref_bench_{cfg}_{isa}")
    if not os.path.exists(exe):
        return None, f"{exe} not built"
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "oracle", "_ref") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    pack = wl["kind"] == "pack"
    chunks = []
    try:
        for _ in range(queries if pack else 1):
            # ref_bench: <nu1> <nu2> <idx> <queries> [reference flags] (it inserts the dummy file-name argument itself)
            out = subprocess.run([exe, str(nu1), str(nu2), "1234", "1" if pack else str(queries)] + wl["flags"],
                                 capture_output=True, text=True, timeout=timeout, env=env)
            if out.returncode != 0:
                return None, f"reference exited {out.returncode}: {out.stderr[-200:]}"
            chunks += out.stdout.split("=== ref_bench query")[1:]
    except Exception as e:  # noqa: BLE001
        return None, f"reference run failed: {e}"
    res, prev = [], dict(expansion=0, conversion=0)
    for ch in chunks:
        st = {}
        for k, rx in STAGE_RE.items():
            m = re.search(rx, ch)
            if not m and not (k == "packing" and not pack):
                return None, f"could not parse '{k}' from the reference output"
            st[k] = int(m.group(1)) if m else 0
        correct = re.search(r"Is correct\?\s*:\s*(\d)", ch)
        if pack:
            q = {k: st[k] / 1e3 for k in STAGES}
        else:
            q = dict(expansion=(st["expansion"] - prev["expansion"]) / 1e3, conversion=(st["conversion"] - prev["conversion"]) / 1e3,
                     first_dim=st["first_dim"] / 1e3, folding=st["folding"] / 1e3, packing=0.0)
            prev = dict(expansion=st["expansion"], conversion=st["conversion"])
        q["correct"] = bool(correct and correct.group(1) == "1")
        q["total"] = sum(q[k] for k in STAGES)
        res.append(q)
    return res, isa


def expansion_ntts(prm, nu1, nu2):
    """Forward-NTT count of coefficientExpansion (src/testing.cpp:40-105) - the scaling proxy for a reduced Pack sample."""
    nbits = prm["t_gsw"] * nu2
    g, stop = ceil_log2(nbits + (1 << nu1)), ceil_log2(max(nbits, 1))
    total = 0
    for r in range(g):
        for i in range(2 << r):
            if stop > 0 and r > stop and i % 2 == 1:
                continue
            if stop > 0 and r == stop and i % 2 == 1 and i // 2 > nbits:
                continue
            total += prm["t_exp_right"] if i % 2 else prm["t_exp"]
    return total


def pack_sample_scaled(cfg, nu1, nu2):
    """Pack workloads are too large for the reference on a host (cfg3: 64 GiB packed + 128 GiB of MatPoly copies):
    one query at the reduced shape WORKLOADS[cfg]['sample'] with the same parameter macros, each stage scaled by its own
    work ratio (first_dim ~ records, folding ~ 2^nu2 - 1, expansion ~ forward NTTs, conversion ~ nu2, packing ~ 1)."""
    wl = WORKLOADS[cfg]
    s1, s2 = wl["sample"]
    s1, s2 = min(s1, nu1), min(s2, nu2)
    res, info = run_reference(cfg, s1, s2, 1)
    if res is None:
        return None, info, None
    q = res[0]
    prm = wl["prm"]
    scale = dict(first_dim=float(1 << (nu1 + nu2 - s1 - s2)), folding=((1 << nu2) - 1) / max((1 << s2) - 1, 1), packing=1.0,
                 conversion=nu2 / max(s2, 1),
                 expansion=(float(1 << (nu1 - s1)) if wl.get("direct") else expansion_ntts(prm, nu1, nu2) / expansion_ntts(prm, s1, s2)))
    scaled = {k: q[k] * scale[k] for k in STAGES}
    scaled["total"] = sum(scaled[k] for k in STAGES)
    scaled["correct"] = q["correct"]
    text = (f"ONE query of the unmodified reference at the reduced shape ./spiral {s1} {s2} a {' '.join(wl['flags'])} (same macros; measured "
            f"{q['total']:.0f} ms: " + ", ".join(f"{k} {q[k]:.0f}" for k in STAGES) + "), each stage scaled by its work ratio "
            + ", ".join(f"{k} x{scale[k]:.3g}" for k in STAGES))
    return scaled, info, text


def spiral_sample_scaled(cfg, nu1, nu2):
    """Spiral workloads beyond the headline (cfg5: 8 GiB, ~1 min of host-side database generation alone): the 2nd of 2 queries of
    the unmodified reference at the reduced second dimension WORKLOADS[cfg]['sample'], each stage scaled by its work ratio
    (first_dim ~ records, folding ~ 2^nu2 - 1, expansion ~ forward NTTs, conversion ~ the reference's own cost model,
    select_params.py:185-187: ScalToMat ~ 2^nu1, RegevToGSW ~ nu2 * t_GSW)."""
    wl = WORKLOADS[cfg]
    s1, s2 = wl.get("sample", (nu1, nu2))
    s1, s2 = min(s1, nu1), min(s2, nu2)
    res, info = run_reference(cfg, s1, s2, 2)
    if res is None:
        return None, info, None
    q = res[-1]
    prm = wl["prm"]
    conv = lambda a, b: 185451.0 * (2 ** a / 512) * (prm["t_conv"] / 4) + 93709.0 * (b * prm["t_gsw"] / 40) * (prm["t_conv"] / 4)  # noqa: E731
    scale = dict(first_dim=float(1 << (nu1 + nu2 - s1 - s2)), folding=((1 << nu2) - 1) / max((1 << s2) - 1, 1), packing=1.0,
                 conversion=conv(nu1, nu2) / conv(s1, s2), expansion=expansion_ntts_spiral(prm, nu1, nu2) / expansion_ntts_spiral(prm, s1, s2))
    scaled = {k: q[k] * scale[k] for k in STAGES}
    scaled["total"] = sum(scaled[k] for k in STAGES)
    scaled["correct"] = q["correct"]
    text = (f"2nd of 2 queries of the unmodified reference at the reduced shape ./spiral {s1} {s2} (same macros; measured {q['total']:.0f} ms: "
            + ", ".join(f"{k} {q[k]:.0f}" for k in STAGES if k != "packing") + "), each stage scaled by its work ratio "
            + ", ".join(f"{k} x{scale[k]:.3g}" for k in STAGES if k != "packing"))
    return scaled, info, text


def expansion_ntts_spiral(prm, nu1, nu2):
    """Forward-NTT count of expandImproved (src/spiral.cpp:1664-1743) with runConversionImproved's stopround rule (:2080-2085)."""
    nbits = prm["t_gsw"] * nu2
    g = ceil_log2(nbits + (1 << nu1))
    stop = ceil_log2(max(nbits, 1)) if nbits <= (1 << nu1) else 0
    total = 0
    for r in range(g):
        for i in range(2 << r):
            if stop > 0 and r > stop and i % 2 == 1:
                continue
            if stop > 0 and r == stop and i % 2 == 1 and i // 2 > nbits:
                continue
            total += prm["t_exp_right"] if i % 2 else prm["t_exp"]
    return total


def oracle_port_baseline(nu1, nu2):
    """Fallback CPU baseline when oracle/_ref is absent: the oracle port (scalar C) on one core."""
    from tests import oracle_lib as ol
    lib = ol.load()
    s = ol.SpiralSession(lib, "cfg1", nu1, nu2, seed=1)
    Bbuf = s.reference_db()
    q = s.query(3)
    t0 = time.perf_counter()
    s.oracle_answer(q, Bbuf)
    dt = (time.perf_counter() - t0) * 1e3
    s.close()
    return dt


def mem_available():
    try:
        return int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:  # noqa: BLE001
        return 1 << 35


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cfg, wl = args.workload, WORKLOADS[args.workload]
    # the reference is one single-threaded process: it answers the SAME workload our arm runs at this GPU count,
    # unless that database would not fit comfortably in host memory
    nu1, nu2, scaling = workload_shape(args, args.gpus)
    base = {"impl": "reference", "metric": "server_ms_per_query", "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": False, "scaling": scaling, "vs_baseline": None, "dtype": "u64",
            "data": "synthetic", "config": {"workload": workload_name(cfg, nu1, nu2)}}
    if wl["kind"] == "pack":
        q, info, text = pack_sample_scaled(cfg, nu1, nu2)
        if q is None:
            print(json.dumps(dict(base, unavailable=info)))
            return 0
        stages = {k: q[k] for k in STAGES}
        line = dict(base, steps=1, warmup=0, value=q["total"], ms_per_step=q["total"], stages_ms=stages,
                    db_gbs_scanned=db_bytes_total(wl, nu1, nu2) / (stages["first_dim"] * 1e-3) / 1e9,
                    cpu_baseline={"value": q["total"], "unit": "ms", "cores": 1, "kind": "reference",
                                  "sample": text + f"; single thread on {cpu_model()} ({info}); decoded correctly: {q['correct']}"},
                    e2e={"value": q["total"], "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line))
        return 0
    nu2_run = nu2
    if db_bytes_total(wl, nu1, nu2) * 2.5 > mem_available():
        nu2_run = wl["nu2"] if args.nu2 is None else args.nu2
        while db_bytes_total(wl, nu1, nu2_run) * 2.5 > mem_available() and nu2_run > 1:
            nu2_run -= 1
        base["config"]["note"] = (f"host memory too small for the {db_bytes_total(wl, nu1, nu2) >> 30} GiB database of the {args.gpus}-GPU workload: "
                                  f"reference ran ./spiral {nu1} {nu2_run}")
        base["config"]["workload"] = workload_name(cfg, nu1, nu2_run)
    queries = args.steps + args.warmup
    if db_bytes_total(wl, nu1, nu2_run) > (2 << 30):     # ~1.4 s per CPU query and ~10 s of database generation per 2 GiB
        queries = min(queries, 3)
        args.warmup = min(args.warmup, 1)
        args.steps = queries - args.warmup
        base["steps"], base["warmup"] = args.steps, args.warmup
    res, info = run_reference(cfg, nu1, nu2_run, queries)
    if res is None:
        try:
            ms = oracle_port_baseline(6, 4)
            sample = "oracle port (scalar C restatement), ONE query at ./spiral 6 4 (1/32 of the records), 1 core: " + info
        except Exception as e:  # noqa: BLE001
            print(json.dumps(dict(base, unavailable=f"{info}; oracle port failed: {e}")))
            return 0
        line = dict(base, value=ms, ms_per_step=ms, cpu_baseline={"value": ms, "unit": "ms", "cores": 1, "kind": "port", "sample": sample},
                    e2e={"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line))
        return 0
    timed = res[args.warmup:] if len(res) > args.warmup else res
    ms = sum(q["total"] for q in timed) / len(timed)
    stages = {k: sum(q[k] for q in timed) / len(timed) for k in ("expansion", "conversion", "first_dim", "folding")}
    sample = (f"unmodified reference (oracle/_ref, g++ -O3 -march=x86-64-{'v4' if info == 'avx512' else 'v3'}, {info} scan path), "
              f"{len(timed)} full queries at the same workload after {args.warmup} warm-up, single thread (the reference is single-threaded, "
              f"src/spiral.cpp:1231) on {cpu_model()}; all queries decoded correctly: {all(q['correct'] for q in res)}")
    line = dict(base, value=ms, ms_per_step=ms, stages_ms=stages,
                db_gbs_scanned=db_bytes_total(wl, nu1, nu2_run) / (stages["first_dim"] * 1e-3) / 1e9,
                cpu_baseline={"value": ms, "unit": "ms", "cores": 1, "kind": "reference", "sample": sample},
                e2e={"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# clocks sampling (recipe in /opt/skills/guides/B200_PROFILING.md)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: an NVML polling thread (5 ms period - the timed region of the
    default run is only tens of milliseconds), with `nvidia-smi -lms` as the fallback when pynvml is unavailable."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index
        self.nvml, self.samples, self.reasons, self.stop_flag, self.max_mhz = None, [], set(), False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = (("hw_slowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))
        while not self.stop_flag:
            try:
                self.samples.append(int(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                for name, bit in bits:
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.t.join(timeout=1)
            sm = sorted(self.samples)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "nvml, 5 ms period over the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 8 and r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if len(r) > 8 and r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) > 8:
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# our arm: one driver per scheme variant, one shared timing harness
# ------------------------------------------------------------------------------------------------
CLIENT_SEED = bytes((37 * i + 11) & 0xFF for i in range(32))     # every rank derives the same client, keys and queries


def rnd_ntt_factory(np, seed):
    rng = np.random.default_rng(seed)

    def rnd_ntt(npolys):
        a = rng.integers(0, 249561089, size=(npolys, 2, N_POLY), dtype=np.uint64)
        return np.ascontiguousarray(a.reshape(-1))
    return rnd_ntt


def planted_targets(nu1, nu2, world, count=None):
    """Record indices to verify: one owned by EVERY rank (second-dimension index ii = rank mod world), spread over the first
    dimension too, so each shard's scan, folds and its slot of the peer exchange carry a checked answer."""
    num_per, dim0 = 1 << nu2, 1 << nu1
    local = num_per // world
    n = count or max(world, 2)
    out = []
    for k in range(n):
        g = k % world
        ii = g + world * ((5 * k + 3) % local)
        j = (k * 2654435761 + 12345) % dim0
        out.append(j * num_per + ii)
    return out


def planted_record(np, idx, polys, p_db):
    return np.random.default_rng(0xB200 + idx).integers(0, p_db, size=(polys, N_POLY), dtype=np.uint64)


class SpiralDriver:
    """Spiral / SpiralStream (matrix-Regev) path: resident SpiralServer, one 64 KiB query in, one 96 KiB response out.
    Keys and queries are REAL: the GPU client (sb200_client_*) generates the public parameters and encrypts the queries, known
    records are planted in the synthetic database, and verify() decodes the answers of the timed calls."""
    kernel = "k_scan_spiral"

    def __init__(self, args, cfg, nu1, nu2, rank, world, local_rank, torch, dist, np):
        from spiral_b200 import SpiralParams
        from spiral_b200.client import SpiralClient
        from spiral_b200.server import SpiralServer
        self.torch, self.dist, self.rank, self.world, self.np = torch, dist, rank, world, np
        self.nu1, self.nu2 = nu1, nu2
        p = WORKLOADS[cfg]["prm"]
        prm = SpiralParams(nu1, nu2, p["t_gsw"], p["t_conv"], p["t_exp"], p["t_exp_right"], p["qp_bits"], p["out_n"], p["p_db"])
        self.srv = srv = SpiralServer(prm, device=local_rank, rank=rank, world=world)
        self.implicit = bool(WORKLOADS[cfg].get("implicit"))
        self.const_value = 41                             # the one record of an implicit database (the reference draws rand % (p_db / 4))
        if self.implicit:
            srv.load_db_implicit_constant(self.const_value, implicit_slices(nu1, nu2))
        else:
            srv.load_db_random(seed=1000 + rank)
        self.use_p2p = world > 1 and args.exchange == "p2p"
        if self.use_p2p:                                  # cudaIpc handles of every rank's exchange buffer, rank order
            handles = [None] * world
            dist.all_gather_object(handles, srv.xchg_export())
            srv.xchg_connect(handles)
        self.p = p
        ell_bits = p["t_gsw"] * nu2
        self.g = ceil_log2(ell_bits + (1 << nu1))
        stop = ceil_log2(ell_bits) if ell_bits <= (1 << nu1) else 0
        self.n_right = stop + 1 if stop else self.g
        self.rnd_ntt = rnd_ntt_factory(np, 7)             # synthetic keys for the extra clients of the serving-throughput section
        # the client: same seed on every rank -> identical public parameters and queries without any broadcast
        self.client = SpiralClient(prm, CLIENT_SEED, device=local_rank)
        srv.set_public_params(*self.client.public_params())
        # planted records: the owner of ii = idx mod 2^nu2 overwrites its local item
        self.targets = planted_targets(nu1, nu2, world)
        num_per, local = 1 << nu2, (1 << nu2) // world
        for idx in self.targets:
            j, ii = divmod(idx, num_per)
            if ii % world == rank and not self.implicit:
                srv.load_db_items(planted_record(np, idx, 4, p["p_db"]).astype(np.uint16)[None], item_begin=j * local + ii // world)
        self.q_host = torch.empty(2 * 2 * N_POLY, dtype=torch.int64).pin_memory()
        self.set_query(self.targets[0], 1)
        self.resp_host = torch.empty(6 * N_POLY, dtype=torch.int64).pin_memory()
        self.gathered = torch.empty(world * 6 * N_POLY, dtype=torch.int64, device="cuda")
        self.part = torch.empty(6 * N_POLY, dtype=torch.int64, device="cuda")
        self.resp_dev = torch.empty(6 * N_POLY, dtype=torch.int64, device="cuda")
        self.h2d_bytes, self.d2h_bytes = int(self.q_host.numel() * 8), int(self.resp_host.numel() * 8)
        self.db_bytes = srv.db_bytes
        self.exchange = ("none (1 GPU)" if world == 1 else "peer-memory stores + flags over NVLink, fused into the stream (no NCCL call per query); the "
                         "expansion is sharded too: every rank expands / converts 1/N of the first-dimension and GSW ciphertexts and its "
                         "ScalToMat / RegevToGSW kernels store them into all ranks' buffers" if self.use_p2p else "NCCL all_gather of one 96 KiB ciphertext per GPU")
        if world == 1 or self.use_p2p:
            srv.prepare(self.resp_dev.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def set_query(self, idx, query_id):
        """Encrypt a query for record idx (GPU client, seeded wire form) and expand it to the 64 KiB in-memory ciphertext that
        sb200_server_upload_query / sb200_server_answer take (sb200_dev_query_from_wire + sb200_dev_ntt_to_ref)."""
        torch, lib = self.torch, self.srv.lib
        wire = self.client.query_wire(idx, query_id)
        wdev = torch.zeros(wire.size + 64, dtype=torch.uint8, device="cuda")
        wdev[:wire.size] = torch.from_numpy(wire).cuda()
        cv = torch.empty(2 * 2 * N_POLY, dtype=torch.int32, device="cuda")
        out = torch.empty(2 * 2 * N_POLY, dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for rc in (lib.sb200_dev_query_from_wire(cv.data_ptr(), wdev.data_ptr(), 1, st), lib.sb200_dev_ntt_to_ref(out.data_ptr(), cv.data_ptr(), 2, st)):
            if rc != 0:
                raise SystemExit("query expansion failed: " + lib.sb200_last_error().decode())
        torch.cuda.current_stream().synchronize()
        self.q_host.copy_(out.cpu())

    def expected(self, idx):
        if self.implicit:                                 # every record: the constant in coefficient 0 of its four polynomials
            rec = self.np.zeros((4, N_POLY), dtype=self.np.uint64)
            rec[:, 0] = self.const_value
            return rec
        return planted_record(self.np, idx, 4, self.p["p_db"])

    def decode_matches(self, idx):
        got = self.client.decode(self.resp_host.numpy().view(self.np.uint64))
        return bool(self.np.array_equal(got, self.expected(idx)))

    def upload(self, stream):
        self.srv.upload_query_ptr(self.q_host.data_ptr(), stream)

    def stage_convert(self, stream):
        self.srv.expand_and_convert(stream)

    def stage_scan(self, stream):
        self.srv.scan(stream)

    def stage_rest(self, stream):
        srv = self.srv
        srv.lift(stream)
        srv.fold_local(stream)
        if self.use_p2p:
            # exchange step fused into the stream: stores into rank 0's HBM over NVLink + flag, rank 0 waits and folds
            srv.exchange_and_tail(self.resp_dev.data_ptr(), stream)
        elif self.world > 1:
            srv.copy_partial(self.part.data_ptr(), stream)
            self.dist.all_gather_into_tensor(self.gathered, self.part)    # one 96 KiB ciphertext per GPU over NVLink (NCCL)
            if self.rank == 0:
                srv.fold_tail(self.gathered.data_ptr(), self.resp_dev.data_ptr(), stream)
        else:
            srv.fold_tail(srv.partial_ct_ptr(), self.resp_dev.data_ptr(), stream)

    def download(self):
        if self.rank == 0:
            self.resp_host.copy_(self.resp_dev, non_blocking=True)

    # one GPU: the whole resident query / the whole host-buffer query as ONE C-ABI call each (sb200_server_process,
    # sb200_server_answer) - the call a C++ host makes; the staged calls above remain for the NCCL-exchange path
    def process(self, stream, marks):
        self.srv.process(self.resp_dev.data_ptr(), stream, marks)

    def answer_host(self, stream):
        rc = self.srv.lib.sb200_server_answer(self.srv.h, self.q_host.data_ptr(), self.resp_host.data_ptr(), stream)
        if rc != 0:
            raise SystemExit("sb200_server_answer failed: " + self.srv.lib.sb200_last_error().decode())

    def check(self, stream):
        if self.use_p2p and self.srv.xchg_error(stream) != 0:
            raise SystemExit(f"rank {self.rank}: peer exchange timed out (error {self.srv.xchg_error(stream)})")

    def close(self):
        self.client.close()
        self.srv.close()


class PackDriver:
    """SpiralPack (packed query + expansion) / SpiralStreamPack (direct upload): resident PackServer, out_n^2 planes.
    Real keys and queries from the GPU Pack client, planted records, the exchange inside the C-ABI (sb200_pack_server_process).
    N > 1: SpiralPack shards the second dimension of every plane (tail folds on rank 0); SpiralStreamPack shards whole planes
    (its 8-column planes cannot be split 8 ways) and every rank uploads only 1/N of the direct-upload query."""
    kernel = "k_scan_pack"

    def __init__(self, args, cfg, nu1, nu2, rank, world, local_rank, torch, dist, np):
        from spiral_b200 import SpiralParams
        from spiral_b200.client import PackClient
        from spiral_b200.server import PackServer
        self.torch, self.dist, self.rank, self.world, self.np = torch, dist, rank, world, np
        self.nu1, self.nu2 = nu1, nu2
        wl = WORKLOADS[cfg]
        self.p = p = wl["prm"]
        self.direct = wl["direct"]
        self.shard = "planes" if (self.direct and world > 1) else "nu2"
        prm = SpiralParams(nu1, nu2, p["t_gsw"], p["t_conv"], p["t_exp"], p["t_exp_right"], p["qp_bits"], p["out_n"], p["p_db"])
        self.srv = srv = PackServer(prm, device=local_rank, rank=rank, world=world, shard=self.shard)
        srv.load_random(seed=1000 + (rank if self.shard == "nu2" else 0))
        self.use_p2p = world > 1
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, srv.xchg_export())
            srv.xchg_connect(handles)
        self.client = PackClient(prm, CLIENT_SEED, device=local_rank)
        self.pub = self.client.public_params(direct=self.direct)
        srv.set_public_params(*self.pub)
        self.view_params = lambda: self.pub
        # planted records: one polynomial per plane at each target index
        self.planes, self.dim0 = p["out_n"] ** 2, 1 << nu1
        self.targets = planted_targets(nu1, nu2, world if self.shard == "nu2" else 1)
        num_per = 1 << nu2
        for idx in self.targets:
            j, ii = divmod(idx, num_per)
            for pl in range(self.planes):
                if not srv.owns_plane(pl):
                    continue
                if self.shard == "nu2":
                    if ii % world == rank:
                        srv.set_plane_item(pl, j, ii // world, self.planted(idx)[pl].astype(np.uint16))
                else:
                    srv.set_plane_item(pl, j, ii, self.planted(idx)[pl].astype(np.uint16))
        PL = 2 * N_POLY
        if self.direct:
            self.vf_host = torch.empty(self.dim0 * 2 * PL, dtype=torch.int64).pin_memory()
            self.vg_host = torch.empty(max(nu2, 1) * 2 * 2 * p["t_gsw"] * PL, dtype=torch.int64).pin_memory()
            jc = self.dim0 // world
            self.slice_words = jc * 2 * PL
            self.h2d_bytes = int((self.slice_words if world > 1 else self.vf_host.numel()) + self.vg_host.numel()) * 8
        else:
            self.q_host = torch.empty(2 * PL, dtype=torch.int64).pin_memory()
            self.h2d_bytes = int(self.q_host.numel() * 8)
        self.set_query(self.targets[0], 1)
        self.resp_host = torch.empty(srv.response_words, dtype=torch.int64).pin_memory()
        self.resp_dev = torch.empty(srv.response_words, dtype=torch.int64, device="cuda")
        self.d2h_bytes = int(self.resp_host.numel() * 8)
        self.db_bytes = srv.db_bytes
        self.exchange = ("none (1 GPU)" if world == 1 else
                         (f"plane sharding: peer-memory stores of one folded 32 KiB ciphertext per plane to rank 0 (no fold after the exchange); "
                          f"the direct-upload query is uploaded 1/{world} per rank and all-gathered by the reorientation kernel over NVLink") if self.shard == "planes"
                         else f"peer-memory stores of {self.planes} surviving 32 KiB ciphertexts per GPU + flags over NVLink, fused into the stream"
                              + ("; the expansion is sharded too: every rank expands 1/N of the first-dimension ciphertexts and its reorientation kernel "
                                 "stores them into all ranks' query buffers" if self.direct is False and world > 1 else ""))

    def planted(self, idx):
        return self.np.stack([planted_record(self.np, idx * 64 + pl, 1, self.p["p_db"])[0] for pl in range(self.planes)])

    def set_query(self, idx, query_id):
        torch, np, lib = self.torch, self.np, self.srv.lib
        if self.direct:
            vf, vg = self.client.query_direct(idx, query_id % 256)
            self.vf_host.copy_(torch.from_numpy(vf.view(np.int64)))
            self.vg_host[:vg.size].copy_(torch.from_numpy(vg.view(np.int64)))
            return
        wire = self.client.query_wire(idx, query_id)
        wdev = torch.zeros(wire.size + 64, dtype=torch.uint8, device="cuda")
        wdev[:wire.size] = torch.from_numpy(wire).cuda()
        cv = torch.empty(2 * 2 * N_POLY, dtype=torch.int32, device="cuda")
        out = torch.empty(2 * 2 * N_POLY, dtype=torch.int64, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for rc in (lib.sb200_dev_query_from_wire(cv.data_ptr(), wdev.data_ptr(), 1, st), lib.sb200_dev_ntt_to_ref(out.data_ptr(), cv.data_ptr(), 2, st)):
            if rc != 0:
                raise SystemExit("query expansion failed: " + lib.sb200_last_error().decode())
        torch.cuda.current_stream().synchronize()
        self.q_host.copy_(out.cpu())

    def decode_matches(self, idx):
        got = self.client.decode(self.resp_host.numpy().view(self.np.uint64))
        return bool(self.np.array_equal(got, self.planted(idx)))

    def upload(self, stream):
        if not self.direct:
            self.srv.upload_query_ptr(self.q_host.data_ptr(), stream)
        elif self.world > 1:   # this rank's 1/N of the first-dimension ciphertexts; the reorientation kernel all-gathers over NVLink
            self.srv.upload_direct_split_ptr(self.vf_host.data_ptr() + self.rank * self.slice_words * 8, self.vg_host.data_ptr(), stream)
        else:
            self.srv.upload_direct_ptr(self.vf_host.data_ptr(), self.vg_host.data_ptr(), stream)

    def process(self, stream, marks):
        self.srv.process(self.resp_dev.data_ptr(), stream, marks)

    def answer_host(self, stream):
        self.upload(stream)
        self.srv.process(self.resp_dev.data_ptr(), stream, None)
        self.resp_host.copy_(self.resp_dev, non_blocking=True)
        self.torch.cuda.current_stream().synchronize()

    def download(self):
        if self.rank == 0:
            self.resp_host.copy_(self.resp_dev, non_blocking=True)

    def check(self, stream):
        if self.world > 1 and self.srv.xchg_error(stream) != 0:
            raise SystemExit(f"rank {self.rank}: peer exchange timed out (error {self.srv.xchg_error(stream)})")

    def close(self):
        self.client.close()
        self.srv.close()


class Ctx:
    pass


def run_workload(ctx, args, cfg, steps, headline):
    """Builds the resident server(s) of one BASELINE.json configuration, verifies planted records through the timed calls,
    times `steps` queries (device-resident and end to end) and returns the result dict on rank 0 (None elsewhere)."""
    torch, dist, np, lib = ctx.torch, ctx.dist, ctx.np, ctx.lib
    world, rank, local_rank, tstream, stream = ctx.world, ctx.rank, ctx.local_rank, ctx.tstream, ctx.stream
    wl = WORKLOADS[cfg]
    shape_args = args if headline else argparse.Namespace(workload=cfg, nu1=None, nu2=None, scaling=None)
    nu1, nu2, scaling = workload_shape(shape_args, world)
    lib.sb200_kernel_log_reset()                             # kernel names are recorded when a launch is issued or CAPTURED: from here on, this workload's
    drv = (SpiralDriver if wl["kind"] == "spiral" else PackDriver)(args, cfg, nu1, nu2, rank, world, local_rank, torch, dist, np)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    import ctypes
    fused = hasattr(drv, "process") and (world == 1 or drv.use_p2p)
    mark_pool = []

    def new_marks():
        """Four CUDA events for one query, created up front (torch creates the cudaEvent_t lazily at the first record)."""
        es = [ev() for _ in range(4)]
        for e in es:
            e.record()
        return es, (ctypes.c_void_p * 4)(*[e.cuda_event for e in es])

    def step(timed_events=None, e2e=False):
        """One query.  e2e=True includes the H2D of the query and the D2H of the response."""
        if fused:
            if e2e and world == 1:
                drv.answer_host(stream)                        # upload, all stages, download, stream synchronised
            elif e2e:
                drv.upload(stream); drv.process(stream, None); drv.download()
            elif timed_events is None or not mark_pool:
                drv.process(stream, None)
            else:
                es, handles = mark_pool.pop()
                drv.process(stream, handles)
                timed_events.append(es)
            return
        if e2e:
            drv.upload(stream)
        marks = []
        if timed_events is not None:
            marks.append(ev()); marks[-1].record()
        drv.stage_convert(stream)
        if timed_events is not None:
            marks.append(ev()); marks[-1].record()
        drv.stage_scan(stream)
        if timed_events is not None:
            marks.append(ev()); marks[-1].record()
        drv.stage_rest(stream)
        if timed_events is not None:
            marks.append(ev()); marks[-1].record()
            timed_events.append(marks)
        if e2e:
            drv.download()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness of the benchmarked configuration, through the exact calls that are timed below --------------------
    # every planted record (one owned by each rank) is queried with a real encryption, answered by the (sharded) servers -
    # at N > 1 through the cudaIpc / NVLink exchange - and decoded by the GPU client on rank 0
    verified = None
    if drv.targets:
        ok, modes = [], []
        for k, idx in enumerate(drv.targets):
            drv.set_query(idx, 100 + k)
            barrier()
            if k % 2 == 0:                                     # host-buffer call path (e2e)
                step(e2e=True); modes.append("host-buffer call")
            else:                                              # device-resident call path (value), response fetched afterwards
                drv.upload(stream); step(timed_events=[]); drv.download(); modes.append("resident call")
            torch.cuda.current_stream().synchronize()
            drv.check(stream)
            if rank == 0:
                ok.append(drv.decode_matches(idx))
        flag = torch.tensor([int(all(ok)) if rank == 0 else 1], device="cuda")
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        verified = {"queries": len(drv.targets), "decoded_equal_planted": int(sum(ok)) if rank == 0 else None, "record_indices": drv.targets,
                    "owner_ranks": [int((i % (1 << nu2)) % world) for i in drv.targets], "paths": sorted(set(modes)),
                    "how": "GPU client (sb200_client_*: keys, public parameters, encrypted queries) -> the timed server calls on every rank "
                           "(peer exchange included at N > 1) -> GPU decode on rank 0 == the record planted in the synthetic database"}
        if int(flag[0]) != 1:
            raise SystemExit(f"bench verification FAILED for {cfg}: decoded records {ok} for targets {drv.targets}")
        drv.set_query(drv.targets[0], 1)

    # resident upload once for the device-timed loop
    drv.upload(stream)
    for _ in range(max(args.warmup, 3)):
        step(timed_events=[])
    barrier()
    if os.environ.get("SB200_BENCH_TRACE"):                  # profiling aid: rank 0's in-graph timeline of one query of this (sharded) workload
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import trace_query
        if rank == 0:
            with open(os.environ["SB200_BENCH_TRACE"] + f"_{cfg}_{world}gpu.md", "w") as f:
                trace_query.trace_steps(lib, lambda: step(timed_events=None), 4, f"{workload_name(cfg, nu1, nu2)}, rank 0 of {world}", f)
        else:
            for _ in range(4):
                step(timed_events=None)
        barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # the timed region of `value`: EXACTLY `steps` queries, nothing but the server call inside (one graph launch per query)
    launches0 = lib.sb200_launch_count()
    t_begin, t_end = ev(), ev()
    barrier()
    t_begin.record()
    for _ in range(steps):
        step(timed_events=None)
    t_end.record()
    barrier()
    launches = lib.sb200_launch_count() - launches0
    total_ms = t_begin.elapsed_time(t_end)
    drv.check(stream)
    from spiral_b200.lib import kernel_log
    scan_kernels = sorted(k.strip("()") for k in kernel_log(lib) if "k_scan" in k)       # what the timed queries actually dispatched
    # the same `steps` queries again with four CUDA events per query on the launching stream (stage breakdown, scan duration for
    # the roofline).  The events sit between the stages, so the stages run as separate graphs here: their sum is a few percent
    # above `value`, which has no event inside.
    if fused:
        mark_pool.extend(new_marks() for _ in range(steps))
        torch.cuda.synchronize()
    events = []
    m_begin, m_end = ev(), ev()
    barrier()
    m_begin.record()
    for _ in range(steps):
        step(timed_events=events)
    m_end.record()
    barrier()
    marked_ms = m_begin.elapsed_time(m_end)
    drv.check(stream)

    # end to end through the host-buffer call path (H2D query + D2H response inside the timed region)
    e_begin, e_end = ev(), ev()
    barrier()
    e_begin.record()
    for _ in range(steps):
        step(timed_events=None, e2e=True)
        torch.cuda.current_stream().synchronize()             # the caller holds the response before the next query
    e_end.record()
    barrier()
    e2e_ms = e_begin.elapsed_time(e_end)
    clocks = sampler.stop() if rank == 0 else None

    # sustained: the same device-resident step back to back for >= 2 s, clocks and power sampled over the whole loop
    sustained = None
    if headline and args.sustained_s > 0:
        n_sus = max(steps, int(args.sustained_s * 1e3 / max(total_ms / steps, 1e-3)) + 1)
        if world > 1:
            t_n = torch.tensor([n_sus], device="cuda")
            dist.all_reduce(t_n, op=dist.ReduceOp.MAX)
            n_sus = int(t_n[0])
        s_sampler = ClockSampler(local_rank)
        s0, s1 = ev(), ev()
        barrier()
        if rank == 0:
            s_sampler.start()
        s0.record()
        for _ in range(n_sus):
            step(timed_events=None)
        s1.record()
        barrier()
        sus_ms = s0.elapsed_time(s1)
        s_clocks = s_sampler.stop() if rank == 0 else None
        t_s = torch.tensor([sus_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_s, op=dist.ReduceOp.MAX)
        sustained = {"value": float(t_s[0]) / n_sus, "unit": "ms", "queries": n_sus, "seconds": float(t_s[0]) / 1e3, "clocks": s_clocks}
        drv.check(stream)

    extras = headline_extras(ctx, args, cfg, drv, steps, ev) if headline else {}

    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = float(t[0]), float(t[1])

    scan_ms = sorted(m[1].elapsed_time(m[2]) for m in events)
    exp_ms = sum(m[0].elapsed_time(m[1]) for m in events) / len(events)
    rest_ms = sum(m[2].elapsed_time(m[3]) for m in events) / len(events)
    scan_avg = sum(scan_ms) / len(scan_ms)
    sc = torch.tensor([scan_avg], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(sc, op=dist.ReduceOp.MAX)
    scan_max = float(sc[0])

    line = None
    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        db_bytes_gpu = drv.db_bytes                            # algorithmic bytes per launch: 8 B x NTT coefficients on this GPU
        achieved = db_bytes_gpu / (scan_avg * 1e-3) / 1e9
        traffic, traffic_note = scan_traffic(cfg, world, scaling)
        ms_per_query = total_ms / steps
        l2_note = (f"database shard ({db_bytes_gpu / 2**30:.2f} GiB) is {db_bytes_gpu / 126e6:.0f}x the 126 MB L2 and is streamed once per query - no flush needed"
                   if db_bytes_gpu > 4 * 126e6 else f"database shard is only {db_bytes_gpu / 2**20:.0f} MiB: partly L2-resident between queries")
        line = {
            "metric": "server_ms_per_query", "value": ms_per_query, "unit": "ms", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_query, "higher_is_better": False, "scaling": scaling, "vs_baseline": None, "dtype": "u64 (28-bit CRT residues, 32x32->64 MAC)",
            "data": "synthetic",
            "config": {"workload": workload_name(cfg, nu1, nu2), "sharding": f"second dimension strided over {world} GPU(s), {db_bytes_gpu / 2**30:.2f} GiB shard per GPU",
                       "exchange": drv.exchange, "l2": l2_note},
            "stages_ms": {"expansion_conversion": exp_ms, "first_dim_scan": scan_avg, "lift_fold_modswitch": rest_ms},
            "stages_note": f"stage times: a second loop of {steps} queries with CUDA events between the stages ({marked_ms / steps:.4f} ms/query there; "
                           "`value` is timed with no event inside the query)",
            "db_gbs_scanned": world * db_bytes_gpu / (scan_max * 1e-3) / 1e9,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_source": traffic_note, "kernel": ", ".join(scan_kernels) or drv.kernel, "peak_source": peak_src,
                         "peak_note": "peak is the measured COPY bandwidth (read + write streams); a read-only stream reaches 7.2-7.4 TB/s on a B200 of this pool "
                                      "(scripts/micro/pipes.cu, profiles/r02_micro_pipes.txt), so frac can exceed 1 for a scan that only reads",
                         "scan_ms_min_med_max": [scan_ms[0], scan_ms[len(scan_ms) // 2], scan_ms[-1]]},
            "e2e": {"value": e2e_ms / steps, "unit": "ms", "h2d_bytes_per_step": drv.h2d_bytes, "d2h_bytes_per_step": drv.d2h_bytes},
            "gpu_launches": int(launches), "clocks": clocks, "verified": verified,
        }
        if sustained is not None:
            line["sustained"] = sustained
        line.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_for(cfg, wl, nu1, nu2, headline)
    drv.close()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    return line


def scan_traffic(cfg, world, scaling):
    """ncu dram__bytes (read + write) per launch of the scan kernel, from profiles/scan_traffic.json - valid only while the kernel
    sources it was captured with are the ones built now (the file records their SHA-256; scripts/update_scan_traffic.py)."""
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    if not os.path.exists(tp):
        return None, "no ncu capture committed"
    try:
        import hashlib
        tj = json.load(open(tp))
        for rel, want in tj.get("kernel_source_sha256", {}).items():
            have = hashlib.sha256(open(os.path.join(ROOT, rel), "rb").read()).hexdigest()
            if have != want:
                return None, f"stale: {rel} changed since the ncu capture of commit {tj.get('commit', '?')}"
        key = "dram_bytes_per_launch" if cfg == "cfg1" and scaling == "weak" else f"dram_bytes_per_launch_{cfg}_{world}gpu"
        if "kernel_source_sha256" not in tj:
            return None, "capture has no source hashes (made before round 2)"
        return tj.get(key), f"ncu --set full capture of commit {tj.get('commit', '?')} ({tj.get('captured', '?')})"
    except Exception as e:  # noqa: BLE001
        return None, f"unreadable: {e}"


def cpu_baseline_for(cfg, wl, nu1, nu2, headline):
    """The unmodified reference (oracle/_ref) on ONE host core: the same workload when it is small (cfg1), otherwise a bounded
    sample at a reduced shape with every stage scaled by its own work ratio (stated in `sample`)."""
    if wl["kind"] == "pack":
        q, info, text = pack_sample_scaled(cfg, nu1, nu2)
        if q is not None:
            return {"value": q["total"], "unit": "ms", "cores": 1, "kind": "reference", "stages_ms": {k: q[k] for k in STAGES},
                    "sample": text + f"; 1 thread on {cpu_model()} ({info}); decoded correctly: {q['correct']}"}
        return {"value": None, "unit": "ms", "cores": 0, "kind": "reference", "sample": f"unavailable: {info}"}
    if not headline:
        q, info, text = spiral_sample_scaled(cfg, nu1, nu2)
        if q is not None:
            return {"value": q["total"], "unit": "ms", "cores": 1, "kind": "reference", "stages_ms": {k: q[k] for k in STAGES if k != "packing"},
                    "sample": text + f"; 1 thread on {cpu_model()} ({info}); decoded correctly: {q['correct']}"}
        return {"value": None, "unit": "ms", "cores": 0, "kind": "reference", "sample": f"unavailable: {info}"}
    res, info = (None, "database too large for the host") if db_bytes_total(wl, nu1, nu2) * 2.5 > mem_available() else run_reference(cfg, nu1, nu2, 2)
    if res is not None:
        qd = res[-1]
        return {"value": qd["total"], "unit": "ms", "cores": 1, "kind": "reference",
                "stages_ms": {k: qd[k] for k in ("expansion", "conversion", "first_dim", "folding")},
                "sample": f"unmodified reference (oracle/_ref, {info}), 2nd of 2 full queries at the same workload, 1 thread on {cpu_model()}, decoded correctly: {qd['correct']}"}
    try:
        ms = oracle_port_baseline(6, 4)
        return {"value": ms, "unit": "ms", "cores": 1, "kind": "port", "sample": f"oracle port, one query at ./spiral 6 4 (1/32 of the records), 1 core ({info})"}
    except Exception as e:  # noqa: BLE001
        return {"value": None, "unit": "ms", "cores": 0, "kind": "port", "sample": f"unavailable: {e}"}


def headline_extras(ctx, args, cfg, drv, steps, ev):
    """Sections only the headline workload carries: the wire-format call path and the serving-throughput experiments."""
    torch, np, lib = ctx.torch, ctx.np, ctx.lib
    world, tstream, stream = ctx.world, ctx.tstream, ctx.stream
    wl = WORKLOADS[cfg]
    from spiral_b200.server import SpiralServer
    out = {}
    # the same exchange in wire form: seed-compressed query in (14 376 B), QPBITS-packed response out;
    # seed expansion + unpacking are the first node of the expansion graph, the response packer one extra launch
    if world == 1 and wl["kind"] == "spiral":
        srv = drv.srv
        wbytes = int(lib.sb200_wire_query_bytes(1))
        wire = drv.client.query_wire(drv.targets[1], 7)
        wire_host = torch.from_numpy(wire).pin_memory()
        pbytes = int(lib.sb200_server_packed_response_bytes(srv.h))
        packed_host = torch.empty(pbytes // 8, dtype=torch.int64).pin_memory()

        def wire_query():
            rc = lib.sb200_server_answer_wire(srv.h, wire_host.data_ptr(), wbytes, packed_host.data_ptr(), stream)
            if rc != 0:
                raise SystemExit("answer_wire failed: " + lib.sb200_last_error().decode())
        for _ in range(3):
            wire_query()
        w_begin, w_end = ev(), ev()
        torch.cuda.synchronize()
        w_begin.record()
        for _ in range(steps):
            wire_query()                                       # synchronises the stream: the caller holds the packed response
        w_end.record()
        torch.cuda.synchronize()
        got = drv.client.decode(srv.unpack_response(packed_host.numpy().view(np.uint64)))
        wire_ok = bool(np.array_equal(got, drv.expected(drv.targets[1])))
        if not wire_ok:
            raise SystemExit("bench verification FAILED: wire-format answer does not decode to the planted record")
        out["e2e_wire"] = {"value": w_begin.elapsed_time(w_end) / steps, "unit": "ms", "h2d_bytes_per_step": wbytes, "d2h_bytes_per_step": pbytes,
                           "decoded_equal_planted": wire_ok,
                           "note": "sb200_server_answer_wire: seeded wire query (ChaCha20 row 0) in, packed response out"}
        drv.upload(stream)                                     # back to the plain query for the sections below

    # serving throughput on ONE GPU: several clients in flight (views over the same resident database, one stream
    # each).  The latency-bound expansion / fold chains of one query overlap the HBM-bound scan of another.
    pipelined = None
    if world == 1 and args.clients > 1 and wl["kind"] == "spiral":
        srv, p, rnd_ntt = drv.srv, drv.p, drv.rnd_ntt
        q_host, resp_host, resp_dev = drv.q_host, drv.resp_host, drv.resp_dev
        clients = [srv] + [srv.view() for _ in range(args.clients - 1)]
        streams = [tstream] + [torch.cuda.Stream() for _ in range(args.clients - 1)]
        resp_hosts = [resp_host] + [torch.empty(6 * N_POLY, dtype=torch.int64).pin_memory() for _ in range(args.clients - 1)]
        resp_devs = [resp_dev] + [torch.empty(6 * N_POLY, dtype=torch.int64, device="cuda") for _ in range(args.clients - 1)]
        for c in clients[1:]:
            c.set_public_params(rnd_ntt(drv.g * 2 * p["t_exp"]), rnd_ntt(drv.n_right * 2 * p["t_exp_right"]),
                                rnd_ntt(3 * 2 * p["t_conv"]), rnd_ntt(3 * 2 * p["t_conv"]))

        def one(ci):
            c, st = clients[ci], streams[ci]
            with torch.cuda.stream(st):
                s_ = st.cuda_stream
                c.upload_query_ptr(q_host.data_ptr(), s_)
                c.expand_and_convert(s_); c.scan(s_); c.lift(s_); c.fold_local(s_)
                c.fold_tail(c.partial_ct_ptr(), resp_devs[ci].data_ptr(), s_)
                resp_hosts[ci].copy_(resp_devs[ci], non_blocking=True)
        for _ in range(3):
            for ci in range(len(clients)):
                one(ci)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            for ci in range(len(clients)):
                one(ci)
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        nq = steps * len(clients)
        pipelined = {"clients": len(clients), "queries": nq, "ms_per_query_amortised": wall_ms / nq, "queries_per_s": nq / (wall_ms * 1e-3),
                     "db_gbs_scanned": srv.db_bytes * nq / (wall_ms * 1e-3) / 1e9,
                     "note": "host wall clock around all streams; every query includes its H2D query upload and D2H response"}
        # tensor-core batched first dimension (tc_scan.cu): up to 16 clients' converted queries answered by ONE tcgen05 pass
        # over the limb-tile copy of the database; expansions and folds of the clients overlap on their own streams
        if args.tc_batch > 1 and lib.sb200_tc_supported(srv.dim0, srv.num_per) and not drv.implicit:
            nb = min(args.tc_batch, 16)
            while len(clients) < nb:
                c = srv.view()
                c.set_public_params(rnd_ntt(drv.g * 2 * p["t_exp"]), rnd_ntt(drv.n_right * 2 * p["t_exp_right"]),
                                    rnd_ntt(3 * 2 * p["t_conv"]), rnd_ntt(3 * 2 * p["t_conv"]))
                clients.append(c); streams.append(torch.cuda.Stream())
                resp_hosts.append(torch.empty(6 * N_POLY, dtype=torch.int64).pin_memory())
                resp_devs.append(torch.empty(6 * N_POLY, dtype=torch.int64, device="cuda"))
            tcc, tcs = clients[:nb], streams[:nb]
            srv.enable_tc(nb)
            evs = [torch.cuda.Event() for _ in tcc]
            ev_scan = torch.cuda.Event()

            def tc_round():
                for ci, (c, st) in enumerate(zip(tcc, tcs)):
                    with torch.cuda.stream(st):
                        c.upload_query_ptr(q_host.data_ptr(), st.cuda_stream)
                        c.expand_and_convert(st.cuda_stream)
                        evs[ci].record(st)
                with torch.cuda.stream(tcs[0]):
                    for ci in range(1, nb):
                        tcs[0].wait_event(evs[ci])
                    SpiralServer.scan_batched_tc(tcc, tcs[0].cuda_stream)
                    ev_scan.record(tcs[0])
                for ci, (c, st) in enumerate(zip(tcc, tcs)):
                    with torch.cuda.stream(st):
                        st.wait_event(ev_scan)
                        c.lift(st.cuda_stream); c.fold_local(st.cuda_stream)
                        c.fold_tail(c.partial_ct_ptr(), resp_devs[ci].data_ptr(), st.cuda_stream)
                        resp_hosts[ci].copy_(resp_devs[ci], non_blocking=True)
            for _ in range(3):
                tc_round()
            torch.cuda.synchronize()
            # client 0 holds the real keys and the real query: its answer out of the shared tensor-core pass must decode
            tc_ok = bool(np.array_equal(drv.client.decode(resp_hosts[0].numpy().view(np.uint64)), drv.expected(drv.targets[0])))
            if not tc_ok:
                raise SystemExit("bench verification FAILED: tensor-core batched answer does not decode to the planted record")
            rounds = max(3, steps // 2)
            t0 = time.perf_counter()
            for _ in range(rounds):
                tc_round()
            torch.cuda.synchronize()
            wall_ms = (time.perf_counter() - t0) * 1e3
            sb0, sb1 = ev(), ev()
            with torch.cuda.stream(tcs[0]):
                sb0.record(tcs[0])
                for _ in range(5):
                    SpiralServer.scan_batched_tc(tcc, tcs[0].cuda_stream)
                sb1.record(tcs[0])
            torch.cuda.synchronize()
            pass_ms = sb0.elapsed_time(sb1) / 5
            nq = rounds * nb
            pipelined["tensor_core_batch"] = {
                "queries_per_pass": nb, "ms_per_query_amortised": wall_ms / nq, "queries_per_s": nq / (wall_ms * 1e-3),
                "first_dim_pass_ms": pass_ms, "first_dim_ms_per_query": pass_ms / nb,
                "db_gbs_per_pass": srv.db_bytes / (pass_ms * 1e-3) / 1e9,
                "effective_db_gbs_x_queries": srv.db_bytes * nb / (pass_ms * 1e-3) / 1e9,
                "int8_tensor_tops": 2.0 * 16 * 3 * nb * (srv.db_bytes / 8) * 2 / (pass_ms * 1e-3) / 1e12,
                "decoded_equal_planted": tc_ok,
                "note": "pass = query tiles (k_query_to_tc x queries) + k_scan_tc (tcgen05 kind::i8, u8 limbs) + k_tc_untile; bit-exact (tests/test_gpu_tc.py)"}
        for c in clients[1:]:
            c.close()
        torch.cuda.set_stream(tstream)

    # SpiralPack: the scan is most of a query (cfg3: 11.5 of 15 ms), so one tensor-core pass over the planes for a batch of
    # clients multiplies the serving throughput; expansions / folds of the clients overlap on their own streams
    if world == 1 and args.tc_batch > 1 and wl["kind"] == "pack" and not wl["direct"] and drv.srv.dim0 % 128 == 0 and drv.srv.num_per % 128 == 0:
        from spiral_b200.server import PackServer
        srv = drv.srv
        nb = min(args.tc_batch, 16)      # cfg3: 64 GiB planes + 64 GiB limb tiles + 16 client contexts of ~1.2 GB fit the 180 GB
        srv.enable_tc(nb)
        tcc = [srv] + [srv.view() for _ in range(nb - 1)]
        for c in tcc[1:]:
            c.set_public_params(*drv.view_params())
        tcs = [tstream] + [torch.cuda.Stream() for _ in range(nb - 1)]
        resp_devs = [drv.resp_dev] + [torch.empty(srv.response_words, dtype=torch.int64, device="cuda") for _ in range(nb - 1)]
        resp_hosts = [drv.resp_host] + [torch.empty(srv.response_words, dtype=torch.int64).pin_memory() for _ in range(nb - 1)]
        evs = [torch.cuda.Event() for _ in tcc]
        ev_scan = torch.cuda.Event()

        def pack_tc_round():
            for ci, (c, st) in enumerate(zip(tcc, tcs)):
                with torch.cuda.stream(st):
                    c.upload_query_ptr(drv.q_host.data_ptr(), st.cuda_stream)
                    c.expand_and_convert(st.cuda_stream)
                    evs[ci].record(st)
            with torch.cuda.stream(tcs[0]):
                for ci in range(1, nb):
                    tcs[0].wait_event(evs[ci])
                PackServer.scan_batched_tc(tcc, tcs[0].cuda_stream)
                ev_scan.record(tcs[0])
            for ci, (c, st) in enumerate(zip(tcc, tcs)):
                with torch.cuda.stream(st):
                    st.wait_event(ev_scan)
                    c.fold_local(st.cuda_stream)
                    c.fold_tail(c.partial_cts_ptr(), resp_devs[ci].data_ptr(), st.cuda_stream)
                    resp_hosts[ci].copy_(resp_devs[ci], non_blocking=True)
        for _ in range(2):
            pack_tc_round()
        torch.cuda.synchronize()
        rounds = max(2, steps // 4)
        t0 = time.perf_counter()
        for _ in range(rounds):
            pack_tc_round()
        torch.cuda.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        sb0, sb1 = ev(), ev()
        with torch.cuda.stream(tcs[0]):
            sb0.record(tcs[0])
            for _ in range(3):
                PackServer.scan_batched_tc(tcc, tcs[0].cuda_stream)
            sb1.record(tcs[0])
        torch.cuda.synchronize()
        pass_ms = sb0.elapsed_time(sb1) / 3
        nq = rounds * nb
        pipelined = {"tensor_core_batch": {
            "queries_per_pass": nb, "ms_per_query_amortised": wall_ms / nq, "queries_per_s": nq / (wall_ms * 1e-3),
            "first_dim_pass_ms": pass_ms, "first_dim_ms_per_query": pass_ms / nb,
            "db_gbs_per_pass": srv.db_bytes / (pass_ms * 1e-3) / 1e9,
            "effective_db_gbs_x_queries": srv.db_bytes * nb / (pass_ms * 1e-3) / 1e9,
            "int8_tensor_tops": 2.0 * 16 * 2 * nb * (srv.db_bytes / 8) * 2 / (pass_ms * 1e-3) / 1e12,
            "note": "pass = query tiles + k_scan_tc over all planes (tcgen05 kind::i8, u8 limbs) + k_tc_untile; bit-exact (tests/test_gpu_tc.py); host wall clock, H2D/D2H per query included"}}
        for c in tcc[1:]:
            c.close()
        torch.cuda.set_stream(tstream)
    if pipelined is not None:
        out["pipelined"] = pipelined
    return out


def b200_main(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from spiral_b200.lib import load_library

    ctx = Ctx()
    ctx.torch, ctx.dist, ctx.np = torch, dist, np
    ctx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    ctx.rank = rank = int(os.environ.get("RANK", "0"))
    ctx.local_rank = local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner) write to fd 1: keep the real stdout for the single JSON line
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx.lib = load_library()
    # a dedicated (non-legacy) stream: kernels, graph replays, events, NCCL and copies all run on it
    ctx.tstream = torch.cuda.Stream()
    torch.cuda.set_stream(ctx.tstream)
    ctx.stream = ctx.tstream.cuda_stream

    line = run_workload(ctx, args, args.workload, args.steps, headline=True)
    # the other north-star databases ride on the same line, so the driver's BENCH / SCALE records carry them at every N:
    # cfg5 (Spiral 2^22 x 256 B, 8 GiB) and cfg3 (SpiralPack 2^18 x 30 KB, 64 GiB), both strong-scaled over the N GPUs
    extra = [w for w in args.workloads.split(",") if w] if args.workload == "cfg1" and args.nu1 is None and args.nu2 is None else []
    subs = {}
    for w in extra:
        if w not in WORKLOADS or w == args.workload:
            continue
        sub = run_workload(ctx, args, w, max(3, min(args.steps, 10 if WORKLOADS[w]["kind"] == "spiral" else 5)), headline=False)
        if sub is not None:
            subs[w] = {k: sub[k] for k in ("value", "unit", "steps", "scaling", "config", "stages_ms", "db_gbs_scanned", "roofline", "e2e",
                                           "gpu_launches", "clocks", "verified", "cpu_baseline") if k in sub}
    if rank == 0:
        if subs:
            line["workloads"] = subs
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg1", choices=sorted(WORKLOADS), help="BASELINE.json configuration (cfg1 = the headline)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"], help="default: weak for cfg1 (2 GiB shard per GPU), strong for the others")
    ap.add_argument("--nu1", type=int, default=None)
    ap.add_argument("--nu2", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clients", type=int, default=4, help="N = 1: concurrent clients for the serving-throughput figure (0/1 = skip)")
    ap.add_argument("--tc-batch", type=int, default=16, help="N = 1, Spiral: queries per tensor-core database pass in the serving-throughput section (0/1 = skip)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N > 1: how the surviving ciphertexts reach rank 0")
    ap.add_argument("--workloads", default="cfg5,cfg3", help="with the default headline (cfg1): further BASELINE.json configurations measured into "
                    "the line's `workloads` section ('' = none)")
    ap.add_argument("--sustained-s", type=float, default=2.0, help="seconds of back-to-back queries for the `sustained` figure (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.steps > 8:
            args.steps = 8                                    # ~3 s per CPU query; keep the whole run within minutes
        return reference_main(args)
    return b200_main(args)


if __name__ == "__main__":
    sys.exit(main())
