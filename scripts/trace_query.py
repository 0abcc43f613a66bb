#!/usr/bin/env python
"""Timeline of ONE query inside the replayed CUDA graphs (sb200_trace_enable / sb200_trace_read): for every kernel the time its
first CTA was scheduled and the time its dependencies were resolved (griddepcontrol.wait returned).  The difference between
consecutive dependency-resolved times is what each link of the launch chain really costs - ncu's serialised, cold-cache
durations cannot show that.  usage: python scripts/trace_query.py [cfg1|cfg5] [--nu1 N --nu2 N] > profiles/rNN_trace.md"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from spiral_b200 import SpiralParams  # noqa: E402
from spiral_b200.lib import load_library  # noqa: E402
from spiral_b200.server import SpiralServer  # noqa: E402


def kernel_lines():
    """source line of every pdl_prologue() call -> the kernel(s) it sits in (lines can collide across the unity build's files)."""
    import re
    out = {}
    csrc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "spiral_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        if not f.endswith((".cu", ".cuh")):
            continue
        cur = None
        for n, text in enumerate(open(os.path.join(csrc, f)), 1):
            m = re.search(r"\b(k_[a-z0-9_]+)\s*\(", text)
            if m and "__global__" in text or (m and cur is None and "launch_pdl" not in text and "void" in text):
                cur = m.group(1)
            elif "__global__" in text:
                cur = None                                  # name on the next line
            elif cur is None and m and "launch" not in text:
                cur = m.group(1)
            if ("pdl_prologue()" in text or "pdl_prologue_no_early_dependents()" in text or "pdl_wait()" in text) and "#define" not in text and cur:
                out.setdefault(n, []).append(cur)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", nargs="?", default="cfg1")
    ap.add_argument("--nu1", type=int)
    ap.add_argument("--nu2", type=int)
    ap.add_argument("--queries", type=int, default=3)
    a = ap.parse_args()
    wl = bench.WORKLOADS[a.workload]
    p = wl["prm"]
    nu1, nu2 = a.nu1 or wl["nu1"], a.nu2 or wl["nu2"]
    lib = load_library()
    prm = SpiralParams(nu1, nu2, p["t_gsw"], p["t_conv"], p["t_exp"], p["t_exp_right"], p["qp_bits"], p["out_n"], p["p_db"])
    srv = SpiralServer(prm)
    srv.load_db_random(1)
    rnd = bench.rnd_ntt_factory(np, 7)
    nbits = p["t_gsw"] * nu2
    g = bench.ceil_log2(nbits + (1 << nu1))
    stop = bench.ceil_log2(nbits) if nbits <= (1 << nu1) else 0
    srv.set_public_params(rnd(g * 2 * p["t_exp"]), rnd((stop + 1 if stop else g) * 2 * p["t_exp_right"]), rnd(6 * p["t_conv"]), rnd(6 * p["t_conv"]))
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    q = torch.from_numpy(rnd(2).view(np.int64)).pin_memory()
    srv.upload_query_ptr(q.data_ptr(), st.cuda_stream)
    for _ in range(5):
        srv.process(None, st.cuda_stream, None)
    torch.cuda.synchronize()
    trace_steps(lib, lambda: srv.process(None, st.cuda_stream, None), a.queries, bench.workload_name(a.workload, nu1, nu2), sys.stdout)
    srv.close()


def trace_steps(lib, step, queries, title, out, sync=None):
    """Runs `step` `queries` times with the in-graph trace on and prints the LAST query's timeline to `out`."""
    sync = sync or torch.cuda.synchronize
    sync()
    assert lib.sb200_trace_enable(4096) == 0
    for _ in range(queries):
        step()
    sync()
    buf = np.zeros(4096 * 3, dtype=np.uint64)
    n = lib.sb200_trace_read(buf.ctypes.data, 4096, 1)
    lib.sb200_trace_enable(0)
    rec = buf[:3 * n].reshape(n, 3)
    per = n // queries
    rec = rec[(queries - 1) * per:]                         # the last query
    rec = rec[np.argsort(rec[:, 1], kind="stable")]
    t0 = int(rec[0, 1])
    print(f"# {title}: one query, {per} kernels, {(int(rec[-1, 1]) - t0) / 1e3:.1f} us from the first to the last dependency-resolved time", file=out)
    names = kernel_lines()
    print("| # | kernel | grid | block | scheduled us | ready us | to next ready us | waited us |", file=out)
    print("|---:|---|---|---:|---:|---:|---:|---:|", file=out)
    for i, (ts, tr, shape) in enumerate(rec):
        shape = int(shape)
        gx, gy, bx, line = shape & 0xFFFFF, (shape >> 20) & 0xFFFF, (shape >> 36) & 0xFFF, shape >> 48
        if line >= 0xF000:
            print(f"| {i} | mark {line & 0xFFF} | | | | {(int(tr) - t0) / 1e3:.1f} | | |", file=out)
            continue
        nxt = (int(rec[i + 1, 1]) - int(tr)) / 1e3 if i + 1 < len(rec) else 0.0
        print(f"| {i} | {'/'.join(names.get(line, ['?']))} | ({gx},{gy}) | {bx} | {(int(ts) - t0) / 1e3:.1f} | {(int(tr) - t0) / 1e3:.1f} | {nxt:.1f} | {(int(tr) - int(ts)) / 1e3:.1f} |", file=out)


if __name__ == "__main__":
    main()
