#!/usr/bin/env python
"""Timing of the SpiralPack / SpiralStreamPack resident server (not the headline bench): whole-query
wall time through the C-ABI with host buffers, and the scan's share via ncu when run under it.
usage: python scripts/bench_pack.py cfg3|cfg4 nu1 nu2 [out_n] [steps]"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from spiral_b200 import SpiralParams  # noqa: E402
from spiral_b200.lib import check, load_library  # noqa: E402

CFG = {"cfg3": dict(t_gsw=8, t_conv=4, t_exp=16, t_exp_right=56, qp_bits=20, out_n=4, p_db=256),
       "cfg4": dict(t_gsw=3, t_conv=56, t_exp=56, t_exp_right=56, qp_bits=27, out_n=5, p_db=65536)}
N = 2048


def main():
    cfg, nu1, nu2 = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    c = dict(CFG[cfg])
    if len(sys.argv) > 4:
        c["out_n"] = int(sys.argv[4])
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
    sb = load_library()
    prm = SpiralParams(nu1, nu2, c["t_gsw"], c["t_conv"], c["t_exp"], c["t_exp_right"], c["qp_bits"], c["out_n"], c["p_db"])
    h = C.c_void_p()
    check(sb.sb200_pack_server_create(C.byref(h), C.byref(prm), 0), sb)
    t0 = time.perf_counter()
    check(sb.sb200_pack_server_load_random(h, 5), sb)
    load_s = time.perf_counter() - t0
    rng = np.random.default_rng(3)

    def rnd(npolys):
        return np.ascontiguousarray(rng.integers(0, 249561089, size=(npolys, 2, N), dtype=np.uint64).reshape(-1))
    n, ell, dim0 = c["out_n"], c["t_gsw"], 1 << nu1
    nbits = ell * nu2
    g = int(np.ceil(np.log2(nbits + dim0)))
    stop = int(np.ceil(np.log2(max(nbits, 1))))
    vW = rnd(n * (n + 1) * c["t_conv"])
    resp = np.zeros((n + 1) * n * N, dtype=np.uint64)
    direct = cfg == "cfg4"
    if direct:
        check(sb.sb200_pack_server_set_public_params(h, None, None, None, vW.ctypes.data_as(C.POINTER(C.c_uint64))), sb)
        vf, vg = rnd(dim0 * 2), rnd(max(nu2, 1) * 2 * 2 * ell)
        call = lambda: check(sb.sb200_pack_server_answer_direct(h, vf.ctypes.data, vg.ctypes.data, resp.ctypes.data, None, None), sb)  # noqa: E731
    else:
        P = C.POINTER(C.c_uint64)
        Wl, Wr, V = rnd(g * 2 * c["t_exp"]), rnd((stop + 1) * 2 * c["t_exp_right"]), rnd(2 * 2 * c["t_conv"])
        check(sb.sb200_pack_server_set_public_params(h, Wl.ctypes.data_as(P), Wr.ctypes.data_as(P), V.ctypes.data_as(P), vW.ctypes.data_as(P)), sb)
        q = rnd(2)
        call = lambda: check(sb.sb200_pack_server_answer(h, q.ctypes.data, resp.ctypes.data, None, None), sb)  # noqa: E731
    call(); call()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    db = sb.sb200_pack_server_db_bytes(h)
    print(json.dumps({"workload": f"{cfg} shape nu1={nu1} nu2={nu2} out_n={n} ({'direct upload' if direct else 'packed query + expansion'})",
                      "db_GiB": db / 2**30, "db_load_s": load_s, "e2e_ms_per_query_host_buffers": ms,
                      "db_GBs_if_scan_were_everything": db / (ms * 1e-3) / 1e9}))
    sb.sb200_pack_server_destroy(h)


if __name__ == "__main__":
    main()
