"""Known-answer vectors of the wire formats and of the counter-based client (tests/golden/wire_kat.json).

A wire / on-disk format is a contract between machines: these digests freeze the byte-level behaviour of the plain-C statement
(oracle/wire_format.c, oracle/client_sim.c so_client_new_chacha); tests/test_oracle_wire.py checks the oracle against them and the
GPU suites check the CUDA code against the oracle, so an accidental change of either side shows up as a test failure.
    python scripts/make_wire_golden.py            # rewrites the fixture (only when a format version is bumped on purpose)
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import oracle_lib as ol  # noqa: E402


def digest(lib, a):
    a = np.ascontiguousarray(a)
    if a.dtype != np.uint64:
        pad = (-a.nbytes) % 8
        a = np.frombuffer(a.tobytes() + bytes(pad), dtype=np.uint64).copy()
    return f"{lib.so_fnv1a64(ol.ptr(a), a.size):016x}"


def vectors(lib):
    out = {}
    seed = bytes(range(32))
    row0 = np.zeros(2 * ol.N, dtype=np.uint64)
    lib.so_wire_seeded_row0(ol.ptr8(np.frombuffer(seed, dtype=np.uint8).copy()), ol.ptr(row0))
    out["seeded_row0(seed=00..1f)"] = {"digest": digest(lib, row0), "first": [int(x) for x in row0[:4]], "first_mod_b": [int(x) for x in row0[ol.N:ol.N + 4]]}
    s = ol.SpiralSession(lib, "cfg1", 3, 2, seed=5, chacha_seed=seed)
    sr, Sp = s.secret()
    out["chacha_client(cfg1,3,2,seed=00..1f)"] = {
        "sr": digest(lib, sr), "Sp": digest(lib, Sp),
        "W_exp_left": digest(lib, ol.canon(s.W_left, ol.KIND_NTT)), "W_exp_right": digest(lib, ol.canon(s.W_right, ol.KIND_NTT)),
        "W_conv": digest(lib, ol.canon(s.W_conv, ol.KIND_NTT)), "V_conv": digest(lib, ol.canon(s.V_conv, ol.KIND_NTT)),
        "sr_first": [int(x) for x in sr[:8]]}
    wire = s.chacha_query_wire(13, 7, bytes([9] * 32))
    out["chacha_query_wire(idx=13,query_id=7,wire_seed=09..09)"] = {"bytes": int(wire.size), "header": wire[:8].tolist(), "digest": digest(lib, wire)}
    full = np.zeros(lib.so_wire_query_bytes(ol.WIRE_FULL), dtype=np.uint8)
    lib.so_wire_query_pack_full(ol.ptr(ol.wire_expand(lib, wire)), ol.ptr8(full))
    out["full_wire_of_the_same_query"] = {"bytes": int(full.size), "header": full[:8].tolist(), "digest": digest(lib, full)}
    thr = np.zeros(128, dtype=np.uint64)
    lib.so_client_gaussian_thresholds(s.client, ol.ptr(thr))
    out["gaussian_thresholds"] = {"digest": digest(lib, thr), "t63_t64_t65": [int(thr[63]), int(thr[64]), int(thr[65])]}
    rec = np.arange(64, dtype=np.uint8)
    pts = np.zeros(32, dtype=np.uint64)
    lib.so_records_to_plaintexts(ol.ptr(pts), ol.ptr8(rec), 32, 65536)
    out["records_to_plaintexts(bytes 0..63, p=65536)"] = [int(x) for x in pts[:4]]
    s.close()
    ps = ol.PackSession(lib, "cfg3", 4, 2, False, seed=1, chacha_seed=seed)
    out["chacha_pack_client(cfg3,4,2,seed=00..1f)"] = {
        "v_W": digest(lib, ol.canon(ps.v_W, ol.KIND_NTT)), "V": digest(lib, ol.canon(ps.V, ol.KIND_NTT)),
        "W_exp_left": digest(lib, ol.canon(ps.W_left, ol.KIND_NTT)), "W_exp_right": digest(lib, ol.canon(ps.W_right, ol.KIND_NTT)),
        "query_wire(idx=5,query_id=1,wire_seed=07..07)": digest(lib, ps.chacha_query_wire(5, 1, bytes([7] * 32)))}
    ps.close()
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "wire_kat.json")
    with open(path, "w") as f:
        json.dump(vectors(ol.load()), f, indent=1)
    print("wrote", path)
