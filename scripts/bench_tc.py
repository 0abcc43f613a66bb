"""Time the tensor-core batched first dimension (k_scan_tc) against the single-query scan at a cfg1-shaped database.
usage: python scripts/bench_tc.py [nu1 nu2]   (default 8 7 = 2 GiB).  Prints one JSON line per batch size."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spiral_b200 import lib  # noqa: E402

N = 2048
nu1, nu2 = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 7)
dim0, num_per = 1 << nu1, 1 << nu2
sb = lib.load_library()
assert sb.sb200_init(0) == 0, sb.sb200_last_error()
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0


def rnd_pb(n):
    g = torch.Generator(device="cuda"); g.manual_seed(n & 0xffff)
    return torch.randint(0, 1 << 28, (n,), dtype=torch.int64, device="cuda", generator=g) | (
        torch.randint(0, 1 << 27, (n,), dtype=torch.int64, device="cuda", generator=g) << 32)


db_words = sb.sb200_db_words(nu1, nu2)
db = rnd_pb(db_words)
db_tc = torch.empty(db_words * 8, dtype=torch.uint8, device="cuda")
assert sb.sb200_dev_db_to_tc(db_tc.data_ptr(), db.data_ptr(), dim0, num_per, None) == 0, sb.sb200_last_error()
torch.cuda.synchronize()
out_words = num_per * 6 * 2 * N
qmask = torch.tensor([1, 1, 1, 0], dtype=torch.int64, device="cuda").repeat(N * dim0 * 2)
queries = [rnd_pb(N * dim0 * 2 * 4 + b) [: N * dim0 * 2 * 4] * qmask for b in range(16)]
outs = [torch.empty(out_words, dtype=torch.int32, device="cuda") for _ in range(16)]
ref = torch.empty(out_words, dtype=torch.int32, device="cuda")


def timed(fn, iters=None, warm=3):
    iters = iters or int(os.environ.get("TC_ITERS", "10"))
    for _ in range(warm):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(iters))
    return ts[0], ts[len(ts) // 2]


ts = timed(lambda: sb.sb200_dev_first_dim(ref.data_ptr(), queries[0].data_ptr(), db.data_ptr(), dim0, num_per, None))
print(json.dumps({"kernel": "k_scan_spiral", "queries": 1, "ms_min": ts[0], "ms_med": ts[1], "db_gbs": db_words * 8 / ts[1] / 1e6}))
COUNTS = [int(c) for c in os.environ.get("TC_COUNTS", "1,4,5,8,10,12,16").split(",")]
ITERS = int(os.environ.get("TC_ITERS", "10"))
for count in COUNTS:
    q_tc = torch.zeros(sb.sb200_tc_query_bytes(dim0, count), dtype=torch.uint8, device="cuda")
    for b in range(count):
        assert sb.sb200_dev_query_to_tc(q_tc.data_ptr(), queries[b].data_ptr(), b, count, dim0, None) == 0, sb.sb200_last_error()
    arr = (C.c_void_p * count)(*[outs[b].data_ptr() for b in range(count)])
    t1 = torch.empty(sb.sb200_tc_scratch_bytes(num_per, count), dtype=torch.uint8, device="cuda")
    tq = timed(lambda: sb.sb200_dev_query_to_tc(q_tc.data_ptr(), queries[0].data_ptr(), 0, count, dim0, None))
    t = timed(lambda: sb.sb200_dev_first_dim_tc(arr, count, count, q_tc.data_ptr(), db_tc.data_ptr(), dim0, num_per, t1.data_ptr(), None))
    same = bool(torch.equal(outs[0], ref))
    bytes_moved = db_words * 8 + q_tc.numel() + count * out_words * 4
    print(json.dumps({"kernel": "k_scan_tc", "queries": count, "ms_min": t[0], "ms_med": t[1], "ms_per_query": t[1] / count,
                      "db_gbs_per_pass": db_words * 8 / t[1] / 1e6, "db_gbs_x_queries": count * db_words * 8 / t[1] / 1e6,
                      "frac_of_hbm_peak_per_pass": db_words * 8 / t[1] / 1e6 / peak, "all_bytes_gbs": bytes_moved / t[1] / 1e6,
                      "tensor_tops": 2.0 * N * 2 * (num_per * 2) * (dim0 * 2) * 16 * 3 * count / t[1] / 1e9,
                      "query_to_tc_ms_each": tq[1], "query0_equals_single_scan": same}))
