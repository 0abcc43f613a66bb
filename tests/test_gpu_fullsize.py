"""GPU suite: BASELINE.json's full cfg1 size (2^20 x 256 B = 2^15 plaintext matrices, 2 GiB database in HBM),
checked through size-independent properties - the oracle would need minutes here:
  * selection: a first-dimension "query" that is the constant 1 in lane (j0, m0, r0) must return, after the
    scan and the INTT + CRT lift, exactly the centred plaintext polynomials of items j0*num_per + i
    (database preprocessing + scan + lift at full size, bit-exact against the planted data);
  * linearity of the scan modulo both primes: scan(q1 + q2) = scan(q1) + scan(q2);
  * error behaviour of the resident server (no silent fallbacks)."""
import ctypes as C

import numpy as np
import pytest

from spiral_b200 import SB200Error, SpiralParams
from spiral_b200.server import SpiralServer

pytestmark = pytest.mark.gpu
N, P, B = 2048, 268369921, 249561089
Q = P * B
CFG1 = dict(t_gsw=8, t_conv=4, t_exp=8, t_exp_right=56, qp_bits=20, out_n=2, p_db=256)


def params(nu1, nu2):
    return SpiralParams(nu1, nu2, CFG1["t_gsw"], CFG1["t_conv"], CFG1["t_exp"], CFG1["t_exp_right"], CFG1["qp_bits"], CFG1["out_n"], CFG1["p_db"])


def pack(lo, hi):
    return lo.astype(np.uint64) | (hi.astype(np.uint64) << np.uint64(32))


@pytest.fixture(scope="module")
def full_server(sb):
    import torch
    nu1, nu2 = 8, 7
    rng = np.random.default_rng(2022)
    pts = rng.integers(0, 256, size=(1 << (nu1 + nu2), 4, N), dtype=np.uint16)      # 512 MiB of plaintext
    srv = SpiralServer(params(nu1, nu2))
    srv.load_db_items(pts)
    assert srv.db_bytes == 2 << 30
    yield srv, pts, torch
    srv.close()


def run_scan(sb, srv, torch, query_np, lift=True):
    dim0, num_per = srv.dim0, srv.num_per
    q = torch.from_numpy(query_np.view(np.int64)).cuda()
    out = torch.empty(num_per * 6 * 2 * N, dtype=torch.int32, device="cuda")
    db_ptr = sb.sb200_server_db_ptr(srv.h)
    assert sb.sb200_dev_first_dim(out.data_ptr(), q.data_ptr(), db_ptr, dim0, num_per, None) == 0, sb.sb200_last_error()
    if not lift:
        torch.cuda.synchronize()
        return out.cpu().numpy().view(np.uint32)
    raw = torch.empty(num_per * 6 * N, dtype=torch.int64, device="cuda")
    assert sb.sb200_dev_from_ntt(raw.data_ptr(), out.data_ptr(), num_per * 6, None) == 0, sb.sb200_last_error()
    torch.cuda.synchronize()
    return raw.cpu().numpy().view(np.uint64)


def test_selection_query_returns_planted_items(sb, full_server):
    srv, pts, torch = full_server
    dim0, num_per = srv.dim0, srv.num_per
    j0, m0, r0 = 5, 1, 2
    query = np.zeros((N, dim0, 2, 4), dtype=np.uint64)
    query[:, j0, m0, r0] = np.uint64(1 | (1 << 32))                 # the constant polynomial 1 in NTT form, both primes
    raw = run_scan(sb, srv, torch, np.ascontiguousarray(query.reshape(-1))).reshape(num_per, 3, 2, N)
    for i in (0, 17, num_per - 1):
        item = pts[j0 * num_per + i].astype(np.int64)               # (4, N): entry (m, c) at m*2 + c
        for c in range(2):
            v = item[m0 * 2 + c]
            want = np.where(v >= 128, Q - (256 - v), v).astype(np.uint64)   # centre-lift, src/spiral.cpp:1116-1127
            assert np.array_equal(raw[i, r0, c], want), f"item {j0 * num_per + i} entry ({m0},{c})"
        assert not raw[i, [0, 1]].any(), "rows that the query does not select must be zero"


def test_scan_is_linear_at_full_size(sb, full_server):
    srv, pts, torch = full_server
    rng = np.random.default_rng(5)
    shape = (N * srv.dim0 * 2, 4)

    def rnd():
        lo, hi = rng.integers(0, P, size=shape, dtype=np.uint64), rng.integers(0, B, size=shape, dtype=np.uint64)
        lo[:, 3] = 0; hi[:, 3] = 0
        return lo, hi
    (a_lo, a_hi), (b_lo, b_hi) = rnd(), rnd()
    s_lo, s_hi = (a_lo + b_lo) % np.uint64(P), (a_hi + b_hi) % np.uint64(B)
    oa = run_scan(sb, srv, torch, np.ascontiguousarray(pack(a_lo, a_hi).reshape(-1)), lift=False).astype(np.uint64).reshape(-1, 2, N)
    ob = run_scan(sb, srv, torch, np.ascontiguousarray(pack(b_lo, b_hi).reshape(-1)), lift=False).astype(np.uint64).reshape(-1, 2, N)
    os_ = run_scan(sb, srv, torch, np.ascontiguousarray(pack(s_lo, s_hi).reshape(-1)), lift=False).astype(np.uint64).reshape(-1, 2, N)
    assert np.array_equal((oa[:, 0] + ob[:, 0]) % np.uint64(P), os_[:, 0])
    assert np.array_equal((oa[:, 1] + ob[:, 1]) % np.uint64(B), os_[:, 1])
    assert os_.max() < P


def test_server_error_paths(sb):
    with pytest.raises(SB200Error):
        SpiralServer(params(4, 2), world=3)                      # world must be a power of two
    with pytest.raises(SB200Error):
        SpiralServer(params(4, 1), rank=0, world=4)              # more shards than second-dimension entries
    srv = SpiralServer(params(3, 2))
    q = np.zeros(2 * 2 * N, dtype=np.uint64)
    with pytest.raises(SB200Error, match="public parameters"):
        srv.answer(q)                                            # no keys yet
    z = np.zeros(1, dtype=np.uint64)
    g, n_right = 6, 6
    srv.set_public_params(np.zeros(g * 2 * 8 * 2 * N, dtype=np.uint64), np.zeros(n_right * 2 * 56 * 2 * N, dtype=np.uint64),
                          np.zeros(3 * 8 * 2 * N, dtype=np.uint64), np.zeros(3 * 8 * 2 * N, dtype=np.uint64))
    with pytest.raises(SB200Error, match="database not loaded"):
        srv.answer(q)
    srv.close()
    assert sb.sb200_dev_first_dim(None, None, None, 3, 4, None) == -3       # dim0 not a power of two -> SB200_ERR_ARG
    del z
