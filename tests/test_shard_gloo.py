"""CPU suite, part 4: the multi-GPU host logic on two `gloo` ranks.

Each rank takes the strided shard ii = rank (mod world) of the second dimension, runs the first
dimension and the local fold rounds on it (oracle arithmetic - no GPU here), all-gathers the one
surviving ciphertext and rank 0 runs the last log2(world) folds.  The result must equal the
unsharded pipeline bit for bit: this pins the shard ownership rule, the fold-round / GSW-dimension
bookkeeping and the gather order that bench.py and the tier-3 C-ABI use on NCCL."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import oracle_lib as ol

N = ol.N


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _inputs(nu1, nu2, t_gsw, seed):
    rng = np.random.default_rng(seed)
    dim0, num_per, m2 = 1 << nu1, 1 << nu2, 3 * t_gsw

    def packed(n):
        return (rng.integers(0, ol.P, size=n, dtype=np.uint64) | (rng.integers(0, ol.B, size=n, dtype=np.uint64) << np.uint64(32)))
    query = packed(dim0 * 2 * 4 * N).reshape(N, dim0, 2, 4)
    query[..., 3] = 0
    db = packed(N * num_per * 2 * dim0 * 2).reshape(N, num_per, 2 * dim0 * 2)       # B[z][ii][c,j,m]
    stride = 3 * m2 * 2 * N                                                            # reference stride per GSW dimension
    q = np.zeros(nu2 * stride, dtype=np.uint64)
    qn = np.zeros(nu2 * stride, dtype=np.uint64)
    for d in range(nu2):
        q[d * stride:d * stride + 3 * m2 * N] = packed(3 * m2 * N)
        qn[d * stride:d * stride + 3 * m2 * N] = packed(3 * m2 * N)
    return np.ascontiguousarray(query.reshape(-1)), db, q, qn


def _pipeline(lib, query, db_rows, dim0, q, qn, first_dim, t_gsw):
    """scan + lift + all fold rounds over the given rows; GSW dimensions start at first_dim."""
    num_per = db_rows.shape[1]
    scratch = np.zeros(num_per * 6 * 2 * N, dtype=np.uint64)
    dbc = np.ascontiguousarray(db_rows.reshape(-1))
    lib.so_multiply_query_by_database(ol.ptr(scratch), ol.ptr(query), ol.ptr(dbc), dim0, num_per)
    cts = np.zeros(max(num_per, 1) * 6 * N, dtype=np.uint64)
    lib.so_ntt_inv_and_crt_lift(ol.ptr(cts), ol.ptr(scratch), num_per)
    return _fold(lib, cts, num_per, q, qn, first_dim, t_gsw)


def _fold(lib, cts, count, q, qn, first_dim, t_gsw):
    np_, d = count, first_dim
    while np_ >= 2:
        np_ //= 2
        lib.so_fold_one_further_dimension(d, np_, ol.ptr(q), ol.ptr(qn), ol.ptr(cts), t_gsw)
        d += 1
    return cts[:6 * N].copy()


def _worker(rank, world, port, nu1, nu2, t_gsw, seed, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lib = ol.load()
    query, db, q, qn = _inputs(nu1, nu2, t_gsw, seed)
    shard = db[:, rank::world, :]                                  # ii = rank (mod world): strided ownership
    part = _pipeline(lib, query, shard, 1 << nu1, q, qn, 0, t_gsw)
    mine = torch.from_numpy(part.view(np.int64))
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)                                # one 96 KiB ciphertext per rank, order = rank
    if rank == 0:
        cts = np.concatenate([g.numpy().view(np.uint64) for g in gathered])
        log_w = world.bit_length() - 1
        final = _fold(lib, np.ascontiguousarray(cts), world, q, qn, nu2 - log_w, t_gsw)
        np.save(out_path, final)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nu2", [(2, 2), (2, 3)])
def test_strided_shards_reproduce_unsharded_fold(tmp_path, oracle, world, nu2):
    nu1, t_gsw, seed = 1, 8, 42
    out_path = str(tmp_path / "final.npy")
    mp.spawn(_worker, args=(world, _free_port(), nu1, nu2, t_gsw, seed, out_path), nprocs=world, join=True)
    got = np.load(out_path)
    query, db, q, qn = _inputs(nu1, nu2, t_gsw, seed)
    want = _pipeline(oracle, query, db, 1 << nu1, q, qn, 0, t_gsw)
    assert np.array_equal(got, want)


def test_contiguous_block_ownership_would_be_wrong(oracle):
    """Negative control: block (not strided) shards pair the wrong ciphertexts in the first rounds."""
    nu1, nu2, t_gsw, world = 1, 2, 8, 2
    query, db, q, qn = _inputs(nu1, nu2, t_gsw, 7)
    want = _pipeline(oracle, query, db, 1 << nu1, q, qn, 0, t_gsw)
    per = (1 << nu2) // world
    parts = [_pipeline(oracle, query, db[:, r * per:(r + 1) * per, :], 1 << nu1, q, qn, 0, t_gsw) for r in range(world)]
    got = _fold(oracle, np.ascontiguousarray(np.concatenate(parts)), world, q, qn, nu2 - 1, t_gsw)
    assert not np.array_equal(got, want)
