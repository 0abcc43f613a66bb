"""GPU suite: a real client/server split through files - the C++ hosts spiral_b200/csrc/host/pir_client.cpp and pir_server.cpp
(plain g++ above the C-ABI, no Python, no oracle on the data path).  Client: seed -> public-parameter file, wire queries,
decoded items; server: record file (or snapshot) + public parameters + wire queries -> packed responses.  The decoded item
must be the bytes of the record file at that index (the reference's "Is correct?: 1" across process boundaries)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "spiral_b200", "csrc", "host")


def run(*cmd):
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, f"{' '.join(cmd)}\n{r.stdout}\n{r.stderr}"
    return r


@pytest.mark.parametrize("params,item_bytes", [("5,3,8,4,8,56,20,256", 8192), ("4,2,9,4,8,56,21,256", 8192)])
def test_client_and_server_processes_exchange_files(sb, tmp_path, params, item_bytes):
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    client, server = os.path.join(HOST, "pir_client"), os.path.join(HOST, "pir_server")
    nu1, nu2 = (int(x) for x in params.split(",")[:2])
    total = 1 << (nu1 + nu2)
    rng = np.random.default_rng(9)
    records = rng.integers(0, 256, total * item_bytes, dtype=np.uint8)
    t = lambda name: str(tmp_path / name)  # noqa: E731
    records.tofile(t("records.bin"))
    np.arange(32, dtype=np.uint8).tofile(t("seed.bin"))
    run(client, "keygen", "--params", params, "--seed", t("seed.bin"), "--out", t("pp.bin"))
    idxs = [0, total - 1, total // 3]
    for k, idx in enumerate(idxs):
        np.full(32, 50 + k, dtype=np.uint8).tofile(t(f"ws{k}.bin"))
        run(client, "query", "--params", params, "--seed", t("seed.bin"), "--idx", str(idx), "--query-id", str(k), "--wire-seed", t(f"ws{k}.bin"),
            "--out", t(f"q{k}.bin"))
        assert os.path.getsize(t(f"q{k}.bin")) == 8 + 32 + 14336
    # first server process: from the record file, also writes a snapshot; second one: from the snapshot
    r = run(server, "--params", params, "--db", t("records.bin"), "--save-snapshot", t("db.sb2d"), "--pp", t("pp.bin"),
            "--query", t("q0.bin"), "--query", t("q1.bin"), "--out-prefix", t("resp"))
    assert "answered" in r.stderr
    run(server, "--params", params, "--snapshot", t("db.sb2d"), "--pp", t("pp.bin"), "--query", t("q2.bin"), "--out-prefix", t("resp2"))
    os.replace(t("resp2.0"), t("resp.2"))
    for k, idx in enumerate(idxs):
        run(client, "decode", "--params", params, "--seed", t("seed.bin"), "--in", t(f"resp.{k}"), "--out", t(f"item{k}.bin"))
        got = np.fromfile(t(f"item{k}.bin"), dtype=np.uint8)
        assert np.array_equal(got, records[idx * item_bytes:(idx + 1) * item_bytes]), f"item {idx} not recovered"
    # a query made under another key decodes to something else (the check above is not vacuous)
    np.arange(1, 33, dtype=np.uint8).tofile(t("seed2.bin"))
    run(client, "decode", "--params", params, "--seed", t("seed2.bin"), "--in", t("resp.0"), "--out", t("wrong.bin"))
    assert not np.array_equal(np.fromfile(t("wrong.bin"), dtype=np.uint8), records[:item_bytes])
    # errors are loud
    bad = subprocess.run([server, "--params", params, "--db", t("seed.bin"), "--pp", t("pp.bin"), "--out-prefix", t("x")], capture_output=True, text=True)
    assert bad.returncode != 0 and "record stream" in bad.stderr
