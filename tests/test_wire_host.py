"""CPU suite: the host-side pieces of the wire formats and of the GPU client that need no device - format sizes and the Gaussian
sampler's integer thresholds agree between the product (C-ABI) and the oracle's plain-C statement, and the C++ hosts of the
client/server split refuse to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests import oracle_lib as ol

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "spiral_b200", "csrc", "host")


def test_wire_query_sizes_agree_with_the_oracle_and_the_reference_accounting(sb, oracle):
    for kind in range(0, 5):
        assert sb.sb200_wire_query_bytes(kind) == oracle.so_wire_query_bytes(kind)
    b_per_elem = 2048 * 56 // 8                                  # print_summary, src/spiral.cpp:219
    assert sb.sb200_wire_query_bytes(1) == 8 + 32 + b_per_elem   # "query_size": 14336.0 in all_parameter_choices.txt + seed + header
    assert sb.sb200_wire_query_bytes(2) == 8 + 2 * b_per_elem


def test_gaussian_thresholds_agree_on_the_host(sb, oracle):
    """Both sides derive the 128 thresholds from libm's exp(); the CUDA sampler and the plain-C client must walk the same table."""
    s = ol.SpiralSession(oracle, "cfg1", 2, 1, chacha_seed=bytes(32))
    want = np.zeros(128, dtype=np.uint64)
    oracle.so_client_gaussian_thresholds(s.client, ol.ptr(want))
    got = np.zeros(128, dtype=np.uint64)
    assert sb.sb200_client_gaussian_thresholds(ol.ptr(got)) == 0
    assert np.array_equal(got, want)
    # width 6.4 (sigma = 2.55): P(|x| <= 2) = 0.6756, P(|x| <= 6) = 0.9896 (cdf[k] = P(x <= k - 64))
    cdf = got.astype(np.float64) / 2.0 ** 53
    assert abs((cdf[66] - cdf[61]) - 0.6756) < 1e-3 and abs((cdf[70] - cdf[57]) - 0.9896) < 1e-3
    s.close()


def test_packed_response_sizes(sb, oracle):
    # row 0 at QPBITS, rows 1-2 at log2(4p) bits (src/spiral.cpp:229-232): cfg1 = 2*2048*20 + 4*2048*10 bits = 20 480 B
    assert sb.sb200_packed_response_words(2 * 2048, 4 * 2048, 20, 256) * 8 == 20480
    assert sb.sb200_packed_words(2048, 56) * 8 == 14336


def test_cxx_hosts_fail_loudly_without_a_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    seed = tmp_path / "seed.bin"
    seed.write_bytes(bytes(32))
    r = subprocess.run([os.path.join(HOST, "pir_client"), "keygen", "--params", "4,2,8,4,8,56,20,256", "--seed", str(seed), "--out", str(tmp_path / "pp.bin")],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    r = subprocess.run([os.path.join(HOST, "pir_server"), "--params", "4,2,8,4,8,56,20,256", "--db", str(seed), "--pp", str(seed), "--out-prefix", "x"],
                       capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
