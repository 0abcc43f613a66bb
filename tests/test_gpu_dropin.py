"""GPU suite: drop-in check.  The UNMODIFIED reference harness (oracle/_ref/libspiral_ref_*.so: its own
client, key generation, query, decoding and "Is correct?" gate) runs with the reference's server
functions re-defined by spiral_b200/csrc/host/host_mirror.cpp, i.e. bound to the CUDA path through
the C-ABI.  SB200_PARITY=1 additionally runs the reference's own function behind every mirrored call
and aborts on the first differing byte."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _driver(cfg):
    flags = open("/proc/cpuinfo").read()
    isa = "avx512" if all(f" {x}" in flags for x in ("avx512f", "avx512dq", "avx512bw", "avx512vl")) else "avx2"
    return os.path.join(ROOT, "oracle", "_ref", f"spiral_b200_driver_{cfg}_{isa}")


@pytest.mark.parametrize("cfg,args", [("cfg1", ["6", "3", "77"]), ("cfg1", ["7", "4", "1500"]), ("cfg5", ["5", "3", "100"])])
def test_reference_harness_with_cuda_server(cfg, args):
    exe = _driver(cfg)
    if not os.path.exists(exe):
        pytest.skip("prebuilt reference driver not present (built only where /root/reference exists)")
    env = dict(os.environ, SB200_PARITY="1")
    out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Is correct?: 1" in out.stdout, out.stdout[-2000:]
    for leaf in ("reorientCiphertexts", "multiplyQueryByDatabase", "nttInvAndCrtLiftCiphertexts", "foldOneFurtherDimension",
                 "expandImproved", "regevToGSW", "scalToMat", "modswitch", "getRescaled"):
        assert f"parity ok: {leaf}" in out.stderr, f"{leaf} was not exercised:\n{out.stderr[-2000:]}"
    assert "PARITY FAIL" not in out.stderr
    assert "database resident on the GPU" in out.stderr


@pytest.mark.parametrize("cfg,args,leaves", [
    ("cfg3", ["5", "2", "77", "a", "--high-rate"],
     ("convertDb", "coefficientExpansion", "reorientCiphertextsDim1", "regevToSimpleGsw", "fastMultiplyQueryByDatabaseDim1",
      "foldCiphertextsDim1", "pack", "getRescaled")),
    ("cfg4", ["6", "3", "100", "a", "--high-rate", "--direct-upload"],
     ("convertDb", "fastMultiplyQueryByDatabaseDim1", "foldCiphertextsDim1", "pack", "getRescaled")),
    ("cfg1", ["6", "2", "9", "a", "--high-rate"],
     ("convertDb", "coefficientExpansion", "regevToSimpleGsw", "fastMultiplyQueryByDatabaseDim1", "foldCiphertextsDim1", "pack")),
])
def test_reference_pack_harness_with_cuda_server(cfg, args, leaves):
    """SpiralPack / SpiralStreamPack: the reference's testHighRate (src/testing.cpp:777) drives the CUDA path through
    host_mirror.cpp's definitions of its own leaf functions; the planes convertDb produced stay resident in HBM."""
    exe = _driver(cfg)
    if not os.path.exists(exe):
        pytest.skip("prebuilt reference driver not present (built only where /root/reference exists)")
    env = dict(os.environ, SB200_PARITY="1")
    out = subprocess.run([exe] + args, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Is correct? : 1" in out.stdout, out.stdout[-2000:]
    for leaf in leaves:
        assert f"parity ok: {leaf}" in out.stderr, f"{leaf} was not exercised:\n{out.stderr[-2000:]}"
    assert "PARITY FAIL" not in out.stderr
    assert "resident on the GPU" in out.stderr
    assert "CUDA kernel launches" in out.stderr
