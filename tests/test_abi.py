"""CPU suite, part 2: the C-ABI library builds for sm_100a, loads without a GPU, exports every symbol
that include/spiral_b200.h declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes as C
import subprocess

import numpy as np
import pytest

from spiral_b200 import lib as sblib


def test_library_exports_every_declared_symbol(sb):
    declared = sblib.declared_symbols()
    assert len(declared) >= 50
    out = subprocess.run(["nm", "-D", "--defined-only", sblib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in include/spiral_b200.h but not exported: {missing}"
    bound = set(sb._sb200_signatures)
    assert not [s for s in declared if s not in bound], "python binding is missing declared entry points"


def test_no_torch_or_cxx_types_in_header():
    import re
    text = re.sub(r"/\*.*?\*/", "", open(sblib.HEADER_PATH).read(), flags=re.S)   # declarations only
    for bad in ("torch", "at::", "std::", "MatPoly", "vector<", "Tensor"):
        assert bad not in text


def test_abi_scalars(sb):
    assert sb.sb200_abi_version() == 1
    assert sb.sb200_arb_qprime(20) == 786433 and sb.sb200_arb_qprime(27) == 132120577   # include/values.h:74-76
    assert sb.sb200_db_words(8, 7) * 8 == 2 << 30                                           # cfg1: 2 GiB


def test_compute_fails_loudly_without_gpu(sb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    a = np.zeros(2048, dtype=np.uint64)
    out = np.zeros(4096, dtype=np.uint64)
    p64 = C.POINTER(C.c_uint64)
    rc = sb.sb200_to_ntt(out.ctypes.data_as(p64), a.ctypes.data_as(p64), 1)
    assert rc == -1, "must return SB200_ERR_NO_DEVICE, never compute on the CPU"
    assert b"no CPU fallback" in sb.sb200_last_error()
