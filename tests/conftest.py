import os
import sys

import pytest

# Several tests run up to 8 shards of a sharded server on ONE device, each on its own streams, with kernels that wait for each
# other's flags.  With the default 8 hardware queues distinct streams can share a queue, and a waiting kernel then blocks the very
# kernel it waits for.  Must be set before the CUDA context exists; one process per GPU (the deployed form) never needs it.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# ... and lazy kernel loading may need a context-wide synchronisation that a spinning kernel of the same context never allows
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Run order: the oracle's own pins first, then the hot path from its leaves outwards (parity cases, whole queries, full
# sizes, Pack, tensor cores, the reference harness), then the rows of SURVEY 8f that sit either side of the path.
_ORDER = ["test_abi", "test_cost_model", "test_oracle_golden", "test_oracle_e2e", "test_oracle_wire", "test_wire_host", "test_shard_gloo", "test_gpu_parity", "test_gpu_e2e",
          "test_gpu_fullsize", "test_gpu_pack", "test_gpu_tc", "test_gpu_dropin", "test_gpu_wire", "test_gpu_client", "test_gpu_cli", "test_gpu_pack_client"]


def pytest_collection_modifyitems(session, config, items):
    def key(item):
        name = os.path.splitext(os.path.basename(str(item.fspath)))[0]
        return _ORDER.index(name) if name in _ORDER else len(_ORDER)
    items.sort(key=key)            # stable: the order inside a file is kept


@pytest.fixture(scope="session")
def oracle():
    from tests import oracle_lib
    return oracle_lib.load()


@pytest.fixture(scope="session")
def sb():
    """The product library, through its C-ABI.  Never falls back to anything else."""
    from spiral_b200 import lib
    lib.build()
    return lib.load_library()
