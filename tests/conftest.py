import importlib
import os
import sys

import pytest

os.environ.setdefault("OMP_NUM_THREADS", "1")  # reference/oracle parity runs are single-threaded

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_pkg():
    """import the product package (directory name has a hyphen)"""
    return importlib.import_module("md-bench_b200")


@pytest.fixture(scope="session")
def mdb():
    return load_pkg()


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def ref_usable(variant="vl_dp_aos"):
    """the prebuilt reference library is present AND this CPU can run it (built -march=x86-64-v4)"""
    from refbind import ref_available
    if not ref_available(variant):
        return False
    try:
        with open("/proc/cpuinfo") as f:
            return "avx512f" in f.read()
    except OSError:
        return False
