import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def cpu_oracle():
    """oracle/sht_cpu.c built on demand (checker only)."""
    from oracle import sht_cpu
    sht_cpu.build()
    return sht_cpu


@pytest.fixture(scope="session")
def shtlib():
    """libcmdr_sht.so through the sharp.f90 mirror; built on demand."""
    lib_path = os.path.join(ROOT, "commander_b200", "lib", "libcmdr_sht.so")
    if not os.path.exists(lib_path):
        import __graft_entry__
        __graft_entry__.build()
    from commander_b200 import sharp
    sharp.lib()
    return sharp
