import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def pm():
    import __graft_entry__ as ge
    lib_missing = not os.path.exists(os.path.join(ge.PKG_DIR, "libpiet_metal_b200.so"))
    oracle_missing = not os.path.exists(os.path.join(ROOT, "oracle", "libpm_oracle.so"))
    harness_missing = not os.path.exists(os.path.join(ROOT, "tests", "native", "libpm_host_harness.so"))
    if lib_missing or oracle_missing or harness_missing:
        ge.build()
    return ge.load_package()


@pytest.fixture(scope="session")
def oracle(pm):
    import oracle_api
    return oracle_api
