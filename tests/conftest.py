import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "offline"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE = "/root/reference"  # present in the build container only; GPU-box tests never need it


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA GPU (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def tables():
    from noahmp_b200.tables import ParamTables
    return ParamTables.default()


@pytest.fixture(scope="session")
def oracle_libs():
    """Build the three oracle variants once (test infrastructure; the product never loads them)."""
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "all"], check=True)
    return os.path.join(ROOT, "oracle")
