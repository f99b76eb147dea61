import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# `dump` commands of the shipped input scripts write relative to the working directory, as LAMMPS does: keep the
# repository clean while testing
os.environ.setdefault("SEDI_DUMP_DIR", __import__("tempfile").mkdtemp(prefix="sedi_dump_"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    """the CPU checker (test infrastructure only)"""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle
