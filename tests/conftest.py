import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """the CPU oracle (test infrastructure; built on demand with gcc)"""
    from oracle import splat_oracle

    splat_oracle.lib()
    return splat_oracle


@pytest.fixture(scope="session")
def splat():
    """the product package; the library must already be built (python -m splatapult_b200.build)"""
    import splatapult_b200

    if not os.path.exists(splatapult_b200.library_path()):
        from splatapult_b200 import build

        build.build()
    splatapult_b200.lib()
    return splatapult_b200


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def scratch(tmp_path_factory):
    return tmp_path_factory.mktemp("splat")
