import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _native_libs():
    """Host-side native libraries are built once per session (seconds); the CUDA library is built by
    __graft_entry__.build() / on first use."""
    from pingpongpro_b200 import _build

    _build.build_host()
    _build.build_synth()
    _build.build_oracle()
