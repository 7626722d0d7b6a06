import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


@pytest.fixture(scope='session', autouse=True)
def _built_library():
    """The C-ABI library must exist for both suites (the CPU suite only loads it and checks its exports)."""
    import __graft_entry__
    __graft_entry__.build()
