import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, 'baseline') not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))       # ref_harness: runs the unmodified reference from baseline/_ref


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDEN = os.path.join(ROOT, 'tests', 'golden')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
