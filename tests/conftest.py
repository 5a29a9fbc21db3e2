import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / 'tests'), str(ROOT / 'tests' / 'golden')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line(
        'markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def golden():
    from helpers import load_golden
    return load_golden()


@pytest.fixture(scope='session')
def dsp_golden():
    from helpers import load_dsp_golden
    return load_dsp_golden()


@pytest.fixture(scope='session')
def ns():
    from helpers import b200_namespace
    return b200_namespace()
