import os
import random
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One libb200g16 context on cuda:0 for the whole GPU session (fails loudly without a GPU)."""
    from gnark_whir_b200 import lib
    c = lib.Context(0)
    yield c
    c.close()


@pytest.fixture
def rng():
    return random.Random(20261018)
