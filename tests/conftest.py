import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def ednafull():
    from oracle import oracle as O
    return O.make_matrix()   # EDNAFULL restricted to A,C,G,T,N (Align.pyx:63-99 documents the equality)
