import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200, sm_100)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def chess_root():
    import nbvh_b200
    root = nbvh_b200.chess_fixture_dir()
    if root is None:
        pytest.skip("Chess fixture (baseline/_ref/chess) not present; run tools/make_fixtures.py in the authoring container")
    return root


@pytest.fixture(scope="session")
def synthetic():
    import nbvh_b200
    return nbvh_b200.synthetic_scene()
