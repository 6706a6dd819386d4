import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def state_dict():
    from dual_space_nerf_b200 import net as N

    return N.synthetic_net(0).state_dict()


@pytest.fixture(scope="session")
def scene64():
    from dual_space_nerf_b200 import scene as S

    return S.make_scene(64, 64)
