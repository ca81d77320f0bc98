import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def capi_mod():
    from acvd_b200 import build, capi
    if not os.path.exists(capi.LIB_PATH):
        build.build_library()
    return capi


@pytest.fixture(scope="session")
def gpu_ctx_factory(capi_mod):
    """Creates contexts on cuda:0; fails loudly (no CPU fallback) when there is no device."""
    made = []

    def make():
        ctx = capi_mod.Context(0)
        made.append(ctx)
        return ctx

    yield make
    for c in made:
        c.close()
