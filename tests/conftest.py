"""pytest configuration: `gpu` marker, oracle fixtures, shared tolerances.

The oracle (oracle/) is the CHECKER in these tests; the thing under test on the
`-m gpu` side is the CUDA path reached through the C ABI (include/staar_b200.h).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, 'tests')):
    if _p not in sys.path:
        sys.path.insert(0, _p)



def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle, build
    build(with_ref=True)
    return Oracle("port")


@pytest.fixture(scope="session")
def reference(oracle):
    from oracle.pyoracle import Oracle, have_reference
    if not have_reference():
        pytest.skip("oracle/_ref/libstaar_ref.so not built (needs /root/reference)")
    return Oracle("reference")


@pytest.fixture(scope="session")
def gpu_ctx():
    """The CUDA path under test, reached through the C ABI.  No fallback: a missing library or device FAILS."""
    from staarpipeline_b200.api import Context
    ctx = Context(0)
    yield ctx
    ctx.close()
