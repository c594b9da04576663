import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def checkers():
    """(ref, port): the compiled reference if its prebuilt .so is present, and the C restatement."""
    import oracle
    if not oracle.have_port() or (os.path.isdir(oracle.REFERENCE_ROOT) and not oracle.have_ref()):
        oracle.build()
    ref = oracle.Ref() if oracle.have_ref() else None
    return ref, oracle.Port()


@pytest.fixture(scope="session")
def comparand(checkers):
    """genFFT's own CPU output is the parity target; the bit-exact restatement stands in when the
    compiled reference did not travel."""
    ref, port = checkers
    return ref if ref is not None else port
