import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# GENFFT_TEST_BACKEND=emu: the -m gpu tests run against the kernel-logic emulator (tests/emu/backend.py; test
# infrastructure that executes the real kernel sources on host fibers).  Only tests/test_emu_suite.py sets it, for a
# pytest subprocess; a GPU box never does, and genfft_b200 itself cannot load the emulator.
EMU = os.environ.get("GENFFT_TEST_BACKEND") == "emu"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "no_emu: meaningless or too large for the kernel-logic emulator")
    if EMU:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from emu import backend
        backend.install()


# -m gpu tests the emulator run leaves out: sizes that take the host fibers more than a few seconds, binaries linked
# against the real library, the bench contract, NCCL process groups.
EMU_SKIP = re.compile("|".join([
    r"test_gpu_cpp\.py", r"test_gpu_bench_contract\.py", r"test_gpu_dist\.py", r"test_dist_fft1d_on_one_rank",
    r"test_c4_real_2pow22", r"test_c2_full_size", r"beyond_the_reference_maximum", r"test_c3_2pow24", r"test_c2c_large",
    r"chain_is_bit_identical\[(True|False)-float(32|64)-2[0-9]-", r"test_fft2d_chain_is_bit_identical\[.*(65536|4096-4096)",
    r"unfused_split_path\[4096-70000", r"test_r2c_chain_is_bit_identical\[2[0-9]-",
    r"test_c5_shape_2d", r"test_chain_group_size_and_lag\[(16384-2|1024-3|65536-1)", r"\[256-4096-float", r"\[256-8192-float",
    r"test_real_fft_vs_reference\[22-",
]))


def pytest_collection_modifyitems(config, items):
    if not EMU:
        return
    skip = pytest.mark.skip(reason="not run on the emulator")
    for item in items:
        if item.get_closest_marker("no_emu") or EMU_SKIP.search(item.nodeid):
            item.add_marker(skip)


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
