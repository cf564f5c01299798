import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """A device-side protocol error must fail a test, not hang the box: every GPU test gets a wall-clock limit
    (pytest-timeout; the kernels themselves trap after a few seconds of waiting on a barrier that never completes)."""
    try:
        import pytest_timeout  # noqa: F401
    except Exception:
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(900))


@pytest.fixture(scope="session", autouse=True)
def _built_artifacts():
    """Builds the oracle, the emulation harness and libqsv.so if they are missing (CPU box: nvcc cross-compiles)."""
    import subprocess
    need = [os.path.join(ROOT, "oracle", "liboracle.so"), os.path.join(ROOT, "tests", "emu", "libqsv_emu.so"),
            os.path.join(ROOT, "quantr_b200", "libqsv.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-C", ROOT, "all"], check=True, capture_output=True)
