"""The pass kernel exists in two builds: the pipelined one (large registers) and the synchronous one (small registers).
The launcher picks by register size; these tests force each build onto the sizes the other one normally serves, so
both are checked against the oracle over the whole parity suite.  The switch (QSV_ASYNC) is read once per process,
hence the subprocesses."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode,select", [("2", "golden or random or qft16 or x3sudoko or layered or rerun"), ("0", "layered or qft_closed_form"), ("2-nofast", "golden or qft16 or random")])
def test_parity_suite_with_forced_kernel(mode, select):
    if os.environ.get("QSV_VARIANT_INNER"):
        pytest.skip("inner run")
    env = dict(os.environ, QSV_VARIANT_INNER="1", QSV_ASYNC=mode.split("-")[0])
    if mode.endswith("nofast"):
        env["QSV_NO_FAST"] = "1"
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-k", select],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("mode", ["1", "2"])
def test_fused_basis_initialisation(mode):
    """Opt-in path (QSV_FUSED_INIT, pass_kernel_init.cu): written at the end of round 1 with no GPU time left, so it
    is not part of the default suite yet - run with QSV_TEST_FUSED_INIT=1 to validate it on hardware."""
    if os.environ.get("QSV_VARIANT_INNER"):
        pytest.skip("inner run")
    if not os.environ.get("QSV_TEST_FUSED_INIT"):
        pytest.skip("opt-in: set QSV_TEST_FUSED_INIT=1")
    env = dict(os.environ, QSV_VARIANT_INNER="1", QSV_FUSED_INIT=mode)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-k", "qft_closed_form or layered or golden or random"],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
