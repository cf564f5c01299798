"""The pass kernel exists in two builds: the pipelined TMA kernel (large registers, pass_kernel_tma.cu) and the
synchronous one (small registers, pass_kernel.cu).  The launcher picks by register size; these tests force each build
onto the sizes the other one normally serves, so both are checked against the oracle over the whole parity suite - and
the fused basis initialisation of the pipelined kernel in each of its modes.  The switches (QSV_ASYNC, QSV_FUSED_INIT)
are read once per process, hence the subprocesses."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_inner(env_extra, select):
    if os.environ.get("QSV_VARIANT_INNER"):
        pytest.skip("inner run")
    env = dict(os.environ, QSV_VARIANT_INNER="1", **env_extra)
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "--timeout", "300", "-k", select],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.parametrize("mode,select", [("2", "golden or random or qft16 or x3sudoko or layered or rerun or none_overwrite or x_gate_runs"),
                                         ("0", "layered or qft_closed_form"),
                                         ("2-nofast", "golden or qft16 or random or x_gate_runs")])
def test_parity_suite_with_forced_kernel(mode, select):
    env = {"QSV_ASYNC": mode.split("-")[0]}
    if mode.endswith("nofast"):
        env["QSV_NO_FAST"] = "1"
    run_inner(env, select)


@pytest.mark.parametrize("mode", ["0", "1", "2"])
def test_fused_basis_initialisation_modes(mode):
    """0: memset + ordinary first pass; 1: every tile synthesised and computed; 2 (default): zero tiles written by bulk
    stores.  Forced onto small registers as well (QSV_ASYNC=2) so the golden vectors and random circuits run through it."""
    run_inner({"QSV_FUSED_INIT": mode, "QSV_ASYNC": "2"}, "qft_closed_form or layered or golden or rerun or config3_depth100_live")


def test_dense_custom_round_stays_covered():
    """Multi-controlled Custom gates are lowered to controlled ops by default; with that switched off every Custom gate
    of the reference's tests goes through the dense (CSR) round again."""
    run_inner({"QSV_STRUCTURED_CUSTOM": "0"}, "x3sudoko or golden or none_overwrite or qft16 or post_select")


def test_folded_prefix_through_the_device_sub_register():
    """Leading gates on the top qubits of a basis state are folded into the initial amplitudes (plan.cpp build_plan): up to 14
    local qubits on the host, wider prefixes (QFT-33: 17) as a plan of their own on a sub-register on the device.  Here every
    prefix goes through the sub-register, and the ones of small registers are folded too."""
    run_inner({"QSV_HOST_PREFIX_BITS": "0"}, "qft_closed_form or layered or config3_depth100_live or grover")
    run_inner({"QSV_HOST_PREFIX_BITS": "0", "QSV_PREFIX_MIN_LOCAL": "0", "QSV_PREFIX_KEEP_BITS": "8", "QSV_ASYNC": "2"}, "golden or qft16 or rerun or x3sudoko")
    run_inner({"QSV_PREFIX_MIN_LOCAL": "0", "QSV_PREFIX_KEEP_BITS": "6"}, "golden or random or qft16 or none_overwrite")
