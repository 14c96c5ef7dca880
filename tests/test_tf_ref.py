"""oracle/tf_ref.py: the reference's own wflib as the pin of the golden vectors and as the timed baseline.

Here (no TensorFlow) the replay runs on oracle/tf_shim and must reproduce every committed vector bit for bit -- that pins
the replay logic (fixture keys, call signatures, the GradientTape plumbing).  Where `import tensorflow` gives the real
library the second test holds the vectors to TensorFlow's own kernels at 1e-6; it is skipped, not absent, elsewhere."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import tf_ref  # noqa: E402

needs_reference = pytest.mark.skipif(tf_ref.find_reference() is None, reason="no reference checkout on this machine")


@needs_reference
def test_replay_of_the_reference_on_the_shim_reproduces_the_goldens_bit_for_bit():
    # separate interpreter: importing the shim as `tensorflow` must not leak into this test session
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "tf_ref.py")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "PINNED" in out.stdout and out.stdout.count("0.00e+00") >= 18


@needs_reference
@pytest.mark.skipif(tf_ref.real_tensorflow() is None, reason="TensorFlow is not installed: the oracle stays pinned to the reference source on the shim")
def test_real_tensorflow_pins_the_goldens():
    ref = tf_ref.Reference()
    assert ref.kind == "tf"
    worst = tf_ref.check_goldens(ref)
    bad = {k: v for k, v in worst.items() if v > 1e-6}
    assert not bad, bad


def test_bench_reference_arm_reports_which_implementation_ran():
    """bench.py --impl reference prefers the real reference under TensorFlow and says so in `cpu_baseline.kind`."""
    import bench
    kind, why = bench.reference_kind()
    assert kind in ("tf", "port")
    if tf_ref.real_tensorflow() is None:
        assert kind == "port" and "TensorFlow" in why
