"""examples/another_example.py (the reference's another-example.py recipe as a launch script): runs, learns, writes
TensorFlow-format checkpoints under the names a Saver over the reference's graph would use, and --resume continues a run
that was stopped MID-WINDOW exactly as if it had never stopped (another-example.py:126-155, 209, 323-327)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "examples", "another_example.py")


def _run(*args):
    r = subprocess.run([sys.executable, SCRIPT, *args], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.gpu
def test_another_example_learns_and_resumes_mid_window_from_a_tf_checkpoint(tmp_path):
    from gaccum_b200 import tf_checkpoint as ck
    a, b = str(tmp_path / "straight"), str(tmp_path / "interrupted")
    _run("--steps", "11", "--model-dir", a)
    out = _run("--steps", "5", "--model-dir", b)                     # N = 3: applies at g = 0, 3; g = 4 only accumulated
    assert "saved model.ckpt-5" in out
    mid = ck.read_bundle(ck.latest_checkpoint(b))
    # the keys of the reference's Saver: Keras layer names, TF1 Adam slots and beta powers, unnamed accumulator variables
    want = {"global_step", "beta1_power", "beta2_power"}
    for i, layer in enumerate(["dense", "dense_1", "dense_2", "dense_3"]):
        for j, kind in enumerate(["kernel", "bias"]):
            k = 2 * i + j
            want |= {f"{layer}/{kind}", f"{layer}/{kind}/Adam", f"{layer}/{kind}/Adam_1", "Variable" if k == 0 else f"Variable_{k}"}
    assert set(mid) == want
    assert mid["dense/kernel"].shape == (13, 16) and mid["dense_3/kernel"].shape == (4, 1) and int(mid["global_step"]) == 5
    assert np.any(mid["Variable"] != 0)                               # stopped mid-window: the accumulators are live
    assert abs(float(mid["beta1_power"]) - 0.729) < 1e-6 and abs(float(mid["beta2_power"]) - 0.999 ** 3) < 1e-6   # two applies so far
    out = _run("--steps", "6", "--model-dir", b, "--resume")
    assert "restored model.ckpt-5" in out and "saved model.ckpt-11" in out
    x, y = ck.read_bundle(os.path.join(a, "model.ckpt-11")), ck.read_bundle(os.path.join(b, "model.ckpt-11"))
    assert set(x) == set(y)
    for k in x:
        assert np.array_equal(x[k], y[k]), k                          # bit-identical to the uninterrupted run
    # and it learns: Adam at its default 1e-3, N = 3
    out = _run("--steps", "3000")
    losses = [float(v) for v in re.findall(r"loss ([0-9.]+)", out)]
    assert len(losses) >= 3 and losses[-1] < 0.5 * losses[0], out
