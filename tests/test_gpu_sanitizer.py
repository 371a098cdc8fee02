"""-m gpu: compute-sanitizer over one small forward of every product kernel (GSC and TSM): synccheck (mbarrier / named
barrier protocol of the warp-specialised kernels), memcheck and racecheck (shared-memory hazards: the row staging of the
ShareLayer reduce, the exchange buffers of the epilogues).  Skipped only when the tool is not installed."""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(tool, variant):
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    env = dict(os.environ)
    env.pop("BSR_DEBUG_KEEP", None)
    cmd = [exe, "--tool", tool, "--print-limit", "5", sys.executable, os.path.join(ROOT, "tools", "profile_forward.py"), "2", variant]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    out = r.stdout + r.stderr
    # the FIRST report says which kernel / barrier; the tail is only the host backtrace of the last one
    assert "launches per forward" in out, out[:3000] + "\n...\n" + out[-1500:]
    return out


@pytest.mark.parametrize("variant", ["gsc", "tsm"])
def test_synccheck_clean(variant):
    out = _run("synccheck", variant)
    assert "ERROR SUMMARY: 0 errors" in out, out[-3000:]


def test_memcheck_clean():
    out = _run("memcheck", "tsm")
    assert "ERROR SUMMARY: 0 errors" in out, out[-3000:]


def test_racecheck_clean():
    out = _run("racecheck", "tsm")
    assert "RACECHECK SUMMARY: 0 hazards displayed (0 errors, 0 warnings)" in out, out[:3000]


def test_memcheck_glue_pass_clean():
    """The kernels outside the generator's launches - chunk split through shared memory, caller glue, composite, the UCB
    post-processing, the compact host path - under memcheck (tools/profile_glue.py, 2 images)."""
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    env = dict(os.environ)
    env.pop("BSR_DEBUG_KEEP", None)
    cmd = [exe, "--tool", "memcheck", "--print-limit", "5", sys.executable, os.path.join(ROOT, "tools", "profile_glue.py"), "2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    out = r.stdout + r.stderr
    assert "glue pass done" in out, out[:3000] + "\n...\n" + out[-1500:]
    assert "ERROR SUMMARY: 0 errors" in out, out[-3000:]
