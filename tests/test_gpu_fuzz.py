"""GPU test (-m gpu): >= 1000 corrupted streams through the real kernel, under a host watchdog (a corrupt stream must
end in a page status, never in a hang, and never write outside its buffers)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_corrupt_streams_on_the_gpu_under_a_watchdog():
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "gpu_fuzz.py"), "1200", "20261017"]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, cwd=ROOT)      # the watchdog
    except subprocess.TimeoutExpired:
        pytest.fail("the GPU fuzz run did not end: a corrupt stream hung the page kernel")
    assert r.returncode == 0 and "GPU FUZZ OK trials 1200" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
