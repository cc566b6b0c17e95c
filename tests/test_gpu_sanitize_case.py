"""tests/sanitize_case.py -- the compact every-kernel workload written for compute-sanitizer (closed on this GPU pool:
profiles/r2_sanitizer.txt) -- as part of the suite: every result is compared with the CPU oracle; with a library built with
-DSVFM_DEBUG_CHECKS (SVFM_LIB_PATH) the kernels also assert their own bounds and protocol invariants."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sanitize_case_matches_the_oracle(oracle):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "sanitize_case.py")], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "sanitize_case: all results match the oracle" in out.stdout
