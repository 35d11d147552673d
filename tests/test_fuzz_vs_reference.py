"""Differential fuzz of the host logic against the LIVE reference (authoring container only; skipped elsewhere).
A fixed seed of tools/fuzz_vs_reference.py: random shapes, ridge values, centring flags, view weights, confounds,
feature groups and dtypes must give the reference's scores, weights, means and pairwise correlations wherever the
problem is well posed and the component is determined."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.reference
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fixed_seed_fuzz_has_no_mismatch():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_vs_reference.py"), "20240924", "200"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "0 mismatches" in out.stdout


def test_fixed_seed_loss_fuzz_has_no_mismatch():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_loss_vs_reference.py"), "20240924", "200"],
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert "0 mismatches" in out.stdout
