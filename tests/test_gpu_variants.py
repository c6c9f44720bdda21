"""The non-default launch structures must reproduce the same golden vectors as the default (one kernel per fixed-grid solve).
The switches are read when the library is loaded, so each variant runs in a fresh interpreter."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
    "one_launch_per_evaluation": {"SCLDM_SOLVE": "0"},
    "no_pdl_modulation_per_evaluation": {"SCLDM_SOLVE": "0", "SCLDM_PDL": "0", "SCLDM_MOD_BATCH": "0"},
}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(VARIANTS))
def test_kernel_variant_matches_golden(name):
    env = dict(os.environ, **VARIANTS[name])
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_dit.py"), "-m", "gpu", "-q", "-x",
                          "-p", "no:cacheprovider", "-k", "golden or intermediates or ode or large_batch or batch_invariance"],
                         env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert " passed" in res.stdout
