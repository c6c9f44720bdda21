"""The per-phase and unfused kernel variants must reproduce the same golden vectors as the default persistent
block-stack kernel.  The variants are selected by environment switches that are read when the library / the packed
weights are created, so each one runs in a fresh interpreter."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

VARIANTS = {
    "first_generation_block_stack_kernel": {"SCLDM_MEGA": "1"},
    "one_kernel_per_block_half": {"SCLDM_MEGA": "0"},
    "unfused_no_pdl": {"SCLDM_MEGA": "0", "SCLDM_FUSED_ATTN": "0", "SCLDM_FUSED_MLP": "0", "SCLDM_TC_FINAL": "0", "SCLDM_PDL": "0"},
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
