"""--neg_deg_sample.  CPU: the oracle's restatement is pinned to the reference's fixtures by
tests/test_oracle_golden.py (negdeg_* cases).  GPU: tests/negdeg_check.py, run in its OWN process and allowed to fail --
the kge_negdeg.cu kernels were written after this round's GPU budget was spent and have not run on a device yet; a
device-side fault in them must not take the rest of the suite's CUDA context down with it."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="kge_negdeg.cu has not run on a GPU yet (written after the round's GPU budget was spent)")
def test_neg_deg_sample_matches_reference_fixtures_and_oracle():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "negdeg_check.py")], capture_output=True, text=True,
                         timeout=900, cwd=ROOT)
    assert "NEGDEG_CHECK_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]


def test_neg_deg_sample_step_configuration():
    """host side: the flag travels in the step configuration; chunk bookkeeping of the extra columns"""
    from dglke_b200 import _lib
    from dglke_b200.engine import Hyper
    import ctypes as C
    cfg = _lib.make_cfg("TransE_l2", 8, 8, 12.0, 0.1, 0.1, 0.0, 3, False, 1.0, False, 16, 8, 4, neg_deg_sample=True)
    assert cfg.neg_deg_sample == 1 and cfg.neg_sample_size == 4 and C.sizeof(_lib.StepCfg) == 80
    assert _lib.StepCfg.neg_deg_sample.offset == 76
    assert Hyper(neg_deg_sample=True).neg_deg_sample and not Hyper().neg_deg_sample
