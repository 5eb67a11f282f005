"""BASELINE config #1: the reference's demo (demo/fit_identity.ipynb) end to end on the native layer --
RSLM initialisation, fused LM + AMIS forward, native backward, Adam.  The Monte-Carlo pose loss must fall and
the test pose error must improve."""
import os
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_fit_identity_trains(cuda_device):
    sys.path.insert(0, os.path.join(ROOT, "demo"))
    import fit_identity
    out = fit_identity.run(steps=160, batch_size=256, verbose=False)
    assert out["finite"]
    assert out["loss_mc_last"] < out["loss_mc_first"] - 0.2
    assert out["test_t_err_after"] < 0.7 * out["test_t_err_before"]
    assert out["test_r_err_after"] < 0.7 * out["test_r_err_before"]
    print(out)
