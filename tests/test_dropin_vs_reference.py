"""Drop-in checks that need the LIVE reference (authoring container only): the reference's own model-selection
wrapper drives this package's estimators (clone / get_params / set_params / fit / score contract), host logic on the
torch-CPU stand-in."""
import subprocess
import sys
import os

import pytest

pytestmark = pytest.mark.reference
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys, warnings
sys.path.insert(0, %r)
from tests import fake_ops
from oracle import refshim
refshim.install()
import numpy as np, pytest
fake_ops.install(pytest.MonkeyPatch())
from cca_zoo.model_selection import GridSearchCV
import cca_zoo.linear as ref
from cca_zoo_b200 import linear as ours
rng = np.random.default_rng(0)
lat = rng.standard_normal((150, 2))
views = [lat @ rng.standard_normal((2, 8)) + rng.standard_normal((150, 8)),
         lat @ rng.standard_normal((2, 6)) + rng.standard_normal((150, 6)),
         lat @ rng.standard_normal((2, 5)) + rng.standard_normal((150, 5))]
for name, nv, grid in (("rCCA", 2, {"c": [0.0, 0.1, 0.5, 0.9]}), ("MCCA", 3, {"c": [0.0, 0.3], "eps": [1e-6, 1e-3]}),
                       ("GCCA", 3, {"c": [0.1, 0.6]})):
    res = []
    for lib in (ref, ours):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            gs = GridSearchCV(getattr(lib, name)(latent_dimensions=2), param_grid=grid, cv=3).fit(views[:nv])
        res.append((gs.best_params_, gs.best_score_, gs.cv_results_["mean_test_score"]))
    assert res[0][0] == res[1][0], (name, res[0][0], res[1][0])
    assert np.allclose(res[0][2], res[1][2], atol=1e-8), (name, res[0][2], res[1][2])
    assert type(gs.best_estimator_).__module__.startswith("cca_zoo_b200")
print("DROPIN_OK")
"""


def test_reference_gridsearch_drives_our_estimators():
    out = subprocess.run([sys.executable, "-c", SCRIPT % ROOT], capture_output=True, text=True, timeout=900)
    assert "DROPIN_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
