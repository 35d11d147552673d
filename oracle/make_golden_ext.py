"""Generate tests/golden/reference_outputs_ext.{npz,json} from the UNMODIFIED reference: the estimators
that CALL the MCCA core with extra fit arguments (PartialCCA: ``partials``; GRCCA: ``feature_groups``).

    python oracle/make_golden_ext.py

TEST INFRASTRUCTURE ONLY (see make_golden.py).  Kept in a second fixture file so that the round-1 vectors in
reference_outputs.npz stay byte-identical.
"""
from __future__ import annotations

import json
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import refshim  # noqa: E402

refshim.install()

from cca_zoo.linear import GCCA, GRCCA, MCCA, PartialCCA, rCCA  # noqa: E402

import torch  # noqa: E402
from cca_zoo.deep.objectives import GCCALoss  # noqa: E402

from oracle.make_golden import DATASETS, build_dataset, loss_inputs  # noqa: E402


# extra input recipes (rebuilt by tests/golden_io.py): a base data set of make_golden.DATASETS, then optionally
# "shift" (adds a constant, so that centring matters) and "dup" = [view, column] (appends a copy of that column)
EXT_DATASETS = {
    "three_views_shifted": ("derived", {"base": "three_views", "shift": 1.5}),
    "joint4_shifted": ("derived", {"base": "joint4_gcca", "shift": -0.8}),
    "two_views_dup": ("derived", {"base": "two_views", "dup": [1, 0]}),
}


def build_ext_dataset(name):
    if name not in EXT_DATASETS:
        return build_dataset(name)
    args = EXT_DATASETS[name][1]
    views = [v.copy() for v in build_dataset(args["base"])]
    if "shift" in args:
        views = [v + args["shift"] for v in views]
    if "dup" in args:
        j, col = args["dup"]
        views[j] = np.hstack([views[j], views[j][:, col:col + 1]])
    return views


# estimators whose ``center=False`` semantics differ from "skip the centring everywhere" (np.cov centres inside
# MCCA / GCCA / GRCCA), and the ridge-regularised fit of a rank-deficient view (nothing is dropped when c > 0)
CENTER_CASES = [
    # (name, model, kwargs, dataset, dtype)
    ("mcca_nocenter", "MCCA", dict(latent_dimensions=2, c=0.1, center=False, pca=False), "three_views_shifted", "f64"),
    ("mcca_nocenter_pca", "MCCA", dict(latent_dimensions=3, center=False), "three_views_shifted", "f64"),
    ("gcca_nocenter", "GCCA", dict(latent_dimensions=2, c=0.1, center=False), "three_views_shifted", "f64"),
    ("gcca_nocenter_w", "GCCA", dict(latent_dimensions=3, center=False, view_weights=[1.0, 2.0, 0.5, 1.0]),
     "joint4_shifted", "f64"),
    ("gcca_nocenter32", "GCCA", dict(latent_dimensions=3, c=0.2, center=False), "joint4_shifted", "f32"),
    ("rcca_nocenter_shifted", "rCCA", dict(latent_dimensions=2, c=0.3, center=False), "two_views_dup", "f64"),
    ("rcca_dup_ridge", "rCCA", dict(latent_dimensions=9, c=0.2), "two_views_dup", "f64"),
]
CENTER_MODELS = {"MCCA": MCCA, "GCCA": GCCA, "rCCA": rCCA}


def confounds(n, q, seed):
    """Seeded confounds with non-zero means (so that centring of the views vs no centring of P matters)."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, q)) + np.linspace(0.3, 1.2, q)


def groups(dims, n_groups, seed):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, g, size=d) for d, g in zip(dims, n_groups)]


PARTIAL_CASES = [
    # (name, kwargs, dataset, dtype, q, seed)
    ("pcca_two", dict(latent_dimensions=2), "two_views", "f64", 3, 1),
    ("pcca_two_c", dict(latent_dimensions=2, c=[0.1, 0.3]), "two_views", "f64", 1, 2),
    ("pcca_two_nocenter", dict(latent_dimensions=2, c=0.05, center=False), "two_views", "f64", 2, 3),
    ("pcca_three", dict(latent_dimensions=2, c=0.2), "three_views", "f64", 3, 4),
    ("pcca_med", dict(latent_dimensions=5, c=0.05), "joint3_med", "f64", 4, 5),
    ("pcca_med32", dict(latent_dimensions=5, c=0.05), "joint3_med", "f32", 4, 5),
]

GROUP_CASES = [
    # (name, kwargs, dataset, dtype, groups per view, seed)
    ("grcca_two", dict(latent_dimensions=2, c=0.5), "two_views", "f64", [3, 3], 2),
    ("grcca_two_pv", dict(latent_dimensions=1, c=[0.5, 0.0]), "two_views", "f64", [3, 3], 2),
    ("grcca_three_mu", dict(latent_dimensions=2, c=[0.3, 0.6, 0.2], mu=[0.5, 2.0, 1.0]), "three_views", "f64",
     [2, 2, 3], 3),
    ("grcca_c0", dict(latent_dimensions=2, c=0.0), "two_views", "f64", [3, 3], 2),
    ("grcca_med", dict(latent_dimensions=5, c=0.2, mu=0.5), "joint3_med", "f64", [8, 6, 5], 4),
    ("grcca_med32", dict(latent_dimensions=5, c=0.2, mu=0.5), "joint3_med", "f32", [8, 6, 5], 4),
    ("grcca_nocenter", dict(latent_dimensions=2, c=0.4, mu=2.0, center=False), "three_views_shifted", "f64", [3, 2, 2], 5),
]


GLOSS_CASES = [
    # (name, batch, widths, eps, seed) -- inputs: make_golden.loss_inputs (shared latent + noise)
    ("gloss_32x4x3", 32, [4, 4, 4], 1e-4, 10),
    ("gloss_256x3", 256, [8, 6, 10], 1e-5, 11),
    ("gloss_12x5x3", 12, [5, 5, 5], 1e-4, 12),          # fewer samples than total width
    ("gloss_1024x32x4", 1024, [32, 32, 24, 16], 1e-5, 13),
]


def main():
    out, meta = {}, {"datasets": {**DATASETS, **EXT_DATASETS}, "partial_cases": [], "group_cases": [],
                     "gloss_cases": [], "center_cases": []}
    for name, model, kwargs, ds, dt in CENTER_CASES:
        views = build_ext_dataset(ds)
        if dt == "f32":
            views = [v.astype(np.float32) for v in views]
        est = CENTER_MODELS[model](**kwargs).fit(views)
        for i, (w, mu) in enumerate(zip(est.weights_, est.means_)):
            out[f"{name}/w{i}"], out[f"{name}/mean{i}"] = np.asarray(w), np.asarray(mu)
        out[f"{name}/score"] = np.asarray(est.score(views))
        meta["center_cases"].append(dict(name=name, model=model, kwargs=kwargs, dataset=ds, dtype=dt))
        print(name, out[f"{name}/score"])
    for name, batch, widths, eps, seed in GLOSS_CASES:
        zs = [z.clone().requires_grad_(True) for z in loss_inputs(batch, widths, seed)]
        loss = GCCALoss(eps=eps)(zs)
        loss.backward()
        out[f"{name}/loss"] = np.array(loss.item())
        for i, z in enumerate(zs):
            out[f"{name}/grad{i}"] = z.grad.numpy()
        meta["gloss_cases"].append(dict(name=name, batch=batch, widths=widths, eps=eps, seed=seed))
        print(name, loss.item())
    for name, kwargs, ds, dt, q, seed in PARTIAL_CASES:
        views = build_dataset(ds)
        if dt == "f32":
            views = [v.astype(np.float32) for v in views]
        Z = confounds(views[0].shape[0], q, seed)
        est = PartialCCA(**kwargs).fit(views, partials=Z)
        for i, (w, mu, b) in enumerate(zip(est.weights_, est.means_, est.confound_betas_)):
            out[f"{name}/w{i}"], out[f"{name}/mean{i}"], out[f"{name}/beta{i}"] = np.asarray(w), np.asarray(mu), b
        out[f"{name}/score"] = np.asarray(est.score(views))
        zs = est.transform(views, partials=Z)
        out[f"{name}/partial_corr"] = np.array([abs(np.corrcoef(zs[0][:, d], zs[1][:, d])[0, 1])
                                                for d in range(zs[0].shape[1])])
        meta["partial_cases"].append(dict(name=name, kwargs=kwargs, dataset=ds, dtype=dt, q=q, seed=seed))
        print(name, out[f"{name}/partial_corr"])
    for name, kwargs, ds, dt, ng, seed in GROUP_CASES:
        views = build_ext_dataset(ds)
        if dt == "f32":
            views = [v.astype(np.float32) for v in views]
        gs = groups([v.shape[1] for v in views], ng, seed)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            est = GRCCA(**kwargs).fit(views, feature_groups=gs)
        for i, (w, mu) in enumerate(zip(est.weights_, est.means_)):
            out[f"{name}/w{i}"], out[f"{name}/mean{i}"] = np.asarray(w), np.asarray(mu)
        out[f"{name}/score"] = np.asarray(est.score(views))
        meta["group_cases"].append(dict(name=name, kwargs=kwargs, dataset=ds, dtype=dt, n_groups=ng, seed=seed))
        print(name, out[f"{name}/score"])
    gdir = os.path.join(ROOT, "tests", "golden")
    np.savez_compressed(os.path.join(gdir, "reference_outputs_ext.npz"), **out)
    with open(os.path.join(gdir, "reference_outputs_ext.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
