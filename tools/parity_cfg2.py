"""Parity at BASELINE config 2 scale: CUDA path (each covariance arithmetic, both solvers) vs the float64 oracle
(covariance form, numpy/LAPACK) on the SAME float32 JointData.  Prints per-vector weight error, score error,
subspace distance and the minimum gap between adjacent canonical correlations."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
from cca_zoo_b200.datasets import joint_data
from cca_zoo_b200.linear import rCCA
from oracle import restatement as R
n, d, k, c = 100000, 1024, 64, 0.1
out = []
for snr in [2.0 / 1024, 1.0]:
    views = joint_data(2, n, k, [d, d], snr, 0, np.float32)
    t0 = time.time()
    X = np.hstack(views).astype(np.float64)
    mu = X.mean(0); Xc = X - mu
    C = Xc.T @ Xc / (n - 1)
    w_ref, sv = R.cov_rcca_fit(C, [d, d], k, c, n)
    print(f"snr={snr:.4g}: oracle fp64 in {time.time()-t0:.1f}s; sv range [{sv.min():.4f},{sv.max():.4f}] min gap {np.min(-np.diff(sv)):.2e}", flush=True)
    sub = np.random.default_rng(0).choice(n, 20000, replace=False)
    vs = [v[sub] for v in views]
    sc_ref = R.score(vs, [mu[:d], mu[d:]], w_ref)
    dev = [torch.from_numpy(v).cuda() for v in views]
    for prec in ["tf32", "tf32x3", "exact"]:
        for solver in ["cholesky", "eigen"]:
            if solver == "eigen" and prec != "tf32x3":
                continue
            est = rCCA(latent_dimensions=k, c=c, precision=prec, solver=solver).fit(dev)
            w = [x.astype(np.float64) for x in est.weights_]
            pv = R.max_rel_err_per_vector(w, w_ref)
            ws = R.align_signs(w, w_ref)
            med = float(np.median(np.linalg.norm(ws[0] - w_ref[0], axis=0) / np.linalg.norm(w_ref[0], axis=0)))
            sd = max(R.subspace_distance(w[i], w_ref[i]) for i in range(2))
            sc = est.score(vs)
            se = float(np.max(np.abs(sc - sc_ref) / np.abs(sc_ref)))
            row = dict(snr=snr, precision=prec, solver=solver, max_vec_rel_err=pv, median_vec_rel_err=med,
                       subspace_dist=sd, score_rel_err=se)
            out.append(row); print(row, flush=True)
json.dump(out, open("gpurun_out/parity_cfg2.json", "w"), indent=1)
