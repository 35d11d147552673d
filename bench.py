#!/usr/bin/env python
"""Benchmark of the hot path: rCCA.fit() on BASELINE.json configs[1]
(2-view rCCA, n=100000 rows per GPU, d=[1024,1024], k=64, c=0.1, float32 inputs).

    python bench.py --gpus 1 --steps 10 --warmup 3            # our CUDA path
    python bench.py --impl reference --steps 1 --warmup 0     # reference algorithm on the host cores
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   # sample-sharded, one all-reduce

One "step" = one fit.  Under N ranks every rank holds its own 100000-row shard (weak scaling): the job
is ONE fit over N*100000 samples per step, its throughput is reported in units of the 1-GPU workload
(`value` = N fit-units / s; at N=1 this is plain fit()/s).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS, DIMS, K, C_RIDGE = 100_000, [1024, 1024], 64, 0.1
SNR = 2.0 / 1024
WORKLOAD = "rCCA.fit 2 views n=100000 rows/GPU d=[1024,1024] k=64 c=0.1 float32 (JointData snr=2/1024)"
MODEL = "rcca"


def select_workload(name: str) -> None:
    """configs[1] (default) or configs[3]: MCCA 4 views d=512 k=32, 125000 rows per GPU (n=1e6 on 8 GPUs)."""
    global N_ROWS, DIMS, K, C_RIDGE, SNR, WORKLOAD, MODEL
    if name == "mcca4":
        N_ROWS, DIMS, K, C_RIDGE, SNR, MODEL = 125_000, [512] * 4, 32, 0.0, 2.0 / 512, "mcca"
        WORKLOAD = "MCCA.fit 4 views n=125000 rows/GPU d=[512]*4 k=32 c=0 float32 (JointData snr=2/512)"


def make_views(seed: int, n_rows: int = N_ROWS):
    """Rows of ONE JointData-style population (cca_zoo/datasets/_simulated.py:113-125): the loading matrices
    W_i come from a fixed stream shared by every rank, the latent draws and the noise from `seed`, so that
    the row shards of different ranks are samples of the same model (a sharded data set, not N unrelated ones)."""
    rng_w = np.random.default_rng(20240924)
    weights = [rng_w.standard_normal((p, K)) for p in DIMS]
    rng = np.random.default_rng(seed)
    z = rng.standard_normal((n_rows, K))
    views = []
    for w in weights:
        noise = rng.standard_normal((n_rows, w.shape[0])) * (1.0 / np.sqrt(SNR))
        views.append((z @ w.T + noise).astype(np.float32))
    return views


# ----------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled every ~10 ms DURING the timed region (NVML; nvidia-smi fallback)."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index = index
        self.sm, self.mx, self.reasons, self.power = [], [], set(), []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        nv = self._nvml
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM)))
        try:
            self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
        except Exception:
            pass
        try:
            mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        except Exception:
            mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        for bit, name in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
        if out.returncode == 0 and out.stdout.strip():
            r = [x.strip() for x in out.stdout.strip().split(",")]
            self.sm.append(float(r[0]))
            self.mx.append(float(r[1]))
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[2:]):
                if val == "Active":
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml() if self._nvml else self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.01 if self._nvml else 0.2)

    def start(self):
        self._t.start()
        return self

    def stop(self):
        try:  # one sample taken by the caller's thread at the end of the timed region (the Python-side loop
            self._sample_nvml() if self._nvml else self._sample_smi()   # can starve the sampler thread)
        except Exception:
            pass
        self._stop.set()
        self._t.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None,
                "sm_max_mhz": max(self.mx) if self.mx else None, "reasons": sorted(self.reasons),
                "samples": len(self.sm), "power_w_max": max(self.power) if self.power else None,
                "source": "nvml" if self._nvml else "nvidia-smi"}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference algorithm (oracle port: same LAPACK calls as cca_zoo/linear/_rcca.py:83-101)
# ----------------------------------------------------------------------------------------------
def cpu_fit_seconds(views):
    from oracle import restatement as R

    t0 = time.perf_counter()
    R.ref_rcca_fit(views, K, C_RIDGE)
    return time.perf_counter() - t0


def cpu_threads():
    try:
        from threadpoolctl import threadpool_info

        return max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(budget_s: float = 20.0):
    """Reference algorithm on a bounded row sample; tall-SVD cost is linear in n, so the full-size
    time is extrapolated as t_sample * (N_ROWS / n_sample)."""
    views = make_views(0, 12_500)
    t_cal = cpu_fit_seconds([v[:4000] for v in views])          # calibration (also warms BLAS)
    n_s = int(min(12_500, max(4000, 4000 * budget_s / max(t_cal, 1e-3))))
    t = cpu_fit_seconds([v[:n_s] for v in views])
    full = t * (N_ROWS / n_s)
    return {"value": 1.0 / full, "unit": "fit/s", "cores": cpu_threads(), "kind": "port",
            "sample": f"oracle.ref_rcca_fit (numpy LAPACK gesdd path of _rcca.py) on {n_s} of {N_ROWS} rows, "
                      f"{t:.2f} s measured, x{N_ROWS / n_s:.1f} linear-in-n extrapolation"}


def parity_vs_oracle(est, views):
    """Outside the timed region: the fitted weights / canonical correlations against the float64 oracle
    (covariance form of cca_zoo/linear/_rcca.py:83-101, numpy LAPACK) on the very same float32 inputs."""
    from oracle import restatement as R

    X = np.hstack(views).astype(np.float64)
    n = X.shape[0]
    mu = X.mean(axis=0)
    X -= mu
    C = X.T @ X / (n - 1)
    del X
    w_ref, sv = R.cov_rcca_fit(C, DIMS, K, C_RIDGE, n)
    w = [x.astype(np.float64) for x in est.weights_]
    ws = R.align_signs(w, w_ref)
    per_vec = np.concatenate([np.linalg.norm(a - b, axis=0) / np.linalg.norm(b, axis=0) for a, b in zip(ws, w_ref)])
    sub = np.random.default_rng(0).choice(n, 20_000, replace=False)
    vs = [v[sub] for v in views]
    sc = est.score(vs)
    sc_ref = R.score(vs, [mu[:DIMS[0]], mu[DIMS[0]:]], w_ref)
    return {"oracle": "oracle.restatement.cov_rcca_fit, float64, same inputs",
            "max_weight_rel_err": float(per_vec.max()), "median_weight_rel_err": float(np.median(per_vec)),
            "canonical_corr_max_rel_err": float(np.max(np.abs(sc - sc_ref) / np.abs(sc_ref))),
            "subspace_distance": float(max(R.subspace_distance(w[i], w_ref[i]) for i in range(2))),
            "min_gap_of_reference_spectrum": float(np.min(-np.diff(sv))), "tolerance_float32": 1e-3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    views = make_views(0, 12_500)
    t_cal = cpu_fit_seconds([v[:3000] for v in views])
    total_steps = max(args.steps + args.warmup, 1)
    n_s = int(min(12_500, max(3000, 3000 * (150.0 / total_steps) / max(t_cal, 1e-3))))
    sample = [v[:n_s] for v in views]
    for _ in range(args.warmup):
        cpu_fit_seconds(sample)
    times = [cpu_fit_seconds(sample) for _ in range(max(args.steps, 1))]
    t = float(np.mean(times)) * (N_ROWS / n_s)
    val = args.gpus / t  # same unit as our arm: 100000-row fit-units per second (CPU does them serially)
    val = 1.0 / t
    line = {
        "impl": "reference", "metric": "rcca_fit_per_s", "value": val, "unit": "fit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": val, "unit": "fit/s", "cores": cpu_threads(), "kind": "port",
                         "sample": f"{n_s} of {N_ROWS} rows per step, linear-in-n extrapolation x{N_ROWS / n_s:.1f}"},
        "e2e": {"value": val, "unit": "fit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from cca_zoo_b200 import _lib
    from cca_zoo_b200.linear import rCCA

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: cca_zoo_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    dev = torch.device("cuda", local)

    host = [torch.from_numpy(v).pin_memory() for v in make_views(1000 + rank)]
    views = [h.to(dev) for h in host]
    if MODEL == "mcca":
        from cca_zoo_b200.linear import MCCA

        est = MCCA(latent_dimensions=K, c=C_RIDGE, precision=args.precision)
    else:
        est = rCCA(latent_dimensions=K, c=C_RIDGE, precision=args.precision)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm (`value`) with live timing of the tcgen05 kernel ----
    for _ in range(args.warmup):
        est.fit(views)
    lib.ccab_profile_moments(1)
    k1_ms = []

    def step_dev():
        est.fit(views)
        k1_ms.append(lib.ccab_profile_moments_last_ms())

    sampler = ClockSampler(local).start() if rank == 0 else None
    l0 = lib.ccab_launch_count()
    total_ms = timed(step_dev, args.steps)
    launches = lib.ccab_launch_count() - l0
    clocks = sampler.stop() if sampler else None
    lib.ccab_profile_moments(0)
    ms_per_step = total_ms / args.steps
    value = world / (ms_per_step * 1e-3)

    # ---- end-to-end arm: pinned host inputs -> fit -> numpy weights ----
    def step_e2e():
        est.fit(host)

    for _ in range(min(args.warmup, 2)):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps) / args.steps
    h2d = sum(h.numel() * h.element_size() for h in host)
    d2h = sum(w.nbytes for w in est.weights_) + sum(m.nbytes for m in est.means_)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant data-parallel kernel (K1, moments_tf32_kernel) ----
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    bf16 = peaks.get("bf16_tflops", 1590.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops/2 (TF32 runs at half the dense bf16 rate)" if peaks else \
        "fallback 1590/2 TFLOP/s (B200_PROFILING.md)"
    D = sum(DIMS)
    flops = N_ROWS * D * (D + 1)  # algorithmic: symmetric product, SURVEY.md §8d (per rank)
    k1 = float(np.mean([m for m in k1_ms if m and m > 0])) if any(m and m > 0 for m in k1_ms) else None
    passes = 3 if args.precision == "tf32x3" else 1
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_k1_traffic.json")) as f:
            tr = json.load(f)[args.precision]
        traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]   # one ncu --set full capture, per launch
    except Exception:
        pass
    roof = None
    if k1:
        ach = flops / (k1 * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "moments_tf32_2cta_kernel", "achieved": ach, "peak": bf16 / 2,
                "unit": "TFLOP/s", "frac": ach / (bf16 / 2), "traffic": traffic, "kernel_ms": k1, "mma_passes": passes,
                "algorithmic_flops": flops, "algorithmic_bytes": N_ROWS * D * 4,
                "frac_of_issued": passes * ach / (bf16 / 2), "peak_source": peak_src,
                "share_of_step": k1 / ms_per_step}

    line = {
        "metric": f"{MODEL}_fit_per_s", "value": value, "unit": "fit/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "tf32x3+f32" if args.precision == "tf32x3" else args.precision + "+f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_gpu": N_ROWS, "total_rows": N_ROWS * world,
                   "parallelism": f"sample-sharded x{world}, one all-reduce of the moment buffer" if world > 1 else "single GPU",
                   "l2": "inputs (819 MB per GPU) exceed the 126 MB L2; no explicit flush",
                   "unit_note": "value = (n_gpus x 100000-row fit-units) / step time"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": world / (e2e_ms * 1e-3), "unit": "fit/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
        "roofline": roof,
    }
    if world == 1 and not args.no_cpu and MODEL == "rcca":
        line["cpu_baseline"] = cpu_baseline()
        line["parity"] = parity_vs_oracle(est, [h.numpy() for h in host])
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tf32x3", choices=["tf32", "tf32x3", "exact"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="rcca", choices=["rcca", "mcca4"],
                    help="rcca = BASELINE configs[1] (the headline); mcca4 = configs[3] shard (scaling study)")
    args = ap.parse_args()
    select_workload(args.workload)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
