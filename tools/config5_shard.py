"""BASELINE.json configs[4] (GCCA 8 x 2048, k=128, float64, n=5e5 over 8 GPUs) in units of ONE RANK'S ROW SHARD
(n=62500 rows per GPU): device-resident fit time (CUDA events, max over ranks), the K1 (fp64 DMMA) share and the
property residuals of tests/test_config5_gpu.py.  Prints one JSON line on rank 0 (committed under profiles/).

    python tools/config5_shard.py                                   # one shard on one GPU
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/config5_shard.py   # N shards, one all-reduce
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cca_zoo_b200 import _lib, ops  # noqa: E402
from cca_zoo_b200.linear import GCCA  # noqa: E402
from tests.test_config5_gpu import D1, K, M, N, check_gcca_properties, make_shard  # noqa: E402

world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:   # every rank holds its own 62500-row shard; fit() all-reduces the 2.1 GB moment buffer once
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))


def timed(fn, reps):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / reps], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)      # max over ranks, timed on the device
    return float(ms.item())


def main():
    views = make_shard(seed=rank)
    est = GCCA(latent_dimensions=K)
    lib = _lib.load()
    t_warm = timed(lambda: est.fit(views), 1)
    l0 = lib.ccab_launch_count()
    t_fit = timed(lambda: est.fit(views), 2)
    launches = (lib.ccab_launch_count() - l0) // 2
    t_k1 = timed(lambda: ops.moments(views), 2)
    res = check_gcca_properties(views, est.weights_, allreduce=(dist.all_reduce if world > 1 else None))
    D = M * D1
    flops = N * D * (D + 1)
    if rank == 0:
        print(json.dumps({
            "workload": f"GCCA.fit {M} views d={D1} k={K} float64, {N} rows per GPU (the rank shard of n=5e5 on 8 GPUs), "
                        f"device-resident, {world} GPU(s), total rows {N * world}",
            "n_gpus": world, "fit_ms": t_fit, "first_fit_ms": t_warm, "k1_ms": t_k1,
            "k1_tflops_fp64_per_gpu": flops / (t_k1 * 1e-3) / 1e12, "k1_share": t_k1 / t_fit,
            "after_k1_ms": t_fit - t_k1, "gpu_launches_per_fit": int(launches),
            "allreduce_bytes": (D * D + D + 1) * 8 if world > 1 else 0,
            "properties": {k: (v if not hasattr(v, "tolist") else [float(v[0]), float(v[-1])])
                           for k, v in res.items() if k != "C"},
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
