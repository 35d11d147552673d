"""Sample-sharded fit on 2 GPUs (NCCL): every rank fits its row shard, the moments are all-reduced once,
the replicated solve must give the same weights on every rank and match a single-GPU fit of the whole
data.  Skipped on boxes with fewer than 2 GPUs (the CPU/gloo logic test is tests/test_parallel_cpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from cca_zoo_b200 import parallel
    from cca_zoo_b200.datasets import joint_data
    from cca_zoo_b200.linear import MCCA, rCCA

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        views = joint_data(n_views=3, n_samples=6001, n_features=[160, 96, 40], latent_dimensions=5,
                           signal_to_noise=0.05, random_state=9)
        lo, hi = parallel.shard_rows(6001, rank, world)
        shard = [v[lo:hi] for v in views]
        est = rCCA(latent_dimensions=5, c=0.1).fit(shard[:2])
        assert est.n_samples_ == 6001
        np.save(os.path.join(out_dir, f"rcca_w0_rank{rank}.npy"), est.weights_[0])
        np.save(os.path.join(out_dir, f"rcca_mean0_rank{rank}.npy"), est.means_[0])
        m = MCCA(latent_dimensions=5, c=0.05).fit(shard)
        np.save(os.path.join(out_dir, f"mcca_w2_rank{rank}.npy"), m.weights_[2])
        # wide views: the device-side fit, which reads the all-reduced sample count on the device (no read-back)
        wide = joint_data(n_views=2, n_samples=5003, n_features=[256, 320], latent_dimensions=4,
                          signal_to_noise=0.02, random_state=10, dtype=np.float32)
        lo, hi = parallel.shard_rows(5003, rank, world)
        dev = rCCA(latent_dimensions=4, c=0.1).fit([torch.from_numpy(v[lo:hi]).cuda() for v in wide])
        assert dev._fit_info["route"] == "device" and dev.n_samples_ == 5003
        np.save(os.path.join(out_dir, f"wide_w1_rank{rank}.npy"), dev.weights_[1])
        # which exchange ran?  (fused NVLS kernel on symmetric memory when the box offers multicast, NCCL otherwise)
        used_nvls = any(v is not None for v in parallel._NvlsExchange._cache.values())
        os.environ["CCAB_EXCHANGE"] = "nccl"
        ref = rCCA(latent_dimensions=4, c=0.1).fit([torch.from_numpy(v[lo:hi]).cuda() for v in wide])
        os.environ.pop("CCAB_EXCHANGE")
        np.save(os.path.join(out_dir, f"wide_w1_nccl_rank{rank}.npy"), ref.weights_[1])
        with open(os.path.join(out_dir, f"exchange_rank{rank}.txt"), "w") as f:
            f.write("nvls" if used_nvls else "nccl")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_fit_matches_single_gpu(tmp_path):
    import torch.multiprocessing as mp

    from cca_zoo_b200.datasets import joint_data
    from cca_zoo_b200.linear import MCCA, rCCA
    from oracle import restatement as R

    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    views = joint_data(n_views=3, n_samples=6001, n_features=[160, 96, 40], latent_dimensions=5,
                       signal_to_noise=0.05, random_state=9)
    single = rCCA(latent_dimensions=5, c=0.1).fit(views[:2])
    w0 = [np.load(tmp_path / f"rcca_w0_rank{r}.npy") for r in range(2)]
    assert np.array_equal(w0[0], w0[1]), "replicated solve must be bit-identical across ranks"
    assert R.max_rel_err_per_vector([w0[0]], [single.weights_[0]]) < 1e-9
    np.testing.assert_allclose(np.load(tmp_path / "rcca_mean0_rank0.npy"), single.means_[0], rtol=1e-12, atol=1e-12)
    ms = MCCA(latent_dimensions=5, c=0.05).fit(views)
    w2 = np.load(tmp_path / "mcca_w2_rank1.npy")
    assert R.max_rel_err_per_vector([w2], [ms.weights_[2]]) < 1e-9
    wide = joint_data(n_views=2, n_samples=5003, n_features=[256, 320], latent_dimensions=4,
                      signal_to_noise=0.02, random_state=10, dtype=np.float32)
    one = rCCA(latent_dimensions=4, c=0.1).fit(wide)
    ww = [np.load(tmp_path / f"wide_w1_rank{r}.npy") for r in range(2)]
    assert np.array_equal(ww[0], ww[1])
    assert R.max_rel_err_per_vector([ww[0].astype(np.float64)], [one.weights_[1].astype(np.float64)]) < 2e-4
    # the fused NVLS exchange and the NCCL all-reduce sum the same two shards: identical totals, identical weights
    wn = np.load(tmp_path / "wide_w1_nccl_rank0.npy")
    print("exchange path:", open(tmp_path / "exchange_rank0.txt").read())
    assert R.max_rel_err_per_vector([ww[0].astype(np.float64)], [wn.astype(np.float64)]) < 1e-6


def test_exchange_message_round_trip():
    """pack -> unpack restores the moment buffer (upper block triangle, column sums) and carries n."""
    from cca_zoo_b200 import ops

    views = [torch.randn(700, d, device="cuda", dtype=torch.float64) for d in (130, 64, 300)]
    dims = [130, 64, 300]
    mom = ops.moments(views)
    packed = ops.moments_pack(mom, dims, 700)
    assert packed.numel() < mom.numel() * 0.62
    back, n_dev = ops.moments_unpack(packed, dims)
    assert float(n_dev.item()) == 700.0
    assert torch.equal(back, mom)
