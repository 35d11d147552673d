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
