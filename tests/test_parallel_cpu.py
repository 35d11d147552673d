"""Host-side logic of the sample-sharded path (N > 1) on CPU with the gloo backend, world_size 2.

The moment kernel itself needs a GPU; here the oracle's numpy moments stand in for it so that the
exchange step (packing, ONE all-reduce carrying M, s and n, row sharding) is exercised end to end:
the all-reduced buffer must equal the moments of the whole data set and give the same covariance.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cca_zoo_b200 import parallel
from cca_zoo_b200.datasets import joint_data
from oracle import restatement as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _padded_moment_buffer(views):
    """Oracle moments laid out like ccab_moments' output for widths that are multiples of 128-free:
    here simply [M.flatten(), s] (the layout is opaque to the all-reduce)."""
    M, s, n = R.moments(views)
    return torch.from_numpy(np.concatenate([M.ravel(), s])), n


def _worker(rank, world, port, n_rows, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        views = joint_data(n_views=2, n_samples=n_rows, n_features=[12, 9], latent_dimensions=3,
                           signal_to_noise=0.5, random_state=5)
        lo, hi = parallel.shard_rows(n_rows, rank, world)
        shard = [v[lo:hi] for v in views]
        buf, n_local = _padded_moment_buffer(shard)
        assert n_local == hi - lo
        tot, n_total = parallel.allreduce_moments(buf, n_local)
        assert n_total == n_rows
        np.save(os.path.join(out_dir, f"rank{rank}.npy"), tot.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [1000, 1001])
def test_allreduce_of_shard_moments_equals_whole(tmp_path, n_rows):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_rows, str(tmp_path)), nprocs=world, join=True)
    views = joint_data(n_views=2, n_samples=n_rows, n_features=[12, 9], latent_dimensions=3,
                       signal_to_noise=0.5, random_state=5)
    whole, _ = _padded_moment_buffer(views)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1), "every rank must end with bit-identical moments (replicated eigensolve)"
    np.testing.assert_allclose(r0, whole.numpy(), rtol=1e-12, atol=1e-9)
    D = 21
    M = r0[: D * D].reshape(D, D)
    s = r0[D * D:]
    C = R.covariance_from_moments(M, s, n_rows)
    Mw, sw, _ = R.moments(views)
    np.testing.assert_allclose(C, R.covariance_from_moments(Mw, sw, n_rows), rtol=1e-10, atol=1e-12)


def test_shard_rows_partitions_exactly():
    for n in (1, 7, 100, 100_000):
        for world in (1, 2, 3, 8):
            bounds = [parallel.shard_rows(n, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            for (a, b), (c, d) in zip(bounds[:-1], bounds[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in bounds]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_is_identity():
    buf = torch.arange(5, dtype=torch.float64)
    out, n = parallel.allreduce_moments(buf, 17)
    assert out is buf and n == 17
