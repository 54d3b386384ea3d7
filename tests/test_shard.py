"""
CPU test of the multi-rank host logic (world_size 2, gloo): partition by
object index, per-rank fit on the shard, gather of the result records.  The
per-rank "fit" here is a deterministic stand-in producing ngx_lm_result-shaped
records from its inputs (the real fit needs a GPU; its sharding is exercised
by `bench.py --gpus N`).
"""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _fake_fit(images, weights, jacobians, psf_pars, guesses):
    from ngmix_b200 import _lib
    n = images.shape[0]
    res = np.zeros(n, dtype=_lib.LM_RESULT_DTYPE)
    res["pars"][:, :6] = guesses + images.reshape(n, -1).sum(axis=1)[:, None]
    res["npix"] = (weights.reshape(n, -1) > 0).sum(axis=1)
    res["flux"] = jacobians[:, 0]
    return res


def _worker(rank, world, port, n, outdir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from ngmix_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(0)          # same global batch on all ranks
    arrays = dict(images=rng.normal(size=(n, 5, 5)),
                  weights=(rng.uniform(size=(n, 5, 5)) > 0.2) * 1.0,
                  jacobians=rng.normal(size=(n, 8)), psf_pars=None,
                  guesses=rng.normal(size=(n, 6)))
    lo, hi = shard.shard_range(n, rank, world)
    out = shard.run_sharded(_fake_fit, arrays)
    full = _fake_fit(**{k: v for k, v in arrays.items()})
    assert out.shape == (n,)
    assert np.array_equal(out["pars"], full["pars"])
    assert np.array_equal(out["npix"], full["npix"])
    local = shard.run_sharded(_fake_fit, arrays, gather=False)
    assert np.array_equal(local["flux"], full["flux"][lo:hi])
    np.save(os.path.join(outdir, "r%d.npy" % rank), out["flux"])
    dist.destroy_process_group()


def test_shard_ranges():
    from ngmix_b200.shard import shard_range
    for n in (0, 1, 7, 10000, 1000003):
        for world in (1, 2, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans[:-1], spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo(tmp_path):
    n, world, port = 37, 2, 29613
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world,
             join=True)
    a = np.load(tmp_path / "r0.npy")
    b = np.load(tmp_path / "r1.npy")
    assert np.array_equal(a, b) and a.size == n
