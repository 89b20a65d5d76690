"""World-size-2 gloo test (CPU) of the batch-sharding host logic used by bench.py --gpus N."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG_NAME, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, global_batch, ret):
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module(PKG_NAME).sharding
    lo, hi = sh.shard_bounds(global_batch, rank, world)
    full = torch.arange(global_batch * 6, dtype=torch.float32).reshape(global_batch, 2, 3)
    local = full[lo:hi] * 2.0                       # stand-in for the per-sample forward
    got = sh.gather_shards(local, global_batch)
    ok = torch.equal(got, full * 2.0)
    mx = sh.max_over_ranks([1.0 + rank, 5.0 - rank])
    ret[rank] = (lo, hi, bool(ok), mx)
    dist.destroy_process_group()


def test_shards_cover_batch_and_timings_take_the_max():
    for global_batch in (32, 33, 1):
        world, port = 2, _free_port()
        with mp.Manager() as m:
            ret = m.dict()
            mp.spawn(_worker, args=(world, port, global_batch, ret), nprocs=world, join=True)
            ret = dict(ret)
        assert ret[0][0] == 0 and ret[0][1] == ret[1][0] and ret[1][1] == global_batch
        assert ret[0][2] and ret[1][2]
        assert ret[0][3] == ret[1][3] == [2.0, 5.0]


def _worker_grad(rank, world, port, ret):
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tt = importlib.import_module(PKG_NAME).trunk_train
    flat = torch.arange(10, dtype=torch.float32) * (rank + 1)          # this rank's flat gradient buffer
    tt.allreduce_flat_(flat)                                           # the train step's single collective
    ret[rank] = flat.tolist()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_sums_over_ranks():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker_grad, args=(world, port, ret), nprocs=world, join=True)
        ret = dict(ret)
    want = (torch.arange(10, dtype=torch.float32) * 3).tolist()       # 1x + 2x; the 1/world lives in the RMSprop kernel
    assert ret[0] == want and ret[1] == want


def _worker_two_buckets(rank, world, port, ret):
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ts = importlib.import_module(PKG_NAME).train_sun
    flat = torch.arange(12, dtype=torch.float32) * (rank + 1)          # [conv / norm gradients | Dense gradients], split at 5
    work = ts.start_tail_allreduce(flat, 5)                            # Dense part first, asynchronously (SunTrainer.sun_train_step)
    flat[:5] += 100.0 * (rank + 1)                                     # "the rest of the backward pass" still writes the head
    ts.finish_allreduce(flat, 5, work)
    ret[rank] = flat.tolist()
    dist.destroy_process_group()


def test_two_bucket_gradient_allreduce():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker_two_buckets, args=(world, port, ret), nprocs=world, join=True)
        ret = dict(ret)
    base = torch.arange(12, dtype=torch.float32) * 3
    base[:5] += 300.0
    assert ret[0] == base.tolist() and ret[1] == base.tolist()


def test_shard_bounds_edge_cases(pkg):
    sb = pkg.sharding.shard_bounds
    assert [sb(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert [sb(2, r, 4) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]      # empty shards
    assert sb(0, 0, 1) == (0, 0)


def _worker_ragged(rank, world, port, ret):
    """Ragged shards (3 samples over 2 ranks) and an empty shard (1 sample over 2 ranks): with the adjoints normalised by the GLOBAL
    batch, the two-bucket all-reduce protocol of train_sun / train returns the global-batch mean gradient on every rank."""
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = importlib.import_module(PKG_NAME)
    ts, sh = pkg.train_sun, pkg.sharding
    out = {}
    for gb in (3, 1):
        per_sample = torch.arange(gb * 6, dtype=torch.float32).reshape(gb, 6) + 1.0      # sample i's gradient of its own loss term
        lo, hi = sh.shard_bounds(gb, rank, world)
        flat_g = per_sample[lo:hi].sum(0) / gb if hi > lo else torch.zeros(6)             # adjoints scaled by 1 / global batch; empty shard: zeros
        work = ts.start_tail_allreduce(flat_g, 4)                                         # "Dense" bucket first ...
        ts.finish_allreduce(flat_g, 4, work)                                              # ... then the rest, then join
        out[gb] = (flat_g.tolist(), per_sample.mean(0).tolist())
    ret[rank] = out
    dist.destroy_process_group()


def test_ragged_and_empty_shards_give_the_global_mean():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker_ragged, args=(world, port, ret), nprocs=world, join=True)
        ret = dict(ret)
    for rank in (0, 1):
        for gb in (3, 1):
            got, want = ret[rank][gb]
            assert np.allclose(got, want, rtol=1e-6), (rank, gb, got, want)
