"""Batch sharding of the path across the GPUs of one box: every conv / instance-norm term is per-sample,
so rank r simply owns a contiguous slice of the panoramas and no data-path collective exists for inference (north_star:
"independent per-GPU shards with no collective").  The only torch.distributed traffic of an inference run is the timing
protocol of bench.py (barrier + max over ranks).

Two terms of the reference couple the samples of a batch, and under sharding both are evaluated PER SHARD (a documented
deviation): tf.reduce_max(sunpose_pred) over the whole batch (generator.py:160) — so a sharded inference equals
the reference run on each shard's batch, not on the concatenated batch — and, in training, the BatchNormalization batch
statistics of sunRadNet / the discriminator.  Training normalises every loss adjoint by the GLOBAL batch (train.Step.train_step /
train_sun.SunTrainer.sun_train_step take `global_batch`), so the all-reduced SUM of the per-rank gradients is the gradient of
the global-batch mean for any shard sizes; a rank whose shard is empty contributes zeros and still joins the collectives."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(global_batch: int, rank: int, world: int):
    """[lo, hi) of the samples rank `rank` processes: ceil-sized leading shards, ragged tail, empty shards allowed."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    per = -(-global_batch // world)
    lo = min(rank * per, global_batch)
    return lo, min(lo + per, global_batch)


def max_over_ranks(values, device="cpu"):
    """Element-wise max of a list of floats over all ranks (device timings are reported as the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def gather_shards(local: torch.Tensor, global_batch: int):
    """Concatenate per-rank shards along batch on every rank (used by tests to compare a sharded run with a single-rank
    run; not on the inference hot path)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    per = -(-global_batch // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    parts = []
    for r, o in enumerate(out):
        lo, hi = shard_bounds(global_batch, r, world)
        parts.append(o[: hi - lo])
    return torch.cat(parts, dim=0)
