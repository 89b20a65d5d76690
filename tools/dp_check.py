"""Data-parallel correctness of the trunk train step on N GPUs (run under torchrun; `gpurun --gpus 2`):
every rank computes gradients on its own shard, ONE all-reduce sums the flat gradient buffers, and the averaged gradient
must equal what a single GPU computes on the concatenated batch (the trunk has no cross-sample coupling: instance norm is
per sample).  Prints one JSON line on rank 0."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from __graft_entry__ import load_package

def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pkg = load_package()
    Bl, h, w, C, k, nb = 4, 8, 32, 128, 3, 2
    rng = np.random.default_rng(0)
    xg = rng.standard_normal((Bl * world, h, w, C)).astype(np.float32)
    tg = rng.standard_normal((Bl * world, h, w, C)).astype(np.float32)
    lim = (6.0 / (k * k * C + C)) ** 0.5
    wts = [{f"conv{i}_kernel": rng.uniform(-lim, lim, (k * k * C, C)).astype(np.float32) for i in (1, 2)} for _ in range(nb)]
    def build(B):
        trunk = pkg.resLayer((C,) * nb, C, k_h=k, k_w=k)
        trunk.build((B, h, w, C))
        for unit, wt in zip(trunk.sequence, wts):
            unit.conv1.kernel.copy_(torch.from_numpy(wt["conv1_kernel"])); unit.conv2.kernel.copy_(torch.from_numpy(wt["conv2_kernel"]))
        return pkg.trunk_train.TrunkTrainer(trunk, (B, h, w, C), lr=1e-3)
    lo, hi = pkg.sharding.shard_bounds(Bl * world, rank, world)
    tr = build(hi - lo)
    x, t = torch.from_numpy(xg[lo:hi]).cuda(), torch.from_numpy(tg[lo:hi]).cuda()
    y = tr.forward(x); _, dy = tr.loss_and_grad(y, t); tr.backward(dy)
    pkg.trunk_train.allreduce_flat_(tr.flat_g)
    g_dp = tr.flat_g / world                       # mean over ranks of per-shard mean losses == global-batch mean loss
    out = {}
    if rank == 0:
        full = build(Bl * world)
        xf, tf = torch.from_numpy(xg).cuda(), torch.from_numpy(tg).cuda()
        yf = full.forward(xf); _, dyf = full.loss_and_grad(yf, tf); full.backward(dyf)
        rel = ((g_dp - full.flat_g).double().norm() / full.flat_g.double().norm()).item()
        out = {"check": "dp_gradients_equal_single_gpu", "world": world, "rel_l2": rel, "ok": rel < 1e-4,
               "flat_gradient_bytes": int(tr.flat_g.numel() * 4)}
        print(json.dumps(out), flush=True)
    dist.barrier(); dist.destroy_process_group()
    if rank == 0 and not out["ok"]:
        sys.exit(1)

if __name__ == "__main__":
    main()
