"""N>1 path on CPU: world_size-2 gloo ranks shard a batch with shard_bounds, run a per-image function on their
slice and gather -- the result must equal the single-process result (inference needs no data-path collective;
the only cross-rank traffic is the gather of results / the max-over-ranks of the timing)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _per_image_fn(x):
    """Stand-in for predict+explain on CPU ranks: the oracle's tiny network (per-image independent)."""
    sys.path.insert(0, ROOT)
    from oracle import cnn as ocnn
    cfg = ocnn.NetConfig.torch_flavour((12, 12, 1), 2, [(2, 3), (3, 3)], [4])
    p = ocnn.init_params(cfg, seed=3, bias_std=0.1)
    return ocnn.forward(cfg, p, x).logits.float()


def _worker(rank, world, port, n, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bcad_b200
    x = np.random.default_rng(0).standard_normal((n, 12, 12, 1)).astype(np.float32)
    bounds = bcad_b200.shard_bounds(n, world)
    s, e = bounds[rank]
    local = _per_image_fn(x[s:e]) if e > s else torch.zeros((0, 2))
    # gather variable-size shards on rank 0 (pad to the largest shard)
    cap = max(b[1] - b[0] for b in bounds)
    padded = torch.zeros((cap, 2))
    padded[: e - s] = local
    bufs = [torch.zeros((cap, 2)) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, bufs, dst=0)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the bench's max-over-ranks timing reduction
    if rank == 0:
        full = torch.cat([bufs[r][: bounds[r][1] - bounds[r][0]] for r in range(world)])
        torch.save({"full": full, "tmax": t}, out_path)
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [7, 2, 1])
def test_world2_sharded_equals_single(tmp_path, n):
    out = str(tmp_path / "out.pt")
    port = 29500 + (os.getpid() % 2000) + n
    mp.spawn(_worker, args=(2, port, n, out), nprocs=2, join=True)
    got = torch.load(out)
    x = np.random.default_rng(0).standard_normal((n, 12, 12, 1)).astype(np.float32)
    want = _per_image_fn(x)
    assert torch.equal(got["full"], want)
    assert float(got["tmax"]) == 2.0


def _ar_worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from bcad_b200.training import allreduce_mean_
    torch.manual_seed(rank)
    g = torch.randn(1000)
    mine = g.clone()
    allreduce_mean_(g, slice(100, 900))                  # fc1-weight bucket first, then the two small ones
    gathered = [torch.zeros(1000) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        torch.save({"avg": g, "want": sum(gathered) / world}, out_path)
    dist.destroy_process_group()


def test_gradient_allreduce_mean_world2(tmp_path):
    """The training step's only collective: bucketed all-reduce + /world == mean of the ranks' gradient vectors."""
    out = str(tmp_path / "ar.pt")
    mp.spawn(_ar_worker, args=(2, 29500 + (os.getpid() % 2000) + 17, out), nprocs=2, join=True)
    got = torch.load(out)
    assert torch.allclose(got["avg"], got["want"], atol=1e-6)
