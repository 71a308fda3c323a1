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


class _FakeEngine:
    """Host stand-in for bcad_b200.Engine in the trainer's N>1 logic: the 'gradient' of a shard is the mean of per-sample
    vectors, the dense slice written by part 1 and the conv slice by part 2 (like bcad_train_backward_part)."""
    uses_tensor_path = False
    tdev = torch.device("cpu")
    N, DENSE0 = 64, 16                                   # flat vector: [conv 0..16 | dense 16..64], fc1 bucket = 24..56

    def __init__(self):
        self.applied = None

    def grad_layout(self, is_dense, index):
        return (24, 32, 56, 8)

    def _as_device_input(self, x):
        return torch.as_tensor(x)

    def predict(self, x):
        return torch.zeros(x.shape[0], dtype=torch.int32), None, None

    @staticmethod
    def sample_grad(i):
        return torch.arange(_FakeEngine.N, dtype=torch.float32) * 0.01 + float(i)

    def train_backward(self, x, labels, grads=None, part=0, loss=None):
        g = torch.zeros(self.N) if grads is None else grads
        mean = torch.stack([self.sample_grad(int(v)) for v in x[:, 0]]).mean(dim=0)
        if part in (0, 1):
            g[self.DENSE0:] = mean[self.DENSE0:]
        if part in (0, 2):
            g[:self.DENSE0] = mean[:self.DENSE0]
        return g, torch.zeros(x.shape[0])

    def apply_update(self, grads, *a, **k):
        self.applied = grads.clone()


def _trainer_worker(rank, world, port, out_path, overlap, equal):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from bcad_b200.training import DataParallelTrainer
    ids = [[0, 1, 2], [3, 4, 5]] if equal else [[0, 1, 2], [3]]              # sample ids per rank; unequal: 3 + 1
    eng = _FakeEngine()
    tr = DataParallelTrainer(eng, overlap=overlap, equal_shards=equal)
    tr.step(torch.tensor(ids[rank], dtype=torch.float32)[:, None], None)
    if rank == 0:
        n = sum(len(v) for v in ids)
        want = torch.stack([_FakeEngine.sample_grad(i) for v in ids for i in v]).mean(dim=0)
        torch.save({"got": eng.applied, "want": want, "n": n}, out_path)
    dist.destroy_process_group()


@pytest.mark.parametrize("overlap,equal", [(True, True), (True, False), (False, False)])
def test_data_parallel_trainer_step_is_the_global_batch_mean(tmp_path, overlap, equal):
    """DataParallelTrainer over 2 gloo ranks: two-part backward with the dense bucket all-reduced before the conv part runs, and
    shard-size weighting -- the update every rank applies is the mean gradient of the GLOBAL batch (Classes/CNNModel.py:459-464),
    also when the shards differ in size."""
    out = str(tmp_path / "tr.pt")
    mp.spawn(_trainer_worker, args=(2, 29500 + (os.getpid() % 2000) + 31 + 2 * overlap + equal, out, overlap, equal), nprocs=2, join=True)
    got = torch.load(out)
    assert torch.allclose(got["got"], got["want"], atol=1e-5)
