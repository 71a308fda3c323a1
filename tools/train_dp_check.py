"""Data-parallel training step over NCCL vs the same global batch on one handle (run under torchrun, >= 2 GPUs).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/train_dp_check.py

Every rank trains on its shard with DataParallelTrainer (one bucketed gradient all-reduce per step); rank 0 also trains a
second handle on the whole batch.  After 3 Adam steps the weights must agree to 1e-5 and all ranks must hold the same
weights bit for bit.  Prints 'DP-OK' from rank 0, exits non-zero on mismatch.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bcad_b200  # noqa: E402
from bcad_b200.training import DataParallelTrainer  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec = bcad_b200.NetSpec.torch_flavour((32, 32, 1), 2, [(8, 3), (16, 3)], [32, 16], 0.01)
    rng = np.random.default_rng(0)
    cw = [rng.standard_normal((8, 3, 3, 1)).astype(np.float32) * 0.3, rng.standard_normal((16, 3, 3, 8)).astype(np.float32) * 0.1]
    cb = [rng.standard_normal(8).astype(np.float32) * 0.05, rng.standard_normal(16).astype(np.float32) * 0.05]
    dw = [rng.standard_normal((32, 8 * 8 * 16)).astype(np.float32) * 0.05, rng.standard_normal((16, 32)).astype(np.float32) * 0.2,
          rng.standard_normal((2, 16)).astype(np.float32) * 0.3]
    db = [rng.standard_normal(32).astype(np.float32) * 0.05, rng.standard_normal(16).astype(np.float32) * 0.05,
          rng.standard_normal(2).astype(np.float32) * 0.05]
    per = 8
    X = rng.standard_normal((3, per * world, 32, 32, 1)).astype(np.float32)
    Y = rng.integers(0, 2, (3, per * world))
    eng = bcad_b200.Engine(spec, precision="fp32", max_batch=per, keep_all_activations=True, device=local)
    eng.set_weights(cw, cb, dw, db)
    tr = DataParallelTrainer(eng, opt="adam", lr=1e-2)
    for s in range(3):
        tr.step(X[s, rank * per:(rank + 1) * per], Y[s, rank * per:(rank + 1) * per])
    mine = np.concatenate([a.ravel() for grp in eng.get_weights() for a in grp])
    t = torch.from_numpy(mine).cuda()
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    ok = True
    if rank == 0:
        for r in range(1, world):
            if not torch.equal(gathered[0], gathered[r]):
                print(f"rank {r} weights differ from rank 0: {float((gathered[0] - gathered[r]).abs().max())}")
                ok = False
        full = bcad_b200.Engine(spec, precision="fp32", max_batch=per * world, keep_all_activations=True, device=local)
        full.set_weights(cw, cb, dw, db)
        for s in range(3):
            x = torch.from_numpy(X[s]).cuda()
            full.predict(x)
            g, _ = full.train_backward(x, Y[s])
            full.apply_update(g, "adam", lr=1e-2)
        ref = np.concatenate([a.ravel() for grp in full.get_weights() for a in grp])
        err = float(np.abs(ref - mine).max())
        print(f"max |w_dp - w_full| = {err:.3e}")
        ok = ok and err < 1e-5
        print("DP-OK" if ok else "DP-MISMATCH")
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
