"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: shard bounds, ragged all-gather of acquisition
values, sharded raw-sample sweep and restart split inside optimize_acqf give the single-process answer."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class QuadraticAcqf(torch.nn.Module):
    X_pending = None

    def __init__(self, c):
        super().__init__()
        self.c = c

    def set_X_pending(self, X):
        self.X_pending = X

    def forward(self, X):
        return -((X - self.c) ** 2).sum(dim=(-1, -2)) + 0.1 * torch.sin(7 * X).sum(dim=(-1, -2))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from botorch_b200.optim import optimize_acqf
    from botorch_b200.optim.sharded import all_gather_values, global_argmax, shard_bounds, sharded_evaluate

    torch.manual_seed(0)
    b = 11  # ragged: 6 + 5
    X = torch.rand(b, 2, 3, dtype=torch.float64)
    acqf = QuadraticAcqf(torch.tensor([0.3, 0.6, 0.4], dtype=torch.float64))
    full = sharded_evaluate(acqf, X)
    lo, hi = shard_bounds(b, rank, world)
    gathered = all_gather_values(acqf(X[lo:hi]), b)
    bounds = torch.stack([torch.zeros(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64)])
    torch.manual_seed(0)
    cand, val = optimize_acqf(acqf, bounds, q=2, num_restarts=5, raw_samples=37, options={"seed": 1},
                              shard_across_ranks=True)
    torch.save({"full": full, "gathered": gathered, "argmax": global_argmax(full), "cand": cand, "val": val},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


def test_shard_bounds_cover_range():
    from botorch_b200.optim.sharded import shard_bounds

    for b in (0, 1, 7, 64, 65536):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(b, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == b
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(180)
def test_world2_sharded_evaluate_and_optimize(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0 = torch.load(tmp_path / "rank0.pt")
    r1 = torch.load(tmp_path / "rank1.pt")
    torch.manual_seed(0)
    X = torch.rand(11, 2, 3, dtype=torch.float64)
    acqf = QuadraticAcqf(torch.tensor([0.3, 0.6, 0.4], dtype=torch.float64))
    ref = acqf(X)
    for r in (r0, r1):
        assert torch.equal(r["full"], ref) and torch.equal(r["gathered"], ref)
        assert r["argmax"] == (float(ref.max()), int(ref.argmax()))
    # every rank ends with the same candidate, equal to the single-process result
    from botorch_b200.optim import optimize_acqf

    bounds = torch.stack([torch.zeros(3, dtype=torch.float64), torch.ones(3, dtype=torch.float64)])
    torch.manual_seed(0)
    cand, val = optimize_acqf(acqf, bounds, q=2, num_restarts=5, raw_samples=37, options={"seed": 1})
    assert torch.equal(r0["cand"], r1["cand"]) and torch.equal(r0["val"], r1["val"])
    assert torch.equal(r0["cand"], cand) and torch.equal(r0["val"], val)
