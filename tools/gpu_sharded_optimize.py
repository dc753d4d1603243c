"""Developer check (run under torchrun on N GPUs): `optimize_acqf(..., shard_across_ranks=True)` over NCCL must return
exactly what the unsharded call returns on one GPU (kernels are batch-split invariant, restarts are independent)."""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, ".")
from botorch_b200.benchmarks import configs
from botorch_b200.optim import optimize_acqf

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
    os.environ.pop("NCCL_DEBUG")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ok = True
for cfg in sys.argv[1:] or ["C1", "C2"]:
    spec = configs.CONFIGS[cfg]
    data = configs.make_problem(spec); model = configs.build_model(data, dev); acqf = configs.build_acqf(data, model)
    bounds = torch.stack([torch.zeros(spec.d), torch.ones(spec.d)]).to(dev, torch.float64)
    kw = dict(bounds=bounds, q=spec.q, num_restarts=spec.num_restarts, raw_samples=spec.raw_samples, options={"maxiter": 30, "seed": 0})
    out = {}
    for sharded in (False, True):
        optimize_acqf(acqf, shard_across_ranks=sharded, **kw)
        torch.manual_seed(1234)  # the Boltzmann selection of initial conditions draws from the global RNG
        torch.cuda.synchronize(); (dist.barrier() if world > 1 else None); t0 = time.perf_counter()
        c, v = optimize_acqf(acqf, shard_across_ranks=sharded, **kw)
        torch.cuda.synchronize(); out[sharded] = (c, v, time.perf_counter() - t0)
    same = torch.equal(out[False][0], out[True][0]) and torch.equal(out[False][1], out[True][1])
    ok &= same
    if rank == 0:
        print(f"{cfg}: world {world}: unsharded {out[False][2]*1e3:.1f} ms, sharded {out[True][2]*1e3:.1f} ms, value {float(out[True][1]):.9f}, identical: {same}")
if world > 1:
    flag = torch.tensor([int(ok)], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0: print("all ranks identical:", bool(flag.item()))
    dist.destroy_process_group()
