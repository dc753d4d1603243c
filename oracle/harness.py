"""Build the CPU oracle for a synthetic problem description (TEST INFRASTRUCTURE ONLY).

`data` is duck-typed: any object with `.spec.{kernel, outputscale, acqf, S}`, `.train_X`, `.train_Y`,
`.lengthscale`, `.noise`, `.best_f`, `.X_baseline` (what botorch_b200.benchmarks.configs.make_problem returns);
the oracle never imports the product package."""
from __future__ import annotations

import os
import time

import torch

from .acquisition import OracleQLogEI, OracleQLogNEI, value_and_grad
from .gp import OracleGP


def build_oracle(data, seed: int = 1234):
    gp = OracleGP(data.train_X, data.train_Y, data.lengthscale, torch.tensor(data.noise, dtype=torch.float64),
                  kernel=data.spec.kernel, outputscale=data.spec.outputscale, mean_constant=0.0)
    if data.spec.acqf == "qLogEI":
        return OracleQLogEI(gp, torch.tensor(data.best_f, dtype=torch.float64), data.spec.S, seed)
    return OracleQLogNEI(gp, data.X_baseline, data.spec.S, seed)


def time_cpu_fwd_bwd(acqf, X: torch.Tensor, chunk: int, warmup: int = 1, reps: int = 1, threads: int | None = None):
    """Time forward+backward of the oracle exactly like generation/gen.py:466-469 (losses.sum(), autograd.grad),
    chunked like `init_batch_limit`.  Returns (seconds_per_pass, threads)."""
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    chunks = X.split(chunk)
    for _ in range(warmup):
        value_and_grad(acqf, chunks[0])
    t0 = time.perf_counter()
    for _ in range(reps):
        for c in chunks:
            value_and_grad(acqf, c)
    return (time.perf_counter() - t0) / reps, threads
