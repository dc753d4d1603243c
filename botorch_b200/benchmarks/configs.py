"""The BASELINE.md configurations (C1-C3) as deterministic synthetic problems (BASELINE.md section 2,
SURVEY.md section 8d): scrambled-Sobol train inputs, Hartmann-6 / Ackley-20 targets (+ N(0, 0.05^2) noise for
the NEI configs), FIXED hyper-parameters (prior-mode lengthscales x U[0.5, 2], noise 1e-3, zero mean), Sobol
evaluation points, explicit baseline = the r best training points."""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
from torch import Tensor
from torch.quasirandom import SobolEngine

from ..test_functions.synthetic import Ackley, Hartmann
from ..utils.sampling import draw_sobol_samples


@dataclass
class ProblemSpec:
    name: str
    acqf: str  # "qLogEI" | "qLogNEI"
    kernel: str  # "rbf" | "matern52"
    n: int
    d: int
    q: int
    S: int
    raw_samples: int
    num_restarts: int
    r: int = 16
    outputscale: float | None = None


C1 = ProblemSpec("C1", "qLogEI", "rbf", n=64, d=6, q=4, S=512, raw_samples=512, num_restarts=20, r=0)
C2 = ProblemSpec("C2", "qLogNEI", "rbf", n=1024, d=20, q=8, S=1024, raw_samples=8192, num_restarts=64, r=16)
C3 = ProblemSpec("C3", "qLogNEI", "matern52", n=4096, d=20, q=8, S=1024, raw_samples=65536, num_restarts=64, r=16,
                 outputscale=1.0)
CONFIGS = {"C1": C1, "C2": C2, "C3": C3}


@dataclass
class ProblemData:
    spec: ProblemSpec
    train_X: Tensor  # n x d in [0, 1]^d
    train_Y: Tensor  # n x 1
    lengthscale: Tensor  # d
    noise: float
    X_baseline: Tensor | None  # r x d
    best_f: float
    bounds: Tensor  # 2 x d (unit cube)


def make_problem(spec: ProblemSpec, n: int | None = None) -> ProblemData:
    """All randomness is seeded; tensors are CPU fp64 (callers move them to their device)."""
    n = spec.n if n is None else n
    d = spec.d
    train_X = SobolEngine(d, scramble=True, seed=0).draw(n, dtype=torch.float64)
    gen = torch.Generator().manual_seed(0)
    if spec.acqf == "qLogEI":
        f = Hartmann(dim=6, negate=True) if d == 6 else Ackley(dim=d, negate=True)
    else:
        f = Ackley(dim=d, negate=True)
    lo, hi = f.bounds[0], f.bounds[1]
    Y = f(lo + (hi - lo) * train_X, noise=False).unsqueeze(-1)
    if spec.acqf == "qLogNEI":
        Y = Y + 0.05 * torch.randn(Y.shape, generator=gen, dtype=torch.float64)
    mode = math.exp(math.sqrt(2.0) + 0.5 * math.log(d) - 3.0)
    lengthscale = mode * (0.5 + 1.5 * torch.rand(d, generator=gen, dtype=torch.float64))
    r = min(spec.r, n)
    X_baseline = train_X[Y.squeeze(-1).topk(r).indices].clone() if r > 0 else None
    bounds = torch.stack([torch.zeros(d, dtype=torch.float64), torch.ones(d, dtype=torch.float64)])
    return ProblemData(spec, train_X, Y, lengthscale, 1e-3, X_baseline, float(Y.max()), bounds)


def eval_points(data: ProblemData, b: int, seed: int = 0) -> Tensor:
    """b x q x d Sobol q-batches: the generator `gen_batch_initial_conditions` uses (initializers.py:384)."""
    return draw_sobol_samples(bounds=data.bounds, n=b, q=data.spec.q, seed=seed)


def build_model(data: ProblemData, device):
    from ..models import MaternKernel, RBFKernel, ScaleKernel, SingleTaskGP

    d = data.spec.d
    base = (RBFKernel if data.spec.kernel == "rbf" else MaternKernel)(ard_num_dims=d, lengthscale=data.lengthscale)
    covar = ScaleKernel(base, outputscale=data.spec.outputscale) if data.spec.outputscale is not None else base
    model = SingleTaskGP(data.train_X.to(device), data.train_Y.to(device), covar_module=covar)
    model.likelihood.noise = data.noise
    model.mean_module.constant = 0.0
    return model.to(device)


def build_acqf(data: ProblemData, model, seed: int = 1234):
    from ..acquisition import qLogExpectedImprovement, qLogNoisyExpectedImprovement
    from ..sampling import SobolQMCNormalSampler

    sampler = SobolQMCNormalSampler(sample_shape=torch.Size([data.spec.S]), seed=seed)
    dev = model.train_inputs[0].device
    if data.spec.acqf == "qLogEI":
        return qLogExpectedImprovement(model, best_f=torch.tensor(data.best_f, dtype=torch.float64, device=dev), sampler=sampler)
    return qLogNoisyExpectedImprovement(model, X_baseline=data.X_baseline.to(dev), sampler=sampler, prune_baseline=False)
