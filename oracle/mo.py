"""Oracle for the qLogEHVI-style multi-objective value on independent GPs (TEST INFRASTRUCTURE ONLY).

Follows botorch/acquisition/multi_objective/logei.py:272-435 (`_compute_log_qehvi`, steps 1-9) with the fused-kernel
arithmetic of oracle/log_areas.py (pinned to the compiled reference kernel) and independent-output sampling as in
models/gpytorch.py:786-884 (block-diagonal MTMVN, non-interleaved base samples: output k <-> base-sample column k).
"""
from __future__ import annotations

from itertools import combinations

import torch
from torch import Tensor

from . import safe_math as sm
from .gp import OracleGP, mvn_rsample_from_base_samples
from .log_areas import log_areas
from .sampling import draw_sobol_normal_samples


class OracleQLogEHVI:
    def __init__(self, gps: list[OracleGP], cell_lower: Tensor, cell_upper: Tensor, S: int, seed: int,
                 tau_relu: float = sm.TAU_RELU, tau_max: float = sm.TAU_MAX) -> None:
        self.gps, self.cl, self.cu, self.S, self.seed = gps, cell_lower, cell_upper, S, seed
        self.tau_relu, self.tau_max = tau_relu, tau_max

    def __call__(self, X: Tensor) -> Tensor:
        if X.dim() == 2:
            X = X.unsqueeze(0)
        b, q, _ = X.shape
        m = len(self.gps)
        Z = draw_sobol_normal_samples(q * m, self.S, X.dtype, self.seed).view(self.S, 1, q, m)
        outs = []
        for k, gp in enumerate(self.gps):
            mean, cov = gp.posterior_mvn(X)
            zk = Z[..., k].expand(self.S, b, q)
            outs.append(mvn_rsample_from_base_samples(mean, cov, zk, (self.S,)).squeeze(-1))
        obj = torch.stack(outs, dim=-1)  # S x b x q x m
        nc = self.cl.shape[-2]
        seg = torch.full((self.S, b, nc, 2), -torch.inf, dtype=X.dtype)
        flat = obj.reshape(-1, q, m)
        for i in range(1, q + 1):
            idx = torch.tensor(list(combinations(range(q), i)), dtype=torch.long)
            sub = flat.index_select(-2, idx.view(-1)).view(flat.shape[0], *idx.shape, m)
            la = log_areas(sub, self.cl, self.cu, self.tau_relu, self.tau_max)
            la = sm.logsumexp(la.view(self.S, b, nc, -1), dim=-1)
            seg[..., i % 2] = sm.logplusexp(seg[..., i % 2].clone(), la)
        diff = sm.logdiffexp(log_a=seg[..., 0], log_b=seg[..., 1])
        return sm.logmeanexp(sm.logsumexp(diff, dim=-1), dim=0)
