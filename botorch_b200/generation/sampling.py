"""`MaxPosteriorSampling` -- Thompson sampling over a discrete candidate set (reference:
botorch/generation/sampling.py:58-155; TuRBO's candidate selection, tutorials/turbo_1).

Joint posterior over the N candidates on the GPU: cross-covariance kernel, `K R` contraction and the `A A^T` SYRK
run in the hand-written DMMA kernels; the N x N Cholesky is one cuSOLVER call (`psd_safe_cholesky` semantics); the
samples `mean + L z` use the triangular-aware DMMA kernel.  Like the reference, base samples are iid `torch.randn`
on the model's device (NOT seeded Sobol samples), so only statistical parity with the reference is defined unless
the same device generator state is used.  Independent t-batches of X (e.g. TuRBO trust regions) can be split
across ranks with `optim.sharded.shard_bounds`.
"""
from __future__ import annotations

import torch
from torch import Tensor
from torch.nn import Module

from ..acquisition.objective import IdentityMCObjective, MCAcquisitionObjective, PosteriorTransform
from ..models.gp_regression import SingleTaskGP
from ..models.prediction_strategy import psd_safe_cholesky


def _flip_sub_unique(x: Tensor, k: int) -> Tensor:
    """First k unique elements of a 1-d tensor, traversing it from the back (reference:
    botorch/generation/utils.py:21-52)."""
    seen, keep = set(), []
    n = len(x)
    for j, xi in enumerate(reversed(x.tolist())):
        if xi not in seen:
            seen.add(xi)
            keep.append(n - 1 - j)
        if len(seen) >= k:
            break
    return x[torch.tensor(keep, dtype=torch.long, device=x.device)]


class SamplingStrategy(Module):
    pass


class MaxPosteriorSampling(SamplingStrategy):
    def __init__(self, model, objective: MCAcquisitionObjective | None = None,
                 posterior_transform: PosteriorTransform | None = None, replacement: bool = True) -> None:
        super().__init__()
        self.model = model
        self.objective = IdentityMCObjective() if objective is None else objective
        self.posterior_transform = posterior_transform
        self.replacement = replacement

    def forward(self, X: Tensor, num_samples: int = 1, observation_noise: bool = False) -> Tensor:
        """X: batch_shape x N x d -> batch_shape x num_samples x d rows of X picked by posterior arg-max."""
        fast = isinstance(self.model, SingleTaskGP) and self.posterior_transform is None and not observation_noise
        if not fast:
            posterior = self.model.posterior(X, observation_noise=observation_noise,
                                             posterior_transform=self.posterior_transform)
            samples = posterior.rsample(sample_shape=torch.Size([num_samples]))
            return self.maximize_samples(X, samples, num_samples)
        strat = self.model.prediction_strategy()
        batch_shape, N, d = X.shape[:-2], X.shape[-2], X.shape[-1]
        Xf = X.reshape(-1, N, d).to(device=strat.device, dtype=torch.float64)
        outs = []
        with torch.no_grad():
            for xb in Xf:
                if N % 2:
                    # the DMMA SYRK / TRMM kernels move 16-byte row chunks (even leading dimensions): evaluate the joint
                    # posterior with one extra point (the centroid) and keep the N x N block of the candidates
                    mean, covar = strat.joint_posterior(torch.cat([xb, xb.mean(dim=0, keepdim=True)]))
                    mean, covar = mean[:N], covar[:N, :N].contiguous()
                else:
                    mean, covar = strat.joint_posterior(xb)
                chol = psd_safe_cholesky(covar, max_tries=6)
                Z = torch.randn(num_samples, N, device=strat.device, dtype=torch.float64)
                if N % 2:
                    # the triangular DMMA kernel moves 16-byte row chunks: pad the factor with one decoupled unit row /
                    # column and the base samples with a zero column (the extra output row is dropped)
                    chol = torch.nn.functional.pad(chol, (0, 1, 0, 1))
                    chol[N, N] = 1.0
                    Zp = torch.nn.functional.pad(Z, (0, 1))
                    Y = strat.lower_times_samples(chol, Zp)[:N]
                else:
                    Y = strat.lower_times_samples(chol, Z)  # N x num_samples
                outs.append((Y + mean.unsqueeze(-1)).t())
        samples = torch.stack(outs, dim=1).reshape(num_samples, *batch_shape, N, 1)
        return self.maximize_samples(X, samples, num_samples)

    def maximize_samples(self, X: Tensor, samples: Tensor, num_samples: int = 1) -> Tensor:
        obj = self.objective(samples, X=X)  # num_samples x batch_shape x N
        if self.replacement:
            idcs = torch.argmax(obj, dim=-1)
        else:
            # de-duplication exactly as the reference (:116-136): lower triangle of the per-sample top-k index
            # matrix in row-major order, traversed from the back, first `num_samples` unique entries
            _, idcs_full = torch.topk(obj, num_samples, dim=-1)
            ridx, cindx = torch.tril_indices(num_samples, num_samples)
            sub_idcs = idcs_full[ridx, ..., cindx]
            if sub_idcs.ndim == 1:
                idcs = _flip_sub_unique(sub_idcs, num_samples)
            elif sub_idcs.ndim == 2:
                idcs = torch.stack([_flip_sub_unique(sub_idcs[:, i], num_samples) for i in range(sub_idcs.size(-1))], dim=-1)
            else:
                raise NotImplementedError("MaxPosteriorSampling without replacement for more than a single batch "
                                          "dimension is not yet implemented.")
        if idcs.ndim > 1:
            idcs = idcs.permute(*range(1, idcs.ndim), 0)
        idcs = idcs.unsqueeze(-1).expand(*idcs.shape, X.size(-1)).contiguous()
        Xe = X.expand(*obj.shape[1:], X.size(-1))
        return torch.gather(Xe, -2, idcs)
