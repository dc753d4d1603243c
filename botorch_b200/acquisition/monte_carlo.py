"""MC acquisition skeleton (reference: botorch/acquisition/monte_carlo.py:60-135, 146-348):
posterior -> sampler -> objective -> `_sample_forward` -> q-reduction -> sample-reduction.
This generic torch-op route is the UNFUSED path (custom objectives / constraints); it still obtains
the posterior from the CUDA kernels.  qLogEI / qLogNEI override `forward` with the fused kernel when
their configuration allows it (acquisition/logei.py)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from functools import partial
from typing import Callable

import torch
from torch import Tensor

from ..exceptions.errors import UnsupportedError
from ..sampling.base import MCSampler
from ..utils.objective import compute_smoothed_feasibility_indicator
from ..utils.transforms import concatenate_pending_points, is_ensemble, t_batch_mode_transform
from .acquisition import AcquisitionFunction, MCSamplerMixin
from .objective import IdentityMCObjective, MCAcquisitionObjective, PosteriorTransform


class MCAcquisitionFunction(AcquisitionFunction, MCSamplerMixin, ABC):
    def __init__(self, model, sampler: MCSampler | None = None, objective: MCAcquisitionObjective | None = None,
                 posterior_transform: PosteriorTransform | None = None, X_pending: Tensor | None = None) -> None:
        super().__init__(model=model)
        MCSamplerMixin.__init__(self, sampler=sampler)
        if objective is None and model.num_outputs != 1:
            if posterior_transform is None:
                raise UnsupportedError("Must specify an objective or a posterior transform when using a multi-output model.")
            elif not posterior_transform.scalarize:
                raise UnsupportedError("If using a multi-output model without an objective, posterior_transform must "
                                       "scalarize the output.")
        self._identity_objective = objective is None or isinstance(objective, IdentityMCObjective)
        if objective is None:
            objective = IdentityMCObjective()
        self.posterior_transform = posterior_transform
        self.objective = objective
        self.set_X_pending(X_pending)

    def _get_samples_and_objectives(self, X: Tensor) -> tuple[Tensor, Tensor]:
        posterior = self.model.posterior(X=X, posterior_transform=self.posterior_transform)
        samples = self.get_posterior_samples(posterior)
        return samples, self.objective(samples=samples, X=X)


class SampleReducingMCAcquisitionFunction(MCAcquisitionFunction):
    _log: bool = False

    def __init__(self, model, sampler=None, objective=None, posterior_transform=None, X_pending=None,
                 sample_reduction: Callable = torch.mean, q_reduction: Callable = torch.amax,
                 constraints: list[Callable[[Tensor], Tensor]] | None = None, eta: Tensor | float = 1e-3,
                 fat: bool = False) -> None:
        super().__init__(model=model, sampler=sampler, objective=objective, posterior_transform=posterior_transform,
                         X_pending=X_pending)
        sample_dim = tuple(range(len(self.sample_shape)))
        if is_ensemble(model):
            sample_dim = sample_dim + (-1,)
        self._sample_reduction = partial(sample_reduction, dim=sample_dim)
        self._q_reduction = partial(q_reduction, dim=-1)
        self._constraints = constraints
        self._eta = eta
        self._fat = fat

    @concatenate_pending_points
    @t_batch_mode_transform()
    def forward(self, X: Tensor) -> Tensor:
        return self._sample_reduction(self._q_reduction(self._non_reduced_forward(X=X)))

    def _non_reduced_forward(self, X: Tensor) -> Tensor:
        samples, obj = self._get_samples_and_objectives(X)
        acqval = self._sample_forward(obj)
        return self._apply_constraints(acqval=acqval, samples=samples)

    @abstractmethod
    def _sample_forward(self, obj: Tensor) -> Tensor:
        ...

    def _apply_constraints(self, acqval: Tensor, samples: Tensor) -> Tensor:
        """Weight the per-sample utility by the smoothed feasibility indicator (reference :322-348)."""
        if self._constraints is not None:
            if not self._log and (acqval < 0).any():
                raise ValueError("Constraint-weighting requires unconstrained acquisition values to be non-negative.")
            ind = compute_smoothed_feasibility_indicator(constraints=self._constraints, samples=samples, eta=self._eta,
                                                         log=self._log, fat=self._fat)
            acqval = acqval.add(ind) if self._log else acqval.mul(ind)
        return acqval


def __getattr__(name: str):
    """qExpectedImprovement & co. live in `mc_improvement` (they build on the fused classes of `logei`, which imports
    this module); expose them under the reference's module path lazily."""
    if name in ("qExpectedImprovement", "qNoisyExpectedImprovement", "qProbabilityOfImprovement", "qSimpleRegret",
                "qUpperConfidenceBound", "qLowerConfidenceBound", "qPosteriorStandardDeviation"):
        from . import mc_improvement

        return getattr(mc_improvement, name)
    raise AttributeError(name)
