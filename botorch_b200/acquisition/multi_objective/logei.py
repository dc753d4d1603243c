"""qLogEHVI-style multi-objective MC acquisition on independent-output models (reference:
botorch/acquisition/multi_objective/logei.py:172-435, `_compute_log_qehvi` :272-435).

The per-output posteriors come from the CUDA posterior kernels, the inclusion-exclusion inner loop (steps 1-4 of the
reference) is the CUDA `fused_log_areas` kernel -- the reference only has that kernel on the CPU -- and the outer
log-space reductions (steps 6-9) are the torch ops of `utils/safe_math.py`.  Box decompositions stay on the host: the
class takes a `partitioning` object exposing `get_hypercell_bounds()` (as `FastNondominatedPartitioning` does) or the
cell bounds directly.
"""
from __future__ import annotations

from itertools import combinations

import torch
from torch import Tensor

from ...sampling.base import MCSampler
from ...utils.safe_math import logdiffexp, logmeanexp, logplusexp, logsumexp
from ...utils.transforms import concatenate_pending_points, t_batch_mode_transform
from ..acquisition import AcquisitionFunction, MCSamplerMixin
from ..logei import TAU_MAX, TAU_RELU, check_tau
from ... import settings
from .fused_log_areas import fused_log_areas, fused_log_hvi, log_hvi_fusable


def compute_subset_indices(q: int, device=None) -> dict[str, Tensor]:
    """All size-i subsets of range(q), i = 1..q (reference: utils/multi_objective/hypervolume.py compute_subset_indices)."""
    return {f"q_choose_{i}": torch.tensor(list(combinations(range(q), i)), dtype=torch.long, device=device)
            for i in range(1, q + 1)}


class qLogExpectedHypervolumeImprovement(AcquisitionFunction, MCSamplerMixin):
    _log = True

    def __init__(self, model, ref_point=None, partitioning=None, sampler: MCSampler | None = None, X_pending=None,
                 cell_bounds: tuple[Tensor, Tensor] | None = None, fat: bool = True, tau_relu: float = TAU_RELU,
                 tau_max: float = TAU_MAX) -> None:
        AcquisitionFunction.__init__(self, model=model)
        MCSamplerMixin.__init__(self, sampler=sampler)
        if not fat:
            raise NotImplementedError("The CUDA log-areas kernel implements the fat (default) variant.")
        if cell_bounds is None:
            if partitioning is None:
                raise ValueError("Provide either `partitioning` or `cell_bounds`.")
            cell_bounds = tuple(partitioning.get_hypercell_bounds())
        self.register_buffer("cell_lower_bounds", cell_bounds[0])
        self.register_buffer("cell_upper_bounds", cell_bounds[1])
        self.tau_relu = check_tau(tau_relu, name="tau_relu")
        self.tau_max = check_tau(tau_max, name="tau_max")
        self.fat = fat
        self.q_out = -1
        self.q_subset_indices: dict[str, Tensor] = {}
        self.set_X_pending(X_pending)

    def compute_q_subset_indices(self, q_out: int, device) -> dict[str, Tensor]:
        if q_out != self.q_out:
            self.q_subset_indices = compute_subset_indices(q_out, device=device)
            self.q_out = q_out
        return self.q_subset_indices

    def _compute_log_qehvi(self, samples: Tensor, X: Tensor | None = None) -> Tensor:
        obj = samples  # mc_samples x batch_shape x q x m  (identity multi-output objective)
        q = obj.shape[-2]
        flat3 = obj.reshape(-1, q, obj.shape[-1])
        if settings.fused_log_hvi.value() and log_hvi_fusable(flat3, self.cell_lower_bounds):
            # steps 1-8 in one launch: every li[j][k] once per cell, subsets as bit masks, streaming log-sum-exps
            per_sample = fused_log_hvi(flat3, self.cell_lower_bounds, self.cell_upper_bounds, self.tau_relu, self.tau_max)
            return logmeanexp(per_sample.view(obj.shape[:-2]), dim=0)
        idx = self.compute_q_subset_indices(q_out=q, device=obj.device)
        batch_shape = obj.shape[:-2]
        nc = self.cell_lower_bounds.shape[-2]
        log_areas_per_segment = torch.full((*batch_shape, nc, 2), -torch.inf, dtype=obj.dtype, device=obj.device)
        flat = obj.reshape(-1, q, obj.shape[-1])
        for i in range(1, q + 1):
            q_choose_i = idx[f"q_choose_{i}"]
            sub = flat.index_select(dim=-2, index=q_choose_i.view(-1)).view(flat.shape[0], *q_choose_i.shape, flat.shape[-1])
            log_areas_i = fused_log_areas(sub.contiguous(), self.cell_lower_bounds, self.cell_upper_bounds,
                                          self.tau_relu, self.tau_max)  # (B, nc, n_sub)
            log_areas_i = logsumexp(log_areas_i.view(*batch_shape, nc, -1), dim=-1)
            log_areas_per_segment[..., i % 2] = logplusexp(log_areas_per_segment[..., i % 2], log_areas_i)
        seg = logdiffexp(log_a=log_areas_per_segment[..., 0], log_b=log_areas_per_segment[..., 1])
        return logmeanexp(logsumexp(seg, dim=-1), dim=0)

    @concatenate_pending_points
    @t_batch_mode_transform()
    def forward(self, X: Tensor) -> Tensor:
        posterior = self.model.posterior(X)
        samples = self.get_posterior_samples(posterior)
        return self._compute_log_qehvi(samples=samples, X=X)
