"""Low-rank Cholesky-update sampling, torch-op version for the UNFUSED route (reference:
botorch/utils/low_rank.py:84-172).  The fused route does the same arithmetic inside
`csrc/sample_reduce.cu`."""
from __future__ import annotations

import torch
from torch import Tensor

from ..exceptions.errors import NanError
from ..models.prediction_strategy import psd_safe_cholesky


def sample_cached_cholesky(posterior, baseline_L: Tensor, q: int, base_samples: Tensor, sample_shape: torch.Size,
                           max_tries: int = 6) -> Tensor:
    """Samples at the q new points of a joint (X_baseline, X) posterior given chol of the baseline block."""
    mvn = posterior.distribution
    covar = mvn.covariance_matrix
    bottom = covar[..., -q:, :]
    r = bottom.shape[-1] - q
    bl, br = bottom.split([r, q], dim=-1)
    bl_chol = torch.linalg.solve_triangular(baseline_L, bl.transpose(-2, -1), upper=False).transpose(-2, -1)
    br_chol = psd_safe_cholesky(br - bl_chol @ bl_chol.transpose(-2, -1), max_tries=max_tries)
    new_Lq = torch.cat([bl_chol, br_chol], dim=-1)
    mean = mvn.mean
    S = sample_shape.numel()
    # base samples S x 1 x (r+q) x 1 -> (r+q) x S, broadcast over the t-batch
    Z = base_samples.reshape(S, r + q).t()
    res = new_Lq.matmul(Z) + mean[..., -q:].unsqueeze(-1)  # b x q x S
    res = res.permute(-1, *range(mean.dim() - 1), -2).unsqueeze(-1).contiguous()
    if torch.isnan(res).any() or torch.isinf(res).any():
        raise NanError("Samples contain nans or infs.")
    return res.view(*sample_shape, *res.shape[1:])
