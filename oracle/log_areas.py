"""Pure-torch restatement of the qLogEHVI/qLogNEHVI fused log-area loop (TEST INFRASTRUCTURE ONLY).

Follows botorch/csrc/logei_fused.cpp: safe_softplus :39-48, cauchy :62-65, log_fatplus_fwd :85-92,
compute_fatmin :116-177 (n == 1 and hard-min branches), forward :184-272.  Pinned against the COMPILED reference
(`oracle/_ref`, built by oracle/build_ref.py from the reference source where it lies) in tests/test_log_areas.py.
Gradients by autograd.
"""
from __future__ import annotations

import torch
from torch import Tensor


def _safe_softplus(y: Tensor) -> Tensor:
    mid = torch.log1p(torch.exp(y.clamp(-20.0, 20.0)))
    return torch.where(y > 20, y, torch.where(y < -20, torch.exp(y.clamp_max(-20.0)), mid))


def _log_fatplus(x: Tensor, tau: float) -> Tensor:
    y = x * (1.0 / tau)
    f = _safe_softplus(y) + 0.1 / (1 + y * y)
    val = tau * f
    return torch.where(val > 0, torch.log(val.clamp_min(torch.finfo(x.dtype).tiny)), torch.full_like(val, -1e30))


def _fatmin(x: Tensor, tau: float) -> Tensor:
    """Smooth minimum over the last dim (reference compute_fatmin)."""
    if x.shape[-1] == 1:
        return x[..., 0]
    mn = x.amin(dim=-1, keepdim=True)
    z = (x - mn.detach() * 0 - mn) * (1.0 / tau)
    S = (2.0 / (2.0 + 2.0 * z + z * z)).sum(dim=-1)
    soft = mn.squeeze(-1) - tau * torch.log(S)
    return torch.where(mn.squeeze(-1) < -1e29, mn.squeeze(-1), soft)


def log_areas(obj_subsets: Tensor, cell_lower: Tensor, cell_upper: Tensor, tau_relu: float, tau_max: float) -> Tensor:
    """obj (B, n_sub, i, m); cells (nc, m) or (B, nc, m) -> (B, nc, n_sub)."""
    cl = cell_lower if cell_lower.dim() == 3 else cell_lower.unsqueeze(0)
    cu = cell_upper if cell_upper.dim() == 3 else cell_upper.unsqueeze(0)
    clamp = 1e10 if obj_subsets.dtype == torch.float64 else 1e8
    lcl = torch.log(cu.clamp_max(clamp) - cl)  # (Bc, nc, m)
    # (B, 1, n_sub, i, m) - (Bc, nc, 1, 1, m)
    diff = obj_subsets.unsqueeze(1) - cl.unsqueeze(2).unsqueeze(3)
    li = _log_fatplus(diff, tau_relu)  # (B, nc, n_sub, i, m)
    lim = _fatmin(li.transpose(-1, -2), tau_max)  # over i -> (B, nc, n_sub, m)
    pair = torch.stack([lim, lcl.unsqueeze(2).expand_as(lim)], dim=-1)
    return _fatmin(pair, tau_max).sum(dim=-1)
