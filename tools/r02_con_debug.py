import sys, math, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_gpu_fused_n3 import _setup, _sampler, DEV
from botorch_b200.acquisition import qLogExpectedImprovement
from botorch_b200.acquisition.objective import GenericMCObjective
from botorch_b200.utils.safe_math import log_fatplus, fatmax, logmeanexp, log_fatmoid
spec, data, model, X = _setup(S=128)
med = float(data.train_Y.median())
cons = [lambda Y: Y[..., 0] - (med + 0.4), lambda Y: -2.0 * Y[..., 0] + 2.0 * (med - 1.5)]
best = torch.tensor(med - 0.3, dtype=torch.float64, device=DEV)
for eta in (2e-2, torch.tensor([2e-2, 5e-2])):
    f = qLogExpectedImprovement(model, best_f=best, sampler=_sampler(128), constraints=cons, eta=eta)
    g = qLogExpectedImprovement(model, best_f=best, sampler=_sampler(128), constraints=cons, eta=eta, objective=GenericMCObjective(lambda Y, X=None: Y[..., 0]))
    vf, vg = f(X), g(X)
    post = model.posterior(X)
    mean, cov = post.mean.squeeze(-1), post.distribution.covariance_matrix
    L = torch.linalg.cholesky(cov)
    Z = f.sampler.base_samples.reshape(128, spec.q).to(DEV)
    y = mean.unsqueeze(0) + torch.einsum("bij,sj->sbi", L, Z)
    e = float(eta if not isinstance(eta, torch.Tensor) else eta[0]); e2 = float(eta if not isinstance(eta, torch.Tensor) else eta[1])
    li = log_fatplus(y - best, tau=1e-6) + log_fatmoid(-(y - (med + 0.4)) / e) + log_fatmoid(-(-2.0 * y + 2.0 * (med - 1.5)) / e2)
    man = logmeanexp(fatmax(li, dim=-1, tau=1e-2), dim=0)
    nocon = logmeanexp(fatmax(log_fatplus(y - best, tau=1e-6), dim=-1, tau=1e-2), dim=0)
    print("eta", eta, "cons tuple", f._fused_constraints())
    print(" fused  ", vf[:4].tolist()); print(" generic", vg[:4].tolist()); print(" manual ", man[:4].tolist()); print(" nocon  ", nocon[:4].tolist())
