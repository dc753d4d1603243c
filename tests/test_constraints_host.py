"""Outcome-constraint helpers: the oracle restatement AND the product's host functions against vectors generated from
the reference's botorch/utils/objective.py / safe_math.py (tests/golden/make_golden_constraints.py).  CPU only."""
import os

import pytest
import torch

G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "constraints_ref.pt"))
CONS = [lambda Y: Y[..., 1] - 0.2, lambda Y: -Y[..., 0] - 1.0]
ETA = torch.tensor([1e-2, 0.5], dtype=torch.float64)


def _close(a, b, tol=1e-13):
    fin = torch.isfinite(b)
    assert torch.equal(torch.isfinite(a), fin)
    assert torch.equal(a[~fin], b[~fin])
    assert float(((a[fin] - b[fin]).abs() / b[fin].abs().clamp_min(1e-300)).max()) <= tol


def test_oracle_scalar_functions_match_reference():
    from oracle import constraints as oc

    _close(oc.log1pexp(G["x"]), G["log1pexp"])
    _close(oc.logexpit(G["x"]), G["logexpit"])
    _close(oc.fatmoid(G["x"], tau=0.7), G["fatmoid"])
    _close(torch.log(oc.fatmoid(G["x"], tau=0.7)), G["log_fatmoid"])
    _close(torch.exp(oc.logexpit(G["x"])), G["sigmoid"])


def test_host_scalar_functions_bit_exact():
    from botorch_b200.utils import safe_math as sm

    assert torch.equal(sm.log1pexp(G["x"]), G["log1pexp"])
    assert torch.equal(sm.logexpit(G["x"]), G["logexpit"])
    assert torch.equal(sm.fatmoid(G["x"], tau=0.7), G["fatmoid"])
    assert torch.equal(sm.log_fatmoid(G["x"], tau=0.7), G["log_fatmoid"])
    assert torch.equal(sm.sigmoid(G["x"], log=True, fat=True), G["sigmoid_fat_log"])
    assert torch.equal(sm.sigmoid(G["x"]), G["sigmoid"])


@pytest.mark.parametrize("log", [False, True])
@pytest.mark.parametrize("fat", [False, True])
def test_smoothed_indicator_matches_reference(log, fat):
    from botorch_b200.utils.objective import compute_smoothed_feasibility_indicator
    from oracle.constraints import smoothed_feasibility

    want = G[f"smoothed_log{int(log)}_fat{int(fat)}"]
    assert torch.equal(compute_smoothed_feasibility_indicator(CONS, G["samples"], eta=ETA, log=log, fat=fat), want)
    _close(smoothed_feasibility(CONS, G["samples"], ETA, log=log, fat=fat), want, tol=1e-12)


def test_indicator_marginalisation_gradient_and_errors():
    from botorch_b200.utils.objective import compute_feasibility_indicator, compute_smoothed_feasibility_indicator
    from oracle.constraints import feasibility_indicator

    assert torch.equal(compute_feasibility_indicator(CONS, G["samples"]), G["indicator"])
    assert torch.equal(feasibility_indicator(CONS, G["samples"]), G["indicator"])
    assert torch.equal(compute_feasibility_indicator(CONS, G["samples"], marginalize_dim=-3), G["indicator_marg"])
    assert torch.equal(compute_smoothed_feasibility_indicator(CONS, G["samples"], eta=1e-3, log=True, fat=True),
                       G["smoothed_scalar_eta"])
    x = G["samples"].clone().requires_grad_(True)
    val = compute_smoothed_feasibility_indicator(CONS, x, eta=1e-1, log=True, fat=True)
    assert torch.equal(torch.autograd.grad(val.sum(), x)[0], G["smoothed_grad"])
    with pytest.raises(ValueError):
        compute_smoothed_feasibility_indicator(CONS, G["samples"], eta=torch.tensor([1e-3]))
    with pytest.raises(ValueError):
        compute_smoothed_feasibility_indicator(CONS, G["samples"], eta=1e-3, fat=[True])
    with pytest.raises(ValueError):
        compute_smoothed_feasibility_indicator(CONS, G["samples"], eta=-1.0)
    # fat = None: the callable already returns a probability
    p = compute_smoothed_feasibility_indicator([lambda Y: torch.sigmoid(Y[..., 0])], G["samples"], eta=1.0, log=False, fat=[None])
    assert torch.allclose(p, torch.sigmoid(G["samples"][..., 0]))
