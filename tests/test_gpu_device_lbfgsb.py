"""GPU: the device-resident L-BFGS-B (csrc/lbfgsb.cu).

(1) `mcacq_lbfgsb_step` against its CPU restatement oracle/lbfgsb.py (itself pinned to scipy in
    tests/test_lbfgsb_device_model.py): the same trial points round by round, the same iteration / evaluation counts and
    termination words.
(2) `gen_candidates_device` against `gen_candidates_scipy` on the benchmark problems: tolerance-based candidate parity
    (same minimisers where both converge, best value not worse), as section 8f N4 anticipates.
(3) `optimize_acqf` with `settings.optimizer("device")`, including the generic (autograd) evaluation route."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _quad(seed, N, D):
    rng = np.random.default_rng(seed)
    A = rng.normal(size=(N, D, D))
    Q = np.einsum("nij,nkj->nik", A, A) + 0.5 * np.eye(D)
    c = rng.normal(size=(N, D))

    def fun(X):
        d = X - c
        return 0.5 * np.einsum("ni,nij,nj->n", d, Q, d) + np.cos(3 * X).sum(-1), np.einsum("nij,nj->ni", Q, d) - 3 * np.sin(3 * X)

    return fun, rng.normal(size=(N, D))


def _rosen(X):
    f = (100 * (X[:, 1:] - X[:, :-1] ** 2) ** 2 + (1 - X[:, :-1]) ** 2).sum(-1)
    g = np.zeros_like(X)
    g[:, :-1] += -400 * X[:, :-1] * (X[:, 1:] - X[:, :-1] ** 2) - 2 * (1 - X[:, :-1])
    g[:, 1:] += 200 * (X[:, 1:] - X[:, :-1] ** 2)
    return f, g


CASES = {
    "quad6": (*_quad(0, 8, 6), -1.0, 1.5),
    "quad40": (*_quad(1, 5, 40), -0.5, 0.8),
    "quad161": (*_quad(2, 3, 161), -0.3, 0.4),       # odd D > 128: strided loops, padding of the state rows
    "rosen10": (_rosen, np.random.default_rng(3).uniform(-1.5, 1.5, size=(5, 10)), -2.0, 2.0),
    "rosen10_tight": (_rosen, np.random.default_rng(4).uniform(-0.5, 0.5, size=(5, 10)), -0.7, 0.9),
    "rosen12_unbounded": (_rosen, np.random.default_rng(5).uniform(-1.0, 1.0, size=(4, 12)), -np.inf, np.inf),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("maxiter", [200, 6])
def test_step_kernel_follows_the_cpu_model_round_by_round(name, maxiter):
    from botorch_b200.generation.device_gen import DeviceLBFGSB
    from oracle.lbfgsb import FG, LbfgsbState

    fun, x0, lo, hi = CASES[name]
    N, D = x0.shape
    l, u = np.full(D, lo), np.full(D, hi)
    states = [LbfgsbState(x0[i], l, u, maxiter=maxiter) for i in range(N)]
    opt = DeviceLBFGSB(torch.from_numpy(x0).to(DEV), torch.from_numpy(l).to(DEV), torch.from_numpy(u).to(DEV), maxiter=maxiter)
    rounds = 0
    worst = early = 0.0
    dev_active = N
    while any(s.task == FG for s in states) or dev_active > 0:
        Xh = np.stack([s.x for s in states])
        Xd = opt.X.cpu().numpy()
        # both sides evaluate at THEIR OWN points (like production), and the points must agree
        drift = float(np.abs(Xd - Xh).max() / max(1.0, np.abs(Xh).max()))
        worst = max(worst, drift)
        if rounds < 8:
            early = max(early, drift)
        f, g = fun(Xh)
        fd, gd = fun(Xd)
        for i, s in enumerate(states):
            if s.task == FG:
                s.step(f[i], g[i])
        opt.step(torch.from_numpy(fd).to(DEV), torch.from_numpy(np.ascontiguousarray(gd)).to(DEV))
        rounds += 1
        dev_active = opt.active()
        assert rounds < 2000
    assert opt.active() == 0
    fdev, status = opt.summary()
    for i, s in enumerate(states):
        task, msg, nit, nfev = status[i].tolist()
        if maxiter == 6:
            assert (nit, nfev) == (s.iter, s.nfev), (i, status[i].tolist(), s.iter, s.nfev, s.message)
        else:   # ~80 Rosenbrock iterations: a rounding difference may cost or save a line-search evaluation
            assert abs(nit - s.iter) <= 2 and abs(nfev - s.nfev) <= 3, (i, status[i].tolist(), s.iter, s.nfev, s.message)
        assert task == s.task
        assert abs(float(fdev[i]) - s.f) <= 1e-9 * max(1.0, abs(s.f))
    # trial points agree to rounding over the first rounds and for short runs; long runs (80 Rosenbrock iterations along a
    # curved valley) amplify the different summation orders through the curvature pairs -- mid-run trial points drift apart
    # by up to a few 1e-3 and come back together: iteration / evaluation counts, final values and final points agree
    assert early <= 1e-11
    assert worst <= (1e-12 if maxiter == 6 else 2e-2)
    Xfin = opt.X.cpu().numpy()
    assert np.abs(Xfin - np.stack([s.x for s in states])).max() <= (1e-12 if maxiter == 6 else 1e-3)


def _problem(cfg, n=None, **over):
    from dataclasses import replace

    from botorch_b200.benchmarks import configs

    spec = replace(configs.CONFIGS[cfg], **over)
    data = configs.make_problem(spec, n=n)
    model = configs.build_model(data, DEV)
    acqf = configs.build_acqf(data, model)
    return spec, data, model, acqf


@pytest.mark.parametrize("cfg,n,nr", [("C1", None, 12), ("C2", 256, 10), ("C3", 512, 8)])
def test_gen_candidates_device_matches_scipy_driver_to_tolerance(cfg, n, nr):
    from botorch_b200.benchmarks import configs
    from botorch_b200.generation import gen_candidates_device, gen_candidates_scipy

    spec, data, model, acqf = _problem(cfg, n=n)
    ics = configs.eval_points(data, nr, seed=7).to(DEV)
    opts = {"maxiter": 60}
    c_h, v_h = gen_candidates_scipy(ics, acqf, lower_bounds=0.0, upper_bounds=1.0, options=opts)
    c_d, v_d = gen_candidates_device(ics, acqf, lower_bounds=0.0, upper_bounds=1.0, options=opts)
    assert c_d.shape == c_h.shape and v_d.shape == v_h.shape
    assert float(c_d.min()) >= 0.0 and float(c_d.max()) <= 1.0
    # the two drivers run the same algorithm on values that agree to ~1e-12: restarts end in the same basin, and where the
    # iteration limit cuts both off the paths have diverged by rounding only
    dv = (v_d - v_h).abs() / v_h.abs().clamp_min(1e-12)
    assert float(dv.median()) < 1e-6
    assert float(v_d.max()) >= float(v_h.max()) - 1e-6 * abs(float(v_h.max()))
    status = gen_candidates_device.last_status
    assert int((status[:, 0] == 0).sum()) == 0   # no restart left active
    assert int(status[:, 2].max()) <= 60


def test_cuda_graph_and_eager_rounds_are_bit_identical():
    from botorch_b200.benchmarks import configs
    from botorch_b200.generation import gen_candidates_device

    spec, data, model, acqf = _problem("C2", n=192)
    ics = configs.eval_points(data, 6, seed=3).to(DEV)
    c_g, v_g = gen_candidates_device(ics, acqf, lower_bounds=0.0, upper_bounds=1.0, options={"maxiter": 25})
    c_e, v_e = gen_candidates_device(ics, acqf, lower_bounds=0.0, upper_bounds=1.0, options={"maxiter": 25, "cuda_graph": False})
    assert torch.equal(c_g, c_e) and torch.equal(v_g, v_e)


def test_optimize_acqf_with_the_device_optimizer_and_generic_route():
    from botorch_b200 import settings
    from botorch_b200.acquisition.objective import GenericMCObjective
    from botorch_b200.acquisition import qLogExpectedImprovement
    from botorch_b200.optim import optimize_acqf
    from botorch_b200.sampling import SobolQMCNormalSampler

    spec, data, model, acqf = _problem("C1")
    bounds = torch.stack([torch.zeros(spec.d), torch.ones(spec.d)]).to(DEV, torch.float64)
    kw = dict(bounds=bounds, q=spec.q, num_restarts=8, raw_samples=128, options={"maxiter": 40, "seed": 0})
    torch.manual_seed(0)
    c_h, v_h = optimize_acqf(acqf, **kw)
    torch.manual_seed(0)
    with settings.optimizer("device"):
        c_d, v_d = optimize_acqf(acqf, **kw)
    assert abs(float(v_d) - float(v_h)) <= 1e-5 * abs(float(v_h))
    # a custom objective keeps the acquisition function off the fused kernels: autograd feeds the same device state machines
    generic = qLogExpectedImprovement(model, best_f=torch.tensor(data.best_f, dtype=torch.float64, device=DEV),
                                      sampler=SobolQMCNormalSampler(sample_shape=torch.Size([spec.S]), seed=1234),
                                      objective=GenericMCObjective(lambda Y, X=None: Y[..., 0]))
    torch.manual_seed(0)
    with settings.optimizer("device"):
        c_g, v_g = optimize_acqf(generic, **kw)
    assert abs(float(v_g) - float(v_h)) <= 1e-5 * abs(float(v_h))
