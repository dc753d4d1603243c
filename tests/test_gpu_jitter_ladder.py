"""GPU: `psd_safe_cholesky`'s jitter ladder inside `sample_reduce` with covariance blocks of a KNOWN escalation level
(SURVEY.md hard part #2; reference: linear_operator psd_safe_cholesky as called by botorch/utils/low_rank.py:137-140 and
posteriors/gpytorch.py:119-124, `cholesky_max_tries = 6`, jitter 1e-8 * 10^i on the failing batch elements only).

The blocks are built as Q diag(lambda) Q^T with a prescribed smallest eigenvalue: lambda_min = -3 * 10^-(9-k) needs exactly
k + 1 escalations (jitter 10^-(8-k) is the first one above 3 * 10^-(9-k)), with a factor 3 of margin on either side, so the
level does not depend on rounding.  Values and factors are compared with the oracle's `psd_safe_cholesky` at that level,
through the C ABI (`mcacq_sample_reduce_forward`).  Also: NotPSD signalling, and the qLogNEI fallback branch."""
import ctypes as C
import warnings

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _blocks(q, lam_mins, seed=0):
    g = torch.Generator().manual_seed(seed)
    mats = []
    for lm in lam_mins:
        Q, _ = torch.linalg.qr(torch.randn(q, q, generator=g, dtype=torch.float64))
        lam = torch.cat([torch.tensor([lm], dtype=torch.float64), 0.5 + torch.rand(q - 1, generator=g, dtype=torch.float64)])
        A = (Q * lam) @ Q.T
        mats.append(0.5 * (A + A.T))
    return torch.stack(mats)


def _run(mean, Sxx, Sxb, L_base, Zt, best, fat=1):
    from botorch_b200 import _lib

    L = _lib.lib()
    b, q = mean.shape
    r = 0 if Sxb is None else Sxb.shape[-1]
    f64 = dict(device=DEV, dtype=torch.float64)
    dv = lambda t: None if t is None else t.to(**f64).contiguous()  # noqa: E731
    mean_d, Sxx_d, Sxb_d, Lb_d, Zt_d, best_d = dv(mean), dv(Sxx), dv(Sxb), dv(L_base), dv(Zt), dv(best)
    acq = torch.empty(b, **f64)
    info = torch.empty(b, dtype=torch.int32, device=DEV)
    Bm = torch.empty(b, q, max(r, 1), **f64)
    Cm = torch.empty(b, q, q, **f64)
    mc = _lib.MC(S=Zt.shape[1], fat=fat, tau_relu=1e-6, tau_max=1e-2, Zt=Zt_d.data_ptr(), best=best_d.data_ptr(), obj_weight=1.0,
                 obj_offset=0.0, util_param=0.0, Zbar=None, n_con=0, con_fat=0,
                 jitter_f32=int(torch.get_default_dtype() == torch.float32))
    base = None
    if r > 0:
        base = _lib.Baseline(r=r, _pad=0, U_base=None, A_base=None, L_base=Lb_d.data_ptr(), A_base_absmax=None)
    rc = L.mcacq_sample_reduce_forward(C.byref(base) if base is not None else None, C.byref(mc), mean_d.data_ptr(),
                                       Sxx_d.data_ptr(), _lib.ptr(Sxb_d), b, q, acq.data_ptr(), info.data_ptr(), Bm.data_ptr(),
                                       Cm.data_ptr(), _lib.stream_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    return acq.cpu(), info.cpu(), Bm.cpu(), Cm.cpu()


def _oracle_value(mean, T, Bmat, Z, best):
    """logmeanexp_S fatmax_q log_fatplus(mean + B z_b + C z_q - best) with C = the oracle's psd_safe_cholesky(T)."""
    from oracle.gp import psd_safe_cholesky
    from oracle.safe_math import fatmax, log_fatplus, logmeanexp

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        Cf = psd_safe_cholesky(T, max_tries=6)
    r = 0 if Bmat is None else Bmat.shape[-1]
    y = mean + Z[:, r:] @ Cf.T
    if r > 0:
        y = y + Z[:, :r] @ Bmat.T
    li = log_fatplus(y - best.unsqueeze(-1), tau=1e-6)
    return logmeanexp(fatmax(li, dim=-1, tau=1e-2), dim=0), Cf


@pytest.mark.parametrize("q", [3, 8])
def test_known_escalation_levels_match_the_oracle(q):
    from botorch_b200 import _lib

    lam = [1e-3] + [-3.0 * 10.0 ** -(9 - k) for k in range(6)] + [-1.0]   # levels 0, 1..6, not PSD
    Sxx = _blocks(q, lam)
    b, S = len(lam), 64
    g = torch.Generator().manual_seed(1)
    mean = torch.randn(b, q, generator=g, dtype=torch.float64)
    Z = torch.randn(S, q, generator=g, dtype=torch.float64)
    best = torch.full((S,), 0.3, dtype=torch.float64)
    acq, info, _, Cm = _run(mean, Sxx, None, None, Z.T.contiguous(), best)
    levels = (info & _lib.INFO_JITTER_MASK).tolist()
    assert levels[:7] == [0, 1, 2, 3, 4, 5, 6]
    assert all((int(info[i]) & _lib.INFO_NOT_PSD) == 0 for i in range(7))
    assert int(info[7]) & _lib.INFO_NOT_PSD and torch.isnan(acq[7])
    eps = torch.finfo(torch.float64).eps
    for i in range(7):
        ref, Cf = _oracle_value(mean[i], Sxx[i], None, Z, best)
        # two correctly rounded fp64 factorisations of a matrix whose smallest eigenvalue (after the jitter) is lam_eff agree
        # to ~eps / lam_eff, not to eps: the bar is 1e-9 down to lam_eff = 2e-6 and 10 eps / lam_eff below (3e-7 at level 1)
        lam_eff = lam[i] + (0.0 if i == 0 else 1e-8 * 10.0 ** (i - 1))
        tol = max(1e-9, 10.0 * eps / lam_eff)
        assert abs(float(acq[i]) - float(ref)) <= tol * max(abs(float(ref)), 1.0), (i, float(acq[i]), float(ref))
        assert float((Cm[i] - Cf).abs().max()) <= tol * float(Cf.abs().max()), i
    # the Python layer turns the status words into the reference's warnings / errors
    from botorch_b200.acquisition._fused import _raise_on_info
    from botorch_b200.exceptions.errors import NotPSDError
    from botorch_b200.exceptions.warnings import NumericalWarning

    with pytest.warns(NumericalWarning) as rec:
        _raise_on_info(info[:7].to(DEV))
    msgs = [str(w.message) for w in rec]
    assert len(msgs) == 6 and "1.0e-08" in msgs[0] and "1.0e-03" in msgs[5]
    with pytest.raises(NotPSDError):
        _raise_on_info(info.to(DEV))


def test_cached_root_path_with_known_levels():
    """r > 0 (`sample_cached_cholesky`): the jitter applies to Sxx - B B^T, B = Sxb L_base^{-T}."""
    from botorch_b200 import _lib

    q, r, S = 4, 6, 48
    lam = [2e-3, -3e-8, -3e-6]   # levels 0, 2, 4
    T = _blocks(q, lam, seed=5)
    g = torch.Generator().manual_seed(2)
    A = torch.randn(r, r, generator=g, dtype=torch.float64)
    L_base = torch.linalg.cholesky(A @ A.T + 0.5 * torch.eye(r, dtype=torch.float64))
    Bmat = 0.3 * torch.randn(len(lam), q, r, generator=g, dtype=torch.float64)
    Sxb = Bmat @ L_base.T            # B = Sxb L^{-T}
    Sxx = T + Bmat @ Bmat.transpose(-1, -2)
    mean = torch.randn(len(lam), q, generator=g, dtype=torch.float64)
    Z = torch.randn(S, r + q, generator=g, dtype=torch.float64)
    best = torch.randn(S, generator=g, dtype=torch.float64) * 0.1
    acq, info, Bm, Cm = _run(mean, Sxx, Sxb, L_base, Z.T.contiguous(), best)
    assert (info & _lib.INFO_JITTER_MASK).tolist() == [0, 2, 4]
    assert float((Bm - Bmat).abs().max()) < 1e-12
    for i in range(len(lam)):
        # the kernel forms T = Sxx - B B^T itself; feed the oracle the same difference
        ref, Cf = _oracle_value(mean[i], Sxx[i] - Bmat[i] @ Bmat[i].T, Bmat[i], Z, best)
        lam_eff = lam[i] + (0.0 if lam[i] > 0 else 1e-8 * 10.0 ** ([0, 2, 4][i] - 1))
        tol = max(1e-9, 10.0 * torch.finfo(torch.float64).eps / lam_eff)
        assert abs(float(acq[i]) - float(ref)) <= tol * max(abs(float(ref)), 1.0)
        assert float((Cm[i] - Cf).abs().max()) <= max(tol, 1e-8) * float(Cf.abs().max())


def test_not_psd_in_the_cached_root_path_falls_back_to_joint_sampling(monkeypatch):
    """reference cached_cholesky.py:143-170: NanError / NotPSDError in the low-rank update -> BotorchWarning + standard
    sampling of the joint posterior.  The fused call is forced to fail; the result must equal the `cache_root=False` route."""
    from dataclasses import replace

    from botorch_b200.acquisition import logei, qLogNoisyExpectedImprovement
    from botorch_b200.benchmarks import configs
    from botorch_b200.exceptions.errors import NotPSDError
    from botorch_b200.exceptions.warnings import BotorchWarning
    from botorch_b200.sampling import SobolQMCNormalSampler

    spec = replace(configs.C2, S=64)
    data = configs.make_problem(spec, n=128)
    model = configs.build_model(data, DEV)
    X = configs.eval_points(data, 5).to(DEV)
    mk = lambda **kw: qLogNoisyExpectedImprovement(model, X_baseline=data.X_baseline.to(DEV), prune_baseline=False,  # noqa: E731
                                                   sampler=SobolQMCNormalSampler(torch.Size([64]), seed=3), **kw)
    acqf = mk()
    ok = acqf(X)

    def boom(*a, **k):
        raise NotPSDError("Matrix not positive definite after repeatedly adding jitter up to 1.0e-03.")

    monkeypatch.setattr(logei, "_chunked_fused", boom)
    with pytest.warns(BotorchWarning, match="Low-rank cholesky updates failed"):
        fb = acqf(X)
    assert acqf._cache_root is True   # restored after the fallback
    monkeypatch.undo()
    # the fallback samples the joint posterior over cat[X_baseline, X] with the base samples of the (r + q)-dim draw; the
    # cached-root value conditions on the same baseline draws: both are valid qLogNEI estimates that agree to MC accuracy
    assert torch.isfinite(fb).all() and fb.shape == ok.shape
    assert float((fb - ok).abs().max()) < 0.5
    # and it is exactly what the generic route returns with the cached root switched off on the same object
    acqf._cache_root = False
    try:
        direct = acqf._sample_reduction(acqf._q_reduction(acqf._non_reduced_forward(X=X)))
    finally:
        acqf._cache_root = True
    assert float((direct - fb).abs().max() / fb.abs().max()) < 1e-12
