"""The reference's OWN known-answer tests for the MC acquisition arithmetic, restated against this package.

The reference pins `_sample_forward` / the q- and sample reductions with a `MockPosterior` that returns prescribed samples
(test/acquisition/test_logei.py:100-455, test/acquisition/test_monte_carlo.py:98-150, 388-453, 470-495, 570-584): the
expected numbers below are the ones those tests assert.  Two legs:

* CPU (`not gpu`): the Python mirror's classes (same names / constructor arguments) over `botorch_b200.utils.testing.MockModel`
  -- host semantics of the generic route;
* GPU: the same expectations through the C ABI (`mcacq_sample_reduce_forward`, the kernel the fused route runs), with zero
  base samples so that the kernel's reparameterised samples ARE the prescribed ones (`y = mean + C * 0`).
"""
import ctypes as C
import math

import pytest
import torch

TAU_RELU = 1e-6


def _mock(samples):
    from botorch_b200.utils.testing import MockModel, MockPosterior

    return MockModel(MockPosterior(samples=samples))


# ---------------------------------------------------------------------------------------------------------------------
# CPU leg: reference test bodies against the mirror classes
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float, torch.double])
def test_q_log_expected_improvement(dtype):
    """test/acquisition/test_logei.py:101-207"""
    from botorch_b200.acquisition import qExpectedImprovement, qLogExpectedImprovement
    from botorch_b200.sampling import IIDNormalSampler, SobolQMCNormalSampler

    mm = _mock(torch.zeros(1, 1, 1, dtype=dtype))
    X = torch.zeros(1, 1, dtype=dtype)
    sampler = IIDNormalSampler(sample_shape=torch.Size([2]))
    acqf = qExpectedImprovement(model=mm, best_f=0, sampler=sampler)
    log_acqf = qLogExpectedImprovement(model=mm, best_f=0, sampler=sampler)
    assert not acqf._fat and log_acqf._fat
    res = acqf(X)
    assert res.dtype == dtype and res.item() == 0.0
    exp_log_res = log_acqf(X).exp().item()
    assert 0 < exp_log_res <= log_acqf.tau_relu
    # best_f shifted downward: non-zero improvement
    acqf = qExpectedImprovement(model=mm, best_f=-1, sampler=sampler)
    log_acqf = qLogExpectedImprovement(model=mm, best_f=-1, sampler=sampler)
    res, exp_log_res = acqf(X), log_acqf(X).exp()
    assert res.item() == 1.0 and exp_log_res.dtype == dtype
    assert 1.0 <= exp_log_res.item() <= 1.0 + log_acqf.tau_relu
    # best_f shifted upward: EI is exactly 0, LogEI stays finite
    acqf = qExpectedImprovement(model=mm, best_f=1, sampler=sampler)
    log_acqf = qLogExpectedImprovement(model=mm, best_f=1, sampler=sampler)
    res, log_res = acqf(X), log_acqf(X)
    assert res.item() == 0
    assert 0 <= log_res.exp().item() <= log_acqf.tau_relu
    assert -100 < log_res.item() < -1
    for sampler in (IIDNormalSampler(sample_shape=torch.Size([2]), seed=12345), SobolQMCNormalSampler(sample_shape=torch.Size([2]))):
        res = qLogExpectedImprovement(model=mm, best_f=0, sampler=sampler)(X)
        assert 0 < res.exp().item() < TAU_RELU
    with pytest.raises(ValueError, match="tau_max is not a scalar:"):
        qLogExpectedImprovement(model=mm, best_f=0, tau_max=torch.tensor([1, 2]))
    with pytest.raises(ValueError, match="tau_relu is non-positive:"):
        qLogExpectedImprovement(model=mm, best_f=0, tau_relu=-2)


@pytest.mark.parametrize("dtype", [torch.float, torch.double])
def test_q_log_expected_improvement_batch(dtype):
    """test/acquisition/test_logei.py:209-250 and test_monte_carlo.py:118-148"""
    from botorch_b200.acquisition import qExpectedImprovement, qLogExpectedImprovement
    from botorch_b200.sampling import IIDNormalSampler

    samples = torch.zeros(2, 2, 1, dtype=dtype)
    samples[0, 0, 0] = 1.0
    mm = _mock(samples)
    X = torch.zeros(2, 2, 1, dtype=dtype)
    for S, best_f in ((2, 0), (3, torch.Tensor([0, 0]))):
        sampler = IIDNormalSampler(sample_shape=torch.Size([S]))
        res = qLogExpectedImprovement(model=mm, best_f=best_f, sampler=sampler)(X).exp()
        assert res.dtype == dtype
        assert 1.0 <= res[0].item() <= 1.0 + TAU_RELU
        assert 0 < res[1].item() <= TAU_RELU
        res = qExpectedImprovement(model=mm, best_f=best_f, sampler=sampler)(X)
        assert res[0].item() == 1.0 and res[1].item() == 0.0
    res = qExpectedImprovement(model=mm, best_f=-1, sampler=sampler)(X)
    assert res[0].item() == 2.0 and res[1].item() == 1.0


@pytest.mark.parametrize("dtype", [torch.float, torch.double])
def test_q_log_noisy_expected_improvement(dtype):
    """test/acquisition/test_logei.py:306-383 (cache_root = False: a mock posterior has no covariance to factor)"""
    from botorch_b200.acquisition import qLogNoisyExpectedImprovement, qNoisyExpectedImprovement
    from botorch_b200.sampling import IIDNormalSampler

    mm_noisy = _mock(torch.tensor([0.0, 1.0], dtype=dtype).view(1, 2, 1))
    X_baseline = torch.zeros(1, 1, dtype=dtype)
    X = torch.zeros(1, 1, dtype=dtype)
    sampler = IIDNormalSampler(sample_shape=torch.Size([2]))
    kwargs = dict(model=mm_noisy, X_baseline=X_baseline, sampler=sampler, prune_baseline=False, cache_root=False)
    assert qNoisyExpectedImprovement(**kwargs)(X).item() == 1.0
    log_res = qLogNoisyExpectedImprovement(**kwargs)(X)
    assert log_res.dtype == dtype
    assert math.isclose(log_res.exp().item(), 1.0, rel_tol=1e-5)
    # a pending point is equivalent to one more baseline point (incremental = True)
    mm_pending = _mock(torch.tensor([1.0, 0.0, 0.0], dtype=dtype).view(1, 3, 1))
    X2 = torch.zeros(1, 1, 1, dtype=dtype, requires_grad=True)
    kw = dict(model=mm_pending, sampler=sampler, prune_baseline=False, cache_root=False)
    log_acqf = qLogNoisyExpectedImprovement(X_baseline=X_baseline, incremental=True, **kw)
    log_acqf.set_X_pending(X)
    assert log_acqf.X_pending is None
    assert torch.equal(log_acqf.X_baseline, torch.cat([X_baseline, X], dim=0))
    af_val1 = log_acqf(X2)
    log_acqf2 = qLogNoisyExpectedImprovement(X_baseline=torch.cat([X_baseline, X], dim=-2), incremental=False, **kw)
    af_val2 = log_acqf2(X2)
    assert math.isclose(af_val1.item(), af_val2.item(), rel_tol=1e-5)
    log_acqf.set_X_pending(None)
    assert torch.equal(log_acqf.X_baseline, X_baseline)


@pytest.mark.parametrize("dtype", [torch.float, torch.double])
def test_q_log_noisy_expected_improvement_batch(dtype):
    """test/acquisition/test_logei.py:385-455"""
    from botorch_b200.acquisition import qLogNoisyExpectedImprovement
    from botorch_b200.sampling import IIDNormalSampler, SobolQMCNormalSampler

    samples_noisy = torch.zeros(2, 3, 1, dtype=dtype)
    samples_noisy[0, -1, 0] = 1.0
    mm_noisy = _mock(samples_noisy)
    X = torch.zeros(2, 2, 1, dtype=dtype)
    X_baseline = torch.zeros(1, 1, dtype=dtype)
    for sampler in (IIDNormalSampler(sample_shape=torch.Size([2])), IIDNormalSampler(sample_shape=torch.Size([2]), seed=12345),
                    SobolQMCNormalSampler(sample_shape=torch.Size([2]))):
        acqf = qLogNoisyExpectedImprovement(model=mm_noisy, X_baseline=X_baseline, sampler=sampler, prune_baseline=False,
                                            cache_root=False)
        res = acqf(X).exp()
        assert torch.allclose(res, torch.tensor([1.0, 0.0], dtype=dtype), atol=acqf.tau_relu, rtol=0)
        assert 0.0 < res[1].item() < acqf.tau_relu
        # the base samples keep the t-batch dimension collapsed and are not redrawn
        assert acqf.sampler.base_samples.shape == torch.Size([2, 1, 3, 1])
        bs = acqf.sampler.base_samples.clone()
        acqf(X)
        assert torch.equal(acqf.sampler.base_samples, bs)


@pytest.mark.parametrize("dtype", [torch.float, torch.double])
def test_other_sample_reducing_utilities(dtype):
    """test/acquisition/test_monte_carlo.py:388-453, 470-495, 570-584"""
    from botorch_b200.acquisition import (qLowerConfidenceBound, qPosteriorStandardDeviation, qProbabilityOfImprovement,
                                          qSimpleRegret, qUpperConfidenceBound)
    from botorch_b200.sampling import IIDNormalSampler

    one = _mock(torch.zeros(1, 1, 1, dtype=dtype))
    X1 = torch.zeros(1, 1, dtype=dtype)
    samples = torch.zeros(2, 2, 1, dtype=dtype)
    samples[0, 0, 0] = 1.0
    two = _mock(samples)
    X2 = torch.zeros(2, 2, 1, dtype=dtype)
    sampler = IIDNormalSampler(sample_shape=torch.Size([2]))
    assert qProbabilityOfImprovement(model=one, best_f=0, sampler=sampler)(X1).item() == 0.5
    res = qProbabilityOfImprovement(model=two, best_f=0, sampler=sampler)(X2)
    assert res[0].item() == 1.0 and res[1].item() == 0.5
    assert qSimpleRegret(model=one, sampler=sampler)(X1).item() == 0.0
    res = qSimpleRegret(model=two, sampler=sampler)(X2)
    assert res[0].item() == 1.0 and res[1].item() == 0.0
    for cls in (qUpperConfidenceBound, qLowerConfidenceBound):
        assert cls(model=one, beta=0.5, sampler=sampler)(X1).item() == 0.0
        res = cls(model=two, beta=0.5, sampler=sampler)(X2)
        assert res[0].item() == 1.0 and res[1].item() == 0.0
    res = qPosteriorStandardDeviation(model=two, sampler=IIDNormalSampler(sample_shape=torch.Size([8])))(X2)
    assert res[0].item() == 0.0 and res[1].item() == 0.0


# ---------------------------------------------------------------------------------------------------------------------
# GPU leg: the same expectations through the sample / reduce kernel of the fused route (C ABI)
# ---------------------------------------------------------------------------------------------------------------------
def _kernel_value(samples, best, mode, S=2, util_param=0.0, tau_relu=TAU_RELU):
    """acq[b] of `mcacq_sample_reduce_forward` for prescribed samples [b x q] (identity covariance, ZERO base samples)."""
    from botorch_b200 import _lib

    dev = torch.device("cuda:0")
    f64 = dict(device=dev, dtype=torch.float64)
    L = _lib.lib()
    b, q = samples.shape
    mean = samples.to(**f64).contiguous()
    Sxx = torch.eye(q, **f64).expand(b, q, q).contiguous()
    Zt = torch.zeros(q, S, **f64)
    best_d = torch.as_tensor(best, dtype=torch.float64).expand(S).to(dev).contiguous()
    Zbar = torch.zeros(q, **f64)
    acq = torch.empty(b, **f64)
    info = torch.empty(b, dtype=torch.int32, device=dev)
    Cm = torch.empty(b, q, q, **f64)
    mc = _lib.MC(S=S, fat=mode, tau_relu=tau_relu, tau_max=1e-2, Zt=Zt.data_ptr(), best=best_d.data_ptr(), obj_weight=1.0,
                 obj_offset=0.0, util_param=util_param, Zbar=Zbar.data_ptr(), n_con=0, con_fat=0, jitter_f32=0)
    rc = L.mcacq_sample_reduce_forward(None, C.byref(mc), mean.data_ptr(), Sxx.data_ptr(), None, b, q, acq.data_ptr(),
                                       info.data_ptr(), None, Cm.data_ptr(), _lib.stream_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    assert int(info.max()) & _lib.INFO_FLAG_MASK == 0
    return acq.cpu()


@pytest.mark.gpu
def test_kernel_reproduces_the_reference_known_answers():
    one = torch.zeros(1, 1)
    two = torch.zeros(2, 2)
    two[0, 0] = 1.0
    # qLogEI (mode 1), test_logei.py:101-250
    v = _kernel_value(one, 0.0, 1).exp().item()
    assert 0 < v <= TAU_RELU
    v = _kernel_value(one, -1.0, 1).exp().item()
    assert 1.0 <= v <= 1.0 + TAU_RELU
    lv = _kernel_value(one, 1.0, 1).item()
    assert -100 < lv < -1 and 0 <= math.exp(lv) <= TAU_RELU
    for S in (2, 3):
        v = _kernel_value(two, 0.0, 1, S=S).exp()
        assert 1.0 <= v[0].item() <= 1.0 + TAU_RELU and 0 < v[1].item() <= TAU_RELU
    # the kernel value IS the reference arithmetic: log_fatplus -> fatmax -> logmeanexp of the prescribed samples
    from oracle.safe_math import fatmax, log_fatplus, logmeanexp

    want = logmeanexp(fatmax(log_fatplus(two.double().expand(3, 2, 2) - 0.25, tau=TAU_RELU), dim=-1, tau=1e-2), dim=0)
    got = _kernel_value(two, 0.25, 1, S=3)
    assert float(((got - want).abs() / want.abs()).max()) < 1e-13
    # qEI (mode 2), test_monte_carlo.py:98-148
    assert _kernel_value(one, 0.0, 2).item() == 0.0 and _kernel_value(one, -1.0, 2).item() == 1.0
    assert _kernel_value(two, 0.0, 2, S=3).tolist() == [1.0, 0.0]
    assert _kernel_value(two, -1.0, 2, S=3).tolist() == [2.0, 1.0]
    # qSimpleRegret (mode 3), :429-453; qProbabilityOfImprovement (mode 4, tau = 1e-3), :388-423
    assert _kernel_value(one, 0.0, 3).item() == 0.0 and _kernel_value(two, 0.0, 3).tolist() == [1.0, 0.0]
    assert _kernel_value(one, 0.0, 4, tau_relu=1e-3).item() == 0.5
    assert _kernel_value(two, 0.0, 4, tau_relu=1e-3).tolist() == [1.0, 0.5]
    # qUCB (mode 5, beta' = sqrt(beta pi / 2)) and qPSTD (mode 6), :470-495, 570-584: zero-variance samples
    bp = math.sqrt(0.5 * math.pi / 2)
    assert _kernel_value(one, 0.0, 5, util_param=bp).item() == 0.0
    assert _kernel_value(two, 0.0, 5, util_param=bp).tolist() == [1.0, 0.0]
    assert _kernel_value(two, 0.0, 6, S=8, util_param=math.sqrt(math.pi / 2)).tolist() == [0.0, 0.0]
