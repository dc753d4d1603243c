"""CPU tests: the C-ABI library loads and exports every symbol include/mcacq_b200.h declares; host-side logic
(samplers, transforms, error conventions) behaves like the reference; the product fails loudly without CUDA."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def so_path():
    from botorch_b200 import _lib

    if not _lib._SO.exists():
        _lib.build()
    return str(_lib._SO)


def test_header_symbols_exported(so_path):
    header = open(os.path.join(ROOT, "include", "mcacq_b200.h")).read()
    declared = set(re.findall(r"\b(mcacq_[a-z_0-9]+)\s*\(", header))
    declared -= {"mcacq_model", "mcacq_baseline", "mcacq_mc"}
    lib = ctypes.CDLL(so_path)
    for sym in sorted(declared):
        assert hasattr(lib, sym), f"{sym} declared in the header but not exported"
    from botorch_b200 import _lib

    assert declared == set(_lib.EXPORTS)
    lib.mcacq_version.restype = ctypes.c_char_p
    assert b"sm_100a" in lib.mcacq_version()


def test_struct_layout_matches_header():
    from botorch_b200 import _lib

    assert ctypes.sizeof(_lib.Model) == 4 * 4 + 4 * 8 + 7 * 8 + 4 * 4 + 4 * 8
    assert ctypes.sizeof(_lib.Baseline) == 8 + 4 * 8
    # S, fat | tau_relu, tau_max | Zt, best | obj_weight, obj_offset, util_param | Zbar | n_con, con_fat | 3 x double[4] |
    # jitter_f32, pad
    assert ctypes.sizeof(_lib.MC) == 8 + 2 * 8 + 2 * 8 + 3 * 8 + 8 + 8 + 3 * 4 * 8 + 8


def test_bad_arguments_return_error_codes(so_path):
    """Argument validation happens before any CUDA call, so it is testable without a GPU."""
    from botorch_b200 import _lib

    L = _lib.lib()
    assert L.mcacq_dgemm_tri(0, 128, 100, None, None, None, None, None) == -1  # null pointers / np % 16 != 0
    assert L.mcacq_cov_cross(7, 1.0, None, 1, None, 1, 1, None, 1, None) == -1
    assert L.mcacq_workspace_bytes(4, 2, 3, 16, 0) > 0
    assert L.mcacq_workspace_bytes(-1, 2, 3, 16, 0) == 0
    assert L.mcacq_acq_forward(None, None, None, None, 1, 1, None, None, None, 0, None) == -1


def test_sass_uses_fp64_tensor_pipe(so_path):
    """The contraction must be on the FP64 tensor pipe: DMMA in SASS, cp.async (LDGSTS) staging."""
    out = subprocess.run(["cuobjdump", "-sass", so_path], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert out.stdout.count("DMMA") >= 128
    assert "LDGSTS" in out.stdout


def test_no_cpu_fallback():
    from botorch_b200 import _lib
    from botorch_b200.models import SingleTaskGP

    X = torch.rand(8, 2, dtype=torch.float64)
    Y = torch.rand(8, 1, dtype=torch.float64)
    model = SingleTaskGP(X, Y)
    with pytest.raises(_lib.McacqError):
        model.posterior(torch.rand(2, 2, dtype=torch.float64))


def test_product_never_imports_oracle():
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "botorch_b200")):
        for fn in fns:
            if fn.endswith(".py"):
                src = open(os.path.join(dp, fn)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad


def test_sampler_semantics():
    from botorch_b200.acquisition.logei import _ShapePosterior
    from botorch_b200.exceptions import InputDataError
    from botorch_b200.sampling import IIDNormalSampler, SobolQMCNormalSampler, get_sampler

    with pytest.raises(InputDataError):
        SobolQMCNormalSampler(sample_shape=4)
    post = _ShapePosterior(torch.Size([7]), 3, torch.device("cpu"), torch.float64)
    s = SobolQMCNormalSampler(torch.Size([16]), seed=1234)
    s._construct_base_samples(post)
    assert s.base_samples.shape == (16, 1, 3)  # t-batch collapsed (sampling/base.py:98-116)
    first = s.base_samples.clone()
    s._construct_base_samples(_ShapePosterior(torch.Size([2]), 3, torch.device("cpu"), torch.float64))
    assert torch.equal(first, s.base_samples)  # re-used across t-batch shapes
    s2 = SobolQMCNormalSampler(torch.Size([16]), seed=1234)
    s2._construct_base_samples(post)
    assert torch.equal(first, s2.base_samples)  # deterministic under the seed
    # frozen leading columns (sampling/normal.py:68-135)
    wide = SobolQMCNormalSampler(torch.Size([16]), seed=1234)
    wide._update_base_samples(_ShapePosterior(torch.Size([7]), 5, torch.device("cpu"), torch.float64), base_sampler=s)
    assert wide.base_samples.shape == (16, 1, 5)
    assert torch.equal(wide.base_samples[..., :3], first)
    iid = IIDNormalSampler(torch.Size([8]), seed=3)
    iid._construct_base_samples(post)
    assert iid.base_samples.shape == (8, 1, 3)
    assert isinstance(get_sampler(post, torch.Size([4]), seed=0), SobolQMCNormalSampler)


def test_t_batch_and_tau_errors():
    from botorch_b200.acquisition.logei import check_tau
    from botorch_b200.utils.transforms import t_batch_mode_transform

    class A:
        X_pending = None

        @t_batch_mode_transform()
        def f(self, X):
            return X.sum(dim=(-1, -2))

    a = A()
    assert a.f(torch.zeros(3, 2)).shape == (1,)
    assert a.f(torch.zeros(5, 3, 2)).shape == (5,)
    with pytest.raises(ValueError):
        a.f(torch.zeros(3))
    with pytest.raises(ValueError):
        check_tau(0.0, "tau")
    with pytest.raises(ValueError):
        check_tau(torch.ones(2), "tau")


def test_standardize_and_normalize_transforms():
    from botorch_b200.models.transforms import Normalize, Standardize

    Y = torch.tensor([[1.0], [2.0], [4.0]], dtype=torch.float64)
    st = Standardize(m=1)
    st.train()
    Yt, _ = st(Y)
    assert torch.allclose(Yt.mean(), torch.zeros((), dtype=torch.float64), atol=1e-15)
    assert torch.allclose(Yt.std(), torch.ones((), dtype=torch.float64))
    back, _ = st.untransform(Yt)
    assert torch.allclose(back, Y)
    nz = Normalize(d=2, bounds=torch.tensor([[0.0, -1.0], [2.0, 3.0]], dtype=torch.float64))
    X = torch.tensor([[1.0, 1.0]], dtype=torch.float64)
    assert torch.allclose(nz(X), torch.tensor([[0.5, 0.5]], dtype=torch.float64))
    learn = Normalize(d=2)
    learn.train()
    out = learn(torch.tensor([[1.0, 5.0], [3.0, 5.0]], dtype=torch.float64))
    assert torch.allclose(out[:, 0], torch.tensor([0.0, 1.0], dtype=torch.float64))
    assert torch.allclose(out[:, 1], torch.tensor([5.0, 5.0], dtype=torch.float64))  # degenerate range untouched


def test_fused_supported_is_a_host_only_query():
    """`mcacq_fused_supported` (compiled limits + shared memory of the sample / reduce kernels) answers without a CUDA device."""
    from botorch_b200 import _lib

    assert _lib.fused_supported(8, 16, 1024) and _lib.fused_supported(8, 0, 512) and _lib.fused_supported(8, 512, 1024)
    assert not _lib.fused_supported(8, _lib.MAX_R + 1, 1024)
    assert not _lib.fused_supported(_lib.MAX_Q + 1, 16, 1024)
    assert not _lib.fused_supported(32, 512, 1024)            # limits fine, q x r too large for the kernels' shared memory
    assert not _lib.fused_supported(8, 16, 0) and not _lib.fused_supported(0, 16, 8)
