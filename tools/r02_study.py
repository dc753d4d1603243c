"""Developer tool (round 2, run under gpurun): accuracy of the contraction modes AT and NEAR the training points and
timing of the int8 contraction variants.  Writes gpurun_out/r02_study_*.txt.

  python tools/r02_study.py accuracy   # dmma / int8 (6,5) / int8 (7,6) / int8 auto against the CPU oracle, C1-C3 full n
  python tools/r02_study.py contract   # ozaki_imma_kernel: G, debug switches (no-MMA / no-TMA / one box), stages, BN
"""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from botorch_b200 import _lib, settings  # noqa: E402
from botorch_b200.benchmarks import configs  # noqa: E402

dev = torch.device("cuda:0")
OUT = "gpurun_out"
os.makedirs(OUT, exist_ok=True)


def near_train_points(data, per=64, deltas=(0.0, 1e-6, 1e-3), seed=0):
    g = torch.Generator().manual_seed(seed)
    n, d = data.train_X.shape
    idx = torch.linspace(0, n - 1, min(n, per)).round().long()
    T = data.train_X[idx]
    dirs = torch.randn(T.shape, generator=g, dtype=torch.float64)
    dirs = dirs / dirs.norm(dim=-1, keepdim=True)
    return {dl: (T + dl * dirs).clamp(0.0, 1.0) for dl in deltas}


def accuracy():
    from oracle.acquisition import value_and_grad
    from oracle.harness import build_oracle

    lines = []
    for name in ("C1", "C2", "C3"):
        spec = configs.CONFIGS[name]
        data = configs.make_problem(spec)
        t0 = time.time()
        orc = build_oracle(data)
        pts = near_train_points(data)
        box = configs.eval_points(data, 16).reshape(-1, spec.d)
        sets = {f"train+{dl:g}": p for dl, p in pts.items()}
        sets["sobol box"] = box
        ref = {}
        for k, P in sets.items():
            m, c = orc.gp.posterior_mvn(P.unsqueeze(1))
            ref[k] = (m.reshape(-1), c.reshape(-1))
        # q-batches made of (displaced) training points: the acquisition value where the variance has collapsed
        q = spec.q
        qb = {}
        for dl, p in pts.items():
            nb = p.shape[0] // q
            qb[f"train+{dl:g}"] = p[: nb * q].reshape(nb, q, spec.d)[:8]
        qb["sobol box"] = configs.eval_points(data, 8)
        qref = {k: value_and_grad(orc, X) for k, X in qb.items()}
        lines.append(f"== {name}: n={spec.n} d={spec.d} q={spec.q} {spec.kernel} (oracle built in {time.time() - t0:.1f}s)")
        for mode, forced in (("dmma", None), ("int8", (6, 5)), ("int8", (7, 6)), ("int8", (7, 5)), ("int8", None)):
            with settings.contraction(mode), settings.int8_slices(forced):
                model = configs.build_model(data, dev)
                strat = model.prediction_strategy()
                acqf = configs.build_acqf(data, model)
                tag = f"{mode}{'' if forced is None else forced}"
                if mode == "int8" and forced is None:
                    tag += f" auto->{strat.contraction}({strat.g_fwd},{strat.g_bwd}) probe v={strat.int8_probe_error:.1e} g={strat.int8_probe_grad_error:.1e}"
                for k, P in sets.items():
                    post = model.posterior(P.unsqueeze(1).to(dev))
                    m = post.mean.reshape(-1).cpu()
                    v = post.variance.reshape(-1).cpu()
                    mr, vr = ref[k]
                    em = float(((m - mr).abs() / mr.abs().clamp_min(1e-300)).max())
                    ev = float(((v - vr).abs() / vr.abs()).max())
                    lines.append(f"  {tag:70s} {k:12s} mean rel {em:.2e}  var rel max {ev:.2e}  (var min {float(vr.min()):.2e})")
                for k, X in qb.items():
                    Xg = X.to(dev).requires_grad_(True)
                    val = acqf(Xg)
                    (gr,) = torch.autograd.grad(val.sum(), Xg)
                    vo, go = qref[k]
                    ea = float(((val.detach().cpu() - vo).abs() / vo.abs()).max())
                    eg = float((gr.cpu() - go).abs().max() / go.abs().max())
                    lines.append(f"  {tag:70s} {k:12s} acq rel max {ea:.2e}  grad rel {eg:.2e}")
            del model, acqf, strat
            torch.cuda.empty_cache()
        print("\n".join(lines[-60:]), flush=True)
    open(f"{OUT}/r02_study_accuracy.txt", "w").write("\n".join(lines) + "\n")


def contract():
    L, st, f64 = _lib.lib(), _lib.stream_ptr(), dict(device=dev, dtype=torch.float64)
    M, n = 65536, 4096
    A = torch.rand(M, n, **f64)
    R = torch.triu(torch.randn(n, n, **f64))
    Rt = R.t().contiguous()
    ref = A[:128] @ R
    lines = []

    def run(G, env, tri=0):
        for k in ("MCACQ_OZ_DEBUG", "MCACQ_OZ_STAGES", "MCACQ_OZ_BN", "MCACQ_OZ_BK", "MCACQ_OZ_GROUP", "MCACQ_OZ_HINT", "MCACQ_OZ_TMASTORE", "MCACQ_OZ_RING", "MCACQ_OZ_NA"):
            os.environ.pop(k, None)
        os.environ.update({k: str(v) for k, v in env.items()})
        As = torch.empty(G, M, n, dtype=torch.int8, device=dev)
        ra = torch.empty(M, **f64)
        Bs = torch.empty(G, n, n, dtype=torch.int8, device=dev)
        cb = torch.empty(n, **f64)
        L.mcacq_slice_rows(A.data_ptr(), M, n, n, n, G, 1, 0, As.data_ptr(), ra.data_ptr(), st)
        L.mcacq_slice_rows((Rt if tri == 0 else R).data_ptr(), n, n, n, n, G, 0, 0, Bs.data_ptr(), cb.data_ptr(), st)
        C = torch.zeros(M, n, **f64)
        fn = lambda: L.mcacq_ozaki_contract(tri, M, n, n, G, As.data_ptr(), ra.data_ptr(), Bs.data_ptr(), cb.data_ptr(), C.data_ptr(), n, st)  # noqa: E731
        rc = fn()
        torch.cuda.synchronize()
        if rc != 0:
            lines.append(f"G={G} {env}: rc={rc}")
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        err = float((C[:128] - ref).abs().max() / ref.abs().max()) if tri == 0 else float("nan")
        pairs = G * (G + 1) // 2
        lines.append(f"G={G} tri={tri} {str(env):60s} {ms:8.3f} ms   int8 {M * n * (n + 64.0) * pairs / ms * 1e-12:5.2f} POP/s   err {err:.1e}")
        print(lines[-1], flush=True)
        del As, Bs, C

    which = sys.argv[2] if len(sys.argv) > 2 else "full"
    if which == "ring":
        for G in (4, 5, 6, 7):
            run(G, {})
            run(G, {"MCACQ_OZ_RING": 0})
            run(G, {"MCACQ_OZ_DEBUG": 1})
            run(G, {"MCACQ_OZ_DEBUG": 2})
            run(G, {"MCACQ_OZ_NA": 8})
            run(G, {"MCACQ_OZ_NA": 12})
            run(G, {"MCACQ_OZ_HINT": 0})
            run(G, {"MCACQ_OZ_GROUP": 4})
            run(G, {"MCACQ_OZ_GROUP": 16})
        for G in (5, 6):
            run(G, {}, tri=1)
            run(G, {"MCACQ_OZ_RING": 0}, tri=1)
        open(f"{OUT}/r02_study_contract_{which}.txt", "w").write("\n".join(lines) + "\n")
        return
    if which == "epilogue":
        for G in (5, 6, 7):
            run(G, {})
            run(G, {"MCACQ_OZ_TMASTORE": 0})
            run(G, {"MCACQ_OZ_DEBUG": 2})
            run(G, {"MCACQ_OZ_DEBUG": 3})
        run(6, {}, tri=1)
        run(5, {}, tri=1)
        # the tensor core's shared-memory operand rate: pure N = 256 instructions (G = 1, BN = 256), MMA only
        for G in (1, 2):
            run(G, {"MCACQ_OZ_BK": 64})
            run(G, {"MCACQ_OZ_BK": 64, "MCACQ_OZ_DEBUG": 2})
            run(G, {"MCACQ_OZ_BK": 64, "MCACQ_OZ_DEBUG": 3})
            run(G, {"MCACQ_OZ_BK": 128})
            run(G, {"MCACQ_OZ_BK": 128, "MCACQ_OZ_DEBUG": 2})
    else:
        for G in (5, 6, 7):
            run(G, {})
            run(G, {"MCACQ_OZ_DEBUG": 1})   # no UMMA: TMA ingest alone
            run(G, {"MCACQ_OZ_DEBUG": 2})   # no TMA: UMMA + epilogue alone
            run(G, {"MCACQ_OZ_DEBUG": 3})   # neither: barrier/epilogue skeleton
            run(G, {"MCACQ_OZ_DEBUG": 4})   # one TMA box per operand over all G slices
            run(G, {"MCACQ_OZ_STAGES": 1})
            run(G, {"MCACQ_OZ_BN": 64})
            run(G, {"MCACQ_OZ_HINT": 0})
        run(6, {}, tri=1)
        run(5, {}, tri=1)
    open(f"{OUT}/r02_study_contract_{which}.txt", "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    {"accuracy": accuracy, "contract": contract}[sys.argv[1]]()
