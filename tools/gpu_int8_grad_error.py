"""Developer tool: gradient / value error of the int8 mode against the FP64 DMMA mode for a given number of backward
diagonals (env MCACQ_G_BWD, MCACQ_G_FWD), on C1 / C2 / C3 at random points."""
import os, sys
import torch
sys.path.insert(0, ".")
from botorch_b200.benchmarks import configs
from botorch_b200 import settings
dev = torch.device("cuda:0")
for cfg, b in (("C1", 256), ("C2", 256), ("C3", 512)):
    data = configs.make_problem(configs.CONFIGS[cfg])
    X = configs.eval_points(data, b).to(dev)
    res = {}
    for mode in ("dmma", "int8"):
        settings.contraction.set(mode)
        model = configs.build_model(data, dev); acqf = configs.build_acqf(data, model)
        strat = model.prediction_strategy()
        if mode == "int8":
            strat.desc.g_bwd = int(os.environ.get("MCACQ_G_BWD", "5")); strat.desc.g_fwd = int(os.environ.get("MCACQ_G_FWD", "6"))
        Xg = X.clone().requires_grad_(True); v = acqf(Xg); (g,) = torch.autograd.grad(v.sum(), Xg)
        res[mode] = (v.detach(), g)
    v0, g0 = res["dmma"]; v1, g1 = res["int8"]
    print(f"{cfg}: G fwd/bwd {os.environ.get('MCACQ_G_FWD','6')}/{os.environ.get('MCACQ_G_BWD','5')}: value rel err {float(((v1-v0).abs()/v0.abs()).max()):.2e}  grad err / max|grad| {float((g1-g0).abs().max()/g0.abs().max()):.2e}  per-batch worst {float(((g1-g0).abs().amax((1,2))/g0.abs().amax((1,2))).max()):.2e}")
settings.contraction.set("dmma")
