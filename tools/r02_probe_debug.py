import sys, torch
sys.path.insert(0, ".")
from botorch_b200.benchmarks import configs
dev = torch.device("cuda:0")
for cfg in ("C2", "C3"):
    data = configs.make_problem(configs.CONFIGS[cfg]); model = configs.build_model(data, dev)
    st = model.prediction_strategy(); print(cfg, st.contraction, st.g_fwd, st.g_bwd, st.int8_var_ratio_limit)
