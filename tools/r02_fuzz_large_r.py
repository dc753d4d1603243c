import os, sys, warnings
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
warnings.simplefilter("ignore")
import pytest, torch
import test_gpu_fuzz as F
bad = []
for seed in range(700, 1500):
    if seed % 11 != 5 or seed % 3 == 0: continue
    for flag in ("1", "0"):
        os.environ["MCACQ_BIGR_GEMM"] = flag
        try:
            F.test_random_configuration_matches_oracle(seed)
        except AssertionError as e:
            c = F._case(seed)
            msg = str(e).split("\n")[0][:60]
            import re
            m = re.search(r"assert ([0-9.e+-]+) < ([0-9.e+-]+)", str(e))
            print("FAIL seed", seed, "gemm" if flag == "1" else "loop", {k: c[k] for k in ("n","d","q","r","S","b","kernel","contraction","fixed_noise","normalize")}, m.group(0) if m else msg, flush=True)
        except BaseException as e:  # pytest.skip
            pass
print("done")
