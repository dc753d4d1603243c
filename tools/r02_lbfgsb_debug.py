import sys
import numpy as np, torch
sys.path.insert(0, ".")
from botorch_b200.generation.device_gen import DeviceLBFGSB
from oracle.lbfgsb import FG, LbfgsbState
sys.path.insert(0, "tests")
from test_gpu_device_lbfgsb import CASES
DEV = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "quad6"
maxiter = int(sys.argv[2]) if len(sys.argv) > 2 else 6
fun, x0, lo, hi = CASES[name]
N, D = x0.shape
l, u = np.full(D, lo), np.full(D, hi)
states = [LbfgsbState(x0[i], l, u, maxiter=maxiter) for i in range(N)]
opt = DeviceLBFGSB(torch.from_numpy(x0).to(DEV), torch.from_numpy(l).to(DEV), torch.from_numpy(u).to(DEV), maxiter=maxiter)
for rnd in range(40):
    Xh = np.stack([s.x for s in states]); Xd = opt.X.cpu().numpy()
    diff = np.abs(Xd - Xh).max(axis=1)
    fdv, st = opt.summary()
    print(f"round {rnd}: max|dX| per problem {np.array2string(diff, precision=1)}")
    print("   cpu  (task,iter,nfev,col,stp):", [(s.task, s.iter, s.nfev, s.col, round(getattr(s, 'ls', {}).get('stp', 0), 6) if hasattr(s, 'ls') else None) for s in states])
    print("   dev  (task,msg,iter,nfev):    ", st.tolist())
    if not any(s.task == FG for s in states) and opt.active() == 0 and rnd > 0:
        break
    f, g = fun(Xh); fd, gd = fun(Xd)
    for i, s in enumerate(states):
        if s.task == FG: s.step(f[i], g[i])
    opt.step(torch.from_numpy(fd).to(DEV), torch.from_numpy(np.ascontiguousarray(gd)).to(DEV))
