"""CPU restatement of the device-resident L-BFGS-B step (TEST INFRASTRUCTURE ONLY).

The reference optimises every restart with scipy's L-BFGS-B (`scipy.optimize._lbfgsb.setulb`, the C translation of
L-BFGS-B 3.0 by Zhu, Byrd, Lu, Nocedal and Morales) through the batched driver botorch/optim/batched_lbfgs_b.py:365-634;
`fmin_l_bfgs_b_batched` there steps one state machine per restart and evaluates all active restarts at once.  scipy is a
third-party dependency of the reference (pyproject.toml: scipy), not part of its tree, so this module restates the
PUBLISHED algorithm -- R. H. Byrd, P. Lu, J. Nocedal, C. Zhu, "A limited memory algorithm for bound constrained
optimization", SIAM J. Sci. Comput. 16 (1995): generalised Cauchy point (Algorithm CP), direct primal subspace
minimisation (section 5.1) with the projection step of J. L. Morales, J. Nocedal, "Remark on Algorithm 778" (2011), the
More'-Thuente line search (MINPACK-2 dcsrch / dcstep, ftol 1e-3, gtol 0.9, xtol 0.1), the compact limited-memory matrices
(theta, S, Y, S'S, S'Y) and L-BFGS-B's stopping tests (projected gradient <= pgtol, relative reduction <= factr * epsmch) --
in the exact order of operations the CUDA kernel `csrc/lbfgsb.cu` uses, with dense (2 col x 2 col) solves in place of
L-BFGS-B's incremental factorisations (same mathematics).  It is pinned against scipy itself
(tests/test_lbfgsb_device_model.py: same minimisers, same iteration counts to within the tolerance-based contract of
SURVEY.md section 8f N4) and is the checker of the CUDA kernel on a GPU (tests/test_gpu_device_lbfgsb.py).

One `LbfgsbState` per problem; reverse communication like `setulb`:
    st = LbfgsbState(x0, lower, upper); while st.task == FG: f, g = fun(st.x); st.step(f, g)
"""
from __future__ import annotations

import numpy as np

FG, NEW_X, CONVERGED, STOPPED, ABNORMAL = 0, 1, 2, 3, 4
EPSMCH = np.finfo(np.float64).eps
BIG = 1.0e10
FTOL, GTOL, XTOL = 1.0e-3, 0.9, 0.1


class LbfgsbState:
    def __init__(self, x0, lower, upper, m: int = 10, factr: float = 1e7, pgtol: float = 1e-5, maxiter: int = 15000,
                 maxfun: int = 15000, maxls: int = 20):
        n = x0.shape[0]
        self.n, self.m = n, m
        self.l = np.asarray(lower, dtype=np.float64).copy()
        self.u = np.asarray(upper, dtype=np.float64).copy()
        self.x = np.clip(np.asarray(x0, dtype=np.float64), self.l, self.u)
        self.factr, self.pgtol, self.maxiter, self.maxfun, self.maxls = factr, pgtol, maxiter, maxfun, maxls
        self.f = 0.0
        self.g = np.zeros(n)
        self.S = np.zeros((m, n))   # rows = correction pairs, oldest first among the first `col`
        self.Y = np.zeros((m, n))
        self.SS = np.zeros((m, m))
        self.SY = np.zeros((m, m))  # SY[i][j] = s_i . y_j
        self.col = 0
        self.theta = 1.0
        self.iter = 0
        self.nfev = 0
        self.task = FG
        self.phase = "start"
        self.message = ""
        self.boxed = bool(np.all(np.isfinite(self.l)) and np.all(np.isfinite(self.u)))

    # ------------------------------------------------------------------ helpers
    def _projgr(self) -> float:
        g, x = self.g, self.x
        pg = np.where(g < 0, np.maximum(x - self.u, g), np.minimum(x - self.l, g))
        return float(np.abs(pg).max()) if pg.size else 0.0

    def _minv(self):
        """M^{-1} = [[-D, L^T], [L, theta S^T S]] of the compact representation B = theta I - W M W^T, W = [Y, theta S]."""
        c = self.col
        SY = self.SY[:c, :c]
        D = np.diag(np.diag(SY))
        L = np.tril(SY, -1)
        return np.block([[-D, L.T], [L, self.theta * self.SS[:c, :c]]])

    @staticmethod
    def _solve(A, b):
        """Gaussian elimination with partial pivoting (the kernel's dense solve); returns None on a zero pivot."""
        A = A.copy()
        b = b.copy()
        k = A.shape[0]
        for i in range(k):
            p = i + int(np.argmax(np.abs(A[i:, i])))
            if not (abs(A[p, i]) > 0.0):
                return None
            if p != i:
                A[[i, p]] = A[[p, i]]
                b[[i, p]] = b[[p, i]]
            for r in range(i + 1, k):
                fct = A[r, i] / A[i, i]
                A[r, i:] -= fct * A[i, i:]
                b[r] -= fct * b[i]
        xs = np.zeros(k)
        for i in range(k - 1, -1, -1):
            xs[i] = (b[i] - A[i, i + 1:] @ xs[i + 1:]) / A[i, i]
        return xs

    def _reset_memory(self):
        self.col = 0
        self.theta = 1.0

    # ------------------------------------------------------------------ generalised Cauchy point
    def _cauchy(self):
        n, c = self.n, self.col
        x, g, l, u, theta = self.x, self.g, self.l, self.u, self.theta
        xcp = x.copy()
        tl, tu = x - l, u - x
        neg = -g
        fixed = ((tl <= 0) & (neg <= 0)) | ((tu <= 0) & (neg >= 0) & ~(tl <= 0)) | (neg == 0)
        d = np.where(fixed, 0.0, neg)
        tbreak = np.full(n, np.inf)
        lo = (~fixed) & (neg < 0) & np.isfinite(l)
        up = (~fixed) & (neg > 0) & np.isfinite(u)
        tbreak[lo] = tl[lo] / (-neg[lo])
        tbreak[up] = tu[up] / neg[up]
        W = np.concatenate([self.Y[:c], theta * self.S[:c]], axis=0) if c > 0 else np.zeros((0, n))  # 2c x n
        p = W @ d
        cvec = np.zeros(2 * c)
        f1 = -float(d @ d)
        f2 = -theta * f1
        Minv = self._minv() if c > 0 else None
        if c > 0:
            v = self._solve(Minv, p)
            if v is None:
                return None
            f2 -= float(p @ v)
        f2_org = f2
        dtm = -f1 / f2
        tsum = 0.0
        order = np.argsort(tbreak, kind="stable")
        nbreak = int(np.isfinite(tbreak).sum())
        tj0 = 0.0
        free_mask = ~fixed
        all_fixed_exit = False
        for k in range(nbreak):
            b = order[k]
            tj = tbreak[b]
            dt = tj - tj0
            if dtm < dt:
                break
            tsum += dt
            dibp = d[b]
            d[b] = 0.0
            if dibp > 0:
                zibp = u[b] - x[b]
                xcp[b] = u[b]
            else:
                zibp = l[b] - x[b]
                xcp[b] = l[b]
            free_mask[b] = False
            if k == nbreak - 1 and nbreak == n:
                dtm = dt
                all_fixed_exit = True
                break
            dibp2 = dibp * dibp
            f1 = f1 + dt * f2 + dibp2 - theta * dibp * zibp
            f2 = f2 - theta * dibp2
            if c > 0:
                cvec += dt * p
                wbp = W[:, b]
                v = self._solve(Minv, wbp)
                if v is None:
                    return None
                wmc, wmp, wmw = float(cvec @ v), float(p @ v), float(wbp @ v)
                p = p - dibp * wbp
                f1 += dibp * wmc
                f2 += 2.0 * dibp * wmp - dibp2 * wmw
            f2 = max(EPSMCH * f2_org, f2)
            if k < nbreak - 1:
                dtm = -f1 / f2
            elif nbreak == n:  # unreachable (handled above); kept for symmetry with the published listing
                f1 = f2 = dtm = 0.0
            else:
                dtm = -f1 / f2
            tj0 = tj
        if not all_fixed_exit:
            dtm = max(dtm, 0.0)
            tsum += dtm
            xcp = np.where(d != 0.0, x + tsum * d, xcp)
        if c > 0:
            cvec += dtm * p
        # variables free at the Cauchy point: not fixed at entry, breakpoint not reached
        return xcp, cvec, free_mask

    # ------------------------------------------------------------------ subspace minimisation
    def _subsm(self, xcp, cvec, free_mask):
        c, theta = self.col, self.theta
        x, g, l, u = self.x, self.g, self.l, self.u
        F = np.nonzero(free_mask)[0]
        if F.size == 0 or c == 0:
            return xcp
        W = np.concatenate([self.Y[:c], theta * self.S[:c]], axis=0)
        Minv = self._minv()
        mc = self._solve(Minv, cvec)
        if mc is None:
            return None
        r = -theta * (xcp[F] - x[F]) - g[F] + W[:, F].T @ mc
        Wz = W[:, F]
        K3 = Minv - (Wz @ Wz.T) / theta
        v = self._solve(K3, Wz @ r)
        if v is None:
            return None
        dF = (r + (Wz.T @ v) / theta) / theta
        xbar = xcp.copy()
        proj = np.clip(xcp[F] + dF, l[F], u[F])
        xbar[F] = proj
        hit = bool(np.any((proj == l[F]) | (proj == u[F])))
        if hit:
            dd_p = float((xbar - x) @ g)
            if dd_p > 0.0:
                # the projected point is not a descent direction: truncate the Newton step at the first bound instead
                xbar = xcp.copy()
                alpha, ibd, at_upper = 1.0, -1, False
                for i, k in enumerate(F):
                    dk = dF[i]
                    if dk < 0 and np.isfinite(l[k]):
                        t2 = l[k] - xcp[k]
                        t1 = 0.0 if t2 >= 0 else (t2 / dk if dk * alpha < t2 else alpha)
                        if t1 < alpha:
                            alpha, ibd, at_upper = t1, i, False
                    elif dk > 0 and np.isfinite(u[k]):
                        t2 = u[k] - xcp[k]
                        t1 = 0.0 if t2 <= 0 else (t2 / dk if dk * alpha > t2 else alpha)
                        if t1 < alpha:
                            alpha, ibd, at_upper = t1, i, True
                xbar[F] = xcp[F] + alpha * dF
                if alpha < 1.0 and ibd >= 0:
                    xbar[F[ibd]] = u[F[ibd]] if at_upper else l[F[ibd]]
        return xbar

    # ------------------------------------------------------------------ More'-Thuente line search (dcsrch / dcstep)
    def _dcstep(self, fp, dp):
        ls = self.ls
        stx, fx, dx, sty, fy, dy, stp = ls["stx"], ls["fx"], ls["gx"], ls["sty"], ls["fy"], ls["gy"], ls["stp"]
        brackt, stpmin, stpmax = ls["brackt"], ls["stmin"], ls["stmax"]
        sgnd = dp * (dx / abs(dx))
        if fp > fx:
            theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
            s = max(abs(theta), abs(dx), abs(dp))
            gamma = s * np.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
            if stp < stx:
                gamma = -gamma
            p = (gamma - dx) + theta
            q = ((gamma - dx) + gamma) + dp
            r = p / q
            stpc = stx + r * (stp - stx)
            stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
            stpf = stpc if abs(stpc - stx) < abs(stpq - stx) else stpc + (stpq - stpc) / 2.0
            brackt = True
        elif sgnd < 0.0:
            theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
            s = max(abs(theta), abs(dx), abs(dp))
            gamma = s * np.sqrt((theta / s) ** 2 - (dx / s) * (dp / s))
            if stp > stx:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dx
            r = p / q
            stpc = stp + r * (stx - stp)
            stpq = stp + (dp / (dp - dx)) * (stx - stp)
            stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
            brackt = True
        elif abs(dp) < abs(dx):
            theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
            s = max(abs(theta), abs(dx), abs(dp))
            gamma = s * np.sqrt(max(0.0, (theta / s) ** 2 - (dx / s) * (dp / s)))
            if stp > stx:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = (gamma + (dx - dp)) + gamma
            r = p / q
            if r < 0.0 and gamma != 0.0:
                stpc = stp + r * (stx - stp)
            elif stp > stx:
                stpc = stpmax
            else:
                stpc = stpmin
            stpq = stp + (dp / (dp - dx)) * (stx - stp)
            if brackt:
                stpf = stpc if abs(stpc - stp) < abs(stpq - stp) else stpq
                if stp > stx:
                    stpf = min(stp + 0.66 * (sty - stp), stpf)
                else:
                    stpf = max(stp + 0.66 * (sty - stp), stpf)
            else:
                stpf = stpc if abs(stpc - stp) > abs(stpq - stp) else stpq
                stpf = min(stpmax, stpf)
                stpf = max(stpmin, stpf)
        else:
            if brackt:
                theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp
                s = max(abs(theta), abs(dy), abs(dp))
                gamma = s * np.sqrt((theta / s) ** 2 - (dy / s) * (dp / s))
                if stp > sty:
                    gamma = -gamma
                p = (gamma - dp) + theta
                q = ((gamma - dp) + gamma) + dy
                r = p / q
                stpf = stp + r * (sty - stp)
            elif stp > stx:
                stpf = stpmax
            else:
                stpf = stpmin
        if fp > fx:
            sty, fy, dy = stp, fp, dp
        else:
            if sgnd < 0.0:
                sty, fy, dy = stx, fx, dx
            stx, fx, dx = stp, fp, dp
        ls.update(stx=stx, fx=fx, gx=dx, sty=sty, fy=fy, gy=dy, stp=stpf, brackt=brackt)

    def _dcsrch(self, f, g, start: bool) -> str:
        """Returns 'FG', 'CONVERGENCE', 'WARNING' or 'ERROR'; the trial step is ls['stp']."""
        ls = self.ls
        stpmin, stpmax = 0.0, ls["stpmx"]
        if start:
            if ls["stp"] < stpmin or ls["stp"] > stpmax or g >= 0.0:
                return "ERROR"
            ls.update(brackt=False, stage=1, finit=f, ginit=g, gtest=FTOL * g, width=stpmax - stpmin,
                      width1=2.0 * (stpmax - stpmin), stx=0.0, fx=f, gx=g, sty=0.0, fy=f, gy=g, stmin=0.0,
                      stmax=ls["stp"] + 4.0 * ls["stp"])
            return "FG"
        stp = ls["stp"]
        ftest = ls["finit"] + stp * ls["gtest"]
        if ls["stage"] == 1 and f <= ftest and g >= 0.0:
            ls["stage"] = 2
        if ls["brackt"] and (stp <= ls["stmin"] or stp >= ls["stmax"]):
            return "WARNING"
        if ls["brackt"] and ls["stmax"] - ls["stmin"] <= XTOL * ls["stmax"]:
            return "WARNING"
        if stp == stpmax and f <= ftest and g <= ls["gtest"]:
            return "WARNING"
        if stp == stpmin and (f > ftest or g >= ls["gtest"]):
            return "WARNING"
        if f <= ftest and abs(g) <= GTOL * (-ls["ginit"]):
            return "CONVERGENCE"
        if ls["stage"] == 1 and f <= ls["fx"] and f > ftest:
            gt = ls["gtest"]
            fm, gm = f - stp * gt, g - gt
            ls["fx"], ls["fy"] = ls["fx"] - ls["stx"] * gt, ls["fy"] - ls["sty"] * gt
            ls["gx"], ls["gy"] = ls["gx"] - gt, ls["gy"] - gt
            self._dcstep(fm, gm)
            ls["fx"], ls["fy"] = ls["fx"] + ls["stx"] * gt, ls["fy"] + ls["sty"] * gt
            ls["gx"], ls["gy"] = ls["gx"] + gt, ls["gy"] + gt
        else:
            self._dcstep(f, g)
        if ls["brackt"]:
            if abs(ls["sty"] - ls["stx"]) >= 0.66 * ls["width1"]:
                ls["stp"] = ls["stx"] + 0.5 * (ls["sty"] - ls["stx"])
            ls["width1"] = ls["width"]
            ls["width"] = abs(ls["sty"] - ls["stx"])
        if ls["brackt"]:
            ls["stmin"], ls["stmax"] = min(ls["stx"], ls["sty"]), max(ls["stx"], ls["sty"])
        else:
            ls["stmin"] = ls["stp"] + 1.1 * (ls["stp"] - ls["stx"])
            ls["stmax"] = ls["stp"] + 4.0 * (ls["stp"] - ls["stx"])
        ls["stp"] = min(max(ls["stp"], stpmin), stpmax)
        if (ls["brackt"] and (ls["stp"] <= ls["stmin"] or ls["stp"] >= ls["stmax"])) or \
                (ls["brackt"] and ls["stmax"] - ls["stmin"] <= XTOL * ls["stmax"]):
            ls["stp"] = ls["stx"]
        return "FG"

    # ------------------------------------------------------------------ one iteration's set-up: direction + first trial
    def _begin_iteration(self):
        while True:
            cp = self._cauchy()
            if cp is None:
                self._reset_memory()
                continue
            xcp, cvec, free_mask = cp
            z = self._subsm(xcp, cvec, free_mask)
            if z is None:
                self._reset_memory()
                continue
            break
        x, g, l, u = self.x, self.g, self.l, self.u
        d = z - x
        dtd = float(d @ d)
        dnorm = np.sqrt(dtd)
        stpmx = BIG
        if self.iter == 0:
            stpmx = 1.0
        else:
            for i in range(self.n):
                a1 = d[i]
                if a1 < 0 and np.isfinite(l[i]):
                    a2 = l[i] - x[i]
                    if a2 >= 0:
                        stpmx = 0.0
                    elif a1 * stpmx < a2:
                        stpmx = a2 / a1
                elif a1 > 0 and np.isfinite(u[i]):
                    a2 = u[i] - x[i]
                    if a2 <= 0:
                        stpmx = 0.0
                    elif a1 * stpmx > a2:
                        stpmx = a2 / a1
        stp = min(1.0 / dnorm, stpmx) if (self.iter == 0 and not self.boxed) else 1.0
        self.d, self.z, self.t, self.gold, self.fold = d, z, x.copy(), g.copy(), self.f
        self.dtd = dtd
        gd = float(g @ d)
        self.gdold = gd
        self.ls = dict(stp=stp, stpmx=stpmx, ifun=0, iback=0)
        if gd >= 0.0:
            return self._line_search_failed()
        res = self._dcsrch(self.f, gd, start=True)
        if res != "FG":
            return self._line_search_failed()
        self._take_trial()

    def _take_trial(self):
        ls = self.ls
        ls["ifun"] += 1
        ls["iback"] = ls["ifun"] - 1
        self.x = self.z.copy() if ls["stp"] == 1.0 else ls["stp"] * self.d + self.t
        self.task = FG
        self.phase = "linesearch"

    def _line_search_failed(self):
        self.x, self.f, self.g = self.t.copy(), self.fold, self.gold.copy()
        if self.col == 0:
            self.task, self.message = ABNORMAL, "ABNORMAL: "
            return
        self._reset_memory()
        self._begin_iteration()

    # ------------------------------------------------------------------ reverse communication
    def step(self, f: float, g) -> None:
        """Feed f, g at `self.x`; advances until the next evaluation is needed (`task == FG`) or the run ends."""
        self.f, self.g = float(f), np.asarray(g, dtype=np.float64).copy()
        self.nfev += 1
        if self.phase == "start":
            if self._projgr() <= self.pgtol:
                self.task, self.message = CONVERGED, "CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL"
                return
            self._begin_iteration()
            return
        ls = self.ls
        gd = float(self.g @ self.d)
        res = self._dcsrch(self.f, gd, start=False)
        if res == "FG":
            if ls["iback"] + 1 >= self.maxls:   # the trial just judged was number iback + 1
                return self._line_search_failed()
            return self._take_trial()
        if res == "ERROR":
            return self._line_search_failed()
        # line search done (CONVERGENCE or WARNING): a new iterate
        stp = ls["stp"]
        self.iter += 1
        sbgnrm = self._projgr()
        if self.iter >= self.maxiter:
            self.task, self.message = STOPPED, "STOP: TOTAL NO. OF ITERATIONS REACHED LIMIT"
            return
        if self.nfev > self.maxfun:
            self.task, self.message = STOPPED, "STOP: TOTAL NO. OF F,G EVALUATIONS EXCEEDS LIMIT"
            return
        if sbgnrm <= self.pgtol:
            self.task, self.message = CONVERGED, "CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL"
            return
        ddum = max(abs(self.fold), abs(self.f), 1.0)
        if (self.fold - self.f) <= EPSMCH * self.factr * ddum:
            self.task, self.message = CONVERGED, "CONVERGENCE: RELATIVE REDUCTION OF F <= FACTR*EPSMCH"
            return
        y = self.g - self.gold
        rr = float(y @ y)
        if stp == 1.0:
            dr, ddum = gd - self.gdold, -self.gdold
            s = self.d
        else:
            dr, ddum = (gd - self.gdold) * stp, -self.gdold * stp
            s = stp * self.d
        if dr > EPSMCH * ddum:
            m = self.m
            if self.col == m:  # drop the oldest pair
                self.S[:-1], self.Y[:-1] = self.S[1:].copy(), self.Y[1:].copy()
                self.SS[:-1, :-1], self.SY[:-1, :-1] = self.SS[1:, 1:].copy(), self.SY[1:, 1:].copy()
                self.col -= 1
            c = self.col
            self.S[c], self.Y[c] = s, y
            for j in range(c + 1):
                self.SS[c, j] = self.SS[j, c] = float(self.S[j] @ s)
                self.SY[c, j] = float(s @ self.Y[j])
                self.SY[j, c] = float(self.S[j] @ y)
            self.SS[c, c] = self.dtd if stp == 1.0 else stp * stp * self.dtd
            self.SY[c, c] = dr
            self.col = c + 1
            self.theta = rr / dr
        self._begin_iteration()


def minimize_batched(fun, x0, lower, upper, **kw):
    """Drive N independent state machines with ONE batched evaluation per round (all problems, finished ones included,
    like the device loop).  fun(X: N x D) -> (f: N, g: N x D).  Returns (xs, fs, states, rounds)."""
    N = x0.shape[0]
    states = [LbfgsbState(x0[i], lower, upper, **kw) for i in range(N)]
    rounds = 0
    while any(s.task == FG for s in states):
        X = np.stack([s.x for s in states])
        f, g = fun(X)
        rounds += 1
        for i, s in enumerate(states):
            if s.task == FG:
                s.step(f[i], g[i])
    return np.stack([s.x for s in states]), np.array([s.f for s in states]), states, rounds
