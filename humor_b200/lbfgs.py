"""L-BFGS with strong-Wolfe line search for the fitting stages — same algorithm and defaults as the
``torch.optim.LBFGS(max_iter=20, lr=1.0, line_search_fn='strong_wolfe')`` the reference builds
(humor/fitting/motion_optimizer.py:24,228-231,281-284,461-478), re-planned for two things the library version cannot do:

* **joint optimisation across GPUs** (SURVEY.md §8e): the reference runs ONE L-BFGS over all B sub-sequences (shared step
  length and curvature history).  When the sub-sequences are sharded over ranks, every inner product of the algorithm is
  a global quantity.  Here each rank keeps only its own slice of the variables and of the history; the scalars the
  algorithm branches on (loss, g.d, y.s, ... ) are all-reduced, so every rank takes the same decisions and the iterates
  equal the single-process ones up to summation order.  ``group=None`` / world size 1: no collective is issued.
* **few host synchronisations, few launches**: the library version launches ~4 kernels per history pair per iteration
  (two-loop recursion on the vectors) and reads ~6 scalars per closure evaluation one by one.  Here the two-loop recursion
  runs on the host in COEFFICIENT space over the Gram matrices S.Y, Y.Y (m x m, fp64, updated by one row/column per
  iteration), the device does one (2m+3) x n by n x 3 product per outer iteration and one (2m+3)-term linear combination
  for the direction, and each closure evaluation costs exactly one packed device->host read.

The mathematics (update rule, H0 scaling, step-length initialisation, bracketing / zoom phases, termination tests and
their constants) follows torch.optim.LBFGS so that at world size 1 the iterates agree with it to rounding
(tests/test_lbfgs.py).  ``MotionOptimizer`` keeps using the library class at world size 1 unless told otherwise.
"""
import math

import numpy as np
import torch
from torch.optim import Optimizer


def _cubic_interpolate(x1, f1, g1, x2, f2, g2, bounds=None):
    """Minimiser of the cubic through (x1,f1,g1), (x2,f2,g2), clamped to the bounds (polyinterp)."""
    if bounds is not None:
        lo, hi = bounds
    else:
        lo, hi = (x1, x2) if x1 <= x2 else (x2, x1)
    d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)
    d2sq = d1 * d1 - g1 * g2
    if d2sq >= 0:
        d2 = math.sqrt(d2sq)
        if x1 <= x2:
            pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))
        else:
            pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2 * d2))
        return min(max(pos, lo), hi)
    return (lo + hi) / 2.0


class LBFGS(Optimizer):
    """Drop-in for ``torch.optim.LBFGS`` (strong-Wolfe) with an optional process group for sharded variables.

    ``group``: a ``torch.distributed`` process group (or ``True`` for the default group).  Every rank must call
    ``step`` with a closure that returns ITS share of the loss (terms that couple two ranks counted once overall)
    and leaves its own gradients in ``.grad``.
    """

    def __init__(self, params, lr=1.0, max_iter=20, max_eval=None, tolerance_grad=1e-7, tolerance_change=1e-9,
                 history_size=100, line_search_fn='strong_wolfe', group=None):
        if max_eval is None:
            max_eval = max_iter * 5 // 4
        if line_search_fn not in (None, 'strong_wolfe'):
            raise RuntimeError("only 'strong_wolfe' is supported")
        defaults = dict(lr=lr, max_iter=max_iter, max_eval=max_eval, tolerance_grad=tolerance_grad,
                        tolerance_change=tolerance_change, history_size=history_size, line_search_fn=line_search_fn)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("LBFGS doesn't support per-parameter options (parameter groups)")
        self._params = self.param_groups[0]['params']
        self._n = sum(p.numel() for p in self._params)
        self._group = group
        self._dist = None
        if group is not None:
            import torch.distributed as dist
            self._dist = dist
            if group is True:
                self._group = dist.group.WORLD
        self.syncs = 0                       # packed device->host reads so far (one per closure evaluation + one per iteration)
        self._st = None

    # ------------------------------------------------------------------------------------------------ helpers
    def _flat_grad(self):
        views = []
        for p in self._params:
            if p.grad is None:
                views.append(p.new_zeros(p.numel()))
            else:
                views.append(p.grad.reshape(-1))
        return torch.cat(views, 0)

    def _flat_params(self):
        return torch.cat([p.detach().reshape(-1) for p in self._params], 0)

    def _set_params(self, flat):
        off = 0
        for p in self._params:
            n = p.numel()
            p.copy_(flat[off:off + n].view_as(p))
            off += n

    def _read(self, sums=(), maxs=()):
        """One packed read: SUM-reduced and MAX-reduced scalars (across ranks when sharded) -> python floats."""
        out = []
        ns = 0
        if len(sums):
            s = torch.cat([x.reshape(-1).double() for x in sums])
            ns = s.numel()
            if self._dist is not None:
                self._dist.all_reduce(s, op=self._dist.ReduceOp.SUM, group=self._group)
            out.append(s)
        if len(maxs):
            m = torch.cat([x.reshape(-1).double() for x in maxs])
            if self._dist is not None:
                self._dist.all_reduce(m, op=self._dist.ReduceOp.MAX, group=self._group)
            out.append(m)
        vals = torch.cat(out).tolist()
        self.syncs += 1
        return vals[:ns], vals[ns:]

    def _init_state(self, like, m):
        n = self._n
        st = {
            'func_evals': 0, 'n_iter': 0,
            # rows [0,m): s_i, [m,2m): y_i, 2m: candidate s, 2m+1: candidate y, 2m+2: current gradient
            'H': torch.zeros(2 * m + 3, n, dtype=like.dtype, device=like.device),
            'order': [],                       # history slots, oldest first
            'SY': np.zeros((m, m)), 'YY': np.zeros((m, m)),         # s_i.y_j and y_i.y_j by slot (fp64, host)
            'gamma': 1.0, 'd': None, 't': None, 'prev_g': None, 'prev_loss': None,
        }
        return st

    # ------------------------------------------------------------------------------------------------ direction
    def _direction(self, st, g, m):
        """History update with the pair (s, y) of the previous iteration, then d = -H g by the two-loop recursion in
        coefficient space.  Returns (d, gtd, device max|d|)."""
        H = st['H']
        H[2 * m] = st['d'] * st['t']                  # s
        H[2 * m + 1] = g - st['prev_g']               # y
        H[2 * m + 2] = g
        P = H @ H[2 * m:2 * m + 3].t()                # columns: .s  .y  .g
        p, _ = self._read(sums=[P])
        P = np.asarray(p, np.float64).reshape(2 * m + 3, 3)
        ys, yy = P[2 * m, 1], P[2 * m + 1, 1]
        order, SY, YY = st['order'], st['SY'], st['YY']
        if ys > 1e-10:
            slot = order.pop(0) if len(order) == m else next(i for i in range(m) if i not in order)
            o = np.asarray(order, np.int64)
            SY[slot, o] = P[m + o, 0]                 # s_new . y_j
            SY[o, slot] = P[o, 1]                     # s_j . y_new
            YY[slot, o] = YY[o, slot] = P[m + o, 1]   # y_j . y_new
            SY[slot, slot], YY[slot, slot] = ys, yy
            H[slot] = H[2 * m]
            H[m + slot] = H[2 * m + 1]
            P[slot], P[m + slot] = P[2 * m], P[2 * m + 1]
            order.append(slot)
            st['gamma'] = ys / yy
        o = np.asarray(order, np.int64)
        k = len(order)
        Sg, Yg, gg = P[o, 2], P[m + o, 2], P[2 * m + 2, 2]
        SYo, YYo = SY[np.ix_(o, o)], YY[np.ix_(o, o)]
        rho_inv = np.diag(SYo).copy()                 # y_i . s_i
        gamma = st['gamma']
        # q = cg*g + sum cy_j y_j ; first loop newest -> oldest
        cg = -1.0
        cy = np.zeros(k)
        al = np.zeros(k)
        for i in range(k - 1, -1, -1):
            al[i] = (cg * Sg[i] + SYo[i] @ cy) / rho_inv[i]
            cy[i] -= al[i]
        # r = gamma*q + sum cs_j s_j ; second loop oldest -> newest
        cs = np.zeros(k)
        for i in range(k):
            be = (gamma * (cg * Yg[i] + YYo[i] @ cy) + SYo[:, i] @ cs) / rho_inv[i]
            cs[i] += al[i] - be
        coef = np.zeros(2 * m + 3)
        coef[o] = cs
        coef[m + o] = gamma * cy
        coef[2 * m + 2] = gamma * cg
        c = torch.from_numpy(coef).to(H.dtype).to(H.device, non_blocking=True)
        d = c @ H
        gtd = gamma * cg * gg + gamma * float(cy @ Yg) + float(cs @ Sg)
        return d, gtd, d.abs().max()

    # ------------------------------------------------------------------------------------------------ step
    @torch.no_grad()
    def step(self, closure):
        closure = torch.enable_grad()(closure)
        grp = self.param_groups[0]
        lr, max_iter, max_eval = float(grp['lr']), grp['max_iter'], grp['max_eval']
        tol_g, tol_x, m = grp['tolerance_grad'], grp['tolerance_change'], grp['history_size']
        line_search = grp['line_search_fn']

        orig_loss = closure()
        g = self._flat_grad()
        if self._st is None:
            self._st = self._init_state(g, m)
        st = self._st
        extra = [g.abs().sum(), g.dot(g)] if st['n_iter'] == 0 else []
        (loss, *l1), (gmax,) = self._read(sums=[orig_loss.detach()] + extra, maxs=[g.abs().max()])
        evals = 1
        st['func_evals'] += 1
        if gmax <= tol_g:
            return orig_loss
        d, t, prev_loss = st['d'], st['t'], st['prev_loss']
        n_iter = 0
        while n_iter < max_iter:
            n_iter += 1
            st['n_iter'] += 1
            if st['n_iter'] == 1:
                d = g.neg()
                st['order'] = []
                st['gamma'] = 1.0
                # |g|_1 and g.g came with the loss; max|d| = max|g|
                gtd, dmax_dev, dmax = -l1[1], None, gmax
                t = min(1.0, 1.0 / l1[0]) * lr
            else:
                st['d'], st['t'] = d, t
                d, gtd, dmax_dev = self._direction(st, g, m)
                dmax = None
                t = lr
            if st['prev_g'] is None:
                st['prev_g'] = g.clone()
            else:
                st['prev_g'].copy_(g)
            prev_loss = loss
            if gtd > -tol_x:
                break
            ls_evals = 0
            x0 = self._flat_params().clone()

            def evaluate(tt, want_dmax=False):
                """f, g, g.d, max|g| (and max|d| on request) at x0 + tt*d — one packed read."""
                self._set_params(torch.add(x0, d, alpha=tt))
                f_dev = closure()
                gn = self._flat_grad()
                mx = [gn.abs().max()] + ([dmax_dev] if want_dmax else [])
                (f, gd), mxs = self._read(sums=[f_dev.detach(), gn.dot(d)], maxs=mx)
                return f, gn, gd, mxs

            if line_search is None:
                self._set_params(torch.add(x0, d, alpha=t))
                if n_iter != max_iter:
                    f_dev = closure()
                    g = self._flat_grad()
                    mx = [g.abs().max()] + ([dmax_dev] if dmax is None else [])
                    (loss,), mxs = self._read(sums=[f_dev.detach()], maxs=mx)
                    gmax = mxs[0]
                    dmax = mxs[1] if dmax is None else dmax
                    ls_evals = 1
                elif dmax is None:
                    _, (dmax,) = self._read(maxs=[dmax_dev])
            else:
                # ---- strong-Wolfe line search (bracketing, then zoom); all comparisons on host scalars
                max_ls = max_eval - evals
                c1, c2, ls_tol = 1e-4, 0.9, 1e-9
                f0, g0, gtd0 = loss, g, gtd
                f_new, g_new, gtd_new, mxs = evaluate(t, want_dmax=dmax is None)
                if dmax is None:
                    dmax = mxs[1]
                gmax_new = mxs[0]
                ls_evals = 1
                t_prev, f_prev, g_prev, gtd_prev, gmax_prev = 0.0, f0, g0, gtd0, gmax
                done = False
                ls_iter = 0
                br = None
                while ls_iter < max_ls:
                    if f_new > (f0 + c1 * t * gtd0) or (ls_iter > 1 and f_new >= f_prev):
                        br = [[t_prev, f_prev, g_prev, gtd_prev, gmax_prev], [t, f_new, g_new, gtd_new, gmax_new]]
                        break
                    if abs(gtd_new) <= -c2 * gtd0:
                        br = [[t, f_new, g_new, gtd_new, gmax_new]]
                        done = True
                        break
                    if gtd_new >= 0:
                        br = [[t_prev, f_prev, g_prev, gtd_prev, gmax_prev], [t, f_new, g_new, gtd_new, gmax_new]]
                        break
                    min_step = t + 0.01 * (t - t_prev)
                    max_step = t * 10
                    tmp = t
                    t = _cubic_interpolate(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, bounds=(min_step, max_step))
                    t_prev, f_prev, g_prev, gtd_prev, gmax_prev = tmp, f_new, g_new, gtd_new, gmax_new
                    f_new, g_new, gtd_new, (gmax_new,) = evaluate(t)
                    ls_evals += 1
                    ls_iter += 1
                if br is None:                       # ran out of evaluations while extrapolating
                    br = [[0.0, f0, g0, gtd0, gmax], [t, f_new, g_new, gtd_new, gmax_new]]
                insuf = False
                lo, hi = (0, 1) if br[0][1] <= br[-1][1] else (1, 0)
                while not done and ls_iter < max_ls:
                    if abs(br[1][0] - br[0][0]) * dmax < ls_tol:
                        break
                    t = _cubic_interpolate(br[0][0], br[0][1], br[0][3], br[1][0], br[1][1], br[1][3])
                    bmax, bmin = max(br[0][0], br[1][0]), min(br[0][0], br[1][0])
                    eps = 0.1 * (bmax - bmin)
                    if min(bmax - t, t - bmin) < eps:
                        if insuf or t >= bmax or t <= bmin:
                            t = bmax - eps if abs(t - bmax) < abs(t - bmin) else bmin + eps
                            insuf = False
                        else:
                            insuf = True
                    else:
                        insuf = False
                    f_new, g_new, gtd_new, (gmax_new,) = evaluate(t)
                    ls_evals += 1
                    ls_iter += 1
                    if f_new > (f0 + c1 * t * gtd0) or f_new >= br[lo][1]:
                        br[hi] = [t, f_new, g_new, gtd_new, gmax_new]
                        lo, hi = (0, 1) if br[0][1] <= br[1][1] else (1, 0)
                    else:
                        if abs(gtd_new) <= -c2 * gtd0:
                            done = True
                        elif gtd_new * (br[hi][0] - br[lo][0]) >= 0:
                            br[hi] = br[lo]
                        br[lo] = [t, f_new, g_new, gtd_new, gmax_new]
                if len(br) == 1:
                    lo = 0
                t, loss, g, _, gmax = br[lo]
                self._set_params(torch.add(x0, d, alpha=t))
            evals += ls_evals
            st['func_evals'] += ls_evals
            if n_iter == max_iter:
                break
            if evals >= max_eval:
                break
            if gmax <= tol_g:
                break
            if dmax * abs(t) <= tol_x:
                break
            if abs(loss - prev_loss) < tol_x:
                break
        st['d'], st['t'], st['prev_loss'] = d, t, prev_loss
        return orig_loss
