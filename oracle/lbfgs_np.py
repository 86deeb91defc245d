"""numpy restatement of ``torch.optim.LBFGS`` with ``line_search_fn='strong_wolfe'``
as the reference runs it (reference optimizer.py:261-270: lr=2, max_iter=25,
tolerance_change=1e-6, defaults max_eval = max_iter*5//4, tolerance_grad=1e-7,
history_size=100).  Follows torch 2.11.0 ``torch/optim/lbfgs.py`` — the
third-party dependency the reference calls; SURVEY.md Appendix B.

Oracle: test infrastructure only (see oracle/__init__.py).

Scalar typing is emulated, not simplified: torch mixes Python floats (loss
values, step sizes that came from Python arithmetic) with fp32 0-d tensors
(dot products, interpolated steps).  Under numpy >= 2 (NEP 50) ``np.float32``
scalars combine with Python floats exactly the way fp32 0-d tensors do (the
Python scalar is cast to fp32, the result stays fp32), so "tensor" scalars are
``np.float32`` here and "Python" scalars are ``float``.

The closure maps a flat fp32 vector to ``(float(loss), grad fp32)``.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32

# ---------------------------------------------------------------------------
# Conditioning diagnostics (not part of the algorithm).  torch's line search takes
# branch decisions on quantities that can be pure round-off: on a nearly linear ray the
# cubic interpolation's discriminant d1^2 - g1*g2 is a difference of two equal numbers
# and its SIGN (bisection vs. extrapolation to the bound) flips with a 1-ulp change of
# the loss.  Each decision records how far it was from flipping; tests use the first
# evaluation that depends on a marginal decision as the point after which two fp32
# implementations (or the reference on 1 vs 8 threads) may part ways.
# ---------------------------------------------------------------------------
_DIAG = None


def _note(kind, margin):
    if _DIAG is not None:
        _DIAG["events"].append((len(_DIAG["trace"]), kind, float(margin)))


def _rel_gap(a, b, scale):
    scale = max(abs(float(scale)), 1e-30)
    return abs(float(a) - float(b)) / scale


def _cubic_interpolate(x1, f1, g1, x2, f2, g2, bounds=None):
    """lbfgs.py:12-37."""
    if bounds is not None:
        xmin_bound, xmax_bound = bounds
    else:
        xmin_bound, xmax_bound = (x1, x2) if x1 <= x2 else (x2, x1)
    d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)
    d2_square = d1 ** 2 - g1 * g2
    _note("cubic_disc", abs(float(d2_square)) / max(float(d1) ** 2, abs(float(g1) * float(g2)), 1e-30))
    if d2_square >= 0:
        d2 = np.sqrt(d2_square)
        if x1 <= x2:
            min_pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))
        else:
            min_pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2 * d2))
        return min(max(min_pos, xmin_bound), xmax_bound)
    return (xmin_bound + xmax_bound) / 2.0


def _strong_wolfe(obj_func, t, d, f, g, gtd, c1=1e-4, c2=0.9, tolerance_change=1e-9, max_ls=25):
    """lbfgs.py:40-209.  obj_func(t) evaluates at x + t*d and returns (loss, grad)."""
    d_norm = np.abs(d).max()
    g = g.copy()
    f_new, g_new = obj_func(t)
    ls_func_evals = 1
    gtd_new = np.dot(g_new, d)

    t_prev, f_prev, g_prev, gtd_prev = 0, f, g, gtd
    done = False
    ls_iter = 0
    while ls_iter < max_ls:
        _note("armijo", _rel_gap(f_new, f + c1 * t * gtd, f) / 6e-8)
        if ls_iter > 1:
            _note("f_vs_prev", _rel_gap(f_new, f_prev, f) / 6e-8)
        _note("wolfe", _rel_gap(abs(gtd_new), -c2 * gtd, gtd))
        _note("gtd_sign", _rel_gap(gtd_new, 0.0, gtd))
        if f_new > (f + c1 * t * gtd) or (ls_iter > 1 and f_new >= f_prev):
            bracket = [t_prev, t]
            bracket_f = [f_prev, f_new]
            bracket_g = [g_prev, g_new.copy()]
            bracket_gtd = [gtd_prev, gtd_new]
            break
        if abs(gtd_new) <= -c2 * gtd:
            bracket = [t]
            bracket_f = [f_new]
            bracket_g = [g_new]
            done = True
            break
        if gtd_new >= 0:
            bracket = [t_prev, t]
            bracket_f = [f_prev, f_new]
            bracket_g = [g_prev, g_new.copy()]
            bracket_gtd = [gtd_prev, gtd_new]
            break
        min_step = t + 0.01 * (t - t_prev)
        max_step = t * 10
        tmp = t
        t = _cubic_interpolate(t_prev, f_prev, gtd_prev, t, f_new, gtd_new, bounds=(min_step, max_step))
        t_prev = tmp
        f_prev = f_new
        g_prev = g_new.copy()
        gtd_prev = gtd_new
        f_new, g_new = obj_func(t)
        ls_func_evals += 1
        gtd_new = np.dot(g_new, d)
        ls_iter += 1

    if ls_iter == max_ls:
        bracket = [0, t]
        bracket_f = [f, f_new]
        bracket_g = [g, g_new]

    insuf_progress = False
    low_pos, high_pos = (0, 1) if bracket_f[0] <= bracket_f[-1] else (1, 0)
    while not done and ls_iter < max_ls:
        if abs(bracket[1] - bracket[0]) * d_norm < tolerance_change:
            break
        t = _cubic_interpolate(bracket[0], bracket_f[0], bracket_gtd[0], bracket[1], bracket_f[1], bracket_gtd[1])
        eps = 0.1 * (max(bracket) - min(bracket))
        if min(max(bracket) - t, t - min(bracket)) < eps:
            if insuf_progress or t >= max(bracket) or t <= min(bracket):
                if abs(t - max(bracket)) < abs(t - min(bracket)):
                    t = max(bracket) - eps
                else:
                    t = min(bracket) + eps
                insuf_progress = False
            else:
                insuf_progress = True
        else:
            insuf_progress = False

        f_new, g_new = obj_func(t)
        ls_func_evals += 1
        gtd_new = np.dot(g_new, d)
        ls_iter += 1

        _note("armijo", _rel_gap(f_new, f + c1 * t * gtd, f) / 6e-8)
        _note("f_vs_low", _rel_gap(f_new, bracket_f[low_pos], f) / 6e-8)
        _note("wolfe", _rel_gap(abs(gtd_new), -c2 * gtd, gtd))
        if f_new > (f + c1 * t * gtd) or f_new >= bracket_f[low_pos]:
            bracket[high_pos] = t
            bracket_f[high_pos] = f_new
            bracket_g[high_pos] = g_new.copy()
            bracket_gtd[high_pos] = gtd_new
            low_pos, high_pos = (0, 1) if bracket_f[0] <= bracket_f[1] else (1, 0)
        else:
            if abs(gtd_new) <= -c2 * gtd:
                done = True
            elif gtd_new * (bracket[high_pos] - bracket[low_pos]) >= 0:
                bracket[high_pos] = bracket[low_pos]
                bracket_f[high_pos] = bracket_f[low_pos]
                bracket_g[high_pos] = bracket_g[low_pos]
                bracket_gtd[high_pos] = bracket_gtd[low_pos]
            bracket[low_pos] = t
            bracket_f[low_pos] = f_new
            bracket_g[low_pos] = g_new.copy()
            bracket_gtd[low_pos] = gtd_new

    t = bracket[low_pos]
    return bracket_f[low_pos], bracket_g[low_pos], t, ls_func_evals


def first_marginal_eval(events, cubic_tol=1e-2, ulp_tol=64.0, rel_tol=1e-3):
    """Index of the first closure evaluation whose position (or whose acceptance)
    depends on a decision that round-off could flip; None if the run is well conditioned."""
    for n_evals_so_far, kind, margin in events:
        tol = cubic_tol if kind == "cubic_disc" else (ulp_tol if kind in ("armijo", "f_vs_prev", "f_vs_low", "dloss")
                                                      else rel_tol)
        if margin < tol:
            return n_evals_so_far
    return None


def lbfgs_minimize(closure, x0, lr=2, max_iter=25, max_eval=None, tolerance_grad=1e-7,
                   tolerance_change=1e-6, history_size=100):
    """One ``LBFGS.step(closure)`` from a fresh optimiser (lbfgs.py:332-537).
    Returns (x, info) with info = dict(n_iter, func_evals, t, f, trace, events)."""
    global _DIAG
    if max_eval is None:
        max_eval = max_iter * 5 // 4
    x = np.array(x0, dtype=F32).reshape(-1).copy()
    trace = []
    _DIAG = dict(events=[], trace=trace)
    try:
        return _lbfgs_minimize(closure, x, trace, lr, max_iter, max_eval, tolerance_grad, tolerance_change,
                               history_size, _DIAG["events"])
    finally:
        _DIAG = None


def _lbfgs_minimize(closure, x, trace, lr, max_iter, max_eval, tolerance_grad, tolerance_change, history_size,
                    events):

    def evaluate(z):
        f, g = closure(z)
        trace.append((float(f), z.copy()))
        return float(f), np.asarray(g, dtype=F32).reshape(-1)

    loss, flat_grad = evaluate(x)
    current_evals = 1
    info = dict(n_iter=0, func_evals=1, t=None, f=loss, trace=trace, events=events)
    if np.abs(flat_grad).max() <= tolerance_grad:
        return x, info

    old_dirs, old_stps, ro = [], [], []
    H_diag = 1
    d = t = prev_flat_grad = None
    n_iter = 0
    while n_iter < max_iter:
        n_iter += 1
        if n_iter == 1:
            d = -flat_grad
        else:
            y = flat_grad - prev_flat_grad
            s = d * F32(t)
            ys = np.dot(y, s)
            _note("ys", _rel_gap(ys, 1e-10, 1e-10))
            if ys > 1e-10:
                if len(old_dirs) == history_size:
                    old_dirs.pop(0), old_stps.pop(0), ro.pop(0)
                old_dirs.append(y)
                old_stps.append(s)
                ro.append(1.0 / ys)
                H_diag = ys / np.dot(y, y)
            num_old = len(old_dirs)
            al = [None] * num_old
            q = -flat_grad
            for i in range(num_old - 1, -1, -1):
                al[i] = np.dot(old_stps[i], q) * ro[i]
                q = q + old_dirs[i] * F32(-al[i])
            d = r = q * F32(H_diag)
            for i in range(num_old):
                be_i = np.dot(old_dirs[i], r) * ro[i]
                r = r + old_stps[i] * F32(al[i] - be_i)
            d = r
        prev_flat_grad = flat_grad.copy()
        prev_loss = loss

        if n_iter == 1:
            t = min(1.0, 1.0 / np.abs(flat_grad).sum(dtype=F32)) * lr
        else:
            t = lr
        gtd = np.dot(flat_grad, d)
        if gtd > -tolerance_change:
            break

        def obj_func(step):
            return evaluate(x + F32(step) * d)

        # NB: torch does not forward tolerance_change here (lbfgs.py:478-487), so the
        # line search runs with its own default of 1e-9.
        loss, flat_grad, t, ls_func_evals = _strong_wolfe(
            obj_func, t, d, loss, flat_grad, gtd, max_ls=max_eval - current_evals)
        x = x + F32(t) * d
        opt_cond = np.abs(flat_grad).max() <= tolerance_grad
        current_evals += ls_func_evals

        if n_iter == max_iter:
            break
        if current_evals >= max_eval:
            break
        if opt_cond:
            break
        if np.abs(d * F32(t)).max() <= tolerance_change:
            break
        _note("dloss", _rel_gap(abs(loss - prev_loss), tolerance_change, loss) / 6e-8)
        if abs(loss - prev_loss) < tolerance_change:
            break

    info.update(n_iter=n_iter, func_evals=current_evals, t=float(t), f=loss)
    return x, info
