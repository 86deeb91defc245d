"""Shared helpers of the distributional parity tests (tests/test_oracle_dist.py on CPU,
tests/test_gpu_dist.py on the B200) over tests/golden/dist.npz.

A "run" is one L-BFGS stage solve of one window: its per-evaluation energies E[0..n-1] and its
final window pose.  Two runs of the same problem are compared by

  * lead      = number of leading evaluations whose energies agree to 1e-4 of the run's largest
                energy (the north star's per-iteration bar);
  * joints_mm = largest final-joint coordinate difference in millimetres (bar: 0.5 mm);
  * strict    = every evaluation within 1e-4, same evaluation count, joints within 0.5 mm.

A divergence is *explained* when the reference's own trace shows that the trial point of the first
diverging evaluation (or one of the three before it) came out of a `_cubic_interpolate` call
(torch/optim/lbfgs.py:12-37) whose result moves by more than 1e-4 relative when the newer loss value
changes by at most two fp32 ulps — the bisection / extrapolation-bound branch flip of the
discriminant d1^2 - g1*g2 (tests/golden/make_golden_dist.py records the
perturbed results).  Divergences that first show after `LATE_EVAL` evaluations are accumulated
round-off (the reference against itself shows the same), not a single decision.
"""
from __future__ import annotations

import numpy as np

E_TOL = 1e-4
JOINT_TOL_MM = 0.5
SENS_TOL = 1e-4          # relative change of the interpolated step under a <= 2 ulp change of the loss
LATE_EVAL = 8            # first divergence at or after this evaluation: accumulated round-off


def compare_runs(E_a, n_a, pose_a, E_b, n_b, pose_b):
    """(lead, joints_mm, strict) of run b against run a (a = the reference at one thread)."""
    n = int(min(n_a, n_b))
    a, b = np.asarray(E_a[:n], np.float64), np.asarray(E_b[:n], np.float64)
    rel = np.abs(a - b) / np.abs(np.asarray(E_a[:int(n_a)], np.float64)).max()
    bad = np.nonzero(rel > E_TOL)[0]
    lead = int(bad[0]) if len(bad) else n
    mm = float(np.abs(np.asarray(pose_a, np.float64) - np.asarray(pose_b, np.float64)).max() * 1000.0)
    strict = lead == n and int(n_a) == int(n_b) and mm < JOINT_TOL_MM
    return lead, mm, strict


def cubic_rows(g, max_iter):
    """{(stage, window): rows} of the reference's recorded _cubic_interpolate calls."""
    cub, idx = g[f"mi{max_iter}_cubic"], g[f"mi{max_iter}_cubic_index"]
    cols = {str(c): i for i, c in enumerate(g["cubic_columns"])}
    out, off = {}, 0
    for si, wi, n in idx:
        out[(int(si), int(wi))] = cub[off:off + int(n)]
        off += int(n)
    return out, cols


def cubic_sensitivity(rows, cols, k):
    """Largest relative change of an interpolated step under a <= 2 ulp change of the loss, over the calls
    that produced the trial points of evaluations k-3 .. k."""
    best = 0.0
    for r in rows:
        ev = int(r[cols["eval_index"]])
        if k - 3 <= ev <= k:
            t = r[cols["t"]]
            pert = r[[cols["t_m2ulp"], cols["t_m1ulp"], cols["t_p1ulp"], cols["t_p2ulp"]]]
            if np.isfinite(pert).any() and t != 0:
                best = max(best, float(np.nanmax(np.abs(pert - t)) / abs(t)))
    return best


def classify(lead, n_a, n_b, rows, cols):
    """'' (no divergence), 'cubic' (ill-conditioned interpolation), 'late' (accumulated round-off) or
    'unexplained'."""
    if lead == min(n_a, n_b) and n_a == n_b:
        return ""
    if cubic_sensitivity(rows, cols, lead) > SENS_TOL:
        return "cubic"
    if lead >= LATE_EVAL:
        return "late"
    return "unexplained"


def summarize(name, mm, strict, kinds, n_eval_diff):
    mm = np.asarray(mm)
    return {"name": name, "n": len(mm), "q50_mm": float(np.quantile(mm, 0.5)), "q75_mm": float(np.quantile(mm, 0.75)),
            "q90_mm": float(np.quantile(mm, 0.9)), "max_mm": float(mm.max()),
            "frac_within_0.5mm": float((mm < JOINT_TOL_MM).mean()), "strict": int(np.sum(strict)),
            "cubic": kinds.count("cubic"), "late": kinds.count("late"), "unexplained": kinds.count("unexplained"),
            "mean_abs_n_eval_diff": float(np.mean(np.abs(n_eval_diff)))}


def reference_self_noise(g, max_iter, stage):
    """The reference at N threads against the reference at 1 thread (same inputs), summarised."""
    E, P, ne = g[f"mi{max_iter}_E"], g[f"mi{max_iter}_pose"], g[f"mi{max_iter}_n_eval"]
    rows, cols = cubic_rows(g, max_iter)
    W = E.shape[2]
    mm, strict, kinds = [], [], []
    for w in range(W):
        lead, d, s = compare_runs(E[0, stage, w], ne[0, stage, w], P[0, stage, w], E[1, stage, w], ne[1, stage, w],
                                  P[1, stage, w])
        mm.append(d), strict.append(s)
        kinds.append(classify(lead, int(ne[0, stage, w]), int(ne[1, stage, w]), rows[(stage, w)], cols))
    return summarize(f"reference {int(g['threads'][1])} threads vs 1 thread", mm, strict, kinds,
                     ne[1, stage].astype(int) - ne[0, stage].astype(int))


def format_table(rows):
    keys = ["name", "n", "q50_mm", "q75_mm", "q90_mm", "max_mm", "frac_within_0.5mm", "strict", "cubic", "late",
            "unexplained", "mean_abs_n_eval_diff"]
    lines = ["  ".join(keys)]
    for r in rows:
        lines.append("  ".join(("%.4g" % r[k]) if isinstance(r[k], float) else str(r[k]) for k in keys))
    return "\n".join(lines)
