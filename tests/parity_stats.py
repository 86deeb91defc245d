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
(torch/optim/lbfgs.py:12-37) whose result moves by more than 1e-4 relative under the perturbations two
independent fp32 implementations of the closure show against each other — the newer loss value off by up
to PROBE_ULPS fp32 ulps (this library's teacher-forced energy is within 5.3e-7 of the reference's, the
reference's own CUDA path within 3.5e-7 of its CPU path), a directional derivative off by 2^-18 relative
(different summation order of a 2048-element dot product) — i.e. the bisection / extrapolation-bound branch
flip of the discriminant d1^2 - g1*g2 (tests/golden/make_golden_dist.py records the
calls' arguments, the probes re-run oracle/lbfgs_np.py's restatement of the interpolation on them).  Divergences that first show after `LATE_EVAL` evaluations are accumulated
round-off (the reference against itself shows the same), not a single decision.
"""
from __future__ import annotations

import numpy as np

E_TOL = 1e-4
JOINT_TOL_MM = 0.5
SENS_TOL = 1e-4          # relative change of the interpolated step under the probes
PROBE_ULPS = 8           # loss perturbation (fp32 ulps)
PROBE_GREL = 2.0 ** -18  # relative perturbation of a directional derivative
LATE_EVAL = 8            # first divergence at or after this evaluation: accumulated round-off
ONSET_TOL = 2e-6         # relative energy difference that marks the ONSET of a divergence: four times the closure's
                         # implementation noise (5e-7); the 1e-4 bar is typically crossed a few evaluations later


def onset_of(E_a, n_a, E_b, n_b):
    """First evaluation at which run b's energy leaves run a's by more than ONSET_TOL (n if never)."""
    n = int(min(n_a, n_b))
    a, b = np.asarray(E_a[:n], np.float64), np.asarray(E_b[:n], np.float64)
    bad = np.nonzero(np.abs(a - b) / np.abs(np.asarray(E_a[:int(n_a)], np.float64)).max() > ONSET_TOL)[0]
    return int(bad[0]) if len(bad) else n


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


def _probe(r, cols):
    """Largest relative change of one recorded call's result under the probes."""
    from oracle.lbfgs_np import _cubic_interpolate
    F = np.float32
    x1, f1, g1, x2, f2, g2 = (float(r[cols[c]]) for c in ("x1", "f1", "g1", "x2", "f2", "g2"))
    lo, hi = float(r[cols["lo"]]), float(r[cols["hi"]])
    bounds = None if np.isnan(lo) else (F(lo), F(hi))
    t0 = float(r[cols["t"]])
    if t0 == 0:
        return 0.0

    def run(f2_, g1_, g2_):
        with np.errstate(all="ignore"):
            return float(_cubic_interpolate(F(x1) if x1 != 0 else 0.0, f1, F(g1_), F(x2), f2_, F(g2_), bounds))

    best = 0.0
    for k in range(1, PROBE_ULPS + 1):
        for sgn in (-1.0, 1.0):
            f2p = float(F(f2) + sgn * k * np.spacing(np.abs(F(f2))))
            best = max(best, abs(run(f2p, g1, g2) - t0) / abs(t0))
    for sgn in (-1.0, 1.0):
        best = max(best, abs(run(f2, g1, g2 * (1 + sgn * PROBE_GREL)) - t0) / abs(t0))
        best = max(best, abs(run(f2, g1 * (1 + sgn * PROBE_GREL), g2) - t0) / abs(t0))
    pert = r[[cols["t_m2ulp"], cols["t_m1ulp"], cols["t_p1ulp"], cols["t_p2ulp"]]]         # recorded with torch's own types
    if np.isfinite(pert).any():
        best = max(best, float(np.nanmax(np.abs(pert - t0)) / abs(t0)))
    return best


def cubic_sensitivity(rows, cols, k):
    """Largest relative change of an interpolated step under the probes, over the calls that produced the trial
    points of evaluations k-2 .. k."""
    best = 0.0
    for r in rows:
        ev = int(r[cols["eval_index"]])
        if k - 2 <= ev <= k:
            best = max(best, _probe(r, cols))
    return best


def classify(lead, n_a, n_b, rows, cols, onset=None):
    """'' (no divergence), 'cubic' (the divergence sets in at an ill-conditioned interpolation), 'late' (accumulated
    round-off) or 'unexplained'.  onset: evaluation at which the energies first part by ONSET_TOL (default: lead)."""
    if lead == min(n_a, n_b) and n_a == n_b:
        return ""
    k = lead if onset is None else min(onset, lead)
    if cubic_sensitivity(rows, cols, k) > SENS_TOL:
        return "cubic"
    if k >= LATE_EVAL:
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
        kinds.append(classify(lead, int(ne[0, stage, w]), int(ne[1, stage, w]), rows[(stage, w)], cols,
                              onset_of(E[0, stage, w], ne[0, stage, w], E[1, stage, w], ne[1, stage, w])))
    return summarize(f"reference {int(g['threads'][1])} threads vs 1 thread", mm, strict, kinds,
                     ne[1, stage].astype(int) - ne[0, stage].astype(int))


def format_table(rows):
    keys = ["name", "n", "q50_mm", "q75_mm", "q90_mm", "max_mm", "frac_within_0.5mm", "strict", "cubic", "late",
            "unexplained", "mean_abs_n_eval_diff"]
    lines = ["  ".join(keys)]
    for r in rows:
        lines.append("  ".join(("%.4g" % r[k]) if isinstance(r[k], float) else str(r[k]) for k in keys))
    return "\n".join(lines)
