"""Distributional free-running parity on the B200 (reference
optimizer.py:261-270): 64 windows x 2 stages at max_iter 3 and 25, every run compared with the reference at one CPU
thread (tests/golden/dist.npz from the unmodified reference; tests/test_oracle_dist.py pins it on CPU):

  (a) this library's CUDA path;
  (b) the reference at eight CPU threads (golden) — its energies are bit-identical to the 1-thread run for the first
      evaluations of most windows (same oneDNN kernels, only some reductions reorder), so this is a LOWER bound of
      what any independent fp32 implementation can show;
  (c) the reference on its own CUDA path (optimizer.py:39 picks cuda when torch sees one), run live on this box from
      baseline/_ref: cuBLAS / cuDNN instead of oneDNN, i.e. the reference's own cross-backend noise — plain fp32 and
      torch's stock TF32 settings.

What is asserted, per stage and iteration count:
  * (a) is not worse than (c, plain fp32): as many windows within 0.5 mm and as many strict windows (minus a margin
    for the coin tosses of 64 windows), 90 % quantile and maximum of the final-joint deviation in the same range;
  * every divergence of (a) sets in at an ill-conditioned `_cubic_interpolate` decision in the REFERENCE's own
    trace (tests/parity_stats.py) or late, as accumulated round-off — a closure that is simply too inaccurate shows
    up as "unexplained" (the reference's stock TF32 CUDA path does: 58-64 of 64 windows);
  * the first two evaluations agree for every window."""
import os

import numpy as np
import pytest

import parity_stats as ps
from globalegomocap_b200 import synthetic as syn

pytestmark = pytest.mark.gpu

W_LOCAL = (0.01 / 10000, 0.001 / 100, 0.01, 0.0, 0.01)
W_GLOBAL = (0.01, 0.001, 0.01, 0.0, 0)


@pytest.fixture(scope="module")
def setup(golden_dir):
    from globalegomocap_b200.engine import Engine
    g = np.load(os.path.join(golden_dir, "dist.npz"))
    clip = syn.make_clip(int(g["n_frames"]), seed=int(g["clip_seed"]))
    assert float(clip["heatmap_list"].astype(np.float64).sum()) == float(g["heat_checksum"])
    bias = syn.mean_pose_bias(clip)
    eng = Engine(max_windows=64)
    eng.set_camera(*syn.load_camera())
    eng.set_vae(0, syn.make_vae_state_dict(11, perturb_bn=True, pose_bias=bias))
    eng.set_vae(1, syn.make_vae_state_dict(12, perturb_bn=True, pose_bias=bias))
    yield g, clip, eng
    eng.close()


def _run_stage(eng, g, clip, max_iter, stage):
    from globalegomocap_b200.engine import energy_weights, lbfgs_params
    starts = [int(s) for s in g["starts"]]
    W = len(starts)
    params = lbfgs_params(max_iter=max_iter)
    mb = g["mean_bone_length"]
    if stage == 0:
        x0 = np.stack([clip["estimated_local_skeleton"][s:s + 10] for s in starts]).astype(np.float32)
        return eng.solve_stage(0, x0, clip["heatmap_list"], np.asarray(starts, np.int64), np.zeros(W, np.int32), mb,
                               g["eps"][:, 0], energy_weights(*W_LOCAL), params, want_trace=True)
    x0 = g[f"mi{max_iter}_rel_in"].astype(np.float32)       # both reference runs were anchored here too
    return eng.solve_stage(1, x0, None, None, np.zeros(W, np.int32), mb, g["eps"][:, 1], energy_weights(*W_GLOBAL),
                           params, want_trace=True)


@pytest.fixture(scope="module")
def reference_on_cuda(tmp_path_factory):
    """The unmodified reference on ITS OWN CUDA path (optimizer.py:39; baseline/_ref travels to the GPU box) over the
    same 64 windows, in a subprocess: plain fp32 (TF32 off) and torch's stock settings (cuDNN may use TF32).  None when
    the reference is not installed on this box."""
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(repo, "baseline", "_ref", "optimizer.py")):
        return None
    out = {}
    cache = os.environ.get("GEM_REF_CUDA_CACHE")           # optional directory: reuse the runs across pytest invocations
    for tag, tf32 in (("fp32", 0), ("stock", -1)):
        path = (os.path.join(cache, f"ref_cuda_{tag}.npz") if cache else
                str(tmp_path_factory.mktemp("refcuda") / f"ref_{tag}.npz"))
        if cache and os.path.exists(path):
            out[tag] = np.load(path)
            continue
        res = subprocess.run([sys.executable, os.path.join(repo, "tests", "reference_runs.py"), "--out", path, "--tf32", str(tf32)],
                             capture_output=True, text=True, timeout=1500)
        assert res.returncode == 0, res.stderr[-2000:]
        out[tag] = np.load(path)
        assert str(out[tag]["device"]) == "cuda"
    return out


def _against_ref1(name, g, max_iter, stage, E, ne, pose, cub, cols):
    E_ref, P_ref, ne_ref = g[f"mi{max_iter}_E"], g[f"mi{max_iter}_pose"], g[f"mi{max_iter}_n_eval"]
    mm, strict, kinds, leads = [], [], [], []
    for w in range(pose.shape[0]):
        lead, d, s = ps.compare_runs(E_ref[0, stage, w], ne_ref[0, stage, w], P_ref[0, stage, w], E[w], ne[w], pose[w])
        mm.append(d), strict.append(s), leads.append(lead)
        kinds.append(ps.classify(lead, int(ne_ref[0, stage, w]), int(ne[w]), cub[(stage, w)], cols,
                                 ps.onset_of(E_ref[0, stage, w], ne_ref[0, stage, w], E[w], ne[w])))
    return ps.summarize(name, mm, strict, kinds, np.asarray(ne).astype(int) - ne_ref[0, stage].astype(int)), leads


@pytest.mark.parametrize("max_iter", [3, 25])
def test_cuda_deviations_stay_within_the_references_self_noise(setup, reference_on_cuda, max_iter):
    g, clip, eng = setup
    cub, cols = ps.cubic_rows(g, max_iter)
    table, checks = [], []
    for stage in (0, 1):
        res = _run_stage(eng, g, clip, max_iter, stage)
        assert int(res["status"].sum()) == 0
        tr, pose, ev = res["trace"].cpu().numpy(), res["pose"].cpu().numpy(), res["func_evals"].cpu().numpy()
        if os.environ.get("GEM_DIST_DUMP"):                     # raw runs for offline analysis
            np.savez_compressed(os.path.join(os.environ["GEM_DIST_DUMP"], f"ours_mi{max_iter}_stage{stage}.npz"), E=tr, pose=pose,
                                n_eval=ev)
        ours, leads = _against_ref1("this library (CUDA) vs reference 1 CPU thread", g, max_iter, stage, tr, ev, pose, cub, cols)
        # the first two evaluations agree for every window (some global-stage solves stop after their first one)
        early = [(w, l) for w, l in enumerate(leads) if l < min(2, int(g[f"mi{max_iter}_n_eval"][0, stage, w]))]
        # the same with every layer on fp32 CUDA cores (gemm mode 0): the split-fp16 tensor-core operands carry 22-23
        # significant bits, fp32 24 — the one systematic difference to the reference's arithmetic
        eng.set_gemm_mode(0)
        try:
            res0 = _run_stage(eng, g, clip, max_iter, stage)
        finally:
            eng.set_gemm_mode(3)
        simt, _ = _against_ref1("this library, fp32 CUDA-core layers (gemm mode 0) vs reference 1 CPU thread", g, max_iter, stage,
                                res0["trace"].cpu().numpy(), res0["func_evals"].cpu().numpy(), res0["pose"].cpu().numpy(), cub, cols)
        rows = [ours, simt, ps.reference_self_noise(g, max_iter, stage)]
        yard = None
        if reference_on_cuda is not None:
            for tag, label in (("fp32", "reference on CUDA (fp32) vs reference 1 CPU thread"),
                               ("stock", "reference on CUDA (stock TF32 settings) vs reference 1 CPU thread")):
                r = reference_on_cuda[tag]
                row, _ = _against_ref1(label, g, max_iter, stage, r[f"mi{max_iter}_E"][stage], r[f"mi{max_iter}_n_eval"][stage],
                                       r[f"mi{max_iter}_pose"][stage], cub, cols)
                rows.append(row)
            yard = rows[3]                                      # the strict yardstick: the reference's own fp32 CUDA path
        table += rows
        checks.append((stage, ours, simt, rows[2], yard, early))
    print("\nmax_iter", max_iter, "(local stage rows, then global stage rows)\n" + ps.format_table(table))
    for stage, ours, simt, ref_threads, yard, early in checks:
        assert not early, (max_iter, stage, early)
        # every divergence sets in at an ill-conditioned interpolation of the reference's own trace (or late, as accumulated
        # round-off); the reference's CUDA path shows at most one exception per stage, so does this library
        assert ours["unexplained"] <= 2, (max_iter, stage, ours["unexplained"])
        if yard is None:
            continue
        # not worse than the reference's own fp32 CUDA path against its CPU path, up to the coin tosses of 64 windows
        # (measured: 40 / 24 / 17 / 22 strict windows here, 42 / 28 / 28 / 18 there; profiles/r02_parity_distribution.txt)
        assert ours["frac_within_0.5mm"] >= yard["frac_within_0.5mm"] - 0.25, (max_iter, stage, ours, yard)
        assert ours["strict"] >= yard["strict"] - 16, (max_iter, stage, ours["strict"], yard["strict"])
        # with fp32 CUDA-core layers the statistics are the reference-on-CUDA's (41 / 19 / 27 / 20 vs 42 / 28 / 28 / 18)
        assert simt["unexplained"] <= 2 and simt["strict"] >= yard["strict"] - 10, (max_iter, stage, simt["strict"], yard["strict"])
        assert simt["frac_within_0.5mm"] >= yard["frac_within_0.5mm"] - 0.17, (max_iter, stage, simt, yard)
        assert ours["max_mm"] <= 1.5 * max(yard["max_mm"], ref_threads["max_mm"]), (max_iter, stage, ours["max_mm"])
        assert ours["q90_mm"] <= max(2.0 * yard["q90_mm"], 0.05), (max_iter, stage, ours["q90_mm"], yard["q90_mm"])
        assert abs(ours["mean_abs_n_eval_diff"] - yard["mean_abs_n_eval_diff"]) <= 1.0
