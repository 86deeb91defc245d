"""Batched L-BFGS state machine and the full stage solve on the GPU vs torch.optim.LBFGS, the
CPU oracle and the reference's recorded traces.  Needs a B200: `pytest -m gpu`."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

W_LOCAL = (0.01 / 10000, 0.001 / 100, 0.01, 0.0, 0.01)
W_GLOBAL = (0.01, 0.001, 0.01, 0.0, 0)


def rosenbrock(x):
    return (100 * (x[..., 1:] - x[..., :-1] ** 2) ** 2 + (1 - x[..., :-1]) ** 2).sum(-1)


def quartic(x):
    i = torch.arange(1, x.shape[-1] + 1, dtype=x.dtype, device=x.device)
    return (i * (x - 0.5) ** 2).sum(-1) + 0.1 * (x ** 4).sum(-1) + torch.sin(3 * x).sum(-1)


def flat(x):
    return 1e-3 * (x ** 2).sum(-1)


def _f64_closure(fn, x32):
    """loss/grad evaluated in float64 from fp32 points and rounded once: identical values on any
    device, so only the optimiser's own arithmetic differs between the two sides."""
    x = x32.detach().double().requires_grad_(True)
    f = fn(x)
    (g,) = torch.autograd.grad(f.sum(), x)
    return f.float(), g.float()


def _torch_reference(fn, x0, max_iter):
    x = torch.nn.Parameter(torch.tensor(x0, dtype=torch.float32))
    opt = torch.optim.LBFGS([x], lr=2, max_iter=max_iter, tolerance_change=1e-6, line_search_fn="strong_wolfe")
    trace = []

    def closure():
        opt.zero_grad()
        f, g = _f64_closure(fn, x.data[None])
        x.grad = g[0].clone()
        trace.append((float(f[0]), x.detach().clone().numpy()))
        return f[0]

    opt.step(closure)
    st = opt.state[x]
    return x.detach().numpy(), trace, st["n_iter"], st["func_evals"]


@pytest.mark.parametrize("fn,n", [(rosenbrock, 16), (rosenbrock, 64), (quartic, 32), (quartic, 200), (flat, 8)])
@pytest.mark.parametrize("max_iter", [1, 3, 7, 25])
def test_state_machine_matches_torch_lbfgs(fn, n, max_iter):
    from globalegomocap_b200.engine import Engine, lbfgs_params
    W = 5
    rng = np.random.default_rng(n + max_iter)
    x0 = rng.standard_normal((W, n)).astype(np.float32)
    eng = Engine(max_windows=W, latent_dim=n, max_history=max(max_iter - 1, 1))
    params = lbfgs_params(max_iter=max_iter)
    eng.lbfgs_begin(x0, params)
    traces = [[] for _ in range(W)]
    for _ in range(params.max_eval + 1):
        st = eng.lbfgs_stats()
        active = (st["finished"] == 0).cpu().numpy()
        if not active.any():
            break
        z = eng.lbfgs_trial().clone()
        f, g = _f64_closure(fn, z)
        for w in range(W):
            if active[w]:
                traces[w].append((float(f[w]), z[w].cpu().numpy()))
        eng.lbfgs_advance(f, g)
    st = {k: v.cpu().numpy() for k, v in eng.lbfgs_stats().items()}
    assert st["finished"].all()
    xs = eng.lbfgs_x().cpu().numpy()
    for w in range(W):
        xt, trace_t, n_iter_t, evals_t = _torch_reference(fn, x0[w], max_iter)
        assert st["n_iter"][w] == n_iter_t, (w, st["n_iter"][w], n_iter_t)
        assert st["func_evals"][w] == evals_t == len(traces[w]), (w, st["func_evals"][w], evals_t, len(traces[w]))
        # fp32 round-off (dot-product order) is amplified by every line-search decision; the first
        # dozen evaluations must agree tightly, the tail of a 25-iteration run loosely
        for k, ((ft, zt), (fg, zg)) in enumerate(zip(trace_t[:12], traces[w][:12])):
            np.testing.assert_allclose(zg, zt, rtol=2e-3, atol=2e-4)
            assert abs(ft - fg) <= 2e-3 * max(1.0, abs(ft))
        if len(trace_t) <= 12:
            np.testing.assert_allclose(xs[w], xt, rtol=2e-3, atol=2e-4)
        else:   # both ended in a minimum of comparable depth
            assert traces[w][-1][0] <= trace_t[-1][0] + 0.05 * max(1.0, abs(trace_t[-1][0]))
    eng.close()


@pytest.fixture(scope="module")
def engines(vae_weights, vae_weights_g2, camera):
    from globalegomocap_b200.engine import Engine
    out = {}
    for tag, wts in (("g1", vae_weights), ("g2", vae_weights_g2)):
        eng = Engine(max_windows=64)
        eng.set_camera(*camera)
        eng.set_vae(0, wts[0])
        eng.set_vae(1, wts[1])
        out[tag] = eng
    yield out
    for e in out.values():
        e.close()


def test_teacher_forced_closure_matches_reference_and_oracle(engines, golden_dir, clip58, vae_weights, camera):
    """decode -> fused energy/gradient -> decoder bwd-data at the reference's own z for every
    recorded evaluation of the max_iter=25 runs: energy within 1e-4 relative of the reference
    (north-star bar; measured ~1e-6), dE/dz within 1e-3 of the fp64-accumulating oracle."""
    from globalegomocap_b200.engine import energy_weights
    from oracle import energy_np as en
    from oracle.pipeline_np import StageSolver
    eng = engines["g1"]
    g = np.load(os.path.join(golden_dir, "traces.npz"))
    mb = en.mean_bone_length(clip58["estimated_local_skeleton"])
    heat_all = clip58["heatmap_list"]
    worst_e = 0.0
    for wi, s in enumerate(g["starts"]):
        s = int(s)
        for stage, which, wts in (("local", 0, W_LOCAL), ("global", 1, W_GLOBAL)):
            x0 = (clip58["estimated_local_skeleton"][s:s + 10] if stage == "local"
                  else g[f"mi25_w{wi}_local_relglobal"]).astype(np.float32)
            E_ref, Z = g[f"mi25_w{wi}_{stage}_E"], g[f"mi25_w{wi}_{stage}_z"]
            K = len(E_ref)
            pose = eng.decode(which, Z)
            E, _, grad, status = eng.energy_grad(pose, np.stack([x0] * K), heat_all, np.full(K, s, np.int64),
                                                 np.zeros(K, np.int32), mb, energy_weights(*wts))
            dz = eng.decode_vjp(which, grad).cpu().numpy()
            E = E.cpu().numpy()
            rel = np.abs(E - E_ref) / np.abs(E_ref)
            worst_e = max(worst_e, rel.max())
            assert rel.max() <= 1e-4, (wi, stage, rel.max())
            cl = StageSolver(vae_weights[which], camera, mb, wts).closure_for(x0, heat_all[s:s + 10])
            for k in (0, K // 2, K - 1):
                e_o, g_o = cl(Z[k])
                assert np.abs(dz[k] - g_o).max() <= 1e-3 * np.abs(g_o).max(), (wi, stage, k)
    print("worst teacher-forced relative energy error vs reference:", worst_e)


def test_replay_of_reference_closure_stream_reproduces_every_trial_point(golden_dir):
    """Teacher-forced OPTIMISER parity on the GPU: the batched state machine, fed the reference's
    exact (loss, gradient) stream for four 25-iteration solves at once (different lengths: the
    windows finish in different rounds), asks for exactly the points the reference evaluated."""
    from globalegomocap_b200.engine import Engine, lbfgs_params
    g = np.load(os.path.join(golden_dir, "traces_grad.npz"))
    keys = ("w1_local", "w1_global", "w2_local", "w2_global")
    E = [g[k + "_E"] for k in keys]
    Z = [g[k + "_z"] for k in keys]
    G = [g[k + "_g"] for k in keys]
    W = len(keys)
    eng = Engine(max_windows=W)
    params = lbfgs_params(max_iter=25)
    eng.lbfgs_begin(np.stack([z[0] for z in Z]), params)
    worst = 0.0
    for k in range(params.max_eval + 1):
        st = eng.lbfgs_stats()
        fin = st["finished"].cpu().numpy()
        z = eng.lbfgs_trial().cpu().numpy()
        f = np.zeros(W, np.float32)
        gr = np.zeros((W, 2048), np.float32)
        for w in range(W):
            if k < len(E[w]):
                assert not fin[w], (keys[w], k)
                scale = max(np.abs(Z[w][k] - Z[w][0]).max(), 1e-3)
                err = np.abs(z[w] - Z[w][k]).max() / scale
                worst = max(worst, err)
                assert err <= 2e-4 + 1e-6 / scale, (keys[w], k, err)
                f[w], gr[w] = E[w][k], G[w][k]
            else:
                assert fin[w], (keys[w], k)       # stopped after exactly the reference's number of evaluations
        eng.lbfgs_advance(f, gr)
    st = eng.lbfgs_stats()
    assert st["finished"].cpu().numpy().all()
    assert st["func_evals"].cpu().tolist() == [len(e) for e in E]
    print("worst trial-point deviation relative to the distance travelled:", worst)
    eng.close()


def agreement(E, E_ref, pose, pose_ref):
    E = np.asarray(E)[~np.isnan(E)]
    n = min(len(E), len(E_ref))
    rel = np.abs(E[:n] - np.asarray(E_ref[:n])) / np.abs(E_ref).max()
    lead = int(np.argmax(rel > 1e-4)) if (rel > 1e-4).any() else n
    err_mm = float(np.abs(pose - pose_ref).max() * 1000)
    return lead, err_mm, (lead == n and len(E) == len(E_ref) and err_mm < 0.5)


def test_free_running_stage_matches_reference(engines, golden_dir, clip58):
    """gem_solve_stage (encoder -> batched L-BFGS -> decode, both stages, 3 windows per launch)
    against the reference's recorded runs.  See tests/test_oracle_lbfgs.py for why agreement
    beyond the first evaluations is statistical: torch's cubic line-search step is ill-conditioned
    here and the reference itself does not reproduce across thread counts."""
    from globalegomocap_b200.engine import energy_weights, lbfgs_params
    from oracle import energy_np as en
    mb = en.mean_bone_length(clip58["estimated_local_skeleton"])
    heat_all = clip58["heatmap_list"]
    cases = []
    for tag, fname, iters in (("g1", "traces.npz", (1, 2, 3, 5, 25)), ("g2", "traces_g2.npz", (5, 25))):
        g = np.load(os.path.join(golden_dir, fname))
        eng = engines[tag]
        starts = [int(s) for s in g["starts"]]
        fb = np.asarray(starts, dtype=np.int64)
        for mi in iters:
            params = lbfgs_params(max_iter=mi)
            x0 = np.stack([clip58["estimated_local_skeleton"][s:s + 10] for s in starts]).astype(np.float32)
            res = eng.solve_stage(0, x0, heat_all, fb, np.zeros(3, np.int32), mb, g["eps"][:, 0],
                                  energy_weights(*W_LOCAL), params, want_trace=True)
            rel = np.stack([g[f"mi{mi}_w{wi}_local_relglobal"] for wi in range(3)]).astype(np.float32)
            resg = eng.solve_stage(1, rel, None, None, np.zeros(3, np.int32), mb, g["eps"][:, 1],
                                   energy_weights(*W_GLOBAL), params, want_trace=True)
            for stage, r in (("local", res), ("global", resg)):
                tr, pose, ev = r["trace"].cpu().numpy(), r["pose"].cpu().numpy(), r["func_evals"].cpu().numpy()
                assert int(r["status"].sum()) == 0
                for wi in range(3):
                    E_ref = g[f"mi{mi}_w{wi}_{stage}_E"]
                    assert np.isfinite(tr[wi, :ev[wi]]).all() and np.isnan(tr[wi, ev[wi]:]).all()
                    cases.append(((tag, mi, wi, stage),) + agreement(tr[wi], E_ref, pose[wi],
                                                                      g[f"mi{mi}_w{wi}_{stage}_pose"]))
    for key, lead, err_mm, strict in cases:
        print(key, "leading evals in agreement:", lead, "final joints mm: %.4f" % err_mm, "strict" if strict else "")
        assert lead >= 2, key
        if key[1] == 1:
            assert strict, key
    n_strict = sum(c[3] for c in cases)
    print("strict (energy 1e-4 at every evaluation, joints 0.5 mm):", n_strict, "of", len(cases))
    # 19 of 42 measured (round 2; which of the ill-conditioned cases flip moves by a few with every change of the
    # closure's last bits — the statistically meaningful statement is tests/test_gpu_dist.py on 256 solves)
    assert n_strict >= 18, (n_strict, len(cases))
    assert sum(1 for c in cases if c[0][1] == 25 and c[3]) >= 2


def test_stage_is_batch_independent(engines, golden_dir, clip58):
    """A window's whole solve is bit-identical wherever it sits in the batch and whatever its
    neighbours do (no cross-window arithmetic anywhere in the pipeline)."""
    from globalegomocap_b200.engine import energy_weights, lbfgs_params
    from oracle import energy_np as en
    eng = engines["g1"]
    g = np.load(os.path.join(golden_dir, "traces.npz"))
    mb = en.mean_bone_length(clip58["estimated_local_skeleton"])
    starts = [0, 8, 16, 0, 8, 0, 16]
    x0 = np.stack([clip58["estimated_local_skeleton"][s:s + 10] for s in starts]).astype(np.float32)
    eps = np.stack([g["eps"][{0: 0, 8: 1, 16: 2}[s], 0] for s in starts])
    res = eng.solve_stage(0, x0, clip58["heatmap_list"], np.asarray(starts, np.int64), np.zeros(len(starts), np.int32),
                          mb, eps, energy_weights(*W_LOCAL), lbfgs_params(max_iter=25), want_trace=True)
    pose, tr = res["pose"].cpu().numpy(), res["trace"].cpu().numpy()
    for a, b in ((0, 3), (0, 5), (1, 4), (2, 6)):
        assert (pose[a] == pose[b]).all()
        assert np.array_equal(tr[a], tr[b], equal_nan=True)
    ev = res["func_evals"].cpu().numpy()
    assert (ev >= 25).all() and (ev <= 32).all()
