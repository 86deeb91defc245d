"""CPU oracle for the GlobalEgoMocap pose-sequence optimiser hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``globalegomocap_b200``; the only callers are ``tests/``,
``__graft_entry__.smoke()`` (as the checker) and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs.  The product path has no CPU fallback and raises when
the CUDA extension is missing.

What it is: a restatement, in numpy (and, for the timed CPU baseline, in plain
PyTorch CPU ops), of the reference's algorithm for this path — each function
cites the reference file:line it follows.  The reference is pure Python, so
there is no ``oracle/_ref`` binary to build; instead the restatement is PINNED
against golden vectors produced by importing and running the unmodified
reference in the CPU container (``tests/golden/make_golden.py`` → committed
``tests/golden/*.npz``): per-term energies and autograd gradients, VAE
encode/decode/VJP, per-evaluation L-BFGS traces and ``optimizer.main`` end to
end.  ``tests/test_oracle_*.py`` hold those checks.  The reference itself ships
no tests, golden vectors or known-answer fixtures for this path.

Third-party arithmetic the reference calls for this path and that is restated
here: ``torch.optim.LBFGS`` with ``strong_wolfe`` (torch 2.11.0,
``torch/optim/lbfgs.py``; restated in ``lbfgs_np.py`` and checked against the
installed torch on toy problems and on the reference's own traces),
``torch.nn.functional.grid_sample`` (bilinear, zeros padding,
align_corners=True), ``scipy.ndimage.gaussian_filter1d`` (scipy 1.18.1;
restated in ``pipeline_np.gaussian_filter1d_reflect``).

``lift_np.py`` restates the step before the optimiser (heat-map argmax + depth -> local
skeleton, reference ``utils/skeleton.py`` / ``FishEyeCalibrated.camera2world``) and is pinned the same way
(``tests/golden/make_golden_lift.py`` -> ``lift.npz``, ``tests/test_oracle_lift.py``).
"""
