import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from globalegomocap_b200 import synthetic as syn
from globalegomocap_b200.engine import Engine, energy_weights, lbfgs_params
from oracle import energy_np as en
from oracle.pipeline_np import StageSolver
from oracle import lbfgs_np
d = np.load('tests/golden/clip58.npz'); clip = {k: d[k] for k in d.files}
g = np.load('tests/golden/traces.npz')
bias = syn.mean_pose_bias(clip)
sd = syn.make_vae_state_dict(11, perturb_bn=True, pose_bias=bias)
cam = syn.load_camera(); mb = en.mean_bone_length(clip['estimated_local_skeleton'])
W_LOCAL = (0.01 / 10000, 0.001 / 100, 0.01, 0.0, 0.01)
eng = Engine(max_windows=8); eng.set_camera(*cam); eng.set_vae(0, sd)
wi, s = 1, 8
x0 = clip['estimated_local_skeleton'][s:s+10].astype(np.float32)
heat = clip['heatmap_list']
# drive the state machine by hand to see every trial point
z0, mu, std = eng.encode(0, x0[None], g['eps'][wi:wi+1, 0])
print('z0 diff vs ref', float(np.abs(z0.cpu().numpy()[0] - g['mi3_w1_local_z'][0]).max()))
eng.lbfgs_begin(z0, lbfgs_params(max_iter=3))
Zref = g['mi3_w1_local_z']; Eref = g['mi3_w1_local_E']
for k in range(5):
    st = eng.lbfgs_stats()
    if int(st['finished'][0]): break
    z = eng.lbfgs_trial().clone()
    pose = eng.decode(0, z)
    E, _, grad, _ = eng.energy_grad(pose, x0[None], heat, np.array([s]), np.zeros(1, np.int32), mb, energy_weights(*W_LOCAL))
    dz = eng.decode_vjp(0, grad)
    zr = Zref[k] if k < len(Zref) else None
    step = (z[0] - z0[0]).cpu().numpy()
    print(k, 'E', float(E[0]), 'Eref', Eref[k] if k < len(Eref) else None, 'max|z-zref|', None if zr is None else float(np.abs(z[0].cpu().numpy() - zr).max()),
          'step ratio vs ref', None if zr is None or k == 0 else float(np.abs(step).max() / max(np.abs(zr - Zref[0]).max(), 1e-30)), 't', float(st['t'][0]))
    eng.lbfgs_advance(E, dz)
# oracle decisions with events
sol = StageSolver(sd, cam, mb, W_LOCAL, max_iter=3)
pose, info = sol.solve(x0, heat[s:s+10], g['eps'][wi, 0])
print('oracle E', [t[0] for t in info['trace']])
print('oracle events', [e for e in info['events'] if e[1] == 'cubic_disc'])
orig = lbfgs_np._cubic_interpolate
