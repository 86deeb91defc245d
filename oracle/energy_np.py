"""numpy restatement of the reference's energy terms and their analytic
gradients (oracle: test infrastructure only, see oracle/__init__.py).

E = w3d*E_3d + ws*E_smooth + wb*E_bone + wv*E_vae + wr*E_reproj
(reference optimizer.py:226-240).  All functions take a pose ``x`` of shape
(T, 15, 3) and work in ``x.dtype`` (float32 to mimic the reference, float64 as
a tie-breaker).
"""
from __future__ import annotations

import numpy as np

PARENTS = np.array([0, 0, 1, 2, 0, 4, 5, 1, 7, 8, 9, 4, 11, 12, 13])  # optimizer.py:34


# ---------------------------------------------------------------- fisheye
def fisheye_project(p: np.ndarray, poly, cx, cy):
    """FishEyeCalibrated.py:96-129 (world2camera_pytorch): (n,3) -> (n,2), plus
    the intermediate values the Jacobian needs.  Raises like the reference when
    a point has x = y = 0."""
    dt = p.dtype
    x, y = p[:, 0], p[:, 1]
    zn = -p[:, 2]
    r = np.sqrt(x * x + y * y)
    if not (r != 0).all():
        raise Exception("norm is zero!")
    theta = np.arctan(zn / r)
    c = np.asarray(poly, dtype=np.float64).astype(dt)          # python scalars are cast to the tensor dtype
    rho = np.full_like(theta, c[0])
    drho = np.zeros_like(theta)
    t_i = np.ones_like(theta)                                   # theta^(i-1) when used for drho
    for i in range(1, len(c)):
        drho = drho + dt.type(i) * c[i] * t_i
        t_i = t_i * theta
        rho = rho + t_i * c[i]
    inv = dt.type(1.0) / r
    u = x * inv * rho + dt.type(cx)
    v = y * inv * rho + dt.type(cy)
    return np.stack([u, v], axis=1), (r, theta, rho, drho)


def fisheye_jacobian(p: np.ndarray, aux):
    """d(u,v)/d(x,y,z) per point -> (n,2,3); SURVEY.md A.5."""
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    r, theta, rho, drho = aux
    r2 = r * r
    q = r2 + z * z
    dth_dx = z * x / (r * q)
    dth_dy = z * y / (r * q)
    dth_dz = -r / q
    r3 = r2 * r
    J = np.empty((p.shape[0], 2, 3), dtype=p.dtype)
    J[:, 0, 0] = rho / r - x * x * rho / r3 + (x / r) * drho * dth_dx
    J[:, 0, 1] = -x * y * rho / r3 + (x / r) * drho * dth_dy
    J[:, 0, 2] = (x / r) * drho * dth_dz
    J[:, 1, 0] = -x * y * rho / r3 + (y / r) * drho * dth_dx
    J[:, 1, 1] = rho / r - y * y * rho / r3 + (y / r) * drho * dth_dy
    J[:, 1, 2] = (y / r) * drho * dth_dz
    return J


# ---------------------------------------------------------------- bilinear
def bilinear_zeros(maps: np.ndarray, ix: np.ndarray, iy: np.ndarray):
    """grid_sample(bilinear, padding_mode='zeros', align_corners=True) for one
    sample per map (optimizer.py:147).  maps (n,H,W); ix, iy (n,) pixel coords.
    Returns (S, dS/dix, dS/diy)."""
    n, H, W = maps.shape
    dt = ix.dtype
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    wx1, wx0 = ix - x0, x1 - ix          # weight of column x1 / x0
    wy1, wy0 = iy - y0, y1 - iy
    xi0, yi0 = x0.astype(np.int64), y0.astype(np.int64)
    xi1, yi1 = xi0 + 1, yi0 + 1
    idx = np.arange(n)

    def tex(yy, xx):
        ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
        val = maps[idx, np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(dt)
        return np.where(ok, val, dt.type(0))

    nw, ne, sw, se = tex(yi0, xi0), tex(yi0, xi1), tex(yi1, xi0), tex(yi1, xi1)
    S = nw * (wx0 * wy0) + ne * (wx1 * wy0) + sw * (wx0 * wy1) + se * (wx1 * wy1)
    dS_dix = -nw * wy0 + ne * wy0 - sw * wy1 + se * wy1
    dS_diy = -nw * wx0 - ne * wx1 + sw * wx0 + se * wx1
    return S, dS_dix, dS_diy


# ---------------------------------------------------------------- terms
def e3d(x, x0):
    """pose_energy_3d, optimizer.py:210-213."""
    d = x - x0
    return (d * d).sum(dtype=x.dtype), 2 * d


def esmooth(x):
    """smooth_accelerate, optimizer.py:202-208."""
    v = x[:-1] - x[1:]
    a = v[:-1] - v[1:]                    # a_t = x_t - 2 x_{t+1} + x_{t+2}
    g = np.zeros_like(x)
    g[:-2] += 2 * a
    g[1:-1] -= 4 * a
    g[2:] += 2 * a
    return (a * a).sum(dtype=x.dtype), g


def ebone(x, mean_bone):
    """bone_length_energy + calculate_bone_length, optimizer.py:172-177, 89-94."""
    b = x - x[:, PARENTS, :]
    ell = np.sqrt((b * b).sum(-1))
    diff = ell - mean_bone[None, :].astype(x.dtype)
    safe = np.where(ell > 0, ell, 1)
    coef = np.where(ell > 0, 2 * diff / safe, 0)[..., None] * b       # torch.norm subgradient at 0 is 0
    g = coef.copy()
    np.subtract.at(g, (slice(None), PARENTS), coef)
    return (diff * diff).sum(dtype=x.dtype), g


def evae(x):
    """vae_energy applied to the decoded pose, optimizer.py:215-218, 238."""
    return (x * x).sum(dtype=x.dtype), 2 * x


def ereproj(x, heat_hwc, poly, cx, cy):
    """reprojection_energy_heatmap_fast, optimizer.py:139-149.  heat_hwc is the
    window's (T,H,W,15) heatmap block as stored in the pickle; map index t*15+j
    (optimizer.py:251-252)."""
    T, J, _ = x.shape
    dt = x.dtype
    H, W = heat_hwc.shape[1], heat_hwc.shape[2]
    maps = np.transpose(heat_hwc, (0, 3, 1, 2)).reshape(T * J, H, W)
    p = x.reshape(-1, 3)
    uv, aux = fisheye_project(p, poly, cx, cy)
    gx = ((uv[:, 0] - dt.type(128)) - dt.type(512)) / dt.type(512)
    gy = (uv[:, 1] - dt.type(512)) / dt.type(512)
    ix = ((gx + 1) / 2) * dt.type(W - 1)
    iy = ((gy + 1) / 2) * dt.type(H - 1)
    S, dsx, dsy = bilinear_zeros(maps, ix, iy)
    Jm = fisheye_jacobian(p, aux)
    kx = dt.type((W - 1) / 2.0 / 512.0)
    ky = dt.type((H - 1) / 2.0 / 512.0)
    g = -(dsx * kx)[:, None] * Jm[:, 0, :] - (dsy * ky)[:, None] * Jm[:, 1, :]
    return -S.sum(dtype=dt), g.reshape(T, J, 3)


def total_energy(x, x0, heat_hwc, mean_bone, weights, poly, cx, cy):
    """weights = (w3d, ws, wb, wv, wr).  Returns (E, dE/dx, per-term energies[5]).
    E_reproj is skipped when wr == 0 (optimizer.py:232-235)."""
    w3d, ws, wb, wv, wr = [x.dtype.type(w) for w in weights]
    e1, g1 = e3d(x, x0.astype(x.dtype))
    e2, g2 = esmooth(x)
    e3, g3 = ebone(x, mean_bone)
    e4, g4 = evae(x)
    if wr != 0:
        e5, g5 = ereproj(x, heat_hwc, poly, cx, cy)
    else:
        e5, g5 = x.dtype.type(0), np.zeros_like(x)
    E = w3d * e1 + ws * e2 + wb * e3 + wv * e4 + wr * e5
    G = w3d * g1 + ws * g2 + wb * g3 + wv * g4 + wr * g5
    return E, G, np.array([e1, e2, e3, e4, e5], dtype=x.dtype)


def mean_bone_length(skeleton):
    """BodyPoseOptimizer.__init__, optimizer.py:42-43: mean over all frames of
    the clip's local estimate (cast to fp32 first, optimizer.py:333)."""
    s = np.asarray(skeleton).astype(np.float32).reshape(-1, 15, 3)
    b = s - s[:, PARENTS, :]
    return np.sqrt((b * b).sum(-1)).mean(0, dtype=np.float32)
