// Fused energy + analytic gradient for W independent windows (one launch).
//
// Replaces, per closure evaluation, the reference's pose_energy_3d / smooth_accelerate /
// bone_length_energy / vae_energy / reprojection_energy_heatmap_fast (optimizer.py:139-149,
// 172-177, 202-218, 226-240), FishEyeCameraCalibrated.world2camera_pytorch
// (FishEyeCalibrated.py:96-129), grid_sample(bilinear, zeros, align_corners=True) and the
// autograd backward of all of them (~70 tiny ATen kernels + one host sync) by ONE kernel:
//   * a CTA owns two consecutive windows (2 x T x J joints, one thread per joint);
//   * the windows' pose and anchor tiles (2 x 1800 B each) are staged into shared memory by
//     the TMA engine (cp.async.bulk + mbarrier) and the gradient tile leaves the same way;
//   * the heatmaps are gathered in place from the pickle's HWC layout: 4 texels per joint,
//     never the 2.4 MB of maps per window;
//   * per-term energies are reduced with warp shuffles, combined in the reference's order.
// Compiled with --fmad=false: the pixel-coordinate chain must round like ATen's separate
// fp32 ops, because floor() of the pixel coordinate selects the texels.
#include <string.h>

#include "kernels.cuh"

namespace gem {

constexpr int kSlot = 160;             // threads per window (>= T*J, multiple of 32)
constexpr int kWinPerCta = 2;
constexpr int kThreads = kSlot * kWinPerCta;
constexpr int kSplitMax = 512;         // floats per window of the split gradient tile: T * pp <= 512

struct EnergyArgs {
    const float* pose;
    const float* pose0;
    const float* heat;
    const int64_t* frame_base;
    const int32_t* clip;
    const float* mean_bone;
    float* energy;
    float* terms;
    float* grad;
    uint32_t* status;
    int W, T, J, H, Wd;
    float w3d, ws, wb, wv, wr;
    int use_bulk;
    // optional: dE/dpose also as TF32 hi / lo parts in the token-major, zero-padded [W*T][pp] layout the
    // tensor-core bwd-data layers read (saves a separate split pass)
    float *gp_hi, *gp_lo;
    int pp;
    // gp_f16: the split is the fp16 scheme's instead (uint16 arrays in the same buffers): the window's gradient is
    // first scaled by the power of two that brings its largest entry to ~2^4 (gradients shrink to 1e-7 near
    // convergence, far below fp16's normal range); the exponent goes to row_exp[w] and is removed again, exactly,
    // by the epilogue of the T*256 -> latent GEMM at the end of the (linear) bwd-data chain
    int gp_f16;
    int32_t* row_exp;
    // optional texel cache (used when the heat maps stay in pinned HOST memory and are read over PCIe): per joint
    // an 8x8 window of its map, 256 contiguous bytes in HBM, the map coordinate of its corner and a valid bit per
    // texel.  The same few texels are read by every evaluation of a stage (joints move by a fraction of a texel per
    // step), so only the texels the optimiser actually samples cross the bus, once each while the joint stays in
    // its window; values are copies of the map's, so the energy is bit-identical with and without the cache.
    float* patch;              // [W][T*J][kPatchW * kPatchW]
    short2* patch_origin;      // [W][T*J]: map coordinate of the window's corner
    unsigned long long* patch_valid;   // [W][T*J]: one bit per texel of the window (0: empty window)
    unsigned long long* patch_stats;   // optional {lookups, texels fetched}
    // camera and skeleton of the calling ctx: they travel as kernel parameters (constant bank), so two ctxs with
    // different calibrations never see each other's (FishEyeCalibrated.py:8-14 is per optimiser in the reference too)
    CameraConst cam;
    SkeletonConst skel;
};

__device__ __forceinline__ float texel(const float* __restrict__ heat, int64_t frame, int y, int x, int j, int H,
                                       int Wd, int J) {
    if (x < 0 || x >= Wd || y < 0 || y >= H) return 0.f;     // padding_mode='zeros'
    return __ldg(heat + ((frame * H + y) * (int64_t)Wd + x) * J + j);
}


// Demand-fetched 8x8 window of one joint's map: origin and a valid bit per texel; only texels of the bilinear footprint
// (x0, y0) .. (x0+1, y0+1) that are not in the window yet are read from the map (over PCIe when it is host memory).
// A footprint that leaves the window re-centres it (and empties it).
__device__ __forceinline__ void cache_lookup(const EnergyArgs& a, size_t pk, int64_t frame, int j, int x0, int y0,
                                             bool count_lookup, float& nw, float& ne, float& sw, float& se) {
    float* pe = a.patch + pk * (kPatchW * kPatchW);
    const short2 o = a.patch_origin[pk];
    unsigned long long valid = a.patch_valid[pk];
    int ox = o.x, oy = o.y;
    int dx = x0 - ox, dy = y0 - oy;
    if (valid == 0ull || dx < 0 || dx > kPatchW - 2 || dy < 0 || dy > kPatchW - 2)
        ox = x0 - (kPatchW / 2 - 1), oy = y0 - (kPatchW / 2 - 1), dx = dy = kPatchW / 2 - 1, valid = 0ull;
    const int b00 = dy * kPatchW + dx;
    const unsigned long long foot = (3ull | (3ull << kPatchW)) << b00;
    const unsigned long long missing = foot & ~valid;
    if (a.patch_stats) {
        if (count_lookup) atomicAdd(a.patch_stats, 1ull);
        if (missing) atomicAdd(a.patch_stats + 1, (unsigned long long)__popcll(missing));
    }
    float* p0 = pe + b00;
    if (missing) {                            // up to four independent loads in flight
        const bool m0 = (missing >> b00) & 1ull, m1 = (missing >> (b00 + 1)) & 1ull;
        const bool m2 = (missing >> (b00 + kPatchW)) & 1ull, m3 = (missing >> (b00 + kPatchW + 1)) & 1ull;
        nw = m0 ? texel(a.heat, frame, y0, x0, j, a.H, a.Wd, a.J) : p0[0];
        ne = m1 ? texel(a.heat, frame, y0, x0 + 1, j, a.H, a.Wd, a.J) : p0[1];
        sw = m2 ? texel(a.heat, frame, y0 + 1, x0, j, a.H, a.Wd, a.J) : p0[kPatchW];
        se = m3 ? texel(a.heat, frame, y0 + 1, x0 + 1, j, a.H, a.Wd, a.J) : p0[kPatchW + 1];
        if (m0) p0[0] = nw;
        if (m1) p0[1] = ne;
        if (m2) p0[kPatchW] = sw;
        if (m3) p0[kPatchW + 1] = se;
        a.patch_origin[pk] = make_short2((short)ox, (short)oy);
        a.patch_valid[pk] = valid | missing;
    } else {
        nw = p0[0], ne = p0[1], sw = p0[kPatchW], se = p0[kPatchW + 1];
    }
}

// Fisheye projection of a joint and the map cell its bilinear footprint starts at (FishEyeCalibrated.py:96-129,
// optimizer.py:139-149); the arithmetic (and this translation unit's --fmad=false) is what decides the cell, so the
// energy kernel and the texel prefetch kernel share it.  Returns false when r == 0.
struct Proj {
    float r, inv, rho, drho, ix, iy, fx0, fy0;
};
__device__ __forceinline__ bool project_joint(const CameraConst& c_cam, float x, float y, float z, int H, int Wd, Proj& p) {
    const float zn = -z;
    p.r = sqrtf(x * x + y * y);
    if (p.r == 0.f) return false;
    const float theta = atanf(zn / p.r);
    float rho = c_cam.poly[0], drho = 0.f, ti = 1.f;
    for (int i = 1; i < c_cam.n_poly; ++i) {          // power accumulation, not Horner
        drho += (float)i * c_cam.poly[i] * ti;
        ti *= theta;
        rho += ti * c_cam.poly[i];
    }
    p.rho = rho, p.drho = drho;
    p.inv = 1.0f / p.r;
    const float u = x * p.inv * rho + c_cam.cx;
    const float v = y * p.inv * rho + c_cam.cy;
    // pose_2d[:,0] -= 128; (pose_2d - 512)/512; grid_sample unnormalise (align_corners)
    const float gxn = ((u - 128.f) - 512.f) / 512.f;
    const float gyn = (v - 512.f) / 512.f;
    p.ix = ((gxn + 1.f) / 2.f) * (float)(Wd - 1);
    p.iy = ((gyn + 1.f) / 2.f) * (float)(H - 1);
    p.fx0 = floorf(p.ix), p.fy0 = floorf(p.iy);
    return true;
}

// Zero-copy maps: fetches the texels the next energy evaluation will sample into the joints' cache windows.  A
// handful of CTAs (grid-stride over all joints) so that the wait for PCIe occupies a few SMs instead of every SM:
// energy CTAs stalled on PCIe reads are small but sit on all SMs, and no 200 KB tensor-core CTA of another slice can
// be scheduled beside them — the solve then serialises with the transfers instead of overlapping them.
__global__ void __launch_bounds__(512) texel_prefetch_kernel(const __grid_constant__ EnergyArgs a) {
    const int TJ = a.T * a.J;
    const size_t total = (size_t)a.W * TJ;
    for (size_t i = (size_t)blockIdx.x * 512 + threadIdx.x; i < total; i += (size_t)gridDim.x * 512) {
        const int w = (int)(i / TJ), k = (int)(i - (size_t)w * TJ);
        const int t = k / a.J, j = k - t * a.J;
        const float x = a.pose[i * 3 + 0], y = a.pose[i * 3 + 1], z = a.pose[i * 3 + 2];
        Proj p;
        if (!project_joint(a.cam, x, y, z, a.H, a.Wd, p)) continue;
        if (!(p.fx0 >= -1.f && p.fx0 <= (float)a.Wd && p.fy0 >= -1.f && p.fy0 <= (float)a.H)) continue;
        float nw, ne, sw, se;
        cache_lookup(a, i, a.frame_base[w] + t, j, (int)p.fx0, (int)p.fy0, false, nw, ne, sw, se);
    }
}

__global__ void __launch_bounds__(kThreads) energy_grad_kernel(const __grid_constant__ EnergyArgs a) {
    __shared__ __align__(16) float s_x[kWinPerCta * kSlot * 3];
    __shared__ __align__(16) float s_x0[kWinPerCta * kSlot * 3];
    __shared__ __align__(16) float s_g[kWinPerCta * kSlot * 3];
    __shared__ float s_red[kThreads / 32][6];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(16) float s_gh[kWinPerCta * kSplitMax];
    __shared__ __align__(16) float s_gl[kWinPerCta * kSplitMax];

    const int tid = threadIdx.x;
    const int TJ = a.T * a.J;
    const int n = TJ * 3;                         // floats per window
    const int w0 = blockIdx.x * kWinPerCta;
    const int nwin = min(kWinPerCta, a.W - w0);
    const bool bulk = a.use_bulk && nwin == kWinPerCta;

    // ---- stage pose / anchor tiles -----------------------------------------------------------
    if (bulk) {
        // windows are contiguous in memory: one bulk copy covers both (2*n*4 bytes, 16-B multiple)
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            fence_barrier_init();
        }
        __syncthreads();
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)(kWinPerCta * n * sizeof(float));
            mbar_arrive_expect_tx(&s_bar, 2 * bytes);
            bulk_g2s(s_x, a.pose + (size_t)w0 * n, bytes, &s_bar);
            bulk_g2s(s_x0, a.pose0 + (size_t)w0 * n, bytes, &s_bar);
        }
        mbar_wait(&s_bar, 0);
    } else {
        for (int i = tid; i < nwin * n; i += kThreads) {
            s_x[i] = a.pose[(size_t)w0 * n + i];
            s_x0[i] = a.pose0[(size_t)w0 * n + i];
        }
        __syncthreads();
    }
    // with the bulk path both windows sit back to back (stride n); keep the same packing otherwise
    const int wl = tid / kSlot;                   // window slot inside the CTA
    const int k = tid - wl * kSlot;               // joint-frame index t*J + j
    const bool active = wl < nwin && k < TJ;
    const int w = w0 + wl;
    const float* X = s_x + wl * n;
    const float* X0 = s_x0 + wl * n;

    float e3d = 0.f, esm = 0.f, ebn = 0.f, eva = 0.f, erp = 0.f;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (active) {
        const int t = k / a.J, j = k - t * a.J;
        const float x = X[k * 3 + 0], y = X[k * 3 + 1], z = X[k * 3 + 2];

        // E_reproj, first half: the projection and the four texel loads are issued before the other terms so that
        // their DRAM latency overlaps that arithmetic (the terms are still added to the gradient in the old order)
        bool rp = false;
        Proj pj;
        float nw = 0.f, ne = 0.f, sw = 0.f, se = 0.f;
        if (a.wr != 0.f) {
            if (!project_joint(a.cam, x, y, z, a.H, a.Wd, pj)) {
                if (a.status) atomicOr(a.status + w, GEM_WIN_NORM_ZERO);
            } else if (pj.fx0 >= -1.f && pj.fx0 <= (float)a.Wd && pj.fy0 >= -1.f && pj.fy0 <= (float)a.H) {
                // (anything further than one texel outside contributes exactly 0)
                rp = true;
                const int x0 = (int)pj.fx0, y0 = (int)pj.fy0;
                const int64_t frame = a.frame_base[w] + t;
                if (a.patch) {
                    cache_lookup(a, (size_t)w * TJ + k, frame, j, x0, y0, true, nw, ne, sw, se);
                } else {
                    nw = texel(a.heat, frame, y0, x0, j, a.H, a.Wd, a.J);
                    ne = texel(a.heat, frame, y0, x0 + 1, j, a.H, a.Wd, a.J);
                    sw = texel(a.heat, frame, y0 + 1, x0, j, a.H, a.Wd, a.J);
                    se = texel(a.heat, frame, y0 + 1, x0 + 1, j, a.H, a.Wd, a.J);
                }
            }
        }

        // E_3d = sum (x - x0)^2                                     optimizer.py:210-213
        {
            const float dx = x - X0[k * 3 + 0], dy = y - X0[k * 3 + 1], dz = z - X0[k * 3 + 2];
            e3d = dx * dx + dy * dy + dz * dz;
            gx += a.w3d * (2.f * dx), gy += a.w3d * (2.f * dy), gz += a.w3d * (2.f * dz);
        }
        // E_vae = sum x^2 on the decoded pose                        optimizer.py:215-218,238
        {
            eva = x * x + y * y + z * z;
            gx += a.wv * (2.f * x), gy += a.wv * (2.f * y), gz += a.wv * (2.f * z);
        }
        // E_smooth: a_s = (x_s - x_{s+1}) - (x_{s+1} - x_{s+2}), s = 0..T-3   optimizer.py:202-208
        {
            float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 3; ++q) {        // q = 0: a_t (coef 1), 1: a_{t-1} (coef -2), 2: a_{t-2} (coef 1)
                const int s = t - q;
                if (s < 0 || s > a.T - 3) continue;
                const float coef = (q == 1) ? -2.f : 1.f;
                const float* p0 = X + ((s)*a.J + j) * 3;
                const float* p1 = X + ((s + 1) * a.J + j) * 3;
                const float* p2 = X + ((s + 2) * a.J + j) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float ac = (p0[c] - p1[c]) - (p1[c] - p2[c]);
                    acc[c] += coef * ac;
                    if (q == 0) esm += ac * ac;
                }
            }
            gx += a.ws * (2.f * acc[0]), gy += a.ws * (2.f * acc[1]), gz += a.ws * (2.f * acc[2]);
        }
        // E_bone = sum (|x_j - x_parent| - Lbar_j)^2                 optimizer.py:172-177, 89-94
        {
            const float* mb = a.mean_bone + (size_t)a.clip[w] * a.J;
            const int p = a.skel.parent[j];
            const float bx = x - X[(t * a.J + p) * 3 + 0], by = y - X[(t * a.J + p) * 3 + 1],
                        bz = z - X[(t * a.J + p) * 3 + 2];
            const float len = sqrtf(bx * bx + by * by + bz * bz);
            const float diff = len - mb[j];
            ebn = diff * diff;
            if (len > 0.f) {                        // torch.norm's subgradient at 0 is 0
                const float c = a.wb * (2.f * diff) / len;
                gx += c * bx, gy += c * by, gz += c * bz;
            }
            for (int ci = a.skel.child_start[j]; ci < a.skel.child_start[j + 1]; ++ci) {   // this joint as a parent
                const int cj = a.skel.child_list[ci];
                const float cx = X[(t * a.J + cj) * 3 + 0] - x, cy = X[(t * a.J + cj) * 3 + 1] - y,
                            cz = X[(t * a.J + cj) * 3 + 2] - z;
                const float cl = sqrtf(cx * cx + cy * cy + cz * cz);
                if (cl > 0.f) {
                    const float c = a.wb * (2.f * (cl - mb[cj])) / cl;
                    gx -= c * cx, gy -= c * cy, gz -= c * cz;
                }
            }
        }
        // E_reproj = -sum bilinear(H_tj; pix(project(x)))            optimizer.py:139-149
        if (rp) {
            {
                const float r = pj.r, inv = pj.inv, rho = pj.rho, drho = pj.drho;
                const float ix = pj.ix, iy = pj.iy, fx0 = pj.fx0, fy0 = pj.fy0;
                {
                    const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix;
                    const float wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
                    erp = -(nw * (wx0 * wy0) + ne * (wx1 * wy0) + sw * (wx0 * wy1) + se * (wx1 * wy1));
                    const float ds_dix = -nw * wy0 + ne * wy0 - sw * wy1 + se * wy1;
                    const float ds_diy = -nw * wx0 - ne * wx1 + sw * wx0 + se * wx1;
                    const float du = ds_dix * ((float)(a.Wd - 1) * 0.5f / 512.f);   // dS/du
                    const float dv = ds_diy * ((float)(a.H - 1) * 0.5f / 512.f);    // dS/dv
                    // fisheye Jacobian.  Only the energy's forward chain has to round like ATen's
                    // ops; the gradient is held to 1e-4, so one reciprocal replaces the dozen IEEE divisions.
                    const float r2 = r * r, q = r2 + z * z;
                    const float inv_q = __frcp_rn(q), inv_rq = inv * inv_q;
                    const float dth_dx = z * x * inv_rq, dth_dy = z * y * inv_rq, dth_dz = -r * inv_q;
                    const float xr = x * inv, yr = y * inv;
                    const float rho_r = rho * inv, rho_r3 = rho_r * inv * inv;
                    const float xd = xr * drho, yd = yr * drho;
                    const float du_dx = rho_r - x * x * rho_r3 + xd * dth_dx;
                    const float du_dy = -x * y * rho_r3 + xd * dth_dy;
                    const float du_dz = xd * dth_dz;
                    const float dv_dx = -x * y * rho_r3 + yd * dth_dx;
                    const float dv_dy = rho_r - y * y * rho_r3 + yd * dth_dy;
                    const float dv_dz = yd * dth_dz;
                    gx -= a.wr * (du * du_dx + dv * dv_dx);
                    gy -= a.wr * (du * du_dy + dv * dv_dy);
                    gz -= a.wr * (du * du_dz + dv * dv_dz);
                }
            }
        }
    }
    if (wl < kWinPerCta && k < TJ) {
        float* G = s_g + wl * n;
        G[k * 3 + 0] = gx, G[k * 3 + 1] = gy, G[k * 3 + 2] = gz;
        if (a.gp_hi && !a.gp_f16) {
            const int t = k / a.J, j = k - t * a.J;
            const int o = wl * a.T * a.pp + t * a.pp + j * 3;
            const float g3[3] = {gx, gy, gz};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float h = __uint_as_float(__float_as_uint(g3[c]) & 0xFFFFE000u);
                s_gh[o + c] = h, s_gl[o + c] = g3[c] - h;
            }
            if (j == 0)
                for (int c = a.J * 3; c < a.pp; ++c) s_gh[wl * a.T * a.pp + t * a.pp + c] = 0.f, s_gl[wl * a.T * a.pp + t * a.pp + c] = 0.f;
        }
    }

    // ---- per-window energy reduction: shuffle inside each warp, then kSlot/32 partials --------
    const float r0 = warp_sum(e3d), r1 = warp_sum(esm), r2 = warp_sum(ebn), r3 = warp_sum(eva), r4 = warp_sum(erp);
    const float r5 = warp_max(fmaxf(fabsf(gx), fmaxf(fabsf(gy), fabsf(gz))));      // (inactive threads hold zeros)
    if ((tid & 31) == 0) {
        float* d = s_red[tid >> 5];
        d[0] = r0, d[1] = r1, d[2] = r2, d[3] = r3, d[4] = r4, d[5] = r5;
    }
    if (bulk) fence_proxy_async_smem();           // make s_g visible to the bulk-copy engine
    __syncthreads();

    if (tid < kWinPerCta && tid < nwin) {
        constexpr int wpw = kSlot / 32;
        float t5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int q = 0; q < wpw; ++q)
#pragma unroll
            for (int c = 0; c < 5; ++c) t5[c] += s_red[tid * wpw + q][c];
        const int ww = w0 + tid;
        if (a.terms) {
#pragma unroll
            for (int c = 0; c < 5; ++c) a.terms[(size_t)ww * 5 + c] = t5[c];
        }
        // optimizer.py:239-240 (left to right; the reproj product is skipped when its weight is 0)
        float E = a.w3d * t5[0] + a.ws * t5[1] + a.wb * t5[2] + a.wv * t5[3];
        if (a.wr != 0.f) E += a.wr * t5[4];
        a.energy[ww] = E;
    }
    int ns = a.T * a.pp;                           // floats per window of the split tile
    if (a.gp_hi && a.gp_f16) {
        // fp16 scheme: scale the window by 2^e (max entry -> ~2^4), split, stage as uint16
        constexpr int wpw = kSlot / 32;
        if (wl < kWinPerCta && k < TJ) {
            float mx = 0.f;
            for (int q = 0; q < wpw; ++q) mx = fmaxf(mx, s_red[wl * wpw + q][5]);
            int e = 0;
            if (mx > 0.f && mx < 3.0e38f) {
                e = 4 - ilogbf(mx);
                e = e > 100 ? 100 : (e < -100 ? -100 : e);
            }
            const float sc = exp2f((float)e);
            if (k == 0 && wl < nwin) a.row_exp[w] = e;
            const int t = k / a.J, j = k - t * a.J;
            uint16_t* gh = reinterpret_cast<uint16_t*>(s_gh) + wl * a.T * a.pp + t * a.pp;
            uint16_t* gl = reinterpret_cast<uint16_t*>(s_gl) + wl * a.T * a.pp + t * a.pp;
            const float g3[3] = {gx * sc, gy * sc, gz * sc};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                uint16_t h, l;
                asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(g3[c]));
                float hf;
                asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(h));
                asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l) : "f"((g3[c] - hf) * 2048.f));
                gh[j * 3 + c] = h, gl[j * 3 + c] = l;
            }
            if (j == 0)
                for (int c = a.J * 3; c < a.pp; ++c) gh[c] = 0, gl[c] = 0;
        }
        if (bulk) fence_proxy_async_smem();
        __syncthreads();
        ns = ns / 2;                               // staged as 16-bit: the copies below count 4-byte words
    }
    if (bulk) {
        if (tid == 0) {
            bulk_s2g(a.grad + (size_t)w0 * n, s_g, (uint32_t)(kWinPerCta * n * sizeof(float)));
            if (a.gp_hi) {
                bulk_s2g(a.gp_hi + (size_t)w0 * ns, s_gh, (uint32_t)(kWinPerCta * ns * sizeof(float)));
                bulk_s2g(a.gp_lo + (size_t)w0 * ns, s_gl, (uint32_t)(kWinPerCta * ns * sizeof(float)));
            }
            bulk_commit();
            bulk_wait_read0();
        }
    } else {
        for (int i = tid; i < nwin * n; i += kThreads) a.grad[(size_t)w0 * n + i] = s_g[i];
        if (a.gp_hi)
            for (int i = tid; i < nwin * ns; i += kThreads) {
                a.gp_hi[(size_t)w0 * ns + i] = s_gh[i];
                a.gp_lo[(size_t)w0 * ns + i] = s_gl[i];
            }
    }
}

int launch_energy_grad(cudaStream_t stream, const CameraConst* cam, const SkeletonConst* skel, int W, int T, int J, int H,
                       int Wd, const float* pose, const float* pose0, const float* heat, const int64_t* frame_base, const int32_t* clip,
                       const float* mean_bone, const gem_energy_weights& wt, float* energy, float* terms,
                       float* grad, uint32_t* status, float* gp_hi, float* gp_lo, int pp, float* patch,
                       short2* patch_origin, unsigned long long* patch_stats, int gp_f16, int32_t* row_exp,
                       unsigned long long* patch_valid) {
    if (W <= 0) return GEM_OK;
    GEM_REQUIRE(skel != nullptr && skel->num_joints == J, "skeleton not set for this joint count");
    GEM_REQUIRE(wt.reproj == 0.f || cam != nullptr, "camera not set");
    GEM_REQUIRE(T * J <= kSlot, "T*J must be <= 160");
    GEM_REQUIRE(T >= 3, "seq_len must be >= 3");
    GEM_REQUIRE(wt.reproj == 0.f || (heat != nullptr && frame_base != nullptr), "heatmaps required when reproj != 0");
    EnergyArgs a;
    memset(&a.cam, 0, sizeof(a.cam));
    if (cam) a.cam = *cam;
    a.skel = *skel;
    a.pose = pose, a.pose0 = pose0, a.heat = heat, a.frame_base = frame_base, a.clip = clip, a.mean_bone = mean_bone;
    a.energy = energy, a.terms = terms, a.grad = grad, a.status = status;
    a.gp_hi = gp_hi, a.gp_lo = gp_hi ? gp_lo : nullptr, a.pp = gp_hi ? pp : 0;
    a.gp_f16 = gp_hi ? gp_f16 : 0, a.row_exp = row_exp;
    GEM_REQUIRE(!a.gp_f16 || (row_exp && (T * pp) % 8 == 0), "fp16 gradient output needs row_exp and T*pp % 8 == 0");
    a.patch = patch && patch_valid ? patch : nullptr;
    a.patch_origin = a.patch ? patch_origin : nullptr, a.patch_stats = a.patch ? patch_stats : nullptr;
    a.patch_valid = a.patch ? patch_valid : nullptr;
    GEM_REQUIRE(!gp_hi || (gp_lo && pp >= J * 3 && T * pp <= (gp_f16 ? 2 : 1) * kSplitMax && (T * pp) % 4 == 0),
                "bad split gradient layout");
    a.W = W, a.T = T, a.J = J, a.H = H, a.Wd = Wd;
    a.w3d = wt.w3d, a.ws = wt.smooth, a.wb = wt.bone, a.wv = wt.vae, a.wr = wt.reproj;
    const size_t pair_bytes = (size_t)kWinPerCta * T * J * 3 * sizeof(float);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    a.use_bulk = (pair_bytes % 16 == 0) && al16(pose) && al16(pose0) && al16(grad) &&
                 (!gp_hi || (al16(gp_hi) && al16(gp_lo)));
    const int grid = (W + kWinPerCta - 1) / kWinPerCta;
    energy_grad_kernel<<<grid, kThreads, 0, stream>>>(a);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

// fetches, into the joints' cache windows, the texels the energy evaluation of `pose` will sample (zero-copy maps)
int launch_texel_prefetch(cudaStream_t stream, const CameraConst* cam, int W, int T, int J, int H, int Wd, const float* pose,
                          const float* heat, const int64_t* frame_base, float* patch, short2* patch_origin,
                          unsigned long long* patch_valid, unsigned long long* patch_stats, int ctas) {
    if (W <= 0) return GEM_OK;
    GEM_REQUIRE(cam && heat && frame_base && patch && patch_origin && patch_valid, "texel prefetch needs the camera, the maps and the cache");
    EnergyArgs a;
    memset(&a, 0, sizeof(a));
    a.cam = *cam;
    a.pose = pose, a.heat = heat, a.frame_base = frame_base;
    a.patch = patch, a.patch_origin = patch_origin, a.patch_valid = patch_valid, a.patch_stats = patch_stats;
    a.W = W, a.T = T, a.J = J, a.H = H, a.Wd = Wd;
    const size_t total = (size_t)W * T * J;
    int grid = (int)((total + 511) / 512);
    if (grid > ctas) grid = ctas;
    texel_prefetch_kernel<<<grid, 512, 0, stream>>>(a);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

}  // namespace gem
