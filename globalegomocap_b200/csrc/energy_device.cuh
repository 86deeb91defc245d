// Per-joint part of the fused energy + analytic gradient (total_loss with the decoder bypassed: reference
// optimizer.py:139-149, 172-177, 202-218, 226-240; FishEyeCalibrated.py:96-129; grid_sample(bilinear, zeros,
// align_corners=True) and the autograd backward of all of them), shared by the stand-alone energy kernel (energy.cu)
// and by the backward chain kernel's energy prologue (gemm_tap_tc.cu), so that both evaluate the SAME arithmetic
// in the SAME order: results do not depend on which kernel runs the term.
//
// Every translation unit that includes this header is compiled with --fmad=false: the pixel-coordinate chain must
// round like ATen's separate fp32 ops, because floor() of the pixel coordinate selects the texels.
#pragma once
#include "common.cuh"

namespace gem {

constexpr int kEnergySlot = 160;       // joint slots per window (>= T*J, multiple of 32): 5 warps of partial sums

// everything about one energy evaluation that is the same for every joint of the launch
struct EnergyCommon {
    const float* heat;
    const int64_t* frame_base;
    const int32_t* clip;
    const float* mean_bone;
    uint32_t* status;
    int W, T, J, H, Wd;
    int planar;                // heat-map layout: 2 = tiled (below), 0 = [frames][H][Wd][J] (the pickle's HWC), 1 = [frames][J][H][Wd] (planar:
                               // what the reference itself permutes every window to, optimizer.py:251; x-neighbours share a sector)
    float w3d, ws, wb, wv, wr;
    // optional texel cache (used when the heat maps stay in pinned HOST memory and are read over PCIe): per joint
    // an 8x8 window of its map, 256 contiguous bytes in HBM, the map coordinate of its corner and a valid bit per
    // texel.  The same few texels are read by every evaluation of a stage (joints move by a fraction of a texel per
    // step), so only the texels the optimiser actually samples cross the bus, once each while the joint stays in
    // its window; values are copies of the map's, so the energy is bit-identical with and without the cache.
    float* patch;              // [W][T*J][kPatchWd * kPatchWd]
    short2* patch_origin;      // [W][T*J]: map coordinate of the window's corner
    unsigned long long* patch_valid;   // [W][T*J]: one bit per texel of the window (0: empty window)
    unsigned long long* patch_stats;   // optional {lookups, texels fetched}
    // camera and skeleton of the calling ctx: they travel as kernel parameters (constant bank), so two ctxs with
    // different calibrations never see each other's (FishEyeCalibrated.py:8-14 is per optimiser in the reference too)
    CameraConst cam;
    SkeletonConst skel;
};

// Problem shape as seen by the per-joint code.  ShapeDyn reads the launch's runtime values; ShapeFix carries them as
// compile-time constants (the reference's own configuration: 10 frames x 15 joints, 64 x 64 maps, 11 or 14 polynomial
// coefficients), which turns the divisions by J into multiplications, the texel addresses into shifts and unrolls the
// polynomial.  Same operations on the same values in the same order: results do not depend on the policy.
struct ShapeDyn {
    int T, J, H, Wd, n_poly;
    __device__ __forceinline__ explicit ShapeDyn(const EnergyCommon& a) : T(a.T), J(a.J), H(a.H), Wd(a.Wd), n_poly(a.cam.n_poly) {}
};
template <int kT, int kJ, int kH, int kW, int kNP>
struct ShapeFix {
    static constexpr int T = kT, J = kJ, H = kH, Wd = kW, n_poly = kNP;
    __device__ __forceinline__ explicit ShapeFix(const EnergyCommon&) {}
};

__device__ __forceinline__ float texel(const float* __restrict__ heat, int64_t frame, int y, int x, int j, int H,
                                       int Wd, int J, int planar = 0) {
    if (x < 0 || x >= Wd || y < 0 || y >= H) return 0.f;     // padding_mode='zeros'
    if (planar == 2)       // tiled: [H/4][Wd/8] tiles of 4 rows x 8 texels
        return __ldg(heat + (frame * J + j) * (int64_t)H * Wd + ((int64_t)(y >> 2) * (Wd >> 3) + (x >> 3)) * 32 + (y & 3) * 8 + (x & 7));
    if (planar) return __ldg(heat + ((frame * J + j) * H + y) * (int64_t)Wd + x);
    return __ldg(heat + ((frame * H + y) * (int64_t)Wd + x) * J + j);
}

// Demand-fetched 8x8 window of one joint's map: origin and a valid bit per texel; only texels of the bilinear footprint
// (x0, y0) .. (x0+1, y0+1) that are not in the window yet are read from the map (over PCIe when it is host memory).
// A footprint that leaves the window re-centres it (and empties it).
__device__ __forceinline__ void cache_lookup(const EnergyCommon& a, size_t pk, int64_t frame, int j, int x0, int y0,
                                             bool count_lookup, float& nw, float& ne, float& sw, float& se) {
    constexpr int kPatchW = kPatchWd;
    float* pe = a.patch + pk * kPatchFloats;
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

// The window over PLANAR maps.  A map row is contiguous, and a read of pinned host memory costs the same whether it
// brings 16 or 128 contiguous bytes (tools/pcie_gather_bench.cu: ~16 M requests/s per SM, ~150 M/s per GPU, at any size
// up to a 128-byte line), so the fetch unit is a whole window ROW: kPlanarW texels (64 B), origin aligned to
// kPlanarAlign texels in x, one valid bit per row.  Rows are fetched by texel_fetch_kernel (energy.cu), kPlanarW / 4
// lanes per row so that a row is one request; the per-thread path below is the fallback when nothing prefetched
// (resident maps with the cache forced on).  Requires Wd % 4 == 0.
__device__ __forceinline__ void planar_recentre(int x0, int y0, int& ox, int& oy) {
    // x0 - ox lands in [W/2 - A/2, W/2 + A/2 - 1], y0 - oy = H/2 - 1
    ox = (x0 - (kPlanarW / 2 - kPlanarAlign / 2)) & ~(kPlanarAlign - 1), oy = y0 - (kPlanarH / 2 - 1);
}
__device__ __forceinline__ bool planar_outside(unsigned long long valid, int dx, int dy) {
    return valid == 0ull || dx < 0 || dx > kPlanarW - 2 || dy < 0 || dy > kPlanarH - 2;
}
// true when the footprint at (x0, y0) is not in the joint's window: empty window, outside it, or a row not fetched yet
__device__ __forceinline__ bool planar_window_miss(const EnergyCommon& a, size_t pk, int x0, int y0) {
    const short2 o = a.patch_origin[pk];
    const unsigned long long valid = a.patch_valid[pk];
    const int dx = x0 - o.x, dy = y0 - o.y;
    if (planar_outside(valid, dx, dy)) return true;
    return ((3ull << dy) & ~valid) != 0ull;
}

__device__ __forceinline__ void cache_lookup_planar(const EnergyCommon& a, size_t pk, int64_t frame, int j, int x0, int y0,
                                                    bool count_lookup, float& nw, float& ne, float& sw, float& se) {
    float* pe = a.patch + pk * kPatchFloats;
    const short2 o = a.patch_origin[pk];
    unsigned long long valid = a.patch_valid[pk];
    int ox = o.x, oy = o.y;
    int dx = x0 - ox, dy = y0 - oy;
    if (planar_outside(valid, dx, dy)) {
        planar_recentre(x0, y0, ox, oy);
        valid = 0ull, dx = x0 - ox, dy = y0 - oy;
    }
    const unsigned long long need = (3ull << dy) & ~valid;
    int requests = 0;
    if (need) {                                   // fallback: this thread reads the missing rows itself, unit by unit
        const float* plane = a.heat + (frame * a.J + j) * (int64_t)a.H * a.Wd;
        for (int r = 0; r < 2; ++r) {
            if (!((need >> (dy + r)) & 1ull)) continue;
            const int y = oy + dy + r;
#pragma unroll 1
            for (int u = 0; u < kPlanarW / 4; ++u) {
                const int xs = ox + 4 * u;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);           // padding_mode='zeros' outside the map
                if (y >= 0 && y < a.H && xs >= 0 && xs < a.Wd) v = __ldg(reinterpret_cast<const float4*>(plane + (int64_t)y * a.Wd + xs));
                *reinterpret_cast<float4*>(pe + (dy + r) * kPlanarW + 4 * u) = v;
            }
            if (y >= 0 && y < a.H) requests += kPlanarW * 4 / 32;     // counted in 32-byte sectors
        }
        a.patch_origin[pk] = make_short2((short)ox, (short)oy);
        a.patch_valid[pk] = valid | need;
    }
    if (a.patch_stats) {
        if (count_lookup) atomicAdd(a.patch_stats, 1ull);
        if (requests) atomicAdd(a.patch_stats + 1, (unsigned long long)requests);
    }
    const float* p0 = pe + dy * kPlanarW + dx;
    nw = p0[0], ne = p0[1], sw = p0[kPlanarW], se = p0[kPlanarW + 1];
}

// One fetch event of texel_fetch_kernel, executed by a group of kRows * kPlanarW / 4 lanes (lane `gl` of the group): the
// footprint's two rows and (kRows - 2) / 2 rows above and below, kPlanarW / 4 lanes per row.
template <int kRows>
__device__ __forceinline__ void planar_fetch_rows(const EnergyCommon& a, size_t pk, int64_t frame, int j, int x0, int y0, int gl,
                                                  unsigned group_mask) {
    constexpr int kLanesPerRow = kPlanarW / 4;
    float* pe = a.patch + pk * kPatchFloats;
    const short2 o = a.patch_origin[pk];
    unsigned long long valid = a.patch_valid[pk];
    __syncwarp(group_mask);                       // every lane has read the window's state before lane 0 rewrites it
    int ox = o.x, oy = o.y;
    int dx = x0 - ox, dy = y0 - oy;
    if (planar_outside(valid, dx, dy)) {
        planar_recentre(x0, y0, ox, oy);
        valid = 0ull, dx = x0 - ox, dy = y0 - oy;
    }
    const int r = gl / kLanesPerRow, u = gl - r * kLanesPerRow;
    const int row = dy - (kRows - 2) / 2 + r;
    int lo = dy - (kRows - 2) / 2, hi = lo + kRows - 1;
    lo = lo < 0 ? 0 : lo, hi = hi > kPlanarH - 1 ? kPlanarH - 1 : hi;
    const unsigned long long span = ((2ull << hi) - 1ull) & ~((1ull << lo) - 1ull);
    const unsigned long long need = span & ~valid;
    if (row >= 0 && row < kPlanarH && ((need >> row) & 1ull)) {
        const int y = oy + row, xs = ox + 4 * u;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);                   // padding_mode='zeros' outside the map
        if (y >= 0 && y < a.H && xs >= 0 && xs < a.Wd) {
            const float* plane = a.heat + (frame * a.J + j) * (int64_t)a.H * a.Wd;
            v = __ldg(reinterpret_cast<const float4*>(plane + (int64_t)y * a.Wd + xs));
        }
        *reinterpret_cast<float4*>(pe + row * kPlanarW + 4 * u) = v;
        if (a.patch_stats && u == 0 && y >= 0 && y < a.H) atomicAdd(a.patch_stats + 1, (unsigned long long)(kPlanarW * 4 / 32));
    }
    if (gl == 0) {
        a.patch_origin[pk] = make_short2((short)ox, (short)oy);
        a.patch_valid[pk] = valid | need;
    }
}

// The window over TILED maps (layout 2: every map stored as [H/4][Wd/8] tiles of 4 rows x 8 texels, one 128-byte line
// each).  The bus charges per request, not per byte (section 4 of DESIGN.md), and a row-major map can only offer one
// row per request; a tile brings 4 x 8 texels of the joint's neighbourhood in ONE request, so a bilinear footprint
// costs 1.4 requests on average (1, 2 or 4 tiles) instead of 2 rows, and the joint's later positions mostly fall into
// tiles that are already there.  The window is the same 16 x 16 texels of HBM (row-major), its origin aligned to the
// tile grid, with one valid bit per tile (2 across, 4 down).  Requires H % 4 == 0 and Wd % 8 == 0.
__device__ __forceinline__ void tiled_recentre(int x0, int y0, int& ox, int& oy) {
    ox = (x0 - 4) & ~(kTileW - 1), oy = (y0 - 6) & ~(kTileH - 1);           // x0 - ox in [4, 11], y0 - oy in [6, 9]
}
// bits of the window's tiles the footprint at window coordinates (dx, dy) touches
__device__ __forceinline__ unsigned tiled_foot(int dx, int dy) {
    const unsigned cols = (1u << (dx >> 3)) | (1u << ((dx + 1) >> 3));
    return (cols << (2 * (dy >> 2))) | (cols << (2 * ((dy + 1) >> 2)));
}
__device__ __forceinline__ bool tiled_window_miss(const EnergyCommon& a, size_t pk, int x0, int y0) {
    const short2 o = a.patch_origin[pk];
    const unsigned long long valid = a.patch_valid[pk];
    const int dx = x0 - o.x, dy = y0 - o.y;
    if (planar_outside(valid, dx, dy)) return true;
    return (tiled_foot(dx, dy) & ~(unsigned)valid) != 0u;
}
// address of the tile that holds map texel (y, x) (both multiples of the tile size here), or NULL outside the map
__device__ __forceinline__ const float* tile_ptr(const EnergyCommon& a, int64_t frame, int j, int y, int x) {
    if (y < 0 || y >= a.H || x < 0 || x >= a.Wd) return nullptr;           // a tile is wholly inside or wholly outside
    const float* map = a.heat + (frame * a.J + j) * (int64_t)a.H * a.Wd;
    return map + ((int64_t)(y / kTileH) * (a.Wd / kTileW) + x / kTileW) * kTileFloats;
}
// per-thread path (nothing prefetched: resident maps, or the probe was switched off)
__device__ __forceinline__ void cache_lookup_tiled(const EnergyCommon& a, size_t pk, int64_t frame, int j, int x0, int y0,
                                                   bool count_lookup, float& nw, float& ne, float& sw, float& se) {
    float* pe = a.patch + pk * kPatchFloats;
    const short2 o = a.patch_origin[pk];
    unsigned long long valid = a.patch_valid[pk];
    int ox = o.x, oy = o.y;
    int dx = x0 - ox, dy = y0 - oy;
    if (planar_outside(valid, dx, dy)) {
        tiled_recentre(x0, y0, ox, oy);
        valid = 0ull, dx = x0 - ox, dy = y0 - oy;
    }
    const unsigned need = tiled_foot(dx, dy) & ~(unsigned)valid;
    int requests = 0;
    if (need) {
#pragma unroll 1
        for (int q = 0; q < 8; ++q) {
            if (!((need >> q) & 1u)) continue;
            const int ty = q >> 1, tx = q & 1;
            const float* src = tile_ptr(a, frame, j, oy + ty * kTileH, ox + tx * kTileW);
            requests += src ? kTileFloats * 4 / 32 : 0;                      // counted in 32-byte sectors
#pragma unroll 1
            for (int l = 0; l < 8; ++l) {
                const float4 v = src ? __ldg(reinterpret_cast<const float4*>(src) + l) : make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(pe + (ty * kTileH + (l >> 1)) * kPlanarW + tx * kTileW + (l & 1) * 4) = v;
            }
        }
        a.patch_origin[pk] = make_short2((short)ox, (short)oy);
        a.patch_valid[pk] = valid | need;
    }
    if (a.patch_stats) {
        if (count_lookup) atomicAdd(a.patch_stats, 1ull);
        if (requests) atomicAdd(a.patch_stats + 1, (unsigned long long)requests);
    }
    const float* p0 = pe + dy * kPlanarW + dx;
    nw = p0[0], ne = p0[1], sw = p0[kPlanarW], se = p0[kPlanarW + 1];
}
// One fetch event of texel_fetch_tiles_kernel, executed by a group of eight lanes (lane l of the group = one float4 of a
// tile's 128 bytes: row l / 2, half l % 2, so a tile is ONE request): the loads of the (up to) four tiles of the
// footprint are all issued before the first store.
__device__ __forceinline__ void tiled_fetch_event(const EnergyCommon& a, size_t pk, int64_t frame, int j, int x0, int y0, int l,
                                                  unsigned group_mask) {
    float* pe = a.patch + pk * kPatchFloats;
    const short2 o = a.patch_origin[pk];
    unsigned long long valid = a.patch_valid[pk];
    __syncwarp(group_mask);                       // every lane has read the window's state before lane 0 rewrites it
    int ox = o.x, oy = o.y;
    int dx = x0 - ox, dy = y0 - oy;
    if (planar_outside(valid, dx, dy)) {
        tiled_recentre(x0, y0, ox, oy);
        valid = 0ull, dx = x0 - ox, dy = y0 - oy;
    }
    const unsigned need = tiled_foot(dx, dy) & ~(unsigned)valid;
    float4 v[4];
    bool take[4];
    int ty[4], tx[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        ty[q] = ((q >> 1) ? (dy + 1) : dy) >> 2, tx[q] = ((q & 1) ? (dx + 1) : dx) >> 3;
        // the second row / column of tiles only when the footprint really crosses into it (else it repeats the first)
        const bool distinct = (!(q >> 1) || ty[q] != (dy >> 2)) && (!(q & 1) || tx[q] != (dx >> 3));
        take[q] = distinct && ((need >> (ty[q] * 2 + tx[q])) & 1u);
        v[q] = make_float4(0.f, 0.f, 0.f, 0.f);   // padding_mode='zeros' outside the map
        if (take[q]) {
            const float* src = tile_ptr(a, frame, j, oy + ty[q] * kTileH, ox + tx[q] * kTileW);
            if (src) {
                v[q] = __ldg(reinterpret_cast<const float4*>(src) + l);
                if (a.patch_stats && l == 0) atomicAdd(a.patch_stats + 1, (unsigned long long)(kTileFloats * 4 / 32));
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (take[q]) *reinterpret_cast<float4*>(pe + (ty[q] * kTileH + (l >> 1)) * kPlanarW + tx[q] * kTileW + (l & 1) * 4) = v[q];
    if (l == 0) {
        a.patch_origin[pk] = make_short2((short)ox, (short)oy);
        a.patch_valid[pk] = valid | need;
    }
}

// Fisheye projection of a joint and the map cell its bilinear footprint starts at (FishEyeCalibrated.py:96-129,
// optimizer.py:139-149); the arithmetic (and --fmad=false) is what decides the cell, so the energy code and the
// texel prefetch kernel share it.  Returns false when r == 0.
struct Proj {
    float r, inv, rho, drho, ix, iy, fx0, fy0;
};
__device__ __forceinline__ bool project_joint(const CameraConst& c_cam, float x, float y, float z, int H, int Wd, Proj& p,
                                              int n_poly = -1) {
    if (n_poly < 0) n_poly = c_cam.n_poly;
    const float zn = -z;
    p.r = sqrtf(x * x + y * y);
    if (p.r == 0.f) return false;
    const float theta = atanf(zn / p.r);
    float rho = c_cam.poly[0], drho = 0.f, ti = 1.f;
#pragma unroll
    for (int i = 1; i < n_poly; ++i) {                // power accumulation, not Horner
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

// First half of a joint's reprojection term: the projection and the four texel loads of its bilinear footprint.  Kept
// apart from the arithmetic so that callers can have the loads of several joints in flight before they consume any
// (the chain kernel's prologue evaluates four joints per thread); the stand-alone kernel calls both halves back to back.
struct JointTexels {
    Proj pj;
    float nw, ne, sw, se;
    bool rp;
};
template <class S>
__device__ __forceinline__ void joint_gather(const EnergyCommon& a, const S& sh, const float* X, int w, int k, JointTexels& jt) {
    jt.rp = false;
    jt.nw = jt.ne = jt.sw = jt.se = 0.f;
    if (a.wr == 0.f) return;
    const int t = k / sh.J, j = k - t * sh.J;
    const float x = X[k * 3 + 0], y = X[k * 3 + 1], z = X[k * 3 + 2];
    if (!project_joint(a.cam, x, y, z, sh.H, sh.Wd, jt.pj, sh.n_poly)) {
        if (a.status) atomicOr(a.status + w, GEM_WIN_NORM_ZERO);
    } else if (jt.pj.fx0 >= -1.f && jt.pj.fx0 <= (float)sh.Wd && jt.pj.fy0 >= -1.f && jt.pj.fy0 <= (float)sh.H) {
        // (anything further than one texel outside contributes exactly 0)
        jt.rp = true;
        const int x0 = (int)jt.pj.fx0, y0 = (int)jt.pj.fy0;
        const int64_t frame = a.frame_base[w] + t;
        if (a.patch && a.planar == 2) {
            cache_lookup_tiled(a, (size_t)w * (sh.T * sh.J) + k, frame, j, x0, y0, true, jt.nw, jt.ne, jt.sw, jt.se);
        } else if (a.patch && a.planar) {
            cache_lookup_planar(a, (size_t)w * (sh.T * sh.J) + k, frame, j, x0, y0, true, jt.nw, jt.ne, jt.sw, jt.se);
        } else if (a.patch) {
            cache_lookup(a, (size_t)w * (sh.T * sh.J) + k, frame, j, x0, y0, true, jt.nw, jt.ne, jt.sw, jt.se);
        } else {
            // one 64-bit base per map (or frame), 32-bit offsets inside it
            const int H = sh.H, Wd = sh.Wd, J = sh.J;
            const bool xl = x0 >= 0 && x0 < Wd, xr = x0 + 1 >= 0 && x0 + 1 < Wd;           // padding_mode='zeros'
            const bool yt = y0 >= 0 && y0 < H, yb = y0 + 1 >= 0 && y0 + 1 < H;
            if (a.planar == 2) {           // tiled: [H/4][Wd/8] tiles of 4 rows x 8 texels
                const float* map = a.heat + (frame * J + j) * (int64_t)(H * Wd);
                auto at = [&](int yy, int xx) { return __ldg(map + (((yy >> 2) * (Wd >> 3) + (xx >> 3)) * 32 + (yy & 3) * 8 + (xx & 7))); };
                jt.nw = (xl && yt) ? at(y0, x0) : 0.f;
                jt.ne = (xr && yt) ? at(y0, x0 + 1) : 0.f;
                jt.sw = (xl && yb) ? at(y0 + 1, x0) : 0.f;
                jt.se = (xr && yb) ? at(y0 + 1, x0 + 1) : 0.f;
            } else if (a.planar) {
                const float* map = a.heat + (frame * J + j) * (int64_t)(H * Wd);
                jt.nw = (xl && yt) ? __ldg(map + (y0 * Wd + x0)) : 0.f;
                jt.ne = (xr && yt) ? __ldg(map + (y0 * Wd + x0 + 1)) : 0.f;
                jt.sw = (xl && yb) ? __ldg(map + ((y0 + 1) * Wd + x0)) : 0.f;
                jt.se = (xr && yb) ? __ldg(map + ((y0 + 1) * Wd + x0 + 1)) : 0.f;
            } else {
                const float* map = a.heat + frame * (int64_t)(H * Wd * J) + j;
                jt.nw = (xl && yt) ? __ldg(map + (y0 * Wd + x0) * J) : 0.f;
                jt.ne = (xr && yt) ? __ldg(map + (y0 * Wd + x0 + 1) * J) : 0.f;
                jt.sw = (xl && yb) ? __ldg(map + ((y0 + 1) * Wd + x0) * J) : 0.f;
                jt.se = (xr && yb) ? __ldg(map + ((y0 + 1) * Wd + x0 + 1) * J) : 0.f;
            }
        }
    }
}

// The five energy terms of joint-frame k = t*J + j of window w and its dE/dx: X / X0 are the window's pose and anchor
// (T*J*3 floats, shared memory), jt the joint's gathered texels.  e = {E_3d, E_smooth, E_bone, E_vae, E_reproj}
// contributions, g = weighted gradient.
template <class S>
__device__ __forceinline__ void joint_terms(const EnergyCommon& a, const S& sh, const float* X, const float* X0, int w, int k,
                                            const JointTexels& jt, float (&e)[5], float (&g)[3]) {
    float e3d = 0.f, esm = 0.f, ebn = 0.f, eva = 0.f, erp = 0.f;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    {
        const int t = k / sh.J, j = k - t * sh.J;
        const float x = X[k * 3 + 0], y = X[k * 3 + 1], z = X[k * 3 + 2];
        const bool rp = jt.rp;
        const Proj& pj = jt.pj;
        const float nw = jt.nw, ne = jt.ne, sw = jt.sw, se = jt.se;

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
            // the five frames t-2 .. t+2 of this joint, each read once (rows outside the window are never used)
            float v[5][3];
#pragma unroll
            for (int m = 0; m < 5; ++m) {
                const int f = t - 2 + m;
                const bool in = (m == 2) || (f >= 0 && f < sh.T);
                const float* pf = X + (k + (in ? (m - 2) * sh.J : 0)) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) v[m][c] = (m == 2) ? (c == 0 ? x : (c == 1 ? y : z)) : pf[c];
            }
            float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int q = 0; q < 3; ++q) {        // q = 0: a_t (coef 1), 1: a_{t-1} (coef -2), 2: a_{t-2} (coef 1)
                const int s = t - q;
                if (s < 0 || s > sh.T - 3) continue;
                const float coef = (q == 1) ? -2.f : 1.f;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float ac = (v[2 - q][c] - v[3 - q][c]) - (v[3 - q][c] - v[4 - q][c]);
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
            const float* Xt = X + (k - j) * 3;                     // frame t's joints
            const float bx = x - Xt[p * 3 + 0], by = y - Xt[p * 3 + 1], bz = z - Xt[p * 3 + 2];
            const float len = sqrtf(bx * bx + by * by + bz * bz);
            const float diff = len - mb[j];
            ebn = diff * diff;
            if (len > 0.f) {                        // torch.norm's subgradient at 0 is 0
                const float c = a.wb * (2.f * diff) / len;
                gx += c * bx, gy += c * by, gz += c * bz;
            }
            for (int ci = a.skel.child_start[j]; ci < a.skel.child_start[j + 1]; ++ci) {   // this joint as a parent
                const int cj = a.skel.child_list[ci];
                const float cx = Xt[cj * 3 + 0] - x, cy = Xt[cj * 3 + 1] - y, cz = Xt[cj * 3 + 2] - z;
                const float cl = sqrtf(cx * cx + cy * cy + cz * cz);
                if (cl > 0.f) {
                    const float c = a.wb * (2.f * (cl - mb[cj])) / cl;
                    gx -= c * cx, gy -= c * cy, gz -= c * cz;
                }
            }
        }
        // E_reproj = -sum bilinear(H_tj; pix(project(x)))            optimizer.py:139-149
        if (rp) {
            const float r = pj.r, inv = pj.inv, rho = pj.rho, drho = pj.drho;
            const float ix = pj.ix, iy = pj.iy, fx0 = pj.fx0, fy0 = pj.fy0;
            const float wx1 = ix - fx0, wx0 = (fx0 + 1.f) - ix;
            const float wy1 = iy - fy0, wy0 = (fy0 + 1.f) - iy;
            erp = -(nw * (wx0 * wy0) + ne * (wx1 * wy0) + sw * (wx0 * wy1) + se * (wx1 * wy1));
            const float ds_dix = -nw * wy0 + ne * wy0 - sw * wy1 + se * wy1;
            const float ds_diy = -nw * wx0 - ne * wx1 + sw * wx0 + se * wx1;
            const float du = ds_dix * ((float)(sh.Wd - 1) * 0.5f / 512.f);  // dS/du
            const float dv = ds_diy * ((float)(sh.H - 1) * 0.5f / 512.f);   // dS/dv
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
    e[0] = e3d, e[1] = esm, e[2] = ebn, e[3] = eva, e[4] = erp;
    g[0] = gx, g[1] = gy, g[2] = gz;
}

template <class S>
__device__ __forceinline__ void joint_energy_grad(const EnergyCommon& a, const S& sh, const float* X, const float* X0, int w, int k,
                                                  float (&e)[5], float (&g)[3]) {
    JointTexels jt;
    joint_gather(a, sh, X, w, k, jt);
    joint_terms(a, sh, X, X0, w, k, jt, e, g);
}

// optimizer.py:239-240 (left to right; the reproj product is skipped when its weight is 0) from the window's five
// term sums
__device__ __forceinline__ float combine_energy(const EnergyCommon& a, const float (&t5)[5]) {
    float E = a.w3d * t5[0] + a.ws * t5[1] + a.wb * t5[2] + a.wv * t5[3];
    if (a.wr != 0.f) E += a.wr * t5[4];
    return E;
}

// fp16 scheme: the power of two that brings a window's largest |dE/dx| entry to ~2^4 (gradients shrink to 1e-7 near
// convergence, far below fp16's normal range); removed again, exactly, by the T*256 -> latent GEMM's epilogue
__device__ __forceinline__ int grad_exponent(float mx) {
    int e = 0;
    if (mx > 0.f && mx < 3.0e38f) {
        e = 4 - ilogbf(mx);
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
    }
    return e;
}
__device__ __forceinline__ void split_f16_energy(float x, uint16_t& h, uint16_t& l) {      // = tc::split_f16
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
    float hf;
    asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(h));
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l) : "f"((x - hf) * 2048.f));
}

}  // namespace gem
