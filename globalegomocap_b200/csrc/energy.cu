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
#include <stdlib.h>
#include <string.h>

#include "kernels.cuh"
#include "energy_device.cuh"

namespace gem {

int g_energy_fixed = -1;

constexpr int kSlot = kEnergySlot;     // threads per window (>= T*J, multiple of 32)
constexpr int kWinPerCta = 2;
constexpr int kThreads = kSlot * kWinPerCta;
constexpr int kSplitMax = 512;         // floats per window of the split gradient tile: T * pp <= 512
static_assert(kPatchW == kPatchWd, "texel window side");

struct EnergyArgs : EnergyCommon {
    const float* pose;
    const float* pose0;
    float* energy;
    float* terms;
    float* grad;
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
};

// Zero-copy maps: fetches the texels the next energy evaluation will sample into the joints' cache windows.  A
// handful of CTAs (grid-stride over all joints) so that the wait for PCIe occupies a few SMs instead of every SM:
// energy CTAs stalled on PCIe reads are small but sit on all SMs, and no 200 KB tensor-core CTA of another slice can
// be scheduled beside them — the solve then serialises with the transfers instead of overlapping them.
__global__ void __launch_bounds__(512) texel_prefetch_kernel(const __grid_constant__ EnergyArgs a) {
    const int TJ = a.T * a.J;
    const size_t total = (size_t)a.W * TJ;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(i / TJ), k = (int)(i - (size_t)w * TJ);
        const int t = k / a.J, j = k - t * a.J;
        const float x = a.pose[i * 3 + 0], y = a.pose[i * 3 + 1], z = a.pose[i * 3 + 2];
        Proj p;
        if (!project_joint(a.cam, x, y, z, a.H, a.Wd, p)) continue;
        if (!(p.fx0 >= -1.f && p.fx0 <= (float)a.Wd && p.fy0 >= -1.f && p.fy0 <= (float)a.H)) continue;
        float nw, ne, sw, se;
        if (a.planar == 2) cache_lookup_tiled(a, i, a.frame_base[w] + t, j, (int)p.fx0, (int)p.fy0, false, nw, ne, sw, se);
        else if (a.planar) cache_lookup_planar(a, i, a.frame_base[w] + t, j, (int)p.fx0, (int)p.fy0, false, nw, ne, sw, se);
        else cache_lookup(a, i, a.frame_base[w] + t, j, (int)p.fx0, (int)p.fy0, false, nw, ne, sw, se);
    }
}

// Planar maps, two launches instead of one: the PROBE (one thread per joint, every SM, no bus traffic: projection and
// a look at the joint's window) appends the joints whose footprint is not in their window to a miss list; the FETCH
// kernel (a few CTAs) walks only that list - a few per cent of the joints in a typical round - and fetches whole window
// rows, kPlanarW / 4 lanes per row, so that a row is ONE request on the bus (energy_device.cuh: planar_fetch_rows).  The list's order varies from run to run, its
// contents do not, and the fetched values are copies of the map's.
__global__ void __launch_bounds__(128) texel_probe_kernel(const __grid_constant__ EnergyArgs a, uint32_t* __restrict__ count,
                                                          uint2* __restrict__ list) {
    const int TJ = a.T * a.J;
    const size_t total = (size_t)a.W * TJ;
    const size_t i = (size_t)blockIdx.x * 128 + threadIdx.x;
    bool miss = false;
    int x0 = 0, y0 = 0;
    if (i < total) {
        const float x = a.pose[i * 3 + 0], y = a.pose[i * 3 + 1], z = a.pose[i * 3 + 2];
        Proj p;
        if (project_joint(a.cam, x, y, z, a.H, a.Wd, p) &&
            (p.fx0 >= -1.f && p.fx0 <= (float)a.Wd && p.fy0 >= -1.f && p.fy0 <= (float)a.H)) {
            x0 = (int)p.fx0, y0 = (int)p.fy0;
            miss = a.planar == 2 ? tiled_window_miss(a, i, x0, y0) : planar_window_miss(a, i, x0, y0);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, miss);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(count, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (miss) list[base + __popc(m & ((1u << lane) - 1u))] = make_uint2((uint32_t)i, ((uint32_t)(uint16_t)(short)x0) | ((uint32_t)(uint16_t)(short)y0 << 16));
}

template <int kRows>
__global__ void __launch_bounds__(512) texel_fetch_kernel(const __grid_constant__ EnergyArgs a, uint32_t* __restrict__ count,
                                                          const uint2* __restrict__ list) {
    constexpr int kPlanarFetchLanes = kRows * (kPlanarW / 4);               // lanes that share one fetch event
    static_assert(kRows % 2 == 0 && kPlanarFetchLanes <= 32 && 32 % kPlanarFetchLanes == 0, "a fetch event is handled inside one warp");
    const int TJ = a.T * a.J;
    const uint32_t n = *reinterpret_cast<volatile uint32_t*>(count);
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const int gl = (int)(tid % kPlanarFetchLanes);                          // lane within the event's group
    const unsigned group_mask = (unsigned)(((1ull << kPlanarFetchLanes) - 1ull) << ((threadIdx.x & 31) - gl));
    const uint32_t groups = gridDim.x * blockDim.x / kPlanarFetchLanes;
    for (uint32_t e = tid / kPlanarFetchLanes; e < n; e += groups) {
        const uint2 m = list[e];
        const size_t pk = m.x;
        const int x0 = (short)(m.y & 0xffffu), y0 = (short)(m.y >> 16);
        const int w = (int)(pk / TJ), k = (int)(pk - (size_t)w * TJ);
        const int t = k / a.J, j = k - t * a.J;
        planar_fetch_rows<kRows>(a, pk, a.frame_base[w] + t, j, x0, y0, gl, group_mask);
    }
    // the last CTA to finish empties the list for the next round (every CTA has read n by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(count + 1, 1u) == gridDim.x - 1) count[0] = 0u, count[1] = 0u;
    }
}

// tiled maps: eight lanes per fetch event, a 128-byte tile per load instruction of the group (energy_device.cuh:
// tiled_fetch_event)
__global__ void __launch_bounds__(512) texel_fetch_tiles_kernel(const __grid_constant__ EnergyArgs a, uint32_t* __restrict__ count,
                                                                const uint2* __restrict__ list) {
    const int TJ = a.T * a.J;
    const uint32_t n = *reinterpret_cast<volatile uint32_t*>(count);
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t groups = gridDim.x * blockDim.x / 8;
    const int l = (int)(threadIdx.x & 7);
    const unsigned group_mask = 0xffu << ((threadIdx.x & 31) - l);
    for (uint32_t e = tid / 8; e < n; e += groups) {
        const uint2 m = list[e];
        const size_t pk = m.x;
        const int x0 = (short)(m.y & 0xffffu), y0 = (short)(m.y >> 16);
        const int w = (int)(pk / TJ), k = (int)(pk - (size_t)w * TJ);
        const int t = k / a.J, j = k - t * a.J;
        tiled_fetch_event(a, pk, a.frame_base[w] + t, j, x0, y0, l, group_mask);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(count + 1, 1u) == gridDim.x - 1) count[0] = 0u, count[1] = 0u;
    }
}

// (Round 2 also measured a persistent, double-buffered variant — CTAs walking over window pairs with the next pair's
// tiles in flight and stores not waited for, bit-identical: 27.7 us against 27.8 us at 1870 windows, 1.02 ms against
// 0.95 ms at 100 980.  The kernel is bound by instruction issue — about 500 instructions per joint with --fmad=false,
// IEEE divisions / square roots and the eleven-term polynomial — not by the load latency that pipelining hides.)
//
// S is the shape policy (energy_device.cuh): ShapeFix for the reference's configuration, ShapeDyn for anything else.
template <class S>
__global__ void __launch_bounds__(kThreads, 4) energy_grad_kernel(const __grid_constant__ EnergyArgs a) {
    const S sh(a);
    __shared__ __align__(16) float s_x[kWinPerCta * kSlot * 3];
    __shared__ __align__(16) float s_x0[kWinPerCta * kSlot * 3];
    __shared__ __align__(16) float s_g[kWinPerCta * kSlot * 3];
    __shared__ float s_red[kThreads / 32][6];
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(16) float s_gh[kWinPerCta * kSplitMax];
    __shared__ __align__(16) float s_gl[kWinPerCta * kSplitMax];

    const int tid = threadIdx.x;
    const int TJ = sh.T * sh.J;
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
        float e5[5], g3[3];
        joint_energy_grad(a, sh, X, X0, w, k, e5, g3);        // energy_device.cuh: shared with the chain kernel's prologue
        e3d = e5[0], esm = e5[1], ebn = e5[2], eva = e5[3], erp = e5[4];
        gx = g3[0], gy = g3[1], gz = g3[2];
    }
    if (wl < kWinPerCta && k < TJ) {
        if (a.grad) {
            float* G = s_g + wl * n;
            G[k * 3 + 0] = gx, G[k * 3 + 1] = gy, G[k * 3 + 2] = gz;
        }
        if (a.gp_hi && !a.gp_f16) {
            const int t = k / sh.J, j = k - t * sh.J;
            const int o = wl * sh.T * a.pp + t * a.pp + j * 3;
            const float g3[3] = {gx, gy, gz};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float h = __uint_as_float(__float_as_uint(g3[c]) & 0xFFFFE000u);
                s_gh[o + c] = h, s_gl[o + c] = g3[c] - h;
            }
            for (int c = sh.J * 3 + j; c < a.pp; c += sh.J)      // the row's padding columns, dealt over its J threads
                s_gh[wl * sh.T * a.pp + t * a.pp + c] = 0.f, s_gl[wl * sh.T * a.pp + t * a.pp + c] = 0.f;
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
        a.energy[ww] = combine_energy(a, t5);
    }
    int ns = sh.T * a.pp;                           // floats per window of the split tile
    if (a.gp_hi && a.gp_f16) {
        // fp16 scheme: scale the window by 2^e (max entry -> ~2^4), split, stage as uint16
        constexpr int wpw = kSlot / 32;
        if (wl < kWinPerCta && k < TJ) {
            float mx = 0.f;
            for (int q = 0; q < wpw; ++q) mx = fmaxf(mx, s_red[wl * wpw + q][5]);
            const int e = grad_exponent(mx);
            const float sc = exp2f((float)e);
            if (k == 0 && wl < nwin) a.row_exp[w] = e;
            const int t = k / sh.J, j = k - t * sh.J;
            uint16_t* gh = reinterpret_cast<uint16_t*>(s_gh) + wl * sh.T * a.pp + t * a.pp;
            uint16_t* gl = reinterpret_cast<uint16_t*>(s_gl) + wl * sh.T * a.pp + t * a.pp;
            const float g3[3] = {gx * sc, gy * sc, gz * sc};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                uint16_t h, l;
                split_f16_energy(g3[c], h, l);
                gh[j * 3 + c] = h, gl[j * 3 + c] = l;
            }
            for (int c = sh.J * 3 + j; c < a.pp; c += sh.J) gh[c] = 0, gl[c] = 0;      // padding columns, dealt over the row's J threads
        }
        if (bulk) fence_proxy_async_smem();
        __syncthreads();
        ns = ns / 2;                               // staged as 16-bit: the copies below count 4-byte words
    }
    if (bulk) {
        if (tid == 0) {
            if (a.grad) bulk_s2g(a.grad + (size_t)w0 * n, s_g, (uint32_t)(kWinPerCta * n * sizeof(float)));
            if (a.gp_hi) {
                bulk_s2g(a.gp_hi + (size_t)w0 * ns, s_gh, (uint32_t)(kWinPerCta * ns * sizeof(float)));
                bulk_s2g(a.gp_lo + (size_t)w0 * ns, s_gl, (uint32_t)(kWinPerCta * ns * sizeof(float)));
            }
            bulk_commit();
            bulk_wait_read0();
        }
    } else {
        if (a.grad)
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
                       unsigned long long* patch_valid, int planar) {
    if (W <= 0) return GEM_OK;
    GEM_REQUIRE(planar >= 0 && planar <= 2 && (!planar || Wd % 4 == 0), "planar heat maps need a width that is a multiple of 4");
    GEM_REQUIRE(planar != 2 || (Wd % kTileW == 0 && H % kTileH == 0), "tiled heat maps need H % 4 == 0 and W % 8 == 0");
    GEM_REQUIRE(skel != nullptr && skel->num_joints == J, "skeleton not set for this joint count");
    GEM_REQUIRE(wt.reproj == 0.f || cam != nullptr, "camera not set");
    GEM_REQUIRE(T * J <= kSlot, "T*J must be <= 160");
    GEM_REQUIRE(T >= 3, "seq_len must be >= 3");
    GEM_REQUIRE(grad != nullptr || gp_hi != nullptr, "no gradient output requested");
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
    a.W = W, a.T = T, a.J = J, a.H = H, a.Wd = Wd, a.planar = planar;
    a.w3d = wt.w3d, a.ws = wt.smooth, a.wb = wt.bone, a.wv = wt.vae, a.wr = wt.reproj;
    const size_t pair_bytes = (size_t)kWinPerCta * T * J * 3 * sizeof(float);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    a.use_bulk = (pair_bytes % 16 == 0) && al16(pose) && al16(pose0) && al16(grad) &&
                 (!gp_hi || (al16(gp_hi) && al16(gp_lo)));
    const int grid = (W + kWinPerCta - 1) / kWinPerCta;
    // the reference's configuration as compile-time constants (identical results: tests/test_gpu_kernels.py runs both)
    static const bool fixed_env = [] { const char* e = getenv("GEM_ENERGY_FIXED"); return !e || e[0] != '0'; }();
    const bool fixed_ok = g_energy_fixed >= 0 ? g_energy_fixed != 0 : fixed_env;
    const bool ref_shape = fixed_ok && T == 10 && J == 15 && H == 64 && Wd == 64;
    if (ref_shape && a.cam.n_poly == 11) energy_grad_kernel<ShapeFix<10, 15, 64, 64, 11>><<<grid, kThreads, 0, stream>>>(a);
    else if (ref_shape && a.cam.n_poly == 14) energy_grad_kernel<ShapeFix<10, 15, 64, 64, 14>><<<grid, kThreads, 0, stream>>>(a);
    else energy_grad_kernel<ShapeDyn><<<grid, kThreads, 0, stream>>>(a);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

// fetches, into the joints' cache windows, the texels the energy evaluation of `pose` will sample (zero-copy maps)
int launch_texel_prefetch(cudaStream_t stream, const CameraConst* cam, int W, int T, int J, int H, int Wd, const float* pose,
                          const float* heat, const int64_t* frame_base, float* patch, short2* patch_origin,
                          unsigned long long* patch_valid, unsigned long long* patch_stats, int ctas, int planar, int threads) {
    if (W <= 0) return GEM_OK;
    GEM_REQUIRE(threads >= 32 && threads <= 512 && threads % 32 == 0, "texel prefetch: 32..512 threads per CTA");
    GEM_REQUIRE(planar >= 0 && planar <= 2 && (!planar || Wd % 4 == 0), "planar heat maps need a width that is a multiple of 4");
    GEM_REQUIRE(planar != 2 || (Wd % kTileW == 0 && H % kTileH == 0), "tiled heat maps need H % 4 == 0 and W % 8 == 0");
    GEM_REQUIRE(cam && heat && frame_base && patch && patch_origin && patch_valid, "texel prefetch needs the camera, the maps and the cache");
    EnergyArgs a;
    memset(&a, 0, sizeof(a));
    a.cam = *cam;
    a.pose = pose, a.heat = heat, a.frame_base = frame_base;
    a.patch = patch, a.patch_origin = patch_origin, a.patch_valid = patch_valid, a.patch_stats = patch_stats;
    a.W = W, a.T = T, a.J = J, a.H = H, a.Wd = Wd, a.planar = planar;
    const size_t total = (size_t)W * T * J;
    int grid = (int)((total + threads - 1) / threads);
    if (grid > ctas) grid = ctas;
    texel_prefetch_kernel<<<grid, threads, 0, stream>>>(a);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

// planar maps: probe (all joints) + fetch (the probe's miss list); miss_count = {entries, CTAs done}, both zero on entry
int launch_texel_probe_fetch(cudaStream_t stream, const CameraConst* cam, int W, int T, int J, int H, int Wd, const float* pose,
                             const float* heat, const int64_t* frame_base, float* patch, short2* patch_origin,
                             unsigned long long* patch_valid, unsigned long long* patch_stats, int ctas, int threads,
                             uint32_t* miss_count, uint2* miss_list, int rows, int layout) {
    if (W <= 0) return GEM_OK;
    GEM_REQUIRE(layout == 1 || layout == 2, "probe + fetch read planar (1) or tiled (2) maps");
    GEM_REQUIRE(Wd % 4 == 0, "planar heat maps need a width that is a multiple of 4");
    GEM_REQUIRE(layout != 2 || (Wd % kTileW == 0 && H % kTileH == 0), "tiled heat maps need H % 4 == 0 and W % 8 == 0");
    GEM_REQUIRE(cam && heat && frame_base && patch && patch_origin && patch_valid && miss_count && miss_list,
                "texel prefetch needs the camera, the maps, the cache and the miss list");
    GEM_REQUIRE(ctas >= 1 && threads >= 32 && threads <= 512 && threads % 32 == 0, "texel fetch: 32..512 threads per CTA");
    GEM_REQUIRE(H < 32767 && Wd < 32767, "map side beyond the miss list's 16-bit coordinates");
    EnergyArgs a;
    memset(&a, 0, sizeof(a));
    a.cam = *cam;
    a.pose = pose, a.heat = heat, a.frame_base = frame_base;
    a.patch = patch, a.patch_origin = patch_origin, a.patch_valid = patch_valid, a.patch_stats = patch_stats;
    a.W = W, a.T = T, a.J = J, a.H = H, a.Wd = Wd, a.planar = layout;
    const size_t total = (size_t)W * T * J;
    texel_probe_kernel<<<(unsigned)((total + 127) / 128), 128, 0, stream>>>(a, miss_count, miss_list);
    GEM_CHECK_LAUNCH();
    constexpr int kMaxRows = 32 / (kPlanarW / 4);                           // rows one warp can fetch per event
    if (layout == 2) texel_fetch_tiles_kernel<<<ctas, threads, 0, stream>>>(a, miss_count, miss_list);
    else if (rows >= 8 && kMaxRows >= 8) texel_fetch_kernel<(kMaxRows >= 8 ? 8 : 2)><<<ctas, threads, 0, stream>>>(a, miss_count, miss_list);
    else if (rows >= 4 && kMaxRows >= 4) texel_fetch_kernel<(kMaxRows >= 4 ? 4 : 2)><<<ctas, threads, 0, stream>>>(a, miss_count, miss_list);
    else texel_fetch_kernel<2><<<ctas, threads, 0, stream>>>(a, miss_count, miss_list);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

}  // namespace gem
