// Initial 3-D lift: heat-map argmax + per-joint depth -> local skeleton, the step that produces
// the optimiser's `estimated_local_skeleton` input.
//
// Reference: utils/skeleton.py:33-46 (Skeleton.set_skeleton), :80-82 (set_skeleton_from_file: cv2 nearest resize of
// the 64x64 maps to 1024x1024 and 128-pixel padding left and right), :176-204 (get_max_preds) and
// utils/fisheye/FishEyeCalibrated.py:18-33 (camera2world, float64).  The resize and the padding are not
// materialised: the first row-major maximum of the resized, padded image is the top-left pixel of the block of the
// source map's first row-major maximum, (up x0 + pad, up y0); a map without a positive texel has its maximum in
// the zero padding and is masked to pixel (0, 0) (get_max_preds' pred_mask).
//
// The kernel is a pure stream over the maps in their pickle layout ([frame][H][W][J], HWC): one CTA per frame, the
// frame travels through a ring of shared-memory buffers filled by the TMA engine (cp.async.bulk + mbarrier, 30 KB per
// copy), every thread keeps the running (value, first index) of all J channels for its pixels, and a shuffle / shared
// memory reduction leaves the J maxima; thread j then does the float64 back-projection.  Algorithmic bytes: the maps,
// once (245,760 B per frame).  Compiled with --fmad=false: the float64 chain rounds like numpy's separate operations.
#include "kernels.cuh"

namespace gem {

namespace {

constexpr int kLiftThreads = 256;
constexpr int kLiftPix = 512;            // pixels per staged chunk (two per thread)
constexpr int kLiftBuf = 3;              // chunks in flight per CTA
constexpr int kLiftMaxJ = 16;
constexpr int kLiftMaxPoly = 16;

struct LiftArgs {
    const float* heat;       // [N][H*W][J]
    const double* depth;     // [N][J]
    double* points;          // [N][J][3]
    float* preds;            // optional [N][J][2]
    float* maxvals;          // optional [N][J]
    int32_t* argmax;         // optional [N][J]: y0 * W + x0 of the source map's first maximum
    double poly[kLiftMaxPoly];   // polynomialC2W, constant term first
    int n_poly;
    double cx, cy;
    int N, HW, Wd, J, up, pad_x;
};

__device__ __forceinline__ bool better(float v2, int i2, float v1, int i1) { return v2 > v1 || (v2 == v1 && i2 < i1); }

__global__ void __launch_bounds__(kLiftThreads) lift_kernel(LiftArgs a) {
    extern __shared__ __align__(16) float sbuf[];                 // [kLiftBuf][kLiftPix * J]
    __shared__ __align__(8) uint64_t full_bar[kLiftBuf];
    __shared__ float s_val[kLiftThreads / 32][kLiftMaxJ];
    __shared__ int s_idx[kLiftThreads / 32][kLiftMaxJ];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int f = blockIdx.x;
    const int J = a.J;
    const int chunks = a.HW / kLiftPix;
    const uint32_t chunk_bytes = (uint32_t)(kLiftPix * J * sizeof(float));
    const float* src = a.heat + (size_t)f * a.HW * J;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kLiftBuf; ++i) mbar_init(&full_bar[i], 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        for (int c = 0; c < kLiftBuf && c < chunks; ++c) {
            mbar_arrive_expect_tx(&full_bar[c], chunk_bytes);
            bulk_g2s(sbuf + (size_t)c * kLiftPix * J, src + (size_t)c * kLiftPix * J, chunk_bytes, &full_bar[c]);
        }
    }
    float best[kLiftMaxJ];
    int bidx[kLiftMaxJ];
#pragma unroll
    for (int j = 0; j < kLiftMaxJ; ++j) best[j] = -INFINITY, bidx[j] = 0x7fffffff;
    for (int c = 0; c < chunks; ++c) {
        const int slot = c % kLiftBuf;
        mbar_wait(&full_bar[slot], (uint32_t)((c / kLiftBuf) & 1));
        const float* s = sbuf + (size_t)slot * kLiftPix * J;
#pragma unroll
        for (int h = 0; h < kLiftPix / kLiftThreads; ++h) {
            const int pl = tid + h * kLiftThreads;                // pixel inside the chunk (stride J words: no bank conflicts for odd J)
            const int p = c * kLiftPix + pl;                      // pixel of the frame, increasing per thread
            const float* px = s + pl * J;
#pragma unroll
            for (int j = 0; j < kLiftMaxJ; ++j) {
                if (j < J) {
                    const float v = px[j];
                    if (v > best[j]) best[j] = v, bidx[j] = p;    // strict: the first maximum stays
                }
            }
        }
        __syncthreads();                                          // the whole CTA has consumed the slot
        if (tid == 0 && c + kLiftBuf < chunks) {
            mbar_arrive_expect_tx(&full_bar[slot], chunk_bytes);
            bulk_g2s(sbuf + (size_t)slot * kLiftPix * J, src + (size_t)(c + kLiftBuf) * kLiftPix * J, chunk_bytes, &full_bar[slot]);
        }
    }
    // J arg-max reductions: shuffles inside each warp, then one value per warp
#pragma unroll
    for (int j = 0; j < kLiftMaxJ; ++j) {
        if (j < J) {
            float v = best[j];
            int i = bidx[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float v2 = __shfl_xor_sync(0xffffffffu, v, o);
                const int i2 = __shfl_xor_sync(0xffffffffu, i, o);
                if (better(v2, i2, v, i)) v = v2, i = i2;
            }
            if (lane == 0) s_val[warp][j] = v, s_idx[warp][j] = i;
        }
    }
    __syncthreads();
    if (tid < J) {
        const int j = tid;
        float v = s_val[0][j];
        int i = s_idx[0][j];
        for (int w = 1; w < kLiftThreads / 32; ++w)
            if (better(s_val[w][j], s_idx[w][j], v, i)) v = s_val[w][j], i = s_idx[w][j];
        // get_max_preds on the resized, padded image (utils/skeleton.py:176-204)
        const bool pos = v > 0.f;
        const float px = pos ? (float)((i % a.Wd) * a.up + a.pad_x) : 0.f;
        const float py = pos ? (float)((i / a.Wd) * a.up) : 0.f;
        const size_t o = (size_t)f * J + j;
        if (a.preds) a.preds[o * 2 + 0] = px, a.preds[o * 2 + 1] = py;
        if (a.maxvals) a.maxvals[o] = pos ? v : 0.f;              // (the padding's zeros are the maximum otherwise)
        if (a.argmax) a.argmax[o] = i;
        // camera2world (FishEyeCalibrated.py:18-33), float64
        const double x = (double)px - a.cx, y = (double)py - a.cy;
        const double dist = sqrt(x * x + y * y);
        double z = 0.0;                                           // np.polyval(p[::-1], dist): Horner from the top
        for (int k = a.n_poly - 1; k >= 0; --k) z = z * dist + a.poly[k];
        const double nz = -z;
        const double norm = sqrt((x * x + y * y) + nz * nz);
        const double d = a.depth[o];
        a.points[o * 3 + 0] = x / norm * d;
        a.points[o * 3 + 1] = y / norm * d;
        a.points[o * 3 + 2] = nz / norm * d;
    }
}

}  // namespace

int launch_lift(cudaStream_t stream, int N, int H, int Wd, int J, const float* heat, const double* depth,
                const double* poly_c2w, int n_poly, double cx, double cy, int up, int pad_x, double* points, float* preds,
                float* maxvals, int32_t* argmax) {
    if (N <= 0) return GEM_OK;
    GEM_REQUIRE(heat && depth && points && poly_c2w, "NULL argument");
    GEM_REQUIRE(J >= 1 && J <= kLiftMaxJ, "at most 16 joints");
    GEM_REQUIRE(n_poly >= 1 && n_poly <= kLiftMaxPoly, "at most 16 polynomial coefficients");
    GEM_REQUIRE((H * Wd) % kLiftPix == 0 && H * Wd >= kLiftPix, "H*W must be a multiple of 512");
    GEM_REQUIRE((reinterpret_cast<uintptr_t>(heat) & 15u) == 0, "heat maps must be 16-byte aligned");
    LiftArgs a;
    a.heat = heat, a.depth = depth, a.points = points, a.preds = preds, a.maxvals = maxvals, a.argmax = argmax;
    for (int i = 0; i < kLiftMaxPoly; ++i) a.poly[i] = i < n_poly ? poly_c2w[i] : 0.0;
    a.n_poly = n_poly, a.cx = cx, a.cy = cy;
    a.N = N, a.HW = H * Wd, a.Wd = Wd, a.J = J, a.up = up, a.pad_x = pad_x;
    const size_t smem = (size_t)kLiftBuf * kLiftPix * J * sizeof(float);
    static PerDeviceOnce attr_set;
    if (bool* once_ = attr_set.flag(); !*once_) {
        GEM_CUDA(cudaFuncSetAttribute(lift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(kLiftBuf * kLiftPix * kLiftMaxJ * sizeof(float))));
        *once_ = true;
    }
    lift_kernel<<<N, kLiftThreads, smem, stream>>>(a);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

}  // namespace gem
