// Small pose-space kernels around the two optimisation stages: reparameterisation, SLAM camera
// transforms (float64, like the reference's numpy), overlap-average stitching, Gaussian smoothing.
#include "kernels.cuh"

namespace gem {

// z0 = eps * exp(0.5 * logvar) + mu        (SeqConvVAE.py:159-169, 184-189); fc = [W][2n] (mu | logvar)
__global__ void reparam_kernel(const float* __restrict__ fc, const float* __restrict__ eps, size_t eps_stride,
                               float* __restrict__ z0, float* __restrict__ mu_out, float* __restrict__ std_out, int W,
                               int n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)W * n) return;
    const size_t w = i / n, c = i - w * n;
    const float mu = fc[w * 2 * n + c];
    const float lv = fc[w * 2 * n + n + c];
    const float sd = expf(0.5f * lv);
    z0[i] = eps[w * eps_stride + c] * sd + mu;
    if (mu_out) mu_out[i] = mu;
    if (std_out) std_out[i] = sd;
}

int launch_reparam(cudaStream_t stream, const float* fc, const float* eps, size_t eps_stride, float* z0, float* mu,
                   float* sd, int W, int n) {
    if (W <= 0) return GEM_OK;
    const size_t total = (size_t)W * n;
    reparam_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(fc, eps, eps_stride, z0, mu, sd, W, n);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

// ---- 4x4 float64 helpers ------------------------------------------------------------------
__device__ void inv4x4(const double* a, double* out) {
    // Gauss-Jordan with partial pivoting (np.linalg.inv is LAPACK LU with partial pivoting)
    double m[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            m[r][c] = a[r * 4 + c];
            m[r][4 + c] = (r == c) ? 1.0 : 0.0;
        }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        double best = fabs(m[col][col]);
        for (int r = col + 1; r < 4; ++r)
            if (fabs(m[r][col]) > best) best = fabs(m[r][col]), piv = r;
        if (piv != col)
            for (int c = 0; c < 8; ++c) {
                const double tmp = m[col][c];
                m[col][c] = m[piv][c];
                m[piv][c] = tmp;
            }
        const double d = 1.0 / m[col][col];
        for (int c = 0; c < 8; ++c) m[col][c] *= d;
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            const double f = m[r][col];
            if (f != 0.0)
                for (int c = 0; c < 8; ++c) m[r][c] -= f * m[col][c];
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) out[r * 4 + c] = m[r][4 + c];
}

// mode 0: out = inv(C[w][0]) C[w][t] x   (utils/utils.py:99-112)
// mode 1: out = C[w][0] x                 (optimizer.py:302-308)
template <typename TIn>
__global__ void transform_kernel(const TIn* __restrict__ pose, const double* __restrict__ cams, double* out64,
                                 float* out32, int T, int J, int mode) {
    __shared__ double M[32][12];     // first three rows of each frame's matrix (T <= 32)
    const int w = blockIdx.x;
    const double* Cw = cams + (size_t)w * T * 16;
    if (threadIdx.x < T) {
        const int t = threadIdx.x;
        if (mode == 0) {
            double inv0[16];
            inv4x4(Cw, inv0);
            const double* Ct = Cw + t * 16;
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 4; ++c) {
                    double acc = 0.0;
                    for (int k = 0; k < 4; ++k) acc += inv0[r * 4 + k] * Ct[k * 4 + c];
                    M[t][r * 4 + c] = acc;
                }
        } else {
            for (int i = 0; i < 12; ++i) M[t][i] = Cw[i];
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < T * J; k += blockDim.x) {
        const int t = k / J;
        const size_t base = ((size_t)w * T * J + k) * 3;
        const double x = (double)pose[base], y = (double)pose[base + 1], z = (double)pose[base + 2];
        for (int r = 0; r < 3; ++r) {
            const double v = ((M[t][r * 4 + 0] * x + M[t][r * 4 + 1] * y) + M[t][r * 4 + 2] * z) + M[t][r * 4 + 3];
            if (out64) out64[base + r] = v;
            if (out32) out32[base + r] = (float)v;
        }
    }
}

int launch_transform(cudaStream_t stream, int W, int T, int J, const void* pose, int pose_is_f64, const double* cams,
                     double* out64, float* out32, int mode) {
    if (W <= 0) return GEM_OK;
    GEM_REQUIRE(T <= 32, "seq_len must be <= 32");
    if (pose_is_f64)
        transform_kernel<double><<<W, 160, 0, stream>>>((const double*)pose, cams, out64, out32, T, J, mode);
    else
        transform_kernel<float><<<W, 160, 0, stream>>>((const float*)pose, cams, out64, out32, T, J, mode);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

// merge_batches (optimizer.py:425-437): frame f of the merged sequence
__global__ void merge_kernel(const double* __restrict__ win, double* __restrict__ out, int W, int T, int overlap,
                             int row) {
    const int S = T - overlap;
    const int nframes = S * W + overlap;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)nframes * row) return;
    const int f = (int)(i / row), c = (int)(i - (size_t)f * row);
    int wi = f / S, r = f - wi * S;
    double v;
    if (wi >= W) {                                   // tail: last window's final `overlap` frames
        v = win[((size_t)(W - 1) * T + S + r) * row + c];
    } else if (r < overlap && wi > 0) {              // overlap: (first[-ov:] + second[:ov]) / 2
        v = (win[((size_t)(wi - 1) * T + S + r) * row + c] + win[((size_t)wi * T + r) * row + c]) / 2;
    } else {
        v = win[((size_t)wi * T + r) * row + c];
    }
    out[i] = v;
}

int launch_merge(cudaStream_t stream, int W, int T, int overlap, int row, const double* win, double* out) {
    if (W <= 0) return GEM_OK;
    GEM_REQUIRE(overlap >= 0 && overlap * 2 <= T, "overlap must be in [0, T/2]");
    const size_t total = (size_t)((T - overlap) * W + overlap) * row;
    merge_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(win, out, W, T, overlap, row);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

constexpr int kMaxRadius = 32;
struct GaussWeights {
    double w[kMaxRadius + 1];   // w[0] centre, w[j] at +-j
    int radius;
};

// scipy.ndimage.correlate1d with a symmetric kernel, mode='reflect' (d c b a | a b c d | d c b a)
__global__ void gauss_kernel(const double* __restrict__ seq, double* __restrict__ out, int N, int row, GaussWeights g) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)N * row) return;
    const int f = (int)(i / row), c = (int)(i - (size_t)f * row);
    auto at = [&](int idx) {
        const int period = 2 * N;
        idx %= period;
        if (idx < 0) idx += period;
        if (idx >= N) idx = period - 1 - idx;
        return seq[(size_t)idx * row + c];
    };
    double acc = at(f) * g.w[0];
    for (int j = g.radius; j >= 1; --j) acc += (at(f - j) + at(f + j)) * g.w[j];
    out[i] = acc;
}

int launch_gauss(cudaStream_t stream, int N, int row, double sigma, const double* seq, double* out) {
    if (N <= 0) return GEM_OK;
    GaussWeights g;
    g.radius = (int)(4.0 * sigma + 0.5);
    GEM_REQUIRE(g.radius <= kMaxRadius && sigma > 0, "sigma out of range");
    double sum = 0.0, phi[2 * kMaxRadius + 1];
    for (int x = -g.radius; x <= g.radius; ++x) {
        phi[x + g.radius] = exp(-0.5 / (sigma * sigma) * (double)(x * x));
        sum += phi[x + g.radius];
    }
    for (int j = 0; j <= g.radius; ++j) g.w[j] = phi[g.radius + j] / sum;
    const size_t total = (size_t)N * row;
    gauss_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(seq, out, N, row, g);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

}  // namespace gem
