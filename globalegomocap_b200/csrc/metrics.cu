// Per-pose similarity alignment for the error metrics `main` returns.
//
// Reference: calculate_errors.py:114-179 (`calculate_errors`: the `aligned_*` and `bone_length_aligned_*` MPJPEs align
// every estimated pose to its ground truth on its own), utils/rigid_transform_with_scale.py:18-43 (`umeyama`) and
// utils/skeleton.py:123-135 (`_skeleton_resize` with the mean skeleton's bone lengths).  The reference loops over the
// frames in Python, one 3x3 numpy SVD per pose and variant; here one thread owns a frame (J <= 32 joints), float64
// throughout.  The rotation V W^T of the SVD C = V S W^T (with the reference's reflection fix) and sum(S) do not depend
// on the SVD algorithm, so a one-sided Jacobi on the 3x3 covariance reproduces numpy's LAPACK result to rounding.
#include "kernels.cuh"

namespace gem {

namespace {

constexpr int kMetJ = 32;

struct MetricArgs {
    const double* est;       // [N][J][3]
    const double* gt;        // [N][J][3]
    double* aligned;         // optional [N][J][3]: c est R + t
    double* gt_out;          // optional [N][J][3]: the (resized) ground truth the errors refer to
    double* err;             // [N][J]: |aligned - gt_out|
    double bone[kMetJ];      // bone lengths in mm (resize) — used when resize != 0
    int parent[kMetJ];
    int N, J, resize;
};

// Skeleton._skeleton_resize (utils/skeleton.py:123-135): bones rescaled along the kinematic chain, joint by joint
__device__ void resize_pose(double (*p)[3], const MetricArgs& a) {
    double vec[kMetJ][3];
    for (int i = 0; i < a.J; ++i) {
        const int pa = a.parent[i];
        const double vx = p[i][0] - p[pa][0], vy = p[i][1] - p[pa][1], vz = p[i][2] - p[pa][2];
        const double len = sqrt(vx * vx + vy * vy + vz * vz);
        const double m = i == 0 ? 0.0 : a.bone[i] / len;
        vec[i][0] = vx * m / 1000, vec[i][1] = vy * m / 1000, vec[i][2] = vz * m / 1000;
    }
    for (int i = 0; i < a.J; ++i) {
        const int pa = a.parent[i];
        p[i][0] = p[pa][0] + vec[i][0], p[i][1] = p[pa][1] + vec[i][1], p[i][2] = p[pa][2] + vec[i][2];
    }
}

// one-sided Jacobi SVD of a 3x3 matrix: A = U diag(s) V^T, U's columns orthonormal (for rank-deficient A the
// null directions are completed by cross products)
__device__ void svd3(const double A[3][3], double U[3][3], double s[3], double V[3][3]) {
    double B[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) B[i][j] = A[i][j], V[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int i = 0; i < 3; ++i) alpha += B[i][p] * B[i][p], beta += B[i][q] * B[i][q], gamma += B[i][p] * B[i][q];
                off = fmax(off, fabs(gamma) / sqrt(fmax(alpha * beta, 1e-300)));
                if (gamma == 0.0) continue;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
                for (int i = 0; i < 3; ++i) {
                    const double bp = B[i][p], bq = B[i][q];
                    B[i][p] = c * bp - sn * bq, B[i][q] = sn * bp + c * bq;
                    const double vp = V[i][p], vq = V[i][q];
                    V[i][p] = c * vp - sn * vq, V[i][q] = sn * vp + c * vq;
                }
            }
        if (off < 1e-15) break;
    }
    for (int j = 0; j < 3; ++j) {
        s[j] = sqrt(B[0][j] * B[0][j] + B[1][j] * B[1][j] + B[2][j] * B[2][j]);
        for (int i = 0; i < 3; ++i) U[i][j] = s[j] > 0 ? B[i][j] / s[j] : 0.0;
    }
    // sort singular values in decreasing order (numpy's convention; the reflection fix flips the LAST one)
    for (int a = 0; a < 2; ++a)
        for (int b = a + 1; b < 3; ++b)
            if (s[b] > s[a]) {
                const double ts = s[a]; s[a] = s[b], s[b] = ts;
                for (int i = 0; i < 3; ++i) {
                    const double tu = U[i][a]; U[i][a] = U[i][b], U[i][b] = tu;
                    const double tv = V[i][a]; V[i][a] = V[i][b], V[i][b] = tv;
                }
            }
    if (s[2] <= 1e-13 * s[0]) {                           // rank <= 2 (to rounding): complete U's last column
        U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
        U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
        U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    }
}

__device__ double det3(const double M[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

__global__ void __launch_bounds__(64) pose_align_kernel(MetricArgs a) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.N) return;
    const int J = a.J;
    double P[kMetJ][3], Q[kMetJ][3];
    for (int i = 0; i < J; ++i)
        for (int c = 0; c < 3; ++c) P[i][c] = a.est[((size_t)f * J + i) * 3 + c], Q[i][c] = a.gt[((size_t)f * J + i) * 3 + c];
    if (a.resize) resize_pose(P, a), resize_pose(Q, a);
    // umeyama (utils/rigid_transform_with_scale.py:18-43): C = (P - mp)^T (Q - mq) / n
    double mp[3] = {0, 0, 0}, mq[3] = {0, 0, 0};
    for (int i = 0; i < J; ++i)
        for (int c = 0; c < 3; ++c) mp[c] += P[i][c], mq[c] += Q[i][c];
    for (int c = 0; c < 3; ++c) mp[c] /= J, mq[c] /= J;
    double C[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, varp = 0.0;
    for (int i = 0; i < J; ++i)
        for (int r = 0; r < 3; ++r) {
            const double pr = P[i][r] - mp[r];
            varp += pr * pr;
            for (int c = 0; c < 3; ++c) C[r][c] += pr * (Q[i][c] - mq[c]);
        }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C[r][c] /= J;
    varp /= J;                                            // np.var(P, axis=0).sum()
    double U[3][3], S[3], V[3][3];
    svd3(C, U, S, V);                                     // C = U S V^T  (numpy: V, S, W^T = svd(C) with W^T = V^T here)
    double Vt[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) Vt[r][c] = V[c][r];
    if (det3(U) * det3(Vt) < 0.0) {
        S[2] = -S[2];
        for (int r = 0; r < 3; ++r) U[r][2] = -U[r][2];
    }
    double R[3][3];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[r][c] = U[r][0] * Vt[0][c] + U[r][1] * Vt[1][c] + U[r][2] * Vt[2][c];
    const double sc = (S[0] + S[1] + S[2]) / varp;
    double t[3];
    for (int c = 0; c < 3; ++c) t[c] = mq[c] - sc * (mp[0] * R[0][c] + mp[1] * R[1][c] + mp[2] * R[2][c]);
    for (int i = 0; i < J; ++i) {
        double e2 = 0.0;
        for (int c = 0; c < 3; ++c) {
            const double v = (P[i][0] * R[0][c] + P[i][1] * R[1][c] + P[i][2] * R[2][c]) * sc + t[c];
            if (a.aligned) a.aligned[((size_t)f * J + i) * 3 + c] = v;
            if (a.gt_out) a.gt_out[((size_t)f * J + i) * 3 + c] = Q[i][c];
            const double d = v - Q[i][c];
            e2 += d * d;
        }
        a.err[(size_t)f * J + i] = sqrt(e2);
    }
}

}  // namespace

int launch_pose_align(cudaStream_t stream, int N, int J, const double* est, const double* gt, const int32_t* parents_h,
                      const double* bone_len_mm_h, double* aligned, double* gt_out, double* err) {
    if (N <= 0) return GEM_OK;
    GEM_REQUIRE(est && gt && err && parents_h, "NULL argument");
    GEM_REQUIRE(J >= 3 && J <= kMetJ, "3 <= joints <= 32");
    MetricArgs a;
    a.est = est, a.gt = gt, a.aligned = aligned, a.gt_out = gt_out, a.err = err;
    a.N = N, a.J = J, a.resize = bone_len_mm_h != nullptr;
    for (int i = 0; i < kMetJ; ++i) {
        a.parent[i] = i < J ? parents_h[i] : 0;
        a.bone[i] = (bone_len_mm_h && i < J) ? bone_len_mm_h[i] : 0.0;
        GEM_REQUIRE(a.parent[i] >= 0 && a.parent[i] < J, "bad parent index");
    }
    pose_align_kernel<<<(N + 63) / 64, 64, 0, stream>>>(a);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

}  // namespace gem
