// Launch wrappers shared between the translation units of libgem_b200.so.
#pragma once
#include "common.cuh"

namespace gem {

// cam / skel: host copies owned by the calling ctx (passed to the kernel by value; cam may be NULL when reproj == 0)
int launch_energy_grad(cudaStream_t stream, const CameraConst* cam, const SkeletonConst* skel, int W, int T, int J, int H,
                       int Wd, const float* pose, const float* pose0,
                       const float* heat, const int64_t* frame_base, const int32_t* clip, const float* mean_bone,
                       const gem_energy_weights& wt, float* energy, float* terms, float* grad, uint32_t* status,
                       float* gp_hi = nullptr, float* gp_lo = nullptr, int pp = 0, float* patch = nullptr,
                       short2* patch_origin = nullptr, unsigned long long* patch_stats = nullptr, int gp_f16 = 0,
                       int32_t* row_exp = nullptr, unsigned long long* patch_valid = nullptr, int planar = 0);
int launch_texel_prefetch(cudaStream_t stream, const CameraConst* cam, int W, int T, int J, int H, int Wd, const float* pose,
                          const float* heat, const int64_t* frame_base, float* patch, short2* patch_origin,
                          unsigned long long* patch_valid, unsigned long long* patch_stats, int ctas, int planar = 0,
                          int threads = 512);
int launch_texel_probe_fetch(cudaStream_t stream, const CameraConst* cam, int W, int T, int J, int H, int Wd, const float* pose,
                             const float* heat, const int64_t* frame_base, float* patch, short2* patch_origin,
                             unsigned long long* patch_valid, unsigned long long* patch_stats, int ctas, int threads,
                             uint32_t* miss_count, uint2* miss_list, int rows, int layout);
constexpr int kPatchW = kPatchWd;

enum { EPI_NONE = 0, EPI_LRELU = 1, EPI_MASK = 2 };
struct TapGemmArgs {
    const float* A;      // [M][lda]
    const void* A_hi = nullptr;    // tensor-core path only: A already split into hi / lo parts ([M][lda] each; fp32 TF32
    const void* A_lo = nullptr;    //   values, or fp16 with the fp16 scheme)
    const int32_t* row_exp = nullptr;   // fp16 scheme: row m of A carries a factor 2^row_exp[m] that the epilogue removes
    float* C_lo = nullptr;         // tensor-core path only: write the result split (hi to C, lo to C_lo)
    uint32_t* C_sign = nullptr;    // tensor-core path only: packed sign bits of the result [M][N/32]
    int out16 = 0;                 // with C_lo: write the result as fp16 hi / scaled fp16 lo (uint16 arrays) instead
    uint32_t* row_flag = nullptr;  // optional [M]: OR-ed with GEM_WIN_F16_RANGE when a row's fp16 output left fp16's range
    const float* B;      // [taps][K][ldb]
    const float* bias;   // [N] or NULL
    const float* aux;    // [M][ldaux] saved activation for EPI_MASK ...
    const uint32_t* aux_bits = nullptr;   // ... or its packed sign bits [M][N/32] (N % 32 == 0)
    float* C;            // [M][ldc]
    int M, N, K, taps, T;
    int lda, ldb, ldc, ldaux;   // ldb = row stride of B[tap][k][:] (N rounded up to a multiple of 4, zero padded)
    int epi;
};
int launch_tap_gemm_simt(cudaStream_t stream, const TapGemmArgs& g);
// scheme: 1 = 3xTF32, 2 = fp16 hi + scaled fp16 lo (kind::f16, cross terms in their own accumulator)
int launch_tap_gemm_tc(cudaStream_t stream, const TapGemmArgs& g, void* owner, size_t scheme);
extern int g_gemm_pair;
extern int g_energy_fixed;   // debug override of GEM_ENERGY_FIXED (gem_debug_energy_fixed): -1 = environment / default
extern long long* g_gemm_dbg;
int tc_gemm_prepare_weight(void* owner, cudaStream_t stream, const float* B, int ldb, int K, int N, int scheme = 1);
int launch_split_f16(cudaStream_t stream, const float* A, int lda, int M, int K, const int32_t* row_exp, uint16_t* hi,
                     uint16_t* lo, uint32_t* row_flag = nullptr);
int launch_rowscale_split_f16(cudaStream_t stream, const float* xh, const float* xl, int M, int K, uint16_t* hi,
                              uint16_t* lo, int32_t* row_exp);
int launch_split_tf32(cudaStream_t stream, const float* A, int lda, int M, int K, float* hi, float* lo);
void tc_gemm_forget_weight(void* owner, const float* B);
void tc_gemm_release(void* owner);
bool tc_gemm_available();

// tcgen05 kernel for the k=3 convolutions (gemm_tap_tc.cu); activations travel as TF32 hi / lo pairs
struct TapTcLaunch {
    const float* B;                 // the layer's weight pointer (key of tc_tap_prepare_weight)
    int scheme = 1;                 // 1 = 3xTF32 (fp32 hi / lo activations), 2 = fp16 hi + scaled fp16 lo (uint16 arrays)
    const void *A_hi, *A_lo;        // [W*T][lda]
    int lda, Kreal;                 // row pitch and valid columns (the rest of a 32-wide K block reads zero)
    const float* bias;              // [N] or NULL
    const float* aux;               // EPI_MASK: [W*T][ldaux], only the sign is used ...
    int ldaux;
    const uint32_t* aux_bits = nullptr;   // ... or its packed sign bits [W*T][N/32] (preferred)
    uint32_t* sign_out = nullptr;         // optional: packed sign bits of the output [W*T][N/32]
    void *out_hi, *out_lo;          // [W*T][ldo] (scheme's element type); out_lo == NULL writes the plain fp32 result to out_hi
    int ldo;
    int W, T, epi;
};
bool tc_tap_supported(int K, int N, int T);
int tc_tap_prepare_weight(void* owner, cudaStream_t stream, const float* B, int ldb, int K, int N, int scheme = 1);
int launch_rowscale_split_pad_f16(cudaStream_t stream, const float* src, int C, int T, int W, int ldo, uint16_t* hi,
                                  uint16_t* lo, int32_t* row_exp);
int launch_tap_tc(cudaStream_t stream, void* owner, const TapTcLaunch& L);
// Up to four consecutive k=3 layers in ONE launch (fp16 scheme): the activation tile stays in shared memory between
// the layers (the epilogue's staged output IS the next layer's swizzled operand), only sign bits and the last
// layer's output reach HBM; weights stream through a two-slot ring.  Layers after the first must be 64 -> N with
// N = 64 for every layer but the last (which may be 48-padded plain fp32 output, or N = 128 split output).
constexpr int kTapChainMax = 5;
// Optional energy prologue of a backward chain on CTA pairs (launch_tap_chain_pair): the chain kernel evaluates the
// fused energy + gradient itself (energy_device.cuh) from `pose` and feeds dE/dpose straight into its first layer.
struct ChainEnergyLaunch {
    const CameraConst* cam;                     // host copies owned by the ctx (cam may be NULL when wt.reproj == 0)
    const SkeletonConst* skel;
    const float *pose, *pose0, *heat;           // [W][T][J][3] decoded pose and anchor; heat maps (NULL without reproj)
    const int64_t* frame_base;
    const int32_t* clip;
    const float* mean_bone;
    gem_energy_weights wt;
    float* energy;                              // [W]
    uint32_t* status;                           // [W] or NULL
    int32_t* row_exp;                           // [W]
    int J, H, Wd;
    int planar = 0;                             // heat-map layout (gem_ctx_set_heat_layout)
    float* patch = nullptr;                     // optional texel cache (zero-copy heat maps)
    short2* patch_origin = nullptr;
    unsigned long long* patch_valid = nullptr;
    unsigned long long* patch_stats = nullptr;
};
bool tap_chain_energy_supported(int T, int J, int first_layer_k);
struct TapChainLaunch {
    int nl;
    const float* B[kTapChainMax];               // weight pointers (keys of tc_tap_prepare_weight, scheme 2)
    const float* bias[kTapChainMax];            // [N] or NULL
    const uint32_t* aux_bits[kTapChainMax];     // EPI_MASK: packed sign bits of the saved activation [W*T][N/32]
    uint32_t* sign_out[kTapChainMax];           // optional: packed sign bits of the layer's output [W*T][N/32]
    int epi[kTapChainMax];
    const void *A_hi, *A_lo;                    // first layer's input [W*T][lda] fp16 hi / lo
    int lda, Kreal;
    void *out_hi, *out_lo;                      // last layer's output [W*T][ldo]: split fp16, or plain fp32 when out_lo == NULL
    int ldo;
    int W, T;
    uint32_t* status = nullptr;                 // optional [W]: OR-ed with GEM_WIN_F16_RANGE when a split output saturates
    const ChainEnergyLaunch* energy = nullptr;  // launch_tap_chain_pair only: energy prologue instead of loading A_hi / A_lo
};
int launch_tap_chain(cudaStream_t stream, void* owner, const TapChainLaunch& L);
// the same on CTA pairs (cta_group::2): up to five layers with up to 256 channels in and out (gemm_tap_tc.cu)
int launch_tap_chain_pair(cudaStream_t stream, void* owner, const TapChainLaunch& L);
int launch_split_pad(cudaStream_t stream, const float* src, int C, size_t tokens, int ldo, float* hi, float* lo);
void tc_tap_forget_weight(void* owner, const float* B);
void tc_tap_release(void* owner);
extern long long* g_tap_dbg;   // debug: per-CTA phase timestamps of the tap kernel (NULL in production)

struct LbfgsWin;
struct LbfgsBuffers {
    LbfgsWin* st;
    float *X, *D, *G, *GP, *BG0, *BG1, *ZT;   // [W][n]  (prev_flat_grad is G itself, see lbfgs.cu)
    float *ZT_hi, *ZT_lo;                           // optional [W][n]: the trial point as TF32 hi / lo parts, or,
    int zt_f16;                                     //   when zt_f16, as fp16 hi / scaled lo (uint16 [W][n] in the same buffers)
    uint32_t* status;                               // optional [W]: OR-ed with GEM_WIN_F16_RANGE when a trial point saturates
    float *Y, *S;                                   // [W][m][n]
    float* RO;                                      // [W][m]
    float* trace;                                   // [W][trace_stride] or NULL
    int n, m, trace_stride;
    double lr, tol_grad, tol_change;
    int max_iter, max_eval;
};
size_t lbfgs_state_bytes();
int launch_lbfgs_begin(cudaStream_t stream, const LbfgsBuffers& b, const float* z0, int W);
int launch_lbfgs_advance(cudaStream_t stream, const LbfgsBuffers& b, const float* loss, const float* grad, int W);
int launch_lbfgs_stats(cudaStream_t stream, const LbfgsBuffers& b, int W, int32_t* n_iter, int32_t* evals,
                       int32_t* finished, double* t, double* loss);

// initial 3-D lift (lift.cu): heat-map argmax + depth -> local skeleton
int launch_lift(cudaStream_t stream, int N, int H, int Wd, int J, const float* heat, const double* depth,
                const double* poly_c2w, int n_poly, double cx, double cy, int up, int pad_x, double* points, float* preds,
                float* maxvals, int32_t* argmax);

// per-pose similarity alignment of the error metrics (metrics.cu)
int launch_pose_align(cudaStream_t stream, int N, int J, const double* est, const double* gt, const int32_t* parents_h,
                      const double* bone_len_mm_h, double* aligned, double* gt_out, double* err);

int launch_reparam(cudaStream_t stream, const float* fc, const float* eps, size_t eps_stride, float* z0, float* mu,
                   float* sd, int W, int n);
int launch_transform(cudaStream_t stream, int W, int T, int J, const void* pose, int pose_is_f64, const double* cams,
                     double* out64, float* out32, int mode);
int launch_merge(cudaStream_t stream, int W, int T, int overlap, int row, const double* win, double* out);
int launch_gauss(cudaStream_t stream, int N, int row, double sigma, const double* seq, double* out);

}  // namespace gem
