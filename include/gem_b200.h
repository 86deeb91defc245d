/*
 * gem_b200.h — C ABI of the B200-native GlobalEgoMocap pose-sequence optimiser.
 *
 * The reference (jianwang-mpi/GlobalEgoMocap) is pure Python and has no FFI of its
 * own; its boundary for this path is the Python surface
 *     BodyPoseOptimizer.optimize_pose_seq_pytorch_LBFGS   (optimizer.py:242-276)
 *     BodyPoseOptimizer.total_loss                        (optimizer.py:226-240)
 *     optimizer.main window loop + stitching              (optimizer.py:370-450)
 * The Python shim `globalegomocap_b200.optimizer` keeps those names/signatures and
 * binds the entry points below with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns 0 on success, a negative gem_status otherwise;
 *     gem_last_error() gives a thread-local message.  No C++ exceptions cross the ABI.
 *   - all `*_d` / pointer arguments are DEVICE pointers owned by the caller unless
 *     the name ends in `_h` (host).  The library never frees caller memory; scratch
 *     (activations, L-BFGS vectors/history) is owned by the gem_ctx, sized for
 *     `max_windows` at creation.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the
 *     calls do not synchronise unless stated.
 *   - one host thread per gem_ctx; one gem_ctx per device (multi-GPU = one process
 *     per GPU).
 *   - sizes: a per-window call with W = 0 is a no-op that returns GEM_OK (its buffers
 *     may be NULL); W above the ctx's max_windows returns GEM_ERR_CAPACITY.
 *   - window geometry: T frames x J joints x 3, pose tensors are [W][T][J][3] fp32,
 *     heatmaps are the pickle's HWC layout [frames][H][Wd][J] fp32 and are gathered
 *     in place (map of window w, frame t, joint j = frame_base[w]+t, channel j;
 *     optimizer.py:251-252 permutes to t*15+j).
 */
#ifndef GEM_B200_H
#define GEM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gem_ctx gem_ctx;

enum gem_status {
    GEM_OK = 0,
    GEM_ERR_INVALID = -1,   /* bad argument */
    GEM_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
    GEM_ERR_STATE = -3,     /* call order / missing weights */
    GEM_ERR_CAPACITY = -4   /* more windows than the ctx was created for */
};

/* per-window status bits */
#define GEM_WIN_NORM_ZERO 1u /* a joint had x = y = 0 in the camera frame: the reference raises
                                Exception("norm is zero!") (FishEyeCalibrated.py:108,124-127) */
#define GEM_WIN_F16_RANGE 2u /* an activation or latent entry of the window exceeded fp16's finite range (65504) while the
                                tensor-core layers ran in the split-fp16 scheme (gemm modes 2, 3): operands saturate, the
                                result is no longer fp32-faithful.  Re-run with gem_ctx_set_gemm_mode(ctx, 1) (3xTF32). */

/* weights of total_loss in the reference's order of appearance (optimizer.py:239-240) */
typedef struct gem_energy_weights {
    float w3d;     /* weight_3d          * pose_energy_3d      (optimizer.py:210-213) */
    float smooth;  /* smooth_weight      * smooth_accelerate   (optimizer.py:202-208) */
    float bone;    /* bone_length_weight * bone_length_energy  (optimizer.py:172-177) */
    float vae;     /* vae_weight         * vae_energy(pose)    (optimizer.py:215-218,238) */
    float reproj;  /* reproj_weight      * reprojection_energy_heatmap_fast (optimizer.py:139-149);
                      exactly 0 skips the term like optimizer.py:232-235 */
} gem_energy_weights;

/* torch.optim.LBFGS(lr, max_iter, max_eval, tolerance_grad, tolerance_change,
 * line_search_fn='strong_wolfe') as constructed at optimizer.py:261-262 */
typedef struct gem_lbfgs_params {
    double lr;               /* 2 */
    int32_t max_iter;        /* 25 */
    int32_t max_eval;        /* max_iter*5/4 = 31 */
    double tolerance_grad;   /* 1e-7 */
    double tolerance_change; /* 1e-6 (outer loop; the line search keeps torch's own 1e-9) */
} gem_lbfgs_params;

/* One generic layer: out[m][n] = act( bias[n] + sum_tap sum_k in[m + tap - taps/2][k] * w[tap][k][n] ),
 * rows m are (window, frame) tokens, taps that leave the window's T frames read zeros. */
typedef struct gem_layer {
    const float* w_d;    /* [taps][k][n] fp32 */
    const float* bias_d; /* [n] or NULL */
    int32_t taps, k, n;
} gem_layer;

/* BatchNorm-folded, tap-arranged weights of one ConvVAE (SeqConvVAE.py:27-92) prepared by the
 * host shim (globalegomocap_b200/vae_prep.py).  dec[0] is decoder_input fused with decoder.0's
 * ConvTranspose1d (both linear, no activation between them): latent -> [T][256]. */
typedef struct gem_vae_weights {
    gem_layer dec[6];      /* forward: latent->T*256, 256->128, 128->64, 64->64, 64->64, 64->45 */
    gem_layer dec_bwd[6];  /* bwd-data of dec[5]..dec[0] in execution order (no bias) */
    gem_layer enc[6];      /* 45->64, 64->64, 64->128, 128->256, 256->512, T*512 -> 2*latent (mu|logvar) */
} gem_vae_weights;

int gem_version(void);
const char* gem_last_error(void);

/* ---- context ---------------------------------------------------------------------------- */
int gem_ctx_create(gem_ctx** out, int device, int max_windows, int latent_dim, int seq_len, int num_joints,
                   int heat_h, int heat_w, int max_history);
int gem_ctx_destroy(gem_ctx* ctx);
/* polynomialW2C coefficients + principal point (FishEyeCalibrated.py:8-14) */
int gem_ctx_set_camera(gem_ctx* ctx, const double* poly_h, int n_poly, double cx, double cy);
/* kinematic_parents (optimizer.py:34) */
int gem_ctx_set_skeleton(gem_ctx* ctx, const int32_t* parents_h, int num_joints);
/* which: 0 = local-stage VAE, 1 = global-stage VAE (optimizer.py:334,344); struct is copied */
int gem_ctx_set_vae(gem_ctx* ctx, int which, const gem_vae_weights* weights_h);
/* bytes of device scratch owned by the ctx */
int64_t gem_ctx_scratch_bytes(const gem_ctx* ctx);
/* 0: hand-written SIMT fp32 layers everywhere; 1: tcgen05/TMEM, 3xTF32 arithmetic in every layer; 2: like 1 but the
 * plain GEMMs (latent<->T*256, encoder fc) use the fp16 scheme: x ~ fp16 hi + 2^-11 fp16 lo, three kind::f16 MMAs
 * per product, gradient rows rescaled by a power of two — twice the tensor rate of 3xTF32 and (measured) a smaller
 * error against float64; 3 (default): the fp16 scheme in every tensor-core layer, activations travel as fp16 pairs */
int gem_ctx_set_gemm_mode(gem_ctx* ctx, int mode);

/* A stage's windows are independent: gem_solve_stage splits them into n_chunks slices (boundaries at
 * multiples of 12 windows, at least 96 windows each) that run the same kernel sequence on internal streams,
 * forked from and joined to the caller's stream, so one slice's launch gaps and HBM-bound L-BFGS updates
 * overlap another's tensor-core layers.  Default 0 = automatic (4 slices, 8 when the heat maps are host memory read over
 * PCIe; env GEM_CHUNKS); 1 = everything on the caller's stream.
 * Results do not depend on the setting.  Profiling (gem_ctx_set_profiling) forces 1.
 * Closure rounds 1.. of a stage are one CUDA graph of round 0's launches (captured once per slice and
 * configuration, cached in the ctx; env GEM_GRAPHS=0 disables). */
int gem_ctx_set_chunks(gem_ctx* ctx, int n_chunks);
/* Explicit slice boundaries instead of n_chunks equal parts: first_window_h[0] = 0 < first_window_h[1] < ...
 * (all even).  Used by calls whose W exceeds the last entry; n = 0 returns to automatic slicing.  A caller
 * that uploads its clips one after the other makes the slices the clips. */
int gem_ctx_set_slices(gem_ctx* ctx, int n, const int32_t* first_window_h);
/* One-shot, consumed by the next gem_solve_stage / gem_solve_windows: cudaEvent_t events_h[i] completes when
 * the inputs (heatmaps, poses, cameras) of all windows in [first_window_h[i], first_window_h[i+1]) are in
 * device memory.  A slice starts once every event covering windows below its end has completed, so the
 * optimisation of the first clips overlaps the host-to-device copy of the later ones. */
int gem_ctx_set_ready_events(gem_ctx* ctx, int n, const int32_t* first_window_h, void* const* events_h);

/* Layout of every heat_d argument: 0 (default) = [frames][H][Wd][J], the pickle's HWC layout, gathered in place;
 * 1 = planar [frames][J][H][Wd] — what the reference itself permutes each window's maps to before grid_sample
 * (optimizer.py:251); 2 = tiled [frames][J][H/4][Wd/8][4][8]: every map as tiles of 4 rows x 8 texels, one 128-byte
 * line per tile.  In the planar layout the x-neighbours of a bilinear footprint share a 32-byte sector (half the DRAM
 * sectors per sample on resident maps) and the zero-copy texel window fetches whole 64-byte rows; in the tiled layout
 * one PCIe request brings a joint's 4 x 8 neighbourhood (reads of host memory are charged per request, not per byte:
 * half the requests of the planar layout again).  Values are the same: results are bit-identical in all layouts.
 * Layout 1 needs Wd % 4 == 0, layout 2 H % 4 == 0 and Wd % 8 == 0.  gem_lift_skeleton always takes HWC maps. */
int gem_ctx_set_heat_layout(gem_ctx* ctx, int planar);

/* heat_d may be pinned (or registered) HOST memory: the energy kernel then reads the maps over PCIe through a
 * per-joint texel window in HBM (HWC maps: 8 x 8 texels with a valid bit per texel; planar maps: 16 x 16 with a valid
 * bit per 64-byte row; tiled maps: 16 x 16 with a valid bit per 4 x 8 tile.  A unit is fetched the first time a
 * bilinear footprint needs it — for planar and tiled maps by a probe kernel that lists the joints with a missing
 * footprint and a few fetch CTAs that walk that list before every energy evaluation; the window is re-centred,
 * empty, when the footprint leaves it), so only the few per cent of the maps the optimiser ever samples cross the
 * bus and no up-front copy is needed.  The window buffers (1 KB per joint-frame) are allocated by the first call that
 * uses them.  Values are copies: results are bit-identical.  mode -1 (default): cache on exactly when heat_d is host
 * memory; 0 off; 1 on. */
int gem_ctx_set_texel_cache(gem_ctx* ctx, int mode);
/* synchronises; returns the cache's lookups (one per joint per evaluation) and the texels it fetched from the map
 * (counted in 32-byte sectors: one per texel with HWC maps, two per window row with planar maps, four per tile with
 * tiled maps) since the previous call while counting was enabled, then clears them and sets counting on/off */
int gem_ctx_texel_cache_stats(gem_ctx* ctx, int enable, uint64_t* lookups_h, uint64_t* texels_fetched_h);

/* ---- instrumentation ---------------------------------------------------------------------- */
/* kernel classes reported by gem_ctx_read_profile */
#define GEM_TAG_DEC 100           /* +i: decoder forward layer i (0 = latent -> T*256 GEMM) */
#define GEM_TAG_DEC_BWD 200       /* +i: decoder bwd-data layer i (5 = T*256 -> latent GEMM, 6 = its row-scaled fp16 split) */
#define GEM_TAG_ENC 300           /* +i: encoder layer i (5 = T*512 -> 2*latent GEMM) */
#define GEM_TAG_ENERGY 1
#define GEM_TAG_LBFGS_BEGIN 2
#define GEM_TAG_LBFGS_ADVANCE 3
#define GEM_TAG_REPARAM 4
#define GEM_TAG_TRANSFORM 5
#define GEM_TAG_STITCH 6
#define GEM_TAG_OTHER 7
/* number of CUDA kernels this ctx has launched since creation */
int64_t gem_ctx_launch_count(const gem_ctx* ctx);
/* when enabled, every kernel launch is bracketed by a CUDA event pair on its stream */
int gem_ctx_set_profiling(gem_ctx* ctx, int enable);
/* synchronises the device, aggregates the recorded event pairs per tag into the host arrays
 * (count and total milliseconds), writes the number of distinct tags and clears the record */
int gem_ctx_read_profile(gem_ctx* ctx, int max_tags, int32_t* tags_h, int32_t* counts_h, float* total_ms_h,
                         int32_t* n_tags_h);

/* ---- the fused energy + analytic gradient (total_loss with the decoder bypassed) -------- */
/* pose_d, pose0_d, grad_d: [W][T][J][3]; frame_base_d: [W] int64 first frame of each window inside
 * heat_d (may be NULL when weights->reproj == 0); clip_d: [W] int32 row of mean_bone_d [clips][J];
 * energy_d [W], terms_d [W][5] (E_3d, E_smooth, E_bone, E_vae, E_reproj; may be NULL),
 * status_d [W] uint32 OR-ed with GEM_WIN_* bits (may be NULL). */
int gem_energy_grad(gem_ctx* ctx, void* stream, int W, const float* pose_d, const float* pose0_d,
                    const float* heat_d, const int64_t* frame_base_d, const int32_t* clip_d,
                    const float* mean_bone_d, const gem_energy_weights* weights_h, float* energy_d,
                    float* terms_d, float* grad_d, uint32_t* status_d);

/* ---- the GEMM both VAE contractions run on (exported for tests) ---------------------------- */
/* c[M][N] = act( bias[N] + a[M][K] b[K][N] ), row-major fp32, N % 4 == 0.  use_tensor_cores = 1 takes the
 * tcgen05 3xTF32 kernel (needs K % 32 == 0, N % 128 == 0, M >= 64), 0 the CUDA-core fp32 kernel. */
int gem_gemm(gem_ctx* ctx, void* stream, int M, int N, int K, const float* a_d, int lda, const float* b_d,
             const float* bias_d, int leaky_relu, float* c_d, int ldc, int use_tensor_cores);

/* ---- VAE pieces (ConvVAE.decode_to_bodypose / get_latent_space) -------------------------- */
/* z_d [W][latent] -> pose_d [W][T][J][3]; keeps the activations for gem_decode_vjp */
int gem_decode(gem_ctx* ctx, void* stream, int which, int W, const float* z_d, float* pose_d);
/* dpose_d [W][T][J][3] -> dz_d [W][latent] (bwd-data only; uses the last gem_decode's activations) */
int gem_decode_vjp(gem_ctx* ctx, void* stream, int which, int W, const float* dpose_d, float* dz_d);
/* pose_d [W][T][J*3], eps_d [W][latent] -> z0 = mu + eps*exp(0.5*logvar); mu_d/std_d optional */
int gem_encode(gem_ctx* ctx, void* stream, int which, int W, const float* pose_d, const float* eps_d,
               float* z0_d, float* mu_d, float* std_d);

/* ---- batched L-BFGS (torch.optim.LBFGS.step semantics, one independent solve per window) - */
/* start W solves from z0_d [W][n]; afterwards gem_lbfgs_trial() is the point to evaluate */
int gem_lbfgs_begin(gem_ctx* ctx, void* stream, int W, const float* z0_d, const gem_lbfgs_params* params_h);
/* device pointer [W][n] of the points whose (loss, gradient) the next gem_lbfgs_advance expects */
const float* gem_lbfgs_trial(gem_ctx* ctx);
/* feed loss_d [W] and grad_d [W][n] evaluated at gem_lbfgs_trial(); advances every unfinished
 * window to its next evaluation point (or to termination) without host synchronisation */
int gem_lbfgs_advance(gem_ctx* ctx, void* stream, int W, const float* loss_d, const float* grad_d);
/* current iterate [W][n] (the optimum once finished) */
const float* gem_lbfgs_x(gem_ctx* ctx);
/* copies per-window counters to caller device arrays (any may be NULL): n_iter, func_evals int32,
 * finished int32 (1 when the solve has terminated), final step t (double), loss (double),
 * trace_d [W][trace_stride] fp32 losses in evaluation order */
int gem_lbfgs_stats(gem_ctx* ctx, void* stream, int W, int32_t* n_iter_d, int32_t* func_evals_d,
                    int32_t* finished_d, double* t_d, double* loss_d, float* trace_d, int trace_stride);

/* ---- one full stage: optimize_pose_seq_pytorch_LBFGS for W windows at once --------------- */
/* pose0_d [W][T][J][3] (the stage's input pose = x0 of E_3d), eps_d [W][latent] reparameterisation
 * noise, outputs pose_out_d [W][T][J][3]; energy_trace_d [W][max_eval+1] (optional),
 * n_iter_d / func_evals_d [W] (optional), status_d [W] (optional). */
int gem_solve_stage(gem_ctx* ctx, void* stream, int which, int W, const float* pose0_d, const float* heat_d,
                    const int64_t* frame_base_d, const int32_t* clip_d, const float* mean_bone_d,
                    const float* eps_d, const gem_energy_weights* weights_h, const gem_lbfgs_params* params_h,
                    float* pose_out_d, float* energy_trace_d, int32_t* n_iter_d, int32_t* func_evals_d,
                    uint32_t* status_d);

/* ---- the whole per-window path: local stage -> SLAM transform -> global stage (optimizer.py:386-419) ----
 * Per slice of windows, on the slice's stream: gem_solve_stage(which = 0) with weights_local_h, then
 * gem_relative_global of its result with cams_d [W][T][4][4] float64, then gem_solve_stage(which = 1) with
 * weights_global_h (reproj must be 0) anchored at the transformed pose.  eps_d is [W][2][latent] (local, global).
 * Outputs: local_pose_d, global_pose_d [W][T][J][3] fp32; rel_f32_d (required) / rel_f64_d (optional) the
 * transformed local result; n_iter_d, func_evals_d [2][W] (optional); status_d [W] (optional: GEM_WIN_* bits of both stages OR-ed). */
int gem_solve_windows(gem_ctx* ctx, void* stream, int W, const float* pose0_d, const float* heat_d,
                      const int64_t* frame_base_d, const int32_t* clip_d, const float* mean_bone_d,
                      const double* cams_d, const float* eps_d, const gem_energy_weights* weights_local_h,
                      const gem_energy_weights* weights_global_h, const gem_lbfgs_params* params_h,
                      float* local_pose_d, double* rel_f64_d, float* rel_f32_d, float* global_pose_d,
                      int32_t* n_iter_d, int32_t* func_evals_d, uint32_t* status_d);

/* ---- SLAM camera transforms and stitching (float64 like the reference's numpy) ----------- */
/* get_relative_global_pose_with_camera_matrix (utils/utils.py:99-112): out[w][t] = inv(C[w][0]) C[w][t] x.
 * pose_d fp32 [W][T][J][3] (a stage result) or, when pose_is_f64, float64; cams_d float64 [W][T][4][4];
 * out_f64_d [W][T][J][3] float64 and/or out_f32_d (the fp32 cast the next stage consumes). */
int gem_relative_global(gem_ctx* ctx, void* stream, int W, const void* pose_d, int pose_is_f64,
                        const double* cams_d, double* out_f64_d, float* out_f32_d);
/* relative_global_pose_to_global_pose (optimizer.py:302-308): out[w][t] = C[w][0] x */
int gem_to_global(gem_ctx* ctx, void* stream, int W, const void* pose_d, int pose_is_f64, const double* cams_d,
                  double* out_f64_d);
/* merge_batches (optimizer.py:425-437) of W consecutive windows of one sequence, overlap frames
 * averaged: windows_d float64 [W][T][J][3] -> out_d float64 [(T-overlap)*W+overlap][J][3] */
int gem_merge_windows(gem_ctx* ctx, void* stream, int W, int overlap, const double* windows_d, double* out_d);
/* scipy.ndimage.gaussian_filter1d(seq, sigma, axis=0, mode='reflect') (optimizer.py:450) on
 * seq_d float64 [N][row] -> out_d */
int gem_gaussian_smooth(gem_ctx* ctx, void* stream, int N, int row, double sigma, const double* seq_d,
                        double* out_d);


/* ---- initial 3-D lift (the step that produces the optimiser's `estimated_local_skeleton`) --------
 * Skeleton.set_skeleton for n_frames frames (utils/skeleton.py:33-46): get_max_preds (:176-204) on the maps as
 * set_skeleton_from_file prepares them (:80-82: nearest resize by `up` = 16, `pad_x` = 128 zero columns left and
 * right) followed by FishEyeCameraCalibrated.camera2world (FishEyeCalibrated.py:18-33, float64).  heat_d float32
 * [n_frames][H][W][J] in the pickle's HWC layout (H*W a multiple of 512, J <= 16, 16-byte aligned), depth_d float64
 * [n_frames][J], poly_c2w_h = the calibration's polynomialC2W (constant term first, host memory), (cx, cy) =
 * intrinsic[0][2], intrinsic[1][2].  points_d float64 [n_frames][J][3]; optional outputs (NULL to skip): preds_d
 * float32 [n_frames][J][2] image coordinates, maxvals_d float32 [n_frames][J], argmax_d int32 [n_frames][J] = y0*W+x0
 * of the source map's first row-major maximum.  No ctx: the call only needs the current device. */
int gem_lift_skeleton(void* stream, int n_frames, int H, int W, int J, const float* heat_d, const double* depth_d,
                      const double* poly_c2w_h, int n_poly, double cx, double cy, int up, int pad_x, double* points_d,
                      float* preds_d, float* maxvals_d, int32_t* argmax_d);


/* ---- error metrics ---------------------------------------------------------------------------
 * The per-pose part of calculate_errors (calculate_errors.py:114-179): every estimated pose est_d[f] is aligned to
 * its ground truth gt_d[f] by the similarity transform of umeyama (utils/rigid_transform_with_scale.py:18-43); with
 * bone_len_mm_h != NULL both poses are first rescaled bone by bone along the kinematic chain parents_h
 * (Skeleton._skeleton_resize, utils/skeleton.py:123-135; lengths in millimetres, host memory).  est_d, gt_d float64
 * [n_frames][J][3] (3 <= J <= 32); err_d float64 [n_frames][J] = |aligned - gt|; optional aligned_d / gt_out_d
 * float64 [n_frames][J][3] = the aligned estimate and the (rescaled) ground truth it is compared with.  No ctx. */
int gem_pose_align_errors(void* stream, int n_frames, int J, const double* est_d, const double* gt_d,
                          const int32_t* parents_h, const double* bone_len_mm_h, double* aligned_d, double* gt_out_d,
                          double* err_d);

#ifdef __cplusplus
}
#endif
#endif /* GEM_B200_H */
