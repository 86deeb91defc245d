// C ABI of libgem_b200.so (declared in include/gem_b200.h): context, scratch ownership and the
// per-stage pipeline that strings the kernels together without host synchronisation.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "kernels.cuh"

namespace gem {

// ---- error plumbing ------------------------------------------------------------------------
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_last_error = buf;
    return GEM_ERR_CUDA;
}

}  // namespace gem

using namespace gem;

struct gem_ctx {
    int device = 0;
    int Wmax = 0, n = 0, T = 0, J = 0, H = 0, Wd = 0, m = 0;
    int gemm_mode = 0;
    int heat_planar = 0;                         // heat-map layout: 0 = [frames][H][W][J] (pickle), 1 = [frames][J][H][W], 2 = tiled
    int fuse_energy = 0;                         // opt-in (GEM_FUSE_ENERGY=1): the backward chain on CTA pairs evaluates the energy
                                                 // itself; measured 0.5 ms per step SLOWER than the separate kernel
    int tap_chain = 2;                           // mode 3: 0 one launch per k=3 layer, 1 the four K<=128 layers of each direction
                                                 // in one launch, 2 all five on CTA pairs (cta_group::2) in one launch
    bool have_camera = false, have_skeleton = false;
    CameraConst cam;                             // this ctx's calibration and kinematic tree: kernel parameters of the energy
    SkeletonConst skel;                          // kernels (captured graphs bake them in, so setting either drops the graphs)
    bool have_vae[2] = {false, false};
    gem_vae_weights vae[2];
    int64_t scratch_bytes = 0;
    std::vector<void*> allocs;

    // decoder activations (token-major [W*T][C]) and their gradients
    float *act[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *pose = nullptr;
    float *gact[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *gpose = nullptr;
    // the same activations as TF32 hi / lo pairs: what the tensor-core layers read and write (gemm_mode 1)
    float *act_hi[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *act_lo[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float *gact_hi[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *gact_lo[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    uint32_t* act_sign[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // packed signs of act[i]: the bwd-data masks
    float *gp_hi = nullptr, *gp_lo = nullptr;   // d pose, [W*T][kPosePad]
    // gemm_mode 2: d act0 as row-scaled fp16 hi / lo for the T*256 -> latent GEMM, and the rows' exponents
    uint16_t *g0_h16 = nullptr, *g0_l16 = nullptr;
    int32_t* row_exp = nullptr;
    bool act_split = false;                      // the last decode left its activations in act_hi/act_lo
    bool tap_tc[2] = {false, false};             // the VAE's conv layers are prepared for the tcgen05 tap kernel
    // encoder activations, fc output
    float *eact[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *fc = nullptr, *z0 = nullptr;
    float *e4_hi = nullptr, *e4_lo = nullptr;    // last encoder conv activation split for the tcgen05 fc GEMM
    uint16_t *e2_hi = nullptr, *e2_lo = nullptr, *e3_hi = nullptr, *e3_lo = nullptr;   // fp16 pairs: enc[2], enc[3] outputs (mode 3)
    bool enc_tc[2] = {false, false};             // enc[3], enc[4] are prepared for the tcgen05 tap kernel
    int enc_tc_on = 1;                           // encoder's wide k=3 layers on tcgen05 (GEM_ENC_TC=0: CUDA cores): see encode_impl
    // closure outputs
    float *f_new = nullptr, *g_new = nullptr;
    LbfgsBuffers lb;
    gem_lbfgs_params lb_params;
    bool lb_started = false;
    // chunked execution of a stage (gem_ctx_set_chunks): internal streams and fork/join events
    int n_chunks = 1;
    std::vector<cudaStream_t> streams;
    std::vector<cudaEvent_t> join_ev;
    cudaEvent_t fork_ev = nullptr;
    // graph-stable copies of a stage's per-window inputs / outputs (so that a captured round can be replayed
    // by later calls whatever buffers the caller passes)
    float *pose0_own = nullptr, *mb_own = nullptr, *trace_own = nullptr;
    int64_t* fb_own = nullptr;
    int32_t* clip_own = nullptr;
    uint32_t* status_own = nullptr;
    // energy kernel's texel cache (heat maps read from pinned host memory): [W][T*J][kPatchFloats] + origins + valid
    // masks, allocated by the first call that needs them (ensure_texel_cache); and counters
    float* patch = nullptr;
    short2* patch_origin = nullptr;
    unsigned long long* patch_valid = nullptr;   // one bit per texel of a joint's window
    unsigned long long* patch_stats = nullptr;   // {lookups, texels fetched}, counted while patch_stats_on
    uint2* miss_list = nullptr;                  // planar maps: joints whose footprint the probe kernel found missing [W][T*J]
    uint32_t* miss_count = nullptr;              // [W][2]: {entries, fetch CTAs done} of the slice that starts at the window
    bool patch_stats_on = false;
    int texel_cache = -1;                        // -1 auto (on when the maps are host memory), 0 off, 1 on
    int texel_keep = 0;                          // measurement only (GEM_TEXEL_KEEP=1): stages do not empty the texel windows
    int texel_prefetch_ctas = 8;                 // CTAs of the texel fetch / prefetch kernel (0: the energy kernel fetches itself)
    int texel_prefetch_threads = 512;            // threads per CTA of the texel prefetch kernel
    int texel_probe = 1;                         // planar maps: probe + fetch launches instead of the one prefetch kernel (GEM_TEXEL_PROBE=0)
    // a stage's FIRST round misses on every joint: its fetches are issued slice after slice (each slice's fetch waits
    // for the previous slice's), with more CTAs and only the footprint's rows, so that the first slices compute while
    // the later ones still wait for the bus - instead of all slices sharing the bus and finishing together
    int texel_cold_chain = 1, texel_cold_ctas = 32, texel_cold_rows = 2, texel_rows = 4;
    std::vector<cudaEvent_t> cold_ev;            // one per slice of the running call
    int cold_used = 0;                           // events of cold_ev recorded by the running call
    int trace_cap = 0;                           // columns of trace_own
    bool use_graphs = true;
    struct RoundGraph {
        int which, w0, Wk, gemm_mode, has_heat, trace_stride, launches;
        const void* heat;
        gem_energy_weights wt;
        gem_lbfgs_params p;
        cudaGraphExec_t exec;
    };
    std::vector<RoundGraph> graphs;
    std::vector<int> user_slices;                // first window of each slice (gem_ctx_set_slices); empty = automatic
    struct Ready {
        int first_window;
        cudaEvent_t ev;
    };
    std::vector<Ready> ready;                    // one-shot: inputs of windows >= first_window are complete after ev
    void* tc_workspace = nullptr;
    size_t tc_workspace_bytes = 0;
    // instrumentation: kernel-launch counter and optional CUDA-event pairs around every launch
    int64_t launches = 0;
    bool prof_on = false;
    struct ProfEntry {
        int tag;
        cudaEvent_t a, b;
    };
    std::vector<ProfEntry> prof;
    size_t prof_used = 0;
};

// Runs one kernel launch `f` under the ctx's instrumentation (tag identifies the kernel class).
template <typename F>
static int timed(gem_ctx* c, cudaStream_t s, int tag, F&& f) {
    c->launches += 1;
    if (!c->prof_on) return f();
    if (c->prof_used == c->prof.size()) {
        gem_ctx::ProfEntry e;
        e.tag = tag;
        GEM_CUDA(cudaEventCreate(&e.a));
        GEM_CUDA(cudaEventCreate(&e.b));
        c->prof.push_back(e);
    }
    gem_ctx::ProfEntry& e = c->prof[c->prof_used++];
    e.tag = tag;
    GEM_CUDA(cudaEventRecord(e.a, s));
    const int rc = f();
    GEM_CUDA(cudaEventRecord(e.b, s));
    return rc;
}

// Captured rounds bake in kernel parameters (weight pointers, TMA descriptors, camera, skeleton): whenever one of
// those changes the cached graphs are dropped and re-captured by the next solve.
static void drop_graphs(gem_ctx* c) {
    for (auto& g : c->graphs) cudaGraphExecDestroy(g.exec);
    c->graphs.clear();
}

static const int kDecC[6] = {256, 128, 64, 64, 64, 0};   // channels after dec[0..4]; dec[5] -> J*3
// row pitch (elements) of the split d pose: 16-byte multiples for TMA; a whole 64-element K block with fp16 elements
static int pose_pad(const gem_ctx* c) { return c->gemm_mode == 3 ? (c->J * 3 + 63) & ~63 : (c->J * 3 + 3) & ~3; }
static const int kEncC[5] = {64, 64, 128, 256, 512};

static int ensure_streams(gem_ctx* c, int n) {
    if (n <= 0) return GEM_OK;
    if (!c->fork_ev) GEM_CUDA(cudaEventCreateWithFlags(&c->fork_ev, cudaEventDisableTiming));
    while ((int)c->streams.size() < n) {
        cudaStream_t st;
        cudaEvent_t ev;
        GEM_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        GEM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c->streams.push_back(st), c->join_ev.push_back(ev);
    }
    return GEM_OK;
}

template <typename T>
static int ctx_alloc(gem_ctx* c, T** p, size_t count) {
    void* q = nullptr;
    const size_t bytes = count * sizeof(T);
    GEM_CUDA(cudaMalloc(&q, bytes ? bytes : 16));
    c->allocs.push_back(q);
    c->scratch_bytes += (int64_t)bytes;
    *p = static_cast<T*>(q);
    return GEM_OK;
}

// The texel cache's buffers (1 KB per joint-frame with 16 x 16 windows) exist only in ctxs that read maps through it.
static int ensure_texel_cache(gem_ctx* c) {
    if (c->patch) return GEM_OK;
    const size_t Weven = ((size_t)c->Wmax + 1) & ~(size_t)1, joints = Weven * c->T * c->J;
    float* patch = nullptr;
    int rc = ctx_alloc(c, &patch, joints * kPatchFloats);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->patch_origin, joints);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->patch_valid, joints);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->miss_list, joints);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->miss_count, 2 * Weven);
    if (rc != GEM_OK) return rc;
    GEM_CUDA(cudaMemset(c->patch_valid, 0, joints * sizeof(unsigned long long)));
    GEM_CUDA(cudaMemset(c->patch_origin, 0, joints * sizeof(short2)));
    GEM_CUDA(cudaMemset(c->miss_count, 0, 2 * Weven * sizeof(uint32_t)));
    c->patch = patch;
    return GEM_OK;
}

#define GEM_TRY(expr)            \
    do {                         \
        int _rc = (expr);        \
        if (_rc != GEM_OK) return _rc; \
    } while (0)

extern "C" {

int gem_version(void) { return 100; }
const char* gem_last_error(void) { return g_last_error.c_str(); }

int gem_ctx_create(gem_ctx** out, int device, int max_windows, int latent_dim, int seq_len, int num_joints, int heat_h,
                   int heat_w, int max_history) {
    GEM_REQUIRE(out != nullptr, "out is NULL");
    GEM_REQUIRE(max_windows > 0 && latent_dim > 0 && latent_dim % 4 == 0 && latent_dim <= 2048,
                "latent_dim must be a multiple of 4 in (0, 2048]");
    GEM_REQUIRE(seq_len >= 3 && seq_len <= 32 && num_joints > 0 && num_joints <= kMaxJoints &&
                    seq_len * num_joints <= 160,
                "window geometry out of range");
    GEM_REQUIRE(max_history >= 1 && max_history <= 64, "max_history must be in [1, 64]");
    GEM_CUDA(cudaSetDevice(device));
    gem_ctx* c = new gem_ctx();
    c->device = device, c->Wmax = max_windows, c->n = latent_dim, c->T = seq_len, c->J = num_joints;
    c->H = heat_h, c->Wd = heat_w, c->m = max_history;
    const size_t W = (size_t)max_windows, tok = W * seq_len, n = latent_dim;
    const size_t Weven = (W + 1) & ~(size_t)1;   // energy kernel copies window pairs
    int rc = GEM_OK;
    auto A = [&](float** p, size_t cnt) { if (rc == GEM_OK) rc = ctx_alloc(c, p, cnt); };
    for (int i = 0; i < 5; ++i) A(&c->act[i], tok * kDecC[i]), A(&c->gact[i], tok * kDecC[i]);
    for (int i = 0; i < 5; ++i) {
        A(&c->act_hi[i], tok * kDecC[i]), A(&c->act_lo[i], tok * kDecC[i]);
        A(&c->gact_hi[i], tok * kDecC[i]), A(&c->gact_lo[i], tok * kDecC[i]);
        if (rc == GEM_OK) rc = ctx_alloc(c, &c->act_sign[i], tok * kDecC[i] / 32);
    }
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->g0_h16, tok * kDecC[0]);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->g0_l16, tok * kDecC[0]);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->row_exp, W);
    {
        const size_t cnt = tok * (size_t)((num_joints * 3 + 3) & ~3);
        A(&c->gp_hi, cnt), A(&c->gp_lo, cnt);
        if (rc == GEM_OK && cudaMemset(c->gp_hi, 0, cnt * sizeof(float)) != cudaSuccess) rc = GEM_ERR_CUDA;
        if (rc == GEM_OK && cudaMemset(c->gp_lo, 0, cnt * sizeof(float)) != cudaSuccess) rc = GEM_ERR_CUDA;
    }
    A(&c->pose, Weven * seq_len * num_joints * 3), A(&c->gpose, Weven * seq_len * num_joints * 3);
    for (int i = 0; i < 5; ++i) A(&c->eact[i], tok * kEncC[i]);
    A(&c->e4_hi, tok * kEncC[4]), A(&c->e4_lo, tok * kEncC[4]);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->e2_hi, tok * kEncC[2]);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->e2_lo, tok * kEncC[2]);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->e3_hi, tok * kEncC[3]);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->e3_lo, tok * kEncC[3]);
    A(&c->fc, W * 2 * n), A(&c->z0, W * n), A(&c->f_new, W), A(&c->g_new, W * n);
    A(&c->pose0_own, Weven * seq_len * num_joints * 3), A(&c->mb_own, W * num_joints);
    c->trace_cap = max_history + 8;              // max_eval + 1 <= (max_history + 1) * 5 / 4 + 1
    A(&c->trace_own, W * (size_t)c->trace_cap);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->fb_own, W);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->clip_own, W);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->status_own, W);
    if (rc == GEM_OK) rc = ctx_alloc(c, &c->patch_stats, 2);
    if (rc == GEM_OK && cudaMemset(c->patch_stats, 0, 2 * sizeof(unsigned long long)) != cudaSuccess) rc = GEM_ERR_CUDA;
    LbfgsBuffers& b = c->lb;
    memset(&b, 0, sizeof(b));
    A(&b.X, W * n), A(&b.D, W * n), A(&b.G, W * n), A(&b.GP, W * n), A(&b.BG0, W * n),
        A(&b.BG1, W * n), A(&b.ZT, W * n), A(&b.ZT_hi, W * n), A(&b.ZT_lo, W * n);
    A(&b.Y, W * max_history * n), A(&b.S, W * max_history * n), A(&b.RO, W * max_history);
    if (rc == GEM_OK) {
        void* st = nullptr;
        cudaError_t e = cudaMalloc(&st, W * lbfgs_state_bytes());
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(lbfgs state)", __FILE__, __LINE__);
        else c->allocs.push_back(st), c->scratch_bytes += (int64_t)(W * lbfgs_state_bytes()), b.st = (LbfgsWin*)st;
    }
    b.n = latent_dim, b.m = max_history;
    if (rc != GEM_OK) {
        gem_ctx_destroy(c);
        return rc;
    }
    // default: tcgen05 everywhere in the fp16 scheme (3); GEM_GEMM_MODE / gem_ctx_set_gemm_mode select 2 (fp16-scheme
    // GEMMs, 3xTF32 convolutions), 1 (3xTF32 everywhere) or 0 (fp32 CUDA cores)
    c->gemm_mode = 3;
    if (const char* env = getenv("GEM_GEMM_MODE")) c->gemm_mode = (env[0] >= '0' && env[0] <= '3') ? env[0] - '0' : 3;
    if (const char* env = getenv("GEM_TAP_CHAIN")) c->tap_chain = (env[0] >= '0' && env[0] <= '2') ? env[0] - '0' : 2;
    if (const char* env = getenv("GEM_ENC_TC")) c->enc_tc_on = env[0] != '0';
    if (const char* env = getenv("GEM_FUSE_ENERGY")) c->fuse_energy = env[0] != '0';
    c->n_chunks = 0;      // automatic (slice_bounds)
    if (const char* env = getenv("GEM_CHUNKS")) c->n_chunks = atoi(env) >= 0 ? atoi(env) : 0;
    if (c->n_chunks > 16) c->n_chunks = 16;
    if (const char* env = getenv("GEM_GRAPHS")) c->use_graphs = env[0] != '0';
    if (const char* env = getenv("GEM_TEXEL_CACHE")) c->texel_cache = atoi(env);
    if (const char* env = getenv("GEM_TEXEL_KEEP")) c->texel_keep = atoi(env) != 0;
    if (const char* env = getenv("GEM_TEXEL_PREFETCH_CTAS")) c->texel_prefetch_ctas = atoi(env) > 0 ? atoi(env) : 0;
    if (const char* env = getenv("GEM_TEXEL_PREFETCH_THREADS")) c->texel_prefetch_threads = atoi(env);
    if (const char* env = getenv("GEM_TEXEL_PROBE")) c->texel_probe = env[0] != '0';
    if (const char* env = getenv("GEM_TEXEL_COLD_CHAIN")) c->texel_cold_chain = env[0] != '0';
    if (const char* env = getenv("GEM_TEXEL_COLD_CTAS")) c->texel_cold_ctas = atoi(env) > 0 ? atoi(env) : 1;
    if (const char* env = getenv("GEM_TEXEL_COLD_ROWS")) c->texel_cold_rows = atoi(env);
    if (const char* env = getenv("GEM_TEXEL_ROWS")) c->texel_rows = atoi(env);
    // default skeleton: the reference's 15-joint kinematic tree (optimizer.py:34)
    if (num_joints == 15) {
        static const int32_t parents[15] = {0, 0, 1, 2, 0, 4, 5, 1, 7, 8, 9, 4, 11, 12, 13};
        rc = gem_ctx_set_skeleton(c, parents, 15);
        if (rc != GEM_OK) {
            gem_ctx_destroy(c);
            return rc;
        }
    }
    *out = c;
    return GEM_OK;
}

int gem_ctx_destroy(gem_ctx* c) {
    if (!c) return GEM_OK;
    cudaSetDevice(c->device);
    for (void* p : c->allocs) cudaFree(p);
    drop_graphs(c);
    for (cudaStream_t st : c->streams) cudaStreamDestroy(st);
    for (cudaEvent_t ev : c->join_ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : c->cold_ev) cudaEventDestroy(ev);
    if (c->fork_ev) cudaEventDestroy(c->fork_ev);
    tc_gemm_release(c);
    tc_tap_release(c);
    for (auto& e : c->prof) cudaEventDestroy(e.a), cudaEventDestroy(e.b);
    delete c;
    return GEM_OK;
}

int64_t gem_ctx_scratch_bytes(const gem_ctx* c) { return c ? c->scratch_bytes : 0; }

int gem_ctx_set_gemm_mode(gem_ctx* c, int mode) {
    GEM_REQUIRE(c != nullptr, "ctx is NULL");
    GEM_REQUIRE(mode >= 0 && mode <= 3, "mode must be 0, 1, 2 or 3");
    if (mode >= 1 && !tc_gemm_available()) {
        set_error("tcgen05 GEMM path not available in this build");
        return GEM_ERR_STATE;
    }
    c->gemm_mode = mode;
    return GEM_OK;
}

/* debug hook (not in the public header): per-CTA phase timestamps of the tcgen05 tap kernel */
int gem_debug_tap_timestamps(long long* buf_d) {
    gem::g_tap_dbg = buf_d;
    return GEM_OK;
}

/* debug hook (not in the public header): 1 = CTA-pair (cta_group::2) fp16-scheme GEMM, 0 = one CTA per tile, -1 = default */
/* debug hook: 0 = one launch per k=3 layer, 1 = one-CTA chains of four layers, 2 = CTA-pair chains of five (default) */
int gem_debug_tap_chain(gem_ctx* c, int mode) {
    GEM_REQUIRE(c != nullptr && mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
    c->tap_chain = mode;
    return GEM_OK;
}
/* debug hook: 1 = the encoder's 128 -> 256 -> 512 layers on the tcgen05 tap kernel (mode 3), 0 = CUDA cores (default) */
/* debug hook: 1 = the backward chain evaluates the energy in its prologue (default), 0 = separate energy kernel */
int gem_debug_fuse_energy(gem_ctx* c, int on) {
    GEM_REQUIRE(c != nullptr, "ctx is NULL");
    c->fuse_energy = on ? 1 : 0;
    return GEM_OK;
}
int gem_debug_enc_tc(gem_ctx* c, int on) {
    GEM_REQUIRE(c != nullptr, "ctx is NULL");
    c->enc_tc_on = on ? 1 : 0;
    return GEM_OK;
}
int gem_debug_gemm_timestamps(long long* buf_d) {
    gem::g_gemm_dbg = buf_d;
    return GEM_OK;
}
int gem_debug_gemm_pair(int mode) {
    gem::g_gemm_pair = mode;
    return GEM_OK;
}
int gem_debug_energy_fixed(int mode) {      // 0: the runtime-shape energy kernel for every launch, 1: compile-time shape where it applies, -1: default
    gem::g_energy_fixed = mode;
    return GEM_OK;
}

int gem_ctx_set_chunks(gem_ctx* c, int n_chunks) {
    GEM_REQUIRE(c != nullptr && n_chunks >= 0 && n_chunks <= 16, "n_chunks must be in [0, 16] (0 = automatic)");
    c->n_chunks = n_chunks;
    return GEM_OK;
}

int gem_ctx_set_heat_layout(gem_ctx* c, int planar) {
    GEM_REQUIRE(c != nullptr && planar >= 0 && planar <= 2, "layout must be 0 (HWC), 1 (planar CHW) or 2 (tiled CHW)");
    GEM_REQUIRE(!planar || c->Wd % 4 == 0, "planar heat maps need a width that is a multiple of 4");
    GEM_REQUIRE(planar != 2 || (c->Wd % kTileW == 0 && c->H % kTileH == 0), "tiled heat maps need H % 4 == 0 and W % 8 == 0");
    c->heat_planar = planar;
    return GEM_OK;
}

int gem_ctx_set_texel_cache(gem_ctx* c, int mode) {
    GEM_REQUIRE(c != nullptr && mode >= -1 && mode <= 1, "mode must be -1 (auto), 0 or 1");
    c->texel_cache = mode;
    return GEM_OK;
}

int gem_ctx_texel_cache_stats(gem_ctx* c, int enable, uint64_t* lookups_h, uint64_t* texels_fetched_h) {
    GEM_REQUIRE(c != nullptr, "ctx is NULL");
    GEM_CUDA(cudaSetDevice(c->device));
    GEM_CUDA(cudaDeviceSynchronize());
    unsigned long long v[2] = {0, 0};
    GEM_CUDA(cudaMemcpy(v, c->patch_stats, sizeof(v), cudaMemcpyDeviceToHost));
    if (lookups_h) *lookups_h = v[0];
    if (texels_fetched_h) *texels_fetched_h = v[1];
    GEM_CUDA(cudaMemset(c->patch_stats, 0, sizeof(v)));
    c->patch_stats_on = enable != 0;
    return GEM_OK;
}

int gem_ctx_set_profiling(gem_ctx* c, int enable) {
    GEM_REQUIRE(c != nullptr, "ctx is NULL");
    c->prof_on = enable != 0;
    c->prof_used = 0;
    return GEM_OK;
}

int64_t gem_ctx_launch_count(const gem_ctx* c) { return c ? c->launches : 0; }

int gem_ctx_read_profile(gem_ctx* c, int max_tags, int32_t* tags_h, int32_t* counts_h, float* total_ms_h,
                         int32_t* n_tags_h) {
    GEM_REQUIRE(c && tags_h && counts_h && total_ms_h && n_tags_h && max_tags > 0, "bad arguments");
    GEM_CUDA(cudaSetDevice(c->device));
    GEM_CUDA(cudaDeviceSynchronize());
    int n = 0;
    for (size_t i = 0; i < c->prof_used; ++i) {
        float ms = 0.f;
        GEM_CUDA(cudaEventElapsedTime(&ms, c->prof[i].a, c->prof[i].b));
        int k = 0;
        while (k < n && tags_h[k] != c->prof[i].tag) ++k;
        if (k == n) {
            if (n == max_tags) continue;
            tags_h[n] = c->prof[i].tag, counts_h[n] = 0, total_ms_h[n] = 0.f;
            ++n;
        }
        counts_h[k] += 1;
        total_ms_h[k] += ms;
    }
    *n_tags_h = n;
    c->prof_used = 0;
    return GEM_OK;
}

int gem_ctx_set_camera(gem_ctx* c, const double* poly_h, int n_poly, double cx, double cy) {
    GEM_REQUIRE(c && poly_h && n_poly >= 1 && n_poly <= kMaxPoly, "bad camera polynomial");
    CameraConst cam;
    memset(&cam, 0, sizeof(cam));
    for (int i = 0; i < n_poly; ++i) cam.poly[i] = (float)poly_h[i];   // python scalars are cast to the tensor dtype
    cam.n_poly = n_poly, cam.cx = (float)cx, cam.cy = (float)cy;
    GEM_CUDA(cudaSetDevice(c->device));
    if (!c->have_camera || memcmp(&c->cam, &cam, sizeof(cam)) != 0) {
        GEM_CUDA(cudaDeviceSynchronize());       // (replayed graphs of an earlier solve may still be running)
        drop_graphs(c);
        c->cam = cam;
    }
    c->have_camera = true;
    return GEM_OK;
}

int gem_ctx_set_skeleton(gem_ctx* c, const int32_t* parents_h, int num_joints) {
    GEM_REQUIRE(c && parents_h && num_joints == c->J, "skeleton size must match the ctx");
    SkeletonConst sk;
    memset(&sk, 0, sizeof(sk));
    for (int j = 0; j < num_joints; ++j) {
        GEM_REQUIRE(parents_h[j] >= 0 && parents_h[j] < num_joints, "parent index out of range");
        sk.parent[j] = parents_h[j];
    }
    sk.num_joints = num_joints;
    int nc = 0;
    for (int j = 0; j < num_joints; ++j) {
        sk.child_start[j] = nc;
        for (int cj = 0; cj < num_joints; ++cj)
            if (cj != j && parents_h[cj] == j) sk.child_list[nc++] = cj;
    }
    for (int j = num_joints; j <= kMaxJoints; ++j) sk.child_start[j] = nc;
    GEM_CUDA(cudaSetDevice(c->device));
    if (!c->have_skeleton || memcmp(&c->skel, &sk, sizeof(sk)) != 0) {
        GEM_CUDA(cudaDeviceSynchronize());
        drop_graphs(c);
        c->skel = sk;
    }
    c->have_skeleton = true;
    return GEM_OK;
}

int gem_ctx_set_vae(gem_ctx* c, int which, const gem_vae_weights* w) {
    GEM_REQUIRE(c && w && (which == 0 || which == 1), "bad arguments");
    const int T = c->T, n = c->n, P = c->J * 3;
    const int dk[6] = {n, 256, 128, 64, 64, 64}, dn[6] = {T * 256, 128, 64, 64, 64, P}, dt[6] = {1, 3, 3, 3, 3, 3};
    const int bk[6] = {P, 64, 64, 64, 128, T * 256}, bn[6] = {64, 64, 64, 128, 256, n}, bt[6] = {3, 3, 3, 3, 3, 1};
    const int ek[6] = {P, 64, 64, 128, 256, T * 512}, en_[6] = {64, 64, 128, 256, 512, 2 * n},
              et[6] = {3, 3, 3, 3, 3, 1};
    for (int i = 0; i < 6; ++i) {
        GEM_REQUIRE(w->dec[i].w_d && w->dec[i].k == dk[i] && w->dec[i].n == dn[i] && w->dec[i].taps == dt[i],
                    "decoder layer shape mismatch");
        GEM_REQUIRE(w->dec_bwd[i].w_d && w->dec_bwd[i].k == bk[i] && w->dec_bwd[i].n == bn[i] &&
                        w->dec_bwd[i].taps == bt[i],
                    "decoder bwd layer shape mismatch");
        GEM_REQUIRE(w->enc[i].w_d && w->enc[i].k == ek[i] && w->enc[i].n == en_[i] && w->enc[i].taps == et[i],
                    "encoder layer shape mismatch");
    }
    GEM_CUDA(cudaSetDevice(c->device));
    // New weights: every captured round still points at the old matrices and their prepared slabs — wait
    // for work in flight, drop the graphs, and free the prepared copies of matrices no slot uses any more.
    GEM_CUDA(cudaDeviceSynchronize());
    drop_graphs(c);
    if (c->have_vae[which]) {
        const gem_vae_weights old = c->vae[which];
        auto still_used = [&](const float* p) {
            for (int i = 0; i < 6; ++i) {
                if (w->dec[i].w_d == p || w->dec_bwd[i].w_d == p || w->enc[i].w_d == p) return true;
                if (c->have_vae[1 - which]) {
                    const gem_vae_weights& o = c->vae[1 - which];
                    if (o.dec[i].w_d == p || o.dec_bwd[i].w_d == p || o.enc[i].w_d == p) return true;
                }
            }
            return false;
        };
        for (int i = 0; i < 6; ++i)
            for (const gem_layer* L : {&old.dec[i], &old.dec_bwd[i], &old.enc[i]})
                if (!still_used(L->w_d)) tc_gemm_forget_weight(c, L->w_d), tc_tap_forget_weight(c, L->w_d);
    }
    c->vae[which] = *w;
    c->have_vae[which] = true;
    // K-major TF32 hi/lo copies of the three plain-GEMM weight matrices for the tcgen05 path
    const gem_layer* big[3] = {&w->dec[0], &w->dec_bwd[5], &w->enc[5]};
    for (const gem_layer* L : big)
        if (L->k % 32 == 0 && L->n % 128 == 0)
        {
            GEM_TRY(tc_gemm_prepare_weight(c, 0, L->w_d, (L->n + 3) & ~3, L->k, L->n, 1));
            if (L->k % 64 == 0) GEM_TRY(tc_gemm_prepare_weight(c, 0, L->w_d, (L->n + 3) & ~3, L->k, L->n, 2));
        }
    // K-major, tap-concatenated hi/lo slabs of the decoder's k=3 convolutions and their bwd-data
    bool tap_ok = c->T <= 32;
    for (int i = 1; i <= 5 && tap_ok; ++i) tap_ok = tc_tap_supported(w->dec[i].k, w->dec[i].n, c->T);
    for (int i = 0; i <= 4 && tap_ok; ++i) tap_ok = tc_tap_supported(w->dec_bwd[i].k, w->dec_bwd[i].n, c->T);
    tap_ok = tap_ok && c->vae[which].dec[0].k % 32 == 0 && c->vae[which].dec[0].n % 128 == 0;
    if (tap_ok) {
        for (int scheme = 1; scheme <= 2; ++scheme) {
            for (int i = 1; i <= 5; ++i)
                GEM_TRY(tc_tap_prepare_weight(c, 0, w->dec[i].w_d, (w->dec[i].n + 3) & ~3, w->dec[i].k, w->dec[i].n, scheme));
            for (int i = 0; i <= 4; ++i)
                GEM_TRY(tc_tap_prepare_weight(c, 0, w->dec_bwd[i].w_d, (w->dec_bwd[i].n + 3) & ~3, w->dec_bwd[i].k,
                                              w->dec_bwd[i].n, scheme));
        }
    }
    c->tap_tc[which] = tap_ok;
    // the encoder's two wide k=3 layers (128 -> 256 -> 512) in the fp16 scheme
    c->enc_tc[which] = false;
    if (tap_ok && tc_tap_supported(w->enc[3].k, w->enc[3].n, c->T) && tc_tap_supported(w->enc[4].k, w->enc[4].n, c->T) &&
        w->enc[3].k % 64 == 0 && w->enc[3].n % 64 == 0 && w->enc[4].n % 64 == 0) {
        for (int i = 3; i <= 4; ++i)
            GEM_TRY(tc_tap_prepare_weight(c, 0, w->enc[i].w_d, (w->enc[i].n + 3) & ~3, w->enc[i].k, w->enc[i].n, 2));
        c->enc_tc[which] = true;
    }
    GEM_CUDA(cudaStreamSynchronize(0));
    return GEM_OK;
}

}  // extern "C"

// ---- layer runner ----------------------------------------------------------------------------
static int run_layer(gem_ctx* c, cudaStream_t s, int tag, const gem_layer& L, const float* A, int lda, int M, float* C,
                     int ldc, int epi, const float* aux, const void* A_hi = nullptr, const void* A_lo = nullptr,
                     float* C_lo = nullptr, uint32_t* C_sign = nullptr, const int32_t* row_exp = nullptr, int out16 = 0,
                     const uint32_t* aux_bits = nullptr, uint32_t* row_flag = nullptr) {
    TapGemmArgs g;
    g.A = A, g.B = L.w_d, g.bias = L.bias_d, g.aux = aux, g.C = C, g.aux_bits = aux_bits;
    g.A_hi = A_hi, g.A_lo = A_lo, g.C_lo = C_lo, g.C_sign = C_sign, g.row_exp = row_exp, g.out16 = out16;
    g.row_flag = row_flag;
    g.M = M, g.N = L.n, g.K = L.k, g.taps = L.taps, g.T = c->T;
    g.lda = lda, g.ldb = (L.n + 3) & ~3, g.ldc = ldc, g.ldaux = L.n, g.epi = epi;
    return timed(c, s, tag, [&]() {
        if (c->gemm_mode >= 1 && L.taps == 1 && (M >= 64 || A_hi || C_lo) && L.k % (c->gemm_mode >= 2 ? 64 : 32) == 0 &&
            L.n % 128 == 0)
            return launch_tap_gemm_tc(s, g, c, (size_t)(c->gemm_mode >= 2 ? 2 : 1));
        return launch_tap_gemm_simt(s, g);
    });
}

// one k=3 convolution on the tcgen05 tap kernel; activations are TF32 hi / lo pairs (gemm_mode 1, 2) or fp16 hi /
// scaled fp16 lo pairs held in the same buffers (gemm_mode 3)
static int run_tap_tc(gem_ctx* c, cudaStream_t s, int tag, const gem_layer& L, const float* A_hi, const float* A_lo, int lda,
                      int W, float* out_hi, float* out_lo, int ldo, int epi, const float* aux,
                      const uint32_t* aux_bits = nullptr, uint32_t* sign_out = nullptr) {
    TapTcLaunch t;
    t.scheme = c->gemm_mode == 3 ? 2 : 1;
    t.aux_bits = aux_bits, t.sign_out = sign_out;
    t.B = L.w_d, t.A_hi = A_hi, t.A_lo = A_lo, t.lda = lda, t.Kreal = lda;
    t.bias = L.bias_d, t.aux = aux, t.ldaux = L.n, t.out_hi = out_hi, t.out_lo = out_lo, t.ldo = ldo;
    t.W = W, t.T = c->T, t.epi = epi;
    return timed(c, s, tag, [&]() { return launch_tap_tc(s, c, t); });
}

static bool use_tc_chain(const gem_ctx* c, int which, int W) { return c->gemm_mode >= 1 && c->tap_tc[which] && W >= 1; }
// the backward chain on CTA pairs can evaluate the energy terms in its prologue (gemm_tap_tc.cu: Chain2Energy)
static bool fused_energy_ok(const gem_ctx* c, int which, int W) {
    return c->fuse_energy && use_tc_chain(c, which, W) && c->gemm_mode == 3 && c->tap_chain == 2 &&
           tap_chain_energy_supported(c->T, c->J, pose_pad(c));
}

// Scratch of the windows [w0, w0 + W): every per-window buffer of the ctx shifted by w0.  Chunks of one
// stage run concurrently on different streams, each on its own slice.
struct Slice {
    int w0;
    size_t tok0;                                  // first token
    float *act[5], *gact[5], *act_hi[5], *act_lo[5], *gact_hi[5], *gact_lo[5];
    uint32_t* act_sign[5];
    float *pose, *gpose, *gp_hi, *gp_lo, *f_new, *g_new;
    uint16_t *g0_h16, *g0_l16;
    int32_t* row_exp;
    float *eact[5], *e4_hi, *e4_lo, *fc, *z0;    // encoder activations, fc output, initial latent
    uint16_t *e2_hi, *e2_lo, *e3_hi, *e3_lo;
    float *pose0_own, *mb_own, *trace_own;       // staged inputs / outputs of this slice
    int64_t* fb_own;
    int32_t* clip_own;
    uint32_t* status_own;
    float* patch;
    short2* patch_origin;
    unsigned long long* patch_valid;
    uint2* miss_list;
    uint32_t* miss_count;
    LbfgsBuffers lb;
};
static Slice slice_of(gem_ctx* c, int w0) {
    Slice v;
    v.w0 = w0, v.tok0 = (size_t)w0 * c->T;
    for (int i = 0; i < 5; ++i) {
        const size_t o = v.tok0 * kDecC[i];
        v.act[i] = c->act[i] + o, v.gact[i] = c->gact[i] + o;
        v.act_hi[i] = c->act_hi[i] + o, v.act_lo[i] = c->act_lo[i] + o;
        v.gact_hi[i] = c->gact_hi[i] + o, v.gact_lo[i] = c->gact_lo[i] + o;
        v.act_sign[i] = c->act_sign[i] + o / 32;
    }
    const size_t P = (size_t)c->J * 3, n = c->n, m = c->m;
    v.pose = c->pose + v.tok0 * P, v.gpose = c->gpose + v.tok0 * P;
    const size_t gp_stride = (size_t)((c->J * 3 + 3) & ~3);      // floats per token the buffers were allocated with
    v.gp_hi = c->gp_hi + v.tok0 * gp_stride, v.gp_lo = c->gp_lo + v.tok0 * gp_stride;   // (64 fp16 fit in 48 floats)
    v.f_new = c->f_new + w0, v.g_new = c->g_new + (size_t)w0 * n;
    v.g0_h16 = c->g0_h16 + v.tok0 * kDecC[0], v.g0_l16 = c->g0_l16 + v.tok0 * kDecC[0], v.row_exp = c->row_exp + w0;
    for (int i = 0; i < 5; ++i) v.eact[i] = c->eact[i] + v.tok0 * kEncC[i];
    v.e4_hi = c->e4_hi + v.tok0 * kEncC[4], v.e4_lo = c->e4_lo + v.tok0 * kEncC[4];
    v.e2_hi = c->e2_hi + v.tok0 * kEncC[2], v.e2_lo = c->e2_lo + v.tok0 * kEncC[2];
    v.e3_hi = c->e3_hi + v.tok0 * kEncC[3], v.e3_lo = c->e3_lo + v.tok0 * kEncC[3];
    v.fc = c->fc + (size_t)w0 * 2 * n, v.z0 = c->z0 + (size_t)w0 * n;
    v.pose0_own = c->pose0_own + v.tok0 * P, v.mb_own = c->mb_own;     // (mean bones are indexed by absolute window)
    v.trace_own = c->trace_own + (size_t)w0 * c->trace_cap;
    v.fb_own = c->fb_own + w0, v.clip_own = c->clip_own + w0, v.status_own = c->status_own + w0;
    v.patch = nullptr, v.patch_origin = nullptr, v.patch_valid = nullptr, v.miss_list = nullptr, v.miss_count = nullptr;
    if (c->patch) {                                   // (ensure_texel_cache)
        v.patch = c->patch + v.tok0 * c->J * kPatchFloats, v.patch_origin = c->patch_origin + v.tok0 * c->J;
        v.patch_valid = c->patch_valid + v.tok0 * c->J;
        v.miss_list = c->miss_list + v.tok0 * c->J, v.miss_count = c->miss_count + 2 * (size_t)w0;
    }
    v.lb = c->lb;
    LbfgsBuffers& b = v.lb;
    const size_t on = (size_t)w0 * n;
    b.X += on, b.D += on, b.G += on, b.GP += on, b.BG0 += on, b.BG1 += on, b.ZT += on, b.ZT_hi += on, b.ZT_lo += on;
    b.Y += on * m, b.S += on * m, b.RO += (size_t)w0 * m;
    b.st = reinterpret_cast<LbfgsWin*>(reinterpret_cast<char*>(b.st) + (size_t)w0 * lbfgs_state_bytes());
    if (b.trace) b.trace += (size_t)w0 * b.trace_stride;
    return v;
}

// z: plain latent [W][n]; or, when z_hi/z_lo are given, the same already split into TF32 parts
// status: optional per-window status words of the slice (GEM_WIN_F16_RANGE is raised when an fp16 operand saturates)
static int decode_impl(gem_ctx* c, cudaStream_t s, int which, int W, const Slice& v_, const float* z, float* pose_out,
                       const void* z_hi = nullptr, const void* z_lo = nullptr, uint32_t* status = nullptr) {
    const gem_vae_weights& v = c->vae[which];
    const int T = c->T, M = W * T, P = c->J * 3;
    if (use_tc_chain(c, which, W)) {
        // latent -> [T][256] on the tcgen05 GEMM, its epilogue writes the activation already split
        GEM_TRY(run_layer(c, s, GEM_TAG_DEC + 0, v.dec[0], z, c->n, W, v_.act_hi[0], T * 256, EPI_LRELU, nullptr, z_hi,
                          z_lo, v_.act_lo[0], v_.act_sign[0], nullptr, c->gemm_mode == 3, nullptr, status));
        c->act_split = true;
        if (c->gemm_mode == 3 && c->tap_chain == 2) {
            // 256 -> 128 -> 64 -> 64 -> 64 -> pose in ONE launch on CTA pairs
            TapChainLaunch t;
            t.nl = 5;
            for (int i = 1; i <= 5; ++i) {
                t.B[i - 1] = v.dec[i].w_d, t.bias[i - 1] = v.dec[i].bias_d, t.aux_bits[i - 1] = nullptr;
                t.sign_out[i - 1] = i <= 4 ? v_.act_sign[i] : nullptr, t.epi[i - 1] = i <= 4 ? EPI_LRELU : EPI_NONE;
            }
            t.A_hi = v_.act_hi[0], t.A_lo = v_.act_lo[0], t.lda = v.dec[1].k, t.Kreal = v.dec[1].k;
            t.out_hi = pose_out, t.out_lo = nullptr, t.ldo = P, t.W = W, t.T = T, t.status = status;
            return timed(c, s, GEM_TAG_DEC + 1, [&]() { return launch_tap_chain_pair(s, c, t); });
        }
        if (c->gemm_mode == 3 && c->tap_chain) {
            // 256 -> 128 on its own, then 128 -> 64 -> 64 -> 64 -> pose in ONE launch: the activation tile stays in
            // shared memory, only the sign bits (the bwd-data masks) and the pose reach HBM
            GEM_TRY(run_tap_tc(c, s, GEM_TAG_DEC + 1, v.dec[1], v_.act_hi[0], v_.act_lo[0], v.dec[1].k, W, v_.act_hi[1],
                               v_.act_lo[1], v.dec[1].n, EPI_LRELU, nullptr, nullptr, v_.act_sign[1]));
            TapChainLaunch t;
            t.nl = 4;
            for (int i = 2; i <= 5; ++i) {
                t.B[i - 2] = v.dec[i].w_d, t.bias[i - 2] = v.dec[i].bias_d, t.aux_bits[i - 2] = nullptr;
                t.sign_out[i - 2] = i <= 4 ? v_.act_sign[i] : nullptr, t.epi[i - 2] = i <= 4 ? EPI_LRELU : EPI_NONE;
            }
            t.A_hi = v_.act_hi[1], t.A_lo = v_.act_lo[1], t.lda = v.dec[2].k, t.Kreal = v.dec[2].k;
            t.out_hi = pose_out, t.out_lo = nullptr, t.ldo = P, t.W = W, t.T = T, t.status = status;
            return timed(c, s, GEM_TAG_DEC + 2, [&]() { return launch_tap_chain(s, c, t); });
        }
        for (int i = 1; i <= 4; ++i)
            GEM_TRY(run_tap_tc(c, s, GEM_TAG_DEC + i, v.dec[i], v_.act_hi[i - 1], v_.act_lo[i - 1], v.dec[i].k, W,
                               v_.act_hi[i], v_.act_lo[i], v.dec[i].n, EPI_LRELU, nullptr, nullptr, v_.act_sign[i]));
        return run_tap_tc(c, s, GEM_TAG_DEC + 5, v.dec[5], v_.act_hi[4], v_.act_lo[4], 64, W, pose_out, nullptr, P, EPI_NONE,
                          nullptr);
    }
    GEM_TRY(run_layer(c, s, GEM_TAG_DEC + 0, v.dec[0], z, c->n, W, v_.act[0], T * 256, EPI_LRELU, nullptr));
    const float* in = v_.act[0];
    for (int i = 1; i <= 4; ++i) {
        GEM_TRY(run_layer(c, s, GEM_TAG_DEC + i, v.dec[i], in, v.dec[i].k, M, v_.act[i], v.dec[i].n, EPI_LRELU, nullptr));
        in = v_.act[i];
    }
    c->act_split = false;
    return run_layer(c, s, GEM_TAG_DEC + 5, v.dec[5], in, 64, M, pose_out, P, EPI_NONE, nullptr);
}

// dpose_is_split: the slice's gp_hi / gp_lo already hold dpose (written by the energy kernel)
// en: the chain kernel evaluates the energy itself (fused_energy_ok) and no d pose exists anywhere
static int decode_vjp_impl(gem_ctx* c, cudaStream_t s, int which, int W, const Slice& v_, const float* dpose, float* dz,
                           bool dpose_is_split = false, uint32_t* status = nullptr, const ChainEnergyLaunch* en = nullptr) {
    const gem_vae_weights& v = c->vae[which];
    const int T = c->T, M = W * T, P = c->J * 3;
    // the LeakyReLU derivative only needs the sign of the saved activation, which its TF32 hi part keeps
    float* const* saved = c->act_split ? v_.act_hi : v_.act;
    // dec_bwd[i] is the bwd-data of dec[5-i]; its output is d(pre-activation of dec[4-i])
    if (use_tc_chain(c, which, W)) {
        const int pp = pose_pad(c);
        if (!dpose_is_split) {
            if (c->gemm_mode == 3)
                GEM_TRY(timed(c, s, GEM_TAG_DEC_BWD + 0, [&]() {
                    return launch_rowscale_split_pad_f16(s, dpose, P, T, W, pp, (uint16_t*)v_.gp_hi, (uint16_t*)v_.gp_lo,
                                                         v_.row_exp);
                }));
            else
                GEM_TRY(timed(c, s, GEM_TAG_DEC_BWD + 0,
                              [&]() { return launch_split_pad(s, dpose, P, (size_t)M, pp, v_.gp_hi, v_.gp_lo); }));
        }
        const float *in_hi = v_.gp_hi, *in_lo = v_.gp_lo;
        int lda = pp;
        int i0 = 0;
        if (c->gemm_mode == 3 && c->tap_chain == 2 && c->act_split) {
            // d pose -> 64 -> 64 -> 64 -> 128 -> 256 in ONE launch on CTA pairs (masks from the forward pass's sign bits)
            TapChainLaunch t;
            t.nl = 5;
            for (int i = 0; i < 5; ++i) {
                t.B[i] = v.dec_bwd[i].w_d, t.bias[i] = v.dec_bwd[i].bias_d, t.aux_bits[i] = v_.act_sign[4 - i];
                t.sign_out[i] = nullptr, t.epi[i] = EPI_MASK;
            }
            t.A_hi = in_hi, t.A_lo = in_lo, t.lda = pp, t.Kreal = pp;
            t.out_hi = v_.gact_hi[0], t.out_lo = v_.gact_lo[0], t.ldo = v.dec_bwd[4].n, t.W = W, t.T = T, t.status = status;
            t.energy = en;
            GEM_TRY(timed(c, s, GEM_TAG_DEC_BWD + 0, [&]() { return launch_tap_chain_pair(s, c, t); }));
            i0 = 5;
        } else if (c->gemm_mode == 3 && c->tap_chain && c->act_split) {
            // d pose -> 64 -> 64 -> 64 -> 128 in ONE launch (masks from the forward pass's sign bits), 128 -> 256 after it
            TapChainLaunch t;
            t.nl = 4;
            for (int i = 0; i < 4; ++i) {
                t.B[i] = v.dec_bwd[i].w_d, t.bias[i] = v.dec_bwd[i].bias_d, t.aux_bits[i] = v_.act_sign[4 - i];
                t.sign_out[i] = nullptr, t.epi[i] = EPI_MASK;
            }
            t.A_hi = in_hi, t.A_lo = in_lo, t.lda = pp, t.Kreal = pp;
            t.out_hi = v_.gact_hi[1], t.out_lo = v_.gact_lo[1], t.ldo = v.dec_bwd[3].n, t.W = W, t.T = T, t.status = status;
            GEM_TRY(timed(c, s, GEM_TAG_DEC_BWD + 0, [&]() { return launch_tap_chain(s, c, t); }));
            in_hi = v_.gact_hi[1], in_lo = v_.gact_lo[1], lda = v.dec_bwd[3].n;
            i0 = 4;
        }
        for (int i = i0; i < 5; ++i) {
            const int a = 4 - i;
            GEM_TRY(run_tap_tc(c, s, GEM_TAG_DEC_BWD + i, v.dec_bwd[i], in_hi, in_lo, lda, W, v_.gact_hi[a], v_.gact_lo[a],
                               v.dec_bwd[i].n, EPI_MASK, saved[a], c->act_split ? v_.act_sign[a] : nullptr));
            in_hi = v_.gact_hi[a], in_lo = v_.gact_lo[a];
            lda = v.dec_bwd[i].n;
        }
        if (c->gemm_mode == 3)      // the chain is fp16 end to end; the rows still carry the energy kernel's 2^row_exp
            return run_layer(c, s, GEM_TAG_DEC_BWD + 5, v.dec_bwd[5], nullptr, T * 256, W, dz, c->n, EPI_NONE, nullptr,
                             v_.gact_hi[0], v_.gact_lo[0], nullptr, nullptr, v_.row_exp);
        if (c->gemm_mode == 2) {
            // fp16 scheme: gradient rows are rescaled to ~2^10 before the split (they shrink to 1e-7 near convergence)
            GEM_TRY(timed(c, s, GEM_TAG_DEC_BWD + 6, [&]() {
                return launch_rowscale_split_f16(s, v_.gact_hi[0], v_.gact_lo[0], W, T * 256, v_.g0_h16, v_.g0_l16, v_.row_exp);
            }));
            return run_layer(c, s, GEM_TAG_DEC_BWD + 5, v.dec_bwd[5], nullptr, T * 256, W, dz, c->n, EPI_NONE, nullptr,
                             v_.g0_h16, v_.g0_l16, nullptr, nullptr, v_.row_exp);
        }
        return run_layer(c, s, GEM_TAG_DEC_BWD + 5, v.dec_bwd[5], nullptr, T * 256, W, dz, c->n, EPI_NONE, nullptr,
                         v_.gact_hi[0], v_.gact_lo[0]);
    }
    const float* in = dpose;
    int lda = P;
    for (int i = 0; i < 5; ++i) {
        const int a = 4 - i;   // activation whose LeakyReLU derivative masks this output
        GEM_TRY(run_layer(c, s, GEM_TAG_DEC_BWD + i, v.dec_bwd[i], in, lda, M, v_.gact[a], v.dec_bwd[i].n, EPI_MASK,
                          saved[a], nullptr, nullptr, nullptr, nullptr, nullptr, 0,
                          c->act_split ? v_.act_sign[a] : nullptr));
        in = v_.gact[a];
        lda = v.dec_bwd[i].n;
    }
    return run_layer(c, s, GEM_TAG_DEC_BWD + 5, v.dec_bwd[5], v_.gact[0], T * 256, W, dz, c->n, EPI_NONE, nullptr);
}

static int encode_impl(gem_ctx* c, cudaStream_t s, int which, int W, const Slice& v_, const float* pose, const float* eps,
                       float* z0, float* mu, float* sd, size_t eps_stride = 0, uint32_t* status = nullptr) {
    const gem_vae_weights& v = c->vae[which];
    const int T = c->T, M = W * T, P = c->J * 3;
    const float* in = pose;
    int lda = P;
    const gem_layer& fcL = v.enc[5];
    // The encoder's 128 -> 256 -> 512 layers run on the tcgen05 tap kernel (6x faster than the CUDA-core layers).  z0
    // seeds a chaotic line search, so the choice was held against the distribution test
    // (tests/test_gpu_dist.py, 64 windows x 2 stages x {3, 25} iterations): strict windows 42 / 23 / 16 / 20 with the
    // tensor-core layers, 40 / 24 / 17 / 22 with the CUDA-core ones — the same within the coin tosses (round 1 kept the
    // CUDA-core layers on a 21-vs-19-of-42 count, which was noise).  GEM_ENC_TC=0 selects the CUDA-core layers.
    const bool enc_tc = c->enc_tc_on && c->gemm_mode == 3 && c->enc_tc[which] && fcL.k % 64 == 0 && fcL.n % 128 == 0;
    for (int i = 0; i < (enc_tc ? 3 : 5); ++i) {
        GEM_TRY(run_layer(c, s, GEM_TAG_ENC + i, v.enc[i], in, lda, M, v_.eact[i], v.enc[i].n, EPI_LRELU, nullptr));
        in = v_.eact[i];
        lda = v.enc[i].n;
    }
    if (enc_tc) {
        // 128 -> 256 -> 512 on the tcgen05 tap kernel (fp16 scheme); its last epilogue writes the fc GEMM's operand
        GEM_TRY(timed(c, s, GEM_TAG_ENC + 3, [&]() {
            return launch_split_f16(s, in, v.enc[3].k, M, v.enc[3].k, nullptr, v_.e2_hi, v_.e2_lo);
        }));
        GEM_TRY(run_tap_tc(c, s, GEM_TAG_ENC + 3, v.enc[3], (const float*)v_.e2_hi, (const float*)v_.e2_lo, v.enc[3].k, W,
                           (float*)v_.e3_hi, (float*)v_.e3_lo, v.enc[3].n, EPI_LRELU, nullptr));
        GEM_TRY(run_tap_tc(c, s, GEM_TAG_ENC + 4, v.enc[4], (const float*)v_.e3_hi, (const float*)v_.e3_lo, v.enc[4].k, W,
                           v_.e4_hi, v_.e4_lo, v.enc[4].n, EPI_LRELU, nullptr));
        GEM_TRY(run_layer(c, s, GEM_TAG_ENC + 5, fcL, nullptr, T * 512, W, v_.fc, 2 * c->n, EPI_NONE, nullptr, v_.e4_hi,
                          v_.e4_lo));
    } else if (c->gemm_mode >= 2 && fcL.k % 64 == 0 && fcL.n % 128 == 0) {
        GEM_TRY(timed(c, s, GEM_TAG_ENC + 5, [&]() {
            return launch_split_f16(s, in, T * 512, W, T * 512, nullptr, (uint16_t*)v_.e4_hi, (uint16_t*)v_.e4_lo, status);
        }));
        GEM_TRY(run_layer(c, s, GEM_TAG_ENC + 5, fcL, nullptr, T * 512, W, v_.fc, 2 * c->n, EPI_NONE, nullptr, v_.e4_hi,
                          v_.e4_lo));
    } else if (c->gemm_mode == 1 && fcL.k % 32 == 0 && fcL.n % 128 == 0) {
        // split into the slice's own buffers (the GEMM's internal split scratch is shared by the whole ctx)
        GEM_TRY(timed(c, s, GEM_TAG_ENC + 5,
                      [&]() { return launch_split_tf32(s, in, T * 512, W, T * 512, v_.e4_hi, v_.e4_lo); }));
        GEM_TRY(run_layer(c, s, GEM_TAG_ENC + 5, fcL, nullptr, T * 512, W, v_.fc, 2 * c->n, EPI_NONE, nullptr, v_.e4_hi,
                          v_.e4_lo));
    } else {
        GEM_TRY(run_layer(c, s, GEM_TAG_ENC + 5, fcL, in, T * 512, W, v_.fc, 2 * c->n, EPI_NONE, nullptr));
    }
    return timed(c, s, GEM_TAG_REPARAM, [&]() {
        return launch_reparam(s, v_.fc, eps, eps_stride ? eps_stride : (size_t)c->n, z0, mu, sd, W, c->n);
    });
}

// copies a stage's per-window inputs into ctx-owned buffers: anchor pose, first heat-map frame, the window's
// mean bone lengths (gathered through its clip index, so the staged clip index is the window itself)
__global__ void stage_inputs_kernel(int w_abs0, int TJ3, int J, const float* __restrict__ pose0, const int64_t* __restrict__ fb,
                                    const int32_t* __restrict__ clip, const float* __restrict__ mean_bone,
                                    float* __restrict__ pose0_own, int64_t* __restrict__ fb_own,
                                    int32_t* __restrict__ clip_own, float* __restrict__ mb_own,
                                    uint32_t* __restrict__ status_own, unsigned long long* __restrict__ patch_valid, int TJ,
                                    int keep_patch, uint32_t* __restrict__ miss_count) {
    // all pointers are the slice's (window blockIdx.x of the slice); mb_own is the ctx-wide table indexed by the
    // absolute window w_abs0 + blockIdx.x, which is what the staged clip index points at
    const int w = blockIdx.x;
    for (int i = threadIdx.x; i < TJ3; i += blockDim.x) pose0_own[(size_t)w * TJ3 + i] = pose0[(size_t)w * TJ3 + i];
    if (threadIdx.x < J) mb_own[(size_t)(w_abs0 + w) * J + threadIdx.x] = mean_bone[(size_t)clip[w] * J + threadIdx.x];
    // a new stage reads new maps: every joint's cached texel patch is stale
    if (!keep_patch && patch_valid)
        for (int i = threadIdx.x; i < TJ; i += blockDim.x) patch_valid[(size_t)w * TJ + i] = 0ull;      // empty window
    if (miss_count && w == 0 && threadIdx.x < 2) miss_count[threadIdx.x] = 0u;     // (an aborted call may have left entries)
    if (threadIdx.x == 0) {
        fb_own[w] = fb ? fb[w] : 0;
        clip_own[w] = w_abs0 + w;
        status_own[w] = 0;
    }
}

#define GEM_ENTER(c, W)                                                          \
    GEM_REQUIRE((c) != nullptr, "ctx is NULL");                                  \
    if ((W) > (c)->Wmax) {                                                       \
        set_error("more windows than the ctx was created for");                  \
        return GEM_ERR_CAPACITY;                                                 \
    }                                                                            \
    GEM_REQUIRE((W) >= 0, "W must be >= 0");                                      \
    if ((W) == 0) return GEM_OK; /* no windows: a no-op (empty buffers have no address to check) */ \
    GEM_CUDA(cudaSetDevice((c)->device))

extern "C" {

int gem_energy_grad(gem_ctx* c, void* stream, int W, const float* pose_d, const float* pose0_d, const float* heat_d,
                    const int64_t* frame_base_d, const int32_t* clip_d, const float* mean_bone_d,
                    const gem_energy_weights* wt, float* energy_d, float* terms_d, float* grad_d, uint32_t* status_d) {
    GEM_ENTER(c, W);
    GEM_REQUIRE(pose_d && pose0_d && clip_d && mean_bone_d && wt && energy_d && grad_d, "NULL argument");
    if (wt->reproj != 0.f && !c->have_camera) {
        set_error("camera not set");
        return GEM_ERR_STATE;
    }
    return timed(c, (cudaStream_t)stream, GEM_TAG_ENERGY, [&]() {
        return launch_energy_grad((cudaStream_t)stream, c->have_camera ? &c->cam : nullptr, &c->skel, W, c->T, c->J, c->H,
                                  c->Wd, pose_d, pose0_d, heat_d,
                                  frame_base_d, clip_d, mean_bone_d, *wt, energy_d, terms_d, grad_d, status_d, nullptr, nullptr,
                                  0, nullptr, nullptr, nullptr, 0, nullptr, nullptr, c->heat_planar);
    });
}

int gem_gemm(gem_ctx* c, void* stream, int M, int N, int K, const float* a_d, int lda, const float* b_d,
             const float* bias_d, int leaky_relu, float* c_d, int ldc, int use_tensor_cores) {
    GEM_REQUIRE(c && a_d && b_d && c_d && M >= 0 && N > 0 && K > 0, "bad arguments");
    GEM_REQUIRE(N % 4 == 0, "N must be a multiple of 4");
    GEM_CUDA(cudaSetDevice(c->device));
    gem_layer L;
    L.w_d = b_d, L.bias_d = bias_d, L.taps = 1, L.k = K, L.n = N;
    const int saved = c->gemm_mode;
    if (use_tensor_cores) GEM_TRY(tc_gemm_prepare_weight(c, (cudaStream_t)stream, b_d, N, K, N, use_tensor_cores == 2 ? 2 : 1));
    c->gemm_mode = use_tensor_cores == 2 ? 2 : (use_tensor_cores ? 1 : 0);
    const int rc = run_layer(c, (cudaStream_t)stream, GEM_TAG_OTHER, L, a_d, lda, M, c_d, ldc,
                             leaky_relu ? EPI_LRELU : EPI_NONE, nullptr);
    c->gemm_mode = saved;
    return rc;
}

int gem_decode(gem_ctx* c, void* stream, int which, int W, const float* z_d, float* pose_d) {
    GEM_ENTER(c, W);
    GEM_REQUIRE((which == 0 || which == 1) && z_d && pose_d, "bad arguments");
    if (!c->have_vae[which]) {
        set_error("VAE weights not set");
        return GEM_ERR_STATE;
    }
    return decode_impl(c, (cudaStream_t)stream, which, W, slice_of(c, 0), z_d, pose_d);
}

int gem_decode_vjp(gem_ctx* c, void* stream, int which, int W, const float* dpose_d, float* dz_d) {
    GEM_ENTER(c, W);
    GEM_REQUIRE((which == 0 || which == 1) && dpose_d && dz_d, "bad arguments");
    if (!c->have_vae[which]) {
        set_error("VAE weights not set");
        return GEM_ERR_STATE;
    }
    return decode_vjp_impl(c, (cudaStream_t)stream, which, W, slice_of(c, 0), dpose_d, dz_d);
}

int gem_encode(gem_ctx* c, void* stream, int which, int W, const float* pose_d, const float* eps_d, float* z0_d,
               float* mu_d, float* std_d) {
    GEM_ENTER(c, W);
    GEM_REQUIRE((which == 0 || which == 1) && pose_d && eps_d && z0_d, "bad arguments");
    if (!c->have_vae[which]) {
        set_error("VAE weights not set");
        return GEM_ERR_STATE;
    }
    return encode_impl(c, (cudaStream_t)stream, which, W, slice_of(c, 0), pose_d, eps_d, z0_d, mu_d, std_d);
}

static int lbfgs_configure(gem_ctx* c, const gem_lbfgs_params* p, float* trace, int trace_stride) {
    GEM_REQUIRE(p && p->max_iter >= 1 && p->max_eval >= 1 && p->lr > 0, "bad L-BFGS parameters");
    GEM_REQUIRE(p->max_iter - 1 <= c->m, "max_iter - 1 exceeds the ctx's max_history");
    c->lb.lr = p->lr, c->lb.tol_grad = p->tolerance_grad, c->lb.tol_change = p->tolerance_change;
    c->lb.max_iter = p->max_iter, c->lb.max_eval = p->max_eval;
    c->lb.trace = trace, c->lb.trace_stride = trace_stride;
    c->lb_params = *p;
    return GEM_OK;
}

int gem_lbfgs_begin(gem_ctx* c, void* stream, int W, const float* z0_d, const gem_lbfgs_params* params_h) {
    GEM_ENTER(c, W);
    GEM_REQUIRE(z0_d != nullptr, "z0 is NULL");
    GEM_TRY(lbfgs_configure(c, params_h, nullptr, 0));
    c->lb_started = true;
    return timed(c, (cudaStream_t)stream, GEM_TAG_LBFGS_BEGIN,
                 [&]() { return launch_lbfgs_begin((cudaStream_t)stream, c->lb, z0_d, W); });
}

const float* gem_lbfgs_trial(gem_ctx* c) { return c ? c->lb.ZT : nullptr; }
const float* gem_lbfgs_x(gem_ctx* c) { return c ? c->lb.X : nullptr; }

int gem_lbfgs_advance(gem_ctx* c, void* stream, int W, const float* loss_d, const float* grad_d) {
    GEM_ENTER(c, W);
    GEM_REQUIRE(loss_d && grad_d, "NULL argument");
    if (!c->lb_started) {
        set_error("gem_lbfgs_begin has not been called");
        return GEM_ERR_STATE;
    }
    return timed(c, (cudaStream_t)stream, GEM_TAG_LBFGS_ADVANCE,
                 [&]() { return launch_lbfgs_advance((cudaStream_t)stream, c->lb, loss_d, grad_d, W); });
}

__global__ void copy_trace_kernel(const float* src, int src_stride, float* dst, int dst_stride, int W, int cols) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)W * cols) return;
    const int w = (int)(i / cols), k = (int)(i - (size_t)w * cols);
    dst[(size_t)w * dst_stride + k] = src[(size_t)w * src_stride + k];
}

int gem_lbfgs_stats(gem_ctx* c, void* stream, int W, int32_t* n_iter_d, int32_t* func_evals_d, int32_t* finished_d,
                    double* t_d, double* loss_d, float* trace_d, int trace_stride) {
    GEM_ENTER(c, W);
    if (trace_d && c->lb.trace && W > 0) {
        const int cols = trace_stride < c->lb.trace_stride ? trace_stride : c->lb.trace_stride;
        const size_t total = (size_t)W * cols;
        copy_trace_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
            c->lb.trace, c->lb.trace_stride, trace_d, trace_stride, W, cols);
        GEM_CHECK_LAUNCH();
    }
    return launch_lbfgs_stats((cudaStream_t)stream, c->lb, W, n_iter_d, func_evals_d, finished_d, t_d, loss_d);
}

}  // extern "C"

// Everything one stage does for the windows [w0, w0 + Wk) of a call, enqueued on stream q (the slice's own):
// input staging, encoder, L-BFGS begin, max_eval + 1 closure rounds (round 0 launched directly, the rest a
// replayed CUDA graph), final decode and counters.  All pointers are the call's (un-sliced) arguments.
struct StageCall {
    int which;
    const float* pose0;         // [W][T][J][3]
    const float* heat;
    const int64_t* frame_base;  // [W] or NULL
    const int32_t* clip;        // [W]
    const float* mean_bone;
    const float* eps;           // [W][eps_stride]
    size_t eps_stride;
    gem_energy_weights wt;
    gem_lbfgs_params p;
    float* pose_out;            // [W][T][J][3]
    float* trace;               // [W][max_eval + 1] or NULL
    int32_t *n_iter, *func_evals;
    uint32_t* status;
    bool status_or = false;     // OR this stage's status bits into `status` instead of overwriting it
    bool texel_cache;           // read the maps through the per-joint texel cache
    bool host_maps = false;     // the maps are host memory read over PCIe: texel prefetch kernel, more slices
};

__global__ void or_status_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] |= src[i];
}

// The texel cache pays when the maps are read over PCIe from (pinned / registered) host memory.  On device-resident
// maps it cuts the energy kernel's DRAM traffic but never its run time: measured again in round 2 at 100 980 windows
// (0.95 ms per launch with the cache, 0.88 ms without: the kernel is bound by the latency chain of its short-lived
// CTAs, not by sectors), so "auto" means host memory only.
static bool is_host_memory(const void* p) {
    cudaPointerAttributes at;
    if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}
static bool want_texel_cache(const gem_ctx* c, const void* heat, float reproj, int W) {
    if (!heat || reproj == 0.f || c->texel_cache == 0) return false;
    if (c->texel_cache == 1) return true;
    (void)W;
    return is_host_memory(heat);
}

static int enqueue_stage_slice(gem_ctx* c, cudaStream_t q, const StageCall& a, int w0, int Wk, bool graphs) {
    if (Wk <= 0) return GEM_OK;
    const int which = a.which;
    const int TJ3 = c->T * c->J * 3;
    const size_t P = (size_t)TJ3;
    const int trace_cols = a.p.max_eval + 1;
    Slice v = slice_of(c, w0);
    // L-BFGS view of this slice with the stage's parameters
    LbfgsBuffers& lb = v.lb;
    lb.lr = a.p.lr, lb.tol_grad = a.p.tolerance_grad, lb.tol_change = a.p.tolerance_change;
    lb.max_iter = a.p.max_iter, lb.max_eval = a.p.max_eval;
    lb.zt_f16 = c->gemm_mode >= 2;
    lb.status = v.status_own;
    lb.trace = a.trace ? v.trace_own : nullptr;
    lb.trace_stride = a.trace ? c->trace_cap : 0;
    if (a.trace) GEM_CUDA(cudaMemsetAsync(v.trace_own, 0xff, (size_t)Wk * c->trace_cap * sizeof(float), q));   // NaN
    stage_inputs_kernel<<<Wk, 128, 0, q>>>(w0, TJ3, c->J, a.pose0 + w0 * P, a.frame_base ? a.frame_base + w0 : nullptr,
                                           a.clip + w0, a.mean_bone, v.pose0_own, v.fb_own, v.clip_own, v.mb_own,
                                           v.status_own, v.patch_valid, c->T * c->J, c->texel_keep, v.miss_count);
    GEM_CHECK_LAUNCH();
    c->launches += 1;
    // z0 = mu + eps * std                                   optimizer.py:255-259
    GEM_TRY(encode_impl(c, q, which, Wk, v, a.pose0 + w0 * P, a.eps + (size_t)w0 * a.eps_stride, v.z0, nullptr, nullptr,
                        a.eps_stride, v.status_own));
    GEM_TRY(timed(c, q, GEM_TAG_LBFGS_BEGIN, [&]() { return launch_lbfgs_begin(q, lb, v.z0, Wk); }));
    const bool tc = use_tc_chain(c, which, Wk);
    const bool fused = fused_energy_ok(c, which, Wk);
    // one closure round: decode -> fused energy/gradient -> decoder bwd-data -> L-BFGS advance
    bool first_round = true;                     // (round 0 is launched directly, never captured)
    auto enqueue_round = [&]() -> int {
        GEM_TRY(decode_impl(c, q, which, Wk, v, lb.ZT, v.pose, tc ? lb.ZT_hi : nullptr, tc ? lb.ZT_lo : nullptr, v.status_own));
        if (a.host_maps && c->texel_prefetch_ctas > 0 && c->heat_planar && c->texel_probe) {
            // zero-copy planar maps: a probe over all joints lists the misses, a few CTAs fetch them over PCIe
            const bool cold = first_round && c->texel_cold_chain && !c->texel_keep;
            if (cold && c->cold_used > 0) GEM_CUDA(cudaStreamWaitEvent(q, c->cold_ev[c->cold_used - 1], 0));
            c->launches += 1;
            GEM_TRY(timed(c, q, GEM_TAG_ENERGY, [&]() {
                return launch_texel_probe_fetch(q, &c->cam, Wk, c->T, c->J, c->H, c->Wd, v.pose, a.heat, v.fb_own, v.patch,
                                                v.patch_origin, v.patch_valid, c->patch_stats_on ? c->patch_stats : nullptr,
                                                cold ? c->texel_cold_ctas : c->texel_prefetch_ctas, c->texel_prefetch_threads,
                                                v.miss_count, v.miss_list, cold ? c->texel_cold_rows : c->texel_rows, c->heat_planar);
            }));
            if (cold) {
                if ((size_t)c->cold_used == c->cold_ev.size()) {
                    cudaEvent_t ev;
                    GEM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                    c->cold_ev.push_back(ev);
                }
                GEM_CUDA(cudaEventRecord(c->cold_ev[c->cold_used++], q));
            }
            first_round = false;
        } else if (a.host_maps && c->texel_prefetch_ctas > 0)
            // zero-copy maps: the wait for PCIe happens in a few CTAs, not in energy CTAs parked on every SM
            GEM_TRY(timed(c, q, GEM_TAG_ENERGY, [&]() {
                return launch_texel_prefetch(q, &c->cam, Wk, c->T, c->J, c->H, c->Wd, v.pose, a.heat, v.fb_own, v.patch, v.patch_origin,
                                             v.patch_valid, c->patch_stats_on ? c->patch_stats : nullptr,
                                             c->texel_prefetch_ctas, c->heat_planar, c->texel_prefetch_threads);
            }));
        if (fused) {
            // the backward chain evaluates the energy in its prologue: no energy launch, no d pose in memory
            ChainEnergyLaunch en;
            en.cam = c->have_camera ? &c->cam : nullptr, en.skel = &c->skel;
            en.pose = v.pose, en.pose0 = v.pose0_own, en.heat = a.heat, en.frame_base = v.fb_own, en.clip = v.clip_own;
            en.mean_bone = v.mb_own, en.wt = a.wt, en.energy = v.f_new, en.status = v.status_own, en.row_exp = v.row_exp;
            en.J = c->J, en.H = c->H, en.Wd = c->Wd, en.planar = c->heat_planar;
            if (a.texel_cache) {
                en.patch = v.patch, en.patch_origin = v.patch_origin, en.patch_valid = v.patch_valid;
                en.patch_stats = c->patch_stats_on ? c->patch_stats : nullptr;
            }
            GEM_TRY(decode_vjp_impl(c, q, which, Wk, v, nullptr, v.g_new, true, v.status_own, &en));
        } else {
            GEM_TRY(timed(c, q, GEM_TAG_ENERGY, [&]() {
                return launch_energy_grad(q, c->have_camera ? &c->cam : nullptr, &c->skel, Wk, c->T, c->J, c->H, c->Wd, v.pose,
                                          v.pose0_own, a.heat, v.fb_own, v.clip_own,
                                          v.mb_own, a.wt, v.f_new, nullptr, tc ? nullptr : v.gpose, v.status_own, tc ? v.gp_hi : nullptr,
                                          v.gp_lo, pose_pad(c), a.texel_cache ? v.patch : nullptr, v.patch_origin,
                                          c->patch_stats_on ? c->patch_stats : nullptr, c->gemm_mode == 3, v.row_exp, v.patch_valid,
                                          c->heat_planar);
            }));
            GEM_TRY(decode_vjp_impl(c, q, which, Wk, v, v.gpose, v.g_new, tc, v.status_own));
        }
        return timed(c, q, GEM_TAG_LBFGS_ADVANCE, [&]() { return launch_lbfgs_advance(q, lb, v.f_new, v.g_new, Wk); });
    };
    // LBFGS.step: at most max_eval + 1 closure evaluations per window (lbfgs.py:478-487, App. B)
    GEM_TRY(enqueue_round());                        // round 0: direct (also performs every lazy initialisation)
    cudaGraphExec_t exec = nullptr;
    int round_launches = 0;
    if (graphs && a.p.max_eval >= 2) {
        for (auto& g : c->graphs) {
            if (g.which == which && g.w0 == w0 && g.Wk == Wk && g.gemm_mode == (c->gemm_mode | (c->tap_chain << 8) | (c->fuse_energy << 12)) && g.heat == a.heat &&
                g.has_heat == ((a.wt.reproj != 0.f) + 2 * (int)a.texel_cache + 4 * (int)c->patch_stats_on + 8 * (a.host_maps ? c->texel_prefetch_ctas + 512 * (c->texel_prefetch_threads / 32) : 0) + (1 << 20) * c->heat_planar + (1 << 22) * c->texel_probe) &&
                g.trace_stride == lb.trace_stride &&
                memcmp(&g.wt, &a.wt, sizeof(a.wt)) == 0 && g.p.lr == a.p.lr && g.p.max_iter == a.p.max_iter &&
                g.p.max_eval == a.p.max_eval && g.p.tolerance_grad == a.p.tolerance_grad &&
                g.p.tolerance_change == a.p.tolerance_change) {
                exec = g.exec, round_launches = g.launches;
                break;
            }
        }
        if (!exec) {
            cudaGraph_t graph = nullptr;
            GEM_CUDA(cudaStreamBeginCapture(q, cudaStreamCaptureModeRelaxed));
            const int64_t launches_before = c->launches;
            int rc_round = GEM_OK;
            for (int round = 1; round <= a.p.max_eval && rc_round == GEM_OK; ++round) rc_round = enqueue_round();   // ONE graph of all replayed rounds: kernel-to-kernel edges instead of graph-to-graph launch gaps
            round_launches = (int)(c->launches - launches_before);
            c->launches = launches_before;               // captured, not executed
            cudaError_t e = cudaStreamEndCapture(q, &graph);
            if (rc_round != GEM_OK) {
                if (graph) cudaGraphDestroy(graph);
                return rc_round;
            }
            GEM_CUDA(e);
            gem_ctx::RoundGraph g;
            g.which = which, g.w0 = w0, g.Wk = Wk, g.gemm_mode = c->gemm_mode | (c->tap_chain << 8) | (c->fuse_energy << 12), g.heat = a.heat;
            g.has_heat = (a.wt.reproj != 0.f) + 2 * (int)a.texel_cache + 4 * (int)c->patch_stats_on + 8 * (a.host_maps ? c->texel_prefetch_ctas + 512 * (c->texel_prefetch_threads / 32) : 0) + (1 << 20) * c->heat_planar + (1 << 22) * c->texel_probe;
            g.trace_stride = lb.trace_stride, g.wt = a.wt, g.p = a.p;
            g.launches = round_launches;
            e = cudaGraphInstantiate(&g.exec, graph, 0);
            cudaGraphDestroy(graph);
            GEM_CUDA(e);
            if (c->graphs.size() >= 64) {
                cudaGraphExecDestroy(c->graphs.front().exec);
                c->graphs.erase(c->graphs.begin());
            }
            c->graphs.push_back(g);
            exec = g.exec;
        }
    }
    if (exec) {
        GEM_CUDA(cudaGraphLaunch(exec, q));                   // rounds 1 .. max_eval
        c->launches += round_launches;
    } else {
        for (int round = 1; round <= a.p.max_eval; ++round) GEM_TRY(enqueue_round());
    }
    // final decode of the optimum (every window is parked with its trial point = x)      optimizer.py:273-276
    GEM_TRY(decode_impl(c, q, which, Wk, v, lb.X, a.pose_out + w0 * P, tc ? lb.ZT_hi : nullptr, tc ? lb.ZT_lo : nullptr,
                        v.status_own));
    GEM_TRY(timed(c, q, GEM_TAG_OTHER, [&]() {
        return launch_lbfgs_stats(q, lb, Wk, a.n_iter ? a.n_iter + w0 : nullptr, a.func_evals ? a.func_evals + w0 : nullptr,
                                  nullptr, nullptr, nullptr);
    }));
    if (a.status && a.status_or) {
        or_status_kernel<<<(Wk + 255) / 256, 256, 0, q>>>(a.status + w0, v.status_own, Wk);
        GEM_CHECK_LAUNCH();
        c->launches += 1;
    } else if (a.status) {
        GEM_CUDA(cudaMemcpyAsync(a.status + w0, v.status_own, (size_t)Wk * sizeof(uint32_t), cudaMemcpyDeviceToDevice, q));
    }
    if (a.trace)
        GEM_CUDA(cudaMemcpy2DAsync(a.trace + (size_t)w0 * trace_cols, trace_cols * sizeof(float), v.trace_own,
                                   c->trace_cap * sizeof(float), trace_cols * sizeof(float), Wk, cudaMemcpyDeviceToDevice, q));
    return GEM_OK;
}

// Slice boundaries of a call over W windows: the caller's (gem_ctx_set_slices) when they fit, else n_chunks
// equal parts at multiples of 12 windows (the tap kernel's M tile), at least 96 windows each.
static std::vector<int> slice_bounds(const gem_ctx* c, int W, bool zero_copy, bool shared_scratch) {
    std::vector<int> w0;
    // layer shapes the chain / tap kernels do not take (T > 32, odd channel counts) fall back to GEMM launches that
    // split their operand into ONE ctx-wide scratch buffer: concurrent slices would overwrite each other's operand
    if (shared_scratch) {
        w0 = {0, W};
        return w0;
    }
    if (!c->prof_on && !c->user_slices.empty() && c->user_slices.back() < W) {
        w0 = c->user_slices;
        w0.push_back(W);
        return w0;
    }
    // automatic: up to 4 slices; up to 8 when the maps are read over PCIe (more slices hide more of the transfers).
    // The GEMMs run 256-row CTA-pair tiles (a row = a window): among the admissible slice counts take the one whose
    // slices pad to the fewest tile rows in total, the larger count on ties (468 windows: 2 slices of 234, not 4 of 117
    // that would each fill less than half a tile: DESIGN.md section 7).
    const int kAlign = 12, kMinChunk = 96, kTile = 256;
    int n;
    if (c->prof_on) {
        n = 1;
    } else if (c->n_chunks > 0) {
        n = c->n_chunks;
    } else {
        const int n_max = zero_copy ? 8 : 4;
        const int k_max = n_max < W / kMinChunk ? n_max : (W / kMinChunk > 1 ? W / kMinChunk : 1);
        auto padded = [&](int k) { return (long)k * (((W + k - 1) / k + kTile - 1) / kTile) * kTile; };
        long least = padded(1);
        for (int k = 2; k <= k_max; ++k) least = padded(k) < least ? padded(k) : least;
        n = 1;
        for (int k = 2; k <= k_max; ++k)          // concurrency is worth ~20 % of a step: accept up to 25 % more tile rows for it
            if (4 * padded(k) <= 5 * least) n = k;
    }
    if (n > W / kMinChunk) n = W / kMinChunk;
    if (n < 1) n = 1;
    w0.assign(n + 1, W);
    w0[0] = 0;
    for (int k = 1; k < n; ++k) w0[k] = (int)((int64_t)W * k / n) / kAlign * kAlign;
    return w0;
}

struct Fork {
    std::vector<cudaStream_t> cs;
    bool forked = false;
};
// slice k runs on its own stream, after everything already on the caller's stream `s` and after the ready
// events that cover its windows (gem_ctx_set_ready_events; consumed by this call)
static int fork_slices(gem_ctx* c, cudaStream_t s, const std::vector<int>& w0, bool graphs, Fork* f) {
    c->cold_used = 0;
    const int n = (int)w0.size() - 1;
    f->forked = n > 1 || graphs || !c->ready.empty();      // (the caller's stream may be the legacy one: not capturable)
    f->cs.assign(n, s);
    if (!f->forked) return GEM_OK;
    GEM_TRY(ensure_streams(c, n));
    GEM_CUDA(cudaEventRecord(c->fork_ev, s));
    for (int k = 0; k < n; ++k) {
        f->cs[k] = c->streams[k];
        GEM_CUDA(cudaStreamWaitEvent(f->cs[k], c->fork_ev, 0));
        for (auto& r : c->ready)           // events are in window order: wait for all that cover windows below the slice end
            if (r.first_window < w0[k + 1]) GEM_CUDA(cudaStreamWaitEvent(f->cs[k], r.ev, 0));
    }
    c->ready.clear();
    return GEM_OK;
}
static int join_slices(gem_ctx* c, cudaStream_t s, const Fork& f) {
    if (!f.forked) return GEM_OK;
    for (size_t k = 0; k < f.cs.size(); ++k) {
        GEM_CUDA(cudaEventRecord(c->join_ev[k], f.cs[k]));
        GEM_CUDA(cudaStreamWaitEvent(s, c->join_ev[k], 0));
    }
    return GEM_OK;
}

static int check_stage_args(gem_ctx* c, int which, const gem_energy_weights* wt, const gem_lbfgs_params* p, bool want_trace) {
    GEM_REQUIRE(p && p->max_iter >= 1 && p->max_eval >= 1 && p->lr > 0, "bad L-BFGS parameters");
    GEM_REQUIRE(p->max_iter - 1 <= c->m, "max_iter - 1 exceeds the ctx's max_history");
    if (!c->have_vae[which] || (wt->reproj != 0.f && !c->have_camera)) {
        set_error("VAE weights / camera not set");
        return GEM_ERR_STATE;
    }
    if (want_trace && p->max_eval + 1 > c->trace_cap) {
        set_error("max_eval exceeds the trace capacity of the ctx (raise max_history)");
        return GEM_ERR_CAPACITY;
    }
    return GEM_OK;
}

extern "C" {

int gem_ctx_set_slices(gem_ctx* c, int n, const int32_t* first_window_h) {
    GEM_REQUIRE(c != nullptr && n >= 0 && n <= 64 && (n == 0 || first_window_h), "bad arguments");
    std::vector<int> v;
    for (int i = 0; i < n; ++i) {
        GEM_REQUIRE(first_window_h[i] % 2 == 0 && (i == 0 ? first_window_h[i] == 0 : first_window_h[i] > first_window_h[i - 1]),
                    "slice starts must be even, strictly increasing and begin at 0");
        v.push_back(first_window_h[i]);
    }
    c->user_slices = v;
    return GEM_OK;
}

int gem_ctx_set_ready_events(gem_ctx* c, int n, const int32_t* first_window_h, void* const* events_h) {
    GEM_REQUIRE(c != nullptr && n >= 0 && (n == 0 || (first_window_h && events_h)), "bad arguments");
    c->ready.clear();
    for (int i = 0; i < n; ++i) c->ready.push_back({first_window_h[i], (cudaEvent_t)events_h[i]});
    return GEM_OK;
}

int gem_solve_stage(gem_ctx* c, void* stream, int which, int W, const float* pose0_d, const float* heat_d,
                    const int64_t* frame_base_d, const int32_t* clip_d, const float* mean_bone_d, const float* eps_d,
                    const gem_energy_weights* wt, const gem_lbfgs_params* params_h, float* pose_out_d,
                    float* energy_trace_d, int32_t* n_iter_d, int32_t* func_evals_d, uint32_t* status_d) {
    GEM_ENTER(c, W);
    GEM_REQUIRE((which == 0 || which == 1) && pose0_d && clip_d && mean_bone_d && eps_d && wt && params_h && pose_out_d,
                "bad arguments");
    GEM_REQUIRE(wt->reproj == 0.f || (heat_d && frame_base_d), "heatmaps required when reproj != 0");
    GEM_TRY(check_stage_args(c, which, wt, params_h, energy_trace_d != nullptr));
    cudaStream_t s = (cudaStream_t)stream;
    if (W == 0) return GEM_OK;
    GEM_TRY(lbfgs_configure(c, params_h, nullptr, 0));
    StageCall a;
    a.which = which, a.pose0 = pose0_d, a.heat = heat_d, a.frame_base = frame_base_d, a.clip = clip_d;
    a.mean_bone = mean_bone_d, a.eps = eps_d, a.eps_stride = (size_t)c->n, a.wt = *wt, a.p = *params_h;
    a.pose_out = pose_out_d, a.trace = energy_trace_d, a.n_iter = n_iter_d, a.func_evals = func_evals_d, a.status = status_d;
    a.texel_cache = want_texel_cache(c, heat_d, wt->reproj, W);
    a.host_maps = a.texel_cache && is_host_memory(heat_d);
    if (a.texel_cache) GEM_TRY(ensure_texel_cache(c));
    const bool graphs = c->use_graphs && !c->prof_on;
    const std::vector<int> w0 = slice_bounds(c, W, a.host_maps, c->gemm_mode >= 1 && !c->tap_tc[which]);
    Fork f;
    GEM_TRY(fork_slices(c, s, w0, graphs, &f));
    for (size_t k = 0; k + 1 < w0.size(); ++k) GEM_TRY(enqueue_stage_slice(c, f.cs[k], a, w0[k], w0[k + 1] - w0[k], graphs));
    c->lb_started = true;
    return join_slices(c, s, f);
}

int gem_solve_windows(gem_ctx* c, void* stream, int W, const float* pose0_d, const float* heat_d,
                      const int64_t* frame_base_d, const int32_t* clip_d, const float* mean_bone_d, const double* cams_d,
                      const float* eps_d, const gem_energy_weights* wt_local, const gem_energy_weights* wt_global,
                      const gem_lbfgs_params* params_h, float* local_pose_d, double* rel_f64_d, float* rel_f32_d,
                      float* global_pose_d, int32_t* n_iter_d, int32_t* func_evals_d, uint32_t* status_d) {
    GEM_ENTER(c, W);
    GEM_REQUIRE(pose0_d && clip_d && mean_bone_d && cams_d && eps_d && wt_local && wt_global && params_h && local_pose_d &&
                    rel_f32_d && global_pose_d,
                "bad arguments");
    GEM_REQUIRE(wt_local->reproj == 0.f || (heat_d && frame_base_d), "heatmaps required when reproj != 0");
    GEM_REQUIRE(wt_global->reproj == 0.f, "the global stage has no reprojection term (optimizer.py:344-358)");
    GEM_TRY(check_stage_args(c, 0, wt_local, params_h, false));
    GEM_TRY(check_stage_args(c, 1, wt_global, params_h, false));
    cudaStream_t s = (cudaStream_t)stream;
    if (W == 0) return GEM_OK;
    GEM_TRY(lbfgs_configure(c, params_h, nullptr, 0));
    StageCall a, b;
    a.which = 0, a.pose0 = pose0_d, a.heat = heat_d, a.frame_base = frame_base_d, a.clip = clip_d;
    a.mean_bone = mean_bone_d, a.eps = eps_d, a.eps_stride = 2 * (size_t)c->n, a.wt = *wt_local, a.p = *params_h;
    a.pose_out = local_pose_d, a.trace = nullptr, a.n_iter = n_iter_d, a.func_evals = func_evals_d, a.status = status_d;
    a.texel_cache = want_texel_cache(c, heat_d, wt_local->reproj, W);
    a.host_maps = a.texel_cache && is_host_memory(heat_d);
    if (a.texel_cache) GEM_TRY(ensure_texel_cache(c));
    b = a;
    b.texel_cache = false, b.host_maps = false;
    b.which = 1, b.pose0 = rel_f32_d, b.heat = nullptr, b.frame_base = nullptr, b.eps = eps_d + c->n, b.wt = *wt_global;
    b.pose_out = global_pose_d, b.status_or = true;      // (status_d keeps the local stage's bits, the global stage ORs its own in)
    b.n_iter = n_iter_d ? n_iter_d + W : nullptr, b.func_evals = func_evals_d ? func_evals_d + W : nullptr;
    const bool graphs = c->use_graphs && !c->prof_on;
    const std::vector<int> w0 = slice_bounds(c, W, a.host_maps, c->gemm_mode >= 1 && !(c->tap_tc[0] && c->tap_tc[1]));
    Fork f;
    GEM_TRY(fork_slices(c, s, w0, graphs, &f));
    const size_t P = (size_t)c->T * c->J * 3;
    for (size_t k = 0; k + 1 < w0.size(); ++k) {
        const int Wk = w0[k + 1] - w0[k];
        cudaStream_t q = f.cs[k];
        GEM_TRY(enqueue_stage_slice(c, q, a, w0[k], Wk, graphs));                      // local stage     optimizer.py:386
        GEM_TRY(timed(c, q, GEM_TAG_TRANSFORM, [&]() {                                  // SLAM transform  optimizer.py:394-398
            return launch_transform(q, Wk, c->T, c->J, local_pose_d + w0[k] * P, 0, cams_d + (size_t)w0[k] * c->T * 16,
                                    rel_f64_d ? rel_f64_d + w0[k] * P : nullptr, rel_f32_d + w0[k] * P, 0);
        }));
        GEM_TRY(enqueue_stage_slice(c, q, b, w0[k], Wk, graphs));                      // global stage    optimizer.py:414
    }
    c->lb_started = true;
    return join_slices(c, s, f);
}

int gem_relative_global(gem_ctx* c, void* stream, int W, const void* pose_d, int pose_is_f64, const double* cams_d,
                        double* out_f64_d, float* out_f32_d) {
    GEM_ENTER(c, W);
    GEM_REQUIRE(pose_d && cams_d && (out_f64_d || out_f32_d), "NULL argument");
    return timed(c, (cudaStream_t)stream, GEM_TAG_TRANSFORM, [&]() {
        return launch_transform((cudaStream_t)stream, W, c->T, c->J, pose_d, pose_is_f64, cams_d, out_f64_d, out_f32_d, 0);
    });
}

int gem_to_global(gem_ctx* c, void* stream, int W, const void* pose_d, int pose_is_f64, const double* cams_d,
                  double* out_f64_d) {
    GEM_ENTER(c, W);
    GEM_REQUIRE(pose_d && cams_d && out_f64_d, "NULL argument");
    return timed(c, (cudaStream_t)stream, GEM_TAG_TRANSFORM, [&]() {
        return launch_transform((cudaStream_t)stream, W, c->T, c->J, pose_d, pose_is_f64, cams_d, out_f64_d, nullptr, 1);
    });
}

int gem_merge_windows(gem_ctx* c, void* stream, int W, int overlap, const double* windows_d, double* out_d) {
    GEM_REQUIRE(c != nullptr && W >= 0, "bad arguments");
    GEM_REQUIRE(windows_d && out_d, "NULL argument");
    GEM_CUDA(cudaSetDevice(c->device));
    return timed(c, (cudaStream_t)stream, GEM_TAG_STITCH,
                 [&]() { return launch_merge((cudaStream_t)stream, W, c->T, overlap, c->J * 3, windows_d, out_d); });
}

int gem_gaussian_smooth(gem_ctx* c, void* stream, int N, int row, double sigma, const double* seq_d, double* out_d) {
    GEM_REQUIRE(c != nullptr && N >= 0 && row > 0, "bad arguments");
    GEM_REQUIRE(seq_d && out_d, "NULL argument");
    GEM_CUDA(cudaSetDevice(c->device));
    return timed(c, (cudaStream_t)stream, GEM_TAG_STITCH,
                 [&]() { return launch_gauss((cudaStream_t)stream, N, row, sigma, seq_d, out_d); });
}

int gem_lift_skeleton(void* stream, int n_frames, int H, int W, int J, const float* heat_d, const double* depth_d,
                      const double* poly_c2w_h, int n_poly, double cx, double cy, int up, int pad_x, double* points_d,
                      float* preds_d, float* maxvals_d, int32_t* argmax_d) {
    GEM_REQUIRE(n_frames >= 0 && H > 0 && W > 0 && up >= 1 && pad_x >= 0, "bad arguments");
    return launch_lift((cudaStream_t)stream, n_frames, H, W, J, heat_d, depth_d, poly_c2w_h, n_poly, cx, cy, up, pad_x,
                       points_d, preds_d, maxvals_d, argmax_d);
}

int gem_pose_align_errors(void* stream, int n_frames, int J, const double* est_d, const double* gt_d,
                          const int32_t* parents_h, const double* bone_len_mm_h, double* aligned_d, double* gt_out_d,
                          double* err_d) {
    GEM_REQUIRE(n_frames >= 0, "bad arguments");
    return launch_pose_align((cudaStream_t)stream, n_frames, J, est_d, gt_d, parents_h, bone_len_mm_h, aligned_d, gt_out_d,
                             err_d);
}

}  // extern "C"
