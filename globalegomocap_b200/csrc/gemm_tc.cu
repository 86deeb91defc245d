// tcgen05 / TMEM / TMA GEMM with fp32-faithful arithmetic (3xTF32) for the VAE's large
// contractions: decoder latent -> T*256 (fused decoder_input + ConvT, SeqConvVAE.py:134-136), its
// bwd-data transpose, and the encoder's fc_mu|fc_var (SeqConvVAE.py:107-108).
//
//   C[M][N] = epi( bias[N] + A[M][K] * B[K][N] ),  A = activations (row-major), B = weights.
//
// Precision: a tensor-core TF32 product keeps 11 significand bits, which would not hold the
// reference's energy to 1e-4 across an L-BFGS run.  Both operands are split as x = hi + lo with
// hi = x with the low 13 mantissa bits cleared (exactly a TF32 value) and lo = x - hi (exact in fp32),
// and three MMAs accumulate  hi*hi + hi*lo + lo*hi  into one fp32 TMEM accumulator; the dropped
// lo*lo term and lo's own truncation are ~2^-22 relative, the level of fp32 accumulation noise.
//
// The tensor core's fp32 accumulation truncates rather than rounds to nearest (measured: the error of a
// single accumulator grows linearly with the number of accumulation steps, 1.5e-5 relative at
// K = 2560), so the K blocks are dealt round-robin to four TMEM accumulators that the epilogue adds
// with ordinary round-to-nearest fp32 adds: a quarter of the steps, a quarter of the bias.
//
// Kernel anatomy (one 128x128 output tile per CTA, 192 threads):
//   warp 0   TMA producer: per 32-wide K block four cp.async.bulk.tensor loads (A_hi, A_lo, B_hi, B_lo;
//            128 rows x 128 B each, SWIZZLE_128B) into a 3-stage shared-memory ring, mbarrier full/empty
//   warp 1   allocates all 512 TMEM columns (4 accumulators); one elected lane issues 12 tcgen05.mma.kind::tf32 (M128 N128 K8)
//            per stage and tcgen05.commit's the stage back to the producer
//   warps 2-5 epilogue: tcgen05.ld the accumulator (32 lanes x 32 columns per instruction), bias /
//            LeakyReLU, 128-byte row segments straight to global memory
// Weights are transposed to K-major [N][K] and split once per checkpoint; activations are split by a
// small elementwise kernel before each GEMM.
#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"

namespace gem {

using namespace tc;

namespace {

constexpr int BM = 128, BN = 128, BK = 32;        // BK fp32 = 128 bytes = one swizzle row
constexpr int kStages = 3;
constexpr int kTileBytes = BM * BK * 4;           // 16 KB
constexpr int kStageBytes = 4 * kTileBytes;       // A_hi, A_lo, B_hi, B_lo
constexpr int kAcc = 4;                           // independent fp32 accumulators (see below)
constexpr int kTmemCols = kAcc * BN;              // 512 = all of TMEM
constexpr int kThreads = 192;
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

constexpr uint32_t kInstrDesc = instr_desc_tf32(BM, BN);

struct TcArgs {
    const float* bias;
    float* C;        // result, or its TF32 hi part when C_lo is set
    float* C_lo;     // optional: the epilogue writes the result already split for the next tensor-core layer
    uint32_t* sign;  // optional: packed sign bits of the result, [M][N/32]
    int M, N, K, ldc, epi;
};

__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, TcArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int num_kb = g.K / BK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a_hi), prefetch_tmap(&map_a_lo), prefetch_tmap(&map_b_hi), prefetch_tmap(&map_b_lo);
        for (int s = 0; s < kStages; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t parity = ((kb / kStages) & 1) ^ 1;
                mbar_wait(&empty_bar[s], parity);
                uint8_t* st = smem + s * kStageBytes;
                mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
                tma_load_2d(st + 0 * kTileBytes, &map_a_hi, kb * BK, m0, &full_bar[s]);
                tma_load_2d(st + 1 * kTileBytes, &map_a_lo, kb * BK, m0, &full_bar[s]);
                tma_load_2d(st + 2 * kTileBytes, &map_b_hi, kb * BK, n0, &full_bar[s]);
                tma_load_2d(st + 3 * kTileBytes, &map_b_lo, kb * BK, n0, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                mbar_wait(&full_bar[s], (kb / kStages) & 1);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + s * kStageBytes);
                const uint64_t a_hi = make_smem_desc(base + 0 * kTileBytes), a_lo = make_smem_desc(base + 1 * kTileBytes);
                const uint64_t b_hi = make_smem_desc(base + 2 * kTileBytes), b_lo = make_smem_desc(base + 3 * kTileBytes);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    const uint64_t adv = (uint64_t)((k * 8 * 4) >> 4);     // 32 bytes per K=8 step inside the 128-B row
                    // small terms first, then the dominant one
                    const uint32_t acc = tmem_base + (uint32_t)((kb % kAcc) * BN);
                    umma_tf32(acc, a_lo + adv, b_hi + adv, kInstrDesc, (kb >= kAcc) || (k != 0));
                    umma_tf32(acc, a_hi + adv, b_lo + adv, kInstrDesc, 1);
                    umma_tf32(acc, a_hi + adv, b_hi + adv, kInstrDesc, 1);
                }
                umma_commit(&empty_bar[s]);          // frees the smem stage when these MMAs have read it
            }
            umma_commit(tmem_full_bar);              // accumulator complete
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int m = m0 + q * 32 + lane;
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
            tmem_ld_32x32b_x32(taddr, v);
            tmem_ld_wait();
            const int nacc = num_kb < kAcc ? num_kb : kAcc;
            for (int a = 1; a < nacc; ++a) {
                uint32_t t[32];
                tmem_ld_32x32b_x32(taddr + (uint32_t)(a * BN), t);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(t[j]));
            }
            if (m < g.M) {
                const int nb = n0 + c * 32;
                uint32_t sbits = 0;
                float* dst = g.C + (size_t)m * g.ldc + nb;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    float4 o;
                    float* po = reinterpret_cast<float*>(&o);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float x = __uint_as_float(v[j + e]);
                        if (g.bias) x += __ldg(g.bias + nb + j + e);
                        if (g.epi == EPI_LRELU) x = x > 0.f ? x : x * 0.01f;
                        po[e] = x;
                        sbits |= (x > 0.f ? 1u : 0u) << (j + e);
                    }
                    if (g.C_lo) {
                        float4 l;
                        split_tf32(o.x, o.x, l.x), split_tf32(o.y, o.y, l.y), split_tf32(o.z, o.z, l.z), split_tf32(o.w, o.w, l.w);
                        *reinterpret_cast<float4*>(g.C_lo + (size_t)m * g.ldc + nb + j) = l;
                    }
                    *reinterpret_cast<float4*>(dst + j) = o;
                }
                if (g.sign) g.sign[(size_t)m * (g.N >> 5) + (nb >> 5)] = sbits;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// hi = x with the 13 low mantissa bits cleared (an exact TF32 value), lo = x - hi
__global__ void split_tf32_kernel(const float* __restrict__ A, int lda, int M, int K, float* __restrict__ hi,
                                  float* __restrict__ lo) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one float4 per thread
    const int kv = K / 4;
    if (i >= (size_t)M * kv) return;
    const int m = (int)(i / kv), k = (int)(i - (size_t)m * kv) * 4;
    const float4 x = *reinterpret_cast<const float4*>(A + (size_t)m * lda + k);
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u), l.x = x.x - h.x;
    h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u), l.y = x.y - h.y;
    h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u), l.z = x.z - h.z;
    h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u), l.w = x.w - h.w;
    *reinterpret_cast<float4*>(hi + (size_t)m * K + k) = h;
    *reinterpret_cast<float4*>(lo + (size_t)m * K + k) = l;
}

// weights [K][ldb] -> K-major [N][K] hi / lo (once per checkpoint)
__global__ void transpose_split_kernel(const float* __restrict__ B, int ldb, int K, int N, float* __restrict__ hi,
                                       float* __restrict__ lo) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int k = k0 + r, n = n0 + threadIdx.x;
        tile[r][threadIdx.x] = (k < K && n < N) ? B[(size_t)k * ldb + n] : 0.f;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int n = n0 + r, k = k0 + threadIdx.x;
        if (n < N && k < K) {
            const float x = tile[threadIdx.x][r];
            const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
            hi[(size_t)n * K + k] = h;
            lo[(size_t)n * K + k] = x - h;
        }
    }
}

// ---------------------------------------------------------------- host side
// rows x K fp32 row-major (pitch K), box = 128 rows x 32 floats, 128-byte swizzle, OOB rows read zero
int make_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t K, uint64_t pitch = 0) {
    const uint64_t dims[2] = {K, rows};
    const uint64_t strides[1] = {(pitch ? pitch : K) * sizeof(float)};
    const uint32_t box[2] = {(uint32_t)BK, (uint32_t)BM};
    return make_map_f32(map, base, 2, dims, strides, box);
}

struct WeightSplit {
    float *hi = nullptr, *lo = nullptr;
    int K = 0, N = 0;
    CUtensorMap map_hi, map_lo;
};
struct TcState {
    std::unordered_map<const float*, WeightSplit> weights;   // keyed by the caller's weight pointer
    float *a_hi = nullptr, *a_lo = nullptr;
    size_t a_capacity = 0;                                    // floats in each of a_hi / a_lo
    bool attr_set = false;
};
std::mutex g_mu;
std::unordered_map<void*, TcState*> g_states;                // one per workspace owner (ctx)

}  // namespace

namespace tc {

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

int make_map_f32(CUtensorMap* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return GEM_ERR_CUDA;
    }
    cuuint64_t d[5], st[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; ++i) d[i] = dims[i], bx[i] = box[i], estr[i] = 1;
    for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), d, st, bx, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return GEM_ERR_CUDA;
    }
    return GEM_OK;
}

}  // namespace tc

bool tc_gemm_available() { return true; }

// `owner` identifies the ctx; scratch grows on demand and lives until tc_gemm_release(owner).
int launch_tap_gemm_tc(cudaStream_t stream, const TapGemmArgs& g, void* owner, size_t) {
    if (g.M <= 0) return GEM_OK;
    GEM_REQUIRE(g.taps == 1, "tcgen05 path handles plain GEMMs only");
    GEM_REQUIRE(g.K % BK == 0 && g.N % BN == 0, "K must be a multiple of 32 and N of 128 for the tcgen05 path");
    GEM_REQUIRE(g.lda % 4 == 0 && g.ldc % 4 == 0, "lda/ldc must be multiples of 4");
    GEM_REQUIRE(g.epi == EPI_NONE || g.epi == EPI_LRELU, "unsupported epilogue on the tcgen05 path");
    TcState* st;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_states.find(owner);
        if (it == g_states.end()) it = g_states.emplace(owner, new TcState()).first;
        st = it->second;
    }
    if (!st->attr_set) {
        GEM_CUDA(cudaFuncSetAttribute(tc_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        st->attr_set = true;
    }
    // weights: K-major hi/lo copies registered by tc_gemm_prepare_weight
    auto wit = st->weights.find(g.B);
    if (wit == st->weights.end() || wit->second.K != g.K || wit->second.N != g.N) {
        set_error("tcgen05 path: weight matrix was not prepared (tc_gemm_prepare_weight)");
        return GEM_ERR_STATE;
    }
    // activations: already split by the producing kernel, or split here into the ctx-owned scratch
    const float *a_hi = g.A_hi, *a_lo = g.A_lo;
    uint64_t pitch = (uint64_t)g.lda;
    if (!a_hi) {
        const size_t need = (size_t)g.M * g.K;
        if (need > st->a_capacity) {
            GEM_CUDA(cudaStreamSynchronize(stream));
            if (st->a_hi) cudaFree(st->a_hi), cudaFree(st->a_lo);
            GEM_CUDA(cudaMalloc(&st->a_hi, need * sizeof(float)));
            GEM_CUDA(cudaMalloc(&st->a_lo, need * sizeof(float)));
            st->a_capacity = need;
        }
        const size_t n4 = (size_t)g.M * (g.K / 4);
        split_tf32_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(g.A, g.lda, g.M, g.K, st->a_hi, st->a_lo);
        GEM_CHECK_LAUNCH();
        a_hi = st->a_hi, a_lo = st->a_lo, pitch = (uint64_t)g.K;
    }
    CUtensorMap map_a_hi, map_a_lo;
    int rc = make_map(&map_a_hi, a_hi, (uint64_t)g.M, (uint64_t)g.K, pitch);
    if (rc == GEM_OK) rc = make_map(&map_a_lo, a_lo, (uint64_t)g.M, (uint64_t)g.K, pitch);
    if (rc != GEM_OK) return rc;
    TcArgs a;
    a.bias = g.bias, a.C = g.C, a.C_lo = g.C_lo, a.sign = g.C_sign, a.M = g.M, a.N = g.N, a.K = g.K, a.ldc = g.ldc, a.epi = g.epi;
    dim3 grid((g.M + BM - 1) / BM, g.N / BN);
    tc_gemm_kernel<<<grid, kThreads, kSmemBytes, stream>>>(map_a_hi, map_a_lo, wit->second.map_hi, wit->second.map_lo, a);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

// x -> (hi, lo) TF32 parts for M rows of K floats (K % 4 == 0; output pitch K)
int launch_split_tf32(cudaStream_t stream, const float* A, int lda, int M, int K, float* hi, float* lo) {
    if (M <= 0) return GEM_OK;
    GEM_REQUIRE(K % 4 == 0 && lda % 4 == 0, "K and lda must be multiples of 4");
    const size_t n4 = (size_t)M * (K / 4);
    split_tf32_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(A, lda, M, K, hi, lo);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

// Transposes B [K][ldb] to K-major [N][K] and splits it into TF32 hi / lo parts; replaces any earlier
// registration of the same pointer (the caller may have refilled the buffer).
int tc_gemm_prepare_weight(void* owner, cudaStream_t stream, const float* B, int ldb, int K, int N) {
    GEM_REQUIRE(K % BK == 0 && N % BN == 0, "K must be a multiple of 32 and N of 128 for the tcgen05 path");
    TcState* st;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_states.find(owner);
        if (it == g_states.end()) it = g_states.emplace(owner, new TcState()).first;
        st = it->second;
    }
    auto old = st->weights.find(B);
    if (old != st->weights.end()) {
        GEM_CUDA(cudaStreamSynchronize(stream));
        cudaFree(old->second.hi), cudaFree(old->second.lo);
        st->weights.erase(old);
    }
    WeightSplit ws;
    ws.K = K, ws.N = N;
    GEM_CUDA(cudaMalloc(&ws.hi, (size_t)N * K * sizeof(float)));
    GEM_CUDA(cudaMalloc(&ws.lo, (size_t)N * K * sizeof(float)));
    dim3 grid((K + 31) / 32, (N + 31) / 32), block(32, 8);
    transpose_split_kernel<<<grid, block, 0, stream>>>(B, ldb, K, N, ws.hi, ws.lo);
    GEM_CHECK_LAUNCH();
    int rc = make_map(&ws.map_hi, ws.hi, (uint64_t)N, (uint64_t)K);
    if (rc == GEM_OK) rc = make_map(&ws.map_lo, ws.lo, (uint64_t)N, (uint64_t)K);
    if (rc != GEM_OK) return rc;
    st->weights.emplace(B, ws);
    return GEM_OK;
}

void tc_gemm_release(void* owner) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_states.find(owner);
    if (it == g_states.end()) return;
    TcState* st = it->second;
    for (auto& kv : st->weights) cudaFree(kv.second.hi), cudaFree(kv.second.lo);
    if (st->a_hi) cudaFree(st->a_hi), cudaFree(st->a_lo);
    delete st;
    g_states.erase(it);
}

}  // namespace gem
