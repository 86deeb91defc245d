// tcgen05 / TMEM / TMA kernel for the VAE's k=3 convolutions (Conv1d / ConvTranspose1d, stride 1,
// pad 1, SeqConvVAE.py:27-92) and their bwd-data, in fp32-faithful 3xTF32 arithmetic.
//
//   out[m][n] = epi( bias[n] + sum_{tap<3} sum_k A[m + tap - 1][k] * B[tap][k][n] ),   m = (window, frame) token,
//   taps that leave the window's T frames read zeros.
//
// Instead of shifting the activation rows per tap (which a swizzled shared-memory operand cannot do), the
// three taps sit side by side in the accumulator:  P[m][tap*N + n] = sum_k A[m][k] * B[tap][k][n]  is ONE
// GEMM with 3N output columns, and the convolution is finished in the epilogue:
//   out[m] = P_0[m-1] + P_1[m] + P_2[m+1]   (rows of the same window only).
// An M tile is 128 TMEM lanes = 4 warps' lane quarters; each quarter holds floor(32/T) whole windows
// (T = 10: 3 windows, 30 rows, 2 idle), so the +-1 row shift is a warp shuffle and never crosses a warp.
// The activation tile is fetched by 3-D TMA boxes {32 channels, T frames, 3 windows} from the token-major
// [W][T][C] tensor; windows beyond W are zero-filled by the TMA unit.
//
// 3xTF32: x = hi + lo (hi = top 19 bits, exact TF32; lo = x - hi); hi*hi + hi*lo + lo*hi accumulate in one
// fp32 TMEM accumulator (K <= 256 here: few enough accumulation steps that the tensor core's truncating
// adds stay at the 1e-6 level, see gemm_tc.cu).  Activations travel between layers already split: the
// epilogue of the producing kernel writes hi and lo, so no separate split pass touches HBM.
//
// Second scheme, same kernel (template parameter F16), twice the tensor rate and half the operand bytes:
// activations x ~ hi + 2^-11 lo as two fp16 arrays (tc_common.cuh: split_f16); weights pre-multiplied by 2^8 and
// stored as THREE fp16 slabs  hi' = fp16(256 w),  lo = fp16(256 w - hi')  (unscaled: normal numbers thanks to the
// 2^8)  and  hs = hi' 2^-11,  so that  hi_a hi'_b + hi_a lo_b + lo_a hs_b  all carry the same scale and share ONE
// accumulator (a second accumulator for scaled cross terms would not fit two CTAs' TMEM); the epilogue multiplies
// by 2^-8.  A 128-byte row holds 64 fp16: the K = 64 layers are a single K block.
//
// CTA anatomy (320 threads, one stage of shared memory; two CTAs share an SM and overlap each other's
// load / MMA / epilogue phases): warp 0 = TMA producer, warp 1 = TMEM allocation + single-thread
// tcgen05.mma issue (M128, N = 3*NCTA, K8, kind::tf32), warps 2-9 = epilogue, two per TMEM lane quarter
// (tcgen05.ld, shuffle shift-add, bias / LeakyReLU / LeakyReLU-derivative mask, hi/lo split, staging in the
// idle operand memory, 3-D TMA store).
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>
#include <unordered_map>

#include "energy_device.cuh"
#include "tc_common.cuh"

namespace gem {

using namespace tc;

namespace {

constexpr int kRows = 128;                 // UMMA M
constexpr int kATile = kRows * 128;        // 16 KB: 128 rows of one 128-byte swizzle row (32 fp32 or 64 fp16)
constexpr int kQuarterBytes = 32 * 128;    // 4 KB: one warp's 32 rows
constexpr float kWScale16 = 256.f;         // fp16 scheme: weights are stored times 2^8
constexpr int kThreads = 320;              // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr int kTmemCols = 256;

template <int NCTA, bool F16>
struct Cfg {
    static constexpr int NP = 3 * NCTA;                       // accumulator columns: the taps side by side
    static constexpr int BK = F16 ? 64 : 32;                  // elements per 128-byte row
    static constexpr int kBTile = NP * 128;
    static constexpr int kNB = F16 ? 3 : 2;                   // weight slabs per K block
    static constexpr int kStage = 2 * kATile + kNB * kBTile;  // A_hi, A_lo, B slabs
    static constexpr size_t kSmem = (size_t)kStage + 1024 /*align slack*/ + 128 /*barriers*/;
    static_assert(NP % 16 == 0 && NP <= 256, "UMMA N");
    static_assert(kBTile % 1024 == 0, "B tile must keep the swizzle atom alignment");
};

struct TapTcArgs {
    const float* bias;   // [N] or NULL
    const float* aux;    // [tokens][ldaux]: sign of the saved activation (EPI_MASK) ...
    const uint32_t* aux_bits;   // ... or, preferred, its packed sign bits [tokens][N/32]
    uint32_t* sign_out;  // optional: packed sign bits of this layer's output [tokens][N/32]
    float* out_hi;       // [tokens][ldo]: hi part (fp32 TF32 value / fp16), or the plain fp32 result when out_lo is NULL
    void* out_lo;
    int W, T, wpq;       // windows, frames per window, windows per 32-lane quarter
    int N, ldo, ldaux, epi, num_kb;
    long long* dbg;      // optional per-CTA phase timestamps (SM clock), 16 slots per CTA; NULL in production
};
#define TAP_DBG(slot)                                                                                      \
    do {                                                                                                   \
        if (g.dbg) g.dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = clock64();          \
    } while (0)

template <int NCTA, bool F16>
__global__ void __launch_bounds__(kThreads)
tc_tap_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
              const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
              const __grid_constant__ CUtensorMap map_b_hs,
              const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo, TapTcArgs g) {
    using C = Cfg<NCTA, F16>;
    constexpr int BK = C::BK;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sA_hi = smem;
    uint8_t* sA_lo = smem + kATile;
    uint8_t* sB_hi = smem + 2 * kATile;
    uint8_t* sB_lo = sB_hi + C::kBTile;
    uint8_t* sB_hs = sB_lo + C::kBTile;              // fp16 scheme only
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kStage);
    uint64_t* empty_bar = full_bar + 1;
    uint64_t* tmem_full_bar = full_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(full_bar + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) TAP_DBG(0);
    if (threadIdx.x == 0 && g.dbg) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g.dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 14] = (long long)gt;
    }
    const int win0 = blockIdx.x * 4 * g.wpq;      // first window of this M tile
    const int y = blockIdx.y;                      // output-channel slab of NCTA channels
    const int rows_q = g.wpq * g.T;                // rows in use per quarter

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a_hi), prefetch_tmap(&map_a_lo), prefetch_tmap(&map_b_hi), prefetch_tmap(&map_b_lo);
        mbar_init(full_bar, 1), mbar_init(empty_bar, 1), mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    // the idle rows of each quarter are never written by TMA: clear them once so the MMA reads zeros
    {
        const int idle = 32 - rows_q;                          // rows per quarter
        const int total = 2 * 4 * idle * 8;                    // float4 slots in both A tiles
        for (int i = threadIdx.x; i < total; i += kThreads) {
            const int tile = i / (4 * idle * 8), r = i % (4 * idle * 8);
            const int q = r / (idle * 8), rr = (r % (idle * 8)) / 8, c16 = r % 8;
            uint8_t* p = (tile ? sA_lo : sA_hi) + q * kQuarterBytes + (rows_q + rr) * 128 + c16 * 16;
            *reinterpret_cast<float4*>(p) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TAP_DBG(1);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            const uint32_t box_bytes = (uint32_t)rows_q * 128;
            for (int kb = 0; kb < g.num_kb; ++kb) {
                mbar_wait(empty_bar, (uint32_t)((kb & 1) ^ 1));
                mbar_arrive_expect_tx(full_bar, 8 * box_bytes + C::kNB * C::kBTile);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    tma_load_3d(sA_hi + q * kQuarterBytes, &map_a_hi, kb * BK, 0, win0 + q * g.wpq, full_bar);
                    tma_load_3d(sA_lo + q * kQuarterBytes, &map_a_lo, kb * BK, 0, win0 + q * g.wpq, full_bar);
                }
                tma_load_2d(sB_hi, &map_b_hi, kb * BK, y * C::NP, full_bar);
                tma_load_2d(sB_lo, &map_b_lo, kb * BK, y * C::NP, full_bar);
                if (F16) tma_load_2d(sB_hs, &map_b_hs, kb * BK, y * C::NP, full_bar);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = F16 ? instr_desc_f16(kRows, C::NP) : instr_desc_tf32(kRows, C::NP);
            const uint64_t a_hi = make_smem_desc(smem_u32(sA_hi)), a_lo = make_smem_desc(smem_u32(sA_lo));
            const uint64_t b_hi = make_smem_desc(smem_u32(sB_hi)), b_lo = make_smem_desc(smem_u32(sB_lo));
            const uint64_t b_hs = make_smem_desc(smem_u32(sB_hs));
            for (int kb = 0; kb < g.num_kb; ++kb) {
                mbar_wait(full_bar, (uint32_t)(kb & 1));
                tc_fence_after();
                if (kb < 2) TAP_DBG(2 + kb);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);        // 32 bytes per MMA K step inside the 128-B row
                    const uint32_t first = (kb > 0) || (k != 0);
                    if (F16) {                                             // small terms first
                        umma_f16(tmem_base, a_lo + adv, b_hs + adv, idesc, first);
                        umma_f16(tmem_base, a_hi + adv, b_lo + adv, idesc, 1);
                        umma_f16(tmem_base, a_hi + adv, b_hi + adv, idesc, 1);
                    } else {
                        umma_tf32(tmem_base, a_lo + adv, b_hi + adv, idesc, first);
                        umma_tf32(tmem_base, a_hi + adv, b_lo + adv, idesc, 1);
                        umma_tf32(tmem_base, a_hi + adv, b_hi + adv, idesc, 1);
                    }
                }
                umma_commit(empty_bar);              // the stage may be refilled once these MMAs have read it
            }
            umma_commit(tmem_full_bar);
        }
    } else {
        // ===== epilogue warps 2..5: TMEM lane quarter = warp % 4 =====
        // All MMAs have completed when tmem_full_bar fires, so the operand stage is free: the result tile
        // is staged there (split outputs in the 128-byte-swizzled layout a 3-D TMA store expects, which is
        // also the next layer's operand layout) and leaves by TMA / coalesced rows.  Per-thread row-segment
        // stores would cost 32 L1 wavefronts per instruction.
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;                       // the quarter's two warps split the column chunks
        constexpr int kChunks = NCTA / 16;
        const int c_begin = half ? (kChunks + 1) / 2 : 0, c_end = half ? kChunks : (kChunks + 1) / 2;
        const int wl = lane / g.T, t = lane - wl * g.T;         // window inside the quarter, frame
        const int winq = win0 + q * g.wpq;                      // first window of this quarter
        const int win = winq + wl;
        const bool row_ok = lane < rows_q && win < g.W;
        const bool has_prev = t > 0, has_next = t < g.T - 1;
        const size_t token = (size_t)win * g.T + t;
        const int words = g.N >> 5;                             // sign words per token
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        if (threadIdx.x == 64) TAP_DBG(4);
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
        uint32_t sbits = 0, mbits = 0;
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
            uint32_t p0[16], p1[16], p2[16];
            tmem_ld_32x32b_x16(trow + (uint32_t)(0 * NCTA + c * 16), p0);
            tmem_ld_32x32b_x16(trow + (uint32_t)(1 * NCTA + c * 16), p1);
            tmem_ld_32x32b_x16(trow + (uint32_t)(2 * NCTA + c * 16), p2);
            const int nb = y * NCTA + c * 16;
            const int sh = (c & 1) * 16;
            if (g.epi == EPI_MASK && g.aux_bits && sh == 0 && row_ok) mbits = __ldg(g.aux_bits + token * words + (nb >> 5));
            tmem_ld_wait();
            if (threadIdx.x == 64 && c < 4) TAP_DBG(8 + c);
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(p0[j]), 1);     // P_0 of the previous frame
                const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(p2[j]), 1);   // P_2 of the next frame
                float v = __uint_as_float(p1[j]);
                v += has_prev ? up : 0.f;
                v += has_next ? dn : 0.f;
                o[j] = F16 ? v * (1.f / kWScale16) : v;          // fp16 scheme: the weights carry a factor 2^8
            }
            if (g.bias) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (nb + j < g.N) o[j] += __ldg(g.bias + nb + j);
            }
            if (g.epi == EPI_LRELU) {
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] = o[j] > 0.f ? o[j] : o[j] * 0.01f;
            } else if (g.epi == EPI_MASK) {
                if (g.aux_bits) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) o[j] = ((mbits >> (sh + j)) & 1u) ? o[j] : o[j] * 0.01f;
                } else if (row_ok) {
                    const float* ax = g.aux + token * g.ldaux + nb;
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 s4 = *reinterpret_cast<const float4*>(ax + j);
                        o[j + 0] = s4.x > 0.f ? o[j + 0] : o[j + 0] * 0.01f;
                        o[j + 1] = s4.y > 0.f ? o[j + 1] : o[j + 1] * 0.01f;
                        o[j + 2] = s4.z > 0.f ? o[j + 2] : o[j + 2] * 0.01f;
                        o[j + 3] = s4.w > 0.f ? o[j + 3] : o[j + 3] * 0.01f;
                    }
                }
            }
            if (g.sign_out) {                     // LeakyReLU keeps the sign: one bit per channel for the bwd-data mask
#pragma unroll
                for (int j = 0; j < 16; ++j) sbits |= (o[j] > 0.f ? 1u : 0u) << (sh + j);
                if (sh == 16) {
                    if (row_ok) g.sign_out[token * words + (nb >> 5)] = sbits;
                    sbits = 0;
                }
            }
            if (g.out_lo && F16) {
                // one 64-channel block of fp16: 128-byte rows, this chunk is two 16-byte pieces of them
                uint8_t* th = smem + q * kQuarterBytes + lane * 128;
                uint8_t* tl = th + kATile;
#pragma unroll
                for (int j = 0; j < 16; j += 8) {
                    uint16_t h[8], l[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) split_f16(o[j + e], h[e], l[e]);
                    const int off = (((c * 2 + (j >> 3)) ^ (lane & 7)) << 4);            // SWIZZLE_128B
                    *reinterpret_cast<uint4*>(th + off) =
                        make_uint4((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16),
                                   (uint32_t)h[4] | ((uint32_t)h[5] << 16), (uint32_t)h[6] | ((uint32_t)h[7] << 16));
                    *reinterpret_cast<uint4*>(tl + off) =
                        make_uint4((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16),
                                   (uint32_t)l[4] | ((uint32_t)l[5] << 16), (uint32_t)l[6] | ((uint32_t)l[7] << 16));
                }
            } else if (g.out_lo) {
                uint8_t* th = smem + ((c >> 1) * 2 + 0) * kATile + q * kQuarterBytes + lane * 128;
                uint8_t* tl = th + kATile;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 h, l;
                    split_tf32(o[j + 0], h.x, l.x), split_tf32(o[j + 1], h.y, l.y);
                    split_tf32(o[j + 2], h.z, l.z), split_tf32(o[j + 3], h.w, l.w);
                    const int off = ((((c & 1) * 4 + (j >> 2)) ^ (lane & 7)) << 4);     // SWIZZLE_128B
                    *reinterpret_cast<float4*>(th + off) = h;
                    *reinterpret_cast<float4*>(tl + off) = l;
                }
            } else {
                float* sp = reinterpret_cast<float*>(smem + q * 2 * kQuarterBytes) + lane * g.N + c * 16;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (nb + j < g.N) sp[j] = o[j];
            }
        }
        if (threadIdx.x == 64) TAP_DBG(7);
        if (g.out_lo && F16) {
            // the quarter's two warps fill one 64-channel block together: meet, then one of them stores it
            fence_proxy_async_smem();
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            if (half == 0 && lane == 0 && winq < g.W) {
                tma_store_3d(&map_o_hi, smem + q * kQuarterBytes, y * NCTA, 0, winq);
                tma_store_3d(&map_o_lo, smem + kATile + q * kQuarterBytes, y * NCTA, 0, winq);
                bulk_commit();
                bulk_wait_read0();
            }
        } else if (g.out_lo) {
            fence_proxy_async_smem();             // generic-proxy writes -> visible to the TMA store
            __syncwarp();
            if (lane == 0 && winq < g.W) {
                for (int b = c_begin >> 1; b < (c_end + 1) >> 1; ++b) {       // this warp's 32-channel blocks
                    tma_store_3d(&map_o_hi, smem + (b * 2 + 0) * kATile + q * kQuarterBytes, y * NCTA + b * 32, 0, winq);
                    tma_store_3d(&map_o_lo, smem + (b * 2 + 1) * kATile + q * kQuarterBytes, y * NCTA + b * 32, 0, winq);
                }
                bulk_commit();
                bulk_wait_read0();                // the staging memory must outlive the engine's reads, not its writes
            }
        } else {
            // plain output (the pose): the quarter's windows are consecutive, dense rows in global memory;
            // the quarter's two warps meet on a named barrier and share the copy
            asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
            int nwin = g.W - winq;
            nwin = nwin < 0 ? 0 : (nwin > g.wpq ? g.wpq : nwin);
            const int count = nwin * g.T * g.N;
            const float* sp = reinterpret_cast<const float*>(smem + q * 2 * kQuarterBytes);
            float* dp = g.out_hi + (size_t)winq * g.T * g.N;
            for (int i = half * 32 + lane; i < count; i += 64) dp[i] = sp[i];
        }
    }
    if (threadIdx.x == 64) TAP_DBG(5);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
    if (threadIdx.x == 0) TAP_DBG(6);
    if (threadIdx.x == 0 && g.dbg) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g.dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + 15] = (long long)gt;
    }
}

// weights [3][K][ldb] (tap, in, out) -> K-major slabs [gridy][3][NCTA][Kp], split into TF32 hi / lo
__global__ void tap_weight_prep_kernel(const float* __restrict__ B, int ldb, int K, int N, int Kp, int ncta, int gridy,
                                       float* __restrict__ hi, float* __restrict__ lo) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)gridy * 3 * ncta * Kp;
    if (i >= total) return;
    const int k = (int)(i % Kp);
    const size_t row = i / Kp;
    const int np = (int)(row % ncta), tap = (int)((row / ncta) % 3), yy = (int)(row / (3 * (size_t)ncta));
    const int n = yy * ncta + np;
    float x = 0.f;
    if (k < K && n < N) x = B[((size_t)tap * K + k) * ldb + n];
    float h, l;
    split_tf32(x, h, l);
    hi[i] = h, lo[i] = l;
}

// the fp16 scheme's three slabs: hi' = fp16(2^8 w), lo = fp16(2^8 w - hi'), hs = hi' 2^-11
__global__ void tap_weight_prep_f16_kernel(const float* __restrict__ B, int ldb, int K, int N, int Kp, int ncta, int gridy,
                                           uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, uint16_t* __restrict__ hs) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)gridy * 3 * ncta * Kp;
    if (i >= total) return;
    const int k = (int)(i % Kp);
    const size_t row = i / Kp;
    const int np = (int)(row % ncta), tap = (int)((row / ncta) % 3), yy = (int)(row / (3 * (size_t)ncta));
    const int n = yy * ncta + np;
    float x = 0.f;
    if (k < K && n < N) x = B[((size_t)tap * K + k) * ldb + n] * kWScale16;
    uint16_t h, l, s;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
    float hf;
    asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(h));
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l) : "f"(x - hf));
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(s) : "f"(hf * (1.f / kF16LoScale)));
    hi[i] = h, lo[i] = l, hs[i] = s;
}

// [tokens][C] -> zero-padded [tokens][ldo] hi / lo (the bwd-data entry: d pose with 45 -> 48 columns)
__global__ void split_pad_kernel(const float* __restrict__ src, int C, size_t tokens, int ldo, float* __restrict__ hi,
                                 float* __restrict__ lo) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= tokens * (size_t)ldo) return;
    const size_t m = i / ldo;
    const int c = (int)(i - m * ldo);
    float h = 0.f, l = 0.f;
    if (c < C) split_tf32(src[m * C + c], h, l);
    hi[i] = h, lo[i] = l;
}

// fp16 scheme: one CTA per window; the window's gradient [T*C] is scaled by the power of two that brings its
// largest entry to ~2^4, split into fp16 hi / lo and zero-padded to [T][ldo]; the exponent goes to row_exp[w]
__global__ void __launch_bounds__(128) rowscale_split_pad_f16_kernel(const float* __restrict__ src, int C, int T, int ldo,
                                                                     uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                                     int32_t* __restrict__ row_exp) {
    __shared__ float red[4];
    const size_t w = blockIdx.x;
    const float* p = src + w * T * C;
    float mx = 0.f;
    for (int i = threadIdx.x; i < T * C; i += 128) mx = fmaxf(mx, fabsf(p[i]));
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    int e = 0;
    if (mx > 0.f && mx < 3.0e38f) {
        e = 4 - ilogbf(mx);
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
    }
    const float sc = exp2f((float)e);
    if (threadIdx.x == 0) row_exp[w] = e;
    for (int i = threadIdx.x; i < T * ldo; i += 128) {
        const int t = i / ldo, c = i - t * ldo;
        uint16_t h = 0, l = 0;
        if (c < C) split_f16(p[t * C + c] * sc, h, l);
        hi[w * T * ldo + i] = h, lo[w * T * ldo + i] = l;
    }
}


// ---- chain of k=3 layers in one launch (fp16 scheme) ---------------------------------------------------------------
// A K = 64 layer is latency, not work: 0.7 us of setup, 1.3 us until the first tile lands, 0.5 us of MMAs, an
// epilogue, then the launch gap — per layer, ten times per closure round.  The epilogue already stages its result in
// the 128-byte-swizzled layout the next layer's MMA reads, so a chain of layers can keep the activation tile in
// shared memory: one setup, one activation load, and only sign bits (the bwd-data masks) plus the last layer's
// output go to HBM.  Structure per CTA (12 windows = one M tile, 576 threads, one CTA per SM):
//   warp 0    TMA: the first layer's activation tile (3-D boxes per TMEM lane quarter), then every layer's weight
//             blocks (three fp16 slabs of 3*NCTA rows per K block) through a two-slot ring (full / empty mbarriers)
//   warp 1    one thread issues the MMAs of layer l as soon as its operand tile is ready (act_ready: the previous
//             layer's epilogue warps arrive after a proxy fence) and commits to acc_full
//   warps 2-17  epilogue, four per TMEM lane quarter (one 16-column chunk each): shuffle shift-add of the three
//             taps, bias / LeakyReLU / mask, sign bits, fp16 hi/lo split into the other activation buffer
// Accumulators alternate between TMEM columns [0, 192) and [192, 384); activation buffers alternate likewise.
constexpr int kChainThreads = 64 + 16 * 32;
constexpr int kChainEpiWarps = 16;
constexpr int kChainWSlot = 3 * 192 * 128;                    // 73,728 B: three slabs of 192 rows x 128 B
constexpr size_t kChainSmem = 4 * kATile + 2 * kChainWSlot + 1024 /*align slack*/ + 256 /*barriers*/;

struct ChainLayer {
    const float* bias;
    const uint32_t* aux_bits;
    uint32_t* sign_out;
    int ncta, gridy, num_kb, N, epi;
    int out_kind;            // 0: operand of the next layer (shared memory only), 1: split fp16 to global (TMA store),
};                           // 2: plain fp32 rows to global
struct ChainMaps {
    CUtensorMap a_hi, a_lo;
    CUtensorMap w[kTapChainMax][3];
    CUtensorMap o_hi, o_lo;
};
struct ChainArgs {
    ChainLayer L[kTapChainMax];
    int nl, W, T, wpq;
    float* out_plain;
    uint32_t* status;            // optional [W]: GEM_WIN_F16_RANGE is OR-ed in when a split activation saturates
    long long* dbg;
};

__global__ void __launch_bounds__(kChainThreads, 1)
tc_tap_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* act = smem;                                   // two buffers of (hi 16 KB, lo 16 KB)
    uint8_t* wring = smem + 4 * kATile;                    // two weight slots
    uint64_t* bars = reinterpret_cast<uint64_t*>(wring + 2 * kChainWSlot);
    uint64_t* act_full = bars;                             // the first layer's activation tile has landed
    uint64_t* w_full = bars + 1;                           // [2]
    uint64_t* w_empty = bars + 3;                          // [2]
    uint64_t* acc_full = bars + 5;                         // layer l's accumulators complete (phase l)
    uint64_t* act_ready = bars + 6;                        // layer l's output tile is in shared memory (phase l)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
    __shared__ float sbias[kTapChainMax * 128];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) TAP_DBG(0);
    if (threadIdx.x == 0 && g.dbg) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g.dbg[(size_t)blockIdx.x * 16 + 14] = (long long)gt;
    }
    const int win0 = blockIdx.x * 4 * g.wpq;
    const int rows_q = g.wpq * g.T;
    const int nkb0 = g.L[0].num_kb;
    // buffer plan: layer 0 reads buffers 0 .. nkb0-1; layer l >= 1 reads what layer l-1 wrote and writes the other one
    auto in_buf = [&](int l) { return l == 0 ? 0 : ((nkb0 == 2 ? 0 : 1) + (l - 1)) & 1; };
    auto out_buf = [&](int l) { return ((nkb0 == 2 ? 0 : 1) + l) & 1; };

    // weight blocks in issue order: (layer, slab, K block); block b goes to ring slot b & 1
    auto issue_weight_block = [&](int b, int l, int y, int kb) {
        const int np = 3 * g.L[l].ncta;
        const uint32_t bt = (uint32_t)np * 128;
        const int slot = b & 1;
        uint8_t* dst = wring + slot * kChainWSlot;
        mbar_arrive_expect_tx(&w_full[slot], 3 * bt);
        tma_load_2d(dst, &maps.w[l][0], kb * 64, y * np, &w_full[slot]);
        tma_load_2d(dst + bt, &maps.w[l][1], kb * 64, y * np, &w_full[slot]);
        tma_load_2d(dst + 2 * bt, &maps.w[l][2], kb * 64, y * np, &w_full[slot]);
    };
    if (warp == 0 && lane == 0) {
        mbar_init(act_full, 1);
        mbar_init(&w_full[0], 1), mbar_init(&w_full[1], 1), mbar_init(&w_empty[0], 1), mbar_init(&w_empty[1], 1);
        mbar_init(acc_full, 1), mbar_init(act_ready, kChainEpiWarps);
        fence_barrier_init();
        // the first tiles are requested before the CTA-wide setup barrier: the activation tile and the first two
        // weight blocks (both ring slots are empty) travel while TMEM is allocated and the idle rows are cleared
        const uint32_t box_bytes = (uint32_t)rows_q * 128;
        mbar_arrive_expect_tx(act_full, (uint32_t)nkb0 * 8 * box_bytes);
        for (int kb = 0; kb < nkb0; ++kb) {
            uint8_t* dst = act + kb * 2 * kATile;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                tma_load_3d(dst + q * kQuarterBytes, &maps.a_hi, kb * 64, 0, win0 + q * g.wpq, act_full);
                tma_load_3d(dst + kATile + q * kQuarterBytes, &maps.a_lo, kb * 64, 0, win0 + q * g.wpq, act_full);
            }
        }
        int b = 0;
        for (int l = 0; l < g.nl && b < 2; ++l)
            for (int y = 0; y < g.L[l].gridy && b < 2; ++y)
                for (int kb = 0; kb < g.L[l].num_kb && b < 2; ++kb, ++b) issue_weight_block(b, l, y, kb);
        for (int l = 1; l < g.nl; ++l) prefetch_tmap(&maps.w[l][0]), prefetch_tmap(&maps.w[l][1]), prefetch_tmap(&maps.w[l][2]);
    }
    // biases of every layer, zero beyond N (the epilogue adds them unguarded)
    for (int i = threadIdx.x; i < g.nl * 128; i += kChainThreads) {
        const ChainLayer& L = g.L[i >> 7];
        const int n = i & 127;
        sbias[i] = (L.bias && n < L.N) ? __ldg(L.bias + n) : 0.f;
    }
    // the idle rows of each quarter are never written by TMA: clear them once so the first layer's MMA reads zeros
    {
        const int idle = 32 - rows_q;
        const int total = 4 * 4 * idle * 8;                    // 16-byte slots in the four tiles
        for (int i = threadIdx.x; i < total; i += kChainThreads) {
            const int tile = i / (4 * idle * 8), r = i % (4 * idle * 8);
            const int q = r / (idle * 8), rr = (r % (idle * 8)) / 8, c16 = r % 8;
            uint8_t* p = act + tile * kATile + q * kQuarterBytes + (rows_q + rr) * 128 + c16 * 16;
            *reinterpret_cast<float4*>(p) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TAP_DBG(1);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int b = 0;                                       // (blocks 0 and 1 were requested during setup)
            for (int l = 0; l < g.nl; ++l)
                for (int y = 0; y < g.L[l].gridy; ++y)
                    for (int kb = 0; kb < g.L[l].num_kb; ++kb, ++b) {
                        if (b < 2) continue;
                        mbar_wait(&w_empty[b & 1], (uint32_t)(((b >> 1) & 1) ^ 1));
                        issue_weight_block(b, l, y, kb);
                    }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            int b = 0;
            for (int l = 0; l < g.nl; ++l) {
                if (l == 0) mbar_wait(act_full, 0);
                else mbar_wait(act_ready, (uint32_t)((l - 1) & 1));
                tc_fence_after();
                if (l < 4) TAP_DBG(2 + l);
                const int np = 3 * g.L[l].ncta;
                const uint32_t bt = (uint32_t)np * 128;
                const uint32_t idesc = instr_desc_f16(kRows, np);
                for (int y = 0; y < g.L[l].gridy; ++y) {
                    const uint32_t acc = tmem_base + (uint32_t)(((l + y) & 1) * 192);
                    for (int kb = 0; kb < g.L[l].num_kb; ++kb, ++b) {
                        const int slot = b & 1;
                        mbar_wait(&w_full[slot], (uint32_t)((b >> 1) & 1));
                        tc_fence_after();
                        const uint8_t* in = act + (l == 0 ? kb : in_buf(l)) * 2 * kATile;
                        const uint32_t wb = smem_u32(wring + slot * kChainWSlot);
                        const uint64_t a_hi = make_smem_desc(smem_u32(in)), a_lo = make_smem_desc(smem_u32(in + kATile));
                        const uint64_t b_hi = make_smem_desc(wb), b_lo = make_smem_desc(wb + bt), b_hs = make_smem_desc(wb + 2 * bt);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);
                            umma_f16(acc, a_lo + adv, b_hs + adv, idesc, (kb > 0) || (k != 0));      // small terms first
                            umma_f16(acc, a_hi + adv, b_lo + adv, idesc, 1);
                            umma_f16(acc, a_hi + adv, b_hi + adv, idesc, 1);
                        }
                        umma_commit(&w_empty[slot]);
                    }
                }
                umma_commit(acc_full);
            }
        }
    } else {
        // ===== epilogue warps 2..17: TMEM lane quarter = warp % 4, one 16-column chunk per warp =====
        const int q = warp & 3;
        const int sub = (warp - 2) >> 2;
        const int wl = lane / g.T, t = lane - wl * g.T;         // window inside the quarter, frame
        const int winq = win0 + q * g.wpq;
        const int win = winq + wl;
        const bool row_ok = lane < rows_q && win < g.W;
        const bool has_prev = t > 0, has_next = t < g.T - 1;
        const size_t token = (size_t)win * g.T + t;
        for (int l = 0; l < g.nl; ++l) {
            const ChainLayer& L = g.L[l];
            const int halves = L.N >> 4;                        // 16-bit sign halves per token
            // the masks of this layer's chunks do not depend on the accumulators: fetch them while the MMAs run
            uint32_t mbits_y[2] = {0u, 0u};
            if (L.epi == EPI_MASK && row_ok && sub * 16 < L.ncta) {
                for (int y = 0; y < L.gridy; ++y)
                    mbits_y[y] = __ldg(reinterpret_cast<const uint16_t*>(L.aux_bits) + token * halves + ((y * L.ncta + sub * 16) >> 4));
            }
            mbar_wait(acc_full, (uint32_t)(l & 1));
            tc_fence_after();
            if (threadIdx.x == 64 && l < 4) TAP_DBG(8 + l);
            for (int y = 0; y < L.gridy; ++y) {
                uint8_t* stage = act + (y == 0 ? out_buf(l) : in_buf(l)) * 2 * kATile;
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(((l + y) & 1) * 192);
                const int c = sub;
                if (c * 16 < L.ncta) {
                    uint32_t p0[16], p1[16], p2[16];
                    tmem_ld_32x32b_x16(trow + (uint32_t)(0 * L.ncta + c * 16), p0);
                    tmem_ld_32x32b_x16(trow + (uint32_t)(1 * L.ncta + c * 16), p1);
                    tmem_ld_32x32b_x16(trow + (uint32_t)(2 * L.ncta + c * 16), p2);
                    const int nb = y * L.ncta + c * 16;
                    const uint32_t mbits = mbits_y[y];
                    tmem_ld_wait();
                    float o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(p0[j]), 1);     // P_0 of the previous frame
                        const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(p2[j]), 1);   // P_2 of the next frame
                        float v = __uint_as_float(p1[j]);
                        v += has_prev ? up : 0.f;
                        v += has_next ? dn : 0.f;
                        o[j] = v * (1.f / kWScale16);
                    }
                    if (L.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(sbias + l * 128 + nb);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 bv = bp[j >> 2];
                            o[j] += bv.x, o[j + 1] += bv.y, o[j + 2] += bv.z, o[j + 3] += bv.w;
                        }
                    }
                    if (L.epi == EPI_LRELU) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[j] = o[j] > 0.f ? o[j] : o[j] * 0.01f;
                    } else if (L.epi == EPI_MASK) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[j] = ((mbits >> j) & 1u) ? o[j] : o[j] * 0.01f;
                    }
                    if (L.sign_out) {
                        uint32_t sbits = 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) sbits |= (o[j] > 0.f ? 1u : 0u) << j;
                        if (row_ok) reinterpret_cast<uint16_t*>(L.sign_out)[token * halves + (nb >> 4)] = (uint16_t)sbits;
                    }
                    if (L.out_kind != 2) {
                        if (g.status) {             // fp16 range check (the split saturates silently)
                            float amax = 0.f;
#pragma unroll
                            for (int j = 0; j < 16; ++j) amax = fmaxf(amax, fabsf(o[j]));
                            if (amax > 65504.f && row_ok) atomicOr(g.status + win, GEM_WIN_F16_RANGE);
                        }
                        uint8_t* th = stage + q * kQuarterBytes + lane * 128;
                        uint8_t* tl = th + kATile;
#pragma unroll
                        for (int j = 0; j < 16; j += 8) {
                            uint16_t h[8], lo8[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) split_f16(o[j + e], h[e], lo8[e]);
                            const int off = (((c * 2 + (j >> 3)) ^ (lane & 7)) << 4);            // SWIZZLE_128B
                            *reinterpret_cast<uint4*>(th + off) =
                                make_uint4((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16),
                                           (uint32_t)h[4] | ((uint32_t)h[5] << 16), (uint32_t)h[6] | ((uint32_t)h[7] << 16));
                            *reinterpret_cast<uint4*>(tl + off) =
                                make_uint4((uint32_t)lo8[0] | ((uint32_t)lo8[1] << 16), (uint32_t)lo8[2] | ((uint32_t)lo8[3] << 16),
                                           (uint32_t)lo8[4] | ((uint32_t)lo8[5] << 16), (uint32_t)lo8[6] | ((uint32_t)lo8[7] << 16));
                        }
                    } else {
                        float* sp = reinterpret_cast<float*>(stage + q * 2 * kQuarterBytes) + lane * L.N + c * 16;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (nb + j < L.N) sp[j] = o[j];
                    }
                }
                if (L.out_kind == 0) {
                    // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
                    fence_proxy_async_smem();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(act_ready);
                } else if (L.out_kind == 1) {
                    fence_proxy_async_smem();
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
                    if (sub == 0 && lane == 0 && winq < g.W) {
                        tma_store_3d(&maps.o_hi, stage + q * kQuarterBytes, y * L.ncta, 0, winq);
                        tma_store_3d(&maps.o_lo, stage + kATile + q * kQuarterBytes, y * L.ncta, 0, winq);
                        bulk_commit();
                        bulk_wait_read0();
                    }
                } else {
                    // plain output (the pose): the quarter's windows are consecutive, dense rows in global memory
                    asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
                    int nwin = g.W - winq;
                    nwin = nwin < 0 ? 0 : (nwin > g.wpq ? g.wpq : nwin);
                    const int count = nwin * g.T * L.N;
                    const float* sp = reinterpret_cast<const float*>(stage + q * 2 * kQuarterBytes);
                    float* dp = g.out_plain + (size_t)winq * g.T * L.N;
                    for (int i = sub * 32 + lane; i < count; i += 128) dp[i] = sp[i];
                }
            }
        }
    }
    if (threadIdx.x == 64) TAP_DBG(5);
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (threadIdx.x == 0) TAP_DBG(6);
    if (threadIdx.x == 0 && g.dbg) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g.dbg[(size_t)blockIdx.x * 16 + 15] = (long long)gt;
    }
}


// ---- CTA-pair chain: a whole direction of the decoder's k=3 layers in one launch -----------------------------------
// The 256 <-> 128 channel layers do not fit the one-CTA chain (their input alone is four activation buffers) and, on
// their own, they are bound by the 590 KB of weight slabs every 128-row tile has to pull through L2.  Two CTAs of a
// cluster run each layer as ONE M = 256 MMA (tcgen05.mma.cta_group::2): each CTA keeps its own twelve windows'
// activations but only HALF of every weight block (96 of the 192 slab rows), so the weight traffic per window halves,
// the ring slots shrink to 36 KB, and four activation buffers fit beside them.  Chains:
//   forward   256 -> 128 -> 64 -> 64 -> 64 -> pose        backward   d pose -> 64 -> 64 -> 64 -> 128 -> 256
// A layer with more than two output slabs is split into pseudo-layers of at most two slabs (two accumulators of
// 192 TMEM columns); every pseudo-layer but the last arrives on act_ready when its accumulators are drained and its
// output (if it feeds the next layer) is in shared memory.  Barriers the peer signals live in the LEADER: act_full
// and w_full (both CTAs' TMA bytes), act_ready (both CTAs' epilogue warps); the leader's commits are multicast to
// the w_empty / acc_full barriers of both CTAs.
constexpr int kChain2Max = 6;
constexpr int kChain2WSlot = 3 * 96 * 128;                     // 36,864 B: three half slabs of 96 rows x 128 B
// shared memory: n_act activation buffers (hi + lo tile each) and n_slots weight slots: 4 + 2 (200 KB), or - when the
// first layer's K blocks stream through two buffers and no later layer needs more - 2 + 4 (208 KB): four weight
// blocks in flight instead of two
constexpr int kChain2SlotsMax = 4;
constexpr size_t kChain2Smem = (8 * kATile + 2 * kChain2WSlot > 4 * kATile + 4 * kChain2WSlot ? 8 * kATile + 2 * kChain2WSlot
                                                                                             : 4 * kATile + 4 * kChain2WSlot) +
                               1024 /*align slack*/ + 256 /*barriers*/;

struct Chain2Layer {
    const float* bias;           // full-layer bias [N] or NULL
    const uint32_t* aux_bits;
    uint32_t* sign_out;
    int ncta, y0, ny, num_kb, N, epi, out_kind;      // slabs y0 .. y0+ny-1 (ny <= 2) of a layer with N outputs
    unsigned char in_buf[4], stage_buf[2];
    int wmap;                    // index of the layer's weight maps
};
struct Chain2Maps {
    CUtensorMap a_hi, a_lo;
    CUtensorMap w[kChain2Max][3];
    CUtensorMap o_hi, o_lo;
};
struct Chain2Args {
    Chain2Layer L[kChain2Max];
    int nl, W, T, wpq;
    int n_act, n_slots;          // activation buffers (2 or 4) and weight ring slots (4 or 2)
    int stream0;                 // the first layer's K blocks stream through buffers 0 / 1 (loop order: K block outer, slab inner)
    float* out_plain;
    uint32_t* status;            // optional [W]: GEM_WIN_F16_RANGE is OR-ed in when a split activation saturates
    long long* dbg;
};

// Energy prologue of the backward chain (optional): instead of loading d pose tiles that a separate energy kernel
// wrote, the epilogue warps evaluate the energy terms and their analytic gradient for the CTA's twelve windows
// themselves (energy_device.cuh: the same arithmetic as energy_grad_kernel, the same reduction tree) from the pose the
// forward chain left in global memory, scale and split dE/dpose into the first layer's swizzled operand tile, and
// arrive on act_full.  dE/dpose, its split copies and one launch per round never exist.
struct Chain2Energy : EnergyCommon {
    int enabled;
    int in_buf, scratch0, scratch1;      // activation buffers: first layer's operand; pose / anchor + partial sums
    const float* pose;                   // [W][T][J][3] decoded pose (the forward chain's output)
    const float* pose0;                  // [W][T][J][3] the stage's anchor
    float* energy;                       // [W]
    int32_t* row_exp;                    // [W] exponent of the window's gradient scale
};
static_assert(sizeof(Chain2Maps) + sizeof(Chain2Args) + sizeof(Chain2Energy) <= 4096, "kernel parameters beyond 4 KB");
constexpr int kEnergyQuarterFloats = 2048;      // scratch floats per TMEM lane quarter in each scratch buffer
constexpr int kEnergyRedOffset = 1440;          // partial sums live behind the anchor poses in scratch1

// kEnergy: the variant with the energy prologue compiled in (its per-joint code is 6 k instructions; the plain
// chain stays at its 4 k, which its instruction-cache behaviour depends on: stall_no_inst was 18 % of the samples with
// both in one kernel)
template <bool kEnergy>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kChainThreads, 1)
tc_tap_chain2_kernel(const __grid_constant__ Chain2Maps maps, const __grid_constant__ Chain2Args g,
                     const __grid_constant__ Chain2Energy en_arg) {
    const Chain2Energy& en = en_arg;
    const bool en_on = kEnergy && en_arg.enabled;      // (folds to false at compile time in the plain variant)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* act = smem;                                   // n_act buffers of (hi 16 KB, lo 16 KB)
    uint8_t* wring = smem + (size_t)g.n_act * 2 * kATile;  // n_slots half-block slots
    uint64_t* bars = reinterpret_cast<uint64_t*>(wring + (size_t)g.n_slots * kChain2WSlot);
    uint64_t* act_full = bars;                             // [2] leader: both CTAs' first-layer activation tiles landed
                                                           //     (streaming: per buffer, phase = use of the buffer)
    uint64_t* act_empty = bars + 2;                        // [2] each CTA (multicast commit): streaming buffer consumed
    uint64_t* w_full = bars + 4;                           // [n_slots] leader: both halves of a weight block landed
    uint64_t* w_empty = bars + 4 + kChain2SlotsMax;        // [n_slots] each CTA (multicast commit)
    uint64_t* acc_full = bars + 4 + 2 * kChain2SlotsMax;   // each CTA (multicast commit), phase = pseudo-layer
    uint64_t* act_ready = acc_full + 1;                    // leader: 32 epilogue warps of the pair, phase = pseudo-layer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(act_ready + 1);
    const int nslots = g.n_slots;
    __shared__ float sbias[kChain2Max * 128];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) TAP_DBG(0);
    if (threadIdx.x == 0 && g.dbg) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g.dbg[(size_t)blockIdx.x * 16 + 14] = (long long)gt;
    }
    const int win0 = blockIdx.x * 4 * g.wpq;               // this CTA's twelve windows
    const int rows_q = g.wpq * g.T;
    const int nkb0 = g.L[0].num_kb;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&maps.a_hi), prefetch_tmap(&maps.a_lo);
        mbar_init(act_full, en_on ? 2 * kChainEpiWarps : 1);      // energy prologue: the pair's epilogue warps arrive
        mbar_init(&act_full[1], 1), mbar_init(&act_empty[0], 1), mbar_init(&act_empty[1], 1);
        for (int i = 0; i < kChain2SlotsMax; ++i) mbar_init(&w_full[i], 1), mbar_init(&w_empty[i], 1);
        mbar_init(acc_full, 1), mbar_init(act_ready, 2 * kChainEpiWarps);
        fence_barrier_init();
    }
    // biases of every pseudo-layer's slabs, zero beyond N (the epilogue adds them unguarded)
    for (int i = threadIdx.x; i < g.nl * 128; i += kChainThreads) {
        const Chain2Layer& L = g.L[i >> 7];
        const int n = L.y0 * L.ncta + (i & 127);
        sbias[i] = (L.bias && (i & 127) < L.ny * L.ncta && n < L.N) ? __ldg(L.bias + n) : 0.f;
    }
    // idle rows of each quarter are never written by TMA: clear them once so the first layer's MMA reads zeros
    {
        const int idle = 32 - rows_q;
        const int total = 2 * g.n_act * 4 * idle * 8;          // 16-byte slots in the activation tiles
        for (int i = threadIdx.x; i < total; i += kChainThreads) {
            const int tile = i / (4 * idle * 8), r = i % (4 * idle * 8);
            const int q = r / (idle * 8), rr = (r % (idle * 8)) / 8, c16 = r % 8;
            uint8_t* p = act + tile * kATile + q * kQuarterBytes + (rows_q + rr) * 128 + c16 * 16;
            *reinterpret_cast<float4*>(p) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (en_on) {
            // the energy prologue writes only the real channels of valid windows: padding columns, idle rows and the
            // rows of windows beyond W must read zero (the TMA unit zero-fills them on the loading path)
            float4* p = reinterpret_cast<float4*>(act + en.in_buf * 2 * kATile);
            for (int i = threadIdx.x; i < 2 * kATile / 16; i += kChainThreads) p[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();                  // the peer's barriers exist before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) TAP_DBG(1);

    if (warp == 0) {
        // ===== TMA producer (both CTAs: own activations, own half of every weight block) =====
        if (lane == 0) {
            const uint32_t box_bytes = (uint32_t)rows_q * 128;
            const uint32_t act_bar = mapa_u32(smem_u32(act_full), 0);
            const bool stream0 = g.stream0 && !en_on;
            auto load_act = [&](int kb, int buf, uint32_t bar) {
                uint8_t* dst = act + buf * 2 * kATile;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    tma_load_3d_pair(dst + q * kQuarterBytes, &maps.a_hi, kb * 64, 0, win0 + q * g.wpq, bar);
                    tma_load_3d_pair(dst + kATile + q * kQuarterBytes, &maps.a_lo, kb * 64, 0, win0 + q * g.wpq, bar);
                }
            };
            int b = 0;
            auto load_weights = [&](const Chain2Layer& L, int i, int kb) {
                const int half = 3 * L.ncta / 2;                           // slab rows per CTA
                const uint32_t bt = (uint32_t)half * 128;
                const int slot = b % nslots;
                mbar_wait(&w_empty[slot], (uint32_t)(((b / nslots) & 1) ^ 1));
                uint8_t* dst = wring + slot * kChain2WSlot;
                if (rank == 0) mbar_arrive_expect_tx(&w_full[slot], 2u * 3u * bt);
                const uint32_t bar = mapa_u32(smem_u32(&w_full[slot]), 0);
                const int row = (L.y0 + i) * 3 * L.ncta + (int)rank * half;
                tma_load_2d_pair(dst, &maps.w[L.wmap][0], kb * 64, row, bar);
                tma_load_2d_pair(dst + bt, &maps.w[L.wmap][1], kb * 64, row, bar);
                tma_load_2d_pair(dst + 2 * bt, &maps.w[L.wmap][2], kb * 64, row, bar);
                ++b;
            };
            if (!stream0) {
                if (rank == 0 && !en_on) mbar_arrive_expect_tx(act_full, 2u * (uint32_t)nkb0 * 8 * box_bytes);
                for (int kb = 0; kb < (en_on ? 0 : nkb0); ++kb) load_act(kb, g.L[0].in_buf[kb], act_bar);
            }
            for (int l = 0; l < g.nl; ++l) prefetch_tmap(&maps.w[g.L[l].wmap][0]);
            for (int l = 0; l < g.nl; ++l) {
                const Chain2Layer& L = g.L[l];
                if (l == 0 && stream0) {
                    // K block outer, slab inner: the activation K blocks stream through buffers 0 / 1
                    for (int kb = 0; kb < L.num_kb; ++kb) {
                        const int ab = kb & 1, use = kb >> 1;
                        if (use > 0) mbar_wait(&act_empty[ab], (uint32_t)((use - 1) & 1));
                        if (rank == 0) mbar_arrive_expect_tx(&act_full[ab], 2u * 8 * box_bytes);
                        load_act(kb, ab, mapa_u32(smem_u32(&act_full[ab]), 0));
                        for (int i = 0; i < L.ny; ++i) load_weights(L, i, kb);
                    }
                    continue;
                }
                for (int i = 0; i < L.ny; ++i)
                    for (int kb = 0; kb < L.num_kb; ++kb) load_weights(L, i, kb);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (lane == 0 && rank == 0) {
            int b = 0;
            for (int l = 0; l < g.nl; ++l) {
                const Chain2Layer& L = g.L[l];
                if (l == 0 && g.stream0 && !en_on) {}                     // (waits per K block below)
                else if (l == 0 && !en_on) mbar_wait(act_full, 0);
                else if (l == 0) mbar_wait_cluster(act_full, 0);          // the pair's energy prologues have written the tiles
                else mbar_wait_cluster(act_ready, (uint32_t)((l - 1) & 1));
                tc_fence_after();
                if (l < 3) TAP_DBG(2 + l);
                const int np = 3 * L.ncta;
                const uint32_t bt = (uint32_t)(np / 2) * 128;
                const uint32_t idesc = instr_desc_f16(2 * kRows, np);
                auto block = [&](int i, int kb, int buf) {
                    const uint32_t acc = tmem_base + (uint32_t)(((l + i) & 1) * 192);
                    const int slot = b % nslots;
                    mbar_wait(&w_full[slot], (uint32_t)((b / nslots) & 1));
                    tc_fence_after();
                    const uint8_t* in = act + buf * 2 * kATile;
                    const uint32_t wb = smem_u32(wring + slot * kChain2WSlot);
                    const uint64_t a_hi = make_smem_desc(smem_u32(in)), a_lo = make_smem_desc(smem_u32(in + kATile));
                    const uint64_t b_hi = make_smem_desc(wb), b_lo = make_smem_desc(wb + bt), b_hs = make_smem_desc(wb + 2 * bt);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        umma_f16_pair(acc, a_lo + adv, b_hs + adv, idesc, (kb > 0) || (k != 0));      // small terms first
                        umma_f16_pair(acc, a_hi + adv, b_lo + adv, idesc, 1);
                        umma_f16_pair(acc, a_hi + adv, b_hi + adv, idesc, 1);
                    }
                    umma_commit_pair(&w_empty[slot], 3);
                    ++b;
                };
                if (l == 0 && g.stream0 && !en_on) {
                    // K block outer, slab inner (every accumulator still sums its K blocks in ascending order)
                    for (int kb = 0; kb < L.num_kb; ++kb) {
                        const int ab = kb & 1;
                        mbar_wait(&act_full[ab], (uint32_t)((kb >> 1) & 1));
                        tc_fence_after();
                        for (int i = 0; i < L.ny; ++i) block(i, kb, ab);
                        umma_commit_pair(&act_empty[ab], 3);           // the buffer may take the K block after next
                    }
                } else {
                    for (int i = 0; i < L.ny; ++i)
                        for (int kb = 0; kb < L.num_kb; ++kb) block(i, kb, L.in_buf[kb]);
                }
                umma_commit_pair(acc_full, 3);
            }
        }
    } else {
        // ===== epilogue warps 2..17 (both CTAs): TMEM lane quarter = warp % 4, one 16-column chunk per warp =====
        const int q = warp & 3;
        const int sub = (warp - 2) >> 2;
        const int wl = lane / g.T, t = lane - wl * g.T;
        const int winq = win0 + q * g.wpq;
        const int win = winq + wl;
        const bool row_ok = lane < rows_q && win < g.W;
        const bool has_prev = t > 0, has_next = t < g.T - 1;
        const size_t token = (size_t)win * g.T + t;
        const uint32_t ready_bar = mapa_u32(smem_u32(act_ready), 0);
        const int c = sub;
        if constexpr (kEnergy) if (en_on) {
            // ===== energy prologue: this quarter's windows, 32 joint-frames per (virtual) warp as in energy_grad_kernel =====
            constexpr int wpw = kEnergySlot / 32;                         // partial sums per window
            const int TJ = en.T * en.J, n = TJ * 3;
            float* s_x = reinterpret_cast<float*>(act + en.scratch0 * 2 * kATile) + q * kEnergyQuarterFloats;
            float* s_x0 = reinterpret_cast<float*>(act + en.scratch1 * 2 * kATile) + q * kEnergyQuarterFloats;
            float* s_red = s_x0 + kEnergyRedOffset;                      // [wpq][wpw][6]
            const int tq = sub * 32 + lane;                               // thread within the quarter's four warps
            int nwin = g.W - winq;
            nwin = nwin < 0 ? 0 : (nwin > g.wpq ? g.wpq : nwin);
            {
                // the quarter's windows are one contiguous run of floats (8-byte aligned: T*J*3 is even here, else scalar);
                // all loads are issued before the first store so that one memory latency covers them
                const float* gp = en.pose + (size_t)winq * n;
                const float* gp0 = en.pose0 + (size_t)winq * n;
                const int total = nwin * n;
                if ((n & 1) == 0) {
                    constexpr int kMaxV = (kEnergyRedOffset / 2 + 127) / 128;          // float2 loads per thread
                    float2 a[kMaxV], b[kMaxV];
#pragma unroll
                    for (int u = 0; u < kMaxV; ++u) {
                        const int i = tq + u * 128;
                        a[u] = b[u] = make_float2(0.f, 0.f);
                        if (2 * i < total) {
                            a[u] = __ldg(reinterpret_cast<const float2*>(gp) + i);
                            b[u] = __ldg(reinterpret_cast<const float2*>(gp0) + i);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kMaxV; ++u) {
                        const int i = tq + u * 128;
                        if (2 * i < total) {
                            reinterpret_cast<float2*>(s_x)[i] = a[u];
                            reinterpret_cast<float2*>(s_x0)[i] = b[u];
                        }
                    }
                } else {
                    for (int i = tq; i < total; i += 128) s_x[i] = gp[i], s_x0[i] = gp0[i];
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
            float gv[4][3];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                // two joints per thread at a time: both joints' texel loads are in flight before either is consumed
                JointTexels jt[2];
                int kk[2], ww[2];
                bool on[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int v = sub + (2 * half + u) * 4;               // virtual warp: window v / wpw, joints (v % wpw) * 32 ..
                    ww[u] = v / wpw, kk[u] = (v - ww[u] * wpw) * 32 + lane;
                    on[u] = ww[u] < nwin && kk[u] < TJ;
                    jt[u].rp = false;
                    if (on[u]) joint_gather(en, ShapeDyn(en), s_x + ww[u] * n, winq + ww[u], kk[u], jt[u]);
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int it = 2 * half + u;
                    const int v = sub + it * 4;
                    float e5[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, g3[3] = {0.f, 0.f, 0.f};
                    if (on[u]) joint_terms(en, ShapeDyn(en), s_x + ww[u] * n, s_x0 + ww[u] * n, winq + ww[u], kk[u], jt[u], e5, g3);
                    gv[it][0] = g3[0], gv[it][1] = g3[1], gv[it][2] = g3[2];
                    if (v < g.wpq * wpw) {                                // (warp-uniform)
                        const float r0 = warp_sum(e5[0]), r1 = warp_sum(e5[1]), r2 = warp_sum(e5[2]), r3 = warp_sum(e5[3]),
                                    r4 = warp_sum(e5[4]);
                        const float r5 = warp_max(fmaxf(fabsf(g3[0]), fmaxf(fabsf(g3[1]), fabsf(g3[2]))));
                        if (lane == 0) {
                            float* d = s_red + v * 6;
                            d[0] = r0, d[1] = r1, d[2] = r2, d[3] = r3, d[4] = r4, d[5] = r5;
                        }
                    }
                }
            }
            asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
            if (tq < nwin) {                                              // one thread per window: the sums in order
                float t5[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
                for (int qq = 0; qq < wpw; ++qq)
#pragma unroll
                    for (int cc = 0; cc < 5; ++cc) t5[cc] += s_red[(tq * wpw + qq) * 6 + cc];
                en.energy[winq + tq] = combine_energy(en, t5);
            }
            uint8_t* tile_hi = act + en.in_buf * 2 * kATile + q * kQuarterBytes;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int v = sub + it * 4;
                const int wl_e = v / wpw, k = (v - wl_e * wpw) * 32 + lane;
                if (wl_e < nwin && k < TJ) {
                    float mx = 0.f;
                    for (int qq = 0; qq < wpw; ++qq) mx = fmaxf(mx, s_red[(wl_e * wpw + qq) * 6 + 5]);
                    const int e = grad_exponent(mx);
                    const float sc = exp2f((float)e);
                    if (k == 0) en.row_exp[winq + wl_e] = e;
                    const int tt = k / en.J, j = k - tt * en.J;
                    const int row = wl_e * en.T + tt;                    // TMEM lane / operand row within the quarter
                    uint8_t* th = tile_hi + row * 128;
                    uint8_t* tl = th + kATile;
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) {
                        uint16_t h, lo16;
                        split_f16_energy(gv[it][cc] * sc, h, lo16);
                        const int col = j * 3 + cc;
                        const int off = (((col >> 3) ^ (row & 7)) << 4) + (col & 7) * 2;       // SWIZZLE_128B
                        *reinterpret_cast<uint16_t*>(th + off) = h;
                        *reinterpret_cast<uint16_t*>(tl + off) = lo16;
                    }
                }
            }
            fence_proxy_async_smem();              // generic-proxy writes -> visible to the tensor core's operand reads
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(act_full), 0));
        }
        for (int l = 0; l < g.nl; ++l) {
            const Chain2Layer& L = g.L[l];
            const int halves = L.N >> 4;
            const bool has_chunk = c * 16 < L.ncta;
            uint32_t mbits_i[2] = {0u, 0u};
            if (L.epi == EPI_MASK && row_ok && has_chunk) {
                for (int i = 0; i < L.ny; ++i)
                    mbits_i[i] = __ldg(reinterpret_cast<const uint16_t*>(L.aux_bits) + token * halves +
                                       (((L.y0 + i) * L.ncta + c * 16) >> 4));
            }
            mbar_wait(acc_full, (uint32_t)(l & 1));
            tc_fence_after();
            if (threadIdx.x == 64 && l < 6) TAP_DBG(8 + l);
            for (int i = 0; i < L.ny; ++i) {
                const int y = L.y0 + i;
                uint8_t* stage = act + L.stage_buf[i] * 2 * kATile;
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(((l + i) & 1) * 192);
                if (has_chunk) {
                    uint32_t p0[16], p1[16], p2[16];
                    tmem_ld_32x32b_x16(trow + (uint32_t)(0 * L.ncta + c * 16), p0);
                    tmem_ld_32x32b_x16(trow + (uint32_t)(1 * L.ncta + c * 16), p1);
                    tmem_ld_32x32b_x16(trow + (uint32_t)(2 * L.ncta + c * 16), p2);
                    const int nb = y * L.ncta + c * 16;
                    const uint32_t mbits = mbits_i[i];
                    tmem_ld_wait();
                    float o[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(p0[j]), 1);     // P_0 of the previous frame
                        const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(p2[j]), 1);   // P_2 of the next frame
                        float v = __uint_as_float(p1[j]);
                        v += has_prev ? up : 0.f;
                        v += has_next ? dn : 0.f;
                        o[j] = v * (1.f / kWScale16);
                    }
                    if (L.bias) {
                        const float4* bp = reinterpret_cast<const float4*>(sbias + l * 128 + i * L.ncta + c * 16);
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            const float4 bv = bp[j >> 2];
                            o[j] += bv.x, o[j + 1] += bv.y, o[j + 2] += bv.z, o[j + 3] += bv.w;
                        }
                    }
                    if (L.epi == EPI_LRELU) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[j] = o[j] > 0.f ? o[j] : o[j] * 0.01f;
                    } else if (L.epi == EPI_MASK) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) o[j] = ((mbits >> j) & 1u) ? o[j] : o[j] * 0.01f;
                    }
                    if (L.sign_out) {
                        uint32_t sbits = 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) sbits |= (o[j] > 0.f ? 1u : 0u) << j;
                        if (row_ok) reinterpret_cast<uint16_t*>(L.sign_out)[token * halves + (nb >> 4)] = (uint16_t)sbits;
                    }
                    if (L.out_kind != 2) {
                        if (g.status) {             // fp16 range check (the split saturates silently)
                            float amax = 0.f;
#pragma unroll
                            for (int j = 0; j < 16; ++j) amax = fmaxf(amax, fabsf(o[j]));
                            if (amax > 65504.f && row_ok) atomicOr(g.status + win, GEM_WIN_F16_RANGE);
                        }
                        uint8_t* th = stage + q * kQuarterBytes + lane * 128;
                        uint8_t* tl = th + kATile;
#pragma unroll
                        for (int j = 0; j < 16; j += 8) {
                            uint16_t h[8], lo8[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) split_f16(o[j + e], h[e], lo8[e]);
                            const int off = (((c * 2 + (j >> 3)) ^ (lane & 7)) << 4);            // SWIZZLE_128B
                            *reinterpret_cast<uint4*>(th + off) =
                                make_uint4((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16),
                                           (uint32_t)h[4] | ((uint32_t)h[5] << 16), (uint32_t)h[6] | ((uint32_t)h[7] << 16));
                            *reinterpret_cast<uint4*>(tl + off) =
                                make_uint4((uint32_t)lo8[0] | ((uint32_t)lo8[1] << 16), (uint32_t)lo8[2] | ((uint32_t)lo8[3] << 16),
                                           (uint32_t)lo8[4] | ((uint32_t)lo8[5] << 16), (uint32_t)lo8[6] | ((uint32_t)lo8[7] << 16));
                        }
                    } else {
                        float* sp = reinterpret_cast<float*>(stage + q * 2 * kQuarterBytes) + lane * L.N + c * 16;
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            if (nb + j < L.N) sp[j] = o[j];
                    }
                }
            }
            // accumulators drained, operand tiles (if any) written: the leader may issue the next pseudo-layer
            if (l + 1 < g.nl) {
                fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's operand reads
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(ready_bar);
            }
            if (L.out_kind == 1) {
                fence_proxy_async_smem();
                asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
                if (sub == 0 && lane == 0 && winq < g.W) {
                    for (int i = 0; i < L.ny; ++i) {
                        const uint8_t* stage = act + L.stage_buf[i] * 2 * kATile;
                        tma_store_3d(&maps.o_hi, stage + q * kQuarterBytes, (L.y0 + i) * L.ncta, 0, winq);
                        tma_store_3d(&maps.o_lo, stage + kATile + q * kQuarterBytes, (L.y0 + i) * L.ncta, 0, winq);
                    }
                    bulk_commit();
                    bulk_wait_read0();
                }
            } else if (L.out_kind == 2) {
                // plain output (the pose): the quarter's windows are consecutive, dense rows in global memory
                asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");
                int nwin = g.W - winq;
                nwin = nwin < 0 ? 0 : (nwin > g.wpq ? g.wpq : nwin);
                const int count = nwin * g.T * L.N;
                const float* sp = reinterpret_cast<const float*>(act + L.stage_buf[0] * 2 * kATile + q * 2 * kQuarterBytes);
                float* dp = g.out_plain + (size_t)winq * g.T * L.N;
                for (int k = sub * 32 + lane; k < count; k += 128) dp[k] = sp[k];
            }
        }
    }
    if (threadIdx.x == 64) TAP_DBG(5);
    tc_fence_before();
    cluster_sync_all();                  // neither CTA leaves (or frees TMEM) while the pair's MMAs / reads are in flight
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
    if (threadIdx.x == 0) TAP_DBG(6);
    if (threadIdx.x == 0 && g.dbg) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        g.dbg[(size_t)blockIdx.x * 16 + 15] = (long long)gt;
    }
}

struct TapWeight {
    void *hi = nullptr, *lo = nullptr, *hs = nullptr;
    int K = 0, Kp = 0, N = 0, ncta = 0, gridy = 0;
    CUtensorMap map_hi, map_lo, map_hs;
    CUtensorMap map_hi2, map_lo2, map_hs2;     // half-height boxes (fp16 scheme): each CTA of a pair loads half a block
};
struct AMaps {
    CUtensorMap hi, lo;
};
struct KeyHash {
    size_t operator()(const std::pair<const float*, int>& k) const {
        return std::hash<const void*>()(k.first) ^ ((size_t)k.second * 0x9E3779B97F4A7C15ull);
    }
};
struct TapState {
    std::unordered_map<std::pair<const float*, int>, TapWeight, KeyHash> weights;     // (layer's weight pointer, scheme)
    std::map<std::tuple<const void*, const void*, int, int, int, int, int>, AMaps> amaps;   // (hi, lo, ld, K, W, T, f16)
    bool attr_set = false;
};
std::mutex g_mu;
std::unordered_map<void*, TapState*> g_states;

TapState* state_of(void* owner) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_states.find(owner);
    if (it == g_states.end()) it = g_states.emplace(owner, new TapState()).first;
    return it->second;
}

}  // namespace

bool tc_tap_supported(int K, int N, int T) {
    return T >= 1 && T <= 32 && K >= 1 && K <= 1024 && N >= 1 && (N <= 48 || N % 64 == 0);
}

// scheme: 1 = 3xTF32 (fp32 hi / lo slabs), 2 = fp16 (three fp16 slabs)
int tc_tap_prepare_weight(void* owner, cudaStream_t stream, const float* B, int ldb, int K, int N, int scheme) {
    GEM_REQUIRE(tc_tap_supported(K, N, 1), "layer shape not supported by the tcgen05 tap kernel");
    const bool f16 = scheme == 2;
    const int bk = f16 ? 64 : 32;
    TapState* st = state_of(owner);
    const auto key = std::make_pair(B, f16 ? 2 : 1);
    auto old = st->weights.find(key);
    if (old != st->weights.end()) {
        GEM_CUDA(cudaStreamSynchronize(stream));
        cudaFree(old->second.hi), cudaFree(old->second.lo);
        if (old->second.hs) cudaFree(old->second.hs);
        st->weights.erase(old);
    }
    TapWeight w;
    w.K = K, w.N = N, w.Kp = (K + bk - 1) / bk * bk;
    w.ncta = N <= 48 ? 48 : 64;
    w.gridy = N <= 48 ? 1 : N / 64;
    const size_t total = (size_t)w.gridy * 3 * w.ncta * w.Kp;
    const size_t esz = f16 ? 2 : 4;
    GEM_CUDA(cudaMalloc(&w.hi, total * esz));
    GEM_CUDA(cudaMalloc(&w.lo, total * esz));
    if (f16) {
        GEM_CUDA(cudaMalloc(&w.hs, total * esz));
        tap_weight_prep_f16_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
            B, ldb, K, N, w.Kp, w.ncta, w.gridy, (uint16_t*)w.hi, (uint16_t*)w.lo, (uint16_t*)w.hs);
    } else {
        tap_weight_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(B, ldb, K, N, w.Kp, w.ncta, w.gridy,
                                                                                  (float*)w.hi, (float*)w.lo);
    }
    GEM_CHECK_LAUNCH();
    const uint64_t dims[2] = {(uint64_t)w.Kp, (uint64_t)w.gridy * 3 * w.ncta};
    const uint64_t strides[1] = {(uint64_t)w.Kp * esz};
    const uint32_t box[2] = {(uint32_t)bk, (uint32_t)(3 * w.ncta)};
    int rc;
    if (f16) {
        rc = make_map_u16(&w.map_hi, w.hi, 2, dims, strides, box);
        if (rc == GEM_OK) rc = make_map_u16(&w.map_lo, w.lo, 2, dims, strides, box);
        if (rc == GEM_OK) rc = make_map_u16(&w.map_hs, w.hs, 2, dims, strides, box);
        const uint32_t box2[2] = {(uint32_t)bk, (uint32_t)(3 * w.ncta / 2)};
        if (rc == GEM_OK) rc = make_map_u16(&w.map_hi2, w.hi, 2, dims, strides, box2);
        if (rc == GEM_OK) rc = make_map_u16(&w.map_lo2, w.lo, 2, dims, strides, box2);
        if (rc == GEM_OK) rc = make_map_u16(&w.map_hs2, w.hs, 2, dims, strides, box2);
    } else {
        rc = make_map_f32(&w.map_hi, (float*)w.hi, 2, dims, strides, box);
        if (rc == GEM_OK) rc = make_map_f32(&w.map_lo, (float*)w.lo, 2, dims, strides, box);
        w.map_hs = w.map_hi;
    }
    if (rc != GEM_OK) return rc;
    st->weights.emplace(key, w);
    return GEM_OK;
}

int launch_split_pad(cudaStream_t stream, const float* src, int C, size_t tokens, int ldo, float* hi, float* lo) {
    const size_t total = tokens * (size_t)ldo;
    if (total == 0) return GEM_OK;
    split_pad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, C, tokens, ldo, hi, lo);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

int launch_rowscale_split_pad_f16(cudaStream_t stream, const float* src, int C, int T, int W, int ldo, uint16_t* hi,
                                  uint16_t* lo, int32_t* row_exp) {
    if (W <= 0) return GEM_OK;
    rowscale_split_pad_f16_kernel<<<W, 128, 0, stream>>>(src, C, T, ldo, hi, lo, row_exp);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

long long* g_tap_dbg = nullptr;   // set by gem_debug_tap_timestamps (tests / profiling only)

template <int NCTA, bool F16>
static int launch_tap_cfg(cudaStream_t stream, dim3 grid, const AMaps& am, const TapWeight& w, const AMaps& om,
                          const TapTcArgs& a) {
    static PerDeviceOnce attr_set;
    if (bool* once_ = attr_set.flag(); !*once_) {
        GEM_CUDA(cudaFuncSetAttribute(tc_tap_kernel<NCTA, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)Cfg<NCTA, F16>::kSmem));
        *once_ = true;
    }
    tc_tap_kernel<NCTA, F16><<<grid, kThreads, Cfg<NCTA, F16>::kSmem, stream>>>(am.hi, am.lo, w.map_hi, w.map_lo, w.map_hs,
                                                                                 om.hi, om.lo, a);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

int launch_tap_tc(cudaStream_t stream, void* owner, const TapTcLaunch& L) {
    if (L.W <= 0) return GEM_OK;
    const bool f16 = L.scheme == 2;
    const int bk = f16 ? 64 : 32;
    const size_t esz = f16 ? 2 : 4;
    TapState* st = state_of(owner);
    auto wit = st->weights.find(std::make_pair(L.B, f16 ? 2 : 1));
    if (wit == st->weights.end()) {
        set_error("tcgen05 tap path: weights were not prepared (tc_tap_prepare_weight)");
        return GEM_ERR_STATE;
    }
    const TapWeight& w = wit->second;
    GEM_REQUIRE(L.T >= 1 && L.T <= 32, "seq_len must be <= 32 on the tcgen05 tap path");
    GEM_REQUIRE((L.lda * esz) % 16 == 0 && L.Kreal <= L.lda && L.Kreal <= w.Kp && L.Kreal >= w.K, "bad activation layout");
    GEM_REQUIRE(L.out_lo == nullptr || ((L.ldo * esz) % 16 == 0 && w.ncta == 64), "split output needs N % 64 == 0");
    GEM_REQUIRE(L.epi != EPI_MASK || L.aux_bits || (L.aux && L.ldaux % 4 == 0 && w.N % 16 == 0),
                "mask epilogue needs sign bits or an aligned aux");
    GEM_REQUIRE((!L.aux_bits && !L.sign_out) || w.N % 32 == 0, "sign bits need N % 32 == 0");
    GEM_REQUIRE(L.out_lo != nullptr || (L.ldo == w.N && w.N <= 48), "plain output must be dense and N <= 48");
    const int wpq = 32 / L.T;
    auto maps_for = [&](const void* hi, const void* lo, int ld, int cols, AMaps** out) -> int {
        const auto key = std::make_tuple(hi, lo, ld, cols, L.W, L.T, (int)f16);
        auto mit = st->amaps.find(key);
        if (mit == st->amaps.end()) {
            AMaps m;
            const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)L.T, (uint64_t)L.W};
            const uint64_t strides[2] = {(uint64_t)ld * esz, (uint64_t)L.T * ld * esz};
            const uint32_t box[3] = {(uint32_t)bk, (uint32_t)L.T, (uint32_t)wpq};
            int rc;
            if (f16) {
                rc = make_map_u16(&m.hi, hi, 3, dims, strides, box);
                if (rc == GEM_OK) rc = make_map_u16(&m.lo, lo, 3, dims, strides, box);
            } else {
                rc = make_map_f32(&m.hi, (const float*)hi, 3, dims, strides, box);
                if (rc == GEM_OK) rc = make_map_f32(&m.lo, (const float*)lo, 3, dims, strides, box);
            }
            if (rc != GEM_OK) return rc;
            mit = st->amaps.emplace(key, m).first;
        }
        *out = &mit->second;
        return GEM_OK;
    };
    if (st->amaps.size() > 512) st->amaps.clear();
    AMaps *am = nullptr, *om = nullptr;
    {
        int rc = maps_for(L.A_hi, L.A_lo, L.lda, L.Kreal, &am);
        if (rc == GEM_OK) rc = L.out_lo ? maps_for(L.out_hi, L.out_lo, L.ldo, w.N, &om) : GEM_OK;
        if (rc != GEM_OK) return rc;
        if (!om) om = am;      // unused by the kernel when the output is plain
    }
    TapTcArgs a;
    a.bias = L.bias, a.aux = L.aux, a.aux_bits = L.aux_bits, a.sign_out = L.sign_out;
    a.out_hi = (float*)L.out_hi, a.out_lo = L.out_lo;
    a.W = L.W, a.T = L.T, a.wpq = wpq, a.N = w.N, a.ldo = L.ldo, a.ldaux = L.ldaux, a.epi = L.epi;
    a.num_kb = w.Kp / bk;
    a.dbg = g_tap_dbg;
    dim3 grid((L.W + 4 * wpq - 1) / (4 * wpq), w.gridy);
    if (w.ncta == 64)
        return f16 ? launch_tap_cfg<64, true>(stream, grid, *am, w, *om, a) : launch_tap_cfg<64, false>(stream, grid, *am, w, *om, a);
    return f16 ? launch_tap_cfg<48, true>(stream, grid, *am, w, *om, a) : launch_tap_cfg<48, false>(stream, grid, *am, w, *om, a);
}


int launch_tap_chain(cudaStream_t stream, void* owner, const TapChainLaunch& L) {
    if (L.W <= 0) return GEM_OK;
    GEM_REQUIRE(L.nl >= 1 && L.nl <= kTapChainMax, "chain length");
    GEM_REQUIRE(L.T >= 1 && L.T <= 32, "seq_len must be <= 32 on the tcgen05 tap path");
    TapState* st = state_of(owner);
    const TapWeight* w[kTapChainMax];
    for (int l = 0; l < L.nl; ++l) {
        auto wit = st->weights.find(std::make_pair(L.B[l], 2));
        if (wit == st->weights.end()) {
            set_error("tcgen05 tap chain: weights were not prepared in the fp16 scheme (tc_tap_prepare_weight)");
            return GEM_ERR_STATE;
        }
        w[l] = &wit->second;
        const bool last = l == L.nl - 1;
        GEM_REQUIRE(w[l]->Kp / 64 <= (l == 0 ? 2 : 1), "chain layers after the first take one 64-channel K block");
        GEM_REQUIRE(last || (w[l]->N == 64 && w[l]->gridy == 1), "inner chain layers must produce 64 channels");
        GEM_REQUIRE(w[l]->gridy <= 2, "the last chain layer may have at most two output slabs");
        GEM_REQUIRE(L.epi[l] != EPI_MASK || L.aux_bits[l], "mask epilogue needs sign bits");
        GEM_REQUIRE((!L.aux_bits[l] && !L.sign_out[l]) || w[l]->N % 16 == 0, "sign bits need N % 16 == 0");
    }
    const TapWeight& wl = *w[L.nl - 1];
    GEM_REQUIRE((L.lda * 2) % 16 == 0 && L.Kreal <= L.lda && L.Kreal <= w[0]->Kp && L.Kreal >= w[0]->K, "bad activation layout");
    GEM_REQUIRE(L.out_lo == nullptr || ((L.ldo * 2) % 16 == 0 && wl.ncta == 64), "split output needs N % 64 == 0");
    GEM_REQUIRE(L.out_lo != nullptr || (L.ldo == wl.N && wl.N <= 48 && wl.gridy == 1), "plain output must be dense and N <= 48");
    const int wpq = 32 / L.T;
    auto make3 = [&](CUtensorMap* m, const void* base, int ld, int cols) -> int {
        const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)L.T, (uint64_t)L.W};
        const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L.T * ld * 2};
        const uint32_t box[3] = {64u, (uint32_t)L.T, (uint32_t)wpq};
        return make_map_u16(m, base, 3, dims, strides, box);
    };
    ChainMaps maps;
    memset(&maps, 0, sizeof(maps));
    int rc = make3(&maps.a_hi, L.A_hi, L.lda, L.Kreal);
    if (rc == GEM_OK) rc = make3(&maps.a_lo, L.A_lo, L.lda, L.Kreal);
    if (rc == GEM_OK && L.out_lo) rc = make3(&maps.o_hi, L.out_hi, L.ldo, wl.N);
    if (rc == GEM_OK && L.out_lo) rc = make3(&maps.o_lo, L.out_lo, L.ldo, wl.N);
    if (rc != GEM_OK) return rc;
    if (!L.out_lo) maps.o_hi = maps.a_hi, maps.o_lo = maps.a_lo;      // unused by the kernel
    ChainArgs a;
    memset(&a, 0, sizeof(a));
    for (int l = 0; l < L.nl; ++l) {
        maps.w[l][0] = w[l]->map_hi, maps.w[l][1] = w[l]->map_lo, maps.w[l][2] = w[l]->map_hs;
        ChainLayer& c = a.L[l];
        c.bias = L.bias[l], c.aux_bits = L.aux_bits[l], c.sign_out = L.sign_out[l];
        c.ncta = w[l]->ncta, c.gridy = w[l]->gridy, c.num_kb = w[l]->Kp / 64, c.N = w[l]->N, c.epi = L.epi[l];
        c.out_kind = l < L.nl - 1 ? 0 : (L.out_lo ? 1 : 2);
    }
    for (int l = L.nl; l < kTapChainMax; ++l) maps.w[l][0] = maps.w[l][1] = maps.w[l][2] = maps.a_hi;
    a.nl = L.nl, a.W = L.W, a.T = L.T, a.wpq = wpq, a.out_plain = L.out_lo ? nullptr : (float*)L.out_hi;
    a.dbg = g_tap_dbg;
    a.status = L.status;
    static PerDeviceOnce attr_set;
    if (bool* once_ = attr_set.flag(); !*once_) {
        GEM_CUDA(cudaFuncSetAttribute(tc_tap_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmem));
        *once_ = true;
    }
    const int grid = (L.W + 4 * wpq - 1) / (4 * wpq);
    tc_tap_chain_kernel<<<grid, kChainThreads, kChainSmem, stream>>>(maps, a);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}


// A whole direction of the decoder's k=3 layers on CTA pairs (see tc_tap_chain2_kernel).  Layers with more than two
// output slabs are split into pseudo-layers; activation buffers are assigned here: a pseudo-layer's outputs go to
// buffers that none of its own inputs occupy unless it is the layer's last pseudo-layer (whose MMAs have all
// completed before its epilogue runs).
// the fused chain's energy prologue handles windows of at most 160 joint-frames whose TMEM lane quarter holds at most
// three of them (T >= 10) and a first layer of one K block (45 pose channels padded to 64)
bool tap_chain_energy_supported(int T, int J, int first_layer_k) {
    if (T < 3 || T > 32 || J < 1 || J > kMaxJoints) return false;
    const int wpq = 32 / T, TJ = T * J;
    return TJ <= kEnergySlot && wpq * (kEnergySlot / 32) <= 16 && wpq * TJ * 3 <= kEnergyRedOffset &&
           kEnergyRedOffset + wpq * (kEnergySlot / 32) * 6 <= kEnergyQuarterFloats && first_layer_k <= 64;
}

int launch_tap_chain_pair(cudaStream_t stream, void* owner, const TapChainLaunch& L) {
    if (L.W <= 0) return GEM_OK;
    GEM_REQUIRE(L.nl >= 1 && L.nl <= kTapChainMax, "chain length");
    GEM_REQUIRE(L.T >= 1 && L.T <= 32, "seq_len must be <= 32 on the tcgen05 tap path");
    TapState* st = state_of(owner);
    const TapWeight* w[kTapChainMax];
    for (int l = 0; l < L.nl; ++l) {
        auto wit = st->weights.find(std::make_pair(L.B[l], 2));
        if (wit == st->weights.end()) {
            set_error("tcgen05 tap chain: weights were not prepared in the fp16 scheme (tc_tap_prepare_weight)");
            return GEM_ERR_STATE;
        }
        w[l] = &wit->second;
        GEM_REQUIRE(w[l]->Kp / 64 <= 4 && w[l]->gridy <= 4, "chain layers take at most 256 channels in and out");
        GEM_REQUIRE(l == 0 || w[l]->Kp == (w[l - 1]->N + 63) / 64 * 64, "chain layers must fit together");
        GEM_REQUIRE(l == L.nl - 1 || w[l]->ncta == 64, "inner chain layers must produce multiples of 64 channels");
        GEM_REQUIRE(L.epi[l] != EPI_MASK || L.aux_bits[l], "mask epilogue needs sign bits");
        GEM_REQUIRE((!L.aux_bits[l] && !L.sign_out[l]) || w[l]->N % 16 == 0, "sign bits need N % 16 == 0");
    }
    const TapWeight& wl = *w[L.nl - 1];
    GEM_REQUIRE((L.lda * 2) % 16 == 0 && L.Kreal <= L.lda && L.Kreal <= w[0]->Kp && L.Kreal >= w[0]->K, "bad activation layout");
    GEM_REQUIRE(L.out_lo == nullptr || ((L.ldo * 2) % 16 == 0 && wl.ncta == 64), "split output needs N % 64 == 0");
    GEM_REQUIRE(L.out_lo != nullptr || (L.ldo == wl.N && wl.N <= 48 && wl.gridy == 1), "plain output must be dense and N <= 48");
    const int wpq = 32 / L.T;
    auto make3 = [&](CUtensorMap* m, const void* base, int ld, int cols) -> int {
        const uint64_t dims[3] = {(uint64_t)cols, (uint64_t)L.T, (uint64_t)L.W};
        const uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L.T * ld * 2};
        const uint32_t box[3] = {64u, (uint32_t)L.T, (uint32_t)wpq};
        return make_map_u16(m, base, 3, dims, strides, box);
    };
    Chain2Maps maps;
    memset(&maps, 0, sizeof(maps));
    int rc = make3(&maps.a_hi, L.A_hi, L.lda, L.Kreal);
    if (rc == GEM_OK) rc = make3(&maps.a_lo, L.A_lo, L.lda, L.Kreal);
    if (rc == GEM_OK && L.out_lo) rc = make3(&maps.o_hi, L.out_hi, L.ldo, wl.N);
    if (rc == GEM_OK && L.out_lo) rc = make3(&maps.o_lo, L.out_lo, L.ldo, wl.N);
    if (rc != GEM_OK) return rc;
    if (!L.out_lo) maps.o_hi = maps.a_hi, maps.o_lo = maps.a_lo;      // unused by the kernel
    for (int l = 0; l < kChain2Max; ++l) maps.w[l][0] = maps.w[l][1] = maps.w[l][2] = maps.a_hi;
    Chain2Args a;
    int np = 0;
    // buffer plan over nbuf activation buffers.  stream: the first layer's K blocks pass through buffers 0 / 1 while its
    // MMAs run (they are dead once its accumulators are complete, so its outputs may take them)
    auto plan = [&](int nbuf, bool stream) -> bool {
        memset(&a, 0, sizeof(a));
        np = 0;
        int cur[4], ncur = w[0]->Kp / 64;             // buffers holding the current layer's input K blocks
        for (int i = 0; i < ncur; ++i) cur[i] = stream ? (i & 1) : i;
        if (!stream && ncur > nbuf) return false;
        for (int l = 0; l < L.nl; ++l) {
            maps.w[l][0] = w[l]->map_hi2, maps.w[l][1] = w[l]->map_lo2, maps.w[l][2] = w[l]->map_hs2;
            const bool last = l == L.nl - 1;
            const int gridy = w[l]->gridy;
            if (l == 0 && stream && gridy > 2) return false;           // (one pseudo-layer: both slabs accumulate per K block)
            int next[4], nnext = 0;
            for (int y0 = 0; y0 < gridy; y0 += 2) {
                if (np >= kChain2Max) return false;
                Chain2Layer& c = a.L[np++];
                c.bias = L.bias[l], c.aux_bits = L.aux_bits[l], c.sign_out = L.sign_out[l];
                c.ncta = w[l]->ncta, c.y0 = y0, c.ny = gridy - y0 < 2 ? gridy - y0 : 2;
                c.num_kb = w[l]->Kp / 64, c.N = w[l]->N, c.epi = L.epi[l], c.wmap = l;
                c.out_kind = last ? (L.out_lo ? 1 : 2) : 0;
                for (int i = 0; i < c.num_kb; ++i) c.in_buf[i] = (unsigned char)cur[i];
                const bool last_pseudo = y0 + 2 >= gridy;
                for (int i = 0; i < c.ny; ++i) {
                    // a free buffer: not an input (unless this is the layer's last pseudo-layer), not an output kept for
                    // the next layer, not one a TMA store of an earlier pseudo-layer of this layer may still be reading
                    int pick = -1;
                    for (int bfr = 0; bfr < nbuf && pick < 0; ++bfr) {
                        bool used = false;
                        for (int k = 0; k < nnext; ++k) used |= next[k] == bfr;
                        if (!last_pseudo)
                            for (int k = 0; k < ncur; ++k) used |= cur[k] == bfr;
                        for (int k = 0; k < i; ++k) used |= c.stage_buf[k] == bfr;
                        if (!used) pick = bfr;
                    }
                    if (pick < 0) return false;
                    c.stage_buf[i] = (unsigned char)pick;
                    next[nnext++] = pick;
                }
            }
            if (nnext > 4) return false;
            for (int i = 0; i < nnext; ++i) cur[i] = next[i];
            ncur = nnext;
        }
        a.n_act = nbuf, a.n_slots = nbuf == 2 ? 4 : 2, a.stream0 = stream ? 1 : 0;
        return true;
    };
    static const bool stream_ok = [] { const char* e = getenv("GEM_CHAIN_STREAM"); return !e || e[0] != '0'; }();
    const bool want_stream = stream_ok && !L.energy && w[0]->Kp / 64 == 4;
    if (!(want_stream && plan(2, true))) GEM_REQUIRE(plan(4, false), "no free activation buffer for the chain");
    a.nl = np, a.W = L.W, a.T = L.T, a.wpq = wpq, a.out_plain = L.out_lo ? nullptr : (float*)L.out_hi;
    a.dbg = g_tap_dbg;
    a.status = L.status;
    Chain2Energy en;
    memset(&en, 0, sizeof(en));
    if (L.energy) {
        const ChainEnergyLaunch& E = *L.energy;
        GEM_REQUIRE(tap_chain_energy_supported(L.T, E.J, w[0]->Kp) && E.J * 3 <= L.Kreal,
                    "energy prologue: window geometry not supported by the fused chain");
        GEM_REQUIRE(E.skel && E.skel->num_joints == E.J && E.pose && E.pose0 && E.clip && E.mean_bone && E.energy && E.row_exp,
                    "energy prologue: missing argument");
        GEM_REQUIRE(E.wt.reproj == 0.f || (E.cam && E.heat && E.frame_base), "energy prologue: camera / heat maps required");
        en.enabled = 1;
        en.in_buf = a.L[0].in_buf[0];
        int nfree = 0, free_buf[4];
        for (int bfr = 0; bfr < 4; ++bfr)
            if (bfr != en.in_buf) free_buf[nfree++] = bfr;
        en.scratch0 = free_buf[0], en.scratch1 = free_buf[1];
        en.pose = E.pose, en.pose0 = E.pose0, en.energy = E.energy, en.row_exp = E.row_exp;
        en.heat = E.heat, en.frame_base = E.frame_base, en.clip = E.clip, en.mean_bone = E.mean_bone, en.status = E.status;
        en.W = L.W, en.T = L.T, en.J = E.J, en.H = E.H, en.Wd = E.Wd, en.planar = E.planar;
        en.w3d = E.wt.w3d, en.ws = E.wt.smooth, en.wb = E.wt.bone, en.wv = E.wt.vae, en.wr = E.wt.reproj;
        en.patch = E.patch && E.patch_valid ? E.patch : nullptr;
        en.patch_origin = en.patch ? E.patch_origin : nullptr, en.patch_valid = en.patch ? E.patch_valid : nullptr;
        en.patch_stats = en.patch ? E.patch_stats : nullptr;
        if (E.cam) en.cam = *E.cam;
        en.skel = *E.skel;
    }
    static PerDeviceOnce attr_set;
    if (bool* once_ = attr_set.flag(); !*once_) {
        GEM_CUDA(cudaFuncSetAttribute(tc_tap_chain2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChain2Smem));
        GEM_CUDA(cudaFuncSetAttribute(tc_tap_chain2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChain2Smem));
        *once_ = true;
    }
    const int tiles = (L.W + 4 * wpq - 1) / (4 * wpq);
    const int grid = 2 * ((tiles + 1) / 2);
    if (en.enabled) tc_tap_chain2_kernel<true><<<grid, kChainThreads, kChain2Smem, stream>>>(maps, a, en);
    else tc_tap_chain2_kernel<false><<<grid, kChainThreads, kChain2Smem, stream>>>(maps, a, en);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

// frees the prepared slabs of one layer's weights (every scheme); the caller has synchronised the device
void tc_tap_forget_weight(void* owner, const float* B) {
    TapState* st = state_of(owner);
    for (int scheme = 1; scheme <= 2; ++scheme) {
        auto it = st->weights.find(std::make_pair(B, scheme));
        if (it == st->weights.end()) continue;
        cudaFree(it->second.hi), cudaFree(it->second.lo);
        if (it->second.hs) cudaFree(it->second.hs);
        st->weights.erase(it);
    }
}

void tc_tap_release(void* owner) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_states.find(owner);
    if (it == g_states.end()) return;
    TapState* st = it->second;
    for (auto& kv : st->weights) {
        cudaFree(kv.second.hi), cudaFree(kv.second.lo);
        if (kv.second.hs) cudaFree(kv.second.hs);
    }
    delete st;
    g_states.erase(it);
}

}  // namespace gem
