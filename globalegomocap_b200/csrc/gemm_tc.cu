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
// Kernel anatomy (one 128 x BN output tile per CTA, 576 threads):
//   warp 0   TMA producer: per 32-wide K block four cp.async.bulk.tensor loads (A_hi, A_lo, B_hi, B_lo;
//            128 rows x 128 B each, SWIZZLE_128B) into a 3-stage shared-memory ring, mbarrier full/empty
//   warp 1   allocates all 512 TMEM columns (4 accumulators); one elected lane issues 12 tcgen05.mma.kind::tf32 (M128 N128 K8)
//            per stage and tcgen05.commit's the stage back to the producer
//   warps 2-17 epilogue (four per TMEM lane quarter): tcgen05.ld the accumulators (32 lanes x 16 columns per instruction), bias /
//            LeakyReLU, 128-byte row segments straight to global memory
// Weights are transposed to K-major [N][K] and split once per checkpoint; activations are split by a
// small elementwise kernel before each GEMM, or arrive already split from the producing kernel.
//
// Second scheme, same kernel (template parameter F16), twice the tensor rate and half the operand bytes:
// x ~ hi + 2^-11 lo with hi = fp16(x), lo = fp16((x - hi) 2^11) (tc_common.cuh: split_f16), the products issued
// as tcgen05.mma.kind::f16 (K = 16):  hi*hi into the main accumulators,  hi*lo + lo*hi into a separate "cross"
// accumulator that the epilogue scales by 2^-11.  22 significant bits for |x| in [2^-14, 65504], an absolute
// 2^-35 below; callers whose rows can be tiny (gradients) pre-scale them by a power of two and pass the
// exponents (TapGemmArgs::row_exp) so that the epilogue undoes it exactly.
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "tc_common.cuh"

namespace gem {

using namespace tc;

namespace {

constexpr int BM = 128;
constexpr int kStages = 3;
constexpr int kATileBytes = BM * 128;             // 16 KB: 128 rows of one 128-byte swizzle row (32 fp32 or 64 fp16)
constexpr int kEpiSub = 4;                        // epilogue warps per TMEM lane quarter
constexpr int kThreads = 64 + 4 * 32 * kEpiSub;   // TMA warp, MMA warp, 16 epilogue warps
constexpr int kTmemCols = 512;                    // all of TMEM: up to four BN-column accumulators

// The N tile is a template parameter chosen per launch (launch_tap_gemm_tc): with ~15 M tiles the tile count
// of a 128-wide N tile lands just above a multiple of the 148 SMs (300 tiles = 2.03 waves for the latent ->
// T*256 layer); 160 / 112 / 144 columns bring it under the wave boundary.
template <int BN>
struct GemmCfg {
    static constexpr int kBTileBytes = BN * 128;
    static constexpr int kStageBytes = 2 * kATileBytes + 2 * kBTileBytes;      // A_hi, A_lo, B_hi, B_lo
    static constexpr int kAcc = (kTmemCols / BN) < 4 ? (kTmemCols / BN) : 4;   // independent fp32 accumulators
    static constexpr int kMain16 = (kAcc - 1) < 2 ? (kAcc - 1) : 2;            // fp16 scheme: main accumulators ...
    static constexpr int kCross16 = kMain16;                                   // ... and the index of the cross one
    static constexpr int kPitch = BN + 4;                                      // staging row pitch (floats)
    static constexpr int kPitch16 = BN + 8;                                    // ... of fp16 outputs (elements)
    static constexpr int kStageOut = 32 * kPitch * 4;                          // one warp's staged 32 x BN block
    static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(BN % 16 == 0 && BN <= 256, "UMMA N");
    static_assert(kBTileBytes % 1024 == 0, "B tile must keep the swizzle atom alignment");
    static_assert(8 * kStageOut <= kStages * kStageBytes, "epilogue staging must fit in the pipeline's memory");
};

struct TcArgs {
    const float* bias;
    float* C;        // result, or its TF32 hi part when C_lo is set
    float* C_lo;     // optional: the epilogue writes the result already split for the next tensor-core layer
    uint32_t* sign;  // optional: packed sign bits of the result, [M][N/32] (needs BN % 32 == 0)
    const int32_t* row_exp;   // optional: row m of A was scaled by 2^row_exp[m]; the result row is scaled back
    int out16;                // C / C_lo are fp16 hi / scaled fp16 lo arrays (uint16), not TF32 hi / lo floats
    uint32_t* row_flag;       // optional [M]: GEM_WIN_F16_RANGE is OR-ed in when a row's fp16 output saturates
    int M, N, K, ldc, epi;
    long long* dbg;           // debug: per-CTA phase timestamps (gem_debug_gemm_timestamps), NULL in production
};


// debug timestamps: slot s of CTA (blockIdx.y * gridDim.x + blockIdx.x), 16 slots per CTA
#define GEMM_TS(slot)                                                                                   \
    do {                                                                                                \
        if (g.dbg) g.dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = clock64();      \
    } while (0)
#define GEMM_TS_ADD(slot, v)                                                                            \
    do {                                                                                                \
        if (g.dbg) g.dbg[(size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 16 + (slot)] = (v);            \
    } while (0)

// Epilogue of one 128 x BN accumulator tile, run by kEpiWarps = 16 warps: a warp can only read the TMEM lane
// quarter (warp % 4), so four warps share each quarter and deal its 16-column chunks round-robin.  (With one warp
// per quarter the epilogue was a single dependent instruction stream per scheduler — ~460 instructions per chunk
// at an IPC of 0.2 — and took 23k of a tile's 63k cycles; per-CTA timestamps, tools/dbg_gemm.py.)
// The accumulators' sum, bias and activation are formed one row per thread, staged in the (now idle) pipeline
// memory and, after a named barrier over the quarter's four warps, written out along the rows: 16-byte vectors,
// whole row segments per instruction instead of 32 row segments of 16 bytes.
template <int BN, bool F16>
__device__ __forceinline__ void epilogue_tile(const TcArgs& g, uint8_t* smem, uint32_t tmem_base, int m0, int n0, int warp,
                                              int lane, int num_kb) {
    using C = GemmCfg<BN>;
    constexpr int kAcc = C::kAcc;
    const int q = warp & 3;                         // TMEM lane quarter this warp may read
    const int sub = (warp - 2) >> 2;                // 0 .. kEpiSub-1: which of the quarter's warps
    const int m = m0 + q * 32 + lane;
    // staging regions, as small as the output format allows (the CTA-pair kernel keeps fewer pipeline bytes)
    const size_t region = g.out16 ? (size_t)32 * C::kPitch16 * 2 : (size_t)C::kStageOut;
    float* stage_hi = reinterpret_cast<float*>(smem + (size_t)(g.C_lo ? q * 2 : q) * region);
    float* stage_lo = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stage_hi) + region);
    const int nmain = F16 ? C::kMain16 : kAcc;
    const int nacc = num_kb < nmain ? num_kb : nmain;
    float row_scale = 1.f;
    if (g.row_exp && m < g.M) row_scale = exp2f(-(float)g.row_exp[m]);      // exact: a power of two
#pragma unroll 1
    for (int c = sub; c < BN / 16; c += kEpiSub) {
        const int nb = n0 + c * 16;
        if (nb >= g.N) break;                       // N % 16 == 0: a chunk is inside the matrix or outside
        uint32_t v[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 16);
        if (F16) {
            // main accumulator(s) and the cross accumulator in flight together, one wait
            uint32_t t1[16], tx[16];
            tmem_ld_32x32b_x16(taddr, v);
            if (C::kMain16 > 1 && nacc > 1) tmem_ld_32x32b_x16(taddr + (uint32_t)BN, t1);
            tmem_ld_32x32b_x16(taddr + (uint32_t)(C::kCross16 * BN), tx);
            tmem_ld_wait();
            if (C::kMain16 > 1 && nacc > 1) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(t1[j]));
            }
#pragma unroll
            for (int j = 0; j < 16; ++j)      // + 2^-11 (hi*lo + lo*hi)
                v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(tx[j]) * (1.f / kF16LoScale));
        } else {
            tmem_ld_32x32b_x16(taddr, v);
            tmem_ld_wait();
            for (int a = 1; a < nacc; ++a) {
                uint32_t t[16];
                tmem_ld_32x32b_x16(taddr + (uint32_t)(a * BN), t);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(t[j]));
            }
        }
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) o[j] = __uint_as_float(v[j]) * row_scale;
        if (g.bias) {
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
                const float4 bv = __ldg(reinterpret_cast<const float4*>(g.bias + nb + j));
                o[j] += bv.x, o[j + 1] += bv.y, o[j + 2] += bv.z, o[j + 3] += bv.w;
            }
        }
        if (g.epi == EPI_LRELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = o[j] > 0.f ? o[j] : o[j] * 0.01f;
        }
        if (g.sign) {                               // bit (n % 32) of word [m][n / 32]: this chunk is one 16-bit half
            uint32_t sbits = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) sbits |= (o[j] > 0.f ? 1u : 0u) << j;
            if (m < g.M) reinterpret_cast<uint16_t*>(g.sign)[(size_t)m * (g.N >> 4) + (nb >> 4)] = (uint16_t)sbits;
        }
        if (g.out16) {
            if (g.row_flag) {                       // fp16 range check (the conversions below saturate silently)
                float amax = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) amax = fmaxf(amax, fabsf(o[j]));
                if (amax > 65504.f && m < g.M) atomicOr(g.row_flag + m, GEM_WIN_F16_RANGE);
            }
            uint16_t* sh_row = reinterpret_cast<uint16_t*>(stage_hi) + lane * C::kPitch16 + c * 16;
            uint16_t* sl_row = reinterpret_cast<uint16_t*>(stage_lo) + lane * C::kPitch16 + c * 16;
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                uint16_t h0, l0, h1, l1;
                split_f16(o[j], h0, l0), split_f16(o[j + 1], h1, l1);
                *reinterpret_cast<uint32_t*>(sh_row + j) = (uint32_t)h0 | ((uint32_t)h1 << 16);
                *reinterpret_cast<uint32_t*>(sl_row + j) = (uint32_t)l0 | ((uint32_t)l1 << 16);
            }
        } else {
            float* sh_row = stage_hi + lane * C::kPitch + c * 16;
            if (g.C_lo) {
                float* sl_row = stage_lo + lane * C::kPitch + c * 16;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 h, l;
                    split_tf32(o[j + 0], h.x, l.x), split_tf32(o[j + 1], h.y, l.y);
                    split_tf32(o[j + 2], h.z, l.z), split_tf32(o[j + 3], h.w, l.w);
                    *reinterpret_cast<float4*>(sh_row + j) = h;
                    *reinterpret_cast<float4*>(sl_row + j) = l;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    *reinterpret_cast<float4*>(sh_row + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            }
        }
    }
    // the quarter's four warps have staged their chunks of the same 32 rows
    asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * kEpiSub) : "memory");
    // coalesced write-out of the quarter's 32 x BN block in 16-byte vectors along the rows
    const int rows = min(32, g.M - (m0 + q * 32));
    const int tq = sub * 32 + lane;                 // thread index within the quarter's warps
    if (g.out16) {
        constexpr int kV = BN / 8;                                    // 8 halves per vector (N, ldc % 8 == 0)
        const int ncols = min(kV, (g.N - n0) / 8);
        const uint16_t* sh = reinterpret_cast<const uint16_t*>(stage_hi);
        const uint16_t* sl = reinterpret_cast<const uint16_t*>(stage_lo);
        for (int i = tq; i < rows * kV; i += 32 * kEpiSub) {
            const int r = i / kV, cv = i - r * kV;
            if (cv >= ncols) continue;
            const size_t off = (size_t)(m0 + q * 32 + r) * g.ldc + n0 + cv * 8;
            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(g.C) + off) =
                *reinterpret_cast<const uint4*>(sh + r * C::kPitch16 + cv * 8);
            *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(g.C_lo) + off) =
                *reinterpret_cast<const uint4*>(sl + r * C::kPitch16 + cv * 8);
        }
    } else {
        constexpr int kV4 = BN / 4;
        const int ncols4 = min(kV4, (g.N - n0) / 4);
        for (int i = tq; i < rows * kV4; i += 32 * kEpiSub) {
            const int r = i / kV4, c4 = i - r * kV4;
            if (c4 >= ncols4) continue;
            const size_t off = (size_t)(m0 + q * 32 + r) * g.ldc + n0 + c4 * 4;
            *reinterpret_cast<float4*>(g.C + off) = *reinterpret_cast<const float4*>(stage_hi + r * C::kPitch + c4 * 4);
            if (g.C_lo)
                *reinterpret_cast<float4*>(g.C_lo + off) = *reinterpret_cast<const float4*>(stage_lo + r * C::kPitch + c4 * 4);
        }
    }
}

template <int BN, bool F16>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, TcArgs g) {
    using C = GemmCfg<BN>;
    constexpr int kAcc = C::kAcc;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * C::kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    constexpr int BK = F16 ? 64 : 32;             // elements per 128-byte row
    const int num_kb = g.K / BK;
    if (threadIdx.x == 0 && g.dbg) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        GEMM_TS_ADD(14, (long long)gt);
        GEMM_TS(0);
    }

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
    if (threadIdx.x == 0) GEMM_TS(1);

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            long long wait_empty = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t parity = ((kb / kStages) & 1) ^ 1;
                const long long t0 = g.dbg ? clock64() : 0;
                mbar_wait(&empty_bar[s], parity);
                if (g.dbg) wait_empty += clock64() - t0;
                uint8_t* st = smem + s * C::kStageBytes;
                mbar_arrive_expect_tx(&full_bar[s], C::kStageBytes);
                tma_load_2d(st, &map_a_hi, kb * BK, m0, &full_bar[s]);
                tma_load_2d(st + kATileBytes, &map_a_lo, kb * BK, m0, &full_bar[s]);
                tma_load_2d(st + 2 * kATileBytes, &map_b_hi, kb * BK, n0, &full_bar[s]);
                tma_load_2d(st + 2 * kATileBytes + C::kBTileBytes, &map_b_lo, kb * BK, n0, &full_bar[s]);
            }
            GEMM_TS_ADD(8, wait_empty);
            GEMM_TS(2);
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            constexpr uint32_t idesc = F16 ? instr_desc_f16(BM, BN) : instr_desc_tf32(BM, BN);
            long long wait_full = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const long long t0 = g.dbg ? clock64() : 0;
                mbar_wait(&full_bar[s], (kb / kStages) & 1);
                if (g.dbg) wait_full += clock64() - t0;
                if (kb == 0) GEMM_TS(3);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + s * C::kStageBytes);
                const uint64_t a_hi = make_smem_desc(base), a_lo = make_smem_desc(base + kATileBytes);
                const uint64_t b_hi = make_smem_desc(base + 2 * kATileBytes);
                const uint64_t b_lo = make_smem_desc(base + 2 * kATileBytes + C::kBTileBytes);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);        // 32 bytes per MMA K step inside the 128-B row
                    if (F16) {
                        const uint32_t cross = tmem_base + (uint32_t)(C::kCross16 * BN);
                        const uint32_t acc = tmem_base + (uint32_t)((kb % C::kMain16) * BN);
                        umma_f16(cross, a_lo + adv, b_hi + adv, idesc, (kb != 0) || (k != 0));
                        umma_f16(cross, a_hi + adv, b_lo + adv, idesc, 1);
                        umma_f16(acc, a_hi + adv, b_hi + adv, idesc, (kb >= C::kMain16) || (k != 0));
                    } else {
                        // small terms first, then the dominant one
                        const uint32_t acc = tmem_base + (uint32_t)((kb % kAcc) * BN);
                        umma_tf32(acc, a_lo + adv, b_hi + adv, idesc, (kb >= kAcc) || (k != 0));
                        umma_tf32(acc, a_hi + adv, b_lo + adv, idesc, 1);
                        umma_tf32(acc, a_hi + adv, b_hi + adv, idesc, 1);
                    }
                }
                umma_commit(&empty_bar[s]);          // frees the smem stage when these MMAs have read it
            }
            umma_commit(tmem_full_bar);              // accumulators complete
            GEMM_TS_ADD(9, wait_full);
            GEMM_TS(4);
        }
    } else {
        // ===== epilogue warps 2..17 =====
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        if (threadIdx.x == 64) GEMM_TS(5);
        epilogue_tile<BN, F16>(g, smem, tmem_base, m0, n0, warp, lane, num_kb);
        if (threadIdx.x == 64) GEMM_TS(6);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
    if (threadIdx.x == 0 && g.dbg) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        GEMM_TS_ADD(15, (long long)gt);
        GEMM_TS(7);
    }
}

// ---- CTA-pair variant of the fp16 scheme (tcgen05.mma.cta_group::2, M = 256) --------------------------------------
// The 1-CTA kernel is bound by shared-memory bandwidth, not by the tensor pipe: an M128 x N160 x K16 MMA reads
// 4 KB of A and 5 KB of B from shared memory every 80 cycles (112 B/clk of the SM's 128) while TMA writes the next
// stage at 75 B/clk on top.  Two CTAs of one cluster (the two SMs of a TPC) run ONE M256 MMA: each keeps its own
// 128 rows of A but only HALF of the B tile, the hardware feeds both tensor cores from the two halves, so per SM
// the MMA reads 6.5 KB per 80 cycles and TMA writes 52 KB instead of 72 KB per K block (and L2 -> SM traffic drops
// by the same 28 %).  Protocol: both CTAs' producers load into their own shared memory but signal the LEADER's
// full barrier (which expects both CTAs' bytes); the leader's MMA thread issues for the pair and multicasts its
// commits to the empty / accumulator barriers of both CTAs; each CTA's epilogue drains its own TMEM.
#ifndef GEM_PAIR_STAGES
#define GEM_PAIR_STAGES 4
#endif
template <int BN>
struct PairCfg {
    static constexpr int kStages = GEM_PAIR_STAGES;
    static constexpr int kBHalfBytes = (BN / 2) * 128;
    static constexpr int kStageBytes = 2 * kATileBytes + 2 * kBHalfBytes;      // A_hi, A_lo, half of B_hi, half of B_lo
    static constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
    static_assert(BN % 16 == 0 && BN <= 256, "UMMA N (M = 256 needs N % 16 == 0; halves keep the 8-row swizzle atoms)");
    static_assert(kBHalfBytes % 1024 == 0, "B half tile must keep the swizzle atom alignment");
    // the pair kernel writes fp16 hi/lo pairs or plain fp32 (never fp32 hi/lo pairs): at most 8 fp16 or 4 fp32 regions
    static_assert(8 * 32 * GemmCfg<BN>::kPitch16 * 2 <= kStages * kStageBytes && 4 * GemmCfg<BN>::kStageOut <= kStages * kStageBytes,
                  "epilogue staging must fit in the pipeline's memory");
    static_assert(kSmemBytes <= 227 * 1024, "shared memory");
};

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tc_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                    const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, TcArgs g) {
    using C = GemmCfg<BN>;
    using P = PairCfg<BN>;
    constexpr int kSt = P::kStages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kSt * P::kStageBytes);
    uint64_t* empty_bar = full_bar + kSt;
    uint64_t* tmem_full_bar = empty_bar + kSt;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();                 // 0 = leader; the pair covers rows [m0 - rank*BM, +256)
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    constexpr int BK = 64;
    const int num_kb = g.K / BK;
    if (threadIdx.x == 0 && g.dbg) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        GEMM_TS_ADD(14, (long long)gt);
        GEMM_TS(0);
    }

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a_hi), prefetch_tmap(&map_a_lo), prefetch_tmap(&map_b_hi), prefetch_tmap(&map_b_lo);
        for (int s = 0; s < kSt; ++s) mbar_init(&full_bar[s], 1), mbar_init(&empty_bar[s], 1);
        mbar_init(tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_pair(tmem_slot, kTmemCols);
    tc_fence_before();
    cluster_sync_all();                 // the peer's barriers exist before anything arrives on them remotely
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) GEMM_TS(1);

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            const int nb = n0 + (int)rank * (BN / 2);
            long long wait_empty = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kSt;
                const uint32_t parity = ((kb / kSt) & 1) ^ 1;
                const long long t0 = g.dbg ? clock64() : 0;
                mbar_wait(&empty_bar[s], parity);
                if (g.dbg) wait_empty += clock64() - t0;
                uint8_t* st = smem + s * P::kStageBytes;
                if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * P::kStageBytes);
                const uint32_t lead_bar = mapa_u32(smem_u32(&full_bar[s]), 0);
                tma_load_2d_pair(st, &map_a_hi, kb * BK, m0, lead_bar);
                tma_load_2d_pair(st + kATileBytes, &map_a_lo, kb * BK, m0, lead_bar);
                tma_load_2d_pair(st + 2 * kATileBytes, &map_b_hi, kb * BK, nb, lead_bar);
                tma_load_2d_pair(st + 2 * kATileBytes + P::kBHalfBytes, &map_b_lo, kb * BK, nb, lead_bar);
            }
            GEMM_TS_ADD(8, wait_empty);
            GEMM_TS(2);
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = instr_desc_f16(2 * BM, BN);
            const uint32_t cross = tmem_base + (uint32_t)(C::kCross16 * BN);
            long long wait_full = 0;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kSt;
                const long long t0 = g.dbg ? clock64() : 0;
                mbar_wait(&full_bar[s], (kb / kSt) & 1);
                if (g.dbg) wait_full += clock64() - t0;
                if (kb == 0) GEMM_TS(3);
                tc_fence_after();
                const uint32_t base = smem_u32(smem + s * P::kStageBytes);
                const uint64_t a_hi = make_smem_desc(base), a_lo = make_smem_desc(base + kATileBytes);
                const uint64_t b_hi = make_smem_desc(base + 2 * kATileBytes);
                const uint64_t b_lo = make_smem_desc(base + 2 * kATileBytes + P::kBHalfBytes);
                const uint32_t acc = tmem_base + (uint32_t)((kb % C::kMain16) * BN);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);        // 32 bytes per MMA K step inside the 128-B row
                    umma_f16_pair(cross, a_lo + adv, b_hi + adv, idesc, (kb != 0) || (k != 0));
                    umma_f16_pair(cross, a_hi + adv, b_lo + adv, idesc, 1);
                    umma_f16_pair(acc, a_hi + adv, b_hi + adv, idesc, (kb >= C::kMain16) || (k != 0));
                }
                umma_commit_pair(&empty_bar[s], 3);      // frees the stage in both CTAs
            }
            umma_commit_pair(tmem_full_bar, 3);          // accumulators complete in both CTAs
            GEMM_TS_ADD(9, wait_full);
            GEMM_TS(4);
        }
    } else {
        // ===== epilogue warps 2..17 (both CTAs, each its own 128 rows) =====
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        if (threadIdx.x == 64) GEMM_TS(5);
        epilogue_tile<BN, true>(g, smem, tmem_base, m0, n0, warp, lane, num_kb);
        if (threadIdx.x == 64) GEMM_TS(6);
    }
    tc_fence_before();
    cluster_sync_all();                 // neither CTA leaves (or frees TMEM) while the pair's MMAs / reads are in flight
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
    if (threadIdx.x == 0 && g.dbg) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
        GEMM_TS_ADD(15, (long long)gt);
        GEMM_TS(7);
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

// the fp16 scheme's split (tc_common.cuh: split_f16), optionally of rows pre-scaled by 2^row_exp[m]
__global__ void split_f16_kernel(const float* __restrict__ A, int lda, int M, int K, const int32_t* __restrict__ row_exp,
                                 uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, uint32_t* __restrict__ row_flag) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one float4 per thread
    const int kv = K / 4;
    if (i >= (size_t)M * kv) return;
    const int m = (int)(i / kv), k = (int)(i - (size_t)m * kv) * 4;
    float4 x = *reinterpret_cast<const float4*>(A + (size_t)m * lda + k);
    if (row_exp) {
        const float sc = exp2f((float)row_exp[m]);
        x.x *= sc, x.y *= sc, x.z *= sc, x.w *= sc;
    }
    if (row_flag && fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))) > 65504.f)
        atomicOr(row_flag + m, GEM_WIN_F16_RANGE);
    uint16_t h[4], l[4];
    split_f16(x.x, h[0], l[0]), split_f16(x.y, h[1], l[1]), split_f16(x.z, h[2], l[2]), split_f16(x.w, h[3], l[3]);
    *reinterpret_cast<uint2*>(hi + (size_t)m * K + k) =
        make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    *reinterpret_cast<uint2*>(lo + (size_t)m * K + k) =
        make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}

__global__ void transpose_split_f16_kernel(const float* __restrict__ B, int ldb, int K, int N, uint16_t* __restrict__ hi,
                                           uint16_t* __restrict__ lo) {
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
            uint16_t h, l;
            split_f16(tile[threadIdx.x][r], h, l);
            hi[(size_t)n * K + k] = h;
            lo[(size_t)n * K + k] = l;
        }
    }
}

// One CTA per row: x = hi (+ lo) in fp32  ->  the row's power-of-two scale 2^e that brings max|x| to ~2^10, and
// the fp16 split of x 2^e.  Gradient rows can be arbitrarily small (1e-7 near convergence); unscaled they would
// fall into fp16's subnormals and keep only an absolute 2^-35.
__global__ void __launch_bounds__(256) rowscale_split_f16_kernel(const float* __restrict__ xh, const float* __restrict__ xl,
                                                                 int K, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                                 int32_t* __restrict__ row_exp) {
    __shared__ float red[8];
    const size_t row = blockIdx.x;
    const float* ph = xh + row * K;
    const float* pl = xl ? xl + row * K : nullptr;
    float mx = 0.f;
    for (int k = threadIdx.x * 4; k < K; k += 256 * 4) {
        float4 v = *reinterpret_cast<const float4*>(ph + k);
        if (pl) {
            const float4 w = *reinterpret_cast<const float4*>(pl + k);
            v.x += w.x, v.y += w.y, v.z += w.z, v.w += w.w;
        }
        mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
    int e = 0;
    if (mx > 0.f && mx < 3.0e38f) {
        e = 10 - ilogbf(mx);
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
    }
    const float sc = exp2f((float)e);
    if (threadIdx.x == 0) row_exp[row] = e;
    for (int k = threadIdx.x * 4; k < K; k += 256 * 4) {
        float4 v = *reinterpret_cast<const float4*>(ph + k);
        if (pl) {
            const float4 w = *reinterpret_cast<const float4*>(pl + k);
            v.x += w.x, v.y += w.y, v.z += w.z, v.w += w.w;      // hi + lo is exact
        }
        uint16_t h[4], l[4];
        split_f16(v.x * sc, h[0], l[0]), split_f16(v.y * sc, h[1], l[1]);
        split_f16(v.z * sc, h[2], l[2]), split_f16(v.w * sc, h[3], l[3]);
        *reinterpret_cast<uint2*>(hi + row * K + k) =
            make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
        *reinterpret_cast<uint2*>(lo + row * K + k) =
            make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
    }
}

// ---------------------------------------------------------------- host side
// rows x K row-major (pitch in elements), box = box_rows x one 128-byte row, 128-byte swizzle, OOB rows read zero
int make_map(CUtensorMap* map, const void* base, bool f16, uint64_t rows, uint64_t K, uint64_t pitch = 0,
             int box_rows = BM) {
    const uint64_t esz = f16 ? 2 : 4;
    const uint64_t dims[2] = {K, rows};
    const uint64_t strides[1] = {(pitch ? pitch : K) * esz};
    const uint32_t box[2] = {(uint32_t)(128 / esz), (uint32_t)box_rows};
    return f16 ? make_map_u16(map, base, 2, dims, strides, box)
               : make_map_f32(map, static_cast<const float*>(base), 2, dims, strides, box);
}

constexpr int kNumBN = 6;
constexpr int kBNs[kNumBN] = {64, 96, 112, 128, 144, 160};      // (64 and 96: CTA-pair kernel only, small batches)
constexpr int kFirstWideBN = 2;
struct WeightSplit {
    void *hi = nullptr, *lo = nullptr;                 // K-major [N][K], fp32 (TF32 scheme) or fp16
    int K = 0, N = 0;
    CUtensorMap map_hi[kNumBN], map_lo[kNumBN];       // one box height per N-tile width
    CUtensorMap map_hi2[kNumBN], map_lo2[kNumBN];     // half-height boxes: each CTA of a pair loads half of the B tile
};
struct KeyHash {
    size_t operator()(const std::pair<const float*, int>& k) const {
        return std::hash<const void*>()(k.first) ^ ((size_t)k.second * 0x9E3779B97F4A7C15ull);
    }
};
struct TcState {
    std::unordered_map<std::pair<const float*, int>, WeightSplit, KeyHash> weights;   // (caller's weight pointer, scheme)
    // split copies of activations that arrive un-split (the layer-by-layer fallbacks, gem_gemm): one pair of buffers per
    // STREAM, because a ctx's slices run concurrently on their own streams and each may be in one of these launches
    struct Scratch {
        void *hi = nullptr, *lo = nullptr;
        size_t capacity = 0;                                  // bytes in each of hi / lo
    };
    std::unordered_map<cudaStream_t, Scratch> scratch;
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

static int make_map_any(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                        const uint64_t* strides_bytes, const uint32_t* box) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return GEM_ERR_CUDA;
    }
    cuuint64_t d[5], st[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; ++i) d[i] = dims[i], bx[i] = box[i], estr[i] = 1;
    for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
    CUresult r = enc(map, dt, (cuuint32_t)rank, const_cast<void*>(base), d, st, bx, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
        return GEM_ERR_CUDA;
    }
    return GEM_OK;
}

int make_map_f32(CUtensorMap* map, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box) {
    return make_map_any(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box);
}
int make_map_u16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                 const uint32_t* box) {
    return make_map_any(map, CU_TENSOR_MAP_DATA_TYPE_UINT16, base, rank, dims, strides_bytes, box);
}

}  // namespace tc

bool tc_gemm_available() { return true; }
long long* g_gemm_dbg = nullptr;   // set by gem_debug_gemm_timestamps (profiling only)
int g_gemm_pair = -1;      // debug override of GEM_GEMM_PAIR (gem_debug_gemm_pair): -1 = environment / default

// `owner` identifies the ctx; scratch grows on demand and lives until tc_gemm_release(owner).
template <int BN, bool F16>
static int launch_bn(cudaStream_t stream, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const WeightSplit& w, int idx,
                     const TcArgs& a) {
    static PerDeviceOnce attr_set;
    if (bool* once_ = attr_set.flag(); !*once_) {
        GEM_CUDA(cudaFuncSetAttribute(tc_gemm_kernel<BN, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)GemmCfg<BN>::kSmemBytes));
        *once_ = true;
    }
    dim3 grid((a.M + BM - 1) / BM, (a.N + BN - 1) / BN);
    tc_gemm_kernel<BN, F16><<<grid, kThreads, GemmCfg<BN>::kSmemBytes, stream>>>(a_hi, a_lo, w.map_hi[idx], w.map_lo[idx], a);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}
template <int BN>
static int launch_pair_bn(cudaStream_t stream, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const WeightSplit& w, int idx,
                          const TcArgs& a) {
    static PerDeviceOnce attr_set;
    if (bool* once_ = attr_set.flag(); !*once_) {
        GEM_CUDA(cudaFuncSetAttribute(tc_gemm_pair_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)PairCfg<BN>::kSmemBytes));
        *once_ = true;
    }
    dim3 grid(2 * ((a.M + 2 * BM - 1) / (2 * BM)), (a.N + BN - 1) / BN);      // whole pairs along M
    tc_gemm_pair_kernel<BN><<<grid, kThreads, PairCfg<BN>::kSmemBytes, stream>>>(a_hi, a_lo, w.map_hi2[idx], w.map_lo2[idx], a);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}
template <bool F16>
static int launch_scheme(int bn, cudaStream_t stream, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const WeightSplit& w,
                         int idx, const TcArgs& a) {
    switch (bn) {
        case 112: return launch_bn<112, F16>(stream, a_hi, a_lo, w, idx, a);
        case 144: return launch_bn<144, F16>(stream, a_hi, a_lo, w, idx, a);
        case 160: return launch_bn<160, F16>(stream, a_hi, a_lo, w, idx, a);
        default: return launch_bn<128, F16>(stream, a_hi, a_lo, w, idx, a);
    }
}

static TcState* state_of(void* owner) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_states.find(owner);
    if (it == g_states.end()) it = g_states.emplace(owner, new TcState()).first;
    return it->second;
}

// scheme: 1 = 3xTF32 (fp32 hi / lo operands), 2 = fp16 hi + scaled fp16 lo operands
int launch_tap_gemm_tc(cudaStream_t stream, const TapGemmArgs& g, void* owner, size_t scheme) {
    if (g.M <= 0) return GEM_OK;
    const bool f16 = scheme == 2;
    const int bk = f16 ? 64 : 32;
    const size_t esz = f16 ? 2 : 4;
    GEM_REQUIRE(g.taps == 1, "tcgen05 path handles plain GEMMs only");
    GEM_REQUIRE(g.K % bk == 0 && g.N % 128 == 0, "K must be a multiple of the 128-byte K block and N of 128");
    GEM_REQUIRE(g.lda % 8 == 0 && g.ldc % 4 == 0, "lda must be a multiple of 8, ldc of 4");
    GEM_REQUIRE(g.epi == EPI_NONE || g.epi == EPI_LRELU, "unsupported epilogue on the tcgen05 path");
    TcState* st = state_of(owner);
    // weights: K-major hi/lo copies registered by tc_gemm_prepare_weight
    auto wit = st->weights.find(std::make_pair(g.B, (int)(f16 ? 2 : 1)));
    if (wit == st->weights.end() || wit->second.K != g.K || wit->second.N != g.N) {
        set_error("tcgen05 path: weight matrix was not prepared (tc_gemm_prepare_weight)");
        return GEM_ERR_STATE;
    }
    // activations: already split by the producing kernel, or split here into the ctx-owned scratch
    const void *a_hi = g.A_hi, *a_lo = g.A_lo;
    uint64_t pitch = (uint64_t)g.lda;
    if (!a_hi) {
        const size_t need = (size_t)g.M * g.K * esz;
        TcState::Scratch* sc;
        {
            std::lock_guard<std::mutex> lk(g_mu);
            sc = &st->scratch[stream];                        // (references into an unordered_map stay valid)
        }
        if (need > sc->capacity) {
            GEM_CUDA(cudaStreamSynchronize(stream));
            if (sc->hi) cudaFree(sc->hi), cudaFree(sc->lo);
            sc->hi = sc->lo = nullptr, sc->capacity = 0;
            GEM_CUDA(cudaMalloc(&sc->hi, need));
            GEM_CUDA(cudaMalloc(&sc->lo, need));
            sc->capacity = need;
        }
        const size_t n4 = (size_t)g.M * (g.K / 4);
        if (f16)
            split_f16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(g.A, g.lda, g.M, g.K, g.row_exp,
                                                                            (uint16_t*)sc->hi, (uint16_t*)sc->lo, nullptr);
        else
            split_tf32_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(g.A, g.lda, g.M, g.K, (float*)sc->hi,
                                                                             (float*)sc->lo);
        GEM_CHECK_LAUNCH();
        a_hi = sc->hi, a_lo = sc->lo, pitch = (uint64_t)g.K;
    }
    CUtensorMap map_a_hi, map_a_lo;
    int rc = make_map(&map_a_hi, a_hi, f16, (uint64_t)g.M, (uint64_t)g.K, pitch);
    if (rc == GEM_OK) rc = make_map(&map_a_lo, a_lo, f16, (uint64_t)g.M, (uint64_t)g.K, pitch);
    if (rc != GEM_OK) return rc;
    TcArgs a;
    a.bias = g.bias, a.C = g.C, a.C_lo = (float*)g.C_lo, a.sign = g.C_sign, a.M = g.M, a.N = g.N, a.K = g.K, a.ldc = g.ldc;
    a.epi = g.epi, a.row_exp = f16 ? g.row_exp : nullptr, a.out16 = (g.C_lo && g.out16) ? 1 : 0;
    a.dbg = g_gemm_dbg;
    a.row_flag = a.out16 ? g.row_flag : nullptr;
    GEM_REQUIRE(!a.out16 || g.ldc % 8 == 0, "fp16 outputs need ldc % 8 == 0");
    GEM_REQUIRE((reinterpret_cast<uintptr_t>(g.bias) & 15) == 0, "bias must be 16-byte aligned");
    // N-tile width.  Up to one wave of 128-wide tiles: keep 128 (kernels of concurrent slices share the SMs, the
    // least padded tiling wins).  Beyond: the fewest waves x columns over the 148 SMs (sign words need
    // 32-column alignment).
    const int mt = (g.M + BM - 1) / BM;
    int best = 3;                                  // 128
    if ((long)mt * (g.N / 128) > kNumSMs) {
        long best_cost = -1;
        for (int i = kFirstWideBN; i < kNumBN; ++i) {
            if (g.C_sign && kBNs[i] % 32 != 0) continue;
            const long tiles = (long)mt * ((g.N + kBNs[i] - 1) / kBNs[i]);
            const long cost = ((tiles + kNumSMs - 1) / kNumSMs) * kBNs[i];
            if (best_cost < 0 || cost < best_cost || (cost == best_cost && kBNs[i] == 128)) best = i, best_cost = cost;
        }
    }
    if (const char* env = getenv("GEM_GEMM_BN")) {
        for (int i = kFirstWideBN; i < kNumBN; ++i)
            if (atoi(env) == kBNs[i] && !(g.C_sign && kBNs[i] % 32 != 0)) best = i;
    }
    // fp16 scheme: CTA pairs (cta_group::2) unless GEM_GEMM_PAIR=0; 160-wide tiles (the lowest shared-memory traffic
    // per MMA that three accumulators allow), 128 on request
    static const int pair_env = []() {
        const char* env = getenv("GEM_GEMM_PAIR");
        return env ? atoi(env) : 1;
    }();
    const int pair_mode = g_gemm_pair >= 0 ? g_gemm_pair : pair_env;
    if (f16 && pair_mode && !(a.C_lo && !a.out16)) {      // (fp32 hi/lo outputs need the one-CTA kernel's larger staging area)
        // 160-wide tiles have the lowest shared-memory traffic per MMA.  A small batch (a slice of a few hundred windows:
        // strong scaling, a real 30-clip dataset) leaves most SMs without a tile; GEM_GEMM_NARROW=1 then takes the
        // narrowest tile (64 / 96 / 128 columns) that still fits the machine in one wave.  Bit-identical, but measured
        // to gain only 3-7 % of the GEMM at 236 windows (34.5 -> 32 us per launch: the tile's K loop is bound by its
        // operand loads, not by the MMAs) and nothing at 372 / 472 windows: off by default.
        int bn = 160;
        const long pairs_m = (g.M + 2 * BM - 1) / (2 * BM);
        static const int narrow_env = []() {
            const char* env = getenv("GEM_GEMM_NARROW");
            return env ? atoi(env) : 0;
        }();
        if (narrow_env) {
            const int cand[3] = {64, 96, 128};
            for (int c : cand)
                if (pairs_m * ((g.N + c - 1) / c) <= kNumSMs / 2) {
                    bn = c;
                    break;
                }
        }
        if (const char* env = getenv("GEM_GEMM_BN")) bn = atoi(env);
        if (bn == 64) return launch_pair_bn<64>(stream, map_a_hi, map_a_lo, wit->second, 0, a);
        if (bn == 96) return launch_pair_bn<96>(stream, map_a_hi, map_a_lo, wit->second, 1, a);
        if (bn == 128) return launch_pair_bn<128>(stream, map_a_hi, map_a_lo, wit->second, 3, a);
        if (bn == 144) return launch_pair_bn<144>(stream, map_a_hi, map_a_lo, wit->second, 4, a);
        return launch_pair_bn<160>(stream, map_a_hi, map_a_lo, wit->second, 5, a);
    }
    return f16 ? launch_scheme<true>(kBNs[best], stream, map_a_hi, map_a_lo, wit->second, best, a)
               : launch_scheme<false>(kBNs[best], stream, map_a_hi, map_a_lo, wit->second, best, a);
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
// x -> fp16 (hi, scaled lo) for M rows of K floats, rows optionally pre-scaled by 2^row_exp[m]
int launch_split_f16(cudaStream_t stream, const float* A, int lda, int M, int K, const int32_t* row_exp, uint16_t* hi,
                     uint16_t* lo, uint32_t* row_flag) {
    if (M <= 0) return GEM_OK;
    GEM_REQUIRE(K % 4 == 0 && lda % 4 == 0, "K and lda must be multiples of 4");
    const size_t n4 = (size_t)M * (K / 4);
    split_f16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(A, lda, M, K, row_exp, hi, lo, row_flag);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

// rows of x = xh (+ xl) -> per-row power-of-two scale + fp16 split of the scaled row (K % 4 == 0, dense rows)
int launch_rowscale_split_f16(cudaStream_t stream, const float* xh, const float* xl, int M, int K, uint16_t* hi,
                              uint16_t* lo, int32_t* row_exp) {
    if (M <= 0) return GEM_OK;
    GEM_REQUIRE(K % 4 == 0, "K must be a multiple of 4");
    rowscale_split_f16_kernel<<<M, 256, 0, stream>>>(xh, xl, K, hi, lo, row_exp);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

// Transposes B [K][ldb] to K-major [N][K] and splits it into hi / lo parts for one scheme (1 = 3xTF32, 2 = fp16);
// replaces any earlier registration of the same pointer (the caller may have refilled the buffer).
int tc_gemm_prepare_weight(void* owner, cudaStream_t stream, const float* B, int ldb, int K, int N, int scheme) {
    const bool f16 = scheme == 2;
    GEM_REQUIRE(K % (f16 ? 64 : 32) == 0 && N % 128 == 0, "K must be a multiple of the 128-byte K block and N of 128");
    TcState* st = state_of(owner);
    const auto key = std::make_pair(B, (int)(f16 ? 2 : 1));
    auto old = st->weights.find(key);
    if (old != st->weights.end()) {
        GEM_CUDA(cudaStreamSynchronize(stream));
        cudaFree(old->second.hi), cudaFree(old->second.lo);
        st->weights.erase(old);
    }
    WeightSplit ws;
    ws.K = K, ws.N = N;
    const size_t bytes = (size_t)N * K * (f16 ? 2 : 4);
    GEM_CUDA(cudaMalloc(&ws.hi, bytes));
    GEM_CUDA(cudaMalloc(&ws.lo, bytes));
    dim3 grid((K + 31) / 32, (N + 31) / 32), block(32, 8);
    if (f16)
        transpose_split_f16_kernel<<<grid, block, 0, stream>>>(B, ldb, K, N, (uint16_t*)ws.hi, (uint16_t*)ws.lo);
    else
        transpose_split_kernel<<<grid, block, 0, stream>>>(B, ldb, K, N, (float*)ws.hi, (float*)ws.lo);
    GEM_CHECK_LAUNCH();
    for (int i = 0; i < kNumBN; ++i) {
        int rc = make_map(&ws.map_hi[i], ws.hi, f16, (uint64_t)N, (uint64_t)K, 0, kBNs[i]);
        if (rc == GEM_OK) rc = make_map(&ws.map_lo[i], ws.lo, f16, (uint64_t)N, (uint64_t)K, 0, kBNs[i]);
        if (rc == GEM_OK && f16) rc = make_map(&ws.map_hi2[i], ws.hi, f16, (uint64_t)N, (uint64_t)K, 0, kBNs[i] / 2);
        if (rc == GEM_OK && f16) rc = make_map(&ws.map_lo2[i], ws.lo, f16, (uint64_t)N, (uint64_t)K, 0, kBNs[i] / 2);
        if (rc != GEM_OK) return rc;
    }
    st->weights.emplace(key, ws);
    return GEM_OK;
}

// frees the prepared copies of one weight matrix (every scheme); the caller has synchronised the device
void tc_gemm_forget_weight(void* owner, const float* B) {
    TcState* st = state_of(owner);
    for (int scheme = 1; scheme <= 2; ++scheme) {
        auto it = st->weights.find(std::make_pair(B, scheme));
        if (it == st->weights.end()) continue;
        cudaFree(it->second.hi), cudaFree(it->second.lo);
        st->weights.erase(it);
    }
}

void tc_gemm_release(void* owner) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_states.find(owner);
    if (it == g_states.end()) return;
    TcState* st = it->second;
    for (auto& kv : st->weights) cudaFree(kv.second.hi), cudaFree(kv.second.lo);
    for (auto& kv : st->scratch)
        if (kv.second.hi) cudaFree(kv.second.hi), cudaFree(kv.second.lo);
    delete st;
    g_states.erase(it);
}

}  // namespace gem
