// Generic "tap GEMM" in fp32 on the CUDA cores: every VAE layer in token-major form.
//
//   C[m][n] = epi( bias[n] + sum_{tap<taps} sum_{k<K} A[m + tap - taps/2][k] * B[tap][k][n] )
//
// Rows m are (window, frame) tokens, m = w*T + t; a tap that would leave the window's T frames
// reads zeros.  With taps = 3 this is Conv1d / ConvTranspose1d(k=3, s=1, p=1) or the bwd-data of
// either (the host arranges B accordingly, globalegomocap_b200/vae_prep.py); with taps = 1 it is a
// plain GEMM (the fused decoder_input+ConvT layer, its transpose, and the encoder's fc_mu|fc_var).
// Epilogues: bias, LeakyReLU(0.01) (SeqConvVAE.py:36-38,76-78), or multiplication by the LeakyReLU
// derivative of a saved activation (bwd-data; the mask is recomputed from the stored output's sign).
//
// This is the always-available fp32 path and the parity baseline for the tcgen05 kernel
// (gemm_tc.cu) that takes over the large contractions.
#include "kernels.cuh"

namespace gem {

template <int BM, int BN, int BK, int TM, int TN, bool VEC_A>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) tap_gemm_kernel(TapGemmArgs g) {
    constexpr int NT = (BM / TM) * (BN / TN);
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);

    const int kchunks = (g.K + BK - 1) / BK;
    const int total = kchunks * g.taps;
    const int half = g.taps / 2;

    // A tile loader: BM x BK, each thread owns (BM*BK/4)/NT float4 (or 4 scalars)
    constexpr int A_V4 = BM * BK / 4;
    constexpr int A_PER = (A_V4 + NT - 1) / NT;
    constexpr int B_V4 = BK * BN / 4;
    constexpr int B_PER = (B_V4 + NT - 1) / NT;
    float4 ra[A_PER], rb[B_PER];

    auto load_tiles = [&](int it) {
        const int tap = it / kchunks, k0 = (it - tap * kchunks) * BK;
        const int shift = tap - half;
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int v = tid + i * NT;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < A_V4) {
                const int r = v / (BK / 4), kk = (v % (BK / 4)) * 4;
                const int m = m0 + r;
                bool ok = m < g.M;
                if (ok && g.taps > 1) {
                    const int t = m % g.T + shift;
                    ok = t >= 0 && t < g.T;
                }
                if (ok) {
                    const float* src = g.A + (size_t)(m + shift) * g.lda + k0 + kk;
                    if (VEC_A && k0 + kk + 3 < g.K) {
                        val = *reinterpret_cast<const float4*>(src);
                    } else {
                        if (k0 + kk + 0 < g.K) val.x = src[0];
                        if (k0 + kk + 1 < g.K) val.y = src[1];
                        if (k0 + kk + 2 < g.K) val.z = src[2];
                        if (k0 + kk + 3 < g.K) val.w = src[3];
                    }
                }
            }
            ra[i] = val;
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int v = tid + i * NT;
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (v < B_V4) {
                const int kk = v / (BN / 4), nn = (v % (BN / 4)) * 4;
                if (k0 + kk < g.K && n0 + nn < g.ldb)
                    val = __ldg(reinterpret_cast<const float4*>(g.B + ((size_t)tap * g.K + k0 + kk) * g.ldb + n0 + nn));
            }
            rb[i] = val;
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_PER; ++i) {
            const int v = tid + i * NT;
            if (v < A_V4) {
                const int r = v / (BK / 4), kk = (v % (BK / 4)) * 4;
                As[buf][kk + 0][r] = ra[i].x;
                As[buf][kk + 1][r] = ra[i].y;
                As[buf][kk + 2][r] = ra[i].z;
                As[buf][kk + 3][r] = ra[i].w;
            }
        }
#pragma unroll
        for (int i = 0; i < B_PER; ++i) {
            const int v = tid + i * NT;
            if (v < B_V4) {
                const int kk = v / (BN / 4), nn = (v % (BN / 4)) * 4;
                *reinterpret_cast<float4*>(&Bs[buf][kk][nn]) = rb[i];
            }
        }
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    for (int it = 0; it < total; ++it) {
        const int buf = it & 1;
        if (it + 1 < total) load_tiles(it + 1);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float av[TM], bv[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
                av[i] = t4.x, av[i + 1] = t4.y, av[i + 2] = t4.z, av[i + 3] = t4.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                const float4 t4 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * TN + j]);
                bv[j] = t4.x, bv[j + 1] = t4.y, bv[j + 2] = t4.z, bv[j + 3] = t4.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (it + 1 < total) {
            store_tiles(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue ---------------------------------------------------------------------------
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            if (g.bias) v += __ldg(g.bias + n);
            if (g.epi == EPI_LRELU) {
                v = v > 0.f ? v : v * 0.01f;
            } else if (g.epi == EPI_MASK) {
                bool pos;
                if (g.aux_bits) pos = (g.aux_bits[(size_t)m * (g.N >> 5) + (n >> 5)] >> (n & 31)) & 1u;
                else pos = g.aux[(size_t)m * g.ldaux + n] > 0.f;
                v = pos ? v : v * 0.01f;
            }
            g.C[(size_t)m * g.ldc + n] = v;
        }
    }
}

int launch_tap_gemm_simt(cudaStream_t stream, const TapGemmArgs& g) {
    if (g.M <= 0 || g.N <= 0) return GEM_OK;
    GEM_REQUIRE(g.taps == 1 || g.taps == 3, "taps must be 1 or 3");
    GEM_REQUIRE(g.ldb % 4 == 0 && (reinterpret_cast<uintptr_t>(g.B) & 15u) == 0, "B rows must be 16-byte aligned");
    const bool vec_a = (g.lda % 4 == 0) && ((reinterpret_cast<uintptr_t>(g.A) & 15u) == 0);
    if (g.N >= 512 && g.M >= 256) {
        dim3 grid((g.M + 127) / 128, (g.N + 127) / 128);
        if (vec_a)
            tap_gemm_kernel<128, 128, 16, 8, 8, true><<<grid, 256, 0, stream>>>(g);
        else
            tap_gemm_kernel<128, 128, 16, 8, 8, false><<<grid, 256, 0, stream>>>(g);
    } else {
        dim3 grid((g.M + 63) / 64, (g.N + 63) / 64);
        if (vec_a)
            tap_gemm_kernel<64, 64, 16, 4, 4, true><<<grid, 256, 0, stream>>>(g);
        else
            tap_gemm_kernel<64, 64, 16, 4, 4, false><<<grid, 256, 0, stream>>>(g);
    }
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

}  // namespace gem
