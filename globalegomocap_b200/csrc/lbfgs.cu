// Batched L-BFGS with strong-Wolfe line search: W independent solves advanced in lock-step,
// one closure evaluation per round, no host synchronisation.
//
// Semantics are those of torch.optim.LBFGS.step (torch 2.11 optim/lbfgs.py:332-537, _strong_wolfe
// :40-209, _cubic_interpolate :12-37) as constructed by the reference at optimizer.py:261-270, for a
// fresh optimiser per window.  The Python control flow becomes an explicit per-window state machine:
// `advance` consumes (loss, grad) at the requested trial point, runs the optimiser's logic up to
// the next point it needs evaluated and writes that point; finished windows are parked.
//
// One CTA per window; the n-vector work (dots, axpys, two-loop recursion over the (y, s) history in
// HBM, streamed through a shared-memory ring by the TMA engine) is spread over the CTA, the scalar
// decision logic runs on thread 0.  torch mixes Python floats (double) with fp32 0-d tensors; `Num`
// reproduces that typing so that step lengths round the way the reference's do.  Compiled with
// --fmad=false.
#include <stdlib.h>

#include "kernels.cuh"

namespace gem {

// ---- torch-like scalar: Python float (double) or fp32 0-d tensor -------------------------
struct Num {
    double v;
    int is32;
};
__device__ __forceinline__ Num py(double v) { return Num{v, 0}; }
__device__ __forceinline__ Num t32(float v) { return Num{(double)v, 1}; }
__device__ __forceinline__ Num nadd(Num a, Num b) {
    if (a.is32 | b.is32) return t32(__fadd_rn((float)a.v, (float)b.v));
    return py(__dadd_rn(a.v, b.v));
}
__device__ __forceinline__ Num nsub(Num a, Num b) {
    if (a.is32 | b.is32) return t32(__fsub_rn((float)a.v, (float)b.v));
    return py(__dsub_rn(a.v, b.v));
}
__device__ __forceinline__ Num nmul(Num a, Num b) {
    if (a.is32 | b.is32) return t32(__fmul_rn((float)a.v, (float)b.v));
    return py(__dmul_rn(a.v, b.v));
}
__device__ __forceinline__ Num ndiv(Num a, Num b) {
    if (!a.is32 && b.is32)      // Tensor.__rtruediv__: reciprocal() * scalar
        return t32(__fmul_rn(__frcp_rn((float)b.v), (float)a.v));
    if (a.is32 | b.is32) return t32(__fdiv_rn((float)a.v, (float)b.v));
    return py(__ddiv_rn(a.v, b.v));
}
__device__ __forceinline__ double ncmp_val(Num a, Num b, double& bv) {   // promote like torch/NEP 50
    if (a.is32 | b.is32) {
        bv = (double)(float)b.v;
        return (double)(float)a.v;
    }
    bv = b.v;
    return a.v;
}
__device__ __forceinline__ bool nlt(Num a, Num b) { double bv; double av = ncmp_val(a, b, bv); return av < bv; }
__device__ __forceinline__ bool nle(Num a, Num b) { double bv; double av = ncmp_val(a, b, bv); return av <= bv; }
__device__ __forceinline__ bool ngt(Num a, Num b) { return nlt(b, a); }
__device__ __forceinline__ bool nge(Num a, Num b) { return nle(b, a); }
__device__ __forceinline__ Num nabs(Num a) { return Num{fabs(a.v), a.is32}; }
__device__ __forceinline__ Num nmaxp(Num a, Num b) { return ngt(b, a) ? b : a; }   // Python max(a, b)
__device__ __forceinline__ Num nminp(Num a, Num b) { return nlt(b, a) ? b : a; }   // Python min(a, b)

enum { PH_INIT = 0, PH_BRACKET = 1, PH_ZOOM = 2, PH_DONE = 3 };

struct LbfgsWin {
    int phase, n_iter, evals, hist_len;
    int ls_iter, max_ls, ls_evals, ls_first;
    int ls_done, insuf, low_pos, high_pos, nbracket, gp_is_g;   // gp_is_g: the bracket's g_prev is still G
    double loss, prev_loss;
    float h_diag, gtd, d_norm, pad1;
    Num t, t_prev, br[2];
    double f_prev, br_f[2];
    float gtd_prev, br_gtd[2], pad2;
};

constexpr int kLbThreads = 256;
constexpr int kWarps = kLbThreads / 32;
constexpr int kChunks = 2;          // float4 chunks per thread: n <= 2048 keeps a vector in 8 registers
constexpr int kPer = 4 * kChunks;
#ifndef GEM_LBFGS_OCC
#define GEM_LBFGS_OCC 4     // resident CTAs per SM the register budget is compiled for
#endif
#ifndef GEM_LBFGS_RING
#define GEM_LBFGS_RING 4
#endif
#ifndef GEM_LBFGS_SHARED_ADDR
#define GEM_LBFGS_SHARED_ADDR 1   // 1: the recursion addresses its ring and barriers by 32-bit shared addresses held in registers
#endif
#ifndef GEM_LBFGS_L2HINT
#define GEM_LBFGS_L2HINT 0  // 1: history rows of the recursion's first loop are loaded evict_last (the second loop re-reads
#endif                      //    them in reverse order), those of the second loop evict_first
constexpr int kRing = GEM_LBFGS_RING;   // rows (n floats each) of the (y, s) history in flight per CTA
constexpr int kMaxHist = 64;

// ---- per-thread slices: chunk c = tid + i*kLbThreads covers elements [4c, 4c+4) ---------------
__device__ __forceinline__ void ld8(float (&r)[kPer], const float* __restrict__ p, int tid, int n4) {
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
        const int c = tid + i * kLbThreads;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < n4) v = *reinterpret_cast<const float4*>(p + 4 * c);
        r[4 * i + 0] = v.x, r[4 * i + 1] = v.y, r[4 * i + 2] = v.z, r[4 * i + 3] = v.w;
    }
}
// the same from a 32-bit shared-window address (`addr` = this thread's first float4 of the row).  The history ring and
// its barriers are addressed this way inside the recursion: through generic pointers every access re-derived the
// shared window's base from two special registers (S2R), ~25 of the ~80 instructions a row costs a warp.
__device__ __forceinline__ void lds8(float (&r)[kPer], uint32_t addr, int tid, int n4) {
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
        const int c = tid + i * kLbThreads;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < n4)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"(addr + (uint32_t)(i * kLbThreads * 16))
                         : "memory");
        r[4 * i + 0] = v.x, r[4 * i + 1] = v.y, r[4 * i + 2] = v.z, r[4 * i + 3] = v.w;
    }
}
__device__ __forceinline__ void mbar_wait_s(uint32_t bar_addr, uint32_t parity) {      // mbar_wait (common.cuh) on a shared address
    unsigned long long t0 = 0;
    for (uint32_t spin = 0;; ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar_addr), "r"(parity)
            : "memory");
        if (ok) break;
        if ((spin & 255u) == 255u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 4000000000ull) __trap();
        }
    }
}
__device__ __forceinline__ void st8(float* __restrict__ p, const float (&r)[kPer], int tid, int n4) {
#pragma unroll
    for (int i = 0; i < kChunks; ++i) {
        const int c = tid + i * kLbThreads;
        if (c < n4) *reinterpret_cast<float4*>(p + 4 * c) = make_float4(r[4 * i], r[4 * i + 1], r[4 * i + 2], r[4 * i + 3]);
    }
}
// trial point: plain (host-driven closures read it) and, when requested, split into TF32 hi / lo parts so the
// decoder's first tensor-core GEMM can consume it without a separate split pass
__device__ __forceinline__ void split_f16_dev(float x, uint16_t& hi, uint16_t& lo) {     // = tc::split_f16
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(hi) : "f"(x));
    float hf;
    asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(hi));
    const float r = (x - hf) * 2048.f;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(lo) : "f"(r));
}
__device__ __forceinline__ void st_trial(const LbfgsBuffers& b, size_t off, const float (&r)[kPer], int tid, int n4) {
    st8(b.ZT + off, r, tid, n4);
    if (!b.ZT_hi) return;
    if (b.zt_f16) {
        uint16_t* zh = reinterpret_cast<uint16_t*>(b.ZT_hi) + off;
        uint16_t* zl = reinterpret_cast<uint16_t*>(b.ZT_lo) + off;
        if (b.status) {                     // fp16 range check (the split saturates silently)
            float amax = 0.f;
#pragma unroll
            for (int i = 0; i < kPer; ++i) amax = fmaxf(amax, fabsf(r[i]));
            if (amax > 65504.f) atomicOr(b.status + off / (size_t)b.n, GEM_WIN_F16_RANGE);
        }
#pragma unroll
        for (int i = 0; i < kChunks; ++i) {
            const int c = tid + i * kLbThreads;
            if (c >= n4) continue;
            uint16_t h[4], l[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) split_f16_dev(r[4 * i + j], h[j], l[j]);
            *reinterpret_cast<uint2*>(zh + 4 * c) =
                make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
            *reinterpret_cast<uint2*>(zl + 4 * c) =
                make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
        }
    } else {
        float h[kPer], l[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            h[i] = __uint_as_float(__float_as_uint(r[i]) & 0xFFFFE000u);
            l[i] = r[i] - h[i];
        }
        st8(b.ZT_hi + off, h, tid, n4);
        st8(b.ZT_lo + off, l, tid, n4);
    }
}
__device__ __forceinline__ float dot8(const float (&a)[kPer], const float (&b)[kPer]) {
    float p = 0.f;
#pragma unroll
    for (int i = 0; i < kPer; ++i) p += a[i] * b[i];
    return p;
}

// Up to three block-wide reductions with ONE barrier (partials are double-buffered by `par`); bit j of
// MAXMASK makes value j a max instead of a sum.  Fixed tree: deterministic run to run.
struct RedBuf {
    float v[2][3][kWarps];
};
// barrier over the kLbThreads threads of one window group (group 0 of a one-group CTA: the same as __syncthreads)
__device__ __forceinline__ void group_sync(int gid) {
    asm volatile("bar.sync %0, %1;" ::"r"(gid + 1), "n"(kLbThreads) : "memory");
}
template <int NV, int MAXMASK>
__device__ __forceinline__ void block_reduce(float (&x)[NV], RedBuf& red, int& par, int tid, int gid) {
    const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        x[j] = ((MAXMASK >> j) & 1) ? warp_max(x[j]) : warp_sum(x[j]);
        if (lane == 0) red.v[par][j][warp] = x[j];
    }
    group_sync(gid);
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        float r = red.v[par][j][0];
#pragma unroll
        for (int i = 1; i < kWarps; ++i) r = ((MAXMASK >> j) & 1) ? fmaxf(r, red.v[par][j][i]) : r + red.v[par][j][i];
        x[j] = r;
    }
    par ^= 1;
}

// _cubic_interpolate (lbfgs.py:12-37); f's are Python floats, g's fp32 tensors
__device__ Num cubic_interpolate(Num x1, double f1, float g1, Num x2, double f2, float g2, bool has_bounds,
                                 Num lo, Num hi) {
    Num xmin, xmax;
    if (has_bounds) {
        xmin = lo, xmax = hi;
    } else if (nle(x1, x2)) {
        xmin = x1, xmax = x2;
    } else {
        xmin = x2, xmax = x1;
    }
    // d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)
    const Num three_df = py(__dmul_rn(3.0, __dsub_rn(f1, f2)));
    const Num d1 = nsub(t32(__fadd_rn(g1, g2)), ndiv(three_df, nsub(x1, x2)));
    const float d1f = (float)d1.v;
    const float d2_square = __fsub_rn(__fmul_rn(d1f, d1f), __fmul_rn(g1, g2));
    if (d2_square >= 0.f) {
        const float d2 = __fsqrt_rn(d2_square);
        Num min_pos;
        if (nle(x1, x2)) {
            const float frac = __fdiv_rn(__fsub_rn(__fadd_rn(g2, d2), d1f),
                                         __fadd_rn(__fsub_rn(g2, g1), __fmul_rn(2.f, d2)));
            min_pos = nsub(x2, nmul(nsub(x2, x1), t32(frac)));
        } else {
            const float frac = __fdiv_rn(__fsub_rn(__fadd_rn(g1, d2), d1f),
                                         __fadd_rn(__fsub_rn(g1, g2), __fmul_rn(2.f, d2)));
            min_pos = nsub(x1, nmul(nsub(x1, x2), t32(frac)));
        }
        return nminp(nmaxp(min_pos, xmin), xmax);
    }
    return ndiv(nadd(xmin, xmax), py(2.0));
}

__global__ void __launch_bounds__(kLbThreads) lbfgs_begin_kernel(LbfgsBuffers b, const float* __restrict__ z0, int W) {
    const int w = blockIdx.x;
    const size_t off = (size_t)w * b.n;
    for (int i = threadIdx.x; i < b.n; i += kLbThreads) {
        const float v = z0[off + i];
        b.X[off + i] = v;
        b.ZT[off + i] = v;
        if (b.ZT_hi && b.zt_f16) {
            uint16_t h, l;
            if (b.status && fabsf(v) > 65504.f) atomicOr(b.status + w, GEM_WIN_F16_RANGE);
            split_f16_dev(v, h, l);
            reinterpret_cast<uint16_t*>(b.ZT_hi)[off + i] = h;
            reinterpret_cast<uint16_t*>(b.ZT_lo)[off + i] = l;
        } else if (b.ZT_hi) {
            const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
            b.ZT_hi[off + i] = h;
            b.ZT_lo[off + i] = v - h;
        }
    }
    if (threadIdx.x == 0) {
        LbfgsWin s;
        memset(&s, 0, sizeof(s));
        s.phase = PH_INIT;
        s.h_diag = 1.f;
        b.st[w] = s;
    }
}

// One CTA per window.  Vector work is spread over the CTA (8 elements per thread, float4 accesses),
// the scalar decisions run on thread 0.  The two-loop recursion streams the window's (y, s) history
// from HBM through a kRing-row shared-memory ring filled by the TMA engine (cp.async.bulk + mbarrier),
// so the dependent chain dot -> reduce -> axpy never waits on a cold global load.
//
// torch keeps g, prev_flat_grad and the bracket's g_prev as separate clones; here
//   * prev_flat_grad is G itself: y = g_new - G is formed before G is overwritten;
//   * g_prev aliases G until the bracket phase interpolates (LbfgsWin::gp_is_g);
//   * the bracket gradients are only stored while the line search is still running.
// everything the kLbThreads threads that advance one window share
struct GroupSmem {
    RedBuf red;
    LbfgsWin s;                                        // thread 0 mutates; others only read the snapshot below
    int sh_action, sh_code, sh_copy_gp, sh_gsrc, sh_gp_is_g, sh_stop, sh_done2;
    float sh_tf, sh_tf2;
    float al_s[kMaxHist], ro_s[kMaxHist];
    __align__(8) uint64_t full_bar[kRing];
};

// advances window w by one closure evaluation; called by one group of kLbThreads threads (tid = 0 .. kLbThreads-1,
// barrier id gid + 1), `ring` = the group's [kRing][n] floats of shared memory
template <bool kReuse>
__device__ __forceinline__ void lbfgs_advance_window(const LbfgsBuffers& b, const float* __restrict__ loss_in,
                                                     const float* __restrict__ grad_in, int w, int tid, int gid,
                                                     GroupSmem& sm, float* ring, bool first_window, uint32_t& rows_done) {
    RedBuf& red = sm.red;
    LbfgsWin& s = sm.s;
    int &sh_action = sm.sh_action, &sh_code = sm.sh_code, &sh_copy_gp = sm.sh_copy_gp, &sh_gsrc = sm.sh_gsrc;
    int &sh_gp_is_g = sm.sh_gp_is_g, &sh_stop = sm.sh_stop, &sh_done2 = sm.sh_done2;
    float &sh_tf = sm.sh_tf, &sh_tf2 = sm.sh_tf2;
    float* al_s = sm.al_s;
    float* ro_s = sm.ro_s;
    uint64_t* full_bar = sm.full_bar;

    const int n = b.n, n4 = b.n >> 2;
    const size_t off = (size_t)w * n;
    float *X = b.X + off, *D = b.D + off, *G = b.G + off, *GP = b.GP + off, *BG0 = b.BG0 + off, *BG1 = b.BG1 + off,
          *ZT = b.ZT + off;
    (void)ZT;
    int par = 0;

    if (tid == 0) {
        s = b.st[w];
        // a persistent group keeps its barriers across windows (rows_done counts the rows they have carried)
        if (!kReuse || first_window) {
#pragma unroll
            for (int i = 0; i < kRing; ++i) mbar_init(&full_bar[i], 1);
            fence_barrier_init();
        }
    }
    group_sync(gid);
    const int phase0 = s.phase, hist0 = s.hist_len, n_iter0 = s.n_iter;
    const float h_diag0 = s.h_diag;
    if (phase0 == PH_DONE) return;

    const double f_new = (double)loss_in[w];
    float gn[kPer];
    ld8(gn, grad_in + off, tid, n4);
    if (tid == 0 && b.trace) {
        const int idx = s.evals + ((phase0 == PH_INIT) ? 0 : s.ls_evals);
        if (idx < b.trace_stride) b.trace[(size_t)w * b.trace_stride + idx] = (float)f_new;
    }

    enum { ACT_EVAL = 2, ACT_LS_FINISH = 3 };
    float yv[kPer], sv[kPer];          // newest curvature pair (valid when !first_iter)
    bool first_iter;

    if (phase0 == PH_INIT) {
        // lbfgs.py:361-373: first evaluation
        st8(G, gn, tid, n4);
        float r1[1] = {0.f};
#pragma unroll
        for (int i = 0; i < kPer; ++i) r1[0] = fmaxf(r1[0], fabsf(gn[i]));
        block_reduce<1, 1>(r1, red, par, tid, gid);
        if (tid == 0) {
            s.loss = f_new;
            s.evals = 1;
            s.n_iter = 0;
            sh_stop = 0;
            if ((double)r1[0] <= (double)(float)b.tol_grad) {
                s.phase = PH_DONE;
                sh_stop = 1;
                b.st[w] = s;
            }
        }
        group_sync(gid);
        if (sh_stop) return;
        first_iter = true;
    } else {
        // ---- an evaluation requested by the line search has arrived ----------------------------
        float d[kPer];
        ld8(d, D, tid, n4);
        float r1[1] = {dot8(gn, d)};
        block_reduce<1, 0>(r1, red, par, tid, gid);
        const float gtd_new = r1[0];
        if (tid == 0) {
            const double c1 = 1e-4, c2 = 0.9;
            const Num gtd = t32(s.gtd);
            int act = 0, code = 0, copy_gp = 0;
            // what each bracket-gradient slot holds after this call: 0 g_new, 1 BG0, 2 BG1 (as stored), 3 GP, 4 G
            int slot_src[2] = {1, 2};
            s.ls_evals += 1;
            if (s.phase == PH_BRACKET) {
                bool to_zoom_entry = false;
                int set_br = 0;   // 1: bracket=[t_prev,t] g=[g_prev,g_new]; 2: bracket=[t] g=[g_new]; 3: [0,t] g=[g,g_new]
                if (!s.ls_first) s.ls_iter += 1;
                s.ls_first = 0;
                if (s.ls_iter < s.max_ls) {
                    // f_new > (f + c1 * t * gtd) or (ls_iter > 1 and f_new >= f_prev)
                    const Num rhs = nadd(py(s.loss), nmul(nmul(py(c1), s.t), gtd));
                    if (ngt(py(f_new), rhs) || (s.ls_iter > 1 && f_new >= s.f_prev)) {
                        set_br = 1;
                        to_zoom_entry = true;
                    } else if (nle(t32(fabsf(gtd_new)), nmul(py(-c2), gtd))) {
                        set_br = 2;
                        s.ls_done = 1;
                        to_zoom_entry = true;
                    } else if (gtd_new >= 0.f) {
                        set_br = 1;
                        to_zoom_entry = true;
                    } else {
                        // interpolate (lbfgs.py:78-96)
                        const Num min_step = nadd(s.t, nmul(py(0.01), nsub(s.t, s.t_prev)));
                        const Num max_step = nmul(s.t, py(10.0));
                        const Num tmp = s.t;
                        s.t = cubic_interpolate(s.t_prev, s.f_prev, s.gtd_prev, s.t, f_new, gtd_new, true, min_step,
                                                max_step);
                        s.t_prev = tmp;
                        s.f_prev = f_new;
                        s.gtd_prev = gtd_new;
                        copy_gp = 1;                 // g_prev = g_new.clone()
                        act = ACT_EVAL;
                    }
                } else {
                    set_br = 3;          // reached max number of line-search iterations
                    to_zoom_entry = true;
                }
                if (to_zoom_entry) {
                    if (set_br == 1) {
                        s.br[0] = s.t_prev, s.br[1] = s.t;
                        s.br_f[0] = s.f_prev, s.br_f[1] = f_new;
                        s.br_gtd[0] = s.gtd_prev, s.br_gtd[1] = gtd_new;
                        s.nbracket = 2;
                        slot_src[0] = s.gp_is_g ? 4 : 3, slot_src[1] = 0;
                    } else if (set_br == 2) {
                        s.br[0] = s.t;
                        s.br_f[0] = f_new;
                        s.nbracket = 1;
                        slot_src[0] = 0;
                    } else {
                        s.br[0] = py(0.0), s.br[1] = s.t;
                        s.br_f[0] = s.loss, s.br_f[1] = f_new;
                        s.nbracket = 2;
                        slot_src[0] = 4, slot_src[1] = 0;
                    }
                    s.insuf = 0;
                    const double flast = s.br_f[s.nbracket - 1];
                    if (s.br_f[0] <= flast) s.low_pos = 0, s.high_pos = 1;
                    else s.low_pos = 1, s.high_pos = 0;
                    s.phase = PH_ZOOM;
                }
                code = set_br;
            } else {   // PH_ZOOM: an evaluation inside the zoom loop (lbfgs.py:152-199)
                s.ls_iter += 1;
                const Num rhs = nadd(py(s.loss), nmul(nmul(py(c1), s.t), gtd));
                if (ngt(py(f_new), rhs) || f_new >= s.br_f[s.low_pos]) {
                    s.br[s.high_pos] = s.t;
                    s.br_f[s.high_pos] = f_new;
                    s.br_gtd[s.high_pos] = gtd_new;
                    code = 10 + s.high_pos;                   // BG[high] = g_new
                    slot_src[s.high_pos] = 0;
                    if (s.br_f[0] <= s.br_f[1]) s.low_pos = 0, s.high_pos = 1;
                    else s.low_pos = 1, s.high_pos = 0;
                } else {
                    code = 20 + s.low_pos;                    // BG[low] = g_new
                    if (nle(t32(fabsf(gtd_new)), nmul(py(-c2), gtd))) {
                        s.ls_done = 1;
                    } else if (nge(nmul(t32(gtd_new), nsub(s.br[s.high_pos], s.br[s.low_pos])), py(0.0))) {
                        s.br[s.high_pos] = s.br[s.low_pos];
                        s.br_f[s.high_pos] = s.br_f[s.low_pos];
                        s.br_gtd[s.high_pos] = s.br_gtd[s.low_pos];
                        code = 30 + s.low_pos;                // BG[high] = BG[low]; BG[low] = g_new
                        slot_src[s.high_pos] = slot_src[s.low_pos];
                    }
                    slot_src[s.low_pos] = 0;
                    s.br[s.low_pos] = s.t;
                    s.br_f[s.low_pos] = f_new;
                    s.br_gtd[s.low_pos] = gtd_new;
                }
            }
            // ---- zoom loop head (lbfgs.py:109-150): either ask for another evaluation or finish
            if (s.phase == PH_ZOOM && act == 0) {
                bool fin = true;
                if (!s.ls_done && s.ls_iter < s.max_ls) {
                    // abs(bracket[1] - bracket[0]) * d_norm < tolerance_change (line search default 1e-9)
                    const Num width = nmul(nabs(nsub(s.br[1], s.br[0])), t32(s.d_norm));
                    if (!nlt(width, py(1e-9))) {
                        Num t = cubic_interpolate(s.br[0], s.br_f[0], s.br_gtd[0], s.br[1], s.br_f[1], s.br_gtd[1], false,
                                                  py(0), py(0));
                        const Num bmax = nmaxp(s.br[0], s.br[1]), bmin = nminp(s.br[0], s.br[1]);
                        const Num eps = nmul(py(0.1), nsub(bmax, bmin));
                        if (nlt(nminp(nsub(bmax, t), nsub(t, bmin)), eps)) {
                            if (s.insuf || nge(t, bmax) || nle(t, bmin)) {
                                if (nlt(nabs(nsub(t, bmax)), nabs(nsub(t, bmin)))) t = nsub(bmax, eps);
                                else t = nadd(bmin, eps);
                                s.insuf = 0;
                            } else {
                                s.insuf = 1;
                            }
                        } else {
                            s.insuf = 0;
                        }
                        s.t = t;
                        fin = false;
                    }
                }
                act = fin ? ACT_LS_FINISH : ACT_EVAL;
            }
            sh_action = act, sh_code = code, sh_copy_gp = copy_gp, sh_gp_is_g = s.gp_is_g;
            sh_gsrc = slot_src[s.low_pos];
            sh_tf = (act == ACT_EVAL) ? (float)s.t.v : (float)s.br[s.low_pos].v;
            if (copy_gp) s.gp_is_g = 0;
            if (act == ACT_EVAL) b.st[w] = s;     // (a finishing call stores the state after its stop test)
        }
        group_sync(gid);
        const int code = sh_code;
        const float tf = sh_tf;
        if (sh_action == ACT_EVAL) {
            // ---- vector side effects of the decision: the bracket's gradient clones ----------------
            if (sh_copy_gp) st8(GP, gn, tid, n4);
            if (code == 1) {                 // bracket_g = [g_prev, g_new.clone()]
                float t8[kPer];
                ld8(t8, sh_gp_is_g ? G : GP, tid, n4);
                st8(BG0, t8, tid, n4), st8(BG1, gn, tid, n4);
            } else if (code == 2) {          // bracket_g = [g_new]
                st8(BG0, gn, tid, n4);
            } else if (code == 3) {          // bracket_g = [g, g_new]
                float t8[kPer];
                ld8(t8, G, tid, n4);
                st8(BG0, t8, tid, n4), st8(BG1, gn, tid, n4);
            } else if (code >= 10 && code < 30) {
                st8((code % 10) ? BG1 : BG0, gn, tid, n4);
            } else if (code >= 30) {
                float* lo = (code - 30) ? BG1 : BG0;
                float* hi = (code - 30) ? BG0 : BG1;
                float t8[kPer];
                ld8(t8, lo, tid, n4);
                st8(hi, t8, tid, n4), st8(lo, gn, tid, n4);
            }
            float x[kPer];
            ld8(x, X, tid, n4);
#pragma unroll
            for (int i = 0; i < kPer; ++i) x[i] = x[i] + tf * d[i];      // _add_grad(t, d) from x_init
            st_trial(b, off, x, tid, n4);
            return;
        }
        // ---- line search finished (lbfgs.py:201-209, 488-527): accept the low bracket point -----
        float g[kPer], x[kPer];
        {
            const int src = sh_gsrc;
            if (src == 0) {
#pragma unroll
                for (int i = 0; i < kPer; ++i) g[i] = gn[i];
            } else {
                ld8(g, src == 1 ? BG0 : src == 2 ? BG1 : src == 3 ? GP : G, tid, n4);
            }
        }
        ld8(yv, G, tid, n4);           // prev_flat_grad
        ld8(x, X, tid, n4);
        float r2[2] = {0.f, 0.f};
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            yv[i] = g[i] - yv[i];                                   // y = flat_grad - prev_flat_grad
            sv[i] = tf * d[i];                                      // s = d * t
            x[i] = x[i] + sv[i];
            r2[0] = fmaxf(r2[0], fabsf(g[i]));
            r2[1] = fmaxf(r2[1], fabsf(sv[i]));
        }
        st8(G, g, tid, n4);
        st8(X, x, tid, n4);
        block_reduce<2, 3>(r2, red, par, tid, gid);
        if (tid == 0) {
            s.t = s.br[s.low_pos];
            s.loss = s.br_f[s.low_pos];
            s.evals += s.ls_evals;
            bool stop = false;
            if (s.n_iter == b.max_iter) stop = true;
            else if (s.evals >= b.max_eval) stop = true;
            else if ((double)r2[0] <= (double)(float)b.tol_grad) stop = true;
            else if ((double)r2[1] <= (double)(float)b.tol_change) stop = true;
            else if (fabs(s.loss - s.prev_loss) < b.tol_change) stop = true;
            sh_stop = stop;
            if (stop) {
                s.phase = PH_DONE;
                b.st[w] = s;
            }
        }
        group_sync(gid);
        if (sh_stop) {
            st_trial(b, off, x, tid, n4);
            return;
        }
        first_iter = false;
    }

    // ---- next outer iteration: direction (lbfgs.py:388-442), step length, first trial point --------
    float q[kPer];
    int k = hist0;
    float hd = h_diag0;
    if (first_iter) {
#pragma unroll
        for (int i = 0; i < kPer; ++i) q[i] = -gn[i];
    } else {
        float* Yw = b.Y + (size_t)w * b.m * n;
        float* Sw = b.S + (size_t)w * b.m * n;
        float* ROw = b.RO + (size_t)w * b.m;
        const uint32_t row_bytes = (uint32_t)n * 4u;
        // the rows the recursion will stream do not depend on whether this iteration's pair is kept:
        // start filling the ring before the y.s / y.y reductions
        const int kr = k;                       // stored pairs to stream: h = kr-1 .. 0, then 0 .. kr-1
        const int total = 4 * kr;
        int issued = 0;                         // thread 0
        auto row_src = [&](int r) -> const float* {
            if (r < 2 * kr) return ((r & 1) ? Yw : Sw) + (size_t)(kr - 1 - (r >> 1)) * n;
            const int rr = r - 2 * kr;
            return ((rr & 1) ? Sw : Yw) + (size_t)(rr >> 1) * n;
        };
        auto issue_upto = [&](int limit) {
            limit = limit < total ? limit : total;
            while (issued < limit) {
                const int slot = (int)((rows_done + (uint32_t)issued) % kRing);
                mbar_arrive_expect_tx(&full_bar[slot], row_bytes);
#if GEM_LBFGS_L2HINT
                bulk_g2s_hint(ring + (size_t)slot * n, row_src(issued), row_bytes, &full_bar[slot],
                              issued < 2 * kr ? l2_policy_evict_last() : l2_policy_evict_first());
#else
                bulk_g2s(ring + (size_t)slot * n, row_src(issued), row_bytes, &full_bar[slot]);
#endif
                ++issued;
            }
        };
        if (tid == 0) issue_upto(kRing);
        if (tid < k) ro_s[tid] = ROw[tid];
        float r2[2] = {dot8(yv, sv), dot8(yv, yv)};
        block_reduce<2, 0>(r2, red, par, tid, gid);
        const float ys = r2[0], yy = r2[1];
        // ys > 1e-10: keep the pair.  max_iter - 1 <= m (checked by the host), so the history never overflows.
        const bool pushed = (double)ys > (double)1e-10f;
        float ro_new = 0.f;
        if (pushed) {
            st8(Yw + (size_t)k * n, yv, tid, n4);
            st8(Sw + (size_t)k * n, sv, tid, n4);
            ro_new = __frcp_rn(ys);                   // 1. / ys  (reciprocal)
            hd = __fdiv_rn(ys, yy);                   // H_diag = ys / y.dot(y)
            if (tid == 0) ROw[k] = ro_new;
            k += 1;
        }
        // two-loop recursion; G already holds flat_grad
        ld8(q, G, tid, n4);
#pragma unroll
        for (int i = 0; i < kPer; ++i) q[i] = -q[i];
        int r = 0;                                    // rows consumed from the ring
        float row[kPer];
        uint32_t ring_s = smem_u32(ring) + (uint32_t)tid * 16u;            // this thread's first float4 of slot 0
        uint32_t bars_s = smem_u32(full_bar);
        asm volatile("" : "+r"(ring_s), "+r"(bars_s));                     // opaque: kept in registers, not re-derived per row
        auto next_row = [&]() {
            const uint32_t g = rows_done + (uint32_t)r;           // rows this group's ring has carried so far
            const uint32_t slot = g % kRing;
#if GEM_LBFGS_SHARED_ADDR
            mbar_wait_s(bars_s + slot * 8u, (g / kRing) & 1u);
            lds8(row, ring_s + slot * row_bytes, tid, n4);
#else
            mbar_wait(&full_bar[slot], (g / kRing) & 1u);
            ld8(row, ring + (size_t)slot * n, tid, n4);
#endif
            ++r;
        };
        if (pushed) {
            float r1[1] = {dot8(sv, q)};
            block_reduce<1, 0>(r1, red, par, tid, gid);
            const float a = r1[0] * ro_new;
            if (tid == 0) al_s[k - 1] = a;
            const float na = -a;
#pragma unroll
            for (int i = 0; i < kPer; ++i) q[i] = q[i] + yv[i] * na;
        }
        for (int h = kr - 1; h >= 0; --h) {
            next_row();                               // s_h
            float r1[1] = {dot8(row, q)};
            block_reduce<1, 0>(r1, red, par, tid, gid);
            if (tid == 0) issue_upto(r + kRing);      // every row < r has been consumed by the whole CTA
            const float a = r1[0] * ro_s[h];
            if (tid == 0) al_s[h] = a;
            const float na = -a;
            next_row();                               // y_h
#pragma unroll
            for (int i = 0; i < kPer; ++i) q[i] = q[i] + row[i] * na;
        }
#pragma unroll
        for (int i = 0; i < kPer; ++i) q[i] = q[i] * hd;
        for (int h = 0; h < kr; ++h) {
            next_row();                               // y_h
            float r1[1] = {dot8(row, q)};
            block_reduce<1, 0>(r1, red, par, tid, gid);
            if (tid == 0) issue_upto(r + kRing);
            const float c = al_s[h] - r1[0] * ro_s[h];
            next_row();                               // s_h
#pragma unroll
            for (int i = 0; i < kPer; ++i) q[i] = q[i] + row[i] * c;
        }
        if (pushed) {
            float r1[1] = {dot8(yv, q)};
            block_reduce<1, 0>(r1, red, par, tid, gid);
            const float c = al_s[k - 1] - r1[0] * ro_new;
#pragma unroll
            for (int i = 0; i < kPer; ++i) q[i] = q[i] + sv[i] * c;
        }
        rows_done += (uint32_t)total;                 // every issued row was consumed
    }
    // d = q; prev_flat_grad = flat_grad (G itself); gtd = g.d; t
    st8(D, q, tid, n4);
    float r3[3] = {0.f, 0.f, 0.f};
    {
        float g[kPer];
        if (first_iter) {
#pragma unroll
            for (int i = 0; i < kPer; ++i) g[i] = gn[i];
        } else {
            ld8(g, G, tid, n4);
        }
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
            r3[0] += g[i] * q[i];
            r3[1] += fabsf(g[i]);
            r3[2] = fmaxf(r3[2], fabsf(q[i]));
        }
    }
    block_reduce<3, 4>(r3, red, par, tid, gid);
    if (tid == 0) {
        const float gtd = r3[0], l1 = r3[1], dmax = r3[2];
        s.n_iter = n_iter0 + 1;
        s.hist_len = k;
        s.h_diag = hd;
        s.gp_is_g = 1;
        s.prev_loss = s.loss;
        Num t;
        if (first_iter) {
            // t = min(1., 1. / flat_grad.abs().sum()) * lr
            const Num inv = t32(__frcp_rn(l1));
            t = nmul(nminp(py(1.0), inv), py(b.lr));
        } else {
            t = py(b.lr);
        }
        s.t = t;
        s.gtd = gtd;
        if ((double)gtd > (double)(float)(-b.tol_change)) {
            s.phase = PH_DONE;                            // directional derivative below tolerance
        } else {
            s.d_norm = dmax;
            s.t_prev = py(0.0);
            s.f_prev = s.loss;
            s.gtd_prev = gtd;
            s.ls_iter = 0;
            s.ls_evals = 0;
            s.ls_first = 1;
            s.ls_done = 0;
            s.max_ls = b.max_eval - s.evals;
            s.phase = PH_BRACKET;
        }
        sh_done2 = s.phase == PH_DONE;
        sh_tf2 = (float)t.v;
        b.st[w] = s;
    }
    group_sync(gid);
    {
        float x[kPer];
        ld8(x, X, tid, n4);
        if (!sh_done2) {
            const float tf = sh_tf2;
#pragma unroll
            for (int i = 0; i < kPer; ++i) x[i] = x[i] + tf * q[i];
        }
        st_trial(b, off, x, tid, n4);
    }
}

// One CTA per window (the original launch shape): four CTAs per SM, spread over every SM of the GPU.
__global__ void __launch_bounds__(kLbThreads, GEM_LBFGS_OCC) lbfgs_advance_kernel(LbfgsBuffers b, const float* __restrict__ loss_in,
                                                                     const float* __restrict__ grad_in, int W) {
    extern __shared__ __align__(16) float ring[];     // [kRing][n]
    __shared__ GroupSmem sm;
    uint32_t rows_done = 0;
    lbfgs_advance_window<false>(b, loss_in, grad_in, blockIdx.x, threadIdx.x, 0, sm, ring, true, rows_done);
}

// Persistent variant (opt-in, GEM_LBFGS_SMS): a CTA of kGroups independent groups fills one SM (1024 threads x 64
// registers) and walks over windows, so a launch of G CTAs occupies exactly G SMs.  Small CTAs spread over every SM,
// and an SM that holds even one of them has no room for a 200 KB tensor-core CTA of another slice: measured, the
// L-BFGS update does not overlap the tensor-bound layers at all (a step without it is 7.1 ms shorter, its own time is
// 8.2 ms).  Packing it onto fewer SMs does not pay either, see launch_lbfgs_advance.
constexpr int kGroups = 4;
__global__ void __launch_bounds__(kGroups * kLbThreads, 1)
lbfgs_advance_persistent_kernel(LbfgsBuffers b, const float* __restrict__ loss_in, const float* __restrict__ grad_in, int W) {
    extern __shared__ __align__(16) float ring[];     // [kGroups][kRing][n]
    __shared__ GroupSmem sm[kGroups];
    const int gid = threadIdx.x / kLbThreads, tid = threadIdx.x - gid * kLbThreads;
    float* my_ring = ring + (size_t)gid * kRing * b.n;
    bool first = true;
    uint32_t rows_done = 0;
    for (int w = blockIdx.x * kGroups + gid; w < W; w += gridDim.x * kGroups) {
        lbfgs_advance_window<true>(b, loss_in, grad_in, w, tid, gid, sm[gid], my_ring, first, rows_done);
        group_sync(gid);          // the group's shared state is reused by the next window
        first = false;
    }
}

__global__ void lbfgs_stats_kernel(const LbfgsWin* st, int W, int32_t* n_iter, int32_t* evals, int32_t* finished,
                                   double* t, double* loss) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= W) return;
    const LbfgsWin s = st[w];
    if (n_iter) n_iter[w] = s.n_iter;
    if (evals) evals[w] = s.evals + ((s.phase == PH_BRACKET || s.phase == PH_ZOOM) ? s.ls_evals : 0);
    if (finished) finished[w] = s.phase == PH_DONE;
    if (t) t[w] = s.t.v;
    if (loss) loss[w] = s.loss;
}

size_t lbfgs_state_bytes() { return sizeof(LbfgsWin); }

int launch_lbfgs_begin(cudaStream_t stream, const LbfgsBuffers& b, const float* z0, int W) {
    if (W <= 0) return GEM_OK;
    GEM_REQUIRE(b.n % 4 == 0 && b.n <= kLbThreads * kPer, "latent_dim must be a multiple of 4, <= 2048");
    GEM_REQUIRE(b.m <= kMaxHist, "max_history must be <= 64");
    lbfgs_begin_kernel<<<W, kLbThreads, 0, stream>>>(b, z0, W);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}
int launch_lbfgs_advance(cudaStream_t stream, const LbfgsBuffers& b, const float* loss, const float* grad, int W) {
    if (W <= 0) return GEM_OK;
    // one small CTA per window spread over the whole GPU (default), or persistent CTAs on GEM_LBFGS_SMS SMs.  The
    // persistent shape was meant to leave SMs to other slices' tensor-core CTAs; measured, it loses: a group moves one
    // 8 KB row per ~1 us (its four-row ring against a loaded HBM latency of ~4 us), so the kernel needs the
    // shared memory of ALL SMs in flight to reach 4.9 TB/s, and on 64 SMs it takes 2.3x as long (step 22.1 -> 23.2 ms).
    static const int sms = []() {
        const char* env = getenv("GEM_LBFGS_SMS");
        const int v = env ? atoi(env) : 0;
        return v < 0 ? 0 : (v > kNumSMs ? kNumSMs : v);
    }();
    const size_t smem = (size_t)kRing * b.n * sizeof(float);
    if (sms > 0) {
        static PerDeviceOnce attr_set;
        if (bool* once_ = attr_set.flag(); !*once_) {
            GEM_CUDA(cudaFuncSetAttribute(lbfgs_advance_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(kGroups * kRing * 2048 * sizeof(float))));
            *once_ = true;
        }
        int grid = (W + kGroups - 1) / kGroups;
        if (grid > sms) grid = sms;
        lbfgs_advance_persistent_kernel<<<grid, kGroups * kLbThreads, kGroups * smem, stream>>>(b, loss, grad, W);
    } else {
        static PerDeviceOnce attr1_set;
        if (bool* once_ = attr1_set.flag(); !*once_) {
            GEM_CUDA(cudaFuncSetAttribute(lbfgs_advance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)(kRing * 2048 * sizeof(float))));
            *once_ = true;
        }
        lbfgs_advance_kernel<<<W, kLbThreads, smem, stream>>>(b, loss, grad, W);
    }
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}
int launch_lbfgs_stats(cudaStream_t stream, const LbfgsBuffers& b, int W, int32_t* n_iter, int32_t* evals,
                       int32_t* finished, double* t, double* loss) {
    if (W <= 0) return GEM_OK;
    lbfgs_stats_kernel<<<(W + 127) / 128, 128, 0, stream>>>(b.st, W, n_iter, evals, finished, t, loss);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}

}  // namespace gem
