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
// HBM) is spread over the CTA, the scalar decision logic runs on thread 0.  torch mixes Python
// floats (double) with fp32 0-d tensors; `Num` reproduces that typing so that step lengths round
// the way the reference's do.  Compiled with --fmad=false.
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
    int ls_done, insuf, low_pos, high_pos, nbracket, pad0;
    double loss, prev_loss;
    float h_diag, gtd, d_norm, pad1;
    Num t, t_prev, br[2];
    double f_prev, br_f[2];
    float gtd_prev, br_gtd[2], pad2;
};

constexpr int kLbThreads = 256;
constexpr int kMaxPerThread = 8;   // n <= 2048 keeps a vector in registers

// block-wide sum / max with a fixed reduction tree (deterministic run to run)
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < kLbThreads / 32; ++i) r += red[i];
    return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < kLbThreads / 32; ++i) r = fmaxf(r, red[i]);
    return r;
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
    }
    if (threadIdx.x == 0) {
        LbfgsWin s;
        memset(&s, 0, sizeof(s));
        s.phase = PH_INIT;
        s.h_diag = 1.f;
        b.st[w] = s;
    }
}

// Per-thread strided ownership of vector elements: element e = threadIdx.x + i*kLbThreads.
#define FOR_OWN(i, e) for (int i = 0, e = threadIdx.x; i < kMaxPerThread && e < b.n; ++i, e += kLbThreads)

__global__ void __launch_bounds__(kLbThreads) lbfgs_advance_kernel(LbfgsBuffers b, const float* __restrict__ loss_in,
                                                                  const float* __restrict__ grad_in, int W) {
    __shared__ float red[kLbThreads / 32];
    __shared__ LbfgsWin s;          // thread 0 mutates, everyone reads after barriers
    __shared__ int action;          // what the vector phase has to do next
    __shared__ float bcast[4];

    const int w = blockIdx.x;
    const size_t off = (size_t)w * b.n;
    float *X = b.X + off, *D = b.D + off, *G = b.G + off, *PG = b.PG + off, *GP = b.GP + off, *BG0 = b.BG0 + off,
          *BG1 = b.BG1 + off, *ZT = b.ZT + off;
    float* BG[2] = {BG0, BG1};

    if (threadIdx.x == 0) s = b.st[w];
    __syncthreads();
    if (s.phase == PH_DONE) return;

    const double f_new = (double)loss_in[w];
    float gn[kMaxPerThread];
    FOR_OWN(i, e) gn[i] = grad_in[off + e];
    if (threadIdx.x == 0 && b.trace) {
        const int idx = s.evals + ((s.phase == PH_INIT) ? 0 : s.ls_evals);
        if (idx < b.trace_stride) b.trace[(size_t)w * b.trace_stride + idx] = (float)f_new;
    }

    enum { ACT_NEW_ITER = 1, ACT_EVAL = 2, ACT_LS_FINISH = 3, ACT_STOP = 4 };
    bool new_iter = false;

    if (s.phase == PH_INIT) {
        // lbfgs.py:361-373: first evaluation
        float gmax = 0.f;
        FOR_OWN(i, e) {
            G[e] = gn[i];
            gmax = fmaxf(gmax, fabsf(gn[i]));
        }
        gmax = block_max(gmax, red);
        if (threadIdx.x == 0) {
            s.loss = f_new;
            s.evals = 1;
            s.n_iter = 0;
            if ((double)gmax <= (double)(float)b.tol_grad) s.phase = PH_DONE;
        }
        __syncthreads();
        if (s.phase == PH_DONE) {
            if (threadIdx.x == 0) b.st[w] = s;
            return;
        }
        new_iter = true;
    } else {
        // ---- an evaluation requested by the line search has arrived ----------------------------
        float part = 0.f;
        FOR_OWN(i, e) part += gn[i] * D[e];
        const float gtd_new = block_sum(part, red);
        // The scalar logic may need several vector copies; it is run as a small loop of
        // (decide on thread 0) -> (vector action by all threads).
        int copy_gp_from_new = 0;     // GP = g_new.clone()
        if (threadIdx.x == 0) {
            const double c1 = 1e-4, c2 = 0.9;
            const Num gtd = t32(s.gtd);
            int act = 0;
            s.ls_evals += 1;
            bool to_zoom_entry = false;
            int set_br = 0;   // 1: bracket=[t_prev,t] with g=[GP,g_new]; 2: bracket=[t] g=[g_new]; 3: [0,t] g=[G,g_new]
            if (s.phase == PH_BRACKET) {
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
                        copy_gp_from_new = 1;
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
                    } else if (set_br == 2) {
                        s.br[0] = s.t;
                        s.br_f[0] = f_new;
                        s.nbracket = 1;
                    } else {
                        s.br[0] = py(0.0), s.br[1] = s.t;
                        s.br_f[0] = s.loss, s.br_f[1] = f_new;
                        s.nbracket = 2;
                    }
                    s.insuf = 0;
                    const double flast = s.br_f[s.nbracket - 1];
                    if (s.br_f[0] <= flast) s.low_pos = 0, s.high_pos = 1;
                    else s.low_pos = 1, s.high_pos = 0;
                    s.phase = PH_ZOOM;
                }
                bcast[0] = (float)set_br;
            } else {   // PH_ZOOM: an evaluation inside the zoom loop (lbfgs.py:152-199)
                s.ls_iter += 1;
                const Num rhs = nadd(py(s.loss), nmul(nmul(py(c1), s.t), gtd));
                if (ngt(py(f_new), rhs) || f_new >= s.br_f[s.low_pos]) {
                    s.br[s.high_pos] = s.t;
                    s.br_f[s.high_pos] = f_new;
                    s.br_gtd[s.high_pos] = gtd_new;
                    bcast[0] = (float)(10 + s.high_pos);      // BG[high] = g_new
                    if (s.br_f[0] <= s.br_f[1]) s.low_pos = 0, s.high_pos = 1;
                    else s.low_pos = 1, s.high_pos = 0;
                } else {
                    int code = 20 + s.low_pos;                // BG[low] = g_new
                    if (nle(t32(fabsf(gtd_new)), nmul(py(-c2), gtd))) {
                        s.ls_done = 1;
                    } else if (nge(nmul(t32(gtd_new), nsub(s.br[s.high_pos], s.br[s.low_pos])), py(0.0))) {
                        s.br[s.high_pos] = s.br[s.low_pos];
                        s.br_f[s.high_pos] = s.br_f[s.low_pos];
                        s.br_gtd[s.high_pos] = s.br_gtd[s.low_pos];
                        code = 30 + s.low_pos;                // BG[high] = BG[low]; BG[low] = g_new
                    }
                    s.br[s.low_pos] = s.t;
                    s.br_f[s.low_pos] = f_new;
                    s.br_gtd[s.low_pos] = gtd_new;
                    bcast[0] = (float)code;
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
            action = act;
            bcast[1] = (float)copy_gp_from_new;
        }
        __syncthreads();
        // ---- vector side effects of the decision ---------------------------------------------
        {
            const int code = (int)bcast[0];
            if (s.phase == PH_BRACKET || (s.phase == PH_ZOOM && code >= 1 && code <= 3)) {
                if (bcast[1] != 0.f) FOR_OWN(i, e) GP[e] = gn[i];
                if (code == 1) {            // bracket_g = [g_prev, g_new.clone()]
                    FOR_OWN(i, e) { BG0[e] = GP[e]; BG1[e] = gn[i]; }
                } else if (code == 2) {     // bracket_g = [g_new]
                    FOR_OWN(i, e) BG0[e] = gn[i];
                } else if (code == 3) {     // bracket_g = [g, g_new]
                    FOR_OWN(i, e) { BG0[e] = G[e]; BG1[e] = gn[i]; }
                }
            } else if (code >= 10 && code < 20) {
                float* dst = BG[code - 10];
                FOR_OWN(i, e) dst[e] = gn[i];
            } else if (code >= 20 && code < 30) {
                float* dst = BG[code - 20];
                FOR_OWN(i, e) dst[e] = gn[i];
            } else if (code >= 30) {
                float* lo = BG[code - 30];
                float* hi = BG[1 - (code - 30)];
                FOR_OWN(i, e) { hi[e] = lo[e]; lo[e] = gn[i]; }
            }
        }
        if (action == ACT_EVAL) {
            const float tf = (float)s.t.v;
            FOR_OWN(i, e) ZT[e] = X[e] + tf * D[e];          // _add_grad(t, d) from x_init
            if (threadIdx.x == 0) b.st[w] = s;
            return;
        }
        // ---- line search finished (lbfgs.py:201-209, 488-527) ------------------------------
        __syncthreads();
        const float* gsel = BG[s.low_pos];
        const float tf = (float)s.br[s.low_pos].v;
        float gmax = 0.f, stepmax = 0.f;
        FOR_OWN(i, e) {
            const float gv = gsel[e];
            G[e] = gv;
            gmax = fmaxf(gmax, fabsf(gv));
            const float step = tf * D[e];
            X[e] = X[e] + step;
            stepmax = fmaxf(stepmax, fabsf(step));
        }
        gmax = block_max(gmax, red);
        stepmax = block_max(stepmax, red);
        if (threadIdx.x == 0) {
            s.t = s.br[s.low_pos];
            s.loss = s.br_f[s.low_pos];
            s.evals += s.ls_evals;
            bool stop = false;
            if (s.n_iter == b.max_iter) stop = true;
            else if (s.evals >= b.max_eval) stop = true;
            else if ((double)gmax <= (double)(float)b.tol_grad) stop = true;
            else if ((double)stepmax <= (double)(float)b.tol_change) stop = true;
            else if (fabs(s.loss - s.prev_loss) < b.tol_change) stop = true;
            if (stop) s.phase = PH_DONE;
        }
        __syncthreads();
        if (s.phase == PH_DONE) {
            FOR_OWN(i, e) ZT[e] = X[e];
            if (threadIdx.x == 0) b.st[w] = s;
            return;
        }
        new_iter = true;
    }

    if (new_iter) {
        // ---- next outer iteration: direction (lbfgs.py:388-442), step, first trial point ------
        __syncthreads();
        if (threadIdx.x == 0) s.n_iter += 1;
        __syncthreads();
        float dreg[kMaxPerThread];
        if (s.n_iter == 1) {
            FOR_OWN(i, e) dreg[i] = -G[e];
        } else {
            const float tprev = (float)s.t.v;
            float* Yw = b.Y + (size_t)w * b.m * b.n;
            float* Sw = b.S + (size_t)w * b.m * b.n;
            float* ROw = b.RO + (size_t)w * b.m;
            float yv[kMaxPerThread], sv[kMaxPerThread];
            float p_ys = 0.f, p_yy = 0.f;
            FOR_OWN(i, e) {
                yv[i] = G[e] - PG[e];
                sv[i] = D[e] * tprev;
                p_ys += yv[i] * sv[i];
                p_yy += yv[i] * yv[i];
            }
            const float ys = block_sum(p_ys, red);
            const float yy = block_sum(p_yy, red);
            int k = s.hist_len;
            if ((double)ys > (double)1e-10f) {
                if (k == b.m) {
                    // history full (never reached with max_iter <= m+1): drop the oldest pair
                    for (int h = 1; h < k; ++h) {
                        FOR_OWN(i, e) {
                            Yw[(size_t)(h - 1) * b.n + e] = Yw[(size_t)h * b.n + e];
                            Sw[(size_t)(h - 1) * b.n + e] = Sw[(size_t)h * b.n + e];
                        }
                        if (threadIdx.x == 0) ROw[h - 1] = ROw[h];
                    }
                    k -= 1;
                }
                FOR_OWN(i, e) {
                    Yw[(size_t)k * b.n + e] = yv[i];
                    Sw[(size_t)k * b.n + e] = sv[i];
                }
                if (threadIdx.x == 0) {
                    ROw[k] = __frcp_rn(ys);               // 1. / ys  (reciprocal)
                    s.h_diag = __fdiv_rn(ys, yy);
                }
                k += 1;
                __syncthreads();
                if (threadIdx.x == 0) s.hist_len = k;
            }
            __syncthreads();
            // two-loop recursion
            float q[kMaxPerThread];
            float al[64];
            FOR_OWN(i, e) q[i] = -G[e];
            for (int h = k - 1; h >= 0; --h) {
                float part = 0.f;
                FOR_OWN(i, e) part += Sw[(size_t)h * b.n + e] * q[i];
                const float a = block_sum(part, red) * ROw[h];
                if (h < 64) al[h] = a;
                const float na = -a;
                FOR_OWN(i, e) q[i] = q[i] + Yw[(size_t)h * b.n + e] * na;
            }
            const float hd = s.h_diag;
            FOR_OWN(i, e) q[i] = q[i] * hd;
            for (int h = 0; h < k; ++h) {
                float part = 0.f;
                FOR_OWN(i, e) part += Yw[(size_t)h * b.n + e] * q[i];
                const float be = block_sum(part, red) * ROw[h];
                const float c = al[h] - be;
                FOR_OWN(i, e) q[i] = q[i] + Sw[(size_t)h * b.n + e] * c;
            }
            FOR_OWN(i, e) dreg[i] = q[i];
        }
        // prev_flat_grad = flat_grad; prev_loss = loss; t; gtd = g.d
        float p_gtd = 0.f, p_l1 = 0.f, p_dmax = 0.f;
        FOR_OWN(i, e) {
            const float gv = G[e];
            D[e] = dreg[i];
            PG[e] = gv;
            GP[e] = gv;                                       // g_prev = g (clone) for the bracket phase
            p_gtd += gv * dreg[i];
            p_l1 += fabsf(gv);
            p_dmax = fmaxf(p_dmax, fabsf(dreg[i]));
        }
        const float gtd = block_sum(p_gtd, red);
        const float l1 = block_sum(p_l1, red);
        const float dmax = block_max(p_dmax, red);
        if (threadIdx.x == 0) {
            s.prev_loss = s.loss;
            Num t;
            if (s.n_iter == 1) {
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
        }
        __syncthreads();
        if (s.phase == PH_DONE) {
            FOR_OWN(i, e) ZT[e] = X[e];
        } else {
            const float tf = (float)s.t.v;
            FOR_OWN(i, e) ZT[e] = X[e] + tf * dreg[i];
        }
        if (threadIdx.x == 0) b.st[w] = s;
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
    GEM_REQUIRE(b.n <= kLbThreads * kMaxPerThread, "latent_dim must be <= 2048");
    GEM_REQUIRE(b.m <= 64, "max_history must be <= 64");
    lbfgs_begin_kernel<<<W, kLbThreads, 0, stream>>>(b, z0, W);
    GEM_CHECK_LAUNCH();
    return GEM_OK;
}
int launch_lbfgs_advance(cudaStream_t stream, const LbfgsBuffers& b, const float* loss, const float* grad, int W) {
    if (W <= 0) return GEM_OK;
    lbfgs_advance_kernel<<<W, kLbThreads, 0, stream>>>(b, loss, grad, W);
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
