// ctb_phases.cuh -- batched N-field minimum finder / phase follower on the device.
//
// What it replaces in the reference (one scipy.optimize.fmin call -- Nelder-Mead -- per point and temperature):
//   generic_potential.findMinimum                         generic_potential.py:477-484
//   transitionFinder.traceMinimum (fmin corrector)        transitionFinder.py:34-128 (fmin at :127-129)
//   the fmin polish of both minima before a tunnelling    transitionFinder.py:658-664
//
// One WARP owns one (model-parameter row, starting point) and walks a temperature grid.  At every temperature it runs a
// damped Newton iteration on the finite-difference gradient and Hessian of Vtot -- the reference's own derivative
// definitions, helper_functions.gradientFunction / hessianFunction with order 4 and step x_eps (helper_functions.py:
// 446-461, 505-554; generic_potential.py:347-442) -- with ONE STENCIL POINT PER LANE: the 1 + 4 N + 8 N (N - 1)
// evaluations of Vtot that one gradient + Hessian needs (25 for two fields) are independent, so the lanes evaluate them
// side by side and the weighted sums are formed with warp shuffles; the N x N Cholesky solve runs redundantly in every
// lane's registers; the backtracking line search evaluates all step halvings at once (lane k tries 2^-k of the step) and
// a ballot picks the first one that lowers V.  A Newton iteration therefore costs two Vtot latencies instead of ~27.
// The walk is warm-started from a linear extrapolation of the previous two temperatures.  A temperature at which the
// iterate leaves the neighbourhood of the traced minimum (the phase has ended and the iterate fell into another basin)
// or ends on a point whose Hessian is not positive definite ends the trace, like the singular-Hessian test of
// traceMinimum (transitionFinder.py:118-126); the point it ended on is reported, so the host can start the next phase
// there (what traceMultiMin does with its fmin call at transitionFinder.py:466-500).
#pragma once
#include "ctb_device.cuh"

namespace ctb {

// The staged Jb / Jf tables as a model's V sees them (jf.nint = 0: the model has no fermions, nothing staged)
struct Tabs { JTab jb, jf; };

// A MODEL is a struct with
//   static constexpr int ndim, nparam;   number of fields, doubles per parameter row
//   void load(const double *row);        take one parameter row
//   void set_T(double T);                temperature-dependent constants
//   double V(const Tabs &, const double *X) const;   Vtot(X, T) = V0 + V1 + V1T (GP:240-337)
// Built in: Model1XY.  Others are compiled at run time from a few lines of CUDA (cosmotransitions_b200/jit.py).

// model1 (examples/testModel1.py:62-116) at a two-field point: V0 + V1 + V1T (GP:240-337)
struct Model1XY {
    static constexpr int ndim = 2;
    static constexpr int nparam = 8;
    double l1, l2, mu2, Y1, Y2, nb, inv_rs, v2;
    double inv_T2, T4c;
    CTB_DEV void load(const double *p)
    {
        l1 = p[0]; l2 = p[1]; mu2 = p[2]; Y1 = p[3]; Y2 = p[4]; nb = p[5]; inv_rs = 1.0 / p[6]; v2 = p[7];
    }
    CTB_DEV void set_T(double T)
    {
        const double T2 = T * T;
        inv_T2 = 1.0 / (T2 + 1e-100);
        T4c = (T2 * T2) / (2 * CTB_PI * CTB_PI);
    }
    CTB_DEV double V(const Tabs &tb, const double *X) const
    {
        const JTab &jb = tb.jb;
        const Model1Field f = model1_field(l1, l2, mu2, Y1, Y2, nb, inv_rs, v2, X[0], X[1]);
        const double yt = jtab_eval<0>(jb, f.m0 * inv_T2) + jtab_eval<0>(jb, f.m1 * inv_T2) + nb * jtab_eval<0>(jb, f.mb * inv_T2);
        return f.V01 + yt * T4c;
    }
};

// The general one-field model (CTB_POT_THERMAL1F row: quartic V0, <= 8 bosons with thermal masses, <= 4 fermions;
// generic_potential.py:240-337) as a one-dimensional Model: the parameter row stays in global memory (L1-resident)
struct Thermal1FX {
    static constexpr int ndim = 1;
    static constexpr int nparam = CTB_NPARAM_THERMAL1F;
    const double *p;
    double inv_rs, T2, inv_T2, T4c;
    int nb, nf;
    CTB_DEV void load(const double *q)
    {
        p = q;
        inv_rs = 1.0 / q[5];
        nb = min(max((int)q[7], 0), CTB_THERMAL1F_MAX_BOSONS);
        nf = min(max((int)q[8], 0), CTB_THERMAL1F_MAX_FERMIONS);
    }
    CTB_DEV void set_T(double T)
    {
        T2 = T * T;
        inv_T2 = 1.0 / (T2 + 1e-100);
        T4c = (T2 * T2) / (2 * CTB_PI * CTB_PI);
    }
    CTB_DEV double V(const Tabs &tb, const double *X) const
    {
        const PotThermal1F::Parts o = PotThermal1F::eval(p, nb, nf, inv_rs, T2, inv_T2, T4c, tb.jb, tb.jf, X[0]);
        return o.v0 + o.v1 + o.v1t * T4c;
    }
};

#define CTB_TRACE_SUBSTEPS 8
struct MinKnobs {
    double x_eps;     // finite-difference step (generic_potential.x_eps, 1e-3)
    double xtol;      // converged when |step| <= xtol (|X| + 1)
    double jump_tol;  // trace: a minimum farther than jump_tol (|X_start| + 1) from the previous one is another phase
    int max_iter;
};

// In-place Cholesky of the symmetric N x N matrix A + lam I; false if it is not positive definite.
template <int N>
CTB_DEV bool cholesky(const double (&A)[N][N], double lam, double (&L)[N][N])
{
#pragma unroll
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int j = 0; j <= i; j++) {
            double s = A[i][j] + (i == j ? lam : 0.0);
#pragma unroll
            for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
            if (i == j) {
                if (!(s > 0.0)) return false;
                L[i][i] = sqrt(s);
            } else {
                L[i][j] = s / L[j][j];
            }
        }
    }
    return true;
}

template <int N>
CTB_DEV void cholesky_solve(const double (&L)[N][N], const double (&b)[N], double (&x)[N])
{
    double y[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; k++) s -= L[i][k] * y[k];
        y[i] = s / L[i][i];
    }
#pragma unroll
    for (int i = N - 1; i >= 0; i--) {
        double s = y[i];
#pragma unroll
        for (int k = i + 1; k < N; k++) s -= L[k][i] * x[k];
        x[i] = s / L[i][i];
    }
}

template <int N> CTB_DEV double norm_n(const double (&v)[N])
{
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) s += v[i] * v[i];
    return sqrt(s);
}

CTB_DEV double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Gradient and Hessian by the reference's 4th-order stencils, one stencil point per lane.  Point 0 is X itself
// (weight -30 in every diagonal second derivative), points 1 .. 4N move one coordinate by {-2,-1,1,2} h (weights
// {1,-8,8,-1}/12h for the gradient and {-1,16,16,-1}/12h^2 for H_ii), the rest move a pair (i < j) by the 4 x 4
// products (weights c_a c_b / 144 h^2 for H_ij).  Every lane returns the same f0, g, H.
template <class Model>
__device__ __forceinline__ void warp_grad_hess(const Model &m, const Tabs &jb, const double (&X)[Model::ndim], double h,
                                               int lane, double &f0, double (&g)[Model::ndim],
                                               double (&H)[Model::ndim][Model::ndim])
{
    constexpr int N = Model::ndim;
    constexpr int NP = 1 + 4 * N + 8 * N * (N - 1);
    double fc = 0.0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        g[i] = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) H[i][j] = 0.0;
    }
#pragma unroll 1
    for (int p = lane; p < NP; p += 32) {
        // decode the stencil point: coordinates ii / jj move by da / db steps
        int ii = -1, jj = -1, a = 0, b = 0;
        if (p >= 1 + 4 * N) {
            const int q = p - 1 - 4 * N;
            int pr = q >> 4;
            a = (q >> 2) & 3; b = q & 3;
            // pair index -> (i < j), row-major over the upper triangle
            ii = 0; jj = 1;
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int j = i + 1; j < N; j++) {
                    if (pr == 0) { ii = i; jj = j; }
                    pr--;
                }
        } else if (p >= 1) {
            ii = (p - 1) >> 2; a = (p - 1) & 3;
        }
        const double da = (a < 2 ? (double)(a - 2) : (double)(a - 1));      // -2, -1, 1, 2
        const double db = (b < 2 ? (double)(b - 2) : (double)(b - 1));
        const double ca = (a == 0 ? 1.0 : a == 1 ? -8.0 : a == 2 ? 8.0 : -1.0);
        const double cb = (b == 0 ? 1.0 : b == 1 ? -8.0 : b == 2 ? 8.0 : -1.0);
        const double c2 = (a == 0 || a == 3) ? -1.0 : 16.0;
        double Y[N];
#pragma unroll
        for (int i = 0; i < N; i++) Y[i] = X[i] + (i == ii ? da * h : 0.0) + (i == jj ? db * h : 0.0);
        const double v = m.V(jb, Y);
        if (p == 0) fc = v;
#pragma unroll
        for (int i = 0; i < N; i++) {
            if (p == 0) H[i][i] += -30.0 * v;
            if (jj < 0 && i == ii) { g[i] += ca * v; H[i][i] += c2 * v; }
#pragma unroll
            for (int j = i + 1; j < N; j++)
                if (i == ii && j == jj) H[i][j] += (ca * cb) * v;
        }
    }
    const double i12h = 1.0 / (12.0 * h), i12h2 = 1.0 / (12.0 * h * h);
    f0 = __shfl_sync(0xffffffffu, fc, 0);
#pragma unroll
    for (int i = 0; i < N; i++) {
        g[i] = warp_sum(g[i]) * i12h;
        H[i][i] = warp_sum(H[i][i]) * i12h2;
#pragma unroll
        for (int j = i + 1; j < N; j++) H[i][j] = H[j][i] = warp_sum(H[i][j]) * (i12h * i12h);
    }
}

// Damped Newton from X (in/out), the whole warp on one point.  Returns CTB_MIN_OK, CTB_MIN_NOT_CONVERGED or
// CTB_MIN_NOT_A_MINIMUM (converged with a lifted Hessian: a saddle or a maximum); fmin = V at the result.
template <class Model>
__device__ __forceinline__ int warp_newton_minimize(const Model &m, const Tabs &jb, const MinKnobs &k, int lane,
                                                    double (&X)[Model::ndim], double &fmin, int &iters)
{
    constexpr int N = Model::ndim;
    double g[N], H[N][N], L[N][N], step[N], Y[N];
    double f;
    int status = CTB_MIN_NOT_CONVERGED;
    iters = 0;
    for (int it = 0; it < k.max_iter; it++) {
        iters = it + 1;
        warp_grad_hess(m, jb, X, k.x_eps, lane, f, g, H);
        // lift a Hessian that is not positive definite (saddle / maximum side of the start): H + lam I
        double hmax = 0.0;
#pragma unroll
        for (int i = 0; i < N; i++) hmax = fmax(hmax, fabs(H[i][i]));
        double lam = 0.0;
        bool pd = cholesky<N>(H, 0.0, L);
        for (int tries = 0; !pd && tries < 60; tries++) {
            lam = (lam == 0.0) ? 1e-3 * hmax + 1e-12 : 2.0 * lam;
            pd = cholesky<N>(H, lam, L);
        }
        if (!pd) { fmin = f; return CTB_MIN_NOT_CONVERGED; } // non-finite Hessian
        double mg[N];
#pragma unroll
        for (int i = 0; i < N; i++) mg[i] = -g[i];
        cholesky_solve<N>(L, mg, step);
        double sn = norm_n<N>(step);
        const double cap = 0.1 * (norm_n<N>(X) + 1.0);
        if (sn > cap) {
#pragma unroll
            for (int i = 0; i < N; i++) step[i] *= cap / sn;
            sn = cap;
        }
        // backtracking: lane k tries 2^-k of the step; the first k that does not raise V (beyond rounding noise) wins
        const int kk = lane < 29 ? lane : 29;
        const double tk = __hiloint2double((1023 - kk) << 20, 0);
#pragma unroll
        for (int i = 0; i < N; i++) Y[i] = X[i] + tk * step[i];
        const double fk = m.V(jb, Y);
        const unsigned okm = __ballot_sync(0xffffffffu, !(fk > f + 1e-13 * fabs(f))) & 0x3fffffffu;
        const int ks = okm ? (__ffs(okm) - 1) : 29;
        const double ts = okm ? __hiloint2double((1023 - ks) << 20, 0) : __hiloint2double((1023 - 30) << 20, 0);
        const double fn = __shfl_sync(0xffffffffu, fk, ks);
#pragma unroll
        for (int i = 0; i < N; i++) { step[i] *= ts; X[i] += step[i]; }
        sn *= ts;
        fmin = fn;
        if (!(sn == sn)) return CTB_MIN_NOT_CONVERGED; // NaN
        if (sn <= k.xtol * (norm_n<N>(X) + 1.0)) {
            status = (lam == 0.0) ? CTB_MIN_OK : CTB_MIN_NOT_A_MINIMUM;
            break;
        }
    }
    return status;
}

// One warp = one (parameter row, start point).  It starts at temperature index t_begin[i] (default 0) with X0[i] and
// walks t += t_dir[i] (default +1) to the end of the grid; T is one shared grid (T_stride = 0) or one row per instance.
// Entries it does not reach are NaN / CTB_MIN_NOT_TRACED.  param_stride = 0: one model for all instances.
template <class Model>
__global__ void __launch_bounds__(256) trace_minima_kernel(const double *__restrict__ params, int64_t param_stride, JTab g,
                                                          JTab gf,
                                                          const double *__restrict__ X0, const double *__restrict__ T,
                                                          int64_t T_stride, const int32_t *__restrict__ t_begin,
                                                          const int32_t *__restrict__ t_dir, int64_t n, int nT, MinKnobs k,
                                                          double *__restrict__ Xout, double *__restrict__ Vout,
                                                          int32_t *__restrict__ flag, int32_t *__restrict__ iters_out)
{
    constexpr int N = Model::ndim;
    Tabs jb;
    jb.jb = stage_table(g, ctb_dynsmem);
    jb.jf = gf;
    if (gf.nint > 0) jb.jf = stage_table(gf, ctb_dynsmem + CTB_TABLE_DOUBLES(g.nint));
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    Model m;
    m.load(params + i * param_stride);
    double X[N], Xp[N], Xpp[N];
#pragma unroll
    for (int d = 0; d < N; d++) X[d] = Xp[d] = Xpp[d] = X0[i * N + d];
    const double scale = norm_n<N>(X) + 1.0;
    const double *Ti = T + i * T_stride;
    const int tb = t_begin ? t_begin[i] : 0;
    const int dir = (t_dir && t_dir[i] < 0) ? -1 : 1;
    // entries outside the walk
    for (int t = lane; t < nT; t += 32) {
        if ((dir > 0) ? (t < tb) : (t > tb)) {
            const int64_t o = i * nT + t;
#pragma unroll
            for (int d = 0; d < N; d++) Xout[o * N + d] = nan("");
            if (Vout) Vout[o] = nan("");
            flag[o] = CTB_MIN_NOT_TRACED;
        }
    }
    double Tp = 0.0, Tpp = 0.0;
    int have = 0, total_iters = 0;
    int t = tb;
    for (; t >= 0 && t < nT; t += dir) {
        const int64_t o = i * nT + t;
        const double Tt = Ti[t];
        m.set_T(Tt);
#pragma unroll
        for (int d = 0; d < N; d++) {
            double xs = Xp[d];
            if (have >= 2 && Tp != Tpp) xs = Xp[d] + (Xp[d] - Xpp[d]) * ((Tt - Tp) / (Tp - Tpp));
            X[d] = xs;
        }
        double fmin = nan("");
        int iters = 0;
        int st = warp_newton_minimize(m, jb, k, lane, X, fmin, iters);
        if (have >= 1) {
            double dd[N];
#pragma unroll
            for (int d = 0; d < N; d++) dd[d] = X[d] - Xp[d];
            if (st != CTB_MIN_OK || norm_n<N>(dd) > k.jump_tol * scale) {
                // Suspicious: no minimum from the extrapolated start, or a long way from the previous one.  Near the end of
                // a phase the minimum moves like sqrt(T_end - T), so a long move on the caller's grid need not be a jump:
                // walk the interval again in CTB_TRACE_SUBSTEPS equal steps from the previous minimum (traceMinimum's
                // adaptive dt, transitionFinder.py:150-215, on a fixed refinement).  Only a sub-step that itself fails or
                // moves by more than the tolerance ends the phase.
                double Xs[N];
#pragma unroll
                for (int d = 0; d < N; d++) Xs[d] = Xp[d];
                st = CTB_MIN_OK;
                for (int sub = 1; sub <= CTB_TRACE_SUBSTEPS && st == CTB_MIN_OK; sub++) {
                    const double Ts = (sub == CTB_TRACE_SUBSTEPS) ? Tt : Tp + (Tt - Tp) * ((double)sub / CTB_TRACE_SUBSTEPS);
                    m.set_T(Ts);
                    double Xq[N];
#pragma unroll
                    for (int d = 0; d < N; d++) Xq[d] = Xs[d];
                    int it2 = 0;
                    st = warp_newton_minimize(m, jb, k, lane, Xs, fmin, it2);
                    iters += it2;
#pragma unroll
                    for (int d = 0; d < N; d++) dd[d] = Xs[d] - Xq[d];
                    if (st == CTB_MIN_OK && norm_n<N>(dd) > k.jump_tol * scale) st = CTB_MIN_JUMPED;
                    if (st != CTB_MIN_OK && sub != CTB_TRACE_SUBSTEPS) {
                        // the phase ended inside the interval: report where the iterate ends up at T[t] itself
                        m.set_T(Tt);
                        const int st2 = warp_newton_minimize(m, jb, k, lane, Xs, fmin, it2);
                        iters += it2;
                        if (st == CTB_MIN_JUMPED && st2 != CTB_MIN_OK) st = st2;
                    }
                }
#pragma unroll
                for (int d = 0; d < N; d++) X[d] = Xs[d];
            }
        }
        total_iters += iters;
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < N; d++) Xout[o * N + d] = X[d];
            if (Vout) Vout[o] = fmin;
            flag[o] = st;
        }
        if (st != CTB_MIN_OK) { t += dir; break; }
#pragma unroll
        for (int d = 0; d < N; d++) { Xpp[d] = Xp[d]; Xp[d] = X[d]; }
        Tpp = Tp; Tp = Tt;
        have++;
    }
    // the rest of the grid after the trace ended
    for (int u = t + dir * lane; u >= 0 && u < nT; u += dir * 32) {
        const int64_t o = i * nT + u;
#pragma unroll
        for (int d = 0; d < N; d++) Xout[o * N + d] = nan("");
        if (Vout) Vout[o] = nan("");
        flag[o] = CTB_MIN_NOT_TRACED;
    }
    if (iters_out && lane == 0) iters_out[i] = total_iters;
}

// host-side launcher (8 warps = 8 instances per 256-thread block; both tables staged as 48-byte records)
template <class Model>
inline int launch_trace_minima(const double *d_params, int64_t param_stride, const JTab &jb, const JTab &jf, const double *d_X0,
                        const double *d_T, int64_t T_stride, const int32_t *d_t_begin, const int32_t *d_t_dir, int64_t n, int nT,
                        const MinKnobs &k, double *d_X, double *d_V, int32_t *d_flag, int32_t *d_iters, cudaStream_t st)
{
    const int64_t blocks = (n + 7) / 8; // 8 warps = 8 instances per 256-thread block
    if (blocks > 0x7fffffff) return CTB_ERR_UNSUPPORTED;
    const size_t sm = (CTB_TABLE_DOUBLES(jb.nint) + (jf.nint > 0 ? CTB_TABLE_DOUBLES(jf.nint) : 0)) * sizeof(double);
    launch_k(trace_minima_kernel<Model>, (unsigned)blocks, 256, sm, st, d_params, param_stride, jb, jf, d_X0, d_T, T_stride,
             d_t_begin, d_t_dir, n, nT, k, d_X, d_V, d_flag, d_iters);
    return (int)cudaGetLastError();
}


} // namespace ctb
