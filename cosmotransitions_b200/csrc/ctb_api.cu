// ctb_api.cu -- streaming kernels and the extern "C" entry points declared in include/ctbounce.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cstdio>
#include "ctb_device.cuh"
#include "ctb_bounce.cuh"
#include "ctb_wall.cuh"
#include "ctb_phases.cuh"

using namespace ctb;

namespace ctb {
int launch_trace_minima_model1(const double *d_params, int64_t param_stride, const JTab &jb, const double *d_X0,
                               const double *d_T, int64_t T_stride, const int32_t *d_t_begin, const int32_t *d_t_dir,
                               int64_t n, int nT, const MinKnobs &k, double *d_X, double *d_V, int32_t *d_flag,
                               int32_t *d_iters, cudaStream_t st);
int launch_trace_minima_thermal1f(const double *d_params, int64_t param_stride, const JTab &jb, const JTab &jf,
                                  const double *d_X0, const double *d_T, int64_t T_stride, const int32_t *d_t_begin,
                                  const int32_t *d_t_dir, int64_t n, int nT, const MinKnobs &k, double *d_X, double *d_V,
                                  int32_t *d_flag, int32_t *d_iters, cudaStream_t st);
}

namespace {

// Jb_spline / Jf_spline (finiteT.py:167-174,197-204).  HBM-streaming: 8 B in + 8 B out per element.
// Persistent blocks stage the 40 KB table in shared memory once; each thread keeps UNROLL
// independent elements in flight.
template <int DER>
__global__ void __launch_bounds__(256) jspline_kernel(JTab g, const double *__restrict__ theta,
                                                      double *__restrict__ out, int64_t n)
{
    extern __shared__ double smem[];
    const JTab t = stage_table(g, smem);
    __syncthreads();
    // vectorised main part: one double2 (16 B) per thread and access, UNROLL independent accesses in flight;
    // requires 16-byte aligned arrays (the scalar tail loop handles the rest, or everything if unaligned)
    constexpr int UNROLL = 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t done = 0;
    if ((((uintptr_t)theta | (uintptr_t)out) & 15) == 0) {
        const double2 *in2 = reinterpret_cast<const double2 *>(theta);
        double2 *out2 = reinterpret_cast<double2 *>(out);
        const int64_t n2 = n >> 1;
        int64_t i = tid;
        for (; i + (UNROLL - 1) * stride < n2; i += UNROLL * stride) {
            double2 x[UNROLL], y[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; u++) x[u] = __ldcs(in2 + i + u * stride);
#pragma unroll
            for (int u = 0; u < UNROLL; u++) { y[u].x = jtab_eval<DER>(t, x[u].x); y[u].y = jtab_eval<DER>(t, x[u].y); }
#pragma unroll
            for (int u = 0; u < UNROLL; u++) __stcs(out2 + i + u * stride, y[u]);
        }
        for (; i < n2; i += stride) {
            const double2 x = in2[i];
            out2[i] = make_double2(jtab_eval<DER>(t, x.x), jtab_eval<DER>(t, x.y));
        }
        done = n2 << 1;
    }
    for (int64_t i = done + tid; i < n; i += stride) out[i] = jtab_eval<DER>(t, theta[i]);
}

// ---- Jb / Jf series variants (finiteT.py:207-296) --------------------------------------------------------
// K_0(z), K_1(z), z > 0, from K_nu(z) = int_0^inf exp(-z cosh t) cosh(nu t) dt with the trapezoidal rule:
// the integrand is analytic in the strip |Im t| < pi/2, so step h = 1/4 gives ~exp(-pi^2/h) = 7e-18.
CTB_DEV void bessel_k01(double z, double &k0, double &k1)
{
    const double h = 0.25, eh = 1.2840254166877414; // exp(1/4)
    double f0 = exp(-z);
    double s0 = 0.5 * f0, s1 = 0.5 * f0;
    double E = 1.0;
    for (int k = 1; k < 4096; k++) {
        E *= eh;
        const double ch = 0.5 * (E + 1.0 / E);
        const double f = exp(-z * ch);
        const double g = f * ch;
        s0 += f;
        s1 += g;
        if (g <= 1e-18 * s1) break;
    }
    k0 = h * s0;
    k1 = h * s1;
}

// nan_to_num for a product that is 0 * inf at x = 0 (finiteT.py:260,265,275)
CTB_DEV double nan0(double y)
{
    if (isnan(y)) return 0.0;
    if (isinf(y)) return y > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return y;
}

// one term K(k, x) of Jb_high / Jf_high: x2K2, dx2K2, d2x2K2, d3x2K2 (finiteT.py:249-275)
CTB_DEV double high_term(int deriv, double k, double x)
{
    double k0 = 0.0, k1 = 0.0;
    const double ax = fabs(x);
    if (deriv == 0) {
        if (x == 0) return -2.0 / (k * k * k * k);
        if (x < 0) return nan(""); // special.kn(2, negative) is NaN
        bessel_k01(k * x, k0, k1);
        const double k2 = k0 + 2.0 * k1 / (k * x);
        return -x * x * k2 / (k * k);
    }
    if (ax > 0) bessel_k01(k * ax, k0, k1);
    else { k0 = INFINITY; k1 = INFINITY; }
    if (deriv == 1) return nan0(x * ax * k1 / k);
    if (deriv == 2) return (ax == 0) ? 1.0 / (k * k) : nan0(ax * (k1 / k - ax * k0));
    return nan0(x * (ax * k * k1 - 3 * k0));
}

struct LowCoef { double c[56]; };

template <int WHICH, int APPROX>
__global__ void __launch_bounds__(256) jseries_kernel(int deriv, int nterms, LowCoef lc, const double *__restrict__ xin,
                                                      double *__restrict__ out, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double x = xin[i];
        double y;
        if (APPROX == CTB_APPROX_HIGH) { // finiteT.py:278-296
            y = 0.0;
            double sign = 1.0;
            for (int k = 1; k <= nterms; k++) {
                y += sign * high_term(deriv, (double)k, x);
                if (WHICH == 1) sign = -sign;
            }
        } else { // finiteT.py:225-244
            const double x2 = x * x;
            double lg = log(x2);
            lg = isnan(lg) ? 0.0 : (isinf(lg) ? (lg > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308) : lg);
            const double *g;
            if (WHICH == 0) {
                y = lc.c[0] + x * x * (lc.c[1] + x * (lc.c[2] + lc.c[3] * x * (lg - lc.c[4])));
                g = lc.c + 5;
            } else {
                y = lc.c[0] + x * x * (lc.c[1] + lc.c[2] * x * x * (lg - lc.c[3]));
                g = lc.c + 4;
            }
            double p = x2 * x2 * x2; // x^(2i+4), i = 1
            for (int k = 1; k <= nterms; k++) {
                y += g[k - 1] * p;
                p *= x2;
            }
        }
        out[i] = y;
    }
}

// ---- Jb / Jf "exact": the defining integrals (finiteT.py:40-147), by quadrature on the device -------------
// The reference hands them to QUADPACK (scipy.integrate.quad, tolerance 1.49e-8).  Here:
//  theta = x^2 >= 0:  y = x sinh u makes the integrand y^2 log(1 -+ exp(-sqrt(y^2 + x^2))) dy an even, entire
//    function of u on the real line with double-exponential decay (nearest singularity at |Im u| = pi/2), so the
//    plain trapezoidal rule in u converges like exp(-pi^2/h): h = 0.2 (narrowed like 1/sqrt(x) for the Gaussian
//    peak at large x) is exact to rounding.  Same substitution for dJ/dx (finiteT.py:104-111).
//  theta < 0 (finiteT.py:54-66,84-96), a = sqrt(-theta), s = sqrt(|y^2 - a^2|):
//    int_0^inf s sqrt(s^2 + a^2) log(1 -+ e^-s) ds      -- double-exponential rule s = exp(pi/2 sinh t)
//    + int_0^a s sqrt(a^2 - s^2) log(2 |sin or cos (s/2)|) ds -- tanh-sinh panels between the logarithmic
//      singularities (s = 2 pi k for Jb, pi (2k+1) for Jf), integrand evaluated from the distance to the nearer
//      panel end so that the singular factors keep full relative accuracy.
CTB_DEV double log_one_minus_exp(double w) // log(1 - e^-w), w > 0
{
    return (w < 0.6931471805599453) ? log(-expm1(-w)) : log1p(-exp(-w));
}

template <int WHICH>
__device__ double jexact_pos(double x, int deriv)
{
    const double pi4 = 97.40909103400243723644;
    if (x < 1e-20) {
        if (deriv) return (WHICH == 0 ? 9.869604401089358 / 6.0 : 9.869604401089358 / 12.0) * x;
        return WHICH == 0 ? -pi4 / 45.0 : -7.0 * pi4 / 360.0;
    }
    const double h = 0.2 / fmax(1.0, 0.5 * sqrt(x));
    double sum = 0.0;
    for (int k = 1; k < 100000; k++) { // (the k = 0 node has y = 0: no contribution)
        const double u = k * h;
        const double eu = exp(u), emu = 1.0 / eu;
        const double y = x * (0.5 * (eu - emu)), w = x * (0.5 * (eu + emu));
        if (w > 745.0) break;
        double g;
        if (deriv == 0) g = (WHICH == 0) ? y * y * w * log_one_minus_exp(w) : -(y * y * w) * log1p(exp(-w));
        else {
            const double e = exp(-w);
            g = (WHICH == 0) ? (y * y * x) * e / (-expm1(-w)) : (y * y * x) * e / (1.0 + e);
        }
        sum += g;
    }
    return sum * h;
}

template <int WHICH>
__device__ double jexact_neg(double theta)
{
    const double a = sqrt(-theta), hd = 1.0 / 16.0, hpi = 1.5707963267948966, pi = 3.141592653589793;
    double p1 = 0.0;
    for (int k = -120; k <= 120; k++) { // s in (e^-40, e^7)
        const double t = k * hd, q = hpi * sinh(t);
        if (q > 7.0 || q < -40.0) continue;
        const double s = exp(q), wgt = s * hpi * cosh(t);
        const double f = (WHICH == 0) ? log_one_minus_exp(s) : -log1p(exp(-s));
        p1 += (s * sqrt(s * s + a * a) * f) * wgt;
    }
    double p2 = 0.0;
    // panel ends: Jb 0, 2 pi, 4 pi, ...; Jf 0, pi, 3 pi, ...   (clipped at a)
    for (int i = 0; i < 100000; i++) {
        const double p = (WHICH == 0) ? 2.0 * pi * i : (i == 0 ? 0.0 : pi * (2 * i - 1));
        if (p >= a) break;
        const double qn = (WHICH == 0) ? 2.0 * pi * (i + 1) : pi * (2 * i + 1);
        const bool right_sing = qn <= a, left_sing = !(WHICH == 1 && i == 0);
        const double q_ = right_sing ? qn : a, L = q_ - p;
        double acc = 0.0;
        for (int k = -80; k <= 80; k++) {
            const double t = k * hd, al = hpi * sinh(t);
            if (fabs(al) > 36.0) continue;
            const double e2 = exp(2.0 * al);
            const double dl = L * e2 / (1.0 + e2), dr = L / (1.0 + e2); // distances from the panel ends
            const double ch = cosh(al);
            const double wgt = L * hpi * cosh(t) / (2.0 * ch * ch);
            const bool left = dl < dr;
            const double s = left ? p + dl : q_ - dr;
            const double root = sqrt(((a - q_) + dr) * (a + s));
            double tr; // |sin(s/2)| (Jb) or |cos(s/2)| (Jf)
            if (left && left_sing) tr = fabs(sin(0.5 * dl));
            else if (!left && right_sing) tr = fabs(sin(0.5 * dr));
            else tr = (WHICH == 0) ? fabs(sin(0.5 * s)) : fabs(cos(0.5 * s));
            const double f = root * s * log(2.0 * tr);
            acc += ((WHICH == 0) ? f : -f) * wgt;
        }
        p2 += acc;
    }
    return (p1 + p2) * hd;
}

// mode 0: J(x), x = m/T real (J is even in x); 1: dJ/dx; 2: J(theta), theta = x^2 of either sign
template <int WHICH>
__global__ void __launch_bounds__(128) jexact_kernel(int mode, const double *__restrict__ xin, double *__restrict__ out,
                                                     int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = xin[i];
    double y;
    if (isnan(x)) y = x;
    else if (mode == 0) y = jexact_pos<WHICH>(fabs(x), 0);
    else if (mode == 1) y = copysign(jexact_pos<WHICH>(fabs(x), 1), x);
    else y = (x >= 0) ? jexact_pos<WHICH>(sqrt(x), 0) : jexact_neg<WHICH>(x);
    out[i] = y;
}

// exactSolution for rows of (r, phi0, dV, d2V): parity aid for the Bessel start conditions (T1D:294-373)
template <int ALPHA>
__global__ void __launch_bounds__(128) exact_solution_kernel(double alpha, const double *__restrict__ in,
                                                             double *__restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double2 e = exact_solution<ALPHA>(in[4 * i], exact_coef(in[4 * i + 1], in[4 * i + 2], in[4 * i + 3], 0.5 * (alpha - 1.0)));
    out[2 * i] = e.x;
    out[2 * i + 1] = e.y;
}

struct Model8 { double l1, l2, mu2, Y1, Y2, nb, rs, v2; };

CTB_DEV double model1_v1t(const Model1Field &f, double nb, const JTab &jb, double inv_T2, double T4c)
{
    double yt = jtab_eval<0>(jb, f.m0 * inv_T2) + jtab_eval<0>(jb, f.m1 * inv_T2) + nb * jtab_eval<0>(jb, f.mb * inv_T2);
    return yt * T4c;
}
CTB_DEV double model1_thermal(const Model1Field &f, double nb, const JTab &jb, double inv_T2, double T4c)
{
    return f.V01 + model1_v1t(f, nb, jb, inv_T2, T4c);
}

// Vtot(X_i, T_i) pointwise (generic_potential.py:312-337 for model1); `parts` selects V0 / V1 / V1T
__global__ void __launch_bounds__(256) vtot_model1_kernel(Model8 m, JTab g, const double *__restrict__ X,
                                                          const double *__restrict__ T, double *__restrict__ out,
                                                          int64_t n, int parts)
{
    extern __shared__ double smem[];
    const JTab jb = stage_table(g, smem);
    __syncthreads();
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double2 x = reinterpret_cast<const double2 *>(X)[i];
        const double Ti = T[i];
        const Model1Field f = model1_field(m.l1, m.l2, m.mu2, m.Y1, m.Y2, m.nb, 1.0 / m.rs, m.v2, x.x, x.y);
        const double T2 = Ti * Ti + 1e-100;
        const double v1t = (parts & CTB_V_THERMAL)
                               ? model1_v1t(f, m.nb, jb, 1.0 / T2, (Ti * Ti * Ti * Ti) / (2 * CTB_PI * CTB_PI)) : 0.0;
        double v;
        if (parts == CTB_V_TOTAL) v = f.V01 + v1t;
        else {
            v = 0.0;
            if (parts & CTB_V_TREE) v += f.V0;
            if (parts & CTB_V_ONELOOP) v += f.V1;
            v += v1t;
        }
        out[i] = v;
    }
}

// Vtot on the outer-product grid X_i = x0 + s_i nhat, T_j  (out[i * nT + j]).
// Persistent blocks (4 per SM): each stages the Jb table in shared memory once and owns a contiguous
// range of output points, so stores are coalesced and the 148 SMs are evenly loaded whatever the grid
// shape.  The field-dependent part (three masses, V0 + V1: three logs and a sqrt) is computed once per
// field point touched by the block and shared through shared memory.
#define VG_THREADS 256
#define VG_ROWCAP 64
__global__ void __launch_bounds__(VG_THREADS) vtot_model1_grid_kernel(Model8 m, double x0, double x1, double n0,
                                                                      double n1, JTab g, const double *__restrict__ s,
                                                                      int64_t ns, const double *__restrict__ T,
                                                                      int64_t nT, double *__restrict__ out,
                                                                      int64_t chunk)
{
    extern __shared__ double smem[];
    Model1Field *rowc = reinterpret_cast<Model1Field *>(smem);
    const JTab jb = stage_table(g, smem + VG_ROWCAP * 6);
    const int64_t total = ns * nT;
    const int64_t p0 = (int64_t)blockIdx.x * chunk;
    const int64_t p1 = (p0 + chunk < total) ? p0 + chunk : total;
    if (p0 < p1) {
        const int64_t i_first = p0 / nT, i_last = (p1 - 1) / nT;
        for (int64_t r = threadIdx.x; r <= i_last - i_first; r += VG_THREADS) {
            const double si = s[i_first + r];
            rowc[r] = model1_field(m.l1, m.l2, m.mu2, m.Y1, m.Y2, m.nb, 1.0 / m.rs, m.v2, x0 + si * n0, x1 + si * n1);
        }
        __syncthreads();
        int64_t p = p0 + threadIdx.x;
        int64_t i = p / nT, j = p - i * nT;
        for (; p < p1; p += VG_THREADS) {
            const double Tj = __ldg(T + j);
            const double T2 = Tj * Tj;
            const double inv_T2 = 1.0 / (T2 + 1e-100);
            const double T4c = (T2 * T2) / (2 * CTB_PI * CTB_PI);
            __stcs(out + p, model1_thermal(rowc[i - i_first], m.nb, jb, inv_T2, T4c));
            j += VG_THREADS;
            while (j >= nT) { j -= nT; i++; }
        }
    } else {
        __syncthreads();
    }
}

// The same grid, tiled: a block owns VGT_THREADS consecutive temperatures and a range of field points.  Each THREAD
// keeps one temperature (1 / T^2 and T^4 / 2 pi^2 are formed once, in registers) and walks down the block's rows; the
// field-dependent part of a row (three masses, V0 + V1: three logs and a square root) is computed once per block and
// row and broadcast from shared memory.
//
// INTERVAL TRACKING.  The first tiled version looked every theta = m^2 / T^2 up from scratch (FP32 asinh guess, 48-byte
// record from shared memory): ncu showed it bound by the shared-memory data pipe -- 12 wavefronts per warp look-up is the
// floor for delivering 48 B to each of 32 lanes, 5.7e6 wavefronts = 57 % of the run time -- with the XU pipe (sqrt,
// log, conversions of the guess) second.  But down a column theta moves by ~1e-3 of itself per row while a knot
// interval is ~1e-2 wide, so a thread KEEPS the record of each of its three look-ups in registers (b0, b1, c0..c3)
// and re-uses it while theta stays inside [b0, b1): one subtraction and two compares instead of the guess and the
// loads.  When theta has left the interval the neighbouring record is fetched (predicated loads: only the lanes that
// moved cost shared-memory wavefronts); only if that is not enough either (first row of a block, a jump) the closed-form
// guess runs, under a warp-uniform branch.
#define VGT_THREADS 256
#define VGT_ROWS 32
#ifndef VGT_MIN_BLOCKS
#define VGT_MIN_BLOCKS 2
#endif
struct TrackRec { double b0, b1, c0, c1, c2, c3; int j; };

// Staging for the tracked look-up: nint + 1 records.  The last real record ends just above xmax (so that theta = xmax
// still evaluates the spline, finiteT.py:173) and is followed by an all-zero record [nextafter(xmax), inf) -- "theta >
// xmax -> 0" then needs no test of its own.  `smem` 16-byte aligned, (nint + 1) * 6 doubles.
// Split in two so that the global loads are in flight while the block computes its first row cache: stage_tracked_load
// reads the records of this thread (4 per thread at 256 threads) into registers, stage_tracked_store writes them out.
// d_coef must be 16-byte aligned (ctbounce.h asks for 32).
#define VGT_STAGE_PER_THREAD 4
struct StageRegs { double b0[VGT_STAGE_PER_THREAD], b1[VGT_STAGE_PER_THREAD]; double2 c01[VGT_STAGE_PER_THREAD], c23[VGT_STAGE_PER_THREAD]; };
CTB_DEV void stage_tracked_load(const JTab &g, StageRegs &q)
{
    const double2 *c2 = reinterpret_cast<const double2 *>(g.coef);
#pragma unroll
    for (int u = 0; u < VGT_STAGE_PER_THREAD; u++) {
        const int j = threadIdx.x + u * VGT_THREADS;
        if (j < g.nint) {
            q.b0[u] = __ldg(g.brk + j); q.b1[u] = __ldg(g.brk + j + 1);
            q.c01[u] = __ldg(c2 + 2 * j); q.c23[u] = __ldg(c2 + 2 * j + 1);
        }
    }
}
CTB_DEV JTab stage_tracked_store(const JTab &g, const StageRegs &q, double *smem)
{
    const double top = __longlong_as_double(__double_as_longlong(g.xmax) + (g.xmax > 0.0 ? 1 : -1)); // nextafter(xmax, inf), xmax != 0
    double2 *rec = reinterpret_cast<double2 *>(smem);
#pragma unroll
    for (int u = 0; u < VGT_STAGE_PER_THREAD; u++) {
        const int j = threadIdx.x + u * VGT_THREADS;
        if (j < g.nint) {
            rec[3 * j] = make_double2(q.b0[u], q.c01[u].x);
            rec[3 * j + 1] = make_double2(q.c01[u].y, q.c23[u].x);
            rec[3 * j + 2] = make_double2(q.c23[u].y, (j == g.nint - 1) ? top : q.b1[u]);
        } else if (j == g.nint) {
            rec[3 * j] = make_double2(top, 0.0);
            rec[3 * j + 1] = make_double2(0.0, 0.0);
            rec[3 * j + 2] = make_double2(0.0, __longlong_as_double(0x7ff0000000000000LL));
        }
    }
    JTab t = g;
    t.mode = JT_RECORDS;
    t.i_rec = (int)(smem - ctb_dynsmem);
    return t;
}

CTB_DEV void track_load(TrackRec &k, int i_rec, int j)
{
    const double2 *r = reinterpret_cast<const double2 *>(ctb_dynsmem + i_rec) + 3 * j;
    const double2 q0 = r[0], q1 = r[1], q2 = r[2];
    k.b0 = q0.x; k.c0 = q0.y; k.c1 = q1.x; k.c2 = q1.y; k.c3 = q2.x; k.b1 = q2.y; k.j = j;
}

// value of the staged piecewise cubic at theta through the tracked record (same arithmetic as jtab_eval<0>)
CTB_DEV double track_eval(TrackRec &k, int i_rec, int nint, double xmin, float loc_scale, float loc_inv_du, float loc_c,
                          double theta)
{
    const double x = (theta < xmin) ? xmin : theta;
    if (!(x >= k.b0 && x < k.b1)) {
        track_load(k, i_rec, min(max(k.j + ((x >= k.b1) ? 1 : -1), 0), nint)); // the neighbouring interval
        if (!(x >= k.b0 && x < k.b1)) { // far away (first row, a jump, NaN): locate from scratch (one-sided guess)
            const int j = jtab_guess_f(x, loc_scale, loc_inv_du, loc_c, nint - 1);
            track_load(k, i_rec, j);
            if (x >= k.b1) track_load(k, i_rec, j + 1);
        }
    }
    return jtab_poly<0>(__dsub_rn(x, k.b0), k.c0, k.c1, k.c2, k.c3);
}

__global__ void __launch_bounds__(VGT_THREADS, VGT_MIN_BLOCKS) vtot_model1_grid_tiled_kernel(Model8 m, double x0, double x1, double n0,
                                                                               double n1, JTab g, const double *__restrict__ s,
                                                                               int64_t ns, const double *__restrict__ T,
                                                                               int64_t nT, double *__restrict__ out,
                                                                               int rows_per_block, int col_tiles)
{
    extern __shared__ double smem[];
    double4 *rowc = reinterpret_cast<double4 *>(smem); // {m0, m1, mb, V0 + V1} per cached row
    StageRegs sq;
    stage_tracked_load(g, sq);
    JTab jb = g;
    // consecutive block ids go down the rows of one column tile: the block scheduler hands an SM ids k, k + #SM, ... ,
    // i.e. (when the grid is one wave) one block of EACH column tile -- the tiles differ in cost (share of far look-ups)
    const int row_chunks = gridDim.x / col_tiles;
    const int ct = blockIdx.x / row_chunks, rb = blockIdx.x - ct * row_chunks;
    const int64_t j = (int64_t)ct * VGT_THREADS + threadIdx.x;
    const int64_t i0 = (int64_t)rb * rows_per_block;
    const int64_t i1 = (i0 + rows_per_block < ns) ? i0 + rows_per_block : ns;
    const bool on = j < nT;
    // columns past the end of the grid follow the last one (nothing is stored)
    const double Tj = T[on ? j : nT - 1], T2 = Tj * Tj;
    const double inv_T2 = 1.0 / (T2 + 1e-100);
    const double T4c = (T2 * T2) / (2 * CTB_PI * CTB_PI);
    const double inv_rs = 1.0 / m.rs;
    TrackRec k0, k1, kb;
    k0.b0 = k1.b0 = kb.b0 = 1.0; k0.b1 = k1.b1 = kb.b1 = 0.0; // empty interval: the first look-up locates from scratch
    k0.c0 = k0.c1 = k0.c2 = k0.c3 = k1.c0 = k1.c1 = k1.c2 = k1.c3 = kb.c0 = kb.c1 = kb.c2 = kb.c3 = 0.0;
    k0.j = k1.j = kb.j = 0;
    for (int64_t ib = i0; ib < i1; ib += VGT_ROWS) {
        const int nrows = (int)((i1 - ib < VGT_ROWS) ? i1 - ib : VGT_ROWS);
        if (ib != i0) __syncthreads(); // previous pass has finished reading the row cache
        if (threadIdx.x < nrows) {
            const double si = s[ib + threadIdx.x];
            const Model1Field f = model1_field(m.l1, m.l2, m.mu2, m.Y1, m.Y2, m.nb, inv_rs, m.v2, x0 + si * n0, x1 + si * n1);
            rowc[threadIdx.x] = make_double4(f.m0, f.m1, f.mb, f.V01);
        }
        if (ib == i0) jb = stage_tracked_store(g, sq, smem + VGT_ROWS * 4);
        __syncthreads();
        double *o = out + ib * nT + (on ? j : 0);
#pragma unroll 1
        for (int r = 0; r < nrows; r++) {
            const double4 q = rowc[r];
            const double yt = track_eval(k0, jb.i_rec, jb.nint, jb.xmin, jb.loc_scale, jb.loc_inv_du, jb.loc_c, q.x * inv_T2)
                              + track_eval(k1, jb.i_rec, jb.nint, jb.xmin, jb.loc_scale, jb.loc_inv_du, jb.loc_c, q.y * inv_T2)
                              + m.nb * track_eval(kb, jb.i_rec, jb.nint, jb.xmin, jb.loc_scale, jb.loc_inv_du, jb.loc_c, q.z * inv_T2);
            if (on) __stcs(o + (int64_t)r * nT, q.w + yt * T4c);
        }
    }
}

// generic_potential.Vtot / V1T_from_X (GP:258-337) for the general one-field model (CTB_POT_THERMAL1F layout, one
// model, per-point temperature): X_i, T_i -> out_i.  `rad` = (num_boson_dof - sum n_b) pi^4/45 + (num_fermion_dof -
// sum n_f) 7 pi^4/360, the field-independent radiation term of V1T (GP:289-295), 0 without it.
struct Thermal1FModel { double v[CTB_NPARAM_THERMAL1F]; };
__global__ void __launch_bounds__(256) vtot_thermal1f_kernel(Thermal1FModel m, double rad, JTab gb, JTab gf,
                                                             const double *__restrict__ phi, const double *__restrict__ T,
                                                             double *__restrict__ out, int64_t n, int parts)
{
    extern __shared__ double smem[];
    __shared__ double row[CTB_NPARAM_THERMAL1F];
    const JTab jb = stage_table_compact(gb, smem);
    const JTab jf = stage_table_compact(gf, smem + ((CTB_TABLE_DOUBLES_COMPACT(gb.nint) + 1) & ~(size_t)1)); // 16-byte aligned
    if (threadIdx.x < CTB_NPARAM_THERMAL1F) row[threadIdx.x] = m.v[threadIdx.x];
    __syncthreads();
    const double inv_rs = 1.0 / row[5];
    const int nb = min(max((int)row[7], 0), CTB_THERMAL1F_MAX_BOSONS), nf = min(max((int)row[8], 0), CTB_THERMAL1F_MAX_FERMIONS);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const double Ti = T[i], T2 = Ti * Ti;
        const PotThermal1F::Parts o = PotThermal1F::eval(row, nb, nf, inv_rs, T2, 1.0 / (T2 + 1e-100), 0.0, jb, jf, phi[i]);
        double v = 0.0;
        if (parts & CTB_V_TREE) v += o.v0;
        if (parts & CTB_V_ONELOOP) v += o.v1;
        if (parts & CTB_V_THERMAL) v += (o.v1t - rad) * ((T2 * T2) / (2 * CTB_PI * CTB_PI));
        out[i] = v;
    }
}

template <class Pot>
__global__ void potential_kernel(const double *params, JTab jb, JTab jf, const double *__restrict__ phi,
                                 double *__restrict__ out, int64_t n)
{
    extern __shared__ double smem[];
    if (!Pot::analytic) {
        jb = stage_table_compact(jb, smem);
        __syncthreads();
    }
    Pot pot;
    pot.load(params, &jb, &jf);
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = pot.V(phi[i]);
}

// V_i(phi_i) for n parameter rows / fminbound of V_i on [lo_i, hi_i]: one thread per row
template <class Pot, bool MINIMIZE>
__global__ void __launch_bounds__(128) rows_kernel(const double *__restrict__ params, JTab jb, JTab jf,
                                                   const double *__restrict__ a, const double *__restrict__ b,
                                                   double xtol_rel, double *__restrict__ out0,
                                                   double *__restrict__ out1, int64_t n)
{
    extern __shared__ double smem[];
    if (!Pot::analytic) {
        jb = stage_table_compact(jb, smem);
        __syncthreads();
    }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Pot pot;
    pot.load(params + i * Pot::nparam, &jb, &jf);
    if (!MINIMIZE) {
        out0[i] = pot.V(a[i]);
    } else {
        const double lo = a[i], hi = b[i];
        const double x = fminbound([&](double p) { return pot.V(p); }, lo, hi, xtol_rel * (hi - lo));
        out0[i] = x;
        if (out1) out1[i] = pot.V(x);
    }
}

__global__ void fp64_peak_kernel(int64_t iters, double *sink)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    double b0 = a0 + 8, b1 = a0 + 9, b2 = a0 + 10, b3 = a0 + 11, b4 = a0 + 12, b5 = a0 + 13, b6 = a0 + 14, b7 = a0 + 15;
    const double m = 0.999999, c = 1e-7;
    for (int64_t k = 0; k < iters; k++) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        b0 = fma(b0, m, c); b1 = fma(b1, m, c); b2 = fma(b2, m, c); b3 = fma(b3, m, c);
        b4 = fma(b4, m, c); b5 = fma(b5, m, c); b6 = fma(b6, m, c); b7 = fma(b7, m, c);
    }
    double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7)) + ((b0 + b1) + (b2 + b3)) + ((b4 + b5) + (b6 + b7));
    if (r == 12345.678) sink[0] = r; // never true; keeps the loop alive
}

size_t table_bytes(const ctb_jtable *t) { return t ? CTB_TABLE_DOUBLES(t->nint) * sizeof(double) : 0; }
size_t table_bytes_compact(const ctb_jtable *t) { return t ? CTB_TABLE_DOUBLES_COMPACT(t->nint) * sizeof(double) : 0; }

// every kernel that stages a table keeps it in at most 48 KB of dynamic shared memory (no opt-in needed):
// 48-byte records -> nint <= 1000 incl. the 8.5 KB of static staging in find_action_kernel (the shipped tables have 997 intervals)
#define CTB_JTAB_MAX_NINT 1000
bool jtab_ok(const ctb_jtable *t) { return t && t->d_brk && t->d_coef && t->nint >= 1 && t->nint <= CTB_JTAB_MAX_NINT; }

template <bool MINIMIZE>
int launch_rows(int pot_kind, const double *d_params, const ctb_jtable *jb, const ctb_jtable *jf, const double *a,
                const double *b, double xtol_rel, double *o0, double *o1, int64_t n, cudaStream_t st)
{
    unsigned g = (unsigned)((n + 127) / 128);
    switch (pot_kind) {
    case CTB_POT_POLY4:
        rows_kernel<PotPoly4, MINIMIZE><<<g, 128, 0, st>>>(d_params, to_jtab(jb), to_jtab(jf), a, b, xtol_rel, o0, o1, n);
        break;
    case CTB_POT_SPLINE:
        rows_kernel<PotSpline, MINIMIZE><<<g, 128, 0, st>>>(d_params, to_jtab(jb), to_jtab(jf), a, b, xtol_rel, o0, o1, n);
        break;
    case CTB_POT_THERMAL1F:
        if (!(jtab_ok(jb) && jtab_ok(jf))) return CTB_ERR_BAD_ARG;
        launch_k(rows_kernel<PotThermal1F, MINIMIZE>, g, 128, pot_smem_bytes<PotThermal1F>(jb->nint), st, d_params, to_jtab(jb), to_jtab(jf), a, b,
                                                                            xtol_rel, o0, o1, n);
        break;
    case CTB_POT_MODEL1_LINE:
        if (!jtab_ok(jb)) return CTB_ERR_BAD_ARG;
        launch_k(rows_kernel<PotModel1Line, MINIMIZE>, g, 128, pot_smem_bytes<PotModel1Line>(jb->nint), st, d_params, to_jtab(jb), to_jtab(jf), a, b,
                                                                             xtol_rel, o0, o1, n);
        break;
    case CTB_POT_MODEL1F:
        if (!jtab_ok(jb)) return CTB_ERR_BAD_ARG;
        launch_k(rows_kernel<PotModel1F, MINIMIZE>, g, 128, pot_smem_bytes<PotModel1F>(jb->nint), st, d_params, to_jtab(jb), to_jtab(jf), a, b,
                                                                          xtol_rel, o0, o1, n);
        break;
    default: return CTB_ERR_UNSUPPORTED;
    }
    return (int)cudaGetLastError();
}


// findAction for arbitrary Profile1D arrays (tunneling1D.py:763-791): one WARP per profile.
// scipy.integrate.simpson on the non-uniform R is linear in the integrand, S = sum_j w_j y_j: lane-strided
// points, each lane forms its point's weight from the five neighbouring abscissae (coalesced reads) and
// evaluates the integrand once; warp-shuffle reduction; lane 0 adds the bulk term of the bubble interior.
// one Simpson pair of scipy.integrate.simpson on unequal spacing (_quadrature.py:_basic_simpson)
CTB_DEV double simpson_pair_sum(double x0, double x1, double x2, double y0, double y1, double y2)
{
    const double h0 = x1 - x0, h1 = x2 - x1, hsum = h0 + h1, hprod = h0 * h1;
    if (hprod > 0 && hprod < 1e300) { // regular spacing: h0/h1, h1/h0 and hsum/hprod from one Newton reciprocal
        const double rp = rcp_fast(hprod);
        return (hsum * (1.0 / 6.0)) * (y0 * (2.0 - h1 * h1 * rp) + y1 * (hsum * (hsum * rp)) + y2 * (2.0 - h0 * h0 * rp));
    }
    const double h0divh1 = (h1 != 0) ? h0 / h1 : 0.0; // scipy's guarded divisions (repeated abscissae)
    const double inv = (h0divh1 != 0) ? 1.0 / h0divh1 : 0.0;
    const double q = (hprod != 0) ? hsum / hprod : 0.0;
    return (hsum / 6.0) * (y0 * (2.0 - inv) + y1 * (hsum * q) + y2 * (2.0 - h0divh1));
}

#ifndef FA_MIN_BLOCKS
#define FA_MIN_BLOCKS 6
#endif
#define FA_U 4                     // points per lane and chunk
#define FA_PTS (FA_U * 32)         // a chunk = 64 Simpson pairs = 129 points (the last one is shared with the next chunk)
#define FA_ROW (FA_PTS + 8)
// findAction for given profiles (tunneling1D.py:763-791): one WARP per profile, HBM-streaming (24 B per point in).
// Per chunk of 128 points: phase A -- every lane turns its four points (r, phi, phi') into the integrand and stages
// (r, y) in shared memory; the loads of the NEXT chunk are issued before the arithmetic of this one (double
// buffering in registers); phase B -- every lane sums two Simpson pairs from the staged values (one reciprocal per
// pair, no divergence between odd and even points).
template <class Pot>
__global__ void __launch_bounds__(128, FA_MIN_BLOCKS) find_action_kernel(const double *__restrict__ params, JTab jb, JTab jf,
                                                          const double *__restrict__ phi_meta,
                                                          const double *__restrict__ R, const double *__restrict__ Phi,
                                                          const double *__restrict__ dPhi, int64_t stride,
                                                          const int32_t *__restrict__ npts, int64_t n, double alpha,
                                                          double afac, double vfac, double *__restrict__ S)
{
    extern __shared__ double smem[];
    __shared__ double fa_r[4 * FA_ROW], fa_y[4 * FA_ROW];
    if (!Pot::analytic) {
        jb = stage_table_compact(jb, smem);
        __syncthreads();
    }
    const int lane = threadIdx.x & 31;
    const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return; // whole warp
    Pot pot;
    pot.load(params + i * Pot::nparam, &jb, &jf);
    const double *r = R + i * stride, *ph = Phi + i * stride, *dp = dPhi + i * stride;
    const int N = npts[i];
    if (N < 1 || N > stride) { if (lane == 0) S[i] = nan(""); return; }
    const double Vm = pot.V_quad(phi_meta[i]);
    const int nb = (N & 1) ? N : N - 1; // composite rule on the first nb points
    const int ialpha = (alpha == (double)(int)alpha) ? (int)alpha : -1;
    auto integrand = [&](double rj, double phj, double dpj) {
        const double ra = (ialpha == 2) ? rj * rj : (ialpha == 3) ? rj * rj * rj : (ialpha == 0) ? 1.0 : pow(rj, alpha);
        return (0.5 * (dpj * dpj) + pot.V_quad(phj) - Vm) * ra;
    };
    double *sr = fa_r + (threadIdx.x >> 5) * FA_ROW, *sy = fa_y + (threadIdx.x >> 5) * FA_ROW;
    double rn[FA_U + 1], pn[FA_U + 1], dn[FA_U + 1]; // slot FA_U: the chunk's 129th point (lane 0 only)
    auto fetch = [&](int base) {
#pragma unroll
        for (int u = 0; u <= FA_U; u++) {
            const int j = base + 32 * u + lane;
            const bool on = j < N && (u < FA_U || lane == 0);
            rn[u] = on ? __ldcs(r + j) : 0.0; pn[u] = on ? __ldcs(ph + j) : 0.0; dn[u] = on ? __ldcs(dp + j) : 0.0;
        }
    };
    double acc = 0.0;
    fetch(0);
    for (int base = 0; base < nb - 1; base += FA_PTS) {
        __syncwarp();
#pragma unroll
        for (int u = 0; u <= FA_U; u++) {
            if (u < FA_U || lane == 0) { sr[32 * u + lane] = rn[u]; sy[32 * u + lane] = integrand(rn[u], pn[u], dn[u]); }
        }
        if (base + FA_PTS < nb - 1) fetch(base + FA_PTS); // in flight while the pairs of this chunk are summed
        __syncwarp();
#pragma unroll
        for (int v = 0; v < FA_U / 2; v++) {
            const int a = 2 * (lane + 32 * v);            // local index of the pair's first point
            if (base + a + 2 <= nb - 1)
                acc += simpson_pair_sum(sr[a], sr[a + 1], sr[a + 2], sy[a], sy[a + 1], sy[a + 2]);
        }
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        if (N == 2) acc = 0.5 * (r[1] - r[0]) * (integrand(r[1], ph[1], dp[1]) + integrand(r[0], ph[0], dp[0]));
        else if (!(N & 1)) { // even N: Cartwright's correction for the last interval (scipy.integrate.simpson)
            const double h0 = r[N - 2] - r[N - 3], h1 = r[N - 1] - r[N - 2];
            double num, den, al, be, et;
            num = 2 * h1 * h1 + 3 * h0 * h1; den = 6 * (h1 + h0); al = (den != 0) ? num / den : 0.0;
            num = h1 * h1 + 3.0 * h0 * h1; den = 6 * h0; be = (den != 0) ? num / den : 0.0;
            num = 1 * h1 * h1 * h1; den = 6 * h0 * (h0 + h1); et = (den != 0) ? num / den : 0.0;
            acc += al * integrand(r[N - 1], ph[N - 1], dp[N - 1]) + be * integrand(r[N - 2], ph[N - 2], dp[N - 2])
                   - et * integrand(r[N - 3], ph[N - 3], dp[N - 3]);
        }
        const double r0 = r[0];
        const double rd = (ialpha == 2) ? r0 * r0 * r0 : (ialpha == 3) ? (r0 * r0) * (r0 * r0) : pow(r0, alpha + 1.0);
        S[i] = acc * afac + (rd * vfac) * (pot.V_quad(ph[0]) - Vm);
    }
}

template <class Pot>
int launch_find_action(const double *params, const ctb_jtable *jb, const ctb_jtable *jf, const double *phi_meta,
                       const double *R, const double *Phi, const double *dPhi, int64_t stride, const int32_t *npts,
                       int64_t n, double alpha, double *S, cudaStream_t st)
{
    const double d = alpha + 1.0; // T1D:782-790
    const double afac = 2.0 * pow(CTB_PI, 0.5 * d) / tgamma(0.5 * d), vfac = pow(CTB_PI, 0.5 * d) / tgamma(0.5 * d + 1.0);
    const size_t sm = Pot::analytic ? 0 : pot_smem_bytes<Pot>(jb->nint);
    launch_k(find_action_kernel<Pot>, (unsigned)((n + 3) / 4), 128, sm, st, params, to_jtab(jb), to_jtab(jf), phi_meta, R, Phi,
             dPhi, stride, npts, n, alpha, afac, vfac, S);
    return (int)cudaGetLastError();
}

// number of SMs of the current device (148 on B200), queried once per device
int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int grid_for(int64_t n, int block)
{
    int64_t g = (n + block - 1) / block;
    const int64_t cap = (int64_t)sm_count() * 5; // persistent grid-stride blocks (5 x 40 KB table copies per SM)
    return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

} // namespace

extern "C" {

CTB_API int ctb_abi_version(void) { return CTB_ABI_VERSION; }

CTB_API const char *ctb_strerror(int code)
{
    switch (code) {
    case CTB_OK: return "ok";
    case CTB_ERR_BAD_ARG: return "bad argument";
    case CTB_ERR_UNSUPPORTED: return "unsupported option (alpha must be 1..4, 2 <= npoints <= 4096, known pot_kind)";
    case CTB_ERR_NO_DEVICE: return "no usable CUDA device (needs sm_100)";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

CTB_API const char *ctb_status_string(int status)
{
    static const char *names[] = {"converged", "xtol reached", "stable, not metastable", "no barrier", "r > rmax",
                                  "dr < drmin", "Stepsize rounds down to zero.",
                                  "non-finite value in initialConditions root find", "brentq: same sign at both ends",
                                  "non-finite initial conditions on the first shot", "not run",
                                  "more than CTB_MAX_STEPS integration steps"};
    if (status < 0 || status > 11) return "unknown status";
    return names[status];
}

CTB_API int ctb_device_ok(void)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return CTB_ERR_NO_DEVICE;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return CTB_ERR_NO_DEVICE;
    return (p.major == 10) ? CTB_OK : CTB_ERR_NO_DEVICE;
}

CTB_API int ctb_sm_count(void)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    return sm_count();
}

CTB_API int ctb_bounce_batch(int pot_kind, const ctb_bounce_in *in, const ctb_opts *o, const ctb_jtable *jb,
                             const ctb_jtable *jf, const ctb_bounce_out *out, void *stream)
{
    BounceArgs a;
    BounceLaunch L;
    int rc = ctb::fill_bounce_args(a, in, jb, jf, out);
    if (rc) return rc;
    if (a.n == 0) return o ? CTB_OK : CTB_ERR_BAD_ARG;
    rc = ctb::fill_bounce_knobs(a, o, out, pot_kind == CTB_POT_POLY4, L);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int block = L.block, stage_cap = L.stage_cap;
    const bool warp = L.warp;
    if ((pot_kind == CTB_POT_MODEL1_LINE || pot_kind == CTB_POT_MODEL1F) && !jtab_ok(jb)) return CTB_ERR_BAD_ARG;
    if (pot_kind == CTB_POT_THERMAL1F && !(jtab_ok(jb) && jtab_ok(jf))) return CTB_ERR_BAD_ARG;
    switch (pot_kind) {
    case CTB_POT_POLY4:
        return warp ? ctb::launch_bounce_warp_poly4(a, o->alpha, stage_cap, st)
                    : ctb::launch_bounce_poly4(a, o->alpha, block, st);
    case CTB_POT_SPLINE:
        return warp ? ctb::launch_bounce_warp_spline(a, o->alpha, stage_cap, st)
                    : ctb::launch_bounce_spline(a, o->alpha, block, st);
    case CTB_POT_THERMAL1F:
        return warp ? ctb::launch_bounce_warp_thermal1f(a, o->alpha, stage_cap, st)
                    : ctb::launch_bounce_thermal1f(a, o->alpha, block, st);
    case CTB_POT_MODEL1_LINE:
        return warp ? ctb::launch_bounce_warp_model1_line(a, o->alpha, stage_cap, st)
                    : ctb::launch_bounce_model1_line(a, o->alpha, block, st);
    case CTB_POT_MODEL1F:
        return warp ? ctb::launch_bounce_warp_model1f(a, o->alpha, stage_cap, st)
                    : ctb::launch_bounce_model1f(a, o->alpha, block, st);
    }
    return CTB_ERR_UNSUPPORTED;
}

CTB_API int ctb_bounce_setup(int pot_kind, const ctb_bounce_in *in, const ctb_jtable *jb, const ctb_jtable *jf,
                             const ctb_bounce_out *out, float *d_work_key, void *stream)
{
    BounceArgs a;
    int rc = ctb::fill_bounce_args(a, in, jb, jf, out);
    if (rc) return rc;
    if (a.n == 0) return CTB_OK;
    a.work_key = d_work_key;
    a.knobs = Knobs{};
    cudaStream_t st = (cudaStream_t)stream;
    switch (pot_kind) {
    case CTB_POT_POLY4: return ctb::launch_setup_poly4(a, st);
    case CTB_POT_SPLINE: return ctb::launch_setup_spline(a, st);
    case CTB_POT_THERMAL1F:
        if (!(jtab_ok(jb) && jtab_ok(jf))) return CTB_ERR_BAD_ARG;
        return ctb::launch_setup_thermal1f(a, st);
    case CTB_POT_MODEL1_LINE:
        if (!jtab_ok(jb)) return CTB_ERR_BAD_ARG;
        return ctb::launch_setup_model1_line(a, st);
    case CTB_POT_MODEL1F:
        if (!jtab_ok(jb)) return CTB_ERR_BAD_ARG;
        return ctb::launch_setup_model1f(a, st);
    }
    return CTB_ERR_UNSUPPORTED;
}

CTB_API int ctb_jspline(const ctb_jtable *tab, int deriv, const double *d_theta, double *d_out, int64_t n, void *stream)
{
    if (!jtab_ok(tab) || n < 0) return CTB_ERR_BAD_ARG;
    if (n == 0) return CTB_OK;
    if (!d_theta || !d_out) return CTB_ERR_BAD_ARG;
    JTab t = to_jtab(tab);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t sm = table_bytes(tab);
    int64_t gneed = (n + 256 * 4 - 1) / (256 * 4);
    const int64_t gcap = (int64_t)sm_count() * 4; // persistent: 4 blocks x 48 KB per SM
    int g = (int)(gneed < gcap ? (gneed < 1 ? 1 : gneed) : gcap);
    switch (deriv) {
    case 0: jspline_kernel<0><<<g, 256, sm, st>>>(t, d_theta, d_out, n); break;
    case 1: jspline_kernel<1><<<g, 256, sm, st>>>(t, d_theta, d_out, n); break;
    case 2: jspline_kernel<2><<<g, 256, sm, st>>>(t, d_theta, d_out, n); break;
    case 3: jspline_kernel<3><<<g, 256, sm, st>>>(t, d_theta, d_out, n); break;
    default: return CTB_ERR_UNSUPPORTED;
    }
    return (int)cudaGetLastError();
}

CTB_API int ctb_jseries(int which, int approx, int deriv, int nterms, const double *lowcoef, const double *d_x,
                        double *d_out, int64_t n, void *stream)
{
    if (n < 0 || (which != 0 && which != 1) || nterms < 0) return CTB_ERR_BAD_ARG;
    if (approx == CTB_APPROX_HIGH && (deriv < 0 || deriv > 3)) return CTB_ERR_UNSUPPORTED;
    if (approx == CTB_APPROX_LOW && (deriv != 0 || nterms > 50 || !lowcoef)) return CTB_ERR_UNSUPPORTED;
    if (approx != CTB_APPROX_HIGH && approx != CTB_APPROX_LOW) return CTB_ERR_UNSUPPORTED;
    if (n == 0) return CTB_OK;
    if (!d_x || !d_out) return CTB_ERR_BAD_ARG;
    LowCoef lc;
    for (int i = 0; i < 56; i++) lc.c[i] = 0.0;
    if (approx == CTB_APPROX_LOW)
        for (int i = 0; i < (which == 0 ? 55 : 54); i++) lc.c[i] = lowcoef[i];
    cudaStream_t st = (cudaStream_t)stream;
    int g = grid_for(n, 256);
    if (which == 0 && approx == CTB_APPROX_HIGH) jseries_kernel<0, CTB_APPROX_HIGH><<<g, 256, 0, st>>>(deriv, nterms, lc, d_x, d_out, n);
    else if (which == 1 && approx == CTB_APPROX_HIGH) jseries_kernel<1, CTB_APPROX_HIGH><<<g, 256, 0, st>>>(deriv, nterms, lc, d_x, d_out, n);
    else if (which == 0) jseries_kernel<0, CTB_APPROX_LOW><<<g, 256, 0, st>>>(deriv, nterms, lc, d_x, d_out, n);
    else jseries_kernel<1, CTB_APPROX_LOW><<<g, 256, 0, st>>>(deriv, nterms, lc, d_x, d_out, n);
    return (int)cudaGetLastError();
}

CTB_API int ctb_jexact(int which, int mode, const double *d_x, double *d_out, int64_t n, void *stream)
{
    if (n < 0 || (which != 0 && which != 1) || mode < 0 || mode > 2) return CTB_ERR_BAD_ARG;
    if (n == 0) return CTB_OK;
    if (!d_x || !d_out) return CTB_ERR_BAD_ARG;
    const unsigned g = (unsigned)((n + 127) / 128);
    if (which == 0) jexact_kernel<0><<<g, 128, 0, (cudaStream_t)stream>>>(mode, d_x, d_out, n);
    else jexact_kernel<1><<<g, 128, 0, (cudaStream_t)stream>>>(mode, d_x, d_out, n);
    return (int)cudaGetLastError();
}

CTB_API int ctb_vtot_model1(const double *model8, const ctb_jtable *jb, const double *d_X, const double *d_T, double *d_out,
                            int64_t n, int parts, void *stream)
{
    if (!model8 || !jtab_ok(jb) || n < 0 || parts < 1 || parts > CTB_V_TOTAL) return CTB_ERR_BAD_ARG;
    if (n == 0) return CTB_OK;
    if (!d_X || !d_T || !d_out) return CTB_ERR_BAD_ARG;
    if ((uintptr_t)d_X & 15) return CTB_ERR_BAD_ARG; // X is read as one 16-byte (phi1, phi2) pair per point
    Model8 m{model8[0], model8[1], model8[2], model8[3], model8[4], model8[5], model8[6], model8[7]};
    vtot_model1_kernel<<<grid_for(n, 256), 256, table_bytes(jb), (cudaStream_t)stream>>>(m, to_jtab(jb), d_X, d_T, d_out, n, parts);
    return (int)cudaGetLastError();
}

CTB_API int ctb_vtot_model1_grid(const double *model8, const double *line4, const ctb_jtable *jb, const double *d_s,
                         int64_t ns, const double *d_T, int64_t nT, double *d_out, void *stream)
{
    if (!model8 || !line4 || !jtab_ok(jb) || ns < 0 || nT < 0) return CTB_ERR_BAD_ARG;
    if (ns == 0 || nT == 0) return CTB_OK;
    if (!d_s || !d_T || !d_out) return CTB_ERR_BAD_ARG;
    Model8 m{model8[0], model8[1], model8[2], model8[3], model8[4], model8[5], model8[6], model8[7]};
    if (nT >= VGT_THREADS / 2 && jb->loc_inv_du != 0.0 && jb->xmax != 0.0 && ((uintptr_t)jb->d_coef & 15) == 0 &&
        jb->nint < VGT_STAGE_PER_THREAD * VGT_THREADS) { // tiled form: one temperature per thread
        const int64_t col_tiles = (nT + VGT_THREADS - 1) / VGT_THREADS;
        int64_t row_chunks = ((int64_t)sm_count() * VGT_MIN_BLOCKS + col_tiles - 1) / col_tiles;
        if (row_chunks > ns) row_chunks = ns;
        if (row_chunks < 1) row_chunks = 1;
        const int64_t rows_per_block = (ns + row_chunks - 1) / row_chunks;
        row_chunks = (ns + rows_per_block - 1) / rows_per_block;
        if (col_tiles * row_chunks > 0x7fffffff || rows_per_block > 0x7fffffff) return CTB_ERR_UNSUPPORTED;
        const size_t tsm = table_bytes(jb) + (6 + VGT_ROWS * 4) * sizeof(double);
        cudaFuncSetAttribute(vtot_model1_grid_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
        vtot_model1_grid_tiled_kernel<<<(unsigned)(col_tiles * row_chunks), VGT_THREADS, tsm, (cudaStream_t)stream>>>(
            m, line4[0], line4[1], line4[2], line4[3], to_jtab(jb), d_s, ns, d_T, nT, d_out, (int)rows_per_block, (int)col_tiles);
        return (int)cudaGetLastError();
    }
    // persistent grid: 4 blocks per SM unless that would make a block span more than VG_ROWCAP field points
    const int64_t total = ns * nT;
    int64_t nblk = (int64_t)sm_count() * 4;
    const int64_t min_blocks = (total + (VG_ROWCAP - 2) * nT - 1) / ((VG_ROWCAP - 2) * nT);
    if (nblk > (total + VG_THREADS - 1) / VG_THREADS) nblk = (total + VG_THREADS - 1) / VG_THREADS;
    // last word: a block never spans more than VG_ROWCAP field points (chunk <= (VG_ROWCAP - 2) nT points touch at
    // most VG_ROWCAP - 1 rows), also for nT < 4 where 256 points per block would be 64+ rows
    if (nblk < min_blocks) nblk = min_blocks;
    const int64_t chunk = (total + nblk - 1) / nblk;
    const size_t vg_smem = table_bytes(jb) + VG_ROWCAP * 6 * sizeof(double);
    cudaFuncSetAttribute(vtot_model1_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vg_smem);
    vtot_model1_grid_kernel<<<(unsigned)nblk, VG_THREADS, vg_smem,
                              (cudaStream_t)stream>>>(m, line4[0], line4[1], line4[2], line4[3], to_jtab(jb), d_s, ns,
                                                      d_T, nT, d_out, chunk);
    return (int)cudaGetLastError();
}

CTB_API int ctb_vtot_thermal1f(const double *model64, double rad, const ctb_jtable *jb, const ctb_jtable *jf,
                               const double *d_phi, const double *d_T, double *d_out, int64_t n, int parts, void *stream)
{
    if (!model64 || !jtab_ok(jb) || !jtab_ok(jf) || n < 0 || parts < 1 || parts > CTB_V_TOTAL) return CTB_ERR_BAD_ARG;
    if (n == 0) return CTB_OK;
    if (!d_phi || !d_T || !d_out) return CTB_ERR_BAD_ARG;
    Thermal1FModel m;
    for (int i = 0; i < CTB_NPARAM_THERMAL1F; i++) m.v[i] = model64[i];
    const size_t sm = table_bytes_compact(jb) + table_bytes_compact(jf) + 16;
    cudaFuncSetAttribute(vtot_thermal1f_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int64_t g = (n + 255) / 256;
    if (g > (int64_t)sm_count() * 2) g = (int64_t)sm_count() * 2; // persistent: two 80 KB table pairs per SM
    vtot_thermal1f_kernel<<<(unsigned)g, 256, sm, (cudaStream_t)stream>>>(m, rad, to_jtab(jb), to_jtab(jf), d_phi, d_T, d_out,
                                                                         n, parts);
    return (int)cudaGetLastError();
}

CTB_API int ctb_potential(int pot_kind, const double *d_params, const ctb_jtable *jb, const ctb_jtable *jf,
                  const double *d_phi, double *d_out, int64_t n, void *stream)
{
    if (n < 0) return CTB_ERR_BAD_ARG;
    if (n == 0) return CTB_OK;
    if (!d_params || !d_phi || !d_out) return CTB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    int g = grid_for(n, 128);
    switch (pot_kind) {
    case CTB_POT_POLY4: potential_kernel<PotPoly4><<<g, 256, 0, st>>>(d_params, to_jtab(jb), to_jtab(jf), d_phi, d_out, n); break;
    case CTB_POT_SPLINE: potential_kernel<PotSpline><<<g, 256, 0, st>>>(d_params, to_jtab(jb), to_jtab(jf), d_phi, d_out, n); break;
    case CTB_POT_THERMAL1F:
        if (!(jtab_ok(jb) && jtab_ok(jf))) return CTB_ERR_BAD_ARG;
        launch_k(potential_kernel<PotThermal1F>, g, 128, pot_smem_bytes<PotThermal1F>(jb->nint), st, d_params, to_jtab(jb), to_jtab(jf), d_phi, d_out, n);
        break;
    case CTB_POT_MODEL1_LINE:
        if (!jtab_ok(jb)) return CTB_ERR_BAD_ARG;
        launch_k(potential_kernel<PotModel1Line>, g, 128, pot_smem_bytes<PotModel1Line>(jb->nint), st, d_params, to_jtab(jb), to_jtab(jf), d_phi, d_out, n);
        break;
    case CTB_POT_MODEL1F:
        if (!jtab_ok(jb)) return CTB_ERR_BAD_ARG;
        launch_k(potential_kernel<PotModel1F>, g, 128, pot_smem_bytes<PotModel1F>(jb->nint), st, d_params, to_jtab(jb), to_jtab(jf), d_phi, d_out, n);
        break;
    default: return CTB_ERR_UNSUPPORTED;
    }
    return (int)cudaGetLastError();
}

CTB_API int ctb_potential_batch(int pot_kind, const double *d_params, const ctb_jtable *jb, const ctb_jtable *jf,
                                const double *d_phi, double *d_out, int64_t n, void *stream)
{
    if (n < 0) return CTB_ERR_BAD_ARG;
    if (n == 0) return CTB_OK;
    if (!d_params || !d_phi || !d_out) return CTB_ERR_BAD_ARG;
    return launch_rows<false>(pot_kind, d_params, jb, jf, d_phi, nullptr, 0.0, d_out, nullptr, n, (cudaStream_t)stream);
}

CTB_API int ctb_minimize_1d(int pot_kind, const double *d_params, const ctb_jtable *jb, const ctb_jtable *jf,
                            const double *d_lo, const double *d_hi, double xtol_rel, int64_t n, double *d_xmin,
                            double *d_vmin, void *stream)
{
    if (n < 0 || !(xtol_rel > 0)) return CTB_ERR_BAD_ARG;
    if (n == 0) return CTB_OK;
    if (!d_params || !d_lo || !d_hi || !d_xmin) return CTB_ERR_BAD_ARG;
    return launch_rows<true>(pot_kind, d_params, jb, jf, d_lo, d_hi, xtol_rel, d_xmin, d_vmin, n, (cudaStream_t)stream);
}

CTB_API int ctb_wall_batch(int pot_kind, const ctb_bounce_in *in, const ctb_wall_opts *o, const ctb_jtable *jb,
                           const ctb_jtable *jf, const ctb_bounce_out *out, void *stream)
{
    if (!in || !out || !o || in->n < 0) return CTB_ERR_BAD_ARG;
    if (in->n == 0) return CTB_OK;
    if (!in->d_params || !in->d_phi_abs || !in->d_phi_meta) return CTB_ERR_BAD_ARG;
    if (o->npoints < 2 || o->npoints > 4096) return CTB_ERR_UNSUPPORTED;
    if (out->d_prof_R && (!out->d_prof_Phi || !out->d_prof_dPhi || out->prof_stride < o->npoints)) return CTB_ERR_BAD_ARG;
    WallArgs a;
    a.params = in->d_params; a.phi_abs = in->d_phi_abs; a.phi_meta = in->d_phi_meta; a.n = in->n;
    a.knobs = WallKnobs{o->Fguess, o->Ftol, o->phitol, o->rmin, o->rmax, o->phi0_rel, o->phi_eps, o->npoints};
    a.jb = to_jtab(jb); a.jf = to_jtab(jf);
    a.out = *out;
    cudaStream_t st = (cudaStream_t)stream;
    switch (pot_kind) {
    case CTB_POT_POLY4: return ctb::launch_wall_poly4(a, st);
    case CTB_POT_SPLINE: return ctb::launch_wall_spline(a, st);
    case CTB_POT_THERMAL1F:
        if (!(jtab_ok(jb) && jtab_ok(jf))) return CTB_ERR_BAD_ARG;
        return ctb::launch_wall_thermal1f(a, st);
    case CTB_POT_MODEL1_LINE:
        if (!jtab_ok(jb)) return CTB_ERR_BAD_ARG;
        return ctb::launch_wall_model1_line(a, st);
    case CTB_POT_MODEL1F:
        if (!jtab_ok(jb)) return CTB_ERR_BAD_ARG;
        return ctb::launch_wall_model1f(a, st);
    }
    return CTB_ERR_UNSUPPORTED;
}

CTB_API int ctb_find_action(int pot_kind, const double *d_params, const ctb_jtable *jb, const ctb_jtable *jf,
                            const double *d_phi_meta, const double *d_R, const double *d_Phi, const double *d_dPhi,
                            int64_t stride, const int32_t *d_npts, int64_t n, double alpha, double *d_S, void *stream)
{
    if (n < 0 || stride < 1 || !(alpha >= 0)) return CTB_ERR_BAD_ARG;
    if (n == 0) return CTB_OK;
    if (!d_params || !d_phi_meta || !d_R || !d_Phi || !d_dPhi || !d_npts || !d_S) return CTB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    switch (pot_kind) {
    case CTB_POT_POLY4:
        return launch_find_action<PotPoly4>(d_params, jb, jf, d_phi_meta, d_R, d_Phi, d_dPhi, stride, d_npts, n, alpha, d_S, st);
    case CTB_POT_SPLINE:
        return launch_find_action<PotSpline>(d_params, jb, jf, d_phi_meta, d_R, d_Phi, d_dPhi, stride, d_npts, n, alpha, d_S, st);
    case CTB_POT_THERMAL1F:
        if (!(jtab_ok(jb) && jtab_ok(jf))) return CTB_ERR_BAD_ARG;
        return launch_find_action<PotThermal1F>(d_params, jb, jf, d_phi_meta, d_R, d_Phi, d_dPhi, stride, d_npts, n, alpha, d_S, st);
    case CTB_POT_MODEL1_LINE:
        if (!jtab_ok(jb)) return CTB_ERR_BAD_ARG;
        return launch_find_action<PotModel1Line>(d_params, jb, jf, d_phi_meta, d_R, d_Phi, d_dPhi, stride, d_npts, n, alpha, d_S, st);
    case CTB_POT_MODEL1F:
        if (!jtab_ok(jb)) return CTB_ERR_BAD_ARG;
        return launch_find_action<PotModel1F>(d_params, jb, jf, d_phi_meta, d_R, d_Phi, d_dPhi, stride, d_npts, n, alpha, d_S, st);
    }
    return CTB_ERR_UNSUPPORTED;
}

CTB_API int ctb_trace_minima(int model_kind, const double *d_params, int64_t param_stride, const ctb_jtable *jb,
                             const ctb_jtable *jf, const double *d_X0, const double *d_T, int64_t T_stride,
                             const int32_t *d_t_begin, const int32_t *d_t_dir, int64_t n, int32_t nT, const ctb_min_opts *opts,
                             double *d_X, double *d_V, int32_t *d_flag, int32_t *d_iters, void *stream)
{
    if (n < 0 || nT < 0 || !opts || param_stride < 0 || T_stride < 0) return CTB_ERR_BAD_ARG;
    if (model_kind != CTB_MODEL_MODEL1 && model_kind != CTB_MODEL_THERMAL1F) return CTB_ERR_UNSUPPORTED;
    if (n == 0 || nT == 0) return CTB_OK;
    if (!d_params || !d_X0 || !d_T || !d_X || !d_flag || !jtab_ok(jb)) return CTB_ERR_BAD_ARG;
    if (!(opts->x_eps > 0.0) || !(opts->xtol > 0.0) || !(opts->jump_tol > 0.0) || opts->max_iter < 1) return CTB_ERR_BAD_ARG;
    const MinKnobs k{opts->x_eps, opts->xtol, opts->jump_tol, opts->max_iter};
    if (model_kind == CTB_MODEL_THERMAL1F) {
        if (!jtab_ok(jf)) return CTB_ERR_BAD_ARG;
        return launch_trace_minima_thermal1f(d_params, param_stride, to_jtab(jb), to_jtab(jf), d_X0, d_T, T_stride, d_t_begin,
                                             d_t_dir, n, nT, k, d_X, d_V, d_flag, d_iters, (cudaStream_t)stream);
    }
    return launch_trace_minima_model1(d_params, param_stride, to_jtab(jb), d_X0, d_T, T_stride, d_t_begin, d_t_dir, n, nT, k,
                                      d_X, d_V, d_flag, d_iters, (cudaStream_t)stream);
}

CTB_API int ctb_exact_solution(double alpha, const double *d_in, double *d_out, int64_t n, void *stream)
{
    if (n < 0 || !(alpha >= 0.0 && alpha <= 64.0)) return CTB_ERR_BAD_ARG;
    if (n == 0) return CTB_OK;
    if (!d_in || !d_out) return CTB_ERR_BAD_ARG;
    const unsigned g = (unsigned)((n + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
    if (alpha == 1.0) exact_solution_kernel<1><<<g, 128, 0, st>>>(alpha, d_in, d_out, n);
    else if (alpha == 2.0) exact_solution_kernel<2><<<g, 128, 0, st>>>(alpha, d_in, d_out, n);
    else if (alpha == 3.0) exact_solution_kernel<3><<<g, 128, 0, st>>>(alpha, d_in, d_out, n);
    else if (alpha == 4.0) exact_solution_kernel<4><<<g, 128, 0, st>>>(alpha, d_in, d_out, n);
    else exact_solution_kernel<CTB_ALPHA_REAL><<<g, 128, 0, st>>>(alpha, d_in, d_out, n);
    return (int)cudaGetLastError();
}

CTB_API int ctb_copy_blocks(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                            void *stream)
{
    if (width == 0 || height == 0) return CTB_OK;
    if (!dst || !src || dpitch < width || spitch < width) return CTB_ERR_BAD_ARG;
    return (int)cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDefault, (cudaStream_t)stream);
}

CTB_API int ctb_host_register(void *p, size_t bytes)
{
    if (!p || !bytes) return CTB_ERR_BAD_ARG;
    return (int)cudaHostRegister(p, bytes, cudaHostRegisterPortable);
}

CTB_API int ctb_host_unregister(void *p)
{
    if (!p) return CTB_ERR_BAD_ARG;
    return (int)cudaHostUnregister(p);
}

CTB_API int ctb_fp64_peak(int64_t iters, int blocks, int threads, double *d_sink, float *ms_out, void *stream)
{
    if (!d_sink || !ms_out || iters <= 0 || blocks <= 0 || threads <= 0 || threads > 1024) return CTB_ERR_BAD_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fp64_peak_kernel<<<blocks, threads, 0, st>>>(iters / 8 + 1, d_sink); // warm-up
    cudaEventRecord(e0, st);
    fp64_peak_kernel<<<blocks, threads, 0, st>>>(iters, d_sink);
    cudaEventRecord(e1, st);
    cudaError_t err = cudaEventSynchronize(e1);
    if (err == cudaSuccess) cudaEventElapsedTime(ms_out, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (err != cudaSuccess) return (int)err;
    return (int)cudaGetLastError();
}

} // extern "C"
