// ctb_device.cuh -- device-side building blocks of the batched bounce solver (sm_100a, FP64).
//
// Behavioural spec: SURVEY.md Appendix A.  Reference locations are cited per function
// (cosmoTransitions/tunneling1D.py = T1D, helper_functions.py = HF, finiteT.py = FT,
// generic_potential.py = GP).  One THREAD owns one instance; all state lives in registers.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include "../../include/ctbounce.h"

namespace ctb {

#define CTB_DEV __device__ __forceinline__
// PINNED ARITHMETIC.  Everything that steers control flow of the integration (Cash-Karp stages, error
// norm, step controller, dense-output cubic, root finds, potential derivatives) is written with explicit
// fma() / __dmul_rn / __dadd_rn so that the rounding of every operation is fixed in the source instead of
// being left to the compiler's FMA-contraction choices.  The shoot kernel, the save kernel, the fused
// kernel and the warp kernel are separate compilations of this code and must produce the SAME trajectory
// bit for bit (e.g. the save pass has to find the end point rf inside the very step in which the shot found it).
#define CTB_MUL(a, b) __dmul_rn((a), (b))
#define CTB_ADD(a, b) __dadd_rn((a), (b))
#define CTB_SUB(a, b) __dsub_rn((a), (b))
#ifndef CTB_THERMAL_MIN_BLOCKS
#define CTB_THERMAL_MIN_BLOCKS 4
#endif

// ------------------------------------------------------------------------------------------------
// Jb / Jf spline tables (piecewise cubic, see ctbounce.h)                      FT:167-174,197-204
//
// Every kernel that evaluates a table stages it in DYNAMIC shared memory first.  The staged copy is addressed by
// element offsets into the dynamic shared-memory window (ctb_dynsmem), not by pointers: the compiler then emits
// LDS / LDS.128 directly.  (Through generic pointers held in a struct the same accesses were generic LD plus
// register <-> uniform-register traffic: half of the instructions of the out-of-line V of round 1.)
extern __shared__ double ctb_dynsmem[];

enum { JT_GLOBAL = 0, JT_COMPACT = 1, JT_RECORDS = 2 };
struct JTab {
    const double *brk;  // [nint + 1] global memory
    const double *coef; // [nint * 4] global memory
    int mode;           // JT_GLOBAL: brk / coef above; JT_COMPACT: i_brk / i_coef; JT_RECORDS: i_rec
    int i_brk, i_coef;  // offsets (in doubles) into ctb_dynsmem
    int i_rec;          // records {brk_j, c0, c1, c2, c3, brk_{j+1}} (48 B each)
    int nint;
    double xmin, xmax;
    float loc_scale, loc_u0, loc_inv_du; // closed-form locator (see ctbounce.h), slope per unit of log2; loc_inv_du == 0: bisection
    float loc_c;                         // -loc_u0 loc_inv_du - 1 - 1e-3: offset of the one-sided interval guess
};

#define CTB_TABLE_DOUBLES(nint) ((size_t)(nint) * 6)

// host side: the C-ABI table descriptor -> the kernels' view of it (NULL: an empty table)
inline JTab to_jtab(const ctb_jtable *t)
{
    JTab j{nullptr, nullptr, JT_GLOBAL, 0, 0, 0, 0, 0.0, 0.0, 0.0f, 0.0f, 0.0f, 0.0f};
    if (t) {
        j.brk = t->d_brk; j.coef = t->d_coef; j.nint = t->nint; j.xmin = t->xmin; j.xmax = t->xmax;
        j.loc_scale = (float)t->loc_scale; j.loc_u0 = (float)t->loc_u0;
        j.loc_inv_du = (float)(t->loc_inv_du * 0.69314718055994531); // per unit of log2: asinh(z) = ln 2 * log2(|z| + sqrt(z^2 + 1))
        j.loc_c = (float)(-t->loc_u0 * t->loc_inv_du - 1.001);
    }
    return j;
}

// Copies a table into shared memory (whole block cooperates; caller syncs) as one 48-byte record per knot
// interval -- left breakpoint, the four coefficients, right breakpoint -- so that an evaluation touches
// exactly three 16-byte words: the interval test and the polynomial come from the same record.
// `smem` must point into the dynamic shared-memory window, 16-byte aligned.
CTB_DEV JTab stage_table(const JTab &g, double *smem)
{
    for (int j = threadIdx.x; j < g.nint; j += blockDim.x) {
        double *r = smem + 6 * j;
        r[0] = g.brk[j];
        r[1] = g.coef[4 * j]; r[2] = g.coef[4 * j + 1]; r[3] = g.coef[4 * j + 2]; r[4] = g.coef[4 * j + 3];
        r[5] = g.brk[j + 1];
    }
    JTab t = g;
    t.mode = JT_RECORDS;
    t.i_rec = (int)(smem - ctb_dynsmem);
    return t;
}

// Compact staging (40 KB) for the bounce / minimiser kernels: a 32-byte header (so that the out-of-line V needs
// one offset, not a descriptor), breakpoints, coefficients; after it come the per-thread parameter slots of the
// potential functor (CTB_SLOT_STRIDE doubles per parameter, slot k of thread t at [k * CTB_SLOT_STRIDE + t]).
// Four 128-thread blocks (table + slots <= 55 KB) fit one SM.
#define CTB_SLOT_STRIDE 128
#define CTB_TABLE_DOUBLES_COMPACT(nint) ((size_t)(nint) * 5 + 6)
struct JHeader { double xmin, xmax; int last; float loc_scale, loc_inv_du, loc_c; };
CTB_DEV int jtab_compact_nb(int nint) { return (nint + 2) & ~1; }
CTB_DEV JTab stage_table_compact(const JTab &g, double *smem)
{
    const int nb = jtab_compact_nb(g.nint);
    if (threadIdx.x == 0) {
        JHeader *h = reinterpret_cast<JHeader *>(smem);
        h->xmin = g.xmin; h->xmax = g.xmax; h->last = g.nint - 1;
        h->loc_scale = g.loc_scale; h->loc_inv_du = g.loc_inv_du; h->loc_c = g.loc_c;
    }
    for (int i = threadIdx.x; i < g.nint + 1; i += blockDim.x) smem[4 + i] = g.brk[i];
    for (int i = threadIdx.x; i < 4 * g.nint; i += blockDim.x) smem[4 + nb + i] = g.coef[i];
    JTab t = g;
    t.mode = JT_COMPACT;
    t.i_brk = (int)(smem - ctb_dynsmem) + 4;
    t.i_coef = t.i_brk + nb;
    return t;
}
// offset of the parameter slots that follow a compact table staged at ctb_dynsmem[i_brk - 4]
CTB_DEV int jtab_slot_base(const JTab &t) { return t.i_coef + 4 * t.nint; }

// One-sided FP32 guess of the knot interval from the sinh spacing of the reference's sample grid
// (x_k = |xmin| sinh(u_k)/20, u_k uniform; not-a-knot spline: interval j starts at sample j + 1 for j >= 1):
// k = (asinh(x scale) - u0) / du computed with the MUFU approximations is good to ~3e-4 of an interval, so
// g = floor(k - 1 - 1e-3), clamped, is either the interval that holds x or the one below it -- never above
// (checked against bisection on 3e6 arguments incl. every breakpoint +- 1 ulp, with the MUFU results perturbed by
// +- 2^-21: 0.14-0.19 % one low, none high).
CTB_DEV int jtab_guess_f(double x, float loc_scale, float loc_inv_du, float loc_c, int last)
{
    const float z = (float)x * loc_scale;
    const float a = fabsf(z);
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(a, a, 1.0f)));
    // (a + r >= 1: the raw MUFU log2 needs none of __logf's denormal handling; ln 2 is folded into loc_inv_du by to_jtab)
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(a + r));
    const float u = copysignf(l2, z);
    const int g = __float2int_rd(fmaf(u, loc_inv_du, loc_c));
    return min(max(g, 0), last);
}
CTB_DEV int jtab_guess(const JTab &t, double x) { return jtab_guess_f(x, t.loc_scale, t.loc_inv_du, t.loc_c, t.nint - 1); }

CTB_DEV double jtab_brk(const JTab &t, int j)
{
    return t.mode == JT_COMPACT ? ctb_dynsmem[t.i_brk + j] : t.mode == JT_RECORDS ? ctb_dynsmem[t.i_rec + 6 * j] : t.brk[j];
}

CTB_DEV int jtab_locate(const JTab &t, double x)
{
    // largest j in [0, nint-1] with brk[j] <= x  (j = 0 below the table: first piece extrapolates)
    int lo = 0, hi = t.nint - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (x >= jtab_brk(t, mid)) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <int DER>
CTB_DEV double jtab_poly(double dx, double c0, double c1, double c2, double c3)
{
    if (DER == 0) return fma(dx, fma(dx, fma(dx, c3, c2), c1), c0);
    if (DER == 1) return fma(dx, fma(dx, __dmul_rn(3.0, c3), __dmul_rn(2.0, c2)), c1);
    if (DER == 2) return fma(dx, __dmul_rn(6.0, c3), __dmul_rn(2.0, c2));
    return __dmul_rn(6.0, c3);
}

// Compact staged table, closed-form locator: guess, the two breakpoints around it, select, the coefficient pair.
// Branch-free (two look-ups of one V interleave); relies on the one-sidedness of the guess.
template <int DER>
CTB_DEV double jtab_eval_compact(int i_brk, int i_coef, int last, double xmin, double xmax, float loc_scale,
                                 float loc_inv_du, float loc_c, double theta)
{
    const double x = (theta < xmin) ? xmin : theta;
    int j = jtab_guess_f(x, loc_scale, loc_inv_du, loc_c, last);
    const double bg = ctb_dynsmem[i_brk + j], bn = ctb_dynsmem[i_brk + j + 1];
    const bool up = (x >= bn) && (j < last);
    const double b0 = up ? bn : bg;
    j += up;
    const double2 *c = reinterpret_cast<const double2 *>(ctb_dynsmem + i_coef) + 2 * j;
    const double2 c01 = c[0], c23 = c[1];
    const double v = jtab_poly<DER>(__dsub_rn(x, b0), c01.x, c01.y, c23.x, c23.y);
    return (theta > xmax) ? 0.0 : v;
}

// Piecewise-cubic evaluation, any staging mode.
template <int DER>
CTB_DEV double jtab_eval(const JTab &t, double theta)
{
    const int last = t.nint - 1;
    if (t.loc_inv_du != 0.0f && t.mode == JT_COMPACT)
        return jtab_eval_compact<DER>(t.i_brk, t.i_coef, last, t.xmin, t.xmax, t.loc_scale, t.loc_inv_du, t.loc_c, theta);
    const double x = (theta < t.xmin) ? t.xmin : theta;
    double b0, c0, c1, c2, c3;
    if (t.loc_inv_du != 0.0f && t.mode == JT_RECORDS) {
        // one staged 48-byte record (three 16-byte loads); the next one when the guess was one low (~0.2 %)
        int j = jtab_guess(t, x);
        const double2 *r = reinterpret_cast<const double2 *>(ctb_dynsmem + t.i_rec) + 3 * j;
        double2 q0 = r[0], q1 = r[1], q2 = r[2];
        if (x >= q2.y && j < last) {
            q0 = r[3]; q1 = r[4]; q2 = r[5];
        }
        b0 = q0.x; c0 = q0.y; c1 = q1.x; c2 = q1.y; c3 = q2.x;
    } else {
        int j;
        if (t.loc_inv_du != 0.0f) {
            j = jtab_guess(t, x);
            j += (j < last && x >= jtab_brk(t, j + 1));
        } else {
            j = jtab_locate(t, x);
        }
        b0 = jtab_brk(t, j);
        if (t.mode == JT_RECORDS) {
            const double2 *r = reinterpret_cast<const double2 *>(ctb_dynsmem + t.i_rec) + 3 * j;
            const double2 q0 = r[0], q1 = r[1], q2 = r[2];
            c0 = q0.y; c1 = q1.x; c2 = q1.y; c3 = q2.x;
        } else {
            const double2 *c = (t.mode == JT_COMPACT ? reinterpret_cast<const double2 *>(ctb_dynsmem + t.i_coef)
                                                     : reinterpret_cast<const double2 *>(t.coef)) + 2 * j;
            const double2 c01 = c[0], c23 = c[1];
            c0 = c01.x; c1 = c01.y; c2 = c23.x; c3 = c23.y;
        }
    }
    const double v = jtab_poly<DER>(__dsub_rn(x, b0), c0, c1, c2, c3);
    return (theta > t.xmax) ? 0.0 : v;
}

// log(x) for positive, finite, NORMAL x (the callers add 1e-100 to |m^2 / mu^2|, GP:250): branch-free, < 1 ulp.
// Argument reduction x = 2^k m, m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = f / (2 + f), f = m - 1, with the
// classic degree-14 even remainder polynomial (Lg1..Lg7) split into two chains; the division is a Newton reciprocal
// plus one residual correction.  ~35 instructions where the library routine (special-case branches, call ABI) costs
// ~80 inside the out-of-line V.  Non-finite x passes through (NaN stays NaN, inf stays inf).
CTB_DEV double log_pos(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
    const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                 Lg7 = 1.479819860511658591e-01;
    int hx = __double2hiint(x);
    const int lx = __double2loint(x);
    hx += 0x3ff00000 - 0x3fe6a09e;
    const int k = (hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffff) + 0x3fe6a09e;
    const double m = __hiloint2double(hx, lx);
    const double f = __dadd_rn(m, -1.0);
    const double hfsq = __dmul_rn(__dmul_rn(0.5, f), f);
    const double d = __dadd_rn(2.0, f);
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    double e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    e = fma(-d, r, 1.0);
    r = fma(r, e, r);
    double s = __dmul_rn(f, r);
    s = fma(fma(-s, d, f), r, s);
    const double z = __dmul_rn(s, s);
    const double w = __dmul_rn(z, z);
    const double t1 = __dmul_rn(w, fma(w, fma(w, Lg6, Lg4), Lg2));
    const double t2 = __dmul_rn(z, fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1));
    const double R = __dadd_rn(t2, t1);
    const double dk = (double)k;
    // s (hfsq + R) + dk ln2_lo - hfsq + f + dk ln2_hi, summed in fdlibm's order
    double v = fma(s, __dadd_rn(hfsq, R), __dmul_rn(dk, ln2_lo));
    v = __dadd_rn(__dsub_rn(v, hfsq), f);
    v = fma(dk, ln2_hi, v);
    return (fabs(x) < 1.7976931348623157e308) ? v : x;
}

// ------------------------------------------------------------------------------------------------
// potentials
#define CTB_PI 3.14159265358979323846

#ifndef CTB_POLY4_MIN_BLOCKS
#define CTB_POLY4_MIN_BLOCKS 3
#endif
struct PotPoly4 {
    static constexpr bool analytic = true;
    static constexpr int min_blocks = CTB_POLY4_MIN_BLOCKS; // 128-thread blocks per SM the bounce kernel is compiled for (<= 168 regs)
    static constexpr int nparam = CTB_NPARAM_POLY4;
    static constexpr int nslot = 0;
    double c0, c1, c2, c3, c4;
    CTB_DEV void load(const double *p, const JTab *, const JTab *)
    {
        c0 = p[0]; c1 = p[1]; c2 = p[2]; c3 = p[3]; c4 = p[4];
    }
    // V steers the barrier bisection and fminbound (discrete decisions): keep the reference's
    // rounding (no FMA contraction).  dV / d2V feed the integrator and may contract.
    CTB_DEV double V(double x) const
    {
        double v = __dadd_rn(c3, __dmul_rn(x, c4));
        v = __dadd_rn(c2, __dmul_rn(x, v));
        v = __dadd_rn(c1, __dmul_rn(x, v));
        return __dadd_rn(c0, __dmul_rn(x, v));
    }
    CTB_DEV double V_quad(double x) const { return c0 + x * (c1 + x * (c2 + x * (c3 + x * c4))); } // integrand only
    CTB_DEV double dV(double x) const
    {
        return fma(x, fma(x, fma(x, CTB_MUL(4.0, c4), CTB_MUL(3.0, c3)), CTB_MUL(2.0, c2)), c1);
    }
    CTB_DEV double d2V(double x) const { return fma(x, fma(x, CTB_MUL(12.0, c4), CTB_MUL(6.0, c3)), CTB_MUL(2.0, c2)); }
};

// Tabulated potential (CTB_POT_SPLINE): one cubic per knot interval, per-instance table in global memory
// (L1/L2-resident), located by bisection.  V, dV, d2V are the polynomial and its derivatives, i.e. what
// splev(x, tck, der=0|1|2) returns for the reference's SplinePath.
struct PotSpline {
    static constexpr bool analytic = true;
    static constexpr int min_blocks = 3;
    static constexpr int nparam = CTB_NPARAM_SPLINE;
    static constexpr int nslot = 0;
    const double *brk, *coef;
    int nint;
    CTB_DEV void load(const double *p, const JTab *, const JTab *)
    {
        nint = (int)p[0];
        if (nint < 1) nint = 1;
        if (nint > CTB_SPLINE_MAX_PIECES) nint = CTB_SPLINE_MAX_PIECES;
        brk = p + 1;
        coef = p + 1 + (CTB_SPLINE_MAX_PIECES + 1);
    }
    CTB_DEV int locate(double x) const
    {
        int lo = 0, hi = nint - 1;
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (x >= __ldg(brk + mid)) lo = mid; else hi = mid - 1;
        }
        return lo;
    }
    CTB_DEV double V(double x) const
    {
        const int j = locate(x);
        const double dx = x - __ldg(brk + j);
        const double *c = coef + 4 * j;
        return fma(dx, fma(dx, fma(dx, __ldg(c + 3), __ldg(c + 2)), __ldg(c + 1)), __ldg(c));
    }
    CTB_DEV double V_quad(double x) const { return V(x); }
    CTB_DEV double dV(double x) const
    {
        const int j = locate(x);
        const double dx = x - __ldg(brk + j);
        const double *c = coef + 4 * j;
        return fma(dx, fma(dx, CTB_MUL(3.0, __ldg(c + 3)), CTB_MUL(2.0, __ldg(c + 2))), __ldg(c + 1));
    }
    CTB_DEV double d2V(double x) const
    {
        const int j = locate(x);
        const double dx = x - __ldg(brk + j);
        const double *c = coef + 4 * j;
        return fma(dx, CTB_MUL(6.0, __ldg(c + 3)), CTB_MUL(2.0, __ldg(c + 2)));
    }
};

// Field-dependent part of model1's Vtot: the three boson masses (examples/testModel1.py:89-93) and
// V0 + V1 (examples/testModel1.py:78-80, GP:240-256).  Independent of T.
struct Model1Field { double m0, m1, mb, V01, V0, V1; };
CTB_DEV Model1Field model1_field(double l1, double l2, double mu2, double Y1, double Y2, double nb, double inv_rs, double v2,
                                 double phi1, double phi2)
{
    double a = l1 * (3 * phi1 * phi1 - v2);
    double b = l2 * (3 * phi2 * phi2 - v2);
    double A = .5 * (a + b);
    double B = sqrt(.25 * (a - b) * (a - b) + mu2 * mu2);
    Model1Field f;
    f.mb = Y1 * (phi1 * phi1 + phi2 * phi2) + Y2 * phi1 * phi2;
    f.m0 = A + B; f.m1 = A - B;
    double r = .25 * l1 * (phi1 * phi1 - v2) * (phi1 * phi1 - v2) + .25 * l2 * (phi2 * phi2 - v2) * (phi2 * phi2 - v2);
    r -= mu2 * phi1 * phi2;
    double y1 = 0.0;
    y1 += 1.0 * f.m0 * f.m0 * (log_pos(fabs(f.m0 * inv_rs) + 1e-100) - 1.5);
    y1 += 1.0 * f.m1 * f.m1 * (log_pos(fabs(f.m1 * inv_rs) + 1e-100) - 1.5);
    y1 += nb * f.mb * f.mb * (log_pos(fabs(f.mb * inv_rs) + 1e-100) - 1.5);
    f.V0 = r;
    f.V1 = y1 * (1.0 / (64 * CTB_PI * CTB_PI));
    f.V01 = r + f.V1;
    return f;
}

// Thermal functors with per-thread parameter SLOTS in shared memory.  V is ONE out-of-line function per functor
// (it is called from ~15 places: 5-point stencils, bisection, fminbound, integrand; one copy keeps the kernel inside
// the instruction cache) that takes the field value and two shared-memory offsets: the staged table and the
// thread's slots.  Everything it needs comes in by LDS; nothing goes through a `this` pointer (which would put the
// functor into local memory and turn every field access into a generic load).
CTB_DEV double &pot_slot(int i_slot, int k) { return ctb_dynsmem[i_slot + k * CTB_SLOT_STRIDE + (threadIdx.x & (CTB_SLOT_STRIDE - 1))]; }

struct TabView { // staged compact table as V sees it
    int i_brk, i_coef, last;
    double xmin, xmax;
    float loc_scale, loc_inv_du, loc_c;
};
CTB_DEV TabView tab_view(int i_tab)
{
    const JHeader *h = reinterpret_cast<const JHeader *>(ctb_dynsmem + i_tab);
    TabView t;
    t.xmin = h->xmin; t.xmax = h->xmax; t.last = h->last;
    t.loc_scale = h->loc_scale; t.loc_inv_du = h->loc_inv_du; t.loc_c = h->loc_c;
    t.i_brk = i_tab + 4;
    t.i_coef = t.i_brk + jtab_compact_nb(t.last + 1);
    return t;
}
CTB_DEV double tab_J(const TabView &t, double theta)
{
    return jtab_eval_compact<0>(t.i_brk, t.i_coef, t.last, t.xmin, t.xmax, t.loc_scale, t.loc_inv_du, t.loc_c, theta);
}

// examples/testModel1.py:62-116 through GP:312-337, on the line X = x0 + phi*nhat
struct PotModel1Line {
    static constexpr bool analytic = false;
    static constexpr int min_blocks = CTB_THERMAL_MIN_BLOCKS;
    static constexpr int nparam = CTB_NPARAM_MODEL1_LINE;
    static constexpr int nslot = 14;
    enum { S_L1, S_L2, S_MU2, S_Y1, S_Y2, S_NB, S_INV_RS, S_V2, S_INV_T2, S_T4C, S_X0, S_X1, S_N0, S_N1 };
    int i_tab, i_slot;
    CTB_DEV void load(const double *p, const JTab *jb_, const JTab *)
    {
        i_tab = jb_->i_brk - 4;
        i_slot = jtab_slot_base(*jb_);
        const double T = p[8];
        pot_slot(i_slot, S_L1) = p[0]; pot_slot(i_slot, S_L2) = p[1]; pot_slot(i_slot, S_MU2) = p[2];
        pot_slot(i_slot, S_Y1) = p[3]; pot_slot(i_slot, S_Y2) = p[4]; pot_slot(i_slot, S_NB) = p[5];
        pot_slot(i_slot, S_INV_RS) = 1.0 / p[6]; pot_slot(i_slot, S_V2) = p[7];
        pot_slot(i_slot, S_INV_T2) = 1.0 / (T * T + 1e-100);
        pot_slot(i_slot, S_T4C) = (T * T * T * T) / (2 * CTB_PI * CTB_PI);
        pot_slot(i_slot, S_X0) = p[9]; pot_slot(i_slot, S_X1) = p[10]; pot_slot(i_slot, S_N0) = p[11]; pot_slot(i_slot, S_N1) = p[12];
    }
    static __device__ __noinline__ double V_s(double phi, int i_tab, int i_slot)
    {
        const TabView t = tab_view(i_tab);
        const double nb = pot_slot(i_slot, S_NB), inv_T2 = pot_slot(i_slot, S_INV_T2);
        const Model1Field f = model1_field(pot_slot(i_slot, S_L1), pot_slot(i_slot, S_L2), pot_slot(i_slot, S_MU2),
                                           pot_slot(i_slot, S_Y1), pot_slot(i_slot, S_Y2), nb, pot_slot(i_slot, S_INV_RS),
                                           pot_slot(i_slot, S_V2), fma(phi, pot_slot(i_slot, S_N0), pot_slot(i_slot, S_X0)),
                                           fma(phi, pot_slot(i_slot, S_N1), pot_slot(i_slot, S_X1)));
        const double yt = tab_J(t, f.m0 * inv_T2) + tab_J(t, f.m1 * inv_T2) + nb * tab_J(t, f.mb * inv_T2);
        return f.V01 + yt * pot_slot(i_slot, S_T4C);
    }
    CTB_DEV double V(double phi) const { return V_s(phi, i_tab, i_slot); }
    CTB_DEV double V_quad(double phi) const { return V(phi); }
};

struct PotModel1F {
    static constexpr bool analytic = false;
    static constexpr int min_blocks = CTB_THERMAL_MIN_BLOCKS;
    static constexpr int nparam = CTB_NPARAM_MODEL1F;
    static constexpr int nslot = 7;
    enum { S_LAM, S_V2, S_Y, S_NB, S_INV_RS, S_INV_T2, S_T4C };
    int i_tab, i_slot;
    CTB_DEV void load(const double *p, const JTab *jb_, const JTab *)
    {
        i_tab = jb_->i_brk - 4;
        i_slot = jtab_slot_base(*jb_);
        const double T = p[5];
        pot_slot(i_slot, S_LAM) = p[0]; pot_slot(i_slot, S_V2) = p[1]; pot_slot(i_slot, S_Y) = p[2]; pot_slot(i_slot, S_NB) = p[3];
        // the reference divides by renormScaleSq, T^2 + 1e-100 and 2 pi^2 at every evaluation
        // (GP:250,282,296); the reciprocals are formed once here (<= 1 ulp per factor)
        pot_slot(i_slot, S_INV_RS) = 1.0 / p[4];
        pot_slot(i_slot, S_INV_T2) = 1.0 / (T * T + 1e-100);
        pot_slot(i_slot, S_T4C) = (T * T * T * T) / (2 * CTB_PI * CTB_PI);
    }
    static __device__ __noinline__ double V_s(double phi, int i_tab, int i_slot)
    {
        const TabView t = tab_view(i_tab);
        const double lam = pot_slot(i_slot, S_LAM), v2 = pot_slot(i_slot, S_V2), nb = pot_slot(i_slot, S_NB);
        const double inv_rs = pot_slot(i_slot, S_INV_RS), inv_T2 = pot_slot(i_slot, S_INV_T2);
        const double p2 = phi * phi;
        const double m0 = lam * (3 * p2 - v2), m1 = pot_slot(i_slot, S_Y) * p2;
        // (the two logarithms and the two table look-ups are branch-free: their instruction streams interleave)
        const double l0 = log_pos(fabs(m0 * inv_rs) + 1e-100), l1 = log_pos(fabs(m1 * inv_rs) + 1e-100);
        const double j0 = tab_J(t, m0 * inv_T2), j1 = tab_J(t, m1 * inv_T2);
        double r = .25 * lam * (p2 - v2) * (p2 - v2);
        double y1 = m0 * m0 * (l0 - 1.5);
        y1 += nb * m1 * m1 * (l1 - 1.5);
        r += y1 * (1.0 / (64 * CTB_PI * CTB_PI));
        const double yt = j0 + nb * j1;
        return r + yt * pot_slot(i_slot, S_T4C);
    }
    CTB_DEV double V(double phi) const { return V_s(phi, i_tab, i_slot); }
    CTB_DEV double V_quad(double phi) const { return V(phi); }
};

// General one-field thermal potential (CTB_POT_THERMAL1F, layout in ctbounce.h).  The parameter row stays
// in global memory (L1-resident, 512 B per instance); V is one out-of-line function.
struct PotThermal1F {
    static constexpr bool analytic = false;
    static constexpr int min_blocks = CTB_THERMAL_MIN_BLOCKS;
    static constexpr int nparam = CTB_NPARAM_THERMAL1F;
    static constexpr int nslot = 0;
    const double *p;
    double inv_rs, T2, inv_T2, T4c;
    int nb, nf;
    JTab jb, jf;
    CTB_DEV void load(const double *p_, const JTab *jb_, const JTab *jf_)
    {
        p = p_;
        const double T = p[6];
        inv_rs = 1.0 / p[5];
        T2 = T * T;
        inv_T2 = 1.0 / (T2 + 1e-100);
        T4c = (T2 * T2) / (2 * CTB_PI * CTB_PI);
        nb = min(max((int)p[7], 0), CTB_THERMAL1F_MAX_BOSONS);
        nf = min(max((int)p[8], 0), CTB_THERMAL1F_MAX_FERMIONS);
        jb = *jb_;
        jf = *jf_;
    }
    // V0 / V1 / V1T pieces (GP:240-296) for one field value; p may live in global or shared memory
    struct Parts { double v0, v1, v1t; };
    static CTB_DEV Parts eval(const double *p, int nb, int nf, double inv_rs, double T2, double inv_T2, double T4c,
                              const JTab &jb, const JTab &jf, double phi)
    {
        const double p2 = phi * phi;
        Parts o;
        o.v0 = p[0] + phi * (p[1] + phi * (p[2] + phi * (p[3] + phi * p[4])));
        double y1 = 0.0, yt = 0.0;
        for (int i = 0; i < nb; i++) { // GP:249-251, 285-286
            const double *q = p + 9 + 5 * i;
            const double m2 = q[0] + q[1] * p2 + q[2] * T2, dof = q[3];
            y1 += dof * m2 * m2 * (log_pos(fabs(m2 * inv_rs) + 1e-100) - q[4]);
            yt += dof * jtab_eval<0>(jb, m2 * inv_T2);
        }
        for (int j = 0; j < nf; j++) { // GP:252-255, 287-288
            const double *q = p + 49 + 3 * j;
            const double m2 = q[0] + q[1] * p2, dof = q[2];
            y1 -= dof * m2 * m2 * (log_pos(fabs(m2 * inv_rs) + 1e-100) - 1.5);
            yt += dof * jtab_eval<0>(jf, m2 * inv_T2);
        }
        o.v1 = y1 * (1.0 / (64 * CTB_PI * CTB_PI));
        o.v1t = yt; // in units of T^4 / (2 pi^2): the caller subtracts the radiation constant before scaling
        return o;
    }
    __device__ __noinline__ double V(double phi) const
    {
        const Parts o = eval(p, nb, nf, inv_rs, T2, inv_T2, T4c, jb, jf, phi);
        return o.v0 + o.v1 + o.v1t * T4c;
    }
    CTB_DEV double V_quad(double phi) const { return V(phi); }
};

// Dynamic shared memory of a kernel that evaluates functor Pot on a staged compact table: the table plus the
// functor's per-thread parameter slots (blocks of at most CTB_SLOT_STRIDE threads).
template <class Pot>
inline size_t pot_smem_bytes(int nint)
{
    return Pot::analytic ? 0 : (CTB_TABLE_DOUBLES_COMPACT(nint) + (size_t)Pot::nslot * CTB_SLOT_STRIDE) * sizeof(double);
}
// Launch with the opt-in for more than 48 KB of dynamic shared memory where needed.
template <class... KArgs, class... Args>
inline void launch_k(void (*kernel)(KArgs...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, Args... args)
{
    if (smem > 48 * 1024) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kernel<<<grid, block, smem, st>>>(args...);
}

// ------------------------------------------------------------------------------------------------
// small helpers
CTB_DEV double sgn(double x) { return (double)((x > 0) - (x < 0)); }

// 1/x for normal, finite, non-zero x: MUFU.RCP64H seed + two Newton steps (<= 1 ulp; no special-case
// slow path, unlike the IEEE division sequence).  Used where the operand is a radius or a tolerance.
CTB_DEV double rcp_fast(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    return y;
}

// x^(-1/5) for x in (1e-5, 1]: FP32 seed (ex2/lg2 MUFU, ~1e-6 relative) + two Newton steps on
// y -> y (6 - x y^5) / 5 in FP64 (quadratic: 1e-6 -> 1e-12 -> rounding).  Replaces pow(errmax, -0.2).
CTB_DEV double inv_fifth_root(double x)
{
    double y = (double)exp2f(__fmul_rn(-0.2f, __log2f((float)x)));
    const double mx = CTB_MUL(-0.2, x);
    double y2 = CTB_MUL(y, y);
    double y5 = CTB_MUL(CTB_MUL(y2, y2), y);
    y = CTB_MUL(y, fma(mx, y5, 1.2));
    y2 = CTB_MUL(y, y);
    y5 = CTB_MUL(CTB_MUL(y2, y2), y);
    y = CTB_MUL(y, fma(mx, y5, 1.2));
    return y;
}

// scipy.optimize.brentq (Zeros/brentq.c) restated; F: double -> double.
// err: 0 ok, 1 same sign, 2 NaN function value (scipy raises ValueError), 3 maxiter
template <class F>
CTB_DEV double brentq(F f, double xa, double xb, int &err)
{
    const double xtol = 2e-12, rtol = 8.881784197001252e-16;
    double xpre = xa, xcur = xb;
    double xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis;
    double delta, stry, dpre, dblk;
    err = 0;
    fpre = f(xpre);
    fcur = f(xcur);
    if (isnan(fpre) || isnan(fcur)) { err = 2; return nan(""); }
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) { err = 1; return 0.; }
    for (int i = 0; i < 100; i++) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre; fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = CTB_ADD(xtol, CTB_MUL(rtol, fabs(xcur))) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = CTB_MUL(-fcur, xcur - xpre) / (fcur - fpre);
            } else {
                dpre = (fpre - fcur) / (xpre - xcur);
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = CTB_MUL(-fcur, CTB_SUB(CTB_MUL(fblk, dblk), CTB_MUL(fpre, dpre)))
                       / CTB_MUL(CTB_MUL(dblk, dpre), fblk - fpre);
            }
            double m1 = fabs(spre), m2 = CTB_SUB(CTB_MUL(3.0, fabs(sbis)), delta);
            double mn = (m1 < m2) ? m1 : m2;
            if (2 * fabs(stry) < mn) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = f(xcur);
        if (isnan(fcur)) { err = 2; return nan(""); }
    }
    err = 3;
    return xcur;
}

// scipy.optimize.fminbound (_optimize.py:_minimize_scalar_bounded) restated; returns xf
template <class F>
CTB_DEV double fminbound(F func, double x1, double x2, double xatol)
{
    const double sqrt_eps = 1.4832396974191326e-08; // sqrt(2.2e-16)
    const double golden_mean = 0.3819660112501051;  // 0.5*(3-sqrt(5))
    double a = x1, b = x2;
    double fulc = a + golden_mean * (b - a);
    double nfc = fulc, xf = fulc;
    double rat = 0.0, e = 0.0;
    double x = xf;
    double fx = func(x);
    int num = 1;
    double fu;
    double ffulc = fx, fnfc = fx;
    double xm = 0.5 * (a + b);
    double tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
    double tol2 = 2.0 * tol1;
    while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
        bool golden = true;
        if (fabs(e) > tol1) {
            golden = false;
            double r = (xf - nfc) * (fx - ffulc);
            double q = (xf - fulc) * (fx - fnfc);
            double p = (xf - fulc) * q - (xf - nfc) * r;
            q = 2.0 * (q - r);
            if (q > 0.0) p = -p;
            q = fabs(q);
            r = e;
            e = rat;
            if ((fabs(p) < fabs(0.5 * q * r)) && (p > q * (a - xf)) && (p < q * (b - xf))) {
                rat = (p + 0.0) / q;
                x = xf + rat;
                if (((x - a) < tol2) || ((b - x) < tol2)) {
                    double si = sgn(xm - xf) + ((xm - xf) == 0);
                    rat = tol1 * si;
                }
            } else {
                golden = true;
            }
        }
        if (golden) {
            e = (xf >= xm) ? a - xf : b - xf;
            rat = golden_mean * e;
        }
        double si = sgn(rat) + (rat == 0);
        double ar = fabs(rat);
        x = xf + si * (ar > tol1 ? ar : tol1);
        fu = func(x);
        num += 1;
        if (fu <= fx) {
            if (x >= xf) a = xf; else b = xf;
            fulc = nfc; ffulc = fnfc;
            nfc = xf; fnfc = fx;
            xf = x; fx = fu;
        } else {
            if (x < xf) a = x; else b = x;
            if ((fu <= fnfc) || (nfc == xf)) {
                fulc = nfc; ffulc = fnfc;
                nfc = x; fnfc = fu;
            } else if ((fu <= ffulc) || (fulc == xf) || (fulc == nfc)) {
                fulc = x; ffulc = fu;
            }
        }
        xm = 0.5 * (a + b);
        tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
        tol2 = 2.0 * tol1;
        if (num >= 500) break;
    }
    return xf;
}

// scipy.integrate.simpson on a non-uniform grid, as a STREAM: points are fed in order and only the
// last three are kept.  N (total number of points) must be known up front.   T1D:787 + scipy
struct SimpsonStream {
    // General part (unequal spacing, scipy's formula): (xa, ya) first point of the open pair, (xb, yb) its
    // middle point, (xp, yp) middle point of the previous pair (needed by the even-N end correction).
    double xa, ya, xb, yb, xp, yp;
    double sum;
    int idx, nbasic;
    // Uniform part: on R = linspace(r0, rf, npoints) (T1D:730) consecutive spacings agree to ~1e-13, the
    // unequal-spacing weights collapse to h/3 (1, 4, 2, 4, ..., 4, 1) and the rule becomes two running sums
    // over odd / even global indices (one add per point): see begin_uniform / feed_uniform / finish.
    double s_odd, s_even, y_g0, yb_prev;
    bool uniform_on;
    CTB_DEV void init(int N)
    {
        sum = 0.0; idx = 0;
        nbasic = (N & 1) ? N : N - 1; // even N: composite rule on the first N-1 points
        xa = ya = xb = yb = xp = yp = 0.0;
        s_odd = s_even = y_g0 = yb_prev = 0.0;
        uniform_on = false;
    }
    static CTB_DEV double pair(double x0, double y0, double x1, double y1, double x2, double y2)
    {
        double h0 = x1 - x0, h1 = x2 - x1;
        double hsum = h0 + h1, hprod = h0 * h1;
        if (hprod > 0 && hprod < 1e300) { // regular case: h0/h1, h1/h0, hsum/hprod from one reciprocal
            double rp = rcp_fast(hprod);
            double h0divh1 = h0 * h0 * rp, inv = h1 * h1 * rp, q = hsum * rp;
            return hsum * (1.0 / 6.0) * (y0 * (2.0 - inv) + y1 * (hsum * q) + y2 * (2.0 - h0divh1));
        }
        double h0divh1 = (h1 != 0) ? h0 / h1 : 0.0; // scipy's guarded divisions (repeated abscissae)
        double inv = (h0divh1 != 0) ? 1.0 / h0divh1 : 0.0;
        double q = (hprod != 0) ? hsum / hprod : 0.0;
        return hsum / 6.0 * (y0 * (2.0 - inv) + y1 * (hsum * q) + y2 * (2.0 - h0divh1));
    }
    CTB_DEV void feed(double x, double y)
    {
        if (idx & 1) { // odd index: middle (or, for even N, the very last) point
            xb = x; yb = y;
        } else {
            if (idx >= 2 && idx < nbasic) { sum += pair(xa, ya, xb, yb, x, y); xp = xb; yp = yb; }
            xa = x; ya = y;
        }
        idx++;
    }
    // Call right after feed() of the even-index point g0 from which every further pair lies on the uniform grid.
    CTB_DEV void begin_uniform()
    {
        uniform_on = true;
        s_even = ya; y_g0 = ya; s_odd = 0.0;
    }
    // Point with global index parity `odd` on the uniform part (no x needed).
    CTB_DEV void feed_uniform(double y, bool odd)
    {
        if (odd) { yb_prev = yb; yb = y; s_odd += y; }
        else { s_even += y; ya = y; }
    }
    // x_m2, x_m1, x_last: abscissae of the last three points (only used for even N); h: uniform spacing
    CTB_DEV double finish(int N, double h, double x_m2, double x_m1, double x_last) const
    {
        if (!uniform_on) return finish_general(N);
        const bool even = !(N & 1);
        // even N: the very last point (odd index) is outside the composite rule; it was added to s_odd
        const double so = even ? s_odd - yb : s_odd;
        const double total = sum + (h * (1.0 / 3.0)) * (4.0 * so + 2.0 * s_even - y_g0 - ya);
        if (!even) return total;
        return total + end_correction(x_m2, yb_prev, x_m1, ya, x_last, yb);
    }
    static CTB_DEV double end_correction(double x0, double y0, double x1, double y1, double x2, double y2)
    {
        // Cartwright's last-interval correction (scipy.integrate.simpson, even N)
        double h0 = x1 - x0, h1 = x2 - x1;
        double num, den, al, be, et;
        num = 2 * h1 * h1 + 3 * h0 * h1; den = 6 * (h1 + h0);
        al = (den != 0) ? num / den : 0.0;
        num = h1 * h1 + 3.0 * h0 * h1; den = 6 * h0;
        be = (den != 0) ? num / den : 0.0;
        num = 1 * h1 * h1 * h1; den = 6 * h0 * (h0 + h1);
        et = (den != 0) ? num / den : 0.0;
        return al * y2 + be * y1 - et * y0;
    }
    CTB_DEV double finish_general(int N) const
    {
        if (N & 1) return sum;
        if (N == 2) return 0.5 * (xb - xa) * (yb + ya);
        // points N-3, N-2, N-1 = (xp, yp), (xa, ya), (xb, yb)
        return sum + end_correction(xp, yp, xa, ya, xb, yb);
    }
};

// ------------------------------------------------------------------------------------------------
// the instance solver
struct Knobs {
    double xtol, phitol, thinCutoff, rmin, rmax, phi_eps, xguess;
    int npoints, max_interior_pts;
    double alpha_real; // the friction coefficient when the kernel is the ALPHA = CTB_ALPHA_REAL instantiation
};

struct ProfileSink { // optional Profile1D dump
    double *R, *Phi, *dPhi;
};

enum { CT_CONVERGED = 0, CT_UNDERSHOOT = 1, CT_OVERSHOOT = 2 };

// Geometry factors of findAction (T1D:782-790) for d = ALPHA + 1 dimensions: area of the unit ALPHA-sphere
// 2 pi^(d/2) / Gamma(d/2), volume of the unit d-ball pi^(d/2) / Gamma(d/2 + 1).  ALPHA in {1, 2, 3, 4}: constants;
// ALPHA = CTB_ALPHA_REAL (0): any real alpha >= 0, passed at run time.
#define CTB_ALPHA_REAL 0
template <int ALPHA> CTB_DEV double area_fac(double alpha)
{
    if (ALPHA == CTB_ALPHA_REAL) return 2.0 * pow(CTB_PI, 0.5 * (alpha + 1.0)) / tgamma(0.5 * (alpha + 1.0));
    return (ALPHA == 1) ? 6.283185307179586 : (ALPHA == 2) ? 12.566370614359172 : (ALPHA == 3) ? 19.739208802178716
                                                                                               : 26.318945069571623;
}
template <int ALPHA> CTB_DEV double vol_fac(double alpha)
{
    if (ALPHA == CTB_ALPHA_REAL) return pow(CTB_PI, 0.5 * (alpha + 1.0)) / tgamma(0.5 * (alpha + 1.0) + 1.0);
    return (ALPHA == 1) ? 3.141592653589793 : (ALPHA == 2) ? 4.1887902047863905 : (ALPHA == 3) ? 4.934802200544679
                                                                                               : 5.263789013914325;
}
template <int ALPHA> CTB_DEV double pow_alpha(double r, double alpha) // r^ALPHA
{
    if (ALPHA == CTB_ALPHA_REAL) return pow(r, alpha);
    return (ALPHA == 1) ? r : (ALPHA == 2) ? r * r : (ALPHA == 3) ? r * r * r : (r * r) * (r * r);
}
template <int ALPHA> CTB_DEV double pow_dim(double r, double alpha) // r^(ALPHA+1)
{
    if (ALPHA == CTB_ALPHA_REAL) return pow(r, alpha + 1.0);
    return (ALPHA == 1) ? r * r : (ALPHA == 2) ? r * r * r : (ALPHA == 3) ? (r * r) * (r * r) : (r * r) * (r * r) * r;
}

// T1D:294-373  phi(r), phi'(r) of the linearised equation of motion around phi0.
// nu = (ALPHA-1)/2: closed forms for nu = 1/2, I0/I1/J0/J1 + recurrences for nu = 1.
// ONE out-of-line copy per ALPHA: the transcendental bodies (exp, sinh, sincos, Bessel) are large and
// every call site sharing them keeps the kernel's hot code inside the instruction cache.
// (beta = sqrt|V''| and ratio = V'/V'' are constant over the ~16 evaluations of a root find and the up
// to 250 interior points: the caller computes them once, see exact_coef)
struct ExactCoef {
    double phi0, dv, d2v, beta, ratio;
    double nu; // (alpha - 1) / 2, read by the ALPHA = CTB_ALPHA_REAL instantiation only
};
CTB_DEV ExactCoef exact_coef(double phi0, double dv, double d2v, double nu = 0.0)
{
    return ExactCoef{phi0, dv, d2v, sqrt(fabs(d2v)), dv / d2v, nu};
}

// General order nu >= -1/2 (any real alpha >= 0, T1D:336-372): with Lambda_nu(z) = Gamma(nu+1) (z/2)^-nu I_nu(z) =
// 0F1(; nu+1; z^2/4) (J_nu: 0F1(; nu+1; -z^2/4)) the reference's expressions are phi - phi0 = (Lambda_nu - 1) V'/V'' and
// phi' = beta Lambda_nu'(z) V'/V'' with Lambda_nu' = +-z / (2 (nu+1)) Lambda_{nu+1}.  hyp0f1m1(b, z, modified) returns
// 0F1(; b; +-z^2/4) - 1:
//   ascending series (all terms positive for I: no cancellation; for J it loses ~e^z eps, so only below z = 12);
//   I, z >= 25 and z >= 2 nu^2 + 10: Hankel's asymptotic expansion e^z / sqrt(2 pi z) sum (-1)^k a_k(nu) / z^k;
//   J, z >= 12: sqrt(2 / pi z) (P cos chi - Q sin chi), chi = z - (nu/2 + 1/4) pi.
// Truncated at the smallest term.  Checked against 50-digit mpmath values for nu in [-1/2, 7/2] (alpha in [0, 8], the
// range the API accepts): I to 2e-15 relative, J (value and derivative) to 4e-12 absolute.  exp overflow (z > 709) gives
// inf like scipy's iv.
CTB_DEV double hyp0f1m1(double b, double z, bool modified)
{
    const double nu = b - 1.0;
    if (modified ? (z < 25.0 || z < 2.0 * nu * nu + 10.0) : (z < 12.0 + 2.0 * fabs(nu))) {
        const double t = (modified ? 0.25 : -0.25) * z * z;
        double term = 1.0, sum = 0.0;
        for (int k = 1; k < 2000; k++) {
            term *= t / ((double)k * (b + (double)(k - 1)));
            sum += term;
            if (fabs(term) <= 1e-17 * fabs(sum) && (double)k > 0.5 * z) break;
        }
        return sum;
    }
    const double mu = 4.0 * nu * nu;
    const double pref = tgamma(b) * pow(0.5 * z, -nu); // Gamma(nu+1) (z/2)^-nu
    if (modified) {
        double term = 1.0, sum = 1.0, last = 1.0;
        for (int k = 1; k < 200; k++) {
            const double q = (double)(2 * k - 1);
            term *= -(mu - q * q) / ((double)k * 8.0 * z);
            if (fabs(term) > last) break; // the series is asymptotic: stop at the smallest term
            sum += term;
            last = fabs(term);
            if (last < 1e-18) break;
        }
        const double amp = sum * rsqrt(2.0 * CTB_PI * z) * pref;
        if (z < 700.0) return exp(z) * amp - 1.0;
        const double e2 = exp(0.5 * z);
        return (e2 * amp) * e2 - 1.0;
    }
    double P = 1.0, Q = 0.0, term = 1.0, last = 1.0;
    for (int k = 1; k < 200; k++) {
        const double q = (double)(2 * k - 1);
        term *= (mu - q * q) / ((double)k * 8.0 * z);
        if (fabs(term) > last) break;
        last = fabs(term);
        // a_k / z^k enters Q for odd k and P for even k, with signs (+ - - + + - - ...) -> (-1)^(k/2)
        const double sg = ((k >> 1) & 1) ? -1.0 : 1.0;
        if (k & 1) Q += sg * term; else P += sg * term;
        if (last < 1e-18) break;
    }
    double sn, cs;
    sincos(z - (0.5 * nu + 0.25) * CTB_PI, &sn, &cs);
    return sqrt(2.0 / (CTB_PI * z)) * (P * cs - Q * sn) * pref - 1.0;
}

// I0(z), I1(z) for z >= 0 together (the alpha = 1, 3 start conditions evaluate both at every call: the library's
// cyl_bessel_i0 + cyl_bessel_i1 are two independent ~120-instruction routines, 11.5 % of all C3 instructions).
//   z <= 7.75: the ascending series in t = z^2 / 4, 21 all-positive terms each (no cancellation; <= 4 ulp), summed from
//              k = 1 so that I0 - 1 and (2 / z) I1 - 1 -- what the start conditions actually need -- come without the
//              cancellation the subtraction has at small z;
//   z >  7.75: e^z / sqrt(2 pi z) times degree-22 polynomials in 1 / z (Chebyshev fits of sqrt(2 pi z) e^-z I_nu(z) on
//              0 < 1/z <= 1/7.75 made with 50-digit mpmath values; 2.3e-16 relative), ONE exponential and ONE reciprocal
//              square root for both; e^z is split for z > 700 so that the result overflows where I_nu does, not before.
// Checked against mpmath on 1e-3 <= z <= 1e3: I0, I1 <= 5 ulp.
struct BesselI01 { double i0, i1, i0m1, i1hm1; }; // i0m1 = I0 - 1, i1hm1 = (2 / z) I1 - 1
CTB_DEV BesselI01 bessel_i01(double z)
{
    BesselI01 r;
    if (z <= 7.75) {
        const double t = 0.25 * z * z, t2 = t * t;
        // (even / odd halves: four independent Horner chains of ten terms instead of two of twenty)
        double p0e = 6.75788438580422547e-35;
        p0e = fma(p0e, t2, 7.90429189301205402e-30);
        p0e = fma(p0e, t2, 5.84791131412603850e-25);
        p0e = fma(p0e, t2, 2.57892888952958276e-20);
        p0e = fma(p0e, t2, 6.27608134555919329e-16);
        p0e = fma(p0e, t2, 7.59405842812662392e-12);
        p0e = fma(p0e, t2, 3.93675988914084175e-08);
        p0e = fma(p0e, t2, 6.94444444444444444e-05);
        p0e = fma(p0e, t2, 2.77777777777777762e-02);
        p0e = fma(p0e, t2, 1.00000000000000000e+00);
        double p0o = 1.68947109645105643e-37;
        p0o = fma(p0o, t2, 2.43959626327532528e-32);
        p0o = fma(p0o, t2, 2.28434035708048379e-27);
        p0o = fma(p0o, t2, 1.31578004567835862e-22);
        p0o = fma(p0o, t2, 4.35838982330499500e-18);
        p0o = fma(p0o, t2, 7.59405842812662337e-14);
        p0o = fma(p0o, t2, 6.15118732678256523e-10);
        p0o = fma(p0o, t2, 1.92901234567901239e-06);
        p0o = fma(p0o, t2, 1.73611111111111101e-03);
        p0o = fma(p0o, t2, 2.50000000000000000e-01);
        const double p0 = fma(p0o, t, p0e);
        double p1e = 3.37894219290211260e-36;
        p1e = fma(p1e, t2, 4.39127327389558567e-31);
        p1e = fma(p1e, t2, 3.65494457132877406e-26);
        p1e = fma(p1e, t2, 1.84209206394970202e-21);
        p1e = fma(p1e, t2, 5.23006778796599400e-17);
        p1e = fma(p1e, t2, 7.59405842812662312e-13);
        p1e = fma(p1e, t2, 4.92094986142605219e-09);
        p1e = fma(p1e, t2, 1.15740740740740735e-05);
        p1e = fma(p1e, t2, 6.94444444444444406e-03);
        p1e = fma(p1e, t2, 5.00000000000000000e-01);
        double p1o = 8.04510045929074426e-39;
        p1o = fma(p1o, t2, 1.28399803330280280e-33);
        p1o = fma(p1o, t2, 1.34372962181204914e-28);
        p1o = fma(p1o, t2, 8.77186697118905747e-24);
        p1o = fma(p1o, t2, 3.35260755638845791e-19);
        p1o = fma(p1o, t2, 6.90368948011511223e-15);
        p1o = fma(p1o, t2, 6.83465258531396137e-11);
        p1o = fma(p1o, t2, 2.75573192239858883e-07);
        p1o = fma(p1o, t2, 3.47222222222222235e-04);
        p1o = fma(p1o, t2, 8.33333333333333287e-02);
        const double p1 = fma(p1o, t, p1e);
        r.i0m1 = p0 * t;
        r.i1hm1 = p1 * t;
        r.i0 = r.i0m1 + 1.0;
        r.i1 = (0.5 * z) * (r.i1hm1 + 1.0);
    } else {
        const double u = rcp_fast(z), u2 = u * u;
        double g0e = 5.10439856463674160e+16;
        g0e = fma(g0e, u2, 4.30710600204572800e+16);
        g0e = fma(g0e, u2, 4.26278082774414200e+15);
        g0e = fma(g0e, u2, 1.12849228600115922e+14);
        g0e = fma(g0e, u2, 1.00993467373931482e+12);
        g0e = fma(g0e, u2, 3.30983866591941309e+09);
        g0e = fma(g0e, u2, 3.96114130742453830e+06);
        g0e = fma(g0e, u2, 1.60629422238111147e+03);
        g0e = fma(g0e, u2, 7.54450335052397558e-01);
        g0e = fma(g0e, u2, 1.12156100369509099e-01);
        g0e = fma(g0e, u2, 7.03125000067696959e-02);
        g0e = fma(g0e, u2, 1.00000000000000000e+00);
        double g0o = -6.90847115994374720e+16;
        g0o = fma(g0o, u2, -1.63983679029149040e+16);
        g0o = fma(g0o, u2, -8.01728730945816125e+14);
        g0o = fma(g0o, u2, -1.21369303670207148e+13);
        g0o = fma(g0o, u2, -6.54529112951727753e+10);
        g0o = fma(g0o, u2, -1.30254789946553946e+08);
        g0o = fma(g0o, u2, -9.19298952676435292e+04);
        g0o = fma(g0o, u2, -1.86106739466121525e+01);
        g0o = fma(g0o, u2, 2.26021504737211332e-01);
        g0o = fma(g0o, u2, 7.32421795032049128e-02);
        g0o = fma(g0o, u2, 1.24999999999999348e-01);
        const double g0 = fma(g0o, u, g0e);
        double g1e = -5.39685914928520960e+16;
        g1e = fma(g1e, u2, -4.57949971183415680e+16);
        g1e = fma(g1e, u2, -4.56594665224959400e+15);
        g1e = fma(g1e, u2, -1.22096213504144156e+14);
        g1e = fma(g1e, u2, -1.10837511090765234e+12);
        g1e = fma(g1e, u2, -3.71040044690264654e+09);
        g1e = fma(g1e, u2, -4.59308217909604032e+06);
        g1e = fma(g1e, u2, -1.97488628412734852e+03);
        g1e = fma(g1e, u2, -9.27920846064049454e-01);
        g1e = fma(g1e, u2, -1.44202791873929648e-01);
        g1e = fma(g1e, u2, -1.17187500026180808e-01);
        g1e = fma(g1e, u2, 1.00000000000000000e+00);
        double g1o = 7.32348604794205440e+16;
        g1o = fma(g1o, u2, 1.74952333397986120e+16);
        g1o = fma(g1o, u2, 8.62720432009044125e+14);
        g1o = fma(g1o, u2, 1.32166579033533594e+13);
        g1o = fma(g1o, u2, 7.25176396518202820e+10);
        g1o = fma(g1o, u2, 1.48193099420835495e+08);
        g1o = fma(g1o, u2, 1.09296831122297372e+05);
        g1o = fma(g1o, u2, 2.42323657130146302e+01);
        g1o = fma(g1o, u2, -2.75910766256493467e-01);
        g1o = fma(g1o, u2, -1.02539043477997735e-01);
        g1o = fma(g1o, u2, -3.74999999999987566e-01);
        const double g1 = fma(g1o, u, g1e);
        double rs;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(rs) : "d"(z));
        rs = rs * fma(-0.5 * z * rs, rs, 1.5); // two Newton steps on 1 / sqrt(z)
        rs = rs * fma(-0.5 * z * rs, rs, 1.5);
        rs *= 0.39894228040143268; // 1 / sqrt(2 pi)
        if (z < 700.0) {
            const double E = exp(z) * rs;
            r.i0 = g0 * E; r.i1 = g1 * E;
        } else { // (also z = inf: exp -> inf, rs -> 0 would give NaN without the split order below)
            const double e2 = exp(0.5 * z);
            r.i0 = (g0 * rs * e2) * e2; r.i1 = (g1 * rs * e2) * e2;
            if (z > 1e300) r.i0 = r.i1 = INFINITY;
        }
        r.i0m1 = r.i0 - 1.0;
        r.i1hm1 = (2.0 * u) * r.i1 - 1.0;
    }
    return r;
}

template <int ALPHA>
__device__ __noinline__ double2 exact_solution(double r, const ExactCoef c)
{
    double phi, dphi;
    const double phi0 = c.phi0, dv = c.dv, d2v = c.d2v, beta = c.beta;
    double z = beta * r;
    if (ALPHA == CTB_ALPHA_REAL) { // any real alpha >= 0: general-order Bessel functions (see hyp0f1m1)
        const double nu = c.nu;
        if (z < 1e-2) {
            const double s = (d2v > 0) ? 1.0 : -1.0;
            const double g = tgamma(nu + 1.0);
            const double h2 = (0.5 * z) * (0.5 * z);
            double p = 0.0, dp = 0.0, hk = 1.0, sk = 1.0;
            for (int k = 1; k <= 3; k++) {
                sk *= s;
                const double q = hk * sk / (tgamma((double)k + 1.0) * tgamma((double)k + 1.0 + nu));
                p += q; dp += q * (2.0 * k);
                hk *= h2;
            }
            return make_double2(p * (0.25 * g * (r * r) * dv * s) + phi0, dp * (0.25 * g * r * dv * s));
        }
        const bool modified = d2v > 0;
        const double fm1 = hyp0f1m1(nu + 1.0, z, modified);
        const double fp = (modified ? 0.5 : -0.5) * z / (nu + 1.0) * (hyp0f1m1(nu + 2.0, z, modified) + 1.0);
        return make_double2(fm1 * c.ratio + phi0, (beta * fp) * c.ratio);
    }
    if (z < 1e-2) {
        const double s = (d2v > 0) ? 1.0 : -1.0;
        // Gamma(k+1) Gamma(k+1+nu), k = 1..3 ; Gamma(nu+1)
        const double G1 = (ALPHA == 1) ? 1.0 : (ALPHA == 2) ? 1.3293403881791370 : (ALPHA == 3) ? 2.0 : 3.3233509704478426;
        const double G2 = (ALPHA == 1) ? 4.0 : (ALPHA == 2) ? 2.0 * 3.3233509704478426 : (ALPHA == 3) ? 12.0
                                                                                                   : 2.0 * 11.631728396567448;
        const double G3 = (ALPHA == 1) ? 36.0 : (ALPHA == 2) ? 6.0 * 11.631728396567448 : (ALPHA == 3) ? 144.0
                                                                                                     : 6.0 * 52.34277778455352;
        const double g = (ALPHA == 1) ? 1.0 : (ALPHA == 2) ? 0.88622692545275801 : (ALPHA == 3) ? 1.0 : 1.3293403881791370;
        double h2 = (0.5 * z) * (0.5 * z);
        // (reciprocals of the constants are folded at compile time: no IEEE division sequence on this path)
        double t1 = s * (1.0 / G1), t2 = h2 * (1.0 / G2), t3 = h2 * h2 * s * (1.0 / G3);
        double p = t1 + t2 + t3;
        double dp = t1 * 2 + t2 * 4 + t3 * 6;
        phi = p * (0.25 * g * (r * r) * dv * s) + phi0;
        dphi = dp * (0.25 * g * r * dv * s);
        return make_double2(phi, dphi);
    }
    const bool modified = d2v > 0;
    const double ratio = c.ratio;
    if (ALPHA == 2) {
        // Gamma(3/2) (z/2)^(-1/2) I_{1/2}(z) = sinh(z)/z ; J: sin(z)/z
        double sh, ch;
        if (modified) {
            if (z < 1.0) { sh = sinh(z); ch = cosh(z); }
            else if (z < 700.0) { // one exp for both; no cancellation for z >= 1
                double ez = exp(z), emz = rcp_fast(ez);
                sh = 0.5 * (ez - emz); ch = 0.5 * (ez + emz);
            } else { double e2 = exp(0.5 * z); sh = (0.5 * e2) * e2; ch = sh; } // overflows with I_nu
        } else {
            sincos(z, &sh, &ch);
        }
        // (1 / z by the Newton reciprocal, 1 / r = beta / z: no IEEE division sequence on this path)
        const double inv_z = rcp_fast(z), shz = sh * inv_z;
        phi = (shz - 1.0) * ratio + phi0;
        dphi = ((ch - shz) * (beta * inv_z)) * ratio;
    } else if (ALPHA == 1) {
        // nu = 0: I_0(z) - 1 and (I_{-1} + I_1)/2 = I_1 ; J: J_0(z) - 1 and (J_{-1} - J_1)/2 = -J_1
        if (modified) {
            const BesselI01 b = bessel_i01(z);
            phi = b.i0m1 * ratio + phi0;
            dphi = (beta * b.i1) * ratio;
        } else {
            phi = (j0(z) - 1.0) * ratio + phi0;
            dphi = (-beta * j1(z)) * ratio;
        }
    } else if (ALPHA == 4) {
        // nu = 3/2: f(z) = Gamma(5/2) (z/2)^(-3/2) I_{3/2}(z) = 3 (cosh z - sinh z / z) / z^2  (J: sin z / z - cos z),
        // phi - phi0 = (f - 1) V'/V'', phi' = beta f'(z) V'/V''.  The closed forms cancel like z^-4 for small z:
        // below z = 1 the ascending series sum_k (+-z^2/4)^k Gamma(5/2) / (k! Gamma(k + 5/2)) is used instead.
        double fm1, fp;
        if (z < 1.0) {
            const double q = (modified ? 0.25 : -0.25) * z * z;
            double t = 1.0, sf = 0.0, sd = 0.0;
#pragma unroll
            for (int k = 1; k <= 11; k++) {
                t *= q / ((double)k * ((double)k + 1.5));
                sf += t;
                sd += (double)k * t;
            }
            fm1 = sf;
            fp = 2.0 * sd / z;
        } else {
            const double iz = 1.0 / z, iz2 = iz * iz;
            double sh, ch;
            if (modified && z >= 700.0) { // sinh = cosh = e^z / 2 (overflows with I_nu: phi -> inf, not NaN)
                const double e2 = exp(0.5 * z), E = (0.5 * e2) * e2;
                fm1 = (3.0 * iz2 * (1.0 - iz)) * E - 1.0;
                fp = (3.0 * iz2 * (1.0 - 3.0 * iz + 3.0 * iz2)) * E;
            } else if (modified) {
                double ez = exp(z), emz = rcp_fast(ez);
                sh = 0.5 * (ez - emz); ch = 0.5 * (ez + emz);
                fm1 = 3.0 * (ch - sh * iz) * iz2 - 1.0;
                fp = 3.0 * iz2 * (sh - 3.0 * ch * iz + 3.0 * sh * iz2);
            } else {
                sincos(z, &sh, &ch);
                fm1 = 3.0 * (sh * iz - ch) * iz2 - 1.0;
                fp = 3.0 * iz2 * (sh + 3.0 * ch * iz - 3.0 * sh * iz2);
            }
        }
        phi = fm1 * ratio + phi0;
        dphi = (beta * fp) * ratio;
    } else {
        // nu = 1: f = (z/2)^-1 I_1(z), f' = (z/2)^-1 I_2(z) with I_2 = I_0 - (2/z) I_1 (J: f' = -(z/2)^-1 J_2,
        // J_2 = (2/z) J_1 - J_0), i.e. the reference's -nu (..) I_nu / r + (..) beta (I_{nu-1} +- I_{nu+1}) / 2 collected
        const double hp = 2.0 * rcp_fast(z); // (z/2)^-1, Gamma(2) = 1
        if (modified) { // f - 1 = (2/z) I1 - 1 and I2 = I0 - (2/z) I1 = (I0 - 1) - ((2/z) I1 - 1), both free of cancellation
            const BesselI01 b = bessel_i01(z);
            phi = b.i1hm1 * ratio + phi0;
            dphi = ((beta * hp) * (b.i0m1 - b.i1hm1)) * ratio;
        } else {
            const double B0 = j0(z), B1 = j1(z);
            phi = (hp * B1 - 1.0) * ratio + phi0;
            dphi = ((beta * hp) * (B0 - hp * B1)) * ratio;
        }
    }
    return make_double2(phi, dphi);
}

#define CTB_ALPHA_WALL (-1) // WallWithConstFriction (T1D:829-999): friction term F phi' instead of (alpha / r) phi'

template <class Pot, int ALPHA>
struct Sfi {
    Pot pot;
    double phi_abs, phi_meta, phi_eps, inv_12eps;
    double friction; // used by ALPHA == CTB_ALPHA_WALL only
    double alpha_rt; // used by ALPHA == CTB_ALPHA_REAL only: the real friction coefficient alpha
    double inv_r_end; // 1 / (t + dt) of the last Cash-Karp attempt
    int n_rkck;
    CTB_DEV void set_phi_eps(double e) { phi_eps = e; inv_12eps = 1.0 / (12. * e); }

    // T1D:151-161, 191-201 (5-point stencils) or the functor's analytic forms
    CTB_DEV double dV(double phi) const
    {
        if constexpr (Pot::analytic) {
            return pot.dV(phi);
        } else {
            double eps = phi_eps;
            const double va = pot.V(CTB_SUB(phi, CTB_MUL(2.0, eps))), vb = pot.V(CTB_SUB(phi, eps));
            const double vc = pot.V(CTB_ADD(phi, eps)), vd = pot.V(CTB_ADD(phi, CTB_MUL(2.0, eps)));
            return CTB_MUL(CTB_SUB(CTB_ADD(CTB_SUB(va, CTB_MUL(8.0, vb)), CTB_MUL(8.0, vc)), vd), inv_12eps);
        }
    }
    CTB_DEV double d2V(double phi) const
    {
        if constexpr (Pot::analytic) {
            return pot.d2V(phi);
        } else {
            double eps = phi_eps;
            const double va = pot.V(CTB_SUB(phi, CTB_MUL(2.0, eps))), vb = pot.V(CTB_SUB(phi, eps)), v0 = pot.V(phi);
            const double vc = pot.V(CTB_ADD(phi, eps)), vd = pot.V(CTB_ADD(phi, CTB_MUL(2.0, eps)));
            double acc = CTB_ADD(-va, CTB_MUL(16.0, vb));
            acc = CTB_SUB(acc, CTB_MUL(30.0, v0));
            acc = CTB_ADD(acc, CTB_MUL(16.0, vc));
            acc = CTB_SUB(acc, vd);
            return acc / CTB_MUL(CTB_MUL(12., eps), eps);
        }
    }
    // T1D:163-189
    CTB_DEV double dV_from_absMin(double delta_phi) const
    {
        double phi = phi_abs + delta_phi;
        double dv = dV(phi);
        if (phi_eps > 0) {
            double dv_ = CTB_MUL(d2V(phi), delta_phi);
            double q = delta_phi / phi_eps;
            double blend = exp(-CTB_MUL(q, q));
            dv = CTB_ADD(CTB_MUL(dv_, blend), CTB_MUL(dv, 1 - blend));
        }
        return dv;
    }

    // The reference's own root find (T1D:428-434), one out-of-line copy for its two (rare) callers
    __device__ __noinline__ double start_radius_brentq(const ExactCoef ec, double pa, double ac, double lo, double hi,
                                                      int &err) const
    {
        return brentq([&](double r_) { return fabs(exact_solution<ALPHA>(r_, ec).x - pa) - ac; }, lo, hi, err);
    }

    // Root of |phi(r) - phi_abs| = ac on [lo, hi] for the monotone (V'' > 0) exact solution.  err as brentq's.
    CTB_DEV double start_radius_newton(const ExactCoef &ec, double pa, double ac, double lo, double hi, int &err) const
    {
        err = 0;
        const double flo = fabs(exact_solution<ALPHA>(lo, ec).x - pa) - ac;
        double2 e = exact_solution<ALPHA>(hi, ec);
        const double fhi = fabs(e.x - pa) - ac;
        if (isnan(flo) || isnan(fhi)) { err = 2; return nan(""); }
        if (flo == 0) return lo;
        if (fhi == 0) return hi;
        if (signbit(flo) == signbit(fhi)) { err = 1; return 0.; }
        if (flo > 0) return start_radius_brentq(ec, pa, ac, lo, hi, err);
        const double inv_ac = 1.0 / ac;
        double x = hi;
        for (int it = 0; it < 80; it++) {
            const double d = e.x - pa, ad = fabs(d);
            double xn = nan("");
            if (ad < 1.7976931348623157e308 && e.y != 0 && fabs(e.y) < 1.7976931348623157e308) {
                if (ad == ac) return x;
                xn = x - log(ad * inv_ac) * d * rcp_fast(e.y); // Newton on log(|phi - phi_abs| / ac): d/dr = phi' / (phi - phi_abs)
            }
            if (fabs(xn - x) <= 4e-15 * x) return xn; // converged (NaN compares false); may sit a rounding outside [lo, hi]
            if (ad >= ac) hi = x; else lo = x; // (also taken for an overflowed value at x: the root lies below)
            if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
            x = xn;
            if ((hi - lo) <= 4e-16 * hi) break;
            e = exact_solution<ALPHA>(x, ec);
            if (isnan(e.x)) { err = 2; return nan(""); }
        }
        return x;
    }

    // The same root without the reference's x10 bracket search (T1D:418-424), for V'' > 0 and a root well inside the
    // finite range: |phi - phi_abs| = |delta_phi0| + |V'/V''| (S(z) - 1) with S(z) = Gamma(nu+1) (z/2)^-nu I_nu(z), z = beta r,
    // growing monotonically from 1 like 1 + z^2 / (2 (alpha + 1)) and then like C e^z z^(-alpha/2); so S(z*) = K fixes z*
    // up to a Newton polish, started from whichever asymptotic form applies.  ok = false (caller falls back to the
    // bracketing path, which reproduces the reference's statuses in the overflow zone) whenever anything is out of the
    // ordinary: K not in (1, 1e250), non-finite coefficients, no convergence in 24 evaluations.
    CTB_DEV double start_radius_direct(const ExactCoef &ec, double delta_phi0, double pa, double ac, double rmin, bool &ok) const
    {
        ok = false;
        const double ar = fabs(ec.ratio), K = 1.0 + (ac - fabs(delta_phi0)) / ar;
        if (!(ar > 0.0 && K > 1.0 && K < 1e250 && ec.beta > 0.0 && ec.beta < 1e300)) return 0.0;
        const double invC = (ALPHA == 1) ? 2.5066282746310002 : (ALPHA == 2) ? 2.0 : (ALPHA == 3) ? 1.2533141373155001
                                                                                                : 0.66666666666666663;
        double zg;
        if (K < 2.0) zg = sqrt(2.0 * (ALPHA + 1) * (K - 1.0));
        else {
            const double L = log(K * invC);
            zg = fma(0.5 * ALPHA, log(fma(0.5 * ALPHA, log(L + 1.0), L)), L);
        }
        const double inv_ac = 1.0 / ac;
        double lo = rmin, hi = INFINITY, x = fmax(zg / ec.beta, rmin);
        for (int it = 0; it < 24; it++) {
            const double2 e = exact_solution<ALPHA>(x, ec);
            const double d = e.x - pa, ad = fabs(d);
            if (!(ad < 1.7976931348623157e308) || !(fabs(e.y) < 1.7976931348623157e308) || e.y == 0) return 0.0;
            if (ad == ac) { ok = true; return x; }
            double xn = x - log(ad * inv_ac) * d * rcp_fast(e.y);
            if (fabs(xn - x) <= 4e-15 * x) { ok = xn >= rmin; return xn; }
            if (ad >= ac) hi = x; else lo = x;
            if (!(xn > lo && xn < hi)) xn = (hi < INFINITY) ? 0.5 * (lo + hi) : 2.0 * x;
            x = xn;
        }
        return 0.0;
    }

    // T1D:377-434, given dV_from_absMin(delta_phi0) and d2V(phi0); returns 0 or a CTB_ST_* error
    CTB_DEV int initial_conditions(double delta_phi0, double dv, double d2v, double rmin, double cutoff, double &r0,
                                   double &phi_r0, double &dphi_r0) const
    {
        const ExactCoef ec = exact_coef(phi_abs + delta_phi0, dv, d2v, ALPHA == CTB_ALPHA_REAL ? 0.5 * (alpha_rt - 1.0) : 0.0);
        double2 e = exact_solution<ALPHA>(rmin, ec);
        phi_r0 = e.x; dphi_r0 = e.y;
        r0 = rmin;
        if (fabs(phi_r0 - phi_abs) > fabs(cutoff)) return 0;
        if (isnan(dphi_r0) || isnan(delta_phi0) || sgn(dphi_r0) != sgn(delta_phi0)) return 0;
        const double pa = phi_abs, ac = fabs(cutoff);
        // direct solve (no x10 bracket search) where the root is safely inside the finite range.  Only where an
        // evaluation of the exact solution is expensive (I_0 / I_1, alpha = 1, 3): five evaluations saved outweigh the
        // extra code; with the exp-based closed forms (alpha = 2, 4) the larger kernel loses more in the instruction
        // cache than the search costs (measured: C3 -6 %, C2 +16 %).
        if ((ALPHA == 1 || ALPHA == 3) && d2v > 0) {
            bool ok;
            const double rd = start_radius_direct(ec, delta_phi0, pa, ac, rmin, ok);
            if (ok) {
                r0 = rd;
                e = exact_solution<ALPHA>(r0, ec);
                phi_r0 = e.x; dphi_r0 = e.y;
                return 0;
            }
        }
        double r = rmin, rlast = rmin;
        while (isfinite(r) && r > 0) { // (r > 0: a zero start radius would never leave this loop)
            rlast = r;
            r *= 10;
            double p = exact_solution<ALPHA>(r, ec).x;
            if (fabs(p - phi_abs) > fabs(cutoff)) break;
        }
        int err;
        double rr;
        if (d2v > 0) {
            // V'' > 0 (always the case close to the true vacuum, i.e. for every thin-wall shot): |phi(r) - phi_abs|
            // grows monotonically like sinh(beta r) / (beta r) or I_nu, so its logarithm is almost linear in r and a
            // bracketed Newton iteration on log(|phi - phi_abs| / cutoff) -- phi' comes with phi for free -- converges
            // in 3-5 evaluations where brentq needs ~16.  The end-point tests are brentq's own (same statuses).
            rr = start_radius_newton(ec, pa, ac, rlast, r, err);
        } else {
            rr = start_radius_brentq(ec, pa, ac, rlast, r, err);
        }
        if (err == 2) return CTB_ST_NONFINITE;
        if (err == 1) return CTB_ST_BRENT_SIGN;
        r0 = rr;
        e = exact_solution<ALPHA>(r0, ec);
        phi_r0 = e.x; dphi_r0 = e.y;
        return 0;
    }

    // T1D:436-440
    CTB_DEV void eom(double y0, double y1, double r, double &f0, double &f1) const
    {
        eom_inv(y0, y1, (ALPHA == CTB_ALPHA_WALL) ? 0.0 : rcp_fast(r), f0, f1);
    }
    CTB_DEV void eom_inv(double y0, double y1, double inv_r, double &f0, double &f1) const
    {
        f0 = y1;
        if constexpr (ALPHA == CTB_ALPHA_WALL) f1 = fma(-friction, y1, dV(y0)); // T1D:912-916
        else f1 = fma(-CTB_MUL(ALPHA == CTB_ALPHA_REAL ? alpha_rt : (double)ALPHA, y1), inv_r, dV(y0));
    }

    // HF:204-236 Cash-Karp step
    CTB_DEV void rkck(double y0, double y1, double d0, double d1, double t, double dt, double &dy0, double &dy1,
                      double &e0, double &e1)
    {
        const double a2 = 0.2, a3 = 0.3, a4 = 0.6, a5 = 1.0, a6 = 0.875, b21 = 0.2;
        const double b31 = 3.0 / 40.0, b32 = 9.0 / 40.0, b41 = 0.3, b42 = -0.9, b43 = 1.2;
        const double b51 = -11.0 / 54.0, b52 = 2.5, b53 = -70.0 / 27.0, b54 = 35.0 / 27.0;
        const double b61 = 1631.0 / 55296.0, b62 = 175.0 / 512.0, b63 = 575.0 / 13824.0;
        const double b64 = 44275.0 / 110592.0, b65 = 253.0 / 4096.0, c1 = 37.0 / 378.0;
        const double c3 = 250.0 / 621.0, c4 = 125.0 / 594.0, c6 = 512.0 / 1771.0;
        const double dc5 = -277.00 / 14336.0;
        const double dc1 = c1 - 2825.0 / 27648.0, dc3 = c3 - 18575.0 / 48384.0;
        const double dc4 = c4 - 13525.0 / 55296.0, dc6 = c6 - 0.25;
        double k20, k21, k30, k31, k40, k41, k50, k51, k60, k61;
        n_rkck++;
        // (pinned arithmetic: y + dt * (sum_j b_j k_j) with the sum accumulated left to right by FMAs)
        const double h21 = CTB_MUL(b21, dt);
        eom(fma(h21, d0, y0), fma(h21, d1, y1), fma(a2, dt, t), k20, k21);
        eom(fma(dt, fma(b32, k20, CTB_MUL(b31, d0)), y0), fma(dt, fma(b32, k21, CTB_MUL(b31, d1)), y1), fma(a3, dt, t),
            k30, k31);
        eom(fma(dt, fma(b43, k30, fma(b42, k20, CTB_MUL(b41, d0))), y0),
            fma(dt, fma(b43, k31, fma(b42, k21, CTB_MUL(b41, d1))), y1), fma(a4, dt, t), k40, k41);
        // (a5 = 1: this stage sits at t + dt, where the caller evaluates the derivative again after an accepted step;
        //  its reciprocal radius is kept for that)
        inv_r_end = (ALPHA == CTB_ALPHA_WALL) ? 0.0 : rcp_fast(fma(a5, dt, t));
        eom_inv(fma(dt, fma(b54, k40, fma(b53, k30, fma(b52, k20, CTB_MUL(b51, d0)))), y0),
                fma(dt, fma(b54, k41, fma(b53, k31, fma(b52, k21, CTB_MUL(b51, d1)))), y1), inv_r_end, k50, k51);
        eom(fma(dt, fma(b65, k50, fma(b64, k40, fma(b63, k30, fma(b62, k20, CTB_MUL(b61, d0))))), y0),
            fma(dt, fma(b65, k51, fma(b64, k41, fma(b63, k31, fma(b62, k21, CTB_MUL(b61, d1))))), y1), fma(a6, dt, t),
            k60, k61);
        dy0 = CTB_MUL(dt, fma(c6, k60, fma(c4, k40, fma(c3, k30, CTB_MUL(c1, d0)))));
        dy1 = CTB_MUL(dt, fma(c6, k61, fma(c4, k41, fma(c3, k31, CTB_MUL(c1, d1)))));
        e0 = CTB_MUL(dt, fma(dc6, k60, fma(dc5, k50, fma(dc4, k40, fma(dc3, k30, CTB_MUL(dc1, d0))))));
        e1 = CTB_MUL(dt, fma(dc6, k61, fma(dc5, k51, fma(dc4, k41, fma(dc3, k31, CTB_MUL(dc1, d1))))));
    }

    // HF:113-179; returns 0 or CTB_ST_STEP_UNDERFLOW.  dt is in/out (try -> used).
    CTB_DEV int rkqs(double y0, double y1, double d0, double d1, double t, double &dt, double epsfrac,
                     double inv_ea0, double inv_ea1, double &dy0, double &dy1, double &dtnext)
    {
        double errmax;
        for (;;) {
            double e0, e1;
            rkck(y0, y1, d0, d1, t, dt, dy0, dy1, e0, e1);
            double a0 = CTB_MUL(fabs(e0), inv_ea0), b0 = CTB_MUL(fabs(e0), rcp_fast(CTB_MUL(fabs(y0) + 1e-300, epsfrac)));
            double a1 = CTB_MUL(fabs(e1), inv_ea1), b1 = CTB_MUL(fabs(e1), rcp_fast(CTB_MUL(fabs(y1) + 1e-300, epsfrac)));
            double m0 = (a0 < b0) ? a0 : b0, m1 = (a1 < b1) ? a1 : b1;
            errmax = (m0 > m1) ? m0 : m1;
            // numpy min/max propagate NaN; nan_to_num: NaN -> 0, +inf -> DBL_MAX
            if (isnan((a0 + b0) + (a1 + b1))) errmax = 0.0; // all four are >= 0: the sum is NaN iff one is
            else if (isinf(errmax)) errmax = 1.7976931348623157e308;
            if (errmax < 1.0) break;
            double dttemp = CTB_MUL(CTB_MUL(0.9, dt), rsqrt(sqrt(errmax)));
            const double dt10 = CTB_MUL(dt, .1);
            if (dt > 0) dt = (dt10 > dttemp) ? dt10 : dttemp;
            else dt = (dt10 < dttemp) ? dt10 : dttemp;
            if (t + dt == t) return CTB_ST_STEP_UNDERFLOW;
        }
        if (errmax > 1.89e-4) dtnext = CTB_MUL(CTB_MUL(0.9, dt), inv_fifth_root(errmax));
        else dtnext = CTB_MUL(5.0, dt);
        return 0;
    }
};

// HF:583-599 cubic Bezier through (y0, y0'dr) and (y1, y1'dr), one component
// (the Bezier form Y0 (1-t)^3 + 3 Y1 (1-t)^2 t + 3 Y2 (1-t) t^2 + Y3 t^3 expanded in powers of t once per
// step, so that each of the ~500 dense-output points costs three FMAs per component)
struct Cubic {
    double A0, A1, A2, A3, Yend;
    CTB_DEV void init(double y0, double dy0, double y1, double dy1)
    {
        const double third = 1.0 / 3.0;
        double Y1 = fma(dy0, third, y0), Y2 = fma(-dy1, third, y1);
        A0 = y0;
        A1 = CTB_MUL(3.0, Y1 - y0);
        A2 = CTB_MUL(3.0, (y0 - Y1) + (Y2 - Y1));
        A3 = fma(3.0, Y1 - Y2, y1 - y0);
        Yend = y1;
    }
    CTB_DEV double operator()(double t) const { return fma(t, fma(t, fma(t, A3, A2), A1), A0); }
    CTB_DEV double at(double t) const { return (t == 1.0) ? Yend : (*this)(t); } // exact end point (root bracketing)
};

// Where the dense-output cubic of one step crosses `off` (T1D:521-534: brentq on [0, 1], xtol 2e-12): a bracketed
// Newton iteration started from the secant through the end points; the cubic is close to linear within a step, so two
// or three evaluations reach rounding level.  err as brentq's (1: no sign change, 2: NaN).
CTB_DEV double cubic_crossing(const Cubic &f, double off, int &err)
{
    err = 0;
    const double f0 = f.A0 - off, f1 = f.Yend - off;
    if (isnan(f0) || isnan(f1)) { err = 2; return nan(""); }
    if (f0 == 0) return 0.0;
    if (f1 == 0) return 1.0;
    if (signbit(f0) == signbit(f1)) { err = 1; return 0.; }
    double lo = 0.0, hi = 1.0; // f(lo) has the sign of f0
    double t = f0 / (f0 - f1);
    for (int it = 0; it < 60; it++) {
        const double ft = f(t) - off;
        if (ft == 0) return t;
        if (isnan(ft)) { err = 2; return nan(""); }
        if (signbit(ft) == signbit(f0)) lo = t; else hi = t;
        const double dft = fma(t, fma(t, CTB_MUL(3.0, f.A3), CTB_MUL(2.0, f.A2)), f.A1);
        double tn = t - ft * rcp_fast(dft); // (a zero / non-finite slope gives NaN or a point outside the bracket: bisection)
        if (fabs(tn - t) <= 1e-15) return fmin(fmax(tn, 0.0), 1.0); // converged (NaN compares false)
        if (!(tn > lo && tn < hi)) tn = 0.5 * (lo + hi);
        t = tn;
        if ((hi - lo) <= 1e-15) break;
    }
    return t;
}

struct InstanceResult {
    double S, phi0, r0, rf, x, phi_bar, rscale;
    int status, n_shots, n_rkck, n_points;
};

// SingleFieldInstanton.__init__ (T1D:128-149): metastability check, findBarrierLocation (T1D:203-226)
// and findRScale (T1D:228-285).  phi_bar / rscale come in as NaN ("None") or as overrides.
template <class Pot>
__device__ int sfi_init(const Pot &pot, double phi_abs, double phi_meta, double V_meta, double &phi_bar, double &rscale)
{
    if (V_meta <= pot.V(phi_abs)) { phi_bar = rscale = nan(""); return CTB_ST_STABLE; }
    if (isnan(phi_bar)) { // findBarrierLocation
        double phi_tol = fabs(phi_meta - phi_abs) * 1e-12;
        double phi1 = phi_meta, phi2 = phi_abs, phi0 = 0.5 * (phi1 + phi2);
        while (fabs(phi1 - phi2) > phi_tol) {
            double V0 = pot.V(phi0);
            if (V0 > V_meta) phi1 = phi0; else phi2 = phi0;
            phi0 = 0.5 * (phi1 + phi2);
        }
        phi_bar = phi0;
    }
    if (isnan(rscale)) { // findRScale
        double phi_tol = fabs(phi_bar - phi_meta) * 1e-6;
        double x1 = fmin(phi_bar, phi_meta), x2 = fmax(phi_bar, phi_meta);
        double top = fminbound([&](double x) { return -pot.V(x); }, x1, x2, phi_tol);
        if (top + phi_tol > x2 || top - phi_tol < x1) return CTB_ST_NO_BARRIER;
        double Vtop = pot.V(top) - V_meta;
        double xtop = top - phi_meta;
        if (Vtop <= 0) return CTB_ST_NO_BARRIER;
        rscale = fabs(xtop) / sqrt(fabs(6 * Vtop));
    }
    return 0;
}

// Work predictor used to order instances (heavy first, similar neighbours): the starting value of the
// shooting variable x = -log|(phi_bar - phi_abs)/(phi_meta - phi_abs)| (T1D:682-683).  Thin walls
// (phi_bar close to phi_meta) have large x: more x5 growth shots, more steps, 250 interior points.
CTB_DEV double work_key(double phi_abs, double phi_meta, double phi_bar)
{
    return -log(fabs((phi_bar - phi_abs) / (phi_meta - phi_abs)));
}

// __init__ + findProfile + findAction for one instance.                        T1D:128-149, 615-791
//
// Control flow.  ONE integration loop serves both the shooting passes (integrateProfile, T1D:489-536)
// and the final dense-output pass (integrateAndSaveProfile, T1D:578-613): a single copy of the
// Cash-Karp step keeps the kernel's hot code inside the instruction cache.  Every lane of the warp
// takes part in every trip of the outer loop (lanes without an instance, failed instances and lanes
// that finished shooting early just idle): the save pass -- 500 dense-output points, up to 250 interior
// points and the quadrature -- starts for the whole warp at once, on a warp vote, so its long uniform
// code runs once per warp instead of once per group of lanes that happen to finish together.
enum { PH_SHOOT = 0, PH_WAIT = 1, PH_SAVE = 2, PH_DONE = 3 };

// The solve can also be split over two kernels (MODE_SHOOT then MODE_SAVE) that hand the accepted shot
// over through global memory: all warps of an SM are then in the same phase, which halves the hot
// instruction footprint each kernel needs in the instruction cache.  MODE_FUSED is the single-kernel form.
enum { MODE_FUSED = 0, MODE_SHOOT = 1, MODE_SAVE = 2 };
struct HandOff { // state of the last integrated shot (pointers already offset to this instance)
    double *y00, *y01, *dphi0;
};

template <class Pot, int ALPHA, bool PROFILE, int MODE>
__device__ void solve_instance(Sfi<Pot, ALPHA> &s, const Knobs &k, bool valid, double phi_bar_in, double rscale_in,
                               InstanceResult &res, const ProfileSink &sink, const HandOff &h)
{
    const unsigned FULL = 0xffffffffu;
    const double qnan = nan("");
    // MODE_SAVE: res arrives filled with what the shoot kernel wrote (status, r0, rf, x, n_shots, n_rkck)
    const int status_in = res.status;
    const double r0_in = res.r0, rf_in = res.rf;
    if (MODE != MODE_SAVE) {
        res.S = res.phi0 = res.r0 = res.rf = res.x = res.phi_bar = res.rscale = qnan;
        res.n_shots = res.n_rkck = res.n_points = 0;
        res.status = CTB_ST_NOT_RUN;
        s.n_rkck = 0;
    } else {
        s.n_rkck = res.n_rkck;
        valid = valid && (status_in == CTB_ST_CONVERGED || status_in == CTB_ST_XTOL);
    }
    s.phi_eps = 0.0; s.inv_12eps = 0.0;
    s.alpha_rt = (ALPHA == CTB_ALPHA_REAL) ? k.alpha_real : (double)ALPHA;
    const double phi_abs = s.phi_abs, phi_meta = s.phi_meta;
    const Pot &pot = s.pot;
    int phase = valid ? (MODE == MODE_SAVE ? PH_WAIT : PH_SHOOT) : PH_DONE;

    // ---- __init__ : T1D:128-149 ----
    double V_meta = 0.0, phi_bar = phi_bar_in, rscale = rscale_in;
    if (valid) {
        V_meta = pot.V(phi_meta);
        if (MODE != MODE_SAVE) { // (the save kernel receives the phi_bar / rscale the shoot kernel used)
            int st = sfi_init(pot, phi_abs, phi_meta, V_meta, phi_bar, rscale);
            res.phi_bar = phi_bar;
            res.rscale = rscale;
            if (st) { res.status = st; phase = PH_DONE; rscale = 1.0; phi_bar = 0.5 * (phi_abs + phi_meta); }
        }
    }
    s.set_phi_eps(k.phi_eps * fabs(phi_abs - phi_meta));

    // ---- findProfile : T1D:676-700 ----
    const double xtol = k.xtol, phitol = k.phitol;
    double xmin = xtol * 10, xmax = INFINITY, x;
    if (!isnan(k.xguess)) x = k.xguess;
    else x = -log(fabs((phi_bar - phi_abs) / (phi_meta - phi_abs)));
    const double rmin = k.rmin * rscale;
    const double dr0 = rmin, drmin = 0.01 * rmin, rmax_rel = k.rmax * rscale;
    // guard (not in the reference, where this case never returns): zero or non-finite radial scale
    if (phase != PH_DONE && !(rmin > 0.0 && rmin < INFINITY)) { res.status = CTB_ST_NONFINITE; phase = PH_DONE; }
    const double delta_phi = phi_meta - phi_abs;
    const double ea0 = fabs(delta_phi * phitol), ea1 = fabs(delta_phi / rscale * phitol);
    const double epsfrac = phitol;
    const double inv_ea0 = 1.0 / ea0, inv_ea1 = 1.0 / ea1;
    const double cutoff = k.thinCutoff * delta_phi;
    const int npoints = k.npoints;
    const double afac = area_fac<ALPHA>(s.alpha_rt), vfac = vol_fac<ALPHA>(s.alpha_rt);

    bool have_rf = false;
    double rf = qnan, r0 = qnan, y00 = qnan, y01 = qnan, delta_phi0 = qnan, dv0 = qnan, d2v0 = qnan;
    int status = CTB_ST_XTOL, nshots = 0;
    if (MODE == MODE_SAVE && valid) { // take over the last integrated shot
        r0 = r0_in; rf = rf_in; y00 = *h.y00; y01 = *h.y01; delta_phi0 = *h.dphi0;
        status = status_in; nshots = res.n_shots;
        dv0 = s.dV_from_absMin(delta_phi0);
        d2v0 = s.d2V(phi_abs + delta_phi0);
    }
    // dense-output state (save pass only)
    SimpsonStream simp;
    int N = 0, i_pt = 1, pidx = 0;
    double step = 0.0, di = 1.0, Ri = 0.0, R_first = 0.0, Phi_first = 0.0;

    int g_dense = 0, g0 = 0; // global index of the next dense-output point; start of the uniform Simpson part
    auto emit = [&](double r, double phi, double dphi, bool dense) {
        double ra = pow_alpha<ALPHA>(r, s.alpha_rt);
        double integrand = (0.5 * (dphi * dphi) + pot.V_quad(phi) - V_meta) * (ra * afac);
        if (dense && simp.uniform_on) simp.feed_uniform(integrand, g_dense & 1);
        else {
            simp.feed(r, integrand);
            if (dense && g_dense == g0 && npoints >= 8) simp.begin_uniform();
        }
        if (dense) g_dense++;
        if (PROFILE) { sink.R[pidx] = r; sink.Phi[pidx] = phi; sink.dPhi[pidx] = dphi; pidx++; }
    };
    auto fail = [&](int st) { res.status = st; phase = PH_DONE; };

    for (;;) { // one trip = one shot (T1D:702-727) or, once for the whole warp, the save pass (T1D:729-761)
        if (__all_sync(FULL, phase != PH_SHOOT)) {
            if (MODE == MODE_SHOOT) break; // every lane has its accepted shot (or failed): hand over
            if (__all_sync(FULL, phase == PH_DONE)) break;
            if (phase == PH_WAIT) phase = PH_SAVE;
        }
        if (MODE != MODE_SAVE && phase == PH_SHOOT) {
            delta_phi0 = exp(-x) * delta_phi;
            dv0 = s.dV_from_absMin(delta_phi0);
            d2v0 = s.d2V(phi_abs + delta_phi0);
            double r0_, p0, dp0;
            int e = s.initial_conditions(delta_phi0, dv0, d2v0, rmin, cutoff, r0_, p0, dp0);
            if (e) fail(e);
            else if (!isfinite(r0_) || !isfinite(x)) {
                // "use the last finite values instead" (T1D:707-712); note delta_phi0 stays the new one
                if (!have_rf) fail(CTB_ST_FIRST_SHOT);
                else { status = CTB_ST_XTOL; phase = PH_WAIT; }
            } else {
                r0 = r0_; y00 = p0; y01 = dp0;
            }
        }
        if (MODE != MODE_SHOOT && phase == PH_SAVE) { // grid, interior points (T1D:734-760), first dense point
            res.r0 = r0; res.rf = rf;
            res.phi0 = phi_abs + delta_phi0;
            step = (rf - r0) / (npoints - 1);
            const double Rl1 = (npoints == 2) ? rf : __dadd_rn(__dmul_rn(1.0, step), r0);
            int nint = 0;
            double a_int = 0.0, dx0 = Rl1 - r0, st_int = 0.0;
            bool lin_int = true;
            const int max_interior = (k.max_interior_pts < 0) ? npoints / 2 : k.max_interior_pts;
            if (max_interior > 0) {
                if (r0 / dx0 <= max_interior) {
                    nint = (int)ceil(r0 / dx0);
                    st_int = r0 / nint;
                } else {
                    nint = max_interior;
                    lin_int = false;
                    a_int = (r0 / dx0 - nint) * 2 / ((double)nint * (nint + 1));
                }
            }
            N = nint + npoints;
            res.n_points = N;
            simp.init(N);
            const double phi_c = phi_abs + delta_phi0;
            // first profile point: r = 0 when there are interior points, else the start of the integration
            R_first = (nint > 0) ? 0.0 : r0;
            Phi_first = (nint > 0) ? phi_c : y00;
            if (nint > 0) {
                const ExactCoef ec = exact_coef(phi_c, dv0, d2v0, 0.5 * (s.alpha_rt - 1.0));
                emit(0.0, phi_c, 0.0, false);
                for (int i = 1; i < nint; i++) {
                    double ri;
                    if (lin_int) ri = __dadd_rn(__dmul_rn((double)i, st_int), 0.0);
                    else {
                        double Nv = (double)(nint - i);
                        ri = r0 - dx0 * (Nv + 0.5 * a_int * Nv * (Nv + 1));
                    }
                    double2 e = exact_solution<ALPHA>(ri, ec);
                    emit(ri, e.x, e.y, false);
                }
            }
            g_dense = nint;
            g0 = (nint & 1) ? nint + 1 : nint; // first even global index on the dense-output grid
            emit(r0, y00, y01, true);
            i_pt = 1; di = 1.0;
            Ri = Rl1;
        }

        // ---- integrate from (r0, y00, y01): T1D:489-536 (shooting) / T1D:578-613 (saving) ----
        if (phase == PH_SHOOT || phase == PH_SAVE) {
            // compile-time constant in the split kernels: the other mode's code is not generated
            const bool saving = (MODE == MODE_SAVE) ? true : (MODE == MODE_SHOOT) ? false : (phase == PH_SAVE);
            int ctype = -1;
            double r_ = r0, ya = y00, yb = y01, dr = dr0, da, db;
            s.eom(ya, yb, r_, da, db);
            const double ysign = sgn(ya - phi_meta);
            const double rmax = rmax_rel + r0;
            double yf0 = qnan, yf1 = qnan;
            for (;;) {
                double dy0, dy1, drnext;
                int e = s.rkqs(ya, yb, da, db, r_, dr, epsfrac, inv_ea0, inv_ea1, dy0, dy1, drnext);
                if (e) { fail(e); break; }
                if (s.n_rkck >= CTB_MAX_STEPS) { fail(CTB_ST_MAXSTEPS); break; } // guard, see ctbounce.h
                double r1, na, nb, inv_r1;
                if (saving && dr < drmin) { // T1D:594-597
                    na = ya + dy0 * drmin / dr; nb = yb + dy1 * drmin / dr;
                    dr = drnext = drmin;
                    r1 = r_ + dr;
                    inv_r1 = rcp_fast(r1);
                } else {
                    r1 = r_ + dr; na = ya + dy0; nb = yb + dy1;
                    inv_r1 = s.inv_r_end; // the Cash-Karp stage at t + dt already formed 1 / r1
                }
                double nda, ndb;
                s.eom_inv(na, nb, inv_r1, nda, ndb);
                if (!saving) {
                    if (r1 > rmax) { fail(CTB_ST_RMAX); break; }
                    else if (dr < drmin) { fail(CTB_ST_DRMIN); break; }
                    else if (fabs(na - phi_meta) < 3 * ea0 && fabs(nb) < 3 * ea1) {
                        rf = r1; yf0 = na; yf1 = nb; ctype = CT_CONVERGED;
                        break;
                    } else if (nb * ysign > 0 || (na - phi_meta) * ysign < 0) {
                        Cubic f0, f1;
                        f0.init(ya, dr * da, na, dr * nda);
                        f1.init(yb, dr * db, nb, dr * ndb);
                        // undershoot (phi' = 0) is tested first, as in the reference; one root find serves both
                        const bool under = nb * ysign > 0;
                        ctype = under ? CT_UNDERSHOOT : CT_OVERSHOOT;
                        const Cubic &fr = under ? f1 : f0;
                        const double off = under ? 0.0 : phi_meta;
                        int berr;
                        double xx = cubic_crossing(fr, off, berr);
                        if (berr == 1 || berr == 2) { fail((berr == 2) ? CTB_ST_NONFINITE : CTB_ST_BRENT_SIGN); break; }
                        rf = fma(dr, xx, r_);
                        yf0 = f0.at(xx); yf1 = f1.at(xx);
                        break;
                    }
                } else {
                    if (r_ < Ri && Ri <= r1) { // T1D:601-606
                        Cubic f0, f1;
                        f0.init(ya, dr * da, na, dr * nda);
                        f1.init(yb, dr * db, nb, dr * ndb);
                        const double inv_dr = rcp_fast(dr);
                        while (i_pt < npoints && r_ < Ri && Ri <= r1) {
                            double t = (Ri - r_) * inv_dr;
                            emit(Ri, f0(t), f1(t), true);
                            i_pt++;
                            di += 1.0;
                            Ri = (i_pt == npoints - 1) ? rf : __dadd_rn(__dmul_rn(di, step), r0);
                        }
                    }
                    if (i_pt >= npoints) break;
                }
                r_ = r1; ya = na; yb = nb; da = nda; db = ndb;
                dr = drnext;
            }
            if (phase == PH_SAVE) { // findAction : T1D:780-791
                const double x_m2 = (npoints >= 3) ? __dadd_rn(__dmul_rn((double)(npoints - 3), step), r0) : r0;
                const double x_m1 = (npoints >= 2) ? __dadd_rn(__dmul_rn((double)(npoints - 2), step), r0) : r0;
                double S = simp.finish(N, step, x_m2, x_m1, rf);
                double Rd = pow_dim<ALPHA>(R_first, s.alpha_rt);
                S += (Rd * vfac) * (pot.V_quad(Phi_first) - V_meta);
                res.S = S;
                res.status = status;
                phase = PH_DONE;
            } else if (phase == PH_SHOOT) { // bracket update: T1D:716-727
                if (fabs(yf0 - phi_meta) < 3 * ea0 && fabs(yf1) < 3 * ea1) ctype = CT_CONVERGED;
                have_rf = true;
                nshots++;
                res.x = x;
                if (ctype == CT_CONVERGED) { status = CTB_ST_CONVERGED; phase = PH_WAIT; }
                else {
                    if (ctype == CT_UNDERSHOOT) { xmin = x; x = (xmax == INFINITY) ? x * 5.0 : .5 * (xmin + xmax); }
                    else { xmax = x; x = .5 * (xmin + xmax); }
                    if ((xmax - xmin) < xtol) { status = CTB_ST_XTOL; phase = PH_WAIT; }
                }
            }
        }
    }
    res.n_shots = nshots;
    res.n_rkck = s.n_rkck;
    if (MODE == MODE_SHOOT && phase == PH_WAIT) { // accepted shot -> global memory for the save kernel
        res.status = status;
        res.r0 = r0; res.rf = rf;
        res.phi0 = phi_abs + delta_phi0;
        *h.y00 = y00; *h.y01 = y01; *h.dphi0 = delta_phi0;
    }
}

} // namespace ctb
