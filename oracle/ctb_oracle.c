/*
 * ctb_oracle.c -- CPU restatement (plain C, scalar FP64) of CosmoTransitions' 1-D bounce hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path
 * (cosmotransitions_b200/) never links, imports or calls anything in oracle/.
 *
 * Parity status: the reference ships no tests / golden vectors for this path, so
 * the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, generated in the build container by
 * oracle/gen_golden.py (imports /root/reference from a writable copy, numpy 2.3.5 / scipy 1.18.1)
 * and committed under tests/golden/.  tests/test_oracle_golden.py checks this file against them.
 *
 * What it follows (file:line into the reference checkout):
 *   cosmoTransitions/tunneling1D.py:128-149   SingleFieldInstanton.__init__        -> sfi_setup
 *   cosmoTransitions/tunneling1D.py:151-201   dV / d2V finite differences, dV_from_absMin -> sfi_dV, sfi_d2V, sfi_dV_from_absMin
 *   cosmoTransitions/tunneling1D.py:203-226   findBarrierLocation                  -> sfi_setup (bisection block)
 *   cosmoTransitions/tunneling1D.py:228-285   findRScale                           -> sfi_setup (fminbound block)
 *   cosmoTransitions/tunneling1D.py:294-373   exactSolution                        -> exact_solution
 *   cosmoTransitions/tunneling1D.py:377-434   initialConditions                    -> initial_conditions
 *   cosmoTransitions/tunneling1D.py:436-440   equationOfMotion                     -> eom
 *   cosmoTransitions/tunneling1D.py:444-536   integrateProfile                     -> integrate_profile
 *   cosmoTransitions/tunneling1D.py:539-613   integrateAndSaveProfile              -> integrate_and_save
 *   cosmoTransitions/tunneling1D.py:615-761   findProfile                          -> solve_one
 *   cosmoTransitions/tunneling1D.py:763-791   findAction                           -> solve_one (tail)
 *   cosmoTransitions/helper_functions.py:113-179  rkqs ; :204-236 _rkck ; :583-599 cubicInterpFunction
 *   cosmoTransitions/finiteT.py:167-174,197-204   Jf_spline / Jb_spline (evaluation + clamps)   -> J_spline
 *   cosmoTransitions/finiteT.py:207-296           Jb/Jf low- and high-x series                  -> ctbo_jseries
 *   cosmoTransitions/pathDeformation.py:811-834   SplinePath.V / dV / d2V (splev of one tck)    -> POT_SPLINE
 *   cosmoTransitions/generic_potential.py:240-337 V1, V1T, Vtot
 *   examples/testModel1.py:62-116                 model1.V0 / boson_massSq
 *
 * Third-party arithmetic the reference calls on this path (not vendored in the reference; scipy
 * 1.18.1 in the build container) and that is restated here from the published algorithms:
 *   scipy.optimize.brentq      (Zeros/brentq.c, Brent 1973 as coded by C. Harris)   -> brentq
 *   scipy.optimize.fminbound   (_optimize.py:_minimize_scalar_bounded)             -> fminbound
 *   scipy.integrate.simpson    (_quadrature.py: non-uniform composite Simpson + Cartwright
 *                               last-interval correction for even N)                -> simpson
 *   scipy.interpolate.splev    (FITPACK splev / splder / fpbspl, de Boor)           -> splev
 *   scipy.special.iv / jv / gamma for nu = (alpha-1)/2, alpha in {1,2,3,4}:
 *        closed forms (nu=1/2) and series/asymptotic I0,I1 + libm j0,j1 (nu=1)      -> bessel_*
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#define CTBO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* status codes (shared vocabulary with include/ctbounce.h)                                     */
enum {
    ST_CONVERGED = 0,      /* last shot ended "converged"                                       */
    ST_XTOL = 1,           /* bracket shrank below xtol                                         */
    ST_STABLE = 2,         /* PotentialError "stable, not metastable"                          */
    ST_NO_BARRIER = 3,     /* PotentialError "no barrier"                                      */
    ST_RMAX = 4,           /* IntegrationError "r > rmax"                                      */
    ST_DRMIN = 5,          /* IntegrationError "dr < drmin"                                    */
    ST_STEP_UNDERFLOW = 6, /* IntegrationError "Stepsize rounds down to zero."                 */
    ST_NONFINITE = 7,      /* ValueError out of brentq (NaN function value) in initialConditions*/
    ST_BRENT_SIGN = 8,     /* ValueError out of brentq: f(a), f(b) same sign                   */
    ST_FIRST_SHOT = 9,     /* AssertionError: non-finite initial conditions on first shot      */
    ST_BAD_ARG = 10
};

enum { CT_CONVERGED = 0, CT_UNDERSHOOT = 1, CT_OVERSHOOT = 2 };

/* potential kinds */
enum { POT_POLY4 = 0, POT_MODEL1_LINE = 1, POT_MODEL1F = 2, POT_SPLINE = 3, POT_THERMAL1F = 4 };

typedef struct {
    int n;           /* number of knots  */
    int k;           /* degree (3)       */
    const double *t; /* knots  [n]       */
    const double *c; /* coeffs [n]       */
    double xmin, xmax;
} spline_t;

typedef struct {
    int kind;
    const double *p;   /* per-instance parameters */
    const spline_t *jb;
    const spline_t *jf;
} pot_t;

typedef struct {
    double xtol, phitol, thinCutoff, rmin, rmax, phi_eps, xguess; /* xguess NaN -> None */
    int npoints, max_interior_pts; /* max_interior_pts < 0 -> None */
    int alpha;
} opts_t;

/* ------------------------------------------------------------------------------------------ */
/* FITPACK de Boor evaluation (splev for der=0, splder otherwise)                              */

static void fpbspl(const double *t, int k, double x, int l /*1-based*/, double *h)
{
    double hh[6];
    h[0] = 1.0;
    for (int j = 1; j <= k; j++) {
        for (int i = 0; i < j; i++) hh[i] = h[i];
        h[0] = 0.0;
        for (int i = 1; i <= j; i++) {
            int li = l + i, lj = li - j;
            double tli = t[li - 1], tlj = t[lj - 1];
            if (tli == tlj) {
                h[i] = 0.0;
            } else {
                double f = hh[i - 1] / (tli - tlj);
                h[i - 1] = h[i - 1] + f * (tli - x);
                h[i] = f * (x - tlj);
            }
        }
    }
}

static double splev(const spline_t *s, double x, int nu)
{
    const double *t = s->t;
    int n = s->n, k = s->k;
    int k1 = k + 1, nk1 = n - k1;
    /* interval search t(l) <= x < t(l+1), k1 <= l <= nk1 (1-based); FITPACK walks linearly from
     * l = k1, which for non-decreasing knots lands on the largest l in [k1, nk1] with t(l) <= x
     * (l = k1 when x < t(k1): extrapolation with the first piece).  Same l by bisection: */
    int l = k1;
    {
        int lo = k1, hi = nk1; /* invariant: answer in [lo, hi] */
        while (lo < hi) {
            int mid = (lo + hi + 1) >> 1;
            if (x >= t[mid - 1]) lo = mid; else hi = mid - 1;
        }
        l = lo;
    }
    double h[6];
    if (nu == 0) {
        fpbspl(t, k, x, l, h);
        double sp = 0.0;
        int ll = l - k1;
        for (int j = 1; j <= k1; j++) { ll += 1; sp += s->c[ll - 1] * h[j - 1]; }
        return sp;
    }
    /* splder: difference the coefficients nu times (whole array, as FITPACK does) */
    double wrk[1100];
    if (n > 1100) return NAN;
    for (int i = 0; i < nk1; i++) wrk[i] = s->c[i];
    int kk = k, lq = 1, nk2 = nk1;
    for (int j = 1; j <= nu; j++) {
        double ak = kk;
        nk2 -= 1;
        int l1 = lq;
        for (int i = 1; i <= nk2; i++) {
            l1 += 1;
            int l2 = l1 + kk;
            double fac = t[l2 - 1] - t[l1 - 1];
            if (fac > 0.0) wrk[i - 1] = ak * (wrk[i] - wrk[i - 1]) / fac;
        }
        lq += 1;
        kk -= 1;
    }
    if (kk == 0) return wrk[l - k1]; /* piecewise-constant derivative */
    fpbspl(t, kk, x, l, h);
    double sp = 0.0;
    int ll = l - k1;
    int k2 = k1 - nu;
    for (int j = 1; j <= k2; j++) { ll += 1; sp += wrk[ll - 1] * h[j - 1]; }
    return sp;
}

/* finiteT.py:167-174 / 197-204 */
static double J_spline(const spline_t *s, double theta, int der)
{
    double y = splev(s, theta, der);
    if (theta < s->xmin) y = splev(s, s->xmin, der);
    if (theta > s->xmax) y = 0.0;
    return y;
}

/* ------------------------------------------------------------------------------------------ */
/* potentials                                                                                  */

/* generic_potential.py:240-337 + examples/testModel1.py:62-116, along X(phi) = x0 + phi*nhat.
 * p = [l1, l2, mu2, Y1, Y2, nb, renormScaleSq, v2, T, x0_0, x0_1, n_0, n_1]                    */
static double vtot_model1(const double *p, const spline_t *jb, double X0, double X1, double T)
{
    double l1 = p[0], l2 = p[1], mu2 = p[2], Y1 = p[3], Y2 = p[4], nb = p[5], rs = p[6], v2 = p[7];
    double phi1 = X0, phi2 = X1;
    double a = l1 * (3 * phi1 * phi1 - v2);
    double b = l2 * (3 * phi2 * phi2 - v2);
    double A = .5 * (a + b);
    double B = sqrt(.25 * (a - b) * (a - b) + mu2 * mu2);
    double mb = Y1 * (phi1 * phi1 + phi2 * phi2) + Y2 * phi1 * phi2;
    double m2[3] = {A + B, A - B, mb};
    double dof[3] = {1.0, 1.0, nb};
    double c = 1.5;
    /* V0 */
    double r = .25 * l1 * (phi1 * phi1 - v2) * (phi1 * phi1 - v2)
             + .25 * l2 * (phi2 * phi2 - v2) * (phi2 * phi2 - v2);
    r -= mu2 * phi1 * phi2;
    /* V1 (fermion placeholder has dof 0: contributes exactly 0) */
    double y1 = 0.0;
    for (int i = 0; i < 3; i++)
        y1 += dof[i] * m2[i] * m2[i] * (log(fabs(m2[i] / rs) + 1e-100) - c);
    r += y1 / (64 * M_PI * M_PI);
    /* V1T */
    double T2 = T * T + 1e-100, T4 = T * T * T * T;
    double yt = 0.0;
    for (int i = 0; i < 3; i++) yt += dof[i] * J_spline(jb, m2[i] / T2, 0);
    r += yt * T4 / (2 * M_PI * M_PI);
    return r;
}

/* One-field model1-style thermal potential:
 * V0 = lam/4 (phi^2-v2)^2 ; bosons m^2 = {lam(3phi^2-v2), Y phi^2}, dof (1, nb), c = 1.5.
 * p = [lam, v2, Y, nb, renormScaleSq, T]                                                         */
static double vtot_model1f(const double *p, const spline_t *jb, double phi, double T)
{
    double lam = p[0], v2 = p[1], Y = p[2], nb = p[3], rs = p[4];
    double m2[2] = {lam * (3 * phi * phi - v2), Y * phi * phi};
    double dof[2] = {1.0, nb};
    double r = .25 * lam * (phi * phi - v2) * (phi * phi - v2);
    double y1 = 0.0;
    for (int i = 0; i < 2; i++)
        y1 += dof[i] * m2[i] * m2[i] * (log(fabs(m2[i] / rs) + 1e-100) - 1.5);
    r += y1 / (64 * M_PI * M_PI);
    double T2 = T * T + 1e-100, T4 = T * T * T * T;
    double yt = 0.0;
    for (int i = 0; i < 2; i++) yt += dof[i] * J_spline(jb, m2[i] / T2, 0);
    r += yt * T4 / (2 * M_PI * M_PI);
    return r;
}

/* General one-field generic_potential (generic_potential.py:240-337): quartic V0, bosons with
 * m^2 = a + b phi^2 + t T^2 (dof, c), fermions with m^2 = a + b phi^2 (dof).
 * p = [c0..c4, renormScaleSq, T, nb, nf, 8 x {a,b,t,dof,c}, 4 x {a,b,dof}, pad]                       */
static double vtot_thermal1f(const double *p, const spline_t *jb, const spline_t *jf, double phi)
{
    double rs = p[5], T = p[6];
    int nb = (int)p[7], nf = (int)p[8];
    double r = p[0] + phi * (p[1] + phi * (p[2] + phi * (p[3] + phi * p[4])));
    double T2 = T * T + 1e-100, T4 = T * T * T * T;
    double y1 = 0.0, yt = 0.0;
    for (int i = 0; i < nb; i++) {
        const double *q = p + 9 + 5 * i;
        double m2 = q[0] + q[1] * phi * phi + q[2] * T * T;
        y1 += q[3] * m2 * m2 * (log(fabs(m2 / rs) + 1e-100) - q[4]);
        yt += q[3] * J_spline(jb, m2 / T2, 0);
    }
    for (int j = 0; j < nf; j++) {
        const double *q = p + 49 + 3 * j;
        double m2 = q[0] + q[1] * phi * phi;
        y1 -= q[2] * m2 * m2 * (log(fabs(m2 / rs) + 1e-100) - 1.5);
        yt += q[2] * J_spline(jf, m2 / T2, 0);
    }
    return r + y1 / (64 * M_PI * M_PI) + yt * T4 / (2 * M_PI * M_PI);
}

static double pot_V(const pot_t *P, double phi)
{
    const double *p = P->p;
    switch (P->kind) {
    case POT_POLY4: /* V = c0 + c1 phi + c2 phi^2 + c3 phi^3 + c4 phi^4 */
        return p[0] + phi * (p[1] + phi * (p[2] + phi * (p[3] + phi * p[4])));
    case POT_MODEL1_LINE:
        return vtot_model1(p, P->jb, p[9] + phi * p[11], p[10] + phi * p[12], p[8]);
    case POT_MODEL1F:
        return vtot_model1f(p, P->jb, phi, p[5]);
    case POT_THERMAL1F:
        return vtot_thermal1f(p, P->jb, P->jf, phi);
    case POT_SPLINE: { /* pathDeformation.SplinePath.V: splev(x, tck); p = [n, t[n], c[n]] */
        spline_t s = {(int)p[0], 3, p + 1, p + 1 + (int)p[0], 0, 0};
        return splev(&s, phi, 0);
    }
    }
    return NAN;
}

static int pot_has_analytic_derivs(const pot_t *P) { return P->kind == POT_POLY4 || P->kind == POT_SPLINE; }

static double pot_dV_analytic(const pot_t *P, double phi)
{
    const double *p = P->p;
    if (P->kind == POT_SPLINE) { /* SplinePath.dV: splev(x, tck, der=1) */
        spline_t s = {(int)p[0], 3, p + 1, p + 1 + (int)p[0], 0, 0};
        return splev(&s, phi, 1);
    }
    return p[1] + phi * (2 * p[2] + phi * (3 * p[3] + phi * 4 * p[4]));
}
static double pot_d2V_analytic(const pot_t *P, double phi)
{
    const double *p = P->p;
    if (P->kind == POT_SPLINE) { /* SplinePath.d2V: splev(x, tck, der=2) */
        spline_t s = {(int)p[0], 3, p + 1, p + 1 + (int)p[0], 0, 0};
        return splev(&s, phi, 2);
    }
    return 2 * p[2] + phi * (6 * p[3] + phi * 12 * p[4]);
}

/* ------------------------------------------------------------------------------------------ */
/* instance state (SingleFieldInstanton attributes)                                            */
typedef struct {
    pot_t pot;
    double phi_absMin, phi_metaMin, phi_bar, rscale, phi_eps; /* phi_eps already rescaled */
    int alpha;
    int wall;        /* WallWithConstFriction: equation of motion with a constant friction term */
    double friction;
    /* counters */
    int n_rkck, n_exact;
} sfi_t;

/* tunneling1D.py:151-161 */
static double sfi_dV(sfi_t *s, double phi)
{
    if (pot_has_analytic_derivs(&s->pot)) return pot_dV_analytic(&s->pot, phi);
    double eps = s->phi_eps;
    return (pot_V(&s->pot, phi - 2 * eps) - 8 * pot_V(&s->pot, phi - eps)
            + 8 * pot_V(&s->pot, phi + eps) - pot_V(&s->pot, phi + 2 * eps)) / (12. * eps);
}
/* tunneling1D.py:191-201 */
static double sfi_d2V(sfi_t *s, double phi)
{
    if (pot_has_analytic_derivs(&s->pot)) return pot_d2V_analytic(&s->pot, phi);
    double eps = s->phi_eps;
    return (-pot_V(&s->pot, phi - 2 * eps) + 16 * pot_V(&s->pot, phi - eps) - 30 * pot_V(&s->pot, phi)
            + 16 * pot_V(&s->pot, phi + eps) - pot_V(&s->pot, phi + 2 * eps)) / (12. * eps * eps);
}
/* tunneling1D.py:163-189 */
static double sfi_dV_from_absMin(sfi_t *s, double delta_phi)
{
    double phi = s->phi_absMin + delta_phi;
    double dV = sfi_dV(s, phi);
    if (s->phi_eps > 0) {
        double dV_ = sfi_d2V(s, phi) * delta_phi;
        double q = delta_phi / s->phi_eps;
        double blend = exp(-(q * q));
        dV = dV_ * blend + dV * (1 - blend);
    }
    return dV;
}

/* ------------------------------------------------------------------------------------------ */
/* scipy.optimize.brentq restated (Zeros/brentq.c).  err: 0 ok, 1 sign error, 2 NaN value,
 * 3 not converged in maxiter                                                                  */
typedef double (*fn1_t)(double, void *);

static double brentq(fn1_t f, void *ctx, double xa, double xb, double xtol, double rtol, int maxiter,
                     int *err, int *funcalls)
{
    double xpre = xa, xcur = xb;
    double xblk = 0., fpre, fcur, fblk = 0., spre = 0., scur = 0., sbis;
    double delta, stry, dpre, dblk;
    *err = 0;
    fpre = f(xpre, ctx);
    if (isnan(fpre)) { *err = 2; return NAN; }
    fcur = f(xcur, ctx);
    if (isnan(fcur)) { *err = 2; return NAN; }
    if (funcalls) *funcalls = 2;
    if (fpre == 0) return xpre;
    if (fcur == 0) return xcur;
    if (signbit(fpre) == signbit(fcur)) { *err = 1; return 0.; }
    for (int i = 0; i < maxiter; i++) {
        if (fpre != 0 && fcur != 0 && (signbit(fpre) != signbit(fcur))) {
            xblk = xpre;
            fblk = fpre;
            spre = scur = xcur - xpre;
        }
        if (fabs(fblk) < fabs(fcur)) {
            xpre = xcur; xcur = xblk; xblk = xpre;
            fpre = fcur; fcur = fblk; fblk = fpre;
        }
        delta = (xtol + rtol * fabs(xcur)) / 2;
        sbis = (xblk - xcur) / 2;
        if (fcur == 0 || fabs(sbis) < delta) return xcur;
        if (fabs(spre) > delta && fabs(fcur) < fabs(fpre)) {
            if (xpre == xblk) {
                stry = -fcur * (xcur - xpre) / (fcur - fpre);
            } else {
                dpre = (fpre - fcur) / (xpre - xcur);
                dblk = (fblk - fcur) / (xblk - xcur);
                stry = -fcur * (fblk * dblk - fpre * dpre) / (dblk * dpre * (fblk - fpre));
            }
            double m1 = fabs(spre), m2 = 3 * fabs(sbis) - delta;
            double mn = (m1 < m2) ? m1 : m2;
            if (2 * fabs(stry) < mn) { spre = scur; scur = stry; }
            else { spre = sbis; scur = sbis; }
        } else {
            spre = sbis; scur = sbis;
        }
        xpre = xcur; fpre = fcur;
        if (fabs(scur) > delta) xcur += scur;
        else xcur += (sbis > 0 ? delta : -delta);
        fcur = f(xcur, ctx);
        if (funcalls) (*funcalls)++;
        if (isnan(fcur)) { *err = 2; return NAN; }
    }
    *err = 3;
    return xcur;
}
#define BRENT_XTOL 2e-12
#define BRENT_RTOL 8.881784197001252e-16
#define BRENT_MAXITER 100

/* scipy.optimize.fminbound restated (_optimize.py:_minimize_scalar_bounded), returns xf */
static double np_sign(double x) { return (x > 0) - (x < 0) + (isnan(x) ? NAN : 0.0); }

static double fminbound(fn1_t func, void *ctx, double x1, double x2, double xatol, int maxfun, int *nfev)
{
    const double sqrt_eps = sqrt(2.2e-16);
    const double golden_mean = 0.5 * (3.0 - sqrt(5.0));
    double a = x1, b = x2;
    double fulc = a + golden_mean * (b - a);
    double nfc = fulc, xf = fulc;
    double rat = 0.0, e = 0.0;
    double x = xf;
    double fx = func(x, ctx);
    int num = 1;
    double fu = INFINITY;
    double ffulc = fx, fnfc = fx;
    double xm = 0.5 * (a + b);
    double tol1 = sqrt_eps * fabs(xf) + xatol / 3.0;
    double tol2 = 2.0 * tol1;
    while (fabs(xf - xm) > (tol2 - 0.5 * (b - a))) {
        int golden = 1;
        if (fabs(e) > tol1) {
            golden = 0;
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
                    double si = np_sign(xm - xf) + ((xm - xf) == 0);
                    rat = tol1 * si;
                }
            } else {
                golden = 1;
            }
        }
        if (golden) {
            if (xf >= xm) e = a - xf; else e = b - xf;
            rat = golden_mean * e;
        }
        double si = np_sign(rat) + (rat == 0);
        double ar = fabs(rat);
        x = xf + si * (ar > tol1 ? ar : tol1); /* np.maximum */
        fu = func(x, ctx);
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
        if (num >= maxfun) break;
    }
    if (nfev) *nfev = num;
    return xf;
}

/* scipy.integrate.simpson(y, x=x) restated for 1-D non-uniform x (_quadrature.py) */
static double basic_simpson(const double *y, const double *x, int start, int stop)
{
    /* pairs (i, i+1, i+2) for i = start, start+2, ... < stop */
    double sum = 0.0;
    for (int i = start; i < stop; i += 2) {
        double h0 = x[i + 1] - x[i], h1 = x[i + 2] - x[i + 1];
        double hsum = h0 + h1, hprod = h0 * h1;
        double h0divh1 = (h1 != 0) ? h0 / h1 : 0.0;
        double inv = (h0divh1 != 0) ? 1.0 / h0divh1 : 0.0;
        double q = (hprod != 0) ? hsum / hprod : 0.0;
        double tmp = hsum / 6.0 * (y[i] * (2.0 - inv) + y[i + 1] * (hsum * q) + y[i + 2] * (2.0 - h0divh1));
        sum += tmp;
    }
    return sum;
}
static double simpson(const double *y, const double *x, int N)
{
    if (N % 2 == 0) {
        double val = 0.0, result = 0.0;
        if (N == 2) {
            val = 0.5 * (x[1] - x[0]) * (y[1] + y[0]);
        } else {
            result = basic_simpson(y, x, 0, N - 3);
            double h0 = x[N - 2] - x[N - 3], h1 = x[N - 1] - x[N - 2];
            double num, den, alpha, beta, eta;
            num = 2 * h1 * h1 + 3 * h0 * h1; den = 6 * (h1 + h0);
            alpha = (den != 0) ? num / den : 0.0;
            num = h1 * h1 + 3.0 * h0 * h1; den = 6 * h0;
            beta = (den != 0) ? num / den : 0.0;
            num = 1 * h1 * h1 * h1; den = 6 * h0 * (h0 + h1);
            eta = (den != 0) ? num / den : 0.0;
            result += alpha * y[N - 1] + beta * y[N - 2] - eta * y[N - 3];
        }
        return result + val;
    }
    return basic_simpson(y, x, 0, N - 2);
}

/* ------------------------------------------------------------------------------------------ */
/* Bessel functions for nu = (alpha-1)/2, alpha in {1, 2, 3, 4}                                */

/* I_0, I_1: ascending series for z < 30, Hankel asymptotic series beyond (overflow -> inf). */
static double bessel_in_series(int n, double z)
{
    double q = 0.25 * z * z;
    double term = (n == 0) ? 1.0 : 0.5 * z;
    double sum = term;
    for (int k = 1; k < 500; k++) {
        term *= q / ((double)k * (double)(k + n));
        sum += term;
        if (term < 1e-17 * sum) break;
    }
    return sum;
}
static double bessel_in_asym(int n, double z)
{
    double mu = 4.0 * n * n;
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 60; k++) {
        double tk = 2.0 * k - 1.0;
        double nt = -term * (mu - tk * tk) / (8.0 * k * z);
        if (fabs(nt) > fabs(term)) break;
        term = nt;
        sum += term;
        if (fabs(term) < 1e-17 * fabs(sum)) break;
    }
    /* exp(z)/sqrt(2 pi z), computed so that it overflows to inf only when the result does */
    double ez2 = exp(0.5 * z);
    return ez2 * (ez2 * (sum / sqrt(2 * M_PI * z)));
}
static double bessel_i(int n, double z)
{
    if (z < 30.0) return bessel_in_series(n, z);
    return bessel_in_asym(n, z);
}

/* I_nu (sg = +1) or J_nu (sg = -1) for nu = 1/2, 3/2, 5/2 by the ascending series
 * (z/2)^nu sum_k (sg z^2/4)^k / (k! Gamma(k + nu + 1)); used for z < 1 only. */
static void bessel_half_series(double sg, double z, double *b12, double *b32, double *b52)
{
    double nus[3] = {0.5, 1.5, 2.5}, out[3];
    double q = sg * 0.25 * z * z;
    for (int m = 0; m < 3; m++) {
        double nu = nus[m];
        double term = pow(0.5 * z, nu) / tgamma(nu + 1.0), sum = term;
        for (int k = 1; k < 60; k++) {
            term *= q / ((double)k * ((double)k + nu));
            sum += term;
            if (fabs(term) < 1e-18 * fabs(sum)) break;
        }
        out[m] = sum;
    }
    *b12 = out[0]; *b32 = out[1]; *b52 = out[2];
}

/* tunneling1D.py:294-373 */
static void exact_solution(sfi_t *s, double r, double phi0, double dV, double d2V, double *phi_out,
                           double *dphi_out)
{
    s->n_exact++;
    double beta = sqrt(fabs(d2V));
    double beta_r = beta * r;
    double nu = 0.5 * (s->alpha - 1);
    double phi, dphi;
    if (beta_r < 1e-2) {
        double sg = (d2V > 0) ? 1.0 : -1.0;
        phi = 0.0; dphi = 0.0;
        for (int k = 1; k < 4; k++) {
            double t_ = pow(0.5 * beta_r, 2 * k - 2) * pow(sg, k) / (tgamma(k + 1) * tgamma(k + 1 + nu));
            phi += t_;
            dphi += t_ * (2 * k);
        }
        phi *= 0.25 * tgamma(nu + 1) * (r * r) * dV * sg;
        dphi *= 0.25 * tgamma(nu + 1) * r * dV * sg;
        phi += phi0;
    } else {
        double z = beta_r;
        double Inu, Inum1, Inup1; /* or J */
        int modified = d2V > 0;
        if (s->alpha == 2) {
            double pre = sqrt(2.0 / (M_PI * z));
            if (modified) {
                if (z < 700.0) {
                    double sh = sinh(z), ch = cosh(z);
                    Inu = pre * sh; Inum1 = pre * ch; Inup1 = pre * (ch - sh / z);
                } else { /* sinh = cosh = e^z/2 to double precision; overflow only when I_nu does */
                    double e2 = exp(0.5 * z), q = (pre * 0.5 * e2) * e2;
                    Inu = q; Inum1 = q; Inup1 = q - q / z;
                }
            } else {
                double sn = sin(z), cs = cos(z);
                Inu = pre * sn; Inum1 = pre * cs; Inup1 = pre * (sn / z - cs);
            }
        } else if (s->alpha == 3) { /* nu = 1 */
            if (modified) {
                double i0 = bessel_i(0, z), i1 = bessel_i(1, z);
                Inu = i1; Inum1 = i0; Inup1 = i0 - 2.0 * i1 / z;
            } else {
                double j0_ = j0(z), j1_ = j1(z);
                Inu = j1_; Inum1 = j0_; Inup1 = 2.0 * j1_ / z - j0_;
            }
        } else if (s->alpha == 1) { /* nu = 0: I_{-1} = I_1, J_{-1} = -J_1 */
            if (modified) {
                double i0 = bessel_i(0, z), i1 = bessel_i(1, z);
                Inu = i0; Inum1 = i1; Inup1 = i1;
            } else {
                double j0_ = j0(z), j1_ = j1(z);
                Inu = j0_; Inum1 = -j1_; Inup1 = j1_;
            }
        } else { /* alpha == 4, nu = 3/2: spherical Bessel closed forms; ascending series below z = 1,
                    where the closed forms cancel */
            double pre = sqrt(2.0 / (M_PI * z));
            if (z < 1.0) {
                bessel_half_series(modified ? 1.0 : -1.0, z, &Inum1, &Inu, &Inup1);
            } else if (modified && z >= 700.0) { /* sinh = cosh = e^z / 2; overflow only when I_nu does */
                double e2 = exp(0.5 * z), q = (pre * 0.5 * e2) * e2;
                Inum1 = q;
                Inu = q * (1.0 - 1.0 / z);
                Inup1 = q * ((1.0 + 3.0 / (z * z)) - 3.0 / z);
            } else if (modified) {
                double sh = sinh(z), ch = cosh(z);
                Inum1 = pre * sh;
                Inu = pre * (ch - sh / z);
                Inup1 = pre * ((1.0 + 3.0 / (z * z)) * sh - 3.0 * ch / z);
            } else {
                double sn = sin(z), cs = cos(z);
                Inum1 = pre * sn;
                Inu = pre * (sn / z - cs);
                Inup1 = pre * ((3.0 / (z * z) - 1.0) * sn - 3.0 * cs / z);
            }
        }
        double hp = pow(0.5 * beta_r, -nu);
        double g = tgamma(nu + 1);
        if (modified) {
            phi = (g * hp * Inu - 1) * dV / d2V;
            dphi = -nu * (hp / r) * Inu;
            dphi += hp * 0.5 * beta * (Inum1 + Inup1);
            dphi *= g * dV / d2V;
            phi += phi0;
        } else {
            phi = (g * hp * Inu - 1) * dV / d2V;
            dphi = -nu * (hp / r) * Inu;
            dphi += hp * 0.5 * beta * (Inum1 - Inup1);
            dphi *= g * dV / d2V;
            phi += phi0;
        }
    }
    *phi_out = phi;
    *dphi_out = dphi;
}

/* tunneling1D.py:377-434 */
typedef struct { sfi_t *s; double phi0, dV, d2V, cutoff; } dpd_ctx;
static double delta_phi_diff(double r_, void *vctx)
{
    dpd_ctx *c = (dpd_ctx *)vctx;
    double p, dp;
    exact_solution(c->s, r_, c->phi0, c->dV, c->d2V, &p, &dp);
    return fabs(p - c->s->phi_absMin) - fabs(c->cutoff);
}

static int initial_conditions(sfi_t *s, double delta_phi0, double rmin, double delta_phi_cutoff,
                              double *r0_out, double *phi_out, double *dphi_out)
{
    double phi0 = s->phi_absMin + delta_phi0;
    double dV = sfi_dV_from_absMin(s, delta_phi0);
    double d2V = sfi_d2V(s, phi0);
    double phi_r0, dphi_r0;
    exact_solution(s, rmin, phi0, dV, d2V, &phi_r0, &dphi_r0);
    if (fabs(phi_r0 - s->phi_absMin) > fabs(delta_phi_cutoff)) {
        *r0_out = rmin; *phi_out = phi_r0; *dphi_out = dphi_r0;
        return 0;
    }
    {
        double s1 = np_sign(dphi_r0), s2 = np_sign(delta_phi0);
        if (s1 != s2) { /* NaN != anything is true, as in numpy */
            *r0_out = rmin; *phi_out = phi_r0; *dphi_out = dphi_r0;
            return 0;
        }
    }
    double r = rmin, rlast = rmin;
    while (isfinite(r)) {
        rlast = r;
        r *= 10;
        double phi, dphi;
        exact_solution(s, r, phi0, dV, d2V, &phi, &dphi);
        if (fabs(phi - s->phi_absMin) > fabs(delta_phi_cutoff)) break;
    }
    dpd_ctx ctx = {s, phi0, dV, d2V, delta_phi_cutoff};
    int err;
    double r0 = brentq(delta_phi_diff, &ctx, rlast, r, BRENT_XTOL, BRENT_RTOL, BRENT_MAXITER, &err, NULL);
    if (err == 2) return ST_NONFINITE;
    if (err == 1) return ST_BRENT_SIGN;
    exact_solution(s, r0, phi0, dV, d2V, &phi_r0, &dphi_r0);
    *r0_out = r0; *phi_out = phi_r0; *dphi_out = dphi_r0;
    return 0;
}

/* tunneling1D.py:436-440 */
static void eom(sfi_t *s, const double y[2], double r, double out[2])
{
    out[0] = y[1];
    if (s->wall) out[1] = sfi_dV(s, y[0]) - s->friction * y[1]; /* tunneling1D.py:912-916 */
    else out[1] = sfi_dV(s, y[0]) - s->alpha * y[1] / r;
}

/* helper_functions.py:204-236 */
static void rkck(sfi_t *s, const double y[2], const double dydt[2], double t, double dt, double dyout[2],
                 double yerr[2])
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
    double ytemp[2], ak2[2], ak3[2], ak4[2], ak5[2], ak6[2];
    s->n_rkck++;
    for (int i = 0; i < 2; i++) ytemp[i] = y[i] + b21 * dt * dydt[i];
    eom(s, ytemp, t + a2 * dt, ak2);
    for (int i = 0; i < 2; i++) ytemp[i] = y[i] + dt * (b31 * dydt[i] + b32 * ak2[i]);
    eom(s, ytemp, t + a3 * dt, ak3);
    for (int i = 0; i < 2; i++) ytemp[i] = y[i] + dt * (b41 * dydt[i] + b42 * ak2[i] + b43 * ak3[i]);
    eom(s, ytemp, t + a4 * dt, ak4);
    for (int i = 0; i < 2; i++)
        ytemp[i] = y[i] + dt * (b51 * dydt[i] + b52 * ak2[i] + b53 * ak3[i] + b54 * ak4[i]);
    eom(s, ytemp, t + a5 * dt, ak5);
    for (int i = 0; i < 2; i++)
        ytemp[i] = y[i] + dt * (b61 * dydt[i] + b62 * ak2[i] + b63 * ak3[i] + b64 * ak4[i] + b65 * ak5[i]);
    eom(s, ytemp, t + a6 * dt, ak6);
    for (int i = 0; i < 2; i++) {
        dyout[i] = dt * (c1 * dydt[i] + c3 * ak3[i] + c4 * ak4[i] + c6 * ak6[i]);
        yerr[i] = dt * (dc1 * dydt[i] + dc3 * ak3[i] + dc4 * ak4[i] + dc5 * ak5[i] + dc6 * ak6[i]);
    }
}

/* helper_functions.py:113-179.  returns 0 or ST_STEP_UNDERFLOW */
static int rkqs(sfi_t *s, const double y[2], const double dydt[2], double t, double dt_try,
                const double epsfrac[2], const double epsabs[2], double dy[2], double *dt_used,
                double *dt_next)
{
    double dt = dt_try, errmax, yerr[2];
    for (;;) {
        rkck(s, y, dydt, t, dt, dy, yerr);
        /* np.max(np.min([...], axis=0)) with numpy's NaN propagation, then nan_to_num */
        errmax = -INFINITY;
        int nan = 0;
        for (int i = 0; i < 2; i++) {
            double e1 = fabs(yerr[i] / epsabs[i]);
            double e2 = fabs(yerr[i]) / ((fabs(y[i]) + 1e-300) * epsfrac[i]);
            if (isnan(e1) || isnan(e2)) nan = 1;
            double m = (e1 < e2) ? e1 : e2;
            if (m > errmax) errmax = m;
        }
        if (nan) errmax = 0.0;
        else if (isinf(errmax)) errmax = 1.7976931348623157e308;
        if (errmax < 1.0) break;
        double dttemp = 0.9 * dt * pow(errmax, -.25);
        if (dt > 0) dt = (dt * .1 > dttemp) ? dt * .1 : dttemp; /* max(dttemp, dt*.1) */
        else        dt = (dt * .1 < dttemp) ? dt * .1 : dttemp;
        if (t + dt == t) return ST_STEP_UNDERFLOW;
    }
    if (errmax > 1.89e-4) *dt_next = 0.9 * dt * pow(errmax, -.2);
    else *dt_next = 5 * dt;
    *dt_used = dt;
    return 0;
}

/* helper_functions.py:583-599 */
typedef struct { double Y[4][2]; } cubic_t;
static void cubic_init(cubic_t *f, const double y0[2], const double dy0[2], const double y1[2],
                       const double dy1[2])
{
    for (int i = 0; i < 2; i++) {
        f->Y[0][i] = y0[i];
        f->Y[1][i] = y0[i] + dy0[i] / 3.0;
        f->Y[2][i] = y1[i] - dy1[i] / 3.0;
        f->Y[3][i] = y1[i];
    }
}
static double cubic_eval(const cubic_t *f, double t, int i)
{
    double mt = 1 - t;
    return f->Y[0][i] * (mt * mt * mt) + 3 * f->Y[1][i] * mt * mt * t + 3 * f->Y[2][i] * mt * t * t
           + f->Y[3][i] * (t * t * t);
}
typedef struct { const cubic_t *f; int comp; double off; } cub_ctx;
static double cubic_root_fn(double x, void *v)
{
    cub_ctx *c = (cub_ctx *)v;
    return cubic_eval(c->f, x, c->comp) - c->off;
}

/* tunneling1D.py:444-536. returns 0 or an ST_ error */
static int integrate_profile(sfi_t *s, double r0, const double y0_in[2], double dr0, const double epsfrac[2],
                             const double epsabs[2], double drmin, double rmax, double *r_out,
                             double y_out[2], int *ctype)
{
    double dr = dr0;
    double y0[2] = {y0_in[0], y0_in[1]}, dydr0[2];
    eom(s, y0, r0, dydr0);
    double ysign = np_sign(y0[0] - s->phi_metaMin);
    rmax += r0;
    double r, y[2];
    int ct;
    for (;;) {
        double dy[2], drnext, r1, y1[2], dydr1[2];
        int e = rkqs(s, y0, dydr0, r0, dr, epsfrac, epsabs, dy, &dr, &drnext);
        if (e) return e;
        r1 = r0 + dr;
        y1[0] = y0[0] + dy[0]; y1[1] = y0[1] + dy[1];
        eom(s, y1, r1, dydr1);
        if (r1 > rmax) return ST_RMAX;
        else if (dr < drmin) return ST_DRMIN;
        else if (fabs(y1[0] - s->phi_metaMin) < 3 * epsabs[0] && fabs(y1[1] - 0) < 3 * epsabs[1]) {
            r = r1; y[0] = y1[0]; y[1] = y1[1];
            ct = CT_CONVERGED;
            break;
        } else if (y1[1] * ysign > 0 || (y1[0] - s->phi_metaMin) * ysign < 0) {
            cubic_t f;
            double a[2] = {dr * dydr0[0], dr * dydr0[1]}, b[2] = {dr * dydr1[0], dr * dydr1[1]};
            cubic_init(&f, y0, a, y1, b);
            double x;
            int err;
            if (y1[1] * ysign > 0) {
                cub_ctx c = {&f, 1, 0.0};
                x = brentq(cubic_root_fn, &c, 0, 1, BRENT_XTOL, BRENT_RTOL, BRENT_MAXITER, &err, NULL);
                ct = CT_UNDERSHOOT;
            } else {
                cub_ctx c = {&f, 0, s->phi_metaMin};
                x = brentq(cubic_root_fn, &c, 0, 1, BRENT_XTOL, BRENT_RTOL, BRENT_MAXITER, &err, NULL);
                ct = CT_OVERSHOOT;
            }
            if (err == 2) return ST_NONFINITE;
            if (err == 1) return ST_BRENT_SIGN;
            r = r0 + dr * x;
            y[0] = cubic_eval(&f, x, 0);
            y[1] = cubic_eval(&f, x, 1);
            break;
        }
        r0 = r1; y0[0] = y1[0]; y0[1] = y1[1]; dydr0[0] = dydr1[0]; dydr0[1] = dydr1[1];
        dr = drnext;
    }
    if (fabs(y[0] - s->phi_metaMin) < 3 * epsabs[0] && fabs(y[1]) < 3 * epsabs[1]) ct = CT_CONVERGED;
    *r_out = r; y_out[0] = y[0]; y_out[1] = y[1]; *ctype = ct;
    return 0;
}

/* tunneling1D.py:539-613.  Fills Phi[N], dPhi[N] on R[N] */
static int integrate_and_save(sfi_t *s, const double *R, int N, const double y0_in[2], double dr,
                              const double epsfrac[2], const double epsabs[2], double drmin, double *Phi,
                              double *dPhi)
{
    double r0 = R[0];
    double y0[2] = {y0_in[0], y0_in[1]}, dydr0[2];
    for (int i = 0; i < N; i++) { Phi[i] = 0; dPhi[i] = 0; }
    Phi[0] = y0[0]; dPhi[0] = y0[1];
    eom(s, y0, r0, dydr0);
    int i = 1;
    while (i < N) {
        double dy[2], drnext, r1, y1[2], dydr1[2];
        int e = rkqs(s, y0, dydr0, r0, dr, epsfrac, epsabs, dy, &dr, &drnext);
        if (e) return e;
        if (dr >= drmin) {
            r1 = r0 + dr;
            y1[0] = y0[0] + dy[0]; y1[1] = y0[1] + dy[1];
        } else {
            y1[0] = y0[0] + dy[0] * drmin / dr; y1[1] = y0[1] + dy[1] * drmin / dr;
            dr = drnext = drmin;
            r1 = r0 + dr;
        }
        eom(s, y1, r1, dydr1);
        if (r0 < R[i] && R[i] <= r1) {
            cubic_t f;
            double a[2] = {dr * dydr0[0], dr * dydr0[1]}, b[2] = {dr * dydr1[0], dr * dydr1[1]};
            cubic_init(&f, y0, a, y1, b);
            while (i < N && r0 < R[i] && R[i] <= r1) {
                double x = (R[i] - r0) / dr;
                Phi[i] = cubic_eval(&f, x, 0);
                dPhi[i] = cubic_eval(&f, x, 1);
                i++;
            }
        }
        r0 = r1; y0[0] = y1[0]; y0[1] = y1[1]; dydr0[0] = dydr1[0]; dydr0[1] = dydr1[1];
        dr = drnext;
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double S, phi0, r0, rf, x, phi_bar, rscale, yf0, yf1;
    int status, n_shots, n_rkck, n_exact, n_points, ctype;
} result_t;

static double negV(double x, void *v) { return -pot_V((const pot_t *)v, x); }

/* tunneling1D.py:128-149, 203-285 */
static int sfi_setup(sfi_t *s, const pot_t *pot, double phi_absMin, double phi_metaMin, const opts_t *o,
                     double phi_bar_in, double rscale_in)
{
    s->pot = *pot;
    s->phi_absMin = phi_absMin; s->phi_metaMin = phi_metaMin;
    s->n_rkck = s->n_exact = 0;
    s->alpha = o->alpha;
    s->wall = 0; s->friction = 0.0;
    s->phi_eps = 0; s->phi_bar = NAN; s->rscale = NAN;
    if (pot_V(pot, phi_metaMin) <= pot_V(pot, phi_absMin)) return ST_STABLE;
    if (isnan(phi_bar_in)) {
        double phi_tol = fabs(phi_metaMin - phi_absMin) * 1e-12;
        double V_phimeta = pot_V(pot, phi_metaMin);
        double phi1 = phi_metaMin, phi2 = phi_absMin, phi0 = 0.5 * (phi1 + phi2);
        while (fabs(phi1 - phi2) > phi_tol) {
            double V0 = pot_V(pot, phi0);
            if (V0 > V_phimeta) phi1 = phi0; else phi2 = phi0;
            phi0 = 0.5 * (phi1 + phi2);
        }
        s->phi_bar = phi0;
    } else s->phi_bar = phi_bar_in;
    if (isnan(rscale_in)) {
        double phi_tol = fabs(s->phi_bar - phi_metaMin) * 1e-6;
        double x1 = fmin(s->phi_bar, phi_metaMin), x2 = fmax(s->phi_bar, phi_metaMin);
        double top = fminbound(negV, (void *)pot, x1, x2, phi_tol, 500, NULL);
        if (top + phi_tol > x2 || top - phi_tol < x1) return ST_NO_BARRIER;
        double Vtop = pot_V(pot, top) - pot_V(pot, phi_metaMin);
        double xtop = top - phi_metaMin;
        if (Vtop <= 0) return ST_NO_BARRIER;
        s->rscale = fabs(xtop) / sqrt(fabs(6 * Vtop));
    } else s->rscale = rscale_in;
    s->phi_eps = o->phi_eps * fabs(phi_absMin - phi_metaMin);
    return 0;
}

#define MAXPTS 4096

/* findProfile + findAction.  prof_* may be NULL; otherwise arrays of capacity prof_cap.
 * shot_x / shot_type (capacity shot_cap) log the shooting sequence when non-NULL. */
static double g_phi_bar_override = NAN, g_rscale_override = NAN; /* test hook, see ctbo_set_overrides */

static void solve_one(const pot_t *pot, double phi_absMin, double phi_metaMin, const opts_t *o, result_t *res,
                      double *prof_R, double *prof_Phi, double *prof_dPhi, int prof_cap, double *shot_x,
                      int *shot_type, int shot_cap)
{
    sfi_t S_, *s = &S_;
    memset(res, 0, sizeof(*res));
    res->S = res->phi0 = res->r0 = res->rf = res->x = res->yf0 = res->yf1 = NAN;
    int st = sfi_setup(s, pot, phi_absMin, phi_metaMin, o, g_phi_bar_override, g_rscale_override);
    res->phi_bar = s->phi_bar; res->rscale = s->rscale;
    if (st) { res->status = st; return; }
    int npoints = o->npoints;
    if (npoints < 2 || npoints > MAXPTS / 2 || (o->alpha < 1 || o->alpha > 4)) { res->status = ST_BAD_ARG; return; }

    /* tunneling1D.py:676-700 */
    double xtol = o->xtol, phitol = o->phitol;
    double xmin = xtol * 10, xmax = INFINITY, x;
    if (!isnan(o->xguess)) x = o->xguess;
    else x = -log(fabs((s->phi_bar - phi_absMin) / (phi_metaMin - phi_absMin)));
    const double xincrease = 5.0;
    double rmin = o->rmin * s->rscale;
    double dr0 = rmin, drmin = 0.01 * rmin, rmax = o->rmax * s->rscale;
    double delta_phi = phi_metaMin - phi_absMin;
    double epsabs[2] = {fabs(delta_phi * phitol), fabs(delta_phi / s->rscale * phitol)};
    double epsfrac[2] = {phitol, phitol};
    double delta_phi_cutoff = o->thinCutoff * delta_phi;
    int have_rf = 0, nshots = 0, ctype = -1;
    double rf = NAN, r0 = NAN, y0[2] = {NAN, NAN}, yf[2] = {NAN, NAN}, delta_phi0 = NAN;
    int final_status = ST_XTOL;
    for (;;) {
        delta_phi0 = exp(-x) * delta_phi;
        double r0_, phi0, dphi0;
        int e = initial_conditions(s, delta_phi0, rmin, delta_phi_cutoff, &r0_, &phi0, &dphi0);
        if (e) { res->status = e; goto counters; }
        if (!isfinite(r0_) || !isfinite(x)) {
            if (!have_rf) { res->status = ST_FIRST_SHOT; goto counters; }
            final_status = ST_XTOL;
            break;
        }
        r0 = r0_; y0[0] = phi0; y0[1] = dphi0;
        e = integrate_profile(s, r0, y0, dr0, epsfrac, epsabs, drmin, rmax, &rf, yf, &ctype);
        if (e) { res->status = e; goto counters; }
        have_rf = 1;
        if (shot_x && nshots < shot_cap) { shot_x[nshots] = x; shot_type[nshots] = ctype; }
        nshots++;
        res->x = x;
        if (ctype == CT_CONVERGED) { final_status = ST_CONVERGED; break; }
        else if (ctype == CT_UNDERSHOOT) { xmin = x; x = (xmax == INFINITY) ? x * xincrease : .5 * (xmin + xmax); }
        else { xmax = x; x = .5 * (xmin + xmax); }
        if ((xmax - xmin) < xtol) { final_status = ST_XTOL; break; }
    }
    res->n_shots = nshots;
    res->ctype = ctype;
    res->r0 = r0; res->rf = rf; res->yf0 = yf[0]; res->yf1 = yf[1];

    /* second pass: tunneling1D.py:729-761 */
    {
        static __thread double R[MAXPTS], Phi[MAXPTS], dPhi[MAXPTS], integrand[MAXPTS];
        int nint = 0;
        /* R = linspace(r0, rf, npoints) placed after the interior points: compute both first */
        double step = (rf - r0) / (npoints - 1);
        double Rl0 = r0, Rl1 = 1 * step + r0;
        if (npoints == 2) Rl1 = rf;
        int max_interior = (o->max_interior_pts < 0) ? npoints / 2 : o->max_interior_pts;
        double Rint[MAXPTS / 2 + 2];
        if (max_interior > 0) {
            double dx0 = Rl1 - Rl0;
            if (Rl0 / dx0 <= max_interior) {
                int n = (int)ceil(Rl0 / dx0);
                nint = n;
                double st_ = (Rl0 - 0.0) / n; /* linspace(0, R0, n+1)[:-1] */
                for (int i = 0; i < n; i++) Rint[i] = i * st_ + 0.0;
            } else {
                int n = max_interior;
                nint = n;
                double a = (Rl0 / dx0 - n) * 2 / ((double)n * (n + 1));
                for (int i = 0; i < n; i++) {
                    double Nv = n - i;
                    Rint[i] = Rl0 - dx0 * (Nv + 0.5 * a * Nv * (Nv + 1));
                }
                Rint[0] = 0.0;
            }
        }
        if (nint + npoints > MAXPTS) { res->status = ST_BAD_ARG; goto counters; }
        double *Rl = R + nint;
        for (int i = 0; i < npoints; i++) Rl[i] = i * step + r0;
        Rl[npoints - 1] = rf;
        int e = integrate_and_save(s, Rl, npoints, y0, dr0, epsfrac, epsabs, drmin, Phi + nint, dPhi + nint);
        if (e) { res->status = e; goto counters; }
        if (nint > 0) {
            Phi[0] = phi_absMin + delta_phi0;
            dPhi[0] = 0.0;
            R[0] = Rint[0];
            double dV = sfi_dV_from_absMin(s, delta_phi0);
            double d2V = sfi_d2V(s, Phi[0]);
            for (int i = 1; i < nint; i++) {
                R[i] = Rint[i];
                exact_solution(s, Rint[i], Phi[0], dV, d2V, &Phi[i], &dPhi[i]);
            }
        }
        int n = nint + npoints;
        res->n_points = n;
        /* findAction: tunneling1D.py:780-791 */
        int d = s->alpha + 1;
        double Vmeta = pot_V(pot, phi_metaMin);
        double afac_pi = pow(M_PI, d * .5), afac_g = tgamma(d * .5);
        for (int i = 0; i < n; i++) {
            double ra = (s->alpha == 2) ? R[i] * R[i] : pow(R[i], (double)s->alpha);
            double area = ra * 2 * afac_pi / afac_g;
            double ig = 0.5 * (dPhi[i] * dPhi[i]) + pot_V(pot, Phi[i]) - Vmeta;
            integrand[i] = ig * area;
        }
        double Sv = simpson(integrand, R, n);
        double volume = pow(R[0], (double)d) * pow(M_PI, d * .5) / tgamma(d * .5 + 1);
        Sv += volume * (pot_V(pot, Phi[0]) - Vmeta);
        res->S = Sv;
        res->phi0 = phi_absMin + delta_phi0;
        if (prof_R) {
            int m = n < prof_cap ? n : prof_cap;
            memcpy(prof_R, R, m * sizeof(double));
            memcpy(prof_Phi, Phi, m * sizeof(double));
            memcpy(prof_dPhi, dPhi, m * sizeof(double));
        }
    }
    res->status = final_status;
counters:
    res->n_rkck = s->n_rkck;
    res->n_exact = s->n_exact;
}

/* ------------------------------------------------------------------------------------------ */
/* exported C interface (ctypes)                                                               */

typedef struct {
    int kind, nparam, alpha, npoints, max_interior_pts;
    double xtol, phitol, thinCutoff, rmin, rmax, phi_eps, xguess;
    /* spline tables (may be NULL for POLY4) */
    const double *jb_t, *jb_c; int jb_n; double jb_xmin, jb_xmax;
    const double *jf_t, *jf_c; int jf_n; double jf_xmin, jf_xmax;
} ctbo_config;

typedef struct {
    const ctbo_config *cfg;
    const double *params, *phi_abs, *phi_meta;
    int64_t n;
    double *S, *phi0, *r0, *rf, *x, *phi_bar, *rscale;
    int *status, *n_shots, *n_rkck, *n_exact, *n_points, *ctype;
    int tid, nthreads;
} job_t;

static void cfg_to(const ctbo_config *c, opts_t *o, spline_t *jb, spline_t *jf)
{
    o->xtol = c->xtol; o->phitol = c->phitol; o->thinCutoff = c->thinCutoff; o->rmin = c->rmin;
    o->rmax = c->rmax; o->phi_eps = c->phi_eps; o->xguess = c->xguess; o->npoints = c->npoints;
    o->max_interior_pts = c->max_interior_pts; o->alpha = c->alpha;
    jb->n = c->jb_n; jb->k = 3; jb->t = c->jb_t; jb->c = c->jb_c; jb->xmin = c->jb_xmin; jb->xmax = c->jb_xmax;
    jf->n = c->jf_n; jf->k = 3; jf->t = c->jf_t; jf->c = c->jf_c; jf->xmin = c->jf_xmin; jf->xmax = c->jf_xmax;
}

static void *job_run(void *v)
{
    job_t *j = (job_t *)v;
    opts_t o; spline_t jb, jf;
    cfg_to(j->cfg, &o, &jb, &jf);
    for (int64_t i = j->tid; i < j->n; i += j->nthreads) {
        pot_t pot = {j->cfg->kind, j->params + i * j->cfg->nparam, &jb, &jf};
        result_t r;
        solve_one(&pot, j->phi_abs[i], j->phi_meta[i], &o, &r, NULL, NULL, NULL, 0, NULL, NULL, 0);
        j->S[i] = r.S; j->phi0[i] = r.phi0; j->r0[i] = r.r0; j->rf[i] = r.rf; j->x[i] = r.x;
        j->phi_bar[i] = r.phi_bar; j->rscale[i] = r.rscale;
        j->status[i] = r.status; j->n_shots[i] = r.n_shots; j->n_rkck[i] = r.n_rkck;
        j->n_exact[i] = r.n_exact; j->n_points[i] = r.n_points; j->ctype[i] = r.ctype;
    }
    return NULL;
}

/* Batched findProfile+findAction over n instances on nthreads host threads. */
CTBO_API int ctbo_bounce_batch(const ctbo_config *cfg, const double *params, const double *phi_abs,
                               const double *phi_meta, int64_t n, int nthreads, double *S, double *phi0,
                               double *r0, double *rf, double *x, double *phi_bar, double *rscale, int *status,
                               int *n_shots, int *n_rkck, int *n_exact, int *n_points, int *ctype)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 1024) nthreads = 1024;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    job_t *jobs = (job_t *)malloc(sizeof(job_t) * nthreads);
    for (int t = 0; t < nthreads; t++) {
        job_t j = {cfg, params, phi_abs, phi_meta, n, S, phi0, r0, rf, x, phi_bar, rscale,
                   status, n_shots, n_rkck, n_exact, n_points, ctype, t, nthreads};
        jobs[t] = j;
        if (nthreads > 1) pthread_create(&th[t], NULL, job_run, &jobs[t]);
    }
    if (nthreads == 1) job_run(&jobs[0]);
    else for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th); free(jobs);
    return 0;
}

/* One instance with its profile and shot log. Returns number of profile points. */
CTBO_API int ctbo_bounce_one(const ctbo_config *cfg, const double *params, double phi_abs, double phi_meta,
                             double *out9 /*S,phi0,r0,rf,x,phi_bar,rscale,yf0,yf1*/,
                             int *iout6 /*status,n_shots,n_rkck,n_exact,n_points,ctype*/, double *prof_R,
                             double *prof_Phi, double *prof_dPhi, int prof_cap, double *shot_x, int *shot_type,
                             int shot_cap)
{
    opts_t o; spline_t jb, jf;
    cfg_to(cfg, &o, &jb, &jf);
    pot_t pot = {cfg->kind, params, &jb, &jf};
    result_t r;
    solve_one(&pot, phi_abs, phi_meta, &o, &r, prof_R, prof_Phi, prof_dPhi, prof_cap, shot_x, shot_type, shot_cap);
    out9[0] = r.S; out9[1] = r.phi0; out9[2] = r.r0; out9[3] = r.rf; out9[4] = r.x; out9[5] = r.phi_bar;
    out9[6] = r.rscale; out9[7] = r.yf0; out9[8] = r.yf1;
    iout6[0] = r.status; iout6[1] = r.n_shots; iout6[2] = r.n_rkck; iout6[3] = r.n_exact; iout6[4] = r.n_points;
    iout6[5] = r.ctype;
    return r.n_points;
}

/* --- unit-level entry points for the parity tests ------------------------------------------ */

CTBO_API void ctbo_jspline(const double *t, const double *c, int n, double xmin, double xmax, const double *theta,
                           double *out, int64_t m, int der)
{
    spline_t s = {n, 3, t, c, xmin, xmax};
    for (int64_t i = 0; i < m; i++) out[i] = J_spline(&s, theta[i], der);
}

CTBO_API void ctbo_potential(const ctbo_config *cfg, const double *params, const double *phi, double *out,
                             int64_t m)
{
    opts_t o; spline_t jb, jf;
    cfg_to(cfg, &o, &jb, &jf);
    pot_t pot = {cfg->kind, params, &jb, &jf};
    for (int64_t i = 0; i < m; i++) out[i] = pot_V(&pot, phi[i]);
}

/* Vtot of model1 on the outer-product grid X(s_i) = x0 + s_i*nhat, T_j : out[i*nT + j].
 * params as POT_MODEL1_LINE (the T slot is ignored). */
CTBO_API void ctbo_vtot_model1_grid(const ctbo_config *cfg, const double *params, const double *s, int64_t ns,
                                    const double *T, int64_t nT, double *out)
{
    opts_t o; spline_t jb, jf;
    cfg_to(cfg, &o, &jb, &jf);
    for (int64_t i = 0; i < ns; i++)
        for (int64_t j = 0; j < nT; j++)
            out[i * nT + j] = vtot_model1(params, &jb, params[9] + s[i] * params[11],
                                          params[10] + s[i] * params[12], T[j]);
}

typedef struct { double (*f)(double); } pyfn_ctx;
CTBO_API double ctbo_brentq_cubic(const double Y[8], int comp, double off, int *err, int *funcalls)
{
    cubic_t f;
    memcpy(f.Y, Y, sizeof(f.Y));
    cub_ctx c = {&f, comp, off};
    return brentq(cubic_root_fn, &c, 0, 1, BRENT_XTOL, BRENT_RTOL, BRENT_MAXITER, err, funcalls);
}

static double poly_fn(double x, void *v)
{
    const double *p = (const double *)v; /* p[0]=degree, then coefficients high->low */
    int d = (int)p[0];
    double y = p[1];
    for (int i = 1; i <= d; i++) y = y * x + p[1 + i];
    return y;
}
CTBO_API double ctbo_brentq_poly(const double *p, double a, double b, int *err, int *funcalls)
{
    return brentq(poly_fn, (void *)p, a, b, BRENT_XTOL, BRENT_RTOL, BRENT_MAXITER, err, funcalls);
}
CTBO_API double ctbo_fminbound_poly(const double *p, double a, double b, double xatol, int *nfev)
{
    return fminbound(poly_fn, (void *)p, a, b, xatol, 500, nfev);
}
CTBO_API double ctbo_simpson(const double *y, const double *x, int n) { return simpson(y, x, n); }

CTBO_API void ctbo_exact_solution(int alpha, double r, double phi0, double dV, double d2V, double *out2)
{
    sfi_t s; memset(&s, 0, sizeof(s)); s.alpha = alpha;
    exact_solution(&s, r, phi0, dV, d2V, &out2[0], &out2[1]);
}

/* one rkqs call on a POLY4 potential: out = dy0, dy1, dt_used, dt_next ; returns status */
CTBO_API int ctbo_rkqs_poly4(const double *coef5, int alpha, const double y[2], double t, double dt_try,
                             const double epsfrac[2], const double epsabs[2], double out4[4])
{
    sfi_t s; memset(&s, 0, sizeof(s));
    s.alpha = alpha; s.pot.kind = POT_POLY4; s.pot.p = coef5;
    double dydt[2], dy[2], dtu = NAN, dtn = NAN;
    eom(&s, y, t, dydt);
    int e = rkqs(&s, y, dydt, t, dt_try, epsfrac, epsabs, dy, &dtu, &dtn);
    out4[0] = dy[0]; out4[1] = dy[1]; out4[2] = dtu; out4[3] = dtn;
    return e;
}

/* ---- finiteT.Jb / Jf, approx = 'high' and 'low' (finiteT.py:207-296) -------------------------------
 * scipy.special.kn(0|1|2, z) restated through K_nu(z) = int_0^inf exp(-z cosh t) cosh(nu t) dt, trapezoidal
 * rule with step 0.2 (error ~ exp(-pi^2/0.2)), K_2 = K_0 + 2 K_1 / z. */
static void bessel_k01(double z, double *k0, double *k1)
{
    const double h = 0.2;
    double s0 = 0.5 * exp(-z), s1 = s0;
    for (int k = 1; k < 8192; k++) {
        double ch = cosh(k * h);
        double f = exp(-z * ch);
        s0 += f; s1 += f * ch;
        if (f * ch <= 1e-19 * s1) break;
    }
    *k0 = h * s0; *k1 = h * s1;
}
static double nan_to_num(double y)
{
    if (isnan(y)) return 0.0;
    if (isinf(y)) return y > 0 ? 1.7976931348623157e308 : -1.7976931348623157e308;
    return y;
}
static double high_term(int deriv, double k, double x)
{
    double k0 = INFINITY, k1 = INFINITY, ax = fabs(x);
    if (deriv == 0) { /* x2K2, finiteT.py:249-255 */
        if (x == 0) return -2.0 / (k * k * k * k);
        if (x < 0) return NAN;
        bessel_k01(k * x, &k0, &k1);
        return -x * x * (k0 + 2.0 * k1 / (k * x)) / (k * k);
    }
    if (ax > 0) bessel_k01(k * ax, &k0, &k1);
    if (deriv == 1) return nan_to_num(x * ax * k1 / k);                             /* :258-260 */
    if (deriv == 2) return (ax == 0) ? 1.0 / (k * k) : nan_to_num(ax * (k1 / k - ax * k0)); /* :263-270 */
    return nan_to_num(x * (ax * k * k1 - 3 * k0));                                  /* :273-275 */
}
CTBO_API void ctbo_jseries(int which, int approx_low, int deriv, int nterms, const double *lowcoef, const double *x,
                           double *out, int64_t n)
{
    for (int64_t i = 0; i < n; i++) {
        double y = 0.0, xi = x[i];
        if (!approx_low) { /* Jb_high / Jf_high, finiteT.py:278-296 */
            double sg = 1.0;
            for (int k = 1; k <= nterms; k++) { y += sg * high_term(deriv, (double)k, xi); if (which == 1) sg = -sg; }
        } else {           /* Jb_low / Jf_low, finiteT.py:225-244 */
            double lg = nan_to_num(log(xi * xi));
            const double *g;
            if (which == 0) {
                y = lowcoef[0] + xi * xi * (lowcoef[1] + xi * (lowcoef[2] + lowcoef[3] * xi * (lg - lowcoef[4])));
                g = lowcoef + 5;
            } else {
                y = lowcoef[0] + xi * xi * (lowcoef[1] + lowcoef[2] * xi * xi * (lg - lowcoef[3]));
                g = lowcoef + 4;
            }
            for (int k = 1; k <= nterms; k++) y += g[k - 1] * pow(xi, 2 * k + 4);
        }
        out[i] = y;
    }
}

/* V_i(phi_i), one field value per parameter row */
CTBO_API void ctbo_potential_rows(const ctbo_config *cfg, const double *params, const double *phi, double *out,
                                  int64_t n)
{
    opts_t o; spline_t jb, jf;
    cfg_to(cfg, &o, &jb, &jf);
    for (int64_t i = 0; i < n; i++) {
        pot_t pot = {cfg->kind, params + i * cfg->nparam, &jb, &jf};
        out[i] = pot_V(&pot, phi[i]);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* WallWithConstFriction (tunneling1D.py:829-999): findRScale by a quadratic fit (:838-864), start at r = 0 on the
 * exponential solution near phi_absMin (:866-910), shooting in the friction F (:960-990), dense output (:992-997).
 * dout = F, rf, phi_bar, rscale, phi0, dphi0 ; iout = status, n_shots, n_rkck, n_points                          */
CTBO_API void ctbo_wall_one(const ctbo_config *cfg, const double *params, double phi_absMin, double phi_metaMin,
                            double Fguess, double Ftol, double phi0_rel, double *dout, int *iout, double *prof_R,
                            double *prof_Phi, double *prof_dPhi, int prof_cap)
{
    opts_t o; spline_t jb, jf;
    cfg_to(cfg, &o, &jb, &jf);
    pot_t pot = {cfg->kind, params, &jb, &jf};
    sfi_t S_, *s = &S_;
    for (int i = 0; i < 6; i++) dout[i] = NAN;
    iout[0] = ST_BAD_ARG; iout[1] = iout[2] = iout[3] = 0;
    memset(s, 0, sizeof(*s));
    s->pot = pot; s->phi_absMin = phi_absMin; s->phi_metaMin = phi_metaMin;
    s->alpha = o.alpha; s->wall = 1;
    /* __init__ (tunneling1D.py:128-149) with the overridden findRScale */
    if (pot_V(&pot, phi_metaMin) <= pot_V(&pot, phi_absMin)) { iout[0] = ST_STABLE; return; }
    {
        double phi_tol = fabs(phi_metaMin - phi_absMin) * 1e-12;
        double V_phimeta = pot_V(&pot, phi_metaMin);
        double phi1 = phi_metaMin, phi2 = phi_absMin, phi0 = 0.5 * (phi1 + phi2);
        while (fabs(phi1 - phi2) > phi_tol) {
            double V0 = pot_V(&pot, phi0);
            if (V0 > V_phimeta) phi1 = phi0; else phi2 = phi0;
            phi0 = 0.5 * (phi1 + phi2);
        }
        s->phi_bar = phi0;
    }
    dout[2] = s->phi_bar;
    {
        double pA = phi_absMin, pB = 0.5 * (s->phi_bar + phi_metaMin), pC = phi_metaMin;
        double yA = pot_V(&pot, pA), yB = pot_V(&pot, pB), yC = pot_V(&pot, pC);
        double lmda = 2 * ((yA - yB) / (pA - pB) - (yB - yC) / (pB - pC)) / (pC - pA);
        if (lmda <= 0.0) { iout[0] = ST_NO_BARRIER; return; }
        s->rscale = M_PI / sqrt(lmda);
    }
    dout[3] = s->rscale;
    s->phi_eps = o.phi_eps * fabs(phi_absMin - phi_metaMin);
    int npoints = o.npoints;
    if (npoints < 2 || npoints > MAXPTS) { iout[0] = ST_BAD_ARG; return; }
    /* findProfile (tunneling1D.py:918-999) */
    double rmin = o.rmin * s->rscale;
    double dr0 = rmin, drmin = 0.01 * rmin, rmax = o.rmax * s->rscale;
    double delta_phi = phi_metaMin - phi_absMin;
    double epsabs[2] = {fabs(delta_phi * o.phitol), fabs(delta_phi / s->rscale * o.phitol)};
    double epsfrac[2] = {o.phitol, o.phitol};
    double Fmin = 0.0, Fmax = INFINITY, F;
    if (!isnan(Fguess)) F = Fguess;
    else {
        double Delta_V = pot_V(&pot, phi_metaMin) - pot_V(&pot, phi_absMin);
        F = Delta_V * s->rscale / (delta_phi * delta_phi);
    }
    Ftol *= F;
    const double Fincrease = 5.0;
    double rf = NAN, r0 = 0.0, y0[2] = {NAN, NAN}, yf[2];
    int nshots = 0, ctype = -1, status = ST_XTOL;
    double d2V_abs = sfi_d2V(s, phi_absMin);
    for (;;) {
        /* initialConditions (tunneling1D.py:866-910) */
        double k = 0.5 * (sqrt(F * F + 4 * d2V_abs) - F);
        r0 = 0.0;
        y0[0] = phi_absMin + phi0_rel * (phi_metaMin - phi_absMin);
        y0[1] = k * (y0[0] - phi_absMin);
        s->friction = F;
        int e = integrate_profile(s, r0, y0, dr0, epsfrac, epsabs, drmin, rmax, &rf, yf, &ctype);
        if (e) { iout[0] = e; iout[1] = nshots; iout[2] = s->n_rkck; return; }
        nshots++;
        if (ctype == CT_CONVERGED) { status = ST_CONVERGED; break; }
        else if (ctype == CT_UNDERSHOOT) { Fmax = F; F = (Fmin == 0.0) ? F / Fincrease : .5 * (Fmin + Fmax); }
        else { Fmin = F; F = (Fmax == INFINITY) ? F * Fincrease : .5 * (Fmin + Fmax); }
        if ((Fmax - Fmin) < Ftol) { status = ST_XTOL; break; }
    }
    /* second pass with the CURRENT F (tunneling1D.py:992-997) */
    {
        static __thread double R[MAXPTS], Phi[MAXPTS], dPhi[MAXPTS];
        double step = (rf - r0) / (npoints - 1);
        for (int i = 0; i < npoints; i++) R[i] = i * step + r0;
        R[npoints - 1] = rf;
        s->friction = F;
        int e = integrate_and_save(s, R, npoints, y0, dr0, epsfrac, epsabs, drmin, Phi, dPhi);
        if (e) { iout[0] = e; iout[1] = nshots; iout[2] = s->n_rkck; return; }
        if (prof_R) {
            int m = npoints < prof_cap ? npoints : prof_cap;
            memcpy(prof_R, R, m * sizeof(double));
            memcpy(prof_Phi, Phi, m * sizeof(double));
            memcpy(prof_dPhi, dPhi, m * sizeof(double));
        }
    }
    dout[0] = F; dout[1] = rf; dout[4] = y0[0]; dout[5] = y0[1];
    iout[0] = status; iout[1] = nshots; iout[2] = s->n_rkck; iout[3] = npoints;
}

/* SingleFieldInstanton.findAction(profile) for a given profile (tunneling1D.py:763-791); alpha is a real number */
CTBO_API double ctbo_find_action(const ctbo_config *cfg, const double *params, double phi_meta, double alpha,
                                 const double *R, const double *Phi, const double *dPhi, int n)
{
    opts_t o; spline_t jb, jf;
    cfg_to(cfg, &o, &jb, &jf);
    pot_t pot = {cfg->kind, params, &jb, &jf};
    double d = alpha + 1.0;
    double Vmeta = pot_V(&pot, phi_meta);
    double *integrand = (double *)malloc(sizeof(double) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {
        double area = pow(R[i], alpha) * 2 * pow(M_PI, d * .5) / tgamma(d * .5);
        integrand[i] = (0.5 * (dPhi[i] * dPhi[i]) + pot_V(&pot, Phi[i]) - Vmeta) * area;
    }
    double S = simpson(integrand, R, n);
    double volume = pow(R[0], d) * pow(M_PI, d * .5) / tgamma(d * .5 + 1);
    S += volume * (pot_V(&pot, Phi[0]) - Vmeta);
    free(integrand);
    return S;
}

static double posV(double x, void *v) { return pot_V((const pot_t *)v, x); }

/* fminbound of V_i on [lo_i, hi_i], xatol_i = xtol_rel (hi_i - lo_i)  (scipy.optimize.fminbound restated) */
CTBO_API void ctbo_minimize_1d(const ctbo_config *cfg, const double *params, const double *lo, const double *hi,
                               double xtol_rel, int64_t n, double *xmin, double *vmin)
{
    opts_t o; spline_t jb, jf;
    cfg_to(cfg, &o, &jb, &jf);
    for (int64_t i = 0; i < n; i++) {
        pot_t pot = {cfg->kind, params + i * cfg->nparam, &jb, &jf};
        xmin[i] = fminbound(posV, &pot, lo[i], hi[i], xtol_rel * (hi[i] - lo[i]), 500, NULL);
        if (vmin) vmin[i] = pot_V(&pot, xmin[i]);
    }
}

/* phi_bar= / rscale= overrides of __init__ for subsequent calls (NaN = None). Not thread-safe: tests only. */
CTBO_API void ctbo_set_overrides(double phi_bar, double rscale)
{
    g_phi_bar_override = phi_bar;
    g_rscale_override = rscale;
}

CTBO_API int ctbo_abi_version(void) { return 1; }
