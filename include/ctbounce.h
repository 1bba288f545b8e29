/*
 * ctbounce.h -- C ABI of libctbounce.so: the B200 (sm_100a) batched bounce solver.
 *
 * Drop-in boundary for ONE hot path of CosmoTransitions (all file:line into the reference):
 *   tunneling1D.SingleFieldInstanton.__init__ / findProfile / findAction   tunneling1D.py:128-149,615-791
 *   helper_functions.rkqs / _rkck / cubicInterpFunction                   helper_functions.py:113-236,583-599
 *   finiteT.Jb_spline / Jf_spline                                          finiteT.py:167-174,197-204
 *   generic_potential.Vtot (V0 + V1 + V1T) for examples/testModel1.py      generic_potential.py:240-337
 *
 * The reference is pure Python and has no FFI; these entry points are what a ctypes binding of that
 * path binds (see INTEGRATION.md for the reference-side stub).  Conventions:
 *   - plain C types only; every pointer named d_* is a DEVICE pointer owned by the caller
 *     (contiguous FP64 / int32), on the device that is current when the call is made;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous
 *     with respect to the host unless stated otherwise;
 *   - the library keeps no state between calls and allocates nothing;
 *   - return value: 0 on success, CTB_ERR_* (<0) for argument errors, or a positive cudaError_t;
 *   - numerical failures are reported PER INSTANCE in the status array (CTB_ST_*), never as a
 *     process-level error.  The Python host maps them back to the reference's exceptions.
 */
#ifndef CTBOUNCE_H
#define CTBOUNCE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTB_ABI_VERSION 12

#if defined(__GNUC__)
#define CTB_API __attribute__((visibility("default")))
#else
#define CTB_API
#endif

/* ---- return codes --------------------------------------------------------------------------- */
#define CTB_OK 0
#define CTB_ERR_BAD_ARG (-1)      /* NULL pointer, n < 0, bad enum ...                       */
#define CTB_ERR_UNSUPPORTED (-2)  /* alpha out of range, npoints out of range, unknown kind    */
#define CTB_ERR_NO_DEVICE (-3)    /* no CUDA device / not an sm_100 device                   */

/* ---- per-instance status (int32) ------------------------------------------------------------ */
#define CTB_ST_CONVERGED 0       /* last shot ended "converged"            tunneling1D.py:717   */
#define CTB_ST_XTOL 1            /* bracket narrower than xtol             tunneling1D.py:726   */
#define CTB_ST_STABLE 2          /* PotentialError "stable, not metastable" tunneling1D.py:133-135 */
#define CTB_ST_NO_BARRIER 3      /* PotentialError "no barrier"            tunneling1D.py:266-283 */
#define CTB_ST_RMAX 4            /* IntegrationError "r > rmax"            tunneling1D.py:506   */
#define CTB_ST_DRMIN 5           /* IntegrationError "dr < drmin"          tunneling1D.py:508   */
#define CTB_ST_STEP_UNDERFLOW 6  /* IntegrationError "Stepsize rounds down to zero." helper_functions.py:173 */
#define CTB_ST_NONFINITE 7       /* ValueError from brentq on a NaN value  tunneling1D.py:432   */
#define CTB_ST_BRENT_SIGN 8      /* ValueError from brentq: same signs     tunneling1D.py:432,519,523 */
#define CTB_ST_FIRST_SHOT 9      /* AssertionError                         tunneling1D.py:710   */
#define CTB_ST_NOT_RUN 10
/* Guard that the reference does not have: an instance may take at most CTB_MAX_STEPS Cash-Karp attempts in total
 * (valid bounces need 1e2..1e4; the wall solver, whose legitimate "r > rmax" runs take ~1e5 steps, allows 8x as
 * many); beyond that it is abandoned instead of occupying the GPU for hours (the reference would keep stepping until
 * r > rmax, up to rmax/drmin = 1e10 steps per shot).  Likewise a set-up whose radial scale is zero or not finite
 * (where the reference's start-radius search never returns) ends with CTB_ST_NONFINITE.                       */
#define CTB_ST_MAXSTEPS 11
#define CTB_MAX_STEPS (1 << 20)

/* ---- potential functors (device-side V(phi); parameters are rows of d_params) --------------- */
/* V = c0 + c1 phi + c2 phi^2 + c3 phi^3 + c4 phi^4 ; analytic dV, d2V.  params[5] = c0..c4        */
#define CTB_POT_POLY4 0
/* examples/testModel1.py model1 along the line X(phi) = x0 + phi*nhat at temperature T;
 * dV, d2V by the 5-point stencils of tunneling1D.py:151-161,191-201.
 * params[13] = l1, l2, mu2, Y1, Y2, n_boson, renormScaleSq, v2, T, x0[0], x0[1], nhat[0], nhat[1]   */
#define CTB_POT_MODEL1_LINE 1
/* one-field model1-style generic_potential: V0 = lam/4 (phi^2-v2)^2, bosons m^2 =
 * {lam(3phi^2-v2), Y phi^2}, dof (1, n), c = 3/2.  params[6] = lam, v2, Y, n, renormScaleSq, T      */
#define CTB_POT_MODEL1F 2

/* Tabulated potential: an interpolating cubic spline in piecewise-polynomial form -- what
 * pathDeformation.SplinePath hands to the tunneling class (V, dV, d2V = splev of one cubic B-spline with
 * der = 0, 1, 2; pathDeformation.py:801-834, 959-964).  Outside the table the end pieces extrapolate, as
 * splev does.  params[1280] = nint (<= 255, as a double), brk[0..255], coef[255 * 4] (c0..c3 per piece:
 * V = c0 + dx (c1 + dx (c2 + dx c3)), dx = phi - brk[j]), zero padded.                              */
#define CTB_POT_SPLINE 3
#define CTB_SPLINE_MAX_PIECES 255

/* General one-field generic_potential (generic_potential.py:132-337): quartic tree-level potential plus up
 * to 8 bosons and 4 fermions with field-dependent masses m^2 = a + b phi^2 (+ t T^2 for bosons), Vtot = V0 + V1
 * + V1T with Jb for bosons and Jf for fermions; dV, d2V by the 5-point stencils.  params[64] =
 *   c0..c4 (V0 = sum c_k phi^k), renormScaleSq, T, n_bosons, n_fermions,
 *   bosons  8 x {a, b, t, dof, c}   (params[9 + 5 i ...]),
 *   fermions 4 x {a, b, dof}        (params[49 + 3 j ...]),  3 unused.                                      */
#define CTB_POT_THERMAL1F 4
#define CTB_THERMAL1F_MAX_BOSONS 8
#define CTB_THERMAL1F_MAX_FERMIONS 4

#define CTB_NPARAM_POLY4 5
#define CTB_NPARAM_THERMAL1F 64
#define CTB_NPARAM_SPLINE 1280
#define CTB_NPARAM_MODEL1_LINE 13
#define CTB_NPARAM_MODEL1F 6

/* ---- Jb / Jf spline table in piecewise-polynomial form (device memory) ----------------------
 * The reference evaluates scipy's (t, c, k=3) B-spline (finiteT.py:166,196).  The host converts it
 * once to one cubic per knot interval: value = c0 + dx (c1 + dx (c2 + dx c3)), dx = theta - brk[j].  */
typedef struct ctb_jtable {
    const double *d_brk;  /* [nint + 1] strictly increasing breakpoints                           */
    const double *d_coef; /* [nint * 4] c0..c3 per interval, 32-byte aligned                       */
    int32_t nint;
    int32_t reserved;
    double xmin;          /* theta < xmin  -> value (or derivative) at xmin   finiteT.py:172,202   */
    double xmax;          /* theta > xmax  -> 0                               finiteT.py:173,203   */
    /* Optional closed-form interval locator.  The reference samples both functions at
     * x_k = |xmin| sinh(u_k) / 20 with u_k uniform (finiteT.py:159-161,189-191), so
     * k ~ (asinh(theta * loc_scale) - loc_u0) * loc_inv_du and the knot interval is k - 1 (not-a-knot
     * spline: x_1 and x_{n-2} are not breakpoints); the kernels correct the guess by comparing with
     * d_brk.  loc_inv_du = 0 selects plain bisection.                                             */
    double loc_scale, loc_u0, loc_inv_du;
} ctb_jtable;

/* ---- findProfile knobs (tunneling1D.py:128-130, 615-617); same defaults as the reference ---- */
typedef struct ctb_opts {
    double xtol;              /* 1e-4 */
    double phitol;            /* 1e-4 */
    double thinCutoff;        /* 0.01 */
    double rmin;              /* 1e-4 (relative to rscale) */
    double rmax;              /* 1e4  (relative to rscale) */
    double phi_eps;           /* 1e-3 (relative to |phi_absMin - phi_metaMin|) */
    double xguess;            /* NaN = None */
    int32_t npoints;          /* 500; 2 <= npoints <= 4096 */
    int32_t max_interior_pts; /* -1 = None (npoints/2) */
    int32_t alpha;            /* 1..4: d = alpha + 1 dimensions; 2 = O(3) finite T, 3 = O(4) zero T;
                                 0 = the real value in alpha_real (general-order Bessel start, tunneling1D.py:336-372;
                                 thread-per-instance kernel, CTB_POT_POLY4 / CTB_POT_SPLINE functors)               */
    int32_t block_threads;    /* 0 = library default (thread-per-instance kernel only) */
    int32_t kernel;           /* CTB_KERNEL_AUTO / _THREAD / _WARP, see below */
    int32_t reserved;
    double alpha_real;        /* used when alpha == 0: any real 0 <= alpha <= 8 (Bessel order nu = (alpha-1)/2 <= 7/2) */
} ctb_opts;

/* Two kernels implement the same solve (identical shooting sequence and results up to the summation
 * order of the quadrature):
 *   THREAD: one GPU thread per instance -- throughput; the choice for >= ~4096 instances.
 *   WARP:   one warp per instance -- latency; the 31 nodes of the next five levels of the over/undershoot
 *           decision tree are integrated speculatively, one per lane, and resolved by warp shuffles; the
 *           save pass is lane-parallel with the profile staged in shared memory.
 *   AUTO:   WARP below 4096 instances (when npoints + interior points <= 2048), THREAD otherwise. */
#define CTB_KERNEL_AUTO 0
#define CTB_KERNEL_THREAD 1
#define CTB_KERNEL_WARP 2

/* ---- inputs of the batched solve: one row / entry per instance ------------------------------ */
typedef struct ctb_bounce_in {
    const double *d_params;   /* [n, nparam] row-major potential parameters                          */
    const double *d_phi_abs;  /* [n] phi_absMin                                                      */
    const double *d_phi_meta; /* [n] phi_metaMin                                                     */
    const double *d_phi_bar;  /* optional [n]: the phi_bar= override of __init__ (tunneling1D.py:140-143);
                                 NULL or a NaN entry = findBarrierLocation                           */
    const double *d_rscale;   /* optional [n]: the rscale= override (tunneling1D.py:144-147)         */
    const int64_t *d_order;   /* optional [n] permutation: GPU thread t solves instance d_order[t]
                                 (results are always written at the instance's own index)           */
    double *d_scratch;        /* optional [3 n] workspace.  When given (thread-per-instance kernel, no profile
                                 dump, all scalar outputs requested) the solve runs as two kernels -- shooting,
                                 then dense output + action -- handing the accepted shot over through it  */
    int64_t n;
} ctb_bounce_in;

/* ---- outputs of the batched solve: structure of arrays, any pointer may be NULL ------------- */
typedef struct ctb_bounce_out {
    double *d_S;         /* Euclidean action            findAction, tunneling1D.py:763-791          */
    double *d_phi0;      /* phi(r=0) = phi_absMin + exp(-x) (phi_metaMin - phi_absMin)              */
    double *d_r0;        /* radius where the integration started (profile.R[n_interior])            */
    double *d_rf;        /* radius where it ended (profile.R[-1])                                    */
    double *d_x;         /* shooting parameter of the last shot                                      */
    double *d_phi_bar;   /* findBarrierLocation                                                      */
    double *d_rscale;    /* findRScale                                                               */
    int32_t *d_status;   /* CTB_ST_*                                                                 */
    int32_t *d_n_shots;  /* integrateProfile calls                                                   */
    int32_t *d_n_rkck;   /* Cash-Karp step attempts (shooting + save pass)                           */
    int32_t *d_n_points; /* profile length                                                           */
    /* optional profile dump (Profile1D.R / .Phi / .dPhi): row i holds n_points[i] entries          */
    double *d_prof_R;
    double *d_prof_Phi;
    double *d_prof_dPhi;
    int64_t prof_stride; /* row length in doubles (>= npoints + npoints/2)                           */
} ctb_bounce_out;

CTB_API int ctb_abi_version(void);
CTB_API const char *ctb_strerror(int code);
CTB_API const char *ctb_status_string(int status);

/* Device check: returns CTB_OK if the current device can run the sm_100a kernels. */
CTB_API int ctb_device_ok(void);

/* Number of SMs of the current device (148 on B200): the persistent kernels size their grids from it;
 * exported so that callers (bench, tools) can size theirs the same way.  <= 0: no device. */
CTB_API int ctb_sm_count(void);

/* SingleFieldInstanton(phi_absMin, phi_metaMin, V, dV, d2V, phi_eps, alpha).findProfile(**knobs)
 * followed by .findAction(profile), for n independent instances (tunneling1D.py:128-791).
 * jb/jf may be NULL for POLY4. */
CTB_API int ctb_bounce_batch(int pot_kind, const ctb_bounce_in *in, const ctb_opts *opts, const ctb_jtable *jb,
                             const ctb_jtable *jf, const ctb_bounce_out *out, void *stream);

/* WallWithConstFriction(phi_absMin, phi_metaMin, V, dV, d2V, phi_eps).findProfile(Fguess, Ftol, phitol, npoints,
 * rmin, rmax, phi0_rel) (tunneling1D.py:829-999) for n instances: the wall profile phi'' + F phi' = V'(phi) between the
 * two minima, found by over/undershooting in the constant friction F; rscale from the quadratic fit of
 * WallWithConstFriction.findRScale (tunneling1D.py:838-864).  Uses in->d_params / d_phi_abs / d_phi_meta / n only.
 * Outputs (ctb_bounce_out): d_x = F, d_phi0 = phi(r = 0), d_r0 = 0, d_rf, d_phi_bar, d_rscale, d_status, d_n_shots,
 * d_n_rkck, d_n_points (= npoints) and the optional profile dump; d_S is not written.                          */
typedef struct ctb_wall_opts {
    double Fguess;   /* NaN = None: Delta_V rscale / delta_phi^2 (tunneling1D.py:965-969) */
    double Ftol;     /* 1e-4, relative to the first F */
    double phitol;   /* 1e-4 */
    double rmin;     /* 1e-4 (relative to rscale) */
    double rmax;     /* 1e4  (relative to rscale) */
    double phi0_rel; /* 1e-3 */
    double phi_eps;  /* 1e-3 (relative to |phi_absMin - phi_metaMin|) */
    int32_t npoints; /* 500; 2 <= npoints <= 4096 */
    int32_t reserved;
} ctb_wall_opts;
CTB_API int ctb_wall_batch(int pot_kind, const ctb_bounce_in *in, const ctb_wall_opts *opts, const ctb_jtable *jb,
                           const ctb_jtable *jf, const ctb_bounce_out *out, void *stream);

/* SingleFieldInstanton.findAction(profile) (tunneling1D.py:763-791) for n given profiles: row i of d_R / d_Phi /
 * d_dPhi ([n, stride], row-major) holds d_npts[i] points of Profile1D.R / .Phi / .dPhi; S_i = simpson of
 * (phi'^2/2 + V_i(phi) - V_i(phi_meta_i)) Omega_alpha r^alpha over the non-uniform R (scipy.integrate.simpson's
 * rule, including its even-N end correction) + the bulk term of the bubble interior.  alpha: any real >= 0
 * (the quadrature has no Bessel start).  A profile with d_npts[i] < 1 or > stride gives NaN. */
CTB_API int ctb_find_action(int pot_kind, const double *d_params, const ctb_jtable *jb, const ctb_jtable *jf,
                            const double *d_phi_meta, const double *d_R, const double *d_Phi, const double *d_dPhi,
                            int64_t stride, const int32_t *d_npts, int64_t n, double alpha, double *d_S, void *stream);

/* SingleFieldInstanton.__init__ alone (tunneling1D.py:128-149): metastability check,
 * findBarrierLocation, findRScale.  Fills out->d_phi_bar, d_rscale, d_status (0, CTB_ST_STABLE or
 * CTB_ST_NO_BARRIER) and, if d_work_key != NULL, a per-instance work predictor (float32, larger =
 * more shooting work) that the host sorts on to build ctb_bounce_in.d_order. */
CTB_API int ctb_bounce_setup(int pot_kind, const ctb_bounce_in *in, const ctb_jtable *jb, const ctb_jtable *jf,
                             const ctb_bounce_out *out, float *d_work_key, void *stream);

/* finiteT.Jb_spline(theta, n=deriv) / Jf_spline (finiteT.py:167-174,197-204): out[i] = J^(deriv)(theta[i]).
 * deriv in 0..3. */
CTB_API int ctb_jspline(const ctb_jtable *tab, int deriv, const double *d_theta, double *d_out, int64_t n, void *stream);

/* finiteT.Jb / Jf with approx = 'high' or 'low' (finiteT.py:225-296, front ends :302-375).  NOTE the
 * argument is x = m/T here (not theta = x^2 as for the spline variant).
 *   which: 0 = Jb, 1 = Jf;  approx: CTB_APPROX_HIGH (deriv 0..3, sum of nterms Bessel-K terms,
 *   finiteT.py:249-296) or CTB_APPROX_LOW (deriv 0 only, polynomial + log + nterms <= 50 power terms,
 *   finiteT.py:207-244).  lowcoef (HOST pointer, used for 'low' only): Jb: a, b, c, d, logab, g[50];
 *   Jf: a, b, d, logaf, g[50] -- the reference's lowCoef_b / lowCoef_f. */
#define CTB_APPROX_HIGH 0
#define CTB_APPROX_LOW 1
CTB_API int ctb_jseries(int which, int approx, int deriv, int nterms, const double *lowcoef, const double *d_x,
                        double *d_out, int64_t n, void *stream);

/* finiteT.Jb / Jf with approx = 'exact' (finiteT.py:40-147): the defining integrals
 *   Jb = int_0^inf y^2 log(1 - exp(-sqrt(y^2 + theta))) dy,  Jf = -int_0^inf y^2 log(1 + exp(-sqrt(y^2 + theta))) dy
 * by fixed-node double-exponential / trapezoidal quadrature on the device (absolute error ~1e-14; the reference's
 * QUADPACK calls stop at 1.49e-8).  which: 0 = Jb, 1 = Jf.
 *   mode CTB_EXACT_X:      in = x = m/T (real),   out = J(x^2)        Jb_exact / Jf_exact   finiteT.py:40-52,69-81
 *   mode CTB_EXACT_DX:     in = x,                out = dJ/dx          dJb_exact / dJf_exact finiteT.py:104-111
 *   mode CTB_EXACT_THETA:  in = theta (any sign), out = J(theta)       Jb_exact2 / Jf_exact2 finiteT.py:54-66,84-96
 * (negative theta: the real part, log|1 -+ exp(-i sqrt|..|)| below the branch point, as in the reference).  */
#define CTB_EXACT_X 0
#define CTB_EXACT_DX 1
#define CTB_EXACT_THETA 2
CTB_API int ctb_jexact(int which, int mode, const double *d_x, double *d_out, int64_t n, void *stream);

/* generic_potential.Vtot (generic_potential.py:312-337) for model1 (examples/testModel1.py:62-116):
 * pointwise, d_X [n,2], d_T [n] -> d_out [n].  model[8] = l1,l2,mu2,Y1,Y2,n_boson,renormScaleSq,v2 (host). */
/* parts: CTB_V_TOTAL for Vtot, CTB_V_THERMAL alone for generic_potential.V1T_from_X
 * (generic_potential.py:298-310); any OR of the three pieces is allowed. */
#define CTB_V_TREE 1      /* V0   examples/testModel1.py:62-80        */
#define CTB_V_ONELOOP 2   /* V1   generic_potential.py:240-256        */
#define CTB_V_THERMAL 4   /* V1T  generic_potential.py:258-296        */
#define CTB_V_TOTAL 7
CTB_API int ctb_vtot_model1(const double *model8, const ctb_jtable *jb, const double *d_X, const double *d_T,
                            double *d_out, int64_t n, int parts, void *stream);

/* Same on the outer-product grid X_i = x0 + s_i * nhat, T_j:  d_out[i * nT + j]  (row-major [ns, nT]).
 * line4 = x0[0], x0[1], nhat[0], nhat[1] (host). */
CTB_API int ctb_vtot_model1_grid(const double *model8, const double *line4, const ctb_jtable *jb, const double *d_s,
                         int64_t ns, const double *d_T, int64_t nT, double *d_out, void *stream);

/* generic_potential.Vtot / V1T_from_X (generic_potential.py:258-337) for ONE general one-field model
 * (model64: host pointer, the CTB_POT_THERMAL1F row layout; its T entry is ignored) at n (phi_i, T_i) points.
 * rad = (num_boson_dof - sum n_b) pi^4/45 + (num_fermion_dof - sum n_f) 7 pi^4/360 is the field-independent radiation
 * constant of V1T (generic_potential.py:289-295; 0 when include_radiation is off or the totals are None).         */
CTB_API int ctb_vtot_thermal1f(const double *model64, double rad, const ctb_jtable *jb, const ctb_jtable *jf,
                               const double *d_phi, const double *d_T, double *d_out, int64_t n, int parts, void *stream);

/* SingleFieldInstanton.exactSolution(r, phi0, dV, d2V) (tunneling1D.py:294-373) for n argument rows (parity aid):
 * d_in [n, 4] = r, phi0, dV, d2V  ->  d_out [n, 2] = phi(r), phi'(r).  alpha: 1, 2, 3, 4 evaluate the specialised forms
 * the bounce kernels use (closed forms, I0 / I1 / J0 / J1); any other real alpha in [0, 8] the general-order form. */
CTB_API int ctb_exact_solution(double alpha, const double *d_in, double *d_out, int64_t n, void *stream);

/* Potential functor evaluation V(phi_j) for one parameter row (debug / parity aid). */
CTB_API int ctb_potential(int pot_kind, const double *d_params, const ctb_jtable *jb, const ctb_jtable *jf,
                  const double *d_phi, double *d_out, int64_t n, void *stream);

/* V_i(phi_i): one field value per parameter row (d_params [n, nparam], d_phi [n] -> d_out [n]).
 * The batched form of calling the user's V(phi) (tunneling1D.py:132) for n different potentials. */
CTB_API int ctb_potential_batch(int pot_kind, const double *d_params, const ctb_jtable *jb, const ctb_jtable *jf,
                                const double *d_phi, double *d_out, int64_t n, void *stream);

/* Batched bounded 1-D minimisation: x_i = argmin V_i(phi) on [lo_i, hi_i] by scipy.optimize.fminbound's
 * algorithm (the routine findRScale uses, tunneling1D.py:264-265) with xatol_i = xtol_rel * (hi_i - lo_i).
 * Feeds phi_absMin(T) to the finite-temperature scans.  d_vmin may be NULL. */
CTB_API int ctb_minimize_1d(int pot_kind, const double *d_params, const ctb_jtable *jb, const ctb_jtable *jf,
                            const double *d_lo, const double *d_hi, double xtol_rel, int64_t n, double *d_xmin,
                            double *d_vmin, void *stream);

/* ---- batched N-field minimum finder / phase follower -----------------------------------
 * Replaces one scipy.optimize.fmin call per point and temperature in generic_potential.findMinimum
 * (generic_potential.py:477-484), transitionFinder.traceMinimum (transitionFinder.py:34-128) and the polish of both
 * minima before a tunnelling (transitionFinder.py:658-664).  Instance i starts from d_X0[i] at temperature index
 * d_t_begin[i] (NULL: 0) and walks t += d_t_dir[i] (+1 / -1; NULL: +1) to the end of its temperature row
 * d_T[i * T_stride + t] (T_stride = 0: one shared grid; nT = 1 and T_stride = 1: findMinimum(X_i, T_i)).  At every
 * temperature: damped Newton on the reference's 4th-order finite-difference gradient and Hessian of Vtot
 * (helper_functions.py:446-554 with eps = x_eps), warm-started by extrapolating the previous two minima.  The walk of
 * an instance ends at the first temperature whose result is not CTB_MIN_OK (that entry holds the point it ended on);
 * entries never visited are NaN / CTB_MIN_NOT_TRACED.  model_kind: CTB_MODEL_MODEL1 (examples/testModel1.py, two
 * fields): d_params rows of 8 = l1, l2, mu2, Y1, Y2, n_boson, renormScaleSq, v2; CTB_MODEL_THERMAL1F (the general
 * one-field model, one field): rows of CTB_NPARAM_THERMAL1F as for CTB_POT_THERMAL1F (their T entry is ignored; jf
 * required).  Rows are param_stride doubles apart (0: one row for all instances); jf may be NULL for MODEL1.  Outputs: d_X [n, nT, ndim], d_V [n, nT] (may be NULL), d_flag [n, nT], d_iters [n] (Newton
 * iterations, may be NULL).                                                                                         */
#define CTB_MODEL_MODEL1 0
#define CTB_MODEL_THERMAL1F 1
#define CTB_MIN_OK 0             /* converged on a point with a positive definite Hessian                          */
#define CTB_MIN_NOT_CONVERGED 1  /* max_iter iterations (or a non-finite potential)                                 */
#define CTB_MIN_NOT_A_MINIMUM 2  /* converged on a saddle / maximum (Hessian had to be lifted)                      */
#define CTB_MIN_JUMPED 3         /* converged farther than jump_tol (|X0| + 1) from the previous minimum: a different
                                    phase (the traced one has ended, transitionFinder.py:118-126)                  */
#define CTB_MIN_NOT_TRACED 4     /* not visited                                                                     */
typedef struct ctb_min_opts {
    double x_eps;     /* 1e-3: finite-difference step, generic_potential.x_eps (generic_potential.py:104)          */
    double xtol;      /* 1e-9: converged when the Newton step is below xtol (|X| + 1)                               */
    double jump_tol;  /* 0.1                                                                                        */
    int32_t max_iter; /* 60                                                                                         */
    int32_t reserved;
} ctb_min_opts;
CTB_API int ctb_trace_minima(int model_kind, const double *d_params, int64_t param_stride, const ctb_jtable *jb,
                             const ctb_jtable *jf, const double *d_X0, const double *d_T, int64_t T_stride,
                             const int32_t *d_t_begin, const int32_t *d_t_dir, int64_t n, int32_t nT, const ctb_min_opts *opts,
                             double *d_X, double *d_V, int32_t *d_flag, int32_t *d_iters, void *stream);

/* ---- host-side gather plumbing of the multi-GPU partition (no collective on the data path) ----
 * Block-cyclic partition: with blocks of B instances, device r of G owns global blocks r, r + G, r + 2G, ...
 * ctb_copy_blocks moves `height` rows of `width` bytes between two pitched layouts on `stream` (one DMA descriptor:
 * cudaMemcpy2DAsync, direction inferred from the pointers), which is exactly the scatter of a shard out of / the
 * gather of a shard into the GLOBAL host array: host pitch = G * B * elem, device pitch = B * elem, width = B * elem.
 * The host side should be page-locked (ctb_host_register pins an existing allocation, e.g. a shared-memory
 * segment that the other ranks of the box map too; every process that issues copies registers it itself).  */
CTB_API int ctb_copy_blocks(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                            void *stream);
CTB_API int ctb_host_register(void *p, size_t bytes);
CTB_API int ctb_host_unregister(void *p);

/* Register-resident FP64 FMA loop used to measure the FP64 roofline denominator.
 * Runs iters * 16 dependent-chain-free FMAs per thread; returns elapsed ms via *ms_out (synchronous). */
CTB_API int ctb_fp64_peak(int64_t iters, int blocks, int threads, double *d_sink, float *ms_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CTBOUNCE_H */
