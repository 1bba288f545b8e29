// ctb_wall.cuh -- WallWithConstFriction (tunneling1D.py:829-999): the bubble-wall profile with a constant friction
// term, solved by over/undershooting in the friction coefficient F.  One thread = one instance; the Cash-Karp
// stepper, the step controller, the dense-output cubic and the end-point root find are the ones of the bounce solver
// (Sfi<Pot, CTB_ALPHA_WALL>: only the equation of motion differs).
#pragma once
#include "ctb_device.cuh"

namespace ctb {

struct WallKnobs {
    double Fguess, Ftol, phitol, rmin, rmax, phi0_rel, phi_eps;
    int npoints;
};

struct WallArgs {
    const double *params, *phi_abs, *phi_meta;
    int64_t n;
    WallKnobs knobs;
    JTab jb, jf;
    ctb_bounce_out out; // d_x receives F; d_S is not written (the wall has no action)
};

// One pass of the integration from (0, y00, y01) with friction s.friction: shooting (T1D:489-536) or, with
// SAVE, dense output on R = linspace(0, rf, npoints) (T1D:578-613).  Returns 0 or a CTB_ST_* error.
template <class Pot, bool SAVE, bool PROFILE>
__device__ int wall_integrate(Sfi<Pot, CTB_ALPHA_WALL> &s, double y00, double y01, double dr0, double drmin, double rmax,
                              double epsfrac, double ea0, double ea1, double &rf, int &ctype, int npoints,
                              const ProfileSink &sink)
{
    const double phi_meta = s.phi_meta;
    const double inv_ea0 = 1.0 / ea0, inv_ea1 = 1.0 / ea1;
    double r_ = 0.0, ya = y00, yb = y01, dr = dr0, da, db;
    s.eom(ya, yb, r_, da, db);
    const double ysign = sgn(ya - phi_meta);
    double yf0 = nan(""), yf1 = nan("");
    // dense output grid
    const double step = SAVE ? rf / (npoints - 1) : 0.0;
    int i_pt = 1;
    double Ri = SAVE ? ((npoints == 2) ? rf : __dadd_rn(__dmul_rn(1.0, step), 0.0)) : 0.0;
    if (SAVE && PROFILE) { sink.R[0] = 0.0; sink.Phi[0] = y00; sink.dPhi[0] = y01; }
    ctype = -1;
    for (;;) {
        double dy0, dy1, drnext;
        int e = s.rkqs(ya, yb, da, db, r_, dr, epsfrac, inv_ea0, inv_ea1, dy0, dy1, drnext);
        if (e) return e;
        if (s.n_rkck >= 8 * CTB_MAX_STEPS) return CTB_ST_MAXSTEPS; // guard, see ctbounce.h (walls: ~1e5 steps to reach rmax)
        double r1, na, nb;
        if (SAVE && dr < drmin) { // T1D:594-597
            na = ya + dy0 * drmin / dr; nb = yb + dy1 * drmin / dr;
            dr = drnext = drmin;
            r1 = r_ + dr;
        } else {
            r1 = r_ + dr; na = ya + dy0; nb = yb + dy1;
        }
        double nda, ndb;
        s.eom(na, nb, r1, nda, ndb);
        if (!SAVE) {
            if (r1 > rmax) return CTB_ST_RMAX;
            else if (dr < drmin) return CTB_ST_DRMIN;
            else if (fabs(na - phi_meta) < 3 * ea0 && fabs(nb) < 3 * ea1) {
                rf = r1; yf0 = na; yf1 = nb; ctype = CT_CONVERGED;
                break;
            } else if (nb * ysign > 0 || (na - phi_meta) * ysign < 0) {
                Cubic f0, f1;
                f0.init(ya, dr * da, na, dr * nda);
                f1.init(yb, dr * db, nb, dr * ndb);
                const bool under = nb * ysign > 0;
                ctype = under ? CT_UNDERSHOOT : CT_OVERSHOOT;
                const Cubic &fr = under ? f1 : f0;
                const double off = under ? 0.0 : phi_meta;
                int berr;
                double xx = cubic_crossing(fr, off, berr);
                if (berr == 1 || berr == 2) return (berr == 2) ? CTB_ST_NONFINITE : CTB_ST_BRENT_SIGN;
                rf = fma(dr, xx, r_);
                yf0 = f0.at(xx); yf1 = f1.at(xx);
                break;
            }
        } else {
            if (r_ < Ri && Ri <= r1) { // T1D:601-606
                Cubic f0, f1;
                f0.init(ya, dr * da, na, dr * nda);
                f1.init(yb, dr * db, nb, dr * ndb);
                while (i_pt < npoints && r_ < Ri && Ri <= r1) {
                    const double t = (Ri - r_) / dr;
                    if (PROFILE) { sink.R[i_pt] = Ri; sink.Phi[i_pt] = f0(t); sink.dPhi[i_pt] = f1(t); }
                    i_pt++;
                    Ri = (i_pt == npoints - 1) ? rf : __dadd_rn(__dmul_rn((double)i_pt, step), 0.0);
                }
            }
            if (i_pt >= npoints) break;
        }
        r_ = r1; ya = na; yb = nb; da = nda; db = ndb;
        dr = drnext;
    }
    if (!SAVE && fabs(yf0 - phi_meta) < 3 * ea0 && fabs(yf1) < 3 * ea1) ctype = CT_CONVERGED;
    return 0;
}

template <class Pot, bool PROFILE>
__global__ void __launch_bounds__(128) wall_kernel(const WallArgs a)
{
    extern __shared__ double ctb_smem[];
    JTab jb = a.jb;
    if (!Pot::analytic) {
        jb = stage_table_compact(a.jb, ctb_smem);
        __syncthreads();
    }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    Sfi<Pot, CTB_ALPHA_WALL> s;
    s.pot.load(a.params + i * Pot::nparam, &jb, &a.jf);
    const double phi_abs = a.phi_abs[i], phi_meta = a.phi_meta[i];
    s.phi_abs = phi_abs; s.phi_meta = phi_meta; s.n_rkck = 0; s.friction = 0.0;
    const WallKnobs &k = a.knobs;
    const Pot &pot = s.pot;
    const double qn = nan("");
    double F = qn, rf = qn, phi_bar = qn, rscale = qn, y00 = qn;
    int status = CTB_ST_NOT_RUN, nshots = 0, npts = 0;
    ProfileSink sink{nullptr, nullptr, nullptr};
    if (PROFILE) {
        sink.R = a.out.d_prof_R + i * a.out.prof_stride;
        sink.Phi = a.out.d_prof_Phi + i * a.out.prof_stride;
        sink.dPhi = a.out.d_prof_dPhi + i * a.out.prof_stride;
    }
    do {
        // __init__ (T1D:128-149) with the overridden findRScale (T1D:838-864)
        const double V_meta = pot.V(phi_meta), V_abs = pot.V(phi_abs);
        if (V_meta <= V_abs) { status = CTB_ST_STABLE; break; }
        {
            const double phi_tol = fabs(phi_meta - phi_abs) * 1e-12;
            double phi1 = phi_meta, phi2 = phi_abs, phi0 = 0.5 * (phi1 + phi2);
            while (fabs(phi1 - phi2) > phi_tol) {
                if (pot.V(phi0) > V_meta) phi1 = phi0; else phi2 = phi0;
                phi0 = 0.5 * (phi1 + phi2);
            }
            phi_bar = phi0;
        }
        {
            const double pA = phi_abs, pB = 0.5 * (phi_bar + phi_meta), pC = phi_meta;
            const double yB = pot.V(pB);
            const double lmda = 2 * ((V_abs - yB) / (pA - pB) - (yB - V_meta) / (pB - pC)) / (pC - pA);
            if (lmda <= 0.0) { status = CTB_ST_NO_BARRIER; break; }
            rscale = CTB_PI / sqrt(lmda);
        }
        s.set_phi_eps(k.phi_eps * fabs(phi_abs - phi_meta));
        // findProfile (T1D:918-999)
        const double rmin = k.rmin * rscale, dr0 = rmin, drmin = 0.01 * rmin, rmax = k.rmax * rscale;
        if (!(rmin > 0.0 && rmin < INFINITY)) { status = CTB_ST_NONFINITE; break; } // guard, see ctbounce.h
        const double delta_phi = phi_meta - phi_abs;
        const double ea0 = fabs(delta_phi * k.phitol), ea1 = fabs(delta_phi / rscale * k.phitol);
        double Fmin = 0.0, Fmax = INFINITY;
        F = !isnan(k.Fguess) ? k.Fguess : (V_meta - V_abs) * rscale / (delta_phi * delta_phi);
        const double Ftol = k.Ftol * F;
        const double d2v_abs = s.d2V(phi_abs);
        double y01 = qn;
        int ctype = -1, err = 0;
        status = CTB_ST_XTOL;
        for (;;) {
            // initialConditions (T1D:866-910): r0 = 0 on the growing exponential about phi_absMin
            const double kk = 0.5 * (sqrt(F * F + 4 * d2v_abs) - F);
            y00 = phi_abs + k.phi0_rel * (phi_meta - phi_abs);
            y01 = kk * (y00 - phi_abs);
            s.friction = F;
            err = wall_integrate<Pot, false, false>(s, y00, y01, dr0, drmin, rmax, k.phitol, ea0, ea1, rf, ctype, 0, sink);
            if (err) break;
            nshots++;
            if (ctype == CT_CONVERGED) { status = CTB_ST_CONVERGED; break; }
            else if (ctype == CT_UNDERSHOOT) { Fmax = F; F = (Fmin == 0.0) ? F / 5.0 : .5 * (Fmin + Fmax); } // F is too high
            else { Fmin = F; F = (Fmax == INFINITY) ? F * 5.0 : .5 * (Fmin + Fmax); }                        // F is too low
            if ((Fmax - Fmin) < Ftol) { status = CTB_ST_XTOL; break; }
        }
        if (err) { status = err; break; }
        // second pass with the current F (T1D:992-997)
        s.friction = F;
        npts = k.npoints;
        err = wall_integrate<Pot, true, PROFILE>(s, y00, y01, dr0, drmin, rmax, k.phitol, ea0, ea1, rf, ctype, npts, sink);
        if (err) { status = err; npts = 0; }
    } while (0);
    if (a.out.d_x) a.out.d_x[i] = F;
    if (a.out.d_phi0) a.out.d_phi0[i] = y00;
    if (a.out.d_r0) a.out.d_r0[i] = (status <= CTB_ST_XTOL) ? 0.0 : qn;
    if (a.out.d_rf) a.out.d_rf[i] = rf;
    if (a.out.d_phi_bar) a.out.d_phi_bar[i] = phi_bar;
    if (a.out.d_rscale) a.out.d_rscale[i] = rscale;
    if (a.out.d_status) a.out.d_status[i] = status;
    if (a.out.d_n_shots) a.out.d_n_shots[i] = nshots;
    if (a.out.d_n_rkck) a.out.d_n_rkck[i] = s.n_rkck;
    if (a.out.d_n_points) a.out.d_n_points[i] = npts;
}

template <class Pot>
int launch_wall_any(const WallArgs &a, cudaStream_t st)
{
    const unsigned grid = (unsigned)((a.n + 127) / 128);
    const size_t sm = pot_smem_bytes<Pot>(a.jb.nint);
    if (a.out.d_prof_R) launch_k(wall_kernel<Pot, true>, grid, 128, sm, st, a);
    else launch_k(wall_kernel<Pot, false>, grid, 128, sm, st, a);
    return (int)cudaGetLastError();
}

int launch_wall_poly4(const WallArgs &a, cudaStream_t st);
int launch_wall_spline(const WallArgs &a, cudaStream_t st);
int launch_wall_model1_line(const WallArgs &a, cudaStream_t st);
int launch_wall_model1f(const WallArgs &a, cudaStream_t st);
int launch_wall_thermal1f(const WallArgs &a, cudaStream_t st);

} // namespace ctb
