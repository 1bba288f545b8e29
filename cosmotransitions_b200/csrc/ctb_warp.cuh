// ctb_warp.cuh -- warp-per-instance ("latency") variant of the bounce solver.
//
// For batches too small to fill the GPU with one thread per instance (per-temperature scans of a few
// hundred points, the scalar SingleFieldInstanton API) one WARP owns one instance:
//
//  * shooting: the reference's search in x (T1D:702-727) is a deterministic binary decision process --
//    undershoot -> (xmin = x; x = 5x or midpoint), overshoot -> (xmax = x; x = midpoint).  The 31
//    nodes of the next five levels of that decision tree are integrated SPECULATIVELY, one shot per
//    lane; the outcomes are then exchanged with warp shuffles and the tree is walked exactly as the
//    sequential loop would have walked it.  The x values that are actually visited, and therefore
//    every downstream number, are identical to the sequential algorithm; up to five sequential shots
//    collapse into one round.
//  * save pass: all lanes integrate the accepted trajectory redundantly (free under SIMT) and split the
//    dense-output points and the interior points between them; r and the action integrand are staged
//    in shared memory, Simpson's pairs are evaluated lane-parallel and reduced with shuffles.
#pragma once
#include "ctb_device.cuh"
// (included from ctb_bounce.cuh after BounceArgs is defined)

namespace ctb {

enum { SHOT_CONVERGED = 0, SHOT_UNDER = 1, SHOT_OVER = 2, SHOT_NONFINITE = 3, SHOT_ERROR = 4, SHOT_NOT_RUN = 5 };

struct ShootConsts {
    double phi_abs, phi_meta, delta_phi, rmin, dr0, drmin, rmax_rel, ea0, ea1, inv_ea0, inv_ea1, epsfrac, cutoff;
};

struct ShotResult {
    int code, status, nrk;
    double x, r0, y00, y01, rf, delta_phi0;
};

// One trip of the reference's shooting loop for a given x: initialConditions + integrateProfile
// (T1D:703-715, 489-536).  Same arithmetic as the shooting branch of solve_instance.
template <class Pot, int ALPHA>
__device__ void run_shot(Sfi<Pot, ALPHA> &s, const ShootConsts &c, double x, ShotResult &o)
{
    const double qnan = nan("");
    o.code = SHOT_ERROR; o.status = CTB_ST_NOT_RUN; o.x = x;
    o.r0 = o.y00 = o.y01 = o.rf = qnan;
    const int nrk0 = s.n_rkck;
    const double delta_phi0 = exp(-x) * c.delta_phi;
    o.delta_phi0 = delta_phi0;
    const double dv0 = s.dV_from_absMin(delta_phi0);
    const double d2v0 = s.d2V(c.phi_abs + delta_phi0);
    double r0, p0, dp0;
    int e = s.initial_conditions(delta_phi0, dv0, d2v0, c.rmin, c.cutoff, r0, p0, dp0);
    if (e) { o.status = e; o.nrk = 0; return; }
    if (!isfinite(r0) || !isfinite(x)) { o.code = SHOT_NONFINITE; o.nrk = 0; return; }
    o.r0 = r0; o.y00 = p0; o.y01 = dp0;
    double r_ = r0, ya = p0, yb = dp0, dr = c.dr0, da, db;
    s.eom(ya, yb, r_, da, db);
    const double ysign = sgn(ya - c.phi_meta);
    const double rmax = c.rmax_rel + r0;
    double yf0 = qnan, yf1 = qnan, rf = qnan;
    int ctype = -1;
    for (;;) {
        double dy0, dy1, drnext;
        e = s.rkqs(ya, yb, da, db, r_, dr, c.epsfrac, c.inv_ea0, c.inv_ea1, dy0, dy1, drnext);
        if (e) { o.status = e; break; }
        if (s.n_rkck - nrk0 >= CTB_MAX_STEPS) { o.status = CTB_ST_MAXSTEPS; break; } // guard, see ctbounce.h
        const double r1 = r_ + dr, na = ya + dy0, nb = yb + dy1;
        double nda, ndb;
        s.eom_inv(na, nb, s.inv_r_end, nda, ndb);
        if (r1 > rmax) { o.status = CTB_ST_RMAX; break; }
        else if (dr < c.drmin) { o.status = CTB_ST_DRMIN; break; }
        else if (fabs(na - c.phi_meta) < 3 * c.ea0 && fabs(nb) < 3 * c.ea1) {
            rf = r1; yf0 = na; yf1 = nb; ctype = SHOT_CONVERGED;
            break;
        } else if (nb * ysign > 0 || (na - c.phi_meta) * ysign < 0) {
            Cubic f0, f1;
            f0.init(ya, dr * da, na, dr * nda);
            f1.init(yb, dr * db, nb, dr * ndb);
            const bool under = nb * ysign > 0;
            ctype = under ? SHOT_UNDER : SHOT_OVER;
            const Cubic &fr = under ? f1 : f0;
            const double off = under ? 0.0 : c.phi_meta;
            int berr;
            double xx = cubic_crossing(fr, off, berr);
            if (berr == 1 || berr == 2) { o.status = (berr == 2) ? CTB_ST_NONFINITE : CTB_ST_BRENT_SIGN; ctype = -1; break; }
            rf = fma(dr, xx, r_);
            yf0 = f0.at(xx); yf1 = f1.at(xx);
            break;
        }
        r_ = r1; ya = na; yb = nb; da = nda; db = ndb;
        dr = drnext;
    }
    o.nrk = s.n_rkck - nrk0;
    if (ctype < 0) return; // error, status set
    if (fabs(yf0 - c.phi_meta) < 3 * c.ea0 && fabs(yf1) < 3 * c.ea1) ctype = SHOT_CONVERGED;
    o.code = ctype;
    o.rf = rf;
}

CTB_DEV double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// x candidates beyond this are not integrated speculatively: exp(-x) underflows, the start-radius
// search then walks 300 decades before failing, which would stall the whole round.  If the sequential
// walk does reach such a node it becomes the root of the next round and is integrated there.
#define CTB_SPEC_X_MAX 600.0

template <class Pot, int ALPHA, bool PROFILE>
__device__ void solve_instance_warp(Sfi<Pot, ALPHA> &s, const Knobs &k, double phi_bar_in, double rscale_in,
                                    InstanceResult &res, const ProfileSink &sink, double *stage_r, double *stage_y)
{
    const int lane = threadIdx.x & 31;
    const double qnan = nan("");
    res.S = res.phi0 = res.r0 = res.rf = res.x = res.phi_bar = res.rscale = qnan;
    res.n_shots = res.n_rkck = res.n_points = 0;
    res.status = CTB_ST_NOT_RUN;
    s.n_rkck = 0;
    s.phi_eps = 0.0; s.inv_12eps = 0.0;
    const double phi_abs = s.phi_abs, phi_meta = s.phi_meta;
    const Pot &pot = s.pot;

    // ---- __init__ (all lanes, redundantly) ----
    const double V_meta = pot.V(phi_meta);
    double phi_bar = phi_bar_in, rscale = rscale_in;
    {
        int st = sfi_init(pot, phi_abs, phi_meta, V_meta, phi_bar, rscale);
        res.phi_bar = phi_bar;
        res.rscale = rscale;
        if (st) { res.status = st; return; }
    }
    s.set_phi_eps(k.phi_eps * fabs(phi_abs - phi_meta));

    const double xtol = k.xtol, phitol = k.phitol;
    double xmin = xtol * 10, xmax = INFINITY, x;
    if (!isnan(k.xguess)) x = k.xguess;
    else x = -log(fabs((phi_bar - phi_abs) / (phi_meta - phi_abs)));
    ShootConsts c;
    c.phi_abs = phi_abs; c.phi_meta = phi_meta; c.delta_phi = phi_meta - phi_abs;
    c.rmin = k.rmin * rscale; c.dr0 = c.rmin; c.drmin = 0.01 * c.rmin; c.rmax_rel = k.rmax * rscale;
    if (!(c.rmin > 0.0 && c.rmin < INFINITY)) { res.status = CTB_ST_NONFINITE; return; } // guard, see ctbounce.h
    c.ea0 = fabs(c.delta_phi * phitol); c.ea1 = fabs(c.delta_phi / rscale * phitol);
    c.inv_ea0 = 1.0 / c.ea0; c.inv_ea1 = 1.0 / c.ea1; c.epsfrac = phitol;
    c.cutoff = k.thinCutoff * c.delta_phi;
    const int npoints = k.npoints;

    // ---- shooting rounds ----
    bool have_rf = false;
    double r0 = qnan, y00 = qnan, y01 = qnan, rf = qnan, delta_phi0 = qnan, x_last = qnan;
    int status = -1, nshots = 0, nrk_total = 0;
    for (int round = 0; round < 256 && status < 0; round++) {
        // candidate of this lane: node (lane + 1) of the decision tree in heap order; bit 0 = undershoot
        const int node_l = lane + 1;
        double cm = xmin, cM = xmax, cx = x;
        bool runnable = lane < 31;
        {
            const int depth = 31 - __clz(node_l);
            for (int b = depth - 1; b >= 0; --b) {
                if ((node_l >> b) & 1) { cM = cx; cx = .5 * (cm + cM); }
                else { cm = cx; cx = (cM == INFINITY) ? cx * 5.0 : .5 * (cm + cM); }
                if ((cM - cm) < xtol) runnable = false; // the sequential loop stops before this shot
            }
            if (lane > 0 && !(cx <= CTB_SPEC_X_MAX)) runnable = false; // deferred to a later round (see above)
        }
        ShotResult sr;
        sr.code = SHOT_NOT_RUN; sr.status = 0; sr.nrk = 0;
        sr.x = cx; sr.r0 = sr.y00 = sr.y01 = sr.rf = sr.delta_phi0 = qnan;
        if (runnable) run_shot(s, c, cx, sr);
        __syncwarp();
        // walk the tree as the sequential loop would (uniform across lanes)
        int node = 1;
        while (node <= 31 && status < 0) {
            const int src = node - 1;
            const int code = __shfl_sync(0xffffffffu, sr.code, src);
            if (code == SHOT_NOT_RUN) break; // becomes the root of the next round
            const int st = __shfl_sync(0xffffffffu, sr.status, src);
            delta_phi0 = shfl_d(sr.delta_phi0, src);
            nrk_total += __shfl_sync(0xffffffffu, sr.nrk, src);
            if (code == SHOT_NONFINITE) { status = have_rf ? CTB_ST_XTOL : CTB_ST_FIRST_SHOT; break; }
            if (code == SHOT_ERROR) { status = st; break; } // initialConditions / integrateProfile raised
            r0 = shfl_d(sr.r0, src); y00 = shfl_d(sr.y00, src); y01 = shfl_d(sr.y01, src);
            rf = shfl_d(sr.rf, src);
            have_rf = true;
            nshots++;
            x_last = x;
            if (code == SHOT_CONVERGED) { status = CTB_ST_CONVERGED; break; }
            if (code == SHOT_UNDER) { xmin = x; x = (xmax == INFINITY) ? x * 5.0 : .5 * (xmin + xmax); }
            else { xmax = x; x = .5 * (xmin + xmax); }
            if ((xmax - xmin) < xtol) { status = CTB_ST_XTOL; break; }
            node = 2 * node + (code == SHOT_OVER ? 1 : 0);
        }
    }
    s.n_rkck = nrk_total;
    res.n_shots = nshots;
    res.x = x_last;
    if (status != CTB_ST_CONVERGED && status != CTB_ST_XTOL) {
        res.status = status < 0 ? CTB_ST_NOT_RUN : status;
        res.n_rkck = nrk_total;
        return;
    }

    // ---- save pass, lane-parallel: T1D:729-791 ----
    res.r0 = r0; res.rf = rf;
    res.phi0 = phi_abs + delta_phi0;
    const double afac = area_fac<ALPHA>(0.0), vfac = vol_fac<ALPHA>(0.0);
    const double step = (rf - r0) / (npoints - 1);
    const double Rl1 = (npoints == 2) ? rf : __dadd_rn(__dmul_rn(1.0, step), r0);
    int nint = 0;
    double a_int = 0.0, dx0 = Rl1 - r0, st_int = 0.0;
    bool lin_int = true;
    const int max_interior = (k.max_interior_pts < 0) ? npoints / 2 : k.max_interior_pts;
    if (max_interior > 0) {
        if (r0 / dx0 <= max_interior) { nint = (int)ceil(r0 / dx0); st_int = r0 / nint; }
        else { nint = max_interior; lin_int = false; a_int = (r0 / dx0 - nint) * 2 / ((double)nint * (nint + 1)); }
    }
    const int N = nint + npoints;
    res.n_points = N;
    const double phi_c = phi_abs + delta_phi0;
    auto put = [&](int slot, double r, double phi, double dphi) {
        const double ra = pow_alpha<ALPHA>(r, 0.0);
        stage_r[slot] = r;
        stage_y[slot] = (0.5 * (dphi * dphi) + pot.V_quad(phi) - V_meta) * (ra * afac);
        if (PROFILE) { sink.R[slot] = r; sink.Phi[slot] = phi; sink.dPhi[slot] = dphi; }
    };
    if (nint > 0) { // interior points T1D:734-760, one per lane
        const double dv0 = s.dV_from_absMin(delta_phi0);
        const double d2v0 = s.d2V(phi_c);
        const ExactCoef ec = exact_coef(phi_c, dv0, d2v0);
        for (int i = lane; i < nint; i += 32) {
            if (i == 0) { put(0, 0.0, phi_c, 0.0); continue; }
            double ri;
            if (lin_int) ri = __dadd_rn(__dmul_rn((double)i, st_int), 0.0);
            else {
                const double Nv = (double)(nint - i);
                ri = r0 - dx0 * (Nv + 0.5 * a_int * Nv * (Nv + 1));
            }
            const double2 e = exact_solution<ALPHA>(ri, ec);
            put(i, ri, e.x, e.y);
        }
    }
    if (lane == 0) put(nint, r0, y00, y01);
    { // integrateAndSaveProfile T1D:578-613: every lane steps, the points of each step are shared out
        double r_ = r0, ya = y00, yb = y01, dr = c.dr0, da, db;
        s.eom(ya, yb, r_, da, db);
        int i_pt = 1;
        int err = 0;
        int guard = 0;
        while (i_pt < npoints) {
            if (++guard > CTB_MAX_STEPS) { err = CTB_ST_MAXSTEPS; break; }
            double dy0, dy1, drnext;
            err = s.rkqs(ya, yb, da, db, r_, dr, c.epsfrac, c.inv_ea0, c.inv_ea1, dy0, dy1, drnext);
            if (err) break;
            double r1, na, nb;
            if (dr < c.drmin) {
                na = ya + dy0 * c.drmin / dr; nb = yb + dy1 * c.drmin / dr;
                dr = drnext = c.drmin;
                r1 = r_ + dr;
            } else { r1 = r_ + dr; na = ya + dy0; nb = yb + dy1; }
            double nda, ndb;
            s.eom(na, nb, r1, nda, ndb);
            Cubic f0, f1;
            f0.init(ya, dr * da, na, dr * nda);
            f1.init(yb, dr * db, nb, dr * ndb);
            const double inv_dr = rcp_fast(dr);
            for (;;) { // points i_pt, i_pt+1, ... with r_ < R_i <= r1 form a prefix: 32 at a time
                const int kk = i_pt + lane;
                const double Rk = (kk == npoints - 1) ? rf : __dadd_rn(__dmul_rn((double)kk, step), r0);
                const bool in = kk < npoints && r_ < Rk && Rk <= r1;
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (in) {
                    const double t = (Rk - r_) * inv_dr;
                    put(nint + kk, Rk, f0(t), f1(t));
                }
                const int cnt = __popc(m);
                i_pt += cnt;
                if (cnt < 32) break;
            }
            r_ = r1; ya = na; yb = nb; da = nda; db = ndb;
            dr = drnext;
        }
        if (err) { res.status = err; res.n_rkck = s.n_rkck; return; }
    }
    __syncwarp();
    // findAction T1D:780-791: composite Simpson on the staged (r, integrand), pairs shared out over lanes
    const int nbasic = (N & 1) ? N : N - 1;
    double part = 0.0;
    for (int p = lane; 2 * p + 2 < nbasic; p += 32)
        part += SimpsonStream::pair(stage_r[2 * p], stage_y[2 * p], stage_r[2 * p + 1], stage_y[2 * p + 1],
                                    stage_r[2 * p + 2], stage_y[2 * p + 2]);
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    double S = part;
    if (!(N & 1)) {
        if (N == 2) S = 0.5 * (stage_r[1] - stage_r[0]) * (stage_y[1] + stage_y[0]);
        else {
            const double h0 = stage_r[N - 2] - stage_r[N - 3], h1 = stage_r[N - 1] - stage_r[N - 2];
            double num, den, al, be, et;
            num = 2 * h1 * h1 + 3 * h0 * h1; den = 6 * (h1 + h0);
            al = (den != 0) ? num / den : 0.0;
            num = h1 * h1 + 3.0 * h0 * h1; den = 6 * h0;
            be = (den != 0) ? num / den : 0.0;
            num = 1 * h1 * h1 * h1; den = 6 * h0 * (h0 + h1);
            et = (den != 0) ? num / den : 0.0;
            S += al * stage_y[N - 1] + be * stage_y[N - 2] - et * stage_y[N - 3];
        }
    }
    const double R_first = (nint > 0) ? 0.0 : r0, Phi_first = (nint > 0) ? phi_c : y00;
    const double Rd = pow_dim<ALPHA>(R_first, 0.0);
    S += (Rd * vfac) * (pot.V_quad(Phi_first) - V_meta);
    res.S = S;
    res.n_rkck = s.n_rkck;
    res.status = status;
}

// One warp = one instance; 4 warps per block.  Dynamic shared memory: [Jb table (thermal potentials)]
// followed by 2 * stage_cap doubles per warp.
template <class Pot, int ALPHA, bool PROFILE>
__global__ void __launch_bounds__(128, 3) bounce_warp_kernel(const BounceArgs a, int stage_cap)
{
    extern __shared__ double ctb_smem[];
    JTab jb = a.jb;
    double *stage = ctb_smem;
    if (!Pot::analytic) {
        jb = stage_table_compact(a.jb, ctb_smem);
        stage = ctb_smem + CTB_TABLE_DOUBLES_COMPACT(a.jb.nint) + Pot::nslot * CTB_SLOT_STRIDE;
        __syncthreads();
    }
    const int wib = threadIdx.x >> 5;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
    if (w >= a.n) return; // whole warp leaves together
    const int64_t i = a.order ? a.order[w] : w;
    Sfi<Pot, ALPHA> s;
    s.pot.load(a.params + i * Pot::nparam, &jb, &a.jf);
    s.phi_abs = a.phi_abs[i];
    s.phi_meta = a.phi_meta[i];
    InstanceResult r;
    ProfileSink sink{nullptr, nullptr, nullptr};
    if (PROFILE) {
        sink.R = a.out.d_prof_R + i * a.out.prof_stride;
        sink.Phi = a.out.d_prof_Phi + i * a.out.prof_stride;
        sink.dPhi = a.out.d_prof_dPhi + i * a.out.prof_stride;
    }
    const double qn = nan("");
    double *sr = stage + (size_t)wib * 2 * stage_cap;
    solve_instance_warp<Pot, ALPHA, PROFILE>(s, a.knobs, a.phi_bar_in ? a.phi_bar_in[i] : qn,
                                             a.rscale_in ? a.rscale_in[i] : qn, r, sink, sr, sr + stage_cap);
    if ((threadIdx.x & 31) != 0) return;
    if (a.out.d_S) a.out.d_S[i] = r.S;
    if (a.out.d_phi0) a.out.d_phi0[i] = r.phi0;
    if (a.out.d_r0) a.out.d_r0[i] = r.r0;
    if (a.out.d_rf) a.out.d_rf[i] = r.rf;
    if (a.out.d_x) a.out.d_x[i] = r.x;
    if (a.out.d_phi_bar) a.out.d_phi_bar[i] = r.phi_bar;
    if (a.out.d_rscale) a.out.d_rscale[i] = r.rscale;
    if (a.out.d_status) a.out.d_status[i] = r.status;
    if (a.out.d_n_shots) a.out.d_n_shots[i] = r.n_shots;
    if (a.out.d_n_rkck) a.out.d_n_rkck[i] = r.n_rkck;
    if (a.out.d_n_points) a.out.d_n_points[i] = r.n_points;
}

} // namespace ctb
