// ctb_bounce.cuh -- the batched bounce kernel (one thread = one instance) and its launchers.
// Each potential functor is instantiated in its own translation unit (ctb_bounce_<pot>.cu) so that
// the build can compile them in parallel.
#pragma once
#include "ctb_device.cuh"

#ifndef CTB_MAX_BLOCK
#define CTB_MAX_BLOCK 128
#endif
#ifndef CTB_MIN_BLOCKS
#define CTB_MIN_BLOCKS 3
#endif
#ifndef CTB_BIG_OCC
#define CTB_BIG_OCC 4 // 128-thread blocks per SM of the shoot kernel in the throughput regime (128 registers)
#endif

namespace ctb {

struct BounceArgs {
    const double *params, *phi_abs, *phi_meta;
    const double *phi_bar_in, *rscale_in; // optional overrides (NaN entry = compute)
    const int64_t *order;                 // optional permutation: thread t solves instance order[t]
    float *work_key;
    double *scratch;                      // split solve: 3 n doubles (y(r0), y'(r0), delta_phi0 of the accepted shot)                     // setup kernel only
    int64_t n;
    Knobs knobs;
    JTab jb, jf;
    ctb_bounce_out out;
};

// One thread = one instance.  Adjacent instances (adjacent model parameters) share a warp, so that
// lanes follow similar trajectories; blocks are small so that the hardware block scheduler balances
// the 4x spread in work per instance across the 148 SMs.
// Occupancy target (128-thread blocks per SM), measured on B200: the shoot kernel of an
// analytic potential wants 3 blocks (168 registers, no spills) while the batch is small enough for the
// longest warps to set the run time (C2, 1e5 instances: 1.107 vs 1.156 ms) and 4 blocks (128 registers)
// once throughput dominates (C3, 1e6 instances: 10.59 vs 11.05 ms); the save kernel wants 4 in both.
template <class Pot, int MODE, int OCC>
constexpr int bounce_min_blocks()
{
    if (!Pot::analytic) return Pot::min_blocks;
    return MODE == MODE_SHOOT ? OCC : MODE == MODE_SAVE ? 4 : Pot::min_blocks;
}

template <class Pot, int ALPHA, bool PROFILE, int MODE, int OCC = 3>
__global__ void __launch_bounds__(CTB_MAX_BLOCK, bounce_min_blocks<Pot, MODE, OCC>()) bounce_kernel(const BounceArgs a)
{
    // every lane of every warp runs the solver (it votes warp-wide); lanes past the end of the batch
    // carry a dummy instance that is never integrated
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = t < a.n;
    const int64_t i = valid ? (a.order ? a.order[t] : t) : 0;
    // thermal potentials: the Jb table lives in shared memory for the lifetime of the block
    extern __shared__ double ctb_smem[];
    JTab jb = a.jb;
    if (!Pot::analytic) {
        jb = stage_table_compact(a.jb, ctb_smem);
        __syncthreads();
    }
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
    HandOff h{nullptr, nullptr, nullptr};
    // hand-over slots are indexed by the launch-local thread index t (not by the instance): coalesced, and a launch over a
    // slice of a permutation needs only 3 n doubles of its own
    if (MODE != MODE_FUSED) { const int64_t ts = valid ? t : 0; h.y00 = a.scratch + ts; h.y01 = a.scratch + a.n + ts; h.dphi0 = a.scratch + 2 * a.n + ts; }
    const double qn = nan("");
    double phi_bar_in = a.phi_bar_in ? a.phi_bar_in[i] : qn, rscale_in = a.rscale_in ? a.rscale_in[i] : qn;
    if (MODE == MODE_SAVE) { // what the shoot kernel left behind (these outputs are mandatory in split mode)
        r.status = a.out.d_status[i]; r.r0 = a.out.d_r0[i]; r.rf = a.out.d_rf[i];
        r.n_shots = a.out.d_n_shots[i]; r.n_rkck = a.out.d_n_rkck[i];
        r.S = qn; r.n_points = 0;
        phi_bar_in = a.out.d_phi_bar[i]; rscale_in = a.out.d_rscale[i];
    }
    solve_instance<Pot, ALPHA, PROFILE, MODE>(s, a.knobs, valid, phi_bar_in, rscale_in, r, sink, h);
    if (!valid) return;
    if (MODE == MODE_SAVE) {
        if (a.out.d_S) a.out.d_S[i] = r.S;
        a.out.d_status[i] = r.status;
        a.out.d_n_rkck[i] = r.n_rkck;
        if (a.out.d_n_points) a.out.d_n_points[i] = r.n_points;
        return;
    }
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
#include "ctb_warp.cuh"
namespace ctb {

// __init__ only: phi_bar, rscale, status (0 / STABLE / NO_BARRIER) and the work predictor.
template <class Pot>
__global__ void __launch_bounds__(128) setup_kernel(const BounceArgs a)
{
    extern __shared__ double ctb_smem[];
    JTab jb = a.jb;
    if (!Pot::analytic) {
        jb = stage_table_compact(a.jb, ctb_smem);
        __syncthreads();
    }
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    Pot pot;
    pot.load(a.params + i * Pot::nparam, &jb, &a.jf);
    const double pa = a.phi_abs[i], pm = a.phi_meta[i];
    double phi_bar = a.phi_bar_in ? a.phi_bar_in[i] : nan(""), rscale = a.rscale_in ? a.rscale_in[i] : nan("");
    int st = sfi_init(pot, pa, pm, pot.V(pm), phi_bar, rscale);
    if (st) rscale = nan("");
    if (a.out.d_phi_bar) a.out.d_phi_bar[i] = phi_bar;
    if (a.out.d_rscale) a.out.d_rscale[i] = rscale;
    if (a.out.d_status) a.out.d_status[i] = st;
    if (a.work_key) a.work_key[i] = st ? -INFINITY : (float)work_key(pa, pm, phi_bar);
}

template <class Pot>
int launch_setup_any(const BounceArgs &a, cudaStream_t st)
{
    launch_k(setup_kernel<Pot>, (unsigned)((a.n + 127) / 128), 128, pot_smem_bytes<Pot>(a.jb.nint), st, a);
    return (int)cudaGetLastError();
}

template <class Pot>
int launch_bounce_any(const BounceArgs &a, int alpha, int block, cudaStream_t st)
{
    if (block < 32 || block > CTB_MAX_BLOCK || (block & 31)) return CTB_ERR_UNSUPPORTED;
    unsigned grid = (unsigned)((a.n + block - 1) / block);
    const bool prof = a.out.d_prof_R != nullptr;
    const size_t sm = pot_smem_bytes<Pot>(a.jb.nint);
    // split solve (shoot kernel, then save kernel) when the caller provides scratch and all hand-over arrays
    const bool split = a.scratch && !prof && a.out.d_status && a.out.d_r0 && a.out.d_rf && a.out.d_n_shots &&
                       a.out.d_n_rkck && a.out.d_phi_bar && a.out.d_rscale;
    if (alpha < 0 || alpha > 4) return CTB_ERR_UNSUPPORTED;
    if (alpha == CTB_ALPHA_REAL && !Pot::analytic) return CTB_ERR_UNSUPPORTED; // real alpha: polynomial / tabulated functors
    if (split && (alpha == 2 || alpha == 3)) {
#ifdef CTB_FORCE_BIG // tuning builds: -DCTB_FORCE_BIG=0|1 pins the occupancy variant of the shoot kernel
        const bool big = Pot::analytic && CTB_FORCE_BIG;
#else
        const bool big = Pot::analytic && a.n >= 300000; // throughput regime: trade registers for warps
#endif
        if (alpha == 2) {
            if (big) launch_k(bounce_kernel<Pot, 2, false, MODE_SHOOT, CTB_BIG_OCC>, grid, block, sm, st, a);
            else launch_k(bounce_kernel<Pot, 2, false, MODE_SHOOT, 3>, grid, block, sm, st, a);
            launch_k(bounce_kernel<Pot, 2, false, MODE_SAVE>, grid, block, sm, st, a);
        } else {
            if (big) launch_k(bounce_kernel<Pot, 3, false, MODE_SHOOT, CTB_BIG_OCC>, grid, block, sm, st, a);
            else launch_k(bounce_kernel<Pot, 3, false, MODE_SHOOT, 3>, grid, block, sm, st, a);
            launch_k(bounce_kernel<Pot, 3, false, MODE_SAVE>, grid, block, sm, st, a);
        }
    } else {
        // single-kernel form; the only one compiled for alpha = 1 and 4 (d = 2 and 5 dimensions) and for real alpha
#define CTB_LAUNCH_FUSED(AL)                                                                     \
    do {                                                                                         \
        if (prof) launch_k(bounce_kernel<Pot, AL, true, MODE_FUSED>, grid, block, sm, st, a);          \
        else launch_k(bounce_kernel<Pot, AL, false, MODE_FUSED>, grid, block, sm, st, a);              \
    } while (0)
        switch (alpha) {
        case CTB_ALPHA_REAL:
            if constexpr (Pot::analytic) CTB_LAUNCH_FUSED(CTB_ALPHA_REAL);
            break;
        case 1: CTB_LAUNCH_FUSED(1); break;
        case 2: CTB_LAUNCH_FUSED(2); break;
        case 3: CTB_LAUNCH_FUSED(3); break;
        default: CTB_LAUNCH_FUSED(4); break;
        }
#undef CTB_LAUNCH_FUSED
    }
    return (int)cudaGetLastError();
}

// warp-per-instance variant (ctb_warp.cuh): 4 warps per block, stage_cap staged points per warp
template <class Pot>
int launch_bounce_warp_any(const BounceArgs &a, int alpha, int stage_cap, cudaStream_t st)
{
    const unsigned grid = (unsigned)((a.n + 3) / 4);
    const size_t sm = pot_smem_bytes<Pot>(a.jb.nint) + (size_t)4 * 2 * stage_cap * sizeof(double);
    if (sm > 200 * 1024) return CTB_ERR_UNSUPPORTED;
    const bool prof = a.out.d_prof_R != nullptr;
#define CTB_LAUNCH_WARP(AL, PR)                                                                                  \
    do {                                                                                                         \
        cudaFuncSetAttribute(bounce_warp_kernel<Pot, AL, PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); \
        bounce_warp_kernel<Pot, AL, PR><<<grid, 128, sm, st>>>(a, stage_cap);                                   \
    } while (0)
    if (alpha == 2 && !prof) CTB_LAUNCH_WARP(2, false);
    else if (alpha == 3 && !prof) CTB_LAUNCH_WARP(3, false);
    else if (alpha == 2) CTB_LAUNCH_WARP(2, true);
    else if (alpha == 3) CTB_LAUNCH_WARP(3, true);
    else if (alpha == 1 && !prof) CTB_LAUNCH_WARP(1, false);
    else if (alpha == 4 && !prof) CTB_LAUNCH_WARP(4, false);
    else if (alpha == 1) CTB_LAUNCH_WARP(1, true);
    else if (alpha == 4) CTB_LAUNCH_WARP(4, true);
    else return CTB_ERR_UNSUPPORTED;
#undef CTB_LAUNCH_WARP
    return (int)cudaGetLastError();
}

// ---- host side: C-ABI structures -> BounceArgs (shared by ctb_api.cu and the run-time compiled functors) ----
inline int fill_bounce_args(BounceArgs &a, const ctb_bounce_in *in, const ctb_jtable *jb, const ctb_jtable *jf,
                            const ctb_bounce_out *out)
{
    if (!in || !out || in->n < 0) return CTB_ERR_BAD_ARG;
    if (in->n > 0 && (!in->d_params || !in->d_phi_abs || !in->d_phi_meta)) return CTB_ERR_BAD_ARG;
    a.params = in->d_params; a.phi_abs = in->d_phi_abs; a.phi_meta = in->d_phi_meta; a.n = in->n;
    a.phi_bar_in = in->d_phi_bar; a.rscale_in = in->d_rscale; a.order = in->d_order; a.work_key = nullptr;
    a.scratch = in->d_scratch;
    a.jb = to_jtab(jb); a.jf = to_jtab(jf);
    a.out = *out;
    return CTB_OK;
}

struct BounceLaunch { int block, stage_cap; bool warp; };
// findProfile knobs + kernel choice (thread / warp, block size); small_blocks: the quartic functor's 64-thread blocks
inline int fill_bounce_knobs(BounceArgs &a, const ctb_opts *o, const ctb_bounce_out *out, bool small_blocks, BounceLaunch &L)
{
    if (!o) return CTB_ERR_BAD_ARG;
    if (o->npoints < 2 || o->npoints > 4096) return CTB_ERR_UNSUPPORTED;
    if (out->d_prof_R) {
        if (!out->d_prof_Phi || !out->d_prof_dPhi) return CTB_ERR_BAD_ARG;
        int64_t maxint = o->max_interior_pts < 0 ? o->npoints / 2 : o->max_interior_pts;
        if (out->prof_stride < o->npoints + maxint) return CTB_ERR_BAD_ARG;
    }
    a.knobs.xtol = o->xtol; a.knobs.phitol = o->phitol; a.knobs.thinCutoff = o->thinCutoff; a.knobs.rmin = o->rmin;
    a.knobs.rmax = o->rmax; a.knobs.phi_eps = o->phi_eps; a.knobs.xguess = o->xguess;
    a.knobs.npoints = o->npoints; a.knobs.max_interior_pts = o->max_interior_pts;
    a.knobs.alpha_real = o->alpha_real;
    if (o->alpha == 0 && !(o->alpha_real >= 0.0 && o->alpha_real <= 8.0)) return CTB_ERR_UNSUPPORTED;
    // quartic potential, alpha = 2, 64-thread blocks below the throughput regime: finer-grained hand-out of the work-sorted
    // instances (C2: -2 %; alpha = 3 shares of 125 000 instances: 128-thread blocks are 2.5 % faster, tools/shard_timing.py)
    L.block = o->block_threads ? o->block_threads : (small_blocks && a.n < 300000 && o->alpha == 2 ? 64 : 128);
    L.stage_cap = o->npoints + (o->max_interior_pts < 0 ? o->npoints / 2 : o->max_interior_pts);
    L.warp = o->kernel == CTB_KERNEL_WARP;
    if (o->kernel == CTB_KERNEL_AUTO) L.warp = a.n < 4096 && L.stage_cap <= 2048 && o->alpha != 0;
    if (o->kernel < 0 || o->kernel > 2 || (L.warp && L.stage_cap > 2048)) return CTB_ERR_UNSUPPORTED;
    return CTB_OK;
}

int launch_bounce_thermal1f(const BounceArgs &a, int alpha, int block, cudaStream_t st);
int launch_setup_thermal1f(const BounceArgs &a, cudaStream_t st);
int launch_bounce_warp_thermal1f(const BounceArgs &a, int alpha, int stage_cap, cudaStream_t st);
int launch_bounce_spline(const BounceArgs &a, int alpha, int block, cudaStream_t st);
int launch_setup_spline(const BounceArgs &a, cudaStream_t st);
int launch_bounce_warp_spline(const BounceArgs &a, int alpha, int stage_cap, cudaStream_t st);
int launch_bounce_poly4(const BounceArgs &a, int alpha, int block, cudaStream_t st);
int launch_bounce_warp_poly4(const BounceArgs &a, int alpha, int stage_cap, cudaStream_t st);
int launch_bounce_warp_model1_line(const BounceArgs &a, int alpha, int stage_cap, cudaStream_t st);
int launch_bounce_warp_model1f(const BounceArgs &a, int alpha, int stage_cap, cudaStream_t st);
int launch_bounce_model1_line(const BounceArgs &a, int alpha, int block, cudaStream_t st);
int launch_bounce_model1f(const BounceArgs &a, int alpha, int block, cudaStream_t st);
int launch_setup_poly4(const BounceArgs &a, cudaStream_t st);
int launch_setup_model1_line(const BounceArgs &a, cudaStream_t st);
int launch_setup_model1f(const BounceArgs &a, cudaStream_t st);

} // namespace ctb
