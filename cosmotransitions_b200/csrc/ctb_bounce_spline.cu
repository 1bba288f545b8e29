// bounce kernel instantiation for the PotSpline potential functor (see ctb_bounce.cuh)
#include "ctb_bounce.cuh"
#include "ctb_wall.cuh"

namespace ctb {
int launch_bounce_spline(const BounceArgs &a, int alpha, int block, cudaStream_t st)
{
    return launch_bounce_any<PotSpline>(a, alpha, block, st);
}
int launch_setup_spline(const BounceArgs &a, cudaStream_t st) { return launch_setup_any<PotSpline>(a, st); }
int launch_bounce_warp_spline(const BounceArgs &a, int alpha, int stage_cap, cudaStream_t st)
{
    return launch_bounce_warp_any<PotSpline>(a, alpha, stage_cap, st);
}
int launch_wall_spline(const WallArgs &a, cudaStream_t st) { return launch_wall_any<PotSpline>(a, st); }
} // namespace ctb
