// bounce kernel instantiation for the PotPoly4 potential functor (see ctb_bounce.cuh)
#include "ctb_bounce.cuh"
#include "ctb_wall.cuh"

namespace ctb {
int launch_bounce_poly4(const BounceArgs &a, int alpha, int block, cudaStream_t st)
{
    return launch_bounce_any<PotPoly4>(a, alpha, block, st);
}
int launch_setup_poly4(const BounceArgs &a, cudaStream_t st) { return launch_setup_any<PotPoly4>(a, st); }
int launch_bounce_warp_poly4(const BounceArgs &a, int alpha, int stage_cap, cudaStream_t st)
{
    return launch_bounce_warp_any<PotPoly4>(a, alpha, stage_cap, st);
}
int launch_wall_poly4(const WallArgs &a, cudaStream_t st) { return launch_wall_any<PotPoly4>(a, st); }
} // namespace ctb
