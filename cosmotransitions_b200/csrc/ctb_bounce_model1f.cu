// bounce kernel instantiation for the PotModel1F potential functor (see ctb_bounce.cuh)
#include "ctb_bounce.cuh"
#include "ctb_wall.cuh"

namespace ctb {
int launch_bounce_model1f(const BounceArgs &a, int alpha, int block, cudaStream_t st)
{
    return launch_bounce_any<PotModel1F>(a, alpha, block, st);
}
int launch_setup_model1f(const BounceArgs &a, cudaStream_t st) { return launch_setup_any<PotModel1F>(a, st); }
int launch_bounce_warp_model1f(const BounceArgs &a, int alpha, int stage_cap, cudaStream_t st)
{
    return launch_bounce_warp_any<PotModel1F>(a, alpha, stage_cap, st);
}
int launch_wall_model1f(const WallArgs &a, cudaStream_t st) { return launch_wall_any<PotModel1F>(a, st); }
} // namespace ctb
