// bounce kernel instantiation for the PotModel1Line potential functor (see ctb_bounce.cuh)
#include "ctb_bounce.cuh"
#include "ctb_wall.cuh"

namespace ctb {
int launch_bounce_model1_line(const BounceArgs &a, int alpha, int block, cudaStream_t st)
{
    return launch_bounce_any<PotModel1Line>(a, alpha, block, st);
}
int launch_setup_model1_line(const BounceArgs &a, cudaStream_t st) { return launch_setup_any<PotModel1Line>(a, st); }
int launch_bounce_warp_model1_line(const BounceArgs &a, int alpha, int stage_cap, cudaStream_t st)
{
    return launch_bounce_warp_any<PotModel1Line>(a, alpha, stage_cap, st);
}
int launch_wall_model1_line(const WallArgs &a, cudaStream_t st) { return launch_wall_any<PotModel1Line>(a, st); }
} // namespace ctb
