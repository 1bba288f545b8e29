// TUNING ONLY (not part of the product build): stubs for the thermal-potential launchers so that
// poly4-only experiment builds link quickly.
#include "ctb_bounce.cuh"
#include "ctb_wall.cuh"
namespace ctb {
int launch_bounce_model1_line(const BounceArgs &, int, int, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_bounce_model1f(const BounceArgs &, int, int, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_setup_model1_line(const BounceArgs &, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_setup_model1f(const BounceArgs &, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_bounce_warp_model1_line(const BounceArgs &, int, int, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_bounce_warp_model1f(const BounceArgs &, int, int, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_bounce_spline(const BounceArgs &, int, int, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_setup_spline(const BounceArgs &, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_bounce_warp_spline(const BounceArgs &, int, int, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_bounce_thermal1f(const BounceArgs &, int, int, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_setup_thermal1f(const BounceArgs &, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_bounce_warp_thermal1f(const BounceArgs &, int, int, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_wall_spline(const WallArgs &, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_wall_model1_line(const WallArgs &, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_wall_model1f(const WallArgs &, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
int launch_wall_thermal1f(const WallArgs &, cudaStream_t) { return CTB_ERR_UNSUPPORTED; }
} // namespace ctb
