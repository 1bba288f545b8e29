// phase-follower kernel instantiation (see ctb_phases.cuh); the extern "C" entry point is in ctb_api.cu
#include "ctb_phases.cuh"

namespace ctb {
int launch_trace_minima_model1(const double *d_params, int64_t param_stride, const JTab &jb, const double *d_X0,
                               const double *d_T, int64_t T_stride, const int32_t *d_t_begin, const int32_t *d_t_dir,
                               int64_t n, int nT, const MinKnobs &k, double *d_X, double *d_V, int32_t *d_flag,
                               int32_t *d_iters, cudaStream_t st)
{
    JTab none = jb;
    none.nint = 0;
    return launch_trace_minima<Model1XY>(d_params, param_stride, jb, none, d_X0, d_T, T_stride, d_t_begin, d_t_dir, n, nT, k, d_X,
                                         d_V, d_flag, d_iters, st);
}
int launch_trace_minima_thermal1f(const double *d_params, int64_t param_stride, const JTab &jb, const JTab &jf,
                                  const double *d_X0, const double *d_T, int64_t T_stride, const int32_t *d_t_begin,
                                  const int32_t *d_t_dir, int64_t n, int nT, const MinKnobs &k, double *d_X, double *d_V,
                                  int32_t *d_flag, int32_t *d_iters, cudaStream_t st)
{
    return launch_trace_minima<Thermal1FX>(d_params, param_stride, jb, jf, d_X0, d_T, T_stride, d_t_begin, d_t_dir, n, nT, k, d_X,
                                           d_V, d_flag, d_iters, st);
}
} // namespace ctb
