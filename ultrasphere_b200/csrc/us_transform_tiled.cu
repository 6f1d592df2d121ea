// Contiguous fast path — placeholder until the TMA-staged persistent kernels land.
#include "us_common.cuh"
#include "us_params.cuh"

namespace us {
int try_launch_to_cartesian_tiled(const XformParams&, int, cudaStream_t) { return 0; }
int try_launch_from_cartesian_tiled(const XformParams&, int, cudaStream_t) { return 0; }
}  // namespace us
