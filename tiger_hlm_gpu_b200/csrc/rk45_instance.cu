// rk45_instance.cu — one (Model, number type) instance of the two integration kernels of rk45_window.cuh and
// its launcher.  Compiled once per instance (Makefile: -DHLM_INST_MODEL=... -DHLM_INST_T=... -DHLM_INST_NAME=...)
// so that the instances build in parallel; hlm_capi.cu only sees the launchers.
#include <algorithm>

#include "rk45_window.cuh"

#if !defined(HLM_INST_MODEL) || !defined(HLM_INST_T) || !defined(HLM_INST_NAME)
#error "compile with -DHLM_INST_MODEL=<Model204|Model200|DummyModel> -DHLM_INST_T=<double|float> -DHLM_INST_NAME=<launcher name>"
#endif

namespace hlm {

// lanes: the lane-refill schedule (rk45_lanes_kernel) instead of tiles (rk45_window_kernel).  The grid is
// SMs x resident CTAs of the kernel (persistent warps pull work from a counter), fewer when there is less work.
cudaError_t HLM_INST_NAME(bool lanes, const WindowArgs& a, int sm_count, cudaStream_t stream) {
    using Model = HLM_INST_MODEL;
    using T = HLM_INST_T;
    static int blocks_per_sm[2] = {0, 0};
    if (blocks_per_sm[lanes] == 0) {
        const cudaError_t e = lanes ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[1], rk45_lanes_kernel<Model, T>, 128, 0)
                                    : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[0], rk45_window_kernel<Model, T>, 128, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm[lanes] < 1) blocks_per_sm[lanes] = 1;
    }
    const long long grid = std::max<long long>(1, std::min<long long>((a.n_tiles + 3) / 4, (long long)sm_count * blocks_per_sm[lanes]));
    if (lanes) rk45_lanes_kernel<Model, T><<<(unsigned)grid, 128, 0, stream>>>(a);
    else rk45_window_kernel<Model, T><<<(unsigned)grid, 128, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace hlm
