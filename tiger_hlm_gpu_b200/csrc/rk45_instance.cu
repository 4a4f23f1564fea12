// rk45_instance.cu — one (Model, number type) instance of the two integration kernels of rk45_window.cuh and
// its launcher.  Compiled once per instance (Makefile: -DHLM_INST_MODEL=... -DHLM_INST_T=... -DHLM_INST_NAME=...)
// so that the instances build in parallel; hlm_capi.cu only sees the launchers.
#include <algorithm>

#include "rk45_window.cuh"

#if !defined(HLM_INST_MODEL) || !defined(HLM_INST_T) || !defined(HLM_INST_NAME)
#error "compile with -DHLM_INST_MODEL=<Model204|Model200|DummyModel> -DHLM_INST_T=<double|float> -DHLM_INST_NAME=<launcher name>"
#endif

namespace hlm {

// schedule 0: tiles (rk45_window_kernel); 1: lane refill; 2: lane refill with the early-leave test (few attempts
// per link per launch).  The grid is SMs x resident CTAs of the kernel (persistent warps pull work from a
// counter), fewer when there is less work.
cudaError_t HLM_INST_NAME(int schedule, const WindowArgs& a, int sm_count, cudaStream_t stream) {
    using Model = HLM_INST_MODEL;
    using T = HLM_INST_T;
    static int blocks_per_sm[3] = {0, 0, 0};
    if (schedule < 0 || schedule > 2) return cudaErrorInvalidValue;
    if (blocks_per_sm[schedule] == 0) {
        cudaError_t e;
        if (schedule == 0) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[0], rk45_window_kernel<Model, T>, HLM_CTA_THREADS, 0);
        else if (schedule == 1) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[1], rk45_lanes_kernel<Model, T, false>, HLM_CTA_THREADS, 0);
        else e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[2], rk45_lanes_kernel<Model, T, true>, HLM_CTA_THREADS, 0);
        if (e != cudaSuccess) return e;
        if (blocks_per_sm[schedule] < 1) blocks_per_sm[schedule] = 1;
    }
    constexpr int kWarps = HLM_CTA_THREADS / 32;
    const long long grid = std::max<long long>(1, std::min<long long>((a.n_tiles + kWarps - 1) / kWarps, (long long)sm_count * blocks_per_sm[schedule]));
    if (schedule == 0) rk45_window_kernel<Model, T><<<(unsigned)grid, HLM_CTA_THREADS, 0, stream>>>(a);
    else if (schedule == 1) rk45_lanes_kernel<Model, T, false><<<(unsigned)grid, HLM_CTA_THREADS, 0, stream>>>(a);
    else rk45_lanes_kernel<Model, T, true><<<(unsigned)grid, HLM_CTA_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace hlm
