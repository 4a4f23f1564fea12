// rk45_instance.cu — one (Model, number type) instance of the two integration kernels of rk45_window.cuh and
// its launcher.  Compiled once per instance (Makefile: -DHLM_INST_MODEL=... -DHLM_INST_T=... -DHLM_INST_NAME=...)
// so that the instances build in parallel; hlm_capi.cu only sees the launchers.
#include <algorithm>
#include <cstdlib>

#include "rk45_window.cuh"

#if !defined(HLM_INST_MODEL) || !defined(HLM_INST_T) || !defined(HLM_INST_NAME)
#error "compile with -DHLM_INST_MODEL=<Model204|Model200|DummyModel> -DHLM_INST_T=<double|float> -DHLM_INST_NAME=<launcher name>"
#endif

namespace hlm {

// schedule 0: tiles (rk45_window_kernel); 1: lane refill; 2: lane refill with the early-leave test (few attempts
// per link per launch); 3: tiles taken from the launch's order (sorted tiles).  The grid is SMs x resident CTAs of the kernel (persistent warps pull work from a
// counter), fewer when there is less work.
// a template so that `if constexpr` really leaves the kernels an instance does not carry uninstantiated
template <class Model, typename T>
static cudaError_t launch_rk45_impl(int schedule, const WindowArgs& a, int sm_count, cudaStream_t stream) {
    static int blocks_per_sm[4] = {0, 0, 0, 0};
    if (schedule < 0 || schedule > 3) return cudaErrorInvalidValue;
    // schedule 3: the tile kernel that follows the launch's order (sorted tiles).  Models with an inflow term have it as
    // their only tile kernel; the others carry it as a second FP64 instance beside the plain one.
    constexpr bool kHasSortedInstance = sizeof(T) == 8 && !Model::HAS_INFLOW;
    if (schedule == 3 && !kHasSortedInstance) schedule = 0;
    // Which kernels this instance carries (each is minutes of ptxas): lane refill for FP64 only — FP32 has no
    // routed use (the implicit fallback is FP64) — and its early-leave variant for the models routed runs use
    // (those with an inflow term).  A schedule the instance lacks falls back to the next simpler one.
    constexpr bool kHasLanes = sizeof(T) == 8;
    constexpr bool kHasEarly = kHasLanes && Model::HAS_INFLOW;
    if (schedule == 2 && !kHasEarly) schedule = 1;
    if (schedule == 1 && !kHasLanes) schedule = 0;
    constexpr int kWarps = HLM_CTA_THREADS / 32;
    // HLM_TUNE_BLOCKS_PER_SM (environment): cap on resident CTAs per SM, for occupancy experiments only
    static const int bps_cap = [] {
        const char* e = std::getenv("HLM_TUNE_BLOCKS_PER_SM");
        return e ? std::max(1, std::atoi(e)) : 1 << 20;
    }();
    auto grid_for = [&](int bps) {
        bps = std::min(bps, bps_cap);
        return (unsigned)std::max<long long>(1, std::min<long long>((a.n_tiles + kWarps - 1) / kWarps, (long long)sm_count * bps));
    };
    auto occupancy = [&](auto kernel, int slot) -> cudaError_t {
        if (blocks_per_sm[slot] != 0) return cudaSuccess;
        const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm[slot], kernel, HLM_CTA_THREADS, 0);
        if (e == cudaSuccess && blocks_per_sm[slot] < 1) blocks_per_sm[slot] = 1;
        return e;
    };
    if (schedule == 0) {
        if (cudaError_t e = occupancy(rk45_window_kernel<Model, T>, 0)) return e;
        rk45_window_kernel<Model, T><<<grid_for(blocks_per_sm[0]), HLM_CTA_THREADS, 0, stream>>>(a);
    } else if (schedule == 3) {
        if constexpr (kHasSortedInstance) {
            if (cudaError_t e = occupancy(rk45_window_kernel<Model, T, true>, 3)) return e;
            rk45_window_kernel<Model, T, true><<<grid_for(blocks_per_sm[3]), HLM_CTA_THREADS, 0, stream>>>(a);
        }
    } else if (schedule == 1) {
        if constexpr (kHasLanes) {
            if (cudaError_t e = occupancy(rk45_lanes_kernel<Model, T, false>, 1)) return e;
            rk45_lanes_kernel<Model, T, false><<<grid_for(blocks_per_sm[1]), HLM_CTA_THREADS, 0, stream>>>(a);
        }
    } else {
        if constexpr (kHasEarly) {
            if (cudaError_t e = occupancy(rk45_lanes_kernel<Model, T, true>, 2)) return e;
            rk45_lanes_kernel<Model, T, true><<<grid_for(blocks_per_sm[2]), HLM_CTA_THREADS, 0, stream>>>(a);
        }
    }
    return cudaGetLastError();
}

cudaError_t HLM_INST_NAME(int schedule, const WindowArgs& a, int sm_count, cudaStream_t stream) {
    return launch_rk45_impl<HLM_INST_MODEL, HLM_INST_T>(schedule, a, sm_count, stream);
}

}  // namespace hlm
