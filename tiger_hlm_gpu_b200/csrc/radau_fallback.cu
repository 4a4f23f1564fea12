// radau_fallback.cu — translation units of the implicit fallback (radau_fallback.cuh).
//
// Built with -fmad=false: its linear algebra is written as plain C++ expressions and must round exactly
// like the CPU twin the tests compare it with.  The RK45 translation units keep nvcc's default (their
// arithmetic goes through explicit intrinsics, and libdevice's pow must stay the build the reference
// uses), which is why this is a file of its own.  Compiled once per model (-DHLM_INST_MODEL=...
// -DHLM_INST_NAME=...) so the models build side by side, and once without those macros for the dispatcher.
#include <algorithm>

#include "radau_fallback.cuh"

#ifndef HLM_RADAU_WARP
#define HLM_RADAU_WARP 1  // 1: one warp per flagged link (radau_warp_kernel); 0: one thread (radau_window_kernel)
#endif

namespace hlm {

#define HLM_DECLARE_RADAU(name) \
    cudaError_t name(const WindowArgs& a, int* list, unsigned int* n_list, unsigned int* n_radau, int sm_count, cudaStream_t stream)
HLM_DECLARE_RADAU(radau_launch_204);
HLM_DECLARE_RADAU(radau_launch_200);
HLM_DECLARE_RADAU(radau_launch_dummy);

#ifdef HLM_INST_MODEL
HLM_DECLARE_RADAU(HLM_INST_NAME) {
    using Model = HLM_INST_MODEL;
    cudaError_t e = cudaMemsetAsync(n_list, 0, sizeof(unsigned int), stream);
    if (e != cudaSuccess) return e;
    const int tpb = 256;
    const long long lo = a.tile_lo << 5, hi = std::min<long long>(a.ns, (a.tile_lo + a.n_tiles) << 5);
    radau_collect_kernel<<<(unsigned)((hi - lo + tpb - 1) / tpb), tpb, 0, stream>>>(a.status, lo, hi, list, n_list);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    RadauArgs ra;
    ra.w = a;
    ra.list = list;
    ra.n_list = n_list;
    ra.n_radau = n_radau;
    // the list length lives on the device: a fixed grid strides over it (flagged links are rare)
#if HLM_RADAU_WARP
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((hi - lo + 3) / 4, (long long)sm_count * 2));
    radau_warp_kernel<Model><<<grid, 128, 0, stream>>>(ra);
#else
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((hi - lo + 63) / 64, (long long)sm_count * 4));
    radau_window_kernel<Model><<<grid, 64, 0, stream>>>(ra);
#endif
    return cudaGetLastError();
}
#else
cudaError_t radau_launch(int uid, const WindowArgs& a, int* list, unsigned int* n_list, unsigned int* n_radau,
                         int sm_count, cudaStream_t stream) {
    if (uid == Model204::UID) return radau_launch_204(a, list, n_list, n_radau, sm_count, stream);
    if (uid == Model200::UID) return radau_launch_200(a, list, n_list, n_radau, sm_count, stream);
    if (uid == DummyModel::UID) return radau_launch_dummy(a, list, n_list, n_radau, sm_count, stream);
    return cudaErrorInvalidValue;
}
#endif

}  // namespace hlm
