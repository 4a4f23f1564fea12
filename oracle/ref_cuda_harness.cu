// ref_cuda_harness.cu — the UNCHANGED reference kernel, built for sm_100a, behind a C entry point.
//
// TEST INFRASTRUCTURE ONLY (same rules as oracle_rk45.c).  Built by oracle/Makefile into
// oracle/_ref/libref_cuda.so together with the reference's models/model_204_global.cu and
// I_O/forcing_data.cu, from the sources where they lie under /root/reference/src.  No reference
// source is copied into this repository.
//
// The reference kernel rk45_then_radau_multi<Model204> (solver/rk45_kernel.cu:17-176) is pulled in
// TWICE by #include, each time inside its own namespace so the two instantiations do not collide:
//   ref_plain::   the file exactly as it is;
//   ref_counted:: the same file with two call sites re-pointed by macros at counting wrappers:
//                   rk45_step<Model204>(…, sys, …)      -> one attempt        (rk45_kernel.cu:116)
//                   norm_inf_diff(k45[0], k45[1], N_EQ) -> one err<=1 pass,   (rk45_kernel.cu:132)
//                                                          and a jump if > SLOPE_JUMP_THRESH
//                 from which n_accept = errok - jump, n_reject = attempt - errok, n_jump = jump
//                 (the reference itself has no counters).  The wrappers forward to
//                 the reference's own functions, so the arithmetic is untouched;
//                 tests/test_gpu_reference_cuda.py checks ref_counted == ref_plain bit for bit.
// The harness does what src/main.cpp:552-574,633-640,666-703 and rk45_api.hpp:63-116,159-270 do
// (upload, constants, launch, download, reorder) but with the 1-D launch geometry of
// rk45_api.hpp:133-140 — the committed main.cpp:679 geometry integrates only system 0.
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

#include "rk45.h"
#include "rk45_step_dense.cuh"
#include "event_detector.cuh"
#include "small_lu.cuh"
#include "radau_step_dense.cuh"
#include "models/model_204.hpp"
#include "I_O/forcing_data.h"

namespace ref_plain {
#include "solver/rk45_kernel.cu"
}

__device__ int* g_cnt_attempt;
__device__ int* g_cnt_errok;
__device__ int* g_cnt_jump;

template <class M>
__device__ void counted_rk45_step(double t, const double* y, double* y_out, int n, double h, double rtol,
                                  double atol, double* err, double k[7][M::N_EQ], int sys,
                                  const typename M::SP_TYPE* sp, const float* F, int nF) {
    rk45_step<M>(t, y, y_out, n, h, rtol, atol, err, k, sys, sp, F, nF);
    g_cnt_attempt[sys] += 1;
}
__device__ static double counted_norm_inf_diff(const double* a, const double* b, int n, int sys) {
    double v = norm_inf_diff(a, b, n);
    g_cnt_errok[sys] += 1;
    if (v > SLOPE_JUMP_THRESH) g_cnt_jump[sys] += 1;
    return v;
}

namespace ref_counted {
#define rk45_step counted_rk45_step
#define norm_inf_diff(a, b, n) counted_norm_inf_diff(a, b, n, sys)
#include "solver/rk45_kernel.cu"
#undef rk45_step
#undef norm_inf_diff
}

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e_ = (x);                                                         \
        if (e_ != cudaSuccess) {                                                      \
            std::fprintf(stderr, "ref_cuda: %s -> %s\n", #x, cudaGetErrorString(e_)); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

extern "C" int ref_cuda_sizeof_spatial_params() { return (int)sizeof(SpatialParams); }

// Returns 0 on success.  dense_out is [sys][q][comp] (rk45_api.hpp:255-267); slots the kernel never
// writes are 0 (buffers are zero-filled here; the reference leaves them uninitialised).
// cnt_* may be NULL when counted == 0.  kernel_ms receives the CUDA-event time of the launch.
extern "C" int ref_cuda_run204(int counted, const double* prm6, int ns, const double* y0, double t0, double tf,
                               const double* tq, int nq, const void* sp_aos, const float* forc, int nForc_h,
                               const double* dt_h, const unsigned long long* nT, double* final_out,
                               double* dense_out, int* stiff_out, int* cnt_attempt, int* cnt_errok,
                               int* cnt_jump, float* kernel_ms) {
    constexpr int N = Model204::N_EQ;
    if ((long long)ns * N * (long long)nq >= (1LL << 31)) {
        std::fprintf(stderr, "ref_cuda: ns*N_EQ*nq overflows the reference's 32-bit dense index (SURVEY F15)\n");
        return 2;
    }
    Model204::Parameters P;
    P.initialStep = prm6[0]; P.rtol = prm6[1]; P.atol = prm6[2];
    P.safety = prm6[3]; P.minScale = prm6[4]; P.maxScale = prm6[5];
    CK(cudaMemcpyToSymbol(devParams, &P, sizeof(P)));
    size_t nT_h[MAX_FORCINGS] = {0};
    double dt_hh[MAX_FORCINGS] = {0};
    size_t forc_elems = 0;
    for (int j = 0; j < nForc_h; ++j) { nT_h[j] = (size_t)nT[j]; dt_hh[j] = dt_h[j]; forc_elems += nT_h[j] * (size_t)ns; }
    CK(cudaMemcpyToSymbol(c_forc_dt, dt_hh, sizeof(double) * MAX_FORCINGS));
    CK(cudaMemcpyToSymbol(c_forc_nT, nT_h, sizeof(size_t) * MAX_FORCINGS));

    double *d_y0 = nullptr, *d_final = nullptr, *d_tq = nullptr, *d_dense = nullptr;
    int *d_stiff = nullptr, *d_cnt = nullptr;
    float* d_forc = nullptr;
    SpatialParams* d_sp = nullptr;
    size_t bytes_dense = sizeof(double) * (size_t)ns * N * (size_t)(nq > 0 ? nq : 1);
    CK(cudaMalloc(&d_y0, sizeof(double) * ns * N));
    CK(cudaMalloc(&d_final, sizeof(double) * ns * N));
    CK(cudaMalloc(&d_tq, sizeof(double) * (nq > 0 ? nq : 1)));
    CK(cudaMalloc(&d_dense, bytes_dense));
    CK(cudaMalloc(&d_stiff, sizeof(int) * ns));
    CK(cudaMalloc(&d_cnt, sizeof(int) * 3 * ns));
    CK(cudaMalloc(&d_forc, sizeof(float) * (forc_elems ? forc_elems : 1)));
    CK(cudaMalloc(&d_sp, sizeof(SpatialParams) * ns));
    CK(cudaMemset(d_final, 0, sizeof(double) * ns * N));
    CK(cudaMemset(d_dense, 0, bytes_dense));
    CK(cudaMemset(d_stiff, 0, sizeof(int) * ns));
    CK(cudaMemset(d_cnt, 0, sizeof(int) * 3 * ns));
    CK(cudaMemcpy(d_y0, y0, sizeof(double) * ns * N, cudaMemcpyHostToDevice));
    if (nq > 0) CK(cudaMemcpy(d_tq, tq, sizeof(double) * nq, cudaMemcpyHostToDevice));
    if (forc_elems) CK(cudaMemcpy(d_forc, forc, sizeof(float) * forc_elems, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sp, sp_aos, sizeof(SpatialParams) * ns, cudaMemcpyHostToDevice));
    int* p;
    p = d_cnt;          CK(cudaMemcpyToSymbol(g_cnt_attempt, &p, sizeof(p)));
    p = d_cnt + ns;     CK(cudaMemcpyToSymbol(g_cnt_errok, &p, sizeof(p)));
    p = d_cnt + 2 * ns; CK(cudaMemcpyToSymbol(g_cnt_jump, &p, sizeof(p)));

    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int TPB = 128;  // rk45_api.hpp:133-136
    int blocks = (ns + TPB - 1) / TPB;
    CK(cudaEventRecord(e0));
    if (counted)
        ref_counted::rk45_then_radau_multi<Model204><<<blocks, TPB>>>(d_y0, d_final, d_tq, d_dense, ns, nq, t0, tf,
                                                                      d_sp, d_stiff, d_forc, nForc_h);
    else
        ref_plain::rk45_then_radau_multi<Model204><<<blocks, TPB>>>(d_y0, d_final, d_tq, d_dense, ns, nq, t0, tf,
                                                                    d_sp, d_stiff, d_forc, nForc_h);
    CK(cudaEventRecord(e1));
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (kernel_ms) *kernel_ms = ms;

    CK(cudaMemcpy(final_out, d_final, sizeof(double) * ns * N, cudaMemcpyDeviceToHost));
    if (stiff_out) CK(cudaMemcpy(stiff_out, d_stiff, sizeof(int) * ns, cudaMemcpyDeviceToHost));
    if (dense_out && nq > 0) {
        std::vector<double> raw((size_t)ns * N * nq);
        CK(cudaMemcpy(raw.data(), d_dense, bytes_dense, cudaMemcpyDeviceToHost));
        for (int s = 0; s < ns; ++s)
            for (int q = 0; q < nq; ++q)
                for (int c = 0; c < N; ++c)
                    dense_out[((size_t)s * nq + q) * N + c] = raw[(size_t)s * N * nq + (size_t)c * nq + q];
    }
    if (counted && cnt_attempt) {
        CK(cudaMemcpy(cnt_attempt, d_cnt, sizeof(int) * ns, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(cnt_errok, d_cnt + ns, sizeof(int) * ns, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(cnt_jump, d_cnt + 2 * ns, sizeof(int) * ns, cudaMemcpyDeviceToHost));
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_y0); cudaFree(d_final); cudaFree(d_tq); cudaFree(d_dense);
    cudaFree(d_stiff); cudaFree(d_cnt); cudaFree(d_forc); cudaFree(d_sp);
    return 0;
}
