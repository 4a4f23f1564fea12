// ref_host_harness.cpp — host build of the REFERENCE's own step / dense / rhs functions.
//
// TEST INFRASTRUCTURE ONLY (same rules as oracle_rk45.c).  Built by oracle/Makefile into
// oracle/_ref/libref_host.so from the reference sources where they lie under
// /root/reference/src; no reference source is copied into this repository.
//
// What comes from the reference, unchanged, via #include:
//   rk45_step<Model204>, rk45_dense<Model204>   solver/rk45_step_dense.cuh:33-145,171-244
//   norm_inf_diff, SLOPE_JUMP_THRESH, MIN_STEP_FRACTION   solver/event_detector.cuh:11-53
//   Model204::rhs, Model204::Parameters          models/model_204.hpp:15-115
//   SpatialParams                                I_O/parameters_loader.hpp:19-37
// g++ ignores the device/host attributes that <cuda_runtime.h> expands __device__ /
// __host__ / __global__ to, so those templates compile as plain host functions.
//
// What is written here: the driver loop around them.  The reference's only driver is the
// __global__ kernel solver/rk45_kernel.cu:17-176, which cannot be compiled for the host, so
// ref_host_run() restates that loop (same statements, same order) and calls the reference's
// functions for all arithmetic inside a step.
#include <cuda_runtime.h>

#include <cmath>
#include <cstddef>
#include <cstring>

#include "solver/rk45_step_dense.cuh"
#include "solver/event_detector.cuh"
#include "models/model_204.hpp"

extern "C" {

int ref_host_sizeof_spatial_params() { return (int)sizeof(SpatialParams); }

// One Model204::rhs evaluation (float forcing slice, as the kernel passes it).
void ref_host_rhs204(const void* sp, int sys, const double* y, const float* F, int nForc, double* dydt) {
    Model204::rhs(0.0, y, dydt, Model204::N_EQ, sys, (const SpatialParams*)sp, F, nForc);
}

// k0 + rk45_step, exactly the pair of calls at rk45_kernel.cu:114-116.
void ref_host_step204(const void* sp, int sys, const double* y, double h, double rtol, double atol,
                      const float* F, int nForc, double* y_out, double* err, double* k_out /*[7][5]*/) {
    double k[7][Model204::N_EQ];
    Model204::rhs(0.0, y, k[0], Model204::N_EQ, sys, (const SpatialParams*)sp, F, nForc);
    rk45_step<Model204>(0.0, y, y_out, Model204::N_EQ, h, rtol, atol, err, k, sys, (const SpatialParams*)sp, F, nForc);
    std::memcpy(k_out, k, sizeof(k));
}

void ref_host_dense204(const double* y_n, const double* k_in /*[7][5]*/, double h, double theta, double* y_dense) {
    double k[7][Model204::N_EQ];
    std::memcpy(k, k_in, sizeof(k));
    rk45_dense<Model204>(y_n, k, Model204::N_EQ, h, theta, y_dense);
}

// Loop of rk45_kernel.cu:36-175 around the reference's functions for systems [sys_begin, sys_end)
// (callers thread over disjoint ranges; this image has no libgomp).
// forc: float [forcing][time][system] (rk45_kernel.cu:101-105), dt_h hours, nT samples.
// dense is written in the order retrieve_and_free returns, [sys][q][comp] (rk45_api.hpp:255-267).
void ref_host_run204(const double* prm6, int ns, const double* y0, double t0, double tf, const double* tq,
                     int nq, const void* sp_v, const float* forc, int nForc, const double* dt_h,
                     const long long* nT, double* y_final, double* dense, int* stiff_out,
                     long long* n_accept, long long* n_reject, long long* n_jump, int sys_begin, int sys_end) {
    constexpr int N_EQ = Model204::N_EQ;
    const SpatialParams* d_sp = (const SpatialParams*)sp_v;
    Model204::Parameters P;
    P.initialStep = prm6[0]; P.rtol = prm6[1]; P.atol = prm6[2];
    P.safety = prm6[3]; P.minScale = prm6[4]; P.maxScale = prm6[5];
    for (int sys = sys_begin; sys < sys_end; ++sys) {
        double y[N_EQ], y_next[N_EQ], k45[7][N_EQ], err;
        for (int i = 0; i < N_EQ; ++i) y[i] = y0[(size_t)sys * N_EQ + i];
        int next_q = 0, reject_count = 0;
        bool stiff = false;
        long long na = 0, nr = 0, nj = 0;
        double t = t0, h = P.initialStep;
        while (t < tf && !stiff) {
            if (t + h > tf) h = tf - t;
            float F[16];
            for (int j = 0; j < nForc; ++j) {
                double dt_min = dt_h[j] * 60.0;
                double sampleIdxReal = t / dt_min;
                size_t nS = (size_t)nT[j];
                size_t idx = sampleIdxReal < 0.0 ? 0 : (sampleIdxReal >= nS ? nS - 1 : size_t(sampleIdxReal));
                size_t base = 0;
                for (int kk = 0; kk < j; ++kk) base += (size_t)nT[kk] * size_t(ns);
                F[j] = forc[base + idx * size_t(ns) + sys];
            }
            Model204::rhs(t, y, k45[0], N_EQ, sys, d_sp, F, nForc);
            rk45_step<Model204>(t, y, y_next, N_EQ, h, P.rtol, P.atol, &err, k45, sys, d_sp, F, nForc);
            if (err <= 1.0) {
                reject_count = 0;
                double jump = norm_inf_diff(k45[0], k45[1], N_EQ);
                if (jump > SLOPE_JUMP_THRESH) {
                    h = fmax(h * 0.5, P.initialStep * MIN_STEP_FRACTION);
                    ++nj;
                    continue;
                }
                double t1 = t + h;
                while (next_q < nq && tq[next_q] <= t1) {
                    double tqv = tq[next_q];
                    if (tqv > t && dense) {
                        double th = (tqv - t) / h, yd[N_EQ];
                        rk45_dense<Model204>(y, k45, N_EQ, h, th, yd);
                        for (int c = 0; c < N_EQ; ++c) dense[((size_t)sys * nq + next_q) * N_EQ + c] = yd[c];
                    }
                    ++next_q;
                }
                for (int i = 0; i < N_EQ; ++i) y[i] = y_next[i];
                t = t1;
                ++na;
                double fac = P.safety * pow(1.0 / (err + 1e-16), 0.2);
                h *= fmin(fmax(fac, P.minScale), P.maxScale);
            } else {
                ++reject_count;
                ++nr;
                double fac = P.safety * pow(1.0 / (err + 1e-16), 0.2);
                fac = fmin(fac, 1.0);
                fac = fmin(fmax(fac, P.minScale), P.maxScale);
                h *= fac;
                if (reject_count > 5 || h < (tf - t0) * MIN_STEP_FRACTION) stiff = true;
            }
        }
        if (n_accept) n_accept[sys] = na;
        if (n_reject) n_reject[sys] = nr;
        if (n_jump) n_jump[sys] = nj;
        if (stiff && t < tf) {
            if (stiff_out) stiff_out[sys] = 1;
            continue;
        }
        if (y_final)
            for (int i = 0; i < N_EQ; ++i) y_final[(size_t)sys * N_EQ + i] = y[i];
    }
}

}  // extern "C"
