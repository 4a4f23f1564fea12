// radau_fallback.cuh — implicit re-integration of the links the RK45 path flagged stiff.
//
// The reference's run_rk45 hands flagged systems to radau_kernel_multi (solver/rk45_api.hpp:198-247,
// solver/radau_kernel.cu:20-140, solver/radau_step_dense.cuh:39-208).  That code is unfinished in ways
// that make its numbers meaningless (the whole forcing array is passed where the per-step
// slice belongs, the dense interpolant reads a never-filled buffer, the "embedded" weights b_alt do not
// sum to 1 so the error estimate is O(h), the 15x15 elimination has no pivoting).  What is kept here is
// its contract and structure — 3-stage Radau IIA of order 5 in the reference's slope form
// K_s = f(y + h sum_j A_sj K_j), one thread (radau_window_kernel) or one warp (radau_warp_kernel, the one
// launched) per flagged link, finite-difference Jacobian
// (approx_jacobian), forcings sampled at the step-start time and held over the step (as the RK45
// kernel does), tolerances and step-scale limits from Model::Parameters, dense output at the
// query times, final state written when tf is reached — and what is replaced is the numerics:
//   * simplified Newton on the 15x15 system I - h (A (x) J) with the Jacobian of the step start,
//     LU with partial pivoting, convergence measured against the error tolerance;
//   * the error estimate of Hairer & Wanner (Solving ODEs II, IV.8) as SciPy's Radau implements it:
//     err = (gamma0/h I - J)^-1 (f(t,y) + (e1 Z1 + e2 Z2 + e3 Z3)/h), Z_s = Y_s - y, refined once after
//     a rejection; the norm is the RK45 path's max_i |err_i| / (atol + rtol max(|y_i|, |y_new_i|));
//   * the step controller is the reference's form (safety * err^-1/p clamped to [minScale, maxScale],
//     capped at 1 after a rejection) with p = 4, the order of the estimate;
//   * dense output is the collocation polynomial itself (cubic through 0, Z1, Z2, Z3 at 0, c1, c2, 1).
// Unlike the reference the link is not restarted from t0: it continues from the last state the RK45
// path accepted, so the dense records already written stay and nothing is integrated twice.
// Parity: there is nothing in the reference to be identical to; the pin is SciPy's
// solve_ivp(method="Radau") within the solver tolerance and a CPU restatement kept with the test
// infrastructure (tests/test_gpu_radau.py).
#pragma once
#include "rk45_window.cuh"

namespace hlm {

namespace radau {
constexpr double S6 = 2.449489742783178098197284074705891;  // sqrt(6)
#define HLM_RADAU_TABLES                                                                                              \
    static constexpr double C1 = (4.0 - S6) / 10.0, C2 = (4.0 + S6) / 10.0;                                           \
    /* solver/radau_step_dense.cuh:60-64 */                                                                           \
    static constexpr double A[3][3] = {{(88.0 - 7.0 * S6) / 360.0, (296.0 - 169.0 * S6) / 1800.0, (-2.0 + 3.0 * S6) / 225.0}, \
                                       {(296.0 + 169.0 * S6) / 1800.0, (88.0 + 7.0 * S6) / 360.0, (-2.0 - 3.0 * S6) / 225.0}, \
                                       {(16.0 - S6) / 36.0, (16.0 + S6) / 36.0, 1.0 / 9.0}};                          \
    static constexpr double E[3] = {(-13.0 - 7.0 * S6) / 3.0, (-13.0 + 7.0 * S6) / 3.0, -1.0 / 3.0};                  \
    /* real eigenvalue of A^-1: 3 + 3^(2/3) - 3^(1/3) */                                                              \
    static constexpr double MU_REAL = 3.637834252744495732;
constexpr int kNewtonMaxIter = 8;

// LU with partial pivoting of an n x n row-major matrix; false if singular
template <int n> __device__ inline bool lu_factor(double (&M)[n][n], int (&piv)[n]) {
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(M[k][k]);
        for (int i = k + 1; i < n; ++i)
            if (fabs(M[i][k]) > best) { best = fabs(M[i][k]); p = i; }
        piv[k] = p;
        if (!(best > 0.0)) return false;
        if (p != k)
            for (int j = 0; j < n; ++j) { const double tmp = M[k][j]; M[k][j] = M[p][j]; M[p][j] = tmp; }
        const double inv = 1.0 / M[k][k];
        for (int i = k + 1; i < n; ++i) {
            const double m = M[i][k] * inv;
            M[i][k] = m;
            for (int j = k + 1; j < n; ++j) M[i][j] -= m * M[k][j];
        }
    }
    return true;
}
template <int n> __device__ inline void lu_solve(const double (&M)[n][n], const int (&piv)[n], double (&b)[n]) {
    for (int k = 0; k < n; ++k) {
        const double tmp = b[k]; b[k] = b[piv[k]]; b[piv[k]] = tmp;
        for (int i = k + 1; i < n; ++i) b[i] -= M[i][k] * b[k];
    }
    for (int i = n - 1; i >= 0; --i) {
        double s = b[i];
        for (int j = i + 1; j < n; ++j) s -= M[i][j] * b[j];
        b[i] = s / M[i][i];
    }
}
}  // namespace radau

struct RadauArgs {
    WindowArgs w;            // the window being completed (state columns, forcings, queries, dense buffer)
    const int* list;         // indices of the flagged links
    const unsigned int* n_list;
    unsigned int* n_radau;   // [ld] accepted implicit steps per link
};

// status kStiff -> list (order is arbitrary; every link's result depends on that link alone)
static __global__ void radau_collect_kernel(const int* __restrict__ status, long long lo, long long hi, int* __restrict__ list,
                                     unsigned int* __restrict__ n_list) {
    const long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < hi && (status[i] == kStiff || status[i] == kStiffPaused)) list[atomicAdd(n_list, 1u)] = (int)i;
}

template <class Model> __global__ void __launch_bounds__(64) radau_window_kernel(const RadauArgs ra) {
    using namespace radau;
    HLM_RADAU_TABLES
    constexpr int N = Model::N_EQ;
    const WindowArgs& a = ra.w;
    const unsigned int n_list = *ra.n_list;
    const bool run_to_end = (a.q_hi >= a.nq);
    const double rtol = a.prm.rtol, atol = a.prm.atol;
    for (unsigned int item = blockIdx.x * blockDim.x + threadIdx.x; item < n_list; item += gridDim.x * blockDim.x) {
        const long long sys = ra.list[item];
        double y[N];
        for (int i = 0; i < N; ++i) y[i] = a.y[(long long)i * a.ld + sys];
        double t = a.t[sys], h = a.h[sys];
        int next_q = a.next_q[sys];
        unsigned int n_rej = a.n_reject[sys], n_imp = ra.n_radau[sys];
        typename Model::template Link<double> L;
        L.load(a.sp, a.ld, sys);
        if constexpr (Model::HAS_INFLOW) L.set_inflow(a.qin ? __ldg(a.qin + sys) : 0.0);
        const long long col = (Model::N_FORC > 0 && a.n_forc > 0) ? (a.col ? (long long)a.col[sys] : sys) : 0;
        // the RK45 path leaves h below its stiffness threshold; start from the configured initial step
        if (a.status[sys] == kStiff && (!(h > 0.0) || h < a.prm.initialStep)) h = a.prm.initialStep;
        int status = kStiffPaused;
        long long budget = a.max_attempts > 0 ? a.max_attempts : 0x7fffffffffffffffLL;
        bool unused = false;

        while (true) {
            if (!(t < a.tf)) { status = kDoneStiff; break; }
            if (!run_to_end && next_q >= a.q_hi) break;  // window complete: pause, stay flagged
            if (budget-- <= 0 || !(h > 1e-13 * fmax(1.0, fabs(t)))) { status = kStalled; break; }
            if (t + h > a.tf) h = a.tf - t;

            double F[2] = {0.0, 0.0};
            if (Model::N_FORC > 0)
                for (int j = 0; j < Model::N_FORC && j < a.n_forc; ++j) {
                    double lo, hi;
                    const long long idx = forcing_index(t, a.forc_dt_min[j], a.forc_nT[j], lo, hi);
                    long long r = idx - a.forc_i0[j];
                    r = r < 0 ? 0 : (r >= a.forc_nres[j] ? a.forc_nres[j] - 1 : r);
                    F[j] = (double)__ldg(a.forc[j] + r * a.forc_ncols + col);
                }

            // Jacobian at the step start by forward differences (solver/radau_step_dense.cuh:13-31)
            double f0[N], J[N][N];
            Model::template rhs<double, false>(y, F, L, f0, unused);
            {
                const double eps = 1.4901161193847656e-08;  // sqrt(2^-52)
                double yp[N], f1[N];
                for (int j = 0; j < N; ++j) {
                    for (int i = 0; i < N; ++i) yp[i] = y[i];
                    const double d = eps * fmax(1.0, fabs(y[j]));
                    yp[j] = y[j] + d;
                    const double dj = yp[j] - y[j];
                    Model::template rhs<double, false>(yp, F, L, f1, unused);
                    for (int i = 0; i < N; ++i) J[i][j] = (f1[i] - f0[i]) / dj;
                }
            }
            double M[3 * N][3 * N];
            int piv[3 * N];
            for (int s = 0; s < 3; ++s)
                for (int i = 0; i < N; ++i)
                    for (int r = 0; r < 3; ++r)
                        for (int j = 0; j < N; ++j)
                            M[s * N + i][r * N + j] = ((s == r && i == j) ? 1.0 : 0.0) - h * A[s][r] * J[i][j];
            double Mr[N][N];
            int pivr[N];
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < N; ++j) Mr[i][j] = ((i == j) ? MU_REAL / h : 0.0) - J[i][j];
            bool ok = lu_factor<3 * N>(M, piv) && lu_factor<N>(Mr, pivr);

            // simplified Newton for the stage slopes K
            double K[3][N], Y[3][N];
            for (int s = 0; s < 3; ++s)
                for (int i = 0; i < N; ++i) K[s][i] = f0[i];
            bool converged = false;
            double prev = -1.0;
            for (int it = 0; ok && it < kNewtonMaxIter && !converged; ++it) {
                double G[3 * N];
                for (int s = 0; s < 3; ++s) {
                    for (int i = 0; i < N; ++i) Y[s][i] = y[i] + h * (A[s][0] * K[0][i] + A[s][1] * K[1][i] + A[s][2] * K[2][i]);
                    double fs[N];
                    Model::template rhs<double, false>(Y[s], F, L, fs, unused);
                    for (int i = 0; i < N; ++i) G[s * N + i] = fs[i] - K[s][i];
                }
                lu_solve<3 * N>(M, piv, G);
                double norm = 0.0;
                for (int s = 0; s < 3; ++s)
                    for (int i = 0; i < N; ++i) {
                        K[s][i] += G[s * N + i];
                        const double v = fabs(h * G[s * N + i]) / (atol + rtol * fabs(y[i]));
                        if (v > norm || !(v == v)) norm = v;
                    }
                if (!(norm == norm) || (prev >= 0.0 && norm > 2.0 * prev && norm > 1.0)) break;  // diverging
                converged = norm < 0.03;
                prev = norm;
            }
            if (!converged) {  // Newton failure: halve the step (counts as a rejection)
                h *= 0.5;
                ++n_rej;
                continue;
            }
            // stage states with the converged slopes; y_new = Y_3 (stiffly accurate)
            double Z[3][N], y_new[N];
            for (int s = 0; s < 3; ++s)
                for (int i = 0; i < N; ++i) Z[s][i] = h * (A[s][0] * K[0][i] + A[s][1] * K[1][i] + A[s][2] * K[2][i]);
            for (int i = 0; i < N; ++i) y_new[i] = y[i] + Z[2][i];

            double ze[N], e[N];
            for (int i = 0; i < N; ++i) {
                ze[i] = (E[0] * Z[0][i] + E[1] * Z[1][i] + E[2] * Z[2][i]) / h;
                e[i] = f0[i] + ze[i];
            }
            lu_solve<N>(Mr, pivr, e);
            auto err_norm = [&](const double (&ev)[N]) {
                double m = 0.0;
                for (int i = 0; i < N; ++i) {
                    const double v = fabs(ev[i] / (atol + rtol * fmax(fabs(y[i]), fabs(y_new[i]))));
                    if (v > m || !(v == v)) m = v;
                }
                return m;
            };
            double err = err_norm(e);
            if (err > 1.0) {  // one refinement, as SciPy / RADAU5 do before rejecting
                double yp[N], f1[N];
                for (int i = 0; i < N; ++i) yp[i] = y[i] + e[i];
                Model::template rhs<double, false>(yp, F, L, f1, unused);
                for (int i = 0; i < N; ++i) e[i] = f1[i] + ze[i];
                lu_solve<N>(Mr, pivr, e);
                err = err_norm(e);
            }
            double fac = a.prm.safety * sqrt(sqrt(1.0 / (err + 1e-16)));  // err^(-1/4) from IEEE operations only
            if (!(err == err)) { err = 2.0; fac = a.prm.minScale; }  // non-finite stage: treat as rejected

            if (err <= 1.0) {
                const double t1 = t + h;
                bool overshoot = false;
                while (next_q < a.nq) {
                    const double tq = a.tq[next_q];
                    if (!(tq <= t1)) break;
                    if (next_q >= a.q_hi) { overshoot = true; break; }
                    if (tq > t && a.dense != nullptr) {
                        const double th = (tq - t) / h;
                        long long out = dense_base(a, sys, next_q);
                        for (int i = 0; i < N; ++i) {
                            if (!((a.dense_mask >> i) & 1u)) continue;
                            // collocation cubic through (0,0), (c1,Z1), (c2,Z2), (1,Z3): Newton form
                            const double d1 = Z[0][i] / C1, d2 = (Z[1][i] - Z[0][i]) / (C2 - C1), d3 = (Z[2][i] - Z[1][i]) / (1.0 - C2);
                            const double dd1 = (d2 - d1) / C2, dd2 = (d3 - d2) / (1.0 - C1);
                            const double ddd = dd2 - dd1;
                            dense_put(a, out++, y[i] + th * (d1 + (th - C1) * (dd1 + (th - C2) * ddd)));
                        }
                    }
                    ++next_q;
                }
                if (overshoot) break;  // the step reaches past this window's buffer: redo it next window
                for (int i = 0; i < N; ++i) y[i] = y_new[i];
                t = t1;
                ++n_imp;
                h *= fmin(a.prm.maxScale, fmax(a.prm.minScale, fac));
            } else {
                ++n_rej;
                h *= fmin(a.prm.maxScale, fmax(a.prm.minScale, fmin(1.0, fac)));
            }
        }
        for (int i = 0; i < N; ++i) a.y[(long long)i * a.ld + sys] = y[i];
        a.t[sys] = t;
        a.h[sys] = h;
        a.next_q[sys] = next_q;
        a.status[sys] = status;
        a.n_reject[sys] = n_rej;
        ra.n_radau[sys] = n_imp;
        if constexpr (Model::HAS_INFLOW) {
            if (status != kStiffPaused) route_publish(a, sys, y[0]);
        }
    }
}


// ---------------------------------------------------------------------------------------------------------
// The same integrator with one WARP per flagged link.  Flagged links are few and each is a serial chain of
// implicit steps, so the launch lasts as long as its slowest link; with one thread per link every 15x15
// elimination walks 225 doubles of local memory alone.  Here the 15 rows of the Newton matrix live in the
// registers of 15 lanes (the 5 rows of the error-estimate matrix in 5), the five Jacobian columns and the
// three stage slopes are evaluated by different lanes at once, and pivots/rows/solutions travel by shuffle.
// Every arithmetic operation is the one the one-thread kernel (and its CPU twin) performs on the same
// operands in the same order — elimination updates of different rows are independent, pivot search scans the
// replicated column in the twin's order, sums keep their order — so results are bit-identical.
// ---------------------------------------------------------------------------------------------------------
namespace radau {

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// rows of an n x n matrix on lanes 0..n-1 (row[j] = M[lane][j]); piv is replicated
template <int n> __device__ __forceinline__ bool warp_lu_factor(double (&row)[n], int (&piv)[n], int lane) {
#pragma unroll
    for (int k = 0; k < n; ++k) {
        int p = k;
        double best = fabs(shfl_d(row[k], k));
#pragma unroll
        for (int i = k + 1; i < n; ++i) {
            const double v = fabs(shfl_d(row[k], i));
            if (v > best) { best = v; p = i; }
        }
        piv[k] = p;
        if (!(best > 0.0)) return false;
        if (p != k) {  // warp-uniform
            const int src = lane == k ? p : (lane == p ? k : lane);
#pragma unroll
            for (int j = 0; j < n; ++j) row[j] = shfl_d(row[j], src);
        }
        const double inv = 1.0 / shfl_d(row[k], k);
        const double m = row[k] * inv;
        const bool below = lane > k && lane < n;
#pragma unroll
        for (int j = k + 1; j < n; ++j) {
            const double pj = shfl_d(row[j], k);
            if (below) row[j] -= m * pj;
        }
        if (below) row[k] = m;
    }
    return true;
}

// b = this lane's element of the right-hand side; x = the solution, replicated in every lane
template <int n>
__device__ __forceinline__ void warp_lu_solve(const double (&row)[n], const int (&piv)[n], double b, double (&x)[n], int lane) {
#pragma unroll
    for (int k = 0; k < n; ++k) {
        const int p = piv[k];
        if (p != k) {  // warp-uniform
            const int src = lane == k ? p : (lane == p ? k : lane);
            b = shfl_d(b, src);
        }
        const double bk = shfl_d(b, k);
        if (lane > k && lane < n) b -= row[k] * bk;
    }
#pragma unroll
    for (int i = n - 1; i >= 0; --i) {
        double s = b;
#pragma unroll
        for (int j = i + 1; j < n; ++j) s -= row[j] * x[j];
        s = s / row[i];
        x[i] = shfl_d(s, i);
    }
}

// v[idx] for a lane-dependent idx without dynamic register indexing
template <int n> __device__ __forceinline__ double pick(const double (&v)[n], int idx) {
    double r = v[0];
#pragma unroll
    for (int i = 1; i < n; ++i)
        if (idx == i) r = v[i];
    return r;
}
}  // namespace radau

template <class Model> __global__ void __launch_bounds__(128) radau_warp_kernel(const RadauArgs ra) {
    using namespace radau;
    HLM_RADAU_TABLES
    constexpr int N = Model::N_EQ;
    constexpr int N3 = 3 * N;
    static_assert(N3 <= 32, "one warp holds the rows of the Newton matrix");
    const WindowArgs& a = ra.w;
    const int lane = threadIdx.x & 31;
    const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const unsigned int n_list = *ra.n_list;
    const bool run_to_end = (a.q_hi >= a.nq);
    const double rtol = a.prm.rtol, atol = a.prm.atol;
    // this lane's row of the 15 x 15 matrix is (stage rs, component ri); as a row of the 5 x 5 matrix, component lane
    const int rs = lane < N3 ? lane / N : 0, ri = lane < N3 ? lane % N : 0;
    for (unsigned int item = warp; item < n_list; item += n_warps) {
        const long long sys = ra.list[item];
        double y[N];
        for (int i = 0; i < N; ++i) y[i] = a.y[(long long)i * a.ld + sys];
        double t = a.t[sys], h = a.h[sys];
        int next_q = a.next_q[sys];
        unsigned int n_rej = a.n_reject[sys], n_imp = ra.n_radau[sys];
        typename Model::template Link<double> L;
        L.load(a.sp, a.ld, sys);
        if constexpr (Model::HAS_INFLOW) L.set_inflow(a.qin ? __ldg(a.qin + sys) : 0.0);
        const long long col = (Model::N_FORC > 0 && a.n_forc > 0) ? (a.col ? (long long)a.col[sys] : sys) : 0;
        if (a.status[sys] == kStiff && (!(h > 0.0) || h < a.prm.initialStep)) h = a.prm.initialStep;
        int status = kStiffPaused;
        long long budget = a.max_attempts > 0 ? a.max_attempts : 0x7fffffffffffffffLL;
        bool unused = false;
        __syncwarp();  // every lane has read the link's state before lane 0 may write it back

        while (true) {  // every condition below is computed from replicated values: the warp stays together
            if (!(t < a.tf)) { status = kDoneStiff; break; }
            if (!run_to_end && next_q >= a.q_hi) break;
            if (budget-- <= 0 || !(h > 1e-13 * fmax(1.0, fabs(t)))) { status = kStalled; break; }
            if (t + h > a.tf) h = a.tf - t;

            double F[2] = {0.0, 0.0};
            if (Model::N_FORC > 0)
                for (int j = 0; j < Model::N_FORC && j < a.n_forc; ++j) {
                    double lo, hi;
                    const long long idx = forcing_index(t, a.forc_dt_min[j], a.forc_nT[j], lo, hi);
                    long long r = idx - a.forc_i0[j];
                    r = r < 0 ? 0 : (r >= a.forc_nres[j] ? a.forc_nres[j] - 1 : r);
                    F[j] = (double)__ldg(a.forc[j] + r * a.forc_ncols + col);
                }

            double f0[N], J[N][N];
            Model::template rhs<double, false>(y, F, L, f0, unused);
            {   // Jacobian column `lane` by lanes 0..N-1, all at once; then replicated
                const double eps = 1.4901161193847656e-08;
                double yp[N], f1[N], colv[N];
                const double yj = pick<N>(y, lane < N ? lane : 0);
                const double d = eps * fmax(1.0, fabs(yj));
                const double ypj = yj + d;
                const double dj = ypj - yj;
#pragma unroll
                for (int i = 0; i < N; ++i) yp[i] = (lane == i) ? ypj : y[i];
                Model::template rhs<double, false>(yp, F, L, f1, unused);
#pragma unroll
                for (int i = 0; i < N; ++i) colv[i] = (f1[i] - f0[i]) / dj;
#pragma unroll
                for (int i = 0; i < N; ++i)
#pragma unroll
                    for (int j = 0; j < N; ++j) J[i][j] = shfl_d(colv[i], j);
            }
            double Ji[N];  // row ri of the Jacobian
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double v = J[0][j];
#pragma unroll
                for (int i = 1; i < N; ++i)
                    if (ri == i) v = J[i][j];
                Ji[j] = v;
            }
            double Mrow[N3], Mrrow[N];
            int piv[N3], pivr[N];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double As = rs == 0 ? A[0][r] : (rs == 1 ? A[1][r] : A[2][r]);
#pragma unroll
                for (int j = 0; j < N; ++j) Mrow[r * N + j] = ((rs == r && ri == j) ? 1.0 : 0.0) - h * As * Ji[j];
            }
#pragma unroll
            for (int j = 0; j < N; ++j) Mrrow[j] = ((lane == j) ? MU_REAL / h : 0.0) - Ji[j];  // lanes < N: ri == lane
            bool ok = warp_lu_factor<N3>(Mrow, piv, lane) && warp_lu_factor<N>(Mrrow, pivr, lane);

            double K[3][N];
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int i = 0; i < N; ++i) K[s][i] = f0[i];
            bool converged = false;
            double prev = -1.0;
            const int ls = lane < 3 ? lane : 0;  // the stage this lane evaluates
            const double a0 = ls == 0 ? A[0][0] : (ls == 1 ? A[1][0] : A[2][0]);
            const double a1 = ls == 0 ? A[0][1] : (ls == 1 ? A[1][1] : A[2][1]);
            const double a2 = ls == 0 ? A[0][2] : (ls == 1 ? A[1][2] : A[2][2]);
            for (int it = 0; ok && it < kNewtonMaxIter && !converged; ++it) {
                double Ys[N], fs[N], g[N];
#pragma unroll
                for (int i = 0; i < N; ++i) Ys[i] = y[i] + h * (a0 * K[0][i] + a1 * K[1][i] + a2 * K[2][i]);
                Model::template rhs<double, false>(Ys, F, L, fs, unused);
#pragma unroll
                for (int i = 0; i < N; ++i) g[i] = fs[i] - (ls == 0 ? K[0][i] : (ls == 1 ? K[1][i] : K[2][i]));
                double b = 0.0;  // element (rs, ri) of the residual: component ri of the lane that evaluated stage rs
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double v = shfl_d(g[i], rs);
                    if (ri == i) b = v;
                }
                double G[N3];
                warp_lu_solve<N3>(Mrow, piv, b, G, lane);
                double norm = 0.0;
#pragma unroll
                for (int s = 0; s < 3; ++s)
#pragma unroll
                    for (int i = 0; i < N; ++i) {
                        K[s][i] += G[s * N + i];
                        const double v = fabs(h * G[s * N + i]) / (atol + rtol * fabs(y[i]));
                        if (v > norm || !(v == v)) norm = v;
                    }
                if (!(norm == norm) || (prev >= 0.0 && norm > 2.0 * prev && norm > 1.0)) break;
                converged = norm < 0.03;
                prev = norm;
            }
            if (!converged) {
                h *= 0.5;
                ++n_rej;
                continue;
            }
            double Z[3][N], y_new[N];
#pragma unroll
            for (int s = 0; s < 3; ++s)
#pragma unroll
                for (int i = 0; i < N; ++i) Z[s][i] = h * (A[s][0] * K[0][i] + A[s][1] * K[1][i] + A[s][2] * K[2][i]);
#pragma unroll
            for (int i = 0; i < N; ++i) y_new[i] = y[i] + Z[2][i];

            double ze[N], e[N];
#pragma unroll
            for (int i = 0; i < N; ++i) {
                ze[i] = (E[0] * Z[0][i] + E[1] * Z[1][i] + E[2] * Z[2][i]) / h;
                e[i] = f0[i] + ze[i];
            }
            {
                double x[N];
                warp_lu_solve<N>(Mrrow, pivr, pick<N>(e, lane < N ? lane : 0), x, lane);
#pragma unroll
                for (int i = 0; i < N; ++i) e[i] = x[i];
            }
            auto err_norm = [&](const double (&ev)[N]) {
                double m = 0.0;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double v = fabs(ev[i] / (atol + rtol * fmax(fabs(y[i]), fabs(y_new[i]))));
                    if (v > m || !(v == v)) m = v;
                }
                return m;
            };
            double err = err_norm(e);
            if (err > 1.0) {
                double yp[N], f1[N];
#pragma unroll
                for (int i = 0; i < N; ++i) yp[i] = y[i] + e[i];
                Model::template rhs<double, false>(yp, F, L, f1, unused);
#pragma unroll
                for (int i = 0; i < N; ++i) e[i] = f1[i] + ze[i];
                double x[N];
                warp_lu_solve<N>(Mrrow, pivr, pick<N>(e, lane < N ? lane : 0), x, lane);
#pragma unroll
                for (int i = 0; i < N; ++i) e[i] = x[i];
                err = err_norm(e);
            }
            double fac = a.prm.safety * sqrt(sqrt(1.0 / (err + 1e-16)));
            if (!(err == err)) { err = 2.0; fac = a.prm.minScale; }

            if (err <= 1.0) {
                const double t1 = t + h;
                bool overshoot = false;
                while (next_q < a.nq) {
                    const double tq = a.tq[next_q];
                    if (!(tq <= t1)) break;
                    if (next_q >= a.q_hi) { overshoot = true; break; }
                    if (tq > t && a.dense != nullptr && lane == 0) {
                        const double th = (tq - t) / h;
                        long long out = dense_base(a, sys, next_q);
                        for (int i = 0; i < N; ++i) {
                            if (!((a.dense_mask >> i) & 1u)) continue;
                            const double d1 = Z[0][i] / C1, d2 = (Z[1][i] - Z[0][i]) / (C2 - C1), d3 = (Z[2][i] - Z[1][i]) / (1.0 - C2);
                            const double dd1 = (d2 - d1) / C2, dd2 = (d3 - d2) / (1.0 - C1);
                            const double ddd = dd2 - dd1;
                            dense_put(a, out++, y[i] + th * (d1 + (th - C1) * (dd1 + (th - C2) * ddd)));
                        }
                    }
                    ++next_q;
                }
                if (overshoot) break;
#pragma unroll
                for (int i = 0; i < N; ++i) y[i] = y_new[i];
                t = t1;
                ++n_imp;
                h *= fmin(a.prm.maxScale, fmax(a.prm.minScale, fac));
            } else {
                ++n_rej;
                h *= fmin(a.prm.maxScale, fmax(a.prm.minScale, fmin(1.0, fac)));
            }
        }
        __syncwarp();
        if (lane == 0) {
            for (int i = 0; i < N; ++i) a.y[(long long)i * a.ld + sys] = y[i];
            a.t[sys] = t;
            a.h[sys] = h;
            a.next_q[sys] = next_q;
            a.status[sys] = status;
            a.n_reject[sys] = n_rej;
            ra.n_radau[sys] = n_imp;
            if constexpr (Model::HAS_INFLOW) {
                if (status != kStiffPaused) route_publish(a, sys, y[0]);
            }
        }
        __syncwarp();
    }
}

}  // namespace hlm
