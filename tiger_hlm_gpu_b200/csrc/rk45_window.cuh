// rk45_window.cuh — the hot path: batched adaptive Dormand–Prince 5(4) over one output window.
//
// Replaces the reference kernel rk45_then_radau_multi<Model> (solver/rk45_kernel.cu:17-176) and the
// device functions it inlines: rk45_step (solver/rk45_step_dense.cuh:33-145), rk45_dense
// (solver/rk45_step_dense.cuh:171-244) and norm_inf_diff (solver/event_detector.cuh:46-53).
//
// Same per-link algorithm, same FP operations in the same order (see fp_exact.cuh), different
// machine mapping:
//   * one lane per link, one warp per tile of 32 consecutive links, persistent warps that pull
//     tiles from an atomic counter — no __syncthreads anywhere, grid = SMs x resident CTAs;
//   * state, per-link parameters and counters are structure-of-arrays: every load/store of a tile
//     is one coalesced 256-byte line per column;
//   * all seven stage slopes, the state and the link parameters stay in registers for the whole
//     window (the reference keeps k[7][5] and its forcing slice on a 352-byte local stack and
//     re-reads 15 AoS doubles per rhs call);
//   * forcings stay on their GRID ([time][cell], shared by ~500 links per cell) and are re-read
//     only when a lane's time leaves the validity interval of its current sample; the exact
//     IEEE index computation of the reference runs only then;
//   * k0 is re-used whenever that is bit-safe: after a rejected or slope-jump attempt (t, y, F
//     unchanged) and, FSAL, after an accepted step whose stage-6 state equals y_next bit for bit
//     with the forcing sample unchanged;
//   * the run is cut into query windows so the dense output of a window fits in HBM; a lane that
//     has emitted every query of the window pauses with its complete loop state saved, which is
//     invisible to the step sequence (no clipping of h at window edges);
//   * dense records are written as contiguous 40-byte rows of the final [link][query][state]
//     layout the reference's host reorder (solver/rk45_api.hpp:255-267) produces, so there is no
//     reorder pass on either side.
#pragma once
#include "fp_exact.cuh"
#include "models.cuh"

#ifndef HLM_BLOCKS_PER_SM
#define HLM_BLOCKS_PER_SM 3
#endif
namespace hlm {

// Dormand–Prince tableau, same expressions as solver/rk45_step_dense.cuh:54-83 so the constants
// round identically.  Zero entries are kept: the reference multiplies through them and with a
// non-finite slope that matters (0*inf = NaN).  The tables live in __constant__ memory: with every
// loop unrolled the indices are literals and DMUL/DFMA read them as c[bank][offset] operands, with
// no UMOV pair per use (ncu: 105 UMOVs per attempt when they were immediates).
namespace dp {
#define HLM_DP_A                                                                                 \
    {{0, 0, 0, 0, 0, 0},                                                                         \
     {1.0 / 5.0, 0, 0, 0, 0, 0},                                                                 \
     {3.0 / 40.0, 9.0 / 40.0, 0, 0, 0, 0},                                                       \
     {44.0 / 45.0, -56.0 / 15.0, 32.0 / 9.0, 0, 0, 0},                                           \
     {19372.0 / 6561.0, -25360.0 / 2187.0, 64448.0 / 6561.0, -212.0 / 729.0, 0, 0},              \
     {9017.0 / 3168.0, -355.0 / 33.0, 46732.0 / 5247.0, 49.0 / 176.0, -5103.0 / 18656.0, 0},     \
     {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0}}
#define HLM_DP_B {35.0 / 384.0, 0.0, 500.0 / 1113.0, 125.0 / 192.0, -2187.0 / 6784.0, 11.0 / 84.0, 0.0}
#define HLM_DP_BALT \
    {5179.0 / 57600.0, 0.0, 7571.0 / 16695.0, 393.0 / 640.0, -92097.0 / 339200.0, 187.0 / 2100.0, 1.0 / 40.0}
// solver/rk45_step_dense.cuh:193-219
#define HLM_DP_P                                                                                                        \
    {{1.0, -8048581381.0 / 2820520608.0, 8663915743.0 / 2820520608.0, -12715105075.0 / 11282082432.0},                   \
     {0.0, 0.0, 0.0, 0.0},                                                                                              \
     {0.0, 131558114200.0 / 32700410799.0, -68118460800.0 / 10900136933.0, 87487479700.0 / 32700410799.0},              \
     {0.0, -1754552775.0 / 470086768.0, 14199869525.0 / 1410260304.0, -10690763975.0 / 1880347072.0},                   \
     {0.0, 127303824393.0 / 49829197408.0, -318862633887.0 / 49829197408.0, 701980252875.0 / 199316789632.0},           \
     {0.0, -282668133.0 / 205662961.0, 2019193451.0 / 616988883.0, -1453857185.0 / 822651844.0},                        \
     {0.0, 40617522.0 / 29380423.0, -110615467.0 / 29380423.0, 69997945.0 / 29380423.0}}
constexpr double hA[7][6] = HLM_DP_A;
constexpr double hB[7] = HLM_DP_B;
constexpr double hBALT[7] = HLM_DP_BALT;
constexpr double hP[7][4] = HLM_DP_P;
struct Tab64 {
    double A[7][6];
    double E[7];  // b - b_alt, the compile-time difference the reference's constant folding produces
    double B6;
    double P[7][4];
};
struct Tab32 {
    float A[7][6];
    float E[7];
    float B6;
    float P[7][4];
};
#define HLM_E(s) (hB[s] - hBALT[s])
#define HLM_TAB_INIT(T)                                                                                             \
    {                                                                                                               \
        {{(T)hA[0][0], (T)hA[0][1], (T)hA[0][2], (T)hA[0][3], (T)hA[0][4], (T)hA[0][5]},                            \
         {(T)hA[1][0], (T)hA[1][1], (T)hA[1][2], (T)hA[1][3], (T)hA[1][4], (T)hA[1][5]},                            \
         {(T)hA[2][0], (T)hA[2][1], (T)hA[2][2], (T)hA[2][3], (T)hA[2][4], (T)hA[2][5]},                            \
         {(T)hA[3][0], (T)hA[3][1], (T)hA[3][2], (T)hA[3][3], (T)hA[3][4], (T)hA[3][5]},                            \
         {(T)hA[4][0], (T)hA[4][1], (T)hA[4][2], (T)hA[4][3], (T)hA[4][4], (T)hA[4][5]},                            \
         {(T)hA[5][0], (T)hA[5][1], (T)hA[5][2], (T)hA[5][3], (T)hA[5][4], (T)hA[5][5]},                            \
         {(T)hA[6][0], (T)hA[6][1], (T)hA[6][2], (T)hA[6][3], (T)hA[6][4], (T)hA[6][5]}},                           \
            {(T)HLM_E(0), (T)HLM_E(1), (T)HLM_E(2), (T)HLM_E(3), (T)HLM_E(4), (T)HLM_E(5), (T)HLM_E(6)}, (T)hB[6], \
        {                                                                                                           \
            {(T)hP[0][0], (T)hP[0][1], (T)hP[0][2], (T)hP[0][3]}, {(T)hP[1][0], (T)hP[1][1], (T)hP[1][2], (T)hP[1][3]}, \
                {(T)hP[2][0], (T)hP[2][1], (T)hP[2][2], (T)hP[2][3]},                                               \
                {(T)hP[3][0], (T)hP[3][1], (T)hP[3][2], (T)hP[3][3]},                                               \
                {(T)hP[4][0], (T)hP[4][1], (T)hP[4][2], (T)hP[4][3]},                                               \
                {(T)hP[5][0], (T)hP[5][1], (T)hP[5][2], (T)hP[5][3]},                                               \
                {(T)hP[6][0], (T)hP[6][1], (T)hP[6][2], (T)hP[6][3]}                                                \
        }                                                                                                           \
    }
__constant__ Tab64 c64 = HLM_TAB_INIT(double);
__constant__ Tab32 c32 = HLM_TAB_INIT(float);
template <typename T> struct tab;
template <> struct tab<double> {
    static __device__ __forceinline__ const Tab64& get() { return c64; }
};
template <> struct tab<float> {
    static __device__ __forceinline__ const Tab32& get() { return c32; }
};
}  // namespace dp

// solver/event_detector.cuh:11,15
constexpr double kSlopeJumpThresh = 100.0;
constexpr double kMinStepFraction = 1e-6;

// kDoneStiff: flagged stiff by the RK45 path, then carried to tf by the implicit fallback (radau_fallback.cuh);
// kStiffPaused: in the fallback's hands, waiting for the next output window
enum LinkStatus : int { kActive = 0, kDone = 1, kStiff = 2, kStalled = 3, kDoneStiff = 4, kStiffPaused = 5 };

constexpr int kMaxForcings = 16;  // I_O/forcing_data.h:5

struct SolverParams {  // Model::Parameters, models/model_204.hpp:22-30
    double initialStep, rtol, atol, safety, minScale, maxScale;
};

// Everything one window launch needs.  Passed by value (fits the 4 KB parameter space).
struct WindowArgs {
    // resident per-link state, SoA, leading dimension ld (ns rounded up to 32)
    double* y;             // [N_EQ][ld]
    double* t;             // [ld]
    double* h;             // [ld]
    int* next_q;           // [ld]
    int* reject_run;       // [ld]   consecutive rejections (rk45_kernel.cu:130,155)
    int* status;           // [ld]   LinkStatus
    unsigned int* n_accept;  // [ld]
    unsigned int* n_reject;  // [ld]
    unsigned int* n_jump;    // [ld]
    const double* sp;      // [N_SP][ld]  model-prepared parameter columns
    const int* col;        // [ld] forcing column (grid cell) per link, or nullptr = link index
    const float* forc[2];  // forcing j: samples [forc_i0, forc_i0 + forc_nres) of the record, [nres][ncols]
    long long forc_nT[2];    // length of the whole record (the reference's c_forc_nT: the index clamps to it)
    long long forc_i0[2];    // first resident sample
    long long forc_nres[2];  // resident samples
    double forc_dt_min[2];  // c_forc_dt[j] * 60.0, rk45_kernel.cu:90
    long long forc_ncols;
    int n_forc;            // forcings present (0..2 used by the models here)
    const double* tq;      // [nq] ascending
    int nq;
    int q_lo, q_hi;        // this window's dense buffer covers queries [q_lo, q_hi)
    double* dense;         // [ns][q_hi - q_lo][N_EQ] or nullptr
    double t0, tf;
    SolverParams prm;
    long long ns, ld;
    // links of this launch: tiles [tile_lo, tile_lo + n_tiles) of 32 links; dense records of link sys go to
    // row sys - dense_sys0 of the buffer (a launch over a chunk of links fills a buffer of its own, which
    // then leaves in one contiguous copy)
    long long tile_lo, n_tiles, dense_sys0;
    long long max_attempts;  // per link per launch; <=0 = unbounded like the reference
    unsigned int* tile_counter;
    // routed models (Model::HAS_INFLOW): discharge entering each link from upstream, constant over the
    // interval (nullptr = unrouted, 0); and, for links another rank needs, where the epilogue puts the
    // link's discharge once the interval is integrated (send_slot[sys] < 0: not a boundary link)
    const double* qin;     // [ld]
    const int* send_slot;  // [ld] or nullptr
    double* send_buf;
};

template <typename T, int N>
struct StepOut {
    T err;
};

// One DOPRI5 attempt from (y, k0): fills k[1..6], y_next and the FSAL flag, returns err.
// solver/rk45_step_dense.cuh:94-142.  All loops are compile-time unrolled; k stays in registers.
template <class Model, typename T, bool kFast>
__device__ __forceinline__ T dopri_attempt(const T (&y)[Model::N_EQ], T (&k)[7][Model::N_EQ], T h,
                                           const T* F, const typename Model::template Link<T>& L, T rtol, T atol,
                                           T (&y_next)[Model::N_EQ], bool& fsal, bool& bad) {
    using f = fp<T>;
    constexpr int N = Model::N_EQ;
    const auto& TB = dp::tab<T>::get();
#pragma unroll
    for (int s = 1; s < 7; ++s) {
        T ha[6];
#pragma unroll
        for (int j = 0; j < s; ++j) ha[j] = f::mul(h, TB.A[s][j]);
        T yt[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            T acc = y[i];
#pragma unroll
            for (int j = 0; j < s; ++j) acc = f::fma(ha[j], k[j][i], acc);
            yt[i] = acc;
        }
        Model::template rhs<T, kFast>(yt, F, L, k[s], bad);
        if (s == 6) {
            // y_out = y + sum_{s<7} (h*b[s])*k[s].  a[6][j] == b[j] bit for bit for j < 6, so the first
            // six terms ARE the stage-6 state yt; only the last (zero-weight) term remains.  FSAL: k6 =
            // rhs(yt, F) is the next k0 iff y_next == yt bit for bit (that term changed nothing).
            const T hb6 = f::mul(h, TB.B6);
            fsal = true;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                y_next[i] = f::fma(hb6, k[6][i], yt[i]);
                fsal = fsal && f::same_bits(y_next[i], yt[i]);
            }
        }
    }
    T he[7];
#pragma unroll
    for (int s = 0; s < 7; ++s) he[s] = f::mul(h, TB.E[s]);
    T max_ratio = (T)0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        T e = (T)0;
#pragma unroll
        for (int s = 0; s < 7; ++s) e = f::fma(he[s], k[s][i], e);
        const T ymax = f::max_a(f::abs(y[i]), f::abs(y_next[i]));  // y NaN => y_next NaN: same result as fmax
        const T tol = f::fma(rtol, ymax, atol);
        const T ratio = f::abs(f::template div_err<kFast>(e, tol, bad));
        if (ratio > max_ratio) max_ratio = ratio;  // NaN-ignoring form of the reference
    }
    return max_ratio;
}

// Exact forcing sample index of the reference (rk45_kernel.cu:90-98) plus a conservative interval
// [lo, hi) of times for which that index is certain to be the same: 1e-9 relative inside the
// sample's edges, far beyond the rounding of t/dt_min.  Outside it the caller recomputes exactly.
__device__ __forceinline__ long long forcing_index(double t, double dt_min, long long nT, double& lo, double& hi) {
    const double r = __ddiv_rn(t, dt_min);
    long long idx = (r < 0.0) ? 0 : ((r >= (double)nT) ? nT - 1 : (long long)r);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    lo = (idx <= 0) ? -inf : (double)idx * dt_min * (1.0 + 1e-9);
    hi = (idx >= nT - 1) ? inf : (double)(idx + 1) * dt_min * (1.0 - 1e-9);
    if (!(dt_min > 0.0)) { lo = inf; hi = -inf; }  // degenerate spacing: never cache
    return idx;
}

template <class Model, typename T>
__global__ void __launch_bounds__(128, HLM_BLOCKS_PER_SM) rk45_window_kernel(const WindowArgs a) {
    using f = fp<T>;
    constexpr int N = Model::N_EQ;
    const int lane = threadIdx.x & 31;
    const long long n_tiles = a.n_tiles;
    const bool run_to_end = (a.q_hi >= a.nq);
    const int qw = a.q_hi - a.q_lo;
    const T rtol = (T)a.prm.rtol, atol = (T)a.prm.atol;
    const T safety = (T)a.prm.safety, minScale = (T)a.prm.minScale, maxScale = (T)a.prm.maxScale;
    const T h_floor = f::mul((T)a.prm.initialStep, (T)kMinStepFraction);     // rk45_kernel.cu:134
    const T h_stiff = f::mul(f::sub((T)a.tf, (T)a.t0), (T)kMinStepFraction);  // rk45_kernel.cu:160
    const T tf = (T)a.tf;

    for (;;) {
        unsigned int tile = 0;
        if (lane == 0) tile = atomicAdd(a.tile_counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if ((long long)tile >= n_tiles) break;
        const long long sys = ((a.tile_lo + (long long)tile) << 5) + lane;
        if (sys >= a.ns) continue;
        int status = a.status[sys];
        if (status != kActive) continue;

        // ---- load lane state (coalesced columns) ----
        T y[N], k[7][N], y_next[N];
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = (T)a.y[(long long)i * a.ld + sys];
        T t = (T)a.t[sys], h = (T)a.h[sys];
        int next_q = a.next_q[sys];
        int reject_run = a.reject_run[sys];
        unsigned int n_acc = a.n_accept[sys], n_rej = a.n_reject[sys], n_jmp = a.n_jump[sys];
        typename Model::template Link<T> L;
        L.load(a.sp, a.ld, sys);
        if constexpr (Model::HAS_INFLOW) L.set_inflow(a.qin ? (T)__ldg(a.qin + sys) : (T)0);
        const bool fast_ok = Model::template fast_div_ok<T>(L);
        const long long col = (Model::N_FORC > 0 && a.n_forc > 0) ? (a.col ? (long long)a.col[sys] : sys) : 0;

        T F[2] = {(T)0, (T)0};
        double f_lo = fp<double>::inf(), f_hi = -fp<double>::inf();  // empty validity interval
        T tq_next = (next_q < a.nq) ? (T)__ldg(a.tq + next_q) : f::inf();
        bool k0_valid = false;
        int budget = (a.max_attempts > 0x7fffffffLL) ? 0x7fffffff : (int)a.max_attempts;

        for (;;) {
            if (!(t < tf)) { status = kDone; break; }
            if (!run_to_end && next_q >= a.q_hi) break;  // window complete for this link: pause
            if (a.max_attempts > 0 && budget-- <= 0) { status = kStalled; break; }
            if (f::add(t, h) > tf) h = f::sub(tf, t);  // rk45_kernel.cu:54

            // ---- forcing sample at the step-start time, held for all stages ----
            if (Model::N_FORC > 0 && a.n_forc > 0) {
                const double td = (double)t;
                if (!(td >= f_lo && td < f_hi)) {
                    f_lo = -fp<double>::inf();
                    f_hi = fp<double>::inf();
#pragma unroll
                    for (int j = 0; j < Model::N_FORC; ++j) {
                        if (j < a.n_forc) {
                            double lo, hi;
                            const long long idx = forcing_index(td, a.forc_dt_min[j], a.forc_nT[j], lo, hi);
                            // resident chunk of the record (the host checks that it covers the interval; the
                            // clamp only keeps a mis-driven run inside the buffer)
                            long long r = idx - a.forc_i0[j];
                            r = r < 0 ? 0 : (r >= a.forc_nres[j] ? a.forc_nres[j] - 1 : r);
                            const T v = (T)__ldg(a.forc[j] + r * a.forc_ncols + col);  // f32 -> f64 widening as model_204.hpp:82-83
                            if (!f::same_bits(v, F[j])) k0_valid = false;
                            F[j] = v;
                            f_lo = fmax(f_lo, lo);
                            f_hi = fmin(f_hi, hi);
                        }
                    }
                }
            }

            // Fast attempt: constant-divisor divisions without guards, `bad` collects any operand that
            // needs the real div.rn.f64; then (rarely) the attempt is redone with exact divisions.
            // fac0 = safety * pow(1/(err + 1e-16), 0.2): the controller's factor, needed on both the
            // accept and the reject branch (rk45_kernel.cu:151,156), so computed once here.
            bool bad = !fast_ok, fsal = false;
            T err, fac0;
            if (!bad) {
                if (!k0_valid) Model::template rhs<T, true>(y, F, L, k[0], bad);  // rk45_kernel.cu:114
                err = dopri_attempt<Model, T, true>(y, k, h, F, L, rtol, atol, y_next, fsal, bad);
                fac0 = f::mul(safety, f::template pow_pos<true>(f::template rcp_pos<true>(f::add(err, (T)1e-16), bad), (T)0.2, bad));
            }
            if (__builtin_expect(bad, 0)) {
                bool unused = false;
                Model::template rhs<T, false>(y, F, L, k[0], unused);
                err = dopri_attempt<Model, T, false>(y, k, h, F, L, rtol, atol, y_next, fsal, unused);
                fac0 = f::mul(safety, f::template pow_pos<false>(f::rcp(f::add(err, (T)1e-16)), (T)0.2, unused));
            }

            if (err <= (T)1) {
                reject_run = 0;
                // slope-jump detection, event_detector.cuh:46-53 + rk45_kernel.cu:132-136
                T jump = (T)0;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const T d = f::abs(f::sub(k[0][i], k[1][i]));
                    if (d > jump) jump = d;
                }
                if (jump > (T)kSlopeJumpThresh) {
                    h = f::max_a(h_floor, f::mul(h, (T)0.5));
                    ++n_jmp;
                    k0_valid = true;  // same t, y, F
                    continue;
                }
                const T t1 = f::add(t, h);
                // ---- dense output for queries in (t, t1], rk45_kernel.cu:139-148 ----
                bool overshoot = false;
                if (next_q < a.nq && tq_next <= t1) {
                    // Q depends on the step only: built once here, not per query as the reference does.
                    T Q[4][N];
#pragma unroll
                    for (int m = 0; m < 4; ++m)
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            T sum = (T)0;
#pragma unroll
                            for (int j = 0; j < 7; ++j) sum = f::fma(dp::tab<T>::get().P[j][m], k[j][i], sum);
                            Q[m][i] = sum;
                        }
                    do {
                        if (next_q >= a.q_hi) { overshoot = true; break; }
                        if (tq_next > t && a.dense != nullptr) {
                            const T th = f::div(f::sub(tq_next, t), h);
                            double* out = a.dense + ((sys - a.dense_sys0) * qw + (next_q - a.q_lo)) * N;
                            const T th2 = f::mul(th, th), th3 = f::mul(th2, th), th4 = f::mul(th3, th);
#pragma unroll
                            for (int i = 0; i < N; ++i) {
                                T poly = f::fma(Q[0][i], th, (T)0);
                                poly = f::fma(Q[1][i], th2, poly);
                                poly = f::fma(Q[2][i], th3, poly);
                                poly = f::fma(Q[3][i], th4, poly);
                                out[i] = (double)f::fma(h, poly, y[i]);
                            }
                        }
                        ++next_q;
                        tq_next = (next_q < a.nq) ? (T)__ldg(a.tq + next_q) : f::inf();
                    } while (next_q < a.nq && tq_next <= t1);
                }
                if (overshoot) break;  // step spans past this window's buffer: leave it uncommitted, redo next window

                // FSAL (see dopri_attempt); the forcing sample must also be unchanged (checked next iteration)
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    y[i] = y_next[i];
                    k[0][i] = k[6][i];
                }
                k0_valid = fsal;
                t = t1;
                ++n_acc;
                h = f::mul(h, f::min_a(maxScale, f::max_a(minScale, fac0)));
            } else {
                ++reject_run;
                ++n_rej;
                T fac = f::min_a((T)1, fac0);
                fac = f::min_a(maxScale, f::max_a(minScale, fac));
                h = f::mul(h, fac);
                k0_valid = true;  // same t, y, F
                if (reject_run > 5 || h < h_stiff) { status = kStiff; break; }  // rk45_kernel.cu:160-162
            }
        }
        // A stiff bail-out at t >= tf cannot happen (the flag is only set with t < tf unchanged),
        // so kStiff here always means "flagged and unfinished" as rk45_kernel.cu:167-170.

        // ---- store lane state ----
#pragma unroll
        for (int i = 0; i < N; ++i) a.y[(long long)i * a.ld + sys] = (double)y[i];
        a.t[sys] = (double)t;
        a.h[sys] = (double)h;
        a.next_q[sys] = next_q;
        a.reject_run[sys] = reject_run;
        a.status[sys] = status;
        a.n_accept[sys] = n_acc;
        a.n_reject[sys] = n_rej;
        a.n_jump[sys] = n_jmp;
        // boundary exchange packed here instead of by a kernel of its own: the discharge another rank's
        // links need for the next interval goes straight into the send buffer
        if constexpr (Model::HAS_INFLOW) {
            if (a.send_slot != nullptr && status != kActive) {
                const int slot = __ldg(a.send_slot + sys);
                if (slot >= 0) a.send_buf[slot] = (double)y[0];
            }
        }
    }
}

}  // namespace hlm
