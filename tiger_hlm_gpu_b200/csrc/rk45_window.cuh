// rk45_window.cuh — the hot path: batched adaptive Dormand–Prince 5(4) over one output window.
//
// Replaces the reference kernel rk45_then_radau_multi<Model> (solver/rk45_kernel.cu:17-176) and the
// device functions it inlines: rk45_step (solver/rk45_step_dense.cuh:33-145), rk45_dense
// (solver/rk45_step_dense.cuh:171-244) and norm_inf_diff (solver/event_detector.cuh:46-53).
//
// Same per-link algorithm, same FP operations in the same order (see fp_exact.cuh), different
// machine mapping:
//   * one lane per link, one warp per tile of 32 links, persistent warps that pull tiles from an
//     atomic counter — no __syncthreads anywhere, grid = SMs x resident CTAs.  A tile is 32
//     consecutive links where neighbours step alike, and 32 links that took the same number of
//     attempts in the previous launch where they do not (sorted tiles: Model 200, routed runs);
//     lane refill (rk45_lanes_kernel) serves the launches that have no counts to sort by;
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
// FP32 instances of the tile kernel: 4 CTAs/SM (128 registers, 4 bytes of spills).  Measured at 1 M links, dry / wet:
// 3 CTAs 12.09 / 16.25 ms, 4 CTAs 11.33 / 14.77, 5 CTAs (96 registers, 264 B of spills) 11.34 / 15.60, 6 CTAs 11.98 / 16.03
// (profiles/r2t_ab_fp32_occupancy.log).  The FP64 instances stay at 3: their fourth CTA does not fit without spills.
#ifndef HLM_BLOCKS_PER_SM_F32
#define HLM_BLOCKS_PER_SM_F32 4
#endif
#ifndef HLM_BLOCKS_PER_SM_LANES
#define HLM_BLOCKS_PER_SM_LANES HLM_BLOCKS_PER_SM
#endif
#ifndef HLM_CTA_THREADS
#define HLM_CTA_THREADS 128  // the kernels only use the lane id: any multiple of 32 works
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
    void* dense;           // [links of the launch][q_hi - q_lo][dense_ncol] of double (or float), or nullptr
    unsigned int dense_mask;  // states a record carries, bit i = state i (hlm_set_output_states); dense_ncol = popcount
    int dense_ncol;
    int dense_f32;         // records stored as float (hlm_set_output_precision)
    double t0, tf;
    SolverParams prm;
    long long ns, ld;
    // links of this launch: tiles [tile_lo, tile_lo + n_tiles) of 32 links; dense records of link sys go to
    // row sys - dense_sys0 of the buffer (a launch over a chunk of links fills a buffer of its own, which
    // then leaves in one contiguous copy)
    long long tile_lo, n_tiles, dense_sys0;
    long long max_attempts;  // per link per launch; <=0 = unbounded like the reference
    int reject_limit;        // more consecutive rejections than this flag the link stiff (5: rk45_kernel.cu:160)
    // Lane-refill schedule, longest first: `order` (or nullptr = ascending) lists the launch's links by the attempts
    // they took in the previous launch, most first, so that the long links start early and the launch's tail is made
    // of short ones; `cost` (or nullptr) receives the attempts each link takes in this launch.
    const int* order;        // [links of the launch] link indices
    int* cost;               // [ld]
    // sorted tiles: the key a link leaves in `cost` carries, above its 6 bits of attempts, the number of its block of
    // 2^cost_block_shift consecutive links counted from the END (cost_blocks - 1 - block), so that one descending sort
    // lists the blocks in ascending order and, inside a block, the links by attempts; cost_blocks <= 1: attempts only
    int cost_block_shift, cost_blocks;
    unsigned int* tile_counter;
    // routed models (Model::HAS_INFLOW): discharge entering each link from upstream, constant over the
    // interval (nullptr = unrouted, 0); and, for links another rank needs, where the epilogue puts the
    // link's discharge once the interval is integrated (send_slot[sys] < 0: not a boundary link)
    const double* qin;     // [ld]
    const int* send_slot;  // [ld] or nullptr
    double* send_buf;
    // Forcing samples of the launch's start time, worked out once on the host (same arithmetic as forcing_index):
    // a link whose own time lies in [forc_pre_lo, forc_pre_hi) takes rows forc_pre_row[j] without recomputing
    // the index — in routed runs every link is (re)loaded every coupling interval, all at the same time.
    int forc_pre_ok;
    long long forc_pre_row[2];
    double forc_pre_lo, forc_pre_hi;
    // peer-memory exchange (one process per GPU on one node): instead of a send buffer a collective then moves,
    // the discharge is stored straight into every rank's halo vector through its IPC-mapped address, element
    // peer_off + slot (peer_off = parity * halo length + my rank * segment length); nullptr = not in use
    double* const* peer_halo;  // [peer_world] device pointers (own halo included)
    int peer_world;
    long long peer_off;
};

// Hand a boundary link's discharge to whoever needs it for the next interval (epilogue of the integration
// kernels and of the fallback): no pack kernel, and with peer memory no collective moving data either.
__device__ __forceinline__ void route_publish(const WindowArgs& a, long long sys, double q) {
    if (a.send_slot == nullptr) return;
    const int slot = __ldg(a.send_slot + sys);
    if (slot < 0) return;
    if (a.peer_halo != nullptr) {
        for (int r = 0; r < a.peer_world; ++r) a.peer_halo[r][a.peer_off + slot] = q;  // NVLink stores (own copy too)
    } else {
        a.send_buf[slot] = q;
    }
}

// Dense records.  A record holds the selected states of one (link, query) in ascending state order; the window
// kernels write EVERY slot of the launch's buffer exactly once — the interpolated states, or zeros for a query the
// link never reaches (tq <= t0; queries after a stiff bail-out, a stall or tf) — so no memset precedes
// a launch (it was 9.6 GB of extra HBM writes per step at 10 M links, serialised before the kernel).
__device__ __forceinline__ long long dense_base(const WindowArgs& a, long long sys, int q) {
    return ((sys - a.dense_sys0) * (long long)(a.q_hi - a.q_lo) + (q - a.q_lo)) * a.dense_ncol;
}
__device__ __forceinline__ void dense_put(const WindowArgs& a, long long idx, double v) {
    if (a.dense_f32) static_cast<float*>(a.dense)[idx] = (float)v;
    else static_cast<double*>(a.dense)[idx] = v;
}
// zero records for queries [q_from, q_to) of link sys (clipped to the window).  Out of line (it runs for the rare
// link that leaves queries unreached), and with every argument BY VALUE: a reference to the kernel's parameter
// struct would force the whole struct into local memory for the kernel's lifetime.
static __device__ __noinline__ void dense_zero_rows(void* dense, int f32, int ncol, long long first_record, int n_records) {
    for (long long i = first_record * ncol, end = (first_record + n_records) * ncol; i < end; ++i) {
        if (f32) static_cast<float*>(dense)[i] = 0.0f;
        else static_cast<double*>(dense)[i] = 0.0;
    }
}
__device__ __forceinline__ void dense_zero(const WindowArgs& a, long long sys, int q_from, int q_to) {
    if (a.dense == nullptr) return;
    if (q_from < a.q_lo) q_from = a.q_lo;
    if (q_to > a.q_hi) q_to = a.q_hi;
    if (q_to <= q_from) return;
    dense_zero_rows(a.dense, a.dense_f32, a.dense_ncol, (sys - a.dense_sys0) * (long long)(a.q_hi - a.q_lo) + (q_from - a.q_lo), q_to - q_from);
}

// The seven stage slopes of a link.  Default: all in registers.  HLM_K_SHARED (experiment): the slopes of stages 2..5
// in shared memory (one 8-byte column per thread and value, conflict-free), which takes 40 registers off a thread.
template <typename T, int N> struct KStore {
#ifdef HLM_K_SHARED
    T reg[3][N];  // stages 0, 1, 6
    T* sh;        // this thread's column of the CTA's array [4 * N][HLM_CTA_THREADS]
    static __device__ __forceinline__ constexpr int slot(int s) { return s == 0 ? 0 : (s == 1 ? 1 : 2); }
    __device__ __forceinline__ T get(int s, int i) const {
        return (s >= 2 && s <= 5) ? sh[((s - 2) * N + i) * HLM_CTA_THREADS] : reg[slot(s)][i];
    }
    __device__ __forceinline__ void set(int s, int i, T v) {
        if (s >= 2 && s <= 5) sh[((s - 2) * N + i) * HLM_CTA_THREADS] = v;
        else reg[slot(s)][i] = v;
    }
    __device__ __forceinline__ void bind(T* base) { sh = base; }
    __device__ __forceinline__ T* base() const { return sh; }
#else
    T reg[7][N];
    __device__ __forceinline__ T get(int s, int i) const { return reg[s][i]; }
    __device__ __forceinline__ void set(int s, int i, T v) { reg[s][i] = v; }
    __device__ __forceinline__ void bind(T*) {}
    __device__ __forceinline__ T* base() const { return nullptr; }
#endif
};
#ifdef HLM_K_SHARED
#define HLM_K_SHARED_DECL(T, N) __shared__ T hlm_ksh[4 * (N) * HLM_CTA_THREADS]
#define HLM_K_SHARED_BASE (hlm_ksh + threadIdx.x)
#else
#define HLM_K_SHARED_DECL(T, N)
#define HLM_K_SHARED_BASE nullptr
#endif

// One DOPRI5 attempt from (y, k0): fills k[1..6], y_next and the FSAL flag, returns err.
// solver/rk45_step_dense.cuh:94-142.  All loops are compile-time unrolled; k stays in registers.
template <class Model, typename T, bool kFast, typename G>
__device__ __forceinline__ T dopri_attempt(const T (&y)[Model::N_EQ], KStore<T, Model::N_EQ>& k, T h,
                                           const T* F, const typename Model::template Link<T>& L, T rtol, T atol,
                                           T (&y_next)[Model::N_EQ], bool& fsal, G& bad) {
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
            for (int j = 0; j < s; ++j) acc = f::fma(ha[j], k.get(j, i), acc);
            yt[i] = acc;
        }
        T ks[N];
        Model::template rhs<T, kFast>(yt, F, L, ks, bad);
#pragma unroll
        for (int i = 0; i < N; ++i) k.set(s, i, ks[i]);
        if (s == 6) {
            // y_out = y + sum_{s<7} (h*b[s])*k[s].  a[6][j] == b[j] bit for bit for j < 6, so the first
            // six terms ARE the stage-6 state yt; only the last (zero-weight) term remains.  FSAL: k6 =
            // rhs(yt, F) is the next k0 iff y_next == yt bit for bit (that term changed nothing).
            const T hb6 = f::mul(h, TB.B6);
            fsal = true;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                y_next[i] = f::fma(hb6, ks[i], yt[i]);
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
        for (int s = 0; s < 7; ++s) e = f::fma(he[s], k.get(s, i), e);
        const T ymax = f::max_a(f::abs(y[i]), f::abs(y_next[i]));  // y NaN => y_next NaN: same result as fmax
        const T tol = f::fma(rtol, ymax, atol);
        const T ratio = f::abs(f::template div_err<kFast>(e, tol, bad));
        // NaN-ignoring form of the reference: if (ratio > max_ratio) max_ratio = ratio, from 0
        if (i == 0) max_ratio = f::max0(ratio);
        else max_ratio = f::max_a(max_ratio, ratio);  // (ratio > max_ratio) ? ratio : max_ratio
    }
    return max_ratio;
}

// The same attempt for a model whose rhs has a side chain (Model::SPLIT_SURFACE): states nothing on the hillslope
// depends on — Model204's surface store; Model 200's surface store and, behind it, its channel.  The hillslope's slopes
// (Model::rhs_hill, branch-free) of stage s+1 do not wait for the side chain's slopes of stage s (Model::rhs_side: the
// pows), so the two are written side by side — hill(1); then side(s) next to hill(s+1) for s = 1..5; side(6) — in
// straight-line code the instruction scheduler interleaves: the 63-deep pow chain of one stage runs in the shadow of
// the next stage's sums.  Every value is computed by the operations of dopri_attempt on the same operands (the fma
// chains keep their order), so the result is the same bits.
// kWet = false: every stage's surface store is taken to be empty (no surface pow, no branch per stage); `not_dry`
// comes back true for a lane where some stage's store was not, and the caller redoes the attempt with kWet = true.
template <class Model, typename T, bool kWet, typename G>
__device__ __forceinline__ T dopri_attempt_split(const T (&y)[Model::N_EQ], KStore<T, Model::N_EQ>& k, T h, const T* F,
                                                 const typename Model::template Link<T>& L, T rtol, T atol,
                                                 T (&y_next)[Model::N_EQ], bool& fsal, G& bad, bool& not_dry) {
    using f = fp<T>;
    constexpr int N = Model::N_EQ;
    constexpr int NS = Model::N_SIDE, NH = Model::HILL_OUT;
    const auto& TB = dp::tab<T>::get();
    T ha[7][6];      // h * a[s][j]: stage s's coefficients serve hill(s) and, one step later, side(s)
    T ho[7][NH];     // what the hillslope of stage s hands to the side chain
    T yt6[N];        // the stage-6 state: y_next but for the zero-weight term
    auto side = [&](int s) {  // states of the side chain at stage s -> their slopes
        T ys[NS], ks[NS];
#pragma unroll
        for (int c = 0; c < NS; ++c) {
            const int i = Model::side_state(c);
            T acc = y[i];
#pragma unroll
            for (int j = 0; j < 6; ++j)
                if (j < s) acc = f::fma(ha[s][j], k.get(j, i), acc);
            ys[c] = acc;
            if (s == 6) yt6[i] = acc;
        }
        Model::template rhs_side<T, kWet>(ys, ho[s], L, ks, bad, not_dry);
#pragma unroll
        for (int c = 0; c < NS; ++c) k.set(s, Model::side_state(c), ks[c]);
    };
#pragma unroll
    for (int s = 1; s < 7; ++s) {
        // ---- hill(s): the states off the side chain ----
#pragma unroll
        for (int j = 0; j < s; ++j) ha[s][j] = f::mul(h, TB.A[s][j]);
        T yt[N], ks[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (Model::is_side(i)) {
                yt[i] = (T)0;  // not read by rhs_hill
                continue;
            }
            T acc = y[i];
#pragma unroll
            for (int j = 0; j < s; ++j) acc = f::fma(ha[s][j], k.get(j, i), acc);
            yt[i] = acc;
            if (s == 6) yt6[i] = acc;
        }
        Model::template rhs_hill<T>(yt, F, L, ks, ho[s], bad);
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (!Model::is_side(i)) k.set(s, i, ks[i]);
        // ---- side(s - 1) beside it (side(6) after the loop) ----
        if (s > 1) side(s - 1);
    }
    side(6);
    // y_out and FSAL as in dopri_attempt
    const T hb6 = f::mul(h, TB.B6);
    fsal = true;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        y_next[i] = f::fma(hb6, k.get(6, i), yt6[i]);
        fsal = fsal && f::same_bits(y_next[i], yt6[i]);
    }
    T he[7];
#pragma unroll
    for (int s = 0; s < 7; ++s) he[s] = f::mul(h, TB.E[s]);
    T max_ratio = (T)0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        T e = (T)0;
#pragma unroll
        for (int s = 0; s < 7; ++s) e = f::fma(he[s], k.get(s, i), e);
        const T ymax = f::max_a(f::abs(y[i]), f::abs(y_next[i]));
        const T tol = f::fma(rtol, ymax, atol);
        const T ratio = f::abs(f::template div_err<true>(e, tol, bad));
        if (i == 0) max_ratio = f::max0(ratio);
        else max_ratio = f::max_a(max_ratio, ratio);  // (ratio > max_ratio) ? ratio : max_ratio
    }
    return max_ratio;
}

// The exact attempt (real div.rn / rcp.rn / libdevice pow), for the rare attempt whose operands leave the domain
// of the fast forms.  Out of line on copies of the lane's state in local memory: the hot loop then holds one
// attempt, not two, which halves its instruction footprint (the L1.5 instruction cache is 32 KB, the two bodies
// together were 78 KB) and leaves the register allocation of the fast path without a second producer per value.
template <class Model, typename T> struct ExactIO {
    T y[Model::N_EQ], k[7][Model::N_EQ], y_next[Model::N_EQ], F[2];
    T h, err, fac0;
    typename Model::template Link<T> L;
    T* ksh;  // the lane's shared-memory column of stage slopes (HLM_K_SHARED), else unused
    bool fsal;
};
template <class Model, typename T>
__device__ __noinline__ void exact_attempt(ExactIO<Model, T>& io, T rtol, T atol, T safety) {
    using f = fp<T>;
    bool unused = false, fsal = false;
    constexpr int N = Model::N_EQ;
    Model::template rhs<T, false>(io.y, io.F, io.L, io.k[0], unused);  // rk45_kernel.cu:114
    KStore<T, N> k;
    k.bind(io.ksh);
#pragma unroll
    for (int i = 0; i < N; ++i) k.set(0, i, io.k[0][i]);
    io.err = dopri_attempt<Model, T, false>(io.y, k, io.h, io.F, io.L, rtol, atol, io.y_next, fsal, unused);
#pragma unroll
    for (int s = 1; s < 7; ++s)
#pragma unroll
        for (int i = 0; i < N; ++i) io.k[s][i] = k.get(s, i);
    io.fac0 = f::mul(safety, f::template pow_pos<false>(f::rcp(f::add(io.err, (T)1e-16)), (T)0.2, unused));
    io.fsal = fsal;
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

// A link's sort key for the next launch: the attempts it took in this one in 6 bits — exact up to 15, then in steps of
// 2 (to 47), 4 (to 111) and 16 (saturating at 368) — so that links of equal key take nearly equal numbers of attempts
// both where counts are small (a routed interval: 1 to 20) and where they are large (a day of Model 200: 10 to 300).
__device__ __forceinline__ int cost_key(unsigned int attempts) {
    const unsigned int a = attempts;
    const unsigned int k = a < 16u ? a : (a < 48u ? 16u + ((a - 16u) >> 1) : (a < 112u ? 32u + ((a - 48u) >> 2) : 48u + ((a - 112u) >> 4)));
    return (int)(k < 63u ? k : 63u);
}

// The word a link leaves in `cost`: bits 0-5 the SORT key — the larger of this launch's key and the previous launch's
// (a routed link alternates between, say, 2 and 3 attempts per interval as its steps fall across the interval's end;
// a tile lasts as long as its slowest lane, so what should be equal inside a tile is the upper envelope: counted offline on
// a network of 65 536 links, lanes busy 76 % sorted by the last count, 83 % by the larger of the last two) —, above it the block
// number from the end (cost_blocks > 1), and in bits 24-29 this launch's own key for the next launch to read.
__device__ __forceinline__ int cost_word(const WindowArgs& a, long long sys, unsigned int attempts, int previous_word) {
    const int key = cost_key(attempts);
    const int prev = (previous_word >> 24) & 63;
    const int sort_key = key > prev ? key : prev;
    const int block = a.cost_blocks > 1 ? ((a.cost_blocks - 1 - (int)(sys >> a.cost_block_shift)) << 6) : 0;
    return (key << 24) | block | sort_key;
}

// Tile schedule: a warp takes 32 consecutive links and stays with them until the slowest lane leaves.  Right
// when the lanes of a tile run in lockstep (links sorted by forcing cell with like parameters: 31.9 of 32
// threads active per instruction on the Model204 workload).  The other schedule is rk45_lanes_kernel below.
// kFollowOrder: the kernel can take its tiles from the launch's order (sorted tiles) and records the sort key.  Always
// for the models routed runs use; for the others a second instance, so that the plain one — the bench's hot kernel —
// keeps its instructions and registers exactly.
template <class Model, typename T, bool kFollowOrder = Model::HAS_INFLOW>
__global__ void __launch_bounds__(HLM_CTA_THREADS, sizeof(T) == 4 ? HLM_BLOCKS_PER_SM_F32 : HLM_BLOCKS_PER_SM) rk45_window_kernel(const WindowArgs a) {
    using f = fp<T>;
    constexpr int N = Model::N_EQ;
    HLM_K_SHARED_DECL(T, N);
    const int lane = threadIdx.x & 31;
    const long long n_tiles = a.n_tiles;
    const bool run_to_end = (a.q_hi >= a.nq);
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
        long long sys = ((a.tile_lo + (long long)tile) << 5) + lane;
        if (sys >= a.ns) continue;
        // Sorted tiles: the tile is 32 consecutive entries of
        // the launch's order — links that took the same number of attempts in the previous launch — instead of 32
        // consecutive links, so its lanes finish together although neighbouring links do not.
        if constexpr (kFollowOrder) {
            if (a.order != nullptr) sys = (long long)__ldg(a.order + (sys - (a.tile_lo << 5)));
        }
        int status = a.status[sys];
        if (status != kActive) {  // finished, abandoned or in the fallback's hands: no record from this path
            dense_zero(a, sys, a.q_lo, a.q_hi);
            continue;
        }

        // ---- load lane state (coalesced columns) ----
        T y[N], y_next[N];
        KStore<T, N> k;
        k.bind(HLM_K_SHARED_BASE);
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = (T)a.y[(long long)i * a.ld + sys];
        T t = (T)a.t[sys], h = (T)a.h[sys];
        int next_q = a.next_q[sys];
        int reject_run = a.reject_run[sys];
        unsigned int n_acc = a.n_accept[sys], n_rej = a.n_reject[sys], n_jmp = a.n_jump[sys];
        [[maybe_unused]] const unsigned int n_at_load = n_acc + n_rej + n_jmp;
        [[maybe_unused]] int cost_before = 0;
        if constexpr (kFollowOrder) {
            if (a.cost != nullptr) cost_before = a.cost[sys];
        }
        typename Model::template Link<T> L;
        L.load(a.sp, a.ld, sys);
        if constexpr (Model::HAS_INFLOW) L.set_inflow(a.qin ? (T)__ldg(a.qin + sys) : (T)0);
        const bool fast_ok = Model::template fast_div_ok<T>(L) && f::fast_params_ok(rtol, atol);
        const long long col = (Model::N_FORC > 0 && a.n_forc > 0) ? (a.col ? (long long)a.col[sys] : sys) : 0;

        T F[2] = {(T)0, (T)0};
        double f_lo = fp<double>::inf(), f_hi = -fp<double>::inf();  // empty validity interval
        T tq_next = (next_q < a.nq) ? (T)__ldg(a.tq + next_q) : f::inf();
        bool k0_valid = false;
        int budget = (a.max_attempts > 0x7fffffffLL) ? 0x7fffffff : (int)a.max_attempts;

        for (;;) {
#define HLM_LEAVE break
#define HLM_AGAIN continue
#include "rk45_attempt_body.inc"
#undef HLM_LEAVE
#undef HLM_AGAIN
        }
        // A stiff bail-out at t >= tf cannot happen (the flag is only set with t < tf unchanged),
        // so kStiff here always means "flagged and unfinished" as rk45_kernel.cu:167-170.

        if (status != kActive) dense_zero(a, sys, next_q, a.q_hi);  // queries this link will never reach
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
        if constexpr (kFollowOrder) {
            if (a.cost != nullptr) a.cost[sys] = cost_word(a, sys, n_acc + n_rej + n_jmp - n_at_load, cost_before);  // the next launch's sort key
        }
        if constexpr (Model::HAS_INFLOW) {
            if (status != kActive) route_publish(a, sys, (double)y[0]);
        }
    }
}


// Constants of one launch, shared by every link.
template <typename T> struct RunConsts {
    T rtol, atol, safety, minScale, maxScale, h_floor, h_stiff, tf;
    bool run_to_end;
    __device__ __forceinline__ explicit RunConsts(const WindowArgs& a) {
        using f = fp<T>;
        rtol = (T)a.prm.rtol;
        atol = (T)a.prm.atol;
        safety = (T)a.prm.safety;
        minScale = (T)a.prm.minScale;
        maxScale = (T)a.prm.maxScale;
        h_floor = f::mul((T)a.prm.initialStep, (T)kMinStepFraction);     // rk45_kernel.cu:134
        h_stiff = f::mul(f::sub((T)a.tf, (T)a.t0), (T)kMinStepFraction);  // rk45_kernel.cu:160
        tf = (T)a.tf;
        run_to_end = (a.q_hi >= a.nq);
    }
};

// One link's integration state, held in registers by its lane: load() / attempt() ... / store().
// attempt() is one pass of the reference's loop body (rk45_kernel.cu:53-164) and returns true when the lane
// leaves the link: integrated to tf, paused at the window's end, out of attempts, or flagged stiff.
template <class Model, typename T> struct LinkRun {
    using f = fp<T>;
    static constexpr int N = Model::N_EQ;
    long long sys, col;
    T y[N], y_next[N];
    KStore<T, N> k;
    T t, h, tq_next;
    int next_q, reject_run, status, budget;
    unsigned int n_acc, n_rej, n_jmp, n_at_load;
    int cost_before;
    typename Model::template Link<T> L;
    bool fast_ok, k0_valid;
    T F[2];
    double f_lo, f_hi;

    // coalesced columns when the lanes of a warp hold consecutive links
    __device__ __forceinline__ void load(const WindowArgs& a, long long sys_) {
        sys = sys_;
#pragma unroll
        for (int i = 0; i < N; ++i) y[i] = (T)a.y[(long long)i * a.ld + sys];
        t = (T)a.t[sys];
        h = (T)a.h[sys];
        next_q = a.next_q[sys];
        reject_run = a.reject_run[sys];
        n_acc = a.n_accept[sys];
        n_rej = a.n_reject[sys];
        n_jmp = a.n_jump[sys];
        n_at_load = n_acc + n_rej + n_jmp;
        cost_before = a.cost != nullptr ? a.cost[sys] : 0;
        L.load(a.sp, a.ld, sys);
        if constexpr (Model::HAS_INFLOW) L.set_inflow(a.qin ? (T)__ldg(a.qin + sys) : (T)0);
        col = (Model::N_FORC > 0 && a.n_forc > 0) ? (a.col ? (long long)a.col[sys] : sys) : 0;
        finish_load(a);
    }

    __device__ __forceinline__ void finish_load(const WindowArgs& a) {
        fast_ok = Model::template fast_div_ok<T>(L) && f::fast_params_ok((T)a.prm.rtol, (T)a.prm.atol);
        F[0] = F[1] = (T)0;
        f_lo = fp<double>::inf();
        f_hi = -fp<double>::inf();  // empty validity interval: the first attempt looks the samples up
        if (Model::N_FORC > 0 && a.n_forc > 0 && a.forc_pre_ok) {
            const double td = (double)t;
            if (td >= a.forc_pre_lo && td < a.forc_pre_hi) {  // the launch-level lookup holds for this link
#pragma unroll
                for (int j = 0; j < Model::N_FORC; ++j)
                    if (j < a.n_forc) F[j] = (T)__ldg(a.forc[j] + a.forc_pre_row[j] * a.forc_ncols + col);
                f_lo = a.forc_pre_lo;
                f_hi = a.forc_pre_hi;
            }
        }
        tq_next = (next_q < a.nq) ? (T)__ldg(a.tq + next_q) : f::inf();
        k0_valid = false;
        budget = (a.max_attempts > 0x7fffffffLL) ? 0x7fffffff : (int)a.max_attempts;
    }

    __device__ __forceinline__ bool attempt(const WindowArgs& a, const RunConsts<T>& c) {
        const T rtol = c.rtol, atol = c.atol, safety = c.safety, minScale = c.minScale, maxScale = c.maxScale;
        const T h_floor = c.h_floor, h_stiff = c.h_stiff, tf = c.tf;
        const bool run_to_end = c.run_to_end;
#define HLM_LEAVE return true
#define HLM_AGAIN return false
#include "rk45_attempt_body.inc"
#undef HLM_LEAVE
#undef HLM_AGAIN
        return false;
    }

    // The loop's own exit tests (top of the body) would only fire on the NEXT pass, during which the lane would
    // idle while the others of its warp integrate — with 3 attempts per link that is a quarter of its time.
    // Same tests, same outcome, one pass earlier.
    __device__ __forceinline__ bool finished(const WindowArgs& a, const RunConsts<T>& c) {
        if (!(t < c.tf)) {
            status = kDone;
            return true;
        }
        return !c.run_to_end && next_q >= a.q_hi;
    }

    // A stiff bail-out at t >= tf cannot happen (the flag is only set with t < tf unchanged),
    // so kStiff here always means "flagged and unfinished" as rk45_kernel.cu:167-170.
    __device__ __forceinline__ void store(const WindowArgs& a) const {
        if (status != kActive) dense_zero(a, sys, next_q, a.q_hi);  // queries this link will never reach
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
        if (a.cost != nullptr) a.cost[sys] = cost_word(a, sys, n_acc + n_rej + n_jmp - n_at_load, cost_before);  // the sort reads bits 0-5 (+ block)
        // boundary exchange packed here instead of by a kernel of its own: the discharge another rank's
        // links need for the next interval goes straight into the send buffer
        if constexpr (Model::HAS_INFLOW) {
            if (status != kActive) route_publish(a, sys, (double)y[0]);
        }
    }
};

// Lane-refill schedule: every lane owns one link at a time and, the moment it leaves it, takes the next
// unclaimed link of the launch (warp-aggregated claim: one atomicAdd per warp and iteration, ranks by ballot),
// so a warp never idles behind its slowest link.  For workloads whose links take unlike numbers of attempts
// per launch — routed runs: short coupling intervals, discharge growing downstream, 3 attempts on average
// and 20 at the tail — where the tile schedule leaves three quarters of the lanes idle.  Loads and stores of a
// lane are then its own (not coalesced): per link that is ~300 bytes against thousands of FP64 instructions.
// Per-link arithmetic is the same function, so results are bit-identical under either schedule.
// kEarlyLeave: test for "link finished" right after the attempt instead of at the top of the next pass (see
// LinkRun::finished).  Worth it when links take few attempts per launch (routed runs: +10 %); with tens of
// attempts per link the larger register footprint of the merged control flow costs more than the idle pass
// (unrouted Model 200: -10 %), so the launcher picks.
// (Experiments on this kernel that were measured and not kept — the next link's columns staged in shared memory by
// cp.async, positions dealt statically, claims and order lookups two passes ahead through a per-warp ring with L2
// prefetch — are in DESIGN.md section 5 with their logs under profiles/; commit 2f565ba holds the staged version.)
template <class Model, typename T, bool kEarlyLeave>
__global__ void __launch_bounds__(HLM_CTA_THREADS, HLM_BLOCKS_PER_SM_LANES) rk45_lanes_kernel(const WindowArgs a) {
    const unsigned int lane = threadIdx.x & 31;
    const long long first = a.tile_lo << 5;
    const long long last = ((a.tile_lo + a.n_tiles) << 5) < a.ns ? ((a.tile_lo + a.n_tiles) << 5) : a.ns;
    const long long n_links = last - first;
    const RunConsts<T> c(a);
    HLM_K_SHARED_DECL(T, Model::N_EQ);
    LinkRun<Model, T> r;
    r.k.bind(HLM_K_SHARED_BASE);
    bool have = false, exhausted = false;
    for (;;) {
        const unsigned int need = __ballot_sync(0xffffffffu, !have && !exhausted);
        if (need) {
            unsigned int base = 0;
            const int leader = __ffs(need) - 1;
            if ((int)lane == leader) base = atomicAdd(a.tile_counter, (unsigned int)__popc(need));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (!have && !exhausted) {
                const long long idx = (long long)base + __popc(need & ((1u << lane) - 1u));
                if (idx >= n_links) {
                    exhausted = true;
                } else {
                    const long long sys = a.order != nullptr ? (long long)__ldg(a.order + idx) : first + idx;
                    // every column of the link is requested before its status is looked at: the loads are the lane's own
                    // (not coalesced) and a status test in front of them would put one more DRAM latency in series
                    // (routed hour of 2.5 M links 5.91 -> 5.74 ms).  Claiming one link ahead and prefetching its columns
                    // into L2 was tried on top: 6.60 ms — the second link index and 25 prefetches per link cost more
                    // registers (628 B of spills) and issue slots than the shorter load latency returns.
                    r.status = a.status[sys];
                    r.load(a, sys);
                    if (r.status == kActive) {
                        have = true;
                    } else {
                        if (a.cost != nullptr) a.cost[sys] = cost_word(a, sys, 0u, 0);
                        dense_zero(a, sys, a.q_lo, a.q_hi);
                    }
                }
            }
        }
        if (__ballot_sync(0xffffffffu, have) == 0u) {
            if (__ballot_sync(0xffffffffu, !exhausted) == 0u) break;
            continue;
        }
        if (have && (r.attempt(a, c) || (kEarlyLeave && r.finished(a, c)))) {
            r.store(a);
            have = false;
        }
    }
}


}  // namespace hlm
