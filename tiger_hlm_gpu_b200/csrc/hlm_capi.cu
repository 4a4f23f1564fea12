// hlm_capi.cu — the C ABI of include/hlm_b200.h over the kernels in rk45_window.cuh.
//
// Host side of the operator: what the reference spreads over setup_gpu_buffers /
// launch_rk45_kernel / retrieve_and_free (solver/rk45_api.hpp:63-270), setModelParameters
// (model_registry.cpp:18-60) and the upload code in main.cpp:392-404,552-574.  Differences by
// design: buffers are owned by a context and re-used across calls (the reference allocates and
// frees five buffers per call), everything runs on one stream without device-wide synchronisation,
// results land directly in the caller's [link][query][state] layout, and long runs are cut into
// query windows whose D2H copy overlaps the next window's integration.
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/hlm_b200.h"
#include "rk45_window.cuh"

// rk45_instance.cu: the integration kernels, one translation unit per (model, number type)
namespace hlm {
#define HLM_DECLARE_LAUNCHER(name) cudaError_t name(int schedule, const WindowArgs& a, int sm_count, cudaStream_t stream)
HLM_DECLARE_LAUNCHER(launch_rk45_204_f64);
HLM_DECLARE_LAUNCHER(launch_rk45_204_f32);
HLM_DECLARE_LAUNCHER(launch_rk45_200_f64);
HLM_DECLARE_LAUNCHER(launch_rk45_200_f32);
HLM_DECLARE_LAUNCHER(launch_rk45_dummy_f64);
HLM_DECLARE_LAUNCHER(launch_rk45_dummy_f32);
#undef HLM_DECLARE_LAUNCHER
}  // namespace hlm

// radau_fallback.cu (its own translation unit: built without FMA contraction)
namespace hlm {
cudaError_t radau_launch(int uid, const WindowArgs& a, int* list, unsigned int* n_list, unsigned int* n_radau,
                         int sm_count, cudaStream_t stream);
}

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define HLM_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fail(HLM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));          \
    } while (0)

#define HLM_REQUIRE(cond, msg)                            \
    do {                                                  \
        if (!(cond)) return fail(HLM_ERR_INVALID, (msg)); \
    } while (0)

// device buffer that grows but never shrinks
template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;  // elements
    cudaError_t reserve(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

struct ModelInfo {
    int uid, n_eq, n_sp, n_forc;
};
const ModelInfo kModels[] = {
    {hlm::DummyModel::UID, hlm::DummyModel::N_EQ, hlm::DummyModel::N_SP, hlm::DummyModel::N_FORC},
    {hlm::Model204::UID, hlm::Model204::N_EQ, hlm::Model204::N_SP, hlm::Model204::N_FORC},
    {hlm::Model200::UID, hlm::Model200::N_EQ, hlm::Model200::N_SP, hlm::Model200::N_FORC},
};
const ModelInfo* find_model(int uid) {
    for (const auto& m : kModels)
        if (m.uid == uid) return &m;
    return nullptr;
}

}  // namespace

struct hlm_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr;   // compute
    cudaStream_t stream = nullptr;       // = own_stream or the caller's
    cudaStream_t copy_stream = nullptr;  // D2H of finished windows
    cudaStream_t h2d_stream = nullptr;   // y0 of later chunks while the first ones integrate (run_link_chunks)
    DevBuf<double> io_stage;             // y0 in, results out, in the caller's layouts
    std::vector<cudaEvent_t> chunk_ready;
    std::map<int, hlm::SolverParams> params;
    long long max_attempts = 0;
    int reject_limit = 5;  // solver/rk45_kernel.cu:160
    long long dense_window_bytes = 8LL << 30;
    int precision = 64;
    unsigned int out_mask = 0;  // states a dense record carries (0 = all), hlm_set_output_states
    int out_bits = 64;          // dense records as double or float, hlm_set_output_precision

    // per-link parameters (AoS copy kept on device so any model can be prepared lazily)
    DevBuf<unsigned char> sp_aos;
    long long sp_n = 0;
    DevBuf<double> sp_soa;
    int sp_soa_uid = -1;
    long long sp_soa_ld = 0;

    // forcings
    DevBuf<float> forc[2];
    long long forc_nT[2] = {0, 0};    // whole record
    long long forc_i0[2] = {0, 0};    // first resident sample
    long long forc_nres[2] = {0, 0};  // resident samples
    long long forc_ncols = 0;
    double forc_dt_h[2] = {0, 0};
    int n_forc = 0;
    DevBuf<int> forc_col;
    long long forc_col_n = 0;  // 0 = identity
    long long forc_col_min = 0, forc_col_max = -1;  // range of the map's entries (checked against forc_ncols)

    // session
    bool in_session = false;
    int uid = -1;
    int n_eq = 0;
    long long ns = 0, ld = 0, nq = 0;
    double t0 = 0, tf = 0;
    long long q_done = 0;  // queries [0, q_done) have been emitted
    DevBuf<double> y, t, h, tq;
    DevBuf<int> next_q, reject_run, status;
    DevBuf<unsigned int> n_acc, n_rej, n_jump, tile_counter;
    DevBuf<unsigned long long> totals;
    DevBuf<double> dense[2];
    int dense_cur = 0;
    long long win_q_lo = 0, win_q_hi = 0;
    bool win_has_dense = false;
    cudaEvent_t ev_kernel_done[2] = {nullptr, nullptr};
    cudaEvent_t ev_copy_done[2] = {nullptr, nullptr};
    bool copy_pending[2] = {false, false};

    // implicit fallback for links the RK45 path flags stiff (radau_fallback.cuh)
    bool stiff_fallback = false;
    DevBuf<int> radau_list;
    DevBuf<unsigned int> radau_count, n_radau;

    int schedule = HLM_SCHEDULE_AUTO;
    // lane-refill schedule, longest first: attempts per link of the last launch -> order of the next (dispatch_window)
    DevBuf<int> cost, cost_sorted, iota, order;
    DevBuf<unsigned char> sort_tmp;
    long long cost_ns = 0;  // links the costs were recorded for (0 = none yet)
    long long order_ns = 0;  // links `order` lists (0 = not sorted yet)
    int order_age = 0;       // launches dealt by the current `order`
    bool longest_first = true;

    // routed runs (models with upstream inflow): topology of the links this context owns
    bool routed = false;
    long long route_ns = 0, route_nnz = 0, route_n_send = 0, route_halo_need = 0;
    DevBuf<long long> up_ptr;
    DevBuf<int> up_idx, send_idx, send_slot;
    DevBuf<double> qin, own_send;
    double* send_buf = nullptr;  // own_send.p or the caller's (a buffer its collective reads)
    // peer-memory exchange: every rank's halo vector (two parities) mapped here through CUDA IPC
    int peer_world = 0, peer_rank = 0;
    long long peer_max_send = 0, route_epoch = 0;
    double* own_halo = nullptr;           // cudaMalloc'ed: [2][peer_world * peer_max_send]
    std::vector<void*> peer_opened;       // cudaIpcOpenMemHandle results (nullptr at peer_rank)
    DevBuf<double*> peer_ptrs;            // device copy of the world pointers

    long long launches = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing;  // per window kernel
    std::vector<cudaEvent_t> event_pool;
};

namespace {

int use_device(hlm_ctx* c) {
    HLM_CUDA(cudaSetDevice(c->device));
    return 0;
}

// dense records: selected states of the session's model, columns per record, bytes per record
unsigned int dense_mask(const hlm_ctx* c) {
    const unsigned int all = (1u << c->n_eq) - 1u;
    const unsigned int m = c->out_mask & all;
    return m ? m : all;
}
int dense_ncol(const hlm_ctx* c) { return __builtin_popcount(dense_mask(c)); }
size_t dense_elem(const hlm_ctx* c) { return c->out_bits == 32 ? sizeof(float) : sizeof(double); }
size_t dense_record_bytes(const hlm_ctx* c) { return (size_t)dense_ncol(c) * dense_elem(c); }
// DevBuf<double> elements that hold `bytes`
size_t doubles_for(size_t bytes) { return (bytes + sizeof(double) - 1) / sizeof(double); }

// ---- kernels that are not the hot path -----------------------------------------------------------

template <class Model>
__global__ void prepare_params_kernel(const unsigned char* __restrict__ aos, long long n, long long stride,
                                      double* __restrict__ soa, long long ld) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ld) return;
    double out[Model::N_SP > 0 ? Model::N_SP : 1];
    if (i < n) {
        hlm::SpatialParamsAoS rec;
        const double* src = reinterpret_cast<const double*>(aos + i * stride);  // 8-byte aligned records
        double* dst = reinterpret_cast<double*>(&rec);
#pragma unroll
        for (int w = 0; w < 17; ++w) dst[w] = src[w];
        Model::prepare(rec, out);
    } else {
        for (int c = 0; c < Model::N_SP; ++c) out[c] = 1.0;  // padding lanes are never integrated
    }
    for (int c = 0; c < Model::N_SP; ++c) soa[(long long)c * ld + i] = out[c];
    if (Model::kWetStride > 0)  // the surface branch's parameters as one record per link (models.cuh, wet_block)
        Model::prepare_wet(out, soa + (long long)Model::N_SP * ld + i * Model::kWetStride);
}

// links [lo, hi) of the padded range (hi may reach ld: padding lanes are parked as done)
__global__ void init_state_kernel(const double* __restrict__ y0_aos, int n_eq, long long ns, long long ld,
                                  double* __restrict__ y, double* __restrict__ t, double* __restrict__ h,
                                  int* next_q, int* reject_run, int* status, unsigned int* n_acc,
                                  unsigned int* n_rej, unsigned int* n_jump, double t0, double h0, long long lo,
                                  long long hi) {
    long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const bool live = i < ns;
    for (int c = 0; c < n_eq; ++c) y[(long long)c * ld + i] = live ? y0_aos[i * n_eq + c] : 0.0;
    t[i] = t0;
    h[i] = h0;
    next_q[i] = 0;
    reject_run[i] = 0;
    status[i] = live ? hlm::kActive : hlm::kDone;
    n_acc[i] = n_rej[i] = n_jump[i] = 0u;
}

// new interval from the resident final states: what a second run_rk45 call with y0 = previous
// final does, without the host round trip.  Links that did not finish stay flagged.
// keep_step: the interval continues the previous one (hlm_solve_advance) — time and step size stay.
__global__ void restart_state_kernel(long long ld, double* __restrict__ t, double* __restrict__ h, int* next_q,
                                     int* reject_run, int* status, double t0, double h0, bool keep_step) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ld) return;
    if (status[i] == hlm::kDone || status[i] == hlm::kActive || status[i] == hlm::kDoneStiff) {
        if (!keep_step) {
            t[i] = t0;
            h[i] = h0;
        }
        next_q[i] = 0;
        reject_run[i] = 0;
        status[i] = hlm::kActive;
    }
}

// session results in the caller's formats: final states back in [link][state] order (unfinished links give
// zeros), counters widened to 64 bits, status as the HLM_LINK_* codes of hlm_b200.h
__global__ void export_results_kernel(const double* __restrict__ y, const int* __restrict__ status,
                                      const unsigned int* __restrict__ n_acc, const unsigned int* __restrict__ n_rej,
                                      const unsigned int* __restrict__ n_jump, int n_eq, long long ns, long long ld,
                                      double* __restrict__ out_aos, long long* __restrict__ out_cnt,
                                      int* __restrict__ out_code, long long lo, long long hi) {
    long long i = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    const int s = status[i];
    const bool ok = s == hlm::kDone || s == hlm::kDoneStiff;
    if (out_aos)
        for (int c = 0; c < n_eq; ++c) out_aos[i * n_eq + c] = ok ? y[(long long)c * ld + i] : 0.0;
    out_cnt[i] = (long long)n_acc[i];
    out_cnt[ns + i] = (long long)n_rej[i];
    out_cnt[2 * ns + i] = (long long)n_jump[i];
    out_code[i] = (s == hlm::kStiff || s == hlm::kStiffPaused) ? HLM_LINK_STIFF
                  : (s == hlm::kDone ? HLM_LINK_OK : (s == hlm::kDoneStiff ? HLM_LINK_STIFF_SOLVED : HLM_LINK_STALLED));
}


__global__ void totals_kernel(const unsigned int* __restrict__ n_acc, const unsigned int* __restrict__ n_rej,
                              const unsigned int* __restrict__ n_jump, const int* __restrict__ status,
                              long long ns, unsigned long long* __restrict__ out) {
    unsigned long long v[7] = {0, 0, 0, 0, 0, 0, 0};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ns;
         i += (long long)gridDim.x * blockDim.x) {
        v[0] += n_acc[i];
        v[1] += n_rej[i];
        v[2] += n_jump[i];
        const int s = status[i];
        v[3 + (s == hlm::kDoneStiff ? (int)hlm::kDone : (s == hlm::kStiffPaused ? (int)hlm::kStiff : (s < 0 || s > 3 ? 0 : s)))] += 1;
    }
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        unsigned long long x = v[k];
        for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
        if ((threadIdx.x & 31) == 0 && x) atomicAdd(out + k, x);
    }
}

__global__ void iota_kernel(int* __restrict__ p, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = (int)i;
}

// Register-resident FMA throughput probe: 8 independent chains per thread.
template <typename T> __global__ void fma_peak_kernel(T* out, int iters, T seed) {
    T a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6,
      a7 = seed + 7;
    const T m = (T)0.999999, c = (T)1e-6;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = hlm::fp<T>::fma(a0, m, c); a1 = hlm::fp<T>::fma(a1, m, c);
            a2 = hlm::fp<T>::fma(a2, m, c); a3 = hlm::fp<T>::fma(a3, m, c);
            a4 = hlm::fp<T>::fma(a4, m, c); a5 = hlm::fp<T>::fma(a5, m, c);
            a6 = hlm::fp<T>::fma(a6, m, c); a7 = hlm::fp<T>::fma(a7, m, c);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// Element-wise probes of device arithmetic, for tests that compare it with the host's.
__global__ void probe_kernel(int op, const double* __restrict__ x, const double* __restrict__ y,
                             double* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double r;
    switch (op) {
    case 0: r = pow(x[i], y[i]); break;
    case 1: asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[i])); break;
    case 2: r = __ddiv_rn(x[i], y[i]); break;
    case 3: r = __dsqrt_rn(x[i]); break;
    case 4: {  // the kernels' fast pow with its fallback, as the solver uses it
        hlm::fp<double>::guard g;
        r = hlm::fp<double>::pow_pos<true>(x[i], y[i], g);
        if (g.failed()) r = hlm::fp<double>::pow_pos<false>(x[i], y[i], g);
        break;
    }
    case 5: r = hlm::fp<double>::root5(x[i]); break;  // Model 200's x^(1/5)
    case 6: r = hlm::fp<double>::cbrt2(x[i]); break;  // Model 200's x^(2/3)
    default: r = 0.0;
    }
    out[i] = r;
}

// ---- routing: boundary pack and upstream gather (HBM-bound plumbing around the window kernel) ----
// peer_halo != nullptr: store into every rank's halo vector (peer memory) instead of the send buffer
__global__ void route_pack_kernel(const double* __restrict__ q, const int* __restrict__ send_idx, long long n_send,
                                  double* __restrict__ send_buf, double* const* __restrict__ peer_halo, int peer_world,
                                  long long peer_off) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_send) return;
    const double v = q[send_idx[k]];
    if (peer_halo != nullptr)
        for (int r = 0; r < peer_world; ++r) peer_halo[r][peer_off + k] = v;
    else
        send_buf[k] = v;
}

// qin[i] = sum of the discharge of link i's upstream links, added in the order the topology lists them
// (ascending global link id), so the sum does not depend on how links are partitioned over ranks.
__global__ void route_gather_kernel(const double* __restrict__ q, const double* __restrict__ halo,
                                    const long long* __restrict__ up_ptr, const int* __restrict__ up_idx, long long ns,
                                    double* __restrict__ qin) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ns) return;
    double acc = 0.0;
    for (long long e = up_ptr[i]; e < up_ptr[i + 1]; ++e) {
        const int u = up_idx[e];
        acc = __dadd_rn(acc, u >= 0 ? q[u] : halo[-(u + 1)]);
    }
    qin[i] = acc;
}

template <class Model> int prepare_params(hlm_ctx* c) {
    const long long ld = c->ld;
    if (Model::N_SP == 0) return 0;
    if (c->sp_n != c->ns)
        return fail(HLM_ERR_STATE, "model needs per-link parameters: upload exactly ns SpatialParams records first");
    if (c->sp_soa_uid == Model::UID && c->sp_soa_ld == ld) return 0;
    HLM_CUDA(c->sp_soa.reserve((size_t)(Model::N_SP + Model::kWetStride) * ld));
    const int tpb = 256;
    prepare_params_kernel<Model><<<(unsigned)((ld + tpb - 1) / tpb), tpb, 0, c->stream>>>(
        c->sp_aos.p, c->sp_n, (long long)sizeof(hlm::SpatialParamsAoS), c->sp_soa.p, ld);
    HLM_CUDA(cudaGetLastError());
    ++c->launches;
    c->sp_soa_uid = Model::UID;
    c->sp_soa_ld = ld;
    return 0;
}

// SoA columns of the session's model, rebuilt when the records or the model changed since they were made
int prepare_model_params(hlm_ctx* c) {
    if (c->uid == hlm::Model204::UID) return prepare_params<hlm::Model204>(c);
    if (c->uid == hlm::Model200::UID) return prepare_params<hlm::Model200>(c);
    return 0;
}

cudaEvent_t get_event(hlm_ctx* c) {
    if (!c->event_pool.empty()) {
        cudaEvent_t e = c->event_pool.back();
        c->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// One window launch.  Schedule (rk45_window.cuh): tiles of 32 consecutive links; tiles of 32 links that took the same
// number of attempts in the previous launch (sorted tiles); or lane refill.  By default plain tiles, and where links
// take unlike numbers of attempts per launch — Model 200 (the channel's pace grows with its discharge: 24 attempts per
// day at the median, 150 at the 99th percentile) and every routed run — sorted tiles, with lane refill for the launches
// that have no counts to sort by.  The kernels live in rk45_instance.cu, one translation unit per (model, number type).
int dispatch_window(hlm_ctx* c, const hlm::WindowArgs& a_in) {
    using Launcher = cudaError_t (*)(int, const hlm::WindowArgs&, int, cudaStream_t);
    Launcher launch = nullptr;
    bool divergent_model = false;
    const bool f32 = c->precision == 32;
    if (c->uid == hlm::Model204::UID) launch = f32 ? hlm::launch_rk45_204_f32 : hlm::launch_rk45_204_f64;
    else if (c->uid == hlm::Model200::UID) { launch = f32 ? hlm::launch_rk45_200_f32 : hlm::launch_rk45_200_f64; divergent_model = true; }
    else if (c->uid == hlm::DummyModel::UID) launch = f32 ? hlm::launch_rk45_dummy_f32 : hlm::launch_rk45_dummy_f64;
    else return fail(HLM_ERR_INVALID, "unknown model uid");
    // Which schedule.  Where links take unlike numbers of attempts per launch (Model 200, every routed run) AUTO asks
    // for sorted tiles — tiles of 32 links that took the same number of attempts in the previous launch — when the model's
    // run can change its schedule freely (the models with an inflow term: project-defined) and there is an order (from
    // the second launch over all links on); a launch without an order goes to lane refill, which also records the counts
    // the next launch is sorted by.  Any model takes sorted tiles when asked (HLM_SCHEDULE_SORTED_TILES).
    const bool has_inflow = c->uid == hlm::Model200::UID;
    const bool divergent = c->routed || divergent_model;
    const bool lanes_asked = c->schedule == HLM_SCHEDULE_LANES || (c->schedule == HLM_SCHEDULE_AUTO && divergent);
    const bool sorted_tiles = !f32 && (c->schedule == HLM_SCHEDULE_SORTED_TILES || (has_inflow && c->schedule == HLM_SCHEDULE_AUTO && divergent));
    hlm::WindowArgs a = a_in;
    // Longest first.  Under lane refill a launch ends with the lanes that drew a long link late while the others have
    // run out of links (22 of 32 threads active on the routed workload).  A link's attempt count changes slowly from
    // one launch to the next, so the launch deals the links in the order of the attempts they took last time, most
    // first.  One stable 4-bit radix pass (attempts / 4, up to 60): links with like counts keep their ascending order,
    // so neighbouring lanes still touch neighbouring memory.  Only for launches over all links of the session.
    // Sorted tiles use the same order, sorted by the whole count (6 bits): a tile is 32 links of EQUAL predicted count.
    if ((lanes_asked || sorted_tiles) && c->longest_first && !f32 && a.tile_lo == 0 && a.n_tiles == (c->ns + 31) / 32 && c->ns < (1LL << 31)) {
        const int key_lo = sorted_tiles ? 0 : 2;
        // Sorted tiles, block by block.  A tile's 32 links are then 32 different places of every column, and sorted over
        // all links the rest of each 32-byte sector they touch belongs to links of other counts, which other tiles
        // reach much later: 1.3 KB of DRAM reads per link instead of 0.3 (ncu, routed interval of 2.5 M links).  So the
        // order is by attempts INSIDE blocks of 32 768 consecutive links (10 MB of state and parameters: L2-resident
        // while the launch works through the block) and by block before that: routed interval 0.965 -> 0.817 ms.  Only
        // for routed runs: an unrouted launch is a day, a link's columns are read once per ~30 attempts, and what
        // counts there is that the longest tiles of ALL links start first (Model 200 day 2.97 ms, by blocks 3.36).
        int key_hi = 6;
        static const int block_shift = [] {  // HLM_TUNE_COST_BLOCK_SHIFT (environment): for block-size experiments only
            const char* e = std::getenv("HLM_TUNE_COST_BLOCK_SHIFT");
            return e ? std::max(10, std::min(24, std::atoi(e))) : 15;
        }();
        a.cost_block_shift = block_shift;
        a.cost_blocks = 1;
        if (sorted_tiles && c->routed) {
            a.cost_blocks = (int)((c->ns + (1LL << block_shift) - 1) >> block_shift);
            for (int b = a.cost_blocks - 1; b > 0; b >>= 1) ++key_hi;
        }
        const size_t n = (size_t)c->ns;
        if (c->cost.cap < (size_t)c->ld) {  // (the kernels read a link's previous word: defined from the start)
            HLM_CUDA(c->cost.reserve((size_t)c->ld));
            HLM_CUDA(cudaMemsetAsync(c->cost.p, 0, (size_t)c->ld * sizeof(int), c->stream));
        }
        // In routed runs the order serves several launches: the sort is 46 us (60 with the block bits) of a coupling
        // interval of about a millisecond, and an order a few intervals old deals the links nearly as well (any
        // permutation is a valid order).  Lane refill: every fourth launch (routed hour 4.60 -> 4.52 ms); sorted tiles,
        // which depend on the counts being right: every second (3.72 -> 3.62 ms; every third or fourth 3.63).
        // Unrouted launches are long (a day of Model 200) and their counts move more: every launch.
        static const int age_tune = [] {  // HLM_TUNE_SORT_AGE (environment): launches per order, for experiments only
            const char* e = std::getenv("HLM_TUNE_SORT_AGE");
            return e ? std::max(1, std::atoi(e)) : 0;
        }();
        const int max_age = age_tune ? age_tune : (c->routed ? (sorted_tiles ? 2 : 4) : 1);
        if (c->cost_ns == c->ns && (c->order_ns != c->ns || c->order_age >= max_age)) {
            HLM_CUDA(c->cost_sorted.reserve(n));
            HLM_CUDA(c->order.reserve(n));
            if (c->iota.cap < n) {
                HLM_CUDA(c->iota.reserve(n));
                iota_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->iota.p, (long long)n);
                HLM_CUDA(cudaGetLastError());
                ++c->launches;
            }
            size_t tmp_bytes = 0;
            HLM_CUDA(cub::DeviceRadixSort::SortPairsDescending(nullptr, tmp_bytes, c->cost.p, c->cost_sorted.p, c->iota.p, c->order.p,
                                                               (int)n, key_lo, key_hi, c->stream));
            HLM_CUDA(c->sort_tmp.reserve(tmp_bytes));
            HLM_CUDA(cub::DeviceRadixSort::SortPairsDescending(c->sort_tmp.p, tmp_bytes, c->cost.p, c->cost_sorted.p, c->iota.p, c->order.p,
                                                               (int)n, key_lo, key_hi, c->stream));
            c->launches += 2 + (key_hi - key_lo + 7) / 8;  // histogram, scan, one scatter pass per 8 key bits
            c->order_ns = c->ns;
            c->order_age = 0;
        }
        if (c->order_ns == c->ns && c->cost_ns == c->ns) {
            a.order = c->order.p;
            ++c->order_age;
        }
        a.cost = c->cost.p;
        c->cost_ns = c->ns;
    }
    HLM_CUDA(cudaMemsetAsync(c->tile_counter.p, 0, sizeof(unsigned int), c->stream));
    // per-launch timing for hlm_kernel_time_ms: a caller that never asks (a routed run queues ~10^5 launches a year)
    // must not accumulate events — beyond a bound the oldest pair is recycled
    if (c->timing.size() >= 1024) {
        c->event_pool.push_back(c->timing.front().first);
        c->event_pool.push_back(c->timing.front().second);
        c->timing.erase(c->timing.begin());
    }
    cudaEvent_t e0 = get_event(c), e1 = get_event(c);
    HLM_CUDA(cudaEventRecord(e0, c->stream));
    // routed runs take a few attempts per link per launch: the lane kernel that tests for "finished" right
    // after the attempt (rk45_window.cuh, kEarlyLeave)
    const bool lanes = (lanes_asked || sorted_tiles) && !(sorted_tiles && a.order != nullptr);
    HLM_CUDA(launch(lanes ? (c->routed ? 2 : 1) : ((sorted_tiles && a.order != nullptr) ? 3 : 0), a, c->sm_count, c->stream));
    HLM_CUDA(cudaEventRecord(e1, c->stream));
    c->timing.emplace_back(e0, e1);
    ++c->launches;
    return 0;
}

// links flagged stiff by the window kernel just queued -> list -> implicit integration of the same window
int dispatch_radau(hlm_ctx* c, const hlm::WindowArgs& a) {
    HLM_CUDA(hlm::radau_launch(c->uid, a, c->radau_list.p, c->radau_count.p, c->n_radau.p, c->sm_count, c->stream));
    c->launches += 2;
    return 0;
}

}  // namespace

// =================================================================================================
extern "C" {

int hlm_abi_version(void) { return 4; }

const char* hlm_last_error(void) { return g_err.c_str(); }

int hlm_create(int device, hlm_ctx** out) {
    HLM_REQUIRE(out != nullptr, "hlm_create: out is NULL");
    *out = nullptr;
    int n = 0;
    HLM_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(HLM_ERR_INVALID, "hlm_create: no such CUDA device");
    HLM_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HLM_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(HLM_ERR_CUDA, std::string("hlm_create: device '") + prop.name +
                                      "' is not sm_100-class; this library ships sm_100a code only and has no fallback");
    hlm_ctx* c = new (std::nothrow) hlm_ctx();
    if (!c) return fail(HLM_ERR_NOMEM, "hlm_create: out of host memory");
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&c->ev_kernel_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_copy_done[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        hlm_destroy(c);
        return fail(HLM_ERR_CUDA, std::string("hlm_create: ") + cudaGetErrorString(e));
    }
    c->stream = c->own_stream;
    // reference defaults, models/model_204.hpp:22-30
    const hlm::SolverParams def = {0.01, 1e-6, 1e-9, 0.9, 0.2, 10.0};
    for (const auto& m : kModels) c->params[m.uid] = def;
    *out = c;
    return HLM_OK;
}

void hlm_destroy(hlm_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    // every device buffer of the context (a process that creates and destroys contexts must not leak)
    c->io_stage.release();
    c->sp_aos.release(); c->sp_soa.release(); c->forc[0].release(); c->forc[1].release(); c->forc_col.release();
    c->y.release(); c->t.release(); c->h.release(); c->tq.release(); c->next_q.release(); c->reject_run.release();
    c->status.release(); c->n_acc.release(); c->n_rej.release(); c->n_jump.release(); c->tile_counter.release();
    c->totals.release(); c->dense[0].release(); c->dense[1].release();
    c->radau_list.release(); c->radau_count.release(); c->n_radau.release();
    c->cost.release(); c->cost_sorted.release(); c->iota.release(); c->order.release(); c->sort_tmp.release();
    c->up_ptr.release(); c->up_idx.release(); c->send_idx.release(); c->send_slot.release(); c->qin.release(); c->own_send.release();
    c->peer_ptrs.release();
    for (auto& p : c->timing) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (c->ev_kernel_done[i]) cudaEventDestroy(c->ev_kernel_done[i]);
        if (c->ev_copy_done[i]) cudaEventDestroy(c->ev_copy_done[i]);
    }
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    for (void* p : c->peer_opened)
        if (p) cudaIpcCloseMemHandle(p);
    if (c->own_halo) cudaFree(c->own_halo);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->h2d_stream) cudaStreamDestroy(c->h2d_stream);
    for (cudaEvent_t ev : c->chunk_ready) cudaEventDestroy(ev);
    delete c;
}

int hlm_set_stream(hlm_ctx* c, void* s) {
    HLM_REQUIRE(c, "hlm_set_stream: ctx is NULL");
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return HLM_OK;
}

int hlm_synchronize(hlm_ctx* c) {
    HLM_REQUIRE(c, "hlm_synchronize: ctx is NULL");
    if (int r = use_device(c)) return r;
    HLM_CUDA(cudaStreamSynchronize(c->stream));
    HLM_CUDA(cudaStreamSynchronize(c->copy_stream));
    return HLM_OK;
}

int hlm_model_info(int uid, int* n_eq, int* n_sp, int* n_forc) {
    const ModelInfo* m = find_model(uid);
    if (!m) return fail(HLM_ERR_INVALID, "hlm_model_info: unknown model uid " + std::to_string(uid));
    if (n_eq) *n_eq = m->n_eq;
    if (n_sp) *n_sp = m->n_sp;
    if (n_forc) *n_forc = m->n_forc;
    return HLM_OK;
}

int hlm_set_model_parameters(hlm_ctx* c, int uid, const double p[6]) {
    HLM_REQUIRE(c && p, "hlm_set_model_parameters: NULL argument");
    if (!find_model(uid)) return fail(HLM_ERR_INVALID, "hlm_set_model_parameters: unknown model uid " + std::to_string(uid));
    c->params[uid] = hlm::SolverParams{p[0], p[1], p[2], p[3], p[4], p[5]};
    return HLM_OK;
}

int hlm_get_model_parameters(hlm_ctx* c, int uid, double p[6]) {
    HLM_REQUIRE(c && p, "hlm_get_model_parameters: NULL argument");
    auto it = c->params.find(uid);
    if (it == c->params.end()) return fail(HLM_ERR_INVALID, "hlm_get_model_parameters: unknown model uid");
    const hlm::SolverParams& s = it->second;
    p[0] = s.initialStep; p[1] = s.rtol; p[2] = s.atol; p[3] = s.safety; p[4] = s.minScale; p[5] = s.maxScale;
    return HLM_OK;
}

int hlm_upload_spatial_params(hlm_ctx* c, const void* aos, long long n, long long stride) {
    HLM_REQUIRE(c, "hlm_upload_spatial_params: ctx is NULL");
    HLM_REQUIRE(n >= 0 && (n == 0 || aos), "hlm_upload_spatial_params: NULL records");
    HLM_REQUIRE(stride >= (long long)sizeof(hlm::SpatialParamsAoS) && stride % 8 == 0,
                "hlm_upload_spatial_params: stride must be >= 136 and a multiple of 8");
    if (int r = use_device(c)) return r;
    const size_t rec = sizeof(hlm::SpatialParamsAoS);
    HLM_CUDA(c->sp_aos.reserve((size_t)std::max<long long>(n, 1) * rec));
    if (n > 0)
        // cudaMemcpyDefault: `aos` may be host memory or — what the reference's run_rk45 is handed, a cudaMalloc'ed
        // array (main.cpp:392-404) — device memory; unified addressing tells which
        HLM_CUDA(cudaMemcpy2DAsync(c->sp_aos.p, rec, aos, (size_t)stride, rec, (size_t)n, cudaMemcpyDefault, c->stream));
    c->sp_n = n;
    c->sp_soa_uid = -1;  // columns are rebuilt on the next solve
    return HLM_OK;
}

int hlm_upload_forcing_chunk(hlm_ctx* c, int j, double dt_hours, long long nT_total, long long i0, long long nT_chunk,
                             long long ncols, const float* data) {
    HLM_REQUIRE(c && data, "hlm_upload_forcing: NULL argument");
    HLM_REQUIRE(j >= 0 && j < hlm::kMaxForcings, "hlm_upload_forcing: forcing index out of range [0,16)");
    HLM_REQUIRE(nT_total > 0 && ncols > 0, "hlm_upload_forcing: nT and ncols must be positive");
    HLM_REQUIRE(i0 >= 0 && nT_chunk > 0 && i0 + nT_chunk <= nT_total, "hlm_upload_forcing_chunk: chunk outside the record");
    if (j >= 2) return HLM_OK;  // accepted like the reference's 16 slots, but no compiled model reads F[j>=2]
    HLM_REQUIRE(j <= c->n_forc, "hlm_upload_forcing: upload forcings in order 0,1,...");
    for (int k = 0; k < c->n_forc; ++k)
        if (k != j && c->forc_ncols != ncols)
            return fail(HLM_ERR_INVALID, "hlm_upload_forcing: all forcings must share ncols");
    if (c->forc_col_n != 0 && c->forc_col_max >= ncols)
        return fail(HLM_ERR_INVALID, "hlm_upload_forcing: the forcing column map refers to column " + std::to_string(c->forc_col_max) +
                                         " but this forcing has " + std::to_string(ncols) + " columns");
    if (int r = use_device(c)) return r;
    // a larger chunk reallocates: the window kernels queued on the stream must be done with the old buffer
    if ((size_t)nT_chunk * ncols > c->forc[j].cap) HLM_CUDA(cudaStreamSynchronize(c->stream));
    HLM_CUDA(c->forc[j].reserve((size_t)nT_chunk * ncols));
    HLM_CUDA(cudaMemcpyAsync(c->forc[j].p, data, sizeof(float) * (size_t)nT_chunk * ncols, cudaMemcpyHostToDevice, c->stream));
    c->forc_nT[j] = nT_total;
    c->forc_i0[j] = i0;
    c->forc_nres[j] = nT_chunk;
    c->forc_dt_h[j] = dt_hours;
    c->forc_ncols = ncols;
    if (j == c->n_forc) c->n_forc = j + 1;
    return HLM_OK;
}

int hlm_upload_forcing(hlm_ctx* c, int j, double dt_hours, long long nT, long long ncols, const float* data) {
    return hlm_upload_forcing_chunk(c, j, dt_hours, nT, 0, nT, ncols, data);
}

int hlm_set_forcing_columns(hlm_ctx* c, const int* col, long long n) {
    HLM_REQUIRE(c, "hlm_set_forcing_columns: ctx is NULL");
    if (!col || n == 0) { c->forc_col_n = 0; return HLM_OK; }
    // a column outside the forcing arrays would be an out-of-bounds device read in every kernel
    long long lo = col[0], hi = col[0];
    for (long long i = 1; i < n; ++i) {
        lo = std::min<long long>(lo, col[i]);
        hi = std::max<long long>(hi, col[i]);
    }
    if (lo < 0) return fail(HLM_ERR_INVALID, "hlm_set_forcing_columns: negative column index");
    if (c->n_forc > 0 && hi >= c->forc_ncols)
        return fail(HLM_ERR_INVALID, "hlm_set_forcing_columns: column " + std::to_string(hi) + " but the forcings have " +
                                         std::to_string(c->forc_ncols) + " columns");
    c->forc_col_min = lo;
    c->forc_col_max = hi;
    if (int r = use_device(c)) return r;
    const long long ld = (n + 31) / 32 * 32;
    HLM_CUDA(c->forc_col.reserve((size_t)ld));
    HLM_CUDA(cudaMemsetAsync(c->forc_col.p, 0, sizeof(int) * (size_t)ld, c->stream));
    HLM_CUDA(cudaMemcpyAsync(c->forc_col.p, col, sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    c->forc_col_n = n;
    return HLM_OK;
}

int hlm_clear_forcings(hlm_ctx* c) {
    HLM_REQUIRE(c, "hlm_clear_forcings: ctx is NULL");
    c->n_forc = 0;
    c->forc_col_n = 0;
    c->forc_ncols = 0;
    return HLM_OK;
}

int hlm_set_max_attempts(hlm_ctx* c, long long v) {
    HLM_REQUIRE(c, "hlm_set_max_attempts: ctx is NULL");
    c->max_attempts = v;
    return HLM_OK;
}

int hlm_set_reject_limit(hlm_ctx* c, int n) {
    HLM_REQUIRE(c && n >= 0, "hlm_set_reject_limit: need n >= 0");
    c->reject_limit = n;
    return HLM_OK;
}

int hlm_set_dense_window_bytes(hlm_ctx* c, long long v) {
    HLM_REQUIRE(c && v > 0, "hlm_set_dense_window_bytes: need a positive size");
    c->dense_window_bytes = v;
    return HLM_OK;
}

int hlm_set_stiff_fallback(hlm_ctx* c, int enable) {
    HLM_REQUIRE(c, "hlm_set_stiff_fallback: ctx is NULL");
    c->stiff_fallback = enable != 0;
    return HLM_OK;
}

int hlm_set_schedule(hlm_ctx* c, int mode) {
    HLM_REQUIRE(c && (mode == HLM_SCHEDULE_AUTO || mode == HLM_SCHEDULE_TILES || mode == HLM_SCHEDULE_LANES || mode == HLM_SCHEDULE_SORTED_TILES),
                "hlm_set_schedule: mode must be HLM_SCHEDULE_AUTO, _TILES, _LANES or _SORTED_TILES");
    c->schedule = mode;
    return HLM_OK;
}

int hlm_set_precision(hlm_ctx* c, int bits) {
    HLM_REQUIRE(c && (bits == 64 || bits == 32), "hlm_set_precision: bits must be 64 or 32");
    c->precision = bits;
    return HLM_OK;
}

int hlm_set_output_states(hlm_ctx* c, unsigned int mask) {
    HLM_REQUIRE(c, "hlm_set_output_states: ctx is NULL");
    c->out_mask = mask;
    return HLM_OK;
}

int hlm_set_output_precision(hlm_ctx* c, int bits) {
    HLM_REQUIRE(c && (bits == 64 || bits == 32), "hlm_set_output_precision: bits must be 64 or 32");
    c->out_bits = bits;
    return HLM_OK;
}

int hlm_output_layout(hlm_ctx* c, int uid, int* n_columns, int* bytes_per_value) {
    HLM_REQUIRE(c, "hlm_output_layout: ctx is NULL");
    const ModelInfo* m = find_model(uid);
    if (!m) return fail(HLM_ERR_INVALID, "hlm_output_layout: unknown model uid " + std::to_string(uid));
    const unsigned int all = (1u << m->n_eq) - 1u;
    const unsigned int eff = (c->out_mask & all) ? (c->out_mask & all) : all;
    if (n_columns) *n_columns = __builtin_popcount(eff);
    if (bytes_per_value) *bytes_per_value = c->out_bits == 32 ? 4 : 8;
    return HLM_OK;
}

long long hlm_launch_count(hlm_ctx* c) { return c ? c->launches : 0; }

int hlm_kernel_time_ms(hlm_ctx* c, double* sum_ms, long long* n) {
    HLM_REQUIRE(c, "hlm_kernel_time_ms: ctx is NULL");
    if (int r = use_device(c)) return r;
    HLM_CUDA(cudaStreamSynchronize(c->stream));
    double s = 0;
    for (auto& p : c->timing) {
        float ms = 0;
        HLM_CUDA(cudaEventElapsedTime(&ms, p.first, p.second));
        s += ms;
        c->event_pool.push_back(p.first);
        c->event_pool.push_back(p.second);
    }
    if (sum_ms) *sum_ms = s;
    if (n) *n = (long long)c->timing.size();
    c->timing.clear();
    return HLM_OK;
}

// ---- session ------------------------------------------------------------------------------------

static int solve_begin_impl(hlm_ctx* c, int uid, const double* y0, long long ns, double t0, double tf, const double* tq,
                            long long nq, bool defer_state);

int hlm_solve_begin(hlm_ctx* c, int uid, const double* y0, long long ns, double t0, double tf, const double* tq,
                    long long nq) {
    return solve_begin_impl(c, uid, y0, ns, t0, tf, tq, nq, false);
}

// defer_state: the caller uploads y0 and initialises the state itself, chunk by chunk (run_link_chunks)
static int solve_begin_impl(hlm_ctx* c, int uid, const double* y0, long long ns, double t0, double tf, const double* tq,
                            long long nq, bool defer_state) {
    HLM_REQUIRE(c, "hlm_solve_begin: ctx is NULL");
    const ModelInfo* m = find_model(uid);
    if (!m) return fail(HLM_ERR_INVALID, "hlm_solve_begin: unknown model uid " + std::to_string(uid));
    HLM_REQUIRE(ns > 0 && y0, "hlm_solve_begin: need ns > 0 and y0");
    HLM_REQUIRE(ns < (1LL << 31) - 64, "hlm_solve_begin: ns must be below 2^31 per context");
    HLM_REQUIRE(nq >= 0 && nq < (1LL << 31) && (nq == 0 || tq), "hlm_solve_begin: bad query times");
    if (m->n_forc > 0 && c->n_forc > 0 && c->forc_col_n == 0 && c->forc_ncols < ns)
        return fail(HLM_ERR_INVALID, "hlm_solve_begin: forcing has fewer columns than links and no column map");
    if (m->n_forc > 0 && c->forc_col_n != 0 && c->forc_col_n != ns)
        return fail(HLM_ERR_INVALID, "hlm_solve_begin: forcing column map length differs from ns");
    if (m->n_forc > 0 && c->n_forc > 0 && c->forc_col_n != 0 && c->forc_col_max >= c->forc_ncols)
        return fail(HLM_ERR_INVALID, "hlm_solve_begin: the forcing column map refers to column " + std::to_string(c->forc_col_max) +
                                         " but the forcings have " + std::to_string(c->forc_ncols) + " columns");
    if (int r = use_device(c)) return r;
    // a previous call's D2H copies may still read the window buffers
    HLM_CUDA(cudaStreamSynchronize(c->copy_stream));
    c->copy_pending[0] = c->copy_pending[1] = false;
    c->in_session = false;
    c->uid = uid;
    c->n_eq = m->n_eq;
    c->ns = ns;
    c->ld = (ns + 31) / 32 * 32;
    c->nq = nq;
    c->t0 = t0;
    c->tf = tf;
    c->q_done = 0;
    c->win_q_lo = c->win_q_hi = 0;
    c->win_has_dense = false;
    c->cost_ns = 0;  // no attempt counts yet for these links
    c->order_ns = 0;
    const size_t ld = (size_t)c->ld;
    HLM_CUDA(c->y.reserve(ld * m->n_eq));
    HLM_CUDA(c->t.reserve(ld));
    HLM_CUDA(c->h.reserve(ld));
    HLM_CUDA(c->next_q.reserve(ld));
    HLM_CUDA(c->reject_run.reserve(ld));
    HLM_CUDA(c->status.reserve(ld));
    HLM_CUDA(c->n_acc.reserve(ld));
    HLM_CUDA(c->n_rej.reserve(ld));
    HLM_CUDA(c->n_jump.reserve(ld));
    HLM_CUDA(c->tile_counter.reserve(1));
    HLM_CUDA(c->totals.reserve(8));
    HLM_CUDA(c->radau_list.reserve(ld));
    HLM_CUDA(c->radau_count.reserve(1));
    HLM_CUDA(c->n_radau.reserve(ld));
    HLM_CUDA(cudaMemsetAsync(c->n_radau.p, 0, sizeof(unsigned int) * ld, c->stream));
    HLM_CUDA(c->tq.reserve((size_t)std::max<long long>(nq, 1)));
    if (nq > 0) HLM_CUDA(cudaMemcpyAsync(c->tq.p, tq, sizeof(double) * (size_t)nq, cudaMemcpyHostToDevice, c->stream));
    if (!defer_state) {
        // y0 arrives [link][state]; stage it in the (not yet used) dense buffer, then transpose to columns
        HLM_CUDA(c->dense[0].reserve((size_t)ns * m->n_eq));
        HLM_CUDA(cudaMemcpyAsync(c->dense[0].p, y0, sizeof(double) * (size_t)ns * m->n_eq, cudaMemcpyHostToDevice, c->stream));
        const hlm::SolverParams& prm = c->params[uid];
        const int tpb = 256;
        init_state_kernel<<<(unsigned)((c->ld + tpb - 1) / tpb), tpb, 0, c->stream>>>(
            c->dense[0].p, m->n_eq, ns, c->ld, c->y.p, c->t.p, c->h.p, c->next_q.p, c->reject_run.p, c->status.p,
            c->n_acc.p, c->n_rej.p, c->n_jump.p, t0, prm.initialStep, 0, c->ld);
        HLM_CUDA(cudaGetLastError());
        ++c->launches;
    }
    // parameter columns: now when the records are here; otherwise with the first window — the reference hands d_sp to
    // its launch, not to setup_gpu_buffers (solver/rk45_api.hpp:63-66,120-132)
    if (c->sp_n == c->ns)
        if (int r = prepare_model_params(c)) return r;
    c->in_session = true;
    return HLM_OK;
}

static int restart_impl(hlm_ctx* c, double t0, double tf, const double* tq, long long nq, bool keep_step);

int hlm_solve_restart(hlm_ctx* c, double t0, double tf, const double* tq, long long nq) {
    return restart_impl(c, t0, tf, tq, nq, false);
}

int hlm_solve_advance(hlm_ctx* c, double tf, const double* tq, long long nq) {
    HLM_REQUIRE(c, "hlm_solve_advance: ctx is NULL");
    if (!c->in_session) return fail(HLM_ERR_STATE, "hlm_solve_advance: no session (call hlm_solve_begin)");
    HLM_REQUIRE(tf >= c->tf, "hlm_solve_advance: tf must not move backwards");
    return restart_impl(c, c->tf, tf, tq, nq, true);
}

static int restart_impl(hlm_ctx* c, double t0, double tf, const double* tq, long long nq, bool keep_step) {
    HLM_REQUIRE(c, "hlm_solve_restart: ctx is NULL");
    if (!c->in_session) return fail(HLM_ERR_STATE, "hlm_solve_restart: no session (call hlm_solve_begin)");
    HLM_REQUIRE(nq >= 0 && nq < (1LL << 31) && (nq == 0 || tq), "hlm_solve_restart: bad query times");
    if (int r = use_device(c)) return r;
    HLM_CUDA(c->tq.reserve((size_t)std::max<long long>(nq, 1)));
    if (nq > 0) HLM_CUDA(cudaMemcpyAsync(c->tq.p, tq, sizeof(double) * (size_t)nq, cudaMemcpyHostToDevice, c->stream));
    const int tpb = 256;
    // padding lanes beyond ns carry status kDone from init and are re-armed here; the window kernel
    // never touches them (sys >= ns), so that is harmless
    restart_state_kernel<<<(unsigned)((c->ld + tpb - 1) / tpb), tpb, 0, c->stream>>>(
        c->ld, c->t.p, c->h.p, c->next_q.p, c->reject_run.p, c->status.p, t0, c->params[c->uid].initialStep, keep_step);
    HLM_CUDA(cudaGetLastError());
    ++c->launches;
    c->t0 = t0;
    c->tf = tf;
    c->nq = nq;
    c->q_done = 0;
    c->win_q_lo = c->win_q_hi = 0;
    c->win_has_dense = false;
    return HLM_OK;
}

// Queue the window kernel (and the implicit fallback, if on) for queries [q_lo, q_hi) over the links of tiles
// [tile_lo, tile_lo + n_tiles); `dense` (may be NULL) receives the records, row 0 = link dense_sys0.
static int queue_window(hlm_ctx* c, long long q_lo, long long q_hi, void* dense, long long tile_lo, long long n_tiles,
                        long long dense_sys0) {
    if (int r = prepare_model_params(c)) return r;  // a no-op unless parameters were (re)uploaded since the last launch
    hlm::WindowArgs a;
    std::memset(&a, 0, sizeof(a));
    a.y = c->y.p; a.t = c->t.p; a.h = c->h.p; a.next_q = c->next_q.p; a.reject_run = c->reject_run.p;
    a.status = c->status.p; a.n_accept = c->n_acc.p; a.n_reject = c->n_rej.p; a.n_jump = c->n_jump.p;
    a.sp = c->sp_soa.p;
    a.col = c->forc_col_n ? c->forc_col.p : nullptr;
    a.n_forc = c->n_forc;
    for (int j = 0; j < 2; ++j) {
        a.forc[j] = c->forc[j].p;
        a.forc_nT[j] = c->forc_nT[j];
        a.forc_i0[j] = c->forc_i0[j];
        a.forc_nres[j] = c->forc_nres[j];
        a.forc_dt_min[j] = c->forc_dt_h[j] * 60.0;  // rk45_kernel.cu:90
    }
    a.forc_ncols = c->forc_ncols;
    {   // the samples every link starts the launch with, when they all start at t0 (LinkRun::load); the same
        // expressions as the device's forcing_index, in IEEE double on both sides
        const ModelInfo* mi = find_model(c->uid);
        const int nf = mi ? std::min(std::min(c->n_forc, mi->n_forc), 2) : 0;
        double lo = -INFINITY, hi = INFINITY;
        bool ok = nf > 0;
        for (int j = 0; j < nf && ok; ++j) {
            const double dtm = a.forc_dt_min[j];
            const long long nT = a.forc_nT[j];
            if (!(dtm > 0.0) || nT <= 0) { ok = false; break; }
            const double r = c->t0 / dtm;
            const long long idx = (r < 0.0) ? 0 : ((r >= (double)nT) ? nT - 1 : (long long)r);
            const double l = (idx <= 0) ? -INFINITY : (double)idx * dtm * (1.0 + 1e-9);
            const double h = (idx >= nT - 1) ? INFINITY : (double)(idx + 1) * dtm * (1.0 - 1e-9);
            long long row = idx - a.forc_i0[j];
            row = row < 0 ? 0 : (row >= a.forc_nres[j] ? a.forc_nres[j] - 1 : row);
            a.forc_pre_row[j] = row;
            lo = std::max(lo, l);
            hi = std::min(hi, h);
        }
        a.forc_pre_ok = ok ? 1 : 0;
        a.forc_pre_lo = lo;
        a.forc_pre_hi = hi;
    }
    a.tq = c->tq.p;
    a.nq = (int)c->nq;
    a.q_lo = (int)q_lo;
    a.q_hi = (int)q_hi;
    a.dense = dense;
    a.dense_mask = dense_mask(c);
    a.dense_ncol = dense_ncol(c);
    a.dense_f32 = c->out_bits == 32 ? 1 : 0;
    a.t0 = c->t0; a.tf = c->tf;
    a.prm = c->params[c->uid];
    a.ns = c->ns; a.ld = c->ld;
    a.tile_lo = tile_lo; a.n_tiles = n_tiles; a.dense_sys0 = dense_sys0;
    a.max_attempts = c->max_attempts;
    a.reject_limit = c->reject_limit;
    a.tile_counter = c->tile_counter.p;
    if (c->routed) {
        if (c->route_ns != c->ns) return fail(HLM_ERR_STATE, "hlm_solve_window: routing topology was set for another link count");
        a.qin = c->qin.p;
        a.send_slot = c->route_n_send > 0 ? c->send_slot.p : nullptr;
        a.send_buf = c->send_buf;
        if (c->peer_world > 1) {  // results of this interval are read at parity route_epoch % 2 by the next gather
            a.peer_halo = c->peer_ptrs.p;
            a.peer_world = c->peer_world;
            a.peer_off = (c->route_epoch % 2) * (long long)c->peer_world * c->peer_max_send + (long long)c->peer_rank * c->peer_max_send;
        }
    }
    if (int r = dispatch_window(c, a)) return r;
    if (c->stiff_fallback)
        if (int r = dispatch_radau(c, a)) return r;
    return 0;
}

// a chunked forcing record must cover every sample the interval [t0, tf] can index
// (size_t(t / (dt*60)) clamped to the record, solver/rk45_kernel.cu:90-98)
static int check_forcing_cover(hlm_ctx* c) {
    if (const ModelInfo* m = find_model(c->uid))
        for (int j = 0; j < std::min(c->n_forc, m->n_forc); ++j) {
            const double dtm = c->forc_dt_h[j] * 60.0;
            if (!(dtm > 0.0) || c->forc_nres[j] == c->forc_nT[j]) continue;
            auto index = [&](double t) {
                const double r = t / dtm;
                return (r < 0.0) ? 0LL : (r >= (double)c->forc_nT[j] ? c->forc_nT[j] - 1 : (long long)r);
            };
            if (index(c->t0) < c->forc_i0[j] || index(c->tf) >= c->forc_i0[j] + c->forc_nres[j])
                return fail(HLM_ERR_STATE, "hlm_solve_window: the resident chunk of forcing " + std::to_string(j) +
                                               " does not cover the interval");
        }
    return 0;
}

int hlm_solve_window(hlm_ctx* c, long long q_hi, int want_dense) {
    HLM_REQUIRE(c, "hlm_solve_window: ctx is NULL");
    if (!c->in_session) return fail(HLM_ERR_STATE, "hlm_solve_window: no session (call hlm_solve_begin)");
    if (int r = use_device(c)) return r;
    if (q_hi > c->nq) q_hi = c->nq;
    HLM_REQUIRE(q_hi >= c->q_done, "hlm_solve_window: q_hi must not move backwards");
    if (int r = check_forcing_cover(c)) return r;
    const long long q_lo = c->q_done;
    const long long qw = q_hi - q_lo;
    const bool dense = want_dense && qw > 0;
    int buf = c->dense_cur;
    if (dense) {
        buf = c->dense_cur ^ 1;
        if (c->copy_pending[buf]) {  // the D2H that last read this buffer must finish before it is overwritten
            HLM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy_done[buf], 0));
            c->copy_pending[buf] = false;
        }
        // no memset: the window kernel writes every slot, zeros where a link never gets (rk45_window.cuh)
        HLM_CUDA(c->dense[buf].reserve(doubles_for((size_t)c->ns * (size_t)qw * dense_record_bytes(c))));
        c->dense_cur = buf;
    }
    if (int r = queue_window(c, q_lo, q_hi, dense ? c->dense[buf].p : nullptr, 0, (c->ns + 31) / 32, 0)) return r;
    c->q_done = q_hi;
    c->win_q_lo = q_lo;
    c->win_q_hi = q_hi;
    c->win_has_dense = dense;
    if (dense) HLM_CUDA(cudaEventRecord(c->ev_kernel_done[buf], c->stream));
    return HLM_OK;
}

int hlm_solve_window_buffer(hlm_ctx* c, void** dev_ptr, long long* q_lo, long long* q_hi) {
    HLM_REQUIRE(c, "hlm_solve_window_buffer: ctx is NULL");
    if (!c->in_session) return fail(HLM_ERR_STATE, "hlm_solve_window_buffer: no session");
    if (dev_ptr) *dev_ptr = c->win_has_dense ? (void*)c->dense[c->dense_cur].p : nullptr;
    if (q_lo) *q_lo = c->win_q_lo;
    if (q_hi) *q_hi = c->win_q_hi;
    return HLM_OK;
}

int hlm_solve_fetch_window(hlm_ctx* c, void* host_dense) {
    HLM_REQUIRE(c && host_dense, "hlm_solve_fetch_window: NULL argument");
    if (!c->in_session) return fail(HLM_ERR_STATE, "hlm_solve_fetch_window: no session");
    if (!c->win_has_dense) return HLM_OK;
    if (int r = use_device(c)) return r;
    const int buf = c->dense_cur;
    const long long qw = c->win_q_hi - c->win_q_lo;
    const size_t row = (size_t)qw * dense_record_bytes(c);         // one link's records of this window
    const size_t host_pitch = (size_t)c->nq * dense_record_bytes(c);  // one link's records of the run
    HLM_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_kernel_done[buf], 0));
    char* dst = reinterpret_cast<char*>(host_dense) + (size_t)c->win_q_lo * dense_record_bytes(c);
    if (row == host_pitch)
        HLM_CUDA(cudaMemcpyAsync(dst, c->dense[buf].p, row * (size_t)c->ns, cudaMemcpyDeviceToHost, c->copy_stream));
    else
        HLM_CUDA(cudaMemcpy2DAsync(dst, host_pitch, c->dense[buf].p, row, row, (size_t)c->ns, cudaMemcpyDeviceToHost,
                                   c->copy_stream));
    HLM_CUDA(cudaEventRecord(c->ev_copy_done[buf], c->copy_stream));
    c->copy_pending[buf] = true;
    return HLM_OK;
}

int hlm_solve_fetch_window_packed(hlm_ctx* c, void* host_win, int* ticket) {
    HLM_REQUIRE(c && host_win, "hlm_solve_fetch_window_packed: NULL argument");
    if (!c->in_session) return fail(HLM_ERR_STATE, "hlm_solve_fetch_window_packed: no session");
    if (ticket) *ticket = c->dense_cur;
    if (!c->win_has_dense) return HLM_OK;
    if (int r = use_device(c)) return r;
    const int buf = c->dense_cur;
    const size_t bytes = (size_t)c->ns * (size_t)(c->win_q_hi - c->win_q_lo) * dense_record_bytes(c);
    HLM_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_kernel_done[buf], 0));
    HLM_CUDA(cudaMemcpyAsync(host_win, c->dense[buf].p, bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    HLM_CUDA(cudaEventRecord(c->ev_copy_done[buf], c->copy_stream));
    c->copy_pending[buf] = true;
    return HLM_OK;
}

int hlm_solve_wait_copy(hlm_ctx* c, int ticket) {
    HLM_REQUIRE(c, "hlm_solve_wait_copy: ctx is NULL");
    if (int r = use_device(c)) return r;
    if (ticket == 0 || ticket == 1) HLM_CUDA(cudaEventSynchronize(c->ev_copy_done[ticket]));
    else HLM_CUDA(cudaStreamSynchronize(c->copy_stream));
    return HLM_OK;
}

int hlm_host_alloc(void** out, long long bytes) {
    HLM_REQUIRE(out && bytes > 0, "hlm_host_alloc: bad argument");
    HLM_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable));
    return HLM_OK;
}

int hlm_host_free(void* p) {
    if (p) HLM_CUDA(cudaFreeHost(p));
    return HLM_OK;
}

int hlm_solve_totals(hlm_ctx* c, long long totals[7]) {
    HLM_REQUIRE(c && totals, "hlm_solve_totals: NULL argument");
    if (!c->in_session) return fail(HLM_ERR_STATE, "hlm_solve_totals: no session");
    if (int r = use_device(c)) return r;
    HLM_CUDA(cudaMemsetAsync(c->totals.p, 0, 8 * sizeof(unsigned long long), c->stream));
    const int tpb = 256;
    const unsigned grid = (unsigned)std::min<long long>((c->ns + tpb - 1) / tpb, (long long)c->sm_count * 8);
    totals_kernel<<<grid, tpb, 0, c->stream>>>(c->n_acc.p, c->n_rej.p, c->n_jump.p, c->status.p, c->ns, c->totals.p);
    HLM_CUDA(cudaGetLastError());
    ++c->launches;
    unsigned long long h[8];
    HLM_CUDA(cudaMemcpyAsync(h, c->totals.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    HLM_CUDA(cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 7; ++k) totals[k] = (long long)h[k];
    return HLM_OK;
}

int hlm_solve_radau_steps(hlm_ctx* c, long long* out_steps) {
    HLM_REQUIRE(c && out_steps, "hlm_solve_radau_steps: NULL argument");
    if (!c->in_session) return fail(HLM_ERR_STATE, "hlm_solve_radau_steps: no session");
    if (int r = use_device(c)) return r;
    std::vector<unsigned int> tmp;
    try { tmp.resize((size_t)c->ns); } catch (const std::bad_alloc&) { return fail(HLM_ERR_NOMEM, "hlm_solve_radau_steps: out of host memory"); }
    HLM_CUDA(cudaMemcpyAsync(tmp.data(), c->n_radau.p, sizeof(unsigned int) * (size_t)c->ns, cudaMemcpyDeviceToHost, c->stream));
    HLM_CUDA(cudaStreamSynchronize(c->stream));
    for (long long i = 0; i < c->ns; ++i) out_steps[i] = (long long)tmp[(size_t)i];
    return HLM_OK;
}

int hlm_solve_peek(hlm_ctx* c, double* out_t, double* out_h, double* out_y) {
    HLM_REQUIRE(c, "hlm_solve_peek: ctx is NULL");
    if (!c->in_session) return fail(HLM_ERR_STATE, "hlm_solve_peek: no session");
    if (int r = use_device(c)) return r;
    if (out_t) HLM_CUDA(cudaMemcpyAsync(out_t, c->t.p, sizeof(double) * (size_t)c->ns, cudaMemcpyDeviceToHost, c->stream));
    if (out_h) HLM_CUDA(cudaMemcpyAsync(out_h, c->h.p, sizeof(double) * (size_t)c->ns, cudaMemcpyDeviceToHost, c->stream));
    if (out_y)  // raw columns [N_EQ][ns]
        HLM_CUDA(cudaMemcpy2DAsync(out_y, sizeof(double) * (size_t)c->ns, c->y.p, sizeof(double) * (size_t)c->ld,
                                   sizeof(double) * (size_t)c->ns, (size_t)c->n_eq, cudaMemcpyDeviceToHost, c->stream));
    HLM_CUDA(cudaStreamSynchronize(c->stream));
    return HLM_OK;
}

int hlm_solve_end(hlm_ctx* c, double* out_final, int* out_stiff, long long* out_acc, long long* out_rej,
                  long long* out_jump) {
    HLM_REQUIRE(c, "hlm_solve_end: ctx is NULL");
    if (!c->in_session) return fail(HLM_ERR_STATE, "hlm_solve_end: no session");
    if (int r = use_device(c)) return r;
    const long long ns = c->ns;
    if (out_final || out_stiff || out_acc || out_rej || out_jump) {
        // The window buffer not used last stages everything in the caller's formats (final states back in
        // [link][state] order, counters widened to 64 bits, status as HLM_LINK_* codes), so each output is
        // one device-to-host copy straight into the caller's array and the host touches no element.
        const int buf = c->dense_cur ^ 1;
        if (c->copy_pending[buf]) {
            HLM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy_done[buf], 0));
            c->copy_pending[buf] = false;
        }
        const size_t n_final = (size_t)ns * c->n_eq;
        HLM_CUDA(c->dense[buf].reserve(n_final + 3 * (size_t)ns + (size_t)ns / 2 + 1));
        double* d_final = c->dense[buf].p;
        long long* d_cnt = reinterpret_cast<long long*>(d_final + n_final);
        int* d_code = reinterpret_cast<int*>(d_cnt + 3 * (size_t)ns);
        const int tpb = 256;
        export_results_kernel<<<(unsigned)((ns + tpb - 1) / tpb), tpb, 0, c->stream>>>(
            c->y.p, c->status.p, c->n_acc.p, c->n_rej.p, c->n_jump.p, c->n_eq, ns, c->ld, out_final ? d_final : nullptr, d_cnt,
            d_code, 0, ns);
        HLM_CUDA(cudaGetLastError());
        ++c->launches;
        if (out_final) HLM_CUDA(cudaMemcpyAsync(out_final, d_final, sizeof(double) * n_final, cudaMemcpyDeviceToHost, c->stream));
        long long* outs[3] = {out_acc, out_rej, out_jump};
        for (int k = 0; k < 3; ++k)
            if (outs[k])
                HLM_CUDA(cudaMemcpyAsync(outs[k], d_cnt + (size_t)k * ns, sizeof(long long) * (size_t)ns, cudaMemcpyDeviceToHost,
                                         c->stream));
        if (out_stiff) HLM_CUDA(cudaMemcpyAsync(out_stiff, d_code, sizeof(int) * (size_t)ns, cudaMemcpyDeviceToHost, c->stream));
    }
    HLM_CUDA(cudaStreamSynchronize(c->stream));
    HLM_CUDA(cudaStreamSynchronize(c->copy_stream));
    c->copy_pending[0] = c->copy_pending[1] = false;
    return HLM_OK;
}

// All queries of a large run, chunk of links by chunk of links: each chunk's records form one contiguous block
// of the caller's [link][query][state] array, so every device-to-host copy is a plain memcpy at full link
// rate (windows over queries leave as 2-D copies of short rows), and it overlaps the next chunk's integration
// (two buffers, copy stream).  Chunks shrink towards the end: what stays exposed is the copy of a small last one.
// The state travels the same way: y0 of chunk i is uploaded (a third stream; PCIe is full duplex) and
// initialised while earlier chunks integrate, and a chunk's final states, counters and codes are exported and
// copied out right behind its dense records — so the call never waits for a whole-array transfer.
static int run_link_chunks(hlm_ctx* c, const double* y0, void* out_dense, double* out_final, int* out_stiff,
                           long long* out_acc, long long* out_rej, long long* out_jump) {
    const long long ns = c->ns, nq = c->nq;
    const int n_eq = c->n_eq;
    const long long per_link = nq * (long long)dense_record_bytes(c);
    const long long n_tiles_all = (ns + 31) / 32;
    const long long tiles_max = std::max<long long>(1, c->dense_window_bytes / (per_link * 32));
    const long long n_chunks = std::max<long long>((n_tiles_all + tiles_max - 1) / tiles_max, std::min<long long>(8, n_tiles_all));
    if (int r = check_forcing_cover(c)) return r;
    std::vector<long long> ends;  // chunk i = tiles [ends[i-1], ends[i])
    {
        const double total_w = 0.5 * (double)n_chunks * (double)(n_chunks + 1);
        double acc_w = 0.0;
        long long tile = 0;
        for (long long i = 0; tile < n_tiles_all; ++i) {
            acc_w += (double)std::max<long long>(n_chunks - i, 1);
            long long end = (long long)std::llround((double)n_tiles_all * std::min(1.0, acc_w / total_w));
            end = std::min(n_tiles_all, std::max(end, tile + 1));
            end = std::min(end, tile + tiles_max);
            ends.push_back(end);
            tile = end;
        }
    }
    // staging in the caller's layouts: [ns][n_eq] doubles (y0 in, final states out), 3 x [ns] i64, [ns] i32
    const size_t n_state = (size_t)ns * n_eq;
    HLM_CUDA(c->io_stage.reserve(n_state + 3 * (size_t)ns + (size_t)ns / 2 + 1));
    double* d_state = c->io_stage.p;
    long long* d_cnt = reinterpret_cast<long long*>(d_state + n_state);
    int* d_code = reinterpret_cast<int*>(d_cnt + 3 * (size_t)ns);
    while (c->chunk_ready.size() < ends.size()) {
        cudaEvent_t ev;
        HLM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        c->chunk_ready.push_back(ev);
    }
    const hlm::SolverParams& prm = c->params[c->uid];
    const int tpb = 256;
    // every upload is queued now, in chunk order, on its own stream
    for (size_t i = 0; i < ends.size(); ++i) {
        const long long lo = (i ? ends[i - 1] : 0) * 32, hi = std::min(ns, ends[i] * 32);
        const long long hi_pad = (i + 1 == ends.size()) ? c->ld : hi;  // the last chunk parks the padding lanes
        HLM_CUDA(cudaMemcpyAsync(d_state + (size_t)lo * n_eq, y0 + (size_t)lo * n_eq, sizeof(double) * (size_t)(hi - lo) * n_eq,
                                 cudaMemcpyHostToDevice, c->h2d_stream));
        init_state_kernel<<<(unsigned)((hi_pad - lo + tpb - 1) / tpb), tpb, 0, c->h2d_stream>>>(
            d_state, n_eq, ns, c->ld, c->y.p, c->t.p, c->h.p, c->next_q.p, c->reject_run.p, c->status.p, c->n_acc.p, c->n_rej.p,
            c->n_jump.p, c->t0, prm.initialStep, lo, hi_pad);
        HLM_CUDA(cudaGetLastError());
        ++c->launches;
        HLM_CUDA(cudaEventRecord(c->chunk_ready[i], c->h2d_stream));
    }
    for (size_t i = 0; i < ends.size(); ++i) {
        const long long tile = i ? ends[i - 1] : 0, end = ends[i];
        const long long lo = tile * 32, hi = std::min(ns, end * 32);
        const int buf = c->dense_cur ^ 1;
        if (c->copy_pending[buf]) {
            HLM_CUDA(cudaStreamWaitEvent(c->stream, c->ev_copy_done[buf], 0));
            c->copy_pending[buf] = false;
        }
        const size_t chunk_bytes = (size_t)(hi - lo) * (size_t)per_link;
        HLM_CUDA(c->dense[buf].reserve(doubles_for(chunk_bytes)));
        c->dense_cur = buf;
        HLM_CUDA(cudaStreamWaitEvent(c->stream, c->chunk_ready[i], 0));
        if (int r = queue_window(c, 0, nq, c->dense[buf].p, tile, end - tile, lo)) return r;
        export_results_kernel<<<(unsigned)((hi - lo + tpb - 1) / tpb), tpb, 0, c->stream>>>(
            c->y.p, c->status.p, c->n_acc.p, c->n_rej.p, c->n_jump.p, n_eq, ns, c->ld, out_final ? d_state : nullptr, d_cnt, d_code,
            lo, hi);
        HLM_CUDA(cudaGetLastError());
        ++c->launches;
        HLM_CUDA(cudaEventRecord(c->ev_kernel_done[buf], c->stream));
        HLM_CUDA(cudaStreamWaitEvent(c->copy_stream, c->ev_kernel_done[buf], 0));
        HLM_CUDA(cudaMemcpyAsync(static_cast<char*>(out_dense) + (size_t)lo * (size_t)per_link, c->dense[buf].p, chunk_bytes,
                                 cudaMemcpyDeviceToHost, c->copy_stream));
        const size_t n = (size_t)(hi - lo);
        if (out_final)
            HLM_CUDA(cudaMemcpyAsync(out_final + (size_t)lo * n_eq, d_state + (size_t)lo * n_eq, sizeof(double) * n * n_eq,
                                     cudaMemcpyDeviceToHost, c->copy_stream));
        long long* outs[3] = {out_acc, out_rej, out_jump};
        for (int k = 0; k < 3; ++k)
            if (outs[k])
                HLM_CUDA(cudaMemcpyAsync(outs[k] + lo, d_cnt + (size_t)k * ns + lo, sizeof(long long) * n, cudaMemcpyDeviceToHost,
                                         c->copy_stream));
        if (out_stiff) HLM_CUDA(cudaMemcpyAsync(out_stiff + lo, d_code + lo, sizeof(int) * n, cudaMemcpyDeviceToHost, c->copy_stream));
        HLM_CUDA(cudaEventRecord(c->ev_copy_done[buf], c->copy_stream));
        c->copy_pending[buf] = true;
    }
    c->q_done = nq;
    c->win_q_lo = 0;
    c->win_q_hi = nq;
    c->win_has_dense = false;  // the buffers hold chunks of links, not a window a caller could fetch
    HLM_CUDA(cudaStreamSynchronize(c->h2d_stream));
    HLM_CUDA(cudaStreamSynchronize(c->stream));
    HLM_CUDA(cudaStreamSynchronize(c->copy_stream));
    c->copy_pending[0] = c->copy_pending[1] = false;
    return HLM_OK;
}

int hlm_run_rk45(hlm_ctx* c, int uid, const double* y0, long long ns, double t0, double tf, const double* tq,
                 long long nq, double* out_final, void* out_dense, int* out_stiff, long long* out_acc,
                 long long* out_rej, long long* out_jump) {
    HLM_REQUIRE(c, "hlm_run_rk45: ctx is NULL");
    if (!out_dense) nq = 0;  // nothing to emit: one window straight to tf
    // a large output leaves in chunks of links (run_link_chunks), which also pipelines the state in and out
    bool chunked = false;
    if (const ModelInfo* m = find_model(uid)) {
        int ncol = m->n_eq, bytes = 8;
        hlm_output_layout(c, uid, &ncol, &bytes);
        const long long per_q = ns * ncol * (long long)bytes;
        const long long per_link = nq * ncol * (long long)bytes;
        chunked = nq > 0 && ns > 0 && per_q * nq > std::min<long long>(256LL << 20, c->dense_window_bytes) &&
                  per_link * 32 <= c->dense_window_bytes;
    }
    if (int r = solve_begin_impl(c, uid, y0, ns, t0, tf, tq, nq, chunked)) return r;
    if (chunked) return run_link_chunks(c, y0, out_dense, out_final, out_stiff, out_acc, out_rej, out_jump);
    if (nq == 0) {
        if (int r = hlm_solve_window(c, 0, 0)) return r;
    } else {
        // small output (one window), or a query list so long that even 32 links of it exceed a buffer:
        // windows over queries
        const long long per_q = ns * (long long)dense_record_bytes(c);
        const long long qw_max = std::max<long long>(1, c->dense_window_bytes / per_q);
        const long long n_win = (nq + qw_max - 1) / qw_max;
        const long long qw = (nq + n_win - 1) / n_win;  // even windows: the tail is not a sliver
        for (long long q = 0; q < nq;) {
            q = std::min(nq, q + qw);
            if (int r = hlm_solve_window(c, q, 1)) return r;
            if (int r = hlm_solve_fetch_window(c, out_dense)) return r;
        }
    }
    return hlm_solve_end(c, out_final, out_stiff, out_acc, out_rej, out_jump);
}

// ---- routed runs ---------------------------------------------------------------------------------

int hlm_route_set_topology(hlm_ctx* c, const long long* up_ptr, const int* up_idx, long long ns, const int* send_idx,
                           long long n_send) {
    HLM_REQUIRE(c && up_ptr && ns > 0 && n_send >= 0, "hlm_route_set_topology: bad argument");
    HLM_REQUIRE(up_ptr[0] == 0 && up_ptr[ns] >= 0 && (up_ptr[ns] == 0 || up_idx), "hlm_route_set_topology: bad CSR");
    HLM_REQUIRE(n_send == 0 || send_idx, "hlm_route_set_topology: send_idx is NULL");
    if (int r = use_device(c)) return r;
    const long long nnz = up_ptr[ns];
    const long long ld = (ns + 31) & ~31LL;
    for (long long i = 0; i < ns; ++i)
        if (up_ptr[i + 1] < up_ptr[i]) return fail(HLM_ERR_INVALID, "hlm_route_set_topology: up_ptr is not ascending");
    long long halo_need = 0;  // halo elements the topology refers to
    for (long long e = 0; e < nnz; ++e) {
        if (up_idx[e] >= ns) return fail(HLM_ERR_INVALID, "hlm_route_set_topology: upstream index out of range");
        if (up_idx[e] < 0) halo_need = std::max<long long>(halo_need, -((long long)up_idx[e] + 1) + 1);
    }
    std::vector<int> slot;
    try { slot.assign((size_t)ld, -1); } catch (const std::bad_alloc&) { return fail(HLM_ERR_NOMEM, "hlm_route_set_topology: out of host memory"); }
    for (long long k = 0; k < n_send; ++k) {
        if (send_idx[k] < 0 || send_idx[k] >= ns) return fail(HLM_ERR_INVALID, "hlm_route_set_topology: send index out of range");
        slot[(size_t)send_idx[k]] = (int)k;
    }
    HLM_CUDA(cudaStreamSynchronize(c->stream));  // kernels queued with the old topology
    HLM_CUDA(c->up_ptr.reserve((size_t)ns + 1));
    HLM_CUDA(c->up_idx.reserve((size_t)std::max<long long>(nnz, 1)));
    HLM_CUDA(c->send_idx.reserve((size_t)std::max<long long>(n_send, 1)));
    HLM_CUDA(c->send_slot.reserve((size_t)ld));
    HLM_CUDA(c->own_send.reserve((size_t)std::max<long long>(n_send, 1)));
    HLM_CUDA(c->qin.reserve((size_t)ld));
    HLM_CUDA(cudaMemcpyAsync(c->up_ptr.p, up_ptr, sizeof(long long) * (size_t)(ns + 1), cudaMemcpyHostToDevice, c->stream));
    if (nnz > 0) HLM_CUDA(cudaMemcpyAsync(c->up_idx.p, up_idx, sizeof(int) * (size_t)nnz, cudaMemcpyHostToDevice, c->stream));
    if (n_send > 0) HLM_CUDA(cudaMemcpyAsync(c->send_idx.p, send_idx, sizeof(int) * (size_t)n_send, cudaMemcpyHostToDevice, c->stream));
    HLM_CUDA(cudaMemcpyAsync(c->send_slot.p, slot.data(), sizeof(int) * (size_t)ld, cudaMemcpyHostToDevice, c->stream));
    HLM_CUDA(cudaMemsetAsync(c->qin.p, 0, sizeof(double) * (size_t)ld, c->stream));
    HLM_CUDA(cudaMemsetAsync(c->own_send.p, 0, sizeof(double) * (size_t)std::max<long long>(n_send, 1), c->stream));
    HLM_CUDA(cudaStreamSynchronize(c->stream));  // the host arrays (and `slot`) may go away after return
    c->routed = true;
    c->route_ns = ns;
    c->route_nnz = nnz;
    c->route_n_send = n_send;
    c->route_halo_need = halo_need;
    c->send_buf = c->own_send.p;
    return HLM_OK;
}

int hlm_route_clear(hlm_ctx* c) {
    HLM_REQUIRE(c, "hlm_route_clear: ctx is NULL");
    c->routed = false;
    c->route_ns = c->route_nnz = c->route_n_send = 0;
    c->send_buf = nullptr;
    return HLM_OK;
}

int hlm_route_set_send_buffer(hlm_ctx* c, double* dev_buf) {
    HLM_REQUIRE(c, "hlm_route_set_send_buffer: ctx is NULL");
    if (!c->routed) return fail(HLM_ERR_STATE, "hlm_route_set_send_buffer: no topology");
    c->send_buf = dev_buf ? dev_buf : c->own_send.p;
    return HLM_OK;
}

int hlm_route_send_buffer(hlm_ctx* c, void** dev_ptr, long long* n_send) {
    HLM_REQUIRE(c, "hlm_route_send_buffer: ctx is NULL");
    if (!c->routed) return fail(HLM_ERR_STATE, "hlm_route_send_buffer: no topology");
    if (dev_ptr) *dev_ptr = c->send_buf;
    if (n_send) *n_send = c->route_n_send;
    return HLM_OK;
}

int hlm_route_pack(hlm_ctx* c) {
    HLM_REQUIRE(c, "hlm_route_pack: ctx is NULL");
    if (!c->routed || !c->in_session) return fail(HLM_ERR_STATE, "hlm_route_pack: needs a topology and a session");
    if (int r = use_device(c)) return r;
    if (c->route_n_send == 0) {
        c->route_epoch = 0;
        return HLM_OK;
    }
    const int tpb = 256;
    c->route_epoch = 0;  // the initial state is what the first gather reads, at parity 0
    const bool peer = c->peer_world > 1;
    route_pack_kernel<<<(unsigned)((c->route_n_send + tpb - 1) / tpb), tpb, 0, c->stream>>>(
        c->y.p, c->send_idx.p, c->route_n_send, c->send_buf, peer ? c->peer_ptrs.p : nullptr, c->peer_world,
        (long long)c->peer_rank * c->peer_max_send);
    HLM_CUDA(cudaGetLastError());
    ++c->launches;
    return HLM_OK;
}

int hlm_route_gather(hlm_ctx* c, const double* dev_halo) {
    HLM_REQUIRE(c, "hlm_route_gather: ctx is NULL");
    if (!c->routed || !c->in_session) return fail(HLM_ERR_STATE, "hlm_route_gather: needs a topology and a session");
    if (c->route_ns != c->ns) return fail(HLM_ERR_STATE, "hlm_route_gather: topology was set for another link count");
    if (c->peer_world > 1) {  // the halo is this rank's own vector, filled by the peers' kernels
        if (c->route_halo_need > (long long)c->peer_world * c->peer_max_send)
            return fail(HLM_ERR_INVALID, "hlm_route_gather: the topology refers to halo elements beyond the peer halo vector");
        dev_halo = c->own_halo + (c->route_epoch % 2) * (long long)c->peer_world * c->peer_max_send;
        ++c->route_epoch;
    }
    if (c->route_halo_need > 0 && !dev_halo)
        return fail(HLM_ERR_INVALID, "hlm_route_gather: the topology refers to " + std::to_string(c->route_halo_need) +
                                         " halo elements but no halo vector was given");
    if (int r = use_device(c)) return r;
    const int tpb = 256;
    route_gather_kernel<<<(unsigned)((c->ns + tpb - 1) / tpb), tpb, 0, c->stream>>>(c->y.p, dev_halo, c->up_ptr.p,
                                                                                   c->up_idx.p, c->ns, c->qin.p);
    HLM_CUDA(cudaGetLastError());
    ++c->launches;
    return HLM_OK;
}

int hlm_route_peek(hlm_ctx* c, double* out_qin, double* out_send) {
    HLM_REQUIRE(c, "hlm_route_peek: ctx is NULL");
    if (!c->routed) return fail(HLM_ERR_STATE, "hlm_route_peek: no topology");
    if (int r = use_device(c)) return r;
    if (out_qin) HLM_CUDA(cudaMemcpyAsync(out_qin, c->qin.p, sizeof(double) * (size_t)c->route_ns, cudaMemcpyDeviceToHost, c->stream));
    if (out_send && c->route_n_send > 0)
        HLM_CUDA(cudaMemcpyAsync(out_send, c->send_buf, sizeof(double) * (size_t)c->route_n_send, cudaMemcpyDeviceToHost, c->stream));
    HLM_CUDA(cudaStreamSynchronize(c->stream));
    return HLM_OK;
}

// ---- peer-memory exchange for routed runs ----------------------------------------------------------

static int peer_close_impl(hlm_ctx* c) {
    for (void* p : c->peer_opened)
        if (p) cudaIpcCloseMemHandle(p);
    c->peer_opened.clear();
    if (c->own_halo) cudaFree(c->own_halo);
    c->own_halo = nullptr;
    c->peer_world = 0;
    c->peer_rank = 0;
    c->peer_max_send = 0;
    return HLM_OK;
}

int hlm_route_peer_alloc(hlm_ctx* c, int world, int rank, long long max_send, void* ipc_handle_out) {
    HLM_REQUIRE(c && ipc_handle_out && world > 1 && rank >= 0 && rank < world && max_send > 0, "hlm_route_peer_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI documents a 64-byte handle");
    if (int r = use_device(c)) return r;
    HLM_CUDA(cudaStreamSynchronize(c->stream));
    peer_close_impl(c);
    const size_t n = 2 * (size_t)world * (size_t)max_send;
    HLM_CUDA(cudaMalloc(&c->own_halo, n * sizeof(double)));
    HLM_CUDA(cudaMemset(c->own_halo, 0, n * sizeof(double)));
    cudaIpcMemHandle_t h;
    HLM_CUDA(cudaIpcGetMemHandle(&h, c->own_halo));
    std::memcpy(ipc_handle_out, &h, sizeof(h));
    c->peer_world = -world;  // allocated, peers not mapped yet (kernels keep using the send buffer)
    c->peer_rank = rank;
    c->peer_max_send = max_send;
    return HLM_OK;
}

int hlm_route_peer_open(hlm_ctx* c, const void* handles) {
    HLM_REQUIRE(c && handles, "hlm_route_peer_open: NULL argument");
    if (c->peer_world >= 0 || !c->own_halo) return fail(HLM_ERR_STATE, "hlm_route_peer_open: call hlm_route_peer_alloc first");
    if (int r = use_device(c)) return r;
    const int world = -c->peer_world;
    std::vector<double*> ptrs((size_t)world, nullptr);
    c->peer_opened.assign((size_t)world, nullptr);
    for (int r = 0; r < world; ++r) {
        if (r == c->peer_rank) {
            ptrs[(size_t)r] = c->own_halo;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, static_cast<const char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        HLM_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        c->peer_opened[(size_t)r] = p;
        ptrs[(size_t)r] = static_cast<double*>(p);
    }
    HLM_CUDA(c->peer_ptrs.reserve((size_t)world));
    HLM_CUDA(cudaMemcpy(c->peer_ptrs.p, ptrs.data(), sizeof(double*) * (size_t)world, cudaMemcpyHostToDevice));
    c->peer_world = world;
    c->route_epoch = 0;
    return HLM_OK;
}

int hlm_route_peer_close(hlm_ctx* c) {
    HLM_REQUIRE(c, "hlm_route_peer_close: ctx is NULL");
    if (int r = use_device(c)) return r;
    HLM_CUDA(cudaStreamSynchronize(c->stream));
    return peer_close_impl(c);
}

int hlm_debug_eval(hlm_ctx* c, int op, const double* x, const double* y, double* out, long long n) {
    HLM_REQUIRE(c && x && y && out && n > 0 && op >= 0 && op <= 6, "hlm_debug_eval: bad argument");
    if (int r = use_device(c)) return r;
    double *dx = nullptr, *dy = nullptr, *dout = nullptr;
    HLM_CUDA(cudaMalloc(&dx, sizeof(double) * n));
    HLM_CUDA(cudaMalloc(&dy, sizeof(double) * n));
    HLM_CUDA(cudaMalloc(&dout, sizeof(double) * n));
    HLM_CUDA(cudaMemcpyAsync(dx, x, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    HLM_CUDA(cudaMemcpyAsync(dy, y, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(op, dx, dy, dout, n);
    HLM_CUDA(cudaGetLastError());
    ++c->launches;
    HLM_CUDA(cudaMemcpyAsync(out, dout, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    HLM_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(dx); cudaFree(dy); cudaFree(dout);
    return HLM_OK;
}

int hlm_measure_fma_peak(hlm_ctx* c, int bits, double* tflops) {
    HLM_REQUIRE(c && tflops && (bits == 64 || bits == 32), "hlm_measure_fma_peak: bad argument");
    if (int r = use_device(c)) return r;
    const int tpb = 256, blocks = c->sm_count * 8, iters = 4096;
    void* out = nullptr;
    HLM_CUDA(cudaMalloc(&out, (size_t)tpb * blocks * 8));
    cudaEvent_t e0, e1;
    HLM_CUDA(cudaEventCreate(&e0));
    HLM_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        HLM_CUDA(cudaEventRecord(e0, c->stream));
        if (bits == 64) fma_peak_kernel<double><<<blocks, tpb, 0, c->stream>>>((double*)out, iters, 1.0);
        else fma_peak_kernel<float><<<blocks, tpb, 0, c->stream>>>((float*)out, iters, 1.0f);
        HLM_CUDA(cudaEventRecord(e1, c->stream));
        HLM_CUDA(cudaStreamSynchronize(c->stream));
        float ms = 0;
        HLM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
        c->launches++;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    const double flops = 2.0 * 64.0 * (double)iters * tpb * blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    return HLM_OK;
}

}  // extern "C"
