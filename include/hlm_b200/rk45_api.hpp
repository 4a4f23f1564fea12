// hlm_b200/rk45_api.hpp — header-only C++ shims that rebuild the reference's operator surface on the
// C ABI of hlm_b200.h, so reference host code keeps compiling against the same names:
//
//   SpatialParams                                   I_O/parameters_loader.hpp:19-37
//   struct Model204 { UID, N_EQ, SP_TYPE, Parameters }   models/model_204.hpp:15-30
//   rk45_api::setModelParameters<Model>(p)          model_registry.hpp:9-13, model_registry.cpp:18-60
//   rk45_api::run_rk45<Model>(h_y0, t0, tf, h_query_times, d_sp)   solver/rk45_api.hpp:273-313
//   rk45_api::setup_gpu_buffers / launch_rk45_kernel / retrieve_and_free   solver/rk45_api.hpp:63-270
//   rk45_api::FinalType / DenseType                 solver/rk45_api.hpp:55-56
//
// What changes for a caller:
//   * `rhs` is not part of the trait here: a __device__ function cannot cross a C ABI, so models
//     are compiled into libhlm_b200.so and selected by Model::UID;
//   * `d_sp` may be the reference's device pointer (cudaMalloc'ed AoS array, main.cpp:392-404) or a host
//     pointer — the library tells which and transposes the records on the device; its length is the
//     number of systems in h_y0, as in the reference.  hlm_b200::SpView (pointer + length) also works;
//   * FinalType / DenseType are std::vector<double, ...> whose allocator hands out PAGE-LOCKED memory from a
//     pool and does not zero-fill: results arrive by DMA at the PCIe rate and a 10 GB dense array is neither
//     memset nor staged (std::vector<double> would cost both).  Element access, iteration, size(), data()
//     and structured bindings are unchanged; only code that spells the type std::vector<double> needs
//     `rk45_api::DenseType` (the reference's own alias) instead;
//   * the device pointers of setup_gpu_buffers' tuple are opaque handles of the resident session (the
//     buffers are owned by the library and re-used across calls), not addresses a caller's own kernel can use;
//   * forcings are handed over with rk45_api::setForcing(...) instead of cudaMemcpyToSymbol on
//     c_forc_dt / c_forc_nT / d_forc_data (main.cpp:552-574);
//   * errors still surface as std::runtime_error (solver/rk45_api.hpp:87-108).
// The process-global context mirrors the reference's process-global __constant__ state: one model
// parameter set and one forcing set per process and device.  Multi-GPU hosts create one
// hlm_b200::Context per device (hlm_b200::Context ctx(dev)) and call the member functions.
#pragma once

#include <cstddef>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../hlm_b200.h"

// ---- I_O/parameters_loader.hpp:19-37 --------------------------------------------------------------
struct SpatialParams {
    long stream;
    long next_stream;
    double c1, infil, perco, Hu, lat, sw, ss, n_mann, slope, L, A_h, alpha3, alpha4, melt_f, temp_thr;
};
static_assert(sizeof(SpatialParams) == 136, "SpatialParams must keep the reference's 136-byte layout");

namespace hlm_b200 {

struct Parameters {  // models/model_204.hpp:22-30
    double initialStep = 0.01;
    double rtol = 1e-6;
    double atol = 1e-9;
    double safety = 0.9;
    double minScale = 0.2;
    double maxScale = 10.0;
};

/// Host view of a per-link parameter array (what the reference passes as a device pointer).
struct SpView {
    const SpatialParams* ptr = nullptr;
    long long n = 0;
};

inline void check(int rc, const char* what) {
    if (rc != HLM_OK) throw std::runtime_error(std::string(what) + ": " + hlm_last_error());
}

// ---- result buffers: page-locked, pooled, not zero-filled -------------------------------------------
// The reference returns std::vector<double> (pageable: the device-to-host copy is staged through the driver's
// bounce buffers at a fraction of the PCIe rate) after value-initialising it (a 9.6 GB memset for one day of
// hourly output of 10 M links).  Page-locking is expensive too, so freed blocks go back to a pool and the next
// call of the same shape (a run driven interval by interval) reuses them.
namespace detail {
class PinnedPool {
  public:
    static PinnedPool& instance() {
        static PinnedPool* p = new PinnedPool();  // never destroyed: blocks may outlive static destruction order
        return *p;
    }
    void* take(size_t bytes) {
        {
            std::lock_guard<std::mutex> g(m_);
            auto it = free_.find(bytes);
            if (it != free_.end() && !it->second.empty()) {
                void* p = it->second.back();
                it->second.pop_back();
                cached_ -= bytes;
                return p;
            }
        }
        void* p = nullptr;
        if (hlm_host_alloc(&p, (long long)bytes) != HLM_OK || !p) {
            release();  // cached blocks of other sizes may be what is in the way
            if (hlm_host_alloc(&p, (long long)bytes) != HLM_OK || !p) throw std::bad_alloc();
        }
        return p;
    }
    void give(void* p, size_t bytes) {
        std::lock_guard<std::mutex> g(m_);
        free_[bytes].push_back(p);
        cached_ += bytes;
    }
    /// return every cached block to the system
    void release() {
        std::lock_guard<std::mutex> g(m_);
        for (auto& kv : free_)
            for (void* p : kv.second) hlm_host_free(p);
        free_.clear();
        cached_ = 0;
    }
    size_t cached_bytes() const { return cached_; }

  private:
    std::mutex m_;
    std::map<size_t, std::vector<void*>> free_;
    size_t cached_ = 0;
};
}  // namespace detail

template <typename T> struct PinnedAllocator {
    using value_type = T;
    PinnedAllocator() = default;
    template <typename U> PinnedAllocator(const PinnedAllocator<U>&) {}
    T* allocate(size_t n) { return static_cast<T*>(detail::PinnedPool::instance().take(n > 0 ? n * sizeof(T) : 1)); }
    void deallocate(T* p, size_t n) { detail::PinnedPool::instance().give(p, n > 0 ? n * sizeof(T) : 1); }
    // default-initialisation instead of value-initialisation: resize(n) does not touch the n elements
    template <typename U> void construct(U* p) { ::new (static_cast<void*>(p)) U; }
    template <typename U, typename... A> void construct(U* p, A&&... a) { ::new (static_cast<void*>(p)) U(std::forward<A>(a)...); }
    template <typename U> bool operator==(const PinnedAllocator<U>&) const { return true; }
    template <typename U> bool operator!=(const PinnedAllocator<U>&) const { return false; }
};
template <typename T> using HostVector = std::vector<T, PinnedAllocator<T>>;
/// give the pooled page-locked blocks back (e.g. before a phase that needs the host memory)
inline void release_pinned_pool() { detail::PinnedPool::instance().release(); }

/// RAII owner of one hlm_ctx (one CUDA device).
class Context {
  public:
    explicit Context(int device = 0) { check(hlm_create(device, &ctx_), "hlm_create"); }
    ~Context() { hlm_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    hlm_ctx* get() const { return ctx_; }

    void setModelParameters(int uid, const Parameters& p) {
        const double v[6] = {p.initialStep, p.rtol, p.atol, p.safety, p.minScale, p.maxScale};
        check(hlm_set_model_parameters(ctx_, uid, v), "hlm_set_model_parameters");
    }
    void setSpatialParams(const SpatialParams* sp, long long n) {
        check(hlm_upload_spatial_params(ctx_, sp, n, (long long)sizeof(SpatialParams)), "hlm_upload_spatial_params");
    }
    /// forcing j as [nT][ncols] floats sampled every dt_hours; col == nullptr: column c serves link c
    void setForcing(int j, double dt_hours, long long nT, long long ncols, const float* data) {
        check(hlm_upload_forcing(ctx_, j, dt_hours, nT, ncols, data), "hlm_upload_forcing");
    }
    void setForcingColumns(const int* col, long long n) {
        check(hlm_set_forcing_columns(ctx_, col, n), "hlm_set_forcing_columns");
    }
    void clearForcings() { check(hlm_clear_forcings(ctx_), "hlm_clear_forcings"); }
    /// Links the RK45 path flags stiff are carried on by the Radau IIA fallback (run_rk45's documented
    /// behaviour, solver/rk45_api.hpp:198-247) instead of being abandoned; they come back as HLM_LINK_STIFF_SOLVED.
    void setStiffFallback(bool on) { check(hlm_set_stiff_fallback(ctx_, on ? 1 : 0), "hlm_set_stiff_fallback"); }
    /// HLM_SCHEDULE_AUTO / _TILES / _LANES / _SORTED_TILES: how links are dealt to lanes (bit-identical results either way).
    void setSchedule(int mode) { check(hlm_set_schedule(ctx_, mode), "hlm_set_schedule"); }

    struct Result {
        HostVector<double> final_state;  // [ns][N_EQ]
        HostVector<double> dense;        // [ns][nq][N_EQ]
        std::vector<int> stiff;           // HLM_LINK_* per link
        std::vector<long long> n_accept, n_reject, n_jump;
    };
    /// y0 [ns][n_eq] and tq [nq] in host memory; with_counters = false skips the per-link codes and counters
    /// (what the reference's run_rk45 returns is final + dense only)
    Result run(int uid, int n_eq, const double* y0, long long ns, double t0, double tf, const double* tq, long long nq,
               bool with_counters = true) {
        Result r;
        r.final_state.resize((size_t)ns * n_eq);  // page-locked, not zero-filled: every element is written by the copy
        r.dense.resize((size_t)ns * nq * n_eq);
        if (with_counters) {
            r.stiff.resize((size_t)ns);
            r.n_accept.resize((size_t)ns);
            r.n_reject.resize((size_t)ns);
            r.n_jump.resize((size_t)ns);
        }
        check(hlm_run_rk45(ctx_, uid, y0, ns, t0, tf, tq, nq, r.final_state.data(), nq ? r.dense.data() : nullptr,
                           with_counters ? r.stiff.data() : nullptr, with_counters ? r.n_accept.data() : nullptr,
                           with_counters ? r.n_reject.data() : nullptr, with_counters ? r.n_jump.data() : nullptr),
              "hlm_run_rk45");
        return r;
    }
    Result run(int uid, int n_eq, const std::vector<double>& y0, double t0, double tf, const std::vector<double>& tq) {
        return run(uid, n_eq, y0.data(), (long long)(y0.size() / (size_t)n_eq), t0, tf, tq.data(), (long long)tq.size());
    }

  private:
    hlm_ctx* ctx_ = nullptr;
};

inline Context& default_context() {
    static Context ctx(0);
    return ctx;
}

}  // namespace hlm_b200

// ---- the Model trait, minus the device rhs ----------------------------------------------------------
struct Model204 {
    using SP_TYPE = SpatialParams;
    static constexpr unsigned short UID = 204;  // models/model_204.hpp:18
    static constexpr int N_EQ = 5;
    using Parameters = hlm_b200::Parameters;
};
struct Model200 {  // named by the reference (README.md:95) but not defined there: project-defined, see hlm_b200.h
    using SP_TYPE = SpatialParams;
    static constexpr unsigned short UID = 200;
    static constexpr int N_EQ = 5;
    using Parameters = hlm_b200::Parameters;
};
struct DummyModel {  // README.md:24,41-42 name it; defined by model_dummy_python.ipynb:65-89
    using SP_TYPE = SpatialParams;
    static constexpr unsigned short UID = 0;
    static constexpr int N_EQ = 5;
    using Parameters = hlm_b200::Parameters;
};

namespace rk45_api {

// solver/rk45_api.hpp:55-56 (there: std::vector<double>; here page-locked and not zero-filled, see the head of this file)
using DenseType = hlm_b200::HostVector<double>;
using FinalType = hlm_b200::HostVector<double>;

/// rk45_api::setModelParameters<Model>(p) — model_registry.hpp:9-13.
template <typename Model> void setModelParameters(const typename Model::Parameters& p) {
    hlm_b200::default_context().setModelParameters(Model::UID, p);
}

/// Replaces the forcing upload of main.cpp:552-574.
inline void setForcing(int j, double dt_hours, long long nT, long long ncols, const float* data) {
    hlm_b200::default_context().setForcing(j, dt_hours, nT, ncols, data);
}
inline void setForcingColumns(const int* col, long long n) { hlm_b200::default_context().setForcingColumns(col, n); }

/// rk45_api::run_rk45<Model>(h_y0, t0, tf, h_query_times, d_sp) — solver/rk45_api.hpp:273-313, the reference's
/// signature: d_sp is the AoS parameter array of the num_systems = h_y0.size() / N_EQ links, in device memory as
/// the reference passes it (main.cpp:392-404) or in host memory; nullptr keeps the parameters uploaded last.
/// Returns {final [sys][N_EQ], dense [sys][q][N_EQ]} like the reference.  VecY / VecQ: any contiguous container of
/// double (std::vector<double> as in the reference, or the FinalType of a previous call, so that a run chained
/// interval by interval uploads from page-locked memory).
template <class Model, class VecY, class VecQ>
std::pair<FinalType, DenseType> run_rk45(const VecY& h_y0, double t0, double tf, const VecQ& h_query_times,
                                         const typename Model::SP_TYPE* d_sp) {
    auto& ctx = hlm_b200::default_context();
    const long long ns = (long long)(h_y0.size() / (size_t)Model::N_EQ);
    if (d_sp) ctx.setSpatialParams(d_sp, ns);
    auto r = ctx.run(Model::UID, Model::N_EQ, h_y0.data(), ns, t0, tf, h_query_times.data(), (long long)h_query_times.size(),
                     /*with_counters=*/false);
    return {std::move(r.final_state), std::move(r.dense)};
}
/// The same with a host view (pointer + length), or with the parameters uploaded beforehand (sp = {}).
template <class Model, class VecY, class VecQ>
std::pair<FinalType, DenseType> run_rk45(const VecY& h_y0, double t0, double tf, const VecQ& h_query_times,
                                         hlm_b200::SpView sp = {}) {
    auto& ctx = hlm_b200::default_context();
    if (sp.ptr) ctx.setSpatialParams(sp.ptr, sp.n);
    auto r = ctx.run(Model::UID, Model::N_EQ, h_y0.data(), (long long)(h_y0.size() / (size_t)Model::N_EQ), t0, tf,
                     h_query_times.data(), (long long)h_query_times.size(), /*with_counters=*/false);
    return {std::move(r.final_state), std::move(r.dense)};
}

// ---- the three sub-steps run_rk45 is documented to compose (solver/rk45_api.hpp:63-270) ---------------------
// Same names, tuple and argument shapes as the reference, over the resident session of the C ABI:
//   setup_gpu_buffers   = hlm_solve_begin (state + query times to the device; the reference: 5 cudaMalloc + 2 H2D)
//   launch_rk45_kernel  = hlm_solve_restart(t0, tf) + hlm_solve_window + synchronize (the reference: <<<>>> + sync)
//   retrieve_and_free   = hlm_solve_fetch_window + hlm_solve_end (the reference: 3 D2H, Radau on flagged links,
//                         4 cudaFree, host reorder; here the fallback ran with the window if hlm_set_stiff_fallback
//                         is on, and nothing is reordered or freed)
// The five pointers of the tuple are opaque handles of that session: they identify it in the two later calls
// and must not be dereferenced or handed to other kernels.  One session per process at a time.
namespace detail {
struct PendingRun {
    int uid = -1, n_eq = 0, ns = 0, nq = 0;
    bool launched = false;
    std::vector<double> tq;
    double tokens[4] = {0, 0, 0, 0};  // addresses of these serve as the handles
    int stiff_token = 0;
};
inline std::unique_ptr<PendingRun>& pending() {
    static std::unique_ptr<PendingRun> p;
    return p;
}
inline PendingRun& pending_for(const double* d_y0_all, const char* who) {
    auto& p = pending();
    if (!p || d_y0_all != &p->tokens[0]) throw std::runtime_error(std::string(who) + ": these are not the buffers of the last setup_gpu_buffers");
    return *p;
}
}  // namespace detail

template <class Model>
std::tuple<double*, double*, double*, double*, int*, int, int> setup_gpu_buffers(const std::vector<double>& h_y0,
                                                                                 const std::vector<double>& h_query_times) {
    auto& ctx = hlm_b200::default_context();
    auto run = std::make_unique<detail::PendingRun>();
    run->uid = Model::UID;
    run->n_eq = Model::N_EQ;
    run->ns = int(h_y0.size() / Model::N_EQ);
    run->nq = int(h_query_times.size());
    run->tq = h_query_times;
    // the interval is not known yet (it arrives with launch_rk45_kernel): begin on an empty one
    hlm_b200::check(hlm_solve_begin(ctx.get(), Model::UID, h_y0.data(), run->ns, 0.0, 0.0, run->tq.data(), run->nq), "hlm_solve_begin");
    hlm_b200::check(hlm_synchronize(ctx.get()), "hlm_synchronize");  // h_y0 may go away, as after the reference's cudaMemcpy
    detail::PendingRun* r = run.get();
    detail::pending() = std::move(run);
    return std::make_tuple(&r->tokens[0], &r->tokens[1], &r->tokens[2], &r->tokens[3], &r->stiff_token, r->ns, r->nq);
}

template <class Model>
void launch_rk45_kernel(double* d_y0_all, double* /*d_y_final_all*/, double* /*d_query_times*/, double* /*d_dense_all*/,
                        int* /*d_stiff*/, int num_systems, int num_queries, double t0, double tf,
                        const typename Model::SP_TYPE* d_sp) {
    auto& ctx = hlm_b200::default_context();
    detail::PendingRun& r = detail::pending_for(d_y0_all, "launch_rk45_kernel");
    if (num_systems != r.ns || num_queries != r.nq || Model::UID != r.uid)
        throw std::runtime_error("launch_rk45_kernel: sizes differ from setup_gpu_buffers");
    if (d_sp) ctx.setSpatialParams(d_sp, r.ns);
    hlm_b200::check(hlm_solve_restart(ctx.get(), t0, tf, r.tq.data(), r.nq), "hlm_solve_restart");
    hlm_b200::check(hlm_solve_window(ctx.get(), r.nq, r.nq > 0 ? 1 : 0), "hlm_solve_window");
    hlm_b200::check(hlm_synchronize(ctx.get()), "Kernel execution failed");  // solver/rk45_api.hpp:150-153
    r.launched = true;
}

template <class Model>
std::pair<FinalType, DenseType> retrieve_and_free(double* d_y0_all, double* /*d_y_final_all*/, double* /*d_query_times*/,
                                                  double* /*d_dense_all*/, int* /*d_stiff*/, int num_systems, int num_queries,
                                                  double /*t0*/, double /*tf*/, const typename Model::SP_TYPE* /*d_sp*/) {
    auto& ctx = hlm_b200::default_context();
    detail::PendingRun& r = detail::pending_for(d_y0_all, "retrieve_and_free");
    if (!r.launched) throw std::runtime_error("retrieve_and_free: launch_rk45_kernel has not run");
    if (num_systems != r.ns || num_queries != r.nq) throw std::runtime_error("retrieve_and_free: sizes differ from setup_gpu_buffers");
    FinalType fin;
    DenseType dense;
    fin.resize((size_t)r.ns * r.n_eq);
    dense.resize((size_t)r.ns * r.nq * r.n_eq);
    if (r.nq > 0) hlm_b200::check(hlm_solve_fetch_window(ctx.get(), dense.data()), "hlm_solve_fetch_window");
    hlm_b200::check(hlm_solve_end(ctx.get(), fin.data(), nullptr, nullptr, nullptr, nullptr), "hlm_solve_end");
    detail::pending().reset();
    return {std::move(fin), std::move(dense)};
}

}  // namespace rk45_api
