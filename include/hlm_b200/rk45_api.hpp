// hlm_b200/rk45_api.hpp — header-only C++ shims that rebuild the reference's operator surface on the
// C ABI of hlm_b200.h, so reference host code keeps compiling against the same names:
//
//   SpatialParams                                   I_O/parameters_loader.hpp:19-37
//   struct Model204 { UID, N_EQ, SP_TYPE, Parameters }   models/model_204.hpp:15-30
//   rk45_api::setModelParameters<Model>(p)          model_registry.hpp:9-13, model_registry.cpp:18-60
//   rk45_api::run_rk45<Model>(h_y0, t0, tf, h_query_times, d_sp)   solver/rk45_api.hpp:273-313
//   rk45_api::FinalType / DenseType                 solver/rk45_api.hpp:55-56
//
// What changes for a caller:
//   * `rhs` is not part of the trait here: a __device__ function cannot cross a C ABI, so models
//     are compiled into libhlm_b200.so and selected by Model::UID;
//   * `d_sp` is a HOST pointer to the SpatialParams array plus its length (hlm_b200::SpView); the
//     library uploads and transposes it (the reference cudaMallocs it in main.cpp:392-404);
//   * forcings are handed over with rk45_api::setForcing(...) instead of cudaMemcpyToSymbol on
//     c_forc_dt / c_forc_nT / d_forc_data (main.cpp:552-574);
//   * errors still surface as std::runtime_error (solver/rk45_api.hpp:87-108).
// The process-global context mirrors the reference's process-global __constant__ state: one model
// parameter set and one forcing set per process and device.  Multi-GPU hosts create one
// hlm_b200::Context per device (hlm_b200::Context ctx(dev)) and call the member functions.
#pragma once

#include <cstddef>
#include <stdexcept>
#include <string>
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
    /// HLM_SCHEDULE_AUTO / _TILES / _LANES: how links are dealt to lanes (bit-identical results either way).
    void setSchedule(int mode) { check(hlm_set_schedule(ctx_, mode), "hlm_set_schedule"); }

    struct Result {
        std::vector<double> final_state;  // [ns][N_EQ]
        std::vector<double> dense;        // [ns][nq][N_EQ]
        std::vector<int> stiff;           // HLM_LINK_* per link
        std::vector<long long> n_accept, n_reject, n_jump;
    };
    Result run(int uid, int n_eq, const std::vector<double>& y0, double t0, double tf,
               const std::vector<double>& tq) {
        Result r;
        const long long ns = (long long)(y0.size() / (size_t)n_eq), nq = (long long)tq.size();
        r.final_state.assign((size_t)ns * n_eq, 0.0);
        r.dense.assign((size_t)ns * nq * n_eq, 0.0);
        r.stiff.assign((size_t)ns, 0);
        r.n_accept.assign((size_t)ns, 0);
        r.n_reject.assign((size_t)ns, 0);
        r.n_jump.assign((size_t)ns, 0);
        check(hlm_run_rk45(ctx_, uid, y0.data(), ns, t0, tf, tq.data(), nq, r.final_state.data(),
                           nq ? r.dense.data() : nullptr, r.stiff.data(), r.n_accept.data(), r.n_reject.data(),
                           r.n_jump.data()),
              "hlm_run_rk45");
        return r;
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

using DenseType = std::vector<double>;
using FinalType = std::vector<double>;

/// rk45_api::setModelParameters<Model>(p) — model_registry.hpp:9-13.
template <typename Model> void setModelParameters(const typename Model::Parameters& p) {
    hlm_b200::default_context().setModelParameters(Model::UID, p);
}

/// Replaces the forcing upload of main.cpp:552-574.
inline void setForcing(int j, double dt_hours, long long nT, long long ncols, const float* data) {
    hlm_b200::default_context().setForcing(j, dt_hours, nT, ncols, data);
}
inline void setForcingColumns(const int* col, long long n) { hlm_b200::default_context().setForcingColumns(col, n); }

/// rk45_api::run_rk45<Model>(h_y0, t0, tf, h_query_times, d_sp) — solver/rk45_api.hpp:273-313.
/// Returns {final [sys][N_EQ], dense [sys][q][N_EQ]} like the reference.  `sp` is a host view.
template <class Model>
std::pair<FinalType, DenseType> run_rk45(const std::vector<double>& h_y0, double t0, double tf,
                                         const std::vector<double>& h_query_times, hlm_b200::SpView sp = {}) {
    auto& ctx = hlm_b200::default_context();
    if (sp.ptr) ctx.setSpatialParams(sp.ptr, sp.n);
    auto r = ctx.run(Model::UID, Model::N_EQ, h_y0, t0, tf, h_query_times);
    return {std::move(r.final_state), std::move(r.dense)};
}

}  // namespace rk45_api
