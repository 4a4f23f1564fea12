// hlm_hostio_capi.cpp — C entry points over the host-side I/O headers (hlm_netcdf.hpp, hlm_config.hpp,
// hlm_host.hpp) so that the CPU test-suite can drive them through ctypes.  Not part of the solver's
// boundary (that is include/hlm_b200.h); no CUDA, no dependency on libhlm_b200.so.
#include <cstring>
#include <sstream>
#include <string>

#include "hlm_config.hpp"
#include "hlm_netcdf.hpp"
#include "hlm_routing.hpp"

namespace {
thread_local std::string g_err, g_text;
template <typename F> int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
std::string json_escape(const std::string& s) {
    std::string o;
    for (char c : s) {
        if (c == '"' || c == '\\') { o += '\\'; o += c; }
        else if (c == '\n') o += "\\n";
        else o += c;
    }
    return o;
}
}  // namespace

extern "C" {

const char* hlmio_last_error(void) { return g_err.c_str(); }

/// shape of a variable: returns its rank (<= 8), fills shape[]; -1 on error
int hlmio_inquire(const char* path, const char* var, long long* shape, int* elem_size) {
    int rank = -1;
    const int rc = guarded([&] {
        auto r = hlmnc::open_reader(path);
        const hlmnc::VarInfo& v = r->inquire(var);
        if (v.shape.size() > 8) throw std::runtime_error("rank > 8");
        for (size_t k = 0; k < v.shape.size(); ++k) shape[k] = (long long)v.shape[k];
        if (elem_size) *elem_size = v.elem.size;
        rank = (int)v.shape.size();
    });
    return rc ? -1 : rank;
}
/// the solver intervals hlm_run cuts [t_begin, t_end] into: fills bounds[0..n] (at most cap), returns n or -1
long long hlmio_interval_boundaries(double t_begin, double t_end, const char* interval, double* bounds, long long cap) {
    long long n = 0;
    const int rc = guarded([&] {
        const double dt = parse_interval_minutes(interval);
        long long k = 0;
        for (double ta = t_begin; ta < t_end;) {
            const double tb = next_interval_boundary(ta, t_end, dt, k);
            if (n < cap) bounds[n] = tb;
            ++n;
            ta = tb;
        }
    });
    return rc ? -1 : n;
}
/// newline-separated variable names
const char* hlmio_variables(const char* path) {
    g_text.clear();
    if (guarded([&] {
            auto r = hlmnc::open_reader(path);
            for (auto& n : r->variables()) g_text += n + "\n";
        }))
        return nullptr;
    return g_text.c_str();
}
int hlmio_read_double(const char* path, const char* var, const long long* start, const long long* count, int rank, double* out) {
    return guarded([&] {
        auto r = hlmnc::open_reader(path);
        std::vector<uint64_t> s(start, start + rank), c(count, count + rank);
        r->read_into<double>(var, s, c, out);
    });
}
/// NetCDFLoader::loadTimeChunk through the reference's class (float, like nc_get_vara_float)
int hlmio_load_time_chunk(const char* path, const char* var, long long start, long long n, float* out, long long* dims3, double* dt_hours) {
    return guarded([&] {
        NetCDFLoader l(path, var);
        l.verbose = false;
        if (dims3) { dims3[0] = (long long)l.getTimeSize(); dims3[1] = (long long)l.getLatSize(); dims3[2] = (long long)l.getLonSize(); }
        if (dt_hours) *dt_hours = l.timeStepHours();
        if (out) {
            auto d = l.loadTimeChunk((size_t)start, (size_t)n);
            std::memcpy(out, d.get(), sizeof(float) * (size_t)n * l.getLatSize() * l.getLonSize());
        }
    });
}
/// format: 0 = NetCDF-4 (the reference's container), 1 = classic; level = deflate level of `outputs` (NetCDF-4)
void hlmio_write_dense_netcdf(const char* path, const double* dense, const double* t, const int* ids, const int* states, int nq, int ns, int n_eq,
                              int format, int level) {
    write_dense_netcdf(path, dense, t, ids, states, nq, ns, n_eq, level, format == 1 ? NcFormat::kClassic : NcFormat::kNetcdf4);
}
void hlmio_write_final_netcdf(const char* path, const double* fin, const int* ids, const int* states, int ns, int n_eq, int format, int level) {
    write_final_netcdf(path, fin, ids, states, ns, n_eq, level, format == 1 ? NcFormat::kClassic : NcFormat::kNetcdf4);
}
/// the windowed writer: windows of `qw` queries each taken from the full array [ns][nq][n_eq]; store_float: the file
/// variable is NC_FLOAT; staging_bytes bounds the NetCDF-4 writer's slab (small values force many chunk-rows)
int hlmio_write_dense_windows(const char* path, const double* dense, const double* t, const int* ids, const int* states, int n_states,
                              int nq, int ns, int n_eq, int qw, int format, int level, int store_float, long long staging_bytes) {
    return guarded([&] {
        DenseSeriesWriter w(path, std::vector<double>(t, t + nq), std::vector<int>(ids, ids + ns), std::vector<int>(states, states + n_states), n_eq,
                            store_float != 0, format == 1 ? NcFormat::kClassic : NcFormat::kNetcdf4, level,
                            staging_bytes > 0 ? (uint64_t)staging_bytes : (256ULL << 20));
        std::vector<double> win;
        for (int q0 = 0; q0 < nq; q0 += qw) {
            const int q1 = std::min(nq, q0 + qw);
            win.assign((size_t)ns * (q1 - q0) * n_eq, 0.0);
            for (int s = 0; s < ns; ++s)
                std::memcpy(&win[(size_t)s * (q1 - q0) * n_eq], dense + ((size_t)s * nq + q0) * n_eq, sizeof(double) * (q1 - q0) * n_eq);
            w.write_window(win.data(), q0, q1, q1 - q0);
        }
        w.close();
    });
}
/// HDF5's metadata checksum (lookup3), so a test can check it against the checksums in the reference's own files
unsigned int hlmio_lookup3(const unsigned char* p, long long n) { return hlmnc::lookup3(p, (size_t)n); }
/// load_config -> a flat JSON object of everything SimulationConfig holds
const char* hlmio_load_config_json(const char* path) {
    g_text.clear();
    if (guarded([&] {
            const SimulationConfig c = load_config(path);
            std::ostringstream o;
            o.precision(17);
            o << "{\"model.uid\":" << c.model.uid << ",\"model.name\":\"" << json_escape(c.model.name) << "\""
              << ",\"time.start\":\"" << json_escape(c.time.start_text) << "\",\"time.end\":\"" << json_escape(c.time.end_text) << "\""
              << ",\"time.minutes\":" << c.time.minutes() << ",\"initial.mode\":\"" << json_escape(c.initial.mode) << "\""
              << ",\"initial.file\":\"" << json_escape(c.initial.file) << "\",\"global_params\":" << c.global_params.size()
              << ",\"local_params.file\":\"" << json_escape(c.local_params.file) << "\",\"local_params.num_params\":" << c.local_params.num_params
              << ",\"forcings.type\":\"" << json_escape(c.forcings.type) << "\",\"forcings.path\":\"" << json_escape(c.forcings.path) << "\""
              << ",\"forcings.lookup\":\"" << json_escape(c.forcings.lookup_csv) << "\",\"forcings.precipitation\":\"" << json_escape(c.forcings.var_precip)
              << "\",\"forcings.temperature\":\"" << json_escape(c.forcings.var_temp) << "\",\"forcings.dt_precip\":" << c.forcings.dt_precip_hours
              << ",\"forcings.dt_temp\":" << c.forcings.dt_temp_hours << ",\"output.print_interval\":\"" << json_escape(c.output.print_interval) << "\""
              << ",\"output.print_minutes\":" << parse_interval_minutes(c.output.print_interval) << ",\"output.states\":[";
            for (size_t i = 0; i < c.output.states.size(); ++i) o << (i ? "," : "") << c.output.states[i];
            o << "],\"output.dir\":\"" << json_escape(c.output.dir) << "\",\"output.format\":\"" << json_escape(c.output.format) << "\""
              << ",\"output.dense\":" << (c.output.dense ? "true" : "false") << ",\"output.precision\":" << c.output.precision << ",\"solver.method\":\"" << json_escape(c.solver.method) << "\""
              << ",\"solver.override_tolerances\":" << (c.solver.override_tolerances ? "true" : "false") << ",\"solver.rtol\":" << c.solver.rtol
              << ",\"solver.atol\":" << c.solver.atol << ",\"solver.safety\":" << c.solver.safety << ",\"solver.min_scale\":" << c.solver.min_scale
              << ",\"solver.max_scale\":" << c.solver.max_scale << ",\"solver.override_initial_step\":" << (c.solver.override_initial_step ? "true" : "false")
              << ",\"solver.initial_step\":" << c.solver.initial_step << ",\"solver.interval\":\"" << json_escape(c.solver.interval) << "\""
              << ",\"solver.max_attempts\":" << c.solver.max_attempts
              << ",\"solver.stiff_fallback\":" << (c.solver.stiff_fallback ? "true" : "false")
              << ",\"routing.enabled\":" << (c.routing.enabled ? "true" : "false") << ",\"routing.couple\":\"" << json_escape(c.routing.couple) << "\""
              << ",\"routing.couple_minutes\":" << parse_interval_minutes(c.routing.couple)
              << ",\"routing.subbasin_links\":" << c.routing.subbasin_links << ",\"mpi.step_storage\":" << c.mpi.step_storage
              << ",\"mpi.transfer_buffer\":" << c.mpi.transfer_buffer << ",\"mpi.discontinuity_buf\":" << c.mpi.discontinuity_buf
              << ",\"flags.uses_dam\":" << (c.flags.uses_dam ? "true" : "false") << ",\"flags.convert_area\":" << (c.flags.convert_area ? "true" : "false") << "}";
            g_text = o.str();
        }))
        return nullptr;
    return g_text.c_str();
}


/// hlm_b200::plan_routes flattened for the tests.  order[n]; rank_lo[world + 1]; up_ptr_all: every rank's
/// up_ptr one after the other (n + world entries); up_idx_all / send_idx_all: every rank's lists one after the
/// other (at most n entries each); send_counts[world]; meta = {max_send, n_subbasins, n_cut_edges}.
int hlmio_route_plan(const long long* stream, const long long* next_stream, long long n, int world, long long subbasin_links,
                     long long* order, long long* rank_lo, long long* up_ptr_all, int* up_idx_all, int* send_idx_all,
                     long long* send_counts, long long* meta) {
    return guarded([&] {
        std::vector<long long> s(stream, stream + n), nx(next_stream, next_stream + n);
        const hlm_b200::RoutePlan p = hlm_b200::plan_routes(s, nx, world, subbasin_links);
        std::copy(p.order.begin(), p.order.end(), order);
        long long a = 0, b = 0, c = 0;
        for (int r = 0; r < world; ++r) {
            const hlm_b200::RankTopology& t = p.ranks[(size_t)r];
            rank_lo[r] = t.lo;
            rank_lo[r + 1] = t.hi;
            std::copy(t.up_ptr.begin(), t.up_ptr.end(), up_ptr_all + a);
            a += (long long)t.up_ptr.size();
            std::copy(t.up_idx.begin(), t.up_idx.end(), up_idx_all + b);
            b += (long long)t.up_idx.size();
            std::copy(t.send_idx.begin(), t.send_idx.end(), send_idx_all + c);
            c += (long long)t.send_idx.size();
            send_counts[r] = (long long)t.send_idx.size();
        }
        meta[0] = p.max_send;
        meta[1] = p.n_subbasins;
        meta[2] = p.n_cut_edges;
    });
}

}  // extern "C"
