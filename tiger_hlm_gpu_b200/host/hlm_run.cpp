// hlm_run.cpp — the run driver: config.yaml in, final/dense NetCDF (or CSV) out.
//
// Replaces the worker-rank body of the reference's main() (main.cpp:314-823) on top of the C ABI:
//   config.yaml (I_O/config_loader.cpp)            -> every constant main.cpp hard-codes
//   parameters CSV (I_O/parameters_loader.cpp)     -> this rank's contiguous row chunk (main.cpp:275-307;
//                                                     each rank reads its own rows, no MPI scatter)
//   lookup CSV + forcing NetCDF (forcing_loader)   -> forcing GRID on the device, loaded by time chunks,
//                                                     link -> cell column (main.cpp:495-505); the
//                                                     per-link expansion of main.cpp:507-549 is gone
//   rk45 (solver/rk45_api.hpp)                     -> hlm_solve_* session, driven in intervals
//   write_final_netcdf / write_dense_netcdf        -> same files per rank (main.cpp:796-797), the dense
//                                                     series written window by window while the next
//                                                     window integrates (two pinned buffers)
//
// One process per GPU.  Rank and world size come from --rank/--world or RANK/WORLD_SIZE/LOCAL_RANK
// (torchrun / mpirun style launchers).  Uncoupled models: ranks never talk to each other, links are independent.
// Routed runs (routing.enabled): links are dealt to ranks by sub-basin and the ranks all-gather the boundary
// links' discharge once per coupling interval over NCCL (hlm_nccl.hpp; the id file goes to output.dir).
//
// A single process can also own several GPUs (--devices 0,1,2,3): one host thread and one context per listed device,
// thread r taking rank r of len(list) — the reference's one MPI rank per GPU (main.cpp:314-319) folded into one
// process, for uncoupled models (a routed run needs the ranks' collective: one process per GPU).
//
// usage: hlm_run CONFIG.yaml [--rank R] [--world W] [--device D | --devices D0,D1,...] [--root DIR] [--quiet]
#include <dirent.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <future>
#include <set>
#include <thread>

#include "hlm_config.hpp"
#include "hlm_host.hpp"
#include "hlm_routing.hpp"
#include "hlm_netcdf.hpp"
#ifdef HLM_HAVE_NCCL
#include "hlm_nccl.hpp"  // routed runs over several ranks: the boundary all-gather
#endif

namespace {

struct Options {
    std::string config, root;
    int rank = 0, world = 1, device = -1;
    bool quiet = false;
    std::vector<int> devices;  // --devices: one thread (rank) per entry
};

int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

std::string join_path(const std::string& base, const std::string& p) {
    if (p.empty() || p[0] == '/' || base.empty()) return p;
    return base + "/" + p;
}
std::string dir_of(const std::string& p) {
    const size_t k = p.find_last_of('/');
    return k == std::string::npos ? "." : p.substr(0, k);
}
bool ends_with(const std::string& s, const std::string& suf) {
    return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0;
}
void check(int rc, const char* what) {
    if (rc != HLM_OK) throw std::runtime_error(std::string(what) + ": " + hlm_last_error());
}

// rows [lo, hi) of `rank` under the reference's chunk rule (main.cpp:275-307)
void shard_range(long long n, int world, int rank, long long& lo, long long& hi) {
    const long long base = n / world, rem = n % world;
    lo = rank * base + std::min<long long>(rank, rem);
    hi = lo + base + (rank < rem ? 1 : 0);
}

// first *.nc file (sorted by name) of `folder` that holds variable `var`
std::string find_forcing_file(const std::string& folder, const std::string& var) {
    std::set<std::string> names;
    if (DIR* d = opendir(folder.c_str())) {
        while (dirent* e = readdir(d)) {
            const std::string n = e->d_name;
            if (ends_with(n, ".nc") || ends_with(n, ".nc4") || ends_with(n, ".cdf")) names.insert(n);
        }
        closedir(d);
    } else {
        throw std::runtime_error("forcings.path: cannot open folder " + folder);
    }
    for (auto& n : names) {
        try {
            auto r = hlmnc::open_reader(folder + "/" + n);
            if (r->has_variable(var) && r->inquire(var).shape.size() == 3) return folder + "/" + n;
        } catch (const std::exception&) {
        }
    }
    throw std::runtime_error("no NetCDF file in " + folder + " holds a (time, lat, lon) variable '" + var + "'");
}

struct ForcingSource {
    std::unique_ptr<NetCDFLoader> loader;
    double dt_hours = 1.0;
    long long nT = 0;
    long long res_i0 = -1, res_n = 0;  // resident chunk
    long long index(double t_min) const {  // solver/rk45_kernel.cu:90-98
        const double r = t_min / (dt_hours * 60.0);
        return r < 0.0 ? 0 : (r >= (double)nT ? nT - 1 : (long long)r);
    }
};

// initial states: cold = the reference's common y0 (main.cpp:376), hot = a file
std::vector<double> initial_states(const SimulationConfig& cfg, const std::string& root, const std::vector<SpatialParams>& sp,
                                   int n_eq, double& t_start_minutes) {
    const size_t ns = sp.size();
    std::vector<double> y0(ns * n_eq);
    t_start_minutes = 0.0;
    if (cfg.initial.mode == "cold") {
        const double common204[5] = {0.01, 3.0, 0.0, 5.0, 0.2};
        const double common200[5] = {0.5, 3.0, 0.0, 5.0, 0.2};  // channel discharge 0.5 m3/s + Model204's stores
        for (size_t s = 0; s < ns; ++s)
            for (int i = 0; i < n_eq; ++i) y0[s * n_eq + i] = (cfg.model.uid == 204 && i < 5) ? common204[i] : ((cfg.model.uid == 200 && i < 5) ? common200[i] : 1.0);
        return y0;
    }
    if (cfg.initial.mode != "hot") throw std::runtime_error("initial.mode must be 'cold' or 'hot'");
    const std::string path = join_path(root, cfg.initial.file);
    if (ends_with(path, ".nc")) {
        // a final-state file of an earlier run: outputs(system, variable) keyed by the link ids in `system`
        auto r = hlmnc::open_reader(path);
        const std::vector<double> out = r->read_all<double>("outputs");
        const std::vector<long long> ids = r->read_all<long long>("system");
        const hlmnc::VarInfo& v = r->inquire("outputs");
        if (v.shape.size() != 2 || (int)v.shape[1] != n_eq) throw std::runtime_error("hot start: " + path + " is not outputs(system," + std::to_string(n_eq) + ")");
        std::unordered_map<long long, size_t> row;
        for (size_t k = 0; k < ids.size(); ++k) row[ids[k]] = k;
        for (size_t s = 0; s < ns; ++s) {
            auto it = row.find((long long)sp[s].stream);
            if (it == row.end()) throw std::runtime_error("hot start: link " + std::to_string(sp[s].stream) + " is not in " + path);
            for (int i = 0; i < n_eq; ++i) y0[s * n_eq + i] = out[it->second * n_eq + i];
        }
        return y0;
    }
    std::ifstream in(path);
    if (!in.is_open()) throw std::runtime_error("hot start: cannot open " + path);
    if (ends_with(path, ".csv")) {  // final.csv layout (main.cpp:734-748): one row per link, in link order
        std::string line;
        std::getline(in, line);
        for (size_t s = 0; s < ns; ++s) {
            if (!std::getline(in, line)) throw std::runtime_error("hot start: " + path + " has fewer rows than links");
            std::istringstream ss(line);
            std::string cell;
            for (int i = 0; i < n_eq; ++i) {
                if (!std::getline(ss, cell, ',')) throw std::runtime_error("hot start: short row in " + path);
                y0[s * n_eq + i] = std::stod(cell);
            }
        }
        return y0;
    }
    // .uini — uniform initial conditions: model uid, initial time (minutes), then N_EQ values for every link
    std::vector<double> v;
    std::string tok;
    while (in >> tok) {
        if (tok[0] == '%' || tok[0] == '#') { std::getline(in, tok); continue; }
        v.push_back(std::stod(tok));
    }
    if ((int)v.size() < 2 + n_eq) throw std::runtime_error("hot start: " + path + " needs uid, time and " + std::to_string(n_eq) + " values");
    if ((int)v[0] != cfg.model.uid) throw std::runtime_error("hot start: " + path + " is for model " + std::to_string((int)v[0]));
    t_start_minutes = v[1];
    for (size_t s = 0; s < ns; ++s)
        for (int i = 0; i < n_eq; ++i) y0[s * n_eq + i] = v[2 + i];
    return y0;
}

struct Pinned {
    double* p = nullptr;
    size_t elems = 0;
    void reserve(size_t n) {
        if (n <= elems) return;
        if (p) hlm_host_free(p);
        p = nullptr;
        check(hlm_host_alloc((void**)&p, (long long)(n * sizeof(double))), "hlm_host_alloc");
        elems = n;
    }
    ~Pinned() { if (p) hlm_host_free(p); }
};

int run(const Options& opt) {
    using clock = std::chrono::steady_clock;
    const auto wall0 = clock::now();
    const SimulationConfig cfg = load_config(opt.config);
    const std::string root = opt.root.empty() ? dir_of(opt.config) : opt.root;
    auto say = [&](const std::string& s) { if (!opt.quiet) std::printf("[rank %d] %s\n", opt.rank, s.c_str()); };

    int n_eq = 0, n_sp = 0, n_forc = 0;
    check(hlm_model_info(cfg.model.uid, &n_eq, &n_sp, &n_forc), "model.uid");
    if (cfg.solver.method != "RK45") throw std::runtime_error("solver.method: only RK45 is available");

    // ---- per-link parameters: this rank's rows ----
    // (routed runs: whole sub-basins in contiguous runs of the planned order, so that only sub-basin outlets that
    // drain into another rank's sub-basin cross ranks; the plan is the same on every rank)
    const bool routed = cfg.routing.enabled;
    std::vector<SpatialParams> sp;
    hlm_b200::RoutePlan plan;
    if (n_sp > 0 || !cfg.local_params.file.empty()) {
        std::vector<SpatialParams> all = loadSpatialParams(join_path(root, cfg.local_params.file));
        if (routed) {
            std::vector<long long> stream(all.size()), next(all.size());
            for (size_t s = 0; s < all.size(); ++s) {
                stream[s] = all[s].stream;
                next[s] = all[s].next_stream;
            }
            plan = hlm_b200::plan_routes(stream, next, opt.world, cfg.routing.subbasin_links);
            const hlm_b200::RankTopology& tp = plan.ranks[(size_t)opt.rank];
            sp.resize((size_t)tp.n_local());
            for (long long k = 0; k < tp.n_local(); ++k) sp[(size_t)k] = all[(size_t)plan.order[(size_t)(tp.lo + k)]];
            say("links: " + std::to_string(sp.size()) + " of " + std::to_string(all.size()) + " (" + std::to_string(plan.n_subbasins) +
                " sub-basins over " + std::to_string(opt.world) + " rank(s), " + std::to_string(plan.n_cut_edges) + " cut edges)");
        } else {
            long long lo, hi;
            shard_range((long long)all.size(), opt.world, opt.rank, lo, hi);
            sp.assign(all.begin() + lo, all.begin() + hi);
            say("links " + std::to_string(lo) + ".." + std::to_string(hi) + " of " + std::to_string(all.size()));
        }
    }
    const long long ns = (long long)sp.size();
    if (ns == 0) { say("no links for this rank"); return 0; }

    hlm_ctx* ctx = nullptr;
    const int device = opt.device >= 0 ? opt.device : env_int("LOCAL_RANK", 0);
    check(hlm_create(device, &ctx), "hlm_create");
    struct CtxGuard { hlm_ctx* c; ~CtxGuard() { hlm_destroy(c); } } guard{ctx};

    // ---- solver parameters (main.cpp:618-640; the reference's "auto" initial step evaluates to 1e-6) ----
    const double prm[6] = {cfg.solver.override_initial_step ? cfg.solver.initial_step : 1e-6, cfg.solver.rtol, cfg.solver.atol,
                           cfg.solver.safety, cfg.solver.min_scale, cfg.solver.max_scale};
    check(hlm_set_model_parameters(ctx, cfg.model.uid, prm), "hlm_set_model_parameters");
    check(hlm_set_max_attempts(ctx, cfg.solver.max_attempts), "hlm_set_max_attempts");
    check(hlm_upload_spatial_params(ctx, sp.data(), ns, (long long)sizeof(SpatialParams)), "hlm_upload_spatial_params");

    // ---- forcings ----
    std::vector<ForcingSource> forc;
    if (n_forc > 0) {
        if (cfg.forcings.type != "folder_nc") throw std::runtime_error("forcings.type: only folder_nc is available");
        const std::string folder = join_path(root, cfg.forcings.path);
        const std::string vars[2] = {cfg.forcings.var_precip, cfg.forcings.var_temp};
        const double dt_cfg[2] = {cfg.forcings.dt_precip_hours, cfg.forcings.dt_temp_hours};
        const double dt_ref[2] = {1.0, 24.0};  // main.cpp:520-523
        for (int j = 0; j < n_forc && j < 2; ++j) {
            ForcingSource f;
            f.loader.reset(new NetCDFLoader(find_forcing_file(folder, vars[j]), vars[j]));
            f.loader->verbose = false;
            f.nT = (long long)f.loader->getTimeSize();
            const double dt_file = f.loader->timeStepHours();
            f.dt_hours = dt_cfg[j] > 0 ? dt_cfg[j] : (dt_file > 0 ? dt_file : dt_ref[j]);
            say("forcing " + std::to_string(j) + " '" + vars[j] + "' from " + f.loader->getFileName() + ": " + std::to_string(f.nT) +
                " x " + std::to_string(f.loader->getLatSize()) + " x " + std::to_string(f.loader->getLonSize()) + ", dt " +
                std::to_string(f.dt_hours) + " h");
            forc.push_back(std::move(f));
        }
        for (size_t j = 1; j < forc.size(); ++j)
            if (forc[j].loader->getLatSize() != forc[0].loader->getLatSize() || forc[j].loader->getLonSize() != forc[0].loader->getLonSize())
                throw std::runtime_error("forcings: all variables must share one (lat, lon) grid");
        LookupMapper lm(join_path(folder, cfg.forcings.lookup_csv));
        if (!lm.load()) throw std::runtime_error("Lookup load failed");
        const std::vector<int> col = forcingColumns(sp, lm, (int)forc[0].loader->getLonSize());
        const long long ncells = (long long)(forc[0].loader->getLatSize() * forc[0].loader->getLonSize());
        for (int c : col)
            if (c < 0 || c >= ncells) throw std::runtime_error("forcing lookup points outside the grid");
        check(hlm_set_forcing_columns(ctx, col.data(), ns), "hlm_set_forcing_columns");
    }

    // ---- time axis, queries, intervals ----
    // t counts minutes from time.origin (= time.start unless given), which is also forcing sample 0
    double t_uini = 0.0;
    const std::vector<double> y0 = initial_states(cfg, root, sp, n_eq, t_uini);
    const double t_begin = cfg.time.start_minutes() + t_uini;
    const double t_end = cfg.time.end_minutes();
    if (t_begin < 0.0) throw std::runtime_error("time.start must not lie before time.origin");
    if (!(t_end > t_begin)) throw std::runtime_error("time.end must lie after the start of the run");
    const double dq = parse_interval_minutes(cfg.output.print_interval);
    std::vector<double> tq;
    for (double t = t_begin; t <= t_end; t += dq) tq.push_back(t);  // main.cpp:653-657
    // routed run: the interval is the coupling interval and every link continues across it (hlm_solve_advance)
#ifdef HLM_HAVE_NCCL
    std::unique_ptr<hlm_b200::NcclBoundaryExchange> exchange;
#endif
    const bool multi_rank_routed = routed && opt.world > 1;
    if (routed) {
        const hlm_b200::RankTopology& tp = plan.ranks[(size_t)opt.rank];
        check(hlm_route_set_topology(ctx, tp.up_ptr.data(), tp.up_idx.empty() ? nullptr : tp.up_idx.data(), ns,
                                     tp.send_idx.empty() ? nullptr : tp.send_idx.data(), (long long)tp.send_idx.size()),
              "hlm_route_set_topology");
        if (multi_rank_routed) {
#ifdef HLM_HAVE_NCCL
            exchange.reset(new hlm_b200::NcclBoundaryExchange(opt.world, opt.rank, device, join_path(root, cfg.output.dir), plan.max_send,
                                                              plan.halo_len()));
            check(hlm_set_stream(ctx, exchange->stream()), "hlm_set_stream");  // kernels and collectives on one stream
            if (plan.max_send > 0) check(hlm_route_set_send_buffer(ctx, exchange->d_send()), "hlm_route_set_send_buffer");
#else
            throw std::runtime_error("routing with WORLD_SIZE > 1 needs the NCCL build of hlm_run (nccl.h was not found when it was built)");
#endif
        }
        say("routing: " + std::to_string(tp.up_idx.size()) + " upstream entries, " + std::to_string(tp.send_idx.size()) +
            " boundary links, coupling interval " + cfg.routing.couple);
    }
    check(hlm_set_stiff_fallback(ctx, (cfg.solver.stiff_fallback || routed) ? 1 : 0), "hlm_set_stiff_fallback");
    // Model 200 (project-defined), routed or not: the raised limit.  With the reference's 5 the transient of its first
    // day flags 1.5 % of the links stiff that are not (kinks of the min/max terms), and the implicit fallback spends
    // 2.4 s on them where the whole day of 1 M links takes 0.1 s (tools/probe_model200_day0.py).
    if (routed || cfg.model.uid == 200) check(hlm_set_reject_limit(ctx, hlm_b200::kRoutedRejectLimit), "hlm_set_reject_limit");
    const double interval = parse_interval_minutes(routed ? cfg.routing.couple : cfg.solver.interval);
    const double chunk = std::max(interval, parse_interval_minutes("30d"));

    std::vector<int> states = cfg.output.states;
    if (states.empty()) for (int i = 0; i < n_eq; ++i) states.push_back(i);
    std::vector<int> linkids(ns);
    for (long long s = 0; s < ns; ++s) linkids[s] = (int)sp[s].stream;  // real link ids (the reference writes 0..ns-1, main.cpp:788-793)

    const std::string out_dir = join_path(root, cfg.output.dir);
    const std::string suffix = "_rank_" + std::to_string(opt.rank);
    const bool csv = cfg.output.format == "csv";
    const NcFormat nc_format = cfg.output.format == "netcdf3" ? NcFormat::kClassic : NcFormat::kNetcdf4;
    std::unique_ptr<DenseSeriesWriter> dense_writer;
    std::vector<double> dense_all;  // csv only
    // NetCDF output: output.states and output.precision are applied where the records are produced — the window
    // kernel stores only the selected states (as float if asked), so only those bytes cross PCIe (the reference
    // ships nothing selectively: main.cpp:788-793 writes every state).  The CSV writer keeps all states.
    int rec_cols = n_eq;
    size_t rec_elem = sizeof(double);
    if (cfg.output.dense && !csv) {
        const bool f32 = cfg.output.precision == 32;
        dense_writer.reset(new DenseSeriesWriter(out_dir + "/" + cfg.output.prefix + "dense" + suffix + ".nc", tq, linkids, states, n_eq, f32,
                                                 nc_format, cfg.output.compression_level));
        dense_writer->set_packed_source();
        check(hlm_set_output_states(ctx, dense_writer->output_mask()), "hlm_set_output_states");
        check(hlm_set_output_precision(ctx, f32 ? 32 : 64), "hlm_set_output_precision");
        int bytes = 8;
        check(hlm_output_layout(ctx, cfg.model.uid, &rec_cols, &bytes), "hlm_output_layout");
        rec_elem = (size_t)bytes;
    }
    if (cfg.output.dense && csv) dense_all.assign((size_t)ns * tq.size() * n_eq, 0.0);

    // window size in queries: two pinned host buffers of at most 1 GiB each
    const long long per_q = ns * rec_cols * (long long)rec_elem;
    const long long qw_max = std::max<long long>(1, (1LL << 30) / per_q);
    Pinned pin[2];
    std::future<void> writing[2];

    auto upload_forcing_for = [&](double ta, double tb) {
        for (size_t j = 0; j < forc.size(); ++j) {
            ForcingSource& f = forc[j];
            const long long need_lo = f.index(ta), need_hi = f.index(tb);
            if (f.res_i0 >= 0 && need_lo >= f.res_i0 && need_hi < f.res_i0 + f.res_n) continue;
            const long long i0 = need_lo;
            const long long n = std::min<long long>(f.nT - i0, std::max<long long>(need_hi - need_lo + 1, (long long)(chunk / (f.dt_hours * 60.0)) + 1));
            auto data = f.loader->loadTimeChunk((size_t)i0, (size_t)n);
            const long long ncells = (long long)(f.loader->getLatSize() * f.loader->getLonSize());
            check(hlm_upload_forcing_chunk(ctx, (int)j, f.dt_hours, f.nT, i0, n, ncells, data.get()), "hlm_upload_forcing_chunk");
            check(hlm_synchronize(ctx), "hlm_synchronize");  // `data` is pageable and freed on return
            f.res_i0 = i0;
            f.res_n = n;
        }
    };

    // ---- the run ----
    const auto solve0 = clock::now();
    size_t q_next = 0;  // first query not yet assigned to an interval
    bool first = true;
    int slot = 0;
    long long k_int = 0;  // interval counter of next_interval_boundary (hlm_config.hpp)
    for (double ta = t_begin; ta < t_end;) {
        const double tb = next_interval_boundary(ta, t_end, interval, k_int);  // multiples of the interval from the origin
        size_t q_end = q_next;
        while (q_end < tq.size() && tq[q_end] <= tb) ++q_end;
        const long long nq = (long long)(q_end - q_next);
        upload_forcing_for(ta, tb);
        if (first) check(hlm_solve_begin(ctx, cfg.model.uid, y0.data(), ns, ta, tb, tq.data() + q_next, nq), "hlm_solve_begin");
        if (routed) {  // inflow of this interval from the state at its start
            double* halo = nullptr;
#ifdef HLM_HAVE_NCCL
            if (exchange && plan.max_send > 0) {
                if (first) check(hlm_route_pack(ctx), "hlm_route_pack");  // afterwards the window kernel's epilogue has packed it
                exchange->all_gather();
                halo = exchange->d_halo();
            }
#endif
            check(hlm_route_gather(ctx, halo), "hlm_route_gather");
        }
        if (!first) {
            if (routed) check(hlm_solve_advance(ctx, tb, tq.data() + q_next, nq), "hlm_solve_advance");
            else check(hlm_solve_restart(ctx, ta, tb, tq.data() + q_next, nq), "hlm_solve_restart");
        }
        first = false;
        const bool want_dense = cfg.output.dense && nq > 0;
        for (long long q = 0;;) {
            q = std::min(nq, q + qw_max);
            check(hlm_solve_window(ctx, q >= nq ? nq : q, want_dense ? 1 : 0), "hlm_solve_window");
            if (want_dense) {
                long long wlo = 0, whi = 0;
                check(hlm_solve_window_buffer(ctx, nullptr, &wlo, &whi), "hlm_solve_window_buffer");
                if (whi > wlo) {
                    if (writing[slot].valid()) writing[slot].get();  // the writer is done with this buffer
                    pin[slot].reserve(((size_t)ns * (size_t)(whi - wlo) * rec_cols * rec_elem + 7) / 8);
                    int ticket = 0;
                    check(hlm_solve_fetch_window_packed(ctx, pin[slot].p, &ticket), "hlm_solve_fetch_window_packed");
                    const double* src = pin[slot].p;
                    const size_t g_lo = q_next + (size_t)wlo, g_hi = q_next + (size_t)whi;
                    writing[slot] = std::async(std::launch::async, [&, src, g_lo, g_hi, ticket]() {
                        check(hlm_solve_wait_copy(ctx, ticket), "hlm_solve_wait_copy");
                        if (dense_writer) dense_writer->write_window(src, g_lo, g_hi, g_hi - g_lo);
                        else
                            for (long long s = 0; s < ns; ++s)
                                std::memcpy(&dense_all[((size_t)s * tq.size() + g_lo) * n_eq], src + (size_t)s * (g_hi - g_lo) * n_eq,
                                            (g_hi - g_lo) * n_eq * sizeof(double));
                    });
                    slot ^= 1;
                }
            }
            if (q >= nq) break;
        }
        q_next = q_end;
        ta = tb;
    }
    std::vector<double> y_final((size_t)ns * n_eq);
    std::vector<int> code(ns);
    std::vector<long long> n_acc(ns), n_rej(ns), n_jump(ns);
    check(hlm_solve_end(ctx, y_final.data(), code.data(), n_acc.data(), n_rej.data(), n_jump.data()), "hlm_solve_end");
    for (auto& w : writing) if (w.valid()) w.get();
    const double solve_s = std::chrono::duration<double>(clock::now() - solve0).count();

    // ---- outputs ----
    std::vector<int> all_states(n_eq);
    for (int i = 0; i < n_eq; ++i) all_states[i] = i;
    if (csv) {
        write_final_csv(out_dir + "/" + cfg.output.prefix + "final" + suffix + ".csv", y_final, (int)ns, n_eq);
        if (cfg.output.dense) write_dense_csv(out_dir + "/" + cfg.output.prefix + "dense" + suffix + ".csv", dense_all, tq, (int)ns, n_eq);
    } else {
        write_final_netcdf(out_dir + "/" + cfg.output.prefix + "final" + suffix + ".nc", y_final.data(), linkids.data(), all_states.data(), (int)ns, n_eq,
                           cfg.output.compression_level, nc_format);
        if (dense_writer) dense_writer->close();
    }
    long long acc = 0, rej = 0, jump = 0, stiff = 0, stalled = 0, solved = 0;
    for (long long s = 0; s < ns; ++s) {
        acc += n_acc[s]; rej += n_rej[s]; jump += n_jump[s];
        stiff += code[s] == HLM_LINK_STIFF;
        solved += code[s] == HLM_LINK_STIFF_SOLVED;
        stalled += code[s] == HLM_LINK_STALLED;
    }
    const double wall_s = std::chrono::duration<double>(clock::now() - wall0).count();
    std::printf("[rank %d] done: %lld links, %zu queries, t %.1f -> %.1f min; accepted %lld rejected %lld slope-jump %lld; "
                "stiff %lld (+%lld finished implicitly) stalled %lld; solve %.3f s (%.3e accepted steps/s), total %.3f s\n",
                opt.rank, ns, tq.size(), t_begin, t_end, acc, rej, jump, stiff, solved, stalled, solve_s, acc / std::max(solve_s, 1e-9), wall_s);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    Options opt;
    opt.rank = env_int("RANK", env_int("OMPI_COMM_WORLD_RANK", 0));
    opt.world = env_int("WORLD_SIZE", env_int("OMPI_COMM_WORLD_SIZE", 1));
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string {
            if (i + 1 >= argc) throw std::runtime_error("missing value after " + a);
            return argv[++i];
        };
        try {
            if (a == "--rank") opt.rank = std::stoi(next());
            else if (a == "--world") opt.world = std::stoi(next());
            else if (a == "--device") opt.device = std::stoi(next());
            else if (a == "--devices") {
                std::stringstream ss(next());
                for (std::string tok; std::getline(ss, tok, ',');) opt.devices.push_back(std::stoi(tok));
                if (opt.devices.empty()) throw std::runtime_error("--devices needs a comma-separated list of CUDA device ids");
            }
            else if (a == "--root") opt.root = next();
            else if (a == "--quiet") opt.quiet = true;
            else if (!a.empty() && a[0] != '-' && opt.config.empty()) opt.config = a;
            else throw std::runtime_error("unknown argument " + a);
        } catch (const std::exception& e) {
            std::fprintf(stderr, "hlm_run: %s\n", e.what());
            return 2;
        }
    }
    if (opt.config.empty() || opt.world < 1 || opt.rank < 0 || opt.rank >= opt.world) {
        std::fprintf(stderr, "usage: %s CONFIG.yaml [--rank R] [--world W] [--device D | --devices D0,D1,...] [--root DIR] [--quiet]\n", argv[0]);
        return 2;
    }
    auto guarded = [](const Options& o) {
        try {
            return run(o);
        } catch (const std::exception& e) {  // main.cpp prints and returns 1 on any failure
            std::fprintf(stderr, "[rank %d] error: %s\n", o.rank, e.what());
            return 1;
        }
    };
    if (opt.devices.empty()) return guarded(opt);
    // one process, several GPUs: thread r is rank r of devices.size() on device devices[r]
    try {
        if (load_config(opt.config).routing.enabled && opt.devices.size() > 1)
            throw std::runtime_error("routing with --devices: a routed run over several GPUs needs one process per GPU "
                                     "(RANK / WORLD_SIZE / LOCAL_RANK), the ranks all-gather over NCCL");
    } catch (const std::exception& e) {
        std::fprintf(stderr, "hlm_run: %s\n", e.what());
        return 1;
    }
    std::vector<int> rc(opt.devices.size(), 0);
    std::vector<std::thread> threads;
    for (size_t r = 0; r < opt.devices.size(); ++r) {
        Options o = opt;
        o.rank = (int)r;
        o.world = (int)opt.devices.size();
        o.device = opt.devices[r];
        o.devices.clear();
        threads.emplace_back([&rc, r, o, &guarded] { rc[r] = guarded(o); });
    }
    for (auto& t : threads) t.join();
    for (int v : rc)
        if (v) return v;
    return 0;
}
