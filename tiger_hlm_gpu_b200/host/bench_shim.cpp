// bench_shim.cpp — end-to-end throughput of the path through the C++ shim the integration guide tells a
// reference-side caller to use: rk45_api::run_rk45<Model204>() of include/hlm_b200/rk45_api.hpp, the
// reference's own operator (solver/rk45_api.hpp:273-313), with host vectors in and out.
//
//   hlm_bench_shim NS STEPS WARMUP      -> one JSON line on stdout
//
// Workload: bench.py's (BASELINE configs[3]): Model204, NS synthetic links with the constants of
// data/small_test.csv, hourly precipitation + daily temperature on a forcing grid of ~537 links per cell,
// one step = one simulated day with 24 hourly dense records per link; every step hands the previous day's
// final states back in as h_y0 and receives final + dense states by value, as a chained reference run does.
// The inputs are generated here (same distributions as tiger_hlm_gpu_b200/synthetic.py, another generator), so
// the figure is comparable with bench.py's e2e, not bit-identical in its step counts.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../include/hlm_b200/rk45_api.hpp"

int main(int argc, char** argv) {
    const long long ns = argc > 1 ? std::atoll(argv[1]) : 1000000;
    const int steps = argc > 2 ? std::atoi(argv[2]) : 5, warmup = argc > 3 ? std::atoi(argv[3]) : 3;
    const int days = steps + warmup + 1;
    try {
        std::mt19937_64 rng(204);
        std::uniform_real_distribution<double> U(0.0, 1.0);
        const double c1 = 0.001 / 60.0;  // I_O/parameters_loader.cpp:57
        std::vector<SpatialParams> sp((size_t)ns);
        for (long long i = 0; i < ns; ++i) {
            SpatialParams& p = sp[(size_t)i];
            p.stream = 420000000 + i;
            p.next_stream = 420000000 + i / 2;
            p.c1 = c1; p.infil = 4.0 * c1; p.perco = 1.6 * c1; p.Hu = 178.0; p.lat = 40.3; p.sw = 0.11; p.ss = 0.33;
            p.n_mann = 0.1; p.slope = 0.02;
            p.L = 0.09 + (2.1 - 0.09) * U(rng);
            p.A_h = std::exp(std::log(0.13) + (std::log(1.6) - std::log(0.13)) * U(rng));
            p.alpha3 = 2.0 * 1440.0; p.alpha4 = 55.0 * 1440.0; p.melt_f = 3.7; p.temp_thr = 0.0;
        }
        const long long per_cell = 537, ncells = (ns + per_cell - 1) / per_cell;
        std::vector<int> col((size_t)ns);
        for (long long i = 0; i < ns; ++i) col[(size_t)i] = (int)(i / per_cell);
        std::vector<float> pr((size_t)(24 * days) * ncells), t2m((size_t)days * ncells);
        std::exponential_distribution<double> E(0.5);  // mean 2
        std::normal_distribution<double> Nrm(0.0, 2.0);
        for (auto& v : pr) v = (float)(U(rng) >= 0.85 ? c1 * E(rng) : 0.0);
        for (int d = 0; d < days; ++d)
            for (long long c = 0; c < ncells; ++c)
                t2m[(size_t)d * ncells + c] = (float)(10.0 + 8.0 * std::sin(2.0 * M_PI * d / 365.0) + Nrm(rng));

        Model204::Parameters prm;
        prm.initialStep = 1e-6;  // main.cpp:633-640
        rk45_api::setModelParameters<Model204>(prm);
        auto& ctx = hlm_b200::default_context();
        hlm_b200::check(hlm_set_max_attempts(ctx.get(), 5000000), "hlm_set_max_attempts");
        rk45_api::setForcing(0, 1.0, 24 * days, ncells, pr.data());
        rk45_api::setForcing(1, 24.0, days, ncells, t2m.data());
        rk45_api::setForcingColumns(col.data(), ns);
        ctx.setSpatialParams(sp.data(), ns);  // resident; run_rk45's d_sp stays nullptr below (the reference uploads once too)

        rk45_api::FinalType y;  // page-locked: day k's final states are day k+1's h_y0
        y.resize((size_t)ns * 5);
        const double y0[5] = {0.01, 3.0, 0.0, 5.0, 0.2};  // main.cpp:376
        for (long long i = 0; i < ns; ++i)
            for (int c = 0; c < 5; ++c) y[(size_t)i * 5 + c] = y0[c];
        std::vector<double> tq(24);
        long long acc_timed = 0;
        double checksum = 0.0;
        std::chrono::steady_clock::time_point t_start;
        for (int k = 0; k < warmup + steps; ++k) {
            if (k == warmup) t_start = std::chrono::steady_clock::now();
            for (int q = 0; q < 24; ++q) tq[(size_t)q] = 1440.0 * k + 60.0 * (q + 1);
            auto [fin, dense] = rk45_api::run_rk45<Model204>(y, 1440.0 * k, 1440.0 * (k + 1), tq, (const SpatialParams*)nullptr);
            long long tot[7];
            hlm_b200::check(hlm_solve_totals(ctx.get(), tot), "hlm_solve_totals");
            if (k >= warmup) acc_timed += tot[0];
            checksum += dense[dense.size() - 1] + fin[0];
            y = std::move(fin);  // the dense vector goes back to the pinned pool here
        }
        const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
        std::printf("{\"value\": %.6e, \"unit\": \"accepted system-steps/s\", \"ms_per_step\": %.3f, \"steps\": %d, \"warmup\": %d, "
                    "\"links\": %lld, \"h2d_bytes_per_step\": %lld, \"d2h_bytes_per_step\": %lld, \"checksum\": %.17g, "
                    "\"api\": \"rk45_api::run_rk45<Model204>(h_y0, t0, tf, h_query_times, d_sp) of include/hlm_b200/rk45_api.hpp "
                    "(host/bench_shim.cpp): results returned by value in pooled page-locked vectors, not zero-filled\"}\n",
                    (double)acc_timed / sec, 1e3 * sec / steps, steps, warmup, ns, ns * 40 + 24 * 8, ns * 40 + ns * 24 * 40, checksum);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "hlm_bench_shim: %s\n", e.what());
        return 1;
    }
}
