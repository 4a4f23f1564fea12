// routed_example.cpp — a routed run from a parameter CSV: Model 200 (hillslope-link runoff) with links
// coupled through upstream discharge along `stream`/`next_stream`, constant forcing, one dense record per
// coupling interval.  Writes final.csv / dense.csv in the reference's formats (main.cpp:734-773) and prints
// what a partition over WORLD ranks would exchange.  Used by tests/test_gpu_host_cpp.py.
//
// usage: hlm_routed_example PARAMS.csv OUTDIR [hours=6] [couple_min=15] [subbasin_links=4096] [world=1] [rain] [temp]
#include <cstdio>
#include <cstdlib>

#include "hlm_host.hpp"
#include "hlm_routing.hpp"

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s PARAMS.csv OUTDIR [hours] [couple_min] [subbasin_links] [world] [rain] [temp]\n", argv[0]);
        return 2;
    }
    try {
        const std::string csv = argv[1], outdir = argv[2];
        const double hours = argc > 3 ? std::atof(argv[3]) : 6.0;
        const double dt = argc > 4 ? std::atof(argv[4]) : 15.0;
        const long long sub = argc > 5 ? std::atoll(argv[5]) : 4096;
        const int world = argc > 6 ? std::atoi(argv[6]) : 1;
        const float rain = argc > 7 ? (float)std::atof(argv[7]) : 2.0e-5f;
        const float temp = argc > 8 ? (float)std::atof(argv[8]) : 8.0f;

        std::vector<SpatialParams> sp = loadSpatialParams(csv);
        const long long ns = (long long)sp.size();
        std::vector<long long> stream((size_t)ns), next((size_t)ns);
        for (long long i = 0; i < ns; ++i) {
            stream[(size_t)i] = sp[(size_t)i].stream;
            next[(size_t)i] = sp[(size_t)i].next_stream;
        }
        const hlm_b200::RoutePlan parts = hlm_b200::plan_routes(stream, next, world, sub);
        std::printf("partition over %d rank(s): %lld sub-basins, %lld cut edges, halo vector of %lld doubles per interval\n", world,
                    parts.n_subbasins, parts.n_cut_edges, parts.halo_len());
        for (const auto& t : parts.ranks)
            std::printf("  rank %d: links [%lld, %lld), %zu boundary links\n", t.rank, t.lo, t.hi, t.send_idx.size());

        // the run itself on this process's GPU: one rank owning every link
        const hlm_b200::RoutePlan one = hlm_b200::plan_routes(stream, next, 1, sub);
        hlm_b200::Context& ctx = hlm_b200::default_context();
        const long long nT_pr = (long long)(hours + 1.5), nT_t2m = (long long)(hours / 24.0 + 1.5);
        std::vector<float> pr((size_t)nT_pr * ns, rain), t2m((size_t)nT_t2m * ns, temp);
        ctx.setForcing(0, 1.0, nT_pr, ns, pr.data());
        ctx.setForcing(1, 24.0, nT_t2m, ns, t2m.data());
        ctx.setForcingColumns(nullptr, 0);
        Model200::Parameters hp;
        hp.initialStep = 1e-6;
        ctx.setModelParameters(Model200::UID, hp);
        ctx.setSpatialParams(sp.data(), ns);  // plan `one` keeps the original order
        const double y0_common[5] = {0.5, 3.0, 0.0, 5.0, 0.2};
        std::vector<double> y0((size_t)ns * 5);
        for (long long s = 0; s < ns; ++s)
            for (int i = 0; i < 5; ++i) y0[(size_t)s * 5 + i] = y0_common[i];

        const long long n_int = (long long)(hours * 60.0 / dt + 0.5);
        std::vector<double> tq_all, dense((size_t)ns * (size_t)n_int * 5, 0.0), win((size_t)ns * 5);
        hlm_b200::RoutedRun run(ctx, Model200::UID, one.ranks[0], 1, 0);
        for (long long k = 0; k < n_int; ++k) {
            const double tf = dt * (double)(k + 1);
            const std::vector<double> tq{tf};
            if (k == 0) run.begin(y0, ns, 0.0, tf, tq);
            else run.advance(tf, tq);
            int ticket = -1;
            hlm_b200::check(hlm_solve_fetch_window_packed(ctx.get(), win.data(), &ticket), "hlm_solve_fetch_window_packed");
            hlm_b200::check(hlm_solve_wait_copy(ctx.get(), ticket), "hlm_solve_wait_copy");
            for (long long s = 0; s < ns; ++s)
                for (int i = 0; i < 5; ++i) dense[((size_t)s * (size_t)n_int + (size_t)k) * 5 + i] = win[(size_t)s * 5 + i];
            tq_all.push_back(tf);
        }
        std::vector<double> fin((size_t)ns * 5);
        std::vector<int> code((size_t)ns);
        hlm_b200::check(hlm_solve_end(ctx.get(), fin.data(), code.data(), nullptr, nullptr, nullptr), "hlm_solve_end");
        long long solved = 0, lost = 0;
        for (int c : code) {
            solved += c == HLM_LINK_STIFF_SOLVED;
            lost += c == HLM_LINK_STIFF || c == HLM_LINK_STALLED;
        }
        std::printf("%lld links, %lld intervals of %.1f min; %lld finished by the implicit fallback, %lld lost\n", ns, n_int, dt, solved, lost);
        write_final_csv(outdir + "/final.csv", fin, (int)ns, 5);
        write_dense_csv(outdir + "/dense.csv", dense, tq_all, (int)ns, 5);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
