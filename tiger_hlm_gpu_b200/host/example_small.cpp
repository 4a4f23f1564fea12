// example_small.cpp — the reference's small-example driver (main.cpp:376-382,610-657,734-773) against
// the drop-in operator: load a parameter CSV, constant or synthetic forcing, run_rk45<Model204>,
// write final.csv / dense.csv in the reference's formats.  Used by tests/test_gpu_host_cpp.py.
//
// usage: hlm_example PARAMS.csv OUTDIR [days=2] [query_step_minutes=60] [rain=0.001] [temp=1.0]
#include <cstdio>
#include <cstdlib>

#include "hlm_host.hpp"

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: %s PARAMS.csv OUTDIR [days] [query_step_min] [rain] [temp]\n", argv[0]);
        return 2;
    }
    try {
        const std::string csv = argv[1], outdir = argv[2];
        const double days = argc > 3 ? std::atof(argv[3]) : 2.0;
        const double qstep = argc > 4 ? std::atof(argv[4]) : 60.0;
        const float rain = argc > 5 ? (float)std::atof(argv[5]) : 0.001f;
        const float temp = argc > 6 ? (float)std::atof(argv[6]) : 1.0f;

        std::vector<SpatialParams> sp = loadSpatialParams(csv);  // main.cpp:272
        const int ns = (int)sp.size();
        const double t0 = 0.0, tf = days * 24.0 * 60.0;  // main.cpp:610-611

        // forcing: pr hourly, t2m daily, per-link layout [time][system] (main.cpp:543-548)
        const long long nT_pr = (long long)(days * 24.0 + 0.5), nT_t2m = (long long)(days + 0.5);
        std::vector<float> pr((size_t)nT_pr * ns, rain), t2m((size_t)nT_t2m * ns, temp);
        rk45_api::setForcing(0, 1.0, nT_pr, ns, pr.data());
        rk45_api::setForcing(1, 24.0, nT_t2m, ns, t2m.data());

        Model204::Parameters hp;  // main.cpp:633-640 (initialStep is always 1e-6 there)
        hp.initialStep = 1e-6;
        rk45_api::setModelParameters<Model204>(hp);

        const double y0_common[5] = {0.01, 3.0, 0.0, 5.0, 0.2};  // main.cpp:376
        std::vector<double> h_y0((size_t)ns * Model204::N_EQ);
        for (int s = 0; s < ns; ++s)
            for (int i = 0; i < Model204::N_EQ; ++i) h_y0[(size_t)s * Model204::N_EQ + i] = y0_common[i];
        std::vector<double> tq;
        for (double t = t0; t <= tf; t += qstep) tq.push_back(t);  // main.cpp:653-657

        auto res = rk45_api::run_rk45<Model204>(h_y0, t0, tf, tq, hlm_b200::SpView{sp.data(), ns});
        write_final_csv(outdir + "/final.csv", res.first, ns, Model204::N_EQ);
        write_dense_csv(outdir + "/dense.csv", res.second, tq, ns, Model204::N_EQ);
        std::printf("Final states at t = %.1f:\n", tf);
        for (int s = 0; s < ns && s < 3; ++s) {
            std::printf(" System %d:", s);
            for (int i = 0; i < Model204::N_EQ; ++i) std::printf(" y%d=%.6f", i, res.first[(size_t)s * 5 + i]);
            std::printf("\n");
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
