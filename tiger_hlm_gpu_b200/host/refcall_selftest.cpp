// refcall_selftest.cpp — the reference's own call sequence around the operator, compiled against the shim header.
//
// What main.cpp does between "h_y0 and h_query_times are filled" and "h_y_final and h_dense are back"
// (main.cpp:666-673 setup_gpu_buffers, :694-703 the launch — through launch_rk45_kernel as
// solver/rk45_api.hpp:120-154 documents it, since the kernel itself lives in the library — and :727-732
// retrieve_and_free), with the reference's names, tuple shape and argument order, followed by the composed
// run_rk45 with the reference's signature (solver/rk45_api.hpp:273-313) on a DEVICE copy of the parameter
// array, as main.cpp:392-404 makes it.  Both ways must give the same bits.
//
//   hlm_refcall_selftest PARAMS.csv   -> prints "refcall ok ..." and returns 0
#include <cstdio>
#include <cstring>
#include <tuple>
#include <vector>

#include "../../include/hlm_b200/rk45_api.hpp"
#include "hlm_host.hpp"

// the CUDA runtime calls main.cpp:392-404 makes for d_sp, without its headers (the host program links libcudart)
extern "C" int cudaMalloc(void** p, size_t n);
extern "C" int cudaMemcpy(void* dst, const void* src, size_t n, int kind);
extern "C" int cudaFree(void* p);

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s PARAMS.csv\n", argv[0]);
        return 2;
    }
    try {
        using rk45_api::launch_rk45_kernel;
        using rk45_api::retrieve_and_free;
        using rk45_api::setup_gpu_buffers;
        std::vector<SpatialParams> spatialParams = loadSpatialParams(argv[1]);
        const int num_systems = (int)spatialParams.size();
        // main.cpp:392-404: the parameter array in device memory
        SpatialParams* d_sp = nullptr;
        if (cudaMalloc((void**)&d_sp, sizeof(SpatialParams) * spatialParams.size()) != 0) throw std::runtime_error("cudaMalloc d_sp");
        if (cudaMemcpy(d_sp, spatialParams.data(), sizeof(SpatialParams) * spatialParams.size(), 1 /*HostToDevice*/) != 0)
            throw std::runtime_error("cudaMemcpy d_sp");
        // constant stub forcings as the committed goldens used (model_204.hpp:76-77)
        std::vector<float> pr(48 * (size_t)num_systems, 0.001f), t2m(2 * (size_t)num_systems, 1.0f);
        rk45_api::setForcing(0, 1.0, 48, num_systems, pr.data());
        rk45_api::setForcing(1, 24.0, 2, num_systems, t2m.data());
        Model204::Parameters hp;
        hp.initialStep = 1e-6;  // main.cpp:633-640
        rk45_api::setModelParameters<Model204>(hp);
        const double t0 = 0.0, tf = 2880.0;  // main.cpp:610-611
        std::vector<double> h_y0;
        for (int s = 0; s < num_systems; ++s)
            for (double v : {0.01, 3.0, 0.0, 5.0, 0.2}) h_y0.push_back(v);  // main.cpp:376,644-649
        std::vector<double> h_query_times;
        for (double t = t0; t <= tf; t += 60.0) h_query_times.push_back(t);  // main.cpp:653-657

        // ---- the sequence of main.cpp:666-732 ----
        double *d_y0_all, *d_y_final_all, *d_query_times, *d_dense_all;
        int* d_stiff;
        int ns, nq;
        std::tie(d_y0_all, d_y_final_all, d_query_times, d_dense_all, d_stiff, ns, nq) =
            setup_gpu_buffers<Model204>(h_y0, h_query_times);
        launch_rk45_kernel<Model204>(d_y0_all, d_y_final_all, d_query_times, d_dense_all, d_stiff, ns, nq, t0, tf, d_sp);
        auto [h_y_final, h_dense] =
            retrieve_and_free<Model204>(d_y0_all, d_y_final_all, d_query_times, d_dense_all, d_stiff, ns, nq, t0, tf, d_sp);

        // ---- the composed operator, reference signature ----
        auto [f2, d2] = rk45_api::run_rk45<Model204>(h_y0, t0, tf, h_query_times, d_sp);
        cudaFree(d_sp);

        if (ns != num_systems || nq != (int)h_query_times.size()) throw std::runtime_error("sizes");
        if (h_y_final.size() != f2.size() || h_dense.size() != d2.size()) throw std::runtime_error("result sizes differ");
        if (std::memcmp(h_y_final.data(), f2.data(), f2.size() * sizeof(double)) != 0) throw std::runtime_error("final states differ");
        if (std::memcmp(h_dense.data(), d2.data(), d2.size() * sizeof(double)) != 0) throw std::runtime_error("dense states differ");
        std::printf("refcall ok: %d systems, %d queries, final[0] = %.9g %.9g %.9g %.9g %.9g\n", ns, nq, h_y_final[0], h_y_final[1],
                    h_y_final[2], h_y_final[3], h_y_final[4]);
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "refcall_selftest: %s\n", e.what());
        return 1;
    }
}
