// routed_nccl.cpp — a routed run over several GPUs from a C++ host: one process per GPU, the partition of
// hlm_routing.hpp, and NCCL for the only inter-rank traffic (the boundary links' discharge, one ncclAllGather
// per coupling interval on the stream the library's kernels run on).  This is the host side INTEGRATION.md §6
// describes; the reference's own multi-process plumbing is MPI (main.cpp:269-309), which this image lacks, so
// ranks are plain processes launched with RANK / WORLD_SIZE / LOCAL_RANK in the environment (torchrun or a
// shell loop) and the NCCL unique id travels through a file in OUTDIR (hlm_nccl.hpp).
//
// usage: RANK=r WORLD_SIZE=n LOCAL_RANK=r hlm_routed_nccl PARAMS.csv OUTDIR [hours=3] [couple_min=15] [subbasin_links=4096] [rain] [temp]
// Writes OUTDIR/final_rank_<r>.csv: "stream,q,h_stat,h_surf,h_grav,h_aq" for the rank's links (17 significant digits).
#include <cstdio>
#include <cstdlib>
#include <fstream>

#include "hlm_host.hpp"
#include "hlm_nccl.hpp"
#include "hlm_routing.hpp"

namespace {
int env_int(const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}
}  // namespace

int main(int argc, char** argv) {
    if (argc < 3) {
        std::fprintf(stderr, "usage: RANK=r WORLD_SIZE=n LOCAL_RANK=r %s PARAMS.csv OUTDIR [hours] [couple_min] [subbasin_links] [rain] [temp]\n", argv[0]);
        return 2;
    }
    try {
        const int rank = env_int("RANK", 0), world = env_int("WORLD_SIZE", 1), device = env_int("LOCAL_RANK", rank);
        const std::string csv = argv[1], outdir = argv[2];
        const double hours = argc > 3 ? std::atof(argv[3]) : 3.0;
        const double dt = argc > 4 ? std::atof(argv[4]) : 15.0;
        const long long sub = argc > 5 ? std::atoll(argv[5]) : 4096;
        const float rain = argc > 6 ? (float)std::atof(argv[6]) : 2.0e-5f;
        const float temp = argc > 7 ? (float)std::atof(argv[7]) : 8.0f;

        // every rank reads the parameter file and computes the same plan, then keeps its own links
        const std::vector<SpatialParams> all = loadSpatialParams(csv);
        std::vector<long long> stream_id(all.size()), next_id(all.size());
        for (size_t i = 0; i < all.size(); ++i) {
            stream_id[i] = all[i].stream;
            next_id[i] = all[i].next_stream;
        }
        const hlm_b200::RoutePlan plan = hlm_b200::plan_routes(stream_id, next_id, world, sub);
        const hlm_b200::RankTopology& mine = plan.ranks[(size_t)rank];
        const long long ns = mine.n_local();
        std::vector<SpatialParams> sp((size_t)ns);
        for (long long k = 0; k < ns; ++k) sp[(size_t)k] = all[(size_t)plan.order[(size_t)(mine.lo + k)]];

        // stream, communicator (id through OUTDIR, hlm_nccl.hpp) and the two exchange buffers
        hlm_b200::NcclBoundaryExchange ex(world, rank, device, outdir, plan.max_send, plan.halo_len());
        cudaStream_t stream = ex.stream();

        hlm_b200::Context ctx(device);
        hlm_b200::check(hlm_set_stream(ctx.get(), stream), "hlm_set_stream");
        const long long nT_pr = (long long)(hours + 1.5), nT_t2m = (long long)(hours / 24.0 + 1.5);
        std::vector<float> pr((size_t)nT_pr * ns, rain), t2m((size_t)nT_t2m * ns, temp);
        ctx.setForcing(0, 1.0, nT_pr, ns, pr.data());
        ctx.setForcing(1, 24.0, nT_t2m, ns, t2m.data());
        ctx.setForcingColumns(nullptr, 0);
        Model200::Parameters hp;
        hp.initialStep = 1e-6;
        ctx.setModelParameters(Model200::UID, hp);
        ctx.setSpatialParams(sp.data(), ns);
        const double y0_common[5] = {0.5, 3.0, 0.0, 5.0, 0.2};
        std::vector<double> y0((size_t)ns * 5);
        for (long long s = 0; s < ns; ++s)
            for (int i = 0; i < 5; ++i) y0[(size_t)s * 5 + i] = y0_common[i];

        std::vector<double> fin((size_t)ns * 5);
        std::vector<int> code((size_t)ns);
        {
            hlm_b200::RoutedRun run(ctx, Model200::UID, mine, world, plan.max_send,
                                    [&](const double*, long long, double*) { ex.all_gather(); }, ex.d_send(), ex.d_halo());
            const long long n_int = (long long)(hours * 60.0 / dt + 0.5);
            for (long long k = 0; k < n_int; ++k) {
                const double tf = dt * (double)(k + 1);
                if (k == 0) run.begin(y0, ns, 0.0, tf, {});
                else run.advance(tf, {});
            }
            hlm_b200::check(hlm_solve_end(ctx.get(), fin.data(), code.data(), nullptr, nullptr, nullptr), "hlm_solve_end");
        }
        long long lost = 0;
        for (int c : code) lost += c == HLM_LINK_STIFF || c == HLM_LINK_STALLED;
        std::ofstream f(outdir + "/final_rank_" + std::to_string(rank) + ".csv");
        f.precision(17);
        f << "stream,q,h_stat,h_surf,h_grav,h_aq\n";
        for (long long s = 0; s < ns; ++s) {
            f << sp[(size_t)s].stream;
            for (int i = 0; i < 5; ++i) f << "," << fin[(size_t)s * 5 + i];
            f << "\n";
        }
        std::printf("[rank %d/%d] %lld links, %zu boundary links, %lld all-gathers of %lld doubles, %lld links lost\n", rank, world, ns,
                    mine.send_idx.size(), ex.exchanges(), plan.halo_len(), lost);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
