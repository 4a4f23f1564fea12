// host_selftest.cpp — CPU-only checks of the host contracts (no CUDA call): parse a parameter CSV and a
// lookup CSV and print what was read, so tests can compare with the Python mirror and with the
// reference loader's documented behaviour.  usage: hlm_host_selftest params PARAMS.csv | lookup LOOKUP.csv ID...
#include <cstdio>
#include <cstring>

#include "hlm_host.hpp"

int main(int argc, char** argv) {
    try {
        if (argc >= 3 && !std::strcmp(argv[1], "params")) {
            auto sp = loadSpatialParams(argv[2]);
            std::printf("%zu\n", sp.size());
            for (const auto& p : sp)
                std::printf("%ld %ld %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n",
                            p.stream, p.next_stream, p.c1, p.infil, p.perco, p.Hu, p.lat, p.sw, p.ss, p.n_mann, p.slope,
                            p.L, p.A_h, p.alpha3, p.alpha4, p.melt_f, p.temp_thr);
            return 0;
        }
        if (argc >= 3 && !std::strcmp(argv[1], "lookup")) {
            LookupMapper lm(argv[2]);
            if (!lm.load()) return 3;
            std::printf("%zu\n", lm.size());
            for (int i = 3; i < argc; ++i) {
                auto ll = lm.getLatLon(std::atoll(argv[i]));
                std::printf("%s %d %d %d\n", argv[i], (int)lm.hasStream(std::atoll(argv[i])), ll.first, ll.second);
            }
            return 0;
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    std::fprintf(stderr, "usage: %s params FILE | lookup FILE ID...\n", argv[0]);
    return 2;
}
