// hlm_host.hpp — C++ host side around the operator: the input and output contracts of the path.
//
// Mirrors, with the same names, argument meaning and error behaviour:
//   loadSpatialParams(csv)            I_O/parameters_loader.cpp:8-107  (throws std::runtime_error)
//   LookupMapper                      I_O/forcing_loader.hpp:13-32, forcing_loader.cpp:13-64
//   write_final_csv / write_dense_csv the CSV writers of main.cpp:734-773 (examples: src/final_204_a.csv,
//                                     src/dense_204_a.csv)
//   expandForcingColumns              the streamPoint computation of main.cpp:495-505
// Header-only, no CUDA and no third-party dependency; links against nothing but libhlm_b200.so
// through include/hlm_b200/rk45_api.hpp.
#pragma once

#include <fstream>
#include <iomanip>
#include <cstdio>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/hlm_b200/rk45_api.hpp"

// ---- I_O/parameters_loader.cpp ---------------------------------------------------------------------
inline std::vector<SpatialParams> loadSpatialParams(const std::string& csv_path) {
    std::ifstream in(csv_path);
    if (!in.is_open()) throw std::runtime_error("Failed to open parameter file: " + csv_path);
    std::string line;
    if (!std::getline(in, line)) throw std::runtime_error("Empty parameter file: " + csv_path);
    auto split = [](const std::string& s) {
        std::vector<std::string> out;
        std::string cell;
        std::istringstream ss(s);
        while (std::getline(ss, cell, ',')) out.push_back(cell);
        return out;
    };
    if (!line.empty() && line.back() == '\r') line.pop_back();
    const std::vector<std::string> headers = split(line);
    std::unordered_map<std::string, int> idx;
    for (int i = 0; i < (int)headers.size(); ++i) idx[headers[i]] = i;
    static const char* required[] = {"stream", "next_stream", "i2", "i3", "hu", "centroid_lat", "sw", "ss", "n", "slope",
                                     "length_km", "drainage_area_km2", "melt", "t_thres", "res_ss", "res_gw"};
    for (const char* name : required)
        if (idx.find(name) == idx.end()) throw std::runtime_error(std::string("Missing column '") + name + "' in " + csv_path);

    constexpr double c_1 = 0.001 / 60.0;  // mm/hr -> m/min, parameters_loader.cpp:57
    std::vector<SpatialParams> out;
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        const std::vector<std::string> f = split(line);
        if (f.size() < headers.size()) throw std::runtime_error("Bad row with too few fields in " + csv_path);
        SpatialParams p{};
        p.stream = std::stol(f[idx["stream"]]);
        p.next_stream = std::stol(f[idx["next_stream"]]);
        p.c1 = c_1;
        p.Hu = std::stod(f[idx["hu"]]);
        p.lat = std::stod(f[idx["centroid_lat"]]);
        p.sw = std::stod(f[idx["sw"]]);
        p.ss = std::stod(f[idx["ss"]]);
        p.n_mann = std::stod(f[idx["n"]]);
        p.slope = std::stod(f[idx["slope"]]);
        p.L = std::stod(f[idx["length_km"]]);
        p.A_h = std::stod(f[idx["drainage_area_km2"]]);
        p.melt_f = std::stod(f[idx["melt"]]);
        p.temp_thr = std::stod(f[idx["t_thres"]]);
        p.infil = std::stod(f[idx["i2"]]) * c_1;
        p.perco = std::stod(f[idx["i3"]]) * c_1;
        p.alpha3 = std::stod(f[idx["res_ss"]]) * 24.0 * 60.0;
        p.alpha4 = std::stod(f[idx["res_gw"]]) * 24.0 * 60.0;
        out.push_back(p);
    }
    return out;
}

// ---- I_O/forcing_loader.cpp:13-64 -------------------------------------------------------------------
class LookupMapper {
  public:
    explicit LookupMapper(const std::string& filepath) : filepath_(filepath) {}
    bool load() {
        std::ifstream file(filepath_);
        if (!file.is_open()) {
            std::fprintf(stderr, "Failed to open file: %s\n", filepath_.c_str());
            return false;
        }
        std::string line;
        std::getline(file, line);  // header
        while (std::getline(file, line)) {
            if (line.empty()) continue;
            std::istringstream ss(line);
            std::string field;
            std::getline(ss, field, ',');
            const long long stream = std::stoll(field);
            std::getline(ss, field, ',');
            const int lat = std::stoi(field);
            std::getline(ss, field, ',');
            const int lon = std::stoi(field);
            stream_map_[stream] = {lat, lon};
        }
        return true;
    }
    bool hasStream(long long id) const { return stream_map_.find(id) != stream_map_.end(); }
    std::pair<int, int> getLatLon(long long id) const {
        auto it = stream_map_.find(id);
        return it != stream_map_.end() ? it->second : std::pair<int, int>{-1, -1};
    }
    size_t size() const { return stream_map_.size(); }

  private:
    std::string filepath_;
    std::unordered_map<long long, std::pair<int, int>> stream_map_;
};

/// Forcing column (grid cell) of every link: lat*lon_size + lon, main.cpp:501-505.  Unlike the
/// reference (which silently indexes with -1*lon_size-1) an unmapped stream is an error.
inline std::vector<int> forcingColumns(const std::vector<SpatialParams>& sp, const LookupMapper& lm, int lon_size) {
    std::vector<int> col(sp.size());
    for (size_t s = 0; s < sp.size(); ++s) {
        if (!lm.hasStream(sp[s].stream))
            throw std::runtime_error("stream " + std::to_string(sp[s].stream) + " is not in the forcing lookup");
        const auto ll = lm.getLatLon(sp[s].stream);
        col[s] = ll.first * lon_size + ll.second;
    }
    return col;
}

// ---- main.cpp:734-773 ----------------------------------------------------------------------------------
/// final.csv: header "h_snow,var1,...", one row per system (default ostream precision, as the reference).
template <class Vec>  // std::vector<double> or rk45_api::FinalType
inline void write_final_csv(const std::string& path, const Vec& y_final, int num_systems, int n_eq) {
    std::ofstream f(path);
    if (!f.is_open()) throw std::runtime_error("cannot open " + path);
    f << "h_snow";
    for (int i = 1; i < n_eq; ++i) f << ",var" << i;
    f << "\n";
    for (int s = 0; s < num_systems; ++s) {
        for (int i = 0; i < n_eq; ++i) {
            f << y_final[(size_t)s * n_eq + i];
            if (i + 1 < n_eq) f << ",";
        }
        f << "\n";
    }
}

/// dense.csv: header "time,var{i}_sys{s}...", time with setprecision(8) fixed, values setprecision(9).
template <class Vec>  // std::vector<double> or rk45_api::DenseType
inline void write_dense_csv(const std::string& path, const Vec& dense, const std::vector<double>& tq,
                            int num_systems, int n_eq) {
    std::ofstream f(path);
    if (!f.is_open()) throw std::runtime_error("cannot open " + path);
    const int nq = (int)tq.size();
    f << "time";
    for (int s = 0; s < num_systems; ++s)
        for (int i = 0; i < n_eq; ++i) f << ",var" << i << "_sys" << s;
    f << "\n";
    for (int q = 0; q < nq; ++q) {
        f << std::fixed << std::setprecision(8) << tq[q];
        for (int s = 0; s < num_systems; ++s)
            for (int i = 0; i < n_eq; ++i)
                f << "," << std::setprecision(9) << dense[((size_t)s * nq + q) * n_eq + i];
        f << "\n";
    }
}
