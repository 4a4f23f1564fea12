// hlm_config.hpp — the run configuration: data/config.yaml of the reference, same schema.
//
// Mirrors I_O/config_loader.hpp:10-57 (struct SimulationConfig) and I_O/config_loader.cpp:19-84
// (load_config): the same sections, key names, required/optional keys and the same exception type
// (std::runtime_error) on a missing key or an unparsable time.  yaml-cpp is not part of this image,
// and the schema needs only block mappings, block sequences, scalars and comments, so that subset of
// YAML is parsed here (flow sequences "[0, 1]" are accepted too).  Sections the reference parses but
// never uses (global_params, local_params.columns, mpi, flags) are carried in the struct as well.
//
// Extensions, all optional so that the reference's own config.yaml loads unchanged:
//   time.origin: ISO8601       instant of t = 0 and of forcing sample 0 (default: time.start); a run
//                              restarted from a final-state file keeps the origin and moves time.start
//   forcings.dt_hours: {precipitation: 1, temperature: 24}   sample spacing when the files do not say
//   output.dir / output.prefix / output.format (netcdf = NetCDF-4 as the reference writes | netcdf3 | csv) /
//   output.compression_level (0-9, default 4: output_series.hpp) / output.dense (bool) / output.precision (64|32)
//   solver.interval: "1d"      the run is driven in intervals of this length
//   solver.max_attempts        per-link attempt budget per window (0 = unbounded like the reference)
//   solver.stiff_fallback: true   links the RK45 path flags stiff are continued by the Radau IIA fallback
//   routing: {enabled: true, couple: "15m"}   links coupled through next_stream (Model 200): the run advances
//                              in coupling intervals, upstream discharge held over each
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace hlmyaml {

struct Node {
    enum Kind { kNull, kScalar, kMap, kSeq } kind = kNull;
    std::string scalar;
    bool quoted = false;
    std::vector<std::pair<std::string, Node>> map;
    std::vector<Node> seq;

    bool defined() const { return kind != kNull || defined_null; }
    bool IsNull() const { return kind == kNull; }
    bool defined_null = false;  // key present with an explicit null / empty value

    const Node& operator[](const std::string& key) const {
        static const Node missing;
        if (kind != kMap) return missing;
        for (auto& kv : map) if (kv.first == key) return kv.second;
        return missing;
    }
    explicit operator bool() const { return defined(); }

    std::string as_string(const std::string& what) const {
        if (kind != kScalar) throw std::runtime_error("config: key '" + what + "' is missing or not a scalar");
        return scalar;
    }
    std::string as_string_or(const std::string& dflt) const { return kind == kScalar ? scalar : dflt; }
    double as_double(const std::string& what) const {
        const std::string s = as_string(what);
        char* end = nullptr;
        const double v = std::strtod(s.c_str(), &end);
        if (end == s.c_str() || *end != '\0') throw std::runtime_error("config: '" + what + "' is not a number: " + s);
        return v;
    }
    long long as_int(const std::string& what) const {
        const std::string s = as_string(what);
        char* end = nullptr;
        const long long v = std::strtoll(s.c_str(), &end, 10);
        if (end == s.c_str() || *end != '\0') throw std::runtime_error("config: '" + what + "' is not an integer: " + s);
        return v;
    }
    bool as_bool(const std::string& what) const {
        const std::string s = as_string(what);
        if (s == "true" || s == "True" || s == "yes" || s == "on") return true;
        if (s == "false" || s == "False" || s == "no" || s == "off") return false;
        throw std::runtime_error("config: '" + what + "' is not a boolean: " + s);
    }
};

namespace detail {

struct Line { int indent; std::string text; int number; };

inline std::string strip_comment(const std::string& s) {
    bool sq = false, dq = false;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (c == '\'' && !dq) sq = !sq;
        else if (c == '"' && !sq) dq = !dq;
        else if (c == '#' && !sq && !dq && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
    }
    return s;
}
inline std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) ++a;
    while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
    return s.substr(a, b - a);
}
inline Node scalar_node(const std::string& raw) {
    Node n;
    std::string s = trim(raw);
    if (s.empty() || s == "~" || s == "null" || s == "Null" || s == "NULL") {
        n.defined_null = true;
        return n;
    }
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) {
        n.kind = Node::kScalar;
        n.quoted = true;
        n.scalar = s.substr(1, s.size() - 2);
        return n;
    }
    if (s.front() == '[' && s.back() == ']') {  // flow sequence of scalars
        n.kind = Node::kSeq;
        std::string item;
        std::istringstream ss(s.substr(1, s.size() - 2));
        while (std::getline(ss, item, ',')) if (!trim(item).empty()) n.seq.push_back(scalar_node(item));
        return n;
    }
    n.kind = Node::kScalar;
    n.scalar = s;
    return n;
}
// position of the ':' that ends a mapping key ("key: value" or "key:"), or npos
inline size_t key_colon(const std::string& s) {
    bool sq = false, dq = false;
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[i];
        if (c == '\'' && !dq) sq = !sq;
        else if (c == '"' && !sq) dq = !dq;
        else if (c == ':' && !sq && !dq && (i + 1 == s.size() || s[i + 1] == ' ' || s[i + 1] == '\t')) return i;
    }
    return std::string::npos;
}

class Parser {
  public:
    explicit Parser(std::vector<Line> lines) : lines_(std::move(lines)) {}
    Node parse() {
        Node n = block(lines_.empty() ? 0 : lines_[0].indent);
        if (pos_ != lines_.size()) fail("unexpected indentation");
        return n;
    }

  private:
    [[noreturn]] void fail(const std::string& msg) const {
        const int ln = pos_ < lines_.size() ? lines_[pos_].number : -1;
        throw std::runtime_error("config: YAML " + msg + " at line " + std::to_string(ln));
    }
    Node block(int indent) {
        Node n;
        if (pos_ >= lines_.size()) return n;
        if (lines_[pos_].text.compare(0, 2, "- ") == 0 || lines_[pos_].text == "-") {
            n.kind = Node::kSeq;
            while (pos_ < lines_.size() && lines_[pos_].indent == indent &&
                   (lines_[pos_].text.compare(0, 2, "- ") == 0 || lines_[pos_].text == "-")) {
                Line& l = lines_[pos_];
                const std::string rest = l.text.size() > 2 ? trim(l.text.substr(2)) : "";
                if (rest.empty()) {
                    ++pos_;
                    n.seq.push_back(pos_ < lines_.size() && lines_[pos_].indent > indent ? block(lines_[pos_].indent) : Node());
                } else if (key_colon(rest) != std::string::npos && rest.front() != '"' && rest.front() != '\'') {
                    // "- key: value": a mapping whose first key sits on the dash line
                    const int inner = indent + (int)l.text.find_first_not_of(" \t", 1);
                    l.indent = inner;
                    l.text = rest;
                    n.seq.push_back(block(inner));
                } else {
                    n.seq.push_back(scalar_node(rest));
                    ++pos_;
                }
            }
            return n;
        }
        n.kind = Node::kMap;
        while (pos_ < lines_.size() && lines_[pos_].indent == indent) {
            const Line& l = lines_[pos_];
            if (l.text.compare(0, 2, "- ") == 0) break;
            const size_t c = key_colon(l.text);
            if (c == std::string::npos) fail("expected 'key: value'");
            std::string key = trim(l.text.substr(0, c));
            if (key.size() >= 2 && (key.front() == '"' || key.front() == '\'')) key = key.substr(1, key.size() - 2);
            const std::string rest = trim(l.text.substr(c + 1));
            ++pos_;
            if (!rest.empty()) {
                n.map.emplace_back(key, scalar_node(rest));
            } else if (pos_ < lines_.size() && lines_[pos_].indent > indent) {
                n.map.emplace_back(key, block(lines_[pos_].indent));
            } else if (pos_ < lines_.size() && lines_[pos_].indent == indent && lines_[pos_].text.compare(0, 2, "- ") == 0) {
                n.map.emplace_back(key, block(indent));  // sequence at the same indentation as its key
            } else {
                Node nul;
                nul.defined_null = true;
                n.map.emplace_back(key, nul);
            }
        }
        if (pos_ < lines_.size() && lines_[pos_].indent > indent) fail("unexpected indentation");
        return n;
    }
    std::vector<Line> lines_;
    size_t pos_ = 0;
};

}  // namespace detail

inline Node Load(std::istream& in) {
    std::vector<detail::Line> lines;
    std::string raw;
    int number = 0;
    while (std::getline(in, raw)) {
        ++number;
        std::string s = detail::strip_comment(raw);
        while (!s.empty() && (s.back() == ' ' || s.back() == '\t' || s.back() == '\r')) s.pop_back();
        if (s.empty() || s == "---") continue;
        int indent = 0;
        while (indent < (int)s.size() && s[indent] == ' ') ++indent;
        if (indent < (int)s.size() && s[indent] == '\t') throw std::runtime_error("config: YAML tab indentation at line " + std::to_string(number));
        lines.push_back({indent, s.substr(indent), number});
    }
    return detail::Parser(std::move(lines)).parse();
}
inline Node LoadFile(const std::string& filename) {
    std::ifstream in(filename);
    if (!in.is_open()) throw std::runtime_error("config: cannot open " + filename);
    return Load(in);
}

}  // namespace hlmyaml

// I_O/config_loader.hpp:10-57, plus the optional extensions listed at the top of this file.
struct SimulationConfig {
    struct ModelInfo { int uid = 0; std::string name; } model;
    struct TimeInfo {
        std::chrono::system_clock::time_point start, end, origin;
        std::string start_text, end_text, origin_text;
        double minutes() const { return std::chrono::duration<double>(end - start).count() / 60.0; }
        double start_minutes() const { return std::chrono::duration<double>(start - origin).count() / 60.0; }
        double end_minutes() const { return std::chrono::duration<double>(end - origin).count() / 60.0; }
    } time;
    struct InitialInfo { std::string mode; std::string file; } initial;
    struct GlobalParam { std::string name; double value = 0.0; };
    std::vector<GlobalParam> global_params;
    struct LocalParams { std::string file; int stream_id = 0, next_stream_id = 1, params_start = 2, num_params = 15; } local_params;
    struct ForcingInfo {
        std::string type, path, lookup_csv, var_precip, var_temp;
        double dt_precip_hours = 0.0, dt_temp_hours = 0.0;  // 0 = take from the file, else the reference's 1 h / 24 h
    } forcings;
    struct OutputInfo {
        std::string print_interval;
        std::vector<int> states;
        std::string dir = ".", prefix = "", format = "netcdf";
        bool dense = true;
        int precision = 64;  // dense records as double (the reference's type) or float
        int compression_level = 4;  // deflate level of `outputs` in NetCDF-4 files (output_series.hpp:24,44), 0 = none
    } output;
    struct SolverInfo {
        std::string method = "RK45";
        double rtol = 1e-6, atol = 1e-9, safety = 0.9, min_scale = 0.2, max_scale = 10.0;
        bool override_tolerances = false;
        double initial_step = 0.0;
        bool override_initial_step = false;
        std::string interval = "1d";
        long long max_attempts = 0;
        bool stiff_fallback = false;
    } solver;
    struct RoutingInfo { bool enabled = false; std::string couple = "15m"; long long subbasin_links = 4096; } routing;
    struct MPIInfo { int step_storage = 0, transfer_buffer = 0, discontinuity_buf = 0; } mpi;
    struct FlagsInfo { bool uses_dam = false, convert_area = false; } flags;
};

/// "15m", "1h", "1d", "90s" or a bare number of minutes -> minutes.
inline double parse_interval_minutes(const std::string& s) {
    if (s.empty()) throw std::runtime_error("config: empty interval");
    char* end = nullptr;
    const double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || !(v > 0.0)) throw std::runtime_error("config: bad interval '" + s + "'");
    std::string unit(end);
    while (!unit.empty() && unit.front() == ' ') unit.erase(unit.begin());
    if (unit.empty() || unit == "m" || unit == "min") return v;
    if (unit == "s") return v / 60.0;
    if (unit == "h") return v * 60.0;
    if (unit == "d") return v * 1440.0;
    throw std::runtime_error("config: bad interval unit in '" + s + "'");
}

/// End of the solver interval that starts at `ta`: the next multiple of `interval` (minutes from the origin)
/// strictly after ta, capped at t_end.  Multiples are counted with an integer — ta / interval can round just
/// below k once ta = k * interval when the interval is not a whole number of minutes (10s, 0.1m), and
/// floor(ta / interval) + 1 then returns ta itself for ever.  `k` carries the count between calls (start at 0).
inline double next_interval_boundary(double ta, double t_end, double interval, long long& k) {
    if (!(interval > 0.0)) throw std::runtime_error("config: interval must be positive");
    if (k <= 0) k = (long long)std::floor(ta / interval + 1e-9) + 1;
    while ((double)k * interval <= ta) ++k;
    const double tb = std::min(t_end, (double)k * interval);
    if (!(tb > ta)) throw std::runtime_error("config: interval boundary does not advance (interval too small for the time axis)");
    return tb;
}

namespace hlmcfg_detail {
// config_loader.cpp:11-17: "%Y-%m-%dT%H:%M:%S" through std::get_time.  The reference converts with
// mktime (local time zone); differences of two such points equal the UTC differences unless a DST
// change lies between them, so timegm is used here to keep run lengths independent of TZ.
inline std::chrono::system_clock::time_point parse_iso8601(const std::string& s) {
    std::tm tm = {};
    std::istringstream ss(s);
    ss >> std::get_time(&tm, "%Y-%m-%dT%H:%M:%S");
    if (ss.fail()) throw std::runtime_error("Failed to parse time: " + s);
    return std::chrono::system_clock::from_time_t(timegm(&tm));
}
}  // namespace hlmcfg_detail

/// I_O/config_loader.cpp:19-84.
inline SimulationConfig config_from_yaml(const hlmyaml::Node& doc) {
    using hlmyaml::Node;
    SimulationConfig cfg;
    auto need = [](const Node& n, const char* what) -> const Node& {
        if (!n.defined()) throw std::runtime_error(std::string("config: missing key '") + what + "'");
        return n;
    };
    // 1) model
    const Node& m = need(doc["model"], "model");
    cfg.model.uid = (int)need(m["uid"], "model.uid").as_int("model.uid");
    cfg.model.name = m["name"].as_string_or("");
    // 2) time
    const Node& t = need(doc["time"], "time");
    cfg.time.start_text = need(t["start"], "time.start").as_string("time.start");
    cfg.time.end_text = need(t["end"], "time.end").as_string("time.end");
    cfg.time.start = hlmcfg_detail::parse_iso8601(cfg.time.start_text);
    cfg.time.end = hlmcfg_detail::parse_iso8601(cfg.time.end_text);
    cfg.time.origin_text = t["origin"].as_string_or(cfg.time.start_text);
    cfg.time.origin = hlmcfg_detail::parse_iso8601(cfg.time.origin_text);
    // 3) initial
    const Node& init = need(doc["initial"], "initial");
    cfg.initial.mode = need(init["mode"], "initial.mode").as_string("initial.mode");
    if (cfg.initial.mode == "hot") cfg.initial.file = need(init["file"], "initial.file").as_string("initial.file");
    // 4) global parameters (parsed for completeness; no model here uses any)
    if (doc["global_params"].kind == Node::kSeq)
        for (auto& g : doc["global_params"].seq)
            if (g.kind == Node::kMap) cfg.global_params.push_back({g["name"].as_string_or(""), g["value"] ? g["value"].as_double("global_params.value") : 0.0});
    // 5) local parameters
    if (doc["local_params"]) {
        const Node& lp = doc["local_params"];
        cfg.local_params.file = lp["file"].as_string_or("");
        const Node& c = lp["columns"];
        if (c["stream_id"]) cfg.local_params.stream_id = (int)c["stream_id"].as_int("local_params.columns.stream_id");
        if (c["next_stream_id"]) cfg.local_params.next_stream_id = (int)c["next_stream_id"].as_int("local_params.columns.next_stream_id");
        if (c["params_start"]) cfg.local_params.params_start = (int)c["params_start"].as_int("local_params.columns.params_start");
        if (c["num_params"]) cfg.local_params.num_params = (int)c["num_params"].as_int("local_params.columns.num_params");
    }
    // 6) forcings
    const Node& f = need(doc["forcings"], "forcings");
    cfg.forcings.type = need(f["type"], "forcings.type").as_string("forcings.type");
    cfg.forcings.path = need(f["path"], "forcings.path").as_string("forcings.path");
    cfg.forcings.lookup_csv = need(f["lookup"], "forcings.lookup").as_string("forcings.lookup");
    const Node& vars = need(f["vars"], "forcings.vars");
    cfg.forcings.var_precip = need(vars["precipitation"], "forcings.vars.precipitation").as_string("forcings.vars.precipitation");
    cfg.forcings.var_temp = need(vars["temperature"], "forcings.vars.temperature").as_string("forcings.vars.temperature");
    if (f["dt_hours"]["precipitation"]) cfg.forcings.dt_precip_hours = f["dt_hours"]["precipitation"].as_double("forcings.dt_hours.precipitation");
    if (f["dt_hours"]["temperature"]) cfg.forcings.dt_temp_hours = f["dt_hours"]["temperature"].as_double("forcings.dt_hours.temperature");
    // 7) output
    const Node& o = need(doc["output"], "output");
    cfg.output.print_interval = need(o["print_interval"], "output.print_interval").as_string("output.print_interval");
    if (o["states"].kind == Node::kSeq)
        for (auto& v : o["states"].seq) cfg.output.states.push_back((int)v.as_int("output.states"));
    cfg.output.dir = o["dir"].as_string_or(".");
    cfg.output.prefix = o["prefix"].as_string_or("");
    cfg.output.format = o["format"].as_string_or("netcdf");
    if (cfg.output.format == "netcdf4") cfg.output.format = "netcdf";
    if (cfg.output.format == "classic") cfg.output.format = "netcdf3";
    if (cfg.output.format != "netcdf" && cfg.output.format != "netcdf3" && cfg.output.format != "csv")
        throw std::runtime_error("config: output.format must be netcdf, netcdf3 or csv");
    if (o["compression_level"]) {
        cfg.output.compression_level = (int)o["compression_level"].as_int("output.compression_level");
        if (cfg.output.compression_level < 0 || cfg.output.compression_level > 9) throw std::runtime_error("config: output.compression_level must be 0..9");
    }
    if (o["dense"]) cfg.output.dense = o["dense"].as_bool("output.dense");
    if (o["precision"]) {
        cfg.output.precision = (int)o["precision"].as_int("output.precision");
        if (cfg.output.precision != 32 && cfg.output.precision != 64) throw std::runtime_error("config: output.precision must be 32 or 64");
    }
    // 8) solver
    const Node& s = need(doc["solver"], "solver");
    cfg.solver.method = s["method"].as_string_or("RK45");
    const Node& tol = s["tolerances"];
    if (tol.kind == Node::kMap) {
        cfg.solver.override_tolerances = true;
        cfg.solver.rtol = need(tol["rtol"], "solver.tolerances.rtol").as_double("solver.tolerances.rtol");
        cfg.solver.atol = need(tol["atol"], "solver.tolerances.atol").as_double("solver.tolerances.atol");
        cfg.solver.safety = need(tol["safety"], "solver.tolerances.safety").as_double("solver.tolerances.safety");
        cfg.solver.min_scale = need(tol["min_scale"], "solver.tolerances.min_scale").as_double("solver.tolerances.min_scale");
        cfg.solver.max_scale = need(tol["max_scale"], "solver.tolerances.max_scale").as_double("solver.tolerances.max_scale");
    }
    if (s["initial_step"] && !s["initial_step"].IsNull()) {
        cfg.solver.override_initial_step = true;
        cfg.solver.initial_step = s["initial_step"].as_double("solver.initial_step");
    }
    cfg.solver.interval = s["interval"].as_string_or("1d");
    if (s["max_attempts"]) cfg.solver.max_attempts = s["max_attempts"].as_int("solver.max_attempts");
    if (s["stiff_fallback"]) cfg.solver.stiff_fallback = s["stiff_fallback"].as_bool("solver.stiff_fallback");
    if (doc["routing"]) {
        const Node& r = doc["routing"];
        if (r["enabled"]) cfg.routing.enabled = r["enabled"].as_bool("routing.enabled");
        cfg.routing.couple = r["couple"].as_string_or("15m");
        if (r["subbasin_links"]) cfg.routing.subbasin_links = r["subbasin_links"].as_int("routing.subbasin_links");
    }
    // 9) mpi (required by the reference's loader; optional here: no MPI on this path)
    if (doc["mpi"]) {
        const Node& mp = doc["mpi"];
        if (mp["step_storage"]) cfg.mpi.step_storage = (int)mp["step_storage"].as_int("mpi.step_storage");
        if (mp["transfer_buffer"]) cfg.mpi.transfer_buffer = (int)mp["transfer_buffer"].as_int("mpi.transfer_buffer");
        if (mp["discontinuity_buf"]) cfg.mpi.discontinuity_buf = (int)mp["discontinuity_buf"].as_int("mpi.discontinuity_buf");
    }
    // 10) flags
    if (doc["flags"]) {
        if (doc["flags"]["uses_dam"]) cfg.flags.uses_dam = doc["flags"]["uses_dam"].as_bool("flags.uses_dam");
        if (doc["flags"]["convert_area"]) cfg.flags.convert_area = doc["flags"]["convert_area"].as_bool("flags.convert_area");
    }
    return cfg;
}

inline SimulationConfig load_config(const std::string& filename) { return config_from_yaml(hlmyaml::LoadFile(filename)); }
