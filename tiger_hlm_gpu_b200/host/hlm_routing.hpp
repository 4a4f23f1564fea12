// hlm_routing.hpp — host side of routed runs (links coupled through upstream channel discharge).
//
// The reference carries the river topology in every parameter record (`stream`, `next_stream`:
// I_O/parameters_loader.cpp:74, stream.hpp:31,47) and sketches per-step MPI buffers (data/config.yaml:66-70)
// but couples nothing; the partition and the per-interval driver below are this project's design and the
// topology columns are the only contract taken from the reference.  Same algorithm, same results as the
// Python mirror tiger_hlm_gpu_b200/routing.py (tests compare the two plans array by array):
//   * sub-basins = connected sub-trees of `next_stream` closed when they reach `subbasin_links` links;
//   * whole sub-basins are dealt to ranks in contiguous runs balanced by link count;
//   * a rank's links keep their original relative order (links sorted by forcing cell stay sorted);
//   * boundary links (upstream end of a cut edge) are exchanged as one padded segment per rank;
//   * upstream lists are in ascending original index, so inflow sums are partition-independent.
#pragma once

#include <algorithm>
#include <cstdint>
#include <functional>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hlm_b200/rk45_api.hpp"

namespace hlm_b200 {

struct RankTopology {
    int rank = 0;
    long long lo = 0, hi = 0;            // planned positions [lo, hi)
    std::vector<long long> up_ptr;       // [n_local + 1]
    std::vector<int> up_idx;             // >= 0 local link, < 0 halo slot -(e + 1)
    std::vector<int> send_idx;           // boundary links, in halo-segment order
    long long n_local() const { return hi - lo; }
};

struct RoutePlan {
    int world = 1;
    std::vector<long long> order;        // planned position -> original link index
    std::vector<RankTopology> ranks;
    long long max_send = 0;              // length of every rank's segment of the halo vector
    long long n_subbasins = 0, n_cut_edges = 0;
    long long halo_len() const { return (long long)world * max_send; }
};

/// Index of each link's downstream link, -1 for outlets (next_stream not among `stream`, or itself).
inline std::vector<long long> downstream_index(const std::vector<long long>& stream, const std::vector<long long>& next_stream) {
    const long long n = (long long)stream.size();
    std::vector<long long> order((size_t)n);
    std::iota(order.begin(), order.end(), 0LL);
    std::stable_sort(order.begin(), order.end(), [&](long long a, long long b) { return stream[(size_t)a] < stream[(size_t)b]; });
    for (long long i = 1; i < n; ++i)
        if (stream[(size_t)order[(size_t)i]] == stream[(size_t)order[(size_t)i - 1]]) throw std::runtime_error("duplicate stream ids");
    std::vector<long long> down((size_t)n, -1);
    for (long long i = 0; i < n; ++i) {
        const long long want = next_stream[(size_t)i];
        auto it = std::lower_bound(order.begin(), order.end(), want,
                                   [&](long long a, long long v) { return stream[(size_t)a] < v; });
        if (it != order.end() && stream[(size_t)*it] == want && *it != i) down[(size_t)i] = *it;
    }
    return down;
}

inline RoutePlan plan_routes(const std::vector<long long>& stream, const std::vector<long long>& next_stream, int world,
                             long long subbasin_links = 4096) {
    if (world < 1) throw std::runtime_error("plan_routes: world must be >= 1");
    const std::vector<long long> down = downstream_index(stream, next_stream);
    const long long n = (long long)down.size();
    // leaves-first order of the forest (Kahn); a cycle leaves links unvisited
    std::vector<int> indeg((size_t)n, 0);
    for (long long i = 0; i < n; ++i)
        if (down[(size_t)i] >= 0) ++indeg[(size_t)down[(size_t)i]];
    std::vector<long long> topo;
    topo.reserve((size_t)n);
    for (long long i = 0; i < n; ++i)
        if (indeg[(size_t)i] == 0) topo.push_back(i);
    for (size_t k = 0; k < topo.size(); ++k) {
        const long long d = down[(size_t)topo[k]];
        if (d >= 0 && --indeg[(size_t)d] == 0) topo.push_back(d);
    }
    if ((long long)topo.size() != n) throw std::runtime_error("next_stream has a cycle");
    // close a sub-basin when the open sub-tree above a link (itself included) reaches subbasin_links links
    std::vector<long long> size((size_t)n, 1), root((size_t)n);
    std::vector<char> cut((size_t)n, 0);
    for (long long i : topo) {
        const long long d = down[(size_t)i];
        cut[(size_t)i] = (size[(size_t)i] >= subbasin_links || d < 0) ? 1 : 0;
        if (!cut[(size_t)i]) size[(size_t)d] += size[(size_t)i];
    }
    for (auto it = topo.rbegin(); it != topo.rend(); ++it) {
        const long long i = *it;
        root[(size_t)i] = cut[(size_t)i] ? i : root[(size_t)down[(size_t)i]];
    }
    // sub-basins in the order of their first link, contiguous runs per rank balanced by link count
    std::vector<long long> first((size_t)n, -1), count((size_t)n, 0), roots;
    for (long long i = 0; i < n; ++i) {
        const long long r = root[(size_t)i];
        if (first[(size_t)r] < 0) {
            first[(size_t)r] = i;
            roots.push_back(r);  // already in order of first link
        }
        ++count[(size_t)r];
    }
    if ((long long)roots.size() < world)
        throw std::runtime_error(std::to_string(roots.size()) + " sub-basins cannot be dealt to " + std::to_string(world) + " ranks: lower subbasin_links");
    // contiguous runs: a rank is closed when taking half of the next sub-basin would carry it past its share of
    // the links, or when the sub-basins left are only just enough to give every remaining rank one
    std::vector<long long> rank_of_root((size_t)n, 0);
    {
        long long rank = 0, cum = 0, have = 0;
        for (size_t i = 0; i < roots.size(); ++i) {
            const long long cnt = count[(size_t)roots[i]];
            if (rank < world - 1 && have > 0 &&
                ((2 * cum + cnt) * world > 2 * (rank + 1) * n || (long long)(roots.size() - i) <= world - 1 - rank)) {
                ++rank;
                have = 0;
            }
            rank_of_root[(size_t)roots[i]] = rank;
            cum += cnt;
            have += cnt;
        }
    }
    std::vector<long long> owner((size_t)n);
    for (long long i = 0; i < n; ++i) owner[(size_t)i] = rank_of_root[(size_t)root[(size_t)i]];

    RoutePlan p;
    p.world = world;
    p.n_subbasins = (long long)roots.size();
    p.order.resize((size_t)n);
    std::iota(p.order.begin(), p.order.end(), 0LL);
    std::stable_sort(p.order.begin(), p.order.end(), [&](long long a, long long b) { return owner[(size_t)a] < owner[(size_t)b]; });
    std::vector<long long> pos((size_t)n);
    for (long long k = 0; k < n; ++k) pos[(size_t)p.order[(size_t)k]] = k;
    std::vector<long long> per_rank((size_t)world, 0);
    for (long long i = 0; i < n; ++i) ++per_rank[(size_t)owner[(size_t)i]];
    p.ranks.resize((size_t)world);
    long long lo = 0;
    for (int r = 0; r < world; ++r) {
        p.ranks[(size_t)r].rank = r;
        p.ranks[(size_t)r].lo = lo;
        p.ranks[(size_t)r].hi = lo + per_rank[(size_t)r];
        p.ranks[(size_t)r].up_ptr.assign((size_t)per_rank[(size_t)r] + 1, 0);
        lo += per_rank[(size_t)r];
    }
    // boundary links in planned order -> slots
    std::vector<long long> slot_of((size_t)n, -1);
    std::vector<char> is_boundary((size_t)n, 0);
    for (long long u = 0; u < n; ++u) {
        const long long d = down[(size_t)u];
        if (d >= 0 && owner[(size_t)u] != owner[(size_t)d]) {
            is_boundary[(size_t)u] = 1;
            ++p.n_cut_edges;
        }
    }
    for (long long k = 0; k < n; ++k) {
        const long long u = p.order[(size_t)k];
        if (is_boundary[(size_t)u]) {
            RankTopology& t = p.ranks[(size_t)owner[(size_t)u]];
            t.send_idx.push_back((int)(k - t.lo));
        }
    }
    for (const RankTopology& t : p.ranks) p.max_send = std::max<long long>(p.max_send, (long long)t.send_idx.size());
    for (const RankTopology& t : p.ranks)
        for (size_t k = 0; k < t.send_idx.size(); ++k)
            slot_of[(size_t)p.order[(size_t)(t.lo + t.send_idx[k])]] = (long long)t.rank * p.max_send + (long long)k;
    // upstream lists: counting pass, then fill walking upstream ends in ascending ORIGINAL index
    for (long long u = 0; u < n; ++u) {
        const long long d = down[(size_t)u];
        if (d < 0) continue;
        RankTopology& t = p.ranks[(size_t)owner[(size_t)d]];
        ++t.up_ptr[(size_t)(pos[(size_t)d] - t.lo) + 1];
    }
    for (RankTopology& t : p.ranks) {
        for (size_t i = 1; i < t.up_ptr.size(); ++i) t.up_ptr[i] += t.up_ptr[i - 1];
        t.up_idx.assign((size_t)t.up_ptr.back(), 0);
    }
    std::vector<long long> fill((size_t)n, 0);
    for (long long u = 0; u < n; ++u) {
        const long long d = down[(size_t)u];
        if (d < 0) continue;
        RankTopology& t = p.ranks[(size_t)owner[(size_t)d]];
        const long long row = pos[(size_t)d] - t.lo;
        const long long e = t.up_ptr[(size_t)row] + fill[(size_t)d]++;
        t.up_idx[(size_t)e] = owner[(size_t)u] == t.rank ? (int)(pos[(size_t)u] - t.lo) : (int)(-(slot_of[(size_t)u] + 1));
    }
    return p;
}

/// Per-interval driver of one rank's routed run.  `exchange(d_send, n_send_padded, d_halo)` is the caller's
/// collective on device buffers (ncclAllGather on the context's stream, or MPI for a CUDA-aware build) — or, in
/// peer mode (hlm_route_peer_*), just a barrier; it is not called when world == 1 or nothing crosses ranks.  The implicit fallback is switched on for the run: a
/// link the explicit path abandons would freeze and starve everything downstream.
/// consecutive rejections before a routed run flags a link stiff (hlm_set_reject_limit; the Python mirror's
/// routing.ROUTED_REJECT_LIMIT)
constexpr int kRoutedRejectLimit = 20;

class RoutedRun {
  public:
    using Exchange = std::function<void(const double* d_send, long long n, double* d_halo)>;
    RoutedRun(Context& ctx, int uid, const RankTopology& topo, int world, long long max_send, Exchange exchange = nullptr,
              double* d_send = nullptr, double* d_halo = nullptr)
        : ctx_(ctx), uid_(uid), world_(world), max_send_(max_send), exchange_(std::move(exchange)), d_send_(d_send), d_halo_(d_halo) {
        check(hlm_route_set_topology(ctx.get(), topo.up_ptr.data(), topo.up_idx.empty() ? nullptr : topo.up_idx.data(),
                                     topo.n_local(), topo.send_idx.empty() ? nullptr : topo.send_idx.data(),
                                     (long long)topo.send_idx.size()),
              "hlm_route_set_topology");
        check(hlm_set_stiff_fallback(ctx.get(), 1), "hlm_set_stiff_fallback");
        check(hlm_set_reject_limit(ctx.get(), kRoutedRejectLimit), "hlm_set_reject_limit");
        if (world_ > 1 && max_send_ > 0) {
            // with the peer-memory exchange (hlm_route_peer_alloc/open done by the caller) the kernels deliver the
            // data themselves: the callback is then only a barrier on the stream and the two buffers stay null
            if (!exchange_) throw std::runtime_error("RoutedRun: a multi-rank run needs an exchange (a collective, or a barrier in peer mode)");
            if ((d_send_ == nullptr) != (d_halo_ == nullptr)) throw std::runtime_error("RoutedRun: give both device buffers or neither");
            if (d_send_) check(hlm_route_set_send_buffer(ctx.get(), d_send_), "hlm_route_set_send_buffer");
        }
    }
    ~RoutedRun() {
        hlm_set_stiff_fallback(ctx_.get(), 0);
        hlm_set_reject_limit(ctx_.get(), 5);
        hlm_route_clear(ctx_.get());
    }
    /// First interval [t0, tf]; tq = query times inside it.
    void begin(const std::vector<double>& y0, long long ns, double t0, double tf, const std::vector<double>& tq) {
        check(hlm_solve_begin(ctx_.get(), uid_, y0.data(), ns, t0, tf, tq.empty() ? nullptr : tq.data(), (long long)tq.size()),
              "hlm_solve_begin");
        check(hlm_route_pack(ctx_.get()), "hlm_route_pack");
        gather();
        check(hlm_solve_window(ctx_.get(), (long long)tq.size(), tq.empty() ? 0 : 1), "hlm_solve_window");
    }
    /// Next interval, up to tf: exchange + gather the inflow, continue every link (time and step size kept).
    void advance(double tf, const std::vector<double>& tq) {
        gather();
        check(hlm_solve_advance(ctx_.get(), tf, tq.empty() ? nullptr : tq.data(), (long long)tq.size()), "hlm_solve_advance");
        check(hlm_solve_window(ctx_.get(), (long long)tq.size(), tq.empty() ? 0 : 1), "hlm_solve_window");
    }

  private:
    void gather() {
        if (world_ > 1 && max_send_ > 0) {
            exchange_(d_send_, max_send_, d_halo_);
            check(hlm_route_gather(ctx_.get(), d_halo_), "hlm_route_gather");
        } else {
            check(hlm_route_gather(ctx_.get(), nullptr), "hlm_route_gather");
        }
    }
    Context& ctx_;
    int uid_, world_;
    long long max_send_;
    Exchange exchange_;
    double *d_send_, *d_halo_;
};

}  // namespace hlm_b200
