// models.cuh — device-side model definitions (the Model trait's `rhs`, compiled into the library).
//
// The reference's Model trait (models/model_204.hpp:15-115) hands `rhs` a pointer to the AoS
// SpatialParams array and the link index and re-reads 15 doubles from it on every call.  A device
// function cannot cross a C ABI, so models live here and are selected by UID at the boundary.
// Each model declares
//   UID, N_EQ          as in the reference trait
//   N_SP               number of per-link SoA parameter columns it needs
//   N_FORC             number of forcings its rhs reads (F[0] = rain, F[1] = temperature)
//   prepare()          one AoS record -> its SoA columns (run once at upload)
//   Link               the per-link constants held in registers for a whole window
//   rhs()              the right-hand side, written with fp<T> primitives only
#pragma once
#include "fp_exact.cuh"

namespace hlm {

// Byte layout of the reference's SpatialParams (I_O/parameters_loader.hpp:19-37): 2 x i64 + 15 x f64.
struct SpatialParamsAoS {
    long long stream, next_stream;
    double c1, infil, perco, Hu, lat, sw, ss, n_mann, slope, L, A_h, alpha3, alpha4, melt_f, temp_thr;
};
static_assert(sizeof(SpatialParamsAoS) == 136, "must match the reference's 136-byte record");

// ---------------------------------------------------------------------------------------------
// Model 204 — snow / static / surface / gravitational / aquifer storages.
// Arithmetic follows models/model_204.hpp:54-113 operation by operation, with the two
// contractions nvcc makes in the reference build (`d1 - s*Emax`, `d2 - h_surf*w`) written as fma.
// Hoisted out of rhs because they depend on parameters only and are correctly rounded, hence
// identical whenever computed: 1.0/n_mann (rcp.rn) and sqrt(slope) (sqrt.rn); and the Newton-refined
// reciprocals of the four constant divisors (fp_exact.cuh, div_recip/div_by).
// min_a/max_a take the operand that cannot be NaN first (a parameter, a constant, or a product of
// forcing and parameter), so a NaN in the state is dropped exactly as fmin/fmax drop it.
// ---------------------------------------------------------------------------------------------
struct Model204 {
    static constexpr int UID = 204;
    static constexpr int N_EQ = 5;
    static constexpr int N_SP = 15;
    static constexpr int N_FORC = 2;
    static constexpr bool HAS_INFLOW = false;
    // The surface store (state SURF) enters no other state's slope, and its own slope needs of the others only the
    // store's net inflow d2 = x2 - x3: the step can run the hillslope (rhs_hill) ahead and the surface store's chain
    // of pow()s (rhs_surf) beside it (dopri_attempt_split, rk45_window.cuh).  Same operations on the same operands.
    static constexpr bool SPLIT_SURFACE = true;
    static constexpr int SURF = 2;
    static constexpr int N_SIDE = 1;    // states of the side chain, in the order rhs_side takes and returns them
    static constexpr int HILL_OUT = 1;  // values rhs_hill hands to rhs_side: d2
    static __host__ __device__ constexpr int side_state(int k) { return k == 0 ? SURF : -1; }
    static __host__ __device__ constexpr bool is_side(int i) { return i == SURF; }
    enum { INFIL, PERCO, HU, INV_N, SQRT_SLOPE, LEN, A_H, ALPHA3, ALPHA4, MELT_F, TEMP_THR,
           R_HU, R_A_H, R_ALPHA3, R_ALPHA4 };  // R_* = fp<double>::div_recip of the divisor

    // wet_block: the surface branch's five parameters as one 64-byte record per link, stored behind the N_SP
    // columns of the prepared-parameter buffer (total (N_SP + kWetStride) * ld doubles)
    static constexpr int kWetStride = 8;
    static __host__ __device__ constexpr int wet_slot(int c) {
        return c == INV_N ? 0 : c == SQRT_SLOPE ? 1 : c == LEN ? 2 : c == A_H ? 3 : 4 /* R_A_H */;
    }
    static __device__ __forceinline__ void prepare_wet(const double* out, double* rec) {
        rec[0] = out[INV_N]; rec[1] = out[SQRT_SLOPE]; rec[2] = out[LEN]; rec[3] = out[A_H]; rec[4] = out[R_A_H];
        rec[5] = rec[6] = rec[7] = 0.0;
    }

    static __device__ __forceinline__ void prepare(const SpatialParamsAoS& s, double* out) {
        out[INFIL] = s.infil;
        out[PERCO] = s.perco;
        out[HU] = s.Hu;
        out[INV_N] = __drcp_rn(s.n_mann);
        out[SQRT_SLOPE] = __dsqrt_rn(s.slope);
        out[LEN] = s.L;
        out[A_H] = s.A_h;
        out[ALPHA3] = s.alpha3;
        out[ALPHA4] = s.alpha4;
        out[MELT_F] = s.melt_f;
        out[TEMP_THR] = s.temp_thr;
        out[R_HU] = fp<double>::div_recip(s.Hu);
        out[R_A_H] = fp<double>::div_recip(s.A_h);
        // The reference drains these stores only when alpha >= 1 (model_204.hpp:109,112), else the term
        // is 0.  A zero "reciprocal" makes the three-instruction quotient exactly +-0 for every finite
        // numerator, so the fast path needs no select; the exact path keeps the comparison.
        out[R_ALPHA3] = (s.alpha3 >= 1.0) ? fp<double>::div_recip(s.alpha3) : 0.0;
        out[R_ALPHA4] = (s.alpha4 >= 1.0) ? fp<double>::div_recip(s.alpha4) : 0.0;
    }

    // Per-link constants held in registers for a whole window.  The five parameters only the surface
    // (h_surf != 0) branch reads — INV_N, SQRT_SLOPE, LEN, A_H, R_A_H — are NOT kept: that branch
    // re-reads them through the read-only path (L1 hits), which frees 10 registers on a kernel that
    // sits at the 168-register cap of 3 CTAs/SM.  They come from a 64-byte record per link behind the
    // columns (wet_block), so the branch needs ONE address and immediate offsets: with five column
    // addresses the compiler rebuilt all five on every attempt, dry or not (16 instructions, 10 registers).
    template <typename T> struct Link {
        T p[N_SP];  // wet-only slots stay unused (dead registers are eliminated)
        const double* wet;  // the link's record {INV_N, SQRT_SLOPE, LEN, A_H, R_A_H, -, -, -}
        bool recips_ok;
        __device__ __forceinline__ void load(const double* __restrict__ sp, long long ld_, long long sys) {
            const int dry[] = {INFIL, PERCO, HU, ALPHA3, ALPHA4, MELT_F, TEMP_THR, R_HU, R_ALPHA3, R_ALPHA4};
#pragma unroll
            for (int i = 0; i < 10; ++i) p[dry[i]] = (T)__ldg(sp + (long long)dry[i] * ld_ + sys);
            wet = sp + (long long)N_SP * ld_ + sys * kWetStride;
            const double ra = __ldg(sp + (long long)R_A_H * ld_ + sys);
            recips_ok = p[R_HU] == p[R_HU] && ra == ra && p[R_ALPHA3] == p[R_ALPHA3] && p[R_ALPHA4] == p[R_ALPHA4];
        }
        __device__ __forceinline__ T wet_param(int c) const { return (T)__ldg(wet + wet_slot(c)); }
    };

    /// true when every hoisted reciprocal is usable (divisors inside div_recip's exponent range)
    template <typename T> static __device__ __forceinline__ bool fast_div_ok(const Link<T>& P) { return P.recips_ok; }

    template <typename T, bool kFast, typename G>
    static __device__ __forceinline__ void rhs(const T* y, const T* F, const Link<T>& P, T* dydt, G& bad) {
        using f = fp<T>;
        const T h_snow = y[0], h_stat = y[1], h_surf = y[2], h_grav = y[3], h_aq = y[4];
        const T rainfall = F[0], temperature = F[1];

        // 1) snow
        const T snowmelt = (temperature >= P.p[TEMP_THR]) ? f::min_a(f::mul(temperature, P.p[MELT_F]), h_snow) : (T)0;
        const T x1 = f::add(rainfall, snowmelt);
        dydt[0] = f::sub(rainfall, snowmelt);

        // 2) static
        const T x2 = f::max0(f::sub(f::add(x1, h_stat), P.p[HU]));
        const T d1 = f::sub(x1, x2);
        const T Emax = f::min_a(f::mul((T)0.1, temperature), h_stat);
        const T s = f::template div_by<kFast>(h_stat, P.p[HU], P.p[R_HU], bad);
        dydt[1] = f::fma(-s, Emax, d1);

        // 3) surface.  When h_surf is +-0 the reference's expression collapses exactly:
        // pow(+-0, 2/3) = +0, so w is 0 (or 1 if L/A_h make a NaN, which fmin drops) and
        // fma(-h_surf, w, d2) adds a signed zero to d2 >= +0, i.e. returns d2 bit for bit.
        // Skipping pow/div there is a warp-divergent but exact shortcut.
        const T x3 = f::min_a(P.p[INFIL], x2);
        const T d2 = f::sub(x2, x3);
        if (h_surf == (T)0) {
            dydt[2] = d2;
        } else {
            const T alfa2 = f::mul(f::mul(P.wet_param(INV_N), f::template pow_pos<kFast>(h_surf, (T)(2.0 / 3.0), bad)), P.wet_param(SQRT_SLOPE));
            const T w = f::min_a((T)1, f::mul(f::template div_by<kFast>(f::mul(alfa2, P.wet_param(LEN)), P.wet_param(A_H),
                                                                      P.wet_param(R_A_H), bad), (T)60));
            dydt[2] = f::fma(-h_surf, w, d2);
        }

        // 4) gravitational (interflow), 5) aquifer (baseflow)
        const T x4 = f::min_a(P.p[PERCO], x3);
        const T d3 = f::sub(x3, x4);
        if constexpr (kFast) {
            dydt[3] = f::sub(d3, f::template div_by<true>(h_grav, P.p[ALPHA3], P.p[R_ALPHA3], bad));
            dydt[4] = f::sub(x4, f::template div_by<true>(h_aq, P.p[ALPHA4], P.p[R_ALPHA4], bad));
        } else {
            dydt[3] = f::sub(d3, (P.p[ALPHA3] >= (T)1) ? f::div(h_grav, P.p[ALPHA3]) : (T)0);
            dydt[4] = f::sub(x4, (P.p[ALPHA4] >= (T)1) ? f::div(h_aq, P.p[ALPHA4]) : (T)0);
        }
    }

    // ---- the same right-hand side in two parts (fast forms only), for dopri_attempt_split -------------------------
    /// slopes of states 0, 1, 3, 4 (dydt[SURF] is not written; y[SURF] is not read) and, in ho[0], the surface store's
    /// net inflow d2 = x2 - x3.  Branch-free.
    template <typename T, typename G>
    static __device__ __forceinline__ void rhs_hill(const T* y, const T* F, const Link<T>& P, T* dydt, T (&ho)[HILL_OUT], G& bad) {
        using f = fp<T>;
        const T h_snow = y[0], h_stat = y[1], h_grav = y[3], h_aq = y[4];
        const T rainfall = F[0], temperature = F[1];
        const T snowmelt = (temperature >= P.p[TEMP_THR]) ? f::min_a(f::mul(temperature, P.p[MELT_F]), h_snow) : (T)0;
        const T x1 = f::add(rainfall, snowmelt);
        dydt[0] = f::sub(rainfall, snowmelt);
        const T x2 = f::max0(f::sub(f::add(x1, h_stat), P.p[HU]));
        const T d1 = f::sub(x1, x2);
        const T Emax = f::min_a(f::mul((T)0.1, temperature), h_stat);
        const T s = f::template div_by<true>(h_stat, P.p[HU], P.p[R_HU], bad);
        dydt[1] = f::fma(-s, Emax, d1);
        const T x3 = f::min_a(P.p[INFIL], x2);
        ho[0] = f::sub(x2, x3);
        const T x4 = f::min_a(P.p[PERCO], x3);
        const T d3 = f::sub(x3, x4);
        dydt[3] = f::sub(d3, f::template div_by<true>(h_grav, P.p[ALPHA3], P.p[R_ALPHA3], bad));
        dydt[4] = f::sub(x4, f::template div_by<true>(h_aq, P.p[ALPHA4], P.p[R_ALPHA4], bad));
    }
    /// slope of the surface store.  kWet = false: the store is taken to be empty (slope = d2, exactly what rhs gives
    /// for h_surf == +-0) and `not_dry` records a lane for which it is not.  kWet = true: branch-free — the pow chain
    /// runs on a stand-in operand where the store is empty and the result is dropped there, so no flag is raised for it.
    template <typename T, bool kWet, typename G>
    static __device__ __forceinline__ void rhs_side(const T (&ys)[N_SIDE], const T (&ho)[HILL_OUT], const Link<T>& P, T (&ks)[N_SIDE],
                                                    G& bad, bool& not_dry) {
        using f = fp<T>;
        const T h_surf = ys[0], d2 = ho[0];
        if constexpr (!kWet) {
            not_dry = not_dry || !(h_surf == (T)0);
            ks[0] = d2;
        } else {
            const bool empty = h_surf == (T)0;
            const T x = empty ? (T)0.5 : h_surf;
            const T alfa2 = f::mul(f::mul(P.wet_param(INV_N), f::template pow_pos<true>(x, (T)(2.0 / 3.0), bad)), P.wet_param(SQRT_SLOPE));
            const T w = f::min_a((T)1, f::mul(f::template div_by<true>(f::mul(alfa2, P.wet_param(LEN)), P.wet_param(A_H),
                                                                     P.wet_param(R_A_H), bad), (T)60));
            const T wet_slope = f::fma(-h_surf, w, d2);
            ks[0] = empty ? d2 : wet_slope;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Model 200 — hillslope-link runoff: channel discharge + static / surface / gravitational / aquifer
// storages.  PROJECT-DEFINED, NOT REFERENCE-DERIVED: the reference only names "model 200/204"
// (README.md:95) and ships no definition.  It is Model204's hillslope without the
// snow store (same expressions, same two contractions) draining into the link's channel with the
// Hillslope-Link Model's nonlinear-celerity routing equation (Mantilla & Gupta 2005; the form the
// Iowa HLM uses):
//     dq/dt = invtau * q_e^(1/5) * ((runoff*CH + q_in) - q),   q_e = max(q, 1e-6 m3/s)
//     runoff = (h_surf*w + out_grav) + out_aq   [m/min],  CH = A_h * 1e6/60  [km2 * m/min -> m3/s]
//     invtau = 19.8 / ((800 * L) * sqrt(A_h^0.2))   [1/min]   (v_r 0.33 m/s, lambda1 0.2, lambda2 -0.1)
// q_in is the discharge entering from upstream links (0 for an unrouted run); it is constant over an
// interval, set per link through WindowArgs::qin (routing, rk45_window.cuh).  Being project-defined, its two power
// laws — q_e^(1/5) and the surface store's h^(2/3) — are evaluated by fp<T>::root5 / cbrt2 (fp_exact.cuh: a bit-trick
// seed and Newton steps, within 3 ulp, a quarter of libdevice pow's instructions) rather than by pow(): with six right-hand
// sides per attempt the channel's pow alone was a third of the lane kernel's instructions.  Model204, whose arithmetic
// the reference defines, keeps pow().  The CPU restatement the tests compare with lives with the test infrastructure;
// the independent pin is SciPy (tests).
// ---------------------------------------------------------------------------------------------
struct Model200 {
    static constexpr int UID = 200;
    static constexpr int N_EQ = 5;
    static constexpr int N_SP = 15;
    static constexpr int N_FORC = 2;
    static constexpr bool HAS_INFLOW = true;
    // Not split (dopri_attempt_split): the surface store and, behind it, the channel do hang off the hillslope as a side
    // chain, and the split attempt gives the same bits, but with two pow chains beside the hillslope's sums the lane
    // kernel spills 1.3 KB per thread: routed hour 5.72 -> 6.07 ms.  Measured, not kept.
    static constexpr bool SPLIT_SURFACE = false;
    static constexpr int SURF = 2;
    static constexpr int N_SIDE = 0;
    static constexpr int HILL_OUT = 0;
    enum { INFIL, PERCO, HU, INV_N, SQRT_SLOPE, LEN, A_H, ALPHA3, ALPHA4, CH, INVTAU,
           R_HU, R_A_H, R_ALPHA3, R_ALPHA4 };

    // wet_block: the surface branch's five parameters as one 64-byte record per link, stored behind the N_SP
    // columns of the prepared-parameter buffer (total (N_SP + kWetStride) * ld doubles)
    static constexpr int kWetStride = 8;
    static __host__ __device__ constexpr int wet_slot(int c) {
        return c == INV_N ? 0 : c == SQRT_SLOPE ? 1 : c == LEN ? 2 : c == A_H ? 3 : 4 /* R_A_H */;
    }
    static __device__ __forceinline__ void prepare_wet(const double* out, double* rec) {
        rec[0] = out[INV_N]; rec[1] = out[SQRT_SLOPE]; rec[2] = out[LEN]; rec[3] = out[A_H]; rec[4] = out[R_A_H];
        rec[5] = rec[6] = rec[7] = 0.0;
    }

    static __device__ __forceinline__ void prepare(const SpatialParamsAoS& s, double* out) {
        out[INFIL] = s.infil;
        out[PERCO] = s.perco;
        out[HU] = s.Hu;
        out[INV_N] = __drcp_rn(s.n_mann);
        out[SQRT_SLOPE] = __dsqrt_rn(s.slope);
        out[LEN] = s.L;
        out[A_H] = s.A_h;
        out[ALPHA3] = s.alpha3;
        out[ALPHA4] = s.alpha4;
        out[CH] = __dmul_rn(s.A_h, 1.0e6 / 60.0);
        bool unused = false;
        const double a02 = fp<double>::pow_pos<false>(s.A_h, 0.2, unused);
        out[INVTAU] = __ddiv_rn(19.8, __dmul_rn(__dmul_rn(800.0, s.L), __dsqrt_rn(a02)));
        out[R_HU] = fp<double>::div_recip(s.Hu);
        out[R_A_H] = fp<double>::div_recip(s.A_h);
        out[R_ALPHA3] = (s.alpha3 >= 1.0) ? fp<double>::div_recip(s.alpha3) : 0.0;
        out[R_ALPHA4] = (s.alpha4 >= 1.0) ? fp<double>::div_recip(s.alpha4) : 0.0;
    }

    template <typename T> struct Link {
        T p[N_SP];
        T qin;
        const double* wet;  // the link's wet_block record (see Model204::Link)
        bool recips_ok;
        __device__ __forceinline__ void load(const double* __restrict__ sp, long long ld_, long long sys) {
            const int dry[] = {INFIL, PERCO, HU, ALPHA3, ALPHA4, CH, INVTAU, R_HU, R_ALPHA3, R_ALPHA4};
#pragma unroll
            for (int i = 0; i < 10; ++i) p[dry[i]] = (T)__ldg(sp + (long long)dry[i] * ld_ + sys);
            wet = sp + (long long)N_SP * ld_ + sys * kWetStride;
            qin = (T)0;
            const double ra = __ldg(sp + (long long)R_A_H * ld_ + sys);
            recips_ok = p[R_HU] == p[R_HU] && ra == ra && p[R_ALPHA3] == p[R_ALPHA3] && p[R_ALPHA4] == p[R_ALPHA4];
        }
        __device__ __forceinline__ void set_inflow(T v) { qin = v; }
        __device__ __forceinline__ T wet_param(int c) const { return (T)__ldg(wet + wet_slot(c)); }
    };

    template <typename T> static __device__ __forceinline__ bool fast_div_ok(const Link<T>& P) { return P.recips_ok; }

    template <typename T, bool kFast, typename G>
    static __device__ __forceinline__ void rhs(const T* y, const T* F, const Link<T>& P, T* dydt, G& bad) {
        using f = fp<T>;
        const T q = y[0], h_stat = y[1], h_surf = y[2], h_grav = y[3], h_aq = y[4];
        const T rainfall = F[0], temperature = F[1];

        // static store (Model204 without snow: x1 = rainfall)
        const T x2 = f::max0(f::sub(f::add(rainfall, h_stat), P.p[HU]));
        const T d1 = f::sub(rainfall, x2);
        const T Emax = f::min_a(f::mul((T)0.1, temperature), h_stat);
        const T s = f::template div_by<kFast>(h_stat, P.p[HU], P.p[R_HU], bad);
        dydt[1] = f::fma(-s, Emax, d1);

        // surface store; its outflow h_surf*w feeds the channel
        const T x3 = f::min_a(P.p[INFIL], x2);
        const T d2 = f::sub(x2, x3);
        T out_surf;
        if (h_surf == (T)0) {
            dydt[2] = d2;
            out_surf = (T)0;
        } else {
            const T alfa2 = f::mul(f::mul(P.wet_param(INV_N), f::cbrt2(h_surf)), P.wet_param(SQRT_SLOPE));
            const T w = f::min_a((T)1, f::mul(f::template div_by<kFast>(f::mul(alfa2, P.wet_param(LEN)), P.wet_param(A_H),
                                                                      P.wet_param(R_A_H), bad), (T)60));
            dydt[2] = f::fma(-h_surf, w, d2);
            out_surf = f::mul(h_surf, w);
        }

        // gravitational and aquifer stores
        const T x4 = f::min_a(P.p[PERCO], x3);
        const T d3 = f::sub(x3, x4);
        T out_grav, out_aq;
        if constexpr (kFast) {
            out_grav = f::template div_by<true>(h_grav, P.p[ALPHA3], P.p[R_ALPHA3], bad);
            out_aq = f::template div_by<true>(h_aq, P.p[ALPHA4], P.p[R_ALPHA4], bad);
        } else {
            out_grav = (P.p[ALPHA3] >= (T)1) ? f::div(h_grav, P.p[ALPHA3]) : (T)0;
            out_aq = (P.p[ALPHA4] >= (T)1) ? f::div(h_aq, P.p[ALPHA4]) : (T)0;
        }
        dydt[3] = f::sub(d3, out_grav);
        dydt[4] = f::sub(x4, out_aq);

        // channel
        const T runoff = f::add(f::add(out_surf, out_grav), out_aq);
        const T lateral = f::mul(runoff, P.p[CH]);
        const T qe = f::max_a((T)1e-6, q);
        const T cel = f::root5(qe);
        dydt[0] = f::mul(f::mul(P.p[INVTAU], cel), f::sub(f::add(lateral, P.qin), q));
    }
};

// ---------------------------------------------------------------------------------------------
// DummyModel — the 5-state linear test system of model_dummy_python.ipynb:65-89 (code cell,
// I2 = 0.6*H1).  The reference ships no C++ for it; operations are unfused, in the
// order Python evaluates the notebook's expressions, and match oracle/oracle_rk45.c:rhs_dummy.
// UID 0 is this project's choice (the reference never assigned one).
// ---------------------------------------------------------------------------------------------
struct DummyModel {
    static constexpr int UID = 0;
    static constexpr int N_EQ = 5;
    static constexpr int N_SP = 0;
    static constexpr int N_FORC = 0;
    static constexpr bool HAS_INFLOW = false;
    static constexpr bool SPLIT_SURFACE = false;
    static constexpr int SURF = 0;
    static constexpr int N_SIDE = 0;
    static constexpr int HILL_OUT = 0;
    static constexpr int kWetStride = 0;
    static __device__ __forceinline__ void prepare(const SpatialParamsAoS&, double*) {}
    static __device__ __forceinline__ void prepare_wet(const double*, double*) {}
    template <typename T> struct Link {
        __device__ __forceinline__ void load(const double*, long long, long long) {}
    };
    template <typename T> static __device__ __forceinline__ bool fast_div_ok(const Link<T>&) { return true; }
    template <typename T, bool kFast, typename G>
    static __device__ __forceinline__ void rhs(const T* y, const T*, const Link<T>&, T* dydt, G&) {
        using f = fp<T>;
        const T Y0 = f::mul((T)0.5, y[0]);
        const T X2 = f::mul((T)0.3, y[1]);
        const T I2 = f::mul((T)0.6, y[1]);
        const T I3 = f::mul((T)0.4, y[3]);
        dydt[0] = f::sub((T)1.0, Y0);
        dydt[1] = f::sub(f::sub(f::sub(f::add((T)1.2, Y0), X2), (T)0.4), I2);
        dydt[2] = f::sub(X2, (T)0.2);
        dydt[3] = f::sub(f::sub(I2, I3), (T)0.3);
        dydt[4] = f::sub(I3, (T)0.1);
    }
};

}  // namespace hlm
