/* devroot.h — TEST INFRASTRUCTURE.  C twin of fp<double>::root5 / cbrt2 (tiger_hlm_gpu_b200/csrc/fp_exact.cuh): the
 * x^(1/5) and x^(2/3) of Model 200, which is project-defined (the reference ships no Model 200), so
 * these functions ARE the definition: a seed of the inverse root from the exponent field, three Newton steps on the
 * inverse root, one first-order correction of the root.  Every operation is an IEEE multiplication or fma, written
 * out so that this file and the device code perform the same operations on the same operands (compile with
 * -ffp-contract=off).  Accuracy against long-double pow: tests/test_devroot.py. */
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline double devroot_from_hi(uint32_t hi) {
    uint64_t b = (uint64_t)hi << 32;
    double d;
    memcpy(&d, &b, 8);
    return d;
}
static inline uint32_t devroot_hi(double x) {
    uint64_t b;
    memcpy(&b, &x, 8);
    return (uint32_t)(b >> 32);
}
static inline double oracle_root5(double x) {
    double y = devroot_from_hi(0x4cb8a895u - devroot_hi(x) / 5u);
    const double w = 0.2 * x;
    for (int i = 0; i < 3; ++i) {
        const double y2 = y * y, y4 = y2 * y2, y5 = y4 * y;
        y = y * fma(-w, y5, 1.2);
    }
    const double y2 = y * y, y4 = y2 * y2, y5 = y4 * y;
    const double g = x * y4;
    return fma(g, 0.8 * fma(-x, y5, 1.0), g);
}
static inline double oracle_cbrt2(double x) {
    double y = devroot_from_hi(0x553ef0e8u - devroot_hi(x) / 3u);
    const double w = x * (1.0 / 3.0);
    for (int i = 0; i < 3; ++i) {
        const double y2 = y * y, y3 = y2 * y;
        y = y * fma(-w, y3, 4.0 / 3.0);
    }
    const double y2 = y * y, y3 = y2 * y;
    const double g = x * y;
    return fma(g, (1.0 / 3.0) * fma(-x, y3, 1.0), g);
}
