// format_g.hpp -- the text form the reference gives every number it writes: `ostream << double`, i.e. printf's "%g"
// with six significant digits (file.cpp:702-761 for the Phred columns, :1814-1873 in LikelihoodFile mode).
//
// glibc's %g is exact (it rounds the binary value's full decimal expansion, ties to even) but costs ~200 ns; the
// output writer prints six of them per sample.  append_g() produces the same bytes ~8x faster:
//   v * 10^k (k chosen so that the product has six integer digits) is formed in x87 extended precision, where v and
//   10^k (|k| <= 27) are exact and the product carries one rounding error of at most 2^-64 relative (< 6e-14 absolute);
//   the six digits are the product rounded to an integer.  If the fractional part is within 1e-9 of 1/2, or the number
//   is outside the range this argument covers (non-finite, negative, < 1e-22, >= 1e22), snprintf decides.
// cli/format_check.cpp compares the two on hundreds of millions of values (tests/test_host_cpu.py runs it).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

namespace famseq {

inline void append_g_slow(std::string &out, double v) {
    char buf[40];
    const int n = std::snprintf(buf, sizeof buf, "%g", v);
    out.append(buf, (size_t)n);
}

// The "%g" text of the number m * 10^(e10 - 5), m a six-digit integer (100000 .. 999999): scientific notation when the
// decimal exponent is below -4 or at least 6, trailing zeros dropped.  Writes at most 13 characters, returns the length.
inline int emit_decimal6(char *buf, uint32_t m, int e10) {
    char d[6];
    for (int i = 5; i >= 0; i--) {
        d[i] = (char)('0' + m % 10u);
        m /= 10u;
    }
    int nd = 6; // significant digits left after dropping trailing zeros
    while (nd > 1 && d[nd - 1] == '0') nd--;
    int n = 0;
    if (e10 < -4 || e10 >= 6) { // d.ddddde+XX
        buf[n++] = d[0];
        if (nd > 1) {
            buf[n++] = '.';
            for (int i = 1; i < nd; i++) buf[n++] = d[i];
        }
        buf[n++] = 'e';
        int x = e10;
        if (x < 0) {
            buf[n++] = '-';
            x = -x;
        } else {
            buf[n++] = '+';
        }
        if (x >= 100) buf[n++] = (char)('0' + x / 100);
        buf[n++] = (char)('0' + (x / 10) % 10);
        buf[n++] = (char)('0' + x % 10);
    } else if (e10 >= 0) { // e10 + 1 integer digits, then the rest
        for (int i = 0; i <= e10; i++) buf[n++] = d[i];
        if (nd > e10 + 1) {
            buf[n++] = '.';
            for (int i = e10 + 1; i < nd; i++) buf[n++] = d[i];
        }
    } else { // 0.000ddd
        buf[n++] = '0';
        buf[n++] = '.';
        for (int i = -1; i > e10; i--) buf[n++] = '0';
        for (int i = 0; i < nd; i++) buf[n++] = d[i];
    }
    return n;
}

inline void append_g(std::string &out, double v) {
    static const long double kPow10[28] = {1e0L,  1e1L,  1e2L,  1e3L,  1e4L,  1e5L,  1e6L,  1e7L,  1e8L,  1e9L,
                                           1e10L, 1e11L, 1e12L, 1e13L, 1e14L, 1e15L, 1e16L, 1e17L, 1e18L, 1e19L,
                                           1e20L, 1e21L, 1e22L, 1e23L, 1e24L, 1e25L, 1e26L, 1e27L};
    if (v == 0.0 && !std::signbit(v)) {
        out += '0';
        return;
    }
    if (!(v >= 1e-22 && v < 1e22)) { // negative, NaN, infinite, tiny or huge
        append_g_slow(out, v);
        return;
    }
    // decimal exponent: floor(log10(v)) estimated from the binary exponent, corrected after scaling
    int e2;
    std::frexp(v, &e2);
    int e10 = (int)std::floor((e2 - 1) * 0.30102999566398120);
    long double scaled = 0;
    for (int attempt = 0; attempt < 3; attempt++) {
        const int k = 5 - e10;
        scaled = k >= 0 ? (long double)v * kPow10[k] : (long double)v / kPow10[-k];
        if (scaled < 1e5L)
            e10--;
        else if (scaled >= 1e6L)
            e10++;
        else
            break;
    }
    if (!(scaled >= 1e5L && scaled < 1e6L)) {
        append_g_slow(out, v);
        return;
    }
    uint32_t m = (uint32_t)scaled; // truncation
    const long double frac = scaled - (long double)m;
    if (frac > 0.5L - 1e-9L && frac < 0.5L + 1e-9L) { // too close to a tie for this arithmetic
        append_g_slow(out, v);
        return;
    }
    if (frac > 0.5L) m++;
    if (m == 1000000u) {
        m = 100000u;
        e10++;
    }
    char buf[24];
    out.append(buf, (size_t)emit_decimal6(buf, m, e10));
}

} // namespace famseq
