// format_check.cpp -- append_g() (format_g.hpp) against snprintf("%g") on n pseudo-random doubles per family:
// raw bit patterns over the whole covered exponent range, Phred values -10*log10(p), decimal ties, neighbours of
// powers of ten.  usage: format_check [n] [seed]; exit status 1 and the first mismatches on stderr if any differ.
#include <cinttypes>
#include <cstdlib>

#include "../host/format_g.hpp"

namespace {
uint64_t s_state;
uint64_t next_u64() { // splitmix64
    uint64_t z = (s_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
double unit() { return (double)(next_u64() >> 11) * 0x1.0p-53; }
long long g_bad = 0, g_n = 0;
void check(double v) {
    std::string a, b;
    famseq::append_g(a, v);
    famseq::append_g_slow(b, v);
    g_n++;
    if (a != b && g_bad++ < 20) std::fprintf(stderr, "MISMATCH %.17g (%a): fast '%s' libc '%s'\n", v, v, a.c_str(), b.c_str());
}
} // namespace

int main(int argc, char **argv) {
    const long long n = argc > 1 ? std::atoll(argv[1]) : 2000000;
    s_state = argc > 2 ? std::strtoull(argv[2], nullptr, 10) : 1;
    const double edge[] = {0.0, -0.0, 1.0, 10.0, 100000.0, 999999.5, 999999.4999999999, 1e6, 1e-4, 1e-5, 9.9999949999e-5, 0.5, 1.5, 2.5, 123456.5, 1234565.0,
                           1e21, 1e22, 1e-22, 9.99e-23, 4.82164e-16, 99999.0, 1e300, 1e-300, 5e-324, -1.5, 1.0 / 0.0, -1.0 / 0.0, 0.0 / 0.0,
                           100000.5, 100001.5, 0.1000005, 0.0001000005, 1.000005, 1.000015, 1.0000049999999999, 8.5e-5, 0.000123456549999};
    for (double v : edge) check(v);
    for (long long i = 0; i < n; i++) {
        // (1) raw bit patterns, exponents covering [1e-24, 1e24]
        uint64_t bits = next_u64() & 0x000fffffffffffffull;
        bits |= (uint64_t)(1023 - 80 + (int)(next_u64() % 161)) << 52;
        double v;
        std::memcpy(&v, &bits, 8);
        check(v);
        // (2) Phred values of posteriors: p uniform, p near 0, p near 1
        const double u = unit();
        check(std::fabs(-10 * std::log10(u)));
        check(std::fabs(-10 * std::log10(std::pow(10.0, -30 * unit()))));
        check(std::fabs(-10 * std::log10(1.0 - std::pow(10.0, -16 * unit()))));
        // (3) short decimals: exact and near ties at the sixth digit
        const double six = (double)(100000 + next_u64() % 900000);
        const int sh = (int)(next_u64() % 12) - 6;
        check((six + 0.5) * std::pow(10.0, sh));
        check(std::nextafter((six + 0.5) * std::pow(10.0, sh), 0.0));
        check(std::nextafter((six + 0.5) * std::pow(10.0, sh), 1e300));
        check(six * std::pow(10.0, sh));
        // (4) neighbours of powers of ten
        const double p10 = std::pow(10.0, (int)(next_u64() % 40) - 20);
        check(p10 * (1.0 + (unit() - 0.5) * 1e-5));
        check(std::nextafter(p10, 0.0));
    }
    std::printf("%lld values compared, %lld mismatches\n", g_n, g_bad);
    return g_bad ? 1 : 0;
}
