// The fast number path of the scene lexer (csrc/host/parser.cpp, parse_number) against strtof / the i32 grammar,
// on millions of literals of every shape a scene file holds.  Built and run by tests/test_host_parser.py.
#include <cerrno>
#include <cinttypes>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

namespace cpbrt_host { int parse_number(const char* begin, size_t n, int32_t* iv, float* fv); }

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

static bool all_digits(const std::string& w, size_t from) {
    if (from >= w.size()) return false;
    for (size_t k = from; k < w.size(); ++k) if (w[k] < '0' || w[k] > '9') return false;
    return true;
}

int main(int argc, char** argv) {
    const long n = argc > 1 ? atol(argv[1]) : 2000000;
    long checked = 0, ints = 0, floats = 0, bad = 0;
    char buf[128];
    for (long it = 0; it < n; ++it) {
        const uint64_t r = rnd();
        std::string w;
        switch (r % 12) {
            case 0: snprintf(buf, sizeof buf, "%.9g", (double)(float)((double)(int64_t)(rnd() % 2000000) / 1000.0 - 1000.0)); break;          // what tools write
            case 1: snprintf(buf, sizeof buf, "%d", (int)(rnd() % 4000000) - 2000000); break;                                                 // indices
            case 2: { uint32_t bits = (uint32_t)rnd(); float f; memcpy(&f, &bits, 4); if (!(f == f) || f - f != 0.f) f = 1.f; snprintf(buf, sizeof buf, "%.9g", (double)f); break; }   // any finite float
            case 3: { uint32_t bits = (uint32_t)rnd(); float f; memcpy(&f, &bits, 4); if (!(f == f) || f - f != 0.f) f = 1.f; snprintf(buf, sizeof buf, "%.17g", (double)f); break; }
            case 4: snprintf(buf, sizeof buf, "%" PRIu64 ".%06" PRIu64, rnd() % 1000, rnd() % 1000000); break;
            case 5: snprintf(buf, sizeof buf, "-%" PRIu64 ".%03" PRIu64 "e%d", rnd() % 100, rnd() % 1000, (int)(rnd() % 60) - 30); break;
            case 6: snprintf(buf, sizeof buf, "%" PRIu64, (uint64_t)(rnd() % 100000000000ull)); break;                                                     // beyond i32: a float
            case 7: { double d = (double)(rnd() % (1ull << 53)) / (double)(1ull << (rnd() % 60)); snprintf(buf, sizeof buf, "%.20g", d); break; }   // long mantissas
            case 8: { uint32_t bits = (uint32_t)rnd(); float f; memcpy(&f, &bits, 4); if (!(f == f) || f - f != 0.f) f = 1.f;                 // exact midpoints of two floats
                      float g = nextafterf(f, f > 0 ? 3e38f : -3e38f); snprintf(buf, sizeof buf, "%.60g", ((double)f + (double)g) / 2.0); buf[60] = 0; break; }
            case 9: snprintf(buf, sizeof buf, "%s", (rnd() & 1) ? "-0" : "-0.0"); break;
            case 10: snprintf(buf, sizeof buf, "000%" PRIu64 ".5000", rnd() % 1000); break;
            default: snprintf(buf, sizeof buf, "%" PRIu64 "E+%d", rnd() % 100000, (int)(rnd() % 12)); break;
        }
        w = buf;
        if (w.empty() || !(w[0] == '-' || (w[0] >= '0' && w[0] <= '9'))) continue;
        // reference answer: the i32 grammar first, else strtof on the whole word
        int want_kind = 0; int32_t want_i = 0; float want_f = 0.f;
        const size_t from = w[0] == '-' ? 1 : 0;
        if (all_digits(w, from)) {
            errno = 0; long long v = strtoll(w.c_str(), nullptr, 10);
            if (errno == 0 && v >= INT32_MIN && v <= INT32_MAX) { want_kind = 1; want_i = (int32_t)v; }
        }
        if (!want_kind) { char* end = nullptr; want_f = strtof(w.c_str(), &end); want_kind = (end && *end == 0) ? 2 : 0; }
        int32_t iv = 0; float fv = 0.f;
        const int kind = cpbrt_host::parse_number(w.data(), w.size(), &iv, &fv);
        ++checked;
        bool same = kind == want_kind;
        if (same && kind == 1) { same = iv == want_i; ++ints; }
        if (same && kind == 2) { uint32_t a, b; memcpy(&a, &fv, 4); memcpy(&b, &want_f, 4); same = a == b; ++floats; }
        if (!same && ++bad <= 10) printf("MISMATCH \"%s\": kind %d (want %d) i %d (want %d) f %.9g (want %.9g)\n", w.c_str(), kind, want_kind, iv, want_i, (double)fv, (double)want_f);
    }
    printf("lexer_number_check: %ld literals (%ld integers, %ld floats), %ld mismatches\n", checked, ints, floats, bad);
    return bad ? 1 : 0;
}
