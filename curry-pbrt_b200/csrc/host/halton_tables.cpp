// Halton sampler set-up that runs once per scene on the host:
//   PRIMS / PRIMSUM (src/sampler/halton.rs:141-214), HaltonSampler::new (halton.rs:81-102) and
//   the digit-permutation table (halton.rs:216-231).
#include <mutex>
#include <vector>

#include "scene_host.h"

namespace cpbrt_host {

namespace {
uint32_t g_primes[CPBRT_N_PRIMES];
uint32_t g_prime_sums[CPBRT_N_PRIMES];
std::once_flag g_once;
void init_tables() {
    // the reference lists the first 1000 primes literally; a sieve gives the same list
    const int limit = 8000;
    std::vector<char> composite(limit, 0);
    int n = 0;
    for (int i = 2; i < limit && n < CPBRT_N_PRIMES; ++i) {
        if (composite[i]) continue;
        g_primes[n++] = (uint32_t)i;
        for (int j = i * i; j < limit; j += i) composite[j] = 1;
    }
    g_prime_sums[0] = 0;
    for (int i = 1; i < CPBRT_N_PRIMES; ++i) g_prime_sums[i] = g_prime_sums[i - 1] + g_primes[i - 1];
}

// extended_gcd / multiplicative_inverse (halton.rs:5-18), including the isize -> usize cast of
// a possibly negative Bezout coefficient before the modulo (Q4).
void extended_gcd(uint64_t a, uint64_t b, int64_t* x, int64_t* y) {
    if (b == 0) { *x = 1; *y = 0; return; }
    uint64_t d = a / b;
    int64_t xp, yp;
    extended_gcd(b, a % b, &xp, &yp);
    *x = yp;
    *y = xp - (int64_t)d * yp;
}
uint64_t multiplicative_inverse(uint64_t a, uint64_t n) {
    int64_t x, y;
    extended_gcd(a, n, &x, &y);
    return (uint64_t)x % n;
}
}  // namespace

const uint32_t* primes_1000() { std::call_once(g_once, init_tables); return g_primes; }
const uint32_t* prime_sums_1000() { std::call_once(g_once, init_tables); return g_prime_sums; }

void halton_sampler_params(uint32_t spp, uint32_t xres, uint32_t yres, cpbrt_sampler* out) {
    const uint32_t* primes = primes_1000();
    uint64_t res[2] = {xres, yres};
    out->spp = spp;
    for (int i = 0; i < 2; ++i) {
        out->scale[i] = 1; out->exp[i] = 0;
        while (out->scale[i] < res[i]) { out->scale[i] *= primes[i]; out->exp[i] += 1; }
    }
    out->mult_inverse[0] = multiplicative_inverse(out->scale[1], out->scale[0]);
    out->mult_inverse[1] = multiplicative_inverse(out->scale[0], out->scale[1]);
}

// The reference fills PERMUTATION with one thread_rng() shuffle per prime, different in every
// process.  Here the table is reproducible: a Fisher-Yates shuffle per prime block driven by
// Knuth's 64-bit LCG, j = (state >> 33) mod (i+1) for i = p-1 .. 1, one LCG step per swap, the
// state running on from block to block.
void make_permutation(uint64_t seed, uint16_t* out) {
    const uint32_t* primes = primes_1000();
    uint64_t x = seed;
    size_t base = 0;
    for (int k = 0; k < CPBRT_N_PRIMES; ++k) {
        uint32_t p = primes[k];
        uint16_t* perm = out + base;
        for (uint32_t i = 0; i < p; ++i) perm[i] = (uint16_t)i;
        for (uint32_t i = p - 1; i >= 1; --i) {
            x = x * 6364136223846793005ull + 1442695040888963407ull;
            uint32_t j = (uint32_t)((x >> 33) % (i + 1));
            uint16_t t = perm[i]; perm[i] = perm[j]; perm[j] = t;
        }
        base += p;
    }
}

// Digit-group tables of scrambled_radical_inverse (csrc/device/halton.cuh, DevHaltonDim): for every dimension whose
// base^2 still fits `group_max`, the permuted-and-reversed strings of 1, 2, ... g digits (g = the most digits with
// base^g <= group_max), appended behind the caller's CPBRT_PERM_TABLE_LEN single-digit entries.  T_j[x] for
// x < base^j is what j rounds of `reversed = reversed * base + perm[a % base]; a /= base` produce from reversed = 0.
void build_halton_groups(const uint16_t* permutation, uint32_t group_max, std::vector<HaltonGroupDim>* dims, std::vector<uint16_t>* extra) {
    const uint32_t* primes = primes_1000();
    const uint32_t* sums = prime_sums_1000();
    dims->assign(CPBRT_N_PRIMES, HaltonGroupDim());
    extra->clear();
    const uint32_t table0 = CPBRT_PERM_TABLE_LEN;
    for (int i = 0; i < CPBRT_N_PRIMES; ++i) {
        const uint32_t b = primes[i];
        HaltonGroupDim& d = (*dims)[i];
        uint32_t g = 1, G = b;
        while ((uint64_t)G * b <= group_max) { G *= b; ++g; }
        d.base = b; d.group = G; d.group_digits = g; d.group_magic = 0xFFFFFFFFFFFFFFFFull / G + 1;
        if (g == 1) { d.group_off = sums[i]; d.full_off = sums[i]; continue; }
        d.group_off = table0 + (uint32_t)extra->size();
        uint32_t pw = 1;
        for (uint32_t j = 1; j <= g; ++j) {
            pw *= b;
            if (j == g) d.full_off = table0 + (uint32_t)extra->size();
            for (uint32_t x = 0; x < pw; ++x) {
                uint32_t v = 0, t = x;
                for (uint32_t c = 0; c < j; ++c) { v = v * b + permutation[sums[i] + t % b]; t /= b; }
                extra->push_back((uint16_t)v);
            }
        }
    }
}

// The device's evaluation order on the host (tests): one digit at a time while the quotient needs 64 bits, whole
// groups while it is at least `group`, the partial last group from the table of its digit count.
void halton_grouped_reference(const uint16_t* permutation, const HaltonGroupDim& d, const std::vector<uint16_t>& extra, uint32_t primsum,
                              uint64_t index, uint64_t* acc_out, uint32_t* digits_out) {
    auto table = [&](uint32_t i) -> uint64_t { return i < (uint32_t)CPBRT_PERM_TABLE_LEN ? permutation[i] : extra[i - CPBRT_PERM_TABLE_LEN]; };
    uint64_t acc = 0, x = index; uint32_t cnt = 0;
    while ((x >> 32) != 0) { const uint64_t q = x / d.base; acc = acc * d.base + permutation[primsum + (uint32_t)(x - q * d.base)]; ++cnt; x = q; }
    uint32_t xs = (uint32_t)x;
    while (xs >= d.group) {
        const uint32_t q = (uint32_t)(((unsigned __int128)xs * d.group_magic) >> 64);      // the device's __umul64hi
        acc = acc * d.group + table(d.full_off + (xs - q * d.group)); cnt += d.group_digits; xs = q;
    }
    if (xs != 0) {
        uint32_t j = 1, pw = d.base, o = d.group_off;
        while (xs >= pw) { o += pw; ++j; pw *= d.base; }
        acc = acc * pw + table(o + xs); cnt += j;
    }
    *acc_out = acc; *digits_out = cnt;
}

}  // namespace cpbrt_host
