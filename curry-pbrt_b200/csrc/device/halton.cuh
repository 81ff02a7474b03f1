// Halton sampler on the device: HaltonSampler::{set_pixel,next_sample,get_sample} and the
// Sampler trait defaults (src/sampler/halton.rs:24-136, src/sampler/mod.rs:10-36).
// All index arithmetic is integer and must match the reference bit for bit; indices exceed
// 2^32 for large films x sample counts, so they are carried as u64 with a 32-bit fast path
// for the digit loop (the quotient sequence is the same either way).
#pragma once
#include "scene_dev.cuh"

namespace cpbrt {

// `pre`: a 2-D sample drawn ahead of its turn (together with the draw before it, see
// sampler_get_usize_then_2d); the next sampler_get_2d returns it.
struct SamplerState { uint64_t index; uint32_t dim; bool has_pre = false; float pre_x = 0.f, pre_y = 0.f; };

// inverse (halton.rs:36-47) for base_index 0/1, returning base^digit_count in *pw
CP_DEV uint64_t digit_inverse(uint64_t x, uint32_t base, uint64_t* pw, uint32_t* count) {
    uint64_t acc = 0, p = 1; uint32_t c = 0;
    if (base == 2) {
        if (x) { c = 64 - __clzll((long long)x); acc = __brevll(x) >> (64 - c); p = c == 64 ? 0 : (1ull << c); }
    } else {
        while ((x >> 32) != 0) { uint64_t q = x / base; uint64_t digit = x - q * base; acc = acc * base + digit; p *= base; ++c; x = q; }
        uint32_t xs = (uint32_t)x;
        while (xs != 0) { uint32_t q = xs / base; uint32_t digit = xs - q * base; acc = acc * base + digit; p *= base; ++c; xs = q; }
    }
    *pw = p; *count = c;
    return acc;
}

// radical_inverse (halton.rs:65-70)
CP_DEV float radical_inverse(uint64_t x, uint32_t base) {
    uint64_t pw; uint32_t c;
    uint64_t acc = digit_inverse(x, base, &pw, &c);
    float inv_base_n = 1.f / __ull2float_rn(pw);
    return inv_base_n * __ull2float_rn(acc);
}

CP_DEV uint32_t halton_table(const DevSampler& S, uint32_t i) { CP_CHECK(i < S.table_len, 401); return __ldg(&S.permutation[i]); }

// scrambled_radical_inverse (halton.rs:56-64).  The reference peels one digit per round:
//     reversed = reversed * base + perm[a % base];  a /= base;      while a != 0
// Here, once the running quotient fits 32 bits, a round takes a whole GROUP of g digits (DevHaltonDim): with
// G = base^g, reversed * G + T_g[a % G] is what g of the reference's rounds produce when all g digits are still to
// come (a >= G), T_g[r] being the g digits of r permuted and reversed.  The last, partial group (a < G: j <= g digits
// left) uses the table of j-digit strings.  Divisions by G are one 64-bit multiply-high with floor(2^64/G)+1,
// exact for a < 2^32.  Same integer accumulator and digit count as the reference (cpbrt_halton's parity test
// against the oracle, which keeps the one-digit loop, covers 200 000 values incl. indices beyond 2^32).
CP_DEV void halton_finish_chain(const DevSampler& S, const float4 r0, const float4 r1, const float4 r2, uint32_t xs, uint64_t& acc, uint32_t& cnt) {
    const uint32_t G = __float_as_uint(r1.y), gd = __float_as_uint(r2.z), full_off = __float_as_uint(r2.y);
    const uint64_t gm = ((uint64_t)__float_as_uint(r1.w) << 32) | __float_as_uint(r1.z);
    while (xs >= G) {
        const uint32_t q = (uint32_t)__umul64hi((uint64_t)xs, gm);
        acc = acc * G + halton_table(S, full_off + (xs - q * G)); cnt += gd; xs = q;
    }
    if (xs != 0) {
        const uint32_t base = __float_as_uint(r0.z);
        uint32_t j = 1, pw = base, o = __float_as_uint(r2.x);
        while (xs >= pw) { o += pw; ++j; pw *= base; }
        acc = acc * pw + halton_table(S, o + xs); cnt += j;
    }
}
CP_DEV float scrambled_radical_inverse(const DevSampler& S, uint64_t x, uint32_t dim) {
    const float4* rec = reinterpret_cast<const float4*>(S.dims + dim);
    const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
    const uint32_t base = __float_as_uint(r0.z);
    const uint16_t* __restrict__ pm = S.permutation + __float_as_uint(r0.w);
    uint64_t acc = 0; uint32_t count = 0;
    while ((x >> 32) != 0) { uint64_t q = x / base; uint32_t digit = (uint32_t)(x - q * base); acc = acc * base + __ldg(&pm[digit]); ++count; x = q; }
    halton_finish_chain(S, r0, r1, r2, (uint32_t)x, acc, count);
    // inv_base_n and the tail of permuted zero digits were evaluated on the host with the same f32 operations
    const float inv_base_n = __ldg(&S.inv_pow[dim * kHaltonMaxDigits + count]);
    return inv_base_n * (__ull2float_rn(acc) + r1.x);
}

// N consecutive dimensions of the same index at once.  The digit loops of different primes are
// independent dependency chains (multiply-high -> digits -> table look-up -> accumulate); one at a
// time each round waits out its own latencies, and the shading kernels have too few resident
// warps to cover them.  Walking N chains in one loop lets each chain's latency hide behind the
// others' instructions.  Same quotients, digits and accumulators as N calls of
// scrambled_radical_inverse (matte shading stage: -3.5 % with pairs).
#ifndef CPBRT_HALTON_PAIR
#define CPBRT_HALTON_PAIR 1
#endif
template <int N>
CP_DEV void scrambled_radical_inverse_n(const DevSampler& S, uint64_t x, uint32_t dim, float* out) {
    uint64_t gm[N], acc[N]; uint32_t G[N], fo[N], xs[N], cnt[N];
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const float4* rec = reinterpret_cast<const float4*>(S.dims + dim + k);
        const float4 r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
        G[k] = __float_as_uint(r1.y); gm[k] = ((uint64_t)__float_as_uint(r1.w) << 32) | __float_as_uint(r1.z);
        fo[k] = __float_as_uint(r2.y);
        acc[k] = 0; cnt[k] = 0;
    }
    if ((x >> 32) != 0) {      // the quotients of a 64-bit index differ per prime: each chain on its own until it fits 32 bits
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const float4 r0 = __ldg(reinterpret_cast<const float4*>(S.dims + dim + k));
            const uint32_t base = __float_as_uint(r0.z), off = __float_as_uint(r0.w);
            uint64_t xk = x;
            while ((xk >> 32) != 0) { const uint64_t q = xk / base; const uint32_t digit = (uint32_t)(xk - q * base); acc[k] = acc[k] * base + halton_table(S, off + digit); ++cnt[k]; xk = q; }
            xs[k] = (uint32_t)xk;
        }
    } else {
#pragma unroll
        for (int k = 0; k < N; ++k) xs[k] = (uint32_t)x;
    }
    uint32_t rounds = 0;       // full groups taken by every chain in lock step
    for (;;) {
        bool all = true;
#pragma unroll
        for (int k = 0; k < N; ++k) all = all && xs[k] >= G[k];
        if (!all) break;
        uint32_t q[N], pv[N];
#pragma unroll
        for (int k = 0; k < N; ++k) q[k] = (uint32_t)__umul64hi((uint64_t)xs[k], gm[k]);
#pragma unroll
        for (int k = 0; k < N; ++k) pv[k] = halton_table(S, fo[k] + (xs[k] - q[k] * G[k]));
#pragma unroll
        for (int k = 0; k < N; ++k) { acc[k] = acc[k] * G[k] + pv[k]; xs[k] = q[k]; }
        ++rounds;
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const float4* rec = reinterpret_cast<const float4*>(S.dims + dim + k);
        const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1), r2 = __ldg(rec + 2);
        cnt[k] += rounds * __float_as_uint(r2.z);
        halton_finish_chain(S, r0, r1, r2, xs[k], acc[k], cnt[k]);
        out[k] = __ldg(&S.inv_pow[(dim + k) * kHaltonMaxDigits + cnt[k]]) * (__ull2float_rn(acc[k]) + r1.x);
    }
}

// set_pixel (halton.rs:108-119) with inverse_exp (halton.rs:48-55)
CP_DEV uint64_t halton_pixel_index(const DevSampler& S, uint32_t px, uint32_t py) {
    uint64_t idx = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const uint32_t base = i == 0 ? 2u : 3u;
        uint64_t pw; uint32_t c;
        uint64_t r = digit_inverse(i == 0 ? px : py, base, &pw, &c);
        for (uint32_t k = c; k < S.exp[i]; ++k) r *= base;
        idx += r * S.scale[1 - i] * S.mult_inverse[i];
    }
    return idx % S.stride;
}

// Dimensions >= 1000 (a path ~120 bounces long): the reference draws rand::random() from the thread's
// generator (halton.rs:131), which no implementation can reproduce.  A counter-based generator keyed
// on (index, dimension) keeps such draws independent and uniform (a constant would bias roulette and
// light selection on very long paths) and repeatable; the oracle uses the same function.
CP_DEV float halton_overflow_sample(uint64_t index, uint32_t dim) {
    uint64_t z = index * 0x9E3779B97F4A7C15ull + (uint64_t)dim * 0xD1B54A32D192ED03ull + 0x632BE59BD9B4E019ull;   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    return (float)(uint32_t)(z >> 40) * (1.f / 16777216.f);      // 24 bits: exact in f32, < 1
}
// get_sample (halton.rs:124-136)
CP_DEV float sampler_get_sample(const DevSampler& S, SamplerState& st) {
    float r;
    st.has_pre = false;   // a 1-D draw where the stashed 2-D one was expected: drop the stash, st.dim is still exact
    if (st.dim == 0) r = radical_inverse(st.index / S.scale[0], 2u);
    else if (st.dim == 1) r = radical_inverse(st.index / S.scale[1], 3u);
    else if (st.dim < 1000) r = scrambled_radical_inverse(S, st.index, st.dim);
    else r = halton_overflow_sample(st.index, st.dim);   // reference: rand::random(), not reproducible by design
    ++st.dim;
    return r;
}
// get_2d for dimensions 0 and 1 given the two quotients index / scale[0] and index / scale[1]
// (get_sample halton.rs:125-128 + the get_1d clamp): what the film-point jitter of every camera
// sample draws (render.rs:30)
CP_DEV P2 halton_film_sample(uint64_t q0, uint64_t q1) {
    P2 p;
    const float a = radical_inverse(q0, 2u), b = radical_inverse(q1, 3u);
    p.x = a >= 1.f ? 1.f : a; p.y = b >= 1.f ? 1.f : b;
    return p;
}
// ---- radical_inverse(x, 3) without the digit loop -------------------------------------------------
// A rank of a sample-partitioned render starts in the middle of the sample range and first has to
// replay the film-point drift of every sample before it (render.rs:26-31): two radical inverses per
// pixel and skipped sample, ~900 x 810 000 x 2 for the last of eight ranks.  The base-2 one is a bit
// reversal; the base-3 one walks 13 digits.  Here the reversed digit string comes from a table of
// reversed 7-digit chunks (3^7 = 2187 entries, built in shared memory by the block itself):
//     x = c0 + 2187 * c1  (x < 3^14)   ->   R = rev7(c0) * 2187 + rev7(c1)      (14 digits, reversed)
//     acc = R / 3^(14 - count)         exact, done as a multiplication by the inverse of 3^j modulo 2^32
// with count = number of base-3 digits of x.  Same integer acc, same count, same two f32 operations
// as radical_inverse (halton.rs:36-47, 65-70): bit-identical, and k_film_replay re-checks a sample
// of its values against the digit loop on every run.
struct Rev3Tables { uint16_t rev7[2187]; uint32_t pow3[16]; uint32_t inv3mod[16]; float inv_pw[16]; };
CP_DEV void rev3_tables_fill(Rev3Tables& t) {
    for (uint32_t c = threadIdx.x; c < 2187u; c += blockDim.x) {
        uint32_t x = c, r = 0;
#pragma unroll
        for (int j = 0; j < 7; ++j) { r = r * 3u + x % 3u; x /= 3u; }
        t.rev7[c] = (uint16_t)r;
    }
    if (threadIdx.x < 16) {
        uint32_t pw = 1, inv = 1;
        for (uint32_t j = 0; j < threadIdx.x; ++j) { pw *= 3u; inv *= 0xAAAAAAABu; }      // 3 * 0xAAAAAAAB == 1 (mod 2^32)
        t.pow3[threadIdx.x] = pw; t.inv3mod[threadIdx.x] = inv; t.inv_pw[threadIdx.x] = 1.f / __ull2float_rn((unsigned long long)pw);
    }
    __syncthreads();
}
CP_DEV float radical_inverse3_fast32(const Rev3Tables& t, uint32_t x) {      // x < 3^14
    if (x == 0) return 0.f;
    uint32_t c = x >= t.pow3[8] ? 8u : 0u;                       // c = number of digits - 1
    if (x >= t.pow3[c + 4]) c += 4;
    if (x >= t.pow3[c + 2]) c += 2;
    if (x >= t.pow3[c + 1]) c += 1;
    const uint32_t c1 = x / 2187u, c0 = x - c1 * 2187u;
    const uint32_t R = (uint32_t)t.rev7[c0] * 2187u + t.rev7[c1];
    const uint32_t acc = R * t.inv3mod[13u - c];
    return t.inv_pw[c + 1] * __uint2float_rn(acc);               // acc < 2^32: the same value __ull2float_rn gives
}
CP_DEV float radical_inverse3_fast(const Rev3Tables& t, uint64_t x64) {
    if (x64 >= 4782969ull) return radical_inverse(x64, 3u);      // 3^14 and beyond: the digit loop
    return radical_inverse3_fast32(t, (uint32_t)x64);
}
// radical_inverse(x, 2) for x < 2^32: bit reversal; 1 / 2^c is built from its exponent (exact, as the division is)
CP_DEV float radical_inverse2_fast32(uint32_t x) {
    if (x == 0) return 0.f;
    const uint32_t c = 32u - (uint32_t)__clz((int)x);
    const uint32_t acc = __brev(x) >> (32u - c);
    return __uint_as_float((127u - c) << 23) * __uint2float_rn(acc);
}

// get_1d (sampler/mod.rs:10-17): the clamp constant 0.9999999999999 is 1.0 in f32
CP_DEV float sampler_get_1d(const DevSampler& S, SamplerState& st) { float r = sampler_get_sample(S, st); return r >= 1.f ? 1.f : r; }
CP_DEV P2 sampler_get_2d(const DevSampler& S, SamplerState& st) {
    P2 p;
#if CPBRT_HALTON_PAIR
    if (st.has_pre) { st.has_pre = false; st.dim += 2; p.x = st.pre_x; p.y = st.pre_y; return p; }
    if (st.dim >= 2 && st.dim + 1 < 1000) {
        float u[2];
        scrambled_radical_inverse_n<2>(S, st.index, st.dim, u);
        st.dim += 2;
        p.x = u[0] >= 1.f ? 1.f : u[0]; p.y = u[1] >= 1.f ? 1.f : u[1];
        return p;
    }
#endif
    p.x = sampler_get_1d(S, st); p.y = sampler_get_1d(S, st);
    return p;
}
// get_usize (sampler/mod.rs:24-33)
CP_DEV uint32_t sampler_get_usize(const DevSampler& S, SamplerState& st, uint32_t max) {
    float f = sampler_get_sample(S, st) * (float)max;
    uint32_t r = f32_to_u32_sat(truncf(f));
    return r == max ? max - 1 : r;
}

// get_usize followed by a get_2d (uniform_sample_one_light: light choice, then the light's own
// 2-D sample, integrator/mod.rs:20-30): the three dimensions are evaluated together, the 2-D
// sample waits in the state for the sampler_get_2d that the light's sample_li issues.
CP_DEV uint32_t sampler_get_usize_then_2d(const DevSampler& S, SamplerState& st, uint32_t max) {
#if CPBRT_HALTON_PAIR
    if (st.dim >= 2 && st.dim + 2 < 1000 && !st.has_pre) {
        float u[3];
        scrambled_radical_inverse_n<3>(S, st.index, st.dim, u);
        st.dim += 1;
        st.has_pre = true; st.pre_x = u[1] >= 1.f ? 1.f : u[1]; st.pre_y = u[2] >= 1.f ? 1.f : u[2];
        const float f = u[0] * (float)max;
        const uint32_t r = f32_to_u32_sat(truncf(f));
        return r == max ? max - 1 : r;
    }
#endif
    return sampler_get_usize(S, st, max);
}

}  // namespace cpbrt
