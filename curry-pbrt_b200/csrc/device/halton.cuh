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

// scrambled_radical_inverse (halton.rs:56-64): acc = acc * base + perm[digit] over the digits of x,
// least significant first, then (acc + tail) / base^digits.
//
// The digit loop is one dependent chain (divide -> digit -> table look-up -> multiply-add) of 5..14
// steps per dimension, and a matte bounce draws seven dimensions: 28 % of the shading kernel's
// instructions (ncu, profiles/r01_s3_ncu_cornell.txt).  k consecutive digits of x are the base-p
// digits of r = x mod p^k, and what they add to the accumulator,
//     acc' = acc * p^k + (perm[d0] * p^(k-1) + ... + perm[d(k-1)]),
// depends on r alone: it is tabulated per dimension (`chunk_full`, B = p^k entries) for every
// prime whose B fits 16 bits.  The final chunk has only as many digits as the remaining quotient
// (the reference stops at x == 0, leading zeros belong to `tail`), so it has its own table with the
// digit count and p^m beside the value (`chunk_last`).  Same integers in acc and the same digit
// count as the digit-by-digit loop, so the value is bit-identical (cpbrt_halton checks every
// requested value against the one-digit-at-a-time evaluation on the device, and the parity test
// against the oracle runs over it).  A 32-bit index now takes at most 3 steps in any dimension.
// While the running quotient fits 32 bits the division is one 64-bit multiply-high with
// floor(2^64 / B) + 1, exact for x < 2^32 and B < 2^17.
struct HaltonChain { uint64_t magic, acc; uint32_t B, kd, off, cnt, xs; };

CP_DEV void halton_chain_begin(const DevSampler& S, uint64_t x, uint32_t dim, HaltonChain& c, float* tail) {
    const float4* rec = reinterpret_cast<const float4*>(S.dims + dim);
    const float4 r0 = __ldg(rec), r1 = __ldg(rec + 1);
    c.magic = ((uint64_t)__float_as_uint(r0.y) << 32) | __float_as_uint(r0.x);
    c.B = __float_as_uint(r0.z); c.kd = __float_as_uint(r0.w); c.off = __float_as_uint(r1.x);
    *tail = r1.w;
    c.acc = 0; c.cnt = 0;
    if ((x >> 32) != 0 || c.kd == 1) {
        // digit by digit: until a 64-bit index fits 32 bits, and all the way for primes without a chunk table
        const uint32_t base = __float_as_uint(r1.z);
        const uint16_t* __restrict__ pm = S.permutation + __float_as_uint(r1.y);
        while ((x >> 32) != 0) { const uint64_t q = x / base; const uint32_t digit = (uint32_t)(x - q * base); c.acc = c.acc * base + __ldg(&pm[digit]); ++c.cnt; x = q; }
        if (c.kd == 1) {
            uint32_t xs = (uint32_t)x;
            while (xs != 0) { const uint32_t q = (uint32_t)__umul64hi((uint64_t)xs, c.magic); c.acc = c.acc * base + __ldg(&pm[xs - q * base]); ++c.cnt; xs = q; }
            x = 0;
        }
    }
    c.xs = (uint32_t)x;
}
CP_DEV bool halton_chain_full(const HaltonChain& c) { return c.xs >= c.B; }
CP_DEV void halton_chain_last(const DevSampler& S, HaltonChain& c) {
    if (c.xs != 0) {
        const uint2 e = __ldg(&S.chunk_last[c.off + c.xs]);
        c.acc = c.acc * e.x + (e.y & 0xFFFFFFu); c.cnt += e.y >> 24; c.xs = 0;
    }
}
CP_DEV float halton_chain_value(const DevSampler& S, const HaltonChain& c, uint32_t dim, float tail) {
    // inv_base_n and the tail of permuted zero digits were evaluated on the host with the same f32 operations
    return __ldg(&S.inv_pow[dim * kHaltonMaxDigits + c.cnt]) * (__ull2float_rn(c.acc) + tail);
}

CP_DEV float scrambled_radical_inverse(const DevSampler& S, uint64_t x, uint32_t dim) {
    HaltonChain c; float tail;
    halton_chain_begin(S, x, dim, c, &tail);
    while (halton_chain_full(c)) {
        const uint32_t q = (uint32_t)__umul64hi((uint64_t)c.xs, c.magic);
        c.acc = c.acc * c.B + __ldg(&S.chunk_full[c.off + (c.xs - q * c.B)]); c.cnt += c.kd; c.xs = q;
    }
    halton_chain_last(S, c);
    return halton_chain_value(S, c, dim, tail);
}

// The reference loop itself, one digit per step: what cpbrt_halton compares the chunked evaluation with.
CP_DEV float scrambled_radical_inverse_digitwise(const DevSampler& S, uint64_t x, uint32_t dim) {
    const DevHaltonDim& d = S.dims[dim];
    const uint32_t base = d.base;
    const uint16_t* __restrict__ pm = S.permutation + d.primsum;
    uint64_t acc = 0; uint32_t count = 0;
    while (x != 0) { const uint64_t q = x / base; acc = acc * base + __ldg(&pm[(uint32_t)(x - q * base)]); ++count; x = q; }
    return __ldg(&S.inv_pow[dim * kHaltonMaxDigits + count]) * (__ull2float_rn(acc) + d.tail);
}

// N consecutive dimensions of the same index at once.  The chains of different primes are
// independent (multiply-high -> remainder -> table look-up -> accumulate); one at a time each step
// waits out its own latencies, and the shading kernels have too few resident warps to cover them.
// Walking N chains in one loop lets each chain's latency hide behind the others' instructions.
// Same quotients, chunks and accumulators as N calls of scrambled_radical_inverse.
#ifndef CPBRT_HALTON_PAIR
#define CPBRT_HALTON_PAIR 1
#endif
template <int N>
CP_DEV void scrambled_radical_inverse_n(const DevSampler& S, uint64_t x, uint32_t dim, float* out) {
    HaltonChain c[N]; float tail[N];
#pragma unroll
    for (int k = 0; k < N; ++k) halton_chain_begin(S, x, dim + k, c[k], &tail[k]);
    for (;;) {
        bool all = true;
#pragma unroll
        for (int k = 0; k < N; ++k) all = all && halton_chain_full(c[k]);
        if (!all) break;
        uint32_t q[N], pv[N];
#pragma unroll
        for (int k = 0; k < N; ++k) q[k] = (uint32_t)__umul64hi((uint64_t)c[k].xs, c[k].magic);
#pragma unroll
        for (int k = 0; k < N; ++k) pv[k] = __ldg(&S.chunk_full[c[k].off + (c[k].xs - q[k] * c[k].B)]);
#pragma unroll
        for (int k = 0; k < N; ++k) { c[k].acc = c[k].acc * c[k].B + pv[k]; c[k].cnt += c[k].kd; c[k].xs = q[k]; }
    }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        while (halton_chain_full(c[k])) {
            const uint32_t q = (uint32_t)__umul64hi((uint64_t)c[k].xs, c[k].magic);
            c[k].acc = c[k].acc * c[k].B + __ldg(&S.chunk_full[c[k].off + (c[k].xs - q * c[k].B)]); c[k].cnt += c[k].kd; c[k].xs = q;
        }
    }
    uint2 e[N];
#pragma unroll
    for (int k = 0; k < N; ++k) e[k] = c[k].xs != 0 ? __ldg(&S.chunk_last[c[k].off + c[k].xs]) : make_uint2(1u, 0u);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        c[k].acc = c[k].acc * e[k].x + (e[k].y & 0xFFFFFFu); c[k].cnt += e[k].y >> 24;
        out[k] = halton_chain_value(S, c[k], dim + k, tail[k]);
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

// get_sample (halton.rs:124-136)
CP_DEV float sampler_get_sample(const DevSampler& S, SamplerState& st) {
    float r;
    st.has_pre = false;   // a 1-D draw where the stashed 2-D one was expected: drop the stash, st.dim is still exact
    if (st.dim == 0) r = radical_inverse(st.index / S.scale[0], 2u);
    else if (st.dim == 1) r = radical_inverse(st.index / S.scale[1], 3u);
    else if (st.dim < 1000) r = scrambled_radical_inverse(S, st.index, st.dim);
    else r = 0.5f;   // reference draws rand::random() here; unreachable for sane max_depth
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
