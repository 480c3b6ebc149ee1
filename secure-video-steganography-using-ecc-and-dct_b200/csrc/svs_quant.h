// svs_quant.h - the quantiser of the packed kernels, stated once in plain C++.
//
// (1) FastQuant / make_fast_quant: the host-computed constants of the division-free quantiser
//     used by svs_fast.cuh / svs_tile.cuh / svs_row.cuh.
// (2) A scalar restatement of what those kernels compute per coefficient (speculative FMA
//     quantiser + "too close to call" flag, and the IEEE-exact quantiser built from a rounded
//     reciprocal), on std::fmaf.  The kernels run the same arithmetic on packed FFMA2; this
//     restatement exists so that the CPU test-suite (tests/host_math) can check the claims the
//     kernels rely on against the plain IEEE division of the reference
//     (config_and_setup.py:148,160) without a GPU:
//       * an unflagged coefficient always gets the reference's result,
//       * the exact quantiser always does.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

namespace svs {

struct FastQuant {
    // embed: y = fma(c, r2, ke) = M + floor(c/(2 delta) + 1/4) + fraction, k fraction bits
    float r2, ke, d2, k0;
    uint32_t emask, ebit;           // 2^k - 1, 1 << (k-1)
    int erot;                       // k - 1: where the payload bit is inserted (value 1/2)
    // extract: y = fma(c, r, kx) = M + floor(c/delta + 1/2) + fraction; bit xk is the parity
    float r, kx;
    uint32_t xmask;
    int xk;
    float negzero;                  // -0.0f, opaque to the compiler
    int embed_ok, extract_ok;
};

constexpr uint32_t kQuantZone = 4;  // flagged when the fraction field is < 4 (the bias is 2 ulp)

// |c| <= 2040 for any 8x8 block of bytes (orthonormal basis, L1 norm <= 8), so
// x = c/(2 delta) + 1/4 (embed) and c/delta + 1/2 (extract) are bounded and M = 1.5 * 2^(23-k)
// leaves k fraction bits in the mantissa of M + x.  Error budget (units of 2^-k): rounding of
// the FMA 1/2, reciprocal instead of division 1/4, the reference's own quotient rounding 1/4
// -> strictly below 1; the kernels flag 2.
inline FastQuant make_fast_quant(double delta)
{
    FastQuant q;
    std::memset(&q, 0, sizeof q);
    q.negzero = -0.0f;
    const float d32 = (float)delta;
    if (!(delta >= 0x1p-4) || !(delta <= 0x1p20)) return q;
    {
        const double xmax = 1020.0 / d32 + 1.5;
        int k = 22 - (int)std::ceil(std::log2(xmax + 1.0));
        if (k > 20) k = 20;
        const double M = 1.5 * std::ldexp(1.0, 23 - k);
        q.r2 = (float)(0.5 / (double)d32);
        q.ke = (float)(M + 0.25 + 2.0 * std::ldexp(1.0, -k));
        q.d2 = 2.0f * d32;
        const double k0 = -(double)q.d2 * M;
        q.k0 = (float)k0;
        q.emask = (1u << k) - 1u;
        q.ebit = 1u << (k - 1);
        q.erot = k - 1;
        q.embed_ok = k >= 8 && (double)q.k0 == k0 && (double)d32 == delta &&
                     (double)q.ke == M + 0.25 + 2.0 * std::ldexp(1.0, -k);
    }
    {
        const double xmax = 2040.0 / d32 + 1.5;
        int k = 22 - (int)std::ceil(std::log2(xmax + 1.0));
        if (k > 20) k = 20;
        const double M = 1.5 * std::ldexp(1.0, 23 - k);
        q.r = (float)(1.0 / (double)d32);
        q.kx = (float)(M + 0.5 + 2.0 * std::ldexp(1.0, -k));
        q.xmask = (1u << k) - 1u;
        q.xk = k;
        q.extract_ok = k >= 8 && (double)q.kx == M + 0.5 + 2.0 * std::ldexp(1.0, -k);
    }
    return q;
}

#if !defined(__CUDA_ARCH__)
// ---- scalar restatement (host) ----------------------------------------------------------------
inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// what the reference does: q = rint(c / delta) in binary32, q' = q - (q & 1) + bit, float32(q' * delta)
inline float ref_embed(float c, float d32, int bit)
{
    const int q = (int)std::nearbyintf(c / d32);
    return (float)(q - (q & 1) + bit) * d32;
}
inline int ref_parity(float c, float d32) { return (int)std::nearbyintf(c / d32) & 1; }

// speculative quantiser (embed_fast_kernel): returns the value, sets `flagged` when the kernel
// would send the coefficient to the exact path instead
inline float fast_embed(const FastQuant& q, float c, int bit, bool& flagged)
{
    const uint32_t y = f2u(std::fmaf(c, q.r2, q.ke));
    flagged = (y & q.emask) < kQuantZone;
    return std::fmaf(u2f((y & ~q.emask) | ((uint32_t)bit << q.erot)), q.d2, q.k0);
}
inline int fast_parity(const FastQuant& q, float c, bool& flagged)
{
    const uint32_t y = f2u(std::fmaf(c, q.r, q.kx));
    flagged = (y & q.xmask) < kQuantZone;
    return (int)((y >> q.xk) & 1u);
}

// IEEE-exact c / d from the correctly rounded reciprocal (div_exact), then rint through the
// 1.5 * 2^23 constant; parity = lowest mantissa bit (ExactQ in svs_fast.cuh)
inline float exact_rint_plus_magic(float c, float d, float r)
{
    float t = c * r;
    t = std::fmaf(std::fmaf(t, -d, c), r, t);
    t = std::fmaf(std::fmaf(t, -d, c), r, t);
    return t + 12582912.0f;
}
inline float exact_embed(float c, float d, float r, int bit)
{
    const float m = exact_rint_plus_magic(c, d, r);
    const int adj = bit - (int)(f2u(m) & 1u);
    return ((m - 12582912.0f) + (float)adj) * d;
}
inline int exact_parity(float c, float d, float r) { return (int)(f2u(exact_rint_plus_magic(c, d, r)) & 1u); }
#endif

}  // namespace svs
