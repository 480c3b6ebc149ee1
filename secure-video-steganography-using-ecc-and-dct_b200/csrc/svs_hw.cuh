// svs_hw.cuh - the handful of machine operations the packed kernels are written in, each with
// two bodies: the sm_100a instruction (inline PTX / intrinsic) and a plain C++ restatement of
// what that instruction computes.  The host bodies exist so that the CPU test-suite
// (tests/host_math) can run the block-level code of svs_block.cuh - register layouts, the
// scalar/packed stage split, the quantiser, the bit packing - against the oracle without a
// GPU.  They are test infrastructure; nothing in the product calls them (the library has no
// CPU path: every entry point needs a CUDA device).
#pragma once

#include <cstdint>
#if !defined(__CUDA_ARCH__)
#include <cmath>
#include <cstring>
#endif

#include "svs_math.cuh"

namespace hw {

typedef unsigned long long u64;

// two binary32 values in one 64-bit register pair: .lo = bits 0..31, .hi = bits 32..63
struct P2 { u64 v; };

#if defined(__CUDA_ARCH__)
// ---------------------------------------------------------------------------- device ----------
SVS_HD P2 pk(float a, float b) { P2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
SVS_HD P2 pku(uint32_t a, uint32_t b) { P2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "r"(a), "r"(b)); return r; }
SVS_HD void unpk(P2 p, uint32_t& a, uint32_t& b) { asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(p.v)); }
SVS_HD void unpkf(P2 p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); }
SVS_HD P2 add2(P2 a, P2 b) { P2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SVS_HD P2 sub2(P2 a, P2 b) { P2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
SVS_HD P2 fma2(P2 a, P2 b, P2 c) { P2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
SVS_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
SVS_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
SVS_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
SVS_HD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
SVS_HD uint32_t f2u(float f) { return __float_as_uint(f); }
SVS_HD float u2f(uint32_t u) { return __uint_as_float(u); }
SVS_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
SVS_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_l(lo, hi, s); }
SVS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s) { return __funnelshift_r(lo, hi, s); }
SVS_HD uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) { return __dp2a_lo(a, b, c); }
SVS_HD uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) { return __dp2a_hi(a, b, c); }
SVS_HD uint32_t dp4a(uint32_t a, uint32_t b, uint32_t c) { return __dp4a(a, b, c); }
SVS_HD uint32_t absdiff4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }
SVS_HD int f2i_rn(float v) { return __float2int_rn(v); }
SVS_HD float i2f(int v) { return __int2float_rn(v); }
// np.uint8(np.clip(v, 0, 255)): saturate, then truncate toward zero - one F2IP
SVS_HD uint32_t to_u8(float v)
{
    uint32_t r;
    asm("{.reg .u8 t; cvt.rzi.u8.f32 t, %1; cvt.u32.u8 %0, t;}" : "=r"(r) : "f"(v));
    return r;
}
// four pixels -> one word, x0 in the low byte.  Written as float -> s32 (toward zero) followed by
// the saturating byte pack: ptxas fuses each pair into ONE two-source F2IP (convert two floats,
// shift the previous pair up) - 2 instructions per word instead of 4 conversions + 3 PRMT.
// Truncation before saturation = the reference's clip before truncation on every finite input.
SVS_HD uint32_t pack4_u8(float x0, float x1, float x2, float x3)
{
    uint32_t hi, r;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, 0;" : "=r"(hi) : "r"(__float2int_rz(x3)), "r"(__float2int_rz(x2)));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(__float2int_rz(x1)), "r"(__float2int_rz(x0)), "r"(hi));
    return r;
}
SVS_HD uint32_t umin(uint32_t a, uint32_t b) { return min(a, b); }
#else
// ---------------------------------------------------------------------------- host ------------
SVS_HD float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
SVS_HD uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
SVS_HD P2 pku(uint32_t a, uint32_t b) { P2 r; r.v = (u64)a | ((u64)b << 32); return r; }
SVS_HD P2 pk(float a, float b) { return pku(f2u(a), f2u(b)); }
SVS_HD void unpk(P2 p, uint32_t& a, uint32_t& b) { a = (uint32_t)p.v; b = (uint32_t)(p.v >> 32); }
SVS_HD void unpkf(P2 p, float& a, float& b) { a = u2f((uint32_t)p.v); b = u2f((uint32_t)(p.v >> 32)); }
// compiled with -ffp-contract=off: every operator below is one IEEE binary32 operation
SVS_HD float fadd(float a, float b) { return a + b; }
SVS_HD float fsub(float a, float b) { return a - b; }
SVS_HD float fmul(float a, float b) { return a * b; }
SVS_HD float ffma(float a, float b, float c) { return std::fmaf(a, b, c); }
SVS_HD P2 add2(P2 a, P2 b) { float a0, a1, b0, b1; unpkf(a, a0, a1); unpkf(b, b0, b1); return pk(a0 + b0, a1 + b1); }
SVS_HD P2 sub2(P2 a, P2 b) { float a0, a1, b0, b1; unpkf(a, a0, a1); unpkf(b, b0, b1); return pk(a0 - b0, a1 - b1); }
SVS_HD P2 fma2(P2 a, P2 b, P2 c)
{
    float a0, a1, b0, b1, c0, c1;
    unpkf(a, a0, a1); unpkf(b, b0, b1); unpkf(c, c0, c1);
    return pk(std::fmaf(a0, b0, c0), std::fmaf(a1, b1, c1));
}
SVS_HD uint32_t byte_perm(uint32_t a, uint32_t b, uint32_t sel)
{
    const u64 src = (u64)a | ((u64)b << 32);
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) r |= (uint32_t)((src >> (8 * ((sel >> (4 * k)) & 7u))) & 0xffu) << (8 * k);
    return r;
}
SVS_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s)
{
    s &= 31u;
    return (uint32_t)(((((u64)hi << 32) | lo) << s) >> 32);
}
SVS_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s)
{
    s &= 31u;
    return (uint32_t)((((u64)hi << 32) | lo) >> s);
}
SVS_HD uint32_t dp2a_lo(uint32_t a, uint32_t b, uint32_t c) { return c + (a & 0xffffu) * (b & 0xffu) + (a >> 16) * ((b >> 8) & 0xffu); }
SVS_HD uint32_t dp2a_hi(uint32_t a, uint32_t b, uint32_t c) { return c + (a & 0xffffu) * ((b >> 16) & 0xffu) + (a >> 16) * (b >> 24); }
SVS_HD uint32_t dp4a(uint32_t a, uint32_t b, uint32_t c)
{
    for (int k = 0; k < 4; ++k) c += ((a >> (8 * k)) & 0xffu) * ((b >> (8 * k)) & 0xffu);
    return c;
}
SVS_HD uint32_t absdiff4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) {
        const int d = (int)((a >> (8 * k)) & 0xffu) - (int)((b >> (8 * k)) & 0xffu);
        r |= (uint32_t)(d < 0 ? -d : d) << (8 * k);
    }
    return r;
}
SVS_HD int f2i_rn(float v) { return (int)std::nearbyintf(v); }
SVS_HD float i2f(int v) { return (float)v; }
SVS_HD uint32_t to_u8(float v) { return v != v ? 0u : (v <= 0.0f ? 0u : (v >= 255.0f ? 255u : (uint32_t)v)); }
SVS_HD uint32_t pack4_u8(float x0, float x1, float x2, float x3) { return to_u8(x0) | (to_u8(x1) << 8) | (to_u8(x2) << 16) | (to_u8(x3) << 24); }
SVS_HD uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }
#endif

SVS_HD float lo_of(P2 p) { float a, b; unpkf(p, a, b); return a; }
SVS_HD float hi_of(P2 p) { float a, b; unpkf(p, a, b); return b; }
SVS_HD uint32_t bswap(uint32_t v) { return byte_perm(v, 0, 0x0123); }

// Arithmetic policy for svs_math.cuh on packed pairs.
//
// A product must round on its own before it is added to anything, but ptxas contracts
// mul.f32x2 + add.f32x2 into one FFMA2 regardless of .rn and --fmad=false.  So products are
// written fma(a, c, +0.0): the addend is the zero register (RZ) - no register-file read, which
// matters because operand delivery is what these kernels are bound by (profiles/rf_model.py) -
// and since fma(a, c, +0.0) is NOT the same function as a * c (it returns +0 where the product
// is -0) the compiler can neither turn it back into a multiply nor fuse an addition into it.
// The one difference, the sign of an exact zero, cannot reach a result: a zero of either sign
// added to a non-zero value, quantised (fma(+-0, r, k) = k), divided (rint(+-0 / d) = 0) or
// converted to a byte gives the same bits.  tests/test_block_host.py runs this very arithmetic
// against the oracle on the CPU, the -m gpu tests against the scalar kernels' true multiplies.
struct PackedOps {
    typedef P2 T;
    SVS_HDM T add(T a, T b) const { return add2(a, b); }
    SVS_HDM T sub(T a, T b) const { return sub2(a, b); }
    SVS_HDM T mulc(T a, float c) const { return fma2(a, pk(c, c), pku(0u, 0u)); }
    SVS_HDM T cst(float c) const { return pk(c, c); }
    // sum = a + b, diff = a - b back to back (see ScalarOps::bfly)
    SVS_HDM void bfly(T a, T b, T& sum, T& diff) const
    {
#if defined(__CUDA_ARCH__)
        asm("add.rn.f32x2 %0, %2, %3;\n\tsub.rn.f32x2 %1, %2, %3;" : "=&l"(sum.v), "=&l"(diff.v) : "l"(a.v), "l"(b.v));
#else
        sum = add2(a, b);
        diff = sub2(a, b);
#endif
    }
};

}  // namespace hw
