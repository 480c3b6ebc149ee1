// svs_block.cuh - the throughput kernels of the DCT-QIM path (included by svs_b200.cu):
// ONE 8x8 block per thread, packed FP32 inside the block.
//
// Same results as the scalar kernels in svs_b200.cu, bit for bit (reference:
// proses_frame_qim_dct, config_and_setup.py:106-174).  What the organisation is built on
// (profiles/microbench/pipes_b200.txt: FP32 32 lanes/clk/SMSP, FADD2/FFMA2 do two lanes per
// issue slot, the ALU pipe - LOP3/PRMT/SHF - is half rate, F2I quarter):
//   * both halves of a 64-bit register pair belong to the SAME block: two neighbouring columns
//     during a column pass ("column pairs", c[r*4+j] = x[r][2j], x[r][2j+1]) and two
//     neighbouring rows during a row pass ("row pairs", q[i*8+v] = x[2i][v], x[2i+1][v]), so
//     all four passes of embed run on FADD2/FFMA2;
//   * the regrouping between the two layouts is a 2x2 transposition of registers.  It costs
//     nothing on the FP32 pipe: the first butterfly stage of either transform reads each of
//     its 8 inputs exactly once (svs_math.cuh: dct8_*_head), so that stage runs on SCALAR
//     FADD/FMUL (one lane per instruction, same pipe cycles per lane as the packed form), reads
//     the halves where the previous pass left them and writes where the packed tail wants them;
//   * a thread therefore owns 32 register pairs instead of 64: half the registers of the
//     two-blocks-per-thread organisation of round 1 (variants/svs_fast.cuh), twice the resident
//     warps, and a loop body of ~2 k instructions that fits the 32 KB instruction cache - the
//     warps run free (no lockstep barrier) and their FP32-heavy transform phases overlap other
//     warps' ALU-heavy quantiser / conversion phases.
// Everything below the kernels is written on the operations of svs_hw.cuh, which also have a
// plain C++ body: tests/host_math runs block_embed / block_extract on the CPU against the oracle.
//
// Only whole frames that the payload fills completely come here (k == n for every block); the
// frame in which the payload ends, strided/unaligned inputs and non-float32 deltas are handled
// by the scalar kernels.
#pragma once

#include "svs_hw.cuh"
#include "svs_quant.h"

namespace blk {

using hw::P2;
using hw::PackedOps;
using svs::FastQuant;
using svs::ScalarOps;
using svs::Stage1;

// out-of-line on the device (instruction-cache space), a plain static function for the host tests
#if defined(__CUDACC__)
#define SVS_RARE __host__ __device__ __noinline__
#else
#define SVS_RARE static
#endif

constexpr uint32_t kZone = svs::kQuantZone;        // flagged when the fraction field is < kZone
constexpr float kRintMagic = 12582912.0f;          // 1.5 * 2^23

// n / d for 0 <= n < 2^31 as one 32x32->64 multiply and one shift (host: make_div)
struct Div {
    uint32_t mul, shift;
};
SVS_HD uint32_t div_by(uint32_t n, Div d) { return (uint32_t)(((unsigned long long)n * d.mul) >> d.shift); }

// what the quantiser needs, in registers
struct QuantRegs {
    P2 r2, ke, d2, k0;                  // embed (svs_quant.h)
    P2 rr, kx;                          // extract
    uint32_t emask, ebit, xmask;
    int erot, xk;
    float d, r;                         // delta32, RN(1/delta32)
    float r2s, kes, kxs;                // scalar copies for the repair paths
};
SVS_HD QuantRegs make_quant_regs(const FastQuant& q, float delta32)
{
    QuantRegs c;
    c.r2 = hw::pk(q.r2, q.r2); c.ke = hw::pk(q.ke, q.ke); c.d2 = hw::pk(q.d2, q.d2); c.k0 = hw::pk(q.k0, q.k0);
    c.rr = hw::pk(q.r, q.r); c.kx = hw::pk(q.kx, q.kx);
    c.emask = q.emask; c.ebit = q.ebit; c.xmask = q.xmask;
    c.erot = q.erot; c.xk = q.xk;
    c.d = delta32; c.r = q.r;
    c.r2s = q.r2; c.kes = q.ke; c.kxs = q.kx;
    return c;
}

// ------------------------------------------------------------------------------------------
// input: one image row of the block -> 4 column pairs of exact floats
// ------------------------------------------------------------------------------------------
// (2^23 + 256 * byte SEL of v) as float bits: one PRMT  [0x00, v.bSEL, 0x00, 0x4B].  The axis-0
// pass works on these directly (svs_math.cuh: dct8_fwd_tail_impl<true>), no conversion arithmetic.
SVS_HD uint32_t magic_byte(uint32_t v, uint32_t magic_hi, int sel) { return hw::byte_perm(v, magic_hi, 0x7504u | ((uint32_t)sel << 4)); }

// BGR -> gray of the 8 pixels of a row held in six words: cv2 BGR2GRAY,
// (3735 B + 19235 G + 9798 R + 16384) >> 15 (config_and_setup.py:112), computed as two dp2a per
// pixel with doubled weights so that the gray value is bits 16..23 of the sum s[px]; the 16-bit
// weights are placed according to where the pixel's three bytes sit in the words (no shuffles).
SVS_HD void bgr_row_sums(const uint32_t (&v)[6], uint32_t (&s)[8])
{
    constexpr uint32_t WB = 7470u, WG = 38470u, WR = 19596u, RND = 32768u;
#pragma unroll
    for (int px = 0; px < 8; ++px) {
        const int byte0 = 3 * px, wi = byte0 >> 2, off = byte0 & 3;
        if (off == 0)      s[px] = hw::dp2a_hi(WR, v[wi], hw::dp2a_lo((WG << 16) | WB, v[wi], RND));
        else if (off == 1) s[px] = hw::dp2a_hi((WR << 16) | WG, v[wi], hw::dp2a_lo(WB << 16, v[wi], RND));
        else if (off == 2) s[px] = hw::dp2a_lo(WR, v[wi + 1], hw::dp2a_hi((WG << 16) | WB, v[wi], RND));
        else               s[px] = hw::dp2a_lo((WR << 16) | WG, v[wi + 1], hw::dp2a_hi(WB << 16, v[wi], RND));
    }
}

// 8 BGR pixels in six words -> their 8 gray bytes in two words
SVS_HD void row_gray_words(const uint32_t* w, uint32_t& glo, uint32_t& ghi)
{
    const uint32_t v[6] = {w[0], w[1], w[2], w[3], w[4], w[5]};
    uint32_t s[8];
    bgr_row_sums(v, s);
    glo = hw::byte_perm(hw::byte_perm(s[0], s[1], 0x0062u), hw::byte_perm(s[2], s[3], 0x0062u), 0x5410u);
    ghi = hw::byte_perm(hw::byte_perm(s[4], s[5], 0x0062u), hw::byte_perm(s[6], s[7], 0x0062u), 0x5410u);
}

// CH == 1: w[0..1] are the 8 gray bytes; CH == 3: w[0..5] are the 24 BGR bytes.
// c[0..3] = column pairs of the row as 2^23 + 256 * gray; glo/ghi = the row's gray bytes (only
// built when WANT_GRAY).
template <int CH, bool WANT_GRAY>
SVS_HD void row_to_pairs(const uint32_t* w, uint32_t magic_hi, P2* c, uint32_t& glo, uint32_t& ghi)
{
    if (CH == 1) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t a = magic_byte(w[(2 * j) >> 2], magic_hi, (2 * j) & 3);
            const uint32_t b = magic_byte(w[(2 * j + 1) >> 2], magic_hi, (2 * j + 1) & 3);
            c[j] = hw::pku(a, b);
        }
        if (WANT_GRAY) { glo = w[0]; ghi = w[1]; }
    } else {
        const uint32_t v[6] = {w[0], w[1], w[2], w[3], w[4], w[5]};
        uint32_t s[8];
        bgr_row_sums(v, s);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            c[j] = hw::pku(hw::byte_perm(s[2 * j], magic_hi, 0x7524u), hw::byte_perm(s[2 * j + 1], magic_hi, 0x7524u));
        if (WANT_GRAY) {
            glo = hw::byte_perm(hw::byte_perm(s[0], s[1], 0x0062u), hw::byte_perm(s[2], s[3], 0x0062u), 0x5410u);
            ghi = hw::byte_perm(hw::byte_perm(s[4], s[5], 0x0062u), hw::byte_perm(s[6], s[7], 0x0062u), 0x5410u);
        }
    }
}

// ------------------------------------------------------------------------------------------
// the four passes
// ------------------------------------------------------------------------------------------
// axis 0, forward: four packed transforms down the column pairs, in place.  In: 2^23 + 256 *
// pixel (row_to_pairs); out: the coefficients of the axis-0 pass, as the reference has them.
SVS_HD void columns_fwd(const PackedOps& po, P2 (&c)[32])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        P2 X[8];
        svs::dct8_fwd_tail_impl<true>(po, svs::dct8_fwd_head(po, c[j], c[4 + j], c[8 + j], c[12 + j], c[16 + j], c[20 + j], c[24 + j], c[28 + j]), X);
#pragma unroll
        for (int u = 0; u < 8; ++u) c[4 * u + j] = X[u];
    }
}

// Only the first 2*NP coefficient rows of axis 0 (extraction with few coefficients per block:
// the later rows are never read, config_and_setup.py:138-140); ptxas drops the unused outputs
// and what feeds only them.
template <int NP>
SVS_HD void columns_fwd_pruned(const PackedOps& po, P2 (&c)[32])
{
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        P2 X[8];
        svs::dct8_fwd_tail_impl<true>(po, svs::dct8_fwd_head(po, c[j], c[4 + j], c[8 + j], c[12 + j], c[16 + j], c[20 + j], c[24 + j], c[28 + j]), X);
#pragma unroll
        for (int u = 0; u < 2 * NP; ++u) c[4 * u + j] = X[u];
    }
}

// rows 2i and 2i+1 of the column-pair layout (ra, rb: 4 pairs each) through scalar stage 1 ->
// the eight stage-1 values of both rows, packed (lo = row 2i, hi = row 2i+1)
template <bool INVERSE>
SVS_HD Stage1<P2> regroup_rows(const ScalarOps& so, const P2* ra, const P2* rb)
{
    float a[8], b[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        hw::unpkf(ra[j], a[2 * j], a[2 * j + 1]);
        hw::unpkf(rb[j], b[2 * j], b[2 * j + 1]);
    }
    const Stage1<float> ha = INVERSE ? svs::dct8_inv_head(so, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7])
                                     : svs::dct8_fwd_head(so, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
    const Stage1<float> hb = INVERSE ? svs::dct8_inv_head(so, b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7])
                                     : svs::dct8_fwd_head(so, b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7]);
    Stage1<P2> h;
#pragma unroll
    for (int k = 0; k < 8; ++k) h.v[k] = hw::pk(ha.v[k], hb.v[k]);
    return h;
}

// axis 1, forward, rows 2i and 2i+1: X[v] = (coefficient (2i, v), coefficient (2i+1, v))
SVS_HD void rows_fwd_pair(const PackedOps& po, const ScalarOps& so, const P2* c8, P2 (&X)[8])
{
    svs::dct8_fwd_tail(po, regroup_rows<false>(so, c8, c8 + 4), X);
}

// axis 0, inverse, columns 2j and 2j+1: reads the row-pair layout q, writes column pairs c[r*4+j]
SVS_HD void columns_inv_pair(const PackedOps& po, const ScalarOps& so, const P2 (&q)[32], int j, P2 (&c)[32])
{
    float a[8], b[8];                         // coefficient columns 2j and 2j+1, u = 0..7
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        hw::unpkf(q[8 * i + 2 * j], a[2 * i], a[2 * i + 1]);
        hw::unpkf(q[8 * i + 2 * j + 1], b[2 * i], b[2 * i + 1]);
    }
    const Stage1<float> ha = svs::dct8_inv_head(so, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7]);
    const Stage1<float> hb = svs::dct8_inv_head(so, b[0], b[1], b[2], b[3], b[4], b[5], b[6], b[7]);
    Stage1<P2> h;
#pragma unroll
    for (int k = 0; k < 8; ++k) h.v[k] = hw::pk(ha.v[k], hb.v[k]);
    P2 x[8];
    svs::dct8_inv_tail(po, h, x);
#pragma unroll
    for (int r = 0; r < 8; ++r) c[4 * r + j] = x[r];
}

// axis 1, inverse, rows 2i and 2i+1: out[col] = (pixel (2i, col), pixel (2i+1, col))
SVS_HD void rows_inv_pair(const PackedOps& po, const ScalarOps& so, const P2* c8, P2 (&out)[8])
{
    svs::dct8_inv_tail(po, regroup_rows<true>(so, c8, c8 + 4), out);
}

// ------------------------------------------------------------------------------------------
// quantiser (svs_quant.h states the arithmetic and its error budget)
// ------------------------------------------------------------------------------------------
// IEEE-exact c / d from the correctly rounded reciprocal r = RN(1/d): two Newton corrections on
// the quotient (the sequence __fdiv_rn runs after its own reciprocal; all operands are normal
// here).  Checked against the FPU division in tests/test_host_logic.py.
SVS_HD float div_exact(float c, float d, float r)
{
    float q = hw::fmul(c, r);
    q = hw::ffma(hw::ffma(-q, d, c), r, q);
    return hw::ffma(hw::ffma(-q, d, c), r, q);
}

// Coefficients (0,4), (4,0) and (4,4) of a block of integer pixels are exact multiples of 1/8
// (their basis is +-1/8), so c / delta lands EXACTLY on a rounding tie in one block out of
// 8 delta - far too often for the speculate-and-repair quantiser.  They are always quantised
// with the IEEE-exact quotient, on scalar registers (all three sit in the low half of their
// row pair), rint() through the 1.5 * 2^23 constant: round-half-even, the parity is the lowest
// mantissa bit (two's complement for negative quotients).
SVS_HD bool tie_prone(int flat) { return flat == 4 || flat == 32 || flat == 36; }

SVS_HD float exact_embed(float c, uint32_t bit, const QuantRegs& Q)         // config_and_setup.py:148-156
{
    const float m = hw::fadd(div_exact(c, Q.d, Q.r), kRintMagic);
    const float adj = hw::i2f((int)bit - (int)(hw::f2u(m) & 1u));
    return hw::fmul(hw::fadd(hw::fsub(m, kRintMagic), adj), Q.d);
}
SVS_HD uint32_t exact_parity(float c, const QuantRegs& Q)                   // config_and_setup.py:160-161
{
    return hw::f2u(hw::fadd(div_exact(c, Q.d, Q.r), kRintMagic)) & 1u;
}

// payload bit `idx` (0 = first) of the block's 64-bit window (w0 = bits 0..31, MSB first)
SVS_HD uint32_t window_bit(uint32_t w0, uint32_t w1, int idx) { return ((idx < 32 ? w0 : w1) >> (31 - (idx & 31))) & 1u; }

// Rare path, deliberately out of line and looped so that it costs almost no instruction-cache
// space.  `orig` holds the 8 coefficient pairs of rows 2i / 2i+1 before quantisation, `res`
// the results of the division-free quantiser; every coefficient whose fraction was too close
// to a rounding boundary is recomputed as the scalar kernels do (IEEE division, round-half-
// even, float32 product) and patched into `res`.  Scalars only (an array indexed by the lane
// would live in local memory): a call costs 8 loads of `orig` and touches `res` only where a
// coefficient is actually flagged.
// Payload bit numbers of the two halves of pair v: base + v and base + v + hi_off.
SVS_RARE void fix_pair_embed(const P2* orig, P2* res, int base, int hi_off, int n, float d, float r, float r2, float ke,
                    uint32_t emask, uint32_t w0, uint32_t w1)
{
#pragma unroll 1
    for (int v = 0; v < 8; ++v) {
        float ca, cb;
        hw::unpkf(orig[v], ca, cb);
        const int ia = base + v, ib = ia + hi_off;
        const bool fa = ia >= 0 && ia < n && (hw::f2u(hw::ffma(ca, r2, ke)) & emask) < kZone;
        const bool fb = ib < n && (hw::f2u(hw::ffma(cb, r2, ke)) & emask) < kZone;
        if (!(fa || fb)) continue;
        float oa, ob;
        hw::unpkf(res[v], oa, ob);
        if (fa) {
            const int q = hw::f2i_rn(div_exact(ca, d, r));
            oa = hw::fmul(hw::i2f(q - (q & 1) + (int)window_bit(w0, w1, ia)), d);
        }
        if (fb) {
            const int q = hw::f2i_rn(div_exact(cb, d, r));
            ob = hw::fmul(hw::i2f(q - (q & 1) + (int)window_bit(w0, w1, ib)), d);
        }
        res[v] = hw::pk(oa, ob);
    }
}

// Same for extraction: `rows` = (row 2i byte | row 2i+1 byte << 16), coefficient v at bit 7-v;
// flagged coefficients get their parity from the exact quotient.
SVS_RARE uint32_t fix_pair_extract(const P2* in, uint32_t rows, int i, int n, float d, float r, float kx, uint32_t xmask)
{
#pragma unroll 1
    for (int v = 0; v < 8; ++v) {
        float ca, cb;
        hw::unpkf(in[v], ca, cb);
        const int ia = 16 * i + v - 1, ib = ia + 8;
        if (ia >= 0 && ia < n && (hw::f2u(hw::ffma(ca, r, kx)) & xmask) < kZone) {
            const uint32_t par = (uint32_t)hw::f2i_rn(div_exact(ca, d, r)) & 1u;
            rows = (rows & ~(0x80u >> v)) | (par << (7 - v));
        }
        if (ib < n && (hw::f2u(hw::ffma(cb, r, kx)) & xmask) < kZone) {
            const uint32_t par = (uint32_t)hw::f2i_rn(div_exact(cb, d, r)) & 1u;
            rows = (rows & ~(0x800000u >> v)) | (par << (23 - v));
        }
    }
    return rows;
}

// Embed into rows 2i, 2i+1 (X in, quantised coefficients out, in place).  Payload bit idx =
// 8u + v - 1 goes to coefficient (u, v): row-major flat index 1..n (config_and_setup.py:138-141).
// p0/p1 are the window words pre-rotated so that bit idx sits `idx` places below position
// erot: bringing it there is a rotate by the compile-time constant idx.
template <bool NFULL>
SVS_HD void quant_embed_pair(int i, P2 (&X)[8], const QuantRegs& Q, int n, uint32_t w0, uint32_t w1, uint32_t p0, uint32_t p1)
{
    // halves of pair v: coefficient rows 2i and 2i+1
    const int base = 16 * i - 1, hi_off = 8;
    uint32_t worst = 0xffffffffu;
    P2 nx[8];
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const int il = base + v, ih = il + hi_off;              // payload bit numbers of the two halves
        const bool tl = tie_prone(il + 1), th = tie_prone(ih + 1);
        const bool al = il >= 0 && (NFULL || il < n), ah = NFULL || ih < n;
        const P2 y = hw::fma2(X[v], Q.r2, Q.ke);
        uint32_t ya, yb;
        hw::unpk(y, ya, yb);
        if (al && !tl) worst = hw::umin(worst, ya & Q.emask);
        if (ah && !th) worst = hw::umin(worst, yb & Q.emask);
        const int jl = il < 0 ? 0 : il;
        const uint32_t ta = hw::funnel_l(jl < 32 ? p0 : p1, jl < 32 ? p0 : p1, jl & 31) & Q.ebit;
        const uint32_t tb = hw::funnel_l(ih < 32 ? p0 : p1, ih < 32 ? p0 : p1, ih & 31) & Q.ebit;
        // M + floor() + bit/2, then (2e + bit) * delta in one rounding
        const P2 z = hw::fma2(hw::pku((ya & ~Q.emask) | ta, (yb & ~Q.emask) | tb), Q.d2, Q.k0);
        if (NFULL && !tl && !th && il >= 0) {
            nx[v] = z;
        } else {
            float lo = hw::lo_of(z), hi = hw::hi_of(z);
            if (tl) lo = exact_embed(hw::lo_of(X[v]), window_bit(w0, w1, jl), Q);
            if (th) hi = exact_embed(hw::hi_of(X[v]), window_bit(w0, w1, ih), Q);
            if (!al) lo = hw::lo_of(X[v]);
            if (!ah) hi = hw::hi_of(X[v]);
            nx[v] = hw::pk(lo, hi);
        }
    }
    if (worst < kZone) {                                         // rare: a fraction too close to call
        // (copies made HERE: arrays whose address escapes live in local memory, and X / nx must not)
        P2 orig[8], res[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) { orig[v] = X[v]; res[v] = nx[v]; }
        fix_pair_embed(orig, res, base, hi_off, NFULL ? 63 : n, Q.d, Q.r, Q.r2s, Q.kes, Q.emask, w0, w1);
#pragma unroll
        for (int v = 0; v < 8; ++v) nx[v] = res[v];
    }
#pragma unroll
    for (int v = 0; v < 8; ++v) X[v] = nx[v];
}

// Parities of rows 2i, 2i+1: returns (row 2i byte | row 2i+1 byte << 16), coefficient v at bit 7-v.
SVS_HD uint32_t quant_extract_pair(int i, const P2 (&X)[8], const QuantRegs& Q, int n)
{
    uint32_t worst = 0xffffffffu, rl = 0, rh = 0;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const bool tie = tie_prone(16 * i + v);
        const bool dc = i == 0 && v == 0;
        const P2 y = hw::fma2(X[v], Q.rr, Q.kx);
        uint32_t ya, yb;
        hw::unpk(y, ya, yb);
        const int rot = (7 - v - Q.xk) & 31;                     // parity (bit xk) -> bit 7-v
        if (tie) {
            rl |= exact_parity(hw::lo_of(X[v]), Q) << (7 - v);
        } else if (!dc) {
            worst = hw::umin(worst, ya & Q.xmask);
            rl |= hw::funnel_l(ya, ya, rot) & (0x80u >> v);
        }
        worst = hw::umin(worst, yb & Q.xmask);
        rh |= hw::funnel_l(yb, yb, rot) & (0x80u >> v);
    }
    uint32_t rows = rl | (rh << 16);
    if (worst < kZone) {
        P2 in[8];
#pragma unroll
        for (int v = 0; v < 8; ++v) in[v] = X[v];
        rows = fix_pair_extract(in, rows, i, n, Q.d, Q.r, Q.kxs, Q.xmask);
    }
    return rows;
}

// byte of coefficient row u (bit 7-v = coefficient v) -> stream bits 8u-1 .. 8u+6 of the block's
// 64-bit string (hi:lo), stream bit s at bit 63-s; the DC position falls off the top
SVS_HD void place_row(int u, uint32_t byte, uint32_t& hi, uint32_t& lo)
{
    if (u < 4)       hi |= byte << (25 - 8 * u);
    else if (u == 4) { hi |= byte >> 7; lo |= byte << 25; }
    else             lo |= byte << (57 - 8 * u);
}

// ------------------------------------------------------------------------------------------
// whole blocks (shared by the kernels and by tests/host_math)
// ------------------------------------------------------------------------------------------
// rows: 8 x (CH == 3 ? 6 : 2) words of the block's image rows -> column pairs c (and the
// gray bytes, row r at [2r], [2r+1], when WANT_GRAY)
template <int CH, bool WANT_GRAY>
SVS_HD void block_input(const uint32_t* rows, uint32_t magic_hi, P2 (&c)[32], uint32_t* gray)
{
    constexpr int P = CH == 3 ? 6 : 2;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        uint32_t glo = 0, ghi = 0;
        row_to_pairs<CH, WANT_GRAY>(rows + r * P, magic_hi, c + 4 * r, glo, ghi);
        if (WANT_GRAY) { gray[2 * r] = glo; gray[2 * r + 1] = ghi; }
    }
}

// Forward 2-D transform of the block (column pairs c, destroyed) and embedding of its n payload
// bits: q = the quantised coefficients as row pairs.  Every block coming here takes all n bits.
template <bool NFULL>
SVS_HD void block_forward_quant(P2 (&c)[32], const QuantRegs& Q, int n, uint32_t w0, uint32_t w1, P2 (&q)[32])
{
    const PackedOps po;
    const ScalarOps so;
    columns_fwd(po, c);
    const int pre = (Q.erot - 31) & 31;
    const uint32_t p0 = hw::funnel_l(w0, w0, pre), p1 = hw::funnel_l(w1, w1, pre);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        P2 X[8];
        rows_fwd_pair(po, so, c + 8 * i, X);                               // rows 2i and 2i+1
        if (NFULL || 16 * i - 1 < n) quant_embed_pair<NFULL>(i, X, Q, n, w0, w1, p0, p1);
#pragma unroll
        for (int v = 0; v < 8; ++v) q[8 * i + v] = X[v];
    }
}

// Inverse 2-D transform of the row pairs q (destroyed) and conversion to bytes:
// stego = 16 words, row r at [2r], [2r+1].
SVS_HD void block_inverse(P2 (&q)[32], uint32_t* stego)
{
    const PackedOps po;
    const ScalarOps so;
    P2 c[32];
#pragma unroll
    for (int j = 0; j < 4; ++j) columns_inv_pair(po, so, q, j, c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        P2 o[8];
        rows_inv_pair(po, so, c + 8 * i, o);
        // clip then truncate (config_and_setup.py:171); low halves = row 2i, high halves = row 2i+1
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            stego[4 * i + h] = hw::pack4_u8(hw::lo_of(o[4 * h]), hw::lo_of(o[4 * h + 1]), hw::lo_of(o[4 * h + 2]), hw::lo_of(o[4 * h + 3]));
            stego[4 * i + 2 + h] = hw::pack4_u8(hw::hi_of(o[4 * h]), hw::hi_of(o[4 * h + 1]), hw::hi_of(o[4 * h + 2]), hw::hi_of(o[4 * h + 3]));
        }
    }
}

template <bool NFULL>
SVS_HD void block_embed_pairs(P2 (&c)[32], const QuantRegs& Q, int n, uint32_t w0, uint32_t w1, uint32_t* stego)
{
    P2 q[32];
    block_forward_quant<NFULL>(c, Q, n, w0, w1, q);
    block_inverse(q, stego);
}

template <int CH, bool NFULL, bool WANT_GRAY>
SVS_HD void block_embed(const uint32_t* rows, uint32_t magic_hi, const QuantRegs& Q, int n, uint32_t w0, uint32_t w1,
                        uint32_t* stego, uint32_t* gray)
{
    P2 c[32];
    block_input<CH, WANT_GRAY>(rows, magic_hi, c, gray);
    block_embed_pairs<NFULL>(c, Q, n, w0, w1, stego);
}

// NP = number of coefficient row pairs that hold any of the n coefficients: ceil((n + 1) / 16).
// c: the block as column pairs (destroyed).  Returns the block's n parity bits as a 64-bit
// string (hi:lo), stream bit s at bit 63-s.
template <int NP>
SVS_HD void block_extract_pairs(P2 (&c)[32], const QuantRegs& Q, int n, uint32_t& hi, uint32_t& lo)
{
    const PackedOps po;
    const ScalarOps so;
    if (NP == 4) columns_fwd(po, c);
    else columns_fwd_pruned<NP>(po, c);
    hi = 0;
    lo = 0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        P2 X[8];
        rows_fwd_pair(po, so, c + 8 * i, X);
        const uint32_t rows2 = quant_extract_pair(i, X, Q, n);
        place_row(2 * i, rows2 & 0xffu, hi, lo);
        place_row(2 * i + 1, rows2 >> 16, hi, lo);
    }
    // bits n.. of the string are not part of the stream (config_and_setup.py:138-140)
    if (n < 32) { hi &= ~(0xffffffffu >> n); lo = 0; }
    else if (n < 64) lo &= n == 32 ? 0u : ~(0xffffffffu >> (n - 32));
}

template <int CH, int NP>
SVS_HD void block_extract(const uint32_t* rows, uint32_t magic_hi, const QuantRegs& Q, int n, uint32_t& hi, uint32_t& lo)
{
    P2 c[32];
    block_input<CH, false>(rows, magic_hi, c, nullptr);
    block_extract_pairs<NP>(c, Q, n, hi, lo);
}

#if defined(__CUDACC__)
// ==========================================================================================
// kernels
// ==========================================================================================
// ONE CTA of 16 warps per SM.  Two CTAs of 8 warps hold the same 16 warps, but the warp scheduler
// prefers the older CTA: it finished its share at ~60 % of the kernel time and the younger one ran
// the rest alone (ncu: 3.2 of 4 warp slots per sub-partition occupied on average, 4.0 with one
// CTA; embed 1.57 -> 1.50, extract 0.82 -> 0.79 ms per 600 frames).
#ifndef SVS_BLK_THREADS
#define SVS_BLK_THREADS 512
#endif
#ifndef SVS_BLK_MIN_CTAS
#define SVS_BLK_MIN_CTAS 1
#endif
constexpr int kBlkThreads = SVS_BLK_THREADS;
constexpr int kBlkMinCtas = SVS_BLK_MIN_CTAS;
constexpr int kBlkWarps = kBlkThreads / 32;
constexpr int kMaxPeers = 15;

// How the rows of a group reach the registers: the NEXT group's rows are requested as soon as the
// current ones have been converted to floats, with cp.async into a thread-private shared-memory
// slot (BGR input: 48 words, gray input: 16); they arrive while the current group is being
// transformed.  (Measured against LDG at the top of the group, with and without
// prefetch.global.L2 of the next group: profiles/README.md.)  Every instantiation stages through
// shared memory: a load in flight INTO REGISTERS across the quantiser would make its rare
// out-of-line repair call wait for HBM (the callee saves the registers the load is going to
// write) - measured: 8 % of the warp time.
// Dynamic shared memory of a kernel instantiation: the cp.async slots, 8 bytes x rows x words x threads.
template <int CH, bool EMBED>
__host__ __device__ constexpr int blk_smem_bytes() { return 8 * (CH == 3 ? 3 : 1) * 8 * kBlkThreads; }
// SIDE embed kernels park the block's 16 gray words in shared memory between the input stage and
// the epilogue (16 registers the transforms need): word k of thread t at (k * kBlkThreads + t) * 4
template <int CH, bool SIDE>
__host__ __device__ constexpr int blk_embed_smem_bytes() { return blk_smem_bytes<CH, true>() + (SIDE ? 16 * 4 * kBlkThreads : 0); }

// A warp owns 32 consecutive blocks of a frame in raster order ("group"): BGR rows arrive as
// three 8-byte accesses per lane over one 768-byte contiguous span, stego rows leave as
// 256-byte STG.64 runs, and the 32 n extracted bits form n 4-byte aligned words.
struct BlkGeom {
    const uint8_t* frames;
    long long frame_stride, row_stride;
    uint32_t frame_stride32, row_stride32;        // the same as 32-bit values (the launcher checks that they fit)
    int bw, bpf, n;
    int gpf;                                      // groups per frame = ceil(bpf / 32)
    long long total_groups;                       // n_frames * gpf, < 2^31
    Div by_gpf, by_bw;
    uint32_t magic_hi;                            // 0x4B000000, opaque so that it stays in a register
    float delta32;
};

struct BlkEmbedArgs {
    BlkGeom g;
    FastQuant q;
    const uint32_t* payload;
    long long payload_bit_offset, payload_last_word, cap;
    uint8_t* stego;
    long long stego_frame_stride, stego_row_stride;
    uint32_t stego_frame_stride32, stego_row_stride32;
    int64_t* bits_embedded;
    uint8_t* gray;                                // nullable; contiguous H x W frames   (SIDE kernels only)
    unsigned long long* sse;                      // nullable; per-frame sum of squares  (SIDE kernels only)
    long long gray_frame_stride;                  // H * W
    int W;
};

struct BlkExtractArgs {
    BlkGeom g;
    FastQuant q;
    uint8_t* bits;
    long long bits_frame_stride;
    // fused all-gather: the same rows are also stored to these (peer-mapped, NVLink) buffers;
    // multicast != 0: peers[0] is ONE NVSwitch multicast (multimem) address that reaches every
    // rank including this one, and `bits` is not written separately
    uint8_t* peers[kMaxPeers];
    int n_peers;
    int multicast;
};

struct Where {
    int f, base, by, bx;
    bool ok;
};
__device__ __forceinline__ Where locate(const BlkGeom& G, long long g, int lane)
{
    Where w;
    const uint32_t gi = (uint32_t)g;
    w.f = (int)div_by(gi, G.by_gpf);
    w.base = (int)(gi - (uint32_t)w.f * (uint32_t)G.gpf) * 32;
    w.ok = w.base + lane < G.bpf;
    const int b = min(w.base + lane, G.bpf - 1);
    w.by = (int)div_by((uint32_t)b, G.by_bw);
    w.bx = b - w.by * G.bw;
    return w;
}

// 32-bit addressing inside a frame: a frame spans less than 4 GB (svs_b200.cu: blk_addressable), so
// the offset of a block inside its frame and the offsets of its 8 rows are 32-bit values (the row
// offsets r * stride are warp-uniform and loop-invariant: uniform registers) - one 64-bit add per
// row instead of two, no 64-bit multiplies.
template <typename T>
__device__ __forceinline__ T* frame_ptr(T* base, int f, uint32_t frame_stride, uint32_t in_frame)
{
    return base + ((unsigned long long)(uint32_t)f * frame_stride + in_frame);
}
template <int CH>
__device__ __forceinline__ const uint8_t* block_src(const BlkGeom& G, const Where& w)
{
    return frame_ptr(G.frames, w.f, G.frame_stride32, (uint32_t)(w.by * 8) * G.row_stride32 + (uint32_t)(w.bx * (8 * CH)));
}
template <typename T>
__device__ __forceinline__ T* row_ptr(T* p, int r, uint32_t stride32)
{
    return p + (uint32_t)r * stride32;
}

// cp.async staging: slot (r, j) of thread t is 8 bytes at ((r*P + j) * kBlkThreads + t) * 8 of the
// dynamic shared memory - thread-private, conflict-free, no barrier needed.  The slot and word
// offsets are compile-time immediates of the instructions (no address arithmetic per access).
template <int SLOT_OFF, int SRC_OFF>
__device__ __forceinline__ void cp_async8(uint32_t slot0, const uint8_t* p)
{
    asm volatile("cp.async.ca.shared.global [%0 + %2], [%1 + %3], 8;" ::"r"(slot0), "l"(p), "n"(SLOT_OFF), "n"(SRC_OFF) : "memory");
}
template <int SLOT_OFF>
__device__ __forceinline__ void lds8(uint32_t slot0, uint32_t& lo, uint32_t& hi)
{
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2 + %3];" : "=r"(lo), "=r"(hi) : "r"(slot0), "n"(SLOT_OFF) : "memory");
}
template <int CH, int R>
__device__ __forceinline__ void stage_request_row(uint32_t slot0, const uint8_t* p)
{
    constexpr int P = CH == 3 ? 3 : 1;
    cp_async8<(R * P + 0) * kBlkThreads * 8, 0>(slot0, p);
    if (CH == 3) {
        cp_async8<(R * P + 1) * kBlkThreads * 8, 8>(slot0, p);
        cp_async8<(R * P + 2) * kBlkThreads * 8, 16>(slot0, p);
    }
}
template <int CH>
__device__ __forceinline__ void stage_request(const BlkGeom& G, const Where& w, uint32_t slot0)
{
    const uint8_t* p = block_src<CH>(G, w);
    stage_request_row<CH, 0>(slot0, p);
    stage_request_row<CH, 1>(slot0, row_ptr(p, 1, G.row_stride32));
    stage_request_row<CH, 2>(slot0, row_ptr(p, 2, G.row_stride32));
    stage_request_row<CH, 3>(slot0, row_ptr(p, 3, G.row_stride32));
    stage_request_row<CH, 4>(slot0, row_ptr(p, 4, G.row_stride32));
    stage_request_row<CH, 5>(slot0, row_ptr(p, 5, G.row_stride32));
    stage_request_row<CH, 6>(slot0, row_ptr(p, 6, G.row_stride32));
    stage_request_row<CH, 7>(slot0, row_ptr(p, 7, G.row_stride32));
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int CH, int K>
__device__ __forceinline__ void stage_fetch_from(uint32_t slot0, uint32_t* rows)
{
    constexpr int N = 8 * (CH == 3 ? 3 : 1);
    if constexpr (K < N) {
        lds8<K * kBlkThreads * 8>(slot0, rows[2 * K], rows[2 * K + 1]);
        stage_fetch_from<CH, K + 1>(slot0, rows);
    }
}
template <int CH>
__device__ __forceinline__ void stage_fetch(uint32_t slot0, uint32_t* rows)
{
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    stage_fetch_from<CH, 0>(slot0, rows);
}

// the three payload words that hold the block's 64-bit window, as loaded (big-endian bit order inside bytes)
struct PayWords {
    uint32_t a, b, c, s;
};
__device__ __forceinline__ PayWords payload_request(const uint32_t* __restrict__ words, long long last_word, long long pos)
{
    PayWords p;
    // word indices fit 32 bits (launcher), and the first word of a block the payload fills exists
    const uint32_t wi = (uint32_t)((unsigned long long)pos >> 5), left = (uint32_t)last_word - wi;
    const uint32_t* q = words + wi;
    p.s = (uint32_t)pos & 31u;
    p.a = __ldg(q);
    p.b = left >= 1u ? __ldg(q + 1) : 0u;
    p.c = left >= 2u ? __ldg(q + 2) : 0u;
    return p;
}
__device__ __forceinline__ void payload_window(const PayWords& p, uint32_t& w0, uint32_t& w1)
{
    const uint32_t a = hw::bswap(p.a), b = hw::bswap(p.b), c = hw::bswap(p.c);
    w0 = __funnelshift_l(b, a, p.s);
    w1 = __funnelshift_l(c, b, p.s);
}

__device__ __forceinline__ void stg64(uint8_t* p, uint32_t lo, uint32_t hi)
{
    asm volatile("st.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(lo), "r"(hi) : "memory");
}

template <int OUT_CH>
__device__ __forceinline__ void store_row(uint8_t* dst, uint32_t lo4, uint32_t hi4)
{
    if (OUT_CH == 1) {
        stg64(dst, lo4, hi4);
    } else {            // gray replicated to B,G,R (cv2.cvtColor GRAY2BGR, embed_process.py:126)
        stg64(dst, __byte_perm(lo4, 0, 0x1000), __byte_perm(lo4, 0, 0x2211));
        stg64(dst + 8, __byte_perm(lo4, 0, 0x3332), __byte_perm(hi4, 0, 0x1000));
        stg64(dst + 16, __byte_perm(hi4, 0, 0x2211), __byte_perm(hi4, 0, 0x3332));
    }
}

extern __shared__ __align__(16) unsigned char blk_dyn_smem[];

// SIDE: also the gray reference (first return value of the reference function,
// config_and_setup.py:111-114,172) and / or the per-frame sum of squared differences gray vs
// stego (what cv2.PSNR needs, embed_process.py:204-206), from the bytes the thread already holds.
template <int CH, int OUT_CH, bool NFULL, bool SIDE>
__global__ void __launch_bounds__(kBlkThreads, kBlkMinCtas) embed_blk_kernel(const BlkEmbedArgs a)
{
    const BlkGeom& G = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = NFULL ? 63 : G.n;
    QuantRegs Q = make_quant_regs(a.q, G.delta32);
    {   // FFMA2 takes one uniform-register operand: with k0 there, 2*delta has to sit in a register.
        // Left alone ptxas re-creates it (MOV R, UR) in front of every one of the 32 FFMA2 of the
        // quantiser; an opaque move makes it a value that can only be kept.
        __shared__ uint32_t pin;
        uint32_t d2;
        asm volatile("st.volatile.shared.u32 [%1], %2;\n\tld.volatile.shared.u32 %0, [%1];"
                     : "=r"(d2) : "r"((uint32_t)__cvta_generic_to_shared(&pin)), "r"(hw::f2u(a.q.d2)) : "memory");
        Q.d2 = hw::pku(d2, d2);
    }
    // Plain kernels: warp i takes groups i, i + #warps, ... (all warps sweep the batch together).
    // SIDE kernels: warp i takes the CONTIGUOUS range [i K, (i+1) K) so that consecutive groups of a
    // warp belong to the same frame and its squared-error sum stays in a register until the frame
    // changes - one atomic per warp and frame instead of one per group (1013 same-address atomics
    // per 1080p frame cost 25 % of the kernel).
    const long long n_warps = (long long)gridDim.x * kBlkWarps;
    const long long my_warp = (long long)blockIdx.x * kBlkWarps + warp;
    const long long chunk = (G.total_groups + n_warps - 1) / n_warps;
    const long long gstep = SIDE ? 1 : n_warps;
    long long g = SIDE ? my_warp * chunk : my_warp;
    const long long g_end = SIDE ? min(G.total_groups, g + chunk) : G.total_groups;
    if (g >= g_end) return;
    const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(blk_dyn_smem) + threadIdx.x * 8u;
    unsigned long long sse_acc = 0;                // this warp's sum for frame sse_frame (lane 0 holds the total)
    int sse_frame = -1;
    (void)sse_acc; (void)sse_frame;

    Where w = locate(G, g, lane);
    uint32_t rows[CH == 3 ? 48 : 16];
    PayWords pw;
    auto payload_pos = [&](const Where& x) {
        return a.payload_bit_offset + x.f * a.cap + (long long)min(x.base + lane, G.bpf - 1) * n;
    };
    stage_request<CH>(G, w, slot0);
    pw = payload_request(a.payload, a.payload_last_word, payload_pos(w));
    for (;;) {
        stage_fetch<CH>(slot0, rows);
        if (w.base + lane == 0 && a.bits_embedded != nullptr) a.bits_embedded[w.f] = a.cap;
        P2 c[32];
        uint32_t stego[16];
        uint32_t* const park = reinterpret_cast<uint32_t*>(blk_dyn_smem + blk_smem_bytes<CH, true>()) + threadIdx.x;
        {
            uint32_t gray[16];
            block_input<CH, SIDE>(rows, G.magic_hi, c, gray);
            if (SIDE) {
                if (a.gray != nullptr && w.ok) {
                    uint8_t* gd = frame_ptr(a.gray, w.f, (uint32_t)a.gray_frame_stride, (uint32_t)(w.by * 8) * (uint32_t)a.W + (uint32_t)(w.bx * 8));
#pragma unroll
                    for (int r = 0; r < 8; ++r) stg64(row_ptr(gd, r, (uint32_t)a.W), gray[2 * r], gray[2 * r + 1]);
                }
                if (a.sse != nullptr) {
#pragma unroll
                    for (int k = 0; k < 16; ++k) park[k * kBlkThreads] = gray[k];
                }
            }
        }
        uint32_t w0, w1;
        payload_window(pw, w0, w1);

        // the next group of this warp: its rows (and payload words) are requested now
        const long long gn = g + gstep;
        const bool more = gn < g_end;
        Where wn = w;
        if (more) {
            wn = locate(G, gn, lane);
            stage_request<CH>(G, wn, slot0);
        }

        P2 q[32];
        block_forward_quant<NFULL>(c, Q, n, w0, w1, q);
        // (only now: a load in flight across the quantiser would make its rare out-of-line repair
        // call wait for HBM - the callee saves the registers the load is going to write)
        if (more) pw = payload_request(a.payload, a.payload_last_word, payload_pos(wn));
        block_inverse(q, stego);
        if (w.ok) {
            uint8_t* dst = frame_ptr(a.stego, w.f, a.stego_frame_stride32,
                                     (uint32_t)(w.by * 8) * a.stego_row_stride32 + (uint32_t)(w.bx * (8 * OUT_CH)));
#pragma unroll
            for (int r = 0; r < 8; ++r)
                store_row<OUT_CH>(row_ptr(dst, r, a.stego_row_stride32), stego[2 * r], stego[2 * r + 1]);
        }
        if (SIDE) {
            if (a.sse != nullptr) {
                uint32_t sq = 0;
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const uint32_t d = hw::absdiff4(park[k * kBlkThreads], stego[k]);
                    sq = hw::dp4a(d, d, sq);                                 // <= 64 * 255^2 per block
                }
                if (!w.ok) sq = 0;
                sq = __reduce_add_sync(0xffffffffu, sq);
                if (w.f != sse_frame) {
                    if (lane == 0 && sse_acc != 0) atomicAdd(a.sse + sse_frame, sse_acc);
                    sse_acc = 0;
                    sse_frame = w.f;
                }
                sse_acc += sq;
            }
        }
        if (!more) {
            if (SIDE && a.sse != nullptr && lane == 0 && sse_acc != 0) atomicAdd(a.sse + sse_frame, sse_acc);
            break;
        }
        g = gn;
        w = wn;
    }
}

// Stores the (up to 64) packed words of one group: to this rank's buffer and to every peer
// (plain stores into peer-mapped memory), or once to the multicast address.
__device__ __forceinline__ void store_group_words(const BlkExtractArgs& a, long long row_off, int lane, int nwords,
                                                  const uint32_t (&wv)[2])
{
    if (a.multicast) {
        uint32_t* m32 = reinterpret_cast<uint32_t*>(a.peers[0] + row_off);
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (lane + 32 * j < nwords)
                asm volatile("multimem.st.weak.global.b32 [%0], %1;" ::"l"(m32 + lane + 32 * j), "r"(wv[j]) : "memory");
        return;
    }
    uint32_t* o32 = reinterpret_cast<uint32_t*>(a.bits + row_off);
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (lane + 32 * j < nwords) o32[lane + 32 * j] = wv[j];
    for (int e = 0; e < a.n_peers; ++e) {
        uint32_t* p32 = reinterpret_cast<uint32_t*>(a.peers[e] + row_off);
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (lane + 32 * j < nwords) p32[lane + 32 * j] = wv[j];
    }
}

template <int CH, int NP>
__global__ void __launch_bounds__(kBlkThreads, kBlkMinCtas) extract_blk_kernel(const BlkExtractArgs a)
{
    __shared__ uint32_t pack[kBlkWarps][64];
    const BlkGeom& G = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = NP == 4 ? (G.n >= 63 ? 63 : G.n) : G.n;
    const QuantRegs Q = make_quant_regs(a.q, G.delta32);
    const long long gstep = (long long)gridDim.x * kBlkWarps;
    long long g = (long long)blockIdx.x * kBlkWarps + warp;
    if (g >= G.total_groups) return;
    const uint32_t slot0 = (uint32_t)__cvta_generic_to_shared(blk_dyn_smem) + threadIdx.x * 8u;

    Where w = locate(G, g, lane);
    uint32_t rows[CH == 3 ? 48 : 16];
    stage_request<CH>(G, w, slot0);
    for (;;) {
        stage_fetch<CH>(slot0, rows);
        pack[warp][lane] = 0;
        pack[warp][lane + 32] = 0;
        P2 c[32];
        block_input<CH, false>(rows, G.magic_hi, c, nullptr);

        const long long gn = g + gstep;
        const bool more = gn < G.total_groups;
        Where wn = w;
        if (more) {
            wn = locate(G, gn, lane);
            stage_request<CH>(G, wn, slot0);
        }

        uint32_t hi, lo;
        block_extract_pairs<NP>(c, Q, n, hi, lo);
        if (!w.ok) { hi = 0; lo = 0; }
        __syncwarp();
        {   // place the n-bit string at bit offset lane*n of the warp's run of 32 n bits = n words
            uint32_t* p = pack[warp];
            const uint32_t o = (uint32_t)lane * (uint32_t)n, i0 = o >> 5, sh = o & 31;
            const uint32_t q0 = hi >> sh, q1 = __funnelshift_r(lo, hi, sh), q2 = __funnelshift_r(0u, lo, sh);
            if (q0) atomicOr(p + i0, q0);
            if (q1) atomicOr(p + i0 + 1, q1);
            if (q2) atomicOr(p + i0 + 2, q2);
        }
        __syncwarp();
        const int nblk = min(32, G.bpf - w.base);
        const int nwords = (nblk * n + 31) >> 5;
        const long long row_off = w.f * a.bits_frame_stride + (long long)(w.base >> 5) * (4 * n);
        uint32_t wv[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) wv[j] = hw::bswap(pack[warp][lane + 32 * j]);
        store_group_words(a, row_off, lane, nwords, wv);
        __syncwarp();
        if (!more) break;
        g = gn;
        w = wn;
    }
}
#endif  // __CUDACC__

}  // namespace blk
