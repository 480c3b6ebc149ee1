// svs_fast.cuh - the throughput kernels of the DCT-QIM path (included by svs_b200.cu).
//
// Same results as the scalar kernels in svs_b200.cu, bit for bit, but organised around what the
// B200 SM can issue (measured with profiles/microbench/pipes.cu: FP32 32 lanes/clk/SMSP, the
// ALU pipe - LOP3/PRMT/SHF/I2FP - only 16, F2I 4, F2IP.U8 8):
//   * two 8x8 blocks per thread, one in each half of a 64-bit register pair, so both 2-D
//     transforms run on packed FADD2/FFMA2 (half the issue slots of scalar FP32, no repacking);
//   * u8 -> f32 through one PRMT (byte into the mantissa of 2^23) and a packed subtract, and
//     BGR -> gray through two dp2a (doubled weights put the gray value in byte 2 of the sum);
//   * the float32 division of the quantiser (config_and_setup.py:148,160) replaced by one FMA
//     into a "magic" constant whose mantissa then holds floor() and the fraction; whenever the
//     fraction is within 2 ulp of a rounding boundary (this includes every exact tie) the row is
//     redone with the IEEE division, so the result is always the reference's;
//   * f32 -> u8 with saturation and truncation in one F2IP (cvt.rzi.u8.f32) on the otherwise
//     idle conversion pipe - that is np.uint8(np.clip(v, 0, 255)), config_and_setup.py:171.
// Only whole frames that the payload fills completely come here (k == n for every block); the
// frame in which the payload ends, strided/unaligned inputs and non-float32 deltas are handled by
// the scalar kernels; the optional gray / SSE outputs of the frames handled here come from the
// streaming side_outputs_kernel in svs_b200.cu.
#pragma once

namespace fast {

typedef unsigned long long u64;

struct P2 { u64 v; };                       // (block A, block B) as two binary32 values

__device__ __forceinline__ P2 pk(float a, float b)
{
    P2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ P2 pku(uint32_t a, uint32_t b)
{
    P2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ void unpk(P2 p, uint32_t& a, uint32_t& b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(p.v));
}
__device__ __forceinline__ void unpkf(P2 p, float& a, float& b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v));
}
__device__ __forceinline__ P2 add2(P2 a, P2 b)
{
    P2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ P2 sub2(P2 a, P2 b)
{
    P2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v));
    return r;
}
__device__ __forceinline__ P2 fma2(P2 a, P2 b, P2 c)
{
    P2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v));
    return r;
}

// Arithmetic policy for svs_math.cuh.  A product must round on its own before it is added to
// anything, but ptxas contracts mul.f32x2 + add.f32x2 regardless of .rn; fma(a, c, -0.0) with a
// -0.0 the compiler cannot see through (it arrives as a kernel argument) is an exact multiply
// that nothing can be fused into.
struct PackedOps {
    typedef P2 T;
    P2 negzero;
    __device__ __forceinline__ T add(T a, T b) const { return add2(a, b); }
    __device__ __forceinline__ T sub(T a, T b) const { return sub2(a, b); }
    __device__ __forceinline__ T mulc(T a, float c) const { return fma2(a, pk(c, c), negzero); }
    __device__ __forceinline__ void bfly(T a, T b, T& sum, T& diff) const { sum = add2(a, b); diff = sub2(a, b); }
};

using svs::FastQuant;            // host-computed constants of the division-free quantiser (svs_quant.h)

// One CTA per SM, 12 warps, 64 blocks (2 per thread) per warp and per loop iteration.  168
// registers per thread x 384 threads fills the register file; the CTA is persistent and strides
// over the (frame, 64-block group) space.
#ifndef SVS_FAST_THREADS
#define SVS_FAST_THREADS 384
#endif
constexpr int kFastThreads = SVS_FAST_THREADS;
constexpr int kFastCtasPerSm = 384 / kFastThreads;
constexpr int kFastWarps = kFastThreads / 32;
#ifndef SVS_SYNC_LEVEL
#define SVS_SYNC_LEVEL 1
#endif
#ifndef SVS_SYNC_EVERY
#define SVS_SYNC_EVERY 1
#endif
// Skewed lockstep (experiment, OFF by default).  In strict lockstep the three warps that share an
// SM sub-partition (warp % 4) run the same phase at the same time - all in an FP32-heavy
// transform, then all in the ALU-heavy quantiser.  With SVS_SKEW=1 the three "slots" (warp / 4)
// arrive at the SAME group barrier from three different places of the instruction stream, about a
// third of a row period (~50 instructions) apart, so that one warp's quantiser could run under
// the other two warps' transforms at no cost in instruction-cache footprint.  MEASURED: slower
// (embed 2.13 vs 1.85 ms, extract 0.85 vs 0.83 ms per 600 frames; 44 B more spills) - the phases
// of a single warp already interleave well enough and the mid-stream arrivals only add stalls.
#ifndef SVS_SKEW
#define SVS_SKEW 0
#endif
#define SVS_ARRIVE(site, on) do { if ((on) && (SVS_SKEW ? slot == (site) : (site) == 0)) asm volatile("bar.sync 0;" ::: "memory"); } while (0)
#define SVS_LOCKSTEP() do { if (SVS_SYNC_LEVEL >= 1 && (SVS_SYNC_EVERY == 1 || (iter++ % SVS_SYNC_EVERY) == 0)) __syncthreads(); } while (0)
#define SVS_LOCKSTEP2() do { if (SVS_SYNC_LEVEL >= 2) __syncthreads(); } while (0)
#define SVS_LOCKSTEP3() do { if (SVS_SYNC_LEVEL >= 3) __syncthreads(); } while (0)
#ifndef SVS_ZONE
#define SVS_ZONE 4
#endif
constexpr uint32_t kZone = SVS_ZONE;                    // flagged when fraction bits < kZone (shift = 2 ulp)

struct FastGeom {
    const uint8_t* frames;
    long long frame_stride, row_stride;
    int H, W, bw, bpf, n;
    int groups_per_frame;                         // ceil(bpf / 64)
    long long total_groups;                       // n_frames * groups_per_frame
    float delta32;
    uint32_t magic_hi;                            // 0x4B000000, opaque so that it stays in a register
    uint32_t bw_magic;                            // ceil(2^32 / bw): b / bw == umulhi(b, bw_magic) (svs_row.cuh)
};

struct FastEmbedArgs {
    FastGeom g;
    FastQuant q;
    const uint32_t* payload;
    long long payload_bit_offset, payload_last_word, cap;
    uint8_t* stego;
    long long stego_frame_stride, stego_row_stride;
    int64_t* bits_embedded;
};

constexpr int kMaxPeers = 15;

struct FastExtractArgs {
    FastGeom g;
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

// Stores the (up to 128) packed words of one 64-block group: to this rank's buffer and to every
// peer (plain stores into peer-mapped memory), or once to the multicast address.
__device__ __forceinline__ void store_group_words(const FastExtractArgs& a, long long row_off, int lane, int nwords,
                                                  const uint32_t (&wv)[4])
{
    if (a.multicast) {
        uint32_t* m32 = reinterpret_cast<uint32_t*>(a.peers[0] + row_off);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (lane + 32 * j < nwords)
                asm volatile("multimem.st.weak.global.b32 [%0], %1;" ::"l"(m32 + lane + 32 * j), "r"(wv[j]) : "memory");
        return;
    }
    uint32_t* o32 = reinterpret_cast<uint32_t*>(a.bits + row_off);
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (lane + 32 * j < nwords) o32[lane + 32 * j] = wv[j];
    for (int e = 0; e < a.n_peers; ++e) {
        uint32_t* p32 = reinterpret_cast<uint32_t*>(a.peers[e] + row_off);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (lane + 32 * j < nwords) p32[lane + 32 * j] = wv[j];
    }
}

__device__ __forceinline__ uint32_t bswap(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// (2^23 + byte SEL of v) as float bits: one PRMT  [v.bSEL, 0x00, 0x00, 0x4B]
template <int SEL>
__device__ __forceinline__ uint32_t magic_byte(uint32_t v, uint32_t magic_hi)
{
    return __byte_perm(v, magic_hi, 0x7540 | SEL);
}

// Row words of one block (in registers) -> its 8 gray bytes in two words.  BGR: cv2 BGR2GRAY,
// 2*(3735 B + 19235 G + 9798 R + 16384) < 2^24 through two dp2a per pixel with the 16-bit weights
// placed according to where the pixel's three bytes sit in the row's words (no byte shuffles);
// the gray value is bits 16..23 of the sum.
template <int CH>
__device__ __forceinline__ void row_to_gray(const uint2* w, uint32_t& lo, uint32_t& hi)
{
    if (CH == 1) {
        lo = w[0].x;
        hi = w[0].y;
    } else {
        constexpr uint32_t WB = 7470u, WG = 38470u, WR = 19596u, RND = 32768u;
        const uint32_t v[6] = {w[0].x, w[0].y, w[1].x, w[1].y, w[2].x, w[2].y};
        uint32_t s[8];
#pragma unroll
        for (int px = 0; px < 8; ++px) {
            const int byte0 = 3 * px, wi = byte0 >> 2, off = byte0 & 3;
            if (off == 0)      s[px] = __dp2a_hi(WR, v[wi], __dp2a_lo((WG << 16) | WB, v[wi], RND));
            else if (off == 1) s[px] = __dp2a_hi((WR << 16) | WG, v[wi], __dp2a_lo(WB << 16, v[wi], RND));
            else if (off == 2) s[px] = __dp2a_lo(WR, v[wi + 1], __dp2a_hi((WG << 16) | WB, v[wi], RND));
            else               s[px] = __dp2a_lo((WR << 16) | WG, v[wi + 1], __dp2a_hi(WB << 16, v[wi], RND));
        }
        lo = __byte_perm(__byte_perm(s[0], s[1], 0x0062), __byte_perm(s[2], s[3], 0x0062), 0x5410);
        hi = __byte_perm(__byte_perm(s[4], s[5], 0x0062), __byte_perm(s[6], s[7], 0x0062), 0x5410);
    }
}

__device__ __forceinline__ uint32_t to_u8(float v)      // np.uint8(np.clip(v, 0, 255)): one F2IP
{
    uint32_t r;
    asm("{.reg .u8 t; cvt.rzi.u8.f32 t, %1; cvt.u32.u8 %0, t;}" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ uint32_t pack4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3)
{
    return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
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

__device__ __forceinline__ void payload_window(const uint32_t* __restrict__ words, long long last_word,
                                               long long pos, uint32_t& hi, uint32_t& lo)
{
    const long long wi = pos >> 5;
    const uint32_t s = (uint32_t)(pos & 31);
    const uint32_t w0 = wi <= last_word ? bswap(__ldg(words + wi)) : 0u;
    const uint32_t w1 = wi + 1 <= last_word ? bswap(__ldg(words + wi + 1)) : 0u;
    const uint32_t w2 = wi + 2 <= last_word ? bswap(__ldg(words + wi + 2)) : 0u;
    hi = __funnelshift_l(w1, w0, s);
    lo = __funnelshift_l(w2, w1, s);
}

// IEEE-exact c / d from the correctly rounded reciprocal r = RN(1/d): two Newton corrections on
// the quotient (the sequence __fdiv_rn runs after its own reciprocal; all operands are normal
// here).  Checked against the FPU division on 2.5e8 (c, d) pairs covering this value range.
__device__ __forceinline__ float div_exact(float c, float d, float r)
{
    float q = __fmul_rn(c, r);
    q = __fmaf_rn(__fmaf_rn(-q, d, c), r, q);
    return __fmaf_rn(__fmaf_rn(-q, d, c), r, q);
}

// Coefficients (0,4), (4,0) and (4,4) of a block of integer pixels are exact multiples of 1/8
// (their basis is +-1/8), so c / delta lands EXACTLY on a rounding tie in one block out of
// 8 delta - far too often for the speculate-and-fix quantiser (a third of all warps would take
// the out-of-line path in every group, and the slowest warp holds up the lockstep barrier).
// These three are therefore always quantised with the IEEE division, packed for both blocks:
// the quotient exactly as div_exact, rint() through the 1.5 * 2^23 constant (round-half-even,
// the parity is the lowest mantissa bit, two's complement for negative quotients).
__device__ __forceinline__ bool tie_prone(int flat) { return flat == 4 || flat == 32 || flat == 36; }

struct ExactQ {
    P2 d, negd, r, magic, negzero;
    __device__ __forceinline__ P2 rint_quotient_plus_magic(P2 c) const     // 1.5 * 2^23 + rint(c / d)
    {
        P2 q = fma2(c, r, negzero);
        q = fma2(fma2(q, negd, c), r, q);
        q = fma2(fma2(q, negd, c), r, q);
        return add2(q, magic);
    }
    // q' = q - (q & 1) + bit, returned as float32(q' * d)    (config_and_setup.py:148-156)
    __device__ __forceinline__ P2 embed(P2 c, uint32_t bitA, uint32_t bitB) const
    {
        const P2 m = rint_quotient_plus_magic(c);
        uint32_t ma, mb;
        unpk(m, ma, mb);
        const P2 adj = pk(__int2float_rn((int)bitA - (int)(ma & 1u)), __int2float_rn((int)bitB - (int)(mb & 1u)));
        return fma2(add2(sub2(m, magic), adj), d, negzero);
    }
};
__device__ __forceinline__ ExactQ make_exact_q(float d32, float r32, float negzero)
{
    ExactQ e;
    e.d = pk(d32, d32);
    e.negd = pk(-d32, -d32);
    e.r = pk(r32, r32);
    e.magic = pk(12582912.0f, 12582912.0f);
    e.negzero = pk(negzero, negzero);
    return e;
}

// Rare path, deliberately out of line and looped so that it costs almost no instruction-cache
// space.  `orig` holds the 8 coefficient pairs of row u before quantisation, `res` the results
// of the division-free quantiser; every coefficient whose fraction was too close to a rounding
// boundary is recomputed exactly as the scalar kernels do (IEEE division, round-half-even,
// float32 product) and patched into `res`.
__device__ __noinline__ void fix_row_embed(const P2* orig, P2* res, int u, int n, float d32, float r32,
                                           float r2, float ke, uint32_t emask,
                                           uint32_t wA0, uint32_t wA1, uint32_t wB0, uint32_t wB1)
{
    for (int v = 0; v < 8; ++v) {
        const int i = 8 * u + v - 1;
        if (i < 0 || i >= n) continue;
        float ca, cb, ra, rb;
        unpkf(orig[v], ca, cb);
        unpkf(res[v], ra, rb);
        const uint32_t sh = 31u - (uint32_t)(i & 31);
        if ((__float_as_uint(__fmaf_rn(ca, r2, ke)) & emask) < kZone) {
            const int q = __float2int_rn(div_exact(ca, d32, r32));
            ra = __fmul_rn((float)(q - (q & 1) + (int)(((i < 32 ? wA0 : wA1) >> sh) & 1u)), d32);
        }
        if ((__float_as_uint(__fmaf_rn(cb, r2, ke)) & emask) < kZone) {
            const int q = __float2int_rn(div_exact(cb, d32, r32));
            rb = __fmul_rn((float)(q - (q & 1) + (int)(((i < 32 ? wB0 : wB1) >> sh) & 1u)), d32);
        }
        res[v] = pk(ra, rb);
    }
}

// Same for extraction: `rows` = (rowA | rowB << 16), coefficient v at bit 7-v; flagged
// coefficients get their parity from the exact quotient.
__device__ __noinline__ uint32_t fix_row_extract(const P2* in, uint32_t rows, int u, int n, float d32, float r32,
                                                 float kx, uint32_t xmask)
{
    for (int v = 0; v < 8; ++v) {
        const int i = 8 * u + v - 1;
        if (i < 0 || i >= n) continue;
        float ca, cb;
        unpkf(in[v], ca, cb);
        if ((__float_as_uint(__fmaf_rn(ca, r32, kx)) & xmask) < kZone) {
            const uint32_t par = (uint32_t)__float2int_rn(div_exact(ca, d32, r32)) & 1u;
            rows = (rows & ~(0x80u >> v)) | (par << (7 - v));
        }
        if ((__float_as_uint(__fmaf_rn(cb, r32, kx)) & xmask) < kZone) {
            const uint32_t par = (uint32_t)__float2int_rn(div_exact(cb, d32, r32)) & 1u;
            rows = (rows & ~(0x800000u >> v)) | (par << (23 - v));
        }
    }
    return rows;
}

// 64-bit pointer += 64-bit stride as two adds that ptxas does not re-derive from a base + offset.
template <class T>
__device__ __forceinline__ void step(T*& p, long long stride)
{
    asm volatile("add.s64 %0, %0, %1;" : "+l"(p) : "l"(stride));
}

// Where the 64 blocks of (group g) live, and this lane's two blocks (A = base + lane,
// B = A + 32, both clamped into the frame; ok* says whether the lane really owns them).
struct Lane {
    int f;                       // frame
    int base, bA, bB;
    int byA, bxA, byB, bxB;
    bool okA, okB;
};
__device__ __forceinline__ Lane locate(const FastGeom& G, long long g, int lane)
{
    Lane L;
    const unsigned gi = (unsigned)g;                       // total_groups < 2^31 (host-checked)
    const unsigned f = gi / (unsigned)G.groups_per_frame;
    L.f = (int)f;
    L.base = (int)(gi - f * (unsigned)G.groups_per_frame) * 64;
    const int last = G.bpf - 1;
    L.okA = L.base + lane <= last;
    L.okB = L.base + 32 + lane <= last;
    L.bA = min(L.base + lane, last);
    L.bB = min(L.base + 32 + lane, last);
    L.byA = (int)((unsigned)L.bA / (unsigned)G.bw);
    L.bxA = L.bA - L.byA * G.bw;
    if (G.bw >= 32 && L.okB) {                             // B is 32 blocks further along the raster
        L.bxB = L.bxA + 32;
        L.byB = L.byA;
        if (L.bxB >= G.bw) { L.bxB -= G.bw; L.byB += 1; }
    } else {
        L.byB = (int)((unsigned)L.bB / (unsigned)G.bw);
        L.bxB = L.bB - L.byB * G.bw;
    }
    return L;
}

// The same from (frame, first block of the group) with a multiply-high instead of the divisions:
// cheap enough to be recomputed where it is needed instead of being kept (and spilled) across
// a whole group.  Needs bpf * bw < 2^32 (host-checked).
__device__ __forceinline__ Lane relocate(const FastGeom& G, int f, int base, int lane, bool live)
{
    Lane L;
    L.f = f;
    L.base = base;
    const int last = G.bpf - 1;
    L.okA = live && base + lane <= last;
    L.okB = live && base + 32 + lane <= last;
    L.bA = min(base + lane, last);
    L.bB = min(base + 32 + lane, last);
    L.byA = (int)__umulhi((unsigned)L.bA, G.bw_magic);
    L.bxA = L.bA - L.byA * G.bw;
    L.byB = (int)__umulhi((unsigned)L.bB, G.bw_magic);
    L.bxB = L.bB - L.byB * G.bw;
    return L;
}
__device__ __forceinline__ void group_of(const FastGeom& G, long long g, int& f, int& base, bool& live)
{
    live = g < G.total_groups;
    const unsigned gi = (unsigned)(live ? g : G.total_groups - 1);
    const unsigned ff = gi / (unsigned)G.groups_per_frame;
    f = (int)ff;
    base = (int)(gi - ff * (unsigned)G.groups_per_frame) * 64;
}

// Raw row words of one block of the lane (8 rows x P 8-byte words) -> registers.
template <int CH>
__device__ __forceinline__ void load_block_raw(const FastGeom& G, int f, int by, int bx, uint2* raw)
{
    constexpr int P = CH == 3 ? 3 : 1;
    const uint8_t* p = G.frames + f * G.frame_stride + (long long)(by * 8) * G.row_stride + bx * (8 * CH);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int j = 0; j < P; ++j) raw[r * P + j] = __ldg(reinterpret_cast<const uint2*>(p) + j);
        step(p, G.row_stride);
    }
}

// Image rows [R0, R1) of one block of the lane -> their slots of `raw` (the rest is untouched).
template <int CH, int R0, int R1>
__device__ __forceinline__ void load_block_rows(const FastGeom& G, int f, int by, int bx, uint2* raw)
{
    constexpr int P = CH == 3 ? 3 : 1;
    const uint8_t* p = G.frames + f * G.frame_stride + (long long)(by * 8 + R0) * G.row_stride + bx * (8 * CH);
#pragma unroll
    for (int r = R0; r < R1; ++r) {
#pragma unroll
        for (int j = 0; j < P; ++j) raw[r * P + j] = __ldg(reinterpret_cast<const uint2*>(p) + j);
        step(p, G.row_stride);
    }
}

// Pulls the lane's rows of a (future) group into L2; no registers are held.
template <int CH>
__device__ __forceinline__ void prefetch_l2(const FastGeom& G, const Lane& L)
{
    const uint8_t* frame = G.frames + L.f * G.frame_stride;
    const uint8_t* pA = frame + (long long)(L.byA * 8) * G.row_stride + L.bxA * (8 * CH);
    const uint8_t* pB = frame + (long long)(L.byB * 8) * G.row_stride + L.bxB * (8 * CH);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pA));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pB));
        step(pA, G.row_stride);
        step(pB, G.row_stride);
    }
}

template <int CH>
__device__ __forceinline__ void raw_to_gray(const uint2* raw, uint32_t (&g)[16])
{
    constexpr int P = CH == 3 ? 3 : 1;
#pragma unroll
    for (int r = 0; r < 8; ++r) row_to_gray<CH>(raw + r * P, g[2 * r], g[2 * r + 1]);
}

// raw_to_gray of the first block with the barrier arrivals of slots 1 and 2 inside (BGR input:
// 22 instructions per row, so rows 2 and 4 are ~44 and ~88 instructions after the loop top).
template <int CH>
__device__ __forceinline__ void raw_to_gray_arrive(const uint2* raw, uint32_t (&g)[16], int slot, bool on)
{
    constexpr int P = CH == 3 ? 3 : 1;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        row_to_gray<CH>(raw + r * P, g[2 * r], g[2 * r + 1]);
        if (CH == 3 && r == 1) SVS_ARRIVE(1, on);
        if (CH == 3 && r == 3) SVS_ARRIVE(2, on);
    }
}

// The lane's blocks for group g; warps past the end of the work redo the last group with their
// stores masked so that every warp of the CTA executes the same instruction stream.
__device__ __forceinline__ Lane locate_or_idle(const FastGeom& G, long long g, int lane, bool& live)
{
    live = g < G.total_groups;
    Lane L = locate(G, live ? g : G.total_groups - 1, lane);
    if (!live) { L.okA = false; L.okB = false; }
    return L;
}

// Column c of both blocks -> 8 packed floats, then the axis-0 transform of that column; written
// column by column so that the byte->float work (ALU pipe) of column c+1 can overlap the FP32
// work of column c inside one warp.
template <int C>
__device__ __forceinline__ void column_fwd(const PackedOps& ops, const uint32_t (&gA)[16], const uint32_t (&gB)[16],
                                           uint32_t magic_hi, P2 (&x)[64])
{
    const P2 unbias = pk(-8388608.0f, -8388608.0f);
#pragma unroll
    for (int r = 0; r < 8; ++r)
        x[r * 8 + C] = add2(pku(magic_byte<(C & 3)>(gA[2 * r + (C >> 2)], magic_hi),
                                magic_byte<(C & 3)>(gB[2 * r + (C >> 2)], magic_hi)), unbias);      // exact
    svs::dct8_fwd<8>(ops, x + C);
}

// Input stage for BGR frames: image row r of both blocks straight to its eight
// packed floats, without packing the gray bytes into words first (BGR: the dp2a sums keep the
// gray value in byte 2 and the PRMT that builds the 2^23 + gray float reads it from there), then
// the eight column transforms.  96 PRMT fewer per 64-block group, no gA / gB arrays.
template <int CH>
__device__ __forceinline__ void row_to_x(const uint2* wa, const uint2* wb, uint32_t magic_hi, P2* xr)
{
    const P2 unbias = pk(-8388608.0f, -8388608.0f);
    if (CH == 1) {
        const uint32_t a[2] = {wa[0].x, wa[0].y}, b[2] = {wb[0].x, wb[0].y};
#pragma unroll
        for (int c = 0; c < 8; ++c)
            xr[c] = add2(pku(__byte_perm(a[c >> 2], magic_hi, 0x7540 | (c & 3)), __byte_perm(b[c >> 2], magic_hi, 0x7540 | (c & 3))), unbias);
    } else {
        constexpr uint32_t WB = 7470u, WG = 38470u, WR = 19596u, RND = 32768u;
        const uint32_t va[6] = {wa[0].x, wa[0].y, wa[1].x, wa[1].y, wa[2].x, wa[2].y};
        const uint32_t vb[6] = {wb[0].x, wb[0].y, wb[1].x, wb[1].y, wb[2].x, wb[2].y};
#pragma unroll
        for (int px = 0; px < 8; ++px) {
            const int byte0 = 3 * px, wi = byte0 >> 2, off = byte0 & 3;
            uint32_t sa, sb;
            if (off == 0) {
                sa = __dp2a_hi(WR, va[wi], __dp2a_lo((WG << 16) | WB, va[wi], RND));
                sb = __dp2a_hi(WR, vb[wi], __dp2a_lo((WG << 16) | WB, vb[wi], RND));
            } else if (off == 1) {
                sa = __dp2a_hi((WR << 16) | WG, va[wi], __dp2a_lo(WB << 16, va[wi], RND));
                sb = __dp2a_hi((WR << 16) | WG, vb[wi], __dp2a_lo(WB << 16, vb[wi], RND));
            } else if (off == 2) {
                sa = __dp2a_lo(WR, va[wi + 1], __dp2a_hi((WG << 16) | WB, va[wi], RND));
                sb = __dp2a_lo(WR, vb[wi + 1], __dp2a_hi((WG << 16) | WB, vb[wi], RND));
            } else {
                sa = __dp2a_lo((WR << 16) | WG, va[wi + 1], __dp2a_hi(WB << 16, va[wi], RND));
                sb = __dp2a_lo((WR << 16) | WG, vb[wi + 1], __dp2a_hi(WB << 16, vb[wi], RND));
            }
            xr[px] = add2(pku(__byte_perm(sa, magic_hi, 0x7542), __byte_perm(sb, magic_hi, 0x7542)), unbias);   // exact
        }
    }
}
template <int CH>
__device__ __forceinline__ void input_rowwise(const PackedOps& ops, const uint2* rawA, const uint2* rawB, uint32_t magic_hi, P2 (&x)[64])
{
    constexpr int P = CH == 3 ? 3 : 1;
#pragma unroll
    for (int r = 0; r < 8; ++r) row_to_x<CH>(rawA + r * P, rawB + r * P, magic_hi, x + 8 * r);
#pragma unroll
    for (int c = 0; c < 8; ++c) svs::dct8_fwd<8>(ops, x + c);
}

// ------------------------------------------------------------------------------------------
// embed: every block of every frame handled here is completely filled with payload (k == n)
// ------------------------------------------------------------------------------------------
template <int CH, int OUT_CH, bool NFULL>
__global__ void __launch_bounds__(kFastThreads, kFastCtasPerSm) embed_fast_kernel(const FastEmbedArgs a)
{
    const FastGeom& G = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = min(warp >> 2, 2);            // which of the warps of its SM sub-partition this is
    const int n = NFULL ? 63 : G.n;
    PackedOps ops;
    ops.negzero = pk(a.q.negzero, a.q.negzero);
    const P2 r2 = pk(a.q.r2, a.q.r2), ke = pk(a.q.ke, a.q.ke), d2 = pk(a.q.d2, a.q.d2), k0 = pk(a.q.k0, a.q.k0);
    const uint32_t emask = a.q.emask, ebit = a.q.ebit;
    const int erot = a.q.erot;
    const ExactQ exq = make_exact_q(G.delta32, a.q.r, a.q.negzero);

    // The raw input words of a group are loaded into registers while the PREVIOUS group is
    // being written out (its coefficient registers die row by row), so that the HBM latency is
    // not paid by twelve warps at once at the top of every group.
    constexpr int P = CH == 3 ? 3 : 1;
    uint2 rawA[8 * P], rawB[8 * P];
    const long long gstep = (long long)gridDim.x * kFastWarps;
    bool live;
#ifdef SVS_RECOMPUTE_LANE
    int cf, cbase;                       // all that is carried across a group: frame and first block
    group_of(G, (long long)blockIdx.x * kFastWarps + warp, cf, cbase, live);
    {
        const Lane L = relocate(G, cf, cbase, lane, live);
        load_block_raw<CH>(G, L.f, L.byA, L.bxA, rawA);
        load_block_raw<CH>(G, L.f, L.byB, L.bxB, rawB);
    }
#else
    Lane L = locate_or_idle(G, (long long)blockIdx.x * kFastWarps + warp, lane, live);
    load_block_raw<CH>(G, L.f, L.byA, L.bxA, rawA);
    load_block_raw<CH>(G, L.f, L.byB, L.bxB, rawB);
#endif
#ifdef SVS_EARLY_GRAY
    // block A's gray bytes are made at the END of the previous group (its raw words arrived long
    // before): 16 instead of 48 live registers across the top of the loop, where pressure peaks
    uint32_t gA[16];
    raw_to_gray<CH>(rawA, gA);
#endif

    unsigned iter = 0;
    (void)iter;
#ifdef SVS_SPLIT_BARRIER
    // split-phase group barrier on an mbarrier: every warp ARRIVES at the top of the group and only
    // WAITS after the gray conversion (~350 instructions later), so that the normal skew between the
    // warps is absorbed instead of being paid as a stall, while they still share one cache window
    __shared__ unsigned long long group_bar;
    const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&group_bar);
    if (threadIdx.x == 0) asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(bar_addr), "r"(kFastWarps));
    __syncthreads();
    uint32_t bar_phase = 0;
#endif
    for (long long g0 = (long long)blockIdx.x * kFastWarps; g0 < G.total_groups; g0 += gstep) {
#ifdef SVS_SPLIT_BARRIER
        if (lane == 0) asm volatile("{.reg .b64 t; mbarrier.arrive.shared.b64 t, [%0];}" ::"r"(bar_addr) : "memory");
#else
        SVS_ARRIVE(0, SVS_SYNC_LEVEL >= 1);    // keep the warps of the CTA in one instruction-cache window
#endif
#ifdef SVS_RECOMPUTE_LANE
        Lane L = relocate(G, cf, cbase, lane, live);
#endif
        if (live && L.base + lane == 0 && a.bits_embedded != nullptr) a.bits_embedded[L.f] = a.cap;
#ifdef SVS_L2_PREFETCH
        {
            bool nl;
            const Lane N = locate_or_idle(G, g0 + gstep + warp, lane, nl);
            prefetch_l2<CH>(G, N);
        }
#endif

        P2 x[64];
#ifndef SVS_COLUMNWISE_INPUT
        if (CH == 3) {                             // BGR: row-wise, no packed gray words (1 % faster, measured)
            SVS_ARRIVE(1, SVS_SYNC_LEVEL >= 1);
            SVS_ARRIVE(2, SVS_SYNC_LEVEL >= 1);
            input_rowwise<CH>(ops, rawA, rawB, G.magic_hi, x);
        } else
#endif
        {
#ifdef SVS_EARLY_GRAY
            uint32_t gB[16];
#else
            uint32_t gA[16], gB[16];
            raw_to_gray_arrive<CH>(rawA, gA, slot, SVS_SYNC_LEVEL >= 1);
#endif
            raw_to_gray<CH>(rawB, gB);
#ifdef SVS_SPLIT_BARRIER
            asm volatile("{.reg .pred p;\n"
                         "SVS_WAIT_%=: mbarrier.try_wait.parity.shared.b64 p, [%0], %1;\n"
                         "@!p bra SVS_WAIT_%=;}" ::"r"(bar_addr), "r"(bar_phase) : "memory");
            bar_phase ^= 1u;
#endif
            column_fwd<0>(ops, gA, gB, G.magic_hi, x);
            if (CH != 3) SVS_ARRIVE(1, SVS_SYNC_LEVEL >= 1);
            column_fwd<1>(ops, gA, gB, G.magic_hi, x);
            if (CH != 3) SVS_ARRIVE(2, SVS_SYNC_LEVEL >= 1);
            column_fwd<2>(ops, gA, gB, G.magic_hi, x); column_fwd<3>(ops, gA, gB, G.magic_hi, x);
            column_fwd<4>(ops, gA, gB, G.magic_hi, x); column_fwd<5>(ops, gA, gB, G.magic_hi, x);
            column_fwd<6>(ops, gA, gB, G.magic_hi, x); column_fwd<7>(ops, gA, gB, G.magic_hi, x);
        }
        SVS_LOCKSTEP2();

        // payload windows of the two blocks: coefficient i reads bit i (MSB first).  The words are
        // pre-rotated once so that bit i sits `i` places below position erot: bringing it there
        // is then a rotate by the compile-time constant i.
        uint32_t wA[2], wB[2], pA[2], pB[2];
        {
            const long long at = a.payload_bit_offset + L.f * a.cap;
            payload_window(a.payload, a.payload_last_word, at + (long long)L.bA * n, wA[0], wA[1]);
            payload_window(a.payload, a.payload_last_word, at + (long long)L.bB * n, wB[0], wB[1]);
            const int pre = (erot - 31) & 31;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                pA[j] = __funnelshift_l(wA[j], wA[j], pre);
                pB[j] = __funnelshift_l(wB[j], wB[j], pre);
            }
        }

        // axis-1 transform of row u, then its quantisation (FP32 and ALU work interleave)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            svs::dct8_fwd<1>(ops, x + 8 * u);
            if (NFULL || 8 * u - 1 < n) {
                uint32_t worst = 0xffffffffu;
                P2 nx[8];
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const int i = 8 * u + v - 1;                  // payload bit / coefficient number
                    nx[v] = x[8 * u + v];
                    if (i >= 0 && (NFULL || i < n) && tie_prone(8 * u + v)) {
                        nx[v] = exq.embed(x[8 * u + v], (wA[i >> 5] >> (31 - (i & 31))) & 1u, (wB[i >> 5] >> (31 - (i & 31))) & 1u);
                    } else if (i >= 0 && (NFULL || i < n)) {
                        const P2 y = fma2(x[8 * u + v], r2, ke);
                        uint32_t ya, yb;
                        unpk(y, ya, yb);
                        worst = min(worst, min(ya & emask, yb & emask));
                        const uint32_t ta = __funnelshift_l(pA[i >> 5], pA[i >> 5], i & 31) & ebit;
                        const uint32_t tb = __funnelshift_l(pB[i >> 5], pB[i >> 5], i & 31) & ebit;
                        // M + floor() + bit/2, then (2e + bit) * delta in one rounding
                        nx[v] = fma2(pku((ya & ~emask) | ta, (yb & ~emask) | tb), d2, k0);
                    }
                }
                if (worst < kZone) {                              // rare: a fraction too close to call
                    P2 orig[8], res[8];
#pragma unroll
                    for (int v = 0; v < 8; ++v) { orig[v] = x[8 * u + v]; res[v] = nx[v]; }
                    fix_row_embed(orig, res, u, n, G.delta32, a.q.r, a.q.r2, a.q.ke, emask, wA[0], wA[1], wB[0], wB[1]);
#pragma unroll
                    for (int v = 0; v < 8; ++v) nx[v] = res[v];
                }
#pragma unroll
                for (int v = 0; v < 8; ++v) x[8 * u + v] = nx[v];
            }
            if (u == 3) SVS_LOCKSTEP3();
        }
        SVS_LOCKSTEP2();

#pragma unroll
        for (int v = 0; v < 8; ++v) svs::dct8_inv<8>(ops, x + v);

        SVS_LOCKSTEP2();
#ifdef SVS_RECOMPUTE_LANE
        asm volatile("" : "+r"(cbase));                                // recompute, do not keep
        L = relocate(G, cf, cbase, lane, live);
        uint8_t* out = a.stego + L.f * a.stego_frame_stride;
        uint8_t* dstA = out + (long long)(L.byA * 8) * a.stego_row_stride + L.bxA * (8 * OUT_CH);
        uint8_t* dstB = out + (long long)(L.byB * 8) * a.stego_row_stride + L.bxB * (8 * OUT_CH);
        const bool okA = L.okA, okB = L.okB;
        group_of(G, g0 + gstep + warp, cf, cbase, live);               // the next group of this warp
        L = relocate(G, cf, cbase, lane, live);
#else
        uint8_t* out = a.stego + L.f * a.stego_frame_stride;
        uint8_t* dstA = out + (long long)(L.byA * 8) * a.stego_row_stride + L.bxA * (8 * OUT_CH);
        uint8_t* dstB = out + (long long)(L.byB * 8) * a.stego_row_stride + L.bxB * (8 * OUT_CH);
        const bool okA = L.okA, okB = L.okB;
        L = locate_or_idle(G, g0 + gstep + warp, lane, live);          // the next group of this warp
#endif
#ifndef SVS_NO_PAYLOAD_PREFETCH
        {   // its payload words are consumed right after they are loaded (top of the next group): pull
            // the two cache lines of the lane's blocks into L1 now, a whole output phase ahead
            const long long at = a.payload_bit_offset + L.f * a.cap;
            const long long wa = min((at + (long long)L.bA * n) >> 5, a.payload_last_word);
            const long long wb = min((at + (long long)L.bB * n) >> 5, a.payload_last_word);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.payload + wa));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a.payload + wb));
        }
#endif
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            svs::dct8_inv<1>(ops, x + 8 * r);
            uint32_t ba[8], bb[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float va, vb;
                unpkf(x[r * 8 + c], va, vb);
                ba[c] = to_u8(va);
                bb[c] = to_u8(vb);
            }
            const uint32_t a0 = pack4(ba[0], ba[1], ba[2], ba[3]), a1 = pack4(ba[4], ba[5], ba[6], ba[7]);
            const uint32_t b0 = pack4(bb[0], bb[1], bb[2], bb[3]), b1 = pack4(bb[4], bb[5], bb[6], bb[7]);
            if (okA) store_row<OUT_CH>(dstA, a0, a1);
            if (okB) store_row<OUT_CH>(dstB, b0, b1);
            step(dstA, a.stego_row_stride);
            step(dstB, a.stego_row_stride);
            if (r == 3) SVS_LOCKSTEP3();
            // rows 0..r of x are dead now: their registers take the next group's input, a few image
            // rows at a time and as early as the freed registers allow (16 per row of x, 6 per BGR row)
#ifndef SVS_FINE_PREFETCH      /* measured: no gain (1.89 vs 1.87 ms per 600 frames) */
            if (r == 3) load_block_raw<CH>(G, L.f, L.byA, L.bxA, rawA);
            if (r == 6) load_block_raw<CH>(G, L.f, L.byB, L.bxB, rawB);
#else
            if (r == 0) load_block_rows<CH, 0, 2>(G, L.f, L.byA, L.bxA, rawA);
            if (r == 1) load_block_rows<CH, 2, 5>(G, L.f, L.byA, L.bxA, rawA);
            if (r == 2) load_block_rows<CH, 5, 8>(G, L.f, L.byA, L.bxA, rawA);
            if (r == 3) load_block_rows<CH, 0, 2>(G, L.f, L.byB, L.bxB, rawB);
            if (r == 4) load_block_rows<CH, 2, 5>(G, L.f, L.byB, L.bxB, rawB);
            if (r == 5) load_block_rows<CH, 5, 8>(G, L.f, L.byB, L.bxB, rawB);
#endif
        }
#ifdef SVS_EARLY_GRAY
        raw_to_gray<CH>(rawA, gA);
#endif
    }
}

// ------------------------------------------------------------------------------------------
// extract
// ------------------------------------------------------------------------------------------
template <int CH, bool NFULL>
__global__ void __launch_bounds__(kFastThreads, kFastCtasPerSm) extract_fast_kernel(const FastExtractArgs a)
{
    __shared__ uint32_t pack[kFastWarps][128];
    const FastGeom& G = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = min(warp >> 2, 2);
    const int n = NFULL ? 63 : G.n;
    PackedOps ops;
    ops.negzero = pk(a.q.negzero, a.q.negzero);
    const P2 rr = pk(a.q.r, a.q.r), kx = pk(a.q.kx, a.q.kx);
    const uint32_t xmask = a.q.xmask;
    const int xk = a.q.xk;
    const ExactQ exq = make_exact_q(G.delta32, a.q.r, a.q.negzero);

    constexpr int P = CH == 3 ? 3 : 1;
    uint2 rawA[8 * P], rawB[8 * P];
    const long long gstep = (long long)gridDim.x * kFastWarps;
    bool live;
    Lane L = locate_or_idle(G, (long long)blockIdx.x * kFastWarps + warp, lane, live);
    load_block_raw<CH>(G, L.f, L.byA, L.bxA, rawA);
    load_block_raw<CH>(G, L.f, L.byB, L.bxB, rawB);

    unsigned iter = 0;
    (void)iter;
    for (long long g0 = (long long)blockIdx.x * kFastWarps; g0 < G.total_groups; g0 += gstep) {
        const bool sync_now = SVS_SYNC_LEVEL >= 1 && (iter++ & 3u) == 0;   // small code: a loose lockstep is enough
        SVS_ARRIVE(0, sync_now);
#pragma unroll
        for (int j = 0; j < 4; ++j) pack[warp][lane + 32 * j] = 0;

        P2 x[64];
#ifndef SVS_COLUMNWISE_INPUT
        if (CH == 3) {
            SVS_ARRIVE(1, sync_now);
            SVS_ARRIVE(2, sync_now);
            input_rowwise<CH>(ops, rawA, rawB, G.magic_hi, x);
        } else
#endif
        {
            uint32_t gA[16], gB[16];
            raw_to_gray_arrive<CH>(rawA, gA, slot, sync_now);
            raw_to_gray<CH>(rawB, gB);
            column_fwd<0>(ops, gA, gB, G.magic_hi, x);
            if (CH != 3) SVS_ARRIVE(1, sync_now);
            column_fwd<1>(ops, gA, gB, G.magic_hi, x);
            if (CH != 3) SVS_ARRIVE(2, sync_now);
            column_fwd<2>(ops, gA, gB, G.magic_hi, x); column_fwd<3>(ops, gA, gB, G.magic_hi, x);
            column_fwd<4>(ops, gA, gB, G.magic_hi, x); column_fwd<5>(ops, gA, gB, G.magic_hi, x);
            column_fwd<6>(ops, gA, gB, G.magic_hi, x); column_fwd<7>(ops, gA, gB, G.magic_hi, x);
        }
        SVS_LOCKSTEP2();
        const Lane C = L;                                      // current group (for the output)
        const bool clive = live;
        L = locate_or_idle(G, g0 + gstep + warp, lane, live);  // next group: loaded during the quantiser

        uint32_t hiA = 0, loA = 0, hiB = 0, loB = 0;            // bit i at (hi:lo) bit 63-i
        svs::dct8_fwd<1>(ops, x);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            // (row u+1 is transformed BEFORE row u is read out, in the same basic block: its FP32
            // work fills the pipe while the parity read-out keeps the ALU pipe busy)
            if (u < 7 && (NFULL || 8 * (u + 1) - 1 < n)) svs::dct8_fwd<1>(ops, x + 8 * (u + 1));
            if (NFULL || 8 * u - 1 < n) {
                uint32_t worst = 0xffffffffu;
                uint32_t rowA = 0, rowB = 0;                     // coefficient v of this row at bit 7-v
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const int i = 8 * u + v - 1;
                    if (i >= 0 && (NFULL || i < n) && tie_prone(8 * u + v)) {
                        uint32_t ma, mb;                         // parity = lowest mantissa bit
                        unpk(exq.rint_quotient_plus_magic(x[8 * u + v]), ma, mb);
                        rowA |= (ma & 1u) << (7 - v);
                        rowB |= (mb & 1u) << (7 - v);
                    } else if (i >= 0 && (NFULL || i < n)) {
                        const P2 y = fma2(x[8 * u + v], rr, kx);
                        uint32_t ya, yb;
                        unpk(y, ya, yb);
                        worst = min(worst, min(ya & xmask, yb & xmask));
                        const int rot = (7 - v - xk) & 31;       // parity (bit xk) -> bit 7-v
                        rowA |= __funnelshift_l(ya, ya, rot) & (0x80u >> v);
                        rowB |= __funnelshift_l(yb, yb, rot) & (0x80u >> v);
                    }
                }
                if (worst < kZone) {
                    P2 in[8];
#pragma unroll
                    for (int v = 0; v < 8; ++v) in[v] = x[8 * u + v];
                    const uint32_t both = fix_row_extract(in, rowA | (rowB << 16), u, n, G.delta32, a.q.r, a.q.kx, xmask);
                    rowA = both & 0xffu;
                    rowB = both >> 16;
                }
                // bit 7-v of the row -> stream bit 8u+v-1 of the block -> (hi:lo) bit 64-8u-v
                if (u == 0)      { hiA |= rowA << 25; hiB |= rowB << 25; }
                else if (u < 4)  { hiA |= rowA << (25 - 8 * u); hiB |= rowB << (25 - 8 * u); }
                else if (u == 4) { hiA |= rowA >> 7; loA |= rowA << 25; hiB |= rowB >> 7; loB |= rowB << 25; }
                else             { loA |= rowA << (57 - 8 * u); loB |= rowB << (57 - 8 * u); }
            }
            if (u == 3) load_block_raw<CH>(G, L.f, L.byA, L.bxA, rawA);
            if (u == 6) load_block_raw<CH>(G, L.f, L.byB, L.bxB, rawB);
        }
        if (!C.okA) { hiA = 0; loA = 0; }
        if (!C.okB) { hiB = 0; loB = 0; }
        __syncwarp();
        {
            // place the two n-bit strings at bit offsets lane*n and (32+lane)*n of the warp's run
            uint32_t* p = pack[warp];
            uint32_t o = (uint32_t)lane * (uint32_t)n, w0 = o >> 5, sh = o & 31;
            uint32_t p0 = hiA >> sh, p1 = __funnelshift_r(loA, hiA, sh), p2 = __funnelshift_r(0u, loA, sh);
            if (p0) atomicOr(p + w0, p0);
            if (p1) atomicOr(p + w0 + 1, p1);
            if (p2) atomicOr(p + w0 + 2, p2);
            o = (uint32_t)(32 + lane) * (uint32_t)n; w0 = o >> 5; sh = o & 31;
            p0 = hiB >> sh; p1 = __funnelshift_r(loB, hiB, sh); p2 = __funnelshift_r(0u, loB, sh);
            if (p0) atomicOr(p + w0, p0);
            if (p1) atomicOr(p + w0 + 1, p1);
            if (p2) atomicOr(p + w0 + 2, p2);
        }
        __syncwarp();
        const int nblk = clive ? min(64, G.bpf - C.base) : 0;
        const int nwords = (nblk * n + 31) >> 5;
        const long long row_off = C.f * a.bits_frame_stride + (long long)(C.base >> 5) * (4 * n);
        uint32_t wv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) wv[j] = bswap(pack[warp][lane + 32 * j]);
        store_group_words(a, row_off, lane, nwords, wv);
        __syncwarp();
    }
}

}  // namespace fast
