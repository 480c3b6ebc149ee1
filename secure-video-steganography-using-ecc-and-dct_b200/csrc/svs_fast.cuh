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
// frame in which the payload ends, strided/unaligned inputs, non-float32 deltas and the optional
// gray / SSE outputs are handled by the scalar kernels.
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
};

// Host-computed constants of the division-free quantiser (see make_fast_quant in svs_b200.cu).
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

constexpr int kFastThreads = 128;                 // 4 warps, 64 blocks per warp
constexpr int kFastBlocksPerCta = 2 * kFastThreads;
constexpr uint32_t kZone = 4;                     // flagged when fraction bits < kZone (shift = 2 ulp)

struct FastGeom {
    const uint8_t* frames;
    long long frame_stride, row_stride;
    int H, W, bw, bpf, tiles_per_frame, n;
    float delta32;
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

struct FastExtractArgs {
    FastGeom g;
    FastQuant q;
    uint8_t* bits;
    long long bits_frame_stride;
};

__device__ __forceinline__ uint32_t bswap(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// (2^23 + byte SEL of v) as float bits: one PRMT
template <int SEL>
__device__ __forceinline__ uint32_t magic_byte(uint32_t v)
{
    return __byte_perm(v, 0x4B000000u, 0x7540 | SEL);      // [v.bSEL, 0x00, 0x00, 0x4B]
}

// Row r of one block -> 8 "2^23 + gray" float bit patterns.
template <int CH>
__device__ __forceinline__ void load_row_magic(const uint8_t* __restrict__ row, uint32_t (&m)[8])
{
    if (CH == 1) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(row));
        m[0] = magic_byte<0>(v.x); m[1] = magic_byte<1>(v.x); m[2] = magic_byte<2>(v.x); m[3] = magic_byte<3>(v.x);
        m[4] = magic_byte<0>(v.y); m[5] = magic_byte<1>(v.y); m[6] = magic_byte<2>(v.y); m[7] = magic_byte<3>(v.y);
    } else {
        const uint2 a = __ldg(reinterpret_cast<const uint2*>(row));
        const uint2 b = __ldg(reinterpret_cast<const uint2*>(row) + 1);
        const uint2 c = __ldg(reinterpret_cast<const uint2*>(row) + 2);
        const uint32_t w[7] = {a.x, a.y, b.x, b.y, c.x, c.y, 0u};
#pragma unroll
        for (int px = 0; px < 8; ++px) {
            const int byte0 = 3 * px, wi = byte0 >> 2, off = byte0 & 3;
            const uint32_t sel = (uint32_t)(off | ((off + 1) << 4) | ((off + 2) << 8) | ((off + 3) << 12));
            const uint32_t bgr = off == 0 ? w[wi] : __byte_perm(w[wi], w[wi + 1], sel);
            // 2*(3735 B + 19235 G + 9798 R + 16384) < 2^24: gray = bits 16..23 (cv2 BGR2GRAY)
            uint32_t s = __dp2a_lo((38470u << 16) | 7470u, bgr, 32768u);
            s = __dp2a_hi(19596u, bgr, s);
            m[px] = magic_byte<2>(s);
        }
    }
}

// 8 floats of row r of both blocks -> two uint2 of bytes (clip to [0,255], truncate)
__device__ __forceinline__ uint32_t to_u8(float v)
{
    uint32_t r;
    asm("{.reg .u8 t; cvt.rzi.u8.f32 t, %1; cvt.u32.u8 %0, t;}" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ uint32_t pack4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3)
{
    return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
}

template <int OUT_CH>
__device__ __forceinline__ void store_row(uint8_t* dst, uint32_t lo4, uint32_t hi4)
{
    uint2* row = reinterpret_cast<uint2*>(dst);
    if (OUT_CH == 1) {
        row[0] = make_uint2(lo4, hi4);
    } else {            // gray replicated to B,G,R (cv2.cvtColor GRAY2BGR, embed_process.py:126)
        row[0] = make_uint2(__byte_perm(lo4, 0, 0x1000), __byte_perm(lo4, 0, 0x2211));
        row[1] = make_uint2(__byte_perm(lo4, 0, 0x3332), __byte_perm(hi4, 0, 0x1000));
        row[2] = make_uint2(__byte_perm(hi4, 0, 0x2211), __byte_perm(hi4, 0, 0x3332));
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

// Exact (reference-order) requantisation of one coefficient: the scalar kernels' formula.
__device__ __forceinline__ float requant_exact(float c, float d32, uint32_t bit)
{
    const float t = __fdiv_rn(c, d32);
    const int q = __float2int_rn(t);
    return __fmul_rn((float)(q - (q & 1) + (int)bit), d32);
}

// ------------------------------------------------------------------------------------------
// embed: every block of every frame handled here is completely filled with payload (k == n)
// ------------------------------------------------------------------------------------------
template <int CH, int OUT_CH>
__global__ void __launch_bounds__(kFastThreads, 3) embed_fast_kernel(const FastEmbedArgs a)
{
    const FastGeom& G = a.g;
    const long long f = blockIdx.x / G.tiles_per_frame;
    const int tile = (int)(blockIdx.x - f * G.tiles_per_frame);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = tile * kFastBlocksPerCta + warp * 64;
    if (base >= G.bpf) return;
    if (base + lane == 0 && a.bits_embedded != nullptr) a.bits_embedded[f] = a.cap;

    const int last = G.bpf - 1;
    const int bA = min(base + lane, last), bB = min(base + 32 + lane, last);
    const bool okA = base + lane <= last, okB = base + 32 + lane <= last;
    const int byA = bA / G.bw, bxA = bA - byA * G.bw;
    const int byB = bB / G.bw, bxB = bB - byB * G.bw;
    const uint8_t* frame = G.frames + f * G.frame_stride;
    const uint8_t* srcA = frame + (long long)(byA * 8) * G.row_stride + (long long)bxA * (8 * CH);
    const uint8_t* srcB = frame + (long long)(byB * 8) * G.row_stride + (long long)bxB * (8 * CH);

    PackedOps ops;
    ops.negzero = pk(a.q.negzero, a.q.negzero);
    P2 x[64];
    {
        const P2 unbias = pk(-8388608.0f, -8388608.0f);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint32_t mA[8], mB[8];
            load_row_magic<CH>(srcA + r * G.row_stride, mA);
            load_row_magic<CH>(srcB + r * G.row_stride, mB);
#pragma unroll
            for (int c = 0; c < 8; ++c) x[r * 8 + c] = add2(pku(mA[c], mB[c]), unbias);   // exact
        }
    }
    svs::dct2_fwd(ops, x);

    // payload windows of the two blocks: coefficient i reads bit i (MSB first)
    uint32_t wA[2], wB[2];
    {
        const long long at = a.payload_bit_offset + f * a.cap;
        payload_window(a.payload, a.payload_last_word, at + (long long)bA * G.n, wA[0], wA[1]);
        payload_window(a.payload, a.payload_last_word, at + (long long)bB * G.n, wB[0], wB[1]);
    }
    const int n = G.n;
    const P2 r2 = pk(a.q.r2, a.q.r2), ke = pk(a.q.ke, a.q.ke), d2 = pk(a.q.d2, a.q.d2), k0 = pk(a.q.k0, a.q.k0);
    const uint32_t emask = a.q.emask, ebit = a.q.ebit;
    const int erot = a.q.erot;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (8 * u - 1 < n) {                                  // uniform: row u holds indices 8u-1 .. 8u+6
            uint32_t worst = 0xffffffffu;
            P2 keep[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const int i = 8 * u + v - 1;                  // payload bit / coefficient number
                keep[v] = x[8 * u + v];
                if (i >= 0 && i < n) {
                    const P2 y = fma2(x[8 * u + v], r2, ke);
                    uint32_t ya, yb;
                    unpk(y, ya, yb);
                    const uint32_t la = ya & emask, lb = yb & emask;
                    worst = min(worst, min(la, lb));
                    // rotate payload bit i (bit 31-(i&31) of its word) to position erot
                    const int rot = (erot - (31 - (i & 31))) & 31;
                    const uint32_t ta = __funnelshift_l(wA[i >> 5], wA[i >> 5], rot) & ebit;
                    const uint32_t tb = __funnelshift_l(wB[i >> 5], wB[i >> 5], rot) & ebit;
                    // M + floor() + bit/2, then (2e + bit) * delta in one rounding
                    x[8 * u + v] = fma2(pku((ya & ~emask) | ta, (yb & ~emask) | tb), d2, k0);
                }
            }
            if (worst < kZone) {                              // rare: a fraction too close to call
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const int i = 8 * u + v - 1;
                    if (i >= 0 && i < n) {
                        float ca, cb;
                        unpkf(keep[v], ca, cb);
                        const uint32_t bitA = (wA[i >> 5] >> (31 - (i & 31))) & 1u;
                        const uint32_t bitB = (wB[i >> 5] >> (31 - (i & 31))) & 1u;
                        x[8 * u + v] = pk(requant_exact(ca, G.delta32, bitA), requant_exact(cb, G.delta32, bitB));
                    }
                }
            }
        }
    }
    svs::dct2_inv(ops, x);

    uint8_t* out = a.stego + f * a.stego_frame_stride;
    uint8_t* dstA = out + (long long)(byA * 8) * a.stego_row_stride + (long long)bxA * (8 * OUT_CH);
    uint8_t* dstB = out + (long long)(byB * 8) * a.stego_row_stride + (long long)bxB * (8 * OUT_CH);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        uint32_t ba[8], bb[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float va, vb;
            unpkf(x[r * 8 + c], va, vb);
            ba[c] = to_u8(va);
            bb[c] = to_u8(vb);
        }
        if (okA) store_row<OUT_CH>(dstA + r * a.stego_row_stride, pack4(ba[0], ba[1], ba[2], ba[3]), pack4(ba[4], ba[5], ba[6], ba[7]));
        if (okB) store_row<OUT_CH>(dstB + r * a.stego_row_stride, pack4(bb[0], bb[1], bb[2], bb[3]), pack4(bb[4], bb[5], bb[6], bb[7]));
    }
}

// ------------------------------------------------------------------------------------------
// extract
// ------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(kFastThreads, 3) extract_fast_kernel(const FastExtractArgs a)
{
    __shared__ uint32_t pack[kFastThreads / 32][128];
    const FastGeom& G = a.g;
    const long long f = blockIdx.x / G.tiles_per_frame;
    const int tile = (int)(blockIdx.x - f * G.tiles_per_frame);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int base = tile * kFastBlocksPerCta + warp * 64;
    if (base >= G.bpf) return;
    const int n = G.n;

#pragma unroll
    for (int j = 0; j < 4; ++j) pack[warp][lane + 32 * j] = 0;
    __syncwarp();

    const int last = G.bpf - 1;
    const int bA = min(base + lane, last), bB = min(base + 32 + lane, last);
    const bool okA = base + lane <= last, okB = base + 32 + lane <= last;
    const int byA = bA / G.bw, bxA = bA - byA * G.bw;
    const int byB = bB / G.bw, bxB = bB - byB * G.bw;
    const uint8_t* frame = G.frames + f * G.frame_stride;
    const uint8_t* srcA = frame + (long long)(byA * 8) * G.row_stride + (long long)bxA * (8 * CH);
    const uint8_t* srcB = frame + (long long)(byB * 8) * G.row_stride + (long long)bxB * (8 * CH);

    PackedOps ops;
    ops.negzero = pk(a.q.negzero, a.q.negzero);
    P2 x[64];
    {
        const P2 unbias = pk(-8388608.0f, -8388608.0f);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint32_t mA[8], mB[8];
            load_row_magic<CH>(srcA + r * G.row_stride, mA);
            load_row_magic<CH>(srcB + r * G.row_stride, mB);
#pragma unroll
            for (int c = 0; c < 8; ++c) x[r * 8 + c] = add2(pku(mA[c], mB[c]), unbias);
        }
    }
    svs::dct2_fwd(ops, x);

    const P2 rr = pk(a.q.r, a.q.r), kx = pk(a.q.kx, a.q.kx);
    const uint32_t xmask = a.q.xmask;
    const int xk = a.q.xk;
    uint32_t hiA = 0, loA = 0, hiB = 0, loB = 0;            // bit i at (hi:lo) bit 63-i
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        if (8 * u - 1 < n) {
            uint32_t worst = 0xffffffffu;
            uint32_t rowA = 0, rowB = 0;                     // coefficient v of this row at bit 7-v
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const int i = 8 * u + v - 1;
                if (i >= 0 && i < n) {
                    const P2 y = fma2(x[8 * u + v], rr, kx);
                    uint32_t ya, yb;
                    unpk(y, ya, yb);
                    worst = min(worst, min(ya & xmask, yb & xmask));
                    // parity (bit xk) -> bit 7-v
                    const int rot = (7 - v - xk) & 31;
                    rowA |= __funnelshift_l(ya, ya, rot) & (0x80u >> v);
                    rowB |= __funnelshift_l(yb, yb, rot) & (0x80u >> v);
                }
            }
            if (worst < kZone) {
                rowA = 0;
                rowB = 0;
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const int i = 8 * u + v - 1;
                    if (i >= 0 && i < n) {
                        float ca, cb;
                        unpkf(x[8 * u + v], ca, cb);
                        rowA |= ((uint32_t)__float2int_rn(__fdiv_rn(ca, G.delta32)) & 1u) << (7 - v);
                        rowB |= ((uint32_t)__float2int_rn(__fdiv_rn(cb, G.delta32)) & 1u) << (7 - v);
                    }
                }
            }
            // row u covers stream bits 8u-1 .. 8u+6 of the block: bit 7-v of row -> stream bit 8u+v-1
            if (u == 0)      { hiA |= rowA << 25; hiB |= rowB << 25; }            // v=1..7 -> bits 0..6
            else if (u < 4)  { hiA |= rowA << (25 - 8 * u); hiB |= rowB << (25 - 8 * u); }
            else if (u == 4) { hiA |= rowA >> 7; loA |= rowA << 25; hiB |= rowB >> 7; loB |= rowB << 25; }
            else             { loA |= rowA << (57 - 8 * u); loB |= rowB << (57 - 8 * u); }
        }
    }
    if (!okA) { hiA = 0; loA = 0; }
    if (!okB) { hiB = 0; loB = 0; }
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
    const int nblk = min(64, G.bpf - base);
    const int nwords = (nblk * n + 31) >> 5;
    uint32_t* o32 = reinterpret_cast<uint32_t*>(a.bits + f * a.bits_frame_stride + (long long)(base >> 5) * (4 * n));
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (lane + 32 * j < nwords) o32[lane + 32 * j] = bswap(pack[warp][lane + 32 * j]);
}

}  // namespace fast
