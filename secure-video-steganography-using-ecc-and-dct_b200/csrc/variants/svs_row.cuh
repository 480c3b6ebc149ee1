// svs_row.cuh - "row" kernels of the DCT-QIM path (included by svs_b200.cu); kernel family 4.
//
// Same arithmetic and same results as svs_fast.cuh, bit for bit (packed FADD2/FFMA2 transforms in
// scipy's operation order, division-free quantiser with the exact out-of-line fallback), but the
// 8x8 block is spread over EIGHT lanes instead of living in one thread:
//   * a lane owns one row (or column) of a PAIR of horizontally adjacent blocks - block A in the
//     low half, block B in the high half of every 64-bit register pair - so a 1-D pass is ONE
//     8-point transform per lane (54 packed instructions) and the working set is 16 registers;
//   * between the passes the 8x8 tile is transposed through a warp-private, padded (conflict
//     free) shared-memory tile: three 64-bit transposes per embed, one per extract, plus one byte
//     transpose of the gray input; only __syncwarp is ever needed;
//   * the loop body is ~650 instructions (10 KB): it fits the 32 KB L1.5 instruction cache, so
//     the warps of an SM run FREE (svs_fast.cuh is 4 k instructions and has to keep its 12 warps
//     in lockstep with a bar.sync per group, which serialises its load / FP32 / ALU phases);
//   * 64-80 registers per thread -> 24-32 resident warps per SM; the next tile's input and
//     payload words are still requested one tile ahead;
//   * a lane reads 16 (gray) or 48 (BGR) contiguous bytes with LDG.128 and writes its stego row of
//     both blocks with one STG.128.
// One warp = 4 block pairs = 8 consecutive blocks (64 x 8 pixels) per tile, 8 tiles per 64-block
// group (so that the extracted bits of a group start on an 8-byte boundary, as in svs_fast.cuh).
// Needs an even number of blocks per row (W % 16 == 0, W >= 64) and 16-byte aligned rows;
// everything else goes to the other kernel families.
//
// MEASURED (B200, 600 x 1080p, DESIGN.md section 4 / profiles/r1_row_variant_summary.txt): embed
// 2.43 ms, extract 1.26 ms against 1.84 / 0.83 ms for the lockstep kernels.  The transposes, the
// per-16-pixel bookkeeping and the run-time row index cost 35 % more issue slots per pixel and the
// free-running warps reach 58 % issue utilisation (short-scoreboard / MIO stalls on the
// transposes), so this family is NOT the default; it is kept because it is parity-green (every
// parity test runs it) and is the reference point for the alternative organisation.
#pragma once

namespace row {

using namespace fast;

#ifndef SVS_ROW_THREADS
#define SVS_ROW_THREADS 256
#endif
#ifndef SVS_ROW_CTAS
#define SVS_ROW_CTAS 3
#endif
constexpr int kRowThreads = SVS_ROW_THREADS;
constexpr int kRowWarps = kRowThreads / 32;
constexpr int kRowCtasPerSm = SVS_ROW_CTAS;
constexpr int kLineBytes = 72;                    // one line of a pair's 8x8 tile of 64-bit values (padded)
constexpr int kPairBytes = 576;                   // 8 lines; 72 eight-byte units: pairs alternate bank halves
constexpr int kWarpUnits = 4 * kPairBytes / 8;    // 2304 bytes per warp

template <int CH>
struct RawRow { uint4 v[CH == 3 ? 3 : 1]; };              // 16 pixels of one image row

template <int CH>
__device__ __forceinline__ RawRow<CH> load_row(const uint8_t* p)
{
    RawRow<CH> R;
#pragma unroll
    for (int k = 0; k < (CH == 3 ? 3 : 1); ++k) R.v[k] = __ldg(reinterpret_cast<const uint4*>(p) + k);
    return R;
}

// 8 BGR pixels in 6 words -> 8 gray bytes in 2 words (cv2 BGR2GRAY, config_and_setup.py:112):
// 2*(3735 B + 19235 G + 9798 R + 16384) < 2^24 through two dp2a per pixel, gray = bits 16..23.
__device__ __forceinline__ void bgr8_to_gray(const uint32_t* v, uint32_t& lo, uint32_t& hi)
{
    constexpr uint32_t WB = 7470u, WG = 38470u, WR = 19596u, RND = 32768u;
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

// The lane's image row of both blocks -> gray bytes (aLo,aHi | bLo,bHi).
template <int CH>
__device__ __forceinline__ void row_gray(const RawRow<CH>& R, uint32_t& aLo, uint32_t& aHi, uint32_t& bLo, uint32_t& bHi)
{
    if (CH == 1) {
        aLo = R.v[0].x; aHi = R.v[0].y; bLo = R.v[0].z; bHi = R.v[0].w;
    } else {
        const uint32_t w[12] = {R.v[0].x, R.v[0].y, R.v[0].z, R.v[0].w, R.v[CH == 3 ? 1 : 0].x, R.v[CH == 3 ? 1 : 0].y,
                                R.v[CH == 3 ? 1 : 0].z, R.v[CH == 3 ? 1 : 0].w, R.v[CH == 3 ? 2 : 0].x, R.v[CH == 3 ? 2 : 0].y,
                                R.v[CH == 3 ? 2 : 0].z, R.v[CH == 3 ? 2 : 0].w};
        bgr8_to_gray(w, aLo, aHi);
        bgr8_to_gray(w + 6, bLo, bHi);
    }
}

// Shared memory is addressed through 32-bit shared-window addresses (one register per warp tile).
__device__ __forceinline__ void sts64(uint32_t addr, P2 v)
{
    asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v.v) : "memory");
}
__device__ __forceinline__ P2 lds64(uint32_t addr)
{
    P2 r;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r.v) : "r"(addr) : "memory");
    return r;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr)
{
    uint32_t r;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(r) : "r"(addr) : "memory");
    return r;
}

// Gray rows (lane = row r of its pair) -> axis-0 input of column c (lane = column c): a byte
// transpose through the warp's tile.  Row r of a pair is stored as A0 B0 A1 B1 ... A7 B7
// (16 bytes, pair stride 144 bytes: conflict free), column c is eight 16-bit loads.
// `bytes` = tile + pair * 144 (shared address).
__device__ __forceinline__ void gray_rows_to_columns(uint32_t bytes, int sub,
                                                     uint32_t aLo, uint32_t aHi, uint32_t bLo, uint32_t bHi,
                                                     uint32_t magic_hi, P2 (&x)[8])
{
    sts128(bytes + sub * 16, __byte_perm(aLo, bLo, 0x5140), __byte_perm(aLo, bLo, 0x7362),
           __byte_perm(aHi, bHi, 0x5140), __byte_perm(aHi, bHi, 0x7362));
    __syncwarp();
    const P2 unbias = pk(-8388608.0f, -8388608.0f);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint32_t ab = lds16(bytes + r * 16 + sub * 2);
        x[r] = add2(pku(magic_byte<0>(ab, magic_hi), magic_byte<1>(ab, magic_hi)), unbias);      // exact
    }
    __syncwarp();
}

// 8x8 transpose of the packed values of a pair: lane `sub` holds line `sub` (x[k] = element k of
// it) and receives element `sub` of every line.  Line l of a pair starts at byte 72 l of the
// pair's tile (pair stride 576 bytes = 8 mod 16 eight-byte units): eight conflict-free STS.64,
// then eight conflict-free LDS.64.  (128-bit stores need aligned register quads and cost more
// MOVs than they save.)  `wr` = pair tile + sub * 72, `rd` = pair tile + sub * 8.
__device__ __forceinline__ void transpose8(uint32_t wr, uint32_t rd, P2 (&x)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) sts64(wr + k * 8, x[k]);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = lds64(rd + k * kLineBytes);
    __syncwarp();
}

// which of the lane's 8 coefficients (row u of the block: flat indices 8u..8u+7) carry payload:
// bit 7-v set when 0 <= 8u+v-1 < n   (config_and_setup.py:138-140)
__device__ __forceinline__ uint32_t coeff_mask(int u, int n)
{
    uint32_t m = 0;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
        const int i = 8 * u + v - 1;
        if (i >= 0 && i < n) m |= 0x80u >> v;
    }
    return m;
}

// Rare path of the embed quantiser (see fix_row_embed in svs_fast.cuh): bytes hold the payload
// bits of the row, coefficient v at bit 7-v.
__device__ __noinline__ void fix_row_embed_bytes(const P2* orig, P2* res, uint32_t valid, float d32, float r32,
                                                 float r2, float ke, uint32_t emask, uint32_t bitsA, uint32_t bitsB)
{
    for (int v = 0; v < 8; ++v) {
        if (!((valid >> (7 - v)) & 1u)) continue;
        float ca, cb, ra, rb;
        unpkf(orig[v], ca, cb);
        unpkf(res[v], ra, rb);
        if ((__float_as_uint(__fmaf_rn(ca, r2, ke)) & emask) < kZone) {
            const int q = __float2int_rn(div_exact(ca, d32, r32));
            ra = __fmul_rn((float)(q - (q & 1) + (int)((bitsA >> (7 - v)) & 1u)), d32);
        }
        if ((__float_as_uint(__fmaf_rn(cb, r2, ke)) & emask) < kZone) {
            const int q = __float2int_rn(div_exact(cb, d32, r32));
            rb = __fmul_rn((float)(q - (q & 1) + (int)((bitsB >> (7 - v)) & 1u)), d32);
        }
        res[v] = pk(ra, rb);
    }
}

// Rare path of the extract quantiser: rows = rowA | rowB << 16, coefficient v at bit 7-v.
__device__ __noinline__ uint32_t fix_row_extract_bytes(const P2* in, uint32_t rows, uint32_t valid, float d32, float r32,
                                                       float kx, uint32_t xmask)
{
    for (int v = 0; v < 8; ++v) {
        if (!((valid >> (7 - v)) & 1u)) continue;
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

// The two 32-bit words that hold the 8 payload bits starting at stream bit `pos`; the second
// index is clamped (its bits are only needed when they exist).
__device__ __forceinline__ void payload_words(const uint32_t* __restrict__ words, int last_word, long long pos,
                                              uint32_t& w0, uint32_t& w1)
{
    const int wi = (int)(pos >> 5);
    w0 = __ldg(words + wi);
    w1 = __ldg(words + min(wi + 1, last_word));
}
// ... and those 8 bits (MSB first) in bits 7..0.
__device__ __forceinline__ uint32_t payload_byte(uint32_t w0, uint32_t w1, uint32_t shift)
{
    return __funnelshift_l(bswap(w1), bswap(w0), shift) >> 24;
}

// Per-lane position inside the batch.  A warp takes 64-block groups g, g + gstep, ...; inside a
// group it walks 8 tiles of 8 blocks along the raster, so everything is stepped incrementally
// (pointer += constant, wrap at the end of a block row) and only recomputed once per group.
struct Walk {
    unsigned g, gstep;
    int j;                       // tile inside the group
    int f, base;                 // frame and first block of the group
    int b, bx;                   // the lane's block A (even) and its column
    bool ok;                     // the pair exists (not past the end of the frame)
    const uint8_t* src;          // the lane's image row of the pair
};

template <int CH>
__device__ __forceinline__ void walk_group(Walk& w, const FastGeom& G, int pair, int sub)
{
    const unsigned ff = w.g / (unsigned)G.groups_per_frame;
    w.f = (int)ff;
    w.base = (int)(w.g - ff * (unsigned)G.groups_per_frame) * 64;
    w.j = 0;
    const int b = w.base + 2 * pair;
    w.ok = b < G.bpf;
    w.b = min(b, G.bpf - 2);
    const int by = (int)__umulhi((unsigned)w.b, G.bw_magic);       // b / bw (host: bpf * bw < 2^32)
    w.bx = w.b - by * G.bw;
    w.src = G.frames + w.f * G.frame_stride + (long long)(by * 8 + sub) * G.row_stride + w.bx * (8 * CH);
}

// next tile of the group; returns true when the lane's pair moved to the next block row
template <int CH>
__device__ __forceinline__ bool walk_tile(Walk& w, const FastGeom& G, long long wrap_src)
{
    ++w.j;
    bool wrapped = false;
    if (w.b + 8 < G.bpf) {
        w.b += 8;
        w.bx += 8;
        w.src += 64 * CH;
        if (w.bx >= G.bw) { w.bx -= G.bw; w.src += wrap_src; wrapped = true; }
    } else {
        w.ok = false;                                     // stay on the last pair (valid addresses)
    }
    return wrapped;
}

struct RowEmbedArgs {
    FastEmbedArgs e;
    long long wrap_src, wrap_dst;                         // 8 rows down, one block row back
};
struct RowExtractArgs {
    FastExtractArgs x;
    long long wrap_src;
};

// ------------------------------------------------------------------------------------------
// embed: every block of every frame handled here is completely filled with payload (k == n)
// ------------------------------------------------------------------------------------------
template <int CH, int OUT_CH, bool NFULL>
__global__ void __launch_bounds__(kRowThreads, kRowCtasPerSm) embed_row_kernel(const RowEmbedArgs ra)
{
    __shared__ __align__(16) unsigned long long tiles[kRowWarps][kWarpUnits];
    const FastEmbedArgs& a = ra.e;
    const FastGeom& G = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = lane >> 3, sub = lane & 7;
    const int n = NFULL ? 63 : G.n;
    const uint32_t tile = (uint32_t)__cvta_generic_to_shared(tiles[warp]);
    const uint32_t t_bytes = tile + pair * 144;                       // byte transpose of the gray input
    const uint32_t t_wr = tile + pair * kPairBytes + sub * kLineBytes;  // 64-bit transposes
    const uint32_t t_rd = tile + pair * kPairBytes + sub * 8;
    PackedOps ops;
    ops.negzero = pk(a.q.negzero, a.q.negzero);
    const P2 r2 = pk(a.q.r2, a.q.r2), ke = pk(a.q.ke, a.q.ke), d2 = pk(a.q.d2, a.q.d2), k0 = pk(a.q.k0, a.q.k0);
    const uint32_t emask = a.q.emask, ebit = a.q.ebit;
    const int prerot = (a.q.erot - 7) & 31;               // payload byte bit 7 -> bit erot
    const uint32_t valid = coeff_mask(sub, n);            // this lane quantises row u = sub
    const int last_word = (int)a.payload_last_word;
    const int lead = sub ? 8 * sub - 1 : 0;               // first payload bit of row u inside its block

    Walk w;
    w.g = blockIdx.x * kRowWarps + warp;
    w.gstep = gridDim.x * kRowWarps;
    if ((long long)w.g >= G.total_groups) return;
    walk_group<CH>(w, G, pair, sub);
    // output row and payload position travel with the walk
    uint8_t* dst;
    long long pbit;
    auto group_outputs = [&]() {
        const int by = (int)__umulhi((unsigned)w.b, G.bw_magic);
        dst = a.stego + w.f * a.stego_frame_stride + (long long)(by * 8 + sub) * a.stego_row_stride + w.bx * (8 * OUT_CH);
        pbit = a.payload_bit_offset + w.f * a.cap + (long long)w.b * n + lead;
        if (w.ok && w.b == 0 && sub == 0 && a.bits_embedded != nullptr) a.bits_embedded[w.f] = a.cap;
    };
    group_outputs();
    RawRow<CH> raw = load_row<CH>(w.src);
    uint32_t pw[4];
    payload_words(a.payload, last_word, pbit, pw[0], pw[1]);
    payload_words(a.payload, last_word, pbit + n, pw[2], pw[3]);
    bool more = true;

#pragma unroll 1
    while (more) {
        // this tile: gray bytes, payload bits, where the result goes
        uint32_t aLo, aHi, bLo, bHi;
        row_gray<CH>(raw, aLo, aHi, bLo, bHi);
        uint32_t bitsA = payload_byte(pw[0], pw[1], (uint32_t)pbit & 31u);
        uint32_t bitsB = payload_byte(pw[2], pw[3], (uint32_t)(pbit + n) & 31u);
        if (sub == 0) { bitsA >>= 1; bitsB >>= 1; }       // row 0 has no v = 0 (DC)
        uint8_t* const out = dst;
        const bool ok = w.ok;
        // next tile: step (or start the next group) and request its input one tile ahead
        if (w.j < 7) {
            const bool moved = w.ok && w.b + 8 < G.bpf;
            const bool wrapped = walk_tile<CH>(w, G, ra.wrap_src);
            if (moved) {
                dst += 64 * OUT_CH;
                pbit += 8 * n;
                if (wrapped) dst += ra.wrap_dst;
            }
        } else {
            w.g += w.gstep;
            more = (long long)w.g < G.total_groups;
            if (more) {
                walk_group<CH>(w, G, pair, sub);
                group_outputs();
            }
        }
        if (more) {
            raw = load_row<CH>(w.src);
            payload_words(a.payload, last_word, pbit, pw[0], pw[1]);
            payload_words(a.payload, last_word, pbit + n, pw[2], pw[3]);
        }

        P2 x[8];
        gray_rows_to_columns(t_bytes, sub, aLo, aHi, bLo, bHi, G.magic_hi, x);
        svs::dct8_fwd<1>(ops, x);                          // axis 0: lane = column
        transpose8(t_wr, t_rd, x);
        svs::dct8_fwd<1>(ops, x);                          // axis 1: lane = row u, x[v] = coefficient (u, v)

        // quantise / re-parity (config_and_setup.py:146-158), division free (see svs_fast.cuh)
        {
            const uint32_t pA = __funnelshift_l(bitsA, bitsA, prerot), pB = __funnelshift_l(bitsB, bitsB, prerot);
            uint32_t worst = 0xffffffffu;
            P2 nx[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const P2 y = fma2(x[v], r2, ke);
                uint32_t ya, yb;
                unpk(y, ya, yb);
                worst = min(worst, min(ya & emask, yb & emask));
                const uint32_t ta = __funnelshift_l(pA, pA, v) & ebit;
                const uint32_t tb = __funnelshift_l(pB, pB, v) & ebit;
                nx[v] = fma2(pku((ya & ~emask) | ta, (yb & ~emask) | tb), d2, k0);
            }
            if (worst < kZone) {                           // rare: a fraction too close to call
                P2 orig[8], res[8];
#pragma unroll
                for (int v = 0; v < 8; ++v) { orig[v] = x[v]; res[v] = nx[v]; }
                fix_row_embed_bytes(orig, res, valid, G.delta32, a.q.r, a.q.r2, a.q.ke, emask, bitsA, bitsB);
#pragma unroll
                for (int v = 0; v < 8; ++v) nx[v] = res[v];
            }
#pragma unroll
            for (int v = 0; v < 8; ++v)
                if ((NFULL && v > 0) || ((valid >> (7 - v)) & 1u)) x[v] = nx[v];
        }

        transpose8(t_wr, t_rd, x);
        svs::dct8_inv<1>(ops, x);                          // axis 0: lane = column
        transpose8(t_wr, t_rd, x);
        svs::dct8_inv<1>(ops, x);                          // axis 1: lane = image row, x[c] = pixel c

        uint32_t ba[8], bb[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float va, vb;
            unpkf(x[c], va, vb);
            ba[c] = to_u8(va);                             // np.uint8(np.clip(v, 0, 255)), :171
            bb[c] = to_u8(vb);
        }
        const uint32_t a0 = pack4(ba[0], ba[1], ba[2], ba[3]), a1 = pack4(ba[4], ba[5], ba[6], ba[7]);
        const uint32_t b0 = pack4(bb[0], bb[1], bb[2], bb[3]), b1 = pack4(bb[4], bb[5], bb[6], bb[7]);
        if (ok) {
            if (OUT_CH == 1) {
                *reinterpret_cast<uint4*>(out) = make_uint4(a0, a1, b0, b1);
            } else {       // gray replicated to B,G,R (cv2.cvtColor GRAY2BGR, embed_process.py:126)
                uint4* d4 = reinterpret_cast<uint4*>(out);
                d4[0] = make_uint4(__byte_perm(a0, 0, 0x1000), __byte_perm(a0, 0, 0x2211), __byte_perm(a0, 0, 0x3332), __byte_perm(a1, 0, 0x1000));
                d4[1] = make_uint4(__byte_perm(a1, 0, 0x2211), __byte_perm(a1, 0, 0x3332), __byte_perm(b0, 0, 0x1000), __byte_perm(b0, 0, 0x2211));
                d4[2] = make_uint4(__byte_perm(b0, 0, 0x3332), __byte_perm(b1, 0, 0x1000), __byte_perm(b1, 0, 0x2211), __byte_perm(b1, 0, 0x3332));
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// extract
// ------------------------------------------------------------------------------------------
template <int CH, bool NFULL>
__global__ void __launch_bounds__(kRowThreads, kRowCtasPerSm) extract_row_kernel(const RowExtractArgs ra)
{
    __shared__ __align__(16) unsigned long long tiles[kRowWarps][kWarpUnits];
    __shared__ uint32_t pack[kRowWarps][128];
    const FastExtractArgs& a = ra.x;
    const FastGeom& G = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = lane >> 3, sub = lane & 7;
    const int n = NFULL ? 63 : G.n;
    const uint32_t tile = (uint32_t)__cvta_generic_to_shared(tiles[warp]);
    const uint32_t t_bytes = tile + pair * 144;
    const uint32_t t_wr = tile + pair * kPairBytes + sub * kLineBytes;
    const uint32_t t_rd = tile + pair * kPairBytes + sub * 8;
    uint32_t* pk32 = pack[warp];
    PackedOps ops;
    ops.negzero = pk(a.q.negzero, a.q.negzero);
    const P2 rr = pk(a.q.r, a.q.r), kx = pk(a.q.kx, a.q.kx);
    const uint32_t xmask = a.q.xmask;
    const int xk = a.q.xk;
    const uint32_t valid = coeff_mask(sub, n);
    // stream bit (inside the group) of the lane's first coefficient of block A in tile 0
    const uint32_t first0 = (uint32_t)(2 * pair) * (uint32_t)n + (sub ? 8u * sub - 1u : 0u);

    Walk w;
    w.g = blockIdx.x * kRowWarps + warp;
    w.gstep = gridDim.x * kRowWarps;
    if ((long long)w.g >= G.total_groups) return;
    walk_group<CH>(w, G, pair, sub);
    RawRow<CH> raw = load_row<CH>(w.src);
    bool more = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) pk32[lane + 32 * k] = 0;
    __syncwarp();

#pragma unroll 1
    while (more) {
        uint32_t aLo, aHi, bLo, bHi;
        row_gray<CH>(raw, aLo, aHi, bLo, bHi);
        const bool ok = w.ok;
        const int cj = w.j, cf = w.f, cbase = w.base;
        if (w.j < 7) {
            walk_tile<CH>(w, G, ra.wrap_src);
        } else {
            w.g += w.gstep;
            more = (long long)w.g < G.total_groups;
            if (more) walk_group<CH>(w, G, pair, sub);
        }
        if (more) raw = load_row<CH>(w.src);

        P2 x[8];
        gray_rows_to_columns(t_bytes, sub, aLo, aHi, bLo, bHi, G.magic_hi, x);
        svs::dct8_fwd<1>(ops, x);                          // axis 0
        transpose8(t_wr, t_rd, x);
        svs::dct8_fwd<1>(ops, x);                          // axis 1: x[v] = coefficient (u = sub, v)

        // parity of round(c / delta) (config_and_setup.py:159-163), coefficient v at bit 7-v
        uint32_t rowA = 0, rowB = 0, worst = 0xffffffffu;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const P2 y = fma2(x[v], rr, kx);
            uint32_t ya, yb;
            unpk(y, ya, yb);
            worst = min(worst, min(ya & xmask, yb & xmask));
            const int rot = (7 - v - xk) & 31;
            rowA |= __funnelshift_l(ya, ya, rot) & (0x80u >> v);
            rowB |= __funnelshift_l(yb, yb, rot) & (0x80u >> v);
        }
        if (worst < kZone) {
            P2 in[8];
#pragma unroll
            for (int v = 0; v < 8; ++v) in[v] = x[v];
            const uint32_t both = fix_row_extract_bytes(in, rowA | (rowB << 16), valid, G.delta32, a.q.r, a.q.kx, xmask);
            rowA = both & 0xffu;
            rowB = both >> 16;
        }
        rowA &= valid;
        rowB &= valid;
        if (!ok) { rowA = 0; rowB = 0; }
        // the row's bits go to stream bits blk*n + 8u-1 .. of the group (u = 0: DC has no bit)
        {
            const uint32_t first = first0 + (uint32_t)(8 * cj) * (uint32_t)n;
            if (sub == 0) { rowA <<= 1; rowB <<= 1; }
            uint32_t o = first, w0 = o >> 5, sh = o & 31;
            uint32_t hi = (rowA << 24) >> sh, lo = sh > 24 ? rowA << (56 - sh) : 0u;
            if (hi) atomicOr(pk32 + w0, hi);
            if (lo) atomicOr(pk32 + w0 + 1, lo);
            o = first + (uint32_t)n; w0 = o >> 5; sh = o & 31;
            hi = (rowB << 24) >> sh; lo = sh > 24 ? rowB << (56 - sh) : 0u;
            if (hi) atomicOr(pk32 + w0, hi);
            if (lo) atomicOr(pk32 + w0 + 1, lo);
        }

        if (cj == 7) {                                     // the group is complete: write it out
            __syncwarp();
            const int nblk = min(64, G.bpf - cbase);
            const int nwords = (nblk * n + 31) >> 5;
            const long long row_off = cf * a.bits_frame_stride + (long long)(cbase >> 5) * (4 * n);
            uint32_t wv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                wv[k] = bswap(pk32[lane + 32 * k]);
                pk32[lane + 32 * k] = 0;
            }
            store_group_words(a, row_off, lane, nwords, wv);
            __syncwarp();
        }
    }
}

}  // namespace row
