// svs_tile.cuh - "small code" throughput kernels of the DCT-QIM path (included by svs_b200.cu).
//
// Same arithmetic and same results as svs_fast.cuh (two 8x8 blocks per thread in packed
// FADD2/FFMA2 registers, division-free quantiser with an exact out-of-line fallback), but
// organised so that the hot loop FITS THE INSTRUCTION CACHE:
//   * each thread keeps its 64 coefficient pairs in a private, conflict-free slice of shared
//     memory (chunk c = elements 2c,2c+1 of thread t at 16-byte index c*T + t, so every warp access
//     is one contiguous 512-byte LDS.128/STS.128) instead of 128 registers;
//   * the four 1-D passes become real loops over rows / column pairs with a 16-register working
//     set: about 1.5 k instructions of code instead of 4.7 k, and no dependence on keeping the
//     warps of an SM in lockstep (svs_fast.cuh needs a bar.sync per group for that, which also
//     serialises its load, FP32 and ALU phases);
//   * ~70 registers per thread -> 14 free-running warps per SM (2 CTAs x 224 threads, 112 KB of
//     shared memory each) whose load / FP32 / ALU / conversion phases overlap naturally, so the
//     input is read with plain LDG.64 at the top of a group and needs no staging or prefetch.
// Shared-memory traffic is 3 KB per thread and group (768 clk of the SM's 128 B/clk pipe per
// warp and group, against ~1000 clk of FP32 per warp and group on each of the 4 sub-partitions).
#pragma once

namespace tile {

using namespace fast;

constexpr int kTileThreads = 224;                 // 7 warps; 2 CTAs per SM
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kTileCtasPerSm = 2;
constexpr int kTileSmemBytes = kTileThreads * 512;

// chunk c (coefficients 2c and 2c+1, both blocks) of this thread
__device__ __forceinline__ void ld_chunk(const ulonglong2* mine, int c, P2& e0, P2& e1)
{
    const ulonglong2 v = mine[c * kTileThreads];
    e0.v = v.x;
    e1.v = v.y;
}
__device__ __forceinline__ void st_chunk(ulonglong2* mine, int c, P2 e0, P2 e1)
{
    mine[c * kTileThreads] = make_ulonglong2(e0.v, e1.v);
}

// Row words of one block (in registers) -> its 8 gray bytes in two words.
template <int CH>
__device__ __forceinline__ void row_to_gray(const uint2* w, uint32_t& lo, uint32_t& hi)
{
    if (CH == 1) {
        lo = w[0].x;
        hi = w[0].y;
    } else {
        const uint32_t v[7] = {w[0].x, w[0].y, w[1].x, w[1].y, w[2].x, w[2].y, 0u};
        uint32_t s[8];
#pragma unroll
        for (int px = 0; px < 8; ++px) {
            const int byte0 = 3 * px, wi = byte0 >> 2, off = byte0 & 3;
            const uint32_t sel = (uint32_t)(off | ((off + 1) << 4) | ((off + 2) << 8) | ((off + 3) << 12));
            const uint32_t bgr = off == 0 ? v[wi] : __byte_perm(v[wi], v[wi + 1], sel);
            // 2*(3735 B + 19235 G + 9798 R + 16384) < 2^24: gray = bits 16..23 of the sum (cv2 BGR2GRAY)
            s[px] = __dp2a_hi(19596u, bgr, __dp2a_lo((38470u << 16) | 7470u, bgr, 32768u));
        }
        lo = __byte_perm(__byte_perm(s[0], s[1], 0x0062), __byte_perm(s[2], s[3], 0x0062), 0x5410);
        hi = __byte_perm(__byte_perm(s[4], s[5], 0x0062), __byte_perm(s[6], s[7], 0x0062), 0x5410);
    }
}

// Loads both blocks of the lane (all loads first, then the conversions) as packed gray bytes.
template <int CH>
__device__ __forceinline__ void load_gray(const FastGeom& G, const Lane& L, uint32_t (&gA)[16], uint32_t (&gB)[16])
{
    constexpr int P = CH == 3 ? 3 : 1;
    const uint8_t* frame = G.frames + L.f * G.frame_stride;
    const uint8_t* pA = frame + (long long)(L.byA * 8) * G.row_stride + L.bxA * (8 * CH);
    const uint8_t* pB = frame + (long long)(L.byB * 8) * G.row_stride + L.bxB * (8 * CH);
    uint2 raw[16 * P];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
        for (int j = 0; j < P; ++j) {
            raw[r * P + j] = __ldg(reinterpret_cast<const uint2*>(pA) + j);
            raw[(8 + r) * P + j] = __ldg(reinterpret_cast<const uint2*>(pB) + j);
        }
        step(pA, G.row_stride);
        step(pB, G.row_stride);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        row_to_gray<CH>(raw + r * P, gA[2 * r], gA[2 * r + 1]);
        row_to_gray<CH>(raw + (8 + r) * P, gB[2 * r], gB[2 * r + 1]);
    }
}

// Columns 2*CP and 2*CP+1 of both blocks: bytes -> floats, axis-0 transform, into chunks 4u+CP.
template <int CP>
__device__ __forceinline__ void column_pair_fwd(const PackedOps& ops, const uint32_t (&gA)[16], const uint32_t (&gB)[16],
                                                uint32_t magic_hi, ulonglong2* mine)
{
    const P2 unbias = pk(-8388608.0f, -8388608.0f);
    constexpr int C0 = 2 * CP, C1 = 2 * CP + 1;
    P2 c0[8], c1[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        c0[r] = add2(pku(magic_byte<(C0 & 3)>(gA[2 * r + (C0 >> 2)], magic_hi), magic_byte<(C0 & 3)>(gB[2 * r + (C0 >> 2)], magic_hi)), unbias);
        c1[r] = add2(pku(magic_byte<(C1 & 3)>(gA[2 * r + (C1 >> 2)], magic_hi), magic_byte<(C1 & 3)>(gB[2 * r + (C1 >> 2)], magic_hi)), unbias);
    }
    svs::dct8_fwd<1>(ops, c0);
    svs::dct8_fwd<1>(ops, c1);
#pragma unroll
    for (int u = 0; u < 8; ++u) st_chunk(mine, 4 * u + CP, c0[u], c1[u]);
}

// ------------------------------------------------------------------------------------------
// embed: every block of every frame handled here is completely filled with payload (k == n)
// ------------------------------------------------------------------------------------------
template <int CH, int OUT_CH, bool NFULL>
__global__ void __launch_bounds__(kTileThreads, kTileCtasPerSm) embed_tile_kernel(const FastEmbedArgs a)
{
    extern __shared__ ulonglong2 tile_smem[];
    ulonglong2* mine = tile_smem + threadIdx.x;
    const FastGeom& G = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = NFULL ? 63 : G.n;
    PackedOps ops;
    ops.negzero = pk(a.q.negzero, a.q.negzero);
    const P2 r2 = pk(a.q.r2, a.q.r2), ke = pk(a.q.ke, a.q.ke), d2 = pk(a.q.d2, a.q.d2), k0 = pk(a.q.k0, a.q.k0);
    const uint32_t emask = a.q.emask, ebit = a.q.ebit;
    const int sh0 = a.q.erot - 7;                   // row byte (coefficient v at bit 7-v) -> bit erot

    for (long long g = (long long)blockIdx.x * kTileWarps + warp; g < G.total_groups; g += (long long)gridDim.x * kTileWarps) {
        const Lane L = locate(G, g, lane);
        if (L.base + lane == 0 && a.bits_embedded != nullptr) a.bits_embedded[L.f] = a.cap;
        {
            uint32_t gA[16], gB[16];
            load_gray<CH>(G, L, gA, gB);
            column_pair_fwd<0>(ops, gA, gB, G.magic_hi, mine);
            column_pair_fwd<1>(ops, gA, gB, G.magic_hi, mine);
            column_pair_fwd<2>(ops, gA, gB, G.magic_hi, mine);
            column_pair_fwd<3>(ops, gA, gB, G.magic_hi, mine);
        }

        // 64-bit payload windows of the two blocks: stream bit i at bit 63-i
        uint32_t wA0, wA1, wB0, wB1;
        {
            const long long at = a.payload_bit_offset + L.f * a.cap;
            payload_window(a.payload, a.payload_last_word, at + (long long)L.bA * n, wA0, wA1);
            payload_window(a.payload, a.payload_last_word, at + (long long)L.bB * n, wB0, wB1);
        }
        const u64 vA = ((u64)wA0 << 32) | wA1, vB = ((u64)wB0 << 32) | wB1;

        // axis-1 transform of row u, then its quantisation
#pragma unroll 1
        for (int u = 0; u < 8; ++u) {
            P2 x[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) ld_chunk(mine, 4 * u + j, x[2 * j], x[2 * j + 1]);
            svs::dct8_fwd<1>(ops, x);
            if (NFULL || 8 * u - 1 < n) {
                // the 8 payload bits of this row (stream bits 8u-1 .. 8u+6), coefficient v at bit 7-v
                const uint32_t rbA = ((uint32_t)(u == 0 ? vA >> 57 : vA >> (57 - 8 * u)) & (u == 0 ? 0x7fu : 0xffu)) << sh0;
                const uint32_t rbB = ((uint32_t)(u == 0 ? vB >> 57 : vB >> (57 - 8 * u)) & (u == 0 ? 0x7fu : 0xffu)) << sh0;
                uint32_t worst = 0xffffffffu;
                P2 nx[8];
#pragma unroll
                for (int v = 0; v < 8; ++v) {
                    const int i = 8 * u + v - 1;
                    const bool coded = (v > 0 || u > 0) && (NFULL || i < n);
                    const P2 y = fma2(x[v], r2, ke);
                    uint32_t ya, yb;
                    unpk(y, ya, yb);
                    const uint32_t la = coded ? (ya & emask) : 0xffffffffu, lb = coded ? (yb & emask) : 0xffffffffu;
                    worst = min(worst, min(la, lb));
                    const uint32_t ta = (rbA << v) & ebit, tb = (rbB << v) & ebit;
                    // M + floor() + bit/2, then (2e + bit) * delta in one rounding
                    const P2 q = fma2(pku((ya & ~emask) | ta, (yb & ~emask) | tb), d2, k0);
                    nx[v] = coded ? q : x[v];
                }
                if (worst < kZone) {                              // rare: a fraction too close to call
                    P2 orig[8], res[8];
#pragma unroll
                    for (int v = 0; v < 8; ++v) { orig[v] = x[v]; res[v] = nx[v]; }
                    fix_row_embed(orig, res, u, n, G.delta32, a.q.r, a.q.r2, a.q.ke, emask, wA0, wA1, wB0, wB1);
#pragma unroll
                    for (int v = 0; v < 8; ++v) nx[v] = res[v];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) st_chunk(mine, 4 * u + j, nx[2 * j], nx[2 * j + 1]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) st_chunk(mine, 4 * u + j, x[2 * j], x[2 * j + 1]);
            }
        }

        // inverse, axis 0: columns 2cp and 2cp+1
#pragma unroll 1
        for (int cp = 0; cp < 4; ++cp) {
            P2 c0[8], c1[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) ld_chunk(mine, 4 * u + cp, c0[u], c1[u]);
            svs::dct8_inv<1>(ops, c0);
            svs::dct8_inv<1>(ops, c1);
#pragma unroll
            for (int u = 0; u < 8; ++u) st_chunk(mine, 4 * u + cp, c0[u], c1[u]);
        }

        // inverse, axis 1, clip + truncate, store
        uint8_t* out = a.stego + L.f * a.stego_frame_stride;
        uint8_t* dstA = out + (long long)(L.byA * 8) * a.stego_row_stride + L.bxA * (8 * OUT_CH);
        uint8_t* dstB = out + (long long)(L.byB * 8) * a.stego_row_stride + L.bxB * (8 * OUT_CH);
#pragma unroll 1
        for (int r = 0; r < 8; ++r) {
            P2 x[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) ld_chunk(mine, 4 * r + j, x[2 * j], x[2 * j + 1]);
            svs::dct8_inv<1>(ops, x);
            uint32_t ba[8], bb[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float va, vb;
                unpkf(x[c], va, vb);
                ba[c] = to_u8(va);
                bb[c] = to_u8(vb);
            }
            const uint32_t a0 = pack4(ba[0], ba[1], ba[2], ba[3]), a1 = pack4(ba[4], ba[5], ba[6], ba[7]);
            const uint32_t b0 = pack4(bb[0], bb[1], bb[2], bb[3]), b1 = pack4(bb[4], bb[5], bb[6], bb[7]);
            if (L.okA) store_row<OUT_CH>(dstA, a0, a1);
            if (L.okB) store_row<OUT_CH>(dstB, b0, b1);
            step(dstA, a.stego_row_stride);
            step(dstB, a.stego_row_stride);
        }
    }
}

// ------------------------------------------------------------------------------------------
// extract
// ------------------------------------------------------------------------------------------
template <int CH, bool NFULL>
__global__ void __launch_bounds__(kTileThreads, kTileCtasPerSm) extract_tile_kernel(const FastExtractArgs a)
{
    extern __shared__ ulonglong2 tile_smem[];
    ulonglong2* mine = tile_smem + threadIdx.x;
    const FastGeom& G = a.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = NFULL ? 63 : G.n;
    PackedOps ops;
    ops.negzero = pk(a.q.negzero, a.q.negzero);
    const P2 rr = pk(a.q.r, a.q.r), kx = pk(a.q.kx, a.q.kx);
    const uint32_t xmask = a.q.xmask;
    const int xk = a.q.xk;
    // the warp's 32 chunk-0 slots are 512 contiguous bytes: reused as the 128-word bit-packing area
    uint32_t* pack = reinterpret_cast<uint32_t*>(tile_smem + warp * 32);

    for (long long g = (long long)blockIdx.x * kTileWarps + warp; g < G.total_groups; g += (long long)gridDim.x * kTileWarps) {
        const Lane L = locate(G, g, lane);
        {
            uint32_t gA[16], gB[16];
            load_gray<CH>(G, L, gA, gB);
            column_pair_fwd<0>(ops, gA, gB, G.magic_hi, mine);
            column_pair_fwd<1>(ops, gA, gB, G.magic_hi, mine);
            column_pair_fwd<2>(ops, gA, gB, G.magic_hi, mine);
            column_pair_fwd<3>(ops, gA, gB, G.magic_hi, mine);
        }
        u64 vA = 0, vB = 0;                                    // stream bit i of the block at bit 63-i
#pragma unroll 1
        for (int u = 0; u < 8; ++u) {
            if (!(NFULL || 8 * u - 1 < n)) break;
            P2 x[8];
#pragma unroll
            for (int j = 0; j < 4; ++j) ld_chunk(mine, 4 * u + j, x[2 * j], x[2 * j + 1]);
            svs::dct8_fwd<1>(ops, x);
            uint32_t worst = 0xffffffffu;
            uint32_t rowA = 0, rowB = 0;                        // coefficient v of this row at bit 7-v
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const int i = 8 * u + v - 1;
                const bool coded = (v > 0 || u > 0) && (NFULL || i < n);
                const P2 y = fma2(x[v], rr, kx);
                uint32_t ya, yb;
                unpk(y, ya, yb);
                worst = min(worst, coded ? min(ya & xmask, yb & xmask) : 0xffffffffu);
                const int rot = (7 - v - xk) & 31;              // parity (bit xk) -> bit 7-v
                const uint32_t m = coded ? (0x80u >> v) : 0u;
                rowA |= __funnelshift_l(ya, ya, rot) & m;
                rowB |= __funnelshift_l(yb, yb, rot) & m;
            }
            if (worst < kZone) {
                P2 in[8];
#pragma unroll
                for (int v = 0; v < 8; ++v) in[v] = x[v];
                const uint32_t both = fix_row_extract(in, rowA | (rowB << 16), u, n, G.delta32, a.q.r, a.q.kx, xmask);
                rowA = both & 0xffu;
                rowB = both >> 16;
            }
            // bit 7-v of the row -> stream bit 8u+v-1 -> bit 64-8u-v (v = 0 of row 0 is the DC: always 0)
            vA |= u == 0 ? (u64)rowA << 57 : (u64)rowA << (57 - 8 * u);
            vB |= u == 0 ? (u64)rowB << 57 : (u64)rowB << (57 - 8 * u);
        }
        if (!L.okA) vA = 0;
        if (!L.okB) vB = 0;
        __syncwarp();                                          // every lane is done with its chunks
        pack[lane] = 0; pack[lane + 32] = 0; pack[lane + 64] = 0; pack[lane + 96] = 0;
        __syncwarp();
        {
            // place the two n-bit strings at bit offsets lane*n and (32+lane)*n of the warp's run
            uint32_t hi = (uint32_t)(vA >> 32), lo = (uint32_t)vA;
            uint32_t o = (uint32_t)lane * (uint32_t)n, w0 = o >> 5, sh = o & 31;
            uint32_t p0 = hi >> sh, p1 = __funnelshift_r(lo, hi, sh), p2 = __funnelshift_r(0u, lo, sh);
            if (p0) atomicOr(pack + w0, p0);
            if (p1) atomicOr(pack + w0 + 1, p1);
            if (p2) atomicOr(pack + w0 + 2, p2);
            hi = (uint32_t)(vB >> 32); lo = (uint32_t)vB;
            o = (uint32_t)(32 + lane) * (uint32_t)n; w0 = o >> 5; sh = o & 31;
            p0 = hi >> sh; p1 = __funnelshift_r(lo, hi, sh); p2 = __funnelshift_r(0u, lo, sh);
            if (p0) atomicOr(pack + w0, p0);
            if (p1) atomicOr(pack + w0 + 1, p1);
            if (p2) atomicOr(pack + w0 + 2, p2);
        }
        __syncwarp();
        const int nblk = min(64, G.bpf - L.base);
        const int nwords = (nblk * n + 31) >> 5;
        uint32_t* o32 = reinterpret_cast<uint32_t*>(a.bits + L.f * a.bits_frame_stride + (long long)(L.base >> 5) * (4 * n));
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (lane + 32 * j < nwords) o32[lane + 32 * j] = bswap(pack[lane + 32 * j]);
        __syncwarp();
    }
}

}  // namespace tile
