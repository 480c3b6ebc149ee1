// svs_b200.cu - sm_100a kernels + C ABI (include/svs_b200.h) for the per-frame 8x8 block-DCT +
// parity-QIM path of the reference: proses_frame_qim_dct, config_and_setup.py:106-174.
//
// Mapping: one thread owns one 8x8 block.  The whole block lives in 64 FP32 registers, both 2-D
// transforms run in-register (no shuffles, no shared-memory transposes), and the per-block
// payload window is one funnel-shifted 64-bit word.  A warp owns 32 consecutive blocks of a
// frame in raster order, so that
//   * stego rows leave as fully coalesced 256-byte STG.64 runs,
//   * BGR rows arrive as three LDG.64 per lane over one 768-byte contiguous span (every sector
//     fetched once; the 2nd/3rd load of a row hit L1),
//   * the 32 x n extracted bits of a warp form 4n contiguous, 4-byte aligned output bytes.
// Arithmetic: see svs_math.cuh - op-exact float32, no FMA contraction, IEEE division, half-even
// rounding, clip-then-truncate (config_and_setup.py:135,148-158,168,171).
#include <cuda_runtime.h>

#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#include "svs_b200.h"
#include "svs_math.cuh"
#include "svs_quant.h"
#if defined(SVS_WITH_VARIANTS)
#include "variants/svs_fast.cuh"      // round-1 lockstep kernels (two blocks per thread)
#include "variants/svs_tile.cuh"      // measured dead ends, kept for the A/B script (profiles/build_variant.sh)
#include "variants/svs_row.cuh"
#endif
#include "svs_block.cuh"

namespace {

constexpr int kThreads = 128;           // blocks (8x8) per CTA = threads per CTA
constexpr int kWarps = kThreads / 32;

// ------------------------------------------------------------------------------------------
// parameters
// ------------------------------------------------------------------------------------------
struct Geometry {
    const uint8_t* frames;
    long long frame_stride, row_stride;
    int H, W, bw, bpf;          // bw = blocks per block-row, bpf = blocks per frame
    int tiles_per_frame;        // ceil(bpf / kThreads)
    int n;                      // coefficients per block, 0..63
    float delta32;
    double delta;
};

struct EmbedArgs {
    Geometry g;
    const uint32_t* payload;    // 4-byte aligned, MSB-first bit stream
    long long payload_bit_offset, payload_total_bits, payload_last_word;
    long long cap;              // bits per frame (0 when not active)
    int active;                 // n > 0 && delta > 0
    uint8_t* stego;
    long long stego_frame_stride, stego_row_stride;
    uint8_t* gray;
    int64_t* bits_embedded;
    unsigned long long* sse;
};

struct ExtractArgs {
    Geometry g;
    uint8_t* bits;
    long long bits_frame_stride;
    long long frame_bytes;      // ceil(cap/8)
    int positive_delta;
};

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

// gray of the pixel whose B,G,R are bytes 0,1,2 of `px`:  (3735B + 19235G + 9798R + 16384) >> 15
__device__ __forceinline__ uint32_t gray_of_word(uint32_t px)
{
    uint32_t s = __dp2a_lo((19235u << 16) | 3735u, px, 16384u);   // B*3735 + G*19235 + round
    s = __dp2a_hi(9798u, px, s);                                   // + R*9798 (+ byte3 * 0)
    return s >> 15;
}

// Loads one 8x8 block as packed gray bytes: g[2r], g[2r+1] = row r, pixels 0..3 / 4..7.
template <int CH, bool ALIGNED>
__device__ __forceinline__ void load_block_gray(const uint8_t* __restrict__ p, long long row_stride,
                                                uint32_t (&g)[16])
{
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const uint8_t* row = p + r * row_stride;
        if (CH == 1) {
            if (ALIGNED) {
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(row));
                g[2 * r] = v.x;
                g[2 * r + 1] = v.y;
            } else {
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    lo |= (uint32_t)__ldg(row + c) << (8 * c);
                    hi |= (uint32_t)__ldg(row + 4 + c) << (8 * c);
                }
                g[2 * r] = lo;
                g[2 * r + 1] = hi;
            }
        } else {
            uint32_t w[7];
            if (ALIGNED) {
                const uint2 a = __ldg(reinterpret_cast<const uint2*>(row));
                const uint2 b = __ldg(reinterpret_cast<const uint2*>(row) + 1);
                const uint2 c = __ldg(reinterpret_cast<const uint2*>(row) + 2);
                w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y; w[4] = c.x; w[5] = c.y;
            } else {
#pragma unroll
                for (int k = 0; k < 6; ++k) {
                    uint32_t v = 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) v |= (uint32_t)__ldg(row + 4 * k + c) << (8 * c);
                    w[k] = v;
                }
            }
            w[6] = 0;
            uint32_t out[2] = {0u, 0u};
#pragma unroll
            for (int px = 0; px < 8; ++px) {
                const int byte0 = 3 * px, wi = byte0 >> 2, off = byte0 & 3;
                // bytes off..off+3 of the pair (w[wi], w[wi+1]); byte 3 is multiplied by 0
                const uint32_t sel = (uint32_t)(off | ((off + 1) << 4) | ((off + 2) << 8) | ((off + 3) << 12));
                const uint32_t word = off == 0 ? w[wi] : __byte_perm(w[wi], w[wi + 1], sel);
                out[px >> 2] |= gray_of_word(word) << (8 * (px & 3));
            }
            g[2 * r] = out[0];
            g[2 * r + 1] = out[1];
        }
    }
}

__device__ __forceinline__ void unpack_gray(const uint32_t (&g)[16], float (&x)[64])
{
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c)
            x[r * 8 + c] = (float)((g[2 * r + (c >> 2)] >> (8 * (c & 3))) & 0xffu);
}

// 64-bit window of the MSB-first payload starting at absolute bit `pos`:
// bit i of the window (i = 0 first) is (hi >> (31-i)) & 1 for i < 32, (lo >> (63-i)) & 1 after.
__device__ __forceinline__ void payload_window(const uint32_t* __restrict__ words, long long last_word,
                                               long long pos, uint32_t& hi, uint32_t& lo)
{
    const long long wi = pos >> 5;
    const uint32_t s = (uint32_t)(pos & 31);
    const uint32_t w0 = wi <= last_word ? bswap32(__ldg(words + wi)) : 0u;
    const uint32_t w1 = wi + 1 <= last_word ? bswap32(__ldg(words + wi + 1)) : 0u;
    const uint32_t w2 = wi + 2 <= last_word ? bswap32(__ldg(words + wi + 2)) : 0u;
    hi = __funnelshift_l(w1, w0, s);
    lo = __funnelshift_l(w2, w1, s);
}

// ------------------------------------------------------------------------------------------
// embed
// ------------------------------------------------------------------------------------------
template <int CH, int OUT_CH, bool ALIGNED, bool F64>
__global__ void __launch_bounds__(kThreads, 4) embed_kernel(const EmbedArgs a)
{
    const Geometry& G = a.g;
    const long long f = blockIdx.x / G.tiles_per_frame;
    const int tile = (int)(blockIdx.x - f * G.tiles_per_frame);
    const int b = tile * kThreads + (int)threadIdx.x;          // block index inside the frame
    const bool valid = b < G.bpf;

    if (b == 0 && a.bits_embedded != nullptr) {
        long long left = a.payload_total_bits - f * a.cap;
        left = left < 0 ? 0 : (left > a.cap ? a.cap : left);
        a.bits_embedded[f] = a.active ? left : 0;
    }

    unsigned sse = 0;
    if (valid) {
        const int by = b / G.bw, bx = b - by * G.bw;
        const uint8_t* src = G.frames + f * G.frame_stride + (long long)(by * 8) * G.row_stride + (long long)bx * (8 * CH);
        uint32_t g[16];
        load_block_gray<CH, ALIGNED>(src, G.row_stride, g);

        // how much of the payload reaches this block (config_and_setup.py:130-132,141)
        const long long used = a.active ? f * a.cap + (long long)b * G.n : 0;   // bits before this block
        const long long left = a.payload_total_bits - used;
        const bool process = left > 0;
        const int k = a.active ? (left < G.n ? (int)(left < 0 ? 0 : left) : G.n) : 0;

        uint32_t s[16];
        if (process) {
            float x[64];
            unpack_gray(g, x);
            svs::dct2_fwd(svs::ScalarOps(), x);
            if (k > 0) {
                uint32_t hi, lo;
                payload_window(a.payload, a.payload_last_word, a.payload_bit_offset + used, hi, lo);
                const float d32 = G.delta32;
#pragma unroll
                for (int i = 0; i < 63; ++i) {
                    if (i < k) {
                        const uint32_t bit = i < 32 ? (hi >> (31 - i)) & 1u : (lo >> (63 - i)) & 1u;
                        const float t = __fdiv_rn(x[i + 1], d32);             // float32 division (:148)
                        const int q = __float2int_rn(t);                      // round half to even
                        const int qn = q - (q & 1) + (int)bit;                // parity fix-up (:149-155)
                        x[i + 1] = F64 ? (float)((double)qn * G.delta)        // float(q*delta) -> f32 (:156)
                                       : __fmul_rn((float)qn, d32);
                    }
                }
            }
            svs::dct2_inv(svs::ScalarOps(), x);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                uint32_t lo4 = 0, hi4 = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    // np.uint8(np.clip(v, 0, 255)): clip, then truncate toward zero (:171)
                    lo4 |= __float2uint_rz(fminf(fmaxf(x[r * 8 + c], 0.0f), 255.0f)) << (8 * c);
                    hi4 |= __float2uint_rz(fminf(fmaxf(x[r * 8 + 4 + c], 0.0f), 255.0f)) << (8 * c);
                }
                s[2 * r] = lo4;
                s[2 * r + 1] = hi4;
            }
            if (a.sse != nullptr) {
#pragma unroll
                for (int j = 0; j < 16; ++j)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int d = (int)((s[j] >> (8 * c)) & 0xffu) - (int)((g[j] >> (8 * c)) & 0xffu);
                        sse += (unsigned)(d * d);
                    }
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) s[j] = g[j];
        }

        uint8_t* dst = a.stego + f * a.stego_frame_stride + (long long)(by * 8) * a.stego_row_stride + (long long)bx * (8 * OUT_CH);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            uint2* row = reinterpret_cast<uint2*>(dst + r * a.stego_row_stride);
            const uint32_t lo4 = s[2 * r], hi4 = s[2 * r + 1];
            if (OUT_CH == 1) {
                row[0] = make_uint2(lo4, hi4);
            } else {            // gray replicated to B,G,R: cv2.cvtColor(GRAY2BGR), embed_process.py:126
                row[0] = make_uint2(__byte_perm(lo4, 0, 0x1000), __byte_perm(lo4, 0, 0x2211));
                row[1] = make_uint2(__byte_perm(lo4, 0, 0x3332), __byte_perm(hi4, 0, 0x1000));
                row[2] = make_uint2(__byte_perm(hi4, 0, 0x2211), __byte_perm(hi4, 0, 0x3332));
            }
        }
        if (a.gray != nullptr) {
            uint8_t* gd = a.gray + ((f * G.H + by * 8) * (long long)G.W) + bx * 8;
#pragma unroll
            for (int r = 0; r < 8; ++r)
                *reinterpret_cast<uint2*>(gd + (long long)r * G.W) = make_uint2(g[2 * r], g[2 * r + 1]);
        }
    }
    if (a.sse != nullptr) {
        sse = __reduce_add_sync(0xffffffffu, sse);
        if ((threadIdx.x & 31) == 0 && sse != 0) atomicAdd(a.sse + f, (unsigned long long)sse);
    }
}

// ------------------------------------------------------------------------------------------
// extract
// ------------------------------------------------------------------------------------------
template <int CH, bool ALIGNED, bool WORD_STORES>
__global__ void __launch_bounds__(kThreads, 4) extract_kernel(const ExtractArgs a)
{
    __shared__ uint32_t pack[kWarps][64];
    const Geometry& G = a.g;
    const long long f = blockIdx.x / G.tiles_per_frame;
    const int tile = (int)(blockIdx.x - f * G.tiles_per_frame);
    const int b = tile * kThreads + (int)threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool valid = b < G.bpf;
    const int n = G.n;

    pack[warp][lane] = 0;
    pack[warp][lane + 32] = 0;
    __syncwarp();

    if (valid && a.positive_delta) {
        const int by = b / G.bw, bx = b - by * G.bw;
        const uint8_t* src = G.frames + f * G.frame_stride + (long long)(by * 8) * G.row_stride + (long long)bx * (8 * CH);
        uint32_t g[16];
        load_block_gray<CH, ALIGNED>(src, G.row_stride, g);
        float x[64];
        unpack_gray(g, x);
        svs::dct2_fwd(svs::ScalarOps(), x);
        uint32_t hi = 0, lo = 0;                 // bit i of this block at (hi:lo) bit 63-i
        const float d32 = G.delta32;
#pragma unroll
        for (int i = 0; i < 63; ++i) {
            if (i < n) {
                const float t = __fdiv_rn(x[i + 1], d32);
                const uint32_t par = (uint32_t)__float2int_rn(t) & 1u;    // int(round(t)) % 2 (:160-161)
                if (i < 32) hi |= par << (31 - i); else lo |= par << (63 - i);
            }
        }
        // place the n bits at bit offset lane*n of the warp's 32n-bit (= n words) run
        const uint32_t o = (uint32_t)lane * (uint32_t)n;
        const uint32_t w0 = o >> 5, sh = o & 31;
        const uint32_t p0 = hi >> sh;
        const uint32_t p1 = __funnelshift_r(lo, hi, sh);
        const uint32_t p2 = __funnelshift_r(0u, lo, sh);
        if (p0) atomicOr(&pack[warp][w0], p0);
        if (p1) atomicOr(&pack[warp][w0 + 1], p1);
        if (p2) atomicOr(&pack[warp][w0 + 2], p2);
    }
    __syncwarp();

    // the warp's run: blocks [wb, wb+32) -> bits [wb*n, ...) -> bytes from (wb/32)*4n
    const int wb = tile * kThreads + warp * 32;
    if (wb < G.bpf) {
        const int nblk = min(32, G.bpf - wb);
        const int nbits = nblk * n;
        uint8_t* out = a.bits + f * a.bits_frame_stride + (long long)(wb >> 5) * (4 * n);
        if (WORD_STORES) {
            const int nwords = (nbits + 31) >> 5;
            uint32_t* o32 = reinterpret_cast<uint32_t*>(out);
            if (lane < nwords) o32[lane] = bswap32(pack[warp][lane]);
            if (lane + 32 < nwords) o32[lane + 32] = bswap32(pack[warp][lane + 32]);
        } else {
            const int nbytes = (nbits + 7) >> 3;
            for (int j = lane; j < nbytes; j += 32)
                out[j] = (uint8_t)(pack[warp][j >> 2] >> (24 - 8 * (j & 3)));
        }
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
#if defined(SVS_WITH_VARIANTS)
// ------------------------------------------------------------------------------------------
// Side outputs of embed for the frames the packed kernels handle: the gray reference (first
// return value of the reference function, config_and_setup.py:111-114,172) and the per-frame
// sum of squared errors gray vs stego (what cv2.PSNR needs, embed_process.py:204-206).  The
// packed embed kernels have no registers left for either, and a separate streaming pass
// (4 + 1 bytes read, 1 written per pixel, pure HBM) on top of them is faster than the scalar
// embed kernel that produces them in its epilogue.  16 pixels per thread, W % 16 == 0.
// ------------------------------------------------------------------------------------------
struct SideArgs {
    const uint8_t* frames;
    long long frame_stride, row_stride;
    const uint8_t* stego;                   // 1 channel
    long long stego_frame_stride, stego_row_stride;
    uint8_t* gray;                          // nullable, contiguous frames
    unsigned long long* sse;                // nullable
    int H, W16;                             // rows, 16-pixel chunks per row
};

template <int CH>
__global__ void __launch_bounds__(256) side_outputs_kernel(const SideArgs a)
{
    const long long f = blockIdx.y;
    const int per_frame = a.H * a.W16;
    unsigned sse = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_frame; i += gridDim.x * blockDim.x) {
        const int y = i / a.W16, c = i - y * a.W16;
        const uint8_t* src = a.frames + f * a.frame_stride + (long long)y * a.row_stride + (long long)c * (16 * CH);
        uint32_t g[4];
        if (CH == 1) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(src));
            g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
        } else {
            const uint4 v0 = __ldg(reinterpret_cast<const uint4*>(src));
            const uint4 v1 = __ldg(reinterpret_cast<const uint4*>(src) + 1);
            const uint4 v2 = __ldg(reinterpret_cast<const uint4*>(src) + 2);
            const uint32_t w[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
            blk::row_gray_words(w, g[0], g[1]);
            blk::row_gray_words(w + 6, g[2], g[3]);
        }
        if (a.gray != nullptr)
            *reinterpret_cast<uint4*>(a.gray + (f * a.H + y) * (long long)(a.W16 * 16) + c * 16) = make_uint4(g[0], g[1], g[2], g[3]);
        if (a.sse != nullptr) {
            const uint4 s = __ldg(reinterpret_cast<const uint4*>(a.stego + f * a.stego_frame_stride + (long long)y * a.stego_row_stride + c * 16));
            const uint32_t sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t d = __vabsdiffu4(g[k], sv[k]);
                sse = __dp4a(d, d, sse);                    // sum of the four squared byte differences
            }
        }
    }
    if (a.sse != nullptr) {
        __shared__ unsigned long long part[8];
        unsigned long long t = sse;                         // per thread < 2^32 (host caps the chunks per thread)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = t;
        __syncthreads();
        if (threadIdx.x == 0) {
            t = 0;
            for (int w = 0; w < 8; ++w) t += part[w];
            if (t) atomicAdd(a.sse + f, t);
        }
    }
}

#endif  // SVS_WITH_VARIANTS

thread_local char g_err[256] = "";
std::atomic<long long> g_launches{0};
std::atomic<int> g_reserved_sms{0};

int fail(int code, const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what)
{
    snprintf(g_err, sizeof g_err, "%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

inline bool aligned_to(const void* p, long long a, long long b, int m)
{
    return ((reinterpret_cast<uintptr_t>(p) | (uintptr_t)a | (uintptr_t)b) & (uintptr_t)(m - 1)) == 0;
}

int check_geometry(const void* frames, int channels, long long n_frames, int H, int W,
                   long long frame_stride, long long row_stride, double delta)
{
    if (channels != 1 && channels != 3) return fail(SVS_ERR_SHAPE, "channels must be 1 or 3 (got %d)", channels);
    if (H <= 0 || W <= 0 || (H % 8) || (W % 8))
        return fail(SVS_ERR_SHAPE, "height and width must be positive multiples of 8 (got %dx%d)", H, W);
    if (n_frames < 0) return fail(SVS_ERR_SHAPE, "n_frames < 0");
    if (n_frames > 0 && frames == nullptr) return fail(SVS_ERR_POINTER, "frames is NULL");
    if (row_stride < (long long)W * channels) return fail(SVS_ERR_STRIDE, "row_stride smaller than a row");
    if (n_frames > 1 && frame_stride < (long long)(H - 1) * row_stride + (long long)W * channels)
        return fail(SVS_ERR_STRIDE, "frame_stride smaller than a frame");
    if (std::isnan(delta) || std::isinf(delta)) return fail(SVS_ERR_DELTA, "delta is not finite");
    if (delta > 0 && delta < 0x1p-10) return fail(SVS_ERR_DELTA, "0 < delta < 2^-10 is outside the quantiser range");
    return SVS_OK;
}

Geometry make_geometry(const uint8_t* frames, int H, int W, long long frame_stride, long long row_stride,
                       double delta, int num_ac)
{
    Geometry g;
    g.frames = frames;
    g.frame_stride = frame_stride;
    g.row_stride = row_stride;
    g.H = H;
    g.W = W;
    g.bw = W / 8;
    g.bpf = (H / 8) * (W / 8);
    g.tiles_per_frame = (g.bpf + kThreads - 1) / kThreads;
    g.n = num_ac < 0 ? 0 : (num_ac > SVS_MAX_AC ? SVS_MAX_AC : num_ac);
    g.delta = delta;
    g.delta32 = (float)delta;
    return g;
}

template <int CH, int OUT_CH, bool ALIGNED>
void launch_embed(const EmbedArgs& a, bool f64, unsigned grid, cudaStream_t st)
{
    if (f64) embed_kernel<CH, OUT_CH, ALIGNED, true><<<grid, kThreads, 0, st>>>(a);
    else embed_kernel<CH, OUT_CH, ALIGNED, false><<<grid, kThreads, 0, st>>>(a);
}

template <int CH, bool ALIGNED>
void launch_extract(const ExtractArgs& a, bool words, unsigned grid, cudaStream_t st)
{
    if (words) extract_kernel<CH, ALIGNED, true><<<grid, kThreads, 0, st>>>(a);
    else extract_kernel<CH, ALIGNED, false><<<grid, kThreads, 0, st>>>(a);
}


using svs::make_fast_quant;     // svs_quant.h

// SMs the persistent grids may occupy (svs_set_reserved_sms leaves some to a concurrent collective)
int usable_sms()
{
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[dev] = v;
    }
    const int usable = sms[dev] - g_reserved_sms.load();
    return usable < 1 ? 1 : usable;
}

#if defined(SVS_WITH_VARIANTS)
fast::FastGeom make_fast_geometry(const Geometry& g, long long n_frames)
{
    fast::FastGeom f;
    f.frames = g.frames;
    f.frame_stride = g.frame_stride;
    f.row_stride = g.row_stride;
    f.H = g.H;
    f.W = g.W;
    f.bw = g.bw;
    f.bpf = g.bpf;
    f.n = g.n;
    f.groups_per_frame = (g.bpf + 63) / 64;
    f.total_groups = n_frames * f.groups_per_frame;
    f.delta32 = g.delta32;
    f.magic_hi = 0x4B000000u;
    f.bw_magic = (uint32_t)((0x100000000ull + (unsigned long long)g.bw - 1) / (unsigned long long)g.bw);
    return f;
}

// persistent grid: one CTA per SM (or fewer when there is not enough work)
unsigned fast_grid(long long total_groups)
{
    const long long want = (total_groups + fast::kFastWarps - 1) / fast::kFastWarps;
    const long long cap = (long long)usable_sms() * fast::kFastCtasPerSm;
    return (unsigned)(want < cap ? want : cap);
}
#endif

// Kernel family: 0 = automatic (block kernels when applicable, scalar otherwise), 1 = scalar only,
// 5 = packed block kernels (svs_block.cuh: one block per thread, free-running warps).  Builds
// with -DSVS_WITH_VARIANTS (profiles/build_variant.sh) also carry the round-1 organisations the
// A/B measurements compare against: 2 = packed lockstep (svs_fast.cuh: two blocks per thread),
// 3 = packed tile (svs_tile.cuh), 4 = packed row (svs_row.cuh: 8 lanes per block pair).
std::atomic<int> g_family{0};

bool family_available(int f)
{
#if defined(SVS_WITH_VARIANTS)
    return f >= 0 && f <= 5;
#else
    return f == 0 || f == 1 || f == 5;
#endif
}

int family() { return g_family.load(std::memory_order_relaxed); }

// ---- block kernels (svs_block.cuh): one block per thread, free-running warps -----------------
blk::Div make_div(uint32_t d)
{
    int l = 0;
    while ((1ull << l) < d) ++l;
    blk::Div r;
    r.shift = (uint32_t)(31 + l);
    r.mul = (uint32_t)(((1ull << (31 + l)) + d - 1) / d);        // < 2^32 because 2^l < 2 d
    return r;
}

blk::BlkGeom make_blk_geometry(const Geometry& g, long long n_frames)
{
    blk::BlkGeom b;
    b.frames = g.frames;
    b.frame_stride = g.frame_stride;
    b.row_stride = g.row_stride;
    b.frame_stride32 = (uint32_t)g.frame_stride;
    b.row_stride32 = (uint32_t)g.row_stride;
    b.bw = g.bw;
    b.bpf = g.bpf;
    b.n = g.n;
    b.gpf = (g.bpf + 31) / 32;
    b.total_groups = n_frames * b.gpf;
    b.by_gpf = make_div((uint32_t)b.gpf);
    b.by_bw = make_div((uint32_t)g.bw);
    b.magic_hi = 0x4B000000u;
    b.delta32 = g.delta32;
    return b;
}

// The block kernels address inside a frame with 32-bit offsets: the frame extent and (for more
// than one frame) the distance between frames have to fit.  Anything else - no real video comes
// close - goes to the scalar kernels.
bool blk_addressable(long long n_frames, int H, long long frame_stride, long long row_stride)
{
    const long long lim = 0xffffffffLL;
    return (long long)H * row_stride <= lim && (n_frames <= 1 || frame_stride <= lim);
}

// persistent grid: kBlkMinCtas CTAs on every usable SM (or fewer when there is not enough work)
unsigned blk_grid(long long total_groups)
{
    const long long ctas = (long long)usable_sms() * blk::kBlkMinCtas;
    const long long want = (total_groups + blk::kBlkWarps - 1) / blk::kBlkWarps;
    return (unsigned)(want < ctas ? want : ctas);
}

// opt in to the dynamic shared memory of the cp.async slots, once per kernel instantiation and
// device; `ready` belongs to the calling instantiation (all kernels share one pointer TYPE)
cudaError_t blk_smem_optin(const void* kernel, int bytes, bool (&ready)[64])
{
    if (bytes <= 0) return cudaSuccess;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    if (ready[dev]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) ready[dev] = true;
    return e;
}

template <int CH, int OC, bool NFULL, bool SIDE>
cudaError_t launch_blk_embed_one(const blk::BlkEmbedArgs& ba, cudaStream_t st)
{
    constexpr int smem = blk::blk_embed_smem_bytes<CH, SIDE>();
    static bool ready[64] = {false};
    if (cudaError_t e = blk_smem_optin(reinterpret_cast<const void*>(blk::embed_blk_kernel<CH, OC, NFULL, SIDE>), smem, ready)) return e;
    blk::embed_blk_kernel<CH, OC, NFULL, SIDE><<<blk_grid(ba.g.total_groups), blk::kBlkThreads, smem, st>>>(ba);
    return cudaGetLastError();
}

template <int CH, int OC>
cudaError_t launch_blk_embed(const blk::BlkEmbedArgs& ba, bool nfull, bool side, cudaStream_t st)
{
    if (nfull) return side ? launch_blk_embed_one<CH, OC, true, true>(ba, st) : launch_blk_embed_one<CH, OC, true, false>(ba, st);
    return side ? launch_blk_embed_one<CH, OC, false, true>(ba, st) : launch_blk_embed_one<CH, OC, false, false>(ba, st);
}

template <int CH, int NP>
cudaError_t launch_blk_extract_one(const blk::BlkExtractArgs& xa, cudaStream_t st)
{
    constexpr int smem = blk::blk_smem_bytes<CH, false>();
    static bool ready[64] = {false};
    if (cudaError_t e = blk_smem_optin(reinterpret_cast<const void*>(blk::extract_blk_kernel<CH, NP>), smem, ready)) return e;
    blk::extract_blk_kernel<CH, NP><<<blk_grid(xa.g.total_groups), blk::kBlkThreads, smem, st>>>(xa);
    return cudaGetLastError();
}

template <int CH>
cudaError_t launch_blk_extract(const blk::BlkExtractArgs& xa, cudaStream_t st)
{
    switch ((xa.g.n + 16) / 16) {                 // coefficient row pairs that hold any of flat 1..n
    case 1: return launch_blk_extract_one<CH, 1>(xa, st);
    case 2: return launch_blk_extract_one<CH, 2>(xa, st);
    case 3: return launch_blk_extract_one<CH, 3>(xa, st);
    default: return launch_blk_extract_one<CH, 4>(xa, st);
    }
}

#if defined(SVS_WITH_VARIANTS)
// row kernels: free-running warps, kRowCtasPerSm CTAs on every usable SM, grid-stride over groups
unsigned row_grid(long long total_groups)
{
    const long long ctas = (long long)fast_grid(1ll << 40) / fast::kFastCtasPerSm * row::kRowCtasPerSm;
    const long long want = (total_groups + row::kRowWarps - 1) / row::kRowWarps;
    return (unsigned)(want < ctas ? want : ctas);
}

// the row kernels pair horizontally adjacent blocks and use 128-bit loads / stores
bool row_ok(const Geometry& g, const void* frames, long long frame_stride, long long row_stride)
{
    return (g.bw % 2) == 0 && g.bw >= 8 && aligned_to(frames, frame_stride, row_stride, 16) &&
           (unsigned long long)g.bpf * (unsigned long long)g.bw < 0xffffffffull;
}

unsigned tile_grid(long long total_groups)
{
    const unsigned sm_ctas = fast_grid(1ll << 40) * tile::kTileCtasPerSm / fast::kFastCtasPerSm;
    const long long want = (total_groups + tile::kTileWarps - 1) / tile::kTileWarps;
    return (unsigned)(want < sm_ctas ? want : sm_ctas);
}

template <int CH, int OC, bool NFULL>
cudaError_t launch_tile_embed(const fast::FastEmbedArgs& fa, cudaStream_t st)
{
    static bool ready[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !ready[dev]) {
        cudaError_t e = cudaFuncSetAttribute(tile::embed_tile_kernel<CH, OC, NFULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile::kTileSmemBytes);
        if (e != cudaSuccess) return e;
        ready[dev] = true;
    }
    tile::embed_tile_kernel<CH, OC, NFULL><<<tile_grid(fa.g.total_groups), tile::kTileThreads, tile::kTileSmemBytes, st>>>(fa);
    return cudaGetLastError();
}

template <int CH, bool NFULL>
cudaError_t launch_tile_extract(const fast::FastExtractArgs& fa, cudaStream_t st)
{
    static bool ready[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !ready[dev]) {
        cudaError_t e = cudaFuncSetAttribute(tile::extract_tile_kernel<CH, NFULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, tile::kTileSmemBytes);
        if (e != cudaSuccess) return e;
        ready[dev] = true;
    }
    tile::extract_tile_kernel<CH, NFULL><<<tile_grid(fa.g.total_groups), tile::kTileThreads, tile::kTileSmemBytes, st>>>(fa);
    return cudaGetLastError();
}
#endif  // SVS_WITH_VARIANTS

}  // namespace

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" {

int svs_version(void) { return 200; }

const char* svs_last_error_string(void) { return g_err; }

int64_t svs_kernel_launch_count(void) { return g_launches.load(); }

int svs_set_reserved_sms(int n)
{
    const int prev = g_reserved_sms.load();
    if (n >= 0) g_reserved_sms.store(n);
    return prev;
}

int svs_debug_kernel_family(int family_id)
{
    if (family_id < 0) return g_family.load();
    if (!family_available(family_id)) return -1;
    return g_family.exchange(family_id);
}

int svs_memcpy_d2d_async(void* d_dst, const void* d_src, int64_t bytes, void* stream)
{
    g_err[0] = 0;
    if (bytes < 0 || (bytes > 0 && (d_dst == nullptr || d_src == nullptr))) return fail(SVS_ERR_POINTER, "svs_memcpy_d2d_async: bad arguments");
    if (bytes == 0) return SVS_OK;
    cudaError_t e = cudaMemcpyAsync(d_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "svs_memcpy_d2d_async");
    return SVS_OK;
}

int64_t svs_capacity_bits(int height, int width, int num_ac)
{
    if (height <= 0 || width <= 0) return 0;
    const int n = num_ac < 0 ? 0 : (num_ac > SVS_MAX_AC ? SVS_MAX_AC : num_ac);
    return (int64_t)(height / 8) * (width / 8) * n;
}

int64_t svs_bits_row_bytes(int height, int width, int num_ac)
{
    const int64_t cap = svs_capacity_bits(height, width, num_ac);
    const int64_t words = (cap + 31) / 32;
    return (words * 4 + 15) / 16 * 16;
}

static int extract_impl(const uint8_t* d_frames, int channels, int64_t n_frames,
                        int height, int width, int64_t frame_stride, int64_t row_stride,
                        double delta, int num_ac,
                        uint8_t* d_bits_out, int64_t bits_frame_stride,
                        uint8_t* const* peers, int n_peers, bool multicast, void* stream)
{
    g_err[0] = 0;
    if (int rc = check_geometry(d_frames, channels, n_frames, height, width, frame_stride, row_stride, delta)) return rc;
    ExtractArgs a;
    a.g = make_geometry(d_frames, height, width, frame_stride, row_stride, delta, num_ac);
    if (a.g.n == 0 || n_frames == 0) return SVS_OK;           // nothing is extracted (:138)
    if (d_bits_out == nullptr) return fail(SVS_ERR_POINTER, "bits_out is NULL");
    const long long cap = (long long)a.g.bpf * a.g.n;
    a.frame_bytes = (cap + 7) / 8;
    if (bits_frame_stride < a.frame_bytes) return fail(SVS_ERR_STRIDE, "bits_frame_stride < ceil(cap/8)");
    a.bits = d_bits_out;
    a.bits_frame_stride = bits_frame_stride;
    a.positive_delta = delta > 0;
    const long long grid = n_frames * a.g.tiles_per_frame;
    if (grid > 0x7fffffffLL) return fail(SVS_ERR_SHAPE, "batch too large for one launch; split it");
    const bool words = aligned_to(d_bits_out, bits_frame_stride, 0, 4) && bits_frame_stride >= (cap + 31) / 32 * 4;
    const bool al = aligned_to(d_frames, frame_stride, row_stride, 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const svs::FastQuant fq = make_fast_quant(delta);
    const int fam = family();
    if (fam != 1 && al && words && delta > 0 && fq.extract_ok && n_frames * ((a.g.bpf + 31) / 32) < 0x7fffffffLL &&
        blk_addressable(n_frames, height, frame_stride, row_stride)) {
        cudaError_t fe = cudaSuccess;
        if (fam == 0 || fam == 5) {
            blk::BlkExtractArgs xa;
            xa.g = make_blk_geometry(a.g, n_frames);
            xa.q = fq;
            xa.bits = d_bits_out;
            xa.bits_frame_stride = bits_frame_stride;
            xa.n_peers = n_peers;
            xa.multicast = multicast ? 1 : 0;
            for (int e = 0; e < n_peers; ++e) xa.peers[e] = peers[e];
            fe = channels == 3 ? launch_blk_extract<3>(xa, st) : launch_blk_extract<1>(xa, st);
        }
#if defined(SVS_WITH_VARIANTS)
        else {
            fast::FastExtractArgs fa;
            fa.g = make_fast_geometry(a.g, n_frames);
            fa.q = fq;
            fa.bits = d_bits_out;
            fa.bits_frame_stride = bits_frame_stride;
            fa.n_peers = fam == 3 ? 0 : n_peers;
            fa.multicast = multicast ? 1 : 0;
            for (int e = 0; e < fa.n_peers; ++e) fa.peers[e] = peers[e];
            const bool full = a.g.n == SVS_MAX_AC;
            if (fam == 4 && row_ok(a.g, d_frames, frame_stride, row_stride)) {
                const unsigned rgrid = row_grid(fa.g.total_groups);
                row::RowExtractArgs ra;
                ra.x = fa;
                ra.wrap_src = 8 * row_stride - (long long)a.g.bw * 8 * channels;
                if (channels == 3) {
                    if (full) row::extract_row_kernel<3, true><<<rgrid, row::kRowThreads, 0, st>>>(ra);
                    else row::extract_row_kernel<3, false><<<rgrid, row::kRowThreads, 0, st>>>(ra);
                } else {
                    if (full) row::extract_row_kernel<1, true><<<rgrid, row::kRowThreads, 0, st>>>(ra);
                    else row::extract_row_kernel<1, false><<<rgrid, row::kRowThreads, 0, st>>>(ra);
                }
                fe = cudaGetLastError();
            } else if (fam != 3) {
                const unsigned fgrid = fast_grid(fa.g.total_groups);
                if (channels == 3) {
                    if (full) fast::extract_fast_kernel<3, true><<<fgrid, fast::kFastThreads, 0, st>>>(fa);
                    else fast::extract_fast_kernel<3, false><<<fgrid, fast::kFastThreads, 0, st>>>(fa);
                } else {
                    if (full) fast::extract_fast_kernel<1, true><<<fgrid, fast::kFastThreads, 0, st>>>(fa);
                    else fast::extract_fast_kernel<1, false><<<fgrid, fast::kFastThreads, 0, st>>>(fa);
                }
                fe = cudaGetLastError();
            } else if (channels == 3) {
                fe = full ? launch_tile_extract<3, true>(fa, st) : launch_tile_extract<3, false>(fa, st);
            } else {
                fe = full ? launch_tile_extract<1, true>(fa, st) : launch_tile_extract<1, false>(fa, st);
            }
        }
#endif
        if (fe != cudaSuccess) return cuda_fail(fe, "svs_extract_frames launch (packed)");
    } else if (n_peers > 0) {
        return fail(SVS_ERR_ALIGNMENT, "svs_extract_frames_scatter needs the packed kernels (aligned input, padded bit rows, delta >= 1/16)");
    } else if (channels == 3) { if (al) launch_extract<3, true>(a, words, (unsigned)grid, st); else launch_extract<3, false>(a, words, (unsigned)grid, st); }
    else                      { if (al) launch_extract<1, true>(a, words, (unsigned)grid, st); else launch_extract<1, false>(a, words, (unsigned)grid, st); }
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "svs_extract_frames launch");
    return SVS_OK;
}

int svs_extract_frames(const uint8_t* d_frames, int channels, int64_t n_frames,
                       int height, int width, int64_t frame_stride, int64_t row_stride,
                       double delta, int num_ac,
                       uint8_t* d_bits_out, int64_t bits_frame_stride, void* stream)
{
    return extract_impl(d_frames, channels, n_frames, height, width, frame_stride, row_stride, delta, num_ac,
                        d_bits_out, bits_frame_stride, nullptr, 0, false, stream);
}

int svs_extract_frames_scatter(const uint8_t* d_frames, int channels, int64_t n_frames,
                               int height, int width, int64_t frame_stride, int64_t row_stride,
                               double delta, int num_ac,
                               uint8_t* d_bits_out, int64_t bits_frame_stride,
                               uint8_t* const* peer_bits_out, int n_peers, void* stream)
{
    g_err[0] = 0;
    if (n_peers < 0 || n_peers > blk::kMaxPeers) return fail(SVS_ERR_SHAPE, "n_peers must be 0..%d", blk::kMaxPeers);
    if (n_peers > 0 && peer_bits_out == nullptr) return fail(SVS_ERR_POINTER, "peer_bits_out is NULL");
    for (int e = 0; e < n_peers; ++e)
        if (peer_bits_out[e] == nullptr || !aligned_to(peer_bits_out[e], 0, 0, 4))
            return fail(SVS_ERR_ALIGNMENT, "peer buffer %d is NULL or not 4-byte aligned", e);
    if (family() == 1 || family() == 3) return fail(SVS_ERR_ALIGNMENT, "svs_extract_frames_scatter needs the packed kernels");
    return extract_impl(d_frames, channels, n_frames, height, width, frame_stride, row_stride, delta, num_ac,
                        d_bits_out, bits_frame_stride, peer_bits_out, n_peers, false, stream);
}

int svs_extract_frames_multicast(const uint8_t* d_frames, int channels, int64_t n_frames,
                                 int height, int width, int64_t frame_stride, int64_t row_stride,
                                 double delta, int num_ac,
                                 uint8_t* mc_bits_out, uint8_t* d_bits_local, int64_t bits_frame_stride, void* stream)
{
    g_err[0] = 0;
    if (mc_bits_out == nullptr || !aligned_to(mc_bits_out, 0, 0, 4))
        return fail(SVS_ERR_ALIGNMENT, "mc_bits_out is NULL or not 4-byte aligned");
    if (family() == 1 || family() == 3) return fail(SVS_ERR_ALIGNMENT, "svs_extract_frames_multicast needs the packed kernels");
    uint8_t* one[1] = {mc_bits_out};
    return extract_impl(d_frames, channels, n_frames, height, width, frame_stride, row_stride, delta, num_ac,
                        d_bits_local, bits_frame_stride, one, 1, true, stream);
}

int svs_embed_frames(const uint8_t* d_frames, int channels, int64_t n_frames,
                     int height, int width, int64_t frame_stride, int64_t row_stride,
                     const uint8_t* d_payload, int64_t payload_bit_offset, int64_t payload_total_bits,
                     double delta, int num_ac,
                     uint8_t* d_stego_out, int stego_channels,
                     int64_t stego_frame_stride, int64_t stego_row_stride,
                     uint8_t* d_gray_out, int64_t* d_bits_embedded_out,
                     unsigned long long* d_sse_out, void* stream)
{
    g_err[0] = 0;
    if (int rc = check_geometry(d_frames, channels, n_frames, height, width, frame_stride, row_stride, delta)) return rc;
    if (n_frames == 0) return SVS_OK;
    if (d_stego_out == nullptr) return fail(SVS_ERR_POINTER, "stego_out is NULL");
    if (stego_channels != 1 && stego_channels != 3) return fail(SVS_ERR_SHAPE, "stego_channels must be 1 or 3");
    if (stego_row_stride < (long long)width * stego_channels) return fail(SVS_ERR_STRIDE, "stego_row_stride smaller than a row");
    if (n_frames > 1 && stego_frame_stride < (long long)(height - 1) * stego_row_stride + (long long)width * stego_channels)
        return fail(SVS_ERR_STRIDE, "stego_frame_stride smaller than a frame");
    if (!aligned_to(d_stego_out, stego_frame_stride, stego_row_stride, 8))
        return fail(SVS_ERR_ALIGNMENT, "stego_out, stego_frame_stride and stego_row_stride must be multiples of 8");
    if (d_gray_out != nullptr && !aligned_to(d_gray_out, 0, 0, 8)) return fail(SVS_ERR_ALIGNMENT, "gray_out must be 8-byte aligned");
    if (payload_total_bits < 0) payload_total_bits = 0;
    if (payload_bit_offset < 0) return fail(SVS_ERR_SHAPE, "payload_bit_offset < 0");
    if (payload_total_bits > 0) {
        if (d_payload == nullptr) return fail(SVS_ERR_POINTER, "payload is NULL");
        if (!aligned_to(d_payload, 0, 0, 4)) return fail(SVS_ERR_ALIGNMENT, "payload must be 4-byte aligned");
    }
    EmbedArgs a;
    a.g = make_geometry(d_frames, height, width, frame_stride, row_stride, delta, num_ac);
    a.active = a.g.n > 0 && delta > 0;
    a.cap = a.active ? (long long)a.g.bpf * a.g.n : 0;
    a.payload = reinterpret_cast<const uint32_t*>(d_payload);
    a.payload_bit_offset = payload_bit_offset;
    a.payload_total_bits = payload_total_bits;
    a.payload_last_word = payload_total_bits > 0 ? (payload_bit_offset + payload_total_bits - 1) >> 5 : -1;
    a.stego = d_stego_out;
    a.stego_frame_stride = stego_frame_stride;
    a.stego_row_stride = stego_row_stride;
    a.gray = d_gray_out;
    a.bits_embedded = d_bits_embedded_out;
    a.sse = d_sse_out;
    const long long grid = n_frames * a.g.tiles_per_frame;
    if (grid > 0x7fffffffLL) return fail(SVS_ERR_SHAPE, "batch too large for one launch; split it");
    // float(q*delta) in float32 is a single rounding of the exact product whenever delta is
    // itself a float32 (|q| < 2^24 is guaranteed by delta >= 2^-10); otherwise go through double.
    const bool f64 = (double)(float)delta != delta;
    const bool al = aligned_to(d_frames, frame_stride, row_stride, 8);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // Frames the payload fills completely go to the packed-FP32 kernels; the frame in which the
    // payload ends and everything after it (and every special case) to the scalar kernel.
    const svs::FastQuant fq = make_fast_quant(delta);
    const bool want_side = d_gray_out != nullptr || d_sse_out != nullptr;
    const int fam = family();
    if (fam != 1 && al && !f64 && a.active && fq.embed_ok && n_frames * ((a.g.bpf + 31) / 32) < 0x7fffffffLL) {
        long long full = payload_total_bits / a.cap;
        if (full > n_frames) full = n_frames;
        bool done = false;
        const bool nfull = a.g.n == SVS_MAX_AC;
        if (full > 0 && (fam == 0 || fam == 5) && blk_addressable(full, height, frame_stride, row_stride) &&
            blk_addressable(full, height, stego_frame_stride, stego_row_stride) && a.payload_last_word < 0xfffffff0LL) {
            // block kernels: gray / SSE come from the bytes the thread already holds (SIDE instantiation)
            blk::BlkEmbedArgs ba;
            ba.g = make_blk_geometry(a.g, full);
            ba.q = fq;
            ba.payload = a.payload;
            ba.payload_bit_offset = payload_bit_offset;
            ba.payload_last_word = a.payload_last_word;
            ba.cap = a.cap;
            ba.stego = d_stego_out;
            ba.stego_frame_stride = stego_frame_stride;
            ba.stego_row_stride = stego_row_stride;
            ba.stego_frame_stride32 = (uint32_t)stego_frame_stride;
            ba.stego_row_stride32 = (uint32_t)stego_row_stride;
            ba.bits_embedded = d_bits_embedded_out;
            ba.gray = d_gray_out;
            ba.sse = d_sse_out;
            ba.gray_frame_stride = (long long)height * width;
            ba.W = width;
            cudaError_t e;
            if (channels == 3) e = stego_channels == 1 ? launch_blk_embed<3, 1>(ba, nfull, want_side, st) : launch_blk_embed<3, 3>(ba, nfull, want_side, st);
            else               e = stego_channels == 1 ? launch_blk_embed<1, 1>(ba, nfull, want_side, st) : launch_blk_embed<1, 3>(ba, nfull, want_side, st);
            g_launches.fetch_add(1);
            if (e != cudaSuccess) return cuda_fail(e, "svs_embed_frames launch (block)");
            done = true;
        }
#if defined(SVS_WITH_VARIANTS)
        // round-1 families: their optional gray / SSE outputs come from a separate streaming kernel
        // when its 16-pixel accesses apply; otherwise such a call goes to the scalar kernel altogether
        const bool side_ok = (width % 16) == 0 && aligned_to(d_frames, frame_stride, row_stride, 16) && stego_channels == 1 &&
                             aligned_to(d_stego_out, stego_frame_stride, stego_row_stride, 16) &&
                             (d_gray_out == nullptr || aligned_to(d_gray_out, 0, 0, 16)) && n_frames <= 65535;
        if (full > 0 && !done && (!want_side || side_ok)) {
            fast::FastEmbedArgs fa;
            fa.g = make_fast_geometry(a.g, full);
            fa.q = fq;
            fa.payload = a.payload;
            fa.payload_bit_offset = payload_bit_offset;
            fa.payload_last_word = a.payload_last_word;
            fa.cap = a.cap;
            fa.stego = d_stego_out;
            fa.stego_frame_stride = stego_frame_stride;
            fa.stego_row_stride = stego_row_stride;
            fa.bits_embedded = d_bits_embedded_out;
            const bool use_row = fam == 4 && row_ok(a.g, d_frames, frame_stride, row_stride) &&
                                 aligned_to(d_stego_out, stego_frame_stride, stego_row_stride, 16) &&
                                 a.payload_last_word < 0x7fffffffLL;
            if (use_row) {
                const unsigned rgrid = row_grid(fa.g.total_groups);
                row::RowEmbedArgs ra;
                ra.e = fa;
                ra.wrap_src = 8 * row_stride - (long long)a.g.bw * 8 * channels;
                ra.wrap_dst = 8 * stego_row_stride - (long long)a.g.bw * 8 * stego_channels;
#define SVS_LAUNCH_ROW_EMBED(CH, OC)                                                                     \
    do {                                                                                                 \
        if (nfull) row::embed_row_kernel<CH, OC, true><<<rgrid, row::kRowThreads, 0, st>>>(ra);          \
        else row::embed_row_kernel<CH, OC, false><<<rgrid, row::kRowThreads, 0, st>>>(ra);               \
    } while (0)
                if (channels == 3) { if (stego_channels == 1) SVS_LAUNCH_ROW_EMBED(3, 1); else SVS_LAUNCH_ROW_EMBED(3, 3); }
                else               { if (stego_channels == 1) SVS_LAUNCH_ROW_EMBED(1, 1); else SVS_LAUNCH_ROW_EMBED(1, 3); }
#undef SVS_LAUNCH_ROW_EMBED
            } else if (fam != 3) {
                const unsigned fgrid = fast_grid(fa.g.total_groups);
#define SVS_LAUNCH_EMBED(CH, OC)                                                                         \
    do {                                                                                                 \
        if (nfull) fast::embed_fast_kernel<CH, OC, true><<<fgrid, fast::kFastThreads, 0, st>>>(fa);      \
        else fast::embed_fast_kernel<CH, OC, false><<<fgrid, fast::kFastThreads, 0, st>>>(fa);           \
    } while (0)
                if (channels == 3) { if (stego_channels == 1) SVS_LAUNCH_EMBED(3, 1); else SVS_LAUNCH_EMBED(3, 3); }
                else               { if (stego_channels == 1) SVS_LAUNCH_EMBED(1, 1); else SVS_LAUNCH_EMBED(1, 3); }
#undef SVS_LAUNCH_EMBED
            } else {
                cudaError_t fe;
                if (channels == 3) {
                    if (stego_channels == 1) fe = nfull ? launch_tile_embed<3, 1, true>(fa, st) : launch_tile_embed<3, 1, false>(fa, st);
                    else                     fe = nfull ? launch_tile_embed<3, 3, true>(fa, st) : launch_tile_embed<3, 3, false>(fa, st);
                } else {
                    if (stego_channels == 1) fe = nfull ? launch_tile_embed<1, 1, true>(fa, st) : launch_tile_embed<1, 1, false>(fa, st);
                    else                     fe = nfull ? launch_tile_embed<1, 3, true>(fa, st) : launch_tile_embed<1, 3, false>(fa, st);
                }
                if (fe != cudaSuccess) return cuda_fail(fe, "svs_embed_frames launch (tile)");
            }
            g_launches.fetch_add(1);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) return cuda_fail(e, "svs_embed_frames launch (packed)");
            if (want_side) {
                SideArgs sa;
                sa.frames = d_frames;
                sa.frame_stride = frame_stride;
                sa.row_stride = row_stride;
                sa.stego = d_stego_out;
                sa.stego_frame_stride = stego_frame_stride;
                sa.stego_row_stride = stego_row_stride;
                sa.gray = d_gray_out;
                sa.sse = d_sse_out;
                sa.H = height;
                sa.W16 = width / 16;
                const int per_frame = height * (width / 16);
                int gx = (per_frame + 255) / 256;
                if (gx > 64) gx = 64;
                if (per_frame / (gx * 256) > 4000) gx = per_frame / (256 * 4000) + 1;   // 32-bit per-thread sums: <= 4000 chunks each
                const dim3 grid((unsigned)gx, (unsigned)full);
                if (channels == 3) side_outputs_kernel<3><<<grid, 256, 0, st>>>(sa);
                else side_outputs_kernel<1><<<grid, 256, 0, st>>>(sa);
                g_launches.fetch_add(1);
                e = cudaGetLastError();
                if (e != cudaSuccess) return cuda_fail(e, "svs_embed_frames launch (gray / SSE)");
            }
            done = true;
        }
#endif
        if (done) {
            if (full == n_frames) return SVS_OK;
            // the remaining frames: shift every per-frame quantity by `full`
            a.g.frames += full * frame_stride;
            a.payload_bit_offset += full * a.cap;
            a.payload_total_bits -= full * a.cap;
            a.stego += full * stego_frame_stride;
            if (a.bits_embedded) a.bits_embedded += full;
            if (a.gray) a.gray += full * (long long)height * width;
            if (a.sse) a.sse += full;
            n_frames -= full;
        }
    }
    const unsigned gr = (unsigned)(n_frames * a.g.tiles_per_frame);
    if (channels == 3) {
        if (stego_channels == 1) { if (al) launch_embed<3, 1, true>(a, f64, gr, st); else launch_embed<3, 1, false>(a, f64, gr, st); }
        else                     { if (al) launch_embed<3, 3, true>(a, f64, gr, st); else launch_embed<3, 3, false>(a, f64, gr, st); }
    } else {
        if (stego_channels == 1) { if (al) launch_embed<1, 1, true>(a, f64, gr, st); else launch_embed<1, 1, false>(a, f64, gr, st); }
        else                     { if (al) launch_embed<1, 3, true>(a, f64, gr, st); else launch_embed<1, 3, false>(a, f64, gr, st); }
    }
    g_launches.fetch_add(1);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "svs_embed_frames launch");
    return SVS_OK;
}

// ------------------------------------------------------------------------------------------
// Host-buffer entry points: chunked, three streams deep (H2D of chunk i+1 | kernel of chunk i |
// D2H of chunk i-1).  Work inside one slot is ordered by its stream, so device staging buffers
// are reused without host-side waits; the only host synchronisation is at the very end.
// ------------------------------------------------------------------------------------------
struct svs_slot {
    cudaStream_t stream;
    uint8_t* buf[4];            // 0: frames in, 1: stego / bits out, 2: gray out, 3: counters
    size_t cap[4];
};

struct svs_ctx {
    uint32_t magic;
    int device;
    long long budget;           // staging bytes per slot
    svs_slot slot[3];
    uint8_t* payload;
    size_t payload_cap;
    cudaEvent_t payload_ready;
};

namespace {
constexpr uint32_t kMagic = 0x53565342u;   // "SVSB"

int ensure(uint8_t*& p, size_t& cap, size_t need)
{
    if (need <= cap) return SVS_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    need = (need + 0xfffff) & ~(size_t)0xfffff;
    cudaError_t e = cudaMalloc(&p, need);
    if (e != cudaSuccess) return cuda_fail(e, "svs_ctx staging cudaMalloc");
    cap = need;
    return SVS_OK;
}

// host (strided) -> device (compact) copy of `nf` frames
cudaError_t copy_frames_in(uint8_t* dst, const uint8_t* src, long long nf, int H, long long row_bytes,
                           long long frame_stride, long long row_stride, cudaStream_t st)
{
    if (row_stride == row_bytes && frame_stride == row_bytes * H)
        return cudaMemcpyAsync(dst, src, (size_t)(nf * H * row_bytes), cudaMemcpyHostToDevice, st);
    if (frame_stride == row_stride * H)
        return cudaMemcpy2DAsync(dst, (size_t)row_bytes, src, (size_t)row_stride, (size_t)row_bytes, (size_t)(nf * H),
                                 cudaMemcpyHostToDevice, st);
    for (long long f = 0; f < nf; ++f) {
        cudaError_t e = cudaMemcpy2DAsync(dst + f * H * row_bytes, (size_t)row_bytes, src + f * frame_stride,
                                          (size_t)row_stride, (size_t)row_bytes, (size_t)H, cudaMemcpyHostToDevice, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int sync_all(svs_ctx* c)
{
    for (auto& s : c->slot) {
        cudaError_t e = cudaStreamSynchronize(s.stream);
        if (e != cudaSuccess) return cuda_fail(e, "svs host path synchronise");
    }
    return SVS_OK;
}

// Error exit of a chunk loop: copies of earlier chunks may still be reading / writing the
// caller's host buffers, so the streams are drained (result ignored, the message of `rc` is kept)
// before the caller gets its buffers back.
int drain_and_return(svs_ctx* c, int rc)
{
    char keep[sizeof g_err];
    memcpy(keep, g_err, sizeof keep);
    for (auto& s : c->slot) cudaStreamSynchronize(s.stream);
    memcpy(g_err, keep, sizeof keep);
    return rc;
}
}  // namespace

int svs_ctx_create(int device, int64_t staging_bytes_hint, svs_ctx** out)
{
    g_err[0] = 0;
    if (out == nullptr) return fail(SVS_ERR_POINTER, "out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount (no CUDA device: this library has no CPU fallback)");
    if (device < 0 || device >= count) return fail(SVS_ERR_CONTEXT, "device %d out of range (%d visible)", device, count);
    DeviceGuard guard(device);
    svs_ctx* c = new svs_ctx();
    memset(c, 0, sizeof *c);
    c->magic = kMagic;
    c->device = device;
    // default 64 MB per slot: small chunks keep the fill / drain bubbles of the 3-slot pipeline short
    // (measured through bench.py's e2e leg: 192 MB total 6,317 frames/s, 3 GB total 6,045 frames/s)
    c->budget = staging_bytes_hint > 0 ? staging_bytes_hint / 3 : (64ll << 20);
    for (auto& s : c->slot) {
        e = cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaStreamCreate"); }
    }
    e = cudaEventCreateWithFlags(&c->payload_ready, cudaEventDisableTiming);
    if (e != cudaSuccess) { delete c; return cuda_fail(e, "cudaEventCreate"); }
    *out = c;
    return SVS_OK;
}

int svs_ctx_destroy(svs_ctx* c)
{
    g_err[0] = 0;
    if (c == nullptr) return SVS_OK;
    if (c->magic != kMagic) return fail(SVS_ERR_CONTEXT, "not a live svs_ctx");
    DeviceGuard guard(c->device);
    for (auto& s : c->slot) {
        cudaStreamSynchronize(s.stream);
        for (auto& b : s.buf) if (b) cudaFree(b);
        cudaStreamDestroy(s.stream);
    }
    if (c->payload) cudaFree(c->payload);
    cudaEventDestroy(c->payload_ready);
    c->magic = 0;
    delete c;
    return SVS_OK;
}

int svs_extract_frames_host(svs_ctx* c, const uint8_t* h_frames, int channels, int64_t n_frames,
                            int height, int width, int64_t frame_stride, int64_t row_stride,
                            double delta, int num_ac, uint8_t* h_bits_out, int64_t bits_frame_stride)
{
    g_err[0] = 0;
    if (c == nullptr || c->magic != kMagic) return fail(SVS_ERR_CONTEXT, "not a live svs_ctx");
    if (int rc = check_geometry(h_frames, channels, n_frames, height, width, frame_stride, row_stride, delta)) return rc;
    const long long cap = svs_capacity_bits(height, width, num_ac);
    if (cap == 0 || n_frames == 0) return SVS_OK;
    if (h_bits_out == nullptr) return fail(SVS_ERR_POINTER, "bits_out is NULL");
    const long long frame_bytes = (cap + 7) / 8;
    if (bits_frame_stride < frame_bytes) return fail(SVS_ERR_STRIDE, "bits_frame_stride < ceil(cap/8)");
    DeviceGuard guard(c->device);
    const long long row_bytes = (long long)width * channels, in_bytes = row_bytes * height;
    const long long dev_bits_stride = svs_bits_row_bytes(height, width, num_ac);
    long long chunk = c->budget / (in_bytes + dev_bits_stride);
    chunk = chunk < 1 ? 1 : (chunk > n_frames ? n_frames : chunk);
    int k = 0;
    for (long long f0 = 0; f0 < n_frames; f0 += chunk, ++k) {
        svs_slot& s = c->slot[k % 3];
        const long long nf = n_frames - f0 < chunk ? n_frames - f0 : chunk;
        if (int rc = ensure(s.buf[0], s.cap[0], (size_t)(nf * in_bytes))) return drain_and_return(c, rc);
        if (int rc = ensure(s.buf[1], s.cap[1], (size_t)(nf * dev_bits_stride))) return drain_and_return(c, rc);
        cudaError_t e = copy_frames_in(s.buf[0], h_frames + f0 * frame_stride, nf, height, row_bytes, frame_stride, row_stride, s.stream);
        if (e != cudaSuccess) return drain_and_return(c, cuda_fail(e, "host->device frame copy"));
        if (int rc = svs_extract_frames(s.buf[0], channels, nf, height, width, in_bytes, row_bytes, delta, num_ac,
                                        s.buf[1], dev_bits_stride, s.stream)) return drain_and_return(c, rc);
        e = cudaMemcpy2DAsync(h_bits_out + f0 * bits_frame_stride, (size_t)bits_frame_stride, s.buf[1], (size_t)dev_bits_stride,
                              (size_t)frame_bytes, (size_t)nf, cudaMemcpyDeviceToHost, s.stream);
        if (e != cudaSuccess) return drain_and_return(c, cuda_fail(e, "device->host bits copy"));
    }
    return sync_all(c);
}

int svs_embed_frames_host(svs_ctx* c, const uint8_t* h_frames, int channels, int64_t n_frames,
                          int height, int width, int64_t frame_stride, int64_t row_stride,
                          const uint8_t* h_payload, int64_t payload_bit_offset, int64_t payload_total_bits,
                          double delta, int num_ac, uint8_t* h_stego_out, int stego_channels,
                          uint8_t* h_gray_out, int64_t* h_bits_embedded_out, unsigned long long* h_sse_out)
{
    g_err[0] = 0;
    if (c == nullptr || c->magic != kMagic) return fail(SVS_ERR_CONTEXT, "not a live svs_ctx");
    if (int rc = check_geometry(h_frames, channels, n_frames, height, width, frame_stride, row_stride, delta)) return rc;
    if (n_frames == 0) return SVS_OK;
    if (h_stego_out == nullptr) return fail(SVS_ERR_POINTER, "stego_out is NULL");
    if (stego_channels != 1 && stego_channels != 3) return fail(SVS_ERR_SHAPE, "stego_channels must be 1 or 3");
    if (payload_total_bits < 0) payload_total_bits = 0;
    if (payload_bit_offset < 0) return fail(SVS_ERR_SHAPE, "payload_bit_offset < 0");
    if (payload_total_bits > 0 && h_payload == nullptr) return fail(SVS_ERR_POINTER, "payload is NULL");
    DeviceGuard guard(c->device);
    const int n = num_ac < 0 ? 0 : (num_ac > SVS_MAX_AC ? SVS_MAX_AC : num_ac);
    const bool active = n > 0 && delta > 0;
    const long long cap = active ? svs_capacity_bits(height, width, n) : 0;
    const long long row_bytes = (long long)width * channels, in_bytes = row_bytes * height;
    const long long px = (long long)height * width, out_bytes = px * stego_channels;

    // payload: only the bytes this batch can consume, 4-byte aligned start, on slot 0's stream
    long long dev_bit_offset = 0;
    if (payload_total_bits > 0) {
        long long usable = payload_total_bits;
        if (active && usable > n_frames * cap) usable = n_frames * cap;
        if (!active) usable = 1;                                  // only "non-empty" matters
        const long long byte0 = (payload_bit_offset >> 5) << 2;
        const long long byte1 = (payload_bit_offset + usable + 7) >> 3;
        dev_bit_offset = payload_bit_offset - 8 * byte0;
        // the kernels read whole 32-bit words, up to two words past the one that holds a block's
        // first bit (masked by payload_last_word, which svs_embed_frames derives from the bit count):
        // 12 zeroed bytes of slack keep every such read inside the allocation
        if (int rc = ensure(c->payload, c->payload_cap, (size_t)(byte1 - byte0 + 12))) return rc;
        cudaError_t e = cudaMemsetAsync(c->payload + (byte1 - byte0), 0, 12, c->slot[0].stream);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(c->payload, h_payload + byte0, (size_t)(byte1 - byte0), cudaMemcpyHostToDevice, c->slot[0].stream);
        if (e != cudaSuccess) return drain_and_return(c, cuda_fail(e, "host->device payload copy"));
        cudaEventRecord(c->payload_ready, c->slot[0].stream);
        cudaStreamWaitEvent(c->slot[1].stream, c->payload_ready, 0);
        cudaStreamWaitEvent(c->slot[2].stream, c->payload_ready, 0);
    }
    long long chunk = c->budget / (in_bytes + out_bytes + (h_gray_out ? px : 0) + 16);
    chunk = chunk < 1 ? 1 : (chunk > n_frames ? n_frames : chunk);
    int k = 0;
    for (long long f0 = 0; f0 < n_frames; f0 += chunk, ++k) {
        svs_slot& s = c->slot[k % 3];
        const long long nf = n_frames - f0 < chunk ? n_frames - f0 : chunk;
        if (int rc = ensure(s.buf[0], s.cap[0], (size_t)(nf * in_bytes))) return drain_and_return(c, rc);
        if (int rc = ensure(s.buf[1], s.cap[1], (size_t)(nf * out_bytes))) return drain_and_return(c, rc);
        if (h_gray_out) if (int rc = ensure(s.buf[2], s.cap[2], (size_t)(nf * px))) return drain_and_return(c, rc);
        if (int rc = ensure(s.buf[3], s.cap[3], (size_t)(nf * 16))) return drain_and_return(c, rc);
        int64_t* d_nbits = reinterpret_cast<int64_t*>(s.buf[3]);
        unsigned long long* d_sse = reinterpret_cast<unsigned long long*>(s.buf[3] + nf * 8);
        cudaError_t e = copy_frames_in(s.buf[0], h_frames + f0 * frame_stride, nf, height, row_bytes, frame_stride, row_stride, s.stream);
        if (e != cudaSuccess) return drain_and_return(c, cuda_fail(e, "host->device frame copy"));
        if (h_sse_out) cudaMemsetAsync(d_sse, 0, (size_t)(nf * 8), s.stream);
        const long long first = active ? f0 * cap : 0;            // payload bits consumed by earlier chunks
        if (int rc = svs_embed_frames(s.buf[0], channels, nf, height, width, in_bytes, row_bytes,
                                      c->payload, dev_bit_offset + first, payload_total_bits - first, delta, num_ac,
                                      s.buf[1], stego_channels, out_bytes, (long long)width * stego_channels,
                                      h_gray_out ? s.buf[2] : nullptr, h_bits_embedded_out ? d_nbits : nullptr,
                                      h_sse_out ? d_sse : nullptr, s.stream)) return drain_and_return(c, rc);
        e = cudaMemcpyAsync(h_stego_out + f0 * out_bytes, s.buf[1], (size_t)(nf * out_bytes), cudaMemcpyDeviceToHost, s.stream);
        if (e == cudaSuccess && h_gray_out) e = cudaMemcpyAsync(h_gray_out + f0 * px, s.buf[2], (size_t)(nf * px), cudaMemcpyDeviceToHost, s.stream);
        if (e == cudaSuccess && h_bits_embedded_out) e = cudaMemcpyAsync(h_bits_embedded_out + f0, d_nbits, (size_t)(nf * 8), cudaMemcpyDeviceToHost, s.stream);
        if (e == cudaSuccess && h_sse_out) e = cudaMemcpyAsync(h_sse_out + f0, d_sse, (size_t)(nf * 8), cudaMemcpyDeviceToHost, s.stream);
        if (e != cudaSuccess) return drain_and_return(c, cuda_fail(e, "device->host result copy"));
    }
    return sync_all(c);
}

}  // extern "C"
