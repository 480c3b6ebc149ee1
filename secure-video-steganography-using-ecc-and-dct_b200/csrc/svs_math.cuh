// Op-exact float32 8-point DCT-II / DCT-III and the parity quantiser, shared by the sm_100a
// kernels (svs_kernels.cu) and by a host-only g++ build used in the CPU tests
// (tests/host_math/host_math.cpp) so that the exact operation sequence the kernels execute
// can be checked against the oracle without a GPU.
//
// What is reproduced: scipy.fftpack.dct / idct (type 2, norm='ortho') on
// float32 as called by the reference at config_and_setup.py:135 and :168 - pocketfft's N=8 real
// FFT plan (radix-2 then radix-4) inside its DCT pre/post-processing.  Every FA/FS/FM below is
// ONE IEEE binary32 operation, round-to-nearest-even, never contracted into an FMA:
//   device: __fadd_rn / __fsub_rn / __fmul_rn (documented as never fused);
//   host:   plain operators, compiled with -ffp-contract=off.
// The reference's multiplications by exact powers of two (2, 0.25, 0.5) commute with rounding in
// this value range (no underflow: |values| are 0 or > 2^-100), so they are folded into the
// constants; that is the only algebra applied.  54 operations per 8-point transform.
#pragma once

#if defined(__CUDACC__)
#define SVS_HD __host__ __device__ __forceinline__
#else
#define SVS_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define FA(a, b) __fadd_rn((a), (b))
#define FS(a, b) __fsub_rn((a), (b))
#define FM(a, b) __fmul_rn((a), (b))
#else
#define FA(a, b) ((a) + (b))
#define FS(a, b) ((a) - (b))
#define FM(a, b) ((a) * (b))
#endif

namespace svs {

// fl32(cos(k*pi/16)) * 0.25, k = 1..7 (exact scalings of 0x3F7B14BE ... 0x3E47C5C2)
#define SVS_Q1 0x1.f6297cp-3f   /* 0x3E7B14BE */
#define SVS_Q2 0x1.d906bcp-3f   /* 0x3E6C835E */
#define SVS_Q3 0x1.a9b662p-3f   /* 0x3E54DB31 */
#define SVS_Q4 0x1.6a09e6p-3f   /* 0x3E3504F3 */
#define SVS_Q5 0x1.1c73b4p-3f   /* 0x3E0E39DA */
#define SVS_Q6 0x1.87de2ap-4f   /* 0x3DC3EF15 */
#define SVS_Q7 0x1.8f8b84p-5f   /* 0x3D47C5C2 */
#define SVS_W  0x1.6a09e6p-1f   /* 0x3F3504F3 = fl32(sqrt(1/2)) */
#define SVS_HW 0x1.6a09e6p-2f   /* 0x3EB504F3 = W/2 = fl32(sqrt 2)/4 */

// Forward: x[0], x[S], ..., x[7S] in place.
template <int S>
SVS_HD void dct8_fwd(float* x)
{
    const float x0 = x[0], x1 = x[S], x2 = x[2 * S], x3 = x[3 * S];
    const float x4 = x[4 * S], x5 = x[5 * S], x6 = x[6 * S], x7 = x[7 * S];
    const float a1 = FA(x1, x2), a2 = FS(x2, x1);
    const float a3 = FA(x3, x4), a4 = FS(x4, x3);
    const float a5 = FA(x5, x6), a6 = FS(x6, x5);
    const float s07 = FA(x0, x7), d07 = FS(x0, x7);
    const float h1 = FA(a1, a5), tr = FS(a1, a5);
    const float ti = FA(a2, a6), h2 = FS(a2, a6);
    const float wti = FM(SVS_W, ti), wtr = FM(SVS_W, tr);
    const float h6 = FA(wti, wtr), h5 = FS(wtr, wti);
    const float p0 = FA(s07, a3), m0 = FS(s07, a3);
    const float p1 = FS(d07, a4), m1 = FA(d07, a4);
    const float e0 = FA(p0, h1), e4 = FS(p0, h1);
    const float e6 = FA(m0, h2), e2 = FS(m0, h2);
    const float e1 = FA(p1, h5), e5 = FS(p1, h5);
    const float e7 = FA(m1, h6), e3 = FS(m1, h6);
    float t1, t2;
    t1 = FA(FM(SVS_Q1, e7), FM(SVS_Q7, e1));
    t2 = FS(FM(SVS_Q1, e1), FM(SVS_Q7, e7));
    x[S] = FA(t1, t2);
    x[7 * S] = FS(t1, t2);
    t1 = FA(FM(SVS_Q2, e6), FM(SVS_Q6, e2));
    t2 = FS(FM(SVS_Q2, e2), FM(SVS_Q6, e6));
    x[2 * S] = FA(t1, t2);
    x[6 * S] = FS(t1, t2);
    t1 = FA(FM(SVS_Q3, e5), FM(SVS_Q5, e3));
    t2 = FS(FM(SVS_Q3, e3), FM(SVS_Q5, e5));
    x[3 * S] = FA(t1, t2);
    x[5 * S] = FS(t1, t2);
    x[4 * S] = FM(e4, SVS_HW);
    x[0] = FM(e0, SVS_HW);
}

// Inverse (DCT-III): x[0], x[S], ..., x[7S] in place.
template <int S>
SVS_HD void dct8_inv(float* x)
{
    const float X0 = x[0], X1 = x[S], X2 = x[2 * S], X3 = x[3 * S];
    const float X4 = x[4 * S], X5 = x[5 * S], X6 = x[6 * S], X7 = x[7 * S];
    const float c0 = FM(X0, SVS_HW);
    const float c4 = FM(X4, SVS_HW);
    float t1, t2;
    t1 = FA(X1, X7); t2 = FS(X1, X7);
    const float c1 = FA(FM(SVS_Q1, t2), FM(SVS_Q7, t1));
    const float c7 = FS(FM(SVS_Q1, t1), FM(SVS_Q7, t2));
    t1 = FA(X2, X6); t2 = FS(X2, X6);
    const float c2 = FA(FM(SVS_Q2, t2), FM(SVS_Q6, t1));
    const float c6 = FS(FM(SVS_Q2, t1), FM(SVS_Q6, t2));
    t1 = FA(X3, X5); t2 = FS(X3, X5);
    const float c3 = FA(FM(SVS_Q3, t2), FM(SVS_Q5, t1));
    const float c5 = FS(FM(SVS_Q3, t1), FM(SVS_Q5, t2));
    float r1, r2;
    r1 = FA(c6, c2); const float h2 = FS(c6, c2);
    r2 = FA(c0, c4); const float h1 = FS(c0, c4);
    const float h0 = FA(r2, r1), h3 = FS(r2, r1);
    r1 = FA(c7, c3); const float h6 = FS(c7, c3);
    r2 = FA(c1, c5); const float h5 = FS(c1, c5);
    const float h4 = FA(r2, r1), h7 = FS(r2, r1);
    const float wh5 = FM(SVS_W, h5), wh6 = FM(SVS_W, h6);
    const float tr = FA(wh5, wh6), ti = FS(wh6, wh5);
    const float o1 = FA(h1, tr), o5 = FS(h1, tr);
    const float o2 = FA(ti, h2), o6 = FS(ti, h2);
    x[0] = FA(h0, h4);
    x[7 * S] = FS(h0, h4);
    x[S] = FS(o1, o2);
    x[2 * S] = FA(o2, o1);
    x[3 * S] = FA(h3, h7);
    x[4 * S] = FS(h3, h7);
    x[5 * S] = FS(o5, o6);
    x[6 * S] = FA(o6, o5);
}

// 2-D: axis 0 (down the columns) first, then axis 1 (config_and_setup.py:135,168).  b[u*8+v].
SVS_HD void dct2_fwd(float* b)
{
#pragma unroll
    for (int v = 0; v < 8; ++v) dct8_fwd<8>(b + v);
#pragma unroll
    for (int u = 0; u < 8; ++u) dct8_fwd<1>(b + 8 * u);
}

SVS_HD void dct2_inv(float* b)
{
#pragma unroll
    for (int v = 0; v < 8; ++v) dct8_inv<8>(b + v);
#pragma unroll
    for (int u = 0; u < 8; ++u) dct8_inv<1>(b + 8 * u);
}

// OpenCV 4.13 BGR2GRAY (15-bit fixed point), config_and_setup.py:112.
SVS_HD unsigned gray_bgr(unsigned b, unsigned g, unsigned r)
{
    return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15;
}

}  // namespace svs
