// Op-exact float32 8-point DCT-II / DCT-III, shared by the sm_100a kernels (svs_b200.cu) and by
// a host-only g++ build used in the CPU tests (tests/host_math/host_math.cpp), so that the exact
// operation sequence the kernels execute can be checked against the oracle without a GPU.
//
// What is reproduced: scipy.fftpack.dct / idct (type 2, norm='ortho') on
// float32 as called by the reference at config_and_setup.py:135 and :168 - pocketfft's N=8 real
// FFT plan (radix-2 then radix-4) inside its DCT pre/post-processing.  Every add / sub / mulc
// below is ONE IEEE binary32 operation, round-to-nearest-even, never contracted into an FMA.
// The transforms are written once over an arithmetic policy `A`:
//   ScalarOps  float; device: __fadd_rn / __fsub_rn / __fmul_rn (documented as never fused),
//              host: plain operators compiled with -ffp-contract=off;
//   PackedOps  (svs_b200.cu) two blocks per thread in one 64-bit register pair, Blackwell
//              add/sub/fma .f32x2 (FADD2 / FFMA2), products written as fma(a, c, -0.0) with an
//              opaque -0.0 because ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2
//              even with explicit .rn and --fmad=false.
// The reference's multiplications by exact powers of two (2, 0.25, 0.5) commute with rounding in
// this value range (no underflow: |values| are 0 or > 2^-100), so they are folded into the
// constants; that is the only algebra applied.  54 operations per 8-point transform.
#pragma once

#if defined(__CUDACC__)
#define SVS_HD __host__ __device__ __forceinline__
#define SVS_HDM __host__ __device__ __forceinline__
#else
#define SVS_HD static inline
#define SVS_HDM inline
#endif

namespace svs {

// fl32(cos(k*pi/16)) / 4, k = 1..7, written as exact hex floats
#define SVS_Q1 0x1.f6297cp-3f   /* 0x3E7B14BE */
#define SVS_Q2 0x1.d906bcp-3f   /* 0x3E6C835E */
#define SVS_Q3 0x1.a9b662p-3f   /* 0x3E54DB31 */
#define SVS_Q5 0x1.1c73b4p-3f   /* 0x3E0E39DA */
#define SVS_Q6 0x1.87de2ap-4f   /* 0x3DC3EF15 */
#define SVS_Q7 0x1.8f8b84p-5f   /* 0x3D47C5C2 */
#define SVS_W  0x1.6a09e6p-1f   /* 0x3F3504F3 = fl32(sqrt(1/2)) */
#define SVS_HW 0x1.6a09e6p-2f   /* 0x3EB504F3 = W/2 = fl32(sqrt 2)/4 */

struct ScalarOps {
    typedef float T;
#if defined(__CUDA_ARCH__)
    SVS_HDM T add(T a, T b) const { return __fadd_rn(a, b); }
    SVS_HDM T sub(T a, T b) const { return __fsub_rn(a, b); }
    SVS_HDM T mulc(T a, float c) const { return __fmul_rn(a, c); }
    SVS_HDM T cst(float c) const { return c; }
    // sum = a + b, diff = a - b, emitted back to back: the second instruction finds both operands
    // in the operand-reuse cache instead of the register file (profiles/microbench/coissue.cu:
    // register-file read bandwidth, not the FP32 pipe, is what the packed kernels run out of)
    SVS_HDM void bfly(T a, T b, T& sum, T& diff) const
    {
        asm("add.rn.f32 %0, %2, %3;\n\tsub.rn.f32 %1, %2, %3;" : "=&f"(sum), "=&f"(diff) : "f"(a), "f"(b));
    }
#else
    SVS_HDM T add(T a, T b) const { return a + b; }
    SVS_HDM T sub(T a, T b) const { return a - b; }
    SVS_HDM T mulc(T a, float c) const { return a * c; }
    SVS_HDM T cst(float c) const { return c; }
    SVS_HDM void bfly(T a, T b, T& sum, T& diff) const { sum = a + b; diff = a - b; }
#endif
};

// The first butterfly stage of either transform reads every input exactly once (8 operations);
// everything after it only combines the 8 stage-1 values.  They are split so that the packed
// one-block-per-thread kernels (svs_block.cuh) can run stage 1 on SCALAR registers - which
// regroups the 8x8 block from column pairs to row pairs for free - and the tail packed.
// dct8_fwd / dct8_inv below are head + tail, so every user executes the same 54 operations.
template <class T>
struct Stage1 {
    T v[8];
};

// forward head: s07, a1, a2, a3, a4, a5, a6, d07   (A.1 steps 1-3, the input butterflies)
template <class A>
SVS_HD Stage1<typename A::T> dct8_fwd_head(const A& o, typename A::T x0, typename A::T x1, typename A::T x2,
                                           typename A::T x3, typename A::T x4, typename A::T x5,
                                           typename A::T x6, typename A::T x7)
{
    Stage1<typename A::T> h;
    o.bfly(x0, x7, h.v[0], h.v[7]);
    o.bfly(x2, x1, h.v[1], h.v[2]);              // x1 + x2 (commutative: same bits), x2 - x1
    o.bfly(x4, x3, h.v[3], h.v[4]);
    o.bfly(x6, x5, h.v[5], h.v[6]);
    return h;
}

// forward tail: the remaining 46 operations; X[0..7] out.
//
// BIASED (axis-0 pass of the packed kernels, svs_block.cuh): the 8 inputs of the head were not
// the pixels p but 2^23 + 256 p - a byte dropped into the mantissa of 2^23 by one PRMT, no
// subtraction.  Every operation up to and including e0..e7's integer part is then still exact:
// differences cancel the 2^23, sums double it (2^24 + 256 (p + p'), a multiple of the ulp), and
// the only stage-3 value that still carries it is e0 = 2^26 + 256 (sum of the 8 pixels), one
// subtraction.  All values are 256 x what the reference holds at that point; scaling by a power
// of two commutes with every rounding (no underflow: non-zero magnitudes stay > 2^-100), so the
// 2^-8 is folded into the constants of the final products.  Net: 7 operations fewer per
// transform than converting each byte to a float first.
template <bool BIASED, class A>
SVS_HD void dct8_fwd_tail_impl(const A& o, const Stage1<typename A::T>& h, typename A::T (&X)[8])
{
    typedef typename A::T T;
    constexpr float S = BIASED ? 0x1p-8f : 1.0f;
    const T s07 = h.v[0], a1 = h.v[1], a2 = h.v[2], a3 = h.v[3], a4 = h.v[4], a5 = h.v[5], a6 = h.v[6], d07 = h.v[7];
    T h1, tr, ti, h2, h6, h5, p0, m0, p1, m1, e0, e4, e6, e2, e1, e5, e7, e3;
    o.bfly(a1, a5, h1, tr);
    o.bfly(a2, a6, ti, h2);
    const T wti = o.mulc(ti, SVS_W), wtr = o.mulc(tr, SVS_W);
    o.bfly(wtr, wti, h6, h5);                    // wti + wtr, wtr - wti
    o.bfly(s07, a3, p0, m0);
    o.bfly(d07, a4, m1, p1);                     // d07 + a4, d07 - a4
    o.bfly(p0, h1, e0, e4);
    if constexpr (BIASED) e0 = o.sub(e0, o.cst(67108864.0f));          // 8 x 2^23, exact
    o.bfly(m0, h2, e6, e2);
    o.bfly(p1, h5, e1, e5);
    o.bfly(m1, h6, e7, e3);
    T t1, t2;
    t1 = o.add(o.mulc(e7, SVS_Q1 * S), o.mulc(e1, SVS_Q7 * S));
    t2 = o.sub(o.mulc(e1, SVS_Q1 * S), o.mulc(e7, SVS_Q7 * S));
    o.bfly(t1, t2, X[1], X[7]);
    t1 = o.add(o.mulc(e6, SVS_Q2 * S), o.mulc(e2, SVS_Q6 * S));
    t2 = o.sub(o.mulc(e2, SVS_Q2 * S), o.mulc(e6, SVS_Q6 * S));
    o.bfly(t1, t2, X[2], X[6]);
    t1 = o.add(o.mulc(e5, SVS_Q3 * S), o.mulc(e3, SVS_Q5 * S));
    t2 = o.sub(o.mulc(e3, SVS_Q3 * S), o.mulc(e5, SVS_Q5 * S));
    o.bfly(t1, t2, X[3], X[5]);
    X[4] = o.mulc(e4, SVS_HW * S);
    X[0] = o.mulc(e0, SVS_HW * S);
}

template <class A>
SVS_HD void dct8_fwd_tail(const A& o, const Stage1<typename A::T>& h, typename A::T (&X)[8])
{
    dct8_fwd_tail_impl<false>(o, h, X);
}

// Forward DCT-II (norm='ortho'): x[0], x[S], ..., x[7S] in place.
template <int S, class A>
SVS_HD void dct8_fwd(const A& o, typename A::T* x)
{
    typename A::T X[8];
    dct8_fwd_tail(o, dct8_fwd_head(o, x[0], x[S], x[2 * S], x[3 * S], x[4 * S], x[5 * S], x[6 * S], x[7 * S]), X);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k * S] = X[k];
}

// inverse head: c0, c4, X1+X7, X1-X7, X2+X6, X2-X6, X3+X5, X3-X5   (A.2 steps 1-3, first half)
template <class A>
SVS_HD Stage1<typename A::T> dct8_inv_head(const A& o, typename A::T X0, typename A::T X1, typename A::T X2,
                                           typename A::T X3, typename A::T X4, typename A::T X5,
                                           typename A::T X6, typename A::T X7)
{
    Stage1<typename A::T> h;
    h.v[0] = o.mulc(X0, SVS_HW);
    h.v[1] = o.mulc(X4, SVS_HW);
    o.bfly(X1, X7, h.v[2], h.v[3]);
    o.bfly(X2, X6, h.v[4], h.v[5]);
    o.bfly(X3, X5, h.v[6], h.v[7]);
    return h;
}

// inverse tail: the remaining 46 operations; x[0..7] out
template <class A>
SVS_HD void dct8_inv_tail(const A& o, const Stage1<typename A::T>& h, typename A::T (&x)[8])
{
    typedef typename A::T T;
    const T c0 = h.v[0], c4 = h.v[1];
    T t1, t2;
    t1 = h.v[2]; t2 = h.v[3];
    const T c1 = o.add(o.mulc(t2, SVS_Q1), o.mulc(t1, SVS_Q7));
    const T c7 = o.sub(o.mulc(t1, SVS_Q1), o.mulc(t2, SVS_Q7));
    t1 = h.v[4]; t2 = h.v[5];
    const T c2 = o.add(o.mulc(t2, SVS_Q2), o.mulc(t1, SVS_Q6));
    const T c6 = o.sub(o.mulc(t1, SVS_Q2), o.mulc(t2, SVS_Q6));
    t1 = h.v[6]; t2 = h.v[7];
    const T c3 = o.add(o.mulc(t2, SVS_Q3), o.mulc(t1, SVS_Q5));
    const T c5 = o.sub(o.mulc(t1, SVS_Q3), o.mulc(t2, SVS_Q5));
    T r1, r2, h0, h1, h2, h3, h4, h5, h6, h7, tr, ti, o1, o2, o5, o6;
    o.bfly(c6, c2, r1, h2);
    o.bfly(c0, c4, r2, h1);
    o.bfly(r2, r1, h0, h3);
    o.bfly(c7, c3, r1, h6);
    o.bfly(c1, c5, r2, h5);
    o.bfly(r2, r1, h4, h7);
    const T wh5 = o.mulc(h5, SVS_W), wh6 = o.mulc(h6, SVS_W);
    o.bfly(wh6, wh5, tr, ti);                    // wh5 + wh6, wh6 - wh5
    o.bfly(h1, tr, o1, o5);
    o.bfly(ti, h2, o2, o6);
    o.bfly(h0, h4, x[0], x[7]);
    o.bfly(o1, o2, x[2], x[1]);                  // o2 + o1, o1 - o2
    o.bfly(h3, h7, x[3], x[4]);
    o.bfly(o5, o6, x[6], x[5]);                  // o6 + o5, o5 - o6
}

// Inverse (DCT-III, scipy idct type=2 norm='ortho'): x[0], x[S], ..., x[7S] in place.
template <int S, class A>
SVS_HD void dct8_inv(const A& o, typename A::T* x)
{
    typename A::T y[8];
    dct8_inv_tail(o, dct8_inv_head(o, x[0], x[S], x[2 * S], x[3 * S], x[4 * S], x[5 * S], x[6 * S], x[7 * S]), y);
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k * S] = y[k];
}

// 2-D: axis 0 (down the columns) first, then axis 1 (config_and_setup.py:135,168).  b[u*8+v].
template <class A>
SVS_HD void dct2_fwd(const A& o, typename A::T* b)
{
#pragma unroll
    for (int v = 0; v < 8; ++v) dct8_fwd<8>(o, b + v);
#pragma unroll
    for (int u = 0; u < 8; ++u) dct8_fwd<1>(o, b + 8 * u);
}

template <class A>
SVS_HD void dct2_inv(const A& o, typename A::T* b)
{
#pragma unroll
    for (int v = 0; v < 8; ++v) dct8_inv<8>(o, b + v);
#pragma unroll
    for (int u = 0; u < 8; ++u) dct8_inv<1>(o, b + 8 * u);
}

}  // namespace svs
