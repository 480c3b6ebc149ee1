// Dependent-issue latency microbenchmark for sm_100a: one warp, one dependency chain.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 2048
#define OPL(name, body) struct name { static constexpr const char* label = #name; \
    __device__ __forceinline__ static void run(unsigned& r, unsigned long long& q) { body } };
OPL(FADD,  asm volatile("add.rn.f32 %0, %0, 0f3F800001;" : "+r"(r));)
OPL(FFMA,  asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f3F000000;" : "+r"(r));)
OPL(FADD2, asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(q));)
OPL(FFMA2, asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(q));)
OPL(LOP3,  asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(r));)
OPL(PRMT,  asm volatile("prmt.b32 %0, %0, 0x4b000000, 0x7440;" : "+r"(r));)
OPL(SHF,   asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(r));)
OPL(IDP2A, asm volatile("dp2a.lo.u32.u32 %0, 0x4b230e97, %0, 0x4000;" : "+r"(r));)
OPL(F2IPU8, asm volatile("{.reg .u8 t; .reg .b32 u; cvt.rzi.u8.f32 t, %0; cvt.u32.u8 u, t; or.b32 %0, u, 0x3f800000;}" : "+r"(r));)
OPL(FADD2_then_LOP3, asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(q)); { unsigned lo = (unsigned)q; asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(lo)); q = (q & 0xffffffff00000000ull) | lo; })
OPL(LOP3_then_FADD2, { unsigned lo = (unsigned)q; asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(lo)); q = (q & 0xffffffff00000000ull) | lo; } asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(q));)

template <class Op>
__global__ void lat(unsigned long long* out, unsigned seed)
{
    unsigned r = seed + threadIdx.x;
    unsigned long long q = ((unsigned long long)(0x3f800000u + seed) << 32) | (0x3f900000u + threadIdx.x);
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) Op::run(r, q);
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = r ^ (unsigned)q ^ (unsigned)(q >> 32); }
}
__global__ void lds_lat(unsigned long long* out, unsigned seed)
{
    __shared__ unsigned long long buf[64];
    for (int i = threadIdx.x; i < 64; i += 32) buf[i] = (i * 7 + 3) % 64;   // pointer chase over 8-byte elements
    __syncwarp();
    unsigned long long idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; ++i) idx = buf[idx & 63];
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = idx + seed; }
}
template <class Op> void run(unsigned long long* d)
{
    lat<Op><<<1, 32>>>(d, 3); lat<Op><<<1, 32>>>(d, 5);
    unsigned long long h[2];
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-18s %6.2f clk per dependent step\n", Op::label, (double)h[0] / N);
}
int main()
{
    unsigned long long* d; cudaMalloc(&d, 64);
    run<FADD>(d); run<FFMA>(d); run<FADD2>(d); run<FFMA2>(d); run<LOP3>(d); run<PRMT>(d); run<SHF>(d); run<IDP2A>(d);
    run<F2IPU8>(d); run<FADD2_then_LOP3>(d); run<LOP3_then_FADD2>(d);
    lds_lat<<<1, 32>>>(d, 1); lds_lat<<<1, 32>>>(d, 2);
    unsigned long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%-18s %6.2f clk per dependent step\n", "LDS.64", (double)h[0] / N);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
