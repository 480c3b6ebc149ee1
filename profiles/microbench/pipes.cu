// Pipe-throughput microbenchmark for sm_100a (B200): warp-instructions per clock per SM for the
// instruction classes the DCT-QIM kernels are built from, alone and in mixes.  One CTA of 512
// threads (4 warps per SM sub-partition) per SM, 8 independent dependency chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run: ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 512
#define CHAINS 8

template <class Op>
__global__ void __launch_bounds__(512, 1) bench(unsigned long long* cycles, unsigned* sink, unsigned seed)
{
    unsigned r[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
    unsigned long long q[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) q[i] = ((unsigned long long)(0x3f800000u + i) << 32) | (0x3f900000u + threadIdx.x);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) Op::run(r[i], q[i]);
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc ^= r[i] ^ (unsigned)q[i] ^ (unsigned)(q[i] >> 32);
    if (acc == 0x12345678u) sink[0] = acc;
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

#define OP(name, n, body) struct name { static constexpr int count = n; static constexpr const char* label = #name; \
    __device__ __forceinline__ static void run(unsigned& r, unsigned long long& q) { body } };

OP(FADD,   1, asm volatile("add.rn.f32 %0, %0, 0f3F800001;" : "+r"(r));)
OP(FMUL,   1, asm volatile("mul.rn.f32 %0, %0, 0f3F800001;" : "+r"(r));)
OP(FFMA,   1, asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f3F000000;" : "+r"(r));)
OP(FFMA_R, 1, asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+r"(r));)
OP(FADD_RZ,1, asm volatile("add.rz.f32 %0, %0, 0f3F800001;" : "+r"(r));)
OP(FADD2,  1, asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(q));)
OP(FMUL2,  1, asm volatile("mul.rn.f32x2 %0, %0, %0;" : "+l"(q));)
OP(FFMA2,  1, asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(q));)
OP(LOP3,   1, asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(r));)
OP(PRMT,   1, asm volatile("prmt.b32 %0, %0, 0x4b000000, 0x7440;" : "+r"(r));)
OP(SHF,    1, asm volatile("shf.l.wrap.b32 %0, %0, %0, 7;" : "+r"(r));)
OP(IADD3,  1, asm volatile("add.u32 %0, %0, 0x1234567;" : "+r"(r));)
OP(IMAD,   1, asm volatile("mad.lo.u32 %0, %0, 0x10dcd, 0x1234567;" : "+r"(r));)
OP(IDP2A,  1, asm volatile("dp2a.lo.u32.u32 %0, 0x4b230e97, %0, 0x4000;" : "+r"(r));)
OP(I2FP_U, 1, asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(r));)
OP(I2F_S,  1, asm volatile("cvt.rn.f32.s32 %0, %0;" : "+r"(r));)
OP(I2F_U8, 1, asm volatile("{.reg .b8 t<4>; mov.b32 {t0,t1,t2,t3}, %0; cvt.rn.f32.u8 %0, t2;}" : "+r"(r));)
OP(F2I_RZ, 1, asm volatile("cvt.rzi.u32.f32 %0, %0;" : "+r"(r));)
OP(F2I_RN, 1, asm volatile("cvt.rni.s32.f32 %0, %0;" : "+r"(r));)
OP(F2I_U8, 1, asm volatile("{.reg .u8 t; .reg .b32 u; cvt.rzi.u8.f32 t, %0; cvt.u32.u8 u, t; or.b32 %0, u, 0x3f800000;}" : "+r"(r));)
OP(FMNMX,  1, asm volatile("min.f32 %0, %0, 0f437F0000;" : "+r"(r));)
OP(SELP,   1, asm volatile("{.reg .pred p; setp.lt.u32 p, %0, 0x40000000; selp.b32 %0, 0x3f800001, %0, p;}" : "+r"(r));)
OP(MUFU_RCP,1, asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+r"(r));)
OP(FRND,   1, asm volatile("cvt.rni.f32.f32 %0, %0;" : "+r"(r));)
// mixes (2 instructions per run)
OP(FFMA2_LOP3, 2, asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(q)); asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(r));)
OP(FADD2_LOP3, 2, asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(q)); asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(r));)
OP(FADD_LOP3,  2, asm volatile("add.rn.f32 %0, %0, 0f3F800001;" : "+r"(r)); { unsigned lo = (unsigned)q; asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(lo)); q = (q & 0xffffffff00000000ull) | lo; })
OP(FADD2_FADD2_LOP3, 3, asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(q)); asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(r)); asm volatile("add.rn.f32x2 %0, %0, %0;" : "+l"(q));)
OP(FADD_FADD,  2, asm volatile("add.rn.f32 %0, %0, 0f3F800001;" : "+r"(r)); { unsigned lo = (unsigned)q; asm volatile("add.rn.f32 %0, %0, 0f3F800001;" : "+r"(lo)); q = (q & 0xffffffff00000000ull) | lo; })
OP(FADD_IDP,   2, asm volatile("add.rn.f32 %0, %0, 0f3F800001;" : "+r"(r)); { unsigned lo = (unsigned)q; asm volatile("dp2a.lo.u32.u32 %0, 0x4b230e97, %0, 0x4000;" : "+r"(lo)); q = (q & 0xffffffff00000000ull) | lo; })
OP(FADD_I2FP,  2, asm volatile("add.rn.f32 %0, %0, 0f3F800001;" : "+r"(r)); { unsigned lo = (unsigned)q; asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(lo)); q = (q & 0xffffffff00000000ull) | lo; })
OP(FADD_F2I,   2, asm volatile("add.rn.f32 %0, %0, 0f3F800001;" : "+r"(r)); { unsigned lo = (unsigned)q; asm volatile("cvt.rzi.u32.f32 %0, %0;" : "+r"(lo)); q = (q & 0xffffffff00000000ull) | lo; })
OP(LOP3_PRMT,  2, asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(r)); { unsigned lo = (unsigned)q; asm volatile("prmt.b32 %0, %0, 0x4b000000, 0x7440;" : "+r"(lo)); q = (q & 0xffffffff00000000ull) | lo; })
OP(LOP3_I2FP,  2, asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(r)); { unsigned lo = (unsigned)q; asm volatile("cvt.rn.f32.u32 %0, %0;" : "+r"(lo)); q = (q & 0xffffffff00000000ull) | lo; })
OP(LOP3_FMNMX, 2, asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(r)); { unsigned lo = (unsigned)q; asm volatile("min.f32 %0, %0, 0f437F0000;" : "+r"(lo)); q = (q & 0xffffffff00000000ull) | lo; })
OP(LOP3_IDP,   2, asm volatile("lop3.b32 %0, %0, 0x5a5a5a5a, 0x0f0f0f0f, 0x96;" : "+r"(r)); { unsigned lo = (unsigned)q; asm volatile("dp2a.lo.u32.u32 %0, 0x4b230e97, %0, 0x4000;" : "+r"(lo)); q = (q & 0xffffffff00000000ull) | lo; })

template <class Op>
void run(unsigned long long* d_cycles, unsigned* d_sink, int sms)
{
    bench<Op><<<sms, 512>>>(d_cycles, d_sink, 3);
    bench<Op><<<sms, 512>>>(d_cycles, d_sink, 5);
    cudaDeviceSynchronize();
    unsigned long long h[256];
    cudaMemcpy(h, d_cycles, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double)h[i];
    mean /= sms;
    const double warp_instr = 16.0 * ITERS * CHAINS * Op::count;      // per SM
    printf("%-20s %8.3f warp-instr/clk/SM  (%6.2f clk per warp-instr per SMSP)\n", Op::label, warp_instr / mean,
           mean / (warp_instr / 4.0));
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    unsigned long long* d_cycles;
    unsigned* d_sink;
    cudaMalloc(&d_cycles, 256 * sizeof(unsigned long long));
    cudaMalloc(&d_sink, 64);
    run<FADD>(d_cycles, d_sink, sms); run<FMUL>(d_cycles, d_sink, sms); run<FFMA>(d_cycles, d_sink, sms);
    run<FFMA_R>(d_cycles, d_sink, sms); run<FADD_RZ>(d_cycles, d_sink, sms);
    run<FADD2>(d_cycles, d_sink, sms); run<FMUL2>(d_cycles, d_sink, sms); run<FFMA2>(d_cycles, d_sink, sms);
    run<LOP3>(d_cycles, d_sink, sms); run<PRMT>(d_cycles, d_sink, sms); run<SHF>(d_cycles, d_sink, sms);
    run<IADD3>(d_cycles, d_sink, sms); run<IMAD>(d_cycles, d_sink, sms); run<IDP2A>(d_cycles, d_sink, sms);
    run<I2FP_U>(d_cycles, d_sink, sms); run<I2F_S>(d_cycles, d_sink, sms); run<I2F_U8>(d_cycles, d_sink, sms);
    run<F2I_RZ>(d_cycles, d_sink, sms); run<F2I_RN>(d_cycles, d_sink, sms); run<F2I_U8>(d_cycles, d_sink, sms);
    run<FMNMX>(d_cycles, d_sink, sms); run<SELP>(d_cycles, d_sink, sms); run<MUFU_RCP>(d_cycles, d_sink, sms);
    run<FRND>(d_cycles, d_sink, sms);
    run<FFMA2_LOP3>(d_cycles, d_sink, sms); run<FADD2_LOP3>(d_cycles, d_sink, sms); run<FADD_LOP3>(d_cycles, d_sink, sms);
    run<FADD2_FADD2_LOP3>(d_cycles, d_sink, sms); run<FADD_FADD>(d_cycles, d_sink, sms);
    run<FADD_IDP>(d_cycles, d_sink, sms); run<FADD_I2FP>(d_cycles, d_sink, sms); run<FADD_F2I>(d_cycles, d_sink, sms);
    run<LOP3_PRMT>(d_cycles, d_sink, sms); run<LOP3_I2FP>(d_cycles, d_sink, sms); run<LOP3_FMNMX>(d_cycles, d_sink, sms);
    run<LOP3_IDP>(d_cycles, d_sink, sms);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
