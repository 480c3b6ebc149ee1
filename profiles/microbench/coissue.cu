// Co-issue microbenchmark for sm_100a (B200), round 2: do the FP32 pipe (FADD2 / FFMA2) and the
// ALU pipe (LOP3 / SHF / PRMT) really overlap when the instructions read DISTINCT registers, as
// the DCT-QIM kernels' do (pipes.cu uses one register for every operand), and when the two
// kinds of work come from different warps of an SM sub-partition rather than from one warp?
// One CTA of 512 threads (4 warps per sub-partition) per SM, 8 chains per thread.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o coissue coissue.cu ; run: ./coissue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 256
#define N 8

typedef unsigned long long u64;

// MODE: which instruction stream a warp runs.  fp = packed FP32 per step, alu = ALU ops per step.
template <int FP, int ALU, bool SPLIT_WARPS, bool FMA3>
__global__ void __launch_bounds__(512, 1) bench(u64* cycles, unsigned* sink, unsigned seed)
{
    unsigned r[N];
    u64 q[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        r[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
        q[i] = ((u64)(0x3f800000u + i) << 32) | (0x3f900000u + threadIdx.x);
    }
    const int warp = threadIdx.x >> 5;
    // SPLIT_WARPS: warps 0-7 (two per sub-partition) run only the FP stream, warps 8-15 only the ALU stream
    const bool do_fp = !SPLIT_WARPS || warp < 8, do_alu = !SPLIT_WARPS || warp >= 8;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (do_fp) {
#pragma unroll
                for (int k = 0; k < FP; ++k) {
                    if (FMA3) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(q[i]) : "l"(q[(i + 1 + k) % N]), "l"(q[(i + 3 + k) % N]));
                    else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(q[i]) : "l"(q[(i + 1 + k) % N]));
                }
            }
            if (do_alu) {
#pragma unroll
                for (int k = 0; k < ALU; ++k)
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(r[(i + 1 + k) % N]), "r"(r[(i + 3 + k) % N]));
            }
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) acc ^= r[i] ^ (unsigned)q[i] ^ (unsigned)(q[i] >> 32);
    if (acc == 0x12345678u) sink[0] = acc;
    __syncthreads();
    if (threadIdx.x == 0) cycles[blockIdx.x] = (u64)(t1 - t0);
}

template <int FP, int ALU, bool SPLIT, bool FMA3>
void run(const char* label, u64* d_cycles, unsigned* d_sink, int sms)
{
    bench<FP, ALU, SPLIT, FMA3><<<sms, 512>>>(d_cycles, d_sink, 3);
    bench<FP, ALU, SPLIT, FMA3><<<sms, 512>>>(d_cycles, d_sink, 5);
    cudaDeviceSynchronize();
    u64 h[256];
    cudaMemcpy(h, d_cycles, sms * sizeof(u64), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < sms; ++i) mean += (double)h[i];
    mean /= sms;
    // per sub-partition: 4 warps; SPLIT: 2 run FP, 2 run ALU
    const double steps = (double)ITERS * N;
    const double fp_instr = steps * FP * (SPLIT ? 2 : 4), alu_instr = steps * ALU * (SPLIT ? 2 : 4);
    printf("%-44s %8.0f clk | FP32 pipe busy %5.1f %% | ALU pipe busy %5.1f %% | issue %5.1f %%\n", label, mean,
           100.0 * fp_instr * 2 / mean, 100.0 * alu_instr * 2 / mean, 100.0 * (fp_instr + alu_instr) / mean);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs; packed FP32 and LOP3 are both 2 clk per warp-instruction on their pipe\n", p.name, sms);
    u64* d_cycles;
    unsigned* d_sink;
    cudaMalloc(&d_cycles, 256 * sizeof(u64));
    cudaMalloc(&d_sink, 64);
    run<1, 0, false, false>("FADD2 (2 distinct regs) alone", d_cycles, d_sink, sms);
    run<1, 0, false, true>("FFMA2 (3 distinct regs) alone", d_cycles, d_sink, sms);
    run<0, 1, false, false>("LOP3 (3 distinct regs) alone", d_cycles, d_sink, sms);
    run<1, 1, false, false>("same warp: 1 FADD2 + 1 LOP3", d_cycles, d_sink, sms);
    run<1, 1, false, true>("same warp: 1 FFMA2 + 1 LOP3", d_cycles, d_sink, sms);
    run<2, 1, false, false>("same warp: 2 FADD2 + 1 LOP3", d_cycles, d_sink, sms);
    run<3, 2, false, false>("same warp: 3 FADD2 + 2 LOP3", d_cycles, d_sink, sms);
    run<3, 2, false, true>("same warp: 3 FFMA2 + 2 LOP3", d_cycles, d_sink, sms);
    run<1, 1, true, false>("split warps: 2 warps FADD2, 2 warps LOP3", d_cycles, d_sink, sms);
    run<1, 1, true, true>("split warps: 2 warps FFMA2, 2 warps LOP3", d_cycles, d_sink, sms);
    run<2, 1, true, false>("split warps: FADD2 x2 per step vs LOP3 x1", d_cycles, d_sink, sms);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
