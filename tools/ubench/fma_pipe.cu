// Microbenchmark: issue/pipe throughput of scalar vs packed FP32 on sm_100a (FFMA, FMUL, FADD vs FFMA2, FMUL2, FADD2,
// and mixes).  Prints warp-instructions per cycle per SM sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// --fmad=false -O3 -o fma_pipe fma_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kAcc = 8;       // independent accumulators per thread (ILP)
constexpr int kIter = 4096;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) bench(float *out, float a0, float b0)
{
    float2 acc[kAcc];
    float s[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) { acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f); s[i] = threadIdx.x * 1e-3f - i; }
    const float2 a = make_float2(a0, a0 * 1.0001f), b = make_float2(b0, b0 * 0.9999f);
    const float2 aa = make_float2(a0, a0);
    for (int it = 0; it < kIter; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < kAcc; ++i) {
                if (MODE == 0) { s[i] = __fmaf_rn(s[i], a0, b0); }                                   // FFMA scalar
                if (MODE == 1) { acc[i] = __ffma2_rn(acc[i], a, b); }                                // FFMA2, full 64-bit operands
                if (MODE == 2) { acc[i] = __ffma2_rn(acc[i], aa, b); }                               // FFMA2, one broadcast operand
                if (MODE == 3) { acc[i] = __fmul2_rn(acc[i], a); }                                   // FMUL2
                if (MODE == 4) { acc[i] = __fadd2_rn(acc[i], a); }                                   // FADD2
                if (MODE == 5) { acc[i] = __ffma2_rn(acc[i], a, b); s[i] = __fmaf_rn(s[i], a0, b0); }   // 1 FFMA2 : 1 FFMA
                if (MODE == 6) { acc[i] = __ffma2_rn(acc[i], a, b); s[i] = __fmaf_rn(s[i], a0, b0); s[i] = __fmaf_rn(s[i], b0, a0); }  // 1 : 2
                if (MODE == 7) { s[i] = __fmul_rn(s[i], a0); }                                       // FMUL scalar
                if (MODE == 8) { acc[i] = __ffma2_rn(acc[i], a, b); s[i] = (float)(__float_as_int(s[i]) + i) ; }   // FFMA2 + int op (ALU pipe)
                if (MODE == 10) { acc[i] = __ffma2_rn(acc[i], make_float2(0.2249999f, 0.2249999f), b); }                 // FFMA2, immediate multiplier
                if (MODE == 11) { acc[i] = __fmul2_rn(acc[i], make_float2(1.0001f, 1.0001f)); }                           // FMUL2, immediate
                if (MODE == 12) { acc[i] = __ffma2_rn(a, make_float2(0.2249999f, 0.2249999f), acc[i]); }                 // FFMA2 acc + a*imm (accumulate form)
                if (MODE == 13) { acc[0] = __ffma2_rn(acc[0], a, b); }                                                    // dependent chain: latency
                if (MODE == 14) { s[0] = __fmaf_rn(s[0], a0, b0); }                                                       // scalar dependent chain
                if (MODE == 15) { acc[i] = __ffma2_rn(acc[i], acc[(i + 1) & 7], acc[(i + 3) & 7]); }                      // FFMA2, three distinct 64-bit regs
                if (MODE == 16) { s[i] = __fmaf_rn(s[i], s[(i + 1) & 7], s[(i + 3) & 7]); }                                 // FFMA, three distinct registers
                if (MODE == 17) { s[i] = __fmaf_rn(s[i], s[(i + 1) & 7], b0); }                                           // FFMA, two registers + uniform
                if (MODE == 18) { acc[i] = __fmul2_rn(acc[i], acc[(i + 1) & 7]); }                                        // FMUL2, two distinct pairs
                if (MODE == 19) { acc[i] = __ffma2_rn(acc[(i + 1) & 7], make_float2(0.2249999f, 0.2249999f), acc[i]); }  // FFMA2, two distinct pairs + imm
                if (MODE == 20) { s[i] = __fmul_rn(s[i], s[(i + 1) & 7]); }                                               // FMUL, two distinct registers
                if (MODE == 9) { acc[i].x = __fmaf_rn(acc[i].x, a.x, b.x); acc[i].y = __fmaf_rn(acc[i].y, a.y, b.y); }   // two scalar FFMA on a pair
            }
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < kAcc; ++i) r += acc[i].x + acc[i].y + s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
void run(const char *name, double inst_per_inner, int sms, double clock_ghz, int threads = 1024)
{
    float *out;
    cudaMalloc(&out, sizeof(float) * sms * 1024);
    bench<MODE><<<sms, threads>>>(out, 1.0001f, 0.0001f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<sms, threads>>>(out, 1.0001f, 0.0001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_inst = (double)sms * (threads / 32) /*warps*/ * kIter * 4 * kAcc * inst_per_inner;
    const double cycles = ms * 1e-3 * clock_ghz * 1e9;
    printf("%-44s %8.3f ms  %6.3f warp-inst/clk/SMSP  (%.3f cyc per inner group per warp-slot)\n", name, ms,
           warp_inst / cycles / (sms * 4.0), cycles * sms * 4.0 / (warp_inst / inst_per_inner));
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz (max boost clock; real clock may differ)\n", p.name, p.multiProcessorCount, ghz);
    const int sms = p.multiProcessorCount;
    run<0>("FFMA (scalar, reg*uniform+uniform)", 1, sms, ghz);
    run<7>("FMUL (scalar)", 1, sms, ghz);
    run<1>("FFMA2 (3 x 64-bit operands)", 1, sms, ghz);
    run<2>("FFMA2 (one broadcast operand)", 1, sms, ghz);
    run<3>("FMUL2", 1, sms, ghz);
    run<4>("FADD2", 1, sms, ghz);
    run<5>("1 FFMA2 : 1 FFMA", 2, sms, ghz);
    run<6>("1 FFMA2 : 2 FFMA", 3, sms, ghz);
    run<8>("1 FFMA2 : (IADD + I2F)", 3, sms, ghz);
    run<9>("2 scalar FFMA on a register pair", 2, sms, ghz);
    run<10>("FFMA2 acc*imm+reg", 1, sms, ghz);
    run<11>("FMUL2 acc*imm", 1, sms, ghz);
    run<12>("FFMA2 reg*imm+acc", 1, sms, ghz);
    run<13>("FFMA2 dependent chain, 1 warp/SMSP (latency = 1/rate)", 1, sms, ghz, 128);
    run<14>("FFMA dependent chain, 1 warp/SMSP", 1, sms, ghz, 128);
    run<1>("FFMA2 3x64-bit, ILP 8, 1 warp/SMSP", 1, sms, ghz, 128);
    run<1>("FFMA2 3x64-bit, ILP 8, 2 warps/SMSP", 1, sms, ghz, 256);
    run<1>("FFMA2 3x64-bit, ILP 8, 4 warps/SMSP", 1, sms, ghz, 512);
    run<15>("FFMA2 three distinct register pairs", 1, sms, ghz);
    run<16>("FFMA three distinct registers", 1, sms, ghz);
    run<17>("FFMA two registers + uniform", 1, sms, ghz);
    run<20>("FMUL two distinct registers", 1, sms, ghz);
    run<18>("FMUL2 two distinct pairs", 1, sms, ghz);
    run<19>("FFMA2 two distinct pairs + imm", 1, sms, ghz);
    return 0;
}
