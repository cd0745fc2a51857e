// ubench_pipes.cu -- measures per-SM throughput (lanes/clk) of the instructions the DP recurrence
// leans on, to find which pipe bounds it on B200.  Build: nvcc -arch=sm_100a -O3 -o ubench ubench_pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP> __global__ void k(float *out, float seed, double dseed, long long *cycles)
{
    float f[ILP];
    double d[ILP];
    unsigned u[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { f[i] = seed + i + threadIdx.x; d[i] = dseed + i + threadIdx.x; u[i] = threadIdx.x * 7 + i; }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) d[i] = (double)f[i], f[i] = __double2float_rn(d[i]) ;           // F2F both ways (2 ops)
            if (OP == 1) asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d[i]) : "f"(f[i]));  // F2F.F64.F32
            if (OP == 2) asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f[i]) : "d"(d[i]));// F2F.F32.F64
            if (OP == 3) d[i] = __dadd_rn(d[i], dseed);
            if (OP == 4) d[i] = __dmul_rn(d[i], dseed);
            if (OP == 5) f[i] = __fadd_rn(f[i], seed);
            if (OP == 6) u[i] = (u[i] & 0x8fffffffu) + 0x38000000u;
            if (OP == 7) asm volatile("{.reg .pred p; setp.gt.f64 p, %1, %2; selp.u32 %0, 1, 0, p;}" : "=r"(u[i]) : "d"(d[i]), "d"(dseed));
            if (OP == 8) f[i] = __shfl_up_sync(0xffffffffu, f[i], 1);
            if (OP == 9) asm volatile("ex2.approx.f32 %0, %1;" : "=f"(f[i]) : "f"(f[i]));
        }
    }
    long long t1 = clock64();
    float acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc += f[i] + (float)d[i] + (float)u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char *name, int ops_per_iter, int threads)
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    k<OP><<<148, threads>>>(out, 1.5f, 1.000001, cyc);
    k<OP><<<148, threads>>>(out, 1.5f, 1.000001, cyc);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    double lanes = (double)threads * ITERS * ILP * ops_per_iter / avg;
    printf("%-22s threads/SM=%4d  %8.1f cycles  %6.2f lanes/clk/SM\n", name, threads, avg, lanes);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int threads : {128, 512, 1024}) {
        run<1>("F2F.F64.F32", 1, threads);
        run<2>("F2F.F32.F64", 1, threads);
        run<0>("F2F roundtrip (2 ops)", 2, threads);
        run<3>("DADD", 1, threads);
        run<4>("DMUL", 1, threads);
        run<7>("DSETP+SEL", 1, threads);
        run<5>("FADD", 1, threads);
        run<6>("LOP3+IADD (2 ops)", 2, threads);
        run<8>("SHFL.UP", 1, threads);
        run<9>("MUFU.EX2", 1, threads);
    }
    return 0;
}
