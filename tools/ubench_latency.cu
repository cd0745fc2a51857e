// ubench_latency.cu -- dependent-chain latency (cycles per instruction, one warp on one SM) of the
// instructions on the serial chain of the DP recurrence.  Build: nvcc -arch=sm_100a -O3 -o ubench_latency ubench_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

template <int OP> __global__ void k(float *out, float seed, double dseed, long long *cycles)
{
    float f = seed + threadIdx.x;
    double d = dseed + threadIdx.x;
    unsigned u = threadIdx.x * 7 + 1;
    long long t0 = clock64();
#pragma unroll 16
    for (int it = 0; it < ITERS; ++it) {
        if (OP == 0) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f)); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(d)); }
        if (OP == 1) d = __dadd_rn(d, dseed);
        if (OP == 2) d = __dmul_rn(d, dseed);
        if (OP == 3) f = __fadd_rn(f, seed);
        if (OP == 4) f = __shfl_up_sync(0xffffffffu, f, 1);
        if (OP == 5) asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %0, %1, p;}" : "+f"(f) : "f"(seed));
        if (OP == 6) asm volatile("max.f32 %0, %0, %1;" : "+f"(f) : "f"(seed));
        if (OP == 7) u = (u & 0x8fffffffu) + 0x38000000u;
        if (OP == 8) { asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(f)); d = __dadd_rn(d, dseed); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(f) : "d"(d)); }
        if (OP == 9) asm volatile("ex2.approx.f32 %0, %0;" : "+f"(f));
        if (OP == 10) d = __fma_rn(d, dseed, dseed);
        if (OP == 11) { long long x = __double_as_longlong(d); x += 0x0010000000000000ll; d = __longlong_as_double(x); d = __dadd_rn(d, dseed); }
    }
    long long t1 = clock64();
    out[threadIdx.x] = f + (float)d + (float)u;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}

template <int OP> void run(const char *name, int ops)
{
    float *out; long long *cyc;
    cudaMalloc(&out, 1024 * 4); cudaMalloc(&cyc, 8);
    k<OP><<<1, 32>>>(out, 1.5f, 1.000001, cyc);
    k<OP><<<1, 32>>>(out, 1.5f, 1.000001, cyc);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %7.1f cycles per iteration (%d dependent ops)\n", name, (double)h / ITERS, ops);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("F2F.F64.F32 -> F2F.F32.F64", 2);
    run<1>("DADD", 1);
    run<2>("DMUL", 1);
    run<10>("DFMA", 1);
    run<3>("FADD", 1);
    run<4>("SHFL.UP", 1);
    run<5>("FSETP + FSEL", 2);
    run<6>("FMNMX", 1);
    run<7>("LOP3 + IADD", 2);
    run<8>("F2F.F64.F32 -> DADD -> F2F.F32.F64", 3);
    run<9>("MUFU.EX2", 1);
    run<11>("IADD64(hi) -> DADD", 2);
    return 0;
}
