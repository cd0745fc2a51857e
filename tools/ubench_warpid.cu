// ubench_warpid.cu -- where do the warps of small CTAs land?  Prints, for a grid of 2-warp CTAs that
// all stay resident, the histogram of (%warpid % 4) of warp 0 and warp 1 (the scheduler partition a
// warp slot belongs to) and the number of CTAs per SM.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int *out, int spin)
{
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    long long t0 = clock64();
    while (clock64() - t0 < spin) { }
    if ((threadIdx.x & 31) == 0) {
        out[(blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 2] = smid;
        out[(blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32) * 2 + 1] = warpid;
    }
}
int main()
{
    for (int wpc : {1, 2, 4}) {
        const int n = 592;
        int *d; cudaMalloc(&d, n * wpc * 8);
        k<<<n, 32 * wpc, 12 * 1024>>>(d, 2000000);
        cudaDeviceSynchronize();
        int *h = new int[n * wpc * 2];
        cudaMemcpy(h, d, n * wpc * 8, cudaMemcpyDeviceToHost);
        printf("warps per CTA = %d\n", wpc);
        for (int w = 0; w < wpc; ++w) {
            int hist[4] = {0, 0, 0, 0};
            for (int b = 0; b < n; ++b) hist[h[(b * wpc + w) * 2 + 1] & 3]++;
            printf("  warp %d: warpid%%4 histogram = %d %d %d %d\n", w, hist[0], hist[1], hist[2], hist[3]);
        }
        printf("  first CTAs on SM of block 0 (smid %d): ", h[0]);
        for (int b = 0; b < n; ++b) if (h[b * wpc * 2] == h[0]) { printf("[b%d:", b); for (int w = 0; w < wpc; ++w) printf(" %d", h[(b * wpc + w) * 2 + 1]); printf("] "); }
        printf("\n");
        cudaFree(d); delete[] h;
    }
    return 0;
}
