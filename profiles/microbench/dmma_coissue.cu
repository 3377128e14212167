// Does mma.sync.m8n8k4.f64 (DMMA) issue beside DFMA on sm_100, i.e. could the linear forms of a WENO face leave the
// FP64 vector port?  NF x4 independent DFMA chains + ND x4 independent DMMA chains per inner step, 4 warps per
// sub-partition.  One DMMA = 8x8x4 = 256 FMA = 8 per lane.
#include <cuda_runtime.h>
#include <stdio.h>
template <int NF, int ND>
__global__ void k(double *out, long long *cyc, int iters, double a, double b)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    double c00 = 0, c01 = 0, c10 = 0, c11 = 0, c20 = 0, c21 = 0, c30 = 0, c31 = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NF; ++i) {
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x0) : "d"(a), "d"(b));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x1) : "d"(a), "d"(b));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x2) : "d"(a), "d"(b));
                asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x3) : "d"(a), "d"(b));
            }
#pragma unroll
            for (int i = 0; i < ND; ++i) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c00), "+d"(c01) : "d"(a), "d"(b));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c10), "+d"(c11) : "d"(a), "d"(b));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c20), "+d"(c21) : "d"(a), "d"(b));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c30), "+d"(c31) : "d"(a), "d"(b));
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + c00 + c01 + c10 + c11 + c20 + c21 + c30 + c31;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NF, int ND>
void run(int wps)
{
    double *o; long long *c; cudaMalloc(&o, 148 * 1024 * 8); cudaMalloc(&c, 8);
    int iters = 1024;
    k<NF, ND><<<148, 128 * wps>>>(o, c, iters, 1.0000001, 1e-9);
    k<NF, ND><<<148, 128 * wps>>>(o, c, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    double nf = (double)iters * 4 * NF * 4 * wps, nd = (double)iters * 4 * ND * 4 * wps;
    printf("wps=%d DFMA=%d DMMA=%d: cycles=%lld | per SMSP per clk: DFMA %.3f DMMA %.3f | cycles per DFMA-only would be %.0f; FMA/clk/SMSP total %.1f\n",
           wps, NF, ND, h, nf / h, nd / h, 2 * nf, (nf * 32 + nd * 256) / h);
    cudaFree(o); cudaFree(c);
}
int main()
{
    run<1, 0>(4); run<0, 1>(4); run<1, 1>(4); run<2, 1>(4); run<4, 1>(4); run<8, 1>(4); run<0, 1>(1); run<0, 1>(2);
    return 0;
}
