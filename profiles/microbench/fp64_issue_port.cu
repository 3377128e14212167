#include <cuda_runtime.h>
#include <stdio.h>
// NF x4 DFMAs (4 independent chains) and NI x4 FP32 FFMAs / integer ops (4 independent chains) per inner step
template <int NF, int NI, int KIND>
__global__ void k(double *out, long long *cyc, int iters, double a, double b, float fa, float fb)
{
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
    float y0 = threadIdx.x, y1 = y0 + 1, y2 = y0 + 2, y3 = y0 + 3;
    unsigned z0 = threadIdx.x, z1 = z0 + 1, z2 = z0 + 2, z3 = z0 + 3;
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
            for (int i = 0; i < NI; ++i) {
                if (KIND == 0) {
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y0) : "f"(fa), "f"(fb));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y1) : "f"(fa), "f"(fb));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y2) : "f"(fa), "f"(fb));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(y3) : "f"(fa), "f"(fb));
                } else {
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(z0) : "r"(z1));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(z1) : "r"(z2));
                    asm volatile("xor.b32 %0, %0, %1;" : "+r"(z2) : "r"(z3));
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(z3) : "r"(z0));
                }
            }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + y0 + y1 + y2 + y3 + z0 + z1 + z2 + z3;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int NF, int NI, int KIND>
void run(int wps)
{
    double *o; long long *c; cudaMalloc(&o, 148 * 1024 * 8); cudaMalloc(&c, 8);
    int iters = 1024;
    k<NF, NI, KIND><<<148, 128 * wps>>>(o, c, iters, 1.0000001, 1e-9, 1.0001f, 1e-6f);
    k<NF, NI, KIND><<<148, 128 * wps>>>(o, c, iters, 1.0000001, 1e-9, 1.0001f, 1e-6f);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    double nf = (double)iters * 4 * NF * 4 * wps, ni = (double)iters * 4 * NI * 4 * wps;
    printf("wps=%d NF=%d NI=%d kind=%s: cycles=%lld | per SMSP: fp64 instr/clk=%.3f other instr/clk=%.3f total=%.3f | model serial(2F+I)=%.0f overlap max(2F,F+I)=%.0f\n",
           wps, NF, NI, KIND ? "int" : "f32", h, nf / h, ni / h, (nf + ni) / h, 2 * nf + ni, (2 * nf > nf + ni ? 2 * nf : nf + ni));
    cudaFree(o); cudaFree(c);
}
int main()
{
    run<1, 0, 0>(4); run<1, 1, 0>(4); run<1, 2, 0>(4); run<2, 1, 0>(4); run<1, 1, 1>(4); run<1, 2, 1>(4); run<2, 1, 1>(4); run<2, 3, 1>(4);
    run<1, 1, 1>(3); run<1, 1, 1>(2);
    return 0;
}
