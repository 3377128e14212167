#include <cuda_runtime.h>
#include <stdio.h>
template <int ILP, int MIX>
__global__ void k(double *out, long long *cyc, int iters, double a, double b)
{
    double x[ILP];
    int y = threadIdx.x;
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3 + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
            if (MIX) {
#pragma unroll
                for (int i = 0; i < MIX; ++i) y = y * 3 + r + i;      // integer filler: MIX per ILP DFMAs
            }
        }
    }
    long long t1 = clock64();
    double s = y;
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP, int MIX>
void run(int warps_per_smsp)
{
    double *o; long long *c; cudaMalloc(&o, 148 * 1024 * 8); cudaMalloc(&c, 8);
    int threads = 128 * warps_per_smsp, iters = 2048;
    k<ILP, MIX><<<148, threads>>>(o, c, iters, 1.0000001, 1e-9);
    k<ILP, MIX><<<148, threads>>>(o, c, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    double dfma_per_warp = (double)iters * 8 * ILP;
    printf("warps/SMSP=%d ILP=%d MIX=%d: cycles=%lld  cycles per DFMA per warp=%.2f  DFMA warp-instr/clk/SMSP=%.3f\n", warps_per_smsp, ILP, MIX, h,
           h / dfma_per_warp, dfma_per_warp * warps_per_smsp / h);
    cudaFree(o); cudaFree(c);
}
int main()
{
    run<1, 0>(1); run<2, 0>(1); run<4, 0>(1); run<8, 0>(1);
    run<1, 0>(4); run<2, 0>(4); run<4, 0>(4);
    run<1, 0>(3); run<2, 0>(3);
    run<1, 0>(8);
    run<2, 1>(4); run<2, 2>(4); run<2, 4>(4); run<4, 4>(4); run<4, 8>(4);
    return 0;
}
