// Probe: a TMA tensor map whose row stride (32 doubles) is SMALLER than its inner extent (40 doubles), i.e. rows that
// overlap in memory.  This is what lets the 3-D z-march kernel run 2-D (x, z) grids: a line of Nx cells is viewed as
// rows of 32 cells with a 4-cell halo on each side (row r = cells [32 r - 4, 32 r + 36)).  Also checks that rows
// outside [0, R) (negative / beyond the extent) are zero-filled.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include <vector>
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
#define BX 40
#define BY 14
__global__ void k(const __grid_constant__ CUtensorMap m, double *out, int y, int z)
{
    extern __shared__ __align__(1024) double raw[];
    unsigned sbase = (unsigned)__cvta_generic_to_shared(raw);
    unsigned bar = sbase + 8 * BX * BY;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(BX * BY * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(sbase), "l"(&m), "r"(0), "r"(y), "r"(z), "r"(bar) : "memory");
    }
    __syncwarp();
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(bar), "r"(0) : "memory");
    for (int i = threadIdx.x; i < BX * BY; i += blockDim.x) out[i] = raw[i];
}
int main()
{
    const int Nx = 400, pitch = 432, plane = 2 * pitch, planes = 12, xoff = 16, R = (Nx + 31) / 32;
    std::vector<double> h((size_t)plane * planes);
    for (size_t i = 0; i < h.size(); ++i) h[i] = 1.0 + (double)i;
    double *d, *o;
    cudaMalloc(&d, h.size() * 8); cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&o, BX * BY * 8);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    enc_fn enc = (enc_fn)p;
    CUtensorMap m;
    cuuint64_t dim[3] = {40, (cuuint64_t)R, (cuuint64_t)planes}; cuuint64_t str[2] = {32 * 8, (cuuint64_t)plane * 8};
    cuuint32_t box[3] = {BX, BY, 1}, es[3] = {1, 1, 1};
    CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d + xoff - 4, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode overlapping-stride map: %d\n", (int)r);
    if (r != CUDA_SUCCESS) return 1;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * BX * BY + 64);
    int total_bad = 0;
    for (int y0 : {-3, 2, R - 8}) {
        const int z = 5;
        k<<<1, 128, 8 * BX * BY + 64>>>(m, o, y0, z);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<double> ho(BX * BY);
        cudaMemcpy(ho.data(), o, ho.size() * 8, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int j = 0; j < BY; ++j) for (int i = 0; i < BX; ++i) {
            const int row = y0 + j;
            double want = (row < 0 || row >= R) ? 0.0 : 1.0 + (double)(xoff - 4 + i + 32 * row + (size_t)plane * z);
            if (ho[j * BX + i] != want) ++bad;
        }
        printf("rows %d..%d: %s, mismatches %d (first %g, last %g)\n", y0, y0 + BY - 1, cudaGetErrorString(e), bad, ho[0], ho[BX * BY - 1]);
        total_bad += bad;
    }
    printf(total_bad ? "FAIL\n" : "OK: overlapping rows load as expected, out-of-range rows are zero-filled\n");
    return total_bad != 0;
}
