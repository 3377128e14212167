#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__global__ void k(const __grid_constant__ CUtensorMap m, const float *src, float *out, int mode, const __grid_constant__ CUtensorMap m3, int cx, int cy, int cz)
{
    __shared__ alignas(128) float tile_s[32 * 32];
    extern __shared__ float dyn[];
    float *tile = mode == 3 ? (float *)(((size_t)dyn + 127) & ~(size_t)127) : tile_s;
    __shared__ alignas(8) unsigned long long bar;
    unsigned sb = (unsigned)__cvta_generic_to_shared(&bar), st = (unsigned)__cvta_generic_to_shared(tile);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sb), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(4096) : "memory");
        if (mode == 0)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(st), "l"(&m), "r"(0), "r"(0), "r"(sb) : "memory");
        else if (mode == 2)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                         ::"r"(st), "l"(&m3), "r"(cx), "r"(cy), "r"(cz), "r"(sb) : "memory");
        else if (mode == 3)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(st), "l"(&m), "r"(0), "r"(0), "r"(sb) : "memory");
        else
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(st), "l"(src), "r"(4096), "r"(sb) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(sb), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char **argv)
{
    int mode = argc > 1 ? atoi(argv[1]) : 0;
    std::vector<float> h(64 * 64);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&o, 4096);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t ge = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    printf("entry point: %d %d %p\n", (int)ge, (int)q, p);
    CUtensorMap m; memset(&m, 0, sizeof(m));
    cuuint64_t dim[2] = {64, 64}; cuuint64_t str[1] = {64 * 4};
    cuuint32_t box[2] = {32, 32}, es[2] = {1, 1};
    CUresult r = ((enc_fn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dim, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("enc %d mode %d\n", (int)r, mode);
    const unsigned long long *w = (const unsigned long long *)&m;
    for (int i = 0; i < 16; ++i) printf("%016llx%c", w[i], i % 4 == 3 ? '\n' : ' ');
    CUtensorMap m3; memset(&m3, 0, sizeof(m3));
    cuuint64_t dim3_[3] = {64, 16, 4}; cuuint64_t str3[2] = {64 * 4, 64 * 16 * 4};
    cuuint32_t box3[3] = {32, 32, 1}, es3[3] = {1, 1, 1};
    box3[1] = 16; box3[2] = 2;
    r = ((enc_fn)p)(&m3, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, dim3_, str3, box3, es3, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("enc3 %d\n", (int)r);
    k<<<1, 128, 8192>>>(m, d, o, mode, m3, argc > 2 ? atoi(argv[2]) : 0, argc > 3 ? atoi(argv[3]) : 0, argc > 4 ? atoi(argv[4]) : 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    std::vector<float> ho(1024);
    cudaMemcpy(ho.data(), o, 4096, cudaMemcpyDeviceToHost);
    printf("out[0..3]=%g %g %g %g out[32]=%g\n", ho[0], ho[1], ho[2], ho[3], ho[32]);
    return 0;
}
