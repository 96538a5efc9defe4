// minimal TMA probe: which box shapes / coordinates work
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int VZ, int VY>
__global__ void probe(const __grid_constant__ CUtensorMap tm, float *out, int c0, int c1, int c2)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *smem = (unsigned char *)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
    float *tile = (float *)smem;
    uint64_t *bar = (uint64_t *)(smem + ((VZ * VY * 4 + 127) / 128) * 128);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(VZ * VY * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     ::"r"(smem_u32(tile)), "l"(&tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
    }
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n"
                 ::"r"(smem_u32(bar)), "r"(0) : "memory");
    for (int i = threadIdx.x; i < VZ * VY; i += blockDim.x) out[i] = tile[i];
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int VZ, int VY> int run(EncodeTiledFn encode, float *d, int d1, int d2, int d3, int pitch, int c0, int c1, int c2, const std::vector<float> &h)
{
    CUtensorMap tm;
    cuuint64_t gdim[3] = {(cuuint64_t)d3, (cuuint64_t)d2, (cuuint64_t)d1};
    cuuint64_t gstride[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * d2 * 4};
    cuuint32_t box[3] = {VZ, VY, 1}, estr[3] = {1, 1, 1};
    CUresult rc = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%d coords (%d,%d,%d): encode rc=%d ", VZ, VY, c0, c1, c2, (int)rc);
    if (rc) { printf("\n"); return 1; }
    float *out;
    cudaMalloc(&out, VZ * VY * 4);
    int smem = ((VZ * VY * 4 + 127) / 128) * 128 + 256;
    cudaFuncSetAttribute(probe<VZ, VY>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<VZ, VY><<<1, 128, smem>>>(tm, out, c0, c1, c2);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> o(VZ * VY);
        cudaMemcpy(o.data(), out, VZ * VY * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int y = 0; y < VY; ++y)
            for (int z = 0; z < VZ; ++z) {
                int gy = c1 + y, gz = c0 + z;
                float want = (gy >= 0 && gy < d2 && gz >= 0 && gz < d3) ? h[((size_t)c2 * d2 + gy) * pitch + gz] : 0.f;
                if (o[y * VZ + z] != want) ++bad;
            }
        printf("mismatches=%d", bad);
    }
    printf("\n");
    cudaFree(out);
    return e != cudaSuccess;
}
int main(int argc, char **argv)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeTiledFn encode = (EncodeTiledFn)fn;
    const int d1 = 20, d2 = 50, d3 = 77, pitch = 96;
    std::vector<float> h((size_t)d1 * d2 * pitch);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 9973) + 1.0f;
    float *d;
    cudaMalloc(&d, h.size() * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    int c0 = atoi(argv[1]), c1 = atoi(argv[2]), c2 = atoi(argv[3]), big = atoi(argv[4]);
    if (big) run<68, 20>(encode, d, d1, d2, d3, pitch, c0, c1, c2, h);
    else run<64, 16>(encode, d, d1, d2, d3, pitch, c0, c1, c2, h);
    return 0;
}
