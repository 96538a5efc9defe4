// pattern_bench.cu -- probe: what DRAM throughput do multi-stream tile-march access patterns sustain on B200?
//
// The fused step kernel reads 9 and writes 9 streams; its memory skeleton runs at ~4.7 TB/s of DRAM traffic where
// a 2-stream copy reaches ~6.5 TB/s.  This probe separates the candidates: number of streams,
// length of the contiguous row pieces (tile width), reads vs writes, LDG/STG vs TMA.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/scratch/pattern_bench tools/scratch/pattern_bench.cu
//   ./pattern_bench [dimx dimy dimz]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct Ptrs { float4 *p[18]; };

// CTA = tz4 x ty threads (256), marches along x through one x-chunk; NR read streams, NW write streams
template <int NR, int NW, int UNR>
__global__ void __launch_bounds__(256, 2)
tile_march(Ptrs P, int tz4, int ty, int ntz, int nty, int nchunks, int dimx, int dimy, int row4, long pitch4, long plane4)
{
    int b = blockIdx.x;
    const int bz = b % ntz; b /= ntz;
    const int by = b % nty;
    const int bc = b / nty;
    const int t = threadIdx.x;
    const int lz = t % tz4, ly = t / tz4;
    const int z4 = bz * tz4 + lz, y = by * ty + ly;
    if (z4 >= row4 || y >= dimy) return;
    const int x0 = (int)((long)dimx * bc / nchunks), x1 = (int)((long)dimx * (bc + 1) / nchunks);
    const long off = (long)y * pitch4 + z4;
    for (int x = x0; x < x1; x += UNR) {
        float4 v[UNR][NR > 0 ? NR : 1];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int k = 0; k < NR; ++k)
                if (x + u < x1) v[u][k] = __ldg(P.p[k] + (long)(x + u) * plane4 + off);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            float4 s = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
            for (int k = 0; k < NR; ++k) { s.x += v[u][k].x; s.y += v[u][k].y; s.z += v[u][k].z; s.w += v[u][k].w; }
            if (NW == 0) { if (s.x == 123.456f) P.p[9][0] = s; }
#pragma unroll
            for (int k = 0; k < NW; ++k)
                if (x + u < x1) __stcs(P.p[9 + k] + (long)(x + u) * plane4 + off, s);
        }
    }
}

// flat streams: every CTA walks the arrays in 256-thread float4 lines, grid-stride
template <int NR, int NW, int UNR>
__global__ void __launch_bounds__(256, 2) linear_streams(Ptrs P, long n4)
{
    const long stride = (long)gridDim.x * 256 * UNR;
    for (long i = (long)blockIdx.x * 256 * UNR + threadIdx.x; i < n4; i += stride) {
        float4 v[UNR][NR > 0 ? NR : 1];
#pragma unroll
        for (int u = 0; u < UNR; ++u)
#pragma unroll
            for (int k = 0; k < NR; ++k)
                if (i + u * 256 < n4) v[u][k] = __ldg(P.p[k] + i + u * 256);
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            float4 s = make_float4(1.f, 2.f, 3.f, 4.f);
#pragma unroll
            for (int k = 0; k < NR; ++k) { s.x += v[u][k].x; s.y += v[u][k].y; s.z += v[u][k].z; s.w += v[u][k].w; }
            if (NW == 0) { if (s.x == 123.456f) P.p[9][0] = s; }
#pragma unroll
            for (int k = 0; k < NW; ++k)
                if (i + u * 256 < n4) __stcs(P.p[9 + k] + i + u * 256, s);
        }
    }
}

// ---------------------------------------------------------------- TMA tile march: box loads into a ring, box stores out of it
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tmap, const void *src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
struct Maps { CUtensorMap m[18]; };

// One thread per CTA drives everything: per plane, NR box loads into ring slot s, wait, then NW box stores from the slots
// of the first NW loaded boxes (a pure TMA copy pipeline -- the upper bound of what a TMA-fed stencil kernel can move).
// ring: DEPTH planes x NR boxes.  HALO: the loaded box is (bz + 2*hz) x (by + 2*hy); the stored one bz x by (interior).
template <int NR, int NW, int DEPTH>
__global__ void __launch_bounds__(32, 1)
tma_march(const __grid_constant__ Maps LM, const __grid_constant__ Maps SM_, int bzl, int byl, int bzs, int bys, int hz, int hy,
          int ntz, int nty, int nchunks, int dimx)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const uint32_t box_bytes = (uint32_t)bzl * byl * 4;
    const uint32_t box_al = (box_bytes + 127) / 128 * 128;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)DEPTH * NR * box_al);
    int b = blockIdx.x;
    const int bz = b % ntz; b /= ntz;
    const int by = b % nty;
    const int bc = b / nty;
    const int x0 = (int)((long)dimx * bc / nchunks), x1 = (int)((long)dimx * (bc + 1) / nchunks);
    if (threadIdx.x != 0) return;
    for (int i = 0; i < DEPTH; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int c0 = bz * bzs - hz, c1 = by * bys - hy;
    auto issue = [&](int x) {
        const int s = (x - x0) % DEPTH;
        mbar_expect(&bars[s], box_bytes * NR);
#pragma unroll
        for (int k = 0; k < NR; ++k) tma_load_3d(smem + (size_t)(s * NR + k) * box_al, &LM.m[k], &bars[s], c0, c1, x);
    };
    // prologue: DEPTH-1 planes in flight
    for (int x = x0; x < x0 + DEPTH - 1 && x < x1; ++x) issue(x);
    for (int x = x0; x < x1; ++x) {
        const int i = x - x0, s = i % DEPTH;
        mbar_wait(&bars[s], (uint32_t)(i / DEPTH) & 1u);
        // the slot refilled now (plane x + DEPTH - 1 -> slot (i-1)%DEPTH) was stored from at plane x-1: wait until that store has READ it
        if (NW > 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        if (x + DEPTH - 1 < x1) issue(x + DEPTH - 1);
        if (NW > 0) {
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                // interior of box k % NR (store box = load box when there is no halo)
                const unsigned char *src = smem + (size_t)(s * NR + (k % (NR > 0 ? NR : 1))) * box_al;
                tma_store_3d(&SM_.m[k], src, bz * bzs, by * bys, x);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (NW > 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char **argv)
{
    int dimx = 1029, dimy = 1029, dimz = 1029;
    if (argc >= 4) { dimx = atoi(argv[1]); dimy = atoi(argv[2]); dimz = atoi(argv[3]); }
    const long pitch = (dimz + 31) / 32 * 32, pitch4 = pitch / 4, plane4 = pitch4 * dimy, n4 = plane4 * dimx;
    const int row4 = (dimz + 3) / 4;
    printf("dims %d x %d x %d, pitch %ld floats, %.2f GB per array\n", dimx, dimy, dimz, pitch, n4 * 16.0 / 1e9);
    Ptrs P;
    for (int k = 0; k < 18; ++k) { CK(cudaMalloc(&P.p[k], n4 * 16)); CK(cudaMemset(P.p[k], 0, n4 * 16)); }
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int REPS = 3;
    auto report = [&](const char *name, int nr, int nw, double valid_frac, float ms) {
        const double bytes = (double)(nr + nw) * n4 * 16.0 * valid_frac;
        printf("%-44s %2dr+%2dw  %8.3f ms  %7.1f GB/s\n", name, nr, nw, ms / REPS, bytes / (ms / REPS * 1e-3) / 1e9);
        fflush(stdout);
    };
    // ---- cudaMemcpy D2D
    {
        CK(cudaMemcpy(P.p[9], P.p[0], n4 * 16, cudaMemcpyDeviceToDevice));
        CK(cudaEventRecord(e0));
        for (int r = 0; r < REPS; ++r) CK(cudaMemcpyAsync(P.p[9], P.p[0], n4 * 16, cudaMemcpyDeviceToDevice));
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        report("cudaMemcpy D2D", 1, 1, 1.0, ms);
    }
#define RUN_LINEAR(NR, NW, UNR, GRID)                                                              \
    {                                                                                              \
        linear_streams<NR, NW, UNR><<<GRID, 256>>>(P, n4);                                         \
        CK(cudaEventRecord(e0));                                                                   \
        for (int r = 0; r < REPS; ++r) linear_streams<NR, NW, UNR><<<GRID, 256>>>(P, n4);          \
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());             \
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));                                           \
        char nm[96]; snprintf(nm, 96, "linear unr%d grid%d", UNR, GRID);                           \
        report(nm, NR, NW, 1.0, ms);                                                               \
    }
    RUN_LINEAR(1, 1, 4, 148 * 8)
    RUN_LINEAR(1, 1, 8, 148 * 4)
    RUN_LINEAR(9, 9, 2, 148 * 2)
    RUN_LINEAR(9, 9, 2, 148 * 16)
    RUN_LINEAR(9, 0, 2, 148 * 2)
    RUN_LINEAR(0, 9, 2, 148 * 2)
    RUN_LINEAR(3, 3, 4, 148 * 2)
    const double vf = (double)row4 / pitch4;   // tile kernels touch row4 of the pitch4 float4 of a row
#define RUN_TILE(NR, NW, UNR, TZ4, TY, NCH)                                                                                     \
    {                                                                                                                           \
        const int ntz = (row4 + TZ4 - 1) / TZ4, nty = (dimy + TY - 1) / TY;                                                     \
        const int grid = ntz * nty * NCH;                                                                                       \
        tile_march<NR, NW, UNR><<<grid, 256>>>(P, TZ4, TY, ntz, nty, NCH, dimx, dimy, row4, pitch4, plane4);                    \
        CK(cudaEventRecord(e0));                                                                                                \
        for (int r = 0; r < REPS; ++r) tile_march<NR, NW, UNR><<<grid, 256>>>(P, TZ4, TY, ntz, nty, NCH, dimx, dimy, row4, pitch4, plane4); \
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());                                          \
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));                                                                        \
        char nm[96]; snprintf(nm, 96, "tile %4d x %3d (z x y) unr%d chunks%d ctas%d", TZ4 * 4, TY, UNR, NCH, grid);              \
        report(nm, NR, NW, vf, ms);                                                                                             \
    }
    // 18 streams, tile width sweep (chunks chosen so that every grid is a few thousand CTAs)
    RUN_TILE(9, 9, 2, 16, 16, 2)
    RUN_TILE(9, 9, 2, 32, 8, 2)
    RUN_TILE(9, 9, 2, 64, 4, 2)
    RUN_TILE(9, 9, 2, 128, 2, 2)
    RUN_TILE(9, 9, 2, 256, 1, 2)
    RUN_TILE(9, 9, 2, 256, 1, 1)
    RUN_TILE(9, 9, 1, 16, 16, 2)
    RUN_TILE(9, 9, 1, 256, 1, 2)
    // reads only / writes only
    RUN_TILE(9, 0, 2, 16, 16, 2)
    RUN_TILE(9, 0, 2, 256, 1, 2)
    RUN_TILE(0, 9, 2, 16, 16, 2)
    RUN_TILE(0, 9, 2, 256, 1, 2)
    // fewer streams
    RUN_TILE(1, 1, 2, 16, 16, 2)
    RUN_TILE(1, 1, 2, 256, 1, 2)
    RUN_TILE(3, 3, 2, 16, 16, 2)
    RUN_TILE(3, 3, 2, 256, 1, 2)
    // one wave only (296 CTAs at 2 per SM): tiles in lockstep
    RUN_TILE(9, 9, 2, 16, 16, 1)
    RUN_TILE(9, 9, 2, 16, 16, 8)

    // ---- TMA
    EncodeTiledFn encode = nullptr;
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        encode = (EncodeTiledFn)fn;
    }
    auto make_maps = [&](Maps &Mm, int first, int bz, int by) {
        for (int k = 0; k < 9; ++k) {
            cuuint64_t gdim[3] = {(cuuint64_t)dimz, (cuuint64_t)dimy, (cuuint64_t)dimx};
            cuuint64_t gstride[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * dimy * 4};
            cuuint32_t box[3] = {(cuuint32_t)bz, (cuuint32_t)by, 1};
            cuuint32_t estr[3] = {1, 1, 1};
            CUresult rc = encode(&Mm.m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, P.p[first + k], gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (rc != CUDA_SUCCESS) { printf("encode failed %d (box %d x %d)\n", (int)rc, bz, by); exit(1); }
        }
    };
#define RUN_TMA(NR, NW, DEPTH, BZ, BY, HZ, HY, NCH)                                                                              \
    {                                                                                                                           \
        Maps LM, SMm;                                                                                                           \
        const int bzl = BZ + 2 * HZ, byl = BY + 2 * HY;                                                                         \
        make_maps(LM, 0, bzl, byl);                                                                                             \
        make_maps(SMm, 9, BZ, BY);                                                                                              \
        const int ntz = (dimz + BZ - 1) / BZ, nty = (dimy + BY - 1) / BY;                                                       \
        const int grid = ntz * nty * NCH;                                                                                       \
        const size_t box_al = ((size_t)bzl * byl * 4 + 127) / 128 * 128;                                                        \
        const size_t sm = (size_t)DEPTH * (NR > 0 ? NR : 1) * box_al + DEPTH * 8 + 64;                                          \
        if (sm > 227 * 1024) { printf("skip tma %d x %d: smem %zu\n", BZ, BY, sm); }                                            \
        else {                                                                                                                  \
            CK(cudaFuncSetAttribute(tma_march<NR, NW, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));           \
            tma_march<NR, NW, DEPTH><<<grid, 32, sm>>>(LM, SMm, bzl, byl, BZ, BY, HZ, HY, ntz, nty, NCH, dimx);                  \
            CK(cudaEventRecord(e0));                                                                                            \
            for (int r = 0; r < REPS; ++r) tma_march<NR, NW, DEPTH><<<grid, 32, sm>>>(LM, SMm, bzl, byl, BZ, BY, HZ, HY, ntz, nty, NCH, dimx); \
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());                                      \
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));                                                                    \
            char nm[96]; snprintf(nm, 96, "tma %3d x %2d halo %d,%d depth%d ch%d ctas%d smem%zuK", BZ, BY, HZ, HY, DEPTH, NCH, grid, sm / 1024); \
            report(nm, NR, NW, (double)dimz / pitch, ms);                                                                       \
        }                                                                                                                       \
    }
    // pure TMA copy pipelines, 9 in + 9 out, no halo (GB/s counts the useful bytes: 18 x array)
    RUN_TMA(9, 9, 4, 64, 12, 0, 0, 2)
    RUN_TMA(9, 9, 4, 64, 12, 0, 0, 1)
    RUN_TMA(9, 9, 4, 128, 6, 0, 0, 2)
    RUN_TMA(9, 9, 4, 256, 3, 0, 0, 2)
    RUN_TMA(9, 9, 4, 256, 4, 0, 0, 2)
    RUN_TMA(9, 9, 3, 128, 12, 0, 0, 2)
    RUN_TMA(9, 9, 2, 256, 8, 0, 0, 2)
    RUN_TMA(9, 9, 6, 64, 8, 0, 0, 4)
    RUN_TMA(9, 0, 4, 64, 12, 0, 0, 2)
    RUN_TMA(9, 0, 4, 256, 3, 0, 0, 2)
    // the fused kernel's shape: 60 x 12 stored, halo 4 (z, keeps 16-B alignment) and 4 (y) on the loads -- useful bytes only
    RUN_TMA(9, 9, 4, 64, 12, 4, 4, 2)
    RUN_TMA(9, 9, 3, 128, 12, 4, 4, 2)
    RUN_TMA(9, 9, 2, 256, 12, 4, 4, 2)
    RUN_TMA(9, 9, 2, 128, 24, 4, 4, 2)
    RUN_TMA(3, 3, 4, 64, 12, 4, 4, 2)
    printf("done\n");
    return 0;
}
