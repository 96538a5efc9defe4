// fused.cuh -- fused stress+velocity leapfrog step for sm_100a (so <= 4, fp32).
//
// One launch advances ALL nine fields by one time step over the interior:
//   reads  U,V,W[t0] and the six T[t0]      (9 words / point)
//   writes the six T[t1] and U,V,W[t1]      (9 words / point)      => 72 B / point update,
// SURVEY.md 8d's algorithmic figure, instead of the 108 B of the two-pass path (the stress
// field is not re-read from HBM by a second kernel).
//
// Replaces the reference's stress loop + velocity loop of one time step
// (opesci/staggeredgrid.py:728-748 emitted through opesci/regulargrid.py:566-619) wherever the
// velocity update cannot see a ghost-cell / free-surface modification of the new stresses:
// stresses are stored for the whole interior [m, dim-m)^3, velocities only for the "deep"
// interior [2m+1, dim-2m-1)^3 whose stress stencil touches cells no boundary loop rewrites
// (staggeredgrid.py:750-813 writes only cells with an index <= m or >= dim-m-1).  The thin shell
// that is left is updated after the stress ghost loops by velocity_shell_kernel (below), so the
// per-step order stress -> stress BC -> velocity -> velocity BC of
// opesci/templates/staggered3d_tmpl.py:40-58 is preserved cell for cell.
//
// Structure (B200): one CTA owns an (EY-2M) x (EZ-2M) tile of (y,z) and marches along x (the
// slowest axis) through an x-chunk.  Velocity planes arrive by TMA (cp.async.bulk.tensor.3d,
// zero-filled outside the array) into a ring of RD planes per field, signalled through
// mbarriers; every thread computes the new stresses of one point of the EY x EZ tile (the
// outer M rows/columns are recomputed halo, never stored), keeps its own-column x-windows of
// Txx/Txy/Txz in registers, exchanges the in-plane operands (Txy,Txz,Tyy,Tyz,Tzz) through a
// 4-slot shared-memory ring, and M planes later computes the velocities of its point.
#pragma once
#include <cuda.h>

#include "hetero.cuh"
#include "kernels.cuh"

namespace opesci {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tmap, uint64_t *bar, int c0, int c1, int c2)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// windows over pre-gathered register arrays: v[j], j = 0..2M-1,
//   FWD: offset o = j - M + 1 (-M+1..M)      BWD: offset o = j - M (-M..M-1)
template <int M, typename T, bool FWD>
__device__ __forceinline__ void window_ref_arr(T &acc, bool &first, const T *v, const float *c)
{
    if (FWD) {
#pragma unroll
        for (int o = 1; o <= M; ++o) term<T>(acc, first, c[o - 1], v[o + M - 1]);
#pragma unroll
        for (int o = 1; o <= M - 1; ++o) term<T>(acc, first, -c[o], v[-o + M - 1]);
        term<T>(acc, first, -c[0], v[M - 1]);
    } else {
#pragma unroll
        for (int o = 1; o <= M - 1; ++o) term<T>(acc, first, c[o], v[o + M]);
#pragma unroll
        for (int o = 1; o <= M; ++o) term<T>(acc, first, -c[o - 1], v[-o + M]);
        term<T>(acc, first, c[0], v[M]);
    }
}
template <int M, typename T, bool FWD>
__device__ __forceinline__ T window_fast_arr(const T *v, const float *c)
{
    T d = 0;
#pragma unroll
    for (int k = 1; k <= M; ++k) {
        const T a = FWD ? v[k + M - 1] : v[k - 1 + M];
        const T b = FWD ? v[-k + 1 + M - 1] : v[-k + M];
        d += (T)c[k - 1] * (a - b);
    }
    return d;
}

// heterogeneous mode: every term is (c*g)*media; NV = 2 emits the lambda term then the mu term (c2) per offset
template <int M, bool FWD, int NV>
__device__ __forceinline__ void window_ref_arr_h(float &acc, bool &first, const float *v, const float *c, const float *c2,
                                                 float med, float med2)
{
#define OPESCI_EMIT(idx, sgn, k)                                   \
    {                                                              \
        term_h(acc, first, (sgn) * c[k], v[idx], med);             \
        if (NV == 2) term_h(acc, first, (sgn) * c2[k], v[idx], med2); \
    }
    if (FWD) {
#pragma unroll
        for (int o = 1; o <= M; ++o) OPESCI_EMIT(o + M - 1, 1.0f, o - 1)
#pragma unroll
        for (int o = 1; o <= M - 1; ++o) OPESCI_EMIT(-o + M - 1, -1.0f, o)
        OPESCI_EMIT(M - 1, -1.0f, 0)
    } else {
#pragma unroll
        for (int o = 1; o <= M - 1; ++o) OPESCI_EMIT(o + M, 1.0f, o)
#pragma unroll
        for (int o = 1; o <= M; ++o) OPESCI_EMIT(-o + M, -1.0f, o - 1)
        OPESCI_EMIT(M, 1.0f, 0)
    }
#undef OPESCI_EMIT
}

template <int M> struct FusedCfg {
#ifndef OPESCI_FUSED_EZ
#define OPESCI_FUSED_EZ 64
#endif
#ifndef OPESCI_FUSED_EY
#define OPESCI_FUSED_EY 16
#endif
#ifndef OPESCI_FUSED_MINB
#define OPESCI_FUSED_MINB 1
#endif
    static constexpr int EZ = OPESCI_FUSED_EZ, EY = OPESCI_FUSED_EY;     // threads = stress tile (incl. recomputed halo)
    // stored tile.  TMA needs the innermost box coordinate 16-B aligned (measured on B200: an
    // unaligned start raises "illegal instruction"), so tiles advance in multiples of 4 floats
    // along z and the box starts OFFZ >= M floats left of the stress tile.
    // Sector-aligned stores (M = 2): the stored z range of a tile is [bx*CZ, bx*CZ + CZ) with CZ a multiple of 8 floats,
    // so every row a CTA writes is made of whole 32-byte sectors.  With the tile origin at m + bx*60 (ZS = 0) each row
    // started and ended inside a sector shared with the neighbouring CTA; written with streaming stores at different
    // times, those partial sectors reach DRAM as read-modify-writes.  ZS = thread-tile shift: thread column tz is
    // z = bx*CZ + tz - ZS; it has to stay even (8-byte global accesses of two z-adjacent points), hence M = 2 only.
#ifndef OPESCI_FUSED_ALIGN
#define OPESCI_FUSED_ALIGN 0   /* measured on B200: 20.2 vs 19.9 ms -- whole-sector stores, but 19 tile columns of 56 instead of 18 of 60 and 5 % more halo re-reads */
#endif
    static constexpr int ZS = (OPESCI_FUSED_ALIGN && M == 2) ? 2 : 0;
#ifndef OPESCI_EXP_CY_EXTRA
#define OPESCI_EXP_CY_EXTRA 0   /* TIMING PROBE ONLY (wrong results): tile rows advance by EY - 2M + this */
#endif
    static constexpr int CZ = ZS ? (EZ - 2 * M) / 8 * 8 : (EZ - 2 * M) / 4 * 4, CY = EY - 2 * M + OPESCI_EXP_CY_EXTRA;
    static constexpr int OFFZ = (M + ZS + 3) / 4 * 4;
    static constexpr int VZ = (EZ + M + OFFZ - ZS + 3) / 4 * 4, VY = EY + 2 * M;   // velocity tile delivered by TMA
    // number of tiles along z for an array of dim3 = dim
    static constexpr int ztiles(int dim) { return (dim - 2 * M + ZS + CZ - 1) / CZ; }
#ifndef OPESCI_FUSED_RD_EXTRA
#define OPESCI_FUSED_RD_EXTRA 0     /* A/B: extra planes of TMA prefetch */
#endif
#ifndef OPESCI_FUSED_SR
#define OPESCI_FUSED_SR 4           /* (1 only makes sense with OPESCI_SKELETON) */
#endif
    static constexpr int RD = 2 * M + 2 + OPESCI_FUSED_RD_EXTRA;   // velocity ring depth (planes)
    static constexpr int SR = OPESCI_FUSED_SR;             // in-plane stress ring slots
    static constexpr int VTILE = ((VZ * VY * 4 + 127) / 128) * 128;   // bytes, 128-B aligned for TMA
    static constexpr int STILE = EZ * EY * 4;
    static constexpr int SMEM = 3 * RD * VTILE + 5 * SR * STILE + 3 * RD * 8 + 128;
    static constexpr int THREADS = EZ / 2 * EY;             // two z-adjacent points per thread
    // ---- PAIR kernels: two CTAs stacked in y form a thread-block cluster.  Each stores EY - M rows (the recomputed halo
    // rows exist only on the outer side of the pair); the M stress rows one CTA needs from the other for its velocity
    // y-windows arrive through distributed shared memory, so a ring slot holds EY + M rows.
    static constexpr int PCY = EY - M;                      // rows stored per CTA of a pair
    static constexpr int PSTILE = EZ * (EY + M) * 4;
    static constexpr int PSMEM = 3 * RD * VTILE + 5 * SR * PSTILE + 3 * RD * 8 + 128 + 2 * SR * 8;
};

struct FusedArgs {
    FieldPtrs F;
    GridGeom G;
    StaggeredCoefs C;
    MediaPtrs MD;     // heterogeneous mode (HET kernels): per-cell media, +32 B/point of read traffic
    HeteroCoefs HC;
    int t0, t1;
#define OPESCI_MAX_CHUNKS 20
    int xs[OPESCI_MAX_CHUNKS + 1];   // x-chunk c covers planes [xs[c], xs[c+1]); blockIdx.z + chunk0 selects the chunk
    int chunk0;
    int *pace;        // OPESCI_PACE: progress of every tile (plane index; -1 not started; INT_MAX done), else null
    int cluster_sync; // launched as clusters of OPESCI_CLUSTER_Z z-adjacent CTAs (A/B experiment)
    // tile column of a CTA: blockIdx.x + bx0 (interior launch) or bxs[blockIdx.x] (z-edge launch, ZF kernels)
    int bx0;
    int bxs[2];
    // ---- z-fold (ZF kernels): the z-face stress ghost loops and the z slabs of the velocity shell are done in the
    // z-edge tiles, for the cells whose operands only z-face loops touch: planes [zf_xlo, zf_xhi), rows [M+1, dim2-M-1)
    int zf_side[2];   // this launch handles the low / high z face ...
    int zf_bx[2];     // ... in the tiles of this column
    int zf_c[2];      // tile column (0..EZ-1) of the face plane b = M / b' = dim3-M-1 inside tile bxs[side]
    int zf_xlo, zf_xhi;
    float zf_lev[2][2][2];   // Levander recompute on the z faces: [Txx, Tyy][d/dx U, d/dy V][k] (lev_stress[2][e][f])
    // The reference mirrors Txx across the x faces (reading plane m+1 / dim1-m-2) and Tyy across the y faces (reading row
    // m+1 / dim2-m-2) BEFORE the z-face Levander loops rewrite those cells.  On these planes (Txx) / rows (Tyy) the z-edge
    // tiles therefore keep the recomputed value in registers only (their own velocity update needs it) and store the plain
    // interior value; the face kernels recompute the cells afterwards, in the reference's order (Stepper::stress_bc).
    int zf_raw_x[2];         // planes (-1: none -- an artificial slab end has no x face)
    int zf_raw_y[2];         // rows
};

// six consecutive floats p[-2..3] as three aligned 8-byte loads (p must be 8-byte aligned)
__device__ __forceinline__ void load6(const float *p, float *f)
{
    const float2 a = *reinterpret_cast<const float2 *>(p - 2);
    const float2 b = *reinterpret_cast<const float2 *>(p);
    const float2 c = *reinterpret_cast<const float2 *>(p + 2);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y;
}

// global stores of results that nobody re-reads during this launch: streaming (evict-first) so
// they do not push the halo rows that neighbouring CTAs still need out of L2
#ifndef OPESCI_STREAM_STORES
#define OPESCI_STREAM_STORES 1
#endif
#ifndef OPESCI_T0_NOALLOC
#define OPESCI_T0_NOALLOC 0
#endif
// T[t0] is read exactly once per launch (plus the recomputed halo): do not allocate it in L1, whose
// capacity is what is left of the 228 KB after the shared-memory rings
__device__ __forceinline__ float2 gload2_stream(const float *p)
{
    float2 v;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
#ifndef OPESCI_SKEL_NOSTORE
#define OPESCI_SKEL_NOSTORE 0   /* diagnostics of the memory skeleton: drop one class of traffic */
#endif
#ifndef OPESCI_SKEL_NOT0
#define OPESCI_SKEL_NOT0 0
#endif
__device__ __forceinline__ void gstore(float *p, float v)
{
#if OPESCI_SKEL_NOSTORE
    if (v == 123.456f) *p = v;
    return;
#endif
#if OPESCI_STREAM_STORES
    __stcs(p, v);
#else
    *p = v;
#endif
}
__device__ __forceinline__ void gstore2(float *p, float a, float b)
{
#if OPESCI_SKEL_NOSTORE
    if (a == 123.456f) *p = b;
    return;
#endif
#if OPESCI_STREAM_STORES
    __stcs(reinterpret_cast<float2 *>(p), make_float2(a, b));
#else
    *reinterpret_cast<float2 *>(p) = make_float2(a, b);
#endif
}

// Each thread owns TWO z-adjacent points (lanes L = 0,1 at z0, z0+1): every shared/global access
// along y and x is one 8-byte access for both points and the z-windows of both come from three
// 8-byte loads, which halves the load/store-unit instruction count (the busiest pipe of the
// one-point-per-thread version, ncu: profiles/r01_fused_v2_ncu_summary.txt).
// The five stress fields that are published to the shared-memory ring anyway (Txy, Txz, Tyy, Tyz, Tzz) leave the SM as
// TMA tensor stores straight from their ring slot (one 64 x 12 box per field and plane, issued by one thread) instead of
// 5 x 512 STG.64: measured 19.9 -> 19.3 ms at 1024^3.  The box is the full tile width, so the two recomputed halo
// columns on either side are written as well -- with the bits the neighbouring tile writes there (same operands, same
// operation order).  Tiles whose box would reach outside the interior -- the first tile column, whose halo columns are
// the ghost cells z < m, and a ragged last column / row -- keep ordinary stores.
#ifndef OPESCI_TMA_STORE
#define OPESCI_TMA_STORE 1
#endif
struct StoreMaps { CUtensorMap m[5]; };   // Txy, Txz, Tyy, Tyz, Tzz (ring order), box = EZ x CY x 1
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tmap, const void *src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap), "r"(smem_u32(src)), "r"(c0),
                 "r"(c1), "r"(c2)
                 : "memory");
}
// ZF = true: the z-edge variant (tile columns 0 and last).  After the interior stress update of a plane it applies,
// in registers, what the reference's z-face stress loops do to the cells of that plane (opesci/staggeredgrid.py:750-813
// through opesci/fields.py:313-381, so = 4): Levander recompute of Txx, Tyy on the face plane from level t0, Tzz = 0 on
// it, antisymmetric mirrors of Tzz, Tyz, Txz into the ghost columns -- in the reference's own operation order -- and
// stores the ghost columns; the velocity update then also covers the z slabs of the shell [M, 2M+1) / [dim-2M-1, dim-M),
// whose stress operands are now final.  Restricted to the "pure" zone (planes [zf_xlo, zf_xhi), rows [M+1, dim2-M-1)):
// there no x-/y-face loop writes any cell these operations read or write, so the result is cell for cell what the
// separate face kernels produce; the thin remainder next to the x / y faces stays with them (Stepper::stress_bc).
// A separate instantiation, so the interior tiles keep their register budget.
// PAIR = true: the interior launch as 2-CTA clusters stacked in y (see FusedCfg::PCY).  After a CTA has published the
// stresses of a plane it pushes the M rows next to its partner -- Txy, Tyy, Tyz, the fields with y-windows -- into the
// partner's ring slot with st.async (distributed shared memory; the bytes count on the partner's `full` mbarrier of that
// slot), and the partner waits for them only M planes later, when its velocity update reads the slot.  A slot is
// refilled only after the partner has reported -- one remote arrival on this CTA's `empty` mbarrier -- that it has read the
// plane that lived there.  Same arithmetic on the same operands as the single-CTA tiles: bit-identical results; what
// changes is that 2M fewer halo rows per pair are loaded, recomputed and re-fetched.
template <int SO, int ARITH, bool HET = false, bool ZF = false, bool PAIR = false>
__global__ void __launch_bounds__(FusedCfg<SO / 2>::THREADS, OPESCI_FUSED_MINB)
fused_step(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmV,
           const __grid_constant__ CUtensorMap tmW, const FusedArgs A
#if OPESCI_TMA_STORE
           , const __grid_constant__ StoreMaps SMAPS
#endif
)
{
    constexpr int M = SO / 2;
    using K = FusedCfg<M>;
    typedef float T;
    constexpr int RD = K::RD;
    constexpr int VT = K::VTILE / 4;      // floats per velocity tile
    constexpr int ST = (PAIR ? K::PSTILE : K::STILE) / 4;      // floats per stress tile
    static_assert(!PAIR || !ZF, "pairs: not for the z-edge variant");
    // dynamic shared memory (no static __shared__ in this kernel, so the window starts 1024-B
    // aligned): [3][RD] velocity tiles | [5][SR] stress tiles | mbarriers
    extern __shared__ __align__(1024) unsigned char smem[];
    const T *vring = reinterpret_cast<const T *>(smem);
    T *sring = reinterpret_cast<T *>(smem + (size_t)3 * RD * K::VTILE);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)3 * RD * K::VTILE + (size_t)5 * K::SR * (PAIR ? K::PSTILE : K::STILE));
    uint64_t *pfull = bars + 3 * RD + 1, *pempty = pfull + K::SR;     // PAIR: halo rows arrived / slot read by the partner

#ifndef OPESCI_SPLIT_BARRIER
#define OPESCI_SPLIT_BARRIER 0
#endif
#ifndef OPESCI_SKELETON
#define OPESCI_SKELETON 0
#endif
#ifndef OPESCI_PACE
#define OPESCI_PACE 0   /* > 0: a tile never runs more than this many planes ahead of a running y-neighbour (A/B experiment) */
#endif
#ifndef OPESCI_PACE_MAX
#define OPESCI_PACE_MAX 48
#endif
#ifndef OPESCI_CLUSTER_Z
#define OPESCI_CLUSTER_Z 1   /* > 1: z-adjacent CTAs form a thread-block cluster and march in lockstep (cluster barrier per plane) */
#endif
    const GridGeom &G = A.G;
    const int tid = threadIdx.x;
    // Programmatic dependent launch: the interior launch that follows the z-edge launch in the stream does not depend on
    // it (disjoint cells, both read level t0 only) -- let its CTAs start as soon as SMs free up instead of after the last,
    // partly filled wave of z-edge CTAs.  (The interior kernel waits for the z-edge grid just before it exits, below.)
    if (ZF) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int bx = ZF ? A.bxs[blockIdx.x] : (int)blockIdx.x + A.bx0;       // tile column
    const int tz = 2 * (tid % (K::EZ / 2)), ty = tid / (K::EZ / 2);          // lane 0 sits at tz, lane 1 at tz+1
    uint32_t prank = 0;       // PAIR: 0 = upper CTA of the pair (smaller y), 1 = lower
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(prank));
    // global y of thread row 0: tiles advance by CY rows; a pair advances by 2 PCY rows and its lower CTA starts EY rows down
    const int tile_y0 = PAIR ? (int)(blockIdx.y >> 1) * (2 * K::PCY) + (int)prank * K::EY : (int)blockIdx.y * K::CY;
    const int row_first = PAIR ? (prank == 0 ? M : 0) : M;          // first stored thread row
    const int row_count = PAIR ? K::PCY : K::CY;                    // stored rows
    const int ring_row = PAIR ? (prank == 0 ? 0 : M) : 0;           // slot row of thread row 0 (the lower CTA keeps M rows above)
    const int ye = tile_y0 + ty, ze = bx * K::CZ + tz - K::ZS;   // global coords of lane 0 (ze < 0: outside)
    const int chunk = blockIdx.z + A.chunk0;
    const int xa = A.xs[chunk];
    const int xb = A.xs[chunk + 1];
    const int xs_begin = max(M, xa - M), xs_end = min(G.dim[0] - M, xb + M);
    // plane p of U lives in slot (p - pbaseU) % RD, of V/W in slot (p - pbaseVW) % RD; with the
    // x loop unrolled RD times every slot index below is a compile-time constant
    const int pbaseU = xs_begin - M, pbaseVW = xs_begin - M + 1;
    const int lastU = xs_end + M - 2, lastVW = xs_end + M - 1;
    const int c0 = bx * K::CZ - K::OFFZ, c1 = tile_y0 - M;   // TMA box origin (may be negative)
    const int lvl0 = A.t0 * G.dim[0];
    constexpr uint32_t TILE_BYTES = K::VZ * K::VY * 4;

    if (tid == 0) {
        for (int i = 0; i < 3 * RD; ++i) mbar_init(&bars[i], 1);
#if OPESCI_SPLIT_BARRIER
        mbar_init(&bars[3 * RD], K::THREADS / 32);   // step barrier: one arrival per warp and plane
#endif
        if (PAIR)
            for (int i = 0; i < K::SR; ++i) { mbar_init(&pfull[i], 1); mbar_init(&pempty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (PAIR) {
        // the partner's barriers exist before anything is pushed to it
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    // PAIR: rows this CTA pushes to its partner (the M stored rows next to it) and where they land in the partner's slot
    const bool push_row = PAIR && (prank == 0 ? ty >= K::EY - M : ty < M);
    const int push_dst_row = prank == 0 ? ty - (K::EY - M) : ty + K::EY;   // partner's slot row (its ring_row included)
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < RD; ++k) {
            if (pbaseU + k <= lastU) {
                mbar_arrive_expect_tx(&bars[0 * RD + k], TILE_BYTES);
                tma_load_3d((void *)(vring + (0 * RD + k) * VT), &tmU, &bars[0 * RD + k], c0, c1, lvl0 + pbaseU + k);
            }
            if (pbaseVW + k <= lastVW) {
                mbar_arrive_expect_tx(&bars[1 * RD + k], TILE_BYTES);
                tma_load_3d((void *)(vring + (1 * RD + k) * VT), &tmV, &bars[1 * RD + k], c0, c1, lvl0 + pbaseVW + k);
                mbar_arrive_expect_tx(&bars[2 * RD + k], TILE_BYTES);
                tma_load_3d((void *)(vring + (2 * RD + k) * VT), &tmW, &bars[2 * RD + k], c0, c1, lvl0 + pbaseVW + k);
            }
        }
    }

    // per-lane predicates
    bool inb[2], st_yz[2], vf_yz[2];
#pragma unroll
    for (int L = 0; L < 2; ++L) {
        const int z = ze + L, t = tz + L;
        const bool core = ty >= row_first && ty < row_first + row_count && t >= M && t < M + K::CZ;
        inb[L] = ye < G.dim[1] && z >= 0 && z < G.dim[2];
        st_yz[L] = core && ye >= M && ye < G.dim[1] - M && z >= M && z < G.dim[2] - M;
        // velocities: deep interior; the z-edge variant also owns the z slabs of the shell (whole interior z range)
        vf_yz[L] = core && ye >= 2 * M + 1 && ye < G.dim[1] - 2 * M - 1 &&
                   (ZF ? (z >= M && z < G.dim[2] - M) : (z >= 2 * M + 1 && z < G.dim[2] - 2 * M - 1));
    }
    // z-fold: rows whose z-face cells see no y-face loop (warp-uniform: a warp is one tile row)
    const bool zf_row = ZF && ye >= M + 1 && ye < G.dim[1] - M - 1;
    const bool inb2 = inb[0] && inb[1], st2 = st_yz[0] && st_yz[1], vf2 = vf_yz[0] && vf_yz[1];
#ifndef OPESCI_HALO_SKIP
#define OPESCI_HALO_SKIP 0   /* measured on B200: 20.85 vs 20.05 ms -- fewer loads and flops, but 113 instead of 106 registers and a warp-level branch per plane */
#endif
    // The recomputed halo ROWS of the tile (a whole warp each) feed only the y-windows of the core rows next to them:
    // Txy, Tyy, Tyz.  Their Txx, Tzz, Txz are never read by anyone, so those warps neither load nor update them.
    const bool halo_row = OPESCI_HALO_SKIP && (ty < M || ty >= M + K::CY);
    // CTA-uniform: tiles whose whole 64 x 12 store box lies inside the interior (not the first tile column, whose halo
    // columns are ghost cells; not a ragged last column / row)
    const bool tma_st = OPESCI_TMA_STORE && !OPESCI_SPLIT_BARRIER && !ZF && bx > 0 &&
                        bx * K::CZ - K::ZS + K::EZ <= G.dim[2] - M && tile_y0 + row_first + row_count <= G.dim[1] - M;
    const int xv_lo = max(xa, 2 * M + 1), xv_hi = min(xb, G.dim[0] - 2 * M - 1);
    const long long pyz = (long long)ye * G.s[1] + ze;
    const long long lv0 = (long long)A.t0 * G.level, lv1 = (long long)A.t1 * G.level;
    const long long sx = G.s[0];
    const T *__restrict__ gT0[6];
    T *__restrict__ gT1[6];
    T *__restrict__ gV1[3];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        gT0[k] = (const T *)A.F.f[F_TXX + k] + lv0 + pyz;
        gT1[k] = (T *)A.F.f[F_TXX + k] + lv1 + pyz;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) gV1[k] = (T *)A.F.f[F_U + k] + lv1 + pyz;

    // register state per lane: own-column x-windows of the new stresses
    T txx[2][2 * M], txy[2][2 * M + 1], txz[2][2 * M + 1];
#pragma unroll
    for (int L = 0; L < 2; ++L) {
#pragma unroll
        for (int k = 0; k < 2 * M; ++k) txx[L][k] = 0;
#pragma unroll
        for (int k = 0; k < 2 * M + 1; ++k) txy[L][k] = txz[L][k] = 0;
    }
    T vself[2] = {0, 0}, wself[2] = {0, 0};   // V,W[t0] at plane xs-M (saved one iteration earlier)
    const int lo = (ty + M) * K::VZ + tz + K::OFFZ - K::ZS;   // lane 0's element inside a velocity tile (even => 8-B aligned)
    const T *vlo = vring + lo;
    T *slo = sring + (ty + ring_row) * K::EZ + tz;

#ifndef OPESCI_T0_BAND_POLICY
#define OPESCI_T0_BAND_POLICY 0
#endif
#if OPESCI_T0_BAND_POLICY
    // Rows / columns within 2M of a tile edge are read twice (once as this tile's data, once as the neighbour's
    // recomputed halo): keep those in L2 (evict_last) and let the private middle of the tile go first (evict_first).
    const bool band = ty < 2 * M || ty >= K::EY - 2 * M || tz < 2 * M || tz >= K::EZ - 2 * M;
    uint64_t pol_last, pol_first;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_last));
#if OPESCI_T0_BAND_POLICY == 2
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_first));
#else
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_first));
#endif
    const uint64_t pol = band ? pol_last : pol_first;
#endif
    auto load_told = [&](T (*told)[6], long long px) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            if (halo_row && (k == 0 || k == 2 || k == 5)) { told[0][k] = told[1][k] = 0; continue; }
#if OPESCI_SKEL_NOT0
            told[0][k] = told[1][k] = (T)k; continue;
#endif
            if (inb2) {
#if OPESCI_T0_BAND_POLICY
                float2 v;
                asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;" : "=f"(v.x), "=f"(v.y) : "l"(gT0[k] + px), "l"(pol));
#elif OPESCI_T0_NOALLOC
                const float2 v = gload2_stream(gT0[k] + px);
#else
                const float2 v = *reinterpret_cast<const float2 *>(gT0[k] + px);
#endif
                told[0][k] = v.x; told[1][k] = v.y;
            } else {
                told[0][k] = inb[0] ? gT0[k][px] : (T)0;
                told[1][k] = inb[1] ? gT0[k][px + 1] : (T)0;
            }
        }
    };
#ifndef OPESCI_T0_AHEAD
#define OPESCI_T0_AHEAD 1
#endif
    // T[t0] of the planes ahead: AHEAD = 1 keeps one plane in registers, AHEAD = 2 two (buffers alternate with the
    // parity of the unrolled plane index r; RD is even)
    T told_buf[OPESCI_T0_AHEAD][2][6];
    long long px = (long long)xs_begin * sx;
    load_told(told_buf[0], px);
    // ZF, high face: Tzz two columns beyond the face plane (z = dim3-1) is written by no loop, but the W update of the face
    // plane reads it: the window must hold what the array holds (level t1).  Fetched a plane ahead like T[t0].
    T zz_far = 0;
    const int zf_far_lane = ZF ? ((A.zf_side[1] && bx == A.zf_bx[1]) ? ((tz == A.zf_c[1] + 2) ? 0 : (tz + 1 == A.zf_c[1] + 2) ? 1 : -1) : -1) : -1;
    const bool zf_far = zf_far_lane >= 0 && ze + zf_far_lane < G.dim[2] && ye < G.dim[1];
    if (ZF && zf_far) zz_far = gT1[2][px + zf_far_lane];
    // ZF: one z face per tile column (the host requires at least two tile columns).  Everything that does not change from
    // plane to plane is worked out once: the face side of this column, the source columns of the mirrors (CTA-uniform)
    // and, per thread, which of its two columns takes which part -- `zrole`, 4 bits per column:
    //   1 face plane b (Levander recompute of Txx, Tyy; Tzz = 0)      2 Tzz mirror target b + dir
    //   4 / 8 first / second shear mirror target
    const int zside = (ZF && A.zf_side[1] && bx == A.zf_bx[1]) ? 1 : 0;
    const bool zf_tile = ZF && A.zf_side[zside] && bx == A.zf_bx[zside];
    const int zc = A.zf_c[zside], zdir = zside == 0 ? -1 : 1;
    const int zs_zz = zc - zdir;
    const int zd0 = zside == 0 ? zc - 1 : zc, zs0 = zside == 0 ? zc : zc - 1;
    const int zd1 = zside == 0 ? zc - 2 : zc + 1, zs1 = zside == 0 ? zc + 1 : zc - 2;
    unsigned zrole = 0;
    if (ZF && zf_tile && zf_row) {
#pragma unroll
        for (int L = 0; L < 2; ++L) {
            const int j = tz + L;
            zrole |= (unsigned)((j == zc ? 1 : 0) | (j == zc + zdir ? 2 : 0) | (j == zd0 ? 4 : 0) | (j == zd1 ? 8 : 0)) << (4 * L);
        }
    }
    const bool zrow_stores = ty >= M && ty < M + K::CY;
#if OPESCI_T0_AHEAD == 2
    static_assert(RD % 2 == 0, "plane parity must survive the unrolled loop");
    if (xs_begin + 1 < xs_end) load_told(told_buf[1], px + sx);
#endif

    // planes of the first window (first use of their slots: phase parity 0)
#pragma unroll
    for (int k = 0; k < 2 * M - 1; ++k) {
        mbar_wait(&bars[0 * RD + k], 0);
        mbar_wait(&bars[1 * RD + k], 0);
        mbar_wait(&bars[2 * RD + k], 0);
    }

    for (int xs0 = xs_begin, q = 0; xs0 < xs_end; xs0 += RD, ++q) {
#pragma unroll
        for (int r = 0; r < RD; ++r) {
            const int xs = xs0 + r;
            if (xs >= xs_end) break;
            T (*told)[6] = told_buf[r % OPESCI_T0_AHEAD];
#if OPESCI_PACE > 0
            if (r == 0 && A.pace) {
                if (tid == 0) {
                    volatile int *pg = A.pace + ((size_t)chunk * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
                    *pg = xs;
                    // neighbours that have not started (-1) or have finished (INT_MAX) never hold anybody back, and the
                    // tile that is furthest behind waits for nobody: no deadlock
                    for (int dy = -1; dy <= 1; dy += 2) {
                        const int by = (int)blockIdx.y + dy;
                        if (by < 0 || by >= (int)gridDim.y) continue;
                        volatile int *pn = pg + dy * (int)gridDim.x;
                        int v = *pn;
                        // (a neighbour more than OPESCI_PACE_MAX planes behind belongs to a later wave: waiting for it
                        // would stall this tile for most of its run)
                        while (v >= 0 && v < xs - OPESCI_PACE && v >= xs - OPESCI_PACE_MAX) { __nanosleep(100); v = *pn; }
                    }
                }
                __syncthreads();
            }
#endif
            // heterogeneous mode: lambda, mu, mu12, mu23, mu13 of the two cells (plane xs), issued before the waits so
            // that their latency overlaps the TMA wait and the operand gather
            T med[2][5];
            if (HET) {
                const int ids[5] = {OPESCI_MEDIA_LAMBDA, OPESCI_MEDIA_MU, OPESCI_MEDIA_MU12, OPESCI_MEDIA_MU23, OPESCI_MEDIA_MU13};
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    const float *mp = A.MD.m[ids[k]] + pyz + px;
                    if (inb2) { const float2 v = *reinterpret_cast<const float2 *>(mp); med[0][k] = v.x; med[1][k] = v.y; }
                    else { med[0][k] = inb[0] ? mp[0] : 0.f; med[1][k] = inb[1] ? mp[1] : 0.f; }
                }
            }
            // ---- the newest planes of this iteration: relative plane index RD*q + r + 2M-1
            {
                const int slot = (r + 2 * M - 1) % RD;
                const uint32_t par = (uint32_t)(q + (r + 2 * M - 1) / RD) & 1u;
                mbar_wait(&bars[0 * RD + slot], par);
                mbar_wait(&bars[1 * RD + slot], par);
                mbar_wait(&bars[2 * RD + slot], par);
            }
            // ---- gather the operands of the six stress updates (all offsets are immediates)
            T ux[2][2 * M], vx[2][2 * M], wx[2][2 * M];   // x-windows: U bwd (xs-M..xs+M-1), V,W fwd (xs-M+1..xs+M)
#pragma unroll
            for (int j = 0; j < 2 * M; ++j) {
                const int slot = (r + j) % RD;
                const float2 a = *reinterpret_cast<const float2 *>(vlo + (0 * RD + slot) * VT);
                const float2 b = *reinterpret_cast<const float2 *>(vlo + (1 * RD + slot) * VT);
                const float2 c = *reinterpret_cast<const float2 *>(vlo + (2 * RD + slot) * VT);
                ux[0][j] = a.x; ux[1][j] = a.y;
                vx[0][j] = b.x; vx[1][j] = b.y;
                wx[0][j] = c.x; wx[1][j] = c.y;
            }
            // plane xs: U is window entry M (slot r+M), V/W window entry M-1 (slot r+M-1)
            const T *pu = vlo + (0 * RD + (r + M) % RD) * VT;
            const T *pv = vlo + (1 * RD + (r + M - 1) % RD) * VT;
            const T *pw = vlo + (2 * RD + (r + M - 1) % RD) * VT;
            T vy_b[2][2 * M], uy_f[2][2 * M], wy_f[2][2 * M];   // y-windows (bwd for the normal, fwd for the shear stresses)
#pragma unroll
            for (int j = 0; j < 2 * M; ++j) {
                const float2 a = *reinterpret_cast<const float2 *>(pv + (j - M) * K::VZ);
                const float2 b = *reinterpret_cast<const float2 *>(pu + (j - M + 1) * K::VZ);
                const float2 c = *reinterpret_cast<const float2 *>(pw + (j - M + 1) * K::VZ);
                vy_b[0][j] = a.x; vy_b[1][j] = a.y;
                uy_f[0][j] = b.x; uy_f[1][j] = b.y;
                wy_f[0][j] = c.x; wy_f[1][j] = c.y;
            }
            T fu[6], fv[6], fw[6];                // z rows z0-2 .. z0+3 of plane xs
            load6(pu, fu);
            load6(pv, fv);
            load6(pw, fw);
            T tn[2][6], uself[2], vself_next[2], wself_next[2];
#pragma unroll
            for (int L = 0; L < 2; ++L) {
                T wz_b[2 * M], uz_f[2 * M], vz_f[2 * M];
#pragma unroll
                for (int j = 0; j < 2 * M; ++j) {
                    wz_b[j] = fw[2 + L - M + j];
                    uz_f[j] = fu[2 + L - M + 1 + j];
                    vz_f[j] = fv[2 + L - M + 1 + j];
                }
                uself[L] = ux[L][0];               // U[t0] at plane xs-M (velocity self term of this iteration)
                vself_next[L] = vx[L][0];          // V,W[t0] at plane xs-M+1
                wself_next[L] = wx[L][0];
                if (HET && ARITH == OPESCI_ARITH_REFERENCE) {
                    const float lam = med[L][0], mu = med[L][1];
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        if (halo_row && a != 1) { tn[L][a] = 0; continue; }
                        T acc = told[L][a];
                        bool first = false;
                        if (a == 0) window_ref_arr_h<M, false, 2>(acc, first, ux[L], A.HC.c[0], A.HC.c2[0], lam, mu);
                        else window_ref_arr_h<M, false, 1>(acc, first, ux[L], A.HC.c[0], A.HC.c2[0], lam, mu);
                        if (a == 1) window_ref_arr_h<M, false, 2>(acc, first, vy_b[L], A.HC.c[1], A.HC.c2[1], lam, mu);
                        else window_ref_arr_h<M, false, 1>(acc, first, vy_b[L], A.HC.c[1], A.HC.c2[1], lam, mu);
                        if (a == 2) window_ref_arr_h<M, false, 2>(acc, first, wz_b, A.HC.c[2], A.HC.c2[2], lam, mu);
                        else window_ref_arr_h<M, false, 1>(acc, first, wz_b, A.HC.c[2], A.HC.c2[2], lam, mu);
                        tn[L][a] = acc;
                    }
                    {
                        T acc = told[L][3]; bool first = false;   // Txy (mu12): D_y U, D_x V
                        window_ref_arr_h<M, true, 1>(acc, first, uy_f[L], A.HC.c[1], A.HC.c2[1], med[L][2], 0.f);
                        window_ref_arr_h<M, true, 1>(acc, first, vx[L], A.HC.c[0], A.HC.c2[0], med[L][2], 0.f);
                        tn[L][3] = acc;
                    }
                    {
                        T acc = told[L][4]; bool first = false;   // Tyz (mu23): D_z V, D_y W
                        window_ref_arr_h<M, true, 1>(acc, first, vz_f, A.HC.c[2], A.HC.c2[2], med[L][3], 0.f);
                        window_ref_arr_h<M, true, 1>(acc, first, wy_f[L], A.HC.c[1], A.HC.c2[1], med[L][3], 0.f);
                        tn[L][4] = acc;
                    }
                    if (halo_row) tn[L][5] = 0;
                    else {
                        T acc = told[L][5]; bool first = false;   // Txz (mu13): D_z U, D_x W
                        window_ref_arr_h<M, true, 1>(acc, first, uz_f, A.HC.c[2], A.HC.c2[2], med[L][4], 0.f);
                        window_ref_arr_h<M, true, 1>(acc, first, wx[L], A.HC.c[0], A.HC.c2[0], med[L][4], 0.f);
                        tn[L][5] = acc;
                    }
                } else if (HET) {
                    const T du = window_fast_arr<M, T, false>(ux[L], A.HC.c[0]), dv = window_fast_arr<M, T, false>(vy_b[L], A.HC.c[1]),
                            dw = window_fast_arr<M, T, false>(wz_b, A.HC.c[2]);
                    const T tr = med[L][0] * (du + dv + dw), mu2 = 2.0f * med[L][1];
                    tn[L][0] = halo_row ? 0.f : told[L][0] + (tr + mu2 * du);
                    tn[L][1] = told[L][1] + (tr + mu2 * dv);
                    tn[L][2] = halo_row ? 0.f : told[L][2] + (tr + mu2 * dw);
                    tn[L][3] = told[L][3] + med[L][2] * (window_fast_arr<M, T, true>(uy_f[L], A.HC.c[1]) + window_fast_arr<M, T, true>(vx[L], A.HC.c[0]));
                    tn[L][4] = told[L][4] + med[L][3] * (window_fast_arr<M, T, true>(vz_f, A.HC.c[2]) + window_fast_arr<M, T, true>(wy_f[L], A.HC.c[1]));
                    tn[L][5] = halo_row ? 0.f : told[L][5] + med[L][4] * (window_fast_arr<M, T, true>(uz_f, A.HC.c[2]) + window_fast_arr<M, T, true>(wx[L], A.HC.c[0]));
                } else if (ARITH == OPESCI_ARITH_REFERENCE) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        if (halo_row && a != 1) { tn[L][a] = 0; continue; }
                        T acc = told[L][a];
                        bool first = false;
                        window_ref_arr<M, T, false>(acc, first, ux[L], A.C.sn[a][0]);
                        window_ref_arr<M, T, false>(acc, first, vy_b[L], A.C.sn[a][1]);
                        window_ref_arr<M, T, false>(acc, first, wz_b, A.C.sn[a][2]);
                        tn[L][a] = acc;
                    }
                    {
                        T acc = told[L][3]; bool first = false;   // Txy: D_y U, D_x V
                        window_ref_arr<M, T, true>(acc, first, uy_f[L], A.C.ss[0][0]);
                        window_ref_arr<M, T, true>(acc, first, vx[L], A.C.ss[0][1]);
                        tn[L][3] = acc;
                    }
                    {
                        T acc = told[L][4]; bool first = false;   // Tyz: D_z V, D_y W
                        window_ref_arr<M, T, true>(acc, first, vz_f, A.C.ss[1][0]);
                        window_ref_arr<M, T, true>(acc, first, wy_f[L], A.C.ss[1][1]);
                        tn[L][4] = acc;
                    }
                    if (halo_row) tn[L][5] = 0;
                    else {
                        T acc = told[L][5]; bool first = false;   // Txz: D_z U, D_x W
                        window_ref_arr<M, T, true>(acc, first, uz_f, A.C.ss[2][0]);
                        window_ref_arr<M, T, true>(acc, first, wx[L], A.C.ss[2][1]);
                        tn[L][5] = acc;
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        if (halo_row && a != 1) { tn[L][a] = 0; continue; }
                        tn[L][a] = told[L][a] + (window_fast_arr<M, T, false>(ux[L], A.C.sn[a][0]) +
                                                 window_fast_arr<M, T, false>(vy_b[L], A.C.sn[a][1]) +
                                                 window_fast_arr<M, T, false>(wz_b, A.C.sn[a][2]));
                    }
                    tn[L][3] = told[L][3] + (window_fast_arr<M, T, true>(uy_f[L], A.C.ss[0][0]) + window_fast_arr<M, T, true>(vx[L], A.C.ss[0][1]));
                    tn[L][4] = told[L][4] + (window_fast_arr<M, T, true>(vz_f, A.C.ss[1][0]) + window_fast_arr<M, T, true>(wy_f[L], A.C.ss[1][1]));
                    tn[L][5] = halo_row ? (T)0 : told[L][5] + (window_fast_arr<M, T, true>(uz_f, A.C.ss[2][0]) + window_fast_arr<M, T, true>(wx[L], A.C.ss[2][1]));
                }
            }
#if OPESCI_SKELETON
            // DIAGNOSTIC ONLY (wrong results): keep every global load / store, TMA transfer, barrier and the shared-memory
            // publish, drop the operand gathers and the arithmetic -- how long does the memory skeleton of the kernel take?
#pragma unroll
            for (int L = 0; L < 2; ++L)
#pragma unroll
                for (int k = 0; k < 6; ++k) tn[L][k] = told[L][k] + ux[L][0];
#endif
            // ZF: on the planes / rows of FusedArgs::zf_raw_x / zf_raw_y the plain interior Txx / Tyy go to global memory
            // (stored here, before the z-face operations below touch the registers)
            bool early[2] = {false, false};
            if constexpr (ZF) {
                early[0] = xs == A.zf_raw_x[0] || xs == A.zf_raw_x[1];
                early[1] = ye == A.zf_raw_y[0] || ye == A.zf_raw_y[1];
                if (xs >= xa && xs < xb) {
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (early[e]) {
                            if (st_yz[0]) gstore(gT1[e] + px, tn[0][e]);
                            if (st_yz[1]) gstore(gT1[e] + px + 1, tn[1][e]);
                        }
                }
            }
            if constexpr (ZF) {
                // ---- z-face stress ghost loops of this plane, in registers (see the kernel's header comment).
                // tn[L][k]: 0 Txx, 1 Tyy, 2 Tzz, 3 Txy, 4 Tyz, 5 Txz.  A warp is one tile row, so every source
                // column lives in this warp: shuffles, executed by all lanes (warp-uniform branch).
                if (zf_tile && zf_row && xs >= A.zf_xlo && xs < A.zf_xhi) {
                    // Tzz[b + dir] = -Tzz[b - dir]; shear: low T[b-1] = -T[b], T[b-2] = -T[b+1]; high T[b'] = -T[b'-1],
                    // T[b'+1] = -T[b'-2] (opesci/fields.py:355-381)
                    const T zz = __shfl_sync(0xffffffffu, (zs_zz & 1) ? tn[1][2] : tn[0][2], zs_zz >> 1);
                    const T yz0 = __shfl_sync(0xffffffffu, (zs0 & 1) ? tn[1][4] : tn[0][4], zs0 >> 1);
                    const T xz0 = __shfl_sync(0xffffffffu, (zs0 & 1) ? tn[1][5] : tn[0][5], zs0 >> 1);
                    const T yz1 = __shfl_sync(0xffffffffu, (zs1 & 1) ? tn[1][4] : tn[0][4], zs1 >> 1);
                    const T xz1 = __shfl_sync(0xffffffffu, (zs1 & 1) ? tn[1][5] : tn[0][5], zs1 >> 1);
                    if (zrole != 0) {      // the two or three lanes of the row that hold a face / ghost column
                        const bool own = xs >= xa && xs < xb && zrow_stores;
#pragma unroll
                        for (int L = 0; L < 2; ++L) {
                            const unsigned r4 = (zrole >> (4 * L)) & 15u;
                            if (r4 & 1u) {
                                // Levander: T_ee[t1] = T_ee[t0] + (bwd window of U along x) + (bwd window of V along y), emitted
                                // order +1, -1, -2, 0 per window, separate multiply and add (face_batch evaluates the same
                                // term table the same way in either arithmetic mode)
#pragma unroll
                                for (int e = 0; e < 2; ++e) {
                                    T acc = told[L][e];
                                    const float *cu = A.zf_lev[e][0], *cv = A.zf_lev[e][1];
                                    acc = __fadd_rn(acc, __fmul_rn(cu[1], ux[L][3]));
                                    acc = __fadd_rn(acc, __fmul_rn(-cu[0], ux[L][1]));
                                    acc = __fadd_rn(acc, __fmul_rn(-cu[1], ux[L][0]));
                                    acc = __fadd_rn(acc, __fmul_rn(cu[0], ux[L][2]));
                                    acc = __fadd_rn(acc, __fmul_rn(cv[1], vy_b[L][3]));
                                    acc = __fadd_rn(acc, __fmul_rn(-cv[0], vy_b[L][1]));
                                    acc = __fadd_rn(acc, __fmul_rn(-cv[1], vy_b[L][0]));
                                    acc = __fadd_rn(acc, __fmul_rn(cv[0], vy_b[L][2]));
                                    tn[L][e] = acc;
                                }
                                tn[L][2] = (T)0;
                            }
                            if (r4 & 2u) tn[L][2] = -zz;
                            if (r4 & 4u) { tn[L][4] = -yz0; tn[L][5] = -xz0; }
                            if (r4 & 8u) { tn[L][4] = -yz1; tn[L][5] = -xz1; }
                            // the ghost columns this plane's loops wrote (owned planes only; the interior columns go out below)
                            if (own) {
                                if (r4 & 2u) gstore(gT1[2] + px + L, tn[L][2]);
                                if ((r4 & 8u) || (zside == 0 && (r4 & 4u))) {
                                    gstore(gT1[4] + px + L, tn[L][4]);
                                    gstore(gT1[5] + px + L, tn[L][5]);
                                }
                            }
                        }
                    }
                    // Tzz two columns beyond the high face (z = dim3-1) is written by no loop: the W update of the face plane
                    // reads it, so the window must hold what the array holds
                    if (zf_far) { if (zf_far_lane == 0) tn[0][2] = zz_far; else tn[1][2] = zz_far; }
                }
            }
            // ---- store the new stresses (owned tile, owned planes), prefetch next T[t0]
            if (xs >= xa && xs < xb) {
                if (st2) {
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        if (ZF && k < 2 && early[k]) continue;
                        if (k == 0 || !tma_st) gstore2(gT1[k] + px, tn[0][k], tn[1][k]);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        if (k != 0 && tma_st) continue;
                        if (ZF && k < 2 && early[k]) continue;
                        if (st_yz[0]) gstore(gT1[k] + px, tn[0][k]);
                        if (st_yz[1]) gstore(gT1[k] + px + 1, tn[1][k]);
                    }
                }
            }
            px += sx;
            if (xs + OPESCI_T0_AHEAD < xs_end) load_told(told, px + (OPESCI_T0_AHEAD - 1) * sx);
            if (ZF && zf_far && xs + 1 < xs_end) zz_far = gT1[2][px + zf_far_lane];
            // ---- shift the register windows, publish the in-plane operands
#pragma unroll
            for (int L = 0; L < 2; ++L) {
#pragma unroll
                for (int k = 0; k < 2 * M - 1; ++k) txx[L][k] = txx[L][k + 1];
                txx[L][2 * M - 1] = tn[L][0];
#pragma unroll
                for (int k = 0; k < 2 * M; ++k) { txy[L][k] = txy[L][k + 1]; txz[L][k] = txz[L][k + 1]; }
                txy[L][2 * M] = tn[L][3];
                txz[L][2 * M] = tn[L][5];
            }
#if OPESCI_SPLIT_BARRIER
            // Split-phase step barrier instead of a rendezvous: a warp ARRIVES once it has gathered its operands of plane
            // xs and published its stresses, and only WAITS -- here, one plane later -- for everybody's arrival of the
            // previous plane.  That covers all three hazards with a plane of slack: the slot written below was last read
            // by the velocity phase two planes ago (before that warp's previous arrival); the slots the velocity phase
            // reads were published two planes ago; and the TMA refill of the previous plane's dead slot (thread 0,
            // below) follows every warp's gather of that plane.  Warps may drift up to one plane apart, so the
            // shared-memory phase of one overlaps the arithmetic of another.
            if (xs > xs_begin) mbar_wait(&bars[3 * RD], (uint32_t)((r + 1) & 1));
#endif
            {
                T *s = slo + (xs & (K::SR - 1)) * ST;
                *reinterpret_cast<float2 *>(s + 0 * K::SR * ST) = make_float2(tn[0][3], tn[1][3]);   // Txy
                *reinterpret_cast<float2 *>(s + 1 * K::SR * ST) = make_float2(tn[0][5], tn[1][5]);   // Txz
                *reinterpret_cast<float2 *>(s + 2 * K::SR * ST) = make_float2(tn[0][1], tn[1][1]);   // Tyy
                *reinterpret_cast<float2 *>(s + 3 * K::SR * ST) = make_float2(tn[0][4], tn[1][4]);   // Tyz
                *reinterpret_cast<float2 *>(s + 4 * K::SR * ST) = make_float2(tn[0][2], tn[1][2]);   // Tzz
            }
            if constexpr (PAIR) {
                if (push_row) {
                    const int slot = xs & (K::SR - 1);
                    // the partner has read the plane that lived in this slot (SR planes ago)?
                    if (xs - K::SR >= xs_begin) mbar_wait(&pempty[slot], (uint32_t)(((xs - K::SR - xs_begin) / K::SR) & 1));
                    // Txy, Tyy, Tyz of this row -> same columns of the partner's slot row; 8 bytes each, counted on its `full` barrier
                    const uint32_t dst0 = smem_u32(sring + slot * ST + push_dst_row * K::EZ + tz);
                    uint32_t rdst, rbar;
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rdst) : "r"(dst0), "r"(prank ^ 1u));
                    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(&pfull[slot])), "r"(prank ^ 1u));
                    const int kf[3] = {0, 2, 3};      // ring order: Txy, Txz, Tyy, Tyz, Tzz
                    const int kt[3] = {3, 1, 4};      // tn index of Txy, Tyy, Tyz
#pragma unroll
                    for (int i = 0; i < 3; ++i)
                        asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(
                                         rdst + (uint32_t)(kf[i] * K::SR * ST * 4)),
                                     "r"(__float_as_uint(tn[0][kt[i]])), "r"(__float_as_uint(tn[1][kt[i]])), "r"(rbar)
                                     : "memory");
                }
            }
            T bet[2][3];   // heterogeneous mode: beta1, beta2, beta3 of the two cells (plane xv = xs - M), issued before the barrier
            if (HET) {
                const bool vel_here = (vf_yz[0] || vf_yz[1]) && xs - M >= xv_lo && xs - M < xv_hi;
                const long long pxv_ = px - (long long)(M + 1) * sx;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float2 v = make_float2(0.f, 0.f);
                    if (vel_here) v = *reinterpret_cast<const float2 *>(A.MD.m[OPESCI_MEDIA_BETA1 + k] + pyz + pxv_);
                    bet[0][k] = v.x; bet[1][k] = v.y;
                }
            }
#if OPESCI_SPLIT_BARRIER
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&bars[3 * RD]);
            // ---- the oldest planes of the PREVIOUS iteration (slot r-1) are dead for every warp: refill their slots
            if (tid == 0 && xs > xs_begin) {
                const int rp = (r + RD - 1) % RD;
                const int pU = xs - 1 - M + RD, pVW = xs - M + RD;
#else
#if OPESCI_TMA_STORE
            if (tma_st) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the publish above -> visible to the TMA engine
                // the slot published NEXT plane was stored from three planes ago: that store must have read it before
                // anybody passes the barrier below
                if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            }
#endif
            __syncthreads();
            if constexpr (PAIR) {
                if (tid == 0) {
                    // this plane's halo rows: 3 fields x M rows x EZ columns from the partner
                    mbar_arrive_expect_tx(&pfull[xs & (K::SR - 1)], 3u * M * K::EZ * 4u);
                    // every warp is past the velocity update of the previous iteration (plane xs-1-M): its slot may be refilled
                    const int done = xs - 1 - M;
                    if (done >= xs_begin) {
                        uint32_t rbar;
                        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(&pempty[done & (K::SR - 1)])), "r"(prank ^ 1u));
                        asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(rbar) : "memory");
                    }
                }
            }
#if OPESCI_TMA_STORE
            if (tma_st && tid == 0 && xs >= xa && xs < xb) {
                const T *s0 = sring + (xs & (K::SR - 1)) * ST + (row_first + ring_row) * K::EZ;      // first stored row of the slot
#pragma unroll
                for (int k = 0; k < 5; ++k)
                    tma_store_3d(&SMAPS.m[k], s0 + k * K::SR * ST, bx * K::CZ - K::ZS, tile_y0 + row_first,
                                 A.t1 * G.dim[0] + xs);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
#endif
            // ---- the oldest planes (window entry 0, slot r) are dead: refill their slots
            if (tid == 0) {
                const int rp = r;
                const int pU = xs - M + RD, pVW = xs - M + 1 + RD;
#endif
                if (pU <= lastU) {
                    mbar_arrive_expect_tx(&bars[0 * RD + rp], TILE_BYTES);
                    tma_load_3d((void *)(vring + (0 * RD + rp) * VT), &tmU, &bars[0 * RD + rp], c0, c1, lvl0 + pU);
                }
                if (pVW <= lastVW) {
                    mbar_arrive_expect_tx(&bars[1 * RD + rp], TILE_BYTES);
                    tma_load_3d((void *)(vring + (1 * RD + rp) * VT), &tmV, &bars[1 * RD + rp], c0, c1, lvl0 + pVW);
                    mbar_arrive_expect_tx(&bars[2 * RD + rp], TILE_BYTES);
                    tma_load_3d((void *)(vring + (2 * RD + rp) * VT), &tmW, &bars[2 * RD + rp], c0, c1, lvl0 + pVW);
                }
            }
#if OPESCI_CLUSTER_Z > 1
            // lockstep with the z-neighbours of the cluster: their row pieces of the same plane reach L2 / DRAM together
            if (A.cluster_sync) {
                asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
                asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            }
#endif
            // ---- velocities of plane xv = xs - M from the new stresses xv-M .. xv+M
            const int xv = xs - M;
            if constexpr (PAIR) {
                // the partner's rows of plane xv (pushed M planes ago) are in the slot
                if (xv >= xs_begin) mbar_wait(&pfull[xv & (K::SR - 1)], (uint32_t)(((xv - xs_begin) / K::SR) & 1));
            }
            if ((vf_yz[0] || vf_yz[1]) && xv >= xv_lo && xv < xv_hi) {
                const T *s = slo + (xv & (K::SR - 1)) * ST;
                const T *sxy = s, *sxz = s + 1 * K::SR * ST, *syy = s + 2 * K::SR * ST, *syz = s + 3 * K::SR * ST,
                        *szz = s + 4 * K::SR * ST;
                T xy_yb[2][2 * M], yy_yf[2][2 * M], yz_yb[2][2 * M];
#pragma unroll
                for (int j = 0; j < 2 * M; ++j) {
                    const float2 a = *reinterpret_cast<const float2 *>(sxy + (j - M) * K::EZ);
                    const float2 b = *reinterpret_cast<const float2 *>(syy + (j - M + 1) * K::EZ);
                    const float2 c = *reinterpret_cast<const float2 *>(syz + (j - M) * K::EZ);
                    xy_yb[0][j] = a.x; xy_yb[1][j] = a.y;
                    yy_yf[0][j] = b.x; yy_yf[1][j] = b.y;
                    yz_yb[0][j] = c.x; yz_yb[1][j] = c.y;
                }
                T fxz[6], fyz[6], fzz[6];
                load6(sxz, fxz);
                load6(syz, fyz);
                load6(szz, fzz);
                T vout[2][3];
#pragma unroll
                for (int L = 0; L < 2; ++L) {
                    T xz_zb[2 * M], yz_zb[2 * M], zz_zf[2 * M];
#pragma unroll
                    for (int j = 0; j < 2 * M; ++j) {
                        xz_zb[j] = fxz[2 + L - M + j];
                        yz_zb[j] = fyz[2 + L - M + j];
                        zz_zf[j] = fzz[2 + L - M + 1 + j];
                    }
                    // x-windows from registers: Txx fwd = planes xv-M+1..xv+M = txx[0..2M-1];
                    // Txy, Txz bwd = planes xv-M..xv+M-1 = txy[0..2M-1]
                    if (HET && ARITH == OPESCI_ARITH_REFERENCE) {
                        T acc = 0; bool first = true;
                        window_ref_arr_h<M, true, 1>(acc, first, txx[L], A.HC.c[0], A.HC.c2[0], bet[L][0], 0.f);
                        window_ref_arr_h<M, false, 1>(acc, first, xy_yb[L], A.HC.c[1], A.HC.c2[1], bet[L][0], 0.f);
                        window_ref_arr_h<M, false, 1>(acc, first, xz_zb, A.HC.c[2], A.HC.c2[2], bet[L][0], 0.f);
                        vout[L][0] = add_rn<T>(acc, uself[L]);
                        acc = 0; first = true;
                        window_ref_arr_h<M, false, 1>(acc, first, txy[L], A.HC.c[0], A.HC.c2[0], bet[L][1], 0.f);
                        window_ref_arr_h<M, true, 1>(acc, first, yy_yf[L], A.HC.c[1], A.HC.c2[1], bet[L][1], 0.f);
                        window_ref_arr_h<M, false, 1>(acc, first, yz_zb, A.HC.c[2], A.HC.c2[2], bet[L][1], 0.f);
                        vout[L][1] = add_rn<T>(acc, vself[L]);
                        acc = 0; first = true;
                        window_ref_arr_h<M, false, 1>(acc, first, txz[L], A.HC.c[0], A.HC.c2[0], bet[L][2], 0.f);
                        window_ref_arr_h<M, false, 1>(acc, first, yz_yb[L], A.HC.c[1], A.HC.c2[1], bet[L][2], 0.f);
                        window_ref_arr_h<M, true, 1>(acc, first, zz_zf, A.HC.c[2], A.HC.c2[2], bet[L][2], 0.f);
                        vout[L][2] = add_rn<T>(acc, wself[L]);
                    } else if (HET) {
                        vout[L][0] = uself[L] + bet[L][0] * (window_fast_arr<M, T, true>(txx[L], A.HC.c[0]) + window_fast_arr<M, T, false>(xy_yb[L], A.HC.c[1]) +
                                                             window_fast_arr<M, T, false>(xz_zb, A.HC.c[2]));
                        vout[L][1] = vself[L] + bet[L][1] * (window_fast_arr<M, T, false>(txy[L], A.HC.c[0]) + window_fast_arr<M, T, true>(yy_yf[L], A.HC.c[1]) +
                                                             window_fast_arr<M, T, false>(yz_zb, A.HC.c[2]));
                        vout[L][2] = wself[L] + bet[L][2] * (window_fast_arr<M, T, false>(txz[L], A.HC.c[0]) + window_fast_arr<M, T, false>(yz_yb[L], A.HC.c[1]) +
                                                             window_fast_arr<M, T, true>(zz_zf, A.HC.c[2]));
                    } else if (ARITH == OPESCI_ARITH_REFERENCE) {
                        T acc = 0; bool first = true;
                        window_ref_arr<M, T, true>(acc, first, txx[L], A.C.v[0][0]);
                        window_ref_arr<M, T, false>(acc, first, xy_yb[L], A.C.v[0][1]);
                        window_ref_arr<M, T, false>(acc, first, xz_zb, A.C.v[0][2]);
                        vout[L][0] = add_rn<T>(acc, uself[L]);
                        acc = 0; first = true;
                        window_ref_arr<M, T, false>(acc, first, txy[L], A.C.v[1][0]);
                        window_ref_arr<M, T, true>(acc, first, yy_yf[L], A.C.v[1][1]);
                        window_ref_arr<M, T, false>(acc, first, yz_zb, A.C.v[1][2]);
                        vout[L][1] = add_rn<T>(acc, vself[L]);
                        acc = 0; first = true;
                        window_ref_arr<M, T, false>(acc, first, txz[L], A.C.v[2][0]);
                        window_ref_arr<M, T, false>(acc, first, yz_yb[L], A.C.v[2][1]);
                        window_ref_arr<M, T, true>(acc, first, zz_zf, A.C.v[2][2]);
                        vout[L][2] = add_rn<T>(acc, wself[L]);
                    } else {
                        vout[L][0] = uself[L] + (window_fast_arr<M, T, true>(txx[L], A.C.v[0][0]) + window_fast_arr<M, T, false>(xy_yb[L], A.C.v[0][1]) +
                                                 window_fast_arr<M, T, false>(xz_zb, A.C.v[0][2]));
                        vout[L][1] = vself[L] + (window_fast_arr<M, T, false>(txy[L], A.C.v[1][0]) + window_fast_arr<M, T, true>(yy_yf[L], A.C.v[1][1]) +
                                                 window_fast_arr<M, T, false>(yz_zb, A.C.v[1][2]));
                        vout[L][2] = wself[L] + (window_fast_arr<M, T, false>(txz[L], A.C.v[2][0]) + window_fast_arr<M, T, false>(yz_yb[L], A.C.v[2][1]) +
                                                 window_fast_arr<M, T, true>(zz_zf, A.C.v[2][2]));
                    }
                }
#if OPESCI_SKELETON
#pragma unroll
                for (int L = 0; L < 2; ++L) { vout[L][0] = uself[L] + xy_yb[L][0]; vout[L][1] = vself[L] + fxz[2 + L]; vout[L][2] = wself[L] + txx[L][0]; }
#endif
                const long long pxv = px - (long long)(M + 1) * sx;   // px already points at plane xs+1
                if (vf2) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) gstore2(gV1[k] + pxv, vout[0][k], vout[1][k]);
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (vf_yz[0]) gstore(gV1[k] + pxv, vout[0][k]);
                        if (vf_yz[1]) gstore(gV1[k] + pxv + 1, vout[1][k]);
                    }
                }
            }
#pragma unroll
            for (int L = 0; L < 2; ++L) { vself[L] = vself_next[L]; wself[L] = wself_next[L]; }
        }
    }
#if OPESCI_TMA_STORE
    if (tma_st && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#endif
    if (PAIR) {
        // neither CTA leaves while the other may still push into it or signal it
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    // the kernels after this one read what the z-edge grid wrote: this grid completes only after that one has
    // (no-op when the launch has no programmatic dependency)
    if (!ZF) asm volatile("griddepcontrol.wait;" ::: "memory");
#if OPESCI_PACE > 0
    if (A.pace && tid == 0) A.pace[((size_t)chunk * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = 0x7fffffff;
#endif
}

// All six shell slabs in one launch: blockIdx.x is a flat block index over the boxes.
struct ShellBoxes {
    Range3 r[6];
    int zwide[6];        // 1: 64 x 4 threads (z x y), 0: 4 x 64
    int nbz[6], nby[6];  // blocks along z and y
    int start[7];        // prefix sums of the block counts
};
template <int SO, typename T, int ARITH, bool HET = false>
__global__ void __launch_bounds__(OPESCI_FACE_THREADS)
velocity_shell_kernel(FieldPtrs F, GridGeom G, StaggeredCoefs C, int t0, int t1, const __grid_constant__ ShellBoxes B,
                      MediaPtrs MD, HeteroCoefs HC)
{
    constexpr int M = SO / 2;
    int b = 0;
#pragma unroll
    for (int k = 1; k < 6; ++k)
        if ((int)blockIdx.x >= B.start[k]) b = k;
    const Range3 &R = B.r[b];
    int rem = blockIdx.x - B.start[b];
    const int bz = rem % B.nbz[b];
    rem /= B.nbz[b];
    const int by = rem % B.nby[b], bx = rem / B.nby[b];
    const int tw = B.zwide[b] ? 64 : 4, th = OPESCI_FACE_THREADS / tw;
    const int z = R.lo[2] + bz * tw + (int)(threadIdx.x % tw);
    const int y = R.lo[1] + by * th + (int)(threadIdx.x / tw);
    const int x = R.lo[0] + bx;
    if (z >= R.hi[2] || y >= R.hi[1]) return;
    const long long p = (long long)x * G.s[0] + (long long)y * G.s[1] + z;
    const long long r = (long long)t0 * G.level + p, w = (long long)t1 * G.level + p;
    const long long st[3] = {G.s[0], G.s[1], 1};
    const int opnd[3][3] = {{F_TXX, F_TXY, F_TXZ}, {F_TXY, F_TYY, F_TYZ}, {F_TXZ, F_TYZ, F_TZZ}};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        T *Va = (T *)F.f[F_U + a];
        if constexpr (HET) {
            const float b = MD.m[OPESCI_MEDIA_BETA1 + a][p];
            if (ARITH == OPESCI_ARITH_REFERENCE) {
                float acc = 0;
                bool first = true;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float *g = (const float *)F.f[opnd[a][d]] + w;
                    if (d == a) window_ref_h<M, true, 1>(acc, first, g, st[d], HC.c[d], HC.c2[d], b, b);
                    else window_ref_h<M, false, 1>(acc, first, g, st[d], HC.c[d], HC.c2[d], b, b);
                }
                Va[w] = __fadd_rn(acc, Va[r]);
            } else {
                float acc = 0;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    const float *g = (const float *)F.f[opnd[a][d]] + w;
                    acc += (d == a) ? window_fast<M, float, true>(g, st[d], HC.c[d]) : window_fast<M, float, false>(g, st[d], HC.c[d]);
                }
                Va[w] = Va[r] + b * acc;
            }
        } else if (ARITH == OPESCI_ARITH_REFERENCE) {
            T acc = 0;
            bool first = true;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const T *g = (const T *)F.f[opnd[a][d]] + w;
                if (d == a) window_ref<M, T, true>(acc, first, g, st[d], C.v[a][d]);
                else window_ref<M, T, false>(acc, first, g, st[d], C.v[a][d]);
            }
            Va[w] = add_rn<T>(acc, Va[r]);
        } else {
            T acc = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const T *g = (const T *)F.f[opnd[a][d]] + w;
                acc += (d == a) ? window_fast<M, T, true>(g, st[d], C.v[a][d])
                                : window_fast<M, T, false>(g, st[d], C.v[a][d]);
            }
            Va[w] = Va[r] + acc;
        }
    }
}

}  // namespace opesci
