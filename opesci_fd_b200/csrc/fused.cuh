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
// that is left is updated after the stress ghost loops by velocity_box (kernels.cuh), so the
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

template <int M> struct FusedCfg {
    static constexpr int EZ = 64, EY = 16;                 // threads = stress tile (incl. recomputed halo)
    // stored tile.  TMA needs the innermost box coordinate 16-B aligned (measured on B200: an
    // unaligned start raises "illegal instruction"), so tiles advance in multiples of 4 floats
    // along z and the box starts OFFZ >= M floats left of the stress tile.
    static constexpr int CZ = (EZ - 2 * M) / 4 * 4, CY = EY - 2 * M;
    static constexpr int OFFZ = (M + 3) / 4 * 4;
    static constexpr int VZ = (EZ + M + OFFZ + 3) / 4 * 4, VY = EY + 2 * M;   // velocity tile delivered by TMA
    static constexpr int RD = 2 * M + 2;                   // velocity ring depth (planes)
    static constexpr int SR = 4;                           // in-plane stress ring slots
    static constexpr int VTILE = ((VZ * VY * 4 + 127) / 128) * 128;   // bytes, 128-B aligned for TMA
    static constexpr int STILE = EZ * EY * 4;
    static constexpr int SMEM = 3 * RD * VTILE + 5 * SR * STILE + 3 * RD * 8 + 128;
    static constexpr int THREADS = EZ * EY;
};

struct FusedArgs {
    FieldPtrs F;
    GridGeom G;
    StaggeredCoefs C;
    int t0, t1;
    int xchunk;       // planes per x-chunk
};

template <int SO, int ARITH>
__global__ void __launch_bounds__(FusedCfg<SO / 2>::THREADS, 1)
fused_step(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmV,
           const __grid_constant__ CUtensorMap tmW, const FusedArgs A)
{
    constexpr int M = SO / 2;
    using K = FusedCfg<M>;
    typedef float T;
    constexpr int RD = K::RD;
    constexpr int VT = K::VTILE / 4;      // floats per velocity tile
    constexpr int ST = K::STILE / 4;      // floats per stress tile
    // dynamic shared memory (no static __shared__ in this kernel, so the window starts 1024-B
    // aligned): [3][RD] velocity tiles | [5][SR] stress tiles | mbarriers
    extern __shared__ __align__(1024) unsigned char smem[];
    const T *vring = reinterpret_cast<const T *>(smem);
    T *sring = reinterpret_cast<T *>(smem + (size_t)3 * RD * K::VTILE);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)3 * RD * K::VTILE + (size_t)5 * K::SR * K::STILE);

    const GridGeom &G = A.G;
    const int tid = threadIdx.x;
    const int tz = tid % K::EZ, ty = tid / K::EZ;
    const int ye = blockIdx.y * K::CY + ty, ze = blockIdx.x * K::CZ + tz;   // global coords of this thread's point
    const int xa = M + blockIdx.z * A.xchunk;
    const int xb = min(xa + A.xchunk, G.dim[0] - M);
    const int xs_begin = max(M, xa - M), xs_end = min(G.dim[0] - M, xb + M);
    // plane p of U lives in slot (p - pbaseU) % RD, of V/W in slot (p - pbaseVW) % RD; with the
    // x loop unrolled RD times every slot index below is a compile-time constant
    const int pbaseU = xs_begin - M, pbaseVW = xs_begin - M + 1;
    const int lastU = xs_end + M - 2, lastVW = xs_end + M - 1;
    const int c0 = blockIdx.x * K::CZ - K::OFFZ, c1 = blockIdx.y * K::CY - M;   // TMA box origin (may be negative)
    const int lvl0 = A.t0 * G.dim[0];
    constexpr uint32_t TILE_BYTES = K::VZ * K::VY * 4;

    if (tid == 0) {
        for (int i = 0; i < 3 * RD; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
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

    const bool inb = ye < G.dim[1] && ze < G.dim[2];
    const bool core = ty >= M && ty < M + K::CY && tz >= M && tz < M + K::CZ;
    const bool st_yz = core && ye >= M && ye < G.dim[1] - M && ze >= M && ze < G.dim[2] - M;
    const bool vf_yz = core && ye >= 2 * M + 1 && ye < G.dim[1] - 2 * M - 1 && ze >= 2 * M + 1 && ze < G.dim[2] - 2 * M - 1;
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

    // register state: own-column x-windows of the new stresses
    T txx[2 * M], txy[2 * M + 1], txz[2 * M + 1];
#pragma unroll
    for (int k = 0; k < 2 * M; ++k) txx[k] = 0;
#pragma unroll
    for (int k = 0; k < 2 * M + 1; ++k) txy[k] = txz[k] = 0;
    T vself = 0, wself = 0;                // V,W[t0] at plane xs-M (saved one iteration earlier)
    const int lo = (ty + M) * K::VZ + tz + K::OFFZ;   // this thread's element inside a velocity tile
    const T *vlo = vring + lo;
    T *slo = sring + ty * K::EZ + tz;

    // T[t0] of the first plane
    T told[6];
    long long px = (long long)xs_begin * sx;
#pragma unroll
    for (int k = 0; k < 6; ++k) told[k] = inb ? gT0[k][px] : (T)0;

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
            // ---- the newest planes of this iteration: relative plane index RD*q + r + 2M-1
            {
                constexpr int rel = 0;   // placeholder to keep the block scoped
                (void)rel;
                const int slot = (r + 2 * M - 1) % RD;
                const uint32_t par = (uint32_t)(q + (r + 2 * M - 1) / RD) & 1u;
                mbar_wait(&bars[0 * RD + slot], par);
                mbar_wait(&bars[1 * RD + slot], par);
                mbar_wait(&bars[2 * RD + slot], par);
            }
            // ---- gather the operands of the six stress updates (all offsets are immediates)
            T ux[2 * M], vx[2 * M], wx[2 * M];   // x-windows: U bwd (xs-M..xs+M-1), V,W fwd (xs-M+1..xs+M)
#pragma unroll
            for (int j = 0; j < 2 * M; ++j) {
                const int slot = (r + j) % RD;
                ux[j] = vlo[(0 * RD + slot) * VT];
                vx[j] = vlo[(1 * RD + slot) * VT];
                wx[j] = vlo[(2 * RD + slot) * VT];
            }
            // plane xs: U is window entry M (slot r+M), V/W window entry M-1 (slot r+M-1)
            const T *pu = vlo + (0 * RD + (r + M) % RD) * VT;
            const T *pv = vlo + (1 * RD + (r + M - 1) % RD) * VT;
            const T *pw = vlo + (2 * RD + (r + M - 1) % RD) * VT;
            T vy_b[2 * M], wz_b[2 * M];          // backward windows in-plane (normal stresses)
            T uy_f[2 * M], uz_f[2 * M], vz_f[2 * M], wy_f[2 * M];   // forward windows in-plane (shear stresses)
#pragma unroll
            for (int j = 0; j < 2 * M; ++j) {
                vy_b[j] = pv[(j - M) * K::VZ];
                wz_b[j] = pw[(j - M)];
                uy_f[j] = pu[(j - M + 1) * K::VZ];
                uz_f[j] = pu[(j - M + 1)];
                vz_f[j] = pv[(j - M + 1)];
                wy_f[j] = pw[(j - M + 1) * K::VZ];
            }
            const T uself = ux[0];               // U[t0] at plane xs-M (velocity self term of this iteration)
            const T vself_next = vx[0], wself_next = wx[0];   // V,W[t0] at plane xs-M+1
            T tn[6];
            if (ARITH == OPESCI_ARITH_REFERENCE) {
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    T acc = told[a];
                    bool first = false;
                    window_ref_arr<M, T, false>(acc, first, ux, A.C.sn[a][0]);
                    window_ref_arr<M, T, false>(acc, first, vy_b, A.C.sn[a][1]);
                    window_ref_arr<M, T, false>(acc, first, wz_b, A.C.sn[a][2]);
                    tn[a] = acc;
                }
                {
                    T acc = told[3]; bool first = false;   // Txy: D_y U, D_x V
                    window_ref_arr<M, T, true>(acc, first, uy_f, A.C.ss[0][0]);
                    window_ref_arr<M, T, true>(acc, first, vx, A.C.ss[0][1]);
                    tn[3] = acc;
                }
                {
                    T acc = told[4]; bool first = false;   // Tyz: D_z V, D_y W
                    window_ref_arr<M, T, true>(acc, first, vz_f, A.C.ss[1][0]);
                    window_ref_arr<M, T, true>(acc, first, wy_f, A.C.ss[1][1]);
                    tn[4] = acc;
                }
                {
                    T acc = told[5]; bool first = false;   // Txz: D_z U, D_x W
                    window_ref_arr<M, T, true>(acc, first, uz_f, A.C.ss[2][0]);
                    window_ref_arr<M, T, true>(acc, first, wx, A.C.ss[2][1]);
                    tn[5] = acc;
                }
            } else {
#pragma unroll
                for (int a = 0; a < 3; ++a)
                    tn[a] = told[a] + (window_fast_arr<M, T, false>(ux, A.C.sn[a][0]) + window_fast_arr<M, T, false>(vy_b, A.C.sn[a][1]) +
                                       window_fast_arr<M, T, false>(wz_b, A.C.sn[a][2]));
                tn[3] = told[3] + (window_fast_arr<M, T, true>(uy_f, A.C.ss[0][0]) + window_fast_arr<M, T, true>(vx, A.C.ss[0][1]));
                tn[4] = told[4] + (window_fast_arr<M, T, true>(vz_f, A.C.ss[1][0]) + window_fast_arr<M, T, true>(wy_f, A.C.ss[1][1]));
                tn[5] = told[5] + (window_fast_arr<M, T, true>(uz_f, A.C.ss[2][0]) + window_fast_arr<M, T, true>(wx, A.C.ss[2][1]));
            }
            // ---- store the new stresses (owned tile, owned planes), prefetch next T[t0]
            if (st_yz && xs >= xa && xs < xb) {
#pragma unroll
                for (int k = 0; k < 6; ++k) gT1[k][px] = tn[k];
            }
            px += sx;
            if (xs + 1 < xs_end) {
#pragma unroll
                for (int k = 0; k < 6; ++k) told[k] = inb ? gT0[k][px] : (T)0;
            }
            // ---- shift the register windows, publish the in-plane operands
#pragma unroll
            for (int k = 0; k < 2 * M - 1; ++k) txx[k] = txx[k + 1];
            txx[2 * M - 1] = tn[0];
#pragma unroll
            for (int k = 0; k < 2 * M; ++k) { txy[k] = txy[k + 1]; txz[k] = txz[k + 1]; }
            txy[2 * M] = tn[3];
            txz[2 * M] = tn[5];
            {
                T *s = slo + (xs & (K::SR - 1)) * ST;
                s[0 * K::SR * ST] = tn[3];   // Txy
                s[1 * K::SR * ST] = tn[5];   // Txz
                s[2 * K::SR * ST] = tn[1];   // Tyy
                s[3 * K::SR * ST] = tn[4];   // Tyz
                s[4 * K::SR * ST] = tn[2];   // Tzz
            }
            __syncthreads();
            // ---- the oldest planes (window entry 0, slot r) are dead: refill their slots
            if (tid == 0) {
                const int pU = xs - M + RD, pVW = xs - M + 1 + RD;
                if (pU <= lastU) {
                    mbar_arrive_expect_tx(&bars[0 * RD + r], TILE_BYTES);
                    tma_load_3d((void *)(vring + (0 * RD + r) * VT), &tmU, &bars[0 * RD + r], c0, c1, lvl0 + pU);
                }
                if (pVW <= lastVW) {
                    mbar_arrive_expect_tx(&bars[1 * RD + r], TILE_BYTES);
                    tma_load_3d((void *)(vring + (1 * RD + r) * VT), &tmV, &bars[1 * RD + r], c0, c1, lvl0 + pVW);
                    mbar_arrive_expect_tx(&bars[2 * RD + r], TILE_BYTES);
                    tma_load_3d((void *)(vring + (2 * RD + r) * VT), &tmW, &bars[2 * RD + r], c0, c1, lvl0 + pVW);
                }
            }
            // ---- velocities of plane xv = xs - M from the new stresses xv-M .. xv+M
            const int xv = xs - M;
            if (vf_yz && xv >= xv_lo && xv < xv_hi) {
                const T *s = slo + (xv & (K::SR - 1)) * ST;
                const T *sxy = s, *sxz = s + 1 * K::SR * ST, *syy = s + 2 * K::SR * ST, *syz = s + 3 * K::SR * ST,
                        *szz = s + 4 * K::SR * ST;
                T xy_yb[2 * M], xz_zb[2 * M], yy_yf[2 * M], yz_zb[2 * M], yz_yb[2 * M], zz_zf[2 * M];
#pragma unroll
                for (int j = 0; j < 2 * M; ++j) {
                    xy_yb[j] = sxy[(j - M) * K::EZ];
                    xz_zb[j] = sxz[(j - M)];
                    yy_yf[j] = syy[(j - M + 1) * K::EZ];
                    yz_zb[j] = syz[(j - M)];
                    yz_yb[j] = syz[(j - M) * K::EZ];
                    zz_zf[j] = szz[(j - M + 1)];
                }
                // x-windows from registers: Txx fwd = planes xv-M+1..xv+M = txx[0..2M-1];
                // Txy, Txz bwd = planes xv-M..xv+M-1 = txy[0..2M-1]
                T un, vn, wn;
                if (ARITH == OPESCI_ARITH_REFERENCE) {
                    T acc = 0; bool first = true;
                    window_ref_arr<M, T, true>(acc, first, txx, A.C.v[0][0]);
                    window_ref_arr<M, T, false>(acc, first, xy_yb, A.C.v[0][1]);
                    window_ref_arr<M, T, false>(acc, first, xz_zb, A.C.v[0][2]);
                    un = add_rn<T>(acc, uself);
                    acc = 0; first = true;
                    window_ref_arr<M, T, false>(acc, first, txy, A.C.v[1][0]);
                    window_ref_arr<M, T, true>(acc, first, yy_yf, A.C.v[1][1]);
                    window_ref_arr<M, T, false>(acc, first, yz_zb, A.C.v[1][2]);
                    vn = add_rn<T>(acc, vself);
                    acc = 0; first = true;
                    window_ref_arr<M, T, false>(acc, first, txz, A.C.v[2][0]);
                    window_ref_arr<M, T, false>(acc, first, yz_yb, A.C.v[2][1]);
                    window_ref_arr<M, T, true>(acc, first, zz_zf, A.C.v[2][2]);
                    wn = add_rn<T>(acc, wself);
                } else {
                    un = uself + (window_fast_arr<M, T, true>(txx, A.C.v[0][0]) + window_fast_arr<M, T, false>(xy_yb, A.C.v[0][1]) +
                                  window_fast_arr<M, T, false>(xz_zb, A.C.v[0][2]));
                    vn = vself + (window_fast_arr<M, T, false>(txy, A.C.v[1][0]) + window_fast_arr<M, T, true>(yy_yf, A.C.v[1][1]) +
                                  window_fast_arr<M, T, false>(yz_zb, A.C.v[1][2]));
                    wn = wself + (window_fast_arr<M, T, false>(txz, A.C.v[2][0]) + window_fast_arr<M, T, false>(yz_yb, A.C.v[2][1]) +
                                  window_fast_arr<M, T, true>(zz_zf, A.C.v[2][2]));
                }
                const long long pxv = px - (long long)(M + 1) * sx;   // px already points at plane xs+1
                gV1[0][pxv] = un;
                gV1[1][pxv] = vn;
                gV1[2][pxv] = wn;
            }
            vself = vself_next;
            wself = wself_next;
        }
    }
}

// velocity update on a box of interior points (the shell the fused kernel leaves out), operands
// from global memory: identical arithmetic to velocity_interior
template <int SO, typename T, int ARITH>
__global__ void __launch_bounds__(256)
velocity_box(FieldPtrs F, GridGeom G, StaggeredCoefs C, int t0, int t1, Range3 R)
{
    constexpr int M = SO / 2;
    const int z = R.lo[2] + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = R.lo[1] + blockIdx.y * blockDim.y + threadIdx.y;
    const int x = R.lo[0] + blockIdx.z;
    if (z >= R.hi[2] || y >= R.hi[1]) return;
    const long long p = (long long)x * G.s[0] + (long long)y * G.s[1] + z;
    const long long r = (long long)t0 * G.level + p, w = (long long)t1 * G.level + p;
    const long long st[3] = {G.s[0], G.s[1], 1};
    const int opnd[3][3] = {{F_TXX, F_TXY, F_TXZ}, {F_TXY, F_TYY, F_TYZ}, {F_TXZ, F_TYZ, F_TZZ}};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        T *Va = (T *)F.f[F_U + a];
        if (ARITH == OPESCI_ARITH_REFERENCE) {
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

// All six shell slabs in one launch: blockIdx.x is a flat block index over the boxes.
struct ShellBoxes {
    Range3 r[6];
    int zwide[6];        // 1: 64 x 4 threads (z x y), 0: 4 x 64
    int nbz[6], nby[6];  // blocks along z and y
    int start[7];        // prefix sums of the block counts
};
template <int SO, typename T, int ARITH>
__global__ void __launch_bounds__(256)
velocity_shell_kernel(FieldPtrs F, GridGeom G, StaggeredCoefs C, int t0, int t1, const __grid_constant__ ShellBoxes B)
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
    const int tw = B.zwide[b] ? 64 : 4, th = 256 / tw;
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
        if (ARITH == OPESCI_ARITH_REFERENCE) {
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
