// tiled.cuh -- two-pass stress / velocity kernels for every spatial order and precision (sm_100a).
//
// Replace the one-thread-per-point stress_interior / velocity_interior (kernels.cuh) for the configurations the
// fused kernel does not cover (so >= 6, fp64; BASELINE config 4): the same two loops of the reference
// (opesci/staggeredgrid.py:728-748 emitted through opesci/regulargrid.py:566-619).  Compulsory traffic
// 15 + 12 = 27 words per point and step (108 B fp32, 216 B fp64).
//
// One CTA (TY x TZ threads, one point each) owns a (y,z) tile and marches along x through an x-chunk.
//   * the y- and z-windows of the centre plane come from shared memory: the centre-plane tiles (+ m halo rows /
//     columns, zero-filled outside the array) of the operand fields arrive by TMA (cp.async.bulk.tensor.3d) into
//     an NB-deep ring, signalled through mbarriers, NB-1 planes ahead of their use;
//   * the x-windows of the thread's own column live in registers: one coalesced global load per field and plane,
//     issued one plane ahead (these planes are re-read from L2 by the TMA m planes later);
//   * one __syncthreads per plane recycles the ring slot.
// The one-thread-per-point kernels are latency-bound (ncu, so=8: 64 % warps active, every pipe < 50 %) because each
// point needs 6m dependent-latency loads; here 4m of them are shared-memory reads and the HBM stream is
// asynchronous.  Reference arithmetic: same term order and roundings as the emitted code (window_ref_arr).
#pragma once
#include "fused.cuh"
#include "hetero.cuh"
#include "kernels.cuh"

namespace opesci {

template <int M, typename T> struct TileCfg {
    static constexpr int TY = 8, TZ = 32;                       // threads = points of the tile
    static constexpr int VY = TY + 2 * M;
    static constexpr int AL = 16 / (int)sizeof(T);              // elements per 16 bytes (TMA box granularity)
    static constexpr int VZ = (TZ + 2 * M + AL - 1) / AL * AL;  // row length of a tile in shared memory
    static constexpr int NB = 3;                                // ring depth (planes)
    static constexpr int TILE = ((VZ * VY * (int)sizeof(T) + 127) / 128) * 128;   // bytes, 128-B aligned for TMA
    static constexpr int THREADS = TY * TZ;
#ifndef OPESCI_TILED_MINB_F64
#define OPESCI_TILED_MINB_F64 2
#endif
#ifndef OPESCI_TILED_MINB_F32
#define OPESCI_TILED_MINB_F32 0   /* 0 = unspecified: the compiler keeps its own occupancy heuristic (so=8: 71 / 64 registers, 3-4 CTAs per SM) */
#endif
    // resident CTAs per SM the register allocation must leave room for: the fp64 kernels otherwise take
    // 160+ registers (every window array live at once) and run with a single 8-warp CTA per SM
    // (so=8 fp64 768^3: 15.7 -> 19.8 Gpts/s; at m >= 5 the 128-register cap spills and loses)
    static constexpr int MINB = (sizeof(T) == 8 && M <= 4) ? OPESCI_TILED_MINB_F64 : OPESCI_TILED_MINB_F32;
    static constexpr int smem(int nfields) { return nfields * NB * TILE + nfields * NB * 8 + 128; }
};

struct TileArgs {
    FieldPtrs F;
    GridGeom G;
    StaggeredCoefs C;
    MediaPtrs MD;     // heterogeneous mode (HET kernels, fp32): per-cell media, see hetero.cuh
    HeteroCoefs HC;
    int t0, t1;
    int xchunk;
};

// stress pass: T[t1] = T[t0] + windows of U,V,W[t0]
template <int SO, typename T, int ARITH, bool HET = false>
__global__ void __launch_bounds__(TileCfg<SO / 2, T>::THREADS, TileCfg<SO / 2, T>::MINB)
stress_tiled(const __grid_constant__ CUtensorMap tmU, const __grid_constant__ CUtensorMap tmV,
             const __grid_constant__ CUtensorMap tmW, const TileArgs A)
{
    constexpr int M = SO / 2;
    using K = TileCfg<M, T>;
    constexpr int NB = K::NB, VZ = K::VZ;
    constexpr int TE = K::TILE / (int)sizeof(T);
    extern __shared__ __align__(1024) unsigned char smem[];
    const T *ring = reinterpret_cast<const T *>(smem);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)3 * NB * K::TILE);
    const GridGeom &G = A.G;
    const int tid = threadIdx.x, tz = tid % K::TZ, ty = tid / K::TZ;
    const int y = M + blockIdx.y * K::TY + ty, z = M + blockIdx.x * K::TZ + tz;
    const int xa = M + blockIdx.z * A.xchunk;
    const int xb = min(xa + A.xchunk, G.dim[0] - M);
    const bool ok = y < G.dim[1] - M && z < G.dim[2] - M;
    const int c0 = blockIdx.x * K::TZ, c1 = blockIdx.y * K::TY;   // box origin = tile origin - m = multiple of the tile size
    const int lvl0 = A.t0 * G.dim[0];
    constexpr uint32_t BYTES = K::VZ * K::VY * sizeof(T);
    const CUtensorMap *tms[3] = {&tmU, &tmV, &tmW};
    if (tid == 0) {
        for (int i = 0; i < 3 * NB; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < NB; ++k)
            if (xa + k < xb)
                for (int f = 0; f < 3; ++f) {
                    mbar_arrive_expect_tx(&bars[f * NB + k], BYTES);
                    tma_load_3d((void *)(ring + (size_t)(f * NB + k) * TE), tms[f], &bars[f * NB + k], c0, c1, lvl0 + xa + k);
                }
    }
    const long long sx = G.s[0];
    const long long col = (long long)(ok ? y : M) * G.s[1] + (ok ? z : M);   // threads outside the interior shadow a valid column
    const T *U = (const T *)A.F.f[F_U] + (long long)A.t0 * G.level + col;
    const T *V = (const T *)A.F.f[F_V] + (long long)A.t0 * G.level + col;
    const T *W = (const T *)A.F.f[F_W] + (long long)A.t0 * G.level + col;
    const T *T0[6];
    T *T1[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        T0[k] = (const T *)A.F.f[F_TXX + k] + (long long)A.t0 * G.level + col;
        T1[k] = (T *)A.F.f[F_TXX + k] + (long long)A.t1 * G.level + col;
    }
    // x-windows of the own column: U backward (planes x-M .. x+M-1), V and W forward (x-M+1 .. x+M)
    T uw[2 * M], vw[2 * M], ww[2 * M];
#pragma unroll
    for (int j = 0; j < 2 * M - 1; ++j) {
        uw[j + 1] = U[(long long)(xa - M + j) * sx];
        vw[j + 1] = V[(long long)(xa - M + 1 + j) * sx];
        ww[j + 1] = W[(long long)(xa - M + 1 + j) * sx];
    }
    T nu = U[(long long)(xa + M - 1) * sx], nv = V[(long long)(xa + M) * sx], nw = W[(long long)(xa + M) * sx];
    T told[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) told[k] = T0[k][(long long)xa * sx];
    // heterogeneous mode: lambda, mu, mu12, mu23, mu13 of the own cell, loaded one plane ahead like T[t0]
    const int MID[5] = {OPESCI_MEDIA_LAMBDA, OPESCI_MEDIA_MU, OPESCI_MEDIA_MU12, OPESCI_MEDIA_MU23, OPESCI_MEDIA_MU13};
    float mnext[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    if (HET) {
#pragma unroll
        for (int k = 0; k < 5; ++k) mnext[k] = A.MD.m[MID[k]][col + (long long)xa * sx];
    }
    const int ctr = (ty + M) * VZ + tz + M;   // own element inside a tile
    for (int x = xa, it = 0; x < xb; ++x, ++it) {
        const int slot = it % NB;
        const uint32_t par = (uint32_t)(it / NB) & 1u;
        const long long px = (long long)x * sx;
#pragma unroll
        for (int j = 0; j < 2 * M - 1; ++j) { uw[j] = uw[j + 1]; vw[j] = vw[j + 1]; ww[j] = ww[j + 1]; }
        uw[2 * M - 1] = nu; vw[2 * M - 1] = nv; ww[2 * M - 1] = nw;
        T tcur[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) tcur[k] = told[k];
        if (x + 1 < xb) {   // next plane's compulsory loads, one iteration ahead
            nu = U[px + (long long)M * sx];
            nv = V[px + (long long)(M + 1) * sx];
            nw = W[px + (long long)(M + 1) * sx];
#pragma unroll
            for (int k = 0; k < 6; ++k) told[k] = T0[k][px + sx];
        }
        float med[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) med[k] = mnext[k];
        if (HET && x + 1 < xb) {
#pragma unroll
            for (int k = 0; k < 5; ++k) mnext[k] = A.MD.m[MID[k]][col + px + sx];
        }
        mbar_wait(&bars[0 * NB + slot], par);
        mbar_wait(&bars[1 * NB + slot], par);
        mbar_wait(&bars[2 * NB + slot], par);
        const T *su = ring + (size_t)(0 * NB + slot) * TE + ctr;
        const T *sv = ring + (size_t)(1 * NB + slot) * TE + ctr;
        const T *sw = ring + (size_t)(2 * NB + slot) * TE + ctr;
        T vy[2 * M], uy[2 * M], wy[2 * M], wzb[2 * M], uzf[2 * M], vzf[2 * M];
#pragma unroll
        for (int j = 0; j < 2 * M; ++j) {
            vy[j] = sv[(j - M) * VZ];        // V backward in y
            uy[j] = su[(j - M + 1) * VZ];    // U, W forward in y
            wy[j] = sw[(j - M + 1) * VZ];
            wzb[j] = sw[j - M];              // W backward in z
            uzf[j] = su[j - M + 1];          // U, V forward in z
            vzf[j] = sv[j - M + 1];
        }
        T out[6];
        if constexpr (HET) {
            // per-cell media: every term is (literal*G)*media in reference arithmetic, factored in fast arithmetic
            const float lam = med[0], mu = med[1];
            if (ARITH == OPESCI_ARITH_REFERENCE) {
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    float acc = tcur[a];
                    bool first = false;
                    if (a == 0) window_ref_arr_h<M, false, 2>(acc, first, uw, A.HC.c[0], A.HC.c2[0], lam, mu);
                    else window_ref_arr_h<M, false, 1>(acc, first, uw, A.HC.c[0], A.HC.c2[0], lam, mu);
                    if (a == 1) window_ref_arr_h<M, false, 2>(acc, first, vy, A.HC.c[1], A.HC.c2[1], lam, mu);
                    else window_ref_arr_h<M, false, 1>(acc, first, vy, A.HC.c[1], A.HC.c2[1], lam, mu);
                    if (a == 2) window_ref_arr_h<M, false, 2>(acc, first, wzb, A.HC.c[2], A.HC.c2[2], lam, mu);
                    else window_ref_arr_h<M, false, 1>(acc, first, wzb, A.HC.c[2], A.HC.c2[2], lam, mu);
                    out[a] = acc;
                }
                {
                    float acc = tcur[3]; bool first = false;   // Txy (mu12)
                    window_ref_arr_h<M, true, 1>(acc, first, uy, A.HC.c[1], A.HC.c2[1], med[2], 0.f);
                    window_ref_arr_h<M, true, 1>(acc, first, vw, A.HC.c[0], A.HC.c2[0], med[2], 0.f);
                    out[3] = acc;
                }
                {
                    float acc = tcur[4]; bool first = false;   // Tyz (mu23)
                    window_ref_arr_h<M, true, 1>(acc, first, vzf, A.HC.c[2], A.HC.c2[2], med[3], 0.f);
                    window_ref_arr_h<M, true, 1>(acc, first, wy, A.HC.c[1], A.HC.c2[1], med[3], 0.f);
                    out[4] = acc;
                }
                {
                    float acc = tcur[5]; bool first = false;   // Txz (mu13)
                    window_ref_arr_h<M, true, 1>(acc, first, uzf, A.HC.c[2], A.HC.c2[2], med[4], 0.f);
                    window_ref_arr_h<M, true, 1>(acc, first, ww, A.HC.c[0], A.HC.c2[0], med[4], 0.f);
                    out[5] = acc;
                }
            } else {
                const T du = window_fast_arr<M, T, false>(uw, A.HC.c[0]), dv = window_fast_arr<M, T, false>(vy, A.HC.c[1]),
                        dw = window_fast_arr<M, T, false>(wzb, A.HC.c[2]);
                const T tr = lam * (du + dv + dw), mu2 = 2.0f * mu;
                out[0] = tcur[0] + (tr + mu2 * du);
                out[1] = tcur[1] + (tr + mu2 * dv);
                out[2] = tcur[2] + (tr + mu2 * dw);
                out[3] = tcur[3] + med[2] * (window_fast_arr<M, T, true>(uy, A.HC.c[1]) + window_fast_arr<M, T, true>(vw, A.HC.c[0]));
                out[4] = tcur[4] + med[3] * (window_fast_arr<M, T, true>(vzf, A.HC.c[2]) + window_fast_arr<M, T, true>(wy, A.HC.c[1]));
                out[5] = tcur[5] + med[4] * (window_fast_arr<M, T, true>(uzf, A.HC.c[2]) + window_fast_arr<M, T, true>(ww, A.HC.c[0]));
            }
        } else if (ARITH == OPESCI_ARITH_REFERENCE) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                T acc = tcur[a];
                bool first = false;
                window_ref_arr<M, T, false>(acc, first, uw, A.C.sn[a][0]);
                window_ref_arr<M, T, false>(acc, first, vy, A.C.sn[a][1]);
                window_ref_arr<M, T, false>(acc, first, wzb, A.C.sn[a][2]);
                out[a] = acc;
            }
            {
                T acc = tcur[3]; bool first = false;   // Txy: D_y U, D_x V
                window_ref_arr<M, T, true>(acc, first, uy, A.C.ss[0][0]);
                window_ref_arr<M, T, true>(acc, first, vw, A.C.ss[0][1]);
                out[3] = acc;
            }
            {
                T acc = tcur[4]; bool first = false;   // Tyz: D_z V, D_y W
                window_ref_arr<M, T, true>(acc, first, vzf, A.C.ss[1][0]);
                window_ref_arr<M, T, true>(acc, first, wy, A.C.ss[1][1]);
                out[4] = acc;
            }
            {
                T acc = tcur[5]; bool first = false;   // Txz: D_z U, D_x W
                window_ref_arr<M, T, true>(acc, first, uzf, A.C.ss[2][0]);
                window_ref_arr<M, T, true>(acc, first, ww, A.C.ss[2][1]);
                out[5] = acc;
            }
        } else {
#pragma unroll
            for (int a = 0; a < 3; ++a)
                out[a] = tcur[a] + (window_fast_arr<M, T, false>(uw, A.C.sn[a][0]) + window_fast_arr<M, T, false>(vy, A.C.sn[a][1]) +
                                    window_fast_arr<M, T, false>(wzb, A.C.sn[a][2]));
            out[3] = tcur[3] + (window_fast_arr<M, T, true>(uy, A.C.ss[0][0]) + window_fast_arr<M, T, true>(vw, A.C.ss[0][1]));
            out[4] = tcur[4] + (window_fast_arr<M, T, true>(vzf, A.C.ss[1][0]) + window_fast_arr<M, T, true>(wy, A.C.ss[1][1]));
            out[5] = tcur[5] + (window_fast_arr<M, T, true>(uzf, A.C.ss[2][0]) + window_fast_arr<M, T, true>(ww, A.C.ss[2][1]));
        }
        if (ok) {
#pragma unroll
            for (int k = 0; k < 6; ++k) T1[k][px] = out[k];
        }
        __syncthreads();   // every thread has read the tiles of plane x: refill the slot with plane x+NB
        if (tid == 0 && x + NB < xb) {
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                mbar_arrive_expect_tx(&bars[f * NB + slot], BYTES);
                tma_load_3d((void *)(ring + (size_t)(f * NB + slot) * TE), tms[f], &bars[f * NB + slot], c0, c1, lvl0 + x + NB);
            }
        }
    }
}

// velocity pass: V[t1] = windows of T[t1] + V[t0].  Tiles: Txy, Tyy, Tyz, Txz, Tzz; x-windows: Txx, Txy, Txz.
template <int SO, typename T, int ARITH, bool HET = false>
__global__ void __launch_bounds__(TileCfg<SO / 2, T>::THREADS, TileCfg<SO / 2, T>::MINB)
velocity_tiled(const __grid_constant__ CUtensorMap tmXY, const __grid_constant__ CUtensorMap tmYY,
               const __grid_constant__ CUtensorMap tmYZ, const __grid_constant__ CUtensorMap tmXZ,
               const __grid_constant__ CUtensorMap tmZZ, const TileArgs A)
{
    constexpr int M = SO / 2;
    using K = TileCfg<M, T>;
    constexpr int NB = K::NB, VZ = K::VZ;
    constexpr int TE = K::TILE / (int)sizeof(T);
    extern __shared__ __align__(1024) unsigned char smem[];
    const T *ring = reinterpret_cast<const T *>(smem);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)5 * NB * K::TILE);
    const GridGeom &G = A.G;
    const int tid = threadIdx.x, tz = tid % K::TZ, ty = tid / K::TZ;
    const int y = M + blockIdx.y * K::TY + ty, z = M + blockIdx.x * K::TZ + tz;
    const int xa = M + blockIdx.z * A.xchunk;
    const int xb = min(xa + A.xchunk, G.dim[0] - M);
    const bool ok = y < G.dim[1] - M && z < G.dim[2] - M;
    const int c0 = blockIdx.x * K::TZ, c1 = blockIdx.y * K::TY;
    const int lvl1 = A.t1 * G.dim[0];
    constexpr uint32_t BYTES = K::VZ * K::VY * sizeof(T);
    const CUtensorMap *tms[5] = {&tmXY, &tmYY, &tmYZ, &tmXZ, &tmZZ};
    if (tid == 0) {
        for (int i = 0; i < 5 * NB; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        for (int k = 0; k < NB; ++k)
            if (xa + k < xb)
                for (int f = 0; f < 5; ++f) {
                    mbar_arrive_expect_tx(&bars[f * NB + k], BYTES);
                    tma_load_3d((void *)(ring + (size_t)(f * NB + k) * TE), tms[f], &bars[f * NB + k], c0, c1, lvl1 + xa + k);
                }
    }
    const long long sx = G.s[0];
    const long long col = (long long)(ok ? y : M) * G.s[1] + (ok ? z : M);
    const long long l1 = (long long)A.t1 * G.level + col, l0 = (long long)A.t0 * G.level + col;
    const T *Txx = (const T *)A.F.f[F_TXX] + l1, *Txy = (const T *)A.F.f[F_TXY] + l1, *Txz = (const T *)A.F.f[F_TXZ] + l1;
    // x-windows of the own column: Txx forward (planes x-M+1 .. x+M), Txy and Txz backward (x-M .. x+M-1)
    T xx[2 * M], xy[2 * M], xz[2 * M];
#pragma unroll
    for (int j = 0; j < 2 * M - 1; ++j) {
        xx[j + 1] = Txx[(long long)(xa - M + 1 + j) * sx];
        xy[j + 1] = Txy[(long long)(xa - M + j) * sx];
        xz[j + 1] = Txz[(long long)(xa - M + j) * sx];
    }
    T na = Txx[(long long)(xa + M) * sx], nb = Txy[(long long)(xa + M - 1) * sx], nc = Txz[(long long)(xa + M - 1) * sx];
    T snext[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) snext[k] = ((const T *)A.F.f[F_U + k])[l0 + (long long)xa * sx];
    float bnext[3] = {0.f, 0.f, 0.f};   // heterogeneous mode: beta1, beta2, beta3 of the own cell, one plane ahead
    if (HET) {
#pragma unroll
        for (int k = 0; k < 3; ++k) bnext[k] = A.MD.m[OPESCI_MEDIA_BETA1 + k][col + (long long)xa * sx];
    }
    const int ctr = (ty + M) * VZ + tz + M;
    for (int x = xa, it = 0; x < xb; ++x, ++it) {
        const int slot = it % NB;
        const uint32_t par = (uint32_t)(it / NB) & 1u;
        const long long px = (long long)x * sx;
#pragma unroll
        for (int j = 0; j < 2 * M - 1; ++j) { xx[j] = xx[j + 1]; xy[j] = xy[j + 1]; xz[j] = xz[j + 1]; }
        xx[2 * M - 1] = na; xy[2 * M - 1] = nb; xz[2 * M - 1] = nc;
        T self[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) self[k] = snext[k];
        if (x + 1 < xb) {
            na = Txx[px + (long long)(M + 1) * sx];
            nb = Txy[px + (long long)M * sx];
            nc = Txz[px + (long long)M * sx];
#pragma unroll
            for (int k = 0; k < 3; ++k) snext[k] = ((const T *)A.F.f[F_U + k])[l0 + px + sx];
        }
        float bet[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) bet[k] = bnext[k];
        if (HET && x + 1 < xb) {
#pragma unroll
            for (int k = 0; k < 3; ++k) bnext[k] = A.MD.m[OPESCI_MEDIA_BETA1 + k][col + px + sx];
        }
#pragma unroll
        for (int f = 0; f < 5; ++f) mbar_wait(&bars[f * NB + slot], par);
        const T *sxy = ring + (size_t)(0 * NB + slot) * TE + ctr, *syy = ring + (size_t)(1 * NB + slot) * TE + ctr;
        const T *syz = ring + (size_t)(2 * NB + slot) * TE + ctr, *sxz = ring + (size_t)(3 * NB + slot) * TE + ctr;
        const T *szz = ring + (size_t)(4 * NB + slot) * TE + ctr;
        T xy_y[2 * M], yy_y[2 * M], yz_y[2 * M], xz_z[2 * M], yz_z[2 * M], zz_z[2 * M];
#pragma unroll
        for (int j = 0; j < 2 * M; ++j) {
            xy_y[j] = sxy[(j - M) * VZ];       // Txy backward in y (U)
            yy_y[j] = syy[(j - M + 1) * VZ];   // Tyy forward in y (V)
            yz_y[j] = syz[(j - M) * VZ];       // Tyz backward in y (W)
            xz_z[j] = sxz[j - M];              // Txz backward in z (U)
            yz_z[j] = syz[j - M];              // Tyz backward in z (V)
            zz_z[j] = szz[j - M + 1];          // Tzz forward in z (W)
        }
        T out[3];
        if constexpr (HET) {
            if (ARITH == OPESCI_ARITH_REFERENCE) {
                float acc = 0; bool first = true;
                window_ref_arr_h<M, true, 1>(acc, first, xx, A.HC.c[0], A.HC.c2[0], bet[0], 0.f);
                window_ref_arr_h<M, false, 1>(acc, first, xy_y, A.HC.c[1], A.HC.c2[1], bet[0], 0.f);
                window_ref_arr_h<M, false, 1>(acc, first, xz_z, A.HC.c[2], A.HC.c2[2], bet[0], 0.f);
                out[0] = __fadd_rn(acc, self[0]);
                acc = 0; first = true;
                window_ref_arr_h<M, false, 1>(acc, first, xy, A.HC.c[0], A.HC.c2[0], bet[1], 0.f);
                window_ref_arr_h<M, true, 1>(acc, first, yy_y, A.HC.c[1], A.HC.c2[1], bet[1], 0.f);
                window_ref_arr_h<M, false, 1>(acc, first, yz_z, A.HC.c[2], A.HC.c2[2], bet[1], 0.f);
                out[1] = __fadd_rn(acc, self[1]);
                acc = 0; first = true;
                window_ref_arr_h<M, false, 1>(acc, first, xz, A.HC.c[0], A.HC.c2[0], bet[2], 0.f);
                window_ref_arr_h<M, false, 1>(acc, first, yz_y, A.HC.c[1], A.HC.c2[1], bet[2], 0.f);
                window_ref_arr_h<M, true, 1>(acc, first, zz_z, A.HC.c[2], A.HC.c2[2], bet[2], 0.f);
                out[2] = __fadd_rn(acc, self[2]);
            } else {
                out[0] = self[0] + bet[0] * (window_fast_arr<M, T, true>(xx, A.HC.c[0]) + window_fast_arr<M, T, false>(xy_y, A.HC.c[1]) +
                                             window_fast_arr<M, T, false>(xz_z, A.HC.c[2]));
                out[1] = self[1] + bet[1] * (window_fast_arr<M, T, false>(xy, A.HC.c[0]) + window_fast_arr<M, T, true>(yy_y, A.HC.c[1]) +
                                             window_fast_arr<M, T, false>(yz_z, A.HC.c[2]));
                out[2] = self[2] + bet[2] * (window_fast_arr<M, T, false>(xz, A.HC.c[0]) + window_fast_arr<M, T, false>(yz_y, A.HC.c[1]) +
                                             window_fast_arr<M, T, true>(zz_z, A.HC.c[2]));
            }
        } else if (ARITH == OPESCI_ARITH_REFERENCE) {
            T acc = 0; bool first = true;
            window_ref_arr<M, T, true>(acc, first, xx, A.C.v[0][0]);
            window_ref_arr<M, T, false>(acc, first, xy_y, A.C.v[0][1]);
            window_ref_arr<M, T, false>(acc, first, xz_z, A.C.v[0][2]);
            out[0] = add_rn<T>(acc, self[0]);
            acc = 0; first = true;
            window_ref_arr<M, T, false>(acc, first, xy, A.C.v[1][0]);
            window_ref_arr<M, T, true>(acc, first, yy_y, A.C.v[1][1]);
            window_ref_arr<M, T, false>(acc, first, yz_z, A.C.v[1][2]);
            out[1] = add_rn<T>(acc, self[1]);
            acc = 0; first = true;
            window_ref_arr<M, T, false>(acc, first, xz, A.C.v[2][0]);
            window_ref_arr<M, T, false>(acc, first, yz_y, A.C.v[2][1]);
            window_ref_arr<M, T, true>(acc, first, zz_z, A.C.v[2][2]);
            out[2] = add_rn<T>(acc, self[2]);
        } else {
            out[0] = self[0] + (window_fast_arr<M, T, true>(xx, A.C.v[0][0]) + window_fast_arr<M, T, false>(xy_y, A.C.v[0][1]) +
                                window_fast_arr<M, T, false>(xz_z, A.C.v[0][2]));
            out[1] = self[1] + (window_fast_arr<M, T, false>(xy, A.C.v[1][0]) + window_fast_arr<M, T, true>(yy_y, A.C.v[1][1]) +
                                window_fast_arr<M, T, false>(yz_z, A.C.v[1][2]));
            out[2] = self[2] + (window_fast_arr<M, T, false>(xz, A.C.v[2][0]) + window_fast_arr<M, T, false>(yz_y, A.C.v[2][1]) +
                                window_fast_arr<M, T, true>(zz_z, A.C.v[2][2]));
        }
        if (ok) {
#pragma unroll
            for (int k = 0; k < 3; ++k) ((T *)A.F.f[F_U + k])[(long long)A.t1 * G.level + col + px] = out[k];
        }
        __syncthreads();
        if (tid == 0 && x + NB < xb) {
#pragma unroll
            for (int f = 0; f < 5; ++f) {
                mbar_arrive_expect_tx(&bars[f * NB + slot], BYTES);
                tma_load_3d((void *)(ring + (size_t)(f * NB + slot) * TE), tms[f], &bars[f * NB + slot], c0, c1, lvl1 + x + NB);
            }
        }
    }
}

}  // namespace opesci
