// hetero.cuh -- heterogeneous (`read`) mode of the staggered elastic path.
//
//   media_pointwise / media_averaged   derive beta, lambda, mu and beta1-3, mu12/13/23 from rho, vp, vs
//                                      opesci/staggeredgrid.py:522-598 (ranges of the patched oracle,
//                                      oracle/refgen/make_ref.py:patch_media_ranges)
//   stress_interior_h<SO,ARITH>        stress loop with `literal*G[...]*media[x][y][z]` terms
//   velocity_interior_h<SO,ARITH>      velocity loop, media beta1/beta2/beta3
//                                      opesci/staggeredgrid.py:293-359, 728-748
// fp32 only (the reference reader is float*, src/opesciIO.cpp:319).  ARITH_REFERENCE reproduces the emitted
// evaluation ((literal*G)*media, flat left-to-right sum) bit for bit; ARITH_FAST factors the media out of
// the windows.  Algorithmic traffic 104 B/point (72 + 8 media words).
#pragma once
#include "kernels.cuh"

namespace opesci {

struct HeteroCoefs {
    float c[3][OPESCI_MAX_M];    // c_k*dt/dx_d
    float c2[3][OPESCI_MAX_M];   // 2*c_k*dt/dx_d
};

struct MediaOut {
    float *m[OPESCI_MEDIA_COUNT];
};

// beta = 1.0F/rho; lambda = (pow(vp,2) - 2*pow(vs,2))*rho; mu = rho*pow(vs,2) -- pow(float,int) is double in
// C++, so lambda and mu are double expressions rounded once on assignment.  Whole array.
__global__ void media_pointwise(const float *__restrict__ rho, const float *__restrict__ vp, const float *__restrict__ vs,
                                MediaOut O, GridGeom G)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = blockIdx.z;
    if (z >= G.dim[2] || y >= G.dim[1]) return;
    const long long i = (long long)x * G.s[0] + (long long)y * G.s[1] + z;
    const float r = rho[i];
    const double p2 = __dmul_rn((double)vp[i], (double)vp[i]), s2 = __dmul_rn((double)vs[i], (double)vs[i]);
    O.m[OPESCI_MEDIA_BETA][i] = __fdiv_rn(1.0f, r);
    O.m[OPESCI_MEDIA_LAMBDA][i] = (float)__dmul_rn(__dsub_rn(p2, __dmul_rn(2.0, s2)), (double)r);
    O.m[OPESCI_MEDIA_MU][i] = (float)__dmul_rn((double)r, s2);
}

__device__ __forceinline__ float harmonic4(float a, float b, float c, float d)
{
    // 1.0F/(2.5e-1F/a + 2.5e-1F/b + 2.5e-1F/c + 2.5e-1F/d), left to right
    float s = __fadd_rn(__fdiv_rn(2.5e-1f, a), __fdiv_rn(2.5e-1f, b));
    s = __fadd_rn(s, __fdiv_rn(2.5e-1f, c));
    s = __fadd_rn(s, __fdiv_rn(2.5e-1f, d));
    return __fdiv_rn(1.0f, s);
}

// beta_d = 0.5*beta[+1 along d] + 0.5*beta;  mu_de = harmonic mean over the four cell corners.  [0,dim-1)^3.
__global__ void media_averaged(MediaOut O, GridGeom G)
{
    const int z = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    const int x = blockIdx.z;
    if (z >= G.dim[2] - 1 || y >= G.dim[1] - 1 || x >= G.dim[0] - 1) return;
    const long long sx = G.s[0], sy = G.s[1], sz = 1;
    const long long i = (long long)x * sx + (long long)y * sy + z;
    const float *beta = O.m[OPESCI_MEDIA_BETA], *mu = O.m[OPESCI_MEDIA_MU];
    const float b0 = __fmul_rn(5.0e-1f, beta[i]);
    O.m[OPESCI_MEDIA_BETA1][i] = __fadd_rn(__fmul_rn(5.0e-1f, beta[i + sx]), b0);
    O.m[OPESCI_MEDIA_BETA2][i] = __fadd_rn(__fmul_rn(5.0e-1f, beta[i + sy]), b0);
    O.m[OPESCI_MEDIA_BETA3][i] = __fadd_rn(__fmul_rn(5.0e-1f, beta[i + sz]), b0);
    O.m[OPESCI_MEDIA_MU12][i] = harmonic4(mu[i], mu[i + sy], mu[i + sx], mu[i + sx + sy]);
    O.m[OPESCI_MEDIA_MU13][i] = harmonic4(mu[i], mu[i + sz], mu[i + sx], mu[i + sx + sz]);
    O.m[OPESCI_MEDIA_MU23][i] = harmonic4(mu[i], mu[i + sz], mu[i + sy], mu[i + sy + sz]);
}

// acc (+)= (c*g)*med, separate roundings
__device__ __forceinline__ void term_h(float &acc, bool &first, float c, float g, float med)
{
    const float prod = __fmul_rn(__fmul_rn(c, g), med);
    acc = first ? prod : __fadd_rn(acc, prod);
    first = false;
}
// one window in the printer's order; NV = 2 emits the lambda term then the mu term (coefficient c2) per offset
template <int M, bool FWD, int NV>
__device__ __forceinline__ void window_ref_h(float &acc, bool &first, const float *__restrict__ g, long long stride,
                                             const float *c, const float *c2, float med, float med2)
{
#define OPESCI_EMIT(o, sgn, k)                                              \
    {                                                                       \
        const float gv = g[(long long)(o) * stride];                        \
        term_h(acc, first, (sgn) * c[k], gv, med);                          \
        if (NV == 2) term_h(acc, first, (sgn) * c2[k], gv, med2);           \
    }
    if (FWD) {
#pragma unroll
        for (int o = 1; o <= M; ++o) OPESCI_EMIT(o, 1.0f, o - 1)
#pragma unroll
        for (int o = 1; o <= M - 1; ++o) OPESCI_EMIT(-o, -1.0f, o)
        OPESCI_EMIT(0, -1.0f, 0)
    } else {
#pragma unroll
        for (int o = 1; o <= M - 1; ++o) OPESCI_EMIT(o, 1.0f, o)
#pragma unroll
        for (int o = 1; o <= M; ++o) OPESCI_EMIT(-o, -1.0f, o - 1)
        OPESCI_EMIT(0, 1.0f, 0)
    }
#undef OPESCI_EMIT
}

template <int SO, int ARITH>
__global__ void __launch_bounds__(256)
stress_interior_h(FieldPtrs F, MediaPtrs MD, GridGeom G, HeteroCoefs C, int t0, int t1, int x0 = SO / 2, int z0 = SO / 2)
{
    constexpr int M = SO / 2;
    const int z = z0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = M + blockIdx.y * blockDim.y + threadIdx.y;
    const int x = x0 + blockIdx.z;
    if (z >= G.dim[2] - M || y >= G.dim[1] - M) return;
    const long long p = (long long)x * G.s[0] + (long long)y * G.s[1] + z;
    const long long r = (long long)t0 * G.level + p, w = (long long)t1 * G.level + p;
    const float *U = (const float *)F.f[F_U] + r, *V = (const float *)F.f[F_V] + r, *W = (const float *)F.f[F_W] + r;
    const long long st[3] = {G.s[0], G.s[1], 1};
    const float *vel[3] = {U, V, W};
    const float lam = MD.m[OPESCI_MEDIA_LAMBDA][p], mu = MD.m[OPESCI_MEDIA_MU][p];
    if (ARITH == OPESCI_ARITH_REFERENCE) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float *Tn = (float *)F.f[F_TXX + a];
            float acc = Tn[r];
            bool first = false;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                if (d == a) window_ref_h<M, false, 2>(acc, first, vel[d], st[d], C.c[d], C.c2[d], lam, mu);
                else window_ref_h<M, false, 1>(acc, first, vel[d], st[d], C.c[d], C.c2[d], lam, mu);
            }
            Tn[w] = acc;
        }
    } else {
        float dv[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) dv[d] = window_fast<M, float, false>(vel[d], st[d], C.c[d]);
        const float tr = lam * (dv[0] + dv[1] + dv[2]);
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            float *Tn = (float *)F.f[F_TXX + a];
            Tn[w] = Tn[r] + (tr + 2.0f * mu * dv[a]);
        }
    }
    // shear stresses Txy (mu12), Tyz (mu23), Txz (mu13): D_b V_a then D_a V_b, forward windows
    {
        const int A[3] = {0, 1, 0}, B[3] = {1, 2, 2};
        const int MU[3] = {OPESCI_MEDIA_MU12, OPESCI_MEDIA_MU23, OPESCI_MEDIA_MU13};
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            float *Ts = (float *)F.f[F_TXY + s];
            const float ms = MD.m[MU[s]][p];
            if (ARITH == OPESCI_ARITH_REFERENCE) {
                float acc = Ts[r];
                bool first = false;
                window_ref_h<M, true, 1>(acc, first, vel[A[s]], st[B[s]], C.c[B[s]], C.c2[B[s]], ms, ms);
                window_ref_h<M, true, 1>(acc, first, vel[B[s]], st[A[s]], C.c[A[s]], C.c2[A[s]], ms, ms);
                Ts[w] = acc;
            } else {
                Ts[w] = Ts[r] + ms * (window_fast<M, float, true>(vel[A[s]], st[B[s]], C.c[B[s]]) +
                                      window_fast<M, float, true>(vel[B[s]], st[A[s]], C.c[A[s]]));
            }
        }
    }
}

template <int SO, int ARITH>
__global__ void __launch_bounds__(256)
velocity_interior_h(FieldPtrs F, MediaPtrs MD, GridGeom G, HeteroCoefs C, int t0, int t1)
{
    constexpr int M = SO / 2;
    const int z = M + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = M + blockIdx.y * blockDim.y + threadIdx.y;
    const int x = M + blockIdx.z;
    if (z >= G.dim[2] - M || y >= G.dim[1] - M) return;
    const long long p = (long long)x * G.s[0] + (long long)y * G.s[1] + z;
    const long long r = (long long)t0 * G.level + p, w = (long long)t1 * G.level + p;
    const long long st[3] = {G.s[0], G.s[1], 1};
    const int opnd[3][3] = {{F_TXX, F_TXY, F_TXZ}, {F_TXY, F_TYY, F_TYZ}, {F_TXZ, F_TYZ, F_TZZ}};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float *Va = (float *)F.f[F_U + a];
        const float b = MD.m[OPESCI_MEDIA_BETA1 + a][p];
        if (ARITH == OPESCI_ARITH_REFERENCE) {
            float acc = 0;
            bool first = true;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float *g = (const float *)F.f[opnd[a][d]] + w;
                if (d == a) window_ref_h<M, true, 1>(acc, first, g, st[d], C.c[d], C.c2[d], b, b);
                else window_ref_h<M, false, 1>(acc, first, g, st[d], C.c[d], C.c2[d], b, b);
            }
            Va[w] = __fadd_rn(acc, Va[r]);
        } else {
            float acc = 0;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float *g = (const float *)F.f[opnd[a][d]] + w;
                acc += (d == a) ? window_fast<M, float, true>(g, st[d], C.c[d]) : window_fast<M, float, false>(g, st[d], C.c[d]);
            }
            Va[w] = Va[r] + b * acc;
        }
    }
}

}  // namespace opesci
