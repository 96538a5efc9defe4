/*
 * opesci_oracle.c -- CPU restatement of opesci-fd's generated time-stepping code.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: it may be built, loaded or
 * executed only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  It is
 * never on the product path (the product is opesci_fd_b200/csrc, CUDA only).
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement
 * BIT-FOR-BIT against the reference's own generated OpenMP C++ (oracle/_ref, built by
 * oracle/refgen/make_ref.py from /root/reference) for so = 2..12, fp32 and fp64, staggered
 * and regular grids, and against the committed golden fixtures in tests/golden/.
 *
 * It exports the same C ABI as the CUDA library (include/opesci_b200.h) so that one ctypes
 * binding drives both.  What is restated (reference file:line):
 *   - stress / velocity interior updates          opesci/staggeredgrid.py:728-748,
 *                                                  opesci/regulargrid.py:566-590, 601-619
 *   - stress free-surface / ghost loops            opesci/staggeredgrid.py:750-813, opesci/fields.py:294-381
 *   - velocity free-surface / ghost loops          opesci/staggeredgrid.py:815-864, opesci/fields.py:192-261
 *   - time-level rotation                          opesci/regulargrid.py:408-433
 *   - initialisation + initial BC pass             opesci/staggeredgrid.py:612-659, 866-879
 *   - regular-grid update and second initialisation opesci/regulargrid.py:498-564, 592-619
 *   - L2 convergence                               opesci/staggeredgrid.py:892-945, opesci/regulargrid.py:650-700
 *   - allocate / store / free                      opesci/regulargrid.py:445-453, 474-496, 621-634
 *
 * Arithmetic contract (what makes bit-exactness possible): the generator prints every
 * update as a flat left-to-right sum of `literal*G[...]` products (expand=True,
 * eval_const=True, opesci/regulargrid.py:329-342) with float literals
 * (opesci/codeprinter.py:30,63); gcc without -ffast-math evaluates that as one rounded
 * multiply and one rounded add per term.  Build this file with -ffp-contract=off.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/opesci_b200.h"
#include "../include/opesci_slab.h"
#include "opesci_oracle_slab.h"

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ model ---- */
enum { TERM_MUL = 0, TERM_PLUS = 1, TERM_MINUS = 2 };

/* media factors of one emitted term in heterogeneous (`read`) mode, evaluated left to right like
 * the printed C:  v = coef*G;  then
 *   MK_A       v*A                       e.g. 6.3e-2F*U[..]*lambda[x][y][z]
 *   MK_A_DIV   v*A/D                     D = (da*lambda + db*mu) of the equation
 *   MK_AB_DIV  v*A*B/D
 *   MK_SQ_DIV  v*pow(B, 2)/D             pow(float,int) is double: the term, and from it on the
 *                                        running sum, are double
 *   MK_RATIO   v*A/B                     e.g. V[..]*mu12[2][y][z]/mu12[1][y][z]                    */
enum { MK_NONE = 0, MK_A, MK_A_DIV, MK_AB_DIV, MK_SQ_DIV, MK_RATIO };

typedef struct {
    int kind;      /* TERM_MUL: acc += coef*G ; TERM_PLUS/MINUS: acc +/-= G (exact self terms) */
    int field;
    int level;     /* index into the current {t0,t1,t2} triple */
    long off;      /* element offset inside one time level */
    float coef;    /* signed float literal */
    int mk, ma, mb;        /* media kind and the OPESCI_MEDIA_* arrays A, B */
    long moffa, moffb;     /* their element offsets relative to the output cell */
} Term;

#define MAX_TERMS 96
typedef struct {
    int out, out_level, nterm;
    float da, db;  /* denominator literals of the MK_*_DIV terms */
    long doff;     /* element offset of the lambda/mu cell of the denominator */
    Term term[MAX_TERMS];
} Equation;

typedef struct {
    OpesciB200Params p;
    int m;
    long s[3];              /* strides of axes x,y,z inside one level */
    size_t level_elems;
    Equation stress[6], velocity[3], acoustic, acoustic_init;
    Equation lev_stress_eq[3][3];    /* [face axis d][normal stress e] */
    Equation lev_vel_eq[3][3][2];    /* [face axis d][velocity a][side] */
    double *tables;         /* private copy of every 1-D table */
    int configured;
    OpesciSlab slab;        /* x-slab of this rank; p.dim[0] holds the LOCAL plane count after configure */
    int gdim1;              /* global dim1 */
    float *media[OPESCI_MEDIA_COUNT];   /* heterogeneous mode: derived arrays on the local slab */
} Model;

static Model g_model;
static char g_err[512] = "";
static double g_loop_seconds = 0.0;
static opesci_oracle_exchange_fn g_exchange = NULL;
static void *g_exchange_user = NULL;

void opesci_oracle_set_exchange(opesci_oracle_exchange_fn fn, void *user)
{
    g_exchange = fn;
    g_exchange_user = user;
}

static int fail(const char *msg)
{
    snprintf(g_err, sizeof g_err, "%s", msg);
    return 1;
}

const char *opesci_b200_last_error(void) { return g_err; }
int opesci_b200_is_cuda(void) { return 0; }

/* field ids in struct order (opesci/staggeredgrid.py:69-71) */
enum { F_U = 0, F_V, F_W, F_TXX, F_TYY, F_TZZ, F_TXY, F_TYZ, F_TXZ };
static const int NORMAL_OF_AXIS[3] = {F_TXX, F_TYY, F_TZZ};
static const int VEL_OF_AXIS[3] = {F_U, F_V, F_W};
/* shear field for an axis pair */
static int shear_of(int a, int b)
{
    if (a > b) { int t = a; a = b; b = t; }
    if (a == 0 && b == 1) return F_TXY;
    if (a == 1 && b == 2) return F_TYZ;
    return F_TXZ;
}

static void push(Equation *eq, int kind, int field, int level, long off, float coef)
{
    Term *t = &eq->term[eq->nterm++];
    t->kind = kind; t->field = field; t->level = level; t->off = off; t->coef = coef;
    t->mk = MK_NONE; t->ma = t->mb = 0; t->moffa = t->moffb = 0;
}

/* same, with media factors */
static void push_m(Equation *eq, int kind, int field, int level, long off, float coef,
                   int mk, int ma, long moffa, int mb, long moffb)
{
    push(eq, kind, field, level, off, coef);
    Term *t = &eq->term[eq->nterm - 1];
    t->mk = mk; t->ma = ma; t->mb = mb; t->moffa = moffa; t->moffb = moffb;
}

/* First-derivative window of G along one axis, in the order the printer emits it
 * (verified on the generated files): positive offsets ascending, negative
 * offsets by ascending magnitude, offset 0 last.
 *   forward  (F staggered along the axis, G not; opesci/fields.py:127-128):
 *            sum_k c_k (G[i+k] - G[i-k+1])    offsets -m+1 .. m
 *   backward (G staggered along the axis, F not; opesci/fields.py:131-132):
 *            sum_k c_k (G[i+k-1] - G[i-k])    offsets -m .. m-1
 * c[k-1] carries c_k*dt/dx*material, sign included. */
static void push_window(Equation *eq, int field, int level, long stride, int m, const float *c, int forward)
{
    if (forward) {
        for (int o = 1; o <= m; ++o) push(eq, TERM_MUL, field, level, o * stride, c[o - 1]);
        for (int o = 1; o <= m - 1; ++o) push(eq, TERM_MUL, field, level, -o * stride, -c[o]);
        push(eq, TERM_MUL, field, level, 0, -c[0]);
    } else {
        for (int o = 1; o <= m - 1; ++o) push(eq, TERM_MUL, field, level, o * stride, c[o]);
        for (int o = 1; o <= m; ++o) push(eq, TERM_MUL, field, level, -o * stride, -c[o - 1]);
        push(eq, TERM_MUL, field, level, 0, c[0]);
    }
}

/* opesci/staggeredgrid.py:728-748 via regulargrid.py:601-619; term order = alphabetical by
 * field name (Txx<Txy<Txz<Tyy<Tyz<Tzz<U<V<W), so the self term leads the stress sums and
 * closes the velocity sums. */
static void build_staggered(Model *M)
{
    const OpesciB200Params *p = &M->p;
    const int m = M->m;
    /* normal stresses: T_aa[t1] = T_aa[t0] + sum_d coef(a,d) * D_d V_d  (backward windows) */
    for (int a = 0; a < 3; ++a) {
        Equation *eq = &M->stress[a];
        eq->out = NORMAL_OF_AXIS[a]; eq->out_level = 1; eq->nterm = 0;
        push(eq, TERM_PLUS, eq->out, 0, 0, 1.0f);
        for (int d = 0; d < 3; ++d)
            push_window(eq, VEL_OF_AXIS[d], 0, M->s[d], m, p->c_stress_normal[a][d], 0);
    }
    /* shear stresses, emitted order Txy, Tyz, Txz: T_ab += mu*(D_b V_a + D_a V_b), forward windows */
    static const int PAIR[3][2] = {{0, 1}, {1, 2}, {0, 2}};
    for (int s = 0; s < 3; ++s) {
        const int a = PAIR[s][0], b = PAIR[s][1];
        Equation *eq = &M->stress[3 + s];
        eq->out = shear_of(a, b); eq->out_level = 1; eq->nterm = 0;
        push(eq, TERM_PLUS, eq->out, 0, 0, 1.0f);
        push_window(eq, VEL_OF_AXIS[a], 0, M->s[b], m, p->c_stress_shear[s][0], 1);
        push_window(eq, VEL_OF_AXIS[b], 0, M->s[a], m, p->c_stress_shear[s][1], 1);
    }
    /* velocities: V_a[t1] = sum_d coef(a,d) * D_d T_ad[t1] + V_a[t0] */
    for (int a = 0; a < 3; ++a) {
        Equation *eq = &M->velocity[a];
        eq->out = VEL_OF_AXIS[a]; eq->out_level = 1; eq->nterm = 0;
        for (int d = 0; d < 3; ++d) {
            const int g = (d == a) ? NORMAL_OF_AXIS[a] : shear_of(a, d);
            push_window(eq, g, 1, M->s[d], m, p->c_velocity[a][d], d == a);
        }
        push(eq, TERM_PLUS, eq->out, 0, 0, 1.0f);
    }
}


/* Levander free surface, so == 4 (m == 2) only.
 * Stress (opesci/fields.py:313-353): on face d, T_ee (e != d) is recomputed from level t0:
 *   T_ee[t1] = 1.0F*T_ee[t0] + sum_{f != d} lev_stress[d][e][f][.] * D_f V_f   (backward windows)
 * Velocity (opesci/fields.py:208-242), all operands at level t1:
 *   normal     V_d[n]   = V_d[n+-1] +/- sum_{e != d} a_e (V_e[e:0] - V_e[e:-1])   at plane b
 *   tangential V_e[n]   = 2 V_e[b] - V_e[b+-1] +/- g (D+_e V_d[n'] - D+_e V_d[b])
 * Term order = alphabetical by field name, then lexicographic by index with "+1 before 0" on
 * loop indices and the nearer plane first (verified against the generated files). */
static void build_levander(Model *M)
{
    const OpesciB200Params *p = &M->p;
    const int m = M->m;
    for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) {
            Equation *eq = &M->lev_stress_eq[d][e];
            eq->out = NORMAL_OF_AXIS[e]; eq->out_level = 1; eq->nterm = 0;
            if (e == d) continue;
            push(eq, TERM_PLUS, eq->out, 0, 0, 1.0f);
            for (int f = 0; f < 3; ++f)
                if (f != d) push_window(eq, VEL_OF_AXIS[f], 0, M->s[f], m, p->lev_stress[d][e][f], 0);
        }
    for (int d = 0; d < 3; ++d)
        for (int a = 0; a < 3; ++a)
            for (int side = 0; side < 2; ++side) {
                Equation *eq = &M->lev_vel_eq[d][a][side];
                const long sd = M->s[d];
                eq->out = VEL_OF_AXIS[a]; eq->out_level = 0; eq->nterm = 0;
                if (a == d) {
                    /* target plane n: low n = m-1 (reads plane m), high n = dim-m-1 (reads itself
                     * for the tangential differences and plane n-1 for the self term) */
                    const float sgn = side == 0 ? 1.0f : -1.0f;
                    const long plane = side == 0 ? sd : 0;       /* offset of the plane the differences live on */
                    const long selfoff = side == 0 ? sd : -sd;
                    for (int g = 0; g < 3; ++g) {
                        if (g == d) {
                            push(eq, TERM_PLUS, VEL_OF_AXIS[d], 0, selfoff, 1.0f);
                        } else {
                            const float c = p->lev_vnormal[d][g];
                            push(eq, TERM_MUL, VEL_OF_AXIS[g], 0, plane - M->s[g], -sgn * c);
                            push(eq, TERM_MUL, VEL_OF_AXIS[g], 0, plane, sgn * c);
                        }
                    }
                } else {
                    /* tangential field V_a on face d; target plane n: low n = m-1 (b = n+1),
                     * high n = dim-m (b' = n-1) */
                    const int e = a;
                    const float g = p->lev_vtang[d][e];
                    const float sgn = side == 0 ? 1.0f : -1.0f;
                    const long se = M->s[e];
                    /* planes of V_d used: low: n (=b-1) and n+1 (=b); high: n-1 (=b') and n-2 */
                    const long pl0 = side == 0 ? 0 : -sd, pl1 = side == 0 ? sd : -2 * sd;
                    /* self planes: low 2V[b]-V[b+1]; high 2V[b']-V[b'-1] */
                    const long sf0 = side == 0 ? sd : -sd, sf1 = side == 0 ? 2 * sd : -2 * sd;
                    if (d < e) {
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl0 + se, sgn * g);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl0, -sgn * g);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl1 + se, -sgn * g);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl1, sgn * g);
                        push(eq, TERM_MUL, VEL_OF_AXIS[e], 0, sf0, 2.0f);
                        push(eq, TERM_MINUS, VEL_OF_AXIS[e], 0, sf1, 1.0f);
                    } else {
                        push(eq, TERM_MUL, VEL_OF_AXIS[e], 0, sf0, 2.0f);
                        push(eq, TERM_MINUS, VEL_OF_AXIS[e], 0, sf1, 1.0f);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, se + pl0, sgn * g);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, se + pl1, -sgn * g);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl0, -sgn * g);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl1, sgn * g);
                    }
                }
            }
}


/* ---- heterogeneous (`read`) mode ---------------------------------------------------------
 * opesci/staggeredgrid.py:147-148, 284-359: beta, lambda, mu in the PDEs become per-cell arrays,
 * half-index accesses map to the effective arrays (beta1-3 for U,V,W; mu12/mu23/mu13 for
 * Txy/Tyz/Txz), and every emitted term becomes `literal*G[...]*media[x][y][z]` with the literal
 * c_k*dt/dx_d (2*c_k*dt/dx_d for the mu terms of T_dd).  Term order as in the homogeneous case; where
 * one operand carries a lambda and a mu term, lambda comes first (verified on the generated files). */
typedef struct { const float *c; int mk, ma, mb; } Variant;

static void push_window_v(Equation *eq, int field, int level, long stride, int m, int forward,
                          const Variant *v, int nv, long moff)
{
#define EMIT(o, sign, k)                                                                          \
    for (int j_ = 0; j_ < nv; ++j_)                                                               \
        push_m(eq, TERM_MUL, field, level, (long)(o) * stride, (sign) * v[j_].c[k], v[j_].mk, v[j_].ma, moff, v[j_].mb, moff)
    if (forward) {
        for (int o = 1; o <= m; ++o) EMIT(o, 1.0f, o - 1);
        for (int o = 1; o <= m - 1; ++o) EMIT(-o, -1.0f, o);
        EMIT(0, -1.0f, 0);
    } else {
        for (int o = 1; o <= m - 1; ++o) EMIT(o, 1.0f, o);
        for (int o = 1; o <= m; ++o) EMIT(-o, -1.0f, o - 1);
        EMIT(0, 1.0f, 0);
    }
#undef EMIT
}

static int mu_of_pair(int a, int b)
{
    if (a > b) { int t = a; a = b; b = t; }
    if (a == 0 && b == 1) return OPESCI_MEDIA_MU12;
    if (a == 1 && b == 2) return OPESCI_MEDIA_MU23;
    return OPESCI_MEDIA_MU13;
}

static void build_staggered_hetero(Model *M)
{
    const OpesciB200Params *p = &M->p;
    const int m = M->m;
    for (int a = 0; a < 3; ++a) {
        Equation *eq = &M->stress[a];
        eq->out = NORMAL_OF_AXIS[a]; eq->out_level = 1; eq->nterm = 0;
        push(eq, TERM_PLUS, eq->out, 0, 0, 1.0f);
        for (int d = 0; d < 3; ++d) {
            const Variant v[2] = {{p->h_c[d], MK_A, OPESCI_MEDIA_LAMBDA, 0}, {p->h_c2[d], MK_A, OPESCI_MEDIA_MU, 0}};
            push_window_v(eq, VEL_OF_AXIS[d], 0, M->s[d], m, 0, v, d == a ? 2 : 1, 0);
        }
    }
    static const int PAIR[3][2] = {{0, 1}, {1, 2}, {0, 2}};
    for (int s = 0; s < 3; ++s) {
        const int a = PAIR[s][0], b = PAIR[s][1];
        Equation *eq = &M->stress[3 + s];
        eq->out = shear_of(a, b); eq->out_level = 1; eq->nterm = 0;
        push(eq, TERM_PLUS, eq->out, 0, 0, 1.0f);
        const Variant vb = {p->h_c[b], MK_A, mu_of_pair(a, b), 0}, va = {p->h_c[a], MK_A, mu_of_pair(a, b), 0};
        push_window_v(eq, VEL_OF_AXIS[a], 0, M->s[b], m, 1, &vb, 1, 0);
        push_window_v(eq, VEL_OF_AXIS[b], 0, M->s[a], m, 1, &va, 1, 0);
    }
    for (int a = 0; a < 3; ++a) {
        Equation *eq = &M->velocity[a];
        eq->out = VEL_OF_AXIS[a]; eq->out_level = 1; eq->nterm = 0;
        for (int d = 0; d < 3; ++d) {
            const int g = (d == a) ? NORMAL_OF_AXIS[a] : shear_of(a, d);
            const Variant v = {p->h_c[d], MK_A, OPESCI_MEDIA_BETA1 + a, 0};
            push_window_v(eq, g, 1, M->s[d], m, d == a, &v, 1, 0);
        }
        push(eq, TERM_PLUS, eq->out, 0, 0, 1.0f);
    }
}

/* Levander free surface with per-cell media (so == 4), as emitted by opesci/fields.py:208-242, 313-353
 * through sympy's solve() in this container (sympy 1.14; oracle/refgen) -- the patched oracle of
 * SURVEY.md 8c.  P_d = product of the two spacings other than dx_d.
 *  stress, face d, T_ee (e != d), D = 12P_d*lambda + 24P_d*mu at the cell:
 *    12P_d*T_ee[t0]*lambda/D + 24P_d*T_ee[t0]*mu/D
 *    + window(V_e along e) with 48P_d c_k dt/dx_e, each offset as `*lambda*mu/D` then `*pow(mu,2)/D`
 *    + window(V_f along f), f the third axis, with 24P_d c_k dt/dx_f as `*lambda*mu/D`
 *    (fields alphabetical; = T + [4mu(lambda+mu)/(lambda+2mu)] D_e V_e + [2 lambda mu/(lambda+2mu)] D_f V_f)
 *  velocity normal ghost V_d[n], D = P_d*lambda + 2P_d*mu at the boundary-plane cell:
 *    P_d*V_d[self]*lambda/D + 2P_d*V_d[self]*mu/D  -/+ P_g*(V_g[0] - V_g[-1])*lambda/D   (g != d, alphabetical)
 *  velocity tangential ghost V_e[n], r = mu_de[b]/mu_de[n']:
 *    V_e[b] + V_e[b]*r - V_e[b+-1]*r  +/- g (V_d[n][+e] - V_d[n]) -/+ g (V_d[b][+e] - V_d[b])*r      */
static void build_levander_hetero(Model *M)
{
    const OpesciB200Params *p = &M->p;
    const int m = M->m;
    const int LAM = OPESCI_MEDIA_LAMBDA, MU = OPESCI_MEDIA_MU;
    for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) {
            Equation *eq = &M->lev_stress_eq[d][e];
            eq->out = NORMAL_OF_AXIS[e]; eq->out_level = 1; eq->nterm = 0;
            if (e == d) continue;
            eq->da = p->h_lev_den[d][0]; eq->db = p->h_lev_den[d][1]; eq->doff = 0;
            push_m(eq, TERM_MUL, eq->out, 0, 0, eq->da, MK_A_DIV, LAM, 0, 0, 0);
            push_m(eq, TERM_MUL, eq->out, 0, 0, eq->db, MK_A_DIV, MU, 0, 0, 0);
            for (int f = 0; f < 3; ++f) {
                if (f == d) continue;
                if (f == e) {
                    const Variant v[2] = {{p->h_lev_own[d][f], MK_AB_DIV, LAM, MU}, {p->h_lev_own[d][f], MK_SQ_DIV, 0, MU}};
                    push_window_v(eq, VEL_OF_AXIS[f], 0, M->s[f], m, 0, v, 2, 0);
                } else {
                    const Variant v = {p->h_lev_oth[d][f], MK_AB_DIV, LAM, MU};
                    push_window_v(eq, VEL_OF_AXIS[f], 0, M->s[f], m, 0, &v, 1, 0);
                }
            }
        }
    for (int d = 0; d < 3; ++d)
        for (int a = 0; a < 3; ++a)
            for (int side = 0; side < 2; ++side) {
                Equation *eq = &M->lev_vel_eq[d][a][side];
                const long sd = M->s[d];
                const float sgn = side == 0 ? 1.0f : -1.0f;
                eq->out = VEL_OF_AXIS[a]; eq->out_level = 0; eq->nterm = 0;
                if (a == d) {
                    const long plane = side == 0 ? sd : 0;
                    const long selfoff = side == 0 ? sd : -sd;
                    eq->da = p->h_vn[d][0]; eq->db = p->h_vn[d][1]; eq->doff = plane;
                    for (int g = 0; g < 3; ++g) {
                        if (g == d) {
                            push_m(eq, TERM_MUL, VEL_OF_AXIS[d], 0, selfoff, eq->da, MK_A_DIV, LAM, plane, 0, 0);
                            push_m(eq, TERM_MUL, VEL_OF_AXIS[d], 0, selfoff, eq->db, MK_A_DIV, MU, plane, 0, 0);
                        } else {
                            const float c = p->h_vn[g][0];
                            push_m(eq, TERM_MUL, VEL_OF_AXIS[g], 0, plane - M->s[g], -sgn * c, MK_A_DIV, LAM, plane, 0, 0);
                            push_m(eq, TERM_MUL, VEL_OF_AXIS[g], 0, plane, sgn * c, MK_A_DIV, LAM, plane, 0, 0);
                        }
                    }
                } else {
                    const int e = a;
                    const float g = p->lev_vtang[d][e];
                    const long se = M->s[e];
                    const long pl0 = side == 0 ? 0 : -sd, pl1 = side == 0 ? sd : -2 * sd;
                    const long sf0 = side == 0 ? sd : -sd, sf1 = side == 0 ? 2 * sd : -2 * sd;
                    const int mu = mu_of_pair(d, e);
                    const long ra = pl1, rb = pl0;   /* r = mu_de[plane of pl1] / mu_de[plane of pl0] */
#define RATIO MK_RATIO, mu, ra, mu, rb
                    if (d < e) {
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl0 + se, sgn * g);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl0, -sgn * g);
                        push_m(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl1 + se, -sgn * g, RATIO);
                        push_m(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl1, sgn * g, RATIO);
                        push(eq, TERM_PLUS, VEL_OF_AXIS[e], 0, sf0, 1.0f);
                        push_m(eq, TERM_PLUS, VEL_OF_AXIS[e], 0, sf0, 1.0f, RATIO);
                        push_m(eq, TERM_MINUS, VEL_OF_AXIS[e], 0, sf1, 1.0f, RATIO);
                    } else {
                        push(eq, TERM_PLUS, VEL_OF_AXIS[e], 0, sf0, 1.0f);
                        push_m(eq, TERM_PLUS, VEL_OF_AXIS[e], 0, sf0, 1.0f, RATIO);
                        push_m(eq, TERM_MINUS, VEL_OF_AXIS[e], 0, sf1, 1.0f, RATIO);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, se + pl0, sgn * g);
                        push_m(eq, TERM_MUL, VEL_OF_AXIS[d], 0, se + pl1, -sgn * g, RATIO);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl0, -sgn * g);
                        push_m(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl1, sgn * g, RATIO);
                    }
#undef RATIO
                }
            }
}

/* a11, opesci/staggeredgrid.py:522-598: derive the nine media arrays from rho, vp, vs on the local
 * slab, all in float except `pow(x, 2)` which C++ evaluates in double (std::pow(float,int)).
 * Ranges of the patched oracle (oracle/refgen/make_ref.py:patch_media_ranges). */
static int media_effective(Model *M)
{
    const OpesciB200Params *p = &M->p;
    const int D1 = p->dim[0], D2 = p->dim[1], D3 = p->dim[2];
    const size_t n = M->level_elems;
    if (!p->rho || !p->vp || !p->vs) return fail("hetero: rho/vp/vs missing");
    if (p->media_plane0 > M->slab.L0 || p->media_plane0 + p->media_nplanes < M->slab.L1)
        return fail("hetero: rho/vp/vs do not cover this rank's planes");
    const size_t shift = (size_t)(M->slab.L0 - p->media_plane0) * (size_t)M->s[0];
    const float *rho = p->rho + shift, *vp = p->vp + shift, *vs = p->vs + shift;
    for (int k = 0; k < OPESCI_MEDIA_COUNT; ++k) {
        free(M->media[k]);
        M->media[k] = (float *)calloc(n, sizeof(float));
        if (!M->media[k]) return fail("oracle: out of memory");
    }
    float *beta = M->media[OPESCI_MEDIA_BETA], *lam = M->media[OPESCI_MEDIA_LAMBDA], *mu = M->media[OPESCI_MEDIA_MU];
    for (size_t i = 0; i < n; ++i) {
        beta[i] = 1.0F / rho[i];
        lam[i] = (float)((pow((double)vp[i], 2) - 2 * pow((double)vs[i], 2)) * (double)rho[i]);
        mu[i] = (float)((double)rho[i] * pow((double)vs[i], 2));
    }
    const long sx = M->s[0], sy = M->s[1], sz = 1;
    for (int x = 0; x < D1 - 1; ++x)
        for (int y = 0; y < D2 - 1; ++y)
            for (int z = 0; z < D3 - 1; ++z) {
                const long i = (long)x * sx + (long)y * sy + z;
                M->media[OPESCI_MEDIA_BETA1][i] = 5.0e-1F * beta[i + sx] + 5.0e-1F * beta[i];
                M->media[OPESCI_MEDIA_BETA2][i] = 5.0e-1F * beta[i + sy] + 5.0e-1F * beta[i];
                M->media[OPESCI_MEDIA_BETA3][i] = 5.0e-1F * beta[i + sz] + 5.0e-1F * beta[i];
                M->media[OPESCI_MEDIA_MU12][i] = 1.0F / (2.5e-1F / mu[i] + 2.5e-1F / mu[i + sy] + 2.5e-1F / mu[i + sx] + 2.5e-1F / mu[i + sx + sy]);
                M->media[OPESCI_MEDIA_MU13][i] = 1.0F / (2.5e-1F / mu[i] + 2.5e-1F / mu[i + sz] + 2.5e-1F / mu[i + sx] + 2.5e-1F / mu[i + sx + sz]);
                M->media[OPESCI_MEDIA_MU23][i] = 1.0F / (2.5e-1F / mu[i] + 2.5e-1F / mu[i + sz] + 2.5e-1F / mu[i + sy] + 2.5e-1F / mu[i + sy + sz]);
            }
    return 0;
}

/* opesci/regulargrid.py:592-619 (update) and 530-564 (second initialisation).  Central
 * second-derivative windows; emitted order per axis: +1..+m, -1..-m; centre last. */
static void build_regular(Model *M)
{
    const OpesciB200Params *p = &M->p;
    const int m = M->m;
    Equation *eq = &M->acoustic;
    eq->out = 0; eq->out_level = 2; eq->nterm = 0;
    push(eq, TERM_MINUS, 0, 0, 0, 1.0f);
    for (int d = 0; d < 3; ++d) {
        int present = 0;
        for (int k = 0; k < m; ++k) present |= (p->ac_coef[d][k] != 0.0f);
        if (!present) continue;
        for (int o = 1; o <= m; ++o) push(eq, TERM_MUL, 0, 1, o * M->s[d], p->ac_coef[d][o - 1]);
        for (int o = 1; o <= m; ++o) push(eq, TERM_MUL, 0, 1, -o * M->s[d], p->ac_coef[d][o - 1]);
    }
    push(eq, TERM_MUL, 0, 1, 0, p->ac_centre);
    /* level t1 := `1.0F*v*dt` + halved stencil of level t0 (the constant is added first) */
    eq = &M->acoustic_init;
    eq->out = 0; eq->out_level = 1; eq->nterm = 0;
    for (int d = 0; d < 3; ++d) {
        int present = 0;
        for (int k = 0; k < m; ++k) present |= (p->ac_init_coef[d][k] != 0.0f);
        if (!present) continue;
        for (int o = 1; o <= m; ++o) push(eq, TERM_MUL, 0, 0, o * M->s[d], p->ac_init_coef[d][o - 1]);
        for (int o = 1; o <= m; ++o) push(eq, TERM_MUL, 0, 0, -o * M->s[d], p->ac_init_coef[d][o - 1]);
    }
    push(eq, TERM_MUL, 0, 0, 0, p->ac_init_centre);
}

int opesci_b200_configure(const OpesciB200Params *params)
{
    Model *M = &g_model;
    if (!params || params->struct_size != sizeof(OpesciB200Params))
        return fail("opesci_b200_configure: struct_size mismatch");
    if (params->so < 2 || params->so > 12 || (params->so & 1))
        return fail("opesci_b200_configure: so must be even, 2..12");
    free(M->tables);
    memset(M, 0, sizeof *M);
    M->p = *params;
    M->m = params->so / 2;
    M->gdim1 = params->dim[0];
    {
        const int nr = params->slab_nranks > 1 ? params->slab_nranks : 1;
        const int need = opesci_slab_need(params->kind == OPESCI_KIND_REGULAR_ACOUSTIC, params->so);
        if (opesci_slab_make(&M->slab, nr > 1 ? params->slab_rank : 0, nr, params->dim[0], M->m, OPESCI_SLAB_HALO, need))
            return fail("oracle: slabs thinner than the halo");
        if (nr > 1 && !g_exchange) return fail("oracle: slab_nranks > 1 needs opesci_oracle_set_exchange first");
        M->p.dim[0] = M->slab.L1 - M->slab.L0;   /* every loop below runs on the local slab */
    }
    M->s[0] = (long)params->dim[1] * params->dim[2];
    M->s[1] = params->dim[2];
    M->s[2] = 1;
    M->level_elems = (size_t)M->p.dim[0] * params->dim[1] * params->dim[2];
    /* deep-copy tables */
    size_t total = 0;
    for (int f = 0; f < params->nfields; ++f)
        for (int w = 0; w < 2; ++w) {
            const OpesciSolProgram *pr = w ? &params->fields[f].final_ : &params->fields[f].init;
            for (int t = 0; t < pr->n_tables; ++t) total += (size_t)params->dim[pr->table_axis[t]];
        }
    M->tables = (double *)malloc((total ? total : 1) * sizeof(double));
    size_t pos = 0;
    for (int f = 0; f < params->nfields; ++f)
        for (int w = 0; w < 2; ++w) {
            OpesciSolProgram *pr = w ? &M->p.fields[f].final_ : &M->p.fields[f].init;
            for (int t = 0; t < pr->n_tables; ++t) {
                size_t n = (size_t)params->dim[pr->table_axis[t]];
                memcpy(M->tables + pos, pr->table[t], n * sizeof(double));
                /* x tables are indexed with local plane numbers: shift by the slab origin */
                pr->table[t] = M->tables + pos + (pr->table_axis[t] == 0 ? M->slab.L0 : 0);
                pos += n;
            }
        }
    for (int f = 0; f < params->nfields; ++f) {
        OpesciFieldSpec *fs = &M->p.fields[f];
        const OpesciSlab *sl = &M->slab;
        /* init on every stored plane, L2 on the owned planes; global -> local x */
        fs->lo[0] = (fs->lo[0] > sl->L0 ? fs->lo[0] : sl->L0) - sl->L0;
        fs->hi[0] = (fs->hi[0] < sl->L1 ? fs->hi[0] : sl->L1) - sl->L0;
        fs->l2_lo[0] = (fs->l2_lo[0] > sl->own_lo ? fs->l2_lo[0] : sl->own_lo) - sl->L0;
        fs->l2_hi[0] = (fs->l2_hi[0] < sl->own_hi ? fs->l2_hi[0] : sl->own_hi) - sl->L0;
    }
    if ((params->n_receivers > 0 || params->src_nt > 0) && params->kind != OPESCI_KIND_STAGGERED_ELASTIC)
        return fail("point source / receivers: staggered elastic model only");
    for (int r = 0; r < params->n_receivers; ++r)
        for (int d = 0; d < 3; ++d)
            if (params->receiver_cells[3 * r + d] < 0 || params->receiver_cells[3 * r + d] >= params->dim[d]) return fail("receiver cell outside the grid");
    if (params->src_nt > 0)
        for (int d = 0; d < 3; ++d)
            if (params->source_cell[d] < 0 || params->source_cell[d] >= params->dim[d]) return fail("source cell outside the grid");
    if (params->kind == OPESCI_KIND_STAGGERED_ELASTIC) {
        if (params->nfields != 9 || params->nlevels != 2) return fail("staggered: need 9 fields, 2 levels");
        if (params->hetero && params->is_double) return fail("heterogeneous media: fp32 only (the reader is float*)");
        if (params->hetero) build_staggered_hetero(M); else build_staggered(M);
        if (params->free_surface == 1) {
            if (params->so != 4) return fail("Levander free surface needs so == 4");
            if (params->hetero) build_levander_hetero(M); else build_levander(M);
        }
    } else if (params->kind == OPESCI_KIND_REGULAR_ACOUSTIC) {
        if (params->nfields != 1 || params->nlevels != 3) return fail("regular: need 1 field, 3 levels");
        build_regular(M);
    } else {
        return fail("opesci_b200_configure: unknown kind");
    }
    M->configured = 1;
    g_err[0] = 0;
    return 0;
}

static double run_program(const OpesciSolProgram *pr, int x, int y, int z, double fieldval,
                          float *const *media, size_t cell)
{
    double st[OPESCI_PROG_STACK];
    int sp = 0;
    const int idx[3] = {x, y, z};
    for (int i = 0; i < pr->n_instr; ++i) {
        const OpesciSolInstr *in = &pr->instr[i];
        switch (in->op) {
        case OPESCI_OP_TABLE: st[sp++] = pr->table[in->arg][idx[pr->table_axis[in->arg]]]; break;
        case OPESCI_OP_CONST: st[sp++] = in->value; break;
        case OPESCI_OP_FIELD: st[sp++] = fieldval; break;
        case OPESCI_OP_MEDIA: st[sp++] = (double)media[in->arg][cell]; break;
        case OPESCI_OP_SQRT: st[sp - 1] = sqrt(st[sp - 1]); break;
        case OPESCI_OP_COS: st[sp - 1] = cos(st[sp - 1]); break;
        case OPESCI_OP_SIN: st[sp - 1] = sin(st[sp - 1]); break;
        case OPESCI_OP_ROUNDF: st[sp - 1] = (double)(float)st[sp - 1]; break;
        case OPESCI_OP_ADD: sp--; st[sp - 1] = st[sp - 1] + st[sp]; break;
        case OPESCI_OP_SUB: sp--; st[sp - 1] = st[sp - 1] - st[sp]; break;
        case OPESCI_OP_MUL: sp--; st[sp - 1] = st[sp - 1] * st[sp]; break;
        case OPESCI_OP_DIV: sp--; st[sp - 1] = st[sp - 1] / st[sp]; break;
        case OPESCI_OP_NEG: st[sp - 1] = -st[sp - 1]; break;
        default: break;
        }
    }
    return sp > 0 ? st[sp - 1] : 0.0;
}

#include <time.h>
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* ------------------------------------------------- precision-generic section ---- */
#define REAL float
#define SFX(n) n##_f32
#include "opesci_oracle_loops.inc"
#undef REAL
#undef SFX
#define REAL double
#define SFX(n) n##_f64
#include "opesci_oracle_loops.inc"
#undef REAL
#undef SFX

/* ------------------------------------------------------------------ C ABI ---- */
int opesci_execute(OpesciGrid *grid, OpesciProfiling *profiling)
{
    Model *M = &g_model;
    if (!M->configured) return fail("opesci_execute: opesci_b200_configure was not called");
    double t0 = now_s();
    int rc = M->p.is_double ? execute_f64(M, grid) : execute_f32(M, grid);
    if (profiling) {
        profiling->g_rtime = (float)g_loop_seconds;
        profiling->g_ptime = (float)(now_s() - t0);
        profiling->g_mflops = 0.0f;
    }
    return rc;
}

int opesci_convergence(OpesciGrid *grid, OpesciConvergence *conv)
{
    Model *M = &g_model;
    if (!M->configured) return fail("opesci_convergence: not configured");
    return M->p.is_double ? convergence_f64(M, grid, conv, NULL) : convergence_f32(M, grid, conv, NULL);
}

int opesci_b200_convergence_f64(OpesciGrid *grid, double *out)
{
    Model *M = &g_model;
    if (!M->configured) return fail("opesci_b200_convergence_f64: not configured");
    return M->p.is_double ? convergence_f64(M, grid, NULL, out) : convergence_f32(M, grid, NULL, out);
}

int opesci_free(OpesciGrid *grid)
{
    /* opesci/regulargrid.py:621-634 */
    for (int f = 0; f < g_model.p.nfields; ++f) {
        free(grid->field[f]);
        grid->field[f] = NULL;
    }
    return 0;
}

int opesci_b200_comm_unique_id(void *out_id, int nbytes) { (void)out_id; (void)nbytes; return fail("oracle: no NCCL; use opesci_oracle_set_exchange"); }
int opesci_b200_comm_init(int rank, int nranks, const void *id, int nbytes) { (void)rank; (void)nranks; (void)id; (void)nbytes; return fail("oracle: no NCCL; use opesci_oracle_set_exchange"); }
int opesci_b200_comm_finalize(void) { return 0; }
int opesci_b200_halo_transport(void) { return 0; }
int opesci_b200_slab_range(int rank, int nranks, int gdim1, int so, int *L0, int *L1)
{
    OpesciSlab sl;
    if (opesci_slab_make(&sl, rank, nranks, gdim1, so / 2, OPESCI_SLAB_HALO, opesci_slab_need(0, so))) return fail("slabs thinner than the halo");
    if (L0) *L0 = sl.L0;
    if (L1) *L1 = sl.L1;
    return 0;
}
int opesci_b200_time_fused_parts(OpesciGrid *grid, int reps, double *out) { (void)grid; (void)reps; (void)out; return fail("opesci_b200_time_fused_parts: CUDA library only"); }
int opesci_b200_execute_loopback(int nranks, OpesciGrid *grids) { (void)nranks; (void)grids; return fail("opesci_b200_execute_loopback: CUDA library only"); }
int opesci_b200_reserve_host(size_t bytes_per_array, int count) { (void)bytes_per_array; (void)count; return 0; }
int opesci_b200_release_host(void) { return 0; }

int opesci_b200_time_kernels(OpesciGrid *grid, int reps, double *out_ms)
{
    (void)grid; (void)reps; (void)out_ms;
    return fail("opesci_b200_time_kernels: CUDA library only");
}

int opesci_b200_last_timing(double *loop_seconds, double *points_per_step, int64_t *kernel_launches)
{
    const Model *M = &g_model;
    if (loop_seconds) *loop_seconds = g_loop_seconds;
    if (points_per_step)
        *points_per_step = (double)(M->gdim1 - 2 * M->m) * (M->p.dim[1] - 2 * M->m) * (M->p.dim[2] - 2 * M->m);
    if (kernel_launches) *kernel_launches = 0;
    return 0;
}
