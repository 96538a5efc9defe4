/*
 * opesci_b200.h -- C ABI of the B200-native opesci-fd time-stepping library
 *
 * Drop-in boundary.  The reference JIT-compiles one shared object per model
 * and loads it with ctypes (reference: opesci/grid.py:35-42); that object exports exactly
 *
 *     int opesci_execute    (OpesciGrid *grid, OpesciProfiling *profiling);
 *     int opesci_convergence(OpesciGrid *grid, OpesciConvergence *conv);
 *     int opesci_free       (OpesciGrid *grid);
 *
 * (reference: opesci/templates/regular3d_tmpl.py:44-48, 106-118; struct layouts
 * regular3d_tmpl.py:19-23, 41-42 and opesci/regulargrid.py:435-443, 636-642, 349-354).
 * The same three symbols with the same struct layouts are exported here.
 *
 * In the reference every model parameter (dims, dt, ntsteps, the stencil coefficients as
 * decimal float literals, the analytic solution) is baked into the generated source
 * (opesci/regulargrid.py:391-406).  A prebuilt library needs them at run time, so ONE
 * additive call carries them: opesci_b200_configure().  The parameter block holds exactly
 * the values the generator would have printed, already rounded the way the printer rounds
 * them (opesci/codeprinter.py:30,46-63: 15 significant digits, then an `F` float literal).
 *
 * Plain C, plain pointers and sizes only.  No torch types.
 */
#ifndef OPESCI_B200_H
#define OPESCI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OPESCI_MAX_M 6            /* so/2 for so <= 12 */
#define OPESCI_MAX_FIELDS 9
#define OPESCI_MAX_TABLES 12
#define OPESCI_MAX_PROG 48
#define OPESCI_PROG_STACK 12

/* ---- reference structs (field order = grid.fields: U,V,W,Txx,Tyy,Tzz,Txy,Tyz,Txz;
 *      opesci/staggeredgrid.py:69-71, tests/eigenwave3d.py:64-67) ---------------------- */
/* The reference declares `real_t *` members; pointer size is the same for float/double, so
 * one layout serves both precisions (the ctypes side always uses POINTER(c_float),
 * opesci/grid.py:106).  A RegularGrid model has ONE member (its user-named field). */
typedef struct OpesciGrid {
    void *field[OPESCI_MAX_FIELDS];
} OpesciGrid;

/* `real_t <field>_l2` per field (opesci/regulargrid.py:636-642).  Written as float or double
 * according to the configured precision, packed in field order. */
typedef union OpesciConvergence {
    float f32[OPESCI_MAX_FIELDS];
    double f64[OPESCI_MAX_FIELDS];
} OpesciConvergence;

/* opesci/regulargrid.py:349-354 (no PAPI events: GPU build has no PAPI) */
typedef struct OpesciProfiling {
    float g_rtime;   /* seconds spent in the time loop (CUDA events) */
    float g_ptime;   /* seconds of opesci_execute wall time */
    float g_mflops;  /* reference op count (48*so flop/point/step, regulargrid.py:293-327) / g_rtime */
} OpesciProfiling;

/* ---- analytic-solution programs ---------------------------------------------------------
 * The reference prints the user's sympy solution into the init loops and the L2 loops
 * (opesci/staggeredgrid.py:612-659, 892-945; opesci/regulargrid.py:498-528, 650-700) and libm
 * evaluates it per cell in double.  Here the host evaluates every maximal sub-expression that
 * depends on ONE spatial index (e.g. sin(M_PI*y)) into a 1-D table with the same libm, and
 * the remaining + - * tree runs per cell on the device in IEEE double -- same operations,
 * same order, same bits. */
enum {
    OPESCI_OP_TABLE = 1, /* push table[arg][index along its axis] */
    OPESCI_OP_CONST = 2, /* push value */
    OPESCI_OP_ADD = 3,
    OPESCI_OP_SUB = 4,
    OPESCI_OP_MUL = 5,
    OPESCI_OP_NEG = 6,
    OPESCI_OP_DIV = 7,
    OPESCI_OP_FIELD = 8, /* push (double)F[x][y][z]: only in `final_` programs */
    /* heterogeneous (`read`) mode: the solution contains per-cell media, e.g.
     * cos(...*sqrt(beta[_x][_y][_z]*mu[_x][_y][_z])) (opesci/staggeredgrid.py:648-653) */
    OPESCI_OP_MEDIA = 9,   /* push (double)media[arg][x][y][z], arg = OPESCI_MEDIA_* */
    OPESCI_OP_SQRT = 10,
    OPESCI_OP_COS = 11,
    OPESCI_OP_SIN = 12,
    OPESCI_OP_ROUNDF = 13  /* round the top of the stack to float (a `float` typed C sub-expression) */
};

/* media arrays of the heterogeneous mode (opesci/staggeredgrid.py:262-276) */
enum {
    OPESCI_MEDIA_BETA = 0, OPESCI_MEDIA_LAMBDA, OPESCI_MEDIA_MU,
    OPESCI_MEDIA_BETA1, OPESCI_MEDIA_BETA2, OPESCI_MEDIA_BETA3,
    OPESCI_MEDIA_MU12, OPESCI_MEDIA_MU13, OPESCI_MEDIA_MU23, OPESCI_MEDIA_COUNT
};

typedef struct OpesciSolInstr {
    int32_t op;
    int32_t arg;
    double value;
} OpesciSolInstr;

typedef struct OpesciSolProgram {
    int32_t n_instr;
    int32_t n_tables;
    int32_t table_axis[OPESCI_MAX_TABLES];   /* 0,1,2 = generator axes x,y,z (dim1,dim2,dim3) */
    const double *table[OPESCI_MAX_TABLES];  /* HOST pointers, dim[axis] entries each */
    OpesciSolInstr instr[OPESCI_MAX_PROG];
} OpesciSolProgram;

typedef struct OpesciFieldSpec {
    int32_t lo[3], hi[3];        /* init loop ranges [lo,hi) per axis (staggeredgrid.py:632-640, regulargrid.py:508-511) */
    int32_t l2_lo[3], l2_hi[3];  /* L2 loop ranges (staggeredgrid.py:918-926, regulargrid.py:676-681) */
    OpesciSolProgram init;       /* solution at the field's first time (0 or dt/2) */
    OpesciSolProgram final_;     /* the whole printed residual `F - sol(t_last)` (uses OPESCI_OP_FIELD);
                                  * L2 = sqrt(volume_literal * sum residual^2) */
} OpesciFieldSpec;

enum { OPESCI_KIND_STAGGERED_ELASTIC = 1, OPESCI_KIND_REGULAR_ACOUSTIC = 2,
       OPESCI_KIND_REGULAR_GENERIC = 3 /* any PDE system on a RegularGrid: kernels compiled at run time (generic_source) */ };

/* flags */
enum {
    OPESCI_ARITH_REFERENCE = 0,      /* term order + separate mul/add exactly as emitted: bit-exact */
    OPESCI_ARITH_FAST = 1,           /* factored c_k*(a-b), FMA contraction allowed */
    OPESCI_ARITH_MASK = 0x3,
    OPESCI_HOST_MIRROR_FULL = 0 << 4,   /* grid->field[] = host arrays, all time levels (reference ABI) */
    OPESCI_HOST_MIRROR_NONE = 1 << 4,   /* grid->field[] = DEVICE pointers; nothing copied back */
    OPESCI_HOST_MIRROR_MASK = 0x30,
    OPESCI_NO_CUDA_GRAPH = 1 << 8,
    OPESCI_FORCE_UNFUSED = 1 << 9,      /* two-pass stress / velocity kernels (diagnostic) */
    OPESCI_FORCE_TILED = 1 << 11,
    OPESCI_NO_ZFOLD = 1 << 13,          /* diagnostic: keep the z-face stress ghost loops and the z slabs of the velocity shell in
                                         * the separate face / shell kernels instead of the z-edge tiles of the fused kernel */
    OPESCI_NO_PAIR = 1 << 14,           /* diagnostic: interior fused launch as single CTAs instead of 2-CTA clusters stacked in y */
    OPESCI_L2_REFERENCE = 1 << 12,      /* opesci_convergence accumulates like the reference: serially, in real_t, in loop
                                         * order (staggeredgrid.py:916,935) -- reproduces its printed digits; a serial
                                         * chain by definition (seconds at 256^3), single rank only.  Default: double tree */       /* TMA-tiled two-pass kernels also where the fused kernel applies (diagnostic) */
    OPESCI_OVERLAP = 1 << 10            /* accepted and ignored.  Round 1 had an experimental schedule behind it (ghost loops
                                         * of step n-1 concurrent with independent tiles of step n): no gain measured on
                                         * B200, and it raced with the point source, so the path was removed */
};

typedef struct OpesciB200Params {
    uint32_t struct_size;        /* sizeof(OpesciB200Params): layout check */
    int32_t kind;
    int32_t so;                  /* spatial order 2..12; margin m = so/2 (regulargrid.py:128) */
    int32_t is_double;           /* real_t = double (switch `double`, regulargrid.py:27-28) */
    int32_t dim[3];              /* dim1..3 = grid_size + 1 + 2m (regulargrid.py:152) */
    int32_t ntsteps;
    int32_t nfields;             /* 9 (staggered) or 1 (regular) */
    int32_t nlevels;             /* time levels tp: 2 (staggered), 3 (regular, regulargrid.py:118-120) */
    int32_t converge;            /* switch `converge` */
    int32_t free_surface;        /* 1: Levander (so==4), 2: Robertsson (so!=4), 0: none (staggeredgrid.py:223-226) */
    int32_t flags;
    int32_t warmup_steps;        /* the first `warmup_steps` of `ntsteps` are excluded from the loop timing */
    int32_t slab_rank;           /* x-slab decomposition: this process's rank ... */
    int32_t slab_nranks;         /* ... of slab_nranks (0 or 1: single domain).  dim[0] is always the GLOBAL dim1 */
    double dt;
    double dx[3];
    double volume_literal;       /* dx1*dx2*dx3 as printed (float literal) for the L2 scale */

    /* ---- staggered elastic, homogeneous medium: every literal of the emitted kernels ------
     * c_* = c_k * dt/dx_d * {lambda+2mu | lambda | mu | beta}, k = 1..m, magnitude AND sign of
     * the +offset half of the window; the -offset half uses the negated value. */
    float c_stress_normal[3][3][OPESCI_MAX_M];   /* [T_aa: a][axis d][k-1]   staggeredgrid.py:728-737 */
    float c_stress_shear[3][2][OPESCI_MAX_M];    /* [Txy,Tyz,Txz][term][k-1]; term0 = d_b V_a, term1 = d_a V_b */
    float c_velocity[3][3][OPESCI_MAX_M];        /* [V_a][axis d][k-1]       staggeredgrid.py:739-748 */
    /* Levander free surface, so == 4 only (fields.py:208-242, 313-353) */
    float lev_stress[3][3][3][2];  /* [face axis d][T_ee: e][derivative axis f][k-1], e != d, f != d */
    float lev_vnormal[3][3];       /* [d][e]: r*dx_d/dx_e  (fields.py:220-224), e != d */
    float lev_vtang[3][3];         /* [d][e]: dx_d/dx_e    (fields.py:226-230), e != d */

    /* ---- regular acoustic (regulargrid.py:592-619, 530-564) ------------------------------ */
    float ac_coef[3][OPESCI_MAX_M];   /* [axis][k-1]: weight of u[t1][i+-k]; 0 if the axis is absent */
    float ac_centre;                  /* weight of u[t1][i] */
    float ac_init_coef[3][OPESCI_MAX_M];
    float ac_init_centre;
    double ac_init_const;             /* the `1.0F*v*dt` term, evaluated in real_t by the host */

    /* ---- staggered elastic, heterogeneous medium (`read` mode; staggeredgrid.py:249-276, 522-598) ----
     * hetero != 0: the library derives beta, beta1-3, lambda, mu, mu12/13/23 from rho, vp, vs
     * (with the ranges of the patched oracle: pointwise arrays on [0,dim), averaged ones
     * on [0,dim-1)) and every emitted term becomes `literal*G[...]*media[x][y][z]` (a12).  fp32 only
     * (the file reader is float*, src/opesciIO.cpp:319).  The c_* / lev_stress / lev_vnormal tables above
     * are unused; lev_vtang (dx_d/dx_e) still applies. */
    int32_t hetero;
    int32_t media_plane0;            /* global index of the first x plane held by rho/vp/vs ... */
    int32_t media_nplanes;           /* ... and how many planes they hold (whole array: 0, dim1) */
    int32_t fs_faces;            /* bit (2*d + side) set: face (axis d = 0..2, side 0 low / 1 high) carries the free-surface
                                  * treatment (set_free_surface_boundary, staggeredgrid.py:214-232); 0 = all six */
    const float *rho, *vp, *vs;      /* HOST arrays [media_nplanes][dim2][dim3]: the flat float32 layout the
                                      * reference's raw-binary reader expects: opesci/staggeredgrid.py:549-551 */
    float h_c[3][OPESCI_MAX_M];      /* [axis d][k-1]: c_k*dt/dx_d */
    float h_c2[3][OPESCI_MAX_M];     /* 2*c_k*dt/dx_d: the `*mu` terms of the own-axis window of T_dd */
    /* Levander with per-cell media (so == 4), P_d = product of the two spacings other than dx_d: */
    float h_lev_den[3][2];           /* [face d]: denominator D = a*lambda + b*mu, a = 12 P_d, b = 24 P_d */
    float h_lev_own[3][3][2];        /* [face d][axis f][k-1] = 48 P_d c_k dt/dx_f: window of V_f in T_ff, as
                                      * `*lambda*mu/D` and `*pow(mu,2)/D` terms */
    float h_lev_oth[3][3][2];        /* [face d][axis f][k-1] = 24 P_d c_k dt/dx_f: window of V_f in T_ee, e != f,
                                      * as `*lambda*mu/D` terms */
    float h_vn[3][2];                /* velocity normal ghost, [axis g]: P_g and 2 P_g */

    /* ---- point source + receivers.  Not part of the generated code: the semantics are those of
     * the reference's hand-written propagator (tests/src/test_ref_iso_elastic.cpp:227-290).  At the end of time step
     * ti (after the velocity ghost loops):
     *   1. receiver r samples U, V, W and (Txx+Tyy+Tzz)/3 of the new time level at its grid cell
     *      -> receiver_out[((ti*4 + c)*n_receivers) + r], c = 0..3, real_t;
     *   2. if ti < src_nt: Txx, Tyy, Tzz[new level][source cell] -= src_x|y|z[ti]/3   (explosive source).
     * Cells are array indices (x,y,z) including the ghost margin m: round(coordinate/dx_d) + m.  Staggered elastic
     * model only.  With slabs every rank handles the cells on planes it owns; other receivers read 0. */
    int32_t n_receivers;
    int32_t src_nt;                   /* 0: no source */
    int32_t source_cell[3];
    int32_t reserved2_;
    const int32_t *receiver_cells;    /* HOST [n_receivers][3] */
    const float *src_x, *src_y, *src_z;   /* HOST [src_nt] each */
    void *receiver_out;               /* HOST [ntsteps][4][n_receivers] real_t, filled by opesci_execute */

    OpesciFieldSpec fields[OPESCI_MAX_FIELDS];

    /* ---- OPESCI_KIND_REGULAR_GENERIC (reference: opesci/regulargrid.py:230-270, 530-619).
     * CUDA C++ source defining
     *   extern "C" __global__ void opesci_generic_step (real_t *f0, ..., real_t *f{nfields-1}, int _t0, int _t1, int _t2);
     *   extern "C" __global__ void opesci_generic_init2(real_t *f0, ..., real_t *f{nfields-1}, int _t0, int _t1);
     * one thread per interior point (blocks of 64 x 4 threads along z, y; blockIdx.z = plane - m), arrays indexed as
     * [level][dim1][dim2][pitch] with pitch = dim3 rounded up to a multiple of 32.  The host front end prints into it the
     * expressions the reference's generator emits for the same PDEs; the library compiles it for sm_100a with NVRTC
     * (`--fmad=false` in reference arithmetic).  nlevels must be 3; no slabs. */
    const char *generic_source;
} OpesciB200Params;

/* ---- entry points ---------------------------------------------------------------------- */
/* Replaces: the constants baked by opesci/regulargrid.py:391-406 into every generated file.
 * Must be called before opesci_execute; copies everything it needs (tables included). */
int opesci_b200_configure(const OpesciB200Params *params);

/* Replaces generated opesci_execute (templates/regular3d_tmpl.py:44-59, staggered3d_tmpl.py:9-58):
 * allocates the fields (zero-filled), runs init + init BCs + ntsteps steps, stores
 * the base pointers into *grid.  Returns 0, or non-zero with opesci_b200_last_error() set. */
int opesci_execute(OpesciGrid *grid, OpesciProfiling *profiling);

/* Replaces generated opesci_convergence (regular3d_tmpl.py:106-112; staggeredgrid.py:892-945). */
int opesci_convergence(OpesciGrid *grid, OpesciConvergence *conv);

/* Replaces generated opesci_free (regular3d_tmpl.py:114-118; regulargrid.py:621-634). */
int opesci_free(OpesciGrid *grid);

/* Additive helpers (no reference counterpart) */
const char *opesci_b200_last_error(void);
/* L2 sums in double: out[f] = sqrt(volume * sum (F - sol)^2), double accumulation */
int opesci_b200_convergence_f64(OpesciGrid *grid, double *out_l2);
/* timing of the last opesci_execute: seconds in the time loop (device events) and point updates */
int opesci_b200_last_timing(double *loop_seconds, double *points_per_step, int64_t *kernel_launches);
/* Average device time (ms, CUDA events on the launching stream) of each interior kernel of one
 * time step, measured by re-launching it `reps` times on the resident fields of `grid`:
 * out_ms[0] = stress (or fused stress+velocity) kernel, out_ms[1] = velocity kernel (0 if fused),
 * out_ms[2] = all ghost-cell loops of one step.  Advances the fields; call it last. */
int opesci_b200_time_kernels(OpesciGrid *grid, int reps, double *out_ms);
/* z-fold runs (so = 4, homogeneous: the fused step is two concurrent launches): device time of each launch on its own,
 * out[0] = interior tile columns, out[1] = the two z-edge tile columns (ms), out[2] = interior z columns the interior
 * launch stores, out[3] = all interior z columns.  All zeros when the fused step is a single launch.  Results of the
 * timing launches are wrong by construction (each skips the other's cells): call it last, like time_kernels. */
int opesci_b200_time_fused_parts(OpesciGrid *grid, int reps, double *out);
/* Host result arrays (OPESCI_HOST_MIRROR_FULL) come from a process-wide pool of page-locked blocks that
 * outlives opesci_free; reserve_host pre-fills it with `count` blocks of `bytes_per_array`
 * (page-locking tens of GB takes far longer than copying them), release_host frees the unused blocks. */
int opesci_b200_reserve_host(size_t bytes_per_array, int count);
int opesci_b200_release_host(void);

/* ---- multi-GPU: x-slab decomposition (no reference counterpart: the reference is OpenMP only) ----
 * One process per GPU.  dim1 is split into contiguous slabs; every rank keeps OPESCI_SLAB_HALO planes
 * of all fields on each inner side, computes a whole time step on its local slab (x-face loops only
 * on the first / last rank) and then refreshes the halo planes from its neighbours.  Transport: by default every
 * rank maps its neighbours' field allocations (cudaIpc) and its copy engines pull the planes over NVLink, NCCL
 * carrying one 4-byte token per neighbour and exchange for the ordering; OPESCI_HALO_P2P=0 in the environment, or
 * a platform where the mapping fails, moves the planes with ncclSend/ncclRecv instead (same planes, same results).
 * The host distributes the 128-byte NCCL unique id (rank 0 creates it), e.g. with torch.distributed. */
#define OPESCI_SLAB_HALO 8       /* minimum halo; the halo is max(8, need) with `need` from include/opesci_slab.h: 8 planes up to so=8, 2m beyond */
#define OPESCI_COMM_ID_BYTES 128
int opesci_b200_comm_unique_id(void *out_id, int nbytes);
int opesci_b200_comm_init(int rank, int nranks, const void *id, int nbytes);
int opesci_b200_comm_finalize(void);
/* halo transport of this thread's last opesci_execute: 0 no slabs, 1 ncclSend/Recv, 2 peer memory, 3 loopback copies */
int opesci_b200_halo_transport(void);
/* the planes [L0,L1) of global dim1 that rank `rank` of `nranks` stores (its slab plus halos): what a
 * heterogeneous run has to supply in rho/vp/vs (media_plane0 = L0, media_nplanes = L1-L0); staggered elastic model */
int opesci_b200_slab_range(int rank, int nranks, int gdim1, int so, int *L0, int *L1);
/* Loopback slabs: `nranks` logical ranks of the configured model (configured for the WHOLE domain, slab_nranks <= 1)
 * run concurrently on the CURRENT device, one host thread per rank, each through exactly the schedule a real rank runs
 * (x-chunk table, overlap of the halo exchange with the middle chunks, ghost loops, shell); only the transport of the
 * halo planes differs -- device-to-device copies instead of ncclSend / ncclRecv.  grids[r] receives rank r's arrays
 * (planes [L0,L1) of opesci_b200_slab_range, all time levels; host or device per the HOST_MIRROR flag); free each with
 * opesci_free.  Proof of slab exactness on a one-GPU box: the owned planes must equal the single-domain run bit for bit. */
int opesci_b200_execute_loopback(int nranks, OpesciGrid *grids);
/* 1 if this library was built with the CUDA kernels (0 for the CPU oracle build) */
int opesci_b200_is_cuda(void);

#ifdef __cplusplus
}
#endif
#endif /* OPESCI_B200_H */
