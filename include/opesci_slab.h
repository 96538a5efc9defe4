/*
 * opesci_slab.h -- geometry of the x-slab decomposition (shared by the CUDA library and the CPU oracle).
 *
 * The reference has no distributed backend.  Its generated loops run over the
 * whole grid; here generator axis x (dim1, slowest; a plane is one contiguous block) is cut into
 * contiguous slabs.  Rank r owns interior planes [X0,X1) (rank 0 also the low ghost planes, the
 * last rank the high ones) and stores planes [L0,L1) = owned +- H halo planes.  After every time
 * step (and after initialisation) the halo planes of all fields are overwritten with the
 * neighbour's owned planes.  Within one step an error starting at an artificial slab end
 * travels at most 2m+3 planes inwards (stress update m, velocity update m, the Levander ghost
 * loops on the y/z faces chain through 3 more x-neighbours: SURVEY.md 8e), so H >= 2m+3 keeps every
 * owned plane exact: results are bit-identical to the single-domain run.
 * The regular-grid acoustic model has no ghost-cell loops: an error travels m planes per step, H >= m.
 */
#ifndef OPESCI_SLAB_H
#define OPESCI_SLAB_H

typedef struct OpesciSlab {
    int rank, nranks, halo;
    int gdim;          /* global dim1 */
    int X0, X1;        /* owned interior planes */
    int own_lo, own_hi;/* owned planes incl. the physical ghost planes of the end ranks */
    int L0, L1;        /* stored planes; local index = global - L0 */
    int lo_face, hi_face;
} OpesciSlab;

/* Planes an error starting at an artificial slab end can travel inwards in one time step:
 *   staggered elastic, Levander (so == 4):   stress m + velocity m + ghost-loop chain 3 = 2m+3
 *   staggered elastic, Robertsson (so != 4): stress m + velocity m = 2m (its ghost loops only write zeros / in-axis mirrors)
 *   regular acoustic:                        m                                                       */
static inline int opesci_slab_need(int acoustic, int so)
{
    const int m = so / 2;
    if (acoustic) return m;
    return so == 4 ? 2 * m + 3 : 2 * m;
}

/* `need`: see opesci_slab_need; the halo is max(min_halo, need) planes per inner side.
 * returns 0 on success, 1 if the slabs would be thinner than the halo */
static inline int opesci_slab_make(OpesciSlab *s, int rank, int nranks, int gdim, int m, int min_halo, int need)
{
    const int halo = min_halo > need ? min_halo : need;
    if (nranks < 1) nranks = 1;
    const int n_int = gdim - 2 * m;
    const int base = n_int / nranks, rem = n_int % nranks;
    s->rank = rank; s->nranks = nranks; s->halo = halo; s->gdim = gdim;
    s->X0 = m + rank * base + (rank < rem ? rank : rem);
    s->X1 = s->X0 + base + (rank < rem ? 1 : 0);
    s->lo_face = rank == 0;
    s->hi_face = rank == nranks - 1;
    s->own_lo = s->lo_face ? 0 : s->X0;
    s->own_hi = s->hi_face ? gdim : s->X1;
    s->L0 = s->lo_face ? 0 : s->X0 - halo;
    s->L1 = s->hi_face ? gdim : s->X1 + halo;
    if (nranks > 1 && base < halo) return 1;
    return 0;
}

#endif
