// opesci_b200.cu -- C ABI (include/opesci_b200.h) and host orchestration of the CUDA kernels.
//
// Replaces the reference's generated `opesci_execute / opesci_convergence / opesci_free`
// (opesci/templates/regular3d_tmpl.py:44-59, 106-118; staggered3d_tmpl.py:9-58).  Per call:
//   opesci_execute: allocate + zero the fields on the device, analytic
//   initialisation, initial BC pass, `ntsteps` leapfrog steps replayed from a CUDA graph
//   (stress interior -> stress ghost loops -> velocity interior -> velocity ghost loops, the
//   order is a contract: staggered3d_tmpl.py:40-58), then (HOST_MIRROR_FULL) copy every time
//   level back into host arrays whose base pointers are stored into *grid, like the
//   reference does (regulargrid.py:445-453).
// There is no CPU fallback: every entry point fails loudly without a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include <dlfcn.h>

#include "../../include/opesci_slab.h"
#include "fused.cuh"
#include "generic.cuh"
#include "hetero.cuh"
#include "io.cuh"
#include "kernels.cuh"
#include "tiled.cuh"

using namespace opesci;

namespace {

thread_local char g_err[1024] = "";   // per thread: the loopback ranks run on threads of their own
int fail(const char *fmt, const char *a = "", const char *b = "")
{
    snprintf(g_err, sizeof g_err, fmt, a, b);
    return 1;
}
#define CUDA_OK(call)                                                                          \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) return fail("CUDA error: %s at %s", cudaGetErrorString(e_), #call); \
    } while (0)


// ------------------------------------------------------------------ NCCL (bound at run time)
// Only needed for slab_nranks > 1; the library itself does not link against NCCL, it uses the copy
// already loaded into the process (torch's bundled libnccl.so.2) or dlopens one.
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclFloat64 = 8, ncclUint8 = 1 };
enum { ncclSum = 0 };
struct Nccl {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
} g_nccl;

int nccl_bind()
{
    if (g_nccl.GetUniqueId) return 0;
    void *h = dlopen(nullptr, RTLD_NOW);                       // already in the process (torch)?
    if (!h || !dlsym(h, "ncclCommInitRank")) {
        // OPESCI_NCCL_LIB names the library explicitly; otherwise the standard search path decides
        const char *names[] = {getenv("OPESCI_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        h = nullptr;
        for (const char *n : names)
            if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    }
    if (!h) return fail("NCCL not found: set OPESCI_NCCL_LIB or put libnccl.so.2 on the library path (%s)", dlerror());
    g_nccl.handle = h;
#define BIND(field, sym) *(void **)(&g_nccl.field) = dlsym(h, sym); if (!g_nccl.field) return fail("NCCL symbol missing: %s", sym)
    BIND(GetUniqueId, "ncclGetUniqueId");
    BIND(CommInitRank, "ncclCommInitRank");
    BIND(CommDestroy, "ncclCommDestroy");
    BIND(Send, "ncclSend");
    BIND(Recv, "ncclRecv");
    BIND(GroupStart, "ncclGroupStart");
    BIND(GroupEnd, "ncclGroupEnd");
    BIND(AllReduce, "ncclAllReduce");
    BIND(GetErrorString, "ncclGetErrorString");
#undef BIND
    return 0;
}
#define NCCL_OK(call)                                                                           \
    do {                                                                                        \
        int r_ = (call);                                                                        \
        if (r_ != ncclSuccess) return fail("NCCL error: %s at %s", g_nccl.GetErrorString(r_), #call); \
    } while (0)


// halo transport of the last run of this thread (opesci_b200_halo_transport): 0 none, 1 NCCL send/recv,
// 2 peer memory (neighbour's fields mapped with cudaIpc, planes pulled by the copy engines), 3 loopback copies
thread_local int tl_halo_transport = 0;

// ------------------------------------------------------------------ loopback slabs
// N logical ranks of one model executed concurrently on ONE device, each by its own host thread running the very same
// run_model schedule a real rank runs (chunk table, ev_fork / ev_join, exchange on its own stream); only the transport of
// the halo planes differs: device-to-device copies between the ranks' arrays instead of ncclSend / ncclRecv.
// Purpose: proof of slab exactness on a single-GPU box (opesci_b200_execute_loopback, tests/test_gpu_loopback.py).
struct Run;
struct HostBarrier {
    std::mutex mu;
    std::condition_variable cv;
    int n = 0, waiting = 0;
    unsigned long long gen = 0;
    bool aborted = false;
    // returns true if a peer gave up (the caller must fail too instead of waiting for ever)
    bool wait()
    {
        std::unique_lock<std::mutex> lk(mu);
        if (aborted) return true;
        const unsigned long long g = gen;
        if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); return false; }
        cv.wait(lk, [&] { return gen != g || aborted; });
        return aborted;
    }
    void abort()
    {
        std::lock_guard<std::mutex> lk(mu);
        aborted = true;
        cv.notify_all();
    }
};
struct Loopback {
    int nranks = 0;
    std::vector<Run *> runs;
    std::vector<cudaEvent_t> done, xdone;
    HostBarrier barrier;
};
thread_local Loopback *tl_loop = nullptr;


struct Model {
    OpesciB200Params p;
    bool configured = false;
    int m = 0;
    GridGeom G;
    StaggeredCoefs sc;
    HeteroCoefs hc;                               // heterogeneous (`read`) mode
    AcousticCoefs ac, ac_init;
    DevEq lev_stress_eq[3][3];
    DevEq lev_vel_eq[3][3][2];
    std::vector<double> tables;                  // host copy of every table
    std::vector<size_t> table_off[OPESCI_MAX_FIELDS][2];
    OpesciSlab slab;                              // x-slab of this rank (whole domain when nranks == 1)
    std::string generic_source;                   // OPESCI_KIND_REGULAR_GENERIC: private copy of the CUDA source
};
Model g_model;

// device-resident state of one executed model, keyed by the pointer stored in grid->field[0]
struct Run {
    Model M;
    void *dev[OPESCI_MAX_FIELDS] = {};
    void *host[OPESCI_MAX_FIELDS] = {};
    float *media[OPESCI_MEDIA_COUNT] = {};   // heterogeneous mode: derived media arrays (one level each)
    // point source + receivers
    bool hooks = false;
    long long *d_recv_cell = nullptr;   // [n_receivers] element offsets inside one level, -1 = not owned
    void *d_recv_out = nullptr;         // [ntsteps][4][n_receivers] real_t
    float *d_src = nullptr;             // [3][src_nt]
    int *d_step = nullptr;              // time-step counter on the device
    int *d_pace = nullptr;              // fused kernel, OPESCI_PACE: current plane of every tile [chunk][ytile][ztile]
    long long src_cell = -1;
    bool host_pinned = false;
    double *d_tables = nullptr;
    DevProgram *d_prog = nullptr;   // [nfields][2]
    size_t bytes_per_field = 0;     // device bytes (pitched rows)
    size_t host_bytes_per_field = 0;
    bool fused = false;             // fused stress+velocity kernel in use
    CUtensorMap tmap[3];            // U, V, W (both time levels; level selected through the x coordinate)
#if OPESCI_TMA_STORE
    StoreMaps smaps;                // TMA stores of Txy, Txz, Tyy, Tyz, Tzz from the stress ring (fused.cuh)
#endif
    bool tiled = false;             // TMA-tiled two-pass kernels in use (tiled.cuh: so >= 6, fp64)
    CUtensorMap tmap9[OPESCI_MAX_FIELDS];   // all nine fields, box = TileCfg tile
    int nchunks = 1;                // x-chunks of the fused kernel: chunk c covers planes [xs[c], xs[c+1])
    int xs[OPESCI_MAX_CHUNKS + 1] = {};
    int zstrip = 0;                 // > 0: the fused kernel covers z < zstrip only; the thin strip [zstrip, dim-m) is done per point
    bool split_exchange = false;    // slabs: stress / velocity halo exchanges issued separately (setup_fused)
    int mid0 = 0, mid1 = 0;         // slabs: chunks [mid0, mid1) read no halo plane (they overlap the halo exchange)
    opesci_generic::Module gen;     // OPESCI_KIND_REGULAR_GENERIC: the NVRTC-compiled kernels of this model
    bool pair = false;              // interior fused launch as 2-CTA clusters stacked in y (fused.cuh, PAIR)
    StoreMaps smaps_pair;           // TMA store boxes of FusedCfg::PCY rows
    // z-fold (fused.cuh, ZF kernels): the z-face stress ghost loops and the z slabs of the velocity shell are done by the
    // z-edge tiles of the fused kernel, launched beside the interior tiles on a second stream
    bool zfold = false;
    int zf_nzt = 0;                 // tile columns
    cudaStream_t st_edge = nullptr;
    cudaEvent_t ev_edge_fork = nullptr, ev_edge_join = nullptr;
};
std::map<void *, Run *> g_runs;
std::mutex g_mu;

double g_loop_seconds = 0.0;
long long g_launches = 0;

// ------------------------------------------------------------------ pinned host pool
// The reference ABI hands back host arrays.  Page-locking 80 GB costs tens of seconds, far more than
// the copy itself, so result arrays come from a process-wide pool of pinned blocks that survives
// opesci_free (opesci_b200_reserve_host pre-fills it, opesci_b200_release_host frees it).
// Blocks reserved explicitly (opesci_b200_reserve_host) stay until opesci_b200_release_host; blocks the pool had to
// allocate on demand are kept after opesci_free only while the idle ones add up to at most OPESCI_HOST_POOL_IDLE_MB
// (default 1024 MiB), so a caller that never heard of the pool -- the reference front end -- does not keep a whole
// run's worth of page-locked memory behind.
struct HostBlock { void *ptr; size_t bytes; bool in_use; bool pinned; bool reserved; };
std::vector<HostBlock> g_pool;
std::mutex g_pool_mu;
bool g_pool_reserving = false;
size_t pool_idle_cap()
{
    const char *e = getenv("OPESCI_HOST_POOL_IDLE_MB");
    return (size_t)(e && *e ? strtoull(e, nullptr, 10) : 1024ull) << 20;
}

void *pool_alloc(size_t bytes, bool *pinned)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int best = -1;
    for (int i = 0; i < (int)g_pool.size(); ++i)
        if (!g_pool[i].in_use && g_pool[i].bytes >= bytes && (best < 0 || g_pool[i].bytes < g_pool[best].bytes)) best = i;
    if (best >= 0) { g_pool[best].in_use = true; *pinned = g_pool[best].pinned; return g_pool[best].ptr; }
    HostBlock b{nullptr, bytes, true, true, g_pool_reserving};
    if (cudaMallocHost(&b.ptr, bytes) != cudaSuccess) {
        cudaGetLastError();
        b.pinned = false;
        if (posix_memalign(&b.ptr, 4096, bytes) != 0) return nullptr;
    }
    g_pool.push_back(b);
    *pinned = b.pinned;
    return b.ptr;
}
void pool_release(void *ptr)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (auto &b : g_pool)
        if (b.ptr == ptr) b.in_use = false;
    // trim: free idle on-demand blocks, largest first, until they fit under the cap
    const size_t cap = pool_idle_cap();
    for (;;) {
        size_t idle = 0;
        int big = -1;
        for (int i = 0; i < (int)g_pool.size(); ++i) {
            const HostBlock &b = g_pool[i];
            if (b.in_use || b.reserved) continue;
            idle += b.bytes;
            if (big < 0 || b.bytes > g_pool[big].bytes) big = i;
        }
        if (idle <= cap || big < 0) break;
        if (g_pool[big].pinned) cudaFreeHost(g_pool[big].ptr);
        else free(g_pool[big].ptr);
        g_pool.erase(g_pool.begin() + big);
    }
}

// ------------------------------------------------------------------ model construction
void push(DevEq &eq, int kind, int field, int level, long long off, float coef)
{
    DevTerm &t = eq.term[eq.nterm++];
    t.kind = kind; t.field = field; t.level = level; t.off = off; t.coef = coef;
    t.mk = MK_NONE; t.ma = t.mb = 0; t.pad = 0; t.moffa = t.moffb = 0;
}
void push_m(DevEq &eq, int kind, int field, int level, long long off, float coef, int mk, int ma, long long moffa, int mb,
            long long moffb)
{
    push(eq, kind, field, level, off, coef);
    DevTerm &t = eq.term[eq.nterm - 1];
    t.mk = mk; t.ma = ma; t.mb = mb; t.moffa = moffa; t.moffb = moffb;
}
const int NORMAL_OF_AXIS[3] = {F_TXX, F_TYY, F_TZZ};
const int VEL_OF_AXIS[3] = {F_U, F_V, F_W};

// backward window, m == 2, printer order +1, -1, -2, 0 (opesci/fields.py:313-353)
void push_window_bwd2(DevEq &eq, int field, int level, long long stride, const float *c)
{
    push(eq, TERM_MUL, field, level, stride, c[1]);
    push(eq, TERM_MUL, field, level, -stride, -c[0]);
    push(eq, TERM_MUL, field, level, -2 * stride, -c[1]);
    push(eq, TERM_MUL, field, level, 0, c[0]);
}

// Levander free-surface loops (so == 4): term order as emitted by the reference printer,
// alphabetical by field name then lexicographic by index (verified bit-for-bit through the
// oracle, tests/test_oracle_golden.py).
void build_levander(Model &M)
{
    const OpesciB200Params &p = M.p;
    const long long *s = M.G.s;
    for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) {
            DevEq &eq = M.lev_stress_eq[d][e];
            eq.out = NORMAL_OF_AXIS[e]; eq.out_level = 1; eq.nterm = 0;
            if (e == d) continue;
            push(eq, TERM_PLUS, eq.out, 0, 0, 1.0f);   // `1.0F*T[t0]` is exact
            for (int f = 0; f < 3; ++f)
                if (f != d) push_window_bwd2(eq, VEL_OF_AXIS[f], 0, s[f], p.lev_stress[d][e][f]);
        }
    for (int d = 0; d < 3; ++d)
        for (int a = 0; a < 3; ++a)
            for (int side = 0; side < 2; ++side) {
                DevEq &eq = M.lev_vel_eq[d][a][side];
                const long long sd = s[d];
                const float sgn = side == 0 ? 1.0f : -1.0f;
                eq.out = VEL_OF_AXIS[a]; eq.out_level = 0; eq.nterm = 0;
                if (a == d) {
                    // normal component: V_d[n] = V_d[n +- 1] +- sum_e a_e (V_e[e:0] - V_e[e:-1]) on the face plane
                    const long long plane = side == 0 ? sd : 0, selfoff = side == 0 ? sd : -sd;
                    for (int g = 0; g < 3; ++g) {
                        if (g == d) {
                            push(eq, TERM_PLUS, VEL_OF_AXIS[d], 0, selfoff, 1.0f);
                        } else {
                            const float c = p.lev_vnormal[d][g];
                            push(eq, TERM_MUL, VEL_OF_AXIS[g], 0, plane - s[g], -sgn * c);
                            push(eq, TERM_MUL, VEL_OF_AXIS[g], 0, plane, sgn * c);
                        }
                    }
                } else {
                    // tangential component V_e on face d
                    const int e = a;
                    const float g = p.lev_vtang[d][e];
                    const long long se = s[e];
                    const long long pl0 = side == 0 ? 0 : -sd, pl1 = side == 0 ? sd : -2 * sd;
                    const long long sf0 = side == 0 ? sd : -sd, sf1 = side == 0 ? 2 * sd : -2 * sd;
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

// Levander free-surface loops with per-cell media (so == 4): the emitted forms of the patched reference,
// term for term (oracle/opesci_oracle.c:build_levander_hetero documents them; parity is bit-exact through
// the oracle, tests/test_oracle_golden.py + tests/test_gpu_parity.py).
int mu_of_pair(int a, int b)
{
    if (a > b) { int t = a; a = b; b = t; }
    if (a == 0 && b == 1) return OPESCI_MEDIA_MU12;
    if (a == 1 && b == 2) return OPESCI_MEDIA_MU23;
    return OPESCI_MEDIA_MU13;
}
// backward window, m == 2, printer order +1, -1, -2, 0; `nv` media variants per offset
void push_window_bwd2_m(DevEq &eq, int field, long long stride, const float *c, int nv, const int *mk, const int *ma, const int *mb)
{
    const long long off[4] = {stride, -stride, -2 * stride, 0};
    const float coef[4] = {c[1], -c[0], -c[1], c[0]};
    for (int o = 0; o < 4; ++o)
        for (int j = 0; j < nv; ++j) push_m(eq, TERM_MUL, field, 0, off[o], coef[o], mk[j], ma[j], 0, mb[j], 0);
}
void build_levander_hetero(Model &M)
{
    const OpesciB200Params &p = M.p;
    const long long *s = M.G.s;
    const int LAM = OPESCI_MEDIA_LAMBDA, MU = OPESCI_MEDIA_MU;
    for (int d = 0; d < 3; ++d)
        for (int e = 0; e < 3; ++e) {
            DevEq &eq = M.lev_stress_eq[d][e];
            eq.out = NORMAL_OF_AXIS[e]; eq.out_level = 1; eq.nterm = 0;
            if (e == d) continue;
            eq.da = p.h_lev_den[d][0]; eq.db = p.h_lev_den[d][1]; eq.doff = 0;
            push_m(eq, TERM_MUL, eq.out, 0, 0, eq.da, MK_A_DIV, LAM, 0, 0, 0);
            push_m(eq, TERM_MUL, eq.out, 0, 0, eq.db, MK_A_DIV, MU, 0, 0, 0);
            for (int f = 0; f < 3; ++f) {
                if (f == d) continue;
                if (f == e) {
                    const int mk[2] = {MK_AB_DIV, MK_SQ_DIV}, ma[2] = {LAM, 0}, mb[2] = {MU, MU};
                    push_window_bwd2_m(eq, VEL_OF_AXIS[f], s[f], p.h_lev_own[d][f], 2, mk, ma, mb);
                } else {
                    const int mk[1] = {MK_AB_DIV}, ma[1] = {LAM}, mb[1] = {MU};
                    push_window_bwd2_m(eq, VEL_OF_AXIS[f], s[f], p.h_lev_oth[d][f], 1, mk, ma, mb);
                }
            }
        }
    for (int d = 0; d < 3; ++d)
        for (int a = 0; a < 3; ++a)
            for (int side = 0; side < 2; ++side) {
                DevEq &eq = M.lev_vel_eq[d][a][side];
                const long long sd = s[d];
                const float sgn = side == 0 ? 1.0f : -1.0f;
                eq.out = VEL_OF_AXIS[a]; eq.out_level = 0; eq.nterm = 0;
                if (a == d) {
                    const long long plane = side == 0 ? sd : 0, selfoff = side == 0 ? sd : -sd;
                    eq.da = p.h_vn[d][0]; eq.db = p.h_vn[d][1]; eq.doff = plane;
                    for (int g = 0; g < 3; ++g) {
                        if (g == d) {
                            push_m(eq, TERM_MUL, VEL_OF_AXIS[d], 0, selfoff, eq.da, MK_A_DIV, LAM, plane, 0, 0);
                            push_m(eq, TERM_MUL, VEL_OF_AXIS[d], 0, selfoff, eq.db, MK_A_DIV, MU, plane, 0, 0);
                        } else {
                            const float c = p.h_vn[g][0];
                            push_m(eq, TERM_MUL, VEL_OF_AXIS[g], 0, plane - s[g], -sgn * c, MK_A_DIV, LAM, plane, 0, 0);
                            push_m(eq, TERM_MUL, VEL_OF_AXIS[g], 0, plane, sgn * c, MK_A_DIV, LAM, plane, 0, 0);
                        }
                    }
                } else {
                    const int e = a;
                    const float g = p.lev_vtang[d][e];
                    const long long se = s[e];
                    const long long pl0 = side == 0 ? 0 : -sd, pl1 = side == 0 ? sd : -2 * sd;
                    const long long sf0 = side == 0 ? sd : -sd, sf1 = side == 0 ? 2 * sd : -2 * sd;
                    const int mu = mu_of_pair(d, e);
#define OPESCI_RATIO MK_RATIO, mu, pl1, mu, pl0
                    if (d < e) {
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl0 + se, sgn * g);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl0, -sgn * g);
                        push_m(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl1 + se, -sgn * g, OPESCI_RATIO);
                        push_m(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl1, sgn * g, OPESCI_RATIO);
                        push(eq, TERM_PLUS, VEL_OF_AXIS[e], 0, sf0, 1.0f);
                        push_m(eq, TERM_PLUS, VEL_OF_AXIS[e], 0, sf0, 1.0f, OPESCI_RATIO);
                        push_m(eq, TERM_MINUS, VEL_OF_AXIS[e], 0, sf1, 1.0f, OPESCI_RATIO);
                    } else {
                        push(eq, TERM_PLUS, VEL_OF_AXIS[e], 0, sf0, 1.0f);
                        push_m(eq, TERM_PLUS, VEL_OF_AXIS[e], 0, sf0, 1.0f, OPESCI_RATIO);
                        push_m(eq, TERM_MINUS, VEL_OF_AXIS[e], 0, sf1, 1.0f, OPESCI_RATIO);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, se + pl0, sgn * g);
                        push_m(eq, TERM_MUL, VEL_OF_AXIS[d], 0, se + pl1, -sgn * g, OPESCI_RATIO);
                        push(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl0, -sgn * g);
                        push_m(eq, TERM_MUL, VEL_OF_AXIS[d], 0, pl1, sgn * g, OPESCI_RATIO);
                    }
#undef OPESCI_RATIO
                }
            }
}

// number of SMs of the current device (148 on B200); grids are sized in multiples of it
int sm_count()
{
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// ------------------------------------------------------------------ launch helpers
#ifndef OPESCI_ZF_SUB
#define OPESCI_ZF_SUB 8   /* x-chunks of the z-edge launch per x-chunk of the interior launch (B200, 1024^3: 1 -> 22.19, 3 -> 21.83, 8 -> 21.63 ms per step) */
#endif
// Diagnostic (OPESCI_STEP_TRACE=1): CUDA events at the phase boundaries of every timed step of the staggered loop, the
// average time of each phase printed to stderr after the run.  Not part of any measured number: it disables the graph.
struct PhaseTrace {
    static const int NPH = 6;   // step start | fused | stress ghost loops | velocity shell | velocity ghost loops | halo wait
    bool on = false;
    std::vector<cudaEvent_t> ev;
    void mark(cudaStream_t st)
    {
        if (!on || ev.size() >= (size_t)NPH * 256) return;
        cudaEvent_t e = nullptr;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        cudaEventRecord(e, st);
        ev.push_back(e);
    }
    void reset() { for (cudaEvent_t e : ev) cudaEventDestroy(e); ev.clear(); }
    void report(int rank)
    {
        const size_t nst = ev.size() / NPH;
        if (on && nst > 0) {
            static const char *names[NPH] = {"", "fused", "stress_bc", "velocity_shell", "velocity_bc", "halo_wait"};
            double sum[NPH] = {0, 0, 0, 0, 0, 0}, gap = 0.0;
            for (size_t k = 0; k < nst; ++k) {
                for (int ph = 1; ph < NPH; ++ph) {
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, ev[k * NPH + ph - 1], ev[k * NPH + ph]);
                    sum[ph] += ms;
                }
                if (k + 1 < nst) { float ms = 0.f; cudaEventElapsedTime(&ms, ev[k * NPH + NPH - 1], ev[(k + 1) * NPH]); gap += ms; }
            }
            fprintf(stderr, "[opesci trace] rank %d, %zu steps, ms per step:", rank, nst);
            for (int ph = 1; ph < NPH; ++ph) fprintf(stderr, " %s %.3f", names[ph], sum[ph] / nst);
            fprintf(stderr, " between_steps %.3f\n", nst > 1 ? gap / (nst - 1) : 0.0);
        }
        reset();
    }
    ~PhaseTrace() { reset(); }
};

struct Stepper {
    const Run &R;
    cudaStream_t st;
    long long launches = 0;
    cudaError_t err = cudaSuccess;
    PhaseTrace *trace = nullptr;
    void mark() { if (trace) trace->mark(st); }
    Stepper(const Run &r, cudaStream_t s) : R(r), st(s) {}

    FieldPtrs ptrs() const
    {
        FieldPtrs F;
        for (int f = 0; f < OPESCI_MAX_FIELDS; ++f) F.f[f] = R.dev[f];
        return F;
    }
    MediaPtrs media() const
    {
        MediaPtrs MD;
        for (int k = 0; k < OPESCI_MEDIA_COUNT; ++k) MD.m[k] = R.media[k];
        return MD;
    }
    int fused_part = 0;   // timing only (opesci_b200_time_fused_parts): 1 = interior launch only, 2 = z-edge launch only
    void check()
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && err == cudaSuccess) err = e;
        ++launches;
    }
    dim3 interior_grid(dim3 blk) const
    {
        const Model &M = R.M;
        const int nz = M.G.dim[2] - 2 * M.m, ny = M.G.dim[1] - 2 * M.m, nx = M.G.dim[0] - 2 * M.m;
        return dim3((nz + blk.x - 1) / blk.x, (ny + blk.y - 1) / blk.y, nx);
    }

    // TMA-tiled two-pass kernels (tiled.cuh)
    template <int SO, typename T> dim3 tiled_grid(int *xchunk) const
    {
        using K = TileCfg<SO / 2, T>;
        const Model &M = R.M;
        const int nx = M.G.dim[0] - 2 * M.m, ny = M.G.dim[1] - 2 * M.m, nz = M.G.dim[2] - 2 * M.m;
        const int nbz = (nz + K::TZ - 1) / K::TZ, nby = (ny + K::TY - 1) / K::TY;
        int nchunks = (8 * sm_count() + nbz * nby - 1) / (nbz * nby);   // >= 8 CTAs per SM over the launch
        const int maxc = nx / (16 * M.m) > 0 ? nx / (16 * M.m) : 1;   // keep the 2m-plane warm-up of a chunk small
        if (nchunks > maxc) nchunks = maxc;
        if (nchunks < 1) nchunks = 1;
        *xchunk = (nx + nchunks - 1) / nchunks;
        return dim3(nbz, nby, (nx + *xchunk - 1) / *xchunk);
    }
    template <int SO, typename T, int ARITH> void stress_tiled_launch(int t0, int t1)
    {
        using K = TileCfg<SO / 2, T>;
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(stress_tiled<SO, T, ARITH>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::smem(3));
            attr = true;
        }
        TileArgs A;
        A.F = ptrs(); A.G = R.M.G; A.C = R.M.sc; A.MD = media(); A.HC = R.M.hc; A.t0 = t0; A.t1 = t1;
        const dim3 grid = tiled_grid<SO, T>(&A.xchunk);
        if constexpr (sizeof(T) == 4) {
            if (R.M.p.hetero) {
                static bool attr_h = false;
                if (!attr_h) {
                    cudaFuncSetAttribute(stress_tiled<SO, T, ARITH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::smem(3));
                    attr_h = true;
                }
                stress_tiled<SO, T, ARITH, true><<<grid, K::THREADS, K::smem(3), st>>>(R.tmap9[F_U], R.tmap9[F_V], R.tmap9[F_W], A);
                check();
                return;
            }
        }
        stress_tiled<SO, T, ARITH><<<grid, K::THREADS, K::smem(3), st>>>(R.tmap9[F_U], R.tmap9[F_V], R.tmap9[F_W], A);
        check();
    }
    template <int SO, typename T, int ARITH> void velocity_tiled_launch(int t0, int t1)
    {
        using K = TileCfg<SO / 2, T>;
        static bool attr = false;
        if (!attr) {
            cudaFuncSetAttribute(velocity_tiled<SO, T, ARITH>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::smem(5));
            attr = true;
        }
        TileArgs A;
        A.F = ptrs(); A.G = R.M.G; A.C = R.M.sc; A.MD = media(); A.HC = R.M.hc; A.t0 = t0; A.t1 = t1;
        const dim3 grid = tiled_grid<SO, T>(&A.xchunk);
        if constexpr (sizeof(T) == 4) {
            if (R.M.p.hetero) {
                static bool attr_h = false;
                if (!attr_h) {
                    cudaFuncSetAttribute(velocity_tiled<SO, T, ARITH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::smem(5));
                    attr_h = true;
                }
                velocity_tiled<SO, T, ARITH, true><<<grid, K::THREADS, K::smem(5), st>>>(R.tmap9[F_TXY], R.tmap9[F_TYY], R.tmap9[F_TYZ],
                                                                                         R.tmap9[F_TXZ], R.tmap9[F_TZZ], A);
                check();
                return;
            }
        }
        velocity_tiled<SO, T, ARITH><<<grid, K::THREADS, K::smem(5), st>>>(R.tmap9[F_TXY], R.tmap9[F_TYY], R.tmap9[F_TYZ], R.tmap9[F_TXZ],
                                                                             R.tmap9[F_TZZ], A);
        check();
    }

    template <int SO, typename T, int ARITH> void stress(int t0, int t1)
    {
        if (R.tiled) { stress_tiled_launch<SO, T, ARITH>(t0, t1); return; }
        dim3 blk(64, 4);
        if constexpr (sizeof(T) == 4) {
            if (R.M.p.hetero) {
                stress_interior_h<SO, ARITH><<<interior_grid(blk), blk, 0, st>>>(ptrs(), media(), R.M.G, R.M.hc, t0, t1);
                check();
                return;
            }
        }
        stress_interior<SO, T, ARITH><<<interior_grid(blk), blk, 0, st>>>(ptrs(), R.M.G, R.M.sc, t0, t1);
        check();
    }
    template <int SO, typename T, int ARITH> void velocity(int t0, int t1)
    {
        if (R.tiled) { velocity_tiled_launch<SO, T, ARITH>(t0, t1); return; }
        dim3 blk(64, 4);
        if constexpr (sizeof(T) == 4) {
            if (R.M.p.hetero) {
                velocity_interior_h<SO, ARITH><<<interior_grid(blk), blk, 0, st>>>(ptrs(), media(), R.M.G, R.M.hc, t0, t1);
                check();
                return;
            }
        }
        velocity_interior<SO, T, ARITH><<<interior_grid(blk), blk, 0, st>>>(ptrs(), R.M.G, R.M.sc, t0, t1);
        check();
    }
    template <int SO, typename T, int ARITH> void acoustic(int tprev, int tr, int tw, bool init)
    {
        dim3 blk(64, 4);
        if (init) {
            acoustic_interior<SO, T, ARITH, false><<<interior_grid(blk), blk, 0, st>>>(
                ptrs(), R.M.G, R.M.ac_init, tprev, tr, tw, (T)R.M.p.ac_init_const);
        } else {
            bool marched = false;
            if constexpr (sizeof(T) == 4) {
                // no stencil along the contiguous axis (the reference driver's PDE): float4 marching kernel
                if (!R.M.ac.present[2] && !(R.M.p.flags & OPESCI_FORCE_UNFUSED)) {
                    const Model &Md = R.M;
                    const int nx = Md.G.dim[0] - 2 * Md.m, ny = Md.G.dim[1] - 2 * Md.m;
                    const int nbz = (int)((Md.G.s[1] / 4 + 31) / 32), nby = (ny + 7) / 8;
                    int nchunks = (8 * sm_count() + nbz * nby - 1) / (nbz * nby);   // >= 8 blocks per SM in flight
                    if (nchunks < 1) nchunks = 1;
                    {
                        // Short chunks win (measured on B200 at 512^3, so=4, fast / reference arithmetic: 4 chunks 318 / 327 Gpts/s,
                        // 9: 336 / 381, 16: 353 / 383, 32: 351 / 385) although each chunk re-reads 2m planes of ONE of the three
                        // streams: the step is a fraction of a millisecond, and many short blocks keep the machine evenly filled
                        // to the end.  Chunks of about 32 planes, never shorter than 8m.
                        const int want = 32 > 8 * Md.m ? 32 : 8 * Md.m;
                        const int nc = nx / want;
                        if (nc > nchunks) nchunks = nc;
                        static const char *force = getenv("OPESCI_AC_CHUNKS");
                        if (force) nchunks = atoi(force);
                    }
                    if (nchunks > nx / (4 * Md.m) && nx / (4 * Md.m) >= 1) nchunks = nx / (4 * Md.m);
                    const int xchunk = (nx + nchunks - 1) / nchunks;
                    // so <= 4: the emitted-order sum is also the faster form here (512^3: 383 vs 342 Gpts/s, 1024^3: 411 vs 368), so
                    // "fast" arithmetic runs it too and is bit-identical to the reference; so >= 6: the factored form wins
                    constexpr int AR = SO <= 4 ? OPESCI_ARITH_REFERENCE : ARITH;
                    acoustic_march<SO, AR><<<dim3(nbz, nby, (nx + xchunk - 1) / xchunk), 256, 0, st>>>(ptrs(), Md.G, Md.ac, tprev, tr, tw, xchunk);
                    marched = true;
                }
            }
            if (!marched)
                acoustic_interior<SO, T, ARITH, true><<<interior_grid(blk), blk, 0, st>>>(ptrs(), R.M.G, R.M.ac, tprev, tr, tw, (T)0);
        }
        check();
    }

    // ---- batched ghost loops: loops that cannot observe each other run in one launch
    template <typename T> void launch_batch(FaceBatch &B)
    {
        if (B.count == 0) return;
        const Model &M = R.M;
        const bool het = M.p.hetero != 0;
        const int cpt = het ? 1 : OPESCI_FACE_CPT;
        int total = 0;
        for (int k = 0; k < B.count; ++k) {
            const FaceLoop &L = B.loop[k];
            const int w = (L.d == 2 ? 8 : 128) * cpt, h = OPESCI_FACE_THREADS / (L.d == 2 ? 8 : 128);
            B.nbx[k] = (L.hi2 - L.lo + w - 1) / w;
            B.start[k] = total;
            total += B.nbx[k] * ((L.hi1 - L.lo1 + h - 1) / h);
        }
        B.start[B.count] = total;
        if (het) face_batch<T, true, 1><<<total, OPESCI_FACE_THREADS, 0, st>>>(ptrs(), M.G, B, media());
        else face_batch<T, false, OPESCI_FACE_CPT><<<total, OPESCI_FACE_THREADS, 0, st>>>(ptrs(), M.G, B, media());
        check();
    }
    // loop ranges of one ghost loop: the reference uses [lo, dim - himargin) on both free axes
    // (staggeredgrid.py:785-796, 844-845).  On an artificial slab end the x range of the Levander
    // stress loops (lo = m+1) is widened by one plane: plane m / ldim-m-1 is an ordinary interior
    // plane there, and everything stays >= m planes away from the end of the local array.
    void set_ranges(FaceLoop &L, int d, int lo, int himargin) const
    {
        const Model &M = R.M;
        const int e1 = d == 0 ? 1 : 0, e2 = d == 2 ? 1 : 2;
        L.lo1 = lo; L.lo = lo; L.hi1 = M.G.dim[e1] - himargin; L.hi2 = M.G.dim[e2] - himargin;
        if (e1 == 0 && lo > 1) {
            if (!M.slab.lo_face) L.lo1 = lo - 1;
            if (!M.slab.hi_face) L.hi1 = M.G.dim[0] - himargin + 1;
        }
    }
    bool add_mirror(FaceBatch &B, int field, int level, int d, const MirrorOps &ops, int lo, int himargin) const
    {
        FaceLoop &L = B.loop[B.count];
        L.kind = 0; L.d = d; L.n = 0;
        set_ranges(L, d, lo, himargin);
        L.field = field; L.level = level; L.lv0 = L.lv1 = 0; L.ops = ops;
        if (L.hi1 <= L.lo1 || L.hi2 <= L.lo) return false;
        ++B.count;
        return true;
    }
    bool add_equation(FaceBatch &B, const DevEq &eq, int lv0, int lv1, int d, int n, int lo, int himargin) const
    {
        FaceLoop &L = B.loop[B.count];
        L.kind = 1; L.d = d; L.n = n;
        set_ranges(L, d, lo, himargin);
        L.field = 0; L.level = 0; L.lv0 = lv0; L.lv1 = lv1; L.eq = eq; L.ops.count = 0;
        if (L.hi1 <= L.lo1 || L.hi2 <= L.lo) return false;
        ++B.count;
        return true;
    }
    // x-face loops exist only where the slab end is a physical face
    // a face takes part if set_free_surface_boundary was called for it (opesci/staggeredgrid.py:214-232, 766-768) and,
    // along x, if this slab holds the physical face
    bool face_present(int d, int side) const
    {
        const int mask = R.M.p.fs_faces ? R.M.p.fs_faces : 63;
        if (R.M.p.free_surface == 0 || !((mask >> (2 * d + side)) & 1)) return false;
        return d != 0 || (side == 0 ? R.M.slab.lo_face : R.M.slab.hi_face);
    }
    // low and high side of one face pair never touch the same cells once the grid is this large
    bool sides_independent() const
    {
        const Model &M = R.M;
        // slabs: keep the reference's loop order literally.  (Measured: pairing the sides on slab ranks breaks the
        // bit-exact agreement with the single-domain run and gains nothing, 90.4 vs 90.6 Gpts/s on 2 GPUs.)
        if (M.slab.nranks > 1) return false;
        for (int d = 0; d < 3; ++d)
            if (M.G.dim[d] < 4 * M.m + 6) return false;
        return true;
    }

    // stress ghost loops (opesci/staggeredgrid.py:754-813).  The reference's order is field by
    // field (Txx,Tyy,Tzz,Txy,Tyz,Txz), face axis by face axis, low side then high side.  Loops of
    // different fields write different arrays and read only level-t0 data or their own array, so
    // the k-th loop of every field runs in launch k; within a field the order is kept.
    template <typename T> void stress_bc(int t0, int t1, bool init)
    {
        const Model &M = R.M;
        const int m = M.m;
        static const int ORDER[6] = {F_TXX, F_TYY, F_TZZ, F_TXY, F_TYZ, F_TXZ};
        static const int SH_A[3] = {0, 1, 0}, SH_B[3] = {1, 2, 2};
        const bool pair = sides_independent();
        // sequence of loops per field; one reference loop may be cut into several disjoint pieces (z-fold strips)
        std::vector<FaceLoop> seq[6][6];
        int nseq[6] = {0, 0, 0, 0, 0, 0};
        FaceBatch tmp;
        // z-fold: inside planes [zx0, zx1) x rows [m+1, dim2-m-1) the z-face loops are done by the z-edge tiles of the
        // fused kernel (fused.cuh, ZF); what is left of a [0,dim) x [0,dim) mirror loop are four thin strips
        const bool zf = R.zfold && !init;
        const int zx0 = M.slab.lo_face ? m + 1 : m, zx1 = M.slab.hi_face ? M.G.dim[0] - m - 1 : M.G.dim[0] - m;
        const int zy0 = m + 1, zy1 = M.G.dim[1] - m - 1;
        auto add_pieces = [&](std::vector<FaceLoop> &out, int field, int d, const MirrorOps &ops) {
            tmp.count = 0;
            if (!(zf && d == 2)) {
                if (add_mirror(tmp, field, t1, d, ops, 0, 0)) out.push_back(tmp.loop[0]);
                return;
            }
            const int box[4][4] = {{0, zx0, 0, M.G.dim[1]}, {zx1, M.G.dim[0], 0, M.G.dim[1]},
                                   {zx0, zx1, 0, zy0}, {zx0, zx1, zy1, M.G.dim[1]}};
            for (int k = 0; k < 4; ++k) {
                tmp.count = 0;
                if (!add_mirror(tmp, field, t1, d, ops, 0, 0)) continue;
                FaceLoop L = tmp.loop[0];
                L.lo1 = box[k][0]; L.hi1 = box[k][1]; L.lo = box[k][2]; L.hi2 = box[k][3];
                if (L.hi1 > L.lo1 && L.hi2 > L.lo) out.push_back(L);
            }
        };
        for (int fi = 0; fi < 6; ++fi)
            for (int d = 0; d < 3; ++d) {
                const int b_lo = m, b_hi = M.G.dim[d] - m - 1;
                for (int side = 0; side < 2; ++side) {
                    tmp.count = 0;
                    if (!face_present(d, side)) continue;
                    std::vector<FaceLoop> pieces;
                    if (fi < 3 && fi == d) {
                        // own-axis normal stress (opesci/fields.py:355-381), ranges [0,dim)
                        MirrorOps ops;
                        ops.count = 0;
                        const int b = side == 0 ? b_lo : b_hi, dir = side == 0 ? -1 : 1;
                        ops.dst[ops.count] = b; ops.src[ops.count++] = -1;
                        for (int k = 1; k <= m - 1; ++k) { ops.dst[ops.count] = b + dir * k; ops.src[ops.count++] = b - dir * k; }
                        add_pieces(pieces, ORDER[fi], d, ops);
                    } else if (fi < 3) {
                        // Levander recompute of the other normal stresses on this face from level t0
                        // (opesci/fields.py:313-353; not in the initial pass, staggeredgrid.py:771-773)
                        if (M.p.free_surface != 1 || init) continue;
                        if (!add_equation(tmp, M.lev_stress_eq[d][fi], t0, t1, d, side == 0 ? b_lo : b_hi, m + 1, m + 1)) continue;
                        if (zf && d == 2) {
                            // Its range IS the z-fold zone: done by the z-edge tiles -- except where an earlier loop of the
                            // same field still has to read the plain interior value (FusedArgs::zf_raw_x): Txx on the planes
                            // the x-face mirror reads, Tyy on the rows the y-face mirror reads.  Those lines are recomputed
                            // here, after that mirror, as in the reference.
                            const FaceLoop full = tmp.loop[0];
                            if (fi == 0) {
                                const int planes[2] = {M.slab.lo_face ? m + 1 : -1, M.slab.hi_face ? M.G.dim[0] - m - 2 : -1};
                                for (int k = 0; k < 2; ++k) {
                                    if (planes[k] < full.lo1 || planes[k] >= full.hi1 || (k == 1 && planes[1] == planes[0])) continue;
                                    FaceLoop L = full;
                                    L.lo1 = planes[k]; L.hi1 = planes[k] + 1;
                                    pieces.push_back(L);
                                }
                            } else {
                                const int rows[2] = {m + 1, M.G.dim[1] - m - 2};
                                for (int k = 0; k < 2; ++k) {
                                    if (rows[k] < full.lo || rows[k] >= full.hi2 || (k == 1 && rows[1] == rows[0])) continue;
                                    FaceLoop L = full;
                                    L.lo = rows[k]; L.hi2 = rows[k] + 1;
                                    pieces.push_back(L);
                                }
                            }
                        } else {
                            pieces.push_back(tmp.loop[0]);
                        }
                    } else {
                        const int a = SH_A[fi - 3], bb = SH_B[fi - 3];
                        if (d != a && d != bb) continue;
                        // antisymmetric mirror of a shear stress (opesci/fields.py:366-381), ranges [0,dim)
                        MirrorOps ops;
                        ops.count = 0;
                        for (int j = 0; j < m; ++j) {
                            if (side == 0) { ops.dst[ops.count] = m - 1 - j; ops.src[ops.count++] = m + j; }
                            else { ops.dst[ops.count] = b_hi + j; ops.src[ops.count++] = b_hi - 1 - j; }
                        }
                        add_pieces(pieces, ORDER[fi], d, ops);
                    }
                    // slot = face: launch k holds the loops of face k of every field, so the low / high pairing below always
                    // pairs the two sides of ONE axis, also when some faces carry no free surface
                    seq[fi][2 * d + side] = pieces;
                    nseq[fi] = 6;
                }
            }
        const int stride = pair ? 2 : 1;   // low + high side of a face pair together
        for (int k = 0; k < 6; k += stride) {
            FaceBatch B;
            B.count = 0;
            for (int fi = 0; fi < 6; ++fi)
                for (int j = k; j < k + stride && j < nseq[fi]; ++j)
                    for (const FaceLoop &L : seq[fi][j]) {
                        B.loop[B.count++] = L;
                        if (B.count == OPESCI_MAX_BATCH) { launch_batch<T>(B); B.count = 0; }   // pieces are disjoint: any split is fine
                    }
            launch_batch<T>(B);
        }
    }

    // velocity ghost loops (opesci/staggeredgrid.py:815-864): per face axis d the component
    // staggered along d first (both sides), then the two tangential components, which read the
    // freshly written normal ghosts but not each other.  Robertsson loops only write zeros.
    template <typename T> void velocity_bc(int t1)
    {
        const Model &M = R.M;
        const int m = M.m;
        const bool pair = sides_independent();
        FaceBatch robertsson;
        robertsson.count = 0;
        for (int d = 0; d < 3; ++d) {
            if (d == 2 && M.p.free_surface == 1 && !M.p.hetero && pair && !(M.p.flags & OPESCI_NO_ZFOLD) && face_present(2, 0) && face_present(2, 1)) {
                // both z faces, all three components, one launch (kernels.cuh: vel_zface_lev)
                VelZFaceArgs A;
                A.cn[0] = M.p.lev_vnormal[2][0]; A.cn[1] = M.p.lev_vnormal[2][1];
                A.gt[0] = M.p.lev_vtang[2][0]; A.gt[1] = M.p.lev_vtang[2][1];
                A.x0 = 1; A.x1 = M.G.dim[0] - 1; A.y0 = 1; A.y1 = M.G.dim[1] - 1;
                A.m = m; A.dimz = M.G.dim[2];
                dim3 grid((A.y1 - A.y0 + 31) / 32, (A.x1 - A.x0 + 7) / 8, 2);
                vel_zface_lev<T><<<grid, 256, 0, st>>>(ptrs(), M.G, (long long)t1 * M.G.level, A);
                check();
                continue;
            }
            int seq[3];
            seq[0] = d;
            for (int k = 0, n = 1; k < 3; ++k)
                if (k != d) seq[n++] = k;
            FaceBatch stage[2];
            stage[0].count = stage[1].count = 0;
            for (int si = 0; si < 3; ++si) {
                const int a = seq[si];
                for (int side = 0; side < 2; ++side) {
                    if (!face_present(d, side)) continue;
                    if (M.p.free_surface == 1) {
                        int n;
                        if (a == d) n = side == 0 ? m - 1 : M.G.dim[d] - m - 1;
                        else n = side == 0 ? m - 1 : M.G.dim[d] - m;
                        // every operand and the result live on level t1 (slot 0 of the equation)
                        FaceBatch &B = stage[si == 0 ? 0 : 1];
                        add_equation(B, M.lev_vel_eq[d][a][side], t1, t1, d, n, 1, 1);
                        if (!pair) { launch_batch<T>(B); B.count = 0; }
                    } else if (M.p.free_surface == 2) {
                        // Robertsson: m ghost layers := 0 (opesci/fields.py:243-259)
                        MirrorOps ops;
                        ops.count = 0;
                        for (int j = 0; j < m; ++j) {
                            int n;
                            if (side == 0) n = m - 1 - j;
                            else n = (a == d ? M.G.dim[d] - m - 1 : M.G.dim[d] - m) + j;
                            ops.dst[ops.count] = n; ops.src[ops.count++] = -1;
                        }
                        add_mirror(robertsson, VEL_OF_AXIS[a], t1, d, ops, 1, 1);
                        if (robertsson.count == OPESCI_MAX_BATCH) { launch_batch<T>(robertsson); robertsson.count = 0; }
                    }
                }
            }
            if (pair) { launch_batch<T>(stage[0]); launch_batch<T>(stage[1]); }
        }
        launch_batch<T>(robertsson);
    }

    // fused stress+velocity launch (fused.cuh); only instantiated for so <= 4, fp32
    template <int SO, typename T, int ARITH> void fused(int t0, int t1, int chunk0 = 0, int count = -1)
    {
        if constexpr (SO <= 4 && sizeof(T) == 4) {
            constexpr int M = SO / 2;
            using K = FusedCfg<M>;
            const Model &Md = R.M;
            FusedArgs A;
            A.F = ptrs(); A.G = Md.G; A.C = Md.sc; A.MD = media(); A.HC = Md.hc; A.t0 = t0; A.t1 = t1;
            for (int c = 0; c <= OPESCI_MAX_CHUNKS; ++c) A.xs[c] = R.xs[c];
            A.chunk0 = chunk0;
            if (count < 0) count = R.nchunks - chunk0;
            if (count <= 0) return;
            // z columns [M, zend) are covered by tiles; tile bx stores [bx*CZ + M - ZS, bx*CZ + M - ZS + CZ)
            const int zend = R.zstrip > 0 ? R.zstrip : Md.G.dim[2] - M;
            const int nzt = K::ztiles(zend + M);
            const int nyt = (Md.G.dim[1] - 2 * M + K::CY - 1) / K::CY;
            A.cluster_sync = 0;
            A.pace = nullptr;
            // z-edge + interior launch: two streams (fork / join with events; default) or one stream with a programmatic
            // dependency (OPESCI_ZF_PDL=1).  Measured on B200 at 1024^3: 21.50 vs 21.97 ms per step.
            static const bool zf_pdl = getenv("OPESCI_ZF_PDL") && atoi(getenv("OPESCI_ZF_PDL")) != 0;
            A.bx0 = 0; A.bxs[0] = A.bxs[1] = 0;
            A.zf_side[0] = A.zf_side[1] = 0; A.zf_c[0] = A.zf_c[1] = 0; A.zf_bx[0] = A.zf_bx[1] = -1; A.zf_xlo = A.zf_xhi = 0;
            A.zf_raw_x[0] = Md.slab.lo_face ? M + 1 : -1;
            A.zf_raw_x[1] = Md.slab.hi_face ? Md.G.dim[0] - M - 2 : -1;
            A.zf_raw_y[0] = M + 1; A.zf_raw_y[1] = Md.G.dim[1] - M - 2;
            for (int e = 0; e < 2; ++e)
                for (int f = 0; f < 2; ++f) { A.zf_lev[e][f][0] = Md.p.lev_stress[2][e][f][0]; A.zf_lev[e][f][1] = Md.p.lev_stress[2][e][f][1]; }
            if constexpr (SO == 4) {
                if (R.zfold) {
                    // ---- z-edge tile columns (0 and nzt-1) on the second stream, concurrently with the interior columns
                    FusedArgs E = A;
                    E.zf_xlo = Md.slab.lo_face ? M + 1 : M;
                    E.zf_xhi = Md.slab.hi_face ? Md.G.dim[0] - M - 1 : Md.G.dim[0] - M;
                    const int c_hi = (Md.G.dim[2] - M - 1) - (nzt - 1) * K::CZ;    // tile column of the high face plane
                    // one launch, grid.x = the z-edge tile columns: column 0 holds the low face, column nzt-1 the high one
                    const int ne = nzt == 1 ? 1 : 2;
                    E.bxs[0] = 0; E.bxs[1] = nzt - 1;
                    E.zf_side[0] = E.zf_side[1] = 1;
                    E.zf_bx[0] = 0; E.zf_bx[1] = nzt - 1;
                    E.zf_c[0] = M; E.zf_c[1] = c_hi;
                    // The z-edge CTAs are few (2 of nzt tile columns): cut their x-chunks finer than the interior ones, so
                    // that their last, partly filled wave is short (a wave of full-length chunks is ~1 ms at 1024^3)
                    static const int zf_sub_env = getenv("OPESCI_ZF_SUB") ? atoi(getenv("OPESCI_ZF_SUB")) : 0;
                    int sub = zf_sub_env > 0 ? zf_sub_env : OPESCI_ZF_SUB;
                    while (sub > 1 && (count * sub > OPESCI_MAX_CHUNKS || (A.xs[chunk0 + count] - A.xs[chunk0]) / (count * sub) < 16 * M)) --sub;
                    int ecount = 0;
                    for (int c = 0; c < count; ++c) {
                        const int lo = A.xs[chunk0 + c], hi = A.xs[chunk0 + c + 1];
                        for (int k = 0; k < sub; ++k) E.xs[ecount++] = lo + (int)((long long)(hi - lo) * k / sub);
                    }
                    E.xs[ecount] = A.xs[chunk0 + count];
                    E.chunk0 = 0;
                    cudaError_t e = cudaSuccess;
                    if (!zf_pdl) {
                        e = cudaEventRecord(R.ev_edge_fork, st);
                        if (e == cudaSuccess) e = cudaStreamWaitEvent(R.st_edge, R.ev_edge_fork, 0);
                        if (e != cudaSuccess && err == cudaSuccess) err = e;
                    }
                    if (fused_part != 1)
                    fused_step<SO, ARITH, false, true><<<dim3(ne, nyt, ecount), K::THREADS, K::SMEM, zf_pdl ? st : R.st_edge>>>(R.tmap[0], R.tmap[1], R.tmap[2], E
#if OPESCI_TMA_STORE
                        , R.smaps
#endif
                    );
                    check();
                    if (!zf_pdl) {
                        e = cudaEventRecord(R.ev_edge_join, R.st_edge);
                        if (e != cudaSuccess && err == cudaSuccess) err = e;
                    }
                    A.bx0 = 1;
                }
            }
            const int nmain = fused_part == 2 ? 0 : R.zfold ? nzt - 2 : nzt;
            dim3 grid(nmain > 0 ? nmain : 1, nyt, count);
            if (nmain > 0) {
#if OPESCI_PACE > 0
            {
                // every tile publishes the plane it is at; a tile more than OPESCI_PACE planes ahead of a running
                // y-neighbour waits, so the rows both of them read are still in L2 when the second one arrives
                const size_t nb = (size_t)grid.x * grid.y * OPESCI_MAX_CHUNKS * sizeof(int);   // allocated by setup_fused
                cudaMemsetAsync(R.d_pace, 0xFF, nb, st);   // -1: not started
                A.pace = R.d_pace;
            }
#endif
#if OPESCI_CLUSTER_Z > 1
            if (grid.x % OPESCI_CLUSTER_Z == 0) {
                A.cluster_sync = 1;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = grid; cfg.blockDim = dim3(K::THREADS); cfg.dynamicSmemBytes = K::SMEM; cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension;
                at[0].val.clusterDim.x = OPESCI_CLUSTER_Z; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
#if OPESCI_TMA_STORE
                cudaError_t e = Md.p.hetero ? cudaLaunchKernelEx(&cfg, fused_step<SO, ARITH, true>, R.tmap[0], R.tmap[1], R.tmap[2], A, R.smaps)
                                            : cudaLaunchKernelEx(&cfg, fused_step<SO, ARITH, false>, R.tmap[0], R.tmap[1], R.tmap[2], A, R.smaps);
#else
                cudaError_t e = Md.p.hetero ? cudaLaunchKernelEx(&cfg, fused_step<SO, ARITH, true>, R.tmap[0], R.tmap[1], R.tmap[2], A)
                                            : cudaLaunchKernelEx(&cfg, fused_step<SO, ARITH, false>, R.tmap[0], R.tmap[1], R.tmap[2], A);
#endif
                if (e != cudaSuccess && err == cudaSuccess) err = e;
            } else
#endif
#if OPESCI_TMA_STORE
            if (Md.p.hetero && !(SO == 4 && R.pair)) fused_step<SO, ARITH, true><<<grid, K::THREADS, K::SMEM, st>>>(R.tmap[0], R.tmap[1], R.tmap[2], A, R.smaps);
            else if (SO == 4 && R.pair) {
                if constexpr (SO == 4) {
                    // 2-CTA clusters stacked in y: a pair stores 2 * PCY rows
                    const int npairs = (Md.G.dim[1] - 2 * M + 2 * K::PCY - 1) / (2 * K::PCY);
                    cudaLaunchConfig_t cfg = {};
                    cfg.gridDim = dim3(grid.x, 2 * npairs, grid.z); cfg.blockDim = dim3(K::THREADS); cfg.dynamicSmemBytes = K::PSMEM; cfg.stream = st;
                    cudaLaunchAttribute at[1];
                    at[0].id = cudaLaunchAttributeClusterDimension;
                    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 2; at[0].val.clusterDim.z = 1;
                    cfg.attrs = at; cfg.numAttrs = 1;
                    cudaError_t e = Md.p.hetero ? cudaLaunchKernelEx(&cfg, fused_step<SO, ARITH, true, false, true>, R.tmap[0], R.tmap[1], R.tmap[2], A, R.smaps_pair)
                                                : cudaLaunchKernelEx(&cfg, fused_step<SO, ARITH, false, false, true>, R.tmap[0], R.tmap[1], R.tmap[2], A, R.smaps_pair);
                    if (e != cudaSuccess && err == cudaSuccess) err = e;
                }
            }
            else if (R.zfold && zf_pdl) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = grid; cfg.blockDim = dim3(K::THREADS); cfg.dynamicSmemBytes = K::SMEM; cfg.stream = st;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaError_t e = cudaLaunchKernelEx(&cfg, fused_step<SO, ARITH, false, false>, R.tmap[0], R.tmap[1], R.tmap[2], A, R.smaps);
                if (e != cudaSuccess && err == cudaSuccess) err = e;
            }
            else fused_step<SO, ARITH, false><<<grid, K::THREADS, K::SMEM, st>>>(R.tmap[0], R.tmap[1], R.tmap[2], A, R.smaps);
#else
            if (Md.p.hetero) fused_step<SO, ARITH, true><<<grid, K::THREADS, K::SMEM, st>>>(R.tmap[0], R.tmap[1], R.tmap[2], A);
            else fused_step<SO, ARITH, false><<<grid, K::THREADS, K::SMEM, st>>>(R.tmap[0], R.tmap[1], R.tmap[2], A);
#endif
            }
            if (nmain > 0) check();
            if (R.zfold && !zf_pdl) {
                cudaError_t e = cudaStreamWaitEvent(st, R.ev_edge_join, 0);
                if (e != cudaSuccess && err == cudaSuccess) err = e;
            }
            if (R.zstrip > 0) {
                // A few z columns are left over after the last full tile (1025 = 17 x 60 + 5 at 1024^3): a whole row of
                // nearly empty CTAs would cost as much as a full one (5.5 % of the kernel).  Their stresses are computed
                // per point here -- same arithmetic, same operands (level t0 only) -- and their velocities belong to
                // the shell update, whose z-high slab starts at zstrip (velocity_shell).
                const int xa = A.xs[chunk0], xb = A.xs[chunk0 + count];
                dim3 blk(8, 32);
                dim3 sg((Md.G.dim[2] - M - R.zstrip + blk.x - 1) / blk.x, (Md.G.dim[1] - 2 * M + blk.y - 1) / blk.y, xb - xa);
                if (Md.p.hetero) stress_interior_h<SO, ARITH><<<sg, blk, 0, st>>>(ptrs(), media(), Md.G, Md.hc, t0, t1, xa, R.zstrip);
                else stress_interior<SO, T, ARITH><<<sg, blk, 0, st>>>(ptrs(), Md.G, Md.sc, t0, t1, xa, R.zstrip);
                check();
            }
        }
    }
    // velocity update of the shell the fused kernel leaves out: interior minus [2m+1, dim-2m-1)^3,
    // six disjoint slabs in one launch
    template <int SO, typename T, int ARITH> void velocity_shell(int t0, int t1)
    {
        if constexpr (SO <= 4 && sizeof(T) == 4) {
            const Model &Md = R.M;
            const int m = Md.m;
            int lo[3], hi[3], ilo[3], ihi[3];
            for (int d = 0; d < 3; ++d) { lo[d] = m; hi[d] = Md.G.dim[d] - m; ilo[d] = 2 * m + 1; ihi[d] = Md.G.dim[d] - 2 * m - 1; }
            if (R.zstrip > 0 && R.zstrip < ihi[2]) ihi[2] = R.zstrip;   // the fused kernel stops at the z strip (fused())
            ShellBoxes B;
            int nb = 0, total = 0;
            for (int d = 0; d < 3; ++d)
                for (int side = 0; side < 2; ++side) {
                    Range3 rg;
                    const bool folded = R.zfold && d == 2;   // the z slabs belong to the z-edge tiles of the fused kernel
                    for (int e = 0; e < 3; ++e) {
                        if (e < d) { rg.lo[e] = ilo[e]; rg.hi[e] = ihi[e]; }     // already covered by earlier slabs
                        else if (e == d) { rg.lo[e] = side == 0 ? lo[e] : ihi[e]; rg.hi[e] = side == 0 ? ilo[e] : hi[e]; }
                        else { rg.lo[e] = lo[e]; rg.hi[e] = hi[e]; }
                    }
                    B.r[nb] = rg;
                    B.zwide[nb] = d == 2 ? 0 : 1;
                    const int tw = d == 2 ? 4 : 64, th = OPESCI_FACE_THREADS / tw;
                    const int nz = rg.hi[2] - rg.lo[2], ny = rg.hi[1] - rg.lo[1], nx = rg.hi[0] - rg.lo[0];
                    B.nbz[nb] = nz > 0 ? (nz + tw - 1) / tw : 1;
                    B.nby[nb] = ny > 0 ? (ny + th - 1) / th : 1;
                    B.start[nb] = total;
                    total += (nz > 0 && ny > 0 && nx > 0 && !folded) ? B.nbz[nb] * B.nby[nb] * nx : 0;
                    ++nb;
                }
            B.start[6] = total;
            if (total == 0) return;
            if (Md.p.hetero)
                velocity_shell_kernel<SO, T, ARITH, true><<<total, OPESCI_FACE_THREADS, 0, st>>>(ptrs(), Md.G, Md.sc, t0, t1, B, media(), Md.hc);
            else
                velocity_shell_kernel<SO, T, ARITH, false><<<total, OPESCI_FACE_THREADS, 0, st>>>(ptrs(), Md.G, Md.sc, t0, t1, B, media(), Md.hc);
            check();
        }
    }

    // receivers, then the source, at the end of a step (kernels.cuh); the device keeps the step count
    template <typename T> void point_hooks(int t1)
    {
        if (!R.hooks) return;
        const OpesciB200Params &p = R.M.p;
        const long long lvl = (long long)t1 * R.M.G.level;
        if (p.n_receivers > 0) {
            sample_receivers<T><<<(p.n_receivers + 127) / 128, 128, 0, st>>>(ptrs(), lvl, R.d_recv_cell, p.n_receivers, (T *)R.d_recv_out, R.d_step);
            check();
        }
        inject_source<T><<<1, 1, 0, st>>>(ptrs(), lvl, R.src_cell, R.d_src, R.d_src + p.src_nt, R.d_src + 2 * (size_t)p.src_nt, p.src_nt, R.d_step);
        check();
    }

    template <int SO, typename T, int ARITH> void staggered_step(int ti)
    {
        const int t0 = ti % 2, t1 = (t0 + 1) % 2;   // opesci/regulargrid.py:408-433
        if (R.fused) {
            mark();
            fused<SO, T, ARITH>(t0, t1);             // stress everywhere + velocity of the deep interior
            mark();
            stress_bc<T>(t0, t1, false);
            mark();
            velocity_shell<SO, T, ARITH>(t0, t1);    // velocity next to the faces, after the stress ghost loops
            mark();
            velocity_bc<T>(t1);
            mark(); mark();
        } else {
            stress<SO, T, ARITH>(t0, t1);
            stress_bc<T>(t0, t1, false);
            velocity<SO, T, ARITH>(t0, t1);
            velocity_bc<T>(t1);
        }
        point_hooks<T>(t1);
    }
    // OPESCI_KIND_REGULAR_GENERIC (generic.cuh): the run-time compiled kernels, one thread per interior point
    void generic_launch(CUfunction f, const int *levels, int nlevels)
    {
        const Model &M = R.M;
        const int m = M.m;
        if (!opesci_generic::launch(f, R.dev, M.p.nfields, levels, nlevels, M.G.dim[2] - 2 * m, M.G.dim[1] - 2 * m, M.G.dim[0] - 2 * m, st) &&
            err == cudaSuccess)
            err = cudaErrorLaunchFailure;
        check();
    }
    void generic_step(int ti)
    {
        const int t0 = ti % 3, t1 = (t0 + 1) % 3, t2 = (t1 + 1) % 3;   // opesci/regulargrid.py:408-433
        const int lv[3] = {t0, t1, t2};
        generic_launch(R.gen.step, lv, 3);
    }
    template <int SO, typename T, int ARITH> void acoustic_step(int ti)
    {
        if (R.M.p.kind == OPESCI_KIND_REGULAR_GENERIC) { generic_step(ti); return; }
        const int t0 = ti % 3, t1 = (t0 + 1) % 3, t2 = (t1 + 1) % 3;
        acoustic<SO, T, ARITH>(t0, t1, t2, false);
    }
};


// ------------------------------------------------------------------ fused path set-up
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int M, int ARITH> int set_fused_attr()
{
    CUDA_OK(cudaFuncSetAttribute(fused_step<2 * M, ARITH, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfg<M>::SMEM));
    CUDA_OK(cudaFuncSetAttribute(fused_step<2 * M, ARITH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfg<M>::SMEM));
    // No cudaFuncAttributePreferredSharedMemoryCarveout: with the whole 228 KB given to shared memory the kernel takes
    // 23.5 ms instead of 19.9 (measured) -- the T[t0] loads of the next plane (24.5 KB per CTA) land in L1, and the
    // default carve-out (196 KB for the 186 KB of rings) leaves them the 32 KB they need.
    return 0;
}

EncodeTiledFn get_encode()
{
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return nullptr;
        encode = (EncodeTiledFn)fn;
    }
    return encode;
}

// TMA-tiled two-pass kernels (tiled.cuh) for the staggered configurations the fused kernel does not cover
int setup_tiled(Run &R)
{
    const Model &M = R.M;
    const OpesciB200Params &p = M.p;
    R.tiled = false;
    if (p.kind != OPESCI_KIND_STAGGERED_ELASTIC || R.fused || (p.flags & OPESCI_FORCE_UNFUSED)) return 0;
    EncodeTiledFn encode = get_encode();
    if (!encode) return fail("cuTensorMapEncodeTiled not available from the driver");
    const int esz = p.is_double ? 8 : 4;
    // box = the tile the kernels expect (TileCfg<M,T>::VY x VZ): read from the same constexprs the kernels use, so the
    // tensor-map box and the kernels' expect_tx byte count cannot drift apart
    int VY = 0, VZ = 0;
    auto pick = [&](auto mtag) {
        constexpr int MM = decltype(mtag)::value;
        if (p.is_double) { VY = TileCfg<MM, double>::VY; VZ = TileCfg<MM, double>::VZ; }
        else { VY = TileCfg<MM, float>::VY; VZ = TileCfg<MM, float>::VZ; }
    };
    switch (M.m) {
    case 1: pick(std::integral_constant<int, 1>()); break;
    case 2: pick(std::integral_constant<int, 2>()); break;
    case 3: pick(std::integral_constant<int, 3>()); break;
    case 4: pick(std::integral_constant<int, 4>()); break;
    case 5: pick(std::integral_constant<int, 5>()); break;
    case 6: pick(std::integral_constant<int, 6>()); break;
    default: return fail("setup_tiled: unsupported margin");
    }
    for (int f = 0; f < 9; ++f) {
        cuuint64_t gdim[3] = {(cuuint64_t)p.dim[2], (cuuint64_t)p.dim[1], (cuuint64_t)M.G.dim[0] * p.nlevels};
        cuuint64_t gstride[2] = {(cuuint64_t)M.G.s[1] * esz, (cuuint64_t)M.G.s[0] * esz};
        cuuint32_t box[3] = {(cuuint32_t)VZ, (cuuint32_t)VY, 1};
        cuuint32_t estr[3] = {1, 1, 1};
        CUresult rc = encode(&R.tmap9[f], p.is_double ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, R.dev[f],
                             gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (tiled two-pass kernels)");
    }
    R.tiled = true;
    return 0;
}

int setup_fused(Run &R)
{
    const Model &M = R.M;
    const OpesciB200Params &p = M.p;
    R.fused = false;
    if (p.kind != OPESCI_KIND_STAGGERED_ELASTIC || p.is_double || p.so > 4 || (p.flags & (OPESCI_FORCE_UNFUSED | OPESCI_FORCE_TILED))) return 0;
    for (int d = 0; d < 3; ++d)
        if (M.G.dim[d] < 6 * M.m + 4) return 0;   // no deep interior worth fusing
    EncodeTiledFn encode = get_encode();
    if (!encode) return fail("cuTensorMapEncodeTiled not available from the driver");
    const int m = M.m;
    const int VZ = m == 1 ? FusedCfg<1>::VZ : FusedCfg<2>::VZ, VY = m == 1 ? FusedCfg<1>::VY : FusedCfg<2>::VY;
    for (int f = 0; f < 3; ++f) {
        cuuint64_t gdim[3] = {(cuuint64_t)p.dim[2], (cuuint64_t)p.dim[1], (cuuint64_t)M.G.dim[0] * p.nlevels};
        cuuint64_t gstride[2] = {(cuuint64_t)M.G.s[1] * 4, (cuuint64_t)M.G.s[0] * 4};
        cuuint32_t box[3] = {(cuuint32_t)VZ, (cuuint32_t)VY, 1};
        cuuint32_t estr[3] = {1, 1, 1};
#ifndef OPESCI_FUSED_L2PROMO
#define OPESCI_FUSED_L2PROMO CU_TENSOR_MAP_L2_PROMOTION_L2_128B
#endif
        CUresult rc = encode(&R.tmap[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, R.dev[f], gdim, gstride, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, OPESCI_FUSED_L2PROMO,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (rc != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed");
    }
#if OPESCI_TMA_STORE
    {
        // ring order of the published fields: Txy, Txz, Tyy, Tyz, Tzz
        const int ring_field[5] = {F_TXY, F_TXZ, F_TYY, F_TYZ, F_TZZ};
        for (int k = 0; k < 5; ++k) {
            // extents stop at dim - m: box elements beyond the last interior row / column are not written
            cuuint64_t gdim[3] = {(cuuint64_t)(p.dim[2] - m), (cuuint64_t)(p.dim[1] - m), (cuuint64_t)M.G.dim[0] * p.nlevels};
            cuuint64_t gstride[2] = {(cuuint64_t)M.G.s[1] * 4, (cuuint64_t)M.G.s[0] * 4};
            cuuint32_t box[3] = {(cuuint32_t)(m == 1 ? FusedCfg<1>::EZ : FusedCfg<2>::EZ), (cuuint32_t)(m == 1 ? FusedCfg<1>::CY : FusedCfg<2>::CY), 1};
            cuuint32_t estr[3] = {1, 1, 1};
            if (encode(&R.smaps.m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, R.dev[ring_field[k]], gdim, gstride, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
                return fail("cuTensorMapEncodeTiled failed (store maps)");
        }
    }
#endif
    if (m == 1) { if (set_fused_attr<1, OPESCI_ARITH_REFERENCE>() || set_fused_attr<1, OPESCI_ARITH_FAST>()) return 1; }
    else { if (set_fused_attr<2, OPESCI_ARITH_REFERENCE>() || set_fused_attr<2, OPESCI_ARITH_FAST>()) return 1; }
    // x-chunks: enough CTAs to fill the machine in whole waves, few enough to keep the 2m-plane
    // warm-up of every chunk negligible
    const int nsm = sm_count();
    const int CZ = m == 1 ? FusedCfg<1>::CZ : FusedCfg<2>::CZ, CY = m == 1 ? FusedCfg<1>::CY : FusedCfg<2>::CY;
    // z strip: when the last tile row would hold only a few columns, leave them to the per-point kernel
    R.zstrip = 0;
#ifndef OPESCI_ZSTRIP_MAX
#define OPESCI_ZSTRIP_MAX 0   /* measured on B200 at 1024^3: the strip kernel costs more (+0.55 ms fused, +0.24 ms shell) than the row of nearly empty tiles it removes; kept for A/B */
#endif
    const int ZS = m == 1 ? FusedCfg<1>::ZS : FusedCfg<2>::ZS;
    {
        const int nzint = p.dim[2] - 2 * m + ZS, rem = nzint % CZ;   // tile bx stores z in [bx*CZ + m - ZS, +CZ)
        if (rem > 0 && rem <= OPESCI_ZSTRIP_MAX && nzint / CZ >= 2) R.zstrip = m - ZS + (nzint / CZ) * CZ;
    }
    const int nztiles = ((R.zstrip > 0 ? R.zstrip : p.dim[2] - m) - m + ZS + CZ - 1) / CZ;
    const long long tiles = (long long)nztiles * ((p.dim[1] - 2 * m + CY - 1) / CY);
    const int nx = M.G.dim[0] - 2 * m;
    double best = -1.0;
    for (int nc = 1; nc <= 16; ++nc) {
        const int len = (nx + nc - 1) / nc;
        if (len < 8 * m && nc > 1) break;
        const double waves = (double)tiles * nc / nsm;
        const double eff = waves / (double)((long long)(waves + 0.999999)) * len / (len + 2.0 * m + 2.0);
        if (eff > best) { best = eff; R.nchunks = nc; }
    }
#ifdef OPESCI_FUSED_NCHUNKS
    R.nchunks = OPESCI_FUSED_NCHUNKS;   // A/B experiments (tools/ab.py)
#endif
    auto uniform = [&](int lo, int hi, int nc, int first) {   // chunks first .. first+nc-1 cover [lo, hi)
        const int len = (hi - lo + nc - 1) / nc;
        for (int c = 0; c <= nc; ++c) R.xs[first + c] = lo + c * len < hi ? lo + c * len : hi;
    };
    uniform(m, M.G.dim[0] - m, R.nchunks, 0);
    R.mid0 = 0; R.mid1 = 0;
    // Slabs, default schedule (run_model): the stress fields are exchanged right after the stress ghost loops and the
    // velocities after the velocity ghost loops, so ONE fused launch covers the whole slab.  The older schedule (the full
    // exchange overlapped with the middle chunks of the next step, thin end chunks behind it) is kept for runs with point
    // sources -- they write stress after the stress ghost loops -- and for A/B (OPESCI_SLAB_MIDOVERLAP=1).
    R.split_exchange = M.slab.nranks > 1 && !(p.n_receivers > 0 || p.src_nt > 0) &&
                       !(getenv("OPESCI_SLAB_MIDOVERLAP") && atoi(getenv("OPESCI_SLAB_MIDOVERLAP")) != 0);
    if (M.slab.nranks > 1 && !R.split_exchange) {
        // Slabs: the fused kernel at plane x reads planes x-2m .. x+2m-1 (+1).  Thin end chunks hold every plane whose
        // computation reads a halo plane; the chunks in between can run while the halo exchange of the previous step
        // is still in flight (run_model).
        const int E = M.slab.halo + 2 * m;                       // first / last plane index bound of the middle part
        const int lo = M.slab.lo_face ? m : E, hi = M.slab.hi_face ? M.G.dim[0] - m : M.G.dim[0] - E;
        if (hi - lo >= 8 * m) {
            int c = 0;
            if (!M.slab.lo_face) { R.xs[0] = m; R.xs[1] = lo; c = 1; }
            R.mid0 = c;
            const int nmid = R.nchunks < OPESCI_MAX_CHUNKS - 2 ? R.nchunks : OPESCI_MAX_CHUNKS - 2;
            uniform(lo, hi, nmid, c);
            c += nmid;
            R.mid1 = c;
            if (!M.slab.hi_face) { R.xs[c + 1] = M.G.dim[0] - m; ++c; }
            R.nchunks = c;
        }
    }
#if OPESCI_PACE > 0
    if (!R.d_pace) CUDA_OK(cudaMalloc(&R.d_pace, (size_t)tiles * OPESCI_MAX_CHUNKS * sizeof(int)));
#endif
    // ---- z-fold: so = 4 with the Levander free surface, homogeneous medium.  The high face plane b' = dim3-m-1 and the
    // columns its loops touch (b'-2 .. b'+2) must lie inside the last tile column, the z slab of the shell [b'-2, b'] in
    // its stored columns; otherwise (last column narrower than 3 stored cells) everything stays with the face kernels.
    R.zfold = false;
    const int fs_mask = p.fs_faces ? p.fs_faces : 63;
    if (m == 2 && p.free_surface == 1 && !p.hetero && R.zstrip == 0 && !(p.flags & OPESCI_NO_ZFOLD) && !getenv("OPESCI_NO_ZFOLD") &&
        ((fs_mask >> 4) & 3) == 3) {   // both z faces carry the free surface
        const int c_hi = (p.dim[2] - m - 1) - (nztiles - 1) * CZ;
        // (at least two tile columns: the z-edge kernel handles one face per column)
        if (nztiles >= 2 && c_hi >= 2 * m && c_hi + 2 <= FusedCfg<2>::EZ - 1 && p.dim[1] >= 4 * m + 6 && M.G.dim[0] >= 4 * m + 6) {
            int prio_lo = 0, prio_hi = 0;
            CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            // higher priority: the few z-edge CTAs are scheduled as soon as they are ready instead of behind every
            // pending CTA of the interior launch (they would otherwise form the tail of the step)
            if (!R.st_edge) CUDA_OK(cudaStreamCreateWithPriority(&R.st_edge, cudaStreamNonBlocking, prio_hi));
            if (!R.ev_edge_fork) CUDA_OK(cudaEventCreateWithFlags(&R.ev_edge_fork, cudaEventDisableTiming));
            if (!R.ev_edge_join) CUDA_OK(cudaEventCreateWithFlags(&R.ev_edge_join, cudaEventDisableTiming));
            CUDA_OK(cudaFuncSetAttribute(fused_step<4, OPESCI_ARITH_REFERENCE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfg<2>::SMEM));
            CUDA_OK(cudaFuncSetAttribute(fused_step<4, OPESCI_ARITH_FAST, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfg<2>::SMEM));
            R.zfold = true;
            R.zf_nzt = nztiles;
        }
    }
    // ---- pairs: the interior launch of the z-fold configuration as 2-CTA clusters stacked in y (fused.cuh, PAIR)
    R.pair = false;
    // (homogeneous: the interior columns beside the z-edge launch; heterogeneous: every column)
    if (m == 2 && (p.hetero || (R.zfold && nztiles >= 3)) && !(p.flags & OPESCI_NO_PAIR) && !getenv("OPESCI_NO_PAIR") && p.dim[1] >= 4 * m + 6) {
        const int ring_field[5] = {F_TXY, F_TXZ, F_TYY, F_TYZ, F_TZZ};
        bool ok = true;
        for (int k = 0; k < 5 && ok; ++k) {
            cuuint64_t gdim[3] = {(cuuint64_t)(p.dim[2] - m), (cuuint64_t)(p.dim[1] - m), (cuuint64_t)M.G.dim[0] * p.nlevels};
            cuuint64_t gstride[2] = {(cuuint64_t)M.G.s[1] * 4, (cuuint64_t)M.G.s[0] * 4};
            cuuint32_t box[3] = {(cuuint32_t)FusedCfg<2>::EZ, (cuuint32_t)FusedCfg<2>::PCY, 1};
            cuuint32_t estr[3] = {1, 1, 1};
            ok = encode(&R.smaps_pair.m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, R.dev[ring_field[k]], gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
        if (!ok) return fail("cuTensorMapEncodeTiled failed (pair store maps)");
        CUDA_OK(cudaFuncSetAttribute(fused_step<4, OPESCI_ARITH_REFERENCE, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfg<2>::PSMEM));
        CUDA_OK(cudaFuncSetAttribute(fused_step<4, OPESCI_ARITH_FAST, false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfg<2>::PSMEM));
        CUDA_OK(cudaFuncSetAttribute(fused_step<4, OPESCI_ARITH_REFERENCE, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfg<2>::PSMEM));
        CUDA_OK(cudaFuncSetAttribute(fused_step<4, OPESCI_ARITH_FAST, true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FusedCfg<2>::PSMEM));
        R.pair = true;
    }
    R.fused = true;
    return 0;
}

// ------------------------------------------------------------------ execute
template <int SO, typename T, int ARITH> int run_model(Run &R, cudaStream_t st, double *loop_seconds)
{
    const Model &M = R.M;
    const OpesciB200Params &p = M.p;
    Stepper S(R, st);
    // analytic initialisation of level 0 (staggeredgrid.py:612-659 / regulargrid.py:498-528)
    for (int f = 0; f < p.nfields; ++f) {
        Range3 rg;
        for (int d = 0; d < 3; ++d) { rg.lo[d] = p.fields[f].lo[d]; rg.hi[d] = p.fields[f].hi[d]; }
        // global x range -> planes stored by this rank (local index = global - L0; tables are pre-shifted)
        rg.lo[0] = (rg.lo[0] > M.slab.L0 ? rg.lo[0] : M.slab.L0) - M.slab.L0;
        rg.hi[0] = (rg.hi[0] < M.slab.L1 ? rg.hi[0] : M.slab.L1) - M.slab.L0;
        if (rg.hi[0] <= rg.lo[0] || rg.hi[1] <= rg.lo[1] || rg.hi[2] <= rg.lo[2]) continue;
        dim3 blk(64, 4);
        dim3 grid((rg.hi[2] - rg.lo[2] + blk.x - 1) / blk.x, (rg.hi[1] - rg.lo[1] + blk.y - 1) / blk.y, rg.hi[0] - rg.lo[0]);
        init_field<T><<<grid, blk, 0, st>>>((T *)R.dev[f], M.G, rg, R.d_prog + 2 * f);
        S.check();
    }
    const bool staggered = p.kind == OPESCI_KIND_STAGGERED_ELASTIC;
    const int period = staggered ? 2 : 3;
    if (staggered) {
        S.template stress_bc<T>(0, 0, true);   // initialise_bc (staggeredgrid.py:866-879)
        S.template velocity_bc<T>(0);
    } else if (p.kind == OPESCI_KIND_REGULAR_GENERIC) {
        // second initialisation of every field (regulargrid.py:530-564; the reference emits it once per field --
        // it reads level 0 only, so once is the same)
        const int lv[2] = {0, 1};
        S.generic_launch(R.gen.init2, lv, 2);
    } else {
        S.template acoustic<SO, T, ARITH>(0, 0, 1, true);   // second initialisation: level 1 from level 0
    }
    const bool slabs = M.slab.nranks > 1;
    // ---- peer-memory halo transport (one process per GPU over NCCL): every rank exports its field allocations with
    // cudaIpcGetMemHandle, maps its neighbours' and PULLS their owned planes into its halo planes with plain device
    // copies -- the copy engines move them over NVLink at link rate and no SM is taken from the kernels running beside
    // the transfer (ncclSend/Recv of the same planes reached ~180 GB/s per direction next to the fused kernel).  NCCL
    // only carries a 4-byte token per neighbour and exchange: my receive completing means the neighbour's stream has
    // reached the same exchange, i.e. its planes of this level are final; and because a rank enters exchange k+1 only
    // after its pulls of exchange k have completed (stream order), the same token releases the planes read in exchange
    // k for overwriting.  OPESCI_HALO_P2P=0 keeps ncclSend/Recv for the planes themselves.
    struct PeerHalo {
        bool on = false;
        void *peer[2][OPESCI_MAX_FIELDS];      // [side][field]: the neighbour's allocation, mapped into this process
        long long level[2] = {0, 0};           // the neighbour's level stride (elements)
        int L0[2] = {0, 0};                    // first plane the neighbour stores
        int *d_tok = nullptr;                  // tokens: [0,1] sent to lo / hi, [2,3] received
        PeerHalo() { for (auto &sd : peer) for (void *&q : sd) q = nullptr; }
        void close()
        {
            for (auto &sd : peer)
                for (void *&q : sd)
                    if (q) { cudaIpcCloseMemHandle(q); q = nullptr; }
            on = false;
        }
        ~PeerHalo() { close(); if (d_tok) cudaFree(d_tok); }
    } PH;
    auto tokens = [&](cudaStream_t xs_) -> int {
        const OpesciSlab &sl = M.slab;
        NCCL_OK(g_nccl.GroupStart());
        if (!sl.lo_face) {
            NCCL_OK(g_nccl.Send(PH.d_tok + 0, sizeof(int), ncclUint8, sl.rank - 1, g_nccl.comm, xs_));
            NCCL_OK(g_nccl.Recv(PH.d_tok + 2, sizeof(int), ncclUint8, sl.rank - 1, g_nccl.comm, xs_));
        }
        if (!sl.hi_face) {
            NCCL_OK(g_nccl.Send(PH.d_tok + 1, sizeof(int), ncclUint8, sl.rank + 1, g_nccl.comm, xs_));
            NCCL_OK(g_nccl.Recv(PH.d_tok + 3, sizeof(int), ncclUint8, sl.rank + 1, g_nccl.comm, xs_));
        }
        NCCL_OK(g_nccl.GroupEnd());
        return 0;
    };
    tl_halo_transport = !slabs ? 0 : tl_loop ? 3 : 1;
    if (slabs && !tl_loop && !(getenv("OPESCI_HALO_P2P") && atoi(getenv("OPESCI_HALO_P2P")) == 0)) {
        struct Pack { cudaIpcMemHandle_t h[OPESCI_MAX_FIELDS]; int ok; };
        const OpesciSlab &sl = M.slab;
        Pack mine, theirs[2];
        memset(&mine, 0, sizeof mine);
        memset(theirs, 0, sizeof theirs);
        mine.ok = 1;
        for (int f = 0; f < p.nfields; ++f)
            if (cudaIpcGetMemHandle(&mine.h[f], R.dev[f]) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
        Pack *d_pack = nullptr;
        double *d_bad = nullptr;
        CUDA_OK(cudaMalloc(&d_pack, 3 * sizeof(Pack)));
        CUDA_OK(cudaMalloc(&d_bad, sizeof(double)));
        CUDA_OK(cudaMalloc(&PH.d_tok, 4 * sizeof(int)));
        CUDA_OK(cudaMemsetAsync(PH.d_tok, 0, 4 * sizeof(int), st));
        CUDA_OK(cudaMemcpyAsync(d_pack, &mine, sizeof(Pack), cudaMemcpyHostToDevice, st));
        NCCL_OK(g_nccl.GroupStart());
        if (!sl.lo_face) {
            NCCL_OK(g_nccl.Send(d_pack, sizeof(Pack), ncclUint8, sl.rank - 1, g_nccl.comm, st));
            NCCL_OK(g_nccl.Recv(d_pack + 1, sizeof(Pack), ncclUint8, sl.rank - 1, g_nccl.comm, st));
        }
        if (!sl.hi_face) {
            NCCL_OK(g_nccl.Send(d_pack, sizeof(Pack), ncclUint8, sl.rank + 1, g_nccl.comm, st));
            NCCL_OK(g_nccl.Recv(d_pack + 2, sizeof(Pack), ncclUint8, sl.rank + 1, g_nccl.comm, st));
        }
        NCCL_OK(g_nccl.GroupEnd());
        CUDA_OK(cudaMemcpyAsync(theirs, d_pack + 1, 2 * sizeof(Pack), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        double bad = mine.ok ? 0.0 : 1.0;
        const char *why = mine.ok ? "" : "cudaIpcGetMemHandle failed";
        for (int side = 0; side < 2 && bad == 0.0; ++side) {
            if (side == 0 ? sl.lo_face : sl.hi_face) continue;
            if (!theirs[side].ok) { bad = 1.0; why = "a neighbour could not export its fields"; break; }
            OpesciSlab ns;
            opesci_slab_make(&ns, sl.rank + (side == 0 ? -1 : 1), sl.nranks, sl.gdim, M.m, OPESCI_SLAB_HALO, sl.halo);
            PH.L0[side] = ns.L0;
            PH.level[side] = (long long)(ns.L1 - ns.L0) * M.G.s[0];
            for (int f = 0; f < p.nfields; ++f)
                if (cudaIpcOpenMemHandle(&PH.peer[side][f], theirs[side].h[f], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                    PH.peer[side][f] = nullptr; cudaGetLastError(); bad = 1.0; why = "cudaIpcOpenMemHandle failed"; break;
                }
        }
        // all ranks use the same transport: the sum of the failure flags decides
        const double my_bad = bad;
        CUDA_OK(cudaMemcpyAsync(d_bad, &bad, sizeof(double), cudaMemcpyHostToDevice, st));
        NCCL_OK(g_nccl.AllReduce(d_bad, d_bad, 1, ncclFloat64, ncclSum, g_nccl.comm, st));
        CUDA_OK(cudaMemcpyAsync(&bad, d_bad, sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_OK(cudaStreamSynchronize(st));
        cudaFree(d_pack);
        cudaFree(d_bad);
        if (bad == 0.0) { PH.on = true; tl_halo_transport = 2; }
        else {
            PH.close();
            if (my_bad != 0.0 || sl.rank == 0)
                fprintf(stderr, "[opesci_b200] rank %d: peer-memory halo transport not available (%s); the planes travel by ncclSend/Recv\n",
                        sl.rank, my_bad != 0.0 ? why : "another rank could not map its neighbour");
        }
    }
    // halo refresh: fields [f0, f1), one time level, H planes per inner side (contiguous blocks)
    auto exchange = [&](int level, cudaStream_t xs_, int f0 = 0, int f1 = -1) -> int {
        cudaStream_t st = xs_;
        if (f1 < 0) f1 = p.nfields;
        const OpesciSlab &sl = M.slab;
        const size_t plane = (size_t)M.G.s[0], nel = (size_t)sl.halo * plane;
        if (tl_loop) {
            // loopback transport: same planes, same direction, same stream position as the NCCL group below
            Loopback &L = *tl_loop;
            const int r = sl.rank;
            CUDA_OK(cudaEventRecord(L.done[r], st));                 // my planes of this level are final
            if (L.barrier.wait()) return fail("loopback slabs: a peer rank failed");
            for (int side = 0; side < 2; ++side) {
                const int nb = side == 0 ? r - 1 : r + 1;
                if (nb < 0 || nb >= L.nranks) continue;
                const Run &N = *L.runs[nb];
                const OpesciSlab &ns = N.M.slab;
                CUDA_OK(cudaStreamWaitEvent(st, L.done[nb], 0));
                for (int f = f0; f < f1; ++f) {
                    T *mine = (T *)R.dev[f] + (size_t)level * M.G.level;
                    const T *theirs = (const T *)N.dev[f] + (size_t)level * N.M.G.level;
                    // low halo [X0-halo, X0) = the neighbour's last owned planes; high halo [X1, X1+halo) = its first ones
                    const size_t dst = side == 0 ? 0 : (size_t)(sl.X1 - sl.L0);
                    const size_t src = side == 0 ? (size_t)(sl.X0 - sl.halo - ns.L0) : (size_t)(sl.X1 - ns.L0);
                    CUDA_OK(cudaMemcpyAsync(mine + dst * plane, theirs + src * plane, nel * sizeof(T), cudaMemcpyDeviceToDevice, st));
                }
            }
            CUDA_OK(cudaEventRecord(L.xdone[r], st));                // I have read my neighbours' planes
            if (L.barrier.wait()) return fail("loopback slabs: a peer rank failed");
            for (int nb = r - 1; nb <= r + 1; nb += 2)
                if (nb >= 0 && nb < L.nranks) CUDA_OK(cudaStreamWaitEvent(st, L.xdone[nb], 0));
            return 0;
        }
        if (PH.on) {
            if (tokens(st)) return 1;
            for (int side = 0; side < 2; ++side) {
                if (side == 0 ? sl.lo_face : sl.hi_face) continue;
                // low halo [X0-halo, X0) = the neighbour's last owned planes; high halo [X1, X1+halo) = its first ones
                const size_t dst = side == 0 ? 0 : (size_t)(sl.X1 - sl.L0);
                const size_t src = side == 0 ? (size_t)(sl.X0 - sl.halo - PH.L0[0]) : (size_t)(sl.X1 - PH.L0[1]);
                for (int f = f0; f < f1; ++f) {
                    T *mine = (T *)R.dev[f] + (size_t)level * M.G.level;
                    const T *theirs = (const T *)PH.peer[side][f] + (size_t)level * PH.level[side];
                    CUDA_OK(cudaMemcpyAsync(mine + dst * plane, theirs + src * plane, nel * sizeof(T), cudaMemcpyDefault, st));
                }
            }
            return 0;
        }
        NCCL_OK(g_nccl.GroupStart());
        for (int f = f0; f < f1; ++f) {
            T *base = (T *)R.dev[f] + (size_t)level * M.G.level;
            if (!sl.lo_face) {
                NCCL_OK(g_nccl.Send(base + (size_t)(sl.X0 - sl.L0) * plane, nel * sizeof(T), ncclUint8, sl.rank - 1, g_nccl.comm, st));
                NCCL_OK(g_nccl.Recv(base, nel * sizeof(T), ncclUint8, sl.rank - 1, g_nccl.comm, st));
            }
            if (!sl.hi_face) {
                NCCL_OK(g_nccl.Send(base + (size_t)(sl.X1 - sl.halo - sl.L0) * plane, nel * sizeof(T), ncclUint8, sl.rank + 1, g_nccl.comm, st));
                NCCL_OK(g_nccl.Recv(base + (size_t)(sl.X1 - sl.L0) * plane, nel * sizeof(T), ncclUint8, sl.rank + 1, g_nccl.comm, st));
            }
        }
        NCCL_OK(g_nccl.GroupEnd());
        return 0;
    };
    // staggered: level 0 after the initial BC pass; regular: level 1 (level 0 is analytic on every stored plane)
    if (slabs && exchange(staggered ? 0 : 1, st)) return 1;
    CUDA_OK(cudaStreamSynchronize(st));
    if (S.err != cudaSuccess) return fail("kernel launch failed during initialisation: %s", cudaGetErrorString(S.err));

    // per-step field output (include/opesci_io.h: opesci_b200_set_output; reference regulargrid.py:702-719)
    opesci_io::Snapshotter snap;
    if (opesci_io::output_cfg().armed) {
        if (opesci_io::output_cfg().field >= p.nfields) return fail("opesci_b200_set_output: no such field");
        const int ldims[3] = {M.G.dim[0], p.dim[1], p.dim[2]};
        // slabs: the planes this rank owns, physical ghost planes of the end ranks included -- the pieces tile the array
        const int own_lo = slabs ? M.slab.own_lo - M.slab.L0 : 0, own_hi = slabs ? M.slab.own_hi - M.slab.L0 : M.G.dim[0];
        const char *e = snap.init(R.dev[opesci_io::output_cfg().field], sizeof(T), (size_t)M.G.level, (size_t)M.G.s[1], ldims, own_lo, own_hi,
                                  slabs ? M.slab.own_lo : 0, p.dx, M.slab.rank, slabs ? M.slab.nranks : 1);
        if (e) return fail("%s", e);
    }
#define SNAP_OK(call) do { const char *e_ = (call); if (e_) return fail("%s", e_); } while (0)
    // every handle of the time loop lives in one holder whose destructor runs on each early `return` of the error
    // macros too: a capture left open is ended, graphs / events / the second stream are destroyed
    struct LoopHandles {
        cudaStream_t cap = nullptr;     // stream with a capture in progress
        cudaEvent_t e0 = nullptr, e1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t gexec = nullptr;
        cudaStream_t st2 = nullptr;
        ~LoopHandles()
        {
            if (cap) { cudaGraph_t g = nullptr; cudaStreamEndCapture(cap, &g); if (g) cudaGraphDestroy(g); cudaGetLastError(); }
            if (gexec) cudaGraphExecDestroy(gexec);
            if (graph) cudaGraphDestroy(graph);
            if (e0) cudaEventDestroy(e0);
            if (e1) cudaEventDestroy(e1);
            if (ev_fork) cudaEventDestroy(ev_fork);
            if (ev_join) cudaEventDestroy(ev_join);
            if (st2) cudaStreamDestroy(st2);
        }
    } H;
    cudaEvent_t &e0 = H.e0, &e1 = H.e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    PhaseTrace trace;
    trace.on = staggered && getenv("OPESCI_STEP_TRACE") && atoi(getenv("OPESCI_STEP_TRACE")) != 0;
    const int nsteps = p.ntsteps;
    int warm = p.warmup_steps > 0 ? p.warmup_steps : 0;   // the first `warmup_steps` steps run untimed (bench contract)
    if (warm > nsteps) warm = nsteps;
    long long per_period = 0;
    cudaGraph_t &graph = H.graph;
    cudaGraphExec_t &gexec = H.gexec;
    cudaStream_t &st2 = H.st2;
    cudaEvent_t &ev_fork = H.ev_fork, &ev_join = H.ev_join;
    if (slabs && staggered && R.fused && R.split_exchange) {
        // ---- slabs, fused kernel, default: one fused launch per step over the whole slab.  The six stress fields of the
        // new level travel (own high-priority stream) as soon as the stress ghost loops are done, hidden behind the
        // velocity shell and the velocity ghost loops; the three velocities follow after those -- the only transfer the
        // next step waits for.  The kernels that run beside the stress transfer read stress halo planes, but every plane
        // an OWNED cell's computation reads (local index >= m) already holds, bit for bit, what the neighbour sends
        // (include/opesci_slab.h), and the velocity they leave on the halo planes is replaced by the second transfer.
        int prio_lo = 0, prio_hi = 0;
        CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_OK(cudaStreamCreateWithPriority(&st2, cudaStreamNonBlocking, prio_hi));
        CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));   // work on st done -> transfer may start
        CUDA_OK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));   // both transfers done -> halo planes valid
        auto step = [&](int ti) -> int {
            const int t0 = ti % 2, t1 = (t0 + 1) % 2;
            S.mark();
            SNAP_OK(snap.before_step(st, ti));
            S.template fused<SO, T, ARITH>(t0, t1);
            S.mark();
            S.template stress_bc<T>(t0, t1, false);
            S.mark();
            CUDA_OK(cudaEventRecord(ev_fork, st));
            CUDA_OK(cudaStreamWaitEvent(st2, ev_fork, 0));
            if (exchange(t1, st2, F_TXX, p.nfields)) return 1;
            S.template velocity_shell<SO, T, ARITH>(t0, t1);
            S.mark();
            S.template velocity_bc<T>(t1);
            S.mark();
            SNAP_OK(snap.after_step(st, ti, t1));   // owned planes only: they are final before the halo exchange
            CUDA_OK(cudaEventRecord(ev_fork, st));
            CUDA_OK(cudaStreamWaitEvent(st2, ev_fork, 0));
            if (exchange(t1, st2, F_U, F_TXX)) return 1;
            CUDA_OK(cudaEventRecord(ev_join, st2));
            CUDA_OK(cudaStreamWaitEvent(st, ev_join, 0));   // the next step starts with valid halo planes
            S.mark();
            return 0;
        };
        // Two steps (the time-level indices repeat with period 2) are captured into one CUDA graph and replayed, like the
        // single-GPU loop: both streams, the token kernels and the peer copies (or the ncclSend/Recv pairs) become nodes
        // of it -- measured on one GPU the replayed graph saves 0.56 ms of a 21.1 ms step.  Not with loopback ranks
        // (host barriers inside the exchange), per-step output or the phase trace.
        const bool slab_graph = !tl_loop && !(p.flags & OPESCI_NO_CUDA_GRAPH) && !getenv("OPESCI_NO_CUDA_GRAPH") && !snap.armed && !trace.on &&
                                nsteps - warm >= 2 * period;
        int ti = 0;
        for (; ti < warm; ++ti)
            if (step(ti)) return 1;
        int graph_parity = 0;
        if (slab_graph) {
            // (NCCL's connections to both neighbours exist already: the exchange after the initialisation made them)
            CUDA_OK(cudaStreamSynchronize(st));
            graph_parity = ti % 2;
            const long long before = S.launches;
            CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            H.cap = st;
            if (step(ti) || step(ti + 1)) return 1;
            per_period = S.launches - before;
            H.cap = nullptr;
            CUDA_OK(cudaStreamEndCapture(st, &graph));
            CUDA_OK(cudaGraphInstantiate(&gexec, graph, 0));
            S.launches = before;
        }
        if (warm > 0) S.launches = 0;
        if (trace.on) S.trace = &trace;
        CUDA_OK(cudaEventRecord(e0, st));
        while (ti < nsteps) {
            if (gexec && ti % 2 == graph_parity && ti + 2 <= nsteps) {
                CUDA_OK(cudaGraphLaunch(gexec, st));
                S.launches += per_period;
                ti += 2;
            } else {
                if (step(ti)) return 1;
                ++ti;
            }
        }
        CUDA_OK(cudaEventRecord(e1, st));
        CUDA_OK(cudaStreamSynchronize(st));
        CUDA_OK(cudaStreamSynchronize(st2));
    } else if (slabs && staggered && R.fused && R.mid1 > R.mid0) {
        // ---- slabs, fused kernel: the halo exchange of step n (NCCL, own high-priority stream) runs while the middle
        // x-chunks of step n+1 -- which read no halo plane -- are computed; the thin end chunks, the ghost loops and
        // the shell follow once the exchange has landed.  Same kernels, same per-cell order as the serial schedule.
        int prio_lo = 0, prio_hi = 0;
        CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUDA_OK(cudaStreamCreateWithPriority(&st2, cudaStreamNonBlocking, prio_hi));
        CUDA_OK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));   // step done on st -> exchange may start
        CUDA_OK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));   // exchange done -> halo planes valid
        bool pending = false;
        auto step = [&](int ti) -> int {
            const int t0 = ti % 2, t1 = (t0 + 1) % 2;
            SNAP_OK(snap.before_step(st, ti));
            S.template fused<SO, T, ARITH>(t0, t1, R.mid0, R.mid1 - R.mid0);
            if (pending) { CUDA_OK(cudaStreamWaitEvent(st, ev_join, 0)); pending = false; }
            S.template fused<SO, T, ARITH>(t0, t1, 0, R.mid0);
            S.template fused<SO, T, ARITH>(t0, t1, R.mid1, R.nchunks - R.mid1);
            S.template stress_bc<T>(t0, t1, false);
            S.template velocity_shell<SO, T, ARITH>(t0, t1);
            S.template velocity_bc<T>(t1);
            S.template point_hooks<T>(t1);
            SNAP_OK(snap.after_step(st, ti, t1));   // owned planes only: they are final before the halo exchange
            CUDA_OK(cudaEventRecord(ev_fork, st));
            CUDA_OK(cudaStreamWaitEvent(st2, ev_fork, 0));
            if (exchange(t1, st2)) return 1;
            CUDA_OK(cudaEventRecord(ev_join, st2));
            pending = true;
            return 0;
        };
        auto drain = [&]() -> int {
            if (pending) { CUDA_OK(cudaStreamWaitEvent(st, ev_join, 0)); pending = false; }
            return 0;
        };
        for (int ti = 0; ti < warm; ++ti)
            if (step(ti)) return 1;
        if (drain()) return 1;
        S.launches = 0;
        CUDA_OK(cudaEventRecord(e0, st));
        for (int ti = warm; ti < nsteps; ++ti)
            if (step(ti)) return 1;
        if (drain()) return 1;
        CUDA_OK(cudaEventRecord(e1, st));
        CUDA_OK(cudaStreamSynchronize(st));
        CUDA_OK(cudaStreamSynchronize(st2));
    } else {
    const bool use_graph = !(p.flags & OPESCI_NO_CUDA_GRAPH) && !getenv("OPESCI_NO_CUDA_GRAPH") && nsteps >= 2 * period && !slabs && !snap.armed && !trace.on;
    if (use_graph) {
        // one period of steps (time-level indices repeat with it) captured once, replayed
        CUDA_OK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
        H.cap = st;
        const long long before = S.launches;
        for (int ti = 0; ti < period; ++ti) {
            if (staggered) S.template staggered_step<SO, T, ARITH>(ti);
            else S.template acoustic_step<SO, T, ARITH>(ti);
        }
        per_period = S.launches - before;
        H.cap = nullptr;
        CUDA_OK(cudaStreamEndCapture(st, &graph));
        CUDA_OK(cudaGraphInstantiate(&gexec, graph, 0));
        S.launches = before;
    }
    int ti = 0;
    auto run_steps = [&](int upto) -> int {
        while (ti < upto) {
            if (use_graph && ti % period == 0 && ti + period <= upto) {
                CUDA_OK(cudaGraphLaunch(gexec, st));
                S.launches += per_period;
                ti += period;
            } else {
                SNAP_OK(snap.before_step(st, ti));
                if (staggered) S.template staggered_step<SO, T, ARITH>(ti);
                else S.template acoustic_step<SO, T, ARITH>(ti);
                // the level this step wrote: t1 = (ti+1)%2 (staggered), t2 = (ti+2)%3 (regular)
                SNAP_OK(snap.after_step(st, ti, staggered ? (ti + 1) % 2 : (ti + 2) % 3));
                if (slabs && exchange(staggered ? (ti + 1) % 2 : (ti + 2) % 3, st)) return 1;
                ++ti;
            }
        }
        return 0;
    };
    if (warm > 0) {
        if (run_steps(warm)) return 1;
        S.launches = 0;
    }
    if (trace.on) S.trace = &trace;
    CUDA_OK(cudaEventRecord(e0, st));
    if (run_steps(nsteps)) return 1;
    CUDA_OK(cudaEventRecord(e1, st));
    }
    CUDA_OK(cudaStreamSynchronize(st));
    if (PH.on) {
        // my pulls are complete (both streams were synchronised): unmap the neighbours' fields, then meet them once more
        // so that nobody frees an allocation a neighbour still has mapped or is still reading
        if (st2) CUDA_OK(cudaStreamSynchronize(st2));
        PH.close();
        if (tokens(st)) return 1;
        CUDA_OK(cudaStreamSynchronize(st));
    }
    S.trace = nullptr;
    trace.report(M.slab.rank);
    snap.finish();
    if (opesci_io::write_errors().load() > 0) return fail("per-step field output: writing a .vts file failed");
    if (S.err != cudaSuccess) return fail("kernel launch failed in the time loop: %s", cudaGetErrorString(S.err));
    CUDA_OK(cudaGetLastError());
    float ms = 0.f;
    CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    *loop_seconds = ms * 1e-3;
    g_launches = S.launches;
    return 0;   // ~LoopHandles releases the graph, events and the second stream
}

template <typename T, int ARITH> int dispatch_so(Run &R, cudaStream_t st, double *secs)
{
    switch (R.M.p.so) {
    case 2: return run_model<2, T, ARITH>(R, st, secs);
    case 4: return run_model<4, T, ARITH>(R, st, secs);
    case 6: return run_model<6, T, ARITH>(R, st, secs);
    case 8: return run_model<8, T, ARITH>(R, st, secs);
    case 10: return run_model<10, T, ARITH>(R, st, secs);
    case 12: return run_model<12, T, ARITH>(R, st, secs);
    }
    return fail("unsupported spatial order");
}

int dispatch(Run &R, cudaStream_t st, double *secs)
{
    const bool fast = (R.M.p.flags & OPESCI_ARITH_MASK) == OPESCI_ARITH_FAST;
    if (R.M.p.is_double)
        return fast ? dispatch_so<double, OPESCI_ARITH_FAST>(R, st, secs) : dispatch_so<double, OPESCI_ARITH_REFERENCE>(R, st, secs);
    return fast ? dispatch_so<float, OPESCI_ARITH_FAST>(R, st, secs) : dispatch_so<float, OPESCI_ARITH_REFERENCE>(R, st, secs);
}

template <int SO, typename T, int ARITH> int time_kernels_impl(Run &R, int reps, double *out_ms, int part = 0)
{
    cudaStream_t st;
    CUDA_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    Stepper S(R, st);
    S.fused_part = part;
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    const bool staggered = R.M.p.kind == OPESCI_KIND_STAGGERED_ELASTIC;
    out_ms[0] = out_ms[1] = out_ms[2] = 0.0;
    float ms;
    for (int phase = 0; phase < 3; ++phase) {
        if ((!staggered || part != 0) && phase > 0) break;
        CUDA_OK(cudaStreamSynchronize(st));
        CUDA_OK(cudaEventRecord(e0, st));
        for (int r = 0; r < reps; ++r) {
            const int ti = R.M.p.ntsteps + r, t0 = ti % 2, t1 = (t0 + 1) % 2;
            if (!staggered) S.template acoustic_step<SO, T, ARITH>(ti);
            else if (phase == 0) { if (R.fused) S.template fused<SO, T, ARITH>(t0, t1); else S.template stress<SO, T, ARITH>(t0, t1); }
            else if (phase == 1) { if (R.fused) break; S.template velocity<SO, T, ARITH>(t0, t1); }
            else {
                S.template stress_bc<T>(t0, t1, false);
                if (R.fused) S.template velocity_shell<SO, T, ARITH>(t0, t1);
                S.template velocity_bc<T>(t1);
            }
        }
        CUDA_OK(cudaEventRecord(e1, st));
        CUDA_OK(cudaStreamSynchronize(st));
        CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        out_ms[phase] = (R.fused && phase == 1) ? 0.0 : ms / reps;
    }
    if (S.err != cudaSuccess) return fail("time_kernels: launch failed: %s", cudaGetErrorString(S.err));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaStreamDestroy(st);
    return 0;
}
template <typename T, int ARITH> int time_kernels_so(Run &R, int reps, double *out_ms, int part = 0)
{
    if (part != 0) return R.M.p.so == 4 ? time_kernels_impl<4, T, ARITH>(R, reps, out_ms, part) : fail("fused parts: so = 4 only");
    switch (R.M.p.so) {
    case 2: return time_kernels_impl<2, T, ARITH>(R, reps, out_ms);
    case 4: return time_kernels_impl<4, T, ARITH>(R, reps, out_ms);
    case 6: return time_kernels_impl<6, T, ARITH>(R, reps, out_ms);
    case 8: return time_kernels_impl<8, T, ARITH>(R, reps, out_ms);
    case 10: return time_kernels_impl<10, T, ARITH>(R, reps, out_ms);
    case 12: return time_kernels_impl<12, T, ARITH>(R, reps, out_ms);
    }
    return fail("unsupported spatial order");
}

void release(Run *R)
{
    for (int f = 0; f < OPESCI_MAX_FIELDS; ++f) {
        if (R->dev[f]) cudaFree(R->dev[f]);
        if (R->host[f]) pool_release(R->host[f]);
    }
    for (int k = 0; k < OPESCI_MEDIA_COUNT; ++k)
        if (R->media[k]) cudaFree(R->media[k]);
    if (R->d_tables) cudaFree(R->d_tables);
    if (R->d_prog) cudaFree(R->d_prog);
    if (R->d_recv_cell) cudaFree(R->d_recv_cell);
    if (R->d_recv_out) cudaFree(R->d_recv_out);
    if (R->d_src) cudaFree(R->d_src);
    if (R->d_step) cudaFree(R->d_step);
    if (R->d_pace) cudaFree(R->d_pace);
    opesci_generic::unload(R->gen);
    if (R->ev_edge_fork) cudaEventDestroy(R->ev_edge_fork);
    if (R->ev_edge_join) cudaEventDestroy(R->ev_edge_join);
    if (R->st_edge) cudaStreamDestroy(R->st_edge);
    delete R;
}

// upload tables + programs
int upload_programs(Run &R)
{
    const Model &M = R.M;
    const size_t nt = M.tables.size();
    CUDA_OK(cudaMalloc(&R.d_tables, (nt ? nt : 1) * sizeof(double)));
    if (nt) CUDA_OK(cudaMemcpy(R.d_tables, M.tables.data(), nt * sizeof(double), cudaMemcpyHostToDevice));
    std::vector<DevProgram> progs(2 * (size_t)M.p.nfields);
    for (int f = 0; f < M.p.nfields; ++f)
        for (int w = 0; w < 2; ++w) {
            const OpesciSolProgram &src = w ? M.p.fields[f].final_ : M.p.fields[f].init;
            DevProgram &dst = progs[2 * f + w];
            memset(&dst, 0, sizeof dst);
            dst.n_instr = src.n_instr;
            dst.n_tables = src.n_tables;
            for (int t = 0; t < src.n_tables; ++t) {
                dst.table_axis[t] = src.table_axis[t];
                dst.table[t] = R.d_tables + M.table_off[f][w][t] + (src.table_axis[t] == 0 ? M.slab.L0 : 0);
            }
            for (int i = 0; i < src.n_instr; ++i) {
                dst.instr[i] = src.instr[i];
                if (src.instr[i].op == OPESCI_OP_MEDIA && (!M.p.hetero || !R.media[0]))
                    return fail("solution program reads media arrays: needs a heterogeneous grid produced by opesci_execute");
            }
            for (int k = 0; k < OPESCI_MEDIA_COUNT; ++k) dst.media[k] = R.media[k];
        }
    CUDA_OK(cudaMalloc(&R.d_prog, progs.size() * sizeof(DevProgram)));
    CUDA_OK(cudaMemcpy(R.d_prog, progs.data(), progs.size() * sizeof(DevProgram), cudaMemcpyHostToDevice));
    return 0;
}

// a11 (opesci/staggeredgrid.py:522-598): upload this rank's planes of rho, vp, vs and derive the nine
// media arrays on the device.  Arrays are zero-filled first: like the reference's, cells outside the loop
// ranges stay 0.
int setup_media(Run &R, cudaStream_t st)
{
    const Model &M = R.M;
    const OpesciB200Params &p = M.p;
    if (!p.hetero) return 0;
    if (!p.rho || !p.vp || !p.vs) return fail("heterogeneous media: rho/vp/vs missing");
    if (p.media_plane0 > M.slab.L0 || p.media_plane0 + p.media_nplanes < M.slab.L1)
        return fail("heterogeneous media: rho/vp/vs do not cover the planes this rank stores (opesci_b200_slab_range)");
    const size_t bytes = (size_t)M.G.level * sizeof(float);
    float *in[3] = {nullptr, nullptr, nullptr};
    const float *src[3] = {p.rho, p.vp, p.vs};
    const size_t host_row = (size_t)p.dim[2] * sizeof(float), dev_row = (size_t)M.G.s[1] * sizeof(float);
    const size_t rows = (size_t)M.G.dim[0] * p.dim[1];
    const size_t shift = (size_t)(M.slab.L0 - p.media_plane0) * p.dim[1] * p.dim[2];
    auto cleanup = [&]() { for (float *q : in) if (q) cudaFree(q); };
    for (int k = 0; k < 3; ++k) {
        if (cudaMalloc(&in[k], bytes) != cudaSuccess || cudaMemsetAsync(in[k], 0, bytes, st) != cudaSuccess ||
            cudaMemcpy2DAsync(in[k], dev_row, src[k] + shift, host_row, host_row, rows, cudaMemcpyHostToDevice, st) != cudaSuccess) {
            cleanup();
            return fail("heterogeneous media: upload failed (%s)", cudaGetErrorString(cudaGetLastError()));
        }
    }
    MediaOut O;
    for (int k = 0; k < OPESCI_MEDIA_COUNT; ++k) {
        if (cudaMalloc(&R.media[k], bytes) != cudaSuccess || cudaMemsetAsync(R.media[k], 0, bytes, st) != cudaSuccess) {
            cleanup();
            return fail("heterogeneous media: cudaMalloc failed (%s)", cudaGetErrorString(cudaGetLastError()));
        }
        O.m[k] = R.media[k];
    }
    dim3 blk(64, 4);
    dim3 grid((M.G.dim[2] + blk.x - 1) / blk.x, (M.G.dim[1] + blk.y - 1) / blk.y, M.G.dim[0]);
    media_pointwise<<<grid, blk, 0, st>>>(in[0], in[1], in[2], O, M.G);
    media_averaged<<<grid, blk, 0, st>>>(O, M.G);
    cudaError_t e = cudaStreamSynchronize(st);
    cleanup();
    if (e != cudaSuccess || (e = cudaGetLastError()) != cudaSuccess) return fail("heterogeneous media: kernels failed (%s)", cudaGetErrorString(e));
    return 0;
}

// point source + receivers: device copies of the cell offsets and the source time series
int setup_hooks(Run &R)
{
    const Model &M = R.M;
    const OpesciB200Params &p = M.p;
    R.hooks = p.n_receivers > 0 || p.src_nt > 0;
    if (!R.hooks) return 0;
    const size_t esz = p.is_double ? 8 : 4;
    auto offset = [&](const int32_t *c) -> long long {
        if (c[0] < M.slab.own_lo || c[0] >= M.slab.own_hi) return -1;   // another rank's plane
        return (long long)(c[0] - M.slab.L0) * M.G.s[0] + (long long)c[1] * M.G.s[1] + c[2];
    };
    CUDA_OK(cudaMalloc(&R.d_step, sizeof(int)));
    CUDA_OK(cudaMemset(R.d_step, 0, sizeof(int)));
    if (p.n_receivers > 0) {
        std::vector<long long> cells(p.n_receivers);
        for (int r = 0; r < p.n_receivers; ++r) cells[r] = offset(p.receiver_cells + 3 * r);
        CUDA_OK(cudaMalloc(&R.d_recv_cell, cells.size() * sizeof(long long)));
        CUDA_OK(cudaMemcpy(R.d_recv_cell, cells.data(), cells.size() * sizeof(long long), cudaMemcpyHostToDevice));
        const size_t bytes = (size_t)(p.ntsteps > 0 ? p.ntsteps : 1) * 4 * p.n_receivers * esz;
        CUDA_OK(cudaMalloc(&R.d_recv_out, bytes));
        CUDA_OK(cudaMemset(R.d_recv_out, 0, bytes));
    }
    R.src_cell = p.src_nt > 0 ? offset(p.source_cell) : -1;
    const size_t nsrc = p.src_nt > 0 ? p.src_nt : 1;
    CUDA_OK(cudaMalloc(&R.d_src, 3 * nsrc * sizeof(float)));
    if (p.src_nt > 0) {
        CUDA_OK(cudaMemcpy(R.d_src, p.src_x, p.src_nt * sizeof(float), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(R.d_src + p.src_nt, p.src_y, p.src_nt * sizeof(float), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(R.d_src + 2 * (size_t)p.src_nt, p.src_z, p.src_nt * sizeof(float), cudaMemcpyHostToDevice));
    }
    return 0;
}

Run *find_run(OpesciGrid *grid)
{
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_runs.find(grid->field[0]);
    return it == g_runs.end() ? nullptr : it->second;
}

template <typename T> int l2_sums(Run &R, const void *const *level_base_dev, double *sums)
{
    const Model &M = R.M;
    cudaStream_t st = 0;
    double *d_out = nullptr, *d_partial = nullptr;
    CUDA_OK(cudaMalloc(&d_out, OPESCI_MAX_FIELDS * sizeof(double)));
    CUDA_OK(cudaMemset(d_out, 0, OPESCI_MAX_FIELDS * sizeof(double)));
    size_t max_blocks = 1;
    dim3 blk(64, 4);
    std::vector<dim3> grids(M.p.nfields);
    std::vector<Range3> ranges(M.p.nfields);
    for (int f = 0; f < M.p.nfields; ++f) {
        const OpesciFieldSpec &fs = M.p.fields[f];
        Range3 &rg = ranges[f];
        for (int d = 0; d < 3; ++d) { rg.lo[d] = fs.l2_lo[d]; rg.hi[d] = fs.l2_hi[d]; }
        // every rank sums over the planes it OWNS (global range clipped), in local indices
        rg.lo[0] = (rg.lo[0] > M.slab.own_lo ? rg.lo[0] : M.slab.own_lo) - M.slab.L0;
        rg.hi[0] = (rg.hi[0] < M.slab.own_hi ? rg.hi[0] : M.slab.own_hi) - M.slab.L0;
        const int nz = rg.hi[2] - rg.lo[2], ny = rg.hi[1] - rg.lo[1], nx = rg.hi[0] - rg.lo[0];
        grids[f] = dim3(nz > 0 ? (nz + blk.x - 1) / blk.x : 0, ny > 0 ? (ny + blk.y - 1) / blk.y : 0, nx > 0 ? nx : 0);
        const size_t nb = (size_t)grids[f].x * grids[f].y * grids[f].z;
        if (nb > max_blocks) max_blocks = nb;
    }
    CUDA_OK(cudaMalloc(&d_partial, max_blocks * sizeof(double)));
    for (int f = 0; f < M.p.nfields; ++f) {
        const size_t nb = (size_t)grids[f].x * grids[f].y * grids[f].z;
        if (nb == 0) continue;
        l2_partial<T><<<grids[f], blk, 0, st>>>((const T *)level_base_dev[f], M.G, ranges[f], R.d_prog + 2 * f + 1, d_partial);
        l2_final<<<1, 1024, 0, st>>>(d_partial, nb, d_out + f);
    }
    if (M.slab.nranks > 1 && !g_nccl.comm) return fail("L2 norms of a slab need the NCCL communicator (loopback slabs: compare the fields instead)");
    if (M.slab.nranks > 1) {
        // sum of the per-slab partial sums (ncclAllReduce of 9 doubles)
        NCCL_OK(g_nccl.AllReduce(d_out, d_out, OPESCI_MAX_FIELDS, ncclFloat64, ncclSum, g_nccl.comm, st));
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpy(sums, d_out, OPESCI_MAX_FIELDS * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d_out);
    cudaFree(d_partial);
    return 0;
}

// OPESCI_L2_REFERENCE: accumulators of the reference's serial real_t sums (kernels.cuh: l2_terms / l2_serial).
// acc_out[f] holds the final accumulator of field f converted to double (exact).
template <typename T> int l2_reference(Run &R, const void *const *level_base_dev, double *acc_out)
{
    const Model &M = R.M;
    if (M.slab.nranks > 1) return fail("OPESCI_L2_REFERENCE: the serial reference-order sum needs the whole domain on one rank");
    cudaStream_t st = 0;
    const int nf = M.p.nfields;
    Range3 full[OPESCI_MAX_FIELDS];
    size_t plane_cells = 1;
    int xlo = 1 << 30, xhi = 0;
    for (int f = 0; f < nf; ++f) {
        const OpesciFieldSpec &fs = M.p.fields[f];
        for (int d = 0; d < 3; ++d) { full[f].lo[d] = fs.l2_lo[d]; full[f].hi[d] = fs.l2_hi[d]; }
        const size_t ny = full[f].hi[1] > full[f].lo[1] ? full[f].hi[1] - full[f].lo[1] : 0;
        const size_t nz = full[f].hi[2] > full[f].lo[2] ? full[f].hi[2] - full[f].lo[2] : 0;
        if (ny * nz > plane_cells) plane_cells = ny * nz;
        if (full[f].lo[0] < xlo) xlo = full[f].lo[0];
        if (full[f].hi[0] > xhi) xhi = full[f].hi[0];
    }
    // chunks of whole x planes, about 4 M cells per field and chunk
    int px = (int)((size_t)(4u << 20) / plane_cells);
    if (px < 1) px = 1;
    const size_t stride = (size_t)px * plane_cells;
    double *d_terms = nullptr;
    T *d_acc = nullptr;
    CUDA_OK(cudaMalloc(&d_terms, (size_t)nf * stride * sizeof(double)));
    if (cudaMalloc(&d_acc, OPESCI_MAX_FIELDS * sizeof(T)) != cudaSuccess) { cudaFree(d_terms); return fail("OPESCI_L2_REFERENCE: out of device memory"); }
    cudaMemsetAsync(d_acc, 0, OPESCI_MAX_FIELDS * sizeof(T), st);
    dim3 blk(64, 4);
    for (int x0 = xlo; x0 < xhi; x0 += px) {
        L2SerialArgs N;
        for (int f = 0; f < OPESCI_MAX_FIELDS; ++f) N.count[f] = 0;
        for (int f = 0; f < nf; ++f) {
            Range3 rg = full[f];
            rg.lo[0] = rg.lo[0] > x0 ? rg.lo[0] : x0;
            rg.hi[0] = rg.hi[0] < x0 + px ? rg.hi[0] : x0 + px;
            const int nz = rg.hi[2] - rg.lo[2], ny = rg.hi[1] - rg.lo[1], nx = rg.hi[0] - rg.lo[0];
            if (nz <= 0 || ny <= 0 || nx <= 0) continue;
            N.count[f] = (long long)nx * ny * nz;
            dim3 grid((nz + blk.x - 1) / blk.x, (ny + blk.y - 1) / blk.y, nx);
            l2_terms<T><<<grid, blk, 0, st>>>((const T *)level_base_dev[f], M.G, rg, R.d_prog + 2 * f + 1, d_terms + (size_t)f * stride);
        }
        l2_serial<T><<<nf, 256, 0, st>>>(d_terms, stride, N, d_acc);
    }
    T h_acc[OPESCI_MAX_FIELDS];
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpy(h_acc, d_acc, sizeof h_acc, cudaMemcpyDeviceToHost);
    cudaFree(d_terms);
    cudaFree(d_acc);
    if (e != cudaSuccess) return fail("OPESCI_L2_REFERENCE: %s", cudaGetErrorString(e));
    for (int f = 0; f < nf; ++f) acc_out[f] = (double)h_acc[f];
    return 0;
}

// L2 sums of the level `ntsteps % 2` (the reference's `ti`, also for tp == 3: regulargrid.py:664)
int convergence_sums(OpesciGrid *grid, double *sums, Model **model_out, bool reference_order = false)
{
    Run *R = find_run(grid);
    Run *tmp = nullptr;
    std::vector<void *> uploaded;
    const void *base[OPESCI_MAX_FIELDS] = {};
    if (!R) {
        // arrays this library did not produce: the reference semantics are "read the host arrays
        // in *grid" -- upload level ti and reduce on the device
        if (!g_model.configured) return fail("opesci_convergence: not configured");
        tmp = new Run();
        tmp->M = g_model;
        if (upload_programs(*tmp)) { release(tmp); return 1; }
        R = tmp;
    }
    const Model &M = R->M;
    const int ti = M.p.ntsteps % 2;
    const size_t esz = M.p.is_double ? 8 : 4;
    const size_t lvl_bytes = (size_t)M.G.level * esz;
    const size_t host_row = (size_t)M.p.dim[2] * esz, dev_row = (size_t)M.G.s[1] * esz;
    const size_t rows = (size_t)M.G.dim[0] * M.p.dim[1];
    const size_t host_lvl_bytes = rows * host_row;
    for (int f = 0; f < M.p.nfields; ++f) {
        if (!tmp) {
            base[f] = (const char *)R->dev[f] + (size_t)ti * lvl_bytes;
        } else {
            void *d = nullptr;
            if (cudaMalloc(&d, lvl_bytes) != cudaSuccess || cudaMemset(d, 0, lvl_bytes) != cudaSuccess ||
                cudaMemcpy2D(d, dev_row, (const char *)grid->field[f] + (size_t)ti * host_lvl_bytes, host_row, host_row, rows,
                             cudaMemcpyHostToDevice) != cudaSuccess) {
                for (void *u : uploaded) cudaFree(u);
                release(tmp);
                return fail("opesci_convergence: cannot upload host arrays (%s)", cudaGetErrorString(cudaGetLastError()));
            }
            uploaded.push_back(d);
            base[f] = d;
        }
    }
    int rc;
    if (reference_order) rc = M.p.is_double ? l2_reference<double>(*R, base, sums) : l2_reference<float>(*R, base, sums);
    else rc = M.p.is_double ? l2_sums<double>(*R, base, sums) : l2_sums<float>(*R, base, sums);
    for (void *u : uploaded) cudaFree(u);
    if (model_out) *model_out = &g_model;
    if (tmp) release(tmp);
    return rc;
}

// x-slab of rank `rank` of `nranks` and the geometry of the planes it stores (include/opesci_slab.h)
int apply_slab(Model &M, int rank, int nranks)
{
    const OpesciB200Params &p = M.p;
    const int need = opesci_slab_need(p.kind == OPESCI_KIND_REGULAR_ACOUSTIC, p.so);
    if (opesci_slab_make(&M.slab, rank, nranks, p.dim[0], M.m, OPESCI_SLAB_HALO, need)) return 1;
    for (int d = 0; d < 3; ++d) M.G.dim[d] = p.dim[d];
    M.G.dim[0] = M.slab.L1 - M.slab.L0;   // local planes; p.dim[0] stays the global dim1
    M.G.m = M.m;
    // device rows are padded to a multiple of 32 elements (TMA needs 16-B row strides; 128-B rows
    // keep tiles sector-aligned); the host arrays keep the reference's dense layout
    const long long pitch = ((long long)p.dim[2] + 31) / 32 * 32;
    M.G.s[0] = (long long)p.dim[1] * pitch;
    M.G.s[1] = pitch;
    M.G.s[2] = 1;
    M.G.level = (long long)M.G.dim[0] * p.dim[1] * pitch;
    return 0;
}

}  // namespace

// ==================================================================== C ABI
extern "C" {

const char *opesci_b200_last_error(void) { return g_err; }
int opesci_b200_is_cuda(void) { return 1; }
int opesci_b200_halo_transport(void) { return tl_halo_transport; }

int opesci_b200_configure(const OpesciB200Params *params)
{
    if (!params || params->struct_size != sizeof(OpesciB200Params))
        return fail("opesci_b200_configure: struct_size mismatch (header / binding out of date)");
    if (params->so < 2 || params->so > 12 || (params->so & 1)) return fail("opesci_b200_configure: so must be even, 2..12");
    for (int d = 0; d < 3; ++d)
        if (params->dim[d] < 2 * (params->so / 2) + 1) return fail("opesci_b200_configure: grid too small for the stencil");
    Model &M = g_model;
    M = Model();
    M.p = *params;
    M.m = params->so / 2;
    const int nranks = params->slab_nranks > 1 ? params->slab_nranks : 1;
    if (apply_slab(M, nranks > 1 ? params->slab_rank : 0, nranks))
        return fail("opesci_b200_configure: slabs thinner than the halo: use fewer ranks");
    if (nranks > 1 && (!g_nccl.comm || g_nccl.nranks != nranks || g_nccl.rank != params->slab_rank))
        return fail("opesci_b200_configure: slab_nranks > 1 needs opesci_b200_comm_init with the same rank / size first");
    memcpy(M.sc.sn, params->c_stress_normal, sizeof M.sc.sn);
    memcpy(M.sc.ss, params->c_stress_shear, sizeof M.sc.ss);
    memcpy(M.sc.v, params->c_velocity, sizeof M.sc.v);
    for (int d = 0; d < 3; ++d) {
        M.ac.present[d] = M.ac_init.present[d] = 0;
        for (int k = 0; k < OPESCI_MAX_M; ++k) {
            M.ac.c[d][k] = params->ac_coef[d][k];
            M.ac_init.c[d][k] = params->ac_init_coef[d][k];
            if (k < M.m && params->ac_coef[d][k] != 0.0f) M.ac.present[d] = 1;
            if (k < M.m && params->ac_init_coef[d][k] != 0.0f) M.ac_init.present[d] = 1;
        }
    }
    M.ac.centre = params->ac_centre;
    M.ac_init.centre = params->ac_init_centre;
    // private copy of the 1-D tables (the caller's arrays need not outlive this call)
    for (int f = 0; f < params->nfields; ++f)
        for (int w = 0; w < 2; ++w) {
            const OpesciSolProgram &pr = w ? params->fields[f].final_ : params->fields[f].init;
            if (pr.n_instr > OPESCI_MAX_PROG || pr.n_tables > OPESCI_MAX_TABLES) return fail("solution program too large");
            M.table_off[f][w].clear();
            for (int t = 0; t < pr.n_tables; ++t) {
                const int ax = pr.table_axis[t];
                if (ax < 0 || ax > 2 || !pr.table[t]) return fail("bad solution table");
                M.table_off[f][w].push_back(M.tables.size());
                M.tables.insert(M.tables.end(), pr.table[t], pr.table[t] + params->dim[ax]);
            }
        }
    for (int f = 0; f < OPESCI_MAX_FIELDS; ++f)
        for (int w = 0; w < 2; ++w) {
            OpesciSolProgram &pr = w ? M.p.fields[f].final_ : M.p.fields[f].init;
            for (int t = 0; t < OPESCI_MAX_TABLES; ++t) pr.table[t] = nullptr;   // host pointers are not kept
        }
    if ((params->n_receivers > 0 || params->src_nt > 0) && params->kind != OPESCI_KIND_STAGGERED_ELASTIC)
        return fail("point source / receivers: staggered elastic model only");
    if ((params->n_receivers > 0 && (!params->receiver_cells || !params->receiver_out)) ||
        (params->src_nt > 0 && (!params->src_x || !params->src_y || !params->src_z)))
        return fail("point source / receivers: null array");
    for (int r = 0; r < params->n_receivers; ++r)
        for (int d = 0; d < 3; ++d)
            if (params->receiver_cells[3 * r + d] < 0 || params->receiver_cells[3 * r + d] >= params->dim[d]) return fail("receiver cell outside the grid");
    if (params->src_nt > 0)
        for (int d = 0; d < 3; ++d)
            if (params->source_cell[d] < 0 || params->source_cell[d] >= params->dim[d]) return fail("source cell outside the grid");
    if (params->kind == OPESCI_KIND_STAGGERED_ELASTIC) {
        if (params->nfields != 9 || params->nlevels != 2) return fail("staggered: need 9 fields, 2 levels");
        if (params->hetero) {
            if (params->is_double) return fail("heterogeneous media: fp32 only (the reference reader is float*)");
            memcpy(M.hc.c, params->h_c, sizeof M.hc.c);
            memcpy(M.hc.c2, params->h_c2, sizeof M.hc.c2);
        }
        if (params->free_surface == 1) {
            if (params->so != 4) return fail("Levander free surface needs so == 4");
            if (params->hetero) build_levander_hetero(M); else build_levander(M);
        }
    } else if (params->kind == OPESCI_KIND_REGULAR_GENERIC) {
        if (params->nfields < 1 || params->nfields > OPESCI_MAX_FIELDS || params->nlevels != 3) return fail("generic PDEs: 1..9 fields, 3 time levels");
        if (!params->generic_source || !*params->generic_source) return fail("generic PDEs: generic_source is empty");
        if (nranks > 1) return fail("generic PDEs: x-slabs are not supported");
        M.generic_source = params->generic_source;
        M.p.generic_source = nullptr;    // the caller's string need not outlive this call
    } else if (params->kind == OPESCI_KIND_REGULAR_ACOUSTIC) {
        if (params->nfields != 1 || params->nlevels != 3) return fail("regular: need 1 field, 3 levels");
    } else {
        return fail("opesci_b200_configure: unknown kind");
    }
    M.configured = true;
    g_err[0] = 0;
    return 0;
}

static int execute_model(const Model &model, OpesciGrid *grid, OpesciProfiling *profiling)
{
    const auto wall0 = std::chrono::steady_clock::now();
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail("opesci_execute: no CUDA device (this library has no CPU fallback)");
    Run *R = new Run();
    R->M = model;
    if (tl_loop) tl_loop->runs[model.slab.rank] = R;   // peers read it after the first barrier of an exchange
    const OpesciB200Params &p = R->M.p;
    const size_t esz = p.is_double ? 8 : 4;
    R->bytes_per_field = (size_t)R->M.G.level * p.nlevels * esz;
    R->host_bytes_per_field = (size_t)R->M.G.dim[0] * p.dim[1] * p.dim[2] * p.nlevels * esz;
    cudaStream_t st;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) { release(R); return fail("cudaStreamCreate failed"); }
    auto bail = [&](int rc) { cudaStreamDestroy(st); release(R); return rc; };
    for (int f = 0; f < p.nfields; ++f) {
        if (cudaMalloc(&R->dev[f], R->bytes_per_field) != cudaSuccess)
            return bail(fail("opesci_execute: cudaMalloc of a field failed (%s)", cudaGetErrorString(cudaGetLastError())));
        // regulargrid.py:489-490 allocates without clearing; de-facto contract "buffers start as 0"
        if (cudaMemsetAsync(R->dev[f], 0, R->bytes_per_field, st) != cudaSuccess) return bail(fail("cudaMemset failed"));
    }
    if (setup_media(*R, st)) return bail(1);
    if (upload_programs(*R)) return bail(1);
    if (setup_fused(*R)) return bail(1);
    if (setup_tiled(*R)) return bail(1);
    if (setup_hooks(*R)) return bail(1);
    if (p.kind == OPESCI_KIND_REGULAR_GENERIC) {
        const bool fast = (p.flags & OPESCI_ARITH_MASK) == OPESCI_ARITH_FAST;
        std::string detail;
        if (const char *e = opesci_generic::compile(R->M.generic_source, fast, R->gen, detail)) {
            static thread_local std::string msg;
            msg = detail.substr(0, 700);
            return bail(fail("%s: %s", e, msg.c_str()));
        }
    }
    double secs = 0.0;
    if (dispatch(*R, st, &secs)) return bail(1);
    if (!tl_loop || model.slab.rank == 0) g_loop_seconds = secs;
    if (p.n_receivers > 0 && p.receiver_out && p.ntsteps > 0 &&
        cudaMemcpy(p.receiver_out, R->d_recv_out, (size_t)p.ntsteps * 4 * p.n_receivers * esz, cudaMemcpyDeviceToHost) != cudaSuccess)
        return bail(fail("opesci_execute: copying the receiver data back failed"));
    const int mirror = p.flags & OPESCI_HOST_MIRROR_MASK;
    if (mirror == OPESCI_HOST_MIRROR_FULL) {
        for (int f = 0; f < p.nfields; ++f) {
            // pinned host arrays from the pool: the D2H copy runs at PCIe speed
            bool pinned = false;
            R->host[f] = pool_alloc(R->host_bytes_per_field, &pinned);
            if (!R->host[f]) return bail(fail("opesci_execute: host allocation failed"));
            R->host_pinned = pinned;
            // dense host rows (reference layout [tp][dim1][dim2][dim3]) <- pitched device rows
            if (cudaMemcpy2DAsync(R->host[f], (size_t)p.dim[2] * esz, R->dev[f], (size_t)R->M.G.s[1] * esz, (size_t)p.dim[2] * esz,
                                  (size_t)p.nlevels * R->M.G.dim[0] * p.dim[1], cudaMemcpyDeviceToHost, st) != cudaSuccess)
                return bail(fail("opesci_execute: D2H copy failed"));
        }
        if (cudaStreamSynchronize(st) != cudaSuccess) return bail(fail("opesci_execute: D2H copy failed (%s)", cudaGetErrorString(cudaGetLastError())));
        for (int f = 0; f < p.nfields; ++f) grid->field[f] = R->host[f];   // regulargrid.py:445-453
    } else {
        for (int f = 0; f < p.nfields; ++f) grid->field[f] = R->dev[f];
    }
    cudaStreamDestroy(st);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_runs[grid->field[0]] = R;
    }
    if (profiling) {
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - wall0).count();
        const double pts = (double)(p.dim[0] - 2 * R->M.m) * (p.dim[1] - 2 * R->M.m) * (p.dim[2] - 2 * R->M.m);
        const double flop_per_pt = p.kind == OPESCI_KIND_STAGGERED_ELASTIC ? 48.0 * p.so : 4.0 * p.so + 3.0;
        profiling->g_rtime = (float)secs;
        profiling->g_ptime = (float)wall;
        profiling->g_mflops = secs > 0 ? (float)(pts * p.ntsteps * flop_per_pt / secs * 1e-6) : 0.f;
    }
    return 0;
}

int opesci_execute(OpesciGrid *grid, OpesciProfiling *profiling)
{
    if (!g_model.configured) return fail("opesci_execute: opesci_b200_configure was not called");
    return execute_model(g_model, grid, profiling);
}

int opesci_b200_execute_loopback(int nranks, OpesciGrid *grids)
{
    if (!g_model.configured) return fail("opesci_b200_execute_loopback: opesci_b200_configure was not called");
    if (g_model.slab.nranks > 1) return fail("opesci_b200_execute_loopback: configure the whole domain (slab_nranks <= 1)");
    if (nranks < 2 || !grids) return fail("opesci_b200_execute_loopback: need nranks >= 2 and one OpesciGrid per rank");
    if (g_model.p.n_receivers > 0 || g_model.p.src_nt > 0) return fail("opesci_b200_execute_loopback: point source / receivers are per process");
    if (opesci_io::output_cfg().armed) return fail("opesci_b200_execute_loopback: per-step output is per process");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail("opesci_b200_execute_loopback: no CUDA device");
    Loopback L;
    L.nranks = nranks;
    L.runs.assign(nranks, nullptr);
    L.done.assign(nranks, nullptr);
    L.xdone.assign(nranks, nullptr);
    L.barrier.n = nranks;
    std::vector<Model> models(nranks, g_model);
    for (int r = 0; r < nranks; ++r) {
        models[r].p.slab_rank = r;
        models[r].p.slab_nranks = nranks;
        if (apply_slab(models[r], r, nranks)) return fail("opesci_b200_execute_loopback: slabs thinner than the halo: use fewer ranks");
        if (cudaEventCreateWithFlags(&L.done[r], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&L.xdone[r], cudaEventDisableTiming) != cudaSuccess)
            return fail("opesci_b200_execute_loopback: cudaEventCreate failed");
    }
    std::vector<int> rc(nranks, 0);
    std::vector<std::string> errs(nranks);
    std::mutex err_mu;
    std::vector<std::thread> threads;
    for (int r = 0; r < nranks; ++r)
        threads.emplace_back([&, r] {
            cudaSetDevice(dev);
            tl_loop = &L;
            rc[r] = execute_model(models[r], &grids[r], nullptr);
            if (rc[r]) {
                { std::lock_guard<std::mutex> lk(err_mu); errs[r] = g_err; }
                L.barrier.abort();     // peers waiting in an exchange give up instead of waiting for ever
            }
            tl_loop = nullptr;
        });
    for (auto &t : threads) t.join();
    cudaDeviceSynchronize();
    for (int r = 0; r < nranks; ++r) { cudaEventDestroy(L.done[r]); cudaEventDestroy(L.xdone[r]); }
    int bad = -1;
    for (int r = 0; r < nranks; ++r)
        if (rc[r] && (bad < 0 || errs[r].find("peer rank failed") == std::string::npos)) bad = r;
    if (bad >= 0) {
        for (int r = 0; r < nranks; ++r)
            if (!rc[r]) opesci_free(&grids[r]);
        return fail("loopback rank failed: %s", errs[bad].c_str());
    }
    return 0;
}

int opesci_b200_convergence_f64(OpesciGrid *grid, double *out_l2)
{
    double sums[OPESCI_MAX_FIELDS];
    Run *R = find_run(grid);
    const Model &M = R ? R->M : g_model;
    if (convergence_sums(grid, sums, nullptr)) return 1;
    for (int f = 0; f < M.p.nfields; ++f) out_l2[f] = sqrt(sums[f] * (double)(float)M.p.volume_literal);
    return 0;
}

int opesci_convergence(OpesciGrid *grid, OpesciConvergence *conv)
{
    // staggeredgrid.py:892-945: conv->F_l2 = pow(F_l2 * volume_literal, 0.5) in real_t.
    // Default: the sum is accumulated in double by a deterministic tree (the reference accumulates serially in real_t,
    // which loses digits on large grids: SURVEY.md 7 "hard parts").  With OPESCI_L2_REFERENCE the reference's own serial
    // real_t accumulation is performed instead, so the printed digits are the reference's.
    double sums[OPESCI_MAX_FIELDS];
    Run *R = find_run(grid);
    const Model &M = R ? R->M : g_model;
    const bool ref_order = (M.p.flags & OPESCI_L2_REFERENCE) != 0;
    if (convergence_sums(grid, sums, nullptr, ref_order)) return 1;
    for (int f = 0; f < M.p.nfields; ++f) {
        if (ref_order) {
            // `F_l2 = pow(F_l2 * volume, 0.5)` as emitted: real_t product, libm pow in double, result stored as real_t
            if (M.p.is_double) conv->f64[f] = pow(sums[f] * (double)(float)M.p.volume_literal, 0.5);
            else conv->f32[f] = (float)pow((double)((float)sums[f] * (float)M.p.volume_literal), 0.5);
        } else if (M.p.is_double) conv->f64[f] = sqrt(sums[f] * (double)(float)M.p.volume_literal);
        else conv->f32[f] = (float)sqrt((double)((float)sums[f] * (float)M.p.volume_literal));
    }
    return 0;
}

int opesci_free(OpesciGrid *grid)
{
    Run *R = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_runs.find(grid->field[0]);
        if (it != g_runs.end()) { R = it->second; g_runs.erase(it); }
    }
    if (!R) return fail("opesci_free: arrays were not allocated by opesci_execute");
    const int n = R->M.p.nfields;
    release(R);
    for (int f = 0; f < n; ++f) grid->field[f] = nullptr;
    return 0;
}

int opesci_b200_time_kernels(OpesciGrid *grid, int reps, double *out_ms)
{
    Run *R = find_run(grid);
    if (!R) return fail("opesci_b200_time_kernels: unknown grid");
    if (reps < 1) reps = 1;
    const bool fast = (R->M.p.flags & OPESCI_ARITH_MASK) == OPESCI_ARITH_FAST;
    if (R->M.p.is_double)
        return fast ? time_kernels_so<double, OPESCI_ARITH_FAST>(*R, reps, out_ms) : time_kernels_so<double, OPESCI_ARITH_REFERENCE>(*R, reps, out_ms);
    return fast ? time_kernels_so<float, OPESCI_ARITH_FAST>(*R, reps, out_ms) : time_kernels_so<float, OPESCI_ARITH_REFERENCE>(*R, reps, out_ms);
}

int opesci_b200_time_fused_parts(OpesciGrid *grid, int reps, double *out)
{
    Run *R = find_run(grid);
    if (!R) return fail("opesci_b200_time_fused_parts: unknown grid");
    out[0] = out[1] = out[2] = out[3] = 0.0;
    if (!R->fused || !R->zfold || R->M.p.is_double) return 0;     // one launch: opesci_b200_time_kernels covers it
    if (reps < 1) reps = 1;
    const bool fast = (R->M.p.flags & OPESCI_ARITH_MASK) == OPESCI_ARITH_FAST;
    double ms[3];
    for (int part = 1; part <= 2; ++part) {
        if (fast ? time_kernels_so<float, OPESCI_ARITH_FAST>(*R, reps, ms, part) : time_kernels_so<float, OPESCI_ARITH_REFERENCE>(*R, reps, ms, part)) return 1;
        out[part - 1] = ms[0];
    }
    // interior z columns stored by the interior launch (tile columns 1 .. nzt-2) and by both launches together
    const int m = R->M.m, CZ = FusedCfg<2>::CZ;
    out[2] = (double)(R->zf_nzt > 2 ? (R->zf_nzt - 2) * CZ : 0);
    out[3] = (double)(R->M.p.dim[2] - 2 * m);
    return 0;
}

int opesci_b200_comm_unique_id(void *out_id, int nbytes)
{
    if (nbytes < (int)sizeof(ncclUniqueId)) return fail("opesci_b200_comm_unique_id: buffer too small");
    if (nccl_bind()) return 1;
    ncclUniqueId id;
    NCCL_OK(g_nccl.GetUniqueId(&id));
    memcpy(out_id, &id, sizeof id);
    return 0;
}

int opesci_b200_comm_init(int rank, int nranks, const void *id_bytes, int nbytes)
{
    if (nbytes < (int)sizeof(ncclUniqueId) || nranks < 1 || rank < 0 || rank >= nranks) return fail("opesci_b200_comm_init: bad arguments");
    if (nccl_bind()) return 1;
    if (g_nccl.comm) { g_nccl.CommDestroy(g_nccl.comm); g_nccl.comm = nullptr; }
    ncclUniqueId id;
    memcpy(&id, id_bytes, sizeof id);
    NCCL_OK(g_nccl.CommInitRank(&g_nccl.comm, nranks, id, rank));
    g_nccl.rank = rank;
    g_nccl.nranks = nranks;
    return 0;
}

int opesci_b200_slab_range(int rank, int nranks, int gdim1, int so, int *L0, int *L1)
{
    OpesciSlab sl;
    if (opesci_slab_make(&sl, rank, nranks, gdim1, so / 2, OPESCI_SLAB_HALO, opesci_slab_need(0, so))) return fail("opesci_b200_slab_range: slabs thinner than the halo");
    if (L0) *L0 = sl.L0;
    if (L1) *L1 = sl.L1;
    return 0;
}

int opesci_b200_comm_finalize(void)
{
    if (g_nccl.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g_nccl.comm);
    g_nccl.comm = nullptr;
    g_nccl.rank = 0;
    g_nccl.nranks = 1;
    return 0;
}

int opesci_b200_reserve_host(size_t bytes_per_array, int count)
{
    std::vector<void *> got;
    g_pool_reserving = true;     // blocks created from here on survive opesci_free (single-threaded caller, like the whole ABI)
    for (int i = 0; i < count; ++i) {
        bool pinned = false;
        void *p = pool_alloc(bytes_per_array, &pinned);
        if (!p) { g_pool_reserving = false; for (void *q : got) pool_release(q); return fail("opesci_b200_reserve_host: allocation failed"); }
        got.push_back(p);
    }
    g_pool_reserving = false;
    {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        for (auto &b : g_pool)
            for (void *q : got)
                if (b.ptr == q) b.reserved = true;    // an idle on-demand block that was reused counts as reserved too
    }
    for (void *q : got) pool_release(q);
    return 0;
}

int opesci_b200_release_host(void)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    std::vector<HostBlock> keep;
    for (auto &b : g_pool) {
        if (b.in_use) { keep.push_back(b); continue; }
        if (b.pinned) cudaFreeHost(b.ptr);
        else free(b.ptr);
    }
    g_pool.swap(keep);
    return 0;
}

int opesci_b200_last_timing(double *loop_seconds, double *points_per_step, int64_t *kernel_launches)
{
    const Model &M = g_model;
    if (loop_seconds) *loop_seconds = g_loop_seconds;
    if (points_per_step) *points_per_step = (double)(M.p.dim[0] - 2 * M.m) * (M.p.dim[1] - 2 * M.m) * (M.p.dim[2] - 2 * M.m);
    if (kernel_launches) *kernel_launches = g_launches;
    return 0;
}

}  // extern "C"
