// generic.cuh -- run-time compiled kernels for PDE systems outside the fixed-function kernels.
//
// The reference accepts any PDE set: RegularGrid.solve_fd substitutes the finite-difference expressions, solves every
// equation for the newest time level with sympy and pastes the printed expression into the generated C++
// (opesci/regulargrid.py:230-270, 329-342, 530-619).  The B200 front end does the same derivation
// (opesci_fd_b200/regulargrid.py:_solve_fd_generic) and prints the same expressions into CUDA source -- one thread per
// grid point, the emitted expression verbatim -- which is compiled HERE for sm_100a with NVRTC and launched through the
// driver API.  Reference arithmetic = `--fmad=false` (one rounded multiply and one rounded add per emitted term, what gcc
// produces for the generated C++ on x86-64), fast arithmetic = NVRTC's default contraction.
//
// NVRTC and the driver entry points are bound at run time (dlopen / cudaGetDriverEntryPoint): the library does not link
// against libnvrtc or libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <string>
#include <vector>

namespace opesci_generic {

typedef struct _nvrtcProgram *nvrtcProgram;
struct Nvrtc {
    void *handle = nullptr;
    int (*CreateProgram)(nvrtcProgram *, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*CompileProgram)(nvrtcProgram, int, const char *const *) = nullptr;
    int (*GetCUBINSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetCUBIN)(nvrtcProgram, char *) = nullptr;
    int (*GetProgramLogSize)(nvrtcProgram, size_t *) = nullptr;
    int (*GetProgramLog)(nvrtcProgram, char *) = nullptr;
    int (*DestroyProgram)(nvrtcProgram *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
struct Driver {
    CUresult (*ModuleLoadData)(CUmodule *, const void *) = nullptr;
    CUresult (*ModuleGetFunction)(CUfunction *, CUmodule, const char *) = nullptr;
    CUresult (*ModuleUnload)(CUmodule) = nullptr;
    CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void **, void **) = nullptr;
};

inline const char *bind_nvrtc(Nvrtc &N)
{
    if (N.handle) return nullptr;
    const char *names[] = {getenv("OPESCI_NVRTC_LIB"), "libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"};
    void *h = nullptr;
    for (const char *n : names)
        if (n && *n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!h) return "NVRTC not found: set OPESCI_NVRTC_LIB or put libnvrtc.so.12 on the library path";
#define OPESCI_BIND(field, sym) *(void **)(&N.field) = dlsym(h, sym); if (!N.field) return "NVRTC symbol missing: " sym
    OPESCI_BIND(CreateProgram, "nvrtcCreateProgram");
    OPESCI_BIND(CompileProgram, "nvrtcCompileProgram");
    OPESCI_BIND(GetCUBINSize, "nvrtcGetCUBINSize");
    OPESCI_BIND(GetCUBIN, "nvrtcGetCUBIN");
    OPESCI_BIND(GetProgramLogSize, "nvrtcGetProgramLogSize");
    OPESCI_BIND(GetProgramLog, "nvrtcGetProgramLog");
    OPESCI_BIND(DestroyProgram, "nvrtcDestroyProgram");
    OPESCI_BIND(GetErrorString, "nvrtcGetErrorString");
#undef OPESCI_BIND
    N.handle = h;
    return nullptr;
}

inline const char *bind_driver(Driver &D)
{
    if (D.LaunchKernel) return nullptr;
    cudaDriverEntryPointQueryResult q;
    void *fn = nullptr;
#define OPESCI_ENTRY(field, sym) \
    if (cudaGetDriverEntryPoint(sym, &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return "driver entry point missing: " sym; \
    *(void **)(&D.field) = fn
    OPESCI_ENTRY(ModuleLoadData, "cuModuleLoadData");
    OPESCI_ENTRY(ModuleGetFunction, "cuModuleGetFunction");
    OPESCI_ENTRY(ModuleUnload, "cuModuleUnload");
    OPESCI_ENTRY(LaunchKernel, "cuLaunchKernel");
#undef OPESCI_ENTRY
    return nullptr;
}

// one compiled model: the time-step kernel and the second-initialisation kernel
struct Module {
    CUmodule mod = nullptr;
    CUfunction step = nullptr, init2 = nullptr;
    std::string log;
};

inline Nvrtc &nvrtc() { static Nvrtc n; return n; }
inline Driver &driver() { static Driver d; return d; }

// returns nullptr on success, else a message (err receives details: the compile log)
inline const char *compile(const std::string &source, bool fmad, Module &out, std::string &err)
{
    if (const char *e = bind_nvrtc(nvrtc())) return e;
    if (const char *e = bind_driver(driver())) return e;
    Nvrtc &N = nvrtc();
    nvrtcProgram prog = nullptr;
    int rc = N.CreateProgram(&prog, source.c_str(), "opesci_generic.cu", 0, nullptr, nullptr);
    if (rc != 0) { err = N.GetErrorString(rc); return "nvrtcCreateProgram failed"; }
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", fmad ? "--fmad=true" : "--fmad=false", "-lineinfo"};
    rc = N.CompileProgram(prog, 4, opts);
    size_t ls = 0;
    if (N.GetProgramLogSize(prog, &ls) == 0 && ls > 1) {
        out.log.resize(ls);
        N.GetProgramLog(prog, &out.log[0]);
    }
    if (rc != 0) { err = out.log.empty() ? std::string(N.GetErrorString(rc)) : out.log; N.DestroyProgram(&prog); return "NVRTC compilation of the generated kernel failed"; }
    size_t cs = 0;
    if (N.GetCUBINSize(prog, &cs) != 0 || cs == 0) { N.DestroyProgram(&prog); return "nvrtcGetCUBINSize failed"; }
    std::vector<char> cubin(cs);
    if (N.GetCUBIN(prog, cubin.data()) != 0) { N.DestroyProgram(&prog); return "nvrtcGetCUBIN failed"; }
    N.DestroyProgram(&prog);
    Driver &D = driver();
    cudaFree(nullptr);    // make sure the primary context is current on this thread
    if (D.ModuleLoadData(&out.mod, cubin.data()) != CUDA_SUCCESS) return "cuModuleLoadData failed";
    if (D.ModuleGetFunction(&out.step, out.mod, "opesci_generic_step") != CUDA_SUCCESS) return "generated module has no opesci_generic_step";
    if (D.ModuleGetFunction(&out.init2, out.mod, "opesci_generic_init2") != CUDA_SUCCESS) return "generated module has no opesci_generic_init2";
    return nullptr;
}

inline void unload(Module &m)
{
    if (m.mod && driver().ModuleUnload) driver().ModuleUnload(m.mod);
    m.mod = nullptr; m.step = m.init2 = nullptr;
}

// Launch one of the two kernels: arguments = nfields base pointers, then the time-level indices.
// Geometry: one thread per interior point, blocks of 64 x 4 along z, y; blockIdx.z = plane.
inline bool launch(CUfunction f, void *const *fields, int nfields, const int *levels, int nlevels, int nz, int ny, int nx, cudaStream_t st)
{
    void *ptrs[16];
    void *args[24];
    int lv[4];
    int n = 0;
    for (int k = 0; k < nfields; ++k) { ptrs[k] = fields[k]; args[n++] = &ptrs[k]; }
    for (int k = 0; k < nlevels; ++k) { lv[k] = levels[k]; args[n++] = &lv[k]; }
    if (nz <= 0 || ny <= 0 || nx <= 0) return true;
    return driver().LaunchKernel(f, (unsigned)((nz + 63) / 64), (unsigned)((ny + 3) / 4), (unsigned)nx, 64, 4, 1, 0, (CUstream)st, args, nullptr) == CUDA_SUCCESS;
}

}  // namespace opesci_generic
