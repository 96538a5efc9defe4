// TEST INFRASTRUCTURE ONLY (oracle/_ref build).
//
// Wrapper around ONE file of the reference's own generated C++ (emitted by the unmodified
// numerical logic of /root/reference/opesci via oracle/refgen/make_ref.py).  The generated
// translation unit is included verbatim; its `main` (templates/regular3d_tmpl.py:120-134)
// is renamed so that this driver can additionally dump the raw field arrays and time
// opesci_execute.  Nothing numerical is added here.
//
//   prog                      -> same output as the generated main()
//   prog --dump FILE NELEM    -> also fwrite()s every field (NELEM elements each, all time
//                                levels, struct order) to FILE before opesci_free
//   prog --time               -> prints "EXECUTE_SECONDS <s>" (wall time of opesci_execute)
#define main opesci_generated_main
#include OPESCI_GENERATED
#undef main

#include <chrono>
#include <cstring>

#ifndef OPESCI_REAL_T
#define OPESCI_REAL_T float
#endif

int main(int argc, char **argv)
{
    const char *dump = nullptr;
    long nelem = 0;
    bool timeit = false;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--dump") && i + 2 < argc) {
            dump = argv[i + 1];
            nelem = atol(argv[i + 2]);
            i += 2;
        } else if (!strcmp(argv[i], "--time")) {
            timeit = true;
        }
    }
    OpesciGrid grid;
    OpesciConvergence conv;
    OpesciProfiling profiling;
    memset(&conv, 0, sizeof(conv));
    auto t0 = std::chrono::steady_clock::now();
    opesci_execute(&grid, &profiling);
    auto t1 = std::chrono::steady_clock::now();
    opesci_convergence(&grid, &conv);
    const int nfields = (int)(sizeof(OpesciGrid) / sizeof(void *));
    if (dump) {
        FILE *f = fopen(dump, "wb");
        if (!f) { perror(dump); return 2; }
        OPESCI_REAL_T **ptrs = (OPESCI_REAL_T **)&grid;
        for (int k = 0; k < nfields; ++k)
            if (fwrite(ptrs[k], sizeof(OPESCI_REAL_T), (size_t)nelem, f) != (size_t)nelem) return 3;
        fclose(f);
    }
    opesci_free(&grid);
    OPESCI_REAL_T *norms = (OPESCI_REAL_T *)&conv;
    for (int k = 0; k < nfields; ++k)
        printf("L2[%d] %.10f %.9e\n", k, (double)norms[k], (double)norms[k]);
    if (timeit)
        printf("EXECUTE_SECONDS %.6f\n", std::chrono::duration<double>(t1 - t0).count());
    return 0;
}
