/*
 * opesci_io.h -- C ABI of the model-input / field-output helpers around the time-stepping path
 * (the callers and data formats either side of the hot path).
 *
 * The reference keeps these in its C++ support library libopesci (include/opesciIO.h,
 * include/opesciHandy.h; std::string / std::vector signatures).  Each entry point below is the
 * plain-C mirror of one of them -- pointers and sizes only -- with the reference function it
 * replaces cited next to it.  They live in the same shared object as the CUDA kernels
 * (libopesci_b200.so).  Where a function has a device side (per-step snapshots, SEG-Y decoding into
 * device-resident media planes) it says so; the rest is host file I/O, as in the reference.
 */
#ifndef OPESCI_IO_H
#define OPESCI_IO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- per-step field output --------------------------------------------------
 * Replaces: the `output_step` block the generator emits under the `output_vts` switch
 * (opesci/regulargrid.py:702-719, opesci/staggeredgrid.py:882-890,
 * opesci/templates/staggered3d_tmpl.py:54-56): at the end of every time step `_ti`
 *     opesci_dump_field_vts_3d("<label>_" + std::to_string(_ti), dims, spacing, 2, &F[t1][0][0][0]);
 * for the first field (U / the regular grid's field), inside an `omp single` that stalls the time loop.
 * Here: when armed, opesci_execute copies the new time level of `field` to a page-locked staging
 * buffer on a second stream every `every` steps (the time loop only waits before it overwrites that
 * level again, two steps later) and a host callback on that stream writes
 * "<prefix><ti>.vts" with opesci_b200_dump_field_vts_3d (the Points array, identical in every file of a series
 * and three times the size of the field, is compressed once).  fp64 fields are written as Float32
 * (the reference's writer takes float*).  With x-slabs every rank writes the planes it owns to
 * "<prefix><ti>_r<rank>.vts".  every <= 0 or prefix == NULL disarms.  Applies to the next
 * opesci_execute calls of this process. */
int opesci_b200_set_output(const char *prefix, int field, int every);
/* zlib level (0-9) of every .vts written by this library.  Default 1: the reference asks VTK for level 9
 * (src/opesciIO.cpp:653), which decompresses to the same bytes but costs ~10x the time -- it decides whether the
 * output hides behind the time loop.  Returns -1 for a level outside 0-9. */
int opesci_b200_set_output_level(int zlib_level);
/* number of snapshot files written by the last opesci_execute and how many writes failed */
int opesci_b200_output_stats(int *files_written, int *write_errors);

/* Replaces: opesci_dump_field_vts_3d (src/opesciIO.cpp:614-667, include/opesciIO.h).
 * Writes `name`.vts: a VTK XML StructuredGrid, point (i,j,k) at ((i-margin)*spacing[0],
 * (j-margin)*spacing[1], (k-margin)*spacing[2]) in the reference's point order (k fastest), one
 * Float32 point-data array named "field", zlib-compressed like the reference's
 * vtkZLibDataCompressor output (appended raw data, UInt32 block headers; level: opesci_b200_set_output_level).  `x0` shifts the first index
 * (a slab's first owned plane; 0 otherwise).  Returns 0, -1 on I/O failure. */
int opesci_b200_dump_field_vts_3d(const char *name, const int dims[3], const float spacing[3], int margin,
                                  const float *field, int x0);

/* Replaces: opesci_dump_field_vts (src/opesciIO.cpp:143-194), the writer of src/segy2vts.cpp: field[i + j*dims[0] +
 * k*dims[0]*dims[1]] (x fastest, the layout-0 vectors of the SEG-Y reader), point (i,j,k) at (i*spacing[0],
 * j*spacing[1], k*spacing[2]).  `python -m opesci_fd_b200.segy2vts model.segy` is the converter built on it. */
int opesci_b200_dump_field_vts(const char *name, const int dims[3], const float spacing[3], const float *field);

/* ---- model input ------------------------------------------------------------ */
/* Replaces: opesci_read_simple_binary_ptr (src/opesciIO.cpp:319-342): flat float32 file into
 * array[0..size).  Returns 0; -1 if the file cannot be opened; -2 if it holds fewer than `size`
 * floats (the reference reads past its buffer in that case).  A longer file is truncated with a
 * warning on stderr, like the reference. */
int opesci_b200_read_simple_binary_ptr(const char *filename, float *array, size_t size);
/* number of float32 values in a flat binary file (opesci_read_simple_binary sizes its vector this
 * way, src/opesciIO.cpp:296-317); -1 if it cannot be opened */
int64_t opesci_b200_simple_binary_count(const char *filename);

/* Replaces: opesci_read_model_segy (src/opesciIO.cpp:451-612).  SEG-Y rev 1 model volume, one trace
 * per (ix,iy) column, Nx = traces per record (bytes 3213-3214), Nz = samples per trace (3221-3222),
 * Ny = ntraces/Nx, format code 1 (4-byte IBM float) only, byte order detected from the format code
 * like the reference; spacing[0] = spacing[1] = distance between the first two traces (scaled by the
 * coordinate scalar, bytes 71-72 of the trace header), spacing[2] = scalar * (bytes 117-118).
 * Two-call pattern: with array == NULL only dim[] / spacing[] are filled.
 * layout 0: array[ix + iy*Nx + iz*Nx*Ny]  (the reference's vector layout)
 * layout 1: array[(ix*Ny + iy)*Nz + iz]   (the [x][y][z] layout the time-stepping library's
 *           rho / vp / vs inputs use, include/opesci_b200.h)
 * Returns 0; -1 open failure / unsupported format; -2 capacity (floats) too small. */
int opesci_b200_read_model_segy(const char *filename, float *array, size_t capacity, int dim[3], float spacing[3],
                                int layout);
/* Device side of the same reader: `traces` are the raw SEG-Y trace records already in DEVICE memory
 * (ntraces records of 240 + 4*nz bytes, i.e. the file from byte 3600 on); a CUDA kernel decodes the IBM
 * floats and scatters them into the DEVICE array `out` in layout 0 or 1 -- the model never takes a
 * decoded round trip through host memory.  `stream` is a cudaStream_t (NULL: default stream).
 * Bit-identical to the host reader. */
int opesci_b200_segy_decode_device(const void *traces, int ntraces, int nx, int nz, int swap_endian, float *out,
                                   int layout, void *stream);
/* one IBM REAL*4 (4 bytes in file order after the optional byte swap) -> IEEE float, exactly as
 * real2float (src/opesciIO.cpp:400-417) computes it: sign * mantissa/2^24 * 16^(exponent-64) in double */
float opesci_b200_ibm_to_float(const unsigned char bytes[4], int swap_endian);

/* ---- sources and receivers -------------------------------------------------- */
/* Replaces: opesci_read_receivers (src/opesciIO.cpp:374-396) and the coordinate part of
 * opesci_read_souces (src/opesciIO.cpp:345-372): text file, first line is a header, then one
 * "x y z" triple per non-empty line.  Two-call pattern: xyz == NULL returns the number of triples;
 * otherwise fills xyz[3*n] (capacity in triples) and returns n.  -1 open failure, -2 capacity. */
int opesci_b200_read_xyz(const char *filename, float *xyz, int capacity);
/* Replaces: opesci_resample_timeseries (src/opesciHandy.cpp:125-195): DFT, zero-pad or cut the middle
 * of the spectrum, inverse DFT, both normalised by 1/sqrt(n) -- float accumulation, float angle
 * `2*M_PI*t*k/n` rounded to float, term order as in opesci_dft (src/opesciHandy.cpp:100-114).
 * n2 = round(n*sdt/dt) output samples (returned); out == NULL only returns n2; -2 capacity.
 * |dt-sdt| < eps*(dt+sdt) copies the input. */
int opesci_b200_resample_timeseries(const float *src, int n, float dt, double sdt, float *out, int capacity);
/* Replaces: opesci_calculate_dt (src/opesciHandy.cpp:69-91): (6/7)/sqrt(3) * h / max(vp) */
float opesci_b200_calculate_dt(const float *vp, size_t n, float h);
/* Replaces: opesci_calculate_lame_costants (src/opesciHandy.cpp:53-67): mu = rho*vs*vs,
 * lam = rho*(vp*vp - 2.0*vs*vs) (the `2.0*` term in double, as written there) */
void opesci_b200_calculate_lame_constants(const float *vp, const float *vs, const float *rho, size_t n, float *mu,
                                          float *lam);

#ifdef __cplusplus
}
#endif
#endif /* OPESCI_IO_H */
