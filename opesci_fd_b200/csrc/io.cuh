// io.cuh -- model input / field output around the time-stepping path (include/opesci_io.h).
//
// Host side: plain-C mirrors of the reference's libopesci helpers (src/opesciIO.cpp, src/opesciHandy.cpp),
// restated from their behaviour.  Device side: (1) the per-step snapshot pipeline that replaces the reference's
// blocking `omp single { opesci_dump_field_vts_3d(...) }` (opesci/regulargrid.py:702-719) with
// compute | D2H copy | compress+write running concurrently, (2) the SEG-Y IBM-float decode/scatter kernel.
// Included by opesci_b200.cu (one translation unit).
#pragma once
#include <cuda_runtime.h>
#include <zlib.h>

#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/opesci_io.h"

namespace opesci_io {

// ------------------------------------------------------------------ VTK XML StructuredGrid writer
// One appended-data array: zlib blocks with the UInt32 header [nblocks, blocksize, lastblocksize, csize...]
// (the layout vtkXMLWriter emits for compressor="vtkZLibDataCompressor", header_type UInt32).
// `fill(dst, first_byte, nbytes)` produces the raw bytes of [first_byte, first_byte + nbytes).
// zlib level of the .vts writers.  The reference asks VTK for level 9 (src/opesciIO.cpp:653); the decompressed
// bytes are the same at any level, and level 9 costs ~10x the time of level 1 on field data, which is what decides
// whether the per-step output can hide behind the time loop.  opesci_b200_set_output_level() changes it.
inline int &zlib_level() { static int level = 1; return level; }

// an appended array already in its on-disk form (header + compressed blocks): the Points array of a snapshot series
// is the same in every file and three times the size of the field, so it is compressed once
struct PackedArray {
    std::vector<unsigned char> bytes;
    int dims[3] = {0, 0, 0}, margin = 0, x0 = 0, level = -1;
    float spacing[3] = {0, 0, 0};
};

template <typename Fill> bool append_compressed(FILE *fp, size_t total_bytes, Fill fill, size_t *written, PackedArray *keep = nullptr)
{
    const size_t BS = (size_t)1 << 20;
    const size_t nblocks = total_bytes ? (total_bytes + BS - 1) / BS : 0;
    const size_t last = total_bytes % BS;
    std::vector<std::vector<unsigned char>> out(nblocks);
    std::atomic<size_t> next(0);
    std::atomic<bool> ok(true);
    const int level = zlib_level();
    auto work = [&]() {
        std::vector<unsigned char> raw(BS);
        for (;;) {
            const size_t b = next.fetch_add(1);
            if (b >= nblocks) break;
            const size_t n = (b + 1 == nblocks && last) ? last : BS;
            fill(raw.data(), b * BS, n);
            uLongf cap = compressBound((uLong)n);
            out[b].resize(cap);
            if (compress2(out[b].data(), &cap, raw.data(), (uLong)n, level) != Z_OK) { ok = false; break; }
            out[b].resize(cap);
        }
    };
    unsigned nthreads = std::thread::hardware_concurrency();
    if (nthreads > 16) nthreads = 16;
    if (nthreads > nblocks) nthreads = (unsigned)nblocks;
    if (nthreads <= 1) work();
    else {
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nthreads; ++t) pool.emplace_back(work);
        for (auto &t : pool) t.join();
    }
    if (!ok) return false;
    std::vector<uint32_t> head(3 + nblocks);
    head[0] = (uint32_t)nblocks; head[1] = (uint32_t)BS; head[2] = (uint32_t)last;
    for (size_t b = 0; b < nblocks; ++b) head[3 + b] = (uint32_t)out[b].size();
    size_t w = head.size() * 4;
    if (fwrite(head.data(), 4, head.size(), fp) != head.size()) return false;
    if (keep) keep->bytes.assign((const unsigned char *)head.data(), (const unsigned char *)head.data() + w);
    for (size_t b = 0; b < nblocks; ++b) {
        if (fwrite(out[b].data(), 1, out[b].size(), fp) != out[b].size()) return false;
        if (keep) keep->bytes.insert(keep->bytes.end(), out[b].begin(), out[b].end());
        w += out[b].size();
    }
    *written = w;
    return true;
}

// field values as Float32 from a float or double source
// xfast = false: field[(i*dims[1] + j)*dims[2] + k], the generated code's [x][y][z] arrays (opesci_dump_field_vts_3d);
// xfast = true:  field[i + j*dims[0] + k*dims[0]*dims[1]], the model vectors of the SEG-Y reader (opesci_dump_field_vts)
template <typename T>
int dump_vts(const char *name, const int dims[3], const float spacing[3], int margin, const T *field, int x0, PackedArray *points_cache = nullptr,
             bool xfast = false)
{
    const std::string path = std::string(name) + ".vts";
    const size_t npts = (size_t)dims[0] * dims[1] * dims[2];
    // The data goes to a side file first: the XML header needs the byte offset of the second array.
    const std::string tmp = path + ".part";
    FILE *fb = fopen(tmp.c_str(), "wb");
    if (!fb) return -1;
    size_t w_field = 0, w_pts = 0;
    bool ok = append_compressed(fb, npts * 4, [&](unsigned char *dst, size_t first, size_t n) {
        const size_t e0 = first / 4, ne = n / 4;
        if (sizeof(T) == 4) memcpy(dst, (const unsigned char *)field + first, n);
        else {
            float *d = (float *)dst;
            for (size_t e = 0; e < ne; ++e) d[e] = (float)field[e0 + e];
        }
    }, &w_field);
    // points in the reference's order: i (dims[0]) slowest, k (dims[2]) fastest; coordinate = (index - margin) * spacing
    // evaluated in float like the reference (src/opesciIO.cpp:623-629)
    const bool cached = points_cache && !points_cache->bytes.empty() && points_cache->level == zlib_level() && points_cache->margin == margin &&
                        points_cache->x0 == x0 && memcmp(points_cache->dims, dims, sizeof points_cache->dims) == 0 &&
                        memcmp(points_cache->spacing, spacing, sizeof points_cache->spacing) == 0;
    if (cached) {
        ok = ok && fwrite(points_cache->bytes.data(), 1, points_cache->bytes.size(), fb) == points_cache->bytes.size();
        w_pts = points_cache->bytes.size();
    } else {
        if (points_cache) {
            memcpy(points_cache->dims, dims, sizeof points_cache->dims);
            memcpy(points_cache->spacing, spacing, sizeof points_cache->spacing);
            points_cache->margin = margin; points_cache->x0 = x0; points_cache->level = zlib_level();
            points_cache->bytes.clear();
        }
        ok = ok && append_compressed(fb, npts * 12, [&](unsigned char *dst, size_t first, size_t n) {
            float *d = (float *)dst;
            const size_t c0 = first / 4, nc = n / 4;   // float components; 1 MiB blocks are a multiple of 4 B, not of 12 B
            for (size_t c = 0; c < nc; ++c) {
                const size_t comp = c0 + c, pt = comp / 3;
                const int which = (int)(comp % 3);
                int i, j, k;
                if (xfast) { i = (int)(pt % dims[0]); j = (int)((pt / dims[0]) % dims[1]); k = (int)(pt / ((size_t)dims[0] * dims[1])); }
                else { k = (int)(pt % dims[2]); j = (int)((pt / dims[2]) % dims[1]); i = (int)(pt / ((size_t)dims[2] * dims[1])); }
                d[c] = which == 0 ? (float)(i + x0 - margin) * spacing[0] : which == 1 ? (float)(j - margin) * spacing[1] : (float)(k - margin) * spacing[2];
            }
        }, &w_pts, points_cache);
        if (!ok && points_cache) points_cache->bytes.clear();
    }
    ok = (fclose(fb) == 0) && ok;
    if (!ok) { remove(tmp.c_str()); return -1; }
    FILE *fp = fopen(path.c_str(), "wb");
    if (!fp) { remove(tmp.c_str()); return -1; }
    // VTK's structured extents run fastest-first: k (dims[2]) is the fastest index of the point order above
    fprintf(fp,
            "<?xml version=\"1.0\"?>\n"
            "<VTKFile type=\"StructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\" compressor=\"vtkZLibDataCompressor\">\n"
            "  <StructuredGrid WholeExtent=\"0 %d 0 %d 0 %d\">\n"
            "    <Piece Extent=\"0 %d 0 %d 0 %d\">\n"
            "      <PointData>\n"
            "        <DataArray type=\"Float32\" Name=\"field\" format=\"appended\" offset=\"0\"/>\n"
            "      </PointData>\n"
            "      <Points>\n"
            "        <DataArray type=\"Float32\" Name=\"Points\" NumberOfComponents=\"3\" format=\"appended\" offset=\"%zu\"/>\n"
            "      </Points>\n"
            "    </Piece>\n"
            "  </StructuredGrid>\n"
            "  <AppendedData encoding=\"raw\">\n   _",
            dims[xfast ? 0 : 2] - 1, dims[1] - 1, dims[xfast ? 2 : 0] - 1, dims[xfast ? 0 : 2] - 1, dims[1] - 1, dims[xfast ? 2 : 0] - 1, w_field);
    FILE *fr = fopen(tmp.c_str(), "rb");
    bool good = fr != nullptr;
    if (fr) {
        std::vector<unsigned char> buf((size_t)4 << 20);
        size_t n;
        while ((n = fread(buf.data(), 1, buf.size(), fr)) > 0)
            if (fwrite(buf.data(), 1, n, fp) != n) { good = false; break; }
        fclose(fr);
    }
    fprintf(fp, "\n  </AppendedData>\n</VTKFile>\n");
    good = (fclose(fp) == 0) && good;
    remove(tmp.c_str());
    return good ? 0 : -1;
}

// ------------------------------------------------------------------ per-step snapshot pipeline
struct OutputCfg {
    bool armed = false;
    std::string prefix;
    int field = 0, every = 1;
};
inline OutputCfg &output_cfg() { static OutputCfg c; return c; }
inline std::atomic<int> &files_written() { static std::atomic<int> n(0); return n; }
inline std::atomic<int> &write_errors() { static std::atomic<int> n(0); return n; }

// compute stream --(event)--> io stream: cudaMemcpy2DAsync of one time level into a page-locked stage
// --(host callback)--> writer thread: compress + write, then hand the stage back.
struct Snapshotter {
    bool armed = false;
    OutputCfg cfg;
    cudaStream_t io = nullptr;
    cudaEvent_t ev_step = nullptr, ev_copied[2] = {nullptr, nullptr};
    void *stage[2] = {nullptr, nullptr};
    bool stage_free[2] = {true, true};
    int pending_step[2] = {-1, -1};       // time step whose copy into this stage may still be in flight
    size_t esz = 4, level_elems = 0, pitch_elems = 0;
    int dims[3] = {0, 0, 0}, margin = 2, x0 = 0, rank = 0, nranks = 1, own_lo = 0;
    float spacing[3] = {0, 0, 0};
    const unsigned char *dev = nullptr;
    size_t dev_level_elems = 0;
    int next_stage = 0;
    std::mutex mu;
    std::condition_variable cv;
    std::thread writer;
    struct Job { int stage, ti; };
    PackedArray points;                   // the Points array of this series, compressed once
    std::vector<Job> queue;
    bool quit = false;

    struct Ready { Snapshotter *s; int stage, ti; };
    static void CUDART_CB on_copied(void *arg)
    {
        Ready *r = (Ready *)arg;
        {
            std::lock_guard<std::mutex> lk(r->s->mu);
            r->s->queue.push_back({r->stage, r->ti});
        }
        r->s->cv.notify_all();
        delete r;
    }
    void writer_loop()
    {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return quit || !queue.empty(); });
                if (queue.empty()) return;
                j = queue.front();
                queue.erase(queue.begin());
            }
            char name[1024];
            if (nranks > 1) snprintf(name, sizeof name, "%s%d_r%d", cfg.prefix.c_str(), j.ti, rank);
            else snprintf(name, sizeof name, "%s%d", cfg.prefix.c_str(), j.ti);
            const int rc = esz == 4 ? dump_vts<float>(name, dims, spacing, margin, (const float *)stage[j.stage], x0, &points)
                                    : dump_vts<double>(name, dims, spacing, margin, (const double *)stage[j.stage], x0, &points);
            if (rc == 0) files_written()++;
            else write_errors()++;
            {
                std::lock_guard<std::mutex> lk(mu);
                stage_free[j.stage] = true;
            }
            cv.notify_all();
        }
    }
    // dev_field: base of the field's device array; G: level / row strides in elements; local dims; owned x range
    const char *init(const void *dev_field, size_t esz_, size_t dev_level, size_t dev_pitch, const int ldims[3], int own_lo_, int own_hi_,
                     int global_x0, const double dx[3], int rank_, int nranks_)
    {
        cfg = output_cfg();
        armed = cfg.armed;
        if (!armed) return nullptr;
        esz = esz_; dev = (const unsigned char *)dev_field; dev_level_elems = dev_level; pitch_elems = dev_pitch;
        own_lo = own_lo_;
        dims[0] = own_hi_ - own_lo_; dims[1] = ldims[1]; dims[2] = ldims[2];
        x0 = global_x0; rank = rank_; nranks = nranks_;
        for (int d = 0; d < 3; ++d) spacing[d] = (float)dx[d];
        level_elems = (size_t)dims[0] * dims[1] * dims[2];
        if (cudaStreamCreateWithFlags(&io, cudaStreamNonBlocking) != cudaSuccess) return "snapshot: cudaStreamCreate failed";
        if (cudaEventCreateWithFlags(&ev_step, cudaEventDisableTiming) != cudaSuccess) return "snapshot: cudaEventCreate failed";
        for (int s = 0; s < 2; ++s) {
            if (cudaEventCreateWithFlags(&ev_copied[s], cudaEventDisableTiming) != cudaSuccess) return "snapshot: cudaEventCreate failed";
            if (cudaHostAlloc(&stage[s], level_elems * esz, cudaHostAllocDefault) != cudaSuccess) return "snapshot: page-locked staging allocation failed";
        }
        files_written() = 0;
        write_errors() = 0;
        writer = std::thread([this] { writer_loop(); });
        return nullptr;
    }
    // before the kernels of step `ti` are launched on `st`: the level they overwrite must have left the device
    // (a level written at step s is overwritten at step s+2 at the earliest, staggered and regular alike)
    const char *before_step(cudaStream_t st, int ti)
    {
        if (!armed) return nullptr;
        for (int s = 0; s < 2; ++s)
            if (pending_step[s] >= 0 && pending_step[s] <= ti - 2) {
                if (cudaStreamWaitEvent(st, ev_copied[s], 0) != cudaSuccess) return "snapshot: cudaStreamWaitEvent failed";
                pending_step[s] = -1;
            }
        return nullptr;
    }
    // after step `ti` (which wrote time level `level`) has been queued on `st`
    const char *after_step(cudaStream_t st, int ti, int level)
    {
        if (!armed || (ti % cfg.every) != 0) return nullptr;
        const int s = next_stage;
        next_stage ^= 1;
        {
            // the writer must have finished with this stage (blocks the launching thread, not the GPU)
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return stage_free[s]; });
            stage_free[s] = false;
        }
        if (cudaEventRecord(ev_step, st) != cudaSuccess) return "snapshot: cudaEventRecord failed";
        if (cudaStreamWaitEvent(io, ev_step, 0) != cudaSuccess) return "snapshot: cudaStreamWaitEvent failed";
        const unsigned char *src = dev + ((size_t)level * dev_level_elems + (size_t)own_lo * dims[1] * pitch_elems) * esz;
        if (cudaMemcpy2DAsync(stage[s], (size_t)dims[2] * esz, src, pitch_elems * esz, (size_t)dims[2] * esz, (size_t)dims[0] * dims[1],
                              cudaMemcpyDeviceToHost, io) != cudaSuccess)
            return "snapshot: D2H copy failed";
        if (cudaEventRecord(ev_copied[s], io) != cudaSuccess) return "snapshot: cudaEventRecord failed";
        pending_step[s] = ti;
        if (cudaLaunchHostFunc(io, on_copied, new Ready{this, s, ti}) != cudaSuccess) return "snapshot: cudaLaunchHostFunc failed";
        return nullptr;
    }
    void finish()
    {
        if (!armed) return;
        if (io) cudaStreamSynchronize(io);
        if (writer.joinable()) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return queue.empty() && stage_free[0] && stage_free[1]; });
                quit = true;
            }
            cv.notify_all();
            writer.join();
        }
        for (int s = 0; s < 2; ++s) {
            if (stage[s]) cudaFreeHost(stage[s]);
            if (ev_copied[s]) cudaEventDestroy(ev_copied[s]);
            stage[s] = nullptr; ev_copied[s] = nullptr;
        }
        if (ev_step) cudaEventDestroy(ev_step);
        if (io) cudaStreamDestroy(io);
        ev_step = nullptr; io = nullptr;
        armed = false;
    }
    ~Snapshotter() { finish(); }
};

// ------------------------------------------------------------------ IBM REAL*4 and SEG-Y
// src/opesciIO.cpp:400-417: the four bytes are loaded in memory order into an integer word on a little-endian
// host, sign = bit 31, exponent = bits 24-30 minus 64, mantissa = bits 0-23, value = +-mantissa/2^24 * 16^exponent
// evaluated in double and returned as float.  mantissa/2^24 and 16^exponent are exact in double and so is their
// product (|exponent| <= 64), so ldexp() reproduces pow() bit for bit; the one rounding is the conversion to float.
__host__ __device__ inline float ibm_word_to_float(uint32_t w)
{
    const int exponent = (int)((w & 0x7f000000u) >> 24) - 64;
    // the sign is applied to the INTEGER mantissa (`-mantisse/16777216.0*...`): a zero mantissa gives +0 either way
    const long long m = (long long)(w & 0x00ffffffu);
    const double mant = (double)((w & 0x80000000u) ? -m : m) / 16777216.0;
    return (float)ldexp(mant, 4 * exponent);
}
__host__ __device__ inline uint32_t load_word(const unsigned char *b, bool swap)
{
    return swap ? ((uint32_t)b[3] | (uint32_t)b[2] << 8 | (uint32_t)b[1] << 16 | (uint32_t)b[0] << 24)
                : ((uint32_t)b[0] | (uint32_t)b[1] << 8 | (uint32_t)b[2] << 16 | (uint32_t)b[3] << 24);
}
inline int16_t load_i16(const unsigned char *b, bool swap) { return (int16_t)(swap ? (b[1] | b[0] << 8) : (b[0] | b[1] << 8)); }
inline int32_t load_i32(const unsigned char *b, bool swap) { return (int32_t)load_word(b, swap); }

__host__ __device__ inline size_t model_index(int ix, int iy, int iz, int nx, int ny, int nz, int layout)
{
    return layout == 0 ? (size_t)ix + (size_t)iy * nx + (size_t)iz * nx * ny : ((size_t)ix * ny + iy) * nz + iz;
}

// one thread per sample; threads of a warp walk along a trace (coalesced reads of the raw records); layout 1
// writes are coalesced too (iz fastest), layout 0 scatters with stride nx*ny like the reference's loop
__global__ void segy_decode_kernel(const unsigned char *__restrict__ traces, int ntraces, int nx, int ny, int nz, int swap,
                                   float *__restrict__ out, int layout)
{
    const size_t tracesize = 240 + 4 * (size_t)nz;
    const size_t total = (size_t)ntraces * nz;
    for (size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(s / nz), iz = (int)(s % nz);
        const unsigned char *b = traces + (size_t)i * tracesize + 240 + 4 * (size_t)iz;
        out[model_index(i % nx, i / nx, iz, nx, ny, nz, layout)] = ibm_word_to_float(load_word(b, swap != 0));
    }
}

struct SegyHeader { int nx, ny, nz, ntraces, format; bool swap; size_t tracesize; };
// binary file header fields and the byte-order retry of src/opesciIO.cpp:471-553
inline int segy_parse_header(FILE *fp, SegyHeader &H)
{
    if (fseek(fp, 0, SEEK_END) != 0) return -1;
    const long filesize = ftell(fp);
    unsigned char head[3600];
    if (fseek(fp, 0, SEEK_SET) != 0 || fread(head, 1, 3600, fp) != 3600) return -1;
    H.swap = false;
    H.format = load_i16(head + 3224, false);
    if (H.format < 1 || H.format > 8) {
        H.swap = true;
        H.format = load_i16(head + 3224, true);
        if (H.format < 1 || H.format > 8) { fprintf(stderr, "ERROR: unsupported data sample format code %d\n", H.format); return -1; }
    }
    H.nx = load_i16(head + 3212, H.swap);
    H.nz = load_i16(head + 3220, H.swap);
    if (H.nx <= 0 || H.nz <= 0) return -1;
    H.tracesize = 240 + 4 * (size_t)H.nz;
    H.ntraces = (int)((filesize - 3600) / (long)H.tracesize);
    H.ny = H.ntraces / H.nx;
    return 0;
}

// ------------------------------------------------------------------ DFT resampling (src/opesciHandy.cpp:100-195)
// out[k] = sum_t  re[t]*cos(a) + im[t]*sin(a),  -re[t]*sin(a) + im[t]*cos(a),  a = (float)(2*M_PI*t*k/n); the
// sums are float accumulators updated with a double right-hand side, term by term (what the reference's
// `float += float*cos(float)` compiles to with <cmath>'s double ::cos).
inline void dft(const float *re, const float *im, float *ore, float *oim, int n)
{
    for (int k = 0; k < n; ++k) {
        float sr = 0, si = 0;
        for (int t = 0; t < n; ++t) {
            const float a = 2 * M_PI * t * k / n;
            sr += re[t] * ::cos((double)a) + im[t] * ::sin((double)a);
            si += -re[t] * ::sin((double)a) + im[t] * ::cos((double)a);
        }
        ore[k] = sr;
        oim[k] = si;
    }
}

}  // namespace opesci_io

// ==================================================================== C ABI (include/opesci_io.h)
extern "C" {

int opesci_b200_set_output(const char *prefix, int field, int every)
{
    opesci_io::OutputCfg &c = opesci_io::output_cfg();
    c.armed = prefix != nullptr && every > 0 && field >= 0;
    c.prefix = prefix ? prefix : "";
    c.field = field;
    c.every = every > 0 ? every : 1;
    return 0;
}

int opesci_b200_set_output_level(int zlib_level)
{
    if (zlib_level < 0 || zlib_level > 9) return -1;
    opesci_io::zlib_level() = zlib_level;
    return 0;
}

int opesci_b200_output_stats(int *files_written, int *write_errors)
{
    if (files_written) *files_written = opesci_io::files_written().load();
    if (write_errors) *write_errors = opesci_io::write_errors().load();
    return 0;
}

int opesci_b200_dump_field_vts_3d(const char *name, const int dims[3], const float spacing[3], int margin, const float *field, int x0)
{
    if (!name || !dims || !spacing || !field) return -1;
    return opesci_io::dump_vts<float>(name, dims, spacing, margin, field, x0);
}

int opesci_b200_dump_field_vts(const char *name, const int dims[3], const float spacing[3], const float *field)
{
    if (!name || !dims || !spacing || !field) return -1;
    return opesci_io::dump_vts<float>(name, dims, spacing, 0, field, 0, nullptr, true);
}

int64_t opesci_b200_simple_binary_count(const char *filename)
{
    FILE *fp = fopen(filename, "rb");
    if (!fp) return -1;
    fseek(fp, 0, SEEK_END);
    const long bytes = ftell(fp);
    fclose(fp);
    return (int64_t)(bytes / 4);
}

int opesci_b200_read_simple_binary_ptr(const char *filename, float *array, size_t size)
{
    FILE *fp = fopen(filename, "rb");
    if (!fp) { fprintf(stderr, "ERROR: Failed to open binary file %s\n", filename); return -1; }
    fseek(fp, 0, SEEK_END);
    const size_t have = (size_t)ftell(fp) / 4;
    fseek(fp, 0, SEEK_SET);
    if (have > size) fprintf(stderr, "ERROR: Input file %s size larger than array size\n", filename);
    if (have < size) { fclose(fp); return -2; }
    const size_t got = fread(array, 4, size, fp);
    fclose(fp);
    return got == size ? 0 : -1;
}

float opesci_b200_ibm_to_float(const unsigned char bytes[4], int swap_endian)
{
    return opesci_io::ibm_word_to_float(opesci_io::load_word(bytes, swap_endian != 0));
}

int opesci_b200_read_model_segy(const char *filename, float *array, size_t capacity, int dim[3], float spacing[3], int layout)
{
    using namespace opesci_io;
    FILE *fp = fopen(filename, "rb");
    if (!fp) { fprintf(stderr, "ERROR: Failed to open SEG-Y file %s\n", filename); return -1; }
    SegyHeader H;
    if (segy_parse_header(fp, H)) { fclose(fp); return -1; }
    dim[0] = H.nx; dim[1] = H.ny; dim[2] = H.nz;
    if (H.format != 1) { fprintf(stderr, "ERROR: format code %d not yet supported\n", H.format); fclose(fp); return -1; }
    if (array && capacity < (size_t)H.nx * H.ny * H.nz) { fclose(fp); return -2; }
    std::vector<unsigned char> trace(H.tracesize);
    float x0[2] = {0, 0}, scale = 1.f;
    for (int i = 0; i < H.ntraces; ++i) {
        if (fread(trace.data(), 1, H.tracesize, fp) != H.tracesize) { fclose(fp); return -1; }
        // trace header: coordinate scalar (bytes 71-72), source x/y (73-80), sample interval overloaded as metres (117-118)
        if (i == 0) {
            scale = load_i16(trace.data() + 70, H.swap);
            if (scale < 0) scale = 1.0 / fabs(scale);
            x0[0] = scale * load_i32(trace.data() + 72, H.swap);
            x0[1] = scale * load_i32(trace.data() + 76, H.swap);
        } else if (i == 1) {
            const float x1[2] = {scale * load_i32(trace.data() + 72, H.swap), scale * load_i32(trace.data() + 76, H.swap)};
            const float dx = std::max(fabs(x0[0] - x1[0]), fabs(x0[1] - x1[1]));
            spacing[0] = dx; spacing[1] = dx;
            spacing[2] = scale * load_i16(trace.data() + 116, H.swap);
        }
        if (!array) { if (i >= 1) break; continue; }
        const int ix = i % H.nx, iy = i / H.nx;
        if (iy >= H.ny) break;   // trailing traces of an incomplete record
        for (int iz = 0; iz < H.nz; ++iz)
            array[model_index(ix, iy, iz, H.nx, H.ny, H.nz, layout)] = ibm_word_to_float(load_word(trace.data() + 240 + 4 * (size_t)iz, H.swap));
    }
    fclose(fp);
    return 0;
}

int opesci_b200_segy_decode_device(const void *traces, int ntraces, int nx, int nz, int swap_endian, float *out, int layout, void *stream)
{
    if (!traces || !out || ntraces <= 0 || nx <= 0 || nz <= 0) return -1;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return -1;   // no CPU fallback: the host reader is a separate entry point
    const int ny = ntraces / nx;
    const size_t total = (size_t)(ny * nx) * nz;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 32) blocks = 148 * 32;
    opesci_io::segy_decode_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const unsigned char *)traces, ny * nx, nx, ny, nz, swap_endian, out, layout);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int opesci_b200_read_xyz(const char *filename, float *xyz, int capacity)
{
    FILE *fp = fopen(filename, "r");
    if (!fp) { fprintf(stderr, "ERROR: Failed to open file %s\n", filename); return -1; }
    char line[4096];
    int n = 0;
    bool header = true;
    while (fgets(line, sizeof line, fp)) {
        if (header) { header = false; continue; }          // the first line is a header
        size_t len = strlen(line);
        while (len && (line[len - 1] == '\n' || line[len - 1] == '\r')) line[--len] = 0;
        if (len == 0) continue;
        // `ss >> x >> y >> z` into a zero-initialised triple: missing values stay 0
        float v[3] = {0, 0, 0};
        sscanf(line, "%f %f %f", &v[0], &v[1], &v[2]);
        if (xyz) {
            if (n >= capacity) { fclose(fp); return -2; }
            xyz[3 * n] = v[0]; xyz[3 * n + 1] = v[1]; xyz[3 * n + 2] = v[2];
        }
        ++n;
    }
    fclose(fp);
    return n;
}

int opesci_b200_resample_timeseries(const float *src, int n, float dt, double sdt, float *out, int capacity)
{
    using opesci_io::dft;
    if (fabs(dt - sdt) < std::numeric_limits<float>::epsilon() * (dt + sdt)) {
        if (out) { if (capacity < n) return -2; memcpy(out, src, (size_t)n * 4); }
        return n;
    }
    const int n2 = (int)round(n * sdt / dt);
    if (!out) return n2;
    if (capacity < n2) return -2;
    std::vector<float> zero(n, 0.f), fr(n), fi(n), gr(n2, 0.f), gi(n2, 0.f), tmp(n2);
    dft(src, zero.data(), fr.data(), fi.data(), n);
    const float nrm = 1. / ::sqrt((double)n);
    for (int i = 0; i < n; ++i) { fr[i] *= nrm; fi[i] *= nrm; }
    // spectrum bins: the lower half keeps its place, the upper half stays attached to the END of the spectrum;
    // a longer series gets zeros in between (dt < sdt), a shorter one loses the middle (dt > sdt)
    const int mid = (dt < sdt ? n : n2) / 2, keep = dt < sdt ? n : n2, shift = n2 - n;
    for (int i = 0; i < keep; ++i) {
        if (dt < sdt) { const int dst = i < mid ? i : i + shift; gr[dst] = fr[i]; gi[dst] = fi[i]; }
        else { const int srci = i < mid ? i : i - shift; gr[i] = fr[srci]; gi[i] = fi[srci]; }
    }
    for (int i = 0; i < n2; ++i) gi[i] *= -1;   // conjugate: the forward transform of it is the inverse
    dft(gr.data(), gi.data(), out, tmp.data(), n2);
    const float nrm2 = 1.0 / ::sqrt((double)(float)n2);   // `sqrt((float)snt2)` binds to ::sqrt(double) in the reference's translation unit
    for (int i = 0; i < n2; ++i) out[i] *= nrm2;
    return n2;
}

float opesci_b200_calculate_dt(const float *vp, size_t n, float h)
{
    float maxv = 0;
    for (size_t i = 0; i < n; ++i)
        if (vp[i] > maxv) maxv = vp[i];
    return (6.0 / 7.0) * (1. / sqrt(3.0)) * (h / maxv);
}

void opesci_b200_calculate_lame_constants(const float *vp, const float *vs, const float *rho, size_t n, float *mu, float *lam)
{
    for (size_t i = 0; i < n; ++i) {
        mu[i] = rho[i] * vs[i] * vs[i];
        lam[i] = rho[i] * (vp[i] * vp[i] - 2.0 * vs[i] * vs[i]);
    }
}

}  // extern "C"
