// TEST INFRASTRUCTURE ONLY.  extern "C" doors onto the reference's own libopesci helpers
// (/root/reference/src/opesciIO.cpp, opesciHandy.cpp, compiled where they lie by make_io_ref.py) so that
// ctypes can call them: the pin for include/opesci_io.h's readers and resampler.  No reference code here.
#include <cstring>
#include <string>
#include <vector>

#include "opesciHandy.h"
#include "opesciIO.h"

float real2float(const char *xreal);   // src/opesciIO.cpp:400 (external linkage, not in the header)

extern "C" {
int ref_read_model_segy(const char *fn, float *out, long cap, int *dim, float *spacing)
{
    std::vector<float> a;
    const int rc = opesci_read_model_segy(fn, a, dim, spacing);
    if (rc) return rc;
    if ((long)a.size() > cap) return -2;
    memcpy(out, a.data(), a.size() * sizeof(float));
    return 0;
}
int ref_resample(const float *src, int n, float dt, double sdt, float *out, int cap)
{
    std::vector<float> s(src, src + n), r;
    opesci_resample_timeseries(s, dt, sdt, r);
    if ((int)r.size() > cap) return -2;
    memcpy(out, r.data(), r.size() * sizeof(float));
    return (int)r.size();
}
int ref_read_receivers(const char *fn, float *out, int cap_triples)
{
    std::vector<float> a;
    const int rc = opesci_read_receivers(fn, a);
    if (rc) return rc;
    if ((int)a.size() > 3 * cap_triples) return -2;
    memcpy(out, a.data(), a.size() * sizeof(float));
    return (int)a.size() / 3;
}
int ref_read_sources(const char *xyz, const char *fx, const char *fy, const char *fz, float *oxyz, int cap_triples, float *ox, float *oy,
                     float *oz, int cap_samples, int *nsamples)
{
    std::vector<float> a, x, y, z;
    const int rc = opesci_read_souces(xyz, fx, fy, fz, a, x, y, z);
    if (rc) return rc;
    if ((int)a.size() > 3 * cap_triples || (int)x.size() > cap_samples || (int)y.size() > cap_samples || (int)z.size() > cap_samples) return -2;
    memcpy(oxyz, a.data(), a.size() * sizeof(float));
    memcpy(ox, x.data(), x.size() * sizeof(float));
    memcpy(oy, y.data(), y.size() * sizeof(float));
    memcpy(oz, z.data(), z.size() * sizeof(float));
    nsamples[0] = (int)x.size(); nsamples[1] = (int)y.size(); nsamples[2] = (int)z.size();
    return (int)a.size() / 3;
}
int ref_read_simple_binary_ptr(const char *fn, float *array, int size) { return opesci_read_simple_binary_ptr(fn, array, size); }
float ref_real2float(const char *b) { return real2float(b); }
float ref_calculate_dt(const float *vp, long n, float h)
{
    std::vector<float> v(vp, vp + n);
    return opesci_calculate_dt(v, h);
}
void ref_lame(const float *vp, const float *vs, const float *rho, long n, float *mu, float *lam)
{
    std::vector<float> a(vp, vp + n), b(vs, vs + n), c(rho, rho + n), m, l;
    opesci_calculate_lame_costants(a, b, c, m, l);
    memcpy(mu, m.data(), n * sizeof(float));
    memcpy(lam, l.data(), n * sizeof(float));
}
}
