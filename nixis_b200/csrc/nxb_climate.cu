// Per-vertex climate kernels: climate.py:345-372 assign_surface_temp, :415-448
// sample_insolation (and its 360-rotation drivers :450-490, :503-536), :551-577
// interpolate_insolation, :579-597 calc_yearly_insolation.  The reference runs 360 (rotations) or
// 360 x 2 (days x {slice, interpolate}) full passes over the vertex arrays and recomputes each
// vertex's latitude every time; here every driver is ONE pass: positions are read once, the
// rotation / day loop runs in registers with the reference's float32 rounding of the accumulator
// after every step, and the result is written once.
//
// Arithmetic: float64, the reference's operation order, compiled with -fmad=false (numba never
// contracts a*b+c).  cos / sin of the rotation and tilt angles are evaluated on the HOST with libm
// (the values numba itself uses); the per-vertex asin / atan2 / cos run on the device (CUDA libm,
// <= 2 ulp), so results equal the reference to a few float32 ulps, not bit for bit.
#include "nxb_common.cuh"
#include <math.h>

namespace {

constexpr double kPi = 3.141592653589793;
constexpr int kMaxRot = 360, kMaxTilt = 360;

__device__ __forceinline__ void xyz2latlon(double x, double y, double z, double r, double &lat, double &lon)
{
    // util.py:59-75
    double q = z / r;
    if (!(q > -1.0)) q = -1.0;
    if (!(q < 1.0)) q = 1.0;
    lat = asin(q) * (180.0 / kPi);
    lon = atan2(y, x) * (180.0 / kPi);
}

__device__ __forceinline__ double latitude(double z, double r)
{
    double q = z / r;
    if (!(q > -1.0)) q = -1.0;
    if (!(q < 1.0)) q = 1.0;
    return asin(q) * (180.0 / kPi);
}

__global__ void __launch_bounds__(256)
surface_temp_kernel(const double *__restrict__ verts, int64_t n, double radius, double tilt, float *__restrict__ out)
{
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        const double lat = latitude(verts[3 * v + 2], radius);
        double c = cos(fabs(lat - tilt) * kPi / 180);
        if (!(c > 0.0)) c = 0.0;
        out[v] = (float)c;
    }
}

struct Angles { double c[kMaxRot], s[kMaxRot]; };

// arr[t][v] += sum over rotations, float32 rounding after every rotation (climate.py:448).
// One thread per (tilt t, vertex v); rotations loop in registers.
__global__ void __launch_bounds__(256)
insolation_kernel(const double *__restrict__ verts, int64_t n, double radius,
                  const __grid_constant__ Angles rot, int n_rot,
                  const double *__restrict__ tilt_cs /*[n_tilt][2]*/, int n_tilt, float *__restrict__ arr)
{
    const int64_t total = n * n_tilt;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i / n, v = i - t * n;
        const double x = verts[3 * v], y = verts[3 * v + 1], z = verts[3 * v + 2];
        const double ct = tilt_cs[2 * t], st = tilt_cs[2 * t + 1];
        float acc = arr[i];
        for (int r = 0; r < n_rot; ++r) {
            const double cr = rot.c[r], sr = rot.s[r];
            const double rx = x * cr - y * sr;
            const double ry = x * sr + y * cr;
            const double tx = rx * ct + z * st;
            const double tz = z * ct - rx * st;
            double lat, lon;
            xyz2latlon(tx, ry, tz, radius, lat, lon);
            double a = cos(fabs(lat) * kPi / 180), b = cos(lon * kPi / 180);
            if (!(a > 0.0)) a = 0.0;
            if (!(b > 0.0)) b = 0.0;
            acc = (float)((double)acc + a * b);
        }
        arr[i] = acc;
    }
}

// interpolate_insolation (climate.py:551-577) against n_tab lookup tables of 181 float32 each:
// out = sum over tables of the float32 daily value, float32 additions in table order
// (calc_yearly_insolation, climate.py:585-588); n_tab == 1 is the plain daily map.
__global__ void __launch_bounds__(256)
interpolate_kernel(const double *__restrict__ verts, int64_t n, double radius, const float *__restrict__ tables,
                   int n_tab, float *__restrict__ out)
{
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        const double lat = latitude(verts[3 * v + 2], radius);
        const long long lower = (long long)floor(lat), upper = (long long)ceil(lat);
        const int li = (int)(lower < 0 ? 181 + lower : lower), ui = (int)(upper < 0 ? 181 + upper : upper);
        const double frac = lat - (double)lower;
        const double span = (double)(upper - lower);
        float acc = 0.0f;
        for (int d = 0; d < n_tab; ++d) {
            const float tl = __ldg(tables + d * 181 + li);
            float daily;
            if (lower == upper) daily = tl;
            else {
                const float diff = __fsub_rn(__ldg(tables + d * 181 + ui), tl);
                daily = (float)((double)tl + frac * ((double)diff / span));
            }
            acc = n_tab == 1 ? daily : __fadd_rn(acc, daily);
        }
        out[v] = acc;
    }
}

void host_angles(const double *deg, int n, double *c, double *s)
{
    for (int i = 0; i < n; ++i) { c[i] = cos(deg[i] * kPi / 180); s[i] = sin(deg[i] * kPi / 180); }
}

}  // namespace

// climate.py:345-372.  The altitude term is multiplied by the literal alt_intensity = 0
// (climate.py:350-351: h2 = rescale(altitudes, 0, 0) is all zeros, `h2[v] > 0` never holds), so the
// temperature depends on latitude and tilt only.
NXB_API int nxb_climate_surface_temp_f32(const double *verts, int64_t n, double radius, double tilt, float *out, void *stream)
{
    NXB_ARG(n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(verts && out && radius != 0.0);
    surface_temp_kernel<<<nxb_grid_resident(surface_temp_kernel, 256, 0, (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        verts, n, radius, tilt, out);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// sample_insolation and its drivers.  rot_deg / tilt_deg: HOST arrays (degrees); arr: device
// float32 [n_tilt][n], accumulated in place; tilt_scratch: device double[2 * n_tilt].
NXB_API int nxb_climate_insolation_f32(const double *verts, int64_t n, double radius, const double *rot_deg, int n_rot,
                                       const double *tilt_deg, int n_tilt, double *tilt_scratch, float *arr, void *stream)
{
    NXB_ARG(n >= 0 && n_rot >= 0 && n_rot <= kMaxRot && n_tilt >= 1 && n_tilt <= kMaxTilt);
    if (n == 0 || n_rot == 0) return NXB_OK;
    NXB_ARG(verts && rot_deg && tilt_deg && tilt_scratch && arr && radius != 0.0);
    Angles rot;
    host_angles(rot_deg, n_rot, rot.c, rot.s);
    double tc[kMaxTilt], ts[kMaxTilt], packed[2 * kMaxTilt];
    host_angles(tilt_deg, n_tilt, tc, ts);
    for (int i = 0; i < n_tilt; ++i) { packed[2 * i] = tc[i]; packed[2 * i + 1] = ts[i]; }
    // pageable source: the copy is staged before the call returns, so the stack buffer may die
    NXB_CUDA(cudaMemcpyAsync(tilt_scratch, packed, sizeof(double) * 2 * n_tilt, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    const int64_t total = n * n_tilt;
    insolation_kernel<<<nxb_grid_resident(insolation_kernel, 256, 0, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        verts, n, radius, rot, n_rot, tilt_scratch, n_tilt, arr);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// The 181 lookup vertices of calc_insolation_slice (climate.py:507-515, util.py:80-88), host libm;
// row order is the reference's: rows 0..90 = latitude 0..90, rows 91..180 = latitude -90..-1.
NXB_API int nxb_climate_slice_verts(double radius, double *verts_host /*[181][3]*/)
{
    NXB_ARG(verts_host);
    for (int i = -90; i <= 90; ++i) {
        const int row = i < 0 ? 181 + i : i;
        const double lat = (double)i, lon = 0.0;
        verts_host[3 * row + 0] = radius * cos(lat * (kPi / 180)) * cos(lon * (kPi / 180));
        verts_host[3 * row + 1] = radius * cos(lat * (kPi / 180)) * sin(lon * (kPi / 180));
        verts_host[3 * row + 2] = radius * sin(lat * (kPi / 180));
    }
    return NXB_OK;
}

// climate.py:193-201 (host libm)
NXB_API double nxb_climate_seasonal_tilt(double axial_tilt, double degrees)
{
    return sin(degrees * kPi / 180) * axial_tilt;
}

NXB_API int nxb_climate_interpolate_f32(const double *verts, int64_t n, double radius, const float *tables, int n_tab,
                                        float *out, void *stream)
{
    NXB_ARG(n >= 0 && n_tab >= 1);
    if (n == 0) return NXB_OK;
    NXB_ARG(verts && tables && out && radius != 0.0);
    interpolate_kernel<<<nxb_grid_resident(interpolate_kernel, 256, 0, (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        verts, n, radius, tables, n_tab, out);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}
