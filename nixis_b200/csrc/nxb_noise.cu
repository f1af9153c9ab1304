// OpenSimplex noise / fBm kernels and the table handle (C-ABI: include/nixis_b200.h).
#include "nxb_noise.cuh"
#include "nxb_noise3_fast.cuh"
#ifdef NXB_HAVE_NOISE4
#include "nxb_noise4.cuh"
#else
__device__ __forceinline__ float nxb_noise4(float, float, float, float, const uint8_t *, const uint8_t *) { return 0.0f; }
#endif
#include <math.h>
#include <stdlib.h>
#include <string.h>

// ---------------------------------------------------------------------------------
// opensimplex.py:90-112 -- host.  `over` is int32(int32): the 64-bit LCG state is
// truncated to a signed 32-bit value after every step.
static inline int64_t lcg32(int64_t s)
{
    uint64_t u = (uint64_t)s * 6364136223846793005ULL + 1442695040888963407ULL;
    return (int64_t)(int32_t)(uint32_t)u;
}

NXB_API int nxb_init_perm(int64_t seed, int32_t *perm_host, int32_t *pgi_host)
{
    NXB_ARG(perm_host && pgi_host);
    int32_t source[256];
    for (int i = 0; i < 256; ++i) source[i] = i;
    for (int w = 0; w < 3; ++w) seed = lcg32(seed);
    for (int i = 255; i >= 0; --i) {
        seed = lcg32(seed);
        int64_t r = (seed + 31) % (i + 1);
        if (r < 0) r += i + 1;
        perm_host[i] = source[r];
        pgi_host[i] = (perm_host[i] % 24) * 3;
        source[r] = source[i];
    }
    return NXB_OK;
}

NXB_API int nxb_tables_create(const int32_t *perm_host, const int32_t *pgi_host, void **handle_out)
{
    NXB_ARG(perm_host && pgi_host && handle_out);
    NxbTables t;
    for (int i = 0; i < 256; ++i) {
        NXB_ARG(perm_host[i] >= 0 && perm_host[i] < 256);
        NXB_ARG(pgi_host[i] >= 0 && pgi_host[i] < 72 && pgi_host[i] % 3 == 0);
        t.perm8[i] = (uint8_t)perm_host[i];
        int g = pgi_host[i] / 3, q = g / 3, a = g % 3;
        // GRADIENTS_3D (opensimplex.py:52-61): octant q: x positive iff bit0, y negative iff bit1,
        // z negative iff bit2; member a of the octant carries 11 on axis a.
        t.grad8[i] = (uint8_t)(((q & 1) ? 0 : 1) | ((q & 2) ? 2 : 0) | ((q & 4) ? 4 : 0) | (a << 3));
        t.grad4[i] = (uint8_t)(perm_host[i] & 0xFC);
        t.grad2[i] = (uint8_t)(perm_host[i] & 0x0E);
    }
    NxbTables *d = nullptr;
    NXB_CUDA(cudaMalloc(&d, sizeof(NxbTables)));
    NXB_CUDA(cudaMemcpy(d, &t, sizeof(NxbTables), cudaMemcpyHostToDevice));
    *handle_out = d;
    return NXB_OK;
}

NXB_API int nxb_tables_destroy(void *handle)
{
    if (handle) NXB_CUDA(cudaFree(handle));
    return NXB_OK;
}

// ---------------------------------------------------------------------------------
// element-wise array kernels (noisearr2d/3d/4d)
__global__ void __launch_bounds__(256)
noise3_array_kernel(const NxbTables *__restrict__ tab, const float *__restrict__ x, const float *__restrict__ y,
                    const float *__restrict__ z, int64_t n, float *__restrict__ out)
{
    __shared__ __align__(16) uint8_t s_tab[NXB_TABLE_BYTES];
    nxb_stage_tables(tab, s_tab);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = nxb_noise3(x[i], y[i], z[i], s_tab, s_tab + 256);
}

__global__ void __launch_bounds__(256)
noise2_array_kernel(const NxbTables *__restrict__ tab, const float *__restrict__ x, const float *__restrict__ y,
                    int64_t n, float *__restrict__ out)
{
    __shared__ __align__(16) uint8_t s_tab[NXB_TABLE_BYTES];
    nxb_stage_tables(tab, s_tab);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = nxb_noise2(x[i], y[i], s_tab, s_tab + 768);
}

__global__ void __launch_bounds__(256)
noise4_array_kernel(const NxbTables *__restrict__ tab, const float *__restrict__ x, const float *__restrict__ y,
                    const float *__restrict__ z, const float *__restrict__ w, int64_t n, float *__restrict__ out)
{
    __shared__ __align__(16) uint8_t s_tab[NXB_TABLE_BYTES];
    nxb_stage_tables(tab, s_tab);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = nxb_noise4(x[i], y[i], z[i], w[i], s_tab, s_tab + 512);
}

NXB_API int nxb_noise3_f32(void *tables, const float *x, const float *y, const float *z, int64_t n, float *out, void *stream)
{
    NXB_ARG(tables && n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(x && y && z && out);
    noise3_array_kernel<<<nxb_grid_resident(noise3_array_kernel, 256, 0, ((n) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>((const NxbTables *)tables, x, y, z, n, out);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

NXB_API int nxb_noise2_f32(void *tables, const float *x, const float *y, int64_t n, float *out, void *stream)
{
    NXB_ARG(tables && n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(x && y && out);
    noise2_array_kernel<<<nxb_grid_resident(noise2_array_kernel, 256, 0, ((n) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>((const NxbTables *)tables, x, y, n, out);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

NXB_API int nxb_noise4_f32(void *tables, const float *x, const float *y, const float *z, const float *w, int64_t n, float *out, void *stream)
{
    NXB_ARG(tables && n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(x && y && z && w && out);
#ifndef NXB_HAVE_NOISE4
    nxb_set_error("noise4 not built");
    return NXB_ERR_UNSUPPORTED;
#endif
    noise4_array_kernel<<<nxb_grid_resident(noise4_array_kernel, 256, 0, ((n) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>((const NxbTables *)tables, x, y, z, w, n, out);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// ---------------------------------------------------------------------------------
// fBm: all octaves fused, accumulator in a register, one coalesced float4 load and one
// float store per vertex (terrain.py:12-47).  Octave frequencies/amplitudes travel as a
// __grid_constant__ struct so they sit in the constant bank.
#define NXB_MAX_OCT 16
struct FbmParams {
    float freq[NXB_MAX_OCT];
    float amp[NXB_MAX_OCT];     // already * 0.5
    float w[NXB_MAX_OCT];       // 4-D only
    int n_oct;
};

template <int DIM>
__global__ void __launch_bounds__(256)
fbm_kernel(const NxbTables *__restrict__ tab, const float4 *__restrict__ xyz, int64_t n,
           const __grid_constant__ FbmParams prm, const float *__restrict__ init,
           float *__restrict__ out, float *__restrict__ minmax)
{
    __shared__ __align__(16) uint8_t s_tab[NXB_TABLE_BYTES];
    nxb_stage_tables(tab, s_tab);
    float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x) {
        float4 p = __ldg(xyz + v);
        float acc = init ? init[v] : 0.0f;
#pragma unroll 1
        for (int o = 0; o < prm.n_oct; ++o) {
            float f = prm.freq[o];
            float e;
            if (DIM == 3) e = nxb_noise3(p.x * f, p.y * f, p.z * f, s_tab, s_tab + 256);
            else          e = nxb_noise4(p.x * f, p.y * f, p.z * f, prm.w[o], s_tab, s_tab + 512);
            // ((e + 1) * 0.5) * amp    (terrain.py:28; n_strength*radius == amp)
            acc = fmaf(e + 1.0f, prm.amp[o], acc);
        }
        out[v] = acc;
        lo = fminf(lo, acc);
        hi = fmaxf(hi, acc);
    }
    if (minmax) block_minmax_commit(lo, hi, minmax);
}

// ---- branch-free 3-D fBm (nxb_noise3_fast.cuh): one 1024-thread CTA per SM, 128.6 KB of
// conflict-free tables in shared memory, persistent grid-stride over the vertices.  Positions are
// float64 (the reference's `verts`, terrain.py:17) or float4 promoted to float64; the per-octave
// lattice coordinate, the cell and the candidate selection are float64 (exactly the reference's
// decisions), the contributions FP32.
struct FbmFastParams {
    double nr[NXB_MAX_OCT];     // n_freq / world_radius (terrain.py:43), applied to scale * position
    float amp[NXB_MAX_OCT];     // 0.5 * amplitude / 103   (noise3d = v / 103, terrain.py:28)
    double scale;               // nixis.py:249 `points *= world_radius` for positions stored on the unit sphere
    float base;                 // sum of 0.5 * amplitude
    int n_oct;
};

template <bool POS64>
__global__ void __launch_bounds__(1024, 1)
fbm3_fast_kernel(const NxbTables *__restrict__ tab, const void *__restrict__ pos, int64_t n,
                 const __grid_constant__ FbmFastParams prm, const float *__restrict__ init,
                 float *__restrict__ out, float *__restrict__ minmax)
{
    extern __shared__ __align__(128) char nxf_raw[];
    // tables start at a 32 KB-aligned shared address: a lookup address is then (hash & MASK) | lane4
    const uint32_t raw_sa = nxb_smem_u32(nxf_raw);
    const uint32_t base_sa = (raw_sa + 32767u) & ~32767u;
    char *nxf_sm = nxf_raw + (base_sa - raw_sa);
    nxf_build_tables(tab->perm8, tab->grad8, nxf_sm, threadIdx.x, blockDim.x);
    __syncthreads();
    const uint32_t lane4 = base_sa | ((threadIdx.x & 31u) * 4u);
    float lo = __int_as_float(0x7f800000), hi = __int_as_float(0xff800000);
    // The trip count is WARP-UNIFORM (the selection uses full-mask warp votes): a lane past the end
    // of the array evaluates the last vertex again and simply does not store.
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u); base < n;
         base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t vv = base + (threadIdx.x & 31u);
        const bool live = vv < n;
        const int64_t v = live ? vv : n - 1;
        double px, py, pz;
        if (POS64) {
            const double *p = static_cast<const double *>(pos) + 3 * v;
            px = __dmul_rn(__ldg(p), prm.scale); py = __dmul_rn(__ldg(p + 1), prm.scale); pz = __dmul_rn(__ldg(p + 2), prm.scale);
        } else {
            const float4 p = __ldg(static_cast<const float4 *>(pos) + v);
            px = __dmul_rn((double)p.x, prm.scale); py = __dmul_rn((double)p.y, prm.scale); pz = __dmul_rn((double)p.z, prm.scale);
        }
        float acc = init ? init[v] : 0.0f;
#pragma unroll 1
        for (int o = 0; o < prm.n_oct; ++o) {
            const double f = prm.nr[o];                 // terrain.py:17 `verts * n_roughness`
            acc = fmaf(nxf_noise3_x103_d(__dmul_rn(px, f), __dmul_rn(py, f), __dmul_rn(pz, f), nxf_sm, lane4), prm.amp[o], acc);
        }
        acc += prm.base;
        if (live) out[v] = acc;
        lo = fminf(lo, acc);
        hi = fmaxf(hi, acc);
    }
    if (minmax) block_minmax_commit(lo, hi, minmax);
}

static bool g_fbm_fast_attr[64] = {false};

// nr_host[o]: n_freq_o / world_radius; amp_host[o]: octave amplitude in output units (n_strength * radius)
static int fbm3_fast_launch(void *tables, const void *pos, bool pos64, double scale, int64_t n, int cnt,
                            const double *nr_host, const double *amp_host,
                            const float *init, float *out, float *minmax, cudaStream_t st)
{
    FbmFastParams prm;
    memset(&prm, 0, sizeof prm);
    double base = 0.0;
    for (int o = 0; o < cnt; ++o) {
        prm.nr[o] = nr_host[o];
        prm.amp[o] = (float)(0.5 * amp_host[o] / 103.0);
        base += 0.5 * amp_host[o];
    }
    prm.base = (float)base;
    prm.scale = scale;
    prm.n_oct = cnt;
    int dev = 0;
    NXB_CUDA(cudaGetDevice(&dev));
    if (dev < 64 && !g_fbm_fast_attr[dev]) {
        NXB_CUDA(cudaFuncSetAttribute(fbm3_fast_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NXF_SMEM_BYTES + 32768)));
        NXB_CUDA(cudaFuncSetAttribute(fbm3_fast_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NXF_SMEM_BYTES + 32768)));
        g_fbm_fast_attr[dev] = true;
    }
    if (pos64) {
        int grid = nxb_grid_resident(fbm3_fast_kernel<true>, 1024, NXF_SMEM_BYTES + 32768, (n + 1023) / 1024);
        fbm3_fast_kernel<true><<<grid, 1024, NXF_SMEM_BYTES + 32768, st>>>((const NxbTables *)tables, pos, n, prm, init, out, minmax);
    } else {
        int grid = nxb_grid_resident(fbm3_fast_kernel<false>, 1024, NXF_SMEM_BYTES + 32768, (n + 1023) / 1024);
        fbm3_fast_kernel<false><<<grid, 1024, NXF_SMEM_BYTES + 32768, st>>>((const NxbTables *)tables, pos, n, prm, init, out, minmax);
    }
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// 3-D fBm in chunks of NXB_MAX_OCT octaves accumulated through `out`
static int fbm3_run(void *tables, const void *pos, bool pos64, double scale, int64_t n, int n_oct,
                    const double *nr_host, const double *amp_host, const float *init, float *out, float *minmax, void *stream)
{
    NXB_ARG(tables && n >= 0 && n_oct >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(pos && out && (n_oct == 0 || (nr_host && amp_host)));
    const float *cur_init = init;
    int done = 0;
    do {
        const int cnt = n_oct - done < NXB_MAX_OCT ? n_oct - done : NXB_MAX_OCT;
        for (int o = 0; o < cnt; ++o)
            NXB_ARG(fabs(nr_host[done + o] * scale) * 1.75 < NXF_MAX_COORD);        // unit-sphere positions assumed
        float *mm = (done + cnt >= n_oct) ? minmax : nullptr;
        int rc = fbm3_fast_launch(tables, pos, pos64, scale, n, cnt, nr_host + done, amp_host + done, cur_init, out, mm, (cudaStream_t)stream);
        if (rc) return rc;
        done += cnt;
        cur_init = out;
    } while (done < n_oct);
    return NXB_OK;
}

static int fbm4_launch(void *tables, const nxb_float4 *xyz_unit, int64_t n, int n_oct,
                       const double *freq_host, const double *amp_host, const double *w_host,
                       const float *init, float *out, float *minmax, void *stream)
{
    NXB_ARG(tables && n >= 0 && n_oct >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(xyz_unit && out && (n_oct == 0 || (freq_host && amp_host)));
    const float *cur_init = init;
    int done = 0;
    do {   // > NXB_MAX_OCT octaves: accumulate in chunks through `out`
        FbmParams prm;
        memset(&prm, 0, sizeof prm);
        int cnt = n_oct - done < NXB_MAX_OCT ? n_oct - done : NXB_MAX_OCT;
        for (int o = 0; o < cnt; ++o) {
            prm.freq[o] = (float)freq_host[done + o];
            prm.amp[o] = (float)(0.5 * amp_host[done + o]);
            prm.w[o] = w_host ? (float)w_host[done + o] : 0.0f;
        }
        prm.n_oct = cnt;
        float *mm = (done + cnt >= n_oct) ? minmax : nullptr;
        int grid = nxb_grid_resident(fbm_kernel<4>, 256, 0, (n + 255) / 256);
        fbm_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>((const NxbTables *)tables, (const float4 *)xyz_unit, n, prm, cur_init, out, mm);
        NXB_LAUNCH_CHECK();
        done += cnt;
        cur_init = out;
    } while (done < n_oct);
    return NXB_OK;
}

NXB_API int nxb_fbm3_f32(void *tables, const nxb_float4 *xyz_unit, int64_t n, int n_oct,
                         const double *freq_host, const double *amp_host,
                         const float *init, float *out, float *minmax, void *stream)
{
    return fbm3_run(tables, xyz_unit, false, 1.0, n, n_oct, freq_host, amp_host, init, out, minmax, stream);
}

NXB_API int nxb_fbm3_pos64_f32(void *tables, const double *verts, double scale, int64_t n, int n_oct,
                               const double *nr_host, const double *amp_host,
                               const float *init, float *out, float *minmax, void *stream)
{
    return fbm3_run(tables, verts, true, scale, n, n_oct, nr_host, amp_host, init, out, minmax, stream);
}

NXB_API int nxb_fbm4_f32(void *tables, const nxb_float4 *xyz_unit, int64_t n, int n_oct,
                         const double *freq_host, const double *amp_host, const double *w_host,
                         const float *init, float *out, float *minmax, void *stream)
{
    NXB_ARG(w_host || n_oct == 0);
#ifndef NXB_HAVE_NOISE4
    nxb_set_error("noise4 not built");
    return NXB_ERR_UNSUPPORTED;
#endif
    return fbm4_launch(tables, xyz_unit, n, n_oct, freq_host, amp_host, w_host, init, out, minmax, stream);
}
