// Height assembly: min/max, rescale, power_rescale, ocean mask, dtype conversions.
// util.py:110-254, terrain.py:61-72, nixis.py:332-364.  All HBM-bound streaming kernels:
// 16-byte vector accesses, grid = whole waves of 148 SMs, one atomic pair per CTA for reductions.
#include "nxb_common.cuh"
#include <map>
#include <mutex>
#include <utility>
#include <math.h>

#define INF_POS __int_as_float(0x7f800000)
#define INF_NEG __int_as_float(0xff800000)

// ---------------------------------------------------------------------------------
__global__ void minmax_reset_kernel(float *mm) { mm[0] = INF_POS; mm[1] = INF_NEG; }

NXB_API int nxb_minmax_reset(float *minmax, void *stream)
{
    NXB_ARG(minmax);
    minmax_reset_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(minmax);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

__global__ void __launch_bounds__(256)
minmax_kernel(const float *__restrict__ x, int64_t n, float *__restrict__ minmax)
{
    float lo = INF_POS, hi = INF_NEG;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    // vector body (x from cudaMalloc / torch is >= 16-byte aligned; guard anyway)
    if ((reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            float4 v = __ldg(x4 + i);
            lo = fminf(fminf(lo, v.x), fminf(v.y, fminf(v.z, v.w)));
            hi = fmaxf(fmaxf(hi, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += nth) { lo = fminf(lo, x[i]); hi = fmaxf(hi, x[i]); }
    } else {
        for (int64_t i = tid; i < n; i += nth) { lo = fminf(lo, x[i]); hi = fmaxf(hi, x[i]); }
    }
    block_minmax_commit(lo, hi, minmax);
}

NXB_API int nxb_minmax_f32(const float *x, int64_t n, float *minmax, void *stream)
{
    NXB_ARG(n >= 0 && minmax);
    if (n == 0) return NXB_OK;
    NXB_ARG(x);
    minmax_kernel<<<nxb_grid_resident(minmax_kernel, 256, 0, (((n + 3) / 4) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, minmax);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// ---------------------------------------------------------------------------------
// util.py:110-175
struct RescaleArgs {
    float x_min, x_max, lower, upper, mid;
    int has_mid, mode;
};

__device__ __forceinline__ float rescale_one(float v, const RescaleArgs &a)
{
    if (a.mode == 0) {
        if (!a.has_mid) return ((v - a.x_min) / (a.x_max - a.x_min)) * (a.upper - a.lower) + a.lower;
        if (v <= a.mid) return ((v - a.x_min) / (a.mid - a.x_min)) * (a.mid - a.lower) + a.lower;
        return ((v - a.mid) / (a.x_max - a.mid)) * (a.upper - a.mid) + a.mid;
    }
    if (a.mode == 1) return v <= a.mid ? ((v - a.x_min) / (a.mid - a.x_min)) * (a.mid - a.lower) + a.lower : v;
    return v >= a.mid ? ((v - a.mid) / (a.x_max - a.mid)) * (a.upper - a.mid) + a.mid : v;
}

__global__ void __launch_bounds__(256)
rescale_kernel(const float *__restrict__ x, int64_t n, const __grid_constant__ RescaleArgs a, float *__restrict__ out)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        float4 *o4 = reinterpret_cast<float4 *>(out);
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            float4 v = x4[i];
            v.x = rescale_one(v.x, a); v.y = rescale_one(v.y, a); v.z = rescale_one(v.z, a); v.w = rescale_one(v.w, a);
            o4[i] = v;
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += nth) out[i] = rescale_one(x[i], a);
    } else {
        for (int64_t i = tid; i < n; i += nth) out[i] = rescale_one(x[i], a);
    }
}

NXB_API int nxb_rescale_f32(const float *x, int64_t n, float x_min, float x_max, float lower, float upper,
                            int has_mid, float mid, int mode, float *out, void *stream)
{
    NXB_ARG(n >= 0 && mode >= 0 && mode <= 2 && (mode == 0 || has_mid));
    if (n == 0) return NXB_OK;
    NXB_ARG(x && out);
    RescaleArgs a = {x_min, x_max, lower, upper, mid, has_mid, mode};
    rescale_kernel<<<nxb_grid_resident(rescale_kernel, 256, 0, (((n + 3) / 4) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, a, out);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// ---------------------------------------------------------------------------------
// terrain.py:61-72
__global__ void __launch_bounds__(256)
mask_le_kernel(const float *__restrict__ h, int64_t n, float level, uint8_t *__restrict__ mask)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    if (((reinterpret_cast<uintptr_t>(h) & 15) | (reinterpret_cast<uintptr_t>(mask) & 3)) == 0) {
        const float4 *h4 = reinterpret_cast<const float4 *>(h);
        uint32_t *m4 = reinterpret_cast<uint32_t *>(mask);
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += nth) {
            float4 v = __ldg(h4 + i);
            m4[i] = (v.x <= level ? 1u : 0u) | (v.y <= level ? 0x100u : 0u) | (v.z <= level ? 0x10000u : 0u) | (v.w <= level ? 0x1000000u : 0u);
        }
        for (int64_t i = (n4 << 2) + tid; i < n; i += nth) mask[i] = h[i] <= level;
    } else {
        for (int64_t i = tid; i < n; i += nth) mask[i] = h[i] <= level;
    }
}

NXB_API int nxb_mask_le_f32(const float *h, int64_t n, float level, uint8_t *mask, void *stream)
{
    NXB_ARG(n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(h && mask);
    mask_le_kernel<<<nxb_grid_resident(mask_le_kernel, 256, 0, (((n + 3) / 4) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(h, n, level, mask);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// ---------------------------------------------------------------------------------
// power_rescale statistics (util.py:198-214).  The reference scans sequentially with
//     if sel and x < lower: lower = x   elif sel and x > upper: upper = x
// so an element that sets a new running minimum never updates the maximum.  Exact parallel
// form: describe a run of elements by  S = (has, F, U, M):
//     F = first selected value, M = min selected value,
//     U = max over selected values that are >= the minimum of the selected values before them
//         INSIDE the run (those can never be running-minimum records, whatever came earlier).
// Two adjacent runs A,B combine as  F = has_A ? F_A : F_B,  M = min(M_A, M_B),
//     U = max(U_A, U_B, F_B >= M_A ? F_B : -inf)
// (the only element of B whose fate depends on A is its first selected one).  With the scan
// starting from lower = x_max, upper = x_min:
//     mask_lower = min(x_max, M),  mask_upper = max(x_min, U, F >= x_max ? F : -inf).
// The combine is associative but not commutative: every level below combines in index order.
struct PStat { float has, F, U, M; };

__device__ __forceinline__ PStat pstat_empty() { return PStat{0.0f, 0.0f, INF_NEG, INF_POS}; }

__device__ __forceinline__ PStat pstat_combine(const PStat &a, const PStat &b)
{
    if (a.has == 0.0f) return b;
    if (b.has == 0.0f) return a;
    PStat r;
    r.has = 1.0f; r.F = a.F; r.M = fminf(a.M, b.M);
    r.U = fmaxf(fmaxf(a.U, b.U), b.F >= a.M ? b.F : INF_NEG);
    return r;
}

__device__ __forceinline__ void pstat_push(PStat &s, float v)
{
    if (s.has == 0.0f) { s.has = 1.0f; s.F = v; s.M = v; return; }
    if (v >= s.M) s.U = fmaxf(s.U, v); else s.M = v;
}

__device__ __forceinline__ PStat pstat_shfl_down(const PStat &s, int d)
{
    PStat r;
    r.has = __shfl_down_sync(0xffffffffu, s.has, d);
    r.F = __shfl_down_sync(0xffffffffu, s.F, d);
    r.U = __shfl_down_sync(0xffffffffu, s.U, d);
    r.M = __shfl_down_sync(0xffffffffu, s.M, d);
    return r;
}

#define PSTAT_BLOCK 256
#define PSTAT_PER_THREAD 8

// Each CTA owns ONE contiguous slice of the array and walks it tile by tile; inside a tile
// thread t owns PSTAT_PER_THREAD consecutive elements, so every combine is in index order.
__global__ void __launch_bounds__(PSTAT_BLOCK)
power_stats_kernel(const float *__restrict__ x, const uint8_t *__restrict__ mask, int64_t n, int sel_mode,
                   PStat *__restrict__ block_stats, unsigned int *__restrict__ ticket, float *__restrict__ summary)
{
    __shared__ PStat s_warp[PSTAT_BLOCK / 32];
    __shared__ bool s_last;
    const int64_t per_cta = (n + gridDim.x - 1) / gridDim.x;
    const int64_t begin = per_cta * blockIdx.x, end = begin + per_cta < n ? begin + per_cta : n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    PStat run = pstat_empty();          // meaningful in thread 0 only
    for (int64_t tile = begin; tile < end; tile += (int64_t)PSTAT_BLOCK * PSTAT_PER_THREAD) {
        PStat s = pstat_empty();
        int64_t i0 = tile + (int64_t)threadIdx.x * PSTAT_PER_THREAD;
#pragma unroll
        for (int e = 0; e < PSTAT_PER_THREAD; ++e) {
            int64_t i = i0 + e;
            if (i < end) {
                bool sel = (mask[i] != 0) == (sel_mode == 1);
                if (sel) pstat_push(s, x[i]);
            }
        }
        // ordered warp reduction: lane l absorbs lane l+d (which holds later elements)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            PStat o = pstat_shfl_down(s, d);
            if ((lane & (2 * d - 1)) == 0) s = pstat_combine(s, o);
        }
        if (lane == 0) s_warp[warp] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 0; w < PSTAT_BLOCK / 32; ++w) run = pstat_combine(run, s_warp[w]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        block_stats[blockIdx.x] = run;
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        PStat tot = pstat_empty();
        const volatile PStat *bs = block_stats;
        for (unsigned int b = 0; b < gridDim.x; ++b) {
            PStat o = {bs[b].has, bs[b].F, bs[b].U, bs[b].M};
            tot = pstat_combine(tot, o);
        }
        summary[0] = tot.has; summary[1] = tot.F; summary[2] = tot.U; summary[3] = tot.M;
        *ticket = 0;
    }
}

// Scratch (per-CTA summaries + the last-CTA ticket) is keyed by (device, stream): calls on different
// streams of one device never share a ticket or block_stats array, calls on one stream are ordered by
// the stream.  The map itself is guarded by a mutex (host threads).
struct PStatScratch { PStat *blocks; unsigned int *ticket; int cap; };
static std::mutex g_pstat_mutex;
static std::map<std::pair<int, void *>, PStatScratch> g_pstat;

NXB_API int nxb_power_summary_f32(const float *x, const uint8_t *mask, int64_t n, int sel_mode,
                                  float *summary4, void *stream)
{
    NXB_ARG(n >= 0 && summary4 && (sel_mode == 0 || sel_mode == 1));
    NXB_ARG(n == 0 || (x && mask));
    int dev = 0;
    NXB_CUDA(cudaGetDevice(&dev));
    int grid = nxb_grid_resident(power_stats_kernel, PSTAT_BLOCK, 0, (n + PSTAT_BLOCK * PSTAT_PER_THREAD - 1) / (PSTAT_BLOCK * PSTAT_PER_THREAD));
    std::lock_guard<std::mutex> lock(g_pstat_mutex);
    PStatScratch &sc = g_pstat[std::make_pair(dev, stream)];        // value-initialised (null, 0) on first use
    if (sc.cap < grid) {
        if (sc.blocks) {                        // regrow: the stream may still be using the old arrays
            NXB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
            cudaFree(sc.blocks);
        }
        if (sc.ticket) cudaFree(sc.ticket);
        int cap = nxb_sm_count() * 4;
        if (cap < grid) cap = grid;
        NXB_CUDA(cudaMalloc(&sc.blocks, sizeof(PStat) * cap));
        NXB_CUDA(cudaMalloc(&sc.ticket, sizeof(unsigned int)));
        NXB_CUDA(cudaMemset(sc.ticket, 0, sizeof(unsigned int)));
        sc.cap = cap;
    }
    power_stats_kernel<<<grid, PSTAT_BLOCK, 0, (cudaStream_t)stream>>>(x, mask, n, sel_mode, sc.blocks, sc.ticket, summary4);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// util.py:224-252: normalise -> pow -> denormalise for selected elements, fused; `shift` is
// subtracted from every element afterwards (nixis.py:359).
__global__ void __launch_bounds__(256)
power_apply_kernel(const float *__restrict__ x, const uint8_t *__restrict__ mask, int64_t n, int sel_mode,
                   float lo, float hi, float power, float shift, float *__restrict__ out)
{
    const float range = hi - lo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = x[i];
        bool sel = sel_mode >= 0 && ((mask[i] != 0) == (sel_mode == 1));
        if (sel) {
            float t = (v - lo) / range;
            t = powf(t, power);
            v = t * range + lo;
        }
        out[i] = v - shift;
    }
}

NXB_API int nxb_power_apply_f32(const float *x, const uint8_t *mask, int64_t n, int sel_mode,
                                float lo, float hi, float power, float shift, float *out, void *stream)
{
    NXB_ARG(n >= 0 && (sel_mode == 0 || sel_mode == 1 || sel_mode == -1));
    if (n == 0) return NXB_OK;
    NXB_ARG(x && out && (mask || sel_mode == -1));
    // sel_mode -1 is mode=None: no element is selected (util.py:224-252 never fire)
    power_apply_kernel<<<nxb_grid_resident(power_apply_kernel, 256, 0, ((n) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(x, mask, n, sel_mode, lo, hi, power, shift, out);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
f32_to_f64_kernel(const float *__restrict__ s, int64_t n, double *__restrict__ d)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = (double)s[i];
}
__global__ void __launch_bounds__(256)
f64_to_f32_kernel(const double *__restrict__ s, int64_t n, float *__restrict__ d)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) d[i] = (float)s[i];
}

NXB_API int nxb_f32_to_f64(const float *src, int64_t n, double *dst, void *stream)
{
    NXB_ARG(n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(src && dst);
    f32_to_f64_kernel<<<nxb_grid_resident(f32_to_f64_kernel, 256, 0, ((n) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(src, n, dst);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

NXB_API int nxb_f64_to_f32(const double *src, int64_t n, float *dst, void *stream)
{
    NXB_ARG(n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(src && dst);
    f64_to_f32_kernel<<<nxb_grid_resident(f64_to_f32_kernel, 256, 0, ((n) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(src, n, dst);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_kernel(const float *__restrict__ src, const int32_t *__restrict__ idx, int64_t n, float *__restrict__ dst)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}
__global__ void __launch_bounds__(256)
scatter_kernel(const float *__restrict__ src, const int32_t *__restrict__ idx, int64_t n, float *__restrict__ dst)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) dst[idx[i]] = src[i];
}

NXB_API int nxb_gather_f32(const float *src, const int32_t *idx, int64_t n, float *dst, void *stream)
{
    NXB_ARG(n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(src && idx && dst);
    gather_kernel<<<nxb_grid_resident(gather_kernel, 256, 0, ((n) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(src, idx, n, dst);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

NXB_API int nxb_scatter_f32(const float *src, const int32_t *idx, int64_t n, float *dst, void *stream)
{
    NXB_ARG(n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(src && idx && dst);
    scatter_kernel<<<nxb_grid_resident(scatter_kernel, 256, 0, ((n) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(src, idx, n, dst);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}
