// Library-level entry points: version, error text, device info, FP32 peak microbenchmark.
#include "nxb_common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";

void nxb_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

NXB_API int nxb_version(void) { return NXB_ABI_VERSION; }

NXB_API int nxb_last_error(char *buf, int len)
{
    if (!buf || len <= 0) return NXB_ERR_ARG;
    strncpy(buf, g_err, (size_t)len - 1);
    buf[len - 1] = 0;
    return NXB_OK;
}

int nxb_sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

NXB_API int nxb_device_info(int *sm_count, int *sm_clock_khz, int64_t *mem_bytes, int *cc_major, int *cc_minor)
{
    int dev = 0, v = 0;
    NXB_CUDA(cudaGetDevice(&dev));
    if (sm_count) { NXB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev)); *sm_count = v; }
    if (sm_clock_khz) { NXB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, dev)); *sm_clock_khz = v; }
    if (cc_major) { NXB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMajor, dev)); *cc_major = v; }
    if (cc_minor) { NXB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrComputeCapabilityMinor, dev)); *cc_minor = v; }
    if (mem_bytes) {
        size_t fr = 0, tot = 0;
        NXB_CUDA(cudaMemGetInfo(&fr, &tot));
        *mem_bytes = (int64_t)tot;
    }
    return NXB_OK;
}

// ---------------------------------------------------------------------------------
// FP32 roofline denominator: 8 independent FFMA chains per thread, 1024 threads per SM x 2
// CTAs, no memory traffic.  flops = 2 * chains * iters * unroll * threads.
#define FFMA_CHAINS 8
#define FFMA_UNROLL 64

__global__ void __launch_bounds__(512)
ffma_peak_kernel(int iters, float seed, float *sink)
{
    float a[FFMA_CHAINS];
#pragma unroll
    for (int c = 0; c < FFMA_CHAINS; ++c) a[c] = seed + (float)(threadIdx.x + c);
    const float m = 0.999f, b = 0.001f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < FFMA_UNROLL; ++u)
#pragma unroll
            for (int c = 0; c < FFMA_CHAINS; ++c) a[c] = fmaf(a[c], m, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int c = 0; c < FFMA_CHAINS; ++c) s += a[c];
    if (s == 123.456f) sink[0] = s;     // never true; keeps the chains alive
}

NXB_API int nxb_ffma_peak(int iters, double *tflops_out)
{
    NXB_ARG(iters > 0 && tflops_out);
    float *sink = nullptr;
    NXB_CUDA(cudaMalloc(&sink, 4));
    int grid = nxb_sm_count() * 4;
    cudaEvent_t e0, e1;
    NXB_CUDA(cudaEventCreate(&e0));
    NXB_CUDA(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        NXB_CUDA(cudaEventRecord(e0));
        ffma_peak_kernel<<<grid, 512>>>(iters, 1.0f, sink);
        NXB_CUDA(cudaEventRecord(e1));
        NXB_CUDA(cudaEventSynchronize(e1));
        float ms = 0.0f;
        NXB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * FFMA_CHAINS * FFMA_UNROLL * (double)iters * 512.0 * grid;
        double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *tflops_out = best;
    return NXB_OK;
}
