// Shared helpers for libnixis_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/nixis_b200.h"

#define NXB_API extern "C" __attribute__((visibility("default")))

void nxb_set_error(const char *fmt, ...);

#define NXB_CUDA(call)                                                              \
    do {                                                                            \
        cudaError_t e__ = (call);                                                   \
        if (e__ != cudaSuccess) {                                                   \
            nxb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,              \
                          cudaGetErrorString(e__));                                 \
            return NXB_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define NXB_LAUNCH_CHECK()                                                          \
    do {                                                                            \
        cudaError_t e__ = cudaGetLastError();                                       \
        if (e__ != cudaSuccess) {                                                   \
            nxb_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__,          \
                          cudaGetErrorString(e__));                                 \
            return NXB_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define NXB_ARG(cond)                                                               \
    do {                                                                            \
        if (!(cond)) {                                                              \
            nxb_set_error("%s:%d bad argument: %s", __FILE__, __LINE__, #cond);     \
            return NXB_ERR_ARG;                                                     \
        }                                                                           \
    } while (0)

// B200: 148 SMs.  Grids for streaming kernels are sized in whole waves.
int nxb_sm_count();

static inline int nxb_grid_for(int64_t n, int block, int max_ctas_per_sm)
{
    int64_t need = (n + block - 1) / block;
    int64_t cap = (int64_t)nxb_sm_count() * max_ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// Persistent grid for kernel `k`: as many CTAs as are co-resident (occupancy API), never more
// than the work needs.  A fixed "8 CTAs per SM" guess leaves a second, under-occupied wave when
// the register budget allows fewer (seen in profiles/r01_ncu_summary.json: 1184 CTAs launched,
// 888 resident).
template <typename K>
static inline int nxb_grid_resident(K k, int block, size_t dyn_smem, int64_t n_blocks_of_work)
{
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, block, dyn_smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    int64_t cap = (int64_t)nxb_sm_count() * per_sm;
    if (n_blocks_of_work < 1) n_blocks_of_work = 1;
    return (int)(n_blocks_of_work < cap ? n_blocks_of_work : cap);
}

// ---- TMA-style bulk copies (cp.async.bulk, SASS UBLKCP) + mbarrier helpers ----------------------
__device__ __forceinline__ uint32_t nxb_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void nxb_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nxb_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void nxb_fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void nxb_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nxb_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool nxb_mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(nxb_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void nxb_mbar_wait(uint64_t *bar, uint32_t parity)
{
    while (!nxb_mbar_try_wait(bar, parity)) {}
}
// the same on a shared-window address computed once (saves the generic -> shared conversion per call)
__device__ __forceinline__ void nxb_mbar_wait_a(uint32_t bar_addr, uint32_t parity)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar_addr), "r"(parity) : "memory");
    } while (!ok);
}
// the same with a suspend-time hint (ns): the hardware parks the warp until the phase completes or the time is up,
// instead of returning to a poll loop that competes for issue slots with the warps that still have work
__device__ __forceinline__ void nxb_mbar_wait_hint(uint32_t bar_addr, uint32_t parity, uint32_t ns)
{
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar_addr), "r"(parity), "r"(ns) : "memory");
    } while (!ok);
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16; completes on `bar`
__device__ __forceinline__ void nxb_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(nxb_smem_u32(dst)), "l"(src), "r"(bytes), "r"(nxb_smem_u32(bar)) : "memory");
}

// float atomic min/max through the ordered-int trick (works for all finite + inf)
__device__ __forceinline__ void atomic_min_f32(float *addr, float v)
{
    v += 0.0f;  // -0.0 -> +0.0 (its int image would be INT_MIN)
    if (v >= 0.0f) atomicMin((int *)addr, __float_as_int(v));
    else atomicMax((unsigned int *)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f32(float *addr, float v)
{
    v += 0.0f;
    if (v >= 0.0f) atomicMax((int *)addr, __float_as_int(v));
    else atomicMin((unsigned int *)addr, __float_as_uint(v));
}

__device__ __forceinline__ float warp_min(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide min/max merged into gmem minmax[2] with one atomic pair per block.
// NaNs are ignored by fminf/fmaxf, like np.min would NOT -- callers that need NaN
// propagation check finiteness separately (erosion divergence counter).
__device__ __forceinline__ void block_minmax_commit(float lo, float hi, float *minmax)
{
    __shared__ float s_lo[32], s_hi[32];
    lo = warp_min(lo);
    hi = warp_max(hi);
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    if (lane == 0) { s_lo[w] = lo; s_hi[w] = hi; }
    __syncthreads();
    if (w == 0) {
        lo = lane < nw ? s_lo[lane] : __int_as_float(0x7f800000);
        hi = lane < nw ? s_hi[lane] : __int_as_float(0xff800000);
        lo = warp_min(lo);
        hi = warp_max(hi);
        if (lane == 0) { atomic_min_f32(minmax, lo); atomic_max_f32(minmax + 1, hi); }
    }
}
