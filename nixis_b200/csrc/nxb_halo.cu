// Per-sweep halo exchange of the sharded erosion stencil over NVLink peer memory.
//
// The stencil has radius 1 and reads only neighbours' height and water (erosion.py:225-247), so
// after every sweep a rank owes each peer the h / w values of the own vertices that peer's rows
// reference.  The peer's halo slots for this rank are ONE contiguous range (partition.py), so the
// exchange is: gather own[send_idx[i]] and store it straight into the peer's state buffer through
// its NVLink-mapped address -- pack and transfer are one kernel, there is no staging buffer, no
// NCCL call and no host involvement.  Completion is signalled with a per-source flag in the peer's
// memory (release at system scope after every block has fenced); the next sweep is preceded by a
// one-warp kernel that spins until the flags of all peers reach the expected sweep number.
//
// Message sizes are 10^4..10^5 vertices x 8 B per peer: latency, not bandwidth, is what matters,
// which is why this avoids per-peer launches (one fused kernel for all peers).
#include "nxb_common.cuh"
#include <string.h>

#define HALO_MAX_PEERS 8

struct HaloPeer {
    float2 *hw;                 // peer's destination {height, water} buffer (peer-mapped)
    uint32_t *flag;             // peer's flag slot for THIS rank
    int64_t dst_off;            // first halo slot (relative to h / w) this rank fills
    int64_t src_begin;          // offset of this peer's list inside the concatenated send list
    int64_t count;
};

struct HaloPutArgs {
    const float2 *hw;           // this rank's freshly written {height, water} buffer
    const int32_t *send_idx;    // concatenated local indices, peer after peer
    HaloPeer peer[HALO_MAX_PEERS];
    int npeers;
    int64_t total;
    uint32_t flag_value;
    unsigned int *ticket;       // device counter for the last-block pattern (zero between launches)
};

__global__ void __launch_bounds__(256)
halo_put_kernel(const __grid_constant__ HaloPutArgs a)
{
    __shared__ bool s_last;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.total; i += (int64_t)gridDim.x * blockDim.x) {
        int p = 0;
#pragma unroll
        for (int q = 1; q < HALO_MAX_PEERS; ++q) if (q < a.npeers && i >= a.peer[q].src_begin) p = q;
        const int64_t j = i - a.peer[p].src_begin;
        const int32_t v = __ldg(a.send_idx + i);
        a.peer[p].hw[a.peer[p].dst_off + j] = a.hw[v];
    }
    __threadfence_system();                 // this thread's peer stores are visible system-wide ...
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int t = atomicAdd(a.ticket, 1u);
        s_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last && threadIdx.x < a.npeers) {  // ... before the last block raises the flags
        __threadfence_system();
        volatile uint32_t *f = a.peer[threadIdx.x].flag;
        *f = a.flag_value;
        __threadfence_system();
    }
    if (s_last && threadIdx.x == 0) *a.ticket = 0;
}

// peers_dev: array of npeers structs in host memory (copied by value into the launch)
NXB_API int nxb_halo_put_f32(const float *hw, const int32_t *send_idx, int npeers,
                             void *const *peer_hw, void *const *peer_flag,
                             const int64_t *dst_off, const int64_t *src_begin, const int64_t *count,
                             uint32_t flag_value, void *ticket, void *stream)
{
    NXB_ARG(npeers >= 0 && npeers <= HALO_MAX_PEERS);
    if (npeers == 0) return NXB_OK;
    NXB_ARG(hw && send_idx && peer_hw && peer_flag && dst_off && src_begin && count && ticket);
    HaloPutArgs a;
    a.hw = (const float2 *)hw; a.send_idx = send_idx; a.npeers = npeers; a.flag_value = flag_value;
    a.ticket = (unsigned int *)ticket;
    int64_t total = 0;
    for (int p = 0; p < npeers; ++p) {
        NXB_ARG(src_begin[p] == total && count[p] >= 0);
        a.peer[p].hw = (float2 *)peer_hw[p]; a.peer[p].flag = (uint32_t *)peer_flag[p];
        a.peer[p].dst_off = dst_off[p]; a.peer[p].src_begin = src_begin[p]; a.peer[p].count = count[p];
        total += count[p];
    }
    for (int p = npeers; p < HALO_MAX_PEERS; ++p) { a.peer[p] = a.peer[0]; a.peer[p].src_begin = total; a.peer[p].count = 0; }
    a.total = total;
    int64_t blocks = (total + 255) / 256;
    if (blocks < 1) blocks = 1;
    if (blocks > 2 * nxb_sm_count()) blocks = 2 * nxb_sm_count();
    halo_put_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

__global__ void halo_wait_kernel(const uint32_t *flags, const int32_t *src_ranks, int npeers, uint32_t target)
{
    // programmatic dependent launch (no-ops without the launch attribute): the sweep behind this
    // kernel may become resident now; this kernel itself may have started before the previous
    // sweep drained, so it must not EXIT before that sweep is complete (griddepcontrol.wait) --
    // the next sweep's own griddepcontrol.wait only covers this kernel.
    asm volatile("griddepcontrol.launch_dependents;");
    if ((int)threadIdx.x < npeers) {
        const volatile uint32_t *f = flags + src_ranks[threadIdx.x];
        // flags only grow; wrap-safe comparison
        while ((int32_t)(*f - target) < 0) { __nanosleep(20); }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    __threadfence_system();
}

int nxb_halo_wait_launch(const void *flags, const int32_t *src_ranks, int npeers, uint32_t target, int pdl, cudaStream_t st)
{
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof lc);
    lc.gridDim = dim3(1); lc.blockDim = dim3(32); lc.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = at;
    lc.numAttrs = pdl ? 1 : 0;
    NXB_CUDA(cudaLaunchKernelEx(&lc, halo_wait_kernel, (const uint32_t *)flags, src_ranks, npeers, target));
    return NXB_OK;
}

// flags: this rank's flag array (one uint32 per source rank); src_ranks: device int32[npeers]
NXB_API int nxb_halo_wait(const void *flags, const int32_t *src_ranks, int npeers, uint32_t target, void *stream)
{
    NXB_ARG(npeers >= 0 && npeers <= 32);
    if (npeers == 0) return NXB_OK;
    NXB_ARG(flags && src_ranks);
    return nxb_halo_wait_launch(flags, src_ranks, npeers, target, 0, (cudaStream_t)stream);
}

// The same wait without a kernel: one stream memory operation per source flag
// (cuStreamWaitValue32, GEQ = wrap-safe "(int32)(*flag - target) >= 0").  The stream stalls in
// the front end, no SM is occupied and no launch / drain sits between two sweeps.  src_ranks: HOST.
// The driver entry point is resolved at run time (no link dependency on libcuda).
#include <cuda.h>
typedef CUresult (*nxb_wait32_fn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static nxb_wait32_fn nxb_wait32()
{
    static nxb_wait32_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (nxb_wait32_fn)p;
    }
    return fn;
}

NXB_API int nxb_halo_wait_stream(const void *flags, const int32_t *src_ranks_host, int npeers, uint32_t target, void *stream)
{
    NXB_ARG(npeers >= 0 && npeers <= 32);
    if (npeers == 0) return NXB_OK;
    NXB_ARG(flags && src_ranks_host);
    nxb_wait32_fn fn = nxb_wait32();
    if (!fn) { nxb_set_error("cuStreamWaitValue32 is not available from this driver"); return NXB_ERR_CUDA; }
    for (int i = 0; i < npeers; ++i) {
        const CUdeviceptr a = (CUdeviceptr)((const uint32_t *)flags + src_ranks_host[i]);
        const CUresult r = fn((CUstream)stream, a, target, CU_STREAM_WAIT_VALUE_GEQ);
        if (r != CUDA_SUCCESS) { nxb_set_error("cuStreamWaitValue32 -> CUresult %d", (int)r); return NXB_ERR_CUDA; }
    }
    return NXB_OK;
}

// ---------------------------------------------------------------------------------------------
// Peer-mapped state buffers without torch symmetric memory: plain cudaMalloc + CUDA IPC handles.
// The owner allocates and exports a 64-byte handle (exchanged by the host through the process
// group); every peer process opens it and gets an address it can store to from its kernels --
// over NVLink between two GPUs, through local memory when both processes share ONE device (which is
// how tests exercise the peer stores, flags and waits of the fused exchange on a single-GPU box).
NXB_API int nxb_peer_alloc(int64_t bytes, void **ptr_out, void *handle64_host)
{
    NXB_ARG(bytes > 0 && ptr_out && handle64_host);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    void *p = nullptr;
    NXB_CUDA(cudaMalloc(&p, (size_t)bytes));
    NXB_CUDA(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); nxb_set_error("cudaIpcGetMemHandle -> %s", cudaGetErrorString(e)); return NXB_ERR_CUDA; }
    memcpy(handle64_host, &h, 64);
    *ptr_out = p;
    return NXB_OK;
}

NXB_API int nxb_peer_open(const void *handle64_host, void **ptr_out)
{
    NXB_ARG(handle64_host && ptr_out);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_host, 64);
    void *p = nullptr;
    NXB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *ptr_out = p;
    return NXB_OK;
}

NXB_API int nxb_peer_close(void *ptr)
{
    if (ptr) NXB_CUDA(cudaIpcCloseMemHandle(ptr));
    return NXB_OK;
}

NXB_API int nxb_peer_free(void *ptr)
{
    if (ptr) NXB_CUDA(cudaFree(ptr));
    return NXB_OK;
}
