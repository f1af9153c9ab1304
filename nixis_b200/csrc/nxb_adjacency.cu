// Neighbour table (ELL int32[V][6], -1 padded) -- util.py:580-662.
//
// build_adjacency (util.py:591-613) is a serial append in the reference: triangle t = (a,b,c)
// appends b to row a, c to row b, a to row c at the first free slot.  A vertex occurs at most
// once per triangle, so the slot an entry lands in is the RANK of its triangle index among the
// triangles around that vertex.  That makes it parallel and deterministic:
//   pass 1: every triangle corner claims any free slot of its vertex with an atomic counter and
//           stores the key (triangle index << 32 | next vertex)
//   pass 2: every vertex sorts its <= 6 keys (ascending triangle index) and writes the row.
// sort_adjacency (util.py:623-662) is the ring walk, one thread per vertex, reading the unsorted
// table and writing a separate sorted one (the reference's in-place write is a benign race; the
// sequential result is the parity target).
#include "nxb_common.cuh"

struct AdjWorkspace {           // layout of the caller-provided scratch
    // int32 count[V] | int32 overflow flag (+pad to 16 B) | uint64 keys[V][6]
};

static inline int64_t ws_keys_offset(int64_t V) { return ((V + 1) * 4 + 15) / 16 * 16; }

NXB_API int64_t nxb_adj_build_workspace(int64_t V)
{
    return ws_keys_offset(V) + V * 6 * 8;
}

__global__ void __launch_bounds__(256)
adj_collect_kernel(const int32_t *__restrict__ cells, int64_t T, int64_t V, int32_t *__restrict__ count,
                   int32_t *__restrict__ overflow, unsigned long long *__restrict__ keys)
{
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (int64_t)gridDim.x * blockDim.x) {
        int32_t v[3] = {cells[3 * t], cells[3 * t + 1], cells[3 * t + 2]};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            int32_t me = v[c], nx = v[(c + 1) % 3];
            if (me < 0 || me >= V) { atomicOr(overflow, 2); continue; }
            int slot = atomicAdd(count + me, 1);
            if (slot >= 6) { atomicOr(overflow, 1); continue; }
            keys[(int64_t)me * 6 + slot] = ((unsigned long long)t << 32) | (uint32_t)nx;
        }
    }
}

__global__ void __launch_bounds__(256)
adj_rank_kernel(int64_t V, const int32_t *__restrict__ count, const unsigned long long *__restrict__ keys,
                int32_t *__restrict__ adj)
{
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        int n = count[v];
        if (n > 6) n = 6;
        unsigned long long k[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) k[q] = q < n ? keys[v * 6 + q] : ~0ull;
        // 6-element sorting network (12 compare-exchanges)
#define CX(a, b) { unsigned long long lo = k[a] < k[b] ? k[a] : k[b], hi = k[a] < k[b] ? k[b] : k[a]; k[a] = lo; k[b] = hi; }
        CX(0, 5) CX(1, 3) CX(2, 4)
        CX(1, 2) CX(3, 4)
        CX(0, 3) CX(2, 5)
        CX(0, 1) CX(2, 3) CX(4, 5)
        CX(1, 2) CX(3, 4)
#undef CX
#pragma unroll
        for (int q = 0; q < 6; ++q) adj[v * 6 + q] = q < n ? (int32_t)(uint32_t)k[q] : -1;
    }
}

NXB_API int nxb_adj_build(const int32_t *cells, int64_t T, int64_t V, int32_t *adj, void *workspace, void *stream)
{
    NXB_ARG(T >= 0 && V >= 0 && V < (1ll << 31));
    if (V == 0) return NXB_OK;
    NXB_ARG(adj && workspace && (cells || T == 0));
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *count = (int32_t *)workspace;
    int32_t *overflow = count + V;
    unsigned long long *keys = (unsigned long long *)((char *)workspace + ws_keys_offset(V));
    NXB_CUDA(cudaMemsetAsync(workspace, 0, (size_t)ws_keys_offset(V), st));
    if (T > 0) {
        adj_collect_kernel<<<nxb_grid_resident(adj_collect_kernel, 256, 0, ((T) + 256 - 1) / 256), 256, 0, st>>>(cells, T, V, count, overflow, keys);
        NXB_LAUNCH_CHECK();
    }
    adj_rank_kernel<<<nxb_grid_resident(adj_rank_kernel, 256, 0, ((V) + 256 - 1) / 256), 256, 0, st>>>(V, count, keys, adj);
    NXB_LAUNCH_CHECK();
    int32_t flag = 0;
    NXB_CUDA(cudaMemcpyAsync(&flag, overflow, sizeof flag, cudaMemcpyDeviceToHost, st));
    NXB_CUDA(cudaStreamSynchronize(st));
    if (flag & 2) { nxb_set_error("nxb_adj_build: triangle references a vertex id outside [0,V)"); return NXB_ERR_ARG; }
    if (flag & 1) { nxb_set_error("nxb_adj_build: a vertex has more than 6 outgoing edges (inconsistent winding?)"); return NXB_ERR_OVERFLOW; }
    return NXB_OK;
}

// first entry of row_idx (in its stored order) that also occurs in row_nv and is not `v`
// (util.py:623-629 next_vert)
__device__ __forceinline__ int32_t next_in_ring(int32_t v, const int32_t (&row_idx)[6], const int32_t (&row_nv)[6])
{
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        int32_t cand = row_idx[a];
        bool in = false;
#pragma unroll
        for (int q = 0; q < 6; ++q) in |= (row_nv[q] == cand);
        if (in && cand != v) return cand;
    }
    return -1;
}

__device__ __forceinline__ void load_row(const int32_t *__restrict__ adj, int64_t v, int32_t (&row)[6])
{
    const int2 *p = reinterpret_cast<const int2 *>(adj + v * 6);   // rows are 24 B: 8-byte aligned
    int2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    row[0] = a.x; row[1] = a.y; row[2] = b.x; row[3] = b.y; row[4] = c.x; row[5] = c.y;
}

__global__ void __launch_bounds__(256)
adj_sort_kernel(const int32_t *__restrict__ adj_in, int32_t *__restrict__ adj_out, int64_t V)
{
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < V; idx += (int64_t)gridDim.x * blockDim.x) {
        const int n = idx < 12 ? 5 : 6;           // util.py:640-650: vertices 0..11 have valence 5
        int32_t row[6], nrow[6], ring[6] = {-1, -1, -1, -1, -1, -1};
        load_row(adj_in, idx, row);
        int32_t pv = (int32_t)idx, nv = row[0];
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            if (s < n - 1) {
                ring[s] = nv;
                // python indexing: adj[-1] is the last row
                int64_t r = nv >= 0 ? (int64_t)nv : V + nv;
                if (r < 0 || r >= V) r = idx;     // malformed table: stay in bounds
                load_row(adj_in, r, nrow);
                nv = next_in_ring(pv, row, nrow);
                pv = ring[s];
            }
        }
        if (n == 5) ring[4] = nv; else ring[5] = nv;
        int2 *o = reinterpret_cast<int2 *>(adj_out + idx * 6);
        o[0] = make_int2(ring[0], ring[1]);
        o[1] = make_int2(ring[2], ring[3]);
        o[2] = make_int2(ring[4], ring[5]);
    }
}

NXB_API int nxb_adj_sort(const int32_t *adj_in, int32_t *adj_out, int64_t V, void *stream)
{
    NXB_ARG(V >= 0);
    if (V == 0) return NXB_OK;
    NXB_ARG(adj_in && adj_out && adj_in != adj_out);
    adj_sort_kernel<<<nxb_grid_resident(adj_sort_kernel, 256, 0, ((V) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(adj_in, adj_out, V);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}
