// Closed-form icosphere in meshzoo's vertex / cell order (util.py:17-50 -> meshzoo.icosa_sphere;
// layout from SURVEY.md Appendix B, parity unpinned: meshzoo is absent from the reference tree).
//
//   vertices: [12 corners | 30 edges x (k-1) | 20 faces x (k-1)(k-2)/2 interiors, row-major]
//   cells   : 20 faces x k^2 triangles; per face row i: (k-i) "up" then (k-i-1) "down"
//
// One thread per vertex / per triangle; nothing is read from memory except constant tables,
// so the d=2500 mesh (62.5 M vertices, 125 M triangles) is produced at store bandwidth on
// the device instead of 17 s on the host (nixis.py:242) plus a 1.5 GB upload.
//
// Vertex arithmetic is FP64 with every product/sum rounded once (no FMA) so that it is
// bit-identical to the numpy oracle (oracle/icosphere.py) -- see the formulas there.
#include "nxb_common.cuh"

#define PHI 1.618033988749895   /* (1 + sqrt(5)) / 2 rounded to double */

__constant__ double c_corner[12][3] = {
    {-1, PHI, 0}, {1, PHI, 0}, {-1, -PHI, 0}, {1, -PHI, 0},
    {0, -1, PHI}, {0, 1, PHI}, {0, -1, -PHI}, {0, 1, -PHI},
    {PHI, 0, -1}, {PHI, 0, 1}, {-PHI, 0, -1}, {-PHI, 0, 1}};

__constant__ int c_face[20][3] = {
    {0, 11, 5}, {0, 5, 1}, {0, 1, 7}, {0, 7, 10}, {0, 10, 11},
    {1, 5, 9}, {5, 11, 4}, {11, 10, 2}, {10, 7, 6}, {7, 1, 8},
    {3, 9, 4}, {3, 4, 2}, {3, 2, 6}, {3, 6, 8}, {3, 8, 9},
    {4, 9, 5}, {2, 4, 11}, {6, 2, 10}, {8, 6, 7}, {9, 8, 1}};

// CPython set-iteration order of the sorted corner pairs
__constant__ int c_edge[30][2] = {
    {3, 4}, {4, 9}, {8, 9}, {0, 5}, {2, 11}, {1, 9}, {0, 11}, {7, 10}, {6, 8}, {4, 5},
    {3, 9}, {3, 6}, {5, 9}, {4, 11}, {0, 1}, {0, 7}, {2, 4}, {10, 11}, {0, 10}, {1, 5},
    {2, 10}, {1, 8}, {6, 7}, {6, 10}, {3, 8}, {5, 11}, {2, 3}, {1, 7}, {2, 6}, {7, 8}};

// per face, per side (0: c0->c1, 1: c1->c2, 2: c2->c0): edge id * 2 + reversed
// (reversed = the side runs from the larger corner id to the smaller one; edge nodes are
// laid out from the smaller corner id).  Derived from c_face / c_edge; tests/test_mesh.py
// re-derives it.
__constant__ int c_face_edge[20][3] = {
    {12, 51, 7}, {6, 39, 29}, {28, 54, 31}, {30, 14, 37}, {36, 34, 13}, {38, 24, 11}, {50, 27, 18},
    {35, 41, 8}, {15, 45, 46}, {55, 42, 59}, {20, 3, 1}, {0, 33, 52}, {53, 56, 23}, {22, 16, 49},
    {48, 4, 21}, {2, 25, 19}, {32, 26, 9}, {57, 40, 47}, {17, 44, 58}, {5, 43, 10}};

// interiors before row r (r >= 1) of a face: sum_{q=1}^{r-1} (n-q-1)
__device__ __forceinline__ int64_t interior_before_row(int64_t n, int64_t r)
{
    return (r - 1) * (n - 1) - (r - 1) * r / 2;
}

// unit-sphere position of vertex v of the k-division icosphere, FP64, roundings as in oracle/icosphere.py
__device__ __forceinline__ void icosa_point(int64_t n, int64_t v, double &x, double &y, double &z)
{
    const int64_t n_edge = 30 * (n - 1), per_face = (n - 1) * (n - 2) / 2;
    if (v < 12) {
        x = c_corner[v][0]; y = c_corner[v][1]; z = c_corner[v][2];
    } else if (v < 12 + n_edge) {
        int64_t w = v - 12;
        int e = (int)(w / (n - 1));
        int64_t j = w % (n - 1) + 1;
        double t = __ddiv_rn((double)j, (double)n), u = __dsub_rn(1.0, t);
        const double *a = c_corner[c_edge[e][0]], *b = c_corner[c_edge[e][1]];
        x = __dadd_rn(__dmul_rn(u, a[0]), __dmul_rn(t, b[0]));
        y = __dadd_rn(__dmul_rn(u, a[1]), __dmul_rn(t, b[1]));
        z = __dadd_rn(__dmul_rn(u, a[2]), __dmul_rn(t, b[2]));
    } else {
        int64_t w = v - 12 - n_edge;
        int f = (int)(w / per_face);
        int64_t q = w % per_face;
        // row i (1..n-2): largest i with interior_before_row(i) <= q
        double disc = (double)(2 * n - 3) * (double)(2 * n - 3) - 8.0 * (double)q;
        int64_t i = (int64_t)(((double)(2 * n - 3) - sqrt(disc)) * 0.5) + 1;
        if (i < 1) i = 1;
        if (i > n - 2) i = n - 2;
        while (i > 1 && interior_before_row(n, i) > q) --i;
        while (i < n - 2 && interior_before_row(n, i + 1) <= q) ++i;
        int64_t j = q - interior_before_row(n, i) + 1;
        double bi = __ddiv_rn((double)i, (double)n), bj = __ddiv_rn((double)j, (double)n);
        double b0 = __dsub_rn(__dsub_rn(1.0, bi), bj);
        const double *c0 = c_corner[c_face[f][0]], *c1 = c_corner[c_face[f][1]], *c2 = c_corner[c_face[f][2]];
        x = __dadd_rn(__dadd_rn(__dmul_rn(b0, c0[0]), __dmul_rn(bj, c1[0])), __dmul_rn(bi, c2[0]));
        y = __dadd_rn(__dadd_rn(__dmul_rn(b0, c0[1]), __dmul_rn(bj, c1[1])), __dmul_rn(bi, c2[1]));
        z = __dadd_rn(__dadd_rn(__dmul_rn(b0, c0[2]), __dmul_rn(bj, c1[2])), __dmul_rn(bi, c2[2]));
    }
    double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    x = __ddiv_rn(x, nrm); y = __ddiv_rn(y, nrm); z = __ddiv_rn(z, nrm);
}

__global__ void __launch_bounds__(256)
icosa_points_kernel(int k, int64_t v_begin, int64_t v_end, float4 *__restrict__ out32, double *__restrict__ out64)
{
    for (int64_t v = v_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < v_end;
         v += (int64_t)gridDim.x * blockDim.x) {
        double x, y, z;
        icosa_point(k, v, x, y, z);
        int64_t o = v - v_begin;
        if (out32) out32[o] = make_float4((float)x, (float)y, (float)z, 0.0f);
        if (out64) { out64[3 * o] = x; out64[3 * o + 1] = y; out64[3 * o + 2] = z; }
    }
}

// Edge lengths of the radius-scaled icosphere for adjacency rows [v_begin, v_end) (GLOBAL vertex ids
// in adj_rows), positions recomputed in FP64 on the fly: dist[(v - v_begin) * 6 + q] =
// |R p_v - R p_n| (erosion.py:34-40, 227-229), rounded once to FP32; 0 for -1 pads.
__global__ void __launch_bounds__(256)
icosa_edge_lengths_kernel(int k, const int32_t *__restrict__ adj_rows, int64_t v_begin, int64_t v_end, double radius,
                          float *__restrict__ dist)
{
    for (int64_t v = v_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < v_end;
         v += (int64_t)gridDim.x * blockDim.x) {
        double x, y, z;
        icosa_point(k, v, x, y, z);
        x *= radius; y *= radius; z *= radius;
        const int64_t o = (v - v_begin) * 6;
        for (int q = 0; q < 6; ++q) {
            const int32_t n = adj_rows[o + q];
            float r = 0.0f;
            if (n >= 0) {
                double a, b, c;
                icosa_point(k, n, a, b, c);
                const double ax = x - a * radius, ay = y - b * radius, az = z - c * radius;
                r = (float)sqrt(ax * ax + ay * ay + az * az);
            }
            dist[o + q] = r;
        }
    }
}

NXB_API int nxb_mesh_icosa_edge_lengths(int k, const int32_t *adj_rows, int64_t v_begin, int64_t v_end, double radius,
                                        float *dist, void *stream)
{
    NXB_ARG(k >= 1 && k <= 14000 && adj_rows && dist);
    int64_t V = 10 * (int64_t)k * k + 2;
    NXB_ARG(0 <= v_begin && v_begin <= v_end && v_end <= V);
    if (v_end == v_begin) return NXB_OK;
    int grid = nxb_grid_resident(icosa_edge_lengths_kernel, 256, 0, (v_end - v_begin + 255) / 256);
    icosa_edge_lengths_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(k, adj_rows, v_begin, v_end, radius, dist);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// local node (row r, col c) of face f -> global vertex id
__device__ __forceinline__ int32_t face_node(int64_t n, int f, int64_t r, int64_t c)
{
    const int64_t eb = 12, nm1 = n - 1;
    if (r == 0) {
        if (c == 0) return c_face[f][0];
        if (c == n) return c_face[f][1];
        int fe = c_face_edge[f][0];
        int64_t q = c - 1;
        return (int32_t)(eb + (fe >> 1) * nm1 + ((fe & 1) ? n - 2 - q : q));
    }
    if (r == n) return c_face[f][2];
    if (c == 0) {                       // side 2 runs c2->c0, i.e. downwards
        int fe = c_face_edge[f][2];
        int64_t q = r - 1;
        return (int32_t)(eb + (fe >> 1) * nm1 + ((fe & 1) ? q : n - 2 - q));
    }
    if (c == n - r) {
        int fe = c_face_edge[f][1];
        int64_t q = r - 1;
        return (int32_t)(eb + (fe >> 1) * nm1 + ((fe & 1) ? n - 2 - q : q));
    }
    int64_t base = 12 + 30 * nm1 + (int64_t)f * (nm1 * (n - 2) / 2);
    return (int32_t)(base + interior_before_row(n, r) + (c - 1));
}

// triangle t of the k-division icosphere (meshzoo cell order) -> its three global vertex ids
__device__ __forceinline__ void icosa_cell(int64_t n, int64_t t, int32_t &a, int32_t &b, int32_t &c)
{
    const int64_t nn = n * n;
    int f = (int)(t / nn);
    int64_t l = t % nn;
    // triangles before row i: 2*n*i - i*i ; row i = largest with that <= l
    int64_t i = (int64_t)((double)n - sqrt((double)(nn - l)));
    if (i < 0) i = 0;
    if (i > n - 1) i = n - 1;
    while (i > 0 && 2 * n * i - i * i > l) --i;
    while (i < n - 1 && 2 * n * (i + 1) - (i + 1) * (i + 1) <= l) ++i;
    int64_t r = l - (2 * n * i - i * i);
    if (r < n - i) {                 // up: (i,j) (i,j+1) (i+1,j)
        int64_t j = r;
        a = face_node(n, f, i, j); b = face_node(n, f, i, j + 1); c = face_node(n, f, i + 1, j);
    } else {                         // down: (i,j+1) (i+1,j+1) (i+1,j)
        int64_t j = r - (n - i);
        a = face_node(n, f, i, j + 1); b = face_node(n, f, i + 1, j + 1); c = face_node(n, f, i + 1, j);
    }
}

__global__ void __launch_bounds__(256)
icosa_cells_kernel(int k, int64_t t_begin, int64_t t_end, int32_t *__restrict__ cells)
{
    const int64_t n = k;
    for (int64_t t = t_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < t_end;
         t += (int64_t)gridDim.x * blockDim.x) {
        int32_t a, b, c;
        icosa_cell(n, t, a, b, c);
        int64_t o = 3 * (t - t_begin);
        cells[o] = a; cells[o + 1] = b; cells[o + 2] = c;
    }
}

// ---------------------------------------------------------------------------------------------
// Neighbour rows of a vertex RANGE of the closed-form icosphere, without the cell array and without
// any O(V) table: what build_adjacency + sort_adjacency (util.py:591-662) produce for rows
// [v_begin, v_end), from a scan of the closed-form triangle generator (SURVEY 8e "each GPU builds
// rows for its owned vertices ... analytic from layout").
//   pass 1: one thread per triangle of the WHOLE mesh (no memory traffic: the triangle is computed,
//           not read); a corner inside the range claims a slot of its vertex and stores the key
//           (triangle index << 32 | next vertex) -- the slot an entry lands in after sorting the keys
//           is the rank of its triangle among the triangles round the vertex, i.e. the reference's
//           append order;
//   pass 2: one thread per vertex of the range sorts its <= 6 keys, recomputes each incident
//           triangle to learn the vertex BEFORE it as well, and walks the ring (util.py:623-662)
//           on those (next, previous) pairs: two ring vertices are adjacent exactly when they are
//           consecutive round the vertex (every 3-cycle of this mesh is a face), so `next_vert`'s
//           membership test `i in adj[nv]` is answered by the incident triangles alone.
// Workspace: int32 count[n] + uint64 keys[n][6] = 52 B per vertex of the RANGE.
static inline int64_t rows_keys_offset(int64_t n) { return ((n + 1) * 4 + 15) / 16 * 16; }

NXB_API int64_t nxb_mesh_icosa_adj_rows_workspace(int64_t n_rows)
{
    return rows_keys_offset(n_rows) + n_rows * 6 * 8;
}

__global__ void __launch_bounds__(256)
icosa_rows_collect_kernel(int k, int64_t t_begin, int64_t t_end, int64_t v_begin, int64_t v_end,
                          int32_t *__restrict__ count, int32_t *__restrict__ overflow, unsigned long long *__restrict__ keys)
{
    const int64_t n = k;
    for (int64_t t = t_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < t_end;
         t += (int64_t)gridDim.x * blockDim.x) {
        int32_t v[3];
        icosa_cell(n, t, v[0], v[1], v[2]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int64_t me = v[c];
            if (me < v_begin || me >= v_end) continue;
            const int slot = atomicAdd(count + (me - v_begin), 1);
            if (slot >= 6) { atomicOr(overflow, 1); continue; }
            keys[(me - v_begin) * 6 + slot] = ((unsigned long long)t << 32) | (uint32_t)v[(c + 1) % 3];
        }
    }
}

__global__ void __launch_bounds__(256)
icosa_rows_finish_kernel(int k, int64_t v_begin, int64_t v_end, const int32_t *__restrict__ count,
                         const unsigned long long *__restrict__ keys, int32_t *__restrict__ adj_sorted,
                         int32_t *__restrict__ adj_unsorted)
{
    const int64_t nk = k;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < v_end - v_begin; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t idx = v_begin + i;
        int cnt = count[i];
        if (cnt > 6) cnt = 6;
        unsigned long long kk[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) kk[q] = q < cnt ? keys[i * 6 + q] : ~0ull;
#define CX(a, b) { unsigned long long lo = kk[a] < kk[b] ? kk[a] : kk[b], hi = kk[a] < kk[b] ? kk[b] : kk[a]; kk[a] = lo; kk[b] = hi; }
        CX(0, 5) CX(1, 3) CX(2, 4)
        CX(1, 2) CX(3, 4)
        CX(0, 3) CX(2, 5)
        CX(0, 1) CX(2, 3) CX(4, 5)
        CX(1, 2) CX(3, 4)
#undef CX
        int32_t nx[6], pr[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            nx[q] = -1; pr[q] = -2;
            if (q < cnt) {
                nx[q] = (int32_t)(uint32_t)kk[q];
                int32_t a, b, c;
                icosa_cell(nk, (int64_t)(kk[q] >> 32), a, b, c);
                pr[q] = (a == (int32_t)idx) ? c : ((b == (int32_t)idx) ? a : b);     // the corner before idx in (a, b, c)
            }
        }
        if (adj_unsorted) {
#pragma unroll
            for (int q = 0; q < 6; ++q) adj_unsorted[i * 6 + q] = nx[q];
        }
        // ring walk, util.py:638-662 (vertices 0..11 have valence 5)
        const int n = idx < 12 ? 5 : 6;
        int32_t ring[6] = {-1, -1, -1, -1, -1, -1};
        int32_t pv = (int32_t)idx, nv = nx[0];
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            if (s < n - 1) {
                ring[s] = nv;
                // the two ring neighbours of nv round idx: the vertex before idx in the triangle whose
                // `next` is nv, and the `next` of the triangle whose `previous` is nv
                int32_t ra = -3, rb = -3;
#pragma unroll
                for (int q = 0; q < 6; ++q) {
                    if (nx[q] == nv && nx[q] >= 0) ra = pr[q];
                    if (pr[q] == nv) rb = nx[q];
                }
                // next_vert: first entry of the UNSORTED row that is adjacent to nv and is not pv
                int32_t found = -1;
#pragma unroll
                for (int q = 5; q >= 0; --q) {
                    const int32_t cand = nx[q];
                    if (cand >= 0 && (cand == ra || cand == rb) && cand != pv) found = cand;
                }
                nv = found;
                pv = ring[s];
            }
        }
        if (n == 5) ring[4] = nv; else ring[5] = nv;
        int2 *o = reinterpret_cast<int2 *>(adj_sorted + i * 6);
        o[0] = make_int2(ring[0], ring[1]);
        o[1] = make_int2(ring[2], ring[3]);
        o[2] = make_int2(ring[4], ring[5]);
    }
}

NXB_API int nxb_mesh_icosa_adj_rows(int k, int64_t v_begin, int64_t v_end, int32_t *adj_sorted, int32_t *adj_unsorted,
                                    void *workspace, void *stream)
{
    NXB_ARG(k >= 1 && k <= 14000);
    const int64_t V = 10 * (int64_t)k * k + 2, T = 20 * (int64_t)k * k;
    NXB_ARG(0 <= v_begin && v_begin <= v_end && v_end <= V);
    const int64_t n = v_end - v_begin;
    if (n == 0) return NXB_OK;
    NXB_ARG(adj_sorted && workspace && (((uintptr_t)adj_sorted) & 7) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    int32_t *count = (int32_t *)workspace;
    int32_t *overflow = count + n;
    unsigned long long *keys = (unsigned long long *)((char *)workspace + rows_keys_offset(n));
    NXB_CUDA(cudaMemsetAsync(workspace, 0, (size_t)rows_keys_offset(n), st));
    // triangles that can touch the range: a range of face interiors only is touched by its own faces'
    // triangles; the skeleton (corners, edges: ids below 12 + 30 (k-1)) is touched by every face
    const int64_t skel = 12 + 30 * ((int64_t)k - 1), per_face = ((int64_t)k - 1) * ((int64_t)k - 2) / 2, nn = (int64_t)k * k;
    int64_t t_begin = 0, t_end = T;
    if (v_begin >= skel && per_face > 0) {
        t_begin = (v_begin - skel) / per_face * nn;
        t_end = ((v_end - 1 - skel) / per_face + 1) * nn;
    }
    icosa_rows_collect_kernel<<<nxb_grid_resident(icosa_rows_collect_kernel, 256, 0, (t_end - t_begin + 255) / 256), 256, 0, st>>>(
        k, t_begin, t_end, v_begin, v_end, count, overflow, keys);
    NXB_LAUNCH_CHECK();
    icosa_rows_finish_kernel<<<nxb_grid_resident(icosa_rows_finish_kernel, 256, 0, (n + 255) / 256), 256, 0, st>>>(
        k, v_begin, v_end, count, keys, adj_sorted, adj_unsorted);
    NXB_LAUNCH_CHECK();
    int32_t flag = 0;
    NXB_CUDA(cudaMemcpyAsync(&flag, overflow, sizeof flag, cudaMemcpyDeviceToHost, st));
    NXB_CUDA(cudaStreamSynchronize(st));
    if (flag) { nxb_set_error("nxb_mesh_icosa_adj_rows: a vertex has more than 6 outgoing edges"); return NXB_ERR_OVERFLOW; }
    return NXB_OK;
}

NXB_API int nxb_mesh_icosa_points(int k, int64_t v_begin, int64_t v_end, nxb_float4 *xyz_f32, double *xyz_f64, void *stream)
{
    NXB_ARG(k >= 1 && k <= 14000);      // V and T must fit int32 vertex ids
    int64_t V = 10 * (int64_t)k * k + 2;
    NXB_ARG(0 <= v_begin && v_begin <= v_end && v_end <= V);
    if (v_end == v_begin) return NXB_OK;
    icosa_points_kernel<<<nxb_grid_resident(icosa_points_kernel, 256, 0, ((v_end - v_begin) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(
        k, v_begin, v_end, (float4 *)xyz_f32, xyz_f64);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

NXB_API int nxb_mesh_icosa_cells(int k, int64_t t_begin, int64_t t_end, int32_t *cells, void *stream)
{
    NXB_ARG(k >= 1 && k <= 14000 && cells);
    int64_t T = 20 * (int64_t)k * k;
    NXB_ARG(0 <= t_begin && t_begin <= t_end && t_end <= T);
    if (t_end == t_begin) return NXB_OK;
    icosa_cells_kernel<<<nxb_grid_resident(icosa_cells_kernel, 256, 0, ((t_end - t_begin) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(k, t_begin, t_end, cells);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

__global__ void __launch_bounds__(256)
xyz_f64_to_f32_kernel(const double *__restrict__ in, int64_t n, double scale, float4 *__restrict__ out)
{
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (int64_t)gridDim.x * blockDim.x)
        out[v] = make_float4((float)(in[3 * v] * scale), (float)(in[3 * v + 1] * scale), (float)(in[3 * v + 2] * scale), 0.0f);
}

NXB_API int nxb_xyz_f64_to_f32(const double *xyz_f64, int64_t n, double scale, nxb_float4 *xyz_f32, void *stream)
{
    NXB_ARG(n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(xyz_f64 && xyz_f32);
    xyz_f64_to_f32_kernel<<<nxb_grid_resident(xyz_f64_to_f32_kernel, 256, 0, ((n) + 256 - 1) / 256), 256, 0, (cudaStream_t)stream>>>(xyz_f64, n, scale, (float4 *)xyz_f32);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// =================================================================================================
// Equirectangular export (util.py:290-308, 343-367; nixis.py:270-302).
//
// The reference builds a scipy KD-tree over all vertices (26 s at k=2500, nixis.py:268-269) and asks
// it for the 3 nearest vertices of every pixel direction.  On the closed-form icosphere the same
// answer comes from arithmetic: find the icosahedron face(s) whose cone contains the direction,
// convert to the face's barycentric grid, and compare exact FP64 distances to the 4x4 window of grid
// nodes around it (plus the windows of neighbouring faces when the direction is within 1.5 grid
// steps of a face edge).  Checked against scipy.spatial.KDTree in tests/test_gpu_export.py.

// util.py:290-308 make_ll_arr + :79-88 latlon2xyz: xyz of every pixel of a width x height
// equirectangular map (row 0 = north pole; the longitude of the last column stops one step short of
// +180, util.py:298-299).
__global__ void __launch_bounds__(256)
ll_grid_kernel(int width, int height, double radius, double *__restrict__ out)
{
    const double latpercent = 180.0 / (double)(height - 1), lonpercent = 360.0 / (double)width;
    const double D2R = 3.141592653589793 / 180;
    const int64_t n = (int64_t)width * height;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int h = (int)(i / width), w = (int)(i % width);
        // every product / sum rounded once, in the reference's order (no FMA contraction)
        const double lat = fmax(__dsub_rn(90.0, __dmul_rn((double)h, latpercent)), -90.0);
        const double lon = fmax(__dadd_rn(-180.0, __dmul_rn((double)w, lonpercent)), -180.0);
        const double rl = __dmul_rn(lat, D2R), ro = __dmul_rn(lon, D2R);
        const double rc = __dmul_rn(radius, cos(rl));
        out[3 * i] = __dmul_rn(rc, cos(ro));
        out[3 * i + 1] = __dmul_rn(rc, sin(ro));
        out[3 * i + 2] = __dmul_rn(radius, sin(rl));
    }
}

NXB_API int nxb_ll_grid_f64(int width, int height, double radius, double *xyz_out, void *stream)
{
    NXB_ARG(width >= 1 && height >= 2 && xyz_out);
    const int64_t n = (int64_t)width * height;
    ll_grid_kernel<<<nxb_grid_resident(ll_grid_kernel, 256, 0, (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(width, height, radius, xyz_out);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

__device__ __forceinline__ void top3_insert(double d2, int32_t id, double (&bd)[3], int32_t (&bi)[3])
{
    if (id == bi[0] || id == bi[1] || id == bi[2]) return;      // the same vertex seen from two faces
    if (d2 < bd[2]) {
        if (d2 < bd[1]) {
            bd[2] = bd[1]; bi[2] = bi[1];
            if (d2 < bd[0]) { bd[1] = bd[0]; bi[1] = bi[0]; bd[0] = d2; bi[0] = id; }
            else { bd[1] = d2; bi[1] = id; }
        } else { bd[2] = d2; bi[2] = id; }
    }
}

// query: double[n][3] positions (any radius > 0); out: dists double[n][3] ascending, ids int64[n][3]
__global__ void __launch_bounds__(256)
ico_nearest3_kernel(int k, double radius, const double *__restrict__ query, int64_t n_q,
                    double *__restrict__ dists, long long *__restrict__ ids)
{
    __shared__ double s_inv[20][9];
    if (threadIdx.x < 20) {         // rows of M^-1 for M = [c0 c1 c2]: (c1 x c2, c2 x c0, c0 x c1) / det
        const int f = threadIdx.x;
        const double *a = c_corner[c_face[f][0]], *b = c_corner[c_face[f][1]], *c = c_corner[c_face[f][2]];
        const double bc[3] = {b[1] * c[2] - b[2] * c[1], b[2] * c[0] - b[0] * c[2], b[0] * c[1] - b[1] * c[0]};
        const double ca[3] = {c[1] * a[2] - c[2] * a[1], c[2] * a[0] - c[0] * a[2], c[0] * a[1] - c[1] * a[0]};
        const double ab[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        const double det = a[0] * bc[0] + a[1] * bc[1] + a[2] * bc[2];
        for (int q = 0; q < 3; ++q) { s_inv[f][q] = bc[q] / det; s_inv[f][3 + q] = ca[q] / det; s_inv[f][6 + q] = ab[q] / det; }
    }
    __syncthreads();
    const int64_t n = k;
    const double margin = 1.5 / (double)n;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_q; i += (int64_t)gridDim.x * blockDim.x) {
        const double qx = query[3 * i], qy = query[3 * i + 1], qz = query[3 * i + 2];
        const double qn = sqrt(qx * qx + qy * qy + qz * qz);
        const double ux = qx / qn, uy = qy / qn, uz = qz / qn;           // direction on the unit sphere
        double bd[3] = {1e300, 1e300, 1e300};
        int32_t bi[3] = {-1, -2, -3};
        for (int f = 0; f < 20; ++f) {
            const double *m = s_inv[f];
            const double al = m[0] * ux + m[1] * uy + m[2] * uz, be = m[3] * ux + m[4] * uy + m[5] * uz,
                         ga = m[6] * ux + m[7] * uy + m[8] * uz;
            const double s = al + be + ga;
            if (!(s > 0.0)) continue;
            const double b0 = al / s, b1 = be / s, b2 = ga / s;           // barycentric on the flat face
            if (fmin(b0, fmin(b1, b2)) < -margin) continue;
            const int64_t i0 = (int64_t)floor(b2 * (double)n), j0 = (int64_t)floor(b1 * (double)n);
            const double *c0 = c_corner[c_face[f][0]], *c1 = c_corner[c_face[f][1]], *c2 = c_corner[c_face[f][2]];
            for (int64_t r = i0 - 1; r <= i0 + 2; ++r) {
                if (r < 0 || r > n) continue;
                for (int64_t c = j0 - 1; c <= j0 + 2; ++c) {
                    if (c < 0 || r + c > n) continue;
                    const double wi = (double)r / (double)n, wj = (double)c / (double)n, w0 = 1.0 - wi - wj;
                    double px = w0 * c0[0] + wj * c1[0] + wi * c2[0], py = w0 * c0[1] + wj * c1[1] + wi * c2[1],
                           pz = w0 * c0[2] + wj * c1[2] + wi * c2[2];
                    const double inv = 1.0 / sqrt(px * px + py * py + pz * pz);
                    px = px * inv - ux; py = py * inv - uy; pz = pz * inv - uz;
                    top3_insert(px * px + py * py + pz * pz, face_node(n, f, r, c), bd, bi);
                }
            }
        }
        // exact distances to the mesh's own vertex positions (radius-scaled), then order by them
        double dd[3];
        for (int t = 0; t < 3; ++t) {
            if (bi[t] < 0) { dd[t] = 1e300; continue; }
            double x, y, z;
            icosa_point(n, bi[t], x, y, z);
            const double ax = qx - x * radius, ay = qy - y * radius, az = qz - z * radius;
            dd[t] = sqrt(ax * ax + ay * ay + az * az);
        }
#define SWAP3(a, b) if (dd[b] < dd[a]) { double td = dd[a]; dd[a] = dd[b]; dd[b] = td; int32_t ti = bi[a]; bi[a] = bi[b]; bi[b] = ti; }
        SWAP3(0, 1) SWAP3(1, 2) SWAP3(0, 1)
#undef SWAP3
        for (int t = 0; t < 3; ++t) { dists[3 * i + t] = dd[t]; ids[3 * i + t] = bi[t]; }
    }
}

NXB_API int nxb_ico_nearest3_f64(int k, double radius, const double *query_xyz, int64_t n, double *dists, int64_t *ids, void *stream)
{
    NXB_ARG(k >= 1 && k <= 14000 && radius > 0.0 && n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(query_xyz && dists && ids);
    ico_nearest3_kernel<<<nxb_grid_resident(ico_nearest3_kernel, 256, 0, (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        k, radius, query_xyz, n, dists, (long long *)ids);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

// util.py:343-367 make_gray_array: inverse-distance blend of the 3 nearest vertices' values,
// float64, in the reference's operation order; int() truncation; out int32[n].
__global__ void __launch_bounds__(256)
idw_gray_kernel(const double *__restrict__ dists, const long long *__restrict__ ids, const double *__restrict__ colors,
                int64_t n, int32_t *__restrict__ out)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double d0 = dists[3 * i], d1 = dists[3 * i + 1], d2 = dists[3 * i + 2];
        const double sd = __dadd_rn(__dadd_rn(d0, d1), d2);
        const double w0 = __ddiv_rn(1.0, __ddiv_rn(__dadd_rn(d0, 0.00001), sd)), w1 = __ddiv_rn(1.0, __ddiv_rn(__dadd_rn(d1, 0.00001), sd)),
                     w2 = __ddiv_rn(1.0, __ddiv_rn(__dadd_rn(d2, 0.00001), sd));
        const double t = __dadd_rn(__dadd_rn(w0, w1), w2);
        const double v = __dadd_rn(__dadd_rn(__dmul_rn(colors[ids[3 * i]], __ddiv_rn(w0, t)), __dmul_rn(colors[ids[3 * i + 1]], __ddiv_rn(w1, t))),
                                   __dmul_rn(colors[ids[3 * i + 2]], __ddiv_rn(w2, t)));
        out[i] = (int32_t)v;        // int(): truncation toward zero
    }
}

// One export map in one launch (nixis.py:349, 386-389, 417 feeding util.py:393-429): the per-vertex
// colour of `build_image_data` is derived on the fly from the device-resident field,
//     c = ((x - x_min) / x_range) * new_range + lower [+ add]        (util.py:143, FP64, same order)
//     c = (double)(uint16)c  when quantize != 0                       (.astype('uint16') truncation)
// and blended exactly like make_gray_array.  No V-sized temporary exists.  field_kind 0: float32
// heights, 1: uint8 mask.  out_bits 8 / 16.
struct IdwMapArgs {
    const double *dists; const long long *ids; const void *field; int field_kind;
    double x_min, x_range, new_range, lower, add; int quantize; int out_bits; void *out; int64_t n;
};

__global__ void __launch_bounds__(256)
idw_map_kernel(const __grid_constant__ IdwMapArgs a)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
        const double d0 = a.dists[3 * i], d1 = a.dists[3 * i + 1], d2 = a.dists[3 * i + 2];
        double c[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const long long id = a.ids[3 * i + j];
            const double x = a.field_kind == 0 ? (double)__ldg((const float *)a.field + id)
                                               : (double)__ldg((const uint8_t *)a.field + id);
            double v = __dadd_rn(__dmul_rn(__ddiv_rn(__dadd_rn(x, -a.x_min), a.x_range), a.new_range), a.lower);
            v = __dadd_rn(v, a.add);
            if (a.quantize) v = (double)(uint16_t)(long long)v;
            c[j] = v;
        }
        const double sd = __dadd_rn(__dadd_rn(d0, d1), d2);
        const double w0 = __ddiv_rn(1.0, __ddiv_rn(__dadd_rn(d0, 0.00001), sd)), w1 = __ddiv_rn(1.0, __ddiv_rn(__dadd_rn(d1, 0.00001), sd)),
                     w2 = __ddiv_rn(1.0, __ddiv_rn(__dadd_rn(d2, 0.00001), sd));
        const double t = __dadd_rn(__dadd_rn(w0, w1), w2);
        const double v = __dadd_rn(__dadd_rn(__dmul_rn(c[0], __ddiv_rn(w0, t)), __dmul_rn(c[1], __ddiv_rn(w1, t))),
                                   __dmul_rn(c[2], __ddiv_rn(w2, t)));
        const int32_t px = (int32_t)v;
        if (a.out_bits == 8) ((uint8_t *)a.out)[i] = (uint8_t)px; else ((uint16_t *)a.out)[i] = (uint16_t)px;
    }
}

NXB_API int nxb_idw_map(const double *dists, const int64_t *ids, const void *field, int field_kind, int64_t n,
                        double x_min, double x_max, double lower, double upper, double add, int quantize_u16,
                        int out_bits, void *out, void *stream)
{
    NXB_ARG(n >= 0 && (field_kind == 0 || field_kind == 1) && (out_bits == 8 || out_bits == 16));
    if (n == 0) return NXB_OK;
    NXB_ARG(dists && ids && field && out);
    IdwMapArgs a;
    a.dists = dists; a.ids = (const long long *)ids; a.field = field; a.field_kind = field_kind;
    a.x_min = x_min; a.x_range = x_max - x_min; a.new_range = upper - lower; a.lower = lower; a.add = add;
    a.quantize = quantize_u16; a.out_bits = out_bits; a.out = out; a.n = n;
    idw_map_kernel<<<nxb_grid_resident(idw_map_kernel, 256, 0, (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(a);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}

NXB_API int nxb_idw_gray_f64(const double *dists, const int64_t *ids, const double *colors, int64_t n, int32_t *out, void *stream)
{
    NXB_ARG(n >= 0);
    if (n == 0) return NXB_OK;
    NXB_ARG(dists && ids && colors && out);
    idw_gray_kernel<<<nxb_grid_resident(idw_gray_kernel, 256, 0, (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        dists, (const long long *)ids, colors, n, out);
    NXB_LAUNCH_CHECK();
    return NXB_OK;
}
